"""Back-to-back 64-limb forward NTT launches (the bench's roofline kernel pair) for A/B runs of engine switches."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
ctx = pf.PhantomContext(parms)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
bufs = [torch.zeros((4, 16, ps.n), dtype=torch.int64, device="cuda") for _ in range(6)]   # 6 x 32 MiB > L2
for name, fn in (("fwd", pf.lib.pfhe_ntt_forward_inplace_batch), ("inv", pf.lib.pfhe_ntt_backward_inplace_batch)):
    best = 1e9
    for trial in range(3):
        for i in range(12):
            pf.check(fn(ctx._h, bufs[i % 6].data_ptr(), 4, 16, 0, st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            pf.check(fn(ctx._h, bufs[i % 6].data_ptr(), 4, 16, 0, st))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000 / reps)
    print(f"{name} NTT x64: {best:.2f} us  ({64 / best:.3f} M limb-NTT/s)  STAGGER={os.environ.get('PFHE_STAGGER_NS', '0')}", flush=True)
