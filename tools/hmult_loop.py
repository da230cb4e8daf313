"""Back-to-back HMult+Relin loop (the bench's device-timed region) for A/B runs of engine switches (env)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
parms.set_galois_elts(pf.get_elts_from_steps([1], ps.n))
ctx = pf.PhantomContext(parms)
a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
rlk = pf.PhantomRelinKey(ctx, list(H.switch_key(ps, 100)))
glk = pf.PhantomGaloisKey(ctx, [list(H.switch_key(ps, 1000))])
ca, cb = pf.PhantomCiphertext.from_host(ctx, a), pf.PhantomCiphertext.from_host(ctx, b)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
work = ca.data.clone()


def hmult():
    pf.check(pf.lib.pfhe_multiply_and_relin(ctx._h, 1, ca.data.data_ptr(), cb.data.data_ptr(), work.data_ptr(),
                                            rlk.public_keys_ptr(), st))


def rot():
    pf.check(pf.lib.pfhe_rotate_inplace(ctx._h, 1, work.data_ptr(), 1, glk.get_relin_keys(0).public_keys_ptr(), st))


for name, fn in (("HMult+Relin", hmult), ("rotate", rot)):
    best = 1e9
    for trial in range(3):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000 / reps)
    print(f"{name}: {best:.1f} us/op  ({1e6 / best:.0f} ops/s)  env OVERLAP={os.environ.get('PFHE_OVERLAP', '1')}", flush=True)
