"""BFV HMult+Relin (BEHZ, HPS, HPS over Q) at the bfv_bench.cu N=2^14 parameter sets: engine vs unmodified reference, device timed."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402
from harness import P  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
r = H.reference()
for which, tech in [(w, m) for m in (1, 2, 3) for w in (0, 1, 2)]:
    ps = H.params_bfv_bench(which)
    parms = pf.EncryptionParameters(pf.scheme_type.bfv)
    parms.set_poly_modulus_degree(ps.n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(ps.size_P)
    parms.set_plain_modulus(ps.t)
    parms.set_mul_tech(pf.mul_tech_type(tech))
    ctx = pf.PhantomContext(parms)
    a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
    rlk = pf.PhantomRelinKey(ctx, list(H.switch_key(ps, 100)))
    ca = pf.PhantomCiphertext.from_host(ctx, a, is_ntt_form=False)
    cb = pf.PhantomCiphertext.from_host(ctx, b, is_ntt_form=False)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = ca.data.clone()

    def op():
        pf.check(pf.lib.pfhe_multiply_and_relin(ctx._h, 1, ca.data.data_ptr(), cb.data.data_ptr(), out.data_ptr(),
                                                rlk.public_keys_ptr(), st))
    for _ in range(10):
        op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op()
    e1.record()
    torch.cuda.synchronize()
    eng_us = e0.elapsed_time(e1) * 1000 / reps
    ref_us = float("nan")
    if r is not None:
        h = r.ref_create(2, ps.n, P(ps.primes), ps.size_QP, ps.size_P, ps.t, tech, None, 0, 1.0, 1)
        times = np.zeros(60, dtype=np.float64)
        assert r.ref_time_op(h, 0, 1, P(a), P(b), 0, 0, 60, times.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
        ref_us = float(np.median(times[10:]))
        r.ref_destroy(h)
    # throughput of independent pairs through the batched entry point (lanes)
    lane_txt = ""
    if tech == 2:
        cnt = 64
        outs = [ca.data.clone() for _ in range(4)]
        Arr = ctypes.c_void_p * cnt
        aa, bb = Arr(*[ca.data.data_ptr()] * cnt), Arr(*[cb.data.data_ptr()] * cnt)
        oo = Arr(*[outs[i % 4].data_ptr() for i in range(cnt)])
        for lanes in (2, 4):
            pf.check(pf.lib.pfhe_engine_set_lanes(ctx._h, lanes))
            pf.check(pf.lib.pfhe_multiply_and_relin_batch(ctx._h, 1, aa, bb, oo, 8, rlk.public_keys_ptr(), st))
            torch.cuda.synchronize()
            e0.record()
            pf.check(pf.lib.pfhe_multiply_and_relin_batch(ctx._h, 1, aa, bb, oo, cnt, rlk.public_keys_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            lane_txt += f"  {lanes} lanes {e0.elapsed_time(e1) * 1000 / cnt:.1f} us/op"
    print(f"bfv set {which} {pf.mul_tech_type(tech).name}:{lane_txt}  ", end="")
    print(f" l={ps.limbs()} alpha={ps.size_P} engine {eng_us:.1f} us  reference {ref_us:.1f} us  "
          f"speedup {ref_us / eng_us:.2f}x", flush=True)
