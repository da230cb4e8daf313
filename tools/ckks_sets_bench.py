"""HMult+Relin and rotate at the six N = 2^16 parameter sets of benchmark/ckks_bench.cu:320-394 ({60, 40 x k, 60 x P}, P = 1..6,
36..43 primes): engine (one op at a time, device resident) vs the unmodified reference on the same GPU."""
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

SETS = [(41, 1), (39, 2), (38, 3), (35, 4), (34, 5), (29, 6)]   # (number of 40-bit primes, special primes)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
r = H.reference()
dev = torch.device("cuda")
for k40, size_P in SETS:
    bits = [60] + [40] * k40 + [60] * size_P
    ps = H.ParamSet(f"ckks16_P{size_P}", 65536, bits, size_P)
    parms = pf.EncryptionParameters(pf.scheme_type.ckks)
    parms.set_poly_modulus_degree(ps.n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(size_P)
    parms.set_galois_elts(pf.get_elts_from_steps([1], ps.n))
    ctx = pf.PhantomContext(parms)
    l, n = ps.limbs(), ps.n
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)

    def rand(shape, moduli):
        t = torch.empty(shape, dtype=torch.int64, device=dev)
        flat = t.view(-1, len(moduli), n)
        for j, q in enumerate(moduli):
            flat[:, j, :] = torch.randint(0, int(q), (flat.shape[0], n), generator=gen, device=dev, dtype=torch.int64)
        return t

    primes = [int(p) for p in ps.primes]
    a, b = rand((2, l, n), primes[:l]), rand((2, l, n), primes[:l])
    out = torch.empty_like(a)
    dnum = ctx.dnum(1)
    rlk = pf.PhantomRelinKey.from_device(ctx, [rand((2, ps.size_QP, n), primes) for _ in range(dnum)])
    glk = pf.PhantomRelinKey.from_device(ctx, [rand((2, ps.size_QP, n), primes) for _ in range(dnum)])
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    elt = pf.get_elt_from_step(1, n)

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    mul_us = timed(lambda: pf.check(pf.lib.pfhe_multiply_and_relin(ctx._h, 1, a.data_ptr(), b.data_ptr(), out.data_ptr(),
                                                                   rlk.public_keys_ptr(), st)))
    rot = a.clone()
    rot_us = timed(lambda: pf.check(pf.lib.pfhe_apply_galois_inplace(ctx._h, 1, rot.data_ptr(), elt, glk.public_keys_ptr(), st)))
    ref_mul = ref_rot = float("nan")
    if r is not None:
        ah = a.cpu().numpy().view(np.uint64)
        bh = b.cpu().numpy().view(np.uint64)
        arr = (ctypes.c_int * 1)(1)
        h = r.ref_create(3, n, P(ps.primes), ps.size_QP, size_P, 0, 0, arr, 1, float(2 ** 40), 1)
        times = (ctypes.c_double * 40)()
        assert r.ref_time_op(h, 0, 1, P(ah), P(bh), 0, 0, 40, times) == 0
        ref_mul = sorted(times[8:])[16]
        assert r.ref_time_op(h, 1, 1, P(ah), P(ah), 1, 0, 40, times) == 0
        ref_rot = sorted(times[8:])[16]
        r.ref_destroy(h)
    print(f"P={size_P} l={l} beta={ps.beta()}: HMult+Relin engine {mul_us:.1f} us, reference {ref_mul:.1f} us ({ref_mul / mul_us:.2f}x); "
          f"rotate engine {rot_us:.1f} us, reference {ref_rot:.1f} us ({ref_rot / rot_us:.2f}x)", flush=True)
    del ctx, rlk, glk, a, b, out, rot
    torch.cuda.empty_cache()
