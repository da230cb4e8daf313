"""Development timing of the engine vs the unmodified reference on one GPU (not the judged bench)."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from harness import P  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000.0)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "primary"
    ps = H.params_primary() if which == "primary" else H.params_secondary()
    parms = pf.EncryptionParameters(pf.scheme_type.ckks)
    parms.set_poly_modulus_degree(ps.n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(ps.size_P)
    parms.set_galois_elts(pf.get_elts_from_steps([1], ps.n))
    ctx = pf.PhantomContext(parms)
    l, n = ps.limbs(), ps.n
    a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
    rlk = pf.PhantomRelinKey(ctx, list(H.switch_key(ps, 100)))
    glk = pf.PhantomGaloisKey(ctx, [list(H.switch_key(ps, 1000))])
    ca, cb = pf.PhantomCiphertext.from_host(ctx, a), pf.PhantomCiphertext.from_host(ctx, b)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    x = torch.zeros((ps.size_QP, n), dtype=torch.int64, device="cuda")
    med, best = timeit(lambda: pf.nwt_2d_radix8_forward_inplace(x, ctx, ps.size_QP, 0))
    print(f"engine fwd NTT x{ps.size_QP}: median {med:.1f} us  best {best:.1f} us  -> {ps.size_QP / med:.3f} M NTT/s, "
          f"{ps.size_QP * 16 * n / med / 1e3:.1f} GB/s algorithmic")
    med, best = timeit(lambda: pf.nwt_2d_radix8_backward_inplace(x, ctx, ps.size_QP, 0))
    print(f"engine inv NTT x{ps.size_QP}: median {med:.1f} us  best {best:.1f} us  -> {ps.size_QP / med:.3f} M NTT/s")
    big = torch.zeros((100, n), dtype=torch.int64, device="cuda")
    for cnt in (1, 4, 100):
        med, best = timeit(lambda: pf.check(pf.lib.pfhe_ntt_forward_inplace(ctx._h, big.data_ptr(), min(cnt, ps.size_QP), 0, st)))
        print(f"engine fwd NTT x{min(cnt, ps.size_QP)}: median {med:.1f} us best {best:.1f}")

    work = ca.data.clone()

    def hmult():
        pf.check(pf.lib.pfhe_multiply_and_relin(ctx._h, 1, ca.data.data_ptr(), cb.data.data_ptr(), work.data_ptr(),
                                                rlk.public_keys_ptr(), st))

    def copy_only():
        work.copy_(ca.data)

    m_all, _ = timeit(hmult)
    m_copy, _ = timeit(copy_only)
    print(f"engine HMult+Relin: median {m_all:.1f} us -> {1e6 / m_all:.0f} ops/s (D2D copy of one ciphertext: {m_copy:.1f} us)")

    def rot():
        work.copy_(ca.data)
        pf.check(pf.lib.pfhe_rotate_inplace(ctx._h, 1, work.data_ptr(), 1, glk.get_relin_keys(0).public_keys_ptr(), st))

    m_all, _ = timeit(rot)
    print(f"engine rotate: median {m_all - m_copy:.1f} us")
    out = torch.empty((2, l - 1, n), dtype=torch.int64, device="cuda")
    med, _ = timeit(lambda: pf.check(pf.lib.pfhe_rescale_to_next(ctx._h, 1, ca.data.data_ptr(), 2, out.data_ptr(), st)))
    print(f"engine rescale: median {med:.1f} us")

    # stage timings
    up = torch.empty((ps.beta(), l + ps.size_P, n), dtype=torch.int64, device="cuda")
    cx = torch.empty((2, l + ps.size_P, n), dtype=torch.int64, device="cuda")
    d3 = torch.empty((3, l, n), dtype=torch.int64, device="cuda")
    med, _ = timeit(lambda: pf.check(pf.lib.pfhe_tensor_prod_2x2(ctx._h, ca.data.data_ptr(), cb.data.data_ptr(), d3.data_ptr(), l, st)))
    print(f"  tensor: {med:.1f} us")
    med, _ = timeit(lambda: pf.check(pf.lib.pfhe_modup(ctx._h, 1, up.data_ptr(), d3[2].data_ptr(), st)))
    print(f"  modup: {med:.1f} us")
    med, _ = timeit(lambda: pf.check(pf.lib.pfhe_key_switch_inner_prod(ctx._h, 1, cx.data_ptr(), up.data_ptr(), rlk.public_keys_ptr(), st)))
    print(f"  inner product: {med:.1f} us")
    ct = torch.empty((l, n), dtype=torch.int64, device="cuda")
    med, _ = timeit(lambda: pf.check(pf.lib.pfhe_moddown_from_ntt(ctx._h, 1, ct.data_ptr(), cx[0].data_ptr(), st)))
    print(f"  moddown (1 poly): {med:.1f} us")

    r = H.reference()
    if r is not None:
        steps = (ctypes.c_int * 1)(1)
        h = r.ref_create(3, ps.n, P(ps.primes), ps.size_QP, ps.size_P, 0, 0, steps, 1, float(2 ** 40), 1)
        assert h, r.ref_last_error()
        trials = 50
        times = (ctypes.c_double * trials)()
        for op, name, aux in ((3, f"fwd NTT x{ps.size_QP}", ps.size_QP), (4, f"inv NTT x{ps.size_QP}", ps.size_QP),
                              (3, "fwd NTT x1", 1), (3, "fwd NTT x4", 4), (4, "inv NTT x1", 1), (0, "HMult+Relin", 0), (1, "rotate", 1), (2, "rescale", 0)):
            assert r.ref_time_op(h, op, 1, P(a), P(b), aux, 0, trials, times) == 0, r.ref_last_error()
            ts = sorted(times)
            print(f"reference {name}: median {ts[trials // 2]:.1f} us best {ts[0]:.1f} us")
        assert r.ref_time_op(h, 0, 1, P(a), P(b), 0, 1, 20, times) == 0
        ts = sorted(times[:20])
        print(f"reference HMult+Relin e2e (H2D+op+D2H): median {ts[10]:.1f} us")
        r.ref_destroy(h)


if __name__ == "__main__":
    main()
