"""A known defect of the reference that a drop-in has to reproduce: scaleAndRound_HPS_QR_R_kernel (src/rns.cu:1699-1733)
reduces its floating-point correction `alpha` under r_0, then reduces THAT value under r_1, r_2, ... instead of reducing
the original under every r_j.  The base R is generated downwards from min(q_i) (rns.cu:687-694), r_0 > r_1 > ..., so whenever
alpha mod r_0 lands in [r_j, r_0) the later limbs of the scaled product describe a different integer than the earlier ones and
one coefficient of the size-3 product is off by a fixed fraction of Q (about 0.43 Q for the primes below; it decrypts as
+-28268 times 1, s or s^2).  The chance is (r_0 - r_j) / r_0 per coefficient: ~1e-6 with 40-bit primes -- one N = 8192 product
in ~35 decrypts wrongly -- and ~1e-12 with the 60-bit primes of the reference's benchmarks.

tests/golden/bfv_hps_alpha_case.json holds the seeds of such a product (found on the CPU by tools/dbg_cpu_bfv.py).  The
oracle, the engine and -- when it was built -- the unmodified reference have to give the same words, wrong coefficient
included."""
import ctypes
import json
import os

import numpy as np
import pytest

import harness as H
from harness import P

CASE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bfv_hps_alpha_case.json")


def build_case():
    case = json.load(open(CASE))
    ps = H.ParamSet("bfv_alpha_case", case["n"], case["prime_bits"], 1, scheme=2, t=case["t"])
    o, oc = H.oracle(), ps.octx()
    n, l, m = ps.n, ps.size_Q, ps.size_QP
    sd = [bytes.fromhex(v) for v in case["seeds"]]
    sk = np.zeros((m, n), dtype=np.uint64)
    o.orc_gen_secretkey(oc, sd[0], P(sk))
    pk = np.zeros((2, m, n), dtype=np.uint64)
    assert o.orc_encrypt_zero_symmetric(oc, 0, P(sk), sd[1], sd[2], P(pk)) == 0
    kc = o.orc_create(ps.scheme, ps.n, P(ps.primes), m, 0, ps.t)
    sk2 = np.zeros_like(sk)
    o.orc_poly_mul(kc, P(sk), P(sk), P(sk2), m)
    o.orc_destroy(kc)
    rlk = np.zeros((l // ps.size_P, 2, m, n), dtype=np.uint64)
    assert o.orc_gen_kswitch_key(oc, P(sk2), P(sk), bytes.fromhex(case["kswitch_seeds"]), P(rlk)) == 0
    a = np.zeros(n, dtype=np.uint64)
    a[0], a[1] = 3, 5
    b = np.zeros(n, dtype=np.uint64)
    b[0], b[n - 1] = 7, 2
    ca, cb = np.zeros((2, l, n), dtype=np.uint64), np.zeros((2, l, n), dtype=np.uint64)
    assert o.orc_encrypt_zero_symmetric(oc, 1, P(sk), sd[3], sd[4], P(ca)) == 0
    assert o.orc_encrypt_add_plain(oc, l, P(ca), P(a)) == 0
    assert o.orc_encrypt_zero_asymmetric(oc, P(pk), sd[5], sd[6], P(cb)) == 0
    assert o.orc_encrypt_add_plain(oc, l, P(cb), P(b)) == 0
    want = np.zeros(n, dtype=np.uint64)
    want[0], want[1], want[n - 1] = 11, 35, 6   # (3 + 5x)(7 + 2x^(n-1)) mod x^n + 1
    return case, ps, sk, rlk, ca, cb, want


def test_oracle_reproduces_the_wrong_coefficient():
    case, ps, sk, rlk, ca, cb, want = build_case()
    o, oc = H.oracle(), ps.octx()
    n, l = ps.n, ps.size_Q
    for x in (ca, cb):   # both inputs are sound
        dec = np.zeros(n, dtype=np.uint64)
        assert o.orc_decrypt(oc, l, P(x), 2, P(sk), 2, 1, P(dec)) == 0
        assert np.count_nonzero(dec % ps.t) == 2
    prod = np.zeros((2, l, n), dtype=np.uint64)
    assert o.orc_bfv_multiply_relin_hps(oc, P(ca), P(cb), P(rlk), P(prod)) == 0
    dec = np.zeros(n, dtype=np.uint64)
    assert o.orc_decrypt(oc, l, P(prod), 2, P(sk), 2, 1, P(dec)) == 0
    bad = np.nonzero(dec % ps.t != want)[0]
    assert list(bad) == [case["wrong_coefficient"]] and int(dec[bad[0]] % ps.t) == case["decrypts_to"]
    # the cause: with the unreduced alpha reduced under every r_j the same product decrypts correctly
    o.orc_set_hps_alpha_per_limb(1)
    try:
        assert o.orc_bfv_multiply_relin_hps(oc, P(ca), P(cb), P(rlk), P(prod)) == 0
        assert o.orc_decrypt(oc, l, P(prod), 2, P(sk), 2, 1, P(dec)) == 0
        assert np.array_equal(dec % ps.t, want)
    finally:
        o.orc_set_hps_alpha_per_limb(0)
    # the BEHZ product of the same ciphertexts (no floating-point correction, no base R) decrypts to the product
    assert o.orc_bfv_multiply_relin_behz(oc, P(ca), P(cb), P(rlk), P(prod)) == 0
    assert o.orc_decrypt(oc, l, P(prod), 2, P(sk), 1, 1, P(dec)) == 0
    assert np.array_equal(dec % ps.t, want)


@pytest.mark.gpu
def test_engine_and_reference_reproduce_the_wrong_coefficient():
    import torch
    import phantom_fhe_b200 as pf
    case, ps, sk, rlk, ca, cb, want = build_case()
    o, oc = H.oracle(), ps.octx()
    n, l = ps.n, ps.size_Q
    want3 = np.zeros((3, l, n), dtype=np.uint64)
    assert o.orc_bfv_multiply_hps(oc, P(ca), P(cb), P(want3)) == 0
    want2 = np.zeros((2, l, n), dtype=np.uint64)
    assert o.orc_bfv_multiply_relin_hps(oc, P(ca), P(cb), P(rlk), P(want2)) == 0
    parms = pf.EncryptionParameters(pf.scheme_type.bfv)
    parms.set_poly_modulus_degree(n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(ps.size_P)
    parms.set_plain_modulus(ps.t)
    parms.set_mul_tech(pf.mul_tech_type.hps)
    ctx = pf.PhantomContext(parms)
    key = pf.PhantomRelinKey(ctx, list(rlk))
    x = pf.PhantomCiphertext.from_host(ctx, ca, is_ntt_form=False)
    y = pf.PhantomCiphertext.from_host(ctx, cb, is_ntt_form=False)
    pf.multiply_inplace(ctx, x, y)
    assert np.array_equal(x.to_host(), want3), "size-3 HPS product"
    x = pf.PhantomCiphertext.from_host(ctx, ca, is_ntt_form=False)
    pf.multiply_and_relin_inplace(ctx, x, y, key)
    assert np.array_equal(x.to_host(), want2), "HPS product, relinearised"
    secret = pf.PhantomSecretKey(ctx, sk)
    torch.cuda.synchronize()
    dec = secret.decrypt(ctx, x).cpu().numpy().view(np.uint64) % ps.t
    bad = np.nonzero(dec != want)[0]
    assert list(bad) == [case["wrong_coefficient"]] and int(dec[bad[0]]) == case["decrypts_to"]
    r = H.reference()
    if r is None:
        return
    h = r.ref_create(2, n, P(ps.primes), ps.size_QP, ps.size_P, ps.t, 2, None, 0, 1.0, 1)
    assert h, r.ref_last_error()
    try:
        for d in range(rlk.shape[0]):
            assert r.ref_key_set(h, -1, d, P(rlk[d])) == 0, r.ref_last_error()
        got = np.zeros((2, l, n), dtype=np.uint64)
        assert r.ref_multiply_relin(h, 1, P(ca), P(cb), P(got)) == 0, r.ref_last_error()
        assert np.array_equal(got, want2), "the unmodified reference gives the same words"
    finally:
        r.ref_destroy(h)
