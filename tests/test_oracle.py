"""CPU tests of the oracle (test infrastructure): golden fixtures generated from the reference's own host code
(tests/golden/host_tables.json, make_golden.py), the survey's known-answer anchor, algebraic properties, and --
when oracle/_ref was built -- a live comparison with the reference host library."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

import harness as H
from harness import P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "host_tables.json")))


def test_prime_chains_match_reference_create():
    o = H.oracle()
    for name, cfg in GOLD["configs"].items():
        bits = (ctypes.c_int * len(cfg["bits"]))(*cfg["bits"])
        out = np.zeros(len(cfg["bits"]), dtype=np.uint64)
        assert o.orc_create_primes(cfg["n"], bits, len(cfg["bits"]), P(out)) == 0
        assert [int(v) for v in out] == cfg["primes"], name


def test_ntt_tables_match_reference_host_ntt():
    o = H.oracle()
    for key, t in GOLD["tables"].items():
        logn, q = (int(x) for x in key.split(":"))
        n = 1 << logn
        pr = np.array([q], dtype=np.uint64)
        c = o.orc_create(3, n, P(pr), 1, 0, 0)
        assert c
        assert o.orc_minimal_primitive_root(2 * n, q) == t["root"]
        assert o.orc_n_inv(c, 0) == t["n_inv"]
        assert o.orc_shoup(t["n_inv"], q) == t["n_inv_shoup"]
        ratio = np.zeros(3, dtype=np.uint64)
        o.orc_barrett_ratio(q, P(ratio))
        assert [int(v) for v in ratio] == t["ratio"]
        for i, f in enumerate(("orc_twiddle", "orc_twiddle_shoup", "orc_itwiddle", "orc_itwiddle_shoup")):
            arr = np.ctypeslib.as_array(getattr(o, f)(c, 0), shape=(n,))
            assert [int(v) for v in arr[:8]] == t[["tw_head", "tws_head", "itw_head", "itws_head"][i]]
            assert hashlib.sha256(arr.tobytes()).hexdigest() == t["sha256"][i], (key, f)
        o.orc_destroy(c)


def test_galois_elements():
    o = H.oracle()
    for n, table in GOLD["galois"].items():
        for step, elt in table.items():
            assert o.orc_galois_elt_from_step(int(step), int(n)) == elt


def test_known_answer_config1():
    """x_j = mt19937_64(1)() % q at N=4096, q=1125899906826241."""
    o = H.oracle()
    ps = H.params_c1()
    x = np.zeros(ps.n, dtype=np.uint64)
    o.orc_mt19937_64_fill(1, int(ps.primes[0]), P(x), ps.n, 0)
    assert int(x[0]) == 2469588189546311528 % 1125899906826241  # first mt19937_64(1) output
    y = x.copy()
    idx = (ctypes.c_int * 1)(0)
    o.orc_ntt_forward(ps.octx(), P(y), 1, idx)
    assert [int(v) for v in y[:4]] == [213908721093404, 678455973401121, 1034267331304760, 457393895113370]
    o.orc_ntt_inverse(ps.octx(), P(y), 1, idx)
    assert np.array_equal(x, y)


def test_ntt_is_negacyclic_convolution():
    """NTT(a) * NTT(b) = NTT(a * b mod X^N + 1): pins the transform against schoolbook multiplication."""
    o = H.oracle()
    ps = H.ParamSet("tiny", 4096, [40], 0)
    q = int(ps.primes[0])
    rng = np.random.default_rng(5)
    n = ps.n
    a = np.zeros(n, dtype=np.uint64)
    b = np.zeros(n, dtype=np.uint64)
    ia, ib = rng.integers(0, n, 6), rng.integers(0, n, 5)
    a[ia] = rng.integers(1, q, 6).astype(np.uint64)
    b[ib] = rng.integers(1, q, 5).astype(np.uint64)
    want = [0] * n
    for i in np.nonzero(a)[0]:
        for j in np.nonzero(b)[0]:
            k, v = int(i + j), int(a[i]) * int(b[j])
            if k >= n:
                k, v = k - n, -v
            want[k] = (want[k] + v) % q
    fa, fb = a.copy(), b.copy()
    idx = (ctypes.c_int * 1)(0)
    o.orc_ntt_forward(ps.octx(), P(fa), 1, idx)
    o.orc_ntt_forward(ps.octx(), P(fb), 1, idx)
    prod = np.array([(int(x) * int(y)) % q for x, y in zip(fa, fb)], dtype=np.uint64)
    o.orc_ntt_inverse(ps.octx(), P(prod), 1, idx)
    assert [int(v) for v in prod] == want


def _crt_poly(ps, limbs):
    """centered CRT lift of [l][n] residues (python ints)"""
    primes = [int(p) for p in ps.primes[: limbs.shape[0]]]
    Q = 1
    for p in primes:
        Q *= p
    vals = [0] * limbs.shape[1]
    for i, p in enumerate(primes):
        Qi = Q // p
        c = (Qi * pow(Qi, -1, p)) % Q
        for x in range(limbs.shape[1]):
            vals[x] = (vals[x] + int(limbs[i, x]) * c) % Q
    return vals, Q


@pytest.mark.parametrize("log_dim", [8, 11])
def test_nwt_1d_matches_the_context_transform(log_dim):
    """orc_fnwt_1d / orc_inwt_1d (ntt_1d.cu:146-292) on explicit tables = the oracle's table-driven transform; the
    definition NTT(x)[k] = sum_j x_j psi^(j (2 bitrev(k) + 1)) is checked directly at a few outputs."""
    o = H.oracle()
    dim, count = 1 << log_dim, 3
    primes = np.zeros(count, dtype=np.uint64)
    assert o.orc_create_primes(dim, (ctypes.c_int * count)(50, 50, 50), count, P(primes)) == 0
    c = o.orc_create(3, dim, P(primes), count, 0, 0)
    get = lambda f, i: np.ctypeslib.as_array(getattr(o, f)(c, i), shape=(dim,)).copy()
    tw = np.stack([get("orc_twiddle", i) for i in range(count)])
    tws = np.stack([get("orc_twiddle_shoup", i) for i in range(count)])
    itw = np.stack([get("orc_itwiddle", i) for i in range(count)])
    itws = np.stack([get("orc_itwiddle_shoup", i) for i in range(count)])
    ninv = np.array([o.orc_n_inv(c, i) for i in range(count)], dtype=np.uint64)
    ninvs = np.array([o.orc_shoup(int(ninv[i]), int(primes[i])) for i in range(count)], dtype=np.uint64)
    rng = np.random.default_rng(log_dim)
    x = np.stack([rng.integers(0, int(primes[i]), dim, dtype=np.uint64) for i in range(count)])
    a, b = x.copy(), x.copy()
    o.orc_fnwt_1d(P(a), P(tw), P(tws), P(primes), dim, count - 1, 1)
    o.orc_ntt_forward(c, P(b[1:].copy()), 0, (ctypes.c_int * 2)(1, 2))
    bb = x[1:].copy()
    o.orc_ntt_forward(c, P(bb), 2, (ctypes.c_int * 2)(1, 2))
    assert np.array_equal(a[0], x[0]) and np.array_equal(a[1:], bb)
    q = int(primes[1])
    psi = int(tw[1][1 << (log_dim - 1)]) if False else int(o.orc_minimal_primitive_root(2 * dim, q))
    rev = lambda v: int(format(v, f"0{log_dim}b")[::-1], 2)
    for k in (0, 1, dim - 1):
        e = 2 * rev(k) + 1
        want = sum(int(x[1][j]) * pow(psi, j * e, q) for j in range(dim)) % q
        assert int(a[1][k]) == want
    o.orc_inwt_1d(P(a), P(itw), P(itws), P(primes), P(ninv), P(ninvs), dim, count - 1, 1)
    assert np.array_equal(a, x)
    o.orc_destroy(c)


@pytest.mark.parametrize("sizes", [(2, 2), (3, 2), (2, 4), (5, 5)])
def test_tensor_mxn_is_polynomial_product_in_the_key(sizes):
    """orc_tensor_mxn (polymath.cu:546-594): out[j] = sum_{i1+i2=j} a[i1] b[i2] mod q, against Python integers."""
    sa, sb = sizes
    ps = H.params_small(4096, l=2, alpha=1)
    o = H.oracle()
    l, n = ps.limbs(), ps.n
    rng = np.random.default_rng(7 * sa + sb)
    a = np.stack([np.stack([rng.integers(0, int(ps.primes[i]), n, dtype=np.uint64) for i in range(l)]) for _ in range(sa)])
    b = np.stack([np.stack([rng.integers(0, int(ps.primes[i]), n, dtype=np.uint64) for i in range(l)]) for _ in range(sb)])
    a[:, :, 0] = [[int(ps.primes[i]) - 1 for i in range(l)]] * sa   # extreme residues: largest 128-bit sums
    b[:, :, 0] = [[int(ps.primes[i]) - 1 for i in range(l)]] * sb
    out = np.zeros((sa + sb - 1, l, n), dtype=np.uint64)
    o.orc_tensor_mxn(ps.octx(), P(a), sa, P(b), sb, P(out), l)
    for i in range(l):
        q = int(ps.primes[i])
        for x in (0, 1, 17, n - 1):
            for j in range(sa + sb - 1):
                want = sum(int(a[u, i, x]) * int(b[j - u, i, x]) for u in range(sa) if 0 <= j - u < sb) % q
                assert int(out[j, i, x]) == want
    # in place on the first operand (the reference writes into encrypted1's resized buffer)
    buf = np.zeros((sa + sb - 1, l, n), dtype=np.uint64)
    buf[:sa] = a
    o.orc_tensor_mxn(ps.octx(), P(buf), sa, P(b), sb, P(buf), l)
    assert np.array_equal(buf, out)


def test_keyswitch_is_hybrid_keyswitch():
    """Semantic pin of modup/inner-product/moddown: with evk_d = (P * qhat_d * qhat_d^-1-ish gadget) * s' the output of
    the path must equal c2 * s' up to the rounding error of mod-down -- checked through a noise-free gadget key:
    evk_d[0] = P * g_d * s2 (no mask), evk_d[1] = 0, so keyswitch(c2) = (round-ish(c2 * s2), 0)."""
    o = H.oracle()
    ps = H.params_small(4096, l=4, alpha=2)
    oc, n, l = ps.octx(), ps.n, ps.limbs()
    m, beta = l + ps.size_P, ps.beta()
    primes = [int(p) for p in ps.primes]
    Qs, Ps = primes[:l], primes[ps.size_Q:]
    Pbig = 1
    for p in Ps:
        Pbig *= p
    rng = np.random.default_rng(3)
    # s2: small ternary polynomial in NTT form over all QP limbs
    s_coeff = rng.integers(-1, 2, n)
    s_ntt = np.zeros((ps.size_QP, n), dtype=np.uint64)
    for r, p in enumerate(primes):
        s_ntt[r] = np.array([(int(v) % p) for v in s_coeff], dtype=np.uint64)
    idx_all = (ctypes.c_int * ps.size_QP)(*range(ps.size_QP))
    o.orc_ntt_forward(oc, P(s_ntt), ps.size_QP, idx_all)
    # gadget: g_d = Qhat_d * (Qhat_d^-1 mod Q_d) where Q_d = product of digit d's primes
    Q = 1
    for p in Qs:
        Q *= p
    evk = np.zeros((beta, 2, ps.size_QP, n), dtype=np.uint64)
    for d in range(beta):
        dig = Qs[d * ps.size_P:(d + 1) * ps.size_P]
        Qd = 1
        for p in dig:
            Qd *= p
        Qhat = Q // Qd
        g = Pbig * Qhat * pow(Qhat, -1, Qd)
        for r, p in enumerate(primes):
            gm = g % p
            evk[d, 0, r] = np.array([(int(v) * gm) % p for v in s_ntt[r]], dtype=np.uint64)
    c2 = H.uniform_limbs(ps, list(range(l)), 77)[0]
    ct = np.zeros((2, l, n), dtype=np.uint64)
    o.orc_keyswitch(oc, l, P(ct), P(c2), P(evk))
    assert not ct[1].any()
    # expected: c2 * s2 mod Q (negacyclic), compare in coefficient domain with tolerance from mod-down rounding
    idx = (ctypes.c_int * l)(*range(l))
    got = ct[0].copy()
    o.orc_ntt_inverse(oc, P(got), l, idx)
    prod = np.zeros((l, n), dtype=np.uint64)
    for r in range(l):
        p = primes[r]
        prod[r] = np.array([(int(x) * int(y)) % p for x, y in zip(c2[r], s_ntt[r])], dtype=np.uint64)
    o.orc_ntt_inverse(oc, P(prod), l, idx)
    gv, Qm = _crt_poly(ps, got[:, :64])
    pv, _ = _crt_poly(ps, prod[:, :64])
    for x in range(64):
        diff = (gv[x] - pv[x]) % Qm
        diff = min(diff, Qm - diff)
        assert diff <= beta * ps.size_P * n, diff  # fast-base-conversion overflow + floor error, tiny vs Q


def test_rescale_divides_by_last_prime():
    o = H.oracle()
    ps = H.params_small(4096, l=3, alpha=1)
    oc, n, l = ps.octx(), ps.n, ps.limbs()
    ct = H.ciphertext(ps, 4, polys=1)
    out = np.zeros((1, l - 1, n), dtype=np.uint64)
    src = ct.copy()
    o.orc_rescale(oc, l, P(src), 1, P(out))
    idx = (ctypes.c_int * l)(*range(l))
    coeff = ct[0].copy()
    o.orc_ntt_inverse(oc, P(coeff), l, idx)
    res = out[0].copy()
    o.orc_ntt_inverse(oc, P(res), l - 1, idx)
    v, Q = _crt_poly(ps, coeff[:, :32])
    w, Q2 = _crt_poly(ps, res[:, :32])
    ql = int(ps.primes[l - 1])
    for x in range(32):
        assert (v[x] - (v[x] % ql)) // ql % Q2 == w[x]  # floor division by q_last (rns.cu:1141-1158)


@pytest.mark.skipif(H.reference() is None, reason="oracle/_ref not built")
def test_live_against_reference_host_code():
    r, o = H.reference(), H.oracle()
    bits = [60] + [40] * 15 + [60] * 4
    arr = (ctypes.c_int * len(bits))(*bits)
    a, b = np.zeros(20, dtype=np.uint64), np.zeros(20, dtype=np.uint64)
    assert r.ref_host_create_primes(65536, arr, 20, P(a)) == 0
    assert o.orc_create_primes(65536, arr, 20, P(b)) == 0
    assert np.array_equal(a, b)
    for case in GOLD["bconv"]:
        ib, ob = case["ibase"], case["obase"]
        for j, p in enumerate(ob):
            for i in range(len(ib)):
                want = 1
                for k, qk in enumerate(ib):
                    if k != i:
                        want = want * qk % p
                assert case["qhat_mod_p"][j * len(ib) + i] == want


@pytest.mark.parametrize("tech", ["behz", "hps", "hps_overq", "hps_overq_drop1"])
def test_bfv_multiply_decrypts_to_the_plaintext_product(tech):
    """BEHZ / HPS restatements (oracle/fhe_oracle.c; evaluate.cu:451-548,647-801): noiseless-key sanity. Trivial
    encryptions (Delta*m + e, 0) must multiply to a ciphertext whose c0 decodes to m1*m2 mod (X^n+1, t) and whose
    c1, c2 are 0."""
    o = H.oracle()
    mul = o.orc_bfv_multiply_behz if tech == "behz" else o.orc_bfv_multiply_hps
    if tech.startswith("hps_overq"):   # evaluate.cu:647-801 with mul_tech hps_overq / the leveled arithmetic
        drop = 1 if tech.endswith("drop1") else 0
        mul = lambda c, a, b, out: o.orc_bfv_multiply_hps_overq(c, a, b, out, drop)
    t = 65537
    ps = H.ParamSet("bfv_sem", 64, [40, 40, 40, 50], 1, scheme=2, t=t)
    n, lq = ps.n, ps.size_Q
    Q = 1
    for p in ps.primes[:lq]:
        Q *= int(p)
    delta = Q // t
    rng = np.random.default_rng(1)
    m1, m2 = rng.integers(0, t, n), rng.integers(0, t, n)

    def enc(m):
        ct = np.zeros((2, lq, n), dtype=np.uint64)
        e = rng.integers(-5, 6, n)
        for i in range(lq):
            q = int(ps.primes[i])
            ct[0, i] = [(delta * int(v) + int(x)) % q for v, x in zip(m, e)]
        return ct

    c1, c2 = enc(m1), enc(m2)
    out = np.zeros((3, lq, n), dtype=np.uint64)
    assert mul(ps.octx(), H.P(c1), H.P(c2), H.P(out)) == 0
    exp = [0] * n
    for i in range(n):
        for j in range(n):
            v = int(m1[i]) * int(m2[j])
            if i + j >= n:
                exp[i + j - n] -= v
            else:
                exp[i + j] += v
    exp = [v % t for v in exp]
    dec = []
    for j in range(n):
        x = 0
        for i in range(lq):
            q = int(ps.primes[i])
            qh = Q // q
            x += int(out[0, i, j]) * pow(qh, -1, q) % q * qh
        x %= Q
        dec.append(((t * x + Q // 2) // Q) % t)
    assert dec == exp
    assert not out[1].any() and not out[2].any()
    if tech.startswith("hps_overq"):
        if tech.endswith("drop1"):   # ExpandCRTBasis_Ql_Q leaves the dropped limb at zero (rns.cu:1810-1822)
            assert not out[0, lq - 1].any() and out[0, :lq - 1].any()
        return
    if tech == "hps":   # R: size_Q + 1 NTT primes just below min(q_i) (rns.cu:687-694)
        R = np.zeros(72, dtype=np.uint64)
        nr = ctypes.c_int()
        assert o.orc_hps_aux(ps.octx(), H.P(R), ctypes.byref(nr)) == 0
        assert nr.value == lq + 1
        qmin = min(int(p) for p in ps.primes[:lq])
        assert all(int(v) < qmin and int(v) % (2 * n) == 1 for v in R[:nr.value])
        assert list(R[:nr.value]) == sorted(R[:nr.value], reverse=True)
        return
    # the auxiliary base: m_sk is the largest 61-bit NTT prime, B the next ones (rns.cu:414-420)
    bsk = np.zeros(66, dtype=np.uint64)
    nb = ctypes.c_int()
    assert o.orc_behz_aux(ps.octx(), H.P(bsk), ctypes.byref(nb)) == 0
    assert nb.value in (lq + 1, lq + 2)
    assert int(bsk[nb.value - 1]) == max(int(v) for v in bsk[:nb.value])
    assert all(int(v) % (2 * n) == 1 and int(v).bit_length() == 61 for v in bsk[:nb.value])


@pytest.mark.parametrize("scheme,mul_tech", [(3, 0), (1, 0), (2, 1), (2, 2)])
def test_decrypt_recovers_the_message(scheme, mul_tech):
    """orc_decrypt (secretkey.cu:533-691): symmetric encryptions (c0 = -(a s) + payload, c1 = a) built here with a ternary
    secret decrypt to the message -- CKKS: the NTT-form plaintext itself; BGV: m from m + t e; BFV: m from Delta m + e
    through the BEHZ and the HPS scale-and-round; a size-3 ciphertext uses s^2."""
    o = H.oracle()
    t = 65537 if scheme != 3 else 0
    ps = H.ParamSet("dec", 4096, [40, 40, 40, 50], 1, scheme=scheme, t=t)
    oc, n, l = ps.octx(), ps.n, ps.size_Q
    rng = np.random.default_rng(scheme * 10 + mul_tech)
    idx_all = (ctypes.c_int * ps.size_QP)(*range(ps.size_QP))
    idx = (ctypes.c_int * l)(*range(l))
    sec = rng.integers(-1, 2, n)
    s1 = np.stack([np.array([(int(v)) % int(p) for v in sec], dtype=np.uint64) for p in ps.primes])   # key level
    o.orc_ntt_forward(oc, P(s1), ps.size_QP, idx_all)
    s2 = np.stack([np.array([(int(x) * int(x)) % int(p) for x in s1[i]], dtype=np.uint64) for i, p in enumerate(ps.primes)])
    sk_pow = np.stack([s1, s2])
    Q = 1
    for p in ps.primes[:l]:
        Q *= int(p)
    m = rng.integers(0, t if t else 1 << 30, n)
    e = rng.integers(-4, 5, n)
    if scheme == 3:
        payload = [int(v) for v in m]
    elif scheme == 1:
        payload = [int(v) + t * int(x) for v, x in zip(m, e)]
    else:
        payload = [(Q // t) * int(v) + int(x) for v, x in zip(m, e)]
    pay = np.stack([np.array([v % int(p) for v in payload], dtype=np.uint64) for p in ps.primes[:l]])
    a = np.stack([rng.integers(0, int(p), n, dtype=np.uint64) for p in ps.primes[:l]])   # NTT form
    pay_ntt = pay.copy()
    o.orc_ntt_forward(oc, P(pay_ntt), l, idx)
    c0 = np.stack([np.array([(int(pv) - int(av) * int(sv)) % int(p) for pv, av, sv in zip(pay_ntt[i], a[i], s1[i])],
                            dtype=np.uint64) for i, p in enumerate(ps.primes[:l])])
    ct = np.stack([c0, a])
    if scheme == 2:   # BFV ciphertexts live in coefficient form
        for k in range(2):
            o.orc_ntt_inverse(oc, P(ct[k]), l, idx)
    out = np.zeros((l, n) if scheme == 3 else (n,), dtype=np.uint64)
    assert o.orc_decrypt(oc, l, P(ct), 2, P(sk_pow), mul_tech, 1, P(out)) == 0
    if scheme == 3:
        assert np.array_equal(out, pay_ntt)
    else:
        assert [int(v) for v in out] == [int(v) for v in m]
    # size 3: (c0 - b s^2, a, b) decrypts to the same message
    b = np.stack([rng.integers(0, int(p), n, dtype=np.uint64) for p in ps.primes[:l]])
    c0b = np.stack([np.array([(int(cv) - int(bv) * int(sv)) % int(p) for cv, bv, sv in zip(c0[i], b[i], s2[i])],
                             dtype=np.uint64) for i, p in enumerate(ps.primes[:l])])
    ct3 = np.stack([c0b, a, b])
    if scheme == 2:
        for k in range(3):
            o.orc_ntt_inverse(oc, P(ct3[k]), l, idx)
    out3 = np.zeros_like(out)
    assert o.orc_decrypt(oc, l, P(ct3), 3, P(sk_pow), mul_tech, 1, P(out3)) == 0
    assert np.array_equal(out3, out)
    if scheme == 1:   # BGV correction factor: the decryption is multiplied by its inverse mod t
        outc = np.zeros_like(out)
        assert o.orc_decrypt(oc, l, P(ct), 2, P(sk_pow), 0, 3, P(outc)) == 0
        assert [int(v) for v in outc] == [(int(v) * pow(3, -1, t)) % t for v in m]


def test_batch_encoder_is_slotwise():
    """orc_batch_encode / decode (src/batchencoder.cu): decode(encode(v)) = v, and the product of two plaintext polynomials
    mod (X^N + 1, t) decodes to the slot-wise product -- the property batching exists for."""
    o = H.oracle()
    n, t = 64, 65537 if False else 257   # 257 = 1 mod 128
    rng = np.random.default_rng(3)
    a, b = rng.integers(0, t, n).astype(np.uint64), rng.integers(0, t, n).astype(np.uint64)
    pa, pb, back = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    assert o.orc_batch_encode(n, t, P(a), n, P(pa)) == 0 and o.orc_batch_encode(n, t, P(b), n, P(pb)) == 0
    assert o.orc_batch_decode(n, t, P(pa), P(back)) == 0 and np.array_equal(back, a)
    prod = [0] * n
    for i in range(n):
        for j in range(n):
            v = int(pa[i]) * int(pb[j])
            if i + j >= n:
                prod[i + j - n] -= v
            else:
                prod[i + j] += v
    pp = np.array([v % t for v in prod], dtype=np.uint64)
    assert o.orc_batch_decode(n, t, P(pp), P(back)) == 0
    assert [int(v) for v in back] == [(int(x) * int(y)) % t for x, y in zip(a, b)]
    # short input: the remaining slots are zero; negative values (two's complement) are lifted by t
    short = np.array([5, (1 << 64) - 3], dtype=np.uint64)
    assert o.orc_batch_encode(n, t, P(short), 2, P(pa)) == 0 and o.orc_batch_decode(n, t, P(pa), P(back)) == 0
    assert int(back[0]) == 5 and int(back[1]) == t - 3 and not back[2:].any()


def test_ckks_decode_inverts_encode_and_composes_centred_integers():
    """orc_ckks_decode (src/ckks.cu:137-190, compose_array src/rns_base.cu:174-258): decode(encode(z)) = z within the
    encoder's rounding at full, lower and one-limb levels; a plaintext holding the constant integer c in every limb
    (c positive, negative, above 2^64) decodes to c / scale in every slot: CRT composition, centring and the word-wise
    conversion to double checked against Python integers."""
    o = H.oracle()
    dp = ctypes.POINTER(ctypes.c_double)
    for n, bits in ((4096, [50, 40, 40, 50]), (8192, [60, 40, 40, 40, 60])):
        ps = H.ParamSet("dec", n, bits, 1, 3, 0)
        oc, slots = ps.octx(), n // 2
        rng = np.random.default_rng(n)
        z = rng.uniform(-3, 3, slots) + 1j * rng.uniform(-3, 3, slots)
        flat = np.ascontiguousarray(z.view(np.float64))
        for l, scale in ((ps.size_Q, 2.0 ** 40), (2, 2.0 ** 40), (1, 2.0 ** 30)):
            plain = np.zeros((l, n), dtype=np.uint64)
            assert o.orc_ckks_encode(oc, l, flat.ctypes.data_as(dp), slots, scale, P(plain)) == 0
            back = np.zeros(2 * slots)
            assert o.orc_ckks_decode(oc, l, P(plain), scale, back.ctypes.data_as(dp)) == 0
            assert np.max(np.abs(back.view(np.complex128) - z)) * scale < n
        l = ps.size_Q
        for c in (12345, -987654321, (1 << 70) + 12345, -(1 << 90) + 7):
            plain = np.zeros((l, n), dtype=np.uint64)   # NTT form of the constant polynomial c: c in every position
            for i in range(l):
                plain[i, :] = c % int(ps.primes[i])
            back = np.zeros(2 * slots)
            scale = 2.0 ** 20
            assert o.orc_ckks_decode(oc, l, P(plain), scale, back.ctypes.data_as(dp)) == 0
            v = back.view(np.complex128)
            assert np.max(np.abs(v.real - c / scale)) <= abs(c / scale) * 1e-12 and np.max(np.abs(v.imag)) <= abs(c / scale) * 1e-12
        assert o.orc_ckks_decode(oc, l, P(plain), 2.0 ** 400, back.ctypes.data_as(dp)) != 0


def test_ckks_encode_is_the_canonical_embedding():
    """orc_ckks_encode (src/ckks.cu:66-135): the encoded polynomial evaluates to scale * z_j at the 2N-th roots zeta^(5^j),
    every limb holds the same centred integer coefficients, short inputs leave the other slots at zero."""
    o = H.oracle()
    n = 4096
    ps = H.ParamSet("enc", n, [50, 40, 40, 50], 1, 3, 0)
    oc, l, slots = ps.octx(), ps.size_Q, n // 2
    rng = np.random.default_rng(0)
    z = rng.uniform(-1, 1, slots) + 1j * rng.uniform(-1, 1, slots)
    z[5:] = 0
    flat = np.ascontiguousarray(z[:5].view(np.float64))
    out = np.zeros((l, n), dtype=np.uint64)
    scale = 2.0 ** 40
    assert o.orc_ckks_encode(oc, l, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 5, scale, P(out)) == 0
    w = out.copy()
    o.orc_ntt_inverse(oc, P(w), l, (ctypes.c_int * l)(*range(l)))
    coef = None
    for i in range(l):
        q = int(ps.primes[i])
        c = np.array([int(v) if int(v) < q // 2 else int(v) - q for v in w[i]], dtype=float)
        assert coef is None or np.array_equal(c, coef)
        coef = c
    pos, m = 1, 2 * n
    for j in range(8):
        val = np.polyval(coef[::-1], np.exp(2j * np.pi * pos / m)) / scale
        assert abs(val - z[j]) < 1e-9
        pos = (pos * 5) % m
    assert o.orc_ckks_encode(oc, l, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 0, scale, P(out)) == -1
    assert o.orc_ckks_encode(oc, l, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 5, 2.0 ** 200, P(out)) == -2


# ---------------------------------------------------------------------------------------------------------
# samplers, key generation, encryption (src/prng.cu, src/secretkey.cu)
# ---------------------------------------------------------------------------------------------------------
def test_prng_core_is_the_salsa20_core():
    """salsa20_gpu (src/prng.cu:17-133) is the Salsa20 core over {key[0:32], nonce, key[32:56]}: the core's published
    known-answer vector (Salsa20 specification, section 8, second example), laid out that way."""
    o = H.oracle()
    inp = [211, 159, 13, 115, 76, 55, 82, 183, 3, 117, 222, 37, 191, 187, 234, 136, 49, 237, 179, 48, 1, 106, 178, 219, 175,
           199, 166, 48, 86, 16, 179, 207, 31, 240, 32, 63, 15, 83, 93, 161, 116, 147, 48, 113, 238, 55, 204, 36, 79, 201,
           235, 79, 3, 81, 156, 47, 203, 26, 244, 243, 88, 118, 104, 54]
    want = [109, 42, 178, 168, 156, 240, 248, 238, 168, 196, 190, 203, 26, 110, 170, 154, 29, 29, 150, 26, 150, 30, 235,
            249, 190, 163, 251, 48, 69, 144, 51, 57, 118, 40, 152, 157, 180, 57, 27, 94, 107, 42, 236, 35, 27, 111, 114,
            114, 219, 236, 232, 135, 111, 155, 110, 18, 24, 232, 95, 158, 179, 19, 48, 202]
    key = bytes(inp[0:32] + inp[40:64] + [0] * 8)
    out = (ctypes.c_uint8 * 64)()
    o.orc_prng_block(out, key, int.from_bytes(bytes(inp[32:40]), "little"))
    assert list(out) == want


def test_samplers_shape_of_the_distributions():
    """orc_sample_poly (src/prng.cu:142-244): ternary and error polynomials are the same small integers in every limb,
    with the expected spread; uniform residues are below q and spread over the range; a different seed gives a different
    polynomial, the same seed the same one."""
    o = H.oracle()
    n = 4096
    ps = H.ParamSet("smp", n, [60, 40, 60], 1, 3, 0)
    oc, m = ps.octx(), ps.size_QP
    seed, other = bytes(range(64)), bytes(range(1, 65))
    out = np.zeros((m, n), dtype=np.uint64)

    def centred(row, q):
        return np.where(row > q // 2, row.astype(np.int64) - np.int64(q), row.astype(np.int64))

    assert o.orc_sample_poly(oc, 0, m, seed, P(out)) == 0
    tern = centred(out[0], int(ps.primes[0]))
    assert set(np.unique(tern)) == {-1, 0, 1} and all(np.array_equal(centred(out[i], int(ps.primes[i])), tern) for i in range(m))
    assert abs(np.mean(tern == 0) - 85 / 256) < 0.03   # byte mod 3: 86 zeros -> -1, 85 -> 0, 85 -> 1
    assert o.orc_sample_poly(oc, 1, m, seed, P(out)) == 0
    err = centred(out[0], int(ps.primes[0]))
    assert np.max(np.abs(err)) <= 21 and all(np.array_equal(centred(out[i], int(ps.primes[i])), err) for i in range(m))
    assert abs(np.var(err) - 10.5) < 1.0 and abs(np.mean(err)) < 0.2   # centred binomial, 21 against 21 bits
    assert o.orc_sample_poly(oc, 2, m, seed, P(out)) == 0
    for i in range(m):
        q = int(ps.primes[i])
        assert int(out[i].max()) < q and abs(float(np.mean(out[i].astype(np.float64))) / q - 0.5) < 0.02
    again, diff = np.zeros_like(out), np.zeros_like(out)
    o.orc_sample_poly(oc, 2, m, seed, P(again))
    o.orc_sample_poly(oc, 2, m, other, P(diff))
    assert np.array_equal(again, out) and not np.array_equal(diff, out)


@pytest.mark.parametrize("scheme", [2, 1, 3])
def test_keygen_and_encryption_semantics(scheme):
    """Keys and ciphertexts made by the oracle's restatement of src/secretkey.cu decrypt (orc_decrypt) to what went in:
    symmetric and public-key encryption, and a relinearisation key from orc_gen_kswitch_key carries a product through
    orc_multiply_relin.  BFV / BGV: exact plaintexts; CKKS: within the noise."""
    o = H.oracle()
    n, t = 4096, 65537
    if scheme == 2:
        ps = H.params_small(n, l=3, alpha=1, qbits=36, pbits=42, scheme=2, t=t)
    else:
        ps = H.params_small(n, l=4, alpha=2, scheme=scheme, t=t if scheme == 1 else 0)
    oc, l, m = ps.octx(), ps.size_Q, ps.size_QP
    rng = np.random.default_rng(scheme)
    seeds = [bytes(rng.integers(0, 256, 64, dtype=np.uint8)) for _ in range(16)]
    sk = np.zeros((m, n), dtype=np.uint64)
    o.orc_gen_secretkey(oc, seeds[0], P(sk))
    kc = o.orc_create(ps.scheme, ps.n, P(ps.primes), m, 0, ps.t)
    sk2 = np.zeros_like(sk)
    o.orc_poly_mul(kc, P(sk), P(sk), P(sk2), m)
    o.orc_destroy(kc)
    pows = np.stack([sk, sk2])
    pk = np.zeros((2, m, n), dtype=np.uint64)
    assert o.orc_encrypt_zero_symmetric(oc, 0, P(sk), seeds[1], seeds[2], P(pk)) == 0
    mul_tech = 2 if scheme == 2 else 0

    def plaintext():
        if scheme == 3:   # small integer coefficients times 2^20, NTT form
            coef = rng.integers(-1000, 1000, n) * (1 << 20)
            pl = np.stack([np.array([int(v) % int(ps.primes[i]) for v in coef], dtype=np.uint64) for i in range(l)])
            o.orc_ntt_forward(oc, P(pl), l, (ctypes.c_int * l)(*range(l)))
            return pl, coef
        pl = rng.integers(0, t, n).astype(np.uint64)
        return pl, pl

    def decrypted(ct, size=2):
        out = np.zeros((l, n) if scheme == 3 else (n,), dtype=np.uint64)
        assert o.orc_decrypt(oc, l, P(ct), size, P(pows), mul_tech, 1, P(out)) == 0
        if scheme != 3:
            return out
        o.orc_ntt_inverse(oc, P(out), l, (ctypes.c_int * l)(*range(l)))
        q0 = int(ps.primes[0])
        return np.array([int(v) - q0 if int(v) > q0 // 2 else int(v) for v in out[0]])

    def check(ct, want, what):
        got = decrypted(ct)
        if scheme == 3:
            assert np.max(np.abs(got - want)) < 1 << 12, what
        else:
            assert np.array_equal(got, want), what

    pl, want = plaintext()
    ct = np.zeros((2, l, n), dtype=np.uint64)
    assert o.orc_encrypt_zero_symmetric(oc, 1, P(sk), seeds[3], seeds[4], P(ct)) == 0
    assert o.orc_encrypt_add_plain(oc, l, P(ct), P(pl)) == 0
    check(ct, want, "symmetric encryption")
    ct2 = np.zeros((2, l, n), dtype=np.uint64)
    assert o.orc_encrypt_zero_asymmetric(oc, P(pk), seeds[5], seeds[6], P(ct2)) == 0
    assert o.orc_encrypt_add_plain(oc, l, P(ct2), P(pl)) == 0
    check(ct2, want, "public-key encryption")
    if scheme == 3:
        return
    # relinearisation key: Enc(a) * Enc(b) relinearised decrypts to the negacyclic product a * b mod t
    dnum = ps.size_Q // ps.size_P
    rlk = np.zeros((dnum, 2, m, n), dtype=np.uint64)
    kseeds = b"".join(seeds[7:7 + 2 * dnum])
    assert o.orc_gen_kswitch_key(oc, P(sk2), P(sk), kseeds, P(rlk)) == 0
    a = np.zeros(n, dtype=np.uint64)
    a[0], a[1] = 3, 5
    b = np.zeros(n, dtype=np.uint64)
    b[0], b[n - 1] = 7, 2
    cts = []
    for k, p_ in enumerate((a, b)):
        c_ = np.zeros((2, l, n), dtype=np.uint64)
        assert o.orc_encrypt_zero_symmetric(oc, 1, P(sk), seeds[12 + k], seeds[14 + k], P(c_)) == 0
        assert o.orc_encrypt_add_plain(oc, l, P(c_), P(p_)) == 0
        cts.append(c_)
    out = np.zeros((2, l, n), dtype=np.uint64)
    if scheme == 2:
        assert o.orc_bfv_multiply_relin_hps(oc, P(cts[0]), P(cts[1]), P(rlk), P(out)) == 0
    else:
        o.orc_multiply_relin(oc, l, P(cts[0]), P(cts[1]), P(rlk), P(out))
    want = np.zeros(n, dtype=np.int64)   # (3 + 5x)(7 + 2x^(n-1)) = 21 + 35x + 6x^(n-1) - 10
    want[0], want[1], want[n - 1] = 11, 35, 6
    got = decrypted(out).astype(np.int64) % t   # the reference's HPS rounding returns t itself for some zero coefficients
    assert np.array_equal(got, want), ("product through the generated relinearisation key", np.nonzero(got != want)[0][:8], got[got != want][:8])


def test_samplers_known_answers():
    """tests/golden/sampler_kat.json (see make_sampler_kat.py: oracle output that the GPU tests show equal to the reference's
    sample_*_poly kernels): heads and SHA-256 of whole polynomials for three seeds and the three samplers."""
    import hashlib
    import json
    kat = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_kat.json")))
    ps = H.ParamSet("kat", kat["n"], kat["bits"], 1, 3, 0)
    assert [int(p) for p in ps.primes] == kat["primes"]
    o, oc, m = H.oracle(), ps.octx(), ps.size_QP
    seeds = dict(counting=bytes(range(64)), zeros=bytes(64), ones=bytes([255] * 64))
    for case in kat["cases"]:
        out = np.zeros((m, kat["n"]), dtype=np.uint64)
        assert o.orc_sample_poly(oc, case["kind"], m, seeds[case["seed"]], P(out)) == 0
        assert [[int(v) for v in out[i, :8]] for i in range(m)] == case["head"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"]
