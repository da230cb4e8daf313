"""Kernel-level boundary (cut line 2) and the parity holes of round 1, all against the UNMODIFIED
reference on the same words (oracle/_ref/libphantom_ref.so, oracle/ref_shim.cu): every launcher of include/ntt.cuh:172-226,
DBaseConverter::bConv_BEHZ / _var1 / _HPS, DRNSTool::moddown / divide_and_round_q_last[_ntt] / mod_t_and_divide_q_last_ntt,
the stage taps of the key switch for BGV and BFV, hoisting, the alpha = 5 / 6 sets of ckks_bench.cu, BFV at t = 65537,
all 32 rotation steps at N = 2^16 and the NAF recursion of rotate_inplace."""
import ctypes

import numpy as np
import pytest
import torch

import harness as H
from harness import P

pytestmark = pytest.mark.gpu

pf = None
lib = None


def setup_module(module):
    global pf, lib
    import phantom_fhe_b200 as m
    pf = m
    lib = m.lib
    r = H.reference()
    if r is not None and not hasattr(r.ref_nwt, "_typed"):
        sz, vp, u64p = ctypes.c_size_t, ctypes.c_void_p, H.u64p
        r.ref_nwt.argtypes = [vp, ctypes.c_int, ctypes.c_int, u64p, u64p, sz, sz, sz, ctypes.POINTER(sz), u64p, sz, u64p]
        r.ref_table_moduli.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_int]
        r.ref_bconv.argtypes = [vp, sz, ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p, sz, u64p, sz]
        r.ref_bconv_bases.argtypes = [vp, sz, ctypes.c_int, ctypes.c_int, u64p, ctypes.c_int]
        r.ref_moddown_plain.argtypes = [vp, sz, u64p, u64p]
        r.ref_divide_round.argtypes = [vp, ctypes.c_int, sz, u64p, sz, u64p]
        r.ref_hoisting.argtypes = [vp, sz, u64p, H.i32p, ctypes.c_int, sz, u64p]
        r.ref_keyswitch.argtypes = [vp, sz, u64p, u64p, u64p]
        r.ref_nwt._typed = True


def ref_or_skip():
    r = H.reference()
    if r is None:
        pytest.skip("oracle/_ref/libphantom_ref.so was not built")
    return r


def make_context(ps, steps=(), mul_tech=None):
    parms = pf.EncryptionParameters(pf.scheme_type(ps.scheme))
    parms.set_poly_modulus_degree(ps.n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(ps.size_P)
    if ps.t:
        parms.set_plain_modulus(ps.t)
    if mul_tech is not None:
        parms.set_mul_tech(pf.mul_tech_type(mul_tech))
    if steps:
        parms.set_galois_elts(pf.get_elts_from_steps(list(steps), ps.n))
    return pf.PhantomContext(parms)


def ref_context(r, ps, steps=(), mul_tech=0, gen_keys=1):
    arr = (ctypes.c_int * max(1, len(steps)))(*steps)
    h = r.ref_create(ps.scheme, ps.n, P(ps.primes), ps.size_QP, ps.size_P, ps.t, mul_tech, arr, len(steps), float(2 ** 40),
                     gen_keys)
    assert h, r.ref_last_error()
    return h


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy().view(np.uint64)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def uniform(moduli, n, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros((len(moduli), n), dtype=np.uint64)
    for i, q in enumerate(moduli):
        out[i] = rng.integers(0, int(q), n, dtype=np.uint64)
    return out


def shoup(x, q):
    return (int(x) << 64) // int(q)


def keys_from_reference(r, h, ctx, ps, which=-1):
    dnum = r.ref_dnum(h)
    k = np.zeros((dnum, 2, ps.size_QP, ps.n), dtype=np.uint64)
    for d in range(dnum):
        assert r.ref_key_get(h, which, d, P(k[d])) == 0
    return pf.PhantomRelinKey(ctx, list(k))


SZ = ctypes.c_size_t


def call_ref_nwt(r, h, variant, table, buf, aux, count, start, prm=(), scale=None, scale_mod=None):
    prm_arr = (SZ * 4)(*(list(prm) + [0] * (4 - len(prm))))
    want = buf.copy()
    ns = 0 if scale is None else len(scale)
    rc = r.ref_nwt(h, variant, table, P(want), None if aux is None else P(aux), buf.shape[0], count, start, prm_arr,
                   None if scale is None else P(scale), ns, None if scale is None else P(scale_mod))
    assert rc == 0, r.ref_last_error()
    return want


# ---------------------------------------------------------------------------------------------------------
# the 13 launchers of include/ntt.cuh:172-226
# ---------------------------------------------------------------------------------------------------------
def test_nwt_launchers_on_the_key_tables_against_reference():
    r = ref_or_skip()
    ps = H.params_small(8192, l=5, alpha=2)     # size_QP = 7: rows 0..4 = Q, 5..6 = P
    h = ref_context(r, ps, gen_keys=0)
    try:
        ctx = make_context(ps)
        e, st, n = ctx._h, stream(), ps.n
        primes = [int(p) for p in ps.primes]
        R = pf.lib  # noqa: N806
        # (a) plain forms with a start index: limbs [2, 5) of a 7-limb buffer
        x = uniform(primes, n, 1)
        for variant, fn in ((0, R.pfhe_ntt_forward_inplace), (1, R.pfhe_ntt_backward_inplace)):
            want = call_ref_nwt(r, h, variant, 0, x, None, 3, 2)
            d = dev(x)
            pf.check(fn(e, d.data_ptr(), 3, 2, st))
            assert np.array_equal(host(d), want), f"launcher {variant} with start index"
            d = dev(x)
            pf.check((R.pfhe_nwt_2d_radix8_forward_inplace, R.pfhe_nwt_2d_radix8_backward_inplace)[variant](
                e, 0, d.data_ptr(), 3, 2, st))
            assert np.array_equal(host(d), want)
        # out of place inverse (nwt_2d_radix8_backward): untouched limbs of `out` keep their words
        y = uniform(primes, n, 2)
        want = call_ref_nwt(r, h, 2, 0, y, x, 4, 1)
        d, s_ = dev(y), dev(x)
        pf.check(R.pfhe_ntt_backward(e, d.data_ptr(), s_.data_ptr(), 4, 1, st))
        assert np.array_equal(host(d), want), "nwt_2d_radix8_backward"
        assert np.array_equal(host(s_), x)
        # (b) include_special_mod on a packed Ql u P buffer at a lower level: l = 3 data limbs + 2 special limbs
        l = 3
        rows = list(range(l)) + [5, 6]
        pk = uniform([primes[i] for i in rows], n, 3)
        for variant, fn in ((5, R.pfhe_ntt_forward_inplace_include_special_mod),
                            (10, R.pfhe_ntt_backward_inplace_include_special_mod)):
            for count, start in ((l + 2, 0), (2, l), (3, 1)):
                want = call_ref_nwt(r, h, variant, 0, pk, None, count, start, (ps.size_QP, ps.size_P))
                d = dev(pk)
                pf.check(fn(e, d.data_ptr(), count, start, ps.size_QP, ps.size_P, st))
                assert np.array_equal(host(d), want), f"include_special_mod {variant} count {count} start {start}"
        # exclude_range (the mod-up transform that leaps over the digit's own limbs)
        want = call_ref_nwt(r, h, 6, 0, pk, None, l + 2, 0, (ps.size_QP, ps.size_P, 1, 3))
        d = dev(pk)
        pf.check(R.pfhe_nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range(
            e, d.data_ptr(), l + 2, 0, ps.size_QP, ps.size_P, 1, 3, st))
        assert np.array_equal(host(d), want), "exclude_range"
        assert np.array_equal(host(d)[1:3], pk[1:3])
        # (c) fuse_moddown: ct = (cx - NTT(delta)) * c_i with caller-supplied constants
        delta, cx = uniform(primes[:l], n, 4), uniform(primes[:l], n, 5)
        c = np.array([pow(primes[5] * primes[6], -1, primes[i]) for i in range(l)], dtype=np.uint64)
        cs = np.array([shoup(c[i], primes[i]) for i in range(l)], dtype=np.uint64)
        want = call_ref_nwt(r, h, 3, 0, delta, cx, l, 0, (), c, np.array(primes[:l], dtype=np.uint64))
        d_delta, d_cx, d_ct = dev(delta), dev(cx), torch.zeros((l, n), dtype=torch.int64, device="cuda")
        dc, dcs = dev(c), dev(cs)
        pf.check(R.pfhe_nwt_2d_radix8_forward_inplace_fuse_moddown(e, d_ct.data_ptr(), d_cx.data_ptr(), dc.data_ptr(),
                                                                   dcs.data_ptr(), d_delta.data_ptr(), l, 0, st))
        assert np.array_equal(host(d_ct), want), "forward_inplace_fuse_moddown"
        # (d) forward_modup_fuse: a small-modulus polynomial lifted under one prime
        small = uniform([65537] * 2, n, 6)
        for mi in (0, 4):
            want = call_ref_nwt(r, h, 7, 0, np.zeros_like(small), small, 2, 0, (mi,))
            d_out, d_in = torch.zeros((2, n), dtype=torch.int64, device="cuda"), dev(small)
            pf.check(R.pfhe_nwt_2d_radix8_forward_modup_fuse(e, d_out.data_ptr(), d_in.data_ptr(), mi, 2, 0, st))
            assert np.array_equal(host(d_out), want), f"forward_modup_fuse under prime {mi}"
        # (e) scale forms
        sc = np.array([(12345 + 977 * i) % q for i, q in enumerate(primes)], dtype=np.uint64)
        scs = np.array([shoup(sc[i], primes[i]) for i in range(len(primes))], dtype=np.uint64)
        mods = np.array(primes, dtype=np.uint64)
        dsc, dscs = dev(sc), dev(scs)
        want = call_ref_nwt(r, h, 8, 0, y, x, 5, 1, (), sc, mods)
        d, s_ = dev(y), dev(x)
        pf.check(R.pfhe_nwt_2d_radix8_backward_scale(e, 0, d.data_ptr(), s_.data_ptr(), 5, 1, dsc.data_ptr(), dscs.data_ptr(), st))
        assert np.array_equal(host(d), want), "backward_scale"
        want = call_ref_nwt(r, h, 9, 0, x, None, 7, 0, (), sc, mods)
        d = dev(x)
        pf.check(R.pfhe_nwt_2d_radix8_backward_inplace_scale(e, 0, d.data_ptr(), 7, 0, dsc.data_ptr(), dscs.data_ptr(), st))
        assert np.array_equal(host(d), want), "backward_inplace_scale"
    finally:
        r.ref_destroy(h)


def test_nwt_launchers_on_the_bfv_tables_against_reference():
    """include_temp_mod forms over gpu_Bsk_tables (BEHZ) and transforms over gpu_QlRl_tables (HPS)"""
    r = ref_or_skip()
    ps = H.params_bfv_bench(0)
    for mul_tech, table in ((1, 1), (2, 2)):
        h = ref_context(r, ps, mul_tech=mul_tech, gen_keys=0)
        try:
            ctx = make_context(ps, mul_tech=mul_tech)
            e, st, n = ctx._h, stream(), ps.n
            mods = np.zeros(64, dtype=np.uint64)
            cnt = r.ref_table_moduli(h, table, P(mods), 64)
            assert cnt == lib.pfhe_table_size(e, table), "table sizes"
            mods = mods[:cnt]
            assert [int(lib.pfhe_table_modulus(e, table, i)) for i in range(cnt)] == [int(v) for v in mods], "table moduli"
            x = uniform(mods, n, 11)
            want = call_ref_nwt(r, h, 0, table, x, None, cnt, 0)
            d = dev(x)
            pf.check(lib.pfhe_nwt_2d_radix8_forward_inplace(e, table, d.data_ptr(), cnt, 0, st))
            assert np.array_equal(host(d), want), "forward over the auxiliary table"
            want = call_ref_nwt(r, h, 1, table, x, None, cnt - 1, 1)
            d = dev(x)
            pf.check(lib.pfhe_nwt_2d_radix8_backward_inplace(e, table, d.data_ptr(), cnt - 1, 1, st))
            assert np.array_equal(host(d), want), "inverse over the auxiliary table"
            if table == 1:
                want = call_ref_nwt(r, h, 4, table, x, None, cnt, 0, (cnt,))
                d = dev(x)
                pf.check(lib.pfhe_nwt_2d_radix8_forward_inplace_include_temp_mod(e, table, d.data_ptr(), cnt, 0, cnt, st))
                assert np.array_equal(host(d), want), "forward_inplace_include_temp_mod"
                sc = np.array([ps.t % int(q) for q in mods], dtype=np.uint64)   # tModBsk (evaluate.cu:528-531)
                scs = np.array([shoup(sc[i], mods[i]) for i in range(cnt)], dtype=np.uint64)
                want = call_ref_nwt(r, h, 11, table, x, None, cnt, 0, (cnt,), sc, mods)
                d, dsc, dscs = dev(x), dev(sc), dev(scs)
                pf.check(lib.pfhe_nwt_2d_radix8_backward_inplace_include_temp_mod_scale(
                    e, table, d.data_ptr(), cnt, 0, cnt, dsc.data_ptr(), dscs.data_ptr(), st))
                assert np.array_equal(host(d), want), "backward_inplace_include_temp_mod_scale"
        finally:
            r.ref_destroy(h)


# ---------------------------------------------------------------------------------------------------------
# DBaseConverter, DRNSTool members
# ---------------------------------------------------------------------------------------------------------
def _bases(r, h, chain_index, which, aux):
    buf = np.zeros(128, dtype=np.uint64)
    code = r.ref_bconv_bases(h, chain_index, which, aux, P(buf), 128)
    assert code > 0, r.ref_last_error()
    ni, no = code >> 16, code & 0xffff
    return [int(v) for v in buf[:ni]], [int(v) for v in buf[ni:ni + no]]


def _entries(ctx, moduli):
    """(table << 16 | entry) of each modulus, looked up in the engine's tables"""
    e = ctx._h
    where = {}
    for table in (0, 2, 1):
        for i in range(lib.pfhe_table_size(e, table)):
            where.setdefault(int(lib.pfhe_table_modulus(e, table, i)), (table << 16) | i)
    return (ctypes.c_uint32 * len(moduli))(*[where[int(q)] for q in moduli])


def test_bconv_against_reference():
    r = ref_or_skip()
    cases = []
    ps = H.params_small(8192, l=6, alpha=3)
    cases.append((ps, None, [(1, 0, 0, 0), (1, 1, 0, 0), (1, 1, 1, 0), (2, 0, 0, 0), (1, 0, 0, 1)]))   # P->Ql, digits; also var1
    psb = H.params_bfv_bench(0)
    cases.append((psb, 2, [(1, 2, 0, 2), (1, 3, 0, 2), (1, 2, 0, 1), (1, 2, 0, 0)]))   # Ql->Rl, Rl->Ql: HPS, var1, BEHZ
    for ps, mul_tech, convs in cases:
        h = ref_context(r, ps, mul_tech=mul_tech or 0, gen_keys=0)
        try:
            ctx = make_context(ps, mul_tech=mul_tech)
            e, st, n = ctx._h, stream(), ps.n
            for ci, which, aux, mode in convs:
                ib, ob = _bases(r, h, ci, which, aux)
                x = uniform(ib, n, 20 + which + 7 * mode)
                want = np.zeros((len(ob), n), dtype=np.uint64)
                assert r.ref_bconv(h, ci, which, aux, mode, P(x), len(ib), P(want), len(ob)) == 0, r.ref_last_error()
                d_in, d_out = dev(x), torch.zeros((len(ob), n), dtype=torch.int64, device="cuda")
                pf.check(lib.pfhe_bconv(e, mode, _entries(ctx, ib), len(ib), _entries(ctx, ob), len(ob), d_out.data_ptr(),
                                        d_in.data_ptr(), st))
                assert np.array_equal(host(d_out), want), f"bconv which={which} mode={mode} ({len(ib)} -> {len(ob)} limbs)"
        finally:
            r.ref_destroy(h)


@pytest.mark.parametrize("scheme", [3, 1, 2])
def test_moddown_and_divide_round_against_reference(scheme):
    r = ref_or_skip()
    t = 65537 if scheme != 3 else 0
    ps = H.params_small(8192, l=5, alpha=2, scheme=scheme, t=t)
    h = ref_context(r, ps, mul_tech=2 if scheme == 2 else 0, gen_keys=0)
    try:
        ctx = make_context(ps)
        e, st, n = ctx._h, stream(), ps.n
        primes = [int(p) for p in ps.primes]
        for ci in (1, 2):
            l = ps.limbs(ci)
            rows = list(range(l)) + [ps.size_Q + i for i in range(ps.size_P)]
            cx = uniform([primes[i] for i in rows], n, 40 + ci)
            # DRNSTool::moddown
            want = np.zeros((l, n), dtype=np.uint64)
            assert r.ref_moddown_plain(h, ci, P(cx), P(want)) == 0, r.ref_last_error()
            d_cx, d_ct = dev(cx), torch.zeros((l, n), dtype=torch.int64, device="cuda")
            pf.check(lib.pfhe_moddown(e, ci, d_ct.data_ptr(), d_cx.data_ptr(), st))
            assert np.array_equal(host(d_ct), want), f"DRNSTool::moddown scheme {scheme} level {ci}"
            # moddown_from_NTT through the stage tap
            want = np.zeros((l, n), dtype=np.uint64)
            assert r.ref_moddown(h, ci, P(cx), P(want)) == 0, r.ref_last_error()
            d_cx, d_ct = dev(cx), torch.zeros((l, n), dtype=torch.int64, device="cuda")
            pf.check(lib.pfhe_moddown_from_ntt(e, ci, d_ct.data_ptr(), d_cx.data_ptr(), st))
            assert np.array_equal(host(d_ct), want), f"moddown_from_NTT scheme {scheme} level {ci}"
            # divide-and-round family
            src = np.stack([uniform(primes[:l], n, 50 + k) for k in range(2)])
            variant = {3: 0, 2: 1, 1: 2}[scheme]
            fn = {0: lib.pfhe_divide_and_round_q_last_ntt, 1: lib.pfhe_divide_and_round_q_last,
                  2: lib.pfhe_mod_t_and_divide_q_last_ntt}[variant]
            want = np.zeros((2, l - 1, n), dtype=np.uint64)
            assert r.ref_divide_round(h, variant, ci, P(src), 2, P(want)) == 0, r.ref_last_error()
            d_src, d_dst = dev(src), torch.zeros((2, l - 1, n), dtype=torch.int64, device="cuda")
            pf.check(fn(e, ci, d_src.data_ptr(), 2, d_dst.data_ptr(), st))
            assert np.array_equal(host(d_dst), want), f"divide-and-round variant {variant} level {ci}"
            assert np.array_equal(host(d_src), src)
            # the scheme-level form of the same step
            want = np.zeros((2, l - 1, n), dtype=np.uint64)
            if scheme == 3:
                assert r.ref_rescale(h, ci, P(src), 2, P(want)) == 0, r.ref_last_error()
                got = pf.rescale_to_next(ctx, pf.PhantomCiphertext.from_host(ctx, src, chain_index=ci)).to_host()
            else:
                assert r.ref_mod_switch(h, ci, P(src), 2, P(want)) == 0, r.ref_last_error()
                got = pf.mod_switch_to_next(ctx, pf.PhantomCiphertext.from_host(ctx, src, chain_index=ci,
                                                                                 is_ntt_form=(scheme != 2))).to_host()
            assert np.array_equal(got, want), f"mod_switch_to_next / rescale scheme {scheme} level {ci}"
        # add_to_ct
        a, b = uniform(primes[:5], n, 60), uniform(primes[:5], n, 61)
        d_a, d_b = dev(a), dev(b)
        pf.check(lib.pfhe_add_to_ct(e, d_a.data_ptr(), d_b.data_ptr(), 5, st))
        q = np.array(primes[:5], dtype=object).reshape(-1, 1)
        assert np.array_equal(host(d_a), ((a.astype(object) + b.astype(object)) % q).astype(np.uint64))
    finally:
        r.ref_destroy(h)


@pytest.mark.parametrize("scheme", [3, 1, 2])
def test_key_switch_stages_against_reference(scheme):
    """modup -> inner product -> moddown_from_NTT -> keyswitch_inplace, each stage against the reference's own"""
    r = ref_or_skip()
    t = 65537 if scheme != 3 else 0
    ps = H.params_small(8192, l=6, alpha=2, scheme=scheme, t=t)
    h = ref_context(r, ps, mul_tech=2 if scheme == 2 else 0)
    try:
        ctx = make_context(ps)
        e, st, n = ctx._h, stream(), ps.n
        primes = [int(p) for p in ps.primes]
        rlk = keys_from_reference(r, h, ctx, ps)
        for ci in (1, 3):
            l = ps.limbs(ci)
            m, beta = l + ps.size_P, ps.beta(ci)
            c2 = uniform(primes[:l], n, 70 + ci)
            want = np.zeros((beta, m, n), dtype=np.uint64)
            assert r.ref_modup(h, ci, P(c2), P(want)) == 0, r.ref_last_error()
            d_c2, d_t = dev(c2), torch.zeros((beta, m, n), dtype=torch.int64, device="cuda")
            pf.check(lib.pfhe_modup(e, ci, d_t.data_ptr(), d_c2.data_ptr(), st))
            assert np.array_equal(host(d_t), want), f"modup scheme {scheme} level {ci}"
            want_cx = np.zeros((2, m, n), dtype=np.uint64)
            assert r.ref_inner_prod(h, ci, -1, P(want), P(want_cx)) == 0, r.ref_last_error()
            d_cx = torch.zeros((2, m, n), dtype=torch.int64, device="cuda")
            pf.check(lib.pfhe_key_switch_inner_prod(e, ci, d_cx.data_ptr(), d_t.data_ptr(), rlk.public_keys_ptr(), st))
            assert np.array_equal(host(d_cx), want_cx), f"inner product scheme {scheme} level {ci}"
            if scheme == 2 and ci != 1:
                continue   # the reference's BFV keyswitch_inplace takes the first level's tool whatever the level
            ct = np.stack([uniform(primes[:l], n, 80 + k) for k in range(2)])
            want_ks = np.zeros((2, l, n), dtype=np.uint64)
            assert r.ref_keyswitch(h, ci, P(ct), P(c2), P(want_ks)) == 0, r.ref_last_error()
            d_ct, d_c2 = dev(ct), dev(c2)
            pf.check(lib.pfhe_keyswitch_inplace(e, ci, d_ct.data_ptr(), d_c2.data_ptr(), rlk.public_keys_ptr(), st))
            assert np.array_equal(host(d_ct), want_ks), f"keyswitch_inplace scheme {scheme} level {ci}"
    finally:
        r.ref_destroy(h)


# ---------------------------------------------------------------------------------------------------------
# hoisting, rotations
# ---------------------------------------------------------------------------------------------------------
def _galois_keys_from_reference(r, h, ctx, ps):
    cnt = r.ref_galois_count(h)
    assert [int(r.ref_galois_elt_at(h, i)) for i in range(cnt)] == list(ctx.parms.galois_elts), "Galois element order"
    glk = pf.PhantomGaloisKey.__new__(pf.PhantomGaloisKey)
    glk.relin_keys = [keys_from_reference(r, h, ctx, ps, which=i) for i in range(cnt)]
    return glk


@pytest.mark.parametrize("scheme,mul_tech", [(3, 0), (1, 0), (2, 2), (2, 1), (2, 4)])
def test_hoisting_against_reference(scheme, mul_tech):
    r = ref_or_skip()
    t = 65537 if scheme != 3 else 0
    ps = H.params_small(8192, l=6, alpha=2, scheme=scheme, t=t)
    steps = [1, 2, 3, -1, 5]
    h = ref_context(r, ps, steps=steps, mul_tech=mul_tech)
    try:
        ctx = make_context(ps, steps=steps, mul_tech=mul_tech if scheme == 2 else None)
        glk = _galois_keys_from_reference(r, h, ctx, ps)
        primes = [int(p) for p in ps.primes]
        for ci in ((1, 2) if scheme != 2 else (1,)):   # BFV: the reference addresses the first level's tool only
            l = ps.limbs(ci)
            ct = np.stack([uniform(primes[:l], ps.n, 90 + k) for k in range(2)])
            for deg in ((1, 3) if mul_tech == 4 else (1,)):
                want = np.zeros((2, l, ps.n), dtype=np.uint64)
                arr = (ctypes.c_int * len(steps))(*steps)
                assert r.ref_hoisting(h, ci, P(ct), arr, len(steps), deg, P(want)) == 0, r.ref_last_error()
                c = pf.PhantomCiphertext.from_host(ctx, ct, chain_index=ci, is_ntt_form=(scheme != 2))
                c.noise_scale_deg = deg
                pf.hoisting_inplace(ctx, c, glk, steps)
                assert np.array_equal(c.to_host(), want), f"hoisting scheme {scheme} mul_tech {mul_tech} level {ci} deg {deg}"
    finally:
        r.ref_destroy(h)


def test_rotate_32_steps_full_size_against_reference():
    """BASELINE.json configs[3]: rotate_inplace for every step 1..32 at N = 2^16, L = 16 with the reference's own keys"""
    r = ref_or_skip()
    ps = H.params_primary()
    steps = list(range(1, 33))
    h = ref_context(r, ps, steps=steps)
    try:
        ctx = make_context(ps, steps=steps)
        glk = _galois_keys_from_reference(r, h, ctx, ps)
        a = H.ciphertext(ps, 1)
        l, n = ps.limbs(), ps.n
        for s in steps:
            want = np.zeros((2, l, n), dtype=np.uint64)
            assert r.ref_rotate(h, 1, P(a), s, P(want)) == 0, r.ref_last_error()
            c = pf.PhantomCiphertext.from_host(ctx, a)
            pf.rotate_inplace(ctx, c, s, glk)
            assert np.array_equal(c.to_host(), want), f"rotate step {s}"
    finally:
        r.ref_destroy(h)


def test_rotate_naf_recursion_against_reference():
    """no Galois elements given: both sides fall back to the default set and rotate through the NAF decomposition
    (evaluate.cu:1649-1661); N = 2^16 at the primary set, steps that need 2 and 3 key switches"""
    r = ref_or_skip()
    ps = H.params_primary()
    h = ref_context(r, ps, steps=(), gen_keys=2)
    try:
        ctx = make_context(ps)
        glk = _galois_keys_from_reference(r, h, ctx, ps)
        a = H.ciphertext(ps, 3)
        l, n = ps.limbs(), ps.n
        for s in (3, 7, -5, 1, 1000):
            want = np.zeros((2, l, n), dtype=np.uint64)
            assert r.ref_rotate(h, 1, P(a), s, P(want)) == 0, r.ref_last_error()
            c = pf.PhantomCiphertext.from_host(ctx, a)
            pf.rotate_inplace(ctx, c, s, glk)
            assert np.array_equal(c.to_host(), want), f"NAF rotation by {s}"
    finally:
        r.ref_destroy(h)


# ---------------------------------------------------------------------------------------------------------
# parameter sets the first round left out
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits,size_P", [([60] + [40] * 34 + [60] * 5, 5), ([60] + [40] * 29 + [60] * 6, 6)])
def test_ckks_bench_sets_with_five_and_six_special_primes(bits, size_P):
    """benchmark/ckks_bench.cu:371-394: HMult+Relin and rotate at N = 2^16 with alpha = 5 / 6, the reference's own keys"""
    r = ref_or_skip()
    ps = H.ParamSet(f"ckks16_p{size_P}", 65536, bits, size_P)
    h = ref_context(r, ps, steps=[1])
    try:
        ctx = make_context(ps, steps=[1])
        rlk = keys_from_reference(r, h, ctx, ps)
        glk = _galois_keys_from_reference(r, h, ctx, ps)
        a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
        l, n = ps.limbs(), ps.n
        want = np.zeros((2, l, n), dtype=np.uint64)
        assert r.ref_multiply_relin(h, 1, P(a), P(b), P(want)) == 0, r.ref_last_error()
        ca, cb = pf.PhantomCiphertext.from_host(ctx, a), pf.PhantomCiphertext.from_host(ctx, b)
        pf.multiply_and_relin_inplace(ctx, ca, cb, rlk)
        assert np.array_equal(ca.to_host(), want), "HMult+Relin (fused entry point)"
        ca = pf.PhantomCiphertext.from_host(ctx, a)
        pf.multiply_inplace(ctx, ca, cb)
        pf.relinearize_inplace(ctx, ca, rlk)
        assert np.array_equal(ca.to_host(), want), "multiply_inplace + relinearize_inplace"
        want = np.zeros((2, l, n), dtype=np.uint64)
        assert r.ref_rotate(h, 1, P(a), 1, P(want)) == 0, r.ref_last_error()
        c = pf.PhantomCiphertext.from_host(ctx, a)
        pf.rotate_inplace(ctx, c, 1, glk)
        assert np.array_equal(c.to_host(), want), "rotate"
    finally:
        r.ref_destroy(h)


@pytest.mark.parametrize("mul_tech", [1, 2, 3])
def test_bfv_config3_plain_modulus_65537_against_reference(mul_tech):
    """BASELINE.json configs[2]: BFV HMult+Relin, N = 2^14, {54x7, 60}, t = 65537, BEHZ / HPS / HPS over Q"""
    r = ref_or_skip()
    ps = H.ParamSet("bfv14_t65537", 16384, [54] * 7 + [60], 1, scheme=2, t=65537)
    h = ref_context(r, ps, mul_tech=mul_tech)
    try:
        ctx = make_context(ps, mul_tech=mul_tech)
        rlk = keys_from_reference(r, h, ctx, ps)
        a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
        l, n = ps.limbs(), ps.n
        want = np.zeros((2, l, n), dtype=np.uint64)
        assert r.ref_multiply_relin(h, 1, P(a), P(b), P(want)) == 0, r.ref_last_error()
        ca = pf.PhantomCiphertext.from_host(ctx, a, is_ntt_form=False)
        cb = pf.PhantomCiphertext.from_host(ctx, b, is_ntt_form=False)
        pf.multiply_inplace(ctx, ca, cb)
        pf.relinearize_inplace(ctx, ca, rlk)
        assert np.array_equal(ca.to_host(), want), f"BFV multiply + relinearize, mul_tech {mul_tech}"
        ca = pf.PhantomCiphertext.from_host(ctx, a, is_ntt_form=False)
        pf.multiply_and_relin_inplace(ctx, ca, cb, rlk)
        assert np.array_equal(ca.to_host(), want), f"BFV multiply_and_relin, mul_tech {mul_tech}"
    finally:
        r.ref_destroy(h)


# ---------------------------------------------------------------------------------------------------------
# one engine, several host threads (the reference is built --default-stream per-thread, src/CMakeLists.txt:39)
# ---------------------------------------------------------------------------------------------------------
def test_concurrent_host_threads_share_one_context():
    import threading
    ps = H.params_small(8192, l=6, alpha=2)
    ctx = make_context(ps)
    o = H.oracle()
    rlk_h = H.switch_key(ps, 100)
    rlk = pf.PhantomRelinKey(ctx, list(rlk_h))
    l, n = ps.limbs(), ps.n
    n_threads, per = 4, 6
    pairs = [(H.ciphertext(ps, 200 + 2 * i), H.ciphertext(ps, 201 + 2 * i)) for i in range(n_threads)]
    want = []
    for a, b in pairs:
        w = np.zeros((2, l, n), dtype=np.uint64)
        o.orc_multiply_relin(ps.octx(), l, P(a), P(b), P(rlk_h), P(w))
        want.append(w)
    results, errors = [None] * n_threads, []

    def work(k):
        try:
            torch.cuda.set_device(0)
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                a, b = dev(pairs[k][0]), dev(pairs[k][1])
                outs = []
                for _ in range(per):   # every thread keeps several ops in flight on its own stream
                    out = torch.empty_like(a)
                    pf.check(lib.pfhe_multiply_and_relin(ctx._h, 1, a.data_ptr(), b.data_ptr(), out.data_ptr(),
                                                         rlk.public_keys_ptr(), ctypes.c_void_p(s.cuda_stream)))
                    outs.append(out)
                s.synchronize()
                results[k] = [o_.cpu().numpy().view(np.uint64) for o_ in outs]
        except Exception as ex:   # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for k in range(n_threads):
        for got in results[k]:
            assert np.array_equal(got, want[k]), f"thread {k}"
