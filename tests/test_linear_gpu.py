"""GPU parity tests of the linear part of the evaluator surface (src/evaluate.cu:14-338, 1106-1340): negate / add / sub /
add_many with the BGV correction-factor balancing, add_plain / sub_plain / multiply_plain, mod_switch_to -- host mirror over
the C-ABI against the oracle and against the unmodified reference on the same words."""
import ctypes

import numpy as np
import pytest
import torch

import harness as H
from harness import P

pytestmark = pytest.mark.gpu

pf = None


def setup_module(module):
    global pf
    import phantom_fhe_b200 as m
    pf = m


def make_context(ps):
    parms = pf.EncryptionParameters(pf.scheme_type(ps.scheme))
    parms.set_poly_modulus_degree(ps.n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(ps.size_P)
    if ps.t:
        parms.set_plain_modulus(ps.t)
    if ps.scheme == 2:
        parms.set_mul_tech(2)
    return pf.PhantomContext(parms)


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy().view(np.uint64)


def param_set(scheme, n=4096):
    if scheme == 2:
        return H.params_small(n, l=3, alpha=1, qbits=36, pbits=42, scheme=2, t=65537)
    return H.params_small(n, l=4, alpha=2, scheme=scheme, t=65537 if scheme == 1 else 0)


def reference_for(ps, scale):
    r = H.reference()
    if r is None or not hasattr(r, "ref_plain_op"):
        return None, None
    h = r.ref_create(ps.scheme, ps.n, P(ps.primes), ps.size_QP, ps.size_P, ps.t, 2, None, 0, scale, 0)
    assert h, r.ref_last_error()
    return r, h


@pytest.mark.parametrize("scheme", [2, 1, 3])
def test_plain_operands(scheme):
    """add_plain_inplace / sub_plain_inplace / multiply_plain_inplace: mirror == oracle == reference, at the top level and one
    level down, BGV with a correction factor other than one, three-polynomial ciphertexts for the product."""
    ps = param_set(scheme)
    ctx = make_context(ps)
    o, oc = H.oracle(), ps.octx()
    n, t = ps.n, ps.t
    scale = float(2 ** 30) if scheme == 3 else 1.0   # BFV / BGV plaintexts carry scale 1 (are_same_scale)
    r, h = reference_for(ps, scale)
    rng = np.random.default_rng(scheme)
    try:
        for chain_index, size, cf in ((1, 2, 1), (2, 3, 3 if scheme == 1 else 1)):
            l = ps.size_Q - (chain_index - 1)
            words = np.stack([np.stack([rng.integers(0, int(ps.primes[i]), n, dtype=np.uint64) for i in range(l)])
                              for _ in range(size)])
            if scheme == 3:
                plain = np.stack([rng.integers(0, int(ps.primes[i]), n, dtype=np.uint64) for i in range(l)])
            else:
                plain = rng.integers(0, t, n).astype(np.uint64)
                plain[:4] = (0, t - 1, (t + 1) // 2, (t + 1) // 2 - 1)   # both sides of the upper-half threshold
            for op, fn in ((0, pf.add_plain_inplace), (1, pf.sub_plain_inplace), (2, pf.multiply_plain_inplace)):
                want = words.copy()
                if op == 2:
                    assert o.orc_plain_multiply(oc, l, P(want), size, P(plain)) == 0
                else:
                    assert o.orc_plain_add(oc, l, P(want), P(plain), op, cf) == 0
                ct = pf.PhantomCiphertext.from_host(ctx, words, chain_index=chain_index, scale=scale, is_ntt_form=(scheme != 2))
                ct.correction_factor = cf
                d_plain = torch.from_numpy(plain.view(np.int64)).cuda()
                if op == 2:
                    fn(ctx, ct, d_plain, plain_scale=scale)
                    assert ct.scale == scale * scale
                else:
                    fn(ctx, ct, d_plain)
                assert np.array_equal(host(ct.data), want), f"plain op {op} vs oracle, level {chain_index}"
                if h:
                    ref = np.zeros_like(words)
                    assert r.ref_plain_op(h, op, chain_index, P(words), size, P(plain), cf, P(ref)) == 0, r.ref_last_error()
                    assert np.array_equal(ref, want), f"oracle plain op {op} vs reference, level {chain_index}"
            ct = pf.PhantomCiphertext.from_host(ctx, words, chain_index=chain_index, scale=scale, is_ntt_form=(scheme != 2))
            out = pf.add_plain(ctx, ct, d_plain)   # copying form leaves the operand alone
            assert np.array_equal(host(ct.data), words) and not np.array_equal(host(out.data), words)
            if scheme == 3:
                with pytest.raises(ValueError):
                    pf.add_plain_inplace(ctx, ct, d_plain, plain_scale=scale * 2)
                with pytest.raises(ValueError):
                    pf.add_plain_inplace(ctx, ct, d_plain[:l - 1])
    finally:
        if h:
            r.ref_destroy(h)


@pytest.mark.parametrize("scheme", [2, 1, 3])
def test_add_sub_negate(scheme):
    """negate / add / sub / sub(negate) / add_many: mirror == reference word for word; BGV operands with different correction
    factors are balanced the way the reference balances them (same factor, same words)."""
    ps = param_set(scheme)
    ctx = make_context(ps)
    n, l = ps.n, ps.size_Q
    scale = float(2 ** 30) if scheme == 3 else 1.0   # BFV / BGV plaintexts carry scale 1 (are_same_scale)
    r, h = reference_for(ps, scale)
    rng = np.random.default_rng(10 + scheme)
    ntt = scheme != 2

    def rand_ct(size=2):
        return np.stack([np.stack([rng.integers(0, int(ps.primes[i]), n, dtype=np.uint64) for i in range(l)])
                         for _ in range(size)])

    def plain_python(a, b, op):
        out = np.zeros_like(a)
        for k in range(a.shape[0]):
            for i in range(l):
                q = int(ps.primes[i])
                x, y = a[k, i].astype(object), b[k, i].astype(object)
                out[k, i] = {0: (x + y) % q, 1: (x - y) % q, 2: (y - x) % q, 3: (-x) % q}[op].astype(np.uint64)
        return out

    try:
        for cf1, cf2 in ((1, 1),) + (((1, 3), (5, 12345)) if scheme == 1 else ()):
            a, b = rand_ct(), rand_ct()
            for op in (0, 1, 2, 3):
                c1 = pf.PhantomCiphertext.from_host(ctx, a, scale=scale, is_ntt_form=ntt)
                c2 = pf.PhantomCiphertext.from_host(ctx, b, scale=scale, is_ntt_form=ntt)
                c1.correction_factor, c2.correction_factor = cf1, cf2
                if op == 0:
                    pf.add_inplace(ctx, c1, c2)
                elif op == 3:
                    pf.negate_inplace(ctx, c1)
                else:
                    pf.sub_inplace(ctx, c1, c2, negate=(op == 2))
                assert np.array_equal(host(c2.data), b), "the second operand is left alone"
                if cf1 == cf2:
                    assert np.array_equal(host(c1.data), plain_python(a, b, op)), f"op {op} vs integers"
                if h:
                    ref, cf_out = np.zeros_like(a), ctypes.c_uint64(0)
                    assert r.ref_add_sub(h, op, 1, P(a), P(b), 2, cf1, cf2, P(ref), ctypes.byref(cf_out)) == 0, r.ref_last_error()
                    assert np.array_equal(host(c1.data), ref), f"op {op} vs reference, factors {cf1}, {cf2}"
                    assert c1.correction_factor == cf_out.value, "balanced correction factor"
        # add_many = repeated add; copying forms; refusals
        cts = [pf.PhantomCiphertext.from_host(ctx, rand_ct(), scale=scale, is_ntt_form=ntt) for _ in range(3)]
        total = pf.add_many(ctx, cts)
        want = pf.add(ctx, pf.add(ctx, cts[0], cts[1]), cts[2])
        assert np.array_equal(host(total.data), host(want.data))
        other = pf.PhantomCiphertext.from_host(ctx, rand_ct(3), scale=scale, is_ntt_form=ntt)
        with pytest.raises(ValueError, match="poly number mismatch"):
            pf.add_inplace(ctx, cts[0], other)
        if scheme == 3:
            off = cts[1].clone()
            off.scale = scale * 2
            with pytest.raises(ValueError, match="scale mismatch"):
                pf.sub_inplace(ctx, cts[0], off)
        low = pf.mod_switch_to(ctx, cts[0], 3)
        assert low.chain_index == 3 and low.coeff_modulus_size() == l - 2 and cts[0].chain_index == 1
        step = pf.mod_switch_to_next(ctx, pf.mod_switch_to_next(ctx, cts[0]))
        assert np.array_equal(host(low.data), host(step.data))
        with pytest.raises(ValueError, match="parameter mismatch"):
            pf.add_inplace(ctx, cts[1], low)
        with pytest.raises(ValueError, match="higher level"):
            pf.mod_switch_to(ctx, low, 1)
    finally:
        if h:
            r.ref_destroy(h)
