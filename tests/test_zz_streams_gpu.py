"""GPU tests of the remaining stream formats of the host mirror: public keys (include/secretkey.h:85-96) and plaintexts
(include/plaintext.h:69-97).  A public key written by the unmodified reference is loaded here and used to encrypt; the
reference decrypts the result."""
import ctypes
import io

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


@pytest.mark.parametrize("scheme", [2, 1])
def test_public_key_and_plaintext_streams(scheme):
    ps = H.params_small(4096, l=3, alpha=1, qbits=36, pbits=42, scheme=scheme, t=65537)
    ctx = make_context(ps)
    n, l, m, t = ps.n, ps.size_Q, ps.size_QP, ps.t
    rng = np.random.default_rng(scheme)
    plain = torch.from_numpy(rng.integers(0, t, n).astype(np.uint64).view(np.int64)).cuda()
    # round trips inside the mirror
    sk = pf.PhantomSecretKey(ctx)
    pk = sk.gen_publickey(ctx)
    buf = io.BytesIO()
    pk.save(buf)
    pk2 = pf.PhantomPublicKey.load(ctx, io.BytesIO(buf.getvalue()))
    assert np.array_equal(host(pk2.pk), host(pk.pk))
    ct = pk2.encrypt_asymmetric(ctx, plain)
    assert np.array_equal(host(sk.decrypt(ctx, ct)) % t, host(plain))
    buf = io.BytesIO()
    pf.save_plaintext(buf, plain)
    back, ci, scale = pf.load_plaintext(ctx, io.BytesIO(buf.getvalue()))
    assert np.array_equal(host(back), host(plain)) and ci == 0 and scale == 1.0
    # a public key stream written by the reference
    r = H.reference()
    if r is None or not hasattr(r, "ref_public_key_stream"):
        return
    h = r.ref_create(scheme, n, P(ps.primes), m, ps.size_P, t, 2, None, 0, 1.0, 1)
    assert h, r.ref_last_error()
    try:
        cap = 2 * m * n * 8 + 256
        raw = ctypes.create_string_buffer(cap)
        length = r.ref_public_key_stream(h, raw, cap)
        assert length == 58 + 2 * m * n * 8, r.ref_last_error()
        theirs = pf.PhantomPublicKey.load(ctx, io.BytesIO(raw.raw[:length]))
        out = io.BytesIO()
        theirs.save(out)
        assert out.getvalue() == raw.raw[:length], "public key stream re-written byte for byte"
        ct = theirs.encrypt_asymmetric(ctx, plain)
        dec = np.zeros(n, dtype=np.uint64)
        assert r.ref_decrypt(h, 1, P(host(ct.data)), 2, 1, P(dec)) == 0, r.ref_last_error()
        assert np.array_equal(dec % t, host(plain)), "reference decrypts a ciphertext made under its own public key"
    finally:
        r.ref_destroy(h)


def test_standalone_galois_permutations():
    """pfhe_apply_galois_ntt (PhantomGaloisTool::apply_galois_ntt, src/galois.cu:86-102) and pfhe_apply_galois (coefficient
    form, src/galois.cu:20-39) on their own, against the oracle: every key limb, several elements; in-place calls and
    elements the context does not hold are refused."""
    ps = H.params_small(4096, l=3, alpha=1, scheme=3)
    parms = pf.EncryptionParameters(pf.scheme_type.ckks)
    parms.set_poly_modulus_degree(ps.n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(ps.size_P)
    steps = [1, -3, 0]
    parms.set_galois_elts(pf.get_elts_from_steps(steps, ps.n))
    ctx = pf.PhantomContext(parms)
    o, oc = H.oracle(), ps.octx()
    n, m = ps.n, ps.size_QP
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    x = H.uniform_limbs(ps, list(range(m)), 5)[0]
    d_x = torch.from_numpy(x.view(np.int64)).cuda()
    d_y = torch.zeros_like(d_x)
    for elt in pf.get_elts_from_steps(steps, n):
        tab = np.zeros(n, dtype=np.uint32)
        o.orc_galois_table(n, elt, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        want = np.zeros_like(x)
        o.orc_apply_galois_ntt(oc, P(x), P(want), m, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        pf.check(pf.lib.pfhe_apply_galois_ntt(ctx._h, d_x.data_ptr(), m, elt, d_y.data_ptr(), st))
        assert np.array_equal(host(d_y), want), f"NTT-form automorphism, element {elt}"
        o.orc_apply_galois_coeff(oc, P(x), P(want), m, elt)
        pf.check(pf.lib.pfhe_apply_galois(ctx._h, d_x.data_ptr(), m, elt, d_y.data_ptr(), st))
        assert np.array_equal(host(d_y), want), f"coefficient-form automorphism, element {elt}"
    want = np.zeros_like(x)
    o.orc_apply_galois_coeff(oc, P(x), P(want), 2, 3)   # any odd element in coefficient form, fewer limbs
    pf.check(pf.lib.pfhe_apply_galois(ctx._h, d_x.data_ptr(), 2, 3, d_y.data_ptr(), st))
    assert np.array_equal(host(d_y)[:2], want[:2])
    with pytest.raises(ValueError):
        pf.check(pf.lib.pfhe_apply_galois_ntt(ctx._h, d_x.data_ptr(), m, 3, d_y.data_ptr(), st))   # not a context element
    with pytest.raises(ValueError):
        pf.check(pf.lib.pfhe_apply_galois_ntt(ctx._h, d_x.data_ptr(), m, 5, d_x.data_ptr(), st))   # in place
    with pytest.raises(ValueError):
        pf.check(pf.lib.pfhe_apply_galois(ctx._h, d_x.data_ptr(), m, 4, d_y.data_ptr(), st))       # even element
