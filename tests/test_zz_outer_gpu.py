"""GPU tests of the outer layers of the host mirror: the remaining stream formats -- public keys (include/secretkey.h:85-96)
and plaintexts (include/plaintext.h:69-97); a public key written by the unmodified reference is loaded here and used to
encrypt, the reference decrypts the result -- the Galois permutations on their own, and scripts written against the names
of the reference's Python binding (`import pyPhantom as phantom`, python/src/binding.cu)."""
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


def test_pyphantom_surface_ckks():
    """A script in the shape of the reference's python/examples/ckks.py, written against `import pyPhantom as phantom`:
    keys, encode, public-key encryption, multiply_and_relin, rescale, hoisting over seven steps, add, decrypt, decode."""
    import pyPhantom as phantom
    n, scale = 8192, 2.0 ** 40
    steps = [1, 2, 3, 4, 5, 6, 7]
    parms = phantom.params(phantom.scheme_type.ckks)
    parms.set_poly_modulus_degree(n)
    parms.set_coeff_modulus(phantom.create_coeff_modulus(n, [60, 40, 40, 60]))
    parms.set_special_modulus_size(1)
    parms.set_galois_elts(phantom.get_elts_from_steps(steps, n))
    ctx = phantom.context(parms)
    sk = phantom.secret_key(ctx)
    pk, rlk, glk = sk.gen_publickey(ctx), sk.gen_relinkey(ctx), sk.create_galois_keys(ctx)
    enc = phantom.ckks_encoder(ctx)
    slots = enc.slot_count()
    msg = np.zeros(slots)
    msg[:8] = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0]
    pt = enc.encode_double_vector(ctx, list(msg), scale, chain_index=1)
    ct = pk.encrypt_asymmetric(ctx, pt)
    ct = phantom.multiply_and_relin(ctx, ct, ct, rlk)
    ct = phantom.rescale_to_next(ctx, ct)
    ct2 = phantom.hoisting(ctx, ct, glk, steps)
    ct = phantom.add(ctx, ct, ct2)
    got = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, ct)))
    sq = msg * msg
    want = sq + sum(np.roll(sq, -k) for k in steps)
    assert np.max(np.abs(got - want)) < 1e-3, np.max(np.abs(got - want))
    # plaintext operands and level switching through the binding's names
    half = enc.encode_double_vector(ctx, [0.5] * slots, scale, chain_index=1)
    c2 = phantom.multiply_plain(ctx, pk.encrypt_asymmetric(ctx, pt), half)
    c2 = phantom.rescale_to_next(ctx, c2)
    got = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, c2)))
    assert np.max(np.abs(got - 0.5 * msg)) < 1e-4
    low = phantom.mod_switch_to_next(ctx, pt)
    assert low.chain_index() == 2 and low.data.shape[0] == 2
    c3 = phantom.add_plain(ctx, phantom.mod_switch_to(ctx, sk.encrypt_symmetric(ctx, pt), 2), low)
    got = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, c3)))
    assert np.max(np.abs(got - 2 * msg)) < 1e-4
    with pytest.raises(ValueError):
        phantom.add_plain(ctx, c3, pt)   # plaintext at another level


@pytest.mark.parametrize("scheme", ["bfv", "bgv"])
def test_pyphantom_surface_integer_schemes(scheme):
    """The shape of python/examples/bfv.py / bgv.py: batch encoding, public-key encryption, multiply_and_relin (BFV with
    mul_tech hps_overq_leveled), a rotation by one step, decrypt, decode."""
    import pyPhantom as phantom
    n = 8192
    parms = phantom.params(getattr(phantom.scheme_type, scheme))
    parms.set_poly_modulus_degree(n)
    parms.set_coeff_modulus(phantom.create_coeff_modulus(n, [50, 50, 50, 50, 60, 60]))
    parms.set_plain_modulus(phantom.create_plain_modulus(n, 20))
    parms.set_special_modulus_size(2)
    parms.set_galois_elts(phantom.get_elts_from_steps([1], n))
    if scheme == "bfv":
        parms.set_mul_tech(phantom.mul_tech_type.hps_overq_leveled)
    ctx = phantom.context(parms)
    t = parms.plain_modulus
    sk = phantom.secret_key(ctx)
    pk, rlk, glk = sk.gen_publickey(ctx), sk.gen_relinkey(ctx), sk.create_galois_keys(ctx)
    enc = phantom.batch_encoder(ctx)
    assert enc.slot_count() == n
    rng = np.random.default_rng(3)
    msg = [int(v) for v in rng.integers(0, 1000, n)]
    pt = enc.encode(ctx, msg)
    ct = pk.encrypt_asymmetric(ctx, pt)
    assert [v % t for v in enc.decode(ctx, sk.decrypt(ctx, ct))] == msg
    ct = phantom.multiply_and_relin(ctx, ct, ct, rlk)
    sq = [v * v % t for v in msg]
    assert [v % t for v in enc.decode(ctx, sk.decrypt(ctx, ct))] == sq
    ct = phantom.rotate(ctx, ct, 1, glk)
    half = n // 2
    rot = [sq[(i + 1) % half] for i in range(half)] + [sq[half + (i + 1) % half] for i in range(half)]
    assert [v % t for v in enc.decode(ctx, sk.decrypt(ctx, ct))] == rot
    # linear surface
    two = phantom.add(ctx, sk.encrypt_symmetric(ctx, pt), pk.encrypt_asymmetric(ctx, pt))
    assert [v % t for v in enc.decode(ctx, sk.decrypt(ctx, two))] == [2 * v % t for v in msg]
    three = phantom.add_plain(ctx, two, pt)
    assert [v % t for v in enc.decode(ctx, sk.decrypt(ctx, three))] == [3 * v % t for v in msg]
    prod = phantom.multiply_plain(ctx, sk.encrypt_symmetric(ctx, pt), pt)
    assert [v % t for v in enc.decode(ctx, sk.decrypt(ctx, prod))] == sq
    zero = phantom.sub(ctx, two, phantom.add_many(ctx, [sk.encrypt_symmetric(ctx, pt), sk.encrypt_symmetric(ctx, pt)]))
    assert [v % t for v in enc.decode(ctx, sk.decrypt(ctx, zero))] == [0] * n


def test_rotate_under_leveled_mul_tech_against_reference():
    """rotate_inplace for BFV with mul_tech hps_overq_leveled: the reference's key switch drops levels for rotations too
    (keyswitch_inplace with is_relin = false, eval_key_switch.cu:111-174).  Same words as the unmodified reference, with the
    reference's own Galois key."""
    r = H.reference()
    if r is None:
        pytest.skip("oracle/_ref/libphantom_ref.so was not built")
    ps = H.params_bfv_bench(0)
    steps = (ctypes.c_int * 1)(1)
    h = r.ref_create(2, ps.n, P(ps.primes), ps.size_QP, ps.size_P, ps.t, 4, steps, 1, 1.0, 1)
    assert h, r.ref_last_error()
    try:
        n, l, m = ps.n, ps.size_Q, ps.size_QP
        dnum = r.ref_dnum(h)
        glk_h = np.zeros((dnum, 2, m, n), dtype=np.uint64)
        for d in range(dnum):
            assert r.ref_key_get(h, 0, d, P(glk_h[d])) == 0
        parms = pf.EncryptionParameters(pf.scheme_type.bfv)
        parms.set_poly_modulus_degree(n)
        parms.set_coeff_modulus([int(p) for p in ps.primes])
        parms.set_special_modulus_size(ps.size_P)
        parms.set_plain_modulus(ps.t)
        parms.set_mul_tech(pf.mul_tech_type.hps_overq_leveled)
        parms.set_galois_elts(pf.get_elts_from_steps([1], n))
        ctx = pf.PhantomContext(parms)
        glk = pf.PhantomGaloisKey(ctx, [list(glk_h)])
        a = H.ciphertext(ps, 21)
        want = np.zeros((2, l, n), dtype=np.uint64)
        assert r.ref_rotate(h, 1, P(a), 1, P(want)) == 0, r.ref_last_error()
        c = pf.PhantomCiphertext.from_host(ctx, a, is_ntt_form=False)
        pf.rotate_inplace(ctx, c, 1, glk)
        assert np.array_equal(c.to_host(), want), "rotate under hps_overq_leveled vs reference"
    finally:
        r.ref_destroy(h)


def test_cpp_mirror_application():
    """tests/cpp/mirror_demo.cpp: an application written against the reference's C++ names (PhantomContext,
    PhantomSecretKey, multiply_and_relin_inplace, rotate_inplace, rescale_to_next, ...) on include/phantom_b200.hpp --
    BFV (hps_overq_leveled), BGV and CKKS flows; every decrypted result must be the expected one."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "mirror_demo")
    if not os.path.exists(exe):
        if shutil.which("nvcc") is None:
            pytest.skip("mirror_demo was not built and there is no nvcc")
        lib_dir = os.path.join(root, "phantom-fhe_b200")
        subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-I", os.path.join(root, "include"),
                               os.path.join(root, "tests", "cpp", "mirror_demo.cpp"), "-o", exe, "-L", lib_dir, "-lpfhe_b200",
                               "-Xlinker", "-rpath", "-Xlinker", lib_dir, "-Wno-deprecated-gpu-targets"])
    if not os.access(exe, os.X_OK):
        os.chmod(exe, 0o755)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
