"""Diagnostic: reference secret key -> engine relin key (seeds known) -> reference multiply + relinearize + decrypt; on a wrong
coefficient compare the engine's key with the oracle's for the same seeds and check the key's noise directly."""
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
import test_keygen_gpu as T  # noqa: E402

T.setup_module(T)
r = H.reference()
ps = T.param_set(2, 8192)
n, l, m, t = ps.n, ps.size_Q, ps.size_QP, ps.t
o, oc = H.oracle(), ps.octx()
dnum = l // ps.size_P
rng = np.random.default_rng(11)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for rep in range(reps):
    h = r.ref_create(2, n, P(ps.primes), m, ps.size_P, t, 2, None, 0, 1.0, 1)
    ctx = T.make_context(ps)
    s1 = np.zeros((m, n), dtype=np.uint64)
    assert r.ref_secret_key(h, P(s1)) == 0
    sk = pf.PhantomSecretKey(ctx, s1)
    kseeds = b"".join(T.seeds_of(rng, 2 * dnum))
    rlk = sk.gen_relinkey(ctx, seeds=kseeds)
    digs = [T.host(d).copy() for d in rlk.digits]
    for d, digit in enumerate(digs):
        assert r.ref_key_set(h, -1, d, P(digit)) == 0
    a = np.zeros(n, dtype=np.uint64); a[0], a[1] = 3, 5
    b = np.zeros(n, dtype=np.uint64); b[0], b[n - 1] = 7, 2
    ca, cb = np.zeros((2, l, n), dtype=np.uint64), np.zeros((2, l, n), dtype=np.uint64)
    assert r.ref_encrypt(h, 0, 1, P(a), P(ca)) == 0 and r.ref_encrypt(h, 1, 1, P(b), P(cb)) == 0
    prod = np.zeros((2, l, n), dtype=np.uint64)
    assert r.ref_multiply_relin(h, 1, P(ca), P(cb), P(prod)) == 0
    dec = np.zeros(n, dtype=np.uint64)
    assert r.ref_decrypt(h, 1, P(prod), 2, 1, P(dec)) == 0
    want = np.zeros(n, dtype=np.uint64); want[0], want[1], want[n - 1] = 11, 35, 6
    bad = np.nonzero(dec % t != want)[0]
    # the engine's product on the same ciphertexts and key
    cae = pf.PhantomCiphertext.from_host(ctx, ca, is_ntt_form=False)
    cbe = pf.PhantomCiphertext.from_host(ctx, cb, is_ntt_form=False)
    pf.multiply_inplace(ctx, cae, cbe)
    pf.relinearize_inplace(ctx, cae, rlk)
    same_words = np.array_equal(cae.to_host(), prod)
    dec_e = T.host(sk.decrypt(ctx, cae))
    bad_e = np.nonzero(dec_e % t != want)[0]
    # oracle key for the same seeds
    kc = o.orc_create(ps.scheme, ps.n, P(ps.primes), m, 0, ps.t)
    sk2 = np.zeros_like(s1)
    o.orc_poly_mul(kc, P(s1), P(s1), P(sk2), m)
    o.orc_destroy(kc)
    want_rlk = np.zeros((dnum, 2, m, n), dtype=np.uint64)
    assert o.orc_gen_kswitch_key(oc, P(sk2), P(s1), kseeds, P(want_rlk)) == 0
    key_diff = sum(int(np.count_nonzero(digs[d] != want_rlk[d])) for d in range(dnum))
    print(f"rep {rep}: reference wrong coeffs {list(bad[:4])}, engine-product words equal reference {same_words}, engine decrypt wrong {list(bad_e[:4])}, key words != oracle {key_diff}", flush=True)
    r.ref_destroy(h)
    del ctx
