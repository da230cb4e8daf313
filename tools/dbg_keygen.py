"""Diagnostic: gen_relinkey (engine) against the oracle over many random seed sets; reports differing words."""
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
ps = T.param_set(2, 8192)
ctx = T.make_context(ps)
o, oc = H.oracle(), ps.octx()
n, l, m = ps.n, ps.size_Q, ps.size_QP
dnum = l // ps.size_P
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(7)
total = 0
for rep in range(reps):
    sd = T.seeds_of(rng, 1)
    sk = pf.PhantomSecretKey(ctx, seed=sd[0])
    want_sk = np.zeros((m, n), dtype=np.uint64)
    o.orc_gen_secretkey(oc, sd[0], P(want_sk))
    assert np.array_equal(T.host(sk.secret_key_array())[0], want_sk)
    kseeds = b"".join(T.seeds_of(rng, 2 * dnum))
    rlk = sk.gen_relinkey(ctx, seeds=kseeds)
    kc = o.orc_create(ps.scheme, ps.n, P(ps.primes), m, 0, ps.t)
    sk2 = np.zeros_like(want_sk)
    o.orc_poly_mul(kc, P(want_sk), P(want_sk), P(sk2), m)
    o.orc_destroy(kc)
    got_sk2 = T.host(sk.secret_key_array())[1]
    bad2 = np.argwhere(got_sk2 != sk2)
    if len(bad2):
        i, x = bad2[0]
        print(f"rep {rep}: s^2 differs in {len(bad2)} words, first limb {i} idx {x}: engine {got_sk2[i, x]} oracle {sk2[i, x]} q {ps.primes[i]}")
    want_rlk = np.zeros((dnum, 2, m, n), dtype=np.uint64)
    assert o.orc_gen_kswitch_key(oc, P(sk2), P(want_sk), kseeds, P(want_rlk)) == 0
    for d in range(dnum):
        got = T.host(rlk.digits[d])
        bad = np.argwhere(got != want_rlk[d])
        if len(bad):
            total += len(bad)
            k, i, x = bad[0]
            print(f"rep {rep} digit {d}: {len(bad)} words differ, first poly {k} limb {i} idx {x}: engine {got[k, i, x]} oracle {want_rlk[d][k, i, x]} q {ps.primes[i]}")
print(f"{total} differing key words over {reps} keys")
