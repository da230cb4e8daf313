"""Diagnostic: the reference's BFV multiply + relinearize + decrypt with (a) its own relinearisation key and (b) a key generated
by the engine from fresh random seeds, repeated; reports how many coefficients decrypt wrongly."""
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
import test_keygen_gpu as T  # noqa: E402

T.setup_module(T)
r = H.reference()
ps = T.param_set(2, 8192)
n, l, m, t = ps.n, ps.size_Q, ps.size_QP, ps.t
print("primes bits", [int(p).bit_length() for p in ps.primes], "size_P", ps.size_P, "t", t)
for use_engine_key in ((0, 1) if len(sys.argv) < 3 else (int(sys.argv[2]),)):
    bad_runs = 0
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
        h = r.ref_create(2, n, P(ps.primes), m, ps.size_P, t, 2, None, 0, 1.0, 1)
        ctx = T.make_context(ps)
        s1 = np.zeros((m, n), dtype=np.uint64)
        assert r.ref_secret_key(h, P(s1)) == 0
        if use_engine_key:
            sk = pf.PhantomSecretKey(ctx, s1)
            rlk = sk.gen_relinkey(ctx)
            for d, digit in enumerate(rlk.digits):
                assert r.ref_key_set(h, -1, d, P(T.host(digit))) == 0
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
        if len(bad):
            bad_runs += 1
            print(f"  engine_key={use_engine_key} rep {rep}: {len(bad)} wrong coefficients, first {bad[:6]}, values {dec[bad[:6]]}")
        r.ref_destroy(h)
        del ctx
    print(f"engine_key={use_engine_key}: {bad_runs} runs with wrong coefficients")
