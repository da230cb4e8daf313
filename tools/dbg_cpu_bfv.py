"""Diagnostic, CPU only: the oracle's BFV (HPS) key generation, encryption, multiply + relinearize and decrypt from random
seeds, repeated -- looks for the rare wrong decryptions seen with the reference at N = 8192, t = 65537 (cause: DESIGN.md
section 6, "alpha re-reduced limb after limb").  PFHE_WRITE_CASE=<path> writes the seeds of the first run with exactly one
wrong coefficient: tests/golden/bfv_hps_alpha_case.json was made by
    PFHE_WRITE_CASE=tests/golden/bfv_hps_alpha_case.json python tools/dbg_cpu_bfv.py 400"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from harness import P  # noqa: E402

bits = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else "40,40,40,50".split(","))]
n = 8192
ps = H.ParamSet("dbg", n, bits, 1, scheme=2, t=65537)
o, oc = H.oracle(), ps.octx()
l, m, t = ps.size_Q, ps.size_QP, ps.t
dnum = l // ps.size_P
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
mul_tech = int(sys.argv[4]) if len(sys.argv) > 4 else 2   # 1 behz, 2 hps, 3 hps_overq
a = np.zeros(n, dtype=np.uint64); a[0], a[1] = 3, 5
b = np.zeros(n, dtype=np.uint64); b[0], b[n - 1] = 7, 2
want = np.zeros(n, dtype=np.uint64); want[0], want[1], want[n - 1] = 11, 35, 6
kc = o.orc_create(ps.scheme, ps.n, P(ps.primes), m, 0, ps.t)
bad_runs = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    sd = [rng.bytes(64) for _ in range(8)]
    sk = np.zeros((m, n), dtype=np.uint64)
    o.orc_gen_secretkey(oc, sd[0], P(sk))
    pk = np.zeros((2, m, n), dtype=np.uint64)
    assert o.orc_encrypt_zero_symmetric(oc, 0, P(sk), sd[1], sd[2], P(pk)) == 0
    sk2 = np.zeros_like(sk)
    o.orc_poly_mul(kc, P(sk), P(sk), P(sk2), m)
    kseeds = b"".join(rng.bytes(64) for _ in range(2 * dnum))
    rlk = np.zeros((dnum, 2, m, n), dtype=np.uint64)
    assert o.orc_gen_kswitch_key(oc, P(sk2), P(sk), kseeds, P(rlk)) == 0
    ca, cb = np.zeros((2, l, n), dtype=np.uint64), np.zeros((2, l, n), dtype=np.uint64)
    assert o.orc_encrypt_zero_symmetric(oc, 1, P(sk), sd[3], sd[4], P(ca)) == 0
    assert o.orc_encrypt_add_plain(oc, l, P(ca), P(a)) == 0
    assert o.orc_encrypt_zero_asymmetric(oc, P(pk), sd[5], sd[6], P(cb)) == 0
    assert o.orc_encrypt_add_plain(oc, l, P(cb), P(b)) == 0
    prod = np.zeros((2, l, n), dtype=np.uint64)
    if mul_tech == 2:
        assert o.orc_bfv_multiply_relin_hps(oc, P(ca), P(cb), P(rlk), P(prod)) == 0
    elif mul_tech == 3:
        assert o.orc_bfv_multiply_relin_hps_overq(oc, P(ca), P(cb), P(rlk), P(prod), 0) == 0
    else:
        assert o.orc_bfv_multiply_relin_behz(oc, P(ca), P(cb), P(rlk), P(prod)) == 0
    dec = np.zeros(n, dtype=np.uint64)
    assert o.orc_decrypt(oc, l, P(prod), 2, P(sk), mul_tech, 1, P(dec)) == 0
    bad = np.nonzero(dec % t != want)[0]
    if len(bad):
        bad_runs += 1
        print(f"rep {rep}: {len(bad)} wrong, first {bad[:6]}, values {dec[bad[:6]]}", flush=True)
        if len(bad) == 1 and os.environ.get("PFHE_WRITE_CASE"):
            import json
            json.dump({"n": n, "prime_bits": bits, "t": t, "seeds": [v.hex() for v in sd], "kswitch_seeds": kseeds.hex(),
                       "wrong_coefficient": int(bad[0]), "decrypts_to": int(dec[bad[0]] % t), "generator": "tools/dbg_cpu_bfv.py"},
                      open(os.environ["PFHE_WRITE_CASE"], "w"), indent=1)
            print("case written"), sys.exit(0)
print(f"{bad_runs} bad runs")
