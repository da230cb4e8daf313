"""Generates tests/golden/host_tables.json from the reference's OWN host code (oracle/_ref/libphantom_ref.so built
from /root/reference by oracle/Makefile.ref).  Run in the build container (no GPU needed):

    python tests/golden/make_golden.py

The fixture pins the oracle's (and the engine's) table generator: prime chains of every BASELINE.json config,
minimal 2N-th roots, n^-1, Barrett ratios, twiddle samples + digests, base-conversion matrices, Galois elements."""
import ctypes
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
from harness import P  # noqa: E402

r = H.reference()
assert r is not None, "build oracle/_ref first (make -C oracle -f Makefile.ref)"

CONFIGS = {
    "C1_ntt_2^12": (4096, [50]),
    "primary_2^16_L16_a4": (65536, [60] + [40] * 15 + [60] * 4),
    "secondary_2^16_L16_a1": (65536, [60] + [40] * 15 + [60]),
    "bfv_2^14_438bits": (16384, [54] * 7 + [60]),
    "small_2^12": (4096, [50, 40, 40, 40, 40, 50, 50]),
    "mixed_2^13": (8192, [55, 36, 36, 60, 30]),
}

out = {"configs": {}, "tables": {}, "bconv": [], "galois": {}}
for name, (n, bits) in CONFIGS.items():
    arr = (ctypes.c_int * len(bits))(*bits)
    primes = np.zeros(len(bits), dtype=np.uint64)
    assert r.ref_host_create_primes(n, arr, len(bits), P(primes)) == 0, r.ref_last_error()
    out["configs"][name] = {"n": n, "bits": bits, "primes": [int(p) for p in primes]}
    logn = n.bit_length() - 1
    for q in sorted(set(int(p) for p in primes))[:3] + sorted(set(int(p) for p in primes))[-2:]:
        key = f"{logn}:{q}"
        if key in out["tables"]:
            continue
        tw = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
        misc = np.zeros(6, dtype=np.uint64)
        assert r.ref_host_ntt_table(logn, q, *[P(t) for t in tw], P(misc)) == 0, r.ref_last_error()
        out["tables"][key] = {
            "root": int(misc[0]), "n_inv": int(misc[1]), "n_inv_shoup": int(misc[2]),
            "ratio": [int(misc[3]), int(misc[4]), int(misc[5])],
            "tw_head": [int(v) for v in tw[0][:8]], "tws_head": [int(v) for v in tw[1][:8]],
            "itw_head": [int(v) for v in tw[2][:8]], "itws_head": [int(v) for v in tw[3][:8]],
            "sha256": [hashlib.sha256(t.tobytes()).hexdigest() for t in tw],
        }

# base-conversion matrices: the mod-up digits and the P -> Ql converter of the primary set
prim = out["configs"]["primary_2^16_L16_a4"]["primes"]
Q, Pp = prim[:16], prim[16:]
cases = [(Q[0:4], Q[4:16] + Pp), (Q[12:16], Q[0:12] + Pp), (Pp, Q), (Q[0:2], Q[2:5] + Pp), ([Q[3]], Q[0:3] + Pp)]
for ib, ob in cases:
    ia, oa = np.array(ib, dtype=np.uint64), np.array(ob, dtype=np.uint64)
    mat = np.zeros(len(ib) * len(ob), dtype=np.uint64)
    hinv = np.zeros(len(ib), dtype=np.uint64)
    assert r.ref_host_bconv_tables(P(ia), len(ib), P(oa), len(ob), P(mat), P(hinv)) == 0, r.ref_last_error()
    out["bconv"].append({"ibase": ib, "obase": ob, "qhat_mod_p": [int(v) for v in mat],
                         "qhatinv_mod_q": [int(v) for v in hinv]})

for n in (4096, 65536):
    out["galois"][str(n)] = {str(s): int(r.ref_host_galois_elt(s, n)) for s in (0, 1, 2, 3, 7, 32, -1, -5, 100)}

with open(os.path.join(HERE, "host_tables.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote host_tables.json:", {k: len(v) for k, v in out.items()})
