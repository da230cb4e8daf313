"""Per-CTA phase timeline of the forward NTT passes (debug library built by `make -C phantom-fhe_b200/csrc timeline`).

PFHE_B200_LIB=phantom-fhe_b200/libpfhe_b200_tl.so python tools/timeline.py [limbs]
"""
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

limbs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
ctx = pf.PhantomContext(parms)
n = ps.n
x = torch.zeros((limbs, n), dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
npoly, per = limbs // 16, 16


def run():
    pf.check(pf.lib.pfhe_ntt_forward_inplace_batch(ctx._h, x.data_ptr(), npoly, per, 0, st))


for _ in range(5):
    run()
torch.cuda.synchronize()
run()
torch.cuda.synchronize()
words = 2 * 65536 * 16
buf = np.zeros(words, dtype=np.int64)
dll = ctypes.CDLL(pf.LIB_PATH)
dll.pfhe_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
assert dll.pfhe_debug_timeline(buf.ctypes.data, words) == 0
tiles = n // 2048
for rows, name in ((0, "k_fwd_cols"), (1, "k_fwd_rows")):
    t = buf[rows * 65536 * 16:(rows * 65536 + limbs * tiles) * 16].reshape(-1, 16)
    g0 = t[:, 0].min()
    start_us = (t[:, 0] - g0) / 1000.0
    d = lambda a, b: (t[:, b] - t[:, a]).astype(np.float64)
    life = d(2, 12)
    print(f"== {name}: {len(t)} CTAs, kernel span (first CTA start -> last CTA start) {start_us.max():.1f} us")
    print(f"   CTA lifetime cycles: median {np.median(life):.0f}  p10 {np.percentile(life, 10):.0f}  p90 {np.percentile(life, 90):.0f}")
    phases = [("load wait", 2, 3), ("round0 compute", 3, 4), ("round0 exchange", 4, 5), ("round1 load+compute", 5, 6),
              ("round1 exchange", 6, 7), ("round2 load+compute", 7, 8), ("round2 exchange/store", 8, 9), ("final store", 9, 12)]
    for nm, a, b in phases:
        v = d(a, b)
        print(f"   {nm:24s} median {np.median(v):7.0f} cyc  ({100 * np.median(v) / np.median(life):4.1f}% of lifetime)")
    # by arithmetic kind: limbs 0 and 16.. are integer? primary chain: row 0 and rows >= 16 are 60-bit
    for kind, sel in (("FP64 limbs", [i for i in range(limbs) if (i % per) != 0]), ("int limbs", [i for i in range(limbs) if (i % per) == 0])):
        idx = np.concatenate([np.arange(i * tiles, (i + 1) * tiles) for i in sel])
        print(f"   {kind}: lifetime median {np.median(life[idx]):.0f} cycles")
    # CTAs per SM over time
    sm = t[:, 1]
    print(f"   SMs used {len(np.unique(sm))}, CTAs per SM median {np.median(np.bincount(sm.astype(int))[np.unique(sm).astype(int)]):.0f}")
    order = np.argsort(start_us)
    print("   start times (us) of CTA #0, #592, #1184, #1776:", [round(float(start_us[order[min(i, len(order) - 1)]]), 1) for i in (0, 592, 1184, 1776)])
