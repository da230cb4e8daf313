"""Diagnostic: pfhe_sample_poly against the oracle over many random seeds (rare-event paths of the samplers)."""
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
ps = T.param_set(2, 8192)
ctx = T.make_context(ps)
o, oc = H.oracle(), ps.octx()
n, m = ps.n, ps.size_QP
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(1)
for kind, name in ((0, "ternary"), (1, "error"), (2, "uniform")):
    bad_total = 0
    for rep in range(reps):
        seed = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        d = torch.zeros((m, n), dtype=torch.int64, device="cuda")
        pf.check(pf.lib.pfhe_sample_poly(ctx._h, kind, m, seed, d.data_ptr(), st))
        got = T.host(d)
        want = np.zeros((m, n), dtype=np.uint64)
        o.orc_sample_poly(oc, kind, m, seed, P(want))
        bad = np.argwhere(got != want)
        if len(bad):
            bad_total += len(bad)
            i, x = bad[0]
            print(f"  {name} rep {rep}: {len(bad)} words differ, first limb {i} coeff {x}: engine {got[i, x]} oracle {want[i, x]} q {ps.primes[i]}")
    print(f"{name}: {bad_total} differing words over {reps} seeds x {m} x {n}")
