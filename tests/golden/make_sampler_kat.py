"""Known-answer vectors of the reference's device generator and samplers (src/prng.cu), written to sampler_kat.json.

The values are produced by the CPU oracle; tests/test_keygen_gpu.py::test_samplers_against_reference_kernels shows the
oracle equal, word for word, to the reference's own sample_*_poly kernels on a B200 for the same seeds (the kernels cannot
run in the GPU-less container, so the fixture is generated through the oracle and guards it against regressions).

    python tests/golden/make_sampler_kat.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
from harness import P  # noqa: E402


def main():
    o = H.oracle()
    n = 4096
    ps = H.ParamSet("kat", n, [60, 40, 60], 1, 3, 0)
    oc, m = ps.octx(), ps.size_QP
    cases = []
    for name, seed in (("counting", bytes(range(64))), ("zeros", bytes(64)), ("ones", bytes([255] * 64))):
        for kind in (0, 1, 2):
            out = np.zeros((m, n), dtype=np.uint64)
            assert o.orc_sample_poly(oc, kind, m, seed, P(out)) == 0
            cases.append(dict(seed=name, kind=kind, head=[[int(v) for v in out[i, :8]] for i in range(m)],
                              sha256=hashlib.sha256(out.tobytes()).hexdigest()))
    json.dump(dict(n=n, bits=[60, 40, 60], primes=[int(p) for p in ps.primes], cases=cases),
              open(os.path.join(HERE, "sampler_kat.json"), "w"), indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
