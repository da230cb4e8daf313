"""Worker of tests/test_substitution_gpu.py: runs the reference's own evaluate.cu through whatever library PFHE_REF_SO
names (the link-time substitution build oracle/_ref/libphantom_subst.so) on words and keys handed over in .npy files."""
import ctypes
import sys

import numpy as np

import harness as H
from harness import P

work = sys.argv[1]
r = H.reference()
assert r is not None, "library missing"
scheme, n, size_P, t, mul_tech = [int(v) for v in np.load(f"{work}/meta.npy")]
primes = np.load(f"{work}/primes.npy")
steps = (ctypes.c_int * 1)(1)
h = r.ref_create(scheme, n, P(primes), len(primes), size_P, t, mul_tech, steps, 1, float(2 ** 40), 1)
assert h, r.ref_last_error()
rlk, glk = np.load(f"{work}/rlk.npy"), np.load(f"{work}/glk.npy")
for d in range(rlk.shape[0]):
    assert r.ref_key_set(h, -1, d, P(rlk[d])) == 0
    assert r.ref_key_set(h, 0, d, P(glk[d])) == 0
a, b = np.load(f"{work}/a.npy"), np.load(f"{work}/b.npy")
l = a.shape[1]
out = np.zeros((2, l, n), dtype=np.uint64)
assert r.ref_multiply_relin(h, 1, P(a), P(b), P(out)) == 0, r.ref_last_error()
np.save(f"{work}/got_mul.npy", out)
out = np.zeros((2, l, n), dtype=np.uint64)
assert r.ref_rotate(h, 1, P(a), 1, P(out)) == 0, r.ref_last_error()
np.save(f"{work}/got_rot.npy", out)
out = np.zeros((2, l - 1, n), dtype=np.uint64)
if scheme == 3:
    assert r.ref_rescale(h, 1, P(a), 2, P(out)) == 0, r.ref_last_error()
else:
    assert r.ref_mod_switch(h, 1, P(a), 2, P(out)) == 0, r.ref_last_error()
np.save(f"{work}/got_down.npy", out)
times = (ctypes.c_double * 60)()
assert r.ref_time_op(h, 0, 1, P(a), P(b), 0, 0, 60, times) == 0
np.save(f"{work}/time_us.npy", np.array(sorted(times[10:])))
r.ref_destroy(h)
print("worker ok")
