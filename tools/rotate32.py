"""BASELINE.json config 4: CKKS rotate (Galois key switch), N=2^16, L=16, 32 rotation steps, each on a fresh copy of the
same ciphertext: one at a time, through pfhe_rotate_batch (lanes), and the unmodified reference."""
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

steps = list(range(1, 33))
ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
parms.set_galois_elts(pf.get_elts_from_steps(steps, ps.n))
ctx = pf.PhantomContext(parms)
a = H.ciphertext(ps, 1)
glk = pf.PhantomGaloisKey(ctx, [list(H.switch_key(ps, 1000 * s)) for s in steps])   # 32 x 80 MiB
src = pf.PhantomCiphertext.from_host(ctx, a)
copies = [src.clone() for _ in steps]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def refill():
    for c in copies:
        c.data.copy_(src.data)


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        refill()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000)
    return best


def one_by_one():
    for c, s in zip(copies, steps):
        pf.rotate_inplace(ctx, c, s, glk)


t1 = timed(one_by_one)
print(f"32 rotations, one at a time: {t1:.0f} us ({t1 / 32:.1f} us per rotation)")
for lanes in (2, 3):
    pf.check(pf.lib.pfhe_engine_set_lanes(ctx._h, lanes))
    tb = timed(lambda: pf.rotate_batch(ctx, copies, steps, glk))
    print(f"32 rotations, pfhe_rotate_batch, {lanes} lanes: {tb:.0f} us ({tb / 32:.1f} us per rotation)")
hoist = src.clone()


def hoisted():
    pf.hoisting_inplace(ctx, hoist, glk, steps)


th = timed(hoisted)
print(f"hoisting_inplace over the 32 steps (sum of the rotations, one shared mod-up): {th:.0f} us")
r = H.reference()
if r is not None:
    arr = (ctypes.c_int * 1)(1)
    h = r.ref_create(3, ps.n, P(ps.primes), ps.size_QP, ps.size_P, 0, 0, arr, 1, float(2 ** 40), 1)
    times = (ctypes.c_double * 40)()
    assert r.ref_time_op(h, 1, 1, P(a), P(a), 1, 0, 40, times) == 0
    ts = sorted(times[8:])
    print(f"reference rotate_inplace: median {ts[len(ts) // 2]:.1f} us per rotation -> {32 * ts[len(ts) // 2]:.0f} us for 32")
    r.ref_destroy(h)
