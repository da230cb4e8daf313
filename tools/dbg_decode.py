"""debug: the reference's CKKS decode on its own encodings (needs oracle/_ref)"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import harness as H
P = H.P
dp = ctypes.POINTER(ctypes.c_double)
r = H.reference()
r.ref_ckks_roundtrip.argtypes = [ctypes.c_void_p, dp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_double, dp]
for n, bits, ci in ((4096, [50, 40, 40, 50], 1), (4096, [50, 40, 40, 50], 3), (8192, [60, 40, 40, 40, 60], 1), (16384, [60, 40, 40, 40, 60], 1)):
    ps = H.ParamSet("d", n, bits, 1, 3, 0)
    scale = 2.0 ** 40 if ci == 1 else 2.0 ** 30
    h = r.ref_create(3, ps.n, P(ps.primes), ps.size_QP, ps.size_P, 0, 0, None, 0, scale, 0)
    slots = n // 2
    l = ps.size_Q - (ci - 1)
    rng = np.random.default_rng(1)
    z = rng.uniform(-4, 4, slots) + 1j * rng.uniform(-4, 4, slots)
    flat = np.ascontiguousarray(z.view(np.float64))
    out = np.zeros(2 * slots)
    rc = r.ref_ckks_roundtrip(h, flat.ctypes.data_as(dp), slots, ci, scale, out.ctypes.data_as(dp))
    print(n, ci, "roundtrip rc", rc, "max err", np.max(np.abs(out.view(np.complex128) - z)), out[:4], flush=True)
    plain = np.zeros((l, n), dtype=np.uint64)
    r.ref_ckks_encode(h, flat.ctypes.data_as(dp), slots, ci, scale, P(plain))
    out2 = np.zeros(2 * slots)
    rc = r.ref_ckks_decode(h, P(plain), ci, scale, out2.ctypes.data_as(dp))
    print(n, ci, "decode-of-copy rc", rc, "max err", np.max(np.abs(out2.view(np.complex128) - z)), out2[:4], flush=True)
    o, oc = H.oracle(), ps.octx()
    want = np.zeros(2 * slots)
    o.orc_ckks_decode(oc, l, P(plain), scale, want.ctypes.data_as(dp))
    print("   oracle max err", np.max(np.abs(want.view(np.complex128) - z)), "ref==oracle", np.array_equal(out2.view(np.uint64), want.view(np.uint64)),
          "roundtrip==oracle", np.array_equal(out.view(np.uint64), want.view(np.uint64)), flush=True)
    r.ref_destroy(h)
