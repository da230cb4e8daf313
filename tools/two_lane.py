"""Throughput of independent HMult+Relin ops issued round-robin over L engines / streams (potential of multi-lane batching)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ps = H.params_primary()
a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
rlk_h = list(H.switch_key(ps, 100))
for lanes in (1, 2, 3, 4):
    ctxs, streams, cas, cbs, works = [], [], [], [], []
    for _ in range(lanes):
        parms = pf.EncryptionParameters(pf.scheme_type.ckks)
        parms.set_poly_modulus_degree(ps.n)
        parms.set_coeff_modulus([int(p) for p in ps.primes])
        parms.set_special_modulus_size(ps.size_P)
        ctx = pf.PhantomContext(parms)
        ctxs.append(ctx)
        streams.append(torch.cuda.Stream())
        cas.append(pf.PhantomCiphertext.from_host(ctx, a))
        cbs.append(pf.PhantomCiphertext.from_host(ctx, b))
        works.append(cas[-1].data.clone())
    rlk = pf.PhantomRelinKey(ctxs[0], rlk_h)

    def op(i):
        k = i % lanes
        st = ctypes.c_void_p(streams[k].cuda_stream)
        pf.check(pf.lib.pfhe_multiply_and_relin(ctxs[k]._h, 1, cas[k].data.data_ptr(), cbs[k].data.data_ptr(),
                                                works[k].data_ptr(), rlk.public_keys_ptr(), st))

    best = 1e9
    for trial in range(3):
        for i in range(20):
            op(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(reps):
            op(i)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000 / reps)
    print(f"lanes={lanes}: {best:.1f} us/op ({1e6 / best:.0f} ops/s) OVERLAP={os.environ.get('PFHE_OVERLAP', '1')}", flush=True)
    del ctxs, cas, cbs, works, rlk
