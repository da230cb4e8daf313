"""Host-side issue time vs device time of the batched HMult+Relin entry point (is the launch path the limit?)."""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402

ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
ctx = pf.PhantomContext(parms)
rlk = pf.PhantomRelinKey(ctx, list(H.switch_key(ps, 100)))
n_pairs, count = 8, 256
da = [pf.PhantomCiphertext.from_host(ctx, H.ciphertext(ps, 10 + i)).data for i in range(n_pairs)]
db = [pf.PhantomCiphertext.from_host(ctx, H.ciphertext(ps, 30 + i)).data for i in range(n_pairs)]
out = [torch.empty_like(da[0]) for _ in range(n_pairs)]
Arr = ctypes.c_void_p * count
aa = Arr(*[da[i % n_pairs].data_ptr() for i in range(count)])
bb = Arr(*[db[i % n_pairs].data_ptr() for i in range(count)])
oo = Arr(*[out[i % n_pairs].data_ptr() for i in range(count)])
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for lanes in (1, 2, 3):
    pf.check(pf.lib.pfhe_engine_set_lanes(ctx._h, lanes))
    pf.check(pf.lib.pfhe_multiply_and_relin_batch(ctx._h, 1, aa, bb, oo, 16, rlk.public_keys_ptr(), st))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pf.check(pf.lib.pfhe_multiply_and_relin_batch(ctx._h, 1, aa, bb, oo, count, rlk.public_keys_ptr(), st))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"lanes={lanes}: host issue {1e6 * (t1 - t0) / count:.1f} us/op, issue+drain {1e6 * (t2 - t0) / count:.1f} us/op", flush=True)
