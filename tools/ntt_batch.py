"""Forward NTT at the mod-up shape (4 polynomials x 16 limbs per launch pair) for ncu captures."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
ctx = pf.PhantomContext(parms)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
bufs = [torch.from_numpy(H.uniform_limbs(ps, list(range(16)), 3 + i, polys=4).view("int64")).cuda() for i in range(3)]
for i in range(reps):
    pf.check(pf.lib.pfhe_ntt_forward_inplace_batch(ctx._h, bufs[i % 3].data_ptr(), 4, 16, 0, st))
torch.cuda.synchronize()
print("ok")
