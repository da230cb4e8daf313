"""Runs a few forward/inverse NTT launches (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402

limbs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
ctx = pf.PhantomContext(parms)
x = torch.from_numpy(H.uniform_limbs(ps, list(range(limbs)), 1)[0].view("int64")).cuda()
for _ in range(reps):
    pf.nwt_2d_radix8_forward_inplace(x, ctx, limbs, 0)
    pf.nwt_2d_radix8_backward_inplace(x, ctx, limbs, 0)
torch.cuda.synchronize()
print("ok")
