"""Runs a few HMult+Relin ops (for ncu launch lists)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import phantom_fhe_b200 as pf  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ps = H.params_primary()
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(ps.n)
parms.set_coeff_modulus([int(p) for p in ps.primes])
parms.set_special_modulus_size(ps.size_P)
ctx = pf.PhantomContext(parms)
a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
rlk = pf.PhantomRelinKey(ctx, list(H.switch_key(ps, 100)))
ca, cb = pf.PhantomCiphertext.from_host(ctx, a), pf.PhantomCiphertext.from_host(ctx, b)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
work = ca.data.clone()
torch.cuda.synchronize()
for _ in range(reps):
    pf.check(pf.lib.pfhe_multiply_and_relin(ctx._h, 1, ca.data.data_ptr(), cb.data.data_ptr(), work.data_ptr(),
                                            rlk.public_keys_ptr(), st))
torch.cuda.synchronize()
print("ok")
