"""Two ranks: rank 1 runs HMult+Relin on operands that live in rank 0's HBM (CUDA IPC mapping over NVLink) and writes the
result there; rank 0 checks it against the same op on local memory.  torchrun --nproc-per-node 2 tools/peer_probe.py"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phantom_fhe_b200 as pf  # noqa: E402
from phantom_fhe_b200 import lib, check  # noqa: E402
from phantom_fhe_b200.shard import peer_view  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 65536
primes = pf.CoeffModulus.Create(n, [60] + [40] * 15 + [60] * 4)
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(n)
parms.set_coeff_modulus(primes)
parms.set_special_modulus_size(4)
ctx = pf.PhantomContext(parms)
l, words = 16, 2 * 16 * n
gen = torch.Generator(device=dev)
gen.manual_seed(5)
digits = [torch.empty((2, 20, n), dtype=torch.int64, device=dev) for _ in range(4)]
for d in digits:
    for j, q in enumerate(primes):
        d[:, j, :] = torch.randint(0, q, (2, n), generator=gen, device=dev, dtype=torch.int64)
rlk = pf.PhantomRelinKey.from_device(ctx, digits)
units = 4
store_in = torch.empty((units, 2 * words), dtype=torch.int64, device=dev)
store_out = torch.zeros((units, words), dtype=torch.int64, device=dev)
v = store_in.view(units, 4, l, n)
for j in range(l):
    v[:, :, j, :] = torch.randint(0, primes[j], (units, 4, n), generator=gen, device=dev, dtype=torch.int64)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rin, m_in = peer_view(store_in if rank == 0 else None, 0, rank, dist, dev)
rout, m_out = peer_view(store_out if rank == 0 else None, 0, rank, dist, dev)
torch.cuda.synchronize()
dist.barrier()
if rank == 1:
    for i in range(units):
        a = rin[i].data_ptr()
        check(lib.pfhe_multiply_and_relin(ctx._h, 1, a, a + words * 8, rout[i].data_ptr(), rlk.public_keys_ptr(), st))
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    ref = torch.empty_like(store_out)
    for i in range(units):
        a = store_in[i].data_ptr()
        check(lib.pfhe_multiply_and_relin(ctx._h, 1, a, a + words * 8, ref[i].data_ptr(), rlk.public_keys_ptr(), st))
    torch.cuda.synchronize()
    ok = torch.equal(ref, store_out)
    print("peer HMult+Relin over NVLink:", "bit-exact" if ok else "MISMATCH", flush=True)
    assert ok
dist.barrier()
for m in (m_in, m_out):
    if m is not None:
        m.close()
dist.destroy_process_group()
