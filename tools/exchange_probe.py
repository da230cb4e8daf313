"""Two (or more) ranks: the rooted and the spread plan through the NCCL pipeline (Exchange) and the copy-engine pipeline
(PullExchange) with real HMult+Relin ops, every result compared with the same op computed where its operands live.
torchrun --nproc-per-node 2 tools/exchange_probe.py"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phantom_fhe_b200 as pf  # noqa: E402
from phantom_fhe_b200 import lib, check  # noqa: E402
from phantom_fhe_b200.shard import ExchangePlan, Exchange, PullExchange, peer_view  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 16384
primes = pf.CoeffModulus.Create(n, [60] + [40] * 5 + [60] * 2)
parms = pf.EncryptionParameters(pf.scheme_type.ckks)
parms.set_poly_modulus_degree(n)
parms.set_coeff_modulus(primes)
parms.set_special_modulus_size(2)
ctx = pf.PhantomContext(parms)
l, words = 6, 2 * 6 * n
gen = torch.Generator(device=dev)
gen.manual_seed(5)   # same key on every rank
digits = [torch.empty((2, 8, n), dtype=torch.int64, device=dev) for _ in range(ctx.dnum(1))]
for d in digits:
    for j, q in enumerate(primes):
        d[:, j, :] = torch.randint(0, q, (2, n), generator=gen, device=dev, dtype=torch.int64)
rlk = pf.PhantomRelinKey.from_device(ctx, digits)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
B, chunk = 22, 4   # ragged on purpose


def make(units, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    t = torch.empty((units, 2 * words), dtype=torch.int64, device=dev)
    v = t.view(units, 4, l, n)
    for j in range(l):
        v[:, :, j, :] = torch.randint(0, primes[j], (units, 4, n), generator=g, device=dev, dtype=torch.int64)
    return t


def compute(t, views):
    for vin, vout in views:
        for j in range(vin.shape[0]):
            a = vin[j].data_ptr()
            check(lib.pfhe_multiply_and_relin(ctx._h, 1, a, a + words * 8, vout[j].data_ptr(), rlk.public_keys_ptr(), st))


def local_results(store_in):
    out = torch.empty((store_in.shape[0], words), dtype=torch.int64, device=dev)
    compute(0, [(store_in, out)])
    torch.cuda.synchronize()
    return out


def fence():
    torch.cuda.synchronize()
    dist.barrier()


for kind in ("rooted", "spread"):
    plan = getattr(ExchangePlan, kind)(B, world, chunk)
    lo, hi = plan.home[rank]
    store_in = make(hi - lo, 100 + rank)
    want = local_results(store_in)
    # two-sided, NCCL
    store_out = torch.zeros((hi - lo, words), dtype=torch.int64, device=dev)
    ex = Exchange(plan, rank, store_in, store_out, dist)
    for _ in range(2):
        store_out.zero_()
        fence()
        ex.run(compute)
        fence()
        assert torch.equal(store_out, want), f"{kind} over NCCL"
    # one-sided, copy engines over IPC mappings
    store_out2 = torch.zeros((hi - lo, words), dtype=torch.int64, device=dev)
    homes_in, homes_out, maps = [None] * world, [None] * world, []
    for r in range(world):
        if plan.home[r][1] > plan.home[r][0]:
            homes_in[r], m = peer_view(store_in if rank == r else None, r, rank, dist, dev)
            maps.append(m)
            homes_out[r], m = peer_view(store_out2 if rank == r else None, r, rank, dist, dev)
            maps.append(m)
        elif r == rank:
            homes_in[r], homes_out[r] = store_in, store_out2
    px = PullExchange(plan, rank, homes_in, homes_out)
    for _ in range(2):
        store_out2.zero_()
        fence()
        px.run(compute)
        fence()
        assert torch.equal(store_out2, want), f"{kind} over the copy engines"
    del px, homes_in, homes_out
    for m in maps:
        if m is not None:
            m.close()
fence()
if rank == 0:
    print("exchange ok: rooted and spread plans, NCCL and copy-engine pipelines, results equal the local run", flush=True)
dist.destroy_process_group()
