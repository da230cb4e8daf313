"""Batch data plane (phantom-fhe_b200/shard.py): plans cover every unit exactly once, and the double-buffered
scatter / compute / gather pipeline delivers every result to its home -- world_size 2 and 3 over gloo on the CPU
(the GPU run uses the same code over NCCL, bench.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("total", [0, 1, 5, 16, 37, 1024])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("chunk", [1, 4, 8])
def test_plans_cover_the_batch(total, world, chunk):
    from phantom_fhe_b200.shard import ExchangePlan
    for plan in (ExchangePlan.rooted(total, world, chunk), ExchangePlan.local(total, world, chunk),
                 ExchangePlan.spread(total, world, chunk)):
        seen = []
        for r in range(world):
            for tick in plan.ticks[r]:
                size = 0
                for p in tick:
                    lo, hi = plan.home[p.home]
                    assert lo <= p.lo < p.hi <= hi          # a piece lies inside its home's storage
                    assert p.off == size
                    size += len(p)
                    seen.extend(range(p.lo, p.hi))
                assert 0 < size <= plan.slot
        assert sorted(seen) == list(range(total))
        assert sum(plan.computed_by(r) for r in range(world)) == total
        assert sum(hi - lo for lo, hi in plan.home) == total
        if plan.kind == "local":
            assert all(plan.bytes_moved(r, 32, 16) == (0, 0) for r in range(world))
    if world > 1 and total >= world:
        sent, recv = ExchangePlan.rooted(total, world, chunk).bytes_moved(0, 32, 16)
        remote = total - ExchangePlan.rooted(total, world, chunk).computed_by(0)
        assert (sent, recv) == (32 * remote, 16 * remote)


def test_plan_arguments_are_checked():
    from phantom_fhe_b200.shard import ExchangePlan
    with pytest.raises(ValueError):
        ExchangePlan.rooted(8, 2, 0)
    with pytest.raises(ValueError):
        ExchangePlan.rooted(8, 2, 4, root=2)
    with pytest.raises(ValueError):
        ExchangePlan.spread(8, 0, 4)


def test_single_rank_exchange_runs_in_place():
    import torch
    from phantom_fhe_b200.shard import ExchangePlan, Exchange
    total, iw, ow = 11, 6, 3
    src = torch.arange(total * iw, dtype=torch.int64).view(total, iw)
    out = torch.zeros((total, ow), dtype=torch.int64)
    ex = Exchange(ExchangePlan.rooted(total, 1, 4), 0, src, out)
    order = []

    def compute(t, views):
        order.append(t)
        for vin, vout in views:
            vout.copy_(vin[:, :ow] * 3 + vin[:, ow:])
    ex.run(compute)
    assert order == [0, 1, 2]
    assert torch.equal(out, src[:, :ow] * 3 + src[:, ow:])
    assert ex.stage_in == []   # nothing remote: no staging


WORKER = r"""
import sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from phantom_fhe_b200.shard import ExchangePlan, Exchange, shard_range
world = {world}
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=world)
rank = dist.get_rank()
iw, ow = 10, 5

def unit_in(i):
    return torch.arange(iw, dtype=torch.int64) + 1000 * i

def expect(i):
    v = unit_in(i)
    return v[:ow] * 7 + v[ow:] + 1

def compute(t, views):
    for vin, vout in views:
        vout.copy_(vin[:, :ow] * 7 + vin[:, ow:] + 1)

for total in (0, 3, 16, 37):
    for chunk in (1, 4, 6):
        for kind in ("rooted", "spread", "local"):
            plan = getattr(ExchangePlan, kind)(total, world, chunk)
            lo, hi = plan.home[rank]
            src = torch.stack([unit_in(i) for i in range(lo, hi)]) if hi > lo else torch.zeros((0, iw), dtype=torch.int64)
            out = torch.full((hi - lo, ow), -1, dtype=torch.int64)
            ex = Exchange(plan, rank, src, out, dist)
            for rep in range(2):                      # back-to-back passes reuse the staging slots
                out.fill_(-1)
                ex.run(compute)
                dist.barrier()
                for i in range(lo, hi):
                    assert torch.equal(out[i - lo], expect(i)), (kind, total, chunk, i)
            # every unit is computed exactly once over all ranks
            cnt = torch.tensor([plan.computed_by(rank)])
            dist.all_reduce(cnt)
            assert int(cnt) == total
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_over_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, world=world, port=31000 + (os.getpid() * 7 + world) % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_pull_exchange_single_process(world):
    """the one-sided pipeline (PullExchange): ranks run one after another against shared home storage"""
    import torch
    from phantom_fhe_b200.shard import ExchangePlan, PullExchange
    iw, ow = 10, 5

    def unit_in(i):
        return torch.arange(iw, dtype=torch.int64) + 1000 * i

    def compute(t, views):
        for vin, vout in views:
            vout.copy_(vin[:, :ow] * 7 + vin[:, ow:] + 1)

    for total in (0, 3, 16, 37):
        for chunk in (1, 4, 6):
            for kind in ("rooted", "spread", "local"):
                plan = getattr(ExchangePlan, kind)(total, world, chunk)
                homes_in, homes_out = [], []
                for r in range(world):
                    lo, hi = plan.home[r]
                    homes_in.append(torch.stack([unit_in(i) for i in range(lo, hi)]) if hi > lo
                                    else torch.zeros((0, iw), dtype=torch.int64))
                    homes_out.append(torch.full((hi - lo, ow), -1, dtype=torch.int64))
                for r in range(world):
                    PullExchange(plan, r, homes_in, homes_out).run(compute)
                for r in range(world):
                    lo, hi = plan.home[r]
                    for i in range(lo, hi):
                        v = unit_in(i)
                        assert torch.equal(homes_out[r][i - lo], v[:ow] * 7 + v[ow:] + 1), (kind, total, chunk, i)
