"""Two-GPU checks (skipped on a single-GPU box): kernels of rank 1 working on operands in rank 0's HBM through the CUDA IPC
mapping (pfhe_ipc_export / pfhe_ipc_open), and the scatter / compute / gather pipelines of phantom-fhe_b200/shard.py over NCCL
and over the copy engines, each checked against the same ops computed locally."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port, timeout=600):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_kernels_on_a_peer_mapping_are_bit_exact():
    p = _torchrun([os.path.join(ROOT, "tools", "peer_probe.py")], 29561)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "bit-exact" in p.stdout


def test_exchange_pipelines_deliver_the_results_of_the_local_run():
    p = _torchrun([os.path.join(ROOT, "tools", "exchange_probe.py")], 29562)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "exchange ok" in p.stdout


def test_bench_line_at_two_gpus():
    p = _torchrun([os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--batch", "32", "--no-cpu-baseline",
                   "--no-extra"], 29563)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["gpu_launches"] > 0
    for leg in ("rooted_nccl", "spread_nccl", "rooted_pull", "spread_pull"):
        assert line["scatter_gather"][leg]["value"] > 0, leg
