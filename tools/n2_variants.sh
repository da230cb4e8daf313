#!/bin/bash
# scatter/gather variants at N ranks (default 2): NCCL channel count / buffer size, tick size, zero-copy peer leg
N=${1:-2}
run() {
  tag=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --no-cpu-baseline --no-extra $EXTRA > gpurun_out/sg_${tag}.json 2> gpurun_out/sg_${tag}.err
  echo "== $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sg_${tag}.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), {k:(round(v["value"]), round(v.get("vs_compute_only",0),3), round(v.get("rank0_egress_gbs",0))) if "value" in v else v for k,v in d["scatter_gather"].items() if isinstance(v,dict)})
except Exception as e:
    print("no line", e)
PY
}
true
EXTRA="--peer" run peer A=1
exit 0
EXTRA="" run p2p32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
EXTRA="" run buf16 NCCL_BUFFSIZE=16777216
EXTRA="" run ch32buf16 NCCL_MIN_NCHANNELS=32 NCCL_BUFFSIZE=16777216
