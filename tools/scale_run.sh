#!/bin/bash
# bench.py at N = 2, 4, 8 ranks of one box, the way the driver launches it (builder-side scaling table)
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2h_n$N.json 2> gpurun_out/r2h_n$N.err
  echo "== N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_n$N.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "copy-only", round(d["e2e"]["copy_only_ops_per_s"]),
          {k:(round(v["value"]), round(v.get("vs_compute_only",0),3)) for k,v in d["scatter_gather"].items() if isinstance(v,dict) and "value" in v})
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2h_n$N.err").read()[-1500:])
PY
done
