#!/bin/bash
# scatter/gather tuning at N ranks: tick size x staging depth (compute-only + the exchange legs only)
N=${1:-8}
for cfg in "8 2" "8 3" "16 3" "16 4"; do
  set -- $cfg
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 6 --sg-only --chunk $1 --depth $2 2> gpurun_out/tune_$1_$2.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk',d['chunk'],'depth',d['depth'],'value',round(d['value']),{k:(round(v['value']),round(v['vs_compute_only'],3)) for k,v in d['scatter_gather'].items() if isinstance(v,dict) and 'value' in v})"
done
