#!/bin/bash
# One GPU-box pass that produces every artefact summarised under profiles/: tools/profile_round.sh <tag>
tag=${1:-r1b}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference > gpurun_out/${tag}_bench_line_reference.json 2>> gpurun_out/${tag}_bench.err
python tools/quick_bench.py primary > gpurun_out/${tag}_quick_primary.txt 2>&1
python tools/quick_bench.py secondary > gpurun_out/${tag}_quick_secondary.txt 2>&1
python tools/bfv_bench.py 100 > gpurun_out/${tag}_bfv.txt 2>&1
python tools/two_lane.py 200 > gpurun_out/${tag}_lanes.txt 2>&1
python tools/rotate32.py > gpurun_out/${tag}_rotate32.txt 2>&1
./tools/microbench > gpurun_out/${tag}_microbench.txt 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 10 -o gpurun_out/${tag}_hmult -f \
    python tools/hmult_only.py 2 > gpurun_out/${tag}_ncu_hmult.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 4 --launch-count 2 -o gpurun_out/${tag}_ntt -f \
    python tools/ntt_batch.py 4 > gpurun_out/${tag}_ncu_ntt.log 2>&1
ls -la gpurun_out/${tag}_*
