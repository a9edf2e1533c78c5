#!/bin/bash
# ncu captures of the memory-side kernels (run under gpurun; outputs land in gpurun_out/)
set -x
ncu --set full --clock-control none --import-source on -k regex:per_sample_gather -s 4 -c 2 \
    -o gpurun_out/ncu_gather_$1 -f python tools/bench_replay.py --quick > gpurun_out/ncu_gather_$1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:per_update_cta -s 2 -c 1 \
    -o gpurun_out/ncu_update_$1 -f python tools/bench_replay.py --quick >> gpurun_out/ncu_gather_$1.log 2>&1
tail -3 gpurun_out/ncu_gather_$1.log
