#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/bias_probe.py random 120 512 > gpurun_out/bias.log 2>&1; cat gpurun_out/bias.log
CMD="python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 30 -c 2 -o gpurun_out/prof_extend_ww -f $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
