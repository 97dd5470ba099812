#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 10 -c 1 -f -o $out/r02_prof_shade \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/r02_ncu_shade.log 2>&1
tail -2 $out/r02_ncu_shade.log | cut -c1-200
python tools/ncu_to_json.py $out/r02_prof_shade.ncu-rep k_shade 8388608 "ncu --set full --clock-control none --import-source on -k regex:k_shade -s 10 -c 1 env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64" > $out/r02_k_shade.json
python tools/ncu_summary.py $out/r02_prof_shade.ncu-rep 120 > $out/r02_k_shade_sass.txt 2>&1
cat $out/r02_k_shade.json | head -60
