#!/bin/bash
# ncu --set full captures of the three kernels that carry a BASELINE config, each after the same command exited 0 without ncu.
# Usage (under gpurun): bash tools/gpu_ncu.sh <tag>; reports land in gpurun_out/<tag>_prof_*.ncu-rep
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/${tag}_plain_lucy.log 2>&1 || exit 1
tail -1 $out/${tag}_plain_lucy.log | cut -c1-220
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 10 -c 1 -f -o $out/${tag}_prof_extend \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/${tag}_ncu_extend.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_connect -s 10 -c 1 -f -o $out/${tag}_prof_connect \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/${tag}_ncu_connect.log 2>&1
python tools/gpu_perf.py hdri-test 16 > $out/${tag}_plain_hdri.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bounce_flat -s 6 -c 1 -f -o $out/${tag}_prof_bounce \
    python tools/gpu_perf.py hdri-test 16 > $out/${tag}_ncu_bounce.log 2>&1
ls -la $out/${tag}_prof_*.ncu-rep
# rays of the captured launches (the k_iter_begin log of the same command)
# jobs of the captured launches: ncu -s 10 skips ten launches of the kernel = iteration 10 (0-based) of the FIRST pass; -s 6 = iteration 6
RTX_DEBUG_ITER=10 RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 2>&1 | grep "iteration 10" | head -1 | tee $out/${tag}_rays_lucy.txt
RTX_DEBUG_ITER=6 python tools/gpu_perf.py hdri-test 16 2>&1 | grep "iteration 6" | head -1 | tee $out/${tag}_rays_hdri.txt
