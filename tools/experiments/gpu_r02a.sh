#!/bin/bash
# Round-2 GPU call A: compute-sanitizer on small workloads, ncu re-capture of the shipped trace kernels, baseline throughput.
out=gpurun_out; mkdir -p $out
python tools/gpu_sanitize.py cornell random lucy20k > $out/r02a_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 $out/r02a_sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 40 python tools/gpu_sanitize.py cornell random lucy20k > $out/r02a_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $out/r02a_sanitize_$tool.log | tail -3
done
for s in cornell random cornell-glossy cornell-lucy hdri-test; do python tools/gpu_perf.py $s 64 2>&1 | tail -1; done > $out/r02a_scenes.log; cat $out/r02a_scenes.log
RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/r02a_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 10 -c 1 -f -o $out/r02a_prof_extend \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/r02a_ncu.log 2>&1
tail -1 $out/r02a_ncu.log
