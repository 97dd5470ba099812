#!/bin/bash
# The bench lines of a round without the profiler passes: bash tools/gpu_bench_only.sh <tag>
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-200 $out/${tag}_bench.json
python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; cut -c1-200 $out/${tag}_bench_ref.json
for wl in cornell random cornell-glossy hdri-test; do python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$wl.json 2>> $out/${tag}_bench.err; cut -c1-120 $out/${tag}_bench_$wl.json; done
