#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02b_pytest_gpu.log
for s in cornell random cornell-glossy cornell-lucy hdri-test; do python tools/gpu_perf.py $s 64 2>&1 | tail -1; done > $out/r02b_scenes.log; cat $out/r02b_scenes.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/r02b_bench.json 2> $out/r02b_bench.err; echo "bench rc=$?"; cat $out/r02b_bench.json
