#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "lucy or golden or small_batch" > $out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02i_pytest.log
for i in 1 2; do python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-220; done
python bench.py --steps 3 --warmup 3 > $out/r02i_bench.json 2> $out/r02i_bench.err; echo "bench rc=$?"; cat $out/r02i_bench.json; tail -3 $out/r02i_bench.err
python bench.py --workload hdri-test --spp 64 --steps 3 --warmup 3 --no-cpu-baseline > $out/r02i_bench_hdri.json 2>> $out/r02i_bench.err; echo "bench rc=$?"; cat $out/r02i_bench_hdri.json
