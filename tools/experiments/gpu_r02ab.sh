#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "lucy or golden or soup or device_bvh or level1 or small_batch" > $out/r02ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02ab_pytest.log
for i in 1 2; do for v in 1 0; do RTX_OPTS=tri_pretest=$v python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-170; done; done
for v in 1 0; do RTX_OPTS=tri_pretest=$v,count_stats=3 python tools/gpu_perf.py cornell-lucy 8 2>&1 | tail -1 | sed "s/.*iters/iters/"; done
