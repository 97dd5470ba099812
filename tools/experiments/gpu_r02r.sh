#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_batch or level2_lucy or level2_configured" > $out/r02r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02r_pytest.log
for i in 1 2; do for sd in 0 1; do RTX_OPTS=shade_direct=$sd python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-170; done; done
for sd in 0 1; do RTX_OPTS=shade_direct=$sd python tools/gpu_perf.py random 64 2>&1 | tail -1 | cut -c1-170; done
for sd in 0 1; do RTX_OPTS=shade_direct=$sd,flat_max_entries=0 python tools/gpu_perf.py hdri-test 16 2>&1 | tail -1 | cut -c1-170; done
bash tools/ab_run.sh cornell-lucy 64 2>&1 | cut -c1-150
