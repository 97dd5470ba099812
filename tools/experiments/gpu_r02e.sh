#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "lucy or golden or soup or device_bvh or lean or axis" > $out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02e_pytest.log
for o in tlas_flat_max=0 tlas_flat_max=16; do RTX_OPTS=$o python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1; RTX_OPTS=$o,count_stats=3 python tools/gpu_perf.py cornell-lucy 8 2>&1 | tail -1 | sed 's/.*iters/iters/'; done
python -m pytest tests/test_estimator_kats.py -m gpu -x -q > $out/r02e_kats.log 2>&1; echo "kats rc=$?"; tail -3 $out/r02e_kats.log
