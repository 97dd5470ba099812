#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "lucy or golden or small_batch or soup or device_bvh or level1" > $out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02j_pytest.log
for i in 1 2; do python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-220; done
python tools/gpu_perf.py random 64 2>&1 | tail -1 | cut -c1-220
