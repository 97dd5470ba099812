#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r02c_pytest_gpu.log
RTX_DEBUG_BATCH=1 python tools/gpu_perf.py cornell-lucy 16 > $out/r02c_tail16.log 2>&1; grep "iter\|Mpaths" $out/r02c_tail16.log | tail -12
for spp in 8 32 64 128; do python tools/gpu_perf.py cornell-lucy $spp 2>&1 | tail -1 | cut -c1-200; done
