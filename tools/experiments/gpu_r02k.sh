#!/bin/bash
out=gpurun_out; mkdir -p $out
for i in 1 2; do
python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-200
RTX_B200_LIB=$PWD/build/ab/librtx_nocs.so python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-200
done
for s in random hdri-test cornell-glossy; do python tools/gpu_perf.py $s 64 2>&1 | tail -1 | cut -c1-160; RTX_B200_LIB=$PWD/build/ab/librtx_nocs.so python tools/gpu_perf.py $s 64 2>&1 | tail -1 | cut -c1-160; done
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "lucy or golden or level2" > $out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02k_pytest.log
