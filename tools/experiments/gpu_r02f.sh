#!/bin/bash
out=gpurun_out; mkdir -p $out
nvidia-smi -L
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "multi_device" > $out/r02f_multi.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r02f_multi.log
python tools/bench_multi_inlib.py cornell-lucy 64 > $out/r02f_inlib.log 2>&1; tail -4 $out/r02f_inlib.log
