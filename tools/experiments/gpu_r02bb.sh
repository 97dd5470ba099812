#!/bin/bash
for s in "cornell-lucy 64" "random 64" "cornell-lucy 16" "cornell-lucy 64"; do timeout 200 bash tools/ab_run.sh $s 2>&1 | cut -c1-150; done
RTX_B200_LIB=$PWD/build/ab/librtx_q1.so timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_batch or level2_configured or sample_slices" 2>&1 | tail -3
