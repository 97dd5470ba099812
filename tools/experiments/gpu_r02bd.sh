#!/bin/bash
for s in "hdri-test 64" "hdri-test 256" "cornell-lucy 64" "cornell-glossy 256" "random 64"; do for o in 2 1; do timeout 100 env RTX_OPTS=pixel_major=$o python tools/gpu_perf.py $s 2>&1 | tail -1 | cut -c1-150; done; done
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_batch or level2_configured or sample_slices or multi_slice or reduced_depth" 2>&1 | tail -3
