#!/bin/bash
for s in "hdri-test 64" "cornell-glossy 256" "cornell 64" "earth 64" "hdri-test 64"; do timeout 200 bash tools/ab_run.sh $s 2>&1 | cut -c1-150; done
RTX_B200_LIB=$PWD/build/ab/librtx_b1.so timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "level2_configured or level2_other or flat_and_hierarchy or sample_slices or image_textures" 2>&1 | tail -3
