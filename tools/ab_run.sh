#!/bin/bash
# Run every build/ab/librtx_*.so through tools/gpu_perf.py (on the GPU box): tools/ab_run.sh [scene] [spp]
scene=${1:-cornell-lucy}; spp=${2:-64}
for f in build/ab/librtx_*.so; do
  RTX_B200_LIB=$PWD/$f timeout 300 python tools/gpu_perf.py $scene $spp 2>&1 | tail -1
done
