#!/bin/bash
for f in build/ab/librtx_*.so; do for fd in 524288 2097152; do for s in "cornell-lucy 16" "cornell-lucy 64" "random 64"; do
  RTX_B200_LIB=$PWD/$f RTX_OPTS=fuse_drain=$fd timeout 100 python tools/gpu_perf.py $s 2>&1 | tail -1 | cut -c1-100; done; done; done
