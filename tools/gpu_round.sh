#!/bin/bash
for v in default s192b5 s160b6 s224b4 s128b7; do
  if [ $v = default ]; then unset RTX_B200_LIB; else export RTX_B200_LIB=$PWD/go-raytracing_b200/csrc/variants/librtx_$v.so; fi
  timeout 300 python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1
done
