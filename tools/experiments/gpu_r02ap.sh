#!/bin/bash
for o in time_kernels=1 time_kernels=0; do timeout 60 env RTX_OPTS=$o python tools/gpu_perf.py hdri-test 64 2>&1 | tail -1 | cut -c1-150;  timeout 60 env RTX_OPTS=$o python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-150; timeout 60 env RTX_OPTS=$o python tools/gpu_perf.py cornell-glossy 256 2>&1 | tail -1 | cut -c1-150; done
