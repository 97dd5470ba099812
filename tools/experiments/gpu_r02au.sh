#!/bin/bash
for i in 1 2 3; do timeout 60 python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-150; done
timeout 60 env RTX_OPTS=overlap_connect=0 python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-150
