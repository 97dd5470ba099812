#!/bin/bash
out=gpurun_out; mkdir -p $out
# the drain under a watchdog: a hang must not take the box down
for fd in 0 100000 400000; do timeout 60 env RTX_OPTS=fuse_drain=$fd python tools/gpu_perf.py cornell-lucy 16 2>&1 | tail -1 | cut -c1-200; echo "rc=$?"; done
for fd in 0 400000; do timeout 60 env RTX_OPTS=fuse_drain=$fd python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-200; timeout 60 env RTX_OPTS=fuse_drain=$fd python tools/gpu_perf.py random 64 2>&1 | tail -1 | cut -c1-200; done
