#!/bin/bash
for p in 1048576 4194304 8388608; do RTX_OPTS=pool_paths=$p timeout 300 python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1; done
timeout 300 python tools/gpu_check.py cornell 2>&1 | grep -E "trace|render|radiance"
