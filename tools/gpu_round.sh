#!/bin/bash
timeout 300 python tools/gpu_perf.py cornell-lucy 48 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|assert|passed|failed" | head -20
