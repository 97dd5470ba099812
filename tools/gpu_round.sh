#!/bin/bash
timeout 120 python tools/gpu_check.py cornell-lucy cornell 2>&1 | grep -E "trace|render|secondary|Error|error" 
timeout 120 python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1
