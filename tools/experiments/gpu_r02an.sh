#!/bin/bash
timeout 120 python tools/gpu_two_lanes.py cornell-lucy 16 2 2>&1 | tail -3
timeout 120 python tools/gpu_two_lanes.py cornell-lucy 64 2 2>&1 | tail -3
timeout 120 python tools/gpu_two_lanes.py cornell-lucy 64 4 2>&1 | tail -3
timeout 120 python tools/gpu_two_lanes.py random 64 2 2>&1 | tail -2
timeout 120 python tools/gpu_two_lanes.py hdri-test 64 2 2>&1 | tail -2
