#!/bin/bash
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "cornell-glossy-96" 2>&1 | grep -E "^E   .*(assert|Error)|passed|failed" | head -4; done
