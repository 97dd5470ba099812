#!/bin/bash
RTX_OPTS=l2_persist=0 timeout 300 python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1
timeout 300 python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print(p.name, p.L2_cache_size, getattr(p,'persisting_l2_cache_max_size',None), getattr(p,'access_policy_max_window_size',None))
PY
