#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "test_order or device_bvh or level1_lucy or soup" > $out/r02am_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r02am_pytest.log
timeout 60 python - <<'PY'
import importlib, time, os
grt = importlib.import_module("go-raytracing_b200")
os.environ["RTX_NO_RANK_CACHE"] = "1"
sc = grt.config_scene("cornell-lucy")
ctx = grt.Context(0)
for i in range(4):
    t0 = time.perf_counter(); ctx.load(sc); t1 = time.perf_counter()
    st = ctx.stats()
    print(f"upload {1e3*(t1-t0):.2f} ms", {k: round(v, 2) for k, v in st.items() if k.startswith("ms_upload") or k == "ms_bvh_build"})
PY
