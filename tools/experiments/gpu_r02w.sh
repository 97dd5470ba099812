#!/bin/bash
out=gpurun_out; mkdir -p $out
python bench.py --steps 5 --warmup 3 > $out/r02w_bench.json 2> $out/r02w_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/r02w_bench.json')); print('value',d['value'],'e2e',d['e2e']['value'],'cold',d['e2e_cold']['ms_per_step'], d['e2e_cold']['ms_scene_function_and_flatten_host'], d['e2e_cold']['ms_upload_and_device_builds']); r=d['roofline']; print('frac',r['frac'],'layout',r['layout_bytes_frac'],'share',r['kernel_share_of_step']); print(d['tail']); print([ (x['kernel'],round(x['frac'],3),round(x['share_of_step'],3)) for x in d['roofline_stream']]); print(d['cpu_baseline'])"
for s in cornell random cornell-glossy hdri-test; do python tools/gpu_perf.py $s 64 2>&1 | tail -1 | cut -c1-160; done
