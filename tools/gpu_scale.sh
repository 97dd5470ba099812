#!/bin/bash
# 1 / 2 / 4 / 8 GPUs of one box: the torchrun bench (one process per GPU) on cornell-lucy and hdri-test, and the in-library path.
# Usage (under gpurun --gpus 8): bash tools/gpu_scale.sh <tag>
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l
for wl in cornell-lucy hdri-test; do
  for n in 1 2 4 8; do
    if [ $n -eq 1 ]; then
      python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_n${n}_${wl}.json 2> $out/${tag}_bench_n${n}_${wl}.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $wl --steps 3 --warmup 3 --no-cpu-baseline \
        > $out/${tag}_bench_n${n}_${wl}.json 2> $out/${tag}_bench_n${n}_${wl}.err
    fi
    python -c "
import json
d=json.loads(open('$out/${tag}_bench_n${n}_${wl}.json').read().strip().splitlines()[-1])
print('$wl', 'N=$n', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'tail', d['tail']['ms_tail'], 'reduce', d['tail']['ms_reduce'], 'inlib', (d.get('in_library_multi_gpu') or {}).get('value'))"
  done
done
python tools/bench_multi_inlib.py cornell-lucy > $out/${tag}_inlib_lucy.jsonl 2> $out/${tag}_inlib.err; cat $out/${tag}_inlib_lucy.jsonl | cut -c1-330
