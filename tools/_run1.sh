out=gpurun_out; tag=r01q; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-200 $out/${tag}_bench.json
python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; cut -c1-120 $out/${tag}_bench_ref.json
for w in hdri-test cornell-glossy random cornell; do python bench.py --workload $w > $out/${tag}_bench_$w.json 2>> $out/${tag}_bench.err; echo "$w rc=$?"; cut -c1-120 $out/${tag}_bench_$w.json; done
python bench.py --spp 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --spp 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_l.log 2>&1
for s in cornell random cornell-glossy cornell-lucy hdri-test quads earth cornell-smoke checkered simple glossy-metal perlin primitives; do python tools/gpu_perf.py $s 64 2>&1 | tail -1 | sed 's/ nodes\/ray.*//'; done > $out/${tag}_scenes.log
cat $out/${tag}_scenes.log | cut -c1-150
