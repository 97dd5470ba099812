python -m pytest tests -x -q -m gpu 2>&1 | tail -2
P() { tail -1 | sed 's/ nodes\/ray.*//' | cut -c1-190; }
for s in cornell-lucy random; do
  python tools/gpu_perf.py $s 64 2>&1 | P
  for v in t6s128 t7 t8; do RTX_B200_LIB=$PWD/build/ab/librtx_$v.so python tools/gpu_perf.py $s 64 2>&1 | P; done
done
for s in cornell hdri-test cornell-glossy; do python tools/gpu_perf.py $s 64 2>&1 | P; done
python bench.py > gpurun_out/r01n_bench.json 2> gpurun_out/r01n_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r01n_bench.json
