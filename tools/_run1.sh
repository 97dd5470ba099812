python -m pytest tests -x -q -m gpu 2>&1 | tail -2
P() { tail -1 | sed 's/ nodes\/ray.*//' | cut -c1-190; }
for s in cornell cornell-smoke; do for o in lean=0 lean=1; do RTX_OPTS=$o python tools/gpu_perf.py $s 64 2>&1 | P; done; done
