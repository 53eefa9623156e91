# experiment: warps per CTA of k_contour_hull (variants built with -DEMIA_PRESORT_WARPS=w -DEMIA_HULL_MIN_CTAS=32/w)
for t in 256 2048; do for lib in libemia.so libemia_w2.so libemia_w1.so; do
  EMIA_LIB_PATH=$PWD/deepemia_b200/$lib python bench.py --tiles $t --device-pass-only --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t $lib', round(d['ms_per_step'],4))"
done; done
