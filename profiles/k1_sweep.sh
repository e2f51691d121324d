# K1 generic-kernel tuning sweep (run under gpurun): pixels per thread x CTAs per SM, batched s = 2 / 4
for cfg in "8 5" "8 4" "8 6" "4 8" "4 6" "4 10"; do
  set -- $cfg
  T3D_K1_PPT=$1 T3D_K1_MINB=$2 python profiles/bench_kernels.py --only K1 2>/dev/null | grep batch | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('ppt $1 minb $2', d['kernel'][:32], round(d['units_per_s']), 'frames/s frac %.3f' % d['frac'])
"
done
