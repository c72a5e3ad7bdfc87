# bench.py (short) for the main library and every tuning variant under gingr_b200/lib/variants: it/s and the E-step sweeps
cd "$(dirname "$0")/.."
for lib in gingr_b200/lib/libgingr_cuda.so gingr_b200/lib/variants/libgingr_cuda_*.so; do
  GINGR_CUDA_LIB=$PWD/$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms']
print('$lib'.split('/')[-1], 'it/s %.2f' % d['value'], 'A=%.3f B=%.3f' % (p['estep_sweepA'], p['estep_sweepB']))"
done
