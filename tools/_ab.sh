python tools/ab_outputs.py pnr_b200/_lib/libfrangi_gpu_base.so pnr_b200/_lib/libfrangi_gpu.so pnr_b200/_lib/libfrangi_gpu_tx120.so | grep -v identical
for v in "" _tx120 "" _tx120; do
  FRANGI_GPU_LIB=$PWD/pnr_b200/_lib/libfrangi_gpu$v.so python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-exact --no-verify > gpurun_out/ab3$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab3$v.json')); print('lib$v', round(d['ms_per_step'],3), {k:round(x,3) for k,x in d['roofline']['per_class_ms_per_step'].items()})"
done
