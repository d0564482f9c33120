# tuning aid (not a test): A/B of the k_raster_fast instantiations on one GPU.
# usage: bash tests/ab_raster_variants.sh "variant[:ctas[:carveout]] ..." "variants to run the parity suite with"
mkdir -p gpurun_out; : > gpurun_out/ab.log
for spec in ${1:-0 8}; do
  IFS=: read v c k <<< "$spec"
  echo "== variant $v ctas ${c:-default} carveout ${k:-default}" >> gpurun_out/ab.log
  env CHORD_RASTER_VARIANT=$v ${c:+CHORD_RASTER_CTAS=$c} ${k:+CHORD_RASTER_CARVEOUT=$k} timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-frame 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'e2e',d['e2e']['value'])" >> gpurun_out/ab.log
done
for v in $2; do
  echo "== parity variant $v" >> gpurun_out/ab.log
  CHORD_RASTER_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
