mkdir -p gpurun_out; : > gpurun_out/frame_ab.log
for spec in "11:0" "11:4096" "11:16384" "11:100000000" "0:0" "11:16384"; do
  IFS=: read v m <<< "$spec"
  echo "== variant $v warp_min $m" >> gpurun_out/frame_ab.log
  CHORD_RASTER_VARIANT=$v CHORD_RASTER_WARP_MIN=$m timeout 120 python bench.py --frame-only 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); d=d.get('frame',d); print('fps',d['frames_per_s'],'ms',d['ms_per_frame'],{k:round(v,1) for k,v in d['pass_us'].items()}, 'shadow', d['next_rows']['shadow']['ms_per_pass'])" >> gpurun_out/frame_ab.log
done
CHORD_RASTER_WARP_MIN=0 timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/frame_ab.log
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/frame_ab.log
timeout 100 python bench.py --no-frame --no-cpu-baseline | cut -c1-200 >> gpurun_out/frame_ab.log
cat gpurun_out/frame_ab.log
