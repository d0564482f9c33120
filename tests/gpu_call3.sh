#!/bin/bash
# Tuning aid: packed-block raster A/B (variants / CTAs per SM) + the 2-GPU bench (IPC fused exchange, NCCL exchange, parity keys).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log; tail -4 gpurun_out/gpu_tests.log
for v in "11 -1" "14 -1" "13 -1" "0 -1"; do
  set -- $v
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame --raster-variant $1 --raster-ctas $2 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('variant $1 ctas $2: value', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']))" | tee -a gpurun_out/ab.log
done
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG"
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
  echo "2gpu bench exit $?"; tail -5 gpurun_out/bench_2gpu.err
  python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][-1])
    f = d.get('frame', {})
    print('2gpu value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity'))
    print('frame', {k: f.get(k) for k in ('ms_per_frame', 'ms_per_frame_wall', 'graph_replays', 'exchange_timed_out', 'error', 'parity')})
    print('nccl', f.get('nccl_exchange'))
    print('heavy', d.get('frame_heavy'))
except Exception as e:
    print('2gpu line unreadable:', e)
PY
fi
