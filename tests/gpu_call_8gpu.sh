#!/bin/bash
# Tuning aid: the N-GPU bench line + the multi-GPU parity script + the -m gpu multi-rank tests, in one gpurun --gpus N call.
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l > gpurun_out/ngpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_${N}gpu.json') if l.startswith('{')][-1])
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'parity', d.get('parity'))
    f = d.get('frame', {}); print('frame', {k: f.get(k) for k in ('ms_per_frame', 'exchange_timed_out', 'error')}, 'nccl', (f.get('nccl_exchange') or {}).get('ms_per_frame'))
    print('heavy', (d.get('frame_heavy') or {}).get('ms_per_frame'))
except Exception as e:
    print('unreadable', e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/run_multigpu_check.py > gpurun_out/multigpu_check_${N}.log 2>&1
echo "multigpu check exit $?"; tail -3 gpurun_out/multigpu_check_${N}.log
timeout 400 python -m pytest tests -m gpu -x -q -k "multirank or exchange or multi" > gpurun_out/gpu_tests_${N}gpu.log 2>&1
echo "multi-rank gpu tests exit $?"; tail -3 gpurun_out/gpu_tests_${N}gpu.log
