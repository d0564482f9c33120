#!/bin/bash
# Tuning aid: GPU parity suite + the default bench line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log
tail -25 gpurun_out/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
    f = d.get('frame', {}); s2 = d.get('frame_s2', {}); hv = d.get('frame_heavy', {})
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'])
    print('frame', {k: f.get(k) for k in ('ms_per_frame', 'ms_per_frame_wall', 'graph_replays', 'pass_us_sum', 'error')}, {k: round(v) for k, v in f.get('pass_us', {}).items()})
    print('s2', {k: s2.get(k) for k in ('ms_per_frame', 'ms_per_frame_wall', 'graph_replays', 'pass_us_sum', 'error')})
    print('heavy', hv)
except Exception as e:
    print('bench line unreadable:', e)
PY
