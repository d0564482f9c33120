#!/bin/bash
# Tuning aid: parity suite + frame benchmark + ncu of the resolve kernel.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log; tail -12 gpurun_out/gpu_tests.log
timeout 300 python bench.py --frame-only --steps 20 > gpurun_out/frame.json 2> gpurun_out/frame.err; echo "frame exit $?"; tail -3 gpurun_out/frame.err
python - <<'PY'
import json
try:
    f = json.loads([l for l in open('gpurun_out/frame.json') if l.startswith('{')][-1])
    print('frame', {k: f.get(k) for k in ('ms_per_frame', 'graph_replays', 'pass_us_sum')}, {k: round(v) for k, v in f.get('pass_us', {}).items()})
    print('next rows', f.get('next_rows'))
except Exception as e:
    print('unreadable', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lighting -s 3 -c 1 -f -o gpurun_out/lighting \
    python bench.py --frame-only --frame-eager --steps 2 > /dev/null 2>&1
echo "ncu exit $?"
