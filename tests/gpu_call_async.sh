#!/bin/bash
# Tuning aid: the asynchronous-resolve renderer on the GPU: its parity cases, the frame leg of bench.py (async vs serial), tile pass times.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "async or sky or native_renderer or lighting" > gpurun_out/gpu_tests_async.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/gpu_tests_async.log
timeout 300 python bench.py --frame-only --steps 20 > gpurun_out/frame_async.json 2> gpurun_out/frame_async.err
echo "frame exit $?"; tail -3 gpurun_out/frame_async.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/frame_async.json') if l.startswith('{')][-1])
    print('async ms', d['ms_per_frame'], 'serial ms', d.get('ms_per_frame_serial'), {k: round(v) for k, v in d['pass_us'].items()})
except Exception as e:
    print('unreadable', e)
PY
timeout 300 python tests/profile_frame_tiles.py > gpurun_out/frame_tiles.json 2> gpurun_out/frame_tiles.err
echo "tiles exit $?"; tail -2 gpurun_out/frame_tiles.err; cat gpurun_out/frame_tiles.json
