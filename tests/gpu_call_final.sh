#!/bin/bash
# One call: GPU suite, racecheck over the raster-heavy cases, ncu capture of k_raster_fast for the traffic figure, bench line.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -2 gpurun_out/gpu_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 --log-file gpurun_out/sanitizer_racecheck.log \
    python -m pytest tests/test_gpu_parity.py -q -k "not full_size and not city_4k and not native_cpp and not extreme_frame and not random_camera and not selftest" > gpurun_out/sanitizer_racecheck_pytest.log 2>&1
echo "racecheck exit $? : $(grep 'RACECHECK SUMMARY' gpurun_out/sanitizer_racecheck.log | tail -1)"; tail -1 gpurun_out/sanitizer_racecheck_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_raster_fast -s 4 -c 1 -f -o gpurun_out/raster_fast \
    python bench.py --steps 2 --warmup 1 --no-frame --no-cpu-baseline --no-graph > /dev/null 2>&1
echo "ncu exit $?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
    f = d.get('frame', {})
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'],
          '| frame ms', f.get('ms_per_frame'), 'serial', f.get('ms_per_frame_serial'), {k: round(v) for k, v in f.get('pass_us', {}).items()})
    print('S2', d['frame_s2'].get('ms_per_frame'), d['frame_s2'].get('ms_per_frame_serial'), 'heavy', d['frame_heavy'].get('ms_per_frame'))
except Exception as e:
    print('bench line unreadable:', e)
PY
