#!/bin/bash
# Tuning aid (not a test): everything the first GPU call of a round should bring back, in ONE gpurun call (~8 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/gpu_first_call.sh'
# 1. the GPU parity suite                        -> gpurun_out/gpu_tests.log
# 2. bench.py, default arguments                 -> gpurun_out/bench.json      (a number under ncu is never a bench value)
# 3. ncu launch list of a short bench            -> gpurun_out/launches.csv    (shares of the step, cold-cache)
# 4. ncu --set full of the top kernels + source  -> gpurun_out/raster_fast.ncu-rep, lighting.ncu-rep, frame_misc.ncu-rep
# Read the reports here with: ncu -i <rep> --page details | --page raw --csv | --page source --csv --print-source cuda,sass
# and profiles/tools/ncu_source_lines.py for per-line / per-phase instruction and stall shares.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1

timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log

timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1
echo "smoke under ncu exit $?" >> gpurun_out/smoke.log

timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
    f = d.get('frame', {})
    print('value', round(d['value']), 'Mtri/s', 'e2e', round(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 3),
          '| frame fps', round(f.get('frames_per_s', 0)), {k: round(v) for k, v in f.get('pass_us', {}).items()})
except Exception as e:
    print('bench line unreadable:', e)
PY

# launch list: raster-only S5 (eager launches so that every kernel is its own row) + one pass over the city frame
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-frame --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/launches_frame.csv \
    python bench.py --frame-only --steps 2 > /dev/null 2>&1

# full captures, one launch each, with source correlation (-lineinfo is in the build flags)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_raster_fast -s 4 -c 1 -f -o gpurun_out/raster_fast \
    python bench.py --steps 2 --warmup 1 --no-frame --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lighting -s 3 -c 1 -f -o gpurun_out/lighting \
    python bench.py --frame-only --steps 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_raster_deferred|k_raster_large|k_hzb_build|k_tile_marker|k_group_cull' -s 15 -c 5 -f -o gpurun_out/frame_misc \
    python bench.py --frame-only --steps 2 > /dev/null 2>&1
timeout 120 python tests/pcie_probe.py > gpurun_out/pcie.json 2>&1; cat gpurun_out/pcie.json
ls -la gpurun_out
