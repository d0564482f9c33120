#!/bin/bash
# QA aid: compute-sanitizer over the GPU parity cases (memcheck: out-of-bounds / misaligned accesses incl. the TMA superset reads
# and the peer stores of the fused exchange; racecheck: shared-memory hazards of the warp-synchronous phases; synccheck: barrier
# misuse; initcheck: reads of uninitialised device memory).  The full-size cases (100 M triangles, 4K frames) are left out.
set -u
mkdir -p gpurun_out
SKIP='not full_size and not city_4k and not native_cpp and not extreme_frame and not random_camera and not selftest'
for tool in memcheck racecheck synccheck initcheck; do
    extra=''
    [ $tool = memcheck ] && extra='--leak-check no'
    timeout 1200 compute-sanitizer --tool $tool $extra --error-exitcode 7 --print-limit 20 --log-file gpurun_out/sanitizer_$tool.log \
        python -m pytest tests/test_gpu_parity.py -q -k "$SKIP" > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
    echo "$tool exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
    tail -2 gpurun_out/sanitizer_${tool}_pytest.log
done
