#!/bin/bash
# Tuning aid (not a test), run ON THE GPU BOX: time bench.py with every library under profiles/ab_libs/ (built by
# tests/ab_build_libs.sh) swapped in for chord_b200/libchord_b200.so, then restore the snapshot's own library.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/ab_bench_libs.sh'
# An older library lacks newer optional entry points (chord_frame_set_readback_stream): bench.py falls back by itself.
set -u
mkdir -p gpurun_out; : > gpurun_out/ab_libs.log
cp chord_b200/libchord_b200.so /tmp/libchord_b200_snapshot.so
for lib in profiles/ab_libs/libchord_b200_*.so; do
  echo "== $lib" >> gpurun_out/ab_libs.log
  cp "$lib" chord_b200/libchord_b200.so
  timeout 400 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>> gpurun_out/ab_libs.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); f = d.get('frame', {})
        print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'], 4),
              '| frame ms', round(f.get('ms_per_frame', 0), 4), {k: round(v, 1) for k, v in f.get('pass_us', {}).items()})" >> gpurun_out/ab_libs.log
done
cp /tmp/libchord_b200_snapshot.so chord_b200/libchord_b200.so
cat gpurun_out/ab_libs.log
