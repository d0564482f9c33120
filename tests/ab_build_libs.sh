#!/bin/bash
# Tuning aid (not a test), run HERE (CPU container, has git + nvcc): build libchord_b200.so of several commits into
# profiles/ab_libs/ (git-ignored *.so, but they travel to the GPU box), for tests/ab_bench_libs.sh to time there.
#   bash tests/ab_build_libs.sh 42ed6d7 HEAD
# 42ed6d7 = the library that was timed on the B200 at the end of round 1 (67.0 G tri/s, 0.80 ms city frame).
set -eu
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/profiles/ab_libs"
for c in "$@"; do
  sha=$(git -C "$root" rev-parse --short "$c")
  tmp=$(mktemp -d)
  git -C "$root" archive "$c" chord_b200/csrc chord_b200/build.py include | tar -x -C "$tmp"
  touch "$tmp/chord_b200/__init__.py"
  (cd "$tmp" && python -c "import importlib.util, sys; s = importlib.util.spec_from_file_location('b', 'chord_b200/build.py'); m = importlib.util.module_from_spec(s); s.loader.exec_module(m); m.build(force=True)")
  cp "$tmp/chord_b200/libchord_b200.so" "$root/profiles/ab_libs/libchord_b200_$sha.so"
  rm -rf "$tmp"
  echo "built profiles/ab_libs/libchord_b200_$sha.so"
done
