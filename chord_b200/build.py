"""In-tree build of libchord_b200.so (hand-written CUDA for sm_100a; no fallback architecture).

``python -m chord_b200.build`` or ``chord_b200.build.build()``.  nvcc cross-compiles without a GPU.
--fmad=false is part of the numerics contract: only explicit __fmaf_rn calls fuse.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libchord_b200.so')
SOURCES = ['ctx.cu', 'cull.cu', 'raster.cu', 'hzb.cu', 'resolve.cu', 'shadow.cu', 'renderer.cu']
HEADERS = ['common.cuh', '../../include/chord_b200.h', '../../include/chord_abi_types.h']

# Per-file overrides: none.  (Tried in round 1: resolve.cu with --use_fast_math halves the resolve but breaks the colour
# contract: the Bruneton radius r = |p - earthCentre| ~ 6360 km is ill-conditioned in fp32, only identical IEEE arithmetic on
# both sides agrees everywhere.)
PER_FILE_FLAGS = {}

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--fmad=false',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default', '--expt-relaxed-constexpr', '-shared',
]


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


BUILDER_SRC = os.path.join(CSRC, 'nanite_builder.cpp')
BUILDER_OUT = os.path.join(HERE, 'libchord_builder.so')


def build_builder(force=False):
    """chord_b200/libchord_builder.so: the meshlet-DAG builder's C++ core (host only, g++ -O2 -fopenmp; an asset-import tool, not
    part of the per-frame library)."""
    deps = [BUILDER_SRC, os.path.join(HERE, '..', 'include', 'chord_abi_types.h')]
    if not force and os.path.exists(BUILDER_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(BUILDER_OUT) for d in deps):
        return BUILDER_OUT
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.check_call([cxx, '-O2', '-std=c++17', '-fPIC', '-shared', '-fopenmp', '-Wall', '-Wno-misleading-indentation', '-o', BUILDER_OUT, BUILDER_SRC])
    return BUILDER_OUT


def build(force=False, verbose=False, extra=()):
    build_builder(force)
    if not force and not needs_build():
        return OUT
    objs = []
    nvcc = _nvcc()
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(bdir, s.replace('.cu', '.o'))
        flags = [f for f in NVCC_FLAGS if f != '-shared']
        if s in PER_FILE_FLAGS:
            flags = [f for f in flags if f != '--fmad=false'] + PER_FILE_FLAGS[s]
        cmd = [nvcc] + flags + list(extra) + ['-c', os.path.join(CSRC, s), '-o', o]
        if verbose:
            print(' '.join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0 or 'warning' in out):
            sys.stderr.write(f'--- {s}\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', OUT] + objs + ['-ldl']
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='-v' in sys.argv,
          extra=['-Xptxas', '-v'] if '--ptxas' in sys.argv else ())
    print(OUT)
