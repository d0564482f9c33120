"""The per-thread resolve kernels of chord_b200/csrc/resolve.cu, compiled from their own source text for the host and run
thread by thread on the CPU against the oracle.

TEST INFRASTRUCTURE, not a fallback: tests/hostcc/ holds stand-ins for <cuda_runtime.h> / <cuda_fp16.h> under which g++ can
compile the DEVICE FUNCTIONS of a .cu file; this test cuts the launchers out of resolve.cu, builds a scratch library in a
temporary directory and calls it.  The product never sees that library (libchord_b200.so is sm_100a nvcc output and fails
loudly without a GPU).  What a host run proves: the algebra of the kernel and, bit for bit, every chain the numerics
contract pins to IEEE operations - here the barycentrics -> position -> Bruneton coordinate chain;
the approximate intrinsics (__fdividef, rsqrtf) are evaluated exactly on the host, so their last bits remain the GPU
tests' business (tests/test_gpu_parity.py::test_lighting_resolve, ::test_lighting_packed_formats)."""
import ctypes as C
import hashlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
from chord_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTCC = os.path.join(ROOT, 'tests', 'hostcc')


@pytest.fixture(scope='module')
def hostlib(tmp_path_factory):
    d = tmp_path_factory.mktemp('hostcc')
    src = open(os.path.join(ROOT, 'chord_b200', 'csrc', 'resolve.cu')).read()
    src, n = re.subn(r'^int launch_\w+\([^)]*\)\n\{.*?^\}\n', '', src, flags=re.S | re.M)
    src, n2 = re.subn(r'^template <bool kPacked, int kSlot>\nstatic void lighting_launch\([^)]*\)\n\{.*?^\}\n', '', src, flags=re.S | re.M)
    src, n3 = re.subn(r'^template <bool kPacked, int kSlot>\nstatic void sky_launch\([^)]*\)\n\{.*?^\}\n', '', src, flags=re.S | re.M)
    assert n == 2 and n2 == 1 and n3 == 1 and '<<<' not in src, 'expected to cut exactly the launchers out of resolve.cu'
    src = src.replace('#include "common.cuh"', '')
    open(d / 'resolve_device.inc', 'w').write(src)
    so = str(d / 'libhostcc_resolve.so')
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    # -ffp-contract=off: only the explicit __fmaf_rn calls fuse, as under nvcc --fmad=false; -mfma makes those one instruction
    subprocess.check_call([cxx, '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off', '-mfma', '-mf16c', '-fno-fast-math',
                           '-I', HOSTCC, '-I', os.path.join(ROOT, 'chord_b200', 'csrc'), '-I', str(d),
                           os.path.join(HOSTCC, 'lighting_harness.cpp'), os.path.join(HOSTCC, 'hostcc_rt.cpp'), '-o', so])
    lib = C.CDLL(so)
    lib.hostcc_div_uniform_mismatches.restype = C.c_ulonglong
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _host_lighting(lib, scene, f, w, h, vis64, post, tr, ir, packed):
    sizes = [8, 4, 4, 4, 4, 4]
    planes = [np.zeros((h, w, s), dtype=np.uint8) for s in sizes] if packed else [np.zeros((h, w, 4), dtype=np.float32) for _ in range(6)]
    arr = (C.c_void_p * 6)(*[p.ctypes.data for p in planes])
    vis = np.ascontiguousarray(vis64, dtype=np.uint64)
    post = np.ascontiguousarray(post)
    objs = np.ascontiguousarray(f['objects'])
    t16, i16 = np.ascontiguousarray(tr, dtype=np.float16), np.ascontiguousarray(ir, dtype=np.float16)
    view, iview = np.ascontiguousarray(f['view']), np.ascontiguousarray(f['iview'])
    rc = lib.hostcc_lighting(C.byref(scene.desc), _p(view), _p(iview), _p(objs), C.c_uint32(len(objs)), C.c_uint32(w), C.c_uint32(h),
                             _p(vis), _p(post), C.c_uint32(len(post)), _p(t16), _p(i16), C.c_int(int(packed)), arr)
    assert rc == 0
    return planes


# sha256 over the six fp32 planes + the six packed planes of the host run, per frame.  Round 2 re-baselined them: k_lighting now
# restates the oracle operation for operation (every fp32 plane equals the oracle's), so these digests are the ORACLE's bits plus
# their quantisation into the reference's storage formats.  A deliberate change of the contract has to re-baseline them (and say so).
HOST_RUN_SHA256 = ['85658bfbed77995dbde9c3602351a1cd6fb3f5c9787d08a0e6fbe1ef5d940fe4', '86e6081f834e0c6aedbe4f12315e22a116805b3faabb95a9161a0b2e6944c861']      # (normalize = v * rcp(sqrt(dot)) contract)


def test_lighting_kernel_source_matches_oracle(hostlib):
    """Same case and same assertions as the GPU test (tests/test_gpu_parity.py::test_lighting_resolve), plus: the five
    G-buffer planes that depend only on IEEE-pinned arithmetic or exact copies must equal the oracle bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from cases import city_frames
    w, h = 640, 360
    scene, info, frames = city_frames(frames=2, w=w, h=h)
    tr, ir = synth.synthetic_luts()
    prev, digests = None, []
    for f in frames:
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
        g = oracle.lighting(ref['frame'], ref['vis64'], ref['post_cull'], tr, ir)
        got = _host_lighting(hostlib, scene, f, w, h, ref['vis64'], ref['post_cull'], tr, ir, packed=False)
        names = ['color', 'baseColor', 'vertexNormal', 'pixelNormal', 'motion', 'aoRoughMetal']
        # round 2: the kernel is the oracle's arithmetic operation for operation -> every plane EQUAL (round 1: fused / fast
        # normalisations, colour within 1.5e-7 relative and single pixels up to 3.5e-5 at GGX peaks)
        for i, name in enumerate(names):
            assert np.array_equal(got[i], g[i], equal_nan=True), f'{name}: max |difference| {np.nanmax(np.abs(got[i] - g[i]))}'
        assert g[0].max() > 0.05
        prev = ref['hzb']

        # packed storage formats: same checks as test_lighting_packed_formats
        pk = _host_lighting(hostlib, scene, f, w, h, ref['vis64'], ref['post_cull'], tr, ir, packed=True)

        def half_ulps(a, b16):
            a16 = a.astype(np.float16).view(np.int16).astype(np.int32)
            return np.abs(a16 - b16.view(np.int16).astype(np.int32)).max()
        assert half_ulps(g[0], pk[0].view(np.float16).reshape(h, w, 4)) <= 1
        assert half_ulps(g[4][..., :2], pk[4].view(np.float16).reshape(h, w, 2)) <= 1
        for which in (1, 2, 3):
            p = pk[which].view(np.uint32).reshape(h, w)
            want = np.rint(np.clip(g[which][..., :3], 0, 1) * 1023.0).astype(np.int32)
            gotc = np.stack([p & 0x3FF, (p >> 10) & 0x3FF, (p >> 20) & 0x3FF], axis=-1).astype(np.int32)
            assert np.abs(gotc - want).max() <= 1, which
        p8 = pk[5].reshape(h, w, 4).astype(np.int32)
        assert np.abs(p8 - np.rint(np.clip(g[5], 0, 1) * 255.0).astype(np.int32)).max() <= 1
        digest = hashlib.sha256(b''.join(np.ascontiguousarray(x).tobytes() for x in got + pk)).hexdigest()
        digests.append(digest)
    assert digests == HOST_RUN_SHA256, f'host run of the kernel source changed bits: {digests}'


@pytest.mark.parametrize('w', [1000.0, 60.0, float(np.sqrt(np.float32(np.float32(6420.0) ** 2 - np.float32(6360.0) ** 2))), 3.0, 0.1])
def test_div_uniform_is_the_ieee_division(hostlib, w):
    """div_uniform (resolve.cu): x * RN(1/w) corrected by the exact residual must be x / w for EVERY finite numerator whose
    quotient stays normal - here: all positive floats in [2^-60, 2^60] for the denominators of the default atmosphere
    (1000 m per km, top - bottom = 60 km, H = sqrt(top^2 - bottom^2)) and two awkward ones."""
    lo, hi = np.array([2.0 ** -60, 2.0 ** 60], dtype=np.float32).view(np.uint32)
    bad = C.c_uint32(0)
    n = hostlib.hostcc_div_uniform_mismatches(C.c_float(w), C.c_uint32(int(lo)), C.c_uint32(int(hi)), C.byref(bad))
    assert n == 0, f'{n} numerators differ for w = {w}, first bits {bad.value:#010x}'
