"""The GPU parity cases (tests/test_gpu_parity.py), executed WITHOUT a GPU on the source text of the CUDA kernels.

TEST INFRASTRUCTURE, not a fallback: tests/hostcc/ translates chord_b200/csrc/*.cu for g++ (kernel launches become a
coroutine-per-CUDA-thread executor, the inline-PTX helpers become what the instruction does) and builds a scratch library
with the same C ABI in a pytest temporary directory; for the duration of this module chord_b200.runtime is pointed at that
scratch library, afterwards at libchord_b200.so again.  The product never loads it and fails loudly without a GPU
(tests/test_abi.py::test_no_cpu_fallback_without_gpu).

What a pass here means: the kernels' algebra, their integer / fixed-point work, their queues, tickets, look-back scans and
barrier protocols produce the oracle's bits when every IEEE-pinned operation is IEEE.  What it does not cover: the GPU's
approximate instructions (replaced by the exact operation), the memory model, performance - the `-m gpu` run on a B200
remains the parity gate.  Its use: a kernel rewrite can be checked here in seconds before GPU minutes are spent on it.

The cheap cases run by default; CHORD_HOSTCC_FULL=1 adds the slow ones (overflow fallbacks, cascade shadow passes, scissor
tiles, medium / large triangle streams: ~4 more minutes)."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'hostcc'))
FULL = os.environ.get('CHORD_HOSTCC_FULL') == '1'
slow = pytest.mark.skipif(not FULL, reason='slow on the host executor: set CHORD_HOSTCC_FULL=1')


@pytest.fixture(scope='module')
def host_ctx(tmp_path_factory):
    import hostcc_build
    from chord_b200 import runtime
    so = hostcc_build.build(tmp_path_factory.mktemp('hostcc_lib'))
    saved = (runtime.LIB_PATH, runtime._lib)
    runtime.LIB_PATH, runtime._lib = so, None
    try:
        ctx = runtime.Context(0)
        yield ctx
        ctx.close()
    finally:
        runtime.LIB_PATH, runtime._lib = saved


@pytest.fixture(scope='module')
def cases():
    import test_gpu_parity
    return test_gpu_parity


def test_scratch_library_is_not_the_product(host_ctx):
    from chord_b200 import runtime
    assert 'hostcc_scratch' in runtime.LIB_PATH and not runtime.LIB_PATH.startswith(os.path.dirname(HERE) + '/chord_b200')
    # graphs are not modelled: the call reports it instead of pretending
    with pytest.raises(runtime.ChordError):
        runtime.lib().chord_graph_begin_capture(host_ctx.h)


def test_s1_single_meshlet(host_ctx, cases):
    cases.test_s1_single_meshlet(host_ctx)


def test_context_destroyed_before_its_children(host_ctx, cases):
    cases.test_context_destroyed_before_its_children(host_ctx)


def test_patch_stream_small_triangles(host_ctx, cases):
    cases.test_patch_stream_raster(host_ctx, 3000, 640, 360, 1.0)


@slow
@pytest.mark.parametrize('n,w,h,scale', [(1500, 512, 512, 40.0), (600, 256, 256, 2000.0)])
def test_patch_stream_medium_and_large_triangles(host_ctx, cases, n, w, h, scale):
    cases.test_patch_stream_raster(host_ctx, n, w, h, scale)


def test_ragged_meshlets(host_ctx, cases):
    cases.test_ragged_meshlets(host_ctx)


def test_city_frames(host_ctx, cases):
    cases.test_city_frames(host_ctx)


def test_hzb_culling_lists(host_ctx, cases):
    cases.test_hzb_culling_lists(host_ctx)


@slow
def test_tile_scissor(host_ctx, cases):
    cases.test_tile_scissor(host_ctx)


def test_empty_lists_publish_zero_counts(host_ctx, cases):
    cases.test_empty_lists_publish_zero_counts(host_ctx)


def test_extreme_frame_shapes(host_ctx, cases):
    cases.test_extreme_frame_shapes(host_ctx, **({} if FULL else {'shapes': ((64, 64), (4096, 64))}))


def test_native_renderer_camera_path(host_ctx, cases):
    cases.test_native_renderer_camera_path(host_ctx)


def test_many_objects_instance_culling(host_ctx, cases):
    cases.test_many_objects_instance_culling(host_ctx)


def test_native_renderer_async_resolve(host_ctx, cases):
    cases.test_native_renderer_async_resolve(host_ctx, False)


def test_renderer_fused_exchange_in_process(host_ctx, cases):
    cases.test_renderer_fused_exchange_in_process(host_ctx, world=2)
    if FULL:                                   # (four gloo ranks run on this build by default in test_hostcc_multirank.py)
        cases.test_renderer_fused_exchange_in_process(host_ctx, world=4)


def test_exchange_wait_gives_up_on_a_dead_peer(host_ctx, cases):
    cases.test_exchange_wait_gives_up_on_a_dead_peer(host_ctx)


def test_renderer_fused_exchange_async(host_ctx, cases):
    cases.test_renderer_fused_exchange_async(host_ctx, 2, False)


def test_tall_frame_hzb_rows(host_ctx, cases):
    cases.test_tall_frame_hzb_rows(host_ctx)


def test_readback_stream_modes(host_ctx, cases):
    cases.test_readback_stream_modes(host_ctx)


def test_hzb_chain_combinations(host_ctx, cases):
    cases.test_hzb_chain_combinations(host_ctx)


@pytest.mark.parametrize('w,h', [(264, 72), (72, 264), (200, 200)])
def test_tile_marker_odd_tile_counts(host_ctx, cases, w, h):
    cases.test_tile_marker_odd_tile_counts(host_ctx, w, h)


def test_lighting_resolve(host_ctx, cases):
    cases.test_lighting_resolve(host_ctx)


def test_lighting_textured_materials(host_ctx, cases):
    cases.test_lighting_textured_materials(host_ctx)


def test_sky_branch(host_ctx, cases):
    cases.test_sky_branch(host_ctx)


def test_lighting_packed_formats(host_ctx, cases):
    cases.test_lighting_packed_formats(host_ctx)


def test_bvh_traversal(host_ctx, cases):
    cases.test_bvh_traversal(host_ctx)


def test_builder_asset_parity(host_ctx, cases):
    cases.test_builder_asset_parity(host_ctx)


def test_masked_bucket(host_ctx, cases):
    cases.test_masked_bucket(host_ctx)


def test_cascade_setup_on_device(host_ctx, cases):
    cases.test_cascade_setup_on_device(host_ctx)


@slow
def test_queue_overflow_fallbacks(host_ctx, cases):
    cases.test_queue_overflow_fallbacks(host_ctx)


@slow
def test_cascade_shadow_depth(host_ctx, cases):
    cases.test_cascade_shadow_depth(host_ctx)


@slow
def test_kernels_under_ubsan(tmp_path_factory):
    """The same sources built with -fsanitize=undefined (no recovery): signed overflow in the int32 / int64 edge functions,
    out-of-range float -> int conversions, bad shifts would abort the process.  Runs a raster-heavy subset."""
    import subprocess
    import hostcc_build
    so = hostcc_build.build(tmp_path_factory.mktemp('hostcc_ubsan'), extra=('-fsanitize=undefined', '-fno-sanitize-recover=undefined'))
    script = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from chord_b200 import runtime\n"
        "runtime.LIB_PATH, runtime._lib = %r, None\n"
        "import test_gpu_parity as G\n"
        "ctx = runtime.Context(0)\n"
        "G.test_s1_single_meshlet(ctx); G.test_patch_stream_raster(ctx, 3000, 640, 360, 1.0); G.test_patch_stream_raster(ctx, 600, 256, 256, 2000.0)\n"
        "G.test_ragged_meshlets(ctx); G.test_city_frames(ctx); G.test_masked_bucket(ctx)\n"
        "print('ubsan-clean')\n") % (os.path.dirname(HERE), HERE, so)
    r = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and 'ubsan-clean' in r.stdout, r.stderr[-3000:]


@pytest.fixture(scope='module')
def host_demo(host_ctx, tmp_path_factory):
    """examples/chord_b200_demo.cpp (native C++ host over include/chord_b200.hpp) linked against the scratch library."""
    import subprocess
    from chord_b200 import runtime
    root = os.path.dirname(HERE)
    exe = str(tmp_path_factory.mktemp('hostcc_demo') / 'chord_b200_demo_hostcc')
    libdir = os.path.dirname(runtime.LIB_PATH)
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-I', os.path.join(root, 'include'), os.path.join(root, 'examples', 'chord_b200_demo.cpp'),
                           '-o', exe, '-L', libdir, '-l:' + os.path.basename(runtime.LIB_PATH), '-Wl,-rpath,' + libdir])
    return exe


def test_native_cpp_host(host_ctx, cases, host_demo, tmp_path):
    cases.test_native_cpp_host(host_ctx, tmp_path, exe=host_demo)


def test_native_cpp_renderer_camera_path(host_ctx, cases, host_demo, tmp_path):
    cases.test_native_cpp_renderer_camera_path(host_ctx, tmp_path, exe=host_demo)


@slow
def test_kernels_under_asan(tmp_path_factory):
    """The same sources built with -fsanitize=address (hostcc_rt.cpp announces every coroutine switch to the sanitizer): an
    out-of-bounds access of a kernel to "device" memory (heap blocks here), to its shared-memory arrays or to its stack aborts
    the process.  Covers the TMA superset reads of ragged meshlets (which rely on the 64 bytes of padding chord_scene_upload
    gives every array), the HZB edge gathers, the tile marker's clamp and both resolve paths."""
    import subprocess
    import hostcc_build
    so = hostcc_build.build(tmp_path_factory.mktemp('hostcc_asan'), extra=('-fsanitize=address', '-fno-omit-frame-pointer'))
    asan = subprocess.check_output(['gcc', '-print-file-name=libasan.so'], text=True).strip()
    script = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from chord_b200 import runtime\n"
        "runtime.LIB_PATH, runtime._lib = %r, None\n"
        "import test_gpu_parity as G\n"
        "ctx = runtime.Context(0)\n"
        "G.test_s1_single_meshlet(ctx); G.test_patch_stream_raster(ctx, 3000, 640, 360, 1.0); G.test_ragged_meshlets(ctx)\n"
        "G.test_city_frames(ctx); G.test_lighting_resolve(ctx); G.test_hzb_chain_combinations(ctx)\n"
        "for wh in ((264, 72), (72, 264)): G.test_tile_marker_odd_tile_counts(ctx, *wh)\n"
        "print('asan-clean')\n") % (os.path.dirname(HERE), HERE, so)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS='detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1')
    r = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, timeout=2400, env=env)
    assert r.returncode == 0 and 'asan-clean' in r.stdout, r.stderr[-3000:]


@slow
@pytest.mark.parametrize('order', ['reverse', 'random'])
def test_results_do_not_depend_on_thread_order(host_ctx, order):
    """The executor normally resumes a block's threads front to back; HOSTCC_ORDER runs them back to front / shuffled per round
    (threads still run from one synchronisation point to the next).  A kernel that reads what a sibling writes without a
    barrier / rendezvous between the two would now see a different value: every case must still produce the oracle's bits."""
    import subprocess
    from chord_b200 import runtime
    script = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from chord_b200 import runtime\n"
        "runtime.LIB_PATH, runtime._lib = %r, None\n"
        "import test_gpu_parity as G\n"
        "ctx = runtime.Context(0)\n"
        "G.test_s1_single_meshlet(ctx); G.test_patch_stream_raster(ctx, 3000, 640, 360, 1.0); G.test_ragged_meshlets(ctx)\n"
        "G.test_city_frames(ctx); G.test_hzb_culling_lists(ctx); G.test_lighting_resolve(ctx); G.test_bvh_traversal(ctx)\n"
        "G.test_masked_bucket(ctx); G.test_hzb_chain_combinations(ctx); G.test_cascade_setup_on_device(ctx)\n"
        "print('order-independent')\n") % (os.path.dirname(HERE), HERE, runtime.LIB_PATH)
    r = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, timeout=2400, env=dict(os.environ, HOSTCC_ORDER=order))
    assert r.returncode == 0 and 'order-independent' in r.stdout, r.stderr[-3000:]


@slow
def test_committed_golden_fixtures(host_ctx):
    """tests/golden/ (fixtures committed with their generating script) against the host run of the kernels: the body of
    tests/test_golden.py::test_cuda_matches_golden."""
    import test_golden
    test_golden.test_cuda_matches_golden(host_ctx)


def test_random_raster_cases(host_ctx, cases):
    cases.test_random_raster_cases(host_ctx, count=48 if FULL else 8)


def test_random_camera_frames(host_ctx, cases):
    cases.test_random_camera_frames(host_ctx, sequences=10 if FULL else 4)
