#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 visibility-buffer path (contract: see the task statement / DESIGN.md §6).

Workload (BASELINE.json configs[4], the one the >=10 G tri/s target is quoted on): S5 = synthetic meshlet stream,
781 250 meshlets x 128 triangles (V = 81) = 100 M ~1 px^2 triangles at 3840x2160, raster only.
One "step" = clear the packed visibility surface + software-rasterise the whole draw list once.

  value      Mtri/s, inputs resident in HBM, CUDA events on the library's stream, max over ranks
  e2e        same metric through the C ABI with HOST buffers: per step the draw list + uniforms are copied H2D from
             pinned memory and the R32_UINT visibility image is read back D2H, all inside the timed region
  roofline   algorithmic bytes of the raster kernel (1 884 B / meshlet + 8 B / pixel) / its event-timed duration
  cpu_baseline  the CPU oracle (restatement of chord's path; the reference itself is Win32/DXC/Vulkan-only) on a
             bounded sample of the same draw list, all host threads
Multi-GPU (--gpus N under torchrun): the frame is split into N screen tiles (strong scaling); every rank culls the
stream against its tile's sub-frustum and rasterises only its tile; no data-path collective is needed for the
raster-only workload (the HZB all-gather belongs to the full-frame path, see chord_b200/multigpu.py).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
N_MESHLETS = 781250
TRIS_PER_MESHLET = 128
V_PER_MESHLET = 81
WORKLOAD = ('S5 synthetic meshlet stream: 781250 meshlets x 128 tri (V=81), ~1 px^2 triangles, 3840x2160, raster only '
            '(clear + k_raster_prepare/fast/deferred/large)')
DTYPE = 'f32 setup + i32/i64 fixed-point edge functions, u64 atomics'
BYTES_PER_MESHLET = 12 + 64 + 4 * V_PER_MESHLET + 4 * TRIS_PER_MESHLET + 12 * V_PER_MESHLET     # 1884


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='chord_b200', choices=['chord_b200', 'reference'])
    ap.add_argument('--meshlets', type=int, default=N_MESHLETS, help='override the stream length (debug only)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-frame', action='store_true', help='skip the secondary full-frame (S3 stand-in) measurement')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of the recorded step (CUDA graph replay)')
    ap.add_argument('--frame-only', action='store_true', help='profiling aid: run only the full-frame workload and print its dict')
    ap.add_argument('--frame-eager', action='store_true', help='profiling aid: the frame workloads without CUDA graphs (every kernel its own launch)')
    ap.add_argument('--raster-variant', type=int, default=-1, help='tuning aid: chord_ctx_set_raster_tuning variant (0, 11, 13)')
    ap.add_argument('--raster-ctas', type=int, default=-1, help='tuning aid: CTAs per SM of the meshlet rasteriser')
    ap.add_argument('--raster-warp-min', type=int, default=-1, help='tuning aid: list length below which the CTA-per-meshlet kernel is used')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def tile_for_rank(rank, n, w, h):
    from chord_b200 import multigpu
    return multigpu.tile_for_rank(rank, n, w, h)


def run_reference(args):
    """--impl reference: chord itself cannot run here (MSVC/Win32/DXC/Vulkan-only, no lavapipe), so this arm times the
    CPU restatement of its path (the oracle, kind 'port') on all host threads, on a bounded sample of the workload."""
    import oracle
    from chord_b200 import camera as cam, synth
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # the WHOLE stream of the chord_b200 arm's config (781 250 meshlets = 100 M triangles per step), not a head of it: the two
    # arms time the same work per step.  ~0.5 s per step on 16 host threads.
    sample = args.meshlets
    scene, camera, objects, cmds = synth.patch_stream_scene(sample, W4K, H4K, seed=5)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    fr = oracle.Frame(scene, view, iview, objects, W4K, H4K, threads=cores)
    vis = np.zeros((H4K, W4K), dtype=np.uint64)

    def step():
        # one step of the workload = clear the packed surface + rasterise the draw list (same definition as the other arm;
        # the 66 MB clear is ~2 % of a CPU step)
        vis.fill(0)
        fr.render_mesh(cmds, vis)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    mtri = sample * TRIS_PER_MESHLET / dt / 1e6
    line = {
        'impl': 'reference', 'metric': 'Mtri/s software-rasterized', 'value': mtri, 'unit': 'Mtri/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': DTYPE, 'data': 'synthetic',
        'config': {'workload': WORKLOAD if sample == N_MESHLETS else WORKLOAD.replace('781250', str(sample)),
                   'triangles_per_step': sample * TRIS_PER_MESHLET,
                   'sample': f'the whole stream every step: {sample} meshlets = {sample * TRIS_PER_MESHLET} triangles (same work per step as the chord_b200 arm)'},
        'cpu_baseline': {'value': mtri, 'unit': 'Mtri/s', 'cores': cores, 'kind': 'port',
                         'sample': f'all {sample} meshlets x 128 tri per step, OpenMP over meshlets, {cores} threads'},
        'e2e': {'value': mtri, 'unit': 'Mtri/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'reference (qiutang98/chord) is MSVC/Win32/DXC/Vulkan-only and cannot be built or run offline; '
                'this is the CPU restatement of its raster path (oracle/chord_oracle.c)',
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(dev_scene, view, iview, objects, budget_s=12.0):
    """The oracle on the head of the very same stream (copied back from HBM), all host threads."""
    import oracle
    cores = os.cpu_count() or 1
    probe = min(4000, dev_scene.n_meshlets)
    hs = dev_scene.host_prefix(probe)
    fr = oracle.Frame(hs, view, iview, objects, W4K, H4K, threads=cores)
    vis = np.zeros((H4K, W4K), dtype=np.uint64)
    cmds = np.zeros(probe, dtype=__import__('chord_b200').abi.DrawCmd)
    cmds['meshletId'] = np.arange(probe); cmds['slot'] = np.arange(probe)
    fr.render_mesh(cmds, vis)
    t0 = time.perf_counter()
    fr.render_mesh(cmds, vis)
    per_meshlet = (time.perf_counter() - t0) / probe
    sample = int(min(dev_scene.n_meshlets, max(probe, budget_s / max(per_meshlet, 1e-9) / 3)))
    hs = dev_scene.host_prefix(sample)
    fr = oracle.Frame(hs, view, iview, objects, W4K, H4K, threads=cores)
    cmds = np.zeros(sample, dtype=cmds.dtype)
    cmds['meshletId'] = np.arange(sample); cmds['slot'] = np.arange(sample)
    best = None
    for _ in range(3):
        vis[:] = 0
        t0 = time.perf_counter()
        fr.render_mesh(cmds, vis)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {'value': sample * TRIS_PER_MESHLET / best / 1e6, 'unit': 'Mtri/s', 'cores': cores, 'kind': 'port',
            'sample': f'first {sample} meshlets ({sample * TRIS_PER_MESHLET} tri) of the same S5 stream at 3840x2160, best of 3, '
                      f'OpenMP over meshlets; CPU restatement of chord (reference not runnable offline)'}


CITY = dict(n_buildings=24, patches=16, lods=5, n_protos=6, err_scale=0.5, seed=3)     # S3 stand-in: 30 784 LOD-0 meshlets = 3.94 M triangles
CITY_S2 = dict(n_buildings=24, patches=4, lods=3, n_protos=6, err_scale=0.5, seed=2)    # S2 stand-in: 1 984 LOD-0 meshlets = 254 k triangles


def city_camera_path(camera, info, n, width, height, step=(0.15, 0.01, -0.6)):
    """Per-frame inputs of a camera path (built before any timed region): (camera, objects, view, full-frame instance view)."""
    from chord_b200 import camera as cam, synth
    frames, last_view, last_cam = [], None, None
    for fi in range(n):
        c = cam.Camera(camera.position + np.array(step) * fi, camera.front, width=width, height=height)
        objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam)
        view = cam.fill_view(c, last_view=last_view, frame_counter=fi)
        frames.append((c, objs, view, cam.fill_instance_view(view)))
        last_view, last_cam = view, c
    return frames


def pass_rooflines(pass_us, width, height, post_meshlets, stage_meshlets, avg_meshlet_bytes, peak_gbs):
    """Per-pass roofline of one frame: algorithmic bytes (SURVEY 8d / DESIGN section 4 figures) / measured GPU time."""
    px = width * height
    p2 = lambda v: 1 << (int(v) - 1).bit_length()
    chain = (4.0 / 3.0) * 2 * (p2(width) // 2) * (p2(height) // 2)                  # fp16 texels of one chain
    bytes_ = {
        'GLTF Instance Culling': 118 * post_meshlets,                                # 64 B meshlet + 10 B group + 12 B cmd + 32 B HZB texels per candidate
        'GLTF Visibility Stage0': avg_meshlet_bytes * stage_meshlets[0] + 8 * px,    # meshlet stream + the packed surface written once
        'BuildHZB Post Prepass Stage0': 8 * px + chain,
        'GLTF Visibility Stage1': avg_meshlet_bytes * stage_meshlets[1],
        'BuildHZB': 8 * px + 2 * chain,
        'VisibilityTileMarker': 8 * px + 16 * (px // 64),
        'lightingTile': (8 + 28) * px,                                               # + ~208 B per distinct triangle (not counted: lower bound)
    }
    out = []
    for k, us in pass_us.items():
        b = float(bytes_.get(k, 0))
        gbs = b / (us * 1e-6) / 1e9 if us > 0 else 0.0
        out.append({'pass': k, 'bytes': int(b), 'us': us, 'achieved_gbs': gbs, 'frac': gbs / peak_gbs})
    return out


def frame_benchmark(ctx, stream, torch, steps=20, width=W4K, height=H4K, city=None, use_graph=True, label='S3 stand-in (Bistro-sized)', extras=True):
    """BASELINE configs[2] (and, with CITY_S2 at 1080p, configs[1]): procedural city, full reference frame sequence (instance cull
    -> stage 0 on last frame's HZB -> HZB -> stage 1 -> HZB -> tile mark -> PBR resolve) along a camera path, through the
    NATIVE renderer (chord_renderer_render: one C-ABI call per frame; the sequence is recorded as two CUDA graphs - one per
    history slot - and replayed with the frame's uniforms rewritten).  Frame time = CUDA events around `steps` frames / steps
    (frames in flight overlap host and GPU like in a running renderer); per-pass GPU times come from the renderer's own
    timers at the reference's cut points (renderer.cpp:322-363), measured on a second, untimed pass over the same frames."""
    from chord_b200 import runtime, synth
    city = city or CITY
    scene, camera, objects, info = synth.city_scene(width=width, height=height, **city)
    ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    ctx.upload_luts(tr, ir)
    ctx.upload_scattering(synth.synthetic_scattering_luts()[0])      # the sky branch of the lighting pass (lighting.cpp:193-246 runs it every frame)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    WARM = 6                                 # eager frames + the graph captures of both history slots
    frames = city_camera_path(camera, info, steps + WARM, width, height)

    def timed(r):
        def run(fs):
            for (c, objs, view, iview) in fs:
                r.render(view, iview, objs, lod0, groups)
        run(frames[:WARM])
        _ = r.frame                          # (asynchronous resolve: orders the context stream after the last resolve)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        run(frames[WARM:])
        _ = r.frame
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps
    # headline frame time: two frames in flight, the resolve of frame N on its own stream under the visibility passes of N + 1
    ra = runtime.Renderer(ctx, width, height, resolve=True, packed_gbuffer=True, graph=use_graph, async_resolve=True)
    ms, wall_ms = timed(ra)
    replays_async = ra.graph_replays
    ra.close()
    # the same frames one pass after the other on one stream (the reference's single graphics queue), with the per-pass timers
    r = runtime.Renderer(ctx, width, height, resolve=True, packed_gbuffer=True, graph=use_graph, timers=True)
    ms_serial, _ = timed(r)
    # per-pass times: same frames again, one synchronising timer read per frame (untimed)
    acc = {k: 0.0 for k in runtime.PASS_LABELS}
    stage_meshlets = [0, 0]
    r.frame.reset_stats()
    for (c, objs, view, iview) in frames[WARM:]:
        r.render(view, iview, objs, lod0, groups)
        for k, v in r.pass_times().items():
            acc[k] += v
    n = len(frames) - WARM
    pass_us = {k: v / n for k, v in acc.items()}
    stats = r.frame.raster_stats()
    post_n = int(r.post_cull.count())
    d = scene.datas[0]
    vt = d['meshlets']['vertexTriangleCount']
    V, T = float((vt & 0xFF).mean()), float(((vt >> 8) & 0xFF).mean())
    avg_bytes = 12 + 64 + 4 * V + 4 * T + 12 * V
    peak = _peak_gbs()
    per_frame_meshlets = stats['meshletsRastered'] / n
    out = {'workload': f'{label}: procedural city, {lod0} LOD-0 meshlet instances = {lod0 * 128} triangles ({scene.total_meshlets} unique meshlets over '
                       f'{city["lods"]} LODs, {len(frames[0][1])} objects), {width}x{height}, LOD cut at 1 px error, two-phase HZB, HZB builds, tile classify + '
                       f'PBR + sky resolve (packed G-buffer), camera path',
           'submission': (('chord_renderer_render, recorded as CUDA graphs (visibility passes + resolve per history slot) and replayed with rewritten uniforms'
                           if use_graph else 'chord_renderer_render, eager launches')
                          + '; CHORD_RENDERER_ASYNC_RESOLVE: the resolve of frame N runs on a second stream under the visibility passes of frame N + 1'),
           'frames_per_s': 1000.0 / ms, 'ms_per_frame': ms, 'ms_per_frame_wall': wall_ms, 'frames_timed': steps, 'graph_replays': replays_async,
           'ms_per_frame_serial': ms_serial, 'frames_per_s_serial': 1000.0 / ms_serial,
           'serial_note': 'same frames, every pass after the other on one stream (timers renderer); pass_us below are measured on this run',
           'pass_us': pass_us, 'pass_us_sum': sum(pass_us.values()),
           'pass_roofline': pass_rooflines(pass_us, width, height, post_n, [per_frame_meshlets * 0.8, per_frame_meshlets * 0.2], avg_bytes, peak),
           'post_cull_meshlets_last_frame': post_n, 'raster_stats_total': stats}
    r.close()
    if extras:
        ex = {}
        try:
            ex['shadow'] = shadow_benchmark(ctx, stream, torch, (scene, info, frames[-1][2], frames[-1][1]))
        except Exception as e:
            ex['shadow'] = {'error': repr(e)}
        try:
            ex['masked'] = masked_benchmark(ctx, stream, torch)
        except Exception as e:
            ex['masked'] = {'error': repr(e)}
        out['next_rows'] = ex
    return out


def _peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6650.0))
    except Exception:
        return 6650.0


def shadow_benchmark(ctx, stream, torch, scene_pack, steps=8, dim=2048, cascades=4):
    """SURVEY 8f N3 measurement: renderShadow on the city (mesh_raster.cpp:332-546), every cascade re-rendered (no cache):
    per cascade instance cull with the orthographic view -> hzbCullingGeneric against the farther cascade -> depth-only
    raster (two-sided, depth clamp) -> HZB.  CUDA-event time of the whole pass."""
    from chord_b200 import camera as cam, renderer
    scene, info, view, objs = scene_pack
    cfg = cam.CascadeShadowMapConfig(cascadeCount=cascades, realtimeCascadeCount=2, cascadeDim=dim, cascadeEndDistance=80.0, farCascadeEndDistance=800.0)
    pool = renderer._Pool(ctx)
    light = (0.3, -0.8, 0.2)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    from chord_b200 import runtime
    targets = [runtime.Frame(ctx, dim, dim) for _ in range(cascades)]
    res = None
    times = []
    for i in range(steps + 2):
        if res is not None:
            for l in res.lists:
                pool.release(l)
        if i == steps + 1:
            for t in targets:
                t.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = renderer.renderShadow(ctx, pool, view, objs, lod0, groups, None, cfg, light, depth_targets=targets)
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    meshlets = [int(l.count()) for l in res.lists]
    stats = [f.raster_stats() for f in res.depths]
    return {'workload': f'cascade shadow depth: {cascades} cascades x {dim}^2 over the same city, all re-rendered per pass (host cascade setup inside the timed region)',
            'ms_per_pass': float(np.median(times)), 'meshlets_per_cascade': meshlets,
            'triangles_submitted_per_pass': int(sum(s_['trianglesSubmitted'] for s_ in stats))}


def masked_benchmark(ctx, stream, torch, n_meshlets=100000, steps=10):
    """SURVEY 8f N2 measurement: the S5-style patch stream drawn with an alpha-tested textured material (every fragment
    evaluates 3 perspective-correct uv samples + a trilinear A8 fetch before its RED): k_raster_masked throughput."""
    from chord_b200 import abi, camera as cam, runtime, synth
    scene, camera, objects, cmds = synth.patch_stream_scene(n_meshlets, W4K, H4K, seed=5)
    V = 81
    g = (np.arange(9, dtype=np.float32) / 8.0)
    gx, gy = np.meshgrid(g, g)
    uv = np.tile(np.stack([gx.reshape(-1), gy.reshape(-1)], axis=-1), (n_meshlets, 1)).astype(np.float32)
    d = scene.datas[0]
    m = synth.default_material(alpha_mode=1, alpha_cutoff=0.5)
    m['baseColorId'], m['baseColorSampler'] = 1, 0
    table = abi.TextureTable(synth.alpha_test_textures())
    scene = abi.HostScene([dict(meshlets=d['meshlets'], positions=d['positions'], normals=None, uv0=uv, tangents=None, meshletData=d['meshletData'],
                                groups=d['groups'], groupIndices=d['groupIndices'], bvhNodes=d['bvhNodes'])], scene.primitives, np.array([m]), textures=table)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    ctx.upload_scene(scene)
    fr = runtime.Frame(ctx, W4K, H4K)
    lst = runtime.CmdList(ctx, n_meshlets)
    lst.upload(cmds)
    groups = (n_meshlets + 3) // 4
    for _ in range(3):
        fr.begin(view, iview, objects, n_meshlets, groups)
        fr.render_mesh(lst)
    torch.cuda.synchronize()
    fr.reset_stats()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        fr.begin(view, iview, objects, n_meshlets, groups)
        a.record(stream)
        fr.render_mesh(lst)
        b.record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    st = fr.raster_stats()
    ctx.upload_textures(None)
    return {'workload': f'{n_meshlets} meshlets x 128 ~1 px^2 triangles, alpha-tested 64x64 leaf texture (repeat), 3840x2160, k_raster_masked',
            'mtri_per_s': n_meshlets * TRIS_PER_MESHLET / (ms * 1e-3) / 1e6, 'ms': ms,
            'fragments_alpha_killed_per_step': int(st['fragmentsAlphaKilled'] // steps)}


def frame_benchmark_tiled(ctx, stream, torch, dist, rank, world, steps=20, width=W4K, height=H4K, use_graph=True):
    """BASELINE configs[3]: the same city frame split into `world` screen tiles, one per GPU, through the native renderer: tile
    sub-frustum cull, scissored raster, tile-local phase-1 HZB, and the FUSED exchange of the history HZB - behind the final
    buildHZB a put kernel stores every rank's part of mips 0..5 straight into all ranks' chains over NVLink (peer pointers from
    CUDA IPC), arrival flags instead of a collective, top mips rebuilt in front of the next phase 0; the tiled frame is two
    CUDA-graph launches per rank (visibility passes; resolve on a second stream).  Measured beside it: the same frame with pack -> ncclAllGather -> unpack inside the library
    (chord_renderer_exchange_connect_nccl).  Frame time = max over ranks of (CUDA events around `steps` frames / steps).
    Parity (untimed): every rank also renders the FULL frame on its own GPU; its tile must equal the full frame inside the
    tile (depth bits + decoded ids), nothing outside it, and the gathered min / max chains + valid range must equal the
    single-GPU history bit for bit - all-reduced over the ranks into `parity`."""
    from chord_b200 import multigpu, runtime, synth
    scene, camera, objects, info = synth.city_scene(width=width, height=height, **CITY)
    ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    ctx.upload_luts(tr, ir)
    ctx.upload_scattering(synth.synthetic_scattering_luts()[0])      # the sky branch of the lighting pass (lighting.cpp:193-246 runs it every frame)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    WARM = 6
    frames = city_camera_path(camera, info, steps + WARM, width, height)
    tiles = multigpu.all_tiles(world, width, height)
    tile_iviews = [multigpu.tile_instance_view(view, c, tiles[rank]) for (c, objs, view, iview) in frames]      # hoisted out of the timed region
    dev = f'cuda:{ctx.device}'

    def timed(r):
        def run(lo, hi):
            for i in range(lo, hi):
                c, objs, view, _ = frames[i]
                r.render(view, tile_iviews[i], objs, lod0, groups)
        run(0, WARM)
        _ = r.frame
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        run(WARM, len(frames))
        _ = r.frame                                            # asynchronous resolve: the last resolve is inside the timed region
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / steps
        t = torch.tensor([e0.elapsed_time(e1) / steps, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t[0]), float(t[1])

    out = {'workload': f'S3 stand-in city ({lod0 * 128} LOD-0 triangles) at {width}x{height} split into {world} screen tiles (one per GPU), full frame sequence incl. PBR resolve',
           'tile_rank0': list(tiles[0])}
    # ---- fused exchange
    r = runtime.Renderer(ctx, width, height, resolve=True, packed_gbuffer=True, graph=use_graph, async_resolve=True)
    multigpu.connect_fused(r, rank, world, dist, tiles)
    ms, wall = timed(r)
    try:                                                       # the same without the overlap, for the record
        rs = runtime.Renderer(ctx, width, height, resolve=True, packed_gbuffer=True, graph=use_graph)
        multigpu.connect_fused(rs, rank, world, dist, tiles)
        out['ms_per_frame_serial'] = timed(rs)[0]
        rs.close()
    except Exception as e:
        out['ms_per_frame_serial'] = repr(e)
    timed_out = r.exchange_timed_out()
    out.update({'frames_per_s': 1000.0 / ms, 'ms_per_frame': ms, 'ms_per_frame_wall': wall, 'frames_timed': steps,
                'exchange': 'fused: k_hzb_exchange_put pushes the tile parts into every rank\'s history chain over NVLink peer memory (CUDA IPC) behind the HZB build, arrival flags, no collective call',
                'submission': ('two CUDA graph launches per rank and frame (visibility passes; resolve on a second stream, overlapping the next frame)' if use_graph else 'eager launches'),
                'graph_replays': r.graph_replays,
                'exchange_bytes_per_frame': int(2 * sum(r.history_hzb.tile_bytes(t_) for t_ in tiles) * (world - 1)),
                'exchange_timed_out': bool(timed_out)})
    # ---- where a rank's frame goes: per-pass GPU times of this rank (timers renderer, serial) and the residual of its frame
    # time - the wait for the slowest rank's arrival flags, graph launch gaps, uniform upload - max / min over the ranks
    try:
        rt = runtime.Renderer(ctx, width, height, resolve=True, packed_gbuffer=True, graph=use_graph, timers=True)
        multigpu.connect_fused(rt, rank, world, dist, tiles)
        ms_t, _ = timed(rt)
        acc = {k: 0.0 for k in runtime.PASS_LABELS}
        for i in range(WARM, len(frames)):
            c, objs, view, _ = frames[i]
            rt.render(view, tile_iviews[i], objs, lod0, groups)
            for k, v in rt.pass_times().items():
                acc[k] += v
        n = len(frames) - WARM
        mine = [acc[k] / n for k in runtime.PASS_LABELS]
        busy = sum(mine)
        t = torch.tensor(mine + [busy, ms_t * 1000.0 - busy], dtype=torch.float64, device=dev)
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        labels = list(runtime.PASS_LABELS) + ['gpu_busy_sum', 'residual (frame - busy: exchange wait, launch gaps, upload)']
        out['rank_pass_us'] = {'ms_per_frame_timers': ms_t, 'max_over_ranks': {k: float(v) for k, v in zip(labels, tmax.tolist())},
                               'min_over_ranks': {k: float(v) for k, v in zip(labels, tmin.tolist())},
                               'note': 'timers renderer (serial passes, fused exchange); BuildHZB includes the peer stores of the exchange; the per-frame timer '
                                       'read-back of this untimed pass keeps the ranks in lock step, the frame time is from the timed pass before it'}
        rt.close()
    except Exception as e:
        out['rank_pass_us'] = {'error': repr(e)}
    # ---- parity, untimed: tile vs full frame on this GPU, gathered history vs single-GPU history
    single = runtime.Renderer(ctx, width, height)
    r.clear_history()
    ok_img = ok_out = ok_hzb = True
    for i in range(3):
        c, objs, view, iview = frames[i]
        r.render(view, tile_iviews[i], objs, lod0, groups)
        single.render(view, iview, objs, lod0, groups)
        dist.barrier()                                         # (host-side only: the ranks stay in lock step for the read-backs)
        r.history_sync()
        x0, y0, x1, y1 = tiles[rank]
        vt, vs = r.frame.vis64(), single.frame.vis64()
        dt, ds = r.frame.decoded(r.post_cull), single.frame.decoded(single.post_cull)
        ok_img &= bool(np.array_equal(vt[y0:y1, x0:x1] >> np.uint64(32), vs[y0:y1, x0:x1] >> np.uint64(32)) and np.array_equal(dt[y0:y1, x0:x1], ds[y0:y1, x0:x1]))
        outside = vt.copy(); outside[y0:y1, x0:x1] = 0
        ok_out &= not bool(outside.any())
        hr, hs = r.history_hzb, single.history_hzb
        ok_hzb &= bool(np.array_equal(hr.readback_chain('min'), hs.readback_chain('min')) and np.array_equal(hr.readback_chain('max'), hs.readback_chain('max'))
                       and np.array_equal(hr.valid_range(), hs.valid_range()))
    flags = torch.tensor([int(ok_img), int(ok_out), int(ok_hzb)], dtype=torch.int32, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ti, to, th = (bool(int(x)) for x in flags.tolist())
    out['parity'] = {'result': 'pass' if (ti and to and th) else 'FAIL', 'tile_ok': ti, 'outside_clean': to, 'hzb_ok': th,
                     'what': 'untimed, 3 frames, every rank: its tile == the full frame rendered on the same GPU inside the tile (depth bits + decoded object / meshlet / triangle), '
                             'no texel outside the tile, gathered min / max history chains + valid range == the single-GPU history; all-reduced (min) over the ranks'}
    single.close(); r.close()
    # ---- the same frame with the NCCL all-gather inside the library, for comparison
    try:
        rn = runtime.Renderer(ctx, width, height, resolve=True, packed_gbuffer=True, graph=use_graph)
        multigpu.connect_nccl(rn, rank, world, dist, tiles)
        ms_n, wall_n = timed(rn)
        out['nccl_exchange'] = {'ms_per_frame': ms_n, 'ms_per_frame_wall': wall_n, 'frames_per_s': 1000.0 / ms_n,
                                'exchange': 'pack -> ncclAllGather (communicator created by chord_nccl_comm_create) -> unpack -> top rebuild, captured in the same graph'}
        rn.close()
    except Exception as e:
        out['nccl_exchange'] = {'error': repr(e)}
    return out


def heavy_frame_benchmark(ctx, stream, torch, dist, rank, world, dev_scene, view, camera, objects, M, groups, steps=10):
    """The S5 stream (100 M ~1 px^2 triangles) through the WHOLE visibility sequence instead of raster only: per frame the tile
    sub-frustum instance cull of all 781 250 meshlets, two-phase HZB culling against last frame's (exchanged) chain, raster,
    both HZB builds and, at N > 1, the fused exchange - one CUDA-graph launch per rank and frame.  (No resolve: the synthetic
    stream carries no normals.)  Strong scaling: the same frame split over N screen tiles."""
    from chord_b200 import camera as cam, multigpu, runtime
    tiles = multigpu.all_tiles(world, W4K, H4K)
    iview = multigpu.tile_instance_view(view, camera, tiles[rank]) if world > 1 else cam.fill_instance_view(view)
    r = runtime.Renderer(ctx, W4K, H4K, graph=True)
    if world > 1:
        multigpu.connect_fused(r, rank, world, dist, tiles)

    def run(n):
        for _ in range(n):
            r.render(view, iview, objects, M, groups)
    run(4)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=f'cuda:{ctx.device}')
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    n_post = int(r.post_cull.count())
    r.close()
    return {'workload': f'S5 stream through cull -> stage 0 (HZB phase 0) -> HZB -> stage 1 -> HZB (+ fused exchange at N > 1), {M} meshlets = {M * TRIS_PER_MESHLET} triangles, '
                        f'3840x2160, {world} screen tile(s), static camera', 'ms_per_frame': ms, 'frames_per_s': 1000.0 / ms,
            'mtri_per_s': M * TRIS_PER_MESHLET / (ms * 1e-3) / 1e6, 'post_cull_meshlets_rank0': n_post}


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    from chord_b200 import camera as cam, runtime, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    n_gpus = args.gpus
    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG is VERSION
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # anything NCCL / c10d prints while the communicator comes up goes to stderr, not into the JSON stream
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    assert n_gpus == world, f'--gpus {n_gpus} but WORLD_SIZE {world} (launch with torch.distributed.run)'

    ctx = runtime.Context(local_rank)
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    ctx.set_raster_tuning(args.raster_variant, args.raster_ctas, args.raster_warp_min)
    if args.frame_only:
        print(json.dumps(frame_benchmark(ctx, stream, torch, steps=args.steps, use_graph=not args.frame_eager)), flush=True)
        return
    M = args.meshlets
    from chord_b200 import synth_device
    scene, camera, objects, cmds = synth_device.patch_stream_scene_device(M, W4K, H4K, seed=5, device=f'cuda:{local_rank}')
    view = cam.fill_view(camera)
    tile = tile_for_rank(rank, n_gpus, W4K, H4K)
    iview = cam.fill_instance_view(view)
    ctx.upload_scene(scene)
    frame = runtime.Frame(ctx, W4K, H4K)
    groups = (M + 3) // 4

    if n_gpus > 1:
        # per-tile work list: cull the stream against the tile's sub-frustum (host supplies the planes, K2 does the test)
        from chord_b200 import multigpu
        iview_tile = multigpu.tile_instance_view(view, camera, tile)
        frame.begin(view, iview_tile, objects, M, groups, tile)
        lst = runtime.CmdList(ctx, M)
        frame.instance_culling(lst)
        ctx.sync()
        host_cmds = lst.readback()
        iview_used = iview_tile
    else:
        lst = runtime.CmdList(ctx, M)
        lst.upload(cmds)
        host_cmds = cmds
        iview_used = iview
    n_local = len(host_cmds)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        frame.begin(view, iview_used, objects, M, groups, tile if n_gpus > 1 else None)
        frame.render_mesh(lst)

    # pinned host staging for the e2e leg
    pin_cmds = torch.empty(max(n_local, 1) * 12, dtype=torch.uint8).pin_memory()
    pin_cmds.numpy()[:n_local * 12] = np.frombuffer(host_cmds.tobytes(), dtype=np.uint8)
    pin_vis = [torch.empty(W4K * H4K, dtype=torch.int32).pin_memory() for _ in range(2)]
    pin_view = torch.from_numpy(np.frombuffer(view.tobytes(), dtype=np.uint8).copy()).pin_memory()
    import ctypes as C
    L = runtime.lib()

    e2e_i = [0]

    lst_e2e = [lst, runtime.CmdList(ctx, M)]

    def step_e2e():
        # host buffers in, host image out, everything through the C ABI.  Both transfers are pipelined against the raster:
        # the draw list of step i+1 goes H2D (upload stream) while step i rasterises, the image of step i goes D2H (copy
        # stream, two pinned buffers) while step i+1 rasterises; every step still moves one full list up and one full
        # image down, and all copies complete inside the timed region (chord_frame_wait_readbacks + final sync)
        k = e2e_i[0]
        L.chord_cmdlist_upload_async(lst_e2e[(k + 1) & 1].h, C.c_void_p(pin_cmds.data_ptr()), C.c_uint32(n_local))
        frame.begin(view, iview_used, objects, M, groups, tile if n_gpus > 1 else None)
        frame.render_mesh(lst_e2e[k & 1])
        L.chord_readback_visibility32_async(frame.h, C.c_void_p(pin_vis[k & 1].data_ptr()))
        e2e_i[0] += 1

    # ---- resident (kernel) timing; clocks are sampled from the warm-up to the end of the e2e leg
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    # the step (uniform upload, clear, prepare, raster kernels) is recorded once and replayed: one launch per step instead
    # of ~10, which matters when the per-rank step shrinks to ~0.25 ms at 8 GPUs.  The reference records a command
    # buffer per frame as well (command_list.h).
    graph = None
    if not args.no_graph:
        with runtime.Graph.capture(ctx) as graph:
            step_resident()
        for _ in range(3):
            graph.launch()
    barrier()
    launches0 = ctx.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        if graph is not None:
            graph.launch()
        else:
            step_resident()
    ev[1].record(stream)
    barrier()
    launches = ctx.kernel_launches - launches0
    total_ms = ev[0].elapsed_time(ev[1])
    # duration of the raster call alone (roofline numerator): eager launches bracketed by events on the library's stream
    for i in range(args.steps):
        frame.begin(view, iview_used, objects, M, groups, tile if n_gpus > 1 else None)
        kev[i][0].record(stream)
        frame.render_mesh(lst)
        kev[i][1].record(stream)
    barrier()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    # ---- e2e timing
    L.chord_cmdlist_upload_async(lst_e2e[0].h, C.c_void_p(pin_cmds.data_ptr()), C.c_uint32(n_local))      # prologue of the upload pipeline
    for _ in range(2):
        step_e2e()
    L.chord_frame_wait_readbacks(frame.h)
    barrier()
    # sanity: the pipelined read-back delivers the same image as the synchronous one
    chk = np.empty(W4K * H4K, dtype=np.uint32)
    L.chord_readback_visibility32(frame.h, C.c_void_p(chk.ctypes.data))
    a_img = pin_vis[(e2e_i[0] - 1) & 1].numpy().view(np.uint32).reshape(H4K, W4K)[tile[1]:tile[3], tile[0]:tile[2]]
    assert np.array_equal(chk.reshape(H4K, W4K)[tile[1]:tile[3], tile[0]:tile[2]], a_img), 'async read-back differs'
    # Untimed probe: list upload and image read-back on ONE transfer stream (one PCIe transfer at a time) or on two (full
    # duplex)?  Two streams shorten the transfers but slow the raster they overlap; which one wins depends on how the raster
    # time compares with the transfer time, i.e. on the kernel version and on the box's PCIe link.  The timed leg below uses
    # the faster mode (chord_frame_set_readback_stream); both probe figures are reported.
    def e2e_probe(own):
        L.chord_frame_set_readback_stream(frame.h, C.c_int(own))
        for _ in range(3):
            step_e2e()
        L.chord_frame_wait_readbacks(frame.h)
        barrier()
        tp = time.perf_counter()
        for _ in range(10):
            step_e2e()
        L.chord_frame_wait_readbacks(frame.h)
        ctx.sync()
        return (time.perf_counter() - tp) / 10 * 1e3
    try:
        p0, p1 = e2e_probe(0), e2e_probe(1)
    except Exception as e:                      # the headline must not depend on the probe: stay on the one-stream default
        sys.stderr.write(f'e2e probe failed: {e!r}\n')
        p0, p1 = 0.0, float('inf')
    probe = torch.tensor([p0, min(p1, 1e30)], dtype=torch.float64, device=f'cuda:{local_rank}')
    if dist is not None:
        dist.all_reduce(probe, op=dist.ReduceOp.MAX)
    probe_ms = [float(x) for x in probe.tolist()]
    own_stream = 1 if probe_ms[1] < 0.97 * probe_ms[0] else 0          # switch only for a clear win
    try:
        L.chord_frame_set_readback_stream(frame.h, C.c_int(own_stream))
    except Exception as e:
        sys.stderr.write(f'chord_frame_set_readback_stream failed: {e!r}\n')
        own_stream = 0
    for _ in range(2):
        step_e2e()
    L.chord_frame_wait_readbacks(frame.h)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    L.chord_frame_wait_readbacks(frame.h)
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)      # sync read-backs: wall clock bounds it
    # the timed regions last only a few hundred ms (a 100 ms nvidia-smi period gives a handful of samples): keep the same
    # resident workload running for ~0.7 s more, untimed, so that the clock / throttle record is a median under load
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 0.7:
        for _ in range(20):
            step_resident()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks['note'] = 'sampled every 100 ms from warm-up through the timed regions plus a 0.7 s untimed continuation of the same workload'

    frame_heavy = None
    if not args.no_frame:
        try:
            frame_heavy = heavy_frame_benchmark(ctx, stream, torch, dist, rank, world, scene, view, camera, objects, M, groups)
        except Exception as e:
            frame_heavy = {'error': repr(e)}
    frame_tiled = None
    if dist is not None and not args.no_frame:
        try:
            frame_tiled = frame_benchmark_tiled(ctx, stream, torch, dist, rank, world, use_graph=not args.frame_eager)
        except Exception as e:
            frame_tiled = {'error': repr(e)}
    times = torch.tensor([total_ms, kern_ms, e2e_ms], dtype=torch.float64, device=f'cuda:{local_rank}')
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_ms = (float(x) for x in times.tolist())
    stats = frame.raster_stats()

    if rank == 0:
        tris = M * TRIS_PER_MESHLET
        ms_per_step = total_ms / args.steps
        value = tris / (ms_per_step * 1e-3) / 1e6
        e2e_value = tris / (e2e_ms / args.steps * 1e-3) / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        alg_bytes = n_local * BYTES_PER_MESHLET + 8 * (tile[2] - tile[0]) * (tile[3] - tile[1])
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel comes from an ncu --set full capture (a profiler cannot run inside the bench): it
        # is only quoted when the capture was made with THIS raster.cu (sha256 of the kernel source recorded with it) and is
        # scaled by the share of the stream this rank rasterises; otherwise null.
        traffic, traffic_note = None, 'no ncu capture of this kernel source on record (profiles/raster_ncu_summary.json)'
        ncu_extra = None
        try:
            import hashlib
            summ = json.load(open(os.path.join(ROOT, 'profiles', 'raster_ncu_summary.json')))
            sha = hashlib.sha256(open(os.path.join(ROOT, 'chord_b200', 'csrc', 'raster.cu'), 'rb').read()).hexdigest()
            if summ.get('raster_cu_sha256') == sha:
                traffic = int(summ['dram_bytes_per_launch'] * (n_local / float(summ.get('meshlets_per_launch', N_MESHLETS))))
                traffic_note = (f"ncu --set full capture of commit {summ.get('commit', '?')} (same raster.cu): dram__bytes_read.sum + dram__bytes_write.sum of one "
                                f"k_raster_fast launch over {summ.get('meshlets_per_launch', N_MESHLETS)} meshlets, scaled by this rank's {n_local} meshlets")
                # (ii) DRAM rate under ncu, (iii) L2 atomic unit vs its peak, and which limit binds - from the same capture
                ncu_extra = {'dram_gbs_under_ncu': summ['dram_bytes_per_launch'] / (summ['duration_ms_under_ncu'] * 1e-3) / 1e9,
                             'l2_red_sectors_per_launch': summ.get('l2_red_sectors'), 'l2_red_pct_of_peak_sustained': summ.get('l2_red_pct_of_peak_sustained'),
                             'issue_slots_busy_pct': summ.get('issue_slots_busy_pct'), 'warp_instructions_per_triangle': summ.get('warp_instructions_per_triangle'),
                             'binding_limit': summ.get('binding_limit')}
            else:
                traffic_note = 'stale: profiles/raster_ncu_summary.json was captured with another raster.cu (sha256 differs)'
        except Exception:
            pass
        line = {
            'metric': 'Mtri/s software-rasterized', 'value': value, 'unit': 'Mtri/s', 'n_gpus': n_gpus, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': DTYPE, 'data': 'synthetic',
            'config': {'workload': WORKLOAD if M == N_MESHLETS else WORKLOAD.replace('781250', str(M)),
                       'triangles_per_step': tris, 'l2': 'inputs (1.5 GB) exceed the 126 MB L2; no explicit flush',
                       'parallelism': 'single GPU' if n_gpus == 1 else f'{n_gpus} screen tiles, one per GPU, per-tile sub-frustum cull',
                       'submission': 'eager launches' if graph is None else f'recorded step replayed as one CUDA graph ({graph.kernel_nodes} kernel nodes + uniform upload + clears)'},
            'e2e': {'value': e2e_value, 'unit': 'Mtri/s', 'h2d_bytes_per_step': int(n_local * 12 + 1728 + 288 + 224),
                    'd2h_bytes_per_step': int((tile[2] - tile[0]) * (tile[3] - tile[1]) * 4), 'ms_per_step': e2e_ms / args.steps,
                    'transfer_streams': 2 if own_stream else 1, 'probe_ms_per_step': {'one_stream': probe_ms[0], 'two_streams': probe_ms[1]},
                    'note': 'per rank and step: its draw list H2D (pinned) + its screen tile of the R32_UINT image D2H, both pipelined against the raster (list of step i+1 up / image of step i-1 down while step i runs), all inside the timed region; an untimed probe picks one transfer stream or two (full duplex) first'},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_note': traffic_note, 'kernel': 'k_raster_fast (+ prepare / deferred / large, < 4 % of the step)', 'kernel_ms': kern_ms,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else '6650 GB/s (of fallback)',
                         'algorithmic_bytes_per_launch': int(alg_bytes), 'ncu': ncu_extra,
                         'note': 'at 14.7 B/tri the HBM roofline is ~450 G tri/s; the binding limits are per-triangle setup '
                                 'issue rate and L2 atomics (see profiles/)'},
            'raster_stats': stats,
        }
        if not args.no_cpu_baseline and n_gpus == 1:
            line['cpu_baseline'] = cpu_baseline(scene, view, iview, objects)
        if frame_heavy is not None:
            line['frame_heavy'] = frame_heavy
        if frame_tiled is not None:
            line['frame'] = frame_tiled
            if isinstance(frame_tiled.get('parity'), dict):
                line['parity'] = frame_tiled['parity']['result']
        if n_gpus == 1 and not args.no_frame:
            try:
                line['frame'] = frame_benchmark(ctx, stream, torch, use_graph=not args.frame_eager)
            except Exception as e:      # the headline metric must not depend on the secondary workload
                line['frame'] = {'error': repr(e)}
            try:
                line['frame_s2'] = frame_benchmark(ctx, stream, torch, width=1920, height=1080, city=CITY_S2, use_graph=not args.frame_eager,
                                                   label='S2 stand-in (Sponza-sized, BASELINE configs[1])', extras=False)
            except Exception as e:
                line['frame_s2'] = {'error': repr(e)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
