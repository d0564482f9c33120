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


CITY = dict(n_buildings=24, patches=16, lods=5, n_protos=6, err_scale=0.5, seed=3)     # 30 784 LOD-0 meshlets = 3.94 M triangles


def frame_benchmark(ctx, stream, torch, steps=20, width=W4K, height=H4K):
    """S3 stand-in (BASELINE configs[2]): procedural city, ~3.8 M LOD-0 triangles, full reference frame sequence
    (instance cull -> stage 0 on last frame's HZB -> HZB -> stage 1 -> HZB -> tile mark -> PBR resolve) along a camera
    path; per-pass CUDA-event times at the reference's timer cut points (renderer.cpp:322-363)."""
    from chord_b200 import camera as cam, renderer, synth
    scene, camera, objects, info = synth.city_scene(width=width, height=height, **CITY)
    ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    ctx.upload_luts(tr, ir)
    r = renderer.DeferredRenderer(ctx, width, height, do_resolve=True, packed_gbuffer=True)     # reference storage formats, 28 B/px
    labels = ['GLTF Instance Culling', 'GLTF Visibility Stage0', 'BuildHZB Post Prepass Stage0', 'GLTF Visibility Stage1',
              'BuildHZB', 'VisibilityTileMarker', 'lightingTile']
    acc = {k: 0.0 for k in labels}
    frames = []
    last_view, last_cam = None, None
    total = steps + 3
    for fi in range(total):
        c = cam.Camera(camera.position + np.array([0.15 * fi, 0.01 * fi, -0.6 * fi]), camera.front, width=width, height=height)
        objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam)
        view = cam.fill_view(c, last_view=last_view, frame_counter=fi)
        frames.append((c, objs, view, cam.fill_instance_view(view)))
        last_view, last_cam = view, c
    f = r.frame
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    t_frames = []
    for fi, (c, objs, view, iview) in enumerate(frames):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
        ev[0].record(stream)
        f.begin(view, iview, objs, lod0, groups)
        rc = renderer.GLTFRenderContext(f, len(objs), lod0, groups, r.pool)
        ev[1].record(stream)
        r.pool.release(r.post_cull)
        post = renderer.instanceCulling(rc); ev[2].record(stream)
        need1, rejected = renderer.gltfVisibilityRenderingStage0(rc, r.history_hzb, post); ev[3].record(stream)
        if need1:
            tmp = renderer.buildHZB(rc, True, False, False); ev[4].record(stream)
            renderer.gltfVisibilityRenderingStage1(rc, tmp, rejected); ev[5].record(stream)
            r.pool.release(tmp)
        else:
            ev[4].record(stream); ev[5].record(stream)
        r.pool.release(rejected)
        hzb = renderer.buildHZB(rc, True, True, True); ev[6].record(stream)
        renderer.visibilityMark(rc, post); ev[7].record(stream)
        renderer.lighting(rc, post); ev[8].record(stream)
        r.pool.release(r.history_hzb)
        r.history_hzb, r.last_view, r.post_cull = hzb, view, post
        torch.cuda.synchronize()
        if fi >= 3:
            for k, lab in enumerate(labels):
                acc[lab] += ev[k + 1].elapsed_time(ev[k + 2])
            t_frames.append(ev[0].elapsed_time(ev[8]))
    n = len(t_frames)
    ms = float(np.median(t_frames))
    stats = f.raster_stats()
    extras = {}
    try:
        extras['shadow'] = shadow_benchmark(ctx, stream, torch, (scene, info, frames[-1][2], frames[-1][1]))
    except Exception as e:
        extras['shadow'] = {'error': repr(e)}
    try:
        extras['masked'] = masked_benchmark(ctx, stream, torch)
    except Exception as e:
        extras['masked'] = {'error': repr(e)}
    return {'next_rows': extras, 'workload': f'S3 stand-in (Bistro-sized): procedural city, {lod0} LOD-0 meshlet instances = {lod0 * 128} triangles '
                        f'({scene.total_meshlets} unique meshlets over 5 LODs, {len(frames[0][1])} objects), {width}x{height}, '
                        f'LOD cut at 1 px error, two-phase HZB, HZB builds, tile classify + PBR resolve (packed G-buffer), camera path',
            'frames_per_s': 1000.0 / ms, 'ms_per_frame': ms, 'frames_timed': n,
            'pass_us': {k: 1000.0 * v / n for k, v in acc.items()},
            'post_cull_meshlets_last_frame': int(post.count()), 'raster_stats_total': stats}


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


def frame_benchmark_tiled(ctx, stream, torch, dist, rank, world, steps=20, width=W4K, height=H4K):
    """BASELINE configs[3]: the same city frame split into `world` screen tiles, one per GPU: tile sub-frustum cull, scissored
    raster, tile-local phase-1 HZB, ONE NCCL all-gather of the packed min-HZB tile parts per frame (history for the next
    frame's phase 0), local rebuild of mips 6.., tile-local resolve.  Frame time = max over ranks (device events)."""
    from chord_b200 import camera as cam, multigpu, renderer, synth
    scene, camera, objects, info = synth.city_scene(width=width, height=height, **CITY)
    ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    ctx.upload_luts(tr, ir)
    tr_ = multigpu.TileRenderer(ctx, width, height, rank, world, dist)
    tr_.r.do_resolve = True
    tr_.r.frame.set_gbuffer_packed(True)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    last_cam, times = None, []
    for fi in range(steps + 3):
        c = cam.Camera(camera.position + np.array([0.15 * fi, 0.01 * fi, -0.6 * fi]), camera.front, width=width, height=height)
        objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam)
        view = cam.fill_view(c, last_view=tr_.r.last_view, frame_counter=fi)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tr_.render(c, objs, lod0, groups, view=view)
        e1.record(stream)
        torch.cuda.synchronize()
        if fi >= 3:
            times.append(e0.elapsed_time(e1))
        last_cam = c
    # median over frames (NCCL sets up its channels lazily, the first collectives of a new size are outliers), max over ranks
    t = torch.tensor([float(np.median(times)), float(np.mean(times)), float(np.max(times))], dtype=torch.float64, device=f'cuda:{ctx.device}')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_mean, ms_max = (float(x) for x in t.tolist())
    return {'workload': f'S3 stand-in city ({lod0 * 128} LOD-0 triangles) at {width}x{height} split into {world} screen tiles (one per GPU), '
                        f'NCCL all-gather of the min-HZB tile parts per frame', 'frames_per_s': 1000.0 / ms, 'ms_per_frame': ms, 'ms_per_frame_mean': ms_mean, 'ms_per_frame_max': ms_max,
            'frames_timed': len(times), 'tile_rank0': list(tr_.tile), 'hzb_allgather_bytes': int(sum(tr_.r.history_hzb.tile_bytes(t_) for t_ in multigpu.all_tiles(world, width, height)))}


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

    if args.frame_only:
        print(json.dumps(frame_benchmark(ctx, stream, torch, steps=args.steps)), flush=True)
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

    frame_tiled = None
    if dist is not None and not args.no_frame:
        try:
            frame_tiled = frame_benchmark_tiled(ctx, stream, torch, dist, rank, world)
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
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'raster_ncu_summary.json'))).get('dram_bytes_per_launch')
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
                         'traffic': traffic, 'kernel': 'k_raster_fast (+ prepare / deferred / large, < 4 % of the step)', 'kernel_ms': kern_ms,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else '6650 GB/s (of fallback)',
                         'algorithmic_bytes_per_launch': int(alg_bytes),
                         'note': 'at 14.7 B/tri the HBM roofline is ~450 G tri/s; the binding limits are per-triangle setup '
                                 'issue rate and L2 atomics (see profiles/)'},
            'raster_stats': stats,
        }
        if not args.no_cpu_baseline and n_gpus == 1:
            line['cpu_baseline'] = cpu_baseline(scene, view, iview, objects)
        if frame_tiled is not None:
            line['frame'] = frame_tiled
        if n_gpus == 1 and not args.no_frame:
            try:
                line['frame'] = frame_benchmark(ctx, stream, torch)
            except Exception as e:      # the headline metric must not depend on the secondary workload
                line['frame'] = {'error': repr(e)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
