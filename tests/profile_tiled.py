#!/usr/bin/env python
"""Phase timing of the tiled (multi-GPU) frame path; run under torchrun. Diagnostic only."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chord_b200 import camera as cam, multigpu, renderer, runtime, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    w, h = 3840, 2160
    scene, camera, objects, info = synth.city_scene(n_buildings=24, patches=16, lods=5, n_protos=6, err_scale=0.5, seed=3, width=w, height=h)
    ctx = runtime.Context(local)
    ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    ctx.upload_luts(tr, ir)
    t = multigpu.TileRenderer(ctx, w, h, rank, world, dist)
    t.r.do_resolve = True
    t.r.frame.set_gbuffer_packed(True)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    acc = {}

    def lap(name, t0):
        torch.cuda.synchronize()
        acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    last_cam = None
    for fi in range(12):
        c = cam.Camera(camera.position + np.array([0.15 * fi, 0.01 * fi, -0.6 * fi]), camera.front, width=w, height=h)
        t0 = time.perf_counter(); objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam); lap('host refill_objects', t0)
        t0 = time.perf_counter(); view = cam.fill_view(c, last_view=t.r.last_view, frame_counter=fi); lap('host fill_view', t0)
        t0 = time.perf_counter(); iview = multigpu.tile_instance_view(view, c, t.tile); lap('host tile_instance_view', t0)
        dist.barrier()
        t0 = time.perf_counter()
        post = t.r.render(c, objs, lod0, groups, tile=t.tile, view=view, iview=iview)
        lap('render (tile, incl. resolve)', t0)
        t0 = time.perf_counter()
        t._staging = multigpu.allgather_hzb(t.r.frame, t.r.history_hzb, rank, world, dist, ('min',), t._staging)
        lap('allgather_hzb', t0)
        last_cam = c
    for r_ in range(world):
        dist.barrier()
        if rank == r_:
            for k, v in acc.items():
                print(f'[rank {rank}] {k:32s} median {np.median(v[2:]):8.3f} ms   first {v[0]:8.3f} ms', flush=True)
            print(f'[rank {rank}] post-cull', post.count(), 'stats', t.r.frame.raster_stats(), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
