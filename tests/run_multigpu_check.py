#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on 2/4/8 B200s):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multigpu_check.py
Every rank renders its screen tile of a camera path through the city (tile sub-frustum cull, scissored raster, tiled
HZB, NCCL all-gather of the packed tile parts, local rebuild of the top mips) AND the full frame on its own; the tile
must equal the full frame inside the tile (depth bits + decoded (object, meshlet, triangle)) and the gathered HZB must
equal the single-GPU HZB bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from cases import city_frames  # noqa: E402
from chord_b200 import multigpu, renderer, runtime  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    w, h = 1920, 1080
    scene, info, frames = city_frames(n_buildings=16, w=w, h=h, frames=4)
    ctx = runtime.Context(local)
    ctx.upload_scene(scene)
    exchange = sys.argv[1] if len(sys.argv) > 1 else 'fused'
    tiled = multigpu.TileRenderer(ctx, w, h, rank, world, dist, exchange=exchange)
    single = renderer.DeferredRenderer(ctx, w, h)
    ok = True
    for fi, f in enumerate(frames):
        post_t = tiled.render(f['camera'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'], view=f['view'])
        post_s = single.render(f['camera'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'], view=f['view'], iview=f['iview'])
        dist.barrier()
        tiled.r.history_sync()
        x0, y0, x1, y1 = tiled.tile
        vt, vs = tiled.r.frame.vis64(), single.frame.vis64()
        dt, ds = tiled.r.frame.decoded(post_t), single.frame.decoded(post_s)
        same = np.array_equal(vt[y0:y1, x0:x1] >> np.uint64(32), vs[y0:y1, x0:x1] >> np.uint64(32)) and np.array_equal(dt[y0:y1, x0:x1], ds[y0:y1, x0:x1])
        outside = vt.copy(); outside[y0:y1, x0:x1] = 0
        ht, hs = tiled.r.history_hzb, single.history_hzb
        hz = (np.array_equal(ht.readback_chain('min'), hs.readback_chain('min')) and np.array_equal(ht.readback_chain('max'), hs.readback_chain('max'))
              and np.array_equal(ht.valid_range(), hs.valid_range()))
        print(f'[rank {rank}] frame {fi}: tile {tiled.tile} list {post_t.count()}/{post_s.count()} image_ok={same} outside_clean={not outside.any()} hzb_ok={hz}', flush=True)
        ok &= same and (not outside.any()) and hz
    res = torch.tensor([1 if ok else 0], device=f'cuda:{local}')
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('MULTIGPU_CHECK', 'PASS' if int(res.item()) == 1 else 'FAIL', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(res.item()) == 1 else 1)


if __name__ == '__main__':
    main()
