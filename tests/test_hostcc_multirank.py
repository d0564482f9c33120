"""N > 1 on the kernels' own source text, without GPUs: two gloo ranks, each with the host-executor build of the library
(tests/hostcc/, test infrastructure - see test_hostcc_parity.py), split a camera path through the city into screen tiles.
Per frame and rank: tile sub-frustum cull + scissored raster + two-phase HZB culling on the gathered history + tiled HZB
build through the product's own host code (chord_b200/renderer.py, multigpu.tile_instance_view) and kernels; then the exchange
of chord_b200/multigpu.py::allgather_hzb - chord_hzb_pack_tile -> ONE all-gather -> chord_hzb_unpack_tiles ->
chord_hzb_finish_top - with gloo standing in for NCCL (the "device" memory of the host build is host memory, so CPU tensors
can be packed into directly).  The tile must equal the full frame rendered by the same rank inside the tile, nothing may be
written outside it, and the gathered HZB chain must equal the single-rank chain bit for bit: tests/run_multigpu_check.py,
which runs on 2 / 4 / 8 B200s over NCCL, on the CPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, so, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cases import city_frames
    from chord_b200 import multigpu, renderer, runtime
    runtime.LIB_PATH, runtime._lib = so, None            # the scratch library, in this process only
    w, h = 640, 360
    scene, info, frames = city_frames(frames=3, w=w, h=h)
    ctx = runtime.Context(0)
    ctx.upload_scene(scene)
    tiles = multigpu.all_tiles(world, w, h)
    tile = tiles[rank]
    tiled = renderer.DeferredRenderer(ctx, w, h)
    single = renderer.DeferredRenderer(ctx, w, h)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    slot = None
    ok = True
    for fi, f in enumerate(frames):
        iv = multigpu.tile_instance_view(f['view'], f['camera'], tile)
        post_t = tiled.render(f['camera'], f['objects'], lod0, groups, tile=tile, view=f['view'], iview=iv)
        hzb = tiled.history_hzb
        if slot is None:
            slot = (max(hzb.tile_bytes(t) for t in tiles) + 255) // 256 * 256
        # multigpu.allgather_hzb with gloo / host tensors
        send = torch.zeros(slot, dtype=torch.uint8)
        recv = torch.zeros(slot * world, dtype=torch.uint8)
        hzb.pack_tile('min', tile, send.data_ptr())
        parts = [recv[r * slot:(r + 1) * slot] for r in range(world)]
        dist.all_gather(parts, send)
        hzb.unpack_tiles('min', tiles, recv.data_ptr(), slot)
        tiled.frame.hzb_finish_top(hzb, True, False)
        hzb.set_valid(True)
        post_s = single.render(f['camera'], f['objects'], lod0, groups, view=f['view'], iview=f['iview'])
        x0, y0, x1, y1 = tile
        vt, vs = tiled.frame.vis64(), single.frame.vis64()
        dt, ds = tiled.frame.decoded(post_t), single.frame.decoded(post_s)
        same = (np.array_equal(vt[y0:y1, x0:x1] >> np.uint64(32), vs[y0:y1, x0:x1] >> np.uint64(32))
                and np.array_equal(dt[y0:y1, x0:x1], ds[y0:y1, x0:x1]))
        outside = vt.copy()
        outside[y0:y1, x0:x1] = 0
        chain_ok = np.array_equal(hzb.readback_chain('min'), single.history_hzb.readback_chain('min'))
        ok &= same and (not outside.any()) and chain_ok and post_t.count() <= post_s.count()
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(os.path.join(out_dir, 'result'), 'w').write(str(int(res.item())))
    dist.destroy_process_group()


@pytest.fixture(scope='module')
def scratch_lib(tmp_path_factory):
    sys.path.insert(0, os.path.join(HERE, 'hostcc'))
    import hostcc_build
    return hostcc_build.build(tmp_path_factory.mktemp('hostcc_lib_mr'))


@pytest.mark.parametrize('world', [2, 4])
def test_tiled_frames_on_host_executor_gloo(tmp_path, scratch_lib, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, scratch_lib, str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / 'result').read() == '1'
