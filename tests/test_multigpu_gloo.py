"""N > 1 host logic on CPU: two gloo ranks split a frame into screen tiles, cull against their sub-frustum, render
their tile (oracle), exchange the packed HZB tile parts with ONE all-gather and rebuild the chain; the result must be
bit-identical to the single-rank frame.  The same orchestration runs over NCCL on the GPU box (multigpu.allgather_hzb)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle
    from cases import city_frames
    from chord_b200 import multigpu
    w, h = 640, 360
    scene, info, frames = city_frames(frames=2, w=w, h=h)
    tiles = multigpu.all_tiles(world, w, h)
    tile = tiles[rank]
    prev = None
    ok = True
    for f in frames:
        full = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], None)
        iv = multigpu.tile_instance_view(f['view'], f['camera'], tile)
        mine = oracle.run_frame(scene, f['view'], iv, f['objects'], w, h, info['lod0MeshletCount'], prev, tile=tile)
        # the tile-culled list is a subset of the full list, and renders the same pixels inside the tile
        assert len(mine['post_cull']) <= len(full['post_cull'])
        x0, y0, x1, y1 = tile
        a = oracle.decode(mine['vis64'], mine['post_cull'])[y0:y1, x0:x1]
        b = oracle.decode(full['vis64'], full['post_cull'])[y0:y1, x0:x1]
        ok &= np.array_equal(a, b)
        ok &= np.array_equal(mine['vis64'][y0:y1, x0:x1] >> np.uint64(32), full['vis64'][y0:y1, x0:x1] >> np.uint64(32))
        outside = mine['vis64'].copy()
        outside[y0:y1, x0:x1] = 0
        ok &= not outside.any()
        # one all-gather of the packed tile parts of mips 0..5, then local rebuild of mips 6..
        slot = max(multigpu.tile_texels(t, w, h) for t in tiles)
        for which, red in (('min', np.minimum), ('max', np.maximum)):
            send = torch.zeros(slot * 2, dtype=torch.uint8)            # fp16 bit patterns as bytes
            part = multigpu.pack_tile_np(mine['hzb'][which], tile, w, h)
            send[:part.size * 2] = torch.from_numpy(part.view(np.uint8))
            recv = [torch.zeros(slot * 2, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(recv, send)
            chain = np.zeros_like(mine['hzb'][which])
            for r, t in enumerate(tiles):
                n = multigpu.tile_texels(t, w, h)
                multigpu.unpack_tile_np(chain, t, recv[r].numpy().view(np.uint16)[:n], w, h)
            multigpu.finish_top_np(chain, w, h, red)
            ok &= np.array_equal(chain, full['hzb'][which])
            mine['hzb'][which] = chain
        prev = mine['hzb']          # full-frame history for the next frame's phase 0
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(os.path.join(out_dir, 'result'), 'w').write(str(int(res.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4, 8])
def test_tile_partition_and_hzb_allgather_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / 'result').read() == '1'


def test_tiles_cover_the_frame_exactly():
    sys.path.insert(0, ROOT)
    from chord_b200 import multigpu
    for (w, h) in ((3840, 2160), (1920, 1080), (640, 360), (256, 256)):
        for n in (1, 2, 4, 8):
            cover = np.zeros((h, w), dtype=np.int32)
            for t in multigpu.all_tiles(n, w, h):
                x0, y0, x1, y1 = t
                assert x0 % 64 == 0 and y0 % 64 == 0 and (x1 % 64 == 0 or x1 == w) and (y1 % 64 == 0 or y1 == h)
                cover[y0:y1, x0:x1] += 1
            assert (cover == 1).all(), (w, h, n)
            total = sum(multigpu.tile_texels(t, w, h) for t in multigpu.all_tiles(n, w, h))
            w0, h0, _, offs, _ = multigpu.hzb_geometry(w, h)
            assert total == sum((w0 >> m) * (h0 >> m) for m in range(6))
