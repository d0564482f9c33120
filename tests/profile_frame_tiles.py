#!/usr/bin/env python
"""Tuning aid (not a test): per-pass GPU times of ONE rank's screen tile of the 4K city frame, for world = 1, 2, 4, 8, measured on
one GPU (no exchange: the passes' own cost).  Says which pass keeps the tiled frame from scaling.
    python tests/profile_frame_tiles.py > gpurun_out/frame_tiles.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chord_b200 import multigpu, runtime, synth  # noqa: E402


def main():
    ctx = runtime.Context(0)
    w, h = bench.W4K, bench.H4K
    scene, camera, objects, info = synth.city_scene(width=w, height=h, **bench.CITY)
    ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    ctx.upload_luts(tr, ir)
    ctx.upload_scattering(synth.synthetic_scattering_luts()[0])
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    frames = bench.city_camera_path(camera, info, 14, w, h)
    out = {}
    for world in (1, 2, 4, 8):
        tiles = multigpu.all_tiles(world, w, h)
        per_tile = {}
        for ti in sorted({0, world // 2, world - 1}):
            r = runtime.Renderer(ctx, w, h, resolve=True, packed_gbuffer=True, graph=True, timers=True)
            if world > 1:
                r.set_tile(tiles[ti])
            acc = {k: 0.0 for k in runtime.PASS_LABELS}
            n = 0
            for i, (c, objs, view, iview) in enumerate(frames):
                iv = multigpu.tile_instance_view(view, c, tiles[ti]) if world > 1 else iview
                r.render(view, iv, objs, lod0, groups)
                if i >= 6:
                    for k, v in r.pass_times().items():
                        acc[k] += v
                    n += 1
            per_tile[str(tiles[ti])] = {k: round(v / n, 1) for k, v in acc.items()}
            per_tile[str(tiles[ti])]['sum'] = round(sum(acc.values()) / n, 1)
            per_tile[str(tiles[ti])]['post_cull'] = int(r.post_cull.count())
            r.close()
        out[f'world_{world}'] = per_tile
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
