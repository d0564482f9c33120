#!/usr/bin/env python
"""Tuning aid (not a test): frame time of the 4K city on ONE GPU for the full frame and for one rank's tile of 8, per renderer
mode: visibility passes only / + resolve serial / + timers / asynchronous resolve.  Separates the cost of the passes from the
cost of measuring them.   python tests/profile_frame_modes.py"""
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
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    w, h = bench.W4K, bench.H4K
    scene, camera, objects, info = synth.city_scene(width=w, height=h, **bench.CITY)
    ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    ctx.upload_luts(tr, ir)
    ctx.upload_scattering(synth.synthetic_scattering_luts()[0])
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    steps, warm = 20, 6
    frames = bench.city_camera_path(camera, info, steps + warm, w, h)
    if '--ncu-tile' in sys.argv:
        # for `ncu --metrics gpu__time_duration.sum`: the tile's frames with eager launches, so that every kernel is a row
        tile = multigpu.all_tiles(8, w, h)[0]
        r = runtime.Renderer(ctx, w, h, resolve=True, packed_gbuffer=True, graph=False)
        r.set_tile(tile)
        for i in range(8):
            c, objs, view, iview = frames[i]
            r.render(view, multigpu.tile_instance_view(view, c, tile), objs, lod0, groups)
        _ = r.frame
        torch.cuda.synchronize()
        r.close()
        return
    out = {}
    for label, tile in (('full', None), ('tile_1_of_8', multigpu.all_tiles(8, w, h)[0])):
        ivs = [multigpu.tile_instance_view(view, c, tile) if tile else iview for (c, objs, view, iview) in frames]
        res = {}
        for mode, kw in (('visibility_only', dict()), ('resolve_serial', dict(resolve=True, packed_gbuffer=True)),
                         ('resolve_serial_timers', dict(resolve=True, packed_gbuffer=True, timers=True)),
                         ('resolve_async', dict(resolve=True, packed_gbuffer=True, async_resolve=True)),
                         ('resolve_serial_eager', dict(resolve=True, packed_gbuffer=True, graph=False))):
            kw = dict(dict(graph=True), **kw)
            r = runtime.Renderer(ctx, w, h, **kw)
            if tile:
                r.set_tile(tile)
            for i in range(warm):
                r.render(frames[i][2], ivs[i], frames[i][1], lod0, groups)
            _ = r.frame
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(warm, warm + steps):
                r.render(frames[i][2], ivs[i], frames[i][1], lod0, groups)
            _ = r.frame
            e1.record(stream)
            torch.cuda.synchronize()
            res[mode] = round(e0.elapsed_time(e1) / steps * 1000.0, 1)
            r.close()
        out[label] = res
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
