#!/usr/bin/env python
"""Single-GPU prediction of the S5 tile-parallel scaling (diagnostic, not a bench value): for N in 1, 2, 4, 8 every
rank's screen tile is rendered in turn on cuda:0 exactly as bench.py's rank would do it (tile sub-frustum cull outside
the timed region, then begin + render_mesh per step) and timed with CUDA events.  Predicted speed-up at N = t(1) / max
over ranks t(N): what is lost here is lost to per-step fixed costs and to meshlets straddling tile borders, before any
cross-GPU effect."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chord_b200 import camera as cam, multigpu, runtime, synth_device  # noqa: E402

W, H, M = 3840, 2160, 781250


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    ctx = runtime.Context(0)
    torch.cuda.set_device(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    scene, camera, objects, cmds = synth_device.patch_stream_scene_device(M, W, H, seed=5, device='cuda:0')
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    ctx.upload_scene(scene)
    frame = runtime.Frame(ctx, W, H)
    groups = (M + 3) // 4
    out = {}
    for n in (1, 2, 4, 8):
        per_rank = []
        for rank in range(n):
            tile = multigpu.tile_for_rank(rank, n, W, H)
            lst = runtime.CmdList(ctx, M)
            if n == 1:
                lst.upload(cmds)
                iv, tl = iview, None
            else:
                iv, tl = multigpu.tile_instance_view(view, camera, tile), tile
                frame.begin(view, iv, objects, M, groups, tl)
                frame.instance_culling(lst)
                ctx.sync()
            n_local = lst.count()
            for _ in range(4):
                frame.begin(view, iv, objects, M, groups, tl)
                frame.render_mesh(lst)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            e0.record(stream)
            for i in range(steps):
                frame.begin(view, iv, objects, M, groups, tl)
                kev[i][0].record(stream)
                frame.render_mesh(lst)
                kev[i][1].record(stream)
            e1.record(stream)
            torch.cuda.synchronize()
            with runtime.Graph.capture(ctx) as g:
                frame.begin(view, iv, objects, M, groups, tl)
                frame.render_mesh(lst)
            for _ in range(3):
                g.launch()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            for i in range(steps):
                g.launch()
            g1.record(stream)
            torch.cuda.synchronize()
            per_rank.append({'rank': rank, 'tile': list(tile), 'meshlets': int(n_local), 'ms_per_step': e0.elapsed_time(e1) / steps,
                             'render_mesh_ms': float(np.mean([a.elapsed_time(b) for a, b in kev])),
                             'graph_ms_per_step': g0.elapsed_time(g1) / steps, 'graph_kernel_nodes': g.kernel_nodes})
            g.close()
            lst.close()
        out[n] = per_rank
    t1 = out[1][0]['ms_per_step']
    summary = {n: {'max_ms_per_step': max(r['ms_per_step'] for r in out[n]), 'max_render_mesh_ms': max(r['render_mesh_ms'] for r in out[n]),
                   'sum_meshlets': sum(r['meshlets'] for r in out[n]),
                   'predicted_speedup': t1 / max(r['ms_per_step'] for r in out[n]),
                   'max_graph_ms_per_step': max(r['graph_ms_per_step'] for r in out[n]),
                   'predicted_speedup_graph': out[1][0]['graph_ms_per_step'] / max(r['graph_ms_per_step'] for r in out[n])} for n in out}
    print(json.dumps({'summary': summary, 'per_rank': out}), flush=True)


if __name__ == '__main__':
    main()
