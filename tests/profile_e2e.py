#!/usr/bin/env python
"""Tuning aid (not a test): which part of bench.py's e2e step costs the ~0.1 ms it has over the resident step?  S5 on one GPU,
eager launches, CUDA events on the context stream: render only / + list upload / + image read-back / both (= the e2e step)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chord_b200 import camera as cam, runtime, synth_device  # noqa: E402


def main():
    ctx = runtime.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    M, W, H = bench.N_MESHLETS, bench.W4K, bench.H4K
    scene, camera, objects, cmds = synth_device.patch_stream_scene_device(M, W, H, seed=5, device='cuda:0')
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    ctx.upload_scene(scene)
    frame = runtime.Frame(ctx, W, H)
    groups = (M + 3) // 4
    lists = [runtime.CmdList(ctx, M), runtime.CmdList(ctx, M)]
    for l in lists:
        l.upload(cmds)
    pin_cmds = torch.empty(M * 12, dtype=torch.uint8).pin_memory()
    pin_cmds.numpy()[:] = np.frombuffer(cmds.tobytes(), dtype=np.uint8)
    pin_vis = [torch.empty(W * H, dtype=torch.int32).pin_memory() for _ in range(2)]
    L = runtime.lib()
    out = {}
    for name, up, down in (('render only', False, False), ('+ list upload', True, False), ('+ image read-back', False, True), ('e2e step (both)', True, True)):
        k = [0]

        def step():
            i = k[0]
            if up:
                L.chord_cmdlist_upload_async(lists[(i + 1) & 1].h, C.c_void_p(pin_cmds.data_ptr()), C.c_uint32(M))
            frame.begin(view, iview, objects, M, groups)
            frame.render_mesh(lists[i & 1])
            if down:
                L.chord_readback_visibility32_async(frame.h, C.c_void_p(pin_vis[i & 1].data_ptr()))
            k[0] += 1
        for _ in range(4):
            step()
        L.chord_frame_wait_readbacks(frame.h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(30):
            step()
        L.chord_frame_wait_readbacks(frame.h)
        e1.record(stream)
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / 30 * 1000.0, 1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
