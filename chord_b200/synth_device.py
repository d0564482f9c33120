"""Device-side generator for the S5 meshlet stream (same construction as synth.patch_stream_scene, built with torch
directly in HBM so that the 1.5 GB stream never has to be first-touched in host memory).  torch is plumbing here:
it only fills the input buffers; the scene is handed to libchord_b200 through the same chord_scene_desc (the
library copies with cudaMemcpyDefault, so device pointers are accepted)."""
import ctypes as C
import math

import numpy as np
import torch

from . import abi, camera as cam, synth


class DeviceScene:
    """Duck-types abi.HostScene for runtime.Context.upload_scene: ``desc`` points at device arrays."""

    def __init__(self, tensors, prim, materials, n_meshlets):
        self._keep = tensors
        self.primitives = np.ascontiguousarray(prim, dtype=abi.GLTFPrimitiveBuffer)
        self.materials = np.ascontiguousarray(materials, dtype=abi.GLTFMaterialGPUData)
        self._descs = (abi.PrimitiveDatasDesc * 1)()
        d = self._descs[0]
        d.meshlets, d.meshletCount = tensors['meshlets'].data_ptr(), n_meshlets
        d.positions, d.vertexCount = tensors['positions'].data_ptr(), tensors['positions'].numel() // 3
        d.meshletData, d.meshletDataCount = tensors['meshletData'].data_ptr(), tensors['meshletData'].numel()
        d.groups, d.groupCount = tensors['groups'].data_ptr(), tensors['groups'].numel() * 4 // 40
        d.groupIndices, d.groupIndicesCount = tensors['groupIndices'].data_ptr(), tensors['groupIndices'].numel()
        self.desc = abi.SceneDesc()
        self.desc.datas = C.cast(self._descs, C.POINTER(abi.PrimitiveDatasDesc))
        self.desc.datasCount = 1
        self.desc.primitives, self.desc.primitiveCount = self.primitives.ctypes.data, 1
        self.desc.materials, self.desc.materialCount = self.materials.ctypes.data, len(self.materials)
        self.n_meshlets = n_meshlets

    def host_prefix(self, n):
        """abi.HostScene of the first ``n`` meshlets (for the CPU baseline: literally the head of the same stream)."""
        t = self._keep
        V, stride = 81, 209
        meshlets = t['meshlets'][:n * 16].cpu().numpy().view(abi.GPUGLTFMeshlet).copy()
        pos = t['positions'][:n * V * 3].cpu().numpy().reshape(-1, 3).copy()
        data = t['meshletData'][:n * stride].cpu().numpy().view(np.uint32).copy()
        G = (n + 3) // 4
        groups = t['groups'][:G * 10].cpu().numpy().view(abi.GPUGLTFMeshletGroup).copy()
        groups['meshletCount'] = np.minimum(4, n - np.arange(G) * 4)
        prim = self.primitives.copy()
        prim['vertexCount'], prim['meshletGroupCount'] = len(pos), G
        return abi.HostScene([dict(meshlets=meshlets, positions=pos, meshletData=data, groups=groups,
                                   groupIndices=np.arange(n, dtype=np.uint32))], prim, self.materials.copy())


def patch_stream_scene_device(n_meshlets, width=3840, height=2160, seed=5, device='cuda:0', area_range=(0.5, 1.5),
                              depth_range=(1.0, 100.0), overdraw_scale=1.0):
    q, V, T = 8, 81, 128
    M = int(n_meshlets)
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    fovy = math.radians(45.0)
    tan_half = math.tan(fovy * 0.5)
    f32 = torch.float32

    def uni(lo, hi, shape):
        return torch.rand(shape, generator=gen, device=dev, dtype=f32) * (hi - lo) + lo

    depth = uni(depth_range[0], depth_range[1], (M,))
    area = uni(area_range[0], area_range[1], (M,)) * overdraw_scale
    side_px = torch.sqrt(2.0 * area)
    px_world = depth * (2.0 * tan_half / height)
    cx, cy, ang = uni(0, width, (M,)), uni(0, height, (M,)), uni(0, 2 * math.pi, (M,))
    g = torch.arange(q + 1, device=dev, dtype=f32) - q / 2.0
    gy, gx = torch.meshgrid(g, g, indexing='ij')
    gx, gy = gx.reshape(-1), gy.reshape(-1)
    ca, sa = torch.cos(ang)[:, None], torch.sin(ang)[:, None]
    pos = torch.empty((M, V, 3), device=dev, dtype=f32)
    sx = cx[:, None] + (gx[None] * ca - gy[None] * sa) * side_px[:, None]
    sy = cy[:, None] + (gx[None] * sa + gy[None] * ca) * side_px[:, None]
    pos[:, :, 0] = (sx - width / 2.0) * px_world[:, None]
    pos[:, :, 1] = -(sy - height / 2.0) * px_world[:, None]
    del sx, sy
    pos[:, :, 2] = -depth[:, None] + (torch.rand((M, V), generator=gen, device=dev, dtype=f32) - 0.5) * (0.05 * side_px * px_world)[:, None]
    topo = synth._patch_topology(q)
    i0, i1, i2 = topo & 0xFF, (topo >> 8) & 0xFF, (topo >> 16) & 0xFF
    topo = torch.from_numpy((i0 | (i2 << 8) | (i1 << 16)).astype(np.int64)).to(dev).to(torch.int32)
    stride = V + T
    data = torch.empty((M, stride), device=dev, dtype=torch.int32)
    data[:, :V] = (torch.arange(M, device=dev, dtype=torch.int64)[:, None] * V + torch.arange(V, device=dev)[None]).to(torch.int32)
    data[:, V:] = topo[None]
    # meshlet records as 16 x u32 words
    ml = torch.zeros((M, 16), device=dev, dtype=torch.int32)
    mlf = ml.view(f32)
    pmin, pmax, centers = pos.amin(dim=1), pos.amax(dim=1), pos.mean(dim=1)
    mlf[:, 0:3] = pmin
    ml[:, 3] = (torch.arange(M, device=dev, dtype=torch.int64) * stride).to(torch.int32)
    mlf[:, 4:7] = pmax
    ml[:, 7] = int(abi.pack_vertex_triangle_count(V, T))
    mlf[:, 8:11] = torch.tensor([0.0, 0.0, 1.0], device=dev)
    mlf[:, 11] = 0.0
    mlf[:, 12:15] = centers
    G = (M + 3) // 4
    gr = torch.zeros((G, 10), device=dev, dtype=torch.int32)
    grf = gr.view(f32)
    c4 = centers[::4][:G]
    grf[:, 0:3] = c4
    grf[:, 3] = float(np.finfo(np.float32).max)
    grf[:, 4:7] = c4
    grf[:, 7] = -1.0
    gr[:, 8] = (torch.arange(G, device=dev, dtype=torch.int64) * 4).to(torch.int32)
    gr[:, 9] = torch.clamp(M - torch.arange(G, device=dev, dtype=torch.int64) * 4, max=4).to(torch.int32)
    gi = torch.arange(M, device=dev, dtype=torch.int32)
    prim = np.zeros(1, dtype=abi.GLTFPrimitiveBuffer)
    prim['posMin'] = pmin.amin(dim=0).cpu().numpy()
    prim['posMax'] = pmax.amax(dim=0).cpu().numpy()
    prim['posAverage'] = centers.mean(dim=0).cpu().numpy()
    prim['vertexCount'], prim['meshletGroupCount'] = M * V, G
    tensors = dict(meshlets=ml.reshape(-1), positions=pos.reshape(-1), meshletData=data.reshape(-1),
                   groups=gr.reshape(-1), groupIndices=gi)
    torch.cuda.synchronize(dev)
    scene = DeviceScene(tensors, prim, np.array([synth.default_material()]), M)
    camera = cam.Camera((0, 0, 0), (0, 0, -1), fovy=fovy, width=width, height=height)
    objects = np.array([cam.fill_object(np.eye(4), camera, 0, 0)])
    cmds = np.zeros(M, dtype=abi.DrawCmd)
    cmds['meshletId'] = np.arange(M, dtype=np.uint32)
    cmds['slot'] = np.arange(M, dtype=np.uint32)
    return scene, camera, objects, cmds
