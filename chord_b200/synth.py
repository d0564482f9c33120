"""Seeded synthetic scenes in chord's GPU asset layout (S1, S5, S2/S3 stand-ins).

The layout follows the reference's asset packing: per meshlet ``[vertex remap u32 x V][tri (i0|i1<<8|i2<<16) u32 x T]``
appended to ``meshletData`` with ``dataOffset`` in u32 units (source/asset/gltf/asset_gltf_helper.cpp:522-548),
separate tightly packed position / normal / uv / tangent arrays (:572-584), cluster groups of <= 4 meshlets with
(error, parentError) spheres (source/asset/nanite_builder.cpp:851-868; LOD-0 error = -1, root parentError = FLT_MAX,
:420, :891).  Real Sponza/Bistro assets are not available offline, so S2/S3 are procedural "cities".
"""
import math

import numpy as np

from . import abi, camera as cam

F = np.float32
FLT_MAX = np.finfo(np.float32).max


def _patch_topology(q=8):
    """(q+1)^2 vertices, 2*q*q triangles, CCW when seen from +normal (u x v). Returns packed u32 tris."""
    n = q + 1
    tris = []
    for j in range(q):
        for i in range(q):
            a, b, c, d = j * n + i, j * n + i + 1, (j + 1) * n + i, (j + 1) * n + i + 1
            tris.append((a, b, d))
            tris.append((a, d, c))
    t = np.array(tris, dtype=np.uint32)
    return (t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)).astype(np.uint32)


def default_material(two_sided=0, alpha_mode=0, material_type=abi.LIGHTING_TYPE_GLTF_PBR, base_color=(0.8, 0.8, 0.8, 1.0),
                     metallic=0.0, roughness=0.5, alpha_cutoff=0.5):
    """Untextured material as the reference fills it (source/asset/gltf/asset_gltf.cpp:295-358)."""
    m = np.zeros((), dtype=abi.GLTFMaterialGPUData)
    m['alphaMode'], m['alphaCutOff'], m['bTwoSided'] = alpha_mode, alpha_cutoff, two_sided
    m['baseColorId'] = 0                      # builtin white
    m['baseColorFactor'] = base_color
    m['emissiveFactor'] = (0, 0, 0)
    m['emissiveTexture'] = 0                  # builtin transparent
    m['metallicFactor'], m['roughnessFactor'] = metallic, roughness
    m['metallicRoughnessTexture'] = abi.UNVALID_ID
    m['normalTexture'] = abi.UNVALID_ID
    m['normalFactorScale'] = 1.0
    m['occlusionTextureStrength'] = 1.0
    m['materialType'] = material_type
    return m


def build_bvh(groups, group_indices):
    """Cluster-group BVH over the parent-error spheres, restating NaniteBuilder::buildBVHTree
    (source/asset/nanite_builder.cpp:77-416): groups without a parent (parentError == FLT_MAX) are leaves of the
    root; the others are split 2x2x2 by median along the longest axis of their parent-sphere bounds until fewer than
    8 remain (or depth 13); nodes are flattened breadth-first and the GROUP ARRAY IS REORDERED so that every node's
    leaves are contiguous (flattenBVH :215-311).  Returns (groups, group_indices, nodes) in the flattened order."""
    groups = np.asarray(groups, dtype=abi.GPUGLTFMeshletGroup)
    n = len(groups)
    pc = groups['parentPosCenter'].astype(np.float32)
    pe = groups['parentError'].astype(np.float32)
    has_parent = pe < np.float32(3e38)
    ids = np.nonzero(has_parent)[0]

    class Node:
        __slots__ = ('mn', 'mx', 'children', 'leaves', 'depth', 'flat')

        def __init__(self, depth):
            self.mn = np.full(3, FLT_MAX, dtype=np.float32); self.mx = np.full(3, -FLT_MAX, dtype=np.float32)
            self.children, self.leaves, self.depth, self.flat = [], [], depth, -1

    def bounds(sel):
        if len(sel) == 0:
            return np.full(3, FLT_MAX, dtype=np.float32), np.full(3, -FLT_MAX, dtype=np.float32)
        return (pc[sel] - pe[sel, None]).min(axis=0), (pc[sel] + pe[sel, None]).max(axis=0)

    def longest(mn, mx):
        d = mx - mn
        ax = 0
        if d[1] >= d[0] and d[1] >= d[2]:
            ax = 1
        if d[2] >= d[0] and d[2] >= d[1]:
            ax = 2
        return ax

    def halves(sel, mn, mx):
        ax = longest(mn, mx)
        order = sel[np.argsort(pc[sel, ax], kind='stable')]
        m = len(order)
        return [order[i * m // 2:(i + 1) * m // 2] for i in range(2)]

    root = Node(0)
    root.leaves = list(np.nonzero(~has_parent)[0])
    root.mn, root.mx = bounds(ids)
    queue = [(root, ids)]
    while queue:
        node, sel = queue.pop(0)
        if len(sel) == 0:
            continue
        if len(sel) < 8 or node.depth == 13:
            node.leaves += list(sel)
            continue
        for h0 in halves(sel, node.mn, node.mx):
            mn0, mx0 = bounds(h0)
            for h1 in halves(h0, mn0, mx0):
                mn1, mx1 = bounds(h1)
                for h2 in halves(h1, mn1, mx1):
                    c = Node(node.depth + 1)
                    c.mn, c.mx = bounds(h2)
                    node.children.append(c)
                    queue.append((c, h2))
    # flatten breadth-first; leaves of each node become contiguous in the new group array
    order, flat, q = [], [], [root]
    while q:
        nd = q.pop(0)
        nd.flat = len(flat)
        flat.append(nd)
        q += nd.children
    nodes = np.zeros(len(flat), dtype=abi.GPUBVHNode)
    new_groups, new_indices = [], []
    for nd in flat:
        rec = nodes[nd.flat]
        with np.errstate(over='ignore', invalid='ignore'):
            c = np.float32(0.5) * (nd.mx + nd.mn)
        with np.errstate(over='ignore', invalid='ignore'):       # a root without parented groups has infinite bounds
            rec['sphere'] = (c[0], c[1], c[2], np.float32(0.5) * np.float32(np.linalg.norm((nd.mx - nd.mn).astype(np.float32))))
        rec['leafMeshletGroupOffset'], rec['leafMeshletGroupCount'] = len(new_groups), len(nd.leaves)
        for gi in nd.leaves:
            g = groups[gi].copy()
            idx = group_indices[g['meshletOffset']:g['meshletOffset'] + g['meshletCount']]
            g['meshletOffset'] = len(new_indices)
            new_indices += list(idx)
            new_groups.append(g)
        ch = np.full(8, abi.UNVALID_ID, dtype=np.uint32)
        for k, cnode in enumerate(nd.children):
            ch[k] = cnode.flat
        rec['children'] = ch
    for nd in reversed(flat):
        nodes[nd.flat]['bvhNodeCount'] = 1 + sum(int(nodes[c.flat]['bvhNodeCount']) for c in nd.children)
    assert nodes[0]['bvhNodeCount'] == len(flat) and len(new_groups) == n
    return np.array(new_groups, dtype=abi.GPUGLTFMeshletGroup), np.array(new_indices, dtype=np.uint32), nodes


def _meshlet_records(pos_by_meshlet, data_offsets, V, T, lod, normal):
    """Bounds / cone for meshlets given their vertex positions (M, V, 3)."""
    M = pos_by_meshlet.shape[0]
    ms = np.zeros(M, dtype=abi.GPUGLTFMeshlet)
    ms['posMin'] = pos_by_meshlet.min(axis=1)
    ms['posMax'] = pos_by_meshlet.max(axis=1)
    ms['dataOffset'] = data_offsets
    ms['vertexTriangleCount'] = abi.pack_vertex_triangle_count(V, T)
    ms['coneAxis'] = normal
    ms['coneCutOff'] = 0.0
    ms['coneApex'] = pos_by_meshlet.mean(axis=1)
    ms['lod'] = lod
    return ms


def single_meshlet_scene(width=256, height=256, seed=1):
    """S1: one 9x9-vertex grid meshlet (128 tri), plane + seeded +-5% z jitter, identity object matrix,
    camera fovy 45deg looking down -Z so the patch covers ~60% of the viewport."""
    rng = np.random.default_rng(seed)
    q = 8
    dist = 2.0
    half = dist * math.tan(math.radians(22.5)) * 0.78
    g = np.linspace(-half, half, q + 1, dtype=F)
    xx, yy = np.meshgrid(g, g)
    zz = (-dist + (rng.random((q + 1, q + 1)) - 0.5) * 0.1 * half).astype(F)
    pos = np.stack([xx, yy, zz], axis=-1).reshape(-1, 3).astype(F)
    V, T = (q + 1) ** 2, 2 * q * q
    data = np.concatenate([np.arange(V, dtype=np.uint32), _patch_topology(q)])
    meshlets = _meshlet_records(pos[None], np.array([0], dtype=np.uint32), V, T, 0, (0, 0, 1))
    groups = np.zeros(1, dtype=abi.GPUGLTFMeshletGroup)
    groups['clusterPosCenter'] = pos.mean(axis=0)
    groups['parentPosCenter'] = pos.mean(axis=0)
    groups['error'], groups['parentError'] = -1.0, FLT_MAX
    groups['meshletOffset'], groups['meshletCount'] = 0, 1
    normals = np.tile(np.array([0, 0, 1], dtype=F), (V, 1))
    uv = np.stack([(xx / (2 * half) + 0.5), (yy / (2 * half) + 0.5)], axis=-1).reshape(-1, 2).astype(F)
    tang = np.tile(np.array([1, 0, 0, 1], dtype=F), (V, 1))
    prim = np.zeros(1, dtype=abi.GLTFPrimitiveBuffer)
    prim['posMin'], prim['posMax'], prim['posAverage'] = pos.min(axis=0), pos.max(axis=0), pos.mean(axis=0)
    prim['vertexCount'], prim['meshletGroupCount'] = V, 1
    groups, gidx, nodes = build_bvh(groups, np.zeros(1, dtype=np.uint32))
    scene = abi.HostScene([dict(meshlets=meshlets, positions=pos, normals=normals, uv0=uv, tangents=tang,
                                meshletData=data, groups=groups, groupIndices=gidx, bvhNodes=nodes)],
                          prim, np.array([default_material()]))
    camera = cam.Camera((0, 0, 0), (0, 0, -1), width=width, height=height)
    objects = np.array([cam.fill_object(np.eye(4), camera, 0, 0)])
    return scene, camera, objects


def patch_stream_scene(n_meshlets, width=3840, height=2160, seed=5, area_range=(0.5, 1.5), depth_range=(1.0, 100.0),
                       overdraw_scale=1.0, chunk=16384):
    """S5: ``n_meshlets`` camera-facing 8x8-quad patches (V=81, T=128) whose projected triangle area is uniform in
    ``area_range`` px^2, patch centres uniform over the viewport, depth uniform in ``depth_range`` metres; one
    opaque one-sided object, identity matrix, camera at the origin looking down -Z.  Generated in chunks into
    preallocated arrays (the full stream is 1.5 GB)."""
    q, V, T = 8, 81, 128
    fovy = math.radians(45.0)
    tan_half = math.tan(fovy * 0.5)
    M = int(n_meshlets)
    stride = V + T
    pos = np.empty((M, V, 3), dtype=F)
    data = np.empty((M, stride), dtype=np.uint32)
    meshlets = np.zeros(M, dtype=abi.GPUGLTFMeshlet)
    centers = np.empty((M, 3), dtype=F)
    g = (np.arange(q + 1) - q / 2.0)
    gx, gy = np.meshgrid(g, g)
    gx, gy = gx.reshape(-1).astype(F), gy.reshape(-1).astype(F)
    # the patch is laid out with screen-y down; flip winding so the camera-facing side is CCW in y-up NDC
    topo = _patch_topology(q)
    i0, i1, i2 = topo & 0xFF, (topo >> 8) & 0xFF, (topo >> 16) & 0xFF
    topo = (i0 | (i2 << 8) | (i1 << 16)).astype(np.uint32)
    lane = np.arange(V, dtype=np.uint32)[None]
    for c0 in range(0, M, chunk):
        c1 = min(M, c0 + chunk)
        n = c1 - c0
        rng = np.random.default_rng([seed, c0 // chunk])
        depth = rng.uniform(depth_range[0], depth_range[1], n).astype(F)
        area = (rng.uniform(area_range[0], area_range[1], n) * overdraw_scale).astype(F)
        side_px = np.sqrt(F(2.0) * area)                    # quad side in pixels
        px_world = (F(2.0 * tan_half / height) * depth)     # world units per pixel at that depth
        cx = rng.uniform(0, width, n).astype(F)
        cy = rng.uniform(0, height, n).astype(F)
        ang = rng.uniform(0, 2 * math.pi, n).astype(F)
        ca, sa = np.cos(ang)[:, None], np.sin(ang)[:, None]
        sx = cx[:, None] + (gx[None] * ca - gy[None] * sa) * side_px[:, None]
        sy = cy[:, None] + (gx[None] * sa + gy[None] * ca) * side_px[:, None]
        p = pos[c0:c1]
        p[:, :, 0] = (sx - F(width / 2.0)) * px_world[:, None]
        p[:, :, 1] = -(sy - F(height / 2.0)) * px_world[:, None]
        p[:, :, 2] = -depth[:, None] + (rng.random((n, V), dtype=F) - F(0.5)) * (F(0.05) * side_px * px_world)[:, None]
        data[c0:c1, :V] = (np.arange(c0, c1, dtype=np.uint32)[:, None] * np.uint32(V) + lane)
        data[c0:c1, V:] = topo[None]
        m = meshlets[c0:c1]
        m['posMin'] = p.min(axis=1)
        m['posMax'] = p.max(axis=1)
        centers[c0:c1] = p.mean(axis=1)
    meshlets['dataOffset'] = (np.arange(M, dtype=np.uint64) * stride).astype(np.uint32)
    meshlets['vertexTriangleCount'] = abi.pack_vertex_triangle_count(V, T)
    meshlets['coneAxis'] = (0, 0, 1)
    meshlets['coneCutOff'] = 0.0
    meshlets['coneApex'] = centers
    G = (M + 3) // 4
    groups = np.zeros(G, dtype=abi.GPUGLTFMeshletGroup)
    groups['error'], groups['parentError'] = -1.0, FLT_MAX
    groups['meshletOffset'] = np.arange(G, dtype=np.uint32) * 4
    groups['meshletCount'] = np.minimum(4, M - np.arange(G) * 4)
    groups['clusterPosCenter'] = centers[::4][:G]
    groups['parentPosCenter'] = centers[::4][:G]
    flat = pos.reshape(-1, 3)
    prim = np.zeros(1, dtype=abi.GLTFPrimitiveBuffer)
    prim['posMin'] = meshlets['posMin'].min(axis=0)
    prim['posMax'] = meshlets['posMax'].max(axis=0)
    prim['posAverage'] = centers.mean(axis=0)
    prim['vertexCount'], prim['meshletGroupCount'] = len(flat), G
    nodes = np.zeros(1, dtype=abi.GPUBVHNode)          # every group is a root-level leaf (no parent): one node
    nodes['children'] = abi.UNVALID_ID
    nodes['bvhNodeCount'], nodes['leafMeshletGroupOffset'], nodes['leafMeshletGroupCount'] = 1, 0, G
    scene = abi.HostScene([dict(meshlets=meshlets, positions=flat, meshletData=data.reshape(-1), groups=groups,
                                groupIndices=np.arange(M, dtype=np.uint32), bvhNodes=nodes)],
                          prim, np.array([default_material()]))
    camera = cam.Camera((0, 0, 0), (0, 0, -1), fovy=fovy, width=width, height=height)
    objects = np.array([cam.fill_object(np.eye(4), camera, 0, 0)])
    cmds = np.zeros(M, dtype=abi.DrawCmd)
    cmds['meshletId'] = np.arange(M, dtype=np.uint32)
    cmds['slot'] = np.arange(M, dtype=np.uint32)
    return scene, camera, objects, cmds


def ragged_stream_scene(n_meshlets, width=640, height=360, seed=9, two_sided_every=5):
    """Edge-case stream for the raster: meshlets of every shape the format allows (``base.h:406-445``): T from 1 to 128
    triangles, V from 3 up to the maximum of 255 vertices (triangles partly or fully un-welded, so that V > 128 exercises
    the large-vertex path of the fast raster kernel), odd data offsets (the TMA stage reads 16-byte supersets), patch
    sizes from sub-pixel to tens of pixels.  Returns (scene, camera, objects, cmds) like ``patch_stream_scene``."""
    rng = np.random.default_rng(seed)
    q = 8
    fovy = math.radians(45.0)
    tan_half = math.tan(fovy * 0.5)
    topo = _patch_topology(q)
    i0, i1, i2 = topo & 0xFF, (topo >> 8) & 0xFF, (topo >> 16) & 0xFF
    corners = np.stack([i0, i2, i1], axis=-1).astype(np.int64)          # camera-facing winding (see patch_stream_scene)
    g = (np.arange(q + 1) - q / 2.0)
    gx, gy = np.meshgrid(g, g)
    gx, gy = gx.reshape(-1).astype(F), gy.reshape(-1).astype(F)
    shapes = [(1, 0.0), (1, 1.0), (2, 1.0), (85, 1.0), (128, 0.0), (128, 0.55), (127, 0.45), (43, 1.0), (96, 0.7), (128, 0.3)]
    pos_all, data_all, meshlets, centers = [], [], [], []
    data_len, vert_base = 0, 0
    for mi in range(int(n_meshlets)):
        T, unweld = shapes[mi % len(shapes)] if mi < 4 * len(shapes) else (int(rng.integers(1, 129)), float(rng.random()))
        tri = corners[rng.permutation(128)[:T]] if mi % 3 == 0 else corners[:T]
        depth = F(rng.uniform(1.0, 60.0))
        side_px = F(np.sqrt(2.0 * rng.choice([0.3, 1.0, 1.0, 4.0, 30.0, 400.0])))
        px_world = F(2.0 * tan_half / height) * depth
        cx, cy, ang = F(rng.uniform(-20, width + 20)), F(rng.uniform(-20, height + 20)), F(rng.uniform(0, 2 * math.pi))
        sx = cx + (gx * np.cos(ang) - gy * np.sin(ang)) * side_px
        sy = cy + (gx * np.sin(ang) + gy * np.cos(ang)) * side_px
        grid = np.stack([(sx - F(width / 2.0)) * px_world, -(sy - F(height / 2.0)) * px_world,
                         -depth + (rng.random(81).astype(F) - F(0.5)) * F(0.05) * side_px * px_world], axis=-1).astype(F)
        # local vertex list: welded grid vertices first come first served, then duplicates for un-welded corners
        local, verts, idx = {}, [], np.zeros((T, 3), dtype=np.uint32)
        for t in range(T):
            for c in range(3):
                gv = int(tri[t, c])
                fresh = rng.random() < unweld and len(verts) + 3 * (T - t) - c <= 255     # keep V <= 255 whatever comes
                if not fresh and gv in local:
                    idx[t, c] = local[gv]
                    continue
                if len(verts) >= 255:
                    idx[t, c] = local.get(gv, 0)
                    continue
                idx[t, c] = len(verts)
                if not fresh:
                    local[gv] = len(verts)
                verts.append(grid[gv])
        V = len(verts)
        verts = np.asarray(verts, dtype=F)
        if mi % 7 == 3:                                                  # odd data offset
            data_all.append(np.full(int(rng.integers(1, 4)), 0xDEADBEEF, dtype=np.uint32)); data_len += len(data_all[-1])
        rec = np.zeros(1, dtype=abi.GPUGLTFMeshlet)
        rec['posMin'], rec['posMax'] = verts.min(axis=0), verts.max(axis=0)
        rec['dataOffset'] = data_len
        rec['vertexTriangleCount'] = abi.pack_vertex_triangle_count(V, T)
        rec['coneAxis'], rec['coneCutOff'], rec['coneApex'] = (0, 0, 1), 0.0, verts.mean(axis=0)
        meshlets.append(rec)
        centers.append(verts.mean(axis=0))
        data_all.append(np.arange(vert_base, vert_base + V, dtype=np.uint32))
        data_all.append((idx[:, 0] | (idx[:, 1] << 8) | (idx[:, 2] << 16)).astype(np.uint32))
        data_len += V + T
        vert_base += V
        pos_all.append(verts)
    M = int(n_meshlets)
    meshlets = np.concatenate(meshlets)
    centers = np.asarray(centers, dtype=F)
    flat = np.concatenate(pos_all)
    G = (M + 3) // 4
    groups = np.zeros(G, dtype=abi.GPUGLTFMeshletGroup)
    groups['error'], groups['parentError'] = -1.0, FLT_MAX
    groups['meshletOffset'] = np.arange(G, dtype=np.uint32) * 4
    groups['meshletCount'] = np.minimum(4, M - np.arange(G) * 4)
    groups['clusterPosCenter'] = centers[::4][:G]
    groups['parentPosCenter'] = centers[::4][:G]
    prims = np.zeros(2, dtype=abi.GLTFPrimitiveBuffer)
    prims['posMin'], prims['posMax'], prims['posAverage'] = flat.min(axis=0), flat.max(axis=0), flat.mean(axis=0)
    prims['vertexCount'], prims['meshletGroupCount'] = len(flat), G
    nodes = np.zeros(1, dtype=abi.GPUBVHNode)
    nodes['children'] = abi.UNVALID_ID
    nodes['bvhNodeCount'], nodes['leafMeshletGroupOffset'], nodes['leafMeshletGroupCount'] = 1, 0, G
    scene = abi.HostScene([dict(meshlets=meshlets, positions=flat, meshletData=np.concatenate(data_all), groups=groups,
                                groupIndices=np.arange(M, dtype=np.uint32), bvhNodes=nodes)],
                          prims, np.array([default_material(), default_material(two_sided=1)]))
    camera = cam.Camera((0, 0, 0), (0, 0, -1), fovy=fovy, width=width, height=height)
    # object 0 one-sided, object 1 two-sided (same asset): every ``two_sided_every``-th command draws through object 1
    objects = np.array([cam.fill_object(np.eye(4), camera, 0, 0), cam.fill_object(np.eye(4), camera, 1, 1)])
    cmds = np.zeros(M, dtype=abi.DrawCmd)
    cmds['meshletId'] = np.arange(M, dtype=np.uint32)
    cmds['slot'] = np.arange(M, dtype=np.uint32)
    cmds['objectId'] = (np.arange(M) % two_sided_every == two_sided_every - 1).astype(np.uint32)
    return scene, camera, objects, cmds


# ----------------------------------------------------------------------------------------------- city (S2/S3)
def _face_lod_chain(origin, du, dv, patches, normal, vert_base, data_base, meshlet_base, lods, err_scale=0.05):
    """One planar face of ``patches x patches`` LOD-0 patches (8x8 quads each) with a quadtree LOD chain.

    Returns dict of arrays to append: positions/normals/uv/tangents, meshletData, meshlets, and a list of group
    tuples (lod, meshlet indices (local to the primitive), centre, radius-ish error, parent info)."""
    q = 8
    out_pos, out_nrm, out_uv, out_tan, out_data, out_meshlets = [], [], [], [], [], []
    groups = []     # (meshlet_ids, own_center, own_error, parent_center, parent_error)
    topo = _patch_topology(q)
    origin, du, dv, normal = (np.asarray(a, dtype=np.float64) for a in (origin, du, dv, normal))
    tangent = du / np.linalg.norm(du)
    level_ids = []
    vcount, dcount, mcount = 0, 0, 0
    p = patches
    lod = 0
    face_len = max(np.linalg.norm(du), np.linalg.norm(dv))
    while True:
        n = p * q + 1
        s = np.linspace(0.0, 1.0, n)
        uu, vv = np.meshgrid(s, s)
        pos = (origin[None, None] + uu[..., None] * du[None, None] + vv[..., None] * dv[None, None]).reshape(-1, 3)
        out_pos.append(pos.astype(F))
        out_nrm.append(np.tile(normal.astype(F), (n * n, 1)))
        out_uv.append(np.stack([uu, vv], axis=-1).reshape(-1, 2).astype(F))
        out_tan.append(np.tile(np.append(tangent, 1.0).astype(F), (n * n, 1)))
        ids = np.zeros((p, p), dtype=np.int64)
        for pj in range(p):
            for pi in range(p):
                jj, ii = np.meshgrid(np.arange(q + 1) + pj * q, np.arange(q + 1) + pi * q, indexing='ij')
                remap = (vert_base + vcount + jj * n + ii).reshape(-1).astype(np.uint32)
                mpos = pos[(jj * n + ii).reshape(-1)]
                rec = _meshlet_records(mpos[None].astype(F), np.array([data_base + dcount], dtype=np.uint32),
                                       (q + 1) ** 2, 2 * q * q, lod, normal.astype(F))
                out_data.append(np.concatenate([remap, topo]))
                out_meshlets.append(rec)
                ids[pj, pi] = meshlet_base + mcount
                dcount += len(remap) + len(topo)
                mcount += 1
        level_ids.append(ids)
        vcount += n * n
        lod += 1
        if p == 1 or lod >= lods:
            break
        p //= 2
    # DAG (nanite_builder.cpp:851-868 semantics on a quadtree): the level-k meshlet at (j, i) was produced by
    # simplifying the 2x2 block of level-(k-1) meshlets under it, so its own sphere is Sph(k, j, i) and the parent
    # sphere of those four children is the same Sph(k, j, i).  A GPU cluster group collects meshlets sharing both
    # spheres: 2x2 blocks at level 0, single meshlets above.
    nl = len(level_ids)

    def sph(level, j, i):
        pl = level_ids[level].shape[0]
        c = origin + ((i + 0.5) / pl) * du + ((j + 0.5) / pl) * dv
        err = err_scale * face_len / (level_ids[level - 1].shape[0] * q)     # nominal error of removing level k-1 detail
        return c, err

    for level in range(nl):
        ids = level_ids[level]
        pl = ids.shape[0]
        if level == 0 and nl > 1:
            for bj in range(pl // 2):
                for bi in range(pl // 2):
                    mids = ids[bj * 2:bj * 2 + 2, bi * 2:bi * 2 + 2].reshape(-1)
                    pc, pe = sph(1, bj, bi)
                    groups.append((mids.astype(np.uint32), pc, -1.0, pc, pe))
        elif level == 0:
            for bj in range(0, pl, 2):
                for bi in range(0, pl, 2):
                    mids = ids[bj:bj + 2, bi:bi + 2].reshape(-1)
                    c = origin + 0.5 * du + 0.5 * dv
                    groups.append((mids.astype(np.uint32), c, -1.0, c, FLT_MAX))
        else:
            for j in range(pl):
                for i in range(pl):
                    oc, oe = sph(level, j, i)
                    if level == nl - 1:
                        pc, pe = oc, FLT_MAX
                    else:
                        pc, pe = sph(level + 1, j // 2, i // 2)
                    groups.append((np.array([ids[j, i]], dtype=np.uint32), oc, oe, pc, pe))
    return dict(positions=np.concatenate(out_pos), normals=np.concatenate(out_nrm), uv0=np.concatenate(out_uv),
                tangents=np.concatenate(out_tan), meshletData=np.concatenate(out_data),
                meshlets=np.concatenate(out_meshlets), groups=groups, lod0=int(level_ids[0].size))


def _box_primitive(size, patches, lods, five_faces=True):
    """Axis-aligned box [0,sx]x[0,sy]x[0,sz] in local space; faces outward, optional no bottom."""
    sx, sy, sz = size
    faces = [
        ((0, 0, sz), (sx, 0, 0), (0, sy, 0), (0, 0, 1)),     # +z
        ((sx, 0, 0), (-sx, 0, 0), (0, sy, 0), (0, 0, -1)),   # -z
        ((sx, 0, sz), (0, 0, -sz), (0, sy, 0), (1, 0, 0)),   # +x
        ((0, 0, 0), (0, 0, sz), (0, sy, 0), (-1, 0, 0)),     # -x
        ((0, sy, sz), (sx, 0, 0), (0, 0, -sz), (0, 1, 0)),   # +y (roof)
    ]
    if not five_faces:
        faces.append(((0, 0, 0), (sx, 0, 0), (0, 0, sz), (0, -1, 0)))
    return faces


class _PrimBuilder:
    def __init__(self):
        self.pos, self.nrm, self.uv, self.tan, self.data, self.meshlets = [], [], [], [], [], []
        self.groups, self.group_indices = [], []
        self.prims = []
        self.nv = self.nd = self.nm = self.ng = self.ngi = self.nn = 0
        self.nodes = []
        self.lod0 = []

    def add_primitive(self, faces, patches, lods, err_scale=0.05):
        v0, m0, g0, gi0 = self.nv, self.nm, self.ng, self.ngi
        lod0 = 0
        allpos = []
        pgroups, pindices = [], []
        for (o, du, dv, nrm) in faces:
            f = _face_lod_chain(o, du, dv, patches, nrm, self.nv - v0, self.nd, self.nm - m0, lods, err_scale)
            self.pos.append(f['positions']); self.nrm.append(f['normals']); self.uv.append(f['uv0'])
            self.tan.append(f['tangents']); self.data.append(f['meshletData']); self.meshlets.append(f['meshlets'])
            allpos.append(f['positions'])
            self.nv += len(f['positions']); self.nd += len(f['meshletData']); self.nm += len(f['meshlets'])
            lod0 += f['lod0']
            for (mids, oc, oe, pc, pe) in f['groups']:
                g = np.zeros((), dtype=abi.GPUGLTFMeshletGroup)
                g['clusterPosCenter'], g['error'] = oc, oe
                g['parentPosCenter'], g['parentError'] = pc, pe
                g['meshletOffset'], g['meshletCount'] = sum(len(x) for x in pindices), len(mids)
                pgroups.append(g)
                pindices.append(mids)
        # cluster-group BVH of this primitive; the builder reorders the groups (nanite_builder.cpp:215-311)
        bg, bi, bn = build_bvh(np.array(pgroups), np.concatenate(pindices))
        self.groups += list(bg)
        self.group_indices.append(bi)
        self.nodes.append(bn)
        self.ng += len(bg); self.ngi += len(bi)
        ap = np.concatenate(allpos)
        p = np.zeros((), dtype=abi.GLTFPrimitiveBuffer)
        p['posMin'], p['posMax'], p['posAverage'] = ap.min(axis=0), ap.max(axis=0), ap.mean(axis=0)
        p['vertexOffset'], p['vertexCount'] = v0, self.nv - v0
        p['meshletOffset'] = m0
        p['meshletGroupOffset'], p['meshletGroupIndicesOffset'], p['meshletGroupCount'] = g0, gi0, self.ng - g0
        p['bvhNodeOffset'] = self.nn
        self.nn += len(bn)
        self.prims.append(p)
        self.lod0.append(lod0)
        return len(self.prims) - 1

    def finish(self, materials):
        datas = dict(meshlets=np.concatenate(self.meshlets), positions=np.concatenate(self.pos),
                     normals=np.concatenate(self.nrm), uv0=np.concatenate(self.uv), tangents=np.concatenate(self.tan),
                     meshletData=np.concatenate(self.data), groups=np.array(self.groups),
                     groupIndices=np.concatenate(self.group_indices), bvhNodes=np.concatenate(self.nodes))
        return abi.HostScene([datas], np.array(self.prims), np.array(materials))


def city_scene(n_buildings=26, patches=4, lods=3, width=1920, height=1080, seed=2, ground_patches=8, n_protos=3, err_scale=0.05):
    """S2/S3 stand-in: a grid of box buildings (instanced from 3 prototypes with a quadtree LOD DAG per face) on a
    coarse ground slab, camera at street level so that near-plane clipping, large triangles, LOD cuts and occlusion
    all occur.  ``n_buildings=26, patches=4`` ~ 266 k LOD-0 triangles (S2); ``n_buildings=371`` ~ 3.8 M (S3)."""
    rng = np.random.default_rng(seed)
    b = _PrimBuilder()
    sizes = [(10.0, 30.0, 10.0), (14.0, 18.0, 12.0), (8.0, 45.0, 8.0)]
    if n_protos > 3:
        prng = np.random.default_rng(seed + 1000)
        sizes += [(float(prng.uniform(7, 15)), float(prng.uniform(12, 50)), float(prng.uniform(7, 15))) for _ in range(n_protos - 3)]
    protos = [b.add_primitive(_box_primitive(sz, patches, lods), patches, lods, err_scale) for sz in sizes[:max(n_protos, 1)]]
    side = int(math.ceil(math.sqrt(n_buildings)))
    pitch = 40.0
    extent = side * pitch
    ground = b.add_primitive([((0, 0, extent), (extent, 0, 0), (0, 0, -extent), (0, 1, 0))], ground_patches, 1)
    mats = [default_material(base_color=(0.7, 0.7, 0.72, 1.0)),
            default_material(two_sided=1, base_color=(0.5, 0.6, 0.5, 1.0)),
            default_material(alpha_mode=1, alpha_cutoff=0.5, base_color=(0.6, 0.4, 0.3, 1.0)),
            default_material(base_color=(0.3, 0.3, 0.3, 1.0), roughness=0.9)]
    scene = b.finish(mats)
    camera = cam.Camera((pitch - 7.0, 7.0, extent + 12.0), (0.06, -0.12, -1.0), width=width, height=height)
    objs = []
    transforms = []
    k = 0
    for j in range(side):
        for i in range(side):
            if k >= n_buildings:
                break
            proto = protos[int(rng.integers(0, len(protos)))]
            m = np.eye(4)
            yaw = rng.uniform(-0.3, 0.3)
            s = rng.uniform(0.8, 1.25)
            c, sn = math.cos(yaw), math.sin(yaw)
            m[:3, :3] = np.array([[c, 0, sn], [0, 1, 0], [-sn, 0, c]]) * s
            m[:3, 3] = (i * pitch + rng.uniform(2, 8), 0.0, j * pitch + rng.uniform(2, 8))
            transforms.append((m, proto, int(rng.integers(0, 3))))
            k += 1
    transforms.append((np.eye(4), ground, 3))
    for (m, proto, mat) in transforms:
        objs.append(cam.fill_object(m, camera, proto, mat))
    objects = np.array(objs)
    lod0_total = int(sum(b.lod0[p] for (_, p, _) in transforms))
    group_total = int(sum(int(scene.primitives[p]['meshletGroupCount']) for (_, p, _) in transforms))
    return scene, camera, objects, dict(lod0MeshletCount=lod0_total, meshletGroupCount=group_total,
                                        transforms=transforms)


def refill_objects(transforms, camera, camera_last=None):
    """Per-frame GPUObjectGLTFPrimitive array for a moved camera (static geometry)."""
    return np.array([cam.fill_object(m, camera, proto, mat, camera_last=camera_last) for (m, proto, mat) in transforms])


def synthetic_luts(seed=7):
    """Seeded stand-ins for the Bruneton transmittance (256x64) / irradiance (64x16) RGBA16F LUTs."""
    rng = np.random.default_rng(seed)
    u = np.linspace(0, 1, 256)[None, :, None]
    v = np.linspace(0, 1, 64)[:, None, None]
    tr = np.clip(0.2 + 0.8 * u * (0.5 + 0.5 * v) + 0.02 * rng.random((64, 256, 1)), 0, 1) * np.array([1.0, 0.95, 0.85])
    tr = np.concatenate([tr, np.ones((64, 256, 1))], axis=-1).astype(np.float16)
    u = np.linspace(0, 1, 64)[None, :, None]
    v = np.linspace(0, 1, 16)[:, None, None]
    ir = (0.02 + 0.1 * u * (1.0 - 0.5 * v)) * np.array([0.6, 0.8, 1.0])
    ir = np.concatenate([ir, np.ones((16, 64, 1))], axis=-1).astype(np.float16)
    return np.ascontiguousarray(tr), np.ascontiguousarray(ir)


def synthetic_scattering_luts(seed=8):
    """Seeded stand-ins for the Bruneton scattering (combined: rgb = Rayleigh + multiple scattering, a = single Mie red) and
    single-Mie 4-D LUTs packed as 256 x 128 x 32 RGBA16F volumes (atmosphere.hlsli:43-49): smooth, positive, with structure in
    every texture dimension so that a wrong coordinate shows."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0, 1, 256)[None, None, :, None]
    y = np.linspace(0, 1, 128)[None, :, None, None]
    z = np.linspace(0, 1, 32)[:, None, None, None]
    base = 0.02 + 0.3 * (0.5 + 0.5 * np.sin(6.0 * x + 2.0 * y)) * (0.3 + 0.7 * y) * (1.0 - 0.6 * z) + 0.01 * rng.random((32, 128, 256, 1))
    sc = np.concatenate([base * np.array([0.35, 0.6, 1.0]), 0.05 + 0.2 * (1.0 - y) * (0.5 + 0.5 * np.cos(9.0 * x)) * np.ones((32, 1, 1, 1))], axis=-1)
    mie = np.concatenate([(0.03 + 0.15 * (1.0 - y) * (0.5 + 0.5 * np.cos(5.0 * x + z))) * np.array([1.0, 0.95, 0.9]) * np.ones((32, 1, 1, 1)), np.ones((32, 128, 256, 1))], axis=-1)
    return np.ascontiguousarray(sc.astype(np.float16)), np.ascontiguousarray(mie.astype(np.float16))


# ----------------------------------------------------------------------------------------------- masked bucket (N2)
def alpha_test_textures(seed=7):
    """Procedural RGBA8 alpha-test textures with full box-filtered mip chains: [None (builtin white), 64x64 leaf blobs,
    128x32 chain-link stripes (non-square), 16x16 checker]."""
    rng = np.random.default_rng(seed)

    def finish(alpha):
        h, w = alpha.shape
        img = np.empty((h, w, 4), dtype=np.uint8)
        img[..., :3] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        img[..., 3] = alpha
        return abi.box_mip_chain(img)

    yy, xx = np.mgrid[0:64, 0:64]
    blobs = np.zeros((64, 64), dtype=np.float64)
    for _ in range(14):
        cx, cy, r = rng.uniform(0, 64), rng.uniform(0, 64), rng.uniform(5, 12)
        d = np.hypot((xx - cx + 32) % 64 - 32, (yy - cy + 32) % 64 - 32)
        blobs = np.maximum(blobs, np.clip(1.5 - d / r, 0.0, 1.0))
    leaf = (blobs * 255.0 + 0.5).astype(np.uint8)
    yy, xx = np.mgrid[0:32, 0:128]
    link = (((xx + 2 * yy) % 16 < 5) | ((xx - 2 * yy) % 16 < 5)).astype(np.uint8) * np.uint8(255)
    link = np.where(link > 0, link, rng.integers(0, 90, link.shape).astype(np.uint8)).astype(np.uint8)
    yy, xx = np.mgrid[0:16, 0:16]
    checker = (((xx // 2 + yy // 2) % 2) * 200 + 30).astype(np.uint8)
    return [None, finish(leaf), finish(link), finish(checker)]


def masked_scene(width=512, height=384, seed=11, n_random=24):
    """N2 test scene: instances of one 8x8-quad patch (uv 0..1) with alpha-tested textured materials at varied distance
    (all mip levels incl. the 1x1 tail), tilt (perspective-correct uv), uv scale via per-material sampler address modes
    (repeat / clamp / mirror), sidedness, cut-offs; one instance crosses the near plane (clip path with uv), one fills
    the screen (CTA-cooperative path); untextured masked, blend and opaque materials share the frame."""
    rng = np.random.default_rng(seed)
    q = 8
    g = np.linspace(-1.0, 1.0, q + 1, dtype=F)
    xx, yy = np.meshgrid(g, g)
    V, T = (q + 1) ** 2, 2 * q * q

    def patch(uv_scale, uv_offset):
        pos = np.stack([xx, yy, np.zeros_like(xx)], axis=-1).reshape(-1, 3).astype(F)
        uv = (np.stack([xx * 0.5 + 0.5, 0.5 - yy * 0.5], axis=-1).reshape(-1, 2) * uv_scale + uv_offset).astype(F)
        return pos, uv

    # two primitives in one asset: uv in [0, 1] and uv in [-1.3, 2.9] (exercises the address modes)
    p0, uv0 = patch(1.0, 0.0)
    p1, uv1 = patch(4.2, -1.3)
    pos = np.concatenate([p0, p1])
    uv = np.concatenate([uv0, uv1])
    topo = _patch_topology(q)
    data = np.concatenate([np.arange(V, dtype=np.uint32), topo, np.arange(V, dtype=np.uint32), topo])
    meshlets = np.concatenate([_meshlet_records(p0[None], np.array([0], dtype=np.uint32), V, T, 0, (0, 0, 1)),
                               _meshlet_records(p1[None], np.array([V + T], dtype=np.uint32), V, T, 0, (0, 0, 1))])
    groups = np.zeros(2, dtype=abi.GPUGLTFMeshletGroup)
    groups['clusterPosCenter'] = 0.0
    groups['parentPosCenter'] = 0.0
    groups['error'], groups['parentError'] = -1.0, FLT_MAX
    groups['meshletOffset'], groups['meshletCount'] = 0, 1            # index into the primitive's own groupIndices range
    normals = np.tile(np.array([0, 0, 1], dtype=F), (2 * V, 1))
    tang = np.tile(np.array([1, 0, 0, 1], dtype=F), (2 * V, 1))
    prim = np.zeros(2, dtype=abi.GLTFPrimitiveBuffer)
    for k in range(2):
        prim[k]['posMin'], prim[k]['posMax'], prim[k]['posAverage'] = (-1, -1, 0), (1, 1, 0), (0, 0, 0)
        prim[k]['vertexOffset'], prim[k]['vertexCount'] = k * V, V
        prim[k]['meshletGroupOffset'], prim[k]['meshletGroupIndicesOffset'], prim[k]['meshletGroupCount'] = k, k, 1
        prim[k]['meshletOffset'], prim[k]['bvhNodeOffset'] = k, k
    mats = []
    for tex, smp, two, cut, fac in ((1, 0, 0, 0.5, 1.0), (1, 1, 1, 0.35, 1.0), (2, 2, 1, 0.5, 0.9), (3, 0, 0, 0.45, 1.0),
                                    (2, 1, 0, 0.6, 1.0), (3, 2, 1, 0.2, 0.7)):
        m = default_material(two_sided=two, alpha_mode=1, alpha_cutoff=cut, base_color=(0.8, 0.8, 0.8, fac))
        m['baseColorId'], m['baseColorSampler'] = tex, smp
        mats.append(m)
    mats.append(default_material(alpha_mode=1, alpha_cutoff=0.5))                                   # masked, builtin white: kept
    mats.append(default_material(alpha_mode=1, alpha_cutoff=0.5, base_color=(1, 1, 1, 0.4)))        # masked, builtin white: dropped
    mats.append(default_material(alpha_mode=2))                                                     # blend: never drawn
    mats.append(default_material())                                                                 # opaque
    mats = np.array(mats)
    table = abi.TextureTable(alpha_test_textures(), samplers=((abi.ADDRESS_REPEAT, abi.ADDRESS_REPEAT),
                                                              (abi.ADDRESS_CLAMP_TO_EDGE, abi.ADDRESS_MIRRORED_REPEAT),
                                                              (abi.ADDRESS_MIRRORED_REPEAT, abi.ADDRESS_CLAMP_TO_EDGE)))
    parts = [build_bvh(groups[k:k + 1], np.zeros(1, dtype=np.uint32)) for k in range(2)]
    groups = np.concatenate([p_[0] for p_ in parts]); gidx = np.concatenate([p_[1] for p_ in parts]); nodes = np.concatenate([p_[2] for p_ in parts])
    scene = abi.HostScene([dict(meshlets=meshlets, positions=pos, normals=normals, uv0=uv, tangents=tang,
                                meshletData=data, groups=groups, groupIndices=gidx, bvhNodes=nodes)], prim, mats, textures=table)
    camera = cam.Camera((0, 0, 0), (0, 0, -1), width=width, height=height)

    def xform(center, scale, yaw, pitch, roll=0.0):
        cy_, sy_, cp, sp, cr, sr = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
        ry = np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
        rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
        m = np.eye(4)
        m[:3, :3] = (ry @ rx @ rz) * scale
        m[:3, 3] = center
        return m

    objs = []
    # screen-filling masked quad far back (CTA-cooperative path, low mips), a ground plane through the near plane (clip path)
    objs.append(cam.fill_object(xform((0, 0, -30.0), 22.0, 0.1, -0.05), camera, 1, 0))
    objs.append(cam.fill_object(xform((0, -0.6, -1.0), 30.0, 0.3, -1.45), camera, 1, 1))
    objs.append(cam.fill_object(xform((0.5, 0.2, -20.0), 0.5, 0.4, 0.3), camera, 1, 2))         # ~23 texels / px: level 4..5
    objs.append(cam.fill_object(xform((-2.0, 1.0, -20.0), 0.13, 0.2, -0.3), camera, 1, 3))      # a few pixels: the 1x1 mip tail
    for i in range(n_random):
        z = -float(rng.uniform(1.5, 60.0))
        ext = -z * 0.45
        c = (float(rng.uniform(-ext, ext)), float(rng.uniform(-ext * 0.7, ext * 0.7)), z)
        objs.append(cam.fill_object(xform(c, float(rng.uniform(0.3, 2.5)) * (-z) * 0.12, float(rng.uniform(-1.2, 1.2)),
                                          float(rng.uniform(-1.2, 1.2)), float(rng.uniform(0, 6.28))),
                                    camera, int(rng.integers(0, 2)), int(rng.integers(0, len(mats)))))
    objects = np.array(objs)
    return scene, camera, objects


def textured_scene(width=512, height=384, seed=12, n_random=30):
    """a11 test scene (material.hlsli:30-154 with textures): the patch instances of ``masked_scene`` drawn OPAQUE with textured
    PBR materials - sRGB base colour maps (all mip levels, three address-mode pairs), a tangent-space normal map with a normal
    scale, an occlusion / roughness / metallic map with and without occlusion - next to untextured ones."""
    scene, camera, objects = masked_scene(width, height, seed, n_random=n_random)
    rng = np.random.default_rng(seed + 100)
    texs = alpha_test_textures(seed)                      # [None (builtin white), leaf 64x64, link 128x32, checker 16x16], random RGB
    yy, xx = np.mgrid[0:64, 0:64]
    nx = 0.55 * np.sin(xx * (2 * math.pi / 16.0)) * np.cos(yy * (2 * math.pi / 32.0))
    ny = 0.55 * np.cos(xx * (2 * math.pi / 64.0)) * np.sin(yy * (2 * math.pi / 8.0))
    nmap = np.zeros((64, 64, 4), dtype=np.uint8)
    nmap[..., 0] = np.rint((nx * 0.5 + 0.5) * 255.0); nmap[..., 1] = np.rint((ny * 0.5 + 0.5) * 255.0); nmap[..., 2] = 255; nmap[..., 3] = 255
    yy, xx = np.mgrid[0:32, 0:32]
    mr = np.zeros((32, 32, 4), dtype=np.uint8)
    mr[..., 0] = 120 + (xx * 4) % 130                       # occlusion
    mr[..., 1] = rng.integers(30, 250, (32, 32))            # roughness
    mr[..., 2] = ((xx // 4 + yy // 4) % 2) * 255            # metallic checker
    mr[..., 3] = 255
    texs = texs + [abi.box_mip_chain(nmap), abi.box_mip_chain(mr)]          # ids 4, 5
    scene.textures = abi.TextureTable(texs, samplers=((abi.ADDRESS_REPEAT, abi.ADDRESS_REPEAT),
                                                      (abi.ADDRESS_CLAMP_TO_EDGE, abi.ADDRESS_MIRRORED_REPEAT),
                                                      (abi.ADDRESS_MIRRORED_REPEAT, abi.ADDRESS_CLAMP_TO_EDGE)), srgb=(1, 2, 3))
    mats = scene.materials
    for i in range(len(mats)):
        m = mats[i]
        m['alphaMode'], m['bTwoSided'] = 0, i % 2
        m['baseColorFactor'] = (0.9, 0.8 + 0.02 * i, 0.7, 1.0)
        v = i % 5
        m['baseColorId'], m['baseColorSampler'] = (0 if v == 4 else 1 + i % 3), i % 3
        m['normalTexture'], m['normalSampler'] = (4 if v in (1, 2) else abi.UNVALID_ID), (i + 1) % 3
        m['normalFactorScale'] = 0.7 if v == 1 else 1.3
        m['metallicRoughnessTexture'], m['metallicRoughnessSampler'] = (5 if v in (2, 3) else abi.UNVALID_ID), (i + 2) % 3
        m['bExistOcclusion'], m['occlusionTextureStrength'] = (1 if v == 2 else 0), 0.8
        m['roughnessFactor'], m['metallicFactor'] = 0.3 + 0.05 * i, 0.1 * (i % 4)
    return scene, camera, objects
