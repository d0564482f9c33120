"""Meshlet-DAG builder: triangle mesh -> chord's GPU asset layout.

Restates the pipeline of source/asset/nanite_builder.cpp:432-921 (NaniteBuilder::build) with self-contained pieces where
the reference calls third-party code that is not available offline:

    reference                                            here
    ------------------------------------------------    ---------------------------------------------------------
    meshopt_buildMeshlets (255 verts / 128 tris)          greedy region growing over triangle adjacency
    METIS_PartGraphKway groups of <= 4 meshlets (:584)    greedy grouping by shared boundary edges
    meshopt_simplifyWithAttributes, LockBorder (:806)     quadric edge collapse onto existing vertices, group border locked
    accept if <= 80 % of the triangles remain (:840)      same
    parent error = simplify error + max child (:851)      same; centre = AABB centre of the simplified vertices (:857-866)
    meshopt_computeMeshletBounds (cone)                   normal cone with meshopt's apex construction
    buildBVHTree (:313-416)                               synth.build_bvh

Offline tool (the reference runs it at asset-import time), numpy + heapq; sized for meshes of a few 10^4 triangles.
The per-frame path never calls it.  METIS / meshopt are not bit-reproduced (parity unpinned there too):
any valid DAG exercises the same per-frame kernels.
"""
import heapq
from collections import defaultdict

import numpy as np

from . import abi
from .synth import FLT_MAX, build_bvh, default_material

MAX_V, MAX_T = abi.MESHLET_MAX_VERTICES, abi.MESHLET_MAX_TRIANGLES
SIMPLIFY_TARGET, MIN_REDUCE, MAX_LODS = 0.5, 0.8, 12


def _tri_adjacency(tris):
    """triangle -> neighbours across shared edges."""
    edge = defaultdict(list)
    for t, (a, b, c) in enumerate(tris):
        for e in ((a, b), (b, c), (c, a)):
            edge[(min(e), max(e))].append(t)
    adj = [[] for _ in range(len(tris))]
    for ts in edge.values():
        for i in ts:
            for j in ts:
                if i != j:
                    adj[i].append(j)
    return adj, edge


def build_meshlets(tris, pos=None):
    """Greedy region growing: clusters of <= 128 triangles / <= 255 vertices, grown breadth first over shared edges.  When the
    frontier runs dry before the meshlet is full and positions are given, growth continues from the nearest unused triangle
    (meshopt_buildMeshlets does the same through its kd-tree), which keeps the fill rate near 100 % -- one meshlet is one warp
    iteration of the raster, so half-empty meshlets cost throughput.  Returns a list of triangle-index arrays."""
    tris = np.asarray(tris)
    adj, _ = _tri_adjacency(tris)
    used = np.zeros(len(tris), dtype=bool)
    cent = None if pos is None else np.asarray(pos, dtype=np.float64)[tris].mean(axis=1)
    out = []
    order = np.argsort(tris.min(axis=1), kind='stable')          # deterministic seeds
    cursor = 0
    while True:
        while cursor < len(order) and used[order[cursor]]:
            cursor += 1
        if cursor >= len(order):
            break
        seed = order[cursor]
        if cent is not None and out:                              # start next to the previous meshlet: compact neighbours
            free = np.nonzero(~used)[0]
            seed = free[np.argmin(((cent[free] - cent[out[-1]].mean(axis=0)) ** 2).sum(axis=1))]
        cluster, verts, frontier = [], set(), [seed]
        while len(cluster) < MAX_T:
            if not frontier:
                if cent is None or not cluster:
                    break
                free = np.nonzero(~used)[0]
                if len(free) == 0:
                    break
                frontier = [free[np.argmin(((cent[free] - cent[cluster].mean(axis=0)) ** 2).sum(axis=1))]]
            t = frontier.pop(0)
            if used[t]:
                continue
            nv = verts | set(tris[t].tolist())
            if len(nv) > MAX_V:
                if not frontier:
                    break
                continue
            used[t] = True
            cluster.append(t)
            verts = nv
            frontier += [n for n in adj[t] if not used[n]]
        out.append(np.array(cluster, dtype=np.int64))
    return out


def _group_meshlets(meshlet_tris, tris):
    """Groups of <= 4 meshlets sharing the most boundary edges (METIS stand-in, nanite_builder.cpp:584-726)."""
    edge_owner = defaultdict(set)
    for m, ts in enumerate(meshlet_tris):
        for t in ts:
            a, b, c = tris[t]
            for e in ((a, b), (b, c), (c, a)):
                edge_owner[(min(e), max(e))].add(m)
    link = defaultdict(lambda: defaultdict(int))
    for owners in edge_owner.values():
        owners = list(owners)
        for i in owners:
            for j in owners:
                if i != j:
                    link[i][j] += 1
    free = set(range(len(meshlet_tris)))
    groups = []
    for m in range(len(meshlet_tris)):
        if m not in free:
            continue
        free.discard(m)
        g = [m]
        while len(g) < abi.CLUSTER_GROUP_MERGE_MAX:
            cand = defaultdict(int)
            for i in g:
                for j, w in link[i].items():
                    if j in free:
                        cand[j] += w
            if not cand:
                break
            best = max(sorted(cand), key=lambda j: cand[j])
            free.discard(best)
            g.append(best)
        groups.append(g)
    return groups


def _simplify(pos, tris, target_count):
    """Quadric-error edge collapse onto existing vertices with the open border locked (meshopt_SimplifyLockBorder).
    Returns (new triangle array, error in object units)."""
    tris = [tuple(int(v) for v in t) for t in tris]
    edge_count = defaultdict(int)
    for a, b, c in tris:
        for e in ((a, b), (b, c), (c, a)):
            edge_count[(min(e), max(e))] += 1
    locked = set()
    for (a, b), n in edge_count.items():
        if n == 1:
            locked.add(a); locked.add(b)
    verts = sorted({v for t in tris for v in t})
    Q = {v: np.zeros((4, 4)) for v in verts}
    W = {v: 0.0 for v in verts}
    for a, b, c in tris:
        p0, p1, p2 = pos[a].astype(np.float64), pos[b].astype(np.float64), pos[c].astype(np.float64)
        n = np.cross(p1 - p0, p2 - p0)
        area = np.linalg.norm(n)
        if area == 0:
            continue
        n /= area
        pl = np.append(n, -np.dot(n, p0))
        wgt = np.sqrt(area)                       # meshoptimizer's quadricFromTriangle: sqrt(|cross|) so that the error scales linearly
        K = np.outer(pl, pl) * wgt
        for v in (a, b, c):
            Q[v] += K
            W[v] += wgt
    vt = defaultdict(set)
    for i, t in enumerate(tris):
        for v in t:
            vt[v].add(i)
    alive = [True] * len(tris)

    def cost(src, dst):
        """weighted MEAN squared distance of dst to the planes accumulated around src and dst (meshoptimizer's quadricError:
        v Q v / Q.w), so that its square root is a length in object units"""
        q = Q[src] + Q[dst]
        w = W[src] + W[dst]
        p = np.append(pos[dst].astype(np.float64), 1.0)
        return (float(abs(p @ q @ p)) / w if w > 0 else 0.0), 0.0

    heap = []
    for (a, b) in edge_count:
        for src, dst in ((a, b), (b, a)):
            if src in locked:
                continue
            c, _ = cost(src, dst)
            heapq.heappush(heap, (c, src, dst))
    remap = {}

    def find(v):
        while v in remap:
            v = remap[v]
        return v

    n_alive, max_err = len(tris), 0.0
    while heap and n_alive > target_count:
        c, src, dst = heapq.heappop(heap)
        if src in remap or find(dst) != dst or src == dst:
            continue
        # reject collapses that flip a surviving triangle
        flip = False
        for ti in vt[src]:
            if not alive[ti]:
                continue
            t = [find(v) for v in tris[ti]]
            if dst in t:
                continue
            old = [pos[v].astype(np.float64) for v in t]
            new = [pos[dst if v == src else v].astype(np.float64) for v in t]
            n0 = np.cross(old[1] - old[0], old[2] - old[0]); n1 = np.cross(new[1] - new[0], new[2] - new[0])
            if np.dot(n0, n1) <= 1e-12 * max(np.dot(n0, n0), 1e-30):
                flip = True
                break
        if flip:
            continue
        for ti in list(vt[src]):
            if not alive[ti]:
                continue
            t = [find(v) for v in tris[ti]]
            if dst in t:
                alive[ti] = False
                n_alive -= 1
            else:
                vt[dst].add(ti)
        remap[src] = dst
        Q[dst] = Q[dst] + Q[src]
        W[dst] = W[dst] + W[src]
        max_err = max(max_err, c)
        nb = set()
        for ti in vt[dst]:
            if alive[ti]:
                nb |= {find(v) for v in tris[ti]}
        nb.discard(dst)
        for v in nb:
            if v not in locked:
                heapq.heappush(heap, (cost(v, dst)[0], v, dst))
            if dst not in locked:
                heapq.heappush(heap, (cost(dst, v)[0], dst, v))
    out = []
    for i, t in enumerate(tris):
        if alive[i]:
            a, b, c = (find(v) for v in t)
            if a != b and b != c and a != c:
                out.append((a, b, c))
    # the collapse cost is a mean squared distance: report its root, a length in object units (meshopt_SimplifyErrorAbsolute)
    return np.array(out, dtype=np.int64).reshape(-1, 3), float(np.sqrt(max_err))


def _bounding_sphere(points):
    """meshoptimizer's computeBoundingSphere (Ritter: the most distant pair of the six axis extrema, then grown point by point
    in input order), float32 like the original."""
    F = np.float32
    p = np.asarray(points, dtype=F)
    pmin, pmax = p.argmin(axis=0), p.argmax(axis=0)                       # first extremum per axis, like the strict < / > scan
    d2 = [F(((p[pmax[k]] - p[pmin[k]]) ** 2).sum()) for k in range(3)]
    k = 0
    for axis in range(3):
        if d2[axis] > d2[k]:
            k = axis
    center = ((p[pmin[k]] + p[pmax[k]]) / F(2)).astype(F)
    radius = F(np.sqrt(d2[k]) / F(2)) if d2[k] > 0 else F(0)
    for q in p:
        dd = F(((q - center) ** 2).sum())
        if dd > radius * radius:
            d = F(np.sqrt(dd))
            kk = F(0.5) + (radius / d) / F(2)
            center = (center * kk + q * (F(1) - kk)).astype(F)
            radius = F((radius + d) / F(2))
    return center, radius


def _cone(pos, tris):
    """meshopt_computeClusterBounds normal cone, restated (meshopt_clusterizer.cpp:712-850; checked against the reference's
    own build of it in tests/test_nanite_builder_ref.py): the axis is the centre of the bounding sphere of the triangle
    normals, the apex lies on the ray from the bounding-sphere centre of the corners against the axis.  -> (axis, cutoff, apex)."""
    F = np.float32
    p0, p1, p2 = pos[tris[:, 0]].astype(F), pos[tris[:, 1]].astype(F), pos[tris[:, 2]].astype(F)
    n = np.cross(p1 - p0, p2 - p0).astype(F)
    area = np.sqrt((n * n).sum(axis=1)).astype(F)
    ok = area > 0
    if not ok.any():
        return np.zeros(3), 0.0, np.zeros(3)                          # degenerate cluster: cone data is 0 (trivial reject upstream)
    n = (n[ok] / area[ok, None]).astype(F)
    corners = np.stack([p0[ok], p1[ok], p2[ok]], axis=1).reshape(-1, 3)
    center, _ = _bounding_sphere(corners)
    axis, _ = _bounding_sphere(n)
    la = F(np.sqrt((axis * axis).sum()))
    axis = (axis * (F(0) if la == 0 else F(1) / la)).astype(F)
    dp = (n * axis).sum(axis=1).astype(F)
    mindp = min(F(1), dp.min())
    if mindp <= F(0.1):
        return np.zeros(3), 1.0, np.zeros(3)                          # wider than ~168 degrees: never culls (cutoff 1, axis 0)
    maxt = max(F(0), F((((center[None] - p0[ok]) * n).sum(axis=1) / dp).max()))
    return axis.astype(np.float64), float(np.sqrt(F(1) - mindp * mindp)), (center - axis * maxt).astype(np.float64)


class NaniteAsset:
    """Result: arrays in the reference layout (one GLTFPrimitiveBuffer worth of data)."""


def build(positions, triangles, normals=None, uv0=None, max_lods=MAX_LODS):
    """positions (N,3) f32, triangles (T,3) -> dict ready for abi.HostScene (one primitive) + stats."""
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    tris = np.ascontiguousarray(triangles, dtype=np.int64)
    if normals is None:
        nacc = np.zeros_like(pos, dtype=np.float64)
        fn = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]])
        for k in range(3):
            np.add.at(nacc, tris[:, k], fn)
        normals = (nacc / np.maximum(np.linalg.norm(nacc, axis=1, keepdims=True), 1e-30)).astype(np.float32)
    if uv0 is None:
        uv0 = np.zeros((len(pos), 2), dtype=np.float32)
    meshlets = []          # dict(tris (k,3) global vertex ids, lod, error, center, parent_error, parent_center)
    cur = [dict(tris=tris[ts], lod=0, error=-1.0, center=np.zeros(3)) for ts in build_meshlets(tris, pos)]
    stats = []
    for lod in range(max_lods - 1):
        stats.append((lod, len(cur), int(sum(len(m['tris']) for m in cur))))
        if len(cur) <= 1:
            break
        all_tris = np.concatenate([m['tris'] for m in cur])
        offs = np.cumsum([0] + [len(m['tris']) for m in cur])
        groups = _group_meshlets([np.arange(offs[i], offs[i + 1]) for i in range(len(cur))], all_tris)
        nxt = []
        for g in groups:
            gt = np.concatenate([cur[i]['tris'] for i in g])
            simp, err = _simplify(pos, gt, int(len(gt) * SIMPLIFY_TARGET))
            if 0 < len(simp) < len(gt) * MIN_REDUCE:
                passed = max([0.0] + [cur[i]['error'] for i in g])
                cerr = err + passed
                used = np.unique(simp)
                center = 0.5 * (pos[used].max(axis=0) + pos[used].min(axis=0))
                for i in g:
                    cur[i]['parent_error'], cur[i]['parent_center'] = cerr, center
                for ts in build_meshlets(simp, pos):
                    nxt.append(dict(tris=simp[ts], lod=lod + 1, error=cerr, center=center))
        meshlets += cur
        if not nxt:
            cur = []
            break
        cur = nxt
    meshlets += cur
    # ---- pack (asset_gltf_helper.cpp:522-548)
    recs = np.zeros(len(meshlets), dtype=abi.GPUGLTFMeshlet)
    data, off = [], 0
    for i, m in enumerate(meshlets):
        vs, inv = np.unique(m['tris'], return_inverse=True)
        local = inv.reshape(-1, 3).astype(np.uint32)
        assert len(vs) <= MAX_V and len(local) <= MAX_T
        data.append(vs.astype(np.uint32))
        data.append((local[:, 0] | (local[:, 1] << 8) | (local[:, 2] << 16)).astype(np.uint32))
        axis, cutoff, apex = _cone(pos, m['tris'])
        r = recs[i]
        r['posMin'], r['posMax'] = pos[vs].min(axis=0), pos[vs].max(axis=0)
        r['dataOffset'], r['vertexTriangleCount'] = off, abi.pack_vertex_triangle_count(len(vs), len(local))
        r['coneAxis'], r['coneCutOff'], r['coneApex'], r['lod'] = axis, cutoff, apex, m['lod']
        off += len(vs) + len(local)
    # ---- cluster groups: meshlets sharing own + parent spheres, <= 4 each (nanite_builder.cpp:330-395)
    gmap, glist = {}, []
    for i, m in enumerate(meshlets):
        pe = m.get('parent_error', FLT_MAX)
        pc = m.get('parent_center', m['center'])
        key = (np.float32(m['error']).tobytes(), np.asarray(m['center'], np.float32).tobytes(), np.float32(pe).tobytes(), np.asarray(pc, np.float32).tobytes())
        gi = gmap.get(key)
        if gi is None or len(glist[gi][0]) >= abi.CLUSTER_GROUP_MERGE_MAX:
            gmap[key] = len(glist)
            glist.append(([i], m['error'], m['center'], pe, pc))
        else:
            glist[gi][0].append(i)
    groups = np.zeros(len(glist), dtype=abi.GPUGLTFMeshletGroup)
    gidx = []
    for k, (ms, e, c, pe, pc) in enumerate(glist):
        groups[k]['clusterPosCenter'], groups[k]['error'] = c, e
        groups[k]['parentPosCenter'], groups[k]['parentError'] = pc, pe
        groups[k]['meshletOffset'], groups[k]['meshletCount'] = len(gidx), len(ms)
        gidx += ms
    groups, gidx, nodes = build_bvh(groups, np.array(gidx, dtype=np.uint32))
    prim = np.zeros(1, dtype=abi.GLTFPrimitiveBuffer)
    prim['posMin'], prim['posMax'], prim['posAverage'] = pos.min(axis=0), pos.max(axis=0), pos.mean(axis=0)
    prim['vertexCount'], prim['meshletGroupCount'] = len(pos), len(groups)
    tang = np.tile(np.array([1, 0, 0, 1], dtype=np.float32), (len(pos), 1))
    datas = dict(meshlets=recs, positions=pos, normals=normals, uv0=uv0, tangents=tang, meshletData=np.concatenate(data),
                 groups=groups, groupIndices=gidx, bvhNodes=nodes)
    lod0 = int((recs['lod'] == 0).sum())
    return dict(datas=datas, primitive=prim, lod0MeshletCount=lod0, meshletGroupCount=len(groups), lod_stats=stats,
                lods=int(recs['lod'].max()) + 1)


_native = None


def _native_lib():
    """chord_b200/libchord_builder.so (chord_b200/csrc/nanite_builder.cpp, built by chord_b200/build.py)."""
    global _native
    if _native is None:
        import ctypes as C
        import os
        from . import build as _build
        L = C.CDLL(_build.build_builder())
        for n in ('chord_builder_meshlet_count', 'chord_builder_data_count', 'chord_builder_lod_count'):
            getattr(L, n).restype = C.c_uint32
        L.chord_builder_free.restype = None
        _native = L
    return _native


def build_native(positions, triangles, normals=None, uv0=None, max_lods=MAX_LODS):
    """Same result layout as ``build`` with the hot loops (meshlet clustering, grouping, quadric simplification, cones, packing)
    in C++ (chord_b200/csrc/nanite_builder.cpp): millions of triangles in minutes.  Cluster groups and the BVH are formed here
    from the per-meshlet (error, centre, parent error, parent centre) the native core returns."""
    import ctypes as C
    L = _native_lib()
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    idx = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
    h = C.c_void_p()
    rc = L.chord_builder_build(C.c_void_p(pos.ctypes.data), C.c_uint32(len(pos)), C.c_void_p(idx.ctypes.data), C.c_uint32(len(idx)), C.c_uint32(max_lods), C.byref(h))
    if rc != 0:
        raise ValueError('chord_builder_build: bad input (index out of range / empty mesh)')
    try:
        nm, nd, nl = L.chord_builder_meshlet_count(h), L.chord_builder_data_count(h), L.chord_builder_lod_count(h)
        recs = np.zeros(nm, dtype=abi.GPUGLTFMeshlet)
        data = np.zeros(nd, dtype=np.uint32)
        err, perr = np.zeros(nm, dtype=np.float32), np.zeros(nm, dtype=np.float32)
        cen, pcen = np.zeros((nm, 3), dtype=np.float32), np.zeros((nm, 3), dtype=np.float32)
        lm, lt = np.zeros(nl, dtype=np.uint32), np.zeros(nl, dtype=np.uint32)
        p = lambda a: C.c_void_p(a.ctypes.data)
        L.chord_builder_copy(h, p(recs), p(data), p(err), p(cen), p(perr), p(pcen), p(lm), p(lt))
    finally:
        L.chord_builder_free(h)
    if normals is None:
        nacc = np.zeros_like(pos, dtype=np.float64)
        t64 = idx.astype(np.int64)
        fn = np.cross(pos[t64[:, 1]] - pos[t64[:, 0]], pos[t64[:, 2]] - pos[t64[:, 0]])
        for k in range(3):
            np.add.at(nacc, t64[:, k], fn)
        normals = (nacc / np.maximum(np.linalg.norm(nacc, axis=1, keepdims=True), 1e-30)).astype(np.float32)
    if uv0 is None:
        uv0 = np.zeros((len(pos), 2), dtype=np.float32)
    # ---- cluster groups: meshlets sharing own + parent spheres, <= 4 each (nanite_builder.cpp:330-395), vectorised: sort by key
    key = np.concatenate([err.view(np.uint32)[:, None], cen.view(np.uint32), perr.view(np.uint32)[:, None], pcen.view(np.uint32)], axis=1)
    order = np.lexsort(key.T[::-1])
    sk = key[order]
    new = np.ones(nm, dtype=bool)
    new[1:] = (sk[1:] != sk[:-1]).any(axis=1)
    run_id = np.cumsum(new) - 1
    pos_in_run = np.arange(nm) - np.maximum.accumulate(np.where(new, np.arange(nm), 0))
    gid_sorted = run_id * 0
    # split runs into chunks of 4
    chunk = pos_in_run // abi.CLUSTER_GROUP_MERGE_MAX
    gkey = np.stack([run_id, chunk], axis=1)
    gnew = np.ones(nm, dtype=bool)
    gnew[1:] = (gkey[1:] != gkey[:-1]).any(axis=1)
    gid_sorted = np.cumsum(gnew) - 1
    ng = int(gid_sorted[-1]) + 1 if nm else 0
    first = np.nonzero(gnew)[0]
    groups = np.zeros(ng, dtype=abi.GPUGLTFMeshletGroup)
    rep = order[first]
    groups['clusterPosCenter'], groups['error'] = cen[rep], err[rep]
    groups['parentPosCenter'], groups['parentError'] = pcen[rep], perr[rep]
    groups['meshletOffset'] = first
    groups['meshletCount'] = np.diff(np.append(first, nm))
    gidx = order.astype(np.uint32)
    groups, gidx, nodes = build_bvh(groups, gidx)
    prim = np.zeros(1, dtype=abi.GLTFPrimitiveBuffer)
    prim['posMin'], prim['posMax'], prim['posAverage'] = pos.min(axis=0), pos.max(axis=0), pos.mean(axis=0)
    prim['vertexCount'], prim['meshletGroupCount'] = len(pos), len(groups)
    tang = np.tile(np.array([1, 0, 0, 1], dtype=np.float32), (len(pos), 1))
    datas = dict(meshlets=recs, positions=pos, normals=normals, uv0=uv0, tangents=tang, meshletData=data,
                 groups=groups, groupIndices=gidx, bvhNodes=nodes)
    lod0 = int((recs['lod'] == 0).sum())
    stats = [(i, int(lm[i]), int(lt[i])) for i in range(nl)]
    return dict(datas=datas, primitive=prim, lod0MeshletCount=lod0, meshletGroupCount=len(groups), lod_stats=stats,
                lods=int(recs['lod'].max()) + 1 if nm else 0)


def host_scene(asset, materials=None):
    return abi.HostScene([asset['datas']], asset['primitive'], np.array(materials if materials is not None else [default_material()]))


def terrain_mesh(n=64, size=40.0, seed=1):
    """Seeded height-field test mesh: (n+1)^2 vertices, 2 n^2 triangles, CCW seen from +y."""
    rng = np.random.default_rng(seed)
    g = np.linspace(-size / 2, size / 2, n + 1)
    xx, zz = np.meshgrid(g, g)
    ph = rng.uniform(0, 6.28, 4)
    yy = 1.5 * np.sin(xx * 0.35 + ph[0]) * np.cos(zz * 0.27 + ph[1]) + 0.6 * np.sin(xx * 0.9 + ph[2]) * np.sin(zz * 1.1 + ph[3])
    pos = np.stack([xx, yy, zz], axis=-1).reshape(-1, 3).astype(np.float32)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    tris = np.concatenate([np.stack([a, c, b], axis=1), np.stack([b, c, d], axis=1)])
    uv = np.stack([(xx / size + 0.5), (zz / size + 0.5)], axis=-1).reshape(-1, 2).astype(np.float32)
    return pos, tris, uv
