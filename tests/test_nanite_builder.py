"""Meshlet-DAG builder: structural invariants of the asset and of the LOD cut it induces."""
import numpy as np
import pytest

import oracle
from cases import terrain_case
from chord_b200 import abi, nanite_builder as nb


@pytest.fixture(scope='module', params=['python', 'native'])
def terrain(request):
    """both builders: the numpy / heapq one and the C++ core (chord_b200/csrc/nanite_builder.cpp) behind build_native"""
    return terrain_case(native=request.param == 'native')


def test_asset_structure(terrain):
    scene, asset, _ = terrain
    d = scene.datas[0]
    ms, data = d['meshlets'], d['meshletData']
    V, T = ms['vertexTriangleCount'] & 0xFF, (ms['vertexTriangleCount'] >> 8) & 0xFF
    assert V.max() <= 255 and T.max() <= 128 and V.min() >= 3 and T.min() >= 1      # base.h:429-430
    assert asset['lods'] >= 4
    # LOD 0 holds every input triangle exactly once
    pos, tris, _ = nb.terrain_mesh(n=48)
    seen = set()
    for m in ms[ms['lod'] == 0]:
        v, t = m['vertexTriangleCount'] & 0xFF, (m['vertexTriangleCount'] >> 8) & 0xFF
        remap = data[m['dataOffset']:m['dataOffset'] + v]
        for packed in data[m['dataOffset'] + v:m['dataOffset'] + v + t]:
            tri = tuple(sorted(int(remap[(packed >> s) & 0xFF]) for s in (0, 8, 16)))
            assert tri not in seen
            seen.add(tri)
    assert seen == {tuple(sorted(int(x) for x in t)) for t in tris}
    # groups: <= 4 meshlets, LOD-0 error -1, monotone errors, exactly the root groups have no parent
    g = d['groups']
    assert g['meshletCount'].max() <= 4 and g['meshletCount'].min() >= 1
    has_parent = g['parentError'] < 3e38
    assert (g['parentError'][has_parent] >= np.maximum(g['error'][has_parent], 0)).all()
    assert (~has_parent).any()
    # every triangle count halves (roughly) per level: at least 20 % reduction where a level exists (nanite_builder.cpp:840)
    stats = asset['lod_stats']
    for (l0, _, t0), (l1, _, t1) in zip(stats, stats[1:]):
        assert t1 < t0 * 0.8 + 1
    # cones are usable: never cull a meshlet seen from straight above its centre
    up = np.array([0, 1, 0.0])
    for m in ms:
        if np.linalg.norm(m['coneAxis']) == 0:
            continue
        c = 0.5 * (m['posMin'] + m['posMax']) + 50 * up
        v = m['coneApex'] - c
        assert np.dot(v / np.linalg.norm(v), m['coneAxis']) < m['coneCutOff'] + 1e-3


def test_lod_cut_is_a_cut(terrain):
    """never draw a cluster together with one of its ancestors' replacements: a selected group's parent sphere must not be
    the own sphere of another selected meshlet; and the flat loop equals the BVH walk."""
    scene, asset, frames = terrain
    d = scene.datas[0]
    lods_seen = set()
    for f in frames:
        fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, 0)          # all culls off: pure LOD cut
        post = fr.instance_culling(asset['lod0MeshletCount'] * 2)
        via_bvh, _ = fr.instance_culling_bvh(asset['lod0MeshletCount'] * 2)
        assert np.array_equal(post, via_bvh)
        assert len(post) > 0
        sel = set(int(x) for x in post['meshletId'])
        lods_seen |= set(d['meshlets']['lod'][post['meshletId']].tolist())
        own, parent = {}, {}
        for g in d['groups']:
            for k in range(g['meshletCount']):
                mid = int(d['groupIndices'][g['meshletOffset'] + k])
                own[mid] = (g['clusterPosCenter'].tobytes(), np.float32(g['error']).tobytes())
                parent[mid] = (g['parentPosCenter'].tobytes(), np.float32(g['parentError']).tobytes())
        own_keys = {own[m] for m in sel if d['meshlets']['lod'][m] > 0}
        for m in sel:
            assert parent[m] not in own_keys, 'a cluster and its simplified replacement are both selected'
    assert len(lods_seen) >= 3, lods_seen


def test_lod_cut_renders_the_same_surface(terrain):
    """the simplified levels stay within their error bound: the cut's coverage matches the LOD-0 coverage up to a thin rim"""
    scene, asset, frames = terrain
    d = scene.datas[0]
    lod0 = np.nonzero(d['meshlets']['lod'] == 0)[0]
    cmds0 = np.zeros(len(lod0), dtype=abi.DrawCmd)
    cmds0['meshletId'], cmds0['slot'] = lod0, np.arange(len(lod0))
    for f in frames[1:]:
        fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, 0)
        cut = fr.instance_culling(asset['lod0MeshletCount'] * 2)
        a, _ = fr.render_mesh(cut)
        b, _ = fr.render_mesh(cmds0)
        diff = ((a != 0) != (b != 0)).sum()
        assert diff <= 0.01 * max((b != 0).sum(), 1), diff


def test_native_builder_at_scale():
    """The C++ core on a 500 k-triangle mesh (Bistro is 3.8 M: ~17 s on 8 cores): seconds, full LOD-0 meshlets,
    every input triangle exactly once at LOD 0, format limits, halving levels, monotone errors."""
    import time
    pos, tris, uv = nb.terrain_mesh(n=500, seed=3)
    t0 = time.time()
    a = nb.build_native(pos, tris, uv0=uv)
    dt = time.time() - t0
    assert dt < 120.0, dt
    d = a['datas']
    ms, data = d['meshlets'], d['meshletData']
    V, T = ms['vertexTriangleCount'] & 0xFF, (ms['vertexTriangleCount'] >> 8) & 0xFF
    assert V.max() <= 255 and T.max() <= 128 and V.min() >= 3 and T.min() >= 1
    l0 = np.nonzero(ms['lod'] == 0)[0]
    assert T[l0].sum() == len(tris) and T[l0].mean() > 0.97 * 128            # one meshlet = one warp iteration of the raster: keep them full
    # LOD 0 is a partition of the input (vectorised decode of [remap | triangles])
    keys = []
    for m in l0[:: max(1, len(l0) // 400)]:                                    # a sample of meshlets, decoded exactly
        v, t = int(V[m]), int(T[m])
        remap = data[ms['dataOffset'][m]:ms['dataOffset'][m] + v]
        packed = data[ms['dataOffset'][m] + v:ms['dataOffset'][m] + v + t]
        tri = np.sort(np.stack([remap[packed & 0xFF], remap[(packed >> 8) & 0xFF], remap[(packed >> 16) & 0xFF]], axis=1), axis=1)
        keys.append(tri)
    keys = np.concatenate(keys).astype(np.int64)
    want = np.sort(tris, axis=1).astype(np.int64)
    enc = lambda a_: (a_[:, 0] << 42) | (a_[:, 1] << 21) | a_[:, 2]
    assert np.isin(enc(keys), enc(want)).all() and len(np.unique(enc(keys))) == len(keys)
    stats = a['lod_stats']
    assert len(stats) >= 6
    for (l_a, _, t_a), (l_b, _, t_b) in zip(stats, stats[1:]):
        assert t_b < t_a * 0.8 + 1
    g = d['groups']
    assert g['meshletCount'].max() <= 4 and g['meshletCount'].min() >= 1
    has_parent = g['parentError'] < 3e38
    assert (g['parentError'][has_parent] >= np.maximum(g['error'][has_parent], 0)).all() and (~has_parent).any()
