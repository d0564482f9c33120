"""Meshlet-DAG builder: structural invariants of the asset and of the LOD cut it induces."""
import numpy as np
import pytest

import oracle
from cases import terrain_case
from chord_b200 import abi, nanite_builder as nb


@pytest.fixture(scope='module')
def terrain():
    return terrain_case()


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
