"""The N1 builder held against the REAL reference code: chord's vendored meshoptimizer 0.21 (source/asset/meshoptimizer),
compiled from its sources in place into oracle/_ref/libmeshopt_ref.so (oracle/Makefile target `ref`) and called exactly as
NaniteBuilder calls it (nanite_builder.cpp:443-486, :811-835).  METIS (a Windows .lib in the reference) stays substituted.

Skipped when neither /root/reference nor a prebuilt oracle/_ref is available."""
import ctypes as C

import numpy as np
import pytest

import oracle
from chord_b200 import abi, nanite_builder as nb

REF = oracle.meshopt_ref()
pytestmark = pytest.mark.skipif(REF is None, reason='oracle/_ref/libmeshopt_ref.so not available (no /root/reference)')

MAX_V, MAX_T = abi.MESHLET_MAX_VERTICES, abi.MESHLET_MAX_TRIANGLES


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def ref_build_meshlets(pos, tris, cone_weight=0.0):
    """meshopt_buildMeshlets + meshopt_optimizeMeshlet as in nanite_builder.cpp:443-478 -> list of (global vertex ids (V,), local triangles (T, 3))."""
    idx = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    bound = REF.meshopt_ref_build_meshlets_bound(C.c_size_t(len(idx)), C.c_size_t(MAX_V), C.c_size_t(MAX_T))
    ms = np.zeros((bound, 4), dtype=np.uint32)
    mv = np.zeros(bound * MAX_V, dtype=np.uint32)
    mt = np.zeros(bound * MAX_T * 3, dtype=np.uint8)
    n = REF.meshopt_ref_build_meshlets(_ptr(idx), C.c_size_t(len(idx)), _ptr(pos), C.c_size_t(len(pos)), C.c_size_t(12), C.c_size_t(MAX_V), C.c_size_t(MAX_T),
                                       C.c_float(cone_weight), _ptr(ms), C.c_size_t(bound), _ptr(mv), _ptr(mt))
    assert 0 < n <= bound
    out = []
    for vo, to, vc, tc in ms[:n]:
        out.append((mv[vo:vo + vc].copy(), mt[to:to + 3 * tc].reshape(-1, 3).copy()))
    return out


def ref_bounds(pos, verts, local_tris):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    v = np.ascontiguousarray(verts, dtype=np.uint32)
    t = np.ascontiguousarray(local_tris, dtype=np.uint8).reshape(-1)
    out = np.zeros(11, dtype=np.float32)
    REF.meshopt_ref_meshlet_bounds(_ptr(v), _ptr(t), C.c_size_t(len(t) // 3), _ptr(pos), C.c_size_t(len(pos)), C.c_size_t(12), _ptr(out))
    return dict(center=out[0:3], radius=out[3], apex=out[4:7], axis=out[7:10], cutoff=out[10])


def ref_simplify(pos, tris, target_index_count, target_error=3.0e38):
    """meshopt_simplifyWithAttributes(LockBorder | Sparse | ErrorAbsolute) without attributes -> (triangles, absolute error)."""
    idx = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    dst = np.zeros(len(idx), dtype=np.uint32)
    err = C.c_float()
    n = REF.meshopt_ref_simplify(_ptr(dst), _ptr(idx), C.c_size_t(len(idx)), _ptr(pos), C.c_size_t(len(pos)), C.c_size_t(12),
                                 None, C.c_size_t(0), None, C.c_size_t(0), C.c_size_t(target_index_count), C.c_float(target_error), C.byref(err))
    return dst[:n].reshape(-1, 3).astype(np.int64), float(err.value)


def test_reference_meshoptimizer_is_the_vendored_version():
    assert REF.meshopt_ref_version() == 210                       # source/asset/meshoptimizer/meshoptimizer.h:15


def test_normal_cones_match_meshopt_compute_meshlet_bounds():
    """chord_b200/nanite_builder.py::_cone restates meshopt_computeMeshletBounds (nanite_builder.cpp:480-486 stores its
    cone_axis / cone_cutoff / cone_apex in the GPU meshlet): same axis, cutoff and apex on our own meshlets."""
    pos, tris, _ = nb.terrain_mesh(n=40, seed=5)
    checked = 0
    for ts in nb.build_meshlets(tris, pos):
        mt = tris[ts]
        verts, inv = np.unique(mt, return_inverse=True)
        want = ref_bounds(pos, verts, inv.reshape(-1, 3))
        axis, cutoff, apex = nb._cone(pos, mt)
        if want['cutoff'] >= 1.0:                                 # degenerate cone on both sides: never culls
            assert cutoff >= 1.0
            continue
        assert np.allclose(axis, want['axis'], atol=2e-5), (axis, want['axis'])
        assert abs(cutoff - want['cutoff']) <= 2e-5
        assert np.allclose(apex, want['apex'], atol=2e-3 * max(1.0, np.abs(want['apex']).max())), (apex, want['apex'])
        checked += 1
    assert checked >= 10


def test_meshlets_are_comparable_to_meshopt_build_meshlets():
    """Both builders partition the mesh (every triangle exactly once) within the format limits; ours needs no more than 15 %
    more meshlets than meshopt's clusterizer on the same mesh (fill rate matters: one meshlet = one warp iteration)."""
    pos, tris, _ = nb.terrain_mesh(n=48, seed=1)
    ours = nb.build_meshlets(tris, pos)
    ref = ref_build_meshlets(pos, tris)
    # reference output: a partition, decoded through its own vertex tables
    seen = set()
    for verts, lt in ref:
        assert len(verts) <= MAX_V and len(lt) <= MAX_T
        for t in verts[lt.astype(np.int64)]:
            key = tuple(sorted(int(v) for v in t))
            assert key not in seen
            seen.add(key)
    assert seen == {tuple(sorted(int(v) for v in t)) for t in tris}
    assert sorted(np.concatenate(ours).tolist()) == list(range(len(tris)))
    assert len(ours) <= 1.15 * len(ref), (len(ours), len(ref))
    # a scene made of the REFERENCE's meshlets through our packer renders exactly like one made of ours
    from chord_b200 import camera as cam
    import oracle as orc
    scene_ours = nb.host_scene(nb.build(pos, tris, max_lods=1))
    w, h = 320, 180
    c = cam.Camera((0.0, 12.0, 26.0), (0.0, -0.45, -1.0), width=w, height=h)
    view = cam.fill_view(c); iview = cam.fill_instance_view(view)
    objs = np.array([cam.fill_object(np.eye(4), c, 0, 0)])

    def depth_image(scene, n_meshlets):
        fr = orc.Frame(scene, view, iview, objs, w, h)
        cmds = np.zeros(n_meshlets, dtype=abi.DrawCmd)
        cmds['meshletId'] = np.arange(n_meshlets); cmds['slot'] = np.arange(n_meshlets)
        vis, _ = fr.render_mesh(cmds)
        return vis >> np.uint64(32)

    d_ours = depth_image(scene_ours, len(scene_ours.datas[0]['meshlets']))
    # pack the reference's meshlets into the asset layout by hand (asset_gltf_helper.cpp:522-548: [remap][packed triangles])
    recs = np.zeros(len(ref), dtype=abi.GPUGLTFMeshlet)
    data, off = [], 0
    for i, (verts, lt) in enumerate(ref):
        lt = lt.astype(np.uint32)
        data += [verts.astype(np.uint32), (lt[:, 0] | (lt[:, 1] << 8) | (lt[:, 2] << 16)).astype(np.uint32)]
        b = ref_bounds(pos, verts, lt)
        recs[i]['posMin'], recs[i]['posMax'] = pos[verts].min(axis=0), pos[verts].max(axis=0)
        recs[i]['dataOffset'], recs[i]['vertexTriangleCount'] = off, abi.pack_vertex_triangle_count(len(verts), len(lt))
        recs[i]['coneAxis'], recs[i]['coneCutOff'], recs[i]['coneApex'] = b['axis'], b['cutoff'], b['apex']
        off += len(verts) + len(lt)
    d0 = dict(scene_ours.datas[0])
    d0['meshlets'], d0['meshletData'] = recs, np.concatenate(data)
    scene_ref = abi.HostScene([d0], scene_ours.primitives, scene_ours.materials)
    d_ref = depth_image(scene_ref, len(ref))
    assert np.array_equal(d_ours, d_ref) and (d_ours != 0).mean() > 0.3


def test_simplify_is_comparable_to_meshopt_simplify():
    """One cluster group: both simplifiers keep the open border (LockBorder), reach the 50 % target within the 80 % acceptance
    bound (nanite_builder.cpp:840) and report an absolute error of the same order on the same patch."""
    n = 24
    pos, tris, _ = nb.terrain_mesh(n=n, seed=2)
    cells = np.array([i * n + j for i in range(4, 20) for j in range(4, 20)])                 # a connected 16 x 16 cell patch
    group = np.concatenate([tris[cells], tris[n * n + cells]])
    target = int(len(group) * nb.SIMPLIFY_TARGET)

    def border(ts):
        from collections import Counter
        cnt = Counter()
        for a, b, c in ts:
            for e in ((a, b), (b, c), (c, a)):
                cnt[(min(e), max(e))] += 1
        return {e for e, k in cnt.items() if k == 1}

    ours, err_ours = nb._simplify(pos, group, target)
    ref, err_ref = ref_simplify(pos, group, target * 3)
    assert 0 < len(ours) < len(group) * nb.MIN_REDUCE and 0 < len(ref) < len(group) * nb.MIN_REDUCE
    assert border([tuple(int(v) for v in t) for t in ours]) == border([tuple(int(v) for v in t) for t in group])
    assert border([tuple(int(v) for v in t) for t in ref]) == border([tuple(int(v) for v in t) for t in group])
    assert err_ref > 0 and err_ours > 0
    assert 0.25 <= err_ours / err_ref <= 4.0, (err_ours, err_ref)          # same metric (root of the weighted mean squared plane distance)


def test_native_builder_against_meshopt():
    """The C++ core (chord_b200/csrc/nanite_builder.cpp) against the reference's own meshoptimizer build on the same mesh:
    LOD-0 meshlet count within 15 % of meshopt_buildMeshlets', cones equal to meshopt_computeMeshletBounds' on the native
    meshlets, group simplification error of the first level of the same order as meshopt_simplify's on the same patch."""
    pos, tris, uv = nb.terrain_mesh(n=96, seed=1)
    a = nb.build_native(pos, tris, uv0=uv)
    d = a['datas']
    ms, data = d['meshlets'], d['meshletData']
    ref = ref_build_meshlets(pos, tris)
    n0 = int((ms['lod'] == 0).sum())
    assert n0 <= 1.15 * len(ref), (n0, len(ref))
    checked = 0
    for m in np.nonzero(ms['lod'] == 0)[0][:40]:
        v, t = int(ms['vertexTriangleCount'][m] & 0xFF), int((ms['vertexTriangleCount'][m] >> 8) & 0xFF)
        verts = data[ms['dataOffset'][m]:ms['dataOffset'][m] + v]
        packed = data[ms['dataOffset'][m] + v:ms['dataOffset'][m] + v + t]
        lt = np.stack([packed & 0xFF, (packed >> 8) & 0xFF, (packed >> 16) & 0xFF], axis=1)
        want = ref_bounds(pos, verts, lt)
        if want['cutoff'] >= 1.0:
            assert ms['coneCutOff'][m] >= 1.0
            continue
        assert np.allclose(ms['coneAxis'][m], want['axis'], atol=2e-5)
        assert abs(ms['coneCutOff'][m] - want['cutoff']) <= 2e-5
        assert np.allclose(ms['coneApex'][m], want['apex'], atol=2e-3 * max(1.0, np.abs(want['apex']).max()))
        checked += 1
    assert checked >= 10
    # level-1 errors (native QEM with the group border locked) vs meshopt_simplify on one 16 x 16-cell patch of the same mesh
    n = 96
    cells = np.array([i * n + j for i in range(8, 24) for j in range(8, 24)])
    group = np.concatenate([tris[cells], tris[n * n + cells]])
    _, err_ref = ref_simplify(pos, group, int(len(group) * nb.SIMPLIFY_TARGET) * 3)
    e1 = d['groups']['error'][(d['groups']['error'] > 0)]
    lvl1 = np.unique(ms['lod'])
    err_native = float(np.median(e1[e1 <= np.percentile(e1, 40)]))               # the smallest positive errors are level 1's
    assert err_ref > 0 and 0.2 <= err_native / err_ref <= 5.0, (err_native, err_ref)
