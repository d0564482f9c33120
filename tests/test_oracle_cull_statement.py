"""K1 + K2 stated a second time, straight from the reference HLSL, in numpy float32 -- independently of oracle/chord_oracle.c.

Sources transcribed here: ``instance_culling.hlsl:47-131`` (instanceCullingCS: object-level OBB frustum test, one entry per
cluster group), ``instance_culling.hlsl:133-208`` (clusterGroupCullingCS), ``nanite_shared.hlsli:15-91``
(isMeshletGroupVisibile, isMeshletVisible), ``base.hlsli:233-305`` (transformSphere, frustumCulling) and
``base.hlsli:503-518`` (projectSphereToScreen).  Only the rounding conventions are the project's own pins (DESIGN.md
section 3): a dot product / matrix row accumulates left to right with one fma per further term, ``normalize`` divides by the
square root, ``tan(fovy / 2)`` is libm's tanf evaluated once per view.  The list the oracle emits (canonical order: object,
group, meshlet in group) must be reproduced exactly, for every switch-flag combination."""
import ctypes
import ctypes.util

import numpy as np
import pytest

import oracle
from cases import city_frames, terrain_case

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library('m') or 'libm.so.6')
_libm.tanf.restype = ctypes.c_float
_libm.tanf.argtypes = [ctypes.c_float]


def fma(a, b, c):
    """float32 fma emulated in float64: the product of two float32 is exact in float64 (48 bits)."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64) + np.asarray(c, dtype=np.float64)).astype(F)


def dot(a, b):
    """dot(a, b) over the LAST axis, left to right: a0 b0, then one fma per term."""
    r = (a[..., 0] * b[..., 0]).astype(F)
    for k in range(1, a.shape[-1]):
        r = fma(a[..., k], b[..., k], r)
    return r


def mat(field):
    """structs keep HLSL float4x4 column after column: math element (row, col) = field[col][row]."""
    return np.asarray(field, dtype=F).T


def mul_mv(M, v):
    """mul(M, v): one dot per row."""
    return np.stack([dot(np.broadcast_to(M[r], v.shape), v) for r in range(4)], axis=-1)


def mul_mm(A, B):
    return np.stack([np.stack([dot(A[i], B[:, j]) for j in range(4)]) for i in range(4)])


def point4(p):
    return np.concatenate([p, np.ones(p.shape[:-1] + (1,), dtype=F)], axis=-1)


EXTENT_FACTOR = np.array([(1, 1, 1), (-1, -1, -1), (1, 1, -1), (1, -1, 1), (-1, 1, 1), (1, -1, -1), (-1, -1, 1), (-1, 1, -1)], dtype=F)


def frustum_culled(planes, center, extent, l2w):
    """base.hlsli:275-305: culled when all 8 box corners are behind one plane (dot(n, p) > -w means in front)."""
    corners = fma(extent[..., None, :], EXTENT_FACTOR, center[..., None, :])            # (..., 8, 3)
    rs = mul_mv(l2w, point4(corners))[..., :3]                                           # (..., 8, 3)
    culled = np.zeros(center.shape[:-1], dtype=bool)
    for i in range(6):
        n, w = planes[i, :3], planes[i, 3]
        front = dot(np.broadcast_to(n, rs.shape), rs) > -w                                # (..., 8)
        culled |= ~front.any(axis=-1)
    return culled


def project_sphere(center_vs, radius, height, tan_half):
    """base.hlsli:503-518: pixel radius of the projected sphere, -1 when the eye is inside."""
    d2 = dot(center_vs, center_vs)
    r2 = (radius * radius).astype(F)
    inside = d2 <= r2
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        pix = (((F(height) * F(0.5)) / F(tan_half)) * radius / np.sqrt((d2 - r2).astype(F))).astype(F)
    return np.where(inside, F(-1.0), pix)


def cull_statement(scene, view, iview, objects, flags):
    view = np.asarray(view).reshape(())
    iview = np.asarray(iview).reshape(())
    tan_half = _libm.tanf(F(0.5) * F(view['cameraFovy']))
    height = F(view['renderDimension'][1])
    planes = np.asarray(iview['frustumPlanesRS'], dtype=F)
    w2v = mat(view['translatedWorldToView'])
    out = []
    for oi, obj in enumerate(objects):
        prim = scene.primitives[obj['GLTFPrimitiveDetail']]
        material = scene.materials[obj['GLTFMaterialData']]
        datas = scene.datas[prim['primitiveDatasBufferId']]
        l2w = mat(obj['localToTranslatedWorld'])
        # ---- instanceCullingCS: the primitive's box against the frustum
        if flags & 1:
            pmin, pmax = np.asarray(prim['posMin'], dtype=F), np.asarray(prim['posMax'], dtype=F)
            center = ((pmin + pmax) * F(0.5)).astype(F)
            if frustum_culled(planes, center, (pmax - center).astype(F), l2w):
                continue
        # ---- clusterGroupCullingCS
        g0 = int(prim['meshletGroupOffset'])
        groups = datas['groups'][g0:g0 + int(prim['meshletGroupCount'])]
        l2v = mul_mm(w2v, l2w)
        max_scale = F(obj['scaleExtractFromMatrix'][3])
        parent_err, err = groups['parentError'].astype(F), groups['error'].astype(F)
        parent_vs = mul_mv(l2v, point4(groups['parentPosCenter'].astype(F)))[..., :3]
        cluster_vs = mul_mv(l2v, point4(groups['clusterPosCenter'].astype(F)))[..., :3]
        with np.errstate(over='ignore'):
            parent_px = project_sphere(parent_vs, (max_scale * parent_err).astype(F), height, tan_half)
            cluster_px = project_sphere(cluster_vs, (max_scale * err).astype(F), height, tan_half)
        final_lod, first_lod = parent_err > F(3e38), err < F(-0.5)
        group_visible = ~(~final_lod & (parent_px > 0) & (parent_px <= F(1.0))) & ~(~first_lod & ((cluster_px < 0) | (cluster_px > F(1.0))))
        # every meshlet of the primitive at once
        m0 = int(prim['meshletOffset'])
        meshlets = datas['meshlets']
        visible = np.ones(len(meshlets), dtype=bool)
        if material['bTwoSided'] == 0 and (flags & 4):
            cam_ls = mul_mv(mat(obj['translatedWorldToLocal']), np.array([0, 0, 0, 1], dtype=F))[:3]
            d = (meshlets['coneApex'].astype(F) - cam_ls).astype(F)
            with np.errstate(invalid='ignore', divide='ignore'):
                d = (d / np.sqrt(dot(d, d))[:, None]).astype(F)
            visible &= ~(dot(d, meshlets['coneAxis'].astype(F)) >= meshlets['coneCutOff'].astype(F))
        if flags & 1:
            pmin, pmax = meshlets['posMin'].astype(F), meshlets['posMax'].astype(F)
            center = (F(0.5) * (pmin + pmax)).astype(F)
            visible &= ~frustum_culled(planes, center, (pmax - center).astype(F), l2w)
        gi0 = int(prim['meshletGroupIndicesOffset'])
        for g, gv in zip(groups, group_visible):
            if not gv:
                continue
            for i in range(int(g['meshletCount'])):
                mi = m0 + int(datas['groupIndices'][gi0 + int(g['meshletOffset']) + i])
                if visible[mi]:
                    out.append((oi, mi, len(out)))
    return np.array(out, dtype=np.uint32).reshape(-1, 3)


def _as_triples(cmds):
    return np.stack([cmds['objectId'], cmds['meshletId'], cmds['slot']], axis=-1).astype(np.uint32)


@pytest.mark.parametrize('low', [False, True])
def test_city_lists_match_the_hlsl_statement(low):
    scene, info, frames = city_frames(frames=3, low=low)
    total = 0
    for f in frames:
        for flags in (0b111, 0b101, 0b011, 0b000):
            got = _as_triples(oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, flags).instance_culling(info['lod0MeshletCount']))
            want = cull_statement(scene, f['view'], f['iview'], f['objects'], flags)
            assert np.array_equal(got, want), (low, flags, len(got), len(want))
            total += len(got)
    assert total > 500


def test_builder_dag_lists_match_the_hlsl_statement():
    """a genuine multi-level DAG (chord_b200/nanite_builder.py): the LOD cut moves through the levels with distance."""
    scene, asset, frames = terrain_case()
    sizes = []
    for f in frames:
        got = _as_triples(oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, 0b111).instance_culling(len(scene.datas[0]['meshlets'])))
        want = cull_statement(scene, f['view'], f['iview'], f['objects'], 0b111)
        assert np.array_equal(got, want), (len(got), len(want))
        sizes.append(len(got))
    assert len(set(sizes)) > 1


# ---------------------------------------------------------------------------------------------------------------- K3
def _firstbithigh(x):
    return int(x).bit_length() - 1            # HLSL firstbithigh: -1 for 0


def hzb_cull_statement(scene, view, objects, cmds, hzb, phase0, flags=0b111):
    """hzb_mainview_culling.hlsl:35-213 for one list: returns (visible, rejected) keeping the input order."""
    view = np.asarray(view).reshape(())
    dim = np.asarray(view['renderDimension'], dtype=F)
    w2c = mat(view['translatedWorldToClipLastFrame'] if phase0 else view['translatedWorldToClip'])
    chain = hzb['min'].view(np.float16)
    mip0w, mip0h = hzb['mip0']
    vis, rej = [], []
    for cmd in cmds:
        obj = objects[cmd['objectId']]
        prim = scene.primitives[obj['GLTFPrimitiveDetail']]
        m = scene.datas[prim['primitiveDatasBufferId']]['meshlets'][cmd['meshletId']]
        visible = True
        if flags & 2:
            pmin, pmax = m['posMin'].astype(F), m['posMax'].astype(F)
            center = ((pmin + pmax) * F(0.5)).astype(F)
            extent = (pmax - center).astype(F)
            mvp = mul_mm(w2c, mat(obj['localToTranslatedWorldLastFrame'] if phase0 else obj['localToTranslatedWorld']))
            corners = fma(extent[None, :], EXTENT_FACTOR, center[None, :])
            hs = mul_mv(mvp, point4(corners))
            with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
                ndc = (hs[:, :3] / hs[:, 3:4]).astype(F)
            uvz = np.stack([fma(ndc[:, 0], F(0.5), F(0.5)), fma(ndc[:, 1], F(-0.5), F(0.5)), ndc[:, 2]], axis=-1)
            mn = np.full(3, 10.0, dtype=F); mx = np.full(3, -10.0, dtype=F)
            for k in range(8):                                # HLSL min / max: a NaN operand yields the other one
                mn = np.fmin(mn, uvz[k]); mx = np.fmax(mx, uvz[k])
            z_in_range = bool(mx[2] < 1.0 and mn[2] > 0.0)
            if z_in_range and ((mn[:2] >= 1).any() or (mx[:2] <= 0).any()):
                visible = False
            if visible and z_in_range:
                lo, hi = np.clip(mn[:2], 0, 1).astype(F), np.clip(mx[:2], 0, 1).astype(F)
                rect = np.concatenate([fma(lo, dim[:2], F(0.5)), fma(hi, dim[:2], F(-0.5))]).astype(np.int64)      # int4(): truncation
                rect[:2] = np.maximum(0, rect[:2])
                rect[2:] = np.minimum((dim[:2] - F(1.0)).astype(np.int64), rect[2:])
                if (rect[2:] < rect[:2]).any():
                    visible = False
                else:
                    c0 = rect >> 1
                    level = max(0, max(_firstbithigh(c0[2] - c0[0]), _firstbithigh(c0[3] - c0[1])) - 1)
                    if (((c0[2:] >> level) - (c0[:2] >> level)) >= 4).any():
                        level += 1
                    c = c0 >> level
                    mw = max(mip0w >> level, 1)
                    base = int(hzb['offsets'][level])
                    zmin = F(10.0)
                    for x in range(4):
                        for y in range(4):
                            sx, sy = min(c[2], c[0] + x), min(c[3], c[1] + y)
                            zmin = min(zmin, F(chain[base + int(sy) * mw + int(sx)]))
                    if zmin > mx[2]:
                        visible = False
        (vis if visible else rej).append((cmd['objectId'], cmd['meshletId'], cmd['slot']))
    return np.array(vis, dtype=np.uint32).reshape(-1, 3), np.array(rej, dtype=np.uint32).reshape(-1, 3)


@pytest.mark.parametrize('low', [False, True])
def test_two_phase_hzb_lists_match_the_hlsl_statement(low):
    """phase 0 against LAST frame's HZB with last frame's matrices, phase 1 on the rejects against the mid-frame HZB."""
    scene, info, frames = city_frames(frames=3, low=low)
    prev, occluded = None, 0
    for f in frames:
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], 640, 360, info['lod0MeshletCount'], prev)
        if prev is not None:
            v0, r0 = hzb_cull_statement(scene, f['view'], f['objects'], ref['post_cull'], prev, True)
            assert np.array_equal(v0, _as_triples(ref['stage0_visible'])) and np.array_equal(r0, _as_triples(ref['stage0_rejected']))
            v1, r1 = hzb_cull_statement(scene, f['view'], f['objects'], ref['stage0_rejected'], ref['mid_hzb'], False)
            assert np.array_equal(v1, _as_triples(ref['stage1_visible']))
            occluded += len(r1)
        prev = ref['hzb']
    assert occluded > 0 or low


# ------------------------------------------------------------------------------------------------ K6 triangle culls
def triangle_cull_statement(scene, iview, objects, cmds):
    """mesh_raster.hlsl:84-184: the four per-triangle culls of the mesh shader (#0 homogeneous back face on (x, y, w), #1 all
    w <= 0, #2 uv box outside [0, 1], #3 no pixel centre between the rounded box corners); returns (submitted, culled).
    ``determinant`` is pinned as the cofactor expansion along the first row with fma-fused minors."""
    iview = np.asarray(iview).reshape(())
    dim = np.asarray(iview['renderDimension'], dtype=F)[:2]
    w2c = mat(iview['translatedWorldToClip'])
    submitted = culled = 0
    for cmd in cmds:
        obj = objects[cmd['objectId']]
        material = scene.materials[obj['GLTFMaterialData']]
        if material['alphaMode'] > 1:                                                   # blend is never drawn (mesh_raster.cpp:224)
            continue
        if material['alphaMode'] == 1 and fma(F(1.0), F(material['baseColorFactor'][3]), -F(material['alphaCutOff'])) < 0:
            continue                                                                    # clip() on the builtin white texture
        prim = scene.primitives[obj['GLTFPrimitiveDetail']]
        datas = scene.datas[prim['primitiveDatasBufferId']]
        m = datas['meshlets'][cmd['meshletId']]
        V, T = int(m['vertexTriangleCount'] & 0xFF), int((m['vertexTriangleCount'] >> 8) & 0xFF)
        data = datas['meshletData'][int(m['dataOffset']):int(m['dataOffset']) + V + T]
        pos = datas['positions'][int(prim['vertexOffset']) + data[:V].astype(np.int64)].astype(F)
        mvp = mul_mm(w2c, mat(obj['localToTranslatedWorld']))
        hs = mul_mv(mvp, point4(pos))                                                   # (V, 4)
        tri = np.stack([data[V:] & 0xFF, (data[V:] >> 8) & 0xFF, (data[V:] >> 16) & 0xFF], axis=-1).astype(np.int64)
        p = hs[tri][..., [0, 1, 3]]                                                     # (T, 3 vertices, (x, y, w))
        c = np.zeros(T, dtype=bool)
        if material['bTwoSided'] == 0:
            def minor(a, b, cc, d):
                return fma(a, b, -(cc * d).astype(F))
            m0 = minor(p[:, 1, 1], p[:, 2, 2], p[:, 1, 2], p[:, 2, 1])
            m1 = minor(p[:, 1, 0], p[:, 2, 2], p[:, 1, 2], p[:, 2, 0])
            m2 = minor(p[:, 1, 0], p[:, 2, 1], p[:, 1, 1], p[:, 2, 0])
            det = fma(p[:, 0, 2], m2, fma(-p[:, 0, 1], m1, (p[:, 0, 0] * m0).astype(F)))
            c |= det <= 0
        c |= (p[:, :, 2] <= 0).all(axis=1)
        with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
            q = (p[:, :, :2] / np.abs(p[:, :, 2:3])).astype(F)
        uv = np.stack([fma(q[..., 0], F(0.5), F(0.5)), fma(q[..., 1], F(-0.5), F(0.5))], axis=-1)
        mx = np.fmax(uv[:, 0], np.fmax(uv[:, 1], uv[:, 2])); mn = np.fmin(uv[:, 0], np.fmin(uv[:, 1], uv[:, 2]))
        with np.errstate(invalid='ignore'):
            c |= (mn >= 1).any(axis=1) | (mx <= 0).any(axis=1)
            c |= (np.rint((mn * dim).astype(F)) == np.rint((mx * dim).astype(F))).any(axis=1)      # round(): half to even
        submitted += T
        culled += int(c.sum())
    return submitted, culled


def test_triangle_cull_counts_match_the_hlsl_statement():
    from cases import case_ragged, case_stream
    cases = [case_stream(600, 640, 360, 1.0), case_stream(300, 512, 512, 40.0), case_stream(150, 256, 256, 2000.0), case_ragged(200)]
    for c in cases:
        _, st = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], c['w'], c['h']).render_mesh(c['cmds'])
        sub, cul = triangle_cull_statement(c['scene'], c['iview'], c['objects'], c['cmds'])
        assert (sub, cul) == (st['trianglesSubmitted'], st['trianglesCulled']), (c['name'], sub, cul, st)
    for low in (False, True):
        scene, info, frames = city_frames(frames=2, low=low)
        for f in frames:
            fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360)
            post = fr.instance_culling(info['lod0MeshletCount'])
            _, st = fr.render_mesh(post)
            sub, cul = triangle_cull_statement(scene, f['iview'], f['objects'], post)
            assert (sub, cul) == (st['trianglesSubmitted'], st['trianglesCulled']), (low, sub, cul, st)
            assert cul > 0


# ------------------------------------------------------------------------------------------------------ K9 tile marker
def test_tile_marker_matches_the_hlsl_statement():
    """visibility_tile.hlsl:39-131: a 64-thread group gathers 32x32 px, every 2x2 threads OR their 4x4 px into one uint4 ->
    marker texel (tx, ty) = OR of 1 << shadingType over pixels [8 tx, 8 tx + 8) x [8 ty, 8 ty + 8), clamp-to-edge gather."""
    from chord_b200 import abi
    scene, info, frames = city_frames(frames=2)
    w, h = 640, 360
    for f in frames:
        fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], w, h)
        post = fr.instance_culling(info['lod0MeshletCount'])
        vis, _ = fr.render_mesh(post)
        got = fr.visibility_mark(vis, post)
        ids = (vis & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        _, inst = abi.decode_triangle_instance(ids)
        inst = np.where(ids != 0, inst, 0).astype(np.int64)
        types = np.where(ids != 0, scene.materials['materialType'][f['objects']['GLTFMaterialData'][post['objectId'][inst]]], abi.LIGHTING_TYPE_NONE)
        mh, mw = (h + 7) // 8, (w + 7) // 8
        ys = np.minimum(np.arange(mh * 8), h - 1); xs = np.minimum(np.arange(mw * 8), w - 1)
        t = types[np.ix_(ys, xs)].reshape(mh, 8, mw, 8).transpose(0, 2, 1, 3).reshape(mh, mw, 64).astype(np.uint64)
        want = np.zeros((mh, mw, 4), dtype=np.uint32)
        for k in range(4):
            sel = (t // 32) == k
            want[..., k] = np.bitwise_or.reduce(np.where(sel, np.uint64(1) << (t % 32), 0), axis=-1).astype(np.uint32)
        assert np.array_equal(got, want)
        assert (want[..., 0] & 2).any() and (want[..., 0] & 1).any()         # PBR tiles and empty (sky) tiles both present


# ------------------------------------------------------------------------------------- hzbCullingGeneric (shadow views)
def hzb_cull_generic_statement(scene, view, hzb_view, objects, cmds, hzb, extent_scale, use_last_frame=False, flags=0b111):
    """hzb_culling_generic.hlsl:37-172: the list against the HZB of ANOTHER view; the object matrix is re-based from the
    main camera's translated world to that view's (double subtraction of the two camera positions, then float)."""
    view = np.asarray(view).reshape(())
    hv = np.asarray(hzb_view).reshape(())
    dim = np.asarray(hv['renderDimension'], dtype=F)
    w2c = mat(hv['translatedWorldToClip'])
    rel = (np.asarray(view['cameraWorldPos'], dtype=np.uint32).view(np.float64)[:3, 0] - np.asarray(hv['cameraWorldPos'], dtype=np.uint32).view(np.float64)[:3, 0]).astype(F)
    chain = hzb['min'].view(np.float16)
    mip0w = hzb['mip0'][0]
    out = []
    for cmd in cmds:
        obj = objects[cmd['objectId']]
        prim = scene.primitives[obj['GLTFPrimitiveDetail']]
        m = scene.datas[prim['primitiveDatasBufferId']]['meshlets'][cmd['meshletId']]
        visible = True
        if flags & 2:
            pmin, pmax = m['posMin'].astype(F), m['posMax'].astype(F)
            center = ((pmin + pmax) * F(0.5)).astype(F)
            extent = (pmax - center).astype(F)
            l2w = mat(obj['localToTranslatedWorldLastFrame'] if use_last_frame else obj['localToTranslatedWorld']).copy()
            l2w[:3, 3] = (l2w[:3, 3] + rel).astype(F)
            hs = mul_mv(mul_mm(w2c, l2w), point4(fma(extent[None, :], EXTENT_FACTOR, center[None, :])))
            with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
                ndc = (hs[:, :3] / hs[:, 3:4]).astype(F)
            uvz = np.stack([fma(ndc[:, 0], F(0.5), F(0.5)), fma(ndc[:, 1], F(-0.5), F(0.5)), ndc[:, 2]], axis=-1)
            mn = np.full(3, 10.0, dtype=F); mx = np.full(3, -10.0, dtype=F)
            for k in range(8):
                mn = np.fmin(mn, uvz[k]); mx = np.fmax(mx, uvz[k])
            if bool(((uvz < 1.0) & (uvz > 0.0)).all()):
                es = F(extent_scale)
                rect = np.concatenate([fma(mn[:2], dim[:2], (es * F(-0.5))), fma(mx[:2], dim[:2], (es * F(0.5)))]).astype(np.int64)
                rect[:2] = np.maximum(0, rect[:2])
                rect[2:] = np.minimum((dim[:2] - F(1.0)).astype(np.int64), rect[2:])
                if (rect[2:] < rect[:2]).any():
                    visible = False
                else:
                    c0 = rect >> 1
                    level = max(0, max(_firstbithigh(c0[2] - c0[0]), _firstbithigh(c0[3] - c0[1])))
                    if (((c0[2:] >> level) - (c0[:2] >> level)) >= 2).any():
                        level += 1
                    c = c0 >> level
                    mw = max(mip0w >> level, 1)
                    base = int(hzb['offsets'][level])
                    zmin = F(10.0)
                    for x in range(2):
                        for y in range(2):
                            zmin = min(zmin, F(chain[base + int(min(c[3], c[1] + y)) * mw + int(min(c[2], c[0] + x))]))
                    if zmin > mx[2]:
                        visible = False
        if visible:
            out.append((cmd['objectId'], cmd['meshletId'], cmd['slot']))
    return np.array(out, dtype=np.uint32).reshape(-1, 3)


def test_generic_hzb_lists_match_the_hlsl_statement():
    from cases import case_stream
    from chord_b200 import camera as cam
    from test_oracle_shadow import STREAM_LIGHT, _stream_cfg
    c = case_stream(1200, 256, 256, 100.0)
    scene, lod0 = c['scene'], c['lod0']
    views = cam.cascade_views(c['view'], STREAM_LIGHT, _stream_cfg())
    far = oracle.Frame(scene, c['view'], views[3], c['objects'], 256, 256)
    depth_far, _ = far.render_mesh_depth(far.instance_culling(lod0))
    hzb = far.build_hzb(depth_far, True, False, False)
    near = oracle.Frame(scene, c['view'], views[2], c['objects'], 256, 256)
    post = near.instance_culling(lod0)
    for scale in (1.5, 8.0):
        got = _as_triples(near.hzb_culling_generic(hzb, views[3], scale, False, post))
        want = hzb_cull_generic_statement(scene, c['view'], views[3], c['objects'], post, hzb, scale)
        assert np.array_equal(got, want), (scale, len(got), len(want))
        assert 0 < len(got) < len(post)
