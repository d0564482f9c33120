"""The whole raster (K6 + K6' + K7) stated a second time in numpy -- independently of oracle/chord_oracle.c -- and compared on
whole packed surfaces, depth bits and ids, not on single triangles.

Transcribed / specified here: the mesh shader (``mesh_raster.hlsl:51-185``: vertex transform, culls #0-#3), the fixed-function
rules the reference leaves to the GPU as DESIGN.md section 3 pins them (viewport with flipped y, snap to 1/256 px half-even,
exact integer edge functions, top-left rule, CCW front / cull back, screen-linear depth ``z0 + b1 (z1 - z0) + b2 (z2 - z0)`` with
one fma per term, reversed-Z GREATER_OR_EQUAL kept as the max of ``depthBits << 32 | id``) and the pixel shader's id
(``base.hlsli:437-441``).  Scenes without near-plane crossings (the clip path has its own exact tests in
test_oracle_raster.py); every vertex of a drawn triangle must be inside ``0 <= z <= w`` -- asserted."""
import numpy as np
import pytest

import oracle
from cases import case_ragged, case_stream
from chord_b200 import abi
from test_oracle_cull_statement import F, fma, mat, mul_mm, mul_mv, point4


GUARD_BAND = F(8.0)


def _plane_dist(v, k):
    """DESIGN.md section 3, clip path: w - z, z, then the guard band of 8 viewports: 8 w - x, 8 w + x, 8 w - y, 8 w + y."""
    if k == 0:
        return F(v[3] - v[2])
    if k == 1:
        return F(v[2])
    g = F(GUARD_BAND * v[3])
    return F(g - v[0]) if k == 2 else F(g + v[0]) if k == 3 else F(g - v[1]) if k == 4 else F(g + v[1])


def _clip_polygon(verts):
    """Sutherland-Hodgman in clip space; a new vertex is interpolated FROM the inside vertex TOWARD the outside one, so an edge
    shared by two triangles is cut identically in both."""
    poly = [np.asarray(v, dtype=F) for v in verts]
    for k in range(6):
        if len(poly) < 3:
            break
        out = []
        for i in range(len(poly)):
            a, b = poly[i], poly[(i + 1) % len(poly)]
            da, db = _plane_dist(a, k), _plane_dist(b, k)
            ina, inb = da >= 0, db >= 0
            if ina:
                out.append(a)
            if ina != inb:
                pin, pout, din, dout = (a, b, da, db) if ina else (b, a, db, da)
                with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
                    tt = F(din / F(din - dout))
                    out.append(fma(np.full(4, tt, dtype=F), (pout - pin).astype(F), pin))
        poly = out
    return poly


def _to_screen(v, half_w, half_h):
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        sx, sy, sz = fma(F(v[0] / v[3]), half_w, half_w), fma(F(v[1] / v[3]), -half_h, half_h), F(v[2] / v[3])
        return F(sx), F(sy), sz, bool(abs(sx) <= 32768 and abs(sy) <= 32768)


def _raster_screen_triangle(surface, sx, sy, sz, two_sided, vis_id):
    """snap to 1/256 px (half-even), exact integer edge functions with the top-left rule, depth plane, max of the packed key."""
    height, width = surface.shape
    x = [int(np.rint(F(v * F(256.0)))) for v in sx]
    y = [int(np.rint(F(v * F(256.0)))) for v in sy]
    area = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0])
    if area == 0 or (not two_sided and area > 0):                 # front facing <=> area < 0 (CCW front, flipped viewport)
        return 0
    sg = 1 if area > 0 else -1
    px0, px1 = max((min(x) + 127) >> 8, 0), min((max(x) - 128) >> 8, width - 1)       # pixel centres 256 p + 128 within the bbox
    py0, py1 = max((min(y) + 127) >> 8, 0), min((max(y) - 128) >> 8, height - 1)
    if px0 > px1 or py0 > py1:
        return 0
    SX = (np.arange(px0, px1 + 1, dtype=np.int64) * 256 + 128)[None, :]
    SY = (np.arange(py0, py1 + 1, dtype=np.int64) * 256 + 128)[:, None]
    cover = np.ones((py1 - py0 + 1, px1 - px0 + 1), dtype=bool)
    E = []
    for a in range(3):
        b = (a + 1) % 3
        dx, dy = sg * (x[b] - x[a]), sg * (y[b] - y[a])
        e = dx * (SY - y[a]) - dy * (SX - x[a])
        cover &= (e > 0) | ((e == 0) & ((dy == 0 and dx > 0) or dy < 0))
        E.append(e)
    if not cover.any():
        return 0
    inv_a = F(F(1.0) / F(abs(area)))
    z0, z1, z2 = F(sz[0]), F(sz[1]), F(sz[2])
    b1 = (E[2].astype(F) * inv_a).astype(F)                      # edge v2 -> v0 weighs vertex 1
    b2 = (E[0].astype(F) * inv_a).astype(F)                      # edge v0 -> v1 weighs vertex 2
    z = fma(b2, np.full(b2.shape, F(z2 - z0), dtype=F), fma(b1, np.full(b1.shape, F(z1 - z0), dtype=F), np.full(b1.shape, z0, dtype=F)))
    z = np.minimum(np.maximum(z, F(0.0)), F(1.0)).astype(F)
    key = (z.view(np.uint32).astype(np.uint64) << np.uint64(32)) | vis_id
    view = surface[py0:py1 + 1, px0:px1 + 1]
    view[cover] = np.maximum(view[cover], key[cover])
    return 1


def raster_statement(scene, iview, objects, cmds, width, height, allow_clip=False):
    iview = np.asarray(iview).reshape(())
    dim = np.asarray(iview['renderDimension'], dtype=F)[:2]
    w2c = mat(iview['translatedWorldToClip'])
    half_w, half_h = F(0.5) * F(width), F(0.5) * F(height)
    surface = np.zeros((height, width), dtype=np.uint64)
    drawn = clipped = 0
    for cmd in cmds:
        obj = objects[cmd['objectId']]
        material = scene.materials[obj['GLTFMaterialData']]
        if material['alphaMode'] > 1:
            continue
        if material['alphaMode'] == 1 and fma(F(1.0), F(material['baseColorFactor'][3]), -F(material['alphaCutOff'])) < 0:
            continue
        two_sided = material['bTwoSided'] != 0
        prim = scene.primitives[obj['GLTFPrimitiveDetail']]
        datas = scene.datas[prim['primitiveDatasBufferId']]
        m = datas['meshlets'][cmd['meshletId']]
        V, T = int(m['vertexTriangleCount'] & 0xFF), int((m['vertexTriangleCount'] >> 8) & 0xFF)
        data = datas['meshletData'][int(m['dataOffset']):int(m['dataOffset']) + V + T]
        pos = datas['positions'][int(prim['vertexOffset']) + data[:V].astype(np.int64)].astype(F)
        hs = mul_mv(mul_mm(w2c, mat(obj['localToTranslatedWorld'])), point4(pos))
        tri = np.stack([data[V:] & 0xFF, (data[V:] >> 8) & 0xFF, (data[V:] >> 16) & 0xFF], axis=-1).astype(np.int64)
        # ---- mesh-shader culls on (x, y, w)
        p = hs[tri][..., [0, 1, 3]]
        c = np.zeros(T, dtype=bool)
        if not two_sided:
            def minor(a, b, cc, d):
                return fma(a, b, -(cc * d).astype(F))
            m0 = minor(p[:, 1, 1], p[:, 2, 2], p[:, 1, 2], p[:, 2, 1])
            m1 = minor(p[:, 1, 0], p[:, 2, 2], p[:, 1, 2], p[:, 2, 0])
            m2 = minor(p[:, 1, 0], p[:, 2, 1], p[:, 1, 1], p[:, 2, 0])
            c |= fma(p[:, 0, 2], m2, fma(-p[:, 0, 1], m1, (p[:, 0, 0] * m0).astype(F))) <= 0
        c |= (p[:, :, 2] <= 0).all(axis=1)
        with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
            q = (p[:, :, :2] / np.abs(p[:, :, 2:3])).astype(F)
        uv = np.stack([fma(q[..., 0], F(0.5), F(0.5)), fma(q[..., 1], F(-0.5), F(0.5))], axis=-1)
        mx = np.fmax(uv[:, 0], np.fmax(uv[:, 1], uv[:, 2])); mn = np.fmin(uv[:, 0], np.fmin(uv[:, 1], uv[:, 2]))
        with np.errstate(invalid='ignore'):
            c |= (mn >= 1).any(axis=1) | (mx <= 0).any(axis=1)
            c |= (np.rint((mn * dim).astype(F)) == np.rint((mx * dim).astype(F))).any(axis=1)
        # ---- fixed function, per vertex
        with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
            ndc = (hs[:, :3] / hs[:, 3:4]).astype(F)
        sx, sy = fma(ndc[:, 0], half_w, half_w), fma(ndc[:, 1], -half_h, half_h)
        inside = (hs[:, 2] >= 0) & (hs[:, 2] <= hs[:, 3]) & (np.abs(sx) <= 32768) & (np.abs(sy) <= 32768)
        for t in np.nonzero(~c)[0]:
            i0, i1, i2 = tri[t]
            vis_id = np.uint64(abi.encode_triangle_instance(t, cmd['slot']))
            if inside[i0] and inside[i1] and inside[i2]:
                drawn += _raster_screen_triangle(surface, [sx[i0], sx[i1], sx[i2]], [sy[i0], sy[i1], sy[i2]], [ndc[i0, 2], ndc[i1, 2], ndc[i2, 2]],
                                                 two_sided, vis_id)
                continue
            assert allow_clip, 'this scene was declared free of near-plane / guard-band crossings'
            clipped += 1
            poly = _clip_polygon([hs[i0], hs[i1], hs[i2]])
            if len(poly) < 3:
                continue
            scr = [_to_screen(v, half_w, half_h) for v in poly]
            for i in range(1, len(poly) - 1):
                a, b, cc = scr[0], scr[i], scr[i + 1]
                if a[3] and b[3] and cc[3]:
                    drawn += _raster_screen_triangle(surface, [a[0], b[0], cc[0]], [a[1], b[1], cc[1]], [a[2], b[2], cc[2]], two_sided, vis_id)
    return surface, drawn, clipped


@pytest.mark.parametrize('case', ['stream_1', 'stream_40', 'stream_2000', 'ragged'])
def test_surface_matches_the_statement(case):
    c = {'stream_1': lambda: case_stream(500, 640, 360, 1.0), 'stream_40': lambda: case_stream(250, 512, 512, 40.0),
         'stream_2000': lambda: case_stream(120, 256, 256, 2000.0), 'ragged': lambda: case_ragged(160)}[case]()
    got, st = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], c['w'], c['h']).render_mesh(c['cmds'])
    assert st['trianglesClipped'] == 0
    want, drawn, _ = raster_statement(c['scene'], c['iview'], c['objects'], c['cmds'], c['w'], c['h'])
    assert drawn > 100
    diff = got != want
    assert not diff.any(), f'{int(diff.sum())} texels differ, first at {np.argwhere(diff)[0]}: {got[diff][0]:#x} vs {want[diff][0]:#x}'
    assert (want != 0).sum() > 1000


@pytest.mark.parametrize('low', [False, True])
def test_city_surface_matches_the_statement(low):
    """whole city frames through the statement, clip path included (the low camera skims the ground: triangles cross the near
    plane and leave the guard band; walls fill the view with large triangles)."""
    from cases import city_frames
    scene, info, frames = city_frames(frames=2, low=low)
    clipped_total = 0
    for f in frames[1:]:
        fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360)
        post = fr.instance_culling(info['lod0MeshletCount'])
        got, st = fr.render_mesh(post)
        want, drawn, clipped = raster_statement(scene, f['iview'], f['objects'], post, 640, 360, allow_clip=True)
        assert clipped == st['trianglesClipped']
        diff = got != want
        assert not diff.any(), f'{int(diff.sum())} texels differ, first at {np.argwhere(diff)[0]}: {got[diff][0]:#x} vs {want[diff][0]:#x}'
        clipped_total += clipped
    assert clipped_total > 0 or not low
