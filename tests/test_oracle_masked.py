"""SURVEY §8f N2 — masked (alpha-tested) raster bucket, CPU leg: the specified sampler against an independent numpy
statement, and invariants of the masked path of the oracle (mesh_raster.hlsl:107-112, :198-204)."""
import numpy as np

import oracle
from cases import case_masked
from chord_b200 import abi, synth

F = np.float32


def _fma(a, b, c):
    # a * b is exact in float64 (24 + 24 significant bits); one rounding to float64, then to float32
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def _wrap(x, size, mode):
    if mode == abi.ADDRESS_CLAMP_TO_EDGE:
        return min(max(x, 0), size - 1)
    if mode == abi.ADDRESS_MIRRORED_REPEAT:
        m = x % (2 * size)
        return m if m < size else 2 * size - 1 - m
    return x % size


def _bilinear(level, mode_u, mode_v, u, v):
    h, w = level.shape[:2]
    x, y = _fma(u, F(w), F(-0.5)), _fma(v, F(h), F(-0.5))
    fx0, fy0 = F(np.floor(x)), F(np.floor(y))
    fx, fy = F(x - fx0), F(y - fy0)
    x0, x1 = _wrap(int(fx0), w, mode_u), _wrap(int(fx0) + 1, w, mode_u)
    y0, y1 = _wrap(int(fy0), h, mode_v), _wrap(int(fy0) + 1, h, mode_v)
    a = lambda yy, xx: F(F(level[yy, xx, 3]) / F(255.0))
    top = _fma(fx, F(a(y0, x1) - a(y0, x0)), a(y0, x0))
    bot = _fma(fx, F(a(y1, x1) - a(y1, x0)), a(y1, x0))
    return _fma(fy, F(bot - top), top)


def numpy_sample_alpha(levels, mode_u, mode_v, u, v, dudx, dvdx, dudy, dvdy):
    """DESIGN.md section 3, 'masked bucket sampler', stated independently of oracle/chord_oracle.c."""
    h0, w0 = levels[0].shape[:2]
    ax, bx, ay, by = F(dudx * F(w0)), F(dvdx * F(h0)), F(dudy * F(w0)), F(dvdy * F(h0))
    r2 = max(_fma(bx, bx, F(ax * ax)), _fma(by, by, F(ay * ay)))
    level, frac = 0, F(0)
    if r2 > 1:
        rho = F(np.sqrt(r2))
        mant, exp = np.frexp(rho)                     # rho = mant * 2^exp, mant in [0.5, 1)
        lg = int(exp) - 1
        if lg >= len(levels) - 1:
            level, frac = len(levels) - 1, F(0)
        else:
            level, frac = lg, F(F(mant * 2) - F(1))
    a = _bilinear(levels[level], mode_u, mode_v, u, v)
    if frac > 0:
        a1 = _bilinear(levels[level + 1], mode_u, mode_v, u, v)
        a = _fma(frac, F(a1 - a), a)
    return a


def test_sampler_against_numpy_statement():
    tex = synth.alpha_test_textures()
    modes = ((abi.ADDRESS_REPEAT, abi.ADDRESS_REPEAT), (abi.ADDRESS_CLAMP_TO_EDGE, abi.ADDRESS_MIRRORED_REPEAT),
             (abi.ADDRESS_MIRRORED_REPEAT, abi.ADDRESS_CLAMP_TO_EDGE))
    table = abi.TextureTable(tex, samplers=modes)
    rng = np.random.default_rng(3)
    levels_hit = set()
    for k in range(1500):
        ti, si = int(rng.integers(1, 4)), int(rng.integers(0, 3))
        u, v = F(rng.uniform(-3, 4)), F(rng.uniform(-3, 4))
        scale = F(10.0 ** rng.uniform(-4, 0.3))
        d = (rng.normal(size=4) * scale).astype(F)
        got = oracle.sample_alpha(table, ti, si, u, v, *[float(x) for x in d])
        want = numpy_sample_alpha(tex[ti], modes[si][0], modes[si][1], u, v, *d)
        assert F(got).tobytes() == F(want).tobytes(), (k, ti, si, u, v, d, got, want)
        h0, w0 = tex[ti][0].shape[:2]
        r = np.sqrt(max((d[0] * w0) ** 2 + (d[1] * h0) ** 2, (d[2] * w0) ** 2 + (d[3] * h0) ** 2))
        levels_hit.add(0 if r <= 1 else min(int(np.floor(np.log2(r))), len(tex[ti]) - 1))
    assert levels_hit >= {0, 1, 2, 3, 4}, levels_hit
    # builtin white and out-of-range ids
    assert oracle.sample_alpha(table, 0, 0, 0.3, 0.3, 0, 0, 0, 0) == 1.0
    assert oracle.sample_alpha(table, 99, 0, 0.3, 0.3, 0, 0, 0, 0) == 1.0
    # texel centres at magnification return the texel itself; NaN / inf coordinates are defined (no crash, finite result)
    lv = tex[3][0]
    for (tx, ty) in ((0, 0), (5, 9), (15, 15)):
        a = oracle.sample_alpha(table, 3, 1, (tx + 0.5) / 16, (ty + 0.5) / 16, 0, 0, 0, 0)
        assert F(a) == F(F(lv[ty, tx, 3]) / F(255.0))
    for bad in (float('nan'), float('inf'), -float('inf'), 1e30):
        assert np.isfinite(oracle.sample_alpha(table, 1, 0, bad, 0.5, bad, 0, 0, bad))


def _render(c, cmds=None, textures='keep', tile=None):
    scene = c['scene']
    saved = scene.textures
    if textures != 'keep':
        scene.textures = textures
    try:
        fr = oracle.Frame(scene, c['view'], c['iview'], c['objects'], c['w'], c['h'], tile=tile)
        post = fr.instance_culling(c['lod0'])
        vis, st = fr.render_mesh(post if cmds is None else post[cmds])
    finally:
        scene.textures = saved
    return post, vis, st


def test_masked_bucket_invariants():
    c = case_masked()
    post, vis, st = _render(c)
    assert st['fragmentsAlphaKilled'] > 10000 and st['trianglesClipped'] > 0
    mats = c['scene'].materials
    mat_of = np.array([int(c['objects'][o]['GLTFMaterialData']) for o in post['objectId']])
    slot = ((vis & 0xFFFFFFFF) >> 8).astype(np.int64) - 1
    drawn = np.bincount(slot[slot >= 0], minlength=len(post))
    # blend materials and constant-alpha-below-cutoff materials are never drawn (mesh_raster.cpp:224; clip())
    for i, m in enumerate(mat_of):
        if mats[m]['alphaMode'] == 2 or (mats[m]['alphaMode'] == 1 and mats[m]['baseColorId'] == 0 and mats[m]['baseColorFactor'][3] < mats[m]['alphaCutOff']):
            assert drawn[i] == 0
    # without a texture table every masked material is the builtin white constant: nothing is alpha-killed, and the
    # textured image is a per-pixel subset of that one for every meshlet rendered alone
    _, vis_w, st_w = _render(c, textures=None)
    assert st_w['fragmentsAlphaKilled'] == 0
    assert st_w['trianglesCulled'] == st['trianglesCulled'] and st_w['trianglesClipped'] == st['trianglesClipped']
    for k in range(6):
        _, a, sa = _render(c, cmds=slice(k, k + 1))
        _, b, sb = _render(c, cmds=slice(k, k + 1), textures=None)
        assert not (a != 0)[b == 0].any()
        assert np.array_equal(a[a != 0], b[a != 0])                       # surviving fragments keep depth and id
        assert int((b != 0).sum()) - int((a != 0).sum()) <= sa['fragmentsAlphaKilled']
    # alpha == 1 everywhere behaves like builtin white; alpha == 0 everywhere draws nothing from textured masked materials
    ones = [None] + [abi.box_mip_chain(np.full((8, 8, 4), 255, dtype=np.uint8)) for _ in range(3)]
    zeros = [None] + [abi.box_mip_chain(np.zeros((8, 8, 4), dtype=np.uint8)) for _ in range(3)]
    _, vis_1, st_1 = _render(c, textures=abi.TextureTable(ones))
    assert np.array_equal(vis_1, vis_w) and st_1['fragmentsAlphaKilled'] == 0
    _, vis_0, _ = _render(c, textures=abi.TextureTable(zeros))
    slot0 = ((vis_0 & 0xFFFFFFFF) >> 8).astype(np.int64) - 1
    tex_masked = np.array([mats[m]['alphaMode'] == 1 and mats[m]['baseColorId'] != 0 for m in mat_of])
    assert not tex_masked[slot0[slot0 >= 0]].any()
    # screen-tile partition: union of tiles == full frame
    acc = np.zeros_like(vis)
    for tile in ((0, 0, 256, 192), (256, 0, 512, 192), (0, 192, 256, 384), (256, 192, 512, 384)):
        _, part, _ = _render(c, tile=tile)
        acc = np.maximum(acc, part)
    assert np.array_equal(acc, vis)
