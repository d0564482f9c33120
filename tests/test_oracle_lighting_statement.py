"""K11 (material resolve) stated a second time, in vectorised numpy float32, straight from the reference HLSL -- independently
of oracle/chord_oracle.c.

Transcribed here: ``lighting.hlsl:203-268`` (gltfMetallicRoughnessPBR_Lighting) and ``:270-381`` (mainCS), ``material.hlsli:30-154``
(loadGLTFMetallicRoughnessPBRMaterial, untextured branch: builtin white base colour, transparent emissive), ``nanite_shared.hlsli:
111-179`` (getTriangleMiscInfo), ``base.hlsli:457-498`` (calculateTriangleBarycentrics), ``lighting.hlsli:15-186``, ``bsdf.hlsli:7-55``,
``atmosphere.hlsli:73-77, 183-215, 269-277, 315-324, 847-856, 894-902, 1063-1088`` and ``colorspace.h:9-109``.

Pins shared with the oracle: 4x4 ``mul`` = fma chain per row, everything else plain fp32 in the
HLSL's order, ``SampleLevel`` = exact fp32 bilinear with clamp-to-edge on the RGBA16F texels, ``pow(x, 5)`` = repeated
multiplication, ``sRGB_2_AP1`` folded in double.  The two implementations must agree within the 1e-5 contract on every pixel."""
import numpy as np

import oracle
from cases import city_frames
from chord_b200 import abi, synth
from test_oracle_cull_statement import F, mat, mul_mm, mul_mv, point4


def _d(u2):
    """GPUStorageDouble4 component -> float64 (base.h:90-119: low word first)."""
    return np.asarray(u2, dtype=np.uint32).view(np.float64)[0]


def _srgb_to_ap1():
    x2a = np.array([[1.6410233797, -0.3248032942, -0.2364246952], [-0.6636628587, 1.6153315917, 0.0167563477], [0.0117218943, -0.0082844420, 0.9883948585]])
    cat = np.array([[1.0130349146, 0.0061052578, -0.0149709436], [0.0076982301, 0.9981633521, -0.0050320385], [-0.0028413174, 0.0046851567, 0.9245061375]])
    s2x = np.array([[0.4123907993, 0.3575843394, 0.1804807884], [0.2126390059, 0.7151686788, 0.0721923154], [0.0193308187, 0.1191947798, 0.9505321522]])
    return (x2a @ (cat @ s2x)).astype(F)


def _dot3(a, b):
    return ((a[..., 0] * b[..., 0]).astype(F) + (a[..., 1] * b[..., 1]).astype(F) + (a[..., 2] * b[..., 2]).astype(F)).astype(F)


def _normalize(a):
    """normalize(v) = v * rcp(sqrt(dot(v, v))): IEEE sqrt, IEEE reciprocal, three products."""
    with np.errstate(invalid='ignore', divide='ignore'):
        r = (F(1.0) / np.sqrt(_dot3(a, a)).astype(F)).astype(F)
        return (a * r[..., None]).astype(F)


def _mul33(M, v):
    return np.stack([_dot3(np.broadcast_to(M[r], v.shape), v) for r in range(3)], axis=-1)


def _bary(a, b):
    """a[0] * b.x + a[1] * b.y + a[2] * b.z for (N, 3 vertices, C) attributes."""
    return (((a[:, 0] * b[:, 0:1]).astype(F) + (a[:, 1] * b[:, 1:2]).astype(F)).astype(F) + (a[:, 2] * b[:, 2:3]).astype(F)).astype(F)


def _sample_level0(lut, u, v):
    """SampleLevel(linear clamp, 0) on RGBA16F texels, xyz only."""
    H, W = lut.shape[:2]
    x, y = (u * F(W) - F(0.5)).astype(F), (v * F(H) - F(0.5)).astype(F)
    x0f, y0f = np.floor(x), np.floor(y)
    fx, fy = (x - x0f).astype(F)[:, None], (y - y0f).astype(F)[:, None]
    x0, y0 = x0f.astype(np.int64), y0f.astype(np.int64)
    xa, xb = np.clip(x0, 0, W - 1), np.clip(x0 + 1, 0, W - 1)
    ya, yb = np.clip(y0, 0, H - 1), np.clip(y0 + 1, 0, H - 1)
    t = lut[..., :3].astype(F)
    one = F(1.0)
    return ((t[ya, xa] * ((one - fx) * (one - fy)) + t[ya, xb] * (fx * (one - fy))).astype(F)
            + (t[yb, xa] * ((one - fx) * fy) + t[yb, xb] * (fx * fy)).astype(F)).astype(F)


def _coord_from_unit(x, size):
    return (F(0.5) / F(size) + x * (F(1.0) - F(1.0) / F(size))).astype(F)


def lighting_statement(scene, view, objects, vis64, post, tr_lut, ir_lut):
    view = np.asarray(view).reshape(())
    H, W = vis64.shape
    planes = np.zeros((6, H, W, 4), dtype=F)
    ids = (vis64 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ys, xs = np.nonzero(ids)
    tri_id, inst = abi.decode_triangle_instance(ids[ys, xs])
    cmd = post[inst]
    obj_id = cmd['objectId'].astype(np.int64)
    mat_id = objects['GLTFMaterialData'][obj_id]
    keep = scene.materials['materialType'][mat_id] == abi.LIGHTING_TYPE_GLTF_PBR
    ys, xs, tri_id, obj_id, mat_id, meshlet_id = ys[keep], xs[keep], tri_id[keep].astype(np.int64), obj_id[keep], mat_id[keep], cmd['meshletId'][keep].astype(np.int64)
    n = len(ys)
    # ---- getTriangleMiscInfo, per pixel (the shader's per-thread cache only skips recomputation)
    pos = np.zeros((n, 3, 3), dtype=F); nrm = np.zeros((n, 3, 3), dtype=F)
    hs = np.zeros((n, 3, 4), dtype=F); hs_nj = np.zeros((n, 3, 3), dtype=F); hs_last = np.zeros((n, 3, 3), dtype=F); pos_rs = np.zeros((n, 3, 3), dtype=F)
    for oi in np.unique(obj_id):
        sel = np.nonzero(obj_id == oi)[0]
        obj = objects[oi]
        prim = scene.primitives[obj['GLTFPrimitiveDetail']]
        datas = scene.datas[prim['primitiveDatasBufferId']]
        m = datas['meshlets'][meshlet_id[sel]]
        V = (m['vertexTriangleCount'] & 0xFF).astype(np.int64)
        base = m['dataOffset'].astype(np.int64)
        idx = datas['meshletData'][base + V + tri_id[sel]]
        l2w = mat(obj['localToTranslatedWorld'])
        l2c = mul_mm(mat(view['translatedWorldToClip']), l2w)
        l2c_nj = mul_mm(mat(view['translatedWorldToClip_NoJitter']), l2w)
        l2c_last = mul_mm(mat(view['translatedWorldToClipLastFrame_NoJitter']), mat(obj['localToTranslatedWorldLastFrame']))
        w2l = mat(obj['translatedWorldToLocal'])
        for k in range(3):
            vi = int(prim['vertexOffset']) + datas['meshletData'][base + ((idx >> (8 * k)) & 0xFF)].astype(np.int64)
            p4 = point4(datas['positions'][vi].astype(F))
            pos_rs[sel, k] = mul_mv(l2w, p4)[:, :3]
            hs[sel, k] = mul_mv(l2c, p4)
            hs_nj[sel, k] = mul_mv(l2c_nj, p4)[:, [0, 1, 3]]
            hs_last[sel, k] = mul_mv(l2c_last, p4)[:, [0, 1, 3]]
            nl = datas['normals'][vi].astype(F)
            # mul(float4(n, 0), M): row vector times matrix -> component j = sum_i n_i M[i][j]
            nrm[sel, k] = _normalize(np.stack([_dot3(nl, np.broadcast_to(w2l[:3, j], nl.shape)) for j in range(3)], axis=-1))
    # ---- calculateTriangleBarycentrics
    texel = np.array([F(1.0) / F(W), F(1.0) / F(H)], dtype=F)
    uv = ((np.stack([xs, ys], axis=-1).astype(F) + F(0.5)) * texel).astype(F)
    pc = np.stack([F(2.0) * (uv[:, 0] - F(0.5)), F(2.0) * (F(0.5) - uv[:, 1])], axis=-1).astype(F)      # screenUvToNdcUv
    with np.errstate(divide='ignore', invalid='ignore'):
        rcp_w = (F(1.0) / hs[:, :, 3]).astype(F)
    p = (hs[:, :, :2] * rcp_w[:, :, None]).astype(F)
    p120, p201 = p[:, [1, 2, 0]], p[:, [2, 0, 1]]
    c_dx = (p201[:, :, 1] - p120[:, :, 1]).astype(F)
    c_dy = (p120[:, :, 0] - p201[:, :, 0]).astype(F)
    c = ((c_dx * (pc[:, 0:1] - p120[:, :, 0]).astype(F)).astype(F) + (c_dy * (pc[:, 1:2] - p120[:, :, 1]).astype(F)).astype(F)).astype(F)
    g = (c * rcp_w).astype(F)
    h = (((c[:, 0] * rcp_w[:, 0]).astype(F) + (c[:, 1] * rcp_w[:, 1]).astype(F)).astype(F) + (c[:, 2] * rcp_w[:, 2]).astype(F)).astype(F)
    with np.errstate(divide='ignore', invalid='ignore'):
        b = (g * (F(1.0) / h).astype(F)[:, None]).astype(F)
    # ---- loadGLTFMetallicRoughnessPBRMaterial (untextured)
    mp, ml = _bary(hs_nj, b), _bary(hs_last, b)
    with np.errstate(divide='ignore', invalid='ignore'):
        motion = ((ml[:, :2] / ml[:, 2:3]).astype(F) - (mp[:, :2] / mp[:, 2:3]).astype(F)).astype(F) * np.array([0.5, -0.5], dtype=F)
    position_rs = _bary(pos_rs, b)
    ap1 = _srgb_to_ap1()
    mats = scene.materials[mat_id]
    base_color = _mul33(ap1, (F(1.0) * mats['baseColorFactor'][:, :3]).astype(F))                    # builtin white texel (1, 1, 1, 1)
    vert_normal = _bary(nrm, b)
    pixel_normal = vert_normal
    rough = mats['roughnessFactor'].astype(F)
    metal = np.where(mats['metallicFactor'] >= 1.0, F(0.0), mats['metallicFactor']).astype(F)          # getFallbackMetallic (gltf.h:53-58)
    ao = np.ones(n, dtype=F)
    # ---- PBRMaterial::initGltfMetallicRoughnessPBR
    f0 = F(0.04)
    diffuse_color = ((base_color * (F(1.0) - f0)).astype(F) * (F(1.0) - metal)[:, None]).astype(F)
    spec_color = (f0 + ((base_color - f0).astype(F) * metal[:, None]).astype(F)).astype(F)          # lerp(f0, baseColor, metallic)
    r90 = np.clip((spec_color.max(axis=1) * F(50.0)).astype(F), 0, 1).astype(F)
    # ---- sun + sky irradiance (atmosphere.hlsli:1063-1088)
    at = view['basicData']['atmosphere']
    pix2cam = _normalize(position_rs)
    cam_km = np.array([_d(view['cameraPositionWS_km'][k]) for k in range(3)])
    pos_km = ((position_rs / F(1000.0)).astype(F).astype(np.float64) + cam_km).astype(F)
    target = (pos_km - np.asarray(at['earthCenterKm'], dtype=F)).astype(F)
    sun_dir = (-np.asarray(view['basicData']['sunInfo']['direction'], dtype=F)).astype(F)
    bottom, top = F(at['bottom_radius']), F(at['top_radius'])
    r = np.sqrt(_dot3(target, target)).astype(F)
    mu_s = (_dot3(target, np.broadcast_to(sun_dir, target.shape)) / r).astype(F)
    lum = at['luminanceMode'] != 0
    sky_l = np.asarray(at['skySpectralRadianceToLumiance'], dtype=F) if lum else np.ones(3, dtype=F)
    sun_l = np.asarray(at['sunSpectralRadianceToLumiance'], dtype=F) if lum else np.ones(3, dtype=F)
    x_r = ((r - bottom) / (top - bottom)).astype(F)
    x_mu_s = (mu_s * F(0.5) + F(0.5)).astype(F)
    irr = _sample_level0(ir_lut, _coord_from_unit(x_mu_s, 64), _coord_from_unit(x_r, 16))
    sky = ((irr * (F(1.0) + (_dot3(pixel_normal, target) / r).astype(F))[:, None]).astype(F) * F(0.5)).astype(F)
    sky = _mul33(ap1, (sky * sky_l).astype(F))
    sin_h = (bottom / r).astype(F)
    cos_h = (-np.sqrt(np.maximum(F(1.0) - (sin_h * sin_h).astype(F), F(0.0)))).astype(F)
    big_h = np.sqrt((top * top) - (bottom * bottom)).astype(F)
    rho = np.sqrt(np.maximum(((r * r).astype(F) - bottom * bottom).astype(F), F(0.0))).astype(F)
    disc = (((r * r).astype(F) * ((mu_s * mu_s).astype(F) - F(1.0)).astype(F)).astype(F) + top * top).astype(F)
    d = np.maximum(((-r * mu_s).astype(F) + np.sqrt(np.maximum(disc, F(0.0))).astype(F)).astype(F), F(0.0))
    d_min, d_max = (top - r).astype(F), (rho + big_h).astype(F)
    x_mu = ((d - d_min).astype(F) / (d_max - d_min).astype(F)).astype(F)
    transmittance = _sample_level0(tr_lut, _coord_from_unit(x_mu, 256), _coord_from_unit((rho / big_h).astype(F), 64))
    sar = F(at['sun_angular_radius'])
    e0, e1 = (-sin_h * sar).astype(F), (sin_h * sar).astype(F)
    t = np.clip((((mu_s - cos_h).astype(F) - e0).astype(F) / (e1 - e0).astype(F)).astype(F), 0, 1).astype(F)      # smoothstep
    smooth = ((t * t).astype(F) * (F(3.0) - (F(2.0) * t).astype(F)).astype(F)).astype(F)
    n_dot_sun = np.maximum(_dot3(pixel_normal, np.broadcast_to(sun_dir, pixel_normal.shape)), F(0.0))
    sun_radiance = ((np.asarray(at['solar_irradiance'], dtype=F) * (transmittance * smooth[:, None]).astype(F)).astype(F) * n_dot_sun[:, None]).astype(F)
    sun = _mul33(ap1, (sun_l * sun_radiance).astype(F))
    radiance = (sun + sky).astype(F)
    # ---- evaluateDirectionalDirectLight / getPointShade (lighting.hlsli:96-186)
    nn, vv, ll = _normalize(pixel_normal), _normalize((-pix2cam).astype(F)), _normalize(np.broadcast_to(sun_dir, pixel_normal.shape).copy())
    hh = _normalize((ll + vv).astype(F))
    n_l, n_v = np.clip(_dot3(nn, ll), 0, 1).astype(F), np.clip(_dot3(nn, vv), 0, 1).astype(F)
    n_h, v_h = np.clip(_dot3(nn, hh), 0, 1).astype(F), np.clip(_dot3(vv, hh), 0, 1).astype(F)
    om = np.clip((F(1.0) - v_h).astype(F), 0, 1).astype(F)
    om2 = (om * om).astype(F)
    p5 = ((om2 * om2).astype(F) * om).astype(F)
    fres = (spec_color + ((r90[:, None] - spec_color).astype(F) * p5[:, None]).astype(F)).astype(F)
    a2 = (rough * rough).astype(F)
    # V_SmithGGXCorrelated(NdotL, NdotV, roughness): the call site passes NdotL as the parameter named NoV (lighting.hlsli:150)
    ggx_v = (n_v * np.sqrt((((n_l * n_l).astype(F) * (F(1.0) - a2)).astype(F) + a2).astype(F))).astype(F)
    ggx_l = (n_l * np.sqrt((((n_v * n_v).astype(F) * (F(1.0) - a2)).astype(F) + a2).astype(F))).astype(F)
    ggx = (ggx_v + ggx_l).astype(F)
    with np.errstate(divide='ignore', invalid='ignore'):
        vis = np.where(ggx > 0, (F(0.5) / ggx).astype(F), F(0.0)).astype(F)
        ff = ((((n_h * a2).astype(F) - n_h).astype(F) * n_h).astype(F) + F(1.0)).astype(F)
        dd = (a2 / (((F(np.pi) * ff).astype(F) * ff).astype(F) + F(0.000001)).astype(F)).astype(F)
    lambert = (diffuse_color / F(np.pi)).astype(F)
    lit = (n_l > 0) | (n_v > 0)
    diff = np.where(lit[:, None], (n_l[:, None] * ((F(1.0) - fres).astype(F) * lambert).astype(F)).astype(F), F(0.0)).astype(F)
    spec = np.where(lit[:, None], (n_l[:, None] * ((fres * vis[:, None]).astype(F) * dd[:, None]).astype(F)).astype(F), F(0.0)).astype(F)
    diff, spec = (diff * radiance).astype(F), (spec * radiance).astype(F)
    k = F(1.0) if not lum else F(1e-5)
    diff, spec = ((F(10.0) * diff).astype(F) * k).astype(F), ((F(10.0) * spec).astype(F) * k).astype(F)      # finalRadianceExposureModify
    color = (diff + spec).astype(F)
    color[np.isnan(color).any(axis=1)] = 0
    zeros, ones = np.zeros(n, dtype=F), np.ones(n, dtype=F)
    planes[0, ys, xs] = np.concatenate([color, zeros[:, None]], axis=1)
    planes[1, ys, xs] = np.concatenate([base_color, zeros[:, None]], axis=1)
    planes[2, ys, xs] = np.concatenate([(vert_normal * F(0.5) + F(0.5)).astype(F), ones[:, None]], axis=1)
    planes[3, ys, xs] = np.concatenate([(pixel_normal * F(0.5) + F(0.5)).astype(F), ones[:, None]], axis=1)
    planes[4, ys, xs] = np.concatenate([motion, zeros[:, None], zeros[:, None]], axis=1)
    planes[5, ys, xs] = np.stack([ao, rough, metal, zeros], axis=1)
    return planes


# ---------------------------------------------------------------------------------------------- the sky branch
def _sample3d_level0(lut, u, v, w):
    """Texture3D.SampleLevel(linear clamp, 0) on a (D, H, W, 4) RGBA16F volume: weighted sum of the eight texels."""
    D, H, W = lut.shape[:3]
    t = lut.astype(F)
    out = np.zeros((len(u), 4), dtype=np.float64)
    x, y, z = (u * F(W) - F(0.5)).astype(F), (v * F(H) - F(0.5)).astype(F), (w * F(D) - F(0.5)).astype(F)
    x0, y0, z0 = np.floor(x), np.floor(y), np.floor(z)
    fx, fy, fz = (x - x0).astype(np.float64), (y - y0).astype(np.float64), (z - z0).astype(np.float64)
    x0, y0, z0 = x0.astype(np.int64), y0.astype(np.int64), z0.astype(np.int64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                wgt = (fx if dx else 1 - fx) * (fy if dy else 1 - fy) * (fz if dz else 1 - fz)
                out += t[np.clip(z0 + dz, 0, D - 1), np.clip(y0 + dy, 0, H - 1), np.clip(x0 + dx, 0, W - 1)].astype(np.float64) * wgt[:, None]
    return out.astype(F)


def _sun_and_sky_irradiance(at, ap1, point, normal, sun_dir, tr_lut, ir_lut):
    """GetSunAndSkyIrradiance (atmosphere.hlsli:1063-1088) -> (sun, sky), both through sRGB_2_AP1."""
    bottom, top = F(at['bottom_radius']), F(at['top_radius'])
    lum = at['luminanceMode'] != 0
    sky_l = np.asarray(at['skySpectralRadianceToLumiance'], dtype=F) if lum else np.ones(3, dtype=F)
    sun_l = np.asarray(at['sunSpectralRadianceToLumiance'], dtype=F) if lum else np.ones(3, dtype=F)
    r = np.sqrt(_dot3(point, point)).astype(F)
    sun_b = np.broadcast_to(sun_dir, point.shape)
    mu_s = (_dot3(point, sun_b) / r).astype(F)
    irr = _sample_level0(ir_lut, _coord_from_unit((mu_s * F(0.5) + F(0.5)).astype(F), 64), _coord_from_unit(((r - bottom) / (top - bottom)).astype(F), 16))
    sky = ((irr * (F(1.0) + (_dot3(normal, point) / r).astype(F))[:, None]).astype(F) * F(0.5)).astype(F)
    sky = _mul33(ap1, (sky * sky_l).astype(F))
    sin_h = (bottom / r).astype(F)
    cos_h = (-np.sqrt(np.maximum(F(1.0) - (sin_h * sin_h).astype(F), F(0.0)))).astype(F)
    big_h = np.sqrt((top * top) - (bottom * bottom)).astype(F)
    rho = np.sqrt(np.maximum(((r * r).astype(F) - bottom * bottom).astype(F), F(0.0))).astype(F)
    disc = (((r * r).astype(F) * ((mu_s * mu_s).astype(F) - F(1.0)).astype(F)).astype(F) + top * top).astype(F)
    d = np.maximum(((-r * mu_s).astype(F) + np.sqrt(np.maximum(disc, F(0.0))).astype(F)).astype(F), F(0.0))
    x_mu = ((d - (top - r)).astype(F) / ((rho + big_h) - (top - r)).astype(F)).astype(F)
    transmittance = _sample_level0(tr_lut, _coord_from_unit(x_mu, 256), _coord_from_unit((rho / big_h).astype(F), 64))
    sar = F(at['sun_angular_radius'])
    e0, e1 = (-sin_h * sar).astype(F), (sin_h * sar).astype(F)
    t = np.clip((((mu_s - cos_h).astype(F) - e0).astype(F) / (e1 - e0).astype(F)).astype(F), 0, 1).astype(F)
    smooth = ((t * t).astype(F) * (F(3.0) - (F(2.0) * t).astype(F)).astype(F)).astype(F)
    n_dot_sun = np.maximum(_dot3(normal, sun_b), F(0.0))
    sun_radiance = ((np.asarray(at['solar_irradiance'], dtype=F) * (transmittance * smooth[:, None]).astype(F)).astype(F) * n_dot_sun[:, None]).astype(F)
    return _mul33(ap1, (sun_l * sun_radiance).astype(F)), sky


def sky_statement(view, vis64, tr_lut, ir_lut, sc_lut, mie_lut=None):
    """skyColorEvaluate (lighting.hlsl:75-166) for every texel without geometry: GetSkyRadiance (atmosphere.hlsli:948-997),
    GetCombinedScattering (:916-946), GetScatteringTextureUvwzFromRMuMuSNu (:405-462), the ground branch through
    GetSunAndSkyIrradiance, rayleighPhase / cornetteShanksMiePhase (base.hlsli:601-615), finalRadianceExposureModify."""
    view = np.asarray(view).reshape(())
    H, W = vis64.shape
    color = np.zeros((H, W, 4), dtype=F)
    ys, xs = np.nonzero((vis64 & np.uint64(0xFFFFFFFF)) == 0)
    at = view['basicData']['atmosphere']
    ap1 = _srgb_to_ap1()
    texel = np.array([F(1.0) / F(W), F(1.0) / F(H)], dtype=F)
    uv = ((np.stack([xs, ys], axis=-1).astype(F) + F(0.5)) * texel).astype(F)
    ndc = np.stack([F(2.0) * (uv[:, 0] - F(0.5)), F(2.0) * (F(0.5) - uv[:, 1]), np.full(len(xs), 1.192092896e-07, dtype=F), np.ones(len(xs), dtype=F)], axis=-1).astype(F)
    vp = mul_mv(mat(view['clipToTranslatedWorld']), ndc)
    ray = _normalize((vp[:, :3] / vp[:, 3:4]).astype(F))
    sun_dir = (-np.asarray(view['basicData']['sunInfo']['direction'], dtype=F)).astype(F)
    cam_earth = np.array([_d(view['cameraToEarthCenter_km'][k]) for k in range(3)]).astype(F)
    cam_km = np.array([_d(view['cameraPositionWS_km'][k]) for k in range(3)]).astype(F)
    bottom, top = F(at['bottom_radius']), F(at['top_radius'])
    camera = np.broadcast_to(cam_earth, ray.shape).astype(F)
    # ---- ground
    pov, pop = _dot3(camera, ray), _dot3(camera, camera)
    with np.errstate(invalid='ignore'):
        d2i = ((-pov).astype(F) - np.sqrt(((bottom * bottom) - (pop - (pov * pov).astype(F)).astype(F)).astype(F)).astype(F)).astype(F)
    hit = d2i > 0
    ground = np.zeros_like(ray)
    if hit.any():
        ip = (cam_km + (ray[hit] * d2i[hit, None]).astype(F)).astype(F)
        rel = (ip - np.asarray(at['earthCenterKm'], dtype=F)).astype(F)
        sun_i, sky_i = _sun_and_sky_irradiance(at, ap1, rel, _normalize(rel), sun_dir, tr_lut, ir_lut)
        ground[hit] = ((np.asarray(at['ground_albedo'], dtype=F) * (F(1.0) / F(np.pi))).astype(F) * (sun_i + sky_i).astype(F)).astype(F)
    # ---- GetSkyRadiance (the camera of these tests is inside the atmosphere: the move-to-the-top branch is asserted away)
    r = np.sqrt(pop).astype(F)
    rmu = pov
    with np.errstate(invalid='ignore'):
        dtop = ((-rmu).astype(F) - np.sqrt((((rmu * rmu).astype(F) - (r * r).astype(F)).astype(F) + top * top).astype(F))).astype(F)
    assert not (dtop > 0).any() and (r <= top).all()
    mu = (rmu / r).astype(F)
    sun_b = np.broadcast_to(sun_dir, ray.shape)
    mu_s = (_dot3(camera, sun_b) / r).astype(F)
    nu = _dot3(ray, sun_b)
    hits_ground = (mu < 0) & ((((r * r).astype(F) * ((mu * mu).astype(F) - F(1.0)).astype(F)).astype(F) + bottom * bottom).astype(F) >= 0)
    big_h = np.sqrt((top * top) - (bottom * bottom)).astype(F)
    rho = np.sqrt(np.maximum(((r * r).astype(F) - bottom * bottom).astype(F), F(0.0))).astype(F)
    u_r = _coord_from_unit((rho / big_h).astype(F), 32)
    r_mu = (r * mu).astype(F)
    disc = (((r_mu * r_mu).astype(F) - (r * r).astype(F)).astype(F) + bottom * bottom).astype(F)
    with np.errstate(invalid='ignore', divide='ignore'):
        dg = ((-r_mu).astype(F) - np.sqrt(np.maximum(disc, F(0.0))).astype(F)).astype(F)
        dg_min, dg_max = (r - bottom).astype(F), rho
        xg = np.where(dg_max == dg_min, F(0.0), ((dg - dg_min).astype(F) / (dg_max - dg_min).astype(F)).astype(F)).astype(F)
        u_mu_g = (F(0.5) - (F(0.5) * _coord_from_unit(xg, 64)).astype(F)).astype(F)
        dt = ((-r_mu).astype(F) + np.sqrt(np.maximum((disc + (big_h * big_h).astype(F)).astype(F), F(0.0))).astype(F)).astype(F)
        dt_min, dt_max = (top - r).astype(F), (rho + big_h).astype(F)
        u_mu_t = (F(0.5) + (F(0.5) * _coord_from_unit(((dt - dt_min).astype(F) / (dt_max - dt_min).astype(F)).astype(F), 64)).astype(F)).astype(F)
    u_mu = np.where(hits_ground, u_mu_g, u_mu_t).astype(F)

    def dist_top(mu_):          # DistanceToTopAtmosphereBoundary(atmosphere, bottom_radius, mu)
        dsc = (((bottom * bottom) * ((mu_ * mu_).astype(F) - F(1.0)).astype(F)).astype(F) + top * top).astype(F)
        return np.maximum(((-bottom * mu_).astype(F) + np.sqrt(np.maximum(dsc, F(0.0))).astype(F)).astype(F), F(0.0)).astype(F)
    d_min, d_max = F(top - bottom), big_h
    a = ((dist_top(mu_s) - d_min).astype(F) / F(d_max - d_min)).astype(F)
    big_a = F((dist_top(np.array([at['mu_s_min']], dtype=F))[0] - d_min) / F(d_max - d_min))
    u_mu_s = _coord_from_unit((np.maximum((F(1.0) - (a / big_a).astype(F)).astype(F), F(0.0)) / (F(1.0) + a).astype(F)).astype(F), 32)
    u_nu = ((nu + F(1.0)).astype(F) / F(2.0)).astype(F)
    tcx = (u_nu * F(7.0)).astype(F)
    tex_x = np.floor(tcx).astype(F)
    lerp = (tcx - tex_x).astype(F)
    ua, ub = ((tex_x + u_mu_s).astype(F) / F(8.0)).astype(F), (((tex_x + F(1.0)).astype(F) + u_mu_s).astype(F) / F(8.0)).astype(F)
    s0, s1 = _sample3d_level0(sc_lut, ua, u_mu, u_r), _sample3d_level0(sc_lut, ub, u_mu, u_r)
    w0 = (F(1.0) - lerp).astype(F)[:, None]
    comb = ((s0 * w0).astype(F) + (s1 * lerp[:, None]).astype(F)).astype(F)
    scattering = comb[:, :3]
    if at['bCombineScattering'] != 0:
        ray_s, mie_s = np.asarray(at['rayleigh_scattering'], dtype=F), np.asarray(at['mie_scattering'], dtype=F)
        with np.errstate(invalid='ignore', divide='ignore'):
            ext = ((((scattering * comb[:, 3:4]).astype(F) / comb[:, 0:1]).astype(F) * F(ray_s[0] / mie_s[0])).astype(F) * (mie_s / ray_s).astype(F)).astype(F)
        mie = np.where(comb[:, 0:1] <= 0, F(0.0), ext).astype(F)
    else:
        ml = mie_lut if mie_lut is not None else sc_lut
        m0, m1 = _sample3d_level0(ml, ua, u_mu, u_r), _sample3d_level0(ml, ub, u_mu, u_r)
        mie = ((m0[:, :3] * w0).astype(F) + (m1[:, :3] * lerp[:, None]).astype(F)).astype(F)
    one_nu2 = (F(1.0) + (nu * nu).astype(F)).astype(F)
    rayleigh_phase = (F(F(3.0) / (F(16.0) * F(np.pi))) * one_nu2).astype(F)
    g = F(at['mie_phase_function_g'])
    g2 = F(g * g)
    kk = F(F(F(3.0) / F(F(8.0) * F(np.pi))) * F(F(1.0) - g2)) / F(F(2.0) + g2)
    base = ((F(1.0) + g2) - (F(F(2.0) * g) * nu).astype(F)).astype(F)
    mie_phase = ((F(kk) * one_nu2).astype(F) / (base * np.sqrt(base).astype(F)).astype(F)).astype(F)          # pow(base, 1.5) pinned as base * sqrt(base)
    lum = at['luminanceMode'] != 0
    sky_l = np.asarray(at['skySpectralRadianceToLumiance'], dtype=F) if lum else np.ones(3, dtype=F)
    rad = _mul33(ap1, (((scattering * rayleigh_phase[:, None]).astype(F) + (mie * mie_phase[:, None]).astype(F)).astype(F) * sky_l).astype(F))
    alpha = hit.astype(F)[:, None]
    rad = (rad + ((ground - rad).astype(F) * alpha).astype(F)).astype(F)                              # lerp(radiance, groundRadiance, groundAlpha)
    k = F(1.0) if not lum else F(1e-5)
    out = ((F(10.0) * rad).astype(F) * k).astype(F)
    out[np.isnan(out).any(axis=1)] = 0
    color[ys, xs, :3] = out
    return color


def test_sky_matches_the_hlsl_statement():
    """The oracle's sky branch against the numpy statement, on the single-meshlet frame (half of the empty texels look above the
    horizon, half hit the planet) and between the buildings of the city, combined and separate single-Mie volumes, radiance and
    luminance mode."""
    from chord_b200 import camera as cam
    tr, ir = synth.synthetic_luts()
    sc, mie = synth.synthetic_scattering_luts()
    scene, camera, objects = synth.single_meshlet_scene()
    for combine, m, lum in ((1, None, 0), (0, mie, 0), (1, None, 1)):
        view = cam.fill_view(camera)
        view['basicData']['atmosphere']['bCombineScattering'] = combine
        view['basicData']['atmosphere']['luminanceMode'] = lum
        iview = cam.fill_instance_view(view)
        fr = oracle.Frame(scene, view, iview, objects, 256, 256)
        vis, _ = fr.render_mesh(fr.instance_culling(4))
        got = oracle.sky(fr, vis, tr, ir, sc, m)
        want = sky_statement(view, vis, tr, ir, sc, m)
        err = np.abs(got - want)
        assert (err <= 1e-5 + 1e-5 * np.abs(want)).all(), (combine, lum, err.max(), np.unravel_index(err.argmax(), err.shape))
        assert np.array_equal(got.any(axis=-1), want.any(axis=-1)) and (want[..., :3].sum(axis=-1) > 0).sum() > 10000
    scene, info, frames = city_frames(frames=1)
    f = frames[0]
    ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], 640, 360, info['lod0MeshletCount'], None)
    got = oracle.sky(ref['frame'], ref['vis64'], tr, ir, sc)
    want = sky_statement(f['view'], ref['vis64'], tr, ir, sc)
    err = np.abs(got - want)
    assert (err <= 1e-5 + 1e-5 * np.abs(want)).all(), err.max()
    assert (want[..., :3].sum(axis=-1) > 0).sum() > 5000


def test_resolve_matches_the_hlsl_statement():
    scene, info, frames = city_frames(frames=2)
    tr, ir = synth.synthetic_luts()
    prev = None
    names = ['color', 'baseColor', 'vertexNormal', 'pixelNormal', 'motion', 'aoRoughMetal']
    for f in frames:
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], 640, 360, info['lod0MeshletCount'], prev)
        got = oracle.lighting(ref['frame'], ref['vis64'], ref['post_cull'], tr, ir)
        want = lighting_statement(scene, f['view'], f['objects'], ref['vis64'], ref['post_cull'], tr, ir)
        for i, name in enumerate(names):
            err = np.abs(got[i] - want[i])
            tol = 1e-5 + 1e-5 * np.abs(want[i])
            assert (err <= tol).all(), f'{name}: max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}'
            assert np.array_equal(got[i].any(axis=-1), want[i].any(axis=-1)), f'{name}: coverage differs'
        assert want[0].max() > 0.05 and (want[0][..., :3].sum(axis=-1) > 0).mean() > 0.2
        prev = ref['hzb']
