"""Host mirror of the reference camera / per-view uniform fill.

Follows source/utils/camera.cpp:17-154 (ICamera::fillViewUniformParameter, computeRelativeWorldFrustum),
source/utils/utils.cpp:186-198 (infiniteInvertZPerspectiveRH_ZO), application/flower/widget/viewport.cpp:435-445
(relative lookAt view, projection), source/renderer/renderer.cpp:175-263 (PerframeCameraView /
InstanceCullingViewInfo fill, last-frame hand-off) and source/scene/scene_node.cpp:42-90 (object data).
Math-convention 4x4 matrices (M @ v) in numpy float32; the dtype fields store them column-major (``M.T``).
"""
import math

import numpy as np

from . import abi

F = np.float32


def infinite_invert_z_perspective_rh_zo(aspect, fovy, z_near):
    """utils.cpp:186-198.  Returns math-convention M (clip = M @ view)."""
    t = F(math.tan(F(fovy) * F(0.5)))
    m = np.zeros((4, 4), dtype=F)
    m[0, 0] = F(1.0) / (F(aspect) * t)
    m[1, 1] = F(1.0) / t
    m[3, 2] = F(-1.0)          # glm result[2][3] = -1  -> row 3, col 2
    m[2, 3] = F(z_near)        # glm result[3][2] = zNear -> row 2, col 3
    return m


def perspective_rh_zo(fovy, aspect, z_near, z_far):
    """glm::perspectiveRH_ZO (called with near/far swapped by the reference for reversed z)."""
    t = F(math.tan(F(fovy) / F(2.0)))
    m = np.zeros((4, 4), dtype=F)
    m[0, 0] = F(1.0) / (F(aspect) * t)
    m[1, 1] = F(1.0) / t
    m[2, 2] = F(z_far) / (F(z_near) - F(z_far))
    m[3, 2] = F(-1.0)
    m[2, 3] = -(F(z_far) * F(z_near)) / (F(z_far) - F(z_near))
    return m


def _normalize(v):
    v = np.asarray(v, dtype=F)
    return v / F(np.sqrt(np.dot(v, v)))


def look_at_rh(eye, center, up):
    """glm::lookAtRH."""
    f = _normalize(np.asarray(center, dtype=F) - np.asarray(eye, dtype=F))
    s = _normalize(np.cross(f, np.asarray(up, dtype=F)))
    u = np.cross(s, f)
    m = np.eye(4, dtype=F)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m.astype(F)


class Camera:
    """ViewportCamera stand-in: double-precision world position, relative (camera at origin) view."""

    def __init__(self, position, front, up=(0.0, 1.0, 0.0), fovy=math.radians(45.0), z_near=0.001, z_far=20000.0,
                 width=1920, height=1080):
        self.position = np.asarray(position, dtype=np.float64)
        self.front = _normalize(front)
        self.right = _normalize(np.cross(self.front, np.asarray(up, dtype=F)))
        self.up = _normalize(np.cross(self.right, self.front))
        self.fovy, self.z_near, self.z_far = float(fovy), float(z_near), float(z_far)
        self.width, self.height = int(width), int(height)

    @property
    def aspect(self):
        return F(self.width) / F(self.height)

    def view_matrix(self):
        return look_at_rh(np.zeros(3, dtype=F), self.front, self.up)

    def project_matrix(self):
        return infinite_invert_z_perspective_rh_zo(self.aspect, self.fovy, self.z_near)

    def frustum_planes(self):
        """camera.cpp:80-154 computeRelativeWorldFrustum, order left, down, right, top, front, back."""
        fwd, up, right = self.front, self.up, self.right
        t = F(math.tan(F(self.fovy) * F(0.5)))
        near_c, far_c = fwd * F(self.z_near), fwd * F(self.z_far)
        yn, yf = F(self.z_near) * t, F(self.z_far) * t
        ynv, xnv = yn * up, yn * self.aspect * right
        yfv, xfv = yf * up, yf * self.aspect * right
        nrt, nrd, nlt, nld = near_c + xnv + ynv, near_c + xnv - ynv, near_c - xnv + ynv, near_c - xnv - ynv
        frt, frd, flt, fld = far_c + xfv + yfv, far_c + xfv - yfv, far_c - xfv + yfv, far_c - xfv - yfv

        def plane(pc, p1, p2):
            n = _normalize(np.cross(p1 - pc, p2 - pc))
            return np.array([n[0], n[1], n[2], -np.dot(n, pc)], dtype=F)

        return np.stack([plane(fld, flt, nld), plane(frd, fld, nrd), plane(frt, frd, nrt),
                         plane(flt, frt, nlt), plane(nrt, nrd, nlt), plane(frt, flt, frd)]).astype(F)


def fill_view(camera, jitter=(0.0, 0.0), last_view=None, frame_counter=0, sun_direction=(0.3, -0.8, 0.2)):
    """PerframeCameraView record (camera.cpp:17-78 + renderer.cpp:175-211)."""
    v = np.zeros((), dtype=abi.PerframeCameraView)
    W, H = camera.width, camera.height
    v['renderDimension'] = [W, H, F(1.0) / F(W), F(1.0) / F(H)]
    v['jitterData'][:2] = jitter
    rel_view = camera.view_matrix()
    proj = camera.project_matrix()
    v['translatedWorldToView'] = rel_view.T
    v['viewToTranslatedWorld'] = np.linalg.inv(rel_view.astype(np.float64)).astype(F).T
    v['translatedWorldToClip_NoJitter'] = (proj @ rel_view).astype(F).T
    proj_zfar = perspective_rh_zo(camera.fovy, camera.aspect, camera.z_far, camera.z_near)
    v['clipToTranslatedWorldWithZFar_NoJitter'] = np.linalg.inv((proj_zfar @ rel_view).astype(np.float64)).astype(F).T
    jm = np.eye(4, dtype=F)
    jm[0, 3] += F(2.0) * F(jitter[0]) / F(W)
    jm[1, 3] += F(-2.0) * F(jitter[1]) / F(H)
    pj = (jm @ proj).astype(F)
    v['viewToClip'] = pj.T
    v['clipToView'] = np.linalg.inv(pj.astype(np.float64)).astype(F).T
    w2c = (pj @ rel_view).astype(F)
    v['translatedWorldToClip'] = w2c.T
    v['clipToTranslatedWorld'] = np.linalg.inv(w2c.astype(np.float64)).astype(F).T
    v['frustumPlane'] = camera.frustum_planes()
    # kilometre-space camera (manager_atmosphere: earth centre below the origin)
    cam_km = camera.position / 1000.0
    earth_center_km = np.array([0.0, -6360.0, 0.0])
    v['cameraPositionWS_km'] = abi.double3_to_storage(cam_km)
    v['cameraToEarthCenter_km'] = abi.double3_to_storage(cam_km - earth_center_km)
    v['cameraFovy'], v['zNear'], v['zFar'] = camera.fovy, camera.z_near, camera.z_far
    v['cameraWorldPos'] = abi.double3_to_storage(camera.position)
    v['forwardRS'] = (np.linalg.inv(rel_view.astype(np.float64)) @ np.array([0, 0, -1, 0.0]))[:3].astype(F)
    bd = v['basicData']
    bd['frameCounter'] = frame_counter
    bd['frameCounterMod8'] = frame_counter % 8
    bd['sunInfo']['direction'] = _normalize(sun_direction)
    at = bd['atmosphere']
    at['earthCenterKm'] = earth_center_km
    # Bruneton earth defaults (manager_atmosphere.cpp:518-544 units: km); only the fields lighting.hlsl reads matter
    at['solar_irradiance'] = (1.474, 1.8504, 1.91198)
    at['sun_angular_radius'] = 0.004675
    at['bottom_radius'], at['top_radius'] = 6360.0, 6420.0
    at['luminanceMode'] = 0
    at['skySpectralRadianceToLumiance'] = (114974.9, 71305.95, 65310.55)
    at['sunSpectralRadianceToLumiance'] = (98242.79, 69954.4, 66475.0)
    at['ground_albedo'] = (0.1, 0.1, 0.1)
    at['mu_s_min'] = -0.2
    # read by the sky branch only (lighting.hlsl:75-166; Bruneton earth constants in km^-1, manager_atmosphere.cpp:518-544)
    at['rayleigh_scattering'] = (0.005802, 0.013558, 0.0331)
    at['mie_scattering'] = (0.003996, 0.003996, 0.003996)
    at['mie_phase_function_g'] = 0.8
    at['bCombineScattering'] = 1
    if last_view is None:
        last_view = v
        v['bCameraCut'] = 1
    v['translatedWorldToClipLastFrame'] = last_view['translatedWorldToClip']
    v['translatedWorldToClipLastFrame_NoJitter'] = last_view['translatedWorldToClip_NoJitter']
    v['clipToTranslatedWorld_LastFrame'] = last_view['clipToTranslatedWorld']
    v['jitterData'][2:] = last_view['jitterData'][:2]
    v['cameraWorldPosLastFrame'] = last_view['cameraWorldPos']
    v['frustumPlaneLastFrame'] = last_view['frustumPlane']
    return v


def fill_instance_view(view):
    """renderer.cpp:252-263 main-view InstanceCullingViewInfo."""
    iv = np.zeros((), dtype=abi.InstanceCullingViewInfo)
    iv['frustumPlanesRS'] = view['frustumPlane']
    iv['translatedWorldToClip'] = view['translatedWorldToClip']
    iv['clipToTranslatedWorld'] = np.linalg.inv(view['translatedWorldToClip'].T.astype(np.float64)).astype(F).T
    iv['cameraWorldPos'] = view['cameraWorldPos']
    iv['renderDimension'] = view['renderDimension']
    return iv


def fill_object(local_to_world, camera, primitive_detail, material, local_to_world_last=None, camera_last=None):
    """scene_node.cpp:42-90 getObjectBasicData + component_gltf_mesh.cpp:58-122."""
    o = np.zeros((), dtype=abi.GPUObjectGLTFPrimitive)
    m = np.array(local_to_world, dtype=np.float64)
    m[:3, 3] -= camera.position
    mf = m.astype(F)
    o['localToTranslatedWorld'] = mf.T
    o['translatedWorldToLocal'] = np.linalg.inv(mf.astype(np.float64)).astype(F).T
    s = np.array([np.linalg.norm(mf[:3, 0]), np.linalg.norm(mf[:3, 1]), np.linalg.norm(mf[:3, 2])], dtype=F)
    o['scaleExtractFromMatrix'] = [s[0], s[1], s[2], np.abs(s).max()]
    ml = np.array(local_to_world if local_to_world_last is None else local_to_world_last, dtype=np.float64)
    ml[:3, 3] -= (camera if camera_last is None else camera_last).position
    o['localToTranslatedWorldLastFrame'] = ml.astype(F).T
    o['GLTFPrimitiveDetail'], o['GLTFMaterialData'] = primitive_detail, material
    return o


# ----------------------------------------------------------------------------------------------- cascade shadow views (N3)
class CascadeShadowMapConfig:
    """render_helper.h:464-500 (only the fields the depth path reads)."""

    def __init__(self, cascadeCount=8, realtimeCascadeCount=3, cascadeDim=2048, cascadeStartDistance=0.0, cascadeEndDistance=80.0,
                 farCascadeEndDistance=800.0, splitLambda=0.8, farCascadeSplitLambda=0.8, radiusScaleFixed=10.0):
        self.cascadeCount, self.realtimeCascadeCount, self.cascadeDim = int(cascadeCount), int(realtimeCascadeCount), int(cascadeDim)
        self.cascadeStartDistance, self.cascadeEndDistance = float(cascadeStartDistance), float(cascadeEndDistance)
        self.farCascadeEndDistance, self.splitLambda = float(farCascadeEndDistance), float(splitLambda)
        self.farCascadeSplitLambda, self.radiusScaleFixed = float(farCascadeSplitLambda), float(radiusScaleFixed)


def is_cascade_cache_valid(cache_valid, realtime_count, cascade_count, tick, cascade_id):
    """cascade_setup.hlsl:8-22."""
    if (not cache_valid) or cascade_id < realtime_count:
        return False
    return (tick % (cascade_count - realtime_count)) != (cascade_id - realtime_count)


def ortho_rh_zo(left, right, bottom, top, z_near, z_far):
    """base.hlsli:668-682 ortho_RH_ZeroOne (math convention)."""
    m = np.zeros((4, 4), dtype=F)
    m[0, 0] = F(2.0) / (F(right) - F(left))
    m[1, 1] = F(2.0) / (F(top) - F(bottom))
    m[2, 2] = F(-1.0) / (F(z_far) - F(z_near))
    m[0, 3] = -(F(right) + F(left)) / (F(right) - F(left))
    m[1, 3] = -(F(top) + F(bottom)) / (F(top) - F(bottom))
    m[2, 3] = -F(z_near) / (F(z_far) - F(z_near))
    m[3, 3] = F(1.0)
    return m


def cascade_views(view, light_dir, cfg, valid_depth_range=None, previous=None, cache_valid=False):
    """cascade_setup.hlsl:89-371 cascadeComputeCS on the host: one InstanceCullingViewInfo per cascade (orthographic,
    reversed z, texel-snapped).  float32 arithmetic in the shader's order; pow() keeps this tolerance-level only, which
    is why the views are INPUTS of the parity-checked depth path (the reference computes them in a one-workgroup kernel
    so that the SDSM range never leaves the GPU).  ``valid_depth_range`` = (minBits, maxBits) of buildHZB's valid range
    or None (kUnvalidIdUint32); ``previous`` = last frame's views, kept where the cascade cache is valid."""
    near_z, far_z = F(view['zNear']), F(view['zFar'])
    clip_range = far_z - near_z
    tick = int(view['basicData']['frameCounter'])
    n_c, n_rt = cfg.cascadeCount, cfg.realtimeCascadeCount
    inv = view['clipToTranslatedWorldWithZFar_NoJitter'].T.astype(F)
    light = _normalize(light_dir)

    def log_split(far_plane, near_plane, cid, count, lam):
        rng, ratio = F(far_plane) - F(near_plane), F(far_plane) / F(near_plane)
        p = F(cid + 1) / F(count)
        log_scale = F(near_plane) * F(np.power(np.abs(ratio), p, dtype=F))
        uniform = F(near_plane) + rng * p
        d = F(lam) * (log_scale - uniform) + uniform
        return (d - near_z) / clip_range

    corners_ndc = np.array([[-1, 1, 1], [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 0], [1, 1, 0], [1, -1, 0], [-1, -1, 0]], dtype=F)
    base = np.empty((8, 3), dtype=F)
    for i in range(8):
        h = inv @ np.append(corners_ndc[i], F(1.0)).astype(F)
        base[i] = h[:3] / h[3]
    per = []
    for cid in range(n_c):
        if cid < n_rt:
            min_z, max_z = near_z + F(cfg.cascadeStartDistance), near_z + F(cfg.cascadeEndDistance)
            lam, count, sid = cfg.splitLambda, n_rt, cid
            if valid_depth_range is not None:
                vmin, vmax = (np.array(valid_depth_range, dtype=np.uint32).view(F))
                if vmax > 0:
                    min_z = max(min_z, near_z / vmax)
                if vmin > 0:
                    stable = F(cfg.cascadeEndDistance) - F(cfg.cascadeStartDistance)
                    max_z = max(max_z, min_z * F(1.1))
                    max_z = min(max_z, min_z + stable)
                    max_z = min(max_z, near_z / vmin)
                assert max_z > min_z
            split_start = min_z - near_z
        else:
            max_z, lam = near_z + F(cfg.farCascadeEndDistance), cfg.farCascadeSplitLambda
            min_z, split_start = near_z + F(cfg.cascadeEndDistance), F(cfg.cascadeEndDistance)
            count, sid = n_c - n_rt, cid - n_rt
        split = log_split(max_z, min_z, sid, count, lam)
        prev = split_start / clip_range if sid == 0 else log_split(max_z, min_z, sid - 1, count, lam)
        split0 = log_split(near_z, near_z + F(cfg.farCascadeEndDistance), 0, n_c, cfg.farCascadeSplitLambda)
        c, c0 = base.copy(), np.empty((8, 3), dtype=F)
        for i in range(4):
            ray = base[i + 4] - base[i]
            c0[i + 4], c0[i] = base[i] + ray * split0, base[i] + ray * F(0.0)
            c[i + 4], c[i] = base[i] + ray * split, base[i] + ray * prev
        center, center0 = c.sum(axis=0, dtype=F) / F(8.0), c0.sum(axis=0, dtype=F) / F(8.0)
        radius = max(F(np.sqrt(np.dot(p_ - center, p_ - center))) for p_ in c)
        radius0 = max(F(np.sqrt(np.dot(p_ - center0, p_ - center0))) for p_ in c0)
        per.append(dict(center=center, radius=F(radius), radius0=F(radius0), min_z=min_z,
                        csr=F(np.ceil(F(radius) * F(16.0))) / F(16.0)))
    max_csr = max(p_['csr'] for p_ in per)                 # WaveActiveMax over the cascades
    out = np.zeros(n_c, dtype=abi.InstanceCullingViewInfo)
    dim = F(cfg.cascadeDim)
    for cid, p_ in enumerate(per):
        if previous is not None and is_cascade_cache_valid(cache_valid, n_rt, n_c, tick, cid):
            out[cid] = previous[cid]
            continue
        ext_xy, ext_z = p_['csr'], max_csr * F(2.0)
        if cid >= n_rt:
            radius_scale, z_bias_scale = p_['radius0'] / p_['radius'], F(1.0)
        else:
            radius_scale = F(10.0) * F(cfg.radiusScaleFixed) / p_['radius']
            radius_scale = radius_scale / (radius_scale + F(1.0))
            z_bias_scale = F(0.25) + (p_['min_z'] - near_z) / (F(cfg.cascadeEndDistance) - F(cfg.cascadeStartDistance))
        radius_scale = min(radius_scale, F(1.0))
        cam_pos = (p_['center'] - light * ext_z * F(0.5)).astype(F)
        sview = look_at_rh(cam_pos, p_['center'], np.array([0, 1, 0], dtype=F))
        sproj = ortho_rh_zo(-ext_xy, ext_xy, -ext_xy, ext_xy, ext_z, F(0.0))       # near / far swapped: reversed z
        origin = ((sproj @ sview).astype(F) @ np.array([0, 0, 0, 1], dtype=F)) * (dim / F(2.0))
        rounded = np.rint(origin[:3]).astype(F)
        off = (rounded - origin[:3]) * (F(2.0) / dim)
        sproj[0, 3] += off[0]
        sproj[1, 3] += off[1]
        vp = (sproj @ sview).astype(F)
        inv_vp = np.linalg.inv(vp.astype(np.float64)).astype(F)
        p = np.empty((8, 3), dtype=F)
        for i in range(8):
            h = inv_vp @ np.append(corners_ndc[i], F(1.0)).astype(F)
            p[i] = h[:3] / h[3]

        def plane(a, b, o, d=None):
            n = _normalize(np.cross(a - o, b - o))
            return np.array([n[0], n[1], n[2], -np.dot(n if d is None else d, o)], dtype=F), n

        left, _ = plane(p[4], p[3], p[7])
        down, _ = plane(p[6], p[3], p[2])
        right, _ = plane(p[6], p[1], p[5])
        top, _ = plane(p[5], p[0], p[4])
        front, front_n = plane(p[1], p[3], p[0])
        back, _ = plane(p[5], p[7], p[6], d=front_n)       # cascade_setup.hlsl:349 uses frontN in the distance term
        iv = out[cid]
        iv['translatedWorldToClip'], iv['clipToTranslatedWorld'] = vp.T, inv_vp.T
        iv['cameraWorldPos'] = view['cameraWorldPos']
        iv['orthoDepthConvertToView'] = [sproj[2, 2], sproj[2, 3], z_bias_scale, radius_scale]
        iv['renderDimension'] = [dim, dim, F(1.0) / dim, F(1.0) / dim]
        iv['frustumPlanesRS'] = np.stack([left, down, right, top, front, back])
    return out
