"""numpy / ctypes mirrors of include/chord_abi_types.h and include/chord_b200.h.

The numpy dtypes are byte-for-byte the reference structs (install/resource/shader/gltf.h:16-153,
base.h:90-360); sizes are asserted at import.  Matrices are stored glm column-major: the dtype
field has shape (4, 4) indexed [col][row], so assign ``rec['m'] = M.T`` for a math-convention M.
"""
import ctypes as C

import numpy as np

f4, u4 = np.float32, np.uint32

MESHLET_MAX_VERTICES = 255
MESHLET_MAX_TRIANGLES = 128
CLUSTER_GROUP_MERGE_MAX = 4
HZB_MAX_MIP_COUNT = 12
UNVALID_ID = 0xFFFFFFFF
FLAG_FRUSTUM, FLAG_HZB, FLAG_CONE = 1 << 0, 1 << 1, 1 << 2
LIGHTING_TYPE_NONE, LIGHTING_TYPE_GLTF_PBR = 0, 1

mat4 = (f4, (4, 4))
vec3 = (f4, (3,))
vec4 = (f4, (4,))
double4 = (u4, (4, 2))

GPUBVHNode = np.dtype([
    ('sphere', *vec4), ('children', u4, (8,)),
    ('bvhNodeCount', u4), ('leafMeshletGroupOffset', u4), ('leafMeshletGroupCount', u4)])
GPUGLTFMeshletGroup = np.dtype([
    ('clusterPosCenter', *vec3), ('parentError', f4), ('parentPosCenter', *vec3), ('error', f4),
    ('meshletOffset', u4), ('meshletCount', u4)])
GPUGLTFMeshlet = np.dtype([
    ('posMin', *vec3), ('dataOffset', u4), ('posMax', *vec3), ('vertexTriangleCount', u4),
    ('coneAxis', *vec3), ('coneCutOff', f4), ('coneApex', *vec3), ('lod', u4)])
GLTFPrimitiveBuffer = np.dtype([
    ('posMin', *vec3), ('primitiveDatasBufferId', u4), ('posMax', *vec3), ('vertexOffset', u4),
    ('posAverage', *vec3), ('vertexCount', u4),
    ('meshletOffset', u4), ('color0Offset', u4), ('smoothNormalOffset', u4), ('textureCoord1Offset', u4),
    ('bvhNodeOffset', u4), ('meshletGroupOffset', u4), ('meshletGroupIndicesOffset', u4), ('meshletGroupCount', u4),
    ('lod0IndicesOffset', u4), ('lod0IndicesCount', u4), ('pad0', u4), ('pad1', u4)])
GLTFMaterialGPUData = np.dtype([
    ('alphaMode', u4), ('alphaCutOff', f4), ('bTwoSided', u4), ('baseColorId', u4),
    ('baseColorFactor', *vec4), ('emissiveFactor', *vec3), ('emissiveTexture', u4),
    ('metallicFactor', f4), ('roughnessFactor', f4), ('metallicRoughnessTexture', u4), ('normalTexture', u4),
    ('baseColorSampler', u4), ('emissiveSampler', u4), ('normalSampler', u4), ('metallicRoughnessSampler', u4),
    ('normalFactorScale', f4), ('bExistOcclusion', u4), ('occlusionTextureStrength', f4), ('materialType', u4)])
InstanceCullingViewInfo = np.dtype([
    ('translatedWorldToClip', *mat4), ('clipToTranslatedWorld', *mat4), ('cameraWorldPos', *double4),
    ('orthoDepthConvertToView', *vec4), ('renderDimension', *vec4), ('frustumPlanesRS', f4, (6, 4))])
DensityProfileLayer = np.dtype([
    ('width', f4), ('exp_term', f4), ('exp_scale', f4), ('linear_term', f4),
    ('constant_term', f4), ('pad0', f4), ('pad1', f4), ('pad2', f4)])
DensityProfile = np.dtype([('layers', DensityProfileLayer, (2,))])
AtmosphereParameters = np.dtype([
    ('solar_irradiance', *vec3), ('sun_angular_radius', f4),
    ('bottom_radius', f4), ('top_radius', f4), ('luminanceMode', u4), ('pad1', f4),
    ('rayleigh_density', DensityProfile), ('rayleigh_scattering', *vec3), ('bCombineScattering', u4),
    ('mie_density', DensityProfile), ('mie_scattering', *vec3), ('pad3', f4),
    ('mie_extinction', *vec3), ('mie_phase_function_g', f4),
    ('absorption_density', DensityProfile), ('absorption_extinction', *vec3), ('pad4', f4),
    ('ground_albedo', *vec3), ('mu_s_min', f4),
    ('skySpectralRadianceToLumiance', *vec3), ('pad5', f4),
    ('sunSpectralRadianceToLumiance', *vec3), ('pad6', f4),
    ('earthCenterKm', *vec3), ('pad7', f4)])
SkyLightInfo = np.dtype([('direction', *vec3), ('pad0', f4), ('cascadeGlobalMatrix', *mat4)])
GPUBasicData = np.dtype([
    ('blueNoiseCtx', u4, (4,)), ('atmosphere', AtmosphereParameters), ('sunInfo', SkyLightInfo),
    ('frameCounter', u4), ('frameCounterMod8', u4), ('GLTFPrimitiveDetailBuffer', u4), ('GLTFPrimitiveDataBuffer', u4),
    ('GLTFMaterialBuffer', u4), ('GLTFObjectCount', u4), ('GLTFObjectBuffer', u4), ('debuglineVertices', u4),
    ('debuglineCount', u4), ('debuglineMaxCount', u4), ('pointClampEdgeSampler', u4), ('linearClampEdgeSampler', u4),
    ('halton23_16tap', f4, (16, 2)), ('brdfLut', u4), ('pad0', u4), ('pad1', u4), ('pad2', u4)])
PerframeCameraView = np.dtype([
    ('basicData', GPUBasicData),
    ('translatedWorldToView', *mat4), ('viewToTranslatedWorld', *mat4),
    ('viewToClip', *mat4), ('clipToView', *mat4),
    ('translatedWorldToClip', *mat4), ('translatedWorldToClip_NoJitter', *mat4), ('clipToTranslatedWorld', *mat4),
    ('frustumPlane', f4, (6, 4)),
    ('translatedWorldToClipLastFrame', *mat4), ('translatedWorldToClipLastFrame_NoJitter', *mat4),
    ('clipToTranslatedWorld_LastFrame', *mat4),
    ('frustumPlaneLastFrame', f4, (6, 4)),
    ('renderDimension', *vec4),
    ('cameraToEarthCenter_km', *double4), ('cameraPositionWS_km', *double4),
    ('cameraWorldPos', *double4), ('cameraWorldPosLastFrame', *double4),
    ('clipToTranslatedWorldWithZFar_NoJitter', *mat4),
    ('cameraFovy', f4), ('zNear', f4), ('zFar', f4), ('bCameraCut', u4),
    ('jitterData', *vec4), ('forwardRS', *vec3), ('pad0', f4)])
GPUObjectGLTFPrimitive = np.dtype([
    ('localToTranslatedWorld', *mat4), ('translatedWorldToLocal', *mat4),
    ('localToTranslatedWorldLastFrame', *mat4), ('scaleExtractFromMatrix', *vec4),
    ('GLTFPrimitiveDetail', u4), ('GLTFMaterialData', u4), ('pad1', u4), ('pad2', u4)])
DrawCmd = np.dtype([('objectId', u4), ('meshletId', u4), ('slot', u4)])

_SIZES = {
    'GPUBVHNode': (GPUBVHNode, 60), 'GPUGLTFMeshletGroup': (GPUGLTFMeshletGroup, 40),
    'GPUGLTFMeshlet': (GPUGLTFMeshlet, 64), 'GLTFPrimitiveBuffer': (GLTFPrimitiveBuffer, 96),
    'GLTFMaterialGPUData': (GLTFMaterialGPUData, 96), 'InstanceCullingViewInfo': (InstanceCullingViewInfo, 288),
    'AtmosphereParameters': (AtmosphereParameters, 352), 'GPUBasicData': (GPUBasicData, 640),
    'PerframeCameraView': (PerframeCameraView, 1728), 'GPUObjectGLTFPrimitive': (GPUObjectGLTFPrimitive, 224),
    'DrawCmd': (DrawCmd, 12),
}
for _n, (_d, _s) in _SIZES.items():
    assert _d.itemsize == _s, f'{_n}: {_d.itemsize} != {_s}'


def encode_triangle_instance(triangle_id, instance_id):
    """base.hlsli:437-441 encodeTriangleIdInstanceId."""
    return ((((np.uint32(instance_id) + np.uint32(1)) & np.uint32(0xFFFFFF)) << np.uint32(8))
            | (np.uint32(triangle_id) & np.uint32(0xFF)))


def decode_triangle_instance(pack):
    """base.hlsli:443-447 decodeTriangleIdInstanceId -> (triangleId, instanceId)."""
    pack = np.asarray(pack, dtype=np.uint32)
    return pack & np.uint32(0xFF), ((pack >> np.uint32(8)) & np.uint32(0xFFFFFF)) - np.uint32(1)


def pack_vertex_triangle_count(v, t):
    """gltf.h:60 packVertexCountTriangleCount."""
    return (np.uint32(v) & np.uint32(0xFF)) | ((np.uint32(t) & np.uint32(0xFF)) << np.uint32(8))


def double3_to_storage(v):
    """base.h:90-119 GPUStorageDouble4 (low word first)."""
    out = np.zeros((4, 2), dtype=np.uint32)
    out[:3] = np.asarray(v, dtype=np.float64).view(np.uint32).reshape(3, 2)
    return out


# ---------------------------------------------------------------- ctypes descs (chord_b200.h)
class PrimitiveDatasDesc(C.Structure):
    _fields_ = [
        ('meshlets', C.c_void_p), ('meshletCount', C.c_uint32),
        ('positions', C.c_void_p), ('vertexCount', C.c_uint32),
        ('normals', C.c_void_p), ('uv0', C.c_void_p), ('tangents', C.c_void_p),
        ('meshletData', C.c_void_p), ('meshletDataCount', C.c_uint32),
        ('bvhNodes', C.c_void_p), ('bvhNodeCount', C.c_uint32),
        ('groups', C.c_void_p), ('groupCount', C.c_uint32),
        ('groupIndices', C.c_void_p), ('groupIndicesCount', C.c_uint32),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ('datas', C.POINTER(PrimitiveDatasDesc)), ('datasCount', C.c_uint32),
        ('primitives', C.c_void_p), ('primitiveCount', C.c_uint32),
        ('materials', C.c_void_p), ('materialCount', C.c_uint32),
    ]


class Rect(C.Structure):
    _fields_ = [('x0', C.c_uint32), ('y0', C.c_uint32), ('x1', C.c_uint32), ('y1', C.c_uint32)]


class RasterStats(C.Structure):
    _fields_ = [('trianglesSubmitted', C.c_uint64), ('trianglesCulled', C.c_uint64),
                ('trianglesClipped', C.c_uint64), ('trianglesLarge', C.c_uint64),
                ('largeQueueOverflow', C.c_uint64), ('meshletsRastered', C.c_uint64),
                ('fragmentsAlphaKilled', C.c_uint64), ('pad0', C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


ADDRESS_REPEAT, ADDRESS_CLAMP_TO_EDGE, ADDRESS_MIRRORED_REPEAT = 0, 1, 2
PASS_DEPTH, PASS_DEPTH_CLAMP = 1, 2


TEXTURE_SRGB = 1


class TextureDesc(C.Structure):
    _fields_ = [('rgba8', C.c_void_p), ('width', C.c_uint32), ('height', C.c_uint32), ('mipCount', C.c_uint32), ('flags', C.c_uint32)]


class SamplerDesc(C.Structure):
    _fields_ = [('addressU', C.c_uint32), ('addressV', C.c_uint32)]


def box_mip_chain(level0):
    """RGBA8 mip chain of an (h, w, 4) uint8 image by 2x2 box filtering (round half up), down to 1x1."""
    levels = [np.ascontiguousarray(level0, dtype=np.uint8)]
    while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:
        a = levels[-1].astype(np.uint32)
        h, w = a.shape[:2]
        nh, nw = max(h // 2, 1), max(w // 2, 1)
        ys = np.minimum(np.arange(nh)[:, None] * 2 + np.array([0, 1])[None], h - 1)
        xs = np.minimum(np.arange(nw)[:, None] * 2 + np.array([0, 1])[None], w - 1)
        s = a[ys[:, 0]][:, xs[:, 0]] + a[ys[:, 0]][:, xs[:, 1]] + a[ys[:, 1]][:, xs[:, 0]] + a[ys[:, 1]][:, xs[:, 1]]
        levels.append(((s + 2) // 4).astype(np.uint8))
    return levels


class TextureTable:
    """The material texture table chord_textures_upload takes: ``textures`` is a list whose entries are None (builtin
    white) or a list of (h, w, 4) uint8 mip levels (level k is max(1, w >> k) x max(1, h >> k)); ``samplers`` is a list
    of (addressU, addressV)."""

    def __init__(self, textures, samplers=((ADDRESS_REPEAT, ADDRESS_REPEAT),), srgb=()):
        """``srgb``: indices of the textures whose r, g, b are sRGB encoded (CHORD_TEXTURE_SRGB: base colour / emissive images)."""
        self._keep = []
        self.tex = (TextureDesc * max(len(textures), 1))()
        self.n_tex = len(textures)
        for i, t in enumerate(textures):
            if t is None:
                continue
            h, w = t[0].shape[:2]
            for k, lv in enumerate(t):
                assert lv.shape == (max(h >> k, 1), max(w >> k, 1), 4) and lv.dtype == np.uint8, (k, lv.shape)
            blob = np.concatenate([np.ascontiguousarray(lv).reshape(-1) for lv in t])
            self._keep.append(blob)
            self.tex[i].rgba8, self.tex[i].width, self.tex[i].height, self.tex[i].mipCount = blob.ctypes.data, w, h, len(t)
            self.tex[i].flags = TEXTURE_SRGB if i in srgb else 0
        self.smp = (SamplerDesc * max(len(samplers), 1))()
        self.n_smp = len(samplers)
        for i, (u, v) in enumerate(samplers):
            self.smp[i].addressU, self.smp[i].addressV = u, v


class CascadeSetupDesc(C.Structure):
    """chord_cascade_setup_desc = CascadeSetupPushConsts (cascade_setup.hlsl:24-50)."""
    _fields_ = [('lightDir', C.c_float * 3), ('cascadeStartDistance', C.c_float), ('cascadeEndDistance', C.c_float),
                ('splitLambda', C.c_float), ('farCascadeSplitLambda', C.c_float), ('farCascadeEndDistance', C.c_float),
                ('radiusScale', C.c_float), ('cascadeCount', C.c_uint32), ('realtimeCount', C.c_uint32), ('cascadeDim', C.c_uint32),
                ('bCacheValid', C.c_uint32), ('tickCount', C.c_uint32)]


class AtmosphereLuts(C.Structure):
    _fields_ = [('transmittance', C.c_void_p), ('irradiance', C.c_void_p)]


def _ptr(a):
    return None if a is None else a.ctypes.data


class HostScene:
    """Host-side scene arrays in the reference layout + the ctypes desc that points at them.

    ``datas`` is a list of dicts with numpy arrays: meshlets (GPUGLTFMeshlet), positions (N,3) f32,
    normals (N,3) / uv0 (N,2) / tangents (N,4) or None, meshletData u32, bvhNodes or None,
    groups (GPUGLTFMeshletGroup), groupIndices u32.
    """

    def __init__(self, datas, primitives, materials, textures=None):
        self.datas = datas
        self.textures = textures          # optional TextureTable (masked bucket)
        self.primitives = np.ascontiguousarray(primitives, dtype=GLTFPrimitiveBuffer)
        self.materials = np.ascontiguousarray(materials, dtype=GLTFMaterialGPUData)
        for d in self.datas:
            d['meshlets'] = np.ascontiguousarray(d['meshlets'], dtype=GPUGLTFMeshlet)
            d['positions'] = np.ascontiguousarray(d['positions'], dtype=np.float32).reshape(-1, 3)
            d['meshletData'] = np.ascontiguousarray(d['meshletData'], dtype=np.uint32)
            d['groups'] = np.ascontiguousarray(d['groups'], dtype=GPUGLTFMeshletGroup)
            d['groupIndices'] = np.ascontiguousarray(d['groupIndices'], dtype=np.uint32)
            for k, w in (('normals', 3), ('uv0', 2), ('tangents', 4)):
                if d.get(k) is not None:
                    d[k] = np.ascontiguousarray(d[k], dtype=np.float32).reshape(-1, w)
                else:
                    d[k] = None
            if d.get('bvhNodes') is not None:
                d['bvhNodes'] = np.ascontiguousarray(d['bvhNodes'], dtype=GPUBVHNode)
            else:
                d['bvhNodes'] = None
        self._descs = (PrimitiveDatasDesc * len(self.datas))()
        for i, d in enumerate(self.datas):
            x = self._descs[i]
            x.meshlets, x.meshletCount = _ptr(d['meshlets']), len(d['meshlets'])
            x.positions, x.vertexCount = _ptr(d['positions']), len(d['positions'])
            x.normals, x.uv0, x.tangents = _ptr(d['normals']), _ptr(d['uv0']), _ptr(d['tangents'])
            x.meshletData, x.meshletDataCount = _ptr(d['meshletData']), len(d['meshletData'])
            x.bvhNodes = _ptr(d['bvhNodes'])
            x.bvhNodeCount = 0 if d['bvhNodes'] is None else len(d['bvhNodes'])
            x.groups, x.groupCount = _ptr(d['groups']), len(d['groups'])
            x.groupIndices, x.groupIndicesCount = _ptr(d['groupIndices']), len(d['groupIndices'])
        self.desc = SceneDesc()
        self.desc.datas = C.cast(self._descs, C.POINTER(PrimitiveDatasDesc))
        self.desc.datasCount = len(self.datas)
        self.desc.primitives, self.desc.primitiveCount = _ptr(self.primitives), len(self.primitives)
        self.desc.materials, self.desc.materialCount = _ptr(self.materials), len(self.materials)

    @property
    def total_meshlets(self):
        return sum(len(d['meshlets']) for d in self.datas)

    @property
    def total_triangles(self):
        return int(sum(((d['meshlets']['vertexTriangleCount'] >> 8) & 0xFF).sum() for d in self.datas))
