/*
 * chord_abi_types.h — byte-exact C mirrors of the GPU data model chord shares between
 * C++ and HLSL for the visibility-buffer path.  Plain C so that the C-ABI library,
 * the CPU oracle and foreign-language bindings (ctypes / cgo / JNI) all agree.
 *
 * Every struct cites the reference definition it mirrors (paths relative to the
 * chord tree).  Matrices are glm column-major: m[col*4 + row]; `mul(M, v)` in the
 * reference HLSL is the mathematical product M·v.
 *
 * No code here: only layouts, constants and the visibility-id packing.
 */
#ifndef CHORD_ABI_TYPES_H
#define CHORD_ABI_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
#define CHORD_STATIC_ASSERT(c, m) static_assert(c, m)
extern "C" {
#else
#define CHORD_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* ---- constants (install/resource/shader/base.h:396-445, nanite_shared.hlsli:8-12, gltf.h:10-12) */
#define CHORD_HZB_MAX_MIP_COUNT            12          /* kHZBMaxMipmapCount */
#define CHORD_MAX_INSTANCE_ID_COUNT        0xFFFFFFu   /* kMaxInstanceIdCount */
#define CHORD_MAX_SHADING_TYPE             0x7F
#define CHORD_LIGHTING_TYPE_NONE           0
#define CHORD_LIGHTING_TYPE_GLTF_PBR       1
#define CHORD_MESHLET_MAX_VERTICES         255
#define CHORD_MESHLET_MAX_TRIANGLES        128
#define CHORD_NANITE_MAX_LOD_COUNT         12
#define CHORD_BVH_LEVEL_NODE_COUNT         8
#define CHORD_CLUSTER_GROUP_MERGE_MAX      4
#define CHORD_SUBPIXEL_BITS                8
#define CHORD_UNVALID_ID                   0xFFFFFFFFu
#define CHORD_ERROR_PIXEL_THRESHOLD        1.0f
#define CHORD_ERROR_RADIUS_ROOT            3e38f
#define CHORD_FLAG_FRUSTUM_CULL_BIT        0           /* kFrustumCullingEnableBit  */
#define CHORD_FLAG_HZB_CULL_BIT            1           /* kHZBCullingEnableBit      */
#define CHORD_FLAG_CONE_CULL_BIT           2           /* kMeshletConeCullEnableBit */
#define CHORD_MAX_RENDER_DIM               4096        /* source/renderer/renderer.h:53 */
#define CHORD_MIN_RENDER_DIM               64

typedef struct { float x, y; }        chord_float2;
typedef struct { float x, y, z; }     chord_float3;
typedef struct { float x, y, z, w; }  chord_float4;
typedef struct { float m[16]; }       chord_float4x4;  /* column-major */
typedef struct { uint32_t x, y; }     chord_uint2;

/* gltf.h:16-24 */
typedef struct {
    chord_float4 sphere;
    uint32_t children[CHORD_BVH_LEVEL_NODE_COUNT];
    uint32_t bvhNodeCount;
    uint32_t leafMeshletGroupOffset;
    uint32_t leafMeshletGroupCount;
} chord_GPUBVHNode;
CHORD_STATIC_ASSERT(sizeof(chord_GPUBVHNode) == 60, "GPUBVHNode");

/* gltf.h:26-36 */
typedef struct {
    chord_float3 clusterPosCenter;
    float parentError;
    chord_float3 parentPosCenter;
    float error;
    uint32_t meshletOffset;
    uint32_t meshletCount;
} chord_GPUGLTFMeshletGroup;
CHORD_STATIC_ASSERT(sizeof(chord_GPUGLTFMeshletGroup) == 40, "GPUGLTFMeshletGroup");

/* gltf.h:38-51 */
typedef struct {
    chord_float3 posMin;
    uint32_t dataOffset;           /* in u32 units into meshletData */
    chord_float3 posMax;
    uint32_t vertexTriangleCount;  /* bits 0-7 vertex count, bits 8-15 triangle count */
    chord_float3 coneAxis;
    float coneCutOff;
    chord_float3 coneApex;
    uint32_t lod;
} chord_GPUGLTFMeshlet;
CHORD_STATIC_ASSERT(sizeof(chord_GPUGLTFMeshlet) == 64, "GPUGLTFMeshlet");

/* gltf.h:65-91 */
typedef struct {
    chord_float3 posMin;
    uint32_t primitiveDatasBufferId;
    chord_float3 posMax;
    uint32_t vertexOffset;
    chord_float3 posAverage;
    uint32_t vertexCount;
    uint32_t meshletOffset;
    uint32_t color0Offset;
    uint32_t smoothNormalOffset;
    uint32_t textureCoord1Offset;
    uint32_t bvhNodeOffset;
    uint32_t meshletGroupOffset;
    uint32_t meshletGroupIndicesOffset;
    uint32_t meshletGroupCount;
    uint32_t lod0IndicesOffset;
    uint32_t lod0IndicesCount;
    uint32_t pad0;
    uint32_t pad1;
} chord_GLTFPrimitiveBuffer;
CHORD_STATIC_ASSERT(sizeof(chord_GLTFPrimitiveBuffer) == 96, "GLTFPrimitiveBuffer");

/* gltf.h:94-116 — in chord these are bindless buffer ids; the C-ABI resolves them to
 * device pointers at chord_scene_upload time (see chord_primitive_datas_desc). */
typedef struct {
    uint32_t meshletBuffer, positionBuffer, normalBuffer, textureCoord0Buffer;
    uint32_t tangentBuffer, textureCoord1Buffer, color0Buffer, smoothNormalsBuffer;
    uint32_t meshletDataBuffer, bvhNodeBuffer, meshletGroupBuffer, meshletGroupIndicesBuffer;
    uint32_t lod0IndicesBuffer, pad0, pad1, pad2;
} chord_GLTFPrimitiveDatasBuffer;
CHORD_STATIC_ASSERT(sizeof(chord_GLTFPrimitiveDatasBuffer) == 64, "GLTFPrimitiveDatasBuffer");

/* gltf.h:118-153 */
typedef struct {
    uint32_t alphaMode;            /* 0 opaque, 1 masked, 2 blend */
    float alphaCutOff;
    uint32_t bTwoSided;
    uint32_t baseColorId;
    chord_float4 baseColorFactor;
    chord_float3 emissiveFactor;
    uint32_t emissiveTexture;
    float metallicFactor;
    float roughnessFactor;
    uint32_t metallicRoughnessTexture;
    uint32_t normalTexture;
    uint32_t baseColorSampler, emissiveSampler, normalSampler, metallicRoughnessSampler;
    float normalFactorScale;
    uint32_t bExistOcclusion;
    float occlusionTextureStrength;
    uint32_t materialType;
} chord_GLTFMaterialGPUData;
CHORD_STATIC_ASSERT(sizeof(chord_GLTFMaterialGPUData) == 96, "GLTFMaterialGPUData");

/* base.h:90-119 */
typedef struct { chord_uint2 x, y, z, w; } chord_GPUStorageDouble4;
CHORD_STATIC_ASSERT(sizeof(chord_GPUStorageDouble4) == 32, "GPUStorageDouble4");

/* base.h:121-135 */
typedef struct {
    chord_float4x4 translatedWorldToClip;
    chord_float4x4 clipToTranslatedWorld;
    chord_GPUStorageDouble4 cameraWorldPos;
    chord_float4 orthoDepthConvertToView;
    chord_float4 renderDimension;           /* w, h, 1/w, 1/h */
    chord_float4 frustumPlanesRS[6];
} chord_InstanceCullingViewInfo;
CHORD_STATIC_ASSERT(sizeof(chord_InstanceCullingViewInfo) == 288, "InstanceCullingViewInfo");

/* base.h:138-151 */
typedef struct {
    float width, exp_term, exp_scale, linear_term;
    float constant_term, pad0, pad1, pad2;
} chord_DensityProfileLayer;
typedef struct { chord_DensityProfileLayer layers[2]; } chord_DensityProfile;

/* base.h:156-231 */
typedef struct {
    chord_float3 solar_irradiance;   float sun_angular_radius;
    float bottom_radius;             float top_radius;  uint32_t luminanceMode; float pad1;
    chord_DensityProfile rayleigh_density;
    chord_float3 rayleigh_scattering; uint32_t bCombineScattering;
    chord_DensityProfile mie_density;
    chord_float3 mie_scattering;      float pad3;
    chord_float3 mie_extinction;      float mie_phase_function_g;
    chord_DensityProfile absorption_density;
    chord_float3 absorption_extinction; float pad4;
    chord_float3 ground_albedo;       float mu_s_min;
    chord_float3 skySpectralRadianceToLumiance; float pad5;
    chord_float3 sunSpectralRadianceToLumiance; float pad6;
    chord_float3 earthCenterKm;       float pad7;
} chord_AtmosphereParameters;
CHORD_STATIC_ASSERT(sizeof(chord_AtmosphereParameters) == 352, "AtmosphereParameters");

/* base.h:233-246 */
typedef struct { chord_float4x4 globalMatrix; } chord_GPUCascadeShadowMapConfig;
typedef struct { chord_float3 direction; float pad0; chord_GPUCascadeShadowMapConfig cascadeConfig; } chord_SkyLightInfo;
typedef struct { uint32_t STBN_scalar, STBN_vec2, STBN_vec3, pad0; } chord_GPUBlueNoiseContext;

/* base.h:256-284 */
typedef struct {
    chord_GPUBlueNoiseContext blueNoiseCtx;
    chord_AtmosphereParameters atmosphere;
    chord_SkyLightInfo sunInfo;
    uint32_t frameCounter, frameCounterMod8, GLTFPrimitiveDetailBuffer, GLTFPrimitiveDataBuffer;
    uint32_t GLTFMaterialBuffer, GLTFObjectCount, GLTFObjectBuffer, debuglineVertices;
    uint32_t debuglineCount, debuglineMaxCount, pointClampEdgeSampler, linearClampEdgeSampler;
    chord_float2 halton23_16tap[16];
    uint32_t brdfLut, pad0, pad1, pad2;
} chord_GPUBasicData;
CHORD_STATIC_ASSERT(sizeof(chord_GPUBasicData) == 640, "GPUBasicData");

/* base.h:292-340 */
typedef struct {
    chord_GPUBasicData basicData;
    chord_float4x4 translatedWorldToView;
    chord_float4x4 viewToTranslatedWorld;
    chord_float4x4 viewToClip;
    chord_float4x4 clipToView;
    chord_float4x4 translatedWorldToClip;
    chord_float4x4 translatedWorldToClip_NoJitter;
    chord_float4x4 clipToTranslatedWorld;
    chord_float4 frustumPlane[6];
    chord_float4x4 translatedWorldToClipLastFrame;
    chord_float4x4 translatedWorldToClipLastFrame_NoJitter;
    chord_float4x4 clipToTranslatedWorld_LastFrame;
    chord_float4 frustumPlaneLastFrame[6];
    chord_float4 renderDimension;                   /* w, h, 1/w, 1/h */
    chord_GPUStorageDouble4 cameraToEarthCenter_km;
    chord_GPUStorageDouble4 cameraPositionWS_km;
    chord_GPUStorageDouble4 cameraWorldPos;
    chord_GPUStorageDouble4 cameraWorldPosLastFrame;
    chord_float4x4 clipToTranslatedWorldWithZFar_NoJitter;
    float cameraFovy, zNear, zFar; uint32_t bCameraCut;
    chord_float4 jitterData;
    chord_float3 forwardRS; float pad0;
} chord_PerframeCameraView;
CHORD_STATIC_ASSERT(sizeof(chord_PerframeCameraView) == 1728, "PerframeCameraView");

/* base.h:343-360 */
typedef struct {
    chord_float4x4 localToTranslatedWorld;
    chord_float4x4 translatedWorldToLocal;
    chord_float4x4 localToTranslatedWorldLastFrame;
    chord_float4 scaleExtractFromMatrix;            /* xyz scale, w = max |scale| */
} chord_GPUObjectBasicData;
typedef struct {
    chord_GPUObjectBasicData basicData;
    uint32_t GLTFPrimitiveDetail;
    uint32_t GLTFMaterialData;
    uint32_t pad1, pad2;
} chord_GPUObjectGLTFPrimitive;
CHORD_STATIC_ASSERT(sizeof(chord_GPUObjectGLTFPrimitive) == 224, "GPUObjectGLTFPrimitive");

/* draw command currency between every stage: instance_culling.hlsl:28-33, :205 */
typedef struct { uint32_t objectId, meshletId, slot; } chord_draw_cmd;
CHORD_STATIC_ASSERT(sizeof(chord_draw_cmd) == 12, "draw cmd");

/* 32-bit visibility id: base.hlsli:437-447.  0 == no geometry. */
static inline uint32_t chord_encode_triangle_instance(uint32_t triangleId, uint32_t instanceId)
{ return (((instanceId + 1u) & CHORD_MAX_INSTANCE_ID_COUNT) << 8) | (triangleId & 0xFFu); }
static inline void chord_decode_triangle_instance(uint32_t pack, uint32_t* triangleId, uint32_t* instanceId)
{ *triangleId = pack & 0xFFu; *instanceId = ((pack >> 8) & CHORD_MAX_INSTANCE_ID_COUNT) - 1u; }

/* Packed 64-bit visibility-buffer texel used by this implementation (north_star):
 * high 32 bits = IEEE bits of the reversed-Z depth in [0,1] (render_textures.cpp:81-101
 * clears depth to 0.0, helper.h:7-13 compares GREATER_OR_EQUAL), low 32 bits = the
 * reference's 32-bit id above.  A clear texel is 0.  Larger key wins (atomicMax). */
typedef uint64_t chord_vis64;

#ifdef __cplusplus
}
#endif
#endif /* CHORD_ABI_TYPES_H */
