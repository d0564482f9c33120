// shadow.cu — cascade shadow view setup on the device: cascade_setup.hlsl:89-371 (CascadeSetupCS, one
// thread per cascade) so that the SDSM depth range (buildHZB's valid range) never leaves the GPU and the shadow pass
// stays asynchronous.  Output: InstanceCullingViewInfo[cascadeCount] in HBM, consumed by chord_frame_begin_cascade and
// chord_hzb_culling_generic_cascade.  pow() makes this tolerance-level against any other implementation (the host
// mirror chord_b200/camera.py::cascade_views included); everything downstream is bit-exact GIVEN these views.
#include "common.cuh"

#include <cstring>

namespace chordb {

struct CascadeSetupParams {
    chord_cascade_setup_desc d;
    const chord_PerframeCameraView* view;        // device
    const uint32_t* validRange;                  // device {minBits, maxBits} or null (kUnvalidIdUint32)
    chord_InstanceCullingViewInfo* out;          // device
};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 normalize(V3 a) { const float l = sqrtf(dot(a, a)); return v3(a.x / l, a.y / l, a.z / l); }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }

// 4x4 matrices in math convention m[row][col]
struct M4 { float m[4][4]; };
__device__ M4 mul(const M4& a, const M4& b)
{
    M4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0.0f; for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
    return r;
}
__device__ void mul_point(const M4& a, float x, float y, float z, float w, float o[4])
{ for (int i = 0; i < 4; i++) o[i] = a.m[i][0] * x + a.m[i][1] * y + a.m[i][2] * z + a.m[i][3] * w; }

__device__ float det3(float a, float b, float c, float d, float e, float f, float g, float h, float i)
{ return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g); }

// adjugate / determinant (base.hlsli:684-724 matrixInverse computes the same cofactors)
__device__ M4 inverse(const M4& a)
{
    M4 r;
    float cof[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s[9]; int n = 0;
            for (int y = 0; y < 4; y++) if (y != i) for (int x = 0; x < 4; x++) if (x != j) s[n++] = a.m[y][x];
            const float d = det3(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
            cof[i][j] = ((i + j) & 1) ? -d : d;
        }
    const float det = a.m[0][0] * cof[0][0] + a.m[0][1] * cof[0][1] + a.m[0][2] * cof[0][2] + a.m[0][3] * cof[0][3];
    const float idet = 1.0f / det;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = cof[j][i] * idet;
    return r;
}

__device__ M4 load_mat(const chord_float4x4& c)        // column-major storage -> math convention
{ M4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = c.m[j * 4 + i]; return r; }
__device__ void store_mat(chord_float4x4& c, const M4& a)
{ for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c.m[j * 4 + i] = a.m[i][j]; }

// cascade_setup.hlsl:57-79
__device__ float log_cascade_split(float nearZ, float farPlane, float nearPlane, float clipRange, uint32_t cascadeId, uint32_t cascadeCount, float lambda)
{
    const float range = farPlane - nearPlane, ratio = farPlane / nearPlane;
    const float p = (float)(cascadeId + 1u) / (float)cascadeCount;
    const float logScale = nearPlane * powf(fabsf(ratio), p);
    const float uniformScale = nearPlane + range * p;
    const float d = lambda * (logScale - uniformScale) + uniformScale;
    return (d - nearZ) / clipRange;
}

__device__ __forceinline__ bool cascade_cache_valid(uint32_t bCacheValid, uint32_t realtimeCount, uint32_t cascadeCount, uint32_t tick, uint32_t cascadeId)
{
    if (!bCacheValid || cascadeId < realtimeCount) return false;
    return (tick % (cascadeCount - realtimeCount)) != (cascadeId - realtimeCount);
}

__global__ void __launch_bounds__(32) k_cascade_setup(CascadeSetupParams prm)
{
    const chord_cascade_setup_desc& pc = prm.d;
    const uint32_t cascadeId = threadIdx.x;
    const bool active = cascadeId < pc.cascadeCount;
    const chord_PerframeCameraView& perView = *prm.view;
    const V3 upDir = v3(0.0f, 1.0f, 0.0f);
    const float nearZ = perView.zNear, farZ = perView.zFar, clipRange = farZ - nearZ;
    const uint32_t cid = active ? cascadeId : 0u;

    float minZ, maxZ, splitLambda, splitStart; uint32_t splitCount, splitId;
    if (cid < pc.realtimeCount) {
        minZ = nearZ + pc.cascadeStartDistance; maxZ = nearZ + pc.cascadeEndDistance;
        splitLambda = pc.splitLambda; splitCount = pc.realtimeCount; splitId = cid;
        if (prm.validRange) {
            const float minZValid = __uint_as_float(prm.validRange[0]), maxZValid = __uint_as_float(prm.validRange[1]);
            if (maxZValid > 0.0f) minZ = fmaxf(minZ, nearZ / maxZValid);
            if (minZValid > 0.0f) {
                const float stable = pc.cascadeEndDistance - pc.cascadeStartDistance;
                maxZ = fmaxf(maxZ, minZ * 1.1f);
                maxZ = fminf(maxZ, minZ + stable);
                maxZ = fminf(maxZ, nearZ / minZValid);
            }
        }
        splitStart = minZ - nearZ;
    } else {
        maxZ = nearZ + pc.farCascadeEndDistance; splitLambda = pc.farCascadeSplitLambda;
        minZ = nearZ + pc.cascadeEndDistance; splitStart = pc.cascadeEndDistance;
        splitCount = pc.cascadeCount - pc.realtimeCount; splitId = cid - pc.realtimeCount;
    }
    const float splitDist = log_cascade_split(nearZ, maxZ, minZ, clipRange, splitId, splitCount, splitLambda);
    const float prevSplitDist = (splitId == 0u) ? splitStart / clipRange : log_cascade_split(nearZ, maxZ, minZ, clipRange, splitId - 1u, splitCount, splitLambda);
    const float splitDist0 = log_cascade_split(nearZ, nearZ, nearZ + pc.farCascadeEndDistance, clipRange, 0u, pc.cascadeCount, pc.farCascadeSplitLambda);

    const float ndc[8][3] = { {-1, 1, 1}, {1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {-1, 1, 0}, {1, 1, 0}, {1, -1, 0}, {-1, -1, 0} };
    const M4 invView = load_mat(perView.clipToTranslatedWorldWithZFar_NoJitter);
    V3 base[8];
    for (int i = 0; i < 8; i++) { float h[4]; mul_point(invView, ndc[i][0], ndc[i][1], ndc[i][2], 1.0f, h); base[i] = v3(h[0] / h[3], h[1] / h[3], h[2] / h[3]); }
    V3 c[8], c0[8];
    for (int i = 0; i < 4; i++) {
        const V3 ray = base[i + 4] - base[i];
        c0[i + 4] = base[i] + ray * splitDist0; c0[i] = base[i] + ray * 0.0f;
        c[i + 4] = base[i] + ray * splitDist;   c[i] = base[i] + ray * prevSplitDist;
    }
    V3 center = v3(0, 0, 0), center0 = v3(0, 0, 0);
    for (int i = 0; i < 8; i++) { center = center + c[i]; center0 = center0 + c0[i]; }
    center = v3(center.x / 8.0f, center.y / 8.0f, center.z / 8.0f); center0 = v3(center0.x / 8.0f, center0.y / 8.0f, center0.z / 8.0f);
    float sphereRadius = 0.0f, sphereRadius0 = 0.0f;
    for (int i = 0; i < 8; i++) { sphereRadius = fmaxf(sphereRadius, length(c[i] - center)); sphereRadius0 = fmaxf(sphereRadius0, length(c0[i] - center0)); }
    const float cascadeSphereRadius = ceilf(sphereRadius * 16.0f) / 16.0f;
    float maxCascadeSphereRadius = active ? cascadeSphereRadius : 0.0f;                    // WaveActiveMax over the cascades
    for (int o = 16; o > 0; o >>= 1) maxCascadeSphereRadius = fmaxf(maxCascadeSphereRadius, __shfl_xor_sync(0xffffffffu, maxCascadeSphereRadius, o));
    if (!active) return;

    const float extXY = cascadeSphereRadius, extZ = maxCascadeSphereRadius * 2.0f;
    float radiusScale, zStartBiasScale;
    if (cascadeId >= pc.realtimeCount) { radiusScale = sphereRadius0 / sphereRadius; zStartBiasScale = 1.0f; }
    else {
        radiusScale = 10.0f * pc.radiusScale / sphereRadius;
        radiusScale = radiusScale / (radiusScale + 1.0f);
        zStartBiasScale = 0.25f + (minZ - nearZ) / (pc.cascadeEndDistance - pc.cascadeStartDistance);
    }
    radiusScale = fminf(radiusScale, 1.0f);

    const V3 light = normalize(v3(pc.lightDir[0], pc.lightDir[1], pc.lightDir[2]));
    const V3 eye = center - light * extZ * 0.5f;
    // lookAt_RH (base.hlsli:637-666), ortho_RH_ZeroOne with near / far swapped for reversed z (base.hlsli:668-682)
    const V3 fwd = normalize(center - eye), side = normalize(cross(fwd, upDir)), up = cross(side, fwd);
    M4 view; memset(&view, 0, sizeof(view));
    view.m[0][0] = side.x; view.m[0][1] = side.y; view.m[0][2] = side.z; view.m[0][3] = -dot(side, eye);
    view.m[1][0] = up.x;   view.m[1][1] = up.y;   view.m[1][2] = up.z;   view.m[1][3] = -dot(up, eye);
    view.m[2][0] = -fwd.x; view.m[2][1] = -fwd.y; view.m[2][2] = -fwd.z; view.m[2][3] = dot(fwd, eye);
    view.m[3][3] = 1.0f;
    const float l = -extXY, r = extXY, b = -extXY, t = extXY, zn = extZ, zf = 0.0f;
    M4 proj; memset(&proj, 0, sizeof(proj));
    proj.m[0][0] = 2.0f / (r - l); proj.m[1][1] = 2.0f / (t - b); proj.m[2][2] = -1.0f / (zf - zn);
    proj.m[0][3] = -(r + l) / (r - l); proj.m[1][3] = -(t + b) / (t - b); proj.m[2][3] = -zn / (zf - zn); proj.m[3][3] = 1.0f;

    // texel alignment (cascade_setup.hlsl:283-299)
    const float sMapSize = (float)pc.cascadeDim;
    float origin[4]; { const M4 vp0 = mul(proj, view); mul_point(vp0, 0.0f, 0.0f, 0.0f, 1.0f, origin); }
    const float ox = origin[0] * (sMapSize / 2.0f), oy = origin[1] * (sMapSize / 2.0f);
    proj.m[0][3] += (rintf(ox) - ox) * (2.0f / sMapSize);
    proj.m[1][3] += (rintf(oy) - oy) * (2.0f / sMapSize);
    const M4 vp = mul(proj, view), inv = inverse(vp);

    V3 p[8];
    for (int i = 0; i < 8; i++) { float h[4]; mul_point(inv, ndc[i][0], ndc[i][1], ndc[i][2], 1.0f, h); p[i] = v3(h[0] / h[3], h[1] / h[3], h[2] / h[3]); }
    chord_InstanceCullingViewInfo vi; memset(&vi, 0, sizeof(vi));
    auto plane = [&](int k, V3 a, V3 bq, V3 o, V3* nOut, const V3* dNormal) {
        const V3 n = normalize(cross(a - o, bq - o));
        vi.frustumPlanesRS[k].x = n.x; vi.frustumPlanesRS[k].y = n.y; vi.frustumPlanesRS[k].z = n.z;
        vi.frustumPlanesRS[k].w = -dot(dNormal ? *dNormal : n, o);
        if (nOut) *nOut = n;
    };
    V3 frontN;
    plane(0, p[4], p[3], p[7], nullptr, nullptr);          // left
    plane(1, p[6], p[3], p[2], nullptr, nullptr);          // down
    plane(2, p[6], p[1], p[5], nullptr, nullptr);          // right
    plane(3, p[5], p[0], p[4], nullptr, nullptr);          // top
    plane(4, p[1], p[3], p[0], &frontN, nullptr);          // front
    plane(5, p[5], p[7], p[6], nullptr, &frontN);          // back: cascade_setup.hlsl:349 uses frontN in the distance term
    store_mat(vi.translatedWorldToClip, vp);
    store_mat(vi.clipToTranslatedWorld, inv);
    vi.cameraWorldPos = perView.cameraWorldPos;
    vi.orthoDepthConvertToView.x = proj.m[2][2]; vi.orthoDepthConvertToView.y = proj.m[2][3];
    vi.orthoDepthConvertToView.z = zStartBiasScale; vi.orthoDepthConvertToView.w = radiusScale;
    vi.renderDimension.x = sMapSize; vi.renderDimension.y = sMapSize; vi.renderDimension.z = 1.0f / sMapSize; vi.renderDimension.w = 1.0f / sMapSize;
    if (!cascade_cache_valid(pc.bCacheValid, pc.realtimeCount, pc.cascadeCount, pc.tickCount, cascadeId)) prm.out[cascadeId] = vi;
}

int launch_cascade_setup(chord_frame* mainFrame, const chord_cascade_setup_desc* desc, const chord_hzb* validRangeSource, chord_InstanceCullingViewInfo* dViews)
{
    CascadeSetupParams prm;
    prm.d = *desc; prm.view = mainFrame->dView;
    prm.validRange = validRangeSource ? validRangeSource->validRange : nullptr;
    prm.out = dViews;
    k_cascade_setup<<<1, 32, 0, mainFrame->ctx->stream>>>(prm);
    mainFrame->ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

}  // namespace chordb
