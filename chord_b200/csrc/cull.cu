// cull.cu — cluster culling for the visibility-buffer path (sm_100a).
//
//   k_object_prepare     per-object derived matrices (what every HLSL thread recomputes with mul()) + the object -> primitive ->
//                        datas chain resolved once per object (ObjectResolved)
//   k_object_cull_scan   K1  instance_culling.hlsl:47-131   object OBB vs frustum -> group count, and - last block done - the
//                        exclusive scan of the counts (canonical object order replaces the InterlockedAdd)
//   k_group_cull         K2  instance_culling.hlsl:133-208  LOD-DAG cut + cone/frustum per meshlet, ordered compaction
//   k_hzb_cull<phase>    K3  hzb_mainview_culling.hlsl:35-213 two-phase HZB occlusion, ordered compaction x2
//
// All compactions are single-pass decoupled look-back scans (common.cuh) so that the output order is the
// canonical (object, group, meshlet-in-group) order, fixing the reference's run-to-run nondeterministic slot ids
//  while staying one pass over the data.
#include "common.cuh"

#include <cstring>

namespace chordb {

constexpr int kCullBlock = 256;

// ------------------------------------------------------------------------------------------------ helpers
struct UVZ { float x, y, z; };

// base.hlsli:161-171 projectPosToUVz
__device__ __forceinline__ UVZ project_pos_to_uvz(const float* __restrict__ M, float px, float py, float pz)
{
    f4 h = mat_mul_vec(M, px, py, pz, 1.0f);
    UVZ o;
    o.x = __fmaf_rn(__fdiv_rn(h.x, h.w), 0.5f, 0.5f);
    o.y = __fmaf_rn(__fdiv_rn(h.y, h.w), -0.5f, 0.5f);
    o.z = __fdiv_rn(h.z, h.w);
    return o;
}

// base.hlsli:275-305: true if culled
__device__ bool frustum_culling(const chord_float4* __restrict__ planes, f3 c, f3 e, const float* __restrict__ l2w)
{
    f3 p[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float x = __fmaf_rn(e.x, kExtentApplyFactor[k][0], c.x);
        float y = __fmaf_rn(e.y, kExtentApplyFactor[k][1], c.y);
        float z = __fmaf_rn(e.z, kExtentApplyFactor[k][2], c.z);
        f4 w = mat_mul_vec(l2w, x, y, z, 1.0f);
        p[k].x = w.x; p[k].y = w.y; p[k].z = w.z;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const chord_float4 pl = planes[i];
        bool allBack = true;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (dot3f(pl.x, pl.y, pl.z, p[j].x, p[j].y, p[j].z) > -pl.w) allBack = false;
        if (allBack) return true;
    }
    return false;
}

// base.hlsli:248-272: true if culled
__device__ bool ortho_frustum_culling(f3 c, f3 e, const float* __restrict__ localToClip)
{
    float mn0 = 10.0f, mn1 = 10.0f, mx0 = -10.0f, mx1 = -10.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float x = __fmaf_rn(e.x, kExtentApplyFactor[k][0], c.x);
        float y = __fmaf_rn(e.y, kExtentApplyFactor[k][1], c.y);
        float z = __fmaf_rn(e.z, kExtentApplyFactor[k][2], c.z);
        UVZ u = project_pos_to_uvz(localToClip, x, y, z);
        mn0 = fminf(mn0, u.x); mn1 = fminf(mn1, u.y);
        mx0 = fmaxf(mx0, u.x); mx1 = fmaxf(mx1, u.y);
    }
    return (mn0 >= 1.0f) || (mn1 >= 1.0f) || (mx0 <= 0.0f) || (mx1 <= 0.0f);
}

// base.hlsli:500-518
__device__ __forceinline__ float project_sphere_to_screen(float sx, float sy, float sz, float r, float height, float tanHalfFovy)
{
    const float d2 = dot3f(sx, sy, sz, sx, sy, sz);
    const float r2 = __fmul_rn(r, r);
    if (d2 <= r2) return -1.0f;
    return __fdiv_rn(__fmul_rn(__fdiv_rn(__fmul_rn(height, 0.5f), tanHalfFovy), r), __fsqrt_rn(__fsub_rn(d2, r2)));
}

// nanite_shared.hlsli:15-49
__device__ bool is_meshlet_group_visible(float renderHeight, float tanHalfFovy, float maxScale, const float* __restrict__ localToView,
                                         const chord_GPUGLTFMeshletGroup& g)
{
    const bool bFinalLod = g.parentError > CHORD_ERROR_RADIUS_ROOT;
    const bool bFirstLod = g.error < -0.5f;
    if (!bFinalLod) {
        f4 v = mat_mul_vec(localToView, g.parentPosCenter.x, g.parentPosCenter.y, g.parentPosCenter.z, 1.0f);
        float pe = project_sphere_to_screen(v.x, v.y, v.z, __fmul_rn(maxScale, g.parentError), renderHeight, tanHalfFovy);
        if (pe > 0.0f && pe <= CHORD_ERROR_PIXEL_THRESHOLD) return false;
    }
    if (!bFirstLod) {
        f4 v = mat_mul_vec(localToView, g.clusterPosCenter.x, g.clusterPosCenter.y, g.clusterPosCenter.z, 1.0f);
        float e = project_sphere_to_screen(v.x, v.y, v.z, __fmul_rn(maxScale, g.error), renderHeight, tanHalfFovy);
        if (e < 0.0f || e > CHORD_ERROR_PIXEL_THRESHOLD) return false;
    }
    return true;
}

__device__ __forceinline__ void meshlet_center_extent(const chord_GPUGLTFMeshlet& m, bool halfFirst, f3& c, f3& e)
{
    if (halfFirst) {    // nanite_shared.hlsli:74: 0.5 * (posMin + posMax)
        c.x = __fmul_rn(0.5f, __fadd_rn(m.posMin.x, m.posMax.x)); c.y = __fmul_rn(0.5f, __fadd_rn(m.posMin.y, m.posMax.y)); c.z = __fmul_rn(0.5f, __fadd_rn(m.posMin.z, m.posMax.z));
    } else {            // hzb_mainview_culling.hlsl:67: (posMin + posMax) * 0.5  (same value: multiplication commutes)
        c.x = __fmul_rn(__fadd_rn(m.posMin.x, m.posMax.x), 0.5f); c.y = __fmul_rn(__fadd_rn(m.posMin.y, m.posMax.y), 0.5f); c.z = __fmul_rn(__fadd_rn(m.posMin.z, m.posMax.z), 0.5f);
    }
    e.x = __fsub_rn(m.posMax.x, c.x); e.y = __fsub_rn(m.posMax.y, c.y); e.z = __fsub_rn(m.posMax.z, c.z);
}

// nanite_shared.hlsli:51-91
__device__ bool is_meshlet_visible(uint32_t switchFlags, const chord_float4* __restrict__ planes, const chord_GPUObjectGLTFPrimitive& obj,
                                   const float* __restrict__ localToClip, const chord_GPUGLTFMeshlet& m, uint32_t bTwoSided)
{
    bool visible = true;
    if (bTwoSided == 0 && has_flag(switchFlags, CHORD_FLAG_CONE_CULL_BIT)) {
        f4 cam = mat_mul_vec(obj.basicData.translatedWorldToLocal.m, 0.0f, 0.0f, 0.0f, 1.0f);
        float vx = __fsub_rn(m.coneApex.x, cam.x), vy = __fsub_rn(m.coneApex.y, cam.y), vz = __fsub_rn(m.coneApex.z, cam.z);
        float len = __fsqrt_rn(dot3f(vx, vy, vz, vx, vy, vz));
        vx = __fdiv_rn(vx, len); vy = __fdiv_rn(vy, len); vz = __fdiv_rn(vz, len);
        if (dot3f(vx, vy, vz, m.coneAxis.x, m.coneAxis.y, m.coneAxis.z) >= m.coneCutOff) visible = false;
    }
    if (visible && has_flag(switchFlags, CHORD_FLAG_FRUSTUM_CULL_BIT)) {
        f3 c, e; meshlet_center_extent(m, true, c, e);
        if (localToClip[15] == 1.0f) visible = !ortho_frustum_culling(c, e, localToClip);
        else                         visible = !frustum_culling(planes, c, e, obj.basicData.localToTranslatedWorld.m);
    }
    return visible;
}

// ------------------------------------------------------------------------------------------------ per-object prepare
__global__ void __launch_bounds__(96) k_object_prepare(FrameDev f, ObjectDerived* __restrict__ out, ObjectResolved* __restrict__ res)
{
    // one object per block, 96 threads = 6 matrices x 16 elements
    const uint32_t o = blockIdx.x;
    const uint32_t which = threadIdx.x >> 4, e = threadIdx.x & 15u, c = e >> 2, r = e & 3u;
    const chord_GPUObjectGLTFPrimitive& obj = f.objects[o];
    const float* A; const float* B; float* dst;
    switch (which) {
    case 0:  A = f.iview->translatedWorldToClip.m;          B = obj.basicData.localToTranslatedWorld.m;          dst = out[o].localToClip.m; break;
    case 1:  A = f.view->translatedWorldToView.m;           B = obj.basicData.localToTranslatedWorld.m;          dst = out[o].localToView.m; break;
    case 2:  A = f.view->translatedWorldToClipLastFrame.m;  B = obj.basicData.localToTranslatedWorldLastFrame.m; dst = out[o].localToClipHzb0.m; break;
    case 4:  A = f.view->translatedWorldToClip_NoJitter.m;  B = obj.basicData.localToTranslatedWorld.m;          dst = out[o].localToClipNoJitter.m; break;
    case 5:  A = f.view->translatedWorldToClipLastFrame_NoJitter.m; B = obj.basicData.localToTranslatedWorldLastFrame.m; dst = out[o].localToClipLastNoJitter.m; break;
    default: A = f.view->translatedWorldToClip.m;           B = obj.basicData.localToTranslatedWorld.m;          dst = out[o].localToClipHzb1.m; break;
    }
    dst[c * 4 + r] = dot4f(A[r], A[4 + r], A[8 + r], A[12 + r], B[c * 4], B[c * 4 + 1], B[c * 4 + 2], B[c * 4 + 3]);
    if (threadIdx.x == 95u) {
        // object -> primitive -> datas, followed once per object instead of once per cull thread
        const chord_GLTFPrimitiveBuffer& prim = f.scene.prims[obj.GLTFPrimitiveDetail];
        const DatasDev& datas = f.scene.datas[prim.primitiveDatasBufferId];
        ObjectResolved rr;
        rr.meshlets = datas.meshlets;
        rr.groups = datas.groups + prim.meshletGroupOffset;
        rr.groupIndices = datas.groupIndices + prim.meshletGroupIndicesOffset;
        rr.meshletOffset = prim.meshletOffset;
        rr.twoSided = f.scene.mats[obj.GLTFMaterialData].bTwoSided;
        rr.groupCount = prim.meshletGroupCount;
        rr.pad = 0u;
        res[o] = rr;
    }
}

// ------------------------------------------------------------------------------------------------ K1
// instanceCullingCS (instance_culling.hlsl:47-131): per-object frustum test -> number of cluster groups the object contributes,
// and - in the same launch - the exclusive scan of those counts (the reference's InterlockedAdd slot allocation made ordered):
// the block that finishes last scans all counts ("last block done"; one block when there are <= 128 objects).
__global__ void __launch_bounds__(128) k_object_cull_scan(FrameDev f, uint32_t* __restrict__ groupCount, uint32_t* __restrict__ groupBase,
                                                           uint32_t* __restrict__ total, uint32_t* __restrict__ blocksDone)
{
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o < f.objectCount) {
        const chord_GPUObjectGLTFPrimitive& obj = f.objects[o];
        const chord_GLTFPrimitiveBuffer& prim = f.scene.prims[obj.GLTFPrimitiveDetail];
        bool visible = true;
        if (has_flag(f.switchFlags, CHORD_FLAG_FRUSTUM_CULL_BIT)) {
            f3 c, e;
            c.x = __fmul_rn(__fadd_rn(prim.posMin.x, prim.posMax.x), 0.5f); c.y = __fmul_rn(__fadd_rn(prim.posMin.y, prim.posMax.y), 0.5f); c.z = __fmul_rn(__fadd_rn(prim.posMin.z, prim.posMax.z), 0.5f);
            e.x = __fsub_rn(prim.posMax.x, c.x); e.y = __fsub_rn(prim.posMax.y, c.y); e.z = __fsub_rn(prim.posMax.z, c.z);
            const float* l2c = f.derived[o].localToClip.m;
            if (l2c[15] == 1.0f) visible = !ortho_frustum_culling(c, e, l2c);
            else                 visible = !frustum_culling(f.iview->frustumPlanesRS, c, e, obj.basicData.localToTranslatedWorld.m);
        }
        groupCount[o] = visible ? prim.meshletGroupCount : 0u;
    }
    // ---- last block done
    __shared__ uint32_t sLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sLast = (atomicAdd(blocksDone, 1u) + 1u == gridDim.x) ? 1u : 0u;
    __syncthreads();
    if (!sLast) return;
    if (threadIdx.x == 0) *blocksDone = 0u;                                  // re-armed for the next launch (stream-ordered)
    __threadfence();
    // single block exclusive scan of the counts written by all blocks (read past L1: other SMs wrote them)
    __shared__ uint32_t warpSums[4];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n = f.objectCount;
    for (uint32_t base = 0; base < n; base += 128) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < n) ? __ldcg(groupCount + i) : 0u;
        uint32_t s = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, s, k); if (lane >= (uint32_t)k) s += t; }
        if (lane == 31) warpSums[warp] = s;
        __syncthreads();
        uint32_t warpBase = 0, blockTotal = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) { const uint32_t t = warpSums[w]; if (w < (int)warp) warpBase += t; blockTotal += t; }
        if (i < n) groupBase[i] = carry + warpBase + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += blockTotal;
        __syncthreads();
    }
    if (threadIdx.x == 0) { *total = carry; groupBase[n] = carry; }
}

// ------------------------------------------------------------------------------------------------ K2
__global__ void __launch_bounds__(kCullBlock) k_group_cull(FrameDev f, const uint32_t* __restrict__ objGroupBase, const uint32_t* __restrict__ totalGroups,
                                                           unsigned long long* __restrict__ scanStatus, uint32_t* __restrict__ ticket,
                                                           chord_draw_cmd* __restrict__ outCmds, uint32_t* __restrict__ outCount, uint32_t outCapacity,
                                                           const uint32_t* __restrict__ candidateBits, const uint32_t* __restrict__ bvhCtrl)
{
    // candidateBits != null: only groups the BVH traversal reached are evaluated (bvhCtrl[4] != 0: queue overflowed, ignore the bits)
    const bool useBits = candidateBits != nullptr && bvhCtrl[4] == 0u;
    __shared__ uint32_t sTile;
    __shared__ uint32_t sWarp[kCullBlock / 32];
    __shared__ uint32_t sBase;
    const uint32_t total = *totalGroups;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // the last tile publishes the list count (below); with nothing to do nobody would, so the count is zeroed here instead of
    // by a memset in front of every cull kernel
    if (total == 0u && blockIdx.x == 0 && threadIdx.x == 0) *outCount = 0u;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = sTile;
        if ((unsigned long long)tile * kCullBlock >= total) break;
        const uint32_t t = tile * kCullBlock + threadIdx.x;

        uint32_t visCount = 0, visId[CHORD_CLUSTER_GROUP_MERGE_MAX], objectId = 0;
        if (t < total && (!useBits || ((candidateBits[t >> 5] >> (t & 31u)) & 1u))) {
            // canonical order: thread t <-> (object, group) through the scanned bases (binary search: last base <= t)
            uint32_t lo = 0, hi = f.objectCount;
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (objGroupBase[mid] <= t) lo = mid; else hi = mid; }
            // skip culled objects that share the same base (count 0): lo is the last one with base <= t, which is the owner
            objectId = lo;
            const uint32_t groupId = t - objGroupBase[lo];
            const chord_GPUObjectGLTFPrimitive& obj = f.objects[objectId];
            const ObjectResolved res = f.resolved[objectId];                     // object -> primitive -> datas, resolved by k_object_prepare
            const chord_GPUGLTFMeshletGroup g = res.groups[groupId];
            const ObjectDerived& der = f.derived[objectId];
            if (is_meshlet_group_visible(f.view->renderDimension.y, f.tanHalfFovy, obj.basicData.scaleExtractFromMatrix.w, der.localToView.m, g)) {
                for (uint32_t i = 0; i < g.meshletCount && i < CHORD_CLUSTER_GROUP_MERGE_MAX; i++) {
                    const uint32_t meshletIndex = res.meshletOffset + res.groupIndices[i + g.meshletOffset];
                    const chord_GPUGLTFMeshlet m = res.meshlets[meshletIndex];
                    if (is_meshlet_visible(f.switchFlags, f.iview->frustumPlanesRS, obj, der.localToClip.m, m, res.twoSided))
                        visId[visCount++] = meshletIndex;
                }
            }
        }
        // block exclusive scan of visCount
        uint32_t s = visCount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (uint32_t)o) s += v; }
        if (lane == 31) sWarp[warp] = s;
        __syncthreads();
        uint32_t warpBase = 0, blockTotal = 0;
#pragma unroll
        for (int w = 0; w < kCullBlock / 32; w++) { uint32_t v = sWarp[w]; if (w < (int)warp) warpBase += v; blockTotal += v; }
        if (warp == 0) {
            unsigned long long ex = scan_tile_exclusive_prefix(scanStatus, tile, blockTotal);
            if (lane == 0) {
                sBase = (uint32_t)ex;
                if ((unsigned long long)(tile + 1) * kCullBlock >= total) *outCount = (uint32_t)ex + blockTotal;   // last tile publishes the count
            }
        }
        __syncthreads();
        const uint32_t dst = sBase + warpBase + s - visCount;
        for (uint32_t i = 0; i < visCount; i++) {
            const uint32_t slot = dst + i;
            if (slot < outCapacity) outCmds[slot] = chord_draw_cmd{objectId, visId[i], slot};
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3
struct HzbDev {
    const __half* minChain;
    uint64_t mipOffset[CHORD_HZB_MAX_MIP_COUNT];
    uint32_t mip0W, mip0H, mipCount;
};

template <bool kFirstStage>
__global__ void __launch_bounds__(kCullBlock) k_hzb_cull(FrameDev f, HzbDev hzb,
                                                         const chord_draw_cmd* __restrict__ inCmds, const uint32_t* __restrict__ inCount, uint32_t inCapacity,
                                                         unsigned long long* __restrict__ scanStatus, uint32_t* __restrict__ ticket,
                                                         chord_draw_cmd* __restrict__ outVis, uint32_t* __restrict__ outVisCount,
                                                         chord_draw_cmd* __restrict__ outRej, uint32_t* __restrict__ outRejCount)
{
    __shared__ uint32_t sTile;
    __shared__ uint32_t sWarpV[kCullBlock / 32], sWarpR[kCullBlock / 32];
    __shared__ uint32_t sBaseV, sBaseR;
    uint32_t total = *inCount; if (total > inCapacity) total = inCapacity;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const float W = f.view->renderDimension.x, H = f.view->renderDimension.y;
    if (total == 0u && blockIdx.x == 0 && threadIdx.x == 0) { *outVisCount = 0u; if (kFirstStage) *outRejCount = 0u; }      // see k_group_cull
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = sTile;
        if ((unsigned long long)tile * kCullBlock >= total) break;
        const uint32_t t = tile * kCullBlock + threadIdx.x;
        bool active = t < total, visible = true;
        chord_draw_cmd cmd = chord_draw_cmd{0, 0, 0};
        if (active) {
            cmd = inCmds[t];
            if (has_flag(f.switchFlags, CHORD_FLAG_HZB_CULL_BIT)) {
                const chord_GPUGLTFMeshlet m = f.resolved[cmd.objectId].meshlets[cmd.meshletId];      // (object -> primitive -> datas resolved by k_object_prepare)
                f3 c, e; meshlet_center_extent(m, false, c, e);
                const float* mvp = kFirstStage ? f.derived[cmd.objectId].localToClipHzb0.m : f.derived[cmd.objectId].localToClipHzb1.m;
                float mn0 = 10.0f, mn1 = 10.0f, mn2 = 10.0f, mx0 = -10.0f, mx1 = -10.0f, mx2 = -10.0f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float x = __fmaf_rn(e.x, kExtentApplyFactor[i][0], c.x);
                    float y = __fmaf_rn(e.y, kExtentApplyFactor[i][1], c.y);
                    float z = __fmaf_rn(e.z, kExtentApplyFactor[i][2], c.z);
                    UVZ u = project_pos_to_uvz(mvp, x, y, z);
                    mn0 = fminf(mn0, u.x); mn1 = fminf(mn1, u.y); mn2 = fminf(mn2, u.z);
                    mx0 = fmaxf(mx0, u.x); mx1 = fmaxf(mx1, u.y); mx2 = fmaxf(mx2, u.z);
                }
                const bool zInRange = (mx2 < 1.0f) && (mn2 > 0.0f);
                if (zInRange && (mn0 >= 1.0f || mn1 >= 1.0f || mx0 <= 0.0f || mx1 <= 0.0f)) visible = false;
                if (visible && zInRange) {
                    const float u0 = fminf(fmaxf(mn0, 0.0f), 1.0f), v0 = fminf(fmaxf(mn1, 0.0f), 1.0f);
                    const float u1 = fminf(fmaxf(mx0, 0.0f), 1.0f), v1 = fminf(fmaxf(mx1, 0.0f), 1.0f);
                    int rx0 = __float2int_rz(__fmaf_rn(u0, W, 0.5f)), ry0 = __float2int_rz(__fmaf_rn(v0, H, 0.5f));
                    int rx1 = __float2int_rz(__fmaf_rn(u1, W, -0.5f)), ry1 = __float2int_rz(__fmaf_rn(v1, H, -0.5f));
                    rx0 = max(rx0, 0); ry0 = max(ry0, 0);
                    rx1 = __float2int_rz(fminf(__fsub_rn(W, 1.0f), __int2float_rn(rx1)));
                    ry1 = __float2int_rz(fminf(__fsub_rn(H, 1.0f), __int2float_rn(ry1)));
                    if (rx1 < rx0 || ry1 < ry0) {
                        visible = false;
                    } else {
                        const int hx0 = rx0 >> 1, hy0 = ry0 >> 1, hx1 = rx1 >> 1, hy1 = ry1 >> 1;
                        const int dxv = hx1 - hx0, dyv = hy1 - hy0;
                        const int mlx = dxv > 0 ? 31 - __clz(dxv) : -1, mly = dyv > 0 ? 31 - __clz(dyv) : -1;
                        int mip = max(0, max(mlx, mly) - 1);
                        if (((hx1 >> mip) - (hx0 >> mip) >= 4) || ((hy1 >> mip) - (hy0 >> mip) >= 4)) mip += 1;
                        const int cx0 = hx0 >> mip, cy0 = hy0 >> mip, cx1 = hx1 >> mip, cy1 = hy1 >> mip;
                        const uint32_t pitch = max(1u, hzb.mip0W >> mip);      // mips narrower than one texel are stored 1 wide (hzb.cu mip_dim)
                        const __half* mp = hzb.minChain + hzb.mipOffset[mip];
                        float zMin = 10.0f;
#pragma unroll
                        for (int x = 0; x < 4; x++)
#pragma unroll
                            for (int y = 0; y < 4; y++) {
                                const int sx = min(cx1, cx0 + x), sy = min(cy1, cy0 + y);
                                zMin = fminf(zMin, __half2float(mp[(size_t)sy * pitch + (uint32_t)sx]));
                            }
                        if (zMin > mx2) visible = false;
                    }
                }
            }
        }
        const uint32_t vbit = (active && visible) ? 1u : 0u;
        const uint32_t rbit = (kFirstStage && active && !visible) ? 1u : 0u;
        const uint32_t vmask = __ballot_sync(0xffffffffu, vbit), rmask = __ballot_sync(0xffffffffu, rbit);
        const uint32_t lt = (1u << lane) - 1u;
        if (lane == 0) { sWarpV[warp] = __popc(vmask); sWarpR[warp] = __popc(rmask); }
        __syncthreads();
        uint32_t baseV = 0, baseR = 0, totV = 0, totR = 0;
#pragma unroll
        for (int w = 0; w < kCullBlock / 32; w++) {
            const uint32_t a = sWarpV[w], b = sWarpR[w];
            if (w < (int)warp) { baseV += a; baseR += b; }
            totV += a; totR += b;
        }
        if (warp == 0) {
            unsigned long long ex = scan_tile_exclusive_prefix(scanStatus, tile, (unsigned long long)totV | ((unsigned long long)totR << 31));
            if (lane == 0) {
                sBaseV = (uint32_t)(ex & 0x7FFFFFFFull); sBaseR = (uint32_t)((ex >> 31) & 0x7FFFFFFFull);
                if ((unsigned long long)(tile + 1) * kCullBlock >= total) {
                    *outVisCount = sBaseV + totV;
                    if (kFirstStage) *outRejCount = sBaseR + totR;
                }
            }
        }
        __syncthreads();
        if (vbit) outVis[sBaseV + baseV + __popc(vmask & lt)] = cmd;
        if (rbit) outRej[sBaseR + baseR + __popc(rmask & lt)] = cmd;
    }
}

// ------------------------------------------------------------------------------------------------ N3: hzbCullingGeneric
// hzb_culling_generic.hlsl:37-172: one-pass 2x2 test against the HZB of another view (previous cascade / cached cascade)
struct GenericCullParams {
    chord_float4x4 viewProj;            // hzbView.translatedWorldToClip
    float rel[3];                       // float3(perView.cameraWorldPos - hzbView.cameraWorldPos)
    float W, H, lo, hi;                 // hzbView.renderDimension.xy, extentScale * -/+ 0.5
    uint32_t useLastFrame;
    const chord_InstanceCullingViewInfo* dView;    // when set, viewProj / rel / W / H are read from this device record instead
};

__global__ void __launch_bounds__(kCullBlock) k_hzb_cull_generic(FrameDev f, HzbDev hzb, GenericCullParams prm,
                                                                 const chord_draw_cmd* __restrict__ inCmds, const uint32_t* __restrict__ inCount, uint32_t inCapacity,
                                                                 unsigned long long* __restrict__ scanStatus, uint32_t* __restrict__ ticket,
                                                                 chord_draw_cmd* __restrict__ outVis, uint32_t* __restrict__ outVisCount)
{
    __shared__ uint32_t sTile;
    __shared__ uint32_t sWarpV[kCullBlock / 32];
    __shared__ uint32_t sBaseV;
    uint32_t total = *inCount; if (total > inCapacity) total = inCapacity;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (total == 0u && blockIdx.x == 0 && threadIdx.x == 0) *outVisCount = 0u;      // see k_group_cull
    __shared__ GenericCullParams sPrm;
    if (prm.dView) {
        // the HZB's view lives on the device (cascade buffer): resolve the derived parameters once per block
        if (threadIdx.x < 16) sPrm.viewProj.m[threadIdx.x] = prm.dView->translatedWorldToClip.m[threadIdx.x];
        if (threadIdx.x == 16) {
            const chord_uint2* a = &f.view->cameraWorldPos.x; const chord_uint2* b = &prm.dView->cameraWorldPos.x;
            for (int k = 0; k < 3; k++) {
                const double da = __longlong_as_double((long long)((unsigned long long)a[k].x | ((unsigned long long)a[k].y << 32)));
                const double db = __longlong_as_double((long long)((unsigned long long)b[k].x | ((unsigned long long)b[k].y << 32)));
                sPrm.rel[k] = (float)(da - db);
            }
            sPrm.W = prm.dView->renderDimension.x; sPrm.H = prm.dView->renderDimension.y;
            sPrm.lo = prm.lo; sPrm.hi = prm.hi; sPrm.useLastFrame = prm.useLastFrame;
        }
        __syncthreads();
        prm = sPrm;
    }
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = sTile;
        if ((unsigned long long)tile * kCullBlock >= total) break;
        const uint32_t t = tile * kCullBlock + threadIdx.x;
        bool active = t < total, visible = true;
        chord_draw_cmd cmd = chord_draw_cmd{0, 0, 0};
        if (active) {
            cmd = inCmds[t];
            if (has_flag(f.switchFlags, CHORD_FLAG_HZB_CULL_BIT)) {
                const chord_GPUGLTFMeshlet m = f.resolved[cmd.objectId].meshlets[cmd.meshletId];      // (object -> primitive -> datas resolved by k_object_prepare)
                f3 c, e; meshlet_center_extent(m, false, c, e);
                const chord_GPUObjectGLTFPrimitive& obj = f.objects[cmd.objectId];
                float l2w[16], mvp[16];
                const float* src = prm.useLastFrame ? obj.basicData.localToTranslatedWorldLastFrame.m : obj.basicData.localToTranslatedWorld.m;
#pragma unroll
                for (int i = 0; i < 16; i++) l2w[i] = src[i];
                l2w[12] = __fadd_rn(l2w[12], prm.rel[0]); l2w[13] = __fadd_rn(l2w[13], prm.rel[1]); l2w[14] = __fadd_rn(l2w[14], prm.rel[2]);
                mat_mul(prm.viewProj.m, l2w, mvp);
                float mn0 = 10.0f, mn1 = 10.0f, mn2 = 10.0f, mx0 = -10.0f, mx1 = -10.0f, mx2 = -10.0f;
                bool can = true;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float x = __fmaf_rn(e.x, kExtentApplyFactor[i][0], c.x);
                    float y = __fmaf_rn(e.y, kExtentApplyFactor[i][1], c.y);
                    float z = __fmaf_rn(e.z, kExtentApplyFactor[i][2], c.z);
                    UVZ u = project_pos_to_uvz(mvp, x, y, z);
                    mn0 = fminf(mn0, u.x); mn1 = fminf(mn1, u.y); mn2 = fminf(mn2, u.z);
                    mx0 = fmaxf(mx0, u.x); mx1 = fmaxf(mx1, u.y); mx2 = fmaxf(mx2, u.z);
                    can = can && (u.x < 1.0f && u.y < 1.0f && u.z < 1.0f) && (u.x > 0.0f && u.y > 0.0f && u.z > 0.0f);
                }
                if (can) {
                    int rx0 = __float2int_rz(__fmaf_rn(mn0, prm.W, prm.lo)), ry0 = __float2int_rz(__fmaf_rn(mn1, prm.H, prm.lo));
                    int rx1 = __float2int_rz(__fmaf_rn(mx0, prm.W, prm.hi)), ry1 = __float2int_rz(__fmaf_rn(mx1, prm.H, prm.hi));
                    rx0 = max(rx0, 0); ry0 = max(ry0, 0);
                    rx1 = __float2int_rz(fminf(__fsub_rn(prm.W, 1.0f), __int2float_rn(rx1)));
                    ry1 = __float2int_rz(fminf(__fsub_rn(prm.H, 1.0f), __int2float_rn(ry1)));
                    if (rx1 < rx0 || ry1 < ry0) {
                        visible = false;
                    } else {
                        const int hx0 = rx0 >> 1, hy0 = ry0 >> 1, hx1 = rx1 >> 1, hy1 = ry1 >> 1;
                        const int dxv = hx1 - hx0, dyv = hy1 - hy0;
                        const int mlx = dxv > 0 ? 31 - __clz(dxv) : -1, mly = dyv > 0 ? 31 - __clz(dyv) : -1;
                        int mip = max(0, max(mlx, mly));
                        if (((hx1 >> mip) - (hx0 >> mip) >= 2) || ((hy1 >> mip) - (hy0 >> mip) >= 2)) mip += 1;     // 2x2 sample: bias one level
                        const int cx0 = hx0 >> mip, cy0 = hy0 >> mip, cx1 = hx1 >> mip, cy1 = hy1 >> mip;
                        const uint32_t pitch = max(1u, hzb.mip0W >> mip);      // mips narrower than one texel are stored 1 wide (hzb.cu mip_dim)
                        const __half* mp = hzb.minChain + hzb.mipOffset[mip];
                        float zMin = 10.0f;
#pragma unroll
                        for (int x = 0; x < 2; x++)
#pragma unroll
                            for (int y = 0; y < 2; y++) {
                                const int sx = min(cx1, cx0 + x), sy = min(cy1, cy0 + y);
                                zMin = fminf(zMin, __half2float(mp[(size_t)sy * pitch + (uint32_t)sx]));
                            }
                        if (zMin > mx2) visible = false;
                    }
                }
            }
        }
        const uint32_t vbit = (active && visible) ? 1u : 0u;
        const uint32_t vmask = __ballot_sync(0xffffffffu, vbit);
        const uint32_t lt = (1u << lane) - 1u;
        if (lane == 0) sWarpV[warp] = __popc(vmask);
        __syncthreads();
        uint32_t baseV = 0, totV = 0;
#pragma unroll
        for (int w = 0; w < kCullBlock / 32; w++) { const uint32_t a = sWarpV[w]; if (w < (int)warp) baseV += a; totV += a; }
        if (warp == 0) {
            unsigned long long ex = scan_tile_exclusive_prefix(scanStatus, tile, (unsigned long long)totV);
            if (lane == 0) {
                sBaseV = (uint32_t)ex;
                if ((unsigned long long)(tile + 1) * kCullBlock >= total) *outVisCount = sBaseV + totV;
            }
        }
        __syncthreads();
        if (vbit) outVis[sBaseV + baseV + __popc(vmask & lt)] = cmd;
    }
}

// ------------------------------------------------------------------------------------------------ N4: BVH traversal
// Persistent-warp traversal of the 8-ary cluster-group BVH (gltf.h:16-24; builder nanite_builder.cpp:77-416), which the
// reference uploads but never walks.  One warp per node: lanes 0..7 test the eight child spheres,
// a ballot compacts the survivors into a global work queue, all 32 lanes mark the node's leaf groups as candidates.
// ctrl = {head, tail, pending, done, overflow}.  A node sphere bounds the PARENT-error spheres below it, so a node that
// projects to <= 1 px (minus a 1e-4 margin) cannot contain a group that passes the LOD test: pruning never changes the
// final list, the exact per-group test still runs in k_group_cull on the candidates.
constexpr float kBvhPruneThreshold = CHORD_ERROR_PIXEL_THRESHOLD * (1.0f - 1e-4f);

__global__ void __launch_bounds__(128) k_bvh_seed(FrameDev f, const uint32_t* __restrict__ objGroupCount, unsigned long long* __restrict__ entries,
                                                  uint32_t capacity, uint32_t* __restrict__ ctrl)
{
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= f.objectCount || objGroupCount[o] == 0u) return;
    const uint32_t slot = atomicAdd(&ctrl[1], 1u);
    if (slot >= capacity) { ctrl[4] = 1u; return; }
    atomicAdd(&ctrl[2], 1u);
    st_volatile_u64(&entries[slot], (unsigned long long)o);           // node 0 in the high word
}

__global__ void __launch_bounds__(128) k_bvh_traverse(FrameDev f, const uint32_t* __restrict__ objGroupBase, unsigned long long* __restrict__ entries,
                                                      uint32_t capacity, uint32_t* __restrict__ ctrl, uint32_t* __restrict__ bits, uint32_t bitWords)
{
    const uint32_t lane = threadIdx.x & 31u;
    volatile uint32_t* vctrl = ctrl;
    // Termination protocol: ctrl[2] (pending) counts queue entries that are published or about to be and not yet finished.
    // A node RESERVES its children in pending before it publishes them and only then retires itself, so pending can reach 0
    // exactly once - when every entry ever published has been processed; the warp that brings it to 0 raises ctrl[3] (done).
    // Waiters leave on `done` alone (never on pending == 0, which a racing reader could see between a child's retirement and
    // its parent's reservation under the old "publish, then count" order).
    if (vctrl[1] == 0u) return;                                         // k_bvh_seed queued nothing: no warp will ever raise done
    while (true) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(&ctrl[0], 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= capacity) return;
        unsigned long long e;
        while (true) {
            e = ld_volatile_u64(&entries[idx]);
            if (e != ~0ull) break;
            if (vctrl[3] != 0u) {                                       // traversal finished ...
                e = ld_volatile_u64(&entries[idx]);                     // ... and this slot was never filled (re-read after done)
                if (e == ~0ull) return;
                break;
            }
            __nanosleep(64);
        }
        const uint32_t objectId = (uint32_t)(e & 0xFFFFFFFFull), nodeIndex = (uint32_t)(e >> 32);
        const chord_GPUObjectGLTFPrimitive& obj = f.objects[objectId];
        const chord_GLTFPrimitiveBuffer& prim = f.scene.prims[obj.GLTFPrimitiveDetail];
        const DatasDev& datas = f.scene.datas[prim.primitiveDatasBufferId];
        const chord_GPUBVHNode* node = &datas.bvhNodes[prim.bvhNodeOffset + nodeIndex];
        // leaves -> candidate bits (canonical index = base of the object + group index inside the primitive)
        const uint32_t leafOffset = node->leafMeshletGroupOffset, leafCount = node->leafMeshletGroupCount;
        const uint32_t base = objGroupBase[objectId] + leafOffset;
        for (uint32_t k = lane; k < leafCount; k += 32u) {
            const uint32_t b = base + k;
            if ((b >> 5) < bitWords) atomicOr(&bits[b >> 5], 1u << (b & 31u)); else ctrl[4] = 1u;     // caller under-reported meshletGroupCount
        }
        // children
        uint32_t child = CHORD_UNVALID_ID; bool keep = false;
        if (lane < CHORD_BVH_LEVEL_NODE_COUNT) child = node->children[lane];
        if (child != CHORD_UNVALID_ID) {
            const chord_float4 sp = datas.bvhNodes[prim.bvhNodeOffset + child].sphere;
            const f4 v = mat_mul_vec(f.derived[objectId].localToView.m, sp.x, sp.y, sp.z, 1.0f);
            const float pe = project_sphere_to_screen(v.x, v.y, v.z, __fmul_rn(obj.basicData.scaleExtractFromMatrix.w, sp.w), f.view->renderDimension.y, f.tanHalfFovy);
            keep = !(pe > 0.0f && pe <= kBvhPruneThreshold);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        uint32_t n = (uint32_t)__popc(mask), slotBase = 0;
        if (n) {
            if (lane == 0) slotBase = atomicAdd(&ctrl[1], n);
            slotBase = __shfl_sync(0xffffffffu, slotBase, 0);
            if (slotBase + n > capacity) { if (lane == 0) ctrl[4] = 1u; n = 0; }
            else {
                if (lane == 0) { atomicAdd(&ctrl[2], n); __threadfence(); }      // reserve BEFORE the entries become visible
                __syncwarp();
                if (keep) st_volatile_u64(&entries[slotBase + __popc(mask & ((1u << lane) - 1u))], (unsigned long long)objectId | ((unsigned long long)child << 32));
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            const uint32_t after = atomicAdd(&ctrl[2], 0xFFFFFFFFu) - 1u;       // retire this node
            if (after == 0u) { __threadfence(); vctrl[3] = 1u; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
int launch_object_prepare(chord_frame* fr)
{
    chord_ctx* ctx = fr->ctx;
    FrameDev f = make_frame_dev(fr);
    k_object_prepare<<<fr->objectCount, 96, 0, ctx->stream>>>(f, fr->dDerived, fr->dResolved);
    ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

static int launch_instance_culling_impl(chord_frame* fr, chord_cmdlist* out, bool useBvh);
int launch_instance_culling(chord_frame* fr, chord_cmdlist* out) { return launch_instance_culling_impl(fr, out, false); }
int launch_instance_culling_bvh(chord_frame* fr, chord_cmdlist* out) { return launch_instance_culling_impl(fr, out, true); }

static int launch_instance_culling_impl(chord_frame* fr, chord_cmdlist* out, bool useBvh)
{
    chord_ctx* ctx = fr->ctx;
    FrameDev f = make_frame_dev(fr);
    const uint32_t tiles = (fr->meshletGroupCount + kCullBlock - 1) / kCullBlock + 1;
    int rc = scan_reset(ctx, tiles);
    if (rc) return rc;
    k_object_cull_scan<<<(fr->objectCount + 127) / 128, 128, 0, ctx->stream>>>(f, fr->objGroupCount, fr->objGroupBase, fr->totalGroups, fr->totalGroups + 2);
    const uint32_t* bits = nullptr;
    if (useBvh) {
        for (uint32_t i = 0; i < ctx->scene.nDatas; i++) CHORD_REQUIRE(ctx->datasHost[i].bvhNodes != nullptr, "scene was uploaded without BVH nodes");
        const uint32_t cap = 2u * fr->meshletGroupCount + fr->objectCount + 64u;
        const uint32_t words = (fr->meshletGroupCount + 31u) / 32u + 1u;
        if (cap > fr->bvhCapacity) {
            CHORD_NO_RESIZE_IN_CAPTURE(ctx, "the BVH traversal queue");
            cudaFree(fr->bvhEntries); cudaFree(fr->bvhBits); fr->bvhEntries = nullptr; fr->bvhBits = nullptr; fr->bvhCapacity = 0;
            CHORD_CUDA(cudaMalloc(&fr->bvhEntries, sizeof(unsigned long long) * (size_t)cap));
            CHORD_CUDA(cudaMalloc(&fr->bvhBits, sizeof(uint32_t) * (size_t)words));
            if (!fr->bvhCtrl) CHORD_CUDA(cudaMalloc(&fr->bvhCtrl, 32));
            fr->bvhCapacity = cap; fr->bvhWords = words;
        }
        CHORD_CUDA(cudaMemsetAsync(fr->bvhEntries, 0xFF, sizeof(unsigned long long) * (size_t)cap, ctx->stream));
        CHORD_CUDA(cudaMemsetAsync(fr->bvhBits, 0, sizeof(uint32_t) * (size_t)words, ctx->stream));
        { const int rcFill = launch_fill_u32(ctx, fr->bvhCtrl, 0u, 0u, 8); if (rcFill) return rcFill; }
        k_bvh_seed<<<(fr->objectCount + 127) / 128, 128, 0, ctx->stream>>>(f, fr->objGroupCount, (unsigned long long*)fr->bvhEntries, cap, fr->bvhCtrl);
        k_bvh_traverse<<<(uint32_t)ctx->smCount * 2u, 128, 0, ctx->stream>>>(f, fr->objGroupBase, (unsigned long long*)fr->bvhEntries, cap, fr->bvhCtrl, fr->bvhBits, fr->bvhWords);
        ctx->launches += 2;
        bits = fr->bvhBits;
    }
    uint32_t grid = tiles < (uint32_t)ctx->smCount * 8u ? tiles : (uint32_t)ctx->smCount * 8u;
    k_group_cull<<<grid, kCullBlock, 0, ctx->stream>>>(f, fr->objGroupBase, fr->totalGroups, ctx->scanStatus, ctx->scanTicket,
                                                       out->cmds, out->count, out->capacity, bits, fr->bvhCtrl);
    ctx->launches += 2;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

int launch_hzb_culling(chord_frame* fr, const chord_hzb* hzb, int firstStage, chord_cmdlist* in, chord_cmdlist* outVis, chord_cmdlist* outRej)
{
    chord_ctx* ctx = fr->ctx;
    FrameDev f = make_frame_dev(fr);
    HzbDev h;
    h.minChain = hzb->minChain; h.mip0W = hzb->mip0W; h.mip0H = hzb->mip0H; h.mipCount = hzb->mipCount;
    for (int i = 0; i < CHORD_HZB_MAX_MIP_COUNT; i++) h.mipOffset[i] = hzb->mipOffset[i];
    const uint32_t tiles = (in->capacity + kCullBlock - 1) / kCullBlock + 1;
    int rc = scan_reset(ctx, tiles);
    if (rc) return rc;
    uint32_t grid = tiles < (uint32_t)ctx->smCount * 8u ? tiles : (uint32_t)ctx->smCount * 8u;
    if (firstStage)
        k_hzb_cull<true><<<grid, kCullBlock, 0, ctx->stream>>>(f, h, in->cmds, in->count, in->capacity, ctx->scanStatus, ctx->scanTicket,
                                                               outVis->cmds, outVis->count, outRej->cmds, outRej->count);
    else
        k_hzb_cull<false><<<grid, kCullBlock, 0, ctx->stream>>>(f, h, in->cmds, in->count, in->capacity, ctx->scanStatus, ctx->scanTicket,
                                                                outVis->cmds, outVis->count, nullptr, nullptr);
    ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

int launch_hzb_culling_generic(chord_frame* fr, const chord_hzb* hzb, const chord_InstanceCullingViewInfo* hzbView, float extentScale,
                               int useLastFrame, chord_cmdlist* in, chord_cmdlist* outVis, const chord_InstanceCullingViewInfo* dHzbView)
{
    chord_ctx* ctx = fr->ctx;
    FrameDev f = make_frame_dev(fr);
    HzbDev h;
    h.minChain = hzb->minChain; h.mip0W = hzb->mip0W; h.mip0H = hzb->mip0H; h.mipCount = hzb->mipCount;
    for (int i = 0; i < CHORD_HZB_MAX_MIP_COUNT; i++) h.mipOffset[i] = hzb->mipOffset[i];
    GenericCullParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.dView = dHzbView;
    if (hzbView) {
        prm.viewProj = hzbView->translatedWorldToClip;
        const chord_uint2* a = &fr->hView.cameraWorldPos.x; const chord_uint2* b = &hzbView->cameraWorldPos.x;
        for (int k = 0; k < 3; k++) {
            const uint64_t ua = (uint64_t)a[k].x | ((uint64_t)a[k].y << 32), ub = (uint64_t)b[k].x | ((uint64_t)b[k].y << 32);
            double da, db; memcpy(&da, &ua, 8); memcpy(&db, &ub, 8);
            prm.rel[k] = (float)(da - db);             // float3(perView.cameraWorldPos - instanceView.cameraWorldPos), hzb_culling_generic.hlsl:78
        }
        prm.W = hzbView->renderDimension.x; prm.H = hzbView->renderDimension.y;
    }
    prm.lo = extentScale * -0.5f; prm.hi = extentScale * 0.5f;
    prm.useLastFrame = useLastFrame ? 1u : 0u;
    const uint32_t tiles = (in->capacity + kCullBlock - 1) / kCullBlock + 1;
    int rc = scan_reset(ctx, tiles);
    if (rc) return rc;
    uint32_t grid = tiles < (uint32_t)ctx->smCount * 8u ? tiles : (uint32_t)ctx->smCount * 8u;
    k_hzb_cull_generic<<<grid, kCullBlock, 0, ctx->stream>>>(f, h, prm, in->cmds, in->count, in->capacity, ctx->scanStatus, ctx->scanTicket,
                                                             outVis->cmds, outVis->count);
    ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

}  // namespace chordb
