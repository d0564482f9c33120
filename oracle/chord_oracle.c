/*
 * chord_oracle.c — CPU restatement (plain C + OpenMP) of chord's visibility-buffer path.
 * TEST INFRASTRUCTURE ONLY, see chord_oracle.h.  PARITY UNPINNED (no reference goldens exist).
 *
 * Floating-point rules (compile with -ffp-contract=off so that nothing but the
 * explicit fmaf() calls fuses):
 *   dot products / matrix rows accumulate left to right:  r = a0*b0; r = fma(a1,b1,r); ...
 *   every `a*b + c` of the HLSL is one fma; divisions and sqrt are IEEE round-to-nearest;
 *   min/max are IEEE minNum/maxNum (fminf/fmaxf); float->int is truncation unless stated.
 */
#include "chord_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ small math */
typedef struct { float x, y, z, w; } f4;
typedef struct { float x, y, z; } f3;

static inline float dot3f(float ax, float ay, float az, float bx, float by, float bz)
{ float r = ax * bx; r = fmaf(ay, by, r); r = fmaf(az, bz, r); return r; }
static inline float dot4f(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw)
{ float r = ax * bx; r = fmaf(ay, by, r); r = fmaf(az, bz, r); r = fmaf(aw, bw, r); return r; }

/* C = A * B, column-major (C[c][r] = row r of A . column c of B) */
static void mat_mul(const chord_float4x4* A, const chord_float4x4* B, chord_float4x4* C)
{
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            C->m[c * 4 + r] = dot4f(A->m[0 + r], A->m[4 + r], A->m[8 + r], A->m[12 + r],
                                    B->m[c * 4 + 0], B->m[c * 4 + 1], B->m[c * 4 + 2], B->m[c * 4 + 3]);
}
/* M * (x,y,z,w) */
static inline f4 mat_mul_vec(const chord_float4x4* M, float x, float y, float z, float w)
{
    f4 o;
    o.x = dot4f(M->m[0], M->m[4], M->m[8],  M->m[12], x, y, z, w);
    o.y = dot4f(M->m[1], M->m[5], M->m[9],  M->m[13], x, y, z, w);
    o.z = dot4f(M->m[2], M->m[6], M->m[10], M->m[14], x, y, z, w);
    o.w = dot4f(M->m[3], M->m[7], M->m[11], M->m[15], x, y, z, w);
    return o;
}

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* base.hlsli:175-185 kExtentApplyFactor */
static const float kExtentApplyFactor[8][3] = {
    { 1,  1,  1}, {-1, -1, -1}, { 1,  1, -1}, { 1, -1,  1},
    {-1,  1,  1}, { 1, -1, -1}, {-1, -1,  1}, {-1,  1, -1},
};

/* ------------------------------------------------------------------ fp16 */
uint16_t oracle_f32_to_f16_rtz(float v)
{
    /* round toward zero, subnormals kept, inf/nan passed through */
    uint32_t u = f2u(v);
    uint32_t sign = (u >> 16) & 0x8000u;
    uint32_t e = (u >> 23) & 0xFFu;
    uint32_t m = u & 0x7FFFFFu;
    if (e == 0xFF) return (uint16_t)(sign | 0x7C00u | (m ? 0x200u : 0));
    int32_t he = (int32_t)e - 127 + 15;
    if (he >= 31) return (uint16_t)(sign | 0x7BFFu);        /* RTZ: overflow clamps to max finite */
    if (he <= 0) {
        if (he < -10) return (uint16_t)sign;
        m |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - he);               /* 14..24 */
        return (uint16_t)(sign | (m >> shift));
    }
    return (uint16_t)(sign | ((uint32_t)he << 10) | (m >> 13));
}
float oracle_f16_to_f32(uint16_t h)
{
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1Fu;
    uint32_t m = h & 0x3FFu;
    if (e == 0) {
        if (m == 0) return u2f(sign);
        /* subnormal: m * 2^-24 */
        float f = (float)m * 5.9604644775390625e-08f;
        return sign ? -f : f;
    }
    if (e == 31) return u2f(sign | 0x7F800000u | (m << 13));
    return u2f(sign | ((e - 15 + 127) << 23) | (m << 13));
}

/* ------------------------------------------------------------------ scene access */
typedef struct {
    const chord_GPUObjectGLTFPrimitive* object;
    const chord_GLTFPrimitiveBuffer* prim;
    const chord_primitive_datas_desc* datas;
    const chord_GLTFMaterialGPUData* material;
} object_refs;

static object_refs get_object(const oracle_frame* f, uint32_t objectId)
{
    object_refs r;
    r.object = &f->objects[objectId];
    r.prim = &f->scene->primitives[r.object->GLTFPrimitiveDetail];
    r.datas = &f->scene->datas[r.prim->primitiveDatasBufferId];
    r.material = &f->scene->materials[r.object->GLTFMaterialData];
    return r;
}

static inline int has_flag(uint32_t flags, uint32_t bit) { return (flags & (1u << bit)) != 0u; }

/* base.hlsli:161-171 projectPosToUVz */
static inline f3 project_pos_to_uvz(float px, float py, float pz, const chord_float4x4* M)
{
    f4 h = mat_mul_vec(M, px, py, pz, 1.0f);
    f3 o;
    float nx = h.x / h.w, ny = h.y / h.w, nz = h.z / h.w;
    o.x = fmaf(nx, 0.5f, 0.5f);
    o.y = fmaf(ny, -0.5f, 0.5f);
    o.z = nz;
    return o;
}

/* base.hlsli:275-305 frustumCulling(planes, centerLS, extentLS, localToTranslatedWorld): true if culled */
static int frustum_culling(const chord_float4 planes[6], f3 c, f3 e, const chord_float4x4* l2w)
{
    f3 p[8];
    for (int k = 0; k < 8; k++) {
        float x = fmaf(e.x, kExtentApplyFactor[k][0], c.x);
        float y = fmaf(e.y, kExtentApplyFactor[k][1], c.y);
        float z = fmaf(e.z, kExtentApplyFactor[k][2], c.z);
        f4 w = mat_mul_vec(l2w, x, y, z, 1.0f);
        p[k].x = w.x; p[k].y = w.y; p[k].z = w.z;
    }
    for (int i = 0; i < 6; i++) {
        int allBack = 1;
        for (int j = 0; j < 8; j++) {
            if (dot3f(planes[i].x, planes[i].y, planes[i].z, p[j].x, p[j].y, p[j].z) > -planes[i].w) { allBack = 0; break; }
        }
        if (allBack) return 1;
    }
    return 0;
}

/* base.hlsli:248-272 orthoFrustumCulling: true if culled */
static int ortho_frustum_culling(f3 c, f3 e, const chord_float4x4* localToClip)
{
    float mn[3] = {10.0f, 10.0f, 10.0f}, mx[3] = {-10.0f, -10.0f, -10.0f};
    for (int k = 0; k < 8; k++) {
        float x = fmaf(e.x, kExtentApplyFactor[k][0], c.x);
        float y = fmaf(e.y, kExtentApplyFactor[k][1], c.y);
        float z = fmaf(e.z, kExtentApplyFactor[k][2], c.z);
        f3 uvz = project_pos_to_uvz(x, y, z, localToClip);
        mn[0] = fminf(mn[0], uvz.x); mn[1] = fminf(mn[1], uvz.y); mn[2] = fminf(mn[2], uvz.z);
        mx[0] = fmaxf(mx[0], uvz.x); mx[1] = fmaxf(mx[1], uvz.y); mx[2] = fmaxf(mx[2], uvz.z);
    }
    return (mn[0] >= 1.0f) || (mn[1] >= 1.0f) || (mx[0] <= 0.0f) || (mx[1] <= 0.0f);
}

static inline int is_ortho(const chord_float4x4* M) { return M->m[15] == 1.0f; }   /* base.hlsli:241-244 */

/* base.hlsli:500-518 projectSphereToScreen; tanHalfFovy = tanf(0.5f*fovy) evaluated once per view on the host */
static float project_sphere_to_screen(f4 sphere, float height, float tanHalfFovy)
{
    float d2 = dot3f(sphere.x, sphere.y, sphere.z, sphere.x, sphere.y, sphere.z);
    float r = sphere.w, r2 = r * r;
    if (d2 <= r2) return -1.0f;
    return height * 0.5f / tanHalfFovy * r / sqrtf(d2 - r2);
}

/* nanite_shared.hlsli:15-49 */
static int is_meshlet_group_visible(float renderHeight, float tanHalfFovy, const chord_GPUObjectGLTFPrimitive* obj,
                                    const chord_float4x4* localToView, const chord_GPUGLTFMeshletGroup* g)
{
    const int bFinalLod = (g->parentError > CHORD_ERROR_RADIUS_ROOT);
    const int bFirstLod = (g->error < -0.5f);
    const float maxScale = obj->basicData.scaleExtractFromMatrix.w;
    if (!bFinalLod) {
        f4 v = mat_mul_vec(localToView, g->parentPosCenter.x, g->parentPosCenter.y, g->parentPosCenter.z, 1.0f);
        f4 s = { v.x, v.y, v.z, maxScale * g->parentError };
        float parentError = project_sphere_to_screen(s, renderHeight, tanHalfFovy);
        if (parentError > 0.0f && parentError <= CHORD_ERROR_PIXEL_THRESHOLD) return 0;
    }
    if (!bFirstLod) {
        f4 v = mat_mul_vec(localToView, g->clusterPosCenter.x, g->clusterPosCenter.y, g->clusterPosCenter.z, 1.0f);
        f4 s = { v.x, v.y, v.z, maxScale * g->error };
        float error = project_sphere_to_screen(s, renderHeight, tanHalfFovy);
        if (error < 0.0f || error > CHORD_ERROR_PIXEL_THRESHOLD) return 0;
    }
    return 1;
}

/* nanite_shared.hlsli:51-91 */
static int is_meshlet_visible(uint32_t switchFlags, const chord_float4 planes[6], const chord_GPUObjectGLTFPrimitive* obj,
                              const chord_float4x4* l2w, const chord_float4x4* localToClip,
                              const chord_GPUGLTFMeshlet* m, const chord_GLTFMaterialGPUData* mat)
{
    int visible = 1;
    if (visible && mat->bTwoSided == 0 && has_flag(switchFlags, CHORD_FLAG_CONE_CULL_BIT)) {
        f4 cam = mat_mul_vec(&obj->basicData.translatedWorldToLocal, 0.0f, 0.0f, 0.0f, 1.0f);
        float vx = m->coneApex.x - cam.x, vy = m->coneApex.y - cam.y, vz = m->coneApex.z - cam.z;
        float len = sqrtf(dot3f(vx, vy, vz, vx, vy, vz));
        vx = vx / len; vy = vy / len; vz = vz / len;
        if (dot3f(vx, vy, vz, m->coneAxis.x, m->coneAxis.y, m->coneAxis.z) >= m->coneCutOff) visible = 0;
    }
    f3 c = { 0.5f * (m->posMin.x + m->posMax.x), 0.5f * (m->posMin.y + m->posMax.y), 0.5f * (m->posMin.z + m->posMax.z) };
    f3 e = { m->posMax.x - c.x, m->posMax.y - c.y, m->posMax.z - c.z };
    if (visible && has_flag(switchFlags, CHORD_FLAG_FRUSTUM_CULL_BIT)) {
        if (is_ortho(localToClip)) visible = !ortho_frustum_culling(c, e, localToClip);
        else                        visible = !frustum_culling(planes, c, e, l2w);
    }
    return visible;
}

/* ------------------------------------------------------------------ K1 + K2 */
uint32_t oracle_instance_culling(const oracle_frame* f, chord_draw_cmd* out, uint32_t capacity)
{
    const float tanHalfFovy = tanf(0.5f * f->view->cameraFovy);
    uint32_t count = 0;
    for (uint32_t objectId = 0; objectId < f->objectCount; objectId++) {
        object_refs o = get_object(f, objectId);
        const chord_float4x4* l2w = &o.object->basicData.localToTranslatedWorld;
        chord_float4x4 localToClip; mat_mul(&f->iview->translatedWorldToClip, l2w, &localToClip);

        /* instance_culling.hlsl:67-99, object level */
        int visible = 1;
        f3 c = { (o.prim->posMin.x + o.prim->posMax.x) * 0.5f, (o.prim->posMin.y + o.prim->posMax.y) * 0.5f, (o.prim->posMin.z + o.prim->posMax.z) * 0.5f };
        f3 e = { o.prim->posMax.x - c.x, o.prim->posMax.y - c.y, o.prim->posMax.z - c.z };
        if (has_flag(f->switchFlags, CHORD_FLAG_FRUSTUM_CULL_BIT)) {
            if (is_ortho(&localToClip)) visible = !ortho_frustum_culling(c, e, &localToClip);
            else                         visible = !frustum_culling(f->iview->frustumPlanesRS, c, e, l2w);
        }
        if (!visible) continue;

        /* instance_culling.hlsl:133-208, per cluster group in index order */
        chord_float4x4 localToView; mat_mul(&f->view->translatedWorldToView, l2w, &localToView);
        for (uint32_t g = 0; g < o.prim->meshletGroupCount; g++) {
            const chord_GPUGLTFMeshletGroup* grp = &o.datas->groups[o.prim->meshletGroupOffset + g];
            if (!is_meshlet_group_visible(f->view->renderDimension.y, tanHalfFovy, o.object, &localToView, grp)) continue;
            for (uint32_t i = 0; i < grp->meshletCount; i++) {
                uint32_t idx = o.datas->groupIndices[i + grp->meshletOffset + o.prim->meshletGroupIndicesOffset];
                uint32_t meshletIndex = o.prim->meshletOffset + idx;
                const chord_GPUGLTFMeshlet* m = &o.datas->meshlets[meshletIndex];
                if (is_meshlet_visible(f->switchFlags, f->iview->frustumPlanesRS, o.object, l2w, &localToClip, m, o.material)) {
                    if (count < capacity) { out[count].objectId = objectId; out[count].meshletId = meshletIndex; out[count].slot = count; }
                    count++;
                }
            }
        }
    }
    return count;
}

/* ------------------------------------------------------------------ N4: cluster-group BVH traversal
 * The reference builds and uploads the BVH (nanite_builder.cpp:313-416) but never traverses it.  A node's
 * sphere bounds the PARENT-error spheres of every group below it, so when it projects to <= 1 px (with a 1e-4 safety
 * margin against rounding) every such group fails the "parent still too coarse" half of the LOD test and the subtree is
 * skipped.  Groups of visited nodes are only CANDIDATES; the exact per-group test stays the flat one, so the final list
 * is identical to oracle_instance_culling. bits[] has one bit per (visible-object, group) in canonical order. */
#define BVH_PRUNE_THRESHOLD (CHORD_ERROR_PIXEL_THRESHOLD * (1.0f - 1e-4f))

static void bvh_visit(const oracle_frame* f, const object_refs* o, const chord_float4x4* localToView, float tanHalfFovy,
                      uint32_t nodeIndex, uint32_t groupBase, uint32_t* bits)
{
    const chord_GPUBVHNode* n = &o->datas->bvhNodes[o->prim->bvhNodeOffset + nodeIndex];
    for (uint32_t k = 0; k < n->leafMeshletGroupCount; k++) {
        uint32_t b = groupBase + n->leafMeshletGroupOffset + k;
        bits[b >> 5] |= 1u << (b & 31u);
    }
    for (int c = 0; c < CHORD_BVH_LEVEL_NODE_COUNT; c++) {
        if (n->children[c] == CHORD_UNVALID_ID) continue;
        const chord_GPUBVHNode* ch = &o->datas->bvhNodes[o->prim->bvhNodeOffset + n->children[c]];
        f4 v = mat_mul_vec(localToView, ch->sphere.x, ch->sphere.y, ch->sphere.z, 1.0f);
        f4 s = { v.x, v.y, v.z, o->object->basicData.scaleExtractFromMatrix.w * ch->sphere.w };
        float e = project_sphere_to_screen(s, f->view->renderDimension.y, tanHalfFovy);
        if (e > 0.0f && e <= BVH_PRUNE_THRESHOLD) continue;      /* whole subtree: parent error below one pixel */
        bvh_visit(f, o, localToView, tanHalfFovy, n->children[c], groupBase, bits);
    }
}

/* returns the number of candidate groups; totalGroups receives the number of (visible-object, group) slots */
uint32_t oracle_bvh_candidates(const oracle_frame* f, uint32_t* bits, uint32_t bitWords, uint32_t* totalGroups)
{
    const float tanHalfFovy = tanf(0.5f * f->view->cameraFovy);
    memset(bits, 0, sizeof(uint32_t) * bitWords);
    uint32_t base = 0;
    for (uint32_t objectId = 0; objectId < f->objectCount; objectId++) {
        object_refs o = get_object(f, objectId);
        const chord_float4x4* l2w = &o.object->basicData.localToTranslatedWorld;
        chord_float4x4 localToClip; mat_mul(&f->iview->translatedWorldToClip, l2w, &localToClip);
        int visible = 1;
        f3 c = { (o.prim->posMin.x + o.prim->posMax.x) * 0.5f, (o.prim->posMin.y + o.prim->posMax.y) * 0.5f, (o.prim->posMin.z + o.prim->posMax.z) * 0.5f };
        f3 e = { o.prim->posMax.x - c.x, o.prim->posMax.y - c.y, o.prim->posMax.z - c.z };
        if (has_flag(f->switchFlags, CHORD_FLAG_FRUSTUM_CULL_BIT)) {
            if (is_ortho(&localToClip)) visible = !ortho_frustum_culling(c, e, &localToClip);
            else                         visible = !frustum_culling(f->iview->frustumPlanesRS, c, e, l2w);
        }
        if (!visible) continue;
        chord_float4x4 localToView; mat_mul(&f->view->translatedWorldToView, l2w, &localToView);
        if (o.datas->bvhNodes) bvh_visit(f, &o, &localToView, tanHalfFovy, 0, base, bits);
        base += o.prim->meshletGroupCount;
    }
    if (totalGroups) *totalGroups = base;
    uint32_t n = 0;
    for (uint32_t w = 0; w < bitWords; w++) n += (uint32_t)__builtin_popcount(bits[w]);
    return n;
}

/* instance culling through the BVH: identical output to oracle_instance_culling */
uint32_t oracle_instance_culling_bvh(const oracle_frame* f, chord_draw_cmd* out, uint32_t capacity, uint32_t* candidateGroups)
{
    uint32_t total = 0;
    for (uint32_t i = 0; i < f->objectCount; i++) total += get_object(f, i).prim->meshletGroupCount;
    const uint32_t words = (total + 31) / 32 + 1;
    uint32_t* bits = (uint32_t*)calloc(words, 4);
    uint32_t cand = oracle_bvh_candidates(f, bits, words, NULL);
    if (candidateGroups) *candidateGroups = cand;
    const float tanHalfFovy = tanf(0.5f * f->view->cameraFovy);
    uint32_t count = 0, base = 0;
    for (uint32_t objectId = 0; objectId < f->objectCount; objectId++) {
        object_refs o = get_object(f, objectId);
        const chord_float4x4* l2w = &o.object->basicData.localToTranslatedWorld;
        chord_float4x4 localToClip; mat_mul(&f->iview->translatedWorldToClip, l2w, &localToClip);
        int visible = 1;
        f3 c = { (o.prim->posMin.x + o.prim->posMax.x) * 0.5f, (o.prim->posMin.y + o.prim->posMax.y) * 0.5f, (o.prim->posMin.z + o.prim->posMax.z) * 0.5f };
        f3 e = { o.prim->posMax.x - c.x, o.prim->posMax.y - c.y, o.prim->posMax.z - c.z };
        if (has_flag(f->switchFlags, CHORD_FLAG_FRUSTUM_CULL_BIT)) {
            if (is_ortho(&localToClip)) visible = !ortho_frustum_culling(c, e, &localToClip);
            else                         visible = !frustum_culling(f->iview->frustumPlanesRS, c, e, l2w);
        }
        if (!visible) continue;
        chord_float4x4 localToView; mat_mul(&f->view->translatedWorldToView, l2w, &localToView);
        for (uint32_t g = 0; g < o.prim->meshletGroupCount; g++) {
            const uint32_t b = base + g;
            if (!(bits[b >> 5] & (1u << (b & 31u)))) continue;
            const chord_GPUGLTFMeshletGroup* grp = &o.datas->groups[o.prim->meshletGroupOffset + g];
            if (!is_meshlet_group_visible(f->view->renderDimension.y, tanHalfFovy, o.object, &localToView, grp)) continue;
            for (uint32_t i = 0; i < grp->meshletCount; i++) {
                uint32_t idx = o.datas->groupIndices[i + grp->meshletOffset + o.prim->meshletGroupIndicesOffset];
                uint32_t meshletIndex = o.prim->meshletOffset + idx;
                const chord_GPUGLTFMeshlet* m = &o.datas->meshlets[meshletIndex];
                if (is_meshlet_visible(f->switchFlags, f->iview->frustumPlanesRS, o.object, l2w, &localToClip, m, o.material)) {
                    if (count < capacity) { out[count].objectId = objectId; out[count].meshletId = meshletIndex; out[count].slot = count; }
                    count++;
                }
            }
        }
        base += o.prim->meshletGroupCount;
    }
    free(bits);
    return count;
}

/* ------------------------------------------------------------------ K3 */
static inline int firstbithigh_i(int32_t v)
{
    if (v <= 0) return -1;     /* HLSL firstbithigh(0) == -1; inputs here are never negative */
    int b = 0; uint32_t u = (uint32_t)v;
    while (u >>= 1) b++;
    return b;
}

void oracle_hzb_culling(const oracle_frame* f, int firstStage,
                        const uint16_t* hzbMin, const uint64_t* hzbMipOffsets, uint32_t mip0Width, uint32_t mip0Height, uint32_t mipCount,
                        const chord_draw_cmd* in, uint32_t inCount,
                        chord_draw_cmd* outVisible, uint32_t* outVisibleCount,
                        chord_draw_cmd* outRejected, uint32_t* outRejectedCount)
{
    (void)mip0Height; (void)mipCount;
    uint32_t nv = 0, nr = 0;
    const chord_PerframeCameraView* view = f->view;
    for (uint32_t t = 0; t < inCount; t++) {
        const chord_draw_cmd cmd = in[t];
        object_refs o = get_object(f, cmd.objectId);
        const chord_GPUGLTFMeshlet* m = &o.datas->meshlets[cmd.meshletId];
        f3 c = { (m->posMin.x + m->posMax.x) * 0.5f, (m->posMin.y + m->posMax.y) * 0.5f, (m->posMin.z + m->posMax.z) * 0.5f };
        f3 e = { m->posMax.x - c.x, m->posMax.y - c.y, m->posMax.z - c.z };
        int visible = 1;
        if (has_flag(f->switchFlags, CHORD_FLAG_HZB_CULL_BIT)) {
            float mx[3] = {-10.0f, -10.0f, -10.0f}, mn[3] = {10.0f, 10.0f, 10.0f};
            chord_float4x4 mvp;
            if (firstStage) mat_mul(&view->translatedWorldToClipLastFrame, &o.object->basicData.localToTranslatedWorldLastFrame, &mvp);
            else            mat_mul(&view->translatedWorldToClip, &o.object->basicData.localToTranslatedWorld, &mvp);
            for (int i = 0; i < 8; i++) {
                float x = fmaf(e.x, kExtentApplyFactor[i][0], c.x);
                float y = fmaf(e.y, kExtentApplyFactor[i][1], c.y);
                float z = fmaf(e.z, kExtentApplyFactor[i][2], c.z);
                f3 uvz = project_pos_to_uvz(x, y, z, &mvp);
                mn[0] = fminf(mn[0], uvz.x); mn[1] = fminf(mn[1], uvz.y); mn[2] = fminf(mn[2], uvz.z);
                mx[0] = fmaxf(mx[0], uvz.x); mx[1] = fmaxf(mx[1], uvz.y); mx[2] = fmaxf(mx[2], uvz.z);
            }
            const int zInRange = (mx[2] < 1.0f) && (mn[2] > 0.0f);
            if (zInRange) {
                if (mn[0] >= 1.0f || mn[1] >= 1.0f || mx[0] <= 0.0f || mx[1] <= 0.0f) visible = 0;
            }
            if (visible && zInRange) {
                /* saturate */
                float u0 = fminf(fmaxf(mn[0], 0.0f), 1.0f), v0 = fminf(fmaxf(mn[1], 0.0f), 1.0f);
                float u1 = fminf(fmaxf(mx[0], 0.0f), 1.0f), v1 = fminf(fmaxf(mx[1], 0.0f), 1.0f);
                const float W = view->renderDimension.x, H = view->renderDimension.y;
                int32_t rx0 = (int32_t)fmaf(u0, W, 0.5f), ry0 = (int32_t)fmaf(v0, H, 0.5f);
                int32_t rx1 = (int32_t)fmaf(u1, W, -0.5f), ry1 = (int32_t)fmaf(v1, H, -0.5f);
                if (rx0 < 0) rx0 = 0; if (ry0 < 0) ry0 = 0;
                /* min(renderDimension.xy - 1, pixelRect.zw): float min then int conversion */
                float fx1 = fminf(W - 1.0f, (float)rx1), fy1 = fminf(H - 1.0f, (float)ry1);
                rx1 = (int32_t)fx1; ry1 = (int32_t)fy1;
                if (rx1 < rx0 || ry1 < ry0) {
                    visible = 0;
                } else {
                    int32_t hx0 = rx0 >> 1, hy0 = ry0 >> 1, hx1 = rx1 >> 1, hy1 = ry1 >> 1;
                    int mlx = firstbithigh_i(hx1 - hx0), mly = firstbithigh_i(hy1 - hy0);
                    int mipLevel = (mlx > mly ? mlx : mly) - 1;
                    if (mipLevel < 0) mipLevel = 0;
                    if (((hx1 >> mipLevel) - (hx0 >> mipLevel) >= 4) || ((hy1 >> mipLevel) - (hy0 >> mipLevel) >= 4)) mipLevel += 1;
                    int32_t cx0 = hx0 >> mipLevel, cy0 = hy0 >> mipLevel, cx1 = hx1 >> mipLevel, cy1 = hy1 >> mipLevel;
                    float zMin = 10.0f;
                    const uint32_t pitch = (mip0Width >> mipLevel) ? (mip0Width >> mipLevel) : 1u;   /* mips narrower than one texel are stored 1 wide (Texture2D.Load reads the right row) */
                    const uint16_t* mip = hzbMin + hzbMipOffsets[mipLevel];
                    for (int x = 0; x < 4; x++)
                        for (int y = 0; y < 4; y++) {
                            int32_t sx = cx0 + x; if (sx > cx1) sx = cx1;
                            int32_t sy = cy0 + y; if (sy > cy1) sy = cy1;
                            zMin = fminf(zMin, oracle_f16_to_f32(mip[(uint64_t)sy * pitch + (uint32_t)sx]));
                        }
                    if (zMin > mx[2]) visible = 0;
                }
            }
        }
        if (visible) outVisible[nv++] = cmd;
        else if (firstStage && outRejected) outRejected[nr++] = cmd;
    }
    *outVisibleCount = nv;
    if (outRejectedCount) *outRejectedCount = nr;
}

/* ------------------------------------------------------------------ hzbCullingGeneric (hzb_culling_generic.hlsl:37-172) */
static inline double storage_double(chord_uint2 u) { uint64_t b = (uint64_t)u.x | ((uint64_t)u.y << 32); double d; memcpy(&d, &b, 8); return d; }

void oracle_hzb_culling_generic(const oracle_frame* f, const chord_InstanceCullingViewInfo* hzbView, float extentScale,
                                int bObjectUseLastFrameProject,
                                const uint16_t* hzbMin, const uint64_t* hzbMipOffsets, uint32_t mip0Width,
                                const chord_draw_cmd* in, uint32_t inCount, chord_draw_cmd* outVisible, uint32_t* outVisibleCount)
{
    uint32_t nv = 0;
    /* :78-81 relativeOffset = float3(perView.cameraWorldPos - instanceView.cameraWorldPos), per component */
    const float rel[3] = {
        (float)(storage_double(f->view->cameraWorldPos.x) - storage_double(hzbView->cameraWorldPos.x)),
        (float)(storage_double(f->view->cameraWorldPos.y) - storage_double(hzbView->cameraWorldPos.y)),
        (float)(storage_double(f->view->cameraWorldPos.z) - storage_double(hzbView->cameraWorldPos.z)) };
    for (uint32_t t = 0; t < inCount; t++) {
        const chord_draw_cmd cmd = in[t];
        object_refs o = get_object(f, cmd.objectId);
        const chord_GPUGLTFMeshlet* m = &o.datas->meshlets[cmd.meshletId];
        f3 c = { (m->posMin.x + m->posMax.x) * 0.5f, (m->posMin.y + m->posMax.y) * 0.5f, (m->posMin.z + m->posMax.z) * 0.5f };
        f3 e = { m->posMax.x - c.x, m->posMax.y - c.y, m->posMax.z - c.z };
        int visible = 1;
        if (has_flag(f->switchFlags, CHORD_FLAG_HZB_CULL_BIT)) {
            float mx[3] = {-10.0f, -10.0f, -10.0f}, mn[3] = {10.0f, 10.0f, 10.0f};
            chord_float4x4 l2w = bObjectUseLastFrameProject ? o.object->basicData.localToTranslatedWorldLastFrame : o.object->basicData.localToTranslatedWorld;
            l2w.m[12] += rel[0]; l2w.m[13] += rel[1]; l2w.m[14] += rel[2];       /* [0][3], [1][3], [2][3] (column-major storage) */
            chord_float4x4 mvp; mat_mul(&hzbView->translatedWorldToClip, &l2w, &mvp);
            int can = 1;
            for (int i = 0; i < 8; i++) {
                float x = fmaf(e.x, kExtentApplyFactor[i][0], c.x);
                float y = fmaf(e.y, kExtentApplyFactor[i][1], c.y);
                float z = fmaf(e.z, kExtentApplyFactor[i][2], c.z);
                f3 uvz = project_pos_to_uvz(x, y, z, &mvp);
                mn[0] = fminf(mn[0], uvz.x); mn[1] = fminf(mn[1], uvz.y); mn[2] = fminf(mn[2], uvz.z);
                mx[0] = fmaxf(mx[0], uvz.x); mx[1] = fmaxf(mx[1], uvz.y); mx[2] = fmaxf(mx[2], uvz.z);
                can = can && (uvz.x < 1.0f && uvz.y < 1.0f && uvz.z < 1.0f) && (uvz.x > 0.0f && uvz.y > 0.0f && uvz.z > 0.0f);
            }
            if (can) {
                const float W = hzbView->renderDimension.x, H = hzbView->renderDimension.y;
                /* int4(uvRect * dim.xyxy + extentScale * (-.5, -.5, .5, .5)): mad per component, truncation */
                const float lo = extentScale * -0.5f, hi = extentScale * 0.5f;      /* exact */
                int32_t rx0 = (int32_t)fmaf(mn[0], W, lo), ry0 = (int32_t)fmaf(mn[1], H, lo);
                int32_t rx1 = (int32_t)fmaf(mx[0], W, hi), ry1 = (int32_t)fmaf(mx[1], H, hi);
                if (rx0 < 0) rx0 = 0; if (ry0 < 0) ry0 = 0;
                float fx1 = fminf(W - 1.0f, (float)rx1), fy1 = fminf(H - 1.0f, (float)ry1);
                rx1 = (int32_t)fx1; ry1 = (int32_t)fy1;
                if (rx1 < rx0 || ry1 < ry0) {
                    visible = 0;
                } else {
                    int32_t hx0 = rx0 >> 1, hy0 = ry0 >> 1, hx1 = rx1 >> 1, hy1 = ry1 >> 1;
                    int mlx = firstbithigh_i(hx1 - hx0), mly = firstbithigh_i(hy1 - hy0);
                    int mipLevel = mlx > mly ? mlx : mly;
                    if (mipLevel < 0) mipLevel = 0;
                    if (((hx1 >> mipLevel) - (hx0 >> mipLevel) >= 2) || ((hy1 >> mipLevel) - (hy0 >> mipLevel) >= 2)) mipLevel += 1;
                    int32_t cx0 = hx0 >> mipLevel, cy0 = hy0 >> mipLevel, cx1 = hx1 >> mipLevel, cy1 = hy1 >> mipLevel;
                    float zMin = 10.0f;
                    const uint32_t pitch = (mip0Width >> mipLevel) ? (mip0Width >> mipLevel) : 1u;   /* mips narrower than one texel are stored 1 wide (Texture2D.Load reads the right row) */
                    const uint16_t* mip = hzbMin + hzbMipOffsets[mipLevel];
                    for (int x = 0; x < 2; x++)
                        for (int y = 0; y < 2; y++) {
                            int32_t sx = cx0 + x; if (sx > cx1) sx = cx1;
                            int32_t sy = cy0 + y; if (sy > cy1) sy = cy1;
                            zMin = fminf(zMin, oracle_f16_to_f32(mip[(uint64_t)sy * pitch + (uint32_t)sx]));
                        }
                    if (zMin > mx[2]) visible = 0;
                }
            }
        }
        if (visible) outVisible[nv++] = cmd;
    }
    *outVisibleCount = nv;
}

/* ------------------------------------------------------------------ K7 masked bucket: the specified sampler
 * mesh_raster.hlsl:198-204: sampleColor = baseColorTexture.Sample(baseColorSampler, uv); clip(sampleColor.w * factor - cutoff).
 * Sample() with implicit derivatives is implementation defined on a GPU; pinned here as: trilinear over RGBA8 unorm,
 * rho^2 = max(|d(uv*dim0)/dx|^2, |d(uv*dim0)/dy|^2); rho <= 1 -> level 0; level = floor(log2 rho) read from the exponent,
 * fraction = mantissa - 1 (piecewise-linear log2); texel centre convention x = u*w - 0.5; all arithmetic fp32, one
 * rounding per written operation (fmaf where written). */
static inline int wrap_texel(int x, int size, uint32_t mode)
{
    if (mode == CHORD_ADDRESS_CLAMP_TO_EDGE) return x < 0 ? 0 : (x >= size ? size - 1 : x);
    if (mode == CHORD_ADDRESS_MIRRORED_REPEAT) {
        const int p = 2 * size; int m = x % p; if (m < 0) m += p;
        return m < size ? m : p - 1 - m;
    }
    int m = x % size; if (m < 0) m += size;
    return m;
}

static float alpha_bilinear(const uint8_t* rgba8, int w, int h, uint32_t modeU, uint32_t modeV, float u, float v)
{
    const float x = fmaf(u, (float)w, -0.5f), y = fmaf(v, (float)h, -0.5f);
    /* clamp before the integer conversion: |x| <= 2^24 keeps floor exact and the conversion defined; NaN -> -2^24 */
    const float fx0 = fminf(fmaxf(floorf(x), -16777216.0f), 16777216.0f), fy0 = fminf(fmaxf(floorf(y), -16777216.0f), 16777216.0f);
    float fx = x - fx0, fy = y - fy0;
    if (!(fx >= 0.0f && fx <= 1.0f)) fx = 0.0f;        /* out-of-range / NaN coordinates: nearest of the clamped texel */
    if (!(fy >= 0.0f && fy <= 1.0f)) fy = 0.0f;
    const int x0 = wrap_texel((int)fx0, w, modeU), x1 = wrap_texel((int)fx0 + 1, w, modeU);
    const int y0 = wrap_texel((int)fy0, h, modeV), y1 = wrap_texel((int)fy0 + 1, h, modeV);
    const float a00 = (float)rgba8[((size_t)y0 * w + x0) * 4 + 3] / 255.0f, a10 = (float)rgba8[((size_t)y0 * w + x1) * 4 + 3] / 255.0f;
    const float a01 = (float)rgba8[((size_t)y1 * w + x0) * 4 + 3] / 255.0f, a11 = (float)rgba8[((size_t)y1 * w + x1) * 4 + 3] / 255.0f;
    const float top = fmaf(fx, a10 - a00, a00), bot = fmaf(fx, a11 - a01, a01);
    return fmaf(fy, bot - top, top);
}

static inline const uint8_t* mip_texels(const chord_texture_desc* t, uint32_t level, int* w, int* h)
{
    const uint8_t* p = t->rgba8;
    uint32_t lw = t->width, lh = t->height;
    for (uint32_t l = 0; l < level; l++) { p += (size_t)lw * lh * 4; lw = lw > 1 ? lw >> 1 : 1; lh = lh > 1 ? lh >> 1 : 1; }
    *w = (int)lw; *h = (int)lh;
    return p;
}

float oracle_sample_alpha(const chord_texture_desc* tex, const chord_sampler_desc* smp, float u, float v,
                          float dudx, float dvdx, float dudy, float dvdy)
{
    if (!tex || !tex->rgba8 || tex->width == 0 || tex->height == 0 || tex->mipCount == 0) return 1.0f;   /* builtin white */
    const uint32_t modeU = smp ? smp->addressU : CHORD_ADDRESS_REPEAT, modeV = smp ? smp->addressV : CHORD_ADDRESS_REPEAT;
    const float W0 = (float)tex->width, H0 = (float)tex->height;
    const float ax = dudx * W0, bx = dvdx * H0, ay = dudy * W0, by = dvdy * H0;
    const float rx2 = fmaf(bx, bx, ax * ax), ry2 = fmaf(by, by, ay * ay);
    const float r2 = fmaxf(rx2, ry2);
    uint32_t level = 0; float frac = 0.0f;
    if (r2 > 1.0f) {
        const float rho = sqrtf(r2);
        const uint32_t bits = f2u(rho);
        const uint32_t e = (bits >> 23) & 0xFFu;                      /* rho > 1 and not NaN: e >= 127 */
        if (e - 127u >= tex->mipCount - 1u) { level = tex->mipCount - 1u; frac = 0.0f; }
        else { level = e - 127u; frac = u2f((bits & 0x007FFFFFu) | 0x3F800000u) - 1.0f; }
    }
    int w, h;
    const uint8_t* t0 = mip_texels(tex, level, &w, &h);
    float a = alpha_bilinear(t0, w, h, modeU, modeV, u, v);
    if (frac > 0.0f) {
        const uint8_t* t1 = mip_texels(tex, level + 1u, &w, &h);
        const float a1 = alpha_bilinear(t1, w, h, modeU, modeV, u, v);
        a = fmaf(frac, a1 - a, a);
    }
    return a;
}

/* ---- K11 material textures: SampleGrad (material.hlsli:66-79, :103-107, :126-130) with the SAME pinned sampler, on all four
 * channels and with the explicit gradients the resolve computes analytically (base.hlsli:457-498).  A texture flagged
 * CHORD_TEXTURE_SRGB decodes r, g, b texel-wise to linear before filtering (what a *_SRGB image format does): the decode is a
 * 256-entry table built once with the IEC 61966-2-1 formula in double precision (the library builds the same table with the
 * same host libm, so both sides hold identical floats). */
static float g_srgb_lut[256]; static int g_srgb_lut_ready = 0;
static const float* srgb_lut(void)
{
    if (!g_srgb_lut_ready) {
        for (int i = 0; i < 256; i++) {
            const double c = (double)i / 255.0;
            g_srgb_lut[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
        }
        __atomic_store_n(&g_srgb_lut_ready, 1, __ATOMIC_RELEASE);
    }
    return g_srgb_lut;
}

static void rgba_bilinear(const uint8_t* rgba8, int w, int h, uint32_t modeU, uint32_t modeV, float u, float v, const float* lut, float out[4])
{
    const float x = fmaf(u, (float)w, -0.5f), y = fmaf(v, (float)h, -0.5f);
    const float fx0 = fminf(fmaxf(floorf(x), -16777216.0f), 16777216.0f), fy0 = fminf(fmaxf(floorf(y), -16777216.0f), 16777216.0f);
    float fx = x - fx0, fy = y - fy0;
    if (!(fx >= 0.0f && fx <= 1.0f)) fx = 0.0f;
    if (!(fy >= 0.0f && fy <= 1.0f)) fy = 0.0f;
    const int x0 = wrap_texel((int)fx0, w, modeU), x1 = wrap_texel((int)fx0 + 1, w, modeU);
    const int y0 = wrap_texel((int)fy0, h, modeV), y1 = wrap_texel((int)fy0 + 1, h, modeV);
    const uint8_t* t00 = rgba8 + ((size_t)y0 * w + x0) * 4; const uint8_t* t10 = rgba8 + ((size_t)y0 * w + x1) * 4;
    const uint8_t* t01 = rgba8 + ((size_t)y1 * w + x0) * 4; const uint8_t* t11 = rgba8 + ((size_t)y1 * w + x1) * 4;
    for (int c = 0; c < 4; c++) {
        const int dec = lut != NULL && c < 3;
        const float a00 = dec ? lut[t00[c]] : (float)t00[c] / 255.0f, a10 = dec ? lut[t10[c]] : (float)t10[c] / 255.0f;
        const float a01 = dec ? lut[t01[c]] : (float)t01[c] / 255.0f, a11 = dec ? lut[t11[c]] : (float)t11[c] / 255.0f;
        const float top = fmaf(fx, a10 - a00, a00), bot = fmaf(fx, a11 - a01, a01);
        out[c] = fmaf(fy, bot - top, top);
    }
}

/* returns 0 (and builtin white) when the id does not name an uploaded texture */
int oracle_sample_rgba(const chord_texture_desc* tex, const chord_sampler_desc* smp, float u, float v,
                       float dudx, float dvdx, float dudy, float dvdy, float out[4])
{
    out[0] = out[1] = out[2] = out[3] = 1.0f;
    if (!tex || !tex->rgba8 || tex->width == 0 || tex->height == 0 || tex->mipCount == 0) return 0;     /* builtin white */
    const uint32_t modeU = smp ? smp->addressU : CHORD_ADDRESS_REPEAT, modeV = smp ? smp->addressV : CHORD_ADDRESS_REPEAT;
    const float* lut = (tex->flags & CHORD_TEXTURE_SRGB) ? srgb_lut() : NULL;
    const float W0 = (float)tex->width, H0 = (float)tex->height;
    const float ax = dudx * W0, bx = dvdx * H0, ay = dudy * W0, by = dvdy * H0;
    const float rx2 = fmaf(bx, bx, ax * ax), ry2 = fmaf(by, by, ay * ay);
    const float r2 = fmaxf(rx2, ry2);
    uint32_t level = 0; float frac = 0.0f;
    if (r2 > 1.0f) {
        const float rho = sqrtf(r2);
        const uint32_t bits = f2u(rho);
        const uint32_t e = (bits >> 23) & 0xFFu;
        if (e - 127u >= tex->mipCount - 1u) { level = tex->mipCount - 1u; frac = 0.0f; }
        else { level = e - 127u; frac = u2f((bits & 0x007FFFFFu) | 0x3F800000u) - 1.0f; }
    }
    int w, h;
    const uint8_t* t0 = mip_texels(tex, level, &w, &h);
    rgba_bilinear(t0, w, h, modeU, modeV, u, v, lut, out);
    if (frac > 0.0f) {
        float o1[4];
        const uint8_t* t1 = mip_texels(tex, level + 1u, &w, &h);
        rgba_bilinear(t1, w, h, modeU, modeV, u, v, lut, o1);
        for (int c = 0; c < 4; c++) out[c] = fmaf(frac, o1[c] - out[c], out[c]);
    }
    return 1;
}

/* per-triangle attributes of the masked bucket: 1/w and uv0 per (possibly clipped) vertex + the material's test */
typedef struct {
    const chord_texture_desc* tex; const chord_sampler_desc* smp;
    float alphaFactor, alphaCutOff;
    float rw[3], uv[3][2];
    uint64_t* killed;
} masked_attrs;

/* perspective-correct uv from the three oriented edge values (E[k] = weight numerator of vertex (k+2)%3) */
static inline void masked_uv(const masked_attrs* m, const int64_t E[3], float invA, float* u, float* v)
{
    const float b0 = (float)E[1] * invA, b1 = (float)E[2] * invA, b2 = (float)E[0] * invA;
    const float q0 = b0 * m->rw[0], q1 = b1 * m->rw[1], q2 = b2 * m->rw[2];
    const float S = (q0 + q1) + q2;
    *u = fmaf(q2, m->uv[2][0], fmaf(q1, m->uv[1][0], q0 * m->uv[0][0])) / S;
    *v = fmaf(q2, m->uv[2][1], fmaf(q1, m->uv[1][1], q0 * m->uv[0][1])) / S;
}

/* ------------------------------------------------------------------ K6' fixed-function rasteriser (specified) */
typedef struct {
    uint32_t width, height;
    chord_rect tile;
    uint64_t* vis64;
    int depthClamp;             /* CHORD_PASS_DEPTH_CLAMP: no near / far clipping (orthographic views), z clamped to [0, 1] */
} raster_target;

static inline void vis_atomic_max(uint64_t* p, uint64_t key)
{
    uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < key && !__atomic_compare_exchange_n(p, &old, key, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
}

static inline int64_t floor_div256(int64_t a) { return a >> 8; }                 /* arithmetic shift == floor */
static inline int64_t ceil_div256(int64_t a) { return -((-a) >> 8); }

/* screen-space triangle: sx, sy in pixels (y down), sz = z/w */
static void raster_screen_triangle(const raster_target* rt, const float sx[3], const float sy[3], const float sz[3],
                                   uint32_t visId, int twoSided, const masked_attrs* masked)
{
    int64_t X[3], Y[3];
    for (int i = 0; i < 3; i++) {
        X[i] = (int64_t)lrintf(sx[i] * 256.0f);     /* round-half-even to 1/256 px (subpixelPrecisionBits 8, base.h:443) */
        Y[i] = (int64_t)lrintf(sy[i] * 256.0f);
    }
    int64_t A = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
    if (A == 0) return;
    if (!twoSided && A > 0) return;                  /* CCW front (helper.h:400) + flipped viewport: front <=> A < 0; cull BACK */
    const int64_t s = (A > 0) ? 1 : -1;
    const int64_t absA = (A > 0) ? A : -A;

    int64_t minX = X[0], maxX = X[0], minY = Y[0], maxY = Y[0];
    for (int i = 1; i < 3; i++) {
        if (X[i] < minX) minX = X[i]; if (X[i] > maxX) maxX = X[i];
        if (Y[i] < minY) minY = Y[i]; if (Y[i] > maxY) maxY = Y[i];
    }
    int64_t px0 = ceil_div256(minX - 128), px1 = floor_div256(maxX - 128);
    int64_t py0 = ceil_div256(minY - 128), py1 = floor_div256(maxY - 128);
    if (px0 < (int64_t)rt->tile.x0) px0 = rt->tile.x0;
    if (py0 < (int64_t)rt->tile.y0) py0 = rt->tile.y0;
    if (px1 > (int64_t)rt->tile.x1 - 1) px1 = (int64_t)rt->tile.x1 - 1;
    if (py1 > (int64_t)rt->tile.y1 - 1) py1 = (int64_t)rt->tile.y1 - 1;
    if (px0 > px1 || py0 > py1) return;

    /* edge i: v_i -> v_(i+1); weight of the opposite vertex (i+2)%3 */
    int64_t dx[3], dy[3]; int tl[3];
    for (int i = 0; i < 3; i++) {
        int j = (i + 1) % 3;
        dx[i] = X[j] - X[i]; dy[i] = Y[j] - Y[i];
        int64_t odx = s * dx[i], ody = s * dy[i];    /* direction after orienting clockwise-in-y-down */
        tl[i] = (ody == 0 && odx > 0) || (ody < 0); /* top edge | left edge */
    }
    const float invA = 1.0f / (float)absA;
    const float dz1 = sz[1] - sz[0], dz2 = sz[2] - sz[0];

    for (int64_t py = py0; py <= py1; py++) {
        const int64_t SY = py * 256 + 128;
        for (int64_t px = px0; px <= px1; px++) {
            const int64_t SX = px * 256 + 128;
            int64_t E[3]; int inside = 1;
            for (int i = 0; i < 3; i++) {
                E[i] = s * (dx[i] * (SY - Y[i]) - dy[i] * (SX - X[i]));
                if (!(E[i] > 0 || (E[i] == 0 && tl[i]))) { inside = 0; break; }
            }
            if (!inside) continue;
            if (masked) {
                /* K7 (mesh_raster.hlsl:198-204): uv here and at the +x / +y neighbour samples (forward differences
                 * stand in for the quad derivatives), specified sampler, clip() */
                int64_t Ex[3], Ey[3];
                for (int i = 0; i < 3; i++) {
                    Ex[i] = s * (dx[i] * (SY - Y[i]) - dy[i] * (SX + 256 - X[i]));
                    Ey[i] = s * (dx[i] * (SY + 256 - Y[i]) - dy[i] * (SX - X[i]));
                }
                float u, v, ux, vx, uy, vy;
                masked_uv(masked, E, invA, &u, &v);
                masked_uv(masked, Ex, invA, &ux, &vx);
                masked_uv(masked, Ey, invA, &uy, &vy);
                const float a = oracle_sample_alpha(masked->tex, masked->smp, u, v, ux - u, vx - v, uy - u, vy - v);
                if (fmaf(a, masked->alphaFactor, -masked->alphaCutOff) < 0.0f) {
                    if (masked->killed) __atomic_fetch_add(masked->killed, 1, __ATOMIC_RELAXED);
                    continue;
                }
            }
            /* weights: w1 = E of edge 2 (v2->v0), w2 = E of edge 0 (v0->v1) */
            float b1 = (float)E[2] * invA;
            float b2 = (float)E[0] * invA;
            float z = fmaf(b2, dz2, fmaf(b1, dz1, sz[0]));
            z = fminf(fmaxf(z, 0.0f), 1.0f);
            uint64_t key = ((uint64_t)f2u(z) << 32) | (uint64_t)visId;
            vis_atomic_max(&rt->vis64[(uint64_t)py * rt->width + (uint64_t)px], key);
        }
    }
}

#define GUARD_BAND 8.0f
#define SCREEN_LIMIT 32768.0f

static inline float clip_plane_dist(const float v[4], int k)
{
    switch (k) {
    case 0: return v[3] - v[2];
    case 1: return v[2];
    case 2: return GUARD_BAND * v[3] - v[0];
    case 3: return GUARD_BAND * v[3] + v[0];
    case 4: return GUARD_BAND * v[3] - v[1];
    default: return GUARD_BAND * v[3] + v[1];
    }
}

static int to_screen(const raster_target* rt, const float c[4], float* sx, float* sy, float* sz)
{
    const float halfW = 0.5f * (float)rt->width, halfH = 0.5f * (float)rt->height;
    *sx = fmaf(c[0] / c[3], halfW, halfW);
    *sy = fmaf(c[1] / c[3], -halfH, halfH);       /* flipped viewport, helper.h:304-323 */
    *sz = c[2] / c[3];
    return (fabsf(*sx) <= SCREEN_LIMIT) && (fabsf(*sy) <= SCREEN_LIMIT);   /* false for NaN / inf */
}

/* returns 1 if the clip path was taken.  uv (masked bucket only, else NULL) rides along as two more coordinates */
static int raster_clip_triangle(const raster_target* rt, const float c[3][4], uint32_t visId, int twoSided,
                                const masked_attrs* masked)
{
    int needClip = 0;
    float sx[3], sy[3], sz[3];
    for (int i = 0; i < 3; i++)
        if (rt->depthClamp ? !(c[i][3] > 0.0f) : !(c[i][2] >= 0.0f && c[i][2] <= c[i][3])) needClip = 1;
    if (!needClip)
        for (int i = 0; i < 3; i++)
            if (!to_screen(rt, c[i], &sx[i], &sy[i], &sz[i])) needClip = 1;
    if (!needClip) {
        if (masked) {
            masked_attrs m = *masked;
            for (int i = 0; i < 3; i++) m.rw[i] = 1.0f / c[i][3];
            raster_screen_triangle(rt, sx, sy, sz, visId, twoSided, &m);
        } else {
            raster_screen_triangle(rt, sx, sy, sz, visId, twoSided, NULL);
        }
        return 0;
    }

    /* Sutherland-Hodgman in clip space; new vertices are always interpolated from the inside
     * vertex towards the outside one so that a shared edge clips identically in both triangles. */
    float poly[2][12][6]; int n = 3, cur = 0;
    const int nc = masked ? 6 : 4;
    for (int i = 0; i < 3; i++) {
        memcpy(poly[0][i], c[i], 16);
        poly[0][i][4] = masked ? masked->uv[i][0] : 0.0f; poly[0][i][5] = masked ? masked->uv[i][1] : 0.0f;
    }
    for (int k = rt->depthClamp ? 2 : 0; k < 6 && n >= 3; k++) {       /* depth clamp: the z planes are not clipped against */
        int m = 0;
        for (int i = 0; i < n; i++) {
            const float* a = poly[cur][i];
            const float* b = poly[cur][(i + 1) % n];
            float da = clip_plane_dist(a, k), db = clip_plane_dist(b, k);
            int ina = (da >= 0.0f), inb = (db >= 0.0f);
            if (ina) { memcpy(poly[cur ^ 1][m++], a, 24); }
            if (ina != inb) {
                const float* pin = ina ? a : b; const float* pout = ina ? b : a;
                float din = ina ? da : db, dout = ina ? db : da;
                float t = din / (din - dout);
                for (int j = 0; j < nc; j++) poly[cur ^ 1][m][j] = fmaf(t, pout[j] - pin[j], pin[j]);
                m++;
            }
        }
        n = m; cur ^= 1;
    }
    if (n < 3) return 1;
    float px[12], py[12], pz[12]; int ok[12];
    for (int i = 0; i < n; i++) ok[i] = to_screen(rt, poly[cur][i], &px[i], &py[i], &pz[i]);
    for (int i = 1; i + 1 < n; i++) {
        if (!(ok[0] && ok[i] && ok[i + 1])) continue;
        float tx[3] = { px[0], px[i], px[i + 1] }, ty[3] = { py[0], py[i], py[i + 1] }, tz[3] = { pz[0], pz[i], pz[i + 1] };
        if (masked) {
            masked_attrs m = *masked;
            const int id[3] = { 0, i, i + 1 };
            for (int k = 0; k < 3; k++) {
                m.rw[k] = 1.0f / poly[cur][id[k]][3];
                m.uv[k][0] = poly[cur][id[k]][4]; m.uv[k][1] = poly[cur][id[k]][5];
            }
            raster_screen_triangle(rt, tx, ty, tz, visId, twoSided, &m);
        } else {
            raster_screen_triangle(rt, tx, ty, tz, visId, twoSided, NULL);
        }
    }
    return 1;
}

void oracle_raster_clip_triangle(uint32_t width, uint32_t height, const chord_rect* tile,
                                 const float clip[12], uint32_t visId, int twoSided, uint64_t* vis64)
{
    raster_target rt = { width, height, {0, 0, width, height}, vis64, 0 };
    if (tile) rt.tile = *tile;
    float c[3][4]; memcpy(c, clip, sizeof(c));
    raster_clip_triangle(&rt, c, visId, twoSided, NULL);
}

/* 3x3 determinant of rows (x,y,w): mesh_raster.hlsl:147 */
static inline float minor2(float a, float b, float c, float d) { return fmaf(a, b, -(c * d)); }
static inline float det3(const float r0[3], const float r1[3], const float r2[3])
{
    float m0 = minor2(r1[1], r2[2], r1[2], r2[1]);
    float m1 = minor2(r1[0], r2[2], r1[2], r2[0]);
    float m2 = minor2(r1[0], r2[1], r1[1], r2[0]);
    return fmaf(r0[2], m2, fmaf(-r0[1], m1, r0[0] * m0));
}

/* ------------------------------------------------------------------ K4 + K6 + K7 */
void oracle_render_mesh(const oracle_frame* f, const chord_draw_cmd* cmds, uint32_t count, uint64_t* vis64,
                        chord_raster_stats* stats)
{
    raster_target rt = { f->width, f->height, f->tile, vis64, (f->passFlags & CHORD_PASS_DEPTH_CLAMP) != 0 };
    uint64_t nSub = 0, nCull = 0, nClip = 0, nMesh = 0, nKilled = 0;
    const float dimW = f->iview->renderDimension.x, dimH = f->iview->renderDimension.y;
#ifdef _OPENMP
    int nthreads = f->threads > 0 ? f->threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+:nSub,nCull,nClip,nMesh,nKilled)
#endif
    for (uint32_t ci = 0; ci < count; ci++) {
        const chord_draw_cmd cmd = cmds[ci];
        object_refs o = get_object(f, cmd.objectId);
        /* pipeline filter (pipeline_filter.hlsl:41-130; mesh_raster.cpp:224-253): alphaMode 0/1 only */
        if (o.material->alphaMode > 1u) continue;
        const int twoSided = o.material->bTwoSided != 0 || (f->passFlags & CHORD_PASS_DEPTH) != 0;   /* mesh_raster.cpp:187-189 */
        masked_attrs mattr; const masked_attrs* masked = NULL;
        if (o.material->alphaMode == 1u) {
            const uint32_t tid = o.material->baseColorId;
            const chord_texture_desc* tex = (f->textures && tid < f->textureCount && f->textures[tid].rgba8) ? &f->textures[tid] : NULL;
            if (!tex || !o.datas->uv0) {
                /* masked, untextured: builtin white texel alpha 1.0 (asset_gltf.cpp:295-326); clip(a*factor - cutoff) */
                float a = fmaf(1.0f, o.material->baseColorFactor.w, -o.material->alphaCutOff);
                if (a < 0.0f) continue;
            } else {
                const uint32_t sid = o.material->baseColorSampler;
                memset(&mattr, 0, sizeof(mattr));
                mattr.tex = tex; mattr.smp = (f->samplers && sid < f->samplerCount) ? &f->samplers[sid] : NULL;
                mattr.alphaFactor = o.material->baseColorFactor.w; mattr.alphaCutOff = o.material->alphaCutOff;
                mattr.killed = &nKilled;
                masked = &mattr;
            }
        }
        const chord_GPUGLTFMeshlet* m = &o.datas->meshlets[cmd.meshletId];
        const uint32_t V = m->vertexTriangleCount & 0xFFu, T = (m->vertexTriangleCount >> 8) & 0xFFu;
        chord_float4x4 mvp; mat_mul(&f->iview->translatedWorldToClip, &o.object->basicData.localToTranslatedWorld, &mvp);
        float hs[CHORD_MESHLET_MAX_VERTICES + 1][4], huv[CHORD_MESHLET_MAX_VERTICES + 1][2];
        for (uint32_t i = 0; i < V; i++) {
            uint32_t vi = o.prim->vertexOffset + o.datas->meshletData[m->dataOffset + i];
            const float* p = &o.datas->positions[(uint64_t)vi * 3];
            f4 h = mat_mul_vec(&mvp, p[0], p[1], p[2], 1.0f);
            hs[i][0] = h.x; hs[i][1] = h.y; hs[i][2] = h.z; hs[i][3] = h.w;
            if (masked) { huv[i][0] = o.datas->uv0[(uint64_t)vi * 2]; huv[i][1] = o.datas->uv0[(uint64_t)vi * 2 + 1]; }   /* mesh_raster.hlsl:107-112 */
        }
        nMesh++; nSub += T;
        for (uint32_t t = 0; t < T; t++) {
            uint32_t idx = o.datas->meshletData[m->dataOffset + V + t];
            uint32_t i0 = idx & 0xFFu, i1 = (idx >> 8) & 0xFFu, i2 = (idx >> 16) & 0xFFu;
            const float p0[3] = { hs[i0][0], hs[i0][1], hs[i0][3] };
            const float p1[3] = { hs[i1][0], hs[i1][1], hs[i1][3] };
            const float p2[3] = { hs[i2][0], hs[i2][1], hs[i2][3] };
            int culled = 0;
            if (!twoSided) culled = det3(p0, p1, p2) <= 0.0f;                                    /* #0 */
            if (!culled) culled = (p0[2] <= 0.0f && p1[2] <= 0.0f && p2[2] <= 0.0f);             /* #1 */
            float u0 = fmaf(p0[0] / fabsf(p0[2]), 0.5f, 0.5f), v0 = fmaf(p0[1] / fabsf(p0[2]), -0.5f, 0.5f);
            float u1 = fmaf(p1[0] / fabsf(p1[2]), 0.5f, 0.5f), v1 = fmaf(p1[1] / fabsf(p1[2]), -0.5f, 0.5f);
            float u2 = fmaf(p2[0] / fabsf(p2[2]), 0.5f, 0.5f), v2 = fmaf(p2[1] / fabsf(p2[2]), -0.5f, 0.5f);
            float maxU = fmaxf(u0, fmaxf(u1, u2)), maxV = fmaxf(v0, fmaxf(v1, v2));
            float minU = fminf(u0, fminf(u1, u2)), minV = fminf(v0, fminf(v1, v2));
            if (!culled) culled = (minU >= 1.0f) || (minV >= 1.0f) || (maxU <= 0.0f) || (maxV <= 0.0f);   /* #2 */
            if (!culled) {                                                                      /* #3 */
                float maxSx = maxU * dimW, maxSy = maxV * dimH, minSx = minU * dimW, minSy = minV * dimH;
                culled = (rintf(minSx) == rintf(maxSx)) || (rintf(minSy) == rintf(maxSy));
            }
            if (culled) { nCull++; continue; }
            const float c[3][4] = {
                { hs[i0][0], hs[i0][1], hs[i0][2], hs[i0][3] },
                { hs[i1][0], hs[i1][1], hs[i1][2], hs[i1][3] },
                { hs[i2][0], hs[i2][1], hs[i2][2], hs[i2][3] } };
            if (masked) {
                masked_attrs mt = *masked;
                mt.uv[0][0] = huv[i0][0]; mt.uv[0][1] = huv[i0][1]; mt.uv[1][0] = huv[i1][0]; mt.uv[1][1] = huv[i1][1];
                mt.uv[2][0] = huv[i2][0]; mt.uv[2][1] = huv[i2][1];
                nClip += (uint64_t)raster_clip_triangle(&rt, c, chord_encode_triangle_instance(t, cmd.slot), twoSided, &mt);
            } else {
                nClip += (uint64_t)raster_clip_triangle(&rt, c, chord_encode_triangle_instance(t, cmd.slot), twoSided, NULL);
            }
        }
    }
    if (stats) {
        stats->trianglesSubmitted += nSub; stats->trianglesCulled += nCull; stats->trianglesClipped += nClip;
        stats->meshletsRastered += nMesh; stats->fragmentsAlphaKilled += nKilled;
    }
}

/* ------------------------------------------------------------------ K8 */
static uint32_t next_pot(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

void oracle_hzb_dims(uint32_t depthW, uint32_t depthH, uint32_t* mip0W, uint32_t* mip0H, uint32_t* mipCount)
{
    /* hzb.cpp:49-56 */
    uint32_t tw = next_pot(depthW) / 2, th = next_pot(depthH) / 2;
    if (tw == depthW) tw /= 2;
    if (th == depthH) th /= 2;
    uint32_t mx = tw > th ? tw : th, mips = 1;
    while (mx > 1) { mx >>= 1; mips++; }
    *mip0W = tw; *mip0H = th; *mipCount = mips;
}

void oracle_build_hzb(const uint64_t* vis64, uint32_t width, uint32_t height,
                      int bMin, int bMax, int bValidRange,
                      uint16_t* minChain, uint16_t* maxChain, const uint64_t* mipOffsets,
                      uint32_t mip0W, uint32_t mip0H, uint32_t mipCount, uint32_t validRangeBits[2])
{
    /* mip0 (hzb.hlsl:127-165 / hzb_one.hlsl:131-156): 2x2 footprint with edge clamp, fp32 min/max */
    float* fmin_ = (float*)malloc(sizeof(float) * (size_t)mip0W * mip0H);
    float* fmax_ = (float*)malloc(sizeof(float) * (size_t)mip0W * mip0H);
    uint32_t vmin = ~0u, vmax = 0u;
    for (uint32_t y = 0; y < mip0H; y++)
        for (uint32_t x = 0; x < mip0W; x++) {
            float mn = 0, mx = 0;
            for (int k = 0; k < 4; k++) {
                uint32_t sx = 2 * x + (uint32_t)(k >> 1), sy = 2 * y + (uint32_t)(k & 1);
                if (sx > width - 1) sx = width - 1;
                if (sy > height - 1) sy = height - 1;
                float d = u2f((uint32_t)(vis64[(uint64_t)sy * width + sx] >> 32));
                if (k == 0) { mn = d; mx = d; } else { mn = fminf(mn, d); mx = fmaxf(mx, d); }
                if (bValidRange && d > 0.0f) {
                    /* wave min/max then InterlockedMin/Max on asuint; guards zMin<1 / zMax>0 (hzb.hlsl:165-172) */
                    uint32_t b = f2u(d);
                    if (d < 1.0f && b < vmin) vmin = b;
                    if (b > vmax) vmax = b;
                }
            }
            fmin_[(size_t)y * mip0W + x] = mn; fmax_[(size_t)y * mip0W + x] = mx;
        }
    if (validRangeBits) { validRangeBits[0] = vmin; validRangeBits[1] = vmax; }

    /* mips 0..5 are reduced in fp32 inside one group and stored as R16F; mip 5 is re-read (fp16) for 6.. .
     * min: stored RTZ (conservative for reversed-Z; hzb.hlsl:69). max: raw RTZ for mips 0-4, mip 5 stored
     * as half_bits(RTZ(v))+1 (hzb.hlsl:70) and everything above derives from that. */
    uint32_t w = mip0W, h = mip0H;
    for (uint32_t mip = 0; mip < mipCount; mip++) {
        uint16_t* dmin = bMin ? minChain + mipOffsets[mip] : NULL;
        uint16_t* dmax = bMax ? maxChain + mipOffsets[mip] : NULL;
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                size_t i = (size_t)y * w + x;
                if (dmin) dmin[i] = oracle_f32_to_f16_rtz(fmin_[i]);
                if (dmax) {
                    uint16_t hb = oracle_f32_to_f16_rtz(fmax_[i]);
                    if (mip == 5) hb = (uint16_t)(hb + 1);
                    dmax[i] = hb;
                }
            }
        if (mip == 5) {   /* later mips read back the stored fp16 values */
            for (size_t i = 0; i < (size_t)w * h; i++) {
                if (dmin) fmin_[i] = oracle_f16_to_f32(dmin[i]);
                if (dmax) fmax_[i] = oracle_f16_to_f32(dmax[i]);
            }
        }
        if (mip + 1 == mipCount) break;
        uint32_t nw = w > 1 ? w / 2 : 1, nh = h > 1 ? h / 2 : 1;
        for (uint32_t y = 0; y < nh; y++)
            for (uint32_t x = 0; x < nw; x++) {
                uint32_t x0 = (w > 1) ? 2 * x : 0, x1 = (w > 1) ? 2 * x + 1 : 0;
                uint32_t y0 = (h > 1) ? 2 * y : 0, y1 = (h > 1) ? 2 * y + 1 : 0;
                float a = fmin_[(size_t)y0 * w + x0], b = fmin_[(size_t)y0 * w + x1], c = fmin_[(size_t)y1 * w + x0], d = fmin_[(size_t)y1 * w + x1];
                float A = fmax_[(size_t)y0 * w + x0], B = fmax_[(size_t)y0 * w + x1], C = fmax_[(size_t)y1 * w + x0], D = fmax_[(size_t)y1 * w + x1];
                float mn = fminf(fminf(fminf(a, b), c), d), mx = fmaxf(fmaxf(fmaxf(A, B), C), D);
                /* write in place is safe: index y*nw+x <= y0*w+x0 */
                fmin_[(size_t)y * nw + x] = mn; fmax_[(size_t)y * nw + x] = mx;
            }
        w = nw; h = nh;
    }
    free(fmin_); free(fmax_);
}

/* ------------------------------------------------------------------ K9 + decode */
void oracle_decode(const uint64_t* vis64, uint32_t pixelCount, const chord_draw_cmd* postCull, uint32_t postCullCount,
                   uint32_t* triples)
{
    for (uint32_t i = 0; i < pixelCount; i++) {
        uint32_t id = (uint32_t)(vis64[i] & 0xFFFFFFFFu);
        if (id == 0) { triples[3 * i] = triples[3 * i + 1] = triples[3 * i + 2] = 0xFFFFFFFFu; continue; }
        uint32_t tri, inst; chord_decode_triangle_instance(id, &tri, &inst);
        if (inst >= postCullCount) { triples[3 * i] = triples[3 * i + 1] = 0xFFFFFFFEu; triples[3 * i + 2] = tri; continue; }
        triples[3 * i] = postCull[inst].objectId; triples[3 * i + 1] = postCull[inst].meshletId; triples[3 * i + 2] = tri;
    }
}

void oracle_visibility_mark(const oracle_frame* f, const uint64_t* vis64,
                            const chord_draw_cmd* postCull, uint32_t postCullCount, uint32_t* markerUint4)
{
    /* visibility_tile.hlsl:39-131: per 8x8 tile OR of (1 << shadingType); Gather with clamp-to-edge sampler */
    const uint32_t mw = (f->width + 7) / 8, mh = (f->height + 7) / 8;
    (void)postCullCount;
    for (uint32_t ty = 0; ty < mh; ty++)
        for (uint32_t tx = 0; tx < mw; tx++) {
            uint32_t mk[4] = {0, 0, 0, 0};
            for (uint32_t y = 0; y < 8; y++)
                for (uint32_t x = 0; x < 8; x++) {
                    uint32_t px = tx * 8 + x, py = ty * 8 + y;
                    if (px > f->width - 1) px = f->width - 1;
                    if (py > f->height - 1) py = f->height - 1;
                    uint32_t id = (uint32_t)(vis64[(uint64_t)py * f->width + px] & 0xFFFFFFFFu);
                    uint32_t type = CHORD_LIGHTING_TYPE_NONE;
                    if (id != 0) {
                        uint32_t tri, inst; chord_decode_triangle_instance(id, &tri, &inst);
                        object_refs o = get_object(f, postCull[inst].objectId);
                        type = o.material->materialType;
                    }
                    mk[type / 32] |= 1u << (type % 32);
                }
            memcpy(&markerUint4[4 * ((size_t)ty * mw + tx)], mk, 16);
        }
}

/* ------------------------------------------------------------------ K11: lighting.hlsl:270-381 (PBR branch)
 * Floating point within 1e-5 (north_star), so plain fp32 expressions in the HLSL's order; no bit-exactness claim.
 * Pins chosen where the reference leaves it to the GPU: LUT SampleLevel = exact fp32 bilinear with clamp-to-edge on the
 * RGBA16F texels; pow(x,5) = repeated multiplication; sRGB_2_AP1 folded in double and rounded to float. */
typedef struct { float x, y, z; } v3;
static inline v3 v3_(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3add(v3 a, v3 b) { return v3_(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3sub(v3 a, v3 b) { return v3_(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3mul(v3 a, v3 b) { return v3_(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3s(v3 a, float s) { return v3_(a.x * s, a.y * s, a.z * s); }
static inline float v3dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* HLSL normalize() in K11: v * rcp(sqrt(dot(v, v))) - IEEE square root, IEEE reciprocal, three products.
 * (DXC lowers normalize to GLSL.std.450 Normalize, which drivers implement as v * inversesqrt(dot(v, v)); the reciprocal form is
 * the faithful one and costs one division instead of three.) */
static inline v3 v3norm(v3 a) { const float r = 1.0f / sqrtf(v3dot(a, a)); return v3_(a.x * r, a.y * r, a.z * r); }
static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

static void srgb_to_ap1_matrix(float m[9])
{
    /* colorspace.h:109: XYZ_2_AP1 * (D65_2_D60 * sRGB_2_XYZ) */
    static const double X2A[9] = { 1.6410233797, -0.3248032942, -0.2364246952, -0.6636628587, 1.6153315917, 0.0167563477, 0.0117218943, -0.0082844420, 0.9883948585 };
    static const double CAT[9] = { 1.0130349146, 0.0061052578, -0.0149709436, 0.0076982301, 0.9981633521, -0.0050320385, -0.0028413174, 0.0046851567, 0.9245061375 };
    static const double S2X[9] = { 0.4123907993, 0.3575843394, 0.1804807884, 0.2126390059, 0.7151686788, 0.0721923154, 0.0193308187, 0.1191947798, 0.9505321522 };
    double t[9], o[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { t[r * 3 + c] = 0; for (int k = 0; k < 3; k++) t[r * 3 + c] += CAT[r * 3 + k] * S2X[k * 3 + c]; }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { o[r * 3 + c] = 0; for (int k = 0; k < 3; k++) o[r * 3 + c] += X2A[r * 3 + k] * t[k * 3 + c]; }
    for (int i = 0; i < 9; i++) m[i] = (float)o[i];
}
static inline v3 mat3_mul(const float m[9], v3 v)
{ return v3_(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z); }

static v3 sample_lut_bilinear(const uint16_t* lut, int W, int H, float u, float v)
{
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float fx = x - fx0, fy = y - fy0;
    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    if (x0 < 0) x0 = 0; if (x0 > W - 1) x0 = W - 1; if (x1 < 0) x1 = 0; if (x1 > W - 1) x1 = W - 1;
    if (y0 < 0) y0 = 0; if (y0 > H - 1) y0 = H - 1; if (y1 < 0) y1 = 0; if (y1 > H - 1) y1 = H - 1;
    float o[3];
    for (int c = 0; c < 3; c++) {
        float a = oracle_f16_to_f32(lut[(y0 * W + x0) * 4 + c]), b = oracle_f16_to_f32(lut[(y0 * W + x1) * 4 + c]);
        float d = oracle_f16_to_f32(lut[(y1 * W + x0) * 4 + c]), e = oracle_f16_to_f32(lut[(y1 * W + x1) * 4 + c]);
        float top = a + (b - a) * fx, bot = d + (e - d) * fx;
        o[c] = top + (bot - top) * fy;
    }
    return v3_(o[0], o[1], o[2]);
}

static inline float tex_coord_from_unit(float x, int size) { return 0.5f / (float)size + x * (1.0f - 1.0f / (float)size); }

/* atmosphere.hlsli:1063-1088 */
static v3 sun_and_sky_irradiance(const chord_AtmosphereParameters* at, const uint16_t* trLut, const uint16_t* irLut, const float ap1[9],
                                 v3 p, v3 normal, v3 sunDir, v3* skyOut)
{
    float r = sqrtf(v3dot(p, p));
    float mu_s = v3dot(p, sunDir) / r;
    v3 skyL = at->luminanceMode != 0 ? v3_(at->skySpectralRadianceToLumiance.x, at->skySpectralRadianceToLumiance.y, at->skySpectralRadianceToLumiance.z) : v3_(1, 1, 1);
    /* GetIrradiance (atmosphere.hlsli:894-902, :847-856) */
    float x_r = (r - at->bottom_radius) / (at->top_radius - at->bottom_radius);
    float x_mu_s = mu_s * 0.5f + 0.5f;
    v3 irr = sample_lut_bilinear(irLut, 64, 16, tex_coord_from_unit(x_mu_s, 64), tex_coord_from_unit(x_r, 16));
    v3 sky = v3s(v3s(irr, 1.0f + v3dot(normal, p) / r), 0.5f);
    sky = v3mul(sky, skyL);
    *skyOut = mat3_mul(ap1, sky);
    v3 sunL = at->luminanceMode != 0 ? v3_(at->sunSpectralRadianceToLumiance.x, at->sunSpectralRadianceToLumiance.y, at->sunSpectralRadianceToLumiance.z) : v3_(1, 1, 1);
    /* GetTransmittanceToSun (atmosphere.hlsli:315-324) + GetTransmittanceTextureUvFromRMu (:193-215) */
    float sin_h = at->bottom_radius / r;
    float cos_h = -sqrtf(fmaxf(1.0f - sin_h * sin_h, 0.0f));
    float Hh = sqrtf(at->top_radius * at->top_radius - at->bottom_radius * at->bottom_radius);
    float rho = sqrtf(fmaxf(r * r - at->bottom_radius * at->bottom_radius, 0.0f));
    float disc = r * r * (mu_s * mu_s - 1.0f) + at->top_radius * at->top_radius;
    float d = fmaxf(-r * mu_s + sqrtf(fmaxf(disc, 0.0f)), 0.0f);
    float d_min = at->top_radius - r, d_max = rho + Hh;
    float x_mu = (d - d_min) / (d_max - d_min), xr = rho / Hh;
    v3 tr = sample_lut_bilinear(trLut, 256, 64, tex_coord_from_unit(x_mu, 256), tex_coord_from_unit(xr, 64));
    float e0 = -sin_h * at->sun_angular_radius, e1 = sin_h * at->sun_angular_radius;
    float t = clampf(((mu_s - cos_h) - e0) / (e1 - e0), 0.0f, 1.0f);
    float sm = t * t * (3.0f - 2.0f * t);
    v3 trs = v3s(tr, sm);
    v3 sunRad = v3s(v3mul(v3_(at->solar_irradiance.x, at->solar_irradiance.y, at->solar_irradiance.z), trs), fmaxf(v3dot(normal, sunDir), 0.0f));
    return mat3_mul(ap1, v3mul(sunL, sunRad));
}

static inline double storage_to_double(chord_uint2 u) { uint64_t b = (uint64_t)u.x | ((uint64_t)u.y << 32); double d; memcpy(&d, &b, 8); return d; }


/* ------------------------------------------------------------------ K11, LIGHTING_TYPE == None: the sky branch
 * lighting.hlsl:75-166 (skyColorEvaluate) + atmosphere.hlsli:948-997 (GetSkyRadiance), :916-946 (GetCombinedScattering),
 * :405-462 (GetScatteringTextureUvwzFromRMuMuSNu), :904-914, :116-121, :99-106; base.hlsli:601-615 (phase functions).
 * Pins: Texture3D SampleLevel(linear clamp, 0) = fp32 trilinear on the RGBA16F texels (x, then y, then z,
 * each a + (b - a) * t); pow(x, 1.5) in the Cornette-Shanks phase = x * sqrt(x) (IEEE), the transcendental is implementation
 * defined on a GPU; everything else plain fp32 in the HLSL's order; normalize = v * rcp(sqrt(dot)). */
#define SC_R 32
#define SC_MU 128
#define SC_MU_S 32
#define SC_NU 8
typedef struct { float x, y, z, w; } v4;
static v4 sample_lut3d(const uint16_t* lut, float u, float v, float w)
{
    const int W = SC_NU * SC_MU_S, H = SC_MU, D = SC_R;
    const float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f, z = w * (float)D - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0, x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
#define CL(a, n) do { if (a < 0) a = 0; if (a > n - 1) a = n - 1; } while (0)
    CL(x0, W); CL(x1, W); CL(y0, H); CL(y1, H); CL(z0, D); CL(z1, D);
#undef CL
    float o[4];
    for (int c = 0; c < 4; c++) {
        float zz[2];
        for (int k = 0; k < 2; k++) {
            const int zi = k ? z1 : z0;
            const uint16_t* s = lut + ((size_t)zi * H * W) * 4;
            const float a = oracle_f16_to_f32(s[((size_t)y0 * W + x0) * 4 + c]), b = oracle_f16_to_f32(s[((size_t)y0 * W + x1) * 4 + c]);
            const float d = oracle_f16_to_f32(s[((size_t)y1 * W + x0) * 4 + c]), e = oracle_f16_to_f32(s[((size_t)y1 * W + x1) * 4 + c]);
            const float top = a + (b - a) * fx, bot = d + (e - d) * fx;
            zz[k] = top + (bot - top) * fy;
        }
        o[c] = zz[0] + (zz[1] - zz[0]) * fz;
    }
    v4 r = { o[0], o[1], o[2], o[3] };
    return r;
}

static inline float safe_sqrt(float a) { return sqrtf(fmaxf(a, 0.0f)); }
static inline float dist_to_top(const chord_AtmosphereParameters* at, float r, float mu)
{ const float disc = r * r * (mu * mu - 1.0f) + at->top_radius * at->top_radius; return fmaxf(-r * mu + safe_sqrt(disc), 0.0f); }

static v3 sky_radiance(const chord_AtmosphereParameters* at, const uint16_t* scLut, const uint16_t* mieLut, const float ap1[9],
                       v3 camera, v3 ray, v3 sunDir)
{
    float r = sqrtf(v3dot(camera, camera));
    float rmu = v3dot(camera, ray);
    const float dtop = -rmu - sqrtf(rmu * rmu - r * r + at->top_radius * at->top_radius);
    if (dtop > 0.0f) { camera = v3add(camera, v3s(ray, dtop)); r = at->top_radius; rmu += dtop; }
    else if (r > at->top_radius) return v3_(0, 0, 0);
    const float mu = rmu / r;
    const float mu_s = v3dot(camera, sunDir) / r;
    const float nu = v3dot(ray, sunDir);
    const int hitsGround = mu < 0.0f && r * r * (mu * mu - 1.0f) + at->bottom_radius * at->bottom_radius >= 0.0f;
    /* GetScatteringTextureUvwzFromRMuMuSNu */
    const float H = sqrtf(at->top_radius * at->top_radius - at->bottom_radius * at->bottom_radius);
    const float rho = safe_sqrt(r * r - at->bottom_radius * at->bottom_radius);
    const float u_r = tex_coord_from_unit(rho / H, SC_R);
    const float r_mu = r * mu;
    const float disc = r_mu * r_mu - r * r + at->bottom_radius * at->bottom_radius;
    float u_mu;
    if (hitsGround) {
        const float d = -r_mu - safe_sqrt(disc), d_min = r - at->bottom_radius, d_max = rho;
        u_mu = 0.5f - 0.5f * tex_coord_from_unit(d_max == d_min ? 0.0f : (d - d_min) / (d_max - d_min), SC_MU / 2);
    } else {
        const float d = -r_mu + safe_sqrt(disc + H * H), d_min = at->top_radius - r, d_max = rho + H;
        u_mu = 0.5f + 0.5f * tex_coord_from_unit((d - d_min) / (d_max - d_min), SC_MU / 2);
    }
    const float d = dist_to_top(at, at->bottom_radius, mu_s);
    const float d_min = at->top_radius - at->bottom_radius, d_max = H;
    const float a = (d - d_min) / (d_max - d_min);
    const float Dd = dist_to_top(at, at->bottom_radius, at->mu_s_min);
    const float A = (Dd - d_min) / (d_max - d_min);
    const float u_mu_s = tex_coord_from_unit(fmaxf(1.0f - a / A, 0.0f) / (1.0f + a), SC_MU_S);
    const float u_nu = (nu + 1.0f) / 2.0f;
    /* GetCombinedScattering */
    const float tcx = u_nu * (float)(SC_NU - 1);
    const float tex_x = floorf(tcx), lerp = tcx - tex_x;
    const float ua = (tex_x + u_mu_s) / (float)SC_NU, ub = (tex_x + 1.0f + u_mu_s) / (float)SC_NU;
    const v4 s0 = sample_lut3d(scLut, ua, u_mu, u_r), s1 = sample_lut3d(scLut, ub, u_mu, u_r);
    const float w0 = 1.0f - lerp;
    v3 scattering = v3_(s0.x * w0 + s1.x * lerp, s0.y * w0 + s1.y * lerp, s0.z * w0 + s1.z * lerp), mie;
    if (at->bCombineScattering != 0) {
        const float ca = s0.w * w0 + s1.w * lerp;
        if (scattering.x <= 0.0f) mie = v3_(0, 0, 0);
        else {
            const float k = at->rayleigh_scattering.x / at->mie_scattering.x;
            mie = v3_(scattering.x * ca / scattering.x * k * (at->mie_scattering.x / at->rayleigh_scattering.x),
                      scattering.y * ca / scattering.x * k * (at->mie_scattering.y / at->rayleigh_scattering.y),
                      scattering.z * ca / scattering.x * k * (at->mie_scattering.z / at->rayleigh_scattering.z));
        }
    } else {
        const v4 m0 = sample_lut3d(mieLut, ua, u_mu, u_r), m1 = sample_lut3d(mieLut, ub, u_mu, u_r);
        mie = v3_(m0.x * w0 + m1.x * lerp, m0.y * w0 + m1.y * lerp, m0.z * w0 + m1.z * lerp);
    }
    const float kPI = 3.14159265358979323846f;
    const float rayleigh = (3.0f / (16.0f * kPI)) * (1.0f + nu * nu);
    const float g = at->mie_phase_function_g, g2 = g * g;
    const float kk = 3.0f / (8.0f * kPI) * (1.0f - g2) / (2.0f + g2);
    const float base = 1.0f + g2 - 2.0f * g * nu;
    const float miePhase = kk * (1.0f + nu * nu) / (base * sqrtf(base));             /* pow(base, 1.5) pinned as base * sqrt(base) */
    const v3 skyL = at->luminanceMode != 0 ? v3_(at->skySpectralRadianceToLumiance.x, at->skySpectralRadianceToLumiance.y, at->skySpectralRadianceToLumiance.z) : v3_(1, 1, 1);
    const v3 sc = v3add(v3s(scattering, rayleigh), v3s(mie, miePhase));
    return mat3_mul(ap1, v3mul(sc, skyL));
}

/* lighting.hlsl:75-166 for every pixel of the tile without geometry; writes the colour plane only (exportGbuffer with
 * LIGHTING_TYPE == None stores sceneColor alone, lighting.hlsl:62-73) */
void oracle_sky(const oracle_frame* f, const uint64_t* vis64, const uint16_t* trLut, const uint16_t* irLut, const uint16_t* scLut,
                const uint16_t* mieLut, float* colorPlane)
{
    float ap1[9]; srgb_to_ap1_matrix(ap1);
    const chord_PerframeCameraView* view = f->view;
    const chord_AtmosphereParameters* at = &view->basicData.atmosphere;
    if (!mieLut) mieLut = scLut;      /* bCombineScattering == 0 without a single-Mie volume: undefined in the reference (bindless id 0); pinned */
    const float tsx = 1.0f / (float)f->width, tsy = 1.0f / (float)f->height;
    const v3 sunDir = v3_(-view->basicData.sunInfo.direction.x, -view->basicData.sunInfo.direction.y, -view->basicData.sunInfo.direction.z);
    const v3 camEarth = v3_((float)storage_to_double(view->cameraToEarthCenter_km.x), (float)storage_to_double(view->cameraToEarthCenter_km.y), (float)storage_to_double(view->cameraToEarthCenter_km.z));
    const v3 camKm = v3_((float)storage_to_double(view->cameraPositionWS_km.x), (float)storage_to_double(view->cameraPositionWS_km.y), (float)storage_to_double(view->cameraPositionWS_km.z));
    const v3 earthC = v3_(at->earthCenterKm.x, at->earthCenterKm.y, at->earthCenterKm.z);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (uint32_t py = f->tile.y0; py < f->tile.y1; py++)
        for (uint32_t px = f->tile.x0; px < f->tile.x1; px++) {
            const size_t pi = (size_t)py * f->width + px;
            if ((uint32_t)(vis64[pi] & 0xFFFFFFFFu) != 0u) continue;
            const float su = ((float)px + 0.5f) * tsx, sv = ((float)py + 0.5f) * tsy;
            const f4 vp = mat_mul_vec(&view->clipToTranslatedWorld, 2.0f * (su - 0.5f), 2.0f * (0.5f - sv), 1.192092896e-07f, 1.0f);
            const v3 dir = v3norm(v3_(vp.x / vp.w, vp.y / vp.w, vp.z / vp.w));
            float groundAlpha = 0.0f; v3 ground = v3_(0, 0, 0);
            const float pov = v3dot(camEarth, dir), pop = v3dot(camEarth, camEarth);
            const float rsq = pop - pov * pov;
            const float d2i = -pov - sqrtf(at->bottom_radius * at->bottom_radius - rsq);
            if (d2i > 0.0f) {
                const v3 ip = v3add(camKm, v3s(dir, d2i));
                const v3 rel = v3sub(ip, earthC);
                const v3 normal = v3norm(rel);
                v3 skyIrr;
                const v3 sunIrr = sun_and_sky_irradiance(at, trLut, irLut, ap1, rel, normal, sunDir, &skyIrr);
                const v3 alb = v3s(v3_(at->ground_albedo.x, at->ground_albedo.y, at->ground_albedo.z), 1.0f / 3.14159265358979323846f);
                ground = v3mul(alb, v3add(v3s(sunIrr, 1.0f), v3s(skyIrr, 1.0f)));
                groundAlpha = 1.0f;
            }
            v3 rad = sky_radiance(at, scLut, mieLut, ap1, camEarth, dir, sunDir);
            rad = v3_(rad.x + (ground.x - rad.x) * groundAlpha, rad.y + (ground.y - rad.y) * groundAlpha, rad.z + (ground.z - rad.z) * groundAlpha);
            const float k = at->luminanceMode == 0 ? 1.0f : 1e-5f;
            v3 color = v3_(10.0f * rad.x * k, 10.0f * rad.y * k, 10.0f * rad.z * k);
            if (isnan(color.x) || isnan(color.y) || isnan(color.z)) color = v3_(0, 0, 0);
            float* o0 = colorPlane + pi * 4; o0[0] = color.x; o0[1] = color.y; o0[2] = color.z; o0[3] = 0.0f;
        }
}

static inline const chord_texture_desc* tex_lookup(const oracle_frame* f, uint32_t id)
{ return (f->textures && id < f->textureCount && f->textures[id].rgba8) ? &f->textures[id] : NULL; }
static inline const chord_sampler_desc* smp_lookup(const oracle_frame* f, uint32_t id)
{ return (f->samplers && id < f->samplerCount) ? &f->samplers[id] : NULL; }

void oracle_lighting(const oracle_frame* f, const uint64_t* vis64, const chord_draw_cmd* postCull, uint32_t postCullCount,
                     const uint16_t* trLut, const uint16_t* irLut, float* const planes[6])
{
    (void)postCullCount;
    float ap1[9]; srgb_to_ap1_matrix(ap1);
    const chord_PerframeCameraView* view = f->view;
    const chord_AtmosphereParameters* at = &view->basicData.atmosphere;
    const float tsx = 1.0f / (float)f->width, tsy = 1.0f / (float)f->height;     /* visibilityTexelSize (lighting.cpp) */
    const size_t npx = (size_t)f->width * f->height;
    for (int k = 0; k < 6; k++) memset(planes[k], 0, npx * 16);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (uint32_t py = f->tile.y0; py < f->tile.y1; py++)
        for (uint32_t px = f->tile.x0; px < f->tile.x1; px++) {
            const size_t pi = (size_t)py * f->width + px;
            uint32_t id = (uint32_t)(vis64[pi] & 0xFFFFFFFFu);
            if (id == 0) continue;
            uint32_t tri, inst; chord_decode_triangle_instance(id, &tri, &inst);
            const chord_draw_cmd cmd = postCull[inst];
            object_refs o = get_object(f, cmd.objectId);
            if (o.material->materialType != CHORD_LIGHTING_TYPE_GLTF_PBR) continue;
            const chord_float4x4* l2w = &o.object->basicData.localToTranslatedWorld;
            chord_float4x4 l2c, l2cNJ, l2cLast;
            mat_mul(&view->translatedWorldToClip, l2w, &l2c);
            mat_mul(&view->translatedWorldToClip_NoJitter, l2w, &l2cNJ);
            mat_mul(&view->translatedWorldToClipLastFrame_NoJitter, &o.object->basicData.localToTranslatedWorldLastFrame, &l2cLast);
            /* getTriangleMiscInfo (nanite_shared.hlsli:111-179) */
            const chord_GPUGLTFMeshlet* m = &o.datas->meshlets[cmd.meshletId];
            const uint32_t V = m->vertexTriangleCount & 0xFFu;
            const uint32_t idx = o.datas->meshletData[m->dataOffset + V + tri];
            v3 posRS[3], nrm[3], tanRS[3], bitRS[3]; f4 hs[3]; v3 hsNJ[3], hsLast[3]; float uv[3][2];
            const chord_GLTFMaterialGPUData* matT = o.material;
            /* a texture id "exists" when it names an uploaded texture; the normal map additionally needs the tangent stream */
            const chord_texture_desc* texBase = tex_lookup(f, matT->baseColorId), *texNrm = tex_lookup(f, matT->normalTexture), *texMR = tex_lookup(f, matT->metallicRoughnessTexture);
            const int bExistNormalTexture = texNrm != NULL && o.datas->tangents != NULL && o.datas->uv0 != NULL;
            for (int i = 0; i < 3; i++) {
                uint32_t vi = o.prim->vertexOffset + o.datas->meshletData[m->dataOffset + ((idx >> (i * 8)) & 0xFFu)];
                const float* p = &o.datas->positions[(size_t)vi * 3];
                f4 w = mat_mul_vec(l2w, p[0], p[1], p[2], 1.0f);
                posRS[i] = v3_(w.x, w.y, w.z);
                hs[i] = mat_mul_vec(&l2c, p[0], p[1], p[2], 1.0f);
                f4 a = mat_mul_vec(&l2cNJ, p[0], p[1], p[2], 1.0f); hsNJ[i] = v3_(a.x, a.y, a.w);
                f4 b = mat_mul_vec(&l2cLast, p[0], p[1], p[2], 1.0f); hsLast[i] = v3_(b.x, b.y, b.w);
                uv[i][0] = o.datas->uv0 ? o.datas->uv0[(size_t)vi * 2] : 0.0f; uv[i][1] = o.datas->uv0 ? o.datas->uv0[(size_t)vi * 2 + 1] : 0.0f;
                const float* n = &o.datas->normals[(size_t)vi * 3];
                /* mul(float4(n,0), translatedWorldToLocal): row vector times matrix */
                const float* M = o.object->basicData.translatedWorldToLocal.m;
                v3 nn = v3_(n[0] * M[0] + n[1] * M[1] + n[2] * M[2], n[0] * M[4] + n[1] * M[5] + n[2] * M[6], n[0] * M[8] + n[1] * M[9] + n[2] * M[10]);
                nrm[i] = v3norm(nn);
                tanRS[i] = bitRS[i] = v3_(0, 0, 0);
                if (bExistNormalTexture) {                                   /* nanite_shared.hlsli:160-176 */
                    const float* tg = &o.datas->tangents[(size_t)vi * 4];
                    f4 tw = mat_mul_vec(l2w, tg[0], tg[1], tg[2], 0.0f);
                    v3 t = v3_(tw.x, tw.y, tw.z);
                    t = v3norm(v3sub(t, v3s(nrm[i], v3dot(t, nrm[i]))));     /* Gram-Schmidt */
                    tanRS[i] = t;
                    const v3 n_ = nrm[i];
                    bitRS[i] = v3s(v3_(n_.y * t.z - n_.z * t.y, n_.z * t.x - n_.x * t.z, n_.x * t.y - n_.y * t.x), tg[3]);
                }
            }
            /* calculateTriangleBarycentrics (base.hlsli:457-498) */
            const float su = ((float)px + 0.5f) * tsx, sv = ((float)py + 0.5f) * tsy;
            const float pcx = 2.0f * (su - 0.5f), pcy = 2.0f * (0.5f - sv);
            const float rw0 = 1.0f / hs[0].w, rw1 = 1.0f / hs[1].w, rw2 = 1.0f / hs[2].w;
            const float p0x = hs[0].x * rw0, p0y = hs[0].y * rw0, p1x = hs[1].x * rw1, p1y = hs[1].y * rw1, p2x = hs[2].x * rw2, p2y = hs[2].y * rw2;
            const float cdx0 = p2y - p1y, cdx1 = p0y - p2y, cdx2 = p1y - p0y;          /* Pos201Y - Pos120Y */
            const float cdy0 = p1x - p2x, cdy1 = p2x - p0x, cdy2 = p0x - p1x;          /* Pos120X - Pos201X */
            const float C0 = cdx0 * (pcx - p1x) + cdy0 * (pcy - p1y);
            const float C1 = cdx1 * (pcx - p2x) + cdy1 * (pcy - p2y);
            const float C2 = cdx2 * (pcx - p0x) + cdy2 * (pcy - p0y);
            const float G0 = C0 * rw0, G1 = C1 * rw1, G2 = C2 * rw2;
            const float Hs = C0 * rw0 + C1 * rw1 + C2 * rw2;
            const float rH = 1.0f / Hs;
            const float b0 = G0 * rH, b1 = G1 * rH, b2 = G2 * rH;
            /* analytic screen-space derivatives of the barycentrics (base.hlsli:478-496) and of uv (material.hlsli:52-54) */
            const float Gdx0 = cdx0 * rw0, Gdx1 = cdx1 * rw1, Gdx2 = cdx2 * rw2, Gdy0 = cdy0 * rw0, Gdy1 = cdy1 * rw1, Gdy2 = cdy2 * rw2;
            const float Hdx = cdx0 * rw0 + cdx1 * rw1 + cdx2 * rw2, Hdy = cdy0 * rw0 + cdy1 * rw1 + cdy2 * rw2;
            const float rH2 = rH * rH, sxs = 2.0f * tsx, sys = -2.0f * tsy;
            const float ddx0 = (Gdx0 * Hs - G0 * Hdx) * rH2 * sxs, ddx1 = (Gdx1 * Hs - G1 * Hdx) * rH2 * sxs, ddx2 = (Gdx2 * Hs - G2 * Hdx) * rH2 * sxs;
            const float ddy0 = (Gdy0 * Hs - G0 * Hdy) * rH2 * sys, ddy1 = (Gdy1 * Hs - G1 * Hdy) * rH2 * sys, ddy2 = (Gdy2 * Hs - G2 * Hdy) * rH2 * sys;
            const float mu = uv[0][0] * b0 + uv[1][0] * b1 + uv[2][0] * b2, mv = uv[0][1] * b0 + uv[1][1] * b1 + uv[2][1] * b2;
            const float mudx = uv[0][0] * ddx0 + uv[1][0] * ddx1 + uv[2][0] * ddx2, mvdx = uv[0][1] * ddx0 + uv[1][1] * ddx1 + uv[2][1] * ddx2;
            const float mudy = uv[0][0] * ddy0 + uv[1][0] * ddy1 + uv[2][0] * ddy2, mvdy = uv[0][1] * ddy0 + uv[1][1] * ddy1 + uv[2][1] * ddy2;
            /* motion vector */
            v3 mp = v3add(v3add(v3s(hsNJ[0], b0), v3s(hsNJ[1], b1)), v3s(hsNJ[2], b2));
            v3 ml = v3add(v3add(v3s(hsLast[0], b0), v3s(hsLast[1], b1)), v3s(hsLast[2], b2));
            const float mvx = (ml.x / ml.z - mp.x / mp.z) * 0.5f, mvy = (ml.y / ml.z - mp.y / mp.z) * -0.5f;
            v3 posR = v3add(v3add(v3s(posRS[0], b0), v3s(posRS[1], b1)), v3s(posRS[2], b2));
            /* material (material.hlsli:66-153).  A base colour id that names no uploaded texture is the builtin white the reference
             * binds (asset_gltf.cpp:295-326); the emissive texture is sampled by the reference but its value is never added to the
             * colour on this path (material.hlsli:81 stores it, lighting.hlsl never reads it), so it is not sampled here */
            const chord_GLTFMaterialGPUData* mat = o.material;
            float bc[4];
            oracle_sample_rgba(texBase, smp_lookup(f, mat->baseColorSampler), mu, mv, mudx, mvdx, mudy, mvdy, bc);
            v3 base = mat3_mul(ap1, v3_(bc[0] * mat->baseColorFactor.x, bc[1] * mat->baseColorFactor.y, bc[2] * mat->baseColorFactor.z));
            v3 vnrm = v3add(v3add(v3s(nrm[0], b0), v3s(nrm[1], b1)), v3s(nrm[2], b2));
            v3 pnrm = vnrm;
            if (bExistNormalTexture) {
                const v3 tRS = v3add(v3add(v3s(tanRS[0], b0), v3s(tanRS[1], b1)), v3s(tanRS[2], b2));
                const v3 bRS = v3add(v3add(v3s(bitRS[0], b0), v3s(bitRS[1], b1)), v3s(bitRS[2], b2));
                float nt[4];
                oracle_sample_rgba(texNrm, smp_lookup(f, mat->normalSampler), mu, mv, mudx, mvdx, mudy, mvdy, nt);
                float nx = nt[0] * 2.0f - 1.0f, ny = nt[1] * 2.0f - 1.0f;
                const float nz = sqrtf(1.0f - (nx * nx + ny * ny));
                nx *= mat->normalFactorScale; ny *= mat->normalFactorScale;
                const v3 q = v3norm(v3_(nx, ny, nz));
                /* mul(normalize(xyz), TBN): row vector times the matrix whose rows are T, B, N */
                pnrm = v3_(q.x * tRS.x + q.y * bRS.x + q.z * vnrm.x, q.x * tRS.y + q.y * bRS.y + q.z * vnrm.y, q.x * tRS.z + q.y * bRS.z + q.z * vnrm.z);
            }
            float rough = mat->roughnessFactor;
            float metal = mat->metallicFactor >= 1.0f ? 0.0f : mat->metallicFactor;      /* getFallbackMetallic, gltf.h:53-58 */
            float ao = 1.0f;
            if (texMR != NULL && o.datas->uv0 != NULL) {
                float mr[4];
                oracle_sample_rgba(texMR, smp_lookup(f, mat->metallicRoughnessSampler), mu, mv, mudx, mvdx, mudy, mvdy, mr);
                rough = mr[1]; metal = mr[2];
                ao = mat->bExistOcclusion ? mat->occlusionTextureStrength * mr[0] : 1.0f;
            }
            /* gltfMetallicRoughnessPBR_Lighting (lighting.hlsl:203-268) */
            const v3 f0 = v3_(0.04f, 0.04f, 0.04f);
            v3 diffuseColor = v3s(v3mul(base, v3_(1.0f - f0.x, 1.0f - f0.y, 1.0f - f0.z)), 1.0f - metal);
            v3 specColor = v3_(f0.x + (base.x - f0.x) * metal, f0.y + (base.y - f0.y) * metal, f0.z + (base.z - f0.z) * metal);
            const float reflectance = fmaxf(fmaxf(specColor.x, specColor.y), specColor.z);
            const float r90 = clampf(reflectance * 50.0f, 0.0f, 1.0f);
            v3 pix2cam = v3norm(posR);
            const double ckx = storage_to_double(view->cameraPositionWS_km.x), cky = storage_to_double(view->cameraPositionWS_km.y), ckz = storage_to_double(view->cameraPositionWS_km.z);
            v3 pkm = v3_((float)((double)(posR.x / 1000.0f) + ckx), (float)((double)(posR.y / 1000.0f) + cky), (float)((double)(posR.z / 1000.0f) + ckz));
            v3 sunDirTo = v3_(-view->basicData.sunInfo.direction.x, -view->basicData.sunInfo.direction.y, -view->basicData.sunInfo.direction.z);
            v3 sky;
            v3 sun = sun_and_sky_irradiance(at, trLut, irLut, ap1, v3sub(pkm, v3_(at->earthCenterKm.x, at->earthCenterKm.y, at->earthCenterKm.z)), pnrm, sunDirTo, &sky);
            v3 radiance = v3add(sun, sky);
            /* evaluateDirectionalDirectLight (lighting.hlsli:166-184): pointToLight = -direction, view = -pixelToCamera */
            v3 n = v3norm(pnrm), vv = v3norm(v3s(pix2cam, -1.0f)), l = v3norm(sunDirTo), hh = v3norm(v3add(l, vv));
            const float NdotL = clampf(v3dot(n, l), 0, 1), NdotV = clampf(v3dot(n, vv), 0, 1), NdotH = clampf(v3dot(n, hh), 0, 1), VdotH = clampf(v3dot(vv, hh), 0, 1);
            v3 diff = v3_(0, 0, 0), spec = v3_(0, 0, 0);
            if (NdotL > 0.0f || NdotV > 0.0f) {
                float om = clampf(1.0f - VdotH, 0.0f, 1.0f); float om2 = om * om; float p5 = om2 * om2 * om;
                v3 F = v3_(specColor.x + (r90 - specColor.x) * p5, specColor.y + (r90 - specColor.y) * p5, specColor.z + (r90 - specColor.z) * p5);
                float a2 = rough * rough;
                /* V_SmithGGXCorrelated(NdotL, NdotV, roughness): first parameter is named NoV in bsdf.hlsli:30 */
                float GGXV = NdotV * sqrtf(NdotL * NdotL * (1.0f - a2) + a2);
                float GGXL = NdotL * sqrtf(NdotV * NdotV * (1.0f - a2) + a2);
                float GGX = GGXV + GGXL;
                float Vis = GGX > 0.0f ? 0.5f / GGX : 0.0f;
                float ff = (NdotH * a2 - NdotH) * NdotH + 1.0f;
                float D = a2 / (3.14159265358979323846f * ff * ff + 0.000001f);
                v3 lam = v3s(diffuseColor, 1.0f / 3.14159265358979323846f);
                v3 dc = v3mul(v3_(1.0f - F.x, 1.0f - F.y, 1.0f - F.z), lam);
                v3 sc = v3s(F, Vis * D);
                diff = v3s(dc, NdotL); spec = v3s(sc, NdotL);
            }
            diff = v3mul(diff, radiance); spec = v3mul(spec, radiance);
            const float expoK = at->luminanceMode == 0 ? 1.0f : 1e-5f;           /* finalRadianceExposureModify: 10.0 * radiance * k, left to right */
            diff = v3s(v3s(diff, 10.0f), expoK); spec = v3s(v3s(spec, 10.0f), expoK);
            v3 color = v3add(diff, spec);
            if (isnan(color.x) || isnan(color.y) || isnan(color.z)) color = v3_(0, 0, 0);
            float* o0 = planes[0] + pi * 4; o0[0] = color.x; o0[1] = color.y; o0[2] = color.z; o0[3] = 0.0f;
            float* o1 = planes[1] + pi * 4; o1[0] = base.x; o1[1] = base.y; o1[2] = base.z; o1[3] = 0.0f;
            float* o2 = planes[2] + pi * 4; o2[0] = vnrm.x * 0.5f + 0.5f; o2[1] = vnrm.y * 0.5f + 0.5f; o2[2] = vnrm.z * 0.5f + 0.5f; o2[3] = 1.0f;
            float* o3 = planes[3] + pi * 4; o3[0] = pnrm.x * 0.5f + 0.5f; o3[1] = pnrm.y * 0.5f + 0.5f; o3[2] = pnrm.z * 0.5f + 0.5f; o3[3] = 1.0f;
            float* o4 = planes[4] + pi * 4; o4[0] = mvx; o4[1] = mvy; o4[2] = 0.0f; o4[3] = 0.0f;
            float* o5 = planes[5] + pi * 4; o5[0] = ao; o5[1] = rough; o5[2] = metal; o5[3] = 0.0f;
        }
}
