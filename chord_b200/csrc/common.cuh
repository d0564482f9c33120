// common.cuh — shared device/host declarations for libchord_b200 (sm_100a only).
//
// Numerics contract: the whole library is compiled with --fmad=false, so the ONLY fused
// multiply-adds are the explicit __fmaf_rn calls below; divisions and square roots are IEEE round-to-nearest
// (__fdiv_rn / __fsqrt_rn); min/max are fminf/fmaxf.  Dot products and matrix rows accumulate left to right.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/chord_b200.h"

namespace chordb {

// ------------------------------------------------------------------------------------------------ device scene
struct DatasDev {
    const chord_GPUGLTFMeshlet*      meshlets;
    const float*                     positions;     // float3
    const float*                     normals;       // float3
    const float*                     uv0;           // float2
    const float*                     tangents;      // float4
    const uint32_t*                  meshletData;
    const chord_GPUGLTFMeshletGroup* groups;
    const uint32_t*                  groupIndices;
    const chord_GPUBVHNode*          bvhNodes;
    // built by chord_scene_upload for the rasteriser (raster.cu: build_meshlet_packs): per meshlet ONE contiguous block
    // [triangle words | float3 positions in meshlet order], its offset in uint4 units, and the vertex base it was gathered with
    const uint4*                     packBlocks;
    const uint32_t*                  packOffset;
    const uint32_t*                  packVertexBase;
};

struct SceneDev {
    const DatasDev*                  datas;
    const chord_GLTFPrimitiveBuffer* prims;
    const chord_GLTFMaterialGPUData* mats;
    uint32_t nDatas, nPrims, nMats;
};

// Per-object derived matrices, computed once per frame by k_object_prepare (same arithmetic the per-thread
// HLSL does: mul(translatedWorldToClip, localToTranslatedWorld) etc.).
struct ObjectDerived {
    chord_float4x4 localToClip;        // iview.translatedWorldToClip * localToTranslatedWorld
    chord_float4x4 localToView;        // view.translatedWorldToView  * localToTranslatedWorld
    chord_float4x4 localToClipHzb0;    // view.translatedWorldToClipLastFrame * localToTranslatedWorldLastFrame
    chord_float4x4 localToClipHzb1;    // view.translatedWorldToClip * localToTranslatedWorld
    chord_float4x4 localToClipNoJitter;      // view.translatedWorldToClip_NoJitter * localToTranslatedWorld
    chord_float4x4 localToClipLastNoJitter;  // view.translatedWorldToClipLastFrame_NoJitter * localToTranslatedWorldLastFrame
};

// Per-object record of what the cull kernels reach through object -> primitive -> datas (three dependent loads per thread in
// the reference's shaders): resolved once per object and frame by k_object_prepare, read with ONE load by k_group_cull and
// the HZB cull kernels (their run time is a chain of dependent loads: 14 us / 9 us for a 4K city frame).
struct ObjectResolved {
    const chord_GPUGLTFMeshlet*      meshlets;          // datas.meshlets
    const chord_GPUGLTFMeshletGroup* groups;            // datas.groups + prim.meshletGroupOffset
    const uint32_t*                  groupIndices;      // datas.groupIndices + prim.meshletGroupIndicesOffset
    uint32_t meshletOffset;                             // prim.meshletOffset
    uint32_t twoSided;                                  // material.bTwoSided
    uint32_t groupCount;                                // prim.meshletGroupCount
    uint32_t pad;
};

// material texture table of the masked raster bucket (chord_textures_upload): alpha channel only, A8 mip chain
// alpha: A8 chain for the masked raster bucket; rgba: RGBA8 chain for the resolve; both nullptr: builtin white
struct TextureDev { const uint8_t* alpha; const uint8_t* rgba; uint32_t width, height, mipCount, flags; };
struct SamplerDev { uint32_t addressU, addressV; };

struct FrameDev {
    SceneDev scene;
    const TextureDev* textures; const SamplerDev* samplers; uint32_t nTextures, nSamplers;
    const chord_PerframeCameraView*       view;
    const chord_InstanceCullingViewInfo*  iview;
    const chord_GPUObjectGLTFPrimitive*   objects;
    const ObjectDerived*                  derived;
    const ObjectResolved*                 resolved;
    uint32_t objectCount;
    uint32_t width, height;
    chord_rect tile;
    uint32_t switchFlags;
    uint32_t passFlags;                 // CHORD_PASS_* of the raster call in flight (renderMeshDepth)
    float tanHalfFovy;                  // tanf(0.5f * cameraFovy) evaluated on the host (libm), see DESIGN.md §3
    unsigned long long* vis64;          // width*height packed texels
};

// ------------------------------------------------------------------------------------------------ multi-GPU exchange
// One block of this layout per rank (one cudaMalloc, exported with CUDA IPC or reached through peer access): the header, then
// the two history HZBs (min + max chains each).  Peers write into it: their part of mips 0..5, their valid range, their
// arrival flag (hzb.cu: k_hzb_exchange_put / k_hzb_exchange_wait_top).
#define CHORD_MAX_EXCHANGE_RANKS 16
struct ExchangeHeader {
    uint32_t arrived[CHORD_MAX_EXCHANGE_RANKS];          // [rank] = last frame sequence number whose stores into THIS block are complete
    uint32_t localSeq; uint32_t pad[15];                 // this rank's own sequence number (advanced on the device: graph replays)
    uint32_t ranges[2][CHORD_MAX_EXCHANGE_RANKS][2];     // [history slot][rank] = {minBits, maxBits} of that rank's tile
};
struct ExchangePeers {                                   // device-resident table: where this rank's texels / flags go on every peer
    __half* min[CHORD_MAX_EXCHANGE_RANKS - 1]; __half* max[CHORD_MAX_EXCHANGE_RANKS - 1];
    ExchangeHeader* header[CHORD_MAX_EXCHANGE_RANKS - 1];
};
struct HzbExchange {                                     // one build's exchange parameters (renderer.cu fills it)
    uint32_t nPeers, rank, parity;
    const ExchangePeers* dPeers;                         // table of the history slot `parity`
    ExchangeHeader* localHeader;
};

// ------------------------------------------------------------------------------------------------ math
struct f4 { float x, y, z, w; };
struct f3 { float x, y, z; };

__device__ __forceinline__ float dot3f(float ax, float ay, float az, float bx, float by, float bz)
{ float r = __fmul_rn(ax, bx); r = __fmaf_rn(ay, by, r); r = __fmaf_rn(az, bz, r); return r; }
__device__ __forceinline__ float dot4f(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw)
{ float r = __fmul_rn(ax, bx); r = __fmaf_rn(ay, by, r); r = __fmaf_rn(az, bz, r); r = __fmaf_rn(aw, bw, r); return r; }

__device__ __forceinline__ f4 mat_mul_vec(const float* __restrict__ m, float x, float y, float z, float w)
{
    f4 o;
    o.x = dot4f(m[0], m[4], m[8],  m[12], x, y, z, w);
    o.y = dot4f(m[1], m[5], m[9],  m[13], x, y, z, w);
    o.z = dot4f(m[2], m[6], m[10], m[14], x, y, z, w);
    o.w = dot4f(m[3], m[7], m[11], m[15], x, y, z, w);
    return o;
}

__device__ __forceinline__ void mat_mul(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Cm)
{
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++)
            Cm[c * 4 + r] = dot4f(A[r], A[4 + r], A[8 + r], A[12 + r], B[c * 4], B[c * 4 + 1], B[c * 4 + 2], B[c * 4 + 3]);
}

// base.hlsli:437-447 (device copies of chord_encode/decode_triangle_instance)
__device__ __forceinline__ uint32_t enc_vis(uint32_t tri, uint32_t inst) { return (((inst + 1u) & CHORD_MAX_INSTANCE_ID_COUNT) << 8) | (tri & 0xFFu); }
__device__ __forceinline__ void dec_vis(uint32_t pack, uint32_t& tri, uint32_t& inst) { tri = pack & 0xFFu; inst = ((pack >> 8) & CHORD_MAX_INSTANCE_ID_COUNT) - 1u; }

// Scheduling fence for four independent dependent-load chains: the four values must be in flight (their loads issued) before
// anything after this point is issued; no instruction is emitted.  Without it the compiler interleaves the chains so that
// chain 2's first load waits behind chain 1's second round trip.
__device__ __forceinline__ void loads_issued(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d)
{ asm volatile("" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); }

__device__ __forceinline__ bool has_flag(uint32_t flags, uint32_t bit) { return (flags & (1u << bit)) != 0u; }

__device__ __constant__ const float kExtentApplyFactor[8][3] = {
    { 1,  1,  1}, {-1, -1, -1}, { 1,  1, -1}, { 1, -1,  1},
    {-1,  1,  1}, { 1, -1, -1}, {-1, -1,  1}, {-1,  1, -1},
};

// ------------------------------------------------------------------------------------------------ ordered compaction
// Single-pass decoupled look-back scan state: one 64-bit word per tile, flag in bits 62..63, value below.
constexpr unsigned long long kScanFlagAggregate = 1ull << 62;
constexpr unsigned long long kScanFlagPrefix    = 2ull << 62;
constexpr unsigned long long kScanValueMask     = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{ unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{ asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// Called by all 32 lanes of ONE warp of the block that owns `tile`; returns the exclusive prefix of the tile
// (sum of the aggregates of all lower tiles) and publishes this tile's inclusive prefix.
__device__ __forceinline__ unsigned long long scan_tile_exclusive_prefix(unsigned long long* status, uint32_t tile,
                                                                         unsigned long long aggregate)
{
    const uint32_t lane = threadIdx.x & 31u;
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(&status[0], kScanFlagPrefix | aggregate);
        return 0ull;
    }
    if (lane == 0) st_volatile_u64(&status[tile], kScanFlagAggregate | aggregate);
    unsigned long long exclusive = 0ull;
    int look = (int)tile - 1;
    while (true) {
        const int idx = look - (int)lane;
        unsigned long long v = (idx >= 0) ? ld_volatile_u64(&status[idx]) : (kScanFlagPrefix | 0ull);
        const uint32_t flag = (uint32_t)(v >> 62);
        if (__any_sync(0xffffffffu, flag == 0u)) continue;               // a predecessor has not published yet
        const uint32_t pmask = __ballot_sync(0xffffffffu, flag == 2u);
        unsigned long long val = v & kScanValueMask;
        if (pmask) {
            const uint32_t first = (uint32_t)__ffs((int)pmask) - 1u;    // nearest tile with an inclusive prefix
            if (lane > first) val = 0ull;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        exclusive += val;
        if (pmask) break;
        look -= 32;
    }
    if (lane == 0) st_volatile_u64(&status[tile], kScanFlagPrefix | (exclusive + aggregate));
    return exclusive;
}

// ------------------------------------------------------------------------------------------------ host helpers
void set_error(const char* fmt, ...);
#define CHORD_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
    chordb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); return CHORD_ERR_CUDA; } } while (0)
// A captured call must not (re)allocate or synchronise (either would invalidate the stream capture, or worse, free memory an
// already-recorded node points at): the buffers are sized by the eager run the capture rules demand (chord_b200.h).
#define CHORD_NO_RESIZE_IN_CAPTURE(ctx, what) do { if ((ctx)->capturing) { \
    chordb::set_error("%s would have to be (re)allocated while a graph is being captured: run the same call sequence once eagerly before chord_graph_begin_capture", what); \
    return CHORD_ERR_CAPACITY; } } while (0)
#define CHORD_REQUIRE(cond, msg) do { if (!(cond)) { chordb::set_error("check(%s) failed: %s", #cond, msg); return CHORD_ERR_INVALID; } } while (0)

}  // namespace chordb

// ------------------------------------------------------------------------------------------------ opaque handles
struct chord_graph_frame_inputs { chord_frame* frame; unsigned char* pinned; uint32_t objectCount; float tanHalfFovy; };

struct chord_ctx {
    int refs;                           // owner + one per live child handle (ctx.cu: ctx_retain / ctx_release)
    bool ownerReleased;                 // chord_ctx_destroy has been called; children keep the context alive
    int device;
    cudaStream_t stream;
    cudaStream_t uploadStream;          // created by the first chord_cmdlist_upload_async
    uint32_t switchFlags;
    uint32_t lightingSlot;              // this context's slot of the resolve kernel's constant-bank uniforms (resolve.cu)
    int rasterVariant, rasterCtasPerSm; uint32_t rasterWarpMin;    // chord_ctx_set_raster_tuning (defaults 11, 0 = built-in, 16384)
    uint64_t launches;
    int smCount;
    // scene
    chordb::SceneDev scene;
    void** sceneAllocs; uint32_t nSceneAllocs;
    chordb::DatasDev* datasHost;
    uint32_t* primGroupCountHost;       // prims[i].meshletGroupCount, for the host-side check of chord_frame_begin's totals
    // scratch for ordered compaction
    unsigned long long* scanBlock;       // [ticket | status...], see scan_reset
    unsigned long long* scanStatus; uint32_t scanTiles;
    uint32_t* scanTicket;
    // atmosphere LUTs
    __half* lutTransmittance; __half* lutIrradiance;
    __half* lutScattering; __half* lutSingleMie;      // 256 x 128 x 32 RGBA16F volumes of the sky branch (null: no sky branch)
    // graph capture (chord_graph_begin_capture .. end_capture): pinned snapshots made while capturing
    bool capturing; void** captureAllocs; uint32_t nCaptureAllocs;
    chord_graph_frame_inputs* captureFrames; uint32_t nCaptureFrames;     // chord_frame_begin calls recorded by the capture in flight
    // material textures (masked bucket)
    chordb::TextureDev* dTextures; chordb::SamplerDev* dSamplers; uint8_t* dTexels; uint8_t* dTexelsRgba; uint32_t nTextures, nSamplers;
    float* dSrgbLut;                    // 256 floats, sRGB byte -> linear (CHORD_TEXTURE_SRGB texel decode in the resolve)
};

struct chord_cascade_views {
    chord_ctx* ctx;
    chord_InstanceCullingViewInfo* dViews;      // [CHORD_MAX_CASCADE_COUNT]
    uint32_t cascadeDim;                        // of the last setup / upload (frame size check)
};

struct chord_graph {
    chord_ctx* ctx;
    cudaGraph_t graph; cudaGraphExec_t exec;
    uint32_t kernelNodes;
    void** pinned; uint32_t nPinned;
    chord_graph_frame_inputs* frames; uint32_t nFrames;
};

struct chord_cmdlist {
    chord_ctx* ctx;
    uint32_t capacity;
    uint32_t* count;            // device
    uint32_t* pinnedCount;      // host-pinned ring for async uploads
    uint32_t pinnedIdx;
    chord_draw_cmd* cmds;       // device
    // overlapped uploads (chord_cmdlist_upload_async runs on the context's upload stream): `uploaded` gates the next
    // consumer on the main stream, `lastUse` (recorded after every consumer) gates the next upload (write after read)
    cudaEvent_t uploaded, lastUse; bool uploadPending, asyncUser;
};

struct chord_hzb {
    chord_ctx* ctx;
    uint32_t depthW, depthH, mip0W, mip0H, mipCount;
    uint64_t mipOffset[CHORD_HZB_MAX_MIP_COUNT];   // in texels
    uint64_t totalTexels;
    __half* minChain; __half* maxChain;
    uint32_t* validRange;        // {minBits, maxBits}
    uint32_t* counter;
    int valid;
    bool externalChains;         // chains live in a renderer's exchange block (not freed with the HZB)
};

struct chord_frame {
    chord_ctx* ctx;
    uint32_t width, height;
    chord_rect tile;
    unsigned long long* vis64;
    // per-frame inputs: ONE device block [view | iview | objects...] fed by ONE H2D copy per chord_frame_begin from a
    // pinned staging ring (dView / dIView / dObjects point into it)
    unsigned char* dInputs; unsigned char* hInputs[4]; cudaEvent_t inputsCopied[4]; uint32_t inputsSlot; size_t inputsBytes;
    chord_PerframeCameraView* dView; chord_InstanceCullingViewInfo* dIView; void* dLightingView; size_t lightingViewOffset, objectsOffset;
    chord_GPUObjectGLTFPrimitive* dObjects; uint32_t objectCapacity, objectCount;
    chord_rect clearedTile; bool clearedValid;     // region that may hold non-zero texels (tile-only clears)
    chordb::ObjectDerived* dDerived; chordb::ObjectResolved* dResolved;
    uint32_t lod0MeshletCount, meshletGroupCount;
    float tanHalfFovy;
    uint32_t passFlags; bool hIViewOrtho;
    // instance culling scratch
    uint32_t* objGroupCount; uint32_t* objGroupBase; uint32_t* totalGroups;
    void* bvhEntries; uint32_t* bvhBits; uint32_t* bvhCtrl; uint32_t bvhCapacity, bvhWords;
    // raster
    void* largeTris; uint32_t* largeCounters; uint32_t largeTriCapacity, largeItemCapacity; void* largeItems;
    unsigned long long* stats;   // device chord_raster_stats
    void* works; uint32_t worksCapacity; uint32_t* redoList; uint32_t* maskedList; void* deferred; uint32_t deferredCapacity;
    // resolve
    uint32_t* tileMarker; uint32_t* tileList; uint32_t* tileCount; float* gbuffer[CHORD_GBUF_COUNT];
    void* gbufferPackedBuf[CHORD_GBUF_COUNT]; int gbufferPacked;
    chord_rect gbufTile; bool gbufValid; int gbufPackedMode;     // rect the resolve planes were last fully written for
    void* resolveRecs; uint32_t resolveCapacity;
    // pipelined read-back of the R32_UINT image: two staging slots, a copy stream, events
    uint32_t* splitVis[2]; cudaStream_t copyStream; bool copyStreamShared; cudaEvent_t splitDone[2], copyDone[2]; int splitIdx;
    cudaStream_t ownCopyStream;       // chord_frame_set_readback_stream(frame, 1): read-backs on a stream of their own
    chord_PerframeCameraView hView;   // host copy
    bool begun;
};

namespace chordb {
FrameDev make_frame_dev(const chord_frame* f);
// kernels' host launchers (each returns a CHORD_* status)
int launch_object_prepare(chord_frame* f);
int launch_instance_culling(chord_frame* f, chord_cmdlist* out);
int launch_instance_culling_bvh(chord_frame* f, chord_cmdlist* out);
int launch_hzb_culling(chord_frame* f, const chord_hzb* hzb, int firstStage, chord_cmdlist* in, chord_cmdlist* outVis, chord_cmdlist* outRej);
int launch_render_mesh(chord_frame* f, chord_cmdlist* list);
int launch_cascade_setup(chord_frame* mainFrame, const chord_cascade_setup_desc* desc, const chord_hzb* validRangeSource, chord_InstanceCullingViewInfo* dViews);
int launch_hzb_culling_generic(chord_frame* f, const chord_hzb* hzb, const chord_InstanceCullingViewInfo* hzbView, float extentScale,
                               int useLastFrame, chord_cmdlist* in, chord_cmdlist* outVis, const chord_InstanceCullingViewInfo* dHzbView = nullptr);
int launch_build_hzb(chord_frame* f, int bMin, int bMax, int bRange, chord_hzb* dst, const HzbExchange* exchange = nullptr, bool deferPut = false);
int launch_hzb_exchange_put(chord_frame* f, int bRange, chord_hzb* dst, const HzbExchange* exchange, uint32_t seq);   // seq 0: the kernel advances the sequence number itself
int launch_hzb_range_reduce(chord_ctx* ctx, const uint32_t* pairs, uint32_t world, uint32_t* validRange);
int launch_hzb_exchange_wait_top(chord_frame* f, chord_hzb* dst, const HzbExchange* exchange, uint32_t world, uint32_t* status, uint32_t maxPolls);
int launch_visibility_mark(chord_frame* f, chord_cmdlist* post);
enum { CHORD_LIGHTING_PREPARE = 1u, CHORD_LIGHTING_SHADE = 2u };
int launch_lighting(chord_frame* f, chord_cmdlist* post, uint32_t parts = CHORD_LIGHTING_PREPARE | CHORD_LIGHTING_SHADE);
int launch_split_vis(chord_frame* f, uint32_t* vis32, float* depth32);
int launch_decode(chord_frame* f, chord_cmdlist* post, uint32_t* triples);
int launch_selftest_vertex_records(chord_ctx* ctx, uint32_t n, uint32_t seed, uint32_t width, uint32_t height, uint32_t clampZ, unsigned long long* dOut);
int scan_reset(chord_ctx* ctx, uint32_t tilesNeeded);
int launch_fill_u32(chord_ctx* ctx, void* dst, uint32_t v0, uint32_t v1, size_t words);    // ctx.cu: small fills as kernels (cheaper graph nodes than memsets)
int build_meshlet_packs(chord_ctx* ctx, DatasDev* d, uint32_t datasId, uint32_t meshletCount, uint32_t vertexCount, uint32_t groupCount, uint32_t groupIndicesCount,
                        const chord_GLTFPrimitiveBuffer* dPrims, uint32_t nPrims, void** outAllocs);
size_t lighting_view_bytes();                                                     // resolve.cu: sizeof(LightingView)
void fill_lighting_view(const chord_PerframeCameraView& view, void* dst);         // host-side derivation of the resolve uniforms
// one frame's inputs, staged as ONE block [view | instance view | LightingView | objects ...] (chord_frame_begin)
void fill_frame_inputs(const chord_frame* f, unsigned char* dst, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                       const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount);
void ctx_retain(chord_ctx* ctx);
void ctx_release(chord_ctx* ctx);
}
