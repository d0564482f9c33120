// ctx.cu — context / scene / list / HZB / frame objects and the C-ABI entry points of libchord_b200.
// Host-side only; every pass body lives in cull.cu / raster.cu / hzb.cu / resolve.cu.
#include "common.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace chordb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

FrameDev make_frame_dev(const chord_frame* f)
{
    FrameDev d;
    d.scene = f->ctx->scene;
    d.view = f->dView; d.iview = f->dIView; d.objects = f->dObjects; d.derived = f->dDerived; d.resolved = f->dResolved;
    d.objectCount = f->objectCount;
    d.width = f->width; d.height = f->height; d.tile = f->tile;
    d.switchFlags = f->ctx->switchFlags;
    d.passFlags = f->passFlags;
    d.tanHalfFovy = f->tanHalfFovy;
    d.vis64 = f->vis64;
    d.textures = f->ctx->dTextures; d.samplers = f->ctx->dSamplers; d.nTextures = f->ctx->nTextures; d.nSamplers = f->ctx->nSamplers;
    return d;
}

// Small fills as KERNELS, not memset / memcpy nodes: inside a CUDA graph on a B200 a dependent empty kernel costs 1.2 us, a
// 32-byte memset node 2.9 us and an 8-byte host-to-device copy node 4.5 us (tests/micro/memset_nodes.cu) - and a frame has
// about ten of them (compaction scratch, queue counters, tile count, valid-range seed).
__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* __restrict__ dst, uint32_t v0, uint32_t v1, uint32_t n)
{
    // words alternate v0, v1 (v0 == v1: plain fill; {~0u, 0u}: the {min, max} seed of a valid depth range)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = (i & 1u) ? v1 : v0;
}

int launch_fill_u32(chord_ctx* ctx, void* dst, uint32_t v0, uint32_t v1, size_t words)
{
    if (!words) return CHORD_OK;
    uint32_t grid = (uint32_t)((words + 255) / 256); if (grid > (uint32_t)ctx->smCount * 8u) grid = (uint32_t)ctx->smCount * 8u;
    k_fill_u32<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(dst), v0, v1, (uint32_t)words);
    ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

// Scratch of the ordered compactions: ONE allocation [ticket word | status[0 .. tiles]] so that one fill per cull
// kernel resets both (two stream operations per cull call, four cull calls per frame, were ~2 us each on the timeline).
int scan_reset(chord_ctx* ctx, uint32_t tilesNeeded)
{
    if (tilesNeeded > ctx->scanTiles || !ctx->scanBlock) {
        CHORD_NO_RESIZE_IN_CAPTURE(ctx, "the compaction scratch");
        if (ctx->scanBlock) cudaFree(ctx->scanBlock);                           // (synchronises the device)
        ctx->scanBlock = nullptr; ctx->scanStatus = nullptr; ctx->scanTicket = nullptr; ctx->scanTiles = 0;
        uint32_t n = tilesNeeded + 1024;
        CHORD_CUDA(cudaMalloc(&ctx->scanBlock, sizeof(unsigned long long) * ((size_t)n + 1)));
        ctx->scanTicket = reinterpret_cast<uint32_t*>(ctx->scanBlock);        // low half of word 0
        ctx->scanStatus = ctx->scanBlock + 1;
        ctx->scanTiles = n;
    }
    return launch_fill_u32(ctx, ctx->scanBlock, 0u, 0u, 2 * ((size_t)tilesNeeded + 2));
}

template <typename T>
static int upload(chord_ctx* ctx, std::vector<void*>& allocs, const T* host, size_t count, const T** out, size_t padBytes = 64)
{
    *out = nullptr;
    if (!host || count == 0) return CHORD_OK;
    void* d = nullptr;
    const size_t bytes = sizeof(T) * count;
    CHORD_CUDA(cudaMalloc(&d, bytes + padBytes));                 // pad: kernels may read aligned supersets
    CHORD_CUDA(cudaMemsetAsync((char*)d + bytes, 0, padBytes, ctx->stream));
    CHORD_CUDA(cudaMemcpyAsync(d, host, bytes, cudaMemcpyDefault, ctx->stream));   // host or device source (UVA)
    allocs.push_back(d);
    *out = (const T*)d;
    return CHORD_OK;
}

// Handle lifetime (reference: pool refs kept alive by CommandBuffer::addReferenceResource, command_list.h:23-38,
// render_helper.h:411-451): every child object (list, HZB, frame, graph, cascade views) holds one reference on its
// context; chord_ctx_destroy only drops the owner's reference, the context is torn down when the last child is gone.
void ctx_retain(chord_ctx* ctx) { __atomic_add_fetch(&ctx->refs, 1, __ATOMIC_ACQ_REL); }

static void free_scene(chord_ctx* ctx);
static void ctx_teardown(chord_ctx* ctx);
void ctx_release(chord_ctx* ctx)
{
    if (__atomic_sub_fetch(&ctx->refs, 1, __ATOMIC_ACQ_REL) == 0) ctx_teardown(ctx);
}

static void free_scene(chord_ctx* ctx)
{
    for (uint32_t i = 0; i < ctx->nSceneAllocs; i++) cudaFree(ctx->sceneAllocs[i]);
    free(ctx->sceneAllocs); ctx->sceneAllocs = nullptr; ctx->nSceneAllocs = 0;
    free(ctx->datasHost); ctx->datasHost = nullptr;
    free(ctx->primGroupCountHost); ctx->primGroupCountHost = nullptr;
    memset(&ctx->scene, 0, sizeof(ctx->scene));
}

void fill_frame_inputs(const chord_frame* f, unsigned char* dst, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                       const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount)
{
    memcpy(dst, view, sizeof(*view));
    memcpy(dst + sizeof(chord_PerframeCameraView), iview, sizeof(*iview));
    fill_lighting_view(*view, dst + f->lightingViewOffset);
    memcpy(dst + f->objectsOffset, objects, sizeof(chord_GPUObjectGLTFPrimitive) * (size_t)objectCount);
}

static void ctx_teardown(chord_ctx* ctx)
{
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_scene(ctx);
    if (ctx->scanBlock) cudaFree(ctx->scanBlock);
    if (ctx->lutTransmittance) cudaFree(ctx->lutTransmittance);
    if (ctx->lutIrradiance) cudaFree(ctx->lutIrradiance);
    cudaFree(ctx->lutScattering); cudaFree(ctx->lutSingleMie);
    cudaFree(ctx->dTextures); cudaFree(ctx->dSamplers); cudaFree(ctx->dTexels); cudaFree(ctx->dTexelsRgba); cudaFree(ctx->dSrgbLut);
    if (ctx->uploadStream) { cudaStreamSynchronize(ctx->uploadStream); cudaStreamDestroy(ctx->uploadStream); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    free(ctx);
}

}  // namespace chordb

using namespace chordb;

extern "C" {

const char* chord_last_error(void) { return g_error; }

int chord_abi_version(void) { return CHORD_B200_ABI_VERSION; }

int chord_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int chord_ctx_create(int device, chord_ctx** out)
{
    CHORD_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { set_error("no CUDA device visible"); return CHORD_ERR_NO_DEVICE; }
    CHORD_REQUIRE(device >= 0 && device < n, "device index");
    CHORD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CHORD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("libchord_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor); return CHORD_ERR_NO_DEVICE; }
    chord_ctx* c = (chord_ctx*)calloc(1, sizeof(chord_ctx));
    CHORD_REQUIRE(c != nullptr, "out of host memory");
    c->device = device;
    c->refs = 1;                                                   // the owner's reference, dropped by chord_ctx_destroy
    c->switchFlags = 0b111;
    { static int nextSlot = 0; c->lightingSlot = (uint32_t)(__atomic_fetch_add(&nextSlot, 1, __ATOMIC_RELAXED) & 3); }
    c->rasterVariant = 11; c->rasterCtasPerSm = 0; c->rasterWarpMin = 16384u;
    c->smCount = prop.multiProcessorCount;
    const int rc = [&]() -> int {
        // greatest priority: a renderer's asynchronous resolve stream (least priority) yields to the passes enqueued here
        int leastPrio = 0, greatestPrio = 0;
        CHORD_CUDA(cudaDeviceGetStreamPriorityRange(&leastPrio, &greatestPrio));
        CHORD_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatestPrio));
        // sRGB byte -> linear float (IEC 61966-2-1), the texel decode a *_SRGB image format performs; double precision on the
        // host, rounded once (the oracle builds the identical table)
        float lut[256];
        for (int i = 0; i < 256; i++) { const double v = (double)i / 255.0; lut[i] = (float)(v <= 0.04045 ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4)); }
        CHORD_CUDA(cudaMalloc(&c->dSrgbLut, sizeof(lut)));
        CHORD_CUDA(cudaMemcpy(c->dSrgbLut, lut, sizeof(lut), cudaMemcpyHostToDevice));
        return CHORD_OK;
    }();
    if (rc != CHORD_OK) { chord_ctx_destroy(c); return rc; }      // nothing half-built survives a failed create
    *out = c;
    return CHORD_OK;
}

void chord_ctx_destroy(chord_ctx* ctx)
{
    // Drops the owner's reference.  Children created from the context keep it (stream, scene, scratch) alive until the last
    // of them is destroyed, in any order; a second chord_ctx_destroy on the same handle is a caller error, as with free().
    if (!ctx) return;
    if (ctx->ownerReleased) return;                                // tolerated while children still hold the context
    ctx->ownerReleased = true;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx_release(ctx);
}
uint32_t chord_ctx_refcount(const chord_ctx* ctx) { return ctx ? (uint32_t)__atomic_load_n(&ctx->refs, __ATOMIC_ACQUIRE) : 0u; }

void* chord_ctx_stream(chord_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int chord_ctx_sync(chord_ctx* ctx)
{
    CHORD_REQUIRE(ctx, "ctx");
    if (ctx->uploadStream) CHORD_CUDA(cudaStreamSynchronize(ctx->uploadStream));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    return CHORD_OK;
}
int chord_ctx_set_switch_flags(chord_ctx* ctx, uint32_t f) { CHORD_REQUIRE(ctx, "ctx"); ctx->switchFlags = f & 7u; return CHORD_OK; }
uint32_t chord_ctx_get_switch_flags(const chord_ctx* ctx) { return ctx ? ctx->switchFlags : 0; }
int chord_ctx_set_raster_tuning(chord_ctx* ctx, int variant, int ctasPerSm, int64_t warpMinMeshlets)
{
    CHORD_REQUIRE(ctx, "ctx");
    CHORD_REQUIRE(variant < 0 || variant == 0 || variant == 11 || variant == 13 || variant == 14, "variant (0, 11, 13 or 14)");
    CHORD_REQUIRE(ctasPerSm <= 32, "ctasPerSm");
    if (variant >= 0) ctx->rasterVariant = variant;
    if (ctasPerSm >= 0) ctx->rasterCtasPerSm = ctasPerSm;
    if (warpMinMeshlets >= 0) ctx->rasterWarpMin = warpMinMeshlets > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)warpMinMeshlets;
    return CHORD_OK;
}
uint64_t chord_ctx_kernel_launches(const chord_ctx* ctx) { return ctx ? ctx->launches : 0; }

int chord_scene_upload(chord_ctx* ctx, const chord_scene_desc* desc)
{
    CHORD_REQUIRE(ctx && desc, "null");
    CHORD_REQUIRE(desc->datasCount > 0 && desc->primitiveCount > 0 && desc->materialCount > 0, "empty scene");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    free_scene(ctx);
    std::vector<void*> allocs;
    ctx->datasHost = (DatasDev*)calloc(desc->datasCount, sizeof(DatasDev));
    int rc = CHORD_OK;
    for (uint32_t i = 0; i < desc->datasCount && rc == CHORD_OK; i++) {
        const chord_primitive_datas_desc& d = desc->datas[i];
        DatasDev& o = ctx->datasHost[i];
        if (!d.meshlets || !d.positions || !d.meshletData || !d.groups || !d.groupIndices) { set_error("datas[%u]: missing required array", i); rc = CHORD_ERR_INVALID; break; }
        if ((rc = upload(ctx, allocs, d.meshlets, d.meshletCount, &o.meshlets))) break;
        if ((rc = upload(ctx, allocs, d.positions, (size_t)d.vertexCount * 3, &o.positions))) break;
        if ((rc = upload(ctx, allocs, d.normals, (size_t)d.vertexCount * 3, &o.normals))) break;
        if ((rc = upload(ctx, allocs, d.uv0, (size_t)d.vertexCount * 2, &o.uv0))) break;
        if ((rc = upload(ctx, allocs, d.tangents, (size_t)d.vertexCount * 4, &o.tangents))) break;
        if ((rc = upload(ctx, allocs, d.meshletData, d.meshletDataCount, &o.meshletData))) break;
        if ((rc = upload(ctx, allocs, d.groups, d.groupCount, &o.groups))) break;
        if ((rc = upload(ctx, allocs, d.groupIndices, d.groupIndicesCount, &o.groupIndices))) break;
        if ((rc = upload(ctx, allocs, d.bvhNodes, d.bvhNodeCount, &o.bvhNodes))) break;
    }
    if (rc == CHORD_OK) rc = upload(ctx, allocs, desc->primitives, desc->primitiveCount, &ctx->scene.prims);
    for (uint32_t i = 0; i < desc->datasCount && rc == CHORD_OK; i++) {
        // the rasteriser's packed per-meshlet blocks (positions in meshlet order next to the triangle words)
        const chord_primitive_datas_desc& d = desc->datas[i];
        void* packAllocs[3] = { nullptr, nullptr, nullptr };
        rc = build_meshlet_packs(ctx, &ctx->datasHost[i], i, d.meshletCount, d.vertexCount, d.groupCount, d.groupIndicesCount, ctx->scene.prims, desc->primitiveCount, packAllocs);
        for (void* a : packAllocs) if (a) allocs.push_back(a);
    }
    if (rc == CHORD_OK) rc = upload(ctx, allocs, ctx->datasHost, desc->datasCount, &ctx->scene.datas);
    if (rc == CHORD_OK) rc = upload(ctx, allocs, desc->materials, desc->materialCount, &ctx->scene.mats);
    ctx->scene.nDatas = desc->datasCount; ctx->scene.nPrims = desc->primitiveCount; ctx->scene.nMats = desc->materialCount;
    ctx->primGroupCountHost = (uint32_t*)malloc(sizeof(uint32_t) * desc->primitiveCount);
    for (uint32_t i = 0; i < desc->primitiveCount; i++) ctx->primGroupCountHost[i] = desc->primitives[i].meshletGroupCount;
    ctx->sceneAllocs = (void**)malloc(sizeof(void*) * (allocs.size() + 1));
    ctx->nSceneAllocs = (uint32_t)allocs.size();
    for (size_t i = 0; i < allocs.size(); i++) ctx->sceneAllocs[i] = allocs[i];
    if (rc != CHORD_OK) { free_scene(ctx); return rc; }
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    return CHORD_OK;
}

// ---------------------------------------------------------------------------------------------- cmd lists
int chord_cmdlist_create(chord_ctx* ctx, uint32_t capacity, chord_cmdlist** out)
{
    CHORD_REQUIRE(ctx && out, "null");
    CHORD_REQUIRE(capacity > 0, "capacity");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    chord_cmdlist* l = (chord_cmdlist*)calloc(1, sizeof(chord_cmdlist));
    CHORD_REQUIRE(l != nullptr, "out of host memory");
    l->ctx = ctx; l->capacity = capacity; ctx_retain(ctx);
    const int rc = [&]() -> int {
        CHORD_CUDA(cudaMalloc(&l->count, 16));
        CHORD_CUDA(cudaMalloc(&l->cmds, sizeof(chord_draw_cmd) * (size_t)capacity + 64));
        CHORD_CUDA(cudaMemsetAsync(l->count, 0, 16, ctx->stream));
        return CHORD_OK;
    }();
    if (rc != CHORD_OK) { chord_cmdlist_destroy(l); return rc; }
    *out = l;
    return CHORD_OK;
}
void chord_cmdlist_destroy(chord_cmdlist* l)
{
    if (!l) return;
    cudaStreamSynchronize(l->ctx->stream);
    if (l->ctx->uploadStream) cudaStreamSynchronize(l->ctx->uploadStream);
    if (l->uploaded) cudaEventDestroy(l->uploaded);
    if (l->lastUse) cudaEventDestroy(l->lastUse);
    cudaFree(l->count); cudaFree(l->cmds); if (l->pinnedCount) cudaFreeHost(l->pinnedCount);
    ctx_release(l->ctx); free(l);
}
// a list that has an overlapped upload in flight: the main stream waits for it before the first consumer
static int list_acquire(chord_cmdlist* l)
{
    if (l && l->uploadPending) { CHORD_CUDA(cudaStreamWaitEvent(l->ctx->stream, l->uploaded, 0)); l->uploadPending = false; }
    return CHORD_OK;
}
// after a consumer: the next overlapped upload into this list must not overtake it
static int list_release(chord_cmdlist* l)
{
    if (l && l->asyncUser && !l->ctx->capturing) CHORD_CUDA(cudaEventRecord(l->lastUse, l->ctx->stream));
    return CHORD_OK;
}
#define CHORD_ACQUIRE(l) do { int rc__ = list_acquire(l); if (rc__) return rc__; } while (0)
#define CHORD_RELEASE(l) do { int rc__ = list_release(l); if (rc__) return rc__; } while (0)

int chord_cmdlist_upload(chord_cmdlist* l, const chord_draw_cmd* host, uint32_t count)
{
    CHORD_REQUIRE(l && (host || count == 0), "null");
    if (count > l->capacity) { set_error("cmdlist upload: %u > capacity %u", count, l->capacity); return CHORD_ERR_CAPACITY; }
    CHORD_ACQUIRE(l);
    CHORD_CUDA(cudaMemcpyAsync(l->count, &count, 4, cudaMemcpyHostToDevice, l->ctx->stream));
    if (count) CHORD_CUDA(cudaMemcpyAsync(l->cmds, host, sizeof(chord_draw_cmd) * (size_t)count, cudaMemcpyHostToDevice, l->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(l->ctx->stream));
    return CHORD_OK;
}
int chord_cmdlist_upload_async(chord_cmdlist* l, const chord_draw_cmd* host, uint32_t count)
{
    // stream-ordered, no host synchronisation: `host` (pinned for a true async copy) must stay valid until the stream passes it
    CHORD_REQUIRE(l && (host || count == 0), "null");
    if (count > l->capacity) { set_error("cmdlist upload: %u > capacity %u", count, l->capacity); return CHORD_ERR_CAPACITY; }
    // the count travels as a 4-byte kernel-free copy from a small pinned scratch owned by the list
    chord_ctx* ctx = l->ctx;
    uint32_t* slot = nullptr;
    if (ctx->capturing) {
        // captured upload: the count is snapshotted into pinned memory OWNED BY THE GRAPH (like chord_frame_begin's uniforms);
        // a slot of the list's ring would be overwritten by later uploads and every replay would copy whatever it holds then
        CHORD_CUDA(cudaHostAlloc((void**)&slot, 64, cudaHostAllocDefault));
        ctx->captureAllocs = (void**)realloc(ctx->captureAllocs, sizeof(void*) * (ctx->nCaptureAllocs + 1));
        ctx->captureAllocs[ctx->nCaptureAllocs++] = slot;
    } else {
        if (!l->pinnedCount) CHORD_CUDA(cudaMallocHost(&l->pinnedCount, 64 * sizeof(uint32_t)));
        slot = &l->pinnedCount[l->pinnedIdx++ & 63u];                       // ring: up to 64 uploads in flight
    }
    *slot = count;
    cudaStream_t st = ctx->stream;
    if (!ctx->capturing) {
        // the copy runs on the context's upload stream so that it overlaps whatever the main stream is executing (the
        // raster of the previous list); events order it after the last consumer of this list and before the next one
        if (!ctx->uploadStream) CHORD_CUDA(cudaStreamCreateWithFlags(&ctx->uploadStream, cudaStreamNonBlocking));
        if (!l->uploaded) { CHORD_CUDA(cudaEventCreateWithFlags(&l->uploaded, cudaEventDisableTiming)); CHORD_CUDA(cudaEventCreateWithFlags(&l->lastUse, cudaEventDisableTiming)); }
        if (!l->asyncUser) { l->asyncUser = true; CHORD_CUDA(cudaEventRecord(l->lastUse, ctx->stream)); }   // everything enqueued so far may touch the list
        st = ctx->uploadStream;
        CHORD_CUDA(cudaStreamWaitEvent(st, l->lastUse, 0));
    }
    CHORD_CUDA(cudaMemcpyAsync(l->count, slot, 4, cudaMemcpyHostToDevice, st));
    if (count) CHORD_CUDA(cudaMemcpyAsync(l->cmds, host, sizeof(chord_draw_cmd) * (size_t)count, cudaMemcpyHostToDevice, st));
    if (!ctx->capturing) { CHORD_CUDA(cudaEventRecord(l->uploaded, st)); l->uploadPending = true; }
    return CHORD_OK;
}

int chord_cmdlist_readback(chord_cmdlist* l, chord_draw_cmd* host, uint32_t hostCapacity, uint32_t* outCount)
{
    CHORD_REQUIRE(l && outCount, "null");
    CHORD_ACQUIRE(l);
    uint32_t n = 0;
    CHORD_CUDA(cudaMemcpyAsync(&n, l->count, 4, cudaMemcpyDeviceToHost, l->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(l->ctx->stream));
    *outCount = n;
    if (n > l->capacity) { set_error("cmdlist overflowed on device: %u > capacity %u", n, l->capacity); return CHORD_ERR_CAPACITY; }
    uint32_t m = n < hostCapacity ? n : hostCapacity;
    if (host && m) {
        CHORD_CUDA(cudaMemcpyAsync(host, l->cmds, sizeof(chord_draw_cmd) * (size_t)m, cudaMemcpyDeviceToHost, l->ctx->stream));
        CHORD_CUDA(cudaStreamSynchronize(l->ctx->stream));
    }
    return CHORD_OK;
}
uint32_t chord_cmdlist_capacity(const chord_cmdlist* l) { return l ? l->capacity : 0; }
void* chord_cmdlist_count_ptr(chord_cmdlist* l) { return l ? l->count : nullptr; }
void* chord_cmdlist_cmds_ptr(chord_cmdlist* l) { return l ? l->cmds : nullptr; }

// ---------------------------------------------------------------------------------------------- HZB objects
static uint32_t next_pot(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

int chord_hzb_create(chord_ctx* ctx, uint32_t depthW, uint32_t depthH, chord_hzb** out)
{
    CHORD_REQUIRE(ctx && out, "null");
    CHORD_REQUIRE(depthW >= CHORD_MIN_RENDER_DIM && depthH >= CHORD_MIN_RENDER_DIM && depthW <= CHORD_MAX_RENDER_DIM && depthH <= CHORD_MAX_RENDER_DIM, "render dimension outside [64, 4096] (renderer.h:52-53)");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    chord_hzb* h = (chord_hzb*)calloc(1, sizeof(chord_hzb));
    CHORD_REQUIRE(h != nullptr, "out of host memory");
    h->ctx = ctx; h->depthW = depthW; h->depthH = depthH;
    // hzb.cpp:49-56
    uint32_t tw = next_pot(depthW) / 2, th = next_pot(depthH) / 2;
    if (tw == depthW) tw /= 2;
    if (th == depthH) th /= 2;
    h->mip0W = tw; h->mip0H = th;
    uint32_t mx = tw > th ? tw : th, mips = 1;
    while (mx > 1) { mx >>= 1; mips++; }
    h->mipCount = mips;
    if (mips < 6 || mips > CHORD_HZB_MAX_MIP_COUNT) { set_error("hzb mip count %u outside [6,12] (hzb.cpp:119)", mips); free(h); return CHORD_ERR_INVALID; }
    ctx_retain(ctx);
    uint64_t o = 0;
    for (uint32_t m = 0; m < mips; m++) {
        h->mipOffset[m] = o;
        uint32_t w = tw >> m, hh = th >> m; if (!w) w = 1; if (!hh) hh = 1;
        o += (uint64_t)w * hh;
    }
    h->totalTexels = o;
    const int rc = [&]() -> int {
        CHORD_CUDA(cudaMalloc(&h->minChain, sizeof(__half) * o));
        CHORD_CUDA(cudaMalloc(&h->maxChain, sizeof(__half) * o));
        CHORD_CUDA(cudaMalloc(&h->validRange, 16));
        CHORD_CUDA(cudaMalloc(&h->counter, 16));
        CHORD_CUDA(cudaMemsetAsync(h->counter, 0, 16, ctx->stream));
        CHORD_CUDA(cudaMemsetAsync(h->minChain, 0, sizeof(__half) * o, ctx->stream));
        CHORD_CUDA(cudaMemsetAsync(h->maxChain, 0, sizeof(__half) * o, ctx->stream));
        return CHORD_OK;
    }();
    if (rc != CHORD_OK) { chord_hzb_destroy(h); return rc; }
    *out = h;
    return CHORD_OK;
}
void chord_hzb_destroy(chord_hzb* h)
{
    if (!h) return;
    cudaStreamSynchronize(h->ctx->stream);
    if (!h->externalChains) { cudaFree(h->minChain); cudaFree(h->maxChain); }
    cudaFree(h->validRange); cudaFree(h->counter);
    ctx_release(h->ctx); free(h);
}
int chord_hzb_info(const chord_hzb* h, uint32_t* w, uint32_t* hh, uint32_t* mips, uint32_t* vw, uint32_t* vh)
{
    CHORD_REQUIRE(h, "null");
    if (w) *w = h->mip0W; if (hh) *hh = h->mip0H; if (mips) *mips = h->mipCount;
    if (vw) *vw = h->depthW / 2; if (vh) *vh = h->depthH / 2;       // hzb.cpp:151 mip0ValidArea
    return CHORD_OK;
}
int chord_hzb_readback(chord_hzb* h, int which, uint32_t mip, uint16_t* host)
{
    CHORD_REQUIRE(h && host, "null");
    CHORD_REQUIRE(mip < h->mipCount, "mip");
    uint32_t w = h->mip0W >> mip, hh = h->mip0H >> mip; if (!w) w = 1; if (!hh) hh = 1;
    const __half* src = (which ? h->maxChain : h->minChain) + h->mipOffset[mip];
    CHORD_CUDA(cudaMemcpyAsync(host, src, sizeof(__half) * (size_t)w * hh, cudaMemcpyDeviceToHost, h->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return CHORD_OK;
}
int chord_hzb_readback_valid_range(chord_hzb* h, uint32_t out[2])
{
    CHORD_REQUIRE(h && out, "null");
    CHORD_CUDA(cudaMemcpyAsync(out, h->validRange, 8, cudaMemcpyDeviceToHost, h->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return CHORD_OK;
}
void* chord_hzb_chain_ptr(chord_hzb* h, int which) { return h ? (void*)(which ? h->maxChain : h->minChain) : nullptr; }
uint64_t chord_hzb_chain_bytes(const chord_hzb* h) { return h ? h->totalTexels * sizeof(__half) : 0; }
uint64_t chord_hzb_mip_offset(const chord_hzb* h, uint32_t mip) { return (h && mip < h->mipCount) ? h->mipOffset[mip] * sizeof(__half) : 0; }
int chord_hzb_set_valid(chord_hzb* h, int v) { CHORD_REQUIRE(h, "null"); h->valid = v; return CHORD_OK; }
int chord_hzb_is_valid(const chord_hzb* h) { return h ? h->valid : 0; }

// ---------------------------------------------------------------------------------------------- frames
static const uint32_t kLargeTriCapacity = 1u << 18;      // 256 k queued large triangles per render_mesh call
static const uint32_t kLargeItemCapacity = 1u << 22;     // 4 M (triangle, tile-row) work items
static const uint32_t kDeferredCapacity = 1u << 22;      // 4 M deferred triangles (64 B each); overflow -> k_raster_general

int chord_frame_create(chord_ctx* ctx, uint32_t width, uint32_t height, chord_frame** out)
{
    CHORD_REQUIRE(ctx && out, "null");
    CHORD_REQUIRE(width >= CHORD_MIN_RENDER_DIM && height >= CHORD_MIN_RENDER_DIM && width <= CHORD_MAX_RENDER_DIM && height <= CHORD_MAX_RENDER_DIM, "render dimension outside [64, 4096] (renderer.h:52-53)");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    chord_frame* f = (chord_frame*)calloc(1, sizeof(chord_frame));
    CHORD_REQUIRE(f != nullptr, "out of host memory");
    f->ctx = ctx; f->width = width; f->height = height; ctx_retain(ctx);
    f->tile = chord_rect{0, 0, width, height};
    const int rc = [&]() -> int {
        CHORD_CUDA(cudaMalloc(&f->vis64, sizeof(unsigned long long) * (size_t)width * height));
        CHORD_CUDA(cudaMalloc(&f->totalGroups, 16));
        CHORD_CUDA(cudaMemsetAsync(f->totalGroups, 0, 16, ctx->stream));      // word 0: total groups; word 2: block counter of k_object_cull_scan
        CHORD_CUDA(cudaMalloc(&f->bvhCtrl, 32));
        CHORD_CUDA(cudaMemsetAsync(f->bvhCtrl, 0, 32, ctx->stream));
        f->largeTriCapacity = kLargeTriCapacity; f->largeItemCapacity = kLargeItemCapacity;
        CHORD_CUDA(cudaMalloc(&f->largeTris, (size_t)48 * f->largeTriCapacity));
        CHORD_CUDA(cudaMalloc(&f->largeItems, (size_t)8 * f->largeItemCapacity));
        CHORD_CUDA(cudaMalloc(&f->largeCounters, 64));
        f->deferredCapacity = kDeferredCapacity;
        CHORD_CUDA(cudaMalloc(&f->deferred, (size_t)64 * f->deferredCapacity));
        CHORD_CUDA(cudaMalloc(&f->stats, sizeof(chord_raster_stats)));
        CHORD_CUDA(cudaMemsetAsync(f->stats, 0, sizeof(chord_raster_stats), ctx->stream));
        CHORD_CUDA(cudaMemsetAsync(f->largeCounters, 0, 64, ctx->stream));
        const uint32_t mw = (width + 7) / 8, mh = (height + 7) / 8;
        CHORD_CUDA(cudaMalloc(&f->tileMarker, (size_t)16 * mw * mh));
        CHORD_CUDA(cudaMalloc(&f->tileList, (size_t)8 * mw * mh));
        CHORD_CUDA(cudaMalloc(&f->tileCount, 16));
        return CHORD_OK;
    }();
    if (rc != CHORD_OK) { chord_frame_destroy(f); return rc; }
    *out = f;
    return CHORD_OK;
}

void chord_frame_destroy(chord_frame* f)
{
    if (!f) return;
    cudaStreamSynchronize(f->ctx->stream);
    cudaFree(f->vis64); cudaFree(f->dInputs); cudaFree(f->dDerived); cudaFree(f->dResolved);
    for (int k = 0; k < 4; k++) { if (f->hInputs[k]) cudaFreeHost(f->hInputs[k]); if (f->inputsCopied[k]) cudaEventDestroy(f->inputsCopied[k]); }
    cudaFree(f->objGroupCount); cudaFree(f->objGroupBase); cudaFree(f->totalGroups);
    cudaFree(f->bvhEntries); cudaFree(f->bvhBits); cudaFree(f->bvhCtrl);
    cudaFree(f->works); cudaFree(f->redoList); cudaFree(f->maskedList); cudaFree(f->deferred);
    if (f->copyStream) { cudaStreamSynchronize(f->copyStream); for (int k = 0; k < 2; k++) { cudaFree(f->splitVis[k]); cudaEventDestroy(f->splitDone[k]); cudaEventDestroy(f->copyDone[k]); } }
    if (f->ownCopyStream) { cudaStreamSynchronize(f->ownCopyStream); cudaStreamDestroy(f->ownCopyStream); }
    cudaFree(f->largeTris); cudaFree(f->largeItems); cudaFree(f->largeCounters); cudaFree(f->stats);
    cudaFree(f->tileMarker); cudaFree(f->tileList); cudaFree(f->tileCount); cudaFree(f->resolveRecs);
    for (int i = 0; i < CHORD_GBUF_COUNT; i++) { cudaFree(f->gbuffer[i]); cudaFree(f->gbufferPackedBuf[i]); }
    ctx_release(f->ctx); free(f);
}

int chord_frame_begin(chord_frame* f, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                      const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount,
                      uint32_t lod0MeshletCount, uint32_t meshletGroupCount, const chord_rect* tile)
{
    CHORD_REQUIRE(f && view && iview && objects, "null");
    CHORD_REQUIRE(objectCount > 0, "objectCount (shouldRenderGLTF, mesh_raster.cpp:45-48)");
    CHORD_REQUIRE(lod0MeshletCount > 0 && meshletGroupCount > 0, "lod0MeshletCount/meshletGroupCount (instance_culling.cpp:108)");
    CHORD_REQUIRE(f->ctx->scene.datas != nullptr, "no scene uploaded");
    // the raster target is sized by the INSTANCE view (main view: == view.renderDimension; cascade views: cascadeDim)
    CHORD_REQUIRE((uint32_t)iview->renderDimension.x == f->width && (uint32_t)iview->renderDimension.y == f->height, "instanceView.renderDimension != frame size");
    chord_ctx* ctx = f->ctx;
    {
        // PerframeCollected totals (component_gltf_mesh.cpp:58-122) size the compaction scratch, the BVH candidate bitmask and
        // the BVH queue, while the kernels walk sum(prim.meshletGroupCount) over the visible objects: an under-reported
        // total would make them run past those buffers, so it is checked here (a few thousand objects at most)
        uint64_t groups = 0;
        for (uint32_t i = 0; i < objectCount; i++) {
            CHORD_REQUIRE(objects[i].GLTFPrimitiveDetail < ctx->scene.nPrims && objects[i].GLTFMaterialData < ctx->scene.nMats, "object refers to a primitive / material outside the uploaded scene");
            groups += ctx->primGroupCountHost[objects[i].GLTFPrimitiveDetail];
        }
        if (groups > meshletGroupCount) { set_error("meshletGroupCount %u is smaller than the %llu groups of the %u objects", meshletGroupCount, (unsigned long long)groups, objectCount); return CHORD_ERR_INVALID; }
    }
    CHORD_CUDA(cudaSetDevice(ctx->device));
    if (tile) {
        CHORD_REQUIRE(tile->x0 < tile->x1 && tile->y0 < tile->y1 && tile->x1 <= f->width && tile->y1 <= f->height, "tile rect");
        f->tile = *tile;
    } else {
        f->tile = chord_rect{0, 0, f->width, f->height};
    }
    constexpr size_t kViewBytes = sizeof(chord_PerframeCameraView), kIViewBytes = (sizeof(chord_InstanceCullingViewInfo) + 15) & ~(size_t)15;
    static_assert(sizeof(chord_PerframeCameraView) % 16 == 0, "view block alignment");
    const size_t kLvBytes = (lighting_view_bytes() + 15) & ~(size_t)15;
    if (objectCount > f->objectCapacity) {
        CHORD_NO_RESIZE_IN_CAPTURE(ctx, "the frame's object buffers");
        CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(f->dInputs); cudaFree(f->dDerived); cudaFree(f->dResolved); cudaFree(f->objGroupCount); cudaFree(f->objGroupBase);
        for (int k = 0; k < 4; k++) { if (f->hInputs[k]) cudaFreeHost(f->hInputs[k]); f->hInputs[k] = nullptr; }
        f->dInputs = nullptr; f->dDerived = nullptr; f->dResolved = nullptr; f->objGroupCount = nullptr; f->objGroupBase = nullptr; f->objectCapacity = 0;
        uint32_t cap = objectCount + objectCount / 4 + 16;
        f->lightingViewOffset = kViewBytes + kIViewBytes; f->objectsOffset = kViewBytes + kIViewBytes + kLvBytes;
        f->inputsBytes = f->objectsOffset + sizeof(chord_GPUObjectGLTFPrimitive) * (size_t)cap;
        CHORD_CUDA(cudaMalloc(&f->dInputs, f->inputsBytes));
        for (int k = 0; k < 4; k++) {
            CHORD_CUDA(cudaHostAlloc((void**)&f->hInputs[k], f->inputsBytes, cudaHostAllocDefault));
            if (!f->inputsCopied[k]) CHORD_CUDA(cudaEventCreateWithFlags(&f->inputsCopied[k], cudaEventDisableTiming));
        }
        f->dView = (chord_PerframeCameraView*)f->dInputs;
        f->dIView = (chord_InstanceCullingViewInfo*)(f->dInputs + kViewBytes);
        f->dLightingView = f->dInputs + f->lightingViewOffset;
        f->dObjects = (chord_GPUObjectGLTFPrimitive*)(f->dInputs + f->objectsOffset);
        CHORD_CUDA(cudaMalloc(&f->dDerived, sizeof(ObjectDerived) * (size_t)cap));
        CHORD_CUDA(cudaMalloc(&f->dResolved, sizeof(ObjectResolved) * (size_t)cap));
        CHORD_CUDA(cudaMalloc(&f->objGroupCount, sizeof(uint32_t) * (size_t)(cap + 1)));
        CHORD_CUDA(cudaMalloc(&f->objGroupBase, sizeof(uint32_t) * (size_t)(cap + 1)));
        f->objectCapacity = cap;
    }
    f->objectCount = objectCount;
    f->lod0MeshletCount = lod0MeshletCount; f->meshletGroupCount = meshletGroupCount;
    f->hView = *view;
    f->hIViewOrtho = iview->translatedWorldToClip.m[15] == 1.0f;       // base.hlsli:241-244
    f->tanHalfFovy = tanf(0.5f * view->cameraFovy);
    const size_t inputBytes = f->objectsOffset + sizeof(chord_GPUObjectGLTFPrimitive) * (size_t)objectCount;
    if (ctx->capturing) {
        // captured frame: the uniforms are snapshotted into pinned memory owned by the graph; every replay re-uploads them
        // (chord_graph_update_frame_inputs rewrites the snapshot between replays)
        unsigned char* h = nullptr;
        CHORD_CUDA(cudaHostAlloc((void**)&h, inputBytes, cudaHostAllocDefault));
        ctx->captureAllocs = (void**)realloc(ctx->captureAllocs, sizeof(void*) * (ctx->nCaptureAllocs + 1));
        ctx->captureAllocs[ctx->nCaptureAllocs++] = h;
        ctx->captureFrames = (chord_graph_frame_inputs*)realloc(ctx->captureFrames, sizeof(chord_graph_frame_inputs) * (ctx->nCaptureFrames + 1));
        ctx->captureFrames[ctx->nCaptureFrames++] = chord_graph_frame_inputs{f, h, objectCount, f->tanHalfFovy};
        fill_frame_inputs(f, h, view, iview, objects, objectCount);
        CHORD_CUDA(cudaMemcpyAsync(f->dInputs, h, inputBytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // one H2D copy for view + iview + resolve uniforms + objects, staged through a 4-slot pinned ring (a slot is reused
        // only after the copy that read it has completed)
        const uint32_t slot = f->inputsSlot++ & 3u;
        CHORD_CUDA(cudaEventSynchronize(f->inputsCopied[slot]));
        unsigned char* h = f->hInputs[slot];
        fill_frame_inputs(f, h, view, iview, objects, objectCount);
        CHORD_CUDA(cudaMemcpyAsync(f->dInputs, h, inputBytes, cudaMemcpyHostToDevice, ctx->stream));
        CHORD_CUDA(cudaEventRecord(f->inputsCopied[slot], ctx->stream));
    }
    // addClearGbufferPass (render_textures.cpp:74-102): visibility = 0, depth = 0.0 -> packed texel 0.  Only texels inside
    // the region that can be non-zero are rewritten: a rank that renders the same screen tile every frame clears just
    // that tile (1/N of the surface); everything outside stays 0.
    {
        const chord_rect& t = f->tile;
        const bool same = f->clearedValid && f->clearedTile.x0 == t.x0 && f->clearedTile.y0 == t.y0 && f->clearedTile.x1 == t.x1 && f->clearedTile.y1 == t.y1;
        if (same && t.x1 - t.x0 < f->width) {
            CHORD_CUDA(cudaMemset2DAsync(f->vis64 + (size_t)t.y0 * f->width + t.x0, sizeof(unsigned long long) * (size_t)f->width, 0,
                                         sizeof(unsigned long long) * (size_t)(t.x1 - t.x0), t.y1 - t.y0, ctx->stream));
        } else if (same) {
            CHORD_CUDA(cudaMemsetAsync(f->vis64 + (size_t)t.y0 * f->width, 0, sizeof(unsigned long long) * (size_t)f->width * (t.y1 - t.y0), ctx->stream));
        } else {
            CHORD_CUDA(cudaMemsetAsync(f->vis64, 0, sizeof(unsigned long long) * (size_t)f->width * f->height, ctx->stream));
        }
        f->clearedTile = t; f->clearedValid = true;
    }
    f->begun = true;
    return launch_object_prepare(f);
}

int chord_instance_culling(chord_frame* f, chord_cmdlist* out)
{
    CHORD_REQUIRE(f && out && f->begun, "frame not begun");
    CHORD_ACQUIRE(out);
    const int rc = launch_instance_culling(f, out);
    CHORD_RELEASE(out);
    return rc;
}

int chord_instance_culling_bvh(chord_frame* f, chord_cmdlist* out)
{
    CHORD_REQUIRE(f && out && f->begun, "frame not begun");
    CHORD_ACQUIRE(out);
    const int rc = launch_instance_culling_bvh(f, out);
    CHORD_RELEASE(out);
    return rc;
}

int chord_readback_bvh_candidates(chord_frame* f, uint32_t* hostBits, uint32_t words, uint32_t* outCtrl5)
{
    CHORD_REQUIRE(f && hostBits && f->bvhBits, "no BVH traversal has run on this frame");
    CHORD_REQUIRE(words <= f->bvhWords, "words");
    CHORD_CUDA(cudaMemcpyAsync(hostBits, f->bvhBits, sizeof(uint32_t) * (size_t)words, cudaMemcpyDeviceToHost, f->ctx->stream));
    if (outCtrl5) CHORD_CUDA(cudaMemcpyAsync(outCtrl5, f->bvhCtrl, 20, cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return CHORD_OK;
}

int chord_hzb_culling(chord_frame* f, const chord_hzb* hzb, int firstStage, chord_cmdlist* in, chord_cmdlist* outVisible, chord_cmdlist* outRejected)
{
    CHORD_REQUIRE(f && hzb && in && outVisible && f->begun, "null / frame not begun");
    CHORD_REQUIRE(!firstStage || outRejected, "phase 0 needs a rejected list");
    CHORD_REQUIRE(hzb->depthW == f->width && hzb->depthH == f->height, "hzb built for another render dimension");
    CHORD_REQUIRE(outVisible->capacity >= in->capacity && (!outRejected || outRejected->capacity >= in->capacity), "output list capacity");
    CHORD_ACQUIRE(in); CHORD_ACQUIRE(outVisible); CHORD_ACQUIRE(outRejected);
    const int rc = launch_hzb_culling(f, hzb, firstStage, in, outVisible, firstStage ? outRejected : nullptr);
    CHORD_RELEASE(in); CHORD_RELEASE(outVisible); CHORD_RELEASE(outRejected);
    return rc;
}

int chord_render_mesh(chord_frame* f, chord_cmdlist* list)
{
    CHORD_REQUIRE(f && list && f->begun, "null / frame not begun");
    f->passFlags = 0u;
    CHORD_ACQUIRE(list);
    const int rc = launch_render_mesh(f, list);
    CHORD_RELEASE(list);
    return rc;
}

int chord_render_mesh_depth(chord_frame* f, chord_cmdlist* list, uint32_t passFlags)
{
    CHORD_REQUIRE(f && list && f->begun, "null / frame not begun");
    CHORD_REQUIRE((passFlags & ~(uint32_t)(CHORD_PASS_DEPTH | CHORD_PASS_DEPTH_CLAMP)) == 0u, "passFlags");
    // depth clamp drops the near / far clip planes: only defined here for orthographic instance views (w == 1)
    CHORD_REQUIRE(!(passFlags & CHORD_PASS_DEPTH_CLAMP) || f->hIViewOrtho, "CHORD_PASS_DEPTH_CLAMP needs an orthographic instance view");
    f->passFlags = passFlags;
    CHORD_ACQUIRE(list);
    const int rc = launch_render_mesh(f, list);
    f->passFlags = 0u;
    CHORD_RELEASE(list);
    return rc;
}

int chord_hzb_culling_generic(chord_frame* f, const chord_hzb* hzb, const chord_InstanceCullingViewInfo* hzbView, float extentScale,
                              int bObjectUseLastFrameProject, chord_cmdlist* in, chord_cmdlist* outVisible)
{
    CHORD_REQUIRE(f && hzb && hzbView && in && outVisible && f->begun, "null / frame not begun");
    CHORD_REQUIRE(hzb->depthW == (uint32_t)hzbView->renderDimension.x && hzb->depthH == (uint32_t)hzbView->renderDimension.y, "hzb was not built for hzbView's render dimension");
    CHORD_REQUIRE(outVisible->capacity >= in->capacity, "output list capacity");
    CHORD_ACQUIRE(in); CHORD_ACQUIRE(outVisible);
    const int rc = launch_hzb_culling_generic(f, hzb, hzbView, extentScale, bObjectUseLastFrameProject, in, outVisible);
    CHORD_RELEASE(in); CHORD_RELEASE(outVisible);
    return rc;
}

int chord_visibility_stage0(chord_frame* f, const chord_hzb* lastFrameHzb, chord_cmdlist* in, chord_cmdlist* scratchVisible,
                            chord_cmdlist* outRejected, int* needStage1)
{
    CHORD_REQUIRE(f && in && needStage1 && f->begun, "null / frame not begun");
    *needStage1 = 0;
    // mesh_raster.cpp:291-309: hzbCtx.isValid() && enableGLTFHZBCulling()
    if (lastFrameHzb && lastFrameHzb->valid && (f->ctx->switchFlags & (1u << CHORD_FLAG_HZB_CULL_BIT))) {
        CHORD_REQUIRE(scratchVisible && outRejected, "stage 0 with HZB needs scratch and rejected lists");
        *needStage1 = 1;
        int rc = chord_hzb_culling(f, lastFrameHzb, 1, in, scratchVisible, outRejected);
        if (rc) return rc;
        f->passFlags = 0u;
        rc = launch_render_mesh(f, scratchVisible);
        CHORD_RELEASE(scratchVisible);
        return rc;
    }
    return chord_render_mesh(f, in);
}

int chord_visibility_stage1(chord_frame* f, const chord_hzb* thisFrameHzb, chord_cmdlist* rejected, chord_cmdlist* scratchVisible)
{
    CHORD_REQUIRE(f && thisFrameHzb && rejected && scratchVisible && f->begun, "null / frame not begun");
    int rc = chord_hzb_culling(f, thisFrameHzb, 0, rejected, scratchVisible, nullptr);
    if (rc) return rc;
    f->passFlags = 0u;
    rc = launch_render_mesh(f, scratchVisible);
    CHORD_RELEASE(scratchVisible);
    return rc;
}

int chord_build_hzb(chord_frame* f, int bMin, int bMax, int bValidRange, chord_hzb* dst)
{
    CHORD_REQUIRE(f && dst && f->begun, "null / frame not begun");
    CHORD_REQUIRE(bMin || bMax, "at least one of min/max (hzb.cpp:44)");
    CHORD_REQUIRE(!bValidRange || (bMin && bMax), "valid range needs min and max (hzb.cpp:47)");
    CHORD_REQUIRE(dst->depthW == f->width && dst->depthH == f->height, "hzb created for another render dimension");
    int rc = launch_build_hzb(f, bMin, bMax, bValidRange, dst);
    if (rc == CHORD_OK) dst->valid = 1;
    return rc;
}

int chord_visibility_mark(chord_frame* f, chord_cmdlist* post)
{
    CHORD_REQUIRE(f && post && f->begun, "null / frame not begun");
    CHORD_ACQUIRE(post);
    const int rc = launch_visibility_mark(f, post);
    CHORD_RELEASE(post);
    return rc;
}

int chord_atmosphere_luts_upload(chord_ctx* ctx, const chord_atmosphere_luts* luts)
{
    CHORD_REQUIRE(ctx && luts && luts->transmittance && luts->irradiance, "null");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->lutTransmittance) CHORD_CUDA(cudaMalloc(&ctx->lutTransmittance, 256 * 64 * 4 * sizeof(__half)));
    if (!ctx->lutIrradiance) CHORD_CUDA(cudaMalloc(&ctx->lutIrradiance, 64 * 16 * 4 * sizeof(__half)));
    CHORD_CUDA(cudaMemcpyAsync(ctx->lutTransmittance, luts->transmittance, 256 * 64 * 4 * sizeof(__half), cudaMemcpyHostToDevice, ctx->stream));
    CHORD_CUDA(cudaMemcpyAsync(ctx->lutIrradiance, luts->irradiance, 64 * 16 * 4 * sizeof(__half), cudaMemcpyHostToDevice, ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    return CHORD_OK;
}

int chord_atmosphere_scattering_upload(chord_ctx* ctx, const uint16_t* scattering, const uint16_t* singleMie)
{
    CHORD_REQUIRE(ctx, "null");
    CHORD_REQUIRE(!ctx->capturing, "LUT uploads are invalid while capturing");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    constexpr size_t kBytes = (size_t)256 * 128 * 32 * 4 * sizeof(__half);
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!scattering) {                                   // removes the sky branch from chord_lighting
        cudaFree(ctx->lutScattering); cudaFree(ctx->lutSingleMie); ctx->lutScattering = nullptr; ctx->lutSingleMie = nullptr;
        return CHORD_OK;
    }
    if (!ctx->lutScattering) CHORD_CUDA(cudaMalloc(&ctx->lutScattering, kBytes));
    CHORD_CUDA(cudaMemcpyAsync(ctx->lutScattering, scattering, kBytes, cudaMemcpyHostToDevice, ctx->stream));
    if (singleMie) {
        if (!ctx->lutSingleMie) CHORD_CUDA(cudaMalloc(&ctx->lutSingleMie, kBytes));
        CHORD_CUDA(cudaMemcpyAsync(ctx->lutSingleMie, singleMie, kBytes, cudaMemcpyHostToDevice, ctx->stream));
    } else { cudaFree(ctx->lutSingleMie); ctx->lutSingleMie = nullptr; }
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    return CHORD_OK;
}

// ---------------------------------------------------------------------------------------------- cascade views (N3)
int chord_cascade_views_create(chord_ctx* ctx, chord_cascade_views** out)
{
    CHORD_REQUIRE(ctx && out, "null");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    chord_cascade_views* v = (chord_cascade_views*)calloc(1, sizeof(chord_cascade_views));
    CHORD_REQUIRE(v != nullptr, "out of host memory");
    v->ctx = ctx; ctx_retain(ctx);
    const int rc = [&]() -> int {
        CHORD_CUDA(cudaMalloc(&v->dViews, sizeof(chord_InstanceCullingViewInfo) * CHORD_MAX_CASCADE_COUNT));
        CHORD_CUDA(cudaMemsetAsync(v->dViews, 0, sizeof(chord_InstanceCullingViewInfo) * CHORD_MAX_CASCADE_COUNT, ctx->stream));
        return CHORD_OK;
    }();
    if (rc != CHORD_OK) { chord_cascade_views_destroy(v); return rc; }
    *out = v;
    return CHORD_OK;
}
void chord_cascade_views_destroy(chord_cascade_views* v)
{
    if (!v) return;
    cudaStreamSynchronize(v->ctx->stream);
    cudaFree(v->dViews); ctx_release(v->ctx); free(v);
}
int chord_cascade_setup(chord_frame* mainFrame, const chord_cascade_setup_desc* d, const chord_hzb* validRangeSource, chord_cascade_views* views)
{
    CHORD_REQUIRE(mainFrame && d && views && mainFrame->begun, "null / main frame not begun");
    CHORD_REQUIRE(d->cascadeCount >= 1 && d->cascadeCount <= CHORD_MAX_CASCADE_COUNT && d->realtimeCount <= d->cascadeCount, "cascade counts (mesh_raster.cpp:363)");
    CHORD_REQUIRE(d->cascadeDim >= CHORD_MIN_RENDER_DIM && d->cascadeDim <= CHORD_MAX_RENDER_DIM, "cascadeDim");
    views->cascadeDim = d->cascadeDim;
    return launch_cascade_setup(mainFrame, d, validRangeSource, views->dViews);
}
int chord_cascade_views_readback(chord_cascade_views* v, chord_InstanceCullingViewInfo* host, uint32_t count)
{
    CHORD_REQUIRE(v && host && count <= CHORD_MAX_CASCADE_COUNT, "args");
    CHORD_CUDA(cudaMemcpyAsync(host, v->dViews, sizeof(chord_InstanceCullingViewInfo) * count, cudaMemcpyDeviceToHost, v->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return CHORD_OK;
}
int chord_cascade_views_upload(chord_cascade_views* v, const chord_InstanceCullingViewInfo* host, uint32_t count)
{
    CHORD_REQUIRE(v && host && count >= 1 && count <= CHORD_MAX_CASCADE_COUNT, "args");
    CHORD_CUDA(cudaMemcpyAsync(v->dViews, host, sizeof(chord_InstanceCullingViewInfo) * count, cudaMemcpyHostToDevice, v->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(v->ctx->stream));
    v->cascadeDim = (uint32_t)host[0].renderDimension.x;
    return CHORD_OK;
}
int chord_frame_begin_cascade(chord_frame* f, const chord_PerframeCameraView* view, chord_cascade_views* views, uint32_t cascadeId,
                              const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount, uint32_t lod0MeshletCount, uint32_t meshletGroupCount)
{
    CHORD_REQUIRE(f && view && views && cascadeId < CHORD_MAX_CASCADE_COUNT, "args");
    CHORD_REQUIRE(views->cascadeDim == f->width && views->cascadeDim == f->height, "depth frame size != cascadeDim");
    // host-side placeholder with what chord_frame_begin checks on the host (size; cascade views are orthographic); the
    // device record is then overwritten by the cascade buffer's entry, stream-ordered behind the setup kernel
    chord_InstanceCullingViewInfo iv; memset(&iv, 0, sizeof(iv));
    iv.renderDimension.x = (float)f->width; iv.renderDimension.y = (float)f->height;
    iv.translatedWorldToClip.m[15] = 1.0f;
    const int rc = chord_frame_begin(f, view, &iv, objects, objectCount, lod0MeshletCount, meshletGroupCount, nullptr);
    if (rc) return rc;
    CHORD_CUDA(cudaMemcpyAsync(f->dIView, views->dViews + cascadeId, sizeof(chord_InstanceCullingViewInfo), cudaMemcpyDeviceToDevice, f->ctx->stream));
    return launch_object_prepare(f);        // the derived per-object matrices depend on the instance view
}
int chord_hzb_culling_generic_cascade(chord_frame* f, const chord_hzb* hzb, chord_cascade_views* views, uint32_t hzbViewId, float extentScale,
                                      int bObjectUseLastFrameProject, chord_cmdlist* in, chord_cmdlist* outVisible)
{
    CHORD_REQUIRE(f && hzb && views && hzbViewId < CHORD_MAX_CASCADE_COUNT && in && outVisible && f->begun, "null / frame not begun");
    CHORD_REQUIRE(hzb->depthW == views->cascadeDim && hzb->depthH == views->cascadeDim, "hzb was not built for the cascade dimension");
    CHORD_REQUIRE(outVisible->capacity >= in->capacity, "output list capacity");
    CHORD_ACQUIRE(in); CHORD_ACQUIRE(outVisible);
    const int rc = launch_hzb_culling_generic(f, hzb, nullptr, extentScale, bObjectUseLastFrameProject, in, outVisible, views->dViews + hzbViewId);
    CHORD_RELEASE(in); CHORD_RELEASE(outVisible);
    return rc;
}

// ---------------------------------------------------------------------------------------------- graphs
int chord_graph_begin_capture(chord_ctx* ctx)
{
    CHORD_REQUIRE(ctx && !ctx->capturing, "null / already capturing");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    // relaxed: the captured calls may touch pinned / device allocations (uniform snapshots) without invalidating the capture
    CHORD_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = true; ctx->captureAllocs = nullptr; ctx->nCaptureAllocs = 0; ctx->captureFrames = nullptr; ctx->nCaptureFrames = 0;
    return CHORD_OK;
}

int chord_graph_end_capture(chord_ctx* ctx, chord_graph** out)
{
    CHORD_REQUIRE(ctx && out && ctx->capturing, "null / not capturing");
    *out = nullptr;
    ctx->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (e != cudaSuccess || !graph) {
        for (uint32_t i = 0; i < ctx->nCaptureAllocs; i++) cudaFreeHost(ctx->captureAllocs[i]);
        free(ctx->captureAllocs); ctx->captureAllocs = nullptr; ctx->nCaptureAllocs = 0;
        free(ctx->captureFrames); ctx->captureFrames = nullptr; ctx->nCaptureFrames = 0;
        set_error("cudaStreamEndCapture failed: %s (a captured call synchronised or failed)", cudaGetErrorString(e));
        cudaGetLastError();
        return CHORD_ERR_CUDA;
    }
    chord_graph* g = (chord_graph*)calloc(1, sizeof(chord_graph));
    g->ctx = ctx; g->graph = graph; ctx_retain(ctx);
    g->pinned = ctx->captureAllocs; g->nPinned = ctx->nCaptureAllocs;
    ctx->captureAllocs = nullptr; ctx->nCaptureAllocs = 0;
    g->frames = ctx->captureFrames; g->nFrames = ctx->nCaptureFrames;
    ctx->captureFrames = nullptr; ctx->nCaptureFrames = 0;
    const int rc = [&]() -> int {
        size_t n = 0;
        CHORD_CUDA(cudaGraphGetNodes(graph, nullptr, &n));
        std::vector<cudaGraphNode_t> nodes(n);
        if (n) CHORD_CUDA(cudaGraphGetNodes(graph, nodes.data(), &n));
        for (size_t i = 0; i < n; i++) {
            cudaGraphNodeType t;
            if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) g->kernelNodes++;
        }
        CHORD_CUDA(cudaGraphInstantiate(&g->exec, graph, 0));
        return CHORD_OK;
    }();
    if (rc != CHORD_OK) { chord_graph_destroy(g); return rc; }     // frees the graph and the pinned uniform snapshots
    *out = g;
    return CHORD_OK;
}

int chord_graph_launch(chord_graph* g)
{
    CHORD_REQUIRE(g && g->exec && !g->ctx->capturing, "null / capturing");
    CHORD_CUDA(cudaGraphLaunch(g->exec, g->ctx->stream));
    g->ctx->launches += g->kernelNodes;
    return CHORD_OK;
}

int chord_graph_update_frame_inputs(chord_graph* g, chord_frame* f, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                                    const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount)
{
    // Rewrites the pinned snapshot(s) the recorded chord_frame_begin call(s) on `f` upload, so that the next chord_graph_launch
    // renders THIS camera / THESE object transforms.  Everything that was baked into kernel parameters at capture time must stay
    // the same: object count, frame size / tile, field of view (tan(fovy / 2) is a kernel parameter of the cull kernels).
    CHORD_REQUIRE(g && f && view && iview && objects, "null");
    uint32_t hits = 0;
    for (uint32_t i = 0; i < g->nFrames; i++) {
        chord_graph_frame_inputs& r = g->frames[i];
        if (r.frame != f) continue;
        CHORD_REQUIRE(r.objectCount == objectCount, "object count differs from the recorded frame");
        CHORD_REQUIRE(r.tanHalfFovy == tanf(0.5f * view->cameraFovy), "cameraFovy differs from the recorded frame (baked into the recorded kernels)");
        CHORD_REQUIRE((uint32_t)iview->renderDimension.x == f->width && (uint32_t)iview->renderDimension.y == f->height, "instanceView.renderDimension != frame size");
        // the previous launch of this graph must have consumed the snapshot: replays of one graph are stream-ordered, and the
        // upload is the first node, so waiting for the context stream here would be too strong; callers alternate between
        // graphs / keep one frame in flight (documented in chord_b200.h)
        fill_frame_inputs(f, r.pinned, view, iview, objects, objectCount);
        hits++;
    }
    CHORD_REQUIRE(hits > 0, "the graph recorded no chord_frame_begin on this frame");
    f->hView = *view;
    return CHORD_OK;
}

uint32_t chord_graph_kernel_nodes(const chord_graph* g) { return g ? g->kernelNodes : 0u; }

void chord_graph_destroy(chord_graph* g)
{
    if (!g) return;
    cudaStreamSynchronize(g->ctx->stream);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    for (uint32_t i = 0; i < g->nPinned; i++) cudaFreeHost(g->pinned[i]);
    free(g->pinned); free(g->frames);
    ctx_release(g->ctx); free(g);
}

int chord_textures_upload(chord_ctx* ctx, const chord_texture_desc* textures, uint32_t textureCount,
                          const chord_sampler_desc* samplers, uint32_t samplerCount)
{
    CHORD_REQUIRE(ctx, "null");
    CHORD_REQUIRE(textureCount == 0 || textures, "textures");
    CHORD_REQUIRE(samplerCount == 0 || samplers, "samplers");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->dTextures); cudaFree(ctx->dSamplers); cudaFree(ctx->dTexels); cudaFree(ctx->dTexelsRgba);
    ctx->dTextures = nullptr; ctx->dSamplers = nullptr; ctx->dTexels = nullptr; ctx->dTexelsRgba = nullptr; ctx->nTextures = ctx->nSamplers = 0;
    if (textureCount == 0) return CHORD_OK;
    // two copies per texture, each one mip chain, 16-byte aligned: the alpha channel alone (all the raster path reads: one byte
    // per texel) and the RGBA8 texels (resolve: base colour / normal map / occlusion-roughness-metallic)
    std::vector<TextureDev> table(textureCount);
    std::vector<size_t> offset(textureCount, 0);
    size_t total = 0;
    for (uint32_t i = 0; i < textureCount; i++) {
        const chord_texture_desc& t = textures[i];
        table[i] = TextureDev{nullptr, nullptr, t.width, t.height, t.mipCount, t.flags};
        if (!t.rgba8) continue;
        CHORD_REQUIRE(t.width >= 1 && t.height >= 1 && t.width <= 16384 && t.height <= 16384 && t.mipCount >= 1 && t.mipCount <= 15, "texture dimensions / mip count");
        CHORD_REQUIRE((t.flags & ~(uint32_t)CHORD_TEXTURE_SRGB) == 0u, "texture flags");
        size_t texels = 0; uint32_t w = t.width, h = t.height;
        for (uint32_t l = 0; l < t.mipCount; l++) { texels += (size_t)w * h; w = w > 1 ? w >> 1 : 1; h = h > 1 ? h >> 1 : 1; }
        offset[i] = total; total += (texels + 15) & ~(size_t)15;
    }
    std::vector<uint8_t> alpha(total + 16, 0), rgba(4 * total + 64, 0);
    for (uint32_t i = 0; i < textureCount; i++) {
        const chord_texture_desc& t = textures[i];
        if (!t.rgba8) continue;
        size_t texels = 0; uint32_t w = t.width, h = t.height;
        for (uint32_t l = 0; l < t.mipCount; l++) { texels += (size_t)w * h; w = w > 1 ? w >> 1 : 1; h = h > 1 ? h >> 1 : 1; }
        for (size_t k = 0; k < texels; k++) alpha[offset[i] + k] = t.rgba8[4 * k + 3];
        memcpy(rgba.data() + 4 * offset[i], t.rgba8, 4 * texels);
    }
    CHORD_CUDA(cudaMalloc(&ctx->dTexels, alpha.size()));
    CHORD_CUDA(cudaMemcpy(ctx->dTexels, alpha.data(), alpha.size(), cudaMemcpyHostToDevice));
    CHORD_CUDA(cudaMalloc(&ctx->dTexelsRgba, rgba.size()));
    CHORD_CUDA(cudaMemcpy(ctx->dTexelsRgba, rgba.data(), rgba.size(), cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < textureCount; i++) if (textures[i].rgba8) { table[i].alpha = ctx->dTexels + offset[i]; table[i].rgba = ctx->dTexelsRgba + 4 * offset[i]; }
    CHORD_CUDA(cudaMalloc(&ctx->dTextures, sizeof(TextureDev) * textureCount));
    CHORD_CUDA(cudaMemcpy(ctx->dTextures, table.data(), sizeof(TextureDev) * textureCount, cudaMemcpyHostToDevice));
    if (samplerCount) {
        std::vector<SamplerDev> st(samplerCount);
        for (uint32_t i = 0; i < samplerCount; i++) {
            CHORD_REQUIRE(samplers[i].addressU <= CHORD_ADDRESS_MIRRORED_REPEAT && samplers[i].addressV <= CHORD_ADDRESS_MIRRORED_REPEAT, "sampler address mode");
            st[i] = SamplerDev{samplers[i].addressU, samplers[i].addressV};
        }
        CHORD_CUDA(cudaMalloc(&ctx->dSamplers, sizeof(SamplerDev) * samplerCount));
        CHORD_CUDA(cudaMemcpy(ctx->dSamplers, st.data(), sizeof(SamplerDev) * samplerCount, cudaMemcpyHostToDevice));
    }
    ctx->nTextures = textureCount; ctx->nSamplers = samplerCount;
    return CHORD_OK;
}

int chord_lighting(chord_frame* f, chord_cmdlist* post)
{
    CHORD_REQUIRE(f && post && f->begun, "null / frame not begun");
    CHORD_REQUIRE(f->ctx->lutTransmittance && f->ctx->lutIrradiance, "atmosphere LUTs not uploaded");
    CHORD_ACQUIRE(post);
    const int rc = launch_lighting(f, post);
    CHORD_RELEASE(post);
    return rc;
}

// ---------------------------------------------------------------------------------------------- read-backs
int chord_readback_vis64(chord_frame* f, uint64_t* host)
{
    CHORD_REQUIRE(f && host, "null");
    CHORD_CUDA(cudaMemcpyAsync(host, f->vis64, sizeof(uint64_t) * (size_t)f->width * f->height, cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return CHORD_OK;
}

static int readback_split(chord_frame* f, uint32_t* hostVis, float* hostDepth)
{
    const size_t n = (size_t)f->width * f->height;
    uint32_t* dVis = nullptr; float* dDepth = nullptr;
    CHORD_CUDA(cudaMallocAsync(&dVis, n * 4, f->ctx->stream));
    CHORD_CUDA(cudaMallocAsync(&dDepth, n * 4, f->ctx->stream));
    int rc = launch_split_vis(f, dVis, dDepth);
    if (rc == CHORD_OK && hostVis) CHORD_CUDA(cudaMemcpyAsync(hostVis, dVis, n * 4, cudaMemcpyDeviceToHost, f->ctx->stream));
    if (rc == CHORD_OK && hostDepth) CHORD_CUDA(cudaMemcpyAsync(hostDepth, dDepth, n * 4, cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaFreeAsync(dVis, f->ctx->stream));
    CHORD_CUDA(cudaFreeAsync(dDepth, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return rc;
}
static int ensure_async_readback(chord_frame* f)
{
    if (f->copyStream) return CHORD_OK;
    const size_t n = (size_t)f->width * f->height;
    // ONE transfer stream per context, shared by the work-list uploads and the image read-backs: when a PCIe upload and a
    // read-back run at the same time the raster they overlap slows down by 17 % (1.89 -> 2.21 ms / step on S5), one after
    // the other costs 5 % (1.98 ms); both still fit inside the step they hide behind.
    if (!f->ctx->uploadStream) CHORD_CUDA(cudaStreamCreateWithFlags(&f->ctx->uploadStream, cudaStreamNonBlocking));
    f->copyStream = f->ctx->uploadStream; f->copyStreamShared = true;
    for (int k = 0; k < 2; k++) {
        CHORD_CUDA(cudaMalloc(&f->splitVis[k], n * 4));
        CHORD_CUDA(cudaEventCreateWithFlags(&f->splitDone[k], cudaEventDisableTiming));
        CHORD_CUDA(cudaEventCreateWithFlags(&f->copyDone[k], cudaEventDisableTiming));
    }
    return CHORD_OK;
}

// Pipelined read-back: the R32_UINT image of the CURRENT surface is extracted on the context stream into one of two
// staging slots and copied to `hostPinned` on a separate copy stream, so that the next frame's kernels overlap the
// PCIe transfer.  chord_frame_wait_readbacks() (or destroying the frame) completes all outstanding copies.
int chord_readback_visibility32_async(chord_frame* f, uint32_t* hostPinned)
{
    CHORD_REQUIRE(f && hostPinned, "null");
    int rc = ensure_async_readback(f);
    if (rc) return rc;
    const int k = f->splitIdx; f->splitIdx ^= 1;
    CHORD_CUDA(cudaStreamWaitEvent(f->ctx->stream, f->copyDone[k], 0));          // slot k no longer being copied out
    rc = launch_split_vis(f, f->splitVis[k], nullptr);
    if (rc) return rc;
    CHORD_CUDA(cudaEventRecord(f->splitDone[k], f->ctx->stream));
    CHORD_CUDA(cudaStreamWaitEvent(f->copyStream, f->splitDone[k], 0));
    // only the frame's screen tile travels (the other tiles belong to other GPUs): a pitched copy into the same
    // position of the host image
    const size_t pitch = (size_t)f->width * 4, off = (size_t)f->tile.y0 * f->width + f->tile.x0;
    if (f->tile.x0 == 0 && f->tile.x1 == f->width) {
        // full-width tile (always the case on one GPU): the rows are contiguous, one linear copy instead of a pitched one
        CHORD_CUDA(cudaMemcpyAsync(hostPinned + off, f->splitVis[k] + off, pitch * (size_t)(f->tile.y1 - f->tile.y0), cudaMemcpyDeviceToHost, f->copyStream));
    } else {
        CHORD_CUDA(cudaMemcpy2DAsync(hostPinned + off, pitch, f->splitVis[k] + off, pitch, (size_t)(f->tile.x1 - f->tile.x0) * 4,
                                     f->tile.y1 - f->tile.y0, cudaMemcpyDeviceToHost, f->copyStream));
    }
    CHORD_CUDA(cudaEventRecord(f->copyDone[k], f->copyStream));
    return CHORD_OK;
}

// Which stream carries the pipelined read-backs: 0 (default) = the context's transfer stream, shared with the work-list
// uploads, so that only one PCIe transfer is in flight at a time; 1 = a stream of the frame's own, so that the upload of the
// next list and the read-back of the last image overlap each other as well (full duplex).  Which one wins depends on how
// the step compares with the transfers: concurrent transfers slow the overlapped raster down (measured once: 17 %), one
// after the other they take longer - bench.py probes both before its timed e2e leg.  Completes pending copies first.
int chord_frame_set_readback_stream(chord_frame* f, int ownStream)
{
    CHORD_REQUIRE(f, "null");
    const int rc = ensure_async_readback(f);
    if (rc) return rc;
    CHORD_CUDA(cudaStreamSynchronize(f->copyStream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    if (ownStream) {
        if (!f->ownCopyStream) CHORD_CUDA(cudaStreamCreateWithFlags(&f->ownCopyStream, cudaStreamNonBlocking));
        f->copyStream = f->ownCopyStream; f->copyStreamShared = false;
    } else {
        f->copyStream = f->ctx->uploadStream; f->copyStreamShared = true;
    }
    return CHORD_OK;
}

int chord_frame_wait_readbacks(chord_frame* f)
{
    CHORD_REQUIRE(f, "null");
    if (f->copyStream) CHORD_CUDA(cudaStreamSynchronize(f->copyStream));
    return CHORD_OK;
}

int chord_readback_visibility32(chord_frame* f, uint32_t* host) { CHORD_REQUIRE(f && host, "null"); return readback_split(f, host, nullptr); }
int chord_readback_depth32(chord_frame* f, float* host) { CHORD_REQUIRE(f && host, "null"); return readback_split(f, nullptr, host); }

int chord_selftest_vertex_records(chord_ctx* ctx, uint32_t n, uint32_t seed, uint32_t width, uint32_t height, uint32_t depthClamp, uint64_t* hostOut3)
{
    CHORD_REQUIRE(ctx && hostOut3 && n > 0 && width > 0 && height > 0, "bad arguments");
    unsigned long long* d = nullptr;
    CHORD_CUDA(cudaMallocAsync(&d, 24, ctx->stream));
    const unsigned long long init[3] = { 0ull, ~0ull, 0ull };
    CHORD_CUDA(cudaMemcpyAsync(d, init, 24, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_selftest_vertex_records(ctx, n, seed, width, height, depthClamp, d);
    if (rc == CHORD_OK) CHORD_CUDA(cudaMemcpyAsync(hostOut3, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CHORD_CUDA(cudaFreeAsync(d, ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    return rc;
}

int chord_readback_decoded(chord_frame* f, chord_cmdlist* post, uint32_t* hostTriples)
{
    CHORD_REQUIRE(f && post && hostTriples, "null");
    CHORD_ACQUIRE(post);
    const size_t n = (size_t)f->width * f->height;
    uint32_t* d = nullptr;
    CHORD_CUDA(cudaMallocAsync(&d, n * 12, f->ctx->stream));
    int rc = launch_decode(f, post, d);
    if (rc == CHORD_OK) CHORD_CUDA(cudaMemcpyAsync(hostTriples, d, n * 12, cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaFreeAsync(d, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return rc;
}

int chord_readback_tile_marker(chord_frame* f, uint32_t* host)
{
    CHORD_REQUIRE(f && host, "null");
    const uint32_t mw = (f->width + 7) / 8, mh = (f->height + 7) / 8;
    CHORD_CUDA(cudaMemcpyAsync(host, f->tileMarker, (size_t)16 * mw * mh, cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return CHORD_OK;
}

int chord_readback_gbuffer_f32(chord_frame* f, int which, float* host)
{
    CHORD_REQUIRE(f && host && which >= 0 && which < CHORD_GBUF_COUNT, "args");
    CHORD_REQUIRE(f->gbuffer[which] != nullptr, "lighting has not run on this frame");
    CHORD_CUDA(cudaMemcpyAsync(host, f->gbuffer[which], sizeof(float) * 4 * (size_t)f->width * f->height, cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return CHORD_OK;
}

int chord_frame_set_queue_limits(chord_frame* f, uint32_t deferredTriangles, uint32_t largeTriangles, uint32_t largeItems)
{
    // test / tuning knob: shrink the logical capacities of the raster queues (never beyond the allocation) to exercise the
    // overflow fallbacks (deferred queue -> k_raster_general redo, large queues -> in-thread rasterisation)
    CHORD_REQUIRE(f, "null");
    if (deferredTriangles) f->deferredCapacity = deferredTriangles < kDeferredCapacity ? deferredTriangles : kDeferredCapacity;
    if (largeTriangles) f->largeTriCapacity = largeTriangles < kLargeTriCapacity ? largeTriangles : kLargeTriCapacity;
    if (largeItems) f->largeItemCapacity = largeItems < kLargeItemCapacity ? largeItems : kLargeItemCapacity;
    return CHORD_OK;
}

int chord_frame_set_gbuffer_packed(chord_frame* f, int packed)
{
    CHORD_REQUIRE(f, "null");
    f->gbufferPacked = packed ? 1 : 0;
    return CHORD_OK;
}

int chord_readback_gbuffer_packed(chord_frame* f, int which, void* host)
{
    static const size_t kPackedBytes[CHORD_GBUF_COUNT] = { 8, 4, 4, 4, 4, 4 };
    CHORD_REQUIRE(f && host && which >= 0 && which < CHORD_GBUF_COUNT, "args");
    CHORD_REQUIRE(f->gbufferPackedBuf[which] != nullptr, "lighting has not run in packed mode on this frame");
    CHORD_CUDA(cudaMemcpyAsync(host, f->gbufferPackedBuf[which], kPackedBytes[which] * (size_t)f->width * f->height, cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return CHORD_OK;
}

void* chord_frame_vis64_ptr(chord_frame* f) { return f ? f->vis64 : nullptr; }
int chord_frame_dims(const chord_frame* f, uint32_t* w, uint32_t* h) { CHORD_REQUIRE(f, "null"); if (w) *w = f->width; if (h) *h = f->height; return CHORD_OK; }

int chord_frame_raster_stats(chord_frame* f, chord_raster_stats* out)
{
    CHORD_REQUIRE(f && out, "null");
    CHORD_CUDA(cudaMemcpyAsync(out, f->stats, sizeof(*out), cudaMemcpyDeviceToHost, f->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return CHORD_OK;
}
int chord_frame_reset_stats(chord_frame* f)
{
    CHORD_REQUIRE(f, "null");
    CHORD_CUDA(cudaMemsetAsync(f->stats, 0, sizeof(chord_raster_stats), f->ctx->stream));
    return CHORD_OK;
}

}  // extern "C"
