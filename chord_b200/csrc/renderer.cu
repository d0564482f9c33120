// renderer.cu — the frame sequence of DeferredRenderer::render (source/renderer/renderer.cpp:311-363, history hand-off :489)
// as ONE native call per frame, plus the multi-GPU exchange of the history HZB inside the C ABI.
//
//   chord_renderer_render      uniforms + clear -> instanceCulling -> gltfVisibilityRenderingStage0 against LAST frame's HZB ->
//                              buildHZB(min) -> gltfVisibilityRenderingStage1 -> buildHZB(min, max, range) kept as history ->
//                              visibilityMark -> lighting.  With CHORD_RENDERER_GRAPH the sequence is recorded once per history
//                              slot (two CUDA graphs: the history ping-pongs between two HZBs) and replayed with the frame's
//                              uniforms rewritten in the graphs' pinned snapshot: one launch per frame instead of ~30.
//   exchange, fused (default)  every rank owns one screen tile.  Behind the final buildHZB of a frame a put kernel pushes the
//                              tile's rows of mips 0..5 straight into EVERY rank's history chain (peer pointers over NVLink: CUDA
//                              IPC between processes, peer access inside one), its last block publishes an arrival flag; the next
//                              frame has a one-block kernel (on a stream of its own, beside the instance culling) that waits for
//                              all flags and rebuilds mips 6.. locally.
//                              No pack / all-gather / unpack, no host involvement, graph-capturable.
//   exchange, NCCL             pack -> ncclAllGather -> unpack -> top rebuild on the context stream (NCCL resolved at run time
//                              with dlopen; the literal "NCCL all-gather of the HZB mip chain" of the north star, and the
//                              fallback where peer access is not available).
// Host code only; the kernels live in cull.cu / raster.cu / hzb.cu / resolve.cu.
#include "common.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>

struct chord_nccl_comm {
    chord_ctx* ctx;
    void* comm;                 // ncclComm_t
    uint32_t rank, world;
};

namespace chordb {

// ---- NCCL, resolved at run time (no link-time dependency: the library loads where NCCL is absent) ---------------------
struct NcclUniqueId { char internal[128]; };                       // nccl.h: NCCL_UNIQUE_ID_BYTES 128
typedef int (*pfn_ncclGetUniqueId)(NcclUniqueId*);
typedef int (*pfn_ncclCommInitRank)(void**, int, NcclUniqueId, int);
typedef int (*pfn_ncclCommDestroy)(void*);
typedef int (*pfn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*pfn_ncclGetErrorString)(int);
struct NcclApi {
    void* lib; pfn_ncclGetUniqueId getUniqueId; pfn_ncclCommInitRank commInitRank; pfn_ncclCommDestroy commDestroy;
    pfn_ncclAllGather allGather; pfn_ncclGetErrorString errorString;
};
static NcclApi* nccl_api()
{
    static NcclApi api; static int state = 0;                       // 0 = not tried, 1 = ok, 2 = unavailable
    if (state == 0) {
        // a process that imported torch already has its bundled libnccl.so.2 mapped: dlopen by soname finds that one
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (api.lib) {
            api.getUniqueId = (pfn_ncclGetUniqueId)dlsym(api.lib, "ncclGetUniqueId");
            api.commInitRank = (pfn_ncclCommInitRank)dlsym(api.lib, "ncclCommInitRank");
            api.commDestroy = (pfn_ncclCommDestroy)dlsym(api.lib, "ncclCommDestroy");
            api.allGather = (pfn_ncclAllGather)dlsym(api.lib, "ncclAllGather");
            api.errorString = (pfn_ncclGetErrorString)dlsym(api.lib, "ncclGetErrorString");
        }
        state = (api.lib && api.getUniqueId && api.commInitRank && api.commDestroy && api.allGather) ? 1 : 2;
    }
    return state == 1 ? &api : nullptr;
}
#define CHORD_NCCL(api, expr) do { int r__ = (expr); if (r__ != 0) { \
    chordb::set_error("%s failed: %s", #expr, (api)->errorString ? (api)->errorString(r__) : "NCCL error"); return CHORD_ERR_CUDA; } } while (0)

enum { kPassCount = 7 };

struct RenderSignature {
    uint32_t objectCount, lod0, groups, switchFlags; chord_rect tile; float tanHalfFovy; int mode; bool valid;
};

}  // namespace chordb

using namespace chordb;

struct chord_renderer {
    chord_ctx* ctx;
    uint32_t width, height, flags;
    chord_frame* frame;                                             // the frame of the most recent chord_renderer_render
    chord_frame* frames[2];                                         // async resolve: one frame per history slot; otherwise frames[0] only
    // CHORD_RENDERER_ASYNC_RESOLVE: visibilityMark + lighting of frame N run on `resolveStream` while the context stream
    // already runs the visibility passes of frame N + 1 (they touch disjoint state: the two frames, post[p])
    bool async; cudaStream_t resolveStream; cudaEvent_t geomDone[2], resolveDone[2]; bool resolvePending[2];
    cudaStream_t auxStream; cudaEvent_t auxFork, auxJoin;            // fused exchange: the consumer-side wait runs beside the instance culling
    cudaEvent_t putDone[2]; bool putPending[2]; uint32_t putSeq;    // putSeq mirrors the device-side sequence number (one build per frame)                     // ... and with the asynchronous resolve the producer-side put runs there too
    chord_graph* graphResolve[2];
    chord_cmdlist* post[2]; chord_cmdlist* visible; chord_cmdlist* rejected; chord_cmdlist* visible1;
    uint32_t listCapacity;
    chord_hzb* tmpHzb; chord_hzb* hist[2];
    bool historyValid; uint32_t frameIndex, eagerWithHistory;
    chord_rect tile; bool tiled;
    // recorded frames: one graph per history slot
    chord_graph* graph[2]; cudaEvent_t graphDone[2]; bool graphLaunched[2];
    RenderSignature sig;
    uint64_t graphReplays, graphCaptures;
    // per-pass GPU timers (the reference's timer cut points, renderer.cpp:322-363)
    cudaEvent_t ev[kPassCount + 1]; bool evValid[kPassCount + 1]; bool timersRecorded;
    // exchange: 0 none, 1 fused peer stores, 2 NCCL all-gather
    int mode; uint32_t rank, world; chord_rect tiles[CHORD_MAX_EXCHANGE_RANKS];
    unsigned char* block; size_t blockBytes, chainBytes, headerBytes;
    ExchangePeers* dPeers;                                          // [2]: one table per history slot
    void* opened[CHORD_MAX_EXCHANGE_RANKS]; bool openedIpc[CHORD_MAX_EXCHANGE_RANKS];
    uint32_t* dStatus; uint32_t exchangeMaxPolls;                  // 0 = default (~4 s)
    chord_nccl_comm* nccl; unsigned char* sendBuf; unsigned char* recvBuf; uint64_t slotBytes;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int ensure_lists(chord_renderer* r, uint32_t capacity)
{
    if (r->listCapacity >= capacity && r->post[0]) return CHORD_OK;
    CHORD_NO_RESIZE_IN_CAPTURE(r->ctx, "the renderer's work lists");
    chord_cmdlist** all[5] = { &r->post[0], &r->post[1], &r->visible, &r->rejected, &r->visible1 };
    for (auto l : all) { chord_cmdlist_destroy(*l); *l = nullptr; }
    for (auto l : all) { const int rc = chord_cmdlist_create(r->ctx, capacity, l); if (rc) return rc; }
    r->listCapacity = capacity;
    return CHORD_OK;
}

static void drop_graphs(chord_renderer* r)
{
    for (int p = 0; p < 2; p++) {
        if (r->graph[p]) { chord_graph_destroy(r->graph[p]); r->graph[p] = nullptr; }
        if (r->graphResolve[p]) { chord_graph_destroy(r->graphResolve[p]); r->graphResolve[p] = nullptr; }
        r->graphLaunched[p] = false;
    }
    r->eagerWithHistory = 0;
}

static HzbExchange make_exchange(const chord_renderer* r, uint32_t parity)
{
    HzbExchange ex;
    ex.nPeers = r->world - 1; ex.rank = r->rank; ex.parity = parity;
    ex.dPeers = r->dPeers + parity; ex.localHeader = reinterpret_cast<ExchangeHeader*>(r->block);
    return ex;
}

static int record_event(chord_renderer* r, int i)
{
    if (!(r->flags & CHORD_RENDERER_TIMERS)) return CHORD_OK;
    if (!r->evValid[i]) { CHORD_CUDA(cudaEventCreate(&r->ev[i])); r->evValid[i] = true; }
    // inside a capture only an EXTERNAL record becomes an event-record node that cudaEventElapsedTime can read after a replay
    if (r->ctx->capturing) CHORD_CUDA(cudaEventRecordWithFlags(r->ev[i], r->ctx->stream, cudaEventRecordExternal));
    else                   CHORD_CUDA(cudaEventRecord(r->ev[i], r->ctx->stream));
    return CHORD_OK;
}

static int nccl_allgather_history(chord_renderer* r, chord_hzb* h);

// the frame sequence itself (eager, or while the context is capturing), in two parts: the visibility passes up to the history
// HZB, and the resolve (visibilityMark + lighting)
// Asynchronous renderer + fused exchange: the put of a frame's history (k_hzb_exchange_put) is not part of the visibility
// passes' stream order; it runs on the auxiliary stream as soon as they are done (after_geometry), so that the next frame's
// uniform upload, clear and instance culling do not wait for 20 MB of NVLink stores nobody on THIS GPU needs.
static bool defer_put(const chord_renderer* r) { return r->async && r->mode == 1 && r->auxStream != nullptr; }

static int record_geometry(chord_renderer* r, uint32_t p, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                           const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount, uint32_t lod0, uint32_t groups)
{
    chord_frame* f = r->frame;
    chord_ctx* ctx = r->ctx;
    int rc;
    chord_hzb* last = r->historyValid ? r->hist[p ^ 1u] : nullptr;
    const bool forkWait = last && r->mode == 1 && r->auxStream && !(r->flags & CHORD_RENDERER_TIMERS);
    if (forkWait) {
        // fused exchange, consumer side, on a stream of its own: the wait for the peers' arrival flags (the slowest rank of last
        // frame) and the rebuild of mips 6.. run beside this frame's uniform upload, clear and instance culling, none of which
        // reads the history; the first reader (phase-0 cull below) joins
        CHORD_CUDA(cudaEventRecord(r->auxFork, ctx->stream));
        CHORD_CUDA(cudaStreamWaitEvent(r->auxStream, r->auxFork, 0));
        const HzbExchange ex = make_exchange(r, p ^ 1u);
        cudaStream_t main = ctx->stream;
        ctx->stream = r->auxStream;
        rc = launch_hzb_exchange_wait_top(f, last, &ex, r->world, r->dStatus, r->exchangeMaxPolls);
        ctx->stream = main;
        if (rc) return rc;
        CHORD_CUDA(cudaEventRecord(r->auxJoin, r->auxStream));
    }
    rc = chord_frame_begin(f, view, iview, objects, objectCount, lod0, groups, r->tiled ? &r->tile : nullptr);      // uniforms + addClearGbufferPass
    if (rc) return rc;
    if (last && r->mode == 1 && !forkWait) {
        // fused exchange, consumer side: all peers' parts of last frame's chain have landed; rebuild mips 6.. from them
        const HzbExchange ex = make_exchange(r, p ^ 1u);
        if ((rc = launch_hzb_exchange_wait_top(f, last, &ex, r->world, r->dStatus, r->exchangeMaxPolls))) return rc;
    }
    if ((rc = record_event(r, 0))) return rc;
    rc = (r->flags & CHORD_RENDERER_BVH) ? chord_instance_culling_bvh(f, r->post[p]) : chord_instance_culling(f, r->post[p]);
    if (rc) return rc;
    if (forkWait) CHORD_CUDA(cudaStreamWaitEvent(ctx->stream, r->auxJoin, 0));
    if ((rc = record_event(r, 1))) return rc;
    int need1 = 0;
    if ((rc = chord_visibility_stage0(f, last, r->post[p], r->visible, r->rejected, &need1))) return rc;     // mesh_raster.cpp:269-312
    if ((rc = record_event(r, 2))) return rc;
    if (need1) {                                                                                                // bShouldStage1GLTF
        if ((rc = chord_build_hzb(f, 1, 0, 0, r->tmpHzb))) return rc;
        if ((rc = record_event(r, 3))) return rc;
        if ((rc = chord_visibility_stage1(f, r->tmpHzb, r->rejected, r->visible1))) return rc;
    } else if ((rc = record_event(r, 3))) return rc;
    if ((rc = record_event(r, 4))) return rc;
    chord_hzb* hzb = r->hist[p];
    if (r->mode == 1) {
        const HzbExchange ex = make_exchange(r, p);
        if ((rc = launch_build_hzb(f, 1, 1, 1, hzb, &ex, defer_put(r)))) return rc;
        hzb->valid = 1;
    } else {
        if ((rc = chord_build_hzb(f, 1, 1, 1, hzb))) return rc;
        if (r->mode == 2 && (rc = nccl_allgather_history(r, hzb))) return rc;
    }
    if ((rc = record_event(r, 5))) return rc;
    // asynchronous resolve: the tile classification (memory-latency bound, 8 B / pixel) stays on the context stream with the
    // other visibility passes; the second stream gets the ALU-bound part only
    if (r->async) {
        if ((rc = chord_visibility_mark(f, r->post[p]))) return rc;
        if ((rc = launch_lighting(f, r->post[p], CHORD_LIGHTING_PREPARE))) return rc;      // tile list + resolve records: same stream
    }
    return CHORD_OK;
}

static int record_resolve(chord_renderer* r, uint32_t p)
{
    chord_frame* f = r->frame;
    int rc;
    if (r->flags & CHORD_RENDERER_RESOLVE) {
        if (!r->async && (rc = chord_visibility_mark(f, r->post[p]))) return rc;
        if ((rc = record_event(r, 6))) return rc;
        if (r->async) {
            CHORD_REQUIRE(f->ctx->lutTransmittance && f->ctx->lutIrradiance, "atmosphere LUTs not uploaded");
            if ((rc = launch_lighting(f, r->post[p], CHORD_LIGHTING_SHADE))) return rc;
        } else if ((rc = chord_lighting(f, r->post[p]))) return rc;
    } else if ((rc = record_event(r, 6))) return rc;
    if ((rc = record_event(r, 7))) return rc;
    r->timersRecorded = (r->flags & CHORD_RENDERER_TIMERS) != 0u;
    return CHORD_OK;
}

static int record_frame(chord_renderer* r, uint32_t p, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                        const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount, uint32_t lod0, uint32_t groups)
{
    const int rc = record_geometry(r, p, view, iview, objects, objectCount, lod0, groups);
    return rc ? rc : record_resolve(r, p);
}

// asynchronous renderer: what follows the visibility passes of slot p once they are enqueued on the context stream
static int after_geometry(chord_renderer* r, uint32_t p)
{
    chord_ctx* ctx = r->ctx;
    CHORD_CUDA(cudaEventRecord(r->geomDone[p], ctx->stream));
    if (defer_put(r)) {
        CHORD_CUDA(cudaStreamWaitEvent(r->auxStream, r->geomDone[p], 0));
        const HzbExchange ex = make_exchange(r, p);
        cudaStream_t main = ctx->stream;
        ctx->stream = r->auxStream;
        const int rc = launch_hzb_exchange_put(r->frame, 1, r->hist[p], &ex, ++r->putSeq);
        ctx->stream = main;
        if (rc) return rc;
        CHORD_CUDA(cudaEventRecord(r->putDone[p], r->auxStream));
        r->putPending[p] = true;
    }
    CHORD_CUDA(cudaStreamWaitEvent(r->resolveStream, r->geomDone[p], 0));
    return CHORD_OK;
}

// Runs `fn` with the context's stream pointing at the resolve stream: every launcher of the library enqueues on ctx->stream, so
// this is how a pass is moved to the second stream (host code is single-threaded per context).
template <typename Fn>
static int on_resolve_stream(chord_renderer* r, Fn fn)
{
    cudaStream_t main = r->ctx->stream;
    r->ctx->stream = r->resolveStream;
    const int rc = fn();
    r->ctx->stream = main;
    return rc;
}

static int nccl_allgather_history(chord_renderer* r, chord_hzb* h)
{
    NcclApi* api = nccl_api();
    CHORD_REQUIRE(api && r->nccl && r->nccl->comm, "NCCL exchange is not connected");
    for (int which = 0; which < 2; which++) {
        int rc = chord_hzb_pack_tile(h, which, &r->tiles[r->rank], r->sendBuf);
        if (rc) return rc;
        CHORD_NCCL(api, api->allGather(r->sendBuf, r->recvBuf, (size_t)r->slotBytes, /* ncclUint8 */ 1, r->nccl->comm, r->ctx->stream));
        if ((rc = chord_hzb_unpack_tiles(h, which, r->tiles, r->world, r->recvBuf, r->slotBytes))) return rc;
    }
    // the tiles' valid depth ranges: one more (8-byte) all-gather, folded on the device
    CHORD_CUDA(cudaMemcpyAsync(r->sendBuf, h->validRange, 8, cudaMemcpyDeviceToDevice, r->ctx->stream));
    CHORD_NCCL(api, api->allGather(r->sendBuf, r->recvBuf, 8, /* ncclUint8 */ 1, r->nccl->comm, r->ctx->stream));
    const int rc = launch_hzb_range_reduce(r->ctx, reinterpret_cast<const uint32_t*>(r->recvBuf), r->world, h->validRange);
    if (rc) return rc;
    return chord_hzb_finish_top(r->frame, 1, 1, h);
}

extern "C" {

int chord_renderer_create(chord_ctx* ctx, uint32_t width, uint32_t height, uint32_t flags, chord_renderer** out)
{
    CHORD_REQUIRE(ctx && out, "null");
    *out = nullptr;
    chord_renderer* r = (chord_renderer*)calloc(1, sizeof(chord_renderer));
    CHORD_REQUIRE(r != nullptr, "out of host memory");
    r->ctx = ctx; r->width = width; r->height = height; r->flags = flags;
    ctx_retain(ctx);
    r->tile = chord_rect{0, 0, width, height};
    const int rc = [&]() -> int {
        int e;
        // timers serialise the passes they measure: the asynchronous resolve is for renderers without them
        r->async = (flags & CHORD_RENDERER_ASYNC_RESOLVE) && (flags & CHORD_RENDERER_RESOLVE) && !(flags & CHORD_RENDERER_TIMERS);
        for (int k = 0; k < (r->async ? 2 : 1); k++) {
            if ((e = chord_frame_create(ctx, width, height, &r->frames[k]))) return e;
            if ((e = chord_frame_set_gbuffer_packed(r->frames[k], (flags & CHORD_RENDERER_PACKED_GBUFFER) ? 1 : 0))) return e;
        }
        r->frame = r->frames[0];
        if (r->async) {
            // the resolve yields to the context stream (created with the greatest priority, chord_ctx_create): CTAs of the
            // latency-bound visibility passes are scheduled ahead of the resolve's whenever both are pending
            int least = 0, greatest = 0;
            CHORD_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CHORD_CUDA(cudaStreamCreateWithPriority(&r->resolveStream, cudaStreamNonBlocking, least));
            for (int k = 0; k < 2; k++) {
                CHORD_CUDA(cudaEventCreateWithFlags(&r->geomDone[k], cudaEventDisableTiming));
                CHORD_CUDA(cudaEventCreateWithFlags(&r->resolveDone[k], cudaEventDisableTiming));
            }
        }
        if ((e = chord_hzb_create(ctx, width, height, &r->tmpHzb))) return e;
        if ((e = chord_hzb_create(ctx, width, height, &r->hist[0]))) return e;
        if ((e = chord_hzb_create(ctx, width, height, &r->hist[1]))) return e;
        CHORD_CUDA(cudaEventCreateWithFlags(&r->graphDone[0], cudaEventDisableTiming));
        CHORD_CUDA(cudaEventCreateWithFlags(&r->graphDone[1], cudaEventDisableTiming));
        CHORD_CUDA(cudaMalloc(&r->dStatus, 16));
        CHORD_CUDA(cudaMemsetAsync(r->dStatus, 0, 16, ctx->stream));
        return CHORD_OK;
    }();
    if (rc != CHORD_OK) { chord_renderer_destroy(r); return rc; }
    *out = r;
    return CHORD_OK;
}

static void exchange_teardown(chord_renderer* r)
{
    for (uint32_t q = 0; q < CHORD_MAX_EXCHANGE_RANKS; q++) {
        if (r->opened[q] && r->openedIpc[q]) cudaIpcCloseMemHandle(r->opened[q]);
        r->opened[q] = nullptr; r->openedIpc[q] = false;
    }
    cudaFree(r->dPeers); r->dPeers = nullptr;
    cudaFree(r->sendBuf); cudaFree(r->recvBuf); r->sendBuf = r->recvBuf = nullptr;
    r->mode = 0;
}

void chord_renderer_destroy(chord_renderer* r)
{
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    if (r->resolveStream) cudaStreamSynchronize(r->resolveStream);
    drop_graphs(r);
    exchange_teardown(r);
    chord_frame_destroy(r->frames[0]); chord_frame_destroy(r->frames[1]);
    if (r->resolveStream) cudaStreamDestroy(r->resolveStream);
    if (r->auxStream) { cudaStreamSynchronize(r->auxStream); cudaStreamDestroy(r->auxStream); cudaEventDestroy(r->auxFork); cudaEventDestroy(r->auxJoin); cudaEventDestroy(r->putDone[0]); cudaEventDestroy(r->putDone[1]); }
    for (int k = 0; k < 2; k++) { if (r->geomDone[k]) cudaEventDestroy(r->geomDone[k]); if (r->resolveDone[k]) cudaEventDestroy(r->resolveDone[k]); }
    chord_cmdlist** all[5] = { &r->post[0], &r->post[1], &r->visible, &r->rejected, &r->visible1 };
    for (auto l : all) chord_cmdlist_destroy(*l);
    chord_hzb_destroy(r->tmpHzb); chord_hzb_destroy(r->hist[0]); chord_hzb_destroy(r->hist[1]);
    cudaFree(r->block);
    cudaFree(r->dStatus);
    for (int p = 0; p < 2; p++) if (r->graphDone[p]) cudaEventDestroy(r->graphDone[p]);
    for (int i = 0; i <= kPassCount; i++) if (r->evValid[i]) cudaEventDestroy(r->ev[i]);
    ctx_release(r->ctx);
    free(r);
}

int chord_renderer_set_tile(chord_renderer* r, const chord_rect* tile)
{
    CHORD_REQUIRE(r, "null");
    if (tile) {
        CHORD_REQUIRE(tile->x0 < tile->x1 && tile->y0 < tile->y1 && tile->x1 <= r->width && tile->y1 <= r->height, "tile rect");
        r->tile = *tile;
        r->tiled = !(tile->x0 == 0 && tile->y0 == 0 && tile->x1 == r->width && tile->y1 == r->height);
    } else {
        r->tile = chord_rect{0, 0, r->width, r->height}; r->tiled = false;
    }
    return CHORD_OK;
}

int chord_renderer_clear_history(chord_renderer* r)
{
    // renderer.cpp:95-105 (resize / camera cut): the next frame culls without occlusion history
    CHORD_REQUIRE(r, "null");
    r->historyValid = false;
    r->hist[0]->valid = r->hist[1]->valid = 0;
    drop_graphs(r);
    return CHORD_OK;
}

int chord_renderer_render(chord_renderer* r, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                          const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount, uint32_t lod0MeshletCount, uint32_t meshletGroupCount)
{
    CHORD_REQUIRE(r && view && iview && objects, "null");
    chord_ctx* ctx = r->ctx;
    CHORD_REQUIRE(!ctx->capturing, "chord_renderer_render records its own graphs: do not call it inside chord_graph_begin_capture");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_lists(r, lod0MeshletCount);
    if (rc) return rc;
    const uint32_t p = r->frameIndex & 1u;
    RenderSignature s;
    memset(&s, 0, sizeof(s));                                                   // (compared with memcmp: no stray padding)
    s.objectCount = objectCount; s.lod0 = lod0MeshletCount; s.groups = meshletGroupCount; s.switchFlags = ctx->switchFlags;
    s.tile = r->tile; s.tanHalfFovy = tanf(0.5f * view->cameraFovy); s.mode = r->mode; s.valid = true;
    const bool same = r->sig.valid && memcmp(&s, &r->sig, sizeof(s)) == 0;
    if (!same) { drop_graphs(r); memset(&r->sig, 0, sizeof(r->sig)); r->sig = s; }
    const bool wantGraph = (r->flags & CHORD_RENDERER_GRAPH) != 0u && r->historyValid;
    if (r->async) {
        // this slot's frame was last read by the resolve of two frames ago, which may still be running on the resolve stream
        r->frame = r->frames[p];
        if (r->resolvePending[p]) CHORD_CUDA(cudaStreamWaitEvent(ctx->stream, r->resolveDone[p], 0));
        // ... and its history chain by the put of two frames ago (auxiliary stream)
        if (r->putPending[p]) CHORD_CUDA(cudaStreamWaitEvent(ctx->stream, r->putDone[p], 0));
    }
    if (wantGraph && r->graph[p]) {
        // replay: rewrite the pinned uniform snapshot once the previous replay of THIS graph (two frames ago) is done with it
        if (r->graphLaunched[p]) CHORD_CUDA(cudaEventSynchronize(r->graphDone[p]));
        if ((rc = chord_graph_update_frame_inputs(r->graph[p], r->frame, view, iview, objects, objectCount))) return rc;
        if ((rc = chord_graph_launch(r->graph[p]))) return rc;
        CHORD_CUDA(cudaEventRecord(r->graphDone[p], ctx->stream));
        r->graphLaunched[p] = true; r->graphReplays++;
        if (r->async) {
            if ((rc = after_geometry(r, p))) return rc;
            if ((rc = on_resolve_stream(r, [&]() { return chord_graph_launch(r->graphResolve[p]); }))) return rc;
        }
    } else if (wantGraph && r->eagerWithHistory >= (r->async ? 2u : 1u)) {
        // every lazily sized buffer (of this slot's frame, when there are two) has seen a frame with history: record this slot's
        // graph(s), then run them
        if ((rc = chord_graph_begin_capture(ctx))) return rc;
        rc = r->async ? record_geometry(r, p, view, iview, objects, objectCount, lod0MeshletCount, meshletGroupCount)
                      : record_frame(r, p, view, iview, objects, objectCount, lod0MeshletCount, meshletGroupCount);
        chord_graph* g = nullptr;
        const int rc2 = chord_graph_end_capture(ctx, &g);
        if (rc) { if (g) chord_graph_destroy(g); return rc; }
        if (rc2) return rc2;
        r->graph[p] = g; r->graphCaptures++;
        if (r->async) {
            chord_graph* gr = nullptr;
            rc = on_resolve_stream(r, [&]() -> int {
                int e = chord_graph_begin_capture(ctx);
                if (e) return e;
                e = record_resolve(r, p);
                const int e2 = chord_graph_end_capture(ctx, &gr);
                return e ? e : e2;
            });
            if (rc) { if (gr) chord_graph_destroy(gr); return rc; }
            r->graphResolve[p] = gr;
        }
        if ((rc = chord_graph_launch(g))) return rc;
        CHORD_CUDA(cudaEventRecord(r->graphDone[p], ctx->stream));
        r->graphLaunched[p] = true;
        if (r->async) {
            if ((rc = after_geometry(r, p))) return rc;
            if ((rc = on_resolve_stream(r, [&]() { return chord_graph_launch(r->graphResolve[p]); }))) return rc;
        }
    } else if (r->async) {
        if ((rc = record_geometry(r, p, view, iview, objects, objectCount, lod0MeshletCount, meshletGroupCount))) return rc;
        if ((rc = after_geometry(r, p))) return rc;
        if ((rc = on_resolve_stream(r, [&]() { return record_resolve(r, p); }))) return rc;
        if (r->historyValid) r->eagerWithHistory++;
    } else {
        if ((rc = record_frame(r, p, view, iview, objects, objectCount, lod0MeshletCount, meshletGroupCount))) return rc;
        if (r->historyValid) r->eagerWithHistory++;
    }
    if (r->async) {
        CHORD_CUDA(cudaEventRecord(r->resolveDone[p], r->resolveStream));
        r->resolvePending[p] = true;
    }
    r->hist[p]->valid = 1;
    r->historyValid = true;                                                     // m_rendererHistory.hzbCtx = hzbCtx (renderer.cpp:489)
    r->frameIndex++;
    return CHORD_OK;
}

chord_frame* chord_renderer_frame(chord_renderer* r)
{
    if (!r) return nullptr;
    // asynchronous resolve: whatever the caller enqueues on the context stream next (read-backs, its own passes) is ordered
    // after the resolve of the frame it asks for
    if (r->async && r->frameIndex) {
        const uint32_t p = (r->frameIndex - 1u) & 1u;
        if (r->resolvePending[p]) { cudaSetDevice(r->ctx->device); cudaStreamWaitEvent(r->ctx->stream, r->resolveDone[p], 0); }
    }
    return r->frame;
}
chord_cmdlist* chord_renderer_post_cull(chord_renderer* r) { return (r && r->frameIndex) ? r->post[(r->frameIndex - 1u) & 1u] : nullptr; }
chord_hzb* chord_renderer_history_hzb(chord_renderer* r) { return (r && r->historyValid) ? r->hist[(r->frameIndex - 1u) & 1u] : nullptr; }
uint64_t chord_renderer_graph_replays(const chord_renderer* r) { return r ? r->graphReplays : 0; }

int chord_renderer_history_sync(chord_renderer* r)
{
    // makes last frame's history chain complete NOW (tests / read-backs): the fused exchange defers the wait for the peers and
    // the rebuild of mips 6.. to the start of the next frame
    CHORD_REQUIRE(r, "null");
    if (r->mode != 1 || !r->historyValid) return CHORD_OK;
    const uint32_t p = (r->frameIndex - 1u) & 1u;
    if (r->putPending[p]) CHORD_CUDA(cudaStreamWaitEvent(r->ctx->stream, r->putDone[p], 0));      // this rank's own put advances its sequence number
    const HzbExchange ex = make_exchange(r, p);
    return launch_hzb_exchange_wait_top(r->frame, r->hist[p], &ex, r->world, r->dStatus, r->exchangeMaxPolls);
}

int chord_renderer_set_exchange_timeout(chord_renderer* r, uint32_t maxPolls)
{
    // the bound is a kernel parameter of the consumer-side wait: recorded frames are dropped so that they pick the new value up
    CHORD_REQUIRE(r, "null");
    r->exchangeMaxPolls = maxPolls;
    drop_graphs(r);
    return CHORD_OK;
}

int chord_renderer_exchange_status(chord_renderer* r, uint32_t* outTimedOut)
{
    CHORD_REQUIRE(r && outTimedOut, "null");
    CHORD_CUDA(cudaMemcpyAsync(outTimedOut, r->dStatus, 4, cudaMemcpyDeviceToHost, r->ctx->stream));
    CHORD_CUDA(cudaStreamSynchronize(r->ctx->stream));
    return CHORD_OK;
}

int chord_renderer_pass_times(chord_renderer* r, float* outMicroseconds)
{
    // GPU time of the seven passes of the last frame at the reference's timer cut points (renderer.cpp:322-363: "GLTF Instance
    // Culling", "GLTF Visibility Stage0", "BuildHZB Post Prepass Stage0", "GLTF Visibility Stage1", "BuildHZB",
    // "VisibilityTileMarker", "lightingTile"); needs CHORD_RENDERER_TIMERS.  Synchronises the stream.
    CHORD_REQUIRE(r && outMicroseconds, "null");
    CHORD_REQUIRE((r->flags & CHORD_RENDERER_TIMERS) && r->timersRecorded, "the renderer was created without CHORD_RENDERER_TIMERS or has not rendered yet");
    CHORD_CUDA(cudaStreamSynchronize(r->ctx->stream));
    for (int i = 0; i < kPassCount; i++) {
        float ms = 0.0f;
        CHORD_CUDA(cudaEventElapsedTime(&ms, r->ev[i], r->ev[i + 1]));
        outMicroseconds[i] = ms * 1000.0f;
    }
    return CHORD_OK;
}

// ---- exchange set-up --------------------------------------------------------------------------------------------------
int chord_renderer_exchange_export(chord_renderer* r, chord_ipc_handle* outHandle, void** outLocalPtr)
{
    CHORD_REQUIRE(r, "null");
    chord_ctx* ctx = r->ctx;
    CHORD_CUDA(cudaSetDevice(ctx->device));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    drop_graphs(r);
    if (!r->block) {
        r->headerBytes = align_up(sizeof(ExchangeHeader), 256);
        r->chainBytes = align_up(sizeof(__half) * (size_t)r->hist[0]->totalTexels, 256);
        r->blockBytes = r->headerBytes + 4 * r->chainBytes;
        CHORD_CUDA(cudaMalloc(&r->block, r->blockBytes));                     // plain cudaMalloc: exportable with CUDA IPC
        CHORD_CUDA(cudaMemset(r->block, 0, r->blockBytes));
        // the two history HZBs move into the block: slot p = [min | max] at header + 2 p chains
        for (int p = 0; p < 2; p++) {
            chord_hzb* h = r->hist[p];
            if (!h->externalChains) { cudaFree(h->minChain); cudaFree(h->maxChain); }
            h->minChain = reinterpret_cast<__half*>(r->block + r->headerBytes + (size_t)(2 * p) * r->chainBytes);
            h->maxChain = reinterpret_cast<__half*>(r->block + r->headerBytes + (size_t)(2 * p + 1) * r->chainBytes);
            h->externalChains = true; h->valid = 0;
        }
        r->historyValid = false;
    }
    if (outHandle) {
        cudaIpcMemHandle_t h;
        CHORD_CUDA(cudaIpcGetMemHandle(&h, r->block));
        static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(chord_ipc_handle), "CUDA IPC handle size");
        memcpy(outHandle, &h, sizeof(h));
    }
    if (outLocalPtr) *outLocalPtr = r->block;
    return CHORD_OK;
}

static int set_tiles(chord_renderer* r, uint32_t rank, uint32_t world, const chord_rect* tiles)
{
    CHORD_REQUIRE(world >= 2 && world <= CHORD_MAX_EXCHANGE_RANKS && rank < world && tiles, "rank / world / tiles");
    for (uint32_t q = 0; q < world; q++) {
        const chord_rect& t = tiles[q];
        CHORD_REQUIRE(t.x0 < t.x1 && t.y0 < t.y1 && t.x1 <= r->width && t.y1 <= r->height, "tile rect");
        CHORD_REQUIRE(t.x0 % 64 == 0 && t.y0 % 64 == 0 && (t.x1 % 64 == 0 || t.x1 == r->width) && (t.y1 % 64 == 0 || t.y1 == r->height), "tiles must be cut on multiples of 64 px");
        r->tiles[q] = t;
    }
    r->rank = rank; r->world = world;
    return chord_renderer_set_tile(r, &tiles[rank]);
}

int chord_renderer_exchange_connect(chord_renderer* r, uint32_t rank, uint32_t world, const chord_rect* tiles,
                                    const chord_ipc_handle* handles, void* const* localPtrs)
{
    CHORD_REQUIRE(r && (handles || localPtrs), "null");
    CHORD_REQUIRE(r->block, "call chord_renderer_exchange_export first");
    chord_ctx* ctx = r->ctx;
    CHORD_CUDA(cudaSetDevice(ctx->device));
    CHORD_CUDA(cudaStreamSynchronize(ctx->stream));
    exchange_teardown(r);
    drop_graphs(r);
    int rc = set_tiles(r, rank, world, tiles);
    if (rc) return rc;
    ExchangePeers host[2];
    memset(host, 0, sizeof(host));
    uint32_t k = 0;
    for (uint32_t q = 0; q < world; q++) {
        if (q == rank) continue;
        unsigned char* base = nullptr;
        if (handles) {
            cudaIpcMemHandle_t h; memcpy(&h, &handles[q], sizeof(h));
            void* ptr = nullptr;
            CHORD_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            r->opened[q] = ptr; r->openedIpc[q] = true; base = (unsigned char*)ptr;
        } else {
            base = (unsigned char*)localPtrs[q];
            CHORD_REQUIRE(base != nullptr, "peer block pointer");
            cudaPointerAttributes at;
            CHORD_CUDA(cudaPointerGetAttributes(&at, base));
            if (at.device != ctx->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d) failed: %s", at.device, cudaGetErrorString(e)); return CHORD_ERR_CUDA; }
                cudaGetLastError();
            }
            r->opened[q] = base; r->openedIpc[q] = false;
        }
        for (int p = 0; p < 2; p++) {
            host[p].header[k] = reinterpret_cast<ExchangeHeader*>(base);
            host[p].min[k] = reinterpret_cast<__half*>(base + r->headerBytes + (size_t)(2 * p) * r->chainBytes);
            host[p].max[k] = reinterpret_cast<__half*>(base + r->headerBytes + (size_t)(2 * p + 1) * r->chainBytes);
        }
        k++;
    }
    CHORD_CUDA(cudaMalloc(&r->dPeers, sizeof(host)));
    CHORD_CUDA(cudaMemcpy(r->dPeers, host, sizeof(host), cudaMemcpyHostToDevice));
    CHORD_CUDA(cudaMemset(r->dStatus, 0, 16));
    if (!r->auxStream) {
        int least = 0, greatest = 0;
        CHORD_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CHORD_CUDA(cudaStreamCreateWithPriority(&r->auxStream, cudaStreamNonBlocking, greatest));
        CHORD_CUDA(cudaEventCreateWithFlags(&r->auxFork, cudaEventDisableTiming));
        CHORD_CUDA(cudaEventCreateWithFlags(&r->auxJoin, cudaEventDisableTiming));
        CHORD_CUDA(cudaEventCreateWithFlags(&r->putDone[0], cudaEventDisableTiming));
        CHORD_CUDA(cudaEventCreateWithFlags(&r->putDone[1], cudaEventDisableTiming));
    }
    r->putPending[0] = r->putPending[1] = false;
    // the host mirror of this rank's sequence number starts where the device-side word stands
    CHORD_CUDA(cudaMemcpy(&r->putSeq, &reinterpret_cast<ExchangeHeader*>(r->block)->localSeq, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    r->mode = 1;
    r->historyValid = false; r->hist[0]->valid = r->hist[1]->valid = 0; r->frameIndex = 0;
    return CHORD_OK;
}

// ---- NCCL ---------------------------------------------------------------------------------------------------------------
int chord_nccl_unique_id(unsigned char outId[128])
{
    CHORD_REQUIRE(outId, "null");
    NcclApi* api = nccl_api();
    CHORD_REQUIRE(api, "libnccl.so.2 could not be loaded");
    NcclUniqueId id;
    CHORD_NCCL(api, api->getUniqueId(&id));
    memcpy(outId, &id, 128);
    return CHORD_OK;
}

int chord_nccl_comm_create(chord_ctx* ctx, uint32_t world, uint32_t rank, const unsigned char id128[128], chord_nccl_comm** out)
{
    CHORD_REQUIRE(ctx && id128 && out && rank < world, "args");
    *out = nullptr;
    NcclApi* api = nccl_api();
    CHORD_REQUIRE(api, "libnccl.so.2 could not be loaded");
    CHORD_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueId id; memcpy(&id, id128, 128);
    void* comm = nullptr;
    CHORD_NCCL(api, api->commInitRank(&comm, (int)world, id, (int)rank));
    chord_nccl_comm* c = (chord_nccl_comm*)calloc(1, sizeof(chord_nccl_comm));
    c->ctx = ctx; c->comm = comm; c->rank = rank; c->world = world;
    ctx_retain(ctx);
    *out = c;
    return CHORD_OK;
}

void chord_nccl_comm_destroy(chord_nccl_comm* c)
{
    if (!c) return;
    NcclApi* api = nccl_api();
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (api && c->comm) api->commDestroy(c->comm);
    ctx_release(c->ctx);
    free(c);
}

int chord_hzb_allgather(chord_frame* frame, chord_hzb* hzb, const chord_rect* tiles, chord_nccl_comm* comm, int bMin, int bMax, void* staging, uint64_t stagingBytes)
{
    // pack this rank's tile part of mips 0..5 -> ncclAllGather on the context stream -> unpack every rank's part -> rebuild
    // mips 6..: the gathered chain is bit-identical to the single-GPU chain.  staging: device memory, (world + 1) slots.
    CHORD_REQUIRE(frame && hzb && tiles && comm && staging && (bMin || bMax), "args");
    NcclApi* api = nccl_api();
    CHORD_REQUIRE(api, "libnccl.so.2 could not be loaded");
    uint64_t slot = 0;
    for (uint32_t q = 0; q < comm->world; q++) { uint64_t b = 0; const int rc = chord_hzb_tile_bytes(hzb, &tiles[q], &b); if (rc) return rc; if (b > slot) slot = b; }
    slot = align_up(slot, 256);
    CHORD_REQUIRE(stagingBytes >= slot * (comm->world + 1u), "staging buffer too small: (world + 1) x the largest tile part, 256-byte aligned");
    unsigned char* send = (unsigned char*)staging; unsigned char* recv = send + slot;
    for (int which = 0; which < 2; which++) {
        if (!(which ? bMax : bMin)) continue;
        int rc = chord_hzb_pack_tile(hzb, which, &tiles[comm->rank], send);
        if (rc) return rc;
        CHORD_NCCL(api, api->allGather(send, recv, (size_t)slot, 1, comm->comm, frame->ctx->stream));
        if ((rc = chord_hzb_unpack_tiles(hzb, which, tiles, comm->world, recv, slot))) return rc;
    }
    const int rc = chord_hzb_finish_top(frame, bMin, bMax, hzb);
    if (rc == CHORD_OK) hzb->valid = 1;
    return rc;
}

int chord_renderer_exchange_connect_nccl(chord_renderer* r, const chord_rect* tiles, chord_nccl_comm* comm)
{
    CHORD_REQUIRE(r && tiles && comm && comm->ctx == r->ctx, "null / communicator of another context");
    CHORD_CUDA(cudaSetDevice(r->ctx->device));
    CHORD_CUDA(cudaStreamSynchronize(r->ctx->stream));
    exchange_teardown(r);
    drop_graphs(r);
    int rc = set_tiles(r, comm->rank, comm->world, tiles);
    if (rc) return rc;
    uint64_t slot = 0;
    for (uint32_t q = 0; q < comm->world; q++) { uint64_t b = 0; if ((rc = chord_hzb_tile_bytes(r->hist[0], &tiles[q], &b))) return rc; if (b > slot) slot = b; }
    r->slotBytes = align_up(slot, 256);
    CHORD_CUDA(cudaMalloc(&r->sendBuf, r->slotBytes));
    CHORD_CUDA(cudaMalloc(&r->recvBuf, r->slotBytes * comm->world));
    r->nccl = comm; r->mode = 2;
    r->historyValid = false; r->hist[0]->valid = r->hist[1]->valid = 0; r->frameIndex = 0;
    return CHORD_OK;
}

}  // extern "C"
