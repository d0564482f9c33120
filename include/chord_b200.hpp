// chord_b200.hpp — header-only C++17 shim over the C ABI (chord_b200.h), shaped like the reference's pass-graph API
// (source/renderer/mesh/gltf_rendering.h:37-108, postprocessing/postprocessing.h:41-54, visibility_tile.h:22-34,
// lighting.h:37).  RAII handles replace chord's ref-counted pool resources; errors throw std::runtime_error carrying
// chord_last_error() (the reference would check()-crash).  No CUDA headers needed by the includer.
// Lifetime: the C ABI ref-counts the context from every child handle (chord_b200.h, chord_ctx_destroy), and the classes
// below keep only the chord_ctx* - never a reference to the Context OBJECT - so a Frame, list, HZB, Graph or
// DeferredRenderer may outlive the Context it was created from, exactly like a PoolTextureRef outliving its pool user.
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "chord_b200.h"

namespace chord_b200 {

inline void check(int rc, const char* what)
{
    if (rc != CHORD_OK) throw std::runtime_error(std::string(what) + ": " + chord_last_error());
}

struct CmdListDeleter { void operator()(chord_cmdlist* p) const { chord_cmdlist_destroy(p); } };
struct HzbDeleter { void operator()(chord_hzb* p) const { chord_hzb_destroy(p); } };
using CmdListRef = std::shared_ptr<chord_cmdlist>;     // CountAndCmdBuffer
using HzbRef = std::shared_ptr<chord_hzb>;             // HZBContext

inline CmdListRef makeCmdList(chord_ctx* ctx, uint32_t capacity)
{
    chord_cmdlist* l = nullptr;
    check(chord_cmdlist_create(ctx, capacity, &l), "chord_cmdlist_create");
    return CmdListRef(l, CmdListDeleter{});
}
inline HzbRef makeHzb(chord_ctx* ctx, uint32_t w, uint32_t h)
{
    chord_hzb* z = nullptr;
    check(chord_hzb_create(ctx, w, h, &z), "chord_hzb_create");
    return HzbRef(z, HzbDeleter{});
}

class Frame;

class Context {
public:
    explicit Context(int device) { check(chord_ctx_create(device, &m_ctx), "chord_ctx_create"); }
    ~Context() { chord_ctx_destroy(m_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;

    void uploadScene(const chord_scene_desc& desc) { check(chord_scene_upload(m_ctx, &desc), "chord_scene_upload"); }
    // material textures of the masked bucket (bindless Texture2D / SamplerState ids -> table indices)
    void uploadTextures(const chord_texture_desc* tex, uint32_t nTex, const chord_sampler_desc* smp, uint32_t nSmp)
    { check(chord_textures_upload(m_ctx, tex, nTex, smp, nSmp), "chord_textures_upload"); }
    // AtmosphereLut of the lighting pass: transmittance 256 x 64 + irradiance 64 x 16 (PBR turn), scattering / optional single-Mie
    // volumes 256 x 128 x 32 (sky turn; leave them out and texels without geometry keep the cleared value), all RGBA16F
    void uploadAtmosphereLuts(const uint16_t* transmittance, const uint16_t* irradiance)
    { const chord_atmosphere_luts l = { transmittance, irradiance }; check(chord_atmosphere_luts_upload(m_ctx, &l), "chord_atmosphere_luts_upload"); }
    void uploadScattering(const uint16_t* scattering, const uint16_t* singleMie = nullptr)
    { check(chord_atmosphere_scattering_upload(m_ctx, scattering, singleMie), "chord_atmosphere_scattering_upload"); }
    void setSwitchFlags(uint32_t f) { check(chord_ctx_set_switch_flags(m_ctx, f), "chord_ctx_set_switch_flags"); }
    void sync() { check(chord_ctx_sync(m_ctx), "chord_ctx_sync"); }
    void* stream() const { return chord_ctx_stream(m_ctx); }
    chord_ctx* handle() const { return m_ctx; }

    CmdListRef createCmdList(uint32_t capacity) const { return makeCmdList(m_ctx, capacity); }
    HzbRef createHzb(uint32_t w, uint32_t h) const { return makeHzb(m_ctx, w, h); }
    inline std::unique_ptr<Frame> createFrame(uint32_t w, uint32_t h) const;

private:
    chord_ctx* m_ctx = nullptr;
};

class Frame {
public:
    Frame(const Context& ctx, uint32_t w, uint32_t h) : m_ctx(ctx.handle()), m_w(w), m_h(h)
    { check(chord_frame_create(m_ctx, w, h, &m_frame), "chord_frame_create"); }
    chord_ctx* context() const { return m_ctx; }
    ~Frame() { chord_frame_destroy(m_frame); }
    Frame(const Frame&) = delete;
    Frame& operator=(const Frame&) = delete;
    uint32_t width() const { return m_w; }
    uint32_t height() const { return m_h; }
    chord_frame* handle() const { return m_frame; }

    // uniforms + addClearGbufferPass
    void begin(const chord_PerframeCameraView& view, const chord_InstanceCullingViewInfo& iview,
               const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount,
               uint32_t lod0MeshletCount, uint32_t meshletGroupCount, const chord_rect* tile = nullptr)
    {
        m_lod0 = lod0MeshletCount;
        check(chord_frame_begin(m_frame, &view, &iview, objects, objectCount, lod0MeshletCount, meshletGroupCount, tile), "chord_frame_begin");
    }
    // instanceCulling
    CmdListRef instanceCulling()
    {
        auto out = makeCmdList(m_ctx, m_lod0);
        check(chord_instance_culling(m_frame, out.get()), "chord_instance_culling");
        return out;
    }
    // gltfVisibilityRenderingStage0: returns bShouldInvokeStage1, fills outRejected
    bool visibilityStage0(const chord_hzb* lastFrameHzb, const CmdListRef& in, CmdListRef& outRejected)
    {
        int need = 0;
        CmdListRef scratch, rejected;
        if (lastFrameHzb && chord_hzb_is_valid(lastFrameHzb)) { scratch = makeCmdList(m_ctx, m_lod0); rejected = makeCmdList(m_ctx, m_lod0); }
        check(chord_visibility_stage0(m_frame, lastFrameHzb, in.get(), scratch.get(), rejected.get(), &need), "chord_visibility_stage0");
        m_keep[0] = scratch;                       // stream-ordered: keep alive until the next frame
        outRejected = need ? rejected : CmdListRef();
        return need != 0;
    }
    // gltfVisibilityRenderingStage1
    void visibilityStage1(const HzbRef& thisFrameHzb, const CmdListRef& rejected)
    {
        auto scratch = makeCmdList(m_ctx, m_lod0);
        check(chord_visibility_stage1(m_frame, thisFrameHzb.get(), rejected.get(), scratch.get()), "chord_visibility_stage1");
        m_keep[1] = scratch; m_keep[2] = rejected;
    }
    // buildHZB
    HzbRef buildHZB(bool bMin, bool bMax, bool bValidRange)
    {
        auto h = makeHzb(m_ctx, m_w, m_h);
        check(chord_build_hzb(m_frame, bMin, bMax, bValidRange, h.get()), "chord_build_hzb");
        return h;
    }
    void renderMesh(const CmdListRef& list) { check(chord_render_mesh(m_frame, list.get()), "chord_render_mesh"); }
    // renderMeshDepth (PASS_TYPE_DEPTH: two-sided, optional depth clamp) into this frame used as a cascade depth target
    void renderMeshDepth(const CmdListRef& list, uint32_t passFlags = CHORD_PASS_DEPTH | CHORD_PASS_DEPTH_CLAMP)
    { check(chord_render_mesh_depth(m_frame, list.get(), passFlags), "chord_render_mesh_depth"); }
    // detail::hzbCullingGeneric against the HZB of another view (previous cascade / cached cascade)
    CmdListRef hzbCullingGeneric(const HzbRef& hzb, const chord_InstanceCullingViewInfo& hzbView, float extentScale,
                                 bool bObjectUseLastFrameProject, const CmdListRef& in)
    {
        auto out = makeCmdList(m_ctx, chord_cmdlist_capacity(in.get()));
        check(chord_hzb_culling_generic(m_frame, hzb.get(), &hzbView, extentScale, bObjectUseLastFrameProject ? 1 : 0, in.get(), out.get()), "chord_hzb_culling_generic");
        return out;
    }
    void visibilityMark(const CmdListRef& post) { check(chord_visibility_mark(m_frame, post.get()), "chord_visibility_mark"); }
    void lighting(const CmdListRef& post) { check(chord_lighting(m_frame, post.get()), "chord_lighting"); }

private:
    chord_ctx* m_ctx;                  // kept alive by m_frame's reference (chord_ctx_destroy only drops the owner's)
    chord_frame* m_frame = nullptr;
    uint32_t m_w, m_h, m_lod0 = 0;
    CmdListRef m_keep[3];
};

// Record once / replay many (the reference records a command buffer per frame): construct, issue the frame's calls,
// call end(); launch() replays them with one CUDA-graph launch.
class Graph {
public:
    explicit Graph(const Context& ctx) : m_ctx(ctx.handle()) { check(chord_graph_begin_capture(m_ctx), "chord_graph_begin_capture"); }
    void end() { check(chord_graph_end_capture(m_ctx, &m_graph), "chord_graph_end_capture"); }
    void launch() { check(chord_graph_launch(m_graph), "chord_graph_launch"); }
    uint32_t kernelNodes() const { return chord_graph_kernel_nodes(m_graph); }
    ~Graph() { chord_graph_destroy(m_graph); }
    Graph(const Graph&) = delete;
    Graph& operator=(const Graph&) = delete;
private:
    chord_ctx* m_ctx;
    chord_graph* m_graph = nullptr;
};

inline std::unique_ptr<Frame> Context::createFrame(uint32_t w, uint32_t h) const { return std::make_unique<Frame>(*this, w, h); }

// chord_renderer: DeferredRenderer::render (renderer.cpp:311-363 + :489) as ONE native call per frame - recorded as CUDA graphs
// and replayed with rewritten uniforms (CHORD_RENDERER_GRAPH), with GPU timers at the reference's cut points
// (CHORD_RENDERER_TIMERS), and, as one rank of a tile-parallel frame, with the fused exchange of the history HZB over peer
// memory (exchangeExport / exchangeConnect) or the NCCL all-gather (chord_renderer_exchange_connect_nccl on handle()).
class Renderer {
public:
    Renderer(const Context& ctx, uint32_t w, uint32_t h, uint32_t flags) : m_w(w), m_h(h)
    { check(chord_renderer_create(ctx.handle(), w, h, flags, &m_r), "chord_renderer_create"); }
    ~Renderer() { chord_renderer_destroy(m_r); }
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;
    chord_renderer* handle() const { return m_r; }
    void setTile(const chord_rect* tile) { check(chord_renderer_set_tile(m_r, tile), "chord_renderer_set_tile"); }
    void render(const chord_PerframeCameraView& view, const chord_InstanceCullingViewInfo& iview, const chord_GPUObjectGLTFPrimitive* objects,
                uint32_t objectCount, uint32_t lod0MeshletCount, uint32_t meshletGroupCount)
    { check(chord_renderer_render(m_r, &view, &iview, objects, objectCount, lod0MeshletCount, meshletGroupCount), "chord_renderer_render"); }
    void clearHistory() { check(chord_renderer_clear_history(m_r), "chord_renderer_clear_history"); }
    chord_frame* frame() const { return chord_renderer_frame(m_r); }
    chord_cmdlist* postCull() const { return chord_renderer_post_cull(m_r); }
    chord_hzb* historyHzb() const { return chord_renderer_history_hzb(m_r); }
    std::vector<float> passTimes() const { std::vector<float> us(CHORD_PASS_COUNT); check(chord_renderer_pass_times(m_r, us.data()), "chord_renderer_pass_times"); return us; }
    // multi-GPU, fused exchange: export this rank's block, gather the handles / pointers of all ranks (any transport), connect
    void* exchangeExport(chord_ipc_handle* outHandle = nullptr)
    { void* p = nullptr; check(chord_renderer_exchange_export(m_r, outHandle, &p), "chord_renderer_exchange_export"); return p; }
    void exchangeConnect(uint32_t rank, uint32_t world, const chord_rect* tiles, const chord_ipc_handle* handles, void* const* localPtrs)
    { check(chord_renderer_exchange_connect(m_r, rank, world, tiles, handles, localPtrs), "chord_renderer_exchange_connect"); }
    void historySync() { check(chord_renderer_history_sync(m_r), "chord_renderer_history_sync"); }
    void setExchangeTimeout(uint32_t maxPolls) { check(chord_renderer_set_exchange_timeout(m_r, maxPolls), "chord_renderer_set_exchange_timeout"); }
    bool exchangeTimedOut() const { uint32_t v = 0; check(chord_renderer_exchange_status(m_r, &v), "chord_renderer_exchange_status"); return v != 0; }
private:
    chord_renderer* m_r = nullptr;
    uint32_t m_w, m_h;
};

// DeferredRenderer::render between "Clear GBuffers" and "lightingTile" (source/renderer/renderer.cpp:311-363) plus the
// history hand-off (:489, m_rendererHistory.hzbCtx): instanceCulling -> gltfVisibilityRenderingStage0 against LAST frame's HZB
// -> buildHZB(min) -> gltfVisibilityRenderingStage1 -> buildHZB(min, max, range), kept as history -> visibilityMark -> lighting.
// Headless: the frame stays in device memory (Frame), read it back with the chord_readback_* calls.
class DeferredRenderer {
public:
    DeferredRenderer(const Context& ctx, uint32_t width, uint32_t height, bool doResolve = false, bool useBvh = false)
        : m_ctx(ctx.handle()), m_frame(ctx.createFrame(width, height)), m_doResolve(doResolve), m_useBvh(useBvh) {}

    Frame& frame() { return *m_frame; }
    const HzbRef& historyHzb() const { return m_history; }
    const CmdListRef& postCull() const { return m_post; }
    // renderer.cpp:95-105 (resize / camera cut): the next frame culls without occlusion history
    void clearHistory() { m_history.reset(); }

    // one frame; returns the post-cull list (the drawMeshletCmdBuffer the resolve passes index).  Lists and HZBs come from a
    // small pool (getContext().getBufferPool() / getTexturePool() stand-in): everything is stream-ordered, so an object can
    // be handed out again as soon as the call that last used it has been enqueued - no per-frame allocation, no host sync.
    CmdListRef render(const chord_PerframeCameraView& view, const chord_InstanceCullingViewInfo& iview,
                      const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount,
                      uint32_t lod0MeshletCount, uint32_t meshletGroupCount, const chord_rect* tile = nullptr)
    {
        chord_frame* f = m_frame->handle();
        m_frame->begin(view, iview, objects, objectCount, lod0MeshletCount, meshletGroupCount, tile);   // uniforms + addClearGbufferPass
        releaseList(m_post);
        CmdListRef post = acquireList(lod0MeshletCount);
        if (m_useBvh) check(chord_instance_culling_bvh(f, post.get()), "chord_instance_culling_bvh");
        else          check(chord_instance_culling(f, post.get()), "chord_instance_culling");
        // gltfVisibilityRenderingStage0 (mesh_raster.cpp:269-312)
        int needStage1 = 0;
        CmdListRef visible, rejected;
        const bool haveHistory = m_history && chord_hzb_is_valid(m_history.get());
        if (haveHistory) { visible = acquireList(lod0MeshletCount); rejected = acquireList(lod0MeshletCount); }
        check(chord_visibility_stage0(f, m_history.get(), post.get(), visible.get(), rejected.get(), &needStage1), "chord_visibility_stage0");
        releaseList(visible);
        if (needStage1) {                                                                               // bShouldStage1GLTF
            HzbRef tmp = acquireHzb();
            check(chord_build_hzb(f, 1, 0, 0, tmp.get()), "chord_build_hzb");
            CmdListRef visible1 = acquireList(lod0MeshletCount);
            check(chord_visibility_stage1(f, tmp.get(), rejected.get(), visible1.get()), "chord_visibility_stage1");
            releaseList(visible1);
            m_freeHzb.push_back(tmp);
        }
        releaseList(rejected);
        HzbRef hzb = acquireHzb();
        check(chord_build_hzb(f, 1, 1, 1, hzb.get()), "chord_build_hzb");
        if (m_doResolve) { m_frame->visibilityMark(post); m_frame->lighting(post); }
        if (m_history) m_freeHzb.push_back(m_history);
        m_history = hzb;                                                                                // m_rendererHistory.hzbCtx = hzbCtx
        m_post = post;
        return post;
    }

private:
    CmdListRef acquireList(uint32_t capacity)
    {
        for (size_t i = 0; i < m_freeLists.size(); i++)
            if (chord_cmdlist_capacity(m_freeLists[i].get()) == capacity) { CmdListRef l = m_freeLists[i]; m_freeLists.erase(m_freeLists.begin() + (long)i); return l; }
        return makeCmdList(m_ctx, capacity);
    }
    void releaseList(CmdListRef& l) { if (l) { m_freeLists.push_back(l); l.reset(); } }
    HzbRef acquireHzb()
    {
        if (!m_freeHzb.empty()) { HzbRef h = m_freeHzb.back(); m_freeHzb.pop_back(); return h; }
        return makeHzb(m_ctx, m_frame->width(), m_frame->height());
    }

    chord_ctx* m_ctx;                  // kept alive by m_frame
    std::unique_ptr<Frame> m_frame;
    HzbRef m_history;
    CmdListRef m_post;
    std::vector<CmdListRef> m_freeLists;
    std::vector<HzbRef> m_freeHzb;
    bool m_doResolve, m_useBvh;
};

}  // namespace chord_b200
