/*
 * chord_b200.h — C ABI of libchord_b200.so: chord's Nanite-style visibility-buffer
 * path (cluster cull -> software raster -> HZB -> tile classify -> material resolve)
 * as hand-written sm_100a CUDA kernels, headless.
 *
 * This is the drop-in boundary for the reference's renderer pass graph.  Each entry
 * point cites the reference function it replaces.  Mapping of the reference's
 * vocabulary onto the ABI:
 *     graphics::GraphicsQueue&       -> the context's CUDA stream (calls are stream-ordered, asynchronous)
 *     bindless buffer / texture ids  -> device pointers resolved at chord_scene_upload
 *     PoolBufferGPUOnlyRef pairs     -> chord_cmdlist  (CountAndCmdBuffer, postprocessing.h:48)
 *     HZBContext                     -> chord_hzb      (render_helper.h:411-451)
 *     GBufferTextures                -> owned by chord_frame (render_textures.h:10-61)
 *     check(...) crash               -> non-zero return + chord_last_error()
 * Every function returns 0 on success, never throws, never falls back to a CPU path.
 * One host thread per context.
 */
#ifndef CHORD_B200_H
#define CHORD_B200_H

#include "chord_abi_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CHORD_B200_ABI_VERSION 1

typedef struct chord_ctx     chord_ctx;
typedef struct chord_frame   chord_frame;
typedef struct chord_cmdlist chord_cmdlist;
typedef struct chord_hzb     chord_hzb;
typedef struct { uint32_t x0, y0, x1, y1; } chord_rect;   /* pixel rect [x0,x1) x [y0,y1) */

enum {
    CHORD_OK = 0,
    CHORD_ERR_INVALID = 1,      /* bad argument (reference: check() failure) */
    CHORD_ERR_CUDA = 2,         /* CUDA runtime error, text in chord_last_error() */
    CHORD_ERR_CAPACITY = 3,     /* a list/queue capacity would be exceeded */
    CHORD_ERR_NO_DEVICE = 4
};

/* Thread-local text of the last failure on this thread. */
const char* chord_last_error(void);
int         chord_abi_version(void);
int         chord_device_count(void);     /* visible CUDA devices (0 when there is none / no driver) */

/* ---- context: replaces graphics::Context + one GraphicsQueue (source/graphics/graphics.h) */
int  chord_ctx_create(int device, chord_ctx** out);
/* Handle lifetime mirrors the reference's ref-counted pool resources (CommandBuffer::addReferenceResource,
 * source/graphics/command_list.h:23-38; PoolTextureRef / PoolBufferRef in render_helper.h:411-451): every object created
 * from a context (chord_cmdlist, chord_hzb, chord_frame, chord_graph, chord_cascade_views) holds a reference on it.
 * chord_ctx_destroy drops the OWNER's reference only; stream, scene and scratch stay alive - and the children stay fully
 * usable - until the last child has been destroyed, in any order. */
void chord_ctx_destroy(chord_ctx* ctx);
/* owner (1 until chord_ctx_destroy) + live children; for tests / leak checks */
uint32_t chord_ctx_refcount(const chord_ctx* ctx);
/* cudaStream_t of the context, for callers that interleave their own work / events. */
void* chord_ctx_stream(chord_ctx* ctx);
int  chord_ctx_sync(chord_ctx* ctx);
/* r.instanceculling.{frustumCulling,hzbCulling,meshletConeCulling} cvars folded into
 * switchFlags bits 0/1/2 (instance_culling.cpp:22-68, gltf.h:10-12). Default 0b111. */
int  chord_ctx_set_switch_flags(chord_ctx* ctx, uint32_t switchFlags);
uint32_t chord_ctx_get_switch_flags(const chord_ctx* ctx);
/* Tuning of the meshlet rasteriser's launch shape (no reference counterpart; nothing in the library reads environment
 * variables).  variant: 11 (default) = one warp per meshlet, 4 warps per CTA, 7 CTAs / SM; 13 = same with V <= 104 and 8 CTAs / SM;
 * 0 = one CTA per meshlet.  ctasPerSm: 0 = the variant's own.  warpMinMeshlets: lists shorter than this are rasterised by
 * the CTA-per-meshlet instantiation (default 16384).  A negative argument keeps the current value.  Every choice produces
 * the same image. */
int  chord_ctx_set_raster_tuning(chord_ctx* ctx, int variant, int ctasPerSm, int64_t warpMinMeshlets);
/* Number of kernels this context has launched since creation (for gpu_launches accounting). */
uint64_t chord_ctx_kernel_launches(const chord_ctx* ctx);

/* ---- record once / replay many: the reference records its passes into a command buffer every frame
 * (graphics::CommandList, command_list.h:23-38); here a fixed sequence of chord_* calls on one context (typically
 * chord_frame_begin .. chord_lighting of one frame) is captured into a CUDA graph and replayed with one launch, which
 * removes the per-launch gaps of a ~10..30 kernel frame.  Rules: run the sequence once eagerly first (lazy buffers get
 * sized), then chord_graph_begin_capture, the same calls, chord_graph_end_capture.  Calls that synchronise (read-backs,
 * chord_ctx_sync, uploads of scenes / textures / LUTs) are invalid while capturing.  The per-frame uniforms given to
 * chord_frame_begin during capture are snapshotted into pinned memory owned by the graph and re-uploaded by every
 * replay; work lists, HZBs and frames referenced by the captured calls must outlive the graph. */
typedef struct chord_graph chord_graph;
int  chord_graph_begin_capture(chord_ctx* ctx);
int  chord_graph_end_capture(chord_ctx* ctx, chord_graph** out);
int  chord_graph_launch(chord_graph* graph);                 /* stream-ordered, asynchronous */
uint32_t chord_graph_kernel_nodes(const chord_graph* graph); /* kernels one replay launches */
/* A recorded frame replayed with NEW per-frame uniforms (the reference re-records its command buffer every frame with the
 * frame's view; here only the uniform snapshot changes): rewrites the pinned block(s) the recorded chord_frame_begin call(s)
 * on `frame` upload, so that the next chord_graph_launch renders this camera / these object transforms.  What was baked into
 * kernel parameters at capture time must not change: objectCount, frame size and tile, cameraFovy, switch flags, which lists /
 * HZBs the calls used.  Host-side only (a memcpy into pinned memory); call it after the previous replay of this graph has
 * started executing its upload node - in practice: alternate two graphs (last frame's HZB ping-pong needs two anyway) or
 * keep at most one frame in flight. */
int  chord_graph_update_frame_inputs(chord_graph* graph, chord_frame* frame, const chord_PerframeCameraView* view,
                                     const chord_InstanceCullingViewInfo* instanceView,
                                     const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount);
void chord_graph_destroy(chord_graph* graph);

/* ---- scene: replaces gpu_scene.{h,cpp} pools + the asset GPU buffers
 * (gltf.h:65-153; packing asset_gltf_helper.cpp:522-548, :572-584).  Arrays are the
 * reference structs byte-for-byte.  One chord_primitive_datas_desc corresponds to one
 * GLTFPrimitiveDatasBuffer record (one imported asset). */
typedef struct {
    const chord_GPUGLTFMeshlet*      meshlets;       uint32_t meshletCount;
    const float*                     positions;      uint32_t vertexCount;      /* float3, tightly packed */
    const float*                     normals;        /* float3 x vertexCount, may be NULL if lighting is not run */
    const float*                     uv0;            /* float2 x vertexCount, may be NULL */
    const float*                     tangents;       /* float4 x vertexCount, may be NULL */
    const uint32_t*                  meshletData;    uint32_t meshletDataCount; /* [remap u32 x V][tri u32 x T] per meshlet */
    const chord_GPUBVHNode*          bvhNodes;       uint32_t bvhNodeCount;     /* may be NULL */
    const chord_GPUGLTFMeshletGroup* groups;         uint32_t groupCount;
    const uint32_t*                  groupIndices;   uint32_t groupIndicesCount;
} chord_primitive_datas_desc;

typedef struct {
    const chord_primitive_datas_desc* datas;      uint32_t datasCount;
    const chord_GLTFPrimitiveBuffer*  primitives; uint32_t primitiveCount;
    const chord_GLTFMaterialGPUData*  materials;  uint32_t materialCount;
} chord_scene_desc;

/* Copies everything to HBM (synchronous on return). The big arrays may be host or device pointers (copied with
 * cudaMemcpyDefault); the desc structs themselves, primitives and materials are host memory. Replaces any previous scene. */
int chord_scene_upload(chord_ctx* ctx, const chord_scene_desc* desc);

/* ---- material textures for the masked (alpha-tested) raster bucket: replaces the bindless Texture2D / SamplerState ids of
 * GLTFMaterialGPUData.baseColorId / baseColorSampler (gltf.h:118-153; mesh_raster.hlsl:107-112, :198-204) by an index
 * into the table uploaded here.  Texels are uncompressed RGBA8 unorm (asset_gltf_material.cpp:80-112 also knows BC
 * formats; decompress before upload), mip k is max(1, w >> k) x max(1, h >> k), mips tightly packed from 0 down.
 * The raster path reads a copy that keeps only the alpha channel (one byte per texel); the resolve (chord_lighting: base
 * colour, normal map, occlusion / roughness / metallic through SampleGrad, material.hlsli:30-154) reads the RGBA8 chains.
 * An id that is out of range, CHORD_UNVALID_ID or whose desc has rgba8 == NULL is the builtin white texture
 * (alpha 1; asset_gltf.cpp:295-326), which makes the alpha test a per-meshlet constant.
 * The sampler is specified in DESIGN.md section 3 (trilinear, LOD from forward-differenced perspective-correct uv). */
enum { CHORD_ADDRESS_REPEAT = 0, CHORD_ADDRESS_CLAMP_TO_EDGE = 1, CHORD_ADDRESS_MIRRORED_REPEAT = 2 };
/* flags: CHORD_TEXTURE_SRGB = the texel's r, g, b are sRGB-encoded and decoded to linear BEFORE filtering, as a *_SRGB image
 * format does (base colour / emissive images: asset_gltf_material.cpp:130-155, :510-511); alpha is always linear. */
enum { CHORD_TEXTURE_SRGB = 1u };
typedef struct { const uint8_t* rgba8; uint32_t width, height, mipCount, flags; } chord_texture_desc;
typedef struct { uint32_t addressU, addressV; } chord_sampler_desc;
int chord_textures_upload(chord_ctx* ctx, const chord_texture_desc* textures, uint32_t textureCount,
                          const chord_sampler_desc* samplers, uint32_t samplerCount);

/* ---- pooled work lists: CountAndCmdBuffer = {uint count, uint3 cmd[capacity]}
 * (instance_culling.cpp:140-141 sizes them by lod0MeshletCount). */
int  chord_cmdlist_create(chord_ctx* ctx, uint32_t capacity, chord_cmdlist** out);
void chord_cmdlist_destroy(chord_cmdlist* list);
int  chord_cmdlist_upload(chord_cmdlist* list, const chord_draw_cmd* hostCmds, uint32_t count);
/* stream-ordered upload without host synchronisation; hostCmds must stay valid (pinned for a truly async copy). */
int  chord_cmdlist_upload_async(chord_cmdlist* list, const chord_draw_cmd* hostCmds, uint32_t count);
int  chord_cmdlist_readback(chord_cmdlist* list, chord_draw_cmd* hostCmds, uint32_t hostCapacity, uint32_t* outCount);
uint32_t chord_cmdlist_capacity(const chord_cmdlist* list);
/* device pointers (count is a single u32, cmds is uint3[capacity]) */
void* chord_cmdlist_count_ptr(chord_cmdlist* list);
void* chord_cmdlist_cmds_ptr(chord_cmdlist* list);

/* ---- HZB: HZBContext (render_helper.h:411-451; hzb.cpp:49-74 sizing: nextPOT(dim)/2,
 * halved again if equal to dim; R16F; floor(log2(max(w,h)))+1 mips, >= 6). */
int  chord_hzb_create(chord_ctx* ctx, uint32_t depthWidth, uint32_t depthHeight, chord_hzb** out);
void chord_hzb_destroy(chord_hzb* hzb);
int  chord_hzb_info(const chord_hzb* hzb, uint32_t* mip0Width, uint32_t* mip0Height, uint32_t* mipCount,
                    uint32_t* validWidth, uint32_t* validHeight);
/* which: 0 = min chain, 1 = max chain.  Host buffer receives (w>>mip)*(h>>mip) fp16 bit patterns. */
int  chord_hzb_readback(chord_hzb* hzb, int which, uint32_t mip, uint16_t* host);
/* valid depth range {minBits, maxBits} as written by hzb.hlsl:165-172 (DIM_COMPUTE_VALID_RANGE). */
int  chord_hzb_readback_valid_range(chord_hzb* hzb, uint32_t outMinMaxBits[2]);
/* Whole min chain as one contiguous device allocation (for NCCL / peer copies), bytes. */
void*    chord_hzb_chain_ptr(chord_hzb* hzb, int which);
uint64_t chord_hzb_chain_bytes(const chord_hzb* hzb);
/* Byte offset of a mip inside the chain allocation. */
uint64_t chord_hzb_mip_offset(const chord_hzb* hzb, uint32_t mip);
/* ---- screen-tile partition across GPUs. A tile rect is cut on multiples of 64 px, so every 64x64 depth
 * block (one texel of mip 5) belongs to exactly one tile. chord_build_hzb on a tiled frame fills only the tile's part
 * of mips 0..5; the parts are exchanged with one all-gather of the packed regions (NCCL over NVLink, issued by the
 * host on chord_ctx_stream) and mips 6.. are then rebuilt locally by chord_hzb_finish_top. */
int chord_hzb_tile_bytes(const chord_hzb* hzb, const chord_rect* tile, uint64_t* bytes);
int chord_hzb_pack_tile(chord_hzb* hzb, int which, const chord_rect* tile, void* dstDevice);
int chord_hzb_unpack_tiles(chord_hzb* hzb, int which, const chord_rect* tiles, uint32_t n, const void* srcDevice, uint64_t slotBytes);
int chord_hzb_finish_top(chord_frame* frame, int bMin, int bMax, chord_hzb* dst);
/* Mark a chain filled from outside (after an all-gather) as valid history. */
int  chord_hzb_set_valid(chord_hzb* hzb, int valid);
int  chord_hzb_is_valid(const chord_hzb* hzb);

/* ---- frame: GBufferTextures + per-view uniforms (renderer.cpp:175-263, render_textures.cpp:20-102) */
int  chord_frame_create(chord_ctx* ctx, uint32_t width, uint32_t height, chord_frame** out);
void chord_frame_destroy(chord_frame* frame);

/* Uploads the per-frame uniforms (host pointers, pinned or pageable) and clears the
 * visibility/depth surface (addClearGbufferPass: vis = 0, depth = 0.0).
 * lod0MeshletCount / meshletGroupCount = PerframeCollected totals (component_gltf_mesh.cpp:58-122).
 * tile = NULL renders the full frame; otherwise raster, HZB and resolve are restricted to the
 * rect (screen-tile partition across GPUs). */
int chord_frame_begin(chord_frame* frame,
                      const chord_PerframeCameraView* view,
                      const chord_InstanceCullingViewInfo* instanceView,
                      const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount,
                      uint32_t lod0MeshletCount, uint32_t meshletGroupCount,
                      const chord_rect* tile);

/* instanceCulling (gltf_rendering.h:37; instance_culling.cpp:83-161): per-object frustum
 * cull, LOD-DAG cut, per-meshlet cone + frustum cull.  out receives (objectId, meshletId, slot)
 * in canonical order (object, group, meshlet-in-group); slot == index. */
int chord_instance_culling(chord_frame* frame, chord_cmdlist* out);

/* Same result through a persistent-warp traversal of the 8-ary cluster-group BVH (GPUBVHNode, gltf.h:16-24; built by
 * nanite_builder.cpp:313-416 and uploaded by the reference, but never walked there: instance_culling.hlsl:96-99). Nodes
 * whose parent-error bound projects below one pixel are skipped; the list is identical to chord_instance_culling. */
int chord_instance_culling_bvh(chord_frame* frame, chord_cmdlist* out);
/* candidate bit per (visible object, group) in canonical order + {head, tail, pending, done, overflow} of the traversal */
int chord_readback_bvh_candidates(chord_frame* frame, uint32_t* hostBits, uint32_t words, uint32_t* outCtrl5);

/* detail::hzbCulling (instance_culling.cpp:286-351; hzb_mainview_culling.hlsl:35-213).
 * phase 0: last frame's matrices + HZB, also writes the rejected list. phase 1: this frame's. */
int chord_hzb_culling(chord_frame* frame, const chord_hzb* hzb, int firstStage,
                      chord_cmdlist* in, chord_cmdlist* outVisible, chord_cmdlist* outRejected /* NULL in phase 1 */);

/* renderMesh (mesh_raster.cpp:208-253): filter by (alphaMode, twoSided) and rasterise the
 * list into the frame's visibility/depth surface (K4+K6+K6'+K7 fused). */
int chord_render_mesh(chord_frame* frame, chord_cmdlist* list);

/* ---- depth-only views (cascade shadow maps): renderMeshDepth (mesh_raster.cpp:159-206, PASS_TYPE_DEPTH)
 * through the same cull + raster kernels.  A depth view is a chord_frame of cascadeDim x cascadeDim begun with the
 * cascade's InstanceCullingViewInfo (cascade_setup.hlsl:89-371; frame size == instanceView.renderDimension); its depth is
 * the high word of the packed surface (chord_readback_depth32), chord_build_hzb builds the cascade HZB from it.
 *   CHORD_PASS_DEPTH        every material is rasterised two-sided (mesh_raster.cpp:187-189: cull NONE, no cull #0)
 *   CHORD_PASS_DEPTH_CLAMP  vkCmdSetDepthClampEnableEXT (mesh_raster.cpp:147-150): no near / far clipping, z clamped to
 *                           [0, 1]; valid for orthographic instance views (w == 1).
 * The reference also records vkCmdSetDepthBias(const, 0, slope) but never enables depth bias (helper.h:405 leaves
 * vkCmdSetDepthBiasEnable FALSE, defaults 0.0 render_helper.h:481-482), so no bias is applied on either side. */
enum { CHORD_PASS_DEPTH = 1u, CHORD_PASS_DEPTH_CLAMP = 2u };
int chord_render_mesh_depth(chord_frame* frame, chord_cmdlist* list, uint32_t passFlags);

/* detail::hzbCullingGeneric (instance_culling.cpp; hzb_culling_generic.hlsl:37-172): one-pass 2x2 HZB test of the list
 * against the HZB of ANOTHER view (the previous cascade, or the cached cascade of last frame) with the box rect widened
 * by extentScale pixels.  hzbView = that view's InstanceCullingViewInfo (host pointer).  Survivors keep input order. */
int chord_hzb_culling_generic(chord_frame* frame, const chord_hzb* hzb, const chord_InstanceCullingViewInfo* hzbView,
                              float extentScale, int bObjectUseLastFrameProject, chord_cmdlist* in, chord_cmdlist* outVisible);

/* Cascade views on the device: CascadeSetupCS (cascade_setup.hlsl:89-371; mesh_raster.cpp:400-431) writes one
 * InstanceCullingViewInfo per cascade into a device buffer, reading buildHZB's valid depth range on the GPU (SDSM) so
 * that the shadow pass needs no read-back.  chord_cascade_setup_desc = CascadeSetupPushConsts (cascade_setup.hlsl:24-50).
 * The main frame supplies the PerframeCameraView already on the device (it must have been begun this frame).
 * chord_frame_begin_cascade / chord_hzb_culling_generic_cascade are chord_frame_begin / chord_hzb_culling_generic with
 * the instance view taken from that buffer. */
#define CHORD_MAX_CASCADE_COUNT 8      /* base.h:88 kMaxCascadeCount */
typedef struct {
    float lightDir[3];
    float cascadeStartDistance, cascadeEndDistance, splitLambda, farCascadeSplitLambda, farCascadeEndDistance, radiusScale;
    uint32_t cascadeCount, realtimeCount, cascadeDim, bCacheValid, tickCount;
} chord_cascade_setup_desc;
typedef struct chord_cascade_views chord_cascade_views;
int  chord_cascade_views_create(chord_ctx* ctx, chord_cascade_views** out);
void chord_cascade_views_destroy(chord_cascade_views* views);
int  chord_cascade_setup(chord_frame* mainFrame, const chord_cascade_setup_desc* desc, const chord_hzb* validRangeSource /* NULL: none */,
                         chord_cascade_views* views);
int  chord_cascade_views_readback(chord_cascade_views* views, chord_InstanceCullingViewInfo* host, uint32_t count);
int  chord_cascade_views_upload(chord_cascade_views* views, const chord_InstanceCullingViewInfo* host, uint32_t count);
int  chord_frame_begin_cascade(chord_frame* depthFrame, const chord_PerframeCameraView* view, chord_cascade_views* views, uint32_t cascadeId,
                               const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount,
                               uint32_t lod0MeshletCount, uint32_t meshletGroupCount);
int  chord_hzb_culling_generic_cascade(chord_frame* frame, const chord_hzb* hzb, chord_cascade_views* views, uint32_t hzbViewId,
                                       float extentScale, int bObjectUseLastFrameProject, chord_cmdlist* in, chord_cmdlist* outVisible);

/* gltfVisibilityRenderingStage0 (mesh_raster.cpp:269-312).  needStage1 mirrors the bool result.
 * scratchVisible is the pooled list that receives the phase-0 survivors. */
int chord_visibility_stage0(chord_frame* frame, const chord_hzb* lastFrameHzb /* may be NULL / invalid */,
                            chord_cmdlist* in, chord_cmdlist* scratchVisible, chord_cmdlist* outRejected,
                            int* needStage1);
/* gltfVisibilityRenderingStage1 (mesh_raster.cpp:314-329). */
int chord_visibility_stage1(chord_frame* frame, const chord_hzb* thisFrameHzb,
                            chord_cmdlist* rejected, chord_cmdlist* scratchVisible);

/* buildHZB (postprocessing.h:41-54; hzb.cpp:38-227) from the frame's depth. */
int chord_build_hzb(chord_frame* frame, int bMin, int bMax, int bValidRange, chord_hzb* dst);

/* visibilityMark + prepareShadingTileParam (visibility_tile.cpp:20-111). */
int chord_visibility_mark(chord_frame* frame, chord_cmdlist* postInstanceCull);

/* lighting (lighting.h:37; lighting.cpp:193-246), PBR branch.  LUT layouts: Appendix B of SURVEY.md. */
typedef struct {
    const uint16_t* transmittance;   /* RGBA16F, 256 x 64, host pointer */
    const uint16_t* irradiance;      /* RGBA16F, 64 x 16, host pointer */
} chord_atmosphere_luts;
int chord_atmosphere_luts_upload(chord_ctx* ctx, const chord_atmosphere_luts* luts);
/* The sky branch of the same pass (LIGHTING_TYPE == None: lighting.hlsl:75-166 skyColorEvaluate, the {None, PBR} loop of
 * lighting.cpp:193-246): once the scattering volumes are uploaded, chord_lighting also writes the sky colour into every texel
 * of the tile rect that holds no geometry (colour plane only).  scattering / singleMie: RGBA16F, 256 x 128 x 32 (x = nu * 32 +
 * mu_s, y = mu, z = r; atmosphere.hlsli:43-49), host pointers.  singleMie is read only when atmosphere.bCombineScattering == 0
 * (lighting.hlsl:84) and may be NULL otherwise; NULL with bCombineScattering == 0 - a state the reference leaves undefined
 * (its push constant stays bindless id 0, lighting.cpp:230-233) - makes the scattering volume stand in.  scattering == NULL
 * removes the branch again (unshaded texels keep the addClearGbufferPass value). */
int chord_atmosphere_scattering_upload(chord_ctx* ctx, const uint16_t* scattering, const uint16_t* singleMie);
int chord_lighting(chord_frame* frame, chord_cmdlist* postInstanceCull);

/* ---- read-backs (synchronise the stream) */
int chord_readback_vis64(chord_frame* frame, uint64_t* host);          /* W*H packed texels */
int chord_readback_visibility32(chord_frame* frame, uint32_t* host);   /* the reference's R32_UINT image */
int chord_readback_depth32(chord_frame* frame, float* host);
/* pipelined variant: extraction on the context stream, PCIe copy on a second stream (two staging slots), so the next
 * frame overlaps the transfer; hostPinned must be page-locked. chord_frame_wait_readbacks completes all pending copies. */
int chord_readback_visibility32_async(chord_frame* frame, uint32_t* hostPinned);
int chord_frame_wait_readbacks(chord_frame* frame);
/* Stream of the pipelined read-backs: 0 (default) = the context's one transfer stream, shared with chord_cmdlist_upload_async
 * (one PCIe transfer at a time); 1 = a stream of the frame's own (uploads and read-backs overlap each other too). */
int chord_frame_set_readback_stream(chord_frame* frame, int ownStream);
/* decoded (objectId, meshletId, triangleId) per pixel through a post-instance-cull list
 * (visibility_tile.hlsl:47-60); 0xFFFFFFFF triple where no geometry. */
int chord_readback_decoded(chord_frame* frame, chord_cmdlist* postInstanceCull, uint32_t* hostTriples);

/* Self test of the raster's per-vertex arithmetic (no reference counterpart): n pseudo-random clip-space vertices (arbitrary
 * bit patterns, extreme exponents, exact pixel boundaries, frame-like distributions) go through the shared-reciprocal vertex
 * record k_raster_fast uses and through the plain statement of mesh_raster.hlsl:84-119 + the viewport transform (three IEEE
 * divisions); hostOut3 = { records that differ, index of the first, records that were inside the guard band }. */
int chord_selftest_vertex_records(chord_ctx* ctx, uint32_t n, uint32_t seed, uint32_t width, uint32_t height, uint32_t depthClamp, uint64_t* hostOut3);
int chord_readback_tile_marker(chord_frame* frame, uint32_t* hostUint4PerTile);
enum { CHORD_GBUF_COLOR = 0, CHORD_GBUF_BASECOLOR, CHORD_GBUF_VERTEX_NORMAL, CHORD_GBUF_PIXEL_NORMAL,
       CHORD_GBUF_MOTION, CHORD_GBUF_AO_ROUGH_METAL, CHORD_GBUF_COUNT };
/* pre-quantisation fp32 planes, 4 floats per pixel */
int chord_readback_gbuffer_f32(chord_frame* frame, int which, float* host);
/* packed = 1: chord_lighting writes the reference's storage formats (render_textures.h:21-59, lighting.hlsl:62-73):
 * colour RGBA16F (8 B/px), baseColor / vertexNormal / pixelNormal A2B10G10R10_UNORM, motion RG16F, aoRoughMetal RGBA8
 * (28 B/px); packed = 0 (default) writes the pre-quantisation fp32 planes used for the 1e-5 parity check. */
int chord_frame_set_gbuffer_packed(chord_frame* frame, int packed);
int chord_readback_gbuffer_packed(chord_frame* frame, int which, void* host);
void* chord_frame_vis64_ptr(chord_frame* frame);
int   chord_frame_dims(const chord_frame* frame, uint32_t* w, uint32_t* h);

/* ---- the whole frame as ONE call: DeferredRenderer::render between "Clear GBuffers" and "lightingTile"
 * (source/renderer/renderer.cpp:311-363) + the history hand-off (:489, m_rendererHistory.hzbCtx):
 *   uniforms + addClearGbufferPass -> instanceCulling -> gltfVisibilityRenderingStage0 against LAST frame's HZB -> buildHZB(min)
 *   -> gltfVisibilityRenderingStage1 -> buildHZB(min, max, range), kept as history -> visibilityMark -> lighting.
 * The renderer owns the frame, the pooled lists and HZBs (getBufferPool() / getTexturePool() stand-in).  Flags:
 *   CHORD_RENDERER_RESOLVE         run visibilityMark + lighting (needs chord_atmosphere_luts_upload and vertex normals)
 *   CHORD_RENDERER_PACKED_GBUFFER  lighting writes the reference's storage formats (28 B / px)
 *   CHORD_RENDERER_BVH             instance culling through the cluster-group BVH traversal
 *   CHORD_RENDERER_GRAPH           record the sequence once per history slot (the reference records its command buffer every
 *                                  frame, command_list.h:23-38) and replay it with the frame's uniforms rewritten: one launch
 *                                  per frame.  The first two frames (and the frame after any change of object count, totals,
 *                                  tile, switch flags or cameraFovy) run eagerly.
 *   CHORD_RENDERER_TIMERS          GPU timers at the reference's seven cut points (renderer.cpp:322-363; the timer lambda of
 *                                  GLTFRenderContext, gltf_rendering.h:15-35), read with chord_renderer_pass_times
 *   CHORD_RENDERER_ASYNC_RESOLVE   visibilityMark + lighting of frame N run on a second, lower-priority stream while the context
 *                                  stream already runs the visibility passes of frame N + 1 (two frames in flight, one per history
 *                                  slot; the reference submits everything to one graphics queue).  The latency-bound cull / raster
 *                                  / HZB kernels leave most issue slots of a B200 idle, the ALU-bound resolve fills them.  Results
 *                                  are identical; chord_renderer_frame orders the context stream after the resolve of the frame it
 *                                  returns.  Needs CHORD_RENDERER_RESOLVE; ignored with CHORD_RENDERER_TIMERS (timers serialise). */
typedef struct chord_renderer chord_renderer;
enum { CHORD_RENDERER_RESOLVE = 1u, CHORD_RENDERER_PACKED_GBUFFER = 2u, CHORD_RENDERER_BVH = 4u, CHORD_RENDERER_GRAPH = 8u, CHORD_RENDERER_TIMERS = 16u,
       CHORD_RENDERER_ASYNC_RESOLVE = 32u };
int  chord_renderer_create(chord_ctx* ctx, uint32_t width, uint32_t height, uint32_t flags, chord_renderer** out);
void chord_renderer_destroy(chord_renderer* renderer);
/* NULL = the full frame; otherwise cull (through the instance view the caller passes), raster, HZB and resolve are restricted to the rect */
int  chord_renderer_set_tile(chord_renderer* renderer, const chord_rect* tile);
int  chord_renderer_render(chord_renderer* renderer, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* instanceView,
                           const chord_GPUObjectGLTFPrimitive* objects, uint32_t objectCount, uint32_t lod0MeshletCount, uint32_t meshletGroupCount);
int  chord_renderer_clear_history(chord_renderer* renderer);       /* renderer.cpp:95-105 (resize / camera cut) */
chord_frame*   chord_renderer_frame(chord_renderer* renderer);
chord_cmdlist* chord_renderer_post_cull(chord_renderer* renderer);   /* drawMeshletCmdBuffer of the last frame */
chord_hzb*     chord_renderer_history_hzb(chord_renderer* renderer); /* m_rendererHistory.hzbCtx */
uint64_t       chord_renderer_graph_replays(const chord_renderer* renderer);
enum { CHORD_PASS_INSTANCE_CULLING = 0, CHORD_PASS_VISIBILITY_STAGE0, CHORD_PASS_BUILD_HZB_STAGE0, CHORD_PASS_VISIBILITY_STAGE1,
       CHORD_PASS_BUILD_HZB, CHORD_PASS_VISIBILITY_TILE_MARKER, CHORD_PASS_LIGHTING_TILE, CHORD_PASS_COUNT };
int  chord_renderer_pass_times(chord_renderer* renderer, float outMicroseconds[CHORD_PASS_COUNT]);

/* ---- multi-GPU: one rank per GPU, one screen tile per rank (cut on multiples of 64 px), scene replicated.
 * Phase 0 of the two-phase occlusion cull needs last frame's WHOLE min-HZB on every rank; the reference is single-GPU.
 *
 * Fused exchange (default): chord_renderer_exchange_export allocates the rank's exchange block (two history HZBs + flags) and
 * returns its CUDA IPC handle (ranks in different processes) and / or its device pointer (ranks in one process); the host
 * gathers the handles / pointers of all ranks by any means and hands them to chord_renderer_exchange_connect.  From then on
 * the final buildHZB of every frame stores this rank's part of mips 0..5 (min + max) straight into every rank's chain over
 * NVLink, publishes its valid range and an arrival flag; the next frame begins with a one-block kernel that waits for all
 * flags (bounded: chord_renderer_exchange_status reports a peer that never arrived) and rebuilds mips 6..: no pack /
 * all-gather / unpack, no host synchronisation, recordable (CHORD_RENDERER_GRAPH).  All ranks must render the same frames.
 * chord_renderer_history_sync completes last frame's chain now (tests / read-backs of chord_renderer_history_hzb).
 *
 * NCCL exchange: chord_nccl_* wrap a communicator (libnccl.so.2 is resolved with dlopen at the first call, there is no
 * link-time dependency); chord_hzb_allgather = pack -> ncclAllGather -> unpack -> rebuild of mips 6.. on the context stream,
 * for callers that sequence the passes themselves; chord_renderer_exchange_connect_nccl makes the renderer use it. */
#define CHORD_MAX_RANKS 16
typedef struct { unsigned char bytes[64]; } chord_ipc_handle;
int  chord_renderer_exchange_export(chord_renderer* renderer, chord_ipc_handle* outHandle /* may be NULL */, void** outLocalPtr /* may be NULL */);
int  chord_renderer_exchange_connect(chord_renderer* renderer, uint32_t rank, uint32_t world, const chord_rect* tiles /* [world] */,
                                     const chord_ipc_handle* handles /* [world] or NULL */, void* const* localPtrs /* [world] or NULL */);
int  chord_renderer_history_sync(chord_renderer* renderer);
int  chord_renderer_exchange_status(chord_renderer* renderer, uint32_t* outTimedOut);
/* Failure detection of the fused exchange: how long a rank waits for its peers' arrival flags before it gives up, raises the
 * status word above and renders on with the history it has (a peer that died must not hang the others).  maxPolls x 200 ns;
 * 0 = the default, 20 000 000 polls = ~4 s. */
int  chord_renderer_set_exchange_timeout(chord_renderer* renderer, uint32_t maxPolls);
typedef struct chord_nccl_comm chord_nccl_comm;
int  chord_nccl_unique_id(unsigned char outId[128]);                 /* rank 0; distribute the 128 bytes to every rank */
int  chord_nccl_comm_create(chord_ctx* ctx, uint32_t world, uint32_t rank, const unsigned char id[128], chord_nccl_comm** out);
void chord_nccl_comm_destroy(chord_nccl_comm* comm);
/* staging: device memory of at least (world + 1) x the largest chord_hzb_tile_bytes (each rounded up to 256 bytes) */
int  chord_hzb_allgather(chord_frame* frame, chord_hzb* hzb, const chord_rect* tiles /* [world] */, chord_nccl_comm* comm,
                         int bMin, int bMax, void* staging, uint64_t stagingBytes);
int  chord_renderer_exchange_connect_nccl(chord_renderer* renderer, const chord_rect* tiles /* [world] */, chord_nccl_comm* comm);

/* ---- statistics written by the raster kernels (device counters, read back on demand) */
typedef struct {
    uint64_t trianglesSubmitted;   /* triangles in rasterised meshlets */
    uint64_t trianglesCulled;      /* removed by the four mesh-shader culls */
    uint64_t trianglesClipped;     /* took the clip (near / guard band) path */
    uint64_t trianglesLarge;       /* routed through the tile queue */
    uint64_t largeQueueOverflow;   /* handled inline because the queue was full */
    uint64_t meshletsRastered;
    uint64_t fragmentsAlphaKilled; /* covered samples discarded by the masked bucket's clip() (mesh_raster.hlsl:203) */
    uint64_t pad0;
} chord_raster_stats;
int chord_frame_raster_stats(chord_frame* frame, chord_raster_stats* out);
/* Shrink the logical capacities of the raster work queues (0 = keep). Overflow never changes the image: deferred
 * triangles that do not fit make their meshlet go through the all-in-one kernel, large triangles are rasterised in place. */
int chord_frame_set_queue_limits(chord_frame* frame, uint32_t deferredTriangles, uint32_t largeTriangles, uint32_t largeItems);
int chord_frame_reset_stats(chord_frame* frame);

#ifdef __cplusplus
}
#endif
#endif /* CHORD_B200_H */
