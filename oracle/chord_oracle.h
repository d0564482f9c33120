/*
 * chord_oracle.h — CPU restatement of chord's visibility-buffer path.  TEST INFRASTRUCTURE ONLY:
 * nothing under chord_b200/ (the product) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, tests or fixtures for this path
 *  and cannot be built or run here (Win32/MSVC/DXC/Vulkan only), so this
 * oracle is pinned only by (a) being a literal restatement of the cited HLSL, (b) the Vulkan
 * fixed-function rules written out in DESIGN.md §3, and (c) the independent exact-arithmetic
 * cross-checks in tests/test_oracle_*.py.
 */
#ifndef CHORD_ORACLE_H
#define CHORD_ORACLE_H

#include "../include/chord_b200.h"   /* public POD types only (scene desc, structs) */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const chord_scene_desc* scene;
    const chord_PerframeCameraView* view;
    const chord_InstanceCullingViewInfo* iview;
    const chord_GPUObjectGLTFPrimitive* objects;
    uint32_t objectCount;
    uint32_t width, height;
    chord_rect tile;            /* scissor; full frame = {0,0,w,h} */
    uint32_t switchFlags;
    int threads;                /* OpenMP threads (<=0: default) */
    /* material texture table of the masked bucket (chord_textures_upload); may be NULL / 0 = everything builtin white */
    const chord_texture_desc* textures; uint32_t textureCount;
    const chord_sampler_desc* samplers; uint32_t samplerCount;
    uint32_t passFlags;         /* CHORD_PASS_DEPTH | CHORD_PASS_DEPTH_CLAMP for oracle_render_mesh (renderMeshDepth) */
} oracle_frame;

/* K1+K2: instance_culling.hlsl:47-208 in canonical (object, group, meshlet) order. returns count. */
uint32_t oracle_instance_culling(const oracle_frame* f, chord_draw_cmd* out, uint32_t capacity);

/* N4: the same list through the cluster-group BVH (nanite_builder.cpp:313-416 layout). */
uint32_t oracle_bvh_candidates(const oracle_frame* f, uint32_t* bits, uint32_t bitWords, uint32_t* totalGroups);
uint32_t oracle_instance_culling_bvh(const oracle_frame* f, chord_draw_cmd* out, uint32_t capacity, uint32_t* candidateGroups);

/* K3: hzb_mainview_culling.hlsl:35-213.  hzbMin = mip chain (fp16 bits), mip m at hzbMipOffsets[m] (in texels),
 * row pitch = mip0Width >> m. */
void oracle_hzb_culling(const oracle_frame* f, int firstStage,
                        const uint16_t* hzbMin, const uint64_t* hzbMipOffsets, uint32_t mip0Width, uint32_t mip0Height, uint32_t mipCount,
                        const chord_draw_cmd* in, uint32_t inCount,
                        chord_draw_cmd* outVisible, uint32_t* outVisibleCount,
                        chord_draw_cmd* outRejected, uint32_t* outRejectedCount);

/* hzbCullingGeneric: hzb_culling_generic.hlsl:37-172.  hzbView = view the HZB was built for. */
void oracle_hzb_culling_generic(const oracle_frame* f, const chord_InstanceCullingViewInfo* hzbView, float extentScale,
                                int bObjectUseLastFrameProject,
                                const uint16_t* hzbMin, const uint64_t* hzbMipOffsets, uint32_t mip0Width,
                                const chord_draw_cmd* in, uint32_t inCount, chord_draw_cmd* outVisible, uint32_t* outVisibleCount);

/* K4+K6+K6'+K7: mesh_raster.hlsl:51-210 + Vulkan fixed-function rules into a packed 64-bit surface (W*H). */
void oracle_render_mesh(const oracle_frame* f, const chord_draw_cmd* cmds, uint32_t count, uint64_t* vis64,
                        chord_raster_stats* stats);

/* Rasterise one clip-space triangle (exposed for the exact-arithmetic cross-check tests). */
void oracle_raster_clip_triangle(uint32_t width, uint32_t height, const chord_rect* tile,
                                 const float clip[12], uint32_t visId, int twoSided, uint64_t* vis64);

/* The specified texture sampler of the masked bucket, exposed for the sampler unit tests: alpha of
 * `tex` at (u, v) with screen-space derivatives (dudx, dvdx, dudy, dvdy) in uv units. */
float oracle_sample_alpha(const chord_texture_desc* tex, const chord_sampler_desc* smp, float u, float v,
                          float dudx, float dvdx, float dudy, float dvdy);

/* K8: hzb.hlsl / hzb_one.hlsl.  Chains receive fp16 bit patterns; offsets as above. */
void oracle_hzb_dims(uint32_t depthW, uint32_t depthH, uint32_t* mip0W, uint32_t* mip0H, uint32_t* mipCount);
void oracle_build_hzb(const uint64_t* vis64, uint32_t width, uint32_t height,
                      int bMin, int bMax, int bValidRange,
                      uint16_t* minChain, uint16_t* maxChain, const uint64_t* mipOffsets,
                      uint32_t mip0W, uint32_t mip0H, uint32_t mipCount, uint32_t validRangeBits[2]);

/* K9: visibility_tile.hlsl:63-131 — one uint4 per 8x8 tile. */
void oracle_visibility_mark(const oracle_frame* f, const uint64_t* vis64,
                            const chord_draw_cmd* postCull, uint32_t postCullCount, uint32_t* markerUint4);

/* K11: lighting.hlsl:270-381, PBR branch, untextured materials.  planes[6] = CHORD_GBUF_* order, 4 floats per pixel,
 * pre-quantisation fp32 values; untouched pixels are 0.  LUTs are RGBA16F bit patterns (256x64 and 64x16). */
/* sky branch of the lighting pass (lighting.hlsl:75-166) for the texels without geometry: scattering = RGBA16F 256 x 128 x 32
 * (atmosphere.hlsli:43-49), singleMie = same shape or NULL when atmosphere.bCombineScattering != 0 */
void oracle_sky(const oracle_frame* f, const uint64_t* vis64, const uint16_t* trLut, const uint16_t* irLut, const uint16_t* scattering,
                const uint16_t* singleMie, float* colorPlane);
void oracle_lighting(const oracle_frame* f, const uint64_t* vis64, const chord_draw_cmd* postCull, uint32_t postCullCount,
                     const uint16_t* transmittanceLut, const uint16_t* irradianceLut, float* const planes[6]);

/* decode through the post-instance-cull list: (object, meshlet, tri) or 0xFFFFFFFF x3 */
void oracle_decode(const uint64_t* vis64, uint32_t pixelCount, const chord_draw_cmd* postCull, uint32_t postCullCount,
                   uint32_t* triples);

uint16_t oracle_f32_to_f16_rtz(float v);
float    oracle_f16_to_f32(uint16_t h);

#ifdef __cplusplus
}
#endif
#endif
