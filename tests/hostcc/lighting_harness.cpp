// TEST INFRASTRUCTURE ONLY: executes the source text of the per-thread resolve kernels (k_resolve_records, k_lighting<>) of
// chord_b200/csrc/resolve.cu on the CPU, one CUDA thread after the other, through the stand-in headers of this directory.
// resolve_device.inc is resolve.cu with its launch_* functions cut out (tests/test_kernel_source_on_host.py writes it);
// the host-side constant fill (fill_lighting_consts) is the library's own function.  See cuda_runtime.h here for what a
// host run does and does not prove.
#include "common.cuh"

#include <cmath>
#include <vector>

namespace chordb { void set_error(const char*, ...) {} }

#include "resolve_device.inc"

using namespace chordb;

template <typename K>
static void run_grid(unsigned grid, unsigned block, K&& body) { hostcc::launch(dim3(grid), dim3(block), body); }

extern "C" int hostcc_lighting(const chord_scene_desc* desc, const chord_PerframeCameraView* view, const chord_InstanceCullingViewInfo* iview,
                               const chord_GPUObjectGLTFPrimitive* objects, uint32_t nObjects, uint32_t W, uint32_t H,
                               unsigned long long* vis64, const chord_draw_cmd* post, uint32_t nPost,
                               const void* trLut, const void* irLut, int packed, void** planes)
{
    std::vector<DatasDev> datas(desc->datasCount);
    for (uint32_t i = 0; i < desc->datasCount; i++) {
        const chord_primitive_datas_desc& d = desc->datas[i];
        datas[i] = DatasDev{d.meshlets, d.positions, d.normals, d.uv0, d.tangents, d.meshletData, d.groups, d.groupIndices, d.bvhNodes};
    }
    // k_object_prepare's arithmetic (cull.cu): derived = view matrix * object matrix, column by column
    std::vector<ObjectDerived> derived(nObjects);
    for (uint32_t o = 0; o < nObjects; o++) {
        const auto& b = objects[o].basicData;
        mat_mul(iview->translatedWorldToClip.m, b.localToTranslatedWorld.m, derived[o].localToClip.m);
        mat_mul(view->translatedWorldToView.m, b.localToTranslatedWorld.m, derived[o].localToView.m);
        mat_mul(view->translatedWorldToClipLastFrame.m, b.localToTranslatedWorldLastFrame.m, derived[o].localToClipHzb0.m);
        mat_mul(view->translatedWorldToClip.m, b.localToTranslatedWorld.m, derived[o].localToClipHzb1.m);
        mat_mul(view->translatedWorldToClip_NoJitter.m, b.localToTranslatedWorld.m, derived[o].localToClipNoJitter.m);
        mat_mul(view->translatedWorldToClipLastFrame_NoJitter.m, b.localToTranslatedWorldLastFrame.m, derived[o].localToClipLastNoJitter.m);
    }
    FrameDev f{};
    f.scene.datas = datas.data(); f.scene.prims = desc->primitives; f.scene.mats = desc->materials;
    f.scene.nDatas = desc->datasCount; f.scene.nPrims = desc->primitiveCount; f.scene.nMats = desc->materialCount;
    f.view = view; f.iview = iview; f.objects = objects; f.derived = derived.data(); f.objectCount = nObjects;
    f.width = W; f.height = H; f.tile = chord_rect{0, 0, W, H}; f.switchFlags = 7u; f.vis64 = vis64;

    std::vector<ResolveRecord> recs(nPost ? nPost : 1);
    run_grid(2, 256, [&] { k_resolve_records(f, post, &nPost, nPost, recs.data()); });

    // K9 + K10 (warp-collective kernels, restated here): every 8x8 tile, PBR bit from the clamp-to-edge gather
    const uint32_t mw = (W + 7) / 8, mh = (H + 7) / 8;
    std::vector<uint32_t> tileList; uint32_t tileCount = 0;
    for (uint32_t ty = 0; ty < mh; ty++) for (uint32_t tx = 0; tx < mw; tx++) {
        bool pbr = false;
        for (uint32_t l = 0; l < 64; l++) {
            const uint32_t px = min(tx * 8 + (l & 7u), W - 1), py = min(ty * 8 + (l >> 3), H - 1);
            const uint32_t id = (uint32_t)(vis64[(size_t)py * W + px] & 0xFFFFFFFFull);
            if (!id) continue;
            uint32_t tri, inst; dec_vis(id, tri, inst);
            pbr |= desc->materials[objects[post[inst].objectId].GLTFMaterialData].materialType == CHORD_LIGHTING_TYPE_GLTF_PBR;
        }
        tileList.push_back(tx * 8); tileList.push_back((ty * 8) | (pbr ? 0x80000000u : 0u)); tileCount++;
    }

    fill_lighting_consts(*view, c_lightingView[0]);          // the constant-bank slot the launcher fills with a device-to-device copy
    LightingTargets lc;
    static float srgb[256];
    for (int i = 0; i < 256; i++) { const double v = (double)i / 255.0; srgb[i] = (float)(v <= 0.04045 ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4)); }
    lc.srgbLut = srgb;
    lc.trLut = (const __half*)trLut; lc.irLut = (const __half*)irLut;
    for (int i = 0; i < CHORD_GBUF_COUNT; i++) { lc.planes[i] = packed ? nullptr : (float*)planes[i]; lc.packed[i] = packed ? planes[i] : nullptr; }
    const unsigned grid = (tileCount + 3) / 4;
    if (packed) run_grid(grid, 64, [&] { k_lighting<true, 0, false>(f, lc, recs.data(), tileList.data(), &tileCount, &nPost, nPost); });
    else        run_grid(grid, 64, [&] { k_lighting<false, 0, false>(f, lc, recs.data(), tileList.data(), &tileCount, &nPost, nPost); });
    return 0;
}

// div_uniform against the IEEE division: exhaustive over every float numerator for one denominator
extern "C" unsigned long long hostcc_div_uniform_mismatches(float w, unsigned lo, unsigned hi, unsigned* firstBad)
{
    const float rw = 1.0f / w;
    unsigned long long bad = 0;
    for (unsigned long long u = lo; u <= hi; u++) {
        const float x = __uint_as_float((unsigned)u);
        const float a = div_uniform(x, w, rw), b = x / w;
        if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) { if (!bad && firstBad) *firstBad = (unsigned)u; bad++; }
    }
    return bad;
}
