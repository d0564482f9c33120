// chord_b200_demo.cpp — native C++ host driving the visibility-buffer path through the C++ shim (include/chord_b200.hpp),
// the way chord's DeferredRenderer::render would (renderer.cpp:311-363).  Reads a scene/frame blob (written by
// tests/test_gpu_parity.py from the seeded cases), renders one frame with the flat or the BVH instance cull, and writes the
// packed visibility surface + the post-cull list so the test can compare them with the oracle.
//
//   g++ -std=c++17 -Iinclude examples/chord_b200_demo.cpp -Lchord_b200 -l:libchord_b200.so -Wl,-rpath,$PWD/chord_b200 -o demo
//   ./demo in.blob out.blob [bvh]
//   ./demo in.blob out.blob path       every frame of the blob through chord_b200::DeferredRenderer (two-phase HZB culling against
//                                      the previous frame's HZB, renderer.cpp:311-363 + :489); one {count, list, surface} per frame
#include <chord_b200.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace {

struct Reader {
    std::vector<unsigned char> buf; size_t off = 0;
    explicit Reader(const char* path)
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) { std::perror(path); std::exit(2); }
        std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
        buf.resize((size_t)n);
        if (std::fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { std::perror("read"); std::exit(2); }
        std::fclose(f);
    }
    uint64_t u64() { uint64_t v; std::memcpy(&v, buf.data() + off, 8); off += 8; return v; }
    const void* bytes(size_t n) { const void* p = n ? buf.data() + off : nullptr; off += (n + 7) & ~size_t(7); return p; }
};

}  // namespace

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s in.blob out.blob [bvh]\n", argv[0]); return 2; }
    const bool useBvh = argc > 3 && std::strcmp(argv[3], "bvh") == 0;
    const bool path = argc > 3 && std::strcmp(argv[3], "path") == 0;
    Reader r(argv[1]);
    // header: width, height, objectCount, lod0MeshletCount, meshletGroupCount, then array sizes
    const uint32_t width = (uint32_t)r.u64(), height = (uint32_t)r.u64(), objectCount = (uint32_t)r.u64();
    const uint32_t lod0 = (uint32_t)r.u64(), groups = (uint32_t)r.u64();
    chord_primitive_datas_desc d{};
    d.meshletCount = (uint32_t)r.u64(); d.vertexCount = (uint32_t)r.u64(); d.meshletDataCount = (uint32_t)r.u64();
    d.bvhNodeCount = (uint32_t)r.u64(); d.groupCount = (uint32_t)r.u64(); d.groupIndicesCount = (uint32_t)r.u64();
    const uint32_t primCount = (uint32_t)r.u64(), matCount = (uint32_t)r.u64();
    d.meshlets = (const chord_GPUGLTFMeshlet*)r.bytes(sizeof(chord_GPUGLTFMeshlet) * d.meshletCount);
    d.positions = (const float*)r.bytes(12ull * d.vertexCount);
    d.meshletData = (const uint32_t*)r.bytes(4ull * d.meshletDataCount);
    d.bvhNodes = (const chord_GPUBVHNode*)r.bytes(sizeof(chord_GPUBVHNode) * d.bvhNodeCount);
    d.groups = (const chord_GPUGLTFMeshletGroup*)r.bytes(sizeof(chord_GPUGLTFMeshletGroup) * d.groupCount);
    d.groupIndices = (const uint32_t*)r.bytes(4ull * d.groupIndicesCount);
    chord_scene_desc scene{};
    scene.datas = &d; scene.datasCount = 1;
    scene.primitives = (const chord_GLTFPrimitiveBuffer*)r.bytes(sizeof(chord_GLTFPrimitiveBuffer) * primCount); scene.primitiveCount = primCount;
    scene.materials = (const chord_GLTFMaterialGPUData*)r.bytes(sizeof(chord_GLTFMaterialGPUData) * matCount); scene.materialCount = matCount;
    const auto* view = (const chord_PerframeCameraView*)r.bytes(sizeof(chord_PerframeCameraView));
    const auto* iview = (const chord_InstanceCullingViewInfo*)r.bytes(sizeof(chord_InstanceCullingViewInfo));
    const auto* objects = (const chord_GPUObjectGLTFPrimitive*)r.bytes(sizeof(chord_GPUObjectGLTFPrimitive) * objectCount);

    // optional tail of the blob: more frames of the same scene (count, then view / instance view / objects per frame)
    struct FrameIn { const chord_PerframeCameraView* view; const chord_InstanceCullingViewInfo* iview; const chord_GPUObjectGLTFPrimitive* objects; };
    std::vector<FrameIn> frames{FrameIn{view, iview, objects}};
    if (r.off < r.buf.size()) {
        const uint64_t more = r.u64();
        for (uint64_t i = 0; i < more; i++) {
            FrameIn fi;
            fi.view = (const chord_PerframeCameraView*)r.bytes(sizeof(chord_PerframeCameraView));
            fi.iview = (const chord_InstanceCullingViewInfo*)r.bytes(sizeof(chord_InstanceCullingViewInfo));
            fi.objects = (const chord_GPUObjectGLTFPrimitive*)r.bytes(sizeof(chord_GPUObjectGLTFPrimitive) * objectCount);
            frames.push_back(fi);
        }
    }

    if (path) {
        try {
            // Lifetime as in the reference (ref-counted pool resources, command_list.h:23-38): the Context OBJECT goes away right
            // after the renderer has been created; the renderer's frame / lists / HZBs keep the chord_ctx alive and usable.
            auto owner = std::make_unique<chord_b200::Context>(0);
            owner->uploadScene(scene);
            chord_b200::DeferredRenderer renderer(*owner, width, height);
            chord_ctx* rawCtx = owner->handle();
            owner.reset();                                                        // chord_ctx_destroy: drops the owner's reference only
            if (chord_ctx_refcount(rawCtx) < 1) { std::fprintf(stderr, "context died with children alive\n"); return 3; }
            FILE* o = std::fopen(argv[2], "wb");
            if (!o) { std::perror(argv[2]); return 2; }
            std::vector<uint64_t> vis((size_t)width * height);
            std::vector<chord_draw_cmd> cmds(lod0);
            for (size_t fi = 0; fi < frames.size(); fi++) {
                auto post = renderer.render(*frames[fi].view, *frames[fi].iview, frames[fi].objects, objectCount, lod0, groups);
                chord_b200::check(chord_readback_vis64(renderer.frame().handle(), vis.data()), "chord_readback_vis64");
                uint32_t n = 0;
                chord_b200::check(chord_cmdlist_readback(post.get(), cmds.data(), lod0, &n), "chord_cmdlist_readback");
                const uint64_t n64 = n;
                std::fwrite(&n64, 8, 1, o);
                std::fwrite(cmds.data(), sizeof(chord_draw_cmd), n, o);
                std::fwrite(vis.data(), 8, vis.size(), o);
                size_t covered = 0; for (uint64_t v : vis) covered += v != 0;
                std::printf("frame %zu: %u meshlets after culling, %zu px covered, history HZB %s\n", fi, n, covered, fi ? "used" : "none");
            }
            std::fclose(o);
            std::printf("%llu kernel launches, %u handles keep the context alive\n", (unsigned long long)chord_ctx_kernel_launches(rawCtx), chord_ctx_refcount(rawCtx));
        } catch (const std::exception& e) {
            std::fprintf(stderr, "chord_b200 error: %s\n", e.what());
            return 1;
        }
        return 0;
    }

    try {
        chord_b200::Context ctx(0);
        ctx.uploadScene(scene);
        auto frame = ctx.createFrame(width, height);
        frame->begin(*view, *iview, objects, objectCount, lod0, groups);
        chord_b200::CmdListRef post;
        if (useBvh) {
            post = ctx.createCmdList(lod0);
            chord_b200::check(chord_instance_culling_bvh(frame->handle(), post.get()), "chord_instance_culling_bvh");
        } else {
            post = frame->instanceCulling();
        }
        chord_b200::CmdListRef rejected;
        const bool stage1 = frame->visibilityStage0(nullptr, post, rejected);       // no history on the first frame
        if (stage1) { std::fprintf(stderr, "unexpected stage 1 without history\n"); return 3; }
        auto hzb = frame->buildHZB(true, true, true);
        frame->visibilityMark(post);

        std::vector<uint64_t> vis((size_t)width * height);
        chord_b200::check(chord_readback_vis64(frame->handle(), vis.data()), "chord_readback_vis64");
        std::vector<chord_draw_cmd> cmds(lod0);
        uint32_t n = 0;
        chord_b200::check(chord_cmdlist_readback(post.get(), cmds.data(), lod0, &n), "chord_cmdlist_readback");
        FILE* o = std::fopen(argv[2], "wb");
        if (!o) { std::perror(argv[2]); return 2; }
        const uint64_t n64 = n;
        std::fwrite(&n64, 8, 1, o);
        std::fwrite(cmds.data(), sizeof(chord_draw_cmd), n, o);
        std::fwrite(vis.data(), 8, vis.size(), o);
        std::fclose(o);
        size_t covered = 0; for (uint64_t v : vis) covered += v != 0;
        std::printf("rendered %ux%u: %u meshlets after culling (%s), %zu px covered, %llu kernel launches\n", width, height, n,
                    useBvh ? "BVH traversal" : "flat loop", covered, (unsigned long long)chord_ctx_kernel_launches(ctx.handle()));
    } catch (const std::exception& e) {
        std::fprintf(stderr, "chord_b200 error: %s\n", e.what());
        return 1;
    }
    return 0;
}
