// chord_b200_multigpu.cpp — native multi-rank host: one thread per rank (the ABI's "one host thread per GPU"), every rank one
// screen tile of the frame through chord_b200::Renderer, history HZB exchanged by the FUSED path (behind each rank's final
// buildHZB a put kernel stores its tile part straight into every rank's chain; in one process the peers are plain device
// pointers, between processes the same call takes CUDA IPC handles).  Ranks map onto the visible GPUs round-robin, so the same binary runs 2 / 4 / 8 ranks
// on an 8-GPU box and on a single GPU.
//
//   g++ -std=c++17 -pthread -Iinclude examples/chord_b200_multigpu.cpp -Lchord_b200 -l:libchord_b200.so -Wl,-rpath,$PWD/chord_b200 -o multigpu
//   ./multigpu in.blob out.blob [graph]
// in.blob  = the demo blob (scene + F frames) followed by: R, R tile rects, F x R tile instance views (tests write it)
// out.blob = per frame: the stitched packed surface (every rank contributes its tile), then rank 0's gathered min chain
#include <chord_b200.hpp>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
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

class Barrier {
public:
    explicit Barrier(unsigned n) : m_n(n) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_m);
        const unsigned gen = m_gen;
        if (++m_count == m_n) { m_count = 0; m_gen++; m_cv.notify_all(); }
        else m_cv.wait(lk, [&] { return gen != m_gen; });
    }
private:
    std::mutex m_m; std::condition_variable m_cv; unsigned m_n, m_count = 0, m_gen = 0;
};

}  // namespace

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s in.blob out.blob [graph]\n", argv[0]); return 2; }
    const bool useGraph = argc > 3 && std::strcmp(argv[3], "graph") == 0;
    Reader r(argv[1]);
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
    struct FrameIn { const chord_PerframeCameraView* view; const chord_InstanceCullingViewInfo* iview; const chord_GPUObjectGLTFPrimitive* objects; };
    std::vector<FrameIn> frames;
    {
        FrameIn f0;
        f0.view = (const chord_PerframeCameraView*)r.bytes(sizeof(chord_PerframeCameraView));
        f0.iview = (const chord_InstanceCullingViewInfo*)r.bytes(sizeof(chord_InstanceCullingViewInfo));
        f0.objects = (const chord_GPUObjectGLTFPrimitive*)r.bytes(sizeof(chord_GPUObjectGLTFPrimitive) * objectCount);
        frames.push_back(f0);
        const uint64_t more = r.u64();
        for (uint64_t i = 0; i < more; i++) {
            FrameIn fi;
            fi.view = (const chord_PerframeCameraView*)r.bytes(sizeof(chord_PerframeCameraView));
            fi.iview = (const chord_InstanceCullingViewInfo*)r.bytes(sizeof(chord_InstanceCullingViewInfo));
            fi.objects = (const chord_GPUObjectGLTFPrimitive*)r.bytes(sizeof(chord_GPUObjectGLTFPrimitive) * objectCount);
            frames.push_back(fi);
        }
    }
    const uint32_t world = (uint32_t)r.u64();
    if (world < 2 || world > CHORD_MAX_RANKS) { std::fprintf(stderr, "rank count %u\n", world); return 2; }
    std::vector<chord_rect> tiles(world);
    for (uint32_t q = 0; q < world; q++) { tiles[q].x0 = (uint32_t)r.u64(); tiles[q].y0 = (uint32_t)r.u64(); tiles[q].x1 = (uint32_t)r.u64(); tiles[q].y1 = (uint32_t)r.u64(); }
    std::vector<const chord_InstanceCullingViewInfo*> tileViews(frames.size() * world);
    for (size_t f = 0; f < frames.size(); f++)
        for (uint32_t q = 0; q < world; q++) tileViews[f * world + q] = (const chord_InstanceCullingViewInfo*)r.bytes(sizeof(chord_InstanceCullingViewInfo));

    const int devices = chord_device_count();
    if (devices < 1) { std::fprintf(stderr, "no CUDA device\n"); return 1; }
    std::vector<void*> blocks(world, nullptr);
    std::vector<std::vector<uint64_t>> stitched(frames.size(), std::vector<uint64_t>((size_t)width * height, 0ull));
    std::vector<uint16_t> chain0;
    std::vector<std::string> errors(world);
    Barrier barrier(world);
    auto rankMain = [&](uint32_t rank) {
        try {
            chord_b200::Context ctx((int)(rank % (uint32_t)devices));
            ctx.uploadScene(scene);
            chord_b200::Renderer ren(ctx, width, height, useGraph ? CHORD_RENDERER_GRAPH : 0u);
            blocks[rank] = ren.exchangeExport();
            barrier.wait();                                                    // every rank's block exists
            ren.exchangeConnect(rank, world, tiles.data(), nullptr, blocks.data());
            barrier.wait();                                                    // every rank has mapped every peer
            std::vector<uint64_t> vis((size_t)width * height);
            for (size_t f = 0; f < frames.size(); f++) {
                ren.render(*frames[f].view, *tileViews[f * world + rank], frames[f].objects, objectCount, lod0, groups);
                chord_b200::check(chord_readback_vis64(ren.frame(), vis.data()), "chord_readback_vis64");
                const chord_rect& t = tiles[rank];
                for (uint32_t y = t.y0; y < t.y1; y++)
                    std::memcpy(&stitched[f][(size_t)y * width + t.x0], &vis[(size_t)y * width + t.x0], 8ull * (t.x1 - t.x0));
                for (size_t i = 0; i < vis.size(); i++) {                       // nothing outside the tile
                    const uint32_t x = (uint32_t)(i % width), y = (uint32_t)(i / width);
                    if (vis[i] && !(x >= t.x0 && x < t.x1 && y >= t.y0 && y < t.y1)) throw std::runtime_error("texel outside the rank's tile");
                }
                barrier.wait();                                                // lock step (the read-backs above synchronise anyway)
            }
            ren.historySync();
            if (ren.exchangeTimedOut()) throw std::runtime_error("a peer never arrived");
            if (rank == 0) {
                uint32_t w0 = 0, h0 = 0, mips = 0;
                chord_b200::check(chord_hzb_info(ren.historyHzb(), &w0, &h0, &mips, nullptr, nullptr), "chord_hzb_info");
                for (uint32_t m = 0; m < mips; m++) {
                    const uint32_t mw = (w0 >> m) ? (w0 >> m) : 1u, mh = (h0 >> m) ? (h0 >> m) : 1u;
                    std::vector<uint16_t> lvl((size_t)mw * mh);
                    chord_b200::check(chord_hzb_readback(ren.historyHzb(), 0, m, lvl.data()), "chord_hzb_readback");
                    chain0.insert(chain0.end(), lvl.begin(), lvl.end());
                }
            }
            barrier.wait();                                                    // nobody tears its block down while a peer may still write
        } catch (const std::exception& e) {
            errors[rank] = e.what();
            std::fprintf(stderr, "[rank %u] %s\n", rank, e.what());
            std::exit(1);                                                      // (the other ranks would wait at the barrier for ever)
        }
    };
    std::vector<std::thread> threads;
    for (uint32_t q = 0; q < world; q++) threads.emplace_back(rankMain, q);
    for (auto& t : threads) t.join();
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) { std::perror(argv[2]); return 2; }
    for (auto& img : stitched) std::fwrite(img.data(), 8, img.size(), o);
    const uint64_t n = chain0.size();
    std::fwrite(&n, 8, 1, o);
    std::fwrite(chain0.data(), 2, chain0.size(), o);
    std::fclose(o);
    std::printf("%u ranks on %d device(s), %zu frames, %s submission, gathered min chain %llu texels\n", world, devices, frames.size(),
                useGraph ? "graph" : "eager", (unsigned long long)n);
    return 0;
}
