// hzb.cu — hierarchical-Z build (K8: hzb.hlsl:127-389, hzb_one.hlsl; host hzb.cpp:38-227) for sm_100a.
//
// A 256-thread block reduces a 64x64 depth region to mips 0..5 through shared memory (10 KB, 8 blocks / SM); mips 6.. are
// finished from the stored mip 5 by a one-block kernel launched right behind (the reference elects the last workgroup
// with a device counter instead, hzb.hlsl:289-300: same values, but the election would make every block carry the top
// levels' 40 KB of shared memory).  Depth comes from the
// high word of the packed 64-bit visibility texels.  Storage is R16F: min is stored round-toward-zero (conservative
// for reversed-Z, hzb.hlsl:69), max is stored RTZ for mips 0-4 and as half_bits(RTZ(v)) + 1 at mip 5 (hzb.hlsl:70),
// with mips 6.. derived from the stored mip 5.  fp32->fp16 RTZ is monotonic, so reducing stored fp16 values equals
// storing the fp32 reduction.
#include "common.cuh"

namespace chordb {

struct HzbBuildParams {
    const unsigned long long* vis64;
    uint32_t width, height;
    __half* minChain; __half* maxChain;     // either may be null
    uint64_t mipOffset[CHORD_HZB_MAX_MIP_COUNT];
    uint32_t mip0W, mip0H, mipCount;
    uint32_t* validRange;                    // null unless requested
    uint32_t* counter;
    uint32_t numBlocks;
    uint32_t blockX0, blockY0;               // first 64x64 block of the dispatched region (tile-restricted builds)
    // ---- fused exchange (multi-GPU): every texel of mips 0..5 this launch stores is ALSO stored into the same
    // position of every peer GPU's chain over NVLink (peer pointers: CUDA IPC / peer access), so that the next frame's
    // phase-0 cull finds the whole min chain on every GPU without a pack -> all-gather -> unpack sequence.  The block that
    // finishes last publishes this rank's valid range and its arrival flag (sequence number) to every peer.
    uint32_t nPeers, rank;                   // nPeers == 0: plain local build
    const chordb::ExchangePeers* peers;      // peer pointer table in device memory (indexed by a loop variable: not a kernel parameter)
    chordb::ExchangeHeader* localHeader;
    uint32_t parity;                         // which of the two history slots this build fills (ranges[parity])
};

__device__ __forceinline__ __half to_half_rtz(float v) { return __float2half_rz(v); }
__device__ __forceinline__ __half half_plus_ulp(__half h) { return __ushort_as_half((unsigned short)(__half_as_ushort(h) + 1)); }

__device__ __forceinline__ uint32_t mip_dim(uint32_t d0, uint32_t mip) { uint32_t d = d0 >> mip; return d ? d : 1u; }

// mips 6.. from the stored mip 5; executed by one block
__device__ void hzb_top_levels(const HzbBuildParams& p, float* __restrict__ sMinP, float* __restrict__ sMinQ,
                               float* __restrict__ sMaxP, float* __restrict__ sMaxQ)
{
    if (p.mipCount <= 6) return;
    uint32_t w = mip_dim(p.mip0W, 5), h = mip_dim(p.mip0H, 5);
    {
        // mip 5 (<= 64 x 64 texels, 256 threads): all of a thread's loads are issued before the first one is consumed - as a
        // rolled loop this was sixteen dependent round trips to L2 (11-13 us for a kernel that moves 16 KB)
        const uint32_t n = w * h;
        // (clamped, unpredicated addresses: with per-load predicates the compiler runs out of predicate registers and falls
        // back to batches of four)
        auto load_mip5 = [&](const __half* __restrict__ g, float* __restrict__ dst) {
            __half v[16];
#pragma unroll
            for (uint32_t k = 0; k < 16; k++) v[k] = __ldcg(g + min(threadIdx.x + k * 256u, n - 1u));
#pragma unroll
            for (uint32_t k = 0; k < 16; k++) { const uint32_t i = threadIdx.x + k * 256u; if (i < n) dst[i] = __half2float(v[k]); }
        };
        if (p.minChain) load_mip5(p.minChain + p.mipOffset[5], sMinP);
        if (p.maxChain) load_mip5(p.maxChain + p.mipOffset[5], sMaxP);
    }
    __syncthreads();
    float* srcMin = sMinP; float* dstMin = sMinQ; float* srcMax = sMaxP; float* dstMax = sMaxQ;
    for (uint32_t mip = 6; mip < p.mipCount; mip++) {
        const uint32_t nw = mip_dim(p.mip0W, mip), nh = mip_dim(p.mip0H, mip);
        for (uint32_t i = threadIdx.x; i < nw * nh; i += blockDim.x) {
            const uint32_t x = i % nw, y = i / nw;
            const uint32_t x0 = (w > 1) ? 2 * x : 0, x1 = (w > 1) ? 2 * x + 1 : 0, y0 = (h > 1) ? 2 * y : 0, y1 = (h > 1) ? 2 * y + 1 : 0;
            if (p.minChain) {
                const float v = fminf(fminf(fminf(srcMin[y0 * w + x0], srcMin[y0 * w + x1]), srcMin[y1 * w + x0]), srcMin[y1 * w + x1]);
                dstMin[i] = v; p.minChain[p.mipOffset[mip] + i] = to_half_rtz(v);
            }
            if (p.maxChain) {
                const float v = fmaxf(fmaxf(fmaxf(srcMax[y0 * w + x0], srcMax[y0 * w + x1]), srcMax[y1 * w + x0]), srcMax[y1 * w + x1]);
                dstMax[i] = v; p.maxChain[p.mipOffset[mip] + i] = to_half_rtz(v);
            }
        }
        __syncthreads();
        float* t = srcMin; srcMin = dstMin; dstMin = t; t = srcMax; srcMax = dstMax; dstMax = t;
        w = nw; h = nh;
    }
}

// Specialised on what the call builds: the min-only build between the two visibility stages (renderer.cpp:334) carries
// neither the max chain's reductions / shared-memory traffic nor the valid-range accumulation.
template <bool kMin, bool kMax, bool kRange>
__global__ void __launch_bounds__(256, (kMin && kMax) ? 6 : 8) k_hzb_build(HzbBuildParams p)
{
    __shared__ float sMinP[kMin ? 1024 : 1], sMinQ[kMin ? 256 : 1], sMaxP[kMax ? 1024 : 1], sMaxQ[kMax ? 256 : 1];
    const uint32_t bx = p.blockX0 + blockIdx.x, by = p.blockY0 + blockIdx.y;
    const uint32_t tid = threadIdx.x;
    constexpr bool doMin = kMin, doMax = kMax;
    const bool evenWidth = (p.width & 1u) == 0u;

    // ---- mip 0: 32x32 texels of this block, 4 per thread (hzb.hlsl:134-173)
    float vMin = 10.0f, vMax = -10.0f;      // valid-range accumulators
    auto mip0_texel = [&](uint32_t l, uint32_t x, uint32_t y, const float (&d)[4]) {
        float mn = d[0], mx = d[0];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (s) { if (kMin) mn = fminf(mn, d[s]); if (kMax) mx = fmaxf(mx, d[s]); }
            if (kRange) { if (d[s] > 0.0f) { vMax = fmaxf(vMax, d[s]); vMin = fminf(vMin, d[s]); } }
        }
        if (kMin) sMinP[l] = mn;
        if (kMax) sMaxP[l] = mx;
        if (x < p.mip0W && y < p.mip0H) {
            const size_t o = p.mipOffset[0] + (size_t)y * p.mip0W + x;
            if (doMin) p.minChain[o] = to_half_rtz(mn);
            if (doMax) p.maxChain[o] = to_half_rtz(mx);
        }
    };
    if (evenWidth && bx * 64 + 64 <= p.width && by * 64 + 64 <= p.height) {
        // block entirely inside the image (uniform per block): all eight 16-byte loads of the thread are issued before the
        // first one is consumed - with the edge handling in the way the compiler serialises the four texels, i.e. four
        // dependent round trips to HBM per thread instead of one
        ulonglong2 r0[4], r1[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t l = tid + k * 256, x = bx * 32 + (l & 31u), y = by * 32 + (l >> 5);
            r0[k] = *reinterpret_cast<const ulonglong2*>(p.vis64 + (size_t)(2 * y) * p.width + 2 * x);
            r1[k] = *reinterpret_cast<const ulonglong2*>(p.vis64 + (size_t)(2 * y + 1) * p.width + 2 * x);
        }
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t l = tid + k * 256, x = bx * 32 + (l & 31u), y = by * 32 + (l >> 5);
            const float d[4] = { __uint_as_float((uint32_t)(r0[k].x >> 32)), __uint_as_float((uint32_t)(r1[k].x >> 32)),
                                 __uint_as_float((uint32_t)(r0[k].y >> 32)), __uint_as_float((uint32_t)(r1[k].y >> 32)) };
            mip0_texel(l, x, y, d);
        }
    } else if (by * 64u >= p.height) {
        // block entirely below the image (at 4K: 30 of the 64 block rows of the power-of-two dispatch): every source row clamps
        // to the last image row, so the four texels of a thread (same column, rows 8 apart) see the same four pixels
        const uint32_t x = bx * 32 + (tid & 31u);
        float g[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t sx = min(2 * x + (uint32_t)(s >> 1), p.width - 1);
            g[s] = __uint_as_float((uint32_t)(p.vis64[(size_t)(p.height - 1) * p.width + sx] >> 32));
        }
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t l = tid + k * 256;
            mip0_texel(l, x, by * 32 + (l >> 5), g);
        }
    } else {
        // edge / off-screen block: clamp-to-edge gather (hzb.hlsl:139-150), again every load before the first use
        float g[4][4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t l = tid + k * 256, x = bx * 32 + (l & 31u), y = by * 32 + (l >> 5);
            const uint32_t sy0 = min(2 * y, p.height - 1), sy1 = min(2 * y + 1, p.height - 1);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint32_t sx = min(2 * x + (uint32_t)(s >> 1), p.width - 1), sy = (s & 1) ? sy1 : sy0;
                g[k][s] = __uint_as_float((uint32_t)(p.vis64[(size_t)sy * p.width + sx] >> 32));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t l = tid + k * 256;
            mip0_texel(l, bx * 32 + (l & 31u), by * 32 + (l >> 5), g[k]);
        }
    }
    if (kRange) {
        // hzb.hlsl:165-172: wave min/max, then InterlockedMin/Max on the raw bits with the <1 / >0 guards.  min / max are
        // associative, so the eight waves of the block are combined in shared memory first and ONE pair of atomics per
        // block reaches the two global words (65 k same-address atomics per 4K build otherwise).
        __shared__ float sRange[2][8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { vMin = fminf(vMin, __shfl_xor_sync(0xffffffffu, vMin, o)); vMax = fmaxf(vMax, __shfl_xor_sync(0xffffffffu, vMax, o)); }
        if ((tid & 31u) == 0) { sRange[0][tid >> 5] = vMin; sRange[1][tid >> 5] = vMax; }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int w = 1; w < 8; w++) { vMin = fminf(vMin, sRange[0][w]); vMax = fmaxf(vMax, sRange[1][w]); }
            if (vMin < 1.0f) atomicMin(&p.validRange[0], __float_as_uint(vMin));
            if (vMax > 0.0f) atomicMax(&p.validRange[1], __float_as_uint(vMax));
        }
    }
    __syncthreads();

    // ---- mips 1..5 through shared memory (ping-pong P <-> Q)
    float* srcMin = sMinP; float* dstMin = sMinQ; float* srcMax = sMaxP; float* dstMax = sMaxQ;
    uint32_t w = 32;
    for (uint32_t mip = 1; mip <= 5 && mip < p.mipCount; mip++) {
        const uint32_t nw = w >> 1;
        if (tid < nw * nw) {
            const uint32_t lx = tid % nw, ly = tid / nw;
            const uint32_t a = (2 * ly) * w + 2 * lx, b = a + 1, c = a + w, d = c + 1;
            float mn = 0.0f, mx = 0.0f;
            if (kMin) { mn = fminf(fminf(fminf(srcMin[a], srcMin[b]), srcMin[c]), srcMin[d]); dstMin[tid] = mn; }
            if (kMax) { mx = fmaxf(fmaxf(fmaxf(srcMax[a], srcMax[b]), srcMax[c]), srcMax[d]); dstMax[tid] = mx; }
            const uint32_t mw = mip_dim(p.mip0W, mip), mh = mip_dim(p.mip0H, mip);
            const uint32_t x = bx * nw + lx, y = by * nw + ly;
            if (x < mw && y < mh) {
                const size_t o = p.mipOffset[mip] + (size_t)y * mw + x;
                if (doMin) p.minChain[o] = to_half_rtz(mn);
                if (doMax) {
                    const __half hv = (mip == 5) ? half_plus_ulp(to_half_rtz(mx)) : to_half_rtz(mx);
                    p.maxChain[o] = hv;
                }
            }
        }
        __syncthreads();
        float* t = srcMin; srcMin = dstMin; dstMin = t; t = srcMax; srcMax = dstMax; dstMax = t;
        w = nw;
    }
}

// ---- fused exchange, producer side (multi-GPU) ----------------------------------------------------------------
// This rank's part of mips 0..5 of both chains - built into its own history chain by k_hzb_build just before - is pushed into
// the same positions of EVERY peer's chain over NVLink (peer pointers: CUDA IPC / peer access): one warp per (chain, mip, row)
// of the tile region, 16-byte loads from the local chain (L2 hits) and 16-byte stores per peer, i.e. rows of up to 960
// contiguous bytes per peer.  (Round 2 first did these stores out of k_hzb_build's own epilogue, one 2-byte store per texel
// and peer - 64 contiguous bytes per warp instruction and a system-scope fence per 64 x 64 block: ~280 GB/s, 40 us for one
// peer at 2 GPUs, ~70 us at 8; profiles/r02_o_*.)  The block that finishes last publishes this rank's valid depth range and -
// after a system-scope fence - its arrival flag to every peer; k_hzb_exchange_wait_top is the consumer.
struct TileRegion { uint32_t bx0, by0, bx1, by1; };
__global__ void __launch_bounds__(256) k_hzb_exchange_put(HzbBuildParams p, TileRegion r, uint32_t rowsPerChain, int kRange, uint32_t seqParam)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (tid >> 5); w < 2u * rowsPerChain; w += warpsPerGrid) {
        const uint32_t which = w / rowsPerChain;                       // 0 = min chain, 1 = max chain
        uint32_t k = w % rowsPerChain, mip = 0;
        while (k >= (((r.by1 - r.by0) * 32u) >> mip)) { k -= ((r.by1 - r.by0) * 32u) >> mip; mip++; }
        const uint32_t rw = ((r.bx1 - r.bx0) * 32u) >> mip;            // texels of this row inside the region
        const uint32_t x0 = (r.bx0 * 32u) >> mip, y = ((r.by0 * 32u) >> mip) + k;
        const size_t o = p.mipOffset[mip] + (size_t)y * mip_dim(p.mip0W, mip) + x0;
        const __half* src = (which ? p.maxChain : p.minChain) + o;
        if (((o | (size_t)rw) & 7u) == 0u) {
            // (every chain starts on a 256-byte boundary: o % 8 == 0 makes source and destinations 16-byte aligned)
            const uint4* src16 = reinterpret_cast<const uint4*>(src);
            for (uint32_t i = lane; i < rw / 8u; i += 32u) {
                const uint4 v = src16[i];
                for (uint32_t q = 0; q < p.nPeers; q++) reinterpret_cast<uint4*>((which ? p.peers->max[q] : p.peers->min[q]) + o)[i] = v;
            }
        } else {
            for (uint32_t i = lane; i < rw; i += 32u) {
                const __half v = src[i];
                for (uint32_t q = 0; q < p.nPeers; q++) (which ? p.peers->max[q] : p.peers->min[q])[o + i] = v;
            }
        }
    }
    {
        // "last block done": every block makes its peer stores visible system-wide and counts itself; the block that brings
        // the count to numBlocks knows that every store of this launch (all blocks, all peers) has been performed, publishes
        // this rank's valid depth range and then - after another system-scope fence - the arrival flag of this frame.
        __shared__ uint32_t sLast;
        __threadfence_system();
        __syncthreads();
        if (tid == 0) sLast = (atomicAdd(p.counter, 1u) + 1u == p.numBlocks) ? 1u : 0u;
        __syncthreads();
        if (sLast) {
            __threadfence_system();
            uint32_t seq = 0;
            if (tid == 0) {
                *p.counter = 0u;                                             // re-armed for the next launch (stream-ordered)
                // seqParam == 0: this kernel advances the rank's sequence number itself (it runs in stream order behind the
                // build); otherwise k_hzb_exchange_advance did that in stream order and the host hands the number in - the
                // deferred put of the asynchronous renderer may run after later passes have read the counter
                if (seqParam == 0u) { seq = p.localHeader->localSeq + 1u; p.localHeader->localSeq = seq; }
            }
            __syncthreads();
            if (tid <= p.nPeers) {
                chordb::ExchangeHeader* dst = tid < p.nPeers ? p.peers->header[tid] : p.localHeader;
                if (kRange) {
                    // stream-ordered put: the chain's valid-range words are final (the build precedes this launch).  Deferred
                    // put: k_hzb_exchange_advance saved them in this rank's own header entry - by now the next frame's
                    // k_hzb_exchange_wait_top may have folded all ranks' ranges into the chain's words
                    const volatile uint32_t* vr = seqParam ? p.localHeader->ranges[p.parity][p.rank] : p.validRange;
                    const uint32_t lo = vr[0], hi = vr[1];
                    dst->ranges[p.parity][p.rank][0] = lo; dst->ranges[p.parity][p.rank][1] = hi;
                }
            }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                // broadcast the sequence number to thread 0's peers: one thread writes all flags so that they follow the fence
                const uint32_t s2 = seqParam ? seqParam : p.localHeader->localSeq;
                for (uint32_t q = 0; q < p.nPeers; q++) *((volatile uint32_t*)&p.peers->header[q]->arrived[p.rank]) = s2;
                *((volatile uint32_t*)&p.localHeader->arrived[p.rank]) = s2;
            }
            (void)seq;
        }
    }
}

// Consumer side of the fused exchange: runs (one block) in front of the first reader of a gathered history chain.  Waits until
// every peer's arrival flag has reached this rank's own sequence number (all ranks render the same frames, so the numbers
// advance in lock step; a peer may be one frame ahead - it then writes the OTHER history slot), folds the ranks' valid
// ranges into the chain's validRange words and rebuilds mips 6.. from the gathered mip 5.  The wait is bounded: if a peer
// never arrives (it died), status[0] is set and the kernel leaves instead of hanging the GPU.
__global__ void __launch_bounds__(256) k_hzb_exchange_wait_top(HzbBuildParams p, uint32_t world, uint32_t* __restrict__ status, uint32_t maxSpins)
{
    __shared__ float sMinP[4096], sMinQ[1024], sMaxP[4096], sMaxQ[1024];
    __shared__ uint32_t sFail;
    if (threadIdx.x == 0) sFail = 0u;
    __syncthreads();
    const uint32_t target = p.localHeader->localSeq;
    if (threadIdx.x < world && threadIdx.x != p.rank) {
        const volatile uint32_t* flag = &p.localHeader->arrived[threadIdx.x];
        uint32_t spins = 0;
        while ((int32_t)(*flag - target) < 0) {
            if (++spins > maxSpins) { sFail = 1u; break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (sFail) { if (threadIdx.x == 0) atomicExch(&status[0], 1u); return; }
    if (threadIdx.x == 0 && p.validRange) {
        uint32_t lo = ~0u, hi = 0u;
        for (uint32_t r = 0; r < world; r++) {
            const volatile uint32_t* rr = p.localHeader->ranges[p.parity][r];
            lo = min(lo, rr[0]); hi = max(hi, rr[1]);
        }
        p.validRange[0] = lo; p.validRange[1] = hi;
    }
    hzb_top_levels(p, sMinP, sMinQ, sMaxP, sMaxQ);
}

__global__ void __launch_bounds__(256) k_hzb_top(HzbBuildParams p)
{
    __shared__ float sMinP[4096], sMinQ[1024], sMaxP[4096], sMaxQ[1024];
    hzb_top_levels(p, sMinP, sMinQ, sMaxP, sMaxQ);
}

static HzbBuildParams make_params(chord_frame* fr, int bMin, int bMax, int bRange, chord_hzb* dst)
{
    HzbBuildParams p;
    p.vis64 = fr->vis64; p.width = fr->width; p.height = fr->height;
    p.minChain = bMin ? dst->minChain : nullptr; p.maxChain = bMax ? dst->maxChain : nullptr;
    for (int i = 0; i < CHORD_HZB_MAX_MIP_COUNT; i++) p.mipOffset[i] = dst->mipOffset[i];
    p.mip0W = dst->mip0W; p.mip0H = dst->mip0H; p.mipCount = dst->mipCount;
    p.validRange = bRange ? dst->validRange : nullptr;
    p.counter = dst->counter;
    p.blockX0 = p.blockY0 = 0; p.numBlocks = 0;
    p.nPeers = 0; p.rank = 0; p.parity = 0; p.localHeader = nullptr; p.peers = nullptr;
    return p;
}

// 64x64-pixel block range a screen tile owns (tiles touching the right / bottom frame edge also own the clamped
// off-screen blocks up to the power-of-two extent, so that the union of all tiles is the reference's full dispatch)
static void tile_blocks(const chord_hzb* h, const chord_rect& t, uint32_t frameW, uint32_t frameH, uint32_t& bx0, uint32_t& by0, uint32_t& bx1, uint32_t& by1)
{
    const uint32_t gw = (h->mip0W + 31) / 32, gh = (h->mip0H + 31) / 32;
    bx0 = t.x0 / 64; by0 = t.y0 / 64;
    bx1 = (t.x1 >= frameW) ? gw : t.x1 / 64;
    by1 = (t.y1 >= frameH) ? gh : t.y1 / 64;
}

static bool tile_aligned(const chord_rect& t, uint32_t w, uint32_t h)
{ return (t.x0 % 64 == 0) && (t.y0 % 64 == 0) && (t.x1 % 64 == 0 || t.x1 == w) && (t.y1 % 64 == 0 || t.y1 == h); }

// the rank's own sequence number, advanced in stream order behind a build whose put is deferred to another stream
// (and this rank's own valid-range entry of the slot: the next frame's k_hzb_exchange_wait_top folds it in whether or not the
// put - which carries it to the peers - has run yet)
__global__ void k_hzb_exchange_advance(chordb::ExchangeHeader* __restrict__ header, const uint32_t* __restrict__ validRange, uint32_t parity, uint32_t rank)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        header->localSeq = header->localSeq + 1u;
        if (validRange) { header->ranges[parity][rank][0] = validRange[0]; header->ranges[parity][rank][1] = validRange[1]; }
    }
}

static uint32_t exchange_put_rows(const chord_hzb* dst, const TileRegion& region)
{
    uint32_t rows = 0;
    for (uint32_t m = 0; m < 6 && m < dst->mipCount; m++) rows += ((region.by1 - region.by0) * 32u) >> m;
    return rows;
}

// producer side of the fused exchange on its own (the asynchronous renderer runs it on a stream beside the next passes): the
// tile's rows of mips 0..5 of `dst` (complete in this rank's own chain) go to every peer, then the valid range and the flags
int launch_hzb_exchange_put(chord_frame* fr, int bRange, chord_hzb* dst, const HzbExchange* ex, uint32_t seq)
{
    chord_ctx* ctx = fr->ctx;
    CHORD_REQUIRE(ex && ex->nPeers, "no exchange");
    HzbBuildParams p = make_params(fr, 1, 1, bRange, dst);
    p.nPeers = ex->nPeers; p.rank = ex->rank; p.parity = ex->parity; p.localHeader = ex->localHeader; p.peers = ex->dPeers;
    CHORD_REQUIRE(tile_aligned(fr->tile, fr->width, fr->height), "tile rect must be cut on multiples of 64 px for a tiled HZB build");
    TileRegion region;
    tile_blocks(dst, fr->tile, fr->width, fr->height, region.bx0, region.by0, region.bx1, region.by1);
    const uint32_t rows = exchange_put_rows(dst, region);
    uint32_t pgrid = (2u * rows + 7u) / 8u; if (pgrid > (uint32_t)ctx->smCount * 8u) pgrid = (uint32_t)ctx->smCount * 8u; if (pgrid == 0u) pgrid = 1u;
    p.numBlocks = pgrid;
    k_hzb_exchange_put<<<pgrid, 256, 0, ctx->stream>>>(p, region, rows, bRange ? 1 : 0, seq);
    ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

int launch_build_hzb(chord_frame* fr, int bMin, int bMax, int bRange, chord_hzb* dst, const HzbExchange* ex, bool deferPut)
{
    chord_ctx* ctx = fr->ctx;
    HzbBuildParams p = make_params(fr, bMin, bMax, bRange, dst);
    if (ex && ex->nPeers) {
        CHORD_REQUIRE(bMin, "the exchanged history chain is the min chain");
        CHORD_REQUIRE(bMax, "an exchanged build stores both chains (the peer table always carries max pointers)");
        p.nPeers = ex->nPeers; p.rank = ex->rank; p.parity = ex->parity; p.localHeader = ex->localHeader; p.peers = ex->dPeers;
    }
    dim3 grid((dst->mip0W + 31) / 32, (dst->mip0H + 31) / 32);            // hzb.cpp:74
    const bool fullFrame = fr->tile.x0 == 0 && fr->tile.y0 == 0 && fr->tile.x1 == fr->width && fr->tile.y1 == fr->height;
    CHORD_REQUIRE(!(p.nPeers && fullFrame), "an exchanged build renders one screen tile of several");
    if (!fullFrame) {
        // screen-tile partition: build only the blocks this tile owns; everything else reads as 0 (= far,
        // never occludes) until the tiles of the other GPUs are gathered in and chord_hzb_finish_top rebuilds mips 6..
        CHORD_REQUIRE(tile_aligned(fr->tile, fr->width, fr->height), "tile rect must be cut on multiples of 64 px for a tiled HZB build");
        uint32_t bx0, by0, bx1, by1; tile_blocks(dst, fr->tile, fr->width, fr->height, bx0, by0, bx1, by1);
        p.blockX0 = bx0; p.blockY0 = by0; grid = dim3(bx1 - bx0, by1 - by0);
        // an exchanged chain is NOT cleared: every texel of mips 0..5 has exactly one owner who rewrites it each frame, and a
        // peer that runs ahead may already have stored its part of this frame into it
        if (bMin && !p.nPeers) CHORD_CUDA(cudaMemsetAsync(dst->minChain, 0, sizeof(__half) * dst->totalTexels, ctx->stream));
        if (bMax && !p.nPeers) CHORD_CUDA(cudaMemsetAsync(dst->maxChain, 0, sizeof(__half) * dst->totalTexels, ctx->stream));
    }
    p.numBlocks = grid.x * grid.y;
    if (bRange) {
        const int rcFill = launch_fill_u32(ctx, dst->validRange, ~0u, 0u, 2);   // hzb.cpp:105-106: {min bits, max bits} seed
        if (rcFill) return rcFill;
    }
    // hzb.cpp:44-47: at least one of min / max, a valid range only with both
    if (bMin && bMax && bRange) k_hzb_build<true, true, true><<<grid, 256, 0, ctx->stream>>>(p);
    else if (bMin && bMax)      k_hzb_build<true, true, false><<<grid, 256, 0, ctx->stream>>>(p);
    else if (bMin)              k_hzb_build<true, false, false><<<grid, 256, 0, ctx->stream>>>(p);
    else                        k_hzb_build<false, true, false><<<grid, 256, 0, ctx->stream>>>(p);
    ctx->launches++;
    // mips 6..: from the (tile-)local mip 5 now, like the phase-1 HZB wants; an exchanged chain gets them from the GATHERED
    // mip 5 later (launch_hzb_exchange_wait_top, in front of the first reader)
    if (dst->mipCount > 6 && !p.nPeers) { k_hzb_top<<<1, 256, 0, ctx->stream>>>(p); ctx->launches++; }
    CHORD_CUDA(cudaGetLastError());
    // producer side of the fused exchange: right behind the build, unless the caller launches it itself (on another stream)
    if (p.nPeers && !deferPut) return launch_hzb_exchange_put(fr, bRange, dst, ex, 0u);
    if (p.nPeers) {
        // deferred put: the sequence number still advances here, in stream order (k_hzb_exchange_wait_top of the next frame
        // reads it as its target); the caller launches the put with the same number
        k_hzb_exchange_advance<<<1, 32, 0, ctx->stream>>>(p.localHeader, bRange ? dst->validRange : nullptr, p.parity, p.rank);
        ctx->launches++;
        CHORD_CUDA(cudaGetLastError());
    }
    return CHORD_OK;
}

// NCCL path: the ranks' {minBits, maxBits} pairs (all-gathered) folded into the chain's valid range
__global__ void k_hzb_range_reduce(const uint32_t* __restrict__ pairs, uint32_t world, uint32_t* __restrict__ validRange)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t lo = ~0u, hi = 0u;
    for (uint32_t r = 0; r < world; r++) { lo = min(lo, pairs[2 * r]); hi = max(hi, pairs[2 * r + 1]); }
    validRange[0] = lo; validRange[1] = hi;
}

int launch_hzb_range_reduce(chord_ctx* ctx, const uint32_t* pairs, uint32_t world, uint32_t* validRange)
{
    k_hzb_range_reduce<<<1, 32, 0, ctx->stream>>>(pairs, world, validRange);
    ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

int launch_hzb_exchange_wait_top(chord_frame* fr, chord_hzb* dst, const HzbExchange* ex, uint32_t world, uint32_t* status, uint32_t maxPolls)
{
    HzbBuildParams p = make_params(fr, 1, 1, 1, dst);
    p.rank = ex->rank; p.parity = ex->parity; p.localHeader = ex->localHeader;
    // 200 ns per poll; the default 20 M polls = ~4 s before the wait gives up (a peer that died must not hang this GPU;
    // chord_renderer_set_exchange_timeout)
    k_hzb_exchange_wait_top<<<1, 256, 0, fr->ctx->stream>>>(p, world, status, maxPolls ? maxPolls : 20u * 1000u * 1000u);
    fr->ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

}  // namespace chordb

namespace chordb {

__device__ __forceinline__ uint64_t tile_region_texels(const TileRegion& r, uint32_t mip)
{ return (uint64_t)(((r.bx1 - r.bx0) * 32u) >> mip) * (((r.by1 - r.by0) * 32u) >> mip); }

// one thread per packed texel of mips 0..5 of one tile; pack = chain -> contiguous, unpack = contiguous -> chain
struct TileCopyParams { TileRegion regions[16]; uint64_t mipOffset6[6]; uint32_t nRegions; };

template <bool kPack>
__global__ void k_hzb_tile_copy(__half* __restrict__ chain, TileCopyParams prm, uint32_t mip0W, __half* __restrict__ packed, uint64_t slotTexels)
{
    const uint32_t ri = blockIdx.y;
    if (ri >= prm.nRegions) return;
    const TileRegion r = prm.regions[ri];
    const uint64_t* mipOffset6 = prm.mipOffset6;
    uint64_t total = 0;
#pragma unroll
    for (uint32_t m = 0; m < 6; m++) total += tile_region_texels(r, m);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k = i; uint32_t mip = 0;
        while (k >= tile_region_texels(r, mip)) { k -= tile_region_texels(r, mip); mip++; }
        const uint32_t rw = ((r.bx1 - r.bx0) * 32u) >> mip;
        const uint32_t x = ((r.bx0 * 32u) >> mip) + (uint32_t)(k % rw), y = ((r.by0 * 32u) >> mip) + (uint32_t)(k / rw);
        __half* c = chain + mipOffset6[mip] + (uint64_t)y * (mip0W >> mip) + x;
        __half* p = packed + (uint64_t)ri * slotTexels + i;
        if (kPack) *p = *c; else *c = *p;
    }
}

}  // namespace chordb

extern "C" int chord_hzb_tile_bytes(const chord_hzb* h, const chord_rect* tile, uint64_t* bytes)
{
    using namespace chordb;
    CHORD_REQUIRE(h && tile && bytes, "null");
    CHORD_REQUIRE(tile_aligned(*tile, h->depthW, h->depthH), "tile rect must be cut on multiples of 64 px");
    uint32_t bx0, by0, bx1, by1; tile_blocks(h, *tile, h->depthW, h->depthH, bx0, by0, bx1, by1);
    uint64_t t = 0;
    for (uint32_t m = 0; m < 6; m++) t += (uint64_t)(((bx1 - bx0) * 32u) >> m) * (((by1 - by0) * 32u) >> m);
    *bytes = t * sizeof(__half);
    return CHORD_OK;
}

static int hzb_tile_copy(chord_hzb* h, int which, const chord_rect* tiles, uint32_t n, void* packed, uint64_t slotBytes, bool pack)
{
    using namespace chordb;
    CHORD_REQUIRE(h && tiles && packed && n > 0 && n <= 16, "args (at most 16 tiles)");
    CHORD_REQUIRE(h->mipCount >= 6, "mips");
    chord_ctx* ctx = h->ctx;
    TileCopyParams prm;                                  // by value: no allocation, no copy, one launch
    prm.nRegions = n;
    for (uint32_t i = 0; i < n; i++) {
        CHORD_REQUIRE(tile_aligned(tiles[i], h->depthW, h->depthH), "tile rect must be cut on multiples of 64 px");
        tile_blocks(h, tiles[i], h->depthW, h->depthH, prm.regions[i].bx0, prm.regions[i].by0, prm.regions[i].bx1, prm.regions[i].by1);
    }
    for (int m = 0; m < 6; m++) prm.mipOffset6[m] = h->mipOffset[m];
    __half* chain = which ? h->maxChain : h->minChain;
    const dim3 grid(128, n);
    if (pack) k_hzb_tile_copy<true><<<grid, 256, 0, ctx->stream>>>(chain, prm, h->mip0W, (__half*)packed, slotBytes / 2);
    else      k_hzb_tile_copy<false><<<grid, 256, 0, ctx->stream>>>(chain, prm, h->mip0W, (__half*)packed, slotBytes / 2);
    ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}

extern "C" int chord_hzb_pack_tile(chord_hzb* h, int which, const chord_rect* tile, void* dstDevice)
{ return hzb_tile_copy(h, which, tile, 1, dstDevice, 0, true); }

extern "C" int chord_hzb_unpack_tiles(chord_hzb* h, int which, const chord_rect* tiles, uint32_t n, const void* srcDevice, uint64_t slotBytes)
{ return hzb_tile_copy(h, which, tiles, n, (void*)srcDevice, slotBytes, false); }

// Multi-GPU helper: after the mips 0..5 of every tile have been gathered into dst, rebuild mips 6...
extern "C" int chord_hzb_finish_top(chord_frame* fr, int bMin, int bMax, chord_hzb* dst)
{
    using namespace chordb;
    CHORD_REQUIRE(fr && dst, "null");
    HzbBuildParams p = make_params(fr, bMin, bMax, 0, dst);
    k_hzb_top<<<1, 256, 0, fr->ctx->stream>>>(p);
    fr->ctx->launches++;
    CHORD_CUDA(cudaGetLastError());
    return CHORD_OK;
}
