// Tuning aid (not part of the product): what does programmatic dependent launch buy on a chain of small dependent kernels
// inside a CUDA graph on this GPU / driver?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_chain pdl_chain.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool kPdl>
__global__ void __launch_bounds__(256) k_step(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned n, int chain)
{
    if (kPdl) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned v = in[i];
    for (int k = 0; k < chain; k++) v = in[(v + i) % n];          // dependent loads: latency-bound like the cull kernels
    out[i] = v + 1u;
}

template <bool kPdl>
static float run(unsigned* a, unsigned* b, unsigned n, int kernels, int grid, int chain, cudaStream_t s)
{
    cudaGraph_t g; cudaGraphExec_t e;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    for (int k = 0; k < kernels; k++) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        if (kPdl) { cfg.attrs = at; cfg.numAttrs = 1; }
        cudaLaunchKernelEx(&cfg, k_step<kPdl>, (const unsigned*)((k & 1) ? b : a), (k & 1) ? a : b, n, chain);
    }
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&e, g, 0);
    for (int i = 0; i < 5; i++) cudaGraphLaunch(e, s);
    cudaStreamSynchronize(s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    for (int i = 0; i < 50; i++) cudaGraphLaunch(e, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaGraphExecDestroy(e); cudaGraphDestroy(g);
    return ms / 50.0f * 1000.0f / kernels;                          // us per kernel
}

int main()
{
    const unsigned n = 1 << 20;
    unsigned *a, *b; cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    const int grids[] = {1, 148, 1184}; const int chains[] = {0, 4, 8};
    for (int g : grids) for (int c : chains) {
        const float t0 = run<false>(a, b, n, 30, g, c, s), t1 = run<true>(a, b, n, 30, g, c, s);
        printf("grid %5d chain %d: %.2f us/kernel plain, %.2f us/kernel PDL (err %s)\n", g, c, t0, t1, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
