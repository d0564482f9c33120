// Tuning aid (not part of the product): what does a small memset / memcpy node between two dependent kernels cost inside a CUDA
// graph on this GPU / driver?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o memset_nodes memset_nodes.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_step(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + 1u;
}

static float run(int mode, unsigned* a, unsigned* b, unsigned* c, const unsigned* hostPinned, cudaStream_t s)
{
    const int kernels = 30;
    cudaGraph_t g; cudaGraphExec_t e;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    for (int k = 0; k < kernels; k++) {
        if (mode == 1) cudaMemsetAsync(c, 0, 32, s);
        if (mode == 2) cudaMemcpyAsync(c, hostPinned, 8, cudaMemcpyHostToDevice, s);
        if (mode == 3) cudaMemcpyAsync(c, c + 64, 256, cudaMemcpyDeviceToDevice, s);
        if (mode == 4) cudaMemsetAsync(c, 0, 8u << 20, s);
        k_step<<<148, 256, 0, s>>>((k & 1) ? b : a, (k & 1) ? a : b, 148 * 256);
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
    return ms / 50.0f * 1000.0f / kernels;
}

int main()
{
    unsigned *a, *b, *c, *h;
    cudaMalloc(&a, 1 << 20); cudaMalloc(&b, 1 << 20); cudaMalloc(&c, 16 << 20); cudaMallocHost(&h, 64);
    cudaMemset(a, 0, 1 << 20); cudaMemset(b, 0, 1 << 20);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    const char* names[] = {"kernel only", "+ 32 B memset node", "+ 8 B H2D memcpy node (pinned)", "+ 256 B D2D memcpy node", "+ 8 MB memset node"};
    for (int m = 0; m < 5; m++) printf("%-34s %.2f us per (node +) kernel   (%s)\n", names[m], run(m, a, b, c, h, s), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
