// TEST INFRASTRUCTURE ONLY -- not a CUDA runtime and not a CPU fallback of the library.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile chord_b200/csrc/*.cu as plain C++ so that the tests can EXECUTE
// THE SOURCE TEXT OF THE KERNELS on the CPU and hold it against the oracle when no GPU is at hand:
//   * tests/test_kernel_source_on_host.py runs the per-thread resolve kernels thread by thread (lighting_harness.cpp);
//   * tests/test_hostcc_parity.py builds the whole library this way (hostcc_build.py) and runs the GPU parity cases on it:
//     every CUDA thread of a block is a coroutine (hostcc_rt.cpp), __syncthreads / __syncwarp / shuffles / ballots are
//     rendezvous points, blocks run one after the other, atomics are plain read-modify-writes.
// Nothing in the product links, loads or knows what is built from this header: libchord_b200.so is nvcc output for sm_100a
// only and fails loudly without a device (tests/test_abi.py::test_no_cpu_fallback_without_gpu).  The scratch libraries live
// in pytest temporary directories.
//
// Semantics: every *_rn intrinsic is the IEEE operation it names; the approximate ones (__fdividef, rsqrtf, MUFU
// sequences) are evaluated exactly here, so a host run checks the ALGEBRA and the synchronisation logic of a kernel (and
// every IEEE-pinned chain bit for bit), not the last bits of the GPU's approximations, not the memory model, not
// performance -- those stay the GPU tests' job.  Compile with -ffp-contract=off.
#pragma once

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __host__
#define __device__
#define __global__
#define __constant__
#define __shared__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// the executing thread's coordinates: set by the harness / the coroutine scheduler whenever a CUDA thread (re)starts
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---------------------------------------------------------------------------------------------- execution model
namespace hostcc {
void launch(dim3 grid, dim3 block, const std::function<void()>& body);      // blocks in sequence, threads as coroutines
void yield();                                                               // inside a spin loop: let the other threads run
void sync_block();
// warp rendezvous: every live lane named in `mask` deposits `v`; returns the deposits (index = lane) and who deposited
const unsigned long long* exchange(unsigned mask, unsigned long long v, unsigned* arrivedMask);
unsigned lane_id();
bool in_kernel();
}

// ---------------------------------------------------------------------------------------------- arithmetic intrinsics
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __saturatef(float a) { return a != a ? 0.0f : fminf(fmaxf(a, 0.0f), 1.0f); }
static inline int hostcc_f2i(float r) { return r != r ? 0 : r >= 2147483648.0f ? INT_MAX : r < -2147483648.0f ? INT_MIN : (int)r; }
static inline int __float2int_rn(float a) { return hostcc_f2i(nearbyintf(a)); }
static inline int __float2int_rz(float a) { return hostcc_f2i(truncf(a)); }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
static inline float __ll2float_rn(long long a) { return (float)a; }
static inline unsigned __float_as_uint(float a) { unsigned u; memcpy(&u, &a, 4); return u; }
static inline int __float_as_int(float a) { int u; memcpy(&u, &a, 4); return u; }
static inline float __uint_as_float(unsigned a) { float u; memcpy(&u, &a, 4); return u; }
static inline float __int_as_float(int a) { float u; memcpy(&u, &a, 4); return u; }
static inline double __longlong_as_double(long long a) { double u; memcpy(&u, &a, 8); return u; }
static inline int __popc(unsigned a) { return __builtin_popcount(a); }
static inline int __ffs(int a) { return __builtin_ffs(a); }
static inline int __clz(int a) { return a ? __builtin_clz((unsigned)a) : 32; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) { hostcc::yield(); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
using std::isnan;
using std::isinf;
using std::isfinite;

// ---------------------------------------------------------------------------------------------- block / warp operations
static inline void __syncthreads() { hostcc::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned a; hostcc::exchange(mask, 0, &a); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    unsigned a, r = 0; const unsigned long long* v = hostcc::exchange(mask, pred ? 1ull : 0ull, &a);
    for (int i = 0; i < 32; i++) if (((a >> i) & 1u) && v[i]) r |= 1u << i;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __all_sync(unsigned mask, int pred) { unsigned a; const unsigned long long* v = hostcc::exchange(mask, pred ? 1ull : 0ull, &a); for (int i = 0; i < 32; i++) if (((a >> i) & 1u) && !v[i]) return 0; return 1; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned x)
{ unsigned a, r = 0; const unsigned long long* v = hostcc::exchange(mask, x, &a); for (int i = 0; i < 32; i++) if ((a >> i) & 1u) r |= (unsigned)v[i]; return r; }
template <typename T> static inline unsigned long long hostcc_bits(T x) { unsigned long long u = 0; static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes"); memcpy(&u, &x, sizeof(T)); return u; }
template <typename T> static inline T hostcc_unbits(unsigned long long u) { T x; memcpy(&x, &u, sizeof(T)); return x; }
template <typename T> static inline T hostcc_shfl(unsigned mask, T x, int src)
{
    unsigned a; const unsigned long long* v = hostcc::exchange(mask, hostcc_bits(x), &a);
    return (src >= 0 && src < 32 && ((a >> src) & 1u)) ? hostcc_unbits<T>(v[src]) : x;        // inactive source: own value
}
template <typename T> static inline T __shfl_sync(unsigned mask, T x, int src) { return hostcc_shfl(mask, x, src & 31); }
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T x, int laneMask) { return hostcc_shfl(mask, x, (int)(hostcc::lane_id() ^ (unsigned)laneMask)); }
template <typename T> static inline T __shfl_up_sync(unsigned mask, T x, unsigned delta) { const int s = (int)hostcc::lane_id() - (int)delta; return hostcc_shfl(mask, x, s < 0 ? -1 : s); }
template <typename T> static inline T __shfl_down_sync(unsigned mask, T x, unsigned delta) { const int s = (int)hostcc::lane_id() + (int)delta; return hostcc_shfl(mask, x, s > 31 ? -1 : s); }

// one coroutine runs at a time: a plain read-modify-write is atomic
template <typename T> static inline T hostcc_rmw_add(T* p, T v) { const T o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return hostcc_rmw_add(p, v); }
static inline int atomicAdd(int* p, int v) { return hostcc_rmw_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return hostcc_rmw_add(p, v); }
static inline unsigned atomicMax(unsigned* p, unsigned v) { const unsigned o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { const unsigned o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { const int o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
static inline unsigned atomicAnd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o & v; return o; }
static inline unsigned atomicExch(unsigned* p, unsigned v) { const unsigned o = *p; *p = v; return o; }
static inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { const unsigned o = *p; if (o == c) *p = v; return o; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long c, unsigned long long v) { const unsigned long long o = *p; if (o == c) *p = v; return o; }

// ---------------------------------------------------------------------------------------------- runtime API (host memory)
typedef struct hostcc_stream* cudaStream_t;
typedef struct hostcc_event* cudaEvent_t;
typedef struct hostcc_graph* cudaGraph_t;
typedef struct hostcc_graph_exec* cudaGraphExec_t;
typedef struct hostcc_graph_node* cudaGraphNode_t;
enum cudaError_t { cudaSuccess = 0, cudaErrorNotSupported = 801 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum cudaGraphNodeType { cudaGraphNodeTypeKernel, cudaGraphNodeTypeMemcpy };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "ok" : "not supported by the host stand-in (tests/hostcc)"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; p->multiProcessorCount = 2; return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)aligned_alloc(256, (n + 255) & ~(size_t)255); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMallocAsync(T** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbolAsync(T& symbol, const void* s, size_t n, size_t offset, cudaMemcpyKind, cudaStream_t)
{ memcpy(reinterpret_cast<char*>(&symbol) + offset, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset2DAsync(void* d, size_t pitch, int v, size_t w, size_t h, cudaStream_t)
{ for (size_t y = 0; y < h; y++) memset((char*)d + y * pitch, v, w); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{ for (size_t y = 0; y < h; y++) memcpy((char*)d + y * dp, (const char*)s + y * sp, w); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
enum { cudaEventRecordExternal = 1, cudaEventWaitExternal = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventRecordWithFlags(cudaEvent_t, cudaStream_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
// multi-GPU plumbing of renderer.cu: every "device" of the stand-in is host memory of this process, so peer pointers are plain
// pointers (the in-process path of chord_renderer_exchange_connect); CUDA IPC between processes is not modelled
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaPointerAttributes { int device; };
static const cudaError_t cudaErrorPeerAccessAlreadyEnabled = (cudaError_t)704;
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->device = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
// graphs: not modelled (chord_graph_begin_capture reports the error; the graph tests need a GPU)
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphGetNodes(cudaGraph_t, cudaGraphNode_t*, size_t* n) { *n = 0; return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphNodeGetType(cudaGraphNode_t, cudaGraphNodeType*) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
