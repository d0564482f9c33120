// TEST INFRASTRUCTURE ONLY -- not a CUDA runtime and not a CPU fallback of the library.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the DEVICE FUNCTIONS of chord_b200/csrc/*.cu as plain C++, so that
// tests/test_kernel_source_on_host.py can execute the very source text of a per-thread kernel (one without cross-thread
// communication) thread by thread on the CPU and hold it against the oracle when no GPU is at hand.  Nothing in the
// product links or loads what is built from this header: libchord_b200.so is nvcc output for sm_100a only and fails
// loudly without a device (tests/test_abi.py::test_no_cpu_fallback_without_gpu).
//
// Semantics: every *_rn intrinsic is the IEEE operation it names; the approximate ones (__fdividef, rsqrtf, MUFU
// sequences) are evaluated exactly here, so a host run checks the ALGEBRA of a kernel (and every IEEE-pinned chain bit for
// bit), not the last bits of the GPU's approximations -- those stay the GPU tests' job.  Compile with -ffp-contract=off.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
struct dim3 { unsigned x = 1, y = 1, z = 1; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// the executing thread's coordinates: the harness sets them before every call of a kernel body
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

typedef struct hostcc_stream* cudaStream_t;
typedef struct hostcc_event* cudaEvent_t;
typedef struct hostcc_graph* cudaGraph_t;
typedef struct hostcc_graph_exec* cudaGraphExec_t;
enum cudaError_t { cudaSuccess = 0 };
static inline const char* cudaGetErrorString(cudaError_t) { return "hostcc"; }

static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __saturatef(float a) { return a != a ? 0.0f : fminf(fmaxf(a, 0.0f), 1.0f); }
static inline int __float2int_rn(float a) { return (int)nearbyintf(a); }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
static inline unsigned __float_as_uint(float a) { unsigned u; memcpy(&u, &a, 4); return u; }
static inline int __float_as_int(float a) { int u; memcpy(&u, &a, 4); return u; }
static inline float __uint_as_float(unsigned a) { float u; memcpy(&u, &a, 4); return u; }
static inline float __int_as_float(int a) { float u; memcpy(&u, &a, 4); return u; }
static inline int __popc(unsigned a) { return __builtin_popcount(a); }
static inline int __ffs(int a) { return __builtin_ffs(a); }
static inline int __clz(int a) { return a ? __builtin_clz((unsigned)a) : 32; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
using std::isnan;

// cross-thread operations cannot be executed one thread at a time: a kernel that reaches one is not host-checkable
[[noreturn]] static inline void hostcc_collective(const char* what) { fprintf(stderr, "hostcc: %s needs a real warp\n", what); abort(); }
static inline unsigned __ballot_sync(unsigned, int) { hostcc_collective("__ballot_sync"); }
static inline int __any_sync(unsigned, int) { hostcc_collective("__any_sync"); }
static inline unsigned __reduce_or_sync(unsigned, unsigned) { hostcc_collective("__reduce_or_sync"); }
template <typename T> static inline T __shfl_sync(unsigned, T, int) { hostcc_collective("__shfl_sync"); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int) { hostcc_collective("__shfl_xor_sync"); }
static inline void __syncthreads() { hostcc_collective("__syncthreads"); }
static inline void __syncwarp(unsigned = 0xffffffffu) { hostcc_collective("__syncwarp"); }
// one thread at a time: plain read-modify-write
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
