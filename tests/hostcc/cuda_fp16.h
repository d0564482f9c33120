// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): <cuda_fp16.h> stand-in on _Float16 (IEEE binary16,
// conversions round to nearest even like __float2half_rn / __half2float).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "cuda_runtime.h"

struct __half { _Float16 v; };
struct __half2 { __half x, y; };
static inline float __half2float(__half h) { return (float)h.v; }
static inline __half __float2half_rn(float f) { __half h; h.v = (_Float16)f; return h; }
static inline __half __float2half_rz(float f)       // round toward zero: step back when round-to-nearest moved away from zero
{
    __half h; h.v = (_Float16)f;
    if (f != f) return h;
    unsigned short u; memcpy(&u, &h.v, 2);
    if (std::fabs((float)h.v) > std::fabs(f)) { u = (unsigned short)(u - 1); memcpy(&h.v, &u, 2); }     // also inf -> largest finite
    return h;
}
static inline unsigned short __half_as_ushort(__half h) { unsigned short u; memcpy(&u, &h.v, 2); return u; }
static inline __half __ushort_as_half(unsigned short u) { __half h; memcpy(&h.v, &u, 2); return h; }
static inline float2 __half22float2(__half2 h) { return float2{(float)h.x.v, (float)h.y.v}; }
static inline float __low2float(__half2 h) { return (float)h.x.v; }
static inline float __high2float(__half2 h) { return (float)h.y.v; }
