// cuda_runtime.h — HOST SHIM for tests/devcore only (never on the product's include path).
//
// Lets g++ compile the per-thread device functions of msdfgen_b200/csrc/distance_core.cuh unchanged, as ordinary C++: the CUDA
// qualifiers vanish, the handful of intrinsics the file uses get host definitions, and "the current thread" is thread 0 of a
// one-thread block.  Only code without warp- or block-level cooperation can be compiled this way; that is exactly the texel
// semantics (segment kernels, root solvers, selectors, combiners), which is what the CPU test checks against the oracle.
#pragma once

#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__ static inline     // kernels are only parsed here, never emitted
#define __launch_bounds__(...)
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ thread_local       // block-scope: static storage; `extern __shared__ T x[]` stays a valid declaration
#define __restrict__

struct double2 { double x, y; };
struct DevcoreIdx { unsigned x, y, z; };
static const DevcoreIdx threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 }, blockDim = { 1, 1, 1 }, gridDim = { 1, 1, 1 };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r = { x, y }; return r; }

// fp32 helpers of the fast foot-point path.  The device versions are approximate (__cosf: absolute error up to 2^-21.4 on
// [-pi, pi], CUDA C Programming Guide) or plain IEEE (__frcp_rn); their values only seed a double-precision polish.  The stand-in
// for __cosf is the host's cosf pushed by b200_shim_cos_bias * 1e-6 (three times the instruction's error bound) in either
// direction, so that the tests cover what the hardware may return.
extern int b200_shim_cos_bias;
static inline float b200_shim_cosf(float x) { return cosf(x)+b200_shim_cos_bias*1e-6f; }
#define __cosf(x) b200_shim_cosf(x)     // (glibc declares an internal function of this name: a macro, not a definition)
static inline float __frcp_rn(float x) { return 1.0f/x; }
using std::isfinite;

template <typename T> static inline T atomicAdd(T *p, T v) { T old = *p; *p += v; return old; }
template <typename T> static inline T atomicOr(T *p, T v) { T old = *p; *p |= v; return old; }
template <typename T> static inline T atomicMax(T *p, T v) { T old = *p; if (v > old) *p = v; return old; }
// what only the kernels' bodies (parsed, not emitted) mention
static inline void __syncthreads() { }
static inline void __syncwarp(unsigned = 0xffffffffu) { }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
template <typename T> static inline T atomicMin(T *p, T v) { T old = *p; if (v < old) *p = v; return old; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t) p; }

// ---- the fp32 filter (distance_filter.cuh) -----------------------------------------------------------------------------------------
#define B200_HOST_SHIM 1
struct float4 { float x, y, z, w; };
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __saturatef(float x) { return x >= 0.f ? (x <= 1.f ? x : 1.f) : 0.f; }       // NaN -> 0 like the instruction
static inline int __float_as_int(float x) { int r; memcpy(&r, &x, 4); return r; }
static inline unsigned __float_as_uint(float x) { unsigned r; memcpy(&r, &x, 4); return r; }
static inline float __int_as_float(int x) { float r; memcpy(&r, &x, 4); return r; }
static inline float __uint_as_float(unsigned x) { float r; memcpy(&r, &x, 4); return r; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int) x); }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
// round-up conversion: the nearest float, stepped up if it fell below
static inline float __double2float_ru(double x) {
    float f = (float) x;
    return (double) f < x ? nextafterf(f, INFINITY) : f;
}
// A "warp" is one lane here, so a warp-wide OR is the lane's own value: exact_texel walks the union of its lanes' survivors and
// every lane acts only on its own bits, hence a lane's result does not depend on its neighbours.
static inline unsigned __reduce_or_sync(unsigned, unsigned x) { return x; }
// sqrt.approx.ftz.f32 has a maximum relative error of 2^-23 (PTX ISA); the stand-in is the correctly rounded root pushed by
// b200_shim_sqrt_bias * 2^-22 in either direction, so the tests cover twice the instruction's error on both sides.
extern int b200_shim_sqrt_bias;
static inline float b200_shim_sqrt_approx(float x) {
    const float r = sqrtf(x);
    return b200_shim_sqrt_bias ? r*(1.f+b200_shim_sqrt_bias*2.384185791015625e-7f) : r;
}
