// Shared device/host helpers for the cnf_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "cnf_b200.h"

namespace cnf {

// ---------------------------------------------------------------------------------------------
// host side: error reporting across the C ABI (thread local text, integer codes)
// ---------------------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);

#define CNF_REQUIRE(cond, ...)                                         \
    do {                                                               \
        if (!(cond)) return ::cnf::fail(CNF_ERR_INVALID_ARG, __VA_ARGS__); \
    } while (0)

#define CNF_SUPPORTED(cond, ...)                                       \
    do {                                                               \
        if (!(cond)) return ::cnf::fail(CNF_ERR_UNSUPPORTED, __VA_ARGS__); \
    } while (0)

#define CNF_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ::cnf::fail(CNF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CNF_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return CNF_OK;
}

// Device-resident view of a coupling mask (built on the host from cnf_mask).
struct MaskView {
    uint64_t cond_c;   // bit c set  -> channel c is a conditioner input
    uint64_t cond_s;   // bit (s % s_period) set -> position is a conditioner input
    int s_period;      // 0 = no position mask
    int n_t;           // number of transformed channels
    int c0;            // first transformed channel when they form one run
    int contiguous;    // transformed channels are c0 .. c0+n_t-1
    unsigned char tch[CNF_MAX_CHANNELS];  // transformed channel ids, ascending
};

int build_mask(const cnf_mask& m, int C, MaskView* out);

int sm_count();

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLog1em22 = -50.656872045869f;  // log(1e-22), the reference's safe_log clamp
constexpr float kInvLn10 = 0.4342944819032518f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_exp(float x) { return ex2(x * kLog2e); }
__device__ __forceinline__ float fast_log(float x) { return lg2(x) * kLn2; }

// Packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 issue two IEEE fp32 operations per instruction - same rounding as the
// scalar forms, half the issue slots).  A pair lives in one 64-bit register pair.
struct f2 {
    unsigned long long v;
};
__device__ __forceinline__ f2 f2_make(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 f2_splat(float x) { return f2_make(x, x); }
__device__ __forceinline__ void f2_get(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}

// tanh(v) with v already multiplied by 2*log2(e): 1 - 2/(1 + 2^{v2}).  Absolute error ~1.2e-7.
__device__ __forceinline__ float tanh_from_2log2e(float v2) { return fmaf(-2.0f, rcp(1.0f + ex2(v2)), 1.0f); }

// softplus(v) + softplus(-v) = |v| + 2 log(1 + e^{-|v|})
__device__ __forceinline__ float softplus_pm(float v) {
    float a = fabsf(v);
    return fmaf(2.0f * kLn2, lg2(1.0f + ex2(-a * kLog2e)), a);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// streaming (read-once / write-once) global accesses
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// exact i / d for 0 <= i < 2^21 via a float reciprocal (inv = 1.0f / d)
__device__ __forceinline__ int fast_div(int i, float inv) { return __float2int_rz(((float)i + 0.5f) * inv); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Sum `v` over runs of equal `key` (keys non-decreasing across the lanes of a warp) and let the
// first lane of each run add the run total to dst[key].
__device__ __forceinline__ void warp_segmented_atomic_add(float* dst, long long key, float v, bool valid) {
    const unsigned lane = threadIdx.x & 31u;
    if (!valid) { key = -1; v = 0.0f; }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float ov = __shfl_down_sync(0xffffffffu, v, d);
        long long ok = __shfl_down_sync(0xffffffffu, key, d);
        if (lane + d < 32 && ok == key) v += ov;
    }
    long long prev = __shfl_up_sync(0xffffffffu, key, 1);
    if (valid && (lane == 0 || prev != key)) atomicAdd(dst + key, v);
}

__device__ __forceinline__ void flag(uint32_t* status, uint32_t bits) {
    if (status != nullptr && bits != 0u) atomicOr(status, bits);
}

}  // namespace cnf
