// Shared device/host helpers for libdkt_stereo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include "../../include/dkt_stereo_b200.h"

#define DKT_CHECK_ARG(cond)  do { if (!(cond)) return DKT_E_INVALID; } while (0)
#define DKT_RETURN_LAST()    do { cudaError_t e__ = cudaGetLastError(); return e__ == cudaSuccess ? 0 : (int)e__; } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember per (call site, device)
// whether it has been raised.  One process may drive several GPUs (nn.DataParallel in the reference's evaluator) from
// several host threads; the bit mask is atomic and setting the attribute twice is harmless.
#define DKT_ENSURE_SMEM(bytes, ...)                                                                              \
    do {                                                                                                         \
        static std::atomic<uint64_t> done__{0};                                                                  \
        int dev__ = 0;                                                                                           \
        if (cudaGetDevice(&dev__) != cudaSuccess) { cudaGetLastError(); dev__ = 0; }                             \
        const uint64_t bit__ = 1ull << (dev__ & 63);                                                             \
        if (dev__ > 63 || !(done__.load(std::memory_order_acquire) & bit__)) {                                   \
            cudaError_t ce__ = cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
            if (ce__ != cudaSuccess) return (int)ce__;                                                           \
            done__.fetch_or(bit__, std::memory_order_release);                                                   \
        }                                                                                                        \
    } while (0)

namespace dkt {

constexpr int kNumSMs = 148;

// SM count of the CURRENT device (cached per device ordinal; 148 if the query fails)
inline int device_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return kNumSMs; }
    if (dev >= 0 && dev < 64) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = kNumSMs;
    }
    if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- 16-bit (hi, lo) operand split: hi = rn16(x), lo = rn16(x - hi) ----------------------------------------------
// DKT_SPLIT_FP16 = 1 (default): IEEE half.  hi alone carries 11 significant bits -- enough, with (hi, lo) WEIGHTS, for
// the convs that run 2 MMAs per K step (profiles/r2_precision_study_*.txt); hi + lo carries 22 bits (abs. floor
// 2^-25 from half's subnormals), range +-65504 (the range the reference's own --mixed_precision autocast mode lives in).
// DKT_SPLIT_FP16 = 0: bfloat16 as in round 1 (8 / 16 significant bits, fp32 range); every conv then needs 3 MMAs.
#ifndef DKT_SPLIT_FP16
#define DKT_SPLIT_FP16 1
#endif

#if DKT_SPLIT_FP16
__device__ __forceinline__ void split16(float x, uint16_t& hi, uint16_t& lo) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// pack two consecutive channels (one packed F2FP conversion per pair instead of two F2F)
__device__ __forceinline__ void split16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);                     // .x = a (low half), .y = b
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// hi word only (destinations that feed 2-MMA convs carry no lo plane)
__device__ __forceinline__ uint32_t pack_hi16x2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// two packed 16-bit values (low, high half of w) -> fp32
__device__ __forceinline__ float2 unpack16x2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ float unpack16(uint16_t v) { return __half2float(__ushort_as_half(v)); }
#else
__device__ __forceinline__ void split16(float x, uint16_t& hi, uint16_t& lo) {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    float r = x - __bfloat162float(h);
    __nv_bfloat16 l = __float2bfloat16_rn(r);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
}

__device__ __forceinline__ void split16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);          // .x = a (low half), .y = b
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ uint32_t pack_hi16x2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float2 unpack16x2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ float unpack16(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }
#endif

// 4 consecutive channels of a 16-bit (hi, lo) pair -> fp32 (hi + lo); read-only (non-coherent) path
__device__ __forceinline__ float4 load_split4(const uint16_t* hi, const uint16_t* lo, int64_t off) {
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + off));
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(lo + off));
    const float2 h0 = unpack16x2(h.x), h1 = unpack16x2(h.y), l0 = unpack16x2(l.x), l1 = unpack16x2(l.y);
    return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- activations.  expf / tanhf (not the __ intrinsics): the recurrence is sensitive ----
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Gate activations of the tensor-core GRU epilogues (-DDKT_FAST_GATES=0: expf / tanhf / IEEE division as above).
// ncu r4d: with expf + a full division per gate value the epilogue warps of the 1/8- and 1/16-resolution GRUs (one MMA
// per K16 step) spent a quarter of their samples in sigmoidf_acc and the MMA warp waited for accumulators.  Here:
// e^x = ex2.approx(t) * (1 + r ln 2) with t + r = x log2(e) carried as a rounded product and its FMA remainder, so the
// argument error does not grow with |x| (ex2.approx: 2 ulp); 1 / d = rcp.approx + one Newton step (< 1 ulp).  Absolute
// error of either gate <= ~1.5e-7 (arguments clamped where the fp32 result is saturated anyway).
#ifndef DKT_FAST_GATES
#define DKT_FAST_GATES 1
#endif
__device__ __forceinline__ float exp_fast(float x) {
    const float t = x * 1.44269502162933349609375f;
    float r = fmaf(x, 1.44269502162933349609375f, -t);
    r = fmaf(x, 1.925963033500011e-8f, r);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return fmaf(e, r * 0.693147182464599609375f, e);
}
__device__ __forceinline__ float rcp_fast(float d) {          // d in [1, 2^45]
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fmaf(r, fmaf(-d, r, 1.0f), r);
}
__device__ __forceinline__ float sigmoidf_gate(float x) {
#if DKT_FAST_GATES
    x = fminf(fmaxf(x, -30.0f), 30.0f);
    return rcp_fast(1.0f + exp_fast(-x));
#else
    return sigmoidf_acc(x);
#endif
}
__device__ __forceinline__ float tanhf_gate(float x) {
#if DKT_FAST_GATES
    x = fminf(fmaxf(x, -15.0f), 15.0f);
    return fmaf(-2.0f, rcp_fast(1.0f + exp_fast(2.0f * x)), 1.0f);
#else
    return tanhf(x);
#endif
}

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case DKT_ACT_RELU:    return fmaxf(x, 0.0f);
        case DKT_ACT_SIGMOID: return sigmoidf_acc(x);
        case DKT_ACT_TANH:    return tanhf(x);
        case DKT_ACT_LEAKY:   return x > 0.f ? x : 0.01f * x;
        default:              return x;
    }
}

// Write one value to every non-null precision of an NHWC tensor slice.
__device__ __forceinline__ void store_all(const dkt_tensor& t, int64_t pixel, int c, float v) {
    int64_t off = pixel * t.C + t.c_begin + c;
    if (t.f32) t.f32[off] = v;
    if (t.hi) {
        uint16_t h, l;
        split16(v, h, l);
        t.hi[off] = h;
        if (t.lo) t.lo[off] = l;
    }
}

// 4 consecutive channels (c % 4 == 0 and slice 4-aligned by construction of the callers)
__device__ __forceinline__ void store_all4(const dkt_tensor& t, int64_t pixel, int c, float4 v) {
    int64_t off = pixel * t.C + t.c_begin + c;
    if (t.f32) *reinterpret_cast<float4*>(t.f32 + off) = v;
    if (t.hi) {
        if (t.lo) {
            uint32_t h0, l0, h1, l1;
            split16x2(v.x, v.y, h0, l0);
            split16x2(v.z, v.w, h1, l1);
            *reinterpret_cast<uint2*>(t.hi + off) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(t.lo + off) = make_uint2(l0, l1);
        } else {
            *reinterpret_cast<uint2*>(t.hi + off) = make_uint2(pack_hi16x2(v.x, v.y), pack_hi16x2(v.z, v.w));
        }
    }
}

// ---- K2: one row sample = 2r+2 adjacent volume entries around x (zero outside the row) + the fractional weight.
// Split from the interpolation so that a thread can have the loads of several samples in flight before it consumes any.
template <int R>
__device__ __forceinline__ void sample_row_load(const float* __restrict__ row, int W, float x, float* v, float& a) {
    const float xf = floorf(x);
    a = x - xf;
    const int i0 = (int)xf - R;
#pragma unroll
    for (int k = 0; k < 2 * R + 2; ++k) {
        int idx = i0 + k;
        v[k] = (idx >= 0 && idx < W) ? __ldg(row + idx) : 0.f;
    }
}

}  // namespace dkt
