// K1 / tensor-core variant: all-pairs 1-D correlation volume on tcgen05, pyramid pooled in the
// epilogue.  Per image row (b,y) this is C[w1][w2] = scale * sum_d f1[w1][d] * f2[w2][d]:
// both operands are K-major in the NHWC feature maps, so A = a 128-row w1 tile and B = the whole
// w2 extent (<= 256) are fetched by 3-D TMA boxes {64 ch, rows, 1} straight into the canonical
// 128-byte-swizzled UMMA layout.  Out-of-range rows / channels are zero-filled by the TMA unit.
// 3-term bf16 split (hi*hi + lo*hi + hi*lo) with fp32 accumulation in TMEM.
//
// The kernel is bound by the HBM write of the 470 MB pyramid (cfg2), not by the MMAs: each
// thread of the 4 epilogue warps owns one w1 row of the accumulator, scales it, streams level 0
// out with 16-byte stores and pools levels 1..3 in registers, so lower levels are never re-read.
// Two CTAs are resident per SM (one smem stage each, 256 TMEM columns each) so that one CTA's
// store phase overlaps the other's load + MMA phase.
#include "common.cuh"
#include "tc.cuh"

namespace dkt {

using namespace tc;

constexpr int CT_THREADS = 192;
constexpr uint32_t CT_A_BYTES = 128 * 128;

struct TcCorrParams {
    CUtensorMap f1[2];      // hi/lo 3-D (D, W1, B*H)
    CUtensorMap f2[2];      // hi/lo 3-D (D, W2, B*H)
    float* pyr[DKT_MAX_LEVELS];
    int pw[DKT_MAX_LEVELS];
    int levels;
    int W1, W2, Npad, kblocks, m_tiles;
    int stages;
    uint32_t tmem_cols;
    float scale;
};

__global__ void __launch_bounds__(CT_THREADS, 2)
corr1d_build_tc_kernel(const __grid_constant__ TcCorrParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_bytes = (uint32_t)prm.Npad * 128u;
    const uint32_t stage_bytes = 2u * CT_A_BYTES + 2u * b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)prm.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + 4;
    uint64_t* tmem_full_bar = empty_bar + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x / prm.m_tiles;          // b*H + y
    const int m0 = (blockIdx.x % prm.m_tiles) * 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&prm.f1[0]); tma_prefetch_desc(&prm.f1[1]);
        tma_prefetch_desc(&prm.f2[0]); tma_prefetch_desc(&prm.f2[1]);
        for (int s = 0; s < prm.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, prm.tmem_cols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < prm.kblocks; ++kb) {
                const int stage = kb % prm.stages;
                const uint32_t phase = (uint32_t)(kb / prm.stages) & 1u;
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                uint8_t* st = smem + (size_t)stage * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                tma_load_3d(st, &prm.f1[0], &full_bar[stage], kb * 64, m0, row);
                tma_load_3d(st + CT_A_BYTES, &prm.f1[1], &full_bar[stage], kb * 64, m0, row);
                tma_load_3d(st + 2 * CT_A_BYTES, &prm.f2[0], &full_bar[stage], kb * 64, 0, row);
                tma_load_3d(st + 2 * CT_A_BYTES + b_bytes, &prm.f2[1], &full_bar[stage], kb * 64, 0, row);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16_m128((uint32_t)prm.Npad);
            for (int kb = 0; kb < prm.kblocks; ++kb) {
                const int stage = kb % prm.stages;
                const uint32_t phase = (uint32_t)(kb / prm.stages) & 1u;
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t a_lo = a_hi + CT_A_BYTES;
                const uint32_t w_hi = a_hi + 2 * CT_A_BYTES;
                const uint32_t w_lo = w_hi + b_bytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t dah = smem_desc_sw128(a_hi + k * 32), dal = smem_desc_sw128(a_lo + k * 32);
                    const uint64_t dwh = smem_desc_sw128(w_hi + k * 32), dwl = smem_desc_sw128(w_lo + k * 32);
                    umma_bf16(tmem_base, dah, dwh, idesc, (kb | k) != 0);
                    umma_bf16(tmem_base, dal, dwh, idesc, 1u);
                    umma_bf16(tmem_base, dah, dwl, idesc, 1u);
                }
                umma_commit(&empty_bar[stage]);
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
        const int q = warp & 3;
        const int w1 = m0 + q * 32 + lane;
        const bool valid = w1 < prm.W1;
        const int64_t prow = (int64_t)row * prm.W1 + w1;
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
        float* o0 = prm.pyr[0] + prow * prm.pw[0];
        float* o1 = prm.levels > 1 ? prm.pyr[1] + prow * prm.pw[1] : nullptr;
        float* o2 = prm.levels > 2 ? prm.pyr[2] + prow * prm.pw[2] : nullptr;
        float* o3 = prm.levels > 3 ? prm.pyr[3] + prow * prm.pw[3] : nullptr;
        const bool vec0 = (prm.pw[0] % 4) == 0;
        const bool vec1 = o1 && (prm.pw[1] % 4) == 0;
        const bool vec2 = o2 && (prm.pw[2] % 4) == 0;
        const bool vec3 = o3 && (prm.pw[3] % 4) == 0;
        for (int c0 = 0; c0 < prm.Npad; c0 += 32) {
            float v[32];
            const int ncols = (prm.Npad - c0 >= 32) ? 32 : 16;
            __syncwarp();                       // tcgen05.ld is .sync.aligned: reconverge first
            if (ncols == 32) tmem_ld32(tbase + c0, v); else tmem_ld16(tbase + c0, v);
            tmem_ld_wait();
            if (valid) do {
            if (ncols == 16) {
#pragma unroll
                for (int j = 16; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= prm.scale;
            // level 0
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int c = c0 + j;
                if (vec0 && c + 3 < prm.pw[0]) {
                    *reinterpret_cast<float4*>(o0 + c) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) if (c + t < prm.pw[0]) o0[c + t] = v[j + t];
                }
            }
            if (!o1) break;
            float l1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) l1[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const int c = c0 / 2 + j;
                if (vec1 && c + 3 < prm.pw[1]) {
                    *reinterpret_cast<float4*>(o1 + c) = make_float4(l1[j], l1[j + 1], l1[j + 2], l1[j + 3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) if (c + t < prm.pw[1]) o1[c + t] = l1[j + t];
                }
            }
            if (!o2) break;
            float l2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) l2[j] = (l1[2 * j] + l1[2 * j + 1]) * 0.5f;
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                const int c = c0 / 4 + j;
                if (vec2 && c + 3 < prm.pw[2]) {
                    *reinterpret_cast<float4*>(o2 + c) = make_float4(l2[j], l2[j + 1], l2[j + 2], l2[j + 3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) if (c + t < prm.pw[2]) o2[c + t] = l2[j + t];
                }
            }
            if (!o3) break;
            float l3[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) l3[j] = (l2[2 * j] + l2[2 * j + 1]) * 0.5f;
            {
                const int c = c0 / 8;
                if (vec3 && c + 3 < prm.pw[3]) {
                    *reinterpret_cast<float4*>(o3 + c) = make_float4(l3[0], l3[1], l3[2], l3[3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) if (c + t < prm.pw[3]) o3[c + t] = l3[t];
                }
            }
            } while (0);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, prm.tmem_cols);
}

}  // namespace dkt

using namespace dkt;

extern "C" int dkt_corr1d_build_tc(const uint16_t* f1_hi, const uint16_t* f1_lo,
                                   const uint16_t* f2_hi, const uint16_t* f2_lo,
                                   float* const* pyr, int B, int D, int H, int W1, int W2,
                                   int levels, float scale, void* stream) {
    DKT_CHECK_ARG(f1_hi && f1_lo && f2_hi && f2_lo && pyr);
    DKT_CHECK_ARG(B > 0 && D > 0 && H > 0 && W1 > 0 && W2 > 0);
    if (levels < 1 || levels > DKT_MAX_LEVELS) return DKT_E_UNSUPPORTED;
    if (W2 > 256) return DKT_E_UNSUPPORTED;          // one UMMA N extent; wider rows use the fp32 kernel
    if (D % 8) return DKT_E_ALIGNMENT;
    TcCorrParams prm{};
    int w = W2;
    for (int l = 0; l < DKT_MAX_LEVELS; ++l) {
        prm.pyr[l] = l < levels ? pyr[l] : nullptr;
        prm.pw[l] = w;
        if (l < levels) {
            DKT_CHECK_ARG(pyr[l] != nullptr && w > 0);
            if (reinterpret_cast<uintptr_t>(pyr[l]) & 15) return DKT_E_ALIGNMENT;
        }
        w /= 2;
    }
    prm.levels = levels;
    prm.W1 = W1;
    prm.W2 = W2;
    prm.Npad = (W2 + 15) / 16 * 16;
    prm.kblocks = ceil_div(D, 64);
    prm.m_tiles = ceil_div(W1, 128);
    prm.scale = scale;
    {
        const uint64_t d1[3] = {(uint64_t)D, (uint64_t)W1, (uint64_t)B * H};
        const uint64_t s1[3] = {1, (uint64_t)D, (uint64_t)D * W1};
        const uint32_t b1[3] = {64, 128, 1};
        if (!make_tmap_bf16(&prm.f1[0], f1_hi, 3, d1, s1, b1)) return DKT_E_DRIVER;
        if (!make_tmap_bf16(&prm.f1[1], f1_lo, 3, d1, s1, b1)) return DKT_E_DRIVER;
        const uint64_t d2[3] = {(uint64_t)D, (uint64_t)W2, (uint64_t)B * H};
        const uint64_t s2[3] = {1, (uint64_t)D, (uint64_t)D * W2};
        const uint32_t b2[3] = {64, (uint32_t)prm.Npad, 1};
        if (!make_tmap_bf16(&prm.f2[0], f2_hi, 3, d2, s2, b2)) return DKT_E_DRIVER;
        if (!make_tmap_bf16(&prm.f2[1], f2_lo, 3, d2, s2, b2)) return DKT_E_DRIVER;
    }
    const uint32_t stage_bytes = 2u * CT_A_BYTES + 2u * (uint32_t)prm.Npad * 128u;
    // aim for two resident CTAs per SM (each <= ~110 KB of shared memory)
    int stages = (int)((108u * 1024u) / stage_bytes);
    if (stages < 1) stages = 1;
    if (stages > 4) stages = 4;
    if (stages > prm.kblocks) stages = prm.kblocks;
    prm.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)prm.Npad) cols <<= 1;
    prm.tmem_cols = cols;
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ce = cudaFuncSetAttribute(corr1d_build_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (ce != cudaSuccess) return (int)ce;
        attr_set = true;
    }
    const int64_t ctas = (int64_t)B * H * prm.m_tiles;
    if (ctas > 0x7fffffff) return DKT_E_UNSUPPORTED;
    corr1d_build_tc_kernel<<<(unsigned)ctas, CT_THREADS, smem_bytes, (cudaStream_t)stream>>>(prm);
    DKT_RETURN_LAST();
}
