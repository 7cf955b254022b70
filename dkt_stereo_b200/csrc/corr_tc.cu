// K1 / tensor-core variant: all-pairs 1-D correlation volume on tcgen05, pyramid pooled in the
// epilogue.  Per image row (b,y) this is C[w1][w2] = scale * sum_d f1[w1][d] * f2[w2][d]:
// both operands are K-major in the NHWC feature maps, so A = a 128-row w1 tile and B = the whole
// w2 extent (in blocks of <= 256 columns) are fetched by 3-D TMA boxes {64 ch, rows, 1} straight into the canonical
// 128-byte-swizzled UMMA layout.  Out-of-range rows / channels are zero-filled by the TMA unit.
// 3-term bf16 split (hi*hi + lo*hi + hi*lo) with fp32 accumulation in TMEM.
//
// The kernel is bound by the HBM write of the pyramid (470 MB at cfg2), not by the MMAs, so it is
// organised around the store stream: persistent CTAs (grid = #SMs) walk (row, w1-tile) work items;
// warp 0 = TMA producer (2-stage ring), warp 1 = MMA issuer into one of two TMEM accumulator
// stages, warps 2..9 = epilogue.  The epilogue of tile i overlaps the loads + MMAs of tile i+1.
// An epilogue warp pulls a 32-column chunk of its 32 rows from TMEM (thread = row), scales it,
// transposes it through a swizzled 4 KB smem buffer and continues with 8 lanes per row, so level 0
// leaves as full 128-byte lines; levels 1..3 are pooled in registers (+ one shuffle for level 3)
// from the same chunk, so lower levels are never re-read from HBM.
#include "common.cuh"
#include "tc.cuh"

namespace dkt {

using namespace tc;

constexpr int CT_THREADS = 320;
constexpr int CT_EPI_WARPS = 8;
constexpr int CT_MAX_STAGES = 4;
constexpr uint32_t CT_A_BYTES = 128 * 128;
constexpr uint32_t CT_EPI_BYTES = CT_EPI_WARPS * 32 * 128;

struct TcCorrParams {
    CUtensorMap f1[2];      // hi/lo 3-D (D, W1, B*H)
    CUtensorMap f2[2];      // hi/lo 3-D (D, W2, B*H)
    float* pyr[DKT_MAX_LEVELS];
    int pw[DKT_MAX_LEVELS];
    int levels;
    int W1, W2, Npad, kblocks, m_tiles, n_blocks, num_tiles;   // Npad = columns (w2) per tile; n_blocks tiles span W2
    int stages;
    uint32_t acc_cols;
    int vec;                // pw[0] % 8 == 0: every level's row start keeps its vector alignment
    float scale;
};

__global__ void __launch_bounds__(CT_THREADS, 1)
corr1d_build_tc_kernel(const __grid_constant__ TcCorrParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_bytes = (uint32_t)prm.Npad * 128u;
    const uint32_t stage_bytes = 2u * CT_A_BYTES + 2u * b_bytes;
    uint8_t* epi_smem = smem + (size_t)prm.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + CT_EPI_BYTES);
    uint64_t* empty_bar = full_bar + CT_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + CT_MAX_STAGES;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&prm.f1[0]); tma_prefetch_desc(&prm.f1[1]);
        tma_prefetch_desc(&prm.f2[0]); tma_prefetch_desc(&prm.f2[1]);
        for (int s = 0; s < prm.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], CT_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * prm.acc_cols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
                const int nbk = tile % prm.n_blocks, rm = tile / prm.n_blocks;
                const int row = rm / prm.m_tiles;            // b*H + y
                const int m0 = (rm - row * prm.m_tiles) * 128;
                const int n0 = nbk * prm.Npad;               // first w2 column of this tile
                for (int kb = 0; kb < prm.kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* st = smem + (size_t)stage * stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                    tma_load_3d(st, &prm.f1[0], &full_bar[stage], kb * 64, m0, row);
                    tma_load_3d(st + CT_A_BYTES, &prm.f1[1], &full_bar[stage], kb * 64, m0, row);
                    tma_load_3d(st + 2 * CT_A_BYTES, &prm.f2[0], &full_bar[stage], kb * 64, n0, row);
                    tma_load_3d(st + 2 * CT_A_BYTES + b_bytes, &prm.f2[1], &full_bar[stage], kb * 64, n0, row);
                    if (++stage == prm.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16_m128((uint32_t)prm.Npad);
            int stage = 0;
            uint32_t phase = 0, t = 0;
            for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, ++t) {
                const uint32_t as = t & 1u, aphase = (t >> 1) & 1u;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * prm.acc_cols;
                for (int kb = 0; kb < prm.kblocks; ++kb) {
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
                        umma_bf16(tmem_d, dah, dwh, idesc, (kb | k) != 0);
                        umma_bf16(tmem_d, dal, dwh, idesc, 1u);
                        umma_bf16(tmem_d, dah, dwl, idesc, 1u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == prm.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tmem_full_bar[as]);
            }
        }
    } else {
        // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4; two warps per quarter =====
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        float* ebuf = reinterpret_cast<float*>(epi_smem + (size_t)ew * 4096);
        const int sub = lane >> 3;           // row within a group of 4
        const int jg = lane & 7;             // 4-column group within the 32-column chunk
        const int pw0 = prm.pw[0], pw1 = prm.pw[1], pw2 = prm.pw[2], pw3 = prm.pw[3];
        const int levels = prm.levels;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, ++t) {
            const int nbk = tile % prm.n_blocks, rm = tile / prm.n_blocks;
            const int row = rm / prm.m_tiles;
            const int m0 = (rm - row * prm.m_tiles) * 128;
            const int n0 = nbk * prm.Npad;
            const uint32_t as = t & 1u, aphase = (t >> 1) & 1u;
            mbar_wait(&tmem_full_bar[as], aphase);
            tcgen05_fence_after();
            const uint32_t tbase = tmem_base + as * prm.acc_cols + ((uint32_t)(q * 32) << 16);
            for (int c0 = half * 32; c0 < prm.Npad; c0 += 64) {
                float v[32];
                const int ncols = (prm.Npad - c0 >= 32) ? 32 : 16;
                __syncwarp();
                if (ncols == 32) {
                    tmem_ld32(tbase + c0, v);
                } else {
                    tmem_ld16(tbase + c0, v);
#pragma unroll
                    for (int j = 16; j < 32; ++j) v[j] = 0.f;
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(ebuf + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_float4(v[4 * j] * prm.scale, v[4 * j + 1] * prm.scale, v[4 * j + 2] * prm.scale, v[4 * j + 3] * prm.scale);
                __syncwarp();
                const int c = n0 + c0 + 4 * jg;                         // first of this lane's 4 columns (w2)
                const bool colok = c0 + 4 * jg < prm.Npad;              // false: zero-filled tail of a 16-column chunk (another tile's columns)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + sub;
                    const float4 a = *reinterpret_cast<const float4*>(ebuf + r * 32 + ((jg ^ (r & 7)) << 2));
                    const int w1 = m0 + q * 32 + r;
                    const bool valid = w1 < prm.W1 && colok;
                    const int64_t prow = (int64_t)row * prm.W1 + w1;
                    // pooled values (same association as F.avg_pool2d chains: pairwise means)
                    const float l1a = (a.x + a.y) * 0.5f, l1b = (a.z + a.w) * 0.5f;
                    const float l2 = (l1a + l1b) * 0.5f;
                    const float l2n = __shfl_xor_sync(0xffffffffu, l2, 1);      // neighbour 4-column group
                    if (!valid) continue;
                    float* o0 = prm.pyr[0] + prow * pw0;
                    if (prm.vec && c + 3 < pw0) {
                        *reinterpret_cast<float4*>(o0 + c) = a;
                    } else {
                        if (c < pw0) o0[c] = a.x;
                        if (c + 1 < pw0) o0[c + 1] = a.y;
                        if (c + 2 < pw0) o0[c + 2] = a.z;
                        if (c + 3 < pw0) o0[c + 3] = a.w;
                    }
                    if (levels > 1) {
                        float* o1 = prm.pyr[1] + prow * pw1;
                        const int c1 = c >> 1;
                        if (prm.vec && c1 + 1 < pw1) {
                            *reinterpret_cast<float2*>(o1 + c1) = make_float2(l1a, l1b);
                        } else {
                            if (c1 < pw1) o1[c1] = l1a;
                            if (c1 + 1 < pw1) o1[c1 + 1] = l1b;
                        }
                    }
                    if (levels > 2) {
                        const int c2 = c >> 2;
                        if (c2 < pw2) prm.pyr[2][prow * pw2 + c2] = l2;
                    }
                    if (levels > 3 && !(jg & 1)) {
                        const int c3 = c >> 3;
                        if (c3 < pw3) prm.pyr[3][prow * pw3 + c3] = (l2 + l2n) * 0.5f;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * prm.acc_cols);
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
    if (W2 > 4096) return DKT_E_UNSUPPORTED;
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
    // one UMMA N extent is <= 256 columns: wider rows are cut into n_blocks equal column blocks (multiples of 16,
    // so every block starts on a level-3 pooling boundary); columns past W2 are zero-filled by TMA and never stored
    prm.n_blocks = ceil_div(W2, 256);
    prm.Npad = (ceil_div(W2, prm.n_blocks) + 15) / 16 * 16;
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
    const uint32_t budget = 227u * 1024u - 1024u - CT_EPI_BYTES - 256u;
    int stages = (int)(budget / stage_bytes);
    if (stages < 1) return DKT_E_UNSUPPORTED;
    if (stages > CT_MAX_STAGES) stages = CT_MAX_STAGES;
    prm.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)prm.Npad) cols <<= 1;
    prm.acc_cols = cols;
    prm.vec = (W2 % 8 == 0) ? 1 : 0;
    const size_t smem_bytes = (size_t)stages * stage_bytes + CT_EPI_BYTES + 1024 + 256;
    DKT_ENSURE_SMEM(227 * 1024, corr1d_build_tc_kernel);
    const int64_t tiles = (int64_t)B * H * prm.m_tiles * prm.n_blocks;
    if (tiles > 0x7fffffff) return DKT_E_UNSUPPORTED;
    prm.num_tiles = (int)tiles;
    const int s_sms = device_sms();
    const unsigned grid = (unsigned)(tiles < s_sms ? tiles : s_sms);
    corr1d_build_tc_kernel<<<grid, CT_THREADS, smem_bytes, (cudaStream_t)stream>>>(prm);
    DKT_RETURN_LAST();
}
