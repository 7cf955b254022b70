// K2 on tensor cores: the indexed multi-level lookup fused with the motion encoder's 1x1 `convc1` + ReLU, with the
// 36 -> 64 (RAFT-Stereo) / 162 -> 64 (IGEV-Stereo) contraction on tcgen05 instead of fp32 FMAs.
//
// Why (profiles/r03m_lookup_phases.txt, profiles/r2g_corr_full.csv): in the CUDA-core kernel (corr.cu) the gather and the
// exact-fp32 encode each take about half of the run time and do not overlap -- both live on the shared-memory pipe
// (scattered 4-byte tap stores, two 16-byte loads per 8 FFMA2).  Here the gather threads write the taps as 16-bit
// (hi, lo) values straight into a K-major SWIZZLE_128B operand tile, one elected thread issues a handful of
// tcgen05.mma (M = 128 pixels, N = 64, K = 48 / 192), and the epilogue pulls the 64 results of a pixel from TMEM.
// The encode leaves the shared-memory pipe and the FMA pipe; what remains is the gather, which is bound by the DRAM
// traffic of its 40-byte tap runs (each costs a 128-byte line).
//
//   CTA            128 pixels per chunk, 256 threads, persistent over chunks
//   shared memory  A[NBUF][AP planes][KB][128 rows x 128 B]   taps, K-major, 128-byte swizzle, written by the gather
//                  W[2 planes][KB][64 rows x 128 B]            convc1, pre-swizzled on the host (ops.pack_lookup_tc)
//   TMEM           128 columns: [0, 64) = taps_hi * w_hi, [64, 128) = the lo products (added in the epilogue in fp32
//                  round-to-nearest: tcgen05 accumulates with round-toward-zero, see conv_tc.cu acc_lo_off)
//   schedule       NBUF = 2: the gather of chunk i+1 runs while the MMAs of chunk i are in flight;
//                  NBUF = 1 (IGEV, 48 KB per tile): two CTAs per SM overlap each other instead.
//   AP             2: taps as (hi, lo) pairs, 3 MMAs per K step; 1: hi only, 2 MMAs (engine policy, update.py menc2)
//
// IGEV reads the geometry volume in the (B,H,W,D,C) layout of dkt_geo_pool_dc: the 10 taps x 8 channels of a
// (pixel, level) are ONE 320-byte run instead of 8 runs of 40 bytes in 8 different 192-byte rows.
#include "common.cuh"
#include "tc.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace dkt {

using namespace tc;

constexpr int LT_M = 128;            // pixels per chunk
constexpr int LT_N = 64;             // convc1 output channels
constexpr int LT_THREADS = 256;
constexpr uint32_t LT_A_KB_BYTES = LT_M * 128;     // one 64-wide K block of an A plane
constexpr uint32_t LT_W_KB_BYTES = LT_N * 128;

struct LookupTcParams {
    // RAFT: pyramid levels; IGEV: geo[l] (B,H,W,D>>l,Cg) and init[l] (B,H,W,W>>l)
    const float* vol[DKT_MAX_LEVELS];
    int          vw[DKT_MAX_LEVELS];
    const float* geo[2];
    int levels, Cg, D, W1;
    float* coords;                   // RAFT: coords_x; IGEV: disparity
    const float* delta;
    int delta_C;
    float* flow;
    const uint16_t* w_img;           // [2][KB][64 x 64] 16-bit, already in the swizzled shared-memory order
    const float* bias;
    dkt_tensor out;                  // 64-channel NHWC slice
    int64_t P;
    int C;                           // K slots in use: row samples per pixel * LT_TS
    int flags;                       // bit 0: RAFT gather through aligned 16-byte windows (else one thread per row sample);
                                     // bit 1: 16-bit output planes through the staging tile (else per-thread stores)
};

// byte offset of element (row m, column k) inside one plane [KB][rows][128 B] of a K-major SWIZZLE_128B tile whose
// base is 1024-byte aligned: 8-row atoms of 1024 bytes, 16-byte chunks XOR-ed with the row index inside the atom
__device__ __forceinline__ uint32_t sw128_off(int m, int k, uint32_t kb_bytes) {
    return (uint32_t)(k >> 6) * kb_bytes + (uint32_t)(m >> 3) * 1024u + (uint32_t)(m & 7) * 128u +
           (uint32_t)((((k & 63) >> 3) ^ (m & 7)) << 4) + (uint32_t)(k & 7) * 2u;
}

// K slots per row sample: its 2r+1 = 9 taps + one zero.  Every sample then starts on an even k, so its taps leave as
// five packed 32-bit stores (a pair never straddles a 16-byte swizzle chunk), and a row's byte offset is simply
// ((k >> 6) * 16 KB + (k & 63) * 2) ^ ((m & 7) << 4).
constexpr int LT_TS = 10;
// staging buffer shared by (a) the RAFT gather's tap windows, [256 samples][LT_WS floats] (a sample's 10 taps lie in an
// ALIGNED 16-float window that four lanes fetch with one 16-byte load each: a request then touches 8 lines instead of
// 32, the L1 wavefronts that bounded the one-thread-per-sample gather, ncu r2j) and (b) the epilogue's output tile,
// [128 pixels][64 x 16 bit], written per pixel and copied out as whole 128-byte rows.
constexpr int LT_WS = 18;                        // window row stride in floats (18: 8-byte aligned, 2-way bank conflicts)
constexpr uint32_t LT_STAGE_BYTES = 256 * LT_WS * 4;          // 18432 >= 128 * 128
constexpr int LT_DEFAULT_FLAGS = 2;              // staged epilogue, one-thread-per-sample gather (profiles/r2o_lookup_tc_sweep.txt)

template <int AP>
__device__ __forceinline__ void put_taps(uint8_t* a_buf, uint32_t plane_bytes, int m, int k0, const float* v, float a) {
    uint8_t* row = a_buf + (uint32_t)(m >> 3) * 1024u + (uint32_t)(m & 7) * 128u;
    const uint32_t sw = (uint32_t)(m & 7) << 4;
#pragma unroll
    for (int t = 0; t < LT_TS; t += 2) {
        const float t0 = (1.f - a) * v[t] + a * v[t + 1];
        const float t1 = (t + 1 < LT_TS - 1) ? (1.f - a) * v[t + 1] + a * v[t + 2] : 0.f;
        const int k = k0 + t;
        const uint32_t off = (((uint32_t)(k >> 6) * LT_A_KB_BYTES) + (uint32_t)(k & 63) * 2u) ^ sw;
        if (AP == 2) {
            uint32_t hi, lo;
            split16x2(t0, t1, hi, lo);
            *reinterpret_cast<uint32_t*>(row + off) = hi;
            *reinterpret_cast<uint32_t*>(row + plane_bytes + off) = lo;
        } else {
            *reinterpret_cast<uint32_t*>(row + off) = pack_hi16x2(t0, t1);
        }
    }
}

template <int R, int KB, int AP, int NBUF, bool GEO>
__global__ void __launch_bounds__(LT_THREADS, GEO ? 2 : 3)
lookup_tc_kernel(const __grid_constant__ LookupTcParams prm) {
    static_assert(2 * R + 2 == LT_TS, "tap slots");
    constexpr uint32_t A_PLANE = KB * LT_A_KB_BYTES;
    constexpr uint32_t A_BUF = AP * A_PLANE;
    constexpr uint32_t W_PLANE = KB * LT_W_KB_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;                                   // [NBUF][AP][KB][128 x 128 B]
    uint8_t* w_tile = a_ring + NBUF * A_BUF;                  // [2][KB][64 x 128 B]
    // staging tile: RAFT has room for its own; IGEV (2 x 48 KB of operands) lets it alias the first K block of the tap
    // tile, which is dead between the MMAs' completion and the next gather and holds tap slots only (every column is
    // rewritten by the next gather, so no stale value can meet a zero weight as inf x 0)
    uint8_t* stage = GEO ? a_ring : w_tile + 2 * W_PLANE;
    float* s_x = reinterpret_cast<float*>(w_tile + 2 * W_PLANE + (GEO ? 0 : LT_STAGE_BYTES));   // [128]
    float* s_bias = s_x + LT_M;                               // [64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_bias + LT_N);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    // zero the operand tiles once: the K padding [C, 64 * KB) is never written again and must not hold NaN patterns
    for (uint32_t i = tid; i < NBUF * A_BUF / 16; i += LT_THREADS) reinterpret_cast<uint4*>(a_ring)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < 2 * W_PLANE / 16; i += LT_THREADS)
        reinterpret_cast<uint4*>(w_tile)[i] = __ldg(reinterpret_cast<const uint4*>(prm.w_img) + i);
    if (tid < LT_N) s_bias[tid] = __ldg(prm.bias + tid);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    fence_proxy_async();                                      // W tile + zeros -> visible to the tensor core's proxy
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = idesc_bf16_m128(LT_N);
    const int ksteps = (prm.C + 15) >> 4;                     // K16 steps that hold tap slots (prm.C = samples * LT_TS)

    const int64_t nchunks = (prm.P + LT_M - 1) / LT_M;

    // ---- coordinate bookkeeping + gather of one chunk into A[buf] (all threads; ends with a CTA barrier) ----
    auto gather = [&](int64_t chunk, int buf) {
        const int64_t p0 = chunk * LT_M;
        const int npix = (int)((prm.P - p0) < LT_M ? (prm.P - p0) : LT_M);
        if (tid < npix) {
            const int64_t p = p0 + tid;
            float cx = prm.coords[p];
            if (prm.delta) {
                cx += prm.delta[p * prm.delta_C];
                prm.coords[p] = cx;
            }
            if (!GEO && prm.flow) prm.flow[p * 2] = cx - (float)(p % prm.W1);
            s_x[tid] = cx;
        }
        __syncthreads();
        uint8_t* a_buf = a_ring + (uint32_t)buf * A_BUF;
        if (!GEO) {
            if (!(prm.flags & 1)) {
                // RAFT: unit = (pixel, level); 4 lanes per pixel; a thread issues its 10 loads back to back
                for (int u = tid; u < npix * DKT_MAX_LEVELS; u += LT_THREADS) {
                    const int px = u >> 2, l = u & 3;
                    if (l >= prm.levels) continue;
                    float v[2 * R + 3], a;
                    sample_row_load<R>(prm.vol[l] + (p0 + px) * prm.vw[l], prm.vw[l], s_x[px] * (1.f / (float)(1 << l)), v, a);
                    v[2 * R + 2] = 0.f;
                    put_taps<AP>(a_buf, A_PLANE, px, l * LT_TS, v, a);
                }
            } else {
            // RAFT: a row sample = (pixel, level), 4 per pixel; two halves of 64 pixels = 256 samples each.
            float* win = reinterpret_cast<float*>(stage);
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int half = 0; half < 2; ++half) {
                // pass 1: lane (sample, part) fetches elements [a0, a0 + 4) of the sample's aligned 16-float window
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int u = tid + it * LT_THREADS;
                    const int sl = u >> 2, part = u & 3;
                    const int px = half * 64 + (sl >> 2), l = sl & 3;
                    float4 val = zero4;
                    if (px < npix && l < prm.levels) {
                        const int Wl = prm.vw[l];
                        const float xf = floorf(fminf(fmaxf(s_x[px] * (1.f / (float)(1 << l)), -1.0e6f), 1.0e6f));
                        const int64_t rs = (p0 + px) * Wl;                 // first element of this pixel's row
                        const int64_t e = rs + ((int)xf - R);              // first element the taps need
                        const int64_t a0 = ((e >> 2) << 2) + 4 * part;     // this lane's 4 elements (16-byte aligned)
                        if (a0 + 3 >= rs && a0 < rs + Wl) {
                            const float* src = prm.vol[l] + a0;
                            if (a0 >= rs && a0 + 4 <= rs + Wl) {
                                val = __ldg(reinterpret_cast<const float4*>(src));
                            } else {                                       // straddles a row end: zero outside the row
                                val.x = (a0 + 0 >= rs && a0 + 0 < rs + Wl) ? __ldg(src + 0) : 0.f;
                                val.y = (a0 + 1 >= rs && a0 + 1 < rs + Wl) ? __ldg(src + 1) : 0.f;
                                val.z = (a0 + 2 >= rs && a0 + 2 < rs + Wl) ? __ldg(src + 2) : 0.f;
                                val.w = (a0 + 3 >= rs && a0 + 3 < rs + Wl) ? __ldg(src + 3) : 0.f;
                            }
                        }
                    }
                    float2* dst = reinterpret_cast<float2*>(win + sl * LT_WS + part * 4);
                    dst[0] = make_float2(val.x, val.y);
                    dst[1] = make_float2(val.z, val.w);
                }
                __syncthreads();
                // pass 2: one thread per sample interpolates its 9 taps out of the window
                {
                    const int sl = tid;
                    const int px = half * 64 + (sl >> 2), l = sl & 3;
                    if (px < npix && l < prm.levels) {
                        const float x = fminf(fmaxf(s_x[px] * (1.f / (float)(1 << l)), -1.0e6f), 1.0e6f);
                        const float xf = floorf(x);
                        const int64_t e = (p0 + px) * prm.vw[l] + ((int)xf - R);
                        const float* wv = win + sl * LT_WS + (int)(e - ((e >> 2) << 2));
                        float v[2 * R + 3];
#pragma unroll
                        for (int k = 0; k < 2 * R + 2; ++k) v[k] = wv[k];
                        v[2 * R + 2] = 0.f;
                        put_taps<AP>(a_buf, A_PLANE, px, l * LT_TS, v, x - xf);
                    }
                }
                __syncthreads();
            }
            }
        } else {
            // IGEV: unit = (pixel, level, j): j < Cg one geometry channel (its taps sit Cg floats apart in the
            // (.., D, Cg) layout, so the Cg lanes of a (pixel, level) read consecutive floats), j == Cg the init-corr row.
            // Output channel order of the reference (geometry.py:36-57): per level [geo (c-major, tap-minor), init].
            constexpr int Cg = 8, G1 = Cg + 1, G = 2 * G1;      // host checks prm.Cg == 8 (IGEV's 8 geometry channels)
            constexpr int GU = 3;                 // row samples whose loads a thread has in flight before interpolating
            for (int u0 = tid; u0 < npix * G; u0 += LT_THREADS * GU) {
                float v[GU][2 * R + 3], av[GU];
#pragma unroll
                for (int g = 0; g < GU; ++g) {
                    const int u = u0 + g * LT_THREADS;
                    if (u >= npix * G) continue;
                    const int px = u / G, gi = u - px * G;
                    const int l = gi >= G1, j = gi - l * G1;
                    const int64_t p = p0 + px;
                    const float d = s_x[px];
                    const float inv = l ? 0.5f : 1.f;
                    v[g][2 * R + 2] = 0.f;
                    if (j < Cg) {
                        const int Dl = l ? prm.D / 2 : prm.D;
                        const float x = d * inv, xf = floorf(x);
                        av[g] = x - xf;
                        const int i0 = (int)fminf(fmaxf(xf, -1.0e6f), 1.0e6f) - R;
                        const float* ptr = prm.geo[l] + (p * Dl + i0) * Cg + j;     // only dereferenced inside [0, Dl)
#pragma unroll
                        for (int k = 0; k < 2 * R + 2; ++k)
                            v[g][k] = ((unsigned)(i0 + k) < (unsigned)Dl) ? __ldg(ptr + k * Cg) : 0.f;
                    } else {
                        const int Wl = prm.vw[l];
                        const float x = (float)(p % prm.W1);
                        sample_row_load<R>(prm.vol[l] + p * Wl, Wl, x * inv - d * inv, v[g], av[g]);
                    }
                }
#pragma unroll
                for (int g = 0; g < GU; ++g) {
                    const int u = u0 + g * LT_THREADS;
                    if (u >= npix * G) continue;
                    const int px = u / G, gi = u - px * G;
                    put_taps<AP>(a_buf, A_PLANE, px, gi * LT_TS, v[g], av[g]);
                }
            }
        }
        fence_proxy_async();                                  // generic-proxy stores -> async proxy (tcgen05.mma operand reads)
        __syncthreads();
    };

    int buf = 0;
    uint32_t phase = 0;
    int64_t chunk = blockIdx.x;
    if (chunk < nchunks) gather(chunk, 0);
    for (; chunk < nchunks; chunk += gridDim.x) {
        // ---- MMAs of this chunk: one elected lane of warp 0 ----
        if (warp == 0) {
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t a_hi = smem_u32(a_ring + (uint32_t)buf * A_BUF), a_lo = a_hi + A_PLANE;
                const uint32_t w_hi = smem_u32(w_tile), w_lo = w_hi + W_PLANE;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t ao = (uint32_t)(ks >> 2) * LT_A_KB_BYTES + (uint32_t)(ks & 3) * 32u;
                    const uint32_t wo = (uint32_t)(ks >> 2) * LT_W_KB_BYTES + (uint32_t)(ks & 3) * 32u;
                    const uint32_t acc = ks != 0;
                    umma_bf16(tmem_base, smem_desc_sw128(a_hi + ao), smem_desc_sw128(w_hi + wo), idesc, acc);
                    if (AP == 2) umma_bf16(tmem_base + LT_N, smem_desc_sw128(a_lo + ao), smem_desc_sw128(w_hi + wo), idesc, acc);
                    umma_bf16(tmem_base + LT_N, smem_desc_sw128(a_hi + ao), smem_desc_sw128(w_lo + wo), idesc, AP == 2 ? 1u : acc);
                }
                umma_commit(bar);
            }
            __syncwarp();
        }
        const int64_t next = chunk + gridDim.x;
        if (NBUF == 2 && next < nchunks) gather(next, buf ^ 1);          // overlaps the MMAs in flight
        mbar_wait(bar, phase);
        phase ^= 1u;
        tcgen05_fence_after();
        // ---- epilogue: warp = TMEM lane quarter (warp & 3) x column half (warp >> 2); thread = pixel ----
        {
            const int64_t p0 = chunk * LT_M;
            const int npix = (int)((prm.P - p0) < LT_M ? (prm.P - p0) : LT_M);
            const int q = warp & 3, c0 = (warp >> 2) * 32;
            const int m = q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            const dkt_tensor& o = prm.out;
            const int64_t off0 = (p0 + m) * o.C + o.c_begin + c0;
            // 16-bit planes leave through the staging tile: a thread owns a pixel's 32 channels (its TMEM lane), the copy-out
            // writes whole 128-byte rows.  Row r keeps its 16-byte chunk c at position c ^ (r & 7): conflict-free both ways.
            const int nplanes = o.hi ? (o.lo ? 2 : 1) : 0;
            const bool staged = (prm.flags & 2) != 0;
            for (int pl = 0; pl < (nplanes ? nplanes : 1); ++pl) {
#pragma unroll
                for (int cc = 0; cc < 32; cc += 16) {         // 16 columns at a time: 32 live accumulator registers
                    float v[16], v2[16];
                    tmem_ld16(taddr + cc, v);
                    tmem_ld16(taddr + LT_N + cc, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + v2[j] + s_bias[c0 + cc + j], 0.f);
                    if (pl == 0 && o.f32 && m < npix) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(o.f32 + off0 + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (nplanes) {
                        uint32_t w16[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            uint32_t h, l;
                            split16x2(v[2 * t], v[2 * t + 1], h, l);
                            w16[t] = pl == 0 ? h : l;
                        }
                        if (staged) {
                            const int ch0 = (c0 + cc) >> 3;   // first of this pass's two 16-byte chunks
                            uint8_t* rowp = stage + m * 128;
                            *reinterpret_cast<uint4*>(rowp + (((ch0 + 0) ^ (m & 7)) << 4)) = make_uint4(w16[0], w16[1], w16[2], w16[3]);
                            *reinterpret_cast<uint4*>(rowp + (((ch0 + 1) ^ (m & 7)) << 4)) = make_uint4(w16[4], w16[5], w16[6], w16[7]);
                        } else if (m < npix) {
                            uint16_t* const dp = (pl == 0 ? o.hi : o.lo) + off0 + cc;
                            *reinterpret_cast<uint4*>(dp) = make_uint4(w16[0], w16[1], w16[2], w16[3]);
                            *reinterpret_cast<uint4*>(dp + 8) = make_uint4(w16[4], w16[5], w16[6], w16[7]);
                        }
                    }
                }
                if (nplanes && staged) {
                    __syncthreads();
                    uint16_t* const dstp = pl == 0 ? o.hi : o.lo;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = tid + i * LT_THREADS;  // 128 rows x 8 chunks
                        const int r = idx >> 3, c = idx & 7;
                        if (r < npix)
                            *reinterpret_cast<uint4*>(dstp + (p0 + r) * o.C + o.c_begin + c * 8) =
                                *reinterpret_cast<const uint4*>(stage + r * 128 + ((c ^ (r & 7)) << 4));
                    }
                    if (pl + 1 < nplanes) __syncthreads();    // the lo plane reuses the tile
                }
            }
        }
        tcgen05_fence_before();
        __syncthreads();                                      // accumulator drained before the next chunk's MMAs overwrite it
        if (NBUF == 2) buf ^= 1;
        else if (next < nchunks) gather(next, 0);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------------------------
// IGEV lookup with the geometry runs fetched by the TMA unit (default for hi-plane taps).
//
// Why (tools/dram_random_probe.cu, profiles/r2w_dram_random_probe.txt, measured on B200): a gather made of per-thread
// loads is bound by the number of L1 -> L2 sector REQUESTS, ~50 G/s on the whole chip no matter how many are in flight
// (10-float runs 24 G runs/s, 80-float runs 5 G runs/s = 1.6 TB/s, against 6.7 TB/s streaming): the 320-byte run a
// (pixel, level) needs from the (B,H,W,D,8) geometry volume costs ten of them.  One cp.async.bulk (1-D TMA) per run
// moves the same bytes as whole-line requests: 17.8 G runs/s = 5.7 TB/s in the same probe.
//
//   warps 0..7    producers: per 32-pixel sub-chunk, lanes 0..15 of each warp own one (pixel, volume) each: read disp
//                 (+ delta; written back), wait for the stage to be free, add their byte count to its mbarrier and issue
//                 one bulk copy -- clamped to the samples inside [0, D) / to the array -- into a 4-stage ring.  (One warp
//                 issues a bulk copy every ~59 clocks through its elect / uniform-register loop: a single producer warp
//                 was the bottleneck of the first version; eight of them overlap.)
//   warps 8..17   gather: wait for a stage, read a (pixel, level, channel pair)'s ten samples back from shared memory
//                 (slot stride 88 floats: conflict-free), interpolate, store the taps as 16-bit values into the K-major
//                 SWIZZLE_128B A tile (put_taps); the init-corr rows come the same way as the aligned 64-byte window
//                 around their 10 floats.  After 4 sub-chunks one elected thread issues the 12 x 2 tcgen05.mma of the
//                 chunk into one of TWO TMEM accumulators and the warps go on to the next chunk
//   warps 18..21  epilogue: accumulator -> bias, ReLU -> fp32 / 16-bit planes through a staging tile -> 128-byte rows
// ---------------------------------------------------------------------------------------------------------------
constexpr int GT_SUB = 32;                       // pixels per ring stage
constexpr int GT_NST = 4;                        // ring stages
constexpr int GT_RUN = 88;                       // floats per (pixel, level) slot (80 used)
constexpr int GT_WIN = 16;                       // floats per init-corr window (the aligned 64 bytes that hold a 10-float run)
constexpr uint32_t GT_GEO_BYTES = GT_SUB * 2 * GT_RUN * 4;        // 22528
constexpr uint32_t GT_STAGE_BYTES = GT_GEO_BYTES + GT_SUB * 2 * GT_WIN * 4;      // + 4096 = 26 x 1024
constexpr int GT_PROD_WARPS = 8, GT_GATHER_WARPS = 10, GT_EPI_WARPS = 4;   // gather: 8 warps of geometry units + 2 of init rows
constexpr int GT_ISSUE_LANES = LT_M / GT_PROD_WARPS;        // 128 bulk copies per sub-chunk: lanes 0..15 of each producer warp
constexpr int GT_THREADS = 32 * (GT_PROD_WARPS + GT_GATHER_WARPS + GT_EPI_WARPS);
constexpr int GT_KB = 3;
constexpr size_t GT_SMEM = 1024 + (size_t)GT_KB * LT_A_KB_BYTES + 2 * (size_t)GT_KB * LT_W_KB_BYTES + GT_NST * GT_STAGE_BYTES +
                           LT_M * 128 + (GT_NST * GT_SUB + LT_N) * 4 + 16 * 8 + 64;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int R>
__global__ void __launch_bounds__(GT_THREADS, 1)
geo_lookup_tma_kernel(const __grid_constant__ LookupTcParams prm) {
    static_assert(2 * R + 2 == LT_TS, "tap slots");
    constexpr int Cg = 8, G1 = Cg + 1;
    constexpr uint32_t A_PLANE = GT_KB * LT_A_KB_BYTES;
    constexpr uint32_t W_PLANE = GT_KB * LT_W_KB_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;                                   // [KB][128 x 128 B], hi plane
    uint8_t* w_tile = a_tile + A_PLANE;                       // [2][KB][64 x 128 B]
    uint8_t* ring = w_tile + 2 * W_PLANE;                     // [NST][32 pixels][2 levels][GT_RUN floats]
    uint8_t* ostage = ring + GT_NST * GT_STAGE_BYTES;         // [128][128 B] output rows
    float* s_xr = reinterpret_cast<float*>(ostage + LT_M * 128);       // [NST][32] disparity of the staged pixels
    float* s_bias = s_xr + GT_NST * GT_SUB;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_bias + LT_N);       // [NST]
    uint64_t* empty = full + GT_NST;                                    // [NST]
    uint64_t* a_free = empty + GT_NST;                                  // MMAs of a chunk retired: A tile reusable
    uint64_t* acc_full = a_free + 1;                                    // [2]
    uint64_t* acc_empty = acc_full + 2;                                 // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (uint32_t i = tid; i < A_PLANE / 16; i += GT_THREADS) reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < 2 * W_PLANE / 16; i += GT_THREADS)
        reinterpret_cast<uint4*>(w_tile)[i] = __ldg(reinterpret_cast<const uint4*>(prm.w_img) + i);
    if (tid < LT_N) s_bias[tid] = __ldg(prm.bias + tid);
    if (tid == 0) {
        for (int i = 0; i < GT_NST; ++i) { mbar_init(&full[i], GT_PROD_WARPS * GT_ISSUE_LANES); mbar_init(&empty[i], GT_GATHER_WARPS); }
        mbar_init(a_free, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], GT_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == GT_PROD_WARPS) tmem_alloc(tmem_slot, 256);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t nchunks = (prm.P + LT_M - 1) / LT_M;
    const int D0 = prm.D, D1 = prm.D / 2;

    constexpr int SUBS = LT_M / GT_SUB;                        // sub-chunks per chunk
    // this CTA's sub-chunk sequence n -> (chunk, sb)
    const int64_t my_chunks = blockIdx.x < nchunks ? (nchunks - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t nsub = my_chunks * SUBS;
    auto sub_pixel0 = [&](int64_t n) { return ((int64_t)blockIdx.x + (n / SUBS) * gridDim.x) * LT_M + (n % SUBS) * GT_SUB; };

    if (warp < GT_PROD_WARPS) {
        // ===== producer warps: op = warp * 16 + lane (lane < 16): pixel op >> 2 of the sub-chunk, kind op & 3 = geo L0, geo L1,
        // init L0, init L1.  Every issuing thread arrives once per stage use with its own byte count.
        if (lane < GT_ISSUE_LANES) {
            const int op = warp * GT_ISSUE_LANES + lane, opx = op >> 2, okind = op & 3, l = okind & 1;
            // disp and delta of the NEXT sub-chunk are requested one iteration ahead and only added when that iteration
            // starts: an add right behind the loads would park the warp for a full memory latency per sub-chunk
            // (ncu r2z: that wait, not the copies, set the pace of the whole kernel)
            auto load_x = [&](int64_t n, float& c0, float& d0) {
                c0 = 0.f;
                d0 = 0.f;
                if (n < nsub) {
                    const int64_t p = sub_pixel0(n) + opx;
                    if (p < prm.P) {
                        c0 = prm.coords[p];
                        if (prm.delta) d0 = __ldg(prm.delta + p * prm.delta_C);
                    }
                }
            };
            float c_nx, d_nx;
            load_x(0, c_nx, d_nx);
            for (int64_t n = 0; n < nsub; ++n) {
                const int st = (int)(n % GT_NST);
                const uint32_t ph = (uint32_t)(n / GT_NST) & 1u;
                const float cx = c_nx + d_nx;
                load_x(n + 1, c_nx, d_nx);
                const int64_t p = sub_pixel0(n) + opx;
                uint32_t bytes = 0u, dst = 0u;
                const float* src = nullptr;
                if (p < prm.P) {
                    if (okind < 2) {
                        const int Dl = l ? D1 : D0;
                        const float xf = floorf(cx * (l ? 0.5f : 1.f));
                        const int i0 = (int)fminf(fmaxf(xf, -1.0e6f), 1.0e6f) - R;
                        const int klo = i0 < 0 ? -i0 : 0;
                        const int khi = (Dl - i0) < (2 * R + 2) ? (Dl - i0) : (2 * R + 2);
                        if (khi > klo) {
                            bytes = (uint32_t)(khi - klo) * (Cg * 4);
                            src = prm.geo[l] + (p * Dl + i0 + klo) * Cg;
                            dst = smem_u32(ring + (size_t)st * GT_STAGE_BYTES) + (uint32_t)(((opx * 2 + l) * GT_RUN + klo * Cg) * 4);
                        }
                    } else {
                        // init-corr row: the aligned 16-float window around the 10-float run, clipped to the level's array (the
                        // host checks that its length is a multiple of 4 floats); a neighbouring row's elements are masked by the reader
                        const int Wl = prm.vw[l];
                        const float inv = l ? 0.5f : 1.f;
                        const float xf = floorf((float)(p % prm.W1) * inv - cx * inv);
                        const int i0 = (int)fminf(fmaxf(xf, -1.0e6f), 1.0e6f) - R;
                        if (i0 + 2 * R + 1 >= 0 && i0 < Wl) {
                            const int64_t e = p * Wl + i0, total = prm.P * Wl;
                            const int64_t a0 = (e >> 2) << 2;
                            const int64_t lo = a0 < 0 ? 0 : a0, hi = (a0 + GT_WIN) > total ? total : (a0 + GT_WIN);
                            if (hi > lo) {
                                bytes = (uint32_t)(hi - lo) * 4u;
                                src = prm.vol[l] + lo;
                                dst = smem_u32(ring + (size_t)st * GT_STAGE_BYTES + GT_GEO_BYTES) +
                                      (uint32_t)(((opx * 2 + l) * GT_WIN + (int)(lo - a0)) * 4);
                            }
                        }
                    }
                }
                mbar_wait_backoff(&empty[st], ph ^ 1u, 64);    // the gather warps are done with the stage's previous use
                if (okind == 0 && p < prm.P) {
                    if (prm.delta) prm.coords[p] = cx;
                    s_xr[st * GT_SUB + opx] = cx;
                }
                mbar_arrive_expect_tx(&full[st], bytes);       // release: this thread's s_xr store is ordered before it
                if (bytes) bulk_g2s(dst, src, bytes, &full[st]);
            }
        }
    } else if (warp < GT_PROD_WARPS + GT_GATHER_WARPS) {
        // ===== gather warps =====
        const int ct = tid - 32 * GT_PROD_WARPS;               // 0..319
        const uint32_t idesc = idesc_bf16_m128(LT_N);
        const int ksteps = (prm.C + 15) >> 4;
        int64_t n = 0;
        int ci = 0;
        for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++ci) {
            if (ci > 0) mbar_wait(a_free, (uint32_t)(ci - 1) & 1u);
            for (int sb = 0; sb < SUBS; ++sb, ++n) {
                const int st = (int)(n % GT_NST);
                const uint32_t ph = (uint32_t)(n / GT_NST) & 1u;
                const int64_t pb = chunk * LT_M + sb * GT_SUB;
                const int npix_s = (int)((prm.P - pb) < GT_SUB ? (prm.P - pb) : GT_SUB);      // may be <= 0
                mbar_wait(&full[st], ph);
                const float* stg = reinterpret_cast<const float*>(ring + (size_t)st * GT_STAGE_BYTES);
                const float* sx = s_xr + st * GT_SUB;
                if (ct < 8 * GT_SUB) {   // geometry unit of this thread: (pixel, level, channel pair): ten 8-byte reads, two tap rows
                    const int cp = ct & 3, l = (ct >> 2) & 1, px = ct >> 3;
                    if (px < npix_s) {
                        const int Dl = l ? D1 : D0;
                        const float x = sx[px] * (l ? 0.5f : 1.f), xf = floorf(x);
                        const int i0 = (int)fminf(fmaxf(xf, -1.0e6f), 1.0e6f) - R;
                        const float2* base = reinterpret_cast<const float2*>(stg + (px * 2 + l) * GT_RUN + 2 * cp);
                        float v0[2 * R + 3], v1[2 * R + 3];
#pragma unroll
                        for (int k = 0; k < 2 * R + 2; ++k) {
                            const float2 q = ((unsigned)(i0 + k) < (unsigned)Dl) ? base[k * (Cg / 2)] : make_float2(0.f, 0.f);
                            v0[k] = q.x;
                            v1[k] = q.y;
                        }
                        v0[2 * R + 2] = 0.f;
                        v1[2 * R + 2] = 0.f;
                        put_taps<1>(a_tile, A_PLANE, sb * GT_SUB + px, (l * G1 + 2 * cp) * LT_TS, v0, x - xf);
                        put_taps<1>(a_tile, A_PLANE, sb * GT_SUB + px, (l * G1 + 2 * cp + 1) * LT_TS, v1, x - xf);
                    }
                }
                // init-corr row of (pixel, level): the 64 threads of the last two gather warps have one each
                const int ipx = (ct - 8 * GT_SUB) >> 1, il = ct & 1;
                if (ct >= 8 * GT_SUB && ipx < npix_s) {
                    const int64_t p = pb + ipx;
                    const int Wl = prm.vw[il];
                    const float inv = il ? 0.5f : 1.f;
                    const float x = (float)(p % prm.W1) * inv - sx[ipx] * inv, xf = floorf(x);
                    const int i0 = (int)fminf(fmaxf(xf, -1.0e6f), 1.0e6f) - R;
                    const int64_t e = p * Wl + i0;
                    const float* win = stg + GT_GEO_BYTES / 4 + (ipx * 2 + il) * GT_WIN + (int)(e - ((e >> 2) << 2));
                    float vi[2 * R + 3];
#pragma unroll
                    for (int k = 0; k < 2 * R + 2; ++k) vi[k] = ((unsigned)(i0 + k) < (unsigned)Wl) ? win[k] : 0.f;
                    vi[2 * R + 2] = 0.f;
                    put_taps<1>(a_tile, A_PLANE, sb * GT_SUB + ipx, (il * G1 + Cg) * LT_TS, vi, x - xf);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
            fence_proxy_async();                               // tap stores -> visible to the tensor core's proxy
            asm volatile("bar.sync 1, %0;" ::"n"(32 * GT_GATHER_WARPS) : "memory");
            if (warp == GT_PROD_WARPS) {
                const uint32_t acc = (uint32_t)ci & 1u;
                if (ci >= 2) mbar_wait(&acc_empty[acc], (uint32_t)((ci >> 1) - 1) & 1u);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(a_tile), w_hi = smem_u32(w_tile), w_lo = w_hi + W_PLANE;
                    const uint32_t d = tmem_base + acc * 128u;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t ao = (uint32_t)(ks >> 2) * LT_A_KB_BYTES + (uint32_t)(ks & 3) * 32u;
                        const uint32_t wo = (uint32_t)(ks >> 2) * LT_W_KB_BYTES + (uint32_t)(ks & 3) * 32u;
                        umma_bf16(d, smem_desc_sw128(a_hi + ao), smem_desc_sw128(w_hi + wo), idesc, ks != 0);
                        umma_bf16(d + LT_N, smem_desc_sw128(a_hi + ao), smem_desc_sw128(w_lo + wo), idesc, ks != 0);
                    }
                    umma_commit(a_free);
                    umma_commit(&acc_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4, thread = pixel =====
        const int et = tid - 32 * (GT_PROD_WARPS + GT_GATHER_WARPS);          // 0..127
        const int q = warp & 3, m = q * 32 + lane;
        const dkt_tensor& o = prm.out;
        const int nplanes = o.hi ? (o.lo ? 2 : 1) : 0;
        int ci = 0;
        for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++ci) {
            const uint32_t acc = (uint32_t)ci & 1u;
            const int64_t p0 = chunk * LT_M;
            const int npix = (int)((prm.P - p0) < LT_M ? (prm.P - p0) : LT_M);
            mbar_wait_backoff(&acc_full[acc], (uint32_t)(ci >> 1) & 1u, 256);   // sleep between polls: a spinning warp per
            tcgen05_fence_after();                                              // scheduler took a quarter of the issue slots
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128u;
            const int64_t off0 = (p0 + m) * o.C + o.c_begin;
            for (int pl = 0; pl < (nplanes ? nplanes : 1); ++pl) {
#pragma unroll
                for (int cc = 0; cc < LT_N; cc += 16) {
                    float v[16], v2[16];
                    tmem_ld16(taddr + cc, v);
                    tmem_ld16(taddr + LT_N + cc, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + v2[j] + s_bias[cc + j], 0.f);
                    if (pl == 0 && o.f32 && m < npix) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(o.f32 + off0 + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (nplanes) {
                        uint32_t w16[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            uint32_t h, l;
                            split16x2(v[2 * t], v[2 * t + 1], h, l);
                            w16[t] = pl == 0 ? h : l;
                        }
                        const int ch0 = cc >> 3;
                        uint8_t* rowp = ostage + m * 128;
                        *reinterpret_cast<uint4*>(rowp + (((ch0 + 0) ^ (m & 7)) << 4)) = make_uint4(w16[0], w16[1], w16[2], w16[3]);
                        *reinterpret_cast<uint4*>(rowp + (((ch0 + 1) ^ (m & 7)) << 4)) = make_uint4(w16[4], w16[5], w16[6], w16[7]);
                    }
                }
                if (pl + 1 >= nplanes) {                       // last read of this accumulator: hand it back to the MMA issuer
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                if (nplanes) {
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    uint16_t* const dstp = pl == 0 ? o.hi : o.lo;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int idx = et + i * 128;          // 128 rows x 8 chunks
                        const int r = idx >> 3, c = idx & 7;
                        if (r < npix)
                            *reinterpret_cast<uint4*>(dstp + (p0 + r) * o.C + o.c_begin + c * 8) =
                                *reinterpret_cast<const uint4*>(ostage + r * 128 + ((c ^ (r & 7)) << 4));
                    }
                    asm volatile("bar.sync 2, 128;" ::: "memory");   // the next plane / chunk reuses the tile
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == GT_PROD_WARPS) tmem_dealloc(tmem_base, 256);
}

// ---- IGEV geometry encoding volume: (B,C,D,H,W) -> level 0 (B,H,W,D,C) + pooled level 1 (B,H,W,D/2,C) ----
// one CTA per (b, y, 32-pixel segment): the [C*D][32] slab is read along w (coalesced) and written per pixel as D x C
// contiguous floats.
__global__ void __launch_bounds__(256)
geo_pool_dc_kernel(const float* __restrict__ gev, float* __restrict__ geo0, float* __restrict__ geo1,
                   int C, int D, int H, int W) {
    extern __shared__ float slab[];              // [C*D][33]
    const int by = blockIdx.y;                   // b*H + y
    const int b = by / H, y = by - b * H;
    const int w0 = blockIdx.x * 32;
    const int CD = C * D;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const float* src = gev + (int64_t)b * CD * H * W + (int64_t)y * W + w0;
    for (int r = wrp; r < CD; r += 8)            // r = c*D + d
        slab[r * 33 + lane] = (w0 + lane < W) ? __ldg(src + (int64_t)r * H * W + lane) : 0.f;
    __syncthreads();
    const int D1 = D / 2;
    for (int px = 0; px < 32 && w0 + px < W; ++px) {
        const int64_t pix = ((int64_t)b * H + y) * W + w0 + px;
        float* o0 = geo0 + pix * D * C;
        float* o1 = geo1 + pix * D1 * C;
        for (int i = threadIdx.x; i < CD; i += 256) {         // i = d*C + c (output order)
            const int d = i / C, c = i - d * C;
            o0[i] = slab[(c * D + d) * 33 + px];
        }
        for (int i = threadIdx.x; i < D1 * C; i += 256) {
            const int d = i / C, c = i - d * C;
            o1[i] = (slab[(c * D + 2 * d) * 33 + px] + slab[(c * D + 2 * d + 1) * 33 + px]) * 0.5f;
        }
    }
}

}  // namespace dkt

using namespace dkt;

static int check_enc_out(const dkt_tensor* t) {
    if (!t || !(t->f32 || t->hi) || t->c_count != LT_N) return DKT_E_INVALID;
    if ((t->C % 8) || (t->c_begin % 8)) return DKT_E_ALIGNMENT;
    if ((reinterpret_cast<uintptr_t>(t->f32) & 15) || (reinterpret_cast<uintptr_t>(t->hi) & 15) ||
        (reinterpret_cast<uintptr_t>(t->lo) & 15))
        return DKT_E_ALIGNMENT;
    return 0;
}

template <int KB, int AP, int NBUF, bool GEO>
static int launch_lookup_tc(LookupTcParams prm, cudaStream_t st) {
    // A/B knob (default = what measured best on B200, profiles/r2m_lookup_variants.txt)
    static const int s_flags = [] { const char* v = getenv("DKT_LOOKUP_FLAGS"); return v ? atoi(v) : -1; }();
    prm.flags = s_flags >= 0 ? s_flags : LT_DEFAULT_FLAGS;
    constexpr size_t smem = 1024 + (size_t)NBUF * AP * KB * LT_A_KB_BYTES + 2 * (size_t)KB * LT_W_KB_BYTES +
                            (GEO ? 0 : LT_STAGE_BYTES) + (LT_M + LT_N) * 4 + 64;
    DKT_ENSURE_SMEM(smem, lookup_tc_kernel<4, KB, AP, NBUF, GEO>);
    // The gather lives on the L1 cache (a row sample's 10 loads touch one or two lines; with the carve-out at its maximum
    // the same kernel is 2x slower, profiles/r2n_lookup_variants.txt), so the shared-memory carve-out is requested
    // explicitly: just enough for the CTAs the grid is sized for.
    static const int s_carve = [] { const char* v = getenv("DKT_LOOKUP_CARVEOUT"); return v ? atoi(v) : -2; }();
    static const int s_ctas = [] { const char* v = getenv("DKT_LOOKUP_CTAS"); return v ? atoi(v) : 0; }();
    const int by_smem = (int)((227 * 1024) / (smem + 1024));
    // __launch_bounds__ of the kernel; RAFT with (hi, lo) taps: 2 CTAs of 68 KB leave the L1 more than a third CTA gives
    const int by_regs = GEO ? 2 : (AP == 2 ? 2 : 3);
    int ctas_sm = by_smem < 1 ? 1 : (by_smem > by_regs ? by_regs : by_smem);
    if (s_ctas > 0 && s_ctas < ctas_sm) ctas_sm = s_ctas;
    {
        static std::atomic<uint64_t> done{0};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
        const uint64_t bit = 1ull << (dev & 63);
        if (dev > 63 || !(done.load(std::memory_order_acquire) & bit)) {
            // percent of the 228 KB maximum that holds ctas_sm CTAs (+1 KB each of system use), rounded up
            int pct = s_carve != -2 ? s_carve : (int)((ctas_sm * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            cudaError_t ce = cudaFuncSetAttribute(lookup_tc_kernel<4, KB, AP, NBUF, GEO>,
                                                  cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (ce != cudaSuccess) return (int)ce;
            done.fetch_or(bit, std::memory_order_release);
        }
    }
    const int64_t chunks = ceil_div64(prm.P, LT_M), cap = (int64_t)device_sms() * ctas_sm;
    lookup_tc_kernel<4, KB, AP, NBUF, GEO><<<(unsigned)(chunks < cap ? chunks : cap), LT_THREADS, smem, st>>>(prm);
    DKT_RETURN_LAST();
}

extern "C" int dkt_corr1d_lookup_enc_tc(const float* const* pyr, int levels, int radius,
                                        float* coords_x, const float* delta, int delta_C, float* flow,
                                        const uint16_t* w_img, const float* enc_b, const dkt_tensor* enc_out, int tap_planes,
                                        int B, int H, int W1, int W2, void* stream) {
    DKT_CHECK_ARG(pyr && coords_x && w_img && enc_b);
    DKT_CHECK_ARG(B > 0 && H > 0 && W1 > 0 && W2 > 0);
    if (int rc = check_enc_out(enc_out)) return rc;
    if (radius != 4 || levels < 1 || levels > DKT_MAX_LEVELS) return DKT_E_UNSUPPORTED;
    if (tap_planes != 1 && tap_planes != 2) return DKT_E_INVALID;
    if (delta) DKT_CHECK_ARG(delta_C > 0);
    if (reinterpret_cast<uintptr_t>(w_img) & 15) return DKT_E_ALIGNMENT;
    LookupTcParams prm{};
    int w = W2;
    for (int l = 0; l < DKT_MAX_LEVELS; ++l) {
        prm.vol[l] = l < levels ? pyr[l] : nullptr;
        prm.vw[l] = w;
        if (l < levels) {
            DKT_CHECK_ARG(pyr[l] != nullptr && w > 0);
            if (reinterpret_cast<uintptr_t>(pyr[l]) & 15) return DKT_E_ALIGNMENT;      // 16-byte window loads
        }
        w /= 2;
    }
    prm.levels = levels; prm.W1 = W1;
    prm.coords = coords_x; prm.delta = delta; prm.delta_C = delta_C; prm.flow = flow;
    prm.w_img = w_img; prm.bias = enc_b; prm.out = *enc_out;
    prm.P = (int64_t)B * H * W1;
    prm.C = levels * LT_TS;
    // one tap tile per CTA (48 / 32 KB of shared memory): three CTAs per SM hide the gather's latency better than a
    // second tile per CTA would
    return tap_planes == 2 ? launch_lookup_tc<1, 2, 1, false>(prm, (cudaStream_t)stream)
                           : launch_lookup_tc<1, 1, 1, false>(prm, (cudaStream_t)stream);
}

extern "C" int dkt_geo_pool_dc(const float* gev, float* geo0, float* geo1, int B, int C, int D, int H, int W, void* stream) {
    DKT_CHECK_ARG(gev && geo0 && geo1);
    DKT_CHECK_ARG(B > 0 && C > 0 && D > 1 && H > 0 && W > 0);
    const size_t smem = (size_t)C * D * 33 * 4;
    if (smem > 200 * 1024 || (int64_t)B * H > 65535) return DKT_E_UNSUPPORTED;
    DKT_ENSURE_SMEM(200 * 1024, geo_pool_dc_kernel);
    geo_pool_dc_kernel<<<dim3(ceil_div(W, 32), B * H), 256, smem, (cudaStream_t)stream>>>(gev, geo0, geo1, C, D, H, W);
    DKT_RETURN_LAST();
}

extern "C" int dkt_geo_lookup_enc_tc(const float* geo0, const float* geo1, const float* init0, const float* init1,
                                     float* disp, const float* delta, int delta_C, int radius, int C, int D,
                                     const uint16_t* w_img, const float* enc_b, const dkt_tensor* enc_out, int tap_planes,
                                     int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(geo0 && geo1 && init0 && init1 && disp && w_img && enc_b);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 1 && C > 0 && D > 1);
    if (int rc = check_enc_out(enc_out)) return rc;
    if (radius != 4 || C != 8) return DKT_E_UNSUPPORTED;      // IGEV's 8 geometry channels: 18 row samples x 10 slots = 180 <= 192
    if (tap_planes != 1 && tap_planes != 2) return DKT_E_INVALID;
    if (delta) DKT_CHECK_ARG(delta_C > 0);
    if (reinterpret_cast<uintptr_t>(w_img) & 15) return DKT_E_ALIGNMENT;
    LookupTcParams prm{};
    prm.geo[0] = geo0; prm.geo[1] = geo1;
    prm.vol[0] = init0; prm.vol[1] = init1;
    prm.vw[0] = W; prm.vw[1] = W / 2;
    prm.levels = 2; prm.Cg = C; prm.D = D; prm.W1 = W;
    prm.coords = disp; prm.delta = delta; prm.delta_C = delta_C; prm.flow = nullptr;
    prm.w_img = w_img; prm.bias = enc_b; prm.out = *enc_out;
    prm.P = (int64_t)B * H * W;
    prm.C = 2 * (C + 1) * LT_TS;
    {
        // hi-plane taps (what the update loop runs): geometry runs through the TMA unit; DKT_LOOKUP_TMA=0 keeps the
        // per-thread gather.  The bulk copies need 16-byte aligned volumes (32-byte runs of 8 channels).
        static const int s_tma = [] { const char* v = getenv("DKT_LOOKUP_TMA"); return (v && v[0] == '0') ? 0 : 1; }();
        const bool win_ok = (prm.P * prm.vw[0]) % 4 == 0 && (prm.P * prm.vw[1]) % 4 == 0 &&
                            !((reinterpret_cast<uintptr_t>(init0) | reinterpret_cast<uintptr_t>(init1)) & 15);
        if (s_tma && tap_planes == 1 && win_ok && !((reinterpret_cast<uintptr_t>(geo0) | reinterpret_cast<uintptr_t>(geo1)) & 15)) {
            DKT_ENSURE_SMEM(GT_SMEM, geo_lookup_tma_kernel<4>);
            const int64_t chunks = ceil_div64(prm.P, LT_M), cap = device_sms();
            geo_lookup_tma_kernel<4><<<(unsigned)(chunks < cap ? chunks : cap), GT_THREADS, GT_SMEM, (cudaStream_t)stream>>>(prm);
            DKT_RETURN_LAST();
        }
    }
    return tap_planes == 2 ? launch_lookup_tc<3, 2, 1, true>(prm, (cudaStream_t)stream)
                           : launch_lookup_tc<3, 1, 1, true>(prm, (cudaStream_t)stream);
}
