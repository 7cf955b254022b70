// K3 / tensor-core variant: implicit-GEMM convolution on tcgen05 with TMEM accumulators.
//
//   M = 128 output pixels (an 8 x 16 patch of one image), N = all output channels (<= 256),
//   K = taps x input channels, walked as (tap, 64-channel block) steps.
//
// No im2col buffer exists anywhere: for filter tap (ky,kx) the A tile is the SAME 8x16 patch
// shifted by (ky-pad, kx-pad), fetched by one 4-D TMA box {64 ch, 16, 8, 1} from the NHWC
// activation; pixels that fall outside the image are zero-filled by the TMA unit, which is
// exactly the conv's zero padding.  The box lands in shared memory as 128 rows of 128 bytes
// with the 128-byte swizzle, i.e. already in the canonical K-major UMMA layout.
//
// Precision: activations and weights are bf16 (hi, lo) pairs; each K step issues
// hi*hi + lo*hi + hi*lo into the same fp32 TMEM accumulator (3-term split, ~16 mantissa bits).
//
// conv_tc_kernel<BK, KIND, ACT> (v2) is PERSISTENT: grid = min(tiles, #SMs), each CTA walks tiles
// blockIdx.x, +gridDim.x, ...  Three pipelines run concurrently inside a CTA:
//   warp 0      TMA producer      smem ring of (A hi, A lo, W hi, W lo) stages, full/empty mbarriers
//   warp 1      MMA issuer        two TMEM accumulator stages (2 x Npad columns), tmem_full/empty
//   warps 4..11 epilogue          tile i is drained while the MMAs of tile i+1 run (warps 2, 3 idle: warpgroup alignment, TC2_EPI_WARP0)
// The epilogue is coalesced: a warp owns 32 pixels (its TMEM lane quarter); it pulls a 32-column
// chunk with tcgen05.ld (thread = pixel), transposes it through a swizzled 4 KB smem buffer and
// continues with lanes = 8 x 16-byte channel groups of 4 pixels, so every global access of the
// fused epilogue (context term, h, z, fp32 / bf16 hi / bf16 lo stores) is a full 128-byte line.
// The two warps that share a lane quarter take alternate column chunks.
// BK = K block in channels = 64 (SWIZZLE_128B rows).  Every kernel is instantiated per epilogue kind and
// activation (KIND, ACT): the epilogue loop is straight-line code for exactly one fused tail.
//
// Strided convolutions (the encoders' stride-2 layers) use the same kernel: the TMA tensor map is
// built with elementStrides = stride on the W/H dimensions, so the box {BK, 16*s, 8*s, 1} starting
// at (s*x0 + kx - pad, s*y0 + ky - pad) delivers exactly the 8x16 input samples of tap (ky,kx).
// Filters are kh x kw with independent padding (3x3, 1x1, and the 7x1 form of the 7x7 stem).
#include "common.cuh"
#include "tc.cuh"
#include <stdlib.h>
#include <type_traits>

namespace dkt {

using namespace tc;

constexpr int TC_TILE_W = 16, TC_TILE_H = 8;
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_ACC = 8;           // accumulator stages (tmem_full / tmem_empty barrier pairs)
constexpr uint32_t TC_BAR_BYTES = 512;  // barrier block at the end of the dynamic shared memory

struct TcConvParams {
    CUtensorMap act[DKT_MAX_SRCS][2];   // [source][hi/lo], 4-D (C, W, H, B)
    CUtensorMap wgt[2];                 // hi/lo, 2-D (Cin_total, taps*Npad)
    int nsrc;
    int c_begin[DKT_MAX_SRCS];
    int kblocks[DKT_MAX_SRCS];
    int c_count[DKT_MAX_SRCS];      // channels of the source's slice (weights: the source's K range starts at the sum before it)
    int klast[DKT_MAX_SRCS];        // K16 steps of the source's LAST K block (4 = full 64 channels; 2 for a 96-channel
                                    // source: the box is still 64 wide -- channels past the tensor are zero-filled by the
                                    // TMA unit, channels past the slice are simply never read by an MMA)
    int kh, kw, pad_y, pad_x, stride, taps;
    int N, Npad;
    int H, W, tiles_x, tiles_y;
    int num_tiles;
    int stages;
    uint32_t acc_cols;              // TMEM columns per accumulator stage
    uint32_t acc_stages;            // accumulator stages in TMEM: 2 (N > 128), 4 (N <= 128), 8 (N <= 64)
    uint32_t acc_lo_off;            // != 0: the lo-term MMAs (x_lo * w_hi, x_hi * w_lo) accumulate in a SECOND accumulator,
                                    // acc_lo_off columns after the main one; the epilogue adds the two in fp32 (round to
                                    // nearest).  Why: tcgen05.mma truncates (rounds toward zero) at every accumulate
                                    // (tools/tc_accum_probe.py: a relative bias of -1.65e-8 per instruction on zero-mean
                                    // data); keeping the small lo products out of the main accumulator divides the
                                    // number of truncations it sees by the number of MMAs per K step.
    // patch kernel: one A stage = a (TILE_H + ygroup - 1)-row halo patch serving `ygroup` vertical taps
    int ygroup, a_stages, w_stages;
    uint32_t a_part_bytes;          // bytes of one precision (hi or lo) of an A stage
    uint32_t tmem_cols;
    int a_parts;                    // 2: activations are (hi, lo) pairs, 3 MMAs per K step; 1: hi only, 2 MMAs (x_hi * w_hi + x_hi * w_lo)
    int b_parts;                    // 2: weights are (hi, lo) pairs; 1: hi only (w_lo == NULL): the x * w_lo MMA is dropped
    int merged_n;                   // pair kernel, 2-MMA convs with N = 64 | 128: the two weight planes are ONE B operand of
                                    // 2N rows (CTA 0 stages w_hi, CTA 1 stages w_lo), so x_hi * w_hi and x_hi * w_lo are one
                                    // MMA of N' = 2N whose columns [N, 2N) are exactly the lo accumulator (acc_lo_off = N).
                                    // Why: an N = 128 MMA reads 4 KB of A and 4 KB of B per 64 clocks = the whole 128 B/clk
                                    // of an SM's shared memory (67 % tensor-active with the TMA writes and the epilogue's
                                    // transposes on top, ncu r2g); merged, A is read once per K step instead of twice.
                                    // merged_n == 2 (3-MMA convs, N = 64: the encoders' full-resolution layers, 192 B/clk
                                    // of operand reads unmerged): the weight stage of a CTA already is [w_hi half | w_lo
                                    // half], so x_hi against all 64 rows of it is one MMA of N' = 128 whose columns come
                                    // out as [main(0:32) | lo(0:32) | main(32:64) | lo(32:64)]; x_lo * w_hi goes to a third
                                    // accumulator at column 128; the epilogue adds the three.
    int w_resident;                 // pair kernel: the whole filter stays in the W ring (loaded once per CTA)
    uint32_t a_tx_bytes;            // pair kernel: bytes one CTA's TMA loads deliver per A stage (hi + lo boxes)
    int patch_rows;                 // pair kernel, x-major patch: RY = TILE_H + kh - 1 (shared-memory row = x * RY + y)
    uint32_t epi_sleep_ns;          // back-off of the epilogue warps' wait for an accumulator (0 = plain polling)
    dkt_epilogue epi;
};

// how the epilogue warps of a CTA walk their share of the tiles
struct TileWalk {
    int first, step, items;     // work items first, first + step, ... < items
    int mul, off;               // tile = item * mul + off (pair kernel: 2 * item + cluster rank; may be >= num_tiles)
    uint32_t empty_remote;      // shared::cluster address of the LEADER's tmem_empty_bar[0]; 0 = arrive locally
};

// Epilogue operands (context term, gates, hidden state, residual) go through the read-only path: every element is
// read by the one thread that later stores to it (h' may overwrite h in place), never by another, so a
// non-coherent line can only ever be stale in elements nobody reads again.  What it buys: the loads no longer alias
// the stores as far as the compiler knows, so all of a chunk's loads are in flight before the first store.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// compile-time activation (the runtime switch of apply_act gets if-converted into ~50 issued instructions per value)
template <int ACT>
__device__ __forceinline__ float act_ct(float x) {
    if constexpr (ACT == DKT_ACT_RELU) return fmaxf(x, 0.0f);
    else if constexpr (ACT == DKT_ACT_SIGMOID) return sigmoidf_acc(x);
    else if constexpr (ACT == DKT_ACT_TANH) return tanhf(x);
    else if constexpr (ACT == DKT_ACT_LEAKY) return x > 0.f ? x : 0.01f * x;
    else return x;
}

// scalar path of a partially valid 4-group of a LINEAR conv whose N is not a multiple of 4 (rare shapes only)
template <int ACT>
__device__ __noinline__ void tc_epilogue1(const dkt_epilogue& e, int64_t p, int n, float a) {
    if (e.bias) a += e.bias[n];
    if (e.ctx) a += e.ctx[p * e.ctx_C + e.ctx_c0 + n];
    a = act_ct<ACT>(a) * e.scale;
    if (e.res) a = fmaxf(a + e.res[p * e.res_C + e.res_c0 + n], 0.f);
    else if (e.res_hi) {
        const int64_t off = p * e.res_C + e.res_c0 + n;
        a = fmaxf(a + unpack16(e.res_hi[off]) + unpack16(e.res_lo[off]), 0.f);
    }
    store_all(e.out, p, n, a);
}

// ---------------------------------------------------------------------------------------------
// v2: persistent, double-buffered accumulators, coalesced epilogue
// ---------------------------------------------------------------------------------------------
// 12 warps = three warpgroups: warps 0 (TMA producer) and 1 (MMA issuer) + two idle warps, then the eight epilogue warps as
// two whole warpgroups -- `setmaxnreg` moves registers between warpGROUPS, and the GRU epilogues spilled 50 - 70 registers
// at the 168 a 10- or 12-warp CTA gets per thread (cuobjdump: 105 LDL + 65 STL in the GRU_Q instantiation)
constexpr int TC2_THREADS = 384;
constexpr int TC2_EPI_WARPS = 8;
constexpr int TC2_EPI_WARP0 = 4;        // first epilogue warp
constexpr int TC2_ROLE_REGS = 56, TC2_EPI_REGS = 224;      // 128 * 56 + 256 * 224 = 64512 = 384 * 168

template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr uint32_t TC2_EPI_TILE_BYTES = TC2_EPI_WARPS * 32 * 128;   // one 32 px x 32 ch fp32 tile per warp
constexpr uint32_t TC2_EPI_BYTES = TC2_EPI_TILE_BYTES + 1024;        // + the conv bias (256 floats, zero beyond N)

template <int BK>
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t addr) {
    constexpr uint32_t ROW = BK * 2;                 // bytes per row (64 or 128) == swizzle span
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * ROW) >> 4) << 32;           // stride between 8-row atoms
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(BK == 64 ? 2 : 4) << 61;         // SWIZZLE_128B : SWIZZLE_64B
    return d;
}

// The role loops of warps 0 (TMA) and 1 (MMA) are executed by the WHOLE warp with one elected lane issuing:
// warp-uniform control flow lets ptxas keep descriptors, coordinates and barrier addresses in uniform registers.
// (Inside `if (lane == 0)` every UTCHMMA / UTMALDG was wrapped in an ELECT + 5x R2UR.BROADCAST retry loop, which
// bounded the issue rate -- and with it the tensor pipe -- of every conv with N <= 128: ncu r01h.)
// same with an explicit stride between 8-row atoms (x-major halo patch: RY rows instead of 8)
template <int BK>
__device__ __forceinline__ uint64_t smem_desc_kmajor_sbo(uint32_t addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(BK == 64 ? 2 : 4) << 61;
    return d;
}

__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// Epilogue warps 4..11 of the persistent kernels (see the header comment of this file).
// TMEM lane quarter = warp % 4; the two warps of a quarter take alternate 32-column chunks.  Per chunk a warp
// pulls 32 columns with tcgen05.ld (thread = pixel), transposes through a swizzled 4 KB smem buffer and continues
// with lanes = 8 x 16-byte channel groups of 4 pixels, so every global access is a full 128-byte line.
//
// This loop, not the tensor pipe, bounds every conv whose MMA work per tile is small (N <= 64, or few K steps), and
// what it is bound by is ISSUE SLOTS: ncu r02b counted 2028 executed warp instructions per 32 x 32 chunk of which
// ~400 were the actual math / loads / stores and ~1300 integer address arithmetic, predicates and branches of the
// run-time-generic code.  It is therefore specialised at compile time three ways:
//   KIND, ACT   epilogue kind and activation
//   FL          which optional operands / outputs exist (EPF_* bits); EPF_GENERIC keeps the run-time tests
//   CHECK       per-pixel bounds tests only for tiles that cross the image border (chosen per tile, uniform)
// and every pointer of a chunk is formed once (64-bit) with 32-bit element offsets per pixel.
enum : int {
    EPF_GENERIC = -1,
    EPF_OUT_F32 = 1, EPF_OUT_SPLIT = 2, EPF_RES_F32 = 4, EPF_RES_SPLIT = 8, EPF_CTX = 16, EPF_STATS = 32, EPF_TAIL = 64,
    EPF_OUT_HI = 128,        // 16-bit hi plane only (the destination feeds 2-MMA convs); EPF_OUT_SPLIT = hi and lo
};

template <int FL> __device__ __forceinline__ bool epf(int bit, bool runtime) { return FL < 0 ? runtime : (FL & bit) != 0; }

// XM ("x-major tile"): how an M row maps to a pixel of the 8 x 16 tile.  false: m = y * 16 + x (row-patch and per-tap
// kernels).  true: m = x * 8 + y (pair kernel with the x-major halo patch, where all taps of a filter read ONE patch).
// A warp's 32 M rows (one TMEM lane quarter q) are then a 2 x 16 / an 8 x 4 pixel group; lane (sub, jg) handles the
// group's pixels r = i*4 + sub, i < 8.
template <bool XM> __device__ __forceinline__ int epi_gy(int q) { return XM ? 0 : 2 * q; }
template <bool XM> __device__ __forceinline__ int epi_gx(int q) { return XM ? 4 * q : 0; }
template <bool XM> __device__ __forceinline__ int epi_dy(int i, int sub) { return XM ? ((i & 1) << 2) + sub : (i >> 2); }
template <bool XM> __device__ __forceinline__ int epi_dx(int i, int sub) { return XM ? (i >> 1) : ((i & 3) << 2) + sub; }

template <int KIND, int ACT, int FL, bool XM>
__device__ __forceinline__ void conv_tc_epilogue_warps(const TcConvParams& prm, uint32_t tmem_base,
                                                       uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar,
                                                       uint8_t* epi_smem, int warp, int lane, int tiles_per_img,
                                                       const TileWalk tw) {
    constexpr bool LIN = KIND == DKT_EPI_LINEAR;
    const int ew = warp - TC2_EPI_WARP0;
    const int q = warp & 3;
    const int half = ew >> 2;
    float* ebuf = reinterpret_cast<float*>(epi_smem + (size_t)ew * 4096);
    const dkt_epilogue& e = prm.epi;
    const int N = prm.N, H = prm.H, W = prm.W;
    const int sub = lane >> 3;           // pixel within a group of 4
    const int jg = lane & 7;             // 16-byte channel group within the 32-column chunk
    // ---- what exists (compile time unless FL is EPF_GENERIC) ----
    const bool has_ctx = LIN ? epf<FL>(EPF_CTX, e.ctx != nullptr) : true;
    const bool has_res = LIN && epf<FL>(EPF_RES_F32, e.res != nullptr);
    const bool has_res2 = LIN && !has_res && epf<FL>(EPF_RES_SPLIT, e.res_hi != nullptr);
    const bool has_tail = LIN && epf<FL>(EPF_TAIL, e.tail != nullptr);
    const bool has_stats = LIN && epf<FL>(EPF_STATS, e.stats_partial != nullptr);
    const bool out_f32 = epf<FL>(EPF_OUT_F32, e.out.f32 != nullptr);
    const bool out_split = epf<FL>(EPF_OUT_SPLIT | EPF_OUT_HI, e.out.hi != nullptr);
    const bool out_lo = epf<FL>(EPF_OUT_SPLIT, e.out.lo != nullptr);
    // ---- loop-invariant parameters ----
    const float* const ctx = e.ctx;
    const float* const res = e.res;
    const uint16_t* const res_hi = e.res_hi;
    const uint16_t* const res_lo = e.res_lo;
    const float* const tail = e.tail;
    const int tail_C = e.tail_C;
    const float scale = e.scale;
    float* const o_f32 = e.out.f32;
    uint16_t* const o_hi = e.out.hi;
    uint16_t* const o_lo = e.out.lo;
    const int oC = e.out.C, oc0 = e.out.c_begin;
    const int Nh = N >> 1;
    // a tail (e.g. the flow field appended to the motion features) that completes the last 4-group of an
    // N % 4 != 0 conv is merged into that group's vector store; otherwise it is copied after the chunk loop
    const bool tail_merged = has_tail && (N & 3) && (((N + tail_C) & 3) == 0) && tail_C < 4;
    const int Nvec = tail_merged ? N + tail_C : N;
    float* const stats_out = e.stats_partial;
    // the bias lives in shared memory: a global load here would sit in the latency chain of every chunk
    float* const s_bias = reinterpret_cast<float*>(epi_smem + TC2_EPI_TILE_BYTES);
    if (LIN) {
        for (int i = (int)threadIdx.x - TC2_EPI_WARP0 * 32; i < 256; i += TC2_EPI_WARPS * 32) s_bias[i] = (i < N && e.bias) ? __ldg(e.bias + i) : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const bool has_ops = !LIN || has_ctx || has_res || has_res2;     // operands fetched from HBM per pixel
    // accumulator stages: the MMA warp may run acc_stages - 1 tiles ahead of this drain
    const uint32_t nacc = prm.acc_stages;
    uint32_t t = 0, as = 0, aphase = 0;
    for (int item = tw.first; item < tw.items; item += tw.step, ++t) {
        const int tile = item * tw.mul + tw.off;
        const bool live = tile < prm.num_tiles;          // the odd tile count's filler of a CTA pair is drained, not stored
        const int b = tile / tiles_per_img;
        const int r = tile - b * tiles_per_img;
        const int ty = r / prm.tiles_x;
        const int y0 = ty * TC_TILE_H, x0 = (r - ty * prm.tiles_x) * TC_TILE_W;
        // this warp's pixel group starts at (gy0, gx0); lane pixel i sits at (+epi_dy(i), +epi_dx(i))
        const int gy0 = y0 + epi_gy<XM>(q), gx0 = x0 + epi_gx<XM>(q);
        const int64_t p00 = ((int64_t)b * H + gy0) * W + gx0;
        auto pix_ok = [&](int i) { return (gy0 + epi_dy<XM>(i, sub)) < H && (gx0 + epi_dx<XM>(i, sub)) < W; };
        auto pix_d = [&](int i) { return epi_dy<XM>(i, sub) * W + epi_dx<XM>(i, sub); };    // pixel delta from p00
        const bool interior = (y0 + TC_TILE_H <= H) && (x0 + TC_TILE_W <= W);   // uniform: no per-pixel bounds tests needed
        // L2 prefetch of the operand lines (context / z / h / residual) of a chunk; one lane per pixel row covers the
        // 128-byte line.  The first chunk of the NEXT tile is requested now (the MMA runs ahead, so this warp does not
        // wait and a same-tile prefetch has no lead time), chunk k+1 of this tile at the start of chunk k.
        auto prefetch_chunk = [&](int64_t pp00, int pgy0, int pgx0, int c0p) {
            if (jg != 0 || c0p >= prm.Npad) return;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (!((pgy0 + epi_dy<XM>(i, sub)) < H && (pgx0 + epi_dx<XM>(i, sub)) < W)) continue;
                const int64_t p = pp00 + epi_dy<XM>(i, sub) * W + epi_dx<XM>(i, sub);
                if (LIN) {
                    if (has_ctx) prefetch_l2(ctx + p * e.ctx_C + e.ctx_c0 + c0p);
                    if (has_res) prefetch_l2(res + p * e.res_C + e.res_c0 + c0p);
                    if (has_res2) {
                        prefetch_l2(res_hi + p * e.res_C + e.res_c0 + c0p);
                        prefetch_l2(res_lo + p * e.res_C + e.res_c0 + c0p);
                    }
                } else {
                    prefetch_l2(ctx + p * e.ctx_C + e.ctx_c0 + c0p);
                    if (KIND == DKT_EPI_GRU_Q) {
                        prefetch_l2(e.z.f32 + p * e.z.C + e.z.c_begin + c0p);
                        prefetch_l2(e.h.f32 + p * e.h.C + e.h.c_begin + c0p);
                    } else if (c0p >= Nh) {
                        prefetch_l2(e.h.f32 + p * e.h.C + e.h.c_begin + (c0p - Nh));
                    }
                }
            }
        };
        if (has_ops) {
            if (t == 0 && live) prefetch_chunk(p00, gy0, gx0, half * 32);
            const int ntile = (item + tw.step) * tw.mul + tw.off;
            if (item + tw.step < tw.items && ntile < prm.num_tiles) {
                const int nb = ntile / tiles_per_img;
                const int nr = ntile - nb * tiles_per_img;
                const int nty = nr / prm.tiles_x;
                const int ny0 = nty * TC_TILE_H, nx0 = (nr - nty * prm.tiles_x) * TC_TILE_W;
                const int ngy0 = ny0 + epi_gy<XM>(q), ngx0 = nx0 + epi_gx<XM>(q);
                prefetch_chunk(((int64_t)nb * H + ngy0) * W + ngx0, ngy0, ngx0, half * 32);
            }
        }
        mbar_wait_backoff(&tmem_full_bar[as], aphase, prm.epi_sleep_ns);
        tcgen05_fence_after();
        const uint32_t tbase = tmem_base + as * prm.acc_cols + ((uint32_t)(q * 32) << 16);
        for (int c0 = half * 32; live && c0 < prm.Npad; c0 += 64) {
            float v[32];
            const int ncols = (prm.Npad - c0 >= 32) ? 32 : 16;
            const int n = c0 + 4 * jg;
            if (has_ops) prefetch_chunk(p00, gy0, gx0, c0 + 64);
            // residual of this lane's 8 pixels, requested BEFORE the accumulator is pulled and transposed so that its
            // latency hides behind that work; kept as raw bits (fp32 x4, or bf16 hi x4 | lo x4) until first use
            uint4 rraw[8];
            if (has_res || has_res2) {
                const int64_t rbase = p00 * e.res_C + e.res_c0 + n;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    rraw[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (n + 3 < N && (interior || pix_ok(i))) {
                        const int64_t off = rbase + (int64_t)(pix_d(i) * e.res_C);
                        if (has_res) {
                            rraw[i] = __ldg(reinterpret_cast<const uint4*>(res + off));
                        } else {
                            const uint2 h2 = __ldg(reinterpret_cast<const uint2*>(res_hi + off));
                            const uint2 l2 = __ldg(reinterpret_cast<const uint2*>(res_lo + off));
                            rraw[i] = make_uint4(h2.x, h2.y, l2.x, l2.y);
                        }
                    }
                }
            }
            __syncwarp();                       // tcgen05.ld is .sync.aligned; also: previous chunk's reads done
            if (prm.merged_n == 2) {            // N = 64: chunk c0 in {0, 32}: main at 2*c0, x_hi*w_lo at 2*c0 + 32, x_lo*w_hi at 128 + c0
                float v2[32], v3[32];
                tmem_ld32(tbase + 2 * c0, v);
                tmem_ld32(tbase + 2 * c0 + 32, v2);
                tmem_ld32(tbase + 128 + c0, v3);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += v2[j] + v3[j];
            } else if (ncols == 32) {
                tmem_ld32(tbase + c0, v);
            } else {
                tmem_ld16(tbase + c0, v);
#pragma unroll
                for (int j = 16; j < 32; ++j) v[j] = 0.f;
            }
            if (prm.acc_lo_off) {               // main + lo accumulator, added here in fp32 round-to-nearest
                float v2[32];
                if (ncols == 32) {
                    tmem_ld32(tbase + prm.acc_lo_off + c0, v2);
                } else {
                    tmem_ld16(tbase + prm.acc_lo_off + c0, v2);
#pragma unroll
                    for (int j = 16; j < 32; ++j) v2[j] = 0.f;
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += v2[j];
            }
            tmem_ld_wait();
            // thread = pixel `lane`: row of 8 x 16 B, chunk j stored at j ^ (lane & 7)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(ebuf + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool vec = (n + 3 < Nvec);
            if (n >= Nvec) {
                // padded columns (lane-divergent only in the last chunk): nothing to store
            } else if (!LIN || vec) {
                // ---------------- vector path: 4 channels x 8 pixels per lane ----------------
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                int ntail = 0;                   // components of this group that come from the tail
                if (LIN) {
                    bv = *reinterpret_cast<const float4*>(s_bias + n);      // zero beyond N
                    if (has_tail && n + 3 >= N) ntail = n + 4 - N;          // the merged-tail group
                }
                // chunk-constant 64-bit bases; per pixel only a 32-bit element offset is added
                // (LINEAR only: the GRU kinds hide their epilogue behind long MMA phases and are register-bound instead,
                // so they form each address per pixel from the few values they keep live)
                const int nout = (KIND == DKT_EPI_GRU_ZR) ? n - Nh : n;
                const int64_t obase = p00 * oC + oc0 + nout;
                float* const pf = o_f32 + obase;
                uint16_t* const ph = o_hi + obase;
                uint16_t* const pl = o_lo + obase;
                const float* const pc = ctx + (p00 * e.ctx_C + e.ctx_c0 + n);
                const float* const pt = tail + p00 * tail_C;
                // two batches of 4 pixels: all operand loads of a batch are issued before its math and stores, so their
                // latencies overlap
                auto pixels = [&](auto check_tag) {
                    constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        float4 av[4], cv[4], zv[4], hv[4];
                        float tv[4][3];
                        bool ok[4];
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const int row = hb * 16 + i4 * 4 + sub;
                            ok[i4] = !CHECK || pix_ok(hb * 4 + i4);
                            if (!ok[i4]) continue;
                            const int d = pix_d(hb * 4 + i4);
                            av[i4] = *reinterpret_cast<const float4*>(ebuf + row * 32 + ((jg ^ (row & 7)) << 2));
                            if (LIN) {
                                if (has_ctx) cv[i4] = ld4(pc + d * e.ctx_C);
                                if (ntail) {
                                    const float* tp = pt + d * tail_C;
                                    tv[i4][0] = tp[0];
                                    if (ntail > 1) tv[i4][1] = tp[1];
                                    if (ntail > 2) tv[i4][2] = tp[2];
                                }
                            } else if (KIND == DKT_EPI_GRU_ZR) {
                                const int64_t p = p00 + d;
                                cv[i4] = ld4(ctx + p * e.ctx_C + e.ctx_c0 + n);
                                if (n >= Nh) hv[i4] = ld4(e.h.f32 + p * e.h.C + e.h.c_begin + (n - Nh));
                            } else {
                                const int64_t p = p00 + d;
                                cv[i4] = ld4(ctx + p * e.ctx_C + e.ctx_c0 + n);
                                zv[i4] = ld4(e.z.f32 + p * e.z.C + e.z.c_begin + n);
                                hv[i4] = ld4(e.h.f32 + p * e.h.C + e.h.c_begin + n);
                            }
                        }
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            if (!ok[i4]) continue;
                            const int d = pix_d(hb * 4 + i4);
                            float4 a = av[i4];
                            if (LIN) {
                                a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
                                if (has_ctx) { a.x += cv[i4].x; a.y += cv[i4].y; a.z += cv[i4].z; a.w += cv[i4].w; }
                                a.x = act_ct<ACT>(a.x) * scale; a.y = act_ct<ACT>(a.y) * scale;
                                a.z = act_ct<ACT>(a.z) * scale; a.w = act_ct<ACT>(a.w) * scale;
                                if (has_res || has_res2) {   // residual block tail: relu(x + y)
                                    const uint4 rr = rraw[hb * 4 + i4];
                                    float4 rv;
                                    if (has_res) {
                                        rv = make_float4(__uint_as_float(rr.x), __uint_as_float(rr.y), __uint_as_float(rr.z), __uint_as_float(rr.w));
                                    } else {
                                        const float2 h0 = unpack16x2(rr.x), h1 = unpack16x2(rr.y), l0 = unpack16x2(rr.z), l1 = unpack16x2(rr.w);
                                        rv = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
                                    }
                                    a.x = fmaxf(a.x + rv.x, 0.f); a.y = fmaxf(a.y + rv.y, 0.f);
                                    a.z = fmaxf(a.z + rv.z, 0.f); a.w = fmaxf(a.w + rv.w, 0.f);
                                }
                                if (ntail) {             // 1..3 trailing components are copied from the tail tensor
                                    if (ntail == 1) a.w = tv[i4][0];
                                    else if (ntail == 2) { a.z = tv[i4][0]; a.w = tv[i4][1]; }
                                    else { a.y = tv[i4][0]; a.z = tv[i4][1]; a.w = tv[i4][2]; }
                                }
                            } else if (KIND == DKT_EPI_GRU_ZR) {
                                a.x = sigmoidf_gate(a.x + cv[i4].x); a.y = sigmoidf_gate(a.y + cv[i4].y);
                                a.z = sigmoidf_gate(a.z + cv[i4].z); a.w = sigmoidf_gate(a.w + cv[i4].w);
                                if (n < Nh) {
                                    *reinterpret_cast<float4*>(e.z.f32 + (p00 + d) * e.z.C + e.z.c_begin + n) = a;
                                    continue;
                                }
                                a.x *= hv[i4].x; a.y *= hv[i4].y; a.z *= hv[i4].z; a.w *= hv[i4].w;
                            } else {                     // GRU_Q
                                const float4 z = zv[i4], h = hv[i4];
                                a.x = (1.f - z.x) * h.x + z.x * tanhf_gate(a.x + cv[i4].x);
                                a.y = (1.f - z.y) * h.y + z.y * tanhf_gate(a.y + cv[i4].y);
                                a.z = (1.f - z.z) * h.z + z.z * tanhf_gate(a.z + cv[i4].z);
                                a.w = (1.f - z.w) * h.w + z.w * tanhf_gate(a.w + cv[i4].w);
                            }
                            if (has_stats) {
                                ssum.x += a.x; ssum.y += a.y; ssum.z += a.z; ssum.w += a.w;
                                ssq.x = fmaf(a.x, a.x, ssq.x); ssq.y = fmaf(a.y, a.y, ssq.y);
                                ssq.z = fmaf(a.z, a.z, ssq.z); ssq.w = fmaf(a.w, a.w, ssq.w);
                            }
                            if (LIN) {
                                const int od = d * oC;
                                if (out_f32) *reinterpret_cast<float4*>(pf + od) = a;
                                if (out_split) {
                                    if (out_lo) {
                                        uint32_t h0, l0, h1, l1;
                                        split16x2(a.x, a.y, h0, l0);
                                        split16x2(a.z, a.w, h1, l1);
                                        *reinterpret_cast<uint2*>(ph + od) = make_uint2(h0, h1);
                                        *reinterpret_cast<uint2*>(pl + od) = make_uint2(l0, l1);
                                    } else {
                                        *reinterpret_cast<uint2*>(ph + od) = make_uint2(pack_hi16x2(a.x, a.y), pack_hi16x2(a.z, a.w));
                                    }
                                }
                            } else {
                                const int64_t off = (p00 + d) * oC + oc0 + nout;
                                if (out_f32) *reinterpret_cast<float4*>(o_f32 + off) = a;
                                if (out_split) {
                                    if (out_lo) {
                                        uint32_t h0, l0, h1, l1;
                                        split16x2(a.x, a.y, h0, l0);
                                        split16x2(a.z, a.w, h1, l1);
                                        *reinterpret_cast<uint2*>(o_hi + off) = make_uint2(h0, h1);
                                        *reinterpret_cast<uint2*>(o_lo + off) = make_uint2(l0, l1);
                                    } else {
                                        *reinterpret_cast<uint2*>(o_hi + off) = make_uint2(pack_hi16x2(a.x, a.y), pack_hi16x2(a.z, a.w));
                                    }
                                }
                            }
                        }
                    }
                };
                if (LIN && interior) pixels(std::false_type{});
                else pixels(std::true_type{});
            } else {
                // ---------------- scalar path: partially valid last group of a LINEAR conv (rare shapes) ----------------
                for (int i = 0; i < 8; ++i) {
                    const int row = i * 4 + sub;
                    if (!pix_ok(i)) continue;
                    const int64_t p = p00 + pix_d(i);
                    const float4 a = *reinterpret_cast<const float4*>(ebuf + row * 32 + ((jg ^ (row & 7)) << 2));
                    const float av1[4] = {a.x, a.y, a.z, a.w};
                    for (int u = 0; u < 4; ++u) if (n + u < N) tc_epilogue1<ACT>(e, p, n + u, av1[u]);
                }
            }
            if (has_stats) {
                // fixed-order reduction (deterministic): 8 pixels per lane above, then the warp's 4 pixel groups by
                // shuffles; the warp's 32-pixel sums go straight to HBM ([tile][quarter][sum, sumsq][N]) -- no cross-warp
                // exchange in the epilogue; quarters and tiles are added up by dkt_instnorm_finalize_tiles
                float rs[8] = {ssum.x, ssum.y, ssum.z, ssum.w, ssq.x, ssq.y, ssq.z, ssq.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    rs[u] += __shfl_xor_sync(0xffffffffu, rs[u], 8);
                    rs[u] += __shfl_xor_sync(0xffffffffu, rs[u], 16);
                }
                if (sub == 0 && n < N) {
                    float* sp = stats_out + (((int64_t)tile * 4 + q) * 2) * N + n;
                    *reinterpret_cast<float4*>(sp) = make_float4(rs[0], rs[1], rs[2], rs[3]);
                    *reinterpret_cast<float4*>(sp + N) = make_float4(rs[4], rs[5], rs[6], rs[7]);
                }
            }
        }
        if (live && has_tail && !tail_merged && half == 0) {
            const int m = q * 32 + lane;
            const int y = y0 + (XM ? (m & 7) : m / TC_TILE_W), x = x0 + (XM ? (m >> 3) : m % TC_TILE_W);
            if (y < H && x < W) {
                const int64_t p = ((int64_t)b * H + y) * W + x;
                for (int u = 0; u < tail_C; ++u) store_all(e.out, p, N + u, __ldg(tail + p * tail_C + u));
            }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (tw.empty_remote) mbar_arrive_cluster(tw.empty_remote + as * 8u);
            else mbar_arrive(&tmem_empty_bar[as]);
        }
        if (++as == nacc) { as = 0; aphase ^= 1u; }
    }
}

// PROJ epilogue (DKT_EPI_PROJ): y = act(acc + bias) stays on chip; each pixel's N values are contracted with a
// small fp32 matrix proj[N][12] (the 9 per-tap channel responses of the one-output 3x3 conv that follows) and only
// those 12 floats per pixel are stored.  Layout: thread = pixel (its TMEM lane), the two warps of a lane quarter
// take alternate 32-column chunks and meet through shared memory (double-buffered by tile parity) on a
// 64-thread named barrier.
constexpr int PROJ_T = 9;
template <int ACT, bool XM>
__device__ __forceinline__ void conv_tc_epilogue_proj(const TcConvParams& prm, uint32_t tmem_base,
                                                      uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar,
                                                      uint8_t* epi_smem, int warp, int lane, int tiles_per_img,
                                                      const TileWalk tw) {
    const int ew = warp - TC2_EPI_WARP0;
    const int q = warp & 3;
    const int half = ew >> 2;
    float* s_w = reinterpret_cast<float*>(epi_smem);            // [256][12]
    float* s_b = s_w + 256 * DKT_PROJ_LD;                        // [256]
    float* s_x = s_b + 256;                                      // [2][128][12]
    const dkt_epilogue& e = prm.epi;
    const int N = prm.N;
    const float scale = e.scale;
    const int et = (int)threadIdx.x - TC2_EPI_WARP0 * 32;
    for (int i = et; i < 256 * DKT_PROJ_LD; i += TC2_EPI_WARPS * 32) s_w[i] = (i < N * DKT_PROJ_LD) ? __ldg(e.proj + i) : 0.f;
    for (int i = et; i < 256; i += TC2_EPI_WARPS * 32) s_b[i] = (i < N && e.bias) ? __ldg(e.bias + i) : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t nacc = prm.acc_stages;
    uint32_t t = 0, as = 0, aphase = 0;
    for (int item = tw.first; item < tw.items; item += tw.step, ++t) {
        const int tile = item * tw.mul + tw.off;
        const bool live = tile < prm.num_tiles;
        const int b = tile / tiles_per_img;
        const int r = tile - b * tiles_per_img;
        const int y0 = (r / prm.tiles_x) * TC_TILE_H, x0 = (r % prm.tiles_x) * TC_TILE_W;
        mbar_wait_backoff(&tmem_full_bar[as], aphase, prm.epi_sleep_ns);
        tcgen05_fence_after();
        const uint32_t tbase = tmem_base + as * prm.acc_cols + ((uint32_t)(q * 32) << 16);
        float acc[PROJ_T];
#pragma unroll
        for (int u = 0; u < PROJ_T; ++u) acc[u] = 0.f;
        for (int c0 = half * 32; live && c0 < prm.Npad; c0 += 64) {
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
            for (int j = 0; j < 32; ++j) {
                const int n = c0 + j;                               // < 256: rows >= N of s_w are zero
                const float y = act_ct<ACT>(v[j] + s_b[n]) * scale;
                const float4 w0 = *reinterpret_cast<const float4*>(s_w + n * DKT_PROJ_LD);
                const float4 w1 = *reinterpret_cast<const float4*>(s_w + n * DKT_PROJ_LD + 4);
                const float w2 = s_w[n * DKT_PROJ_LD + 8];
                acc[0] = fmaf(y, w0.x, acc[0]); acc[1] = fmaf(y, w0.y, acc[1]);
                acc[2] = fmaf(y, w0.z, acc[2]); acc[3] = fmaf(y, w0.w, acc[3]);
                acc[4] = fmaf(y, w1.x, acc[4]); acc[5] = fmaf(y, w1.y, acc[5]);
                acc[6] = fmaf(y, w1.z, acc[6]); acc[7] = fmaf(y, w1.w, acc[7]);
                acc[8] = fmaf(y, w2, acc[8]);
            }
        }
        // accumulator drained: hand it back to the MMA warp before the (slow) exchange + store
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (tw.empty_remote) mbar_arrive_cluster(tw.empty_remote + as * 8u);
            else mbar_arrive(&tmem_empty_bar[as]);
        }
        if (++as == nacc) { as = 0; aphase ^= 1u; }
        if (!live) continue;                          // uniform over the CTA: both warps of a quarter skip the exchange
        float* xb = s_x + ((t & 1u) * 128 + q * 32 + lane) * DKT_PROJ_LD;
        if (half == 1) {
            *reinterpret_cast<float4*>(xb) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(xb + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            xb[8] = acc[8];
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        if (half == 0) {
            const float4 o0 = *reinterpret_cast<const float4*>(xb);
            const float4 o1 = *reinterpret_cast<const float4*>(xb + 4);
            const float o2 = xb[8];
            const int m = q * 32 + lane;
            const int y = y0 + (XM ? (m & 7) : m / TC_TILE_W), x = x0 + (XM ? (m >> 3) : m % TC_TILE_W);
            if (y < prm.H && x < prm.W) {
                const int64_t p = ((int64_t)b * prm.H + y) * prm.W + x;
                float* dst = e.out.f32 + p * e.out.C + e.out.c_begin;
                *reinterpret_cast<float4*>(dst) = make_float4(acc[0] + o0.x, acc[1] + o0.y, acc[2] + o0.z, acc[3] + o0.w);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4] + o1.x, acc[5] + o1.y, acc[6] + o1.z, acc[7] + o1.w);
                *reinterpret_cast<float4*>(dst + 8) = make_float4(acc[8] + o2, 0.f, 0.f, 0.f);
            }
        }
    }
}

template <int BK, int KIND, int ACT, int FL>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcConvParams prm) {
    constexpr bool XMT = false;                      // M row = y * 16 + x
    constexpr uint32_t ROW = BK * 2;
    constexpr uint32_t A_BYTES = 128 * ROW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const uint32_t b_bytes = (uint32_t)prm.Npad * ROW;
    const uint32_t a_bytes = (uint32_t)prm.a_parts * A_BYTES;         // hi [+ lo]
    const uint32_t stage_bytes = a_bytes + (uint32_t)prm.b_parts * b_bytes;
    uint8_t* epi_smem = smem + (size_t)prm.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + TC2_EPI_BYTES);
    uint64_t* empty_bar = full_bar + TC_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + TC_MAX_STAGES;     // [TC_MAX_ACC]
    uint64_t* tmem_empty_bar = tmem_full_bar + TC_MAX_ACC;   // [TC_MAX_ACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + TC_MAX_ACC);

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < prm.nsrc; ++s) { tma_prefetch_desc(&prm.act[s][0]); tma_prefetch_desc(&prm.act[s][1]); }
        tma_prefetch_desc(&prm.wgt[0]);
        tma_prefetch_desc(&prm.wgt[1]);
        for (int s = 0; s < prm.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < (int)prm.acc_stages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], TC2_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, prm.tmem_cols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = prm.tiles_x * prm.tiles_y;

    if (warp == 0) {
        // ===== TMA producer =====
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img;
            const int r = tile - b * tiles_per_img;
            const int y0 = (r / prm.tiles_x) * TC_TILE_H, x0 = (r % prm.tiles_x) * TC_TILE_W;
            for (int tap = 0; tap < prm.taps; ++tap) {
                const int ky = tap / prm.kw, kx = tap - ky * prm.kw;
                const int xs = x0 * prm.stride + kx - prm.pad_x, ys = y0 * prm.stride + ky - prm.pad_y;
                int kofs = 0;
                for (int s = 0; s < prm.nsrc; ++s) {
                    for (int kb = 0; kb < prm.kblocks[s]; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        if (elect_one()) {
                            uint8_t* st = smem + (size_t)stage * stage_bytes;
                            mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                            const int c = prm.c_begin[s] + kb * BK;
                            tma_load_4d(st, &prm.act[s][0], &full_bar[stage], c, xs, ys, b);
                            if (prm.a_parts == 2) tma_load_4d(st + A_BYTES, &prm.act[s][1], &full_bar[stage], c, xs, ys, b);
                            tma_load_2d(st + a_bytes, &prm.wgt[0], &full_bar[stage], kofs + kb * BK, tap * prm.Npad);
                            if (prm.b_parts == 2) tma_load_2d(st + a_bytes + b_bytes, &prm.wgt[1], &full_bar[stage], kofs + kb * BK, tap * prm.Npad);
                        }
                        if (++stage == prm.stages) { stage = 0; phase ^= 1u; }
                    }
                    kofs += prm.c_count[s];
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = idesc_bf16_m128((uint32_t)prm.Npad);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t as = 0, aphase = 0;
        for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);          // epilogue drained this accumulator
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + as * prm.acc_cols;
            int it = 0;
            for (int tap = 0; tap < prm.taps; ++tap)
            for (int s = 0; s < prm.nsrc; ++s)
            for (int kb = 0; kb < prm.kblocks[s]; ++kb, ++it) {
                const int kn = (kb == prm.kblocks[s] - 1) ? prm.klast[s] : BK / 16;   // K16 steps of this block
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t dah = smem_desc_kmajor<BK>(a_hi), dal = smem_desc_kmajor<BK>(a_hi + A_BYTES);
                    const uint64_t dwh = smem_desc_kmajor<BK>(a_hi + a_bytes), dwl = smem_desc_kmajor<BK>(a_hi + a_bytes + b_bytes);
                    const bool a_lo = prm.a_parts == 2, b_lo = prm.b_parts == 2;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {           // +32 bytes per K16 step = +2 in the address field
                        if (k >= kn) break;
                        umma_bf16(tmem_d, dah + 2 * k, dwh + 2 * k, idesc, (it | k) != 0);
                        if (a_lo) umma_bf16(tmem_d + prm.acc_lo_off, dal + 2 * k, dwh + 2 * k, idesc, (prm.acc_lo_off == 0 || (it | k) != 0) ? 1u : 0u);
                        if (b_lo) umma_bf16(tmem_d + prm.acc_lo_off, dah + 2 * k, dwl + 2 * k, idesc, (prm.acc_lo_off == 0 || a_lo || (it | k) != 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);               // smem slot reusable once these MMAs retire
                }
                if (++stage == prm.stages) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(&tmem_full_bar[as]);     // accumulator complete
            if (++as == prm.acc_stages) { as = 0; aphase ^= 1u; }
        }
    } else if (warp >= TC2_EPI_WARP0) {
        const TileWalk tw{(int)blockIdx.x, (int)gridDim.x, prm.num_tiles, 1, 0, 0u};
        if constexpr (KIND == DKT_EPI_PROJ)
            conv_tc_epilogue_proj<ACT, XMT>(prm, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane, tiles_per_img, tw);
        else
            conv_tc_epilogue_warps<KIND, ACT, FL, XMT>(prm, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane, tiles_per_img, tw);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, prm.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// v3 ("row patch"): same roles and epilogue as conv_tc_kernel, but the A operand of the `ygroup`
// vertical taps (ky) of a filter column kx comes from ONE TMA box of TILE_H + ygroup - 1 rows:
// with the 8 x 16 pixel tile, M row m = (my, mx) sits at byte (my*16 + mx) * 128 of the box, so tap ky
// is the same box read from byte offset ky * 2048 -- a 1024-byte-aligned start address, i.e. a
// standard SWIZZLE_128B K-major descriptor (no re-staging, no im2col).  A 3x3 conv thus pulls 3
// patches of 10 rows per 64-channel block instead of 9 tiles of 8 rows (2.4x less activation traffic
// from L2, which is what bounds these kernels: see DESIGN.md "K3 roofline"); the 7x1 stem pulls 1
// patch of 14 rows instead of 7 tiles.  Weights stream through their own ring in K blocks of WK
// channels (32 when N > 128 so that three stages fit).  Strided convs use ygroup = 1 (plain tiles).
// ---------------------------------------------------------------------------------------------
constexpr int TCP_MAX_A = 4, TCP_MAX_W = 9;

template <int WK, int KIND, int ACT, int FL = EPF_GENERIC>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc_patch_kernel(const __grid_constant__ TcConvParams prm) {
    constexpr bool XMT = false;                      // M row = y * 16 + x
    constexpr uint32_t WROW = WK * 2;
    constexpr int WSPLIT = 64 / WK;                   // W steps per 64-channel A block
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const uint32_t a_part = prm.a_part_bytes, a_stage_bytes = (uint32_t)prm.a_parts * prm.a_part_bytes;
    const uint32_t b_bytes = (uint32_t)prm.Npad * WROW, w_stage_bytes = (uint32_t)prm.b_parts * b_bytes;
    uint8_t* a_ring = smem;
    uint8_t* w_ring = a_ring + (size_t)prm.a_stages * a_stage_bytes;
    uint8_t* epi_smem = w_ring + (size_t)prm.w_stages * w_stage_bytes;
    uint64_t* afull = reinterpret_cast<uint64_t*>(epi_smem + TC2_EPI_BYTES);
    uint64_t* aempty = afull + TCP_MAX_A;
    uint64_t* wfull = aempty + TCP_MAX_A;
    uint64_t* wempty = wfull + TCP_MAX_W;
    uint64_t* tmem_full_bar = wempty + TCP_MAX_W;            // [TC_MAX_ACC]
    uint64_t* tmem_empty_bar = tmem_full_bar + TC_MAX_ACC;   // [TC_MAX_ACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + TC_MAX_ACC);

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int ygroups = prm.kh / prm.ygroup;                 // A steps per (kb, kx)

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < prm.nsrc; ++s) { tma_prefetch_desc(&prm.act[s][0]); tma_prefetch_desc(&prm.act[s][1]); }
        tma_prefetch_desc(&prm.wgt[0]);
        tma_prefetch_desc(&prm.wgt[1]);
        for (int s = 0; s < prm.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < prm.w_stages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int s = 0; s < (int)prm.acc_stages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], TC2_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, prm.tmem_cols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = prm.tiles_x * prm.tiles_y;

    if (warp == 0) {
        // ===== TMA producer: A patches and W tiles in the order the MMA warp consumes them =====
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0;
        for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img;
            const int r = tile - b * tiles_per_img;
            const int y0 = (r / prm.tiles_x) * TC_TILE_H, x0 = (r % prm.tiles_x) * TC_TILE_W;
            int kofs = 0;
            for (int s = 0; s < prm.nsrc; ++s) {
                for (int kb = 0; kb < prm.kblocks[s]; ++kb) {
                    const int c = prm.c_begin[s] + kb * 64;
                    const int kn = (kb == prm.kblocks[s] - 1) ? prm.klast[s] : 4;     // K16 steps of this block
                    const int wsn = (kn * 16 + WK - 1) / WK;                          // weight steps that hold them
                    for (int kx = 0; kx < prm.kw; ++kx) {
                        const int xs = x0 * prm.stride + kx - prm.pad_x;
                        for (int yg = 0; yg < ygroups; ++yg) {
                            const int ys = y0 * prm.stride + yg * prm.ygroup - prm.pad_y;
                            mbar_wait(&aempty[as], aph ^ 1u);
                            if (elect_one()) {
                                uint8_t* ast = a_ring + (size_t)as * a_stage_bytes;
                                mbar_arrive_expect_tx(&afull[as], a_stage_bytes);
                                tma_load_4d(ast, &prm.act[s][0], &afull[as], c, xs, ys, b);
                                if (prm.a_parts == 2) tma_load_4d(ast + a_part, &prm.act[s][1], &afull[as], c, xs, ys, b);
                            }
                            if (++as == prm.a_stages) { as = 0; aph ^= 1u; }
                            for (int kyi = 0; kyi < prm.ygroup; ++kyi) {
                                const int tap = (yg * prm.ygroup + kyi) * prm.kw + kx;
                                for (int wh = 0; wh < wsn; ++wh) {
                                    mbar_wait(&wempty[ws], wph ^ 1u);
                                    if (elect_one()) {
                                        uint8_t* wst = w_ring + (size_t)ws * w_stage_bytes;
                                        mbar_arrive_expect_tx(&wfull[ws], w_stage_bytes);
                                        const int kc = kofs + kb * 64 + wh * WK;
                                        tma_load_2d(wst, &prm.wgt[0], &wfull[ws], kc, tap * prm.Npad);
                                        if (prm.b_parts == 2) tma_load_2d(wst + b_bytes, &prm.wgt[1], &wfull[ws], kc, tap * prm.Npad);
                                    }
                                    if (++ws == prm.w_stages) { ws = 0; wph ^= 1u; }
                                }
                            }
                        }
                    }
                }
                kofs += prm.c_count[s];
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = idesc_bf16_m128((uint32_t)prm.Npad);
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0, acs = 0, aphase = 0;
        const int steps_per_kb = prm.kw * ygroups;               // A steps per K block
        for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty_bar[acs], aphase ^ 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acs * prm.acc_cols;
            uint32_t accumulate = 0;
            for (int s = 0; s < prm.nsrc; ++s)
            for (int kb = 0; kb < prm.kblocks[s]; ++kb)
            for (int ai = 0; ai < steps_per_kb; ++ai) {
                const int kn = (kb == prm.kblocks[s] - 1) ? prm.klast[s] : 4;         // K16 steps of this block
                const int wsn = (kn * 16 + WK - 1) / WK;
                mbar_wait(&afull[as], aph);
                tcgen05_fence_after();
                const uint32_t a_hi0 = smem_u32(a_ring + (size_t)as * a_stage_bytes);
                for (int kyi = 0; kyi < prm.ygroup; ++kyi) {
                    const uint32_t a_hi = a_hi0 + (uint32_t)kyi * (TC_TILE_W * 128u);
#pragma unroll
                    for (int wh = 0; wh < WSPLIT; ++wh) {
                        if (wh >= wsn) break;
                        mbar_wait(&wfull[ws], wph);
                        tcgen05_fence_after();
                        if (elect_one()) {
                            const uint32_t w_hi = smem_u32(w_ring + (size_t)ws * w_stage_bytes);
                            const uint64_t dah = smem_desc_kmajor<64>(a_hi + wh * WK * 2), dal = smem_desc_kmajor<64>(a_hi + a_part + wh * WK * 2);
                            const uint64_t dwh = smem_desc_kmajor<WK>(w_hi), dwl = smem_desc_kmajor<WK>(w_hi + b_bytes);
#pragma unroll
                            for (int k = 0; k < WK / 16; ++k) {   // +32 bytes per K16 step = +2 in the address field
                                if (wh * (WK / 16) + k >= kn) break;
                                const uint32_t acc_l = prm.acc_lo_off ? accumulate : 1u;      // first lo MMA of a tile overwrites
                                umma_bf16(tmem_d, dah + 2 * k, dwh + 2 * k, idesc, accumulate);
                                if (prm.a_parts == 2) umma_bf16(tmem_d + prm.acc_lo_off, dal + 2 * k, dwh + 2 * k, idesc, acc_l);
                                if (prm.b_parts == 2) umma_bf16(tmem_d + prm.acc_lo_off, dah + 2 * k, dwl + 2 * k, idesc, prm.a_parts == 2 ? 1u : acc_l);
                                accumulate = 1u;
                            }
                            umma_commit(&wempty[ws]);
                        }
                        accumulate = 1u;
                        if (++ws == prm.w_stages) { ws = 0; wph ^= 1u; }
                    }
                }
                if (elect_one()) umma_commit(&aempty[as]);         // patch reusable once every tap's MMAs retire
                if (++as == prm.a_stages) { as = 0; aph ^= 1u; }
            }
            if (elect_one()) umma_commit(&tmem_full_bar[acs]);
            if (++acs == prm.acc_stages) { acs = 0; aphase ^= 1u; }
        }
    } else if (warp >= TC2_EPI_WARP0) {
        const TileWalk tw{(int)blockIdx.x, (int)gridDim.x, prm.num_tiles, 1, 0, 0u};
        if constexpr (KIND == DKT_EPI_PROJ)
            conv_tc_epilogue_proj<ACT, XMT>(prm, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane, tiles_per_img, tw);
        else
            conv_tc_epilogue_warps<KIND, ACT, FL, XMT>(prm, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane, tiles_per_img, tw);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, prm.tmem_cols);
}

// MMA issue loop of the CTA-pair kernel (warp 1 of the leader).  ncu r4a (source page of the full-resolution 64 -> 64
// encoder convs): at N <= 128 a tap's MMAs take ~400 clocks of tensor time but the single issuing warp needed ~780 clocks
// per tap -- ~150 dependent instructions: an integer division for (ky, kx), descriptor words rebuilt from byte addresses,
// run-time operand-mode tests around every UTCHMMA, constant-bank reloads -- and the tensor pipe sat at 50 %.  Here the
// operand mode (MERGED = TcConvParams::merged_n, A_LO / B_LO = which lo planes exist) is a template parameter, every
// descriptor is a constant high word + a running 14-bit low word (address >> 4) advanced by additions only, the tap's
// patch offset is counted, not divided, and full K blocks run without the partial-block test.
struct PairMmaCtx {
    uint8_t* a_ring; uint8_t* w_ring;
    uint64_t* afull; uint64_t* aempty; uint64_t* wfull; uint64_t* wempty; uint64_t* tmem_full_bar; uint64_t* tmem_empty_bar;
    uint32_t tmem_base;
    int pair_id, pairs, items;
};

__device__ __forceinline__ uint64_t desc_words(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}

template <int KB, bool XMT, int MERGED, bool A_LO, bool B_LO>
__device__ __forceinline__ void pair_mma_loop(const TcConvParams& prm, const PairMmaCtx& cx) {
    constexpr uint32_t KROW = KB * 2;
    constexpr uint32_t KSTEPS = KB / 16;
    const uint32_t idesc = idesc_bf16_m256((uint32_t)(MERGED ? 2 * prm.Npad : prm.Npad));
    const uint32_t idesc_n = idesc_bf16_m256((uint32_t)prm.Npad);          // MERGED == 2: the x_lo * w_hi MMA
    const int ntaps = prm.ygroup;                                          // taps served by one A stage
    const int kw = prm.kw;
    const int steps_per_kb = XMT ? 1 : prm.kw * (prm.kh / prm.ygroup);     // A steps per K block
    const int a_stages = prm.a_stages, w_stages = prm.w_stages, nsrc = prm.nsrc;
    const uint32_t acc_stages = prm.acc_stages, acc_cols = prm.acc_cols, acc_lo_off = prm.acc_lo_off;
    const bool w_stream = !prm.w_resident;
    // descriptor words: low = (address >> 4) | 1 << 16, high = atom stride | version | swizzle (constant per operand)
    const uint32_t sbo = XMT ? (uint32_t)prm.patch_rows * KROW : 8u * KROW;
    const uint32_t da_hi = (uint32_t)(smem_desc_kmajor_sbo<KB>(0u, sbo) >> 32), dw_hi = (uint32_t)(smem_desc_kmajor<KB>(0u) >> 32);
    const uint32_t a_base16 = ((smem_u32(cx.a_ring) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_base16 = ((smem_u32(cx.w_ring) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_part16 = prm.a_part_bytes >> 4, a_stage16 = ((uint32_t)prm.a_parts * prm.a_part_bytes) >> 4;
    const uint32_t b16 = ((uint32_t)(prm.Npad >> 1) * KROW) >> 4, w_stage16 = (uint32_t)prm.b_parts * b16;
    const uint32_t tapx16 = XMT ? ((uint32_t)prm.patch_rows * KROW) >> 4 : (TC_TILE_W * KROW) >> 4;   // next tap in the row (XMT: kx + 1)
    constexpr uint32_t tapy16 = KROW >> 4;                                                           // XMT: ky + 1
    int as = 0, ws = 0;
    uint32_t aph = 0, wph = 0, acs = 0, aphase = 0;
    uint32_t a16 = a_base16, w16 = w_base16;             // low descriptor words of the current A / W stage
    bool first = true;
    for (int item = cx.pair_id; item < cx.items; item += cx.pairs) {
        mbar_wait(&cx.tmem_empty_bar[acs], aphase ^ 1u);     // both CTAs' epilogues drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = cx.tmem_base + acs * acc_cols;
        const uint32_t tmem_lo = tmem_d + acc_lo_off;
        uint32_t accumulate = 0;
        for (int s = 0; s < nsrc; ++s) {
            const int kblocks = prm.kblocks[s], klast = prm.klast[s];
            for (int kb = 0; kb < kblocks; ++kb) {
                const bool full = (kb != kblocks - 1) || klast == (int)KSTEPS;
                for (int ai = 0; ai < steps_per_kb; ++ai) {
                    mbar_wait(&cx.afull[as], aph);
                    tcgen05_fence_after();
                    uint32_t t16 = a16, row16 = a16;     // this tap's patch start; start of the tap row (XMT)
                    int kx = 0;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        if (w_stream || first) {
                            mbar_wait(&cx.wfull[ws], wph);
                            tcgen05_fence_after();
                        }
                        if (elect_one()) {
                            auto step = [&](uint32_t k2, uint32_t acc) {
                                const uint64_t dah = desc_words(t16 + k2, da_hi), dwh = desc_words(w16 + k2, dw_hi);
                                umma_bf16_pair(tmem_d, dah, dwh, idesc, acc);
                                if (MERGED == 2) {
                                    umma_bf16_pair(tmem_d + 128u, desc_words(t16 + a_part16 + k2, da_hi), dwh, idesc_n, acc);
                                } else if (MERGED == 0) {
                                    const uint32_t acc_l = acc_lo_off ? acc : 1u;      // first lo MMA of a tile overwrites
                                    if (A_LO) umma_bf16_pair(tmem_lo, desc_words(t16 + a_part16 + k2, da_hi), dwh, idesc, acc_l);
                                    if (B_LO) umma_bf16_pair(tmem_lo, dah, desc_words(w16 + b16 + k2, dw_hi), idesc, A_LO ? 1u : acc_l);
                                }
                            };
                            if (full) {                      // +32 bytes per K16 step = +2 in the address field
                                step(0u, accumulate);
#pragma unroll
                                for (uint32_t k = 1; k < KSTEPS; ++k) step(2u * k, 1u);
                            } else {
                                step(0u, accumulate);
#pragma unroll
                                for (uint32_t k = 1; k < KSTEPS; ++k) {
                                    if ((int)k >= klast) break;
                                    step(2u * k, 1u);
                                }
                            }
                            if (w_stream) umma_commit_pair(&cx.wempty[ws]);
                        }
                        accumulate = 1u;
                        w16 += w_stage16;
                        if (++ws == w_stages) { ws = 0; wph ^= 1u; w16 = w_base16; }
                        if (XMT) {                           // tap (ky, kx) = the patch read from row kx * RY + ky
                            t16 += tapx16;
                            if (++kx == kw) { kx = 0; row16 += tapy16; t16 = row16; }
                        } else {
                            t16 += tapx16;
                        }
                    }
                    if (elect_one()) umma_commit_pair(&cx.aempty[as]);   // both CTAs' patches reusable once every tap's MMAs retire
                    a16 += a_stage16;
                    if (++as == a_stages) { as = 0; aph ^= 1u; a16 = a_base16; }
                }
            }
        }
        if (elect_one()) umma_commit_pair(&cx.tmem_full_bar[acs]);
        if (++acs == acc_stages) { acs = 0; aphase ^= 1u; }
        first = false;
    }
}

// ---------------------------------------------------------------------------------------------
// v4 ("CTA pair"): the row-patch kernel on tcgen05 cta_group::2.  Two CTAs of a cluster (one TPC) work on two
// tiles at once as ONE M = 256 MMA: each CTA stages the halo patches of its own 128-pixel tile and only HALF of
// every weight block (rows [rank * Npad/2, (rank+1) * Npad/2) of the N x 64 tile); the tensor core of the pair
// reads A from both CTAs and the two B halves from both CTAs.  Why: after the epilogue was specialised every
// conv of the step sat at 8.5 - 10.5 TB/s of L2 -> SM traffic (ncu r01g, l1tex__m_xbar2l1tex_read_bytes), the
// cap of the L2 fabric, and the weight stream was most of it (3.5 MB of 4.3 MB per tile for the gru08 gates).
// KB = channels per K block: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows; the 7x7 stems, whose x-im2col
// rows carry 21 / 14 real channels).  Halving the weight bytes per SM also doubles the K depth a weight stage holds; filters whose half fits
// (3x3 / 7x1 at 64 -> 64) stay resident in the ring for the whole launch (`w_resident`).
//   barriers   afull / wfull / tmem_empty live in the LEADER (cluster rank 0): both producers' TMA bytes are
//              counted there (the leader arms 2x the stage bytes), the follower's epilogue warps arrive remotely;
//              aempty / wempty / tmem_full exist in both CTAs and are signalled by multicast tcgen05.commit.
//   roles      warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (leader only; allocates TMEM in both),
//              warps 4..11 = epilogue of the CTA's own tile (TMEM lanes 0..127 of each CTA = its 128 pixels).
// ---------------------------------------------------------------------------------------------
template <int KIND, int ACT, int KB, int FL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ TcConvParams prm) {
    // KB = 64: x-major halo patch -- ONE TMA box of (8 + kh - 1) x (16 + kw - 1) pixels per 64-channel block, fetched with
    // the tensor dims ordered {C, y, x} so that shared-memory row = x * RY + y, serves every tap: with M row m = x * 8 + y
    // tap (ky,kx) is that patch read from row kx * RY + ky with an atom stride of RY rows.  (tools/umma_stride_probe.cu:
    // a SWIZZLE_128B K-major descriptor accepts any 128-byte-aligned start and atom stride with base_offset 0.)  A 3x3
    // conv pulls 1 patch of 10 x 18 pixels per block instead of 3 patches of 10 x 16: 2.6x less activation traffic.
    // KB = 32 (the 7x1 stems, kw = 1: one box already serves all taps): y-major row patch as in conv_tc_patch_kernel.
    constexpr bool XMT = (KB == 64);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int Nh = prm.Npad >> 1;                                   // weight rows staged by this CTA
    const uint32_t a_part = prm.a_part_bytes, a_stage_bytes = (uint32_t)prm.a_parts * prm.a_part_bytes;
    constexpr uint32_t KROW = KB * 2;                               // bytes of one K block row (128 or 64) == swizzle span
    const uint32_t b_bytes = (uint32_t)Nh * KROW, w_stage_bytes = (uint32_t)prm.b_parts * b_bytes;    // K block of KB channels
    uint8_t* a_ring = smem;
    uint8_t* w_ring = a_ring + (size_t)prm.a_stages * a_stage_bytes;
    uint8_t* epi_smem = w_ring + (size_t)prm.w_stages * w_stage_bytes;
    uint64_t* afull = reinterpret_cast<uint64_t*>(epi_smem + TC2_EPI_BYTES);
    uint64_t* aempty = afull + TCP_MAX_A;
    uint64_t* wfull = aempty + TCP_MAX_A;
    uint64_t* wempty = wfull + TCP_MAX_W;
    uint64_t* tmem_full_bar = wempty + TCP_MAX_W;            // [TC_MAX_ACC]
    uint64_t* tmem_empty_bar = tmem_full_bar + TC_MAX_ACC;   // [TC_MAX_ACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + TC_MAX_ACC);

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int ygroups = XMT ? 1 : prm.kh / prm.ygroup;       // A steps per (kb, kx); x-major: one patch per K block, ygroup = taps
    const int kx_n = XMT ? 1 : prm.kw;
    const int pairs = (int)gridDim.x >> 1, pair_id = (int)blockIdx.x >> 1;
    const int items = (prm.num_tiles + 1) >> 1;              // work item = two consecutive tiles

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < prm.nsrc; ++s) { tma_prefetch_desc(&prm.act[s][0]); tma_prefetch_desc(&prm.act[s][1]); }
        tma_prefetch_desc(&prm.wgt[0]);
        tma_prefetch_desc(&prm.wgt[1]);
        for (int s = 0; s < prm.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < prm.w_stages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int s = 0; s < (int)prm.acc_stages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * TC2_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, prm.tmem_cols);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // both CTAs' barriers and TMEM exist before any remote access
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = prm.tiles_x * prm.tiles_y;
    if (warp >= TC2_EPI_WARP0) {
        // ===== epilogue warpgroups: take the registers the role warpgroup gives back =====
        reg_alloc<TC2_EPI_REGS>();
        const TileWalk tw{pair_id, pairs, items, 2, (int)rank, leader ? 0u : mapa_u32(smem_u32(tmem_empty_bar), 0)};
        if constexpr (KIND == DKT_EPI_PROJ)
            conv_tc_epilogue_proj<ACT, XMT>(prm, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane, tiles_per_img, tw);
        else
            conv_tc_epilogue_warps<KIND, ACT, FL, XMT>(prm, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane, tiles_per_img, tw);
    } else {
    reg_dealloc<TC2_ROLE_REGS>();
    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A patches, own half of the W blocks; bytes counted on the leader =====
        // (per-tap state -- weight row, ring slot address, barrier address -- advances by additions: with one MMA per
        // K16 step a tap lasts ~512 clocks and this loop must keep up with it)
        const uint32_t afull_l = mapa_u32(smem_u32(afull), 0), wfull_l = mapa_u32(smem_u32(wfull), 0);
        const int wrow0 = (int)rank * Nh;
        const int a_stages = prm.a_stages, w_stages = prm.w_stages, ntaps = prm.ygroup, Npad = prm.Npad, nsrc = prm.nsrc;
        const bool a2 = prm.a_parts == 2, b2 = prm.b_parts == 2 && prm.merged_n != 1, w_stream = !prm.w_resident;
        const CUtensorMap* const wmap0 = prm.merged_n == 1 ? &prm.wgt[rank] : &prm.wgt[0];   // merged: this CTA's half of B' = [w_hi; w_lo] is one whole plane
        const int wrow_first = prm.merged_n == 1 ? 0 : wrow0;
        const uint32_t a_tx2 = 2u * prm.a_tx_bytes, w_tx2 = 2u * w_stage_bytes;
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0;
        uint8_t* wst = w_ring;
        bool first = true;
        for (int item = pair_id; item < items; item += pairs) {
            int tile = 2 * item + (int)rank;
            if (tile >= prm.num_tiles) tile = prm.num_tiles - 1;      // filler: valid loads, results dropped
            const int b = tile / tiles_per_img;
            const int r = tile - b * tiles_per_img;
            const int y0 = (r / prm.tiles_x) * TC_TILE_H, x0 = (r % prm.tiles_x) * TC_TILE_W;
            int kofs = 0;
            for (int s = 0; s < nsrc; ++s) {
                const int kblocks = prm.kblocks[s];
                int c = prm.c_begin[s], kc = kofs;
                for (int kb = 0; kb < kblocks; ++kb, c += KB, kc += KB) {
                    for (int kx = 0; kx < kx_n; ++kx) {
                        const int xs = XMT ? x0 - prm.pad_x : x0 + kx - prm.pad_x;
                        for (int yg = 0; yg < ygroups; ++yg) {
                            const int ys = y0 + (XMT ? 0 : yg * prm.ygroup) - prm.pad_y;
                            mbar_wait(&aempty[as], aph ^ 1u);
                            if (elect_one()) {
                                uint8_t* ast = a_ring + (size_t)as * a_stage_bytes;
                                if (leader) mbar_arrive_expect_tx(&afull[as], a_tx2);
                                if (XMT) {           // tensor dims ordered {C, y, x, b}
                                    tma_load_4d_pair(ast, &prm.act[s][0], afull_l + as * 8u, c, ys, xs, b);
                                    if (a2) tma_load_4d_pair(ast + a_part, &prm.act[s][1], afull_l + as * 8u, c, ys, xs, b);
                                } else {
                                    tma_load_4d_pair(ast, &prm.act[s][0], afull_l + as * 8u, c, xs, ys, b);
                                    if (a2) tma_load_4d_pair(ast + a_part, &prm.act[s][1], afull_l + as * 8u, c, xs, ys, b);
                                }
                            }
                            if (++as == a_stages) { as = 0; aph ^= 1u; }
                            if (w_stream || first) {
                                int wrow = (XMT ? 0 : (yg * prm.ygroup) * prm.kw + kx) * Npad + wrow_first;   // tap * Npad (+ this CTA's rows)
                                const int wrow_step = (XMT ? 1 : prm.kw) * Npad;
                                for (int kyi = 0; kyi < ntaps; ++kyi, wrow += wrow_step) {
                                    mbar_wait(&wempty[ws], wph ^ 1u);
                                    if (elect_one()) {
                                        if (leader) mbar_arrive_expect_tx(&wfull[ws], w_tx2);
                                        tma_load_2d_pair(wst, wmap0, wfull_l + ws * 8u, kc, wrow);
                                        if (b2) tma_load_2d_pair(wst + b_bytes, &prm.wgt[1], wfull_l + ws * 8u, kc, wrow);
                                    }
                                    wst += w_stage_bytes;
                                    if (++ws == w_stages) { ws = 0; wph ^= 1u; wst = w_ring; }
                                }
                            }
                        }
                    }
                }
                kofs += prm.c_count[s];
            }
            first = false;
        }
    } else if (warp == 1) {
        if (leader) {
            // ===== MMA issuer (leader): M = 256 over both CTAs' tiles =====
            // one instantiation of the issue loop per operand mode (see pair_mma_loop): what a tap costs THIS warp in
            // issue slots bounds every conv whose MMAs are short (N <= 128)
            const bool a_lo = prm.a_parts == 2, b_lo = prm.b_parts == 2;
            const PairMmaCtx cx{a_ring, w_ring, afull, aempty, wfull, wempty, tmem_full_bar, tmem_empty_bar, tmem_base, pair_id, pairs, items};
            if (prm.merged_n == 2) pair_mma_loop<KB, XMT, 2, true, true>(prm, cx);
            else if (prm.merged_n == 1) pair_mma_loop<KB, XMT, 1, false, true>(prm, cx);
            else if (a_lo && b_lo) pair_mma_loop<KB, XMT, 0, true, true>(prm, cx);
            else if (b_lo) pair_mma_loop<KB, XMT, 0, false, true>(prm, cx);
            else if (a_lo) pair_mma_loop<KB, XMT, 0, true, false>(prm, cx);
            else pair_mma_loop<KB, XMT, 0, false, false>(prm, cx);
        }
    }
    }

    // neither CTA may leave (or free TMEM) while its peer can still reach its shared memory / barriers
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, prm.tmem_cols);
}

}  // namespace dkt

using namespace dkt;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- launch: one instantiation per (kernel family, epilogue kind, activation[, K block, operand flags]) ----
enum ConvFamily { FAM_PATCH32, FAM_PATCH64, FAM_TAP64, FAM_PAIR, FAM_PAIR_K32 };

// which optional operands / outputs a LINEAR epilogue uses, as EPF_* bits (EPF_GENERIC if not expressible)
static int epilogue_flags(const dkt_epilogue& e) {
    if (e.kind == DKT_EPI_GRU_ZR || e.kind == DKT_EPI_GRU_Q) {       // only the output precisions vary
        return (e.out.f32 ? EPF_OUT_F32 : 0) | (e.out.hi ? (e.out.lo ? EPF_OUT_SPLIT : EPF_OUT_HI) : 0);
    }
    if (e.kind != DKT_EPI_LINEAR) return EPF_GENERIC;
    int fl = 0;
    if (e.out.f32) fl |= EPF_OUT_F32;
    if (e.out.hi) fl |= e.out.lo ? EPF_OUT_SPLIT : EPF_OUT_HI;
    if (e.res) fl |= EPF_RES_F32;
    else if (e.res_hi) fl |= EPF_RES_SPLIT;
    if (e.ctx) fl |= EPF_CTX;
    if (e.stats_partial) fl |= EPF_STATS;
    if (e.tail) fl |= EPF_TAIL;
    return fl;
}

template <int KIND, int ACT, int KB, int FL>
static int launch_pair(const TcConvParams& prm, unsigned grid, size_t smem_bytes, cudaStream_t st) {
    DKT_ENSURE_SMEM(227 * 1024, conv_tc_pair_kernel<KIND, ACT, KB, FL>);
    conv_tc_pair_kernel<KIND, ACT, KB, FL><<<grid, TC2_THREADS, smem_bytes, st>>>(prm);   // grid even (cluster of 2)
    DKT_RETURN_LAST();
}

// pair kernel: the hot LINEAR operand combinations of the two models get their own instantiation (see
// conv_tc_epilogue_warps); everything else runs the run-time-generic epilogue
template <int KIND, int ACT, int KB>
static int launch_pair_fl(const TcConvParams& prm, unsigned grid, size_t smem_bytes, cudaStream_t st) {
    if constexpr (KIND == DKT_EPI_LINEAR && ACT == DKT_ACT_RELU) {
        const int fl = epilogue_flags(prm.epi);
        if (fl == EPF_OUT_SPLIT) return launch_pair<KIND, ACT, KB, EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
        if (fl == EPF_OUT_HI) return launch_pair<KIND, ACT, KB, EPF_OUT_HI>(prm, grid, smem_bytes, st);
        if (fl == EPF_OUT_F32) return launch_pair<KIND, ACT, KB, EPF_OUT_F32>(prm, grid, smem_bytes, st);      // MobileNetV2 expand convs
        if constexpr (KB == 64) {
            if (fl == (EPF_OUT_HI | EPF_TAIL)) return launch_pair<KIND, ACT, KB, EPF_OUT_HI | EPF_TAIL>(prm, grid, smem_bytes, st);
            if (fl == (EPF_OUT_SPLIT | EPF_RES_SPLIT)) return launch_pair<KIND, ACT, KB, EPF_OUT_SPLIT | EPF_RES_SPLIT>(prm, grid, smem_bytes, st);
            if (fl == (EPF_OUT_SPLIT | EPF_RES_F32)) return launch_pair<KIND, ACT, KB, EPF_OUT_SPLIT | EPF_RES_F32>(prm, grid, smem_bytes, st);
            if (fl == (EPF_OUT_SPLIT | EPF_TAIL)) return launch_pair<KIND, ACT, KB, EPF_OUT_SPLIT | EPF_TAIL>(prm, grid, smem_bytes, st);
        }
    }
    if constexpr (KIND == DKT_EPI_LINEAR && ACT == DKT_ACT_NONE) {
        const int fl = epilogue_flags(prm.epi);
        if (fl == (EPF_OUT_F32 | EPF_STATS)) return launch_pair<KIND, ACT, KB, EPF_OUT_F32 | EPF_STATS>(prm, grid, smem_bytes, st);
        if constexpr (KB == 64) {
            if (fl == EPF_OUT_F32) return launch_pair<KIND, ACT, KB, EPF_OUT_F32>(prm, grid, smem_bytes, st);
            if (fl == EPF_OUT_SPLIT) return launch_pair<KIND, ACT, KB, EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);   // fnet's output conv
            // MobileNetV2 project convs: fp32 + 16-bit pair out, with / without the linear-bottleneck residual as ctx
            if (fl == (EPF_OUT_F32 | EPF_OUT_SPLIT)) return launch_pair<KIND, ACT, KB, EPF_OUT_F32 | EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
            if (fl == (EPF_OUT_F32 | EPF_OUT_SPLIT | EPF_CTX))
                return launch_pair<KIND, ACT, KB, EPF_OUT_F32 | EPF_OUT_SPLIT | EPF_CTX>(prm, grid, smem_bytes, st);
        }
    }
    if constexpr (KIND == DKT_EPI_LINEAR && ACT == DKT_ACT_TANH && KB == 64) {      // initial hidden states
        const int fl = epilogue_flags(prm.epi);
        if (fl == (EPF_OUT_F32 | EPF_OUT_SPLIT)) return launch_pair<KIND, ACT, KB, EPF_OUT_F32 | EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
        if (fl == (EPF_OUT_F32 | EPF_OUT_HI)) return launch_pair<KIND, ACT, KB, EPF_OUT_F32 | EPF_OUT_HI>(prm, grid, smem_bytes, st);
    }
    if constexpr (KIND == DKT_EPI_GRU_ZR && KB == 64) {            // tensor-core engine: r*h as a 16-bit pair, or hi only
        if (epilogue_flags(prm.epi) == EPF_OUT_SPLIT) return launch_pair<KIND, ACT, KB, EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
        if (epilogue_flags(prm.epi) == EPF_OUT_HI) return launch_pair<KIND, ACT, KB, EPF_OUT_HI>(prm, grid, smem_bytes, st);
    }
    if constexpr (KIND == DKT_EPI_GRU_Q && KB == 64) {             // h' as fp32 (for the next blend) + 16-bit (hi, lo) / hi
        if (epilogue_flags(prm.epi) == (EPF_OUT_F32 | EPF_OUT_SPLIT))
            return launch_pair<KIND, ACT, KB, EPF_OUT_F32 | EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
        if (epilogue_flags(prm.epi) == (EPF_OUT_F32 | EPF_OUT_HI))
            return launch_pair<KIND, ACT, KB, EPF_OUT_F32 | EPF_OUT_HI>(prm, grid, smem_bytes, st);
    }
    return launch_pair<KIND, ACT, KB, EPF_GENERIC>(prm, grid, smem_bytes, st);
}

template <int KIND, int ACT, int FL>
static int launch_tap(const TcConvParams& prm, unsigned grid, size_t smem_bytes, cudaStream_t st) {
    DKT_ENSURE_SMEM(227 * 1024, conv_tc_kernel<64, KIND, ACT, FL>);
    conv_tc_kernel<64, KIND, ACT, FL><<<grid, TC2_THREADS, smem_bytes, st>>>(prm);
    DKT_RETURN_LAST();
}

// per-tap kernel (the encoders' stride-2 convs): same hot operand combinations as the pair kernel
template <int KIND, int ACT>
static int launch_tap_fl(const TcConvParams& prm, unsigned grid, size_t smem_bytes, cudaStream_t st) {
    if constexpr (KIND == DKT_EPI_LINEAR && ACT == DKT_ACT_RELU) {
        if (epilogue_flags(prm.epi) == EPF_OUT_SPLIT) return launch_tap<KIND, ACT, EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
    }
    if constexpr (KIND == DKT_EPI_LINEAR && ACT == DKT_ACT_NONE) {
        const int fl = epilogue_flags(prm.epi);
        if (fl == (EPF_OUT_F32 | EPF_STATS)) return launch_tap<KIND, ACT, EPF_OUT_F32 | EPF_STATS>(prm, grid, smem_bytes, st);
        if (fl == EPF_OUT_F32) return launch_tap<KIND, ACT, EPF_OUT_F32>(prm, grid, smem_bytes, st);
    }
    return launch_tap<KIND, ACT, EPF_GENERIC>(prm, grid, smem_bytes, st);
}

template <int KIND, int ACT, int FL>
static int launch_patch64(const TcConvParams& prm, unsigned grid, size_t smem_bytes, cudaStream_t st) {
    DKT_ENSURE_SMEM(227 * 1024, conv_tc_patch_kernel<64, KIND, ACT, FL>);
    conv_tc_patch_kernel<64, KIND, ACT, FL><<<grid, TC2_THREADS, smem_bytes, st>>>(prm);
    DKT_RETURN_LAST();
}

template <int KIND, int ACT>
static int launch_conv_ka(ConvFamily fam, const TcConvParams& prm, unsigned grid, size_t smem_bytes, cudaStream_t st) {
    if (fam == FAM_PAIR) return launch_pair_fl<KIND, ACT, 64>(prm, grid, smem_bytes, st);
    if (fam == FAM_TAP64) return launch_tap_fl<KIND, ACT>(prm, grid, smem_bytes, st);
    if (fam == FAM_PAIR_K32) {                       // 32-channel K blocks: instantiated for the stems' epilogues only
        if constexpr (KIND == DKT_EPI_LINEAR && (ACT == DKT_ACT_NONE || ACT == DKT_ACT_RELU))
            return launch_pair_fl<KIND, ACT, 32>(prm, grid, smem_bytes, st);
        else
            return DKT_E_UNSUPPORTED;
    }
    // row-patch kernel with 64-channel W steps (N <= 128: the encoders' strided convs, whose small MMA phase leaves the
    // epilogue exposed): the same hot operand combinations as the per-tap kernel get their own instantiation
    if (fam == FAM_PATCH64) {
        if constexpr (KIND == DKT_EPI_LINEAR && ACT == DKT_ACT_RELU) {
            if (epilogue_flags(prm.epi) == EPF_OUT_SPLIT) return launch_patch64<KIND, ACT, EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
        }
        if constexpr (KIND == DKT_EPI_LINEAR && ACT == DKT_ACT_NONE) {
            const int fl = epilogue_flags(prm.epi);
            if (fl == (EPF_OUT_F32 | EPF_STATS)) return launch_patch64<KIND, ACT, EPF_OUT_F32 | EPF_STATS>(prm, grid, smem_bytes, st);
            if (fl == EPF_OUT_SPLIT) return launch_patch64<KIND, ACT, EPF_OUT_SPLIT>(prm, grid, smem_bytes, st);
            if (fl == EPF_OUT_F32) return launch_patch64<KIND, ACT, EPF_OUT_F32>(prm, grid, smem_bytes, st);
        }
    }
    DKT_ENSURE_SMEM(227 * 1024, conv_tc_patch_kernel<32, KIND, ACT>);
    DKT_ENSURE_SMEM(227 * 1024, conv_tc_patch_kernel<64, KIND, ACT>);
    switch (fam) {
        case FAM_PATCH32: conv_tc_patch_kernel<32, KIND, ACT><<<grid, TC2_THREADS, smem_bytes, st>>>(prm); break;
        default:          conv_tc_patch_kernel<64, KIND, ACT><<<grid, TC2_THREADS, smem_bytes, st>>>(prm); break;
    }
    DKT_RETURN_LAST();
}

static int launch_conv(ConvFamily fam, const TcConvParams& prm, unsigned grid, size_t smem_bytes, cudaStream_t st) {
    const int kind = prm.epi.kind, act = prm.epi.act;
    if (kind == DKT_EPI_GRU_ZR) return launch_conv_ka<DKT_EPI_GRU_ZR, DKT_ACT_NONE>(fam, prm, grid, smem_bytes, st);
    if (kind == DKT_EPI_GRU_Q) return launch_conv_ka<DKT_EPI_GRU_Q, DKT_ACT_NONE>(fam, prm, grid, smem_bytes, st);
    if (kind == DKT_EPI_PROJ) {
        if (act == DKT_ACT_RELU) return launch_conv_ka<DKT_EPI_PROJ, DKT_ACT_RELU>(fam, prm, grid, smem_bytes, st);
        if (act == DKT_ACT_NONE) return launch_conv_ka<DKT_EPI_PROJ, DKT_ACT_NONE>(fam, prm, grid, smem_bytes, st);
        return DKT_E_UNSUPPORTED;
    }
    switch (act) {
        case DKT_ACT_NONE:    return launch_conv_ka<DKT_EPI_LINEAR, DKT_ACT_NONE>(fam, prm, grid, smem_bytes, st);
        case DKT_ACT_RELU:    return launch_conv_ka<DKT_EPI_LINEAR, DKT_ACT_RELU>(fam, prm, grid, smem_bytes, st);
        case DKT_ACT_SIGMOID: return launch_conv_ka<DKT_EPI_LINEAR, DKT_ACT_SIGMOID>(fam, prm, grid, smem_bytes, st);
        case DKT_ACT_TANH:    return launch_conv_ka<DKT_EPI_LINEAR, DKT_ACT_TANH>(fam, prm, grid, smem_bytes, st);
        case DKT_ACT_LEAKY:   return launch_conv_ka<DKT_EPI_LINEAR, DKT_ACT_LEAKY>(fam, prm, grid, smem_bytes, st);
        default:              return DKT_E_INVALID;
    }
}

extern "C" int dkt_conv2d_tc_ex(const dkt_tensor* srcs, int nsrc, const uint16_t* w_hi, const uint16_t* w_lo,
                                int kh, int kw, int stride, int N, const dkt_epilogue* epi,
                                int B, int Hin, int Win, int H, int W, void* stream) {
    DKT_CHECK_ARG(srcs && w_hi && epi);
    DKT_CHECK_ARG(nsrc >= 1 && nsrc <= DKT_MAX_SRCS);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && Hin > 0 && Win > 0 && N > 0);
    if (kh < 1 || kw < 1 || kh > 7 || kw > 7 || !(kh & 1) || !(kw & 1)) return DKT_E_UNSUPPORTED;
    if (stride != 1 && stride != 2) return DKT_E_UNSUPPORTED;
    if (N > 256) return DKT_E_UNSUPPORTED;
    const int pad_y = kh / 2, pad_x = kw / 2;
    // output extent of a padded strided conv (PyTorch: floor((in + 2p - k) / s) + 1)
    if (H != (Hin + 2 * pad_y - kh) / stride + 1 || W != (Win + 2 * pad_x - kw) / stride + 1) return DKT_E_INVALID;
    // epilogue sanity (vector accesses need 4-channel granularity)
    const dkt_epilogue& e = *epi;
    if (e.kind == DKT_EPI_LINEAR) {
        DKT_CHECK_ARG(e.out.f32 || e.out.hi);
        if (e.stats_partial && ((N % 32) || e.tail)) return DKT_E_UNSUPPORTED;
    } else if (e.kind == DKT_EPI_PROJ) {
        DKT_CHECK_ARG(e.proj && e.out.f32 && e.out.c_count >= DKT_PROJ_LD);
        if ((e.out.C % 4) || (e.out.c_begin % 4) || !aligned16(e.proj)) return DKT_E_ALIGNMENT;
        if (e.ctx || e.res || e.tail) return DKT_E_UNSUPPORTED;
    } else if (e.kind != DKT_EPI_GRU_ZR && e.kind != DKT_EPI_GRU_Q) {
        return DKT_E_INVALID;
    } else {
        DKT_CHECK_ARG(e.ctx && e.z.f32 && e.h.f32 && (e.out.f32 || e.out.hi));
        if (N % 8) return DKT_E_ALIGNMENT;
        if ((e.z.C % 4) || (e.z.c_begin % 4) || (e.h.C % 4) || (e.h.c_begin % 4)) return DKT_E_ALIGNMENT;
    }
    if (N >= 4 && ((e.out.C % 4) || (e.out.c_begin % 4))) return DKT_E_ALIGNMENT;
    if (e.ctx && ((e.ctx_C % 4) || (e.ctx_c0 % 4) || !aligned16(e.ctx))) return DKT_E_ALIGNMENT;
    if (e.res && ((e.res_C % 4) || (e.res_c0 % 4) || !aligned16(e.res))) return DKT_E_ALIGNMENT;
    if (!e.res && e.res_hi) {
        DKT_CHECK_ARG(e.res_lo != nullptr);
        if ((e.res_C % 4) || (e.res_c0 % 4) || (reinterpret_cast<uintptr_t>(e.res_hi) & 7) || (reinterpret_cast<uintptr_t>(e.res_lo) & 7))
            return DKT_E_ALIGNMENT;
    }
    if (e.bias && !aligned16(e.bias)) return DKT_E_ALIGNMENT;
    if (!aligned16(e.out.f32) || !aligned16(e.out.hi) || !aligned16(e.out.lo)) return DKT_E_ALIGNMENT;

    // kernel choice: the row-patch kernel (default) or the per-tap kernel (strided convs, or DKT_CONV_PATCH=0)
    static const int s_patch = [] { const char* v = getenv("DKT_CONV_PATCH"); return (v && v[0] == '0') ? 0 : 1; }();
    static const int s_pair = [] { const char* v = getenv("DKT_CONV_PAIR"); return (v && v[0] == '0') ? 0 : 1; }();
    const int s_sms = device_sms();
    const int Npad = (N + 15) / 16 * 16;
    // 2-MMA mode: the sources carry no lo plane (all of them, or none)
    int n_lo = 0;
    for (int s = 0; s < nsrc; ++s) n_lo += srcs[s].lo != nullptr;
    if (n_lo != 0 && n_lo != nsrc) return DKT_E_INVALID;
    const uint32_t AP = n_lo ? 2u : 1u;
    const uint32_t BP = w_lo ? 2u : 1u;             // w_lo == NULL: single-plane weights (x * w_lo MMA dropped)
    const uint32_t budget = 227u * 1024u - 1024u /*align*/ - TC2_EPI_BYTES - TC_BAR_BYTES;

    // ring geometry of the row-patch kernel
    const int ygroup = (stride == 1) ? kh : 1;
    const int rows_loaded = (stride == 1) ? TC_TILE_H + ygroup - 1 : TC_TILE_H;
    int kb_sum = 0;        // kb_sum: 64-channel K blocks (a source's last one may be partial: c_count % 16 == 0)
    for (int s = 0; s < nsrc; ++s) { kb_sum += (srcs[s].c_count + 63) / 64; }
    const int64_t tiles64 = (int64_t)ceil_div(W, TC_TILE_W) * ceil_div(H, TC_TILE_H) * B;
    // K blocks of 32 channels (64-byte rows): a single 32-channel source, i.e. the x-im2col rows of a 7x7 stem
    // (any other 32-channel source is the first half of a 64-channel block: klast = 2)
    const bool k32 = nsrc == 1 && srcs[0].c_count == 32 && (srcs[0].c_begin % 32) == 0 && kw == 1 && stride == 1;
    const int KBLK = k32 ? 32 : 64;
    const uint32_t a_part_bytes = (uint32_t)rows_loaded * TC_TILE_W * (uint32_t)KBLK * 2u;

    // CTA-pair kernel (default for stride 1): each CTA stages half of a KBLK-channel weight block
    bool use_pair = s_pair != 0 && s_patch != 0 && stride == 1 && tiles64 >= 2;
    // 64-channel K blocks of the pair kernel: ONE x-major halo patch (RY x RX pixels, shared-memory row = x * RY + y) per
    // K block serves every tap (tools/umma_stride_probe.cu: the SW128 descriptor takes any 128-byte-aligned start and any
    // atom stride); each precision's slot is rounded up to the 1024-byte swizzle period
    const bool xm = !k32;
    const int patch_ry = TC_TILE_H + kh - 1, patch_rx = TC_TILE_W + kw - 1;
    const uint32_t xm_box_bytes = (uint32_t)patch_ry * (uint32_t)patch_rx * 128u;
    if (xm && (patch_ry > 256 || patch_rx > 256)) use_pair = false;
    const uint32_t pair_a_part_bytes = xm ? (xm_box_bytes + 1023u) & ~1023u : a_part_bytes;
    int p_a_stages = 2, p_w_stages = 0, p_resident = 0;
    const uint32_t p_w_stage_bytes = BP * (uint32_t)(Npad / 2) * (uint32_t)KBLK * 2u;
    if (use_pair) {
        const int w_steps = (k32 ? 1 : kb_sum) * kh * kw;          // weight blocks per tile
        if (AP * pair_a_part_bytes * 2 >= budget) use_pair = false;
        else if (w_steps <= TCP_MAX_W && AP * pair_a_part_bytes * 2 + (uint32_t)w_steps * p_w_stage_bytes <= budget) {
            p_resident = 1;                                        // whole filter half stays in the ring
            p_w_stages = w_steps;
            p_a_stages = (int)((budget - (uint32_t)w_steps * p_w_stage_bytes) / (AP * pair_a_part_bytes));
            if (p_a_stages > TCP_MAX_A) p_a_stages = TCP_MAX_A;
        } else {
            p_w_stages = (int)((budget - AP * pair_a_part_bytes * 2) / p_w_stage_bytes);
            if (p_w_stages > 8) p_w_stages = 8;
            if (p_w_stages < 2) use_pair = false;
            else if (p_w_stages >= 7 && AP * pair_a_part_bytes * 3 + 4u * p_w_stage_bytes <= budget) {
                p_a_stages = 3;                                    // small N: a third patch stage is worth more
                p_w_stages = (int)((budget - AP * pair_a_part_bytes * 3) / p_w_stage_bytes);
                if (p_w_stages > 8) p_w_stages = 8;
            }
        }
    }

    const int WK = (Npad > 128) ? 32 : 64;
    const uint32_t w_stage_bytes = BP * (uint32_t)Npad * (uint32_t)WK * 2u;
    int a_stages = 2, w_stages = 0;
    bool use_patch = s_patch != 0;
    if (use_patch && !use_pair) {
        if (AP * a_part_bytes * a_stages >= budget) use_patch = false;
        else {
            w_stages = (int)((budget - AP * a_part_bytes * a_stages) / w_stage_bytes);
            if (w_stages > 8) w_stages = 8;
            if (w_stages < 2) use_patch = false;
            else if (w_stages >= 5 && AP * a_part_bytes * 3 + 4u * w_stage_bytes <= budget) {
                a_stages = 3;
                w_stages = (int)((budget - AP * a_part_bytes * 3) / w_stage_bytes);
                if (w_stages > 8) w_stages = 8;
            }
        }
    }
    if (k32 && !use_pair) return DKT_E_UNSUPPORTED;           // 32-channel K blocks exist in the pair kernel only
    const int BK = KBLK;
    const int WBK = use_pair ? KBLK : (use_patch ? WK : BK);  // K extent of a weight box
    // merged weight planes (TcConvParams::merged_n): pair kernel, hi-plane activations, (hi, lo) weights, N a power of two
    // <= 128 so that the lo accumulator starts exactly N columns after the main one
    static const int s_merge = [] { const char* v = getenv("DKT_CONV_MERGE_N"); return (v && v[0] == '0') ? 0 : 1; }();
    static const int s_split_m = [] { const char* v = getenv("DKT_ACC_SPLIT"); return (v && v[0] == '0') ? 0 : 1; }();
    static const int s_merge3 = [] { const char* v = getenv("DKT_CONV_MERGE3"); return (v && v[0] == '0') ? 0 : 1; }();
    const bool merged = s_merge && s_split_m && use_pair && AP == 1 && BP == 2 && (Npad == 64 || Npad == 128) &&
                        epi->kind != DKT_EPI_PROJ;
    const bool merged3 = s_merge >= 1 && s_split_m && use_pair && AP == 2 && BP == 2 && Npad == 64 && epi->kind == DKT_EPI_LINEAR &&
                         s_merge3;
    const int wbox_rows = use_pair ? (merged ? Npad : Npad / 2) : Npad;         // N extent of a weight box

    TcConvParams prm{};
    prm.nsrc = nsrc;
    prm.a_parts = (int)AP;
    prm.b_parts = (int)BP;
    prm.merged_n = merged ? 1 : (merged3 ? 2 : 0);
    int cin_total = 0;
    for (int s = 0; s < nsrc; ++s) {
        const dkt_tensor& t = srcs[s];
        DKT_CHECK_ARG(t.hi && t.c_count > 0 && t.c_begin >= 0 && t.c_begin + t.c_count <= t.C);
        if ((t.c_begin % BK) || (t.c_count % 16) || (t.C % 8) || !aligned16(t.hi) || !aligned16(t.lo))
            return DKT_E_ALIGNMENT;
        prm.c_begin[s] = t.c_begin;
        prm.kblocks[s] = (t.c_count + BK - 1) / BK;
        prm.c_count[s] = t.c_count;
        prm.klast[s] = (t.c_count - (prm.kblocks[s] - 1) * BK) / 16;
        cin_total += t.c_count;
        if (use_pair && xm) {
            // x-major patch: the tensor is described as {C, y, x, b} so that the box lands with y fastest
            const uint64_t dims[4] = {(uint64_t)t.C, (uint64_t)Hin, (uint64_t)Win, (uint64_t)B};
            const uint64_t strides[4] = {1, (uint64_t)t.C * Win, (uint64_t)t.C, (uint64_t)t.C * Win * Hin};
            const uint32_t box[4] = {(uint32_t)BK, (uint32_t)patch_ry, (uint32_t)patch_rx, 1};
            const uint32_t estr[4] = {1, 1, 1, 1};
            if (!make_tmap_bf16(&prm.act[s][0], t.hi, 4, dims, strides, box, BK * 2, estr)) return DKT_E_DRIVER;
            if (!make_tmap_bf16(&prm.act[s][1], t.lo ? t.lo : t.hi, 4, dims, strides, box, BK * 2, estr)) return DKT_E_DRIVER;
            continue;
        }
        const uint64_t dims[4] = {(uint64_t)t.C, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
        const uint64_t strides[4] = {1, (uint64_t)t.C, (uint64_t)t.C * Win, (uint64_t)t.C * Win * Hin};
        const uint32_t box_rows = use_patch ? (uint32_t)(stride == 1 ? rows_loaded : TC_TILE_H * stride) : (uint32_t)(TC_TILE_H * stride);
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)(TC_TILE_W * stride), box_rows, 1};
        const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        if (!make_tmap_bf16(&prm.act[s][0], t.hi, 4, dims, strides, box, BK * 2, estr)) return DKT_E_DRIVER;
        if (!make_tmap_bf16(&prm.act[s][1], t.lo ? t.lo : t.hi, 4, dims, strides, box, BK * 2, estr)) return DKT_E_DRIVER;
    }
    prm.kh = kh; prm.kw = kw; prm.pad_y = pad_y; prm.pad_x = pad_x; prm.stride = stride;
    prm.taps = kh * kw;
    prm.N = N;
    prm.Npad = Npad;
    {
        const uint64_t dims[2] = {(uint64_t)cin_total, (uint64_t)prm.taps * prm.Npad};
        const uint64_t strides[2] = {1, (uint64_t)cin_total};
        const uint32_t box[2] = {(uint32_t)WBK, (uint32_t)wbox_rows};
        if (!aligned16(w_hi) || !aligned16(w_lo)) return DKT_E_ALIGNMENT;
        if (!make_tmap_bf16(&prm.wgt[0], w_hi, 2, dims, strides, box, WBK * 2)) return DKT_E_DRIVER;
        if (!make_tmap_bf16(&prm.wgt[1], w_lo ? w_lo : w_hi, 2, dims, strides, box, WBK * 2)) return DKT_E_DRIVER;
    }
    prm.H = H;
    prm.W = W;
    prm.tiles_x = ceil_div(W, TC_TILE_W);
    prm.tiles_y = ceil_div(H, TC_TILE_H);
    const int64_t tiles = (int64_t)prm.tiles_x * prm.tiles_y * B;
    if (tiles > 0x7fffffff) return DKT_E_UNSUPPORTED;
    prm.num_tiles = (int)tiles;
    uint32_t cols = 32;
    while (cols < (uint32_t)prm.Npad) cols <<= 1;
    {
        // second accumulator for the lo products (see TcConvParams::acc_lo_off) whenever two stages of two accumulators
        // fit the 512 TMEM columns, i.e. N <= 128; LINEAR / GRU epilogues only (PROJ keeps its own tile walk).
        // DKT_ACC_SPLIT=0 restores the single accumulator (A/B knob).
        static const int s_split = [] { const char* v = getenv("DKT_ACC_SPLIT"); return (v && v[0] == '0') ? 0 : 1; }();
        const bool lo_mmas = AP == 2 || BP == 2;
        if (prm.merged_n == 2) {
            cols = 256;                         // [main | lo_w] interleaved in 128 columns + x_lo * w_hi in 64 more
        } else if (s_split && lo_mmas && cols <= 128 && e.kind != DKT_EPI_PROJ) {
            prm.acc_lo_off = cols;
            cols *= 2;
        }
    }
    prm.acc_cols = cols;
    prm.acc_stages = 512u / cols > (uint32_t)TC_MAX_ACC ? (uint32_t)TC_MAX_ACC : 512u / cols;
    {
        static const int s_acc_env = [] { const char* v = getenv("DKT_CONV_ACC_STAGES"); return v ? atoi(v) : 0; }();
        if (s_acc_env >= 2 && (uint32_t)s_acc_env < prm.acc_stages) prm.acc_stages = (uint32_t)s_acc_env;   // A/B knob (2 = old)
    }
    prm.tmem_cols = prm.acc_stages * cols;            // a power of two >= 32 by construction
    {
        static const int s_sleep = [] { const char* v = getenv("DKT_EPI_SLEEP_NS"); return v ? atoi(v) : 0; }();
        prm.epi_sleep_ns = (uint32_t)s_sleep;
    }
    prm.epi = e;

    if (use_pair) {
        prm.ygroup = xm ? kh * kw : ygroup;
        prm.patch_rows = patch_ry;
        prm.a_stages = p_a_stages;
        prm.w_stages = p_w_stages;
        prm.w_resident = p_resident;
        prm.a_part_bytes = pair_a_part_bytes;
        prm.a_tx_bytes = AP * (xm ? xm_box_bytes : a_part_bytes);
        const uint32_t a_part_bytes = pair_a_part_bytes;
        const int items = (int)((tiles + 1) / 2);
        // as few CTA pairs as finish in the same number of rounds (255 items: 64 pairs x 4 rounds, not 74 x 3.45): the
        // SMs left free run the other stream's kernels (the coarse GRUs and the motion encoder overlap, update.py)
        static const int s_max_pairs = [] { const char* v = getenv("DKT_CONV_MAX_PAIRS"); return v ? atoi(v) : 0; }();   // A/B knob
        const int max_pairs = (s_max_pairs > 0 && s_max_pairs < s_sms / 2) ? s_max_pairs : s_sms / 2;
        const int rounds = (items + max_pairs - 1) / max_pairs;
        const int pairs = (items + rounds - 1) / rounds;
        const size_t smem_bytes = (size_t)p_a_stages * AP * a_part_bytes + (size_t)p_w_stages * p_w_stage_bytes + TC2_EPI_BYTES + 1024 + TC_BAR_BYTES;
        return launch_conv(k32 ? FAM_PAIR_K32 : FAM_PAIR, prm, 2u * (unsigned)pairs, smem_bytes, (cudaStream_t)stream);
    }
    const unsigned grid = (unsigned)(tiles < s_sms ? tiles : s_sms);
    if (use_patch) {
        prm.ygroup = ygroup;
        prm.a_stages = a_stages;
        prm.w_stages = w_stages;
        prm.a_part_bytes = a_part_bytes;
        const size_t smem_bytes = (size_t)a_stages * AP * a_part_bytes + (size_t)w_stages * w_stage_bytes + TC2_EPI_BYTES + 1024 + TC_BAR_BYTES;
        return launch_conv(WK == 32 ? FAM_PATCH32 : FAM_PATCH64, prm, grid, smem_bytes, (cudaStream_t)stream);
    }
    // per-tap kernel: the ring takes what the epilogue buffers and barriers leave of the 227 KB
    const uint32_t stage_bytes = AP * 128u * (uint32_t)BK * 2u + BP * (uint32_t)prm.Npad * (uint32_t)BK * 2u;
    int stages = (int)(budget / stage_bytes);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (stages < 2) return DKT_E_UNSUPPORTED;
    prm.stages = stages;
    const size_t smem_bytes = (size_t)stages * stage_bytes + TC2_EPI_BYTES + 1024 + TC_BAR_BYTES;
    return launch_conv(FAM_TAP64, prm, grid, smem_bytes, (cudaStream_t)stream);
}

extern "C" int dkt_conv2d_tc(const dkt_tensor* srcs, int nsrc, const uint16_t* w_hi, const uint16_t* w_lo,
                             int ksize, int N, const dkt_epilogue* epi, int B, int H, int W, void* stream) {
    if (ksize != 1 && ksize != 3) return DKT_E_UNSUPPORTED;
    return dkt_conv2d_tc_ex(srcs, nsrc, w_hi, w_lo, ksize, ksize, 1, N, epi, B, H, W, H, W, stream);
}
