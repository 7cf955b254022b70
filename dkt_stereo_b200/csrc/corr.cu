// K1 (CUDA-core fp32 variant) and K2: correlation volume build + pyramid, indexed lookup,
// and the IGEV geometry-volume twins.  See include/dkt_stereo_b200.h for the contracts and the
// reference lines each entry point replaces.
#include "common.cuh"

namespace dkt {

struct PyrPtrs {
    float* p[DKT_MAX_LEVELS];
    int    w[DKT_MAX_LEVELS];   // W2 >> l (floor chain)
};

// ---------------------------------------------------------------------------------------------
// K1 / fp32 SIMT:  C[w1][w2] = scale * sum_d f1[d][w1] * f2[d][w2]  per (b, y) row, with the
// W2 pyramid pooled in registers so that levels 1..3 never re-read level 0 from HBM.
// CTA tile 64 (w1) x 128 (w2), 256 threads, 4 x 8 outputs per thread (8 consecutive w2 so that
// three pooling levels stay inside one thread).
// ---------------------------------------------------------------------------------------------
constexpr int CB_BM = 64, CB_BN = 128, CB_BK = 16;

__global__ void __launch_bounds__(256)
corr1d_build_simt_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                         int64_t sb, int64_t sd, int64_t sh, int64_t sw,
                         PyrPtrs pyr, int D, int H, int W1, int W2, int levels, float scale) {
    __shared__ __align__(16) float As[CB_BK][CB_BM];
    __shared__ __align__(16) float Bs[CB_BK][CB_BN];

    const int row = blockIdx.z;              // b*H + y
    const int b = row / H, y = row % H;
    const int m0 = blockIdx.y * CB_BM;
    const int n0 = blockIdx.x * CB_BN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const float* a_base = f1 + b * sb + y * sh;
    const float* b_base = f2 + b * sb + y * sh;
    const bool kfast = (sd == 1);            // channels_last: d is the contiguous axis

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < D; k0 += CB_BK) {
        // A tile: 16 x 64 = 1024 elements, 4 per thread
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;
            int k = kfast ? (e % CB_BK) : (e / CB_BM);
            int m = kfast ? (e / CB_BK) : (e % CB_BM);
            float v = 0.f;
            if (k0 + k < D && m0 + m < W1) v = __ldg(a_base + (int64_t)(k0 + k) * sd + (int64_t)(m0 + m) * sw);
            As[k][m] = v;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = tid + i * 256;
            int k = kfast ? (e % CB_BK) : (e / CB_BN);
            int n = kfast ? (e / CB_BK) : (e % CB_BN);
            float v = 0.f;
            if (k0 + k < D && n0 + n < W2) v = __ldg(b_base + (int64_t)(k0 + k) * sd + (int64_t)(n0 + n) * sw);
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CB_BK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
            float av[4] = {a.x, a.y, a.z, a.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int w2 = n0 + tx * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int w1 = m0 + ty * 4 + i;
        if (w1 >= W1) continue;
        const int64_t prow = (int64_t)row * W1 + w1;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = acc[i][j] * scale;
        float* o0 = pyr.p[0] + prow * pyr.w[0];
        if ((pyr.w[0] & 3) == 0 && w2 + 7 < pyr.w[0]) {
            *reinterpret_cast<float4*>(o0 + w2) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o0 + w2 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (w2 + j < pyr.w[0]) o0[w2 + j] = v[j];
        }
        if (levels > 1) {
            float l1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) l1[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
            float* o1 = pyr.p[1] + prow * pyr.w[1];
#pragma unroll
            for (int j = 0; j < 4; ++j) if (w2 / 2 + j < pyr.w[1]) o1[w2 / 2 + j] = l1[j];
            if (levels > 2) {
                float l2[2] = {(l1[0] + l1[1]) * 0.5f, (l1[2] + l1[3]) * 0.5f};
                float* o2 = pyr.p[2] + prow * pyr.w[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) if (w2 / 4 + j < pyr.w[2]) o2[w2 / 4 + j] = l2[j];
                if (levels > 3) {
                    float* o3 = pyr.p[3] + prow * pyr.w[3];
                    if (w2 / 8 < pyr.w[3]) o3[w2 / 8] = (l2[0] + l2[1]) * 0.5f;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: per pixel, per level: 2r+2 adjacent volume entries -> 2r+1 linearly interpolated taps
// (all taps of a level share one fractional weight because the tap offsets are integers).
//
// A CTA owns LK_PIX = 32 consecutive pixels.  Phase 1 (one thread per (pixel, row-sample))
// gathers the taps into a shared tile [32][C]; phase 2 either streams the tile out with
// fully coalesced 16-byte stores (fp32 + optional bf16 hi/lo, NHWC fast path; any strides on the
// generic path) or -- ENC = true -- applies the motion encoder's 1x1 convolution + bias + ReLU
// (reference core/update.py:79 `convc1`) in exact fp32 straight from the tile, so the 36 / 162
// channel lookup result never travels to HBM.  The coordinate bookkeeping of the loop body
// (`coords1 += delta_flow`, `flow = coords1 - coords0`, IGEV `disp += delta_disp`) is fused in
// front: the CTA reads its 32 coordinates once, updates them in shared memory and writes back.
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void sample_row(const float* __restrict__ row, int W, float x, float* taps) {
    const float xf = floorf(x);
    const float a = x - xf;
    const int i0 = (int)xf - R;
    float v[2 * R + 2];
#pragma unroll
    for (int k = 0; k < 2 * R + 2; ++k) {
        int idx = i0 + k;
        v[k] = (idx >= 0 && idx < W) ? __ldg(row + idx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) taps[k] = (1.f - a) * v[k] + a * v[k + 1];
}

struct ConstPyrPtrs {
    const float* p[DKT_MAX_LEVELS];
    int          w[DKT_MAX_LEVELS];
};

constexpr int LK_PIX = 32;
constexpr int LK_ENC_N = 64;          // output channels of convc1

struct LookupOut {
    float* f32; uint16_t* hi; uint16_t* lo;     // plain output (ENC = false)
    int64_t ob, oc, op;
    int fast;                                    // NHWC, oc == 1, ob == HW * op, op % 4 == 0
    const float* enc_w;                          // ENC: [C][64] fp32 (k-major), bias [64]
    const float* enc_b;
    dkt_tensor enc_out;                          // ENC: 64-channel NHWC slice
    int dbg;                                     // profiling only (DKT_LOOKUP_PHASE): bit 0 skips the gather, bit 1 the encode
};

// phase 2, plain: tile (row stride TS) -> global
__device__ __forceinline__ void lookup_store_tile(const float* tile, int TS, int C, const LookupOut& o,
                                                  int64_t p0, int npix, int HW) {
    if (o.fast) {
        const int C4 = (int)o.op >> 2;           // 16-byte groups per pixel row (incl. zero padding)
        for (int i = threadIdx.x; i < npix * C4; i += blockDim.x) {
            const int px = i / C4, c = (i - px * C4) << 2;
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (c + u < C) ? tile[px * TS + c + u] : 0.f;
            const int64_t off = (p0 + px) * o.op + c;
            *reinterpret_cast<float4*>(o.f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
            if (o.hi) {
                uint32_t h0, l0, h1, l1;
                split16x2(v[0], v[1], h0, l0);
                split16x2(v[2], v[3], h1, l1);
                *reinterpret_cast<uint2*>(o.hi + off) = make_uint2(h0, h1);
                if (o.lo) *reinterpret_cast<uint2*>(o.lo + off) = make_uint2(l0, l1);
            }
        }
    } else {
        // generic strides (e.g. NCHW for the drop-in CorrBlock1D): pixel fastest for coalescing
        for (int i = threadIdx.x; i < npix * C; i += blockDim.x) {
            const int c = i / npix, px = i - c * npix;
            const int64_t p = p0 + px;
            const int64_t b = p / HW;
            const int64_t off = b * o.ob + (p - b * HW) * o.op + (int64_t)c * o.oc;
            const float v = tile[px * TS + c];
            o.f32[off] = v;
            if (o.hi) {
                uint16_t h, l;
                split16(v, h, l);
                o.hi[off] = h;
                if (o.lo) o.lo[off] = l;
            }
        }
    }
}

// phase 2, ENC: out[px][n] = relu(b[n] + sum_k tile[k][px] * w[k][n]), n < 64, for a chunk of LK_ENC_PIX = 64 pixels.
// The tile is K-MAJOR ([k][TSP], pixels contiguous): a thread owns 4 pixels x 4 channels (pg = tid >> 4, cg = tid & 15;
// 256 threads) and per k needs one 16-byte load of its 4 pixels (2 distinct addresses per warp: broadcast) and one
// of its 4 weights -- 2 LDS per 16 FMA, against 3 per 8 with a pixel-major tile and 2 x 4 outputs per thread.
constexpr int LK_ENC_PIX = 64;
constexpr int LK_ENC_TSP = LK_ENC_PIX + 4;      // row stride of the k-major tile (16-byte aligned rows)

__device__ __forceinline__ void lookup_encode_tile(const float* tile, int C, const float* ws,
                                                   const LookupOut& o, int64_t p0, int npix) {
    const int cg = threadIdx.x & 15, pg = threadIdx.x >> 4;
    // packed fp32 FMAs (sm_100 FFMA2: two IEEE fmas per issue slot, bit-identical to fmaf): the contraction is bound by
    // FMA issue, 36 x 64 (RAFT) / 162 x 64 (IGEV) per pixel
    float2 acc[4][2];
    const float4 bias = *reinterpret_cast<const float4*>(o.enc_b + cg * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = make_float2(bias.x, bias.y); acc[i][1] = make_float2(bias.z, bias.w); }
    const float* tcol = tile + pg * 4;
    const float* wcol = ws + cg * 4;
#pragma unroll 2
    for (int k = 0; k < C; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(tcol + k * LK_ENC_TSP);
        const float4 w = *reinterpret_cast<const float4*>(wcol + k * LK_ENC_N);
        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 a2 = make_float2(av[i], av[i]);
            acc[i][0] = __ffma2_rn(a2, w01, acc[i][0]);
            acc[i][1] = __ffma2_rn(a2, w23, acc[i][1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int px = pg * 4 + i;
        if (px < npix)
            store_all4(o.enc_out, p0 + px, cg * 4,
                       make_float4(fmaxf(acc[i][0].x, 0.f), fmaxf(acc[i][0].y, 0.f), fmaxf(acc[i][1].x, 0.f), fmaxf(acc[i][1].y, 0.f)));
    }
}

// RAFT-Stereo.  Plain: 128 threads = 32 pixels x 4 levels per chunk.  ENC: 256 threads = 64 pixels x 4 levels.
// CTAs are persistent over chunks (grid = a few CTAs per SM): the encoder weights are staged once per CTA.
template <int R, bool ENC>
__global__ void __launch_bounds__(ENC ? 256 : 128)
corr1d_lookup_kernel(ConstPyrPtrs pyr, int levels, float* __restrict__ coords_x,
                     const float* __restrict__ delta, int delta_C, float* __restrict__ flow,
                     LookupOut o, int64_t P, int HW, int W1) {
    constexpr int T = 2 * R + 1;
    constexpr int PIX = ENC ? LK_ENC_PIX : LK_PIX;
    constexpr int TS = DKT_MAX_LEVELS * T + 1;                 // plain: odd row stride, conflict-free column walks
    __shared__ float s_x[PIX];
    __shared__ __align__(16) float tile[ENC ? DKT_MAX_LEVELS * T * LK_ENC_TSP : LK_PIX * TS];
    __shared__ __align__(16) float ws[ENC ? DKT_MAX_LEVELS * T * LK_ENC_N : 4];
    const int C = levels * T;
    if (ENC) {
        for (int i = threadIdx.x; i < C * LK_ENC_N / 4; i += blockDim.x)
            reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(o.enc_w) + i);
    }
    const int64_t nchunks = (P + PIX - 1) / PIX;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t p0 = chunk * PIX;
        const int npix = (int)((P - p0) < PIX ? (P - p0) : PIX);
        if ((int)threadIdx.x < npix) {
            const int64_t p = p0 + threadIdx.x;
            float cx = coords_x[p];
            if (delta) {
                cx += delta[p * delta_C];
                coords_x[p] = cx;
            }
            if (flow) flow[p * 2] = cx - (float)(p % W1);
            s_x[threadIdx.x] = cx;
        }
        if (levels == 0) continue;
        __syncthreads();
        if (!(ENC && (o.dbg & 1))) {
            const int px = threadIdx.x >> 2, l = threadIdx.x & 3;
            if (px < npix && l < levels) {
                float taps[T];
                sample_row<R>(pyr.p[l] + (p0 + px) * pyr.w[l], pyr.w[l], s_x[px] * (1.f / (float)(1 << l)), taps);
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    if (ENC) tile[(l * T + k) * LK_ENC_TSP + px] = taps[k];
                    else tile[px * TS + l * T + k] = taps[k];
                }
            }
        }
        __syncthreads();
        if (ENC) { if (!(o.dbg & 2)) lookup_encode_tile(tile, C, ws, o, p0, npix); }
        else lookup_store_tile(tile, TS, C, o, p0, npix, HW);
        __syncthreads();                                       // tile and s_x are reused by the next chunk
    }
}

// IGEV: per pixel 2 levels x (Cg geometry channels + 1 init-corr row) row samples; the channel
// order of the reference is level-major: [geo(Cg*T), init(T)] per level (geometry.py:36-57).
// 256 threads walk the 32 * 2 * (Cg+1) row samples of the CTA's pixels.
struct GeoPtrs {
    const float* geo[2];
    const float* init[2];
};

template <int R, bool ENC>
__global__ void __launch_bounds__(256)
geo_lookup_kernel(GeoPtrs g, float* __restrict__ disp, const float* __restrict__ delta, int delta_C,
                  int Cg, int D, int W, LookupOut o, int64_t P, int HW) {
    constexpr int T = 2 * R + 1;
    constexpr int PIX = ENC ? LK_ENC_PIX : LK_PIX;
    extern __shared__ __align__(16) float gsm[];
    const int G = 2 * (Cg + 1);
    const int C = G * T;
    const int TS = C | 1;                     // plain: pixel-major rows
    float* s_d = gsm;                         // [PIX]
    float* tile = gsm + PIX;                  // plain [32][TS]; ENC k-major [C][LK_ENC_TSP]
    float* ws = tile + (ENC ? C * LK_ENC_TSP : ((LK_PIX * TS + 3) & ~3));   // ENC: [C][64]
    if (ENC) {
        for (int i = threadIdx.x; i < C * LK_ENC_N / 4; i += blockDim.x)
            reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(o.enc_w) + i);
    }
    const int64_t nchunks = (P + PIX - 1) / PIX;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t p0 = chunk * PIX;
        const int npix = (int)((P - p0) < PIX ? (P - p0) : PIX);
        if ((int)threadIdx.x < npix) {
            const int64_t p = p0 + threadIdx.x;
            float d = disp[p];
            if (delta) {
                d += delta[p * delta_C];
                disp[p] = d;
            }
            s_d[threadIdx.x] = d;
        }
        __syncthreads();
        // the gather is latency bound: a thread issues the 10 loads of up to GU row samples before interpolating any
        constexpr int GU = 3;
        const float* const geo0 = g.geo[0];
        const float* const geo1 = g.geo[1];
        const float* const init0 = g.init[0];
        const float* const init1 = g.init[1];
        for (int it0 = threadIdx.x; it0 < ((ENC && (o.dbg & 1)) ? 0 : npix * G); it0 += blockDim.x * GU) {
            float v[GU][2 * R + 2], av[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const int it = it0 + u * (int)blockDim.x;
                if (it >= npix * G) continue;
                const int px = it / G, gi = it - px * G;
                const int l = gi / (Cg + 1), j = gi - l * (Cg + 1);
                const int64_t p = p0 + px;
                const float d = s_d[px];
                const float inv = l ? 0.5f : 1.f;
                if (j < Cg) {
                    const int Dl = l ? D / 2 : D;
                    sample_row_load<R>((l ? geo1 : geo0) + (p * Cg + j) * Dl, Dl, d * inv, v[u], av[u]);
                } else {
                    const int Wl = l ? W / 2 : W;
                    const float x = (float)(p % W);
                    sample_row_load<R>((l ? init1 : init0) + p * Wl, Wl, x * inv - d * inv, v[u], av[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const int it = it0 + u * (int)blockDim.x;
                if (it >= npix * G) continue;
                const int px = it / G, gi = it - px * G;
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    const float tap = (1.f - av[u]) * v[u][k] + av[u] * v[u][k + 1];
                    if (ENC) tile[(gi * T + k) * LK_ENC_TSP + px] = tap;
                    else tile[px * TS + gi * T + k] = tap;
                }
            }
        }
        __syncthreads();
        if (ENC) { if (!(o.dbg & 2)) lookup_encode_tile(tile, C, ws, o, p0, npix); }
        else lookup_store_tile(tile, TS, C, o, p0, npix, HW);
        __syncthreads();                      // tile and s_d are reused by the next chunk
    }
}

// ---------------------------------------------------------------------------------------------
// IGEV geometry encoding volume: (B,C,D,H,W) -> level 0 (B,H,W,C,D) + pooled level 1 (.., D/2)
// 32x32 smem transpose of the [D][W] plane of one (b,c,y).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
geo_pool_kernel(const float* __restrict__ gev, float* __restrict__ geo0, float* __restrict__ geo1,
                int C, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int bcy = blockIdx.z;                  // (b*C + c)*H + y
    const int y = bcy % H;
    const int c = (bcy / H) % C;
    const int b = bcy / (H * C);
    const int d0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
    const float* src = gev + (((int64_t)b * C + c) * D) * H * W + (int64_t)y * W;
    for (int i = ty; i < 32; i += 8) {
        int d = d0 + i, w = w0 + tx;
        tile[i][tx] = (d < D && w < W) ? src[(int64_t)d * H * W + w] : 0.f;
    }
    __syncthreads();
    const int D1 = D / 2;
    for (int i = ty; i < 32; i += 8) {
        int w = w0 + i, d = d0 + tx;
        if (w >= W) continue;
        int64_t pix = ((int64_t)b * H + y) * W + w;
        if (d < D) geo0[(pix * C + c) * D + d] = tile[tx][i];
        if (tx < 16) {
            int dd = d0 / 2 + tx;
            if (dd < D1) geo1[(pix * C + c) * D1 + dd] = (tile[2 * tx][i] + tile[2 * tx + 1][i]) * 0.5f;
        }
    }
}

}  // namespace dkt

using namespace dkt;

extern "C" int dkt_corr1d_build_f32(const float* fmap1, const float* fmap2,
                                    int64_t sb, int64_t sd, int64_t sh, int64_t sw,
                                    float* const* pyr, int B, int D, int H, int W1, int W2,
                                    int levels, float scale, void* stream) {
    DKT_CHECK_ARG(fmap1 && fmap2 && pyr);
    DKT_CHECK_ARG(B > 0 && D > 0 && H > 0 && W1 > 0 && W2 > 0);
    if (levels < 1 || levels > DKT_MAX_LEVELS) return DKT_E_UNSUPPORTED;
    PyrPtrs pp;
    int w = W2;
    for (int l = 0; l < DKT_MAX_LEVELS; ++l) {
        pp.p[l] = l < levels ? pyr[l] : nullptr;
        pp.w[l] = w;
        if (l < levels) { DKT_CHECK_ARG(pyr[l] != nullptr && w > 0); }
        w /= 2;
    }
    dim3 grid(ceil_div(W2, CB_BN), ceil_div(W1, CB_BM), B * H);
    if (grid.z > 65535) return DKT_E_UNSUPPORTED;
    corr1d_build_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(fmap1, fmap2, sb, sd, sh, sw, pp,
                                                                      D, H, W1, W2, levels, scale);
    DKT_RETURN_LAST();
}

static int fill_plain_out(LookupOut& o, float* out, uint16_t* out_hi, uint16_t* out_lo,
                          int64_t ob, int64_t oc, int64_t op, int HW) {
    o = LookupOut{};
    o.f32 = out; o.hi = out_hi; o.lo = out_lo;
    o.ob = ob; o.oc = oc; o.op = op;
    const bool al = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out_hi) & 7) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out_lo) & 7) == 0);
    o.fast = (oc == 1 && op % 4 == 0 && ob == (int64_t)HW * op && al) ? 1 : 0;
    return 0;
}

static int fill_enc_out(LookupOut& o, const float* enc_w, const float* enc_b, const dkt_tensor* enc_out) {
    o = LookupOut{};
    if (!enc_w || !enc_b || !enc_out) return DKT_E_INVALID;
    if (!(enc_out->f32 || enc_out->hi) || enc_out->c_count != LK_ENC_N) return DKT_E_INVALID;
    if ((enc_out->C % 4) || (enc_out->c_begin % 4)) return DKT_E_ALIGNMENT;
    if ((reinterpret_cast<uintptr_t>(enc_w) & 15) || (reinterpret_cast<uintptr_t>(enc_b) & 15)) return DKT_E_ALIGNMENT;
    o.enc_w = enc_w; o.enc_b = enc_b; o.enc_out = *enc_out;
    static const int s_dbg = [] { const char* v = getenv("DKT_LOOKUP_PHASE"); return v ? atoi(v) : 0; }();
    o.dbg = s_dbg;
    return 0;
}

static int corr1d_lookup_launch(const float* const* pyr, int levels, int radius, float* coords_x,
                                const float* delta, int delta_C, float* flow, const LookupOut& o, bool enc, bool want_out,
                                int B, int H, int W1, int W2, void* stream) {
    DKT_CHECK_ARG(coords_x);
    DKT_CHECK_ARG(B > 0 && H > 0 && W1 > 0);
    if (levels < 0 || levels > DKT_MAX_LEVELS) return DKT_E_UNSUPPORTED;
    ConstPyrPtrs pp;
    int w = W2;
    for (int l = 0; l < DKT_MAX_LEVELS; ++l) {
        pp.p[l] = (want_out && l < levels) ? pyr[l] : nullptr;
        pp.w[l] = w;
        if (want_out && l < levels) { DKT_CHECK_ARG(pyr && pyr[l] != nullptr && w > 0); }
        w /= 2;
    }
    if (want_out && radius != 4) return DKT_E_UNSUPPORTED;   // configs/*/base.json: corr_radius = 4
    if (delta) DKT_CHECK_ARG(delta_C > 0);
    const int64_t P = (int64_t)B * H * W1;
    const int lv = want_out ? levels : 0;
    // persistent over pixel chunks: at most 8 CTAs per SM's worth of CTAs, each walking chunks blockIdx.x, +gridDim.x, ..
    const int64_t cap = (int64_t)device_sms() * 8;
    if (enc) {
        const int64_t chunks = ceil_div64(P, LK_ENC_PIX);
        corr1d_lookup_kernel<4, true><<<(unsigned)(chunks < cap ? chunks : cap), 256, 0, (cudaStream_t)stream>>>(
            pp, lv, coords_x, delta, delta_C, flow, o, P, H * W1, W1);
    } else {
        const int64_t chunks = ceil_div64(P, LK_PIX);
        corr1d_lookup_kernel<4, false><<<(unsigned)(chunks < 4 * cap ? chunks : 4 * cap), 128, 0, (cudaStream_t)stream>>>(
            pp, lv, coords_x, delta, delta_C, flow, o, P, H * W1, W1);
    }
    DKT_RETURN_LAST();
}

extern "C" int dkt_corr1d_lookup(const float* const* pyr, int levels, int radius,
                                 float* coords_x, const float* delta, int delta_C, float* flow,
                                 float* out, uint16_t* out_hi, uint16_t* out_lo,
                                 int64_t ob, int64_t oc, int64_t op,
                                 int B, int H, int W1, int W2, void* stream) {
    LookupOut o;
    fill_plain_out(o, out, out_hi, out_lo, ob, oc, op, H * W1);
    return corr1d_lookup_launch(pyr, levels, radius, coords_x, delta, delta_C, flow, o, false, out != nullptr,
                                B, H, W1, W2, stream);
}

extern "C" int dkt_corr1d_lookup_enc(const float* const* pyr, int levels, int radius,
                                     float* coords_x, const float* delta, int delta_C, float* flow,
                                     const float* enc_w, const float* enc_b, const dkt_tensor* enc_out,
                                     int B, int H, int W1, int W2, void* stream) {
    LookupOut o;
    int rc = fill_enc_out(o, enc_w, enc_b, enc_out);
    if (rc) return rc;
    DKT_CHECK_ARG(levels >= 1);
    return corr1d_lookup_launch(pyr, levels, radius, coords_x, delta, delta_C, flow, o, true, true, B, H, W1, W2, stream);
}

// ---- adjoint of the one-level lookup: corr_sampler.backward (reference core/corr.py:25-29) ----
// out[k] = (1-a) v[x0+k-r] + a v[x0+k-r+1]  =>  grad_v[j] = (1-a) g[j-x0+r] + a g[j-x0+r-1] (terms whose tap index falls
// outside [0, 2r] vanish).  One warp per (b, y, x1) row of the volume; lanes stride over W2, so every row is written in
// full, coalesced, zeros included -- no memset and no atomics (each row has exactly one owner).
__global__ void __launch_bounds__(256)
corr1d_lookup_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ coords_x, float* __restrict__ grad_vol,
                         int radius, int64_t rows, int HW, int W2) {
    const int lane = threadIdx.x & 31;
    const int taps = 2 * radius + 1;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const float x = coords_x[row];
        const bool finite = fabsf(x) < 1.0e9f;                 // also false for NaN
        const float xf = floorf(finite ? x : 0.f);
        const float a = finite ? x - xf : 0.f;
        const int x0 = (int)xf;
        const int64_t b = row / HW, p = row - b * HW;
        // grad_out is (B, taps, H, W1): tap k of this row sits HW elements after tap k-1
        float g = 0.f;
        if (finite && lane < taps) g = __ldg(grad_out + (b * taps + lane) * HW + p);
        float* dst = grad_vol + row * W2;
        for (int j0 = 0; j0 < W2; j0 += 32) {
            const int j = j0 + lane;
            const int k = j - x0 + radius;                     // tap whose LEFT sample is j
            const float g0 = __shfl_sync(0xffffffffu, g, k & 31);
            const float g1 = __shfl_sync(0xffffffffu, g, (k - 1) & 31);
            float v = 0.f;
            if (k >= 0 && k < taps) v = (1.f - a) * g0;
            if (k - 1 >= 0 && k - 1 < taps) v = fmaf(a, g1, v);
            if (j < W2) dst[j] = v;
        }
    }
}

extern "C" int dkt_corr1d_lookup_backward(const float* grad_out, const float* coords_x, int radius, float* grad_volume,
                                          int B, int H, int W1, int W2, void* stream) {
    DKT_CHECK_ARG(grad_out && coords_x && grad_volume);
    DKT_CHECK_ARG(B > 0 && H > 0 && W1 > 0 && W2 > 0);
    if (radius < 0 || 2 * radius + 1 > 32) return DKT_E_UNSUPPORTED;
    const int64_t rows = (int64_t)B * H * W1;
    if ((int64_t)H * W1 > 0x7fffffff) return DKT_E_UNSUPPORTED;
    const int64_t blocks = ceil_div64(rows, 8), cap = (int64_t)device_sms() * 16;
    corr1d_lookup_bwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        grad_out, coords_x, grad_volume, radius, rows, H * W1, W2);
    DKT_RETURN_LAST();
}

extern "C" int dkt_geo_pool(const float* gev, float* geo0, float* geo1, int B, int C, int D, int H, int W,
                            void* stream) {
    DKT_CHECK_ARG(gev && geo0 && geo1);
    DKT_CHECK_ARG(B > 0 && C > 0 && D > 1 && H > 0 && W > 0);
    if ((int64_t)B * C * H > 65535) {
        // split over batch to stay inside gridDim.z
        for (int b = 0; b < B; ++b) {
            int rc = dkt_geo_pool(gev + (int64_t)b * C * D * H * W, geo0 + (int64_t)b * H * W * C * D,
                                  geo1 + (int64_t)b * H * W * C * (D / 2), 1, C, D, H, W, stream);
            if (rc) return rc;
        }
        return 0;
    }
    dim3 grid(ceil_div(W, 32), ceil_div(D, 32), B * C * H);
    geo_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gev, geo0, geo1, C, D, H, W);
    DKT_RETURN_LAST();
}

static int geo_lookup_launch(const float* geo0, const float* geo1, const float* init0, const float* init1,
                             float* disp, const float* delta, int delta_C, int radius, int C, int D,
                             const LookupOut& o, bool enc, int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(geo0 && geo1 && init0 && init1 && disp);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 1 && C > 0 && D > 1);
    if (radius != 4) return DKT_E_UNSUPPORTED;
    if (delta) DKT_CHECK_ARG(delta_C > 0);
    const int64_t P = (int64_t)B * H * W;
    const int Cout = 2 * (C + 1) * 9;
    const int TS = Cout | 1;
    size_t smem = enc ? (size_t)(LK_ENC_PIX + Cout * LK_ENC_TSP + Cout * LK_ENC_N) * 4
                      : (size_t)(LK_PIX + ((LK_PIX * TS + 3) & ~3)) * 4;
    if (smem > 200 * 1024) return DKT_E_UNSUPPORTED;
    DKT_ENSURE_SMEM(200 * 1024, geo_lookup_kernel<4, true>);
    DKT_ENSURE_SMEM(200 * 1024, geo_lookup_kernel<4, false>);
    GeoPtrs g{{geo0, geo1}, {init0, init1}};
    if (enc) {
        // persistent: as many CTAs as fit (shared memory bound), each walking 64-pixel chunks with the weights staged once
        const int per_sm = (int)((227 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((227 * 1024) / (smem + 1024));
        const int64_t chunks = ceil_div64(P, LK_ENC_PIX), cap = (int64_t)device_sms() * per_sm;
        geo_lookup_kernel<4, true><<<(unsigned)(chunks < cap ? chunks : cap), 256, smem, (cudaStream_t)stream>>>(
            g, disp, delta, delta_C, C, D, W, o, P, H * W);
    } else {
        const int64_t chunks = ceil_div64(P, LK_PIX), cap = (int64_t)device_sms() * 32;
        geo_lookup_kernel<4, false><<<(unsigned)(chunks < cap ? chunks : cap), 256, smem, (cudaStream_t)stream>>>(
            g, disp, delta, delta_C, C, D, W, o, P, H * W);
    }
    DKT_RETURN_LAST();
}

extern "C" int dkt_geo_lookup(const float* geo0, const float* geo1, const float* init0, const float* init1,
                              float* disp, const float* delta, int delta_C, int radius, int C, int D,
                              float* out, uint16_t* out_hi, uint16_t* out_lo,
                              int64_t ob, int64_t oc, int64_t op,
                              int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(out);
    LookupOut o;
    fill_plain_out(o, out, out_hi, out_lo, ob, oc, op, H * W);
    return geo_lookup_launch(geo0, geo1, init0, init1, disp, delta, delta_C, radius, C, D, o, false, B, H, W, stream);
}

extern "C" int dkt_geo_lookup_enc(const float* geo0, const float* geo1, const float* init0, const float* init1,
                                  float* disp, const float* delta, int delta_C, int radius, int C, int D,
                                  const float* enc_w, const float* enc_b, const dkt_tensor* enc_out,
                                  int B, int H, int W, void* stream) {
    LookupOut o;
    int rc = fill_enc_out(o, enc_w, enc_b, enc_out);
    if (rc) return rc;
    return geo_lookup_launch(geo0, geo1, init0, init1, disp, delta, delta_C, radius, C, D, o, true, B, H, W, stream);
}
