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
// One thread per pixel; the (optional) coordinate update is fused in front.
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

template <int R>
__global__ void __launch_bounds__(128)
corr1d_lookup_kernel(ConstPyrPtrs pyr, int levels, float* __restrict__ coords_x,
                     const float* __restrict__ delta, int delta_C, float* __restrict__ flow,
                     float* __restrict__ out, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                     int64_t ob, int64_t oc, int64_t op, int64_t P, int HW, int W1) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float cx = coords_x[p];
    if (delta) {
        cx += delta[p * delta_C];
        coords_x[p] = cx;
    }
    if (flow) flow[p * 2] = cx - (float)(p % W1);
    if (!out) return;
    const int64_t b = p / HW;
    const int64_t pix = p - b * HW;
    const int64_t obase = b * ob + pix * op;
    float inv = 1.f;
    for (int l = 0; l < levels; ++l) {
        float taps[2 * R + 1];
        sample_row<R>(pyr.p[l] + p * pyr.w[l], pyr.w[l], cx * inv, taps);
        inv *= 0.5f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; ++k) {
            int64_t o = obase + (int64_t)(l * (2 * R + 1) + k) * oc;
            out[o] = taps[k];
            if (out_hi) {
                uint16_t h, lo_;
                split_bf16(taps[k], h, lo_);
                out_hi[o] = h;
                if (out_lo) out_lo[o] = lo_;
            }
        }
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

// One thread per (pixel, group); group g in [0, 2*(C+1)): level = g / (C+1), j = g % (C+1):
// j < C -> geometry channel j sampled at disp/2^l ; j == C -> init-corr sampled at (x - disp)/2^l.
template <int R>
__global__ void __launch_bounds__(128)
geo_lookup_kernel(const float* __restrict__ geo0, const float* __restrict__ geo1,
                  const float* __restrict__ init0, const float* __restrict__ init1,
                  const float* __restrict__ disp, int C, int D, int W,
                  float* __restrict__ out, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                  int64_t ob, int64_t oc, int64_t op, int64_t P, int HW) {
    const int G = 2 * (C + 1);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * G) return;
    const int64_t p = t / G;
    const int g = (int)(t - p * G);
    const int l = g / (C + 1), j = g % (C + 1);
    const float d = disp[p];
    const float inv = l ? 0.5f : 1.f;
    float taps[2 * R + 1];
    if (j < C) {
        const int Dl = l ? D / 2 : D;
        const float* row = (l ? geo1 : geo0) + (p * C + j) * Dl;
        sample_row<R>(row, Dl, d * inv, taps);
    } else {
        const int Wl = l ? W / 2 : W;
        const float* row = (l ? init1 : init0) + p * Wl;
        const float x = (float)(p % W);
        sample_row<R>(row, Wl, x * inv - d * inv, taps);
    }
    const int64_t b = p / HW;
    const int64_t obase = b * ob + (p - b * HW) * op;
    const int cbase = l * (C + 1) * (2 * R + 1) + j * (2 * R + 1);
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
        int64_t o = obase + (int64_t)(cbase + k) * oc;
        out[o] = taps[k];
        if (out_hi) {
            uint16_t h, lo_;
            split_bf16(taps[k], h, lo_);
            out_hi[o] = h;
            if (out_lo) out_lo[o] = lo_;
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

extern "C" int dkt_corr1d_lookup(const float* const* pyr, int levels, int radius,
                                 float* coords_x, const float* delta, int delta_C, float* flow,
                                 float* out, uint16_t* out_hi, uint16_t* out_lo,
                                 int64_t ob, int64_t oc, int64_t op,
                                 int B, int H, int W1, int W2, void* stream) {
    DKT_CHECK_ARG(coords_x);
    DKT_CHECK_ARG(B > 0 && H > 0 && W1 > 0);
    if (levels < 0 || levels > DKT_MAX_LEVELS) return DKT_E_UNSUPPORTED;
    ConstPyrPtrs pp;
    int w = W2;
    for (int l = 0; l < DKT_MAX_LEVELS; ++l) {
        pp.p[l] = (out && l < levels) ? pyr[l] : nullptr;
        pp.w[l] = w;
        if (out && l < levels) { DKT_CHECK_ARG(pyr && pyr[l] != nullptr && w > 0); }
        w /= 2;
    }
    if (out && radius != 4) return DKT_E_UNSUPPORTED;   // configs/*/base.json: corr_radius = 4
    if (delta) DKT_CHECK_ARG(delta_C > 0);
    const int64_t P = (int64_t)B * H * W1;
    const int threads = 128;
    const int64_t blocks = ceil_div64(P, threads);
    corr1d_lookup_kernel<4><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        pp, levels, coords_x, delta, delta_C, flow, out, out_hi, out_lo, ob, oc, op, P, H * W1, W1);
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

extern "C" int dkt_geo_lookup(const float* geo0, const float* geo1, const float* init0, const float* init1,
                              const float* disp, int radius, int C, int D,
                              float* out, uint16_t* out_hi, uint16_t* out_lo,
                              int64_t ob, int64_t oc, int64_t op,
                              int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(geo0 && geo1 && init0 && init1 && disp && out);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 1 && C > 0 && D > 1);
    if (radius != 4) return DKT_E_UNSUPPORTED;
    const int64_t P = (int64_t)B * H * W;
    const int64_t T = P * 2 * (C + 1);
    geo_lookup_kernel<4><<<(unsigned)ceil_div64(T, 128), 128, 0, (cudaStream_t)stream>>>(
        geo0, geo1, init0, init1, disp, C, D, W, out, out_hi, out_lo, ob, oc, op, P, H * W);
    DKT_RETURN_LAST();
}
