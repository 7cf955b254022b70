// Cross-scale plumbing of the multi-level GRU (pool2x / interp), final upsamplers (K4) and
// layout / precision conversions.  All HBM-bound elementwise kernels over NHWC activations.
// Contracts and reference citations: include/dkt_stereo_b200.h.
#include "common.cuh"

namespace dkt {

// ---- pool2x: 3x3 / stride 2 / pad 1 average, divisor 9 (count_include_pad) -----------------
// grid: x = ceil(Wd * C4 / 256), y = B * Hd (one output row per blockIdx.y): no 64-bit div/mod per thread
__global__ void __launch_bounds__(256)
pool2x_kernel(const float* __restrict__ src, int sC, int sc0, dkt_tensor dst, int C4,
              int Hs, int Ws, int Hd, int Wd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Wd * C4) return;
    const int xo = i / C4, q = i - xo * C4;
    const int row = blockIdx.y;                   // b * Hd + yo
    const int b = row / Hd, yo = row - b * Hd;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* base = src + (int64_t)b * Hs * Ws * sC + sc0 + q * 4;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * yo + dy;
        if (y < 0 || y >= Hs) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int x = 2 * xo + dx;
            if (x < 0 || x >= Ws) continue;
            float4 v = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)y * Ws + x) * sC));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    // ATen divides the window sum by 9; keep a true division for bit-level agreement
    s.x /= 9.f; s.y /= 9.f; s.z /= 9.f; s.w /= 9.f;
    store_all4(dst, (int64_t)row * Wd + xo, q * 4, s);
}

// ---- bilinear resize, align_corners=True ------------------------------------------------------
// grid: x = ceil(Wd * C4 / 256), y = ceil(B * Hd / IR): a thread produces IR output rows of one (x, channel group)
// column, loads first -- with one row per thread the kernel was bound by the block launch rate (32640 tiny blocks)
constexpr int IR = 4;
__global__ void __launch_bounds__(256)
interp_kernel(const float* __restrict__ src, int sC, int sc0, dkt_tensor dst, int C4,
              int Hs, int Ws, int Hd, int Wd, int rows_total, float sy, float sx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Wd * C4) return;
    const int xo = i / C4, q = i - xo * C4;
    const float fx = sx * xo;
    const int x0 = (int)fx;
    const int x1 = x0 + (x0 < Ws - 1);
    const float lx = fx - x0, hx = 1.f - lx;
    // Fast path (block-uniform): the IR output rows lie in one image and their source rows span at most 4 rows (always
    // the case when upsampling by ~2): each distinct source row is loaded once, 8 loads instead of 16.
    {
        const int row0 = blockIdx.y * IR;
        const int b0 = row0 / Hd, yo0 = row0 - b0 * Hd;
        const int yb = (int)(sy * yo0);
        const int yl = (int)(sy * (yo0 + IR - 1));
        if (row0 + IR - 1 < rows_total && yo0 + IR - 1 < Hd && yl + (yl < Hs - 1) - yb <= 3) {
            const float* base = src + (int64_t)b0 * Hs * Ws * sC + sc0 + q * 4;
            float4 ra[4], rb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = min(yb + j, Hs - 1);
                ra[j] = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)yy * Ws + x0) * sC));
                rb[j] = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)yy * Ws + x1) * sC));
            }
#pragma unroll
            for (int r = 0; r < IR; ++r) {
                const float fy = sy * (yo0 + r);
                const int y0 = (int)fy;
                const int i0 = y0 - yb, i1 = i0 + (y0 < Hs - 1);
                const float l = fy - y0, hy = 1.f - l;
                const float4 t0 = i0 == 0 ? ra[0] : i0 == 1 ? ra[1] : i0 == 2 ? ra[2] : ra[3];
                const float4 t1 = i0 == 0 ? rb[0] : i0 == 1 ? rb[1] : i0 == 2 ? rb[2] : rb[3];
                const float4 u0 = i1 == 0 ? ra[0] : i1 == 1 ? ra[1] : i1 == 2 ? ra[2] : ra[3];
                const float4 u1 = i1 == 0 ? rb[0] : i1 == 1 ? rb[1] : i1 == 2 ? rb[2] : rb[3];
                float4 o;
                o.x = hy * (hx * t0.x + lx * t1.x) + l * (hx * u0.x + lx * u1.x);
                o.y = hy * (hx * t0.y + lx * t1.y) + l * (hx * u0.y + lx * u1.y);
                o.z = hy * (hx * t0.z + lx * t1.z) + l * (hx * u0.z + lx * u1.z);
                o.w = hy * (hx * t0.w + lx * t1.w) + l * (hx * u0.w + lx * u1.w);
                store_all4(dst, (int64_t)(row0 + r) * Wd + xo, q * 4, o);
            }
            return;
        }
    }
    float4 v00[IR], v01[IR], v10[IR], v11[IR];
    float ly[IR];
#pragma unroll
    for (int r = 0; r < IR; ++r) {
        const int row = blockIdx.y * IR + r;          // b * Hd + yo
        if (row >= rows_total) continue;
        const int b = row / Hd, yo = row - b * Hd;
        const float fy = sy * yo;
        const int y0 = (int)fy;
        const int y1 = y0 + (y0 < Hs - 1);
        ly[r] = fy - y0;
        const float* base = src + (int64_t)b * Hs * Ws * sC + sc0 + q * 4;
        v00[r] = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)y0 * Ws + x0) * sC));
        v01[r] = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)y0 * Ws + x1) * sC));
        v10[r] = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)y1 * Ws + x0) * sC));
        v11[r] = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)y1 * Ws + x1) * sC));
    }
#pragma unroll
    for (int r = 0; r < IR; ++r) {
        const int row = blockIdx.y * IR + r;
        if (row >= rows_total) continue;
        const float hy = 1.f - ly[r];
        float4 o;
        o.x = hy * (hx * v00[r].x + lx * v01[r].x) + ly[r] * (hx * v10[r].x + lx * v11[r].x);
        o.y = hy * (hx * v00[r].y + lx * v01[r].y) + ly[r] * (hx * v10[r].y + lx * v11[r].y);
        o.z = hy * (hx * v00[r].z + lx * v01[r].z) + ly[r] * (hx * v10[r].z + lx * v11[r].z);
        o.w = hy * (hx * v00[r].w + lx * v01[r].w) + ly[r] * (hx * v10[r].w + lx * v11[r].w);
        store_all4(dst, (int64_t)row * Wd + xo, q * 4, o);
    }
}

// ---- K4 RAFT: convex combination upsampling ---------------------------------------------------
// one thread per (low-res pixel, sub-position ij); mask channel = k*f*f + ij
__global__ void __launch_bounds__(256)
convex_upsample_kernel(const float* __restrict__ flow, int flow_C, const float* __restrict__ mask,
                       float* __restrict__ out, int H, int W, int f, int64_t total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int ff = f * f;
    const int ij = (int)(t % ff);
    const int64_t p = t / ff;
    const int x = (int)(p % W);
    int64_t r = p / W;
    const int y = (int)(r % H);
    const int64_t b = r / H;
    const float* m = mask + p * (9 * ff) + ij;
    float w[9];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) { w[k] = __ldg(m + k * ff); mx = fmaxf(mx, w[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { w[k] = expf(w[k] - mx); den += w[k]; }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
            v = (float)f * __ldg(flow + ((b * H + yy) * (int64_t)W + xx) * flow_C);
        acc += (w[k] / den) * v;
    }
    const int i = ij / f, j = ij % f;
    out[(b * (H * f) + (y * f + i)) * (int64_t)(W * f) + x * f + j] = acc;
}

// ---- K4 IGEV: learned-weight upsampling (weights already softmaxed) ----------------------------
__global__ void __launch_bounds__(256)
context_upsample_kernel(const float* __restrict__ disp, const float* __restrict__ wts, float* __restrict__ out,
                        float in_scale, float out_scale, int H, int W, int64_t total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int W4 = W * 4, H4 = H * 4;
    const int X = (int)(t % W4);
    int64_t r = t / W4;
    const int Y = (int)(r % H4);
    const int64_t b = r / H4;
    const int y = Y >> 2, x = X >> 2;
    const int64_t plane = (int64_t)H4 * W4;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = in_scale * __ldg(disp + (b * H + yy) * (int64_t)W + xx);
        acc += v * __ldg(wts + (b * 9 + k) * plane + (int64_t)Y * W4 + X);
    }
    out[t] = out_scale * acc;
}

// ---- IGEV upsample_disp plumbing -------------------------------------------------------------------
// sub-pixel rearrangement of a deconv-as-conv result: one thread per (input pixel, parity, 4-channel group)
__global__ void __launch_bounds__(256)
pixel_shuffle2_kernel(const float* __restrict__ src, int src_C, int group, dkt_tensor dst, int C4, int H, int W, int64_t total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int q = (int)(t % C4);
    int64_t r = t / C4;
    const int par = (int)(r & 3);
    r >>= 2;                                      // input pixel index (b*H + y)*W + x
    const int x = (int)(r % W);
    const int64_t by = r / W;                     // b*H + y
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * src_C + par * group + q * 4));
    const int64_t po = (by * 2 + (par >> 1)) * (2 * W) + 2 * x + (par & 1);      // b*(2H) + 2y+py rows of 2W pixels
    store_all4(dst, po, q * 4, v);
}

// softmax over the 9 logits of a full-resolution pixel + the 3x3 neighbourhood combination of context_upsample
__global__ void __launch_bounds__(256)
context_upsample_logits_kernel(const float* __restrict__ disp, const float* __restrict__ logits, int logit_C,
                               float* __restrict__ out, float in_scale, float out_scale, int H, int W, int64_t total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int W4 = W * 4, H4 = H * 4;
    const int X = (int)(t % W4);
    int64_t r = t / W4;
    const int Y = (int)(r % H4);
    const int64_t b = r / H4;
    const int y = Y >> 2, x = X >> 2;
    // logits live at half resolution, parity-major: pixel (Y >> 1, X >> 1), group (Y & 1) * 2 + (X & 1)
    const float* lp = logits + ((b * (2 * H) + (Y >> 1)) * (int64_t)(2 * W) + (X >> 1)) * logit_C + (((Y & 1) << 1) | (X & 1)) * 9;
    float l[9], mx = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { l[k] = __ldg(lp + k); mx = fmaxf(mx, l[k]); }
    float den = 0.f, acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float e = expf(l[k] - mx);
        den += e;
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc = fmaf(e, in_scale * __ldg(disp + (b * H + yy) * (int64_t)W + xx), acc);
    }
    out[t] = out_scale * (acc / den);
}

// ---- layout conversions ------------------------------------------------------------------------
// generic strided (b, c, y, x) fp32 source -> NHWC slice, 32x32 smem transpose over (c, x) of one (b,y)
__global__ void __launch_bounds__(256)
to_nhwc_kernel(const float* __restrict__ src, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
               const float* __restrict__ bias, dkt_tensor dst, int C, int H, int W) {
    __shared__ float tile[32][33];
    const int by = blockIdx.z;
    const int y = by % H;
    const int64_t b = by / H;
    const int c0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const float* base = src + b * sb + y * sh;
    if (sw == 1 || sc != 1) {
        for (int i = ty; i < 32; i += 8) {     // rows = c, fast = x
            int c = c0 + i, x = x0 + tx;
            tile[i][tx] = (c < C && x < W) ? __ldg(base + c * sc + x * sw) : 0.f;
        }
    } else {
        for (int i = ty; i < 32; i += 8) {     // rows = x, fast = c (channels_last source)
            int x = x0 + i, c = c0 + tx;
            tile[tx][i] = (c < C && x < W) ? __ldg(base + c * sc + x * sw) : 0.f;
        }
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int x = x0 + i, c = c0 + tx;
        if (x < W && c < C) {
            float v = tile[tx][i];
            if (bias) v += __ldg(bias + c);
            store_all(dst, (b * H + y) * (int64_t)W + x, c, v);
        }
    }
}

__global__ void __launch_bounds__(256)
to_nchw_kernel(const float* __restrict__ src, int sC, int sc0, float* __restrict__ dst, int C, int H, int W) {
    __shared__ float tile[32][33];
    const int by = blockIdx.z;
    const int y = by % H;
    const int64_t b = by / H;
    const int c0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    for (int i = ty; i < 32; i += 8) {
        int x = x0 + i, c = c0 + tx;
        tile[i][tx] = (x < W && c < C) ? __ldg(src + ((b * H + y) * (int64_t)W + x) * sC + sc0 + c) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, x = x0 + tx;
        if (c < C && x < W) dst[((b * C + c) * H + y) * (int64_t)W + x] = tile[tx][i];
    }
}

// ---- depthwise 3x3 convolution, NHWC (MobileNetV2 of the IGEV feature pyramid: timm conv_dw + folded BatchNorm + ReLU6) ----
// thread = (4 consecutive output pixels of a row, group of 4 channels): per filter row the 6 (stride 1) / 9 (stride 2) input
// columns are loaded once (16-byte loads, zero outside the image) and feed all four outputs; fp32 FMAs, bias, clamp, every
// non-null precision of dst written.  in_max clamps the INPUT on load: the expand conv that produced it applied ReLU in its
// epilogue, min(., 6) completes its ReLU6 here.
template <int STRIDE>
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const float* __restrict__ src, int sC, int sc0, const float* __restrict__ w, const float* __restrict__ bias,
                 float in_max, float out_min, float out_max, dkt_tensor dst, int C4, int Hin, int Win, int H, int W,
                 int64_t total) {
    constexpr int XO = 4, NIN = (XO - 1) * STRIDE + 3;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int WQ = (W + XO - 1) / XO;
    const int g = (int)(t % C4);
    const int64_t q = t / C4;
    const int xq = (int)(q % WQ), y = (int)((q / WQ) % H);
    const int64_t b = q / ((int64_t)WQ * H);
    const int c = g * 4, x0 = xq * XO;
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c));
    float4 acc[XO];
#pragma unroll
    for (int i = 0; i < XO; ++i) acc[i] = bv;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y * STRIDE + ky - 1;
        if (iy < 0 || iy >= Hin) continue;
        const float* row = src + ((b * Hin + iy) * (int64_t)Win) * sC + sc0 + c;
        float4 v[NIN];
#pragma unroll
        for (int j = 0; j < NIN; ++j) {
            const int ix = x0 * STRIDE + j - 1;
            v[j] = (ix >= 0 && ix < Win) ? __ldg(reinterpret_cast<const float4*>(row + (int64_t)ix * sC)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j].x = fminf(v[j].x, in_max); v[j].y = fminf(v[j].y, in_max); v[j].z = fminf(v[j].z, in_max); v[j].w = fminf(v[j].w, in_max);
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float4 k = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(ky * 3 + kx) * C4 * 4 + c));   // [tap][C]
#pragma unroll
            for (int i = 0; i < XO; ++i) {
                const float4 u = v[i * STRIDE + kx];
                acc[i].x = fmaf(u.x, k.x, acc[i].x); acc[i].y = fmaf(u.y, k.y, acc[i].y);
                acc[i].z = fmaf(u.z, k.z, acc[i].z); acc[i].w = fmaf(u.w, k.w, acc[i].w);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < XO; ++i) {
        const int x = x0 + i;
        if (x >= W) break;
        float4 a = acc[i];
        a.x = fminf(fmaxf(a.x, out_min), out_max); a.y = fminf(fmaxf(a.y, out_min), out_max);
        a.z = fminf(fmaxf(a.z, out_min), out_max); a.w = fminf(fmaxf(a.w, out_min), out_max);
        store_all4(dst, (b * H + y) * (int64_t)W + x, c, a);
    }
}

// ---- depth-padded NDHWC <-> NCDHW (IGEV hourglass layers on the 2-D tensor-core conv: depth planes are its images) ----
// src (B,C,D,H,W) fp32 -> 16-bit (hi, lo) (B,D+2,H,W,C), interior planes 1..D (planes 0 and D+1 are the conv's zero padding
// in depth: zeroed once by the caller, never written here); 32x32 smem transpose over (c, x) of one (b, d, y)
__global__ void __launch_bounds__(256)
ncdhw_to_ndhwc_pad_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                          int C, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int bdy = blockIdx.z;
    const int y = bdy % H, d = (bdy / H) % D;
    const int64_t b = bdy / (H * D);
    const int c0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    for (int i = ty; i < 32; i += 8) {          // rows = c, fast = x
        const int c = c0 + i, x = x0 + tx;
        tile[i][tx] = (c < C && x < W) ? __ldg(src + (((b * C + c) * D + d) * H + y) * (int64_t)W + x) : 0.f;
    }
    __syncthreads();
    const dkt_tensor dst{nullptr, hi, lo, C, 0, C};
    for (int i = ty; i < 32; i += 8) {
        const int x = x0 + i, c = c0 + tx;
        if (x < W && c < C) store_all(dst, ((b * (D + 2) + d + 1) * H + y) * (int64_t)W + x, c, tile[tx][i]);
    }
}

// src fp32 (B,D+2,H,W,C) interior planes -> dst (B,C,D,H,W) fp32, times sigmoid(att[b,c,y,x]) when att != NULL
// (FeatureAtt, reference meta_arch/igev_stereo/submodule.py:227-240)
__global__ void __launch_bounds__(256)
ndhwc_pad_to_ncdhw_kernel(const float* __restrict__ src, const float* __restrict__ att, float* __restrict__ dst,
                          int C, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int bdy = blockIdx.z;
    const int y = bdy % H, d = (bdy / H) % D;
    const int64_t b = bdy / (H * D);
    const int c0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    for (int i = ty; i < 32; i += 8) {          // rows = x, fast = c
        const int x = x0 + i, c = c0 + tx;
        tile[i][tx] = (x < W && c < C) ? __ldg(src + (((b * (D + 2) + d + 1) * H + y) * (int64_t)W + x) * C + c) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, x = x0 + tx;
        if (c < C && x < W) {
            float v = tile[tx][i];
            if (att) v *= sigmoidf_acc(__ldg(att + ((b * C + c) * H + y) * (int64_t)W + x));
            dst[(((b * C + c) * D + d) * H + y) * (int64_t)W + x] = v;
        }
    }
}

}  // namespace dkt

using namespace dkt;

extern "C" int dkt_dwconv3x3(const dkt_tensor* src, const float* weight, const float* bias, float in_max, float out_min,
                             float out_max, const dkt_tensor* dst, int B, int Hin, int Win, int stride, void* stream) {
    DKT_CHECK_ARG(src && src->f32 && weight && bias && dst && (dst->f32 || dst->hi));
    DKT_CHECK_ARG(B > 0 && Hin > 0 && Win > 0 && (stride == 1 || stride == 2));
    const int C = src->c_count;
    DKT_CHECK_ARG(C > 0 && dst->c_count == C);
    if ((C % 4) || (src->C % 4) || (src->c_begin % 4) || (dst->C % 4) || (dst->c_begin % 4)) return DKT_E_ALIGNMENT;
    if ((reinterpret_cast<uintptr_t>(src->f32) & 15) || (reinterpret_cast<uintptr_t>(weight) & 15) ||
        (reinterpret_cast<uintptr_t>(bias) & 15) || (reinterpret_cast<uintptr_t>(dst->f32) & 15) ||
        (reinterpret_cast<uintptr_t>(dst->hi) & 7) || (reinterpret_cast<uintptr_t>(dst->lo) & 7))
        return DKT_E_ALIGNMENT;
    const int H = (Hin - 1) / stride + 1, W = (Win - 1) / stride + 1;        // kernel 3, padding 1
    const int64_t total = (int64_t)B * H * ((W + 3) / 4) * (C / 4);
    if (stride == 1)
        dwconv3x3_kernel<1><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
            src->f32, src->C, src->c_begin, weight, bias, in_max, out_min, out_max, *dst, C / 4, Hin, Win, H, W, total);
    else
        dwconv3x3_kernel<2><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
            src->f32, src->C, src->c_begin, weight, bias, in_max, out_min, out_max, *dst, C / 4, Hin, Win, H, W, total);
    DKT_RETURN_LAST();
}

extern "C" int dkt_ncdhw_to_ndhwc_pad(const float* src, uint16_t* hi, uint16_t* lo, int B, int C, int D, int H, int W,
                                      void* stream) {
    DKT_CHECK_ARG(src && hi && lo);
    DKT_CHECK_ARG(B > 0 && C > 0 && D > 0 && H > 0 && W > 0);
    if ((int64_t)B * D * H > 65535) return DKT_E_UNSUPPORTED;
    dim3 grid(ceil_div(W, 32), ceil_div(C, 32), B * D * H);
    ncdhw_to_ndhwc_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, hi, lo, C, D, H, W);
    DKT_RETURN_LAST();
}

extern "C" int dkt_ndhwc_pad_to_ncdhw(const float* src, const float* att, float* dst, int B, int C, int D, int H, int W,
                                      void* stream) {
    DKT_CHECK_ARG(src && dst);
    DKT_CHECK_ARG(B > 0 && C > 0 && D > 0 && H > 0 && W > 0);
    if ((int64_t)B * D * H > 65535) return DKT_E_UNSUPPORTED;
    dim3 grid(ceil_div(W, 32), ceil_div(C, 32), B * D * H);
    ndhwc_pad_to_ncdhw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, att, dst, C, D, H, W);
    DKT_RETURN_LAST();
}

static int check_slice(const dkt_tensor* t, bool need_f32) {
    if (!t) return DKT_E_INVALID;
    if (need_f32 && !t->f32) return DKT_E_INVALID;
    if (!t->f32 && !t->hi) return DKT_E_INVALID;
    if (t->c_count <= 0 || t->c_begin < 0 || t->c_begin + t->c_count > t->C) return DKT_E_INVALID;
    return 0;
}

static int check_vec4(const dkt_tensor* t) {
    if ((t->C % 4) || (t->c_begin % 4) || (t->c_count % 4)) return DKT_E_ALIGNMENT;
    if ((reinterpret_cast<uintptr_t>(t->f32) & 15) || (reinterpret_cast<uintptr_t>(t->hi) & 7) ||
        (reinterpret_cast<uintptr_t>(t->lo) & 7))
        return DKT_E_ALIGNMENT;
    return 0;
}

extern "C" int dkt_pool2x(const dkt_tensor* src, const dkt_tensor* dst, int B, int Hs, int Ws, int Hd, int Wd,
                          void* stream) {
    int rc;
    if ((rc = check_slice(src, true)) || (rc = check_slice(dst, false))) return rc;
    if ((rc = check_vec4(src)) || (rc = check_vec4(dst))) return rc;
    DKT_CHECK_ARG(B > 0 && Hs > 0 && Ws > 0);
    DKT_CHECK_ARG(src->c_count == dst->c_count);
    DKT_CHECK_ARG(Hd == (Hs - 1) / 2 + 1 && Wd == (Ws - 1) / 2 + 1);
    const int C4 = src->c_count / 4;
    if ((int64_t)B * Hd > 65535 || (int64_t)Wd * C4 > 0x7fffffff) return DKT_E_UNSUPPORTED;
    pool2x_kernel<<<dim3((unsigned)ceil_div(Wd * C4, 256), (unsigned)(B * Hd)), 256, 0, (cudaStream_t)stream>>>(
        src->f32, src->C, src->c_begin, *dst, C4, Hs, Ws, Hd, Wd);
    DKT_RETURN_LAST();
}

extern "C" int dkt_interp(const dkt_tensor* src, const dkt_tensor* dst, int B, int Hs, int Ws, int Hd, int Wd,
                          void* stream) {
    int rc;
    if ((rc = check_slice(src, true)) || (rc = check_slice(dst, false))) return rc;
    if ((rc = check_vec4(src)) || (rc = check_vec4(dst))) return rc;
    DKT_CHECK_ARG(B > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0);
    DKT_CHECK_ARG(src->c_count == dst->c_count);
    const int C4 = src->c_count / 4;
    if ((int64_t)B * Hd > 65535 * IR || (int64_t)Wd * C4 > 0x7fffffff) return DKT_E_UNSUPPORTED;
    const float sy = Hd > 1 ? (float)(Hs - 1) / (float)(Hd - 1) : 0.f;
    const float sx = Wd > 1 ? (float)(Ws - 1) / (float)(Wd - 1) : 0.f;
    interp_kernel<<<dim3((unsigned)ceil_div(Wd * C4, 256), (unsigned)ceil_div(B * Hd, IR)), 256, 0, (cudaStream_t)stream>>>(
        src->f32, src->C, src->c_begin, *dst, C4, Hs, Ws, Hd, Wd, B * Hd, sy, sx);
    DKT_RETURN_LAST();
}

extern "C" int dkt_convex_upsample(const float* flow, int flow_C, const float* mask, float* out,
                                   int B, int H, int W, int factor, void* stream) {
    DKT_CHECK_ARG(flow && mask && out && flow_C > 0);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && factor > 0);
    const int64_t total = (int64_t)B * H * W * factor * factor;
    convex_upsample_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
        flow, flow_C, mask, out, H, W, factor, total);
    DKT_RETURN_LAST();
}

extern "C" int dkt_context_upsample(const float* disp, const float* weights, float* out, float in_scale,
                                    float out_scale, int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(disp && weights && out);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0);
    const int64_t total = (int64_t)B * H * W * 16;
    context_upsample_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
        disp, weights, out, in_scale, out_scale, H, W, total);
    DKT_RETURN_LAST();
}

extern "C" int dkt_pixel_shuffle2(const float* src, int src_C, int group, const dkt_tensor* dst, int B, int H, int W,
                                  void* stream) {
    DKT_CHECK_ARG(src);
    int rc = check_slice(dst, false);
    if (rc) return rc;
    rc = check_vec4(dst);
    if (rc) return rc;
    const int C = dst->c_count;
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && group >= C && src_C >= 3 * group + C);
    if ((C % 4) || (group % 4) || (src_C % 4) || (reinterpret_cast<uintptr_t>(src) & 15)) return DKT_E_ALIGNMENT;
    const int64_t total = (int64_t)B * H * W * 4 * (C / 4);
    pixel_shuffle2_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(src, src_C, group, *dst, C / 4, H, W, total);
    DKT_RETURN_LAST();
}

extern "C" int dkt_context_upsample_logits(const float* disp, const float* logits, int logit_C, float* out, float in_scale,
                                           float out_scale, int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(disp && logits && out);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && logit_C >= 36);
    const int64_t total = (int64_t)B * H * W * 16;
    context_upsample_logits_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
        disp, logits, logit_C, out, in_scale, out_scale, H, W, total);
    DKT_RETURN_LAST();
}

static int launch_to_nhwc(const float* src, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const float* bias,
                          const dkt_tensor& dst, int B, int C, int H, int W, void* stream) {
    if ((int64_t)B * H > 65535) return DKT_E_UNSUPPORTED;
    dim3 grid(ceil_div(W, 32), ceil_div(C, 32), B * H);
    to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, sb, sc, sh, sw, bias, dst, C, H, W);
    DKT_RETURN_LAST();
}

extern "C" int dkt_nchw_to_nhwc(const float* src, const float* bias, const dkt_tensor* dst, int B, int C, int H,
                                int W, void* stream) {
    DKT_CHECK_ARG(src);
    int rc = check_slice(dst, false);
    if (rc) return rc;
    DKT_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && dst->c_count == C);
    return launch_to_nhwc(src, (int64_t)C * H * W, (int64_t)H * W, W, 1, bias, *dst, B, C, H, W, stream);
}

extern "C" int dkt_nhwc_to_nchw(const dkt_tensor* src, float* dst, int B, int C, int H, int W, void* stream) {
    DKT_CHECK_ARG(dst);
    int rc = check_slice(src, true);
    if (rc) return rc;
    DKT_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && src->c_count == C);
    if ((int64_t)B * H > 65535) return DKT_E_UNSUPPORTED;
    dim3 grid(ceil_div(W, 32), ceil_div(C, 32), B * H);
    to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src->f32, src->C, src->c_begin, dst, C, H, W);
    DKT_RETURN_LAST();
}

extern "C" int dkt_split_nchw_to_nhwc_bf16x2(const float* src, int64_t sb, int64_t sd, int64_t sh, int64_t sw,
                                             uint16_t* hi, uint16_t* lo, int B, int D, int H, int W,
                                             void* stream) {
    DKT_CHECK_ARG(src && hi && lo);
    DKT_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    dkt_tensor dst{nullptr, hi, lo, D, 0, D};
    return launch_to_nhwc(src, sb, sd, sh, sw, nullptr, dst, B, D, H, W, stream);
}
