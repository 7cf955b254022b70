// IGEV-Stereo pre-loop volume kernels: the group-wise correlation volume, the 8-channel 3x3x3
// 3-D convolutions on it (corr_stem + BatchNorm + LeakyReLU + feature attention; classifier) and the soft-argmin
// initial disparity.  Reference: meta_arch/igev_stereo/submodule.py:152-170 (build_gwc_volume), :10-36 (BasicConv),
// :227-240 (FeatureAtt), :220-224 (disparity_regression); call sites meta_arch/igev_stereo/igev_stereo.py:169-176.
//
// All three are exact-fp32 SIMT kernels: the volumes are 401 MB at cfg3, the arithmetic intensity of an 8 -> 8 channel
// 3x3x3 stencil is 432 flop per voxel-channel pair read once, i.e. FP32-FMA bound (43 GFLOP, 0.8 ms at the FMA peak)
// with channel counts far too small for a 128-wide MMA tile; the volume build and the soft-argmin are one HBM pass.
#include "common.cuh"

namespace dkt {

// ------------------------------------------------------------------------------------------------------------------
// Group-wise correlation volume.  vol[b,g,d,y,x] = mean_{c in group g} L[b,c,y,x] * R[b,c,y,x-d]  (0 where x < d).
// One CTA = one image row segment of GW_X pixels, all groups and disparities.  L and the D-1 pixel wider R segment are
// staged once in shared memory (every R value is used by up to D outputs); thread = (x, d mod GW_DQ).
// ------------------------------------------------------------------------------------------------------------------
constexpr int GW_X = 64;
constexpr int GW_DQ = 4;
constexpr int GW_MAX_CPG = 16;

__global__ void __launch_bounds__(GW_X * GW_DQ)
gwc_volume_kernel(const float* __restrict__ left, const float* __restrict__ right, float* __restrict__ vol,
                  int C, int G, int D, int H, int W) {
    extern __shared__ float gw_smem[];
    const int RW = GW_X + D - 1;                   // right segment: x0 - (D-1) .. x0 + GW_X - 1
    float* sL = gw_smem;                           // [C][GW_X]
    float* sR = gw_smem + C * GW_X;                // [C][RW]
    const int x0 = blockIdx.x * GW_X, y = blockIdx.y, b = blockIdx.z;
    const int64_t plane = (int64_t)H * W;
    const float* Lb = left + (int64_t)b * C * plane + (int64_t)y * W;
    const float* Rb = right + (int64_t)b * C * plane + (int64_t)y * W;
    for (int i = threadIdx.x; i < C * GW_X; i += blockDim.x) {
        const int c = i / GW_X, x = x0 + (i - c * GW_X);
        sL[i] = (x < W) ? __ldg(Lb + c * plane + x) : 0.f;
    }
    for (int i = threadIdx.x; i < C * RW; i += blockDim.x) {
        const int c = i / RW, x = x0 - (D - 1) + (i - c * RW);
        sR[i] = (x >= 0 && x < W) ? __ldg(Rb + c * plane + x) : 0.f;
    }
    __syncthreads();
    const int xl = threadIdx.x % GW_X, dq = threadIdx.x / GW_X;
    const int x = x0 + xl;
    if (x >= W) return;
    const int cpg = C / G;
    const float inv = 1.0f / (float)cpg;
    for (int g = 0; g < G; ++g) {
        float lv[GW_MAX_CPG];
#pragma unroll
        for (int k = 0; k < GW_MAX_CPG; ++k) lv[k] = (k < cpg) ? sL[(g * cpg + k) * GW_X + xl] : 0.f;
        for (int d = dq; d < D; d += GW_DQ) {
            const float* r = sR + (g * cpg) * RW + (xl + (D - 1) - d);
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < GW_MAX_CPG; ++k)
                if (k < cpg) acc = fmaf(lv[k], r[k * RW], acc);
            vol[(((int64_t)b * G + g) * D + d) * plane + (int64_t)y * W + x] = acc * inv;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3x3x3 convolution, 8 input channels, CO in {8, 1} output channels, stride 1, zero padding 1, no bias; epilogue
//   v = acc * scale[co] + shift[co];  v = v > 0 ? v : slope * v;  v *= sigmoid(att[b,co,y,x])   (each optional).
// Tile: 16 (d) x 4 (y) x 32 (x) outputs per CTA, 256 threads; thread = one (y, x) column of 8 consecutive d and all CO
// channels (64 accumulators for CO = 8).  The input tile with its halo is staged four channels at a time; per
// (ci, dy, dx) a thread reads 10 inputs along d and the 3 x CO weights (broadcast float4 loads) for 24 x CO FMAs:
// 12 FMAs per shared-memory instruction, i.e. FMA-issue bound.
// ------------------------------------------------------------------------------------------------------------------
constexpr int C3_DT = 16, C3_YT = 4, C3_XT = 32, C3_DS = 8, C3_CI = 8, C3_CHUNK = 4;
constexpr int C3_ROW = C3_XT + 2;                                   // 34 floats per tile row
constexpr int C3_TILE = (C3_DT + 2) * (C3_YT + 2) * C3_ROW;          // floats per channel
constexpr int C3_THREADS = C3_XT * C3_YT * (C3_DT / C3_DS);

template <int CO>
__global__ void __launch_bounds__(C3_THREADS, 2)
conv3d_c8_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ att, float slope,
                 float* __restrict__ out, int D, int H, int W, int dtiles) {
    extern __shared__ float c3_smem[];
    float* s_in = c3_smem;                                  // [C3_CHUNK][DT+2][YT+2][ROW]
    float* s_w = c3_smem + C3_CHUNK * C3_TILE;              // [ci][dy][dx][dz][CO]
    const int x0 = blockIdx.x * C3_XT, y0 = blockIdx.y * C3_YT;
    const int b = blockIdx.z / dtiles, d0 = (blockIdx.z - b * dtiles) * C3_DT;
    const int64_t plane = (int64_t)H * W, cvol = (int64_t)D * plane;
    const int tid = threadIdx.x;
    for (int i = tid; i < C3_CI * 27 * CO; i += C3_THREADS) {
        // destination index i = (((ci*3 + dy)*3 + dx)*3 + dz)*CO + co  <-  source [co][ci][dz][dy][dx]
        const int co = i % CO;
        int r = i / CO;
        const int dz = r % 3; r /= 3;
        const int dx = r % 3; r /= 3;
        const int dy = r % 3;
        const int ci = r / 3;
        s_w[i] = __ldg(wgt + ((co * C3_CI + ci) * 27 + dz * 9 + dy * 3 + dx));
    }
    const int xl = tid % C3_XT, yl = (tid / C3_XT) % C3_YT, ds = tid / (C3_XT * C3_YT);
    float acc[C3_DS][CO];
#pragma unroll
    for (int j = 0; j < C3_DS; ++j)
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[j][co] = 0.f;

    for (int chunk = 0; chunk < C3_CI / C3_CHUNK; ++chunk) {
        __syncthreads();                                    // previous chunk consumed (and s_w visible)
        const float* src = in + ((int64_t)b * C3_CI + chunk * C3_CHUNK) * cvol;
        for (int i = tid; i < C3_CHUNK * C3_TILE; i += C3_THREADS) {
            const int cc = i / C3_TILE;
            int r = i - cc * C3_TILE;
            const int dd = r / ((C3_YT + 2) * C3_ROW);
            r -= dd * ((C3_YT + 2) * C3_ROW);
            const int yy = r / C3_ROW, xx = r - yy * C3_ROW;
            const int d = d0 - 1 + dd, y = y0 - 1 + yy, x = x0 - 1 + xx;
            float v = 0.f;
            if (d >= 0 && d < D && y >= 0 && y < H && x >= 0 && x < W) v = __ldg(src + cc * cvol + d * plane + (int64_t)y * W + x);
            s_in[i] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int cc = 0; cc < C3_CHUNK; ++cc) {
            const float* wci = s_w + (chunk * C3_CHUNK + cc) * 27 * CO;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float* p = s_in + cc * C3_TILE + (ds * C3_DS) * ((C3_YT + 2) * C3_ROW) + (yl + dy) * C3_ROW + (xl + dx);
                    float v[C3_DS + 2];
#pragma unroll
                    for (int k = 0; k < C3_DS + 2; ++k) v[k] = p[k * ((C3_YT + 2) * C3_ROW)];
                    float w[3][CO];
                    const float* wp = wci + (dy * 3 + dx) * 3 * CO;
                    if (CO % 4 == 0) {
#pragma unroll
                        for (int q = 0; q < 3 * CO / 4; ++q) {
                            const float4 t = *reinterpret_cast<const float4*>(wp + 4 * q);
                            (&w[0][0])[4 * q] = t.x; (&w[0][0])[4 * q + 1] = t.y; (&w[0][0])[4 * q + 2] = t.z; (&w[0][0])[4 * q + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 3 * CO; ++q) (&w[0][0])[q] = wp[q];
                    }
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                        for (int j = 0; j < C3_DS; ++j)
#pragma unroll
                            for (int co = 0; co < CO; ++co) acc[j][co] = fmaf(v[j + dz], w[dz][co], acc[j][co]);
                }
            }
        }
    }
    const int x = x0 + xl, y = y0 + yl;
    if (x >= W || y >= H) return;
#pragma unroll
    for (int co = 0; co < CO; ++co) {
        const float sc = scale ? __ldg(scale + co) : 1.f, sh = shift ? __ldg(shift + co) : 0.f;
        const float am = att ? sigmoidf_acc(__ldg(att + ((int64_t)b * CO + co) * plane + (int64_t)y * W + x)) : 1.f;
        float* o = out + ((int64_t)b * CO + co) * cvol + (int64_t)y * W + x;
#pragma unroll
        for (int j = 0; j < C3_DS; ++j) {
            const int d = d0 + ds * C3_DS + j;
            if (d >= D) break;
            float v = fmaf(acc[j][co], sc, sh);
            v = v > 0.f ? v : v * slope;
            o[d * plane] = v * am;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// soft-argmin over the disparity axis: disp[b,p] = sum_d d * softmax_d(logits[b,:,p]).  Thread = pixel; the D values of
// neighbouring pixels are neighbouring addresses, so each of the two passes is coalesced (the second one hits L2).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softargmin_kernel(const float* __restrict__ logits, float* __restrict__ disp, int D, int64_t HW, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / HW, p = i - b * HW;
    const float* l = logits + b * D * HW + p;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(l + d * HW));
    float s = 0.f, sd = 0.f;
    for (int d = 0; d < D; ++d) {
        const float e = expf(__ldg(l + d * HW) - m);
        s += e;
        sd = fmaf((float)d, e, sd);
    }
    disp[i] = sd / s;
}

}  // namespace dkt

using namespace dkt;

extern "C" int dkt_gwc_volume(const float* left, const float* right, float* vol, int B, int C, int groups, int D,
                              int H, int W, void* stream) {
    DKT_CHECK_ARG(left && right && vol);
    DKT_CHECK_ARG(B > 0 && C > 0 && groups > 0 && D > 0 && H > 0 && W > 0 && (C % groups) == 0);
    if (C / groups > GW_MAX_CPG || H > 65535 || B > 65535) return DKT_E_UNSUPPORTED;
    const size_t smem = (size_t)C * (GW_X + GW_X + D - 1) * sizeof(float);
    if (smem > 200 * 1024) return DKT_E_UNSUPPORTED;
    static size_t s_attr = 0;
    if (smem > s_attr) {
        cudaError_t ce = cudaFuncSetAttribute(gwc_volume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return (int)ce;
        s_attr = smem;
    }
    dim3 grid(ceil_div(W, GW_X), H, B);
    gwc_volume_kernel<<<grid, GW_X * GW_DQ, smem, (cudaStream_t)stream>>>(left, right, vol, C, groups, D, H, W);
    DKT_RETURN_LAST();
}

extern "C" int dkt_conv3d_c8(const float* in, const float* weight, const float* scale, const float* shift,
                             const float* att, float slope, float* out, int B, int CO, int D, int H, int W,
                             void* stream) {
    DKT_CHECK_ARG(in && weight && out && in != out);
    DKT_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    if (CO != 8 && CO != 1) return DKT_E_UNSUPPORTED;
    const int dtiles = ceil_div(D, C3_DT);
    if ((int64_t)B * dtiles > 65535 || ceil_div(H, C3_YT) > 65535) return DKT_E_UNSUPPORTED;
    const size_t smem = (size_t)(C3_CHUNK * C3_TILE + C3_CI * 27 * CO) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv3d_c8_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv3d_c8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (ce != cudaSuccess) return (int)ce;
        attr_set = true;
    }
    dim3 grid(ceil_div(W, C3_XT), ceil_div(H, C3_YT), B * dtiles);
    if (CO == 8)
        conv3d_c8_kernel<8><<<grid, C3_THREADS, smem, (cudaStream_t)stream>>>(in, weight, scale, shift, att, slope, out, D, H, W, dtiles);
    else
        conv3d_c8_kernel<1><<<grid, C3_THREADS, smem, (cudaStream_t)stream>>>(in, weight, scale, shift, att, slope, out, D, H, W, dtiles);
    DKT_RETURN_LAST();
}

extern "C" int dkt_softargmin(const float* logits, float* disp, int B, int D, int H, int W, void* stream) {
    DKT_CHECK_ARG(logits && disp);
    DKT_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0);
    const int64_t HW = (int64_t)H * W, total = HW * B;
    const int64_t blocks = ceil_div64(total, 256);
    if (blocks > 0x7fffffff) return DKT_E_UNSUPPORTED;
    softargmin_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(logits, disp, D, HW, total);
    DKT_RETURN_LAST();
}
