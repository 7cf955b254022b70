// IGEV-Stereo pre-loop volume kernels: the group-wise correlation volume, the 3-D convolutions on
// it (corr_stem + BatchNorm + LeakyReLU + feature attention; the 3-D hourglass; classifier) and the soft-argmin initial
// disparity.  Reference: meta_arch/igev_stereo/submodule.py:152-170 (build_gwc_volume), :10-36 (BasicConv), :227-240
// (FeatureAtt), :220-224 (disparity_regression); meta_arch/igev_stereo/igev_stereo.py:22-89 (hourglass), call sites
// igev_stereo.py:169-176.
//
// All of them are exact-fp32 SIMT kernels on NCDHW volumes (PyTorch modules produce the inputs and consume the outputs):
// the volumes are 401 MB at cfg3, an 8 -> 8 channel 3x3x3 stencil is 432 flop per voxel and input channel read once,
// i.e. FP32-FMA bound (43 GFLOP, 0.8 ms at the FMA peak), and channel counts of 8..48 are too small for a 128-wide MMA
// tile; the volume build and the soft-argmin are one HBM pass.  ncu launch tables: profiles/r03r_igev_preloop_launches.txt.
#include "common.cuh"

namespace dkt {

// 4-byte cp.async with zero fill (src-size 0): the staging loops below keep every load of a tile in flight at once
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// Group-wise correlation volume.  vol[b,g,d,y,x] = mean_{c in group g} L[b,c,y,x] * R[b,c,y,x-d]  (0 where x < d).
// One CTA = one image row segment of GW_X pixels, all groups and disparities.  L and the wider R segment are staged once
// in shared memory.  Thread = 4 consecutive x times 4 consecutive d: out(d0+j, x+i) needs R[x+i-d0-j], seven consecutive
// values that two aligned 16-byte loads deliver, so a channel costs 3 shared-memory loads for 16 FMAs (the first version,
// one (x, d) per thread, paid one load per FMA and ran at 0.55 ms for a 0.1 ms HBM pass).
// ------------------------------------------------------------------------------------------------------------------
constexpr int GW_X = 64;
constexpr int GW_MAX_DQ = 64;          // D <= 256

__global__ void __launch_bounds__(1024)
gwc_volume_kernel(const float* __restrict__ left, const float* __restrict__ right, float* __restrict__ vol,
                  int C, int G, int D, int H, int W, int OFF) {
    extern __shared__ __align__(16) float gw_smem[];
    const int RW = GW_X + OFF;                     // right segment: x0 - OFF .. x0 + GW_X - 1 (OFF = D rounded up to 4)
    float* sL = gw_smem;                           // [C][GW_X]
    float* sR = gw_smem + C * GW_X;                // [C][RW]
    const int x0 = blockIdx.x * GW_X, y = blockIdx.y, b = blockIdx.z;
    const int64_t plane = (int64_t)H * W;
    const float* Lb = left + (int64_t)b * C * plane + (int64_t)y * W;
    const float* Rb = right + (int64_t)b * C * plane + (int64_t)y * W;
    // asynchronous zero-filling copies: with plain loads the ~90 dependent load -> store rounds per thread were the whole
    // run time of this kernel (0.53 ms for a 0.1 ms HBM pass)
    for (int i = threadIdx.x; i < C * GW_X; i += blockDim.x) {
        const int c = i / GW_X, x = x0 + (i - c * GW_X);
        const bool ok = x < W;
        cp_async4(sL + i, ok ? Lb + c * plane + x : left, ok);
    }
    for (int i = threadIdx.x; i < C * RW; i += blockDim.x) {
        const int c = i / RW, x = x0 - OFF + (i - c * RW);
        const bool ok = x >= 0 && x < W;
        cp_async4(sR + i, ok ? Rb + c * plane + x : right, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int xq = threadIdx.x & 15, dq = threadIdx.x >> 4;
    const int x = x0 + 4 * xq, d0 = 4 * dq;
    if (x >= W || d0 >= D) return;
    const int cpg = C / G;
    const float inv = 1.0f / (float)cpg;
    const bool vec = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(vol) & 15) == 0);
    for (int g = 0; g < G; ++g) {
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
        const float* lp = sL + (g * cpg) * GW_X + 4 * xq;
        const float* rp = sR + (g * cpg) * RW + (4 * xq - d0 - 4 + OFF);      // 16-byte aligned: OFF, d0 multiples of 4
        for (int k = 0; k < cpg; ++k) {
            const float4 l4 = *reinterpret_cast<const float4*>(lp + k * GW_X);
            const float4 ra = *reinterpret_cast<const float4*>(rp + k * RW);
            const float4 rb = *reinterpret_cast<const float4*>(rp + k * RW + 4);
            const float l[4] = {l4.x, l4.y, l4.z, l4.w};
            const float r[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(l[i], r[4 + i - j], acc[j][i]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = d0 + j;
            if (d >= D) break;
            float* o = vol + (((int64_t)b * G + g) * D + d) * plane + (int64_t)y * W + x;
            if (vec) {
                *reinterpret_cast<float4*>(o) = make_float4(acc[j][0] * inv, acc[j][1] * inv, acc[j][2] * inv, acc[j][3] * inv);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (x + i < W) o[i] = acc[j][i] * inv;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3-D convolutions of the cost-volume stage (NCDHW fp32, exact fp32 FMA), all with the epilogue
//   v = acc * scale[co] + shift[co];  v = v > 0 ? v : slope * v;  v *= sigmoid(att[b,co,y,x])   (each optional)
// i.e. BasicConv's eval-mode BatchNorm3d folded + LeakyReLU(0.01) (submodule.py:10-36) and FeatureAtt's broadcast
// product (submodule.py:227-240).
//
// conv3d_k3_kernel<CO_T, DS, STRIDE>: 3x3x3, padding 1, stride 1 or 2.  CTA = (2*DS) x 4 x 32 outputs x CO_T channels,
// 256 threads; thread = one (y, x) column of DS consecutive d and CO_T channels (64 accumulators).  Input tile (with
// halo) and the weight slice are staged per chunk of input channels by 4-byte cp.async with zero fill (= the padding),
// double buffered so the next chunk lands while this one is consumed.  Per (ci, dy, dx) a thread reads DS*STRIDE+3-STRIDE
// inputs along d and 3 x CO_T broadcast weights for 3 x DS x CO_T FMAs (>= 10 FMAs per shared-memory instruction).
// ------------------------------------------------------------------------------------------------------------------

struct Conv3dEpi {
    const float* scale;
    const float* shift;
    const float* att;
    float slope;
};

__device__ __forceinline__ float conv3d_epi(float acc, float sc, float sh, float slope, float am) {
    float v = fmaf(acc, sc, sh);
    v = v > 0.f ? v : v * slope;
    return v * am;
}

constexpr int C3_XT = 32, C3_YT = 4, C3_NSTRIP = 2, C3_THREADS = C3_XT * C3_YT * C3_NSTRIP;

template <int CO_T, int DS, int STRIDE>
struct K3Cfg {
    static constexpr int DT = DS * C3_NSTRIP;
    static constexpr int CH = (STRIDE == 1) ? 2 : 1;                      // input channels per stage
    static constexpr int XIN = (C3_XT - 1) * STRIDE + 3, YIN = (C3_YT - 1) * STRIDE + 3, DIN = (DT - 1) * STRIDE + 3;
    static constexpr int TILE = (DIN * YIN * XIN + 3) & ~3;               // floats per input channel (16-byte multiple)
    static constexpr int WCH = 27 * CO_T;                                 // weights per input channel
    static constexpr int STAGE = CH * (TILE + WCH);
    static constexpr int NV = (DS - 1) * STRIDE + 3;                      // inputs along d per thread
    static constexpr bool TABLE = (STRIDE == 1);                          // per-element fill table (see prefetch)
    static constexpr size_t SMEM = (2 * (size_t)STAGE + (TABLE ? TILE : 0)) * sizeof(float);
};

template <int CO_T, int DS, int STRIDE>
__global__ void __launch_bounds__(C3_THREADS, 2)
conv3d_k3_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const Conv3dEpi epi,
                 float* __restrict__ out, int CI, int CO, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                 int dtiles, int coblocks) {
    using Cfg = K3Cfg<CO_T, DS, STRIDE>;
    extern __shared__ __align__(16) float c3_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int z = blockIdx.z;
    const int cb = z % coblocks; z /= coblocks;
    const int dt = z % dtiles;
    const int b = z / dtiles;
    const int x0 = blockIdx.x * C3_XT, y0 = blockIdx.y * C3_YT, d0 = dt * Cfg::DT, co0 = cb * CO_T;
    const int64_t iplane = (int64_t)Hi * Wi, oplane = (int64_t)Ho * Wo;
    const int xl = tid % C3_XT, yl = (tid / C3_XT) % C3_YT, ds = tid / (C3_XT * C3_YT);

    // fill table: element e of a channel's tile -> offset inside that channel's volume, -1 = padding.  The positions a
    // thread copies are the same for every input channel, so the index arithmetic (57 instructions per cp.async in
    // the first version, profiles/r03d) is done once per CTA instead of once per element and chunk.
    int* tab = reinterpret_cast<int*>(c3_smem + 2 * Cfg::STAGE);
    if (Cfg::TABLE) {
        for (int e = tid; e < Cfg::TILE; e += C3_THREADS) {
            const int dd = e / (Cfg::YIN * Cfg::XIN);
            const int rem = e - dd * (Cfg::YIN * Cfg::XIN);
            const int yy = rem / Cfg::XIN, xx = rem - yy * Cfg::XIN;
            const int d = d0 - 1 + dd, y = y0 - 1 + yy, x = x0 - 1 + xx;
            const bool ok = dd < Cfg::DIN && d >= 0 && d < Di && y >= 0 && y < Hi && x >= 0 && x < Wi;
            tab[e] = ok ? (int)(d * iplane + (int64_t)y * Wi + x) : -1;
        }
        __syncthreads();
    }

    auto prefetch = [&](int chunk, int buf) {
        float* sI = c3_smem + buf * Cfg::STAGE;
        float* sW = sI + Cfg::CH * Cfg::TILE;
        const int ci0 = chunk * Cfg::CH;
        if (Cfg::TABLE) {
            for (int cc = 0; cc < Cfg::CH; ++cc) {
                const float* base = in + ((int64_t)b * CI + ci0 + cc) * Di * iplane;
                float* dst = sI + cc * Cfg::TILE;
#pragma unroll 4
                for (int e = tid; e < Cfg::TILE; e += C3_THREADS) {
                    const int off = tab[e];
                    cp_async4(dst + e, base + max(off, 0), off >= 0);
                }
            }
        } else
        for (int r = warp; r < Cfg::CH * Cfg::DIN * Cfg::YIN; r += C3_THREADS / 32) {
            const int cc = r / (Cfg::DIN * Cfg::YIN);
            const int rem = r - cc * (Cfg::DIN * Cfg::YIN);
            const int dd = rem / Cfg::YIN, yy = rem - dd * Cfg::YIN;
            const int d = d0 * STRIDE - 1 + dd, y = y0 * STRIDE - 1 + yy;
            const bool rowok = d >= 0 && d < Di && y >= 0 && y < Hi;
            const float* src = rowok ? in + (((int64_t)b * CI + ci0 + cc) * Di + d) * iplane + (int64_t)y * Wi : in;
            float* dst = sI + cc * Cfg::TILE + (dd * Cfg::YIN + yy) * Cfg::XIN;
            for (int xx = lane; xx < Cfg::XIN; xx += 32) {
                const int x = x0 * STRIDE - 1 + xx;
                const bool ok = rowok && x >= 0 && x < Wi;
                cp_async4(dst + xx, ok ? src + x : in, ok);
            }
        }
        for (int i = tid; i < Cfg::CH * Cfg::WCH; i += C3_THREADS) {
            // destination i = ((cc*9 + dy*3 + dx)*3 + dz)*CO_T + co   <-   source [co][ci][dz][dy][dx]
            const int co = i % CO_T;
            int r = i / CO_T;
            const int dz = r % 3; r /= 3;
            const int dyx = r % 9;
            const int cc = r / 9;
            const bool ok = co0 + co < CO;
            cp_async4(sW + i, ok ? wgt + ((int64_t)(co0 + co) * CI + ci0 + cc) * 27 + dz * 9 + dyx : wgt, ok);
        }
        cp_async_commit();
    };

    float acc[DS][CO_T];
#pragma unroll
    for (int j = 0; j < DS; ++j)
#pragma unroll
        for (int co = 0; co < CO_T; ++co) acc[j][co] = 0.f;

    const int nchunks = CI / Cfg::CH;
    prefetch(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) {
            prefetch(chunk + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sI = c3_smem + buf * Cfg::STAGE;
        const float* sW = sI + Cfg::CH * Cfg::TILE;
#pragma unroll 1
        for (int cc = 0; cc < Cfg::CH; ++cc) {
#pragma unroll
            for (int dyx = 0; dyx < 9; ++dyx) {
                const int dy = dyx / 3, dx = dyx - dy * 3;
                const float* p = sI + cc * Cfg::TILE + ((ds * DS * STRIDE) * Cfg::YIN + (yl * STRIDE + dy)) * Cfg::XIN + (xl * STRIDE + dx);
                float v[Cfg::NV];
#pragma unroll
                for (int k = 0; k < Cfg::NV; ++k) v[k] = p[k * (Cfg::YIN * Cfg::XIN)];
                const float* wp = sW + (cc * 9 + dyx) * 3 * CO_T;
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    float w[CO_T];
                    if (CO_T % 4 == 0) {
#pragma unroll
                        for (int q = 0; q < CO_T / 4; ++q) {
                            const float4 t = *reinterpret_cast<const float4*>(wp + dz * CO_T + 4 * q);
                            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < CO_T; ++q) w[q] = wp[dz * CO_T + q];
                    }
#pragma unroll
                    for (int j = 0; j < DS; ++j)
#pragma unroll
                        for (int co = 0; co < CO_T; ++co) acc[j][co] = fmaf(v[j * STRIDE + dz], w[co], acc[j][co]);
                }
            }
        }
        __syncthreads();                                    // this buffer is refilled by the prefetch of chunk + 2
    }
    const int x = x0 + xl, y = y0 + yl;
    if (x >= Wo || y >= Ho) return;
#pragma unroll
    for (int co = 0; co < CO_T; ++co) {
        const int cg = co0 + co;
        if (cg >= CO) break;
        const float sc = epi.scale ? __ldg(epi.scale + cg) : 1.f, sh = epi.shift ? __ldg(epi.shift + cg) : 0.f;
        const float am = epi.att ? sigmoidf_acc(__ldg(epi.att + ((int64_t)b * CO + cg) * oplane + (int64_t)y * Wo + x)) : 1.f;
        float* o = out + ((int64_t)b * CO + cg) * Do * oplane + (int64_t)y * Wo + x;
#pragma unroll
        for (int j = 0; j < DS; ++j) {
            const int d = d0 + ds * DS + j;
            if (d >= Do) break;
            o[d * oplane] = conv3d_epi(acc[j][co], sc, sh, epi.slope, am);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// ConvTranspose3d, kernel 4, stride 2, padding 1 (output = 2 x input in every dimension): output o = 2m + p receives
// exactly two taps per dimension, (k = 1, i = m), (k = 3, i = m - 1) for p = 0 and (k = 2, i = m), (k = 0, i = m + 1)
// for p = 1.  Thread = DC_MD consecutive input cells along d at one (y, x): it keeps their 3 x 3 x (DC_MD + 2) input
// neighbourhood in registers and produces the 8 output parities of each cell for DC_CO channels (64 accumulators).
// Weight layout of the reference: [CI][CO][4][4][4].
// ------------------------------------------------------------------------------------------------------------------
constexpr int DC_MD = 2, DC_CO = 4, DC_CH = 4;
constexpr int DC_MDT = DC_MD * C3_NSTRIP;
constexpr int DC_XIN = C3_XT + 2, DC_YIN = C3_YT + 2, DC_DIN = DC_MDT + 2;
constexpr int DC_TILE = DC_DIN * DC_YIN * DC_XIN, DC_WCH = 64 * DC_CO, DC_STAGE = DC_CH * (DC_TILE + DC_WCH);
constexpr size_t DC_SMEM = (2 * (size_t)DC_STAGE + DC_TILE) * sizeof(float);      // two stages + the fill table

__global__ void __launch_bounds__(C3_THREADS, 2)
deconv3d_k4s2_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const Conv3dEpi epi,
                     float* __restrict__ out, int CI, int CO, int Di, int Hi, int Wi, int dtiles, int coblocks) {
    extern __shared__ __align__(16) float c3_smem[];
    const int tid = threadIdx.x;
    int z = blockIdx.z;
    const int cb = z % coblocks; z /= coblocks;
    const int dt = z % dtiles;
    const int b = z / dtiles;
    const int x0 = blockIdx.x * C3_XT, y0 = blockIdx.y * C3_YT, d0 = dt * DC_MDT, co0 = cb * DC_CO;
    const int64_t iplane = (int64_t)Hi * Wi;
    const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
    const int64_t oplane = (int64_t)Ho * Wo;
    const int xl = tid % C3_XT, yl = (tid / C3_XT) % C3_YT, ds = tid / (C3_XT * C3_YT);

    int* tab = reinterpret_cast<int*>(c3_smem + 2 * DC_STAGE);       // fill table, as in conv3d_k3_kernel
    for (int e = tid; e < DC_TILE; e += C3_THREADS) {
        const int dd = e / (DC_YIN * DC_XIN);
        const int rem = e - dd * (DC_YIN * DC_XIN);
        const int yy = rem / DC_XIN, xx = rem - yy * DC_XIN;
        const int d = d0 - 1 + dd, y = y0 - 1 + yy, x = x0 - 1 + xx;
        const bool ok = d >= 0 && d < Di && y >= 0 && y < Hi && x >= 0 && x < Wi;
        tab[e] = ok ? (int)(d * iplane + (int64_t)y * Wi + x) : -1;
    }
    __syncthreads();

    auto prefetch = [&](int chunk, int buf) {
        float* sI = c3_smem + buf * DC_STAGE;
        float* sW = sI + DC_CH * DC_TILE;
        const int ci0 = chunk * DC_CH;
        for (int cc = 0; cc < DC_CH; ++cc) {
            const float* base = in + ((int64_t)b * CI + ci0 + cc) * Di * iplane;
            float* dst = sI + cc * DC_TILE;
#pragma unroll 4
            for (int e = tid; e < DC_TILE; e += C3_THREADS) {
                const int off = tab[e];
                cp_async4(dst + e, base + max(off, 0), off >= 0);
            }
        }
        for (int i = tid; i < DC_CH * DC_WCH; i += C3_THREADS) {
            // destination i = (cc*64 + k)*DC_CO + co   <-   source [ci][co][k]
            const int co = i % DC_CO;
            const int r = i / DC_CO;
            const int k = r % 64, cc = r / 64;
            const bool ok = co0 + co < CO;
            cp_async4(sW + i, ok ? wgt + ((int64_t)(ci0 + cc) * CO + co0 + co) * 64 + k : wgt, ok);
        }
        cp_async_commit();
    };

    float acc[DC_MD][8][DC_CO];
#pragma unroll
    for (int j = 0; j < DC_MD; ++j)
#pragma unroll
        for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int co = 0; co < DC_CO; ++co) acc[j][p][co] = 0.f;

    const int nchunks = CI / DC_CH;
    prefetch(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) {
            prefetch(chunk + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sI = c3_smem + buf * DC_STAGE;
        const float* sW = sI + DC_CH * DC_TILE;
#pragma unroll 1
        for (int cc = 0; cc < DC_CH; ++cc) {
            float v[DC_MD + 2][3][3];
            const float* p0 = sI + cc * DC_TILE + ((ds * DC_MD) * DC_YIN + yl) * DC_XIN + xl;
#pragma unroll
            for (int dd = 0; dd < DC_MD + 2; ++dd)
#pragma unroll
                for (int yy = 0; yy < 3; ++yy)
#pragma unroll
                    for (int xx = 0; xx < 3; ++xx) v[dd][yy][xx] = p0[(dd * DC_YIN + yy) * DC_XIN + xx];
            const float4* w4 = reinterpret_cast<const float4*>(sW + cc * DC_WCH);
#pragma unroll
            for (int pd = 0; pd < 2; ++pd)
#pragma unroll
                for (int td = 0; td < 2; ++td) {
                    const int kd = pd == 0 ? (td == 0 ? 1 : 3) : (td == 0 ? 2 : 0);
                    const int od = td == 0 ? 0 : (pd == 0 ? -1 : 1);
#pragma unroll
                    for (int py = 0; py < 2; ++py)
#pragma unroll
                        for (int ty = 0; ty < 2; ++ty) {
                            const int kh = py == 0 ? (ty == 0 ? 1 : 3) : (ty == 0 ? 2 : 0);
                            const int oy = ty == 0 ? 0 : (py == 0 ? -1 : 1);
#pragma unroll
                            for (int px = 0; px < 2; ++px)
#pragma unroll
                                for (int tx = 0; tx < 2; ++tx) {
                                    const int kw = px == 0 ? (tx == 0 ? 1 : 3) : (tx == 0 ? 2 : 0);
                                    const int ox = tx == 0 ? 0 : (px == 0 ? -1 : 1);
                                    const float4 w = w4[(kd * 4 + kh) * 4 + kw];
                                    const int par = (pd * 2 + py) * 2 + px;
#pragma unroll
                                    for (int j = 0; j < DC_MD; ++j) {
                                        const float a = v[j + 1 + od][1 + oy][1 + ox];
                                        acc[j][par][0] = fmaf(a, w.x, acc[j][par][0]);
                                        acc[j][par][1] = fmaf(a, w.y, acc[j][par][1]);
                                        acc[j][par][2] = fmaf(a, w.z, acc[j][par][2]);
                                        acc[j][par][3] = fmaf(a, w.w, acc[j][par][3]);
                                    }
                                }
                        }
                }
        }
        __syncthreads();
    }
    const int mx = x0 + xl, my = y0 + yl;
    if (mx >= Wi || my >= Hi) return;
#pragma unroll
    for (int co = 0; co < DC_CO; ++co) {
        const int cg = co0 + co;
        if (cg >= CO) break;
        const float sc = epi.scale ? __ldg(epi.scale + cg) : 1.f, sh = epi.shift ? __ldg(epi.shift + cg) : 0.f;
        float* ob = out + ((int64_t)b * CO + cg) * Do * oplane;
#pragma unroll
        for (int j = 0; j < DC_MD; ++j) {
            const int md = d0 + ds * DC_MD + j;
            if (md >= Di) break;
#pragma unroll
            for (int pd = 0; pd < 2; ++pd)
#pragma unroll
                for (int py = 0; py < 2; ++py) {
                    const int par = (pd * 2 + py) * 2;
                    float2 r;
                    r.x = conv3d_epi(acc[j][par][co], sc, sh, epi.slope, 1.f);
                    r.y = conv3d_epi(acc[j][par + 1][co], sc, sh, epi.slope, 1.f);
                    *reinterpret_cast<float2*>(ob + (int64_t)(2 * md + pd) * oplane + (int64_t)(2 * my + py) * Wo + 2 * mx) = r;
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 1x1x1 convolution over the channel concatenation of two volumes (the first layer of the hourglass's agg blocks after
// torch.cat): thread = voxel, K1_CO output channels per thread, weights [CO][C0 + C1] staged transposed in shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int K1_CO = 16, K1_MAX_CI = 128;

__global__ void __launch_bounds__(256)
conv3d_k1_kernel(const float* __restrict__ in0, int C0, const float* __restrict__ in1, int C1,
                 const float* __restrict__ wgt, const Conv3dEpi epi, float* __restrict__ out, int CO,
                 int64_t vol, int64_t plane) {
    __shared__ __align__(16) float sw[K1_MAX_CI * K1_CO];
    const int CI = C0 + C1, co0 = blockIdx.y * K1_CO, b = blockIdx.z;
    for (int i = threadIdx.x; i < CI * K1_CO; i += blockDim.x) {
        const int co = i % K1_CO, ci = i / K1_CO;
        sw[i] = (co0 + co < CO) ? __ldg(wgt + (int64_t)(co0 + co) * CI + ci) : 0.f;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= vol) return;
    float acc[K1_CO];
#pragma unroll
    for (int co = 0; co < K1_CO; ++co) acc[co] = 0.f;
    const float* a = in0 + (int64_t)b * C0 * vol + i;
#pragma unroll 8
    for (int ci = 0; ci < C0; ++ci) {
        const float v = __ldg(a + ci * vol);
        const float4* w = reinterpret_cast<const float4*>(sw + ci * K1_CO);
#pragma unroll
        for (int q = 0; q < K1_CO / 4; ++q) {
            const float4 t = w[q];
            acc[4 * q] = fmaf(v, t.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, t.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, t.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, t.w, acc[4 * q + 3]);
        }
    }
    if (C1 > 0) {
        const float* c = in1 + (int64_t)b * C1 * vol + i;
#pragma unroll 8
        for (int ci = 0; ci < C1; ++ci) {
            const float v = __ldg(c + ci * vol);
            const float4* w = reinterpret_cast<const float4*>(sw + (C0 + ci) * K1_CO);
#pragma unroll
            for (int q = 0; q < K1_CO / 4; ++q) {
                const float4 t = w[q];
                acc[4 * q] = fmaf(v, t.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, t.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(v, t.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, t.w, acc[4 * q + 3]);
            }
        }
    }
    const int64_t pix = i % plane;
#pragma unroll
    for (int co = 0; co < K1_CO; ++co) {
        const int cg = co0 + co;
        if (cg >= CO) break;
        const float sc = epi.scale ? __ldg(epi.scale + cg) : 1.f, sh = epi.shift ? __ldg(epi.shift + cg) : 0.f;
        const float am = epi.att ? sigmoidf_acc(__ldg(epi.att + ((int64_t)b * CO + cg) * plane + pix)) : 1.f;
        out[((int64_t)b * CO + cg) * vol + i] = conv3d_epi(acc[co], sc, sh, epi.slope, am);
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
    const int dq = ceil_div(D, 4), off = 4 * dq;
    if (C / groups > 16 || dq > GW_MAX_DQ || H > 65535 || B > 65535) return DKT_E_UNSUPPORTED;
    const size_t smem = (size_t)C * (GW_X + GW_X + off) * sizeof(float);
    if (smem > 200 * 1024) return DKT_E_UNSUPPORTED;
    static size_t s_attr = 0;
    if (smem > s_attr) {
        cudaError_t ce = cudaFuncSetAttribute(gwc_volume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return (int)ce;
        s_attr = smem;
    }
    dim3 grid(ceil_div(W, GW_X), H, B);
    gwc_volume_kernel<<<grid, 16 * dq, smem, (cudaStream_t)stream>>>(left, right, vol, C, groups, D, H, W, off);
    DKT_RETURN_LAST();
}

template <int CO_T, int DS, int STRIDE>
static int launch_k3(const float* in, const float* weight, const Conv3dEpi& epi, float* out, int B, int CI, int CO,
                     int Di, int Hi, int Wi, int Do, int Ho, int Wo, cudaStream_t stream) {
    using Cfg = K3Cfg<CO_T, DS, STRIDE>;
    if (CI % Cfg::CH) return DKT_E_UNSUPPORTED;
    DKT_ENSURE_SMEM((int)Cfg::SMEM, conv3d_k3_kernel<CO_T, DS, STRIDE>);
    const int dtiles = ceil_div(Do, Cfg::DT), coblocks = ceil_div(CO, CO_T);
    if ((int64_t)B * dtiles * coblocks > 65535 || ceil_div(Ho, C3_YT) > 65535) return DKT_E_UNSUPPORTED;
    dim3 grid(ceil_div(Wo, C3_XT), ceil_div(Ho, C3_YT), B * dtiles * coblocks);
    conv3d_k3_kernel<CO_T, DS, STRIDE><<<grid, C3_THREADS, Cfg::SMEM, stream>>>(in, weight, epi, out, CI, CO, Di, Hi, Wi, Do, Ho, Wo,
                                                                             dtiles, coblocks);
    DKT_RETURN_LAST();
}

extern "C" int dkt_conv3d_k3(const float* in, const float* weight, const float* scale, const float* shift,
                             const float* att, float slope, float* out, int B, int CI, int CO, int D, int H, int W,
                             int stride, void* stream) {
    DKT_CHECK_ARG(in && weight && out && in != out);
    DKT_CHECK_ARG(B > 0 && CI > 0 && CO > 0 && D > 0 && H > 0 && W > 0);
    if (stride != 1 && stride != 2) return DKT_E_UNSUPPORTED;
    if ((int64_t)D * H * W > 0x7fffffff) return DKT_E_UNSUPPORTED;        // 32-bit offsets inside one channel volume (fill table)
    const int Do = (D - 1) / stride + 1, Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const Conv3dEpi epi{scale, shift, att, slope};
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1) {
        if (CO == 1) return launch_k3<1, 8, 1>(in, weight, epi, out, B, CI, CO, D, H, W, Do, Ho, Wo, st);
        if (CO % 16 == 0) return launch_k3<16, 4, 1>(in, weight, epi, out, B, CI, CO, D, H, W, Do, Ho, Wo, st);
        return launch_k3<8, 8, 1>(in, weight, epi, out, B, CI, CO, D, H, W, Do, Ho, Wo, st);
    }
    if (CO % 16 == 0) return launch_k3<16, 4, 2>(in, weight, epi, out, B, CI, CO, D, H, W, Do, Ho, Wo, st);
    return launch_k3<8, 8, 2>(in, weight, epi, out, B, CI, CO, D, H, W, Do, Ho, Wo, st);
}

extern "C" int dkt_conv3d_c8(const float* in, const float* weight, const float* scale, const float* shift,
                             const float* att, float slope, float* out, int B, int CO, int D, int H, int W,
                             void* stream) {
    if (CO != 8 && CO != 1) return DKT_E_UNSUPPORTED;
    return dkt_conv3d_k3(in, weight, scale, shift, att, slope, out, B, 8, CO, D, H, W, 1, stream);
}

extern "C" int dkt_deconv3d_k4s2(const float* in, const float* weight, const float* scale, const float* shift,
                                 float slope, float* out, int B, int CI, int CO, int D, int H, int W, void* stream) {
    DKT_CHECK_ARG(in && weight && out && in != out);
    DKT_CHECK_ARG(B > 0 && CI > 0 && CO > 0 && D > 0 && H > 0 && W > 0);
    if (CI % DC_CH) return DKT_E_UNSUPPORTED;
    if ((int64_t)D * H * W > 0x7fffffff / 8) return DKT_E_UNSUPPORTED;    // 32-bit offsets inside one channel volume (in and out)
    if (reinterpret_cast<uintptr_t>(out) & 7) return DKT_E_ALIGNMENT;
    DKT_ENSURE_SMEM((int)DC_SMEM, deconv3d_k4s2_kernel);
    const int dtiles = ceil_div(D, DC_MDT), coblocks = ceil_div(CO, DC_CO);
    if ((int64_t)B * dtiles * coblocks > 65535 || ceil_div(H, C3_YT) > 65535) return DKT_E_UNSUPPORTED;
    const Conv3dEpi epi{scale, shift, nullptr, slope};
    dim3 grid(ceil_div(W, C3_XT), ceil_div(H, C3_YT), B * dtiles * coblocks);
    deconv3d_k4s2_kernel<<<grid, C3_THREADS, DC_SMEM, (cudaStream_t)stream>>>(in, weight, epi, out, CI, CO, D, H, W, dtiles, coblocks);
    DKT_RETURN_LAST();
}

extern "C" int dkt_conv3d_k1(const float* in0, int C0, const float* in1, int C1, const float* weight,
                             const float* scale, const float* shift, const float* att, float slope, float* out,
                             int B, int CO, int D, int H, int W, void* stream) {
    DKT_CHECK_ARG(in0 && weight && out && C0 > 0 && C1 >= 0 && (C1 == 0 || in1));
    DKT_CHECK_ARG(B > 0 && CO > 0 && D > 0 && H > 0 && W > 0);
    if (C0 + C1 > K1_MAX_CI || B > 65535 || ceil_div(CO, K1_CO) > 65535) return DKT_E_UNSUPPORTED;
    const int64_t plane = (int64_t)H * W, vol = plane * D;
    const int64_t blocks = ceil_div64(vol, 256);
    if (blocks > 0x7fffffff) return DKT_E_UNSUPPORTED;
    const Conv3dEpi epi{scale, shift, att, slope};
    dim3 grid((unsigned)blocks, ceil_div(CO, K1_CO), B);
    conv3d_k1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in0, C0, in1, C1, weight, epi, out, CO, vol, plane);
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
