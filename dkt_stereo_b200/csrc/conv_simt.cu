// K3 / CUDA-core variant: exact-fp32 implicit-GEMM convolution over NHWC activations with the
// fused epilogues of the ConvGRU update block.  This is the arbitrary-shape path (7x7 flow
// stem with 2 input channels, 2-channel flow head output, odd channel counts) and the on-device
// cross-check for the tcgen05 kernel in conv_tc.cu.  Contract: include/dkt_stereo_b200.h.
#include "common.cuh"

namespace dkt {

constexpr int SC_BM = 128, SC_BN = 64, SC_BK = 16;

struct SimtSrc {
    const float* f32;
    int C, c_begin, c_count;
};

struct SimtConvParams {
    SimtSrc src[DKT_MAX_SRCS];
    int nsrc, cin_total;
    const float* weight;      // [taps][cin_total][N]
    int ksize, pad, N;
    int H, W;
    int64_t P;
    int vec_ok;               // every slice 16-channel granular and 16-byte aligned
    dkt_epilogue epi;
};

__device__ __forceinline__ void epilogue_store(const dkt_epilogue& e, int N, int64_t p, int n, float acc) {
    if (e.kind == DKT_EPI_LINEAR) {
        float v = acc;
        if (e.bias) v += __ldg(e.bias + n);
        if (e.ctx) v += __ldg(e.ctx + p * e.ctx_C + e.ctx_c0 + n);
        v = apply_act(v, e.act) * e.scale;
        store_all(e.out, p, n, v);
    } else if (e.kind == DKT_EPI_GRU_ZR) {
        const int Nh = N >> 1;
        float v = acc + __ldg(e.ctx + p * e.ctx_C + e.ctx_c0 + n);
        float s = sigmoidf_acc(v);
        if (n < Nh) {
            e.z.f32[p * e.z.C + e.z.c_begin + n] = s;
        } else {
            int c = n - Nh;
            float hv = e.h.f32[p * e.h.C + e.h.c_begin + c];
            store_all(e.out, p, c, s * hv);
        }
    } else {  // DKT_EPI_GRU_Q
        float q = tanhf(acc + __ldg(e.ctx + p * e.ctx_C + e.ctx_c0 + n));
        float zv = e.z.f32[p * e.z.C + e.z.c_begin + n];
        float hv = e.h.f32[p * e.h.C + e.h.c_begin + n];
        store_all(e.out, p, n, (1.f - zv) * hv + zv * q);
    }
}

__global__ void __launch_bounds__(256)
conv2d_simt_kernel(const SimtConvParams prm) {
    __shared__ __align__(16) float As[SC_BK][SC_BM];
    __shared__ __align__(16) float Bs[SC_BK][SC_BN];

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * SC_BM;
    const int n0 = blockIdx.y * SC_BN;
    const int H = prm.H, W = prm.W, N = prm.N;
    const int taps = prm.ksize * prm.ksize;
    const int Ktot = taps * prm.cin_total;

    // the pixel this thread gathers for (fixed across the K loop)
    const int lm = tid % SC_BM;
    const int64_t lp = m0 + lm;
    const bool lp_ok = lp < prm.P;
    int lx = 0, ly = 0;
    int64_t lb = 0;
    if (lp_ok) {
        lx = (int)(lp % W);
        int64_t t = lp / W;
        ly = (int)(t % H);
        lb = t / H;
    }

    const int tx = tid % 16, ty = tid / 16;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Ktot; k0 += SC_BK) {
        if (prm.vec_ok) {
            // the 16-channel chunk lies inside one tap and one source
            const int tap = k0 / prm.cin_total;
            int c = k0 - tap * prm.cin_total;
            int s = 0;
            while (c >= prm.src[s].c_count) { c -= prm.src[s].c_count; ++s; }
            const SimtSrc& S = prm.src[s];
            const int ky = tap / prm.ksize, kx = tap - ky * prm.ksize;
            const int yy = ly + ky - prm.pad, xx = lx + kx - prm.pad;
            const bool ok = lp_ok && yy >= 0 && yy < H && xx >= 0 && xx < W;
            const float* base = S.f32 + ((lb * H + yy) * (int64_t)W + xx) * S.C + S.c_begin + c;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int q = tid / SC_BM + 2 * i;     // channel quad 0..3
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) v = __ldg(reinterpret_cast<const float4*>(base + q * 4));
                As[q * 4 + 0][lm] = v.x;
                As[q * 4 + 1][lm] = v.y;
                As[q * 4 + 2][lm] = v.z;
                As[q * 4 + 3][lm] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = tid / SC_BM + 2 * i;
                const int kg = k0 + k;
                float v = 0.f;
                if (kg < Ktot && lp_ok) {
                    const int tap = kg / prm.cin_total;
                    int c = kg - tap * prm.cin_total;
                    int s = 0;
                    while (c >= prm.src[s].c_count) { c -= prm.src[s].c_count; ++s; }
                    const SimtSrc& S = prm.src[s];
                    const int ky = tap / prm.ksize, kx = tap - ky * prm.ksize;
                    const int yy = ly + ky - prm.pad, xx = lx + kx - prm.pad;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                        v = __ldg(S.f32 + ((lb * H + yy) * (int64_t)W + xx) * S.C + S.c_begin + c);
                }
                As[k][lm] = v;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = tid % SC_BN, k = tid / SC_BN + 4 * i;
            const int kg = k0 + k;
            float v = 0.f;
            if (kg < Ktot && n0 + n < N) v = __ldg(prm.weight + (int64_t)kg * N + n0 + n);
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SC_BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t p = m0 + ty * 8 + i;
        if (p >= prm.P) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) epilogue_store(prm.epi, N, p, n, acc[i][j]);
        }
        if (prm.epi.tail && n0 == 0 && tx == 0) {
            for (int t = 0; t < prm.epi.tail_C; ++t)
                store_all(prm.epi.out, p, N + t, __ldg(prm.epi.tail + p * prm.epi.tail_C + t));
        }
    }
}

}  // namespace dkt

using namespace dkt;

static int validate_epilogue(const dkt_epilogue* e, int N) {
    if (!e) return DKT_E_INVALID;
    switch (e->kind) {
        case DKT_EPI_LINEAR:
            if (!e->out.f32 && !e->out.hi) return DKT_E_INVALID;
            if (e->tail && e->tail_C <= 0) return DKT_E_INVALID;
            break;
        case DKT_EPI_GRU_ZR:
            if ((N & 1) || !e->ctx || !e->z.f32 || !e->h.f32 || (!e->out.f32 && !e->out.hi)) return DKT_E_INVALID;
            break;
        case DKT_EPI_GRU_Q:
            if (!e->ctx || !e->z.f32 || !e->h.f32 || (!e->out.f32 && !e->out.hi)) return DKT_E_INVALID;
            break;
        default:
            return DKT_E_INVALID;
    }
    return 0;
}

extern "C" int dkt_conv2d_simt(const dkt_tensor* srcs, int nsrc, const float* weight, int ksize, int N,
                               const dkt_epilogue* epi, int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(srcs && weight && epi);
    DKT_CHECK_ARG(nsrc >= 1 && nsrc <= DKT_MAX_SRCS);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && N > 0);
    if (ksize < 1 || ksize > 7 || !(ksize & 1)) return DKT_E_UNSUPPORTED;
    int rc = validate_epilogue(epi, N);
    if (rc) return rc;
    SimtConvParams prm{};
    prm.nsrc = nsrc;
    prm.vec_ok = 1;
    for (int s = 0; s < nsrc; ++s) {
        DKT_CHECK_ARG(srcs[s].f32 && srcs[s].c_count > 0 && srcs[s].c_begin >= 0 &&
                      srcs[s].c_begin + srcs[s].c_count <= srcs[s].C);
        prm.src[s] = SimtSrc{srcs[s].f32, srcs[s].C, srcs[s].c_begin, srcs[s].c_count};
        prm.cin_total += srcs[s].c_count;
        if ((srcs[s].c_count % 16) || (srcs[s].c_begin % 4) || (srcs[s].C % 4) ||
            (reinterpret_cast<uintptr_t>(srcs[s].f32) & 15))
            prm.vec_ok = 0;
    }
    prm.weight = weight;
    prm.ksize = ksize;
    prm.pad = ksize / 2;
    prm.N = N;
    prm.H = H;
    prm.W = W;
    prm.P = (int64_t)B * H * W;
    prm.epi = *epi;
    dim3 grid((unsigned)ceil_div64(prm.P, SC_BM), ceil_div(N, SC_BN));
    conv2d_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(prm);
    DKT_RETURN_LAST();
}
