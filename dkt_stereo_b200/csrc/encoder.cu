// Encoder-side kernels (the feature / context extractors that feed the hot
// path; reference core/extractor.py:122-300).  The convolutions themselves run on conv_tc.cu;
// this file holds what surrounds them:
//
//   dkt_stem_rows_bf16x2   image normalisation 2*(x/255)-1 (reference raft_stereo.py:91-92) fused with
//                          an x-direction im2col of the 7x7 stem: out[b,y,x, kx*Cin + c] =
//                          img'[b,c,y,x+kx-3] (zero outside), padded to 64 channels and written as bf16
//                          (hi, lo).  The 7x7x3 conv then becomes a 7x1 conv over 64 channels, i.e. seven
//                          K=64 tensor-core taps instead of 147 scalar MACs per output on CUDA cores.
//   dkt_instnorm_stats     per (image, channel) mean / rstd of an NHWC fp32 tensor (nn.InstanceNorm2d,
//                          biased variance, eps 1e-5), two deterministic passes (no atomics).
//   dkt_instnorm_apply     y = (x - mean) * rstd, optional ReLU, optional residual relu(res + y)
//                          (ResidualBlock tail, reference core/extractor.py:56-60), written to every
//                          non-null precision of the destination slice.
#include "common.cuh"

namespace dkt {

// ---------------------------------------------------------------------------------------------
// stem rows
// ---------------------------------------------------------------------------------------------
constexpr int SR_PIX = 256;         // pixels of one image row per block (64 made 65 K tiny blocks per full-resolution launch)

__global__ void __launch_bounds__(256)
stem_rows_kernel(const float* __restrict__ img, int64_t sb, int64_t sc, int64_t sy, int64_t sx, float scale, float shift,
                 uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                 int Cin, int H, int W, int kw, int Cpad) {
    extern __shared__ float s_img[];                 // [Cin][SR_PIX + kw - 1]
    const int row = blockIdx.y;                      // b*H + y
    const int b = row / H, y = row - b * H;
    const int x0 = blockIdx.x * SR_PIX;
    const int halo = kw / 2, span = SR_PIX + kw - 1;
    for (int i = threadIdx.x; i < Cin * span; i += blockDim.x) {
        const int c = i / span, j = i - c * span;
        const int x = x0 + j - halo;
        float v = 0.f;
        if (x >= 0 && x < W) v = fmaf(img[b * sb + c * sc + y * sy + x * sx], scale, shift);
        s_img[i] = v;
    }
    __syncthreads();
    const int groups = Cpad >> 2;
    const int real = kw * Cin;
    for (int i = threadIdx.x; i < SR_PIX * groups; i += blockDim.x) {
        const int px = i / groups, g = i - px * groups;
        if (x0 + px >= W) continue;
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ch = g * 4 + u;
            if (ch < real) {
                const int kx = ch / Cin, c = ch - kx * Cin;
                v[u] = s_img[c * span + px + kx];
            } else {
                v[u] = 0.f;
            }
        }
        uint32_t h0, l0, h1, l1;
        split16x2(v[0], v[1], h0, l0);
        split16x2(v[2], v[3], h1, l1);
        const int64_t off = ((int64_t)row * W + x0 + px) * Cpad + g * 4;
        *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
    }
}

// ---------------------------------------------------------------------------------------------
// instance norm
// ---------------------------------------------------------------------------------------------
// pass 1: grid (chunks, B); partial[(b*chunks + chunk)*2*C + {0,1}*C + c] = sum, sum of squares
__global__ void __launch_bounds__(256)
instnorm_partial_kernel(const float* __restrict__ x, int xC, int c0, float* __restrict__ partial,
                        int HW, int C, int chunks) {
    extern __shared__ float s_red[];                 // [lanes][2][C]
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int groups = C >> 2;
    const int lanes = blockDim.x / groups;
    const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
    const int per = (HW + chunks - 1) / chunks;
    const int p_begin = chunk * per, p_end = min(HW, p_begin + per);
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (lane < lanes) {
        const float* base = x + (int64_t)b * HW * xC + c0 + g * 4;
        for (int p = p_begin + lane; p < p_end; p += lanes) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)p * xC));
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s_red[(lane * 2 + 0) * C + g * 4 + u] = s[u];
            s_red[(lane * 2 + 1) * C + g * 4 + u] = q[u];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float acc = 0.f;
        for (int l = 0; l < lanes; ++l) acc += s_red[l * 2 * C + i];        // fixed order: deterministic
        partial[((int64_t)b * chunks + chunk) * 2 * C + i] = acc;
    }
}

// pass 2: grid B, C threads: stats[(b*C + c)*2 + {0,1}] = mean, rstd
__global__ void instnorm_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats,
                                         int HW, int C, int chunks, float eps) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int k = 0; k < chunks; ++k) {
            s += (double)partial[((int64_t)b * chunks + k) * 2 * C + c];
            q += (double)partial[((int64_t)b * chunks + k) * 2 * C + C + c];
        }
        const double mean = s / HW;
        double var = q / HW - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((int64_t)b * C + c) * 2 + 0] = (float)mean;
        stats[((int64_t)b * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// ---------------------------------------------------------------------------------------------
// statistics from per-tile partial sums (written by the conv epilogue, dkt_epilogue.stats_partial):
// stage 1: grid (segments, B): segment sums over a contiguous range of the image's tiles, fp64;
// stage 2: grid B: the segments in order -> mean / rstd.  Every order is fixed: bit-reproducible and
// independent of the batch size and of the image's position in the batch.
// ---------------------------------------------------------------------------------------------
constexpr int IN_SEGS = 128;       // segments per image: stage 1 is a chain of dependent loads per thread, keep it short

__global__ void __launch_bounds__(256)
instnorm_tiles_stage1_kernel(const float* __restrict__ partial, double* __restrict__ seg, int tiles_per_img, int C2) {
    // C2 = 2*C values per tile; thread -> (slice, value): slices walk the segment's tiles interleaved
    extern __shared__ double s_acc[];                // [slices][C2]
    const int b = blockIdx.y, sgm = blockIdx.x;
    const int per = (tiles_per_img + IN_SEGS - 1) / IN_SEGS;
    const int t0 = sgm * per, t1 = min(tiles_per_img, t0 + per);
    const int slices = blockDim.x / C2;
    const int v = threadIdx.x % C2, sl = threadIdx.x / C2;
    if (sl < slices) {
        double acc = 0.0;
        const float* base = partial + ((int64_t)b * tiles_per_img) * C2 + v;
        int t = t0 + sl;
        for (; t + 3 * slices < t1; t += 4 * slices) {          // 4 independent loads in flight, summed in index order
            const float a0 = __ldg(base + (int64_t)t * C2), a1 = __ldg(base + (int64_t)(t + slices) * C2);
            const float a2 = __ldg(base + (int64_t)(t + 2 * slices) * C2), a3 = __ldg(base + (int64_t)(t + 3 * slices) * C2);
            acc += (double)a0; acc += (double)a1; acc += (double)a2; acc += (double)a3;
        }
        for (; t < t1; t += slices) acc += (double)__ldg(base + (int64_t)t * C2);
        s_acc[sl * C2 + v] = acc;
    }
    __syncthreads();
    if (threadIdx.x < C2) {
        double acc = 0.0;
        for (int k = 0; k < slices; ++k) acc += s_acc[k * C2 + threadIdx.x];
        seg[((int64_t)b * IN_SEGS + sgm) * C2 + threadIdx.x] = acc;
    }
}

// one warp per (image, channel): lane k adds segments k, k + 32, ... (independent loads), then a fixed-order shuffle tree --
// deterministic; one thread per channel walking all 128 segments took 28 us per launch, 15 launches per encoder pass
__global__ void __launch_bounds__(256)
instnorm_tiles_stage2_kernel(const double* __restrict__ seg, float* __restrict__ stats, int HW, int C, float eps) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int k = lane; k < IN_SEGS; k += 32) {
        s += seg[((int64_t)b * IN_SEGS + k) * 2 * C + c];
        q += seg[((int64_t)b * IN_SEGS + k) * 2 * C + C + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
        const double mean = s / HW;
        double var = q / HW - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((int64_t)b * C + c) * 2 + 0] = (float)mean;
        stats[((int64_t)b * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// grid: x = ceil(HW * C/4 / 256), y = B: 32-bit index math only, the image's stats address is block-uniform
__global__ void __launch_bounds__(256)
instnorm_apply_kernel(const float* __restrict__ x, int xC, int xc0, const float* __restrict__ stats,
                      const float* __restrict__ res, const uint16_t* __restrict__ res_hi,
                      const uint16_t* __restrict__ res_lo, int res_C, int res_c0, dkt_tensor out,
                      int relu, int HW, int C) {
    const int groups = C >> 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW * groups) return;
    const int pl = i / groups;
    const int c = (i - pl * groups) << 2;
    const int b = blockIdx.y;
    const int64_t p = (int64_t)b * HW + pl;
    const float4 v = *reinterpret_cast<const float4*>(x + p * xC + xc0 + c);
    const float4 st0 = __ldg(reinterpret_cast<const float4*>(stats + ((int64_t)b * C + c) * 2));       // m0 r0 m1 r1
    const float4 st1 = __ldg(reinterpret_cast<const float4*>(stats + ((int64_t)b * C + c) * 2 + 4));   // m2 r2 m3 r3
    float4 y = make_float4((v.x - st0.x) * st0.y, (v.y - st0.z) * st0.w, (v.z - st1.x) * st1.y, (v.w - st1.z) * st1.w);
    if (relu == 1) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
    else if (relu == 2) {                      // LeakyReLU(0.01): BasicConv_IN of the IGEV feature side
        y.x = y.x > 0.f ? y.x : 0.01f * y.x; y.y = y.y > 0.f ? y.y : 0.01f * y.y;
        y.z = y.z > 0.f ? y.z : 0.01f * y.z; y.w = y.w > 0.f ? y.w : 0.01f * y.w;
    }
    if (res || res_hi) {
        const float4 r = res ? *reinterpret_cast<const float4*>(res + p * res_C + res_c0 + c)
                             : load_split4(res_hi, res_lo, p * res_C + res_c0 + c);
        y.x = fmaxf(y.x + r.x, 0.f); y.y = fmaxf(y.y + r.y, 0.f); y.z = fmaxf(y.z + r.z, 0.f); y.w = fmaxf(y.w + r.w, 0.f);
    }
    store_all4(out, p, c, y);
}

// ---------------------------------------------------------------------------------------------
// tap sum: the second half of a 3x3 convolution with one output channel computed as
// (a) a 1x1 tensor-core conv to the 9 per-tap responses T[p][t] = sum_c w[t][c] x[p][c] and
// (b) out[p] = bias + sum_t T[p + (ky-1, kx-1)][t] (zero outside the image == zero padding).
// Used for the flow / disparity head's last conv (reference core/update.py:10,14: 256 -> 2, of which
// only channel 0 is ever used because raft_stereo.py:164 zeroes the y component), whose N = 2 would
// otherwise drag the 256-channel input through nine shifted tensor-core tiles.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tapsum_kernel(const float* __restrict__ T, int TC, float bias, float* __restrict__ out, int out_C,
              int64_t P, int H, int W) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    float acc = bias;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1;
            if (xx < 0 || xx >= W) continue;
            acc += __ldg(T + (p + (int64_t)(ky - 1) * W + (kx - 1)) * TC + ky * 3 + kx);
        }
    }
    out[p * out_C] = acc;
}

}  // namespace dkt

using namespace dkt;

extern "C" int dkt_tapsum3x3(const float* taps, int taps_C, float bias, float* out, int out_C,
                             int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(taps && out);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && taps_C >= 9 && out_C >= 1);
    const int64_t P = (int64_t)B * H * W;
    tapsum_kernel<<<(unsigned)ceil_div64(P, 256), 256, 0, (cudaStream_t)stream>>>(taps, taps_C, bias, out, out_C, P, H, W);
    DKT_RETURN_LAST();
}


extern "C" int dkt_stem_rows_bf16x2(const float* img, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                                    float scale, float shift, uint16_t* hi, uint16_t* lo,
                                    int B, int Cin, int H, int W, int kw, int Cpad, void* stream) {
    DKT_CHECK_ARG(img && hi && lo);
    DKT_CHECK_ARG(B > 0 && Cin > 0 && H > 0 && W > 0 && kw > 0 && (kw & 1));
    if (Cpad % 4 || kw * Cin > Cpad) return DKT_E_INVALID;
    if ((reinterpret_cast<uintptr_t>(hi) & 7) || (reinterpret_cast<uintptr_t>(lo) & 7)) return DKT_E_ALIGNMENT;
    if ((int64_t)B * H > 0x7fffffff) return DKT_E_UNSUPPORTED;
    dim3 grid(ceil_div(W, SR_PIX), B * H);
    if (grid.y > 65535) {
        // gridDim.y limit: one image (or row band) at a time
        for (int b = 0; b < B; ++b) {
            int rc = dkt_stem_rows_bf16x2(img + (int64_t)b * sb, sb, sc, sy, sx, scale, shift, hi + (int64_t)b * H * W * Cpad,
                                          lo + (int64_t)b * H * W * Cpad, 1, Cin, H, W, kw, Cpad, stream);
            if (rc) return rc;
        }
        return 0;
    }
    const size_t smem = (size_t)Cin * (SR_PIX + kw - 1) * sizeof(float);
    stem_rows_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(img, sb, sc, sy, sx, scale, shift, hi, lo, Cin, H, W, kw, Cpad);
    DKT_RETURN_LAST();
}

extern "C" int dkt_instnorm_workspace_floats(int B, int C) {
    return B * 128 * 2 * C;          // 128 pixel chunks per image
}

extern "C" int dkt_instnorm_stats(const dkt_tensor* x, float* workspace, float* stats, float eps,
                                  int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(x && x->f32 && workspace && stats);
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && x->c_count > 0);
    const int C = x->c_count;
    if ((C % 4) || (x->C % 4) || (x->c_begin % 4) || C > 1024) return DKT_E_ALIGNMENT;
    const int HW = H * W;
    const int chunks = HW < 128 ? 1 : 128;
    const int groups = C / 4;
    int threads = 256;
    if (groups > threads) return DKT_E_UNSUPPORTED;
    const int lanes = threads / groups;
    const size_t smem = (size_t)lanes * 2 * C * sizeof(float);
    if (smem > 48 * 1024) return DKT_E_UNSUPPORTED;
    instnorm_partial_kernel<<<dim3(chunks, B), threads, smem, (cudaStream_t)stream>>>(x->f32, x->C, x->c_begin, workspace, HW, C, chunks);
    instnorm_finalize_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(workspace, stats, HW, C, chunks, eps);
    DKT_RETURN_LAST();
}

extern "C" int dkt_instnorm_tiles_workspace_floats(int B, int C) {
    return B * IN_SEGS * 2 * C * 2;  // IN_SEGS segments of 2*C doubles per image
}

extern "C" int dkt_instnorm_finalize_tiles(const float* partial, float* workspace, float* stats, float eps,
                                           int B, int C, int H, int W, void* stream) {
    DKT_CHECK_ARG(partial && workspace && stats);
    DKT_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0);
    if (2 * C > 256 || (reinterpret_cast<uintptr_t>(workspace) & 7)) return DKT_E_UNSUPPORTED;
    const int tiles_per_img = ceil_div(W, 16) * ceil_div(H, 8) * 4;      // entries: one per (tile, 2-row quarter)
    const int C2 = 2 * C;
    const int slices = 256 / C2;
    double* seg = reinterpret_cast<double*>(workspace);
    instnorm_tiles_stage1_kernel<<<dim3(IN_SEGS, B), 256, (size_t)slices * C2 * sizeof(double), (cudaStream_t)stream>>>(
        partial, seg, tiles_per_img, C2);
    instnorm_tiles_stage2_kernel<<<dim3((unsigned)ceil_div(C, 8), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(seg, stats, H * W, C, eps);
    DKT_RETURN_LAST();
}

extern "C" int dkt_instnorm_apply(const dkt_tensor* x, const float* stats, const dkt_tensor* res, const dkt_tensor* out,
                                  int relu, int B, int H, int W, void* stream) {
    DKT_CHECK_ARG(x && x->f32 && stats && out && (out->f32 || out->hi));
    const int C = x->c_count;
    DKT_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && out->c_count == C);
    if ((C % 4) || (x->C % 4) || (x->c_begin % 4) || (out->C % 4) || (out->c_begin % 4)) return DKT_E_ALIGNMENT;
    if (res) {
        DKT_CHECK_ARG((res->f32 || (res->hi && res->lo)) && res->c_count == C);
        if ((res->C % 4) || (res->c_begin % 4)) return DKT_E_ALIGNMENT;
    }
    const int64_t n = (int64_t)H * W * (C / 4);
    if (n > 0x7fffffff || B > 65535) return DKT_E_UNSUPPORTED;
    instnorm_apply_kernel<<<dim3((unsigned)ceil_div64(n, 256), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        x->f32, x->C, x->c_begin, stats, res ? res->f32 : nullptr, res ? res->hi : nullptr, res ? res->lo : nullptr,
        res ? res->C : 0, res ? res->c_begin : 0, *out, relu, H * W, C);
    DKT_RETURN_LAST();
}
