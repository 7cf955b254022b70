/*
 * dkt_stereo_b200.h -- C ABI of the B200-native stereo hot path (libdkt_stereo_b200.so).
 *
 * This is the drop-in boundary for the ONE path this repository accelerates: the
 * correlation-volume build + indexed lookup and the ConvGRU disparity-update loop of
 * RAFT-Stereo / IGEV-Stereo as shipped in jiaw-z/DKT-Stereo.  The reference has no native
 * source of its own; the closest thing it has to an operator ABI is the optional, un-vendored
 * `corr_sampler` extension (reference core/corr.py:17-29,49).  Each entry point below names the
 * reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers + sizes, no framework types.  The caller owns every
 *     buffer (inputs, outputs, workspace); the library allocates nothing persistent and keeps
 *     no global state, so calls are thread-safe per stream and CUDA-graph capturable.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - return value: 0 = ok, < 0 = invalid argument (DKT_E_*), > 0 = a cudaError_t.
 *   - "NHWC" buffers are (B, H, W, C) fp32 / 16-bit with C contiguous.  Activations and weights
 *     on the tensor-core path are carried as a 16-bit (hi, lo) pair, hi = rn16(x), lo = rn16(x - hi),
 *     in the format dkt_split_format() reports: IEEE half (default build; hi alone = 11 significant
 *     bits, hi + lo = 22) or bfloat16 (-DDKT_SPLIT_FP16=0; 8 / 16 bits).  Products are accumulated
 *     in fp32 as  hi*hi + lo*hi + hi*lo  (3 MMAs per K step; every tensor with both planes) or, for a
 *     conv whose SOURCE tensors are passed with lo == NULL, as  hi*w_hi + hi*w_lo  (2 MMAs: half
 *     activations against full-precision weights); w_lo == NULL additionally drops the w_lo product
 *     (1 MMA: plain half x half).  Which convs may run with 2 is a property of the
 *     network, measured in profiles/r2_precision_study_*.txt; the engine in dkt_stereo_b200/update.py
 *     makes that choice, this library only executes it.
 *   - 16-bit values are passed as uint16_t ("bf16" in entry-point names is historical: it means
 *     "the library's 16-bit split format").
 */
#ifndef DKT_STEREO_B200_H
#define DKT_STEREO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DKT_ABI_VERSION 3   /* 2: dkt_epilogue grew (proj, res_hi/res_lo, stats_partial)
                               3: dkt_split_format(); conv sources / destinations may omit the lo plane */

/* 16-bit operand formats (dkt_split_format) */
#define DKT_FMT_BF16 0
#define DKT_FMT_FP16 1

/* negative return codes */
#define DKT_E_INVALID     (-1)  /* null pointer / bad dimension */
#define DKT_E_UNSUPPORTED (-2)  /* shape outside what the kernels were built for */
#define DKT_E_ALIGNMENT   (-3)  /* pointer or channel count not aligned as required */
#define DKT_E_DRIVER      (-4)  /* cuTensorMapEncodeTiled unavailable / failed */

#define DKT_MAX_LEVELS 4
#define DKT_MAX_SRCS   3

/* activation codes */
#define DKT_ACT_NONE    0
#define DKT_ACT_RELU    1
#define DKT_ACT_SIGMOID 2
#define DKT_ACT_TANH    3
#define DKT_ACT_LEAKY   4  /* LeakyReLU, negative slope 0.01 (nn.LeakyReLU() of the reference's BasicConv, igev submodule.py:29) */

/* epilogue kinds of dkt_conv2d_* */
#define DKT_EPI_LINEAR 0  /* y = scale * act(acc + bias[n] + ctx[p][n])                          */
#define DKT_EPI_GRU_ZR 1  /* n <  N/2: z = sigmoid(acc + ctx) -> z ; n >= N/2: r = sigmoid(..),
                             (r*h) -> out           (reference core/update.py:27-28,29 first half) */
#define DKT_EPI_GRU_Q  2  /* q = tanh(acc + ctx); h' = (1-z)*h + z*q -> out (h may alias out)
                             (reference core/update.py:29-31)                                     */
#define DKT_EPI_PROJ   3  /* tensor-core path only: y = act(acc + bias[n]) never reaches HBM; instead
                             out.f32[p][t] = sum_n y[n] * proj[n][t], t < DKT_PROJ_LD -- the channel half of
                             the conv that FOLLOWS this one, applied while the tile is still in TMEM
                             (FlowHead / DispHead: conv1+ReLU, then the 9 per-tap responses of conv2's
                             used output channel; reference core/update.py:9-14)                  */
#define DKT_PROJ_LD 12    /* row length of dkt_epilogue.proj (floats; 16-byte aligned rows)        */

/* One NHWC activation tensor, optionally in several precisions.  Null members are skipped
 * by writers; readers document which member they need. */
typedef struct dkt_tensor {
    float*    f32;      /* fp32 values                                   */
    uint16_t* hi;       /* 16-bit high part (tensor-core path)           */
    uint16_t* lo;       /* 16-bit low part; NULL = hi-only tensor        */
    int32_t   C;        /* channels per pixel in the buffer (the stride) */
    int32_t   c_begin;  /* first channel of the slice used               */
    int32_t   c_count;  /* number of channels in the slice               */
} dkt_tensor;

/* Fused epilogue description for the conv kernels. */
typedef struct dkt_epilogue {
    int32_t      kind;      /* DKT_EPI_*                                                       */
    int32_t      act;       /* DKT_ACT_* (LINEAR only)                                         */
    float        scale;     /* output multiplier applied after the activation (LINEAR only)    */
    const float* bias;      /* [N] or NULL                                                     */
    const float* ctx;       /* per-pixel additive term, NHWC fp32, or NULL                     */
    int32_t      ctx_C;     /* channel stride of ctx                                           */
    int32_t      ctx_c0;    /* channel of ctx that lines up with output channel 0              */
    dkt_tensor   out;       /* destination slice (LINEAR: N channels; GRU_ZR: r*h, N/2; GRU_Q: h', N) */
    dkt_tensor   z;         /* GRU_ZR: written (f32 only, N/2 ch).  GRU_Q: read                */
    dkt_tensor   h;         /* GRU_ZR / GRU_Q: hidden state read (f32)                         */
    const float* tail;      /* LINEAR: optional NHWC fp32 source copied into channels          */
    int32_t      tail_C;    /*   [N_valid, N_valid+tail_C) of `out` (motion features ++ flow,  */
                            /*   reference core/update.py:85).                                 */
    const float* res;       /* LINEAR (tensor-core path): optional NHWC fp32 residual; the     */
    int32_t      res_C;     /*   result becomes relu(y + res[p][res_c0 + n]) -- the tail of a  */
    int32_t      res_c0;    /*   ResidualBlock (reference core/extractor.py:56-60).            */
    const float* proj;      /* PROJ: fp32 [N][DKT_PROJ_LD] projection matrix (unused columns 0)    */
    const uint16_t* res_hi; /* LINEAR (tensor-core path): the residual as a bf16 (hi, lo) pair, used when    */
    const uint16_t* res_lo; /*   res == NULL (same res_C / res_c0): residual = hi + lo, so the block input  */
                            /*   needs no separate fp32 copy in HBM.                                        */
    float*       stats_partial; /* LINEAR (tensor-core path, N % 32 == 0): when non-NULL the kernel also  */
                            /*   writes, per 8 x 16 output tile t (t = (b*tiles_y + ty)*tiles_x + tx) and  */
                            /*   per 2-row quarter q of the tile, [t][q][0][n] = sum and [t][q][1][n] =    */
                            /*   sum of squares of the stored values over the quarter's valid pixels       */
                            /*   (fixed summation order): the statistics pass of nn.InstanceNorm2d fused   */
                            /*   into the conv that feeds it, finished by dkt_instnorm_finalize_tiles      */
                            /*   (reference core/extractor.py:16-33).  Size: tiles * 8 * N floats.         */
} dkt_epilogue;

/* ---- library ---------------------------------------------------------------------------- */
int         dkt_abi_version(void);
const char* dkt_error_string(int code);
/* DKT_FMT_FP16 or DKT_FMT_BF16: the 16-bit format of every hi / lo plane this build reads and writes. */
int         dkt_split_format(void);
/* 1 if the running device is sm_100 (tcgen05/TMA paths usable), 0 otherwise, <0 on error. */
int         dkt_device_supported(int device);

/* ---- K1: all-pairs 1-D correlation volume + W2 pyramid -----------------------------------
 * Replaces CorrBlock1D.corr + the avg_pool2d pyramid (reference core/corr.py:111-125,148-156)
 * and, with scale = 1, Combined_Geo_Encoding_Volume.corr + its init_corr pyramid (reference
 * meta_arch/igev_stereo/geometry.py:14-29,61-69).
 *   fmap1/fmap2 : fp32, element (b,d,y,x) at b*sb + d*sd + y*sh + x*sw (elements); any of
 *                 NCHW or channels_last is expressible.
 *   pyr[l]      : fp32 (B,H,W1,W2>>l) contiguous, l < levels <= DKT_MAX_LEVELS; level l is the
 *                 pairwise average of level l-1 (an odd tail element is dropped).
 *   scale       : multiplier applied to the dot product (1/sqrt(D) for RAFT-Stereo). */
int dkt_corr1d_build_f32(const float* fmap1, const float* fmap2,
                         int64_t sb, int64_t sd, int64_t sh, int64_t sw,
                         float* const* pyr, int B, int D, int H, int W1, int W2,
                         int levels, float scale, void* stream);

/* Tensor-core (tcgen05, 3-term bf16 split, TMA-staged) variant of the above.  Inputs are the
 * feature maps already split to NHWC bf16 (hi, lo) pairs, (B,H,W,D) with D % 64 == 0 -- use
 * dkt_split_nchw_to_nhwc_bf16x2 to produce them from the extractor's fp32 output. */
int dkt_corr1d_build_tc(const uint16_t* f1_hi, const uint16_t* f1_lo,
                        const uint16_t* f2_hi, const uint16_t* f2_lo,
                        float* const* pyr, int B, int D, int H, int W1, int W2,
                        int levels, float scale, void* stream);

/* ---- K2: indexed multi-level lookup, fused with the coordinate update ---------------------
 * Replaces CorrBlock1D.__call__ + bilinear_sampler (reference core/corr.py:127-146,
 * core/utils/utils.py:59-74) and the corr_sampler.forward plugin (core/corr.py:22), plus the
 * loop bookkeeping `delta_flow[:,1]=0; coords1 += delta_flow; flow = coords1 - coords0`
 * (reference meta_arch/raft_stereo/raft_stereo.py:154-155,164-167).
 *   coords_x : (B,H,W) fp32 current x coordinate; updated in place when delta != NULL
 *   delta    : NULL, or the NHWC (B,H,W,delta_C) output of the flow head; channel 0 is added
 *   flow     : NULL, or NHWC (B,H,W,2) fp32: receives (coords_x - x, flow_y unchanged = kept)
 *   out      : level-major taps, channel = l*(2r+1)+k, element (b,c,p) at b*ob + c*oc + p*op
 *              (NCHW: ob=C*H*W, oc=H*W, op=1; NHWC with padding: ob=H*W*Cp, oc=1, op=Cp; the
 *              NHWC form takes the coalesced path and also zero-fills channels [C, Cp)).
 *              out == NULL: only the coordinate / flow bookkeeping runs.
 *   out_hi/lo: optional bf16 split of the same values with the same strides (may be NULL).
 * dkt_corr1d_lookup_enc additionally applies the motion encoder's first layer to the taps while
 * they are still on chip: enc_out[p][n] = relu(enc_b[n] + sum_c taps[p][c] * enc_w[c][n]), n < 64
 * (reference core/update.py:72,79 `convc1`, exact fp32), enc_w fp32 [C][64]; enc_out is a
 * 64-channel NHWC slice (every non-null precision is written).  The taps never reach HBM. */
int dkt_corr1d_lookup(const float* const* pyr, int levels, int radius,
                      float* coords_x, const float* delta, int delta_C, float* flow,
                      float* out, uint16_t* out_hi, uint16_t* out_lo,
                      int64_t ob, int64_t oc, int64_t op,
                      int B, int H, int W1, int W2, void* stream);
int dkt_corr1d_lookup_enc(const float* const* pyr, int levels, int radius,
                          float* coords_x, const float* delta, int delta_C, float* flow,
                          const float* enc_w, const float* enc_b, const dkt_tensor* enc_out,
                          int B, int H, int W1, int W2, void* stream);

/* Tensor-core form of dkt_corr1d_lookup_enc (tcgen05, csrc/lookup_tc.cu): same lookup, same coordinate bookkeeping,
 * `convc1` + ReLU as a 16-bit split contraction with fp32 accumulation instead of fp32 FMAs.
 *   w_img      : convc1 as the kernel's shared-memory image, [2 planes (hi, lo)][KB][64 rows x 64 k] 16-bit values in
 *                K-major SWIZZLE_128B order, KB = ceil(C / 64) (dkt_stereo_b200.ops.pack_lookup_tc writes it)
 *   tap_planes : 2 = taps as (hi, lo) pairs, 3 MMAs per K step; 1 = hi only, 2 MMAs (see the header comment)
 *   enc_out    : 64-channel NHWC slice, C and c_begin multiples of 8, planes 16-byte aligned */
int dkt_corr1d_lookup_enc_tc(const float* const* pyr, int levels, int radius,
                             float* coords_x, const float* delta, int delta_C, float* flow,
                             const uint16_t* w_img, const float* enc_b, const dkt_tensor* enc_out, int tap_planes,
                             int B, int H, int W1, int W2, void* stream);

/* Adjoint of the ONE-level lookup with respect to the volume: replaces corr_sampler.backward(volume, coords,
 * grad_output, radius) -> (grad_volume,) of the un-vendored extension the reference binds in CorrSampler.backward
 * (reference core/corr.py:25-29; used only when corr_implementation = "reg_cuda" is trained).
 *   grad_out    : (B, 2r+1, H, W1) fp32 contiguous -- gradient of the (B, 2r+1, H, W1) tap tensor
 *   coords_x    : (B, H, W1) fp32 -- the x coordinates of the forward call (already divided by 2^level)
 *   grad_volume : (B, H, W1, W2) fp32 contiguous -- written in full (zeros where no tap landed)
 * No gradient flows to the coordinates (the reference returns None for them too). */
int dkt_corr1d_lookup_backward(const float* grad_out, const float* coords_x, int radius, float* grad_volume,
                               int B, int H, int W1, int W2, void* stream);

/* ---- IGEV: geometry-encoding-volume pyramid + combined lookup ------------------------------
 * dkt_geo_pool: (B,C,D,H,W) fp32 -> level 0 (B,H,W,C,D) and level 1 (B,H,W,C,D/2)
 *   (reference meta_arch/igev_stereo/geometry.py:17-26; levels == 2 as in configs/igev_stereo).
 * dkt_geo_lookup: reference geometry.py:34-58 -> per level [C*(2r+1) geo taps, (2r+1) init taps];
 *   disp (B,H,W) fp32 is updated in place by delta[...,0] first when delta != NULL
 *   (`disp = disp + delta_disp`, reference meta_arch/igev_stereo/igev_stereo.py:210).
 * dkt_geo_lookup_enc: same + the IGEV motion encoder's convc1 (igev update.py:76,85), see above. */
int dkt_geo_pool(const float* gev, float* geo0, float* geo1, int B, int C, int D, int H, int W, void* stream);
/* The same pyramid in the layout the tensor-core lookup reads: level 0 (B,H,W,D,C), level 1 (B,H,W,D/2,C) -- the
 * 2r+2 disparity samples x C channels one lookup needs from a (pixel, level) are ONE contiguous run. */
int dkt_geo_pool_dc(const float* gev, float* geo0, float* geo1, int B, int C, int D, int H, int W, void* stream);
int dkt_geo_lookup(const float* geo0, const float* geo1, const float* init0, const float* init1,
                   float* disp, const float* delta, int delta_C, int radius, int C, int D,
                   float* out, uint16_t* out_hi, uint16_t* out_lo,
                   int64_t ob, int64_t oc, int64_t op,
                   int B, int H, int W, void* stream);
int dkt_geo_lookup_enc(const float* geo0, const float* geo1, const float* init0, const float* init1,
                       float* disp, const float* delta, int delta_C, int radius, int C, int D,
                       const float* enc_w, const float* enc_b, const dkt_tensor* enc_out,
                       int B, int H, int W, void* stream);
/* Tensor-core form of dkt_geo_lookup_enc (see dkt_corr1d_lookup_enc_tc); geo0 / geo1 in the dkt_geo_pool_dc layout.
 * With tap_planes == 1 and 16-byte aligned volumes whose lengths are multiples of 4 floats (always true for the model's
 * buffers) the geometry runs and the init-corr windows are fetched by the TMA unit (cp.async.bulk into a shared-memory
 * ring, csrc/lookup_tc.cu geo_lookup_tma_kernel); otherwise, or with DKT_LOOKUP_TMA=0, by per-thread loads.  Same results. */
int dkt_geo_lookup_enc_tc(const float* geo0, const float* geo1, const float* init0, const float* init1,
                          float* disp, const float* delta, int delta_C, int radius, int C, int D,
                          const uint16_t* w_img, const float* enc_b, const dkt_tensor* enc_out, int tap_planes,
                          int B, int H, int W, void* stream);

/* ---- IGEV pre-loop volume kernels, exact fp32 ------------------------------
 * dkt_gwc_volume: group-wise correlation volume, replaces build_gwc_volume + groupwise_correlation
 *   (reference meta_arch/igev_stereo/submodule.py:152-170; call site igev_stereo.py:169):
 *   left, right (B,C,H,W) -> vol (B,groups,D,H,W), vol[b,g,d,y,x] = mean over the group's C/groups channels of
 *   left[b,c,y,x] * right[b,c,y,x-d], 0 where x < d.  C/groups <= 16.
 * dkt_conv3d_k3: 3x3x3 Conv3d (padding 1, stride 1 or 2, no bias) with the epilogue
 *   v = acc * scale[co] + shift[co]; v = v > 0 ? v : slope * v; v *= sigmoid(att[b,co,y,x])  (scale/shift/att may be
 *   NULL; slope = 1 for no activation).  in (B,CI,D,H,W), weight [CO][CI][3][3][3] (PyTorch layout), out
 *   (B,CO,Do,Ho,Wo) with Xo = (X-1)/stride + 1, att (B,CO,Ho,Wo), in != out; CI even.  Replaces BasicConv(is_3d) with
 *   its eval-mode BatchNorm3d folded to scale/shift + LeakyReLU(0.01) (submodule.py:10-36) followed by FeatureAtt's
 *   broadcast product (submodule.py:227-240): corr_stem + corr_feature_att (igev_stereo.py:130-131,170-171), the
 *   3x3x3 layers of `hourglass` (igev_stereo.py:22-89) and `classifier` = nn.Conv3d(8,1,3,1,1) (igev_stereo.py:133,175).
 * dkt_conv3d_c8: dkt_conv3d_k3 with CI = 8, stride 1, CO = 8 or 1.
 * dkt_deconv3d_k4s2: ConvTranspose3d(kernel 4, stride 2, padding 1, no bias) + scale/shift + leaky: in (B,CI,D,H,W),
 *   weight [CI][CO][4][4][4], out (B,CO,2D,2H,2W); CI % 4 == 0.  hourglass conv3_up / conv2_up / conv1_up
 *   (igev_stereo.py:42-49).
 * dkt_conv3d_k1: 1x1x1 Conv3d over the channel concatenation [in0 (C0) | in1 (C1)] (in1 may be NULL with C1 = 0),
 *   weight [CO][C0+C1], same epilogue as dkt_conv3d_k3: torch.cat + the first layer of hourglass agg_0 / agg_1
 *   (igev_stereo.py:51-58,76-77,81-82); C0 + C1 <= 128.
 * dkt_softargmin: logits (B,D,H,W) -> disp (B,1,H,W) = sum_d d * softmax_d(logits): F.softmax(dim=1) followed by
 *   disparity_regression (submodule.py:220-224; igev_stereo.py:175-176). */
int dkt_gwc_volume(const float* left, const float* right, float* vol, int B, int C, int groups, int D, int H, int W,
                   void* stream);
int dkt_conv3d_k3(const float* in, const float* weight, const float* scale, const float* shift, const float* att,
                  float slope, float* out, int B, int CI, int CO, int D, int H, int W, int stride, void* stream);
int dkt_conv3d_c8(const float* in, const float* weight, const float* scale, const float* shift, const float* att,
                  float slope, float* out, int B, int CO, int D, int H, int W, void* stream);
int dkt_deconv3d_k4s2(const float* in, const float* weight, const float* scale, const float* shift, float slope,
                      float* out, int B, int CI, int CO, int D, int H, int W, void* stream);
int dkt_conv3d_k1(const float* in0, int C0, const float* in1, int C1, const float* weight, const float* scale,
                  const float* shift, const float* att, float slope, float* out, int B, int CO, int D, int H, int W,
                  void* stream);
int dkt_softargmin(const float* logits, float* disp, int B, int D, int H, int W, void* stream);
/* The hourglass's stride-1 3x3x3 layers with 16 / 32 / 48 channels run on the 2-D tensor-core conv (dkt_conv2d_tc_ex):
 * the depth planes of a depth-padded NDHWC copy of the volume are its images and the three kz taps three channel-
 * concatenated sources (the same buffer at plane offsets 0, 1, 2).  These two passes are the way in and out:
 * dkt_ncdhw_to_ndhwc_pad : src (B,C,D,H,W) fp32 -> 16-bit (hi, lo) planes (B,D+2,H,W,C), interior depth planes 1..D; planes
 *   0 and D+1 of every sample are the zero padding in depth (the caller zeroes them once; never written here).
 * dkt_ndhwc_pad_to_ncdhw : src fp32 (B,D+2,H,W,C), interior planes -> dst (B,C,D,H,W) fp32, multiplied by
 *   sigmoid(att[b,c,y,x]) when att != NULL (FeatureAtt, reference submodule.py:227-240).  B*D*H <= 65535. */
/* dkt_dwconv3x3 : depthwise 3x3 convolution (padding 1, stride 1 or 2) over an NHWC fp32 slice, for the MobileNetV2 encoder
 *   of IGEV's feature pyramid (timm conv_dw + BatchNorm + ReLU6, reference meta_arch/igev_stereo/extractor.py:331-342):
 *   dst[p][c] = clamp(bias[c] + sum_t min(src[p + t][c], in_max) * weight[t][c], out_min, out_max); weight fp32 [9][C] with
 *   the eval-mode BatchNorm folded in; every non-null precision of dst (B,H,W,.) is written, H = (Hin-1)/stride + 1.
 *   in_max completes the ReLU6 of the expand conv that produced src (its epilogue applied the ReLU half). */
int dkt_dwconv3x3(const dkt_tensor* src, const float* weight, const float* bias, float in_max, float out_min, float out_max,
                  const dkt_tensor* dst, int B, int Hin, int Win, int stride, void* stream);
int dkt_ncdhw_to_ndhwc_pad(const float* src, uint16_t* hi, uint16_t* lo, int B, int C, int D, int H, int W, void* stream);
int dkt_ndhwc_pad_to_ncdhw(const float* src, const float* att, float* dst, int B, int C, int D, int H, int W, void* stream);

/* ---- K3: convolutions of the update block with fused epilogues ------------------------------
 * Replaces nn.Conv2d + bias + activation + the GRU gate algebra of reference core/update.py
 * (FlowHead :6-14, ConvGRU :16-32, BasicMotionEncoder :64-85, mask head :110-113) and the IGEV
 * twins in meta_arch/igev_stereo/update.py.
 * Input  = channel-concatenation of `nsrc` NHWC slices (all B x H x W), stride 1, pad ksize/2.
 * Weight = fp32 [ksize*ksize][Cin_total][N] (SIMT) or bf16 hi/lo [ksize*ksize][Npad][Cin_total]
 *          (tensor core; K contiguous, Npad = N rounded up to 16).
 *   dkt_conv2d_simt : exact fp32 CUDA-core implicit GEMM (any Cin / N); reads src.f32.
 *   dkt_conv2d_tc   : tcgen05 implicit GEMM, TMA im2col-free tap loads with zero-filled halos,
 *                     3-term bf16 split; reads src.hi/lo; needs c_begin % 64 == 0 and c_count % 16 == 0 (a source's
 *                     last 64-channel K block may be partial, e.g. the encoders' 96-channel stages; or one
 *                     32-channel source: the 7x7 stems' x-im2col rows), ksize in {1,3}, N <= 256 (any multiple of 16
 *                     is a native UMMA width; other N are zero-padded to the next one).
 *                     Stride-1 convs with >= 2 output tiles run as CTA pairs (tcgen05 cta_group::2). */
int dkt_conv2d_simt(const dkt_tensor* srcs, int nsrc, const float* weight, int ksize, int N,
                    const dkt_epilogue* epi, int B, int H, int W, void* stream);
int dkt_conv2d_tc(const dkt_tensor* srcs, int nsrc, const uint16_t* w_hi, const uint16_t* w_lo,
                  int ksize, int N, const dkt_epilogue* epi, int B, int H, int W, void* stream);
/* General form used by the encoders (reference core/extractor.py): kh x kw filter (odd, <= 7, padding
 * k/2), stride 1 or 2; sources are (B,Hin,Win,C), outputs (B,H,W,.) with H = (Hin + 2*(kh/2) - kh)/stride + 1.
 * Weight layout as above with taps = kh*kw, tap = ky*kw + kx. */
int dkt_conv2d_tc_ex(const dkt_tensor* srcs, int nsrc, const uint16_t* w_hi, const uint16_t* w_lo,
                     int kh, int kw, int stride, int N, const dkt_epilogue* epi,
                     int B, int Hin, int Win, int H, int W, void* stream);

/* ---- a8: cross-scale plumbing of the multi-level GRU -----------------------------------------
 * dkt_pool2x  : F.avg_pool2d(x,3,stride=2,padding=1), divisor 9 (reference core/update.py:87-88)
 * dkt_interp  : F.interpolate(bilinear, align_corners=True) (reference core/update.py:93-95)
 * Both read src.f32 (B,Hs,Ws) and write every non-null member of dst (B,Hd,Wd). */
int dkt_pool2x(const dkt_tensor* src, const dkt_tensor* dst, int B, int Hs, int Ws, int Hd, int Wd, void* stream);
int dkt_interp(const dkt_tensor* src, const dkt_tensor* dst, int B, int Hs, int Ws, int Hd, int Wd, void* stream);

/* ---- K4: final upsampling ---------------------------------------------------------------------
 * dkt_convex_upsample : RAFTStereo.upsample_flow (reference meta_arch/raft_stereo/raft_stereo.py:70-82)
 *   flow  NHWC (B,H,W,flow_C) fp32, channel 0 is upsampled; mask NHWC (B,H,W,9*f*f) fp32;
 *   out (B,1,f*H,f*W) fp32.
 * dkt_context_upsample : IGEV context_upsample (reference meta_arch/igev_stereo/submodule.py:242-254)
 *   disp (B,H,W) fp32, weights (B,9,4H,4W) fp32 NCHW -> out (B,4H,4W), out = out_scale * sum. */
int dkt_convex_upsample(const float* flow, int flow_C, const float* mask, float* out,
                        int B, int H, int W, int factor, void* stream);
int dkt_context_upsample(const float* disp, const float* weights, float* out, float in_scale, float out_scale,
                         int B, int H, int W, void* stream);
/* IGEV upsample_disp on the library's kernels (reference meta_arch/igev_stereo/igev_stereo.py:140-148): the two
 * ConvTranspose2d(kernel 4, stride 2, padding 1) run as 3x3 convolutions that produce the four output parities as
 * channel groups (dkt_conv2d_tc with weights from dkt_stereo_b200.ops.deconv4x4s2_as_conv3x3), then
 * dkt_pixel_shuffle2 : src fp32 NHWC (B,H,W,src_C), channel (py*2+px)*group + c, c < C  ->  dst (B,2H,2W) slice of C
 *                      channels, dst[b][2y+py][2x+px][c] (every non-null plane of dst is written);
 * dkt_context_upsample_logits : softmax over the 9 logits + context_upsample in one pass.  logits fp32 NHWC
 *                      (B,2H,2W,logit_C) with channel (py*2+px)*9 + k = logit k of full-resolution pixel
 *                      (2*(2y')+py.., see above), disp (B,H,W) fp32  ->  out (B,4H,4W) = out_scale * sum_k softmax_k *
 *                      in_scale * disp[neighbour k] (reference submodule.py:242-254 after F.softmax, :145-146). */
int dkt_pixel_shuffle2(const float* src, int src_C, int group, const dkt_tensor* dst, int B, int H, int W, void* stream);
int dkt_context_upsample_logits(const float* disp, const float* logits, int logit_C, float* out, float in_scale,
                                float out_scale, int B, int H, int W, void* stream);

/* ---- layout / precision plumbing ---------------------------------------------------------------
 * dkt_nchw_to_nhwc : fp32 (B,C,H,W) -> every non-null member of dst (NHWC slice), y = x + bias[c].
 * dkt_nhwc_to_nchw : NHWC slice (f32) -> fp32 (B,C,H,W).
 * dkt_split_nchw_to_nhwc_bf16x2 : strided fp32 feature map -> NHWC bf16 (hi, lo) for dkt_corr1d_build_tc. */
int dkt_nchw_to_nhwc(const float* src, const float* bias, const dkt_tensor* dst, int B, int C, int H, int W, void* stream);
int dkt_nhwc_to_nchw(const dkt_tensor* src, float* dst, int B, int C, int H, int W, void* stream);
int dkt_split_nchw_to_nhwc_bf16x2(const float* src, int64_t sb, int64_t sd, int64_t sh, int64_t sw,
                                  uint16_t* hi, uint16_t* lo, int B, int D, int H, int W, void* stream);

/* ---- encoder side (reference core/extractor.py:122-300 on the same kernels) ----
 * dkt_stem_rows_bf16x2 : image fp32 (element (b,c,y,x) at b*sb + c*sc + y*sy + x*sx) -> NHWC (B,H,W,Cpad)
 *   bf16 (hi,lo) with channel kx*Cin + c =
 *   scale*img[b,c,y,x+kx-kw/2] + shift (0 outside the image; channels >= kw*Cin are 0).  With scale = 2/255,
 *   shift = -1 this is the input normalisation of raft_stereo.py:91-92 fused with an x-im2col, after which the
 *   7x7 stem conv (core/extractor.py:140) is a dkt_conv2d_tc_ex with kh = 7, kw = 1 over Cpad channels.
 *   The motion encoder's 7x7 flow stem (core/update.py:73,81) uses the same pair on the NHWC flow field.
 * dkt_instnorm_stats   : nn.InstanceNorm2d statistics (biased variance) of an NHWC fp32 slice ->
 *   stats (B,C,2) = (mean, 1/sqrt(var+eps)); workspace >= dkt_instnorm_workspace_floats(B,C) floats.
 * dkt_instnorm_apply   : out = (x - mean)*rstd, then ReLU if relu == 1 / LeakyReLU(0.01) if relu == 2 (BasicConv_IN, igev
 *   submodule.py:80-106), then relu(res + out) if res != NULL
 *   (the ResidualBlock tail, core/extractor.py:56-60; res is read as fp32 if res->f32 != NULL, else as hi + lo);
 *   writes every non-null precision of `out`. */
int dkt_stem_rows_bf16x2(const float* img, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                         float scale, float shift, uint16_t* hi, uint16_t* lo,
                         int B, int Cin, int H, int W, int kw, int Cpad, void* stream);
/* dkt_tapsum3x3 : out[p*out_C] = bias + sum_{ky,kx} taps[(p + (ky-1)*W + (kx-1))*taps_C + ky*3+kx] with zero
 *   outside the image: the spatial half of a one-output-channel 3x3 conv whose channel half (a 1x1 conv to
 *   9 tap responses) ran on dkt_conv2d_tc.  Used for FlowHead.conv2 / DispHead.conv2 (reference
 *   core/update.py:10,14; only output channel 0 is consumed, raft_stereo.py:164). */
int dkt_tapsum3x3(const float* taps, int taps_C, float bias, float* out, int out_C,
                  int B, int H, int W, void* stream);
int dkt_instnorm_workspace_floats(int B, int C);
int dkt_instnorm_stats(const dkt_tensor* x, float* workspace, float* stats, float eps,
                       int B, int H, int W, void* stream);
int dkt_instnorm_apply(const dkt_tensor* x, const float* stats, const dkt_tensor* res, const dkt_tensor* out,
                       int relu, int B, int H, int W, void* stream);
/* dkt_instnorm_finalize_tiles : per-tile partial sums written by a conv with dkt_epilogue.stats_partial
 *   (B * tiles_per_img tiles of 8 x 16 pixels, [tile][4][2][C]) -> stats (B,C,2) = (mean, 1/sqrt(var+eps)) over the H*W
 *   pixels of each image; tiles are reduced in a fixed order, in double.  workspace >=
 *   dkt_instnorm_tiles_workspace_floats(B, C) floats. */
int dkt_instnorm_tiles_workspace_floats(int B, int C);
int dkt_instnorm_finalize_tiles(const float* partial, float* workspace, float* stats, float eps,
                                int B, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DKT_STEREO_B200_H */
