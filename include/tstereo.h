/*
 * tstereo.h — C ABI of libtstereo.so, the B200 (sm_100a) cost-volume stereo engine.
 *
 * Drop-in boundary for the per-frame hot path of youmi-zym/TemporalStereo.
 * Every entry point takes plain device pointers + sizes + a cudaStream_t (as void*),
 * launches asynchronously on that stream and returns 0 or a negative TSTEREO_E_* code;
 * tstereo_last_error() returns the message of the calling thread's last failure.
 * No torch types cross this boundary.  All tensors are fp32, contiguous NCHW / NCDHW
 * unless an entry point takes explicit element strides.
 *
 * "ref:" lines cite the reference interface each symbol replaces (paths relative to
 * the reference repository root).
 */
#ifndef TSTEREO_H_
#define TSTEREO_H_

#ifdef __cplusplus
extern "C" {
#endif

#define TSTEREO_VERSION 200            /* 0.2.0 */

#define TSTEREO_OK            0
#define TSTEREO_E_ARG        -1        /* bad size / null pointer / unsupported variant */
#define TSTEREO_E_CUDA       -2        /* CUDA runtime error at launch                  */

#define TSTEREO_ACT_NONE 0
#define TSTEREO_ACT_SILU 1
#define TSTEREO_ACT_RELU 2

int         tstereo_version(void);
const char* tstereo_last_error(void);
/* hash of the sources the library was built from (temporalstereo_b200/build.py:source_id) */
const char* tstereo_build_id(void);
/* number of kernels this library has launched in the process so far (bench.py's gpu_launches) */
long long   tstereo_launch_count(void);

/* ---------------------------------------------------------------- cost volume (a1-a3)
 * ref: architecture/modeling/aggregation/utils/block_cost.py:16-83 (block_cost),
 *      :6-13 (groupwise_correlation), architecture/modeling/layers/inverse_warp_3d.py:4-58.
 * left/right [B,C,H,W]; C % 8 == 0; H,W >= 4; three pooled scales (block_cost_scale = 3).
 * scratch: >= tstereo_block_cost_scratch_floats(B,C,H,W,D) floats.
 */
long long tstereo_block_cost_scratch_floats(int B, int C, int H, int W, int D);

/* int branch (block_cost.py:34-45): out [B, C + 3C/8, D, H, W] =
 *   [ -(L - R_d)^2 , g0, g1, g2 ],  R_d[x] = R[x-d] (0 for x<d), d = 0..D-1 */
int tstereo_block_cost_shift(const float* left, const float* right, float* out, float* scratch,
                             int B, int C, int H, int W, int D, void* stream);

/* tensor branch (block_cost.py:47-58): samples [B,S,H,W];
 * out [B, 2C + 3C/8, S, H, W] = [ L broadcast, R warped to x - sample, g0, g1, g2 ] */
int tstereo_block_cost_warp(const float* left, const float* right, const float* samples, float* out,
                            float* scratch, int B, int C, int H, int W, int S, void* stream);

/* ---------------------------------------------------------------- fused cost volume -> first conv (a1-a3 + a5)
 * ref: block_cost.py:34-58, 64-81 feeding the (1,3,3) half of init3d[0] (module.py:111-147) through coarse.py:82-83,
 *      fine.py:102-103, precise.py:88-90.  The raw cost volume is never written (SURVEY.md section 8d "fused path").
 *
 * tstereo_group_cost_*: only the three group-wise terms, gvol [B, 3C/8, D, H, W] = [g0, g1, g2]
 * (scratch as for tstereo_block_cost_*). */
int tstereo_group_cost_shift(const float* left, const float* right, float* gvol, float* scratch,
                             int B, int C, int H, int W, int D, void* stream);
int tstereo_group_cost_warp(const float* left, const float* right, const float* samples, float* gvol,
                            float* scratch, int B, int C, int H, int W, int S, void* stream);
/* Shift volume in the layout its consumer stages by TMA: the C cost planes -(L - R_d)^2 as S-format chunks [0, C/8) of
 * `sout` (struct tstereo_split, declared with the tensor-core convolutions below; needs both parts) + the compact group terms
 * `gvol` [B, 3C/8, D, H, W] (appended as chunks [C/8, ...) with tstereo_split_pack).  ref: block_cost.py:34-45, 64-81. */
struct tstereo_split;
int tstereo_block_cost_shift_s(const float* left, const float* right, const struct tstereo_split* sout, float* gvol,
                               float* scratch, int B, int C, int H, int W, int D, void* stream);

/* out[B,Cout,D,H,W] (strided view) = act(conv3x3(cost volume) + bias) on the tensor cores, the volume's feature channels
 * rebuilt by the producer from the feature maps and its group channels read from gvol.
 *   warp : virtual input channels [R warped (C) | gvol (3C/8)]; the candidate-invariant left half of the volume enters as
 *          addL [B, Cout, H, W] = conv3x3(left, W[:, :C]) (no bias), computed once per frame by tstereo_conv_hw3_tc2;
 *          wpack = tstereo_conv_hw3_tc2 image of W[:, C:] ([Cout][C + 3C/8][9]).
 *   shift: virtual input channels [-(L - R_d)^2 (C) | gvol (3C/8)]; wpack = image of the whole W ([Cout][C + 3C/8][9]).
 * tstereo_cost_conv_wpack_floats(C, Cout, half) floats. */
long long tstereo_cost_conv_wpack_floats(int C, int Cout, int half);
int tstereo_cost_conv_warp(const float* right, const float* samples, const float* gvol, const float* addL,
                           float* out, long long osB, long long osC, long long osD,
                           const float* wpack, const float* bias, const float* oscale,
                           int B, int C, int Cout, int S, int H, int W, int act, int half, void* stream);
int tstereo_cost_conv_shift(const float* left, const float* right, const float* gvol,
                            float* out, long long osB, long long osC, long long osD,
                            const float* wpack, const float* bias, const float* oscale,
                            int B, int C, int Cout, int D, int H, int W, int act, int half, void* stream);

/* "Tap projection" form of the warp levels' first conv (the engine's default for the fine / precise levels, round 2).
 * The warp is a per-position lerp of two columns of R, the same for every channel, so the channel contraction commutes
 * with it:  sum_c W[co,c,t] * Rw[c,d,p] = wa * T[t,co,y,xa] + wb * T[t,co,y,xa+1]  with  T[t*Cout+co] = sum_c W[co,C+c,t] * R[c]
 * ONE 1x1 convolution of the right features per frame (tstereo_conv_d_tc2, k = 1).  This entry point does the rest:
 *   out[b,co,d,y,x] = act( bias[co] + addL[b,co,y,x] + gconv[b,co,d,y,x]
 *                          + sum_{ky,kx} lerp_x( T[b,(ky*3+kx)*Cout+co, y+ky-1, .]  at  (x+kx-1) - samples[b,d,y+ky-1,x+kx-1] ) )
 * T [B, 9*Cout, H, W]; gconv [B, Cout, S, H, W] = the (1,3,3) conv over the group-wise channels (tstereo_group_cost_warp),
 * no bias / activation, or NULL; addL as in tstereo_cost_conv_warp, or NULL.  `out` (strided fp32) and / or `sout` (S-format,
 * declared below with the tensor-core convolutions).  Cout = 8 | 16 | 32.
 * ref: block_cost.py:47-58 -> module.py:111-147 (zero padding of the VOLUME: neighbours outside the image contribute 0). */
struct tstereo_split;
int tstereo_cost_taps(const float* T, const float* samples, const float* gconv, const float* addL, const float* bias,
                      float* out, long long osB, long long osC, long long osD, const struct tstereo_split* sout,
                      int B, int Cout, int S, int H, int W, int act, void* stream);

/* ---------------------------------------------------------------- convolutions (a4-a7, a9, a12, a13)
 * ref: architecture/modeling/layers/basic_layers.py:194-235 (Conv3d), :340-388 (ConvTranspose3d),
 *      architecture/modeling/aggregation/TemporalStereo/module.py:111-184 (separable pairs).
 * Activations are 5-D views [B, C, D, H, W] with element strides (sB, sC, sD) and a contiguous
 * H*W plane (2-D convs use D = 1).  Weights are pre-packed by the host with eval-mode BatchNorm
 * folded in: w[Cin][taps][CoutP] (CoutP = Cout rounded up to 4), bias[Cout] (may be NULL).
 */

/* (1,3,3) / 3x3 convolution over (H,W): stride 1|2, dilation 1|2, padding = dilation. */
int tstereo_conv_hw3(const float* in, long long isB, long long isC, long long isD,
                     float* out, long long osB, long long osC, long long osD,
                     const float* w, const float* bias,
                     int B, int Cin, int Cout, int D, int Hin, int Win, int Hout, int Wout,
                     int stride, int dilation, int act, void* stream);

/* Tensor-core 3x3 conv (tcgen05 implicit GEMM; stride 1, dilation 1|2, any Cout in groups of 32): 2-D tiles, the three kx
 * taps folded into the MMA's N dimension, error-compensated hi+lo split operands (same results as tstereo_conv_hw3 to fp32
 * rounding).
 * ref: layers/basic_layers.py:194-235 via aggregation/TemporalStereo/module.py:111-147, 424-492.
 * wpack: [ceil(Cin/8)][ky 3][khalf 2][row 2N][4] floats, N = 3*CP, CP = 8|16|32 >= Cout,
 * row = part*N + kx*CP + co (part 0 = hi, 1 = lo); tstereo_conv_hw3_tc2_wpack_floats(Cin, Cout, half) floats.
 * half != 0 (all four tc2 entry points): operands split as fp16 hi + fp16 lo and multiplied with kind::f16 — 16 channels per
 * chunk, each 16-byte row of the image holds 8 halves instead of 4 floats (ops.py pack_*(..., half=True)); activations must
 * stay below 65504 in magnitude.  half == 2: the same fp16 image, but only the A_hi * B_hi product is issued (11-bit
 * operands, one MMA term instead of three) -- for layers whose measured sensitivity allows it (the UNet
 * decoder after the top-2 selection).
 * oscale [Cout] or NULL: multiplier of the accumulator before the bias (1 / the weight pre-scale of ops.fp16_prescale). */
long long tstereo_conv_hw3_tc2_wpack_floats(int Cin, int Cout, int half);
int tstereo_conv_hw3_tc2(const float* in, long long isB, long long isC, long long isD,
                         float* out, long long osB, long long osC, long long osD,
                         const float* wpack, const float* bias, const float* oscale,
                         int B, int Cin, int Cout, int D, int H, int W,
                         int dilation, int act, int half, void* stream);
/* Stride-2 3x3 conv (padding 1) and stride-2 transposed convs through the same tensor-core kernel
 * (tstereo_conv_hw3_tc2 over a virtual tensor: input parity phases stacked on the channel axis / one output parity
 * phase per launch).  Any Cout (groups of 32).  ref: aggregation/TemporalStereo/module.py:111-184 (stride-2
 * "DepthwiseConv3D" halves, "DepthwiseConvTranspose3D" k3 s2 p1 op1), :424-492 (UNet stride-2 convs, 4x4 deconvs).
 * wpack (host repack, temporalstereo_b200/ops.py pack_conv_hw3s2_tc2 / pack_deconv_hw_tc2):
 *   s2:     per 32-channel group the tc2 image of w'[Cout][4*Cin8][3][3], virtual channel = phase*Cin8 + c,
 *           phase = (row parity, col parity), Cin8 = Cin rounded up to 8;
 *   deconv: [phase (py,px) 4] x per 32-channel group the tc2 image of the 3x3 shift kernel of that output phase
 *           (k = 3: padding 1, output_padding 1;  k = 4: padding 1; both give Hout = 2*Hin). */
long long tstereo_conv_hw3s2_tc2_wpack_floats(int Cin, int Cout, int half);
int tstereo_conv_hw3s2_tc2(const float* in, long long isB, long long isC, long long isD,
                           float* out, long long osB, long long osC, long long osD,
                           const float* wpack, const float* bias, const float* oscale,
                           int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream);
long long tstereo_deconv_hw_tc2_wpack_floats(int Cin, int Cout, int half);
int tstereo_deconv_hw_tc2(const float* in, long long isB, long long isC, long long isD,
                          float* out, long long osB, long long osC, long long osD,
                          const float* wpack, const float* bias, const float* oscale,
                          int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream);
/* (k,1,1) conv along D (same argument meaning as tstereo_conv_d) through the tensor-core
 * kernel: the k input planes are K-chunks of a 1x1 conv.  wpack: pack_conv_d_tc2 in temporalstereo_b200/ops.py,
 * tstereo_conv_d_tc2_wpack_floats(Cin, Cout, k) floats. */
long long tstereo_conv_d_tc2_wpack_floats(int Cin, int Cout, int k, int half);
int tstereo_conv_d_tc2(const float* in, long long isB, long long isC, long long isD,
                       float* out, long long osB, long long osC, long long osD,
                       const float* wpack, const float* bias, const float* oscale,
                       int B, int Cin, int Cout, int Din, int Dout, int H, int W,
                       int k, int stride, int dilation, int transposed, int act, int half, void* stream);

/* ---- S-format ("split") activations: the TMA-fed form of the tensor-core convolutions (round 2).
 * An S-format tensor holds the fp16 hi / lo halves of an fp32 activation x (hi = fp16(x), lo = fp16(x - hi): exactly the
 * operand split the tensor-core kernels compute from fp32 inputs), laid out [B][D][part][C8][H][W][8 channels] with
 * C8 = ceil(C / 8) and zero padding channels, so that one K-chunk of the implicit GEMM's A operand is ONE TMA box
 * (cp.async.bulk.tensor) from global memory straight into the K-major shared-memory stage: no producer warps, no
 * per-consumer conversion.  The epilogue of the producing layer writes it (the split is computed once per element).
 * Results are bit-identical to the fp32-input entry points above (same operand values, same MMA order).
 * ref: the same reference layers as the _tc2 entry points (layers/basic_layers.py:194-235, 340-388). */
typedef struct tstereo_split {
    void* ptr;                   /* fp16 bit patterns */
    long long sB, sD, sP, sC8;   /* element (fp16) strides: batch, plane, part (hi -> lo), 8-channel chunk; [H][W][8] is dense */
    int C8;                      /* 8-channel chunks addressable from ptr */
    int parts;                   /* 2: hi + lo;  1: hi only (inputs / outputs of the single-term form, half == 2) */
    int nb;                      /* as an output: only batches [0, nb) are written (0: all) */
} tstereo_split;

/* fp32 [B, C, D, H, W] view (dense H*W plane) -> S-format (for activations produced by a non-convolution kernel or by
 * the caller, e.g. the backbone features). */
int tstereo_split_pack(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sout,
                       int B, int C, int D, int H, int W, void* stream);

/* The four tensor-core convolutions with S-format operands.  Arguments as in the _tc2 forms (same packed weights), plus
 *   sin  : S-format input, or NULL (the fp32 input `in` is converted by the kernel's producer warps);
 *   sout : S-format output written by the epilogue, or NULL;  `out` (fp32) may be NULL when sout is given.
 * half must be 1 or 2.  Stride-2 convolutions read the four parity phases of an S-format input through the TMA map's
 * element strides; transposed convolutions write one output parity phase per launch into either format. */
int tstereo_conv_hw3_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                       float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                       const float* wpack, const float* bias, const float* oscale,
                       int B, int Cin, int Cout, int D, int H, int W,
                       int dilation, int act, int half, void* stream);
int tstereo_conv_hw3s2_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                         float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                         const float* wpack, const float* bias, const float* oscale,
                         int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream);
int tstereo_deconv_hw_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                        float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                        const float* wpack, const float* bias, const float* oscale,
                        int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream);
int tstereo_conv_d_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                     float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                     const float* wpack, const float* bias, const float* oscale,
                     int B, int Cin, int Cout, int Din, int Dout, int H, int W,
                     int k, int stride, int dilation, int transposed, int act, int half, void* stream);


/* (k,1,1) convolution along D: k = 3|5, stride 1|2, dilation 1|2, padding = dilation*(k/2).
 * transposed != 0: ConvTranspose (3,1,1) stride 2, padding 1, output_padding 1 (Dout = 2*Din). */
int tstereo_conv_d(const float* in, long long isB, long long isC, long long isD,
                   float* out, long long osB, long long osC, long long osD,
                   const float* w, const float* bias,
                   int B, int Cin, int Cout, int Din, int Dout, int HW,
                   int k, int stride, int dilation, int transposed, int act, void* stream);

/* Transposed (1,k,k) / kxk convolution over (H,W), stride 2, padding 1:
 * k = 3 with output_padding 1 (module.py:149-184) or k = 4 (module.py:452-457); Hout = 2*Hin. */
int tstereo_deconv_hw(const float* in, long long isB, long long isC, long long isD,
                      float* out, long long osB, long long osC, long long osD,
                      const float* w, const float* bias,
                      int B, int Cin, int Cout, int D, int Hin, int Win,
                      int k, int act, void* stream);

/* out[b, c, :] = in[b, c, :] for dense in [B, C, HW] and an out view with element strides (osB, osC): writes one
 * operand of a channel concat in place (ref: torch.cat at aggregation/TemporalStereo/precise.py:86). */
int tstereo_copy_planes(const float* in, float* out, long long osB, long long osC, int B, int C, int HW, void* stream);

/* out = act( trilinear_align_corners(a -> (D,H,W)) + skip )   (module.py:285-295)
 * a [B,C,Da,Ha,Wa], skip/out [B,C,D,H,W] contiguous; skip may be NULL. */
int tstereo_resize_add_act(const float* a, const float* skip, float* out,
                           int B, int C, int Da, int Ha, int Wa, int D, int H, int W,
                           int act, void* stream);

/* The same operator writing the S-format (`tstereo_split`, declared with the tensor-core convolutions above)
 * that the consumer convolution stages by TMA; values are the fp16 hi / lo split of tstereo_resize_add_act's result. */
int tstereo_resize_add_act_s(const float* a, const float* skip, const tstereo_split* sout,
                             int B, int C, int Da, int Ha, int Wa, int D, int H, int W,
                             int act, void* stream);

/* avg_pool3d and max_pool3d, kernel 5, stride 1, padding 2 (module.py:416-417).
 * x [B,C,D,H,W] view with strides; results written to two channel slices (strided views). D <= 24. */
int tstereo_pool5(const float* x, long long xsB, long long xsC,
                  float* avg, float* mx, long long osB, long long osC,
                  int B, int C, int D, int H, int W, void* stream);

/* ---------------------------------------------------------------- memory merge (a8)
 * ref: coarse.py:84-105, fine.py:104-122.  vol [B,C,D,H,W], samples [B,D,H,W],
 * mem_sample / mem_cost [B,M,H,W] (NULL = zeros), past_conv folded to w[C], b[C] (+SiLU).
 * Stable sort of the D+M candidates per pixel; out_vol [B,C,D+M,H,W] (strided view:
 * osB, osC; plane stride H*W), out_samples [B,D+M,H,W].  D+M <= 32. */
int tstereo_merge_memory(const float* vol, const float* samples,
                         const float* mem_sample, const float* mem_cost,
                         const float* past_w, const float* past_b,
                         float* out_vol, long long osB, long long osC, float* out_samples,
                         int B, int C, int D, int M, int H, int W, void* stream);

/* ---------------------------------------------------------------- heads + regression (a10, a11, a14)
 * ref: module.py:356-398 (PredictionHeads), coarse.py:69-75 (predict_disp), fine.py:78-86.
 * feat [B,2C,D,H,W]: channels [0,C) = cost-head features, [C,2C) = offset-head features
 * (both after the (3,1,1) conv + BN + SiLU).  w [2][C][9].
 * cost, off [B,D,H,W]; off = clamp(tanh(x/100),-1,1)*delta. */
int tstereo_heads(const float* feat, const float* w, float* cost, float* off,
                  int B, int C, int D, int H, int W, float delta, void* stream);

/* top-2 over D -> softmax -> gather(sample+off) -> sum.  disp [B,1,H,W];
 * top_disp / top_cost [B,2,H,W] may be NULL.  Ties: lowest index first. */
int tstereo_predict_disp(const float* cost, const float* samples, const float* off,
                         float* disp, float* top_disp, float* top_cost,
                         int B, int D, int H, int W, void* stream);

/* low = disp - r, high = disp + r, samples[:, c_off + k] = |high-low|*{0,3,4,5,8}/8 + min(low,high)
 * samples is [B, S_total, H, W]; low/high [B,1,H,W]. */
int tstereo_range_samples(const float* disp, float radius, float* low, float* high,
                          float* samples, int S_total, int c_off,
                          int B, int H, int W, void* stream);

/* ---------------------------------------------------------------- up-sampling (a12, a13, a16, a20)
 * ConvexUpsample tail (module.py:318-353): m [B,64,H,W] (mask.0+BN+SiLU output),
 * w [36][64], b [36], disp [B,1,H,W] -> out [B,1,2H,2W]. */
int tstereo_convex_upsample(const float* m, const float* w, const float* b, const float* disp,
                            float* out, int B, int H, int W, void* stream);

/* UNet.upsample (module.py:468-483): logits [B,9,H,W], disp [B,1,h,w] -> full [B,1,H,W]. */
int tstereo_unet_upsample(const float* logits, const float* disp, float* full,
                          int B, int H, int W, int h, int w, void* stream);

/* Bilinear align_corners resize of (in * mul / div): in [B,C,Hi,Wi] ->
 * out[:, c_off : c_off+C] of a [B, C_total, Ho, Wo] tensor.
 * (coarse.py:92-95, fine.py:91, precise.py:100-103, projects/TemporalStereo/TemporalStereo.py:305-309) */
int tstereo_bilinear_resize(const float* in, float* out, float mul, float div,
                            int B, int C, int Hi, int Wi, int Ho, int Wo,
                            int C_total, int c_off, void* stream);

/* ---------------------------------------------------------------- temporal warp (a17-a19)
 * ref: projects/TemporalStereo/TemporalStereo.py:326-461 (update_map),
 *      architecture/modeling/layers/inverse_warp.py:92-178 (project_to_3d),
 *      architecture/modeling/layers/softsplat.py:8-53, 334-360 (softmax splatting).
 *
 * pose_prep: per batch item, T = T_now @ inv_T_prev, down_K = K with rows 0,1 / factor,
 * inv(down_K), P = (down_K @ T)[:3,:].  params [B,24] = invK(9) | P(12) | focal | baseline | pad. */
int tstereo_pose_prep(const float* K, const float* T_now, const float* inv_T_prev,
                      const float* baseline, float factor, float* params, int B, void* stream);

/* disparity -> depth -> 3-D -> new camera -> (flow of channel 0, new disparity of every channel).
 * disp [B,C,h,w]; flow [B,2,h,w] (may be NULL); new_disp [B,C_total,h,w] at c_off (may be NULL). */
int tstereo_reproject_disp(const float* disp, const float* params, float* flow, float* new_disp,
                           int C_total, int c_off, int B, int C, int h, int w, void* stream);

/* project_to_3d drop-in (inverse_warp.py:92-178): depth [B,C,h,w] -> optical_flow [B,2C,h,w]
 * (x,y pairs per channel; may be NULL) and triangular_depth [B,C,h,w] (may be NULL).
 * params from tstereo_pose_prep (factor 1, inv_T_prev = identity reproduces K @ T). */
int tstereo_project_to_3d(const float* depth, const float* params, float* flow, float* tri,
                          int B, int C, int h, int w, void* stream);

/* metric = clamp(pd[:, :1] - mean(pd[:, :1]), -50, 50) over the whole local batch.
 * pd [B,C,h,w]; metric [B,1,h,w]; scratch >= 1024 floats. */
int tstereo_splat_metric(const float* pd, float* metric, float* scratch,
                         int B, int C, int h, int w, void* stream);

/* FunctionSoftsplat(x, flow, metric, 'softmax'): x [B,C,h,w], flow [B,2,h,w], metric [B,1,h,w],
 * acc scratch [B,C+1,h,w] (zeroed by the call), out [B,C,h,w]. */
int tstereo_softsplat(const float* x, const float* flow, const float* metric, float* acc,
                      float* out, int B, int C, int h, int w, void* stream);

/* Fused temporal warp: the whole of update_map in three launches (prep: down-sample prev_disp + partial sums + zeroed
 * accumulators; splat: pose, flow, metric, re-projection of the stored top-2 samples and of the local-map stack, softmax
 * splat of both groups; normalise).  ref: projects/TemporalStereo/TemporalStereo.py:326-461.
 *   prev_disp [B,1,Hf,Wf]; K, T_now, inv_T_prev [B,4,4]; baseline [B];
 *   mem_sample / mem_cost [B,M,h,w] -> out_sample / out_cost [B,M,h,w]          (all four NULL: no cost memory)
 *   local_map [B,n_lm_in,h,w] (NULL when n_lm_in == 0) -> out_lm [B,n_lm_out,h,w], n_lm_out <= n_lm_in + 1: the warped
 *   stack [prev_disp at 1/8, local_map][:n_lm_out]                               (n_lm_out == 0, out_lm NULL: no local map)
 *   scratch: tstereo_update_map_scratch_floats(B, h, w, M, n_lm_out) floats, 8-byte aligned. */
long long tstereo_update_map_scratch_floats(int B, int h, int w, int M, int n_lm_out);
int tstereo_update_map(const float* prev_disp, int Hf, int Wf, const float* K, const float* T_now, const float* inv_T_prev,
                       const float* baseline, const float* mem_sample, const float* mem_cost, int M,
                       const float* local_map, int n_lm_in, int n_lm_out,
                       float* out_sample, float* out_cost, float* out_lm, float* scratch,
                       int B, int h, int w, void* stream);

/* ---------------------------------------------------------------- formats either side of the path
 * Wire format of the images: uint8 HWC in (one quarter of the fp32 bytes over PCIe), ImageNet-normalised fp32 planes out,
 * written into a view with element strides (osB, osC).  mean3 / std3 are HOST pointers to 3 floats.
 * ref: architecture/data/datasets/base.py:120-127 (ToTensor + Normalize, same op order: bit-identical). */
int tstereo_normalize_u8(const unsigned char* in, float* out, long long osB, long long osC, int B, int H, int W,
                         const float* mean3, const float* std3, void* stream);
/* On-device evaluation: acc6 (6 doubles, device) = { sum |gt - est|, mask count, #err>1, #err>2, #err>3, #err>5 } over the
 * pixels with lb < gt < ub (bounds used when use_lb / use_ub != 0) -- a 48-byte read-back instead of two full-resolution maps.
 * ref: architecture/data/evaluation/pixel_error.py:6-71 (calc_error), eval.py:31-35 (.clone().cpu()). */
int tstereo_disp_error(const float* est, const float* gt, float lb, float ub, int use_lb, int use_ub, long long n,
                       double* acc6, void* stream);

/* ---------------------------------------------------------------- training losses, forward only
 * ref: architecture/modeling/losses/smooth_l1_loss.py:49-74, warsserstein_distance_loss.py:53-81 (loss_per_level).
 * gt [B,1,Hg,Wg] is scaled (gt / (Wg/W)) and adaptively pooled (average: dense, max: sparse) onto the level's (H, W) grid;
 * valid where start_disp < gt_s < max_disp / (Wg/W).  acc2 (2 doubles, device): [sum, number of valid pixels] —
 *   smooth L1:    loss = sum / count                     (0 when no pixel is valid)
 *   Wasserstein:  loss = sum / (B*H*W),  sum over valid pixels of  sum_d (softmax_d(cost) + 0.25) * |off_d + sample_d - gt_s| */
int tstereo_loss_smooth_l1(const float* est, const float* gt, int B, int H, int W, int Hg, int Wg, float max_disp, float start_disp,
                           int sparse, double* acc2, void* stream);
int tstereo_loss_wasserstein(const float* cost, const float* off, const float* samples, const float* gt, int B, int D, int H, int W,
                             int Hg, int Wg, float max_disp, float start_disp, int sparse, double* acc2, void* stream);


#ifdef __cplusplus
}
#endif
#endif  /* TSTEREO_H_ */
