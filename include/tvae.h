/* libtvae_b200 — C ABI of the B200-native TEMPO-VAE hot path (sm_100a).
 *
 * The reference (cfpark00/TEMPO-VAE) has no FFI layer: its hot path is eager PyTorch
 * (src/model.py, src/model_with_l2.py, src/train_utils.py:149-183). Each entry point below therefore cites the
 * reference call site(s) whose arithmetic it replaces. INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - every function returns 0 on success, < 0 on error; tvae_last_error() returns a thread-local message;
 *  - the caller owns all device memory (incl. workspaces); the library never allocates device memory, never
 *    synchronises and never changes the current device; work is enqueued on `stream`;
 *  - activations are NHWC ("channels last"): element (n, h, w, c) of a tensor with channel pitch `pitch` lives at
 *    ((n*H + h)*W + w)*pitch + c. bf16 tensors need pitch % 8 == 0, fp32 tensors pitch % 4 == 0, bases 16-B aligned;
 *  - parameters keep the reference's own layouts (Conv2d OIHW fp32, ConvTranspose2d [Cin][Cout][kH][kW] fp32).
 */
#ifndef TVAE_H_
#define TVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVAE_ABI_VERSION 1

typedef struct CUstream_st* tvae_stream_t; /* == cudaStream_t */

const char* tvae_last_error(void);
int32_t tvae_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Convolution forward / data-gradient as an implicit GEMM on tcgen05 tensor cores.
 * Replaces nn.Conv2d / nn.ConvTranspose2d forward (src/model.py:21-42; call sites :107-118,181,205,208-210,
 * 240-247,270-278,358,402,502,544,609-614; src/model_with_l2.py:23,30) and autograd's input-gradient for them.
 *   kind 0: RxR (R in {1,3}) stride-1 "same" convolution over x[N,H,W,C]; flip=1 negates the tap offsets (dgrad).
 *           w packed [rows >= Cout][R*R*c_pad] (tvae_pack_weight).
 *   kind 1: 2x2 stride-2 convolution, x[N,H,W,C] -> [N,H/2,W/2,Cout]; w packed [rows >= Cout][4*c_pad].
 *   kind 2: 2x2 stride-2 transposed convolution, x[N,H,W,C] -> [N,2H,2W,Cout]; w packed [4*Cout][c_pad].
 * Epilogue: + bias[Cout] (optional) + residual (optional, fp32, indexed like the output), then written as fp32
 * and/or bf16 (either pointer may be NULL, not both).
 */
typedef struct {
  const void* x;        /* bf16 NHWC */
  int32_t N, H, W, C, x_pitch;
  int32_t kind, R, flip;
  const void* w;        /* bf16 packed, K-major */
  int32_t w_rows, k_pitch, c_pad;
  int32_t Cout;
  const float* bias;
  const float* residual;
  int32_t res_pitch;
  float* out_f32;
  int32_t out_f32_pitch;
  void* out_bf16;
  int32_t out_bf16_pitch;
  int32_t bn;           /* N tile, 0 = auto */
  /* Optional fused GroupNorm statistics of the OUTPUT (after bias/residual): per-tile partial sums
   * stats_part[slot][stats_groups][2] = (sum, sum of squares), slot = 128-pixel tile index (kind 0/1) or
   * 4*tile + tap (kind 2). Needs Cout/stats_groups to be a multiple of 16 dividing the N tile and images of at least
   * 128 output pixels; tvae_gn_stats_finalize turns the partials into (mean, rstd). NULL = off. */
  float* stats_part;
  int32_t stats_groups;
  /* "fp32 mode" (split-bf16 emulation, ~2^-16 relative product error, 3x the tensor work): low-order halves of the
   * operands, value = hi + lo with hi = bf16(value), lo = bf16(value - hi). Same layouts/pitches as x and w. The
   * kernel accumulates x*w + x*w_lo + x_lo*w in fp32. Both NULL = plain bf16. out_bf16_lo (optional, pitch of
   * out_bf16) receives the low-order half of the bf16 output for a split-bf16 consumer. */
  const void* x_lo;
  const void* w_lo;
  void* out_bf16_lo;
  /* Optional fused reconstruction loss for decoder.conv_out under AutoencoderKL.get_loss (src/model.py:656-663): with
   * nll_x = the target (bf16 NHWC [pixels][nll_x_pitch], Cout channels) the epilogue forms d = (conv + bias) - target
   * straight from the fp32 accumulators, accumulates sum |d| (nll_loss_type 0, "l1") or sum d^2 (1, "l2") and sum d^2,
   * and writes the loss gradient wrt the reconstruction, exp(-logvar) / nll_batch * sign(d) resp. * 2 d, as the bf16
   * output (pad lanes up to out_bf16_pitch zeroed) -- the reconstruction itself never reaches HBM (4 B/element less
   * written, and the separate pass of tvae_nll_fwd, 8 B/element, disappears). Requires kind 0, flip 0, out_bf16 only.
   * nll_sums: double[3] like tvae_nll_fwd's; nll_workspace: tvae_conv_nll_workspace_bytes(pixels, Cout). NULL = off. */
  const void* nll_x;
  int32_t nll_x_pitch;
  int32_t nll_loss_type;
  const float* nll_logvar;
  int32_t nll_batch;
  float* nll_workspace;
  double* nll_sums;
} tvae_conv_args;
int32_t tvae_conv_gemm(const tvae_conv_args* args, tvae_stream_t stream);
int64_t tvae_conv_nll_workspace_bytes(int64_t pixels, int32_t Cout);
/* Scheduling switch (results are bit-identical either way): 1 (default) runs tvae_conv_gemm as clusters of two CTAs
 * that share one 256-row tcgen05 MMA (cta_group::2, each SM stages half of the weight tile, a third less operand
 * traffic per SM); 0 runs one CTA per SM. Returns the previous setting. Process-wide; for A/B measurements and tests. */
int32_t tvae_conv_set_cta_pair(int32_t enable);
/* Profiling aid: a per-tile timeline of the conv kernel's producer / MMA / epilogue warps (SM clock cycles, 8 u64 words
 * per (unit, tile): see ConvParams::trace in csrc/conv_gemm.cu and tools/conv_trace.py). device_buffer = NULL turns it
 * off (default); the buffer holds num_SMs * tiles_per_unit * 8 words. */
int32_t tvae_conv_set_trace(void* device_buffer, int32_t tiles_per_unit);

/* Weight gradient: grad[m][n][tap] (=|+=) sum_pixels P[pixel][m] * Q[pixel (+) tap][n].
 * Replaces autograd's weight-gradient of the same call sites.
 *   kind 0: Q on the same [N,H,W] grid, RxR taps with zero padding   (Conv2d: P = dY, Q = x)
 *   kind 1: Q on the 2x finer grid [N,2H,2W], tap (ty,tx) = pixel (2h+ty, 2w+tx)
 *           (2x2 s2 Conv2d: P = dY, Q = x;  2x2 s2 ConvTranspose2d: P = x, Q = dY)
 * workspace: tvae_wgrad_workspace_bytes(Cm, Cn, ntaps, splits) bytes of fp32 scratch (split-K partials).
 */
typedef struct {
  const void* p;        /* bf16 NHWC [N,H,W,Cm] */
  int32_t p_pitch, Cm;
  const void* q;        /* bf16 NHWC */
  int32_t q_pitch, Cn;
  int32_t N, H, W;      /* grid of P */
  int32_t kind, R;
  int32_t splits;       /* >= 1; tvae_wgrad_splits() suggests a value */
  float* workspace;
  float* grad;          /* fp32 [Cm][Cn][taps] */
  int32_t accumulate;   /* 0: overwrite grad, 1: add into it */
  /* kind 0 only. flip = 1 exchanges the operand roles: P = x (dense, Cm = Cin), Q = dY read at pixel (-) tap
   * (Cn = Cout); grad is then written as [Cn][Cm][taps], i.e. still the Conv2d OIHW layout. Lets the 1028-channel
   * side of encoder.conv_in sit on the GEMM's M dimension (9 row tiles, 11 % padding) instead of its N dimension
   * (5 column tiles of 208 with 23 % wasted operand loads). */
  int32_t flip;
  /* Sub-block of a larger parameter: the GEMM covers only Cm x Cn of a parameter whose INNER channel dimension (Cn without
   * flip, Cm with flip) has grad_ld entries, starting at grad_off. 0 / 0 = the parameter is exactly Cm x Cn. Lets a
   * 1028-channel weight gradient be computed as a 1024-channel GEMM (whole 128-row tiles) plus a skinny 4-channel one
   * with the 4 channels on the N side (16-column MMAs) instead of a 128-row tile that is 97 % padding. */
  int32_t grad_ld, grad_off;
} tvae_wgrad_args;
int32_t tvae_wgrad_gemm(const tvae_wgrad_args* args, tvae_stream_t stream);
/* Scheduling switch like tvae_conv_set_cta_pair: 1 (default) pairs adjacent 128-row M tiles on CTA pairs (cta_group::2);
 * an odd last M tile runs as a second one-CTA-per-SM launch with its own split-K factor. Call before
 * tvae_wgrad_splits / tvae_wgrad_workspace_bytes. Returns the previous setting. */
int32_t tvae_wgrad_set_cta_pair(int32_t enable);
int64_t tvae_wgrad_workspace_bytes(int32_t Cm, int32_t Cn, int32_t ntaps, int32_t splits);
int32_t tvae_wgrad_splits(int32_t Cm, int32_t Cn, int32_t ntaps, int64_t pixels);

/* Weight gradient of the last 1..4 channels of a wide, awkward channel count (1028 = 8 x 128 + 4: the input channels
 * of encoder.conv_in, the output channels of decoder.conv_out; src/model.py:424-431, 634-640) for a 3x3 stride-1,
 * zero-padded convolution -- the part of autograd's convolution_backward (weight) that tvae_wgrad_gemm would pay a
 * whole padded 128-row tile for:
 *   grad[c * stride_c + n * stride_n + tap] (+)= sum_pixels wide[pixel][n] * skinny[pixel + shift_sign * tap][c]
 * wide: bf16 NHWC [N*H*W][wide_pitch], Cw channels (multiple of 64, <= 512); skinny: bf16 [N*H*W][skinny_pitch], Cs
 * channels starting at the pointer. shift_sign = +1: the tap shifts the skinny operand (conv_in: wide = dY, skinny =
 * x[..., 1024:], grad = dW[n][1024 + c][tap]: stride_c = 9, stride_n = 9 * Cin, grad pointer advanced by 9 * 1024);
 * -1: it shifts the wide one (conv_out: wide = x, skinny = dY[..., 1024:], grad = dW[1024 + c][n][tap]: stride_c =
 * 9 * Cin, stride_n = 9). The wide operand is read once (taps on the GEMM's M side), partial sums per CTA are added in
 * fixed order (bit-reproducible). workspace: tvae_wgrad_skinny_workspace_bytes(Cw). */
int64_t tvae_wgrad_skinny_workspace_bytes(int32_t Cw);
int32_t tvae_wgrad_skinny(const void* wide_bf16, int32_t Cw, int32_t wide_pitch, const void* skinny_bf16, int32_t Cs,
                          int32_t skinny_pitch, int32_t N, int32_t H, int32_t W, int32_t shift_sign, float* grad,
                          int64_t stride_c, int64_t stride_n, int32_t accumulate, float* workspace,
                          tvae_stream_t stream);

/* out[(tr*Crow + cr)][tk*c_pad + c] = bf16(w[cr*s_row + c*s_col + (tr+tk)*s_tap]), zero for c in [C, c_pad).
 * One of TR, TK is 1. Row pitch of `out` is TK*c_pad. out_lo (optional, same layout) = bf16(w - out): the low-order
 * half for the split-bf16 "fp32 mode" (every *_lo argument below has the same meaning; NULL = not produced). */
/* All packs that went stale in an optimiser step, in ONE launch. `descs` (device memory) holds one descriptor per pack
 * with the arguments of tvae_pack_weight; block_start (device, n + 1 entries) is the exclusive prefix sum of
 * ceil(elements_i / tvae_pack_chunk_elems()) and total_blocks its last entry. */
typedef struct {
  const float* w;
  void* out_bf16;
  int32_t Crow, TR, TK, C, c_pad, reserved;
  int64_t s_row, s_col, s_tap;
} tvae_pack_desc;
int32_t tvae_pack_chunk_elems(void);
int32_t tvae_pack_weights_batched(const tvae_pack_desc* descs, const int64_t* block_start, int32_t n,
                                  int64_t total_blocks, tvae_stream_t stream);
int32_t tvae_pack_weight(const float* w, void* out_bf16, int32_t Crow, int32_t TR, int32_t TK, int32_t C,
                         int32_t c_pad, int64_t s_row, int64_t s_col, int64_t s_tap, void* out_lo,
                         tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Layout conversion at the API boundary (the reference's tensors are NCHW fp32; src/train_utils.py:154).
 */
int32_t tvae_nchw_f32_to_nhwc_bf16(const float* x, void* out_bf16, int32_t N, int32_t C, int32_t HW,
                                   int32_t out_pitch, void* out_lo, tvae_stream_t stream);
/* Radiance normalisation of the data preparation / analysis scripts (src/scripts/prepare_tempo_tiles.py:67-79 with
 * global statistics): z[r][c] = clamp((log(max(rad[r][c], min_radiance)) - mean[c]) / (std[c] + 1e-8), clip_min, clip_max)
 * for raw radiance rows [rows][C] (a granule is [mirror][track][C]). Writes fp32 rows (out_f32, pitch C) and/or the bf16
 * channels-last operand rows (out_bf16, pitch out_pitch, pad lanes zeroed) in one pass. */
int32_t tvae_normalize_radiance(const float* rad, const float* mean, const float* std, int64_t rows, int32_t C,
                                float min_radiance, float clip_min, float clip_max, float* out_f32, void* out_bf16,
                                int32_t out_pitch, tvae_stream_t stream);
/* Channels-last fp32 pixels (row pitch in_pitch elements, e.g. the reference's on-disk [H][W][1028] tiles or a
 * torch.channels_last tensor) -> channels-last bf16 operand rows (pitch out_pitch, pad lanes zeroed): the layout the
 * conv kernels read, without the NCHW detour of src/tempo_data.py:98-99 + the transpose above. */
int32_t tvae_nhwc_f32_to_nhwc_bf16(const float* x, int64_t in_pitch, int64_t rows, int32_t C, void* out_bf16,
                                   int32_t out_pitch, void* out_lo, tvae_stream_t stream);
int32_t tvae_nhwc_f32_to_nchw_f32(const float* x, float* out, int32_t N, int32_t C, int32_t HW, int32_t in_pitch,
                                  tvae_stream_t stream);
int32_t tvae_nhwc_bf16_to_nchw_f32(const void* x_bf16, float* out, int32_t N, int32_t C, int32_t HW,
                                   int32_t in_pitch, tvae_stream_t stream);
int32_t tvae_f32_to_bf16(const float* x, void* out_bf16, int64_t n, void* out_lo, tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * GroupNorm (+ exact-erf GELU). Replaces nn.GroupNorm + nn.GELU (src/model.py:105,179,202,333-339,400,542;
 * src/model_with_l2.py:24-25) and their backward.
 *  x fp32 NHWC [N,HW,C] (pitch C); stats[N][G][2] = (mean, rstd); act: 0 = identity, 1 = GELU (exact erf), 2 = ReLU, 3 = SiLU.
 */
int32_t tvae_gn_stats(const float* x, int32_t N, int32_t HW, int32_t C, int32_t G, float eps, float* stats,
                      tvae_stream_t stream);
/* stats[n][g] = (mean, rstd) from conv-epilogue partials: image n owns slots [n*slots_per_image, (n+1)*slots_per_image);
 * count = elements per (image, group). */
int32_t tvae_gn_stats_finalize(const float* stats_part, int32_t slots_per_image, int32_t N, int32_t G, double count,
                               float eps, float* stats, tvae_stream_t stream);
/* x: the GroupNorm input, NHWC dense: fp32 (x_is_bf16 = 0), or bf16 (x_is_bf16 = 1; only for geometries of the
 * vectorised kernels: C/8 divides 256 and groups are whole 8-channel octets -- every layer of the reference model). */
int32_t tvae_gn_act_fwd(const void* x, int32_t x_is_bf16, const float* stats, const float* gamma, const float* beta,
                        int32_t N, int32_t HW, int32_t C, int32_t G, int32_t act, void* out_bf16, void* out_lo,
                        tvae_stream_t stream);
/* da: bf16 gradient wrt the activation output; gres (optional bf16) is added to dx (residual branch).
 * dgamma/dbeta are overwritten. dx_colsum (optional, float[C]): column sums of dx over all N*HW pixels, i.e. the bias
 * gradient of the conv whose output x is -- produced by the same pass that writes dx.
 * workspace: tvae_gn_bwd_workspace_bytes(N, HW, C, G). */
int64_t tvae_gn_bwd_workspace_bytes(int32_t N, int32_t HW, int32_t C, int32_t G);
int32_t tvae_gn_act_bwd(const void* x, int32_t x_is_bf16, const float* stats, const float* gamma, const float* beta,
                        const void* da_bf16, const void* gres_bf16, int32_t N, int32_t HW, int32_t C, int32_t G, int32_t act,
                        void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, float* workspace,
                        tvae_stream_t stream);

/* The same two calls with the activation derivative handed from forward to backward: tvae_gn_act_fwd2 also stores
 * act'(gamma * xhat + beta) as bf16 (act_grad, same shape as out; optional -- NULL = tvae_gn_act_fwd; needs act != 0, the
 * vectorised geometry (C / G a multiple of 8, C / 8 dividing 256) and no split-bf16 output); tvae_gn_act_bwd2 given that
 * tensor never evaluates the activation: dy = da * act_grad. Value and derivative of GELU come out of the same Phi / phi
 * evaluation, so the forward pays one FMA and 2 B per element for what costs the backward 16 instructions per element. */
int32_t tvae_gn_act_fwd2(const void* x, int32_t x_is_bf16, const float* stats, const float* gamma, const float* beta,
                         int32_t N, int32_t HW, int32_t C, int32_t G, int32_t act, void* out_bf16, void* out_lo,
                         void* act_grad_bf16, tvae_stream_t stream);
int32_t tvae_gn_act_bwd2(const void* x, int32_t x_is_bf16, const float* stats, const float* gamma, const float* beta,
                         const void* da_bf16, const void* gres_bf16, const void* act_grad_bf16, int32_t N, int32_t HW,
                         int32_t C, int32_t G, int32_t act, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum,
                         float* workspace, tvae_stream_t stream);

/* Scheduling switch of tvae_gn_act_bwd (results agree up to the summation order of the row sums): on = 0 (default) runs
 * the two-pass kernels (row sums, then apply: x and da are read from DRAM twice); on = 1 uses ONE persistent pass for
 * tensors that do not fit the L2 (cp.async.bulk ring; x and da are read from DRAM once, the second read is served by the
 * L2 group by group of `group_mb` megabytes: 38 % less DRAM traffic, but slower on B200 because it is issue-bound, see
 * DESIGN.md); on = 2 forces the single pass whatever the size (tests). group_mb <= 0 keeps the current group size (24). */
int32_t tvae_gn_set_bwd_fused(int32_t on, int32_t group_mb);

/* Column sums: out[c] = sum_rows x[row][c] (bias gradients). x bf16 [rows][pitch]; workspace rows_blocks*C floats:
 * tvae_colsum_workspace_bytes(rows, C). */
int64_t tvae_colsum_workspace_bytes(int64_t rows, int32_t C);
int32_t tvae_colsum_bf16(const void* x_bf16, int64_t rows, int32_t C, int32_t pitch, float* out, float* workspace,
                         tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Mid-block self-attention core. Replaces the einsum/softmax/einsum of AttnBlock.forward (src/model.py:128-139)
 * and its backward. Heads are channel-interleaved exactly as the reference's reshape(b, c_, n_heads, hw):
 * head h owns channels {d*n_heads + h}. q,k,v: fp32 [B*T][pitch] at column offsets (fused qkv GEMM output).
 * out: bf16 [B*T][C]; lse: fp32 [B][heads][T] (saved for backward). scale = (C/heads)^-0.5.
 */
int32_t tvae_attn_fwd(const float* q, const float* k, const float* v, int32_t pitch, int32_t B, int32_t T,
                      int32_t C, int32_t heads, void* out_bf16, float* out_f32, float* lse, tvae_stream_t stream);
/* d_out fp32 [B*T][C]; o fp32 [B*T][C] (forward output); dqkv: bf16 [B*T][3C] = (dq | dk | dv);
 * workspace: B*heads*T floats. */
int32_t tvae_attn_bwd(const float* q, const float* k, const float* v, int32_t pitch, const float* o,
                      const float* d_out, const float* lse, int32_t B, int32_t T, int32_t C, int32_t heads,
                      void* dqkv_bf16, float* workspace, tvae_stream_t stream);

/* Tensor-core (TF32 operands, fp32 accumulate) variants of the two calls above for head dimension 32 (C == 32*heads);
 * same arguments, same layouts, any T. ~5e-4 relative error on the logits; the exact kernels above stay the path
 * for other head sizes and for the fp32 mode. Default implementation: tcgen05.mma kind::tf32 with the scores,
 * probabilities and output accumulators in TMEM (attention_sm100.cu); tvae_attn_set_tcgen05(0) selects the
 * mma.sync.m16n8k8 kernels instead (returns the previous setting; a negative argument only queries). */
int32_t tvae_attn_set_tcgen05(int32_t enable);
int32_t tvae_attn_fwd_tc(const float* q, const float* k, const float* v, int32_t pitch, int32_t B, int32_t T,
                         int32_t C, int32_t heads, void* out_bf16, float* out_f32, float* lse, tvae_stream_t stream);
int32_t tvae_attn_bwd_tc(const float* q, const float* k, const float* v, int32_t pitch, const float* o,
                         const float* d_out, const float* lse, int32_t B, int32_t T, int32_t C, int32_t heads,
                         void* dqkv_bf16, float* workspace, tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Reparameterisation + KL. Replaces DiagonalGaussianDistribution.__init__/sample/kl (src/model.py:47-75).
 * moments fp32 NHWC [B*HW][2Z] = (mean | logvar); logvar is clamped to [-30, 20].
 * eps: NCHW fp32 [B][Z][HW] supplied by the caller (the reference draws it on the CPU, src/model.py:61-65),
 * or NULL to draw it with the Philox4x32-10 counter RNG keyed (seed, sample_offset + b, element).
 * z_bf16: NHWC [B*HW][z_pitch] (operand of post_quant_conv); z_nchw / eps_out (optional) fp32 NCHW;
 * kl[B] = 0.5 * sum(mean^2 + var - 1 - logvar) per sample.
 */
int32_t tvae_reparam_fwd(const float* moments, const float* eps, uint64_t seed, uint64_t sample_offset, int32_t B,
                         int32_t HW, int32_t Z, void* z_bf16, int32_t z_pitch, float* z_nchw, float* eps_out,
                         float* kl, void* z_lo, tvae_stream_t stream);
/* d_moments (bf16 NHWC [B*HW][2Z]) = d/d(moments) of  sum_i <dz_i, z_i> + kl_scale * sum_b kl[b],
 * for up to two samples z_i = mean + std*eps_i (dz_i: fp32 NHWC [B*HW][Z], eps_i: fp32 NCHW; pair 2 optional). */
int32_t tvae_reparam_bwd(const float* moments, const float* dz1, const float* eps1, const float* dz2,
                         const float* eps2, float kl_scale, int32_t B, int32_t HW, int32_t Z, void* dmoments_bf16,
                         tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Reconstruction NLL. Replaces F.l1_loss / F.mse_loss + the logvar scaling (src/model.py:656-663).
 *  x: bf16 NHWC [P][x_pitch]; xhat: fp32 NHWC [P][xh_pitch]; C valid channels; loss_type 0 = l1, 1 = l2.
 *  sums[3] (fp64) = { sum rec, sum (x - xhat)^2, unused }; written by the kernel (no pre-zeroing needed).
 *  dxhat (optional bf16 [P][dx_pitch]; pad lanes are never read by the consumers) = d(rec)/d(xhat) * grad_scale, where the caller passes
 *  grad_scale = exp(-logvar) / B (read from the device scalar `logvar`): dxhat = sign(xhat - x)*s or 2(xhat-x)*s.
 *  dx_colsum (optional fp32 [C], needs dxhat): column sums of dxhat over all P pixels = the bias gradient of the
 *  last decoder conv, produced by the same pass.
 *  workspace: tvae_nll_workspace_bytes(C).
 */
int64_t tvae_nll_workspace_bytes(int32_t C);
int32_t tvae_nll_fwd(const void* x_bf16, int32_t x_pitch, const float* xhat, int32_t xh_pitch, int64_t P, int32_t C,
                     int32_t loss_type, const float* logvar, int32_t batch, void* dxhat_bf16, int32_t dx_pitch,
                     float* dx_colsum, double* sums, double* workspace, tvae_stream_t stream);

/* Per-sample reconstruction metrics (src/scripts/evaluate_reconstruction.py:23-42): out[n] = (MAE, MSE) between the
 * bf16 input rows x and the fp32 reconstruction xhat (both channels-last, [N][HW][pitch]); PSNR follows from the MSE.
 * Fixed-order reductions. workspace: tvae_recon_metrics_workspace_bytes(N). */
int64_t tvae_recon_metrics_workspace_bytes(int32_t N);
int32_t tvae_recon_metrics(const void* x_bf16, int32_t x_pitch, const float* xhat, int32_t xh_pitch, int32_t N,
                           int32_t HW, int32_t C, float* out, double* workspace, tvae_stream_t stream);

/* Scalars of AutoencoderKL.get_loss (src/model.py:660-668) from the reductions above, on the device:
 *  out[0] = loss = nll + kl, out[1] = nll = (sums[0]*exp(-logvar) + logvar*n_elem)/B,
 *  out[2] = kl_weight * sum(kl)/B, out[3] = pixel_mse = sums[1]/n_elem, out[4] = d loss / d logvar. */
int32_t tvae_vae_loss_finalize(const double* sums, const float* kl, int32_t B, const float* logvar, double n_elem,
                               float kl_weight, float* out, tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * L2-product head loss. Replaces AvgPool2d(4) + isnan mask + masked-mean MSE (src/model_with_l2.py:151-168).
 *  pred: fp32 NHWC [B*hw][pred_pitch] (channel p = product p); target: fp32 [B][H][W] per product (NaN = invalid),
 *  H = 4*h, W = 4*w. out[p] = {sum sq err, valid count} (fp64 [nprod][2]); dpred (bf16 NHWC [B*hw][dp_pitch]) =
 *  weight[p] * 2 (pred - tgt) / count[p] on valid pixels, 0 elsewhere (computed by the _bwd call after the counts
 *  are known).
 */
int32_t tvae_l2head_loss_fwd(const float* pred, int32_t pred_pitch, const float* const* targets, int32_t nprod,
                             int32_t B, int32_t h, int32_t w, double* out, tvae_stream_t stream);
int32_t tvae_l2head_loss_bwd(const float* pred, int32_t pred_pitch, const float* const* targets, int32_t nprod,
                             int32_t B, int32_t h, int32_t w, const double* sums, const float* weights,
                             float grad_scale, void* dpred_bf16, int32_t dp_pitch, tvae_stream_t stream);

/* out[0] = vae_scal[0] + sum_p weights[p] * sums[p][0]/sums[p][1] over products with a valid pixel;
 * out[1+p] = that product's masked MSE (NaN when it has no valid pixel, i.e. skipped like the reference). */
int32_t tvae_l2head_finalize(const double* sums, const float* weights, int32_t nprod, const float* vae_scal,
                             float* out, tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser. Replaces clip_grad_norm_(max_norm) + torch.optim.AdamW.step (src/train_utils.py:175-177;
 * src/model.py:756-758). Flat fp32 buffers of n elements.
 *  tvae_sumsq: out[0] (fp64) = sum g^2 (deterministic two-stage); workspace tvae_sumsq_workspace_bytes(n).
 *  tvae_adamw: clip coefficient min(1, max_norm / (sqrt(sumsq) + 1e-6)) is computed on the device from `sumsq`
 *  (NULL = no clipping); decoupled weight decay; bias correction from `step` (1-based, after increment).
 */
int64_t tvae_sumsq_workspace_bytes(int64_t n);
int32_t tvae_sumsq(const float* g, int64_t n, double* out, double* workspace, tvae_stream_t stream);
int32_t tvae_adamw(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int64_t step, const double* sumsq,
                   float max_norm, float grad_scale, tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Data side: what feeds the path above.
 *
 * tvae_gather_rows: dst[j] = src[idx[j]] for n rows of row_bytes (multiple of 16) bytes; idx is an int64 DEVICE array
 *  with values in [0, n_src). The device-resident tile cache builds its batches with it (replaces the reference
 *  loader's shuffle-buffer draw + default_collate copy, src/tempo_data.py:34-110).
 *
 * tvae_extract_tiles: n_tiles square TxT tiles cut out of a raw radiance granule rad[M][NT][C] (fp32), each with the
 *  reference's random augmentation chain -- crop at (row0, col0), torch.flip(dims=[0]) if flags & 1,
 *  torch.flip(dims=[1]) if flags & 2, torch.rot90(k, dims=[0,1]) -- and, when mean/std are given, its normalisation
 *  z = clamp((log(max(rad, min_radiance)) - mean[c]) / (std[c] + 1e-8), clip_min, clip_max) applied on the fly
 *  (src/scripts/prepare_tempo_tiles.py:21-58 and :61-83 in ONE pass: the normalised granule is never materialised).
 *  spec: int32 DEVICE array [n_tiles][4] = {row0, col0, flags, k}. Outputs (either may be NULL): out_f32
 *  [n_tiles][T][T][C] (the on-disk tile format) and out_bf16 [n_tiles][T][T][out_pitch] (the conv operand rows, pad
 *  lanes zeroed). mean == std == NULL copies raw values (augmentation only).
 *
 * tvae_spectrum_stats_*: per-channel mean and population standard deviation of log(max(rad, min_radiance)) over all
 *  pixels of any number of granules (src/scripts/compute_tempo_stats.py:58-86: np.log(np.clip) -> mean / std over the
 *  stacked pixels). acc is a caller-zeroed fp64 DEVICE array [2][C] (running sum, sum of squares) that _accum adds
 *  `rows` pixels to (fixed-order reduction through the workspace); _finalize writes mean[C], std[C] (fp32).
 *  take_log = 0 accumulates the values as they are.
 *
 * tvae_batch_stats: out[4] = {min, max, mean, unbiased std} over the C valid channels of `rows` rows with row pitch
 *  `pitch` elements (fp32, or bf16 when is_bf16) -- the batch statistics Trainer.train_step prints at step 0
 *  (src/train_utils.py:156-159). workspace: tvae_batch_stats_workspace_bytes().
 */
int32_t tvae_gather_rows(const void* src, int64_t n_src, int64_t row_bytes, const int64_t* idx, int32_t n, void* dst,
                         tvae_stream_t stream);
int32_t tvae_extract_tiles(const float* rad, int32_t M, int32_t NT, int32_t C, const int32_t* spec, int32_t n_tiles,
                           int32_t T, const float* mean, const float* std, float min_radiance, float clip_min,
                           float clip_max, float* out_f32, void* out_bf16, int32_t out_pitch, tvae_stream_t stream);
int64_t tvae_spectrum_stats_workspace_bytes(int64_t rows, int32_t C);
int32_t tvae_spectrum_stats_accum(const float* rad, int64_t rows, int32_t C, float min_radiance, int32_t take_log,
                                  double* acc, double* workspace, tvae_stream_t stream);
int32_t tvae_spectrum_stats_finalize(const double* acc, int64_t total_rows, int32_t C, float* mean, float* std,
                                     tvae_stream_t stream);
int64_t tvae_batch_stats_workspace_bytes(void);
int32_t tvae_batch_stats(const void* x, int32_t is_bf16, int64_t rows, int64_t C, int64_t pitch, float* out,
                         double* workspace, tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Probe training on latents (src/scripts/linear_probe_analysis.py:212-353): nn.Linear layers run as
 * 1x1 convolutions through tvae_conv_gemm / tvae_wgrad_gemm and AdamW through tvae_adamw; these are the rest.
 *
 * tvae_act_dropout_fwd: out = dropout_p(act(x)) as bf16 operand rows. x fp32 [rows][x_pitch] (the Linear output, bias
 *  included); act: 0 identity, 1 GELU, 2 ReLU, 3 SiLU, 4 tanh (nn.ReLU / nn.GELU / nn.Tanh of MLPProbe); the keep mask
 *  of element (row, c) is Philox4x32-10 keyed by (seed, offset + row, c / 4): an element is zeroed with probability
 *  p_drop, kept ones are scaled by 1 / (1 - p_drop) (nn.Dropout). p_drop = 0 is plain activation (eval mode).
 *  Pad lanes [C, out_pitch) are zeroed.
 * tvae_act_dropout_bwd: dx = da * mask * act'(x) with the SAME (seed, offset): the mask is regenerated, not stored.
 * tvae_probe_mse: nn.MSELoss() on a [n, 1] prediction (target element r at target[r * target_pitch]): sums[0] = sum (pred - y)^2, sums[1] = sum y, sums[2] = sum y^2
 *  over the first n_valid rows (fp64, fixed order) -- loss = sums[0] / n_valid, R^2 = 1 - sums[0] / (sums[2] -
 *  sums[1]^2 / n); dpred (optional, bf16 [rows_padded][dp_pitch], column 0) = 2 (pred - y) / n_valid, 0 on pad rows.
 */
int32_t tvae_act_dropout_fwd(const float* x, int32_t x_pitch, int64_t rows, int32_t C, int32_t act, float p_drop,
                             uint64_t seed, uint64_t offset, void* out_bf16, int32_t out_pitch, tvae_stream_t stream);
int32_t tvae_act_dropout_bwd(const float* x, int32_t x_pitch, const void* da_bf16, int32_t da_pitch, int64_t rows,
                             int32_t C, int32_t act, float p_drop, uint64_t seed, uint64_t offset, void* dx_bf16,
                             int32_t dx_pitch, tvae_stream_t stream);
int32_t tvae_probe_mse(const float* pred, int32_t pred_pitch, const float* target, int64_t target_pitch, int64_t n_valid,
                       int64_t rows_padded, double* sums, void* dpred_bf16, int32_t dp_pitch, tvae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Probe TARGETS (the data side of the probes): the L2 component fields a probe regresses on are
 * normalised per component and averaged over 4x4 pixel blocks down to the latent grid
 * (src/scripts/linear_probe_analysis.py:60-110 normalize_component, :180-190 reshape + np.nanmean). NaN = invalid pixel.
 *
 * tvae_nan_moments: out5 (fp64) = count, sum (x - center), sum (x - center)^2, min x, max x over the non-NaN elements of
 *  x[0..n) (mean / std / min / max of the "zscore" and "minmax" statistics; a second call with center = mean gives the
 *  centred second moment). min / max are +inf / -inf when nothing is valid.
 * tvae_select_hist: one 8-bit pass of an exact radix select (np.median of the "asinh" statistics: median, then the
 *  median of |x - median|). Keys are the order-preserving uint32 image of v = x (use_abs = 0) or |x - center|
 *  (use_abs = 1), NaN skipped; hist256[b] = number of elements whose key equals `prefix` on the bits of `prefix_mask`
 *  and whose byte (key >> shift) & 255 is b. Four calls (shift 24, 16, 8, 0), the caller narrowing prefix between them,
 *  pin the k-th smallest value exactly. key(v) = ~bits(v) for negative v, bits(v) | 0x80000000 otherwise.
 * tvae_component_pool: f(x) per pixel, mode 0: (x - a) / b; mode 1: asinh(x / b); mode 2: logit(a + (1 - 2a) x); NaN stays
 *  NaN. `normalized` (optional, fp32 [H][W] dense) receives f; pooled[H / pool][W / pool] = mean of the valid f in each
 *  pool x pool block, NaN where a block has none. x is fp32 [H][pitch]. Rows / columns beyond the last whole block are
 *  dropped from `pooled` (the reference crops the field to a multiple of 64 first) and kept in `normalized`.
 */
int32_t tvae_nan_moments(const float* x, int64_t n, float center, double* out5, tvae_stream_t stream);
int32_t tvae_select_hist(const float* x, int64_t n, float center, int32_t use_abs, uint32_t prefix, uint32_t prefix_mask,
                         int32_t shift, uint64_t* hist256, tvae_stream_t stream);
int32_t tvae_component_pool(const float* x, int32_t H, int32_t W, int32_t pitch, int32_t pool, int32_t mode, float a,
                            float b, float* normalized, float* pooled, tvae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TVAE_H_ */
