/*
 * ovb200.h — C ABI of the B200 (sm_100a) hot-path library `libovb200.so`.
 *
 * The reference (SAIS-FUXI/Omni-Video) is pure Python/PyTorch and has no FFI of its own: every
 * FLOP on its denoising hot path goes through a library call made by PyTorch.
 * Each entry point below replaces one such call site (or a fused group of them); the comment above
 * each function names the reference file:line it stands in for.  The Python binding a maintainer
 * adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; pointers are DEVICE pointers unless stated otherwise;
 *   - no allocation, no synchronisation, no global state: the caller owns every buffer and passes
 *     the CUDA stream (a `cudaStream_t` passed as `void*`);
 *   - returns OVB_OK (0) or a negative error code; `ovb_last_error()` returns a thread-local,
 *     human-readable description of the last failure; nothing throws across the ABI;
 *   - bf16 tensors are `uint16_t`-sized IEEE bfloat16, row-major; "ld*" are leading dimensions in
 *     ELEMENTS; every bf16 row pitch must be a multiple of 8 elements (16 bytes) and every base
 *     pointer 16-byte aligned.
 */
#ifndef OVB200_H_
#define OVB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVB_OK 0
#define OVB_ERR_INVALID_ARGUMENT (-1)
#define OVB_ERR_CUDA (-2)
#define OVB_ERR_UNSUPPORTED_ARCH (-3)

#define OVB_ABI_VERSION 4

/* ---- library ---------------------------------------------------------------------------- */
int ovb_abi_version(void);
const char* ovb_last_error(void);
/* Fails with OVB_ERR_UNSUPPORTED_ARCH unless the current device is compute capability 10.x. */
int ovb_check_device(void);
/* Number of kernels this library has launched from the calling process (bench `gpu_launches`). */
int64_t ovb_launch_count(void);
/* Adds n to that counter: for callers that replay a captured CUDA graph of this library's launches. */
void ovb_add_launch_count(int64_t n);

/* ---- GEMM (tcgen05 / TMEM / TMA) ------------------------------------------------------------
 * out = epilogue( A[M,K] * W[N,K]^T + bias[N] )       bf16 operands, fp32 accumulation in TMEM.
 * Replaces every nn.Linear on the DiT path (reference omnivideo/modules/model.py:118-121 q/k/v/o,
 * :211-213 ffn, :384-388 patch/text embedding) together with the elementwise work the reference
 * runs right after it (model.py:245-246, :250, :253-254).
 *
 * epilogue:
 *   OVB_EPI_BF16            out(bf16)  = acc + bias
 *   OVB_EPI_GELU_BF16       out(bf16)  = gelu_tanh(bf16(acc + bias))               (model.py:212)
 *   OVB_EPI_RESID_F32       out(fp32)  = resid + bf16(acc + bias)                  (model.py:250)
 *   OVB_EPI_GATE_RESID_F32  out(fp32)  = resid + bf16(acc + bias) * gate           (model.py:246,254)
 *   OVB_EPI_F32             out(fp32)  = acc + bias
 *   OVB_EPI_RESID_BF16      out(bf16)  = resid(bf16) + bf16(acc + bias)            (vae2_1.py:258-261)
 *
 * gate is fp32; element (row r, col n) is read at
 *   gate[(r / rows_per_sample) * gate_sample_stride + (r % rows_per_sample) * gate_row_stride + n]
 * so a per-sample vector uses gate_row_stride = 0 and a per-token table uses gate_row_stride = ld.
 *
 * Sequence-parallel layouts, both optional:
 *   n_split  > 1 : the N output columns are cut into n_split equal groups; group g is written as its
 *                  own row-major [M, N/n_split] matrix at out + g * out_split_stride (elements),
 *                  i.e. straight into the Ulysses all-to-all send layout (N/n_split % 32 == 0).
 *   k_split  > 1 : A is given as k_split row-major [M, K/k_split] matrices, the g-th at
 *                  A + g * a_split_stride (elements): the all-to-all receive layout
 *                  (K/k_split % 64 == 0).
 */
#define OVB_EPI_BF16 0
#define OVB_EPI_GELU_BF16 1
#define OVB_EPI_RESID_F32 2
#define OVB_EPI_GATE_RESID_F32 3
#define OVB_EPI_F32 4
#define OVB_EPI_RESID_BF16 5

typedef struct ovb_gemm_args {
  const void* A;      /* bf16 [M, K] (or k_split x [M, K/k_split]) */
  int64_t lda;
  const void* W;      /* bf16 [N, K]  (nn.Linear weight layout [out, in]) */
  int64_t ldw;
  const void* bias;   /* bf16 [N] or NULL */
  void* out;          /* bf16 or fp32 [M, N] per epilogue */
  int64_t ldo;
  int32_t M, N, K;
  int32_t epilogue;
  const void* resid;  /* fp32 [M, N], ld = ldr (RESID / GATE_RESID), bf16 for RESID_BF16; may alias out */
  int64_t ldr;
  const void* gate;   /* fp32, see above (GATE_RESID) */
  int64_t gate_row_stride;
  int64_t gate_sample_stride;
  int32_t rows_per_sample;
  int32_t n_split;          /* <=1: off */
  int64_t out_split_stride;
  int32_t k_split;          /* <=1: off */
  int64_t a_split_stride;
} ovb_gemm_args;

int ovb_gemm_bf16(const ovb_gemm_args* args, void* stream);

/* ---- flash attention forward (tcgen05 / TMEM / TMA) ----------------------------------------
 * out[b, i, h, :] = softmax_j( scale * <q[b,i,h,:], k[b,j,h,:]> , j < k_lens[b] ) * v[b, j, h, :]
 * head_dim = 128, non-causal, bf16 in / bf16 out, fp32 softmax statistics.
 * Replaces flash_attention() (reference omnivideo/modules/attention.py:24-133 -> flash_attn
 * varlen kernels at :96/:116) for the two call sites model.py:144-149 (self) and :174 (cross).
 * q/k/v/out are [B, L, H, 128] views: element (b, l, h, c) at
 *   ptr + b * batch_stride + l * row_stride + h * 128 + c      (strides in elements)
 * which lets q, k and v live inside one fused [L, 3*H*128] projection buffer.
 * k_lens: HOST pointer to B int32 (valid keys per sample) or NULL (= Lk for all).
 */
typedef struct ovb_attn_args {
  const void* q; int64_t q_row_stride; int64_t q_batch_stride;
  const void* k; int64_t k_row_stride; int64_t k_batch_stride;
  const void* v; int64_t v_row_stride; int64_t v_batch_stride;
  void* out;     int64_t o_row_stride; int64_t o_batch_stride;
  int32_t B, H, Lq, Lk;
  const int32_t* k_lens;  /* host, may be NULL */
  float softmax_scale;    /* <=0: 1/sqrt(128) */
  /* Optional device scratch (16-byte aligned) of ovb_attn_workspace_size() bytes.  With it, the
   * last partial wave of (256-row q block, head, sample) work items is cut along the keys so that it
   * fills all SMs, and a combine pass merges the partial softmax results; without it (NULL) every
   * item is one CTA.  Results agree to bf16 rounding; no_split != 0 forces the one-CTA form. */
  void* workspace;
  int64_t workspace_bytes;
  int32_t no_split;
  /* Scatter form (B == 1): when out_ptrs is non-NULL it is a HOST array of n_out_ptrs device
   * pointers and query row r is written to out_ptrs[r / rows_per_out] + (r % rows_per_out) *
   * o_row_stride + h * 128: the rows of each sequence-parallel rank go straight into that rank's
   * (peer-mapped) buffer from the kernel's epilogue — the all-to-all after attention
   * (reference distributed/ulysses.py:54) with no separate collective. */
  void* const* out_ptrs; int32_t n_out_ptrs; int32_t rows_per_out;
} ovb_attn_args;

int ovb_attn_fwd(const ovb_attn_args* args, void* stream);
/* Upper bound of the scratch ovb_attn_fwd can use for these shapes (0: the launch has no partial wave). */
int64_t ovb_attn_workspace_size(const ovb_attn_args* args);
/* Host-only: the work decomposition ovb_attn_fwd would use for these args (shapes, k_lens, workspace,
 * no_split): n_whole items run as one CTA each, the n_tail items of the last partial wave are cut
 * `parts` ways along the keys (n_tail = 0, parts = 1: no split).  Launches nothing. */
int ovb_attn_plan(const ovb_attn_args* args, int32_t* n_whole, int32_t* n_tail, int32_t* parts);

/* ---- attention with a relative-position bias (umT5 encoder; warp-level mma.sync, bf16) --------------
 * out[b, i, h, :] = softmax_j( scale * <q[b,i,h,:], k[b,j,h,:]> + rel_bias[h, (j - i) + L - 1], j < k_lens[b] )
 *                   * v[b, j, h, :]
 * All heads and samples in ONE launch.  Replaces T5Attention.forward's score / bias / mask / softmax / PV chain
 * (reference omnivideo/modules/t5.py:96-117; umT5 uses scale = 1, t5.py:101) with the bias of
 * T5RelativeEmbedding (t5.py:229-263), which depends on (j - i) only: rel_bias is fp32 [H, >= 2L-1], head h at
 * rel_bias + h * bias_head_stride.  fp32 scores and softmax statistics, P rounded to bf16 before P V, fp32
 * accumulation.  q/k/v/out are [B, L, H, head_dim] views addressed as in ovb_attn_args (so they may be column
 * blocks of one fused [B*L, 3*H*head_dim] projection); head_dim 64 or 128.
 * k_lens: DEVICE pointer to B int32 (valid keys per sample, clamped to [1, L]) or NULL (= L): no host
 * synchronisation is needed to encode a ragged batch.  Rows i >= k_lens[b] are computed like the reference does
 * (they attend the valid keys); the caller drops them.
 */
typedef struct ovb_attn_relbias_args {
  const void* q; int64_t q_row_stride; int64_t q_batch_stride;
  const void* k; int64_t k_row_stride; int64_t k_batch_stride;
  const void* v; int64_t v_row_stride; int64_t v_batch_stride;
  void* out;     int64_t o_row_stride; int64_t o_batch_stride;
  const float* rel_bias; int64_t bias_head_stride;
  const int32_t* k_lens;  /* device, may be NULL */
  int32_t B, H, L, head_dim;
  float softmax_scale;
} ovb_attn_relbias_args;

int ovb_attn_relbias_fwd(const ovb_attn_relbias_args* args, void* stream);

/* ---- peer transport: cross-GPU signal / wait ---------------------------------------------------
 * Every rank owns a pad of >= world uint32 flags in memory that all ranks have mapped
 * (peer_pads[r] = rank r's pad as seen from this process; HOST array of `world` device pointers).
 * One launch: (1) system-scope fence, so that everything this rank's earlier kernels on `stream`
 * wrote to peer memory is visible; (2) store `epoch` into slot `rank` of every peer's pad;
 * (3) wait until all `world` slots of the own pad hold a value >= epoch.  Epochs must increase by one
 * per call on every rank: with epoch_counter (device uint32, owned by this rank) the kernel uses
 * *epoch_counter + 1 and stores it back, which keeps the launch valid inside a replayed CUDA graph;
 * with epoch_counter == NULL the `epoch` argument is used.  A wait that is still unsatisfied after
 * ~30 s (a peer died, or the epochs diverged) is FATAL: the kernel sets *status (device uint32, may
 * be NULL) to 1 and traps, so the stream never continues into kernels that would read half-written
 * receive buffers; the host sees a launch failure at its next synchronisation.  Ranks of a group
 * must therefore stay in lock-step. */
int ovb_peer_signal_wait(uint32_t* const* peer_pads, int32_t rank, int32_t world, uint32_t epoch,
                         uint32_t* epoch_counter, uint32_t* status, void* stream);
/* cudaDeviceEnablePeerAccess(peer_device) for the current device ("already enabled" is success). */
int ovb_enable_peer_access(int32_t peer_device);

/* ---- LayerNorm (+ AdaLN modulation or affine) ------------------------------------------------
 * y = LN_eps(x) * a + b over the last dim (fp32 statistics, two-pass like torch.layer_norm).
 * Replaces WanLayerNorm + modulation (reference model.py:87-97, :243, :252, :286-289) and the
 * affine norm3 (:205-207, :250).
 *   mode OVB_LN_MODULATE : a = 1 + scale, b = shift   (fp32 vectors/tables)
 *   mode OVB_LN_AFFINE   : a = weight,    b = bias    (bf16 [d])
 * scale/shift element (row r, col c):
 *   p[(r / rows_per_sample) * mod_sample_stride + (r % rows_per_sample) * mod_row_stride + c]
 * x is fp32 (x_is_bf16 = 0) or bf16; with round_ln_bf16 != 0 the normalised value is rounded to
 * bf16 before modulation (what `.type_as(x)` does for a bf16 residual stream, model.py:97).
 * out is bf16 (out_is_f32 = 0) or fp32.
 */
#define OVB_LN_MODULATE 0
#define OVB_LN_AFFINE 1

typedef struct ovb_ln_args {
  const void* x; int64_t ldx; int32_t x_is_bf16;
  void* out;     int64_t ldo; int32_t out_is_f32;
  int32_t rows, d;
  float eps;
  int32_t mode;
  const void* a;  /* MODULATE: fp32 scale ; AFFINE: bf16 weight */
  const void* b;  /* MODULATE: fp32 shift ; AFFINE: bf16 bias   */
  int64_t mod_row_stride, mod_sample_stride;
  int32_t rows_per_sample;
  int32_t round_ln_bf16;
} ovb_ln_args;

int ovb_layernorm(const ovb_ln_args* args, void* stream);

/* ---- RMSNorm over the full model dim (+ optional 3-D RoPE), in place ----------------------------
 * For every row: t = bf16( x * rsqrt(mean_d(x^2) + eps) ) * weight ; then, if rope != 0, each head's
 * 64 interleaved pairs (x[2j], x[2j+1]) are rotated by the angle of the token's (frame,row,col).
 * Replaces WanRMSNorm (reference model.py:68-84, :137-138, :169-170) and rope_apply (:37-65;
 * sequence-parallel variant distributed/full_sequence_parallel.py:22-60).
 * x: bf16 [rows, d] with leading dim ldx, optionally cut into col_split column groups of
 *    d/col_split columns, group g at x + g * col_split_stride (the all-to-all send layout).
 * cs_table: fp32 [table_len, 64, 2] = (cos, sin) of reference `freqs` (model.py:405-411).
 * Token index of row r (within sample) is token_offset + r % rows_per_sample; tokens >= F*H*W are
 * left unrotated (model.py:61; full_sequence_parallel.py:9-19).
 */
typedef struct ovb_rmsnorm_rope_args {
  void* x; int64_t ldx;
  int32_t rows, d;
  const void* weight;  /* bf16 [d] */
  float eps;
  int32_t rope;        /* 0: RMSNorm only */
  const float* cs_table; int32_t table_len;
  int32_t F, H, W;
  int32_t rows_per_sample;
  int32_t token_offset;
  int32_t col_split; int64_t col_split_stride;
  /* Scatter form (sequence parallelism over peer memory, see "peer transport" below): when dst_ptrs
   * is non-NULL it is a HOST array of col_split device pointers, and column group g of row r is
   * written to dst_ptrs[g] + r * dst_ld (+ column inside the group) instead of in place; dst_ptrs[g]
   * may be memory of another GPU mapped into this process.  copy_only != 0 skips the normalisation
   * (plain scatter of a column-split matrix, used for V). */
  void* const* dst_ptrs; int64_t dst_ld; int32_t copy_only;
} ovb_rmsnorm_rope_args;

int ovb_rmsnorm_rope(const ovb_rmsnorm_rope_args* args, void* stream);

/* ---- patchify / unpatchify ---------------------------------------------------------------------
 * patchify: latent x [C, F, H, W] (fp32 or bf16) -> tokens bf16 [F*(H/2)*(W/2), C*4]; column index
 * (c*2 + ph)*2 + pw matches patch_embedding.weight.flatten(1) (reference model.py:384, :455-458).
 * unpatchify: head output fp32 [L, 4*C] (column (ph*2+pw)*C + c) -> fp32 [C, F, 2*Hp, 2*Wp]
 * (reference model.py:526-549, einsum 'fhwpqrc->cfphqwr').
 */
int ovb_patchify(const void* x, int32_t x_is_bf16, void* tokens, int64_t ldt, int32_t C, int32_t F,
                 int32_t H, int32_t W, void* stream);
/* General (pt, ph, pw) patch gather for the conditioning front-end: x [C, F, H, W] (fp32 or bf16) ->
 * tokens bf16 [(F/pt)*(H/ph)*(W/pw), C*pt*ph*pw], column ((c*pt + it)*ph + ih)*pw + iw, i.e. the
 * im2col of VisualContextAdapter.patch_embedding (reference modules/visual_context_adapter.py:56-58,
 * :98); the Conv3d itself is then ovb_gemm_bf16 against weight.flatten(1). */
int ovb_patchify3d(const void* x, int32_t x_is_bf16, void* tokens, int64_t ldt, int32_t C, int32_t F,
                   int32_t H, int32_t W, int32_t pt, int32_t ph, int32_t pw, void* stream);
int ovb_unpatchify(const float* y, int64_t ldy, float* out, int32_t C, int32_t F, int32_t Hp,
                   int32_t Wp, void* stream);

/* ---- small fp32 linears ------------------------------------------------------------------------
 * out[M, N] (fp32) = act_out( act_in(x[M, K]) * W[N, K]^T + bias ), fp32 math on bf16-stored weights:
 * what the reference's `autocast('cuda', dtype=float32)` regions compute (time_embedding /
 * time_projection model.py:390-392, :469-476; head.head :285-289).
 * act: 0 none, 1 SiLU.   x fp32 [M, K] (ldx), W bf16 [N, K] (ldw), bias bf16 [N] or NULL.
 */
int ovb_linear_f32(const float* x, int64_t ldx, const void* W, int64_t ldw, const void* bias,
                   float* out, int64_t ldo, int32_t M, int32_t N, int32_t K, int32_t act_in,
                   int32_t act_out, void* stream);

/* Sinusoidal timestep embedding, fp64 math -> fp32 [n, dim]: cos(t*f_i) || sin(t*f_i),
 * f_i = 10000^(-i/(dim/2))   (reference model.py:13-23).  t: fp32 device [n]. */
int ovb_sinusoidal_embedding(const float* t, float* out, int32_t n, int32_t dim, void* stream);

/* AdaLN modulation tables, built once per forward:
 *   out[l, b, k, c] (fp32) = float(mod[l, k, c]) + e[b, e_has_k ? k*d + c : c]
 * mod: bf16 [n_layers, K, d] (stacked `blocks.{l}.modulation`, K = 6, reference model.py:216,238;
 * or `head.modulation`, K = 2, n_layers = 1, e_has_k = 0, :276,286); e: fp32 [B, K*d] or [B, d]. */
int ovb_modulation_table(const void* mod, const float* e, float* out, int32_t n_layers, int32_t B,
                         int32_t K, int32_t d, int32_t e_has_k, void* stream);

/* fp32 -> bf16 / bf16 -> fp32 casts of n contiguous elements (n % 8 == 0). */
int ovb_cast_f32_to_bf16(const float* x, void* out, int64_t n, void* stream);
int ovb_cast_bf16_to_f32(const void* x, float* out, int64_t n, void* stream);

/* ---- sampler-side linear combinations ------------------------------------------------------------
 * out[i] = sum_{k < n_terms} c[k] * x[k][i]      fp32 math; each tensor fp32 or bf16; n % 4 == 0.
 * One launch for each tensor expression of a denoising step outside the DiT: classifier-free guidance
 * (reference x2x_gen_unified.py:484), x0 prediction (utils/fm_solvers_unipc.py:319-323), UniPC corrector
 * (:610-623) and predictor (:467-475).  Bit k of round_mask rounds term k to bf16 before the sum
 * (torch's promotion rule for `0-dim fp32 tensor * bf16 tensor`, which the reference hits when the
 * latents are bf16). */
typedef struct ovb_lincomb_args {
  const void* x[5];
  float c[5];
  int32_t is_bf16[5];
  int32_t n_terms;
  int32_t round_mask;
  void* out;
  int32_t out_is_bf16;
  int64_t n;
} ovb_lincomb_args;

int ovb_lincomb(const ovb_lincomb_args* args, void* stream);

/* ==== Wan2.1 causal-3D VAE (reference omnivideo/modules/vae2_1.py) ================================
 * Activations are channels-last bf16 [T, H, W, C] (C a multiple of 16); all frames of a clip are
 * processed in one pass (the reference's chunk loop + feat_cache, vae2_1.py:515-567, is causal zero
 * padding: DESIGN.md §6).
 *
 * ---- implicit-GEMM convolution (tcgen05 / TMEM / TMA) ------------------------------------------
 * out[t,h,w,n] = bias[n] + sum_{tap,c} x[t*st + dt[tap], h*sh + dh[tap], w*sw + dw[tap], c]
 *                                       * w[n, tap*Cin + c]    (+ resid[t,h,w,n])
 * Reads outside x contribute zero (TMA zero fill) — that is the zero padding of CausalConv3d
 * (vae2_1.py:16-35: W/H symmetric, T on the left only), of nn.Conv2d(padding=1) (:78,:82) and of
 * ZeroPad2d((0,1,0,1)) (:88,:92).  Replaces every cuDNN conv of the VAE.
 * Output element (t,h,w,n), OVB_CONV_OUT_BF16:
 *     out[out_base + t*os_t + h*os_h + w*os_w + (n / n_group)*group_stride + n % n_group]
 *   (n_group = 0: no grouping).  Doubled os_h/os_w + a phase offset in out_base writes one phase of a
 *   2x nearest-upsampled image (Upsample + Conv2d, :75-84, as four 2x2 convs); n_group = C with
 *   group_stride = os_t interleaves time_conv's two channel halves as consecutive frames (:133-136).
 *   n_group a multiple of 16, or (w_rows % 32 == 0, no resid, no norm) a multiple of 4 below 16: per-group
 *   planes such as [group][T][H][W][4], the input of ovb_tap_sum; strides then need 4-element alignment.
 * OVB_CONV_OUT_F32_PLANAR(_CLAMP): fp32 out[n*plane_stride + out_base + t*os_t + h*os_h + w*os_w],
 *   optionally clamped to [-1,1] (Wan2_1_VAE.decode, :660).
 */
#define OVB_CONV_OUT_BF16 0
#define OVB_CONV_OUT_F32_PLANAR 1
#define OVB_CONV_OUT_F32_PLANAR_CLAMP 2

typedef struct ovb_conv3d_args {
  const void* x;          /* bf16 [Ti, Hi, Wi, Cin] */
  int32_t Ti, Hi, Wi, Cin;
  const void* w;          /* bf16 [w_rows, n_taps*Cin], rows >= Cout are zero */
  int32_t w_rows;
  const float* bias;      /* fp32 [Cout] or NULL */
  int32_t Cout;
  int32_t n_taps;         /* <= 27 */
  const int8_t* tap_dt;   /* HOST arrays [n_taps] */
  const int8_t* tap_dh;
  const int8_t* tap_dw;
  int32_t st, sh, sw;     /* input step per output step (1 or 2) */
  int32_t To, Ho, Wo;     /* output extents */
  void* out;
  int64_t os_t, os_h, os_w, out_base;   /* in elements */
  int32_t n_group;
  int64_t group_stride;
  const void* resid;      /* bf16, addressed like out, or NULL */
  int32_t out_mode;
  int64_t plane_stride;
  /* Fused RMS_norm (+SiLU) over the channels of every output voxel (reference vae2_1.py:38-53 followed by
   * :194-196, the op that consumes nearly every conv output): with norm_gamma (fp32 [Cout], 16-byte
   * aligned) set, norm_out — addressed like out — receives
   *     act( v / max(||v||_2, 1e-12) * sqrt(Cout) * gamma ),   v = the bf16 value `out` gets,
   * in addition to `out`; norm_out == NULL: only the normalised tensor is written, into `out`.
   * Needs OVB_CONV_OUT_BF16, n_group == 0, Cout == w_rows, Cout % 32 == 0 and Cout <= 256 with
   * 256 % ... i.e. all channels in ONE N tile (Cout in {32, 64, 96, 128, 192, 256}). */
  const float* norm_gamma;
  void* norm_out;
  int32_t norm_silu;
} ovb_conv3d_args;

int ovb_conv3d(const ovb_conv3d_args* args, void* stream);

/* Sum of per-tap partial products: the second half of a convolution with very few output channels
 * (the decoder head, CausalConv3d 96 -> 3, vae2_1.py:164-165 / :469).  As an implicit GEMM such a conv pays a
 * full A-operand fetch per MMA for 3 useful columns; instead ovb_conv3d computes, as ONE 1x1x1 conv with
 * n_taps*4 output columns and n_group = 4, the planes  y[k][t][h][w][c] = sum_ci W[c][ci][tap k] * x[t][h][w][ci]
 * (bf16), and this kernel shifts and adds them:
 *   out[c][t][h][w] = bias[c] + sum_k y[k][t + dt_k][h + dh_k][w + dw_k][c]      (fp32, planar; clamp to [-1, 1])
 * with terms outside y's extents equal to zero (causal and spatial zero padding).  HBM-bound:
 * 8 B x n_taps read + 4 B x C written per output voxel.  y: bf16 [n_taps][Ty][Hy][Wy][4], 8-byte aligned. */
typedef struct {
  const void* y;
  int32_t n_taps;
  const int8_t *tap_dt, *tap_dh, *tap_dw;
  int32_t Ty, Hy, Wy;
  const float* bias; /* [C] or NULL */
  float* out;        /* fp32 [C][To][Ho][Wo] */
  int32_t C;         /* 1..4 */
  int32_t To, Ho, Wo;
  int32_t clamp;
} ovb_tap_sum_args;

int ovb_tap_sum(const ovb_tap_sum_args* args, void* stream);

/* RMS_norm over channels (+ optional SiLU) on channels-last voxels (vae2_1.py:38-53, :194-196):
 * out[v, c] = act( x[v, c] / max(||x[v, :]||_2, 1e-12) * sqrt(C) * gamma[c] ).  HBM-bound: 4 B/elem. */
int ovb_rmsnorm_silu_cl(const void* x, const float* gamma, void* out, int64_t n_vox, int32_t C,
                        int32_t silu, void* stream);

/* Layout converters with a per-channel affine (latent de/normalisation, vae2_1.py:535-539,:546-550):
 * planar_to_cl: out[i, c] (bf16, c < Cpad) = c < C ? src[c*n + i]*scale[c] + shift[c] : 0
 * cl_to_planar: out[c*n + i] (fp32, c < C)  = float(src[i*Cs + c])*scale[c] + shift[c]      */
int ovb_planar_to_cl(const float* src, const float* scale, const float* shift, void* out, int32_t C,
                     int32_t Cpad, int64_t n, void* stream);
int ovb_cl_to_planar(const void* src, int32_t Cs, const float* scale, const float* shift, float* out,
                     int32_t C, int64_t n, void* stream);

/* Row softmax for the VAE's single-head attention (vae2_1.py:251): P[r, :] = softmax(scale*S[r, :]),
 * S fp32 [rows, cols] (ld lds) -> P bf16 (ld ldp). */
int ovb_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, int32_t rows, int32_t cols,
                     float scale, void* stream);

/* ---- output video path ---------------------------------------------------------------------------
 * Decoded video -> uint8 frames, the tensor part of `cache_video` (reference omnivideo/utils/utils.py:177-185:
 * clamp to value_range, normalise to [0, 1], x 255, truncate):
 *   out[t * out_frame_stride + y * out_row_stride + x * C + c] = uint8((clamp(v, lo, hi) - lo) / max(hi - lo, 1e-5) * 255)
 * x: planar fp32 [C, T, H, W] (what Wan2_1_VAE.decode returns), C <= 4; out: uint8, pitches in bytes, so one
 * video can be written into a cell of a torchvision-style grid. */
int ovb_video_to_uint8(const float* x, uint8_t* out, int32_t C, int32_t T, int32_t H, int32_t W,
                       int64_t out_frame_stride, int64_t out_row_stride, float lo, float hi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OVB200_H_ */
