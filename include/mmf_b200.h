/*
 * mmf_b200.h -- C ABI of the B200-native MultiMAE hot path (libmmf_b200.so).
 *
 * The reference (Yusin2Chen/incomplete_multimodal_fusion) has NO FFI/plugin boundary on this path:
 * it is plain PyTorch eager.  Each entry point below therefore names the
 * group of reference ATen call sites it replaces (file:line under /root/reference/pretraining/multimae).
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types, no allocation, no exceptions, no global mutable state
 *     (except a lazily resolved driver entry point for cuTensorMapEncodeTiled);
 *   - all pointers are DEVICE pointers unless the name says `_host`; row-major; `ld*` in ELEMENTS;
 *   - every call enqueues on `stream` and returns immediately;
 *   - return value: 0 = ok, < 0 = argument error (-(line-ish code)), > 0 = cudaError_t / CUresult.
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t), f32 = IEEE float.
 */
#ifndef MMF_B200_H_
#define MMF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mmf_stream_t; /* == cudaStream_t */

/* ------------------------------------------------------------------------------------------------
 * Library info
 * ---------------------------------------------------------------------------------------------- */
/* ABI version of this header; bumps when a struct layout or signature changes. */
int mmf_abi_version(void);
/* Number of kernel launches issued through this library since load / last reset (for bench.py's
 * `gpu_launches`). */
int64_t mmf_launch_count(void);
void mmf_reset_launch_count(void);
/* Data-parallel runs: leave `n` SMs out of the persistent GEMM grids so that a concurrent NCCL all-reduce (capped to the
 * same number of CTAs, NCCL_MAX_CTAS) does not delay GEMM CTAs that are statically assigned to the SMs it occupies. */
void mmf_set_gemm_reserved_sms(int32_t n);

/* ------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators), bf16 x bf16 -> fp32 accumulate.
 *   out[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )
 * Replaces every nn.Linear / Conv2d(k=s=P) / einsum->bmm call on the path and their autograd
 * backward: zorro_utils.py:125,127,166-168,179,194 (encoder), input_adapters.py:110 (patch
 * projection as im2col GEMM), multimae_utils.py:143-153,164-181 (decoder), output_adapters_simple.py
 * :168,180.  dgrad uses b_mn=1 (B = W read "transposed"), wgrad uses a_mn=b_mn=1 with split_k.
 * ---------------------------------------------------------------------------------------------- */
typedef struct MmfGemmArgs {
  const void* a;           /* bf16. a_mn=0: [M,K] (lda>=K).  a_mn=1: stored [K,M] (lda>=M) */
  const void* b;           /* bf16. b_mn=0: [N,K] (ldb>=K).  b_mn=1: stored [K,N] (ldb>=N) */
  void* out;               /* bf16 or f32 [M,N] (ldo) -- see out_period for the row mapping */
  const float* bias;       /* f32 [N] or NULL: added before the activation */
  const float* residual;   /* f32 rows of length >= N (ldr) or NULL: added after the activation */
  const float* residual2;  /* optional second residual source: rows >= res_split come from residual2[row - res_split] */
  int64_t res_split;
  const int32_t* res_row_map; /* int32 [res_period] or NULL */
  int64_t M, N, K;
  int64_t lda, ldb, ldo, ldr;
  int32_t a_mn, b_mn;
  int32_t out_f32;         /* 0: bf16 output, 1: f32 output */
  int32_t act;             /* 0: none, 1: exact-erf GELU, 2: GEGLU, 3: GEGLU backward (see below) */
  int32_t split_k;         /* >=1.  >1: f32 atomic accumulation into a ZEROED `out`; requires out_f32=1,
                              act=0, bias=residual=NULL */
  int32_t res_period;      /* 0: residual row = r.  >0: residual row = map ? map[r % p] : r % p */
  int32_t out_period;      /* 0: out row = r.  >0: out row = (r / p) * out_batch_rows + r % p */
  int32_t out_batch_rows;
  int32_t block_n;         /* 0: auto, or 128 / 256 */
  float alpha;             /* scale applied to the accumulator first */
  void* out2;              /* optional second output (bf16, same row mapping, ldo2): a bf16 copy of an f32
                              `out`; or, with act=1 and a bf16 `out`, the PRE-activation (for the backward) */
  int64_t ldo2;
  int32_t accumulate;      /* 1: out += result (f32: atomic adds; bf16: read-modify-write) */
  /* act=2 (GEGLU, zorro_utils.py:115-118): B is the [2*I_pad, K] weight; tile columns pair value
   * row j with gate row I_pad + j; out[M, I_pad] = gelu(gate) * value; `out2` (optional, bf16,
   * [M, 2*I_pad]) receives the pre-activation for the backward.  N must be passed as I_pad.
   * act=3 (backward of the same GEGLU fused into the dgrad GEMM of the FFN's second Linear, autograd of
   * zorro_utils.py:115-128): the accumulator is dg = alpha * A . B^T ([M, I_pad], never written); `out2` is an
   * INPUT, the saved pre-activation [value | gate] (bf16 [M, 2*I_pad]); out (bf16 [M, 2*I_pad]) =
   * [dg * gelu(gate) | dg * value * gelu'(gate)].  N = I_pad must be a multiple of 64; bf16, no bias/residual. */
} MmfGemmArgs;
int mmf_gemm_bf16(const MmfGemmArgs* args, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm, single or fused double (fp32 math, warp per row).
 *   y = LN2( LN1(x) )   LN1: gamma g1 (+ optional bias b1, eps1);  LN2 (optional, g2 != NULL): gamma g2, eps2
 * Replaces zorro_utils.py:103-110 applied twice back to back (:238->:176, :239->:124), the final
 * encoder norm multimae.py:431 and the decoder nn.LayerNorm(eps=1e-6) (multimae_utils.py:217-232).
 * x rows may come from two buffers: rows [0, x_split) from x, rows >= x_split from x2 (row - x_split)
 * (x2 == NULL: single source).  stats: f32 [rows, 4] = mean1, rstd1, mean2, rstd2 (nullable in fwd).
 * Fused residual add (forward, delta != NULL): rows >= delta_row0 are normalised as x + delta[row - delta_row0] with
 * delta the bf16 output of the previous sub-layer's last Linear (`x + attn(...)`, `x + ffn(...)`: zorro_utils.py:238-239;
 * under autocast that Linear output is bf16 and the sum fp32, Appendix A #17), and the sum -- the new fp32 residual
 * stream -- is written to xout[row - delta_row0].
 * The backward also adds the residual-branch gradient `dres` and can emit a bf16 copy of dx.
 * dg1/db1/dg2 are ACCUMULATED with atomics: the caller zeroes them.
 * ---------------------------------------------------------------------------------------------- */
int mmf_layernorm_fwd(const float* x, const float* x2, int64_t x_split, int64_t rows, int32_t D, int64_t ldx,
                      const float* g1, const float* b1, float eps1, const float* g2, float eps2, void* y, int64_t ldy,
                      int32_t y_f32, float* stats, const void* delta, int64_t delta_row0, int64_t lddelta, float* xout,
                      int64_t ldxout, mmf_stream_t stream);
int mmf_layernorm_bwd(const void* dy, int64_t lddy, int32_t dy_f32, const float* x, const float* x2, int64_t x_split,
                      int64_t rows, int32_t D, int64_t ldx, const float* g1, const float* b1, const float* g2,
                      const float* stats, const float* dres, int64_t lddres, float* dx, int64_t lddx, void* dx_bf16,
                      int64_t lddxb, float* dg1, float* db1, float* dg2, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Zorro-masked flash attention (zorro_utils.py:181-193; decoder: multimae_utils.py:170-180).
 * q/k/v/o are bf16 token-major matrices; head h occupies columns [h*dh, (h+1)*dh) of a row.
 * Token (b, i) is row  i < n_head ? b*n_head + i : B*n_head + b*n_tail + (i - n_head).
 * seg (device int32[nseg+1], or NULL): tokens [seg[s], seg[s+1]) form segment s; a segment attends
 * to itself, the LAST segment (fusion tokens) attends to everything (multimae.py:410-426).  With seg,
 * Nq == Nk.  The mask is realised by skipping key blocks, not by adding -inf.
 * ---------------------------------------------------------------------------------------------- */
typedef struct MmfAttnArgs {
  const void* q; const void* k; const void* v;
  void* o;                 /* bf16 [rows, H*dh] */
  float* lse;              /* f32 [B, H, Nq] (fwd: optional output; bwd: input) */
  int64_t ldq, ldk, ldv, ldo;
  int32_t B, H, Nq, Nk, dh;          /* dh in {32, 64} */
  int32_t n_head_q, n_tail_q, n_head_k, n_tail_k;
  float scale;
  const int32_t* seg;
  int32_t nseg;
  /* backward only */
  const void* d_o; int64_t lddo;
  float* delta;            /* f32 [B, H, Nq] scratch */
  void* dq; void* dk; void* dv;
  int64_t lddq, lddk, lddv;
} MmfAttnArgs;
int mmf_attn_fwd(const MmfAttnArgs* args, mmf_stream_t stream);
int mmf_attn_bwd(const MmfAttnArgs* args, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Slot ("modality") attention of Block_Fusion (downstream/.../zorro_utils.py:243-258, call site
 * multimae_crossattn.py:450-470): per position p the fusion token's query attends to S slots:
 * slot s < S-1 = modality s's visible token at p (row b*n_head + seg[s] + slotmap[s*F+p]) or, when
 * slotmap is -1, the batch-invariant mask-embedding row p; slot S-1 = the fusion token itself.
 * kv rows are [k | v] with H*dh columns each.  Only the fusion slot's output is produced.
 * ---------------------------------------------------------------------------------------------- */
typedef struct MmfSlotAttnArgs {
  const void* q;          /* bf16 [B*F, H*dh] */
  const void* kv_tok;     /* bf16 planar token rows [B*n_head + B*F, 2*H*dh] */
  const void* kv_me;      /* bf16 [F, 2*H*dh] */
  const int32_t* slotmap; /* int32 [S-1, F] */
  const int32_t* seg;     /* int32 [>= S-1] segment starts inside the head plane */
  void* out;              /* bf16 [B*F, H*dh] */
  float* probs;           /* f32 [B*F, H, S] or NULL */
  int64_t ldq, ldkv, ldme, ldo;
  int32_t B, F, H, S, dh, n_head;   /* dh must be 64 */
  float scale;
  /* backward */
  const void* dout; void* dq; void* dkv_tok; float* dkv_me; /* dkv_me f32 [F, 2*H*dh], caller zeroes */
  int64_t lddout, lddq, lddkv, lddme;
  float* me_scratch;      /* optional f32 [B*F, H, S-1, 2]: enables the atomic-free backward (masks are shared by the
                             batch, so the mask-embedding gradients are a plain reduction over the samples) */
} MmfSlotAttnArgs;
int mmf_slot_attn_fwd(const MmfSlotAttnArgs* args, mmf_stream_t stream);
int mmf_slot_attn_bwd(const MmfSlotAttnArgs* args, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pool attention: R learned queries over the N tokens of each sample with a dense boolean mask
 * (multimae.py:434-455; per-modality pools multimae_crossattn.py:529-543).  masked_fill(-finfo.max)
 * semantics: a row with no allowed key is UNIFORM over all N keys (mode[r] = 0) -- or, for an empty
 * context (mode[r] = 1), zero.  stat: f32 [B,R,H,2] (max, sum) followed by [B,R,H] scratch.
 * ---------------------------------------------------------------------------------------------- */
typedef struct MmfPoolAttnArgs {
  const void* q;          /* bf16 [Bq, R, H*dh]; q_bstride = 0 if batch-invariant */
  const void* kv;         /* bf16 planar token rows [.., 2*H*dh] */
  const uint8_t* mask;    /* [R, N] */
  const int32_t* mode;    /* [R] */
  void* out;              /* bf16 [B, R, H*dh] */
  float* stat;            /* f32 [B*R*H*3] */
  int64_t q_bstride, ldkv;
  int32_t B, R, H, N, dh, n_head, n_tail;
  float scale;
  /* backward */
  const void* dout;       /* bf16 [B, R, H*dh] */
  float* dq;              /* f32 [Bq, R, H*dh], accumulated (caller zeroes) */
  int64_t dq_bstride;
  void* dkv;              /* bf16 planar token rows [.., 2*H*dh], every row written */
  int64_t lddkv;
} MmfPoolAttnArgs;
int mmf_pool_attn_fwd(const MmfPoolAttnArgs* args, mmf_stream_t stream);
int mmf_pool_attn_bwd(const MmfPoolAttnArgs* args, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Masked reconstruction losses (criterion.py:85-115 MSE, :142-172 L1; norm_pix=False):
 *   per pixel err averaged over channels, weighted by the nearest-upsampled patch mask, summed per
 *   sample / (P*P*#masked patches), nanmean over the batch.  mask: int64 [B, F] (1 = masked) or NULL
 *   (plain mean).  `work`: f32 [2*B + 2] scratch (zeroed by the call).  loss: f32 [1].
 *   The backward writes dpred = dloss * d(loss)/d(pred) (bf16 or f32 like pred).
 * ---------------------------------------------------------------------------------------------- */
int mmf_masked_loss_fwd(const void* pred, int32_t pred_f32, const float* target, const int64_t* mask, int64_t mask_bstride,
                        int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, int32_t kind /*0 mse, 1 l1*/, float* work,
                        float* loss, mmf_stream_t stream);
int mmf_masked_loss_bwd(const void* pred, int32_t pred_f32, const float* target, const int64_t* mask, int64_t mask_bstride,
                        int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, int32_t kind, const float* work,
                        const float* dloss, void* dpred, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Patch im2col of ALL modalities through a device token table (input_adapters.py:110 for every adapter + the token
 * selection of multimae.py:378-383, with no per-modality count on the host): row b * nenc + i of `out` [B * nenc, ld_out]
 * bf16 belongs to token tok[i] = global id; with m the modality that id falls into (tok_off[m] <= id < tok_off[m+1]) and
 * patch = id - tok_off[m], columns [col_off[m], col_off[m] + C_m * P * P) hold that patch of imgs[m] ([B, C_m, H, W] f32,
 * (c, ph, pw) order), column ind_col + m holds 1.0 and every other column 0.  One GEMM of `out` against the column-wise
 * concatenation of the modalities' projection weights then embeds all visible tokens at once, and the weight-gradient
 * GEMM's columns ind_col + m are the bias gradients.  M <= 4 modalities; all arrays below are HOST arrays.
 * ---------------------------------------------------------------------------------------------- */
int mmf_im2col_tokens(const float* const* imgs, const int32_t* chans, const int32_t* col_off, const int32_t* tok_off, int32_t M,
                      const int32_t* tok, int32_t nenc, void* out, int64_t ld_out, int32_t ind_col, int64_t batch, int32_t H,
                      int32_t W, int32_t P, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise / data movement (all HBM-bound)
 * ---------------------------------------------------------------------------------------------- */
/* f32 [rows, cols] -> bf16 [rows_pad, cols_pad] with zero padding and a scale: the per-step bf16
 * weight images (what autocast's weight cast does in the reference, pretrain_mmae.py:466). */
int mmf_cast_f32_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst, int64_t rows_pad,
                      int64_t cols_pad, int64_t ld_dst, float scale, mmf_stream_t stream);
/* GEGLU backward (zorro_utils.py:115-118): u bf16 [rows, 2*ipad] = [value | gate], dg bf16 [rows, ipad] */
int mmf_geglu_bwd(const void* u, const void* dg, void* du, int64_t rows, int64_t ipad, mmf_stream_t stream);
/* dpre = dy * gelu'(pre), bf16, n % 8 == 0 (decoder / pooling Mlp, multimae_utils.py:138-155) */
int mmf_gelu_bwd(const void* pre, const void* dy, void* dpre, int64_t n, mmf_stream_t stream);
/* out[c] += sum_r x[r, c]  (bias gradients); out f32, caller zeroes */
int mmf_colsum(const void* x, int32_t x_f32, int64_t rows, int64_t cols, int64_t ld, float* out, mmf_stream_t stream);
/* dst[b, r, :] = src[r, :] (fusion tokens + pos-emb broadcast, multimae.py:353-354) and its gradient */
int mmf_bcast_rows(const float* src, float* dst, int64_t batch, int64_t rows, int64_t d, int64_t dst_batch_stride,
                   mmf_stream_t stream);
int mmf_reduce_batch(const float* src, float* dst, int64_t batch, int64_t rows, int64_t d, int64_t src_batch_stride,
                     mmf_stream_t stream);
/* im2col of the VISIBLE patches only (input_adapters.py:110 + multimae.py:378-383 fused):
 * out[b*n_keep + i, (c ph pw)] = bf16(img[b, c, py*P+ph, px*P+pw]), patch idx[i] = py*(W/P)+px */
int mmf_im2col_gather(const float* img, const int32_t* idx, void* out, int64_t batch, int32_t C, int32_t H, int32_t W,
                      int32_t P, int32_t n_keep, int64_t ld_out, mmf_stream_t stream);
/* one-hot im2col of a class map for the VISIBLE patches (SemSegInputAdapter.forward, input_adapters.py:299-328, the
 * 4th `dnw` modality of multimae_quadruplet.py): out[b*n_keep + i, cls*P*P + ph*P + pw] = 1 for every pixel of patch
 * idx[i]; `out` (bf16, ld_out >= num_classes*P*P) must be zeroed by the caller; cls is [B, H, W] int64 */
int mmf_onehot_im2col(const int64_t* cls, const int32_t* idx, void* out, int64_t batch, int32_t H, int32_t W, int32_t P,
                      int32_t n_keep, int32_t num_classes, int64_t ld_out, mmf_stream_t stream);
/* Device half of the input pipeline (reference utils/multimodal_dfc2023.py:99-141 load_dsm / load_rgb /
 * load_sar after the rasterio decode, and the RandomCrop of :53-94), one launch per modality:
 *   src [B, C, Hs, Ws] raw raster, src_dtype 0 = uint8, 1 = uint16, 2 = float32, aligned to two elements;
 *   mode 0: nan_to_num -> cv2.resize(INTER_AREA) by the integer `factor` -> (float64(x) - mean[c]) / std[c] -> fp32  (load_rgb)
 *   mode 1: 10 log10(x + 1e-7), clip [-25, 0], nan_to_num -> resize -> the same z-score (load_sar; float32 rasters only)
 *   mode 2: nan_to_num -> resize -> (x - mean) / sqrt(var + 1e-6) with the resized image's own fp32 mean / variance (load_dsm)
 *   then the window [top[b], top[b] + Ho) x [left[b], left[b] + Wo) of the RESIZED image -> out [B, C, Ho, Wo] fp32.
 * mean_host / std_host: HOST arrays of C doubles (modes 0, 1); crop_top / crop_left: DEVICE int32 [B] or both null (then
 * Ho = Hs / factor, Wo = Ws / factor).  The caller guarantees top[b] + Ho <= Hs / factor (same for left). */
int mmf_raster_prep(const void* src, int32_t src_dtype, int64_t batch, int32_t C, int32_t Hs, int32_t Ws, int32_t factor,
                    int32_t mode, const double* mean_host, const double* std_host, const int32_t* crop_top,
                    const int32_t* crop_left, int32_t Ho, int32_t Wo, float* out, mmf_stream_t stream);
/* Truncated depth standardisation (pretrain_mmae.py:452-459, --standardize_depth): per sample of n = C*H*W fp32 values,
 * mean and UNBIASED variance of sorted(x)[k_lo : k_hi] (the reference: k_lo = int(0.1 n), k_hi = int(0.9 n)), then
 * out = (x - mean) / sqrt(var + 1e-6) over the whole sample.  Radix select in shared memory, no sort; n <= 56,320. */
int mmf_trunc_standardize(const float* x, float* out, int64_t batch, int32_t n, int32_t k_lo, int32_t k_hi,
                          mmf_stream_t stream);
/* 'b (nh nw) (c ph pw) -> b c (nh ph) (nw pw)' bf16 (output_adapters_simple.py:183-186); inverse=1 for the gradient */
int mmf_unpatchify_bf16(void* tokens, void* image, int64_t batch, int32_t C, int32_t H, int32_t W, int32_t P,
                        int32_t inverse, mmf_stream_t stream);
/* dst[b*n + i, :] = cast(src[b*src_batch_rows + row_off + (idx ? idx[i] : i), :]) */
int mmf_gather_rows(const void* src, int32_t src_f32, int64_t ld_src, int64_t src_batch_rows, int64_t row_off,
                    const int32_t* idx, void* dst, int32_t dst_f32, int64_t ld_dst, int64_t batch, int32_t n, int32_t d,
                    mmf_stream_t stream);
int mmf_add_inplace_f32(float* y, const float* x, int64_t n, mmf_stream_t stream);
/* out (f32) = x (f32) + d (bf16): the residual stream after the last sub-layer (zorro_utils.py:239) */
int mmf_add_bf16_f32(float* out, const float* x, const void* d, int64_t n, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Mask sampling downstream of the random draws, one single-CTA launch (multimae.py:182-255 generate_random_masks and
 * the per-modality token selection / zorro-mask bookkeeping of forward, :378-426; masks are one row for the batch).
 *   noise1 [sum sizes]: the per-task torch.rand(1, n_t) draws concatenated in task order; noise2 [sum sizes]: the
 *   rand_like draw of :241; share [T]: the Dirichlet sample; want_t = round_half_even(share_t * nenc) (:210).
 * Outputs (device): mask [sum sizes] int64 (0 = visible), ids_restore [sum sizes] / ids_keep [nenc] int64,
 *   idx: task t's ascending visible positions (int32) at offset sum(sizes[:t]), counts [T], seg [T+2] =
 *   [0, c0, c0+c1, .., nenc, nenc + n_fusion], slotmap [T, n_fusion] (rank of a position in idx_t or -1; nullable),
 *   tok [nenc] int32 (nullable): the visible tokens in encoder order (task-major, ascending position) as global ids
 *   sum(sizes[:t]) + position -- the table the token-gather kernels read, so that no count has to reach the host.
 * argsort is stable (rank counting); the reference's CUDA argsort leaves the order of equal keys unspecified.
 * sizes is a HOST array.  sum sizes <= 4096, T <= 8.
 * mmf_mask_explicit: the same bookkeeping for caller-provided masks (multimae.py:372-376: argsort of the 0 / 1 row,
 *   ids_restore, ids_keep, the per-task selections of :378-383) from ONE mask row `given` [sum sizes] int64 (0 = visible),
 *   without the reference's host synchronisations ((mask_all == 0).sum(), three nonzero()).  *err is set to 1 when the row
 *   keeps a number of tokens different from nenc (the reference would silently use that other sequence length).
 * ---------------------------------------------------------------------------------------------- */
int mmf_mask_build(const float* noise1, const float* noise2, const float* share, int32_t T, const int32_t* sizes,
                   int32_t nenc, int32_t n_fusion, int64_t* mask, int64_t* ids_restore, int64_t* ids_keep, int32_t* idx,
                   int32_t* counts, int32_t* seg, int32_t* slotmap, int32_t* tok, mmf_stream_t stream);
int mmf_mask_explicit(const int64_t* given, int32_t T, const int32_t* sizes, int32_t nenc, int32_t n_fusion, int64_t* ids_restore,
                      int64_t* ids_keep, int32_t* idx, int32_t* counts, int32_t* seg, int32_t* slotmap, int32_t* tok, int32_t* err,
                      mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * DINO-style distillation loss, forward and student gradient in one launch (criterion.py:328-335 dino_loss_func;
 * call sites pretrain_mmae.py:489-493).  student / teacher: [B, D] rows (bf16 or f32, is_f32), the teacher is
 * detached.  row_loss [B] f32 = per-sample loss (the caller averages); dstudent [B, D] f32 = d(mean loss)/d student.
 * ---------------------------------------------------------------------------------------------- */
int mmf_dino_loss(const void* student, int64_t lds, const void* teacher, int64_t ldt, int32_t is_f32, int32_t B, int32_t D,
                  float student_temp, float teacher_temp, float* row_loss, float* dstudent, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Debiased hard-negative contrastive loss, forward and both input gradients in one call (criterion.py:214-268
 * HardNegtive_loss.forward incl. get_negative_mask :224-231; call sites pretrain_mmae_s2dsm.py:482-492).
 * out1 / out2: [B, D] f32 rows (row pitches ld1 / ld2), B >= 2, D <= 2048.  easy = 0: estimator 'hard' (tau_plus, beta),
 * 1: estimator 'easy'.  loss: f32 scalar = mean over the 2B rows; dout1 / dout2 [B, D] f32 contiguous = d loss / d out
 * for an upstream gradient of 1.  work: f32 scratch of mmf_hardneg_workspace_floats(B, D) elements.
 * The [2B, 2B] similarity takes bf16-rounded operands with fp32 accumulation (torch.mm under the reference's
 * autocast); normalisation, exp / log and reductions are fp32.
 * ---------------------------------------------------------------------------------------------- */
int64_t mmf_hardneg_workspace_floats(int32_t B, int32_t D);
int mmf_hardneg_loss(const float* out1, int64_t ld1, const float* out2, int64_t ld2, int32_t B, int32_t D, float tau_plus, float beta,
                     float temperature, int32_t easy, float* work, float* loss, float* dout1, float* dout2, mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Masked cross-entropy over class maps (criterion.py:24-58 MaskedCrossEntropyLoss with label_smoothing = 0; the loss of
 * the 4th, semantic, modality `dnw` in pretrain_mmae_my.py:68-75).  logits [B, C, H, W] (bf16 or f32),
 * target [B, H, W] int64 class ids, mask [B, (H/P)*(W/P)] int64 (1 = masked patch, counted) or NULL.  Per pixel
 * logsumexp_c - logit[target]; per sample sum / #masked pixels; batch nanmean; all-zero mask -> 0.  P and W must be
 * multiples of 8.  work: f32 [2B + 2] scratch shared by forward and backward (as for the reconstruction losses).
 * ---------------------------------------------------------------------------------------------- */
int mmf_masked_ce_fwd(const void* logits, int32_t logits_f32, const int64_t* target, const int64_t* mask, int64_t mask_bstride,
                      int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, float* work, float* loss, mmf_stream_t stream);
int mmf_masked_ce_bwd(const void* logits, int32_t logits_f32, const int64_t* target, const int64_t* mask, int64_t mask_bstride,
                      int64_t B, int32_t C, int32_t H, int32_t W, int32_t P, const float* work, const float* dloss, void* dlogits,
                      mmf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training step around the path: multi-tensor AdamW and gradient norm / clip coefficient.
 * Replaces torch.optim.AdamW as configured by utils/optim_factory.py:138-176 (betas (0.9, 0.95), decoupled weight
 * decay on every parameter) and the grad-norm / clip of utils/native_scaler.py:20-82 (one torch.norm per parameter).
 * `tensors` is a DEVICE array, one entry per parameter tensor; (chunk_tensor[c], chunk_index[c]) maps launch block c
 * to elements [chunk_index * chunk_elems, +chunk_elems) of tensor chunk_tensor (DEVICE int32 arrays).  n == 0 skips
 * a tensor (no gradient this step).  Update, per element (torch.optim.AdamW, amsgrad off, maximize off):
 *   g *= *grad_scale (if given);  p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
 *   p -= (lr / bias_correction1) * m / (sqrt(v) / sqrt(bias_correction2) + eps)
 * w16a / w16b (nullable): bf16 images of p refreshed in the same pass (what autocast's weight cast produces each
 * forward in the reference); element i of p goes to image element (i / cols) * pitch16 + i % cols, or i if cols == 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct MmfAdamWTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  void* w16a;
  void* w16b;
  int64_t n;
  int64_t pitch16a, pitch16b;
  int32_t cols;
  float bias_correction1, bias_correction2;   /* 1 - beta^step of THIS tensor */
} MmfAdamWTensor;
int mmf_adamw_step(const MmfAdamWTensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t nchunks,
                   int32_t chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay,
                   const float* grad_scale, mmf_stream_t stream);
/* sqnorm (device scalar, overwritten) = sum of g^2 over all tensors; norm (nullable) = sqrt; clip_coef (nullable) =
 * min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_), 1 if max_norm <= 0.  No host synchronisation. */
int mmf_grad_norm(const MmfAdamWTensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t nchunks,
                  int32_t chunk_elems, float max_norm, float* sqnorm, float* norm, float* clip_coef, mmf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMF_B200_H_ */
