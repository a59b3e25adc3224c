/*
 * mmada_b200 — C ABI of the B200 (sm_100a) kernels behind MMaDA's masked-diffusion denoising path.
 *
 * The reference (MercuryCod/MMaDA) is pure Python/PyTorch and has no FFI: its "operators" for this
 * path are PyTorch library calls.  Each entry point below replaces one of those call sites (cited as
 * file:line under /root/reference) and is what a maintainer would bind with ctypes/cffi from the
 * reference's Python.  Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *   - the library never allocates or frees caller-visible memory and keeps no global state beyond
 *     lazily queried device attributes;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     and returns immediately: 0 on success, 1..99 argument/shape errors, 1000+cudaError_t for
 *     CUDA runtime errors.  No exceptions cross the ABI.
 *   - bf16 = 16-bit bfloat16 storage, row-major, leading dimensions in ELEMENTS.
 */
#ifndef MMADA_B200_H
#define MMADA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMADA_ABI_VERSION 1
int mmada_abi_version(void);
/* Compute capability of the current device as major*10+minor (100 for B200); <0 on error. */
int mmada_device_arch(void);

/* ---- dense projections -------------------------------------------------------------------
 * out = epilogue(A[M,K] . B[N,K]^T), A and B bf16 (K contiguous), fp32 accumulate (tcgen05/TMEM).
 * Replaces nn.Linear at models/modeling_llada.py:901-903 (q/k/v_proj), :724 (attn_out),
 * :924 (ff_proj, up_proj), :930 (ff_out), :1362 (transformer.ff_out == lm_head).            */
enum {
    MMADA_EPI_BF16 = 0,        /* out bf16 [M,N]                                                     */
    MMADA_EPI_F32 = 1,         /* out fp32 [M,N]                                                     */
    MMADA_EPI_RESID_F32 = 2,   /* out fp32 [M,N] = aux fp32 [M,N] (same ld as out; may alias) + acc    */
    MMADA_EPI_SWIGLU_BF16 = 3, /* out bf16 [M,N/2] = silu(gate)*up; B rows interleaved per 128:
                                  rows [256j,256j+128) = ff_proj rows [128j,..), next 128 = up_proj  */
    MMADA_EPI_BIAS_BF16 = 4,   /* out bf16 [M,N] = acc + bias fp32 [N]                               */
    MMADA_EPI_BIAS_F32 = 5,    /* out fp32 [M,N] = acc + bias                                        */
    MMADA_EPI_BIAS_RESID_F32 = 6, /* out fp32 [M,N] = acc + bias + aux fp32 [M,N]                     */
    MMADA_EPI_ROPE_BF16 = 7    /* (mmada_gemm_qkv_rope_bf16 only) bf16 store with rotary embedding       */
};
int mmada_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                    const void* aux, const float* bias, int M, int N, int K, int epilogue, int cta_group,
                    void* stream);

/* Fused q|k|v projection + rotary embedding: out bf16 [M,N] = A . B^T, then NeoX half-split RoPE on every
 * head inside columns [0, rope_cols) (q and k; rope_cols = 2*d_model) with position = row % seq_len and
 * the reference's fp32 sin/cos tables [>= seq_len, head_dim/2]; N % 256 == 0, head_dim 64 or 128.
 * Replaces q/k/v_proj (models/modeling_llada.py:901-903) + RotaryEmbedding.forward (:411-428).     */
int mmada_gemm_qkv_rope_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                             const float* sin_table, const float* cos_table, int M, int N, int K, int rope_cols,
                             int head_dim, int seq_len, int cta_group, void* stream);

/* ---- RMSNorm folded into the projections around it ------------------------------------------------
 * RMSLayerNorm.forward (models/modeling_llada.py:315-329) is x * rsqrt(mean(x^2) + eps) * weight in front of the
 * q|k|v and the ff_proj/up_proj projections (:897-903, :921-924).  Because the projection is linear, the row factor
 * rsqrt(...) can be applied to the ACCUMULATOR rows and the norm weight multiplied into the projection weight's
 * columns (done once by the caller), so the normalised activations never travel through HBM:
 *   producer  mmada_gemm_resid_norm_f32: x fp32 [M,N] += A . B^T (attn_out :724 / ff_out :930 + the residual adds
 *             :915,:933), and in the same epilogue xb bf16 [M,N] = bf16(x) (the next projection's A operand) and
 *             ssq_out fp32 [M, N/256] = per 256-column tile, the row's sum of squares; N % 256 == 0.
 *             mmada_embed_norm_f32 does the same for the embedding rows (ssq_out fp32 [M]).
 *   consumer  mmada_gemm_swiglu_rownorm_bf16 / mmada_gemm_qkv_rope_rownorm_bf16 = MMADA_EPI_SWIGLU_BF16 /
 *             mmada_gemm_qkv_rope_bf16 with every accumulator row m scaled by
 *             rsqrt(sum(row_ssq[m, 0..ssq_tiles)) / norm_dim + eps) first (partials added in index order:
 *             deterministic).                                                                              */
int mmada_gemm_resid_norm_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* x, int64_t ldx,
                              void* xb_bf16, int64_t ld_xb, float* ssq_out, int M, int N, int K, int cta_group,
                              void* stream);
int mmada_gemm_swiglu_rownorm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                                   const float* row_ssq, int ssq_tiles, int norm_dim, float eps, int M, int N, int K,
                                   int cta_group, void* stream);
int mmada_gemm_qkv_rope_rownorm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                                     const float* sin_table, const float* cos_table, const float* row_ssq,
                                     int ssq_tiles, int norm_dim, float eps, int M, int N, int K, int rope_cols,
                                     int head_dim, int seq_len, int cta_group, void* stream);
int mmada_embed_norm_f32(const int64_t* ids, const void* table_bf16, float* out, void* xb_bf16, float* ssq_out, int M,
                         int d, int64_t vocab, void* stream);

/* ---- HBM-bound block kernels ----------------------------------------------------------------
 * embed:   out fp32 [M,d] = table bf16 [vocab,d][ids[m]]         models/modeling_llada.py:1222
 * rmsnorm: out bf16 [M_out,d] = x*rsqrt(mean(x^2)+eps)*weight     models/modeling_llada.py:315-329
 *          x fp32 [*,d] (the fp32 residual stream); rows (int32[M_out], may be NULL) gathers input rows
 * rope:    NeoX half-split rotary embedding in place on the q and k thirds of qkv bf16 [M,ld]
 *          (q at column 0, k at column d_model); row m has position m % seq_len; sin/cos are the
 *          reference's fp32 tables [>=seq_len, head_dim/2]         models/modeling_llada.py:376-428 */
int mmada_embed_f32(const int64_t* ids, const void* table_bf16, float* out, int M, int d, int64_t vocab,
                    void* stream);
int mmada_rmsnorm_bf16(const float* x, const float* weight, void* out_bf16, const int32_t* rows, int M_out,
                       int d, float eps, void* stream);
/* gather_rows: dst[i, 0..row_bytes) = src[rows[i], 0..row_bytes), any element type (row_bytes, ld_src_bytes multiples of
 *          16).  Used to restrict the LAST block's attn_out / MLP (models/modeling_llada.py:914-933) to the token rows
 *          whose logits the caller reads (modeling_mmada.py:167-168 slices the image rows): rows are independent
 *          after the attention, so the results are bit-identical to running every row.                        */
int mmada_gather_rows(const void* src, int64_t ld_src_bytes, const int32_t* rows, void* dst, int n_rows, int row_bytes,
                      void* stream);
int mmada_rope_inplace_bf16(void* qkv_bf16, int64_t ld, const float* sin_table, const float* cos_table, int M,
                            int d_model, int head_dim, int seq_len, void* stream);

/* ---- attention -----------------------------------------------------------------------------
 * out[b*L+t, h*hd:(h+1)*hd] = softmax(q k^T * scale) v for every (b,h); bidirectional, no mask, no
 * KV cache.  q/k/v bf16 [B*L, ld] with head h at columns h*hd (they may be the three thirds of one
 * fused buffer); out bf16 [B*L, ldo].  head_dim 64 or 128.
 * Replaces F.scaled_dot_product_attention at models/modeling_llada.py:653-660.              */
int mmada_attention_bf16(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo,
                         int B, int L, int H, int head_dim, float scale, void* stream);

/* ---- t2i sampling step ------------------------------------------------------------------------
 * One launch = models/modeling_mmada.py:164-209 + models/sampling.py:31-36 for one denoising step
 * on already-sliced logits: CFG mix (uncond may be NULL), softmax, argmax(p/q), known-token
 * override, selected probability, mask_len clamp, log p + T*gumbel(u), k-th smallest cut-off with
 * strict '<', write-back.
 *   cond/uncond/q fp32 [B*N, C] (C in {512,1024,2048,4096,8192}); u fp32 [B,N];
 *   known_ids int64 [B,N] in/out (code id, or mask_id where unknown);
 *   input_ids int64 [B, ld_ids] in/out, image tokens at columns [img_off, img_off+N) (may be NULL);
 *   sampled_out int64 [B,N]; sel_out fp32 [B,N]; masking_out uint8 [B,N] (may be NULL);
 *   raw_out int64 [B,N] (may be NULL): raw argmax at every position, known ones included;
 *   no_remask != 0: commit the merged tokens without re-masking — both for
 *   t2m_generate (models/modelling_ours.py:634-682, Appendix A Q15);
 *   tickets int32 [B], zero on entry and left zero.                                            */
int mmada_t2i_sample_step(const float* cond_logits, const float* uncond_logits, const float* q_noise,
                          const float* u_noise, int64_t* known_ids, int64_t* input_ids, int64_t ld_ids,
                          int64_t img_off, int64_t* sampled_out, float* sel_out, uint8_t* masking_out,
                          int64_t* raw_out, int no_remask, int32_t* tickets, int B, int N, int C,
                          float one_plus_g, float g, float mask_len_raw, float temperature,
                          int64_t mask_id, int64_t text_vocab, void* stream);
/* Output head restricted to the still-masked positions (the reference computes the logits of every position and
 * discards those of known ones, models/modeling_mmada.py:183-184).
 *   mmada_compact_masked_rows: from known_ids int64 [B,N], rows_out int32 [branches*B*cap] = flattened token rows
 *     ((r*B + b)*L + img_off + n) of the masked positions of batch row b, cond branch r = 0 then (branches == 2) uncond
 *     r = 1, padded per row up to `cap` with the row's first image position; slot_out int32 [B,N] = b*cap + j for the
 *     j-th masked position of row b, -1 for known ones.  `cap` = upper bound on the masked positions per row.
 *   mmada_t2i_sample_step_compact: mmada_t2i_sample_step on cond/uncond fp32 [B*cap, C] indexed through logit_slot.   */
int mmada_compact_masked_rows(const int64_t* known_ids, int32_t* rows_out, int32_t* slot_out, int B, int N, int L,
                              int img_off, int cap, int branches, int64_t mask_id, void* stream);
int mmada_t2i_sample_step_compact(const float* cond_logits, const float* uncond_logits, const float* q_noise,
                                  const float* u_noise, int64_t* known_ids, int64_t* input_ids, int64_t ld_ids,
                                  int64_t img_off, int64_t* sampled_out, float* sel_out, uint8_t* masking_out,
                                  int no_remask, int32_t* tickets, int B, int N, int C, float one_plus_g, float g,
                                  float mask_len_raw, float temperature, int64_t mask_id, int64_t text_vocab,
                                  const int32_t* logit_slot, void* stream);
/* masking[b,n] = conf[b,n] < sort(conf[b])[mask_len[b]], conf = log(max(p,1e-20)) + T*gumbel(u).
 * Replaces mask_by_random_topk, models/sampling.py:31-36.                                      */
int mmada_mask_by_random_topk(const float* probs, const float* u_noise, const int64_t* mask_len,
                              uint8_t* masking_out, int B, int N, float temperature, void* stream);

/* ---- text / MMU path: low-confidence remasking ------------------------------------------------
 * text_sample_rows: per candidate row r (R rows of V fp32 logits): optional CFG mix
 *   un + (cfg+1)*(l-un) (un_logits may be NULL), Gumbel-max token x0 = argmax exp(l64)/(-log u)^T in
 *   fp64 (temperature 0 -> plain argmax), and conf = softmax(l64)[x0] in fp64.  u_noise fp64 [R,V]
 *   supplies the uniforms (parity); NULL draws them in-kernel (Philox4x32-10 keyed by seed,row,col).
 *   Replaces generate.py:8-19,86,90-96 == models/modeling_mmada.py:49-60,436,441-448.
 * block_mask_count: cnt[b] = #(x[b, lo:lo+block] == mask_id)           generate.py:76-77,22-28
 * text_transfer: k = cnt/steps + (step < cnt%steps); among the masked positions of the block pick the k
 *   largest conf (ties: lower position) and set x[b, lo+p] = x0[b*block+p].  conf_override (fp64
 *   [B,block], may be NULL) replaces conf ('random' remasking).  generate.py:102-111,30-40        */
int mmada_text_sample_rows(const float* logits, const float* un_logits, float cfg_scale_plus1,
                           const double* u_noise, uint64_t seed, float temperature, int R, int V,
                           int64_t* x0_out, double* conf_out, void* stream);
int mmada_block_mask_count(const int64_t* x, int64_t ld, int lo, int block, int B, int64_t mask_id,
                           int32_t* cnt_out, void* stream);
int mmada_text_transfer(int64_t* x, int64_t ld, int lo, int block, const int64_t* x0, const double* conf,
                        const double* conf_override, const int32_t* cnt, int steps, int step, int B,
                        int64_t mask_id, uint8_t* transfer_out, void* stream);

/* ---- training-time forward: masked cross-entropy rows -------------------------------------------
 * nll[r] = logsumexp(logits[r, 0..V)) - logits[r, labels[r]]  (fp32, row pitch ld elements); 0 where
 * labels[r] == ignore_index.  One pass over each row.  Replaces F.cross_entropy(..., reduction='none') on
 * the loss rows of MMadaModelLM.forward_process, models/modeling_mmada.py:240-243,253-256,264-267.        */
int mmada_cross_entropy_rows_f32(const float* logits, int64_t ld, const int64_t* labels, int64_t ignore_index,
                                 float* nll_out, int R, int V, void* stream);

/* ---- MAGVIT-v2 token -> pixel path ------------------------------------------------------------
 * Activations are NHWC (channels contiguous).  conv: implicit GEMM on tcgen05, zero padding by TMA
 * out-of-bounds fill; in bf16 [B,H,W,C_in] (C_in % 64 == 0), weight bf16 [C_out][taps][C_in] (taps 9 = 3x3
 * pad 1, or 1), bias fp32 [C_out]; epilogue one of MMADA_EPI_BIAS_{BF16,F32,RESID_F32} (resid fp32
 * [B,H,W,C_out]).  Replaces torch.nn.Conv2d in models/modeling_magvitv2.py:309-362 and
 * models/common_modules.py (ResnetBlock, Upsample, AttnBlock).                                       */
int mmada_conv_nhwc_bf16(const void* in, const void* weight, const float* bias, void* out, const void* resid,
                         int B, int H, int W, int C_in, int C_out, int taps, int epilogue, void* stream);
/* LFQuantizer (models/modeling_magvitv2.py:186-221).  lfq_decode: indices [total] -> bf16 NHWC
 * [total,64], channels 0..12 = post_quant_conv (1x1, weight fp32 [13,13], bias [13]) applied to the
 * -1/+1 bits (MSB first), channels 13..63 zero.  indices_to_bits: [B,N] -> fp32 [B,13,N] (NCHW).
 * bits_to_indices: fp32 [B,13,N] -> int64 [B,N].                                                    */
int mmada_lfq_decode_nhwc(const int64_t* indices, const float* pq_weight, const float* pq_bias, void* out_bf16,
                          int total_tokens, void* stream);
int mmada_lfq_indices_to_bits(const int64_t* indices, float* out_nchw, int B, int N, void* stream);
int mmada_lfq_bits_to_indices(const float* z_nchw, int64_t* out, int B, int N, void* stream);
/* GroupNorm(32 groups) over fp32 NHWC [B,P,C]: stats accumulates (sum, sum of squares) per (b, group)
 * into sums fp64 [B,32,2] (zeroed by the call); apply writes bf16 (x-mean)*rstd*gamma+beta, optional
 * swish.  models/common_modules.py:16-24.                                                           */
int mmada_groupnorm_stats(const float* x, double* sums, int B, int P, int C, void* stream);
int mmada_groupnorm_apply_bf16(const float* x, const double* sums, const float* gamma, const float* beta,
                               void* out_bf16, int B, int P, int C, float eps, int swish, void* stream);
int mmada_upsample2x_nhwc_bf16(const float* x, void* out_bf16, int B, int H, int W, int C, void* stream);
int mmada_cast_f32_bf16(const float* x, void* out_bf16, int64_t n, void* stream);
int mmada_softmax_rows_bf16(const float* x, void* out_bf16, int R, int n, float scale, void* stream);
int mmada_nhwc_to_nchw_f32(const float* x, float* out, int B, int P, int C, void* stream);
int mmada_image_to_uint8(const float* x, uint8_t* out, int64_t n, void* stream);
/* Encoder side (MAGVITv2.get_code, models/modeling_magvitv2.py:143-169,423-427).  image_to_nhwc64: pixel_values
 * fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64], channels 3..63 zero (conv_in input).  space_to_depth2: fp32 NHWC
 * [B,H,W,C] -> bf16 NHWC [B,H/2,W/2,4C], channel (2sy+sx)C+c = pixel (2y+sy, 2x+sx): Downsample
 * (models/common_modules.py:73-90, pad (0,1,0,1) + 3x3 stride 2) then runs as a stride-1 3x3 convolution whose
 * weights the host has rearranged (taps reaching above / left of the block are zero).                          */
int mmada_image_to_nhwc64_bf16(const float* pixels_nchw, void* out_bf16, int B, int H, int W, void* stream);
int mmada_space_to_depth2_bf16(const float* x, void* out_bf16, int B, int H, int W, int C, void* stream);
/* Motion VQ-VAE decoder (motion_vqvae/models/encdec.py:35-67, resnet.py:12-81; VQVAE_251.forward_decoder
 * vqvae.py:74-81).  conv1d_gather: x fp32 [B,T_in,C] -> bf16 [B, T_in*upsample, taps*C]: for output frame t the
 * frames t + (k - taps/2)*dilation of the (nearest-2x upsampled, if upsample == 2) sequence side by side, zero
 * outside [0, T_out), optional ReLU on the way; the Conv1d is then mmada_gemm_bf16 with weight [C_out, taps*C].
 * relu_f32: in place.                                                                                       */
int mmada_conv1d_gather_bf16(const float* x, void* out_bf16, int B, int T_in, int C, int taps, int dilation,
                             int upsample, int relu, void* stream);
int mmada_relu_f32(float* x, int64_t n, void* stream);

/* Sequence assembly of the generation tasks (UniversalPrompting, training/prompting_utils.py).  text: every row's
 * pre-tokenised text ids concatenated, text_off [B+1] offsets; body [B,N] image / motion tokens (row pitch ld_body).
 * A text gets `bos` in front unless it starts with it (an empty text becomes [bos]).
 * mode 0 (t2i_gen_prompt :200-233; t2m_prompt :87-144 without conditional drop-out): ids [B, text_slots + N + 2] =
 *   [pad.. task bos text eos] (text_slots = the class's max_text_len + 1 ids, left-padded; a longer text is cut to
 *   text_slots - 1 ids + eos) open body close;  mask [B, L] = 0 on the padding, 1 elsewhere.
 * mode 1 (mmu_gen_prompt :379-425): ids [B, 3 + N + text_slots] = task open body close [bos text eos, eos..] (text_slots =
 *   the class's max_text_len; cut to text_slots - 1 ids + eos);  mask [B] = prompt_length: the head plus the text up to
 *   and including its last `end_header` token (:401-413).                                                         */
int mmada_build_prompts(const int64_t* text, const int64_t* text_off, const int64_t* body, int64_t ld_body, int64_t* ids,
                        int64_t* mask, int B, int N, int text_slots, int mode, int64_t task_token, int64_t bos,
                        int64_t eos, int64_t pad, int64_t open_token, int64_t close_token, int64_t end_header,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMADA_B200_H */
