/* agenda_b200.h — C ABI of libagenda_b200.so (sm_100a).
 *
 * Drop-in boundary for the AGenDA heat-map data-generation hot path.  The reference
 * (humansensinglab/AGenDA) is pure Python and has no FFI of its own; each entry point below
 * replaces the Python code cited beside it (paths relative to the reference root).  The
 * Python binding a maintainer adds on the reference side is a ctypes stub — see INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no framework types; device pointers unless stated "host".
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it, nothing synchronises,
 *     nothing allocates device memory, no global state is kept beyond a per-thread cache of TMA descriptors
 *     (a descriptor is a pure function of address, shape and box; it is passed by value to the kernel).
 *   - returns 0 (AGENDA_OK) or a negative error code; agenda_last_error() gives a thread-local
 *     human-readable message for the last failure on the calling thread.
 *   - re-entrant across streams and devices (uses the calling thread's current device).
 */
#ifndef AGENDA_B200_H_
#define AGENDA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGENDA_OK 0
#define AGENDA_ERR_NULL_POINTER (-1)
#define AGENDA_ERR_BAD_SHAPE (-2)
#define AGENDA_ERR_UNSUPPORTED (-3)
#define AGENDA_ERR_MISALIGNED (-4)
#define AGENDA_ERR_CUDA (-5)

/* element types of Q/K/V/O */
#define AGENDA_F32 0
#define AGENDA_BF16 1

/* ABI version (major*1000 + minor). */
int agenda_version(void);
/* Thread-local message for the last non-zero return on this thread ("" if none). */
const char* agenda_last_error(void);
/* 1 if the calling thread's current device is compute capability 10.x, else 0 (or a negative error). */
int agenda_device_ok(void);

/* ---- a2: attention core of UNetCrossAttentionHooker.__call__ (data_generation/hook.py:104-115) --------
 * q [B,N,H*d], k/v [B,M,H*d], out [B,N,H*d]; row-major, token-major: head h owns columns [h*d,(h+1)*d)
 * (this IS the layout to_q/to_k/to_v produce at hook.py:93,101-102 — head_to_batch_dim/batch_to_head_dim
 * (hook.py:104-106,115) are folded into the kernels' addressing).  scale = d**-0.5 in the reference.
 *
 * Self-attention (encoder_hidden_states is None): flash-style fused softmax(scale*QK^T)V on tcgen05 tensor
 * cores fed by TMA, fp32 softmax/accumulate.  dtype BF16 only; d in {40,64,80,160}; N>=1.  Replaces
 * torch.baddbmm + softmax + torch.bmm (hook.py:108,114) without materialising [B*H,N,N].
 */
int agenda_attn_self_fwd(const void* q, const void* k, const void* v, void* out, int dtype,
                         int B, int H, int N, int d, float scale, void* stream);

/* Same, with q/k/v rows `ld` elements apart (batch stride N*ld): q, k, v may be column slices [.., 0:C], [.., C:2C],
 * [.., 2C:3C] of ONE fused projection output [B,N,3C] (to_q/to_k/to_v of hook.py:93,101-102 act on the same
 * hidden_states in self-attention, so one GEMM with the concatenated weights reads them once).  out stays packed
 * [B,N,H*d].  ld >= H*d, ld % 8 == 0, pointers 16-byte aligned. * scale == 0 means the caller has already multiplied q by scale * log2(e) (the processor folds it into W_q): the
 * scores are then the base-2 exponents themselves and the d = 40 kernel skips the scale/shift FMA. */
int agenda_attn_self_fwd_strided(const void* q, const void* k, const void* v, void* out, int dtype,
                                 int B, int H, int N, int d, long long ld, float scale, void* stream);

/* Test hook: agenda_attn_self_fwd (bf16) with an explicit kernel variant: 0 = default (two 128-query tiles per
 * CTA, ping-pong softmax warpgroups, P through TMEM), 1 = one query tile per CTA with P through TMEM (TS-form
 * tcgen05.mma), 2 = one query tile per CTA with P through a 128B-swizzled shared-memory tile (SS-form);
 * 10+e = variant 0 with a share of the exponentials on the FMA pipe (e in {0,2,3,4,8} -> 0/50/37.5/25/12.5 %);
 * 20+e = the same with three query tiles per CTA and 64-key tiles (d = 40 or 64 only); 30+e two warpgroups per query
 * tile; 40+e d = 80 with 64-key tiles; 50+e three query tiles with 80-key tiles (d = 40).  All of 10..59 track the
 * exact running maximum in a single pass; 60 = the shipped default of each head dim WITH the fast first pass
 * (first-tile maximum, row-sum check, exact second pass inside the CTA when the check fails). */
int agenda_attn_self_fwd_variant(const void* q, const void* k, const void* v, void* out, int B, int H, int N,
                                 int d, float scale, int variant, void* stream);

/* Same contract in full fp32 on the CUDA cores (any d<=160, dtype F32 or BF16 inputs): the exact-precision
 * path used for fp32 pipelines and for parity checks. */
int agenda_attn_self_fwd_f32(const void* q, const void* k, const void* v, void* out, int dtype,
                             int B, int H, int N, int d, float scale, void* stream);

/* Cross-attention with the heat-map epilogue (hook.py:108-114 + _unravel_attn, hook.py:28-56).
 * Besides out, writes for every kept batch element b >= b_first (b_first = B/2 when is_train is False,
 * hook.py:48-49, else 0) and every selected key token t:
 *     maps[b-b_first, t, n] (=|+=) mean over heads of softmax(scale*q k^T)[b, head, n, token_idx[t]]
 * maps is fp32 [B-b_first, T, N].  token_idx: HOST int32[T], or NULL for "all M tokens in order" (T==M), the
 * reference's behaviour.  accumulate!=0 adds into maps (used when N == latent_hw^2, where hook.py:72's
 * bicubic resize is the identity and the clamp a no-op, so the epilogue accumulates straight into the
 * persistent heat buffer); accumulate==0 overwrites (native-resolution scratch for
 * agenda_heat_upsample_accum).  Softmax and the head mean are fp32.  M <= 128.
 */
int agenda_attn_cross_fwd_heat(const void* q, const void* k, const void* v, void* out, int dtype,
                               int B, int H, int N, int M, int d, float scale,
                               const int32_t* token_idx, int T, int b_first,
                               float* maps, int accumulate, void* stream);

/* DAAM-style capture (the `daam` package data_generation.py:57-77 is written against keeps one map per head and
 * upsamples before it averages — SURVEY.md §8 a5): same attention, but
 *     maps[b-b_first, head, t, n] (=|+=) softmax(scale*q k^T)[b, head, n, token_idx[t]]       (no head mean)
 * maps is fp32 [B-b_first, H, T, N] and must not be NULL.  accumulate!=0 adds (DAAM sums over denoising steps at
 * native resolution).  The tensor-core path covers T <= 8; more tokens take the fp32 CUDA-core kernel. */
int agenda_attn_cross_fwd_heat_heads(const void* q, const void* k, const void* v, void* out, int dtype,
                                     int B, int H, int N, int M, int d, float scale,
                                     const int32_t* token_idx, int T, int b_first,
                                     float* maps, int accumulate, void* stream);

/* ---- to_q of the cross-attention heat path (hook.py:93) with an fp32 result, one tcgen05 GEMM ---------------------------
 * out[M,N] (fp32, row-major) = x[M,K] (bf16, row-major) * (w_hi + w_lo)[N,K]^T, w_hi / w_lo bf16 row-major in the
 * nn.Linear layout; w_lo may be NULL (plain fp32-output GEMM).  w_hi = bf16(W), w_lo = bf16(W - w_hi) of an fp32
 * checkpoint weight W (agenda_b200/mixed.py) reproduce the fp32 projection to 2^-17: products of bf16 values are exact
 * in the fp32 accumulator, and both halves accumulate into the same TMEM tile, so x is read once and out written once.
 * K % 64 == 0, N % 160 == 0 (320 / 640 / 1280 in the SD UNets); pointers 16-byte aligned; no bias. */
int agenda_linear_split_f32(const void* x, const void* w_hi, const void* w_lo, float* out, int M, int K, int N,
                            void* stream);
/* The same GEMM with the weights pre-packed for bulk copies (they are constants of the model): the kernel is bound by the
 * TMA engine's per-row cost — per 64-wide K block it issues 128 activation rows + 160 (+160) weight rows of 128 bytes — so
 * agenda_linear_split_pack_w lays w_hi | w_lo out once as blob[N/160][K/64][hi tile | lo tile] (160 rows x 128 bytes each, in
 * the 128B-swizzled UMMA layout: the exact image of a pipeline stage) and agenda_linear_split_f32_packed fetches a K block's
 * weights with ONE bulk copy.  agenda_linear_split_pack_bytes gives the blob size.  Same results bit for bit. */
long long agenda_linear_split_pack_bytes(int N, int K, int has_lo);
int agenda_linear_split_pack_w(const void* w_hi, const void* w_lo, void* blob, int N, int K, void* stream);
int agenda_linear_split_f32_packed(const void* x, const void* w_blob, int has_lo, float* out, int M, int K, int N,
                                   void* stream);
/* The same GEMM with the result in the CHUNK-MAJOR layout agenda_attn_cross_fwd_heat_x3_hm streams: x is [B*rows_per_batch, K]
 * (M = B * rows_per_batch), N = heads * d with d a multiple of 40, and
 *     out[b][h][c][n][j] = (x W^T)[b * rows_per_batch + n][h * d + c * 40 + j]        (fp32, c < d / 40, j < 40),
 * so the 128 queries x 40 columns the attention kernel consumes per step are one dense 20 KB block (one bulk copy instead of
 * a 128-row tensor load) and the GEMM's epilogue writes whole rows of a block straight from its accumulator registers. */
int agenda_linear_split_f32_heads(const void* x, const void* w_hi, const void* w_lo, float* out, int M, int K, int N,
                                  int rows_per_batch, int heads, void* stream);

/* ---- prompt side of the split-precision cross-attention: pack K / V once per prompt --------------------------------
 * to_k / to_v of the prompt embedding (hook.py:101-102) do not depend on the latent.  agenda_pack_context_kv turns the
 * fp32 key projection k32 [B,M,H*d] and the value projection v [B,M,H*d] (v_dtype AGENDA_F32 or AGENDA_BF16) into
 * `blob`: for every (batch, head) the exact shared-memory image agenda_attn_cross_fwd_heat_x3 feeds to the tensor
 * cores — K_hi = bf16(K) chunks | K_lo = bf16(K - K_hi) chunks | V (bf16) chunks, 80-row x 128-byte tiles in the
 * 128B-swizzled UMMA layout, zero padded — so that a head's K and V each arrive with one bulk copy.
 * agenda_context_blob_bytes gives the size of `blob` in bytes (negative error code for an unsupported d).
 * M <= 80; d in {40,64,80,160}; pointers 16-byte aligned. */
long long agenda_context_blob_bytes(int B, int H, int d);
int agenda_pack_context_kv(const float* k32, const void* v, int v_dtype, void* blob, int B, int H, int M, int d,
                           void* stream);

/* Cross-attention with fp32-accurate logits on the bf16 tensor cores ("bf16 x 3" split): the path that holds the heat
 * maps to the 1e-4 tolerance against the reference's fp32 baddbmm + softmax (hook.py:108; the reference pipeline is
 * fp32, data_generation.py:30-31).  q is FP32 [B,N,H*d] (a to_q output with fp32 accumulation); kv_blob is the packed
 * prompt side from agenda_pack_context_kv (same B, H, M, d).  The kernel splits every Q tile into hi / lo bf16 on chip
 * and accumulates Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T in fp32 (products of bf16 values are exact; the dropped
 * lo*lo term is 2^-18 relative).  Softmax, head mean and maps are fp32 as in agenda_attn_cross_fwd_heat; P V runs with
 * bf16 P and V.  out [B,N,H*d] in out_dtype (AGENDA_BF16 or AGENDA_F32).
 * maps: per_head == 0 -> [B-b_first, T, N] (mean over heads; token_idx HOST int32[T], or NULL with T == M for all
 * prompt tokens in order, the reference's behaviour); per_head != 0 -> [B-b_first, H, T, N] (no mean, DAAM-style,
 * T <= 8).  maps NULL skips the epilogue.  T <= 8 takes the two-warpgroup kernel, larger T the all-token form.
 * M <= 80; d in {40,64,80,160}; q / kv_blob / out 16-byte aligned. */
int agenda_attn_cross_fwd_heat_x3(const float* q, const void* kv_blob, void* out, int out_dtype, int B, int H, int N,
                                  int M, int d, float scale, const int32_t* token_idx, int T, int b_first,
                                  int per_head, float* maps, int accumulate, void* stream);
/* The same call with q in the chunk-major layout of agenda_linear_split_f32_heads ([B][H][d/40][N][40] fp32; d in
 * {40,80,160}): identical arithmetic, the Q chunks arrive by bulk copy (a ragged last tile copies its rows, the rest of the
 * stage is zero-filled, which is what the tensor map's out-of-bounds fill gives the row-major call). */
int agenda_attn_cross_fwd_heat_x3_hm(const float* q_hm, const void* kv_blob, void* out, int out_dtype, int B, int H, int N,
                                     int M, int d, float scale, const int32_t* token_idx, int T, int b_first,
                                     int per_head, float* maps, int accumulate, void* stream);

/* Backward of agenda_attn_cross_fwd_heat (training mode, SURVEY.md §8 f N3: autograd of hook.py:104-115 and of
 * _unravel_attn hook.py:28-56 as exercised by finetune_sd_token.py:1043-1069).  Given d_out [B,N,H*d] (same dtype as
 * q/k/v) and d_maps fp32 [B-b_first, T, N] (gradient of the head-mean maps; NULL when the maps were not used), writes
 * dq [B,N,H*d] (dtype of q) and ADDS dK / dV into the fp32 accumulators dk, dv [B,M,H*d], which the caller zero-fills
 * (many CTAs per (batch, head) add into them with atomics: the last bits of dK / dV are not reproducible run to run).
 * P is recomputed in fp32; M <= 96.  token_idx / T / b_first as in the forward call. */
int agenda_attn_cross_bwd(const void* q, const void* k, const void* v, const void* d_out, const float* d_maps,
                          void* dq, float* dk, float* dv, int dtype, int B, int H, int N, int M, int d, float scale,
                          const int32_t* token_idx, int T, int b_first, void* stream);

/* The same backward on the tcgen05 tensor cores, for bf16 tensors and the few-token form (d_maps NULL or 1 <= T <= 8
 * selected tokens; M <= 80; d in {40,64,80,160}).  Two stages: dQ (a thread owns a query row: softmax over the 77 scores in
 * registers, the heat-map gradient patched into the dP tile's token columns, Delta, dS = P (dP - Delta) scale, dQ = dS K as a
 * TS-form MMA), which leaves every row's log-sum-exp and Delta in `workspace`; then dK / dV on the transposed tiles (rows =
 * keys, the query range split over the SMs, partial sums ADDED into the caller-zeroed fp32 dk / dv [B,M,H*d] with atomics).
 * q, k, v, d_out, dq bf16; workspace: agenda_attn_cross_bwd_tc_workspace_bytes(B, H, N) bytes, 16-byte aligned. */
long long agenda_attn_cross_bwd_tc_workspace_bytes(int B, int H, int N);
int agenda_attn_cross_bwd_tc(const void* q, const void* k, const void* v, const void* d_out, const float* d_maps,
                             void* workspace, void* dq, float* dk, float* dv, int B, int H, int N, int M, int d,
                             float scale, const int32_t* token_idx, int T, int b_first, void* stream);

/* Backward of agenda_attn_self_fwd (training mode, SURVEY.md §8 f N3: autograd of hook.py:104-115 with
 * encoder_hidden_states None, as exercised by finetune_sd_token.py:1043-1069,1089) on the tcgen05 tensor cores.
 * q, k, v, out (the forward result), d_out and the gradients dq, dk, dv are bf16 [B,N,H*d].  P is recomputed from a
 * log-sum-exp pass (the forward kernel stores no statistics): five launches — Delta = rowsum(dO o O), LSE, dQ, dK, dV —
 * all enqueued on `stream`.  workspace: agenda_attn_self_bwd_workspace_bytes(B, H, N) bytes of device memory (fp32
 * [2,B,H,N]: log2-domain LSE and Delta).  d in {40,64,80,160}; all pointers 16-byte aligned. */
long long agenda_attn_self_bwd_workspace_bytes(int B, int H, int N);
int agenda_attn_self_bwd(const void* q, const void* k, const void* v, const void* out, const void* d_out,
                         void* workspace, void* dq, void* dk, void* dv, int dtype, int B, int H, int N, int d,
                         float scale, void* stream);
/* The forward / backward pair that skips the backward's log-sum-exp pass: agenda_attn_self_fwd_lse is agenda_attn_self_fwd
 * that also writes lse [B,H,N] fp32 = log2 sum_j 2^(scale log2e q_i.k_j) from the softmax epilogue it runs anyway (only the
 * multi-tile kernels emit it: agenda_attn_self_fwd_emits_lse(N, d) tells; otherwise AGENDA_ERR_UNSUPPORTED), and
 * agenda_attn_self_bwd_lse is agenda_attn_self_bwd with that vector handed in (four launches instead of five; three at
 * d <= 64). */
int agenda_attn_self_fwd_emits_lse(int N, int d);
int agenda_attn_self_fwd_lse(const void* q, const void* k, const void* v, void* out, float* lse, int dtype, int B, int H,
                             int N, int d, float scale, void* stream);
int agenda_attn_self_bwd_lse(const void* q, const void* k, const void* v, const void* out, const void* d_out,
                             const float* lse, void* workspace, void* dq, void* dk, void* dv, int dtype, int B, int H, int N,
                             int d, float scale, void* stream);

/* Test hook: same contract, forcing the exact fp32 CUDA-core kernel (bf16 inputs otherwise take the tcgen05
 * tensor-core kernel; fp32 inputs always take the fp32 kernel). */
int agenda_attn_cross_fwd_heat_f32(const void* q, const void* k, const void* v, void* out, int dtype,
                                   int B, int H, int N, int M, int d, float scale,
                                   const int32_t* token_idx, int T, int b_first,
                                   float* maps, int accumulate, void* stream);

/* Attention with an additive mask: data_generation/hook.py:92 (`attn.prepare_attention_mask`) and :108
 * (`attn.get_attention_scores(query, key, attention_mask)` = baddbmm(mask, q, k^T, alpha = scale) + softmax).
 * mask: fp32 [B*H, mask_rows, M] with mask_rows 1 (broadcast over queries) or N, what prepare_attention_mask returns.
 * One entry for both attention kinds: maps == NULL: out = softmax(scale q k^T + mask) v (self-attention: M = N);
 * maps != NULL: cross-attention with the heat epilogue of agenda_attn_cross_fwd_heat (token_idx, T, b_first,
 * accumulate as there; per_head != 0: planes [B - b_first, H, T, N] without the head mean), M <= 128.
 * q [B,N,H*d], k / v [B,M,H*d], out [B,N,H*d] in `dtype` (f32 or bf16).  Exact fp32 CUDA-core path (the SD UNets pass
 * no mask; this is for pipelines that do).  A fully masked row yields NaN, as torch's softmax does. */
int agenda_attn_fwd_masked(const void* q, const void* k, const void* v, void* out, int dtype, int B, int H, int N,
                           int M, int d, float scale, const float* mask, int mask_rows,
                           const int32_t* token_idx, int T, int b_first, int per_head, float* maps,
                           int accumulate, void* stream);

/* ---- a4: compute_global_heat_map (data_generation/hook.py:59-81), streaming form ----------------------
 * acc[i, y, x] += max(0, bicubic(maps[i])[y, x]) for i < n_planes; maps fp32 [n_planes,h,w] -> acc fp32
 * [n_planes,L,L].  Bicubic = torch F.interpolate(mode='bicubic', align_corners=False): A=-0.75,
 * src=(dst+0.5)*h/L-0.5, border-replicated taps. */
int agenda_heat_upsample_accum(const float* maps, float* acc, int n_planes, int h, int w, int L,
                               void* stream);
/* DAAM-style aggregation: acc[b'*T + t] += sum over g < G of max(0, bicubic(maps[(b'*G + g)*T + t])), g ascending
 * (deterministic).  maps fp32 [n_acc_planes/T, G, T, h, w] (G = heads), acc fp32 [n_acc_planes, L, L]. */
int agenda_heat_upsample_accum_heads(const float* maps, float* acc, int n_acc_planes, int T, int G, int h, int w,
                                     int L, void* stream);
/* out[i] = acc[i] / count  (torch.mean over the (layer x step) list, hook.py:79).  count>=1. */
int agenda_heat_finalize(const float* acc, float* out, int64_t n_elems, int count, void* stream);

/* ---- a7: data_generation/data_generation.py:82-85 --------------------------------------------------------
 * u8 = uint8(trunc((h-min)/((max-min)+1e-8f)*255)) per map, all fp32, IEEE round-to-nearest, no FMA
 * contraction (bit-exact with numpy).  heat fp32 [n,hw] -> out u8 [n,hw]. */
int agenda_heat_normalize_u8(const float* heat, uint8_t* out, int n, int hw, void* stream);
/* PIL Image.resize default (BICUBIC, a=-0.5, 8-bit two-pass fixed point, 22 fractional bits), bit-exact.
 * in u8 [n,Hi,Wi] -> out u8 [n,Ho,Wo]. */
int agenda_resize_bicubic_u8(const uint8_t* in, uint8_t* out, int n, int Hi, int Wi, int Ho, int Wo,
                             void* stream);
/* Fused normalise -> u8 -> resize: heat fp32 [n,Hi,Wi] -> out u8 [n,Ho,Wo]. */
int agenda_heat_to_u8_image(const float* heat, uint8_t* out, int n, int Hi, int Wi, int Ho, int Wo,
                            void* stream);

/* ---- a8: data_generation/postprocess_heatmap.py:44-46 ----------------------------------------------------
 * inv = 255 - bg; stack[...,0]=obj, [...,1]=fg, [...,2]=inv.  u8 [n,H,W] x3 -> stack u8 [n,H,W,3], inv u8
 * [n,H,W] (inv may be NULL). */
int agenda_stack_heatmaps_u8(const uint8_t* obj, const uint8_t* fg, const uint8_t* bg, uint8_t* stack,
                             uint8_t* inv, int n, int H, int W, void* stream);
/* a7+a8 fused for the (object, fg-token, bg-token) triple: heat fp32 [n,3,Hi,Wi] -> planes u8 [n,3,Ho,Wo]
 * (the three daam_{word}_heatmaps PNG payloads; may be NULL), stack u8 [n,Ho,Wo,3], inv u8 [n,Ho,Wo] (may
 * be NULL). */
int agenda_heat_postprocess_stack(const float* heat, uint8_t* planes, uint8_t* stack, uint8_t* inv, int n,
                                  int Hi, int Wi, int Ho, int Wo, void* stream);

/* ---- N4: element-wise glue of the rest of the denoising step (data_generation/data_generation.py:59 runs a diffusers
 * UNet2DConditionModel around the attention processor; SURVEY.md §8 f N4).  bf16, channels-last. -----------------------
 * agenda_groupnorm_nhwc: torch.nn.GroupNorm(G, C, eps) on x [B, HW, C] (a channels-last [B,C,H,W] tensor's memory),
 * statistics in fp32 over each (batch, group) = HW * C/G values, biased variance, y = (x - mean) * rstd * gamma + beta,
 * optionally followed by SiLU (diffusers ResnetBlock2D: conv(silu(norm(x)))); y [B, HW, C] bf16.  gamma / beta bf16 [C]
 * (NULL: 1 / 0).  pre_add bf16 [B, C] (may be NULL): a per-(batch, channel) term added to x BEFORE the normalisation —
 * ResnetBlock2D's conv1 bias + time-embedding projection, which otherwise cost two passes over the tensor.
 * workspace: agenda_groupnorm_workspace_bytes(B, HW, C, G) bytes (per-slab partial sums; no atomics, so
 * results are bit-reproducible).  C % 8 == 0, C % G == 0, G <= 64; x, y, workspace 16-byte aligned.
 * agenda_geglu: diffusers GEGLU after its projection: x [M, 2*inner] = [a | gate] -> y [M, inner] = a * gelu(gate)
 * (erf form), fp32 arithmetic, one rounding to bf16.  inner % 8 == 0. */
long long agenda_groupnorm_workspace_bytes(int B, int HW, int C, int G);
int agenda_groupnorm_nhwc(const void* x, const void* gamma, const void* beta, const void* pre_add, void* y, void* workspace,
                          int B, int HW, int C, int G, float eps, int silu, void* stream);
int agenda_geglu(const void* x, void* y, long long M, int inner, void* stream);
/* y[r, c] = h[r, c] + bias[c] + res[r, c]: the tail of ResnetBlock2D (conv2 bias + skip connection) in one pass.
 * h, res, y bf16 [rows, C] (channels-last memory), bias bf16 [C] or NULL.  C % 8 == 0; 16-byte aligned. */
int agenda_add_bias_residual(const void* h, const void* bias, const void* res, void* y, long long rows, int C,
                             void* stream);
/* torch.nn.LayerNorm(C, eps) over the last dim of x [M, C] bf16 (the three norms of diffusers' BasicTransformerBlock):
 * mean and centred variance in fp32, y = (x - mean) * rstd * gamma + beta, one rounding to bf16.  gamma / beta bf16 [C]
 * (NULL: 1 / 0).  C % 8 == 0; pointers 16-byte aligned. */
int agenda_layernorm(const void* x, const void* gamma, const void* beta, void* y, long long M, int C, float eps,
                     void* stream);

/* ---- a9: threshold / connected components / bbox (NOT in the reference; spec SURVEY.md §8 a9) ------------
 * Per map: nrm=(h-min)/((max-min)+1e-8f) fp32; mask = nrm > thr; 4-connectivity; labels 1..K in raster order
 * of each component's first pixel (scipy.ndimage.label numbering); boxes[k] = {x, y, w, h, area}.
 * heat fp32 [n,H,W]; labels int32 [n,H,W] (may be NULL); counts int32 [n] (= K, even when K > max_boxes);
 * boxes int32 [n,max_boxes,5] (only the first min(K,max_boxes) rows are written).
 * Default: ONE CTA per map (W % 32 == 0, bit mask <= 12288 words, i.e. up to ~640x600, 16-byte aligned heat / labels,
 * counts given): the map is streamed once for min / max and a per-word range table, only the words whose range straddles
 * the threshold are read again, labelling runs on the bit mask in shared memory.  Other shapes, and maps whose pieces
 * overflow the CTA's table, take one thread-block cluster per map with the fp32 map resident in distributed shared
 * memory: H*W*4 bytes must fit 16 CTAs x ~200 KB; otherwise AGENDA_ERR_UNSUPPORTED. */
int agenda_ccl_bbox(const float* heat, float thr, int32_t* labels, int32_t* counts, int32_t* boxes,
                    int max_boxes, int n, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AGENDA_B200_H_ */
