// agenda_attn_cross_fwd_heat: cross-attention + heat-map epilogue dispatcher (data_generation/hook.py:108-114 and
// _unravel_attn, hook.py:28-56).
#include "common.cuh"

namespace agenda {
int attn_common_checks(const char* who, const void* q, const void* k, const void* v, void* out, int dtype, int B,
                       int H, int N, int M, int d);
int build_token_list(const char* who, const int32_t* token_idx, int T, int M, TokenList* tl);
int attn_cross_f32(const void* q, const void* k, const void* v, void* out, int dtype, int B, int H, int N, int M,
                   int d, float scale, const TokenList& tl, int b_first, float* maps, int accumulate, void* stream);
int attn_cross_sm100(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, int d,
                     float scale, const TokenList& tl, int b_first, float* maps, int accumulate, void* stream);
}  // namespace agenda

using namespace agenda;

extern "C" int agenda_attn_cross_fwd_heat(const void* q, const void* k, const void* v, void* out, int dtype, int B,
                                          int H, int N, int M, int d, float scale, const int32_t* token_idx, int T,
                                          int b_first, float* maps, int accumulate, void* stream) {
  int rc = attn_common_checks("attn_cross_fwd_heat", q, k, v, out, dtype, B, H, N, M, d);
  if (rc != AGENDA_OK) return rc;
  if (b_first < 0 || b_first > B) return fail(AGENDA_ERR_BAD_SHAPE, "attn_cross_fwd_heat: b_first=%d, B=%d", b_first, B);
  TokenList tl;
  tl.n = 0;
  tl.per_head = 0;
  if (maps != nullptr) {
    rc = build_token_list("attn_cross_fwd_heat", token_idx, T, M, &tl);
    if (rc != AGENDA_OK) return rc;
    if (reinterpret_cast<uintptr_t>(maps) & 3) return fail(AGENDA_ERR_MISALIGNED, "attn_cross_fwd_heat: maps");
  }
  float* mp = tl.n ? maps : nullptr;
  // bf16 activations: tcgen05 tensor-core kernel (products of bf16 operands are exact in its fp32 accumulators, so
  // the heat maps keep fp32-softmax accuracy).  fp32 activations: the exact fp32 CUDA-core kernel.
  if (dtype == AGENDA_BF16 && M <= 80 && (d == 40 || d == 64 || d == 80 || d == 160))
    return attn_cross_sm100(q, k, v, out, B, H, N, M, d, scale, tl, b_first, mp, accumulate, stream);
  return attn_cross_f32(q, k, v, out, dtype, B, H, N, M, d, scale, tl, b_first, mp, accumulate, stream);
}

// DAAM-style capture: per-head probability planes maps[b', head, t, n] (no head mean) — see include/agenda_b200.h.
extern "C" int agenda_attn_cross_fwd_heat_heads(const void* q, const void* k, const void* v, void* out, int dtype, int B,
                                                int H, int N, int M, int d, float scale, const int32_t* token_idx,
                                                int T, int b_first, float* maps, int accumulate, void* stream) {
  int rc = attn_common_checks("attn_cross_fwd_heat_heads", q, k, v, out, dtype, B, H, N, M, d);
  if (rc != AGENDA_OK) return rc;
  if (b_first < 0 || b_first > B) return fail(AGENDA_ERR_BAD_SHAPE, "attn_cross_fwd_heat_heads: b_first=%d, B=%d", b_first, B);
  if (maps == nullptr) return fail(AGENDA_ERR_NULL_POINTER, "attn_cross_fwd_heat_heads: maps is null");
  if (reinterpret_cast<uintptr_t>(maps) & 3) return fail(AGENDA_ERR_MISALIGNED, "attn_cross_fwd_heat_heads: maps");
  TokenList tl;
  rc = build_token_list("attn_cross_fwd_heat_heads", token_idx, T, M, &tl);
  if (rc != AGENDA_OK) return rc;
  tl.per_head = 1;
  float* mp = tl.n ? maps : nullptr;
  if (dtype == AGENDA_BF16 && M <= 80 && tl.n <= 8 && (d == 40 || d == 64 || d == 80 || d == 160))
    return attn_cross_sm100(q, k, v, out, B, H, N, M, d, scale, tl, b_first, mp, accumulate, stream);
  return attn_cross_f32(q, k, v, out, dtype, B, H, N, M, d, scale, tl, b_first, mp, accumulate, stream);
}

// Test hook: force the fp32 CUDA-core kernel regardless of dtype.
extern "C" int agenda_attn_cross_fwd_heat_f32(const void* q, const void* k, const void* v, void* out, int dtype, int B,
                                              int H, int N, int M, int d, float scale, const int32_t* token_idx,
                                              int T, int b_first, float* maps, int accumulate, void* stream) {
  int rc = attn_common_checks("attn_cross_fwd_heat_f32", q, k, v, out, dtype, B, H, N, M, d);
  if (rc != AGENDA_OK) return rc;
  if (b_first < 0 || b_first > B) return fail(AGENDA_ERR_BAD_SHAPE, "attn_cross_fwd_heat_f32: b_first=%d", b_first);
  TokenList tl;
  tl.n = 0;
  tl.per_head = 0;
  if (maps != nullptr) {
    rc = build_token_list("attn_cross_fwd_heat_f32", token_idx, T, M, &tl);
    if (rc != AGENDA_OK) return rc;
  }
  return attn_cross_f32(q, k, v, out, dtype, B, H, N, M, d, scale, tl, b_first, tl.n ? maps : nullptr, accumulate,
                        stream);
}
