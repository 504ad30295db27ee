// Exact-precision (fp32, CUDA-core) attention: softmax(scale * Q K^T) V with flash-style online softmax, and —
// for cross-attention — the heat-map epilogue of the reference's `_unravel_attn` (data_generation/hook.py:28-56):
// the probabilities of the selected key tokens are summed over heads inside one CTA (deterministic, no atomics),
// divided by H, and written / accumulated as [B', T, N] fp32.
//
// This is the path for fp32 pipelines (the reference runs fp32 end to end) and the small layers
// the tcgen05 kernel does not cover; the heavy self-attention layers run on attn_sm100.cu.
// Layout: q [B,N,H*d], k/v [B,M,H*d], token-major (what to_q/to_k/to_v emit, hook.py:93,101-102).
#include "common.cuh"

namespace agenda {

constexpr int kTQ = 32;        // query rows per CTA
constexpr int kWarps = 8;
constexpr int kRPW = kTQ / kWarps;  // rows per warp
constexpr int kTK = 128;       // keys per shared-memory tile
constexpr int kMaxD = 160;
constexpr int kMaxDJ = kMaxD / 32;  // output columns per lane

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, bool CROSS>
__global__ void __launch_bounds__(kWarps * 32) attn_f32_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                               const T* __restrict__ v, T* __restrict__ out, int B,
                                                               int H, int N, int M, int d, float scale, TokenList tl,
                                                               int b_first, float* __restrict__ maps, int accumulate,
                                                               const float* __restrict__ mask, int mask_rows) {
  extern __shared__ __align__(16) float smem[];
  const int ldk = d + 1;  // padded K rows: lane <-> key reads are bank-conflict free
  float* sQ = smem;                   // [kTQ][d]
  float* sK = sQ + kTQ * d;           // [kTK][d+1]
  float* sV = sK + kTK * ldk;         // [kTK][d]
  float* sHeat = sV + kTK * d;        // [kTQ][T]  (cross only)

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int q0 = blockIdx.x * kTQ;
  const int C = H * d;
  const int dj = (d + 31) >> 5;
  const int b = CROSS ? blockIdx.y : blockIdx.y / H;
  const int h_begin = CROSS ? 0 : blockIdx.y % H;
  const int h_end = CROSS ? H : h_begin + 1;
  const int nT = tl.n;
  const bool want_heat = CROSS && maps != nullptr && b >= b_first;

  if (want_heat)
    for (int i = tid; i < kTQ * nT; i += blockDim.x) sHeat[i] = 0.f;

  for (int h = h_begin; h < h_end; ++h) {
    __syncthreads();  // previous head's readers are done with sQ/sK/sV
    for (int i = tid; i < kTQ * d; i += blockDim.x) {
      const int r = i / d, c = i - r * d;
      const int n = q0 + r;
      sQ[i] = n < N ? to_f32<T>(q[(static_cast<long long>(b) * N + n) * C + h * d + c]) : 0.f;
    }
    float m_run[kRPW], l_run[kRPW], o[kRPW][kMaxDJ], p[kRPW][kTK / 32];
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      m_run[r] = -INFINITY; l_run[r] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxDJ; ++j) o[r][j] = 0.f;
    }
    for (int kv0 = 0; kv0 < M; kv0 += kTK) {
      const int kl = min(kTK, M - kv0);
      __syncthreads();
      for (int i = tid; i < kl * d; i += blockDim.x) {
        const int r = i / d, c = i - r * d;
        const long long g = (static_cast<long long>(b) * M + kv0 + r) * C + h * d + c;
        sK[r * ldk + c] = to_f32<T>(k[g]);
        sV[r * d + c] = to_f32<T>(v[g]);
      }
      __syncthreads();
      // ---- S = scale * Q K^T for this warp's rows: lane <-> key (lane + 32 j) ----
      float s[kRPW][kTK / 32];
#pragma unroll
      for (int r = 0; r < kRPW; ++r)
#pragma unroll
        for (int j = 0; j < kTK / 32; ++j) s[r][j] = 0.f;
      const float* qrow = sQ + (wid * kRPW) * d;
      for (int c = 0; c < d; ++c) {
        float kv[kTK / 32];
#pragma unroll
        for (int j = 0; j < kTK / 32; ++j) kv[j] = (lane + 32 * j < kl) ? sK[(lane + 32 * j) * ldk + c] : 0.f;
#pragma unroll
        for (int r = 0; r < kRPW; ++r) {
          const float qv = qrow[r * d + c];
#pragma unroll
          for (int j = 0; j < kTK / 32; ++j) s[r][j] = fmaf(qv, kv[j], s[r][j]);
        }
      }
      // ---- online softmax ----
#pragma unroll
      for (int r = 0; r < kRPW; ++r) {
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < kTK / 32; ++j) {
          s[r][j] = (lane + 32 * j < kl) ? s[r][j] * scale : -INFINITY;
          if (mask != nullptr && lane + 32 * j < kl) {  // additive mask [B*H, 1 | N, M] (hook.py:92,108: baddbmm input)
            const int n = min(q0 + wid * kRPW + r, N - 1);
            s[r][j] += mask[(static_cast<long long>(b * H + h) * mask_rows + (mask_rows == 1 ? 0 : n)) * M + kv0 + lane + 32 * j];
          }
          tmax = fmaxf(tmax, s[r][j]);
        }
        tmax = warp_max(tmax);
        const float m_new = fmaxf(m_run[r], tmax);
        const float corr = expf(m_run[r] - m_new);  // exp(-inf) = 0 on the first tile
        float psum = 0.f;
#pragma unroll
        for (int j = 0; j < kTK / 32; ++j) {
          p[r][j] = expf(s[r][j] - m_new);
          psum += p[r][j];
        }
        psum = warp_sum(psum);
        l_run[r] = l_run[r] * corr + psum;
        m_run[r] = m_new;
#pragma unroll
        for (int j = 0; j < kMaxDJ; ++j) o[r][j] *= corr;
      }
      // ---- O += P V: lane <-> output column (lane + 32 j) ----
      for (int mI = 0; mI < kl; ++mI) {
        float vv[kMaxDJ];
#pragma unroll
        for (int j = 0; j < kMaxDJ; ++j) vv[j] = (j < dj && lane + 32 * j < d) ? sV[mI * d + lane + 32 * j] : 0.f;
#pragma unroll
        for (int r = 0; r < kRPW; ++r) {
          float pm = 0.f;
#pragma unroll
          for (int j = 0; j < kTK / 32; ++j)
            if ((mI >> 5) == j) pm = p[r][j];
          pm = __shfl_sync(0xffffffffu, pm, mI & 31);
#pragma unroll
          for (int j = 0; j < kMaxDJ; ++j) o[r][j] = fmaf(pm, vv[j], o[r][j]);
        }
      }
    }
    // ---- write O, feed the heat accumulators (single KV tile: p / l is the final probability) ----
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int row = wid * kRPW + r, n = q0 + row;
      const float inv_l = 1.0f / l_run[r];
      if (n < N) {
#pragma unroll
        for (int j = 0; j < kMaxDJ; ++j) {
          const int c = lane + 32 * j;
          if (j < dj && c < d)
            out[(static_cast<long long>(b) * N + n) * C + h * d + c] = from_f32<T>(o[r][j] * inv_l);
        }
      }
      if (want_heat) {
        for (int t = 0; t < nT; ++t) {
          const int tok = tl.idx[t];
          if ((tok & 31) == lane) {
            float pv = 0.f;
#pragma unroll
            for (int j = 0; j < kTK / 32; ++j)
              if ((tok >> 5) == j) pv = p[r][j];
            if (tl.per_head) {  // DAAM-style: one plane per (batch, head, token), no head mean
              if (n < N) {
                float* ptr = maps + ((static_cast<long long>(b - b_first) * H + h) * nT + t) * N + n;
                *ptr = accumulate ? (*ptr + pv * inv_l) : pv * inv_l;
              }
            } else {
              sHeat[row * nT + t] += pv * inv_l;  // this warp owns the row: no race
            }
          }
        }
      }
    }
  }
  if (want_heat && !tl.per_head) {
    __syncthreads();
    const float inv_h = 1.0f / static_cast<float>(H);  // .mean(dim=1) over heads, hook.py:55
    float* dst = maps + static_cast<long long>(b - b_first) * nT * N;
    for (int i = tid; i < nT * kTQ; i += blockDim.x) {
      const int t = i / kTQ, row = i - t * kTQ;  // row fastest: 128-B coalesced segments per token plane
      const int n = q0 + row;
      if (n < N) {
        const float val = sHeat[row * nT + t] * inv_h;
        float* ptr = dst + static_cast<long long>(t) * N + n;
        *ptr = accumulate ? (*ptr + val) : val;
      }
    }
  }
}

static size_t attn_f32_smem(int d, int T) {
  return sizeof(float) * (static_cast<size_t>(kTQ) * d + static_cast<size_t>(kTK) * (d + 1) +
                          static_cast<size_t>(kTK) * d + static_cast<size_t>(kTQ) * T);
}

template <typename T, bool CROSS>
static int launch_attn_f32(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, int d,
                           float scale, const TokenList& tl, int b_first, float* maps, int accumulate, void* stream,
                           const float* mask = nullptr, int mask_rows = 1) {
  const size_t smem = attn_f32_smem(d, CROSS ? tl.n : 0);
  auto kern = attn_f32_kernel<T, CROSS>;
  AGENDA_DYN_SMEM(kern, smem);
  dim3 grid((N + kTQ - 1) / kTQ, CROSS ? B : B * H);
  attn_f32_kernel<T, CROSS><<<grid, kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), static_cast<T*>(out), B, H, N, M,
      d, scale, tl, b_first, maps, accumulate, mask, mask_rows);
  AGENDA_LAUNCH_CHECK("attn_f32_kernel");
  return AGENDA_OK;
}

int attn_common_checks(const char* who, const void* q, const void* k, const void* v, void* out, int dtype, int B,
                       int H, int N, int M, int d) {
  if (!q || !k || !v || !out) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (dtype != AGENDA_F32 && dtype != AGENDA_BF16) return fail(AGENDA_ERR_UNSUPPORTED, "%s: dtype %d", who, dtype);
  if (B <= 0 || H <= 0 || N <= 0 || M <= 0 || d <= 0)
    return fail(AGENDA_ERR_BAD_SHAPE, "%s: B=%d H=%d N=%d M=%d d=%d", who, B, H, N, M, d);
  if (static_cast<long long>(B) * H > 65535) return fail(AGENDA_ERR_BAD_SHAPE, "%s: B*H > 65535", who);
  return AGENDA_OK;
}

int build_token_list(const char* who, const int32_t* token_idx, int T, int M, TokenList* tl) {
  if (T < 0 || T > kMaxTokens) return fail(AGENDA_ERR_BAD_SHAPE, "%s: T=%d (max %d)", who, T, kMaxTokens);
  if (token_idx == nullptr) {
    if (T != M) return fail(AGENDA_ERR_BAD_SHAPE, "%s: token_idx NULL means all tokens, so T must equal M", who);
    for (int i = 0; i < M; ++i) tl->idx[i] = i;
  } else {
    for (int i = 0; i < T; ++i) {
      if (token_idx[i] < 0 || token_idx[i] >= M)
        return fail(AGENDA_ERR_BAD_SHAPE, "%s: token_idx[%d]=%d outside [0,%d)", who, i, token_idx[i], M);
      tl->idx[i] = token_idx[i];
    }
  }
  tl->n = T;
  tl->per_head = 0;
  return AGENDA_OK;
}

int attn_cross_f32(const void* q, const void* k, const void* v, void* out, int dtype, int B, int H, int N, int M,
                   int d, float scale, const TokenList& tl, int b_first, float* maps, int accumulate, void* stream) {
  if (d > kMaxD) return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross (fp32 path): d=%d > %d", d, kMaxD);
  if (M > kTK) return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross: M=%d > %d", M, kTK);
  return dtype == AGENDA_F32
             ? launch_attn_f32<float, true>(q, k, v, out, B, H, N, M, d, scale, tl, b_first, maps, accumulate, stream)
             : launch_attn_f32<__nv_bfloat16, true>(q, k, v, out, B, H, N, M, d, scale, tl, b_first, maps, accumulate,
                                                    stream);
}

}  // namespace agenda

using namespace agenda;

extern "C" int agenda_attn_self_fwd_f32(const void* q, const void* k, const void* v, void* out, int dtype, int B,
                                        int H, int N, int d, float scale, void* stream) {
  int rc = attn_common_checks("attn_self_fwd_f32", q, k, v, out, dtype, B, H, N, N, d);
  if (rc != AGENDA_OK) return rc;
  if (d > kMaxD) return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd_f32: d=%d > %d", d, kMaxD);
  TokenList tl;
  tl.n = 0;
  tl.per_head = 0;
  return dtype == AGENDA_F32
             ? launch_attn_f32<float, false>(q, k, v, out, B, H, N, N, d, scale, tl, 0, nullptr, 0, stream)
             : launch_attn_f32<__nv_bfloat16, false>(q, k, v, out, B, H, N, N, d, scale, tl, 0, nullptr, 0, stream);
}

// Attention with an additive mask (hook.py:92: `attn.prepare_attention_mask` -> [B*H, 1 | N, M], added to the scaled
// logits by `get_attention_scores`' baddbmm, hook.py:108).  The SD UNets pass none; pipelines that do (padding masks
// of a text encoder) get the exact fp32 path: one kernel for self- and cross-attention, heat maps as in
// agenda_attn_cross_fwd_heat when `maps` is given (then M <= 128).  A fully masked row gives NaN, as in torch.
extern "C" int agenda_attn_fwd_masked(const void* q, const void* k, const void* v, void* out, int dtype, int B, int H, int N,
                                      int M, int d, float scale, const float* mask, int mask_rows,
                                      const int32_t* token_idx, int T, int b_first, int per_head, float* maps,
                                      int accumulate, void* stream) {
  int rc = attn_common_checks("attn_fwd_masked", q, k, v, out, dtype, B, H, N, M, d);
  if (rc != AGENDA_OK) return rc;
  if (d > kMaxD) return fail(AGENDA_ERR_UNSUPPORTED, "attn_fwd_masked: d=%d > %d", d, kMaxD);
  if (mask == nullptr) return fail(AGENDA_ERR_NULL_POINTER, "attn_fwd_masked: mask is NULL (use the unmasked entry points)");
  if (mask_rows != 1 && mask_rows != N)
    return fail(AGENDA_ERR_BAD_SHAPE, "attn_fwd_masked: mask_rows=%d must be 1 (broadcast over queries) or N=%d", mask_rows, N);
  TokenList tl;
  tl.n = 0;
  tl.per_head = 0;
  if (maps != nullptr) {
    if (M > kTK) return fail(AGENDA_ERR_UNSUPPORTED, "attn_fwd_masked: heat maps need M=%d <= %d", M, kTK);
    if (b_first < 0 || b_first >= B) return fail(AGENDA_ERR_BAD_SHAPE, "attn_fwd_masked: b_first=%d outside [0,%d)", b_first, B);
    if ((rc = build_token_list("attn_fwd_masked", token_idx, T, M, &tl)) != AGENDA_OK) return rc;
    tl.per_head = per_head ? 1 : 0;
    return dtype == AGENDA_F32 ? launch_attn_f32<float, true>(q, k, v, out, B, H, N, M, d, scale, tl, b_first, maps, accumulate,
                                                              stream, mask, mask_rows)
                               : launch_attn_f32<__nv_bfloat16, true>(q, k, v, out, B, H, N, M, d, scale, tl, b_first, maps,
                                                                      accumulate, stream, mask, mask_rows);
  }
  return dtype == AGENDA_F32
             ? launch_attn_f32<float, false>(q, k, v, out, B, H, N, M, d, scale, tl, 0, nullptr, 0, stream, mask, mask_rows)
             : launch_attn_f32<__nv_bfloat16, false>(q, k, v, out, B, H, N, M, d, scale, tl, 0, nullptr, 0, stream, mask,
                                                     mask_rows);
}
