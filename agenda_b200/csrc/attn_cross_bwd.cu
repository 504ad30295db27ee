// Backward of the cross-attention + heat epilogue (training mode, SURVEY.md §8 f N3; the reference reaches it through
// autograd of data_generation/hook.py:104-115 and _unravel_attn hook.py:28-56 from finetune_sd_token.py:1043-1069).
//
//   P  = softmax(scale * Q K^T)                       (recomputed: M <= 96 keys, the 77-token prompt)
//   dV = P^T dO
//   dP = dO V^T  (+ d_maps[b-b_first, t, n] / H on column token_idx[t] for b >= b_first: maps = mean over heads)
//   dS = P * (dP - rowsum(P * dP))
//   dQ = scale * dS K,      dK = scale * dS^T Q
//
// Exact fp32 CUDA-core kernel (first native version; the work is 5 * 2 * N * 77 * d FLOP per (batch, head) — small
// next to the self-attention layers — so it is bound by reading Q / dO and writing dQ once).  CTA = (32 query rows,
// one (batch, head)); K_h, V_h, the Q and dO tiles and the tile's P / dS live in shared memory as fp32.
//   phase 1 (warp per 4 rows, taken together): scores and dP with lanes over keys — every K / V element read from
//           shared memory feeds 4 FMAs — softmax by warp shuffles, dQ with lanes over channels
//   phase 2 (thread per (key, 4 channels)): the tile's contribution to dK / dV, added to the fp32 outputs with atomicAdd
// dk / dv are fp32 [B, M, H*d] accumulators the CALLER zero-fills (128 CTAs per (batch, head) add into them; the
// order of those additions is not fixed, so dK / dV can differ in the last bits between runs).
#include "common.cuh"

namespace agenda {

int build_token_list(const char* who, const int32_t* token_idx, int T, int M, TokenList* tl);

namespace {

constexpr int kBwdRows = 32;     // query rows per CTA
constexpr int kBwdThreads = 256;
constexpr int kBwdMaxKeys = 96;  // three keys per lane

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_f(float* p, float x) { *p = x; }
__device__ __forceinline__ void from_f(__nv_bfloat16* p, float x) { *p = __float2bfloat16(x); }

constexpr int kBwdRowsPerWarp = kBwdRows / (kBwdThreads / 32);  // 4 rows per warp, processed together

__host__ __device__ inline int bwd_align4(int x) { return (x + 3) & ~3; }

template <typename T>
__global__ void __launch_bounds__(kBwdThreads)
attn_cross_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                      const T* __restrict__ d_out, const float* __restrict__ d_maps, T* __restrict__ dq,
                      float* __restrict__ dk, float* __restrict__ dv, const TokenList tl, int H, int N, int M, int d,
                      int b_first, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ldk = d + 1;   // K / V rows padded: lanes walk keys, so consecutive rows must hit different banks
  const int mp = M + 1;
  float* Ks = reinterpret_cast<float*>(smem_raw);   // [M][d+1]
  float* Vs = Ks + bwd_align4(M * ldk);             // [M][d+1]
  float* Qs = Vs + bwd_align4(M * ldk);             // [32][d]   (16-byte aligned: float4 reads in phase 2)
  float* Os = Qs + bwd_align4(kBwdRows * d);        // [32][d]   dO tile
  float* DS = Os + bwd_align4(kBwdRows * d);        // [32][M+1] dS * scale
  float* PS = DS + kBwdRows * mp;                   // [32][M+1] P
  float* GS = PS + kBwdRows * mp;                   // [8 warps][M+1] heat gradient per key of the warp's current row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y / H, h = blockIdx.y - b * H;
  const int n0 = blockIdx.x * kBwdRows;
  const int C = H * d;
  const long long kv_base = static_cast<long long>(b) * M * C + h * d;
  const long long q_base = static_cast<long long>(b) * N * C + h * d;

  for (int i = tid; i < M * d; i += kBwdThreads) {
    const int j = i / d, c = i - j * d;
    Ks[j * ldk + c] = to_f(k[kv_base + static_cast<long long>(j) * C + c]);
    Vs[j * ldk + c] = to_f(v[kv_base + static_cast<long long>(j) * C + c]);
  }
  for (int i = tid; i < kBwdRows * d; i += kBwdThreads) {
    const int r = i / d, c = i - r * d;
    const bool ok = n0 + r < N;
    Qs[i] = ok ? to_f(q[q_base + static_cast<long long>(n0 + r) * C + c]) : 0.f;
    Os[i] = ok ? to_f(d_out[q_base + static_cast<long long>(n0 + r) * C + c]) : 0.f;
  }
  __syncthreads();

  // ---- phase 1: a warp takes 4 query rows TOGETHER (every K / V element read from shared memory feeds 4 FMAs) ----
  const bool heat = (d_maps != nullptr) && (b >= b_first) && tl.n > 0;
  const float inv_h = 1.0f / static_cast<float>(H);
  float* gs = GS + warp * mp;
  const int rbase = warp * kBwdRowsPerWarp;
  float sc[kBwdRowsPerWarp][3], dp[kBwdRowsPerWarp][3];
#pragma unroll
  for (int r = 0; r < kBwdRowsPerWarp; ++r)
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) { sc[r][jj] = 0.f; dp[r][jj] = 0.f; }
  int jrow[3];  // this lane's keys (clamped: lanes past M recompute the last key and drop the result)
#pragma unroll
  for (int jj = 0; jj < 3; ++jj) jrow[jj] = min(lane + 32 * jj, M - 1) * ldk;
  for (int c = 0; c < d; ++c) {
    float qv[kBwdRowsPerWarp], ov[kBwdRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kBwdRowsPerWarp; ++r) { qv[r] = Qs[(rbase + r) * d + c]; ov[r] = Os[(rbase + r) * d + c]; }
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      const float kk = Ks[jrow[jj] + c], vv = Vs[jrow[jj] + c];
#pragma unroll
      for (int r = 0; r < kBwdRowsPerWarp; ++r) {
        sc[r][jj] = fmaf(qv[r], kk, sc[r][jj]);
        dp[r][jj] = fmaf(ov[r], vv, dp[r][jj]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kBwdRowsPerWarp; ++r) {
    const int n = n0 + rbase + r;
    float* ds_row = DS + (rbase + r) * mp;
    float* p_row = PS + (rbase + r) * mp;
    if (n >= N) {  // padding row: contributes nothing to dK / dV (warp-uniform branch)
      for (int j = lane; j < M; j += 32) { ds_row[j] = 0.f; p_row[j] = 0.f; }
      continue;
    }
    float s[3], g[3];
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      s[jj] = (lane + 32 * jj < M) ? sc[r][jj] * scale : -INFINITY;
      g[jj] = dp[r][jj];
    }
    if (heat) {  // d_maps / H lands on the selected key columns (a token listed twice gets both planes)
      for (int j = lane; j < M; j += 32) gs[j] = 0.f;
      __syncwarp();
      const float* gm = d_maps + (static_cast<long long>(b - b_first) * tl.n) * N + n;
      for (int t = lane; t < tl.n; t += 32) atomicAdd(&gs[tl.idx[t]], gm[static_cast<long long>(t) * N] * inv_h);
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) {
        const int j = lane + 32 * jj;
        if (j < M) g[jj] += gs[j];
      }
      __syncwarp();  // gs is reused by the next row
    }
    const float mx = warp_max(fmaxf(fmaxf(s[0], s[1]), s[2]));
    float e[3], sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      e[jj] = (lane + 32 * jj < M) ? __expf(s[jj] - mx) : 0.f;
      sum += e[jj];
    }
    const float inv_sum = 1.0f / warp_sum(sum);
    float delta = 0.f;
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      e[jj] *= inv_sum;  // P
      delta = fmaf(e[jj], g[jj], delta);
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      const int j = lane + 32 * jj;
      if (j < M) {
        ds_row[j] = e[jj] * (g[jj] - delta) * scale;
        p_row[j] = e[jj];
      }
    }
  }
  __syncwarp();
  // dQ rows of this warp: lanes over channels, the 4 rows share every K element
  for (int c = lane; c < d; c += 32) {
    float acc[kBwdRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kBwdRowsPerWarp; ++r) acc[r] = 0.f;
    for (int j = 0; j < M; ++j) {
      const float kk = Ks[j * ldk + c];
#pragma unroll
      for (int r = 0; r < kBwdRowsPerWarp; ++r) acc[r] = fmaf(DS[(rbase + r) * mp + j], kk, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kBwdRowsPerWarp; ++r) {
      const int n = n0 + rbase + r;
      if (n < N) from_f(dq + q_base + static_cast<long long>(n) * C + c, acc[r]);
    }
  }
  __syncthreads();

  // ---- phase 2: the tile's dK / dV; a thread owns (key j, 4 consecutive channels) ----
  const int d4 = (d + 3) >> 2;
  const bool vec = (d & 3) == 0;
  for (int i = tid; i < M * d4; i += kBwdThreads) {
    const int j = i / d4, c0 = (i - j * d4) * 4;
    float ak[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {
#pragma unroll 4
      for (int r = 0; r < kBwdRows; ++r) {
        const float ds = DS[r * mp + j], ps = PS[r * mp + j];
        const float4 qv = *reinterpret_cast<const float4*>(Qs + r * d + c0);
        const float4 ov = *reinterpret_cast<const float4*>(Os + r * d + c0);
        ak[0] = fmaf(ds, qv.x, ak[0]); ak[1] = fmaf(ds, qv.y, ak[1]); ak[2] = fmaf(ds, qv.z, ak[2]); ak[3] = fmaf(ds, qv.w, ak[3]);
        av[0] = fmaf(ps, ov.x, av[0]); av[1] = fmaf(ps, ov.y, av[1]); av[2] = fmaf(ps, ov.z, av[2]); av[3] = fmaf(ps, ov.w, av[3]);
      }
    } else {
      for (int r = 0; r < kBwdRows; ++r) {
        const float ds = DS[r * mp + j], ps = PS[r * mp + j];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c0 + u < d) {
            ak[u] = fmaf(ds, Qs[r * d + c0 + u], ak[u]);
            av[u] = fmaf(ps, Os[r * d + c0 + u], av[u]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c0 + u < d) {
        atomicAdd(dk + kv_base + static_cast<long long>(j) * C + c0 + u, ak[u]);
        atomicAdd(dv + kv_base + static_cast<long long>(j) * C + c0 + u, av[u]);
      }
    }
  }
}

template <typename T>
int launch_bwd(const void* q, const void* k, const void* v, const void* d_out, const float* d_maps, void* dq, float* dk,
               float* dv, const TokenList& tl, int B, int H, int N, int M, int d, int b_first, float scale,
               cudaStream_t stream) {
  const size_t smem = sizeof(float) * (2 * static_cast<size_t>(bwd_align4(M * (d + 1))) +
                                       2 * static_cast<size_t>(bwd_align4(kBwdRows * d)) +
                                       2 * static_cast<size_t>(kBwdRows) * (M + 1) + (kBwdThreads / 32) * (M + 1));
  if (smem > 200 * 1024)
    return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross_bwd: M=%d, d=%d needs %zu B of shared memory (> 200 KB)", M, d, smem);
  auto kern = attn_cross_bwd_kernel<T>;
  AGENDA_DYN_SMEM(kern, smem);
  dim3 grid((N + kBwdRows - 1) / kBwdRows, B * H);
  kern<<<grid, kBwdThreads, smem, stream>>>(static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v),
                                            static_cast<const T*>(d_out), d_maps, static_cast<T*>(dq), dk, dv, tl, H, N,
                                            M, d, b_first, scale);
  AGENDA_LAUNCH_CHECK("attn_cross_bwd_kernel");
  return AGENDA_OK;
}

}  // namespace
}  // namespace agenda

using namespace agenda;

extern "C" int agenda_attn_cross_bwd(const void* q, const void* k, const void* v, const void* d_out,
                                     const float* d_maps, void* dq, float* dk, float* dv, int dtype, int B, int H,
                                     int N, int M, int d, float scale, const int32_t* token_idx, int T, int b_first,
                                     void* stream) {
  if (!q || !k || !v || !d_out || !dq || !dk || !dv) return fail(AGENDA_ERR_NULL_POINTER, "attn_cross_bwd: null pointer");
  if (B <= 0 || H <= 0 || N <= 0 || M <= 0 || d <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "attn_cross_bwd: bad shape");
  if (M > kBwdMaxKeys) return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross_bwd: M=%d keys (max %d)", M, kBwdMaxKeys);
  if (b_first < 0 || b_first > B) return fail(AGENDA_ERR_BAD_SHAPE, "attn_cross_bwd: b_first=%d, B=%d", b_first, B);
  if (static_cast<long long>(B) * H > 65535) return fail(AGENDA_ERR_BAD_SHAPE, "attn_cross_bwd: B*H=%lld > 65535", static_cast<long long>(B) * H);
  TokenList tl;
  tl.n = 0;
  tl.per_head = 0;
  if (d_maps != nullptr) {
    int rc = build_token_list("attn_cross_bwd", token_idx, T, M, &tl);
    if (rc != AGENDA_OK) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == AGENDA_BF16)
    return launch_bwd<__nv_bfloat16>(q, k, v, d_out, d_maps, dq, dk, dv, tl, B, H, N, M, d, b_first, scale, st);
  if (dtype == AGENDA_F32)
    return launch_bwd<float>(q, k, v, d_out, d_maps, dq, dk, dv, tl, B, H, N, M, d, b_first, scale, st);
  return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross_bwd: dtype %d", dtype);
}
