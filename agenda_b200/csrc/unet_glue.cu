// Glue kernels for the rest of the denoising step (SURVEY.md §8 f N4; data_generation/data_generation.py:59 runs a
// diffusers UNet2DConditionModel around the attention processor): the element-wise hot spots of the UNet skeleton that
// the library path handled worst (profiles/r02_unet_step_torch_profiler.txt: GroupNorm + the NCHW <-> NHWC copies around
// it 28 % of the step, GEGLU 11 %, LayerNorm 7 %).  With them the whole denoising step went from 33.4 to ~21 ms.
//
//   agenda_groupnorm_nhwc   GroupNorm(G) (+ SiLU) on a channels-last bf16 tensor, output channels-last: no layout copy on
//                           either side of the cuDNN convolutions / the transformer's [B, HW, C] view.
//   agenda_geglu            out = a * gelu(gate) for x = [a | gate] (exact erf form, torch.nn.functional.gelu's default).
//   agenda_layernorm        LayerNorm over the last dim, one warp per row, the row in registers.
//
// GroupNorm: statistics in fp32, two kernels, no atomics (bit-reproducible): pass 1 writes per-(batch, slab, group)
// partial sums in a fixed order, pass 2 folds them (<= 64 slabs) and applies y = x * (rstd * gamma) + (beta - mean * rstd
// * gamma).  A thread owns ONE 8-channel vector column (16 bytes) and walks the pixels of its slab, so per-channel
// sums (pass 1) and scale / shift (pass 2) live in registers.  HBM bytes: x read twice (the second read mostly from L2),
// y written once.
#include <algorithm>

#include "common.cuh"

namespace agenda {

constexpr int kGnMaxSlabs = 64;

struct GnPlan {
  int vp;       // 8-channel vectors per pixel
  int rows;     // pixels processed side by side by one CTA
  int threads;  // vp * rows
  int slabs;    // CTAs per batch element
  int ppc;      // pixels per slab
};

static GnPlan gn_plan(int B, int HW, int C) {
  GnPlan p;
  p.vp = C / 8;
  p.rows = std::max(1, 256 / p.vp);
  p.threads = p.vp * p.rows;
  const int want = std::max(1, (4 * 148 + B - 1) / B);                 // ~4 CTAs per SM over the whole launch
  const int most = std::max(1, HW / (4 * p.rows));                     // at least 4 pixels per thread row
  p.slabs = std::min({want, most, kGnMaxSlabs});
  p.ppc = (HW + p.slabs - 1) / p.slabs;
  p.slabs = (HW + p.ppc - 1) / p.ppc;
  return p;
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}

// partial[b][slab][g] = (sum, sum of squares) over the slab's pixels and the group's channels
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ pre_add,
                                float2* __restrict__ partial, int HW, int C, int G, int vp, int rows, int ppc) {
  extern __shared__ float sm[];   // [rows][C][2]
  const int b = blockIdx.y, slab = blockIdx.x;
  const int v = threadIdx.x % vp, r = threadIdx.x / vp;
  const int p0 = slab * ppc, p1 = min(HW, p0 + ppc);
  float s[8], q[8], pa[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; pa[j] = 0.f; }
  if (pre_add) unpack8(__ldg(reinterpret_cast<const uint4*>(pre_add + static_cast<long long>(b) * C) + v), pa);
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * HW * C) + v;
  for (int p = p0 + r; p < p1; p += rows) {
    float f[8];
    unpack8(__ldg(src + static_cast<long long>(p) * vp), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float t = f[j] + pa[j]; s[j] += t; q[j] = fmaf(t, t, q[j]); }
  }
  float* mine = sm + (static_cast<long long>(r) * C + v * 8) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) { mine[2 * j] = s[j]; mine[2 * j + 1] = q[j]; }
  __syncthreads();
  if (threadIdx.x < G) {   // fixed summation order: channels of the group, then the pixel rows
    const int g = threadIdx.x, cg = C / G;
    float ss = 0.f, qq = 0.f;
    for (int rr = 0; rr < rows; ++rr)
      for (int c = g * cg; c < (g + 1) * cg; ++c) { ss += sm[(rr * C + c) * 2]; qq += sm[(rr * C + c) * 2 + 1]; }
    partial[(static_cast<long long>(b) * gridDim.x + slab) * G + g] = make_float2(ss, qq);
  }
}

template <bool kSilu>
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                                const __nv_bfloat16* __restrict__ beta, const __nv_bfloat16* __restrict__ pre_add,
                                const float2* __restrict__ partial, __nv_bfloat16* __restrict__ y, int HW, int C, int G,
                                float eps, int vp, int rows, int ppc) {
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.y, slab = blockIdx.x, slabs = gridDim.x;
  if (threadIdx.x < G) {
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < slabs; ++k) {
      const float2 t = partial[(static_cast<long long>(b) * slabs + k) * G + threadIdx.x];
      ss += t.x; qq += t.y;
    }
    const float inv_n = 1.0f / (static_cast<float>(HW) * static_cast<float>(C / G));
    const float mean = ss * inv_n;
    const float var = fmaxf(qq * inv_n - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int v = threadIdx.x % vp, r = threadIdx.x / vp;
  const int cg = C / G;
  float sc[8], sh[8], pa[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pa[j] = 0.f;
  if (pre_add) unpack8(__ldg(reinterpret_cast<const uint4*>(pre_add + static_cast<long long>(b) * C) + v), pa);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j, g = c / cg;
    const float ga = gamma ? __bfloat162float(gamma[c]) : 1.f, be = beta ? __bfloat162float(beta[c]) : 0.f;
    sc[j] = s_rstd[g] * ga;
    sh[j] = be + (pa[j] - s_mean[g]) * sc[j];   // (x + pre_add - mean) * rstd * gamma + beta
  }
  const int p0 = slab * ppc, p1 = min(HW, p0 + ppc);
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * HW * C) + v;
  uint4* dst = reinterpret_cast<uint4*>(y + static_cast<long long>(b) * HW * C) + v;
  for (int p = p0 + r; p < p1; p += rows) {
    float f[8];
    unpack8(__ldg(src + static_cast<long long>(p) * vp), f);
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float a = fmaf(f[j], sc[j], sh[j]), c = fmaf(f[j + 1], sc[j + 1], sh[j + 1]);
      if (kSilu) { a = a / (1.f + __expf(-a)); c = c / (1.f + __expf(-c)); }
      oh[j >> 1] = __floats2bfloat162_rn(a, c);
    }
    dst[static_cast<long long>(p) * vp] = o;
  }
}

// x [M, 2*inner] = [a | gate] -> y [M, inner] = a * gelu(gate); one thread per 8 outputs
__global__ void geglu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n_vec, int iv) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_vec; i += stride) {
    const long long m = i / iv;
    const int c = static_cast<int>(i - m * iv);
    const uint4* row = reinterpret_cast<const uint4*>(x) + m * 2 * iv;
    float a[8], g[8];
    unpack8(__ldg(row + c), a);
    unpack8(__ldg(row + iv + c), g);
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float g0 = 0.5f * g[j] * (1.f + erff(g[j] * 0.70710678118654752f));
      const float g1 = 0.5f * g[j + 1] * (1.f + erff(g[j + 1] * 0.70710678118654752f));
      oh[j >> 1] = __floats2bfloat162_rn(a[j] * g0, a[j + 1] * g1);
    }
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}

// y[r, c] = h[r, c] + bias[c] + res[r, c]  (the tail of diffusers' ResnetBlock2D: conv2 bias + skip connection), 8 per thread
__global__ void add_bias_residual_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ bias,
                                         const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ y,
                                         long long n_vec, int vp) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_vec; i += stride) {
    float a[8], r[8], bb[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h) + i), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(res) + i), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = 0.f;
    if (bias) unpack8(__ldg(reinterpret_cast<const uint4*>(bias) + (i % vp)), bb);
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 8; j += 2) oh[j >> 1] = __floats2bfloat162_rn(a[j] + bb[j] + r[j], a[j + 1] + bb[j + 1] + r[j + 1]);
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}

// LayerNorm over the last dim of x [M, C] bf16: one warp per row, the row held in registers (<= 5 vectors of 8 per lane,
// i.e. C <= 1280; longer rows are re-read), mean then centred variance in fp32, one rounding to bf16.
constexpr int kLnMaxVec = 5;
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                        const __nv_bfloat16* __restrict__ gamma,
                                                        const __nv_bfloat16* __restrict__ beta,
                                                        __nv_bfloat16* __restrict__ y, long long M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nv = C >> 3;
  for (long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; row < M; row += warps) {
    const uint4* src = reinterpret_cast<const uint4*>(x + row * C);
    uint4 keep[kLnMaxVec];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
        keep[k] = __ldg(src + v);
        float f[8];
        unpack8(keep[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[j];
      }
    }
    for (int v = lane + 32 * kLnMaxVec; v < nv; v += 32) {
      float f[8];
      unpack8(__ldg(src + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(C);
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      if (lane + 32 * k < nv) {
        float f[8];
        unpack8(keep[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; sq = fmaf(d, d, sq); }
      }
    }
    for (int v = lane + 32 * kLnMaxVec; v < nv; v += 32) {
      float f[8];
      unpack8(__ldg(src + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; sq = fmaf(d, d, sq); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / static_cast<float>(C) + eps);
    uint4* dst = reinterpret_cast<uint4*>(y + row * C);
    auto emit = [&](int v, const uint4& raw) {
      float f[8], ga[8], be[8];
      unpack8(raw, f);
      if (gamma) unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + v), ga);
      if (beta) unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + v), be);
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float a = (f[j] - mean) * rstd, c = (f[j + 1] - mean) * rstd;
        if (gamma) { a *= ga[j]; c *= ga[j + 1]; }
        if (beta) { a += be[j]; c += be[j + 1]; }
        oh[j >> 1] = __floats2bfloat162_rn(a, c);
      }
      dst[v] = o;
    };
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k)
      if (lane + 32 * k < nv) emit(lane + 32 * k, keep[k]);
    for (int v = lane + 32 * kLnMaxVec; v < nv; v += 32) emit(v, __ldg(src + v));
  }
}

}  // namespace agenda

using namespace agenda;

extern "C" long long agenda_groupnorm_workspace_bytes(int B, int HW, int C, int G) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || G > 64 || C % G != 0 || C % 8 != 0)
    return fail(AGENDA_ERR_BAD_SHAPE, "groupnorm_workspace_bytes: B=%d HW=%d C=%d G=%d", B, HW, C, G);
  return static_cast<long long>(B) * kGnMaxSlabs * G * 8;
}

extern "C" int agenda_groupnorm_nhwc(const void* x, const void* gamma, const void* beta, const void* pre_add, void* y,
                                     void* workspace, int B, int HW, int C, int G, float eps, int silu, void* stream) {
  const char* who = "groupnorm_nhwc";
  if (B == 0) return AGENDA_OK;
  if (!x || !y || !workspace) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (B < 0 || HW <= 0 || C <= 0 || G <= 0 || G > 64 || C % G != 0 || C % 8 != 0 || C > 8 * 1024 || B > 65535)
    return fail(AGENDA_ERR_BAD_SHAPE, "%s: B=%d HW=%d C=%d G=%d (C %% 8 == 0, C %% G == 0, G <= 64)", who, B, HW, C, G);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(workspace) |
       reinterpret_cast<uintptr_t>(pre_add)) & 15)
    return fail(AGENDA_ERR_MISALIGNED, "%s: x, y, workspace, pre_add must be 16-byte aligned", who);
  const GnPlan p = gn_plan(B, HW, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(p.rows) * C * 2 * sizeof(float);
  AGENDA_DYN_SMEM(gn_stats_kernel, smem);
  const dim3 grid(p.slabs, B);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  float2* part = static_cast<float2*>(workspace);
  const __nv_bfloat16* pa = static_cast<const __nv_bfloat16*>(pre_add);
  gn_stats_kernel<<<grid, p.threads, smem, st>>>(xb, pa, part, HW, C, G, p.vp, p.rows, p.ppc);
  AGENDA_LAUNCH_CHECK("gn_stats_kernel");
  const __nv_bfloat16* ga = static_cast<const __nv_bfloat16*>(gamma);
  const __nv_bfloat16* be = static_cast<const __nv_bfloat16*>(beta);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  if (silu) gn_apply_kernel<true><<<grid, p.threads, 0, st>>>(xb, ga, be, pa, part, yb, HW, C, G, eps, p.vp, p.rows, p.ppc);
  else gn_apply_kernel<false><<<grid, p.threads, 0, st>>>(xb, ga, be, pa, part, yb, HW, C, G, eps, p.vp, p.rows, p.ppc);
  AGENDA_LAUNCH_CHECK("gn_apply_kernel");
  return AGENDA_OK;
}

extern "C" int agenda_geglu(const void* x, void* y, long long M, int inner, void* stream) {
  const char* who = "geglu";
  if (M == 0) return AGENDA_OK;
  if (!x || !y) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (M < 0 || inner <= 0 || inner % 8 != 0) return fail(AGENDA_ERR_BAD_SHAPE, "%s: M=%lld inner=%d (inner %% 8 == 0)", who, M, inner);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15)
    return fail(AGENDA_ERR_MISALIGNED, "%s: x, y must be 16-byte aligned", who);
  const int iv = inner / 8;
  const long long n_vec = M * iv;
  const long long blocks = std::min<long long>((n_vec + 255) / 256, static_cast<long long>(num_sms()) * 16);
  geglu_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n_vec, iv);
  AGENDA_LAUNCH_CHECK("geglu_kernel");
  return AGENDA_OK;
}

extern "C" int agenda_layernorm(const void* x, const void* gamma, const void* beta, void* y, long long M, int C, float eps,
                                void* stream) {
  const char* who = "layernorm";
  if (M == 0) return AGENDA_OK;
  if (!x || !y) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (M < 0 || C <= 0 || C % 8 != 0) return fail(AGENDA_ERR_BAD_SHAPE, "%s: M=%lld C=%d (C %% 8 == 0)", who, M, C);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
       reinterpret_cast<uintptr_t>(beta)) & 15)
    return fail(AGENDA_ERR_MISALIGNED, "%s: x, y, gamma, beta must be 16-byte aligned", who);
  const long long blocks = std::min<long long>((M + 7) / 8, static_cast<long long>(num_sms()) * 8);
  layernorm_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gamma), static_cast<const __nv_bfloat16*>(beta),
      static_cast<__nv_bfloat16*>(y), M, C, eps);
  AGENDA_LAUNCH_CHECK("layernorm_kernel");
  return AGENDA_OK;
}

extern "C" int agenda_add_bias_residual(const void* h, const void* bias, const void* res, void* y, long long rows, int C,
                                        void* stream) {
  const char* who = "add_bias_residual";
  if (rows == 0) return AGENDA_OK;
  if (!h || !res || !y) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (rows < 0 || C <= 0 || C % 8 != 0) return fail(AGENDA_ERR_BAD_SHAPE, "%s: rows=%lld C=%d (C %% 8 == 0)", who, rows, C);
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(y) |
       reinterpret_cast<uintptr_t>(bias)) & 15)
    return fail(AGENDA_ERR_MISALIGNED, "%s: pointers must be 16-byte aligned", who);
  const int vp = C / 8;
  const long long n_vec = rows * vp;
  const long long blocks = std::min<long long>((n_vec + 255) / 256, static_cast<long long>(num_sms()) * 16);
  add_bias_residual_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(bias), static_cast<const __nv_bfloat16*>(res),
      static_cast<__nv_bfloat16*>(y), n_vec, vp);
  AGENDA_LAUNCH_CHECK("add_bias_residual_kernel");
  return AGENDA_OK;
}
