// K2: fused cross-attention + heat-map epilogue for sm_100a (tcgen05 / TMEM / TMA).
//
//   out  = softmax(scale * Q K^T) V                                    (data_generation/hook.py:108,114)
//   maps[b', t, n] (=|+=) mean_h softmax(...)[b, h, n, token_idx[t]]    (_unravel_attn, hook.py:28-56)
//
// The key axis is the 77-token prompt: ONE KV tile (padded to 80 by TMA zero fill), so there is no online softmax.
// CTA = one (batch element, 128-query tile) and loops over ALL heads, which makes the head mean a deterministic
// in-register sum (no atomics, no [B*H,N,77] probability tensor, no 77-iteration Python loop):
//   warps 0-3  softmax warpgroup, thread r <-> query row r: S row (80 fp32) from TMEM, exp2, normalise, add the row
//              into a per-thread 80-float heat accumulator, write P (bf16) back into TMEM over S; then drain the
//              previous head's O (TMEM -> bf16 -> global) while the tensor core runs this head's P V.
//   warp 4     TMA producer: (Q_h, K_h, V_h) per head through a 2/3-stage mbarrier ring
//   warp 5     TMEM allocator + tcgen05.mma issuer: QK_0, QK_1, PV_0, QK_2, PV_1, ...
// TMEM: S/P[2] (96 columns each) | O[2] (round16(d) columns each).
// At the end each thread parks its accumulator in its own shared-memory row and emits the selected token columns:
// consecutive threads = consecutive pixels, so every token plane is written with coalesced 128-byte stores.
#include <cooperative_groups.h>
#include <cstdlib>

#include "sm100_common.cuh"

namespace agenda {


namespace sm100 {

namespace cg = cooperative_groups;

constexpr int kXBlockM = 128;
constexpr int kXThreads = 192;
constexpr int kXMPad = 80;      // 77 prompt tokens padded to a multiple of 16
constexpr int kXSlot = 80;      // TMEM columns per S/P buffer (S: 80 fp32 columns; P: 48 packed columns over it)
constexpr int kXFewTokens = 8;  // heat maps for at most this many tokens take the register-light path
constexpr int kHeatLd = 81;     // padded accumulator row in shared memory (bank-conflict free)

template <int D>
struct XCfg {
  static constexpr int kDP = (D + 15) / 16 * 16;
  // TMEM columns actually needed, rounded to the power of two tcgen05.alloc wants: 256 at d = 40, so two CTAs fit an SM
  static constexpr int kTmemCols = (2 * kXSlot + 2 * kDP <= 256) ? 256 : 512;
  static constexpr int kChunks = (D + 63) / 64;
  static constexpr int kStages = (D <= 80) ? 3 : 2;
  static constexpr int kQBytes = kChunks * kXBlockM * 128;
  static constexpr int kKVBytes = kChunks * kXMPad * 128;
  static constexpr int kStageBytes = kQBytes + 2 * kKVBytes;
  static constexpr int kColO = 2 * kXSlot;
  static_assert(kColO + 2 * kDP <= 512, "TMEM overflow");
  static_assert(kStages * kStageBytes >= kXBlockM * kHeatLd * 4, "heat staging does not fit");
};

struct XBarriers {
  uint64_t in_full[3], in_empty[3];
  uint64_t s_full[2], p_full[2], pv_done[2], o_free[2];
  uint32_t tmem_base;
};

template <int D>
constexpr size_t x_smem_bytes() {
  return 1024 + XCfg<D>::kStages * XCfg<D>::kStageBytes + sizeof(XBarriers) + 64;
}

// kFew: heat maps are wanted for <= kXFewTokens key tokens (what every caller of the reference reads,
// data_generation.py:74-77): only those columns are accumulated (the raw scores are fetched again from TMEM with
// one-column loads and pushed through the same fp32 ops, so the values are bit-identical to the all-token path).
// That frees ~70 registers per thread, and with 256 TMEM columns two CTAs share an SM at d = 40.
template <int D, bool kFew>
__global__ void __launch_bounds__(kXThreads, (kFew && XCfg<D>::kTmemCols == 256) ? 2 : 1)
attn_cross_sm100_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ out,
                        float* __restrict__ maps, const TokenList tl, int H, int N, int M, int b_first,
                        int accumulate, float scale_log2) {
  using C = XCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  XBarriers* bars = reinterpret_cast<XBarriers*>(smem + C::kStages * C::kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * kXBlockM;
  const int b = blockIdx.y;
  const bool want_heat = (maps != nullptr) && (b >= b_first);
  // Small layers (N <= 256: 16-32 (batch, query tile) pairs on 148 SMs) are launched as clusters of gridDim.z CTAs
  // along z, each taking H / gridDim.z heads; the head sum of the heat map is then finished through the leader's
  // shared memory in a fixed order (deterministic, no atomics).  gridDim.z = 1 is the plain one-CTA-per-tile kernel.
  const int hpg = H / static_cast<int>(gridDim.z);  // heads of this CTA
  const int h_begin = static_cast<int>(blockIdx.z) * hpg;

  if (tid == 4 * 32) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int s = 0; s < 3; ++s) { mbar_init(&bars->in_full[s], 1); mbar_init(&bars->in_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s_full[s], 1); mbar_init(&bars->p_full[s], 128);
      mbar_init(&bars->pv_done[s], 1); mbar_init(&bars->o_free[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(&bars->tmem_base, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  constexpr int kAcc = kFew ? kXFewTokens : kXMPad;
  float acc[kAcc];  // softmax threads: per-row heat accumulator over this CTA's heads
  if (warp == 4) {
    // ============================== TMA producer (warp converged, one elected lane issues) ==============================
    int st = 0;
    uint32_t ph = 0;
    for (int hl = 0; hl < hpg; ++hl) {
      const int h = h_begin + hl;
      unsigned char* sQ = smem + st * C::kStageBytes;
      unsigned char* sK = sQ + C::kQBytes;
      unsigned char* sV = sK + C::kKVBytes;
      mbar_wait(&bars->in_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&bars->in_full[st], C::kStageBytes);
        for (int c = 0; c < C::kChunks; ++c) {
          tma_load_4d(&map_q, &bars->in_full[st], sQ + c * kXBlockM * 128, c * 64, h, q0, b);
          tma_load_4d(&map_k, &bars->in_full[st], sK + c * kXMPad * 128, c * 64, h, 0, b);
          tma_load_4d(&map_v, &bars->in_full[st], sV + c * kXMPad * 128, c * 64, h, 0, b);
        }
      }
      __syncwarp();
      if (++st == C::kStages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 5) {
    // ============================== MMA issuer (warp converged; tcgen05.mma / commit under elect.sync) ==============================
    constexpr uint32_t idesc_qk = make_idesc(kXBlockM, kXMPad, 0);
    constexpr uint32_t idesc_pv = make_idesc(kXBlockM, C::kDP, 1);
    const uint64_t q_desc0 = make_sdesc(smem_u32(smem), 16, 1024);
    const uint64_t k_desc0 = make_sdesc(smem_u32(smem + C::kQBytes), 16, 1024);
    const uint64_t v_desc0 = make_sdesc(smem_u32(smem + C::kQBytes + C::kKVBytes), kXMPad * 128, 1024);
    auto issue_pv = [&](int h, int st) {
      const int sb = h & 1;
      const uint32_t ph = (h >> 1) & 1;
      mbar_wait(&bars->p_full[sb], ph);
      mbar_wait(&bars->o_free[sb], ph ^ 1);  // O[sb] drained by the epilogue of head h-2
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kXMPad / 16; ++kk) {
          const uint64_t bdesc = v_desc0 + static_cast<uint64_t>((st * C::kStageBytes + kk * 2048) >> 4);
          umma_ts(tmem + C::kColO + sb * C::kDP, tmem + sb * kXSlot + kk * 8, bdesc, idesc_pv, kk != 0);
        }
        umma_commit(&bars->in_empty[st]);
        umma_commit(&bars->pv_done[sb]);
      }
      __syncwarp();
    };
    int st = 0, st_prev = 0;
    uint32_t ph = 0;
    for (int h = 0; h < hpg; ++h) {  // (local head index: only buffer parities depend on it here)
      const int sb = h & 1;
      mbar_wait(&bars->in_full[st], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < C::kDP / 16; ++kk) {
          const uint64_t adesc = q_desc0 + static_cast<uint64_t>((st * C::kStageBytes + (kk >> 2) * kXBlockM * 128 + (kk & 3) * 32) >> 4);
          const uint64_t bdesc = k_desc0 + static_cast<uint64_t>((st * C::kStageBytes + (kk >> 2) * kXMPad * 128 + (kk & 3) * 32) >> 4);
          umma_ss(tmem + sb * kXSlot, adesc, bdesc, idesc_qk, kk != 0);
        }
        umma_commit(&bars->s_full[sb]);
      }
      __syncwarp();
      if (h > 0) issue_pv(h - 1, st_prev);
      st_prev = st;
      if (++st == C::kStages) { st = 0; ph ^= 1u; }
    }
    issue_pv(hpg - 1, st_prev);
  } else {
    // ============================== softmax warpgroup (thread == query row) ==============================
    const int row = tid;
    const int n = q0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;

    auto drain_o = [&](int g) {  // O of this CTA's g-th head: TMEM -> bf16 -> global
      const int ob = g & 1;
      mbar_wait(&bars->pv_done[ob], (g >> 1) & 1);
      tc_fence_after();
      __nv_bfloat16* orow = out + (static_cast<long long>(b) * N + n) * (H * D) + (h_begin + g) * D;
#pragma unroll
      for (int c = 0; c < C::kDP / 16; ++c) {
        float o[16];
        tmem_ld16(tmem + lane_base + C::kColO + ob * C::kDP + c * 16, o);
        tmem_wait_ld();
        if (n < N) {
          uint4 lo, hi;
          lo.x = pack_bf16(o[0], o[1]); lo.y = pack_bf16(o[2], o[3]); lo.z = pack_bf16(o[4], o[5]); lo.w = pack_bf16(o[6], o[7]);
          hi.x = pack_bf16(o[8], o[9]); hi.y = pack_bf16(o[10], o[11]); hi.z = pack_bf16(o[12], o[13]); hi.w = pack_bf16(o[14], o[15]);
          if (c * 16 + 8 <= D) *reinterpret_cast<uint4*>(orow + c * 16) = lo;
          if (c * 16 + 16 <= D) *reinterpret_cast<uint4*>(orow + c * 16 + 8) = hi;
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->o_free[ob]);
    };

    for (int hl = 0; hl < hpg; ++hl) {
      const int h = h_begin + hl;
      const int sb = hl & 1;
      mbar_wait(&bars->s_full[sb], (hl >> 1) & 1);
      tc_fence_after();
      float sv[96];
      float sel[kFew ? kXFewTokens : 1];
      const uint32_t s_taddr = tmem + lane_base + sb * kXSlot;
      tmem_ld32(s_taddr, sv);
      tmem_ld32(s_taddr + 32, sv + 32);
      tmem_ld16(s_taddr + 64, sv + 64);
      if (kFew && want_heat) {
#pragma unroll
        for (int t = 0; t < kXFewTokens; ++t)
          if (t < tl.n) sel[t] = tmem_ld1(s_taddr + tl.idx[t]);
      }
      tmem_wait_ld();
      if (M >= 64) {  // the usual case (77 prompt tokens): only the last 16 columns can be padding
#pragma unroll
        for (int i = 64; i < kXMPad; ++i)
          if (i >= M) sv[i] = -INFINITY;
      } else {
#pragma unroll
        for (int i = 0; i < kXMPad; ++i)
          if (i >= M) sv[i] = -INFINITY;
      }
      // packed f32x2 forms and the 3-input max: same roundings and the same summation order as the scalar code (even /
      // odd partial sums), half the issue slots — this stretch, not HBM, bounds the kernel at N = 4096
      static_assert((kXMPad - 6) % 4 == 2, "max reduction below assumes 80 columns");
      float mx0 = fmax3(sv[0], sv[1], sv[2]), mx1 = fmax3(sv[3], sv[4], sv[5]);
#pragma unroll
      for (int i = 6; i + 4 <= kXMPad; i += 4) {
        mx0 = fmax3(mx0, sv[i], sv[i + 1]); mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
      }
      mx0 = fmax3(mx0, sv[kXMPad - 2], sv[kXMPad - 1]);
      const float m = fmaxf(mx0, mx1) * scale_log2;
      const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2), negm2 = pack_f32x2(-m, -m);
      uint64_t sum2 = pack_f32x2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kXMPad; i += 2) {
        float a, b2;
        unpack_f32x2(ffma2(pack_f32x2(sv[i], sv[i + 1]), scale2, negm2), a, b2);
        sv[i] = ex2(a); sv[i + 1] = ex2(b2);
        sum2 = fadd2(sum2, pack_f32x2(sv[i], sv[i + 1]));
      }
      float sum0, sum1;
      unpack_f32x2(sum2, sum0, sum1);
      const float inv_l = 1.0f / (sum0 + sum1);
      const uint64_t inv2 = pack_f32x2(inv_l, inv_l);
#pragma unroll
      for (int i = 0; i < kXMPad; i += 2)  // normalised probabilities: PV needs no later division
        unpack_f32x2(fmul2(pack_f32x2(sv[i], sv[i + 1]), inv2), sv[i], sv[i + 1]);
      if (want_heat) {
        if (kFew) {
#pragma unroll
          for (int t = 0; t < kXFewTokens; ++t) {
            if (t < tl.n) {
              const float pt = ex2(fmaf(sel[t], scale_log2, -m)) * inv_l;  // same ops as sv[idx[t]] above
              if (tl.per_head) {  // DAAM-style: one plane per (batch, head, token), no head mean
                if (n < N) {
                  float* ptr = maps + ((static_cast<long long>(b - b_first) * H + h) * tl.n + t) * N + n;
                  if (accumulate) atomicAdd(ptr, pt);  // result-less RED: one thread per element and launch, no load in the softmax warps
                  else *ptr = pt;
                }
              } else {
                acc[t] += pt;
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < kXMPad; ++i) acc[i] += sv[i];
        }
      }
#pragma unroll
      for (int i = kXMPad; i < 96; ++i) sv[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t u[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = pack_bf16(sv[c * 32 + 2 * i], sv[c * 32 + 2 * i + 1]);
        tmem_st16(s_taddr + c * 16, u);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars->p_full[sb]);
      if (hl > 0) drain_o(hl - 1);
    }
    drain_o(hpg - 1);

    // ---- heat epilogue: mean over heads, selected token columns, coalesced per token plane ----
    if (want_heat && !(kFew && tl.per_head) && gridDim.z == 1) {
      const float inv_h = 1.0f / static_cast<float>(H);
      float* dst = maps + static_cast<long long>(b - b_first) * tl.n * N + n;
      if (kFew) {
        if (n < N) {
#pragma unroll
          for (int t = 0; t < kXFewTokens; ++t) {
            if (t < tl.n) {
              const float val = acc[t] * inv_h;
              float* ptr = dst + static_cast<long long>(t) * N;
              if (accumulate) atomicAdd(ptr, val);  // (RED: same value as load + add + store, no load)
          else *ptr = val;
            }
          }
        }
      } else {
        // all MMAs and TMA loads of this CTA have completed (pv_done of the last head was observed)
        float* srow = reinterpret_cast<float*>(smem) + row * kHeatLd;
#pragma unroll
        for (int i = 0; i < kXMPad; ++i) srow[i] = acc[i];
        if (n < N) {
          for (int t = 0; t < tl.n; ++t) {
            const float val = srow[tl.idx[t]] * inv_h;
            float* ptr = dst + static_cast<long long>(t) * N;
            if (accumulate) atomicAdd(ptr, val);  // (RED: same value as load + add + store, no load)
          else *ptr = val;
          }
        }
      }
    }
    tc_fence_before();
  }
  if (kFew && gridDim.z > 1 && want_heat && !tl.per_head) {
    // head groups -> one heat row: partial sums of ranks 1.. go to the leader's (now idle) stage buffers
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(blockIdx.z), hs = static_cast<int>(gridDim.z);
    float* xacc = reinterpret_cast<float*>(smem);  // [hs - 1][kXFewTokens][128]
    cluster.sync();  // every CTA of the cluster has finished its TMA loads and MMAs: the leader's stages are free
    if (warp < 4 && rank > 0) {
      float* remote = cluster.map_shared_rank(xacc, 0) + (rank - 1) * (kXFewTokens * kXBlockM) + tid;
#pragma unroll
      for (int t = 0; t < kXFewTokens; ++t)
        if (t < tl.n) remote[t * kXBlockM] = acc[t];
    }
    cluster.sync();
    if (warp < 4 && rank == 0 && q0 + tid < N) {
      const float inv_h = 1.0f / static_cast<float>(H);
      float* dst = maps + static_cast<long long>(b - b_first) * tl.n * N + q0 + tid;
#pragma unroll
      for (int t = 0; t < kXFewTokens; ++t) {
        if (t < tl.n) {
          float sum = acc[t];
          for (int r = 1; r < hs; ++r) sum += xacc[((r - 1) * kXFewTokens + t) * kXBlockM + tid];  // fixed order
          const float val = sum * inv_h;
          float* ptr = dst + static_cast<long long>(t) * N;
          if (accumulate) atomicAdd(ptr, val);  // (RED: same value as load + add + store, no load)
          else *ptr = val;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

}  // namespace sm100

template <int D, bool kFew>
static int launch_cross_t(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, float scale,
                        const TokenList& tl, int b_first, float* maps, int accumulate, cudaStream_t stream) {
  using C = sm100::XCfg<D>;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_head_map(&mq, q, B, H, N, D, sm100::kXBlockM)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mk, k, B, H, M, D, sm100::kXMPad)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mv, v, B, H, M, D, sm100::kXMPad)) != AGENDA_OK) return rc;
  constexpr size_t smem = sm100::x_smem_bytes<D>();
  auto kern = sm100::attn_cross_sm100_kernel<D, kFew>;
  AGENDA_DYN_SMEM(kern, smem);
  dim3 grid((N + sm100::kXBlockM - 1) / sm100::kXBlockM, B);
  // few (batch, query tile) pairs: split the heads over a cluster along z while the grid still fits one wave
  int hs = 1;
  if (kFew) {
    static const int sms = [] { int dev = 0, n = 148; cudaGetDevice(&dev);
                                cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
    // (measured on B200, tools/bench_cross.py: N = 256 26.7 -> 13.1 us and N = 64 19.3 -> 9.7 us at 4; 8 is slower)
    while (hs < 4 && H % (hs * 2) == 0 && static_cast<long long>(grid.x) * grid.y * hs * 2 <= sms) hs *= 2;
    if (const char* e = knob("AGENDA_XSPLIT")) { const int v = atoi(e); if (v >= 1 && v <= 8 && H % v == 0) hs = v; }
  }
  grid.z = hs;
  const float scale_log2 = scale * 1.4426950408889634f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(sm100::kXThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = hs;
  cfg.attrs = attr; cfg.numAttrs = hs > 1 ? 1 : 0;
  AGENDA_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, static_cast<__nv_bfloat16*>(out), maps, tl, H, N, M, b_first,
                                 accumulate, scale_log2));
  AGENDA_LAUNCH_CHECK("attn_cross_sm100_kernel");
  return AGENDA_OK;
}

template <int D>
static int launch_cross(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, float scale,
                        const TokenList& tl, int b_first, float* maps, int accumulate, cudaStream_t stream) {
  const bool few = (maps == nullptr) || tl.n <= sm100::kXFewTokens;
  return few ? launch_cross_t<D, true>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, stream)
             : launch_cross_t<D, false>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, stream);
}

int attn_cross_sm100_res(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, int d,
                         float scale, const TokenList& tl, int b_first, float* maps, int accumulate, int QT,
                         void* stream);

// bf16 tensor-core path; returns AGENDA_ERR_UNSUPPORTED for shapes it does not cover.
int attn_cross_sm100(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, int d,
                     float scale, const TokenList& tl, int b_first, float* maps, int accumulate, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M > sm100::kXMPad) return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross (tensor-core path): M=%d > %d", M, sm100::kXMPad);
  if (maps && tl.per_head && tl.n > sm100::kXFewTokens)
    return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross (tensor-core path): per-head maps for %d > %d tokens", tl.n,
                sm100::kXFewTokens);
  const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out);
  if (al & 15) return fail(AGENDA_ERR_MISALIGNED, "attn_cross_fwd_heat: q/k/v/out must be 16-byte aligned");
  // Resident-K/V form (attn_cross_sm100_res.cu): d of one swizzle chunk, few heat tokens, all heads' K/V fit beside the
  // Q ring.  QT query tiles per CTA: the smallest count that puts the whole launch into one wave of CTAs.
  {
    bool use_res = (d == 40 || d == 64) && H <= 8 && (maps == nullptr || tl.n <= sm100::kXFewTokens);
    if (const char* e = knob("AGENDA_XRES")) use_res = use_res && atoi(e) != 0;
    if (use_res) {
      const int n_tiles = (N + sm100::kXBlockM - 1) / sm100::kXBlockM, sms = num_sms();
      int QT = 1;
      while (QT < 32 && ((n_tiles + QT - 1) / QT) * B > sms) ++QT;
      if (const char* e = knob("AGENDA_XRES_QT")) { const int t = atoi(e); if (t >= 1 && t <= 64) QT = t; }
      // one CTA per SM: pays off once every CTA amortises its K/V block over several query tiles (B = 16, N = 4096:
      // 49 vs 55 us); small launches (QT = 1) keep the two-CTAs-per-SM kernel below
      if (QT >= 2)
        return attn_cross_sm100_res(q, k, v, out, B, H, N, M, d, scale, tl, b_first, maps, accumulate, QT, stream);
    }
  }
  switch (d) {
    case 40: return launch_cross<40>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, st);
    case 64: return launch_cross<64>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, st);
    case 80: return launch_cross<80>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, st);
    case 160: return launch_cross<160>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, st);
    default: return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross (tensor-core path): head dim %d not in {40,64,80,160}", d);
  }
}

}  // namespace agenda
