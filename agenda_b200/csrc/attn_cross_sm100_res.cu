// K2 (resident-K/V form): cross-attention + heat epilogue for head dims of one swizzle chunk (d <= 64) and few heat
// tokens, the shape every SD layer at or near latent resolution has (d = 40 at 64^2, d = 64 at 96^2).
//
// The per-(query tile, head) form in attn_cross_sm100.cu is bound by the TMA unit: every head step pulls a 128-row Q
// tile AND the 80-row K and V tiles of that head (~5 engine cycles per 80..128-byte row, tools/ubench/tma.cu), although
// K and V of a batch element are the same for all of its query tiles.  Here one CTA owns QT consecutive query tiles of a
// batch element: K and V of ALL heads are loaded once and stay resident in shared memory (H x 2 x 10 KB), only Q tiles
// stream through a 3-stage ring, one per (query tile, head) step.  TMA rows per 128 queries and head: 128 + 160 / QT
// instead of 288.
//
// Warps: 0-3 and 4-7 two softmax warpgroups (thread == query row) that take the even and the odd steps, 8 TMA
// producer, 9 TMEM allocator + MMA issuer (both converged, elect.sync around the issue).  TMEM: S/P[2] (80 columns
// each) | O[2] (round16(d) each), indexed by step parity == warpgroup.  Step s = qt * H + h.  MMA order: QK(0) QK(1)
// PV(0) QK(2) PV(1) ...; a Q stage is free as soon as its QK has run.  With K/V resident the chain of ONE softmax
// warpgroup (one warp per SM sub-partition) was the limit (60 us for B = 16, N = 4096); two warpgroups overlap their
// load / exponential / store phases.  The head sums of a query tile are split between the warpgroups and combined
// through shared memory behind a 256-thread named barrier once per query tile.
#include <cstdlib>

#include "sm100_common.cuh"

namespace agenda {
namespace sm100 {

constexpr int kRThreads = 320;
constexpr int kRMPad = 80;     // 77 prompt tokens padded to a multiple of 16
constexpr int kRSlot = 80;     // TMEM columns per S/P buffer
constexpr int kRFew = 8;       // heat tokens handled (register accumulators)
constexpr int kRStages = 3;
constexpr int kRQBytes = 128 * 128;       // one Q tile: 128 rows x 128-byte swizzled rows
constexpr int kRKVBytes = kRMPad * 128;   // one K (or V) tile of one head

#ifdef AGENDA_XRES_TRACE  // tools/ubench/trace_cross.cu: clock64() stamps of CTA (0,0), first 40 steps
__device__ long long g_xres_trace[4 * 40 * 4];
#define XR_TRACE(actor, s, ev)                                                                          \
  do {                                                                                                  \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (s) < 40 && (actor) < 4)                      \
      g_xres_trace[((actor) * 40 + (s)) * 4 + (ev)] = clock64();                                        \
  } while (0)
#else
#define XR_TRACE(actor, s, ev) do { } while (0)
#endif

struct RBarriers {
  float xacc[2][2][kRFew][128];  // [query-tile parity][warpgroup][token][row]: partial head sums
  uint64_t kv_full;
  uint64_t q_full[kRStages], q_empty[kRStages];
  uint64_t s_full[2], s_free[2], p_full[2], pv_done[2], o_free[2];
  uint32_t tmem_base;
};

constexpr int kRPSlot = 48;    // TMEM columns per P buffer (80 bf16 = 40 packed columns, stored as 3 x 16)

template <int D>
struct RCfg {
  static constexpr int kDP = (D + 15) / 16 * 16;
  static constexpr int kColP = 2 * kRSlot;               // P has its own columns: QK(s+2) may overwrite S[sb] as soon as
  static constexpr int kColO = kColP + 2 * kRPSlot;      // the softmax warpgroup has pulled S(s) into registers
  static constexpr int kTmemCols = 512;
  static_assert(kColO + 2 * kDP <= 512, "TMEM overflow");
  static_assert(D <= 64, "one 64-element swizzle chunk per row");
};

inline size_t r_smem_bytes(int H) { return 1024 + kRStages * kRQBytes + static_cast<size_t>(H) * 2 * kRKVBytes + sizeof(RBarriers) + 64; }

template <int D>
__global__ void __launch_bounds__(kRThreads, 1)
attn_cross_sm100_res_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                            const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ out,
                            float* __restrict__ maps, const TokenList tl, int H, int N, int M, int QT, int b_first,
                            int accumulate, float scale_log2, int prefetch) {
  using C = RCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                                  // kRStages Q tiles
  unsigned char* sK = sQ + kRStages * kRQBytes;              // H resident K tiles
  unsigned char* sV = sK + H * kRKVBytes;                    // H resident V tiles
  RBarriers* bars = reinterpret_cast<RBarriers*>(sV + H * kRKVBytes);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y;
  const int tile0 = blockIdx.x * QT;                          // first query tile of this CTA
  const int n_qt = min(QT, (N + 127) / 128 - tile0);          // query tiles that exist
  const int n_steps = n_qt * H;
  const bool want_heat = (maps != nullptr) && (b >= b_first);

  if (tid == 8 * 32) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < kRStages; ++s) { mbar_init(&bars->q_full[s], 1); mbar_init(&bars->q_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s_full[s], 1); mbar_init(&bars->s_free[s], 128); mbar_init(&bars->p_full[s], 128);
      mbar_init(&bars->pv_done[s], 1); mbar_init(&bars->o_free[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(&bars->tmem_base, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 8) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      // first Q tile ahead of the K/V block, so QK(0) can start as early as possible
      mbar_expect_tx(&bars->q_full[0], kRQBytes);
      tma_load_4d(&map_q, &bars->q_full[0], sQ, 0, 0, tile0 * 128, b);
      mbar_expect_tx(&bars->kv_full, H * 2 * kRKVBytes);
      for (int h = 0; h < H; ++h) {
        tma_load_4d(&map_k, &bars->kv_full, sK + h * kRKVBytes, 0, h, 0, b);
        tma_load_4d(&map_v, &bars->kv_full, sV + h * kRKVBytes, 0, h, 0, b);
      }
    }
    __syncwarp();
    // A streamed Q tile (128 rows of 80..128 bytes, 640-byte stride) takes ~3 us from HBM with three tiles in flight —
    // the whole kernel ran at one step per microsecond.  The tiles are therefore pulled into L2 kPrefetch steps ahead
    // (no shared memory needed); the actual load then only pays the L2 latency.
    const int kPrefetch = prefetch;
    if (elect_one())
      for (int s = 1; s < min(n_steps, kPrefetch); ++s)
        tma_prefetch_l2_4d(&map_q, 0, s % H, (tile0 + s / H) * 128, b);
    __syncwarp();
    int st = 1 % kRStages;
    uint32_t ph = (kRStages == 1) ? 1u : 0u;
    for (int s = 1; s < n_steps; ++s) {
      const int qt = s / H, h = s - qt * H;
      mbar_wait(&bars->q_empty[st], ph ^ 1);
      XR_TRACE(3, s, 0);
      if (elect_one()) {
        const int sp = s + kPrefetch - 1;
        if (kPrefetch > 0 && sp < n_steps) tma_prefetch_l2_4d(&map_q, 0, sp % H, (tile0 + sp / H) * 128, b);
        mbar_expect_tx(&bars->q_full[st], kRQBytes);
        tma_load_4d(&map_q, &bars->q_full[st], sQ + st * kRQBytes, 0, h, (tile0 + qt) * 128, b);
      }
      __syncwarp();
      if (++st == kRStages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 9) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc_qk = make_idesc(128, kRMPad, 0);
    constexpr uint32_t idesc_pv = make_idesc(128, C::kDP, 1);
    const uint64_t q_desc0 = make_sdesc(smem_u32(sQ), 16, 1024);
    const uint64_t k_desc0 = make_sdesc(smem_u32(sK), 16, 1024);
    const uint64_t v_desc0 = make_sdesc(smem_u32(sV), kRMPad * 128, 1024);
    int st = 0;
    uint32_t ph = 0;
    auto issue_qk = [&](int s) {  // S[s & 1] = Q(step s) K_h^T; releases the Q stage when it has run
      const int sb = s & 1, h = s % H;
      mbar_wait(&bars->q_full[st], ph);
      tc_fence_after();
      XR_TRACE(2, s, 0);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < C::kDP / 16; ++kk) {
          const uint64_t adesc = q_desc0 + static_cast<uint64_t>((st * kRQBytes + kk * 32) >> 4);
          const uint64_t bdesc = k_desc0 + static_cast<uint64_t>((h * kRKVBytes + kk * 32) >> 4);
          umma_ss(tmem + sb * kRSlot, adesc, bdesc, idesc_qk, kk != 0);
        }
        umma_commit(&bars->s_full[sb]);
        umma_commit(&bars->q_empty[st]);
      }
      __syncwarp();
      XR_TRACE(2, s, 1);
      if (++st == kRStages) { st = 0; ph ^= 1u; }
    };
    auto issue_pv = [&](int s) {
      const int sb = s & 1, h = s % H;
      const uint32_t php = (s >> 1) & 1;
      mbar_wait(&bars->p_full[sb], php);
      mbar_wait(&bars->o_free[sb], php ^ 1);  // O[sb] drained (step s-2) — done by the warpgroup before its softmax(s)
      tc_fence_after();
      XR_TRACE(2, s, 2);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kRMPad / 16; ++kk) {
          const uint64_t bdesc = v_desc0 + static_cast<uint64_t>((h * kRKVBytes + kk * 2048) >> 4);
          umma_ts(tmem + C::kColO + sb * C::kDP, tmem + C::kColP + sb * kRPSlot + kk * 8, bdesc, idesc_pv, kk != 0);
        }
        umma_commit(&bars->pv_done[sb]);
      }
      __syncwarp();
      XR_TRACE(2, s, 3);
    };
    mbar_wait(&bars->kv_full, 0);
    // QK(0) QK(1) | QK(s+2) as soon as S(s) sits in registers, PV(s) when P(s) is in TMEM: neither warpgroup's next
    // score tile waits for its own exponentials
    issue_qk(0);
    if (n_steps > 1) issue_qk(1);
    for (int s = 0; s < n_steps; ++s) {
      if (s + 2 < n_steps) {
        mbar_wait(&bars->s_free[s & 1], (s >> 1) & 1);
        tc_fence_after();
        issue_qk(s + 2);
      }
      issue_pv(s);
    }
  } else {
    // ============================== softmax warpgroups (thread == query row) ==============================
    const int wg = warp >> 2;  // takes the steps s with s % 2 == wg (its own S/P and O slots)
    const int row = tid & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    float acc[kRFew];
#pragma unroll
    for (int i = 0; i < kRFew; ++i) acc[i] = 0.f;

    auto drain_o = [&](int g) {  // O of step g: TMEM -> bf16 -> global
      const int ob = g & 1, gq = g / H, gh = g - gq * H;
      const int n = (tile0 + gq) * 128 + row;
      mbar_wait(&bars->pv_done[ob], (g >> 1) & 1);
      tc_fence_after();
      __nv_bfloat16* orow = out + (static_cast<long long>(b) * N + n) * (H * D) + gh * D;
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

    // warpgroup 0: mean over heads of query tile `qt` from the two parked halves -> global (after warpgroup 1's signal)
    // Accumulate mode adds with a result-less atomic (RED): every heat element is touched by exactly one thread per
    // launch, so the value is the same as load + add + store, but no global LOAD sits in the softmax warps — a load's
    // scoreboard is shared with the tcgen05.ld of the next score tile, and waiting for that tile then also waited out
    // a full global-memory round trip (event trace: ~2800 cycles at every query-tile boundary).
    auto combine_heat = [&](int qt, int n_row) {
      asm volatile("bar.sync %0, 256;" ::"r"(1 + (qt & 1)) : "memory");
      if (n_row < N) {
        const float inv_h = 1.0f / static_cast<float>(H);
        float* dst = maps + static_cast<long long>(b - b_first) * tl.n * N + n_row;
#pragma unroll
        for (int t = 0; t < kRFew; ++t) {
          if (t < tl.n) {
            const float val = (bars->xacc[qt & 1][0][t][row] + bars->xacc[qt & 1][1][t][row]) * inv_h;
            if (accumulate) atomicAdd(dst + static_cast<long long>(t) * N, val);
            else dst[static_cast<long long>(t) * N] = val;
          }
        }
      }
    };
    int prev_s = -1;  // this warpgroup's previous step, whose O is drained after the next P has been handed over
    for (int qt = 0; qt < n_qt; ++qt) {
      const int n = (tile0 + qt) * 128 + row;
      // The head sums of a query tile are combined ONE TILE LATE (by warpgroup 0, at the end of the next tile): by
      // then warpgroup 1's half has long arrived, so neither warpgroup ever waits for the other at a tile boundary.
      const int n_prev = n - 128;  // row of this thread in the previous query tile
      for (int h = 0; h < H; ++h) {
        const int s = qt * H + h;
        if ((s & 1) != wg) continue;
        const int sb = wg;
        XR_TRACE((warp & 3) == 0 ? wg : 9, s, 0);
        mbar_wait(&bars->s_full[sb], (s >> 1) & 1);
        tc_fence_after();
        XR_TRACE((warp & 3) == 0 ? wg : 9, s, 1);
        float sv[96];
        float sel[kRFew];
        const uint32_t s_taddr = tmem + lane_base + sb * kRSlot;
        tmem_ld32(s_taddr, sv);
        tmem_ld32(s_taddr + 32, sv + 32);
        tmem_ld16(s_taddr + 64, sv + 64);
        if (want_heat) {
#pragma unroll
          for (int t = 0; t < kRFew; ++t)
            if (t < tl.n) sel[t] = tmem_ld1(s_taddr + tl.idx[t]);
        }
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&bars->s_free[sb]);  // S(s) is in registers: the tensor core may compute S(s+2) into this slot
        if (M >= 64) {  // the usual case (77 prompt tokens): only the last 16 columns can be padding
#pragma unroll
          for (int i = 64; i < kRMPad; ++i)
            if (i >= M) sv[i] = -INFINITY;
        } else {
#pragma unroll
          for (int i = 0; i < kRMPad; ++i)
            if (i >= M) sv[i] = -INFINITY;
        }
        // packed f32x2 forms and the 3-input max: same roundings and the same summation order as the scalar code (even /
        // odd partial sums), half the issue slots — this stretch, not HBM, bounds the kernel at N = 4096
        static_assert((kRMPad - 6) % 4 == 2, "max reduction below assumes 80 columns");
        float mx0 = fmax3(sv[0], sv[1], sv[2]), mx1 = fmax3(sv[3], sv[4], sv[5]);
#pragma unroll
        for (int i = 6; i + 4 <= kRMPad; i += 4) {
          mx0 = fmax3(mx0, sv[i], sv[i + 1]); mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
        }
        mx0 = fmax3(mx0, sv[kRMPad - 2], sv[kRMPad - 1]);
        const float m = fmaxf(mx0, mx1) * scale_log2;
        const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2), negm2 = pack_f32x2(-m, -m);
        uint64_t sum2 = pack_f32x2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kRMPad; i += 2) {
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
        for (int i = 0; i < kRMPad; i += 2)  // normalised probabilities: PV needs no later division
          unpack_f32x2(fmul2(pack_f32x2(sv[i], sv[i + 1]), inv2), sv[i], sv[i + 1]);
        if (want_heat) {
#pragma unroll
          for (int t = 0; t < kRFew; ++t) {
            if (t < tl.n) {
              const float pt = ex2(fmaf(sel[t], scale_log2, -m)) * inv_l;  // same ops as sv[idx[t]] above
              if (tl.per_head) {  // DAAM-style: one plane per (batch, head, token), no head mean
                if (n < N) {
                  float* ptr = maps + ((static_cast<long long>(b - b_first) * H + h) * tl.n + t) * N + n;
                  if (accumulate) atomicAdd(ptr, pt);  // result-less RED: one thread per element and launch
                  else *ptr = pt;
                }
              } else {
                acc[t] += pt;
              }
            }
          }
        }
#pragma unroll
        for (int i = kRMPad; i < 96; ++i) sv[i] = 0.f;
        // O of this warpgroup's previous step (s-2) is drained here: after the exponentials of step s (PV(s-2) only
        // started when P(s-2) was handed over, so its result is ready by now without waiting) and before P(s) is
        // stored (pv_done(s-2) also says PV(s-2) is done reading these P columns), so PV(s) finds its accumulator
        // free the moment P(s) is handed over.
        if (prev_s >= 0) drain_o(prev_s);
        XR_TRACE((warp & 3) == 0 ? wg : 9, s, 3);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t u[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = pack_bf16(sv[c * 32 + 2 * i], sv[c * 32 + 2 * i + 1]);
          tmem_st16(tmem + lane_base + C::kColP + sb * kRPSlot + c * 16, u);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->p_full[sb]);
        XR_TRACE((warp & 3) == 0 ? wg : 9, s, 2);
        prev_s = s;
      }
      // ---- heat: both warpgroups park their head sums of this tile; warpgroup 1 signals and moves on; warpgroup 0
      //      combines the PREVIOUS tile (fixed summation order, one coalesced store per token plane).  Barrier ids and
      //      xacc buffers alternate per query tile: the warpgroups cannot drift a whole tile apart (the MMA warp issues
      //      their steps in order and blocks on the slower one's s_free within two steps). ----
      if (want_heat && !tl.per_head) {
#pragma unroll
        for (int t = 0; t < kRFew; ++t) {
          if (t < tl.n) bars->xacc[qt & 1][wg][t][row] = acc[t];
          acc[t] = 0.f;
        }
        if (wg == 1) {
          __threadfence_block();
          asm volatile("bar.arrive %0, 256;" ::"r"(1 + (qt & 1)) : "memory");
        } else if (qt > 0) {
          combine_heat(qt - 1, n_prev);
        }
      }
    }
    if (want_heat && !tl.per_head && wg == 0) {  // the last tile: this one does wait for warpgroup 1
      const int qt = n_qt - 1;
      const int n_last = (tile0 + qt) * 128 + row;
      combine_heat(qt, n_last);
    }
    if (prev_s >= 0) drain_o(prev_s);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

}  // namespace sm100

template <int D>
static int launch_cross_res(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, float scale,
                            const TokenList& tl, int b_first, float* maps, int accumulate, int QT, cudaStream_t stream) {
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_head_map(&mq, q, B, H, N, D, 128)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mk, k, B, H, M, D, sm100::kRMPad)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mv, v, B, H, M, D, sm100::kRMPad)) != AGENDA_OK) return rc;
  const size_t smem = sm100::r_smem_bytes(H);
  auto kern = sm100::attn_cross_sm100_res_kernel<D>;
  AGENDA_DYN_SMEM(kern, smem);
  const int n_tiles = (N + 127) / 128;
  dim3 grid((n_tiles + QT - 1) / QT, B);
  kern<<<grid, sm100::kRThreads, smem, stream>>>(mq, mk, mv, static_cast<__nv_bfloat16*>(out), maps, tl, H, N, M, QT,
                                                 b_first, accumulate, scale * 1.4426950408889634f,
                                                 knob("AGENDA_XRES_PF") ? atoi(knob("AGENDA_XRES_PF")) : 0);
  AGENDA_LAUNCH_CHECK("attn_cross_sm100_res_kernel");
  return AGENDA_OK;
}

// Returns AGENDA_ERR_UNSUPPORTED (without setting an error the caller reports) when the shape is not covered, so the
// dispatcher falls through to the per-(query tile, head) kernel.
int attn_cross_sm100_res(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int M, int d,
                         float scale, const TokenList& tl, int b_first, float* maps, int accumulate, int QT,
                         void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d == 40) return launch_cross_res<40>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, QT, st);
  if (d == 64) return launch_cross_res<64>(q, k, v, out, B, H, N, M, scale, tl, b_first, maps, accumulate, QT, st);
  return AGENDA_ERR_UNSUPPORTED;
}

#ifdef AGENDA_XRES_TRACE
extern "C" int agenda_xres_trace_read(long long* host, int n) {
  if (n < 4 * 40 * 4) return -1;
  return cudaMemcpyFromSymbol(host, sm100::g_xres_trace, sizeof(long long) * 4 * 40 * 4) == cudaSuccess ? 640 : -1;
}
#endif

}  // namespace agenda
