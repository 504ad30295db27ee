// K2x3: cross-attention + heat epilogue with fp32-accurate logits on the bf16 tensor cores (sm_100a).
//
//   out  = softmax(scale * Q K^T) V                                    (data_generation/hook.py:108,114)
//   maps[b', t, n] (=|+=) mean_h softmax(...)[b, h, n, token_idx[t]]    (_unravel_attn, hook.py:28-56)
//
// Why: the reference runs in fp32 and the heat maps are held to 1e-4 max-abs.  A bf16 Q / K carries a
// relative error of 2^-9 per element, i.e. logit errors of ~1e-3..1e-2 and probability errors up to ~2.5e-3; tf32
// (truncated to 10 mantissa bits by the tensor core) is borderline.  This kernel takes Q in fp32 and K as a
// pre-split pair K = K_hi + K_lo (both bf16, K_lo = bf16(K - K_hi)), splits every Q tile the same way on chip and
// accumulates S = Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T in the fp32 TMEM accumulator: products of bf16 values are
// exact in fp32, the dropped Q_lo K_lo^T term is 2^-18 relative, so the logits match an fp32 baddbmm to ~1e-5 and the
// probabilities to ~1e-6.  Cross-attention is 3 % of the attention FLOPs, so the 3x MMA count is free;
// the kernel stays HBM-bound (Q fp32 in, O out).
//
// Prompt side: K and V depend on the prompt only, so they are packed ONCE per prompt (agenda_pack_context_kv) into the
// exact shared-memory image the tensor core reads — per (batch, head): K_hi chunks | K_lo chunks | V chunks, each an
// 80-row x 128-byte tile in the 128B-swizzled UMMA layout, zero padded — and a head's block arrives with ONE
// cp.async.bulk each for K and V (no tensor map, no per-row TMA work: ~290 rows per head in the per-row form).
//
// Loop order is HEAD-OUTER: a CTA owns QT consecutive 128-query tiles of one batch element; for each head it keeps
// K_hi / K_lo / V of that head resident (double-buffered where shared memory allows, so the next head's block arrives
// under the current head's steps) and streams the head's fp32 Q column chunks of all its query tiles.
// Step s = head * QT + tile.
//
//   softmax warpgroups (2; thread == query row)  even / odd steps, own S / P / O TMEM slots
//   producer warp      fp32 Q chunks [128 x W] by TMA through a ring (W = 40 or 64 columns) + the K / V bulk copies,
//                      as one polling state machine (neither stream ever blocks the other)
//   MMA warp           TMEM allocator + tcgen05.mma issuer
//   2 converter warps  fp32 Q chunk -> bf16 hi / lo operand tiles in the 128B-swizzled K-major UMMA layout
//
// Heat, few tokens (T <= 8, what every caller of the reference reads): every softmax thread adds its selected-token
// probabilities into a shared-memory row of its own ([warpgroup][tile][token][row]); after the last head warpgroup 0
// adds the two halves in a fixed order and writes the mean over heads (deterministic, no atomics between CTAs).  Small
// launches split the heads over a cluster along z and finish the sum through the leader's shared memory in rank order.
// Heat, all tokens (kAll; the reference's own behaviour: all 77 maps): one softmax warpgroup, 80 register accumulators
// per thread, parked in shared memory at the end and written per token plane, coalesced.
#include <cooperative_groups.h>
#include <cstdlib>
#include <type_traits>

#include "sm100_common.cuh"

namespace agenda {
namespace sm100 {

namespace cg = cooperative_groups;

constexpr int kTMPad = 80;         // 77 prompt tokens padded to a multiple of 16
constexpr int kTSlot = 80;         // TMEM columns per S buffer
constexpr int kTPSlot = 48;        // TMEM columns per P buffer (80 bf16 = 40 packed columns, stored as 3 x 16)
constexpr int kTFew = 8;           // heat tokens of the few-token form
constexpr int kTMaxQT = 4;         // query tiles per CTA
constexpr int kTTile = kTMPad * 128;  // one 80-row, 128B-swizzled K / V tile
constexpr int kTMaxStages = 6;
constexpr int kTConvThreads = 64;
constexpr int kTHeatLd = 81;       // kAll: padded accumulator row in shared memory (bank-conflict free)

template <int D>
struct TCfg {
  static constexpr int kW = (D % 64 == 0) ? 64 : 40;     // Q / K column chunk (d = 40: 1, 64: 1, 80: 2, 160: 4 chunks)
  static constexpr int kNC = D / kW;
  static constexpr int kWP = (kW + 15) / 16 * 16;        // K extent of a chunk inside the MMAs (zero-padded columns)
  static constexpr int kDP = (D + 15) / 16 * 16;         // N extent of the PV MMA
  static constexpr int kVC = (D + 63) / 64;              // 64-column swizzle chunks of V
  static constexpr int kQ32Bytes = 128 * kW * 4;         // one fp32 Q chunk, dense rows (no swizzle)
  static constexpr int kQBBytes = 2 * 128 * 128;         // bf16 hi tile + lo tile
  static constexpr int kKBytes = 2 * kNC * kTTile;       // K_hi chunks | K_lo chunks of one head
  static constexpr int kVBytes = kVC * kTTile;
  static constexpr int kKVBytes = kKBytes + kVBytes;     // one (batch, head) block of the packed context
  static constexpr int kKBufs = (D <= 80) ? 2 : 1;
  static constexpr int kVBufs = (D <= 64) ? 2 : 1;
  static constexpr int kQBBufs = 2;                    // bf16 operand buffers: conversion of chunk j waits for QK(j - kQBBufs)
  static constexpr bool kAliasP = (2 * kTSlot + 2 * kTPSlot + 2 * kDP > 512);  // d = 160: P overwrites its S slot
  // score slots: QK runs up to kSSlots steps ahead of the softmax that consumes it (P / O keep one slot per warpgroup)
  // (measured, tools/bench_cross_x3.py: a third slot helps d = 64 — 169.8 -> 165.2 us at N = 9216 — and costs ~2 us at d = 40 / 80)
  static constexpr int kSSlots = (D == 64) ? 3 : 2;
  static constexpr int kColP = kAliasP ? 0 : kSSlots * kTSlot;
  static constexpr int kPStride = kAliasP ? kTSlot : kTPSlot;
  static constexpr int kColO = kAliasP ? 2 * kTSlot : kSSlots * kTSlot + 2 * kTPSlot;
  static constexpr int kFixedBytes = kQBBufs * kQBBytes + kKBufs * kKBytes + kVBufs * kVBytes;
  static_assert(D % kW == 0, "head dim must be a whole number of chunks");
  static_assert(kColO + 2 * kDP <= 512, "TMEM overflow");
};

struct TBarriers {
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t q32_full[kTMaxStages], q32_empty[kTMaxStages];
  uint64_t qb_full[3], qb_empty[3];
  uint64_t s_full[3], s_free[3], p_full[2], pv_done[2], o_free[2];
  uint32_t tmem_base;
};

constexpr int kTSmemMax = 232448;  // 227 KB: the per-CTA opt-in limit on sm_100

__device__ __forceinline__ void st_shared_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// contiguous global -> shared bulk copy, completion on an mbarrier (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// kAll = false: T <= 8 heat tokens, two softmax warpgroups (384 threads).  kAll = true: up to 80 heat tokens (all of
// the prompt), one softmax warpgroup with register accumulators (256 threads), one query tile per CTA.
template <int D, typename OutT, bool kAll>
__global__ void __launch_bounds__(kAll ? 256 : 384, 1)
attn_cross_sm100_x3_kernel(const __grid_constant__ CUtensorMap map_q, const unsigned char* __restrict__ kv_blob,
                           OutT* __restrict__ out, float* __restrict__ maps, const TokenList tl, int H, int N, int M,
                           int QT, int n_stages, int b_first, int accumulate, float scale_log2,
                           const float* __restrict__ q_hm) {
  using C = TCfg<D>;
  constexpr int W = C::kW;
  constexpr int kWGs = kAll ? 1 : 2;
  constexpr int kProdWarp = 4 * kWGs, kMmaWarp = kProdWarp + 1, kConvWarp0 = kProdWarp + 2;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQB = smem;                                          // kQBBufs x (hi tile, lo tile)
  unsigned char* sK = sQB + C::kQBBufs * C::kQBBytes;                 // kKBufs x (K_hi chunks, K_lo chunks)
  unsigned char* sV = sK + C::kKBufs * C::kKBytes;                    // kVBufs x V chunks
  unsigned char* sQ32 = sV + C::kVBufs * C::kVBytes;                  // n_stages fp32 Q chunks
  float* xacc = reinterpret_cast<float*>(sQ32 + n_stages * C::kQ32Bytes);  // few: [2][QT][nh][128]
  const int nh = (kAll || maps == nullptr || tl.per_head) ? 0 : tl.n;      // heat tokens summed over heads in xacc
  TBarriers* bars = reinterpret_cast<TBarriers*>(xacc + 2 * QT * nh * 128);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y;
  const int tile0 = blockIdx.x * QT;
  const int n_qt = min(QT, (N + 127) / 128 - tile0);
  const int hpg = H / static_cast<int>(gridDim.z);        // heads of this CTA (gridDim.z > 1: cluster along the heads)
  const int h_begin = static_cast<int>(blockIdx.z) * hpg;
  const int n_steps = hpg * n_qt;                          // step s = hl * n_qt + qt
  const bool want_heat = (maps != nullptr) && (b >= b_first);

  if (tid == kProdWarp * 32) {
    if (q_hm == nullptr) tma_prefetch_desc(&map_q);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->k_full[i], 1); mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1); mbar_init(&bars->v_empty[i], 1);
    }
    for (int i = 0; i < kTMaxStages; ++i) { mbar_init(&bars->q32_full[i], 1); mbar_init(&bars->q32_empty[i], kTConvThreads); }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars->qb_full[i], kTConvThreads); mbar_init(&bars->qb_empty[i], 1);
      mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_free[i], 128);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->p_full[i], 128); mbar_init(&bars->pv_done[i], 1); mbar_init(&bars->o_free[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  float hsum[kTFew];  // few + cluster: warpgroup 0's head sums of this row (query tile 0) for the cluster combine
#pragma unroll
  for (int t = 0; t < kTFew; ++t) hsum[t] = 0.f;

  if (warp == kProdWarp) {
    // ============================== producer: Q chunks (TMA) + K / V blocks (bulk copies) ==============================
    // One polling state machine over three independent streams: a full Q ring never delays the next head's K block
    // and a V buffer still being read by the tensor core never delays Q.
    const unsigned char* blob_b = kv_blob + static_cast<size_t>(b) * H * C::kKVBytes;
    const int n_qchunks = n_steps * C::kNC;
    int qi = 0, st = 0, ki = 0, vi = 0;
    uint32_t ph = 0;
    while (qi < n_qchunks || ki < hpg || vi < hpg) {
      bool progress = false;
      if (ki < hpg && mbar_test(&bars->k_empty[ki % C::kKBufs], ((ki / C::kKBufs) & 1) ^ 1)) {
        if (elect_one()) {
          const int kb = ki % C::kKBufs;
          mbar_expect_tx(&bars->k_full[kb], C::kKBytes);
          bulk_load(sK + kb * C::kKBytes, blob_b + static_cast<size_t>(h_begin + ki) * C::kKVBytes, C::kKBytes, &bars->k_full[kb]);
        }
        __syncwarp();
        ++ki;
        progress = true;
      }
      if (qi < n_qchunks && mbar_test(&bars->q32_empty[st], ph ^ 1)) {
        const int s = qi / C::kNC, c = qi - s * C::kNC;
        const int hl = s / n_qt, qt = s - hl * n_qt;
        if (q_hm != nullptr) {
          // chunk-major Q (agenda_linear_split_f32_heads): the 128-query chunk of (batch, head, column chunk) is one dense
          // block -> ONE bulk copy instead of a 128-row tensor load.  A ragged last tile copies its rows only; the tail of
          // the stage is zero-filled first (what the tensor map's out-of-bounds fill did), by the whole warp.
          const int row0 = (tile0 + qt) * 128;
          const int rows = min(128, N - row0);
          unsigned char* dst = sQ32 + st * C::kQ32Bytes;
          if (rows < 128) {
            const int lane = tid & 31;
            for (int off = rows * W * 4 + lane * 16; off < C::kQ32Bytes; off += 32 * 16)
              *reinterpret_cast<uint4*>(dst + off) = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
          }
          if (elect_one()) {
            const float* src = q_hm + ((((static_cast<size_t>(b) * H + (h_begin + hl)) * C::kNC + c) * N + row0) * W);
            mbar_expect_tx(&bars->q32_full[st], rows * W * 4);
            bulk_load(dst, src, rows * W * 4, &bars->q32_full[st]);
          }
        } else if (elect_one()) {
          mbar_expect_tx(&bars->q32_full[st], C::kQ32Bytes);
          tma_load_4d(&map_q, &bars->q32_full[st], sQ32 + st * C::kQ32Bytes, c * W, h_begin + hl, (tile0 + qt) * 128, b);
        }
        __syncwarp();
        ++qi;
        if (++st == n_stages) { st = 0; ph ^= 1u; }
        progress = true;
      }
      if (vi < hpg && mbar_test(&bars->v_empty[vi % C::kVBufs], ((vi / C::kVBufs) & 1) ^ 1)) {
        if (elect_one()) {
          const int vb = vi % C::kVBufs;
          mbar_expect_tx(&bars->v_full[vb], C::kVBytes);
          bulk_load(sV + vb * C::kVBytes, blob_b + static_cast<size_t>(h_begin + vi) * C::kKVBytes + C::kKBytes, C::kVBytes,
                    &bars->v_full[vb]);
        }
        __syncwarp();
        ++vi;
        progress = true;
      }
      if (!progress) __nanosleep(40);
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer (warp converged; tcgen05.mma / commit under elect.sync) ==============
    constexpr uint32_t idesc_qk = make_idesc(128, kTMPad, 0);
    constexpr uint32_t idesc_pv = make_idesc(128, C::kDP, 1);
    const uint64_t qb_desc0 = make_sdesc(smem_u32(sQB), 16, 1024);
    const uint64_t k_desc0 = make_sdesc(smem_u32(sK), 16, 1024);
    const uint64_t v_desc0 = make_sdesc(smem_u32(sV), kTTile, 1024);
    int gq = 0;                          // running Q chunk counter (operand-buffer ring position)
    int k_waited = -1, v_waited = -1;    // last head whose K / V block has been observed
    auto issue_qk = [&](int j) {  // S[j & 1] = Q(step j) K_h^T as three bf16 MMA groups per column chunk
      const int sb = j % C::kSSlots, hl = j / n_qt, kb = hl % C::kKBufs;
      if (hl > k_waited) {
        mbar_wait(&bars->k_full[kb], (hl / C::kKBufs) & 1);
        k_waited = hl;
      }
      // the previous user of this score slot, S(j - kSSlots), is in its softmax warpgroup's registers
      if (j >= C::kSSlots) mbar_wait(&bars->s_free[sb], ((j - C::kSSlots) / C::kSSlots) & 1);
      for (int c = 0; c < C::kNC; ++c, ++gq) {
        const int cb = gq % C::kQBBufs;
        mbar_wait(&bars->qb_full[cb], (gq / C::kQBBufs) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = cb * C::kQBBytes, a_lo = a_hi + 128 * 128;
          const uint32_t b_hi = kb * C::kKBytes + c * kTTile, b_lo = b_hi + C::kNC * kTTile;
#pragma unroll
          for (int g = 0; g < 3; ++g) {  // hi*hi, lo*hi, hi*lo
            const uint32_t ao = (g == 1) ? a_lo : a_hi, bo = (g == 2) ? b_lo : b_hi;
#pragma unroll
            for (int kk = 0; kk < C::kWP / 16; ++kk)
              umma_ss(tmem + sb * kTSlot, qb_desc0 + static_cast<uint64_t>((ao + kk * 32) >> 4),
                      k_desc0 + static_cast<uint64_t>((bo + kk * 32) >> 4), idesc_qk, !(c == 0 && g == 0 && kk == 0));
          }
          umma_commit(&bars->qb_empty[cb]);
          if (c == C::kNC - 1) {
            umma_commit(&bars->s_full[sb]);
            if (j == hl * n_qt + n_qt - 1) umma_commit(&bars->k_empty[kb]);  // last QK of this head: its K block is free
          }
        }
        __syncwarp();
      }
    };
    auto issue_pv = [&](int s) {
      const int sb = s & 1, hl = s / n_qt, vb = hl % C::kVBufs;
      const uint32_t php = (s >> 1) & 1;
      if (hl > v_waited) {
        mbar_wait(&bars->v_full[vb], (hl / C::kVBufs) & 1);
        v_waited = hl;
      }
      mbar_wait(&bars->p_full[sb], php);
      mbar_wait(&bars->o_free[sb], php ^ 1);  // O[sb] of step s-2 drained
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kTMPad / 16; ++kk)
          umma_ts(tmem + C::kColO + sb * C::kDP, tmem + C::kColP + sb * C::kPStride + kk * 8,
                  v_desc0 + static_cast<uint64_t>((vb * C::kVBytes + kk * 2048) >> 4), idesc_pv, kk != 0);
        umma_commit(&bars->pv_done[sb]);
        if (s == hl * n_qt + n_qt - 1) umma_commit(&bars->v_empty[vb]);  // last PV of this head: its V block is free
      }
      __syncwarp();
    };
    // QK runs up to kSSlots steps ahead of PV (one when P aliases S: QK(s+2) would overwrite P(s) before PV(s) has read it)
    int nq = 0;
    for (int s = 0; s < n_steps; ++s) {
      while (nq < n_steps && nq <= s + (C::kAliasP ? 1 : C::kSSlots)) {
        issue_qk(nq);
        ++nq;
      }
      issue_pv(s);
    }
  } else if (warp >= kConvWarp0) {
    // ============================== converters: fp32 Q chunk -> bf16 hi / lo UMMA operand tiles ======================
    // Work item = 8 consecutive floats of a row (two 16-byte loads -> one 16-byte piece of the hi tile and one of the lo
    // tile).  Consecutive threads take consecutive items: the loads walk shared memory linearly (the fp32 chunk is dense,
    // item i sits at byte 32 i) and a quarter warp's stores fill one 128-byte swizzled row — no 8-way bank conflicts of
    // a thread-per-row mapping (row stride 160 / 256 bytes).
    const int ct = tid - kConvWarp0 * 32;
    constexpr int kPieces = W / 8;                       // 16-byte bf16 pieces per row
    constexpr int kItems = 128 * kPieces / kTConvThreads;  // per thread and chunk: 10 (W = 40) or 16 (W = 64)
    if (C::kWP > W) {  // zero padding up to the MMA K extent (chunk 40 -> 48): written once, never overwritten
      for (int i = ct; i < C::kQBBufs * 2 * 128; i += kTConvThreads) {
        const int r = i & 127;
        unsigned char* tile = sQB + (i >> 7) * (128 * 128);
#pragma unroll
        for (int c16 = kPieces; c16 < C::kWP / 8; ++c16) st_shared_v4(tile + r * 128 + ((c16 ^ (r & 7)) << 4), 0u, 0u, 0u, 0u);
      }
    }
    int st = 0;
    uint32_t ph = 0;
    const int n_chunks = n_steps * C::kNC;
    for (int g = 0; g < n_chunks; ++g) {
      const int cb = g % C::kQBBufs;
      mbar_wait(&bars->q32_full[st], ph);
      mbar_wait(&bars->qb_empty[cb], ((g / C::kQBBufs) & 1) ^ 1);  // the MMAs that read this operand buffer have completed
      tc_fence_after();
      const unsigned char* src = sQ32 + st * C::kQ32Bytes;
      unsigned char* dhi = sQB + cb * C::kQBBytes;
      unsigned char* dlo = dhi + 128 * 128;
      // loads of a batch of items first, then convert + store (the stores are asm volatile: loads never move above them)
      constexpr int kBatch = (kItems % 5 == 0) ? 5 : 4;
#pragma unroll
      for (int it0 = 0; it0 < kItems; it0 += kBatch) {
        float4 x[kBatch], y[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const float4* p = reinterpret_cast<const float4*>(src + (ct + (it0 + u) * kTConvThreads) * 32);
          x[u] = p[0]; y[u] = p[1];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int gi = ct + (it0 + u) * kTConvThreads;
          const int r = gi / kPieces, c16 = gi - r * kPieces;
          const float f[8] = {x[u].x, x[u].y, x[u].z, x[u].w, y[u].x, y[u].y, y[u].z, y[u].w};
          uint32_t uh[4], ul[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uh[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
            const float h0 = __uint_as_float(uh[i] << 16), h1 = __uint_as_float(uh[i] & 0xFFFF0000u);
            ul[i] = pack_bf16(f[2 * i] - h0, f[2 * i + 1] - h1);
          }
          const int off = r * 128 + ((c16 ^ (r & 7)) << 4);  // 128B swizzle: 16-byte piece index XOR row-in-atom
          st_shared_v4(dhi + off, uh[0], uh[1], uh[2], uh[3]);
          st_shared_v4(dlo + off, ul[0], ul[1], ul[2], ul[3]);
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's shared-memory reads
      mbar_arrive(&bars->qb_full[cb]);
      mbar_arrive(&bars->q32_empty[st]);
      if (++st == n_stages) { st = 0; ph ^= 1u; }
    }
  } else {
    // ============================== softmax warpgroup(s) (thread == query row) ==============================
    const int wg = warp >> 2;  // takes the steps s with s % kWGs == wg
    const int row = tid & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    float* myacc = xacc + static_cast<size_t>(wg) * QT * nh * 128 + row;  // few: + (qt * nh + t) * 128
    constexpr int kAcc = kAll ? kTMPad : 1;
    float acc[kAcc];  // kAll: per-row heat accumulator over this CTA's heads
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
    for (int i = 0; i < QT * nh; ++i) myacc[i * 128] = 0.f;

    auto drain_o = [&](int g) {  // O of step g: TMEM -> OutT -> global
      const int ob = g & 1, ghl = g / n_qt, gq = g - ghl * n_qt;
      const int n = (tile0 + gq) * 128 + row;
      mbar_wait(&bars->pv_done[ob], (g >> 1) & 1);
      tc_fence_after();
      OutT* orow = out + (static_cast<long long>(b) * N + n) * (H * D) + (h_begin + ghl) * D;
#pragma unroll
      for (int c = 0; c < C::kDP / 16; ++c) {
        float o[16];
        tmem_ld16(tmem + lane_base + C::kColO + ob * C::kDP + c * 16, o);
        tmem_wait_ld();
        if (n < N) {
          if constexpr (std::is_same<OutT, float>::value) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              if (c * 16 + q4 * 4 + 4 <= D)
                *reinterpret_cast<float4*>(orow + c * 16 + q4 * 4) = make_float4(o[q4 * 4], o[q4 * 4 + 1], o[q4 * 4 + 2], o[q4 * 4 + 3]);
          } else {
            uint4 lo, hi;
            lo.x = pack_bf16(o[0], o[1]); lo.y = pack_bf16(o[2], o[3]); lo.z = pack_bf16(o[4], o[5]); lo.w = pack_bf16(o[6], o[7]);
            hi.x = pack_bf16(o[8], o[9]); hi.y = pack_bf16(o[10], o[11]); hi.z = pack_bf16(o[12], o[13]); hi.w = pack_bf16(o[14], o[15]);
            if (c * 16 + 8 <= D) *reinterpret_cast<uint4*>(orow + c * 16) = lo;
            if (c * 16 + 16 <= D) *reinterpret_cast<uint4*>(orow + c * 16 + 8) = hi;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->o_free[ob]);
    };

    int prev_s = -1;
    int hl = 0, qt = wg;  // step s = hl * n_qt + qt, advanced by kWGs per iteration
    while (qt >= n_qt && n_qt > 0) { qt -= n_qt; ++hl; }
    for (int s = wg; s < n_steps; s += kWGs) {
      const int sb = s & 1, ss = s % C::kSSlots, h = h_begin + hl;   // P / O slot of this warpgroup, score slot of this step
      const int n = (tile0 + qt) * 128 + row;
      mbar_wait(&bars->s_full[ss], (s / C::kSSlots) & 1);
      tc_fence_after();
      float sv[96];
      float sel[kAll ? 1 : kTFew];
      const uint32_t s_taddr = tmem + lane_base + ss * kTSlot;
      tmem_ld32(s_taddr, sv);
      tmem_ld32(s_taddr + 32, sv + 32);
      tmem_ld16(s_taddr + 64, sv + 64);
      if (!kAll && want_heat) {
#pragma unroll
        for (int t = 0; t < kTFew; ++t)
          if (t < tl.n) sel[t] = tmem_ld1(s_taddr + tl.idx[t]);
      }
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&bars->s_free[ss]);  // S(s) is in registers
      if (M >= 64) {
#pragma unroll
        for (int i = 64; i < kTMPad; ++i)
          if (i >= M) sv[i] = -INFINITY;
      } else {
#pragma unroll
        for (int i = 0; i < kTMPad; ++i)
          if (i >= M) sv[i] = -INFINITY;
      }
      static_assert((kTMPad - 6) % 4 == 2, "max reduction below assumes 80 columns");
      float mx0 = fmax3(sv[0], sv[1], sv[2]), mx1 = fmax3(sv[3], sv[4], sv[5]);
#pragma unroll
      for (int i = 6; i + 4 <= kTMPad; i += 4) {
        mx0 = fmax3(mx0, sv[i], sv[i + 1]); mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
      }
      mx0 = fmax3(mx0, sv[kTMPad - 2], sv[kTMPad - 1]);
      const float m = fmaxf(mx0, mx1) * scale_log2;
      const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2), negm2 = pack_f32x2(-m, -m);
      uint64_t sum2 = pack_f32x2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kTMPad; i += 2) {
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
      for (int i = 0; i < kTMPad; i += 2)  // normalised probabilities: PV needs no later division
        unpack_f32x2(fmul2(pack_f32x2(sv[i], sv[i + 1]), inv2), sv[i], sv[i + 1]);
      if (want_heat) {
        if constexpr (kAll) {
#pragma unroll
          for (int i = 0; i < kTMPad; ++i) acc[i] += sv[i];
        } else {
#pragma unroll
          for (int t = 0; t < kTFew; ++t) {
            if (t < tl.n) {
              const float pt = ex2(fmaf(sel[t], scale_log2, -m)) * inv_l;  // same ops as sv[idx[t]] above
              if (tl.per_head) {  // DAAM-style: one plane per (batch, head, token), no head mean
                if (n < N) {
                  float* ptr = maps + ((static_cast<long long>(b - b_first) * H + h) * tl.n + t) * N + n;
                  if (accumulate) atomicAdd(ptr, pt);  // result-less RED: one thread per element and launch
                  else *ptr = pt;
                }
              } else {
                myacc[(qt * nh + t) * 128] += pt;
              }
            }
          }
        }
      }
#pragma unroll
      for (int i = kTMPad; i < 96; ++i) sv[i] = 0.f;
      // O of this warpgroup's previous step: drained after the exponentials of this one (its PV has long finished) and
      // before P(s) is stored — pv_done(prev) also says that PV has finished reading the P columns
      if (prev_s >= 0) drain_o(prev_s);
      // (single warpgroup: slot sb was last used by step s-2, whose PV completion was observed one step ago)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t u[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = pack_bf16(sv[c * 32 + 2 * i], sv[c * 32 + 2 * i + 1]);
        tmem_st16(tmem + lane_base + C::kColP + sb * C::kPStride + c * 16, u);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars->p_full[sb]);
      prev_s = s;
      qt += kWGs;
      while (qt >= n_qt) { qt -= n_qt; ++hl; }
    }
    if (prev_s >= 0) drain_o(prev_s);
    tc_fence_before();
    if constexpr (kAll) {
     if (want_heat) {
      // ---- all tokens: park the accumulator row in shared memory (every MMA and copy of this CTA has completed: the
      //      last PV was observed above), then one coalesced store per token plane ----
      const float inv_h = 1.0f / static_cast<float>(H);
      float* srow = reinterpret_cast<float*>(smem) + row * kTHeatLd;   // over the operand tiles (64 KB >= 128 x 81 x 4)
#pragma unroll
      for (int i = 0; i < kTMPad; ++i) srow[i] = acc[i];
      const int n = tile0 * 128 + row;
      if (n < N) {
        float* dst = maps + static_cast<long long>(b - b_first) * tl.n * N + n;
        for (int t = 0; t < tl.n; ++t) {
          const float val = srow[tl.idx[t]] * inv_h;
          float* ptr = dst + static_cast<long long>(t) * N;
          if (accumulate) atomicAdd(ptr, val);  // (RED: same value as load + add + store, no load)
          else *ptr = val;
        }
      }
     }
    }
    // ---- few tokens: warpgroup 0 adds the two warpgroups' head sums (fixed order) and writes the mean over heads ----
    if (want_heat && !kAll && !tl.per_head) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (wg == 0) {
        const float* other = myacc + static_cast<size_t>(QT) * nh * 128;
        const float inv_h = 1.0f / static_cast<float>(H);
        for (int qt = 0; qt < n_qt; ++qt) {
          const int n = (tile0 + qt) * 128 + row;
#pragma unroll
          for (int t = 0; t < kTFew; ++t) {
            if (t < tl.n) {
              const float sum = myacc[(qt * nh + t) * 128] + other[(qt * nh + t) * 128];
              if (gridDim.z > 1) {
                if (qt == 0) hsum[t] = sum;  // (cluster launches have QT == 1)
              } else if (n < N) {
                float* ptr = maps + (static_cast<long long>(b - b_first) * tl.n + t) * N + n;
                if (accumulate) atomicAdd(ptr, sum * inv_h);  // (RED: same value as load + add + store, no load)
                else *ptr = sum * inv_h;
              }
            }
          }
        }
      }
    }
  }
  if (!kAll && gridDim.z > 1 && want_heat && !tl.per_head) {
    // head groups -> one heat row: partial sums of ranks 1.. go to the leader's (now idle) operand tiles
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(blockIdx.z), hsz = static_cast<int>(gridDim.z);
    float* xc = reinterpret_cast<float*>(smem);  // [hsz - 1][kTFew][128]
    cluster.sync();  // every CTA of the cluster has finished its copies and MMAs
    if (warp < 4 && rank > 0) {
      float* remote = cluster.map_shared_rank(xc, 0) + (rank - 1) * (kTFew * 128) + tid;
#pragma unroll
      for (int t = 0; t < kTFew; ++t)
        if (t < tl.n) remote[t * 128] = hsum[t];
    }
    cluster.sync();
    if (warp < 4 && rank == 0 && tile0 * 128 + tid < N) {
      const float inv_h = 1.0f / static_cast<float>(H);
      float* dst = maps + static_cast<long long>(b - b_first) * tl.n * N + tile0 * 128 + tid;
#pragma unroll
      for (int t = 0; t < kTFew; ++t) {
        if (t < tl.n) {
          float sum = hsum[t];
          for (int r = 1; r < hsz; ++r) sum += xc[((r - 1) * kTFew + t) * 128 + tid];  // fixed order
          float* ptr = dst + static_cast<long long>(t) * N;
          if (accumulate) atomicAdd(ptr, sum * inv_h);
          else *ptr = sum * inv_h;
        }
      }
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- prompt-side packing: K fp32 [B,M,C] and V [B,M,C] -> per (batch, head) the shared-memory image of the kernel ----
// One thread per 16-byte piece (8 bf16) of the blob.  Tile t of a (batch, head) block: t < kNC: K_hi chunk t;
// t < 2 kNC: K_lo chunk t - kNC; else V chunk t - 2 kNC (64 columns).  Piece p of row r lands at piece p ^ (r & 7).
template <int D, typename VT>
__global__ void pack_context_kv_kernel(const float* __restrict__ k32, const VT* __restrict__ v, uint4* __restrict__ blob,
                                       int B, int H, int M) {
  using C = TCfg<D>;
  constexpr int kTiles = 2 * C::kNC + C::kVC;
  const long long total = static_cast<long long>(B) * H * kTiles * kTMPad * 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i & 7);
    long long rest = i >> 3;
    const int r = static_cast<int>(rest % kTMPad); rest /= kTMPad;
    const int t = static_cast<int>(rest % kTiles); rest /= kTiles;
    const int h = static_cast<int>(rest % H);
    const int b = static_cast<int>(rest / H);
    const int sp = p ^ (r & 7);  // source piece whose data is stored at position p
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.f;
    bool lo = false;
    if (r < M) {
      if (t < 2 * C::kNC) {
        const int c = (t < C::kNC) ? t : t - C::kNC;
        lo = (t >= C::kNC);
        const int col0 = sp * 8;  // within the chunk
        if (col0 < C::kW) {
          const float* src = k32 + (static_cast<long long>(b) * M + r) * (H * D) + h * D + c * C::kW + col0;
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = src[e];
        }
      } else {
        const int col0 = (t - 2 * C::kNC) * 64 + sp * 8;
        if (col0 < D) {
          const VT* src = v + (static_cast<long long>(b) * M + r) * (H * D) + h * D + col0;
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = static_cast<float>(src[e]);
        }
      }
    }
    uint32_t u[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t hi = sm100::pack_bf16(f[2 * e], f[2 * e + 1]);
      if (lo) {
        const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
        hi = sm100::pack_bf16(f[2 * e] - h0, f[2 * e + 1] - h1);
      }
      u[e] = hi;
    }
    blob[i] = make_uint4(u[0], u[1], u[2], u[3]);
  }
}

}  // namespace sm100

typedef CUresult (*EncodeTiledFnX3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 [B, rows, H*d] viewed as (d, H, rows, B); box (box_cols, 1, box_rows, 1), dense rows (no swizzle), zero fill
static int make_head_map_f32(CUtensorMap* map, const void* base, int B, int H, int rows, int d, int box_cols, int box_rows) {
  const unsigned long long Cs = static_cast<unsigned long long>(H) * d;
  const TensorMapKey key = {base, {static_cast<unsigned long long>(d), static_cast<unsigned long long>(H),
                                   static_cast<unsigned long long>(rows), static_cast<unsigned long long>(B)},
                            {static_cast<unsigned long long>(d) * 4, Cs * 4, static_cast<unsigned long long>(rows) * Cs * 4},
                            {static_cast<unsigned>(box_cols), 1u, static_cast<unsigned>(box_rows), 1u}, 0, 4, 0};
  if (tensor_map_cache_get(key, map)) return AGENDA_OK;
  bind_primary_context();
  static EncodeTiledFnX3 enc = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFnX3>(p);
  }();
  if (!enc) return fail(AGENDA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t C = static_cast<cuuint64_t>(H) * d;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(d) * 4, C * 4, static_cast<cuuint64_t>(rows) * C * 4};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_cols), 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AGENDA_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 Q) failed (CUresult %d)", static_cast<int>(r));
  tensor_map_cache_put(key, *map);
  return AGENDA_OK;
}

template <int D, typename OutT, bool kAll>
static int launch_cross_x3(const float* q, const void* kv_blob, void* out, int B, int H, int N, int M, float scale,
                           const TokenList& tl, int b_first, float* maps, int accumulate, cudaStream_t stream, int q_hm) {
  using C = sm100::TCfg<D>;
  CUtensorMap mq = {};
  int rc;
  if (q_hm) {
    if (C::kW != 40) return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross_fwd_heat_x3_hm: chunk-major Q needs a head dim that is a multiple of 40 (d=%d)", D);
  } else if ((rc = make_head_map_f32(&mq, q, B, H, N, D, C::kW, 128)) != AGENDA_OK) return rc;
  const int n_tiles = (N + 127) / 128, sms = num_sms();
  const int n_heat = (kAll || maps == nullptr || tl.per_head) ? 0 : tl.n;
  // shared memory: fixed part (operand tiles, K / V buffers) + heat rows + as many fp32 Q stages as fit (2..6)
  auto stages_for = [&](int qt) {
    const int left = sm100::kTSmemMax - 1024 - 64 - static_cast<int>(sizeof(sm100::TBarriers)) - C::kFixedBytes -
                     2 * qt * n_heat * 128 * 4;
    return left / C::kQ32Bytes;
  };
  // query tiles per CTA: the smallest count that puts the launch into one wave (amortises the per-head K/V block),
  // as long as at least two Q stages still fit beside the heat rows
  int QT = 1;
  if (!kAll)
    while (QT < sm100::kTMaxQT && ((n_tiles + QT - 1) / QT) * B > sms && stages_for(QT + 1) >= 2) ++QT;
  int n_stages = stages_for(QT);
  if (n_stages > sm100::kTMaxStages) n_stages = sm100::kTMaxStages;
  if (n_stages < 2) return fail(AGENDA_ERR_UNSUPPORTED, "attn_cross_fwd_heat_x3: shared memory budget (d=%d)", D);
  dim3 grid((n_tiles + QT - 1) / QT, B, 1);
  // few (batch, query tile) pairs: split the heads over a cluster along z while the grid still fits one wave
  int hs = 1;
  if (!kAll && QT == 1 && !(maps && tl.per_head))
    while (hs < 4 && H % (hs * 2) == 0 && static_cast<long long>(grid.x) * grid.y * hs * 2 <= sms) hs *= 2;
  grid.z = hs;
  const size_t smem = 1024 + C::kFixedBytes + static_cast<size_t>(n_stages) * C::kQ32Bytes +
                      static_cast<size_t>(2) * QT * n_heat * 128 * 4 + sizeof(sm100::TBarriers) + 64;
  auto kern = sm100::attn_cross_sm100_x3_kernel<D, OutT, kAll>;
  AGENDA_DYN_SMEM(kern, sm100::kTSmemMax);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kAll ? 256 : 384); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = hs;
  cfg.attrs = attr; cfg.numAttrs = hs > 1 ? 1 : 0;
  AGENDA_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, static_cast<const unsigned char*>(kv_blob), static_cast<OutT*>(out), maps, tl,
                                 H, N, M, QT, n_stages, b_first, accumulate, scale * 1.4426950408889634f,
                                 q_hm ? q : static_cast<const float*>(nullptr)));
  AGENDA_LAUNCH_CHECK("attn_cross_sm100_x3_kernel");
  return AGENDA_OK;
}

template <int D>
static int pack_context(const float* k32, const void* v, int v_dtype, void* blob, int B, int H, int M, cudaStream_t st) {
  using C = sm100::TCfg<D>;
  const long long total = static_cast<long long>(B) * H * (2 * C::kNC + C::kVC) * sm100::kTMPad * 8;
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads < 4096 ? (total + threads - 1) / threads : 4096);
  if (v_dtype == AGENDA_F32)
    sm100::pack_context_kv_kernel<D, float><<<blocks, threads, 0, st>>>(k32, static_cast<const float*>(v),
                                                                        static_cast<uint4*>(blob), B, H, M);
  else
    sm100::pack_context_kv_kernel<D, __nv_bfloat16><<<blocks, threads, 0, st>>>(k32, static_cast<const __nv_bfloat16*>(v),
                                                                                static_cast<uint4*>(blob), B, H, M);
  AGENDA_LAUNCH_CHECK("pack_context_kv_kernel");
  return AGENDA_OK;
}

int build_token_list(const char* who, const int32_t* token_idx, int T, int M, TokenList* tl);

}  // namespace agenda

using namespace agenda;

extern "C" long long agenda_context_blob_bytes(int B, int H, int d) {
  long long per = 0;
  switch (d) {
    case 40: per = sm100::TCfg<40>::kKVBytes; break;
    case 64: per = sm100::TCfg<64>::kKVBytes; break;
    case 80: per = sm100::TCfg<80>::kKVBytes; break;
    case 160: per = sm100::TCfg<160>::kKVBytes; break;
    default: return fail(AGENDA_ERR_UNSUPPORTED, "context_blob_bytes: head dim %d not in {40,64,80,160}", d);
  }
  if (B <= 0 || H <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "context_blob_bytes: B=%d H=%d", B, H);
  return per * B * H;
}

extern "C" int agenda_pack_context_kv(const float* k32, const void* v, int v_dtype, void* blob, int B, int H, int M, int d,
                                      void* stream) {
  const char* who = "pack_context_kv";
  if (!k32 || !v || !blob) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (v_dtype != AGENDA_F32 && v_dtype != AGENDA_BF16) return fail(AGENDA_ERR_UNSUPPORTED, "%s: v_dtype %d", who, v_dtype);
  if (B <= 0 || H <= 0 || M <= 0 || d <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "%s: B=%d H=%d M=%d d=%d", who, B, H, M, d);
  if (M > sm100::kTMPad) return fail(AGENDA_ERR_UNSUPPORTED, "%s: M=%d > %d", who, M, sm100::kTMPad);
  if ((reinterpret_cast<uintptr_t>(k32) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(blob)) & 15)
    return fail(AGENDA_ERR_MISALIGNED, "%s: k32 / v / blob must be 16-byte aligned", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d) {
    case 40: return pack_context<40>(k32, v, v_dtype, blob, B, H, M, st);
    case 64: return pack_context<64>(k32, v, v_dtype, blob, B, H, M, st);
    case 80: return pack_context<80>(k32, v, v_dtype, blob, B, H, M, st);
    case 160: return pack_context<160>(k32, v, v_dtype, blob, B, H, M, st);
    default: return fail(AGENDA_ERR_UNSUPPORTED, "%s: head dim %d not in {40,64,80,160}", who, d);
  }
}

static int cross_x3_impl(const char* who, int q_hm, const float* q, const void* kv_blob, void* out, int out_dtype, int B, int H,
                         int N, int M, int d, float scale, const int32_t* token_idx, int T, int b_first, int per_head,
                         float* maps, int accumulate, void* stream) {
  if (!q || !kv_blob || !out) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (out_dtype != AGENDA_F32 && out_dtype != AGENDA_BF16) return fail(AGENDA_ERR_UNSUPPORTED, "%s: out_dtype %d", who, out_dtype);
  if (B <= 0 || H <= 0 || N <= 0 || M <= 0 || d <= 0 || B > 65535)
    return fail(AGENDA_ERR_BAD_SHAPE, "%s: B=%d H=%d N=%d M=%d d=%d", who, B, H, N, M, d);
  if (M > sm100::kTMPad) return fail(AGENDA_ERR_UNSUPPORTED, "%s: M=%d > %d", who, M, sm100::kTMPad);
  if (b_first < 0 || b_first > B) return fail(AGENDA_ERR_BAD_SHAPE, "%s: b_first=%d, B=%d", who, b_first, B);
  const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(kv_blob) | reinterpret_cast<uintptr_t>(out);
  if (al & 15) return fail(AGENDA_ERR_MISALIGNED, "%s: q / kv_blob / out must be 16-byte aligned", who);
  TokenList tl;
  tl.n = 0;
  tl.per_head = 0;
  if (maps != nullptr) {
    int rc = build_token_list(who, token_idx, T, M, &tl);
    if (rc != AGENDA_OK) return rc;
    if (per_head && tl.n > sm100::kTFew)
      return fail(AGENDA_ERR_UNSUPPORTED, "%s: per-head maps for T=%d > %d tokens", who, tl.n, sm100::kTFew);
    if (reinterpret_cast<uintptr_t>(maps) & 3) return fail(AGENDA_ERR_MISALIGNED, "%s: maps", who);
    tl.per_head = per_head ? 1 : 0;
  }
  float* mp = tl.n ? maps : nullptr;
  const bool all = mp != nullptr && tl.n > sm100::kTFew;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define AGENDA_X3(DD)                                                                                                     \
  case DD:                                                                                                                \
    if (all)                                                                                                              \
      return out_dtype == AGENDA_F32                                                                                      \
                 ? launch_cross_x3<DD, float, true>(q, kv_blob, out, B, H, N, M, scale, tl, b_first, mp, accumulate, st, q_hm)  \
                 : launch_cross_x3<DD, __nv_bfloat16, true>(q, kv_blob, out, B, H, N, M, scale, tl, b_first, mp, accumulate, st, q_hm); \
    return out_dtype == AGENDA_F32                                                                                        \
               ? launch_cross_x3<DD, float, false>(q, kv_blob, out, B, H, N, M, scale, tl, b_first, mp, accumulate, st, q_hm)   \
               : launch_cross_x3<DD, __nv_bfloat16, false>(q, kv_blob, out, B, H, N, M, scale, tl, b_first, mp, accumulate, st, q_hm);
  switch (d) {
    AGENDA_X3(40)
    AGENDA_X3(64)
    AGENDA_X3(80)
    AGENDA_X3(160)
    default: return fail(AGENDA_ERR_UNSUPPORTED, "%s: head dim %d not in {40,64,80,160}", who, d);
  }
#undef AGENDA_X3
}

extern "C" int agenda_attn_cross_fwd_heat_x3(const float* q, const void* kv_blob, void* out, int out_dtype, int B, int H,
                                             int N, int M, int d, float scale, const int32_t* token_idx, int T,
                                             int b_first, int per_head, float* maps, int accumulate, void* stream) {
  return cross_x3_impl("attn_cross_fwd_heat_x3", 0, q, kv_blob, out, out_dtype, B, H, N, M, d, scale, token_idx, T, b_first,
                       per_head, maps, accumulate, stream);
}

extern "C" int agenda_attn_cross_fwd_heat_x3_hm(const float* q_hm, const void* kv_blob, void* out, int out_dtype, int B, int H,
                                                int N, int M, int d, float scale, const int32_t* token_idx, int T,
                                                int b_first, int per_head, float* maps, int accumulate, void* stream) {
  return cross_x3_impl("attn_cross_fwd_heat_x3_hm", 1, q_hm, kv_blob, out, out_dtype, B, H, N, M, d, scale, token_idx, T, b_first,
                       per_head, maps, accumulate, stream);
}
