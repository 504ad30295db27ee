// K1 (v2): fused self-attention for sm_100a with TWO 128-query tiles per CTA and two softmax warpgroups that
// ping-pong on the MUFU/FMA pipes while the tensor core serves the other tile (FlashAttention-4 style schedule).
//
// At SD-1.x head dims (d = 40) a 128x128 score tile costs only ~384 tensor-pipe cycles (QK^T with K-extent 48 +
// PV with N-extent 48) but 16384 exponentials = 1024 MUFU cycles per SM, so the kernel is exponent-bound, not
// tensor-bound.  The schedule therefore keeps the MUFU busy: while warpgroup A runs softmax on S_A(j), the tensor
// core computes S_B(j) / P_B V(j-1), and vice versa.  Per-element instruction count is cut with packed fp32x2 math
// (FFMA2 / FADD2), 3-input max (FMNMX3) and the P tile going back to the tensor core through TMEM (TS-form MMA).
//
// 10 warps: 0-3 softmax WG A (query rows q0..q0+127), 4-7 softmax WG B (q0+128..q0+255), 8 TMA producer,
// 9 TMEM allocator + MMA issuer.  TMEM: S_A | S_B (BLOCK_N fp32 columns each) | P_A | P_B (BLOCK_N/2 columns of
// packed bf16 pairs) | O_A | O_B.  P has its own columns so that QK(t, j+1) can be issued as soon as the softmax
// warpgroup has pulled S(t, j) into registers (s_free), i.e. the next score tile is computed WHILE the current one
// is being exponentiated, and PV(t, j) runs under softmax(t, j+1): the softmax warpgroups never wait on the MMA.
// MMA order: QK_A(0) QK_B(0) | QK_A(1) PV_A(0) QK_B(1) PV_B(0) | QK_A(2) PV_A(1) ...
#include <cstdlib>
#include <type_traits>

#include "sm100_common.cuh"

namespace agenda {
namespace sm100 {


// Optional event trace (tools/ubench/trace_attn.cu builds with -DAGENDA_V2_TRACE): CTA (0,0) records clock64() at
// the pipeline events of the first kTraceTiles key tiles.  Compiled out of the product library.
#ifdef AGENDA_V2_TRACE
constexpr int kTraceTiles = 24, kTraceEvents = 8, kTraceActors = 8;
__device__ long long g_v2_trace[kTraceActors * kTraceTiles * kTraceEvents];
#define V2_TRACE(actor, j, ev)                                                                             \
  do {                                                                                                     \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (actor) < kTraceActors &&                           \
        (j) < kTraceTiles)                                                                                 \
      g_v2_trace[((actor) * kTraceTiles + (j)) * kTraceEvents + (ev)] = clock64();                         \
  } while (0)
#else
#define V2_TRACE(actor, j, ev) do { } while (0)
#endif
constexpr int kV2MaxTiles = 3;

// NT query tiles of 128 rows per CTA (one softmax warpgroup each), keys in tiles of BN.
//   NT = 2, BN = 128: two warpgroups ping-pong (2 softmax warps per SM sub-partition)
//   NT = 3, BN = 64 : three warpgroups (3 softmax warps per sub-partition) — more warps to cover tcgen05.ld /
//                     mbarrier / MUFU latencies at small head dims, at the price of twice as many (half-size) tiles
//   KS = 2 (with NT = 2, BN = 128): TWO warpgroups per query tile, each thread owns half a row (BN/2 columns);
//                     4 softmax warps per sub-partition while the MMAs keep their efficient 128-key shape.  The two
//                     halves of a row agree on the running max through shared memory + a 256-thread named barrier.
template <int D, int NT_, int BN_, int KS_ = 1>
struct V2Cfg {
  static constexpr int kNT = NT_;
  static constexpr int kKS = KS_;
  static constexpr int kThreads = NT_ * KS_ * 128 + 32 + 32 * NT_;  // softmax warpgroups, TMA warp, one MMA warp per query tile
  static constexpr int kDP = (D + 15) / 16 * 16;
  static constexpr int kChunks = (D + 63) / 64;
  static constexpr int kBlockN = BN_;
  // without TMEM room for separate P columns (e.g. d = 80: 2*128 + 2*64 + 2*80 > 512) P overwrites the S buffer
  // it came from and QK(t, j+1) is issued after PV(t, j) (in-order MMA pipe).
  static constexpr bool kAliasP = (NT_ * kBlockN * 3 / 2 + NT_ * kDP > 512);
  static constexpr int kQTileBytes = kChunks * 128 * 128;
  static constexpr int kKVBytes = kChunks * kBlockN * 128;
  static constexpr int kStages = (kKVBytes * 2 * 3 + NT_ * kQTileBytes <= 200 * 1024) ? 3 : 2;
  static constexpr int kColS = 0;                                       // + t * kBlockN
  static constexpr int kColP = kAliasP ? 0 : NT_ * kBlockN;              // + t * kPStride
  static constexpr int kPStride = kAliasP ? kBlockN : kBlockN / 2;
  static constexpr int kColO = kAliasP ? NT_ * kBlockN : NT_ * kBlockN * 3 / 2;  // + t * kDP
  // A head dim that is not a multiple of 16 leaves zero-padded V columns inside the PV MMA's N extent: column D of
  // every V tile is set to 1.0, so O[:, D] accumulates the softmax denominator sum_j P (in fp32, from the same
  // bf16-rounded P the numerator uses) and the softmax warps drop one packed add per pair of scores.
  static constexpr bool kSumInMma = (kDP > D);
  // Q(t) as the A operand FROM TMEM (TS-form QK^T) where the columns are free (d = 40, three tiles: 432 + 72 = 504): the
  // SS form re-reads the 128 x kDP Q tile from shared memory for every key tile (4 KB per K = 16 step, more than the
  // K tile itself) — 48 instead of 32 tensor cycles per MMA at 64 keys, and under the board's power cap (DESIGN.md K1
  // item 6) every byte not moved is time.  kDP / 2 columns of packed bf16 pairs per tile.
  static constexpr int kColQ = kColO + NT_ * kDP;  // + t * kDP / 2
  static constexpr bool kQInTmem = (KS_ == 1) && (kColQ + NT_ * kDP / 2 <= 512);
  static_assert(kColO + NT_ * kDP <= 512, "TMEM overflow");
  static_assert(NT_ <= kV2MaxTiles, "too many query tiles");
};

struct V2Barriers {
  float xmax[2][2][2][128];  // [tile parity][query tile][column half][row]: row-max exchange (KS = 2 only)
  uint64_t q_full, q_tmem[kV2MaxTiles];
  uint64_t k_full[3], k_empty[3], v_full[3], v_empty[3];
  uint64_t s_full[kV2MaxTiles], s_free[kV2MaxTiles], p_full[kV2MaxTiles], pv_done[kV2MaxTiles];
  uint32_t tmem_base;
  int bad_rows;  // fast path: some row of this CTA needs the exact-maximum kernel
};

template <class C>
constexpr size_t v2_smem_bytes() {
  return 1024 + C::kNT * C::kQTileBytes + 2 * C::kStages * C::kKVBytes + sizeof(V2Barriers) + 64;
}

// kEmu: share of exponential pairs evaluated by Ex2Emu: 0 none, 2 -> 50 %, 3 -> 37.5 %, 4 -> 25 %, 8 -> 12.5 %
//
// kFast: the first pass does not track the running row maximum.  Every exponential of a row is taken relative to the
// maximum of its FIRST key tile (softmax is shift-invariant; bf16 / fp32 keep full relative precision over 2^+-100),
// which removes the max reduction (0.5 instructions per score), the rescale vote and the O correction from all later
// tiles.  The row sum tells whether that was legitimate: a row whose later scores exceed the first tile's maximum by
// 2^100, or whose sum underflows, ends with l outside [2^-80, 2^100] (or NaN).  A CTA with such a row re-initialises
// its barriers and runs a second, exact-maximum pass over its own query tiles (same code, kFast switched off).
// kUnit (with kFast): the caller has folded scale * log2(e) into Q (the processor scales W_q once), so the scores ARE
// the exponents: the fast pass uses reference 0 instead of the first tile's maximum and the MUFU columns need no FMA
// at all; the row-sum proof is the same (|exponent| beyond ~100 sends the CTA into the exact pass).
template <int D, int kEmu, int NT, int BN_, int KS, bool kFast, bool kUnit = false>
__global__ void __launch_bounds__(NT * KS * 128 + 32 + 32 * NT, 1)
attn_self_sm100_v2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                          const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ out, int H, int N,
                          float scale_log2, int issue_order, float* __restrict__ lse_out) {
  using C = V2Cfg<D, NT, BN_, KS>;
  constexpr int BN = C::kBlockN;
  constexpr int ST = C::kStages;
  constexpr int kTmaWarp = 4 * NT * KS, kMmaWarp = 4 * NT * KS + 1;
  static_assert(KS == 1 || (KS == 2 && NT <= 2 && !C::kAliasP), "column split needs separate P columns");
  static_assert(!kFast || KS == 1, "the fast path is for one warpgroup per query tile");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                               // NT query tiles
  unsigned char* sK = sQ + NT * C::kQTileBytes;           // ST stages
  unsigned char* sV = sK + ST * C::kKVBytes;              // ST stages
  V2Barriers* bars = reinterpret_cast<V2Barriers*>(sV + ST * C::kKVBytes);

  const int tid = threadIdx.x, warp = tid >> 5;
  // 1-D grid, longest work first: the CTAs whose NT query tiles are all inside the sequence come first (n_full per
  // (batch, head)), the partly filled last CTA of every (batch, head) at the end, where they fill the tail of the
  // last wave instead of being sprinkled through it (N = 4096, NT = 3: 1280 full + 128 two-tile CTAs on 148 SMs).
  const int n_full_ctas = N / (128 * NT);
  const int n_bh = gridDim.x / (n_full_ctas + ((N % (128 * NT)) ? 1 : 0));
  int bh, qt;
  if (static_cast<int>(blockIdx.x) < n_bh * n_full_ctas) {
    bh = blockIdx.x / n_full_ctas;
    qt = blockIdx.x - bh * n_full_ctas;
  } else {
    bh = blockIdx.x - n_bh * n_full_ctas;
    qt = n_full_ctas;
  }
  const int q0 = qt * (128 * NT);
  const int b = bh / H, h = bh - b * H;
  const int n_tiles = (N + BN - 1) / BN;
  const int nt = min(NT, (N - q0 + 127) / 128);  // query tiles of this CTA that hold at least one row
  for (int pass = 0;; ++pass) {
  const bool fast = kFast && pass == 0;
  if (tid == kTmaWarp * 32) {
    if (pass == 0) {
      bars->bad_rows = 0;
      tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    } else {  // second pass: every barrier is quiescent (see the drain at the end of the TMA warp's loop)
      mbar_inval(&bars->q_full);
      for (int s = 0; s < 3; ++s) {
        mbar_inval(&bars->k_full[s]); mbar_inval(&bars->k_empty[s]);
        mbar_inval(&bars->v_full[s]); mbar_inval(&bars->v_empty[s]);
      }
      for (int t = 0; t < NT; ++t) {
        mbar_inval(&bars->s_full[t]); mbar_inval(&bars->s_free[t]);
        mbar_inval(&bars->p_full[t]); mbar_inval(&bars->pv_done[t]); mbar_inval(&bars->q_tmem[t]);
      }
    }
    mbar_init(&bars->q_full, 1);
    const int releasers = ((issue_order & 3) == 2) ? nt : 1;  // MMA warps that must have consumed a K / V stage
    for (int s = 0; s < 3; ++s) {
      mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], releasers);
      mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], releasers);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(&bars->s_full[t], 1); mbar_init(&bars->s_free[t], 128 * KS);
      mbar_init(&bars->p_full[t], 128 * KS); mbar_init(&bars->pv_done[t], 1);
      mbar_init(&bars->q_tmem[t], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp && pass == 0) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == kTmaWarp) {
    // ============================== TMA producer (warp converged, one elected lane issues) ==============================
    if (elect_one()) {
      mbar_expect_tx(&bars->q_full, nt * C::kQTileBytes);
      for (int t = 0; t < nt; ++t)
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(&map_q, &bars->q_full, sQ + t * C::kQTileBytes + c * 128 * 128, c * 64, h, q0 + t * 128, b);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&bars->k_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&bars->k_full[s], C::kKVBytes);
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(&map_k, &bars->k_full[s], sK + s * C::kKVBytes + c * BN * 128, c * 64, h, j * BN, b);
      }
      __syncwarp();
      mbar_wait(&bars->v_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&bars->v_full[s], C::kKVBytes);
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(&map_v, &bars->v_full[s], sV + s * C::kKVBytes + c * BN * 128, c * 64, h, j * BN, b);
      }
      __syncwarp();
      if (++s == ST) { s = 0; ph ^= 1u; }
    }
    if (kFast) {
      // drain: the last releases of the K / V stages (tcgen05.commit arrivals nobody waits for) must have landed
      // before a second pass may invalidate the barriers
      for (int i = 0; i < ST; ++i) {
        mbar_wait(&bars->k_empty[s], ph ^ 1);
        mbar_wait(&bars->v_empty[s], ph ^ 1);
        if (++s == ST) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= kMmaWarp) {
    // ============================== MMA issuer(s) ==============================
    // The whole warp stays converged (every lane polls the mbarriers); only the tcgen05.mma / tcgen05.commit
    // instructions sit under elect.sync.  With `if (lane == 0)` around the loop the compiler cannot prove the
    // operands warp-uniform and wraps every UTCHMMA in an R2UR + ELECT + BRA.U.ANY loop (~14 SASS instructions and
    // ~50 cycles per MMA): 22 MMAs per tile pair made the issuer itself the critical path behind s_full / pv_done.
    constexpr uint32_t idesc_qk = make_idesc(128, BN, 0);
    constexpr uint32_t idesc_pv = make_idesc(128, C::kDP, 1);
    const uint64_t q_desc = make_sdesc(smem_u32(sQ), 16, 1024);
    const uint64_t k_desc = make_sdesc(smem_u32(sK), 16, 1024);
    const uint64_t v_desc = make_sdesc(smem_u32(sV), BN * 128, 1024);
    // byte offsets are added to the 14-bit (addr >> 4) field; shared memory is < 256 KB so it never carries out
    auto issue_qk = [&](int t, int s) {  // S(t) = Q(t) K(stage s)^T
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < C::kDP / 16; ++kk) {
          const uint64_t adesc = q_desc + static_cast<uint64_t>((t * C::kQTileBytes + (kk >> 2) * 128 * 128 + (kk & 3) * 32) >> 4);
          const uint64_t bdesc = k_desc + static_cast<uint64_t>((s * C::kKVBytes + (kk >> 2) * BN * 128 + (kk & 3) * 32) >> 4);
          if (C::kQInTmem) umma_ts(tmem + C::kColS + t * BN, tmem + C::kColQ + t * (C::kDP / 2) + kk * 8, bdesc, idesc_qk, kk != 0);
          else umma_ss(tmem + C::kColS + t * BN, adesc, bdesc, idesc_qk, kk != 0);
        }
        umma_commit(&bars->s_full[t]);
      }
      __syncwarp();
    };
    int trace_j = 0;
    (void)trace_j;
    auto issue_pv = [&](int t, int s, bool first) {  // O(t) (+)= P(t) V(stage s)
      tc_fence_after();  // caller has observed p_full(t, j) and v_full(stage)
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          const uint64_t bdesc = v_desc + static_cast<uint64_t>((s * C::kKVBytes + kk * 2048) >> 4);
          umma_ts(tmem + C::kColO + t * C::kDP, tmem + C::kColP + t * C::kPStride + kk * 8, bdesc, idesc_pv,
                  !(first && kk == 0));
        }
        umma_commit(&bars->pv_done[t]);
      }
      __syncwarp();
    };
    auto commit = [&](uint64_t* bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };
    // kSumInMma: V(stage s)[key r][column D] = 1.0 (bf16), in the 128B-swizzled layout TMA wrote (16-byte piece
    // (D*2/16) ^ (r & 7) of row r); generic-proxy stores, made visible to the tensor core by fence.proxy.async
    auto set_ones_column = [&](int s) {
      if (C::kSumInMma) {
        constexpr int kPiece = (D % 64) * 2 / 16, kInPiece = (D % 64) * 2 % 16;
        unsigned char* vs = sV + s * C::kKVBytes + (D / 64) * BN * 128;
        for (int r = (tid & 31); r < BN; r += 32)
          *reinterpret_cast<unsigned short*>(vs + r * 128 + ((kPiece ^ (r & 7)) << 4) + kInPiece) = 0x3F80;
        fence_proxy_async_smem();
        __syncwarp();
      }
    };
    // s = stage of tile j, s1 = stage of tile j+1; ph / ph1 their ring phases
    int s = 0, s1 = (ST > 1) ? 1 : 0;
    uint32_t ph = 0, ph1 = (ST > 1) ? 0u : 1u;
    if ((issue_order & 3) == 2) {
      // One MMA warp per query tile: each issues QK(t, .) / PV(t, .) for its own tile only, so no tile's MMAs queue
      // behind a barrier of another softmax warpgroup (head-of-line blocking of the single in-order issuer).  The
      // tensor pipe serialises the instructions of the warps; a K / V stage is released once every tile's warp has
      // committed behind its last MMA that reads it (k_empty / v_empty count = number of active tiles).
      const int t = warp - kMmaWarp;
      if (t < nt) {
        mbar_wait(&bars->q_full, 0);
        if (C::kQInTmem) mbar_wait(&bars->q_tmem[t], 0);  // the softmax warpgroup has copied Q(t) into TMEM
        mbar_wait(&bars->k_full[0], 0);
        tc_fence_after();
        issue_qk(t, 0);
        commit(&bars->k_empty[0]);
        for (int j = 0; j < n_tiles; ++j) {
          const bool more = (j + 1 < n_tiles);
          if (more) mbar_wait(&bars->k_full[s1], ph1);
          if (!C::kAliasP && more) {
            mbar_wait(&bars->s_free[t], j & 1);  // S(t, j) is in the softmax warpgroup's registers
            tc_fence_after();
            issue_qk(t, s1);
          }
          mbar_wait(&bars->v_full[s], ph);
          set_ones_column(s);  // (idempotent: every tile's warp writes the same ones before its own PV)
          mbar_wait(&bars->p_full[t], j & 1);
          issue_pv(t, s, j == 0);
          if (C::kAliasP && more) issue_qk(t, s1);  // P(t, j) lived in S(t): QK(t, j+1) may only follow PV(t, j)
          if (more) commit(&bars->k_empty[s1]);
          commit(&bars->v_empty[s]);
          s = s1; ph = ph1;
          if (++s1 == ST) { s1 = 0; ph1 ^= 1u; }
        }
      }
    } else if (warp == kMmaWarp) {
    mbar_wait(&bars->q_full, 0);
    if (C::kQInTmem)
      for (int t = 0; t < nt; ++t) mbar_wait(&bars->q_tmem[t], 0);
    mbar_wait(&bars->k_full[0], 0);
    tc_fence_after();
    for (int t = 0; t < nt; ++t) issue_qk(t, 0);
    commit(&bars->k_empty[0]);
    if (C::kAliasP) {
      // P(t, j) lives in S(t): fixed order PV_A(j) QK_A(j+1) PV_B(j) QK_B(j+1)
      for (int j = 0; j < n_tiles; ++j) {
        const bool more = (j + 1 < n_tiles);
        if (more) mbar_wait(&bars->k_full[s1], ph1);
        mbar_wait(&bars->v_full[s], ph);
        set_ones_column(s);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (t < nt) {
            mbar_wait(&bars->p_full[t], j & 1);
            issue_pv(t, s, j == 0);
            if (more) issue_qk(t, s1);
          }
        }
        if (more) commit(&bars->k_empty[s1]);
        commit(&bars->v_empty[s]);
        s = s1; ph = ph1;
        if (++s1 == ST) { s1 = 0; ph1 ^= 1u; }
      }
    } else {
      // issue_order 0: QK_A(j+1) PV_A(j) QK_B(j+1) PV_B(j) — keeps the warpgroups out of phase (one is in its exponential
      //   stretch while the other loads / reduces), but QK_B(j+1) queues behind warpgroup A's exponentials.
      // issue_order 1: QK_A(j+1) QK_B(j+1) | PV_A(j) PV_B(j) — no score tile waits on another warpgroup; the warpgroups
      //   drift into phase and share the MUFU.
      for (int j = 0; j < n_tiles; ++j) {
        const bool more = (j + 1 < n_tiles);
        if (more) mbar_wait(&bars->k_full[s1], ph1);
        if ((issue_order & 3) == 1) {
          if (more) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              if (t < nt) {
                mbar_wait(&bars->s_free[t], j & 1);  // S(t, j) is in the softmax warpgroup's registers
                tc_fence_after();
                issue_qk(t, s1);
              }
            }
            commit(&bars->k_empty[s1]);
          }
          mbar_wait(&bars->v_full[s], ph);
          set_ones_column(s);
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (t < nt) {
              mbar_wait(&bars->p_full[t], j & 1);
              issue_pv(t, s, j == 0);
            }
          }
        } else {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (t < nt) {
              if (more) {
                mbar_wait(&bars->s_free[t], j & 1);
                tc_fence_after();
                V2_TRACE(7, j + 1, 3 + t);
                issue_qk(t, s1);
                V2_TRACE(6, j + 1, t);
              }
              if (t == 0) {
                mbar_wait(&bars->v_full[s], ph);
                set_ones_column(s);
              }
              mbar_wait(&bars->p_full[t], j & 1);
              V2_TRACE(7, j, t);
              issue_pv(t, s, j == 0);
              V2_TRACE(6, j, 3 + t);
            }
          }
          if (more) commit(&bars->k_empty[s1]);
        }
        commit(&bars->v_empty[s]);
        s = s1; ph = ph1;
        if (++s1 == ST) { s1 = 0; ph1 ^= 1u; }
      }
    }
    }  // single issuer
  } else if (warp < kTmaWarp) {
    // ============================== softmax warpgroups (thread == query row) ==============================
    const int wg = warp >> 2;
    const int t = wg / KS;      // query tile
    const int half = wg % KS;   // which CW-column slice of the tile's rows this warpgroup owns
    if (t < nt) {
    constexpr int CW = BN / KS;  // score columns per thread and key tile
    const int row = tid & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t s_taddr = tmem + lane_base + C::kColS + t * BN + half * CW;
    const uint32_t p_taddr = tmem + lane_base + C::kColP + t * C::kPStride + half * (C::kAliasP ? CW : CW / 2);
    const Ex2Emu ex2_emu;
    const uint32_t o_taddr = tmem + lane_base + C::kColO + t * C::kDP;
    // O column chunks (16 fp32 columns each) this thread rescales / writes out
    constexpr int kOChunks = C::kDP / 16;
    const int oc_begin = (KS == 1) ? 0 : (half == 0 ? 0 : (kOChunks + 1) / 2);
    const int oc_end = (KS == 1) ? kOChunks : (half == 0 ? (kOChunks + 1) / 2 : kOChunks);
    float m_used = (kUnit && fast) ? 0.f : -INFINITY;  // (kUnit fast pass: fixed reference 0)
    float l_run = 0.f;
    const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
    if (C::kQInTmem) {
      // Q(t) -> TMEM, the A-operand layout of a TS-form MMA (lane = row, one column = two consecutive bf16): this
      // thread's row out of the 128B-swizzled tile TMA wrote (16-byte piece p of row r sits at p ^ (r & 7)); the
      // pieces beyond D are the zero fill of the tensor map's out-of-bounds columns
      mbar_wait(&bars->q_full, 0);
      const unsigned char* qrow = sQ + t * C::kQTileBytes + row * 128;  // (+ 128 * 128 per 64-element chunk of the row)
      uint32_t qv[C::kDP / 2];
#pragma unroll
      for (int pc = 0; pc < C::kDP / 8; ++pc) {
        const uint4 v4 = *reinterpret_cast<const uint4*>(qrow + (pc >> 3) * 128 * 128 + (((pc & 7) ^ (row & 7)) << 4));
        qv[4 * pc] = v4.x; qv[4 * pc + 1] = v4.y; qv[4 * pc + 2] = v4.z; qv[4 * pc + 3] = v4.w;
      }
      const uint32_t q_taddr = tmem + lane_base + C::kColQ + t * (C::kDP / 2);
#pragma unroll
      for (int c = 0; c < C::kDP / 2; c += 8) tmem_st8(q_taddr + c, qv + c);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars->q_tmem[t]);
    }
    const bool pingpong = (NT == 2) && (nt == 2) && ((issue_order & 3) >= 2) && (issue_order & 4);
    if (pingpong && t == 1) asm volatile("bar.arrive %0, %1;" ::"r"(3), "r"(2 * KS * 128) : "memory");  // tile 0 goes first
    // One key tile.  kMasked is only instantiated for a ragged last tile: with a run-time test the compiler
    // if-converts the masking into an ISETP + FSEL per score on EVERY tile (2 of ~7 instructions per element).
    auto softmax_tile = [&](const int j, auto masked_c, auto fast_c) {
      constexpr bool kMasked = decltype(masked_c)::value;
      constexpr bool kSkipMax = decltype(fast_c)::value;  // fast pass: only the first tile's maximum is taken
      constexpr bool kNoShift = kUnit && kSkipMax;        // ... and not even that when the scores are the exponents
      V2_TRACE((warp < 6 ? warp : 99), j, 0);
      mbar_wait(&bars->s_full[t], j & 1);
      tc_fence_after();
      // Stagger the query tiles: tile t takes its first score tile t * stagger cycles late.  Two free-running tiles
      // otherwise sit in lockstep (both exponentiate, then both load / reduce / store) and the MUFU idles in between;
      // half a tile apart they fill each other's gaps (d = 64: 3.99 -> 3.80 ms; three tiles are insensitive).
      if (j == 0 && t > 0 && (issue_order >> 8) > 0) {
        const long long t_end = clock64() + static_cast<long long>(t) * (issue_order >> 8);
        while (clock64() < t_end) { }
      }
      V2_TRACE((warp < 6 ? warp : 99), j, 1);
      float sv[CW];
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) tmem_ld32(s_taddr + c * 32, sv + c * 32);
      if (CW % 32) tmem_ld16(s_taddr + (CW / 32) * 32, sv + (CW / 32) * 32);  // CW = 80: 32 + 32 + 16 columns
      tmem_wait_ld();
      if (!C::kAliasP) {
        tc_fence_before();
        mbar_arrive(&bars->s_free[t]);  // the tensor core may overwrite S(t) with QK(t, j+1) now
      }
      if (kMasked) {
        const int kv_left = N - j * BN - half * CW;
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (i >= kv_left) sv[i] = -INFINITY;
      }
      // PV(t, j-1) must have drained P(t) (single buffer) before P(t, j) is stored, and left O(t) quiescent before a
      // rescale touches it.  The rescale is rare, so the wait normally happens right before the first P store, after
      // the first 32 columns have been exponentiated (the PV MMA group needs ~500 cycles from p_full to pv_done).
      bool pv_waited = (j == 0);
      if (!kSkipMax || (j == 0 && !kNoShift)) {
      float mx;
      {
        float mx0 = fmax3(sv[0], sv[1], sv[2]), mx1 = fmax3(sv[3], sv[4], sv[5]);
        float mx2 = fmax3(sv[6], sv[7], sv[8]), mx3 = fmax3(sv[9], sv[10], sv[11]);
#pragma unroll
        for (int i = 12; i + 8 <= CW; i += 8) {
          mx0 = fmax3(mx0, sv[i], sv[i + 1]); mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
          mx2 = fmax3(mx2, sv[i + 4], sv[i + 5]); mx3 = fmax3(mx3, sv[i + 6], sv[i + 7]);
        }
        // the loop covers elements 12 .. 12 + 8*floor((CW-12)/8) - 1; the last (CW - 12) % 8 = 4 elements follow here
        // (CW = 128: 124..127, CW = 80: 76..79, CW = 64: 60..63)
        static_assert(CW % 16 == 0 && (CW - 12) % 8 == 4, "max reduction assumes CW % 16 == 0");
        mx0 = fmax3(mx0, sv[CW - 4], sv[CW - 3]);
        mx1 = fmax3(mx1, sv[CW - 2], sv[CW - 1]);
        mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      if (KS == 2) {
        // both halves of a row must use the same running max: swap partial maxima through shared memory
        bars->xmax[j & 1][t][half][row] = mx;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
        mx = fmaxf(mx, bars->xmax[j & 1][t][half ^ 1][row]);
      }
      const float m_new = mx * scale_log2;
      const bool need = m_new > m_used + kV2RescaleThreshold;
      V2_TRACE((warp < 6 ? warp : 99), j, 2);
      if (j == 0) {
        m_used = m_new;
      } else if (__any_sync(0xffffffffu, need)) {  // (identical in both halves: same rows, same m_new, same m_used)
        mbar_wait(&bars->pv_done[t], (j - 1) & 1);
        tc_fence_after();
        pv_waited = true;
        const float m_next = need ? m_new : m_used;
        const float f = ex2(m_used - m_next);
        l_run *= f;
        m_used = m_next;
#pragma unroll
        for (int c = 0; c < kOChunks; ++c) {
          if (c >= oc_begin && c < oc_end) {
            float o[16];
            tmem_ld16(o_taddr + c * 16, o);
            tmem_wait_ld();
            uint32_t u[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(o[i] * f);
            tmem_st16(o_taddr + c * 16, u);
          }
        }
        tmem_wait_st();
      }
      } else {
        V2_TRACE((warp < 6 ? warp : 99), j, 2);
      }
      // Exponential passes of the two query tiles alternate (FlashAttention-3 style named-barrier ping-pong): with the
      // tiles free-running, all softmax warps of an SM sub-partition drift into the same phase, the MUFU idles while
      // they all load / reduce and is then fought over.  Tile t enters its pass when tile t^1 has left its own.
      if (pingpong) asm volatile("bar.sync %0, %1;" ::"r"(3 + t), "r"(2 * KS * 128) : "memory");
      V2_TRACE((warp < 6 ? warp : 99), j, 3);
      const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
      const float emu_a = scale_log2 * (1.0f / 253.0f), emu_b = (126.0f - m_used) * (1.0f / 253.0f);
      uint64_t sum2a = pack_f32x2(0.f, 0.f), sum2b = sum2a;
      // kCols = 32 (or a 16-column tail when CW % 32 == 16) scores -> exponentials -> kCols/2 packed bf16 P columns
      auto exp_chunk = [&](const int col0, auto ncols_c) {
        constexpr int kCols = decltype(ncols_c)::value;
        uint32_t u[kCols / 2];
#pragma unroll
        for (int i = 0; i < kCols / 2; i += 2) {
          const int e = col0 + 2 * i;
          float a0, a1, b0, b1;
          if (kNoShift) {
            a0 = ex2(sv[e]); a1 = ex2(sv[e + 1]);
          } else {
            const uint64_t xa = ffma2(pack_f32x2(sv[e], sv[e + 1]), scale2, negm2);
            unpack_f32x2(xa, a0, a1);
            a0 = ex2(a0); a1 = ex2(a1);
          }
          // pair b of iteration ib = i/2 is emulated according to kEmu:
          // 2 -> every b pair (50 % of all exponentials), 3 -> ib % 3 != 2 (37.5 %), 4 -> even ib (25 %), 8 -> ib % 4 == 0
          const int ib = i >> 1;
          const bool emu_pair = (kEmu == 2) || (kEmu == 3 && (ib % 3) != 2) || (kEmu == 4 && (ib & 1) == 0) ||
                                (kEmu == 8 && (ib & 3) == 0);
          if (emu_pair) {
            ex2_emu(sv[e + 2], sv[e + 3], emu_a, emu_b, b0, b1);
          } else {
            if (kNoShift) {
              b0 = ex2(sv[e + 2]); b1 = ex2(sv[e + 3]);
            } else {
              const uint64_t xb = ffma2(pack_f32x2(sv[e + 2], sv[e + 3]), scale2, negm2);
              unpack_f32x2(xb, b0, b1);
              b0 = ex2(b0); b1 = ex2(b1);
            }
          }
          if (!C::kSumInMma) {
            sum2a = fadd2(sum2a, pack_f32x2(a0, a1));
            sum2b = fadd2(sum2b, pack_f32x2(b0, b1));
          }
          u[i] = pack_bf16(a0, a1);
          u[i + 1] = pack_bf16(b0, b1);
        }
        if (col0 == 0 && !pv_waited) {
          V2_TRACE((warp < 6 ? warp : 99), j, 6);
          mbar_wait(&bars->pv_done[t], (j - 1) & 1);
          tc_fence_after();
          V2_TRACE((warp < 6 ? warp : 99), j, 7);
        }
        if (kCols == 32) tmem_st16(p_taddr + col0 / 2, u);  // P as packed bf16 pairs
        else tmem_st8(p_taddr + col0 / 2, u);
      };
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) exp_chunk(c * 32, std::integral_constant<int, 32>{});
      if (CW % 32) exp_chunk((CW / 32) * 32, std::integral_constant<int, 16>{});
      if (pingpong && !(t == 1 && j + 1 == n_tiles))  // (tile 1's last pass has nobody left to release)
        asm volatile("bar.arrive %0, %1;" ::"r"(3 + (t ^ 1)), "r"(2 * KS * 128) : "memory");
      float s0, s1, s2, s3;
      unpack_f32x2(sum2a, s0, s1);
      unpack_f32x2(sum2b, s2, s3);
      l_run += (s0 + s1) + (s2 + s3);
      V2_TRACE((warp < 6 ? warp : 99), j, 4);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars->p_full[t]);
      V2_TRACE((warp < 6 ? warp : 99), j, 5);
    };
    const int n_full = N / BN;  // key tiles without padding
    if (kFast && fast) {
      for (int j = 0; j < n_full; ++j) softmax_tile(j, std::false_type{}, std::integral_constant<bool, kFast>{});
      if (n_full < n_tiles) softmax_tile(n_full, std::true_type{}, std::integral_constant<bool, kFast>{});
    } else {
      for (int j = 0; j < n_full; ++j) softmax_tile(j, std::false_type{}, std::false_type{});
      if (n_full < n_tiles) softmax_tile(n_full, std::true_type{}, std::false_type{});
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(&bars->pv_done[t], (n_tiles - 1) & 1);
    tc_fence_after();
    if (C::kSumInMma) {  // the denominator is column D of O
      float tmp[16];
      tmem_ld16(o_taddr + (D / 16) * 16, tmp);
      tmem_wait_ld();
      l_run = tmp[D % 16];
    } else if (KS == 2) {  // row sum = sum of the two halves' partial sums (both were scaled by the same running max)
      bars->xmax[0][t][half][row] = l_run;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      l_run += bars->xmax[0][t][half ^ 1][row];
    }
    const float inv_l = 1.0f / l_run;
    const int n = q0 + t * 128 + row;
    __nv_bfloat16* orow = out + (static_cast<long long>(b) * N + n) * (H * D) + h * D;
    // fast pass: was the first tile's maximum a legitimate reference for this row?  (a NaN sum fails both compares)
    if (fast && n < N && !(l_run > 0x1p-80f && l_run < 0x1p100f)) atomicOr(&bars->bad_rows, 1);
    // training mode: the row's base-2 log-sum-exp of the scaled scores, log2 sum_j 2^(scale log2e s_ij), for the backward
    // kernels (a row the fast pass got wrong is rewritten by the exact second pass of the same CTA)
    if (lse_out != nullptr && n < N && (KS == 1 || half == 0))
      lse_out[(static_cast<long long>(b) * H + h) * N + n] = m_used + log2f(l_run);
#pragma unroll
    for (int c = 0; c < kOChunks; ++c) {
      if (c >= oc_begin && c < oc_end) {
        float o[16];
        tmem_ld16(o_taddr + c * 16, o);
        tmem_wait_ld();
        if (n < N) {
          uint4 lo, hi;
          lo.x = pack_bf16(o[0] * inv_l, o[1] * inv_l); lo.y = pack_bf16(o[2] * inv_l, o[3] * inv_l);
          lo.z = pack_bf16(o[4] * inv_l, o[5] * inv_l); lo.w = pack_bf16(o[6] * inv_l, o[7] * inv_l);
          hi.x = pack_bf16(o[8] * inv_l, o[9] * inv_l); hi.y = pack_bf16(o[10] * inv_l, o[11] * inv_l);
          hi.z = pack_bf16(o[12] * inv_l, o[13] * inv_l); hi.w = pack_bf16(o[14] * inv_l, o[15] * inv_l);
          if (c * 16 + 8 <= D) *reinterpret_cast<uint4*>(orow + c * 16) = lo;
          if (c * 16 + 16 <= D) *reinterpret_cast<uint4*>(orow + c * 16 + 8) = hi;
        }
      }
    }
    tc_fence_before();
    }  // t < nt
  }
  __syncthreads();
  if (!fast || *reinterpret_cast<volatile int*>(&bars->bad_rows) == 0) {
    if (warp == kMmaWarp) {
      tc_fence_after();
      tmem_dealloc(tmem, 512);
    }
    break;
  }
  }  // pass
}

}  // namespace sm100

template <int D, int kEmu, int NT, int BN, int KS = 1, bool kFast = false, bool kUnit = false>
static int launch_v2(const void* q, const void* k, const void* v, void* out, int B, int H, int N, float scale,
                     cudaStream_t stream, long long ld, float* lse = nullptr) {
  using C = sm100::V2Cfg<D, NT, BN, KS>;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_head_map(&mq, q, B, H, N, D, 128, ld)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mk, k, B, H, N, D, C::kBlockN, ld)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mv, v, B, H, N, D, C::kBlockN, ld)) != AGENDA_OK) return rc;
  constexpr size_t smem = sm100::v2_smem_bytes<C>();
  auto kern = sm100::attn_self_sm100_v2_kernel<D, kEmu, NT, BN, KS, kFast, kUnit>;
  AGENDA_DYN_SMEM(kern, smem);
  dim3 grid(static_cast<unsigned>(((N + 128 * NT - 1) / (128 * NT)) * B * H));
  // measured on B200 (tools/bench_attn.py): one MMA warp per query tile (2) wins except where P aliases S (d = 80)
  int issue_order = C::kAliasP ? 0 : 2;
  if (const char* e = knob("AGENDA_V2_ORDER")) issue_order = atoi(e);  // experiments only (+4: exp-pass ping-pong)
  int stagger = (NT == 2 && !C::kAliasP) ? 800 : 0;  // cycles; measured on B200 (tools/bench_attn.py)
  if (const char* e = knob("AGENDA_V2_SKEW")) stagger = atoi(e);
  issue_order |= stagger << 8;
  kern<<<grid, C::kThreads, smem, stream>>>(mq, mk, mv, static_cast<__nv_bfloat16*>(out), H, N,
                                            kUnit ? 1.0f : scale * 1.4426950408889634f, issue_order, lse);
  AGENDA_LAUNCH_CHECK("attn_self_sm100_v2_kernel");
  return AGENDA_OK;
}

// emu: 0 = every exponential on the MUFU; 2/3/4/8 = 50/37.5/25/12.5 % of them on the FMA pipe.
// tiles: 2 = two 128-query tiles per CTA with 128-key tiles (64 for d = 160); 3 = three query tiles with 64-key
// tiles (d = 40 / 64 only); 4 = two query tiles, each served by two warpgroups owning half a row (d = 40 / 64).
int attn_self_sm100_v2(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int d, float scale,
                       int emu, int tiles, void* stream, long long ld, float* lse) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ---- shipped configurations (what agenda_attn_self_fwd / _strided dispatch to) ----
  if (d == 40 && tiles == 203 && emu == 3) return launch_v2<40, 3, 3, 64, 1, true, true>(q, k, v, out, B, H, N, scale, st, ld, lse);
  if (d == 40 && tiles == 103 && emu == 3) return launch_v2<40, 3, 3, 64, 1, true>(q, k, v, out, B, H, N, scale, st, ld, lse);
  if (d == 64 && tiles == 102 && emu == 4) return launch_v2<64, 4, 2, 128, 1, true>(q, k, v, out, B, H, N, scale, st, ld, lse);
  if (d == 80 && tiles == 105 && emu == 4) return launch_v2<80, 4, 2, 64, 1, true>(q, k, v, out, B, H, N, scale, st, ld, lse);
  if (d == 160 && tiles == 2 && emu == 4) return launch_v2<160, 4, 2, 64, 1>(q, k, v, out, B, H, N, scale, st, ld, lse);
  if (lse != nullptr) return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd_lse: only the shipped kernel configurations emit the log-sum-exp");
#ifndef AGENDA_VARIANTS
  return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: kernel variant (d=%d, tiles=%d, emu=%d) exists only in builds with "
              "-DAGENDA_VARIANTS (tests / measurements)", d, tiles, emu);
#else
  // ---- measurement / test variants (python -m agenda_b200.build --variants) ----
  if (tiles >= 100) {  // fast first pass with other emulation shares / tile shapes
    if (d == 40 && tiles == 103 && emu == 2) return launch_v2<40, 2, 3, 64, 1, true>(q, k, v, out, B, H, N, scale, st, ld);
    if (d == 40 && tiles == 103 && emu == 4) return launch_v2<40, 4, 3, 64, 1, true>(q, k, v, out, B, H, N, scale, st, ld);
    if (d == 64 && tiles == 102 && emu == 3) return launch_v2<64, 3, 2, 128, 1, true>(q, k, v, out, B, H, N, scale, st, ld);
    if (d == 64 && tiles == 102 && emu == 2) return launch_v2<64, 2, 2, 128, 1, true>(q, k, v, out, B, H, N, scale, st, ld);
    if (d == 80 && tiles == 102 && emu == 4) return launch_v2<80, 4, 2, 128, 1, true>(q, k, v, out, B, H, N, scale, st, ld);
    if (d == 160 && tiles == 102 && emu == 4) return launch_v2<160, 4, 2, 64, 1, true>(q, k, v, out, B, H, N, scale, st, ld);
    tiles -= 100;
  }
#define AGENDA_V2_EMU(DD, NT, BN, KS)                                                   \
    switch (emu) {                                                                      \
      case 0: return launch_v2<DD, 0, NT, BN, KS>(q, k, v, out, B, H, N, scale, st, ld);    \
      case 2: return launch_v2<DD, 2, NT, BN, KS>(q, k, v, out, B, H, N, scale, st, ld);    \
      case 3: return launch_v2<DD, 3, NT, BN, KS>(q, k, v, out, B, H, N, scale, st, ld);    \
      case 8: return launch_v2<DD, 8, NT, BN, KS>(q, k, v, out, B, H, N, scale, st, ld);    \
      default: return launch_v2<DD, 4, NT, BN, KS>(q, k, v, out, B, H, N, scale, st, ld);   \
    }
  if (tiles == 3) {
    switch (d) {
      case 40: AGENDA_V2_EMU(40, 3, 64, 1)
      case 64: AGENDA_V2_EMU(64, 3, 64, 1)
      default: return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: 3-tile variant needs head dim 40 or 64, got %d", d);
    }
  }
  if (tiles == 6) {  // three query tiles, 80-key tiles: the most keys per tile that fit TMEM at d = 40 (504 columns)
    switch (d) {
      case 40: AGENDA_V2_EMU(40, 3, 80, 1)
      default: return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: 3x80 variant is for head dim 40, got %d", d);
    }
  }
  if (tiles == 5) {  // two query tiles, 64-key tiles: P gets its own TMEM columns at d = 80 (no S/P aliasing)
    switch (d) {
      case 80: AGENDA_V2_EMU(80, 2, 64, 1)
      default: return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: 64-key variant is for head dim 80, got %d", d);
    }
  }
  if (tiles == 4) {  // two query tiles, two warpgroups (half rows) per tile
    switch (d) {
      case 40: AGENDA_V2_EMU(40, 2, 128, 2)
      case 64: AGENDA_V2_EMU(64, 2, 128, 2)
      default: return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: split-row variant needs head dim 40 or 64, got %d", d);
    }
  }
  switch (d) {
    case 40: AGENDA_V2_EMU(40, 2, 128, 1)
    case 64: AGENDA_V2_EMU(64, 2, 128, 1)
    case 80: AGENDA_V2_EMU(80, 2, 128, 1)
    case 160: AGENDA_V2_EMU(160, 2, 64, 1)
    default: return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: head dim %d not in {40,64,80,160}", d);
  }
#undef AGENDA_V2_EMU
#endif  // AGENDA_VARIANTS
}

#ifdef AGENDA_V2_TRACE
extern "C" int agenda_v2_trace_read(long long* host, int n) {
  const int total = sm100::kTraceActors * sm100::kTraceTiles * sm100::kTraceEvents;
  if (n < total) return -total;
  return cudaMemcpyFromSymbol(host, sm100::g_v2_trace, sizeof(long long) * total) == cudaSuccess ? total : -1;
}
#endif

}  // namespace agenda
