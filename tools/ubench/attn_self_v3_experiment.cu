// EXPERIMENT, not part of the product library (built by tools/ubench/bench_self.cu; results in DESIGN.md K1):
// with 20 launches it matches the shipped v2 kernel (0.545 - 0.555 ms at B*H = 128, N = 4096, d = 40), under sustained
// load it is 5 % slower (0.62 vs 0.58 ms): the kernel is power-bound there and this design moves more data per score.
//
// K1 (v3): fused self-attention for sm_100a at small head dims (d <= 64), built around what the event traces of v2
// showed (tools/ubench/trace_attn3.cu): at d = 40 a 64-key score tile holds ~350 cycles of exponential
// work per softmax warp, but the hand-shake s_free -> MMA-warp wake-up -> QK^T -> s_full -> softmax wake-up is a chain of
// four ~100-cycle barrier latencies plus the MMA pipeline fill, and with ONE score buffer per query tile that chain is
// on the critical path of every tile (21 % of the softmax warps' time in v2, and the MMA warp itself was ~1000 cycles
// behind).  v3 takes the chain off the critical path instead of shortening it:
//
//   * TWO 128-query tiles per CTA, 64-key tiles, and a ring of THREE score buffers per query tile in TMEM
//     (2 x 3 x 64 columns) + the two O accumulators (2 x 48 / 64) = 480 / 512 columns.  QK(t, j + 3) is issued right
//     behind PV(t, j), two whole key tiles before the softmax warpgroup asks for it.
//   * P(t, j) (bf16 pairs) overwrites the start of the score columns it came from (their owner holds them in
//     registers by then); PV(t, j) reads it from there (TS-form MMA) and QK(t, j + 3), issued by
//     the same thread after PV(t, j), overwrites the buffer — the tensor pipe executes one thread's MMAs in order, so
//     no s_free / p_free barrier exists at all.
//   * every per-tile barrier (s_full, p_full, pv_done) exists once per ring slot: a warpgroup may run two tiles ahead of
//     the tensor pipe and an mbarrier parity wait can only tell adjacent phases apart.
//   * one MMA-issue warp per query tile (blocking waits, no polling), one TMA warp that also writes the ones column
//     of V (row sums come out of the PV MMA at d = 40), K / V ring of six stages.
//
//   * KS = 2: TWO softmax warpgroups per query tile, each thread owns half a row (32 of the 64 score columns): four
//     softmax warps per SM sub-partition instead of two.  The exponential stream needs the MUFU (8 cycles per warp
//     instruction, 62.5 % of the scores) AND the FMA pipe (the emulated 37.5 %: ~8.5 cycles per score) almost
//     back to back (tools/ubench/mix.cu: one warp can hide 4 FFMA2 under each MUFU, but an in-order warp whose next
//     instruction is a MUFU stalls on a full MUFU queue with its FMA work behind it); two warps with identical
//     streams run in lockstep and leave both pipes ~60 % busy, four cover for each other.  In the fast pass the
//     halves never talk to each other (no running maximum; the row sum comes out of the PV MMA).
//
// warps: 8 * KS softmax (warpgroup w: query tile w / KS, column half w % KS), then TMA producer, two MMA issuers,
// TMEM allocator.
#include <cstdlib>
#include <type_traits>

#include "../../agenda_b200/csrc/sm100_common.cuh"

namespace agenda {
namespace sm100 {

// non-blocking probe, result as 0 / 1 (for callers that look at it much later than they ask)
__device__ __forceinline__ uint32_t mbar_test_u32(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done;
}
// four non-blocking probes in flight at once (a probe is a ~100-cycle shared-memory round trip; issued one after
// the other with a branch on each result a polling loop over several barriers spends most of its time waiting)
__device__ __forceinline__ uint32_t mbar_test4(uint64_t* b0, uint32_t p0, uint64_t* b1, uint32_t p1, uint64_t* b2, uint32_t p2,
                                               uint64_t* b3, uint32_t p3) {
  uint32_t bits;
  asm volatile(
      "{\n\t.reg .pred q0, q1, q2, q3;\n\t.reg .u32 r0, r1, r2, r3;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q2, [%5], %6;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q3, [%7], %8;\n\t"
      "selp.u32 r0, 1, 0, q0;\n\t"
      "selp.u32 r1, 2, 0, q1;\n\t"
      "selp.u32 r2, 4, 0, q2;\n\t"
      "selp.u32 r3, 8, 0, q3;\n\t"
      "or.b32 r0, r0, r1;\n\t"
      "or.b32 r2, r2, r3;\n\t"
      "or.b32 %0, r0, r2;\n\t}"
      : "=r"(bits)
      : "r"(smem_u32(b0)), "r"(p0), "r"(smem_u32(b1)), "r"(p1), "r"(smem_u32(b2)), "r"(p2), "r"(smem_u32(b3)), "r"(p3)
      : "memory");
  return bits;
}


#ifdef AGENDA_V2_TRACE
constexpr int kT3Tiles = 24, kT3Events = 8, kT3Actors = 8, kT3First = 20;
__device__ long long g_v3_trace[kT3Actors * kT3Tiles * kT3Events];
#define V3_TRACE(actor, j, ev)                                                                             \
  do {                                                                                                     \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (actor) < kT3Actors && (j) >= kT3First &&            \
        (j) < kT3First + kT3Tiles)                                                                         \
      g_v3_trace[((actor) * kT3Tiles + (j) - kT3First) * kT3Events + (ev)] = clock64();                    \
  } while (0)
#else
#define V3_TRACE(actor, j, ev) do { } while (0)
#endif

template <int D, int KS_>
struct V3Cfg {
  static constexpr int kKS = KS_;      // softmax warpgroups per query tile (each owns 64 / KS score columns of a row)
  static_assert(D <= 64 && D % 8 == 0, "v3 is for head dims up to 64");
  static constexpr int kNT = 2;        // query tiles (softmax warpgroups) per CTA
  static constexpr int kBlockN = 64;   // keys per tile
  static constexpr int kSBufs = 3;     // score buffers per query tile
  static constexpr int kStages = 6;    // K / V ring
  static constexpr int kThreads = 2 * KS_ * 128 + 128;
  static constexpr int kDP = (D + 15) / 16 * 16;
  static constexpr int kQTileBytes = 128 * 128;
  static constexpr int kKVBytes = kBlockN * 128;
  static constexpr int kColS = 0;                          // + (t * kSBufs + b) * kBlockN
  static constexpr int kColO = kNT * kSBufs * kBlockN;     // + t * kDP
  static constexpr bool kSumInMma = (kDP > D);             // column D of V = 1.0: O[:, D] is the softmax denominator
  static_assert(kColO + kNT * kDP <= 512, "TMEM overflow");
};

struct V3Barriers {
  uint64_t q_full;
  uint64_t k_full[6], k_empty[6], v_full[6], v_empty[6], v_ready[6];
  uint64_t s_full[2][3], p_full[2][3], pv_done[2][3];
  uint32_t tmem_base;
  int bad_rows;
};

template <class C>
constexpr size_t v3_smem_bytes() {
  return 1024 + C::kNT * C::kQTileBytes + 2 * C::kStages * C::kKVBytes + sizeof(V3Barriers) + 64;
}

// kEmu / kFast / kUnit: as in attn_self_sm100_v2.cu (share of exponentials on the FMA pipe; first pass without a
// running maximum + row-sum proof + exact second pass for the CTAs that need it; scores are base-2 exponents already).
template <int D, int kEmu, bool kFast, bool kUnit, int KS>
__global__ void __launch_bounds__(2 * KS * 128 + 128, 1)
attn_self_sm100_v3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                          const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ out, int H, int N,
                          float scale_log2) {
  using C = V3Cfg<D, KS>;
  constexpr int BN = C::kBlockN, ST = C::kStages, NT = C::kNT, SB = C::kSBufs;
  constexpr int kTmaWarp = 8 * KS, kMmaWarp = 8 * KS + 1, kAllocWarp = 8 * KS + 3;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + NT * C::kQTileBytes;
  unsigned char* sV = sK + ST * C::kKVBytes;
  V3Barriers* bars = reinterpret_cast<V3Barriers*>(sV + ST * C::kKVBytes);

  const int tid = threadIdx.x, warp = tid >> 5;
  // 1-D grid, full CTAs first, the partly filled last CTA of every (batch, head) at the end (see v2)
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
  const int nt = min(NT, (N - q0 + 127) / 128);
  for (int pass = 0;; ++pass) {
    const bool fast = kFast && pass == 0;
    if (tid == kTmaWarp * 32) {
      if (pass == 0) {
        bars->bad_rows = 0;
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
      } else {  // second pass: every barrier is quiescent (see the drain at the end of the TMA warp's loop)
        mbar_inval(&bars->q_full);
        for (int s = 0; s < ST; ++s) {
          mbar_inval(&bars->k_full[s]); mbar_inval(&bars->k_empty[s]); mbar_inval(&bars->v_full[s]);
          mbar_inval(&bars->v_empty[s]); mbar_inval(&bars->v_ready[s]);
        }
        for (int t = 0; t < NT; ++t)
          for (int s = 0; s < SB; ++s) {
            mbar_inval(&bars->s_full[t][s]); mbar_inval(&bars->p_full[t][s]); mbar_inval(&bars->pv_done[t][s]);
          }
      }
      mbar_init(&bars->q_full, 1);
      for (int s = 0; s < ST; ++s) {
        mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], nt);
        mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], nt);
        mbar_init(&bars->v_ready[s], 1);
      }
      for (int t = 0; t < NT; ++t)
        for (int s = 0; s < SB; ++s) {
          mbar_init(&bars->s_full[t][s], 1); mbar_init(&bars->p_full[t][s], 128 * KS); mbar_init(&bars->pv_done[t][s], 1);
        }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kAllocWarp && pass == 0) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == kTmaWarp) {
      // ============================== TMA producer ==============================
      if (elect_one()) {
        mbar_expect_tx(&bars->q_full, nt * C::kQTileBytes);
        for (int t = 0; t < nt; ++t) tma_load_4d(&map_q, &bars->q_full, sQ + t * C::kQTileBytes, 0, h, q0 + t * 128, b);
      }
      __syncwarp();
      // V tile jj has landed: set its ones column (16-byte piece (D*2/16) ^ (r & 7) of row r in the 128B-swizzled
      // layout TMA wrote; generic-proxy stores made visible to the tensor core by fence.proxy.async) and publish it
      auto finish = [&](int jj) {
        const int sf = jj % ST;
        mbar_wait(&bars->v_full[sf], (jj / ST) & 1);
        if (C::kSumInMma) {
          constexpr int kPiece = D * 2 / 16, kInPiece = D * 2 % 16;
          unsigned char* vs = sV + sf * C::kKVBytes;
          for (int r = (tid & 31); r < BN; r += 32)
            *reinterpret_cast<unsigned short*>(vs + r * 128 + ((kPiece ^ (r & 7)) << 4) + kInPiece) = 0x3F80;
          fence_proxy_async_smem();
          __syncwarp();
        }
        if (elect_one()) mbar_arrive(&bars->v_ready[sf]);
        __syncwarp();
      };
      int s = 0;
      uint32_t ph = 0;
      // (two tiles behind the loads: waiting for the tile just requested would leave one tile in flight at a time)
      for (int j = 0; j < n_tiles; ++j) {
        if (j > 1) finish(j - 2);
        mbar_wait(&bars->k_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&bars->k_full[s], C::kKVBytes);
          tma_load_4d(&map_k, &bars->k_full[s], sK + s * C::kKVBytes, 0, h, j * BN, b);
        }
        __syncwarp();
        mbar_wait(&bars->v_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&bars->v_full[s], C::kKVBytes);
          tma_load_4d(&map_v, &bars->v_full[s], sV + s * C::kKVBytes, 0, h, j * BN, b);
        }
        __syncwarp();
        if (++s == ST) { s = 0; ph ^= 1u; }
      }
      if (n_tiles > 1) finish(n_tiles - 2);
      finish(n_tiles - 1);
      if (kFast) {
        // drain: the last releases of the K / V stages (tcgen05.commit arrivals nobody waits for) must have landed
        // before a second pass may invalidate the barriers
        for (int i = 0; i < ST; ++i) {
          mbar_wait(&bars->k_empty[s], ph ^ 1);
          mbar_wait(&bars->v_empty[s], ph ^ 1);
          if (++s == ST) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
      // ============================== MMA issuers (one per query tile; warp converged, elect.sync around the MMAs) ===
      const int t = warp - kMmaWarp;
      if (t < nt) {
        constexpr uint32_t idesc_qk = make_idesc(128, BN, 0);
        constexpr uint32_t idesc_pv = make_idesc(128, C::kDP, 1);
        const uint64_t q_desc = make_sdesc(smem_u32(sQ + t * C::kQTileBytes), 16, 1024);
        const uint64_t k_desc = make_sdesc(smem_u32(sK), 16, 1024);
        const uint64_t v_desc = make_sdesc(smem_u32(sV), BN * 128, 1024);
        const uint32_t s_col = tmem + C::kColS + t * SB * BN, o_col = tmem + C::kColO + t * C::kDP;
        auto issue_qk = [&](int sb, int st) {  // S(t)[sb] = Q(t) K(stage st)^T
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < C::kDP / 16; ++kk)
              umma_ss(s_col + sb * BN, q_desc + static_cast<uint64_t>((kk * 32) >> 4),
                      k_desc + static_cast<uint64_t>((st * C::kKVBytes + kk * 32) >> 4), idesc_qk, kk != 0);
            umma_commit(&bars->s_full[t][sb]);
            umma_commit(&bars->k_empty[st]);
          }
          __syncwarp();
        };
        auto issue_pv = [&](int sb, int st, bool first) {  // O(t) (+)= P(t)[sb] V(stage st)
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BN / 16; ++kk)
              // P of key block kk (16 keys = 8 packed columns): KS = 1: columns kk * 8 of the buffer; KS = 2: each
              // warpgroup wrote its 32 keys over the start of ITS OWN 32 score columns (the other half's scores may
              // not have been read yet)
              umma_ts(o_col, s_col + sb * BN + (KS == 1 ? kk * 8 : (kk >> 1) * 32 + (kk & 1) * 8),
                      v_desc + static_cast<uint64_t>((st * C::kKVBytes + kk * 2048) >> 4), idesc_pv, !(first && kk == 0));
            umma_commit(&bars->pv_done[t][sb]);
            umma_commit(&bars->v_empty[st]);
          }
          __syncwarp();
        };
        mbar_wait(&bars->q_full, 0);
        for (int jj = 0; jj < SB && jj < n_tiles; ++jj) {
          mbar_wait(&bars->k_full[jj], 0);
          tc_fence_after();
          issue_qk(jj, jj);
        }
        int sb = 0, st = 0, st3 = SB % ST;
        uint32_t bph = 0, sph = 0, sph3 = (SB / ST) & 1;
        for (int j = 0; j < n_tiles; ++j) {
          mbar_wait(&bars->v_ready[st], sph);
          V3_TRACE(4 + t, j, 0);
          mbar_wait(&bars->p_full[t][sb], bph);  // P(t, j) is in TMEM (and S(t, j) has been read)
          tc_fence_after();
          V3_TRACE(4 + t, j, 1);
          issue_pv(sb, st, j == 0);
          V3_TRACE(4 + t, j, 2);
          if (j + SB < n_tiles) {
            mbar_wait(&bars->k_full[st3], sph3);
            tc_fence_after();
            issue_qk(sb, st3);  // overwrites S / P(t)[sb]: executes after PV(t, j) (same issuing thread)
            V3_TRACE(4 + t, j, 3);
          }
          if (++sb == SB) { sb = 0; bph ^= 1u; }
          if (++st == ST) { st = 0; sph ^= 1u; }
          if (++st3 == ST) { st3 = 0; sph3 ^= 1u; }
        }
      }
    } else if (warp < kTmaWarp) {
      // ============================== softmax warpgroups (thread == query row) ==============================
      const int wg = warp >> 2;
      const int t = wg / KS;     // query tile
      const int half = wg % KS;  // which CW-column slice of the tile's rows this warpgroup owns
      if (t < nt) {
        constexpr int CW = BN / KS;  // score columns per thread and key tile
        const int row = tid & 127;
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t s_base = tmem + lane_base + C::kColS + t * SB * BN;  // + sb * BN: score buffer; P over its start
        const uint32_t o_taddr = tmem + lane_base + C::kColO + t * C::kDP;
        const Ex2Emu ex2_emu;
        constexpr int kOChunks = C::kDP / 16;
        // O column chunks (16 fp32 columns each) this thread writes out
        const int oc_begin = (KS == 1) ? 0 : (half == 0 ? 0 : (kOChunks + 1) / 2);
        const int oc_end = (KS == 1) ? kOChunks : (half == 0 ? (kOChunks + 1) / 2 : kOChunks);
        float m_used = (kUnit && fast) ? 0.f : -INFINITY;
        float l_run = 0.f;
        const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
        int sb = 0;
        uint32_t bph = 0;
        // 32 score columns -> 16 packed bf16 P columns.  The first 16 - kEmuPairs pairs go through the MUFU, the last
        // kEmuPairs pairs through the FMA-pipe emulation (Ex2Emu, written out stage by stage over all pairs).
        // PACING: an in-order warp whose next instruction is a MUFU stalls while the MUFU queue is full, with its FMA
        // work behind it.  ptxas issues MUFUs as early as it can (long latency, no inputs to wait for), so every warp
        // of a sub-partition front-loaded its 20 MUFUs (the warps run in lockstep: same stream, same start), the MUFU
        // was saturated with the FMA pipe idle, then the FMA chains ran with the MUFU idle: time = MUFU time + FMA time
        // (0.295 + 0.242 ms = the measured 0.54 ms at N = 4096).  Each MUFU pair therefore takes a data dependency on
        // an intermediate of the emulation chains (+ 0 * finite value: one FFMA2 per pair), which spreads the MUFUs
        // evenly through the FMA work: ~6 FMA-pipe instructions (8-10 cycles) between two MUFUs (8 cycles each).
        constexpr int kEmuPairs = (kEmu == 2) ? 8 : (kEmu == 3) ? 6 : (kEmu == 4) ? 4 : (kEmu == 8) ? 2 : 0;  // of 16
        auto exp_half = [&](const float* r, const uint32_t p_taddr, const uint64_t negm2, const float emu_a,
                            const float emu_b, uint64_t& sum2a, uint64_t& sum2b, auto noshift_c) {
          constexpr bool kNoShift = decltype(noshift_c)::value;
          constexpr int NE = kEmuPairs, NM = 16 - NE, kSlots = 9;
          uint32_t u[16];
          const uint64_t zero2 = pack_f32x2(0.f, 0.f);
          int m = 0;  // next MUFU pair (compile-time after unrolling)
          auto mufu_pairs = [&](const int slot, const uint64_t dep, const bool has_dep) {
            const int upto = ((slot + 1) * NM) / kSlots;
#pragma unroll
            for (; m < upto; ++m) {
              const int e = 2 * m;
              float a0, a1;
              uint64_t x2 = pack_f32x2(r[e], r[e + 1]);
              if (!kNoShift) x2 = ffma2(x2, scale2, negm2);
              if (has_dep && NE > 0) x2 = ffma2(dep, zero2, x2);  // + 0: pacing only
              unpack_f32x2(x2, a0, a1);
              a0 = ex2(a0); a1 = ex2(a1);
              if (!C::kSumInMma) {
                if (m & 1) sum2b = fadd2(sum2b, pack_f32x2(a0, a1));
                else sum2a = fadd2(sum2a, pack_f32x2(a0, a1));
              }
              u[m] = pack_bf16(a0, a1);
            }
          };
#ifdef AGENDA_DEBUG_NO_EXP  // tools/ubench energy experiments: no exponentials at all (wrong results)
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = pack_bf16(r[2 * i] * 0.001f, r[2 * i + 1] * 0.001f);
          tmem_st16(p_taddr, u);
          return;
#endif
          uint64_t eu[NE > 0 ? NE : 1], et[NE > 0 ? NE : 1], ef[NE > 0 ? NE : 1], ep[NE > 0 ? NE : 1];
          mufu_pairs(0, zero2, false);
          if (NE > 0) {
#pragma unroll
            for (int c = 0; c < NE; ++c) {  // u = sat(s * a + b) in [0, 1]
              float u0, u1;
              asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u0) : "f"(r[2 * (NM + c)]), "f"(emu_a), "f"(emu_b));
              asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u1) : "f"(r[2 * (NM + c) + 1]), "f"(emu_a), "f"(emu_b));
              eu[c] = pack_f32x2(u0, u1);
              if (c == NE / 2 - 1) mufu_pairs(1, eu[c], true);
            }
            mufu_pairs(2, eu[NE - 1], true);
#pragma unroll
            for (int c = 0; c < NE; ++c) et[c] = ffma2(eu[c], ex2_emu.k253_2, ex2_emu.magic_lo2);  // x + 1.5 * 2^23
            mufu_pairs(3, et[NE - 1], true);
#pragma unroll
            for (int c = 0; c < NE; ++c) {
              const uint64_t g2 = fsub2(ex2_emu.magic_lo2, et[c]);  // -(n + 126)
              ef[c] = ffma2(eu[c], ex2_emu.k253_2, g2);              // x - n
              if (c == NE / 2 - 1) mufu_pairs(4, ef[c], true);
            }
            mufu_pairs(5, ef[NE - 1], true);
#pragma unroll
            for (int c = 0; c < NE; ++c) ep[c] = ffma2(ex2_emu.c3_2, ef[c], ex2_emu.c2_2);
            mufu_pairs(6, ep[NE - 1], true);
#pragma unroll
            for (int c = 0; c < NE; ++c) ep[c] = ffma2(ep[c], ef[c], ex2_emu.c1_2);
            mufu_pairs(7, ep[NE - 1], true);
#pragma unroll
            for (int c = 0; c < NE; ++c) ep[c] = ffma2(ep[c], ef[c], ex2_emu.c0_2);
            mufu_pairs(8, ep[NE - 1], true);
#pragma unroll
            for (int c = 0; c < NE; ++c) {  // 2^n: add n to the exponent field
              float t0, t1, p0, p1;
              unpack_f32x2(et[c], t0, t1);
              unpack_f32x2(ep[c], p0, p1);
              const float a0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
              const float a1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
              if (!C::kSumInMma) {
                if (c & 1) sum2b = fadd2(sum2b, pack_f32x2(a0, a1));
                else sum2a = fadd2(sum2a, pack_f32x2(a0, a1));
              }
              u[NM + c] = pack_bf16(a0, a1);
            }
          } else {
            mufu_pairs(kSlots - 1, zero2, false);
          }
          tmem_st16(p_taddr, u);
        };
        auto add_row_sum = [&](uint64_t sum2a, uint64_t sum2b) {
          if (!C::kSumInMma) {
            float s0, s1, s2, s3;
            unpack_f32x2(sum2a, s0, s1);
            unpack_f32x2(sum2b, s2, s3);
            l_run += (s0 + s1) + (s2 + s3);
          }
        };
        auto row_max32 = [&](const float* r, int kv_left) {  // max over the first kv_left (<= 32) entries of r[32]
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            mx = fmax3(mx, i < kv_left ? r[i] : -INFINITY, i + 1 < kv_left ? r[i + 1] : -INFINITY);
          return mx;
        };
        const int actor = ((warp & 3) == 0 && wg < 4) ? wg : 99;  // trace: first warp of the first four warpgroups
        (void)actor;

        // ---- exact pass (also the only pass of a kFast = false build): running maximum, lazy O rescale ----
        uint32_t next_ready = 0;
        auto exact_tile = [&](const int j, auto masked_c) {
          constexpr bool kMasked = decltype(masked_c)::value;
          const uint32_t s_taddr = s_base + sb * BN;
          const int nsb = (sb + 1 == SB) ? 0 : sb + 1;
          const uint32_t nph = (nsb == 0) ? (bph ^ 1u) : bph;
          if (!next_ready) mbar_wait(&bars->s_full[t][sb], bph);
          tc_fence_after();
          float sv[CW];
          tmem_ld32(s_taddr + half * CW, sv);
          if (CW == 64) tmem_ld32(s_taddr + 32, sv + (CW == 64 ? 32 : 0));
          next_ready = (j + 1 < n_tiles) ? mbar_test_u32(&bars->s_full[t][nsb], nph) : 0u;
          tmem_wait_ld();
          const int kv_left = kMasked ? N - j * BN - half * CW : CW;
          float mx = row_max32(sv, kv_left);
          if (CW == 64) mx = fmaxf(mx, row_max32(sv + (CW == 64 ? 32 : 0), kv_left - 32));
          if (KS == 2) {  // the other half of the row: both warpgroups of a tile must use the same running maximum
            float other[32];
            tmem_ld32(s_taddr + (half ^ 1) * 32, other);
            tmem_wait_ld();
            mx = fmaxf(mx, row_max32(other, kMasked ? N - j * BN - (half ^ 1) * 32 : 32));
          }
          if (kMasked) {
#pragma unroll
            for (int i = 0; i < CW; ++i)
              if (i >= kv_left) sv[i] = -INFINITY;
          }
          const float m_new = mx * scale_log2;
          const bool need = m_new > m_used + kV2RescaleThreshold;
          if (j == 0) {
            m_used = m_new;
          } else if (__any_sync(0xffffffffu, need)) {  // (identical in both halves: same rows, same m_new, same m_used)
            const float m_next = need ? m_new : m_used;
            const float f = ex2(m_used - m_next);
            l_run *= f;
            m_used = m_next;
            if (half == 0) {
              // O(t) must be quiescent: PV(t, j - 1) done (PV(t, j) cannot be issued before this tile's p_full)
              mbar_wait(&bars->pv_done[t][(sb + SB - 1) % SB], ((j - 1) / SB) & 1);
              tc_fence_after();
#pragma unroll
              for (int c = 0; c < kOChunks; ++c) {
                float o[16];
                tmem_ld16(o_taddr + c * 16, o);
                tmem_wait_ld();
                uint32_t u[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(o[i] * f);
                tmem_st16(o_taddr + c * 16, u);
              }
            }
          }
          const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
          const float emu_a = scale_log2 * (1.0f / 253.0f), emu_b = (126.0f - m_used) * (1.0f / 253.0f);
          uint64_t sum2a = pack_f32x2(0.f, 0.f), sum2b = sum2a;
          exp_half(sv, s_taddr + half * CW, negm2, emu_a, emu_b, sum2a, sum2b, std::false_type{});
          if (CW == 64) exp_half(sv + (CW == 64 ? 32 : 0), s_taddr + 16, negm2, emu_a, emu_b, sum2a, sum2b, std::false_type{});
          add_row_sum(sum2a, sum2b);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bars->p_full[t][sb]);
          sb = nsb; bph = nph;
        };

        // ---- fast pass, KS = 1: software-pipelined over half tiles ----
        // The 64 score columns of a tile are taken as two halves of 32 (registers ra / rb): while half 0 of tile j is
        // exponentiated the tcgen05.ld of half 1 is in flight, and while half 1 is exponentiated the barrier of tile
        // j + 1 has been probed and the load of ITS half 0 is in flight.
        float ra[32], rb[KS == 1 ? 32 : 1];
        auto fast_tile_1 = [&](const int j, auto masked_c) {  // ra holds columns 0..31 of S(t, j) (load in flight)
          constexpr bool kMasked = decltype(masked_c)::value;
          const uint32_t s_taddr = s_base + sb * BN;
          const int nsb = (sb + 1 == SB) ? 0 : sb + 1;
          const uint32_t nph = (nsb == 0) ? (bph ^ 1u) : bph;
          V3_TRACE(actor, j, 0);
          tmem_wait_ld();                    // ra has landed (requested half a tile ago)
          tmem_ld32(s_taddr + 32, rb);       // columns 32..63, in flight during the first half's exponentials
          const uint32_t ready = (j + 1 < n_tiles) ? mbar_test_u32(&bars->s_full[t][nsb], nph) : 1u;
          V3_TRACE(actor, j, 1);
          const int kv_left = kMasked ? N - j * BN : BN;
          if (!kUnit && j == 0) {  // the reference of the whole pass: the first tile's row maximum
            tmem_wait_ld();
            m_used = fmaxf(row_max32(ra, kv_left), row_max32(rb, kv_left - 32)) * scale_log2;
          }
          if (kMasked) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kv_left) ra[i] = -INFINITY;
          }
          const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
          const float emu_a = scale_log2 * (1.0f / 253.0f), emu_b = (126.0f - m_used) * (1.0f / 253.0f);
          uint64_t sum2a = pack_f32x2(0.f, 0.f), sum2b = sum2a;
          exp_half(ra, s_taddr, negm2, emu_a, emu_b, sum2a, sum2b, std::integral_constant<bool, kUnit>{});
          V3_TRACE(actor, j, 2);
          tmem_wait_ld();                    // rb has landed
          if (j + 1 < n_tiles) {
            if (!ready) mbar_wait(&bars->s_full[t][nsb], nph);
            tc_fence_after();
            tmem_ld32(s_base + nsb * BN, ra);  // columns 0..31 of S(t, j + 1)
          }
          if (kMasked) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kv_left - 32) rb[i] = -INFINITY;
          }
          exp_half(rb, s_taddr + 16, negm2, emu_a, emu_b, sum2a, sum2b, std::integral_constant<bool, kUnit>{});
          add_row_sum(sum2a, sum2b);
          V3_TRACE(actor, j, 3);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bars->p_full[t][sb]);
          V3_TRACE(actor, j, 4);
          sb = nsb; bph = nph;
        };
        // ---- fast pass, KS = 2: 32 columns per thread and tile; the next tile's barrier is probed a tile ahead ----
        auto fast_tile_2 = [&](const int j, auto masked_c) {
          constexpr bool kMasked = decltype(masked_c)::value;
          const uint32_t s_taddr = s_base + sb * BN;
          const int nsb = (sb + 1 == SB) ? 0 : sb + 1;
          const uint32_t nph = (nsb == 0) ? (bph ^ 1u) : bph;
          V3_TRACE(actor, j, 0);
          if (!next_ready) mbar_wait(&bars->s_full[t][sb], bph);
          tc_fence_after();
          V3_TRACE(actor, j, 1);
          tmem_ld32(s_taddr + half * 32, ra);
          next_ready = (j + 1 < n_tiles) ? mbar_test_u32(&bars->s_full[t][nsb], nph) : 0u;
          tmem_wait_ld();
          V3_TRACE(actor, j, 2);
          const int kv_left = kMasked ? N - j * BN - half * 32 : 32;
          if (!kUnit && j == 0) {  // the reference of the whole pass: the first tile's row maximum (both halves of the row)
            float other[32];
            tmem_ld32(s_taddr + (half ^ 1) * 32, other);
            tmem_wait_ld();
            m_used = fmaxf(row_max32(ra, kv_left), row_max32(other, kMasked ? N - j * BN - (half ^ 1) * 32 : 32)) * scale_log2;
          }
          if (kMasked) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kv_left) ra[i] = -INFINITY;
          }
          const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
          const float emu_a = scale_log2 * (1.0f / 253.0f), emu_b = (126.0f - m_used) * (1.0f / 253.0f);
          uint64_t sum2a = pack_f32x2(0.f, 0.f), sum2b = sum2a;
          exp_half(ra, s_taddr + half * 32, negm2, emu_a, emu_b, sum2a, sum2b, std::integral_constant<bool, kUnit>{});
          add_row_sum(sum2a, sum2b);
          V3_TRACE(actor, j, 3);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bars->p_full[t][sb]);
          V3_TRACE(actor, j, 4);
          sb = nsb; bph = nph;
        };
        const int n_full = N / BN;  // key tiles without padding
        if (kFast && fast) {
          if constexpr (KS == 1) {
            mbar_wait(&bars->s_full[t][0], 0);
            tc_fence_after();
            tmem_ld32(s_base, ra);
            for (int j = 0; j < n_full; ++j) fast_tile_1(j, std::false_type{});
            if (n_full < n_tiles) fast_tile_1(n_full, std::true_type{});
          } else {
            for (int j = 0; j < n_full; ++j) fast_tile_2(j, std::false_type{});
            if (n_full < n_tiles) fast_tile_2(n_full, std::true_type{});
          }
        } else {
          for (int j = 0; j < n_full; ++j) exact_tile(j, std::false_type{});
          if (n_full < n_tiles) exact_tile(n_full, std::true_type{});
        }
        // ---- epilogue: O / l -> bf16 -> global ----
        mbar_wait(&bars->pv_done[t][(n_tiles - 1) % SB], ((n_tiles - 1) / SB) & 1);
        tc_fence_after();
        if (C::kSumInMma) {
          float tmp[16];
          tmem_ld16(o_taddr + (D / 16) * 16, tmp);
          tmem_wait_ld();
          l_run = tmp[D % 16];
        } else if (KS == 2) {  // row sum = sum of the two halves' partial sums (both relative to the same reference)
          float* xsum = reinterpret_cast<float*>(sQ) + t * 256;  // Q(t) is dead: every QK^T of this pass has completed
          xsum[half * 128 + row] = l_run;
          asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
          l_run += xsum[(half ^ 1) * 128 + row];
        }
        const float inv_l = 1.0f / l_run;
        const int n = q0 + t * 128 + row;
        __nv_bfloat16* orow = out + (static_cast<long long>(b) * N + n) * (H * D) + h * D;
        if (fast && n < N && !(l_run > 0x1p-80f && l_run < 0x1p100f)) atomicOr(&bars->bad_rows, 1);
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
      if (warp == kAllocWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
      }
      break;
    }
  }  // pass
}

}  // namespace sm100

template <int D, int kEmu, bool kFast, bool kUnit, int KS = 2>
static int launch_v3(const void* q, const void* k, const void* v, void* out, int B, int H, int N, float scale,
                     cudaStream_t stream, long long ld) {
  using C = sm100::V3Cfg<D, KS>;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_head_map(&mq, q, B, H, N, D, 128, ld)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mk, k, B, H, N, D, C::kBlockN, ld)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mv, v, B, H, N, D, C::kBlockN, ld)) != AGENDA_OK) return rc;
  constexpr size_t smem = sm100::v3_smem_bytes<C>();
  auto kern = sm100::attn_self_sm100_v3_kernel<D, kEmu, kFast, kUnit, KS>;
  AGENDA_DYN_SMEM(kern, smem);
  dim3 grid(static_cast<unsigned>(((N + 128 * C::kNT - 1) / (128 * C::kNT)) * B * H));
  kern<<<grid, C::kThreads, smem, stream>>>(mq, mk, mv, static_cast<__nv_bfloat16*>(out), H, N,
                                            kUnit ? 1.0f : scale * 1.4426950408889634f);
  AGENDA_LAUNCH_CHECK("attn_self_sm100_v3_kernel");
  return AGENDA_OK;
}

// unit: q carries scale * log2(e) already (agenda_attn_self_fwd_strided with scale == 0); exact: single pass with the
// running maximum (measurements / tests)
int attn_self_sm100_v3(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int d, float scale,
                       bool unit, bool exact, void* stream, long long ld) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d == 40) {
    if (exact) return launch_v3<40, 3, false, false>(q, k, v, out, B, H, N, scale, st, ld);
#ifdef AGENDA_VARIANTS
    if (const char* e = knob("AGENDA_V3_EMU")) {
      const int emu = atoi(e);
      if (unit && emu == 13) return launch_v3<40, 3, true, true, 1>(q, k, v, out, B, H, N, scale, st, ld);  // KS = 1
      if (unit && emu == 2) return launch_v3<40, 2, true, true>(q, k, v, out, B, H, N, scale, st, ld);
      if (unit && emu == 4) return launch_v3<40, 4, true, true>(q, k, v, out, B, H, N, scale, st, ld);
      if (unit && emu == 0) return launch_v3<40, 0, true, true>(q, k, v, out, B, H, N, scale, st, ld);
    }
#endif
    return unit ? launch_v3<40, 3, true, true>(q, k, v, out, B, H, N, scale, st, ld)
                : launch_v3<40, 3, true, false>(q, k, v, out, B, H, N, scale, st, ld);
  }
  return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd (v3): head dim %d not in {40}", d);
}

#ifdef AGENDA_V2_TRACE
extern "C" int agenda_v3_trace_read(long long* host, int n) {
  const int total = sm100::kT3Actors * sm100::kT3Tiles * sm100::kT3Events;
  if (n < total) return -total;
  return cudaMemcpyFromSymbol(host, sm100::g_v3_trace, sizeof(long long) * total) == cudaSuccess ? total : -1;
}
#endif

}  // namespace agenda
