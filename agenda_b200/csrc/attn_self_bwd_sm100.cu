// Self-attention backward on the tcgen05 tensor cores (training mode, SURVEY.md §8 f N3): autograd of hook.py:104-115
// with encoder_hidden_states None, as exercised by finetune_sd_token.py:1043-1069,1089.
//
//   P = softmax(scale Q K^T), O = P V;   dV = P^T dO;   dP = dO V^T;   dS = P o (dP - Delta),  Delta_i = sum_c dO_ic O_ic;
//   dQ = scale dS K;   dK = scale dS^T Q.
//
// P is never stored by the forward kernel, so it is recomputed from the log-sum-exp of every row.  One kernel template,
// four modes, one 128-row tile of one (batch, head) per CTA and a loop over the 128-row tiles of the other operand:
//
//   LSE  rows = queries i, stream K_j:            T1 = Q_i K_j^T;  online max / sum  ->  lse2[i] = log2 sum_j 2^(c T1)
//   DQ   rows = queries i, stream K_j, V_j:       T1 = Q_i K_j^T, T2 = dO_i V_j^T;  G = E (T2 - Delta_i) scale;  dQ_i += G K_j
//   DK   rows = keys j,    stream Q_i, dO_i:      T1 = K_j Q_i^T, T2 = V_j dO_i^T;  G = E (T2 - Delta_i) scale;  dK_j += G Q_i
//   DV   rows = keys j,    stream Q_i, dO_i:      T1 = K_j Q_i^T;                    G = E;                       dV_j += G dO_i
//   DKV  (head dim <= 64: both accumulators fit TMEM) DK and DV in one pass: E is formed once, two G tiles, two accumulators
//   with E = 2^(c T1 - lse2[query]), c = scale log2(e).
//
// DK / DV work on the TRANSPOSED score tile (rows = keys), so that every MMA uses an operand form the forward kernels
// already use: T1 / T2 are SS-form MMAs of two K-major tiles (contraction over the head dim), G goes back to the
// tensor core through TMEM as packed bf16 (TS form, like P in the forward pass) and the streamed tile is read a second
// time as an MN-major B operand (contraction over its 128 rows), exactly like V in the forward PV product.
// The price is recomputation: S is formed in three kernels and dP in two (9 GEMM units instead of 5), in exchange for
// one accumulator per kernel (TMEM: T1 128 + T2 128 + G 64 + accumulator <= 160 columns), no atomics and no dQ
// round trips through global memory.
//
// Warps: 0-3 and 4-7 row warpgroups (thread == row of the tile == TMEM lane; warpgroup hf owns the 64-column half hf of
// the score tile, i.e. streamed rows hf*64 .. hf*64+63), 8 TMA producer, 9 TMEM allocator + MMA issuer.
//
// Pipeline (second version; the first one ran MMA -> 128 row threads -> MMA strictly in turn, 0.088 of the bf16 peak):
//   * the score tiles T1 / T2 are produced as two N = 64 halves with their own full / free barriers, so a warpgroup starts
//     on its half as soon as it exists and hands it back as soon as its last tcgen05.ld has landed;
//   * the issuer puts the T halves of tile it+1 on the tensor pipe BEFORE the accumulating MMAs of tile it (they only
//     need the half to be free and the next streamed tile to be resident), so the next scores are computed while the
//     row warpgroups are still exponentiating tile it; G is double buffered where TMEM has room (head dim <= 128);
//   * two row warps per SM sub-partition instead of one halve the issue-bound exponential / pack phase.
#include <algorithm>
#include <cstdlib>

#include "sm100_common.cuh"

namespace agenda {
namespace sm100 {

constexpr int kBwdThreads = 320;
constexpr int kBwdLSE = 0, kBwdDQ = 1, kBwdDK = 2, kBwdDV = 3, kBwdDKV = 4;

template <int D>
struct BCfg {
  static constexpr int kDP = (D + 15) / 16 * 16;
  static constexpr int kChunks = (D + 63) / 64;
  static constexpr int kTileBytes = kChunks * 128 * 128;   // 128 rows, 64-column swizzle chunks
  static constexpr int kStages = (kChunks <= 2) ? 2 : 1;
  static constexpr int kNG = (kDP <= 128) ? 2 : 1;          // G buffers (64 columns each)
  static constexpr int kColT1 = 0, kColT2 = 128, kColG = 256, kColAcc = 256 + 64 * kNG;
  static_assert(kColAcc + kDP <= 512, "TMEM overflow");
};

struct BBarriers {
  float vec[2][2][10][64];  // DK / DV: [warpgroup][tile parity][lse2 | Delta | kCross: Delta - heat-map term of token t][its 64
                            // streamed queries]; LSE: (m, l) exchange
  uint64_t fixed_full;
  uint64_t st_full[2], st_empty[2];
  uint64_t t_full[2], t_free[2];      // per half
  uint64_t g_full[2][2];              // [half][G buffer]
  uint64_t g_done[2];                 // [G buffer]
  uint32_t tmem_base;
};

template <int D>
constexpr size_t b_smem_bytes() {
  return 1024 + (2 + 2 * BCfg<D>::kStages) * BCfg<D>::kTileBytes + sizeof(BBarriers) + 64;
}

// Cross-attention use of the DK / DV / DKV modes (kCross; K7 on the tensor cores): the ROWS are the prompt's M <= 80 keys (one
// row tile, N = M), the streamed tiles the Nc queries; the streamed range is split over gridDim.y CTAs whose partial dK / dV are
// added into caller-zeroed fp32 buffers; the gradient of the head-mean heat maps enters dP on the rows that are selected tokens.
struct CrossBwdArgs {
  const float* d_maps;   // [B - b_first, T, Nc] fp32 or NULL
  float* out_f32;        // dK (modes DK, DKV) or dV (mode DV): fp32 [B, M, H*D], atomically accumulated
  float* out2_f32;       // dV (mode DKV)
  int Nc;                // streamed (query) length
  int b_first, T;
  int tok[8];
};

// map_r1 / map_r2: the tensors the fixed row tiles come from; map_c1 / map_c2: the streamed ones (see the mode table).
template <int D, int MODE, bool kCross = false>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_self_bwd_kernel(const __grid_constant__ CUtensorMap map_r1, const __grid_constant__ CUtensorMap map_r2,
                     const __grid_constant__ CUtensorMap map_c1, const __grid_constant__ CUtensorMap map_c2,
                     float* __restrict__ lse2, const float* __restrict__ delta, __nv_bfloat16* __restrict__ out,
                     __nv_bfloat16* __restrict__ out2, int H, int N, float scale, const CrossBwdArgs cx) {
  using C = BCfg<D>;
  static_assert(!kCross || MODE == kBwdDK || MODE == kBwdDV || MODE == kBwdDKV, "cross: key-row modes only");
  constexpr bool kDual = (MODE == kBwdDKV);   // dK and dV together: G (-> dK) and the bare E (-> dV), two accumulators
  constexpr bool kT2 = (MODE == kBwdDQ || MODE == kBwdDK || kDual);
  constexpr bool kAcc = (MODE != kBwdLSE);
  constexpr bool kColVec = (MODE == kBwdDK || MODE == kBwdDV || kDual);  // lse2 / Delta indexed by the streamed (query) tile
  constexpr int ST = C::kStages;
  constexpr int NG = kDual ? 1 : C::kNG;
  constexpr int kColG2 = C::kColG + 64;                                    // kDual: E tile (bf16) for the dV product
  constexpr int kColAcc = kDual ? C::kColG + 128 : C::kColAcc;
  constexpr int kColAcc2 = kColAcc + C::kDP;
  static_assert(!kDual || kColAcc2 + C::kDP <= 512, "DKV needs both accumulators in TMEM (head dim <= 64)");
  constexpr bool kEarlyT = (ST >= 2);   // T(it+1) ahead of acc(it) needs tile it+1 resident while tile it is still read
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sR1 = smem;
  unsigned char* sR2 = sR1 + C::kTileBytes;
  unsigned char* sC = sR2 + C::kTileBytes;   // ST stages x (C1 tile, C2 tile)
  BBarriers* bars = reinterpret_cast<BBarriers*>(sC + 2 * ST * C::kTileBytes);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_row_tiles = (N + 127) / 128;
  const int bh = blockIdx.x / n_row_tiles, rt = blockIdx.x - bh * n_row_tiles;
  const int b = bh / H, h = bh - b * H;
  const int r0 = rt * 128;
  const float c_log2 = scale * 1.4426950408889634f;
  // streamed tiles of this CTA: all of them (self-attention), or its share of the query range (kCross, gridDim.y splits)
  const int Ns = kCross ? cx.Nc : N;                      // streamed length == length of the lse2 / Delta vectors (queries)
  const int all_tiles = (Ns + 127) / 128;
  const int per_split = kCross ? (all_tiles + static_cast<int>(gridDim.y) - 1) / static_cast<int>(gridDim.y) : all_tiles;
  const int t_first = kCross ? static_cast<int>(blockIdx.y) * per_split : 0;
  const int n_tiles = min(per_split, all_tiles - t_first);
  if (n_tiles <= 0) return;                               // (uniform over the CTA, before any barrier or TMEM allocation)

  if (tid == 8 * 32) {
    tma_prefetch_desc(&map_r1); tma_prefetch_desc(&map_c1);
    if (kT2) tma_prefetch_desc(&map_r2);
    if (MODE != kBwdLSE) tma_prefetch_desc(&map_c2);
    mbar_init(&bars->fixed_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->st_full[s], 1); mbar_init(&bars->st_empty[s], 1);
      mbar_init(&bars->t_full[s], 1); mbar_init(&bars->t_free[s], 128);
      mbar_init(&bars->g_full[s][0], 128); mbar_init(&bars->g_full[s][1], 128);
      mbar_init(&bars->g_done[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // which streamed tile feeds the accumulating MMA: DQ -> K_j = C1, DK -> Q_i = C1, DV -> dO_i = C2
  constexpr bool kAccFromC2 = (MODE == kBwdDV);
  constexpr bool kLoadC2 = kT2 || kAccFromC2;

  if (warp == 8) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_expect_tx(&bars->fixed_full, (kT2 ? 2 : 1) * C::kTileBytes);
      for (int c = 0; c < C::kChunks; ++c) {
        tma_load_4d(&map_r1, &bars->fixed_full, sR1 + c * 128 * 128, c * 64, h, r0, b);
        if (kT2) tma_load_4d(&map_r2, &bars->fixed_full, sR2 + c * 128 * 128, c * 64, h, r0, b);
      }
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_tiles; ++it) {
      mbar_wait(&bars->st_empty[s], ph ^ 1);
      if (elect_one()) {
        unsigned char* c1 = sC + (2 * s) * C::kTileBytes;
        unsigned char* c2 = c1 + C::kTileBytes;
        mbar_expect_tx(&bars->st_full[s], (kLoadC2 ? 2 : 1) * C::kTileBytes);
        for (int c = 0; c < C::kChunks; ++c) {
          tma_load_4d(&map_c1, &bars->st_full[s], c1 + c * 128 * 128, c * 64, h, (t_first + it) * 128, b);
          if (kLoadC2) tma_load_4d(&map_c2, &bars->st_full[s], c2 + c * 128 * 128, c * 64, h, (t_first + it) * 128, b);
        }
      }
      __syncwarp();
      if (++s == ST) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 9) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc_t = make_idesc(128, 64, 0);
    constexpr uint32_t idesc_acc = make_idesc(128, C::kDP, 1);
    const uint64_t r1_desc = make_sdesc(smem_u32(sR1), 16, 1024);
    const uint64_t r2_desc = make_sdesc(smem_u32(sR2), 16, 1024);
    const uint64_t c_desc = make_sdesc(smem_u32(sC), 16, 1024);
    const uint64_t cacc_desc = make_sdesc(smem_u32(sC), 128 * 128, 1024);   // MN-major view: LBO = chunk stride
    // both halves of the score tile(s) of streamed tile `it`: waits for the stage and for each half to be free
    auto issue_T = [&](int it) {
      const int s = it % ST;
      mbar_wait(&bars->st_full[s], static_cast<uint32_t>(it / ST) & 1u);
      const uint32_t c1_off = (2 * s) * C::kTileBytes, c2_off = c1_off + C::kTileBytes;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (it > 0) mbar_wait(&bars->t_free[hf], static_cast<uint32_t>(it - 1) & 1u);   // the half is in registers
        tc_fence_after();
        if (elect_one()) {
          const uint32_t row_off = hf * 64 * 128;   // streamed rows hf*64.. inside every 64-column chunk
#pragma unroll
          for (int kk = 0; kk < C::kDP / 16; ++kk) {
            const uint32_t ko = (kk >> 2) * 128 * 128 + (kk & 3) * 32;
            umma_ss(tmem + C::kColT1 + hf * 64, r1_desc + static_cast<uint64_t>(ko >> 4),
                    c_desc + static_cast<uint64_t>((c1_off + ko + row_off) >> 4), idesc_t, kk != 0);
          }
          if (kT2) {
#pragma unroll
            for (int kk = 0; kk < C::kDP / 16; ++kk) {
              const uint32_t ko = (kk >> 2) * 128 * 128 + (kk & 3) * 32;
              umma_ss(tmem + C::kColT2 + hf * 64, r2_desc + static_cast<uint64_t>(ko >> 4),
                      c_desc + static_cast<uint64_t>((c2_off + ko + row_off) >> 4), idesc_t, kk != 0);
            }
          }
          umma_commit(&bars->t_full[hf]);
          if (!kAcc && hf == 1) umma_commit(&bars->st_empty[s]);
        }
        __syncwarp();
      }
    };
    mbar_wait(&bars->fixed_full, 0);
    if (kEarlyT) issue_T(0);
    for (int it = 0; it < n_tiles; ++it) {
      if (kEarlyT) { if (it + 1 < n_tiles) issue_T(it + 1); }
      else issue_T(it);
      if (kAcc) {
        const int s = it % ST, buf = it % NG;
        const uint32_t gph = static_cast<uint32_t>(it / NG) & 1u;
        const uint32_t acc_off = (2 * s + (kAccFromC2 ? 1 : 0)) * C::kTileBytes;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&bars->g_full[hf][buf], gph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {   // contraction over the half's 64 streamed rows, 16 per MMA
              const int kk = hf * 4 + k4;
              umma_ts(tmem + kColAcc, tmem + C::kColG + buf * 64 + kk * 8,
                      cacc_desc + static_cast<uint64_t>((acc_off + kk * 2048) >> 4), idesc_acc, !(it == 0 && kk == 0));
            }
            if (kDual) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {   // dV_j += E dO_i (the streamed C2 tile)
                const int kk = hf * 4 + k4;
                umma_ts(tmem + kColAcc2, tmem + kColG2 + kk * 8,
                        cacc_desc + static_cast<uint64_t>((acc_off + C::kTileBytes + kk * 2048) >> 4), idesc_acc,
                        !(it == 0 && kk == 0));
              }
            }
            if (hf == 1) { umma_commit(&bars->g_done[buf]); umma_commit(&bars->st_empty[s]); }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ============================== row warpgroups (thread == row of this CTA's tile) ==============================
    const int hf = warp >> 2;          // column half of the score tile this warpgroup owns
    const int row = tid & 127;
    const int n_row = r0 + row;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const long long vec_base = static_cast<long long>(bh) * (kColVec ? Ns : N);
    // DK / DV: this thread's element of the streamed tile's lse2 (rows 0-63 of the warpgroup) or Delta (rows 64-127) is
    // fetched one tile ahead into a register, so its global-load latency hides behind the current tile
    auto fetch_vec = [&](int it) -> float {
      const int qn = (t_first + it) * 128 + hf * 64 + (row & 63);
      if (it >= n_tiles) return 0.f;
      if (row < 64) return (qn < Ns) ? lse2[vec_base + qn] : INFINITY;   // out-of-range queries: lse2 = +inf, i.e. E = 0
      return (kT2 && qn < Ns) ? delta[vec_base + qn] : 0.f;
    };
    float vec_next = kColVec ? fetch_vec(0) : 0.f;
    float lse_r = 0.f, delta_r = 0.f;
    if (!kColVec && MODE != kBwdLSE) {
      lse_r = (n_row < N) ? lse2[vec_base + n_row] : 0.f;
      if (kT2) delta_r = (n_row < N) ? delta[vec_base + n_row] : 0.f;
    }
    float m_run = -INFINITY, l_run = 0.f;   // LSE mode (over this warpgroup's columns)
    // kCross: the gradient of the head-mean heat maps enters dP on the key rows that are selected tokens:
    //   G_jq = E_jq (T2_jq + dm_t[q] / H - Delta_q) scale = E_jq (T2_jq - Delta'_tq) scale,  Delta'_tq = Delta_q - sum_{t': tok[t'] == tok[t]} dm_t'[q] / H,
    // so a token row simply reads ITS OWN copy of the Delta vector (slot 2 + t) and the inner loop has no branch.  The
    // threads that park Delta for a column (rows 64-127 of the warpgroup) fetch the column's dm values one tile ahead.
    const bool has_maps = kCross && kT2 && cx.d_maps != nullptr && b >= cx.b_first;
    int my_slot = 1;
    if (has_maps)
      for (int t = cx.T - 1; t >= 0; --t) my_slot = (cx.tok[t] == r0 + row) ? 2 + t : my_slot;
    const float* dm_row = has_maps ? cx.d_maps + static_cast<long long>(b - cx.b_first) * cx.T * Ns : nullptr;
    const float inv_h = 1.0f / static_cast<float>(H);
    float dmn[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) dmn[t] = 0.f;
    auto fetch_dm = [&](int it) {
      if (!has_maps || row < 64 || it >= n_tiles) return;
      const int qn = (t_first + it) * 128 + hf * 64 + (row & 63);
#pragma unroll
      for (int t = 0; t < 8; ++t) dmn[t] = (t < cx.T && qn < Ns) ? dm_row[static_cast<long long>(t) * Ns + qn] : 0.f;
    };
    if (kCross) fetch_dm(0);
    for (int it = 0; it < n_tiles; ++it) {
      const int c0 = (t_first + it) * 128 + hf * 64;   // first streamed row == first score column of this warpgroup's half
      const int buf = it % NG;
      const float* vl = bars->vec[hf][it & 1][0];
      const float* vd = bars->vec[hf][it & 1][kCross ? my_slot : 1];
      if (kColVec) {
        // park the prefetched element in this tile's buffer (double buffered: the writers of tile it+1 passed this
        // barrier, so every reader of tile it-1 — same buffer — had finished), read back as broadcasts
        bars->vec[hf][it & 1][row >> 6][row & 63] = vec_next;
        if (kCross && has_maps && row >= 64) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            if (t < cx.T) {
              float sum = 0.f;
#pragma unroll
              for (int t2 = 0; t2 < 8; ++t2) sum += (t2 < cx.T && cx.tok[t2] == cx.tok[t]) ? dmn[t2] : 0.f;
              bars->vec[hf][it & 1][2 + t][row & 63] = vec_next - sum * inv_h;
            }
          }
        }
        vec_next = fetch_vec(it + 1);
        if (kCross) fetch_dm(it + 1);
        if (hf == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      mbar_wait(&bars->t_full[hf], static_cast<uint32_t>(it) & 1u);
      tc_fence_after();
      if (MODE == kBwdLSE) {
        // pass A over the half's 64 columns: the tile's row maximum (columns beyond N masked)
        float tile_max = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float t1[32];
          tmem_ld32(tmem + lane_base + C::kColT1 + hf * 64 + c * 32, t1);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) tile_max = fmaxf(tile_max, (c0 + c * 32 + i < N) ? t1[i] : -INFINITY);
        }
        const float m_new = fmaxf(m_run, tile_max * c_log2);
        l_run = (m_new == -INFINITY) ? 0.f : l_run * ex2(m_run - m_new);   // (first tile: 2^-inf = 0)
        m_run = m_new;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float t1[32], t2[32];
        tmem_ld32(tmem + lane_base + C::kColT1 + hf * 64 + c * 32, t1);
        if (kT2) tmem_ld32(tmem + lane_base + C::kColT2 + hf * 64 + c * 32, t2);
        tmem_wait_ld();
        if (c == 1) {   // this half of T1 / T2 is in registers: the next tile's scores may overwrite it
          tc_fence_before();
          mbar_arrive(&bars->t_free[hf]);
        }
        if (MODE == kBwdLSE) {
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += (c0 + c * 32 + i < N) ? ex2(fmaf(t1[i], c_log2, -m_run)) : 0.f;
          l_run += sum;
        } else {
          uint32_t u[16], u2[kDual ? 16 : 1];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float g[2], pe[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = c * 32 + i + e;   // within the half
              float lse_x, delta_x;
              if (kColVec) { lse_x = vl[col]; delta_x = kT2 ? vd[col] : 0.f; }
              else { lse_x = lse_r; delta_x = delta_r; }
              float p = ex2(fmaf(t1[i + e], c_log2, -lse_x));
              if (!kColVec && c0 + col >= N) p = 0.f;   // keys beyond the sequence (DQ); DK / DV: lse2 = +inf did it
              g[e] = kT2 ? p * (t2[i + e] - delta_x) * scale : p;
              pe[e] = p;
            }
            u[i >> 1] = pack_bf16(g[0], g[1]);
            if (kDual) u2[i >> 1] = pack_bf16(pe[0], pe[1]);
          }
          if (c == 0 && it >= NG) {   // the accumulating MMAs that read this G buffer last (tile it - NG) are done
            mbar_wait(&bars->g_done[buf], static_cast<uint32_t>(it / NG - 1) & 1u);
            tc_fence_after();
          }
          tmem_st16(tmem + lane_base + C::kColG + buf * 64 + hf * 32 + c * 16, u);
          if (kDual) tmem_st16(tmem + lane_base + kColG2 + hf * 32 + c * 16, u2);
        }
      }
      if (kAcc) {
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->g_full[hf][buf]);
      }
    }
    // ---- epilogue ----
    if (MODE == kBwdLSE) {
      // combine the two halves' (m, l): warpgroup 1 parks its pair in shared memory (the vec area is idle in this mode)
      float* ex_m = &bars->vec[0][0][0][0];
      float* ex_l = ex_m + 128;
      if (hf == 1) { ex_m[row] = m_run; ex_l[row] = l_run; }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (hf == 0 && n_row < N) {
        const float m1 = ex_m[row], l1 = ex_l[row];
        const float m = fmaxf(m_run, m1);
        const float l = ((m_run == -INFINITY) ? 0.f : l_run * ex2(m_run - m)) + ((m1 == -INFINITY) ? 0.f : l1 * ex2(m1 - m));
        lse2[vec_base + n_row] = m + log2f(l);
      }
    } else {
      const int last = n_tiles - 1;
      mbar_wait(&bars->g_done[last % NG], static_cast<uint32_t>(last / NG) & 1u);
      tc_fence_after();
#pragma unroll
      for (int which = 0; which < (kDual ? 2 : 1); ++which) {
        __nv_bfloat16* orow = (which ? out2 : out) + (static_cast<long long>(b) * N + n_row) * (H * D) + h * D;
        float* frow = kCross ? (which ? cx.out2_f32 : cx.out_f32) + (static_cast<long long>(b) * N + n_row) * (H * D) + h * D : nullptr;
        const uint32_t acol = which ? kColAcc2 : kColAcc;
#pragma unroll
        for (int c = 0; c < C::kDP / 16; ++c) {
          if ((c & 1) != hf) continue;   // the two warpgroups take alternate 16-column chunks of the accumulator
          float o[16];
          tmem_ld16(tmem + lane_base + acol + c * 16, o);
          tmem_wait_ld();
          if (kCross) {   // partial sums of this CTA's query range: result-less fp32 adds into the caller-zeroed buffers
            if (n_row < N) {
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (c * 16 + k < D) atomicAdd(frow + c * 16 + k, o[k]);
            }
          } else if (n_row < N) {
            uint4 lo, hi;
            lo.x = pack_bf16(o[0], o[1]); lo.y = pack_bf16(o[2], o[3]); lo.z = pack_bf16(o[4], o[5]); lo.w = pack_bf16(o[6], o[7]);
            hi.x = pack_bf16(o[8], o[9]); hi.y = pack_bf16(o[10], o[11]); hi.z = pack_bf16(o[12], o[13]); hi.w = pack_bf16(o[14], o[15]);
            if (c * 16 + 8 <= D) *reinterpret_cast<uint4*>(orow + c * 16) = lo;
            if (c * 16 + 16 <= D) *reinterpret_cast<uint4*>(orow + c * 16 + 8) = hi;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- K7 on the tensor cores, part 1: dQ of the cross-attention (rows = queries; the prompt's M <= 80 keys are ONE tile) ----
// CTA = one 128-query tile of one (batch, head).  S = Q K^T and dP = dO V^T (N = 80) land in TMEM; a thread owns a query
// row: softmax over its 80 scores in registers, the gradient of the head-mean heat maps added to dP on the selected token
// columns (1 / H each), Delta_i = sum_j P_ij dP_ij, dS = P (dP - Delta) scale back to TMEM as packed bf16, dQ = dS K as a
// TS-form MMA over the 80 keys.  lse2 / Delta of every query row are left in the workspace for the dK / dV kernel.
constexpr int kXqThreads = 192;
struct XqBarriers {
  uint64_t ld_full, t_full, g_full, acc_full;
  uint32_t tmem_base;
};
template <int D>
constexpr size_t xq_smem_bytes() { return 1024 + 4 * BCfg<D>::kTileBytes + sizeof(XqBarriers) + 64; }

template <int D>
__global__ void __launch_bounds__(kXqThreads)
attn_cross_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                         const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                         float* __restrict__ lse2, float* __restrict__ delta, __nv_bfloat16* __restrict__ dq, int H, int N,
                         int M, float scale, const CrossBwdArgs cx) {
  using C = BCfg<D>;
  constexpr int kColS = 0, kColP = 80, kColG = 160, kColAcc = 208;   // S 80 | dP 80 | dS bf16 48 | dQ accumulator
  constexpr int kCols = (kColAcc + C::kDP <= 256) ? 256 : 512;
  static_assert(kColAcc + C::kDP <= 512, "TMEM overflow");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sdO = sQ + C::kTileBytes;
  unsigned char* sK = sdO + C::kTileBytes;
  unsigned char* sV = sK + C::kTileBytes;
  XqBarriers* bars = reinterpret_cast<XqBarriers*>(sV + C::kTileBytes);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (N + 127) / 128;
  const int bh = blockIdx.x / n_tiles, rt = blockIdx.x - bh * n_tiles;
  const int b = bh / H, h = bh - b * H;
  const int r0 = rt * 128;
  const float c_log2 = scale * 1.4426950408889634f;

  if (tid == 4 * 32) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    mbar_init(&bars->ld_full, 1); mbar_init(&bars->t_full, 1); mbar_init(&bars->g_full, 128); mbar_init(&bars->acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(&bars->tmem_base, kCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(&bars->ld_full, 4 * C::kTileBytes);
      for (int c = 0; c < C::kChunks; ++c) {
        tma_load_4d(&map_q, &bars->ld_full, sQ + c * 128 * 128, c * 64, h, r0, b);
        tma_load_4d(&map_do, &bars->ld_full, sdO + c * 128 * 128, c * 64, h, r0, b);
        tma_load_4d(&map_k, &bars->ld_full, sK + c * 128 * 128, c * 64, h, 0, b);
        tma_load_4d(&map_v, &bars->ld_full, sV + c * 128 * 128, c * 64, h, 0, b);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    constexpr uint32_t idesc_t = make_idesc(128, 80, 0);
    constexpr uint32_t idesc_acc = make_idesc(128, C::kDP, 1);
    const uint64_t q_desc = make_sdesc(smem_u32(sQ), 16, 1024), do_desc = make_sdesc(smem_u32(sdO), 16, 1024);
    const uint64_t k_desc = make_sdesc(smem_u32(sK), 16, 1024), v_desc = make_sdesc(smem_u32(sV), 16, 1024);
    const uint64_t kacc_desc = make_sdesc(smem_u32(sK), 128 * 128, 1024);   // MN-major view of the key tile
    mbar_wait(&bars->ld_full, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < C::kDP / 16; ++kk) {
        const uint32_t ko = (kk >> 2) * 128 * 128 + (kk & 3) * 32;
        umma_ss(tmem + kColS, q_desc + static_cast<uint64_t>(ko >> 4), k_desc + static_cast<uint64_t>(ko >> 4), idesc_t, kk != 0);
      }
#pragma unroll
      for (int kk = 0; kk < C::kDP / 16; ++kk) {
        const uint32_t ko = (kk >> 2) * 128 * 128 + (kk & 3) * 32;
        umma_ss(tmem + kColP, do_desc + static_cast<uint64_t>(ko >> 4), v_desc + static_cast<uint64_t>(ko >> 4), idesc_t, kk != 0);
      }
      umma_commit(&bars->t_full);
    }
    __syncwarp();
    mbar_wait(&bars->g_full, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 5; ++kk)   // contraction over the 80 (padded) keys, 16 per MMA
        umma_ts(tmem + kColAcc, tmem + kColG + kk * 8, kacc_desc + static_cast<uint64_t>((kk * 2048) >> 4), idesc_acc, kk != 0);
      umma_commit(&bars->acc_full);
    }
    __syncwarp();
  } else {
    const int row = tid, n_row = r0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    mbar_wait(&bars->t_full, 0);
    tc_fence_after();
    // the gradient of the head-mean heat maps enters dP on the selected token columns (1 / H per head): patched into the
    // TMEM tile one column at a time (the column address is a run-time value; a repeated token is simply added twice)
    if (cx.d_maps != nullptr && b >= cx.b_first) {
      for (int t = 0; t < cx.T; ++t) {
        const float dm = (n_row < N) ? cx.d_maps[(static_cast<long long>(b - cx.b_first) * cx.T + t) * N + n_row] : 0.f;
        const uint32_t addr = tmem + lane_base + kColP + static_cast<uint32_t>(cx.tok[t]);
        const float cur = tmem_ld1(addr);
        tmem_wait_ld();
        tmem_st1(addr, cur + dm / static_cast<float>(H));
        tmem_wait_st();
      }
    }
    float p[80];
#pragma unroll
    for (int c = 0; c < 5; ++c) tmem_ld16(tmem + lane_base + kColS + c * 16, p + c * 16);
    tmem_wait_ld();
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 80; ++j) m = fmaxf(m, (j < M) ? p[j] : -INFINITY);
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 80; ++j) {
      p[j] = (j < M) ? ex2((p[j] - m) * c_log2) : 0.f;
      l += p[j];
    }
    const float inv_l = 1.0f / l;
#pragma unroll
    for (int j = 0; j < 80; ++j) p[j] *= inv_l;
    float dlt = 0.f;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      float raw[16];
      tmem_ld16(tmem + lane_base + kColP + c * 16, raw);
      tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < 16; ++k) dlt = fmaf(p[c * 16 + k], raw[k], dlt);
    }
    if (n_row < N) {
      lse2[static_cast<long long>(bh) * N + n_row] = fmaf(m, c_log2, log2f(l));
      delta[static_cast<long long>(bh) * N + n_row] = dlt;
    }
#pragma unroll
    for (int c2 = 0; c2 < 3; ++c2) {
      uint32_t u[16];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c = c2 * 2 + hh;
        if (c < 5) {
          float raw[16];
          tmem_ld16(tmem + lane_base + kColP + c * 16, raw);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; k += 2) {
            const float g0 = p[c * 16 + k] * (raw[k] - dlt) * scale;
            const float g1 = p[c * 16 + k + 1] * (raw[k + 1] - dlt) * scale;
            u[hh * 8 + (k >> 1)] = pack_bf16(g0, g1);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) u[hh * 8 + k] = 0u;
        }
      }
      tmem_st16(tmem + lane_base + kColG + c2 * 16, u);
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(&bars->g_full);
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after();
    __nv_bfloat16* orow = dq + (static_cast<long long>(b) * N + n_row) * (H * D) + h * D;
#pragma unroll
    for (int c = 0; c < C::kDP / 16; ++c) {
      float o[16];
      tmem_ld16(tmem + lane_base + kColAcc + c * 16, o);
      tmem_wait_ld();
      if (n_row < N) {
        uint4 lo, hi;
        lo.x = pack_bf16(o[0], o[1]); lo.y = pack_bf16(o[2], o[3]); lo.z = pack_bf16(o[4], o[5]); lo.w = pack_bf16(o[6], o[7]);
        hi.x = pack_bf16(o[8], o[9]); hi.y = pack_bf16(o[10], o[11]); hi.z = pack_bf16(o[12], o[13]); hi.w = pack_bf16(o[14], o[15]);
        if (c * 16 + 8 <= D) *reinterpret_cast<uint4*>(orow + c * 16) = lo;
        if (c * 16 + 16 <= D) *reinterpret_cast<uint4*>(orow + c * 16 + 8) = hi;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, kCols);
  }
}

// Delta[b, h, n] = sum_c dO[b, n, h*d + c] * O[b, n, h*d + c]  (fp32): one warp per (b, n), lanes over the row
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                      float* __restrict__ delta, int B, int H, int N, int d) {
  const int lane = threadIdx.x & 31;
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= static_cast<long long>(B) * N) return;
  const int b = static_cast<int>(wid / N), n = static_cast<int>(wid - static_cast<long long>(b) * N);
  const __nv_bfloat16* po = o + wid * (H * d);
  const __nv_bfloat16* pd = d_o + wid * (H * d);
  for (int h = 0; h < H; ++h) {
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc += __bfloat162float(po[h * d + c]) * __bfloat162float(pd[h * d + c]);
    acc = warp_sum(acc);
    if (lane == 0) delta[(static_cast<long long>(b) * H + h) * N + n] = acc;
  }
}

}  // namespace sm100

template <int D, int MODE>
static int launch_bwd_mode(const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& c1, const CUtensorMap& c2,
                           float* lse2, const float* delta, void* out, void* out2, int B, int H, int N, float scale,
                           cudaStream_t st) {
  constexpr size_t smem = sm100::b_smem_bytes<D>();
  auto kern = sm100::attn_self_bwd_kernel<D, MODE>;
  AGENDA_DYN_SMEM(kern, smem);
  const int n_tiles = (N + 127) / 128;
  kern<<<static_cast<unsigned>(n_tiles) * B * H, sm100::kBwdThreads, smem, st>>>(r1, r2, c1, c2, lse2, delta,
                                                                               static_cast<__nv_bfloat16*>(out),
                                                                               static_cast<__nv_bfloat16*>(out2), H, N, scale, sm100::CrossBwdArgs{});
  AGENDA_LAUNCH_CHECK("attn_self_bwd_kernel");
  return AGENDA_OK;
}

template <int D>
static int attn_self_bwd_d(const void* q, const void* k, const void* v, const void* o, const void* d_o, float* ws, void* dq,
                           void* dk, void* dv, int B, int H, int N, float scale, cudaStream_t st, const float* lse_in) {
  CUtensorMap mq, mk, mv, mdo;
  int rc;
  if ((rc = make_head_map(&mq, q, B, H, N, D, 128)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mk, k, B, H, N, D, 128)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mv, v, B, H, N, D, 128)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mdo, d_o, B, H, N, D, 128)) != AGENDA_OK) return rc;
  float* lse2 = lse_in ? const_cast<float*>(lse_in) : ws;   // (the forward kernel's log-sum-exp, or this call's LSE pass)
  float* delta = ws + static_cast<long long>(B) * H * N;
  {
    const long long warps = static_cast<long long>(B) * N;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((warps * 32 + threads - 1) / threads);
    sm100::attn_bwd_delta_kernel<<<blocks, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(o),
                                                             static_cast<const __nv_bfloat16*>(d_o), delta, B, H, N, D);
    AGENDA_LAUNCH_CHECK("attn_bwd_delta_kernel");
  }
  if (!lse_in && (rc = launch_bwd_mode<D, sm100::kBwdLSE>(mq, mq, mk, mk, lse2, delta, nullptr, nullptr, B, H, N, scale, st)) != AGENDA_OK) return rc;
  if ((rc = launch_bwd_mode<D, sm100::kBwdDQ>(mq, mdo, mk, mv, lse2, delta, dq, nullptr, B, H, N, scale, st)) != AGENDA_OK) return rc;
  if constexpr (D <= 64) {   // both accumulators fit TMEM: dK and dV share the recomputed E
    return launch_bwd_mode<D, sm100::kBwdDKV>(mk, mv, mq, mdo, lse2, delta, dk, dv, B, H, N, scale, st);
  } else {
    if ((rc = launch_bwd_mode<D, sm100::kBwdDK>(mk, mv, mq, mdo, lse2, delta, dk, nullptr, B, H, N, scale, st)) != AGENDA_OK) return rc;
    return launch_bwd_mode<D, sm100::kBwdDV>(mk, mk, mq, mdo, lse2, delta, dv, nullptr, B, H, N, scale, st);
  }
}

template <int D>
static int attn_cross_bwd_tc_d(const void* q, const void* k, const void* v, const void* d_o, const float* d_maps, float* ws,
                               void* dq, float* dk, float* dv, int B, int H, int N, int M, float scale, const int32_t* token_idx,
                               int T, int b_first, cudaStream_t st) {
  CUtensorMap mq, mk, mv, mdo;
  int rc;
  if ((rc = make_head_map(&mq, q, B, H, N, D, 128)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mdo, d_o, B, H, N, D, 128)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mk, k, B, H, M, D, 128)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mv, v, B, H, M, D, 128)) != AGENDA_OK) return rc;
  float* lse2 = ws;
  float* delta = ws + static_cast<long long>(B) * H * N;
  sm100::CrossBwdArgs cx{};
  cx.d_maps = d_maps; cx.Nc = N; cx.b_first = b_first; cx.T = d_maps ? T : 0;
  for (int t = 0; t < 8; ++t) cx.tok[t] = (d_maps && t < T) ? token_idx[t] : -1;
  const int n_tiles = (N + 127) / 128;
  {
    constexpr size_t smem = sm100::xq_smem_bytes<D>();
    auto kern = sm100::attn_cross_bwd_dq_kernel<D>;
    AGENDA_DYN_SMEM(kern, smem);
    kern<<<static_cast<unsigned>(n_tiles) * B * H, sm100::kXqThreads, smem, st>>>(mq, mdo, mk, mv, lse2, delta,
                                                                                 static_cast<__nv_bfloat16*>(dq), H, N, M, scale, cx);
    AGENDA_LAUNCH_CHECK("attn_cross_bwd_dq_kernel");
  }
  // dK / dV: rows = keys; the query range is split so that the launch fills the GPU (partials meet in fp32 atomics)
  const int splits = std::max(1, std::min(n_tiles, (num_sms() + B * H - 1) / (B * H)));
  const dim3 grid(static_cast<unsigned>(B) * H, static_cast<unsigned>(splits));
  constexpr size_t smem_b = sm100::b_smem_bytes<D>();
  if constexpr (D <= 64) {
    auto kern = sm100::attn_self_bwd_kernel<D, sm100::kBwdDKV, true>;
    AGENDA_DYN_SMEM(kern, smem_b);
    cx.out_f32 = dk; cx.out2_f32 = dv;
    kern<<<grid, sm100::kBwdThreads, smem_b, st>>>(mk, mv, mq, mdo, lse2, delta, nullptr, nullptr, H, M, scale, cx);
    AGENDA_LAUNCH_CHECK("attn_self_bwd_kernel<DKV, cross>");
  } else {
    auto kern_k = sm100::attn_self_bwd_kernel<D, sm100::kBwdDK, true>;
    AGENDA_DYN_SMEM(kern_k, smem_b);
    cx.out_f32 = dk; cx.out2_f32 = nullptr;
    kern_k<<<grid, sm100::kBwdThreads, smem_b, st>>>(mk, mv, mq, mdo, lse2, delta, nullptr, nullptr, H, M, scale, cx);
    AGENDA_LAUNCH_CHECK("attn_self_bwd_kernel<DK, cross>");
    auto kern_v = sm100::attn_self_bwd_kernel<D, sm100::kBwdDV, true>;
    AGENDA_DYN_SMEM(kern_v, smem_b);
    cx.out_f32 = dv;
    kern_v<<<grid, sm100::kBwdThreads, smem_b, st>>>(mk, mk, mq, mdo, lse2, delta, nullptr, nullptr, H, M, scale, cx);
    AGENDA_LAUNCH_CHECK("attn_self_bwd_kernel<DV, cross>");
  }
  return AGENDA_OK;
}

}  // namespace agenda

using namespace agenda;

extern "C" long long agenda_attn_cross_bwd_tc_workspace_bytes(int B, int H, int N) {
  if (B <= 0 || H <= 0 || N <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "attn_cross_bwd_tc_workspace_bytes: B=%d H=%d N=%d", B, H, N);
  return 2ll * B * H * N * 4;
}

extern "C" int agenda_attn_cross_bwd_tc(const void* q, const void* k, const void* v, const void* d_out, const float* d_maps,
                                        void* workspace, void* dq, float* dk, float* dv, int B, int H, int N, int M, int d,
                                        float scale, const int32_t* token_idx, int T, int b_first, void* stream) {
  const char* who = "attn_cross_bwd_tc";
  if (!q || !k || !v || !d_out || !workspace || !dq || !dk || !dv) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (B <= 0 || H <= 0 || N <= 0 || M <= 0 || M > 80 || d <= 0 || static_cast<long long>(B) * H > 65535 || b_first < 0 ||
      b_first > B)
    return fail(AGENDA_ERR_BAD_SHAPE, "%s: B=%d H=%d N=%d M=%d d=%d b_first=%d (M <= 80)", who, B, H, N, M, d, b_first);
  if (d_maps) {
    if (!token_idx || T < 1 || T > 8) return fail(AGENDA_ERR_UNSUPPORTED, "%s: 1..8 selected tokens (T=%d)", who, T);
    for (int t = 0; t < T; ++t)
      if (token_idx[t] < 0 || token_idx[t] >= M) return fail(AGENDA_ERR_BAD_SHAPE, "%s: token index %d outside [0, %d)", who, token_idx[t], M);
  }
  const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(workspace);
  if (al & 15) return fail(AGENDA_ERR_MISALIGNED, "%s: q, k, v, d_out, dq, workspace must be 16-byte aligned", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  switch (d) {
    case 40: return attn_cross_bwd_tc_d<40>(q, k, v, d_out, d_maps, ws, dq, dk, dv, B, H, N, M, scale, token_idx, T, b_first, st);
    case 64: return attn_cross_bwd_tc_d<64>(q, k, v, d_out, d_maps, ws, dq, dk, dv, B, H, N, M, scale, token_idx, T, b_first, st);
    case 80: return attn_cross_bwd_tc_d<80>(q, k, v, d_out, d_maps, ws, dq, dk, dv, B, H, N, M, scale, token_idx, T, b_first, st);
    case 160: return attn_cross_bwd_tc_d<160>(q, k, v, d_out, d_maps, ws, dq, dk, dv, B, H, N, M, scale, token_idx, T, b_first, st);
    default: return fail(AGENDA_ERR_UNSUPPORTED, "%s: head dim %d not in {40,64,80,160}", who, d);
  }
}

extern "C" long long agenda_attn_self_bwd_workspace_bytes(int B, int H, int N) {
  if (B <= 0 || H <= 0 || N <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "attn_self_bwd_workspace_bytes: B=%d H=%d N=%d", B, H, N);
  return 2ll * B * H * N * 4;
}

static int attn_self_bwd_impl(const char* who, const void* q, const void* k, const void* v, const void* out, const void* d_out,
                              const float* lse, void* workspace, void* dq, void* dk, void* dv, int dtype, int B, int H, int N,
                              int d, float scale, void* stream) {
  if (!q || !k || !v || !out || !d_out || !workspace || !dq || !dk || !dv)
    return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (dtype != AGENDA_BF16) return fail(AGENDA_ERR_UNSUPPORTED, "%s: bf16 only (dtype=1)", who);
  if (B <= 0 || H <= 0 || N <= 0 || d <= 0 || static_cast<long long>(B) * H > 65535)
    return fail(AGENDA_ERR_BAD_SHAPE, "%s: B=%d H=%d N=%d d=%d", who, B, H, N, d);
  const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(dq) |
                       reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(workspace);
  if (al & 15) return fail(AGENDA_ERR_MISALIGNED, "%s: all pointers must be 16-byte aligned", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  switch (d) {
    case 40: return attn_self_bwd_d<40>(q, k, v, out, d_out, ws, dq, dk, dv, B, H, N, scale, st, lse);
    case 64: return attn_self_bwd_d<64>(q, k, v, out, d_out, ws, dq, dk, dv, B, H, N, scale, st, lse);
    case 80: return attn_self_bwd_d<80>(q, k, v, out, d_out, ws, dq, dk, dv, B, H, N, scale, st, lse);
    case 160: return attn_self_bwd_d<160>(q, k, v, out, d_out, ws, dq, dk, dv, B, H, N, scale, st, lse);
    default: return fail(AGENDA_ERR_UNSUPPORTED, "%s: head dim %d not in {40,64,80,160}", who, d);
  }
}

extern "C" int agenda_attn_self_bwd(const void* q, const void* k, const void* v, const void* out, const void* d_out,
                                    void* workspace, void* dq, void* dk, void* dv, int dtype, int B, int H, int N, int d,
                                    float scale, void* stream) {
  return attn_self_bwd_impl("attn_self_bwd", q, k, v, out, d_out, nullptr, workspace, dq, dk, dv, dtype, B, H, N, d, scale, stream);
}

extern "C" int agenda_attn_self_bwd_lse(const void* q, const void* k, const void* v, const void* out, const void* d_out,
                                        const float* lse, void* workspace, void* dq, void* dk, void* dv, int dtype, int B, int H,
                                        int N, int d, float scale, void* stream) {
  if (!lse || (reinterpret_cast<uintptr_t>(lse) & 3)) return fail(AGENDA_ERR_NULL_POINTER, "attn_self_bwd_lse: lse is null or misaligned");
  return attn_self_bwd_impl("attn_self_bwd_lse", q, k, v, out, d_out, lse, workspace, dq, dk, dv, dtype, B, H, N, d, scale, stream);
}
