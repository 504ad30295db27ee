// K1: fused self-attention  O = softmax(scale * Q K^T) V  for sm_100a — tcgen05 tensor cores, TMEM accumulators,
// TMA-fed shared-memory tiles, warp-specialised, flash-style online softmax (the [B*H,N,N] probability tensor that
// the reference materialises at data_generation/hook.py:108 never exists).
//
// Layout: q/k/v/out are [B, N, H*d] bf16, token-major — exactly what to_q/to_k/to_v produce (hook.py:93,101-102);
// head_to_batch_dim/batch_to_head_dim (hook.py:104-106,115) are folded into a 4-D TMA tensor map
// (d, H, N, B) whose box is (64 elems, 1 head, BLOCK rows, 1 batch).  Head dims 40/80/160 are zero-padded to the
// 64-element (128-byte) swizzle atom by TMA out-of-bounds fill, so no padded copy of Q/K/V is ever made; the MMAs
// only run over round16(d) (48/80/160) of K-extent / N-extent.
//
// CTA = one (batch, head, 128-query tile).  6 warps:
//   warps 0-3  softmax warpgroup: thread r owns query row r (tcgen05.ld 32x32b: warp w <-> TMEM lanes 32w..32w+31)
//   warp 4     TMA producer (Q once; K_j, V_j through 2-stage mbarrier rings)
//   warp 5     TMEM allocator + single-thread tcgen05.mma issuer
// TMEM (512 columns): S[0] | S[1] (BLOCK_N fp32 columns each, double buffered so QK_{j+1} overlaps softmax_j) | O.
// P (bf16) goes back to the tensor core either through TMEM (aliasing the S buffer it came from; TS-form MMA) or
// through a 128B-swizzled shared-memory tile (SS-form) — template switch kPTmem.
// MMA issue order: QK_0, QK_1, PV_0, QK_2, PV_1, ...
// O is rescaled lazily (only when a row max grows by more than 2^8), by the softmax warpgroup itself.
#include "sm100_common.cuh"

namespace agenda {

namespace sm100 {

constexpr int kBlockM = 128;
constexpr int kThreads = 192;
constexpr int kSoftmaxThreads = 128;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

template <int D>
struct Cfg {
  static constexpr int kD = D;
  static constexpr int kDP = (D + 15) / 16 * 16;        // MMA extent over the head dim
  static constexpr int kChunks = (D + 63) / 64;         // 64-element (128 B) swizzle atoms per row
  static constexpr int kBlockN = (D <= 80) ? 128 : 64;  // keys per KV tile
  static constexpr int kPChunks = kBlockN / 64;
  static constexpr int kQBytes = kChunks * kBlockM * 128;
  static constexpr int kKVBytes = kChunks * kBlockN * 128;  // one K (or V) stage
  static constexpr int kPBytes = kPChunks * kBlockM * 128;  // one P buffer (SS variant)
  static constexpr int kColS0 = 0, kColS1 = kBlockN, kColO = 2 * kBlockN;
  static_assert(kColO + kDP <= 512, "TMEM overflow");
};

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t s_full[2], p_full[2], pv_done[2];
  uint32_t tmem_base;
};

template <int D, bool kPTmem>
constexpr size_t smem_bytes() {
  using C = Cfg<D>;
  return 1024 /*align slack*/ + C::kQBytes + 4 * C::kKVBytes + (kPTmem ? 0 : 2 * C::kPBytes) + sizeof(Barriers) + 64;
}

template <int D, bool kPTmem>
__global__ void __launch_bounds__(kThreads, 1)
attn_self_sm100_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv_k,
                       const __grid_constant__ CUtensorMap map_kv_v, __nv_bfloat16* __restrict__ out, int H, int N,
                       float scale_log2) {
  using C = Cfg<D>;
  constexpr int BN = C::kBlockN;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + C::kQBytes;          // 2 stages
  unsigned char* sV = sK + 2 * C::kKVBytes;     // 2 stages
  unsigned char* sP = sV + 2 * C::kKVBytes;     // 2 buffers (SS variant only)
  Barriers* bars = reinterpret_cast<Barriers*>(sP + (kPTmem ? 0 : 2 * C::kPBytes));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBlockM;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int n_tiles = (N + BN - 1) / BN;

  if (tid == 4 * 32) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_kv_k); tma_prefetch_desc(&map_kv_v);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1);
      mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1);
      mbar_init(&bars->s_full[s], 1); mbar_init(&bars->p_full[s], kSoftmaxThreads);
      mbar_init(&bars->pv_done[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_expect_tx(&bars->q_full, C::kQBytes);
      for (int c = 0; c < C::kChunks; ++c) tma_load_4d(&map_q, &bars->q_full, sQ + c * kBlockM * 128, c * 64, h, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&bars->k_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->k_full[s], C::kKVBytes);
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(&map_kv_k, &bars->k_full[s], sK + s * C::kKVBytes + c * BN * 128, c * 64, h, j * BN, b);
        mbar_wait(&bars->v_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->v_full[s], C::kKVBytes);
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(&map_kv_v, &bars->v_full[s], sV + s * C::kKVBytes + c * BN * 128, c * 64, h, j * BN, b);
      }
    }
  } else if (warp == 5) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(kBlockM, BN, 0);
      constexpr uint32_t idesc_pv = make_idesc(kBlockM, C::kDP, 1);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      auto issue_pv = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&bars->p_full[s], ph);
        mbar_wait(&bars->v_full[s], ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          // B = V tile, MN-major: 16 keys per MMA = 2 swizzle atoms of 8 rows (SBO 1024 B); d > 64 continues in the
          // next 64-column chunk (LBO = BN*128 B)
          const uint64_t bdesc = make_sdesc(v_addr + s * C::kKVBytes + kk * 2048, BN * 128, 1024);
          if (kPTmem) {
            umma_ts(tmem + C::kColO, tmem + (s ? C::kColS1 : C::kColS0) + kk * 8, bdesc, idesc_pv, (j | kk) != 0);
          } else {
            const uint64_t adesc = make_sdesc(p_addr + s * C::kPBytes + (kk >> 2) * kBlockM * 128 + (kk & 3) * 32, 16, 1024);
            umma_ss(tmem + C::kColO, adesc, bdesc, idesc_pv, (j | kk) != 0);
          }
        }
        umma_commit(&bars->v_empty[s]);
        umma_commit(&bars->pv_done[s]);
      };
      mbar_wait(&bars->q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&bars->k_full[s], ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < C::kDP / 16; ++kk) {
          const uint64_t adesc = make_sdesc(q_addr + (kk >> 2) * kBlockM * 128 + (kk & 3) * 32, 16, 1024);
          const uint64_t bdesc = make_sdesc(k_addr + s * C::kKVBytes + (kk >> 2) * BN * 128 + (kk & 3) * 32, 16, 1024);
          umma_ss(tmem + (s ? C::kColS1 : C::kColS0), adesc, bdesc, idesc_qk, kk != 0);
        }
        umma_commit(&bars->k_empty[s]);
        umma_commit(&bars->s_full[s]);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(n_tiles - 1);
    }
  } else {
    // ============================== softmax warpgroup (thread == query row) ==============================
    const int row = tid;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    float m_used = -INFINITY;  // running max actually used for scaling (log2 units, scale folded in)
    float l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&bars->s_full[s], ph);
      tc_fence_after();
      float sv[BN];
      const uint32_t s_taddr = tmem + lane_base + (s ? C::kColS1 : C::kColS0);
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) tmem_ld32(s_taddr + c * 32, sv + c * 32);
      tmem_wait_ld();
      const int kv_left = N - j * BN;  // keys valid in this tile
      if (kv_left < BN) {
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (i >= kv_left) sv[i] = -INFINITY;
      }
      float mx0 = sv[0], mx1 = sv[1], mx2 = sv[2], mx3 = sv[3];
#pragma unroll
      for (int i = 4; i < BN; i += 4) {
        mx0 = fmaxf(mx0, sv[i]); mx1 = fmaxf(mx1, sv[i + 1]); mx2 = fmaxf(mx2, sv[i + 2]); mx3 = fmaxf(mx3, sv[i + 3]);
      }
      const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
      // lazy rescale: only move the reference max when it grew by more than 2^8
      const bool need = m_new > m_used + kRescaleThreshold;
      if (j == 0) {
        m_used = m_new;
      } else if (__any_sync(0xffffffffu, need)) {
        // O must be quiescent: PV_{j-1} finished (it was triggered by our own p_full arrival of tile j-1)
        mbar_wait(&bars->pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const float m_next = need ? m_new : m_used;
        const float f = ex2(m_used - m_next);
        l_run *= f;
        m_used = m_next;
#pragma unroll
        for (int c = 0; c < C::kDP / 16; ++c) {
          float o[16];
          tmem_ld16(tmem + lane_base + C::kColO + c * 16, o);
          tmem_wait_ld();
          uint32_t u[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(o[i] * f);
          tmem_st16(tmem + lane_base + C::kColO + c * 16, u);
        }
        tmem_wait_st();
      }
      // P = 2^(s*scale_log2 - m_used), row sum, bf16
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int i = 0; i < BN; i += 2) {
        sv[i] = ex2(fmaf(sv[i], scale_log2, -m_used));
        sv[i + 1] = ex2(fmaf(sv[i + 1], scale_log2, -m_used));
        sum0 += sv[i]; sum1 += sv[i + 1];
      }
      l_run += sum0 + sum1;
      if (kPTmem) {
        // P overwrites the first BN/2 columns of the S buffer it was computed from (2 bf16 per 32-bit column)
        if (j >= 2) {
          // nothing to wait for: PV_{j-2} read P from this buffer before QK_j (in-order MMA pipe) overwrote it
        }
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t u[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = pack_bf16(sv[c * 32 + 2 * i], sv[c * 32 + 2 * i + 1]);
          tmem_st16(s_taddr + c * 16, u);
        }
        tmem_wait_st();
        tc_fence_before();
      } else {
        // the P buffer of tile j-2 must have been consumed by PV_{j-2}
        if (j >= 2) mbar_wait(&bars->pv_done[s], ((j >> 1) + 1) & 1);
        unsigned char* prow = sP + s * C::kPBytes + row * 128;
#pragma unroll
        for (int c = 0; c < BN / 8; ++c) {  // 16-byte chunks of 8 keys
          uint4 pk;
          pk.x = pack_bf16(sv[c * 8 + 0], sv[c * 8 + 1]); pk.y = pack_bf16(sv[c * 8 + 2], sv[c * 8 + 3]);
          pk.z = pack_bf16(sv[c * 8 + 4], sv[c * 8 + 5]); pk.w = pack_bf16(sv[c * 8 + 6], sv[c * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + (c >> 3) * kBlockM * 128 + (((c & 7) ^ (row & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        tc_fence_before();
      }
      mbar_arrive(&bars->p_full[s]);
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(&bars->pv_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int n = q0 + row;
    __nv_bfloat16* orow = out + (static_cast<long long>(b) * N + n) * (H * D) + h * D;
#pragma unroll
    for (int c = 0; c < C::kDP / 16; ++c) {
      float o[16];
      tmem_ld16(tmem + lane_base + C::kColO + c * 16, o);
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
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace sm100

// ------------------------------------------------------------------ host side ---------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// [B, N, H*d] bf16 viewed as (d, H, N, B); box (64, 1, rows, 1), 128-byte swizzle, zero fill out of bounds.
// cuTensorMapEncodeTiled is a DRIVER entry point: it needs the primary context current on the calling thread.  A thread
// that has made no runtime call yet (PyTorch's autograd worker on its first backward kernel) has none bound.
int make_head_map_uncached(CUtensorMap* map, const void* base, int B, int H, int N, int d, int box_rows, long long row_stride);

void bind_primary_context() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

namespace {
struct TensorMapCache {
  static constexpr int kSlots = 64;
  TensorMapKey keys[kSlots];
  CUtensorMap maps[kSlots];
  int used = 0, next = 0;
};
TensorMapCache& tensor_map_cache() {
  static thread_local TensorMapCache c;
  return c;
}
bool same_key(const TensorMapKey& a, const TensorMapKey& b) {
  if (a.base != b.base || a.dtype != b.dtype || a.rank != b.rank || a.swizzle != b.swizzle) return false;
  for (int i = 0; i < 4; ++i)
    if (a.dims[i] != b.dims[i] || a.box[i] != b.box[i]) return false;
  for (int i = 0; i < 3; ++i)
    if (a.strides[i] != b.strides[i]) return false;
  return true;
}
}  // namespace

bool tensor_map_cache_get(const TensorMapKey& key, CUtensorMap* out) {
  TensorMapCache& c = tensor_map_cache();
  for (int i = 0; i < c.used; ++i)
    if (same_key(c.keys[i], key)) {
      *out = c.maps[i];
      return true;
    }
  return false;
}
void tensor_map_cache_put(const TensorMapKey& key, const CUtensorMap& map) {
  TensorMapCache& c = tensor_map_cache();
  const int slot = c.next;
  c.keys[slot] = key;
  c.maps[slot] = map;
  c.next = (c.next + 1) % TensorMapCache::kSlots;
  if (c.used < TensorMapCache::kSlots) ++c.used;
}

int make_head_map(CUtensorMap* map, const void* base, int B, int H, int N, int d, int box_rows, long long row_stride) {
  {
    const unsigned long long Cs = row_stride ? static_cast<unsigned long long>(row_stride) : static_cast<unsigned long long>(H) * d;
    TensorMapKey key = {base, {static_cast<unsigned long long>(d), static_cast<unsigned long long>(H),
                               static_cast<unsigned long long>(N), static_cast<unsigned long long>(B)},
                        {static_cast<unsigned long long>(d) * 2, Cs * 2, static_cast<unsigned long long>(N) * Cs * 2},
                        {64u, 1u, static_cast<unsigned>(box_rows), 1u}, 1, 4, 128};
    if (tensor_map_cache_get(key, map)) return AGENDA_OK;
    int rc = make_head_map_uncached(map, base, B, H, N, d, box_rows, row_stride);
    if (rc == AGENDA_OK) tensor_map_cache_put(key, *map);
    return rc;
  }
}

int make_head_map_uncached(CUtensorMap* map, const void* base, int B, int H, int N, int d, int box_rows, long long row_stride) {
  bind_primary_context();
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(AGENDA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t C = row_stride ? static_cast<cuuint64_t>(row_stride) : static_cast<cuuint64_t>(H) * d;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(d) * 2, C * 2, static_cast<cuuint64_t>(N) * C * 2};
  cuuint32_t box[4] = {64, 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AGENDA_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  return AGENDA_OK;
}

template <int D, bool kPTmem>
static int launch_sm100(const void* q, const void* k, const void* v, void* out, int B, int H, int N, float scale,
                        cudaStream_t stream, long long ld) {
  using C = sm100::Cfg<D>;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_head_map(&mq, q, B, H, N, D, sm100::kBlockM, ld)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mk, k, B, H, N, D, C::kBlockN, ld)) != AGENDA_OK) return rc;
  if ((rc = make_head_map(&mv, v, B, H, N, D, C::kBlockN, ld)) != AGENDA_OK) return rc;
  constexpr size_t smem = sm100::smem_bytes<D, kPTmem>();
  auto kern = sm100::attn_self_sm100_kernel<D, kPTmem>;
  AGENDA_DYN_SMEM(kern, smem);
  dim3 grid((N + sm100::kBlockM - 1) / sm100::kBlockM, B * H);
  kern<<<grid, sm100::kThreads, smem, stream>>>(mq, mk, mv, static_cast<__nv_bfloat16*>(out), H, N,
                                                scale * 1.4426950408889634f);
  AGENDA_LAUNCH_CHECK("attn_self_sm100_kernel");
  return AGENDA_OK;
}

int attn_common_checks(const char* who, const void* q, const void* k, const void* v, void* out, int dtype, int B,
                       int H, int N, int M, int d);

int attn_self_sm100_v2(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int d, float scale,
                       int emu, int tiles, void* stream, long long ld, float* lse = nullptr);

// variant: 0 = two query tiles per CTA, ping-pong softmax warpgroups (v2; falls back to variant 1 when N <= 128),
//          1 = one query tile per CTA, P through TMEM (TS-form PV MMA), 2 = same with P through shared memory (SS)
// ld: row stride of q/k/v in elements (0 = packed H*d); q_prescaled: q already carries scale * log2(e) (ABI: scale == 0)
// lse (may be NULL): [B, H, N] fp32, the base-2 log-sum-exp of every row's scaled scores — emitted by the multi-tile kernels
// only (N > 128, default variant); asking for it elsewhere is AGENDA_ERR_UNSUPPORTED and the caller runs the LSE pass.
int attn_self_sm100(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int d, float scale,
                    int variant, void* stream, long long ld = 0, bool q_prescaled = false, float* lse = nullptr) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out);
  if (al & 15) return fail(AGENDA_ERR_MISALIGNED, "attn_self_fwd: q/k/v/out must be 16-byte aligned");
  // variants 10/12/13/14/18: v2 with 0 / 50 / 37.5 / 25 / 12.5 % of the exponentials emulated on the FMA pipe
  // variants 20/22/23/24/28: the same emulation shares with three query tiles per CTA and 64-key tiles (d = 40 / 64)
  // variants 30/32/33/34/38: two query tiles, two softmax warpgroups per tile (half a row per thread; d = 40 / 64)
  // variants 40/42/43/44/48: two query tiles with 64-key tiles (d = 80: separate P columns instead of S/P aliasing)
  // variants 50/52/53/54/58: three query tiles with 80-key tiles (d = 40)
  // variant 60: the shipped fast path (first-tile maximum + row-sum check + exact second pass)
  // (62 / 63 / 64: the same with 50 / 37.5 / 25 % emulated exponentials where instantiated: d = 40, 64)
  if (variant >= 60) {
    const int e = variant - 60;
    if (d == 80 && e == 5) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, 4, 105, stream, ld);  // 64-key tiles
    return d == 40 ? attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, e ? e : 3, 103, stream, ld)
                   : attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, e ? e : 4, 102, stream, ld);
  }
  if (variant >= 50) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, variant - 50, 6, stream, ld);
  if (variant >= 40) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, variant - 40, 5, stream, ld);
  if (variant >= 30) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, variant - 30, 4, stream, ld);
  if (variant >= 20) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, variant - 20, 3, stream, ld);
  if (variant >= 10) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, variant - 10, 2, stream, ld);
  // defaults measured on B200 (tools/bench_attn.py): d = 40: three query tiles with 64-key tiles, 37.5 % of the
  // exponentials on the FMA pipe; otherwise two query tiles with 128-key tiles (64 at d = 160), 25 %
  // and the fast first pass (first-tile maximum + row-sum check, exact second pass for the CTAs that need it);
  // AGENDA_V2_FAST=0 keeps the running maximum in a single pass (measurements)
  // agenda_attn_self_fwd_strided with scale == 0: q already carries scale * log2(e)
  // AGENDA_V2_FAST=0 (variant builds only) keeps the running maximum in a single pass (measurements)
#ifdef AGENDA_VARIANTS
  static const int fast = [] { const char* e = knob("AGENDA_V2_FAST"); return (e && atoi(e) == 0) ? 0 : 100; }();
#else
  constexpr int fast = 100;
#endif
  if (q_prescaled) {
    scale = 0.6931471805599453f;  // kernels without the unit-scale form: scale * log2(e) = 1
    if (variant == 0 && N > 128 && d == 40 && fast) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, 3, 203, stream, ld);
  }
  if (variant == 0 && N > 128) {
    // (d = 160 only occurs at N <= 256 in the SD UNets: four key tiles do not amortise the fast pass's epilogue)
    // d = 80: 64-key tiles so that P gets its own TMEM columns (0.071 vs 0.076 ms with P aliased onto S at N = 1024)
    if (d == 80 && fast) return attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, 4, 105, stream, ld, lse);
    return d == 40 ? attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, 3, 3 + fast, stream, ld, lse)
                   : attn_self_sm100_v2(q, k, v, out, B, H, N, d, scale, 4, 2 + (d == 160 ? 0 : fast), stream, ld, lse);
  }
  if (lse != nullptr) return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd_lse: N=%d (<= 128) takes a kernel that does not emit the log-sum-exp", N);
#define AGENDA_DISPATCH(DD)                                                                                \
  case DD:                                                                                                 \
    return variant <= 1 ? launch_sm100<DD, true>(q, k, v, out, B, H, N, scale, st, ld)                        \
                        : launch_sm100<DD, false>(q, k, v, out, B, H, N, scale, st, ld);
  switch (d) {
    AGENDA_DISPATCH(40)
    AGENDA_DISPATCH(64)
    AGENDA_DISPATCH(80)
    AGENDA_DISPATCH(160)
    default:
      return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: head dim %d not in {40,64,80,160}", d);
  }
#undef AGENDA_DISPATCH
}

}  // namespace agenda

using namespace agenda;

extern "C" int agenda_attn_self_fwd(const void* q, const void* k, const void* v, void* out, int dtype, int B, int H,
                                    int N, int d, float scale, void* stream) {
  int rc = attn_common_checks("attn_self_fwd", q, k, v, out, dtype, B, H, N, N, d);
  if (rc != AGENDA_OK) return rc;
  if (dtype != AGENDA_BF16) return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd: tensor-core path takes bf16 (dtype=1)");
  return attn_self_sm100(q, k, v, out, B, H, N, d, scale, 0, stream);
}

extern "C" int agenda_attn_self_fwd_emits_lse(int N, int d) {
  return (N > 128 && (d == 40 || d == 64 || d == 80 || d == 160)) ? 1 : 0;
}

extern "C" int agenda_attn_self_fwd_lse(const void* q, const void* k, const void* v, void* out, float* lse, int dtype, int B,
                                        int H, int N, int d, float scale, void* stream) {
  int rc = attn_common_checks("attn_self_fwd_lse", q, k, v, out, dtype, B, H, N, N, d);
  if (rc != AGENDA_OK) return rc;
  if (dtype != AGENDA_BF16) return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd_lse: tensor-core path takes bf16 (dtype=1)");
  if (!lse || (reinterpret_cast<uintptr_t>(lse) & 3)) return fail(AGENDA_ERR_NULL_POINTER, "attn_self_fwd_lse: lse is null or misaligned");
  return attn_self_sm100(q, k, v, out, B, H, N, d, scale, 0, stream, 0, false, lse);
}

extern "C" int agenda_attn_self_fwd_strided(const void* q, const void* k, const void* v, void* out, int dtype, int B,
                                            int H, int N, int d, long long ld, float scale, void* stream) {
  int rc = attn_common_checks("attn_self_fwd_strided", q, k, v, out, dtype, B, H, N, N, d);
  if (rc != AGENDA_OK) return rc;
  if (dtype != AGENDA_BF16) return fail(AGENDA_ERR_UNSUPPORTED, "attn_self_fwd_strided: bf16 only (dtype=1)");
  if (ld < static_cast<long long>(H) * d || (ld & 7))
    return fail(AGENDA_ERR_BAD_SHAPE, "attn_self_fwd_strided: ld=%lld must be >= H*d=%d and a multiple of 8", ld, H * d);
  return attn_self_sm100(q, k, v, out, B, H, N, d, scale, 0, stream, ld, scale == 0.0f);
}

// Test hook: same contract, explicit kernel variant (see attn_self_sm100).
extern "C" int agenda_attn_self_fwd_variant(const void* q, const void* k, const void* v, void* out, int B, int H, int N,
                                            int d, float scale, int variant, void* stream) {
  int rc = attn_common_checks("attn_self_fwd_variant", q, k, v, out, AGENDA_BF16, B, H, N, N, d);
  if (rc != AGENDA_OK) return rc;
  return attn_self_sm100(q, k, v, out, B, H, N, d, scale, variant, stream);
}
