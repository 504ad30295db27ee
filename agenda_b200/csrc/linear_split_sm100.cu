// to_q for the cross-attention heat path (hook.py:93) with an fp32 result, as ONE tcgen05 GEMM:
//
//     out[M, N] (fp32) = x[M, K] (bf16) * (W_hi + W_lo)[N, K]^T          W_hi = bf16(W), W_lo = bf16(W - W_hi)
//
// The heat-map logits need the checkpoint's fp32 projection weights (agenda_b200/mixed.py), i.e. two bf16 products per
// activation.  With library GEMMs that is an fp32-output GEMM plus an accumulating second GEMM that reads and rewrites
// the whole fp32 output (65 us against 27 us for one GEMM at M = 65536, K = N = 320, profiles/r02_to_q_gemm_forms.txt).
// Here both weight halves are B operands of the same accumulator: the activation tile is read once, the output is
// written once, and the correction costs only tensor-pipe time the kernel does not need (it is bound by the fp32 store
// stream and the L2 weight stream).  w_lo == NULL gives the plain fp32-output GEMM.
//
// Persistent, warp-specialised (the canonical sm_100 GEMM shape): CTA tile 128 x 160, K in blocks of 64 (128-byte
// swizzled rows) through a 3-stage TMA ring; accumulators double-buffered in TMEM (2 x 160 columns) so the epilogue of
// tile i overlaps the MMAs of tile i + 1; epilogue warps stage 32-column fp32 sub-tiles in the 128B-swizzled layout
// (conflict-free stores) and write them with TMA tensor stores (full 128-byte lines, rows clipped by the hardware).
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2-5: epilogue (TMEM lane quadrant = warp % 4)
#include <algorithm>
#include <cstdlib>

#include "sm100_common.cuh"

namespace agenda {
namespace sm100 {

constexpr int kLThreads = 192;
constexpr int kLBM = 128, kLBN = 160, kLBK = 64;
constexpr int kLStages = 3;
constexpr int kLABytes = kLBM * 128;              // 128 rows x 64 bf16
constexpr int kLBBytes = kLBN * 128;              // 160 rows x 64 bf16
constexpr int kLStageBytes = kLABytes + 2 * kLBBytes;
constexpr int kLSubBytes = 128 * 128;             // one fp32 output sub-tile: 128 rows x 32 columns
constexpr int kLStaging = 3;                      // sub-tiles staged at a time (160 columns = 3 + 2 sub-tiles)

struct LBarriers {
  uint64_t full[kLStages], empty[kLStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

constexpr size_t l_smem_bytes() { return 1024 + kLStages * kLStageBytes + kLStaging * kLSubBytes + sizeof(LBarriers) + 64; }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_load_l(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void st_shared_f4(void* p, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kHeads: the output leaves in the chunk-major layout the split-precision cross-attention kernel streams,
//   out[b][h][c][n][40]   (b = m / rows_per_batch, n = m % rows_per_batch, column h*d + c*40 + j -> (h, c, j)),
// i.e. every (batch, head, 40-column chunk) is a dense [N, 40] fp32 plane, so the kernel fetches a 128-query chunk with ONE
// bulk copy of 20 KB instead of a 128-row TMA tensor load (4-5 TMA-engine cycles per row).  A thread owns a row of the tile:
// 40 contiguous floats per chunk, written straight from the TMEM registers (no staging, no tensor-map store).
template <bool kLo, bool kHeads>
__global__ void __launch_bounds__(kLThreads, 1)
linear_split_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                    const __grid_constant__ CUtensorMap map_wlo, const __grid_constant__ CUtensorMap map_out, int M, int K,
                    int N, float* __restrict__ out_hm, int rows_per_batch, int chunks_per_head,
                    const unsigned char* __restrict__ w_blob) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sStage = smem;
  unsigned char* sOut = sStage + kLStages * kLStageBytes;
  LBarriers* bars = reinterpret_cast<LBarriers*>(sOut + kLStaging * kLSubBytes);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int tiles_m = (M + kLBM - 1) / kLBM, tiles_n = N / kLBN;
  const int n_tiles = tiles_m * tiles_n;
  const int n_kb = K / kLBK;

  if (tid == 0) {
    tma_prefetch_desc(&map_x);
    if (w_blob == nullptr) { tma_prefetch_desc(&map_whi); if (kLo) tma_prefetch_desc(&map_wlo); }
    if (!kHeads) tma_prefetch_desc(&map_out);
    for (int s = 0; s < kLStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->acc_full[a], 1); mbar_init(&bars->acc_empty[a], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;   // n fastest: the CTAs that share an activation tile run together
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&bars->empty[s], ph ^ 1);
        if (elect_one()) {
          unsigned char* a = sStage + s * kLStageBytes;
          mbar_expect_tx(&bars->full[s], kLABytes + (kLo ? 2 : 1) * kLBBytes);
          tma_load_2d(&map_x, &bars->full[s], a, kb * kLBK, tm * kLBM);
          if (w_blob != nullptr) {
            // pre-packed weights (agenda_linear_split_pack_w): the K block's hi | lo tiles are one contiguous image of this
            // stage's B area -> ONE bulk copy instead of 160 (+160) tensor-map rows (the TMA engine's per-row cost bounds
            // this kernel: DESIGN.md section 8 item 2)
            constexpr uint32_t kWBytes = (kLo ? 2 : 1) * kLBBytes;
            bulk_load_l(a + kLABytes, w_blob + (static_cast<size_t>(tn) * n_kb + kb) * kWBytes, kWBytes, &bars->full[s]);
          } else {
            tma_load_2d(&map_whi, &bars->full[s], a + kLABytes, kb * kLBK, tn * kLBN);
            if (kLo) tma_load_2d(&map_wlo, &bars->full[s], a + kLABytes + kLBBytes, kb * kLBK, tn * kLBN);
          }
        }
        __syncwarp();
        if (++s == kLStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc(kLBM, kLBN, 0);
    const uint64_t desc0 = make_sdesc(smem_u32(sStage), 16, 1024);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      mbar_wait(&bars->acc_empty[ab], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator (two tiles ago)
      tc_fence_after();
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&bars->full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_off = s * kLStageBytes, bh_off = a_off + kLABytes, bl_off = bh_off + kLBBytes;
#pragma unroll
          for (int kk = 0; kk < kLBK / 16; ++kk) {
            const uint64_t ad = desc0 + static_cast<uint64_t>((a_off + kk * 32) >> 4);
            umma_ss(tmem + ab * kLBN, ad, desc0 + static_cast<uint64_t>((bh_off + kk * 32) >> 4), idesc, !(kb == 0 && kk == 0));
            if (kLo) umma_ss(tmem + ab * kLBN, ad, desc0 + static_cast<uint64_t>((bl_off + kk * 32) >> 4), idesc, 1);
          }
          umma_commit(&bars->empty[s]);
          if (kb == n_kb - 1) umma_commit(&bars->acc_full[ab]);
        }
        __syncwarp();
        if (++s == kLStages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue: TMEM -> swizzled fp32 sub-tiles -> TMA store ==============================
    const int et = tid - 64;                                   // 0..127
    const int quad = warp & 3;                                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + (tid & 31);                    // row of the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      const int ab = it & 1;
      mbar_wait(&bars->acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      if constexpr (kHeads) {
        const int m = tm * kLBM + row;
        const int bb = m / rows_per_batch, nn = m - bb * rows_per_batch;
        const int heads = N / (chunks_per_head * 40);
        float* base[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = tn * 4 + j, hh = cc / chunks_per_head, c = cc - hh * chunks_per_head;
          base[j] = out_hm + ((((static_cast<long long>(bb) * heads + hh) * chunks_per_head + c) * rows_per_batch + nn) * 40);
        }
#pragma unroll
        for (int g = 0; g < kLBN / 16; ++g) {
          float v[16];
          tmem_ld16(tmem + lane_base + ab * kLBN + g * 16, v);
          tmem_wait_ld();
          if (m < M) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int col = g * 16 + q * 4;   // (compile-time) column of the tile: chunk col / 40, offset col % 40
              *reinterpret_cast<float4*>(base[col / 40] + (col % 40)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
        }
      } else {
#pragma unroll
      for (int sub = 0; sub < kLBN / 32; ++sub) {
        float v[32];
        tmem_ld32(tmem + lane_base + ab * kLBN + sub * 32, v);
        tmem_wait_ld();
        const int slot = sub % kLStaging;
        if (slot == 0) {
          // the TMA stores that read the staging buffers have finished reading (issued by thread et == 0)
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        unsigned char* dst = sOut + slot * kLSubBytes + row * 128;
#pragma unroll
        for (int p = 0; p < 8; ++p)
          st_shared_f4(dst + ((p ^ (row & 7)) << 4), v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);
        if (slot == kLStaging - 1 || sub == kLBN / 32 - 1) {
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            const int first = sub - slot;
            for (int q = first; q <= sub; ++q)
              tma_store_2d(&map_out, sOut + (q % kLStaging) * kLSubBytes, tn * kLBN + q * 32, tm * kLBM);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_empty[ab]);
    }
    if (!kHeads && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Weight pre-packing for the bulk-copy form: blob[tn][kb] = hi tile | lo tile, each 160 rows x 128 bytes in the 128B-swizzled
// K-major UMMA layout (16-byte piece p of row r at p ^ (r & 7)) — exactly what the tensor-map loads write into a stage.
// One thread per 16-byte piece.
__global__ void linear_split_pack_kernel(const uint4* __restrict__ w_hi, const uint4* __restrict__ w_lo, uint4* __restrict__ blob,
                                         int N, int K) {
  const int n_kb = K / kLBK, halves = w_lo ? 2 : 1;
  const long long total = static_cast<long long>(N / kLBN) * n_kb * halves * kLBN * 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i & 7);
    long long rest = i >> 3;
    const int r = static_cast<int>(rest % kLBN); rest /= kLBN;
    const int half = static_cast<int>(rest % halves); rest /= halves;
    const int kb = static_cast<int>(rest % n_kb);
    const int tn = static_cast<int>(rest / n_kb);
    const int sp = p ^ (r & 7);   // source piece stored at position p
    const uint4* src = (half ? w_lo : w_hi) + (static_cast<long long>(tn * kLBN + r) * K + kb * kLBK) / 8 + sp;
    blob[i] = *src;
  }
}

}  // namespace sm100

typedef CUresult (*EncodeTiledFnL)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// row-major [rows, cols] matrix viewed as (cols, rows); box (box_cols, box_rows), 128-byte swizzle
static int make_matrix_map(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem, long long rows, long long cols,
                           int box_cols, int box_rows) {
  const TensorMapKey key = {base, {static_cast<unsigned long long>(cols), static_cast<unsigned long long>(rows), 1ull, 1ull},
                            {static_cast<unsigned long long>(cols) * elem, 0ull, 0ull},
                            {static_cast<unsigned>(box_cols), static_cast<unsigned>(box_rows), 1u, 1u}, static_cast<int>(dt), 2, 128};
  if (tensor_map_cache_get(key, map)) return AGENDA_OK;
  bind_primary_context();
  static EncodeTiledFnL enc = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFnL>(p);
  }();
  if (!enc) return fail(AGENDA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * elem};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AGENDA_ERR_CUDA, "cuTensorMapEncodeTiled (matrix) failed (CUresult %d)", static_cast<int>(r));
  tensor_map_cache_put(key, *map);
  return AGENDA_OK;
}

}  // namespace agenda

using namespace agenda;

// w_blob != NULL: packed weights (w_hi / w_lo are then not dereferenced; has_lo says whether the blob carries the lo half)
static int linear_split_impl(const char* who, const void* x, const void* w_hi, const void* w_lo, float* out, int M, int K, int N,
                             int rows_per_batch, int heads, void* stream, const void* w_blob = nullptr, int has_lo = 0) {
  if (!x || (!w_hi && !w_blob) || !out) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (M <= 0 || K <= 0 || N <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "%s: M=%d K=%d N=%d", who, M, K, N);
  if (K % sm100::kLBK || N % sm100::kLBN)
    return fail(AGENDA_ERR_UNSUPPORTED, "%s: K=%d must be a multiple of %d and N=%d a multiple of %d", who, K, sm100::kLBK, N,
                sm100::kLBN);
  const bool hm = heads > 0;
  int chunks_per_head = 0;
  if (hm) {
    if (rows_per_batch <= 0 || M % rows_per_batch || N % heads || (N / heads) % 40)
      return fail(AGENDA_ERR_UNSUPPORTED, "%s: head-major output needs M %% rows_per_batch == 0 and a head dim that is a multiple "
                  "of 40 (M=%d rows_per_batch=%d N=%d heads=%d)", who, M, rows_per_batch, N, heads);
    chunks_per_head = N / heads / 40;
  }
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(w_lo) |
                       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w_blob);
  if (al & 15) return fail(AGENDA_ERR_MISALIGNED, "%s: x / w_hi / w_lo / w_blob / out must be 16-byte aligned", who);
  CUtensorMap mx, mh, ml, mo;
  int rc;
  if ((rc = make_matrix_map(&mx, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, sm100::kLBK, sm100::kLBM)) != AGENDA_OK) return rc;
  if (w_blob) { mh = mx; ml = mx; }   // (not used with packed weights; any valid descriptor)
  else {
    if ((rc = make_matrix_map(&mh, w_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, sm100::kLBK, sm100::kLBN)) != AGENDA_OK) return rc;
    if ((rc = make_matrix_map(&ml, w_lo ? w_lo : w_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, sm100::kLBK, sm100::kLBN)) != AGENDA_OK) return rc;
  }
  const bool lo = w_blob ? (has_lo != 0) : (w_lo != nullptr);
  if (hm) mo = mx;   // (not used by the head-major epilogue; any valid descriptor)
  else if ((rc = make_matrix_map(&mo, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, 32, sm100::kLBM)) != AGENDA_OK) return rc;
  const int n_tiles = ((M + sm100::kLBM - 1) / sm100::kLBM) * (N / sm100::kLBN);
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  constexpr size_t smem = sm100::l_smem_bytes();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define AGENDA_LSPLIT(LO, HM)                                                                                  \
  do {                                                                                                         \
    auto kern = sm100::linear_split_kernel<LO, HM>;                                                            \
    AGENDA_DYN_SMEM(kern, smem);                                                                               \
    kern<<<grid, sm100::kLThreads, smem, st>>>(mx, mh, ml, mo, M, K, N, out, rows_per_batch, chunks_per_head,  \
                                               static_cast<const unsigned char*>(w_blob));                     \
  } while (0)
  if (lo) { if (hm) AGENDA_LSPLIT(true, true); else AGENDA_LSPLIT(true, false); }
  else { if (hm) AGENDA_LSPLIT(false, true); else AGENDA_LSPLIT(false, false); }
#undef AGENDA_LSPLIT
  AGENDA_LAUNCH_CHECK("linear_split_kernel");
  return AGENDA_OK;
}

extern "C" int agenda_linear_split_f32(const void* x, const void* w_hi, const void* w_lo, float* out, int M, int K, int N,
                                       void* stream) {
  return linear_split_impl("linear_split_f32", x, w_hi, w_lo, out, M, K, N, 0, 0, stream);
}

extern "C" int agenda_linear_split_f32_heads(const void* x, const void* w_hi, const void* w_lo, float* out, int M, int K, int N,
                                             int rows_per_batch, int heads, void* stream) {
  if (heads <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "linear_split_f32_heads: heads=%d", heads);
  return linear_split_impl("linear_split_f32_heads", x, w_hi, w_lo, out, M, K, N, rows_per_batch, heads, stream);
}

extern "C" long long agenda_linear_split_pack_bytes(int N, int K, int has_lo) {
  if (N <= 0 || K <= 0 || K % sm100::kLBK || N % sm100::kLBN)
    return fail(AGENDA_ERR_UNSUPPORTED, "linear_split_pack_bytes: K=%d must be a multiple of %d and N=%d a multiple of %d", K,
                sm100::kLBK, N, sm100::kLBN);
  return static_cast<long long>(N) * K * 2 * (has_lo ? 2 : 1);
}

extern "C" int agenda_linear_split_pack_w(const void* w_hi, const void* w_lo, void* blob, int N, int K, void* stream) {
  const char* who = "linear_split_pack_w";
  if (!w_hi || !blob) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  if (N <= 0 || K <= 0 || K % sm100::kLBK || N % sm100::kLBN)
    return fail(AGENDA_ERR_UNSUPPORTED, "%s: K=%d must be a multiple of %d and N=%d a multiple of %d", who, K, sm100::kLBK, N, sm100::kLBN);
  if ((reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(w_lo) | reinterpret_cast<uintptr_t>(blob)) & 15)
    return fail(AGENDA_ERR_MISALIGNED, "%s: w_hi / w_lo / blob must be 16-byte aligned", who);
  const long long total = static_cast<long long>(N) * K / 8 * (w_lo ? 2 : 1);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 4096));
  sm100::linear_split_pack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(w_hi), static_cast<const uint4*>(w_lo), static_cast<uint4*>(blob), N, K);
  AGENDA_LAUNCH_CHECK("linear_split_pack_kernel");
  return AGENDA_OK;
}

extern "C" int agenda_linear_split_f32_packed(const void* x, const void* w_blob, int has_lo, float* out, int M, int K, int N,
                                              void* stream) {
  if (!w_blob) return fail(AGENDA_ERR_NULL_POINTER, "linear_split_f32_packed: w_blob is null");
  return linear_split_impl("linear_split_f32_packed", x, nullptr, nullptr, out, M, K, N, 0, 0, stream, w_blob, has_lo);
}
