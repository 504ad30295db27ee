// K6: threshold + 4-connected component labelling + bbox (spec: SURVEY.md §8 a9; not present in the reference).
//
// B200 design: ONE THREAD-BLOCK CLUSTER PER MAP, the whole fp32 map resident in distributed shared memory.
//   * each CTA of the cluster bulk-copies (TMA engine, cp.async.bulk + mbarrier) its strip of R rows into smem,
//   * min/max are reduced CTA-locally and exchanged through DSMEM, so the map is read from HBM exactly once
//     even though normalisation needs the global min/max before the first pixel can be thresholded,
//   * the strip buffer is then re-used IN PLACE as the union-find parent array (one u32 per pixel, labels are
//     map-linear pixel indices, roots = smallest index of the component),
//   * runs are seeded with warp ballots (one warp = 32 consecutive pixels of a row), merged vertically with
//     shared-memory atomicMin union-find, strips are stitched through DSMEM atomics,
//   * roots are numbered in raster order with ballot/popc scans (+ a DSMEM exchange of per-strip root counts),
//     which reproduces scipy.ndimage.label's numbering bit-exactly,
//   * labels go back to HBM once (coalesced), boxes via per-run global atomics.
// Algorithmic HBM bytes per map: H*W*4 read + H*W*4 written (+ 20 B per box).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace agenda {

constexpr uint32_t kBG = 0xFFFFFFFFu;
constexpr uint32_t kRankFlag = 0x80000000u;
constexpr int kMaxCluster = 16;

struct CclStatic {
  unsigned long long mbar;
  float red_min[32], red_max[32];
  int red_nan[32];
  float x_min[kMaxCluster], x_max[kMaxCluster];
  int x_nan[kMaxCluster];
  int x_roots[kMaxCluster];
  int warp_roots[32];
  int warp_base[32];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- union-find on the CTA-local strip (labels are map-linear indices, strip covers [base, base+strip_px)) ----
__device__ __forceinline__ uint32_t find_local(const uint32_t* L, uint32_t base, uint32_t x) {
  uint32_t p = L[x - base];
  while (p != x) { x = p; p = L[x - base]; }
  return x;
}
__device__ __forceinline__ void unite_local(uint32_t* L, uint32_t base, uint32_t a, uint32_t b) {
  bool done;
  do {
    a = find_local(L, base, a);
    b = find_local(L, base, b);
    if (a < b) { const uint32_t old = atomicMin(&L[b - base], a); done = (old == b); b = old; }
    else if (b < a) { const uint32_t old = atomicMin(&L[a - base], b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// ---- the same over the whole cluster (a label may live in another CTA's shared memory) ----
struct ClusterLabels {
  cg::cluster_group cluster;
  uint32_t* L;         // this CTA's strip
  uint32_t strip_px;   // R*W
  __device__ __forceinline__ uint32_t* slot(uint32_t idx) const {
    const uint32_t rk = idx / strip_px;
    return cluster.map_shared_rank(L, rk) + (idx - rk * strip_px);
  }
  __device__ __forceinline__ uint32_t find(uint32_t x) const {
    uint32_t p = *reinterpret_cast<volatile uint32_t*>(slot(x));
    while (p != x) { x = p; p = *reinterpret_cast<volatile uint32_t*>(slot(x)); }
    return x;
  }
  __device__ __forceinline__ void unite(uint32_t a, uint32_t b) const {
    bool done;
    do {
      a = find(a);
      b = find(b);
      if (a < b) { const uint32_t old = atomicMin(slot(b), a); done = (old == b); b = old; }
      else if (b < a) { const uint32_t old = atomicMin(slot(a), b); done = (old == a); a = old; }
      else done = true;
    } while (!done);
  }
};

__global__ void ccl_bbox_kernel(const float* __restrict__ heat, float thr, int32_t* __restrict__ labels,
                                int32_t* __restrict__ counts, int32_t* __restrict__ boxes, int max_boxes, int H,
                                int W, int R, int use_bulk) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ CclStatic sh;
  uint32_t* L = reinterpret_cast<uint32_t*>(dyn_smem);
  float* Lf = reinterpret_cast<float*>(dyn_smem);

  cg::cluster_group cluster = cg::this_cluster();
  const int cs = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const long long map = blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;

  const int r0 = rank * R;
  const int rows = min(R, H - r0);
  const int strip_px = R * W;
  const int n_px = rows * W;
  const uint32_t base = static_cast<uint32_t>(r0) * W;  // map-linear index of this strip's first pixel
  const float* src = heat + map * H * W + base;

  // ---- 1. strip -> shared memory (bulk async copy through the TMA engine when 16-B aligned) ----
  if (use_bulk) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(n_px) * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sh.mbar)), "r"(bytes)
                   : "memory");
      const uint32_t chunk = 32768u;
      for (uint32_t off = 0; off < bytes; off += chunk) {
        const uint32_t sz = min(chunk, bytes - off);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(dyn_smem + off)),
            "l"(reinterpret_cast<const unsigned char*>(src) + off), "r"(sz), "r"(smem_u32(&sh.mbar))
            : "memory");
      }
    }
    // all threads wait for phase 0
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&sh.mbar)), "r"(0u)
          : "memory");
    }
  } else {
    for (int i = tid; i < n_px; i += blockDim.x) Lf[i] = __ldg(src + i);
    __syncthreads();
  }

  // ---- 2. min / max: CTA reduce, then all-to-all through distributed shared memory ----
  {
    float lo = INFINITY, hi = -INFINITY;
    int nan = 0;
    if ((n_px & 3) == 0) {
      for (int i = tid; i < (n_px >> 2); i += blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(Lf)[i];
        lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
        hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
      }
    } else {
      for (int i = tid; i < n_px; i += blockDim.x) {
        const float v = Lf[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v); nan |= (v != v);
      }
    }
    lo = warp_min(lo); hi = warp_max(hi);
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) { sh.red_min[wid] = lo; sh.red_max[wid] = hi; sh.red_nan[wid] = nan; }
    __syncthreads();
    if (wid == 0) {
      lo = lane < nwarps ? sh.red_min[lane] : INFINITY;
      hi = lane < nwarps ? sh.red_max[lane] : -INFINITY;
      nan = lane < nwarps ? sh.red_nan[lane] : 0;
      lo = warp_min(lo); hi = warp_max(hi);
      nan = __any_sync(0xffffffffu, nan);
      if (lane < cs) {  // lane j publishes this CTA's partials into CTA j's exchange slots
        *cluster.map_shared_rank(&sh.x_min[rank], lane) = lo;
        *cluster.map_shared_rank(&sh.x_max[rank], lane) = hi;
        *cluster.map_shared_rank(&sh.x_nan[rank], lane) = nan;
      }
    }
  }
  cluster.sync();
  float mn = INFINITY, mx = -INFINITY;
  int any_nan = 0;
  for (int j = 0; j < cs; ++j) {
    mn = fminf(mn, sh.x_min[j]); mx = fmaxf(mx, sh.x_max[j]); any_nan |= sh.x_nan[j];
  }
  const float denom = np_denominator(mn, mx);

  const int segs = (W + 31) >> 5;
  const int n_seg = rows * segs;

  // ---- 3. threshold, seed every pixel with the index of the first pixel of its (32-px-segment) run ----
  for (int s = wid; s < n_seg; s += nwarps) {
    const int y = s / segs, x = ((s - y * segs) << 5) + lane;
    const bool valid = x < W;
    const int i = y * W + x;
    bool fg = false;
    if (valid) fg = !any_nan && (np_normalize(Lf[i], mn, denom) > thr);
    const uint32_t m = __ballot_sync(0xffffffffu, fg);
    if (valid) {
      uint32_t lab = kBG;
      if (fg) {
        const uint32_t zeros_below = ~m & ((1u << lane) - 1u);
        const int start = zeros_below ? (32 - __clz(zeros_below)) : 0;
        lab = base + static_cast<uint32_t>(i - (lane - start));
      }
      L[i] = lab;
    }
  }
  __syncthreads();

  // ---- 4. merge inside the strip: across 32-px segment boundaries and with the row above ----
  for (int s = wid; s < n_seg; s += nwarps) {
    const int y = s / segs, x = ((s - y * segs) << 5) + lane;
    const bool valid = x < W;
    const int i = y * W + x;
    const bool fg = valid && (L[i] != kBG);
    const bool up = fg && y > 0 && (L[i - W] != kBG);
    const uint32_t m = __ballot_sync(0xffffffffu, fg);
    const uint32_t mu = __ballot_sync(0xffffffffu, valid && y > 0 && (L[i - W] != kBG));
    if (fg) {
      bool left, upleft;
      if (lane) { left = (m >> (lane - 1)) & 1u; upleft = (mu >> (lane - 1)) & 1u; }
      else {
        left = x > 0 && (L[i - 1] != kBG);
        upleft = x > 0 && y > 0 && (L[i - W - 1] != kBG);
        if (left) unite_local(L, base, base + i, base + i - 1);
      }
      if (up && !(left && upleft)) unite_local(L, base, base + i, base + i - W);
    }
  }
  cluster.sync();  // every strip holds labels (not floats) and is internally merged

  ClusterLabels CL{cluster, L, static_cast<uint32_t>(strip_px)};

  // ---- 5. stitch with the strip above through distributed shared memory ----
  if (rank > 0) {
    const uint32_t* prev = cluster.map_shared_rank(L, rank - 1) + (R - 1) * W;  // last row of the strip above
    for (int s = wid; s < segs; s += nwarps) {
      const int x = (s << 5) + lane;
      const bool valid = x < W;
      const bool fg = valid && (L[x] != kBG);
      const bool upv = valid && (prev[x] != kBG);
      const uint32_t m = __ballot_sync(0xffffffffu, fg);
      const uint32_t mu = __ballot_sync(0xffffffffu, upv);
      if (fg && upv) {
        bool left, upleft;
        if (lane) { left = (m >> (lane - 1)) & 1u; upleft = (mu >> (lane - 1)) & 1u; }
        else { left = x > 0 && (L[x - 1] != kBG); upleft = x > 0 && (prev[x - 1] != kBG); }
        if (!(left && upleft)) CL.unite(base + x, base + x - W);
      }
    }
  }
  cluster.sync();  // forest is final

  // ---- 6. flatten run heads to their root, count roots in raster order (contiguous segment chunk per warp) ----
  const int chunk = (n_seg + nwarps - 1) / nwarps;
  const int s_begin = min(wid * chunk, n_seg), s_end = min(s_begin + chunk, n_seg);
  int my_roots = 0;
  for (int s = s_begin; s < s_end; ++s) {
    const int y = s / segs, x = ((s - y * segs) << 5) + lane;
    const bool valid = x < W;
    const int i = y * W + x;
    const uint32_t g = base + i;
    const bool fg = valid && (L[i] != kBG);
    const uint32_t m = __ballot_sync(0xffffffffu, fg);
    const bool head = fg && (lane == 0 || !((m >> (lane - 1)) & 1u));
    bool is_root = false;
    if (head) {
      const uint32_t root = CL.find(g);
      is_root = (root == g);
      if (!is_root) L[i] = root;
    }
    my_roots += __popc(__ballot_sync(0xffffffffu, is_root));
  }
  if (lane == 0) sh.warp_roots[wid] = my_roots;
  __syncthreads();
  if (wid == 0) {
    const int v = lane < nwarps ? sh.warp_roots[lane] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    sh.warp_base[lane] = incl - v;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane < cs) *cluster.map_shared_rank(&sh.x_roots[rank], lane) = total;
  }
  cluster.sync();  // root counts exchanged; nobody walks the forest any more
  int rank_base = 0, K = 0;
  for (int j = 0; j < cs; ++j) { if (j < rank) rank_base += sh.x_roots[j]; K += sh.x_roots[j]; }

  // ---- 7. number the roots 1..K in raster order; CTA 0 publishes K and initialises the box accumulators ----
  {
    int running = rank_base + sh.warp_base[wid];
    for (int s = s_begin; s < s_end; ++s) {
      const int y = s / segs, x = ((s - y * segs) << 5) + lane;
      const bool valid = x < W;
      const int i = y * W + x;
      const bool is_root = valid && (L[i] == base + static_cast<uint32_t>(i));
      const uint32_t rm = __ballot_sync(0xffffffffu, is_root);
      if (is_root) L[i] = kRankFlag | static_cast<uint32_t>(running + __popc(rm & ((1u << lane) - 1u)) + 1);
      running += __popc(rm);
    }
  }
  const int n_box = min(K, max_boxes);
  int32_t* mybox = boxes ? boxes + map * static_cast<long long>(max_boxes) * 5 : nullptr;
  if (rank == 0) {
    if (tid == 0 && counts) counts[map] = K;
    if (mybox)
      for (int k = tid; k < n_box; k += blockDim.x) {
        mybox[k * 5 + 0] = 0x7fffffff; mybox[k * 5 + 1] = 0x7fffffff;
        mybox[k * 5 + 2] = -1; mybox[k * 5 + 3] = -1; mybox[k * 5 + 4] = 0;
      }
    __threadfence();
  }
  cluster.sync();

  // ---- 8a. run heads fetch their component number (possibly from another CTA) and feed the boxes ----
  for (int s = wid; s < n_seg; s += nwarps) {
    const int y = s / segs, x = ((s - y * segs) << 5) + lane;
    const bool valid = x < W;
    const int i = y * W + x;
    const uint32_t v = valid ? L[i] : kBG;
    const bool fg = v != kBG;
    const uint32_t m = __ballot_sync(0xffffffffu, fg);
    const bool head = fg && (lane == 0 || !((m >> (lane - 1)) & 1u));
    if (head) {
      uint32_t id;
      if (v & kRankFlag) id = v & ~kRankFlag;
      else { id = *CL.slot(v) & ~kRankFlag; L[i] = kRankFlag | id; }
      if (mybox && static_cast<int>(id) <= n_box) {
        const uint32_t run_end = ~(m >> lane);  // first zero at/after this lane; a full segment from lane 0 has none
        const int len = run_end ? __ffs(run_end) - 1 : 32;
        int32_t* b = mybox + (id - 1) * 5;
        atomicMin(b + 0, x); atomicMin(b + 1, r0 + y);
        atomicMax(b + 2, x + len - 1); atomicMax(b + 3, r0 + y);
        atomicAdd(b + 4, len);
      }
    }
  }
  __syncthreads();

  // ---- 8b. every pixel reads its run head's number; labels leave the SM once, coalesced ----
  if (labels) {
    int32_t* out = labels + map * H * W + base;
    for (int s = wid; s < n_seg; s += nwarps) {
      const int y = s / segs, x = ((s - y * segs) << 5) + lane;
      if (x < W) {
        const int i = y * W + x;
        const uint32_t v = L[i];
        uint32_t id = 0;
        if (v != kBG) id = (v & kRankFlag) ? (v & ~kRankFlag) : (L[v - base] & ~kRankFlag);
        out[i] = static_cast<int32_t>(id);
      }
    }
  }
  __threadfence();
  cluster.sync();  // all box atomics are done; no CTA's shared memory is read after this point

  // ---- 9. (xmin, ymin, xmax, ymax, area) -> (x, y, w, h, area) ----
  if (rank == 0 && mybox) {
    for (int k = tid; k < n_box; k += blockDim.x) {
      const int x0 = __ldcg(mybox + k * 5 + 0), y0 = __ldcg(mybox + k * 5 + 1);
      const int x1 = __ldcg(mybox + k * 5 + 2), y1 = __ldcg(mybox + k * 5 + 3);
      mybox[k * 5 + 2] = x1 - x0 + 1;
      mybox[k * 5 + 3] = y1 - y0 + 1;
    }
  }
}

}  // namespace agenda

using namespace agenda;

extern "C" int agenda_ccl_bbox(const float* heat, float thr, int32_t* labels, int32_t* counts, int32_t* boxes,
                               int max_boxes, int n, int H, int W, void* stream) {
  if (!heat) return fail(AGENDA_ERR_NULL_POINTER, "ccl_bbox: heat is null");
  if (n < 0 || H <= 0 || W <= 0 || max_boxes < 0) return fail(AGENDA_ERR_BAD_SHAPE, "ccl_bbox: bad shape");
  if (boxes == nullptr) max_boxes = 0;
  if (static_cast<long long>(H) * W >= (1ll << 31)) return fail(AGENDA_ERR_UNSUPPORTED, "ccl_bbox: map too large");
  if (n == 0) return AGENDA_OK;
  const size_t budget = 200 * 1024;
  const size_t map_bytes = static_cast<size_t>(H) * W * 4;
  int cs = 1;
  while (cs < kMaxCluster && (static_cast<size_t>((H + cs - 1) / cs) * W * 4 > budget)) cs <<= 1;
  int R = (H + cs - 1) / cs;
  if (static_cast<size_t>(R) * W * 4 > budget)
    return fail(AGENDA_ERR_UNSUPPORTED, "ccl_bbox: %dx%d map (%zu B) does not fit a 16-CTA cluster's shared memory", H,
                W, map_bytes);
  cs = (H + R - 1) / R;  // drop empty trailing strips
  const size_t smem = (static_cast<size_t>(R) * W * 4 + 127) & ~static_cast<size_t>(127);
  const int strip_px = R * W;
  const int threads = strip_px >= 16384 ? 1024 : (strip_px >= 4096 ? 512 : 256);
  const int use_bulk = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(heat) & 15) == 0);

  AGENDA_CUDA(cudaFuncSetAttribute(ccl_bbox_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (cs > 8) AGENDA_CUDA(cudaFuncSetAttribute(ccl_bbox_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(static_cast<long long>(n) * cs));
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AGENDA_CUDA(cudaLaunchKernelEx(&cfg, ccl_bbox_kernel, heat, thr, labels, counts, boxes, max_boxes, H, W, R, use_bulk));
  return AGENDA_OK;
}
