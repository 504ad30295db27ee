// K6: threshold + 4-connected component labelling + bbox (spec: SURVEY.md §8 a9; not present in the reference).
//
// B200 design: ONE THREAD-BLOCK CLUSTER PER MAP, the whole fp32 map resident in distributed shared memory.
//   * each CTA of the cluster bulk-copies (TMA engine, cp.async.bulk + mbarrier) its strip of R rows into smem;
//     min/max are reduced CTA-locally and exchanged through DSMEM, so the map is read from HBM exactly ONCE even
//     though normalisation needs the global min/max before the first pixel can be thresholded;
//   * the normalise-and-compare `((h-min)/((max-min)+1e-8f)) > thr` is monotone in h, so one warp finds, by a
//     32-ary search over ordered float bit patterns (7 ballots), the smallest float that passes; every pixel then
//     needs a single compare and the result is bit-identical to evaluating the IEEE division per pixel;
//   * pixels are thresholded four at a time (float4) into a row-major BIT MASK; from here on the unit of work is a
//     32-pixel mask word, not a pixel: "pieces" (maximal runs of set bits inside one word) are enumerated with
//     carry-ripple bit tricks, each piece owns one union-find slot (16 slots per word, overlaid on the dead fp32
//     strip), pieces are merged with the piece(s) they touch in the row above and across word borders with
//     shared-memory atomicMin union-find, strips are stitched through DSMEM;
//   * slot ids increase in raster order of the piece's first pixel, union always keeps the smaller id, so a
//     component's root is its first pixel in raster order; roots are numbered with a block scan plus a DSMEM
//     exchange of per-strip root counts, which reproduces scipy.ndimage.label's numbering bit-exactly;
//   * labels leave the SM once (int4 per thread, coalesced; all-background quads take a fast path), boxes through
//     per-piece global atomics.
// Algorithmic HBM bytes per map: H*W*4 read + H*W*4 written (+ 20 B per box).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace agenda {

constexpr uint32_t kRankFlag = 0x80000000u;
constexpr int kMaxCluster = 16;
constexpr int kSlotsPerWord = 16;  // at most 16 pieces (alternating bits) in a 32-pixel word
constexpr int32_t kCtaOverflowFlag = -1;  // counts[map] value with which the one-CTA kernel hands a map to the cluster kernel
constexpr int kSmemBoxes = 256;           // boxes accumulated in shared memory by the one-CTA kernel

struct CclStatic {
  unsigned long long mbar;
  float red_min[32], red_max[32];
  int red_nan[32];
  float x_min[kMaxCluster], x_max[kMaxCluster];
  int x_nan[kMaxCluster];
  int x_roots[kMaxCluster];
  int warp_tot[32];
  float hstar;
  int mode;  // 0: compare against hstar, 1: nothing is foreground, 2: everything is foreground
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// order-preserving float <-> uint32 key
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// bits of the piece (maximal run of ones) of word m that starts at bit p (m has bit p set, bit p-1 clear or p == 0)
__device__ __forceinline__ uint32_t piece_from(uint32_t m, int p) {
  const uint32_t t = m + (1u << p);  // the carry ripples through the run
  return (m ^ t) & m;
}
// start bit of the piece of m containing bit q
__device__ __forceinline__ int piece_start(uint32_t m, int q) {
  const uint32_t zeros_below = ~m & ((1u << q) - 1u);
  return zeros_below ? 32 - __clz(zeros_below) : 0;
}
// index (0..15) of the piece of m containing bit q, pieces counted from bit 0
__device__ __forceinline__ int piece_index(uint32_t m, int q) {
  const uint32_t starts = m & ~(m << 1);
  return __popc(starts & (0xFFFFFFFFu >> (31 - q))) - 1;
}

struct Forest {
  cg::cluster_group cluster;
  uint32_t* slots;       // this CTA's union-find slots
  uint32_t strip_slots;  // slots per strip = R * words_per_row * 16
  uint32_t base;         // first slot id of this strip

  // find with path halving: re-pointing a node at its grandparent only ever moves it closer to the root, so it is
  // safe against concurrent atomicMin unions (every stored value stays an ancestor)
  __device__ __forceinline__ uint32_t find_local(uint32_t x) const {
    uint32_t p = slots[x - base];
    while (p != x) {
      const uint32_t g = slots[p - base];
      if (g != p) slots[x - base] = g;
      x = p; p = g;
    }
    return x;
  }
  __device__ __forceinline__ void unite_local(uint32_t a, uint32_t b) const {
    bool done;
    do {
      a = find_local(a);
      b = find_local(b);
      if (a < b) { const uint32_t old = atomicMin(&slots[b - base], a); done = (old == b); b = old; }
      else if (b < a) { const uint32_t old = atomicMin(&slots[a - base], b); done = (old == a); a = old; }
      else done = true;
    } while (!done);
  }
  __device__ __forceinline__ uint32_t* slot(uint32_t id) const {
    const uint32_t rk = id / strip_slots;
    return cluster.map_shared_rank(slots, rk) + (id - rk * strip_slots);
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

// Merge every piece of word `m` (slots first_slot + k) with the pieces of `up` (the word above, slots up_first + k)
// it overlaps.  kLocal: both rows live in this CTA.
template <bool kLocal>
__device__ __forceinline__ void merge_with_row_above(const Forest& F, uint32_t m, uint32_t up, uint32_t first_slot,
                                                     uint32_t up_first) {
  uint32_t starts = m & ~(m << 1);
  int k = 0;
  while (starts) {
    const int p = __ffs(starts) - 1;
    starts &= starts - 1;
    uint32_t ov = piece_from(m, p) & up;
    while (ov) {
      const int q = __ffs(ov) - 1;
      const int pu = piece_start(up, q);
      ov &= ~piece_from(up, pu);
      const uint32_t a = first_slot + k, b = up_first + piece_index(up, q);
      if (kLocal) F.unite_local(a, b); else F.unite(a, b);
    }
    ++k;
  }
}

// One map on one cluster.  `mbar_parity`: phase of the bulk-copy mbarrier (the kernel below may run several maps).
__device__ __forceinline__ void ccl_cluster_map(const float* __restrict__ heat, float thr, int32_t* __restrict__ labels,
                                                int32_t* __restrict__ counts, int32_t* __restrict__ boxes,
                                                int max_boxes, int H, int W, int R, int use_bulk, int strip_bytes,
                                                CclStatic& sh, unsigned char* dyn_smem, long long map,
                                                uint32_t mbar_parity) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5, nthreads = blockDim.x;

  const int r0 = rank * R;
  const int rows = min(R, H - r0);
  const int n_px = rows * W;
  const int wpr = (W + 31) >> 5;           // mask words per row
  const int n_words = R * wpr;             // per strip (rows beyond `rows` stay zero)
  const int strip_slots = n_words * kSlotsPerWord;

  float* Lf = reinterpret_cast<float*>(dyn_smem);                      // fp32 strip, dead after thresholding
  uint32_t* slots = reinterpret_cast<uint32_t*>(dyn_smem);             // union-find slots, overlaid on the strip
  uint32_t* bits = reinterpret_cast<uint32_t*>(dyn_smem + strip_bytes);  // [R][wpr] foreground mask
  const float* src = heat + map * H * W + static_cast<long long>(r0) * W;

  // ---- 1. strip -> shared memory (bulk async copy through the TMA engine when 16-B aligned) ----
  if (use_bulk) {
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
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&sh.mbar)), "r"(mbar_parity)
          : "memory");
    }
  } else {
    for (int i = tid; i < n_px; i += nthreads) Lf[i] = __ldg(src + i);
    __syncthreads();
  }

  // ---- 2. min / max: CTA reduce, then all-to-all through distributed shared memory ----
  {
    float lo = INFINITY, hi = -INFINITY;
    int nan = 0;
    if ((n_px & 3) == 0) {
      for (int i = tid; i < (n_px >> 2); i += nthreads) {
        const float4 v = reinterpret_cast<const float4*>(Lf)[i];
        lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
        hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
      }
    } else {
      for (int i = tid; i < n_px; i += nthreads) {
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

  // ---- 3. warp 0: smallest float h* with ((h*-min)/denom) > thr  (32-ary search over ordered bit patterns) ----
  if (wid == 0) {
    float mn = INFINITY, mx = -INFINITY;
    int any_nan = 0;
    for (int j = 0; j < cs; ++j) { mn = fminf(mn, sh.x_min[j]); mx = fmaxf(mx, sh.x_max[j]); any_nan |= sh.x_nan[j]; }
    const float denom = np_denominator(mn, mx);
    int mode = 0;
    float hstar = 0.f;
    if (any_nan || !(np_normalize(mx, mn, denom) > thr)) mode = 1;       // NaN map (numpy: all False) or max fails
    else if (np_normalize(mn, mn, denom) > thr) mode = 2;                // even the minimum passes
    else {
      // invariant: pass(lo) false, pass(hi) true; keys are monotone in the float order
      uint32_t lo = f2key(mn), hi = f2key(mx);
      while (hi - lo > 1u) {
        const uint32_t span = hi - lo;                       // >= 2
        const uint32_t step = span / 33u + 1u;
        uint32_t cand = lo + step * static_cast<uint32_t>(lane + 1);
        const bool valid = (cand - lo) < span;               // lo < cand < hi (no wrap: step*(33) <= span + 33)
        if (!valid) cand = hi;
        const bool pass = np_normalize(key2f(cand), mn, denom) > thr;
        const uint32_t pm = __ballot_sync(0xffffffffu, pass);  // monotone: 0..0 1..1 from some lane on
        if (pm == 0) {                                         // all 32 probes fail: the cut is above the last one
          lo = __shfl_sync(0xffffffffu, cand, 31);
        } else {
          const int first = __ffs(pm) - 1;
          const uint32_t new_hi = __shfl_sync(0xffffffffu, cand, first);
          const uint32_t new_lo = __shfl_sync(0xffffffffu, cand, first > 0 ? first - 1 : 0);
          hi = new_hi;
          if (first > 0) lo = new_lo;
        }
      }
      hstar = key2f(hi);
    }
    if (lane == 0) { sh.hstar = hstar; sh.mode = mode; }
  }
  __syncthreads();
  const float hstar = sh.hstar;
  const int mode = sh.mode;

  // ---- 4. threshold four pixels per thread -> bit mask words ----
  if ((W & 31) == 0) {
    const int n_quads = n_px >> 2;
    for (int i = tid; i < ((n_quads + 31) & ~31); i += nthreads) {
      uint32_t nib = 0;
      if (i < n_quads) {
        const float4 v = reinterpret_cast<const float4*>(Lf)[i];
        nib = (v.x >= hstar ? 1u : 0u) | (v.y >= hstar ? 2u : 0u) | (v.z >= hstar ? 4u : 0u) | (v.w >= hstar ? 8u : 0u);
        if (mode) nib = (mode == 2) ? 0xFu : 0u;
      }
      // 8 consecutive lanes hold one 32-pixel word
      const uint32_t word = __reduce_or_sync(0xFFu << (lane & 24), nib << ((lane & 7) * 4));
      if ((lane & 7) == 0 && i < n_quads) bits[i >> 3] = word;
    }
    for (int i = (n_px >> 5) + tid; i < n_words; i += nthreads) bits[i] = 0;  // rows past the end of the map
  } else {
    // generic width: one warp per (row, word)
    for (int s = wid; s < n_words; s += nwarps) {
      const int y = s / wpr, x = ((s - y * wpr) << 5) + lane;
      bool fg = false;
      if (y < rows && x < W) fg = mode ? (mode == 2) : (Lf[y * W + x] >= hstar);
      const uint32_t word = __ballot_sync(0xffffffffu, fg);
      if (lane == 0) bits[s] = word;
    }
  }
  __syncthreads();  // the fp32 strip is dead from here on: its memory becomes the union-find slots

  Forest F{cluster, slots, static_cast<uint32_t>(strip_slots), static_cast<uint32_t>(rank) * strip_slots};

  // ---- 5. one slot per piece, initialised to itself ----
  for (int w = tid; w < n_words; w += nthreads) {
    const uint32_t m = bits[w];
    const int np = __popc(m & ~(m << 1));
    for (int k = 0; k < np; ++k) slots[w * kSlotsPerWord + k] = F.base + w * kSlotsPerWord + k;
  }
  __syncthreads();

  // ---- 6. merge inside the strip: across word borders and with the row above ----
  for (int w = tid; w < n_words; w += nthreads) {
    const uint32_t m = bits[w];
    if (m == 0) continue;
    const int y = w / wpr, wx = w - y * wpr;
    const uint32_t first = F.base + w * kSlotsPerWord;
    if (wx > 0 && (m & 1u)) {
      const uint32_t left = bits[w - 1];
      if (left >> 31) F.unite_local(first, first - kSlotsPerWord + __popc(left & ~(left << 1)) - 1);
    }
    if (y > 0) {
      const uint32_t up = bits[w - wpr];
      if (up & m) merge_with_row_above<true>(F, m, up, first, first - wpr * kSlotsPerWord);
    }
  }
  cluster.sync();  // every strip is internally merged

  // ---- 7. stitch with the strip above through distributed shared memory ----
  if (rank > 0) {
    const uint32_t* up_bits = cluster.map_shared_rank(bits, rank - 1) + (R - 1) * wpr;
    for (int wx = tid; wx < wpr; wx += nthreads) {
      const uint32_t m = bits[wx];
      const uint32_t up = up_bits[wx];
      if (m & up)
        merge_with_row_above<false>(F, m, up, F.base + wx * kSlotsPerWord,
                                    F.base - wpr * kSlotsPerWord + wx * kSlotsPerWord);
    }
  }
  cluster.sync();  // the forest is final

  // ---- 8. flatten pieces to their root; count roots in raster order (thread t owns a contiguous word range) ----
  const int wpt = (n_words + nthreads - 1) / nthreads;
  const int w_begin = min(tid * wpt, n_words), w_end = min(w_begin + wpt, n_words);
  int my_roots = 0;
  for (int w = w_begin; w < w_end; ++w) {
    const uint32_t m = bits[w];
    const int np = __popc(m & ~(m << 1));
    for (int k = 0; k < np; ++k) {
      const uint32_t id = F.base + w * kSlotsPerWord + k;
      const uint32_t root = F.find(id);
      if (root == id) ++my_roots; else slots[w * kSlotsPerWord + k] = root;
    }
  }
  // exclusive scan of my_roots over the block (raster order == thread order)
  int incl = my_roots;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh.warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int v = lane < nwarps ? sh.warp_tot[lane] : 0;
    int winc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    sh.warp_tot[lane] = winc - v;  // exclusive prefix per warp
    const int total = __shfl_sync(0xffffffffu, winc, 31);
    if (lane < cs) *cluster.map_shared_rank(&sh.x_roots[rank], lane) = total;
  }
  cluster.sync();  // root counts exchanged; nobody walks the forest any more
  int rank_base = 0, K = 0;
  for (int j = 0; j < cs; ++j) { if (j < rank) rank_base += sh.x_roots[j]; K += sh.x_roots[j]; }

  // ---- 9. number the roots 1..K in raster order; CTA 0 publishes K and initialises the box accumulators ----
  {
    int next = rank_base + sh.warp_tot[wid] + (incl - my_roots) + 1;
    for (int w = w_begin; w < w_end; ++w) {
      const uint32_t m = bits[w];
      const int np = __popc(m & ~(m << 1));
      for (int k = 0; k < np; ++k)
        if (slots[w * kSlotsPerWord + k] == F.base + w * kSlotsPerWord + k)
          slots[w * kSlotsPerWord + k] = kRankFlag | static_cast<uint32_t>(next++);
    }
  }
  const int n_box = min(K, max_boxes);
  int32_t* mybox = boxes ? boxes + map * static_cast<long long>(max_boxes) * 5 : nullptr;
  if (rank == 0) {
    if (tid == 0 && counts) counts[map] = K;
    if (mybox)
      for (int k = tid; k < n_box; k += nthreads) {
        mybox[k * 5 + 0] = 0x7fffffff; mybox[k * 5 + 1] = 0x7fffffff;
        mybox[k * 5 + 2] = -1; mybox[k * 5 + 3] = -1; mybox[k * 5 + 4] = 0;
      }
    __threadfence();
  }
  cluster.sync();

  // ---- 10. every piece fetches its component number (possibly from another CTA) and feeds the boxes ----
  for (int w = tid; w < n_words; w += nthreads) {
    const uint32_t m = bits[w];
    if (m == 0) continue;
    const int y = w / wpr, wx = w - y * wpr;
    uint32_t starts = m & ~(m << 1);
    int k = 0;
    while (starts) {
      const int p = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint32_t v = slots[w * kSlotsPerWord + k];
      uint32_t id;
      if (v & kRankFlag) id = v & ~kRankFlag;
      else { id = *F.slot(v) & ~kRankFlag; slots[w * kSlotsPerWord + k] = kRankFlag | id; }
      if (mybox && static_cast<int>(id) <= n_box) {
        const int len = __popc(piece_from(m, p));
        const int x0 = (wx << 5) + p;
        int32_t* b = mybox + (id - 1) * 5;
        atomicMin(b + 0, x0); atomicMin(b + 1, r0 + y);
        atomicMax(b + 2, x0 + len - 1); atomicMax(b + 3, r0 + y);
        atomicAdd(b + 4, len);
      }
      ++k;
    }
  }
  __syncthreads();

  // ---- 11. labels leave the SM once: four pixels per thread, coalesced; all-background quads are free ----
  if (labels) {
    int32_t* out = labels + map * H * W + static_cast<long long>(r0) * W;
    if ((W & 31) == 0) {
      for (int i = tid; i < (n_px >> 2); i += nthreads) {
        const int w = i >> 3, q0 = (i & 7) << 2;
        const uint32_t m = bits[w];
        int4 o = make_int4(0, 0, 0, 0);
        if ((m >> q0) & 0xFu) {
          const uint32_t* ws = slots + w * kSlotsPerWord;
          if ((m >> q0) & 1u) o.x = static_cast<int>(ws[piece_index(m, q0)] & ~kRankFlag);
          if ((m >> q0) & 2u) o.y = static_cast<int>(ws[piece_index(m, q0 + 1)] & ~kRankFlag);
          if ((m >> q0) & 4u) o.z = static_cast<int>(ws[piece_index(m, q0 + 2)] & ~kRankFlag);
          if ((m >> q0) & 8u) o.w = static_cast<int>(ws[piece_index(m, q0 + 3)] & ~kRankFlag);
        }
        reinterpret_cast<int4*>(out)[i] = o;
      }
    } else {
      for (int s = wid; s < rows * wpr; s += nwarps) {
        const int y = s / wpr, x = ((s - y * wpr) << 5) + lane;
        if (x < W) {
          const uint32_t m = bits[s];
          out[y * W + x] = ((m >> lane) & 1u) ? static_cast<int>(slots[s * kSlotsPerWord + piece_index(m, lane)] & ~kRankFlag) : 0;
        }
      }
    }
  }
  __threadfence();
  cluster.sync();  // all box atomics are done; no CTA's shared memory is read after this point

  // ---- 12. (xmin, ymin, xmax, ymax, area) -> (x, y, w, h, area) ----
  if (rank == 0 && mybox) {
    for (int k = tid; k < n_box; k += nthreads) {
      const int x0 = __ldcg(mybox + k * 5 + 0), y0 = __ldcg(mybox + k * 5 + 1);
      const int x1 = __ldcg(mybox + k * 5 + 2), y1 = __ldcg(mybox + k * 5 + 3);
      mybox[k * 5 + 2] = x1 - x0 + 1;
      mybox[k * 5 + 3] = y1 - y0 + 1;
    }
  }
}

// grid = n_clusters * cluster_size.  Direct use: one cluster per map.  As the second-chance launch behind
// ccl_bbox_cta_kernel (`only_flagged`): a few persistent clusters walk over all maps and redo only those it flagged
// (piece table overflow) — a no-op launch of one cluster per map cost 70 us per 2048 maps, 5 % of the whole call.
// The flag of a map is read by every CTA of its cluster before any of them can overwrite it (first cluster.sync).
__global__ void __launch_bounds__(1024, 1)
ccl_bbox_kernel(const float* __restrict__ heat, float thr, int32_t* __restrict__ labels, int32_t* __restrict__ counts,
                int32_t* __restrict__ boxes, int max_boxes, int H, int W, int R, int use_bulk, int strip_bytes,
                int only_flagged, int n_maps) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ CclStatic sh;
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = static_cast<int>(cluster.num_blocks());
  const long long stride = gridDim.x / cs;
  if (only_flagged) {
    // common case: nothing to redo.  All threads look at this cluster's maps at once (a serial walk costs an L2 round
    // trip per map); the answer is the same in every CTA of the cluster because only its own rank 0 clears its flags.
    int any = 0;
    for (long long map = blockIdx.x / cs + threadIdx.x * stride; map < n_maps; map += blockDim.x * stride)
      any |= (__ldcg(counts + map) == kCtaOverflowFlag);
    if (!__syncthreads_or(any)) return;
  }
  if (use_bulk) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  uint32_t parity = 0;
  for (long long map = blockIdx.x / cs; map < n_maps; map += stride) {
    if (only_flagged && __ldcg(counts + map) != kCtaOverflowFlag) continue;
    ccl_cluster_map(heat, thr, labels, counts, boxes, max_boxes, H, W, R, use_bulk, strip_bytes, sh, dyn_smem, map, parity);
    if (use_bulk) parity ^= 1u;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// One CTA per map (maps whose bit mask fits one CTA's shared memory, W % 32 == 0).
//
// The cluster kernel above keeps the fp32 map resident in distributed shared memory, which caps the number of maps in
// flight at ~21 per GPU and makes every phase a cluster barrier (46 % of its warp samples, profiles/r01_ccl_*).  Here
// only the BIT MASK (1 bit per pixel) and a compact piece table live in shared memory:
//   pass 1  streams the map once for min / max / NaN (float4, 4 loads in flight per thread) and leaves a 16-bit
//           (min, max) key pair per 32-pixel word (the range table, see `wtab` in the kernel);
//   pass 2  fills the mask words from the table and re-reads ONLY the words whose range straddles the threshold (~3 % of
//           a config-5 map); without the table (cluster split, tiny piece tables, AGENDA_CCL_TAB=0) it re-reads the map in
//           REVERSE order — the tail of pass 1 is still in L2 — and thresholds straight into mask words;
//   pieces (runs of set bits inside a word) get compact ids by a block scan (raster order is preserved, so the
//   smallest id of a component is still its first pixel and scipy's numbering falls out of a second scan);
//   union-find, root numbering, boxes (shared-memory accumulators) and the label write need __syncthreads only.
// Extra HBM traffic: the straddling words (or, without the table, the part of pass 2 that misses L2).  Several CTAs per SM,
// no cluster, no DSMEM.
// A map with more pieces than `cap` is flagged in counts[] and redone by the cluster kernel (second launch).
// L2 eviction-priority hints: with the table everything streams with evict_first; without it pass 1 asks L2 to keep the
// tail of the map (evict_last) and pass 2 releases it (evict_first)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

// kCS > 1: the two streaming passes of a map are split over a cluster of kCS CTAs (min / max exchanged and the mask
// words delivered to the leader through distributed shared memory, two cluster barriers); the other CTAs then exit
// and the leader labels the map alone.  A quarter of a map is re-read ~4x sooner after it was first read, so pass 2
// finds it in L2 instead of going back to HBM.
template <int kThreads, int kCS>
__global__ void __launch_bounds__(kThreads)
ccl_bbox_cta_kernel(const float* __restrict__ heat, float thr, int32_t* __restrict__ labels, int32_t* __restrict__ counts,
                    int32_t* __restrict__ boxes, int max_boxes, int H, int W, int cap, int hints, int use_tab) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ float red_min[32], red_max[32];
  __shared__ int red_nan[32], warp_tot[32];
  __shared__ float x_min[kCS], x_max[kCS];
  __shared__ int x_nan[kCS];
  __shared__ float sh_hstar;
  __shared__ int sh_mode, sh_total;
  __shared__ int sbox[kSmemBoxes * 5];

  constexpr int kWarps = kThreads / 32;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long map = blockIdx.x / kCS;
  const int rank = (kCS > 1) ? static_cast<int>(cluster.block_rank()) : 0;
  const int n_px = H * W;
  const int wpr = W >> 5;
  const int n_words = H * wpr;
  uint32_t* bits = reinterpret_cast<uint32_t*>(dyn_smem);
  uint32_t* slots = bits + n_words;
  unsigned short* base = reinterpret_cast<unsigned short*>(slots + cap);  // first piece id of every word
  // Per-word range table (use_tab; kCS == 1 only): pass 1 leaves, for every 32-pixel word, the upper 16 bits of the ordered
  // keys of its minimum and maximum (min | max << 16).  Once h* is known a word whose maximum lies below it is all
  // background and a word whose minimum lies above it all foreground WITHOUT a second look at the pixels; only words whose
  // 16-bit range straddles h* (the contour of the components: ~3 % of a config-5 map) are read again.  The table overlays
  // the slots / base areas, which are not live before pass 2 has finished (needs 2 * cap >= n_words).
  uint32_t* wtab = slots;
  const bool tab = (kCS == 1) && use_tab;
  const float4* src4 = reinterpret_cast<const float4*>(heat + map * n_px);
  const int n4 = n_px >> 2;
  // this CTA's share of the two passes: whole mask words (8 float4 each)
  const int words_per_rank = (n_words + kCS - 1) / kCS;
  const int i_begin = min(n4, rank * words_per_rank * 8), i_end = min(n4, (rank + 1) * words_per_rank * 8);

  // ---- pass 1: min / max / NaN ----
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  // hints >= 2: only the LAST (hints) percent of this CTA's range — read last in pass 1 and first in pass 2 — is marked
  // evict_last; the rest streams with evict_first.  With ~300 maps in flight only a fraction of each can stay in the
  // 126 MB L2: marking everything evict_last marks nothing.
  const int keep_from = (hints >= 2) ? i_end - static_cast<int>(static_cast<long long>(i_end - i_begin) * hints / 100) : i_begin;
  auto ld1 = [&](int i) {
    return hints ? ldg_hint(src4 + i, (hints > 0 && i >= keep_from) ? pol_keep : pol_drop) : __ldg(src4 + i);
  };
  const unsigned group_mask = 0xFFu << (lane & 24);  // the 8 lanes that hold one mask word (8 float4)
  // word range: upper halves of the ordered keys of the float4's minimum / maximum, packed (min | max << 16) and reduced
  // over the word's 8 lanes with three xor-shuffles and per-halfword min / max (redux.sync with a partial member mask
  // compiles to a serialising loop: 27 % of the kernel's samples in the first version).  NaNs drop out of fminf / fmaxf;
  // a map with a NaN never consults the table.
  auto note_word = [&](int i, float vmin, float vmax) {
    const uint32_t bmn = __float_as_uint(vmin), bmx = __float_as_uint(vmax);
    const uint32_t kmn = bmn ^ (static_cast<uint32_t>(static_cast<int32_t>(bmn) >> 31) | 0x80000000u);  // == f2key
    const uint32_t kmx = bmx ^ (static_cast<uint32_t>(static_cast<int32_t>(bmx) >> 31) | 0x80000000u);
    uint32_t p = __byte_perm(kmn, kmx, 0x7632);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const uint32_t q = __shfl_xor_sync(group_mask, p, o);
      p = (__vmaxu2(p, q) & 0xFFFF0000u) | (__vminu2(p, q) & 0x0000FFFFu);
    }
    if ((lane & 7) == 0) wtab[i >> 3] = p;
  };
  auto min4 = [](const float4& v) { return fminf(fminf(v.x, v.y), fminf(v.z, v.w)); };
  auto max4 = [](const float4& v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); };
  auto ld2 = [&](int i) { return hints ? ldg_hint(src4 + i, pol_drop) : __ldg(src4 + i); };
  {
    float lo = INFINITY, hi = -INFINITY;
    int nan = 0;
    int i = i_begin + tid;
    for (; i + 3 * kThreads < i_end; i += 4 * kThreads) {
      const float4 a = ld1(i), b = ld1(i + kThreads), c = ld1(i + 2 * kThreads), d = ld1(i + 3 * kThreads);
      const float na = min4(a), nb = min4(b), nc = min4(c), nd = min4(d);
      const float xa = max4(a), xb = max4(b), xc = max4(c), xd = max4(d);
      lo = fminf(fminf(lo, fminf(na, nb)), fminf(nc, nd));
      hi = fmaxf(fmaxf(hi, fmaxf(xa, xb)), fmaxf(xc, xd));
      nan |= (a.x != a.x) | (a.y != a.y) | (a.z != a.z) | (a.w != a.w) | (b.x != b.x) | (b.y != b.y) | (b.z != b.z) |
             (b.w != b.w) | (c.x != c.x) | (c.y != c.y) | (c.z != c.z) | (c.w != c.w) | (d.x != d.x) | (d.y != d.y) |
             (d.z != d.z) | (d.w != d.w);
      // (i_begin, i_end and kThreads are multiples of 8: the 8 lanes of a word enter and leave these loops together)
      if (tab) { note_word(i, na, xa); note_word(i + kThreads, nb, xb); note_word(i + 2 * kThreads, nc, xc); note_word(i + 3 * kThreads, nd, xd); }
    }
    for (; i < i_end; i += kThreads) {
      const float4 a = ld1(i);
      const float na = min4(a), xa = max4(a);
      lo = fminf(lo, na);
      hi = fmaxf(hi, xa);
      nan |= (a.x != a.x) | (a.y != a.y) | (a.z != a.z) | (a.w != a.w);
      if (tab) note_word(i, na, xa);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) { red_min[wid] = lo; red_max[wid] = hi; red_nan[wid] = nan; }
  }
  __syncthreads();
  if (kCS > 1) {  // CTA partials -> every CTA of the cluster
    if (wid == 0) {
      float mn = lane < kWarps ? red_min[lane] : INFINITY;
      float mx = lane < kWarps ? red_max[lane] : -INFINITY;
      int any_nan = lane < kWarps ? red_nan[lane] : 0;
      mn = warp_min(mn); mx = warp_max(mx);
      any_nan = __any_sync(0xffffffffu, any_nan);
      if (lane < kCS) {
        *cluster.map_shared_rank(&x_min[rank], lane) = mn;
        *cluster.map_shared_rank(&x_max[rank], lane) = mx;
        *cluster.map_shared_rank(&x_nan[rank], lane) = any_nan;
      }
    }
    cluster.sync();
  }

  // ---- warp 0: smallest float h* with ((h*-min)/denom) > thr (32-ary search over ordered bit patterns) ----
  if (wid == 0) {
    float mn, mx;
    int any_nan;
    if (kCS > 1) {
      mn = lane < kCS ? x_min[lane] : INFINITY;
      mx = lane < kCS ? x_max[lane] : -INFINITY;
      any_nan = lane < kCS ? x_nan[lane] : 0;
    } else {
      mn = lane < kWarps ? red_min[lane] : INFINITY;
      mx = lane < kWarps ? red_max[lane] : -INFINITY;
      any_nan = lane < kWarps ? red_nan[lane] : 0;
    }
    mn = warp_min(mn); mx = warp_max(mx);
    any_nan = __any_sync(0xffffffffu, any_nan);
    const float denom = np_denominator(mn, mx);
    int mode = 0;
    float hstar = 0.f;
    if (any_nan || !(np_normalize(mx, mn, denom) > thr)) mode = 1;       // NaN map (numpy: all False) or max fails
    else if (np_normalize(mn, mn, denom) > thr) mode = 2;                // even the minimum passes
    else {
      uint32_t lo = f2key(mn), hi = f2key(mx);                            // pass(lo) false, pass(hi) true
      while (hi - lo > 1u) {
        const uint32_t span = hi - lo;
        const uint32_t step = span / 33u + 1u;
        uint32_t cand = lo + step * static_cast<uint32_t>(lane + 1);
        const bool valid = (cand - lo) < span;
        if (!valid) cand = hi;
        const bool pass = np_normalize(key2f(cand), mn, denom) > thr;
        const uint32_t pm = __ballot_sync(0xffffffffu, pass);
        if (pm == 0) {
          lo = __shfl_sync(0xffffffffu, cand, 31);
        } else {
          const int first = __ffs(pm) - 1;
          const uint32_t new_hi = __shfl_sync(0xffffffffu, cand, first);
          const uint32_t new_lo = __shfl_sync(0xffffffffu, cand, first > 0 ? first - 1 : 0);
          hi = new_hi;
          if (first > 0) lo = new_lo;
        }
      }
      hstar = key2f(hi);
    }
    if (lane == 0) { sh_hstar = hstar; sh_mode = mode; }
  }
  for (int k = tid; k < kSmemBoxes * 5; k += kThreads) {
    const int f = k % 5;
    sbox[k] = (f < 2) ? 0x7fffffff : (f < 4 ? -1 : 0);
  }
  __syncthreads();
  const float hstar = sh_hstar;
  const int mode = sh_mode;

  // ---- pass 2 (reverse order: the end of the range is the most recently used part of L2): threshold -> mask words,
  //      written into the LEADER's mask (its own shared memory when kCS == 1) ----
  uint32_t* lead_bits = (kCS > 1) ? cluster.map_shared_rank(bits, 0) : bits;
  if (mode == 0 && tab) {
    // key(a) < key(h) implies a < h (or a == -0, h == +0, which the search cannot produce: both zeros normalise alike),
    // and comparing the upper halves of two keys is conservative, so equal upper halves count as "straddles"
    const uint32_t ks16 = f2key(hstar) >> 16;
    const int span = i_end - i_begin;
    const int iters = (span + kThreads - 1) / kThreads;
    for (int it = 0; it < iters; ++it) {
      const int i = i_begin + it * kThreads + tid;
      if (i < i_end) {  // (uniform over the 8 lanes of a word)
        const uint32_t t = wtab[i >> 3];
        uint32_t word;
        if ((t >> 16) < ks16) word = 0u;
        else if ((t & 0xFFFFu) > ks16) word = 0xFFFFFFFFu;
        else {
          const float4 v = ld2(i);
          const uint32_t nib = (v.x >= hstar ? 1u : 0u) | (v.y >= hstar ? 2u : 0u) | (v.z >= hstar ? 4u : 0u) | (v.w >= hstar ? 8u : 0u);
          word = nib << ((lane & 7) * 4);
          word |= __shfl_xor_sync(group_mask, word, 1);
          word |= __shfl_xor_sync(group_mask, word, 2);
          word |= __shfl_xor_sync(group_mask, word, 4);
        }
        if ((lane & 7) == 0) bits[i >> 3] = word;
      }
    }
  } else if (mode == 0) {
    const int span = i_end - i_begin;
    const int iters = (span + kThreads - 1) / kThreads;
    for (int it = iters - 1; it >= 0; --it) {
      const int i = i_begin + it * kThreads + tid;
      uint32_t nib = 0;
      if (i < i_end) {
        const float4 v = ld2(i);
        nib = (v.x >= hstar ? 1u : 0u) | (v.y >= hstar ? 2u : 0u) | (v.z >= hstar ? 4u : 0u) | (v.w >= hstar ? 8u : 0u);
      }
      const uint32_t word = __reduce_or_sync(0xFFu << (lane & 24), nib << ((lane & 7) * 4));  // 8 lanes = one word
      if ((lane & 7) == 0 && i < i_end) lead_bits[i >> 3] = word;
    }
  } else {
    const uint32_t fill = (mode == 2) ? 0xFFFFFFFFu : 0u;
    for (int w = (i_begin >> 3) + tid; w < (i_end >> 3); w += kThreads) lead_bits[w] = fill;
  }
  if (kCS > 1) {
    cluster.sync();          // the leader's mask is complete; nobody touches another CTA's shared memory after this
    if (rank != 0) return;
  } else {
    __syncthreads();
  }

  // ---- compact piece ids: thread t owns a contiguous word range; exclusive block scan of the piece counts ----
  const int wpt = (n_words + kThreads - 1) / kThreads;
  const int w_begin = min(tid * wpt, n_words), w_end = min(w_begin + wpt, n_words);
  auto block_exclusive_scan = [&](int mine, int& total) -> int {  // returns the exclusive prefix of `mine`
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();  // warp_tot / sh_total from a previous scan have been consumed
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int v = lane < kWarps ? warp_tot[lane] : 0;
      int winc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_tot[lane] = winc - v;
      if (lane == 31) sh_total = winc;
    }
    __syncthreads();
    total = sh_total;
    return warp_tot[wid] + incl - mine;
  };
  int my_pieces = 0;
  for (int w = w_begin; w < w_end; ++w) {
    const uint32_t m = bits[w];
    my_pieces += __popc(m & ~(m << 1));
  }
  int P;
  int next_id = block_exclusive_scan(my_pieces, P);
  if (P > cap) {  // piece table overflow: hand the map to the cluster kernel
    if (tid == 0) counts[map] = kCtaOverflowFlag;
    return;
  }
  for (int w = w_begin; w < w_end; ++w) {
    const uint32_t m = bits[w];
    const int np = __popc(m & ~(m << 1));
    base[w] = static_cast<unsigned short>(next_id);
    for (int k = 0; k < np; ++k) slots[next_id + k] = next_id + k;
    next_id += np;
  }
  __syncthreads();

  // ---- union-find: merge across word borders and with the row above ----
  Forest F{cg::this_cluster(), slots, 0u, 0u};
  for (int w = tid; w < n_words; w += kThreads) {
    const uint32_t m = bits[w];
    if (m == 0) continue;
    const int y = w / wpr, wx = w - y * wpr;
    const uint32_t first = base[w];
    if (wx > 0 && (m & 1u)) {
      const uint32_t left = bits[w - 1];
      if (left >> 31) F.unite_local(first, base[w - 1] + __popc(left & ~(left << 1)) - 1);
    }
    if (y > 0) {
      const uint32_t up = bits[w - wpr];
      if (up & m) merge_with_row_above<true>(F, m, up, first, base[w - wpr]);
    }
  }
  __syncthreads();

  // ---- flatten; count and number the roots in raster order ----
  int my_roots = 0;
  for (int w = w_begin; w < w_end; ++w) {
    const uint32_t m = bits[w];
    const int np = __popc(m & ~(m << 1));
    const uint32_t id0 = base[w];
    for (int k = 0; k < np; ++k) {
      const uint32_t id = id0 + k;
      uint32_t x = id, p = slots[x];
      while (p != x) { x = p; p = slots[x]; }
      if (x == id) ++my_roots; else slots[id] = x;  // non-roots point straight at their root (roots never change now)
    }
  }
  int K;
  int next_label = block_exclusive_scan(my_roots, K) + 1;
  // (the scan's barriers order every flatten write before the numbering below; a non-root's target is a root, and roots
  //  are only rewritten by their owner after this point, flagged, so readers in the next phase can tell them apart)
  for (int w = w_begin; w < w_end; ++w) {
    const uint32_t m = bits[w];
    const int np = __popc(m & ~(m << 1));
    const uint32_t id0 = base[w];
    for (int k = 0; k < np; ++k)
      if (slots[id0 + k] == id0 + k) slots[id0 + k] = kRankFlag | static_cast<uint32_t>(next_label++);
  }
  __syncthreads();

  // ---- every piece resolves its label and feeds the boxes ----
  const int n_box = min(K, max_boxes);
  int32_t* mybox = boxes ? boxes + map * static_cast<long long>(max_boxes) * 5 : nullptr;
  if (tid == 0) counts[map] = K;
  if (mybox)
    for (int k = kSmemBoxes + tid; k < n_box; k += kThreads) {  // boxes beyond the shared-memory window: global atomics
      mybox[k * 5 + 0] = 0x7fffffff; mybox[k * 5 + 1] = 0x7fffffff;
      mybox[k * 5 + 2] = -1; mybox[k * 5 + 3] = -1; mybox[k * 5 + 4] = 0;
    }
  if (n_box > kSmemBoxes) { __threadfence_block(); __syncthreads(); }
  for (int w = tid; w < n_words; w += kThreads) {
    const uint32_t m = bits[w];
    if (m == 0) continue;
    const int y = w / wpr, wx = w - y * wpr;
    const uint32_t id0 = base[w];
    uint32_t starts = m & ~(m << 1);
    int k = 0;
    while (starts) {
      const int p = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint32_t v = slots[id0 + k];
      uint32_t label;
      if (v & kRankFlag) label = v & ~kRankFlag;
      else { label = slots[v] & ~kRankFlag; slots[id0 + k] = kRankFlag | label; }
      if (mybox && static_cast<int>(label) <= n_box) {
        const int len = __popc(piece_from(m, p));
        const int x0 = (wx << 5) + p;
        if (label <= kSmemBoxes) {
          int* b = sbox + (label - 1) * 5;
          atomicMin(b + 0, x0); atomicMin(b + 1, y); atomicMax(b + 2, x0 + len - 1); atomicMax(b + 3, y); atomicAdd(b + 4, len);
        } else {
          int32_t* b = mybox + (label - 1) * 5;
          atomicMin(b + 0, x0); atomicMin(b + 1, y); atomicMax(b + 2, x0 + len - 1); atomicMax(b + 3, y); atomicAdd(b + 4, len);
        }
      }
      ++k;
    }
  }
  __syncthreads();

  // ---- labels leave the SM once: four pixels per thread, coalesced; all-background quads are free ----
  if (labels) {
    int4* out4 = reinterpret_cast<int4*>(labels + map * n_px);
    for (int i = tid; i < n4; i += kThreads) {
      const int w = i >> 3, q0 = (i & 7) << 2;
      const uint32_t m = bits[w];
      int4 o = make_int4(0, 0, 0, 0);
      if ((m >> q0) & 0xFu) {
        const uint32_t* ws = slots + base[w];
        if ((m >> q0) & 1u) o.x = static_cast<int>(ws[piece_index(m, q0)] & ~kRankFlag);
        if ((m >> q0) & 2u) o.y = static_cast<int>(ws[piece_index(m, q0 + 1)] & ~kRankFlag);
        if ((m >> q0) & 4u) o.z = static_cast<int>(ws[piece_index(m, q0 + 2)] & ~kRankFlag);
        if ((m >> q0) & 8u) o.w = static_cast<int>(ws[piece_index(m, q0 + 3)] & ~kRankFlag);
      }
      __stcs(out4 + i, o);  // streaming store: labels are never re-read, keep L2 for the heat maps
    }
  }
  // ---- boxes: (xmin, ymin, xmax, ymax, area) -> (x, y, w, h, area) ----
  if (mybox) {
    for (int k = tid; k < min(n_box, kSmemBoxes); k += kThreads) {
      const int x0 = sbox[k * 5 + 0], y0 = sbox[k * 5 + 1], x1 = sbox[k * 5 + 2], y1 = sbox[k * 5 + 3];
      mybox[k * 5 + 0] = x0; mybox[k * 5 + 1] = y0; mybox[k * 5 + 2] = x1 - x0 + 1; mybox[k * 5 + 3] = y1 - y0 + 1;
      mybox[k * 5 + 4] = sbox[k * 5 + 4];
    }
    if (n_box > kSmemBoxes) {
      __threadfence();  // (global atomics of this CTA are complete: __syncthreads above)
      for (int k = kSmemBoxes + tid; k < n_box; k += kThreads) {
        const int x0 = __ldcg(mybox + k * 5 + 0), y0 = __ldcg(mybox + k * 5 + 1);
        const int x1 = __ldcg(mybox + k * 5 + 2), y1 = __ldcg(mybox + k * 5 + 3);
        mybox[k * 5 + 2] = x1 - x0 + 1;
        mybox[k * 5 + 3] = y1 - y0 + 1;
      }
    }
  }
}

}  // namespace agenda

using namespace agenda;

extern "C" int agenda_ccl_bbox(const float* heat, float thr, int32_t* labels, int32_t* counts, int32_t* boxes,
                               int max_boxes, int n, int H, int W, void* stream) {
  if (n == 0) return AGENDA_OK;  // (an empty batch has no buffers to point at)
  if (!heat) return fail(AGENDA_ERR_NULL_POINTER, "ccl_bbox: heat is null");
  if (n < 0 || H <= 0 || W <= 0 || max_boxes < 0) return fail(AGENDA_ERR_BAD_SHAPE, "ccl_bbox: bad shape");
  if (boxes == nullptr) max_boxes = 0;
  if (static_cast<long long>(H) * W >= (1ll << 26)) return fail(AGENDA_ERR_UNSUPPORTED, "ccl_bbox: map too large");
  if (n == 0) return AGENDA_OK;
  // Shared-memory budget per CTA.  100 KB lets two CTAs (of different clusters) share an SM, so one cluster's
  // barrier waits are covered by another's work; AGENDA_CCL_SMEM_KB overrides it for experiments.
  size_t budget = 100 * 1024;
  if (const char* e = knob("AGENDA_CCL_SMEM_KB")) { const long kb = atol(e); if (kb >= 8 && kb <= 220) budget = static_cast<size_t>(kb) * 1024; }
  const size_t map_bytes = static_cast<size_t>(H) * W * 4;
  const int wpr = (W + 31) / 32;
  // per strip of R rows: fp32 strip R*W*4 (re-used for the union-find slots: R*wpr*64 B) + bit mask R*wpr*4
  auto strip_smem = [&](int R) {
    const size_t strip = std::max(static_cast<size_t>(R) * W * 4, static_cast<size_t>(R) * wpr * kSlotsPerWord * 4);
    return ((strip + 127) & ~static_cast<size_t>(127)) + static_cast<size_t>(R) * wpr * 4;
  };
  int cs = 1;
  while (cs < kMaxCluster && strip_smem((H + cs - 1) / cs) > budget) cs <<= 1;
  const int R = (H + cs - 1) / cs;
  if (strip_smem(R) > 200 * 1024)
    return fail(AGENDA_ERR_UNSUPPORTED, "ccl_bbox: %dx%d map (%zu B) does not fit a 16-CTA cluster's shared memory", H,
                W, map_bytes);
  cs = (H + R - 1) / R;  // drop empty trailing strips
  const size_t smem = strip_smem(R);
  const int strip_bytes = static_cast<int>(smem - static_cast<size_t>(R) * wpr * 4);  // offset of the bit mask
  const int strip_px = R * W;
  // measured on 512x512 maps (16 strips of 32 rows): 256 threads 2.86 ms / 512: 3.36 / 1024: 6.0 per 2048 maps
  int threads = strip_px >= 65536 ? 1024 : (strip_px >= 32768 ? 512 : 256);
  if (const char* e = knob("AGENDA_CCL_THREADS")) { const int t = atoi(e); if (t >= 64 && t <= 1024 && t % 32 == 0) threads = t; }
  const int use_bulk = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(heat) & 15) == 0);

  // ---- one CTA per map when the bit mask fits (see ccl_bbox_cta_kernel); the cluster kernel then only redoes flagged maps ----
  int only_flagged = 0;
  {
    const long long n_px = static_cast<long long>(H) * W;
    const long long n_words = n_px / 32;
    bool use_cta = counts != nullptr && (W & 31) == 0 && n_words <= 12288 && (reinterpret_cast<uintptr_t>(heat) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(labels) & 15) == 0;
    if (const char* e = knob("AGENDA_CCL_CTA")) use_cta = use_cta && atoi(e) != 0;
    if (use_cta) {
      const size_t fixed = static_cast<size_t>(n_words) * 6;  // mask words (4 B) + first piece id per word (2 B)
      size_t target = std::max<size_t>(72 * 1024, fixed + 16 * 1024);  // 72 KB: three CTAs per SM
      if (const char* e = knob("AGENDA_CCL_CTA_SMEM_KB")) { const long kb = atol(e); if (kb >= 4 && kb <= 200) target = static_cast<size_t>(kb) * 1024; }
      long long cap = std::min<long long>({16 * n_words, 65535ll, static_cast<long long>((target - std::min(target, fixed)) / 4)});
      cap &= ~1ll;  // keeps the u16 table 4-byte aligned behind the slots
      if (cap >= 64 || cap >= 16 * n_words - 1) {
        const size_t smem_cta = static_cast<size_t>(n_words) * 4 + static_cast<size_t>(cap) * 4 + static_cast<size_t>(n_words) * 2 + 16;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
#define AGENDA_CCL_CTA(T)                                                                                              \
  do {                                                                                                                 \
    if (cta_cluster == 2) {                                                                                            \
      AGENDA_CUDA(cudaFuncSetAttribute(ccl_bbox_cta_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                       static_cast<int>(smem_cta)));                                                   \
      cudaLaunchConfig_t c4 = {};                                                                                      \
      c4.gridDim = dim3(static_cast<unsigned>(static_cast<long long>(n) * 2));                                         \
      c4.blockDim = dim3(T);                                                                                           \
      c4.dynamicSmemBytes = smem_cta;                                                                                  \
      c4.stream = st;                                                                                                  \
      cudaLaunchAttribute a4[1];                                                                                       \
      a4[0].id = cudaLaunchAttributeClusterDimension;                                                                  \
      a4[0].val.clusterDim.x = 2; a4[0].val.clusterDim.y = 1; a4[0].val.clusterDim.z = 1;                              \
      c4.attrs = a4; c4.numAttrs = 1;                                                                                  \
      AGENDA_CUDA(cudaLaunchKernelEx(&c4, ccl_bbox_cta_kernel<T, 2>, heat, thr, labels, counts, boxes, max_boxes, H,   \
                                     W, static_cast<int>(cap), hints, use_tab));                                             \
    } else if (cta_cluster == 4) {                                                                                     \
      AGENDA_CUDA(cudaFuncSetAttribute(ccl_bbox_cta_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                       static_cast<int>(smem_cta)));                                                   \
      cudaLaunchConfig_t c4 = {};                                                                                      \
      c4.gridDim = dim3(static_cast<unsigned>(static_cast<long long>(n) * 4));                                         \
      c4.blockDim = dim3(T);                                                                                           \
      c4.dynamicSmemBytes = smem_cta;                                                                                  \
      c4.stream = st;                                                                                                  \
      cudaLaunchAttribute a4[1];                                                                                       \
      a4[0].id = cudaLaunchAttributeClusterDimension;                                                                  \
      a4[0].val.clusterDim.x = 4; a4[0].val.clusterDim.y = 1; a4[0].val.clusterDim.z = 1;                              \
      c4.attrs = a4; c4.numAttrs = 1;                                                                                  \
      AGENDA_CUDA(cudaLaunchKernelEx(&c4, ccl_bbox_cta_kernel<T, 4>, heat, thr, labels, counts, boxes, max_boxes, H,   \
                                     W, static_cast<int>(cap), hints, use_tab));                                             \
    } else {                                                                                                           \
      AGENDA_CUDA(cudaFuncSetAttribute(ccl_bbox_cta_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                       static_cast<int>(smem_cta)));                                                   \
      ccl_bbox_cta_kernel<T, 1><<<n, T, smem_cta, st>>>(heat, thr, labels, counts, boxes, max_boxes, H, W,             \
                                                        static_cast<int>(cap), hints, use_tab);                             \
    }                                                                                                                  \
  } while (0)
        int cta_cluster = 1;  // 4: split the two streaming passes of a map over a 4-CTA cluster
        if (const char* e = knob("AGENDA_CCL_CTA_CLUSTER")) { const int c = atoi(e); cta_cluster = (c == 4 || c == 2) ? c : 1; }
        if (n_words < 64) cta_cluster = 1;
        // 25: the last quarter of the map (read last in pass 1, first in pass 2) is marked evict_last, the rest evict_first.
        // Measured on 2048 config-5 maps: no hints 1.196 ms, everything evict_last 1.184, 15 % 1.165, 25 % 1.154, 35 % 1.159,
        // 50 % 1.173 (profiles/r02_ccl_l2_keep_fraction_sweep.txt): ~300 MB of maps are in flight against 126 MB of L2
        int hints = 25;
        if (const char* e = knob("AGENDA_CCL_HINTS")) hints = atoi(e);
        // per-word range table (see the kernel): pass 2 re-reads only the words whose range straddles the threshold
        int use_tab = (cta_cluster == 1 && 2 * cap >= n_words) ? 1 : 0;
        if (const char* e = knob("AGENDA_CCL_TAB")) use_tab = use_tab && atoi(e) != 0;
        // nothing to keep in L2 for a sparse second pass: everything streams with evict_first (0.945 -> 0.895 ms per 2048 maps)
        if (use_tab && !knob("AGENDA_CCL_HINTS")) hints = -1;
        int cta_threads = n_px >= 65536 ? 1024 : (n_px >= 16384 ? 256 : 128);
        if (const char* e = knob("AGENDA_CCL_CTA_THREADS")) { const int t = atoi(e); if (t == 128 || t == 256 || t == 512 || t == 1024) cta_threads = t; }
        if (cta_threads == 1024) AGENDA_CCL_CTA(1024);
        else if (cta_threads == 512) AGENDA_CCL_CTA(512);
        else if (cta_threads == 256) AGENDA_CCL_CTA(256);
        else AGENDA_CCL_CTA(128);
#undef AGENDA_CCL_CTA
        AGENDA_LAUNCH_CHECK("ccl_bbox_cta_kernel");
        if (cap >= 16 * n_words) return AGENDA_OK;  // the piece table cannot overflow: no second launch
        only_flagged = 1;
      }
    }
  }

  AGENDA_CUDA(cudaFuncSetAttribute(ccl_bbox_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (cs > 8) AGENDA_CUDA(cudaFuncSetAttribute(ccl_bbox_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  const long long n_clusters = only_flagged ? std::min<long long>(n, 8) : n;  // persistent walkers for the fallback
  cfg.gridDim = dim3(static_cast<unsigned>(n_clusters * cs));
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
  AGENDA_CUDA(cudaLaunchKernelEx(&cfg, ccl_bbox_kernel, heat, thr, labels, counts, boxes, max_boxes, H, W, R, use_bulk,
                                 strip_bytes, only_flagged, n));
  return AGENDA_OK;
}
