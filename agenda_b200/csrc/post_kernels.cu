// K4 + K5: heat-map normalise -> u8 -> PIL-exact bicubic resize (data_generation/data_generation.py:82-85)
// and invert + channel stack (data_generation/postprocess_heatmap.py:44-46).  Byte/integer results are
// bit-exact with numpy + PIL.  Everything between the fp32 read and the u8 write lives in registers / shared memory, so
// HBM traffic is the algorithmic minimum (read Hi*Wi*4 per map, write Ho*Wo per plane); the fused kernel runs persistent
// CTAs that build Pillow's coefficient tables once and walk the images.
#include <algorithm>

#include "common.cuh"

namespace agenda {

constexpr int kPilPrecisionBits = 32 - 8 - 2;  // Pillow Resample.c PRECISION_BITS
constexpr int kPostThreads = 256;

// Pillow bicubic_filter (a = -0.5) in double, evaluated without FMA contraction so the fixed-point
// coefficients equal the ones Pillow's C code computes on the host.
__device__ __forceinline__ double pil_bicubic(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0)
    return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

__host__ __device__ inline int pil_ksize(int in_size, int out_size) {
  double scale = static_cast<double>(static_cast<float>(in_size)) / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  double support = 2.0 * filterscale;
  int c = static_cast<int>(support);
  if (static_cast<double>(c) < support) ++c;  // ceil
  return c * 2 + 1;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one axis.  bounds[2*xx] = xmin, [2*xx+1] = count;
// kk[xx*ksize + x] = fixed-point weight.  Thread-parallel over output positions.
__device__ void pil_coeffs(int in_size, int out_size, int ksize, int* bounds, int* kk) {
  const double scale = __ddiv_rn(static_cast<double>(static_cast<float>(in_size)), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
    const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);
    int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double w = pil_bicubic(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
      ww = __dadd_rn(ww, w);
    }
    for (int x = 0; x < ksize; ++x) {
      int q = 0;
      if (x < xmax) {
        double w = pil_bicubic(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        const double s = __dmul_rn(w, static_cast<double>(1 << kPilPrecisionBits));
        q = (w < 0) ? static_cast<int>(__dadd_rn(-0.5, s)) : static_cast<int>(__dadd_rn(0.5, s));
      }
      kk[xx * ksize + x] = q;
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

__device__ __forceinline__ uint8_t pil_clip8(int v) {
  v >>= kPilPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int j) { return __byte_perm(w, 0u, 0x4440u | j); }

// in [Hi,Wi] u8 (smem) -> tmp [Hi,Wo] (smem) -> out [Ho,Wo] (smem or global), Pillow's horizontal-then-vertical
// order.  Coefficient tables must already be in smem.  Caller syncs before.
// Fast forms for Pillow's five-tap support (every upscale; the tables are zero-filled past a window's real length, so
// the five taps are always taken — a tap with weight 0 may read a few bytes past its row or plane, still inside the
// CTA's carve-up, see resize_smem_bytes):
//   horizontal: a thread owns one output COLUMN (window start and the five weights in registers) and walks the rows;
//   vertical:   a thread owns four adjacent output columns (one 32-bit load per tap) and walks the output rows.
// No integer division per output, ~3x fewer instructions than the generic (i / Wo, variable-length window) loops below.
__device__ void pil_resize_smem(const uint8_t* in, uint8_t* tmp, uint8_t* out, int Hi, int Wi, int Ho, int Wo,
                                const int* bx, const int* kx, int ksx, const int* by, const int* ky, int ksy) {
  const uint8_t* hsrc = in;
  const int T = blockDim.x, tid = threadIdx.x;
  constexpr int kHalf = 1 << (kPilPrecisionBits - 1);
  const bool roomy = Hi >= 8 && Ho >= 8;  // (zero-weight taps past the end stay inside the next buffer of the carve-up)
  if (Wo != Wi) {
    if (ksx == 5 && Wo <= T && roomy) {
      const int groups = T / Wo, g = tid / Wo, xx = tid - g * Wo;
      if (g < groups) {
        const int* k = kx + xx * 5;
        const int w0 = k[0], w1 = k[1], w2 = k[2], w3 = k[3], w4 = k[4];
        const uint8_t* r = in + bx[2 * xx] + g * Wi;
        uint8_t* o = tmp + g * Wo + xx;
        const int rs = groups * Wi, os = groups * Wo;
        for (int y = g; y < Hi; y += groups, r += rs, o += os) {
          const int acc = kHalf + r[0] * w0 + r[1] * w1 + r[2] * w2 + r[3] * w3 + r[4] * w4;
          *o = pil_clip8(acc);
        }
      }
    } else {
      for (int i = tid; i < Hi * Wo; i += T) {
        const int y = i / Wo, xx = i - y * Wo;
        const int x0 = bx[2 * xx], n = bx[2 * xx + 1];
        int acc = kHalf;
        for (int x = 0; x < n; ++x) acc += static_cast<int>(in[y * Wi + x0 + x]) * kx[xx * ksx + x];
        tmp[i] = pil_clip8(acc);
      }
    }
    hsrc = tmp;
    __syncthreads();
  }
  if (Ho != Hi) {
    const int q = Wo >> 2;  // four-column groups per row
    if (ksy == 5 && (Wo & 3) == 0 && q <= T && roomy && (reinterpret_cast<uintptr_t>(hsrc) & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
      const int groups = T / q, g = tid / q, x4 = tid - g * q;
      if (g < groups) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(hsrc) + x4;
        uint32_t* dst = reinterpret_cast<uint32_t*>(out) + x4;
        for (int yy = g; yy < Ho; yy += groups) {
          const int* k = ky + yy * 5;
          const uint32_t* r = src + by[2 * yy] * q;
          int a0 = kHalf, a1 = kHalf, a2 = kHalf, a3 = kHalf;
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            const uint32_t w = r[t * q];
            const int c = k[t];
            a0 += static_cast<int>(byte_of(w, 0)) * c;
            a1 += static_cast<int>(byte_of(w, 1)) * c;
            a2 += static_cast<int>(byte_of(w, 2)) * c;
            a3 += static_cast<int>(byte_of(w, 3)) * c;
          }
          dst[yy * q] = static_cast<uint32_t>(pil_clip8(a0)) | (static_cast<uint32_t>(pil_clip8(a1)) << 8) |
                        (static_cast<uint32_t>(pil_clip8(a2)) << 16) | (static_cast<uint32_t>(pil_clip8(a3)) << 24);
        }
      }
    } else {
      for (int i = tid; i < Ho * Wo; i += T) {
        const int yy = i / Wo, x = i - yy * Wo;
        const int y0 = by[2 * yy], n = by[2 * yy + 1];
        int acc = kHalf;
        for (int y = 0; y < n; ++y) acc += static_cast<int>(hsrc[(y0 + y) * Wo + x]) * ky[yy * ksy + y];
        out[i] = pil_clip8(acc);
      }
    }
  } else {
    for (int i = tid; i < Ho * Wo; i += T) out[i] = hsrc[i];
  }
  __syncthreads();
}

// Block-wide reduction of per-thread (min, max, saw-a-NaN); result broadcast through smem scratch red[2*32].
__device__ void block_minmax_reduce(float lo, float hi, bool nan, float* red, float& mn, float& mx) {
  // warp reduce (NaN-propagating like numpy's min / max: fminf drops NaN, so carry a flag)
  unsigned any_nan = __ballot_sync(0xffffffffu, nan);
  lo = warp_min(lo); hi = warp_max(hi);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { red[wid] = any_nan ? NAN : lo; red[32 + wid] = any_nan ? NAN : hi; }
  __syncthreads();
  if (wid == 0) {
    float a = lane < nw ? red[lane] : INFINITY, b = lane < nw ? red[32 + lane] : -INFINITY;
    unsigned nn = __ballot_sync(0xffffffffu, a != a);
    a = warp_min(a); b = warp_max(b);
    if (lane == 0) { red[0] = nn ? NAN : a; red[32] = nn ? NAN : b; }
  }
  __syncthreads();
  mn = red[0]; mx = red[32];
  __syncthreads();
}

// Block-wide min/max of an fp32 map (global).
__device__ void block_minmax(const float* __restrict__ h, int n, float* red, float& mn, float& mx) {
  float lo = INFINITY, hi = -INFINITY;
  bool nan = false;
  if ((reinterpret_cast<uintptr_t>(h) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(h) + i);
      lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
      hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
      nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float v = __ldg(h + i);
      lo = fminf(lo, v); hi = fmaxf(hi, v); nan |= (v != v);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = __ldg(h + i);
      lo = fminf(lo, v); hi = fmaxf(hi, v); nan |= (v != v);
    }
  }
  block_minmax_reduce(lo, hi, nan, red, mn, mx);
}

__device__ __forceinline__ uint8_t quantize_u8(float h, float mn, float denom) {
  // (h-min)/(max-min+1e-8)*255 in fp32, then numpy astype(uint8) == C truncation (data_generation.py:82,84)
  const float f = __fmul_rn(np_normalize(h, mn, denom), 255.0f);
  return static_cast<uint8_t>(static_cast<int>(f));
}

__global__ void __launch_bounds__(kPostThreads) heat_normalize_u8_kernel(const float* __restrict__ heat,
                                                                        uint8_t* __restrict__ out, int hw) {
  __shared__ float red[64];
  const float* h = heat + static_cast<long long>(blockIdx.x) * hw;
  uint8_t* o = out + static_cast<long long>(blockIdx.x) * hw;
  float mn, mx;
  block_minmax(h, hw, red, mn, mx);
  const float denom = np_denominator(mn, mx);
  for (int i = threadIdx.x; i < hw; i += blockDim.x) o[i] = quantize_u8(__ldg(h + i), mn, denom);
}

// smem carve-up shared by the resize kernels
struct ResizeSmem {
  int *bx, *kx, *by, *ky;
  uint8_t *in, *tmp, *out;
};
__host__ __device__ inline size_t resize_smem_bytes(int Hi, int Wi, int Ho, int Wo, int out_planes) {
  const int ksx = pil_ksize(Wi, Wo), ksy = pil_ksize(Hi, Ho);
  size_t ints = static_cast<size_t>(2 * Wo + Wo * ksx + 2 * Ho + Ho * ksy);
  auto al = [](size_t v) { return (v + 15) & ~static_cast<size_t>(15); };
  return al(ints * 4) + al(static_cast<size_t>(Hi) * Wi) + al(static_cast<size_t>(Hi) * Wo) +
         al(static_cast<size_t>(out_planes) * Ho * Wo);
}
__device__ inline ResizeSmem carve(unsigned char* base, int Hi, int Wi, int Ho, int Wo, int ksx, int ksy) {
  ResizeSmem s;
  int* p = reinterpret_cast<int*>(base);
  s.bx = p; p += 2 * Wo;
  s.kx = p; p += Wo * ksx;
  s.by = p; p += 2 * Ho;
  s.ky = p; p += Ho * ksy;
  auto al = [](uintptr_t v) { return (v + 15) & ~static_cast<uintptr_t>(15); };
  s.in = reinterpret_cast<uint8_t*>(al(reinterpret_cast<uintptr_t>(p)));
  s.tmp = reinterpret_cast<uint8_t*>(al(reinterpret_cast<uintptr_t>(s.in + Hi * Wi)));
  s.out = reinterpret_cast<uint8_t*>(al(reinterpret_cast<uintptr_t>(s.tmp + Hi * Wo)));
  return s;
}

// MODE 0: u8 in -> resized u8 out.   MODE 1: fp32 heat in -> normalise -> u8 -> resized u8 out.
template <int MODE>
__global__ void __launch_bounds__(kPostThreads) resize_kernel(const void* __restrict__ in_, uint8_t* __restrict__ out,
                                                             int Hi, int Wi, int Ho, int Wo) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ float red[64];
  const int ksx = pil_ksize(Wi, Wo), ksy = pil_ksize(Hi, Ho);
  ResizeSmem s = carve(smem, Hi, Wi, Ho, Wo, ksx, ksy);
  pil_coeffs(Wi, Wo, ksx, s.bx, s.kx);
  pil_coeffs(Hi, Ho, ksy, s.by, s.ky);
  const int hw = Hi * Wi;
  if (MODE == 0) {
    const uint8_t* src = static_cast<const uint8_t*>(in_) + static_cast<long long>(blockIdx.x) * hw;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) s.in[i] = src[i];
  } else {
    const float* h = static_cast<const float*>(in_) + static_cast<long long>(blockIdx.x) * hw;
    float mn, mx;
    block_minmax(h, hw, red, mn, mx);
    const float denom = np_denominator(mn, mx);
    for (int i = threadIdx.x; i < hw; i += blockDim.x) s.in[i] = quantize_u8(__ldg(h + i), mn, denom);
  }
  __syncthreads();
  pil_resize_smem(s.in, s.tmp, s.out, Hi, Wi, Ho, Wo, s.bx, s.kx, ksx, s.by, s.ky, ksy);
  uint8_t* o = out + static_cast<long long>(blockIdx.x) * Ho * Wo;
  const int n_out = Ho * Wo;
  if ((n_out & 3) == 0) {
    for (int i = threadIdx.x; i < (n_out >> 2); i += blockDim.x)
      reinterpret_cast<uint32_t*>(o)[i] = reinterpret_cast<const uint32_t*>(s.out)[i];
  } else {
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) o[i] = s.out[i];
  }
}

// a8: one thread -> 4 pixels: 3 x 4 B in, 12 B interleaved + 4 B inverted out.
__global__ void __launch_bounds__(256) stack_kernel(const uint8_t* __restrict__ obj, const uint8_t* __restrict__ fg,
                                                    const uint8_t* __restrict__ bg, uint8_t* __restrict__ stack,
                                                    uint8_t* __restrict__ inv, long long n_px, int vec_ok) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = vec_ok ? (n_px >> 2) : 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(obj) + i);
    const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(fg) + i);
    const uint32_t c = ~__ldg(reinterpret_cast<const uint32_t*>(bg) + i);  // 255 - x per byte
    // bytes: a0 b0 c0 a1 | b1 c1 a2 b2 | c2 a3 b3 c3
    uint32_t o0 = (a & 0xffu) | ((b & 0xffu) << 8) | ((c & 0xffu) << 16) | (((a >> 8) & 0xffu) << 24);
    uint32_t o1 = ((b >> 8) & 0xffu) | (((c >> 8) & 0xffu) << 8) | (((a >> 16) & 0xffu) << 16) | (((b >> 16) & 0xffu) << 24);
    uint32_t o2 = ((c >> 16) & 0xffu) | (((a >> 24) & 0xffu) << 8) | (((b >> 24) & 0xffu) << 16) | (((c >> 24) & 0xffu) << 24);
    uint32_t* s = reinterpret_cast<uint32_t*>(stack) + i * 3;
    s[0] = o0; s[1] = o1; s[2] = o2;
    if (inv) reinterpret_cast<uint32_t*>(inv)[i] = c;
  }
  for (long long i = (n4 << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_px; i += stride) {
    const uint8_t c = static_cast<uint8_t>(255 - bg[i]);
    stack[i * 3] = obj[i]; stack[i * 3 + 1] = fg[i]; stack[i * 3 + 2] = c;
    if (inv) inv[i] = c;
  }
}

// a7 + a8 fused: 3 maps per image (object, fg token, bg token).  PERSISTENT CTAs: the Pillow coefficient tables (double
// precision, ~2 x 112 windows) are built once per CTA and serve every image the CTA walks (img = blockIdx.x, += gridDim.x)
// instead of once per image; a map small enough (<= 16 floats per thread) is read from HBM exactly once, into registers
// — min / max and the u8 quantisation both work from there — and the NEXT map's loads are issued before the current
// map's resize, so their latency hides behind the integer filter.  Outputs leave as 32-bit words.
__global__ void __launch_bounds__(kPostThreads, 4) postprocess_stack_kernel(const float* __restrict__ heat,
                                                                        uint8_t* __restrict__ planes,
                                                                        uint8_t* __restrict__ stack,
                                                                        uint8_t* __restrict__ inv, int n, int Hi, int Wi,
                                                                        int Ho, int Wo) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ float red[64];
  const int ksx = pil_ksize(Wi, Wo), ksy = pil_ksize(Hi, Ho);
  ResizeSmem s = carve(smem, Hi, Wi, Ho, Wo, ksx, ksy);
  pil_coeffs(Wi, Wo, ksx, s.bx, s.kx);
  pil_coeffs(Hi, Ho, ksy, s.by, s.ky);
  const int hw = Hi * Wi, n_out = Ho * Wo, hw4 = hw >> 2;
  const int tid = threadIdx.x, T = kPostThreads;
  const bool in_regs = (hw & 3) == 0 && hw4 <= 4 * T && (reinterpret_cast<uintptr_t>(heat) & 15) == 0;
  const bool words_ok = (n_out & 3) == 0 && (reinterpret_cast<uintptr_t>(stack) & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(planes) & 3) == 0 && (reinterpret_cast<uintptr_t>(inv) & 3) == 0;
  float4 nxt[4];
  auto fetch = [&](long long m) {  // map m (= img * 3 + t) -> registers
    const float4* h4 = reinterpret_cast<const float4*>(heat + m * hw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + j * T;
      if (i < hw4) nxt[j] = __ldg(h4 + i);
    }
  };
  if (in_regs && blockIdx.x < n) fetch(static_cast<long long>(blockIdx.x) * 3);
  for (long long img = blockIdx.x; img < n; img += gridDim.x) {
    for (int t = 0; t < 3; ++t) {
      const float* h = heat + (img * 3 + t) * hw;
      float mn, mx;
      if (in_regs) {
        float4 cur[4];
        float lo = INFINITY, hi = -INFINITY;
        bool nan = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cur[j] = nxt[j];
          if (tid + j * T < hw4) {
            const float4 v = cur[j];
            lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
            hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
            nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
          }
        }
        // the next map of this CTA (next plane, or plane 0 of its next image): in flight during this map's resize
        const long long m_next = (t < 2) ? img * 3 + t + 1 : (img + gridDim.x) * 3;
        if (m_next < static_cast<long long>(n) * 3) fetch(m_next);
        block_minmax_reduce(lo, hi, nan, red, mn, mx);
        const float denom = np_denominator(mn, mx);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = tid + j * T;
          if (i < hw4) {
            const float4 v = cur[j];
            reinterpret_cast<uint32_t*>(s.in)[i] =
                static_cast<uint32_t>(quantize_u8(v.x, mn, denom)) | (static_cast<uint32_t>(quantize_u8(v.y, mn, denom)) << 8) |
                (static_cast<uint32_t>(quantize_u8(v.z, mn, denom)) << 16) | (static_cast<uint32_t>(quantize_u8(v.w, mn, denom)) << 24);
          }
        }
      } else {
        block_minmax(h, hw, red, mn, mx);
        const float denom = np_denominator(mn, mx);
        for (int i = tid; i < hw; i += T) s.in[i] = quantize_u8(__ldg(h + i), mn, denom);
      }
      __syncthreads();
      pil_resize_smem(s.in, s.tmp, s.out + t * n_out, Hi, Wi, Ho, Wo, s.bx, s.kx, ksx, s.by, s.ky, ksy);
    }
    // planes (as generated, before inversion), then stack with inverted bg
    if (words_ok) {
      const int n4 = n_out >> 2;
      const uint32_t* a4 = reinterpret_cast<const uint32_t*>(s.out);
      const uint32_t *b4 = a4 + n4, *c4 = b4 + n4;
      if (planes) {
        uint32_t* p = reinterpret_cast<uint32_t*>(planes + img * 3 * n_out);
        for (int i = tid; i < 3 * n4; i += T) p[i] = a4[i];
      }
      uint32_t* st = reinterpret_cast<uint32_t*>(stack + img * 3 * n_out);
      uint32_t* iv = inv ? reinterpret_cast<uint32_t*>(inv + img * n_out) : nullptr;
      for (int i = tid; i < n4; i += T) {
        const uint32_t a = a4[i], b = b4[i], c = ~c4[i];  // 255 - x per byte
        // bytes: a0 b0 c0 a1 | b1 c1 a2 b2 | c2 a3 b3 c3
        st[3 * i + 0] = __byte_perm(__byte_perm(a, b, 0x1040), c, 0x3410);
        st[3 * i + 1] = __byte_perm(__byte_perm(a, b, 0x6205), c, 0x3250);
        st[3 * i + 2] = __byte_perm(__byte_perm(a, b, 0x0730), c, 0x7216);
        if (iv) iv[i] = c;
      }
    } else {
      if (planes) {
        uint8_t* p = planes + img * 3 * n_out;
        for (int i = tid; i < 3 * n_out; i += T) p[i] = s.out[i];
      }
      uint8_t* st = stack + img * 3 * n_out;
      for (int i = tid; i < 3 * n_out; i += T) {
        const int px = i / 3, c = i - px * 3;
        const uint8_t v = s.out[c * n_out + px];
        st[i] = (c == 2) ? static_cast<uint8_t>(255 - v) : v;
      }
      if (inv) {
        uint8_t* iv = inv + img * n_out;
        for (int i = tid; i < n_out; i += T) iv[i] = static_cast<uint8_t>(255 - s.out[2 * n_out + i]);
      }
    }
    __syncthreads();  // s.out is rewritten by the next image
  }
}

static int resize_shape_check(const char* who, int n, int Hi, int Wi, int Ho, int Wo, int planes, size_t* smem) {
  if (n < 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "%s: bad shape", who);
  *smem = resize_smem_bytes(Hi, Wi, Ho, Wo, planes);
  if (*smem > 200 * 1024)
    return fail(AGENDA_ERR_UNSUPPORTED, "%s: %dx%d -> %dx%d needs %zu B of shared memory (> 200 KB)", who, Hi, Wi,
                Ho, Wo, *smem);
  return AGENDA_OK;
}

}  // namespace agenda

using namespace agenda;

extern "C" int agenda_heat_normalize_u8(const float* heat, uint8_t* out, int n, int hw, void* stream) {
  if (n == 0) return AGENDA_OK;  // (an empty batch has no buffers to point at)
  if (!heat || !out) return fail(AGENDA_ERR_NULL_POINTER, "heat_normalize_u8: null pointer");
  if (n < 0 || hw <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "heat_normalize_u8: bad shape");
  if (n == 0) return AGENDA_OK;
  heat_normalize_u8_kernel<<<n, kPostThreads, 0, static_cast<cudaStream_t>(stream)>>>(heat, out, hw);
  AGENDA_LAUNCH_CHECK("heat_normalize_u8_kernel");
  return AGENDA_OK;
}

template <int MODE>
static int launch_resize(const char* who, const void* in, uint8_t* out, int n, int Hi, int Wi, int Ho, int Wo,
                         void* stream) {
  if (n == 0) return AGENDA_OK;
  if (!in || !out) return fail(AGENDA_ERR_NULL_POINTER, "%s: null pointer", who);
  size_t smem = 0;
  int rc = resize_shape_check(who, n, Hi, Wi, Ho, Wo, 1, &smem);
  if (rc != AGENDA_OK) return rc;
  if (n == 0) return AGENDA_OK;
  AGENDA_DYN_SMEM(resize_kernel<MODE>, smem);
  resize_kernel<MODE><<<n, kPostThreads, smem, static_cast<cudaStream_t>(stream)>>>(in, out, Hi, Wi, Ho, Wo);
  AGENDA_LAUNCH_CHECK(who);
  return AGENDA_OK;
}

extern "C" int agenda_resize_bicubic_u8(const uint8_t* in, uint8_t* out, int n, int Hi, int Wi, int Ho, int Wo,
                                        void* stream) {
  return launch_resize<0>("resize_bicubic_u8", in, out, n, Hi, Wi, Ho, Wo, stream);
}

extern "C" int agenda_heat_to_u8_image(const float* heat, uint8_t* out, int n, int Hi, int Wi, int Ho, int Wo,
                                       void* stream) {
  return launch_resize<1>("heat_to_u8_image", heat, out, n, Hi, Wi, Ho, Wo, stream);
}

extern "C" int agenda_stack_heatmaps_u8(const uint8_t* obj, const uint8_t* fg, const uint8_t* bg, uint8_t* stack,
                                        uint8_t* inv, int n, int H, int W, void* stream) {
  if (n == 0) return AGENDA_OK;
  if (!obj || !fg || !bg || !stack) return fail(AGENDA_ERR_NULL_POINTER, "stack_heatmaps_u8: null pointer");
  if (n < 0 || H <= 0 || W <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "stack_heatmaps_u8: bad shape");
  if (n == 0) return AGENDA_OK;
  const long long n_px = static_cast<long long>(n) * H * W;
  const uintptr_t al = reinterpret_cast<uintptr_t>(obj) | reinterpret_cast<uintptr_t>(fg) |
                       reinterpret_cast<uintptr_t>(bg) | reinterpret_cast<uintptr_t>(stack) |
                       reinterpret_cast<uintptr_t>(inv);
  const int vec_ok = (al & 3) == 0;
  long long blocks = ((n_px >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  stack_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(obj, fg, bg, stack, inv,
                                                                                             n_px, vec_ok);
  AGENDA_LAUNCH_CHECK("stack_kernel");
  return AGENDA_OK;
}

extern "C" int agenda_heat_postprocess_stack(const float* heat, uint8_t* planes, uint8_t* stack, uint8_t* inv, int n,
                                             int Hi, int Wi, int Ho, int Wo, void* stream) {
  if (n == 0) return AGENDA_OK;
  if (!heat || !stack) return fail(AGENDA_ERR_NULL_POINTER, "heat_postprocess_stack: null pointer");
  size_t smem = 0;
  int rc = resize_shape_check("heat_postprocess_stack", n, Hi, Wi, Ho, Wo, 3, &smem);
  if (rc != AGENDA_OK) return rc;
  if (n == 0) return AGENDA_OK;
  AGENDA_DYN_SMEM(postprocess_stack_kernel, smem);
  // persistent CTAs: as many as fit the GPU at once (shared memory / 2048 threads per SM), each walking n / grid images
  const int per_sm = std::max(1, std::min(2048 / kPostThreads, static_cast<int>((220 * 1024) / (smem + 1024))));
  const int grid = std::min(n, num_sms() * per_sm);
  postprocess_stack_kernel<<<grid, kPostThreads, smem, static_cast<cudaStream_t>(stream)>>>(heat, planes, stack, inv, n,
                                                                                           Hi, Wi, Ho, Wo);
  AGENDA_LAUNCH_CHECK("postprocess_stack_kernel");
  return AGENDA_OK;
}
