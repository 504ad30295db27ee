// K3: streaming form of compute_global_heat_map (reference data_generation/hook.py:59-81).
//   acc += clamp(bicubic_upsample(map), min=0)   per (layer, step) map;   out = acc / count at the end.
// HBM-bound: per plane read h*w*4 B (stays in L1/L2 for the 16 taps) and read-modify-write L*L*4 B twice.
#include <algorithm>

#include "common.cuh"

namespace agenda {

// torch get_cubic_upsample_coefficients (A = -0.75) in fp32, taps for source offsets -1,0,+1,+2.
__device__ __forceinline__ void cubic_taps(float t, float w[4]) {
  const float A = -0.75f;
  float x = t + 1.0f;
  w[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
  w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  float u = 1.0f - t;
  w[2] = ((A + 2.0f) * u - (A + 3.0f)) * u * u + 1.0f;
  x = u + 1.0f;
  w[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

// area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=true) + border-replicated indices
__device__ __forceinline__ void cubic_src(int dst, float scale, int n_in, int idx[4], float w[4]) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  float fl = floorf(src);
  cubic_taps(src - fl, w);
  int i0 = static_cast<int>(fl);
#pragma unroll
  for (int j = 0; j < 4; ++j) idx[j] = min(max(i0 - 1 + j, 0), n_in - 1);
}

// One thread -> 4 consecutive output pixels of one row (float4 read-modify-write of acc, coalesced).
// G > 1 (DAAM-style aggregation): accumulator plane p = b'*T + t receives the G source planes (b'*G + g)*T + t, each
// upsampled and clamped on its own, summed in the fixed order g = 0..G-1 (deterministic).
__global__ void __launch_bounds__(256) heat_upsample_accum_kernel(const float* __restrict__ maps,
                                                                  float* __restrict__ acc, int n_planes,
                                                                  int h, int w, int L, int T, int G) {
  const int quads_per_row = L >> 2;
  const long long total = static_cast<long long>(n_planes) * L * quads_per_row;
  const float sy = static_cast<float>(h) / static_cast<float>(L);
  const float sx = static_cast<float>(w) / static_cast<float>(L);
#pragma unroll 2
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int qx = static_cast<int>(i % quads_per_row);
    const long long r = i / quads_per_row;
    const int y = static_cast<int>(r % L);
    const long long plane = r / L;
    float4* dst = reinterpret_cast<float4*>(acc + (plane * L + y) * L) + qx;
    float4 a = *dst;
    const int plane_i = static_cast<int>(plane);  // (n_planes is an int)
    const int bp = plane_i / T, tt = plane_i - bp * T;
    for (int g = 0; g < G; ++g) {
    const float* __restrict__ src = maps + (static_cast<long long>(bp * G + g) * T + tt) * h * w;
    float v[4];
    if (h == L && w == L) {  // scale 1: torch returns the input bit-exactly
      const float4 s = *reinterpret_cast<const float4*>(src + static_cast<long long>(y) * w + qx * 4);
      v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w;
    } else {
      int iy[4]; float wy[4];
      cubic_src(y, sy, h, iy, wy);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        int ix[4]; float wx[4];
        cubic_src(qx * 4 + p, sx, w, ix, wx);
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* row = src + iy[j] * w;
          // horizontal taps first, then the 4 rows (torch CPU kernel order)
          float rsum = __ldg(row + ix[0]) * wx[0] + __ldg(row + ix[1]) * wx[1] + __ldg(row + ix[2]) * wx[2] +
                       __ldg(row + ix[3]) * wx[3];
          o += rsum * wy[j];
        }
        v[p] = o;
      }
    }
    a.x += fmaxf(v[0], 0.f);  // .clamp_(min=0), hook.py:72
    a.y += fmaxf(v[1], 0.f);
    a.z += fmaxf(v[2], 0.f);
    a.w += fmaxf(v[3], 0.f);
    }
    *dst = a;
  }
}

// Tiled, separable form for source planes that fit shared memory (h*w <= 4096: every SD-1.x / SD-2.1 layer below the
// latent).  One CTA per accumulator plane: the tap tables (identical for every plane) and the source plane are staged in
// shared memory; pass 1 forms the horizontal 4-tap sums tmp[r][x] for all h source rows, pass 2 combines 4 of those rows
// per output row and adds into the accumulator with float4 read-modify-writes.  Same operations in the same order as
// the streaming kernel above (horizontal taps first, then the 4 rows), so results are bit-identical, with ~10x fewer
// loads per output.  HBM traffic per plane = h*w*4 read + L*L*8 read-modify-write, the algorithmic minimum.
// blockDim.x = L * k (k >= 1), so a thread's output column x = tid % L is fixed in pass 1.
__global__ void __launch_bounds__(256) heat_upsample_accum_tiled_kernel(const float* __restrict__ maps,
                                                                        float* __restrict__ acc, int n_planes, int h,
                                                                        int w, int L, int T, int G) {
  extern __shared__ __align__(16) float up_smem[];
  float* s_src = up_smem;                                   // [h*w]
  float* s_tmp = s_src + ((h * w + 3) & ~3);                // [h][L] horizontal sums
  float* s_wy = s_tmp + h * L;                              // [L][4]
  int* s_iy = reinterpret_cast<int*>(s_wy + 4 * L);         // [L][4] (row offsets into s_tmp)
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float sy = static_cast<float>(h) / static_cast<float>(L);
  const float sx = static_cast<float>(w) / static_cast<float>(L);
  for (int y = tid; y < L; y += nthr) {
    int idx[4]; float wgt[4];
    cubic_src(y, sy, h, idx, wgt);
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_iy[4 * y + j] = idx[j] * L; s_wy[4 * y + j] = wgt[j]; }
  }
  const int x = tid % L, r0 = tid / L, r_step = nthr / L;
  int ix[4]; float wx[4];
  cubic_src(x, sx, w, ix, wx);
  const int quads_per_row = L >> 2;
  const int n_quads = L * quads_per_row;
  // Fast form (one source plane per accumulator plane, at most four output quads and four source floats per thread — the
  // 64 x 64 latent of every SD pipeline): persistent CTAs; the accumulator quads of THIS plane are requested before its
  // source is even staged, and the NEXT plane's source floats are fetched into registers during the two filter passes, so
  // neither HBM round trip (16 KB of accumulator, 4 KB of source per plane) sits on the CTA's critical path.
  if (G == 1 && n_quads <= 4 * nthr && h * w <= 4 * nthr) {
    float nsrc[4];
    auto fetch_src = [&](int plane) {
      const float* __restrict__ src = maps + static_cast<long long>(plane) * h * w;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = tid + k * nthr;
        nsrc[k] = (plane < n_planes && i < h * w) ? __ldg(src + i) : 0.f;
      }
    };
    fetch_src(blockIdx.x);
    for (int plane = blockIdx.x; plane < n_planes; plane += gridDim.x) {
      float4* dst_plane = reinterpret_cast<float4*>(acc + static_cast<long long>(plane) * L * L);
      float4 a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = tid + k * nthr;
        if (q < n_quads) a[k] = dst_plane[q];
      }
      __syncthreads();  // the previous plane's s_src / s_tmp readers are done (and the tables are written)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = tid + k * nthr;
        if (i < h * w) s_src[i] = nsrc[k];
      }
      fetch_src(plane + gridDim.x);
      __syncthreads();
      for (int r = r0; r < h; r += r_step) {
        const float* row = s_src + r * w;
        s_tmp[r * L + x] = row[ix[0]] * wx[0] + row[ix[1]] * wx[1] + row[ix[2]] * wx[2] + row[ix[3]] * wx[3];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = tid + k * nthr;
        if (q < n_quads) {
          const int y = q / quads_per_row, qx = q - y * quads_per_row;
          const int4 iy = *reinterpret_cast<const int4*>(s_iy + 4 * y);
          const float4 wy = *reinterpret_cast<const float4*>(s_wy + 4 * y);
          const float4 t0 = *reinterpret_cast<const float4*>(s_tmp + iy.x + 4 * qx);
          const float4 t1 = *reinterpret_cast<const float4*>(s_tmp + iy.y + 4 * qx);
          const float4 t2 = *reinterpret_cast<const float4*>(s_tmp + iy.z + 4 * qx);
          const float4 t3 = *reinterpret_cast<const float4*>(s_tmp + iy.w + 4 * qx);
          // o = 0; o += rsum_j * wy_j for j = 0..3 (the streaming kernel's order)
          const float o0 = ((0.f + t0.x * wy.x) + t1.x * wy.y) + t2.x * wy.z + t3.x * wy.w;
          const float o1 = ((0.f + t0.y * wy.x) + t1.y * wy.y) + t2.y * wy.z + t3.y * wy.w;
          const float o2 = ((0.f + t0.z * wy.x) + t1.z * wy.y) + t2.z * wy.z + t3.z * wy.w;
          const float o3 = ((0.f + t0.w * wy.x) + t1.w * wy.y) + t2.w * wy.z + t3.w * wy.w;
          float4 v = a[k];
          v.x += fmaxf(o0, 0.f);
          v.y += fmaxf(o1, 0.f);
          v.z += fmaxf(o2, 0.f);
          v.w += fmaxf(o3, 0.f);
          dst_plane[q] = v;
        }
      }
    }
    return;
  }
  for (int plane = blockIdx.x; plane < n_planes; plane += gridDim.x) {  // tables are shared by every plane of the CTA
  const int bp = plane / T, tt = plane - bp * T;
  float4* dst_plane = reinterpret_cast<float4*>(acc + static_cast<long long>(plane) * L * L);
  for (int g = 0; g < G; ++g) {
    const float* __restrict__ src = maps + (static_cast<long long>(bp * G + g) * T + tt) * h * w;
    __syncthreads();  // previous group's s_src / s_tmp readers are done (and the tables are written)
    for (int i = tid; i < h * w; i += nthr) s_src[i] = __ldg(src + i);
    __syncthreads();
    for (int r = r0; r < h; r += r_step) {
      const float* row = s_src + r * w;
      s_tmp[r * L + x] = row[ix[0]] * wx[0] + row[ix[1]] * wx[1] + row[ix[2]] * wx[2] + row[ix[3]] * wx[3];
    }
    __syncthreads();
    for (int q = tid; q < L * quads_per_row; q += nthr) {
      const int y = q / quads_per_row, qx = q - y * quads_per_row;
      const int4 iy = *reinterpret_cast<const int4*>(s_iy + 4 * y);
      const float4 wy = *reinterpret_cast<const float4*>(s_wy + 4 * y);
      const float4 t0 = *reinterpret_cast<const float4*>(s_tmp + iy.x + 4 * qx);
      const float4 t1 = *reinterpret_cast<const float4*>(s_tmp + iy.y + 4 * qx);
      const float4 t2 = *reinterpret_cast<const float4*>(s_tmp + iy.z + 4 * qx);
      const float4 t3 = *reinterpret_cast<const float4*>(s_tmp + iy.w + 4 * qx);
      float4 a = dst_plane[q];
      // o = 0; o += rsum_j * wy_j for j = 0..3 (the streaming kernel's order)
      const float o0 = ((0.f + t0.x * wy.x) + t1.x * wy.y) + t2.x * wy.z + t3.x * wy.w;
      const float o1 = ((0.f + t0.y * wy.x) + t1.y * wy.y) + t2.y * wy.z + t3.y * wy.w;
      const float o2 = ((0.f + t0.z * wy.x) + t1.z * wy.y) + t2.z * wy.z + t3.z * wy.w;
      const float o3 = ((0.f + t0.w * wy.x) + t1.w * wy.y) + t2.w * wy.z + t3.w * wy.w;
      a.x += fmaxf(o0, 0.f);
      a.y += fmaxf(o1, 0.f);
      a.z += fmaxf(o2, 0.f);
      a.w += fmaxf(o3, 0.f);
      dst_plane[q] = a;
    }
  }
  }
}

__global__ void __launch_bounds__(256) heat_finalize_kernel(const float* __restrict__ acc,
                                                            float* __restrict__ out, long long n4,
                                                            long long n, float count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (; i < n4; i += stride) {
    float4 a = reinterpret_cast<const float4*>(acc)[i];
    a.x = __fdiv_rn(a.x, count); a.y = __fdiv_rn(a.y, count);
    a.z = __fdiv_rn(a.z, count); a.w = __fdiv_rn(a.w, count);
    reinterpret_cast<float4*>(out)[i] = a;
  }
  // tail (n not a multiple of 4)
  for (long long j = n4 * 4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < n; j += stride)
    out[j] = __fdiv_rn(acc[j], count);
}

}  // namespace agenda

using namespace agenda;

static int heat_upsample_accum_impl(const float* maps, float* acc, int n_planes, int h, int w, int L, int T, int G,
                                    void* stream);

extern "C" int agenda_heat_upsample_accum(const float* maps, float* acc, int n_planes, int h, int w, int L,
                                          void* stream) {
  return heat_upsample_accum_impl(maps, acc, n_planes, h, w, L, 1, 1, stream);
}

extern "C" int agenda_heat_upsample_accum_heads(const float* maps, float* acc, int n_acc_planes, int T, int G, int h,
                                                int w, int L, void* stream) {
  if (T <= 0 || G <= 0 || (n_acc_planes % T) != 0)
    return fail(AGENDA_ERR_BAD_SHAPE, "heat_upsample_accum_heads: n_acc_planes=%d T=%d G=%d", n_acc_planes, T, G);
  return heat_upsample_accum_impl(maps, acc, n_acc_planes, h, w, L, T, G, stream);
}

static int heat_upsample_accum_impl(const float* maps, float* acc, int n_planes, int h, int w, int L, int T, int G,
                                    void* stream) {
  if (n_planes == 0) return AGENDA_OK;  // (an empty batch has no buffers to point at)
  if (!maps || !acc) return fail(AGENDA_ERR_NULL_POINTER, "heat_upsample_accum: null pointer");
  if (n_planes < 0 || h <= 0 || w <= 0 || L <= 0) return fail(AGENDA_ERR_BAD_SHAPE, "heat_upsample_accum: bad shape");
  if (L % 4 != 0) return fail(AGENDA_ERR_BAD_SHAPE, "heat_upsample_accum: latent_hw must be a multiple of 4 (got %d)", L);
  if ((reinterpret_cast<uintptr_t>(acc) & 15) || ((h == L && w == L) && (reinterpret_cast<uintptr_t>(maps) & 15)))
    return fail(AGENDA_ERR_MISALIGNED, "heat_upsample_accum: buffers must be 16-byte aligned");
  if (n_planes == 0) return AGENDA_OK;
  if (!(h == L && w == L) && static_cast<long long>(h) * w <= 4096 && L <= 256 &&
      static_cast<long long>(h) * L <= 16384) {
    const int threads_t = L * (256 / L);  // a multiple of L: each thread keeps one output column in pass 1
    const size_t smem = sizeof(float) * (((static_cast<size_t>(h) * w + 3) & ~size_t(3)) + static_cast<size_t>(h) * L +
                                         4 * L) + sizeof(int) * 4 * L;
    AGENDA_DYN_SMEM(heat_upsample_accum_tiled_kernel, smem);
    // small source planes: several planes per CTA (the tap tables are built once); larger ones: one CTA per plane
    // persistent: as many CTAs as are resident at once (tap tables once per CTA, HBM latencies overlapped by prefetch)
    int per_sm = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, heat_upsample_accum_tiled_kernel, threads_t, smem) != cudaSuccess ||
        per_sm < 1)
      per_sm = 4;
    const int grid_t = std::min(n_planes, num_sms() * per_sm);
    heat_upsample_accum_tiled_kernel<<<grid_t, threads_t, smem, static_cast<cudaStream_t>(stream)>>>(maps, acc, n_planes,
                                                                                                     h, w, L, T, G);
    AGENDA_LAUNCH_CHECK("heat_upsample_accum_tiled_kernel");
    return AGENDA_OK;
  }
  const long long total = static_cast<long long>(n_planes) * L * (L / 4);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  heat_upsample_accum_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      maps, acc, n_planes, h, w, L, T, G);
  AGENDA_LAUNCH_CHECK("heat_upsample_accum_kernel");
  return AGENDA_OK;
}

extern "C" int agenda_heat_finalize(const float* acc, float* out, int64_t n_elems, int count, void* stream) {
  if (n_elems == 0) return AGENDA_OK;
  if (!acc || !out) return fail(AGENDA_ERR_NULL_POINTER, "heat_finalize: null pointer");
  if (n_elems < 0 || count < 1) return fail(AGENDA_ERR_BAD_SHAPE, "heat_finalize: need n_elems>=0, count>=1");
  if ((reinterpret_cast<uintptr_t>(acc) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(AGENDA_ERR_MISALIGNED, "heat_finalize: buffers must be 16-byte aligned");
  if (n_elems == 0) return AGENDA_OK;
  const long long n4 = n_elems / 4;
  const int threads = 256;
  long long blocks = (n4 + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  heat_finalize_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, out, n4, n_elems, static_cast<float>(count));
  AGENDA_LAUNCH_CHECK("heat_finalize_kernel");
  return AGENDA_OK;
}
