// Micro-benchmark: TMA tile-load rate per SM for the attention kernels' operand tiles — box (64 elems, 1 head, R rows, 1)
// of a [B, N, H*d] bf16 tensor viewed as (d, H, N, B), 128-byte swizzle.  One elected thread per CTA streams tiles
// through a 4-stage ring (no consumer), one CTA per SM, tensor L2-resident.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Iagenda_b200/csrc \
//          -o tools/ubench/tma.bin tools/ubench/tma.cu agenda_b200/csrc/*.cu -lcuda
#include <cstdio>
#include <vector>
#include "sm100_common.cuh"

using namespace agenda;
using namespace agenda::sm100;

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap map, int rows, int n_tiles, int H, int N, long long* res) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[4];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int b = blockIdx.x % 16;
    const long long t0 = clock64();
    for (int i = 0; i < n_tiles + 4; ++i) {
      const int s = i & 3;
      if (i >= 4) mbar_wait(&full[s], ((i - 4) >> 2) & 1);
      if (i < n_tiles && elect_one()) {
        mbar_expect_tx(&full[s], rows * 128);
        tma_load_4d(&map, &full[s], smem + s * rows * 128, 0, i % H, ((i / H) * rows + blockIdx.x * 64) % (N - rows), b);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) res[0] = t1 - t0;
  }
}

// The cross-attention kernel's load pattern: per head one Q tile (128 rows) + K and V tiles (80 rows each) into a
// 3-stage ring, grid (N/128, B) with 1 or 2 CTAs resident per SM; no consumer (the stage is recycled when it lands).
__global__ void __launch_bounds__(64, 2) kx(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                                            const __grid_constant__ CUtensorMap mv, int H, int q_tiles_per_cta,
                                            long long* res) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[3];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int stage_bytes = (q_tiles_per_cta * 128 + 160) * 128;
  if (threadIdx.x < 32) {
    const int b = blockIdx.y, q0 = blockIdx.x * 128 * q_tiles_per_cta;
    const long long t0 = clock64();
    for (int h = 0; h < H + 3; ++h) {
      const int s = h % 3;
      if (h >= 3) mbar_wait(&full[s], ((h - 3) / 3) & 1);
      if (h < H && elect_one()) {
        mbar_expect_tx(&full[s], stage_bytes);
        unsigned char* st = smem + s * stage_bytes;
        for (int t = 0; t < q_tiles_per_cta; ++t) tma_load_4d(&mq, &full[s], st + t * 16384, 0, h, q0 + t * 128, b);
        tma_load_4d(&mk, &full[s], st + q_tiles_per_cta * 16384, 0, h, 0, b);
        tma_load_4d(&mv, &full[s], st + q_tiles_per_cta * 16384 + 10240, 0, h, 0, b);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) res[0] = t1 - t0;
  }
}

static void cross_pattern(long long* d_res) {
  const int B = 16, H = 8, N = 4096, M = 77, d = 40;
  void *q, *k, *v;
  cudaMalloc(&q, size_t(B) * N * H * d * 2); cudaMalloc(&k, size_t(B) * M * H * d * 2); cudaMalloc(&v, size_t(B) * M * H * d * 2);
  cudaMemset(q, 0, size_t(B) * N * H * d * 2);
  CUtensorMap mq, mk, mv;
  make_head_map(&mq, q, B, H, N, d, 128); make_head_map(&mk, k, B, H, M, d, 80); make_head_map(&mv, v, B, H, M, d, 80);
  for (int qt : {1, 2, 4}) {
    const size_t smem = 3 * (qt * 128 + 160) * 128 + 2048;
    cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(N / (128 * qt), B);
    kx<<<grid, 64, smem>>>(mq, mk, mv, H, qt, d_res);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) kx<<<grid, 64, smem>>>(mq, mk, mv, H, qt, d_res);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, d_res, 8, cudaMemcpyDeviceToHost);
    printf("cross-attention load pattern, %d query tile(s) per CTA (%zu KB smem): %.1f us per launch, CTA(0,0) %lld cycles for %d heads\n",
           qt, smem / 1024, ms * 100, c, H);
  }
}

int main() {
  long long* d_res; cudaMalloc(&d_res, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 128 * 128 + 2048);
  for (int d : {40, 64}) {
    const int B = 16, H = 8, N = 4096;
    void* buf; cudaMalloc(&buf, size_t(B) * N * H * d * 2); cudaMemset(buf, 0, size_t(B) * N * H * d * 2);
    for (int rows : {128, 80, 64}) {
      CUtensorMap map;
      if (make_head_map(&map, buf, B, H, N, d, rows) != 0) { printf("map failed\n"); return 1; }
      const int n_tiles = 2000;
      for (int ctas : {1, 148}) {
        k<<<ctas, 64, 4 * 128 * 128 + 2048>>>(map, rows, n_tiles, H, N, d_res);
        k<<<ctas, 64, 4 * 128 * 128 + 2048>>>(map, rows, n_tiles, H, N, d_res);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, d_res, 8, cudaMemcpyDeviceToHost);
        printf("d=%d box rows=%3d CTAs=%3d: %6.1f cycles/tile  %5.2f cycles/row  %6.1f useful B/clk/SM\n", d, rows, ctas,
               double(c) / n_tiles, double(c) / n_tiles / rows, rows * d * 2.0 * n_tiles / c);
      }
    }
    cudaFree(buf);
  }
  cross_pattern(d_res);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
