// Micro-benchmark: how much FMA-pipe work hides under MUFU.EX2 for one / two / three warps per SM sub-partition.
// Body: 8 x (1 MUFU.EX2 + K FFMA2 on independent accumulators [+ 1 cvt.bf16x2 per 2 MUFU]).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mix.bin tools/ubench/mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K, bool kCvt, int kOp>
__global__ void __launch_bounds__(512) k(uint32_t* out, long long* cyc, int iters) {
  uint32_t x[8];
  unsigned long long acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = 0xbf800000u + threadIdx.x * 8 + i; acc[i] = 0x3f8000003f800000ull + i; }
  uint32_t pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (kOp == 0) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(acc[(i * K + j) & 7]));
        if (kOp == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(*reinterpret_cast<uint32_t*>(&acc[(i * K + j) & 7])));
        if (kOp == 2) asm volatile("mad.lo.u32 %0, %0, 8388608, %0;" : "+r"(*reinterpret_cast<uint32_t*>(&acc[(i * K + j) & 7])));
      }
      if (kCvt && (i & 1)) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "r"(x[i]), "r"(x[i - 1]));
    }
  }
  const long long t1 = clock64();
  uint32_t s = pk;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i] ^ uint32_t(acc[i]) ^ uint32_t(acc[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, bool kCvt, int kOp>
void run(const char* name, uint32_t* out, long long* cyc) {
  const int iters = 2000;
  for (int threads : {128, 256, 384}) {
    k<K, kCvt, kOp><<<148, threads>>>(out, cyc, iters);
    k<K, kCvt, kOp><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_mufu = double(h[0]) / (iters * 8.0) / (threads / 128);  // cycles per MUFU warp-instruction per sub-partition
    printf("%-28s K=%d cvt=%d warps/SMSP=%d: %6.2f cycles per (MUFU + K ops) per sub-partition\n", name, K, int(kCvt), threads / 128, per_mufu);
  }
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0, false, 0>("MUFU only", out, cyc);
  run<0, true, 0>("MUFU + cvt", out, cyc);
  run<1, false, 0>("MUFU + FFMA2", out, cyc);
  run<2, false, 0>("MUFU + FFMA2", out, cyc);
  run<3, false, 0>("MUFU + FFMA2", out, cyc);
  run<4, false, 0>("MUFU + FFMA2", out, cyc);
  run<6, false, 0>("MUFU + FFMA2", out, cyc);
  run<4, true, 0>("MUFU + FFMA2 + cvt", out, cyc);
  run<2, false, 1>("MUFU + FFMA", out, cyc);
  run<4, false, 1>("MUFU + FFMA", out, cyc);
  run<6, false, 1>("MUFU + FFMA", out, cyc);
  run<2, false, 2>("MUFU + IMAD", out, cyc);
  run<4, false, 2>("MUFU + IMAD", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
