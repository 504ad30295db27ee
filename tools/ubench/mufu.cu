// Micro-benchmark: MUFU.EX2 throughput per SM for f32 / bf16x2 / f16x2 operands, FFMA2 and mixed streams (sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu tools/ubench/mufu.cu ; run: /tmp/mufu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t* out, long long* cyc, int iters) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0xbf800000u + threadIdx.x * 8 + i;  // ~ -1.0f
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 3) asm volatile("ex2.approx.ftz.bf16 %0, %0;" : "+h"(*reinterpret_cast<unsigned short*>(&x[i])));
      if (MODE == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]));
      if (MODE == 5) {  // the softmax inner loop's mix: two ex2 + one pack per element pair
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
        if (i & 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(x[i]) : "r"(x[i]), "r"(x[i - 1]));
      }
      if (MODE == 6) {  // ex2 + packed fma + packed add (no conversion)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
        if (i & 1) {
          unsigned long long a, b;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(x[i - 1]), "r"(x[i]));
          asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(b) : "l"(a));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(x[i - 1]), "=r"(x[i]) : "l"(b));
        }
      }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[] = {"ex2.f32", "ex2.bf16x2", "ex2.f16x2", "ex2.bf16 (scalar)", "cvt.bf16x2.f32", "2 ex2 + 1 cvt pack", "2 ex2 + 1 ffma2"};
  for (int threads : {128, 512}) {
    for (int mode = 0; mode < 7; ++mode) {
      auto run = [&](auto kern) {
        kern<<<148, threads>>>(out, cyc, iters);
        kern<<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = h[0];
        double ops = double(iters) * 8 * threads;  // instructions-lanes per SM
        double per = (mode == 1 || mode == 2) ? 2.0 : 1.0;
        printf("%-18s threads=%4d: %8.0f cycles, %.2f results/clk/SM (%.2f lane-instr/clk/SM)\n", names[mode], threads, c,
               ops * per / c, ops / c);
      };
      if (mode == 0) run(k<0>); if (mode == 1) run(k<1>); if (mode == 2) run(k<2>); if (mode == 3) run(k<3>);
      if (mode == 4) run(k<4>); if (mode == 5) run(k<5>); if (mode == 6) run(k<6>);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
