// Micro-benchmark: issue + execution rate of tcgen05.mma (kind::f16, bf16 in, fp32 accumulate, M = 128, K = 16 per
// instruction) for the tile shapes the attention kernels use.  One CTA per SM, one elected thread issues `n` MMAs
// back to back, then commits; cycles are taken around the issue loop and up to the commit's mbarrier arrival.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iagenda_b200/csrc -o tools/ubench/umma.bin tools/ubench/umma.cu -lcuda
#include <cstdio>
#include "sm100_common.cuh"

using namespace agenda::sm100;

template <int N, bool kTS, bool kBMn, int kChain>
__global__ void __launch_bounds__(128, 1) k(long long* res, int n_mma) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc(128, N, kBMn ? 1 : 0);
    const uint64_t a_desc = make_sdesc(smem_u32(smem), 16, 1024);
    const uint64_t b_desc = kBMn ? make_sdesc(smem_u32(smem + 32768), N * 128, 1024) : make_sdesc(smem_u32(smem + 32768), 16, 1024);
    long long t0 = 0, t1 = 0, t2 = 0;
    t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < n_mma; ++i) {
        // kChain accumulators in rotation: 1 = every MMA accumulates into the same D (a dependent chain)
        const uint32_t d = tmem + 256 + (i % kChain) * 64;
        if (kTS) umma_ts(d, tmem + (i & 7) * 8, b_desc + ((i & 3) * 2), idesc, 1);
        else umma_ss(d, a_desc + ((i & 3) * 2), b_desc + ((i & 3) * 2), idesc, 1);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    t1 = clock64();
    mbar_wait(&bar, 0);
    t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { res[0] = t1 - t0; res[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool kTS, bool kBMn, int kChain>
void run(const char* name, long long* d_res) {
  auto kern = k<N, kTS, kBMn, kChain>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int n : {8, 64, 512}) {
    kern<<<148, 128, 100 * 1024>>>(d_res, n);
    kern<<<148, 128, 100 * 1024>>>(d_res, n);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d_res, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-34s n=%3d: issue %7lld cyc (%6.1f/MMA)  done %7lld cyc (%6.1f/MMA; ideal %5.1f)\n", name, n, h[0], double(h[0]) / n,
           h[1], double(h[1]) / n, N / 2.0);
  }
}

int main() {
  long long* d_res; cudaMalloc(&d_res, 16);
  run<128, false, false, 1>("SS 128x128x16 K-major B, chain", d_res);
  run<128, false, false, 2>("SS 128x128x16 K-major B, 2 accs", d_res);
  run<64, false, false, 1>("SS 128x64x16 K-major B, chain", d_res);
  run<256, false, false, 1>("SS 128x256x16 K-major B, chain", d_res);
  run<48, false, true, 1>("SS 128x48x16 MN-major B, chain", d_res);
  run<48, true, true, 1>("TS 128x48x16 MN-major B, chain", d_res);
  run<48, true, true, 2>("TS 128x48x16 MN-major B, 2 accs", d_res);
  run<64, true, true, 1>("TS 128x64x16 MN-major B, chain", d_res);
  run<128, true, true, 1>("TS 128x128x16 MN-major B, chain", d_res);
  run<160, true, true, 1>("TS 128x160x16 MN-major B, chain", d_res);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
