// Micro-benchmark: do tcgen05.mma groups issued by DIFFERENT warps of a CTA pipeline as well as groups issued by one
// thread?  kWarps warps each issue `rounds` x (group of 3 SS MMAs 128x64x16 | group of 4 TS MMAs 128x48x16) into their
// own accumulators, concurrently; compared with one warp issuing the same total work.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iagenda_b200/csrc -o tools/ubench/umma_2thr.bin tools/ubench/umma_2thr.cu -lcuda
#include <cstdio>
#include "sm100_common.cuh"

using namespace agenda::sm100;

// kSplit: 0 = warp 0 issues everything (QK PV QK PV ...), 1 = warp 0 issues the QK groups, warp 1 the PV groups,
//         2 = three warps, each its own tile's QK + PV (one issuer per query tile)
template <int kSplit, int kGap>
__global__ void __launch_bounds__(128, 1) k(long long* res, int rounds) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4], sink;
  __shared__ uint32_t tmem_base;
  __shared__ long long t_end[4];
  const int warp = threadIdx.x >> 5;
  constexpr int BN = 64;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    mbar_init(&sink, 1 << 20);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  constexpr uint32_t idesc_qk = make_idesc(128, BN, 0);
  constexpr uint32_t idesc_pv = make_idesc(128, 48, 1);
  const uint64_t q_desc = make_sdesc(smem_u32(smem), 16, 1024);
  const uint64_t k_desc = make_sdesc(smem_u32(smem + 49152), 16, 1024);
  const uint64_t v_desc = make_sdesc(smem_u32(smem + 49152 + 16384), BN * 128, 1024);
  auto qk = [&](int t) {
    for (int kk = 0; kk < 3; ++kk)
      umma_ss(tmem + t * BN, q_desc + ((t * 16384 + kk * 32) >> 4), k_desc + (kk * 32 >> 4), idesc_qk, kk != 0);
    umma_commit(&sink);
  };
  auto pv = [&](int t) {
    for (int kk = 0; kk < BN / 16; ++kk)
      umma_ts(tmem + 3 * BN * 3 / 2 + t * 48, tmem + 3 * BN + t * BN / 2 + kk * 8, v_desc + (kk * 2048 >> 4), idesc_pv, 1);
    umma_commit(&sink);
  };
  auto gap = [&]() {
    if (kGap) { const long long e = clock64() + kGap; while (clock64() < e) { } }
  };
  const long long t0 = clock64();
  const int n_issuers = kSplit == 0 ? 1 : (kSplit == 1 ? 2 : 3);
  if (warp < n_issuers) {
    if (elect_one()) {
      for (int r = 0; r < rounds; ++r) {
        if (kSplit == 0) { qk(0); pv(0); qk(1); pv(1); qk(2); pv(2); }
        else if (kSplit == 1) {
          if (warp == 0) { qk(0); gap(); qk(1); gap(); qk(2); gap(); } else { pv(0); gap(); pv(1); gap(); pv(2); gap(); }
        } else { qk(warp); gap(); pv(warp); gap(); }
      }
      umma_commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    if ((threadIdx.x & 31) == 0) t_end[warp] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long e = 0;
    for (int i = 0; i < n_issuers; ++i) e = t_end[i] > e ? t_end[i] : e;
    res[0] = e - t0;
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int kSplit, int kGap>
void run(const char* name, long long* d_res) {
  auto kern = k<kSplit, kGap>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int rounds = 64;
  kern<<<148, 128, 100 * 1024>>>(d_res, rounds);
  kern<<<148, 128, 100 * 1024>>>(d_res, rounds);
  cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d_res, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-60s gap %4d: %7.1f cyc/round of 3 QK + 3 PV groups  (%s)\n", name, kGap, double(h[0]) / rounds,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_res; cudaMalloc(&d_res, 16);
  run<0, 0>("one warp issues QK PV QK PV QK PV", d_res);
  run<1, 0>("warp 0: QK groups, warp 1: PV groups", d_res);
  run<1, 200>("warp 0: QK groups, warp 1: PV groups", d_res);
  run<1, 400>("warp 0: QK groups, warp 1: PV groups", d_res);
  run<2, 0>("three warps, each QK(t) PV(t)", d_res);
  run<2, 200>("three warps, each QK(t) PV(t)", d_res);
  run<2, 600>("three warps, each QK(t) PV(t)", d_res);
  return 0;
}
