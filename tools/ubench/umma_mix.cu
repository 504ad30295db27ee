// Micro-benchmark: throughput of the attention kernels' MMA mix issued by ONE thread: groups of 3 SS MMAs 128x64x16
// (QK^T of a 64-key tile at d = 40, K-major B) and groups of 4 TS MMAs 128x48x16 (P V, MN-major B), three query tiles
// (separate accumulators), in different interleavings, with / without a tcgen05.commit behind every group.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iagenda_b200/csrc -o tools/ubench/umma_mix.bin tools/ubench/umma_mix.cu -lcuda
#include <cstdio>
#include "sm100_common.cuh"

using namespace agenda::sm100;

// mode: 0 = QK groups only, 1 = PV groups only, 2 = QK PV QK PV QK PV (alternating), 3 = QK QK QK PV PV PV
//       4 = alternating, QK as TS form too (A = Q from TMEM)
template <int kMode, bool kCommit, int BN>
__global__ void __launch_bounds__(128, 1) k(long long* res, int rounds) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, sink;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&sink, 1 << 20); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    constexpr uint32_t idesc_qk = make_idesc(128, BN, 0);
    constexpr uint32_t idesc_pv = make_idesc(128, 48, 1);
    const uint64_t q_desc = make_sdesc(smem_u32(smem), 16, 1024);            // 3 tiles x 16 KB
    const uint64_t k_desc = make_sdesc(smem_u32(smem + 49152), 16, 1024);    // BN x 128 B
    const uint64_t v_desc = make_sdesc(smem_u32(smem + 49152 + 16384), BN * 128, 1024);
    long long t0 = clock64(), t1, t2;
    if (elect_one()) {
      for (int r = 0; r < rounds; ++r) {
        auto qk = [&](int t) {
          for (int kk = 0; kk < 3; ++kk) {
            if (kMode == 4) umma_ts(tmem + t * BN, tmem + 480 + kk * 8, k_desc + (kk * 32 >> 4), idesc_qk, kk != 0);
            else umma_ss(tmem + t * BN, q_desc + ((t * 16384 + kk * 32) >> 4), k_desc + (kk * 32 >> 4), idesc_qk, kk != 0);
          }
          if (kCommit) umma_commit(&sink);
        };
        auto pv = [&](int t) {
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_ts(tmem + 3 * BN * 3 / 2 + t * 48, tmem + 3 * BN + t * BN / 2 + kk * 8, v_desc + (kk * 2048 >> 4), idesc_pv, 1);
          if (kCommit) umma_commit(&sink);
        };
        if (kMode == 0) { qk(0); qk(1); qk(2); }
        else if (kMode == 1) { pv(0); pv(1); pv(2); }
        else if (kMode == 3) { qk(0); qk(1); qk(2); pv(0); pv(1); pv(2); }
        else { qk(0); pv(0); qk(1); pv(1); qk(2); pv(2); }
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

template <int kMode, bool kCommit, int BN>
void run(const char* name, long long* d_res) {
  auto kern = k<kMode, kCommit, BN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int rounds = 64;
  kern<<<148, 128, 100 * 1024>>>(d_res, rounds);
  kern<<<148, 128, 100 * 1024>>>(d_res, rounds);
  cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d_res, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-52s BN=%3d commit=%d: issue %7.1f cyc/round, done %7.1f cyc/round  (%s)\n", name, BN, int(kCommit), double(h[0]) / rounds,
         double(h[1]) / rounds, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_res; cudaMalloc(&d_res, 16);
  run<0, false, 64>("3 QK groups (3 SS MMAs each)", d_res);
  run<0, true, 64>("3 QK groups (3 SS MMAs each)", d_res);
  run<1, false, 64>("3 PV groups (4 TS MMAs each)", d_res);
  run<1, true, 64>("3 PV groups (4 TS MMAs each)", d_res);
  run<2, false, 64>("QK PV QK PV QK PV", d_res);
  run<2, true, 64>("QK PV QK PV QK PV", d_res);
  run<3, false, 64>("QK QK QK PV PV PV", d_res);
  run<3, true, 64>("QK QK QK PV PV PV", d_res);
  run<4, false, 64>("QK(TS) PV QK(TS) PV QK(TS) PV", d_res);
  run<4, true, 64>("QK(TS) PV QK(TS) PV QK(TS) PV", d_res);
  run<0, true, 128>("3 QK groups", d_res);
  run<1, true, 128>("3 PV groups (8 TS MMAs each)", d_res);
  run<2, true, 128>("QK PV QK PV QK PV", d_res);
  run<3, true, 128>("QK QK QK PV PV PV", d_res);
  run<4, true, 128>("QK(TS) PV x3", d_res);
  return 0;
}
