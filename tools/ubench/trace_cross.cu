// Event-trace harness for the resident-K/V cross-attention kernel (debug tool).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DAGENDA_XRES_TRACE \
//          -o tools/ubench/trace_cross.bin tools/ubench/trace_cross.cu agenda_b200/csrc/*.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
extern "C" int agenda_attn_cross_fwd_heat(const void*, const void*, const void*, void*, int, int, int, int, int, int, float,
                                          const int*, int, int, float*, int, void*);
extern "C" int agenda_xres_trace_read(long long*, int);
extern "C" const char* agenda_last_error(void);
int main() {
  const int B = 16, N = 4096, H = 8, d = 40, M = 77, T = 3;
  const size_t nq = size_t(B) * N * H * d, nk = size_t(B) * M * H * d;
  std::vector<__nv_bfloat16> hq(nq), hk(nk);
  srand(1);
  for (auto& x : hq) x = __float2bfloat16((rand() / float(RAND_MAX) - 0.5f) * 3.f);
  for (auto& x : hk) x = __float2bfloat16((rand() / float(RAND_MAX) - 0.5f) * 3.f);
  __nv_bfloat16 *q, *k, *v, *o; float* maps;
  cudaMalloc(&q, nq * 2); cudaMalloc(&k, nk * 2); cudaMalloc(&v, nk * 2); cudaMalloc(&o, nq * 2);
  cudaMalloc(&maps, size_t(B / 2) * T * N * 4); cudaMemset(maps, 0, size_t(B / 2) * T * N * 4);
  cudaMemcpy(q, hq.data(), nq * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(k, hk.data(), nk * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(v, hk.data(), nk * 2, cudaMemcpyHostToDevice);
  int toks[3] = {5, 6, 7};
  for (int it = 0; it < 3; ++it) {
    int rc = agenda_attn_cross_fwd_heat(q, k, v, o, 1, B, H, N, M, d, 1.f / sqrtf(float(d)), toks, T, B / 2, maps, 1, nullptr);
    if (rc) { printf("error %d %s\n", rc, agenda_last_error()); return 1; }
  }
  cudaDeviceSynchronize();
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int it = 0; it < 20; ++it)
      agenda_attn_cross_fwd_heat(q, k, v, o, 1, B, H, N, M, d, 1.f / sqrtf(float(d)), toks, T, B / 2, maps, 1, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("B=%d N=%d H=%d d=%d T=%d: %.1f us per launch (traced build, same buffers every launch)\n", B, N, H, d, T, ms * 50.f);
    for (int it = 0; it < 20; ++it)
      agenda_attn_cross_fwd_heat(q, k, v, o, 1, B, H, N, M, d, 1.f / sqrtf(float(d)), toks, T, B, maps, 1, nullptr);
    cudaEventRecord(e0);
    for (int it = 0; it < 20; ++it)
      agenda_attn_cross_fwd_heat(q, k, v, o, 1, B, H, N, M, d, 1.f / sqrtf(float(d)), toks, 0, B / 2, nullptr, 1, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  without the heat epilogue: %.1f us per launch\n", ms * 50.f);
    // the traced CTA is (tile 0, batch 0): b_first = 0 so that it runs the heat epilogue
    float* maps_all; cudaMalloc(&maps_all, size_t(B) * T * N * 4); cudaMemset(maps_all, 0, size_t(B) * T * N * 4);
    agenda_attn_cross_fwd_heat(q, k, v, o, 1, B, H, N, M, d, 1.f / sqrtf(float(d)), toks, T, 0, maps_all, 1, nullptr);
    cudaDeviceSynchronize();
  }
  std::vector<long long> tr(640);
  if (agenda_xres_trace_read(tr.data(), 640) < 0) { printf("trace read failed\n"); return 1; }
  auto at = [&](int a, int s, int e) { return tr[(a * 40 + s) * 4 + e]; };
  const long long t0 = at(0, 0, 0);
  for (int s = 0; s < 24; ++s) {
    const int wg = s & 1;
    printf("s=%2d WG%d: wait_s@%7lld s_ready +%5lld softmax +%5lld drain +%5lld | MMA: q_ready@%7lld QK issue +%4lld, PV start@%7lld +%4lld | TMA Q issue@%7lld\n",
           s, wg, at(wg, s, 0) - t0, at(wg, s, 1) - at(wg, s, 0), at(wg, s, 2) - at(wg, s, 1), at(wg, s, 3) - at(wg, s, 2),
           at(2, s, 0) - t0, at(2, s, 1) - at(2, s, 0), at(2, s, 2) - t0, at(2, s, 3) - at(2, s, 2), at(3, s, 0) - t0);
  }
  return 0;
}
