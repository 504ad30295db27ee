// Event-trace harness for the v2 self-attention kernel (debug tool, not part of the product library).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DAGENDA_V2_TRACE \
//          -o tools/ubench/trace_attn.bin tools/ubench/trace_attn.cu agenda_b200/csrc/*.cu -lcuda
// run:   tools/ubench/trace_attn.bin [variant] [B N H d]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

extern "C" int agenda_attn_self_fwd_variant(const void*, const void*, const void*, void*, int, int, int, int, float, int, void*);
extern "C" int agenda_v2_trace_read(long long*, int);
extern "C" const char* agenda_last_error(void);

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 14;
  const int B = argc > 5 ? atoi(argv[2]) : 16, N = argc > 5 ? atoi(argv[3]) : 4096, H = argc > 5 ? atoi(argv[4]) : 8,
            d = argc > 5 ? atoi(argv[5]) : 40;
  const size_t n = size_t(B) * N * H * d;
  std::vector<__nv_bfloat16> h(n);
  srand(1);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((rand() / float(RAND_MAX) - 0.5f) * 3.f);
  __nv_bfloat16 *q, *k, *v, *o;
  cudaMalloc(&q, n * 2); cudaMalloc(&k, n * 2); cudaMalloc(&v, n * 2); cudaMalloc(&o, n * 2);
  cudaMemcpy(q, h.data(), n * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(k, h.data(), n * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(v, h.data(), n * 2, cudaMemcpyHostToDevice);
  for (int it = 0; it < 3; ++it) {
    int rc = agenda_attn_self_fwd_variant(q, k, v, o, B, H, N, d, 1.0f / sqrtf(float(d)), variant, nullptr);
    if (rc) { printf("error %d: %s\n", rc, agenda_last_error()); return 1; }
  }
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) agenda_attn_self_fwd_variant(q, k, v, o, B, H, N, d, 1.0f / sqrtf(float(d)), variant, nullptr);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("variant %d B=%d N=%d H=%d d=%d: %.4f ms (traced build)\n", variant, B, N, H, d, ms / 10);
  std::vector<long long> tr(8 * 24 * 8);
  int got = agenda_v2_trace_read(tr.data(), int(tr.size()));
  if (got <= 0) { printf("trace read failed %d\n", got); return 1; }
  auto at = [&](int a, int j, int e) { return tr[(a * 24 + j) * 8 + e]; };
  const long long t0 = at(0, 0, 0);
  printf("softmax events per tile: wait_s | s_ready | max_done | pv_ok | exp_done | p_arrived   (cycles since WG0 start)\n");
  for (int j = 0; j < 16; ++j) {
    for (int a = 0; a < 6; ++a) {
      if (at(a, j, 0) == 0) continue;
      printf("j=%2d warp%d:", j, a);
      for (int e = 0; e < 6; ++e) printf(" %7lld", at(a, j, e) - t0);
      printf("   | waitS %5lld ld+max %5lld waitPV %5lld exp %5lld st %5lld\n", at(a, j, 1) - at(a, j, 0), at(a, j, 2) - at(a, j, 1),
             at(a, j, 3) - at(a, j, 2), at(a, j, 4) - at(a, j, 3), at(a, j, 5) - at(a, j, 4));
      if (at(a, j, 6)) printf("            first 32 cols %5lld, wait pv_done %5lld, rest %5lld\n", at(a, j, 6) - at(a, j, 3), at(a, j, 7) - at(a, j, 6), at(a, j, 4) - at(a, j, 7));
    }
    printf("        MMA: QK issued");
    for (int t = 0; t < 3; ++t) if (at(6, j, t)) printf(" t%d@%7lld", t, at(6, j, t) - t0);
    printf("  PV issued");
    for (int t = 0; t < 3; ++t) if (at(6, j, 3 + t)) printf(" t%d@%7lld", t, at(6, j, 3 + t) - t0);
    printf("\n        issue cost: QK");
    for (int t = 0; t < 3; ++t) if (at(6, j, t)) printf(" t%d %5lld", t, at(6, j, t) - at(7, j, 3 + t));
    printf("  PV");
    for (int t = 0; t < 3; ++t) if (at(6, j, 3 + t)) printf(" t%d %5lld (start@%7lld)", t, at(6, j, 3 + t) - at(7, j, t), at(7, j, t) - t0);
    printf("\n");
  }
  return 0;
}
