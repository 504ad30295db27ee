// Timing / event-trace harness for the d = 40 self-attention kernels: the shipped v2 kernel (default) and the v3
// experiment (tools/ubench/attn_self_v3_experiment.cu, V3=1).  Debug tool, not part of the product library.
// build: SRC="tools/ubench/bench_self.cu tools/ubench/attn_self_v3_experiment.cu agenda_b200/csrc/attn_self_sm100_v2.cu \
//             agenda_b200/csrc/attn_sm100.cu agenda_b200/csrc/abi.cu agenda_b200/csrc/attn_f32.cu"
//        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DAGENDA_VARIANTS \
//             -o tools/ubench/bench_self.bin $SRC -lcuda        (+ -DAGENDA_V2_TRACE: event trace of the v3 kernel)
// run:   [AGENDA_KNOBS=1] [V3=1] [REPS=4000] [VARIANT=n] tools/ubench/bench_self.bin [unit=1] [B N H d]
//        REPS=4000: sustained timing (the B200 reaches its 1000 W power cap after ~0.5 s of this kernel and the SM
//        clock drops from 1965 to ~1800 MHz; 20 launches do not show that)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

extern "C" int agenda_attn_self_fwd_strided(const void*, const void*, const void*, void*, int, int, int, int, int, long long,
                                            float, void*);
#ifdef AGENDA_V2_TRACE
extern "C" int agenda_v3_trace_read(long long*, int);
#endif
extern "C" const char* agenda_last_error(void);
namespace agenda {
int attn_self_sm100_v3(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int d, float scale,
                       bool unit, bool exact, void* stream, long long ld);
}
extern "C" int agenda_attn_self_fwd_variant(const void*, const void*, const void*, void*, int, int, int, int, float, int, void*);

int main(int argc, char** argv) {
  const int unit = argc > 1 ? atoi(argv[1]) : 1;
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
  const float scale = unit ? 0.f : 1.0f / sqrtf(float(d));
  const bool v3 = getenv("V3") && atoi(getenv("V3"));
  auto launch = [&]() {
    if (v3) return agenda::attn_self_sm100_v3(q, k, v, o, B, H, N, d, 1.0f / sqrtf(float(d)), unit != 0, false, nullptr, (long long)H * d);
    return agenda_attn_self_fwd_strided(q, k, v, o, 1, B, H, N, d, (long long)H * d, scale, nullptr);
  };
  for (int it = 0; it < 3; ++it) {
    int rc = launch();
    if (rc) { printf("error %d: %s\n", rc, agenda_last_error()); return 1; }
  }
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const int reps = getenv("REPS") ? atoi(getenv("REPS")) : 20;
  const int variant = getenv("VARIANT") ? atoi(getenv("VARIANT")) : -1;
  for (int it = 0; it < reps; ++it) {
    if (variant >= 0) agenda_attn_self_fwd_variant(q, k, v, o, B, H, N, d, 1.0f / sqrtf(float(d)), variant, nullptr);
    else launch();
  }
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%s unit %d B=%d N=%d H=%d d=%d: %.4f ms %s, %.1f TFLOP/s useful\n", v3 ? "v3" : "v2", unit, B, N, H, d, ms / reps,
#ifdef AGENDA_V2_TRACE
         "(traced build)",
#else
         "",
#endif
         4.0 * B * H * double(N) * N * d / (ms / reps) / 1e9);
  {  // numerics: a few rows of (b, h) = (0, 0) and of the last (b, h) against a CPU evaluation
    std::vector<__nv_bfloat16> ho(n);
    cudaMemcpy(ho.data(), o, n * 2, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int pick = 0; pick < 6; ++pick) {
      const int b = pick < 3 ? 0 : B - 1, hh = pick < 3 ? 0 : H - 1, r = ((pick % 3) * 1531 + 7) % N;
      std::vector<double> p(N);
      double mx = -1e300, l = 0;
      for (int c = 0; c < N; ++c) {
        double a = 0;
        for (int e = 0; e < d; ++e)
          a += double(__bfloat162float(h[(size_t(b) * N + r) * H * d + hh * d + e])) *
               double(__bfloat162float(h[(size_t(b) * N + c) * H * d + hh * d + e]));
        p[c] = unit ? a * 0.6931471805599453 : a * scale;
        mx = p[c] > mx ? p[c] : mx;
      }
      for (int c = 0; c < N; ++c) { p[c] = exp(p[c] - mx); l += p[c]; }
      for (int e = 0; e < d; ++e) {
        double acc = 0;
        for (int c = 0; c < N; ++c) acc += p[c] * double(__bfloat162float(h[(size_t(b) * N + c) * H * d + hh * d + e]));
        const double got = __bfloat162float(ho[(size_t(b) * N + r) * H * d + hh * d + e]);
        const double err = fabs(got - acc / l);
        worst = err > worst ? err : worst;
      }
    }
    printf("max abs error of 6 sampled rows vs CPU fp64: %.3e %s\n", worst, worst < 2e-2 ? "ok" : "MISMATCH");
  }
#ifdef AGENDA_V2_TRACE
  std::vector<long long> tr(8 * 24 * 8);
  int got = agenda_v3_trace_read(tr.data(), int(tr.size()));
  if (got <= 0) { printf("trace read failed %d\n", got); return 1; }
  auto at = [&](int a, int j, int e) { return tr[(a * 24 + j) * 8 + e]; };
  const long long t0 = at(0, 0, 0);
  printf("softmax tile t: wait_s s_ready ld_done exp_done p_arrived | waitS ld exp st | period\n");
  printf("mma warp t:     v_ready p_full pv_issued qk_issued | wait_p pv_issue qk_issue\n");
  for (int j = 0; j < 24; ++j) {
    for (int a = 0; a < 4; ++a) {
      if (at(a, j, 0) == 0) continue;
      printf("j=%2d wg%d:", j + 20, a);
      for (int e = 0; e < 5; ++e) printf(" %7lld", at(a, j, e) - t0);
      printf(" | waitS %5lld ld %5lld exp %5lld st %5lld | %5lld\n", at(a, j, 1) - at(a, j, 0), at(a, j, 2) - at(a, j, 1),
             at(a, j, 3) - at(a, j, 2), at(a, j, 4) - at(a, j, 3), j ? at(a, j, 0) - at(a, j - 1, 0) : 0);
    }
    for (int a = 4; a < 6; ++a) {
      if (at(a, j, 0) == 0) continue;
      printf("j=%2d mma%d:", j + 20, a - 4);
      for (int e = 0; e < 4; ++e) printf(" %7lld", at(a, j, e) ? at(a, j, e) - t0 : -1);
      printf(" | wait_p %5lld pv_issue %5lld qk_issue %5lld\n", at(a, j, 1) - at(a, j, 0), at(a, j, 2) - at(a, j, 1),
             at(a, j, 3) - at(a, j, 2));
    }
  }
#endif
  return 0;
}
