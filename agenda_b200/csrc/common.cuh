// Shared helpers for libagenda_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/agenda_b200.h"

namespace agenda {

// thread-local last-error text (agenda_last_error())
char* last_error_buf();
int fail(int code, const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return AGENDA_OK;
  return fail(AGENDA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define AGENDA_CUDA(call)                                       \
  do {                                                          \
    int _rc = ::agenda::check_cuda((call), #call);              \
    if (_rc != AGENDA_OK) return _rc;                           \
  } while (0)

#define AGENDA_LAUNCH_CHECK(name)                               \
  do {                                                          \
    int _rc = ::agenda::check_cuda(cudaGetLastError(), name);   \
    if (_rc != AGENDA_OK) return _rc;                           \
  } while (0)

inline int num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

// Test / measurement switches.  They are consulted only when AGENDA_KNOBS=1 was in the environment when
// the library was first used (tests/conftest.py sets it): the production hot path makes no getenv call at all.
inline const char* knob(const char* name) {
  static const bool on = [] { const char* e = getenv("AGENDA_KNOBS"); return e != nullptr && atoi(e) != 0; }();
  return on ? getenv(name) : nullptr;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (launch site, device) instead of once per launch: the
// attribute sticks to the function in the device's context.  `bytes` may vary per call: it is raised when a larger
// value comes along.  (A benign race between host threads at worst repeats the call.)
#define AGENDA_DYN_SMEM(kernel, bytes)                                                                              \
  do {                                                                                                              \
    static int _agenda_set[64] = {0};                                                                               \
    int _dev = 0;                                                                                                   \
    cudaGetDevice(&_dev);                                                                                           \
    const int _need = static_cast<int>(bytes);                                                                      \
    if (_agenda_set[_dev & 63] < _need) {                                                                           \
      AGENDA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, _need));                \
      _agenda_set[_dev & 63] = _need;                                                                               \
    }                                                                                                               \
  } while (0)

constexpr int kMaxTokens = 128;
// heat-map token selection, passed to kernels by value
struct TokenList {
  int n;
  int per_head;  // 0: maps[b', t, n] = mean over heads (hook.py:55); 1: maps[b', head, t, n], no mean (DAAM keeps heads)
  int idx[kMaxTokens];
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// numpy's `(h - mn) / ((mx - mn) + 1e-8f)` in fp32 with IEEE rounding and no FMA contraction
// (data_generation/data_generation.py:82).
__device__ __forceinline__ float np_normalize(float h, float mn, float denom) {
  return __fdiv_rn(__fsub_rn(h, mn), denom);
}
__device__ __forceinline__ float np_denominator(float mn, float mx) {
  return __fadd_rn(__fsub_rn(mx, mn), 1e-8f);
}

}  // namespace agenda
