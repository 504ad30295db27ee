// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and UMMA descriptor builders shared by the sm_100a attention kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace agenda {
namespace sm100 {

// ------------------------------------------------------------------ PTX wrappers ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {  // required before re-initialising a live barrier
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking probe
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// true on exactly one lane of a fully converged warp (the compiler then knows the guarded region is single-threaded
// and keeps MMA/TMA operands in uniform registers instead of emitting per-thread R2UR loops)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// bring a tile into L2 only (no shared-memory destination, no barrier): hides HBM latency of a later tma_load_4d
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread t of the warp gets TMEM lane (base_lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
// one 32-bit column: thread t of the warp gets TMEM lane (base_lane + t), column `col` of taddr
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t u;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(u) : "r"(taddr) : "memory");
  return __uint_as_float(u);
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]),
      "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2(float x) {
#ifdef AGENDA_DEBUG_NO_MUFU  // tools/ubench experiments only: takes the MUFU out of the softmax to expose other limits
  return x * 0.001f;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp field layout) ----
// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version
  d |= static_cast<uint64_t>(2) << 61;  // LayoutType::SWIZZLE_128B
  return d;
}
// instruction descriptor: kind::f16, A/B = BF16, D = F32, M x N, A K-major, B K-major (b_mn=0) or MN-major (b_mn=1)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ---- packed fp32 pairs (Blackwell f32x2 ALU forms: one issue slot for two lanes' worth of work) and 3-input max ----
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// ---- softmax pieces shared by the self-attention kernels ----
constexpr float kV2RescaleThreshold = 8.0f;  // lazy O rescale: only when the row maximum grows by more than 2^8

// 2^x for a pair on the FMA/ALU pipes instead of the MUFU (FlashAttention-4's trick: at small head dims the
// 16 exp/clk/SM MUFU is the kernel's bottleneck, so a fraction of the exponentials is moved to idle FMA lanes).
// x = s * scale - m is never formed: u = sat(s * a + b) with a = scale / 253, b = (126 - m) / 253 maps x in
// [-126, 127] onto [0, 1], so ONE saturating FMA per score scales, shifts and clamps on both sides (below -126 the
// exponent add would wrap; 127 is the largest finite exponent, and a clamped 2^127 is caught by the row-sum check of
// the fast path).  Then x = n + f with n = round(x), f in [-0.5, 0.5]: 2^f by a cubic minimax polynomial (max rel.
// error 7.5e-5, far below the bf16 rounding of P), 2^n by adding n to the exponent field.  253 * u carries an
// absolute error <= 253 * 2^-25 = 7.5e-6 in x.  (NaN scores saturate to 0 like they did under fmaxf.)
struct Ex2Emu {
  uint64_t k253_2, magic_lo2, c0_2, c1_2, c2_2, c3_2;
  __device__ __forceinline__ Ex2Emu() {
    k253_2 = pack_f32x2(253.f, 253.f);
    magic_lo2 = pack_f32x2(12582912.f - 126.f, 12582912.f - 126.f);
    c0_2 = pack_f32x2(0.9999280572f, 0.9999280572f);
    c1_2 = pack_f32x2(0.6932609677f, 0.6932609677f);
    c2_2 = pack_f32x2(0.2426111251f, 0.2426111251f);
    c3_2 = pack_f32x2(0.0551716648f, 0.0551716648f);
  }
  __device__ __forceinline__ void operator()(float s0, float s1, float a, float b, float& r0, float& r1) const {
    float u0, u1;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u0) : "f"(s0), "f"(a), "f"(b));
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u1) : "f"(s1), "f"(a), "f"(b));
    const uint64_t u2 = pack_f32x2(u0, u1);
    const uint64_t t2 = ffma2(u2, k253_2, magic_lo2);  // x + 1.5 * 2^23: low mantissa bits = n = round(x)
    const uint64_t g2 = fsub2(magic_lo2, t2);          // -(n + 126)
    const uint64_t f2 = ffma2(u2, k253_2, g2);         // x - n
    uint64_t p2 = ffma2(c3_2, f2, c2_2);
    p2 = ffma2(p2, f2, c1_2);
    p2 = ffma2(p2, f2, c0_2);
    float t0, t1, p0, p1;
    unpack_f32x2(t2, t0, t1);
    unpack_f32x2(p2, p0, p1);
    r0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    r1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
  }
};

}  // namespace sm100

// [B, rows, H*d] bf16 viewed as (d, H, rows, B); box (64, 1, box_rows, 1), 128-byte swizzle, zero fill out of bounds.
void bind_primary_context();

// Descriptor cache (the ABI's one piece of state: "per-device TMA descriptor caches").  A tensor map is a
// pure function of (address, element type, dims, strides, box, swizzle): the calling thread keeps its most recent ones,
// so a processor that is called with the same buffers step after step encodes each map once (cuTensorMapEncodeTiled costs
// about a microsecond; an eager — not graph-captured — denoising step builds ~100 of them).
struct TensorMapKey {
  const void* base;
  unsigned long long dims[4], strides[3];
  unsigned box[4];
  int dtype, rank, swizzle;
};
bool tensor_map_cache_get(const TensorMapKey& key, CUtensorMap* out);
void tensor_map_cache_put(const TensorMapKey& key, const CUtensorMap& map);
// row_stride (elements) = distance between consecutive rows; 0 = packed rows of H*d (q/k/v as column slices of one
// fused-projection output pass the row length of that output).
int make_head_map(CUtensorMap* map, const void* base, int B, int H, int rows, int d, int box_rows, long long row_stride = 0);

}  // namespace agenda
