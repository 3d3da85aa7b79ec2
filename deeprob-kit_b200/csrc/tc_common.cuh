// tc_common.cuh -- device helpers shared by the tcgen05 kernels written in round 2 (mbarrier, bulk copy,
// tcgen05 fences / commit, shared-memory operand layouts).  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpk {
namespace tc {

constexpr uint32_t kSpinLimit = 1u << 26;   // bounded waits: a broken pipeline traps instead of hanging the GPU box

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > kSpinLimit) __trap();
  }
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory stores -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// true in exactly one (elected) lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// K-major operand images with the 64-byte swizzle: rows 64 B apart, 8-row groups 512 B apart (SBO).
// byte offset of 16-byte chunk `c` (0..3) of row `row`
__host__ __device__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t c) {
  return row * 64u + ((c ^ ((row >> 1) & 3u)) << 4);
}
// shared-memory matrix descriptor: high word for SBO = 512 B, descriptor version 1, SWIZZLE_64B;
// low word = (address >> 4) | LBO(1) << 16
constexpr uint32_t kDescHi64 = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return (smem_addr >> 4) | (1u << 16); }

// 128-byte swizzle: rows 128 B apart, 8-row groups 1024 B apart; 16-byte chunk c (0..7) of row `row`
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t c) {
  return row * 128u + ((c ^ (row & 7u)) << 4);
}
constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);

// instruction descriptor: fp32 accumulate, K-major A and B, M = 128, N = n; formats 0 = f16, 1 = bf16, 2 = tf32
__host__ __device__ __forceinline__ uint32_t idesc_m128(uint32_t n, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | (8u << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32 (8 K elements of 32 bits per instruction), one issuing thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi64)
               : "memory");
}

__device__ __forceinline__ void mma_tf32_sw128(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi128)
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 operands, 16 K elements per instruction), 64B-swizzled K-major images
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi64)
               : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMA tensor copy (SASS: UTMALDG): box of a 2-D fp32 tensor described by a host-encoded CUtensorMap -> shared memory,
// completion counted in bytes on `bar`.  c0 = coordinate along the contiguous dimension, c1 = row.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
// exp / log through the SFU without the denormal fix-up sequences of __expf / __logf (results below 2^-126 flush to 0)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_ftz(x * 1.4426950408889634f); }
__device__ __forceinline__ float log_fast(float x) { return lg2_ftz(x) * 0.6931471805599453f; }

// host: encode the tensor map of a row-major 2-D fp32 tensor [rows][cols] (row stride in bytes, multiple of 16) with a
// box of box_rows x box_cols elements, no swizzle.  Resolved through the runtime (no link-time libcuda dependency).
int make_tensor_map_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                           uint32_t box_rows, uint32_t box_cols, int swizzle128 = 0);

}  // namespace tc
}  // namespace dpk
