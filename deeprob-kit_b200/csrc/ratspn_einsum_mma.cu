// ratspn_einsum_mma.cu -- RAT-SPN product + sum level with the mixture contraction on the tensor cores.
//
//   ProductLayer.forward  deeprob/spn/layers/ratspn.py:272-286   x1[..., :, None] + x2[..., None, :]
//   SumLayer.forward      deeprob/spn/layers/ratspn.py:363-378   logsumexp(x + log_softmax(W))
// in the einsum form of ratspn_einsum.cu,
//   y[b,p,o] = ml + mr + log sum_i e^{l_i - ml} * ( sum_j softmax(W)[p,o,i,j] * e^{r_j - mr} ),
// where the inner sum -- 1000 of the 1100 multiply-adds per (sample, partition) at K = O = 10 -- is the GEMM
//   T[b, (o,i)] = sum_j er[b, j] * W_p[(o,i), j]          M = 128 samples, N = O*K (100 -> 112), K = K_in (10 -> 16)
// issued as tcgen05.mma with the accumulator in TMEM; the CUDA cores keep only the 2K exps, the K*O-term
// finish sum_i el_i T[o,i] and the logs.
//
// fp32 accuracy on fp16 tensor cores: er in (0,1] is scaled by 2^14, softmax(W) in [0,1] by 2^15 (both stay inside
// the fp16 range, and the subnormal floor drops to ~1e-12), each is split v = hi + lo (22 bits) and the product
// is taken in three passes hi*hi + lo*hi + hi*lo with fp32 accumulation.  A linear-domain sum below 1e-5 (where
// the split no longer guarantees 1e-5 relative accuracy) or non-finite takes the exact log-domain path of
// ratspn_einsum.cu, like every other underflow there.
//
// Kernel shape: CTA = 128 samples x one partition p, looping over sample tiles; 4 worker warps (thread = sample =
// TMEM lane) + 1 warp that loads the partition's weight image once (TMA bulk copy) and issues the three MMAs of
// every tile from one elected thread.  Four CTAs share an SM (128 TMEM columns each), so the load -> exp ->
// stage -> MMA -> drain chain of one CTA overlaps the others'.  Operand images are K-major with the 32-byte
// swizzle (one 16-element K step = one 32-byte row).
#include <cuda_fp16.h>

#include <algorithm>

#include "ratspn_kernels.cuh"

namespace dpk {

namespace {

constexpr int kEmThreads = 160;           // 4 worker warps + 1 MMA warp
constexpr int kEmTile = 128;              // samples per tile = TMEM lanes
constexpr int kEmRowBytes = 32;           // 16 fp16 per operand row
constexpr float kScaleA = 16384.f;        // 2^14
constexpr float kScaleB = 32768.f;        // 2^15
constexpr float kLogScale = 29.f * 0.693147180559945309417f;   // log(2^29)
constexpr float kMinSum = 1e-5f * 536870912.f;                 // 1e-5 in the scaled domain
constexpr uint32_t kEmSpinLimit = 1u << 26;

struct EinsumMmaArgs {
  const float* in;             // [2P][Kin][Bp]
  const unsigned char* wimg;   // [P][hi | lo][Npad rows x 32 B], 32B-swizzled
  const float* wsoft;          // [P][1][Kin2][OC] softmax (fp32 re-evaluation of doubtful outputs)
  const float* wlog;           // [P][1][Kin2][OC] log-softmax (exact fallback)
  float* out;                  // [P][O][Bp]
  int64_t Bp;
  int P, O, OC, Npad, n_tiles, tile_stride;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
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
    if (spin > kEmSpinLimit) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// descriptor high word: SBO = 256 B (8 rows x 32 B), version 1, SWIZZLE_32B; low word = (addr >> 4) | 1 << 16
constexpr uint32_t kEmDescHi = (256u >> 4) | (1u << 14) | (6u << 29);
__device__ __forceinline__ void em_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kEmDescHi)
               : "memory");
}
__device__ __forceinline__ void em_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// byte offset of 16-byte chunk c (0/1) of row r inside a 32B-swizzled image (Swizzle<1,4,3>: bit 4 ^= bit 7)
__host__ __device__ __forceinline__ uint32_t sw32_off(uint32_t r, uint32_t c) { return r * 32u + ((c ^ ((r >> 2) & 1u)) << 4); }

// exact log-domain value of one output (same as ratspn_einsum.cu): logsumexp_ij(l_i + r_j + logw[ij])
__device__ __noinline__ float einsum_exact_mma(const float* __restrict__ l, const float* __restrict__ r, int64_t stride,
                                               int Kin, const float* __restrict__ wlog, int OC) {
  float m = -INFINITY;
  for (int i = 0; i < Kin; ++i)
    for (int j = 0; j < Kin; ++j) m = fmaxf(m, l[i * stride] + r[j * stride] + wlog[(size_t)(i * Kin + j) * OC]);
  if (!(fabsf(m) <= FLT_MAX)) return m;
  float s = 0.f;
  for (int i = 0; i < Kin; ++i)
    for (int j = 0; j < Kin; ++j) s += expf(l[i * stride] + r[j * stride] + wlog[(size_t)(i * Kin + j) * OC] - m);
  return m + logf(s);
}

// fp32 linear-domain value of one output for a sample whose tensor-core sum was too small to trust (same
// arithmetic as ratspn_einsum_reg_kernel); the exact log-domain path only if that underflows as well.
template <int KIN>
__device__ __noinline__ float einsum_careful(const float* __restrict__ l, const float* __restrict__ r, int64_t stride,
                                             const float* __restrict__ wsoft, const float* __restrict__ wlog, int OC) {
  float lv[KIN], rv[KIN], vl = -INFINITY, vr = -INFINITY;
#pragma unroll
  for (int k = 0; k < KIN; ++k) {
    lv[k] = l[k * stride]; rv[k] = r[k * stride];
    vl = fmaxf(vl, lv[k]); vr = fmaxf(vr, rv[k]);
  }
  const float ml = (fabsf(vl) <= FLT_MAX) ? vl : 0.f, mr = (fabsf(vr) <= FLT_MAX) ? vr : 0.f;
#pragma unroll
  for (int k = 0; k < KIN; ++k) { lv[k] = __expf(lv[k] - ml); rv[k] = __expf(rv[k] - mr); }
  float s = 0.f;
  for (int i = 0; i < KIN; ++i) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < KIN; ++j) t = fmaf(__ldg(wsoft + (size_t)(i * KIN + j) * OC), rv[j], t);
    s = fmaf(lv[i], t, s);
  }
  if (s >= 1e-18f && s <= FLT_MAX) return ml + mr + __logf(s);
  return einsum_exact_mma(l, r, stride, KIN, wlog, OC);
}

// Two tiles are in flight per CTA (two A buffers, two accumulator slots of 128 TMEM columns):
//   iteration k:  stage tile k+1 (exps of the values loaded one iteration ago) -> its MMAs run during ...
//                 issue the loads of tile k+2 ... and ...
//                 wait for the accumulator of tile k, finish sum_i el_i T[o,i], log, store.
template <int KIN>
__global__ void __launch_bounds__(kEmThreads) ratspn_einsum_mma_kernel(const EinsumMmaArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 255u) & ~255u;     // 32B-swizzled images repeat every 256 bytes
  unsigned char* sm = smem_raw + (base - raw);
  // layout: 2 x {A hi (4 KB) | A lo (4 KB)} | B hi (Npad*32) | B lo (Npad*32) | barriers | tmem slot
  constexpr uint32_t kABuf = 2 * kEmTile * kEmRowBytes;
  const uint32_t offBhi = 2 * kABuf, offBlo = offBhi + a.Npad * kEmRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + offBlo + a.Npad * kEmRowBytes);
  uint64_t* b_full = bars;       // weight image landed
  uint64_t* a_full = bars + 1;   // [2] A staged by the 4 worker warps
  uint64_t* d_full = bars + 3;   // [2] accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x;
  const int n_mine = (a.n_tiles - (int)blockIdx.y + a.tile_stride - 1) / a.tile_stride;   // tiles of this CTA

  if (threadIdx.x == 0) {
    mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(a_full + s, 4); mbar_init(d_full + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ---------------- weight image + MMA issue ----------------
    const uint32_t wbytes = 2u * a.Npad * kEmRowBytes;
    if (lane == 0) {
      mbar_expect_tx(b_full, wbytes);
      bulk_g2s(base + offBhi, a.wimg + (size_t)p * wbytes, wbytes, b_full);
    }
    mbar_wait(b_full, 0);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(a.Npad >> 3) << 17) | (8u << 24);   // fp32 accum, fp16, K-major, M=128
    const uint32_t b_hi = ((base + offBhi) >> 4) | (1u << 16), b_lo = b_hi + (a.Npad * kEmRowBytes / 16);
    for (int k = 0; k < n_mine; ++k) {
      const int s = k & 1;
      mbar_wait(a_full + s, (uint32_t)(k >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = ((base + s * kABuf) >> 4) | (1u << 16), a_lo = a_hi + (kEmTile * kEmRowBytes / 16);
      const uint32_t d = tmem + s * 128;
      if (elect_one()) {
        em_mma(d, a_hi, b_hi, idesc, 0u);
        em_mma(d, a_lo, b_hi, idesc, 1u);
        em_mma(d, a_hi, b_lo, idesc, 1u);
        em_commit(d_full + s);
      }
      __syncwarp();
    }
  } else {
    // ---------------- workers: thread = sample = TMEM lane ----------------
    const int tid = threadIdx.x;                 // 0..127
    const float* __restrict__ lin = a.in + (size_t)(2 * p) * KIN * a.Bp;
    const float* __restrict__ rin = lin + (size_t)KIN * a.Bp;
    const float* __restrict__ wlog = a.wlog + (size_t)p * KIN * KIN * a.OC;
    const float* __restrict__ wsoft = a.wsoft + (size_t)p * KIN * KIN * a.OC;
    float* __restrict__ outp = a.out + (size_t)p * a.O * a.Bp;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);

    float raw_l[KIN], raw_r[KIN];                // values of the tile that is staged next
    auto load = [&](int k) {
      const int64_t b = ((int64_t)blockIdx.y + (int64_t)k * a.tile_stride) * kEmTile + tid;
#pragma unroll
      for (int i = 0; i < KIN; ++i) {
        raw_l[i] = __ldcs(lin + (size_t)i * a.Bp + b);
        raw_r[i] = __ldcs(rin + (size_t)i * a.Bp + b);
      }
    };
    // exps of the loaded tile: el (kept for the finish) and the staged, scaled hi/lo rows of er
    auto stage = [&](int k, float (&el)[KIN], float* shift) {
      float vl = -INFINITY, vr = -INFINITY;
#pragma unroll
      for (int i = 0; i < KIN; ++i) { vl = fmaxf(vl, raw_l[i]); vr = fmaxf(vr, raw_r[i]); }
      // a fully -inf (dropped-out) or non-finite side: shift by 0, the careful path sorts it out
      const float ml = (fabsf(vl) <= FLT_MAX) ? vl : 0.f;
      const float mr = (fabsf(vr) <= FLT_MAX) ? vr : 0.f;
      *shift = ml + mr - kLogScale;
      float er[KIN];
#pragma unroll
      for (int i = 0; i < KIN; ++i) {
        el[i] = __expf(raw_l[i] - ml);
        er[i] = __expf(raw_r[i] - mr) * kScaleA;
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float x0 = (2 * q < KIN) ? er[2 * q < KIN ? 2 * q : 0] : 0.f;
        const float x1 = (2 * q + 1 < KIN) ? er[2 * q + 1 < KIN ? 2 * q + 1 : 0] : 0.f;
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 f = __half22float2(h);
        const __half2 l2 = __floats2half2_rn(x0 - f.x, x1 - f.y);
        hi[q] = *reinterpret_cast<const uint32_t*>(&h);
        lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      unsigned char* A = sm + (k & 1) * kABuf;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t off = sw32_off((uint32_t)tid, (uint32_t)c);
        *reinterpret_cast<uint4*>(A + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<uint4*>(A + kEmTile * kEmRowBytes + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // this slot's previous TMEM reads are done
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + (k & 1));
    };

    float el_cur[KIN], el_nxt[KIN], shift_cur = 0.f, shift_nxt = 0.f;
    if (n_mine > 0) { load(0); stage(0, el_cur, &shift_cur); }
    if (n_mine > 1) load(1);
    for (int k = 0; k < n_mine; ++k) {
      if (k + 1 < n_mine) stage(k + 1, el_nxt, &shift_nxt);   // its MMAs overlap the finish of tile k
      if (k + 2 < n_mine) load(k + 2);                        // lands while tile k is finished
      const int64_t b = ((int64_t)blockIdx.y + (int64_t)k * a.tile_stride) * kEmTile + tid;
      mbar_wait(d_full + (k & 1), (uint32_t)(k >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tbase = trow + (uint32_t)((k & 1) * 128);
      // two outputs per round trip: independent TMEM loads and dot products
#pragma unroll 1
      for (int o = 0; o < a.O; o += 2) {
        const bool two = o + 1 < a.O;
        uint32_t r0[16], r1[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]),
              "=r"(r0[8]), "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
            : "r"(tbase + (uint32_t)(o * KIN))
            : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]),
              "=r"(r1[8]), "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
            : "r"(tbase + (uint32_t)((two ? o + 1 : o) * KIN))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;   // two partial sums per output
#pragma unroll
        for (int i = 0; i < KIN; i += 2) {
          s0a = fmaf(el_cur[i], __uint_as_float(r0[i]), s0a);
          s1a = fmaf(el_cur[i], __uint_as_float(r1[i]), s1a);
          if (i + 1 < KIN) {
            s0b = fmaf(el_cur[i + 1 < KIN ? i + 1 : 0], __uint_as_float(r0[i + 1]), s0b);
            s1b = fmaf(el_cur[i + 1 < KIN ? i + 1 : 0], __uint_as_float(r1[i + 1]), s1b);
          }
        }
        const float s0 = s0a + s0b, s1 = s1a + s1b;
        float y0, y1;
        if (s0 >= kMinSum && s0 <= FLT_MAX) y0 = shift_cur + __logf(s0);
        else y0 = einsum_careful<KIN>(lin + b, rin + b, a.Bp, wsoft + o, wlog + o, a.OC);
        outp[(size_t)o * a.Bp + b] = y0;
        if (two) {
          if (s1 >= kMinSum && s1 <= FLT_MAX) y1 = shift_cur + __logf(s1);
          else y1 = einsum_careful<KIN>(lin + b, rin + b, a.Bp, wsoft + o + 1, wlog + o + 1, a.OC);
          outp[(size_t)(o + 1) * a.Bp + b] = y1;
        }
      }
#pragma unroll
      for (int i = 0; i < KIN; ++i) el_cur[i] = el_nxt[i];
      shift_cur = shift_nxt;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

// weight images: row n = o*Kin + i, column j: softmax(W)[p,o,i*Kin+j] * 2^15 as hi/lo fp16
__global__ void ratspn_prep_einsum_mma_kernel(const float* __restrict__ wsoft, int P, int O, int Kin, int OC, int Npad,
                                              unsigned char* __restrict__ wimg) {
  const int64_t total = (int64_t)P * Npad * 16;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 15);
    const int n = (int)((idx >> 4) % Npad);
    const int p = (int)(idx / ((int64_t)Npad * 16));
    const int o = n / Kin, i = n - o * Kin;
    float w = 0.f;
    if (o < O && j < Kin) w = wsoft[((size_t)p * Kin * Kin + (size_t)(i * Kin + j)) * OC + o] * kScaleB;
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    unsigned char* img = wimg + (size_t)p * 2 * Npad * kEmRowBytes;
    const uint32_t off = sw32_off((uint32_t)n, (uint32_t)j >> 3) + (uint32_t)(j & 7) * 2u;
    *reinterpret_cast<__half*>(img + off) = hi;
    *reinterpret_cast<__half*>(img + (size_t)Npad * kEmRowBytes + off) = lo;
  }
}

template <int KIN>
int launch_einsum_mma_t(const EinsumMmaArgs& a, int cat, cudaStream_t st) {
  auto kern = ratspn_einsum_mma_kernel<KIN>;
  // 112 KB of dynamic shared memory per CTA caps the residency at 2 CTAs per SM = 2 x 256 TMEM columns
  const size_t smem = 112 * 1024;
  DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)a.P, (unsigned)a.tile_stride);
  ProfScope prof(cat, st);
  kern<<<grid, kEmThreads, smem, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_einsum_mma_kernel");
  return DPK_OK;
}

}  // namespace

bool ratspn_einsum_mma_eligible(int Kin, int O, int nOc, int64_t Bp) {
  // Opt-in (DPK_EINSUM_MMA=1 for any batch, =2 for batches >= 8192): measured on B200 at config 2 it ties with the
  // packed-FFMA2 kernel (0.40 vs 0.41 ms for both levels) -- with 2K exps, the K*O finish, K logs and the operand
  // split left on the CUDA cores and only 4 accumulator slots per SM, the level is latency bound, not FMA bound.
  const int knob = env_int("DPK_EINSUM_MMA", 0);
  if (knob <= 0 || nOc != 1) return false;
  if (!(Kin == 2 || Kin == 4 || Kin == 8 || Kin == 10 || Kin == 16)) return false;
  if ((O - 1) * Kin + 16 > 128) return false;                 // 16-column TMEM reads stay inside the allocation
  return knob == 1 || Bp >= 8192;
}

size_t ratspn_einsum_mma_image_floats(int P, int Kin, int O) {
  const int Npad = (O * Kin + 15) / 16 * 16;
  return (size_t)P * 2 * Npad * kEmRowBytes / 4;
}

int ratspn_run_prep_einsum_mma(const float* wsoft, int P, int O, int Kin, int OC, float* wimg, cudaStream_t st) {
  const int Npad = (O * Kin + 15) / 16 * 16;
  const int64_t total = (int64_t)P * Npad * 16;
  ratspn_prep_einsum_mma_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 2048), 256, 0, st>>>(
      wsoft, P, O, Kin, OC, Npad, reinterpret_cast<unsigned char*>(wimg));
  DPK_LAUNCH_CHECK("ratspn_prep_einsum_mma_kernel");
  return DPK_OK;
}

int ratspn_run_einsum_mma(const float* in, const float* wimg, const float* wsoft, const float* wlog, float* out, int64_t Bp, int P, int Kin,
                          int O, int OC, int cat, cudaStream_t st) {
  EinsumMmaArgs a;
  a.in = in; a.wimg = reinterpret_cast<const unsigned char*>(wimg); a.wsoft = wsoft; a.wlog = wlog; a.out = out;
  a.Bp = Bp; a.P = P; a.O = O; a.OC = OC;
  a.Npad = (O * Kin + 15) / 16 * 16;
  a.n_tiles = (int)(Bp / kEmTile);
  // 2 CTAs per SM, every CTA loops over the tiles t = blockIdx.y, blockIdx.y + stride, ... of its partition
  // (rounded down: one wave -- a second, partial wave would take as long as the first)
  a.tile_stride = (int)std::max<int64_t>(1, std::min<int64_t>(a.n_tiles, 2 * (int64_t)sm_count() / P));
  switch (Kin) {
    case 2: return launch_einsum_mma_t<2>(a, cat, st);
    case 4: return launch_einsum_mma_t<4>(a, cat, st);
    case 8: return launch_einsum_mma_t<8>(a, cat, st);
    case 10: return launch_einsum_mma_t<10>(a, cat, st);
    case 16: return launch_einsum_mma_t<16>(a, cat, st);
  }
  return set_error(DPK_E_ARG, "unsupported Kin %d for the tensor-core einsum", Kin);
}

}  // namespace dpk
