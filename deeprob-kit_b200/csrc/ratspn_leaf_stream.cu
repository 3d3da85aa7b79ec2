// ratspn_leaf_stream.cu -- RAT-SPN leaf level for NARROW models (G0 * K <= 256 leaf columns), where the path is
// HBM bound instead of tensor bound (R * K <~ 11 columns per feature): x is read from HBM exactly
// once, by TMA, and everything else is sized so that the stream never waits.
//
//   RegionGraphLayer.forward  deeprob/spn/layers/ratspn.py:87-108  (GaussianLayer :160-213 with the frozen unit scale,
//   BernoulliLayer :216-247): out[b, (g,k)] = sum_f x[b,f] W[f,(g,k)] + c[(g,k)]   (+ the sample's -1/2 sum_f x_f^2,
//   which factors out of every sum node and is added at the root, see ratspn_leaf_mma.cu / ratspn_tree_mma.cu)
//
// Same arithmetic as the wide-model GEMM (3-pass hi/lo fp16 on tcgen05, fp32 accumulate in TMEM, exact-kernel redo for
// out-of-range inputs); different data movement.  One CTA per SM, persistent over 64- or 128-sample tiles:
//   warp 0       producer: per 32-feature K block one TMA tensor copy (SASS UTMALDG) of the fp32 x tile
//                [tile rows x 32 features], 128B-swizzled, zero filled past B and D, + two bulk copies (UBLKCP) of the
//                K block's weight images; S stages deep, no register involved
//   warp 1       TMEM allocation + tcgen05.mma issue (one elected thread), two accumulator slots: the epilogue of
//                tile t overlaps the K loop of tile t + 1
//   warps 2-17   converters, all sixteen on every stage: thread = (tile row, quarter of its 32 features): 2 conflict-free
//                LDS.128 of the 128B-swizzled fp32 row, hi/lo fp16 split, 1 + 1 conflict-free STS.128 into the K-major
//                64B-swizzled operand images; running sum of squares and max |x| of the quarter row in one register
//                each, reduced over the row's four lanes at the tile's last K block (fixed order: reproducible).
//                (Groups of warps taking alternate stages would shorten nothing and break the mbarrier parity
//                protocol: a group that skips stages can reach a ring slot a whole phase early.)
//   warps 18-21  epilogue: tcgen05.ld, + per-column constant, coalesced stores to the sample-minor act[0]
// With at most 64 columns a 128 x 64 x 16 MMA takes ~110 cycles when it accumulates onto its predecessor's result
// (measured), 3.5 times its issue cost: the three passes then run as two independent chains -- A_hi against the
// stacked [W_hi; W_lo] image (N = 128, columns 0..127) and A_lo against W_hi (columns 128..191) -- that the epilogue adds.
// The loops carry (stage, phase, K block, tile) incrementally: with a stage every ~700 cycles a 64-bit division per
// role and stage is what bounds the kernel (measured: profiles/leaf_stream_r2.txt).
#include <algorithm>
#include <climits>

#include <cuda_fp16.h>

#include "ratspn_kernels.cuh"
#include "tc_common.cuh"

namespace dpk {

namespace {

using namespace tc;

constexpr int kSThreads = (2 + 16 + 4) * 32;
constexpr int kConvWarps = 16;
constexpr size_t kTailBytes = 256;   // barriers + TMEM address
constexpr uint32_t kRawBytes = 128 * 128;     // fp32 x tile of a stage: 128 rows x 32 features
constexpr uint32_t kAImg = 128 * 64;          // one fp16 operand image: 128 rows x 32 features
constexpr uint32_t kWImg = 256 * 64;          // stride of the hi / lo weight images in global memory (256-column tiles)

struct StreamArgs {
  const unsigned char* wimg;   // [KBn][hi | lo][256 columns x 64 B]   (N tile 0 of ratspn_run_prep_leaf_mma)
  const float* cstm;           // [Ntot]
  float* out;                  // act[0]: element (column c, sample b) at (b >> 7) * out_ts + c * out_cs + (b & 127)
  int64_t out_cs, out_ts;
  float* sqsum;                // [Bp] or NULL: -1/2 sum_f x_f^2
  int* redo;                   // [Bp/32]
  const int* wflag;
  int64_t B, Bp;
  int Ntot, NT, KBn, MT, n_tiles, stages;
  int split;                   // 1: two accumulator chains, 3 * NT columns per slot (NT = 64)
  int dbg;                     // DPK_STREAM_DBG (timing experiments only): 1 = converters skip their work, 2 = epilogue skips its stores
  uint32_t stage_bytes;
  float xlimit;
};

// position of a role in the stream of stages: ring slot + its phase, K block, tile ordinal of this CTA
struct StagePos {
  int s, kb, ti;
  uint32_t ph;
  __device__ __forceinline__ void advance(int n, int S, int KBn) {
    s += n;
    while (s >= S) { s -= S; ph ^= 1u; }
    kb += n;
    while (kb >= KBn) { kb -= KBn; ++ti; }
  }
};

__global__ void __launch_bounds__(kSThreads, 1) ratspn_leaf_stream_kernel(const StreamArgs a, const __grid_constant__ CUtensorMap xmap) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw);
  const int S = a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)S * a.stage_bytes);
  uint64_t* raw_full = bars;             // [S] x tile + weight images landed (tx bytes)
  uint64_t* conv_done = bars + 8;        // [S] operand images written by the four converter warps of the stage
  uint64_t* empty = bars + 16;           // [S] MMAs that read the stage completed
  uint64_t* tfull = bars + 24;           // [2] accumulator slot complete
  uint64_t* tempty = bars + 26;          // [2] accumulator slot drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  const uint32_t slot_cols = (uint32_t)(a.split ? 4 * a.NT : a.NT);   // accumulators on power-of-two column offsets
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_bytes = (uint32_t)a.NT * 64u;

  pdl_launch_dependents();   // persistent, every CTA resident: the next launch may be staged behind this one (common.cuh)
  if (__ldg(a.wflag) != 0) {   // parameters outside the fp16 range: the exact kernel does everything
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.Bp / 32; i += (int64_t)gridDim.x * blockDim.x) a.redo[i] = 1;
    return;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(raw_full + s, 1); mbar_init(conv_done + s, kConvWarps); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
    mbar_init_fence();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows of the operand images past the tile height are never written by the converters: zero them once
  for (uint32_t i = threadIdx.x; i < (uint32_t)S * (2 * kAImg / 16); i += blockDim.x) {
    const uint32_t s = i / (2 * kAImg / 16), o = i % (2 * kAImg / 16);
    *reinterpret_cast<uint4*>(sm + (size_t)s * a.stage_bytes + kRawBytes + o * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int KBn = a.KBn;

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0) {
      const uint32_t tx = (uint32_t)a.MT * 128u + ((a.dbg & 16) ? 0u : 2u * b_bytes);
      for (StagePos p = {0, 0, 0, 0u}; p.ti < my_tiles; p.advance(1, S, KBn)) {
        const int tile = (int)blockIdx.x + p.ti * (int)gridDim.x;
        mbar_wait(empty + p.s, p.ph ^ 1u);
        mbar_expect_tx(raw_full + p.s, tx);
        const uint32_t st = base + (uint32_t)p.s * a.stage_bytes;
        tma_load_2d(st, &xmap, p.kb * 32, tile * a.MT, raw_full + p.s);
        if (!(a.dbg & 16)) {
          const unsigned char* w = a.wimg + (size_t)p.kb * (2 * kWImg);
          bulk_g2s(st + kRawBytes + 2 * kAImg, w, b_bytes, raw_full + p.s);
          bulk_g2s(st + kRawBytes + 2 * kAImg + b_bytes, w + kWImg, b_bytes, raw_full + p.s);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issue ----------------
    const int n16 = (a.Ntot + 15) / 16 * 16;
    const uint32_t idesc = idesc_m128((uint32_t)n16, 0u), idesc2 = idesc_m128(2u * (uint32_t)a.NT, 0u);
    for (StagePos p = {0, 0, 0, 0u}; p.ti < my_tiles; p.advance(1, S, KBn)) {
      const int slot = p.ti & 1;
      if (p.kb == 0 && p.ti >= 2) mbar_wait(tempty + slot, (uint32_t)((p.ti >> 1) - 1) & 1u);
      mbar_wait(conv_done + p.s, p.ph);
      fence_after();
      const uint32_t st = base + (uint32_t)p.s * a.stage_bytes;
      const uint32_t a_hi = desc_lo(st + kRawBytes), a_lo = a_hi + (kAImg >> 4);
      const uint32_t b_hi = desc_lo(st + kRawBytes + 2 * kAImg), b_lo = b_hi + (b_bytes >> 4);
      const uint32_t d = tmem + (uint32_t)slot * slot_cols;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < ((a.dbg & 8) ? 0 : 2); ++ks) {
          const uint32_t acc = (p.kb > 0 || ks > 0) ? 1u : 0u;
          if (a.split) {
            mma_f16(d, a_hi + 2 * ks, b_hi + 2 * ks, idesc2, acc);                       // x_hi * [W_hi; W_lo]
            mma_f16(d + 2u * (uint32_t)a.NT, a_lo + 2 * ks, b_hi + 2 * ks, idesc, acc);  // x_lo * W_hi
          } else {
            mma_f16(d, a_hi + 2 * ks, b_hi + 2 * ks, idesc, acc);
            mma_f16(d, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
            mma_f16(d, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
          }
        }
        commit(empty + p.s);
        if (p.kb == KBn - 1) commit(tfull + slot);
      }
      __syncwarp();
    }
  } else if (warp < 2 + kConvWarps) {
    // ---------------- converters: thread = (tile row, quarter = 8 features) ----------------
    const int t = threadIdx.x - 64;
    const uint32_t row = (uint32_t)t >> 2, q = (uint32_t)t & 3u;
    const bool live = (int)row < a.MT;
    float sq = 0.f;
    uint32_t umax = 0u;
    for (StagePos p = {0, 0, 0, 0u}; p.ti < my_tiles; p.advance(1, S, KBn)) {
      mbar_wait(raw_full + p.s, p.ph);
      unsigned char* st = sm + (size_t)p.s * a.stage_bytes;
      if (live && !(a.dbg & 1)) {
        const unsigned char* rp = st + row * 128u;
        const float4 v0 = *reinterpret_cast<const float4*>(rp + (((2u * q) ^ (row & 7u)) << 4));
        const float4 v1 = *reinterpret_cast<const float4*>(rp + (((2u * q + 1u) ^ (row & 7u)) << 4));
        umax = max(umax, max(max(max(__float_as_uint(v0.x) & 0x7fffffffu, __float_as_uint(v0.y) & 0x7fffffffu),
                                 max(__float_as_uint(v0.z) & 0x7fffffffu, __float_as_uint(v0.w) & 0x7fffffffu)),
                             max(max(__float_as_uint(v1.x) & 0x7fffffffu, __float_as_uint(v1.y) & 0x7fffffffu),
                                 max(__float_as_uint(v1.z) & 0x7fffffffu, __float_as_uint(v1.w) & 0x7fffffffu))));
        sq = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, sq))));
        sq = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, sq))));
        const __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
        const __half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
        const __half2 l0 = __floats2half2_rn(v0.x - f0.x, v0.y - f0.y), l1 = __floats2half2_rn(v0.z - f1.x, v0.w - f1.y);
        const __half2 l2 = __floats2half2_rn(v1.x - f2.x, v1.y - f2.y), l3 = __floats2half2_rn(v1.z - f3.x, v1.w - f3.y);
        const uint32_t off = sw64_off(row, q);
        *reinterpret_cast<uint4*>(st + kRawBytes + off) =
            make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                       *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
        *reinterpret_cast<uint4*>(st + kRawBytes + kAImg + off) =
            make_uint4(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1),
                       *reinterpret_cast<const uint32_t*>(&l2), *reinterpret_cast<const uint32_t*>(&l3));
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_done + p.s);
      if (p.kb == KBn - 1) {   // the tile's last K block: the row's four quarters are neighbouring lanes
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, 1));
        umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, 2));
        const int64_t b = ((int64_t)blockIdx.x + (int64_t)p.ti * gridDim.x) * a.MT + row;
        if (live && q == 0u && b < a.B) {
          if (a.sqsum) a.sqsum[b] = -0.5f * sq;
          if (umax > __float_as_uint(a.xlimit)) a.redo[b >> 5] = 1;
        }
        sq = 0.f; umax = 0u;
      }
    }
  } else {
    // ---------------- epilogue: lanes = samples (TMEM lanes), 32 columns per TMEM load ----------------
    const int q = warp & 3;                      // TMEM lane quarter of this warp
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int slot = ti & 1;
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      mbar_wait(tfull + slot, (uint32_t)(ti >> 1) & 1u);
      fence_after();
      const int r = q * 32 + lane;
      const int64_t b = (int64_t)tile * a.MT + r;
      const bool ok = r < a.MT && b < a.Bp;
      if (q * 32 < a.MT) {
        for (int c0 = 0; c0 < a.Ntot; c0 += 32) {
          uint32_t v[32];
          const uint32_t tcol = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * slot_cols + (uint32_t)c0;
          tmem_ld32(tcol, v);
          tmem_ld_wait();
          if (a.split) {
            uint32_t u[32];
            tmem_ld32(tcol + (uint32_t)a.NT, u);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
            tmem_ld32(tcol + 2u * (uint32_t)a.NT, u);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
          }
          if (ok && !(a.dbg & 2)) {
            float* op = a.out + (b >> 7) * a.out_ts + (size_t)c0 * a.out_cs + (b & 127);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (c0 + i < a.Ntot) __stcs(op, __uint_as_float(v[i]) + __ldg(a.cstm + c0 + i));
              op += a.out_cs;
            }
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + slot);
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace

int ratspn_run_leaf_stream(const RatPlan& p, const float* x, float* ws, cudaStream_t st) {
  StreamArgs a;
  a.wimg = reinterpret_cast<const unsigned char*>(ws + p.off_wimg);
  a.cstm = ws + p.off_cstm;
  a.out = ws + p.off_act[0];
  a.out_cs = p.act0_cs; a.out_ts = p.act0_ts;
  a.sqsum = p.off_sqsum ? ws + p.off_sqsum : nullptr;
  a.redo = reinterpret_cast<int*>(ws + p.off_mflags);
  a.wflag = a.redo + p.Bp / 32 + 3;
  a.B = p.B; a.Bp = p.Bp;
  a.Ntot = p.G0 * p.K;
  a.NT = a.Ntot <= 64 ? 64 : a.Ntot <= 128 ? 128 : 256;
  a.KBn = (int)ceil_div(p.D, 32);
  a.split = (a.NT == 64 && env_int("DPK_STREAM_SPLIT", 1) != 0) ? 1 : 0;
  const int nsm = sm_count();
  // Tile height: a multiple of 16 rows (the MMAs are M = 128 whatever it is -- the tensor pipe is mostly idle -- and rows
  // past it stay zero) picked for the fewest rounds over the SMs, counting ~32 rows' worth of fixed cost per stage:
  // 65536 samples on 148 SMs = 4 rounds of 112-row tiles (586 tiles) instead of 3.46 -> 4 rounds of 128-row ones.
  {
    int64_t best = INT64_MAX;
    a.MT = 128;
    for (int mt = 128; mt >= 64; mt -= 16) {
      const int64_t cost = ceil_div(ceil_div(p.B, mt), nsm) * (mt + 32);
      if (cost < best) { best = cost; a.MT = mt; }
    }
    const int knob = env_int("DPK_STREAM_MT", 0);
    if (knob >= 16 && knob <= 128 && knob % 16 == 0) a.MT = knob;
  }
  a.n_tiles = (int)ceil_div(p.B, a.MT);
  a.stage_bytes = kRawBytes + 2 * kAImg + 2u * (uint32_t)a.NT * 64u;
  const size_t smem_max = (size_t)max_dynamic_smem();
  a.stages = (int)std::min<size_t>(7, (smem_max - 1024 - kTailBytes) / a.stage_bytes);
  a.xlimit = 30000.f;
  a.dbg = env_int("DPK_STREAM_DBG", 0);
  a.stages = std::min(a.stages, std::max(2, env_int("DPK_STREAM_STAGES", 7)));
  if (a.stages < 2) return set_error(DPK_E_ARG, "leaf stream kernel: shared memory too small");
  DPK_CUDA_TRY(cudaMemsetAsync(a.redo, 0, ((size_t)p.Bp / 32 + 3) * 4, st));
  alignas(64) CUtensorMap xmap;
  int rc = make_tensor_map_2d_f32(&xmap, x, (uint64_t)p.B, (uint64_t)p.D, (uint64_t)p.D * 4, (uint32_t)a.MT, 32, 1);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)a.stages * a.stage_bytes + kTailBytes;
  DPK_CUDA_TRY(cudaFuncSetAttribute(ratspn_leaf_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(CAT_LEAF_MMA, st);
  const int grid = std::max(1, std::min(std::min(nsm, a.n_tiles), env_int("DPK_STREAM_GRID", 1 << 30)));   // knob: tests force many tiles per CTA
  ratspn_leaf_stream_kernel<<<grid, kSThreads, smem, st>>>(a, xmap);
  DPK_LAUNCH_CHECK("ratspn_leaf_stream_kernel");
  return DPK_OK;
}

}  // namespace dpk
