// ratspn_leaf.cu -- RAT-SPN leaf level: RegionGraphLayer.forward
// (deeprob/spn/layers/ratspn.py:87-108; GaussianLayer :160-213, BernoulliLayer :216-247).
// The reference gathers x[:, mask] into (B,G0,1,dim), evaluates the broadcast log-density
// (B,G0,K,dim), nan_to_num's it and sums the last axis; here a transposed x tile is staged once in
// shared memory and every region is swept over it without materialising anything.
#include <algorithm>

#include "ratspn_kernels.cuh"

namespace dpk {

struct LeafTabGeom {
  int CH, NCH, CHP, NPK, CF, TB, NST;  // see RatPlan::leaf_*
};

// parameters of table row d of a (region, channel-chunk) block
__device__ __forceinline__ const float* leaf_tab_row(const float* block, int d, const LeafTabGeom& g) {
  const int ch = d / g.CH;
  return block + (size_t)ch * g.CF + g.CHP + (d - ch * g.CH) * g.NPK;
}

template <int KIND>
__global__ void ratspn_prep_leaf_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                        const int32_t* __restrict__ mask, const int32_t* __restrict__ region_len,
                                        int G0, int K, int dim, int KC, int nKc, LeafTabGeom g,
                                        float* __restrict__ tab, float* __restrict__ cd) {
  const int Kp = KC * nKc;
  const int dimp = g.NCH * g.CH;  // rows incl. the padding of the last chunk
  const int64_t total = (int64_t)G0 * Kp * dimp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(idx % dimp);
    const int kk = (int)((idx / dimp) % Kp);
    const int r = (int)(idx / ((int64_t)dimp * Kp));
    const int c = kk / KC, k = kk % KC;
    const bool live = kk < K && d < region_len[r];   // pad dims / pad channels contribute exactly 0
    const size_t src = ((size_t)r * K + kk) * dim + d;
    const int ch = d / g.CH, w = d - ch * g.CH;
    float* chunk = tab + (((size_t)r * nKc + c) * g.NCH + ch) * g.CF;
    float* trow = chunk + g.CHP + w * g.NPK;
    if (k == 0) {
      const int f = (d < dim) ? mask[(size_t)r * dim + d] : 0;
      chunk[w] = __uint_as_float((uint32_t)f * (uint32_t)g.TB * 4u + ((uint32_t)f & 31u) * 4u);
      if (w == 0) for (int z = g.CH; z < g.CHP; ++z) chunk[z] = 0.f;
      const int used = (KIND == DPK_LEAF_GAUSSIAN) ? 2 * KC : KC;
      for (int z = used; z < g.NPK; ++z) trow[z] = 0.f;
    }
    float v0 = 0.f, v1 = 0.f, cdv = 0.f;
    if (KIND == kLeafGaussUnit) {
      if (live) { v0 = -p0[src]; cdv = -kLogSqrt2Pi; }   // scale == 1: t = x - mu
      trow[k] = v0;
    } else if (KIND == DPK_LEAF_GAUSSIAN) {
      if (live) {
        const float sigma = p1[src], mu = p0[src];
        v0 = 1.0f / sigma;
        v1 = -mu * v0;
        cdv = -logf(sigma) - kLogSqrt2Pi;
      }
      trow[k] = v0;
      trow[KC + k] = v1;
    } else {
      if (live) {
        v0 = p0[src];
        cdv = -(fmaxf(v0, 0.f) + log1pf(expf(-fabsf(v0))));  // -softplus(logit)
      }
      trow[k] = v0;
    }
    if (d < dim) cd[(((size_t)r * nKc + c) * dim + d) * KC + k] = cdv;
  }
}

__global__ void ratspn_prep_const_kernel(const float* __restrict__ cd, const int32_t* __restrict__ region_len,
                                         int G0, int dim, int KC, int nKc, float* __restrict__ cst) {
  const int Kp = KC * nKc;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G0 * Kp) return;
  const int g = idx / Kp, kk = idx % Kp, c = kk / KC, k = kk % KC;
  const int len = region_len[g];
  float s = 0.f;
  for (int d = 0; d < len; ++d) s += cd[(((size_t)g * nKc + c) * dim + d) * KC + k];
  cst[idx] = s;
}

// Row-wise softmax / log-softmax of raw mixture logits, scattered into the chunked layouts.
//   mode 0 (inner sum level): src (P, O, Kin2); row = (p, o);     dst [p][o/OC][ij][o%OC]
//   mode 1 (root):            src (C, P*Kin2); row = c;           dst [p][c/OC][ij][c%OC]
// =================================================================================================
// Leaf level
// =================================================================================================
struct LeafArgs {
  const float* x;            // (B, D)
  const int32_t* mask;       // (G0, dim)
  const int32_t* region_len; // (G0)
  const float* tab;          // chunked table, see ratspn_plan.cuh
  const float* cd;           // [G0][nKc][dim][KC]
  const float* cst;          // [G0][Kp]
  float* out;                // [G0][K][Bp], or tile-major (out_cs, out_ts: RatPlan::act0_cs / act0_ts)
  int64_t out_cs, out_ts;
  const int* redo;           // NULL, or [Bp/32]: only tiles with a flagged 32-sample group are computed
  int64_t B, Bp;
  int D, G0, K, dim, nKc, regions_per_cta;
  LeafTabGeom g;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}

// One CTA = one tile of TB = 32*ST samples x a contiguous range of regions, 8 warps.
//  * the x tile is staged ONCE, transposed, with asynchronous 4-byte copies:
//      element (feature f, sample s) at f*TB + ((s&31) ^ (f&31)) + (s & ~31)
//    (conflict-free for the fill, lanes = features, and for the sweep, lanes = samples);
//  * a warp owns one region at a time, lanes = samples, accumulators for ST samples x KC channels
//    live in registers; the (1/sigma, -mu/sigma) rows of the region stream through a warp-private
//    ring of TMA bulk copies (cp.async.bulk + mbarrier) and are read as warp-uniform broadcasts;
//    the chunk header holds the pre-swizzled x-tile offset of every row, so the gather address is
//    one shuffle + one XOR;
//  * rows are software-pipelined through two register sets so the shared loads of row d+1 overlap
//    the packed FFMA2 of row d:  t = x*rs + mr ; acc += t*t  (2 FFMA2 per channel pair and feature).
// A non-finite input makes the fast result non-finite; that is detected per region and the region is
// redone on the exact path, which applies nan_to_num term by term like ratspn.py:103.
template <int KIND>
__device__ __forceinline__ float leaf_term(float xv, float p0, float p1, float cd) {
  if constexpr (KIND == DPK_LEAF_GAUSSIAN) {
    const float t = fmaf(xv, p0, p1);
    return fmaf(-0.5f * t, t, cd);
  } else if constexpr (KIND == kLeafGaussUnit) {
    const float t = xv + p0;
    return fmaf(-0.5f * t, t, cd);
  } else {
    return fmaf(xv, p0, cd);
  }
}

template <int KC, int ST, int KIND>
__global__ void __launch_bounds__(256, 1) ratspn_leaf_kernel(const LeafArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int TB = 32 * ST;
  constexpr int NP = (KIND == DPK_LEAF_GAUSSIAN) ? 2 : 1;
  constexpr int NPK = (NP * KC + 3) / 4 * 4;
  constexpr int KH = KC / 2;
  constexpr bool QUAD = (KIND != DPK_LEAF_BERNOULLI);   // quadratic (Gaussian) vs linear (Bernoulli) term
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int CH = a.g.CH, NCH = a.g.NCH, CHP = a.g.CHP, CF = a.g.CF, NST = a.g.NST;

  float* xs = reinterpret_cast<float*>(smem_raw);
  float* ring = xs + (size_t)a.D * TB + (size_t)warp * NST * CF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + (size_t)a.D * TB + (size_t)kLeafWarps * NST * CF) +
                   warp * kLeafMaxStages;

  // ---- this warp's work list: regions r_begin+warp, +8, ... ; nKc*NCH chunks each ---------------
  const int r_begin = blockIdx.y * a.regions_per_cta;
  const int r_end = min(a.G0, r_begin + a.regions_per_cta);
  const int n_reg = max(0, (r_end - r_begin - warp + 7) / 8);
  const int per_region = a.nKc * NCH;

  if (lane == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_launch_dependents();
  pdl_wait();              // flags, tables and act[0] come from the launches before this one
  int p_stage = 0, c_stage = 0;          // ring cursors: a tile consumes exactly what it issued, so they carry over
  uint32_t c_parity = 0;

  // Tiles of this CTA.  The first pass has one tile per CTA; the clean-up pass behind the tensor-core kernel
  // (a.redo != NULL) runs one CTA per SM over all tiles and skips those without a flagged 32-sample group -- nearly all
  // of them: a CTA per tile would cost ~10 us of launches (227 KB of shared memory each) to find nothing to do.
  const int64_t n_tiles = (a.B + TB - 1) / TB;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const int64_t b0 = tile * TB;
  if (a.redo) {
    int any = 0;
#pragma unroll
    for (int s = 0; s < ST; ++s) any |= __ldg(a.redo + (b0 >> 5) + s);
    if (!any) continue;
  }

  // producer side (lane 0 issues; every lane tracks the cursor so the state stays warp-uniform)
  const float* p_src = a.tab + (size_t)(r_begin + warp) * per_region * CF;
  int p_left = n_reg * per_region, p_in_region = 0;
  const uint32_t chunk_bytes = (uint32_t)CF * 4;
  auto issue = [&]() {
    if (p_left > 0) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bars + p_stage, chunk_bytes);
        bulk_g2s(ring + (size_t)p_stage * CF, p_src, chunk_bytes, bars + p_stage);
      }
      --p_left;
      p_src += CF;
      if (++p_in_region == per_region) { p_in_region = 0; p_src += (size_t)(kLeafWarps - 1) * per_region * CF; }
      p_stage = (p_stage + 1 == NST) ? 0 : p_stage + 1;
    }
  };
  for (int s = 0; s < NST - 1; ++s) issue();

  // ---- x tile: asynchronous transposing fill -------------------------------------------------
  for (int s = warp; s < TB; s += 8) {
    const int64_t b = b0 + s;
    const bool inb = b < a.B;
    const float* row = a.x + (inb ? b : 0) * a.D;
    const int sw = s & 31, hi = s & ~31;
    for (int f = lane; f < a.D; f += 32) cp_async_4(xs + f * TB + ((sw ^ (f & 31)) | hi), row + f, inb ? 4u : 0u);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- sweep ------------------------------------------------------------------------------------
  const char* xs_bytes = reinterpret_cast<const char*>(xs);
  const uint32_t lane_x = (uint32_t)lane << 2;
  for (int ri = 0; ri < n_reg; ++ri) {
    const int r = r_begin + warp + 8 * ri;
    const int len_r = __ldg(a.region_len + r);   // pad slots (d >= len_r) are not swept: they contribute exactly 0
    for (int c = 0; c < a.nKc; ++c) {
      float2 acc[ST][KH];
#pragma unroll
      for (int s = 0; s < ST; ++s)
#pragma unroll
        for (int k = 0; k < KH; ++k) acc[s][k] = make_float2(0.f, 0.f);

      for (int ch = 0; ch < NCH; ++ch) {
        issue();                                         // keep NST-1 chunks in flight
        mbar_wait(bars + c_stage, c_parity);
        const float* __restrict__ chunk = ring + (size_t)c_stage * CF;
        const uint32_t* __restrict__ hdr = reinterpret_cast<const uint32_t*>(chunk);
        const float* __restrict__ rows = chunk + CHP;
        const int nrows = max(0, min(CH, len_r - ch * CH));

        // Software pipeline, per row d: header word fetched 2 rows ahead, x + parameters 1 row ahead,
        // so that no shared-memory latency sits between the FFMA2 blocks of consecutive rows.
        auto load_row_regs = [&](int d, uint32_t h, float (&p)[NPK], float (&xv)[ST]) {
          const float* xp = reinterpret_cast<const float*>(xs_bytes + (h ^ lane_x));
#pragma unroll
          for (int s = 0; s < ST; ++s) xv[s] = xp[32 * s];
          load_row_smem<NPK>(rows + d * NPK, p);
        };
        auto fma_row = [&](const float (&p)[NPK], const float (&xv)[ST]) {
#pragma unroll
          for (int s = 0; s < ST; ++s) {
            const float2 x2 = make_float2(xv[s], xv[s]);
#pragma unroll
            for (int k = 0; k < KH; ++k) {
              const float2 p0 = make_float2(p[2 * k], p[2 * k + 1]);
              if constexpr (KIND == DPK_LEAF_GAUSSIAN) {
                const float2 t = __ffma2_rn(x2, p0, make_float2(p[KC + 2 * k], p[KC + 2 * k + 1]));
                acc[s][k] = __ffma2_rn(t, t, acc[s][k]);
              } else if constexpr (KIND == kLeafGaussUnit) {
                const float2 t = __fadd2_rn(x2, p0);
                acc[s][k] = __ffma2_rn(t, t, acc[s][k]);
              } else {
                acc[s][k] = __ffma2_rn(x2, p0, acc[s][k]);
              }
            }
          }
        };
        float pA[NPK], pB[NPK], xA[ST], xB[ST];
        uint32_t h1 = hdr[nrows > 1 ? 1 : 0];
        load_row_regs(0, hdr[0], pA, xA);
        int d = 0;
        for (; d + 2 <= nrows; d += 2) {
          const uint32_t h2 = hdr[d + 2 < nrows ? d + 2 : d];
          load_row_regs(d + 1, h1, pB, xB);
          fma_row(pA, xA);
          h1 = hdr[d + 3 < nrows ? d + 3 : d];
          if (d + 2 < nrows) load_row_regs(d + 2, h2, pA, xA);
          fma_row(pB, xB);
        }
        if (d < nrows) fma_row(pA, xA);
        __syncwarp();                                    // every lane is done with this stage
        if (++c_stage == NST) { c_stage = 0; c_parity ^= 1u; }
      }

      // ---- finish the (region, channel chunk): constants, non-finite check, store -------------
      const float* __restrict__ cst = a.cst + (size_t)r * (KC * a.nKc) + c * KC;
      float val[ST][KC];
      bool bad = false;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const float cv = __ldg(cst + k);
#pragma unroll
        for (int s = 0; s < ST; ++s) {
          const float av = (k & 1) ? acc[s][k / 2].y : acc[s][k / 2].x;
          val[s][k] = QUAD ? fmaf(-0.5f, av, cv) : av + cv;
          bad |= !(fabsf(val[s][k]) <= FLT_MAX);
        }
      }
      if (__any_sync(0xffffffffu, bad)) {
        // exact path for this region: every term goes through nan_to_num like ratspn.py:103
        const float* __restrict__ block = a.tab + ((size_t)r * a.nKc + c) * NCH * CF;
        const float* __restrict__ cdt = a.cd + ((size_t)r * a.nKc + c) * a.dim * KC;
        const int32_t* __restrict__ m = a.mask + (size_t)r * a.dim;
        const int len = __ldg(a.region_len + r);
#pragma unroll
        for (int s = 0; s < ST; ++s)
#pragma unroll
          for (int k = 0; k < KC; ++k) val[s][k] = 0.f;
        for (int d = 0; d < len; ++d) {
          const int f = __ldg(m + d);
          float p[NPK], qd[KC];
          load_row<NPK>(leaf_tab_row(block, d, a.g), p);
          load_row<KC>(cdt + (size_t)d * KC, qd);
#pragma unroll
          for (int s = 0; s < ST; ++s) {
            const float xv = xs[f * TB + (lane ^ (f & 31)) + 32 * s];
#pragma unroll
            for (int k = 0; k < KC; ++k)
              val[s][k] += nan_to_num(leaf_term<KIND>(xv, p[k], NP == 2 ? p[KC + k] : 0.f, qd[k]));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = c * KC + k;
        if (kk < a.K) {
#pragma unroll
          for (int s = 0; s < ST; ++s) {
            const int64_t b = b0 + lane + 32 * s;
            a.out[(b >> 7) * a.out_ts + ((size_t)r * a.K + kk) * a.out_cs + (b & 127)] = val[s][k];
          }
        }
      }
    }
  }
  __syncthreads();   // the x tile is refilled by the next iteration
  }
}

// Fallback for inputs too wide for a shared-memory tile: x is gathered straight from global/L2 and
// every term takes the exact path.  Correct for any D, not tuned.
template <int KC, int KIND>
__global__ void __launch_bounds__(256) ratspn_leaf_wide_kernel(const LeafArgs a) {
  constexpr int NP = (KIND == DPK_LEAF_GAUSSIAN) ? 2 : 1;
  constexpr int NPK = (NP * KC + 3) / 4 * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * 32 + lane;
  if (a.redo && !__ldg(a.redo + blockIdx.x)) return;
  const int r_begin = blockIdx.y * a.regions_per_cta;
  const int r_end = min(a.G0, r_begin + a.regions_per_cta);
  for (int r = r_begin + warp; r < r_end; r += 8) {
    const int len = __ldg(a.region_len + r);
    const int32_t* __restrict__ m = a.mask + (size_t)r * a.dim;
    for (int c = 0; c < a.nKc; ++c) {
      const float* __restrict__ block = a.tab + ((size_t)r * a.nKc + c) * a.g.NCH * a.g.CF;
      const float* __restrict__ cdt = a.cd + ((size_t)r * a.nKc + c) * a.dim * KC;
      float val[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) val[k] = 0.f;
      for (int d = 0; d < len; ++d) {
        const int f = __ldg(m + d);
        const float xv = (b < a.B) ? __ldg(a.x + b * a.D + f) : 0.f;
        float p[NPK], qd[KC];
        load_row<NPK>(leaf_tab_row(block, d, a.g), p);
        load_row<KC>(cdt + (size_t)d * KC, qd);
#pragma unroll
        for (int k = 0; k < KC; ++k)
          val[k] += nan_to_num(leaf_term<KIND>(xv, p[k], NP == 2 ? p[KC + k] : 0.f, qd[k]));
      }
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = c * KC + k;
        if (kk < a.K && b < a.Bp) a.out[(b >> 7) * a.out_ts + ((size_t)r * a.K + kk) * a.out_cs + (b & 127)] = val[k];
      }
    }
  }
}

__global__ void transpose_to_batch_major(const float* __restrict__ in, float* __restrict__ out, int rows,
                                         int64_t B, int64_t Bp) {
  __shared__ float tile[32][33];
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i;
    const int64_t b = b0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && b < Bp) ? in[(size_t)r * Bp + b] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t b = b0 + i;
    const int r = r0 + threadIdx.x;
    if (b < B && r < rows) out[(size_t)b * rows + r] = tile[threadIdx.x][i];
  }
}


// =================================================================================================
// Host-side launchers
// =================================================================================================
struct LeafLaunch {
  int mode;      // 2: 64-sample tile, 1: 32-sample tile, 0: wide fallback
  dim3 grid;
  size_t smem;
  bool pdl;      // redo pass of a narrow model: staged behind the leaf launch (common.cuh)
};

template <int KC, int KIND>
static int launch_leaf_k(const LeafArgs& a, const LeafLaunch& L, cudaStream_t st) {
  ProfScope prof(CAT_LEAF, st);
  if (L.mode == 4) {
    auto kern = ratspn_leaf_kernel<KC, 4, KIND>;
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    if (a.redo) DPK_CUDA_TRY(launch_pdl(L.pdl, kern, L.grid, dim3(256), L.smem, st, a));
    else kern<<<L.grid, 256, L.smem, st>>>(a);
  } else if (L.mode == 2) {
    auto kern = ratspn_leaf_kernel<KC, 2, KIND>;
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    if (a.redo) DPK_CUDA_TRY(launch_pdl(L.pdl, kern, L.grid, dim3(256), L.smem, st, a));
    else kern<<<L.grid, 256, L.smem, st>>>(a);
  } else if (L.mode == 1) {
    auto kern = ratspn_leaf_kernel<KC, 1, KIND>;
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    if (a.redo) DPK_CUDA_TRY(launch_pdl(L.pdl, kern, L.grid, dim3(256), L.smem, st, a));
    else kern<<<L.grid, 256, L.smem, st>>>(a);
  } else {
    ratspn_leaf_wide_kernel<KC, KIND><<<L.grid, 256, 0, st>>>(a);
  }
  DPK_LAUNCH_CHECK("ratspn_leaf_kernel");
  return DPK_OK;
}

template <int KIND>
static int launch_leaf_kind(int KC, const LeafArgs& a, const LeafLaunch& L, cudaStream_t st) {
  switch (KC) {
    case 2: return launch_leaf_k<2, KIND>(a, L, st);
    case 4: return launch_leaf_k<4, KIND>(a, L, st);
    case 8: return launch_leaf_k<8, KIND>(a, L, st);
    case 10: return launch_leaf_k<10, KIND>(a, L, st);
    case 16: return launch_leaf_k<16, KIND>(a, L, st);
  }
  return set_error(DPK_E_ARG, "unsupported leaf channel chunk %d", KC);
}

static LeafTabGeom leaf_geom(const RatPlan& p) {
  LeafTabGeom g;
  g.CH = p.leaf_ch; g.NCH = p.leaf_nch; g.CHP = p.leaf_chp; g.NPK = p.leaf_npk; g.CF = p.leaf_chunk_floats; g.TB = p.leaf_tb; g.NST = p.leaf_stages;
  return g;
}

int ratspn_run_prep_leaf(const dpk_ratspn_desc* d, const RatPlan& p, float* ws, cudaStream_t st) {
  const int64_t total = (int64_t)p.G0 * p.kc.padded * p.leaf_nch * p.leaf_ch;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 4096);
  if (p.fwd_kind == DPK_LEAF_GAUSSIAN)
    ratspn_prep_leaf_kernel<DPK_LEAF_GAUSSIAN><<<blocks, 256, 0, st>>>(d->leaf_p0, d->leaf_p1, d->mask, d->region_len,
                                                                        p.G0, p.K, p.dim, p.kc.chunk, p.kc.count,
                                                                        leaf_geom(p), ws + p.off_tab, ws + p.off_cd);
  else if (p.fwd_kind == kLeafGaussUnit)
    ratspn_prep_leaf_kernel<kLeafGaussUnit><<<blocks, 256, 0, st>>>(d->leaf_p0, nullptr, d->mask, d->region_len,
                                                                     p.G0, p.K, p.dim, p.kc.chunk, p.kc.count,
                                                                     leaf_geom(p), ws + p.off_tab, ws + p.off_cd);
  else
    ratspn_prep_leaf_kernel<DPK_LEAF_BERNOULLI><<<blocks, 256, 0, st>>>(d->leaf_p0, nullptr, d->mask, d->region_len,
                                                                         p.G0, p.K, p.dim, p.kc.chunk, p.kc.count,
                                                                         leaf_geom(p), ws + p.off_tab, ws + p.off_cd);
  DPK_LAUNCH_CHECK("ratspn_prep_leaf_kernel");
  const int n = p.G0 * p.kc.padded;
  ratspn_prep_const_kernel<<<(n + 127) / 128, 128, 0, st>>>(ws + p.off_cd, d->region_len, p.G0, p.dim, p.kc.chunk,
                                                             p.kc.count, ws + p.off_cst);
  DPK_LAUNCH_CHECK("ratspn_prep_const_kernel");
  if (p.leaf_mma) return ratspn_run_prep_leaf_mma(d, p, ws, st);
  return DPK_OK;
}

int ratspn_run_leaf(const dpk_ratspn_desc* d, const RatPlan& p, const float* x, float* ws, cudaStream_t st) {
  LeafArgs a;
  a.x = x; a.mask = d->mask; a.region_len = d->region_len;
  a.tab = ws + p.off_tab; a.cd = ws + p.off_cd; a.cst = ws + p.off_cst; a.out = ws + p.off_act[0];
  a.redo = nullptr;
  if (p.leaf_mma && ((uintptr_t)x & 15) == 0) {
    // tensor-core pass first; the exact kernel below then redoes only the flagged sample groups
    int rc = p.leaf_stream ? ratspn_run_leaf_stream(p, x, ws, st) : ratspn_run_leaf_mma(p, x, ws, st);
    if (rc) return rc;
    a.redo = reinterpret_cast<const int*>(ws + p.off_mflags);
  } else if (p.leaf_mma && p.off_sqsum) {
    // unaligned x: the exact kernel does everything, quadratic term included -- every group flagged, so that the root
    // does not add the (stale) per-sample term again
    DPK_CUDA_TRY(cudaMemsetAsync(ws + p.off_mflags, 0x01, (size_t)p.Bp / 32 * 4, st));
  }
  a.B = p.B; a.Bp = p.Bp; a.D = p.D; a.G0 = p.G0; a.K = p.K; a.dim = p.dim; a.nKc = p.kc.count;
  a.out_cs = p.act0_cs; a.out_ts = p.act0_ts;
  a.g = leaf_geom(p);
  const int nsm = sm_count();
  LeafLaunch L;
  L.mode = p.leaf_mode;
  const int64_t ntiles = ceil_div(p.B, p.leaf_tb);
  // split the regions over blockIdx.y only when the batch alone cannot fill the SMs
  int rsplit = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div(2 * nsm, ntiles)), ceil_div(p.G0, 8));
  a.regions_per_cta = (int)round_up(ceil_div(p.G0, rsplit), 8);
  rsplit = (int)ceil_div(p.G0, a.regions_per_cta);
  L.grid = dim3((unsigned)((a.redo && L.mode != 0) ? std::min<int64_t>(ntiles, nsm) : ntiles), (unsigned)rsplit);
  L.smem = p.leaf_smem;
  L.pdl = p.leaf_stream != 0;
  if (p.fwd_kind == DPK_LEAF_GAUSSIAN) return launch_leaf_kind<DPK_LEAF_GAUSSIAN>(p.kc.chunk, a, L, st);
  if (p.fwd_kind == kLeafGaussUnit) return launch_leaf_kind<kLeafGaussUnit>(p.kc.chunk, a, L, st);
  return launch_leaf_kind<DPK_LEAF_BERNOULLI>(p.kc.chunk, a, L, st);
}

}  // namespace dpk

using namespace dpk;

extern "C" int dpk_ratspn_leaf_forward(const dpk_ratspn_desc* desc, const float* x, int64_t batch, float* out,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  RatPlan p;
  int rc = make_plan(desc, batch, kPlanLeafOnly, &p);
  if (rc) return rc;
  if (batch == 0) return DPK_OK;
  if (!x || !out || !desc->mask || !desc->region_len || !desc->leaf_p0)
    return set_error(DPK_E_ARG, "null pointer argument");
  if ((rc = ratspn_check_ws(p, workspace, workspace_bytes))) return rc;
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfScope prof(CAT_PREP, st, 2);
    if ((rc = ratspn_run_prep_leaf(desc, p, ws, st))) return rc;
  }
  if ((rc = ratspn_run_leaf(desc, p, x, ws, st))) return rc;
  const int rows = p.G0 * p.K;
  dim3 grid((unsigned)ceil_div(p.B, 32), (unsigned)ceil_div(rows, 32));
  ProfScope prof(CAT_LAYER, st);
  transpose_to_batch_major<<<grid, dim3(32, 8), 0, st>>>(ws + p.off_act[0], out, rows, p.B, p.Bp);
  DPK_LAUNCH_CHECK("transpose_to_batch_major");
  return DPK_OK;
}
