// ratspn_einsum.cu -- RAT-SPN product + sum levels and the root:
//   ProductLayer.forward  deeprob/spn/layers/ratspn.py:272-286  outer sum of sibling regions
//   SumLayer.forward      deeprob/spn/layers/ratspn.py:363-378  logsumexp(x + log_softmax(W))
//   RootLayer.forward     deeprob/spn/layers/ratspn.py:446-458
// fused in the "einsum" form
//   y[b,p,o] = ml + mr + log sum_ij softmax(W)[p,o,ij] * exp(l_i - ml) * exp(r_j - mr)
// (2K exps instead of O*K^2, the K^2 product never materialised), with an exact log-domain fallback
// when the linear-domain sum underflows so that extreme weights still match torch.logsumexp.
// The root is the same contraction with the root weights (normalised over all partitions), followed
// by a logsumexp over the partitions.
#include <algorithm>

#include "ratspn_kernels.cuh"

namespace dpk {

// One CTA per (padded) row; padded rows (o >= O) are written as weight 0 / log-weight -inf.
__global__ void ratspn_prep_weight_kernel(const float* __restrict__ src, int mode, int P, int O, int Kin2,
                                          int OC, int nOc, float* __restrict__ wsoft, float* __restrict__ wlog) {
  __shared__ float red[32];
  const int Op = OC * nOc;
  int p_row, o;
  int64_t len;
  const float* row;
  if (mode == 0) { p_row = blockIdx.x / Op; o = blockIdx.x % Op; len = Kin2; row = src + ((size_t)p_row * O + o) * Kin2; }
  else           { p_row = 0;               o = blockIdx.x;      len = (int64_t)P * Kin2; row = src + (size_t)o * len; }
  const bool live = o < O;
  float lse = 0.f;
  if (live) {
    float m = -INFINITY;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) m = fmaxf(m, row[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
    m = warp_max(m);
    m = __shfl_sync(0xffffffffu, m, 0);
    __syncthreads();
    if (threadIdx.x == 0) red[0] = m;
    __syncthreads();
    m = red[0];
    __syncthreads();
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) s += expf(row[i] - m);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    __syncthreads();
    if (threadIdx.x == 0) red[0] = m + logf(s);
    __syncthreads();
    lse = red[0];
  }
  const int oc = o / OC, ok = o % OC;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
    const int p = (mode == 0) ? p_row : (int)(i / Kin2);
    const int ij = (mode == 0) ? (int)i : (int)(i % Kin2);
    const size_t dst = (((size_t)p * nOc + oc) * Kin2 + ij) * OC + ok;
    const float lw = live ? row[i] - lse : -INFINITY;
    wlog[dst] = lw;
    wsoft[dst] = live ? expf(lw) : 0.f;
  }
}

// =================================================================================================
// Product + Sum ("einsum") level and root
// =================================================================================================
struct EinsumArgs {
  const float* in;     // [2P][Kin][Bp]
  const float* wsoft;  // [P][nOc][Kin2][OC]
  const float* wlog;
  float* out;          // inner: [P][O][Bp]   root: (B, O) row-major
  int64_t B, Bp;
  int P, Kin, O, nOc, rows_per_chunk;  // rows_per_chunk: i-rows of the weight staged in smem at a time
};

constexpr int kEinsumThreads = 128;
constexpr float kTinySum = 1e-18f;

// exact log-domain value of one output (fallback): logsumexp_ij(l_i + r_j + logw[ij])
__device__ __noinline__ float einsum_exact(const float* __restrict__ l, const float* __restrict__ r,
                                           int64_t stride, int Kin, const float* __restrict__ wlog, int OC) {
  float m = -INFINITY;
  for (int i = 0; i < Kin; ++i)
    for (int j = 0; j < Kin; ++j)
      m = fmaxf(m, l[i * stride] + r[j * stride] + wlog[(size_t)(i * Kin + j) * OC]);
  if (!(fabsf(m) <= FLT_MAX)) return m;  // -inf (all dropped), +inf or NaN propagate like torch.logsumexp
  float s = 0.f;
  for (int i = 0; i < Kin; ++i)
    for (int j = 0; j < Kin; ++j)
      s += expf(l[i * stride] + r[j * stride] + wlog[(size_t)(i * Kin + j) * OC] - m);
  return m + logf(s);
}

// thread = ST samples (b = base + tid + s*128) of one partition; el/er (shifted exps) live in smem
// [k][sample] (conflict-free), weights of the partition are staged in smem and read as broadcasts.
template <int OC, int ST, bool ROOT>
__global__ void __launch_bounds__(kEinsumThreads) ratspn_einsum_kernel(const EinsumArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NS = kEinsumThreads * ST;
  float* el = sm;
  float* er = el + (size_t)a.Kin * NS;
  float* wsm = er + (size_t)a.Kin * NS;
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * NS;
  const int Kin = a.Kin, Kin2 = Kin * Kin;

  const int p_first = ROOT ? 0 : blockIdx.y;
  const int p_last = ROOT ? a.P : blockIdx.y + 1;

  for (int oc = 0; oc < a.nOc; ++oc) {
    float run_m[ST], run_acc[ST][OC];  // ROOT: online logsumexp over partitions
    if constexpr (ROOT) {
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        run_m[s] = -INFINITY;
#pragma unroll
        for (int o = 0; o < OC; ++o) run_acc[s][o] = 0.f;
      }
    }
    for (int p = p_first; p < p_last; ++p) {
      const float* __restrict__ lin = a.in + (size_t)(2 * p) * Kin * a.Bp;
      const float* __restrict__ rin = lin + (size_t)Kin * a.Bp;
      float ml[ST], mr[ST];
      __syncthreads();  // previous users of el/er/wsm are done
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        const int64_t b = base + tid + s * kEinsumThreads;
        const bool inb = b < a.Bp;
        float vl = -INFINITY, vr = -INFINITY;
        for (int k = 0; k < Kin; ++k) {
          const float l = inb ? lin[(size_t)k * a.Bp + b] : 0.f;
          const float r = inb ? rin[(size_t)k * a.Bp + b] : 0.f;
          el[k * NS + tid + s * kEinsumThreads] = l;
          er[k * NS + tid + s * kEinsumThreads] = r;
          vl = fmaxf(vl, l); vr = fmaxf(vr, r);
        }
        // a fully -inf (dropped-out) or non-finite side: shift by 0, the exact fallback sorts it out
        ml[s] = (fabsf(vl) <= FLT_MAX) ? vl : 0.f;
        mr[s] = (fabsf(vr) <= FLT_MAX) ? vr : 0.f;
        for (int k = 0; k < Kin; ++k) {
          const int o = k * NS + tid + s * kEinsumThreads;
          el[o] = __expf(el[o] - ml[s]);
          er[o] = __expf(er[o] - mr[s]);
        }
      }
      float acc[ST][OC];
#pragma unroll
      for (int s = 0; s < ST; ++s)
#pragma unroll
        for (int o = 0; o < OC; ++o) acc[s][o] = 0.f;

      const float* __restrict__ wp = a.wsoft + ((size_t)p * a.nOc + oc) * Kin2 * OC;
      for (int i0 = 0; i0 < Kin; i0 += a.rows_per_chunk) {
        const int i1 = min(Kin, i0 + a.rows_per_chunk);
        __syncthreads();
        for (int t = tid; t < (i1 - i0) * Kin * OC; t += kEinsumThreads) wsm[t] = __ldg(wp + (size_t)i0 * Kin * OC + t);
        __syncthreads();
        for (int i = i0; i < i1; ++i) {
          float eli[ST];
#pragma unroll
          for (int s = 0; s < ST; ++s) eli[s] = el[i * NS + tid + s * kEinsumThreads];
          const float* wrow = wsm + (size_t)(i - i0) * Kin * OC;
#pragma unroll 2
          for (int j = 0; j < Kin; ++j) {
            float w[OC];
            load_row_smem<OC>(wrow + j * OC, w);
#pragma unroll
            for (int s = 0; s < ST; ++s) {
              const float pij = eli[s] * er[j * NS + tid + s * kEinsumThreads];
#pragma unroll
              for (int o = 0; o < OC; ++o) acc[s][o] = fmaf(w[o], pij, acc[s][o]);
            }
          }
        }
      }

      if constexpr (!ROOT) {
#pragma unroll
        for (int s = 0; s < ST; ++s) {
          const int64_t b = base + tid + s * kEinsumThreads;
          if (b >= a.Bp) continue;
#pragma unroll
          for (int o = 0; o < OC; ++o) {
            const int oo = oc * OC + o;
            if (oo >= a.O) continue;
            float y;
            if (acc[s][o] >= kTinySum && acc[s][o] <= FLT_MAX)
              y = ml[s] + mr[s] + __logf(acc[s][o]);
            else
              y = einsum_exact(lin + b, rin + b, a.Bp, Kin, a.wlog + ((size_t)p * a.nOc + oc) * Kin2 * OC + o, OC);
            a.out[((size_t)p * a.O + oo) * a.Bp + b] = y;
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < ST; ++s) {
          const float mp = ml[s] + mr[s];
          const float nm = fmaxf(run_m[s], mp);
          const float so = __expf(run_m[s] - nm), sn = __expf(mp - nm);  // run_m=-inf -> so = 0
#pragma unroll
          for (int o = 0; o < OC; ++o) run_acc[s][o] = run_acc[s][o] * so + acc[s][o] * sn;
          run_m[s] = nm;
        }
      }
    }
    if constexpr (ROOT) {
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        const int64_t b = base + tid + s * kEinsumThreads;
        if (b >= a.B) continue;
#pragma unroll
        for (int o = 0; o < OC; ++o) {
          const int oo = oc * OC + o;
          if (oo >= a.O) continue;
          float y;
          if (run_acc[s][o] >= kTinySum && run_acc[s][o] <= FLT_MAX) {
            y = run_m[s] + __logf(run_acc[s][o]);
          } else {
            // exact: logsumexp over every partition
            float m = -INFINITY, ssum = 0.f;
            for (int p = 0; p < a.P; ++p) {
              const float* lin = a.in + (size_t)(2 * p) * Kin * a.Bp + b;
              const float v = einsum_exact(lin, lin + (size_t)Kin * a.Bp, a.Bp, Kin,
                                           a.wlog + ((size_t)p * a.nOc + oc) * Kin2 * OC + o, OC);
              if (v == -INFINITY) continue;
              const float nm = fmaxf(m, v);
              ssum = ssum * expf(m - nm) + expf(v - nm);
              m = nm;
            }
            y = (m == -INFINITY) ? -INFINITY : m + logf(ssum);
          }
          a.out[(size_t)b * a.O + oo] = y;
        }
      }
    }
  }
}

// Register-resident variant for the common sizes (Kin, OC in {2,4,8,10,16}): the right-hand exps of
// the thread's 4 samples live in registers (j loop fully unrolled), the partition's whole weight block
// sits in shared memory, outputs are accumulated pairwise with packed FFMA2.  Per (i,j) a thread issues
// 4 FMUL + 4*OC/2 FFMA2 against OC/4 broadcast LDS.128 -> FMA-pipe bound instead of LDS bound.
template <int OC, int KIN, bool ROOT, int ST = 4>
__global__ void __launch_bounds__(kEinsumThreads) ratspn_einsum_reg_kernel(const EinsumArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NS = kEinsumThreads * ST, K2 = KIN * KIN, OH = OC / 2;
  float* wsm = sm;              // [K2][OC]
  float* el = wsm + K2 * OC;    // [KIN][NS]
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * NS;
  const int p_first = ROOT ? 0 : blockIdx.y;
  const int p_last = ROOT ? a.P : blockIdx.y + 1;

  for (int oc = 0; oc < a.nOc; ++oc) {
    float run_m[ST];
    float2 run_acc[ST][OH];
    if constexpr (ROOT) {
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        run_m[s] = -INFINITY;
#pragma unroll
        for (int o = 0; o < OH; ++o) run_acc[s][o] = make_float2(0.f, 0.f);
      }
    }
    for (int p = p_first; p < p_last; ++p) {
      const float* __restrict__ lin = a.in + (size_t)(2 * p) * KIN * a.Bp;
      const float* __restrict__ rin = lin + (size_t)KIN * a.Bp;
      const float* __restrict__ wp = a.wsoft + ((size_t)p * a.nOc + oc) * K2 * OC;
      __syncthreads();  // previous users of wsm are done
      for (int t = tid; t < K2 * OC / 2; t += kEinsumThreads)
        reinterpret_cast<float2*>(wsm)[t] = __ldg(reinterpret_cast<const float2*>(wp) + t);
      float er[ST][KIN], ml[ST], mr[ST];
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        const int64_t b = base + tid + s * kEinsumThreads;
        const bool inb = b < a.Bp;
        float l[KIN];
        float vl = -INFINITY, vr = -INFINITY;
#pragma unroll
        for (int k = 0; k < KIN; ++k) {
          l[k] = inb ? lin[(size_t)k * a.Bp + b] : 0.f;
          er[s][k] = inb ? rin[(size_t)k * a.Bp + b] : 0.f;
          vl = fmaxf(vl, l[k]); vr = fmaxf(vr, er[s][k]);
        }
        ml[s] = (fabsf(vl) <= FLT_MAX) ? vl : 0.f;
        mr[s] = (fabsf(vr) <= FLT_MAX) ? vr : 0.f;
#pragma unroll
        for (int k = 0; k < KIN; ++k) {
          el[k * NS + tid + s * kEinsumThreads] = __expf(l[k] - ml[s]);
          er[s][k] = __expf(er[s][k] - mr[s]);
        }
      }
      __syncthreads();  // weights staged
      float2 acc[ST][OH];
#pragma unroll
      for (int s = 0; s < ST; ++s)
#pragma unroll
        for (int o = 0; o < OH; ++o) acc[s][o] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int i = 0; i < KIN; ++i) {
        float eli[ST];
#pragma unroll
        for (int s = 0; s < ST; ++s) eli[s] = el[i * NS + tid + s * kEinsumThreads];
        const float* wrow = wsm + i * KIN * OC;
#pragma unroll
        for (int j = 0; j < KIN; ++j) {
          float w[OC];
          load_row_smem<OC>(wrow + j * OC, w);
#pragma unroll
          for (int s = 0; s < ST; ++s) {
            const float pij = eli[s] * er[s][j];
            const float2 p2 = make_float2(pij, pij);
#pragma unroll
            for (int o = 0; o < OH; ++o) acc[s][o] = __ffma2_rn(make_float2(w[2 * o], w[2 * o + 1]), p2, acc[s][o]);
          }
        }
      }
      if constexpr (!ROOT) {
#pragma unroll
        for (int s = 0; s < ST; ++s) {
          const int64_t b = base + tid + s * kEinsumThreads;
          if (b >= a.Bp) continue;
#pragma unroll
          for (int o = 0; o < OC; ++o) {
            const int oo = oc * OC + o;
            if (oo >= a.O) continue;
            const float av = (o & 1) ? acc[s][o / 2].y : acc[s][o / 2].x;
            float y;
            if (av >= kTinySum && av <= FLT_MAX) y = ml[s] + mr[s] + __logf(av);
            else y = einsum_exact(lin + b, rin + b, a.Bp, KIN, a.wlog + ((size_t)p * a.nOc + oc) * K2 * OC + o, OC);
            a.out[((size_t)p * a.O + oo) * a.Bp + b] = y;
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < ST; ++s) {
          const float mp = ml[s] + mr[s];
          const float nm = fmaxf(run_m[s], mp);
          const float so = __expf(run_m[s] - nm), sn = __expf(mp - nm);
#pragma unroll
          for (int o = 0; o < OH; ++o) {
            run_acc[s][o].x = run_acc[s][o].x * so + acc[s][o].x * sn;
            run_acc[s][o].y = run_acc[s][o].y * so + acc[s][o].y * sn;
          }
          run_m[s] = nm;
        }
      }
    }
    if constexpr (ROOT) {
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        const int64_t b = base + tid + s * kEinsumThreads;
        if (b >= a.B) continue;
#pragma unroll
        for (int o = 0; o < OC; ++o) {
          const int oo = oc * OC + o;
          if (oo >= a.O) continue;
          const float av = (o & 1) ? run_acc[s][o / 2].y : run_acc[s][o / 2].x;
          float y;
          if (av >= kTinySum && av <= FLT_MAX) {
            y = run_m[s] + __logf(av);
          } else {
            float m = -INFINITY, ssum = 0.f;
            for (int p = 0; p < a.P; ++p) {
              const float* lin = a.in + (size_t)(2 * p) * KIN * a.Bp + b;
              const float v = einsum_exact(lin, lin + (size_t)KIN * a.Bp, a.Bp, KIN,
                                           a.wlog + ((size_t)p * a.nOc + oc) * K2 * OC + o, OC);
              if (v == -INFINITY) continue;
              const float nm = fmaxf(m, v);
              ssum = ssum * expf(m - nm) + expf(v - nm);
              m = nm;
            }
            y = (m == -INFINITY) ? -INFINITY : m + logf(ssum);
          }
          a.out[(size_t)b * a.O + oo] = y;
        }
      }
    }
  }
}

// [rows][Bp] (sample-minor) -> (B, rows) row-major, for the stand-alone layer entry points
template <int OC, int ST>
static int launch_einsum_t(const EinsumArgs& a, dim3 grid, size_t smem, int cat, cudaStream_t st) {
  auto kern = ratspn_einsum_kernel<OC, ST, false>;
  if (smem > 48 * 1024)
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(cat, st);
  kern<<<grid, kEinsumThreads, smem, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_einsum_kernel");
  return DPK_OK;
}

template <int OC, int KIN>
static int launch_einsum_reg_t(const EinsumArgs& a, int cat, cudaStream_t st) {
  // samples per thread: 2 measured fastest on B200 (fewer registers -> more resident CTAs beats the better
  // amortisation of the weight broadcasts at 4); DPK_EINSUM_ST overrides
  const int knob = env_int("DPK_EINSUM_ST", 2);
  const int ST = (knob == 1 || knob == 4) ? knob : 2;
  auto kern = (ST == 1)   ? ratspn_einsum_reg_kernel<OC, KIN, false, 1>
              : (ST == 2) ? ratspn_einsum_reg_kernel<OC, KIN, false, 2>
                          : ratspn_einsum_reg_kernel<OC, KIN, false, 4>;
  const size_t smem = ((size_t)KIN * KIN * OC + (size_t)KIN * kEinsumThreads * ST) * 4;
  if (smem > 48 * 1024)
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(a.Bp, kEinsumThreads * ST), (unsigned)a.P);
  ProfScope prof(cat, st);
  kern<<<grid, kEinsumThreads, smem, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_einsum_reg_kernel");
  return DPK_OK;
}

template <int OC>
static int launch_einsum_reg_k(const EinsumArgs& a, int cat, cudaStream_t st) {
  switch (a.Kin) {
    case 2: return launch_einsum_reg_t<OC, 2>(a, cat, st);
    case 4: return launch_einsum_reg_t<OC, 4>(a, cat, st);
    case 8: return launch_einsum_reg_t<OC, 8>(a, cat, st);
    case 10: return launch_einsum_reg_t<OC, 10>(a, cat, st);
    case 16: return launch_einsum_reg_t<OC, 16>(a, cat, st);
  }
  return 1;  // not covered
}

static int launch_einsum(EinsumArgs a, int OC, int cat, cudaStream_t st) {
  const size_t smem_max = (size_t)max_dynamic_smem();
  // 4 samples per thread amortise the weight broadcasts; with too few CTAs that would leave SMs idle
  const int64_t ctas4 = ceil_div(a.Bp, 4 * kEinsumThreads) * a.P;
  const bool big = ctas4 >= 2 * sm_count();
  if (big && env_int("DPK_EINSUM_GENERIC", 0) == 0) {   // register-resident fast path for the common sizes
    int rc = 1;
    switch (OC) {
      case 2: rc = launch_einsum_reg_k<2>(a, cat, st); break;
      case 4: rc = launch_einsum_reg_k<4>(a, cat, st); break;
      case 8: rc = launch_einsum_reg_k<8>(a, cat, st); break;
      case 10: rc = launch_einsum_reg_k<10>(a, cat, st); break;
      case 16: rc = launch_einsum_reg_k<16>(a, cat, st); break;
    }
    if (rc <= 0) return rc;
  }
  const size_t wchunk_budget = 8192;  // bytes of weights staged at a time
  int rows = (int)std::max<size_t>(1, wchunk_budget / ((size_t)a.Kin * OC * 4));
  rows = std::min(rows, a.Kin);
  a.rows_per_chunk = rows;
  const size_t wbytes = (size_t)rows * a.Kin * OC * 4;
  int ST = 4;
  size_t smem = 2 * (size_t)a.Kin * kEinsumThreads * ST * 4 + wbytes;
  if (smem > 100 * 1024 || !big) { ST = 1; smem = 2 * (size_t)a.Kin * kEinsumThreads * 4 + wbytes; }
  if (smem > smem_max) return set_error(DPK_E_ARG, "einsum level with %d inputs per region does not fit shared memory", a.Kin);
  dim3 grid((unsigned)ceil_div(a.Bp, kEinsumThreads * ST), (unsigned)a.P);
#define DPK_EINSUM_CASE(oc)                                                             \
  case oc:                                                                              \
    return (ST == 4) ? launch_einsum_t<oc, 4>(a, grid, smem, cat, st) : launch_einsum_t<oc, 1>(a, grid, smem, cat, st);
  switch (OC) {
    DPK_EINSUM_CASE(2)
    DPK_EINSUM_CASE(4)
    DPK_EINSUM_CASE(8)
    DPK_EINSUM_CASE(10)
    DPK_EINSUM_CASE(16)
  }
#undef DPK_EINSUM_CASE
  return set_error(DPK_E_ARG, "unsupported output chunk %d", OC);
}

// out[b, c] = logsumexp_p part[p][c][b]   (partition partials of the root -> class log-likelihoods)
__global__ void ratspn_root_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int P, int C,
                                           int64_t B, int64_t Bp) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int c = 0; c < C; ++c) {
    float m = -INFINITY;
    for (int p = 0; p < P; ++p) m = fmaxf(m, part[((size_t)p * C + c) * Bp + b]);
    float y = m;
    if (fabsf(m) <= FLT_MAX) {
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += __expf(part[((size_t)p * C + c) * Bp + b] - m);
      y = m + __logf(s);
    }
    out[(size_t)b * C + c] = y;
  }
}

int ratspn_run_prep_weights(const dpk_ratspn_desc* d, const RatPlan& p, float* ws, cudaStream_t st) {
  for (int e = 0; e < p.n_sum; ++e) {
    const int P = p.act_regions[e] / 2, kin2 = p.act_ch[e] * p.act_ch[e];
    ratspn_prep_weight_kernel<<<P * p.oc.padded, 128, 0, st>>>(d->sum_weight[e], 0, P, p.O, kin2, p.oc.chunk, p.oc.count,
                                                                ws + p.off_wsoft[e], ws + p.off_wlog[e]);
    DPK_LAUNCH_CHECK("ratspn_prep_weight_kernel");
    if (p.einsum_mma[e]) {
      int rc = ratspn_run_prep_einsum_mma(ws + p.off_wsoft[e], P, p.O, p.act_ch[e], p.oc.chunk, ws + p.off_wmma[e], st);
      if (rc) return rc;
    }
  }
  const int kin = p.act_ch[p.depth - 1];
  ratspn_prep_weight_kernel<<<p.cc.padded, 256, 0, st>>>(d->root_weight, 1, p.R, p.C, kin * kin, p.cc.chunk, p.cc.count,
                                                          ws + p.off_rsoft, ws + p.off_rlog);
  DPK_LAUNCH_CHECK("ratspn_prep_weight_kernel(root)");
  return DPK_OK;
}

int ratspn_run_upper(const RatPlan& p, float* ws, float* out, cudaStream_t st) {
  for (int e = 0; e < p.n_sum; ++e) {
    EinsumArgs a;
    a.in = ws + p.off_act[e]; a.wsoft = ws + p.off_wsoft[e]; a.wlog = ws + p.off_wlog[e];
    a.out = ws + p.off_act[e + 1];
    a.B = p.B; a.Bp = p.Bp; a.P = p.act_regions[e] / 2; a.Kin = p.act_ch[e]; a.O = p.O; a.nOc = p.oc.count;
    int rc = p.einsum_mma[e]
                 ? ratspn_run_einsum_mma(a.in, ws + p.off_wmma[e], a.wsoft, a.wlog, a.out, p.Bp, a.P, a.Kin, a.O, p.oc.chunk, CAT_EINSUM, st)
                 : launch_einsum(a, p.oc.chunk, CAT_EINSUM, st);
    if (rc) return rc;
  }
  // root = the same contraction with the (globally normalised) root weights, one partial per
  // partition, then a logsumexp over the partitions
  EinsumArgs a;
  const int l = p.depth - 1;
  a.in = ws + p.off_act[l]; a.wsoft = ws + p.off_rsoft; a.wlog = ws + p.off_rlog; a.out = ws + p.off_rtmp;
  a.B = p.B; a.Bp = p.Bp; a.P = p.R; a.Kin = p.act_ch[l]; a.O = p.C; a.nOc = p.cc.count;
  int rc = launch_einsum(a, p.cc.chunk, CAT_ROOT, st);
  if (rc) return rc;
  ProfScope prof(CAT_ROOT, st);
  ratspn_root_combine_kernel<<<(unsigned)ceil_div(p.B, 256), 256, 0, st>>>(ws + p.off_rtmp, out, p.R, p.C, p.B, p.Bp);
  DPK_LAUNCH_CHECK("ratspn_root_combine_kernel");
  return DPK_OK;
}

}  // namespace dpk
