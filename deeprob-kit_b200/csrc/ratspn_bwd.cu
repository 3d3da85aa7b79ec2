// ratspn_bwd.cu -- backward / E-step statistics of the RAT-SPN path.
//
// The reference gets its gradients from autograd over the ATen ops of
//   deeprob/spn/layers/ratspn.py:87-108 (leaf), :272-286 (product), :363-378 (sum), :446-458 (root)
// and defines EM statistics only on the node graph (deeprob/spn/learning/em.py:99-107:
// stats = exp(child_ll - root_ll + log-grad)).  On the tensorised model both are the same top-down
// pass: with q_o = dL/dy_o / S_o (S_o the max-shifted mixture sum of the forward),
//   posterior of product (i,j) under output o :  w[o,ij] * el_i * er_j * q_o
//   dL/dl_i = el_i * sum_j er_j * T_ij,   dL/dr_j = er_j * sum_i el_i * T_ij,   T_ij = sum_o q_o w[o,ij]
//   posterior counts  N[o,ij] = w[o,ij] * sum_b q_bo el_bi er_bj   (= EM "n" statistic; the gradient
//   w.r.t. the raw logits is N - softmax(W) * sum_ij N)
//   leaf moments      S0,S1,S2[g,k,d] = sum_b dL/dLL[b,g,k] * {1, x, x^2}   (= EM leaf statistics; the
//   gradients w.r.t. loc/scale/logits are closed forms of them)
#include <algorithm>

#include "ratspn_kernels.cuh"

namespace dpk {

constexpr int kBwdThreads = 128;

struct EinsumBwdArgs {
  const float* in;     // [2P][Kin][Bp]
  const float* wsoft;  // [P][nOc][Kin2][OC]
  const float* y;      // inner: [P][O][Bp]   root: (B, O)
  const float* gy;     // like y; nullptr = all ones (EM)
  float* gin;          // [2P][Kin][Bp]
  float* wstat;        // [P][nOc][Kin2][OC], accumulated with atomics
  int64_t B, Bp;
  int P, Kin, O, nOc, rows_per_chunk, root;
};

// One CTA = one partition x NS = 128*ST samples.
// Phase A (thread = ST samples): recompute el/er, q; sweep (i,j) once accumulating dL/dl, dL/dr.
// Phase B (lanes = product index ij, warps split the samples): M[ij,o] += q_o * el_i * er_j.
// Shared layout [sample][KS] with KS odd: conflict-free for lanes=samples (phase A) and for
// lanes=ij at a fixed sample (phase B).
template <int OC, int ST>
__global__ void __launch_bounds__(kBwdThreads) ratspn_einsum_bwd_kernel(const EinsumBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NS = kBwdThreads * ST;
  constexpr int QS = (OC + 3) / 4 * 4;
  const int Kin = a.Kin, Kin2 = Kin * Kin, KS = Kin | 1;
  float* q_sm = sm;                          // [NS][QS]   (16-byte aligned rows)
  float* el = q_sm + (size_t)NS * QS;        // [NS][KS]
  float* er = el + (size_t)NS * KS;
  float* gl = er + (size_t)NS * KS;
  float* gr = gl + (size_t)NS * KS;
  float* wsm = gr + (size_t)NS * KS;         // [rows_per_chunk][Kin][OC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.y;
  const int64_t base = (int64_t)blockIdx.x * NS;
  const float* __restrict__ lin = a.in + (size_t)(2 * p) * Kin * a.Bp;
  const float* __restrict__ rin = lin + (size_t)Kin * a.Bp;

  float ml[ST], mr[ST];
#pragma unroll
  for (int s = 0; s < ST; ++s) {
    const int row = tid + s * kBwdThreads;
    const int64_t b = base + row;
    const bool inb = b < a.B;
    float vl = -INFINITY, vr = -INFINITY;
    for (int k = 0; k < Kin; ++k) {
      const float l = inb ? lin[(size_t)k * a.Bp + b] : 0.f;
      const float r = inb ? rin[(size_t)k * a.Bp + b] : 0.f;
      el[row * KS + k] = l; er[row * KS + k] = r;
      vl = fmaxf(vl, l); vr = fmaxf(vr, r);
    }
    ml[s] = (fabsf(vl) <= FLT_MAX) ? vl : 0.f;
    mr[s] = (fabsf(vr) <= FLT_MAX) ? vr : 0.f;
    for (int k = 0; k < Kin; ++k) {
      el[row * KS + k] = __expf(el[row * KS + k] - ml[s]);
      er[row * KS + k] = __expf(er[row * KS + k] - mr[s]);
      gl[row * KS + k] = 0.f; gr[row * KS + k] = 0.f;
    }
  }

  for (int oc = 0; oc < a.nOc; ++oc) {
    float q[ST][OC];
#pragma unroll
    for (int s = 0; s < ST; ++s) {
      const int row = tid + s * kBwdThreads;
      const int64_t b = base + row;
      float gq[OC];
#pragma unroll
      for (int o = 0; o < OC; ++o) {   // loads first (clamped addresses, no control flow between them)
        const int oo = min(oc * OC + o, a.O - 1);
        const int64_t bc = min(b, a.B - 1);
        const size_t at = a.root ? (size_t)bc * a.O + oo : ((size_t)p * a.O + oo) * a.Bp + bc;
        q[s][o] = a.y[at];
        gq[o] = a.gy ? a.gy[at] : 1.f;
      }
#pragma unroll
      for (int o = 0; o < OC; ++o) {
        const float yv = q[s][o], g = gq[o];
        const bool live = oc * OC + o < a.O && b < a.B && fabsf(yv) <= FLT_MAX && g != 0.f;
        const float qv = live ? g * __expf(fminf(ml[s] + mr[s] - yv, 80.f)) : 0.f;
        q[s][o] = qv;
        q_sm[row * QS + o] = qv;
      }
    }
    // ---- phase A ----
    const float* __restrict__ wp = a.wsoft + ((size_t)p * a.nOc + oc) * Kin2 * OC;
    for (int i0 = 0; i0 < Kin; i0 += a.rows_per_chunk) {
      const int i1 = min(Kin, i0 + a.rows_per_chunk);
      __syncthreads();
      for (int t = tid; t < (i1 - i0) * Kin * OC; t += kBwdThreads) wsm[t] = __ldg(wp + (size_t)i0 * Kin * OC + t);
      __syncthreads();
      for (int i = i0; i < i1; ++i) {
        float eli[ST], si[ST];
#pragma unroll
        for (int s = 0; s < ST; ++s) { eli[s] = el[(tid + s * kBwdThreads) * KS + i]; si[s] = 0.f; }
        const float* wrow = wsm + (size_t)(i - i0) * Kin * OC;
        for (int j = 0; j < Kin; ++j) {
          float w[OC];
          load_row_smem<OC>(wrow + j * OC, w);
#pragma unroll
          for (int s = 0; s < ST; ++s) {
            const int row = tid + s * kBwdThreads;
            float T = 0.f;
#pragma unroll
            for (int o = 0; o < OC; ++o) T = fmaf(q[s][o], w[o], T);
            si[s] = fmaf(er[row * KS + j], T, si[s]);
            gr[row * KS + j] = fmaf(eli[s], T, gr[row * KS + j]);
          }
        }
#pragma unroll
        for (int s = 0; s < ST; ++s) {
          const int row = tid + s * kBwdThreads;
          gl[row * KS + i] = fmaf(eli[s], si[s], gl[row * KS + i]);
        }
      }
    }
    __syncthreads();  // q_sm, el, er complete for every sample of the CTA
    // ---- phase B ----
    if (a.wstat) {
      float* __restrict__ ws = a.wstat + ((size_t)p * a.nOc + oc) * Kin2 * OC;
      const int s_begin = warp * (NS / 4), s_end = s_begin + NS / 4;
      for (int blk = 0; blk < Kin2; blk += 128) {
        int ii[4], jj[4];
        bool ok[4];
        float m[4][OC];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int ij = blk + lane + 32 * t;
          ok[t] = ij < Kin2;
          ii[t] = ok[t] ? ij / Kin : 0;
          jj[t] = ok[t] ? ij % Kin : 0;
#pragma unroll
          for (int o = 0; o < OC; ++o) m[t][o] = 0.f;
        }
        for (int srow = s_begin; srow < s_end; ++srow) {
          float qv[QS];
          load_row_smem<QS>(q_sm + (size_t)srow * QS, qv);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float pe = el[srow * KS + ii[t]] * er[srow * KS + jj[t]];
#pragma unroll
            for (int o = 0; o < OC; ++o) m[t][o] = fmaf(qv[o], pe, m[t][o]);
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (ok[t]) {
            const int ij = blk + lane + 32 * t;
#pragma unroll
            for (int o = 0; o < OC; ++o)
              if (m[t][o] != 0.f) atomicAdd(ws + (size_t)ij * OC + o, m[t][o]);
          }
      }
    }
    __syncthreads();  // before q_sm is overwritten by the next output chunk
  }

  if (a.gin) {
    float* __restrict__ gl_out = a.gin + (size_t)(2 * p) * Kin * a.Bp;
    float* __restrict__ gr_out = gl_out + (size_t)Kin * a.Bp;
#pragma unroll
    for (int s = 0; s < ST; ++s) {
      const int row = tid + s * kBwdThreads;
      const int64_t b = base + row;
      if (b >= a.Bp) continue;
      for (int k = 0; k < Kin; ++k) {
        gl_out[(size_t)k * a.Bp + b] = gl[row * KS + k];
        gr_out[(size_t)k * a.Bp + b] = gr[row * KS + k] * er[row * KS + k];
      }
    }
  }
}

// Register-resident variant for the common sizes (Kin, OC in {2,4,8,10,16}, one output chunk): a thread keeps the
// right-hand exps and the dL/dr accumulators of its ST samples in registers (j loop fully unrolled), the partition's
// whole weight block sits in shared memory; only what phase B needs (q, el, er) goes through shared memory.
template <int OC, int KIN, int ST, bool PB2 = false>
__global__ void __launch_bounds__(kBwdThreads) ratspn_einsum_bwd_reg_kernel(const EinsumBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NS = kBwdThreads * ST;
  constexpr int QS = (OC + 3) / 4 * 4;
  constexpr int K2 = KIN * KIN, KS = KIN | 1;
  float* q_sm = sm;                          // [NS][QS]   (16-byte aligned rows)
  float* el = q_sm + (size_t)NS * QS;        // [NS][KS]
  float* er = el + (size_t)NS * KS;
  float* wsm = er + (size_t)NS * KS;         // [K2][OC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.y;
  const int64_t base = (int64_t)blockIdx.x * NS;
  const float* __restrict__ lin = a.in + (size_t)(2 * p) * KIN * a.Bp;
  const float* __restrict__ rin = lin + (size_t)KIN * a.Bp;
  const float* __restrict__ wp = a.wsoft + (size_t)p * K2 * OC;
  for (int t = tid; t < K2 * OC / 2; t += kBwdThreads)
    reinterpret_cast<float2*>(wsm)[t] = __ldg(reinterpret_cast<const float2*>(wp) + t);

  float elr[ST][KIN], err[ST][KIN], gr[ST][KIN], q[ST][OC];
  // every global load of the thread is issued before the first use (clamped addresses instead of predicated
  // loads, no control flow in between): one memory latency per CTA instead of one per output
  float gq[ST][OC];
#pragma unroll
  for (int s = 0; s < ST; ++s) {
    const int64_t b = min(base + tid + s * kBwdThreads, a.B - 1);
#pragma unroll
    for (int k = 0; k < KIN; ++k) {
      elr[s][k] = lin[(size_t)k * a.Bp + b];
      err[s][k] = rin[(size_t)k * a.Bp + b];
    }
#pragma unroll
    for (int o = 0; o < OC; ++o) {
      const int oc = min(o, a.O - 1);
      const size_t at = a.root ? (size_t)b * a.O + oc : ((size_t)p * a.O + oc) * a.Bp + b;
      q[s][o] = a.y[at];
      gq[s][o] = a.gy ? a.gy[at] : 1.f;
    }
  }
#pragma unroll
  for (int s = 0; s < ST; ++s) {
    const int row = tid + s * kBwdThreads;
    const bool inb = base + row < a.B;
    float vl = -INFINITY, vr = -INFINITY;
#pragma unroll
    for (int k = 0; k < KIN; ++k) {
      if (!inb) { elr[s][k] = 0.f; err[s][k] = 0.f; }
      vl = fmaxf(vl, elr[s][k]); vr = fmaxf(vr, err[s][k]);
    }
    const float ml = (fabsf(vl) <= FLT_MAX) ? vl : 0.f;
    const float mr = (fabsf(vr) <= FLT_MAX) ? vr : 0.f;
#pragma unroll
    for (int k = 0; k < KIN; ++k) {
      elr[s][k] = __expf(elr[s][k] - ml);
      err[s][k] = __expf(err[s][k] - mr);
      el[row * KS + k] = elr[s][k];
      er[row * KS + k] = err[s][k];
      gr[s][k] = 0.f;
    }
#pragma unroll
    for (int o = 0; o < OC; ++o) {
      const float yv = q[s][o], g = gq[s][o];
      const bool live = o < a.O && inb && fabsf(yv) <= FLT_MAX && g != 0.f;
      const float qv = live ? g * __expf(fminf(ml + mr - yv, 80.f)) : 0.f;
      q[s][o] = qv;
      q_sm[row * QS + o] = qv;
    }
  }
  __syncthreads();  // weights, q, el, er of every sample of the CTA are staged
  // ---- phase A: dL/dl_i = el_i sum_j er_j T_ij,  dL/dr_j = er_j sum_i el_i T_ij,  T_ij = sum_o q_o w[ij][o] ----
  float* __restrict__ gl_out = a.gin ? a.gin + (size_t)(2 * p) * KIN * a.Bp : nullptr;
#pragma unroll 1
  for (int i = 0; i < KIN; ++i) {
    float si[ST];
#pragma unroll
    for (int s = 0; s < ST; ++s) si[s] = 0.f;
    const float* wrow = wsm + i * KIN * OC;
#pragma unroll
    for (int j = 0; j < KIN; ++j) {
      float w[OC];
      load_row_smem<OC>(wrow + j * OC, w);
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        // T = sum_o q_o w_o over output pairs with packed FFMA2 (OC is even), the two halves added at the end
        float2 T2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < OC / 2; ++h)
          T2 = __ffma2_rn(make_float2(q[s][2 * h], q[s][2 * h + 1]), make_float2(w[2 * h], w[2 * h + 1]), T2);
        const float T = T2.x + T2.y;
        si[s] = fmaf(err[s][j], T, si[s]);
        // el_i as a runtime-indexed register would spill: take it from shared memory (conflict-free, KS odd)
        gr[s][j] = fmaf(el[(tid + s * kBwdThreads) * KS + i], T, gr[s][j]);
      }
    }
    if (gl_out) {
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        const int row = tid + s * kBwdThreads;
        const int64_t b = base + row;
        if (b < a.Bp) gl_out[(size_t)i * a.Bp + b] = el[row * KS + i] * si[s];
      }
    }
  }
  if (gl_out) {
    float* __restrict__ gr_out = gl_out + (size_t)KIN * a.Bp;
#pragma unroll
    for (int s = 0; s < ST; ++s) {
      const int64_t b = base + tid + s * kBwdThreads;
      if (b >= a.Bp) continue;
#pragma unroll
      for (int k = 0; k < KIN; ++k) gr_out[(size_t)k * a.Bp + b] = gr[s][k] * err[s][k];
    }
  }
  // ---- phase B: posterior counts M[ij][o] += q_o el_i er_j ----
  if constexpr (PB2) {
    // Register tile: a lane owns one left index i and all (j, o) -- KIN*OC accumulators -- and every KIN lanes
    // take a different sample, so a step of the warp covers 32/KIN samples with {1 el, KIN er, OC q} shared-memory
    // loads against KIN*OC/2 packed FFMA2 (lanes = products needed 11 loads per 44 FMA and left the kernel
    // waiting on the shared-memory queue).  The sample slices are summed with shuffles, the four warps through
    // shared memory, so a CTA issues K2*OC atomics instead of 4*K2*OC.
    if (a.wstat) {   // CTA-uniform
      constexpr int NSL = 32 / KIN, OH = OC / 2;
      static_assert(OC % 2 == 0 && 4 * K2 * OC <= NS * (QS + KS), "partials must fit the q/el staging area");
      const int li = lane % KIN, sl = lane / KIN;
      const bool act = lane < NSL * KIN;
      float2 m2[KIN][OH];
#pragma unroll
      for (int j = 0; j < KIN; ++j)
#pragma unroll
        for (int h = 0; h < OH; ++h) m2[j][h] = make_float2(0.f, 0.f);
      const int s_begin = warp * (NS / 4), s_end = s_begin + NS / 4;
      if (act) {
        for (int srow = s_begin + sl; srow < s_end; srow += NSL) {
          float qv[QS];
          load_row_smem<QS>(q_sm + (size_t)srow * QS, qv);
          const float e_l = el[srow * KS + li];
#pragma unroll
          for (int j = 0; j < KIN; ++j) {
            const float pe = e_l * er[srow * KS + j];
            const float2 pe2 = make_float2(pe, pe);
#pragma unroll
            for (int h = 0; h < OH; ++h) m2[j][h] = __ffma2_rn(make_float2(qv[2 * h], qv[2 * h + 1]), pe2, m2[j][h]);
          }
        }
      }
      // slices -> lane li (sl == 0)
#pragma unroll
      for (int j = 0; j < KIN; ++j)
#pragma unroll
        for (int h = 0; h < OH; ++h) {
          float vx = m2[j][h].x, vy = m2[j][h].y;
#pragma unroll
          for (int dd = 1; dd < NSL; ++dd) {
            const float ox = __shfl_sync(0xffffffffu, m2[j][h].x, (lane + dd * KIN) & 31);
            const float oy = __shfl_sync(0xffffffffu, m2[j][h].y, (lane + dd * KIN) & 31);
            if (lane + dd * KIN < NSL * KIN) { vx += ox; vy += oy; }
          }
          m2[j][h] = make_float2(vx, vy);
        }
      __syncthreads();                      // every warp is done with q / el / er: reuse them for the partials
      float* part = q_sm + (size_t)warp * (K2 * OC);
      if (lane < KIN) {
#pragma unroll
        for (int j = 0; j < KIN; ++j)
#pragma unroll
          for (int h = 0; h < OH; ++h)
            *reinterpret_cast<float2*>(part + (li * KIN + j) * OC + 2 * h) = m2[j][h];
      }
      __syncthreads();
      float* __restrict__ ws = a.wstat + (size_t)p * K2 * OC;
      for (int t = tid; t < K2 * OC; t += kBwdThreads) {
        const float v = (q_sm[t] + q_sm[K2 * OC + t]) + (q_sm[2 * K2 * OC + t] + q_sm[3 * K2 * OC + t]);
        if (v != 0.f) atomicAdd(ws + t, v);
      }
    }
  } else if (a.wstat) {   // lanes = product index ij, warps split the samples
    float* __restrict__ ws = a.wstat + (size_t)p * K2 * OC;
    const int s_begin = warp * (NS / 4), s_end = s_begin + NS / 4;
    for (int blk = 0; blk < K2; blk += 128) {
      int ii[4], jj[4];
      bool ok[4];
      float m[4][OC];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int ij = blk + lane + 32 * t;
        ok[t] = ij < K2;
        ii[t] = ok[t] ? ij / KIN : 0;
        jj[t] = ok[t] ? ij % KIN : 0;
#pragma unroll
        for (int o = 0; o < OC; ++o) m[t][o] = 0.f;
      }
      for (int srow = s_begin; srow < s_end; ++srow) {
        float qv[QS];
        load_row_smem<QS>(q_sm + (size_t)srow * QS, qv);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float pe = el[srow * KS + ii[t]] * er[srow * KS + jj[t]];
#pragma unroll
          for (int o = 0; o < OC; ++o) m[t][o] = fmaf(qv[o], pe, m[t][o]);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (ok[t]) {
          const int ij = blk + lane + 32 * t;
#pragma unroll
          for (int o = 0; o < OC; ++o)
            if (m[t][o] != 0.f) atomicAdd(ws + (size_t)ij * OC + o, m[t][o]);
        }
    }
  }
}

// Posterior counts -> gradient w.r.t. the raw logits (or the counts themselves for EM).
//   mode 0: rows (p, o) of length Kin2, dst (P, O, Kin2);  mode 1: rows c of length P*Kin2, dst (C, P*Kin2)
__global__ void ratspn_weight_finalize_kernel(const float* __restrict__ wsoft, const float* __restrict__ wstat,
                                              int mode, int P, int O, int Kin2, int OC, int nOc, int em,
                                              float* __restrict__ dst) {
  __shared__ float red[32];
  int p_row, o;
  int64_t len;
  if (mode == 0) { p_row = blockIdx.x / O; o = blockIdx.x % O; len = Kin2; }
  else           { p_row = 0;              o = blockIdx.x;     len = (int64_t)P * Kin2; }
  float* out = (mode == 0) ? dst + ((size_t)p_row * O + o) * Kin2 : dst + (size_t)o * len;
  const int oc = o / OC, ok = o % OC;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
    const int p = (mode == 0) ? p_row : (int)(i / Kin2);
    const int ij = (mode == 0) ? (int)i : (int)(i % Kin2);
    const size_t at = (((size_t)p * nOc + oc) * Kin2 + ij) * OC + ok;
    s += wsoft[at] * wstat[at];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  s = warp_sum(s);
  __syncthreads();
  if (threadIdx.x == 0) red[0] = s;
  __syncthreads();
  const float total = red[0];
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
    const int p = (mode == 0) ? p_row : (int)(i / Kin2);
    const int ij = (mode == 0) ? (int)i : (int)(i % Kin2);
    const size_t at = (((size_t)p * nOc + oc) * Kin2 + ij) * OC + ok;
    const float n = wsoft[at] * wstat[at];
    out[i] += em ? n : n - wsoft[at] * total;
  }
}

// =================================================================================================
// Leaf moments  S1,S2[g,k,d] = sum_b g[b,g,k] * {x, x^2},  S0tot[g,k] = sum_b g[b,g,k],
// Snan[g,k,d] = sum over samples whose x is NaN/inf (their term has zero gradient: ratspn.py:103)
// =================================================================================================
struct LeafBwdArgs {
  const float* x;
  const int32_t* mask;
  const int32_t* region_len;
  const float* g0;  // [G0][K][Bp]
  float* s1; float* s2; float* snan; float* s0tot;
  int64_t B, Bp, samples_per_cta;
  int D, G0, K, dim, TS, nKc;
  const int* only_if;   // NULL, or device flag: the kernel runs only when it is != 0 (fallback behind the GEMM path)
};

template <int KC> struct LeafBwdJD { static constexpr int value = (KC <= 10) ? 4 : 2; };

// CTA = 8 consecutive regions (warp = region) x a batch slice; lanes = region dims (JD per lane);
// the moment accumulators stay in registers over the whole slice, samples stream through smem.
template <int KC, int KIND>
__global__ void __launch_bounds__(256) ratspn_leaf_bwd_stats_kernel(const LeafBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int JD = LeafBwdJD<KC>::value;
  constexpr int KS = (KC + 3) / 4 * 4;
  constexpr bool GAUSS = (KIND == DPK_LEAF_GAUSSIAN);
  if (a.only_if && !__ldg(a.only_if)) return;
  const int TS = a.TS;
  float* gs = sm;                         // [8][TS][KS]  (first: rows must stay 16-byte aligned)
  float* xs = gs + (size_t)8 * TS * KS;   // [TS][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = blockIdx.x * 8 + warp;
  const bool live = r < a.G0;
  const int kc0 = (blockIdx.z % a.nKc) * KC;
  const int dblk = (blockIdx.z / a.nKc) * 32 * JD;
  const int len = live ? a.region_len[r] : 0;
  int f[JD];
  bool ok[JD];
#pragma unroll
  for (int t = 0; t < JD; ++t) {
    const int d = dblk + lane + 32 * t;
    ok[t] = d < len;
    f[t] = ok[t] ? a.mask[(size_t)r * a.dim + d] : 0;
  }
  float S1[JD][KC], S2[GAUSS ? JD : 1][KC], S0[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    S0[k] = 0.f;
#pragma unroll
    for (int t = 0; t < JD; ++t) { S1[t][k] = 0.f; if (GAUSS) S2[t][k] = 0.f; }
  }

  const int64_t b_begin = (int64_t)blockIdx.y * a.samples_per_cta;
  const int64_t b_end = min((long long)a.B, (long long)(b_begin + a.samples_per_cta));
  for (int64_t b0 = b_begin; b0 < b_end; b0 += TS) {
    const int nb = (int)min((long long)TS, (long long)(b_end - b0));
    __syncthreads();
    bool bad = false;
    const float* __restrict__ src = a.x + b0 * a.D;
    for (int idx = tid; idx < TS * a.D; idx += 256) {
      const float v = (idx < nb * a.D) ? __ldg(src + idx) : 0.f;
      bad |= !(fabsf(v) <= FLT_MAX);
      xs[idx] = v;
    }
    for (int idx = tid; idx < 8 * KC * TS; idx += 256) {
      const int s = idx % TS, k = (idx / TS) % KC, w = idx / (TS * KC);
      const int rr = blockIdx.x * 8 + w, kk = kc0 + k;
      gs[((size_t)w * TS + s) * KS + k] =
          (rr < a.G0 && kk < a.K && s < nb) ? a.g0[((size_t)rr * a.K + kk) * a.Bp + b0 + s] : 0.f;
    }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (!live) continue;
    for (int s = 0; s < nb; ++s) {
      float g[KS];
      load_row_smem<KS>(gs + ((size_t)warp * TS + s) * KS, g);
#pragma unroll
      for (int k = 0; k < KC; ++k) S0[k] += g[k];
#pragma unroll
      for (int t = 0; t < JD; ++t) {
        if (!ok[t]) continue;
        float xv = xs[(size_t)s * a.D + f[t]];
        if (any_bad && !(fabsf(xv) <= FLT_MAX)) {
          const int d = dblk + lane + 32 * t;
#pragma unroll
          for (int k = 0; k < KC; ++k)
            if (kc0 + k < a.K && g[k] != 0.f) atomicAdd(a.snan + ((size_t)r * a.K + kc0 + k) * a.dim + d, g[k]);
          xv = 0.f;
        }
        const float x2 = xv * xv;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          S1[t][k] = fmaf(g[k], xv, S1[t][k]);
          if (GAUSS) S2[t][k] = fmaf(g[k], x2, S2[t][k]);
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int kk = kc0 + k;
    if (kk >= a.K) continue;
    if (lane == 0 && dblk == 0) atomicAdd(a.s0tot + (size_t)r * a.K + kk, S0[k]);
#pragma unroll
    for (int t = 0; t < JD; ++t) {
      if (!ok[t]) continue;
      const size_t at = ((size_t)r * a.K + kk) * a.dim + dblk + lane + 32 * t;
      atomicAdd(a.s1 + at, S1[t][k]);
      if (GAUSS) atomicAdd(a.s2 + at, S2[t][k]);
    }
  }
}

// moments -> parameter gradients (accumulated) or EM statistics
template <int KIND>
__global__ void ratspn_leaf_finalize_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                            const int32_t* __restrict__ region_len, const float* __restrict__ s1,
                                            const float* __restrict__ s2, const float* __restrict__ snan,
                                            const float* __restrict__ s0tot, int G0, int K, int dim, int em,
                                            float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2) {
  const int64_t total = (int64_t)G0 * K * dim;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(idx % dim);
    const int64_t gk = idx / dim;
    const int g = (int)(gk / K);
    if (d >= region_len[g]) continue;
    const float S0 = s0tot[gk] - snan[idx], S1 = s1[idx];
    if (em) {
      if (o0) o0[idx] += S0;
      if (o1) o1[idx] += S1;
      if (o2 && KIND == DPK_LEAF_GAUSSIAN) o2[idx] += s2[idx];
      continue;
    }
    if (KIND == DPK_LEAF_GAUSSIAN) {
      const float mu = p0[idx], sg = p1 ? p1[idx] : 1.0f, S2 = s2[idx];
      const float inv = 1.0f / sg, inv2 = inv * inv;
      if (o0) o0[idx] += (S1 - mu * S0) * inv2;
      if (o1) o1[idx] += (S2 - 2.0f * mu * S1 + mu * mu * S0) * inv2 * inv - S0 * inv;
    } else {
      const float lg = p0[idx];
      if (o0) o0[idx] += S1 - S0 / (1.0f + expf(-lg));
    }
  }
}

// =================================================================================================
// d/dx: same sweep as the forward leaf kernel (lanes = samples over a transposed x tile), the
// per-feature contributions of the R repetitions are summed in a shared-memory gx tile.
// =================================================================================================
struct LeafBwdXArgs {
  const float* x;
  const int32_t* mask;
  const int32_t* region_len;
  const float* tab;  // forward table (chunked layout, see ratspn_plan.cuh)
  int CH, NCH, CHP, CF;
  const float* g0;   // [G0][K][Bp]
  float* gx;         // (B, D), accumulated
  int64_t B, Bp;
  int D, G0, K, dim, nKc;
};

template <int KC, int KIND, bool STAGE>
__global__ void __launch_bounds__(256) ratspn_leaf_bwd_x_kernel(const LeafBwdXArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NP = (KIND == DPK_LEAF_GAUSSIAN) ? 2 : 1;   // KIND may also be kLeafGaussUnit (scale == 1)
  constexpr int NPK = (NP * KC + 3) / 4 * 4;
  float* xs = sm;                       // [D][32] swizzled
  float* gxs = sm + (size_t)a.D * 32;   // [D][32] swizzled
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  const int64_t b = b0 + lane;
  if constexpr (STAGE) {
    for (int s = warp; s < 32; s += 8) {
      const bool inb = b0 + s < a.B;
      const float* row = a.x + (b0 + s) * a.D;
      for (int ff = lane; ff < a.D; ff += 32) {
        const int at = ff * 32 + (s ^ (ff & 31));
        xs[at] = inb ? __ldg(row + ff) : 0.f;
        gxs[at] = 0.f;
      }
    }
    __syncthreads();
  }
  for (int r = warp; r < a.G0; r += 8) {
    const int len = __ldg(a.region_len + r);
    const int32_t* __restrict__ m = a.mask + (size_t)r * a.dim;
    for (int c = 0; c < a.nKc; ++c) {
      float g[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = c * KC + k;
        g[k] = (kk < a.K && b < a.B) ? a.g0[((size_t)r * a.K + kk) * a.Bp + b] : 0.f;
      }
      const float* __restrict__ block = a.tab + ((size_t)r * a.nKc + c) * a.NCH * a.CF;
      for (int d = 0; d < len; ++d) {
        const int ff = __ldg(m + d);
        float xv;
        if constexpr (STAGE) xv = xs[ff * 32 + (lane ^ (ff & 31))];
        else xv = (b < a.B) ? __ldg(a.x + b * a.D + ff) : 0.f;
        float p[NPK];
        const int chk = d / a.CH;
        load_row<NPK>(block + (size_t)chk * a.CF + a.CHP + (d - chk * a.CH) * NPK, p);
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          if constexpr (KIND == DPK_LEAF_GAUSSIAN) {
            const float t = fmaf(xv, p[k], p[KC + k]);   // (x - mu) / sigma
            acc = fmaf(g[k], -t * p[k], acc);            // d/dx of -t^2/2
          } else if constexpr (KIND == kLeafGaussUnit) {
            acc = fmaf(g[k], -(xv + p[k]), acc);         // sigma == 1
          } else {
            acc = fmaf(g[k], p[k], acc);                 // d/dx of x*logit - softplus
          }
        }
        if (!(fabsf(xv) <= FLT_MAX)) acc = 0.f;          // nan_to_num'ed terms have zero gradient
        if constexpr (STAGE) atomicAdd(gxs + ff * 32 + (lane ^ (ff & 31)), acc);
        else if (b < a.B && acc != 0.f) atomicAdd(a.gx + b * a.D + ff, acc);
      }
    }
  }
  if constexpr (STAGE) {
    __syncthreads();
    for (int s = warp; s < 32; s += 8) {
      if (b0 + s >= a.B) continue;
      float* row = a.gx + (b0 + s) * a.D;
      for (int ff = lane; ff < a.D; ff += 32) row[ff] += gxs[ff * 32 + (s ^ (ff & 31))];
    }
  }
}

// =================================================================================================
// Host-side drivers
// =================================================================================================
template <int OC, int ST>
static int launch_einsum_bwd_t(const EinsumBwdArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = ratspn_einsum_bwd_kernel<OC, ST>;
  if (smem > 48 * 1024)
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(CAT_BWD_EINSUM, st);
  kern<<<grid, kBwdThreads, smem, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_einsum_bwd_kernel");
  return DPK_OK;
}

template <int OC, int KIN>
static int launch_einsum_bwd_reg_t(const EinsumBwdArgs& a, cudaStream_t st) {
  constexpr int ST = 2, QS = (OC + 3) / 4 * 4, KS = KIN | 1;
  // register-tiled posterior counts while the KIN*OC accumulators fit (K, O <= 10); DPK_BWD_PHASEB=0: lanes = products
  constexpr bool kTile = KIN * OC <= 100 && OC % 2 == 0 && 4 * KIN * KIN * OC <= kBwdThreads * ST * (QS + KS);
  auto kern = ratspn_einsum_bwd_reg_kernel<OC, KIN, ST>;
  if constexpr (kTile) {
    if (env_int("DPK_BWD_PHASEB", 1) != 0) kern = ratspn_einsum_bwd_reg_kernel<OC, KIN, ST, true>;
  }
  const size_t smem = ((size_t)kBwdThreads * ST * (QS + 2 * KS) + (size_t)KIN * KIN * OC) * 4;
  if (smem > 48 * 1024)
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(a.Bp, kBwdThreads * ST), (unsigned)a.P);
  ProfScope prof(CAT_BWD_EINSUM, st);
  kern<<<grid, kBwdThreads, smem, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_einsum_bwd_reg_kernel");
  return DPK_OK;
}

template <int OC>
static int launch_einsum_bwd_reg_k(const EinsumBwdArgs& a, cudaStream_t st) {
  switch (a.Kin) {
    case 2: return launch_einsum_bwd_reg_t<OC, 2>(a, st);
    case 4: return launch_einsum_bwd_reg_t<OC, 4>(a, st);
    case 8: return launch_einsum_bwd_reg_t<OC, 8>(a, st);
    case 10: return launch_einsum_bwd_reg_t<OC, 10>(a, st);
    case 16: return launch_einsum_bwd_reg_t<OC, 16>(a, st);
  }
  return 1;  // not covered
}

static int launch_einsum_bwd(EinsumBwdArgs a, int OC, cudaStream_t st) {
  const size_t smem_max = (size_t)max_dynamic_smem();
  if (a.nOc == 1 && a.O <= OC && a.B >= 2 * kBwdThreads && env_int("DPK_BWD_GENERIC", 0) == 0) {
    int rc = 1;   // register-resident fast path for the common sizes
    switch (OC) {
      case 2: rc = launch_einsum_bwd_reg_k<2>(a, st); break;
      case 4: rc = launch_einsum_bwd_reg_k<4>(a, st); break;
      case 8: rc = launch_einsum_bwd_reg_k<8>(a, st); break;
      case 10: rc = launch_einsum_bwd_reg_k<10>(a, st); break;
      case 16: rc = launch_einsum_bwd_reg_k<16>(a, st); break;
    }
    if (rc <= 0) return rc;
  }
  int rows = (int)std::max<size_t>(1, 8192 / ((size_t)a.Kin * OC * 4));
  rows = std::min(rows, a.Kin);
  a.rows_per_chunk = rows;
  const size_t wbytes = (size_t)rows * a.Kin * OC * 4;
  const int KS = a.Kin | 1, QS = (OC + 3) / 4 * 4;
  auto need = [&](int ST) { return (size_t)kBwdThreads * ST * (QS + 4 * KS) * 4 + wbytes; };
  // samples per thread: 4 amortises the weight broadcasts best but leaves one 4-warp CTA per SM (115 KB of shared
  // memory at K = O = 10); 2 keeps three CTAs resident -- measured faster on B200 (DPK_BWD_ST overrides)
  int ST = env_int("DPK_BWD_ST", 2);
  if (ST != 1 && ST != 2 && ST != 4) ST = 2;
  if (need(ST) > 160 * 1024 || a.B < (int64_t)ST * kBwdThreads) ST = 1;
  const size_t smem = need(ST);
  if (smem > smem_max) return set_error(DPK_E_ARG, "einsum backward with %d inputs per region does not fit shared memory", a.Kin);
  dim3 grid((unsigned)ceil_div(a.Bp, kBwdThreads * ST), (unsigned)a.P);
#define DPK_CASE(oc)                                                                              \
  case oc:                                                                                        \
    return (ST == 4) ? launch_einsum_bwd_t<oc, 4>(a, grid, smem, st)                              \
           : (ST == 2) ? launch_einsum_bwd_t<oc, 2>(a, grid, smem, st)                            \
                       : launch_einsum_bwd_t<oc, 1>(a, grid, smem, st);
  switch (OC) { DPK_CASE(2) DPK_CASE(4) DPK_CASE(8) DPK_CASE(10) DPK_CASE(16) }
#undef DPK_CASE
  return set_error(DPK_E_ARG, "unsupported output chunk %d", OC);
}

template <int KC, int KIND>
static int launch_leaf_stats_t(LeafBwdArgs a, cudaStream_t st) {
  constexpr int JD = LeafBwdJD<KC>::value;
  constexpr int KS = (KC + 3) / 4 * 4;
  const size_t smem_max = (size_t)max_dynamic_smem();
  int TS = 32;
  auto need = [&](int ts) { return ((size_t)ts * a.D + (size_t)8 * ts * KS) * 4; };
  while (TS > 1 && need(TS) > 100 * 1024) TS >>= 1;
  if (need(TS) > smem_max) return set_error(DPK_E_ARG, "in_features %d too large for the leaf backward kernel", a.D);
  a.TS = TS;
  const int groups = (int)ceil_div(a.G0, 8);
  const int passes = a.nKc * (int)ceil_div(a.dim, 32 * JD);
  // enough batch slices to fill the machine twice, but at least 4*TS samples per CTA
  int64_t slices = std::max<int64_t>(1, ceil_div(2 * sm_count(), (int64_t)groups * passes));
  slices = std::min<int64_t>(slices, std::max<int64_t>(1, a.B / (4 * TS)));
  a.samples_per_cta = round_up(ceil_div(a.B, slices), TS);
  slices = ceil_div(a.B, a.samples_per_cta);
  auto kern = ratspn_leaf_bwd_stats_kernel<KC, KIND>;
  const size_t smem = need(TS);
  if (smem > 48 * 1024)
    DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(CAT_BWD_LEAF, st);
  kern<<<dim3(groups, (unsigned)slices, passes), 256, smem, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_leaf_bwd_stats_kernel");
  return DPK_OK;
}

template <int KIND>
static int launch_leaf_stats(int KC, const LeafBwdArgs& a, cudaStream_t st) {
  switch (KC) {
    case 2: return launch_leaf_stats_t<2, KIND>(a, st);
    case 4: return launch_leaf_stats_t<4, KIND>(a, st);
    case 8: return launch_leaf_stats_t<8, KIND>(a, st);
    case 10: return launch_leaf_stats_t<10, KIND>(a, st);
    case 16: return launch_leaf_stats_t<16, KIND>(a, st);
  }
  return set_error(DPK_E_ARG, "unsupported leaf channel chunk %d", KC);
}

template <int KC, int KIND>
static int launch_leaf_x_t(const LeafBwdXArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)a.D * 64 * 4;
  dim3 grid((unsigned)ceil_div(a.B, 32));
  ProfScope prof(CAT_BWD_LEAF, st);
  if (smem <= (size_t)max_dynamic_smem()) {
    auto kern = ratspn_leaf_bwd_x_kernel<KC, KIND, true>;
    if (smem > 48 * 1024)
      DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(a);
  } else {
    ratspn_leaf_bwd_x_kernel<KC, KIND, false><<<grid, 256, 0, st>>>(a);
  }
  DPK_LAUNCH_CHECK("ratspn_leaf_bwd_x_kernel");
  return DPK_OK;
}

template <int KIND>
static int launch_leaf_x(int KC, const LeafBwdXArgs& a, cudaStream_t st) {
  switch (KC) {
    case 2: return launch_leaf_x_t<2, KIND>(a, st);
    case 4: return launch_leaf_x_t<4, KIND>(a, st);
    case 8: return launch_leaf_x_t<8, KIND>(a, st);
    case 10: return launch_leaf_x_t<10, KIND>(a, st);
    case 16: return launch_leaf_x_t<16, KIND>(a, st);
  }
  return set_error(DPK_E_ARG, "unsupported leaf channel chunk %d", KC);
}

struct BwdTargets {
  int em;
  float* sum_w[DPK_MAX_LEVELS];
  float* root_w;
  float* leaf0; float* leaf1; float* leaf2;  // grads (loc|logits, scale) or EM (s0, s1, s2)
  float* gx;
};

static int run_backward(const dpk_ratspn_desc* d, const RatPlan& p, const float* x, const float* out,
                        const float* gout, const BwdTargets& t, float* ws, cudaStream_t st) {
  DPK_CUDA_TRY(cudaMemsetAsync(ws + p.stat_begin, 0, (p.stat_end - p.stat_begin) * sizeof(float), st));
  const int top = p.depth - 1;
  bool want_leaf = t.leaf0 || t.leaf1 || t.leaf2 || t.gx;
  bool want_below_root = want_leaf;
  for (int e = 0; e < p.n_sum; ++e) want_below_root |= (t.sum_w[e] != nullptr);
  {
    EinsumBwdArgs a;
    a.in = ws + p.off_act[top]; a.wsoft = ws + p.off_rsoft; a.y = out; a.gy = gout;
    a.gin = want_below_root ? ws + p.off_gact[top] : nullptr;
    a.wstat = t.root_w ? ws + p.off_rstat : nullptr;
    a.B = p.B; a.Bp = p.Bp; a.P = p.R; a.Kin = p.act_ch[top]; a.O = p.C; a.nOc = p.cc.count; a.root = 1;
    if (a.gin || a.wstat) { int rc = launch_einsum_bwd(a, p.cc.chunk, st); if (rc) return rc; }
  }
  for (int e = p.n_sum - 1; e >= 0; --e) {
    bool want_below = want_leaf;
    for (int e2 = 0; e2 < e; ++e2) want_below |= (t.sum_w[e2] != nullptr);
    EinsumBwdArgs a;
    a.in = ws + p.off_act[e]; a.wsoft = ws + p.off_wsoft[e]; a.y = ws + p.off_act[e + 1]; a.gy = ws + p.off_gact[e + 1];
    a.gin = want_below ? ws + p.off_gact[e] : nullptr;
    a.wstat = t.sum_w[e] ? ws + p.off_wstat[e] : nullptr;
    a.B = p.B; a.Bp = p.Bp; a.P = p.act_regions[e] / 2; a.Kin = p.act_ch[e]; a.O = p.O; a.nOc = p.oc.count; a.root = 0;
    if (a.gin || a.wstat) { int rc = launch_einsum_bwd(a, p.oc.chunk, st); if (rc) return rc; }
  }
  // weights: counts -> gradients / EM counts
  for (int e = 0; e < p.n_sum; ++e) {
    if (!t.sum_w[e]) continue;
    const int P = p.act_regions[e] / 2, kin2 = p.act_ch[e] * p.act_ch[e];
    ProfScope prof(CAT_FINALIZE, st);
    ratspn_weight_finalize_kernel<<<P * p.O, 128, 0, st>>>(ws + p.off_wsoft[e], ws + p.off_wstat[e], 0, P, p.O, kin2,
                                                            p.oc.chunk, p.oc.count, t.em, t.sum_w[e]);
    DPK_LAUNCH_CHECK("ratspn_weight_finalize_kernel");
  }
  if (t.root_w) {
    const int kin2 = p.act_ch[top] * p.act_ch[top];
    ProfScope prof(CAT_FINALIZE, st);
    ratspn_weight_finalize_kernel<<<p.C, 256, 0, st>>>(ws + p.off_rsoft, ws + p.off_rstat, 1, p.R, p.C, kin2, p.cc.chunk,
                                                        p.cc.count, t.em, t.root_w);
    DPK_LAUNCH_CHECK("ratspn_weight_finalize_kernel(root)");
  }
  if (t.leaf0 || t.leaf1 || t.leaf2) {
    LeafBwdArgs a;
    a.x = x; a.mask = d->mask; a.region_len = d->region_len; a.g0 = ws + p.off_gact[0];
    a.s1 = ws + p.off_s1; a.s2 = ws + p.off_s2; a.snan = ws + p.off_snan; a.s0tot = ws + p.off_s0tot;
    a.B = p.B; a.Bp = p.Bp; a.D = p.D; a.G0 = p.G0; a.K = p.K; a.dim = p.dim; a.nKc = p.kc.count;
    a.TS = 0; a.samples_per_cta = 0; a.only_if = nullptr;
    if (p.stats_mma && ((uintptr_t)a.g0 & 15) == 0) {
      ProfScope prof(CAT_BWD_LEAF, st, 6);
      int rc2 = ratspn_run_leaf_stats_mma(d, p, x, a.g0, ws, a.s1, a.s2, a.s0tot, &a.only_if, st);
      if (rc2) return rc2;
    }
    int rc = (p.kind == DPK_LEAF_GAUSSIAN) ? launch_leaf_stats<DPK_LEAF_GAUSSIAN>(p.kc.chunk, a, st)
                                           : launch_leaf_stats<DPK_LEAF_BERNOULLI>(p.kc.chunk, a, st);
    if (rc) return rc;
    const int64_t total = (int64_t)p.G0 * p.K * p.dim;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 4096);
    ProfScope prof(CAT_FINALIZE, st);
    if (p.kind == DPK_LEAF_GAUSSIAN)
      ratspn_leaf_finalize_kernel<DPK_LEAF_GAUSSIAN><<<blocks, 256, 0, st>>>(
          d->leaf_p0, d->leaf_p1, d->region_len, a.s1, a.s2, a.snan, a.s0tot, p.G0, p.K, p.dim, t.em, t.leaf0, t.leaf1, t.leaf2);
    else
      ratspn_leaf_finalize_kernel<DPK_LEAF_BERNOULLI><<<blocks, 256, 0, st>>>(
          d->leaf_p0, nullptr, d->region_len, a.s1, a.s2, a.snan, a.s0tot, p.G0, p.K, p.dim, t.em, t.leaf0, t.leaf1, t.leaf2);
    DPK_LAUNCH_CHECK("ratspn_leaf_finalize_kernel");
  }
  if (t.gx && p.stats_mma && env_int("DPK_BWDX_MMA", 1) != 0) {
    // large batches: the transposed leaf GEMM on the tensor cores (0.9 ms instead of 9 ms at config 2)
    ProfScope prof(CAT_BWD_LEAF, st, 6);
    int rc = ratspn_run_leaf_bwd_x_mma(d, p, x, ws + p.off_gact[0], ws, t.gx, st);
    if (rc) return rc;
  } else if (t.gx) {
    LeafBwdXArgs a;
    a.x = x; a.mask = d->mask; a.region_len = d->region_len; a.tab = ws + p.off_tab; a.g0 = ws + p.off_gact[0];
    a.gx = t.gx; a.B = p.B; a.Bp = p.Bp; a.D = p.D; a.G0 = p.G0; a.K = p.K; a.dim = p.dim; a.nKc = p.kc.count;
    a.CH = p.leaf_ch; a.NCH = p.leaf_nch; a.CHP = p.leaf_chp; a.CF = p.leaf_chunk_floats;
    int rc = (p.fwd_kind == DPK_LEAF_GAUSSIAN) ? launch_leaf_x<DPK_LEAF_GAUSSIAN>(p.kc.chunk, a, st)
             : (p.fwd_kind == kLeafGaussUnit)  ? launch_leaf_x<kLeafGaussUnit>(p.kc.chunk, a, st)
                                               : launch_leaf_x<DPK_LEAF_BERNOULLI>(p.kc.chunk, a, st);
    if (rc) return rc;
  }
  return DPK_OK;
}

static int check_bwd_common(const dpk_ratspn_desc* desc, int64_t batch, const float* x, const float* out,
                            void* workspace, size_t workspace_bytes, RatPlan* p) {
  int rc = make_plan(desc, batch, DPK_F_SAVE_ACTIVATIONS, p);
  if (rc) return rc;
  if (!x || !out || !desc->mask || !desc->region_len || !desc->leaf_p0)
    return set_error(DPK_E_ARG, "null pointer argument");
  if (!workspace) return set_error(DPK_E_WORKSPACE, "null workspace");
  if ((uintptr_t)workspace % 256) return set_error(DPK_E_WORKSPACE, "workspace must be 256-byte aligned");
  if (workspace_bytes < p->total_floats * 4)
    return set_error(DPK_E_WORKSPACE, "workspace too small for backward: %zu < %zu bytes (forward must run with "
                     "DPK_F_SAVE_ACTIVATIONS)", workspace_bytes, p->total_floats * 4);
  return DPK_OK;
}

}  // namespace dpk

using namespace dpk;

extern "C" int dpk_ratspn_backward(const dpk_ratspn_desc* desc, const float* x, int64_t batch, const float* out,
                                   const float* grad_out, const dpk_ratspn_grads* grads, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  RatPlan p;
  if (batch == 0) return DPK_OK;
  int rc = check_bwd_common(desc, batch, x, out, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if (!grad_out || !grads) return set_error(DPK_E_ARG, "null grad_out / grads");
  BwdTargets t;
  t.em = 0;
  for (int e = 0; e < DPK_MAX_LEVELS; ++e) t.sum_w[e] = (e < p.n_sum) ? grads->sum_weight[e] : nullptr;
  t.root_w = grads->root_weight;
  t.leaf0 = grads->leaf_p0;
  t.leaf1 = (p.kind == DPK_LEAF_GAUSSIAN) ? grads->leaf_p1 : nullptr;
  t.leaf2 = nullptr;
  t.gx = grads->grad_x;
  return run_backward(desc, p, x, out, grad_out, t, static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}

extern "C" int dpk_ratspn_em_statistics(const dpk_ratspn_desc* desc, const float* x, int64_t batch, const float* out,
                                        const dpk_ratspn_em_stats* stats, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  RatPlan p;
  if (batch == 0) return DPK_OK;
  int rc = check_bwd_common(desc, batch, x, out, workspace, workspace_bytes, &p);
  if (rc) return rc;
  if (!stats) return set_error(DPK_E_ARG, "null stats");
  if (p.C != 1) return set_error(DPK_E_ARG, "EM statistics are defined for out_classes == 1 (got %d)", p.C);
  BwdTargets t;
  t.em = 1;
  for (int e = 0; e < DPK_MAX_LEVELS; ++e) t.sum_w[e] = (e < p.n_sum) ? stats->sum_counts[e] : nullptr;
  t.root_w = stats->root_counts;
  t.leaf0 = stats->s0; t.leaf1 = stats->s1; t.leaf2 = (p.kind == DPK_LEAF_GAUSSIAN) ? stats->s2 : nullptr;
  t.gx = nullptr;
  return run_backward(desc, p, x, out, nullptr, t, static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}
