// ratspn_dropout.cu -- RAT-SPN forward / backward in TRAINING mode with probabilistic dropout.
//
//   RegionGraphLayer.forward  deeprob/spn/layers/ratspn.py:98-100   x[torch.rand_like(x) < dropout] = NaN on the
//                             per-dimension log-densities (B, G0, K, dim), before nan_to_num / pad mask / sum
//   SumLayer.forward          deeprob/spn/layers/ratspn.py:370-372  x[torch.rand_like(x) < dropout] = -inf on the
//                             product layer's output (B, P, K^2), before the weighted logsumexp
// (RootLayer has no dropout.)  The reference draws its Bernoulli variables from torch's RNG stream; a kernel cannot
// reproduce that stream, so parity is distributional: here every draw is a pure function of
// (seed, stream, element index) -- a counter-based generator (SplitMix64 finaliser over the counter) -- which the
// backward regenerates instead of storing masks, and which the tests mirror in NumPy to inject the same masks into
// the oracle.  Streams: 0 = leaf elements ((b*G0+g)*K+k)*dim+d, 1+e = sum level e elements (b*P_e+p)*Kin^2+ij.
//
// Training batches are small (the reference's examples use 100), so these kernels are the plain exact log-domain
// formulation, one thread per output, not the tuned inference path: dropout breaks the factorisation the fused
// kernels rely on (a mask per (i,j) pair of every sample).
#include <algorithm>

#include "ratspn_kernels.cuh"

namespace dpk {

namespace {

__host__ __device__ __forceinline__ uint32_t drop_u24(uint64_t seed, uint32_t stream, uint64_t idx) {
  uint64_t z = (seed ^ ((uint64_t)stream << 56)) + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 40);
}

struct DropPlan {
  int kind, D, depth, R, K, O, C, dim, G0, n_sum;
  int64_t B;
  int regions[DPK_MAX_LEVELS], ch[DPK_MAX_LEVELS];
  size_t off_act[DPK_MAX_LEVELS], off_gact[DPK_MAX_LEVELS], off_wlog[DPK_MAX_LEVELS], off_gsum[DPK_MAX_LEVELS];
  size_t off_rlog, off_gsum_root, total;
};

int drop_plan(const dpk_ratspn_desc* d, int64_t batch, DropPlan* p) {
  if (!d) return set_error(DPK_E_ARG, "null descriptor");
  if (d->in_features <= 0 || d->depth <= 0 || d->depth > DPK_MAX_LEVELS - 1 || d->repetitions <= 0 || d->leaf_channels <= 0 ||
      d->sum_nodes <= 0 || d->out_classes <= 0 || d->dimension <= 0 || batch < 0)
    return set_error(DPK_E_ARG, "descriptor field out of range");
  p->kind = d->leaf_kind; p->D = d->in_features; p->depth = d->depth; p->R = d->repetitions; p->K = d->leaf_channels;
  p->O = d->sum_nodes; p->C = d->out_classes; p->dim = d->dimension; p->G0 = d->repetitions << d->depth;
  p->n_sum = d->depth - 1; p->B = batch;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (off + n + 63) / 64 * 64; return o; };
  for (int l = 0; l < p->depth; ++l) {
    p->regions[l] = p->G0 >> l;
    p->ch[l] = l == 0 ? p->K : p->O;
    p->off_act[l] = take((size_t)batch * p->regions[l] * p->ch[l]);
    p->off_gact[l] = take((size_t)batch * p->regions[l] * p->ch[l]);
  }
  for (int e = 0; e < p->n_sum; ++e) {
    const size_t P = p->regions[e] / 2, kin2 = (size_t)p->ch[e] * p->ch[e];
    p->off_wlog[e] = take(P * p->O * kin2);
    p->off_gsum[e] = take(P * p->O);
  }
  const size_t kin2 = (size_t)p->ch[p->depth - 1] * p->ch[p->depth - 1];
  p->off_rlog = take((size_t)p->C * p->R * kin2);
  p->off_gsum_root = take((size_t)p->C);
  p->total = off;
  return DPK_OK;
}

// dst[row] = log_softmax(src[row]) for `rows` rows of `len` entries (one CTA per row)
__global__ void dk_log_softmax_kernel(const float* __restrict__ src, int64_t len, float* __restrict__ dst) {
  __shared__ float red[32];
  const float* row = src + (size_t)blockIdx.x * len;
  float m = -INFINITY;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) m = fmaxf(m, row[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
  m = warp_max(m);
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
  const float lse = red[0];
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) dst[(size_t)blockIdx.x * len + i] = row[i] - lse;
}

struct LeafDropArgs {
  const float* x; const int32_t* mask; const int32_t* region_len; const float* p0; const float* p1;
  int64_t B; int D, G0, K, dim, kind;
  uint64_t seed; uint32_t thr;
};

// per-dimension log-density exactly as the reference composes it (Normal.log_prob / -BCEWithLogits)
__device__ __forceinline__ float leaf_ll(int kind, float xv, float w0, float w1) {
  if (kind == DPK_LEAF_GAUSSIAN) {
    const float t = xv - w0;
    return -(t * t) / (2.f * w1 * w1) - logf(w1) - kLogSqrt2Pi;
  }
  return xv * w0 - (fmaxf(w0, 0.f) + log1pf(expf(-fabsf(w0))));
}

// thread = (b, g, k): sum over the region's kept dimensions
__global__ void dk_leaf_fwd_kernel(const LeafDropArgs a, float* __restrict__ act0) {
  const int64_t total = a.B * a.G0 * a.K;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t % a.K);
    const int g = (int)((t / a.K) % a.G0);
    const int64_t b = t / ((int64_t)a.K * a.G0);
    const int len = a.region_len[g];
    const size_t pbase = ((size_t)g * a.K + k) * a.dim;
    float s = 0.f;
    for (int d = 0; d < len; ++d) {
      if (a.thr && drop_u24(a.seed, 0u, (uint64_t)t * a.dim + d) < a.thr) continue;   // NaN -> nan_to_num -> 0
      const float xv = a.x[b * a.D + a.mask[(size_t)g * a.dim + d]];
      s += nan_to_num(leaf_ll(a.kind, xv, a.p0[pbase + d], a.p1 ? a.p1[pbase + d] : 1.f));
    }
    act0[t] = s;
  }
}

// thread = (b, g, k): gradients of sum_d keep * ll_d weighted by P = dLoss/d act0[b,g,k]
__global__ void dk_leaf_bwd_kernel(const LeafDropArgs a, const float* __restrict__ gact0, float* __restrict__ g0,
                                   float* __restrict__ g1, float* __restrict__ gx) {
  const int64_t total = a.B * a.G0 * a.K;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const float P = gact0[t];
    if (P == 0.f) continue;
    const int k = (int)(t % a.K);
    const int g = (int)((t / a.K) % a.G0);
    const int64_t b = t / ((int64_t)a.K * a.G0);
    const int len = a.region_len[g];
    const size_t pbase = ((size_t)g * a.K + k) * a.dim;
    for (int d = 0; d < len; ++d) {
      if (a.thr && drop_u24(a.seed, 0u, (uint64_t)t * a.dim + d) < a.thr) continue;
      const int f = a.mask[(size_t)g * a.dim + d];
      const float xv = a.x[b * a.D + f];
      const float w0 = a.p0[pbase + d], w1 = a.p1 ? a.p1[pbase + d] : 1.f;
      const float ll = leaf_ll(a.kind, xv, w0, w1);
      if (!(fabsf(ll) <= FLT_MAX)) continue;     // marginalised (NaN) or clamped (+-inf) by nan_to_num: no gradient
      if (a.kind == DPK_LEAF_GAUSSIAN) {
        const float inv = 1.f / w1, z = (xv - w0) * inv;
        if (g0) atomicAdd(g0 + pbase + d, P * z * inv);
        if (g1) atomicAdd(g1 + pbase + d, P * (z * z - 1.f) * inv);
        if (gx) atomicAdd(gx + b * a.D + f, -P * z * inv);
      } else {
        if (g0) atomicAdd(g0 + pbase + d, P * (xv - 1.f / (1.f + expf(-w0))));
        if (gx) atomicAdd(gx + b * a.D + f, P * w0);
      }
    }
  }
}

struct SumDropArgs {
  const float* in;      // (B, 2P, Kin)
  const float* wlog;    // (P, O, Kin^2)  | root: (C, P*Kin^2)
  float* out;           // (B, P, O)      | root: (B, C)
  int64_t B; int P, Kin, O;
  uint64_t seed; uint32_t stream, thr;
};

// thread = (b, p, o): y = logsumexp over the kept (i,j) of l_i + r_j + logw[p,o,ij]
__global__ void dk_sum_fwd_kernel(const SumDropArgs a) {
  const int64_t total = a.B * a.P * a.O;
  const int K2 = a.Kin * a.Kin;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(t % a.O);
    const int p = (int)((t / a.O) % a.P);
    const int64_t b = t / ((int64_t)a.O * a.P);
    const float* l = a.in + ((size_t)b * 2 * a.P + 2 * p) * a.Kin;
    const float* r = l + a.Kin;
    const float* w = a.wlog + ((size_t)p * a.O + o) * K2;
    const uint64_t ebase = ((uint64_t)b * a.P + p) * K2;
    float m = -INFINITY;
    for (int ij = 0; ij < K2; ++ij) {
      if (a.thr && drop_u24(a.seed, a.stream, ebase + ij) < a.thr) continue;
      m = fmaxf(m, l[ij / a.Kin] + r[ij % a.Kin] + w[ij]);
    }
    float y = m;
    if (fabsf(m) <= FLT_MAX) {
      float s = 0.f;
      for (int ij = 0; ij < K2; ++ij) {
        if (a.thr && drop_u24(a.seed, a.stream, ebase + ij) < a.thr) continue;
        s += expf(l[ij / a.Kin] + r[ij % a.Kin] + w[ij] - m);
      }
      y = m + logf(s);
    }
    a.out[t] = y;
  }
}

// thread = (b, p): top-down pass of one sum level.  gy (B,P,O) -> gin (B,2P,Kin); raw-weight gradient
// dL/draw[p,o,n] = sum_b gy * (q_n - softmax_n)  with q the posterior over the kept entries: the q part is accumulated
// here, gsum[p,o] = sum_b gy over the rows that kept anything, the softmax part is subtracted by dk_weight_finish_kernel.
__global__ void dk_sum_bwd_kernel(const SumDropArgs a, const float* __restrict__ y, const float* __restrict__ gy,
                                  float* __restrict__ gin, float* __restrict__ gw, float* __restrict__ gsum) {
  const int64_t total = a.B * a.P;
  const int K2 = a.Kin * a.Kin;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(t % a.P);
    const int64_t b = t / a.P;
    const float* l = a.in + ((size_t)b * 2 * a.P + 2 * p) * a.Kin;
    const float* r = l + a.Kin;
    float* gl = gin + ((size_t)b * 2 * a.P + 2 * p) * a.Kin;
    float* gr = gl + a.Kin;
    for (int i = 0; i < 2 * a.Kin; ++i) gl[i] = 0.f;
    const uint64_t ebase = ((uint64_t)b * a.P + p) * K2;
    for (int o = 0; o < a.O; ++o) {
      const float g = gy[((size_t)b * a.P + p) * a.O + o];
      const float yo = y[((size_t)b * a.P + p) * a.O + o];
      if (g == 0.f || !(fabsf(yo) <= FLT_MAX)) continue;     // every entry dropped (-inf): no gradient
      const float* w = a.wlog + ((size_t)p * a.O + o) * K2;
      if (gsum) atomicAdd(gsum + (size_t)p * a.O + o, g);
      for (int ij = 0; ij < K2; ++ij) {
        if (a.thr && drop_u24(a.seed, a.stream, ebase + ij) < a.thr) continue;
        const int i = ij / a.Kin, j = ij % a.Kin;
        const float q = g * expf(l[i] + r[j] + w[ij] - yo);
        gl[i] += q;
        gr[j] += q;
        if (gw) atomicAdd(gw + ((size_t)p * a.O + o) * K2 + ij, q);
      }
    }
  }
}

// root (no dropout): thread = (b, c) forward
__global__ void dk_root_fwd_kernel(const SumDropArgs a) {
  const int64_t total = a.B * a.O;
  const int K2 = a.Kin * a.Kin;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % a.O);
    const int64_t b = t / a.O;
    const float* w = a.wlog + (size_t)c * a.P * K2;
    float m = -INFINITY;
    for (int p = 0; p < a.P; ++p) {
      const float* l = a.in + ((size_t)b * 2 * a.P + 2 * p) * a.Kin;
      const float* r = l + a.Kin;
      for (int ij = 0; ij < K2; ++ij) m = fmaxf(m, l[ij / a.Kin] + r[ij % a.Kin] + w[(size_t)p * K2 + ij]);
    }
    float y = m;
    if (fabsf(m) <= FLT_MAX) {
      float s = 0.f;
      for (int p = 0; p < a.P; ++p) {
        const float* l = a.in + ((size_t)b * 2 * a.P + 2 * p) * a.Kin;
        const float* r = l + a.Kin;
        for (int ij = 0; ij < K2; ++ij) s += expf(l[ij / a.Kin] + r[ij % a.Kin] + w[(size_t)p * K2 + ij] - m);
      }
      y = m + logf(s);
    }
    a.out[t] = y;
  }
}

// root backward: thread = (b, p)
__global__ void dk_root_bwd_kernel(const SumDropArgs a, const float* __restrict__ y, const float* __restrict__ gy,
                                   float* __restrict__ gin, float* __restrict__ gw, float* __restrict__ gsum) {
  const int64_t total = a.B * a.P;
  const int K2 = a.Kin * a.Kin;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(t % a.P);
    const int64_t b = t / a.P;
    const float* l = a.in + ((size_t)b * 2 * a.P + 2 * p) * a.Kin;
    const float* r = l + a.Kin;
    float* gl = gin + ((size_t)b * 2 * a.P + 2 * p) * a.Kin;
    float* gr = gl + a.Kin;
    for (int i = 0; i < 2 * a.Kin; ++i) gl[i] = 0.f;
    for (int c = 0; c < a.O; ++c) {
      const float g = gy[(size_t)b * a.O + c];
      const float yo = y[(size_t)b * a.O + c];
      if (g == 0.f || !(fabsf(yo) <= FLT_MAX)) continue;
      const float* w = a.wlog + ((size_t)c * a.P + p) * K2;
      if (gsum && p == 0) atomicAdd(gsum + c, g);
      for (int ij = 0; ij < K2; ++ij) {
        const int i = ij / a.Kin, j = ij % a.Kin;
        const float q = g * expf(l[i] + r[j] + w[ij] - yo);
        gl[i] += q;
        gr[j] += q;
        if (gw) atomicAdd(gw + ((size_t)c * a.P + p) * K2 + ij, q);
      }
    }
  }
}

// gw[row, n] -= softmax(raw)[row, n] * gsum[row]
__global__ void dk_weight_finish_kernel(const float* __restrict__ wlog, const float* __restrict__ gsum, int64_t rows,
                                        int64_t len, float* __restrict__ gw) {
  const int64_t total = rows * len;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    gw[t] -= expf(wlog[t]) * gsum[t / len];
}

inline unsigned grid_for(int64_t n, int threads = 128) { return (unsigned)std::min<int64_t>(ceil_div(std::max<int64_t>(n, 1), threads), 1 << 16); }

uint32_t rate_thr(float rate) {
  if (!(rate > 0.f)) return 0u;
  return (uint32_t)std::min<double>((double)rate * 16777216.0, 16777215.0);
}

}  // namespace

}  // namespace dpk

using namespace dpk;

extern "C" uint32_t dpk_dropout_draw(uint64_t seed, uint32_t stream, uint64_t index) { return drop_u24(seed, stream, index); }

extern "C" size_t dpk_ratspn_dropout_workspace_bytes(const dpk_ratspn_desc* desc, int64_t batch) {
  DropPlan p;
  if (drop_plan(desc, batch, &p)) return 0;
  return std::max<size_t>(p.total, 64) * sizeof(float);
}

extern "C" int dpk_ratspn_forward_dropout(const dpk_ratspn_desc* desc, const float* x, int64_t batch,
                                          const dpk_ratspn_dropout* drop, float* out, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  DropPlan p;
  int rc = drop_plan(desc, batch, &p);
  if (rc) return rc;
  if (batch == 0) return DPK_OK;
  if (!x || !out || !drop || !desc->mask || !desc->region_len || !desc->leaf_p0 || !desc->root_weight)
    return set_error(DPK_E_ARG, "null pointer argument");
  if (!(drop->in_rate >= 0.f && drop->in_rate < 1.f) || !(drop->sum_rate >= 0.f && drop->sum_rate < 1.f))
    return set_error(DPK_E_ARG, "dropout rates must be in [0, 1)");
  if (!workspace || ((uintptr_t)workspace & 255)) return set_error(DPK_E_WORKSPACE, "workspace must be 256-byte aligned");
  if (workspace_bytes < p.total * 4) return set_error(DPK_E_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, p.total * 4);
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfScope prof(CAT_PREP, st, p.n_sum + 1);
    for (int e = 0; e < p.n_sum; ++e) {
      if (!desc->sum_weight[e]) return set_error(DPK_E_ARG, "null sum_weight[%d]", e);
      const int P = p.regions[e] / 2, kin2 = p.ch[e] * p.ch[e];
      dk_log_softmax_kernel<<<P * p.O, 128, 0, st>>>(desc->sum_weight[e], kin2, ws + p.off_wlog[e]);
    }
    const int kin = p.ch[p.depth - 1];
    dk_log_softmax_kernel<<<p.C, 256, 0, st>>>(desc->root_weight, (int64_t)p.R * kin * kin, ws + p.off_rlog);
    DPK_LAUNCH_CHECK("dk_log_softmax_kernel");
  }
  {
    ProfScope prof(CAT_LEAF, st);
    LeafDropArgs a{x, desc->mask, desc->region_len, desc->leaf_p0, desc->leaf_p1, batch, p.D, p.G0, p.K, p.dim, p.kind,
                   drop->seed, rate_thr(drop->in_rate)};
    dk_leaf_fwd_kernel<<<grid_for(batch * p.G0 * p.K), 128, 0, st>>>(a, ws + p.off_act[0]);
    DPK_LAUNCH_CHECK("dk_leaf_fwd_kernel");
  }
  for (int e = 0; e < p.n_sum; ++e) {
    ProfScope prof(CAT_EINSUM, st);
    SumDropArgs a{ws + p.off_act[e], ws + p.off_wlog[e], ws + p.off_act[e + 1], batch, p.regions[e] / 2, p.ch[e], p.O,
                  drop->seed, (uint32_t)(1 + e), rate_thr(drop->sum_rate)};
    dk_sum_fwd_kernel<<<grid_for(batch * a.P * a.O), 128, 0, st>>>(a);
    DPK_LAUNCH_CHECK("dk_sum_fwd_kernel");
  }
  {
    ProfScope prof(CAT_ROOT, st);
    const int l = p.depth - 1;
    SumDropArgs a{ws + p.off_act[l], ws + p.off_rlog, out, batch, p.R, p.ch[l], p.C, 0, 0, 0};
    dk_root_fwd_kernel<<<grid_for(batch * p.C), 128, 0, st>>>(a);
    DPK_LAUNCH_CHECK("dk_root_fwd_kernel");
  }
  return DPK_OK;
}

extern "C" int dpk_ratspn_backward_dropout(const dpk_ratspn_desc* desc, const float* x, int64_t batch,
                                           const dpk_ratspn_dropout* drop, const float* out, const float* grad_out,
                                           const dpk_ratspn_grads* grads, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  DropPlan p;
  int rc = drop_plan(desc, batch, &p);
  if (rc) return rc;
  if (batch == 0) return DPK_OK;
  if (!x || !out || !grad_out || !grads || !drop) return set_error(DPK_E_ARG, "null pointer argument");
  if (!workspace || workspace_bytes < p.total * 4) return set_error(DPK_E_WORKSPACE, "workspace too small");
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfScope prof(CAT_BWD_EINSUM, st, 2 * (p.n_sum + 1));
    DPK_CUDA_TRY(cudaMemsetAsync(ws + p.off_gsum_root, 0, (size_t)p.C * 4, st));
    for (int e = 0; e < p.n_sum; ++e) DPK_CUDA_TRY(cudaMemsetAsync(ws + p.off_gsum[e], 0, (size_t)(p.regions[e] / 2) * p.O * 4, st));
    const int l = p.depth - 1;
    const int kin2 = p.ch[l] * p.ch[l];
    SumDropArgs a{ws + p.off_act[l], ws + p.off_rlog, nullptr, batch, p.R, p.ch[l], p.C, 0, 0, 0};
    dk_root_bwd_kernel<<<grid_for(batch * p.R), 128, 0, st>>>(a, out, grad_out, ws + p.off_gact[l], grads->root_weight,
                                                               ws + p.off_gsum_root);
    if (grads->root_weight)
      dk_weight_finish_kernel<<<grid_for((int64_t)p.C * p.R * kin2, 256), 256, 0, st>>>(ws + p.off_rlog, ws + p.off_gsum_root, p.C,
                                                                                      (int64_t)p.R * kin2, grads->root_weight);
    for (int e = p.n_sum - 1; e >= 0; --e) {
      const int P = p.regions[e] / 2, k2 = p.ch[e] * p.ch[e];
      SumDropArgs s{ws + p.off_act[e], ws + p.off_wlog[e], nullptr, batch, P, p.ch[e], p.O, drop->seed, (uint32_t)(1 + e),
                    rate_thr(drop->sum_rate)};
      dk_sum_bwd_kernel<<<grid_for(batch * P), 128, 0, st>>>(s, ws + p.off_act[e + 1], ws + p.off_gact[e + 1], ws + p.off_gact[e],
                                                              grads->sum_weight[e], ws + p.off_gsum[e]);
      if (grads->sum_weight[e])
        dk_weight_finish_kernel<<<grid_for((int64_t)P * p.O * k2, 256), 256, 0, st>>>(ws + p.off_wlog[e], ws + p.off_gsum[e],
                                                                                     (int64_t)P * p.O, k2, grads->sum_weight[e]);
    }
    DPK_LAUNCH_CHECK("dropout backward (sum levels)");
  }
  {
    ProfScope prof(CAT_BWD_LEAF, st);
    LeafDropArgs a{x, desc->mask, desc->region_len, desc->leaf_p0, desc->leaf_p1, batch, p.D, p.G0, p.K, p.dim, p.kind,
                   drop->seed, rate_thr(drop->in_rate)};
    dk_leaf_bwd_kernel<<<grid_for(batch * p.G0 * p.K), 128, 0, st>>>(a, ws + p.off_gact[0], grads->leaf_p0, grads->leaf_p1,
                                                                      grads->grad_x);
    DPK_LAUNCH_CHECK("dk_leaf_bwd_kernel");
  }
  return DPK_OK;
}
