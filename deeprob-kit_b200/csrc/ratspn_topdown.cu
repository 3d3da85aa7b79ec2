// ratspn_topdown.cu -- the top-down passes of a RAT-SPN: MPE completion and ancestral sampling.
//
//   RatSpn.mpe / RatSpn.sample          deeprob/spn/models/ratspn.py:124-182
//   RootLayer.mpe / .sample             deeprob/spn/layers/ratspn.py:460-490   argmax / categorical over all (partition, i, j)
//   ProductLayer.mpe / .sample          :288-330   (group, i*K+j) -> children (2*group, i), (2*group+1, j)
//   SumLayer.mpe / .sample              :380-417   argmax_n (x[group, n] + log_softmax(W[group, o])[n]) / categorical(W)
//   RegionGraphLayer.mpe / .sample      :118-157   mode / draw of the selected leaf channel, NaN entries filled
//
// The reference walks the layers with batched index tensors (and builds some of them on the CPU, which breaks on a
// CUDA model).  Here one thread owns one sample and walks its own tree: the root picks a repetition and the channels
// (i, j) of its two top regions; every leaf region of that repetition then follows its path down -- at a region of
// level L with channel o the sum node's choice over its K^2 inputs gives the channel of the child on the path -- and
// writes its features.  The inner nodes of a path are re-evaluated per leaf (depth * K^2 terms, no per-thread stack).
// MPE reads the per-level log-likelihoods a forward with DPK_F_SAVE_ACTIVATIONS left in the workspace
// (sample-minor [regions*channels][Bp]) and the log-softmax tables next to them; ties resolve to the lowest flat
// index like torch.argmax.  Sampling needs the tables only; its draws come from the counter-based generator of
// ratspn_dropout.cu (inverse-CDF for the categorical choices, Box-Muller for Gaussian leaves).
#include "ratspn_kernels.cuh"

namespace dpk {

namespace {

struct TopDownArgs {
  const float* x;                    // (B, D) evidence, NaN = to be completed (MPE only)
  const float* out;                  // (B, C) log-likelihoods of the forward (MPE with y == NULL)
  const int32_t* y;                  // (B) class per sample or NULL
  float* filled;                     // (B, D)
  const float* act[DPK_MAX_LEVELS];  // per-level activations [regions*ch][Bp] (MPE)
  const float* wlog[DPK_MAX_LEVELS]; // [P][nOc][Kin2][OC]
  const float* rlog;                 // [R][nCc][Kin2][CC]
  const int32_t* mask; const int32_t* region_len; const float* p0; const float* p1;
  int64_t B, Bp;
  int D, depth, R, K, O, C, dim, kind;
  int OCc, nOc, CCc, nCc;
  uint64_t seed;
};

__device__ __forceinline__ float u01(uint64_t seed, uint64_t idx) {
  uint64_t z = (seed ^ (0xffull << 56)) + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((float)(uint32_t)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

// SAMPLE = false: argmax of value + log-weight;  SAMPLE = true: categorical draw from the weights alone
template <bool SAMPLE>
__global__ void ratspn_topdown_kernel(const TopDownArgs a) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int d = a.depth;
  const int kr = (d == 1) ? a.K : a.O;          // channels of the regions under the root
  const int kr2 = kr * kr;
  const uint64_t cbase = (uint64_t)b << 32;     // private counter range of this sample
  // ---- class ----
  int c = 0;
  if (a.C > 1) {
    if (a.y) c = a.y[b];
    else if (!SAMPLE) {
      float best = a.out[b * a.C];
      for (int k = 1; k < a.C; ++k)
        if (a.out[b * a.C + k] > best) { best = a.out[b * a.C + k]; c = k; }
    }
  }
  // ---- root: repetition + channels of its two top regions ----
  int rep = 0, ci = 0, cj = 0;
  {
    const float* top = SAMPLE ? nullptr : a.act[d - 1];
    float best = -INFINITY, acc = 0.f;
    const float u = SAMPLE ? u01(a.seed, cbase) : 0.f;
    bool done = false;
    for (int r = 0; r < a.R && !done; ++r) {
      const float* w = a.rlog + ((size_t)(r * a.nCc + c / a.CCc) * kr2) * a.CCc + c % a.CCc;
      for (int ij = 0; ij < kr2; ++ij) {
        const int i = ij / kr, j = ij - i * kr;
        if (SAMPLE) {
          acc += expf(w[(size_t)ij * a.CCc]);
          rep = r; ci = i; cj = j;               // the last entry catches rounding of the cumulative sum
          if (acc >= u) { done = true; break; }
        } else {
          const float v = top[((size_t)(2 * r) * kr + i) * a.Bp + b] + top[((size_t)(2 * r + 1) * kr + j) * a.Bp + b] +
                          w[(size_t)ij * a.CCc];
          if (v > best || (r == 0 && ij == 0)) { best = v; rep = r; ci = i; cj = j; }
        }
      }
    }
  }
  // ---- every leaf region of the repetition follows its path ----
  const int n_leaf = 1 << d;
  for (int n = 0; n < n_leaf; ++n) {
    int t = (n >> (d - 1)) & 1;
    int g = 2 * rep + t;                         // region index at level d-1
    int ch = t ? cj : ci;
    for (int L = d - 1; L >= 1; --L) {
      // region g of level L, channel ch = output ch of sum level L-1, partition g over regions (2g, 2g+1) of level L-1
      const int kin = (L - 1 == 0) ? a.K : a.O, k2 = kin * kin;
      const float* w = a.wlog[L - 1] + ((size_t)(g * a.nOc + ch / a.OCc) * k2) * a.OCc + ch % a.OCc;
      int bi = 0, bj = 0;
      if (SAMPLE) {
        // the draw of an inner node must not depend on which leaf asks: counter = f(level, region)
        const float u = u01(a.seed, cbase + ((uint64_t)L << 24) + (uint64_t)g + 1);
        float acc = 0.f;
        for (int ij = 0; ij < k2; ++ij) {
          acc += expf(w[(size_t)ij * a.OCc]);
          bi = ij / kin; bj = ij - bi * kin;
          if (acc >= u) break;
        }
      } else {
        const float* lo = a.act[L - 1];
        float best = -INFINITY;
        for (int ij = 0; ij < k2; ++ij) {
          const int i = ij / kin, j = ij - i * kin;
          const float v = lo[((size_t)(2 * g) * kin + i) * a.Bp + b] + lo[((size_t)(2 * g + 1) * kin + j) * a.Bp + b] +
                          w[(size_t)ij * a.OCc];
          if (v > best || ij == 0) { best = v; bi = i; bj = j; }
        }
      }
      t = (n >> (L - 1)) & 1;
      g = 2 * g + t;
      ch = t ? bj : bi;
    }
    // leaf region g (global index), channel ch
    const int len = a.region_len[g];
    const size_t pbase = ((size_t)g * a.K + ch) * a.dim;
    for (int q = 0; q < len; ++q) {
      const int f = a.mask[(size_t)g * a.dim + q];
      const float w0 = a.p0[pbase + q];
      float v;
      if (SAMPLE) {
        const uint64_t cq = cbase + (1ull << 31) + (uint64_t)f * 2;
        if (a.kind == DPK_LEAF_GAUSSIAN) {
          const float u1 = u01(a.seed, cq), u2 = u01(a.seed, cq + 1);
          v = w0 + (a.p1 ? a.p1[pbase + q] : 1.f) * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
        } else {
          v = (u01(a.seed, cq) < 1.f / (1.f + expf(-w0))) ? 1.f : 0.f;
        }
      } else {
        const float xv = a.x[b * a.D + f];
        const float mode = (a.kind == DPK_LEAF_GAUSSIAN) ? w0 : (w0 >= 0.f ? 1.f : 0.f);   // mean | [p >= 1/2]
        v = (xv != xv) ? mode : xv;
      }
      a.filled[b * a.D + f] = v;
    }
  }
}

int fill_args(const dpk_ratspn_desc* desc, const RatPlan& p, float* ws, TopDownArgs* a) {
  for (int l = 0; l < DPK_MAX_LEVELS; ++l) {
    a->act[l] = (l < p.depth) ? ws + p.off_act[l] : nullptr;
    a->wlog[l] = (l < p.n_sum) ? ws + p.off_wlog[l] : nullptr;
  }
  a->rlog = ws + p.off_rlog;
  a->mask = desc->mask; a->region_len = desc->region_len; a->p0 = desc->leaf_p0; a->p1 = desc->leaf_p1;
  a->B = p.B; a->Bp = p.Bp; a->D = p.D; a->depth = p.depth; a->R = p.R; a->K = p.K; a->O = p.O; a->C = p.C;
  a->dim = p.dim; a->kind = p.kind;
  a->OCc = p.oc.chunk; a->nOc = p.oc.count; a->CCc = p.cc.chunk; a->nCc = p.cc.count;
  return DPK_OK;
}

}  // namespace

}  // namespace dpk

using namespace dpk;

extern "C" int dpk_ratspn_mpe(const dpk_ratspn_desc* desc, const float* x, int64_t batch, const float* out, const int32_t* y,
                              float* filled, void* workspace, size_t workspace_bytes, void* stream) {
  RatPlan p;
  int rc = make_plan(desc, batch, DPK_F_SAVE_ACTIVATIONS, &p);
  if (rc) return rc;
  if (batch == 0) return DPK_OK;
  if (!x || !filled || !workspace || (!out && !y && p.C > 1)) return set_error(DPK_E_ARG, "null pointer argument");
  if ((rc = ratspn_check_ws(p, workspace, workspace_bytes))) return rc;
  TopDownArgs a;
  fill_args(desc, p, static_cast<float*>(workspace), &a);
  a.x = x; a.out = out; a.y = y; a.filled = filled; a.seed = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_LAYER, st);
  ratspn_topdown_kernel<false><<<(unsigned)ceil_div(batch, 128), 128, 0, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_topdown_kernel<mpe>");
  return DPK_OK;
}

extern "C" int dpk_ratspn_sample(const dpk_ratspn_desc* desc, int64_t n_samples, const int32_t* y, uint64_t seed,
                                 float* samples, void* workspace, size_t workspace_bytes, void* stream) {
  RatPlan p;
  int rc = make_plan(desc, n_samples, DPK_F_SAVE_ACTIVATIONS, &p);
  if (rc) return rc;
  if (n_samples == 0) return DPK_OK;
  if (!samples || !workspace || !desc->leaf_p0 || !desc->root_weight) return set_error(DPK_E_ARG, "null pointer argument");
  if (p.depth > 12) return set_error(DPK_E_ARG, "sampling supports region graphs up to depth 12");
  if ((rc = ratspn_check_ws(p, workspace, workspace_bytes))) return rc;
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfScope prof(CAT_PREP, st, 1 + p.n_sum);
    if ((rc = ratspn_run_prep_weights(desc, p, ws, st))) return rc;
  }
  TopDownArgs a;
  fill_args(desc, p, ws, &a);
  a.x = nullptr; a.out = nullptr; a.y = y; a.filled = samples; a.seed = seed;
  ProfScope prof(CAT_LAYER, st);
  ratspn_topdown_kernel<true><<<(unsigned)ceil_div(n_samples, 128), 128, 0, st>>>(a);
  DPK_LAUNCH_CHECK("ratspn_topdown_kernel<sample>");
  return DPK_OK;
}
