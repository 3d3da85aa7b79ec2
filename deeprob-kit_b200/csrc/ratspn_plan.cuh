// ratspn_plan.cuh -- host-side launch plan / workspace layout of the RAT-SPN path.
#pragma once
#include "common.cuh"

namespace dpk {

// Workspace layout (all offsets in floats, each 256-byte aligned):
//   leaf tables   tab   [G0][nKc][dim][ROWP]    row = {feature index (int bits), 0, 0, 0, NP*KC values}
//                                              (Gaussian NP=2: 1/sigma[KC] | -mu/sigma[KC]; Bernoulli NP=1: logit[KC])
//                 cd    [G0][nKc][dim][KC]      per-dim additive constant (-log sigma - log sqrt(2pi) | -softplus)
//                 cst   [G0][Kp]                sum over the real dims of cd
//   sum level e   wsoft/wlog [P_e][nOc][Kin_e^2][OC]   softmax / log-softmax of the raw logits
//   root          rsoft/rlog [R][nCc][Kin^2][CC]
//   activations   act[l] [G_l][ch_l][Bp]  (sample-minor so that lanes = samples coalesce)
//   grad of act   gact[l] same shapes (only with DPK_F_SAVE_ACTIVATIONS)
struct RatPlan {
  int kind, D, depth, R, K, O, C, dim, G0;
  int n_sum;  // depth - 1 inner sum levels
  int64_t B, Bp;
  Chunking kc, oc, cc;
  int np;    // table values per channel on the fast path
  int rowp;  // floats per table row: 4 (feature index + pad) + round_up(np * KC, 4)
  int act_regions[DPK_MAX_LEVELS], act_ch[DPK_MAX_LEVELS];
  size_t off_tab, off_cd, off_cst;
  size_t off_wsoft[DPK_MAX_LEVELS], off_wlog[DPK_MAX_LEVELS], w_floats[DPK_MAX_LEVELS];
  size_t off_rsoft, off_rlog, r_floats;
  size_t off_act[DPK_MAX_LEVELS], off_gact[DPK_MAX_LEVELS];
  // backward scratch (only with DPK_F_SAVE_ACTIVATIONS): posterior-count accumulators in the chunked
  // weight layouts and leaf moment accumulators in the parameter layout (G0,K,dim)
  size_t off_wstat[DPK_MAX_LEVELS], off_rstat, off_s1, off_s2, off_snan, off_s0tot;
  size_t stat_begin, stat_end;  // [stat_begin, stat_end) is zero-filled at the start of a backward
  size_t total_floats;
};

static inline size_t align64(size_t v) { return (v + 63) / 64 * 64; }

// returns 0 or a negative error (message set)
static inline int make_plan(const dpk_ratspn_desc* d, int64_t batch, uint32_t flags, RatPlan* p) {
  if (!d) return set_error(DPK_E_ARG, "null descriptor");
  if (d->leaf_kind != DPK_LEAF_GAUSSIAN && d->leaf_kind != DPK_LEAF_BERNOULLI)
    return set_error(DPK_E_ARG, "unknown leaf kind %d", d->leaf_kind);
  if (d->in_features <= 0 || d->depth <= 0 || d->depth > DPK_MAX_LEVELS - 1 || d->repetitions <= 0 ||
      d->leaf_channels <= 0 || d->sum_nodes <= 0 || d->out_classes <= 0 || d->dimension <= 0 || batch < 0)
    return set_error(DPK_E_ARG, "descriptor field out of range");
  if ((int64_t)d->dimension << d->depth < d->in_features)
    return set_error(DPK_E_ARG, "dimension * 2^depth < in_features");
  p->kind = d->leaf_kind; p->D = d->in_features; p->depth = d->depth; p->R = d->repetitions;
  p->K = d->leaf_channels; p->O = d->sum_nodes; p->C = d->out_classes; p->dim = d->dimension;
  p->G0 = d->repetitions << d->depth;
  p->n_sum = d->depth - 1;
  p->B = batch; p->Bp = round_up(batch > 0 ? batch : 1, 128);
  p->kc = pick_chunk(p->K); p->oc = pick_chunk(p->O); p->cc = pick_chunk(p->C);
  p->np = (p->kind == DPK_LEAF_GAUSSIAN) ? 2 : 1;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align64(off + n); return o; };
  p->rowp = 4 + (p->np * p->kc.chunk + 3) / 4 * 4;
  p->off_tab = take((size_t)p->G0 * p->kc.count * p->dim * p->rowp);
  p->off_cd = take((size_t)p->G0 * p->kc.count * p->dim * p->kc.chunk);
  p->off_cst = take((size_t)p->G0 * p->kc.padded);
  for (int l = 0; l < p->depth; ++l) {
    p->act_regions[l] = p->G0 >> l;
    p->act_ch[l] = (l == 0) ? p->K : p->O;
  }
  for (int e = 0; e < p->n_sum; ++e) {
    size_t kin2 = (size_t)p->act_ch[e] * p->act_ch[e];
    p->w_floats[e] = (size_t)(p->act_regions[e] / 2) * p->oc.count * kin2 * p->oc.chunk;
    p->off_wsoft[e] = take(p->w_floats[e]);
    p->off_wlog[e] = take(p->w_floats[e]);
  }
  {
    size_t kin2 = (size_t)p->act_ch[p->depth - 1] * p->act_ch[p->depth - 1];
    p->r_floats = (size_t)p->R * p->cc.count * kin2 * p->cc.chunk;
    p->off_rsoft = take(p->r_floats);
    p->off_rlog = take(p->r_floats);
  }
  for (int l = 0; l < p->depth; ++l)
    p->off_act[l] = take((size_t)p->act_regions[l] * p->act_ch[l] * p->Bp);
  for (int l = 0; l < p->depth; ++l)
    p->off_gact[l] = (flags & DPK_F_SAVE_ACTIVATIONS) ? take((size_t)p->act_regions[l] * p->act_ch[l] * p->Bp) : 0;
  p->stat_begin = p->stat_end = off;
  if (flags & DPK_F_SAVE_ACTIVATIONS) {
    for (int e = 0; e < p->n_sum; ++e) p->off_wstat[e] = take(p->w_floats[e]);
    p->off_rstat = take(p->r_floats);
    const size_t leaf = (size_t)p->G0 * p->K * p->dim;
    p->off_s1 = take(leaf);
    p->off_s2 = take(leaf);
    p->off_snan = take(leaf);
    p->off_s0tot = take((size_t)p->G0 * p->K);
    p->stat_end = off;
  }
  p->total_floats = off;
  return DPK_OK;
}

}  // namespace dpk
