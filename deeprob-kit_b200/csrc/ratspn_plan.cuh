// ratspn_plan.cuh -- host-side launch plan / workspace layout of the RAT-SPN path.
#pragma once
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace dpk {

// Workspace layout (all offsets in floats, each 256-byte aligned):
//   leaf tables   tab   [G0][nKc][NCH][chunk]   chunk = {CHP header words | CH rows x NPK floats}; header word of
//                                              row d = byte offset of feature f in the swizzled x tile
//                                              (f*TB*4 + (f&31)*4; the reader XORs lane<<2 into it); row =
//                                              Gaussian 1/sigma[KC] | -mu/sigma[KC], Bernoulli logit[KC]
//                 cd    [G0][nKc][dim][KC]      per-dim additive constant (-log sigma - log sqrt(2pi) | -softplus)
//                 cst   [G0][Kp]                sum over the real dims of cd
//   sum level e   wsoft/wlog [P_e][nOc][Kin_e^2][OC]   softmax / log-softmax of the raw logits
//   root          rsoft/rlog [R][nCc][Kin^2][CC]
//   activations   act[l] [G_l][ch_l][Bp]  (sample-minor so that lanes = samples coalesce)
//   grad of act   gact[l] same shapes (only with DPK_F_SAVE_ACTIVATIONS)
constexpr uint32_t kPlanLeafOnly = 1u << 30;   // internal make_plan flag: stand-alone leaf layer (act[0] must be the true leaf LL)
constexpr int kTreeMaxD = 4;   // deepest region graph the fused tree kernel walks (ratspn_tree_mma.cu)
bool ratspn_tree_instantiated(int KL, int O);

struct RatPlan {
  int kind, D, depth, R, K, O, C, dim, G0;
  int fwd_kind;  // table/kernel flavour: DPK_LEAF_GAUSSIAN, DPK_LEAF_BERNOULLI or kLeafGaussUnit (scale == 1 everywhere)
  int n_sum;  // depth - 1 inner sum levels
  int64_t B, Bp;
  Chunking kc, oc, cc;
  int np;    // table values per channel on the fast path
  // leaf kernel geometry (decided here because it fixes the table layout the prep kernel writes)
  int leaf_mode;          // 4: 128-sample tile, 2: 64-sample tile, 1: 32-sample tile, 0: wide fallback (no shared tile)
  int leaf_tb;            // samples per tile
  int leaf_npk;           // floats per table row = round_up(np * KC, 4)
  int leaf_ch;            // table rows per ring chunk (<= 32)
  int leaf_nch;           // chunks per (region, channel chunk) = ceil(dim / leaf_ch)
  int leaf_chp;           // header words per chunk = round_up(leaf_ch, 4)
  int leaf_chunk_floats;  // leaf_chp + leaf_ch * leaf_npk
  int leaf_stages;        // ring depth (2..kLeafMaxStages)
  size_t leaf_smem;       // dynamic shared memory of the leaf kernel
  // tensor-core leaf (ratspn_leaf_mma.cu): 0 = not used for this call
  int leaf_mma;
  int act0_tiled;              // act[0] as [128-sample tile][G0*K columns][128] (fused tree kernel behind it), else [G0*K][Bp]
  int64_t act0_cs, act0_ts;    // element (column c, sample b) of act[0]: (b >> 7) * act0_ts + c * act0_cs + (b & 127)
  int leaf_stream;             // narrow model: streaming leaf kernel (ratspn_leaf_stream.cu); implies leaf_mma and leaf_conv
  int leaf_conv;               // main GEMM converts the fp32 inputs itself: no PREP launch, no x images, -x^2/2 added at the root
  size_t off_sqsum;            // [Bp] -1/2 sum_f x_f^2 (leaf_conv with unit-scale Gaussian leaves), else 0
  int mma_nS, mma_nW, mma_kb;  // x^2 (region indicator) N tiles, weight N tiles, 32-feature K blocks
  size_t off_wimg, off_simg, off_cstm, off_sq, off_mflags, off_aimg;
  int act_regions[DPK_MAX_LEVELS], act_ch[DPK_MAX_LEVELS];
  size_t off_tab, off_cd, off_cst;
  size_t off_wsoft[DPK_MAX_LEVELS], off_wlog[DPK_MAX_LEVELS], w_floats[DPK_MAX_LEVELS];
  int einsum_mma[DPK_MAX_LEVELS];      // level uses the tensor-core contraction (ratspn_einsum_mma.cu)
  size_t off_wmma[DPK_MAX_LEVELS];     // its weight images
  size_t off_rsoft, off_rlog, r_floats;
  size_t off_rtmp;  // [R][C][Bp] per-partition partials of the root
  // fused upper levels on the tensor cores (ratspn_tree_mma.cu): 0 = not used for this call
  int tree_mma, tree_G, tree_tcols;
  size_t off_timg;                       // [R][tree_rep_bytes] tf32 hi/lo weight images of every partition of a repetition
  uint32_t tree_rep_bytes, tree_off[kTreeMaxD + 1], tree_npad[kTreeMaxD + 1];
  size_t off_act[DPK_MAX_LEVELS], off_gact[DPK_MAX_LEVELS];
  // backward scratch (only with DPK_F_SAVE_ACTIVATIONS): posterior-count accumulators in the chunked
  // weight layouts and leaf moment accumulators in the parameter layout (G0,K,dim)
  size_t off_wstat[DPK_MAX_LEVELS], off_rstat, off_s1, off_s2, off_snan, off_s0tot;
  int stats_mma;   // leaf moments of the backward as a tensor-core GEMM (ratspn_leaf_mma.cu)
  size_t off_stats_xt, off_stats_wimg, off_stats_aimg, off_stats_s, off_stats_flags;
  size_t off_bx_aimg, off_bx_wimg, off_bx_tmp, off_bx_flags;   // d/dx as a GEMM (ratspn_run_leaf_bwd_x_mma), with stats_mma
  size_t stat_begin, stat_end;  // [stat_begin, stat_end) is zero-filled at the start of a backward
  size_t total_floats;
};

static inline size_t align64(size_t v) { return (v + 63) / 64 * 64; }
bool ratspn_einsum_mma_eligible(int Kin, int O, int nOc, int64_t Bp);          // ratspn_einsum_mma.cu
size_t ratspn_einsum_mma_image_floats(int P, int Kin, int O);
static inline bool einsum_mma_eligible(int Kin, int O, int nOc, int64_t Bp) { return ratspn_einsum_mma_eligible(Kin, O, nOc, Bp); }
static inline size_t einsum_mma_image_floats(int P, int Kin, int O) { return ratspn_einsum_mma_image_floats(P, Kin, O); }

constexpr int kLeafMaxStages = 4;  // per-warp ring depth of TMA-bulk parameter chunks (upper bound)
constexpr int kLeafWarps = 8;
constexpr int kLeafGaussUnit = 2;  // internal leaf flavour: Gaussian with scale == 1 (desc->leaf_p1 == NULL)
// tensor-core leaf tiling: 256 samples x 256 columns per unit, 32 features per K block, 3-stage ring
constexpr int kMmaTileM = 256, kMmaTileN = 256, kMmaKB = 32, kMmaStages = 3;
constexpr int64_t kMmaMinBatch = 8192;  // below this the persistent grid cannot fill the SMs

static inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Shared memory of the leaf kernel = transposed x tile (D*TB floats) + one ring of kLeafStages chunks
// per warp + mbarriers.  Picks the largest tile that leaves room for a useful ring, then the chunk
// height that wastes the fewest padded rows.
static inline void plan_leaf_geometry(RatPlan* p) {
  const size_t smem_max = (size_t)max_dynamic_smem();
  const int nsm = sm_count();
  p->leaf_npk = (p->np * p->kc.chunk + 3) / 4 * 4;
  const int want_stages = std::min(kLeafMaxStages, std::max(2, env_int("DPK_LEAF_STAGES", 2)));  // tuning knob
  const int want_ch = env_int("DPK_LEAF_CH", 0);                                                  // tuning knob
  auto fit = [&](int TB, int stages) -> int {  // best chunk height for this tile, 0 = does not fit
    const size_t fixed = (size_t)p->D * TB * 4 + kLeafWarps * kLeafMaxStages * 8 + 128;
    if (fixed >= smem_max) return 0;
    const size_t per_stage = (smem_max - fixed) / ((size_t)kLeafWarps * stages * 4);  // floats
    int chmax = 0;
    for (int ch = 1; ch <= 32 && ch <= p->dim; ++ch)
      if ((size_t)((ch + 3) / 4 * 4 + ch * p->leaf_npk) <= per_stage) chmax = ch;
    if (chmax < 2 && chmax < p->dim) return 0;
    if (want_ch > 0 && want_ch <= chmax) return want_ch;
    // fewest chunks first (each chunk costs a barrier wait + a pipeline refill, measured ~450 cycles),
    // then the fewest padded rows
    const int nch = (int)ceil_div(p->dim, chmax);
    int best = chmax;
    for (int ch = chmax; ch >= 1 && ceil_div(p->dim, ch) == nch; --ch) best = ch;
    return best;
  };
  // ring depth: 2 measured as good as 3/4 on B200 (the TMA latency hides behind one chunk of rows) and
  // leaves room for taller chunks; fall back to shallower rings only when nothing else fits
  auto choose = [&](int TB, int* stages) -> int {
    for (int st = want_stages; st >= 2; --st) {
      const int ch = fit(TB, st);
      if (ch >= 8 || ch >= p->dim || (st == 2 && ch > 0)) { *stages = st; return ch; }
    }
    return 0;
  };
  int st64 = 0, st32 = 0, st128 = 0;
  const int ch64 = choose(64, &st64), ch32 = choose(32, &st32), ch128 = choose(128, &st128);
  if (ch128 > 0 && ceil_div(p->B, 128) >= nsm && env_int("DPK_LEAF_NO128", 0) == 0) {
    p->leaf_mode = 4; p->leaf_tb = 128; p->leaf_ch = ch128; p->leaf_stages = st128;
  } else if (ch64 > 0 && ceil_div(p->B, 64) >= nsm) { p->leaf_mode = 2; p->leaf_tb = 64; p->leaf_ch = ch64; p->leaf_stages = st64; }
  else if (ch32 > 0) { p->leaf_mode = 1; p->leaf_tb = 32; p->leaf_ch = ch32; p->leaf_stages = st32; }
  else { p->leaf_mode = 0; p->leaf_tb = 32; p->leaf_ch = p->dim < 16 ? p->dim : 16; p->leaf_stages = 2; }
  p->leaf_nch = (int)ceil_div(p->dim, p->leaf_ch);
  p->leaf_chp = (p->leaf_ch + 3) / 4 * 4;
  p->leaf_chunk_floats = p->leaf_chp + p->leaf_ch * p->leaf_npk;
  p->leaf_smem = p->leaf_mode ? (size_t)p->D * p->leaf_tb * 4 +
                                    (size_t)kLeafWarps * p->leaf_stages * p->leaf_chunk_floats * 4 +
                                    kLeafWarps * kLeafMaxStages * 8
                              : 0;
}

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
  p->fwd_kind = (p->kind == DPK_LEAF_GAUSSIAN && !d->leaf_p1) ? kLeafGaussUnit : p->kind;
  p->np = (p->fwd_kind == DPK_LEAF_GAUSSIAN) ? 2 : 1;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align64(off + n); return o; };
  plan_leaf_geometry(p);
  p->off_tab = take((size_t)p->G0 * p->kc.count * p->leaf_nch * p->leaf_chunk_floats);
  p->off_cd = take((size_t)p->G0 * p->kc.count * p->dim * p->kc.chunk);
  p->off_cst = take((size_t)p->G0 * p->kc.padded);
  // Inference only: all product + sum levels and the root as one tensor-core kernel (DPK_TREE_MMA=0 disables it,
  // =1 forces it for any batch size).  The backward needs every level's activations, so it keeps the layer-wise path.
  p->tree_mma = 0; p->tree_G = 0; p->tree_tcols = 0; p->off_timg = 0; p->tree_rep_bytes = 0;
  {
    const int knob = env_int("DPK_TREE_MMA", -1);
    const int Osel = (p->depth == 1) ? p->K : p->O;
    if (knob != 0 && !(flags & (DPK_F_SAVE_ACTIVATIONS | kPlanLeafOnly)) && (batch >= kMmaMinBatch || knob == 1) &&
        p->depth <= kTreeMaxD && ratspn_tree_instantiated(p->K, Osel)) {
      uint32_t toff = 0, nmax = 0;
      for (int lvl = 0; lvl < p->depth; ++lvl) {
        const int Kin = (lvl == 0) ? p->K : p->O;
        const int Nout = (lvl == p->depth - 1) ? p->C : p->O;
        const uint32_t npad = (uint32_t)round_up((int64_t)Kin * Nout, 16);
        p->tree_off[lvl] = toff; p->tree_npad[lvl] = npad;
        toff += (uint32_t)(1 << (p->depth - 1 - lvl)) * 2u * npad * 64u;
        nmax = std::max(nmax, npad);
      }
      const int Kr = (p->depth == 1) ? p->K : p->O;
      const int tcols = nmax <= 128 ? 128 : 256;
      const int G = 512 / tcols;
      const size_t smem = 1024 + (((size_t)toff + 1023) & ~(size_t)1023) + (size_t)G * (16384 + 2 * p->K * 512) + 128;
      if (nmax <= 256 && (p->C - 1) * Kr + 16 <= tcols && smem <= (size_t)max_dynamic_smem()) {
        p->tree_mma = 1; p->tree_G = G; p->tree_tcols = tcols; p->tree_rep_bytes = toff;
      }
    }
  }
  // Tensor-core leaf (all three flavours), 16-byte loadable rows.  DPK_LEAF_MMA=0 disables it, =1 forces it for any batch size (tests).
  {
    const int knob = env_int("DPK_LEAF_MMA", -1);
    const size_t mma_smem = (size_t)kMmaStages * 4 * kMmaTileN * kMmaKB * 2 + 2 * kMmaTileN * 4 + 8 * 16 * 32 * 4 + 256 + 1024;
    // R*K columns per feature: below ~24 the level is HBM bound, not FMA bound, and the CUDA-core
    // kernel -- which reads x exactly once -- beats the GEMM with its extra pass over the operand images
    const bool wide = p->R * p->K >= env_int("DPK_LEAF_MMA_MIN_RK", 24);
    p->leaf_mma = (p->D % 4 == 0 && knob != 0 &&
                   ((batch >= kMmaMinBatch && wide) || knob == 1) && mma_smem <= (size_t)max_dynamic_smem())
                      ? 1 : 0;
    p->mma_nS = p->mma_nW = p->mma_kb = 0;
    p->off_wimg = p->off_simg = p->off_cstm = p->off_sq = p->off_mflags = p->off_aimg = p->off_sqsum = 0;
    // With the fused upper levels behind it the GEMM converts its inputs itself (DPK_LEAF_CONV=0: separate PREP launch)
    // Narrow models (all leaf columns in one N tile, HBM bound): the streaming kernel of ratspn_leaf_stream.cu, which reads
    // x once by TMA.  DPK_LEAF_STREAM=0 disables it, =1 forces it for any batch size (tests).
    p->leaf_stream = 0;
    {
      const int sknob = env_int("DPK_LEAF_STREAM", -1);
      if (sknob != 0 && knob != 0 && p->tree_mma && p->fwd_kind != DPK_LEAF_GAUSSIAN && p->D % 4 == 0 && p->G0 * p->K <= 256 &&
          ((batch >= kMmaMinBatch && !wide) || sknob == 1) && mma_smem <= (size_t)max_dynamic_smem()) {
        p->leaf_stream = 1; p->leaf_mma = 1;
      }
    }
    p->leaf_conv = (p->leaf_mma && p->tree_mma && p->fwd_kind != DPK_LEAF_GAUSSIAN && (env_int("DPK_LEAF_CONV", 1) != 0 || p->leaf_stream)) ? 1 : 0;
    if (p->leaf_mma) {
      p->mma_nS = (p->fwd_kind == kLeafGaussUnit && !p->leaf_conv) ? (int)ceil_div(p->G0, kMmaTileN) : 0;
      p->mma_nW = (int)ceil_div((int64_t)p->G0 * p->K, kMmaTileN);
      // K blocks of 32 features; a general Gaussian has a second half of them for x^2
      p->mma_kb = (int)ceil_div(p->D, kMmaKB) * (p->fwd_kind == DPK_LEAF_GAUSSIAN ? 2 : 1);
      const size_t img_floats = (size_t)kMmaTileN * kMmaKB * 2 / 4;
      p->off_wimg = take((size_t)p->mma_nW * p->mma_kb * 2 * img_floats);
      p->off_simg = take((size_t)p->mma_nS * p->mma_kb * img_floats);   // directly behind wimg (one memset)
      p->off_cstm = take((size_t)p->G0 * p->K);
      if (!p->leaf_conv) {
        p->off_sq = take((size_t)p->G0 * p->Bp);
        p->off_aimg = take((size_t)ceil_div(p->B, kMmaTileM) * p->mma_kb * 2 * img_floats);   // hi/lo fp16 split of x
      } else if (p->fwd_kind == kLeafGaussUnit) {
        p->off_sqsum = take((size_t)p->Bp);
      }
      p->off_mflags = take((size_t)p->Bp / 32 + 4 + 64);   // redo | wflag | unit counters | debug stats
    }
  }
  // With the tree kernel behind it act[0] is tile-major: a 128-sample tile's G0*K columns are one contiguous block, so the
  // tree kernel's TMA box (2K columns x 128 samples) and a leaf epilogue's stores touch one page instead of 2K resp. 256
  // pages Bp*4 bytes apart (DPK_ACT_TILED=0: column-major like the layer-wise path).
  p->act0_tiled = (p->tree_mma && env_int("DPK_ACT_TILED", 1) != 0) ? 1 : 0;
  p->act0_cs = p->act0_tiled ? 128 : p->Bp;
  p->act0_ts = p->act0_tiled ? (int64_t)p->G0 * p->K * 128 : 128;
  for (int l = 0; l < p->depth; ++l) {
    p->act_regions[l] = p->G0 >> l;
    p->act_ch[l] = (l == 0) ? p->K : p->O;
  }
  for (int e = 0; e < p->n_sum; ++e) {
    size_t kin2 = (size_t)p->act_ch[e] * p->act_ch[e];
    p->w_floats[e] = (size_t)(p->act_regions[e] / 2) * p->oc.count * kin2 * p->oc.chunk;
    p->off_wsoft[e] = take(p->w_floats[e]);
    p->off_wlog[e] = take(p->w_floats[e]);
    p->einsum_mma[e] = einsum_mma_eligible(p->act_ch[e], p->O, p->oc.count, p->Bp) ? 1 : 0;
    p->off_wmma[e] = p->einsum_mma[e] ? take(einsum_mma_image_floats(p->act_regions[e] / 2, p->act_ch[e], p->O)) : 0;
  }
  {
    size_t kin2 = (size_t)p->act_ch[p->depth - 1] * p->act_ch[p->depth - 1];
    p->r_floats = (size_t)p->R * p->cc.count * kin2 * p->cc.chunk;
    p->off_rsoft = take(p->r_floats);
    p->off_rlog = take(p->r_floats);
  }
  p->off_rtmp = take((size_t)p->R * p->C * p->Bp);
  if (p->tree_mma) p->off_timg = take((size_t)p->R * p->tree_rep_bytes / 4);
  for (int l = 0; l < p->depth; ++l)
    p->off_act[l] = take((size_t)p->act_regions[l] * p->act_ch[l] * p->Bp);
  for (int l = 0; l < p->depth; ++l)
    p->off_gact[l] = (flags & DPK_F_SAVE_ACTIVATIONS) ? take((size_t)p->act_regions[l] * p->act_ch[l] * p->Bp) : 0;
  p->stat_begin = p->stat_end = off;
  p->stats_mma = 0;
  p->off_bx_aimg = p->off_bx_wimg = p->off_bx_tmp = p->off_bx_flags = 0;
  if (flags & DPK_F_SAVE_ACTIVATIONS) {
    for (int e = 0; e < p->n_sum; ++e) p->off_wstat[e] = take(p->w_floats[e]);
    p->off_rstat = take(p->r_floats);
    const size_t leaf = (size_t)p->G0 * p->K * p->dim;
    p->off_s1 = take(leaf);
    p->off_s2 = take(leaf);
    p->off_snan = take(leaf);
    p->off_s0tot = take((size_t)p->G0 * p->K);
    p->stat_end = off;
    // leaf moments as a GEMM over the batch (large batches; DPK_STATS_MMA=0 disables, =1 forces)
    {
      const int knob = env_int("DPK_STATS_MMA", -1);
      const size_t mma_smem = (size_t)kMmaStages * 4 * kMmaTileN * kMmaKB * 2 + 2 * kMmaTileN * 4 + 8 * 16 * 32 * 4 + 256 + 1024;
      p->stats_mma = (knob != 0 && (batch >= kMmaMinBatch || knob == 1) && mma_smem <= (size_t)max_dynamic_smem()) ? 1 : 0;
      if (p->stats_mma) {
        const int quad = (p->kind == DPK_LEAF_GAUSSIAN) ? 1 : 0;
        const size_t F = (size_t)(quad ? 2 : 1) * p->D + 1, N = (size_t)p->G0 * p->K;
        const size_t KBn = (size_t)p->Bp / kMmaKB, img_floats = (size_t)kMmaTileN * kMmaKB * 2 / 4;
        p->off_stats_xt = take(F * p->Bp);
        p->off_stats_wimg = take((size_t)ceil_div(F, kMmaTileN) * KBn * 2 * img_floats);
        p->off_stats_aimg = take((size_t)ceil_div(N, kMmaTileM) * KBn * 2 * img_floats);
        p->off_stats_s = take(N * F);
        p->off_stats_flags = take((size_t)round_up(N, 128) / 32 + 8);
        // d/dx: P (batch x N) against [mu/sigma^2 | 1/sigma^2] (N x 2D)  (Bernoulli: logits, N x D)
        const size_t ncol = (size_t)(quad ? 2 : 1) * p->D;
        p->off_bx_aimg = take((size_t)ceil_div(round_up(p->B, 256), kMmaTileM) * ceil_div(N, kMmaKB) * 2 * img_floats);
        p->off_bx_wimg = take((size_t)ceil_div(ncol, kMmaTileN) * ceil_div(N, kMmaKB) * 2 * img_floats);
        p->off_bx_tmp = take((size_t)p->B * ncol);
        p->off_bx_flags = take(64);
      }
    }
  }
  p->total_floats = off;
  return DPK_OK;
}

}  // namespace dpk
