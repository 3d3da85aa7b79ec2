// ratspn_tree_mma.cu -- every product + sum level of a RAT-SPN and its root in ONE kernel, mixtures on tcgen05.
//
//   ProductLayer.forward  deeprob/spn/layers/ratspn.py:272-286   x1[..., :, None] + x2[..., None, :]
//   SumLayer.forward      deeprob/spn/layers/ratspn.py:363-378   logsumexp(x + log_softmax(W))
//   RootLayer.forward     deeprob/spn/layers/ratspn.py:446-458   logsumexp over all partitions x nodes
//
// A repetition of the region graph is a complete binary tree: 2^d leaf regions, 2^(d-1-e) partitions at sum level e,
// one root product.  A thread owns one sample (= one TMEM lane) and walks the tree of one repetition in post-order;
// a region's K values never leave the registers between levels.  They are carried in the scaled linear form
//      region = (m, v[K]),   log-value_k = m + log v_k,   max_k v_k = 1
// so that a product+sum node is
//      s_o = sum_i v_left[i] * ( sum_j W[o,i,j] * v_right[j] ),     m' = m_left + m_right + log max_o s_o,   v' = s / max s
// with NO exp / log between levels (the leaves pay K exps per region, every inner node one log and one reciprocal).
// The inner sum -- K of the K+1 multiply-adds per output -- is the GEMM
//      T[b, (o,i)] = sum_j v_right[b, j] * W_p[(o,i), j]      M = 128 samples, N = O*K_in, K = K_in (padded to 8 / 16)
// issued as tcgen05.mma kind::tf32 with the accumulator in TMEM; the CUDA cores keep the finish sum_i v_left[i] T[o,i]
// (packed FFMA2 over pairs of i).  fp32 accuracy: both operands are split v = hi + lo with hi = the tf32 truncation of v
// (lo is exact in fp32 and is itself truncated to tf32 by the tensor core: 22 significant bits, full fp32 exponent
// range) and the product is taken in three passes hi*hi + lo*hi + hi*lo with fp32 accumulation.
//
// Kernel shape: CTA = (repetition r, batch chunk); the weight images of all partitions of the repetition are loaded
// once by one TMA bulk copy (UBLKCP) and stay in shared memory; G = 512 / Tcols warpgroups work on independent sample
// tiles, each with its own A staging buffer and its own TMEM accumulator T, so the stage -> MMA -> drain chain of one
// warpgroup overlaps the others'.  The MMAs of a warpgroup are issued by one elected thread of its first warp (no
// dedicated MMA warp: tcgen05.mma is a single-thread instruction).  Inputs: the sample-minor leaf activations
// act[0] = [G0*K][Bp] (lanes = samples: coalesced).  Output: per-repetition root partials [R][C][Bp], combined by
// ratspn_root_combine_kernel.
// A sample whose linear-domain sums leave [1e-30, FLT_MAX] anywhere in the tree (weights below e^-69, -inf / +inf
// / NaN activations) is recomputed for that repetition by the exact log-domain path (torch.logsumexp semantics).
#include <algorithm>

#include "ratspn_kernels.cuh"
#include "tc_common.cuh"

namespace dpk {

namespace {

using namespace tc;

constexpr int kTile = 128;                       // samples per tile = TMEM lanes
constexpr uint32_t kAHalf = kTile * 64;          // one A image: 128 rows x 64 B
constexpr uint32_t kABytes = 2 * kAHalf;         // hi | lo
constexpr float kTinySumT = 1e-30f;

struct TreeArgs {
  const float* act0;            // element (column c, sample b) at (b >> 7) * act_ts + c * act_cs + (b & 127)  (RatPlan::act0_*)
  int64_t act_cs, act_ts;
  int tile_rows, c0_step;       // tile-major act[0]: (G0*KL, 0); column-major: (0, 128)
  int pdl;                      // host only: stage this launch behind its predecessor (common.cuh)
  const unsigned char* wimg;    // [R][rep_bytes]
  float* part;                  // [R][C][Bp]
  const float* wlog[kTreeMaxD]; // log-softmax tables of the sum levels (exact path), [P][nOc][Kin2][OC]
  const float* rlog;            // [R][nCc][Kin2][CC]
  int64_t B, Bp;
  int depth, R, C, KL, O;
  int OCc, nOc, CCc, nCc;
  int n_tiles, n_chunks, G, Tcols;
  uint32_t rep_bytes;
  uint32_t lvl_off[kTreeMaxD + 1];
  uint32_t lvl_npad[kTreeMaxD + 1];
};

// exact log-domain value of one output: logsumexp_ij(l_i + r_j + logw[ij]) with torch's non-finite propagation
__device__ __noinline__ float tree_exact_lse(const float* __restrict__ l, const float* __restrict__ r, int64_t stride, int Kin,
                                             const float* __restrict__ wlog, int OC) {
  float m = -INFINITY;
  for (int i = 0; i < Kin; ++i)
    for (int j = 0; j < Kin; ++j) m = fmaxf(m, l[i * stride] + r[j * stride] + __ldg(wlog + (size_t)(i * Kin + j) * OC));
  bool nan = false;
  for (int i = 0; i < Kin; ++i) nan |= (l[i * stride] != l[i * stride]) || (r[i * stride] != r[i * stride]);
  if (nan) return NAN;
  if (!(fabsf(m) <= FLT_MAX)) return m;
  float s = 0.f;
  for (int i = 0; i < Kin; ++i)
    for (int j = 0; j < Kin; ++j) s += expf(l[i * stride] + r[j * stride] + __ldg(wlog + (size_t)(i * Kin + j) * OC) - m);
  return m + logf(s);
}

// one (sample, repetition) entirely in the log domain, like the layer-wise kernels of ratspn_einsum.cu
__device__ __noinline__ void tree_exact_rep(const TreeArgs& a, int r, int64_t b) {
  const int d = a.depth, KL = a.KL, O = a.O;
  float lvl[kTreeMaxD][16];
  float cur[16], tmp[16];
  const int n0 = 1 << (d - 1);
  for (int n = 0; n < n0; ++n) {
    const float* lp = a.act0 + (b >> 7) * a.act_ts + (size_t)(((r << d) + 2 * n) * KL) * a.act_cs + (b & 127);
    const float* rp = lp + (size_t)KL * a.act_cs;
    if (d == 1) {
      for (int c = 0; c < a.C; ++c)
        a.part[((size_t)r * a.C + c) * a.Bp + b] =
            tree_exact_lse(lp, rp, a.act_cs, KL, a.rlog + ((size_t)(r * a.nCc + c / a.CCc) * KL * KL) * a.CCc + c % a.CCc, a.CCc);
      return;
    }
    {
      const int p = r * n0 + n;
      for (int o = 0; o < O; ++o)
        cur[o] = tree_exact_lse(lp, rp, a.act_cs, KL, a.wlog[0] + ((size_t)(p * a.nOc + o / a.OCc) * KL * KL) * a.OCc + o % a.OCc, a.OCc);
    }
    for (int L = 1; L < d; ++L) {
      const int idx = n >> (L - 1);
      if (!(idx & 1)) {
        for (int o = 0; o < O; ++o) lvl[L][o] = cur[o];
        break;
      }
      if (L == d - 1) {
        for (int c = 0; c < a.C; ++c)
          a.part[((size_t)r * a.C + c) * a.Bp + b] =
              tree_exact_lse(lvl[L], cur, 1, O, a.rlog + ((size_t)(r * a.nCc + c / a.CCc) * O * O) * a.CCc + c % a.CCc, a.CCc);
        break;
      }
      const int p = r * (n0 >> L) + (idx >> 1);
      for (int o = 0; o < O; ++o)
        tmp[o] = tree_exact_lse(lvl[L], cur, 1, O, a.wlog[L] + ((size_t)(p * a.nOc + o / a.OCc) * O * O) * a.OCc + o % a.OCc, a.OCc);
      for (int o = 0; o < O; ++o) cur[o] = tmp[o];
    }
  }
}

// per-thread view of its warpgroup's pipeline resources
struct WgCtx {
  unsigned char* a_ptr;   // A staging buffer (generic pointer), this thread's row = tid
  uint32_t a_sm;          // its shared-memory address
  uint32_t w_sm;          // weight block of the repetition
  uint32_t t_addr;        // TMEM address of this thread's lane, first column of the warpgroup's accumulator
  uint32_t d_addr;        // TMEM address of the accumulator (lane 0) for the MMA
  uint64_t* dfull;
  uint32_t phase;
  const float* l_ptr;     // leaf staging buffer of the warpgroup: [2*KL rows][128 samples], filled by TMA bulk copies
  uint32_t l_sm;
  uint64_t* lfull;
  uint32_t lphase;
  int wg, tid, warp_in_wg;
};

// One TMA tensor copy (box = 2*KL activation rows x the tile's 128 samples) brings a pair of sibling leaf regions
// into the warpgroup's staging buffer.  No registers are involved, so the copy stays in flight across a whole
// product+sum node (a register prefetch was spilled by the compiler, i.e. waited for immediately).
template <int KL>
__device__ __forceinline__ void tree_issue_pair(const WgCtx& c, const CUtensorMap* tmap, int sample0, int row0) {
  if ((c.tid & 31) == 0) {
    mbar_expect_tx(c.lfull, 2u * KL * 512u);
    tma_load_2d(c.l_sm, tmap, sample0, row0, c.lfull);
  }
  __syncwarp();
}

// The three passes hi*hi + lo*hi + hi*lo are ONE contraction over the concatenated K axis [hi | lo | hi] x [Whi | Whi | Wlo]
// when 3*KIN tf32 values fit one 128-byte row: ceil(3*KIN / 8) MMAs instead of 3 * ceil(KIN / 8), on 128B-swizzled
// rows.  Used for KIN <= 8; at KIN = 10 (4 MMAs instead of 6: tensor pipe 128k -> 85k cycles per SM at config 2) the
// 30-word row image costs more registers than the 128-register budget of a 512-thread CTA has left -- the level
// slots spill and the kernel gets slower (0.165 -> 0.221 ms, profiles/ r2 notes) -- so wider inputs keep separate
// hi / lo images of 64-byte rows.

// stage v_right (hi/lo tf32 rows), then T = A * W^T on the tensor core; returns when T is complete
template <int KIN, int KPF = 0>
__device__ __forceinline__ void tree_mma(WgCtx& c, const TreeArgs& a, const float (&er)[KIN], int lvl, int part,
                                         const CUtensorMap* tmap = nullptr, int pf_sample = -1, int pf_row = 0) {
  constexpr bool PACK = 3 * KIN <= 24;
  if constexpr (PACK) {
    // word w of the row: hi_j (w = j), lo_j (w = KIN + j), hi_j again (w = 2 KIN + j), 0 beyond; built chunk by chunk
    // (4 words live at a time: a 32-register row image spilled the level slots out of the register file)
    auto word = [&](int w) -> uint32_t {
      if (w >= 3 * KIN) return 0u;
      const int j = w % KIN;
      const uint32_t h = __float_as_uint(er[j]) & 0xffffe000u;
      return (w / KIN == 1) ? __float_as_uint(er[j] - __uint_as_float(h)) : h;
    };
    constexpr int NCH = (3 * KIN + 3) / 4;     // 16-byte chunks that carry data; the K steps read 8 values = 2 chunks each
    constexpr int NCHW = (NCH + 1) / 2 * 2;
#pragma unroll
    for (int q = 0; q < NCHW; ++q) {
      *reinterpret_cast<uint4*>(c.a_ptr + sw128_off((uint32_t)c.tid, (uint32_t)q)) =
          make_uint4(word(4 * q), word(4 * q + 1), word(4 * q + 2), word(4 * q + 3));
      asm volatile("" ::: "memory");
    }
  } else {
    constexpr int NK = (KIN <= 8) ? 8 : 16;
    uint32_t hi[NK], lo[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      if (j < KIN) {
        const uint32_t h = __float_as_uint(er[j < KIN ? j : 0]) & 0xffffe000u;
        hi[j] = h;
        lo[j] = __float_as_uint(er[j < KIN ? j : 0] - __uint_as_float(h));
      } else {
        hi[j] = 0u; lo[j] = 0u;
      }
    }
#pragma unroll
    for (int q = 0; q < NK / 4; ++q) {
      const uint32_t off = sw64_off((uint32_t)c.tid, (uint32_t)q);
      *reinterpret_cast<uint4*>(c.a_ptr + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
      *reinterpret_cast<uint4*>(c.a_ptr + kAHalf + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
    }
  }
  fence_async_smem();   // generic-proxy stores -> visible to the tensor core
  fence_before();       // this thread's TMEM loads of the previous accumulator are ordered before the barrier
  named_bar_sync(1 + c.wg, kTile);
  if (c.warp_in_wg == 0) {
    fence_after();
    const uint32_t npad = a.lvl_npad[lvl];
    const uint32_t b_sm = c.w_sm + a.lvl_off[lvl] + (uint32_t)part * (2u * npad * 64u);
    const uint32_t idesc = idesc_m128(npad, 2u);
    const uint32_t a_hi = desc_lo(c.a_sm), b_hi = desc_lo(b_sm);
    if (elect_one()) {
      if constexpr (PACK) {
        constexpr int NKS = (3 * KIN + 7) / 8;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) mma_tf32_sw128(c.d_addr, a_hi + 2 * ks, b_hi + 2 * ks, idesc, ks > 0 ? 1u : 0u);
      } else {
        constexpr int NKS = ((KIN <= 8) ? 8 : 16) / 8;   // K steps of 8 tf32 = 32 bytes of every 64-byte row
        const uint32_t a_lo = a_hi + (kAHalf >> 4), b_lo = b_hi + ((npad * 64u) >> 4);
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) mma_tf32(c.d_addr, a_hi + 2 * ks, b_hi + 2 * ks, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) mma_tf32(c.d_addr, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) mma_tf32(c.d_addr, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
      }
      commit(c.dfull);
    }
    __syncwarp();
  }
  // every thread of the warpgroup has consumed the staged leaf values (it is past the barrier): refill (second
  // warp, so that the MMA-issuing warp stays off that path)
  if constexpr (KPF > 0) {
    if (c.warp_in_wg == 1 && pf_sample >= 0) tree_issue_pair<KPF>(c, tmap, pf_sample, pf_row);
  }
  mbar_wait(c.dfull, c.phase);
  c.phase ^= 1u;
  fence_after();
}

// s[o] = sum_i el[i] * T[o*KIN + i]   (T = this thread's TMEM lane), packed FFMA2 over pairs of i
template <int KIN, int NOUT>
__device__ __forceinline__ void tree_finish(const WgCtx& c, const float (&el)[KIN], float (&s)[NOUT]) {
  static_assert(KIN % 2 == 0, "pairs of inputs");
  constexpr int N = KIN * NOUT;
  float2 acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t t[32];
    tmem_ld32(c.t_addr + (uint32_t)c0, t);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int col = c0 + 2 * q;
      if (col < N) {
        const int o = col / KIN, i = col % KIN;
        acc[o] = __ffma2_rn(make_float2(el[i], el[i + 1]), make_float2(__uint_as_float(t[2 * q]), __uint_as_float(t[2 * q + 1])),
                            acc[o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) s[o] = acc[o].x + acc[o].y;
}

// linear sums relative to `shift` -> (m, v) with max v = 1
template <int NOUT>
__device__ __forceinline__ void tree_normalize(const float (&s)[NOUT], float shift, float* m, float (&v)[NOUT], bool* bad) {
  float mx = s[0];
  bool fin = s[0] <= FLT_MAX;
#pragma unroll
  for (int o = 1; o < NOUT; ++o) { mx = fmaxf(mx, s[o]); fin = fin && (s[o] <= FLT_MAX); }
  if (!(fin && mx >= kTinySumT)) { *bad = true; mx = 1.f; }
  const float inv = __frcp_rn(mx);
#pragma unroll
  for (int o = 0; o < NOUT; ++o) v[o] = s[o] * inv;
  *m = shift + log_fast(mx);
}

// root product of one repetition: per class c  part = shift + log sum_i el[i] T[c*KIN + i]
template <int KIN, int KPF = 0>
__device__ __forceinline__ void tree_root(WgCtx& c, const TreeArgs& a, const float (&el)[KIN], const float (&er)[KIN], float shift,
                                          int r, int64_t b, bool* bad, const CUtensorMap* tmap = nullptr, int pf_sample = -1,
                                          int pf_row = 0) {
  tree_mma<KIN, KPF>(c, a, er, a.depth - 1, 0, tmap, pf_sample, pf_row);
#pragma unroll 1
  for (int cls = 0; cls < a.C; ++cls) {
    uint32_t t[16];
    tmem_ld16(c.t_addr + (uint32_t)(cls * KIN), t);
    tmem_ld_wait();
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < KIN; i += 2) {
      s0 = fmaf(el[i], __uint_as_float(t[i]), s0);
      s1 = fmaf(el[i + 1], __uint_as_float(t[i + 1]), s1);
    }
    const float s = s0 + s1;
    if (!(s >= kTinySumT && s <= FLT_MAX)) *bad = true;
    a.part[((size_t)r * a.C + cls) * a.Bp + b] = shift + log_fast(s);
  }
}

// MAXD: deepest tree this instantiation walks (MAXD - 1 slots of waiting left children live in registers);
// two warpgroups of 256 accumulator columns when a level has more than 128 (o, i) pairs, else four of 128.
template <int KL, int O, int MAXD>
__global__ void __launch_bounds__((KL * O > 128 || O * O > 128) ? 256 : 512, 1) ratspn_tree_mma_kernel(const TreeArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;      // swizzled images: 1024-byte aligned
  unsigned char* sm = smem_raw + (base - raw);
  const uint32_t w_bytes_al = (a.rep_bytes + 1023u) & ~1023u;
  const uint32_t l_bytes = 2u * KL * 512u;
  const uint32_t l_off = w_bytes_al + (uint32_t)a.G * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + l_off + (uint32_t)a.G * l_bytes);
  uint64_t* wfull = bars;            // weight block landed
  uint64_t* dfull = bars + 1;        // [G] accumulator of the warpgroup complete
  uint64_t* lfull = bars + 1 + 4;    // [G] staged leaf values of the warpgroup landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 8);
  const int warp = threadIdx.x >> 5;
  const int r = blockIdx.x;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int g = 0; g < a.G; ++g) { mbar_init(dfull + g, 1); mbar_init(lfull + g, 1); }
    mbar_init_fence();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(wfull, a.rep_bytes);
    bulk_g2s(base, a.wimg + (size_t)r * a.rep_bytes, a.rep_bytes, wfull);
  }

  WgCtx c;
  c.wg = threadIdx.x >> 7; c.tid = threadIdx.x & 127; c.warp_in_wg = c.tid >> 5;
  c.a_ptr = sm + w_bytes_al + (uint32_t)c.wg * kABytes;
  c.a_sm = base + w_bytes_al + (uint32_t)c.wg * kABytes;
  c.w_sm = base;
  c.d_addr = tmem + (uint32_t)(c.wg * a.Tcols);
  c.t_addr = c.d_addr + ((uint32_t)(c.warp_in_wg * 32) << 16);
  c.dfull = dfull + c.wg;
  c.phase = 0u;
  c.l_ptr = reinterpret_cast<const float*>(sm + l_off + (uint32_t)c.wg * l_bytes);
  c.l_sm = base + l_off + (uint32_t)c.wg * l_bytes;
  c.lfull = lfull + c.wg;
  c.lphase = 0u;

  const int d = a.depth;
  const int n0 = 1 << (d - 1);
  const int t_first = blockIdx.y * a.G + c.wg, t_step = a.n_chunks * a.G;
  // rows of this repetition's leaf regions: region g, channel k -> row (g*KL + k), sample-minor
  const int row_base = (r << d) * KL;
  // everything above (barriers, TMEM, the weight images, which an earlier launch wrote) may overlap the predecessor's tail;
  // act[0], the flags and sqsum are its output
  pdl_wait();
  if (c.warp_in_wg == 1 && t_first < a.n_tiles) tree_issue_pair<KL>(c, &tmap, t_first * a.c0_step, t_first * a.tile_rows + row_base);
  mbar_wait(wfull, 0u);

#pragma unroll 1
  for (int t = t_first; t < a.n_tiles; t += t_step) {
    const int64_t b = (int64_t)t * kTile + c.tid;
    bool bad = false;
    float slot_m[MAXD - 1];
    float slot_v[MAXD - 1][O];
#pragma unroll 1
    for (int n = 0; n < n0; ++n) {
      // ---- leaf regions 2n, 2n+1 -> scaled linear form ----
      float el[KL], er[KL];
      mbar_wait(c.lfull, c.lphase);
      c.lphase ^= 1u;
#pragma unroll
      for (int k = 0; k < KL; ++k) { el[k] = c.l_ptr[k * kTile + c.tid]; er[k] = c.l_ptr[(KL + k) * kTile + c.tid]; }
      float ml = el[0], mr = er[0];
#pragma unroll
      for (int k = 1; k < KL; ++k) { ml = fmaxf(ml, el[k]); mr = fmaxf(mr, er[k]); }
      if (!(fabsf(ml) <= FLT_MAX) || !(fabsf(mr) <= FLT_MAX)) { bad = true; ml = 0.f; mr = 0.f; }
#pragma unroll
      for (int k = 0; k < KL; ++k) { el[k] = exp_fast(el[k] - ml); er[k] = exp_fast(er[k] - mr); }
      // the next pair of leaf regions (of this tile, else the first pair of the warpgroup's next tile)
      // TMA coordinates of the next pair: (sample, column) of the column-major act[0], or (0, tile * G0*K + column) of the
      // tile-major one (c0_step = 128 resp. 0, tile_rows = 0 resp. G0*K)
      int pf_tile = -1, pf_row = row_base;
      if (n + 1 < n0) { pf_tile = t; pf_row = row_base + 2 * (n + 1) * KL; }
      else if (t + t_step < a.n_tiles) pf_tile = t + t_step;
      const int pf_sample = pf_tile < 0 ? -1 : pf_tile * a.c0_step;
      pf_row += max(pf_tile, 0) * a.tile_rows;
      if (d == 1) {
        tree_root<KL, KL>(c, a, el, er, ml + mr, r, b, &bad, &tmap, pf_sample, pf_row);
        break;
      }
      float cur_m, cur_v[O];
      {
        float s[O];
        tree_mma<KL, KL>(c, a, er, 0, n, &tmap, pf_sample, pf_row);
        tree_finish<KL, O>(c, el, s);
        tree_normalize<O>(s, ml + mr, &cur_m, cur_v, &bad);
      }
      // ---- carry upwards: a left child waits in its level's slot, a right child is combined with it ----
      bool active = true;
#pragma unroll
      for (int L = 1; L < MAXD; ++L) {
        if (active && L <= d - 1) {
          const int idx = n >> (L - 1);
          if ((idx & 1) == 0) {
            slot_m[L - 1] = cur_m;
#pragma unroll
            for (int o = 0; o < O; ++o) slot_v[L - 1][o] = cur_v[o];
            active = false;
          } else if (L == d - 1) {
            tree_root<O>(c, a, slot_v[L - 1], cur_v, slot_m[L - 1] + cur_m, r, b, &bad);
            active = false;
          } else {
            float s[O];
            tree_mma<O>(c, a, cur_v, L, idx >> 1);
            tree_finish<O, O>(c, slot_v[L - 1], s);
            tree_normalize<O>(s, slot_m[L - 1] + cur_m, &cur_m, cur_v, &bad);
          }
        }
      }
    }
    if (bad && b < a.B) tree_exact_rep(a, r, b);
  }

  fence_before();
  __syncthreads();
  if (warp == 0) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// weight images of one level: row n = o*Kin + i, column j: softmax weight W[p, o, i*Kin + j] as tf32 hi / lo
__global__ void ratspn_prep_tree_kernel(const float* __restrict__ wsoft, int P_total, int parts, int Kin, int Nout, int OC,
                                        int nOc, int npad, uint32_t lvl_off, uint32_t rep_bytes, int pack,
                                        unsigned char* __restrict__ wimg) {
  const int64_t total = (int64_t)P_total * npad * 16;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 15);
    const int n = (int)((idx >> 4) % npad);
    const int pg = (int)(idx / ((int64_t)npad * 16));
    const int r = pg / parts, q = pg - r * parts;
    const int o = n / Kin, i = n - o * Kin;
    float w = 0.f;
    if (o < Nout && j < Kin) w = wsoft[((size_t)(pg * nOc + o / OC) * Kin * Kin + (size_t)(i * Kin + j)) * OC + o % OC];
    const uint32_t h = __float_as_uint(w) & 0xffffe000u;
    const float lo = w - __uint_as_float(h);
    unsigned char* img = wimg + (size_t)r * rep_bytes + lvl_off + (size_t)q * (2u * npad * 64u);
    if (pack) {
      // one 128-byte row [Whi (Kin) | Whi (Kin) | Wlo (Kin) | 0]: the thread of column j writes its three copies and,
      // for the columns past 3*Kin, the zero padding
      auto put = [&](int col, uint32_t bits) {
        *reinterpret_cast<uint32_t*>(img + sw128_off((uint32_t)n, (uint32_t)col >> 2) + (uint32_t)(col & 3) * 4u) = bits;
      };
      if (j < Kin) {
        put(j, h); put(Kin + j, h); put(2 * Kin + j, __float_as_uint(lo));
      }
      if (3 * Kin + j < 32) put(3 * Kin + j, 0u);
      if (3 * Kin + 16 + j < 32) put(3 * Kin + 16 + j, 0u);
      continue;
    }
    const uint32_t off = sw64_off((uint32_t)n, (uint32_t)j >> 2) + (uint32_t)(j & 3) * 4u;
    *reinterpret_cast<uint32_t*>(img + off) = h;
    *reinterpret_cast<uint32_t*>(img + (size_t)npad * 64u + off) = __float_as_uint(lo);
  }
}

// out[b, c] = logsumexp_r part[r][c][b]  (+ the sample's -1/2 sum_f x_f^2 when the leaf GEMM left it out, see
// ratspn_leaf_mma.cu: groups the exact leaf kernel redid carry the term inside their leaf values already)
__global__ void tree_root_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int P, int C, int64_t B,
                                         int64_t Bp, const float* __restrict__ sqsum, const int* __restrict__ redo) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  pdl_wait();
  if (b >= B) return;
  const float add = (sqsum != nullptr && redo[b >> 5] == 0) ? sqsum[b] : 0.f;
  for (int c = 0; c < C; ++c) {
    // repetitions in blocks of 16 independent loads (the three dependent passes over global memory of the first
    // version made this 4 MB kernel take 13 us); online max / sum across blocks
    float m = -INFINITY, s = 0.f;
    bool nan = false;
    for (int p0 = 0; p0 < P; p0 += 16) {
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = (p0 + q < P) ? __ldcs(part + ((size_t)(p0 + q) * C + c) * Bp + b) : -INFINITY;
      float mb = v[0];
#pragma unroll
      for (int q = 1; q < 16; ++q) mb = fmaxf(mb, v[q]);
#pragma unroll
      for (int q = 0; q < 16; ++q) nan |= (v[q] != v[q]);
      const float mn = fmaxf(m, mb);
      if (fabsf(mn) <= FLT_MAX) {
        float sb = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) sb += __expf(v[q] - mn);        // padding: exp(-inf) = 0
        s = s * __expf(m - mn) + sb;                                 // m = -inf on the first block: s = 0 stays 0
      }
      m = mn;
    }
    float y = m;                    // +-inf maxima pass through like torch.logsumexp
    if (nan) y = NAN;
    else if (fabsf(m) <= FLT_MAX) y = m + __logf(s);
    out[(size_t)b * C + c] = y + add;
  }
}

template <int KL, int O>
int launch_tree(const TreeArgs& a, const CUtensorMap& tmap, size_t smem, cudaStream_t st) {
  auto kern = (a.depth <= 3) ? ratspn_tree_mma_kernel<KL, O, 3> : ratspn_tree_mma_kernel<KL, O, kTreeMaxD>;
  DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)a.R, (unsigned)a.n_chunks);
  DPK_CUDA_TRY(launch_pdl(a.pdl != 0, kern, grid, dim3((unsigned)(a.G * kTile)), smem, st, a, tmap));
  DPK_LAUNCH_CHECK("ratspn_tree_mma_kernel");
  return DPK_OK;
}

}  // namespace

// (K_leaf, O) pairs the kernel is instantiated for
bool ratspn_tree_instantiated(int KL, int O) {
  return (KL == 10 && O == 10) || (KL == 16 && O == 16) || (KL == 8 && O == 8) || (KL == 4 && O == 4) ||
         (KL == 4 && O == 2) || (KL == 2 && O == 2);
}

int ratspn_run_prep_tree(const RatPlan& p, float* ws, cudaStream_t st) {
  unsigned char* wimg = reinterpret_cast<unsigned char*>(ws + p.off_timg);
  for (int lvl = 0; lvl < p.depth; ++lvl) {
    const bool root = lvl == p.depth - 1;
    const int parts = 1 << (p.depth - 1 - lvl);
    const int Kin = p.act_ch[lvl];
    const float* src = root ? ws + p.off_rsoft : ws + p.off_wsoft[lvl];
    const int Nout = root ? p.C : p.O;
    const Chunking& ch = root ? p.cc : p.oc;
    const int64_t total = (int64_t)p.R * parts * p.tree_npad[lvl] * 16;
    ratspn_prep_tree_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 2048), 256, 0, st>>>(
        src, p.R * parts, parts, Kin, Nout, ch.chunk, ch.count, (int)p.tree_npad[lvl], p.tree_off[lvl], p.tree_rep_bytes,
        3 * Kin <= 24 ? 1 : 0, wimg);
    DPK_LAUNCH_CHECK("ratspn_prep_tree_kernel");
  }
  return DPK_OK;
}

int ratspn_run_tree(const RatPlan& p, float* ws, float* out, cudaStream_t st) {
  TreeArgs a;
  a.act0 = ws + p.off_act[0];
  a.act_cs = p.act0_cs; a.act_ts = p.act0_ts; a.tile_rows = p.act0_tiled ? p.G0 * p.K : 0; a.c0_step = p.act0_tiled ? 0 : kTile;
  a.pdl = p.leaf_stream;
  a.wimg = reinterpret_cast<const unsigned char*>(ws + p.off_timg);
  a.part = ws + p.off_rtmp;
  for (int e = 0; e < kTreeMaxD; ++e) a.wlog[e] = (e < p.n_sum) ? ws + p.off_wlog[e] : nullptr;
  a.rlog = ws + p.off_rlog;
  a.B = p.B; a.Bp = p.Bp;
  a.depth = p.depth; a.R = p.R; a.C = p.C; a.KL = p.K; a.O = p.O;
  a.OCc = p.oc.chunk; a.nOc = p.oc.count; a.CCc = p.cc.chunk; a.nCc = p.cc.count;
  a.n_tiles = (int)(p.Bp / kTile);
  a.G = p.tree_G; a.Tcols = p.tree_tcols;
  a.n_chunks = std::max(1, std::min(sm_count() / std::max(1, p.R), (int)ceil_div(a.n_tiles, a.G)));
  a.rep_bytes = p.tree_rep_bytes;
  for (int l = 0; l <= kTreeMaxD; ++l) { a.lvl_off[l] = l < p.depth ? p.tree_off[l] : 0; a.lvl_npad[l] = l < p.depth ? p.tree_npad[l] : 0; }
  const size_t smem = 1024 + (((size_t)p.tree_rep_bytes + 1023) & ~(size_t)1023) + (size_t)a.G * (kABytes + 2u * p.K * 512u) + 128;
  // leaf activations as a 2-D tensor [G0*K rows][Bp samples]; box = one pair of sibling regions x one sample tile
  alignas(64) CUtensorMap tmap;
  int rc = p.act0_tiled ? make_tensor_map_2d_f32(&tmap, a.act0, (uint64_t)(p.Bp / 128) * p.G0 * p.K, 128, 512, 2u * p.K, kTile)
                        : make_tensor_map_2d_f32(&tmap, a.act0, (uint64_t)p.G0 * p.K, (uint64_t)p.Bp, (uint64_t)p.Bp * 4, 2u * p.K, kTile);
  if (rc) return rc;
  rc = DPK_E_ARG;
  {
    ProfScope prof(CAT_EINSUM, st);
    const int O = (p.depth == 1) ? p.K : p.O;
    if (p.K == 10 && O == 10) rc = launch_tree<10, 10>(a, tmap, smem, st);
    else if (p.K == 16 && O == 16) rc = launch_tree<16, 16>(a, tmap, smem, st);
    else if (p.K == 8 && O == 8) rc = launch_tree<8, 8>(a, tmap, smem, st);
    else if (p.K == 4 && O == 4) rc = launch_tree<4, 4>(a, tmap, smem, st);
    else if (p.K == 4 && O == 2) rc = launch_tree<4, 2>(a, tmap, smem, st);
    else if (p.K == 2 && O == 2) rc = launch_tree<2, 2>(a, tmap, smem, st);
    else return set_error(DPK_E_ARG, "tree kernel not instantiated for K=%d O=%d", p.K, p.O);
  }
  if (rc) return rc;
  ProfScope prof(CAT_ROOT, st);
  DPK_CUDA_TRY(launch_pdl(p.leaf_stream != 0, tree_root_combine_kernel, dim3((unsigned)ceil_div(p.B, 256)), dim3(256), 0, st,
                          (const float*)(ws + p.off_rtmp), out, p.R, p.C, p.B, p.Bp,
                          (const float*)(p.off_sqsum ? ws + p.off_sqsum : nullptr),
                          reinterpret_cast<const int*>(ws + p.off_mflags)));
  DPK_LAUNCH_CHECK("tree_root_combine_kernel");
  return DPK_OK;
}

}  // namespace dpk
