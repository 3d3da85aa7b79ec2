// dgcspn.cu -- DGC-SPN layers (NCHW fp32, the reference's tensor layout) with their backward passes.
//
// Reference (deeprob-kit, paths relative to /root/reference):
//   SpatialGaussianLayer.forward  deeprob/spn/layers/dgcspn.py:101-120  per-pixel Normal LL, NaN -> 0, sum over C_in
//   SpatialProductLayer.forward   deeprob/spn/layers/dgcspn.py:224-236  zero pad + 2x2 dilated conv with 0/1 weights
//   SpatialSumLayer.forward       deeprob/spn/layers/dgcspn.py:289-304  logsumexp_i(x + log_softmax_i W) per pixel
//   SpatialRootLayer.forward      deeprob/spn/layers/dgcspn.py:343-355  flatten + weighted logsumexp
// The reference materialises (B,C_out,C_in,H,W) in the sum layer and runs the product as a real
// convolution multiplying by 1.0; here the product is a 4-tap gather-add and the sum layer is a
// per-pixel mixture in the linear domain (one exp per input, max-shifted, exact log-domain fallback).
// Lanes run along the contiguous H*W axis, so every access of x / W / out is coalesced, and the
// per-pixel mixture weights are read once per CTA and reused over a slice of the batch.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace dpk {

// =================================================================================================
// Leaf
// =================================================================================================
__global__ void dgc_leaf_fwd_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                    const float* __restrict__ scale, float* __restrict__ out, int64_t B, int Cin, int K,
                                    int HW) {
  // blockIdx.x = (b, k) plane; threads sweep the pixels (no per-element 64-bit div/mod)
  for (int64_t plane = blockIdx.x; plane < B * K; plane += gridDim.x) {
    const int k = (int)(plane % K);
    const int64_t b = plane / K;
    const float* xb = x + b * Cin * HW;
    float* ob = out + plane * HW;
    for (int hw = threadIdx.x; hw < HW; hw += blockDim.x) {
      float acc = 0.f;
      for (int c = 0; c < Cin; ++c) {
        const float xv = xb[c * HW + hw];
        const int pi = (k * Cin + c) * HW + hw;
        const float sg = __ldg(scale + pi), mu = __ldg(loc + pi);
        const float t = (xv - mu) / sg;
        acc += nan_to_num(-0.5f * t * t - logf(sg) - kLogSqrt2Pi);
      }
      ob[hw] = acc;
    }
  }
}

// thread = pixel, CTA.y = batch slice: the (mu, 1/sigma, -log sigma - log sqrt(2 pi)) of KC components x CIN input
// channels stay in registers over the slice, so a sample costs one load per channel, ~6 instructions per output and
// one coalesced store per component (the per-plane kernel above redoes the division and the logarithm for every
// element: ~40 instructions per output, 4x off the HBM time of its 0.87 GB).
template <int CIN, int KC>
__global__ void __launch_bounds__(128) dgc_leaf_fwd_px_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                                              const float* __restrict__ scale, float* __restrict__ out,
                                                              int64_t B, int K, int HW, int64_t per_slice) {
  const int hw = blockIdx.x * 128 + threadIdx.x;
  if (hw >= HW) return;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  const float* __restrict__ xs = x + (size_t)b0 * CIN * HW;
  float* __restrict__ os = out + (size_t)b0 * K * HW;
  const unsigned nb = (unsigned)(b1 - b0), sHW = (unsigned)HW;
  for (int k0 = 0; k0 < K; k0 += KC) {
    float mu[KC][CIN], inv[KC][CIN], lg[KC][CIN];
#pragma unroll
    for (int k = 0; k < KC; ++k)
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const int pi = (min(k0 + k, K - 1) * CIN + c) * HW + hw;
        const float sg = __ldg(scale + pi);
        mu[k][c] = __ldg(loc + pi);
        inv[k][c] = 1.0f / sg;
        lg[k][c] = -logf(sg) - kLogSqrt2Pi;
      }
    constexpr int NB = 4;
    for (unsigned rb = 0; rb < nb; rb += NB) {
      float xv[NB][CIN];
#pragma unroll
      for (int s = 0; s < NB; ++s)
#pragma unroll
        for (int c = 0; c < CIN; ++c) xv[s][c] = xs[(min(rb + s, nb - 1) * CIN + c) * sHW + hw];
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (rb + s >= nb) continue;
        const unsigned ob = ((rb + s) * (unsigned)K + k0) * sHW + hw;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          if (k0 + k >= K) continue;
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const float t = (xv[s][c] - mu[k][c]) * inv[k][c];
            acc += nan_to_num(fmaf(-0.5f * t, t, lg[k][c]));
          }
          os[ob + k * sHW] = acc;
        }
      }
    }
  }
}

// gx[b,c,hw] = sum_k g * -(x-mu)/s^2   (zero where x is non-finite: nan_to_num'ed terms have no gradient)
__global__ void dgc_leaf_bwd_x_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                      const float* __restrict__ scale, const float* __restrict__ g,
                                      float* __restrict__ gx, int64_t B, int Cin, int K, int HW) {
  for (int64_t plane = blockIdx.x; plane < B * Cin; plane += gridDim.x) {
    const int c = (int)(plane % Cin);
    const int64_t b = plane / Cin;
    const float* gb = g + b * K * HW;
    for (int hw = threadIdx.x; hw < HW; hw += blockDim.x) {
      const float xv = x[plane * HW + hw];
      float acc = 0.f;
      if (fabsf(xv) <= FLT_MAX) {
        for (int k = 0; k < K; ++k) {
          const int pi = (k * Cin + c) * HW + hw;
          const float sg = __ldg(scale + pi), mu = __ldg(loc + pi);
          acc -= gb[k * HW + hw] * (xv - mu) / (sg * sg);
        }
      }
      gx[plane * HW + hw] = acc;
    }
  }
}

// thread = parameter element (k,c,hw); batch split over blockIdx.y, partial sums merged with atomics
__global__ void dgc_leaf_bwd_param_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                          const float* __restrict__ scale, const float* __restrict__ g,
                                          float* __restrict__ gloc, float* __restrict__ gscale, int64_t B, int Cin, int K,
                                          int HW, int64_t per_slice) {
  const int64_t n = (int64_t)K * Cin * HW;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int hw = (int)(idx % HW);
  const int c = (int)((idx / HW) % Cin);
  const int k = (int)(idx / ((int64_t)HW * Cin));
  const float mu = loc[idx], sg = scale[idx];
  const float inv = 1.0f / sg, inv2 = inv * inv;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  float s_mu = 0.f, s_sg = 0.f;
  for (int64_t b = b0; b < b1; ++b) {
    const float xv = x[(b * Cin + c) * HW + hw];
    if (!(fabsf(xv) <= FLT_MAX)) continue;
    const float gv = g[(b * K + k) * HW + hw];
    const float d = xv - mu;
    s_mu = fmaf(gv, d * inv2, s_mu);
    s_sg = fmaf(gv, d * d * inv2 * inv - inv, s_sg);
  }
  if (gloc && s_mu != 0.f) atomicAdd(gloc + idx, s_mu);
  if (gscale && s_sg != 0.f) atomicAdd(gscale + idx, s_sg);
}


// Pixel-stationary variant for at most KC components: thread = (input channel, pixel), the components' (mu, 1/sigma) and
// both accumulators in registers, x read once per sample instead of once per component, loads of sample b + 1 issued
// before the arithmetic of sample b (the kernel above runs at 0.5 TB/s: two dependent loads per two FMAs).
template <int KC>
__global__ void __launch_bounds__(128) dgc_leaf_bwd_param_px_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                                                    const float* __restrict__ scale, const float* __restrict__ g,
                                                                    float* __restrict__ gloc, float* __restrict__ gscale,
                                                                    int64_t B, int Cin, int K, int HW, int64_t per_slice) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Cin * HW) return;
  const int c = p / HW, hw = p - c * HW;
  float mu[KC], inv[KC], s_mu[KC], s_sg[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const size_t idx = ((size_t)min(k, K - 1) * Cin + c) * HW + hw;
    mu[k] = loc[idx];
    inv[k] = 1.0f / scale[idx];
    s_mu[k] = 0.f; s_sg[k] = 0.f;
  }
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  if (b0 >= b1) return;
  const float* xp = x + ((size_t)b0 * Cin + c) * HW + hw;
  const float* gp = g + (size_t)b0 * K * HW + hw;
  const size_t sx = (size_t)Cin * HW, sg = (size_t)K * HW;
  float xn, gn[KC];
  auto load = [&](const float* xq, const float* gq) {
    xn = __ldg(xq);
#pragma unroll
    for (int k = 0; k < KC; ++k) gn[k] = (k < K) ? __ldcs(gq + (size_t)k * HW) : 0.f;
  };
  load(xp, gp);
  for (int64_t b = b0; b < b1; ++b) {
    const float xv = xn;
    float gv[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) gv[k] = gn[k];
    xp += sx; gp += sg;
    if (b + 1 < b1) load(xp, gp);
    if (!(fabsf(xv) <= FLT_MAX)) continue;       // marginalised variable: no gradient
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float d = xv - mu[k], inv2 = inv[k] * inv[k];
      s_mu[k] = fmaf(gv[k], d * inv2, s_mu[k]);
      s_sg[k] = fmaf(gv[k], d * d * inv2 * inv[k] - inv[k], s_sg[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    if (k >= K) break;
    const size_t idx = ((size_t)k * Cin + c) * HW + hw;
    if (gloc && s_mu[k] != 0.f) atomicAdd(gloc + idx, s_mu[k]);
    if (gscale && s_sg[k] != 0.f) atomicAdd(gscale + idx, s_sg[k]);
  }
}

// =================================================================================================
// Product (2x2 taps, dilation, stride, zero padding = log 1)
// =================================================================================================
struct ProdDesc {
  int C, H, W, OC, OH, OW, pad_top, pad_left, sh, sw, dh, dw, depthwise;
};

__device__ __forceinline__ int prod_in_channel(const ProdDesc& d, int oc, int tap) {
  if (d.depthwise) return oc;
  // itertools.product(range(C), repeat=4): tap 0 (kh=0,kw=0) is the most significant digit
  int div = 1;
  for (int t = 3; t > tap; --t) div *= d.C;
  return (oc / div) % d.C;
}

__global__ void dgc_product_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t B, ProdDesc d) {
  // blockIdx.x = (b, oc) output plane; threads sweep its pixels
  const int OHW = d.OH * d.OW;
  for (int64_t plane = blockIdx.x; plane < B * d.OC; plane += gridDim.x) {
    const int oc = (int)(plane % d.OC);
    const int64_t b = plane / d.OC;
    const float* src[4];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) src[tap] = x + (b * d.C + prod_in_channel(d, oc, tap)) * d.H * d.W;
    float* ob = out + plane * OHW;
    for (int p = threadIdx.x; p < OHW; p += blockDim.x) {
      const int oh = p / d.OW, ow = p - oh * d.OW;
      float acc = 0.f;
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        const int y = oh * d.sh + (tap >> 1) * d.dh - d.pad_top;
        const int xx = ow * d.sw + (tap & 1) * d.dw - d.pad_left;
        if (y >= 0 && y < d.H && xx >= 0 && xx < d.W) acc += src[tap][y * d.W + xx];
      }
      ob[p] = acc;
    }
  }
}

// depthwise: gather form (CTA = input plane); otherwise scatter with atomics (CTA = output plane)
__global__ void dgc_product_bwd_depthwise_kernel(const float* __restrict__ g, float* __restrict__ gx, int64_t B, ProdDesc d) {
  const int HWi = d.H * d.W, OHW = d.OH * d.OW;
  for (int64_t plane = blockIdx.x; plane < B * d.C; plane += gridDim.x) {
    const float* gp = g + plane * OHW;
    float* op = gx + plane * HWi;
    for (int p = threadIdx.x; p < HWi; p += blockDim.x) {
      const int y = p / d.W, xx = p - y * d.W;
      float acc = 0.f;
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        const int ny = y + d.pad_top - (tap >> 1) * d.dh, nx = xx + d.pad_left - (tap & 1) * d.dw;
        if (ny < 0 || nx < 0 || ny % d.sh || nx % d.sw) continue;
        const int oh = ny / d.sh, ow = nx / d.sw;
        if (oh < d.OH && ow < d.OW) acc += gp[oh * d.OW + ow];
      }
      op[p] = acc;
    }
  }
}


// Depthwise product, forward and backward, pixel-stationary: thread = destination pixel, its (up to) four source offsets
// are computed once -- the plane-per-CTA kernels above spend their time on two integer divisions and four bounds tests
// per element and reach a tenth of the HBM rate -- and the planes (b, c) of a slice stream through, four at a time.
//   BWD = false: out[plane][q] = sum_tap x[plane][y(q,tap), x(q,tap)]        (src = product input,  dst = its output)
//   BWD = true:  gx[plane][p]  = sum_tap g[plane][oh(p,tap), ow(p,tap)]       (src = output gradient, dst = input gradient)
template <bool BWD>
__global__ void __launch_bounds__(128) dgc_product_px_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int64_t planes, int64_t per_slice, ProdDesc d) {
  const int n_dst = BWD ? d.H * d.W : d.OH * d.OW;
  const int n_src = BWD ? d.OH * d.OW : d.H * d.W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_dst) return;
  int off[4];
  if (BWD) {
    const int y = p / d.W, xx = p - y * d.W;
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int ny = y + d.pad_top - (tap >> 1) * d.dh, nx = xx + d.pad_left - (tap & 1) * d.dw;
      off[tap] = -1;
      if (ny >= 0 && nx >= 0 && ny % d.sh == 0 && nx % d.sw == 0) {
        const int oh = ny / d.sh, ow = nx / d.sw;
        if (oh < d.OH && ow < d.OW) off[tap] = oh * d.OW + ow;
      }
    }
  } else {
    const int oh = p / d.OW, ow = p - oh * d.OW;
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int y = oh * d.sh + (tap >> 1) * d.dh - d.pad_top, xx = ow * d.sw + (tap & 1) * d.dw - d.pad_left;
      off[tap] = (y >= 0 && y < d.H && xx >= 0 && xx < d.W) ? y * d.W + xx : -1;
    }
  }
  const int64_t p0 = blockIdx.y * per_slice, p1 = min((long long)planes, (long long)(p0 + per_slice));
  const float* sp = src + p0 * n_src;
  float* dp = dst + p0 * n_dst + p;
  int64_t pl = p0;
  for (; pl + 4 <= p1; pl += 4) {
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) v[u][tap] = off[tap] >= 0 ? __ldg(sp + (size_t)u * n_src + off[tap]) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) __stcs(dp + (size_t)u * n_dst, (v[u][0] + v[u][1]) + (v[u][2] + v[u][3]));
    sp += (size_t)4 * n_src;
    dp += (size_t)4 * n_dst;
  }
  for (; pl < p1; ++pl) {
    float a = 0.f;
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) a += off[tap] >= 0 ? __ldg(sp + off[tap]) : 0.f;
    *dp = a;
    sp += n_src;
    dp += n_dst;
  }
}

__global__ void dgc_product_bwd_scatter_kernel(const float* __restrict__ g, float* __restrict__ gx, int64_t B, ProdDesc d) {
  const int OHW = d.OH * d.OW;
  for (int64_t plane = blockIdx.x; plane < B * d.OC; plane += gridDim.x) {
    const int oc = (int)(plane % d.OC);
    const int64_t b = plane / d.OC;
    const float* gp = g + plane * OHW;
    float* dst[4];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) dst[tap] = gx + (b * d.C + prod_in_channel(d, oc, tap)) * d.H * d.W;
    for (int p = threadIdx.x; p < OHW; p += blockDim.x) {
      const float gv = gp[p];
      if (gv == 0.f) continue;
      const int oh = p / d.OW, ow = p - oh * d.OW;
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        const int y = oh * d.sh + (tap >> 1) * d.dh - d.pad_top;
        const int xx = ow * d.sw + (tap & 1) * d.dw - d.pad_left;
        if (y >= 0 && y < d.H && xx >= 0 && xx < d.W) atomicAdd(dst[tap] + y * d.W + xx, gv);
      }
    }
  }
}

// =================================================================================================
// Per-pixel mixture (SpatialSumLayer)
// =================================================================================================
// softmax / log-softmax over the input-channel axis of weight (O, I, HW); thread = (o, hw)
__global__ void dgc_sum_prep_kernel(const float* __restrict__ w, float* __restrict__ wsoft, float* __restrict__ wlog,
                                    int O, int I, int HW) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O * HW) return;
  const int hw = idx % HW, o = idx / HW;
  const float* base = w + (size_t)o * I * HW + hw;
  float m = -INFINITY;
  for (int i = 0; i < I; ++i) m = fmaxf(m, base[(size_t)i * HW]);
  float s = 0.f;
  for (int i = 0; i < I; ++i) s += expf(base[(size_t)i * HW] - m);
  const float lse = m + logf(s);
  for (int i = 0; i < I; ++i) {
    const float lw = base[(size_t)i * HW] - lse;
    wlog[(size_t)o * I * HW + (size_t)i * HW + hw] = lw;
    wsoft[(size_t)o * I * HW + (size_t)i * HW + hw] = expf(lw);
  }
}

constexpr int kSumNB = 4;  // samples a thread carries at once (weight loads amortised over them)

// thread = pixel; CTA.y = batch slice; loops over output chunks of OC and over the slice in groups of NB
template <int OC>
__global__ void __launch_bounds__(128) dgc_sum_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wsoft,
                                                          const float* __restrict__ wlog, float* __restrict__ out,
                                                          int64_t B, int I, int O, int HW, int64_t per_slice) {
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= HW) return;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  for (int o0 = 0; o0 < O; o0 += OC) {
    for (int64_t bb = b0; bb < b1; bb += kSumNB) {
      float m[kSumNB], acc[kSumNB][OC];
#pragma unroll
      for (int s = 0; s < kSumNB; ++s) {
        m[s] = -INFINITY;
#pragma unroll
        for (int o = 0; o < OC; ++o) acc[s][o] = 0.f;
      }
      for (int i = 0; i < I; ++i) {
        float w[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) w[o] = (o0 + o < O) ? __ldg(wsoft + ((size_t)(o0 + o) * I + i) * HW + hw) : 0.f;
#pragma unroll
        for (int s = 0; s < kSumNB; ++s) {
          if (bb + s >= b1) continue;
          const float xv = x[((bb + s) * I + i) * HW + hw];
          if (xv > m[s]) {                       // online max: rescale what has been accumulated so far
            const float sc = __expf(m[s] - xv);  // m = -inf -> 0
#pragma unroll
            for (int o = 0; o < OC; ++o) acc[s][o] *= sc;
            m[s] = xv;
          }
          const float e = (m[s] == -INFINITY) ? 0.f : __expf(xv - m[s]);
#pragma unroll
          for (int o = 0; o < OC; ++o) acc[s][o] = fmaf(w[o], e, acc[s][o]);
        }
      }
#pragma unroll
      for (int s = 0; s < kSumNB; ++s) {
        if (bb + s >= b1) continue;
#pragma unroll
        for (int o = 0; o < OC; ++o) {
          if (o0 + o >= O) continue;
          float y;
          if (acc[s][o] >= 1e-18f && acc[s][o] <= FLT_MAX && fabsf(m[s]) <= FLT_MAX) {
            y = m[s] + __logf(acc[s][o]);
          } else {  // exact log-domain evaluation
            float mm = -INFINITY;
            for (int i = 0; i < I; ++i)
              mm = fmaxf(mm, x[((bb + s) * I + i) * HW + hw] + wlog[((size_t)(o0 + o) * I + i) * HW + hw]);
            if (!(fabsf(mm) <= FLT_MAX)) {
              y = mm;
            } else {
              float ss = 0.f;
              for (int i = 0; i < I; ++i)
                ss += expf(x[((bb + s) * I + i) * HW + hw] + wlog[((size_t)(o0 + o) * I + i) * HW + hw] - mm);
              y = mm + logf(ss);
            }
          }
          out[((bb + s) * O + o0 + o) * HW + hw] = y;
        }
      }
    }
  }
}

// Fast path when the (I x OC) weight block of a 128-pixel tile fits shared memory: the weights are staged
// ONCE per CTA as [i][o][pixel] (each lane reads its own column: conflict-free) and reused over the whole batch
// slice; a thread carries NB = 8 samples x OC outputs in registers, so every weight load feeds 8 FMAs.
constexpr int kSumFastNB = 8;
template <int OC>
__global__ void __launch_bounds__(128) dgc_sum_fwd_smem_kernel(const float* __restrict__ x, const float* __restrict__ wsoft,
                                                               const float* __restrict__ wlog, float* __restrict__ out,
                                                               int64_t B, int I, int O, int HW, int64_t per_slice) {
  extern __shared__ float wsm[];  // [I][OC][128]
  const int lane_px = threadIdx.x;
  const int hw = blockIdx.x * 128 + lane_px;
  const bool live = hw < HW;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  for (int o0 = 0; o0 < O; o0 += OC) {
    __syncthreads();
    for (int i = 0; i < I; ++i)
#pragma unroll
      for (int o = 0; o < OC; ++o)
        wsm[(i * OC + o) * 128 + lane_px] = (live && o0 + o < O) ? __ldg(wsoft + ((size_t)(o0 + o) * I + i) * HW + hw) : 0.f;
    __syncthreads();
    if (!live) continue;
    for (int64_t bb = b0; bb < b1; bb += kSumFastNB) {
      float m[kSumFastNB], acc[kSumFastNB][OC];
      const float* xrow[kSumFastNB];   // rows past the slice are clamped: loads stay unconditional (8 in flight)
#pragma unroll
      for (int s = 0; s < kSumFastNB; ++s) {
        xrow[s] = x + (size_t)min((long long)(bb + s), (long long)(b1 - 1)) * I * HW + hw;
        m[s] = -INFINITY;
#pragma unroll
        for (int o = 0; o < OC; ++o) acc[s][o] = 0.f;
      }
      // pass 1: per-sample max over the inputs (they stay hot in L1/L2 for pass 2)
#pragma unroll 2
      for (int i = 0; i < I; ++i) {
        float xv[kSumFastNB];
#pragma unroll
        for (int s = 0; s < kSumFastNB; ++s) xv[s] = xrow[s][(size_t)i * HW];
#pragma unroll
        for (int s = 0; s < kSumFastNB; ++s) m[s] = fmaxf(m[s], xv[s]);
      }
#pragma unroll
      for (int s = 0; s < kSumFastNB; ++s)
        if (!(fabsf(m[s]) <= FLT_MAX)) m[s] = 0.f;     // all -inf / non-finite: the exact path sorts it out
#pragma unroll 2
      for (int i = 0; i < I; ++i) {
        float w[OC], xv[kSumFastNB];
#pragma unroll
        for (int s = 0; s < kSumFastNB; ++s) xv[s] = xrow[s][(size_t)i * HW];
#pragma unroll
        for (int o = 0; o < OC; ++o) w[o] = wsm[(i * OC + o) * 128 + lane_px];
#pragma unroll
        for (int s = 0; s < kSumFastNB; ++s) {
          const float e = __expf(xv[s] - m[s]);
#pragma unroll
          for (int o = 0; o < OC; ++o) acc[s][o] = fmaf(w[o], e, acc[s][o]);
        }
      }
#pragma unroll
      for (int s = 0; s < kSumFastNB; ++s) {
        if (bb + s >= b1) continue;
#pragma unroll
        for (int o = 0; o < OC; ++o) {
          if (o0 + o >= O) continue;
          float y;
          if (acc[s][o] >= 1e-18f && acc[s][o] <= FLT_MAX) {
            y = m[s] + __logf(acc[s][o]);
          } else {  // exact log-domain evaluation
            float mm = -INFINITY;
            for (int i = 0; i < I; ++i)
              mm = fmaxf(mm, x[((bb + s) * I + i) * HW + hw] + wlog[((size_t)(o0 + o) * I + i) * HW + hw]);
            if (!(fabsf(mm) <= FLT_MAX)) {
              y = mm;
            } else {
              float ss = 0.f;
              for (int i = 0; i < I; ++i)
                ss += expf(x[((bb + s) * I + i) * HW + hw] + wlog[((size_t)(o0 + o) * I + i) * HW + hw] - mm);
              y = mm + logf(ss);
            }
          }
          out[((bb + s) * O + o0 + o) * HW + hw] = y;
        }
      }
    }
  }
}

__device__ float g_dgc_zero = 0.f;   // what a tap in the zero padding reads
__device__ __forceinline__ float ld_nc_f32(unsigned long long addr) {
  float v;
  asm("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(addr));
  return v;
}
__device__ __forceinline__ void st_f32(unsigned long long addr, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// lg2.approx.ftz without the denormal prescaling of __logf: the caller has sums >= 1e-18 on this path
__device__ __forceinline__ float lg2_fast(float v) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// Fused depthwise SpatialProductLayer + SpatialSumLayer (inference): the product output -- the largest tensor of
// the model, written once and read once by the unfused pair -- never touches HBM.  thread = output pixel; the four
// taps of every input channel are gathered straight from the product layer's input (lanes = consecutive pixels:
// row-wise coalesced; the 2x2 neighbourhoods overlap between lanes and rows, so most taps hit L1/L2), the IC
// product values of NB samples stay in registers for the max pass, the mixture pass and the exact fallback.
// WG = true (32 channels): the weight block of a tile would need 128 KB of shared memory (one CTA per SM), so the
// weights are read through L1 instead (each thread re-reads its own column for every sample group: they stay
// cached) and the CTA may be 64 threads wide for layers of at most 64 pixels.
template <int IC> struct PsNB { static constexpr int value = IC <= 8 ? 4 : 2; };   // samples a thread carries at once
template <int IC, int OC, bool WG = false>
__global__ void __launch_bounds__(128, 3) dgc_prodsum_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wsoft,
                                                              const float* __restrict__ wlog, float* __restrict__ out,
                                                              int64_t B, int O, ProdDesc d, int64_t per_slice) {
  extern __shared__ float wsm[];  // [IC][OC][128]
  constexpr int kPsNB = PsNB<IC>::value;
  const int lane_px = threadIdx.x;
  const int HW = d.OH * d.OW, IHW = d.H * d.W;
  const int hw = blockIdx.x * blockDim.x + lane_px;
  const bool live = hw < HW;
  const int oh = live ? hw / d.OW : 0, ow = live ? hw - (hw / d.OW) * d.OW : 0;
  int off[4];
  bool ok[4];
#pragma unroll
  for (int tap = 0; tap < 4; ++tap) {
    const int y = oh * d.sh + (tap >> 1) * d.dh - d.pad_top;
    const int xx = ow * d.sw + (tap & 1) * d.dw - d.pad_left;
    ok[tap] = live && y >= 0 && y < d.H && xx >= 0 && xx < d.W;
    off[tap] = ok[tap] ? y * d.W + xx : 0;
  }
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  for (int o0 = 0; o0 < O; o0 += OC) {
    if constexpr (!WG) {
      __syncthreads();
      for (int i = 0; i < IC; ++i)
#pragma unroll
        for (int o = 0; o < OC; ++o)
          wsm[(i * OC + o) * 128 + lane_px] = (live && o0 + o < O) ? __ldg(wsoft + ((size_t)(o0 + o) * IC + i) * HW + hw) : 0.f;
      __syncthreads();
    }
    if (!live) continue;
    const float* __restrict__ wcol = wsoft + (size_t)o0 * IC * HW + hw;   // WG: this thread's weight column
    // 32-bit element offsets against CTA-uniform slice bases (the host keeps a slice below 2^31 elements): one
    // integer add per load instead of a 64-bit address chain -- the kernel is instruction bound, not HBM bound
    const float* __restrict__ xs = x + (size_t)b0 * IC * IHW;
    float* __restrict__ os = out + (size_t)b0 * O * HW;
    const unsigned nb = (unsigned)(b1 - b0);
    const unsigned sIHW = (unsigned)IHW, sHW = (unsigned)HW;
    const bool all_out = o0 + OC <= O;
    // Explicit 64-bit byte addresses, one per tap, stepped from plane to plane (a 64-bit add per load; left to the
    // compiler every load re-derives its address from the kernel argument: six instructions).  A tap in the zero
    // padding reads a zero word with step 0, so the loads need no predicate.
    unsigned long long ta[4];
    unsigned tstep[4], tsamp[4];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      ta[tap] = ok[tap] ? (unsigned long long)(xs + off[tap]) : (unsigned long long)&g_dgc_zero;
      tstep[tap] = ok[tap] ? sIHW * 4u : 0u;
      tsamp[tap] = ok[tap] ? IC * sIHW * 4u : 0u;
    }
    for (unsigned rb = 0; rb < nb; rb += kPsNB) {
      float pv[kPsNB][IC], m[kPsNB];
#pragma unroll
      for (int s = 0; s < kPsNB; ++s) {
        const unsigned sb = min(rb + s, nb - 1);   // clamped: loads unconditional
        unsigned long long q0 = ta[0] + (unsigned long long)sb * tsamp[0];
        unsigned long long q1 = ta[1] + (unsigned long long)sb * tsamp[1];
        unsigned long long q2 = ta[2] + (unsigned long long)sb * tsamp[2];
        unsigned long long q3 = ta[3] + (unsigned long long)sb * tsamp[3];
#pragma unroll
        for (int i = 0; i < IC; ++i) {
          pv[s][i] = (ld_nc_f32(q0) + ld_nc_f32(q1)) + (ld_nc_f32(q2) + ld_nc_f32(q3));   // zero padding = log 1
          q0 += tstep[0]; q1 += tstep[1]; q2 += tstep[2]; q3 += tstep[3];
        }
      }
#pragma unroll
      for (int s = 0; s < kPsNB; ++s) {
        float mm = -INFINITY;
#pragma unroll
        for (int i = 0; i < IC; ++i) mm = fmaxf(mm, pv[s][i]);
        m[s] = (fabsf(mm) <= FLT_MAX) ? mm : 0.f;   // all -inf / non-finite: the exact path sorts it out
      }
      float acc[kPsNB][OC];
      float2 acc2[kPsNB][OC / 2];
#pragma unroll
      for (int s = 0; s < kPsNB; ++s)
#pragma unroll
        for (int h = 0; h < OC / 2; ++h) acc2[s][h] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < IC; ++i) {
        float w[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) {
          if constexpr (WG) w[o] = (o0 + o < O) ? __ldg(wcol + (unsigned)((o * IC + i) * HW)) : 0.f;
          else w[o] = wsm[(i * OC + o) * 128 + lane_px];
        }
#pragma unroll
        for (int s = 0; s < kPsNB; ++s) {
          const float e = __expf(pv[s][i] - m[s]);
          const float2 e2 = make_float2(e, e);
#pragma unroll
          for (int h = 0; h < OC / 2; ++h)          // packed FFMA2: two outputs per instruction
            acc2[s][h] = __ffma2_rn(make_float2(w[2 * h], w[2 * h + 1]), e2, acc2[s][h]);
        }
      }
#pragma unroll
      for (int s = 0; s < kPsNB; ++s)
#pragma unroll
        for (int h = 0; h < OC / 2; ++h) { acc[s][2 * h] = acc2[s][h].x; acc[s][2 * h + 1] = acc2[s][h].y; }
      // common case first, without a branch per output; samples with an out-of-range sum are redone exactly.
      // Range check of all sums at once: the minimum catches tiny / negative values, the total catches inf and NaN.
      float amin = FLT_MAX, atot = 0.f;
      if (all_out && rb + kPsNB <= nb) {          // whole group: no per-output conditions
#pragma unroll
        for (int s = 0; s < kPsNB; ++s) {
          unsigned long long op = (unsigned long long)(os + hw) + (unsigned long long)((rb + s) * (unsigned)O + o0) * (sHW * 4u);
#pragma unroll
          for (int o = 0; o < OC; ++o) {
            amin = fminf(amin, acc[s][o]);
            atot += acc[s][o];
            st_f32(op, fmaf(lg2_fast(acc[s][o]), 0.6931471805599453f, m[s]));
            op += sHW * 4u;
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < kPsNB; ++s) {
          if (rb + s >= nb) continue;
          unsigned long long op = (unsigned long long)(os + hw) + (unsigned long long)((rb + s) * (unsigned)O + o0) * (sHW * 4u);
#pragma unroll
          for (int o = 0; o < OC; ++o) {
            if (o0 + o >= O) continue;
            amin = fminf(amin, acc[s][o]);
            atot += acc[s][o];
            st_f32(op + (unsigned long long)o * (sHW * 4u), fmaf(lg2_fast(acc[s][o]), 0.6931471805599453f, m[s]));
          }
        }
      }
      const bool redo = !(amin >= 1e-18f && atot <= FLT_MAX);
      if (redo) {
#pragma unroll 1
        for (int s = 0; s < kPsNB; ++s) {
          if (rb + s >= nb) continue;
#pragma unroll 1
          for (int o = 0; o < OC; ++o) {
            if (o0 + o >= O) continue;
            float av = 0.f;                      // runtime-indexed registers would go to local memory: select
#pragma unroll
            for (int s2 = 0; s2 < kPsNB; ++s2)
#pragma unroll
              for (int o2 = 0; o2 < OC; ++o2)
                if (s2 == s && o2 == o) av = acc[s2][o2];
            if (av >= 1e-18f && av <= FLT_MAX) continue;
            // exact log-domain evaluation
            float mm = -INFINITY;
            for (int i = 0; i < IC; ++i) {
              float pvi = 0.f;
#pragma unroll
              for (int s2 = 0; s2 < kPsNB; ++s2)
#pragma unroll
                for (int i2 = 0; i2 < IC; ++i2)
                  if (s2 == s && i2 == i) pvi = pv[s2][i2];
              mm = fmaxf(mm, pvi + wlog[((size_t)(o0 + o) * IC + i) * HW + hw]);
            }
            float y = mm;
            if (fabsf(mm) <= FLT_MAX) {
              float ss = 0.f;
              for (int i = 0; i < IC; ++i) {
                float pvi = 0.f;
#pragma unroll
                for (int s2 = 0; s2 < kPsNB; ++s2)
#pragma unroll
                  for (int i2 = 0; i2 < IC; ++i2)
                    if (s2 == s && i2 == i) pvi = pv[s2][i2];
                ss += expf(pvi + wlog[((size_t)(o0 + o) * IC + i) * HW + hw] - mm);
              }
              y = mm + logf(ss);
            }
            os[((rb + s) * (unsigned)O + o0 + o) * sHW + hw] = y;
          }
        }
      }
    }
  }
}

// Backward: posterior of input i under output o is w[o,i]*exp(x_i - y_o).
//   gx[b,i]    = sum_o g[b,o] * w[o,i] * exp(x_i - y_o)
//   N[o,i,hw] += sum_b g[b,o] * w[o,i] * exp(x_i - y_o)       (then grad_raw = N - softmax * sum_i N)
// thread = pixel, CTA.y = batch slice; the N accumulators of an (OC x IC) block stay in registers over
// the slice and are merged with atomics.
template <int OC, int IC>
__global__ void __launch_bounds__(128) dgc_sum_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wsoft,
                                                          const float* __restrict__ y, const float* __restrict__ g,
                                                          float* __restrict__ gx, float* __restrict__ nstat, int64_t B,
                                                          int I, int O, int HW, int64_t per_slice) {
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= HW) return;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  for (int i0 = 0; i0 < I; i0 += IC) {
    for (int o0 = 0; o0 < O; o0 += OC) {
      float w[OC][IC], n[OC][IC];
#pragma unroll
      for (int o = 0; o < OC; ++o)
#pragma unroll
        for (int i = 0; i < IC; ++i) {
          w[o][i] = (o0 + o < O && i0 + i < I) ? __ldg(wsoft + ((size_t)(o0 + o) * I + i0 + i) * HW + hw) : 0.f;
          n[o][i] = 0.f;
        }
      for (int64_t b = b0; b < b1; ++b) {
        float xv[IC], gi[IC];
#pragma unroll
        for (int i = 0; i < IC; ++i) { xv[i] = (i0 + i < I) ? x[(b * I + i0 + i) * HW + hw] : -INFINITY; gi[i] = 0.f; }
#pragma unroll
        for (int o = 0; o < OC; ++o) {
          if (o0 + o >= O) continue;
          const float gv = g[(b * O + o0 + o) * HW + hw];
          const float yv = y[(b * O + o0 + o) * HW + hw];
          if (gv == 0.f || !(fabsf(yv) <= FLT_MAX)) continue;
#pragma unroll
          for (int i = 0; i < IC; ++i) {
            const float post = gv * w[o][i] * __expf(fminf(xv[i] - yv, 80.f));
            n[o][i] += post;
            gi[i] += post;
          }
        }
        if (gx) {
#pragma unroll
          for (int i = 0; i < IC; ++i)
            if (i0 + i < I) {
              float* dst = gx + (b * I + i0 + i) * HW + hw;
              *dst = (o0 == 0) ? gi[i] : *dst + gi[i];   // output chunks accumulate in order within the thread
            }
        }
      }
      if (nstat) {
#pragma unroll
        for (int o = 0; o < OC; ++o)
#pragma unroll
          for (int i = 0; i < IC; ++i)
            if (o0 + o < O && i0 + i < I && n[o][i] != 0.f)
              atomicAdd(nstat + ((size_t)(o0 + o) * I + i0 + i) * HW + hw, n[o][i]);
      }
    }
  }
}


// Same contract, for layers whose channels fit one register block (O <= OC, I <= IC -- every layer of the benchmark
// configurations).  The posterior factorises, w[o,i] * exp(x_i - y_o) = w[o,i] * exp(x_i - m) * exp(m - y_o) with
// m = max_i x_i: OC + IC exps per (pixel, sample) instead of OC * IC (the MUFU pipe bounded the kernel above), and the
// loads of sample b + 1 are issued before the arithmetic of sample b.
// GATHER: x is the INPUT of the depthwise product layer in front of the sum layer (d describes it) and the product value is
// recomputed from its four taps -- the training form of the forward fusion never stores the product output.
template <int OC, int IC, bool GATHER>
__global__ void __launch_bounds__(128) dgc_sum_bwd_fast_kernel(const float* __restrict__ x, const float* __restrict__ wsoft,
                                                               const float* __restrict__ y, const float* __restrict__ g,
                                                               float* __restrict__ gx, float* __restrict__ nstat, int64_t B,
                                                               int I, int O, int HW, int64_t per_slice, ProdDesc d) {
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= HW) return;
  int off[4] = {0, 0, 0, 0};
  if (GATHER) {
    const int oh = hw / d.OW, ow = hw - oh * d.OW;
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int yy = oh * d.sh + (tap >> 1) * d.dh - d.pad_top, xx = ow * d.sw + (tap & 1) * d.dw - d.pad_left;
      off[tap] = (yy >= 0 && yy < d.H && xx >= 0 && xx < d.W) ? yy * d.W + xx : -1;
    }
  }
  const int HWin = GATHER ? d.H * d.W : HW;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  if (b0 >= b1) return;
  float w[OC][IC], n[OC][IC];
#pragma unroll
  for (int o = 0; o < OC; ++o)
#pragma unroll
    for (int i = 0; i < IC; ++i) {
      w[o][i] = (o < O && i < I) ? __ldg(wsoft + ((size_t)o * I + i) * HW + hw) : 0.f;
      n[o][i] = 0.f;
    }
  const size_t sx = (size_t)I * HWin, so = (size_t)O * HW, sg = (size_t)I * HW;
  const float* xp = x + b0 * sx + (GATHER ? 0 : hw);
  const float* yp = y + b0 * so + hw;
  const float* gp = g + b0 * so + hw;
  float* gxp = gx ? gx + b0 * sg + hw : nullptr;
  float xn[IC], yn[OC], gn[OC];
  auto load = [&](const float* xq, const float* yq, const float* gq) {
#pragma unroll
    for (int i = 0; i < IC; ++i) {
      if (GATHER) {
        float a = 0.f;
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) a += off[tap] >= 0 ? __ldg(xq + (size_t)i * HWin + off[tap]) : 0.f;
        xn[i] = (i < I) ? a : -INFINITY;
      } else {
        xn[i] = (i < I) ? __ldcs(xq + (size_t)i * HW) : -INFINITY;
      }
    }
#pragma unroll
    for (int o = 0; o < OC; ++o) {
      gn[o] = (o < O) ? __ldcs(gq + (size_t)o * HW) : 0.f;
      yn[o] = (o < O) ? __ldcs(yq + (size_t)o * HW) : 0.f;
    }
  };
  load(xp, yp, gp);
  for (int64_t b = b0; b < b1; ++b) {
    float xv[IC], f[OC];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < IC; ++i) { xv[i] = xn[i]; m = fmaxf(m, xv[i]); }
    if (!(fabsf(m) <= FLT_MAX)) m = 0.f;
#pragma unroll
    for (int o = 0; o < OC; ++o) f[o] = (gn[o] != 0.f && fabsf(yn[o]) <= FLT_MAX) ? gn[o] * __expf(fminf(m - yn[o], 80.f)) : 0.f;
    if (b + 1 < b1) load(xp + sx, yp + so, gp + so);
    float gi[IC];
#pragma unroll
    for (int i = 0; i < IC; ++i) {
      const float e = __expf(fminf(xv[i] - m, 80.f));
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < OC; ++o) {
        const float post = w[o][i] * (e * f[o]);
        n[o][i] += post;
        acc += post;
      }
      gi[i] = acc;
    }
    if (gxp) {
#pragma unroll
      for (int i = 0; i < IC; ++i)
        if (i < I) __stcs(gxp + (size_t)i * HW, gi[i]);
      gxp += sg;
    }
    xp += sx; yp += so; gp += so;
  }
  if (nstat) {
#pragma unroll
    for (int o = 0; o < OC; ++o)
#pragma unroll
      for (int i = 0; i < IC; ++i)
        if (o < O && i < I && n[o][i] != 0.f) atomicAdd(nstat + ((size_t)o * I + i) * HW + hw, n[o][i]);
  }
}


// More than OC x IC channels (the 16 / 32-channel layers of the MNIST-example model): the same factorised, prefetching
// loop per (input chunk = blockIdx.z, output chunk) block.  x of the chunk is re-read once per output chunk and d/dx is
// accumulated across output chunks by the owning thread (read-modify-write of its own element), all mostly out of L2:
// 2.5 TB/s of DRAM traffic against 0.29 for the one-sample-at-a-time kernel below.  The shift m is the chunk's own maximum:
// any finite shift factorises exactly, and y_o >= x_i + log w_oi bounds m - y_o from above.
template <int OC, int IC>
__global__ void __launch_bounds__(128) dgc_sum_bwd_chunk_kernel(const float* __restrict__ x, const float* __restrict__ wsoft,
                                                                const float* __restrict__ y, const float* __restrict__ g,
                                                                float* __restrict__ gx, float* __restrict__ nstat, int64_t B,
                                                                int I, int O, int HW, int64_t per_slice) {
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= HW) return;
  const int i0 = blockIdx.z * IC;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  if (b0 >= b1) return;
  const size_t sx = (size_t)I * HW, so = (size_t)O * HW;
  for (int o0 = 0; o0 < O; o0 += OC) {
    float w[OC][IC], n[OC][IC];
#pragma unroll
    for (int o = 0; o < OC; ++o)
#pragma unroll
      for (int i = 0; i < IC; ++i) {
        w[o][i] = (o0 + o < O && i0 + i < I) ? __ldg(wsoft + ((size_t)(o0 + o) * I + i0 + i) * HW + hw) : 0.f;
        n[o][i] = 0.f;
      }
    const float* xp = x + b0 * sx + (size_t)i0 * HW + hw;
    const float* yp = y + b0 * so + (size_t)o0 * HW + hw;
    const float* gp = g + b0 * so + (size_t)o0 * HW + hw;
    float* gxp = gx ? gx + b0 * sx + (size_t)i0 * HW + hw : nullptr;
    float xn[IC], yn[OC], gn[OC], on[IC];
    auto load = [&](const float* xq, const float* yq, const float* gq, const float* oq) {
#pragma unroll
      for (int i = 0; i < IC; ++i) {
        xn[i] = (i0 + i < I) ? __ldg(xq + (size_t)i * HW) : -INFINITY;
        on[i] = (oq != nullptr && o0 > 0 && i0 + i < I) ? __ldcg(oq + (size_t)i * HW) : 0.f;   // d/dx so far
      }
#pragma unroll
      for (int o = 0; o < OC; ++o) {
        gn[o] = (o0 + o < O) ? __ldg(gq + (size_t)o * HW) : 0.f;
        yn[o] = (o0 + o < O) ? __ldg(yq + (size_t)o * HW) : 0.f;
      }
    };
    load(xp, yp, gp, gxp);
    for (int64_t b = b0; b < b1; ++b) {
      float xv[IC], f[OC], ov[IC];
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < IC; ++i) { xv[i] = xn[i]; ov[i] = on[i]; m = fmaxf(m, xv[i]); }
      if (!(fabsf(m) <= FLT_MAX)) m = 0.f;
#pragma unroll
      for (int o = 0; o < OC; ++o) f[o] = (gn[o] != 0.f && fabsf(yn[o]) <= FLT_MAX) ? gn[o] * __expf(fminf(m - yn[o], 80.f)) : 0.f;
      float* gcur = gxp;
      xp += sx; yp += so; gp += so;
      if (gxp) gxp += sx;
      if (b + 1 < b1) load(xp, yp, gp, gxp);
#pragma unroll
      for (int i = 0; i < IC; ++i) {
        const float e = __expf(fminf(xv[i] - m, 80.f));
        float acc = ov[i];
#pragma unroll
        for (int o = 0; o < OC; ++o) {
          const float post = w[o][i] * (e * f[o]);
          n[o][i] += post;
          acc += post;
        }
        if (gcur && i0 + i < I) gcur[(size_t)i * HW] = acc;
      }
    }
    if (nstat) {
#pragma unroll
      for (int o = 0; o < OC; ++o)
#pragma unroll
        for (int i = 0; i < IC; ++i)
          if (o0 + o < O && i0 + i < I && n[o][i] != 0.f) atomicAdd(nstat + ((size_t)(o0 + o) * I + i0 + i) * HW + hw, n[o][i]);
    }
  }
}

// grad_raw[o,i,hw] += N - softmax * sum_i N
__global__ void dgc_sum_finalize_kernel(const float* __restrict__ wsoft, const float* __restrict__ nstat,
                                        float* __restrict__ gw, int O, int I, int HW) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O * HW) return;
  const int hw = idx % HW, o = idx / HW;
  const size_t base = (size_t)o * I * HW + hw;
  float tot = 0.f;
  for (int i = 0; i < I; ++i) tot += nstat[base + (size_t)i * HW];
  for (int i = 0; i < I; ++i) gw[base + (size_t)i * HW] += nstat[base + (size_t)i * HW] - wsoft[base + (size_t)i * HW] * tot;
}

// =================================================================================================
// Root: out[b,c] = logsumexp_q(x[b,q] + log_softmax_q W[c,q]);  one CTA per sample
// =================================================================================================
__global__ void dgc_root_prep_kernel(const float* __restrict__ w, float* __restrict__ wlog, int C, int64_t Q) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const float* row = w + (size_t)c * Q;
  float m = -INFINITY;
  for (int64_t q = threadIdx.x; q < Q; q += blockDim.x) m = fmaxf(m, row[q]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int64_t q = threadIdx.x; q < Q; q += blockDim.x) s += expf(row[q] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
  const float lse = m + logf(s);
  for (int64_t q = threadIdx.x; q < Q; q += blockDim.x) wlog[(size_t)c * Q + q] = row[q] - lse;
}

__global__ void __launch_bounds__(256) dgc_root_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wlog,
                                                           float* __restrict__ out, int64_t Q, int C) {
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  const float* row = x + b * Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < C; ++c) {
    const float* wl = wlog + (size_t)c * Q;
    float m = -INFINITY;
    for (int64_t q = threadIdx.x; q < Q; q += 256) m = fmaxf(m, row[q] + __ldg(wl + q));
    m = warp_max(m);
    __syncthreads();
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    float y = m;
    if (fabsf(m) <= FLT_MAX) {
      float s = 0.f;
      for (int64_t q = threadIdx.x; q < Q; q += 256) s += __expf(row[q] + __ldg(wl + q) - m);
      s = warp_sum(s);
      __syncthreads();
      if (lane == 0) red[warp] = s;
      __syncthreads();
      s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i];
      y = m + logf(s);
    }
    if (threadIdx.x == 0) out[b * C + c] = y;
  }
}

// gx[b,q] = sum_c g[b,c] * exp(x_q + lw_cq - out_bc);  N[c,q] += g[b,c] * exp(...)   (thread = q, batch slice = CTA.y)
__global__ void dgc_root_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wlog,
                                    const float* __restrict__ out, const float* __restrict__ g, float* __restrict__ gx,
                                    float* __restrict__ nstat, int64_t B, int64_t Q, int C, int64_t per_slice) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  for (int c = 0; c < C; ++c) {
    const float lw = wlog[(size_t)c * Q + q];
    float n = 0.f;
    for (int64_t b = b0; b < b1; ++b) {
      const float gv = g[b * C + c], ov = out[b * C + c];
      float post = 0.f;
      if (gv != 0.f && fabsf(ov) <= FLT_MAX) post = gv * __expf(fminf(x[b * Q + q] + lw - ov, 80.f));
      n += post;
      if (gx) {
        float* dst = gx + b * Q + q;
        *dst = (c == 0) ? post : *dst + post;
      }
    }
    if (nstat && n != 0.f) atomicAdd(nstat + (size_t)c * Q + q, n);
  }
}

// grad_raw[c,q] += N[c,q] - softmax[c,q] * sum_q N[c,q]
__global__ void dgc_root_finalize_kernel(const float* __restrict__ wlog, const float* __restrict__ nstat,
                                         float* __restrict__ gw, int64_t Q) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int64_t q = threadIdx.x; q < Q; q += blockDim.x) s += nstat[(size_t)c * Q + q];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
  for (int64_t q = threadIdx.x; q < Q; q += blockDim.x)
    gw[(size_t)c * Q + q] += nstat[(size_t)c * Q + q] - expf(wlog[(size_t)c * Q + q]) * s;
}

static int env_int_dgc(const char* n, int d) { const char* v = getenv(n); return (v && *v) ? atoi(v) : d; }

static int plane_grid(int64_t planes) { return (int)std::min<int64_t>(planes, (int64_t)sm_count() * 64); }

static int grid_for(int64_t total, int threads) {
  return (int)std::min<int64_t>(ceil_div(total, threads), (int64_t)sm_count() * 32);
}

// batch slices so that (pixel blocks x slices) fills the machine about 4x
static int64_t slice_len(int64_t B, int64_t blocks_x, int min_len) {
  int64_t slices = std::max<int64_t>(1, ceil_div((int64_t)4 * sm_count(), blocks_x));
  slices = std::min<int64_t>(slices, std::max<int64_t>(1, B / min_len));
  return round_up(ceil_div(B, slices), kSumNB);
}

static ProdDesc to_prod(const dpk_dgc_product_desc* d) {
  ProdDesc p;
  p.C = d->channels; p.H = d->height; p.W = d->width; p.OC = d->out_channels; p.OH = d->out_height; p.OW = d->out_width;
  p.pad_top = d->pad_top; p.pad_left = d->pad_left; p.sh = d->stride_h; p.sw = d->stride_w; p.dh = d->dilation_h;
  p.dw = d->dilation_w; p.depthwise = d->depthwise;
  return p;
}

}  // namespace dpk

using namespace dpk;

extern "C" int dpk_dgc_leaf_forward(const float* x, const float* loc, const float* scale, int64_t batch,
                                    int32_t in_channels, int32_t out_channels, int32_t hw, float* out, void* stream) {
  if (batch < 0 || in_channels <= 0 || out_channels <= 0 || hw <= 0) return set_error(DPK_E_ARG, "dgc_leaf: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !loc || !scale || !out) return set_error(DPK_E_ARG, "dgc_leaf: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_DGC, st);
  if ((in_channels == 1 || in_channels == 3) && env_int_dgc("DPK_DGC_LEAF_PX", 1)) {
    const int64_t bx = ceil_div(hw, 128);
    int64_t per = round_up(ceil_div(batch, std::max<int64_t>(1, ceil_div((int64_t)8 * sm_count(), bx))), 4);
    const int64_t row = (int64_t)std::max(in_channels, out_channels) * hw;
    per = std::max<int64_t>(4, std::min<int64_t>(per, ((int64_t(1) << 31) - 1) / row / 4 * 4));   // 32-bit offsets inside a slice
    dim3 grid((unsigned)bx, (unsigned)ceil_div(batch, per));
    if (in_channels == 1) dgc_leaf_fwd_px_kernel<1, 8><<<grid, 128, 0, st>>>(x, loc, scale, out, batch, out_channels, hw, per);
    else dgc_leaf_fwd_px_kernel<3, 8><<<grid, 128, 0, st>>>(x, loc, scale, out, batch, out_channels, hw, per);
    DPK_LAUNCH_CHECK("dgc_leaf_fwd_px_kernel");
    return DPK_OK;
  }
  dgc_leaf_fwd_kernel<<<plane_grid(batch * out_channels), 256, 0, st>>>(x, loc, scale, out, batch, in_channels,
                                                                                 out_channels, hw);
  DPK_LAUNCH_CHECK("dgc_leaf_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_dgc_leaf_backward(const float* x, const float* loc, const float* scale, const float* grad_out,
                                     int64_t batch, int32_t in_channels, int32_t out_channels, int32_t hw, float* grad_x,
                                     float* grad_loc, float* grad_scale, void* stream) {
  if (batch < 0 || in_channels <= 0 || out_channels <= 0 || hw <= 0) return set_error(DPK_E_ARG, "dgc_leaf_bwd: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !loc || !scale || !grad_out) return set_error(DPK_E_ARG, "dgc_leaf_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grad_x) {
    ProfScope prof(CAT_DGC_BWD, st);
    dgc_leaf_bwd_x_kernel<<<plane_grid(batch * in_channels), 256, 0, st>>>(x, loc, scale, grad_out, grad_x, batch,
                                                                                    in_channels, out_channels, hw);
    DPK_LAUNCH_CHECK("dgc_leaf_bwd_x_kernel");
  }
  if (grad_loc || grad_scale) {
    ProfScope prof(CAT_DGC_BWD, st);
    if (out_channels <= 16 && env_int_dgc("DPK_DGC_BWD_FAST", 1) != 0) {
      const int64_t bx = ceil_div((int64_t)in_channels * hw, 128);
      const int64_t per = std::max<int64_t>(16, ceil_div(batch, std::max<int64_t>(1, ceil_div((int64_t)16 * sm_count(), bx))));
      dim3 grid((unsigned)bx, (unsigned)ceil_div(batch, per));
      if (out_channels <= 8)
        dgc_leaf_bwd_param_px_kernel<8><<<grid, 128, 0, st>>>(x, loc, scale, grad_out, grad_loc, grad_scale, batch, in_channels, out_channels, hw, per);
      else
        dgc_leaf_bwd_param_px_kernel<16><<<grid, 128, 0, st>>>(x, loc, scale, grad_out, grad_loc, grad_scale, batch, in_channels, out_channels, hw, per);
      DPK_LAUNCH_CHECK("dgc_leaf_bwd_param_px_kernel");
      return DPK_OK;
    }
    const int64_t n = (int64_t)out_channels * in_channels * hw;
    const int64_t bx = ceil_div(n, 128);
    const int64_t per = slice_len(batch, bx, 32);
    dgc_leaf_bwd_param_kernel<<<dim3((unsigned)bx, (unsigned)ceil_div(batch, per)), 128, 0, st>>>(
        x, loc, scale, grad_out, grad_loc, grad_scale, batch, in_channels, out_channels, hw, per);
    DPK_LAUNCH_CHECK("dgc_leaf_bwd_param_kernel");
  }
  return DPK_OK;
}

extern "C" int dpk_dgc_product_forward(const dpk_dgc_product_desc* desc, const float* x, int64_t batch, float* out,
                                       void* stream) {
  if (!desc || batch < 0) return set_error(DPK_E_ARG, "dgc_product: bad arguments");
  if (batch == 0) return DPK_OK;
  if (!x || !out) return set_error(DPK_E_ARG, "dgc_product: null pointer");
  const ProdDesc d = to_prod(desc);
  if (d.C <= 0 || d.OC <= 0 || d.OH <= 0 || d.OW <= 0 || d.sh <= 0 || d.sw <= 0)
    return set_error(DPK_E_ARG, "dgc_product: bad descriptor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_DGC, st);
  if (d.depthwise && env_int_dgc("DPK_DGC_PROD_PX", 1) != 0) {
    const int64_t planes = batch * d.OC, bx = ceil_div((int64_t)d.OH * d.OW, 128);
    const int64_t slices = std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)16 * sm_count(), bx), ceil_div(planes, 8)));
    const int64_t per = ceil_div(planes, slices);
    dgc_product_px_kernel<false><<<dim3((unsigned)bx, (unsigned)ceil_div(planes, per)), 128, 0, st>>>(x, out, planes, per, d);
  } else {
    dgc_product_fwd_kernel<<<plane_grid(batch * d.OC), 256, 0, st>>>(x, out, batch, d);
  }
  DPK_LAUNCH_CHECK("dgc_product_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_dgc_product_backward(const dpk_dgc_product_desc* desc, const float* grad_out, int64_t batch,
                                        float* grad_x, void* stream) {
  if (!desc || batch < 0) return set_error(DPK_E_ARG, "dgc_product_bwd: bad arguments");
  if (batch == 0) return DPK_OK;
  if (!grad_out || !grad_x) return set_error(DPK_E_ARG, "dgc_product_bwd: null pointer");
  const ProdDesc d = to_prod(desc);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_DGC_BWD, st, d.depthwise ? 1 : 2);
  if (d.depthwise) {
    if (env_int_dgc("DPK_DGC_PROD_PX", 1) != 0) {
      const int64_t planes = batch * d.C, bx = ceil_div((int64_t)d.H * d.W, 128);
      const int64_t slices = std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)16 * sm_count(), bx), ceil_div(planes, 8)));
      const int64_t per = ceil_div(planes, slices);
      dgc_product_px_kernel<true><<<dim3((unsigned)bx, (unsigned)ceil_div(planes, per)), 128, 0, st>>>(grad_out, grad_x, planes, per, d);
    } else {
      dgc_product_bwd_depthwise_kernel<<<plane_grid(batch * d.C), 256, 0, st>>>(grad_out, grad_x, batch, d);
    }
  } else {
    DPK_CUDA_TRY(cudaMemsetAsync(grad_x, 0, (size_t)batch * d.C * d.H * d.W * sizeof(float), st));
    dgc_product_bwd_scatter_kernel<<<plane_grid(batch * d.OC), 256, 0, st>>>(grad_out, grad_x, batch, d);
  }
  DPK_LAUNCH_CHECK("dgc_product_bwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_dgc_sum_forward(const float* x, const float* weight, int64_t batch, int32_t in_channels,
                                   int32_t out_channels, int32_t hw, float* out, float* scratch, void* stream) {
  if (batch < 0 || in_channels <= 0 || out_channels <= 0 || hw <= 0) return set_error(DPK_E_ARG, "dgc_sum: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !weight || !out || !scratch) return set_error(DPK_E_ARG, "dgc_sum: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nw = (size_t)out_channels * in_channels * hw;
  float* wsoft = scratch;
  float* wlog = scratch + nw;
  ProfScope prof(CAT_DGC, st, 2);
  dgc_sum_prep_kernel<<<(out_channels * hw + 127) / 128, 128, 0, st>>>(weight, wsoft, wlog, out_channels, in_channels, hw);
  DPK_LAUNCH_CHECK("dgc_sum_prep_kernel");
  const int64_t bx = ceil_div(hw, 128);
  const Chunking oc = pick_chunk(out_channels);
  const int OC = oc.chunk > 8 ? 8 : oc.chunk;   // OC = 8 covers 10/16/32 in several passes without blowing up registers
  const size_t smem = (size_t)in_channels * OC * 128 * sizeof(float);
  if (smem <= 96 * 1024) {   // weights of a pixel tile resident in shared memory
    const int64_t per = round_up(ceil_div(batch, std::min<int64_t>(std::max<int64_t>(1, ceil_div((int64_t)env_int_dgc("DPK_DGC_CTAS_PER_SM", 8) * sm_count(), bx)),
                                                                     std::max<int64_t>(1, batch / 64))), kSumFastNB);
    dim3 grid((unsigned)bx, (unsigned)ceil_div(batch, per));
#define DPK_SUM_FAST(oc_)                                                                                           \
  {                                                                                                                 \
    auto kern = dgc_sum_fwd_smem_kernel<oc_>;                                                                       \
    if (smem > 48 * 1024) DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, 128, smem, st>>>(x, wsoft, wlog, out, batch, in_channels, out_channels, hw, per);                  \
  }
    switch (OC) {
      case 2: DPK_SUM_FAST(2) break;
      case 4: DPK_SUM_FAST(4) break;
      default: DPK_SUM_FAST(8) break;
    }
#undef DPK_SUM_FAST
  } else {
    const int64_t per = slice_len(batch, bx, 4 * kSumNB);
    dim3 grid((unsigned)bx, (unsigned)ceil_div(batch, per));
    switch (OC) {
      case 2: dgc_sum_fwd_kernel<2><<<grid, 128, 0, st>>>(x, wsoft, wlog, out, batch, in_channels, out_channels, hw, per); break;
      case 4: dgc_sum_fwd_kernel<4><<<grid, 128, 0, st>>>(x, wsoft, wlog, out, batch, in_channels, out_channels, hw, per); break;
      default: dgc_sum_fwd_kernel<8><<<grid, 128, 0, st>>>(x, wsoft, wlog, out, batch, in_channels, out_channels, hw, per); break;
    }
  }
  DPK_LAUNCH_CHECK("dgc_sum_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_dgc_prodsum_forward(const dpk_dgc_product_desc* desc, const float* x, const float* weight,
                                       int64_t batch, int32_t out_channels, float* out, float* scratch, void* stream) {
  if (!desc || batch < 0 || out_channels <= 0) return set_error(DPK_E_ARG, "dgc_prodsum: bad arguments");
  if (batch == 0) return DPK_OK;
  if (!x || !weight || !out || !scratch) return set_error(DPK_E_ARG, "dgc_prodsum: null pointer");
  const ProdDesc d = to_prod(desc);
  const int I = d.OC, hw = d.OH * d.OW;
  if (!d.depthwise || d.C != d.OC || !(I == 2 || I == 4 || I == 8 || I == 16 || I == 32) || d.sh <= 0 || d.sw <= 0)
    return set_error(DPK_E_ARG, "dgc_prodsum: only depthwise products with 2, 4, 8, 16 or 32 channels are fused");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nw = (size_t)out_channels * I * hw;
  float* wsoft = scratch;
  float* wlog = scratch + nw;
  ProfScope prof(CAT_DGC, st, 2);
  dgc_sum_prep_kernel<<<(out_channels * hw + 127) / 128, 128, 0, st>>>(weight, wsoft, wlog, out_channels, I, hw);
  DPK_LAUNCH_CHECK("dgc_sum_prep_kernel");
  const bool wg = I == 32;                               // weights through L1 instead of shared memory
  const int threads = (wg && hw <= 64) ? 64 : 128;
  const int64_t bx = ceil_div(hw, threads);
  const int OC = out_channels <= 2 ? 2 : (out_channels <= 4 ? 4 : 8);
  const size_t smem = wg ? 0 : (size_t)I * OC * 128 * sizeof(float);
  const int64_t per = round_up(ceil_div(batch, std::min<int64_t>(std::max<int64_t>(1, ceil_div((int64_t)env_int_dgc("DPK_DGC_CTAS_PER_SM", 8) * sm_count(), bx)),
                                                                   std::max<int64_t>(1, batch / 64))), 4);
  // the kernel addresses a slice with 32-bit element offsets
  const int64_t row_elems = std::max<int64_t>((int64_t)I * d.H * d.W, (int64_t)out_channels * hw);
  if (per * row_elems >= (int64_t(1) << 31)) return set_error(DPK_E_ARG, "dgc_prodsum: layer too large for one batch slice");
  dim3 grid((unsigned)bx, (unsigned)ceil_div(batch, per));
#define DPK_PS(ic_, oc_)                                                                                            \
  {                                                                                                                 \
    auto kern = dgc_prodsum_fwd_kernel<ic_, oc_>;                                                                   \
    if (smem > 48 * 1024) DPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, threads, smem, st>>>(x, wsoft, wlog, out, batch, out_channels, d, per);                            \
  }
#define DPK_PS_IC(ic_)                                    \
  switch (OC) {                                           \
    case 2: DPK_PS(ic_, 2) break;                         \
    case 4: DPK_PS(ic_, 4) break;                         \
    default: DPK_PS(ic_, 8) break;                        \
  }
  switch (I) {
    case 2: DPK_PS_IC(2) break;
    case 4: DPK_PS_IC(4) break;
    case 8: DPK_PS_IC(8) break;
    case 16: DPK_PS_IC(16) break;
    default: {
      auto kern = dgc_prodsum_fwd_kernel<32, 8, true>;
      kern<<<grid, threads, 0, st>>>(x, wsoft, wlog, out, batch, out_channels, d, per);
    } break;
  }
#undef DPK_PS_IC
#undef DPK_PS
  DPK_LAUNCH_CHECK("dgc_prodsum_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_dgc_sum_backward(const float* x, const float* weight, const float* out, const float* grad_out,
                                    int64_t batch, int32_t in_channels, int32_t out_channels, int32_t hw, float* grad_x,
                                    float* grad_weight, float* scratch, void* stream) {
  if (batch < 0 || in_channels <= 0 || out_channels <= 0 || hw <= 0) return set_error(DPK_E_ARG, "dgc_sum_bwd: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !weight || !out || !grad_out || !scratch) return set_error(DPK_E_ARG, "dgc_sum_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nw = (size_t)out_channels * in_channels * hw;
  float* wsoft = scratch;
  float* wlog = scratch + nw;
  float* nstat = scratch + 2 * nw;
  ProfScope prof(CAT_DGC_BWD, st, grad_weight ? 3 : 2);
  dgc_sum_prep_kernel<<<(out_channels * hw + 127) / 128, 128, 0, st>>>(weight, wsoft, wlog, out_channels, in_channels, hw);
  DPK_LAUNCH_CHECK("dgc_sum_prep_kernel");
  if (grad_weight) DPK_CUDA_TRY(cudaMemsetAsync(nstat, 0, nw * sizeof(float), st));
  const int64_t bx = ceil_div(hw, 128);
  const int64_t per = slice_len(batch, bx, 16);
  dim3 grid((unsigned)bx, (unsigned)ceil_div(batch, per));
  float* ns = grad_weight ? nstat : nullptr;
  const bool fast = env_int_dgc("DPK_DGC_BWD_FAST", 1) != 0;
  if (fast && out_channels <= 8 && in_channels <= 8) {
    // more, shorter slices than the generic kernel: 16 samples' worth of loads per thread keep HBM busy
    const int64_t per2 = std::max<int64_t>(16, ceil_div(batch, std::max<int64_t>(1, ceil_div((int64_t)12 * sm_count(), bx))));
    dim3 grid2((unsigned)bx, (unsigned)ceil_div(batch, per2));
    if (out_channels <= 4 && in_channels <= 4)
      dgc_sum_bwd_fast_kernel<4, 4, false><<<grid2, 128, 0, st>>>(x, wsoft, out, grad_out, grad_x, ns, batch, in_channels, out_channels, hw, per2, ProdDesc{});
    else
      dgc_sum_bwd_fast_kernel<8, 8, false><<<grid2, 128, 0, st>>>(x, wsoft, out, grad_out, grad_x, ns, batch, in_channels, out_channels, hw, per2, ProdDesc{});
  } else if (fast) {
    const int64_t nz = ceil_div(in_channels, 8);
    const int64_t per2 = std::max<int64_t>(16, ceil_div(batch, std::max<int64_t>(1, ceil_div((int64_t)12 * sm_count(), bx * nz))));
    dgc_sum_bwd_chunk_kernel<8, 8><<<dim3((unsigned)bx, (unsigned)ceil_div(batch, per2), (unsigned)nz), 128, 0, st>>>(
        x, wsoft, out, grad_out, grad_x, ns, batch, in_channels, out_channels, hw, per2);
  } else if (out_channels <= 4 && in_channels <= 4)
    dgc_sum_bwd_kernel<4, 4><<<grid, 128, 0, st>>>(x, wsoft, out, grad_out, grad_x, ns, batch, in_channels, out_channels, hw, per);
  else if (out_channels <= 4)
    dgc_sum_bwd_kernel<4, 8><<<grid, 128, 0, st>>>(x, wsoft, out, grad_out, grad_x, ns, batch, in_channels, out_channels, hw, per);
  else
    dgc_sum_bwd_kernel<8, 8><<<grid, 128, 0, st>>>(x, wsoft, out, grad_out, grad_x, ns, batch, in_channels, out_channels, hw, per);
  DPK_LAUNCH_CHECK("dgc_sum_bwd_kernel");
  if (grad_weight) {
    dgc_sum_finalize_kernel<<<(out_channels * hw + 127) / 128, 128, 0, st>>>(wsoft, nstat, grad_weight, out_channels,
                                                                              in_channels, hw);
    DPK_LAUNCH_CHECK("dgc_sum_finalize_kernel");
  }
  return DPK_OK;
}

extern "C" int dpk_dgc_prodsum_backward(const dpk_dgc_product_desc* desc, const float* x, const float* weight, const float* out,
                                        const float* grad_out, int64_t batch, int32_t out_channels, float* grad_prod,
                                        float* grad_weight, float* scratch, void* stream) {
  if (!desc || batch < 0 || out_channels <= 0) return set_error(DPK_E_ARG, "dgc_prodsum_bwd: bad arguments");
  if (batch == 0) return DPK_OK;
  if (!x || !weight || !out || !grad_out || !scratch) return set_error(DPK_E_ARG, "dgc_prodsum_bwd: null pointer");
  const ProdDesc d = to_prod(desc);
  const int I = d.OC, hw = d.OH * d.OW;
  if (!d.depthwise || d.C != d.OC || I > 8 || out_channels > 8 || d.sh <= 0 || d.sw <= 0)
    return set_error(DPK_E_ARG, "dgc_prodsum_bwd: only depthwise products with at most 8 channels in and out");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nw = (size_t)out_channels * I * hw;
  float* wsoft = scratch;
  float* wlog = scratch + nw;
  float* nstat = scratch + 2 * nw;
  ProfScope prof(CAT_DGC_BWD, st, grad_weight ? 3 : 2);
  dgc_sum_prep_kernel<<<(out_channels * hw + 127) / 128, 128, 0, st>>>(weight, wsoft, wlog, out_channels, I, hw);
  DPK_LAUNCH_CHECK("dgc_sum_prep_kernel");
  if (grad_weight) DPK_CUDA_TRY(cudaMemsetAsync(nstat, 0, nw * sizeof(float), st));
  const int64_t bx = ceil_div(hw, 128);
  const int64_t per = std::max<int64_t>(16, ceil_div(batch, std::max<int64_t>(1, ceil_div((int64_t)12 * sm_count(), bx))));
  dim3 grid((unsigned)bx, (unsigned)ceil_div(batch, per));
  float* ns = grad_weight ? nstat : nullptr;
  if (out_channels <= 4 && I <= 4)
    dgc_sum_bwd_fast_kernel<4, 4, true><<<grid, 128, 0, st>>>(x, wsoft, out, grad_out, grad_prod, ns, batch, I, out_channels, hw, per, d);
  else
    dgc_sum_bwd_fast_kernel<8, 8, true><<<grid, 128, 0, st>>>(x, wsoft, out, grad_out, grad_prod, ns, batch, I, out_channels, hw, per, d);
  DPK_LAUNCH_CHECK("dgc_sum_bwd_fast_kernel<gather>");
  if (grad_weight) {
    dgc_sum_finalize_kernel<<<(out_channels * hw + 127) / 128, 128, 0, st>>>(wsoft, nstat, grad_weight, out_channels, I, hw);
    DPK_LAUNCH_CHECK("dgc_sum_finalize_kernel");
  }
  return DPK_OK;
}

extern "C" int dpk_dgc_root_forward(const float* x, const float* weight, int64_t batch, int64_t features,
                                    int32_t out_classes, float* out, float* scratch, void* stream) {
  if (batch < 0 || features <= 0 || out_classes <= 0) return set_error(DPK_E_ARG, "dgc_root: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !weight || !out || !scratch) return set_error(DPK_E_ARG, "dgc_root: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_DGC, st, 2);
  dgc_root_prep_kernel<<<out_classes, 256, 0, st>>>(weight, scratch, out_classes, features);
  DPK_LAUNCH_CHECK("dgc_root_prep_kernel");
  dgc_root_fwd_kernel<<<(unsigned)batch, 256, 0, st>>>(x, scratch, out, features, out_classes);
  DPK_LAUNCH_CHECK("dgc_root_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_dgc_root_backward(const float* x, const float* weight, const float* out, const float* grad_out,
                                     int64_t batch, int64_t features, int32_t out_classes, float* grad_x,
                                     float* grad_weight, float* scratch, void* stream) {
  if (batch < 0 || features <= 0 || out_classes <= 0) return set_error(DPK_E_ARG, "dgc_root_bwd: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !weight || !out || !grad_out || !scratch) return set_error(DPK_E_ARG, "dgc_root_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nw = (size_t)out_classes * features;
  float* wlog = scratch;
  float* nstat = scratch + nw;
  ProfScope prof(CAT_DGC_BWD, st, grad_weight ? 3 : 2);
  dgc_root_prep_kernel<<<out_classes, 256, 0, st>>>(weight, wlog, out_classes, features);
  DPK_LAUNCH_CHECK("dgc_root_prep_kernel");
  if (grad_weight) DPK_CUDA_TRY(cudaMemsetAsync(nstat, 0, nw * sizeof(float), st));
  const int64_t bx = ceil_div(features, 128);
  const int64_t per = slice_len(batch, bx, 16);
  dgc_root_bwd_kernel<<<dim3((unsigned)bx, (unsigned)ceil_div(batch, per)), 128, 0, st>>>(
      x, wlog, out, grad_out, grad_x, grad_weight ? nstat : nullptr, batch, features, out_classes, per);
  DPK_LAUNCH_CHECK("dgc_root_bwd_kernel");
  if (grad_weight) {
    dgc_root_finalize_kernel<<<out_classes, 256, 0, st>>>(wlog, nstat, grad_weight, features);
    DPK_LAUNCH_CHECK("dgc_root_finalize_kernel");
  }
  return DPK_OK;
}
