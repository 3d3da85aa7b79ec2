// flows.cu -- normalizing-flow bijector kernels: affine coupling transform + log-det, batch-norm
// bijector, dequantize/logit preprocessing, Gaussian prior.
//
// Reference (deeprob-kit, paths relative to /root/reference):
//   CouplingLayer1d.apply_backward/forward   deeprob/flows/layers/coupling.py:72-104
//   CouplingLayer2d.apply_backward/forward   deeprob/flows/layers/coupling.py:179-272
//   AutoregressiveLayer.apply_backward       deeprob/flows/layers/autoregressive.py:72-79
//   ScaledTanh                               deeprob/torch/utils.py:52-70
//   BatchNormLayer1d/2d                      deeprob/flows/utils.py:118-153, 183-221
//   DequantizeLayer / LogitLayer             deeprob/flows/utils.py:244-254, 276-294
//   prior + sum                              deeprob/flows/models/base.py:139-143
// The reference runs ~8 elementwise ATen kernels + a reduction per coupling; here each bijector is one
// pass over the activations: transform, per-sample log-det reduction and accumulation into the
// running log-det vector in the same kernel.  The conditioner networks (MLP / conv) stay library calls.
#include <algorithm>

#include <cstdlib>

#include "common.cuh"

namespace dpk {

static inline int flow_env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  return s;
}

// =================================================================================================
// Affine coupling
// =================================================================================================
struct CouplingArgs {
  const float* x; int64_t x_stride;    // (B, N) input rows
  const float* z; int64_t z_stride;    // conditioner output rows: t = z[0:N], raw s = z[N:2N] (affine only)
  const float* inv_mask;               // (N) 1 where the element is transformed, NULL = all
  const float* w;                      // ScaledTanh weights, index (e / w_inner) % w_count
  float* out; int64_t out_stride;
  float* ldj;                          // (B) log-det accumulator (+=), may be NULL
  int64_t B;
  int N, w_count, w_inner, affine, direction;  // direction 0: u = (x - t) * exp(-s), ldj -= sum s ; 1: x = u * exp(s) + t, ldj += sum s
  const int* zmap; int z_half;         // compact conditioner output (forward only): column of element e, offset of the s half
  const float* post_scale; const float* post_shift; float post_ldj;  // fused per-feature affine after the coupling (eval batch-norm)
  float* side; int64_t side_stride;    // optional compact copy of the transformed elements: side[b][zmap[e]] = out[b][e]
  int vec4;
};

// One element of the coupling: masked-out elements (m == 0) pass through without touching z, so a conditioner that
// only produced the live columns (compact z: t = z[zmap[e]], raw s = z[z_half + zmap[e]]) can be consumed directly.
__device__ __forceinline__ float coupling_elem(const CouplingArgs& a, const float* zr, int e, float m, float xv, float& acc) {
  if (m == 0.f) return xv;
  const int zi = a.zmap ? __ldg(a.zmap + e) : e;
  const float t = m * zr[zi];
  float s = 0.f;
  if (a.affine) s = m * __ldg(a.w + (e / a.w_inner) % a.w_count) * tanhf(zr[a.z_half + zi]);
  acc += s;
  return a.direction == 0 ? (xv - t) * expf(-s) : xv * expf(s) + t;
}

__global__ void __launch_bounds__(256) coupling_fwd_kernel(const CouplingArgs a) {
  __shared__ float red[8];
  for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float* xr = a.x + b * a.x_stride;
    const float* zr = a.z + b * a.z_stride;
    float* orow = a.out + b * a.out_stride;
    float acc = 0.f;
    if (a.vec4) {
      // 16-byte loads/stores of x, mask, out (and of the fused batch-norm affine); z through scalar loads
      float* sr = a.side ? a.side + b * a.side_stride : nullptr;
      for (int e = threadIdx.x * 4; e < a.N; e += 1024) {
        const float4 xv = __ldcs(reinterpret_cast<const float4*>(xr + e));
        float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
        if (a.inv_mask) m = __ldg(reinterpret_cast<const float4*>(a.inv_mask + e));
        float4 r;
        r.x = coupling_elem(a, zr, e, m.x, xv.x, acc);
        r.y = coupling_elem(a, zr, e + 1, m.y, xv.y, acc);
        r.z = coupling_elem(a, zr, e + 2, m.z, xv.z, acc);
        r.w = coupling_elem(a, zr, e + 3, m.w, xv.w, acc);
        if (a.post_scale) {
          const float4 pa = __ldg(reinterpret_cast<const float4*>(a.post_scale + e));
          const float4 pc = __ldg(reinterpret_cast<const float4*>(a.post_shift + e));
          r.x = fmaf(r.x, pa.x, pc.x); r.y = fmaf(r.y, pa.y, pc.y);
          r.z = fmaf(r.z, pa.z, pc.z); r.w = fmaf(r.w, pa.w, pc.w);
        }
        *reinterpret_cast<float4*>(orow + e) = r;
        if (sr) {
          if (m.x != 0.f) sr[__ldg(a.zmap + e)] = r.x;
          if (m.y != 0.f) sr[__ldg(a.zmap + e + 1)] = r.y;
          if (m.z != 0.f) sr[__ldg(a.zmap + e + 2)] = r.z;
          if (m.w != 0.f) sr[__ldg(a.zmap + e + 3)] = r.w;
        }
      }
    } else {
      for (int e = threadIdx.x; e < a.N; e += 256) {
        const float m = a.inv_mask ? __ldg(a.inv_mask + e) : 1.f;
        float r = coupling_elem(a, zr, e, m, xr[e], acc);
        if (a.post_scale) r = fmaf(r, __ldg(a.post_scale + e), __ldg(a.post_shift + e));
        orow[e] = r;
        if (a.side && m != 0.f) a.side[b * a.side_stride + __ldg(a.zmap + e)] = r;
      }
    }
    if (a.ldj) {
      float tot = 0.f;
      if (a.affine) tot = block_sum_256(acc, red);
      if (threadIdx.x == 0 && (a.affine || a.post_ldj != 0.f)) a.ldj[b] += ((a.direction == 0) ? -tot : tot) + a.post_ldj;
    }
  }
}

struct CouplingBwdArgs {
  CouplingArgs f;                  // forward arguments (x, z, mask, w, geometry); out/ldj unused
  const float* gout; int64_t gout_stride;
  const float* gldj;               // (B) gradient w.r.t. the log-det, may be NULL
  float* gx; int64_t gx_stride;    // overwritten, may be NULL
  float* gz; int64_t gz_stride;    // overwritten (both halves)
  float* gw;                       // (w_count) accumulated with atomics, may be NULL
};

__global__ void __launch_bounds__(256) coupling_bwd_kernel(const CouplingBwdArgs a) {
  extern __shared__ float gw_s[];  // [w_count]
  const CouplingArgs& f = a.f;
  for (int i = threadIdx.x; i < f.w_count; i += 256) gw_s[i] = 0.f;
  __syncthreads();
  // gradient of the ScaledTanh weights: a thread keeps the partial sum of the weight index it is currently on in a
  // register and touches shared memory only when that index changes (the 1-D couplings have ONE weight: a
  // shared-memory atomic per element serialised the whole CTA on it -- 1.47 ms per coupling at batch 16384)
  int cur_wi = -1;
  float cur_acc = 0.f;
  for (int64_t b = blockIdx.x; b < f.B; b += gridDim.x) {
    const float* xr = f.x + b * f.x_stride;
    const float* zr = f.z + b * f.z_stride;
    const float* gr = a.gout + b * a.gout_stride;
    const float gl = a.gldj ? a.gldj[b] : 0.f;
    for (int e = threadIdx.x; e < f.N; e += 256) {
      const float m = f.inv_mask ? __ldg(f.inv_mask + e) : 1.f;
      const float t = m * zr[e];
      const float gu = gr[e], xv = xr[e];
      if (!f.affine) {
        if (a.gx) a.gx[b * a.gx_stride + e] = gu;
        a.gz[b * a.gz_stride + e] = (f.direction == 0 ? -m : m) * gu;
        continue;
      }
      const int wi = (e / f.w_inner) % f.w_count;
      const float wv = __ldg(f.w + wi);
      const float th = tanhf(zr[f.N + e]);
      const float s = m * wv * th;
      float gxv, gt, gs;
      if (f.direction == 0) {
        const float es = expf(-s);
        gxv = gu * es;
        gt = -m * gxv;
        gs = -gu * (xv - t) * es - gl;   // d u/d s = -u ; d ildj/d s = -1
      } else {
        const float es = expf(s);
        gxv = gu * es;
        gt = m * gu;
        gs = gu * xv * es + gl;
      }
      if (a.gx) a.gx[b * a.gx_stride + e] = gxv;
      a.gz[b * a.gz_stride + e] = gt;
      a.gz[b * a.gz_stride + f.N + e] = gs * m * wv * (1.f - th * th);
      if (a.gw) {
        if (wi != cur_wi) {
          if (cur_wi >= 0 && cur_acc != 0.f) atomicAdd(gw_s + cur_wi, cur_acc);
          cur_wi = wi; cur_acc = 0.f;
        }
        cur_acc += gs * m * th;
      }
    }
  }
  if (a.gw && cur_wi >= 0 && cur_acc != 0.f) atomicAdd(gw_s + cur_wi, cur_acc);
  __syncthreads();
  if (a.gw)
    for (int i = threadIdx.x; i < f.w_count; i += 256)
      if (gw_s[i] != 0.f) atomicAdd(a.gw + i, gw_s[i]);
}

// =================================================================================================
// Per-feature statistics / affine maps for the batch-norm bijector.  Geometry: element (b, f, i) at
// b*F*I + f*I + i  (1d: I = 1; 2d: I = H*W).
// =================================================================================================
// sum_out[f] += sum x ; with `center`: sum_out[f] += sum (x - center[f])^2 ; with `other` (and center):
// dot_out[f] += sum other * (x - center[f])
__global__ void feature_reduce_kernel(const float* __restrict__ x, const float* __restrict__ center,
                                      const float* __restrict__ other, float* __restrict__ sum_out,
                                      float* __restrict__ dot_out, int64_t B, int F, int I, int mode,
                                      int64_t per_slice) {
  // mode 0: sum x ; mode 1: sum (x-c)^2 ; mode 2: sum other and sum other*(x-c)
  __shared__ float red[8];
  const int64_t b0 = blockIdx.y * per_slice, b1 = min((long long)B, (long long)(b0 + per_slice));
  if (I == 1) {  // thread = feature, coalesced across features
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float c = center ? center[f] : 0.f;
    float s = 0.f, d = 0.f;
#pragma unroll 8
    for (int64_t b = b0; b < b1; ++b) {
      const float xv = x[b * F + f];
      if (mode == 0) s += xv;
      else if (mode == 1) { const float t = xv - c; s = fmaf(t, t, s); }
      else { const float o = other[b * F + f]; s += o; d = fmaf(o, xv - c, d); }
    }
    atomicAdd(sum_out + f, s);
    if (mode == 2) atomicAdd(dot_out + f, d);
  } else {       // CTA = feature, threads over (b, i)
    const int f = blockIdx.x;
    const float c = center ? center[f] : 0.f;
    float s = 0.f, d = 0.f;
    const int64_t n = (b1 - b0) * I;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      const int64_t b = b0 + j / I;
      const int i = (int)(j % I);
      const size_t at = ((size_t)b * F + f) * I + i;
      const float xv = x[at];
      if (mode == 0) s += xv;
      else if (mode == 1) { const float t = xv - c; s = fmaf(t, t, s); }
      else { const float o = other[at]; s += o; d = fmaf(o, xv - c, d); }
    }
    const float st = block_sum_256(s, red);
    const float dt = (mode == 2) ? block_sum_256(d, red) : 0.f;
    if (threadIdx.x == 0) { atomicAdd(sum_out + f, st); if (mode == 2) atomicAdd(dot_out + f, dt); }
  }
}

// out = x * a[f] + c[f]  (+ k[f] * (y - mu[f]) when y is given)
__global__ void feature_affine_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                      const float* __restrict__ c, const float* __restrict__ y,
                                      const float* __restrict__ k, const float* __restrict__ mu,
                                      float* __restrict__ out, int64_t total, int F, int I) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)((idx / I) % F);
    float v = fmaf(x[idx], __ldg(a + f), __ldg(c + f));
    if (y) v = fmaf(__ldg(k + f), y[idx] - __ldg(mu + f), v);
    out[idx] = v;
  }
}

// the same for (B, F) tensors with F % 4 == 0: 16-byte accesses, no division per element (the per-feature vectors are
// read as float4 from L1/L2; a thread's feature offset advances by a constant)
__global__ void __launch_bounds__(256) feature_affine_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ a,
                                                                 const float4* __restrict__ c, const float4* __restrict__ y,
                                                                 const float4* __restrict__ k, const float4* __restrict__ mu,
                                                                 float4* __restrict__ out, int64_t total4, int F4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int f = (int)(idx % F4);
  const int fstep = (int)(stride % F4);
  for (; idx < total4; idx += stride) {
    const float4 xv = __ldcs(x + idx), av = __ldg(a + f), cv = __ldg(c + f);
    float4 v = make_float4(fmaf(xv.x, av.x, cv.x), fmaf(xv.y, av.y, cv.y), fmaf(xv.z, av.z, cv.z), fmaf(xv.w, av.w, cv.w));
    if (y) {
      const float4 yv = __ldcs(y + idx), kv = __ldg(k + f), mv = __ldg(mu + f);
      v.x = fmaf(kv.x, yv.x - mv.x, v.x); v.y = fmaf(kv.y, yv.y - mv.y, v.y);
      v.z = fmaf(kv.z, yv.z - mv.z, v.z); v.w = fmaf(kv.w, yv.w - mv.w, v.w);
    }
    __stcs(out + idx, v);
    f += fstep;
    if (f >= F4) f -= F4;
  }
}

// =================================================================================================
// Dequantize + logit preprocessing (either may be disabled) and the Gaussian prior
// =================================================================================================
// u = logit(alpha + (1-2alpha) * q),  q = (x*(bins-1) + noise)/bins  (bins = 0: q = x ; alpha < 0: u = q)
// ildj[b] += -(sum log y + log(1-y))   (the constant parts are added on the host side)
__global__ void __launch_bounds__(256) preprocess_fwd_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                             float bins, float alpha, float* __restrict__ out,
                                                             float* __restrict__ ildj, int64_t B, int N) {
  __shared__ float red[8];
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    float acc = 0.f;
    for (int e = threadIdx.x; e < N; e += 256) {
      float q = x[b * N + e];
      if (bins > 0.f) q = (q * (bins - 1.f) + noise[b * N + e]) / bins;
      if (alpha >= 0.f) {
        const float y = alpha + (1.f - 2.f * alpha) * q;
        const float ly = logf(y), ry = logf(1.f - y);
        q = ly - ry;
        acc += ly + ry;
      }
      out[b * N + e] = q;
    }
    if (alpha >= 0.f && ildj) {
      const float tot = block_sum_256(acc, red);
      if (threadIdx.x == 0) ildj[b] -= tot;
    }
  }
}

__global__ void preprocess_bwd_kernel(const float* __restrict__ x, const float* __restrict__ noise, float bins,
                                      float alpha, const float* __restrict__ gout, const float* __restrict__ gildj,
                                      float* __restrict__ gx, int64_t B, int N) {
  const int64_t total = B * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / N;
    float q = x[idx], dq = 1.f;
    if (bins > 0.f) { q = (q * (bins - 1.f) + noise[idx]) / bins; dq = (bins - 1.f) / bins; }
    float g = gout[idx];
    if (alpha >= 0.f) {
      const float k = 1.f - 2.f * alpha;
      const float y = alpha + k * q;
      const float iy = 1.f / y, iz = 1.f / (1.f - y);
      g = g * k * (iy + iz) - (gildj ? gildj[b] : 0.f) * k * (iy - iz);
    }
    gx[idx] = g * dq;
  }
}

// out[b] = sum_e logN(z[b,e]; loc[e], scale[e]) + ildj[b]      (loc/scale NULL = standard normal)
__global__ void __launch_bounds__(256) normal_prior_fwd_kernel(const float* __restrict__ z, const float* __restrict__ loc,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ ildj, float* __restrict__ out,
                                                               int64_t B, int N) {
  __shared__ float red[8];
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    float acc = 0.f;
    for (int e = threadIdx.x; e < N; e += 256) {
      const float sg = scale ? __ldg(scale + e) : 1.f;
      const float t = (z[b * N + e] - (loc ? __ldg(loc + e) : 0.f)) / sg;
      acc += -0.5f * t * t - (scale ? logf(sg) : 0.f) - kLogSqrt2Pi;
    }
    const float tot = block_sum_256(acc, red);
    if (threadIdx.x == 0) out[b] = tot + (ildj ? ildj[b] : 0.f);
  }
}

// N % 4 == 0 and 16-byte aligned rows: one warp per row, 16-byte loads
__global__ void __launch_bounds__(256) normal_prior_fwd_vec_kernel(const float* __restrict__ z, const float* __restrict__ loc,
                                                                   const float* __restrict__ scale,
                                                                   const float* __restrict__ ildj, float* __restrict__ out,
                                                                   int64_t B, int N) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
  for (int64_t b = w0; b < B; b += nw) {
    const float* zr = z + b * N;
    float acc = 0.f;
#pragma unroll 4
    for (int e = lane * 4; e < N; e += 128) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(zr + e));
      float4 sg = make_float4(1.f, 1.f, 1.f, 1.f), lc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (scale) sg = __ldg(reinterpret_cast<const float4*>(scale + e));
      if (loc) lc = __ldg(reinterpret_cast<const float4*>(loc + e));
      const float t0 = (v.x - lc.x) / sg.x, t1 = (v.y - lc.y) / sg.y, t2 = (v.z - lc.z) / sg.z, t3 = (v.w - lc.w) / sg.w;
      acc += -0.5f * t0 * t0 - (scale ? logf(sg.x) : 0.f) - kLogSqrt2Pi;
      acc += -0.5f * t1 * t1 - (scale ? logf(sg.y) : 0.f) - kLogSqrt2Pi;
      acc += -0.5f * t2 * t2 - (scale ? logf(sg.z) : 0.f) - kLogSqrt2Pi;
      acc += -0.5f * t3 * t3 - (scale ? logf(sg.w) : 0.f) - kLogSqrt2Pi;
    }
    const float tot = warp_sum(acc);
    if (lane == 0) out[b] = tot + (ildj ? ildj[b] : 0.f);
  }
}

__global__ void normal_prior_bwd_kernel(const float* __restrict__ z, const float* __restrict__ loc,
                                        const float* __restrict__ scale, const float* __restrict__ gout,
                                        float* __restrict__ gz, int64_t B, int N) {
  const int64_t total = B * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % N);
    const float sg = scale ? __ldg(scale + e) : 1.f;
    gz[idx] = -gout[idx / N] * (z[idx] - (loc ? __ldg(loc + e) : 0.f)) / (sg * sg);
  }
}

static int sample_grid(int64_t B) { return (int)std::min<int64_t>(B, (int64_t)sm_count() * 16); }
static int flat_grid(int64_t total) { return (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 32); }

static int fill_coupling(CouplingArgs* a, const dpk_coupling_desc* d, const float* x, const float* z) {
  if (!d || d->batch < 0 || d->features <= 0 || d->w_count <= 0 || d->w_inner <= 0)
    return set_error(DPK_E_ARG, "coupling: bad descriptor");
  if (!x || !z || (d->affine && !d->scale_weight)) return set_error(DPK_E_ARG, "coupling: null pointer");
  a->x = x; a->x_stride = d->x_stride; a->z = z; a->z_stride = d->z_stride; a->inv_mask = d->inv_mask;
  a->w = d->scale_weight; a->out = nullptr; a->out_stride = 0; a->ldj = nullptr; a->B = d->batch; a->N = d->features;
  a->w_count = d->w_count; a->w_inner = d->w_inner; a->affine = d->affine; a->direction = d->direction;
  a->zmap = nullptr; a->z_half = d->features; a->post_scale = a->post_shift = nullptr; a->post_ldj = 0.f; a->vec4 = 0; a->side = nullptr; a->side_stride = 0;
  return DPK_OK;
}

}  // namespace dpk

using namespace dpk;

static int launch_coupling_fwd(CouplingArgs& a, float* out, int64_t out_stride, float* log_det, void* stream) {
  if (!out) return set_error(DPK_E_ARG, "coupling: null output");
  a.out = out; a.out_stride = out_stride; a.ldj = log_det;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  a.vec4 = a.N % 4 == 0 && a.x_stride % 4 == 0 && out_stride % 4 == 0 && al16(a.x) && al16(out) && al16(a.inv_mask) &&
           al16(a.post_scale) && al16(a.post_shift);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_FLOW, st);
  coupling_fwd_kernel<<<sample_grid(a.B), 256, 0, st>>>(a);
  DPK_LAUNCH_CHECK("coupling_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_coupling_forward(const dpk_coupling_desc* desc, const float* x, const float* z, float* out,
                                    int64_t out_stride, float* log_det, void* stream) {
  CouplingArgs a;
  int rc = fill_coupling(&a, desc, x, z);
  if (rc) return rc;
  if (a.B == 0) return DPK_OK;
  return launch_coupling_fwd(a, out, out_stride, log_det, stream);
}

extern "C" int dpk_coupling_forward_compact(const dpk_coupling_desc* desc, const float* x, const float* z,
                                            const int32_t* z_index, int32_t z_half, const float* post_scale,
                                            const float* post_shift, float post_log_det, float* out,
                                            int64_t out_stride, float* live_out, float* log_det, void* stream) {
  CouplingArgs a;
  int rc = fill_coupling(&a, desc, x, z);
  if (rc) return rc;
  if ((post_scale == nullptr) != (post_shift == nullptr)) return set_error(DPK_E_ARG, "coupling: post affine needs scale and shift");
  if (z_index && (!desc->inv_mask || z_half <= 0)) return set_error(DPK_E_ARG, "coupling: compact z needs inv_mask and z_half");
  if (post_log_det != 0.f && !log_det) return set_error(DPK_E_ARG, "coupling: post log-det without accumulator");
  if (live_out && !z_index) return set_error(DPK_E_ARG, "coupling: live_out needs z_index");
  if (a.B == 0) return DPK_OK;
  a.zmap = z_index;
  if (z_index) a.z_half = z_half;
  a.post_scale = post_scale; a.post_shift = post_shift; a.post_ldj = post_log_det;
  a.side = live_out; a.side_stride = z_half;
  return launch_coupling_fwd(a, out, out_stride, log_det, stream);
}

extern "C" int dpk_coupling_backward(const dpk_coupling_desc* desc, const float* x, const float* z,
                                     const float* grad_out, int64_t grad_out_stride, const float* grad_log_det,
                                     float* grad_x, int64_t grad_x_stride, float* grad_z, int64_t grad_z_stride,
                                     float* grad_scale_weight, void* stream) {
  CouplingBwdArgs a;
  int rc = fill_coupling(&a.f, desc, x, z);
  if (rc) return rc;
  if (a.f.B == 0) return DPK_OK;
  if (!grad_out || !grad_z) return set_error(DPK_E_ARG, "coupling_bwd: null pointer");
  a.gout = grad_out; a.gout_stride = grad_out_stride; a.gldj = grad_log_det; a.gx = grad_x; a.gx_stride = grad_x_stride;
  a.gz = grad_z; a.gz_stride = grad_z_stride; a.gw = a.f.affine ? grad_scale_weight : nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_FLOW_BWD, st);
  coupling_bwd_kernel<<<sample_grid(a.f.B), 256, (size_t)a.f.w_count * sizeof(float), st>>>(a);
  DPK_LAUNCH_CHECK("coupling_bwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_feature_reduce(const float* x, const float* center, const float* other, float* sum_out,
                                  float* dot_out, int64_t batch, int32_t features, int32_t inner, int32_t mode,
                                  void* stream) {
  if (batch < 0 || features <= 0 || inner <= 0 || mode < 0 || mode > 2) return set_error(DPK_E_ARG, "feature_reduce: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !sum_out || (mode >= 1 && !center) || (mode == 2 && (!other || !dot_out)))
    return set_error(DPK_E_ARG, "feature_reduce: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t bx = (inner == 1) ? ceil_div(features, 128) : features;
  int64_t slices = std::max<int64_t>(1, ceil_div((int64_t)4 * sm_count(), bx));
  slices = std::min<int64_t>(slices, std::max<int64_t>(1, batch / 8));
  const int64_t per = ceil_div(batch, slices);
  ProfScope prof(mode == 2 ? CAT_FLOW_BWD : CAT_FLOW, st);
  feature_reduce_kernel<<<dim3((unsigned)bx, (unsigned)ceil_div(batch, per)), inner == 1 ? 128 : 256, 0, st>>>(
      x, center, other, sum_out, dot_out, batch, features, inner, mode, per);
  DPK_LAUNCH_CHECK("feature_reduce_kernel");
  return DPK_OK;
}

extern "C" int dpk_feature_affine(const float* x, const float* a, const float* c, const float* y, const float* k,
                                  const float* mu, float* out, int64_t batch, int32_t features, int32_t inner,
                                  void* stream) {
  if (batch < 0 || features <= 0 || inner <= 0) return set_error(DPK_E_ARG, "feature_affine: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !a || !c || !out || (y && (!k || !mu))) return set_error(DPK_E_ARG, "feature_affine: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = batch * features * inner;
  ProfScope prof(y ? CAT_FLOW_BWD : CAT_FLOW, st);
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (inner == 1 && features % 4 == 0 && al16(x) && al16(a) && al16(c) && al16(y) && al16(k) && al16(mu) && al16(out)) {
    feature_affine_vec_kernel<<<flat_grid(total / 4), 256, 0, st>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(c),
        reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(k), reinterpret_cast<const float4*>(mu),
        reinterpret_cast<float4*>(out), total / 4, features / 4);
  } else {
    feature_affine_kernel<<<flat_grid(total), 256, 0, st>>>(x, a, c, y, k, mu, out, total, features, inner);
  }
  DPK_LAUNCH_CHECK("feature_affine_kernel");
  return DPK_OK;
}

extern "C" int dpk_flow_preprocess_forward(const float* x, const float* noise, float bins, float alpha, float* out,
                                           float* inv_log_det, int64_t batch, int32_t features, void* stream) {
  if (batch < 0 || features <= 0) return set_error(DPK_E_ARG, "preprocess: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !out || (bins > 0.f && !noise)) return set_error(DPK_E_ARG, "preprocess: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_FLOW, st);
  preprocess_fwd_kernel<<<sample_grid(batch), 256, 0, st>>>(x, noise, bins, alpha, out, inv_log_det, batch, features);
  DPK_LAUNCH_CHECK("preprocess_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_flow_preprocess_backward(const float* x, const float* noise, float bins, float alpha,
                                            const float* grad_out, const float* grad_inv_log_det, float* grad_x,
                                            int64_t batch, int32_t features, void* stream) {
  if (batch < 0 || features <= 0) return set_error(DPK_E_ARG, "preprocess_bwd: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !grad_out || !grad_x || (bins > 0.f && !noise)) return set_error(DPK_E_ARG, "preprocess_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_FLOW_BWD, st);
  preprocess_bwd_kernel<<<flat_grid(batch * features), 256, 0, st>>>(x, noise, bins, alpha, grad_out, grad_inv_log_det,
                                                                     grad_x, batch, features);
  DPK_LAUNCH_CHECK("preprocess_bwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_normal_prior_forward(const float* z, const float* loc, const float* scale, const float* inv_log_det,
                                        float* out, int64_t batch, int32_t features, void* stream) {
  if (batch < 0 || features <= 0) return set_error(DPK_E_ARG, "normal_prior: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!z || !out) return set_error(DPK_E_ARG, "normal_prior: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_FLOW, st);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (flow_env_int("DPK_PRIOR_WARP", 1) && features % 4 == 0 && features >= 128 && al16(z) && al16(loc) && al16(scale)) {
    normal_prior_fwd_vec_kernel<<<(unsigned)std::min<int64_t>(ceil_div(batch, 8), (int64_t)sm_count() * 16), 256, 0, st>>>(
        z, loc, scale, inv_log_det, out, batch, features);
    DPK_LAUNCH_CHECK("normal_prior_fwd_vec_kernel");
    return DPK_OK;
  }
  normal_prior_fwd_kernel<<<sample_grid(batch), 256, 0, st>>>(z, loc, scale, inv_log_det, out, batch, features);
  DPK_LAUNCH_CHECK("normal_prior_fwd_kernel");
  return DPK_OK;
}

extern "C" int dpk_normal_prior_backward(const float* z, const float* loc, const float* scale, const float* grad_out,
                                         float* grad_z, int64_t batch, int32_t features, void* stream) {
  if (batch < 0 || features <= 0) return set_error(DPK_E_ARG, "normal_prior_bwd: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!z || !grad_out || !grad_z) return set_error(DPK_E_ARG, "normal_prior_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(CAT_FLOW_BWD, st);
  normal_prior_bwd_kernel<<<flat_grid(batch * features), 256, 0, st>>>(z, loc, scale, grad_out, grad_z, batch, features);
  DPK_LAUNCH_CHECK("normal_prior_bwd_kernel");
  return DPK_OK;
}
