// layers.cu -- stand-alone layer entry points in the reference's tensor layouts.
//   ProductLayer.forward  deeprob/spn/layers/ratspn.py:272-286   -> dpk_outer_sum_forward
//   SumLayer.forward      deeprob/spn/layers/ratspn.py:363-378   -> dpk_mixture_forward
//   RootLayer.forward     deeprob/spn/layers/ratspn.py:446-458   -> dpk_mixture_forward with P = 1
// These are NOT the hot path (the model-level call never materialises the K^2 product); they exist
// so that code calling a single layer (e.g. the top-down mpe pass) runs on the same library, and
// they evaluate the mixture exactly in the log domain like torch.logsumexp.
#include <algorithm>

#include "common.cuh"

namespace dpk {

// out[b, p, i*K + j] = x[b, 2p, i] + x[b, 2p+1, j]
__global__ void outer_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t total, int P, int K) {
  const int K2 = K * K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ij = (int)(idx % K2);
    const int64_t bp = idx / K2;  // b * P + p
    const float* row = x + bp * 2 * K;
    out[idx] = row[ij / K] + row[K + ij % K];
  }
}

// lse[p, o] = logsumexp_k w[p, o, k]   (one warp per row)
__global__ void row_lse_kernel(const float* __restrict__ w, float* __restrict__ lse, int rows, int kin) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* r = w + (size_t)row * kin;
  float m = -INFINITY;
  for (int k = lane; k < kin; k += 32) m = fmaxf(m, r[k]);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < kin; k += 32) s += expf(r[k] - m);
  s = warp_sum(s);
  if (lane == 0) lse[row] = m + logf(s);
}

// out[b, p, o] = logsumexp_k (x[b, p, k] + w[p, o, k] - lse[p, o]); one warp per output element
__global__ void mixture_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ lse,
                               float* __restrict__ out, int64_t B, int P, int kin, int O) {
  const int64_t total = B * P * O;
  const int lane = threadIdx.x & 31;
  for (int64_t idx = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); idx < total;
       idx += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int o = (int)(idx % O);
    const int p = (int)((idx / O) % P);
    const int64_t b = idx / ((int64_t)O * P);
    const float* xr = x + (b * P + p) * (int64_t)kin;
    const float* wr = w + ((size_t)p * O + o) * kin;
    const float shift = lse[p * O + o];
    float m = -INFINITY;
    for (int k = lane; k < kin; k += 32) m = fmaxf(m, xr[k] + (wr[k] - shift));
    m = warp_max(m);
    float y;
    if (!(fabsf(m) <= FLT_MAX)) {
      y = m;
    } else {
      float s = 0.f;
      for (int k = lane; k < kin; k += 32) s += expf(xr[k] + (wr[k] - shift) - m);
      s = warp_sum(s);
      y = m + logf(s);
    }
    if (lane == 0) out[idx] = y;
  }
}

}  // namespace dpk

using namespace dpk;

extern "C" int dpk_outer_sum_forward(const float* x, int64_t batch, int32_t partitions, int32_t nodes, float* out,
                                     void* stream) {
  if (batch < 0 || partitions <= 0 || nodes <= 0) return set_error(DPK_E_ARG, "outer_sum: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !out) return set_error(DPK_E_ARG, "outer_sum: null pointer");
  const int64_t total = batch * partitions * (int64_t)nodes * nodes;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  ProfScope prof(CAT_LAYER, static_cast<cudaStream_t>(stream));
  outer_sum_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, total, partitions, nodes);
  DPK_LAUNCH_CHECK("outer_sum_kernel");
  return DPK_OK;
}

extern "C" int dpk_mixture_forward(const float* x, const float* weight, int64_t batch, int32_t partitions,
                                   int32_t in_nodes, int32_t out_nodes, float* out, float* scratch, void* stream) {
  if (batch < 0 || partitions <= 0 || in_nodes <= 0 || out_nodes <= 0) return set_error(DPK_E_ARG, "mixture: bad sizes");
  if (batch == 0) return DPK_OK;
  if (!x || !weight || !out || !scratch) return set_error(DPK_E_ARG, "mixture: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = partitions * out_nodes;
  ProfScope prof(CAT_LAYER, st, 2);
  row_lse_kernel<<<(rows + 3) / 4, 128, 0, st>>>(weight, scratch, rows, in_nodes);
  DPK_LAUNCH_CHECK("row_lse_kernel");
  const int64_t total = batch * partitions * (int64_t)out_nodes;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 8), (int64_t)sm_count() * 32);
  mixture_kernel<<<blocks, 256, 0, st>>>(x, weight, scratch, out, batch, partitions, in_nodes, out_nodes);
  DPK_LAUNCH_CHECK("mixture_kernel");
  return DPK_OK;
}
