// train_glue.cu -- the caller side of the path during training: the loss of the SPN models as ONE kernel.
//
//   RatSpn.loss / DgcSpn.loss   deeprob/spn/models/ratspn.py:184-191, deeprob/spn/models/dgcspn.py:189-196
//     out_classes == 1:  loss = -mean(ll)
//     out_classes  > 1:  loss = F.nll_loss(log_softmax(ll, dim=1), y)  = mean_b( logsumexp_c ll[b,:] - ll[b, y_b] )
// The reference composes it from 2-4 tensor ops and then reads it with loss.item() every step
// (deeprob/torch/routines.py:158-166).  Here one launch writes the scalar loss and the gradient w.r.t. the
// log-likelihoods (kept for the backward), so the training step adds no extra pass or synchronisation of its own.
#include <algorithm>

#include "common.cuh"

namespace dpk {
namespace {

__global__ void nll_loss_kernel(const float* __restrict__ ll, const int64_t* __restrict__ y, int64_t B, int C,
                                float* __restrict__ loss, float* __restrict__ grad) {
  const float inv_b = 1.f / (float)B;
  float local = 0.f;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float* row = ll + b * C;
    if (C == 1) {
      local -= row[0];
      if (grad) grad[b] = -inv_b;
      continue;
    }
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, row[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(row[c] - m);
    const float lse = m + logf(s);
    const int64_t t = y[b];
    local += lse - row[t];
    if (grad)
      for (int c = 0; c < C; ++c) grad[b * C + c] = (expf(row[c] - lse) - (c == t ? 1.f : 0.f)) * inv_b;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, local * inv_b);
}

}  // namespace
}  // namespace dpk

using namespace dpk;

extern "C" int dpk_nll_loss(const float* ll, const int64_t* y, int64_t batch, int32_t classes, float* loss, float* grad,
                            void* stream) {
  if (batch <= 0 || classes <= 0) return set_error(DPK_E_ARG, "nll_loss: bad shape");
  if (!ll || !loss || (classes > 1 && !y)) return set_error(DPK_E_ARG, "nll_loss: null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DPK_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
  ProfScope prof(CAT_LAYER, st);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(batch, 256), 2 * (int64_t)sm_count());
  nll_loss_kernel<<<grid, 256, 0, st>>>(ll, y, batch, classes, loss, grad);
  DPK_LAUNCH_CHECK("nll_loss_kernel");
  return DPK_OK;
}
