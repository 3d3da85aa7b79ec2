// ratspn_fwd.cu -- C-ABI entry points of the RAT-SPN forward (RatSpn.forward,
// deeprob/spn/models/ratspn.py:105-122): parameter tables -> leaf level -> product+sum levels -> root.
#include <algorithm>

#include "ratspn_kernels.cuh"

namespace dpk {

int ratspn_check_ws(const RatPlan& p, const void* ws, size_t bytes) {
  if (!ws) return set_error(DPK_E_WORKSPACE, "null workspace");
  if ((uintptr_t)ws % 256) return set_error(DPK_E_WORKSPACE, "workspace must be 256-byte aligned");
  if (bytes < p.total_floats * 4)
    return set_error(DPK_E_WORKSPACE, "workspace too small: %zu < %zu bytes", bytes, p.total_floats * 4);
  return DPK_OK;
}

}  // namespace dpk

using namespace dpk;

extern "C" size_t dpk_ratspn_workspace_bytes(const dpk_ratspn_desc* desc, int64_t batch, uint32_t flags) {
  // one buffer serves dpk_ratspn_forward and the stand-alone leaf layer (dpk_ratspn_leaf_forward), whose plans differ
  RatPlan p, q;
  if (make_plan(desc, batch, flags & DPK_F_SAVE_ACTIVATIONS, &p)) return 0;
  if (make_plan(desc, batch, (flags & DPK_F_SAVE_ACTIVATIONS) | kPlanLeafOnly, &q)) return 0;
  return std::max(p.total_floats, q.total_floats) * sizeof(float);
}

extern "C" int dpk_ratspn_forward(const dpk_ratspn_desc* desc, const float* x, int64_t batch, float* out,
                                  void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  RatPlan p;
  int rc = make_plan(desc, batch, flags & DPK_F_SAVE_ACTIVATIONS, &p);
  if (rc) return rc;
  if (batch == 0) return DPK_OK;
  if (!x || !out || !desc->mask || !desc->region_len || !desc->leaf_p0 || !desc->root_weight)
    return set_error(DPK_E_ARG, "null pointer argument");
  for (int e = 0; e < p.n_sum; ++e)
    if (!desc->sum_weight[e]) return set_error(DPK_E_ARG, "null sum_weight[%d]", e);
  if ((rc = ratspn_check_ws(p, workspace, workspace_bytes))) return rc;
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(flags & DPK_F_TABLES_VALID)) {
    ProfScope prof(CAT_PREP, st, 3 + 2 * p.n_sum + (p.leaf_mma ? 2 : 0) + (p.tree_mma ? p.depth : 0));
    if ((rc = ratspn_run_prep_leaf(desc, p, ws, st))) return rc;
    if ((rc = ratspn_run_prep_weights(desc, p, ws, st))) return rc;
    if (p.tree_mma && (rc = ratspn_run_prep_tree(p, ws, st))) return rc;
  }
  if ((rc = ratspn_run_leaf(desc, p, x, ws, st))) return rc;
  if (p.tree_mma) return ratspn_run_tree(p, ws, out, st);
  return ratspn_run_upper(p, ws, out, st);
}
