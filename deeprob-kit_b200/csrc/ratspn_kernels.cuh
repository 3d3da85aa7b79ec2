// ratspn_kernels.cuh -- host-side drivers shared by the RAT-SPN translation units.
#pragma once
#include "ratspn_plan.cuh"

namespace dpk {

int ratspn_check_ws(const RatPlan& p, const void* ws, size_t bytes);
// leaf tables (tab / cd / cst) from the leaf parameters           [ratspn_leaf.cu]
int ratspn_run_prep_leaf(const dpk_ratspn_desc* d, const RatPlan& p, float* ws, cudaStream_t st);
// tensor-core leaf: weight images / constants, and x -> act[0] for every clean 32-sample group (flags the rest)
int ratspn_run_prep_leaf_mma(const dpk_ratspn_desc* d, const RatPlan& p, float* ws, cudaStream_t st);
int ratspn_run_leaf_mma(const RatPlan& p, const float* x, float* ws, cudaStream_t st);
// narrow models (G0 * K <= 256): x streamed once through TMA                         [ratspn_leaf_stream.cu]
int ratspn_run_leaf_stream(const RatPlan& p, const float* x, float* ws, cudaStream_t st);
// leaf moments S0/S1/S2 of the backward as a tcgen05 GEMM over the batch; *fallback = device flag (!= 0: inputs
// outside the fast path's range, the caller's exact kernel must run)                [ratspn_leaf_mma.cu]
int ratspn_run_leaf_stats_mma(const dpk_ratspn_desc* d, const RatPlan& p, const float* x, const float* g0, float* ws,
                              float* s1, float* s2, float* s0tot, const int** fallback, cudaStream_t st);
// d LL / d x of the leaf level as a tcgen05 GEMM over the (region, channel) posteriors g0 = [G0*K][Bp]; accumulates
// into gx (B, D)                                                                     [ratspn_leaf_mma.cu]
int ratspn_run_leaf_bwd_x_mma(const dpk_ratspn_desc* d, const RatPlan& p, const float* x, const float* g0, float* ws,
                              float* gx, cudaStream_t st);
// softmax / log-softmax tables of every sum level and of the root [ratspn_einsum.cu]
int ratspn_run_prep_weights(const dpk_ratspn_desc* d, const RatPlan& p, float* ws, cudaStream_t st);
// product+sum level with the contraction on the tensor cores                      [ratspn_einsum_mma.cu]
bool ratspn_einsum_mma_eligible(int Kin, int O, int nOc, int64_t Bp);
size_t ratspn_einsum_mma_image_floats(int P, int Kin, int O);
int ratspn_run_prep_einsum_mma(const float* wsoft, int P, int O, int Kin, int OC, float* wimg, cudaStream_t st);
int ratspn_run_einsum_mma(const float* in, const float* wimg, const float* wsoft, const float* wlog, float* out, int64_t Bp, int P, int Kin,
                          int O, int OC, int cat, cudaStream_t st);
// all product + sum levels and the root fused, mixtures on tcgen05 (inference)      [ratspn_tree_mma.cu]
int ratspn_run_prep_tree(const RatPlan& p, float* ws, cudaStream_t st);
int ratspn_run_tree(const RatPlan& p, float* ws, float* out, cudaStream_t st);
// x -> act[0]                                                      [ratspn_leaf.cu]
int ratspn_run_leaf(const dpk_ratspn_desc* d, const RatPlan& p, const float* x, float* ws, cudaStream_t st);
// act[0] -> ... -> out (B, C)                                      [ratspn_einsum.cu]
int ratspn_run_upper(const RatPlan& p, float* ws, float* out, cudaStream_t st);

}  // namespace dpk
