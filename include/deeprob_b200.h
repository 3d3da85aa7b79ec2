/* deeprob_b200.h -- C ABI of libdeeprob_b200.so (hand-written sm_100a kernels for the deeprob-kit
 * tensorised log-likelihood path).
 *
 * The reference (deeprob-kit @ d96ac30) is pure Python on stock PyTorch ops: there is no FFI to
 * re-bind, so every entry point below cites the reference *Python* interface it replaces.  The
 * Python host side (package deeprob_kit_b200) binds these with ctypes; INTEGRATION.md shows the
 * stub a deeprob-kit maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host"; buffers are borrowed for the
 *     duration of the stream-ordered launches only (the caller -- PyTorch -- owns all memory);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value 0 = success, negative = error (DPK_E_*); text via dpk_last_error() (thread-local);
 *   - no global mutable state besides a per-process cache of device attributes; re-entrant per stream.
 */
#ifndef DEEPROB_B200_H_
#define DEEPROB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPK_ABI_VERSION 2
#define DPK_MAX_LEVELS 16

#define DPK_OK 0
#define DPK_E_ARG (-1)      /* bad shape / null pointer / unsupported size */
#define DPK_E_WORKSPACE (-2) /* workspace too small or misaligned */
#define DPK_E_CUDA (-3)     /* a CUDA runtime call failed (message has the CUDA error string) */

#define DPK_LEAF_GAUSSIAN 0
#define DPK_LEAF_BERNOULLI 1

/* flags */
#define DPK_F_SAVE_ACTIVATIONS 1 /* forward keeps per-level activations in the workspace for backward */
#define DPK_F_TABLES_VALID 2     /* the parameter-derived tables in `workspace` (leaf tables / operand images, softmax
                                  * tables) were built by a previous dpk_ratspn_forward with the same descriptor
                                  * contents, parameter values, batch size and flags: do not rebuild them.  The
                                  * caller (deeprob_kit_b200.spn._engine) keys this on the parameters' version
                                  * counters; the reference recomputes log_softmax(weight) on every call
                                  * (deeprob/spn/layers/ratspn.py:375). */

int dpk_abi_version(void);
const char* dpk_last_error(void);

/* Launch accounting for benchmarks/tests (no reference counterpart): every kernel launch of the
 * library is counted per category; with profiling enabled each launch group is also bracketed by
 * CUDA events on its stream.  dpk_profile_read fills `ms` (elapsed, only while enabled) and
 * `launches` (always) for categories 0..ncat-1 since the previous read, and resets them.
 * Categories: 0 prep, 1 ratspn leaf, 2 ratspn product+sum, 3 ratspn root, 4 ratspn bwd product+sum,
 * 5 ratspn bwd leaf, 6 finalize, 7 stand-alone layers, 8 dgcspn fwd, 9 dgcspn bwd, 10 flow fwd,
 * 11 flow bwd, 12 gemm, 13 ratspn leaf GEMM on the tensor cores (tcgen05),
 * 14 its operand-preparation launch. */
#define DPK_PROFILE_CATEGORIES 16
int dpk_profile_enable(int on);
int dpk_profile_read(double* ms, int64_t* launches, int32_t ncat);

/* ------------------------------------------------------------------------------------------------
 * RAT-SPN  (deeprob/spn/models/ratspn.py:105-122 RatSpn.forward and the layers it chains:
 *   RegionGraphLayer.forward  deeprob/spn/layers/ratspn.py:87-108   (Gaussian :160-213, Bernoulli :216-247)
 *   ProductLayer.forward      deeprob/spn/layers/ratspn.py:272-286
 *   SumLayer.forward          deeprob/spn/layers/ratspn.py:363-378
 *   RootLayer.forward         deeprob/spn/layers/ratspn.py:446-458 )
 * Structure: G0 = repetitions * 2^depth leaf regions of `dimension` gathered features and
 * `leaf_channels` densities each; depth products, depth-1 inner sum levels, one root.
 * ---------------------------------------------------------------------------------------------- */
typedef struct dpk_ratspn_desc {
  int32_t leaf_kind;     /* DPK_LEAF_* */
  int32_t in_features;   /* D */
  int32_t depth;         /* rg_depth >= 1 */
  int32_t repetitions;   /* rg_repetitions */
  int32_t leaf_channels; /* K = rg_batch */
  int32_t sum_nodes;     /* O = rg_sum */
  int32_t out_classes;   /* C */
  int32_t dimension;     /* ceil(D / 2^depth) */
  const int32_t* mask;       /* (G0, dimension) int32 copy of base_layer.mask (ratspn.py:47-56) */
  const int32_t* region_len; /* (G0) number of real (non-pad) features of each region */
  const float* leaf_p0;      /* (G0, K, dimension) loc | logits */
  const float* leaf_p1;      /* (G0, K, dimension) scale, NULL for Bernoulli */
  const float* sum_weight[DPK_MAX_LEVELS]; /* host array: level l raw logits (G0>>(l+1), O, Kin_l) */
  const float* root_weight;  /* (C, repetitions * Kin_last) raw logits */
} dpk_ratspn_desc;

/* gradient / statistic outputs of dpk_ratspn_backward; any pointer may be NULL = not wanted.
 * All are ACCUMULATED INTO (caller zero-fills), fp32, same shapes as the parameters. */
typedef struct dpk_ratspn_grads {
  float* grad_x;        /* (B, D) */
  float* leaf_p0;       /* d/dloc or d/dlogits */
  float* leaf_p1;       /* d/dscale */
  float* sum_weight[DPK_MAX_LEVELS];
  float* root_weight;
} dpk_ratspn_grads;

/* EM sufficient statistics (extension; semantics of deeprob/spn/learning/em.py:99-107 on the
 * tensorised model): posterior counts of every sum/root weight and leaf moments. */
typedef struct dpk_ratspn_em_stats {
  float* sum_counts[DPK_MAX_LEVELS]; /* like sum_weight */
  float* root_counts;                /* like root_weight */
  float* s0;                         /* (G0, K, dimension) sum_b post * [x observed] */
  float* s1;                         /* (G0, K, dimension) sum_b post * x */
  float* s2;                         /* (G0, K, dimension) sum_b post * x^2 (Gaussian only, may be NULL) */
} dpk_ratspn_em_stats;

/* bytes of scratch the forward (and, with DPK_F_SAVE_ACTIVATIONS, the following backward) needs */
size_t dpk_ratspn_workspace_bytes(const dpk_ratspn_desc* desc, int64_t batch, uint32_t flags);

/* x (B, D) fp32 row-major, NaN = marginalised variable; out (B, C) fp32 */
int dpk_ratspn_forward(const dpk_ratspn_desc* desc, const float* x, int64_t batch, float* out,
                       void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* backward of sum_{b,c} grad_out[b,c] * out[b,c]; `workspace` is the one the forward filled with
 * DPK_F_SAVE_ACTIVATIONS; `out` the forward result. */
int dpk_ratspn_backward(const dpk_ratspn_desc* desc, const float* x, int64_t batch, const float* out,
                        const float* grad_out, const dpk_ratspn_grads* grads, void* workspace,
                        size_t workspace_bytes, void* stream);

/* E-step statistics of the batch (grad_out == 1 for the class column `cls`, posterior form). */
int dpk_ratspn_em_statistics(const dpk_ratspn_desc* desc, const float* x, int64_t batch,
                             const float* out, const dpk_ratspn_em_stats* stats, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Training mode with probabilistic dropout (deeprob/spn/layers/ratspn.py:98-100: NaN dropout on the per-dimension
 * leaf log-densities (B, G0, K, dim); :370-372: -inf dropout on every product layer output (B, P, K^2) in front of a
 * sum layer; the root has none).  The reference draws from torch's RNG stream (`torch.rand_like`), which no fused
 * kernel can reproduce: here every Bernoulli draw is a pure function of (seed, stream, element index) -- the
 * backward regenerates it -- so parity with the reference is distributional, and exact against the oracle with
 * the same draws injected (dpk_dropout_draw exposes the generator to the tests).  An element is dropped when
 * dpk_dropout_draw(seed, stream, index) < rate * 2^24; stream 0 = leaf elements ((b*G0+g)*K+k)*dim+d, stream 1+e =
 * sum level e elements (b*P_e+p)*Kin^2+ij. */
typedef struct dpk_ratspn_dropout {
  float in_rate;   /* RegionGraphLayer dropout in [0, 1), 0 = none */
  float sum_rate;  /* SumLayer dropout in [0, 1), 0 = none */
  uint64_t seed;   /* one fresh value per forward; the same value for its backward */
} dpk_ratspn_dropout;
uint32_t dpk_dropout_draw(uint64_t seed, uint32_t stream, uint64_t index); /* host: uniform 24-bit integer */
size_t dpk_ratspn_dropout_workspace_bytes(const dpk_ratspn_desc* desc, int64_t batch);
int dpk_ratspn_forward_dropout(const dpk_ratspn_desc* desc, const float* x, int64_t batch,
                               const dpk_ratspn_dropout* drop, float* out, void* workspace, size_t workspace_bytes,
                               void* stream);
/* gradients as in dpk_ratspn_backward (accumulated into zero-filled buffers); `workspace` from the forward */
int dpk_ratspn_backward_dropout(const dpk_ratspn_desc* desc, const float* x, int64_t batch,
                                const dpk_ratspn_dropout* drop, const float* out, const float* grad_out,
                                const dpk_ratspn_grads* grads, void* workspace, size_t workspace_bytes, void* stream);

/* Top-down passes (deeprob/spn/models/ratspn.py:124-182 RatSpn.mpe / RatSpn.sample and the layers' mpe / sample
 * methods, deeprob/spn/layers/ratspn.py:118-157,288-330,380-417,460-490), one thread per sample.
 * dpk_ratspn_mpe: `workspace` is the one a dpk_ratspn_forward with DPK_F_SAVE_ACTIVATIONS filled for the same x,
 * `out` its result; y (B) int32 classes or NULL (= argmax of out, C > 1); filled (B, D) = x with every NaN entry
 * replaced by the mode of the leaf channel the maximising path selects.  Ties resolve like torch.argmax.
 * dpk_ratspn_sample: ancestral samples (n, D); y (n) int32 classes or NULL (class 0); draws come from the
 * counter-based generator (seed); `workspace` of dpk_ratspn_workspace_bytes(desc, n, DPK_F_SAVE_ACTIVATIONS). */
int dpk_ratspn_mpe(const dpk_ratspn_desc* desc, const float* x, int64_t batch, const float* out, const int32_t* y,
                   float* filled, void* workspace, size_t workspace_bytes, void* stream);
int dpk_ratspn_sample(const dpk_ratspn_desc* desc, int64_t n_samples, const int32_t* y, uint64_t seed, float* samples,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Loss of the SPN models (deeprob/spn/models/ratspn.py:184-191, deeprob/spn/models/dgcspn.py:189-196) as one launch:
 * classes == 1: loss = -mean(ll);  classes > 1: loss = mean_b(logsumexp_c ll[b,:] - ll[b, y[b]])  (y: int64 labels).
 * `loss` (1 float, device) is overwritten; `grad` (B, classes) = d loss / d ll, may be NULL. */
int dpk_nll_loss(const float* ll, const int64_t* y, int64_t batch, int32_t classes, float* loss, float* grad, void* stream);

/* Stand-alone layers with the reference layouts (used by the nn.Module layer classes). */
/* RegionGraphLayer.forward: out (B, G0, K) */
int dpk_ratspn_leaf_forward(const dpk_ratspn_desc* desc, const float* x, int64_t batch, float* out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ProductLayer.forward (deeprob/spn/layers/ratspn.py:272-286): x (B, 2P, K) -> out (B, P, K*K),
 * out[b,p,i*K+j] = x[b,2p,i] + x[b,2p+1,j] */
int dpk_outer_sum_forward(const float* x, int64_t batch, int32_t partitions, int32_t nodes, float* out,
                          void* stream);

/* SumLayer.forward (ratspn.py:363-378) / RootLayer.forward (:446-458, partitions = 1):
 * x (B, P, Kin), weight (P, O, Kin) raw logits -> out (B, P, O) = logsumexp_k(x + log_softmax_k(weight));
 * scratch: P*O floats. Exact log-domain evaluation. */
int dpk_mixture_forward(const float* x, const float* weight, int64_t batch, int32_t partitions,
                        int32_t in_nodes, int32_t out_nodes, float* out, float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DGC-SPN layers (deeprob/spn/layers/dgcspn.py), NCHW fp32 like the reference.  Backward entry
 * points: grad_x is OVERWRITTEN (may be NULL = not wanted), parameter gradients are ACCUMULATED INTO.
 * ---------------------------------------------------------------------------------------------- */
/* SpatialGaussianLayer.forward (dgcspn.py:101-120): x (B,Cin,H,W); loc, scale (K,Cin,H,W) -> out (B,K,H,W) */
int dpk_dgc_leaf_forward(const float* x, const float* loc, const float* scale, int64_t batch,
                         int32_t in_channels, int32_t out_channels, int32_t hw, float* out, void* stream);
int dpk_dgc_leaf_backward(const float* x, const float* loc, const float* scale, const float* grad_out,
                          int64_t batch, int32_t in_channels, int32_t out_channels, int32_t hw, float* grad_x,
                          float* grad_loc, float* grad_scale, void* stream);

/* SpatialProductLayer.forward (dgcspn.py:224-236): zero pad (pad_left/pad_top; right/bottom implied by the
 * output size) + 2x2 conv with all-ones depthwise kernels, or the one-hot "all combinations" kernels
 * (out_channels = channels^4, dgcspn.py:187-193) when depthwise == 0. */
typedef struct dpk_dgc_product_desc {
  int32_t channels, height, width;             /* input  (C, H, W) */
  int32_t out_channels, out_height, out_width; /* output (C | C^4, OH, OW) */
  int32_t pad_top, pad_left;
  int32_t stride_h, stride_w, dilation_h, dilation_w;
  int32_t depthwise;
} dpk_dgc_product_desc;
int dpk_dgc_product_forward(const dpk_dgc_product_desc* desc, const float* x, int64_t batch, float* out, void* stream);
int dpk_dgc_product_backward(const dpk_dgc_product_desc* desc, const float* grad_out, int64_t batch, float* grad_x,
                             void* stream);

/* SpatialSumLayer.forward (dgcspn.py:289-304): x (B,Cin,H,W), weight (Cout,Cin,H,W) raw logits ->
 * out (B,Cout,H,W).  scratch: 2*Cout*Cin*HW floats (forward), 3*Cout*Cin*HW floats (backward). */
int dpk_dgc_sum_forward(const float* x, const float* weight, int64_t batch, int32_t in_channels,
                        int32_t out_channels, int32_t hw, float* out, float* scratch, void* stream);
int dpk_dgc_sum_backward(const float* x, const float* weight, const float* out, const float* grad_out,
                         int64_t batch, int32_t in_channels, int32_t out_channels, int32_t hw, float* grad_x,
                         float* grad_weight, float* scratch, void* stream);

/* Inference fusion of a depthwise SpatialProductLayer (dgcspn.py:224-236) with the SpatialSumLayer that follows it
 * (dgcspn.py:289-304): x (B,C,H,W) -> out (B,Cout,OH,OW) without materialising the product output.
 * weight (Cout, C, OH, OW) raw logits; scratch: 2*Cout*C*OH*OW floats.  C in {2,4,8,16}; other shapes: DPK_E_ARG. */
int dpk_dgc_prodsum_forward(const dpk_dgc_product_desc* desc, const float* x, const float* weight, int64_t batch,
                            int32_t out_channels, float* out, float* scratch, void* stream);

/* Backward of that pair for training (autograd of dgcspn.py:224-236 + :289-304 through the fused forward, which keeps no
 * product output): the product values are recomputed from the four taps of x (B,C,H,W); out / grad_out (B,Cout,OH,OW).
 * grad_prod (B,C,OH,OW) or NULL = gradient w.r.t. the product OUTPUT (dpk_dgc_product_backward then gives d/dx);
 * grad_weight (Cout,C,OH,OW) accumulated or NULL; scratch: 3*Cout*C*OH*OW floats.  C, Cout <= 8, else DPK_E_ARG. */
int dpk_dgc_prodsum_backward(const dpk_dgc_product_desc* desc, const float* x, const float* weight, const float* out,
                             const float* grad_out, int64_t batch, int32_t out_channels, float* grad_prod,
                             float* grad_weight, float* scratch, void* stream);

/* SpatialRootLayer.forward (dgcspn.py:343-355): x (B, Q = C*H*W), weight (classes, Q) -> out (B, classes).
 * scratch: classes*Q floats (forward), 2*classes*Q floats (backward). */
int dpk_dgc_root_forward(const float* x, const float* weight, int64_t batch, int64_t features,
                         int32_t out_classes, float* out, float* scratch, void* stream);
int dpk_dgc_root_backward(const float* x, const float* weight, const float* out, const float* grad_out,
                          int64_t batch, int64_t features, int32_t out_classes, float* grad_x,
                          float* grad_weight, float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Normalizing-flow bijectors (deeprob/flows).  Rows are per-sample flattened (B, N) arrays with an
 * explicit row stride, so channel-wise halves of a (B,C,H,W) tensor can be passed without a copy.
 * ---------------------------------------------------------------------------------------------- */
/* Affine coupling transform + log-det (CouplingLayer1d coupling.py:72-104, CouplingLayer2d :179-272,
 * AutoregressiveLayer autoregressive.py:72-79).  z = conditioner output rows: t = z[0:N], raw scale =
 * z[N:2N]; s = inv_mask * w * tanh(raw) with w = scale_weight[(e / w_inner) % w_count] (ScaledTanh,
 * torch/utils.py:52-70).  direction 0: out = (x - inv_mask*t) * exp(-s), log_det -= sum s ("backward" =
 * density direction); direction 1: out = x * exp(s) + inv_mask*t, log_det += sum s (sampling direction).
 * affine == 0: translation only (NICE), log_det untouched. */
typedef struct dpk_coupling_desc {
  int64_t batch;
  int32_t features;          /* N: transformed elements per sample */
  int32_t affine;
  int32_t direction;
  int32_t w_count, w_inner;  /* ScaledTanh weight broadcast pattern */
  int64_t x_stride, z_stride;
  const float* inv_mask;     /* (N) or NULL (= ones) */
  const float* scale_weight; /* (w_count) */
} dpk_coupling_desc;
int dpk_coupling_forward(const dpk_coupling_desc* desc, const float* x, const float* z, float* out,
                         int64_t out_stride, float* log_det /* (B), accumulated into; may be NULL */, void* stream);
/* Inference variant for a conditioner that evaluated only the live output columns (inv_mask != 0; the others never
 * reach the result because coupling.py:79-80 multiplies t and s by inv_mask): z rows hold [t_live | s_live],
 * z_index[e] = column of element e (ignored where inv_mask[e] == 0), z_half = offset of the s half; z_index NULL =
 * the plain layout.  post_scale/post_shift (N, or both NULL) apply the following eval-mode BatchNormLayer1d
 * (flows/utils.py:118-139) as out = out*post_scale + post_shift and add post_log_det to every log_det entry.
 * live_out (B, z_half; may be NULL, needs z_index) additionally receives the transformed elements alone,
 * live_out[b][z_index[e]] = out[b][e]: with alternating masks these are exactly the columns the next coupling's
 * conditioner reads, which saves its gather. */
int dpk_coupling_forward_compact(const dpk_coupling_desc* desc, const float* x, const float* z,
                                 const int32_t* z_index, int32_t z_half, const float* post_scale,
                                 const float* post_shift, float post_log_det, float* out, int64_t out_stride,
                                 float* live_out, float* log_det, void* stream);
/* grad_x (may be NULL) and grad_z are overwritten; grad_scale_weight (w_count, may be NULL) is accumulated into */
int dpk_coupling_backward(const dpk_coupling_desc* desc, const float* x, const float* z, const float* grad_out,
                          int64_t grad_out_stride, const float* grad_log_det, float* grad_x, int64_t grad_x_stride,
                          float* grad_z, int64_t grad_z_stride, float* grad_scale_weight, void* stream);

/* Building blocks of the batch-norm bijector (flows/utils.py:118-153, 183-221) over (B, F, I) arrays
 * (1d: I = 1; 2d: I = H*W).  mode 0: sum_out[f] += sum x; mode 1: sum_out[f] += sum (x - center[f])^2;
 * mode 2: sum_out[f] += sum other, dot_out[f] += sum other * (x - center[f]). */
int dpk_feature_reduce(const float* x, const float* center, const float* other, float* sum_out, float* dot_out,
                       int64_t batch, int32_t features, int32_t inner, int32_t mode, void* stream);
/* out = x * a[f] + c[f]  (+ k[f] * (y - mu[f]) when y != NULL) */
int dpk_feature_affine(const float* x, const float* a, const float* c, const float* y, const float* k,
                       const float* mu, float* out, int64_t batch, int32_t features, int32_t inner, void* stream);

/* DequantizeLayer + LogitLayer apply_backward fused (flows/utils.py:244-248, 276-284): q = (x*(bins-1) +
 * noise)/bins (bins <= 0: q = x), out = logit(alpha + (1-2*alpha)*q) (alpha < 0: out = q);
 * inv_log_det[b] -= sum (log y + log(1-y)); the constant terms are added by the caller. */
int dpk_flow_preprocess_forward(const float* x, const float* noise, float bins, float alpha, float* out,
                                float* inv_log_det, int64_t batch, int32_t features, void* stream);
int dpk_flow_preprocess_backward(const float* x, const float* noise, float bins, float alpha,
                                 const float* grad_out, const float* grad_inv_log_det, float* grad_x,
                                 int64_t batch, int32_t features, void* stream);

/* Prior + final sum (flows/models/base.py:139-143): out[b] = sum_e logN(z[b,e]; loc[e], scale[e]) +
 * inv_log_det[b]; loc/scale NULL = standard normal. */
int dpk_normal_prior_forward(const float* z, const float* loc, const float* scale, const float* inv_log_det,
                             float* out, int64_t batch, int32_t features, void* stream);
int dpk_normal_prior_backward(const float* z, const float* loc, const float* scale, const float* grad_out,
                              float* grad_z, int64_t batch, int32_t features, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layer of the coupling / autoregressive conditioners (nn.Linear in deeprob/flows/layers/coupling.py:45-56,
 * MaskedLinear deeprob/torch/utils.py:73-96 with weight * mask folded by the caller):
 *   out (B, out_features) = act(x (B, in_features) @ weight (out_features, in_features)^T + bias), act = ReLU or none,
 * as a tcgen05 GEMM with fp32-accurate 3-pass hi/lo fp16 operands (the kernels of the RAT-SPN leaf level); rows with
 * non-finite or |x| > 6e4 inputs are re-evaluated exactly in fp32.  Inference only (no backward entry point).
 * in_features % 4 == 0; x, out 16-byte aligned; `flags` may carry DPK_F_TABLES_VALID (weight images still valid). */
size_t dpk_linear_workspace_bytes(int64_t batch, int32_t in_features, int32_t out_features);
int dpk_linear_forward(const float* x, const float* weight, const float* bias /* may be NULL */, int64_t batch,
                       int32_t in_features, int32_t out_features, int32_t relu, float* out, void* workspace,
                       size_t workspace_bytes, uint32_t flags, void* stream);

/* Backward of the same layer on the same GEMM (tcgen05, 3-pass hi/lo fp16 operands, fp32 accumulate):
 *   g = dy * [y > 0] (relu != 0; y = the forward output) | dy;   dx (B, in) = g . weight;   dw (out, in) = g^T . x
 *   (contraction over the batch: split-K with an atomic fp32 epilogue);   db (out) = column sums of g.
 * dx / dw / db may be NULL (= not wanted); dw and db are OVERWRITTEN.  Operands are scaled into the fp16 range by exact
 * powers of two derived on the device from their max magnitudes (gradients are routinely < 1e-6).
 * Replaces the autograd backward of nn.Linear / MaskedLinear (+ ReLU) in deeprob/flows/layers/coupling.py:45-56 and
 * deeprob/flows/layers/autoregressive.py:72-79. */
size_t dpk_linear_backward_workspace_bytes(int64_t batch, int32_t in_features, int32_t out_features);
int dpk_linear_backward(const float* x, const float* weight, const float* y, const float* dy, int64_t batch,
                        int32_t in_features, int32_t out_features, int32_t relu, float* dx, float* dw, float* db,
                        void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPROB_B200_H_ */
