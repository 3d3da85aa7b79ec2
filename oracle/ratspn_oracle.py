"""CPU oracle for the RAT-SPN log-likelihood path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional torch-CPU restatement of the reference algorithm, layer by layer, materialising the
same intermediates the reference materialises (it is also the timed `cpu_baseline` / `--impl
reference` arm of bench.py, so it deliberately keeps the reference's op sequence).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module.

Pinned against the live reference by `tests/golden/ratspn_*.npz` (made by oracle/make_golden.py,
which imports /root/reference) -- see tests/test_oracle_golden.py.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------
# Region graph (deeprob/utils/region.py:55-99, deeprob/utils/random.py:11-27)
# ----------------------------------------------------------------------------------------------
def region_graph_layers(n_features: int, depth: int, n_repetitions: int, seed) -> List[list]:
    """Layers [root, partitions, regions, ..., leaf regions] over `n_repetitions` random binary splits.

    region.py:64-74: every region of the previous level is permuted with RandomState.permutation and
    cut at len//2, the two halves sorted; region.py:91-97: repetitions are concatenated level-wise.
    """
    if n_features <= 0 or depth <= 0 or depth > int(np.log2(n_features)) or n_repetitions <= 0:
        raise ValueError("invalid region graph arguments")
    rng = seed if isinstance(seed, np.random.RandomState) else np.random.RandomState(seed)
    levels: List[list] = [[tuple(range(n_features))]] + [[] for _ in range(2 * depth)]
    for _ in range(n_repetitions):
        frontier = [tuple(range(n_features))]
        for lvl in range(depth):
            children, parts = [], []
            for reg in frontier:
                order = rng.permutation(reg).tolist()
                half = len(reg) // 2
                left, right = tuple(sorted(order[:half])), tuple(sorted(order[half:]))
                children += [left, right]
                parts.append((left, right))
            levels[2 * lvl + 1] = levels[2 * lvl + 1] + parts
            levels[2 * lvl + 2] = levels[2 * lvl + 2] + children
            frontier = children
    return levels


def leaf_tables(leaf_regions: Sequence[tuple], n_features: int, depth: int):
    """Gather table, pad mask and dimension (deeprob/spn/layers/ratspn.py:41-56)."""
    pad = -n_features % (2 ** depth)
    dim = (n_features + pad) // (2 ** depth)
    mask = np.zeros((len(leaf_regions), dim), dtype=np.int64)
    pad_mask = np.zeros((len(leaf_regions), 1, dim), dtype=bool)
    for g, reg in enumerate(leaf_regions):
        filled = tuple(reg) + (reg[-1],) * (dim - len(reg))
        mask[g] = filled
        pad_mask[g, 0, len(reg):] = True
    return torch.from_numpy(mask), (torch.from_numpy(pad_mask) if pad > 0 else None), dim, pad


# ----------------------------------------------------------------------------------------------
# Layers
# ----------------------------------------------------------------------------------------------
_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def gaussian_log_density(v: torch.Tensor, loc: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Normal.log_prob (torch/distributions/normal.py:87-102)."""
    return -((v - loc) ** 2) / (2 * scale ** 2) - scale.log() - _LOG_SQRT_2PI


def bernoulli_log_density(v: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Bernoulli.log_prob = -BCEWithLogits (torch/distributions/bernoulli.py:121-125)."""
    lg, vv = torch.broadcast_tensors(logits, v)
    return -torch.nn.functional.binary_cross_entropy_with_logits(lg, vv, reduction="none")


def leaf_layer(x, mask, pad_mask, kind: str, params: Dict[str, torch.Tensor], clean_nan: bool = False,
               drop: Optional[torch.Tensor] = None) -> torch.Tensor:
    """RegionGraphLayer.forward (deeprob/spn/layers/ratspn.py:87-108): (B,D)->(B,G0,K).

    `drop` (bool, (B,G0,K,dim)) restates the training-mode input dropout of :98-100,
    `x[torch.rand_like(x) < self.dropout] = np.nan` on the per-dimension log-densities, with the Bernoulli draws
    injected instead of taken from torch's RNG stream (a fused kernel cannot reproduce that stream; the tests
    inject the masks the CUDA path derives from its counter-based generator).

    `clean_nan=True` evaluates marginalised (NaN) inputs at 0 and masks their terms with `where`: the
    values are identical, but autograd then yields the gradient of the marginalised likelihood
    instead of the NaN the reference's op sequence produces (0 * NaN in the backward of :96/:103).
    """
    v = x[:, mask].unsqueeze(2)                                      # :95   (B,G0,1,dim)
    missing = torch.isnan(v)
    if clean_nan:
        v = torch.where(missing, torch.zeros_like(v), v)
    if kind == "gaussian":
        ll = gaussian_log_density(v, params["loc"], params["scale"])  # :96   (B,G0,K,dim)
    elif kind == "bernoulli":
        ll = bernoulli_log_density(v, params["logits"])
    else:
        raise ValueError(kind)
    if drop is not None:
        ll = torch.where(drop, torch.full_like(ll, float("nan")), ll) # :98-100 (out of place: autograd-friendly)
    ll = torch.nan_to_num(ll)                                         # :103  NaN->0, +-inf->+-FLT_MAX
    if clean_nan:
        ll = torch.where(missing, torch.zeros_like(ll), ll)
    if pad_mask is not None:
        ll = ll.masked_fill(pad_mask, 0.0)                            # :106-107
    return ll.sum(-1)                                                 # :108


def product_layer(x: torch.Tensor) -> torch.Tensor:
    """ProductLayer.forward (ratspn.py:272-286): siblings (2p,2p+1), out index i*K+j, i from 2p."""
    left, right = x[:, 0::2], x[:, 1::2]
    out = left.unsqueeze(3) + right.unsqueeze(2)
    return out.reshape(x.shape[0], x.shape[1] // 2, x.shape[2] ** 2)


def sum_layer(x: torch.Tensor, weight: torch.Tensor, drop: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SumLayer.forward (ratspn.py:363-378): (B,P,Kin)+(P,O,Kin) -> (B,P,O).  `drop` (bool, (B,P,Kin)) restates the
    training-mode dropout of :370-372, `x[torch.rand_like(x) < self.dropout] = -np.inf`, with injected draws."""
    if drop is not None:
        x = torch.where(drop, torch.full_like(x, float("-inf")), x)
    w = torch.log_softmax(weight, dim=2)
    return torch.logsumexp(x.unsqueeze(2) + w, dim=3)


def root_layer(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """RootLayer.forward (ratspn.py:446-458): (B,P,Kin)->(B,C)."""
    flat = x.flatten(1)
    w = torch.log_softmax(weight, dim=1)
    return torch.logsumexp(flat.unsqueeze(1) + w, dim=2)


# ----------------------------------------------------------------------------------------------
# Whole model (deeprob/spn/models/ratspn.py:72-122)
# ----------------------------------------------------------------------------------------------
class RatSpnOracle:
    """Structure + parameter holder; `log_prob` follows RatSpn.forward (models/ratspn.py:105-122)."""

    def __init__(self, in_features, kind, depth, repetitions, leaf_channels, sum_nodes, out_classes, seed):
        self.kind = kind
        self.in_features, self.depth, self.repetitions = in_features, depth, repetitions
        self.K, self.O, self.C = leaf_channels, sum_nodes, out_classes
        levels = region_graph_layers(in_features, depth, repetitions, seed)
        self.leaf_regions = levels[-1]                                # models/ratspn.py:77 (reversed)[0]
        self.mask, self.pad_mask, self.dim, self.pad = leaf_tables(self.leaf_regions, in_features, depth)
        self.params: Dict[str, torch.Tensor] = {}
        self.sum_weights: List[torch.Tensor] = []
        self.root_weight: Optional[torch.Tensor] = None

    # state_dict keys of the reference: base_layer.*, layers.{odd}.weight, root_layer.weight
    def load_reference_state(self, state: Dict[str, torch.Tensor]) -> "RatSpnOracle":
        if self.kind == "gaussian":
            self.params = {"loc": state["base_layer.loc"].float(), "scale": state["base_layer.scale"].float()}
        else:
            self.params = {"logits": state["base_layer.logits"].float()}
        assert torch.equal(state["base_layer.mask"].long(), self.mask), "region graph restatement mismatch"
        keys = sorted((k for k in state if k.startswith("layers.") and k.endswith(".weight")),
                      key=lambda k: int(k.split(".")[1]))
        self.sum_weights = [state[k].float() for k in keys]
        self.root_weight = state["root_layer.weight"].float()
        return self

    clean_nan = False   # see leaf_layer

    def double(self) -> "RatSpnOracle":
        """Same model in float64 (ground truth for gradient comparisons)."""
        self.params = {k: v.double() for k, v in self.params.items()}
        self.sum_weights = [w.double() for w in self.sum_weights]
        self.root_weight = self.root_weight.double()
        return self

    def leaf(self, x, drop=None):
        return leaf_layer(x, self.mask, self.pad_mask, self.kind, self.params, self.clean_nan, drop)

    leaf_drop = None    # injected training-mode dropout masks (see leaf_layer / sum_layer): (B,G0,K,dim) bool
    sum_drops = None    # list over the sum levels of (B,P,Kin) bool

    def log_prob(self, x: torch.Tensor, keep: Optional[list] = None) -> torch.Tensor:
        h = self.leaf(x, self.leaf_drop)
        if keep is not None:
            keep.append(h)
        for lvl in range(self.depth):                                 # Product, Sum, ..., Product
            h = product_layer(h)
            if lvl < self.depth - 1:
                h = sum_layer(h, self.sum_weights[lvl], self.sum_drops[lvl] if self.sum_drops is not None else None)
                if keep is not None:
                    keep.append(h)
        return root_layer(h, self.root_weight)

    def mpe(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """RatSpn.mpe (models/ratspn.py:124-160): bottom-up pass keeping the input of every layer, then top-down
        RootLayer.mpe (layers/ratspn.py:460-473: argmax over all (partition, i*K+j) of x + log_softmax(W)[y]),
        ProductLayer.mpe (:288-330: (group, i*K+j) -> (2*group, i), (2*group+1, j)), SumLayer.mpe (:380-396: argmax
        over the inputs of the selected sum node of x[group] + log_softmax(W[group, o])) and the leaf's mode of the
        selected channel (:118-137; Normal mean :212-213, Bernoulli [p >= 1/2] :245-247) written to the NaN entries.
        The last step is restated as an explicit scatter through the gather table: identical to the reference's
        gather(inv_mask) for pad == 0; for pad > 0 the reference's unpad step (:82-84) keeps the PADDED entries --
        a bug this restatement (and the CUDA path) does not reproduce."""
        h = self.leaf(x)
        inputs = []                                   # inputs of the sum layers, bottom-up
        for lvl in range(self.depth):
            h = product_layer(h)
            if lvl < self.depth - 1:
                inputs.append(h)
                h = sum_layer(h, self.sum_weights[lvl])
        n = x.shape[0]
        rows = torch.arange(n).unsqueeze(1)
        if self.C == 1:
            y = torch.zeros(n, dtype=torch.long)
        elif y is None:
            y = torch.argmax(root_layer(h, self.root_weight), dim=1)
        kin2 = h.shape[2]
        idx = torch.argmax(h.flatten(1) + torch.log_softmax(self.root_weight, dim=1)[y], dim=1, keepdim=True)
        group, offset = torch.div(idx, kin2, rounding_mode="floor"), torch.remainder(idx, kin2)
        for lvl in range(self.depth - 1, -1, -1):
            # product layer `lvl`: split every (partition, i*K+j) into its two child regions
            k = int(round(math.sqrt(kin2)))
            group = torch.stack([group * 2, group * 2 + 1], dim=2).flatten(1)
            offset = torch.stack([torch.div(offset, k, rounding_mode="floor"), torch.remainder(offset, k)], dim=2).flatten(1)
            if lvl > 0:
                xin = inputs[lvl - 1][rows, group]                                        # (n, regions, Kin^2)
                w = torch.log_softmax(self.sum_weights[lvl - 1][group, offset], dim=2)
                offset = torch.argmax(xin + w, dim=2)
                kin2 = xin.shape[2]
        mode = self.params["loc"] if self.kind == "gaussian" else (torch.sigmoid(self.params["logits"]) >= 0.5).to(x.dtype)
        out = x.clone()
        picked = mode[group, offset]                                                     # (n, 2^depth, dim)
        for r in range(group.shape[1]):
            for b in range(n):
                g = int(group[b, r])
                feats = list(self.leaf_regions[g])
                vals = picked[b, r, :len(feats)]
                cur = out[b, feats]
                out[b, feats] = torch.where(torch.isnan(cur), vals.to(out.dtype), cur)
        return out

    def log_prob_chunked(self, x: torch.Tensor, chunk: int = 1024) -> torch.Tensor:
        """The leaf temporary is (B,G0,K,dim) floats (501 KB/sample at D=784,R=16,K=10) -> chunk the batch."""
        return torch.cat([self.log_prob(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)], 0)

    # ------------------------------------------------------------------------------------------
    # Gradients / EM statistics by autograd over the restatement
    # ------------------------------------------------------------------------------------------
    def grads(self, x: torch.Tensor, grad_out: torch.Tensor, wrt_x: bool = True, clean_nan: bool = False):
        """d(sum(grad_out*log_prob))/d{x, leaf params, sum weights, root weight} via autograd."""
        prev_clean, self.clean_nan = self.clean_nan, clean_nan
        leaves = {k: v.clone().requires_grad_(True) for k, v in self.params.items()}
        sums = [w.clone().requires_grad_(True) for w in self.sum_weights]
        root = self.root_weight.clone().requires_grad_(True)
        xx = x.clone().requires_grad_(wrt_x)
        saved = (self.params, self.sum_weights, self.root_weight)
        self.params, self.sum_weights, self.root_weight = leaves, sums, root
        try:
            with torch.enable_grad():
                out = self.log_prob(xx)
                (out * grad_out).sum().backward()
        finally:
            self.params, self.sum_weights, self.root_weight = saved
            self.clean_nan = prev_clean
        res = {"out": out.detach(), "root": root.grad, "sums": [w.grad for w in sums]}
        res.update({k: v.grad for k, v in leaves.items()})
        if wrt_x:
            res["x"] = xx.grad
        return res

    def em_statistics(self, x: torch.Tensor):
        """E-step sufficient statistics of one batch (EM extension).

        Node-graph semantics being followed: stats = exp(child_ll - root_ll + log-grad)
        (deeprob/spn/learning/em.py:99-107, deeprob/spn/algorithms/gradient.py:48-55), i.e. the
        derivative of sum_b LL_b w.r.t. each log-weight (sum nodes, structure/node.py:100-111) and
        w.r.t. each leaf log-density (leaves, structure/leaf.py:167-174,536-545).  Obtained here by
        autograd with the log-softmax weights and the leaf LLs injected as differentiable tensors.
        Returns dict: ll_sum, n, sum_counts[l] (P,O,Kin), root_counts (C,P*Kin), s0 (G0,K), s1, s2 (G0,K,dim).
        """
        with torch.enable_grad():
            logw = [torch.log_softmax(w, 2).requires_grad_(True) for w in self.sum_weights]
            logr = torch.log_softmax(self.root_weight, 1).requires_grad_(True)
            leaf = self.leaf(x).detach().requires_grad_(True)
            h = leaf
            for lvl in range(self.depth):
                h = product_layer(h)
                if lvl < self.depth - 1:
                    h = torch.logsumexp(h.unsqueeze(2) + logw[lvl], dim=3)
            out = torch.logsumexp(h.flatten(1).unsqueeze(1) + logr, dim=2)
            out.sum().backward()
        post = leaf.grad                                               # (B,G0,K) posterior of each leaf
        v = x[:, self.mask]                                            # (B,G0,dim)
        ok = ~torch.isnan(v)
        if self.pad_mask is not None:
            ok = ok & ~self.pad_mask[:, 0, :].unsqueeze(0)
        v0 = torch.where(ok, v, torch.zeros_like(v))
        okf = ok.to(post.dtype)
        s0 = torch.einsum("bgk,bgd->gkd", post, okf)
        s1 = torch.einsum("bgk,bgd->gkd", post, v0)
        s2 = torch.einsum("bgk,bgd->gkd", post, v0 * v0)
        return {"ll_sum": out.detach().sum(), "n": x.shape[0], "sum_counts": [w.grad for w in logw],
                "root_counts": logr.grad, "s0": s0, "s1": s1, "s2": s2}
