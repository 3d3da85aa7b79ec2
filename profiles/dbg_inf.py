"""Diagnostic: which rows of the out-of-range test disagree, on both leaf paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import param_gen as pg
from helpers import oracle_for, product_model
from test_ratspn_mma_gpu import CASES
cfg = CASES["gauss784"]
orc = oracle_for(cfg)[0].double()
x, _ = pg.ratspn_inputs(cfg)
x[3, 5] = 1.0e4
x[40, 700] = float("inf")
x[699, 0] = float("nan")
x[300:310, :] *= 300.0
ref = orc.log_prob(x.double())
for mode in ("1", "0"):
    os.environ["DPK_LEAF_MMA"] = mode
    model = product_model(cfg, "cuda:0", scale_grad=False)
    out = model(x.cuda()).cpu().double()
    err = ((out - ref).abs() / ref.abs().clamp_min(1.0)).flatten()
    bad = (~(err < 1e-4)).nonzero().flatten().tolist()
    print("MMA=%s bad rows %s" % (mode, bad))
    for r in bad[:6]:
        print("   row %d got %r ref %r" % (r, float(out[r, 0]), float(ref[r, 0])))
    leaf = model.base_layer(x.cuda()).cpu().double()
    rl = orc.leaf(x.double())
    e2 = ((leaf - rl).abs() / rl.abs().clamp_min(1.0)).flatten(1).max(1).values
    print("   leaf bad rows", (~(e2 < 1e-4)).nonzero().flatten().tolist()[:10])
