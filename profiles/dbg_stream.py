"""Dev tool: per-sample error pattern of the streaming leaf kernel for one test case / tile height / grid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import param_gen as pg
from helpers import oracle_for, product_model
import test_ratspn_stream_gpu as T
name, mt, grid = sys.argv[1], sys.argv[2], sys.argv[3]
os.environ.update({"DPK_LEAF_STREAM": "1", "DPK_TREE_MMA": "1", "DPK_STREAM_MT": mt, "DPK_STREAM_GRID": grid})
cfg = T.CASES[name]
model = product_model(cfg, "cuda:0", scale_grad=False)
orc, _ = oracle_for(cfg)
x, _ = pg.ratspn_inputs(cfg)
for rep in range(2):
    out = model(x.cuda()).cpu()
    ref = orc.log_prob(x)
    err = ((out - ref).abs() / ref.abs().clamp_min(1.0)).max(dim=1).values
    bad = (err > 1e-4).nonzero().flatten().tolist()
    print("DBG", name, mt, grid, "rep", rep, "bad", len(bad), "of", len(err), "first", bad[:8], "last", bad[-8:], "maxerr", float(err.max()))
    if bad:
        print("DBG out", out[bad[0]].tolist()[:3], "ref", ref[bad[0]].tolist()[:3])
