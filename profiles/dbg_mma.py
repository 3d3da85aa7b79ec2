"""Diagnostic for the tensor-core leaf path: per-case error against the float64 oracle with a breakdown
by sample row / column so that a layout bug can be read off one GPU run.  Not a test, not a benchmark."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import param_gen as pg  # noqa: E402
from helpers import oracle_for, product_model  # noqa: E402
from test_ratspn_mma_gpu import CASES  # noqa: E402

DEV = "cuda:0"
names = sys.argv[1:] or ["bern16", "gauss36", "gauss784", "gauss_wide", "bern784"]
for name in names:
    cfg = CASES[name]
    os.environ["DPK_LEAF_MMA"] = "1"
    model = product_model(cfg, DEV, scale_grad=False)
    orc = oracle_for(cfg)[0].double()
    x, _ = pg.ratspn_inputs(cfg)
    ref = orc.leaf(x.double())
    t0 = time.time()
    leaf = model.base_layer(x.to(DEV))
    torch.cuda.synchronize()
    leaf = leaf.cpu().double()
    err = (leaf - ref).abs()
    B, G, K = err.shape
    print("== %s  B=%d G0=%d K=%d  max abs err %.3e  max |ref| %.1f  (%.2fs)" % (name, B, G, K, float(err.max()),
                                                                             float(ref.abs().max()), time.time() - t0))
    if float(err.max()) > 1e-3 or not torch.isfinite(leaf).all():
        flat = err.reshape(B, G * K)
        bad = ~(flat < 1e-3)
        print("   bad fraction %.4f; non-finite %d" % (float(bad.float().mean()), int((~torch.isfinite(leaf)).sum())))
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows: %d of %d, first %s last %s" % (len(rows), B, rows[:12].tolist(), rows[-4:].tolist()))
        print("   bad cols: %d of %d, first %s last %s" % (len(cols), G * K, cols[:12].tolist(), cols[-4:].tolist()))
        print("   sample  got %s" % leaf.reshape(B, -1)[0, :6].tolist())
        print("           ref %s" % ref.reshape(B, -1)[0, :6].tolist())
        r = int(rows[0]) if len(rows) else 0
        print("   row %d  got %s" % (r, leaf.reshape(B, -1)[r, :6].tolist()))
        print("           ref %s" % ref.reshape(B, -1)[r, :6].tolist())
    os.environ["DPK_LEAF_MMA"] = "0"
    leaf0 = model.base_layer(x.to(DEV)).cpu().double()
    print("   CUDA-core kernel max abs err %.3e" % float((leaf0 - ref).abs().max()))
