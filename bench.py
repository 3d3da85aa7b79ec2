#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: RAT-SPN log-likelihood evaluations per second.

Metric (BASELINE.json): log-likelihood evals/sec counted as batch x D, RAT-SPN D=784
(GaussianRatSpn depth 3, 16 repetitions, K = O = 10) at batch 65536 per GPU, fp32 results.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one JSON line on rank 0)
  python bench.py --impl reference ...                            the reference algorithm on the host CPUs
  torchrun --nproc-per-node N bench.py --gpus N ...               N ranks, batch-sharded (weak scaling)

A "step" = one log_prob() pass over one batch of 65536 x 784 synthetic N(0,1) samples per GPU.  The inference
path shards along the batch and has no exchange step, so the timed region contains no collective.

Timed regions (all bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks):
  value      K steps replayed from ONE captured CUDA graph of the step (inputs resident in HBM, kernel profiling
             off; `graph: false` in the line if capture was not possible and the loop ran eagerly)
  sustained  the same graph replayed back to back for >= 3 s with its own clock record
  e2e        the public host-batch API (pinned host x -> log-likelihoods on the host; H2D + kernels + D2H)
  em_step    BASELINE config 5: one batch-EM step (E-step statistics + ONE all-reduce of the flat sufficient
             statistics + M-step) per rank on 65536 samples, every N
  hbm_bound_check  the same kernels at a structure where HBM is the binding roofline (R*K = 8)
Per-kernel-category times (`kernel_ms_per_step`, the roofline's kernel_ms) come from a separate eager pass with
CUDA-event bracketing of every launch group (dpk_profile_enable), after the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

D, DEPTH, REPS, K, O, C = 784, 3, 16, 10, 10, 1
METRIC = "log-likelihood evals/sec (batch x D), RAT-SPN D=784"
UNIT = "evals/s"
WORKLOAD = "GaussianRatSpn(784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10) log_prob, batch 65536/GPU"
ALGO_BYTES_PER_SAMPLE = 4 * D + 4 * C          # x row read + LL written
ALGO_FMA_PER_SAMPLE = 348480                    # one FMA per model parameter
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")   # written by profiles/ncu_traffic.py from an ncu capture


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / em_step / hbm_bound_check / train_step")
    ap.add_argument("--sustain-s", type=float, default=3.0)
    return ap.parse_args()


def config_dict(world, batch):
    """Identical for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": batch,
            "l2": "input batch 205 MB > 126 MB L2, re-read from HBM every step",
            "parallelism": ("batch-sharded x%d, no collective on the inference path" % world) if world > 1 else "single GPU"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs from warm-up to the end of the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.tag = index, [], False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.ok = True
        except Exception:  # noqa: BLE001
            pass

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                rs = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                try:
                    pw = self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:  # noqa: BLE001
                    pw = None
                self.samples.append((self.tag, sm, rs, pw))
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.004 if self.tag == "timed" else 0.02)

    def mem_mhz(self):
        try:
            return self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_MEM) if self.ok else None
        except Exception:  # noqa: BLE001
            return None

    def summary(self, tag):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        use = [s for s in self.samples if s[0] == tag] or [s for s in self.samples if s[0] is not None] or self.samples
        clocks = sorted(s[1] for s in use)
        bits = 0
        for s in use:
            bits |= s[2]
        try:
            mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            mx = None
        reasons = [n for b, n in self.REASONS.items() if bits & b and n != "gpu_idle"]
        pw = [s[3] for s in use if s[3] is not None]
        return {"sm_mhz": clocks[len(clocks) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(use),
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def build_oracle():
    from oracle.ratspn_oracle import RatSpnOracle
    from helpers import oracle_state
    cfg = dict(kind="gaussian", in_features=D, rg_depth=DEPTH, rg_repetitions=REPS, rg_batch=K, rg_sum=O,
               out_classes=C, optimize_scale=False)
    orc = RatSpnOracle(D, "gaussian", DEPTH, REPS, K, O, C, 42)
    orc.load_reference_state(oracle_state(orc, cfg))
    return orc


def time_oracle(steps, warmup, rows=512):
    """Each step = one `rows`-row chunk of the batch (the reference cannot hold the (B,128,10,98) temporary of a
    65536 batch: 32.9 GB -- BASELINE.md 4.3; its throughput is flat in the chunk size), all host threads."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = build_oracle()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(rows, D, generator=g)
    with torch.no_grad():
        for _ in range(warmup):
            orc.log_prob(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.log_prob(x)
        dt = time.perf_counter() - t0
    return rows * D * steps / dt, dt / steps * 1e3, cores, "%d steps of a %d-row chunk of the batch, %d threads" % (steps, rows, cores)


def run_reference(args, rank, world):
    if rank != 0:
        return
    val, ms, cores, sample = time_oracle(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic N(0,1), random-init parameters",
        "config": config_dict(max(world, args.gpus), args.batch),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + "; oracle port of the reference op sequence (oracle/ratspn_oracle.py)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def git_sha():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                              timeout=5).stdout.strip() or None
    except Exception:  # noqa: BLE001
        return None


def load_traffic():
    """Per-launch DRAM bytes of our kernels from the committed ncu capture (profiles/ncu_traffic.py)."""
    try:
        return json.load(open(TRAFFIC_FILE))
    except Exception:  # noqa: BLE001
        return None


class GraphStep:
    """One step captured in a CUDA graph (all launches of the C ABI are stream-ordered and the workspace is
    cached, so the step is replayable); falls back to the eager call when capture fails."""

    def __init__(self, fn, dev):
        self.fn, self.graph, self.out = fn, None, None
        if os.environ.get("DPK_BENCH_GRAPH", "1") == "0":
            return
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = fn()
            self.graph = g
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write("bench.py: CUDA graph capture failed (%s); timing the eager loop\n" % exc)
            self.graph = None
            torch.cuda.synchronize(dev)

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
            return self.out
        return self.fn()


def timed_loop(step, n, barrier):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    from deeprob_kit_b200 import _lib
    from deeprob_kit_b200.spn import em
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    from deeprob_kit_b200.spn.streaming import log_prob_host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.lib()
    torch.manual_seed(0)
    model = GaussianRatSpn(D, rg_depth=DEPTH, rg_repetitions=REPS, rg_batch=K, rg_sum=O, random_state=42).eval().to(dev)
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, D, device=dev, generator=g)
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxr(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    extras = {}
    with torch.no_grad():
        sampler.start()
        for _ in range(warmup):
            model(x)
        step = GraphStep(lambda: model(x), dev)
        for _ in range(warmup):
            step()
        # ---- the headline region --------------------------------------------------------------
        sampler.tag = "timed"
        elapsed_ms = timed_loop(step, args.steps, barrier)
        sampler.tag = None

        # ---- per-category kernel times: separate eager pass, CUDA events around every launch group
        prof_steps = max(3, min(args.steps, 20))
        model(x)
        barrier()
        _lib.profile_read()                      # reset counters
        _lib.profile_enable(True)
        for _ in range(prof_steps):
            model(x)
        barrier()
        _lib.profile_enable(False)
        prof_ms, launches = _lib.profile_read()

        # ---- sustained: seconds-long back-to-back loop with its own clock record ---------------
        if not args.no_extras and args.sustain_s > 0:
            n_sus = max(args.steps, int(args.sustain_s * 1e3 / max(elapsed_ms / args.steps, 1e-3)) + 1)
            sampler.tag = "sustained"
            sus_ms = timed_loop(step, n_sus, barrier)
            sampler.tag = None
            extras["sustained_raw"] = (sus_ms, n_sus)

        # ---- end to end through the host-batch API: pinned host x -> LL on the host -------------
        xh = torch.empty(B, D, dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        oh = torch.empty(B, C, dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(2):
            log_prob_host(model, xh, out_host=oh)
        sampler.tag = "e2e"
        e2e_ms = timed_loop(lambda: log_prob_host(model, xh, out_host=oh), e2e_steps, barrier)
        sampler.tag = None

    # ---- BASELINE config 5: batch-EM step with one all-reduce of the sufficient statistics -----
    if not args.no_extras:
        try:
            extras["em_step"] = bench_em(em, model, x, dev, world, barrier, maxr, dist)
        except Exception as exc:  # noqa: BLE001
            extras["em_step"] = {"error": str(exc)[:300]}
        try:
            extras["train_step"] = bench_train(x, dev, barrier, maxr)
        except Exception as exc:  # noqa: BLE001
            extras["train_step"] = {"error": str(exc)[:300]}
        if rank == 0:
            try:
                extras["hbm_bound_check"] = bench_hbm_bound(dev, B)
            except Exception as exc:  # noqa: BLE001
                extras["hbm_bound_check"] = {"error": str(exc)[:300]}
    sampler.stop_flag = True

    # per-rank view of the timed region (diagnostic: which rank sets the max) before the max-over-ranks
    rank_info = [elapsed_ms / args.steps] + [prof_ms.get(k, 0.0) / prof_steps for k in ("ratspn_leaf_mma", "ratspn_einsum")]
    mclk = sampler.mem_mhz()
    rank_info.append(float(mclk or 0))
    if world > 1:
        gathered = [torch.zeros(len(rank_info), device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor(rank_info, device=dev, dtype=torch.float64))
        per_rank = [[round(float(v), 4) for v in g] for g in gathered]
    else:
        per_rank = [[round(float(v), 4) for v in rank_info]]
    elapsed_ms, e2e_ms = maxr([elapsed_ms, e2e_ms])
    if "sustained_raw" in extras:
        sus_ms, n_sus = extras.pop("sustained_raw")
        sus_ms = maxr([sus_ms])[0]
    else:
        sus_ms = n_sus = None
    if rank != 0:
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    ms_step = elapsed_ms / args.steps
    value = world * B * D / (ms_step * 1e-3)
    cats = {k: round(v / prof_steps, 4) for k, v in prof_ms.items() if v > 0}
    per_step_launches = int(sum(launches.values()) // prof_steps)
    clk = sampler.summary("timed")
    sm_mhz = clk.get("sm_mhz") or 1965.0
    fp32_peak_tf = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    burst_tf = float(peaks.get("bf16_tflops", 1590.0))
    leaf_fma = D * K * REPS                       # x*mu products (the unit-scale expansion)
    mma = launches.get("ratspn_leaf_mma", 0) > 0
    leaf_cat = "ratspn_leaf_mma" if mma else "ratspn_leaf"
    leaf_ms = prof_ms[leaf_cat] / prof_steps     # the dominant kernel: one launch per step
    conv = mma and launches.get("ratspn_leaf_mma_prep", 0) == 0     # the GEMM converts its inputs itself: no PREP launch
    kname = ("ratspn_leaf_mma_kernel<conv>" if conv else "ratspn_leaf_mma_kernel<main>") if mma else "ratspn_leaf_kernel"
    algo_gbs = ALGO_BYTES_PER_SAMPLE * B / (leaf_ms * 1e-3) / 1e9
    traffic = load_traffic()
    t_kernel = None
    if traffic and B == 65536:
        t_kernel = (traffic.get("kernels", {}).get(kname) or {}).get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (leaf products as 3-pass hi/lo fp16 on tcgen05, fp32 accumulate)" if mma else "f32",
        "data": "synthetic N(0,1), random-init parameters",
        "config": config_dict(world, B),
        "samples_per_s": value / D,
        "graph": step.graph is not None,
        "clocks": clk,
        "e2e": {"value": world * B * D / (e2e_ms / e2e_steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * C * 4, "steps": e2e_steps,
                "api": "deeprob_kit_b200.spn.streaming.log_prob_host (pinned host in/out, copy stream + compute stream, 3-buffer ring of 8 192-row chunks, "
                       "returns after the last D2H copy landed)"},
        "gpu_launches": per_step_launches * args.steps,
        "gpu_launches_per_step": per_step_launches,
        "roofline": {"bound": "hbm", "kernel": kname,
                     "achieved": algo_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": algo_gbs / hbm_peak,
                     "traffic": t_kernel,
                     "traffic_source": ("profiles/traffic.json: ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                        "captured at git %s" % traffic.get("git_sha")) if t_kernel else None,
                     "step_traffic": traffic.get("step_dram_bytes") if traffic and B == 65536 else None,
                     "peak_source": peak_src, "kernel_ms": leaf_ms,
                     "whole_step_frac": ALGO_BYTES_PER_SAMPLE * B / (ms_step * 1e-3) / 1e9 / hbm_peak,
                     "note": "the path is compute bound at this config (220 flop/B, SURVEY.md 8d): the binding roofline is "
                             "roofline_tensor; hbm_bound_check is the config where HBM binds"},
        "kernel_ms_per_step": cats,
        "per_rank": {"columns": ["ms_per_step", "leaf_kernel_ms", "tree_kernel_ms", "mem_mhz_after"], "rows": per_rank},
        "kernel_ms_source": "separate eager pass of %d steps with CUDA events around every launch group" % prof_steps,
    }
    if mma:
        # dense tensor-core work issued by the main GEMM: 3 passes x (B x 784 x 1280)
        dense = 2.0 * B * D * 3 * REPS * (1 << DEPTH) * K
        useful = 2.0 * leaf_fma * B
        line["roofline_tensor"] = {"bound": "tensor", "kernel": kname,
                                   "useful_tflops": useful / (leaf_ms * 1e-3) / 1e12,
                                   "useful_frac": useful / (leaf_ms * 1e-3) / 1e12 / burst_tf,
                                   "issued_tflops": dense / (leaf_ms * 1e-3) / 1e12,
                                   "issued_frac": dense / (leaf_ms * 1e-3) / 1e12 / burst_tf,
                                   "peak": burst_tf, "unit": "TFLOP/s",
                                   "peak_source": "measured cuBLAS bf16, burst (MEASURED_PEAKS.json): the kernel is timed in a short loop",
                                   "note": "useful = one multiply-add per (sample, feature, channel, repetition); issued = dense fp16 "
                                           "MMA flops incl. the 8x structurally-zero blocks of the region structure and the 3-pass hi/lo split"}
    else:
        line["roofline_fp32"] = {"kernel": kname, "achieved_tflops": 4 * leaf_fma * B / (leaf_ms * 1e-3) / 1e12,
                                 "peak_tflops_at_observed_clock": fp32_peak_tf,
                                 "frac": 4 * leaf_fma * B / (leaf_ms * 1e-3) / 1e12 / fp32_peak_tf}
    line["whole_step_tflops_fp32_equiv"] = 2 * ALGO_FMA_PER_SAMPLE * B / (ms_step * 1e-3) / 1e12
    if sus_ms is not None:
        line["sustained"] = {"ms_per_step": sus_ms / n_sus, "steps": n_sus, "seconds": sus_ms * 1e-3,
                             "value": world * B * D / (sus_ms / n_sus * 1e-3), "unit": UNIT,
                             "clocks": sampler.summary("sustained")}
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        v, _, cores, sample = time_oracle(16, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)


def bench_em(em, infer_model, x, dev, world, barrier, maxr, dist, steps=10, warmup=3):
    """BASELINE config 5 (model of config 2, 65536 samples per GPU): E-step forward + backward-for-posteriors,
    ONE all-reduce of the flat statistics vector (deeprob/spn/learning/em.py:84-107 semantics), M-step."""
    import copy
    model = copy.deepcopy(infer_model).train()
    timing = {}
    for _ in range(warmup):
        em.em_step(model, x, 0.5, timing=timing)
    barrier()
    timing.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lls = [em.em_step(model, x, 0.5, timing=timing) for _ in range(steps)]
    e1.record()
    barrier()
    total = e0.elapsed_time(e1)
    ar = sum(a.elapsed_time(b) for a, b in timing.get("allreduce", [])) / steps if timing.get("allreduce") else 0.0
    total, ar = maxr([total, ar])
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((hi - lo).abs() <= 1e-9 * hi.abs())
    n = x.shape[0]
    return {"ms_per_em_step": total / steps, "allreduce_ms": ar, "allreduce_bytes": timing.get("bytes"),
            "samples_per_s": world * n / (total / steps * 1e-3), "batch_per_gpu": n, "steps": steps,
            "replicas_identical": same, "mean_ll_first_last": [round(lls[0], 3), round(lls[-1], 3)],
            "collective": "one all_reduce(sum, fp32) of [sum LL, N, sum/root counts, S0, S1] per step (NCCL)" if world > 1
                          else "none (single rank)"}


def bench_train(x, dev, barrier, maxr, steps=10, warmup=3):
    """Gradient step of the config-2 structure (learnable scale): forward + backward (all parameter grads and
    d/dx, as a flow with a RatSpn base needs) through the autograd node of the fused path."""
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    torch.manual_seed(0)
    model = GaussianRatSpn(D, rg_depth=DEPTH, rg_repetitions=REPS, rg_batch=K, rg_sum=O, random_state=42,
                           optimize_scale=True).to(dev).train()
    out = {}
    for name, need_x in (("params", False), ("params_and_x", True)):
        xi = x.detach().clone().requires_grad_(need_x)

        def one():
            model.zero_grad(set_to_none=True)
            if xi.grad is not None:
                xi.grad = None
            loss = model.loss(model(xi))
            loss.backward()
        with torch.enable_grad():
            for _ in range(warmup):
                one()
            ms = timed_loop(one, steps, barrier)
        out["ms_fwd_bwd_" + name] = maxr([ms])[0] / steps
    out["batch_per_gpu"] = x.shape[0]
    return out


def bench_hbm_bound(dev, B, steps=50):
    """north_star's '>= 60 % of HBM' can only bind where R*K is small (R*K <~ 11): the same kernels at
    GaussianRatSpn(784, depth 3, R=1, K=8), batch 65536, reported against the algorithmic bytes (x read once + LL)."""
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    torch.manual_seed(0)
    model = GaussianRatSpn(D, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8, random_state=42).eval().to(dev)
    x = torch.randn(B, D, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    with torch.no_grad():
        for _ in range(3):
            model(x)
        step = GraphStep(lambda: model(x), dev)
        for _ in range(3):
            step()
        ms = timed_loop(step, steps, lambda: torch.cuda.synchronize(dev)) / steps
    gbs = ALGO_BYTES_PER_SAMPLE * B / (ms * 1e-3) / 1e9
    return {"workload": "GaussianRatSpn(784, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8) log_prob, batch %d" % B,
            "ms_per_step": ms, "algorithmic_bytes": ALGO_BYTES_PER_SAMPLE * B, "achieved_gbs": gbs, "peak_gbs": hbm_peak,
            "frac": gbs / hbm_peak, "graph": step.graph is not None}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % args.gpus)
    try:
        run_b200(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
