#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: RAT-SPN log-likelihood evaluations per second.

Metric (BASELINE.json): log-likelihood evals/sec counted as batch x D, RAT-SPN D=784
(GaussianRatSpn depth 3, 16 repetitions, K = O = 10) at batch 65536 per GPU, fp32.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one JSON line on rank 0)
  python bench.py --impl reference ...                            the reference algorithm on the host CPUs
  torchrun --nproc-per-node N bench.py --gpus N ...               N ranks, batch-sharded (weak scaling)

A "step" = one log_prob() pass over one batch of 65536 x 784 synthetic N(0,1) samples per GPU
(+ one all-reduce of the summed log-likelihood when N > 1).  `value` has the inputs resident in
HBM; `e2e` goes through the public host-batch API with pinned host buffers (H2D + kernels + D2H
inside the timed region).  See DESIGN.md "Measurement" for the roofline accounting.
"""
import argparse
import json
import os
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
LEAF_MMA_TRAFFIC = 534.8e6                      # DRAM bytes of one launch of ratspn_leaf_mma_kernel<main> (ncu, profiles/leaf_mma_r1.txt)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs from warm-up to the end of the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.timed = index, [], False, False
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
                self.samples.append((self.timed, sm, rs))
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.005)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        use = [s for s in self.samples if s[0]] or self.samples
        clocks = sorted(s[1] for s in use)
        bits = 0
        for s in use:
            bits |= s[2]
        try:
            mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            mx = None
        reasons = [n for b, n in self.REASONS.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": clocks[len(clocks) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(use)}


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


def time_oracle(steps, warmup, rows=1024):
    """Each step = one `rows`-row chunk (the reference cannot hold the (B,128,10,98) temporary of a
    65536 batch: 32.9 GB -- BASELINE.md 4.3), all host threads."""
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


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    val, ms, cores, sample = time_oracle(steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic N(0,1)",
        "config": {"workload": WORKLOAD, "timed_as": "1024-row chunks on the host CPUs (oracle port of the reference ops)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    from deeprob_kit_b200 import _lib
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        # batch-sharded job: every rank evaluates its own shard, the path has no exchange step, so
        # there is no collective inside the timed region (config 5, profiles/bench_em.py, is the path with one)
        return model(x)

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    with torch.no_grad():
        sampler.start()
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        _lib.profile_read()                      # reset counters
        _lib.profile_enable(True)
        sampler.timed = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        sampler.timed = False
        elapsed_ms = e0.elapsed_time(e1)
        _lib.profile_enable(False)
        prof_ms, launches = _lib.profile_read()

        # ---- end to end through the host-batch API: pinned host x -> LL on the host -------------
        xh = torch.empty(B, D, dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        oh = torch.empty(B, C, dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(2):
            log_prob_host(model, xh, out_host=oh)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(e2e_steps):
            log_prob_host(model, xh, out_host=oh)
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1)
        sampler.stop_flag = True

    t = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])
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
    cats = {k: round(v / args.steps, 4) for k, v in prof_ms.items() if v > 0}
    clk = sampler.summary()
    sm_mhz = clk.get("sm_mhz") or 1965.0
    fp32_peak_tf = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    bf16_peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    leaf_fma = D * K * REPS                       # x*mu products (the unit-scale expansion; x^2 term is D*REPS more)
    mma = launches.get("ratspn_leaf_mma", 0) > 0
    leaf_cat = "ratspn_leaf_mma" if mma else "ratspn_leaf"
    leaf_ms = prof_ms[leaf_cat] / args.steps     # the dominant kernel: one launch per step
    # Algorithmic bytes per launch: every sample's row read once + its LL written once,
    # B samples per launch.
    algo_gbs = ALGO_BYTES_PER_SAMPLE * B / (leaf_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (leaf products as 3-pass hi/lo fp16 on tcgen05, fp32 accumulate)" if mma else "f32",
        "data": "synthetic N(0,1), random-init parameters",
        "config": {"workload": WORKLOAD, "samples_per_s": value / D, "batch_per_gpu": B,
                   "l2": "input batch 205 MB > 126 MB L2, re-read from HBM every step",
                   "parallelism": "batch-sharded x%d, no collective on the inference path" % world if world > 1 else "single GPU"},
        "clocks": clk,
        "e2e": {"value": world * B * D / (e2e_ms / e2e_steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * C * 4,
                "api": "deeprob_kit_b200.spn.streaming.log_prob_host (pinned host in/out, 2-stream chunk pipeline)"},
        "gpu_launches": int(sum(launches.values())),
        "roofline": {"bound": "hbm", "kernel": "ratspn_leaf_mma_kernel<main>" if mma else "ratspn_leaf_kernel",
                     "achieved": algo_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": algo_gbs / hbm_peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at B=65536, ncu --set full
                     # (profiles/leaf_mma_r1.txt resp. profiles/leaf_r1.txt): x images read + leaf activations written
                     "traffic": (LEAF_MMA_TRAFFIC if mma else 490.8e6) if B == 65536 else None,
                     "peak_source": peak_src, "kernel_ms": leaf_ms,
                     "note": "the path is compute bound (220 flop/B, SURVEY.md 8d): the binding roofline is roofline_tensor"},
        "kernel_ms_per_step": cats,
    }
    if mma:
        # dense tensor-core work issued by the main GEMM: 3 passes x (B x 784 x 1280)
        dense = 2.0 * B * D * 3 * REPS * (1 << DEPTH) * K
        line["roofline_tensor"] = {"bound": "tensor", "kernel": "ratspn_leaf_mma_kernel<main>",
                                   "achieved": dense / (leaf_ms * 1e-3) / 1e12, "peak": bf16_peak_tf, "unit": "TFLOP/s",
                                   "frac": dense / (leaf_ms * 1e-3) / 1e12 / bf16_peak_tf,
                                   "useful_tflops": 2.0 * leaf_fma * B / (leaf_ms * 1e-3) / 1e12,
                                   "peak_source": "measured cuBLAS bf16, sustained (MEASURED_PEAKS.json)",
                                   "note": "achieved = dense fp16 MMA flops issued, incl. the zero blocks of the region "
                                           "structure and the 3-pass hi/lo split; ncu: tensor pipe active 66% of elapsed"}
    else:
        line["roofline_fp32"] = {"kernel": "ratspn_leaf_kernel", "achieved_tflops": 4 * leaf_fma * B / (leaf_ms * 1e-3) / 1e12,
                                 "peak_tflops_at_observed_clock": fp32_peak_tf,
                                 "frac": 4 * leaf_fma * B / (leaf_ms * 1e-3) / 1e12 / fp32_peak_tf}
    line["whole_step_tflops_fp32_equiv"] = 2 * ALGO_FMA_PER_SAMPLE * B / (ms_step * 1e-3) / 1e12
    if world == 1 and not args.no_cpu_baseline:
        v, _, cores, sample = time_oracle(8, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
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
