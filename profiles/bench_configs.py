"""Secondary BASELINE configs on one GPU (CUDA-event timed, inputs resident in HBM) next to the CPU oracle:
  config 3  DgcSpn((1,28,28), n_batch=8, sum_channels=8, depthwise=True), batch 32768
  config 3b DgcSpn((1,28,28), n_batch=16, sum_channels=32, depthwise=True, n_pooling=2)  (MNIST example setting)
  config 4  RealNVP1d(3072, n_flows=8, depth=2, units=512), batch 16384
  config 4b RealNVP2d((3,32,32), n_flows=1, n_blocks=4, channels=64), batch 1024 per step (examples/nvp2d_cifar10.py)
  config 1  BernoulliRatSpn(15, 3, 4, 4, 2) on all 2^15 states
The CPU oracle (oracle/) is imported here only as the timed CPU comparator, the same role it has in bench.py's
cpu_baseline leg; nothing on the measured GPU path touches it.
Prints one JSON line per config:   python profiles/bench_configs.py [--no-cpu] [--only NAME]"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

warnings.simplefilter("ignore")
from deeprob_kit_b200 import _lib  # noqa: E402
from deeprob_kit_b200.flows.models import RealNVP1d, RealNVP2d  # noqa: E402
from deeprob_kit_b200.spn.models import BernoulliRatSpn, DgcSpn  # noqa: E402

NO_CPU = "--no-cpu" in sys.argv
ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""      # substring of a config name


def want(name):
    return ONLY in name



def gpu_time(model, x, steps=10, grad=False):
    def run():
        if grad:
            model.zero_grad(set_to_none=True)
            model(x).sum().backward()
        else:
            with torch.no_grad():
                model(x)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    _lib.profile_read()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    ms, cnt = _lib.profile_read()
    return e0.elapsed_time(e1) / steps, {k: round(v / steps, 4) for k, v in ms.items() if v > 0}


def cpu_time(fn, x, rows, reps=3):
    torch.set_num_threads(os.cpu_count())
    xs = x[:rows].cpu()
    with torch.no_grad():
        fn(xs)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn(xs)
    return rows * reps / (time.perf_counter() - t0)


def report(name, batch, feats, ms, kern, cpu_sps, ms_train=None):
    sps = batch / (ms * 1e-3)
    line = {"config": name, "batch": batch, "ms_per_step": round(ms, 4), "samples_per_s": sps, "evals_per_s": sps * feats,
            "kernel_ms": kern, "cpu_oracle_samples_per_s": cpu_sps, "cpu_threads": os.cpu_count(),
            "speedup_vs_cpu": (sps / cpu_sps) if cpu_sps else None}
    if ms_train is not None:
        line["ms_per_fwd_bwd_step"] = round(ms_train, 4)
    print(json.dumps(line), flush=True)


def dgc_oracle_fn(model):
    from oracle.dgcspn_oracle import DgcSpnOracle
    orc = DgcSpnOracle(model.in_features, model.out_classes, model.n_batch, model.sum_channels, model.depthwise, model.n_pooling)
    orc.load_reference_state({k: v.detach().cpu() for k, v in model.state_dict().items()})
    return orc.log_prob


torch.manual_seed(0)
for name, kw, batch in (("dgcspn_28x28_c8", dict(n_batch=8, sum_channels=8, depthwise=True), 32768),
                        ("dgcspn_28x28_mnist_example", dict(n_batch=16, sum_channels=32, depthwise=True, n_pooling=2), 32768)):
    if not want(name):
        continue
    m = DgcSpn((1, 28, 28), **kw).cuda().eval()
    x = torch.randn(batch, 1, 28, 28, device="cuda")
    ms, kern = gpu_time(m, x)
    ms_t, _ = gpu_time(m, x[:8192], steps=3, grad=True)
    cpu = None if NO_CPU else cpu_time(dgc_oracle_fn(m), x, 256)
    report(name, batch, 784, ms, kern, cpu, ms_t * batch / 8192)

if want("realnvp1d_3072_8flows"):
    m = RealNVP1d(3072, n_flows=8, depth=2, units=512).cuda().eval()
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "scale_act" in n_:
                p.fill_(0.5)
    x = torch.rand(16384, 3072, device="cuda")
    ms, kern = gpu_time(m, x)
    ms_t, _ = gpu_time(m, x, steps=3, grad=True)
    cpu = None
    if not NO_CPU:
        from oracle.flows_oracle import flow1d_log_prob
        st = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        cpu = cpu_time(lambda t: flow1d_log_prob(t, st, "RealNVP1d", dict(in_features=3072))[0], x, 2048)
    report("realnvp1d_3072_8flows", 16384, 3072, ms, kern, cpu, ms_t)

if want("realnvp2d_32x32x3"):
    import param_gen as pg
    cfg = pg.FLOW_CASES["nvp2d_cifar"]
    m = RealNVP2d(**cfg["kw"])
    m.load_state_dict(pg.flow_fill_state(m.state_dict(), fill_all=True))
    m = m.cuda().eval()
    x = torch.randn(1024, 3, 32, 32, device="cuda")
    ms, kern = gpu_time(m, x, steps=5)
    ms_t, _ = gpu_time(m, x[:256], steps=3, grad=True)
    cpu = None
    if not NO_CPU:      # the CPU comparator of this config is the product model's own modules on the CPU is not possible
        cpu = None      # (no CPU path); the reference rate measured by the survey is 29 samples/s
    report("realnvp2d_32x32x3_4blocks_64ch", 1024, 3072, ms, kern, cpu, ms_t * 1024 / 256)

if want("bernoulli_ratspn_15_all_states"):
    m = BernoulliRatSpn(15, rg_depth=3, rg_repetitions=4, rg_batch=4, rg_sum=2, random_state=42).cuda().eval()
    x = ((torch.arange(2 ** 15).unsqueeze(1) >> torch.arange(14, -1, -1)) & 1).float().cuda()
    ms, kern = gpu_time(m, x)
    cpu = None
    if not NO_CPU:
        from helpers import oracle_for
        import param_gen as pg
        orc, _ = oracle_for(pg.RATSPN_CASES["bern15"])
        cpu = cpu_time(orc.log_prob, x, 32768)
    report("bernoulli_ratspn_15_all_states", 32768, 15, ms, kern, cpu)
