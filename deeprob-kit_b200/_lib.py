"""Build and bind libdeeprob_b200.so (the C ABI of include/deeprob_b200.h) with ctypes.

`build()` cross-compiles every csrc/*.cu for sm_100a with nvcc (works without a GPU); `lib()`
loads the shared object and fails loudly when it is missing -- there is deliberately no fallback.
"""
import ctypes
import os
import shutil
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIBDIR = os.path.join(_HERE, "lib")
_OBJDIR = os.path.join(_HERE, "build")
LIB_PATH = os.path.join(_LIBDIR, "libdeeprob_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]  # no --use_fast_math: fast intrinsics are chosen explicitly, kernel by kernel

MAX_LEVELS = 16

_lock = threading.Lock()
_lib = None


def _nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: cannot build libdeeprob_b200.so")
    return path


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src, *extra))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> lib/libdeeprob_b200.so (incremental). Returns the library path."""
    os.makedirs(_LIBDIR, exist_ok=True)
    os.makedirs(_OBJDIR, exist_ok=True)
    sources = sorted(f for f in os.listdir(_CSRC) if f.endswith(".cu"))
    headers = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(INCLUDE_DIR, "deeprob_b200.h"))
    nvcc = _nvcc()
    jobs = []
    objs = []
    for s in sources:
        src = os.path.join(_CSRC, s)
        obj = os.path.join(_OBJDIR, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            jobs.append([nvcc, *NVCC_FLAGS, "-I", INCLUDE_DIR, "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stdout + r.stderr))

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(LIB_PATH) or any(_newer(o, LIB_PATH) for o in objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs])
    return LIB_PATH


# ------------------------------------------------------------------------------------------------
# ctypes mirror of include/deeprob_b200.h
# ------------------------------------------------------------------------------------------------
c_i32, c_i64, c_u32, c_vp, c_sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t

F_SAVE_ACTIVATIONS = 1
F_TABLES_VALID = 2
LEAF_GAUSSIAN, LEAF_BERNOULLI = 0, 1


class RatSpnDesc(ctypes.Structure):
    _fields_ = [
        ("leaf_kind", c_i32), ("in_features", c_i32), ("depth", c_i32), ("repetitions", c_i32),
        ("leaf_channels", c_i32), ("sum_nodes", c_i32), ("out_classes", c_i32), ("dimension", c_i32),
        ("mask", c_vp), ("region_len", c_vp), ("leaf_p0", c_vp), ("leaf_p1", c_vp),
        ("sum_weight", c_vp * MAX_LEVELS), ("root_weight", c_vp),
    ]


class RatSpnGrads(ctypes.Structure):
    _fields_ = [
        ("grad_x", c_vp), ("leaf_p0", c_vp), ("leaf_p1", c_vp),
        ("sum_weight", c_vp * MAX_LEVELS), ("root_weight", c_vp),
    ]


class RatSpnEmStats(ctypes.Structure):
    _fields_ = [
        ("sum_counts", c_vp * MAX_LEVELS), ("root_counts", c_vp), ("s0", c_vp), ("s1", c_vp), ("s2", c_vp),
    ]


class RatSpnDropout(ctypes.Structure):
    _fields_ = [("in_rate", ctypes.c_float), ("sum_rate", ctypes.c_float), ("seed", ctypes.c_uint64)]


class DgcProductDesc(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in ("channels", "height", "width", "out_channels", "out_height", "out_width",
                                     "pad_top", "pad_left", "stride_h", "stride_w", "dilation_h", "dilation_w",
                                     "depthwise")]


class CouplingDesc(ctypes.Structure):
    _fields_ = [("batch", c_i64), ("features", c_i32), ("affine", c_i32), ("direction", c_i32), ("w_count", c_i32),
                ("w_inner", c_i32), ("x_stride", c_i64), ("z_stride", c_i64), ("inv_mask", c_vp), ("scale_weight", c_vp)]


# name -> (restype, argtypes); every symbol include/deeprob_b200.h declares must be listed here
# (tests/test_cabi.py cross-checks this table against the header and the built library).
SIGNATURES = {
    "dpk_abi_version": (ctypes.c_int, []),
    "dpk_last_error": (ctypes.c_char_p, []),
    "dpk_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "dpk_profile_read": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i64), c_i32]),
    "dpk_ratspn_workspace_bytes": (c_sz, [ctypes.POINTER(RatSpnDesc), c_i64, c_u32]),
    "dpk_ratspn_forward": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_vp, c_i64, c_vp, c_vp, c_sz, c_u32, c_vp]),
    "dpk_ratspn_backward": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_vp, c_i64, c_vp, c_vp,
                                           ctypes.POINTER(RatSpnGrads), c_vp, c_sz, c_vp]),
    "dpk_ratspn_em_statistics": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_vp, c_i64, c_vp,
                                                ctypes.POINTER(RatSpnEmStats), c_vp, c_sz, c_vp]),
    "dpk_dropout_draw": (c_u32, [ctypes.c_uint64, c_u32, ctypes.c_uint64]),
    "dpk_ratspn_dropout_workspace_bytes": (c_sz, [ctypes.POINTER(RatSpnDesc), c_i64]),
    "dpk_ratspn_forward_dropout": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_vp, c_i64, ctypes.POINTER(RatSpnDropout),
                                                  c_vp, c_vp, c_sz, c_vp]),
    "dpk_ratspn_backward_dropout": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_vp, c_i64, ctypes.POINTER(RatSpnDropout),
                                                   c_vp, c_vp, ctypes.POINTER(RatSpnGrads), c_vp, c_sz, c_vp]),
    "dpk_ratspn_mpe": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "dpk_ratspn_sample": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_i64, c_vp, ctypes.c_uint64, c_vp, c_vp, c_sz, c_vp]),
    "dpk_nll_loss": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "dpk_ratspn_leaf_forward": (ctypes.c_int, [ctypes.POINTER(RatSpnDesc), c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "dpk_outer_sum_forward": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "dpk_mixture_forward": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dpk_dgc_leaf_forward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dpk_dgc_leaf_backward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dpk_dgc_product_forward": (ctypes.c_int, [ctypes.POINTER(DgcProductDesc), c_vp, c_i64, c_vp, c_vp]),
    "dpk_dgc_product_backward": (ctypes.c_int, [ctypes.POINTER(DgcProductDesc), c_vp, c_i64, c_vp, c_vp]),
    "dpk_dgc_sum_forward": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dpk_dgc_sum_backward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dpk_dgc_root_forward": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "dpk_coupling_forward": (ctypes.c_int, [ctypes.POINTER(CouplingDesc), c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "dpk_coupling_forward_compact": (ctypes.c_int, [ctypes.POINTER(CouplingDesc), c_vp, c_vp, c_vp, c_i32, c_vp, c_vp,
                                                    ctypes.c_float, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dpk_coupling_backward": (ctypes.c_int, [ctypes.POINTER(CouplingDesc), c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64,
                                             c_vp, c_i64, c_vp, c_vp]),
    "dpk_feature_reduce": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp]),
    "dpk_feature_affine": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp]),
    "dpk_flow_preprocess_forward": (ctypes.c_int, [c_vp, c_vp, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dpk_flow_preprocess_backward": (ctypes.c_int, [c_vp, c_vp, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp, c_i64,
                                                    c_i32, c_vp]),
    "dpk_normal_prior_forward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dpk_normal_prior_backward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dpk_dgc_root_backward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dpk_dgc_prodsum_forward": (ctypes.c_int, [ctypes.POINTER(DgcProductDesc), c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "dpk_dgc_prodsum_backward": (ctypes.c_int, [ctypes.POINTER(DgcProductDesc), c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dpk_linear_workspace_bytes": (c_sz, [c_i64, c_i32, c_i32]),
    "dpk_linear_backward_workspace_bytes": (c_sz, [c_i64, c_i32, c_i32]),
    "dpk_linear_backward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "dpk_linear_forward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_sz, c_u32, c_vp]),
}


def lib() -> ctypes.CDLL:
    """The loaded shared library (built on demand when nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                try:
                    build()
                except Exception as exc:  # noqa: BLE001
                    raise RuntimeError(
                        "libdeeprob_b200.so is missing and could not be built (%s); "
                        "run `python -c 'import __graft_entry__ as g; g.build()'` -- there is no fallback path" % exc
                    ) from exc
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().dpk_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def require_cuda(t, what: str):
    """Product path guard: CUDA float32 tensors only, never a CPU fallback."""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s: deeprob_kit_b200 runs on CUDA tensors only (got %s); there is no CPU path"
                           % (what, getattr(t, "device", type(t))))
    return t


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


PROFILE_CATEGORIES = ["prep", "ratspn_leaf", "ratspn_einsum", "ratspn_root", "ratspn_bwd_einsum", "ratspn_bwd_leaf",
                      "finalize", "layers", "dgcspn_fwd", "dgcspn_bwd", "flow_fwd", "flow_bwd", "gemm",
                      "ratspn_leaf_mma", "ratspn_leaf_mma_prep", "c15"]


def profile_enable(on: bool) -> None:
    lib().dpk_profile_enable(1 if on else 0)


def profile_read():
    """({category: elapsed ms}, {category: kernel launches}) since the previous read."""
    n = len(PROFILE_CATEGORIES)
    ms = (ctypes.c_double * n)()
    cnt = (c_i64 * n)()
    check(lib().dpk_profile_read(ms, cnt, n), "dpk_profile_read")
    return ({PROFILE_CATEGORIES[i]: ms[i] for i in range(n)}, {PROFILE_CATEGORIES[i]: int(cnt[i]) for i in range(n)})
