"""CPU: libdeeprob_b200.so builds (nvcc cross-compile), loads, and exports every symbol the header declares."""
import ctypes
import os
import re

from deeprob_kit_b200 import _lib

HEADER = os.path.join(_lib.INCLUDE_DIR, "deeprob_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    names = declared_symbols()
    assert "dpk_ratspn_forward" in names and "dpk_ratspn_backward" in names
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    path = _lib.build()
    handle = ctypes.CDLL(path)
    for n in names:
        assert hasattr(handle, n), "missing export " + n


def test_version_and_error_convention():
    lib = _lib.lib()
    assert lib.dpk_abi_version() == 2
    # argument validation happens on the host before any CUDA call: usable without a GPU
    desc = _lib.RatSpnDesc()
    assert lib.dpk_ratspn_workspace_bytes(ctypes.byref(desc), 8, 0) == 0
    assert b"descriptor" in lib.dpk_last_error() or b"leaf" in lib.dpk_last_error()
    rc = lib.dpk_ratspn_forward(ctypes.byref(desc), None, 8, None, None, 0, 0, None)
    assert rc == -1


def test_library_is_sm100a_sass():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "--list-elf", _lib.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
