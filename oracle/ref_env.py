"""Make the upstream deeprob-kit importable in THIS container (test infrastructure only).

The reference at /root/reference imports matplotlib at package import
(deeprob/spn/models/__init__.py:3 -> deeprob/spn/structure/io.py:9) and matplotlib is not
installed here, so a two-file stub is placed ahead of it on sys.path (recipe: SURVEY.md 8c).

/root/reference does not exist on the GPU box: only `oracle/make_golden.py` and CPU-side
validation tests (skipped when the directory is absent) may call `enable()`.
"""
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("DEEPROB_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deeprob"))


def enable() -> None:
    """Put the reference and a matplotlib stub on sys.path (idempotent)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "deeprob" in sys.modules:
        return
    stub = os.path.join(tempfile.gettempdir(), "deeprob_oracle_stub")
    os.makedirs(os.path.join(stub, "matplotlib"), exist_ok=True)
    for name in ("__init__.py", "pyplot.py"):
        path = os.path.join(stub, "matplotlib", name)
        if not os.path.exists(path):
            open(path, "w").close()
    sys.dont_write_bytecode = True  # /root/reference is read-only
    try:
        import matplotlib  # noqa: F401  (a real one wins if present)
    except ImportError:
        sys.path.insert(0, stub)
    sys.path.insert(0, REFERENCE_ROOT)
