"""Recipe for oracle/_ref: an UNMODIFIED copy of the reference's Python packages, taken from the read-only
checkout so that the reference's own Numba path can be timed on the GPU box (which has no /root/reference).

    python oracle/make_ref.py            # /root/reference/{bayesopt,examples} -> oracle/_ref/

Test / measurement infrastructure only.  ``oracle/_ref/`` is git-ignored (no reference source enters the
history) but NOT gpurun-ignored, so it travels with the repository snapshot exactly like the built ``.so``.
Nothing in the product package imports it; its only users are ``bench.py --impl reference``, the
``cpu_baseline`` leg of ``bench.py`` and ``tests/golden/make_golden.py`` (which may equally import
/root/reference directly).  ``__graft_entry__.build()`` runs this when the checkout is present.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ("bayesopt", "examples")  # numba_kernels / acquisition / pareto / loop, and the toy objectives of cfg1


def make_ref(source: str = "/root/reference", force: bool = False) -> str | None:
    """Copy the packages byte for byte; returns the destination, or None when the checkout is absent."""
    if not os.path.isdir(os.path.join(source, "bayesopt")):
        return DEST if os.path.isdir(os.path.join(DEST, "bayesopt")) else None
    os.makedirs(DEST, exist_ok=True)
    for pkg in PACKAGES:
        src, dst = os.path.join(source, pkg), os.path.join(DEST, pkg)
        if not os.path.isdir(src):
            continue
        if os.path.isdir(dst) and not force:
            cmp = filecmp.dircmp(src, dst, ignore=["__pycache__"])
            if not (cmp.left_only or cmp.diff_files or cmp.funny_files):
                continue
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DEST, "SOURCE.txt"), "w") as f:
        f.write(f"verbatim copy of {source}/{{{','.join(PACKAGES)}}} made by oracle/make_ref.py; not tracked by git\n")
    return DEST


def import_reference():
    """Import the copied reference (Numba mode).  Returns the modules, or raises ImportError with the reason.
    The reference prints a banner at import (config.py:97-100); it is sent to stderr so that stdout stays a
    single JSON line for bench.py."""
    import contextlib

    if not os.path.isdir(os.path.join(DEST, "bayesopt")):
        raise ImportError(f"{DEST}/bayesopt is missing (run oracle/make_ref.py where /root/reference exists)")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    sys.dont_write_bytecode = True
    with contextlib.redirect_stdout(sys.stderr):
        import numba  # noqa: F401  (ImportError here = the box has no Numba: the caller logs it and uses the port)
        from bayesopt import acquisition, numba_kernels, pareto
    return numba_kernels, acquisition, pareto


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
