"""Builds lib/libqwen3.so (host C + sm_100a CUDA) in-tree with make/nvcc/gcc."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libqwen3.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (-gencode arch=compute_100a,code=sm_100a -lineinfo)
    and link libqwen3.so. nvcc cross-compiles without a GPU. Returns the library path."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode:
        print(out.stdout)
    if out.returncode:
        raise RuntimeError("building libqwen3.so failed")
    return LIB


def lib_path() -> str:
    """Path of the built library; raises if it has not been built (no fallback exists)."""
    override = os.environ.get("QWEN3_LIB_PATH")  # debug: compare kernel variants built side by side
    if override:
        return override
    if not os.path.exists(LIB):
        raise FileNotFoundError(
            f"{LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the product has no CPU or PyTorch fallback)")
    return LIB
