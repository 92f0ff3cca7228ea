"""Build the CUDA library in-tree: nvcc -> xsdba_b200/libxsdba_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "xsdba_b200.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc")))] + [
    os.path.join(os.path.dirname(HERE), "include", "xsdba_b200.h")]
OUT = os.path.join(HERE, "libxsdba_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]
# `--split-compile=0` (parallel back end, 2.5 min -> 1 min) is opt-in for development only: it partitions the module
# before optimisation, and the code generated for one kernel then depends on unrelated edits elsewhere -- the train
# kernel sits at the 64-register cap of a 1024-thread CTA, where such a perturbation flipped it between 0 and 100
# spill instructions (4.0 vs 4.9 ms per slab on the same box, profiles/README.md).
if os.environ.get("XSDBA_B200_SPLIT_COMPILE"):
    NVCC_FLAGS.append("--split-compile=0")


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: xsdba_b200 has no CPU fallback and needs the CUDA toolkit to build")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
