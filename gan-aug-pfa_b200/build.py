"""Build libgap_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library links against the static CUDA runtime only (no libtorch, no libcuda: the one driver
entry point it needs, cuTensorMapEncodeTiled, is resolved at run time), so it cross-compiles on a
box without a GPU and travels to the GPU box as a plain .so.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB_DIR = HERE / "lib"
LIB_PATH = LIB_DIR / "libgap_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [HERE.parent / "include" / "gap_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library.  Rebuilds only when sources changed."""
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "libgap_b200.stamp"
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = obj_dir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src.name}\n{out.decode()}\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libgap_b200.so")
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs),
           "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.run(cmd, check=True)
    stamp.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
