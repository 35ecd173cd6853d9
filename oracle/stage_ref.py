"""Build recipe for oracle/_ref/: the UNMODIFIED reference compiled to Python bytecode.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python, so "compiling it from the sources where they lie" means byte-compiling
/root/reference/{models,dataset,train_gan,train,evaluate,generate_synthetic_data}.py with this interpreter into
oracle/_ref/<name>.gapref (sourceless-module bytecode, the .pyc format under another extension).  No reference source text enters the repository: oracle/_ref/ is
git-ignored, but it is NOT gpurun-ignored, so the bytecode travels to the GPU box (same image, same interpreter),
where /root/reference does not exist.  Used by
  * tests/test_gpu_reference_scripts.py — runs the reference's real train_gan_one_epoch / train_one_epoch against the
    drop-in modules, and the reference's own modules as the checker;
  * bench.py --impl reference and the `library_baseline` leg — the reference itself as the CPU arm / the cuDNN bar.

    python oracle/stage_ref.py          (also called by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import py_compile
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
# not ".pyc": file synchronisers (the GPU-box snapshot among them) skip *.pyc as throw-away caches
EXT = ".gapref"
MODULES = ("models", "dataset", "train_gan", "train", "evaluate", "generate_synthetic_data")


def stage(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref holds bytecode for every module (freshly built or already there)."""
    if not (REF / "models.py").exists():
        return all((OUT / (m + EXT)).exists() for m in MODULES)
    OUT.mkdir(exist_ok=True)
    for m in MODULES:
        src, dst = REF / f"{m}.py", OUT / (m + EXT)
        # dfile: the path recorded in tracebacks; unchecked-hash pycs never look for the source file
        py_compile.compile(str(src), cfile=str(dst), dfile=f"<reference>/{m}.py", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print(f"staged {dst} ({dst.stat().st_size} bytes)")
    (OUT / "PYTHON").write_text(sys.version)
    return True


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("oracle/_ref ready" if ok else "reference tree absent and oracle/_ref incomplete")
