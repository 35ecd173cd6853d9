"""Import the staged reference (oracle/_ref/*.gapref, see stage_ref.py).  TEST INFRASTRUCTURE ONLY: tests/ and bench.py's
reference / library-baseline legs are the only callers.

`load(models=None)` returns a namespace with the reference's modules (`models`, `dataset`, `train_gan`, `train`).
With `models=<module>` the scripts are imported against THAT module under the name `models` — this is how the
reference's unchanged `train_gan.py` / `train.py` are made to run on the drop-in gan_aug_pfa_b200.models.
Import-time side effects of the scripts are neutralised the way SURVEY.md §4 describes: os.makedirs is a no-op while
they import (train_gan.py:37-39 creates /Users/mac/...), `optuna` and `matplotlib` are empty stub modules
(train.py:12, evaluate.py:8), tqdm is replaced by a pass-through iterator."""
from __future__ import annotations

import contextlib
import importlib.machinery
import importlib.util
import io
import os
import sys
import types
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"
EXT = ".gapref"


def available() -> bool:
    return (REF_DIR / ("models" + EXT)).exists() and (REF_DIR / ("train_gan" + EXT)).exists()


class _NoBar:
    def __init__(self, it, **kw):
        self.it = it

    def __iter__(self):
        return iter(self.it)

    def set_postfix(self, **kw):
        pass


def _import_pyc(name: str, alias: str):
    path = REF_DIR / (name + EXT)
    loader = importlib.machinery.SourcelessFileLoader(alias, str(path))
    spec = importlib.util.spec_from_loader(alias, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def load(models=None, tag: str = "ref") -> types.SimpleNamespace:
    """Import the staged scripts.  The modules are registered under private names (`_gapref_<tag>_*`); the bare names the
    scripts import (`models`, `dataset`) are bound only while they load, so two loads (reference models / drop-in
    models) coexist in one process."""
    if not available():
        raise FileNotFoundError("oracle/_ref is not staged: run `python oracle/stage_ref.py` in the build container")
    for stub in ("optuna", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    saved = {k: sys.modules.get(k) for k in ("models", "dataset")}
    real_makedirs = os.makedirs
    ns = types.SimpleNamespace()
    try:
        os.makedirs = lambda *a, **k: None
        with contextlib.redirect_stdout(io.StringIO()):
            ns.models = models if models is not None else _import_pyc("models", f"_gapref_{tag}_models")
            sys.modules["models"] = ns.models
            ns.dataset = _import_pyc("dataset", f"_gapref_{tag}_dataset")
            sys.modules["dataset"] = ns.dataset
            ns.train_gan = _import_pyc("train_gan", f"_gapref_{tag}_train_gan")
            ns.train = _import_pyc("train", f"_gapref_{tag}_train")
    finally:
        os.makedirs = real_makedirs
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ns.train_gan.tqdm = lambda it, **k: _NoBar(it)
    ns.train.tqdm = lambda it, **k: _NoBar(it)
    return ns
