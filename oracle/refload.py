"""Loader for the REAL reference implementation (container-only helper).

TEST INFRASTRUCTURE — not product code.  Imports the unmodified reference
modules from /root/reference by file path so that `oracle/make_golden.py`
can (a) validate the restatement in `oracle/ref_oracle.py` and (b) generate
the committed fixtures under `tests/golden/`.  /root/reference does not
exist on the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py
imports this module.

How:
  * open_clip/src/open_clip/loss.py imports standalone (torch only).
  * PACL/model/pacl.py needs stub modules for `open_clip.src.open_clip`
    (ftfy / timm are absent, the package import fails) and mutates HF_HOME
    and sys.path on import; both are restored afterwards.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("CLIP_EMBEDS_REFERENCE", "/root/reference")
PACL_PY = os.path.join(REF_ROOT, "Patch-Aligned-Contrastive-Learning", "model", "pacl.py")
LOSS_PY = os.path.join(REF_ROOT, "open_clip", "src", "open_clip", "loss.py")


def available() -> bool:
    return os.path.isfile(PACL_PY) and os.path.isfile(LOSS_PY)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_open_clip_loss():
    if "loss" not in _cache:
        _cache["loss"] = _load(LOSS_PY, "_ref_open_clip_loss")
    return _cache["loss"]


def load_pacl():
    if "pacl" in _cache:
        return _cache["pacl"]
    saved_env = os.environ.get("HF_HOME")
    saved_path = list(sys.path)
    stubs = {}
    for name in ("open_clip", "open_clip.src", "open_clip.src.open_clip"):
        if name not in sys.modules:
            stubs[name] = types.ModuleType(name)
            sys.modules[name] = stubs[name]
    try:
        # attribute chain so `from open_clip.src import open_clip` style imports resolve
        sys.modules["open_clip"].src = sys.modules["open_clip.src"]
        sys.modules["open_clip.src"].open_clip = sys.modules["open_clip.src.open_clip"]
        mod = _load(PACL_PY, "_ref_pacl")
    finally:
        for name in stubs:
            sys.modules.pop(name, None)
        sys.path[:] = saved_path
        if saved_env is None:
            os.environ.pop("HF_HOME", None)
        else:
            os.environ["HF_HOME"] = saved_env
    _cache["pacl"] = mod
    return mod


def make_sparc(pacl_mod, V, L, mask, sigma):
    """Build a reference `sparc` module whose encoders return fixed tensors."""
    import torch.nn as nn

    s = pacl_mod.sparc.__new__(pacl_mod.sparc)
    nn.Module.__init__(s)
    s.sigma = sigma
    s.forward_visual = lambda _images: V
    s.forward_text = lambda _caps: (L, mask)
    return s


TRAIN_PY = os.path.join(REF_ROOT, "open_clip", "src", "open_clip_train", "train.py")
VLM2VEC_LOSS_PY = os.path.join(REF_ROOT, "VLM2Vec", "src", "loss.py")


def load_get_clip_metrics():
    """The reference's own `get_clip_metrics` (open_clip_train/train.py:360-377).  The module cannot be imported here
    (it pulls in the `open_clip` package), so the function's source is taken from the file with `ast` and executed
    unmodified in a namespace that provides what it uses (torch, numpy)."""
    import ast
    import numpy as np
    import torch
    src = open(TRAIN_PY).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_clip_metrics")
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), TRAIN_PY, "exec"), ns)
    return ns["get_clip_metrics"]


def load_vlm2vec_loss():
    """VLM2Vec/src/loss.py imports only torch: loaded by file path."""
    if "vlm2vec_loss" not in _cache:
        _cache["vlm2vec_loss"] = _load(VLM2VEC_LOSS_PY, "_ref_vlm2vec_loss")
    return _cache["vlm2vec_loss"]
