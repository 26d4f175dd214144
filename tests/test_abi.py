"""CPU: the C-ABI library builds, loads without a GPU, exports every symbol include/clipk.h declares, and the
product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "clipk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clipk_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from clip_embeds_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/clipk.h but not exported by libclipk.so"


def test_ctypes_table_covers_header():
    from clip_embeds_b200 import _lib
    assert set(_declared()) == set(_lib.SIGNATURES), set(_declared()) ^ set(_lib.SIGNATURES)


def test_fails_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clip_embeds_b200 import _lib
    L = _lib.lib()
    assert L.clipk_check_device() != 0
    assert b"no CPU fallback" in L.clipk_last_error() or b"sm_100" in L.clipk_last_error()
    import clip_embeds_b200.functional as Fk
    with pytest.raises(_lib.ClipkError):
        Fk.pacl_pool(torch.randn(2, 4, 8), torch.randn(2, 8))
    with pytest.raises(_lib.ClipkError):
        Fk.pacl_scores(torch.randn(2, 4, 8), torch.randn(2, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "clip_embeds_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src, f"{f} references the oracle"
