"""CPU: host-side logic that needs no GPU -- drop-in module structure / state-dict compatibility with the reference,
the all-pairs schedule override, RoPE angle tables."""
import os

import torch

from oracle import ref_oracle as O


def _g7():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "goldens_heads.pt"), weights_only=True)["G7"]


def test_heads_state_dict_is_the_references():
    """`VisualProjection` / `TextProjection` expose exactly the reference's state-dict keys (incl. the
    `linear_projection` / `text_projection` alias of ONE Linear, pacl.py:39) and load a reference state dict."""
    from clip_embeds_b200.heads import Patch_Projection, TextProjection, VisualProjection
    G = _g7()
    vis, txt = VisualProjection(128, 64), TextProjection(64)
    assert sorted(vis.state_dict().keys()) == sorted(G["vis_sd"].keys())
    assert sorted(txt.state_dict().keys()) == sorted(G["txt_sd"].keys())
    vis.load_state_dict(G["vis_sd"])
    txt.load_state_dict(G["txt_sd"])
    assert vis[2].linear_projection[0].weight is vis[2].text_projection[0].weight
    assert len(list(vis.parameters())) == 8                 # LN (2) + three Linear layers (6); the alias is not counted twice
    pp = Patch_Projection()                                 # reference defaults: in_dim=768, out_dim=512 (pacl.py:36)
    assert pp.linear_projection[0].weight.shape == (512, 768) and pp.non_linear_projection[2].weight.shape == (512, 512)


def test_heads_fail_loudly_on_cpu():
    import pytest
    from clip_embeds_b200 import _lib
    from clip_embeds_b200.heads import VisualProjection, apply_rope
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.ClipkError):
        VisualProjection(128, 64).eval()(torch.randn(2, 5, 128))
    with pytest.raises(_lib.ClipkError):
        apply_rope(torch.randn(1, 4, 8))


def test_rope_tables_are_the_references_expressions():
    from clip_embeds_b200.heads import _rope_tables
    S, D = 50, 64
    sn, cs = _rope_tables(S, D, torch.device("cpu"))
    x = O.rn(25, 2, S, D)
    x1, x2 = x[..., 0::2], x[..., 1::2]
    y = torch.cat([x1 * cs - x2 * sn, x1 * sn + x2 * cs], dim=-1)
    assert torch.equal(y, O.apply_rope(x))                  # same tables -> bit-identical rotation on the CPU


def test_allpairs_schedule_override(monkeypatch):
    import clip_embeds_b200.functional as Fk
    monkeypatch.delenv("CLIPK_AP_SCHEDULE", raising=False)
    base = Fk.default_schedule(1024, 576, 768, True)
    assert isinstance(base, tuple) and len(base) == 2
    monkeypatch.setenv("CLIPK_AP_SCHEDULE", "64:3")
    assert Fk.default_schedule(1024, 576, 768, True) == (64, 3)
    monkeypatch.setenv("CLIPK_AP_SCHEDULE", "-16")
    assert Fk.default_schedule(1024, 576, 768, False) == (-16, 1)
