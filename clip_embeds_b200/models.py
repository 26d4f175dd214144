"""Module shells mirroring the reference models' scoring interface AFTER the encoders / projection heads.

The reference modules (PACL/model/pacl.py) wrap pretrained open_clip / LLM2CLIP towers that are out of scope
(SURVEY §2); these shells take the tensors those towers + heads produce and keep the method names and return
conventions of the reference, so a reference model can delegate its scoring lines to them:

    open_clip_pacl.patch_alignment / .forward      pacl.py:120-145  (llm2clip_pacl: :249-275)
    sparc.forward / .scoring                       pacl.py:438-478
"""
import torch
import torch.nn as nn

from . import functional as Fk


class PACLHead(nn.Module):
    """forward(visual_proj [B,P,D], text_proj [B,D]) -> (img_feat, txt_feat), as the reference `forward` does after
    `forward_visual` / `forward_text`.  `eval_only_ones=True` reproduces the checked-in forward that overwrites the
    activations with ones (pacl.py:141-142); False is the activation-weighted pooling (pacl.py:206-207, :362-363)."""

    def __init__(self, eval_only_ones=False):
        super().__init__()
        self.eval_only_ones = eval_only_ones

    def patch_alignment(self, visual_patch_proj, text_cls_proj):
        return Fk.patch_alignment(visual_patch_proj, text_cls_proj)

    def forward(self, visual_proj, text_proj):
        return Fk.pacl_pool(visual_proj, text_proj, "ones" if self.eval_only_ones else "sigmoid")

    def score_items(self, visual_proj, text_proj, c=100.0):
        """Eval protocol (eval_pacl.py:50-57): [items,P,D] x [items,K,D] -> (scores [items,K], top1 [items])."""
        return Fk.pacl_eval_scores(visual_proj, text_proj, c, "ones" if self.eval_only_ones else "sigmoid")


class SparcHead(nn.Module):
    """sparc.forward / sparc.scoring on precomputed patch and token embeddings (pacl.py:438-478)."""

    def __init__(self, sigma):
        super().__init__()
        self.sigma = sigma

    def forward(self, v_patch_embed, l_token_embed, language_mask):
        l_hat, g_hat = Fk.sparc_align(v_patch_embed, l_token_embed, self.sigma)
        return v_patch_embed, l_hat, g_hat, language_mask

    def scoring(self, v_patch_embed, l_token_embed, language_mask, local=False):
        if v_patch_embed.shape[0] == 1 and l_token_embed.shape[0] > 1:
            v_patch_embed = v_patch_embed.expand(l_token_embed.shape[0], *v_patch_embed.shape[1:])
        assert v_patch_embed.shape[0] == l_token_embed.shape[0]
        with torch.no_grad():
            v, l_hat, g_hat, _ = self.forward(v_patch_embed, l_token_embed, language_mask)
            gtxt = Fk.normalize_rows(Fk.mean_dim1(l_hat))
            if not local:
                gimg = Fk.normalize_rows(Fk.mean_dim1(v.contiguous()))
            else:
                gimg = Fk.normalize_rows(Fk.mean_dim1(g_hat))
            return gimg @ gtxt.T
