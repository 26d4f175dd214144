"""Drop-in loss modules with the reference's constructor and call signatures.

  ClipLoss(temperature)(image_features, text_features)                      PACL/model/pacl.py:489-514
  OpenClipLoss(local_loss, gather_with_grad, cache_labels, rank, world_size, use_horovod, usehardtext)
      (image_features, text_features, logit_scale, logit_bias=None, output_dict=False)
                                                                            open_clip/src/open_clip/loss.py:89-193
  PaclAllPairsLoss(temperature)(visual_proj, text_proj)                     north_star all-pairs InfoNCE
      (reference: per-image eval loop + pacl.py:509-512)
"""
import torch
import torch.nn as nn

from . import dist as cdist
from . import functional as Fk


class ClipLoss(nn.Module):
    """Symmetric InfoNCE with fixed logit scale 1/temperature (pacl.py:489-514)."""

    def __init__(self, temperature):
        super().__init__()
        self.logit_scale = 1.0 / temperature

    def forward(self, image_features, text_features):
        return Fk.symmetric_infonce(image_features, text_features, self.logit_scale)


GRAD_REDUCTIONS = ("sum", "mean")


class PaclAllPairsLoss(nn.Module):
    """InfoNCE over the all-pairs text-conditioned score matrix; image-sharded across `group` ranks.

    forward(visual_proj [b,P,D] (this rank's images), text_proj [b,D] (this rank's captions)) -> global loss.
    Texts are all-gathered (with gradient: reduce-scatter in backward), V never moves.

    Gradient convention under sharding (`grad_reduction`): every rank returns the GLOBAL loss L and back-propagates
    through its own rows only.
      "sum"  (default): the inputs of rank r receive exactly dL/d(inputs of rank r).  Parameters shared by the ranks
             (the projection heads) then need their gradients SUM-reduced across ranks.
      "mean": the same gradients multiplied by the world size, so that a MEAN reduction -- what
             torch.nn.parallel.DistributedDataParallel does -- yields dL/d(parameters).  This is the convention of the
             reference's open_clip ClipLoss (every rank differentiates the full loss, the gather's backward sums) and of
             VLM2Vec's loss (`* world_size`, VLM2Vec/src/loss.py:35-41); use it whenever the heads are wrapped in DDP."""

    def __init__(self, temperature=0.1, activation="sigmoid", group=None, image_group=None, grad_reduction="sum"):
        super().__init__()
        if grad_reduction not in GRAD_REDUCTIONS:
            raise ValueError(f"grad_reduction must be one of {GRAD_REDUCTIONS}")
        self.logit_scale = 1.0 / temperature
        self.activation = activation
        self.group = group
        self.image_group = image_group
        self.grad_reduction = grad_reduction

    def forward(self, visual_proj, text_proj, v_sqnorm=None):
        """v_sqnorm (optional): squared row norms of `visual_proj` from `heads.VisualProjection(x, return_sqnorm=True)`."""
        pg = self.group
        if pg is not None and cdist.world_size(pg) > 1:
            all_text = cdist.all_gather_with_grad(text_proj, pg)
            offset = cdist.rank(pg) * visual_proj.shape[0]
        else:
            pg = None
            all_text = text_proj
            offset = 0
        scores = Fk.pacl_scores(visual_proj, all_text, self.logit_scale, self.activation, self.image_group, v_sqnorm)
        loss = Fk.score_infonce(scores, offset, pg)
        if pg is not None and self.grad_reduction == "mean":
            loss = cdist.scale_grad(loss, cdist.world_size(pg))
        return loss


class OpenClipLoss(nn.Module):
    """open_clip ClipLoss incl. the fork's left/right hard-negative texts (`usehardtext`), loss.py:89-193.

    text_features holds the rank's B original captions followed by its H_r hard negatives (collate layout,
    open_clip_train/data.py:122-134).  Horovod is not supported (as in the reference's diffsize gather)."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False, usehardtext=False, group=None):
        super().__init__()
        if use_horovod:
            raise NotImplementedError("horovod gather is not supported (reference: loss.py:76)")
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.usehardtext = usehardtext
        self.group = group
        self._labels = {}

    def _text_labels(self, text, n_orig, offset):
        """labels of the text->image CE (loss.py:127-135): row i < n_orig -> i + offset, hard-negative (and padding) rows
        -> ignore_index.  With cache_labels the tensor is built once per (rows, n_orig, offset, device), as the reference
        caches its ground truth (loss.py:110-125)."""
        key = (text.shape[0], n_orig, offset, text.device)
        if self.cache_labels and key in self._labels:
            return self._labels[key]
        lab = torch.full((text.shape[0],), -100, dtype=torch.int64, device=text.device)
        lab[:n_orig] = torch.arange(n_orig, device=text.device) + offset
        if self.cache_labels:
            self._labels[key] = lab
        return lab

    def forward(self, image_features, text_features, logit_scale, logit_bias=None, output_dict=False):
        b = image_features.shape[0]
        bias = 0.0 if logit_bias is None else logit_bias
        sym_ok = Fk.sym_feat_ce_ok(image_features, text_features)
        if self.world_size == 1 and sym_ok:
            # ONE logits GEMM: the text->image CE is the column CE of the same logits over the first b columns (the
            # originals; hard-negative rows are ignore_index rows of logits_per_text, loss.py:127-135)
            total = Fk.sym_feat_ce(image_features, text_features, logit_scale, bias, 0, b)
            return {"contrastive_loss": total} if output_dict else total
        if self.world_size > 1 and self.local_loss and self.gather_with_grad and sym_ok and (
                not self.usehardtext or b % 32 == 0):
            # local loss, one GEMM per rank: local image rows x all captions; the column sums are all-reduced, the image
            # features are never gathered, the caption gradient goes back through the gather's reduce-scatter
            pg = self.group if self.group is not None else torch.distributed.group.WORLD
            total = Fk.sym_feat_ce_dist(image_features, text_features, logit_scale, bias, self.rank, self.world_size, b,
                                        self.usehardtext, pg)
            return {"contrastive_loss": total} if output_dict else total
        if self.world_size > 1:
            if self.usehardtext:
                assert self.gather_with_grad, "usehardtext requires gather_with_grad (loss.py:77)"
            # device-side masks instead of the reference's size exchange (loss.py:78-86) whenever the slab geometry allows
            # it (b a multiple of 32: the CE kernels mask whole 32-column chunks' tails)
            fixed = self.usehardtext and b % 32 == 0 and image_features.is_cuda
            if fixed:
                all_img, all_txt, counts = cdist.gather_features(image_features, text_features, b, True,
                                                                 self.gather_with_grad, self.local_loss, self.rank,
                                                                 self.world_size, self.group, keep_padding=True)
                slab = (counts, self.world_size * b, b)
            else:
                all_img, all_txt = cdist.gather_features(image_features, text_features, b, self.usehardtext,
                                                         self.gather_with_grad, self.local_loss, self.rank,
                                                         self.world_size, self.group)
                slab = None
            if self.local_loss:
                off = b * self.rank
                li = Fk.feat_row_ce(image_features, all_txt, logit_scale, bias, None, off, slab)
                lt = Fk.feat_row_ce(text_features, all_img, logit_scale, bias, self._text_labels(text_features, b, off), 0)
            else:
                N = all_img.shape[0]
                li = Fk.feat_row_ce(all_img, all_txt, logit_scale, bias, None, 0, slab)
                lt = Fk.feat_row_ce(all_txt, all_img, logit_scale, bias, self._text_labels(all_txt, N, 0), 0)
        else:
            li = Fk.feat_row_ce(image_features, text_features, logit_scale, bias, None, 0)
            lt = Fk.feat_row_ce(text_features, image_features, logit_scale, bias, self._text_labels(text_features, b, 0), 0)
        total = (li + lt) / 2
        return {"contrastive_loss": total} if output_dict else total


class SparcLoss(ClipLoss):
    """SparcLoss (pacl.py:516-584): 0.5 * global InfoNCE on mean-pooled features + 1.0 * masked per-sample
    token <-> grouped-patch contrastive loss (both directions).

    With `group` set (one process per GPU, samples sharded), the global term gathers the two [b,D] pooled
    features (gradient: reduce-scatter) and the local term divides by the GLOBAL mask count, as the reference's
    DataParallel run does (outputs gathered before the loss, train_sparc.py:92-94); the returned value is the
    global loss, identical on every rank."""

    def __init__(self, temperature, group=None, grad_reduction="sum"):
        super().__init__(temperature)
        if grad_reduction not in GRAD_REDUCTIONS:
            raise ValueError(f"grad_reduction must be one of {GRAD_REDUCTIONS}")
        self.global_weight = 0.5
        self.local_weight = 1.0
        self.group = group
        self.grad_reduction = grad_reduction     # see PaclAllPairsLoss: "mean" when the heads are wrapped in DDP

    def forward(self, v_patch_embed, l_token_embed, l_grouped_v_patch_embed, language_mask):
        gi = Fk.normalize_rows(Fk.pooled_patch_mean(v_patch_embed))
        gt = Fk.normalize_rows(Fk.mean_dim1(l_token_embed))
        pg = self.group
        W = cdist.world_size(pg) if pg is not None else 1
        if W > 1:
            b = gi.shape[0]
            all_i = cdist.all_gather_with_grad(gi, pg)
            all_t = cdist.all_gather_with_grad(gt, pg)
            off = cdist.rank(pg) * b
            # rows of this rank against everything; every rank contributes its rows' sum, mean over the global N
            li = Fk.feat_row_ce(gi, all_t, self.logit_scale, 0.0, None, off) * (b / all_i.shape[0])
            lt = Fk.feat_row_ce(gt, all_i, self.logit_scale, 0.0, None, off) * (b / all_i.shape[0])
            global_part = (li + lt) / 2
            msum = language_mask.float().sum()
            cdist.all_reduce_sum_(msum, pg)
            local_part = Fk.sparc_local_loss(l_grouped_v_patch_embed, l_token_embed, language_mask, self.logit_scale, msum)
            total = self.global_weight * global_part + self.local_weight * local_part
            total = cdist.all_reduce_sum_with_grad(total, pg)
            return cdist.scale_grad(total, W) if self.grad_reduction == "mean" else total
        global_loss = Fk.symmetric_infonce(gi, gt, self.logit_scale)
        local_loss = Fk.sparc_local_loss(l_grouped_v_patch_embed, l_token_embed, language_mask, self.logit_scale)
        return self.global_weight * global_loss + self.local_weight * local_loss
