"""Batched eval protocol: a whole What'sUp / COCO-spatial / GQA-spatial or MMVP-style set in a
handful of launches, accuracies accounted on the device.

  reference: one tiny forward per item in a Python loop + dict bookkeeping on the host
      PACL/eval_pacl.py:26-104 (eval), :106-186 (eval_4), :268-349 (MMVP pairs); same protocol in eval_sparc.py,
      eval_llm2pacl.py
  here: `functional.pacl_eval_scores` scores all items at once (fp32, bit-exact top-1), `clipk_eval_*` do the
      bookkeeping; only the final handful of counters crosses to the host.
"""
import torch

from . import _lib
from . import functional as Fk

RELATIONS = {"left": 0, "right": 1, "on": 2, "under": 3, "in-front": 4, "behind": 5}


def _i32(t, device):
    return torch.as_tensor(t, device=device).to(torch.int32).contiguous()


def whatsup_accuracies(visual_proj, text_proj, set_id, rel_id, c=100.0, activation="sigmoid"):
    """visual_proj [items,P,D], text_proj [items,K,D] (caption 0 is the ground truth), set_id [items] (index of the
    object pair), rel_id [items] (RELATIONS) -> dict with the numbers eval_pacl.py:86-104 writes, plus the raw counts
    and the per-item `correct` flags (device tensors)."""
    scores, _ = Fk.pacl_eval_scores(visual_proj, text_proj, c, activation)
    return whatsup_from_scores(scores, set_id, rel_id)


def whatsup_from_scores(scores, set_id, rel_id):
    Fk._need_cuda(scores)
    dev = scores.device
    scores = scores.float().contiguous()
    items, K = scores.shape
    set_id, rel_id = _i32(set_id, dev), _i32(rel_id, dev)
    nsets = int(set_id.max().item()) + 1
    correct = torch.empty(items, dtype=torch.int32, device=dev)
    winner = torch.empty(nsets * 6, dtype=torch.int32, device=dev)
    counts = torch.empty(8, dtype=torch.int32, device=dev)
    st = Fk._stream()
    _lib.call("clipk_eval_correct", scores.data_ptr(), items, K, correct.data_ptr(), st)
    _lib.call("clipk_eval_whatsup", correct.data_ptr(), set_id.data_ptr(), rel_id.data_ptr(), items, nsets,
              winner.data_ptr(), counts.data_ptr(), st)
    ind_lr, ind_ou, ind_fb, pair_lr, pair_ou, pair_fb, sets, total = counts.tolist()      # the one D2H read
    return {
        "Individual accuracy": (ind_lr + ind_ou + ind_fb) * 100 / total,
        "Left Right Individual accuracy": ind_lr * 100 / (total / 2),
        "On Under Individual accuracy": ind_ou * 100 / (total / 2),
        "Front Back Individual accuracy": ind_fb * 100 / (total / 2),
        "Left Right Pair accuracy": pair_lr * 100 / (total / 4),
        "On Under Pair accuracy": pair_ou * 100 / (total / 4),
        "Front Back Pair accuracy": pair_fb * 100 / (total / 4),
        "Pair accuracy": (pair_lr + pair_ou + pair_fb) * 100 / (total / 2),
        "Set accuracy": sets * 100 / (total / 4),
        "counts": [ind_lr, ind_ou, ind_fb, pair_lr, pair_ou, pair_fb, sets, total],
        "correct": correct,
    }


def mmvp_accuracies(visual_proj_1, visual_proj_2, text_proj, gt, pairs_per_category=0, categories=1, c=100.0,
                    activation="sigmoid"):
    """MMVP-style pairs (eval_pacl.py:268-349): visual_proj_1/2 [pairs,P,D] (the two images of a pair), text_proj
    [pairs,2,D] (the two statements), gt [pairs,2] (1 = img1 is the right answer for that statement)."""
    s1, _ = Fk.pacl_eval_scores(visual_proj_1, text_proj, c, activation)
    s2, _ = Fk.pacl_eval_scores(visual_proj_2, text_proj, c, activation)
    return mmvp_from_scores(s1, s2, gt, pairs_per_category, categories)


def mmvp_from_scores(s1, s2, gt, pairs_per_category=0, categories=1):
    Fk._need_cuda(s1, s2)
    dev = s1.device
    s1, s2 = s1.float().contiguous(), s2.float().contiguous()
    pairs = s1.shape[0]
    gt = _i32(gt, dev)
    pred = torch.empty(pairs, 2, dtype=torch.int32, device=dev)
    counts = torch.empty(categories, 2, dtype=torch.int32, device=dev)
    _lib.call("clipk_eval_mmvp", s1.data_ptr(), s2.data_ptr(), gt.data_ptr(), pairs, int(pairs_per_category),
              int(categories), pred.data_ptr(), counts.data_ptr(), Fk._stream())
    cnt = counts.tolist()
    return {
        "Pair": 100 * sum(cpair for cpair, _ in cnt) / pairs,
        "Individual": 100 * sum(cs for _, cs in cnt) / pairs / 2,
        "per_category_counts": cnt,
        "pred": pred,
    }
