"""CPU: hand-checked cases for the oracle restatements of the eval protocol accounting (PACL/eval_pacl.py), open_clip's
get_clip_metrics and VLM2Vec's SimpleContrastiveLoss -- the checkers the GPU tests of SURVEY §8f ranks 2 and 4 rely on."""
import torch

from oracle import ref_oracle as O


def test_whatsup_accounting_hand_case():
    # two object pairs (sets); relations: 0 left, 1 right, 2 on, 3 under
    scores = torch.tensor([[2.0, 1.0], [3.0, 1.0], [1.0, 1.0], [5.0, 4.0],      # set 0: left ok, right ok, on TIE (wrong), under ok
                           [0.0, 1.0], [2.0, 1.0], [2.0, 1.0], [2.0, 1.0],      # set 1: left wrong, right ok, on ok, under ok
                           [9.0, 1.0]])                                         # set 1, left again: overwrites -> ok
    set_id = torch.tensor([0, 0, 0, 0, 1, 1, 1, 1, 1])
    rel_id = torch.tensor([0, 1, 2, 3, 0, 1, 2, 3, 0])
    counts, correct = O.whatsup_accounting(scores, set_id, rel_id)
    assert correct == [1, 1, 0, 1, 0, 1, 1, 1, 1]
    ind_lr, ind_ou, ind_fb, pair_lr, pair_ou, pair_fb, sets, total = counts
    assert (ind_lr, ind_ou, ind_fb) == (4, 3, 0)          # set 1's left was overwritten by the later, correct item
    assert (pair_lr, pair_ou, pair_fb) == (2, 1, 0)
    assert sets == 1 and total == 9                        # only set 1 has all four relations right


def test_mmvp_accounting_hand_case():
    s1 = torch.tensor([[5.0, 1.0], [3.0, 3.0], [1.0, 9.0]])
    s2 = torch.tensor([[4.0, 2.0], [3.0, 4.0], [2.0, 8.0]])
    gt = torch.tensor([[1, 0], [1, 0], [0, 1]])
    counts, pred = O.mmvp_accounting(s1, s2, gt, 2, 2)
    # pair 0: text 1 prefers img1 (5 > 4), text 2 prefers img2 (1 < 2) -> both right
    # pair 1: text 1 ties (probability exactly 0.5 is not > 0.5 -> img2, wrong), text 2 prefers img2 -> right
    # pair 2: text 1 prefers img2 (right), text 2 prefers img1 (9 > 8, right)
    assert pred.tolist() == [[1, 0], [0, 0], [0, 1]]
    assert counts == [[1, 3], [1, 2]]                      # category 0 = pairs 0-1, category 1 = pair 2


def test_clip_metrics_hand_case():
    logits = torch.tensor([[9.0, 1.0, 2.0, 2.5], [1.0, 9.0, 2.0, 2.0], [4.0, 5.0, 6.0, 9.0], [1.0, 2.0, 9.0, 3.0]])
    img, txt = torch.eye(4), logits.t().contiguous()       # img_i . txt_j = logits[i, j]
    m, preds = O.clip_metrics(img, txt, 1.0)
    assert preds["image_to_text"].tolist() == [0, 0, 1, 1]  # rows 2 and 3 have one larger logit than the diagonal
    assert preds["text_to_image"].tolist() == [0, 0, 1, 1]  # columns 2 and 3 too
    for d in ("image_to_text", "text_to_image"):
        assert m[f"{d}_R@1"] == 0.5 and m[f"{d}_R@5"] == 1.0 and m[f"{d}_R@10"] == 1.0
        assert m[f"{d}_mean_rank"] == 1.5 and m[f"{d}_median_rank"] == 1.0      # floor(median([0,0,1,1])) + 1


def test_simple_contrastive_loss_label_map():
    x = O.l2n(O.rn(1, 4, 8))
    y = O.l2n(O.rn(2, 12, 8))                              # 3 candidates per query: targets 0, 3, 6, 9
    want = torch.nn.functional.cross_entropy(x @ y.t() / 0.02, torch.tensor([0, 3, 6, 9]))
    assert torch.allclose(O.simple_contrastive_loss(x, y, 0.02), want)


def test_protocol_accounting_matches_reference_golden_g11():
    """G11 (tests/golden/goldens_protocols.json): what the reference's own `eval`, `eval_4` and `eval_MMVP`
    (PACL/eval_pacl.py:26-104, :106-186, :236-349) wrote to evaluation_results.txt when executed UNMODIFIED on planted
    scores (oracle/make_golden_protocols.py).  The oracle's restatements must reproduce every number."""
    import json
    import os
    from oracle import make_golden_protocols as mg
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "goldens_protocols.json")))
    mg.check_oracle(G)
    # per-category MMVP numbers too (eval_pacl.py:341-349)
    c = G["mmvpvlm"]
    counts, _ = O.mmvp_accounting(torch.tensor(c["s1"]), torch.tensor(c["s2"]), torch.tensor(c["gt"]), 15, 9)
    cats = ['Orientation and Direction', 'Presence of Specific Features', 'State and Condition', 'Quantity and Count',
            'Positional and Relational Context', 'Color and Appearance', 'Structural Characteristics', 'Texts',
            'Viewpoint and Perspective']
    pairs = len(c["s1"])
    for i, name in enumerate(cats):
        assert abs(counts[i][0] / (pairs // 9) * 100 - c["reference"][f"{name} Pair accuracy"]) < 1e-9
        assert abs(counts[i][1] / (pairs * 2 // 9) * 100 - c["reference"][f"{name} Single accuracy"]) < 1e-9
    # the planted exact ties are in the data (strict comparison -> counted wrong)
    assert any(s[0] == s[1] for s in G["eval"]["scores"])
