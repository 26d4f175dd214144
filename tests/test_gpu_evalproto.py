"""GPU parity: batched eval protocol (PACL/eval_pacl.py) -- integer bookkeeping, bit-exact against the
oracle's restatement, on planted score tables (ties, duplicate keys, missing relations) and end to end on seeded
What'sUp- / MMVP-shaped feature sets."""
import pytest
import torch

from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _whatsup_keys(items, seed):
    g = torch.Generator().manual_seed(seed)
    nsets = max(1, items // 4)
    set_id = torch.randint(0, nsets, (items,), generator=g)
    rel_id = torch.randint(0, 6, (items,), generator=g)
    return set_id, rel_id


@pytest.mark.parametrize("items,K", [(412, 2), (408, 4), (37, 3)])
def test_whatsup_accounting_exact(items, K):
    from clip_embeds_b200 import evalproto
    g = torch.Generator().manual_seed(items + K)
    scores = torch.randn(items, K, generator=g)
    scores[::7, 1] = scores[::7, 0]                  # exact ties: the comparison is strict, so these are wrong
    set_id, rel_id = _whatsup_keys(items, 5)
    want, correct = O.whatsup_accounting(scores, set_id, rel_id)
    got = evalproto.whatsup_from_scores(scores.cuda(), set_id, rel_id)
    assert got["correct"].cpu().tolist() == correct
    assert got["counts"] == want


def test_whatsup_end_to_end_matches_oracle():
    from clip_embeds_b200 import evalproto
    items, K, P, D = 64, 4, 196, 256
    V, T = O.rn(71, items, P, D), O.rn(72, items, K, D)
    set_id = torch.arange(items) // 4
    rel_id = torch.arange(items) % 4                 # left, right, on, under of every set
    _, sc = O.eval_top1(V, T, 100.0)
    want, correct = O.whatsup_accounting(sc, set_id, rel_id)
    got = evalproto.whatsup_accuracies(V.cuda(), T.cuda(), set_id, rel_id)
    assert got["correct"].cpu().tolist() == correct
    assert got["counts"] == want
    assert abs(got["Individual accuracy"] - sum(want[:3]) * 100 / items) < 1e-9


def test_mmvp_accounting_exact():
    from clip_embeds_b200 import evalproto
    pairs, ncat = 135, 9
    g = torch.Generator().manual_seed(9)
    s1, s2 = 100 * torch.rand(pairs, 2, generator=g), 100 * torch.rand(pairs, 2, generator=g)
    s2[::5, 0] = s1[::5, 0]                          # ties: probability exactly 0.5 -> "img2"
    gt = torch.randint(0, 2, (pairs, 2), generator=g)
    want, pred = O.mmvp_accounting(s1, s2, gt, 15, ncat)
    got = evalproto.mmvp_from_scores(s1.cuda(), s2.cuda(), gt, 15, ncat)
    assert torch.equal(got["pred"].cpu(), pred)
    assert got["per_category_counts"] == want


def test_mmvp_end_to_end_matches_oracle():
    from clip_embeds_b200 import evalproto
    pairs, P, D = 30, 196, 256
    V1, V2, T = O.rn(81, pairs, P, D), O.rn(82, pairs, P, D), O.rn(83, pairs, 2, D)
    gt = torch.stack([torch.arange(pairs) % 2, (torch.arange(pairs) + 1) % 2], 1)
    _, s1 = O.eval_top1(V1, T, 100.0)
    _, s2 = O.eval_top1(V2, T, 100.0)
    want, pred = O.mmvp_accounting(s1, s2, gt, 15, 2)
    got = evalproto.mmvp_accuracies(V1.cuda(), V2.cuda(), T.cuda(), gt, 15, 2)
    assert torch.equal(got["pred"].cpu(), pred)
    assert got["per_category_counts"] == want


def test_protocols_match_reference_golden_g11():
    """The CUDA protocol kernels against G11: the numbers the reference's own eval / eval_4 / eval_MMVP wrote (executed
    unmodified on planted scores, oracle/make_golden_protocols.py), exact ties included."""
    import json
    import os
    from clip_embeds_b200 import evalproto
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "goldens_protocols.json")))
    for name in ("eval", "eval_4"):
        c = G[name]
        got = evalproto.whatsup_from_scores(torch.tensor(c["scores"]).cuda(), torch.tensor(c["set_id"]), torch.tensor(c["rel_id"]))
        for k, v in c["reference"].items():
            assert abs(got[k] - v) < 1e-9, (name, k, got[k], v)
    for name in ("mmvp", "mmvpvlm"):
        c = G[name]
        got = evalproto.mmvp_from_scores(torch.tensor(c["s1"]).cuda(), torch.tensor(c["s2"]).cuda(), torch.tensor(c["gt"]),
                                         c["pairs_per_cat"], c["ncat"])
        assert got["pred"].cpu().tolist() == c["pred"]
        assert abs(got["Pair"] - c["reference"]["Pair"]) < 1e-9
        assert abs(got["Individual"] - c["reference"]["Individual"]) < 1e-9
