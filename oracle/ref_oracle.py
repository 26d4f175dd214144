"""CPU oracle: a restatement of the reference hot path in plain torch (CPU).

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / reported baseline.
The shipped package (`clip_embeds_b200/`) never imports it and has no CPU
fallback.

Parity status: the reference ships NO test, golden vector or fixture for this
path ("parity unpinned" by the reference's own tests).
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container by `oracle/make_golden.py` (which imports the unmodified
reference through `oracle/refload.py`) and committed as
`tests/golden/goldens.pt` + `tests/golden/goldens.json`;
`tests/test_oracle_golden.py` re-checks the restatement against them.

Every function cites the reference lines it restates (paths relative to the
reference root; PACL = Patch-Aligned-Contrastive-Learning).
All functions are differentiable torch code so autograd supplies the
reference gradients.
"""
import torch
import torch.nn.functional as F


def l2n(x, eps=1e-12):
    """F.normalize(x, dim=-1): x / max(||x||_2, eps).  (torch semantics used at
    PACL/model/pacl.py:122,125,145,475-476)"""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


# ----------------------------------------------------------------------------
# PACL  (PACL/model/pacl.py)
# ----------------------------------------------------------------------------
def patch_alignment(V, T):
    """pacl.py:120-133 (copies :249-263, :341-355).
    V [Bi,P,D], T [Bt,D]; Bi == Bt (paired) or Bi == 1 (broadcast).
    Returns sigmoid(10 * <t^, v^_p>) with shape [Bt,P] (the reference's
    `.squeeze()` is mirrored only for the documented shapes)."""
    vn = l2n(V).transpose(-2, -1)          # [Bi,D,P]
    tn = l2n(T).unsqueeze(1)               # [Bt,1,D]
    act = (tn @ vn).squeeze(1)             # [Bt,P]
    return torch.sigmoid(act * 10)


def pacl_forward(V, T, activation="sigmoid"):
    """pacl.py:135-145 and variants (:190-197, :206-209, :270-275, :362-365).
    activation='ones' is the checked-in "Eval only" forward (:141-142);
    'sigmoid' is the activation-weighted pooling (:206-207, :362-363).
    Pooling uses the RAW V (:143).  Returns (img_feat [Bt,D], txt_feat [Bt,D])."""
    a = patch_alignment(V, T)
    if activation == "ones":
        a = torch.ones_like(a)
    elif activation == "softmax":
        # north_star (2) / SURVEY Appendix A.1 third activation (NO reference implementation: this branch restates
        # the definition, it is not pinned by a reference output): softmax over the patches of 10 * cos
        vn = l2n(V).transpose(-2, -1)
        a = torch.softmax(10.0 * (l2n(T).unsqueeze(1) @ vn).squeeze(1), dim=-1)
    pooled = torch.sum(V * a.unsqueeze(-1), dim=1)
    return l2n(pooled), l2n(T)


def pacl_clip_loss(img, txt, temperature):
    """ClipLoss, pacl.py:489-514: scale features first, two GEMMs, two CEs."""
    s = 1.0 / temperature
    li = s * img @ txt.T
    lt = s * txt @ img.T
    labels = torch.arange(li.shape[0])
    return (F.cross_entropy(li, labels) + F.cross_entropy(lt, labels)) / 2


def pacl_allpairs_scores(V, T, c=100.0, activation="sigmoid"):
    """All-pairs text-conditioned score: loop the reference
    eval call `model(img_i, texts)` (eval_pacl.py:53-57, 303-309: one image x
    all texts, diagonal of `c * image_features @ text_features.T`) over images.
    score[i,k] = c * < n(sum_p a_ikp V_ip), n(t_k) >.  Returns [Bi,Bt]."""
    rows = []
    for i in range(V.shape[0]):
        img, txt = pacl_forward(V[i:i + 1], T, activation)     # [Bt,D] each
        rows.append(c * (img * txt).sum(-1))                   # diagonal of img@txt.T
    return torch.stack(rows, 0)


def pacl_allpairs_loss(V, T, temperature=0.1, activation="sigmoid"):
    """per-image reference loop, then the two
    F.cross_entropy calls of pacl.py:509-512 on the score matrix and its
    transpose, logit scale 1/temperature."""
    L = pacl_allpairs_scores(V, T, 1.0 / temperature, activation)
    labels = torch.arange(L.shape[0])
    return (F.cross_entropy(L, labels) + F.cross_entropy(L.T, labels)) / 2


def eval_top1(V, T_items, c=100.0, activation="sigmoid"):
    """Eval protocol of eval_pacl.py:50-57 / eval_llm2pacl.py:62-67: per item one
    image [P,D] and K captions [K,D]; top-1 = argmax_k of the diagonal scores.
    V [items,P,D], T_items [items,K,D] -> (top1 [items] int64, scores [items,K])."""
    sc = torch.stack([pacl_allpairs_scores(V[i:i + 1], T_items[i], c, activation)[0]
                      for i in range(V.shape[0])], 0)
    return sc.argmax(-1), sc


# ----------------------------------------------------------------------------
# SPARC  (PACL/model/pacl.py)
# ----------------------------------------------------------------------------
def sparc_forward(V, L, mask, sigma):
    """sparc.forward, pacl.py:453-478.  V [B,P,D] raw, L [B,T,D] raw, mask [B,T].
    Returns (V, n(L), n(G), mask)."""
    sim = torch.einsum("btd,bpd->btp", L, V)
    smin = sim.min(dim=-1, keepdim=True)[0]
    smax = sim.max(dim=-1, keepdim=True)[0]
    sim = (sim - smin) / (smax - smin + 1e-8)
    sim = torch.where(sim < sigma, 0.0, sim)
    w = sim / (sim.sum(dim=-1, keepdim=True) + 1e-8)
    G = torch.einsum("btp,bpd->btd", w, V)
    return V, l2n(L), l2n(G), mask


def sparc_scoring(V1, L, mask, sigma, local=False):
    """sparc.scoring, pacl.py:438-451: one image expanded to K captions."""
    K = L.shape[0]
    V = V1.expand(K, *V1.shape[1:]) if (V1.shape[0] == 1 and K > 1) else V1
    assert V.shape[0] == L.shape[0]
    v, lt, g, _ = sparc_forward(V, L, mask, sigma)
    gtxt = l2n(lt.mean(dim=1))
    if not local:
        return l2n(v.mean(dim=1)) @ gtxt.T
    return l2n(g.mean(dim=1)) @ gtxt.T


def _masked_pairwise(a, b, mask, scale):
    """SparcLoss.masked_pairwise_contrastive_loss, pacl.py:522-556."""
    B, T, _ = a.shape
    mask_logits = (1.0 - mask) * (-1e8)
    labels = torch.eye(T, dtype=a.dtype).unsqueeze(0).expand(B, -1, -1).reshape(B * T, -1)
    logits = torch.einsum("bmd,bnd->bmn", a, b) * scale + mask_logits.unsqueeze(1)
    loss = F.cross_entropy(logits.reshape(B * T, -1), labels, reduction="none")
    m = mask.reshape(-1)
    return (loss * m).sum() / m.sum()


def sparc_loss(V, l_hat, g_hat, mask, temperature, global_weight=0.5, local_weight=1.0):
    """SparcLoss.forward, pacl.py:558-584."""
    s = 1.0 / temperature
    gi = l2n(V.mean(dim=1))
    gt = l2n(l_hat.mean(dim=1))
    li = s * gi @ gt.T
    lt = s * gt @ gi.T
    labels = torch.arange(li.shape[0])
    gl = (F.cross_entropy(li, labels) + F.cross_entropy(lt, labels)) / 2
    ll = (_masked_pairwise(g_hat, l_hat, mask, s) + _masked_pairwise(l_hat, g_hat, mask, s)) / 2
    return global_weight * gl + local_weight * ll


# ----------------------------------------------------------------------------
# open_clip ClipLoss with hard negatives  (open_clip/src/open_clip/loss.py)
# ----------------------------------------------------------------------------
def openclip_loss_single(img, txt, logit_scale, logit_bias=None, usehardtext=False):
    """ClipLoss.forward with world_size == 1, loss.py:163-164, :127-135, :172-193.
    txt has B+H rows when usehardtext (the first B are the originals)."""
    li = logit_scale * img @ txt.T
    lt = logit_scale * txt @ img.T
    if logit_bias is not None:
        li = li + logit_bias
        lt = lt + logit_bias
    B = li.shape[0]
    lab_i = torch.arange(B)
    if usehardtext:
        lab_t = torch.cat([lab_i, -100 * torch.ones(lt.shape[0] - B, dtype=torch.long)])
    else:
        lab_t = lab_i
    return (F.cross_entropy(li, lab_i) + F.cross_entropy(lt, lab_t)) / 2


def openclip_loss_ranks(imgs, txts, logit_scale, local_loss=True, usehardtext=True, logit_bias=None):
    """Multi-rank ClipLoss restated in ONE process (loss.py:67-87 gather_features_diffsize,
    :137-170 get_logits, :127-135 labels).  imgs[r] [b,D], txts[r] [b+H_r,D] are
    the per-rank inputs (leaf tensors).  Returns the list of per-rank losses; summing
    them and calling backward() reproduces `gather_with_grad=True` gradients, since
    torch.distributed.nn.all_gather's backward sums every rank's gradient w.r.t. the
    gathered copy.  all_text order = [orig_0..orig_{W-1}, hard_0..hard_{W-1}] (:147-153)."""
    W = len(imgs)
    b = imgs[0].shape[0]
    all_img = torch.cat(imgs, 0)
    if usehardtext:
        all_txt = torch.cat([t[:b] for t in txts] + [t[b:] for t in txts], 0)
    else:
        all_txt = torch.cat(txts, 0)
    losses = []
    for r in range(W):
        if local_loss:
            li = logit_scale * imgs[r] @ all_txt.T
            lt = logit_scale * txts[r] @ all_img.T
            lab_i = torch.arange(b) + b * r
        else:
            li = logit_scale * all_img @ all_txt.T
            lt = li.T
            lab_i = torch.arange(li.shape[0])
        if logit_bias is not None:
            li = li + logit_bias
            lt = lt + logit_bias
        if usehardtext:
            lab_t = torch.cat([lab_i, -100 * torch.ones(lt.shape[0] - lab_i.shape[0], dtype=torch.long)])
        else:
            lab_t = lab_i
        losses.append((F.cross_entropy(li, lab_i) + F.cross_entropy(lt, lab_t)) / 2)
    return losses


# ----------------------------------------------------------------------------
# synthetic inputs (rn(seed, *shape))
# ----------------------------------------------------------------------------
def rn(seed, *shape, dtype=torch.float32):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(dtype)


# ----------------------------------------------------------------------------
# Projection heads  (PACL/model/pacl.py:35-48, :70-79)   -- SURVEY §8f rank 1
# ----------------------------------------------------------------------------
def layer_norm(x, weight, bias, eps=1e-5):
    """nn.LayerNorm(D) (pacl.py:71, :76): biased variance over the last dim, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * weight + bias


def gelu_erf(z):
    """nn.GELU() default (pacl.py:44): 0.5 z (1 + erf(z / sqrt 2))."""
    return 0.5 * z * (1.0 + torch.erf(z * 0.7071067811865476))


def patch_projection(x, sd, prefix=""):
    """Patch_Projection.forward (pacl.py:47-48): linear_projection(x) + non_linear_projection(x), with
    non_linear_projection = Linear -> GELU -> Linear (pacl.py:42-46).  `sd` holds the reference's state-dict keys."""
    W1, b1 = sd[prefix + "linear_projection.0.weight"], sd[prefix + "linear_projection.0.bias"]
    W2, b2 = sd[prefix + "non_linear_projection.0.weight"], sd[prefix + "non_linear_projection.0.bias"]
    W3, b3 = sd[prefix + "non_linear_projection.2.weight"], sd[prefix + "non_linear_projection.2.bias"]
    return x @ W1.T + b1 + gelu_erf(x @ W2.T + b2) @ W3.T + b3


def visual_projection(x, sd):
    """visual_projection = Sequential(LayerNorm, Dropout(0.1), Patch_Projection) in eval mode (pacl.py:70-74)."""
    return patch_projection(layer_norm(x, sd["0.weight"], sd["0.bias"]), sd, "2.")


def text_projection(x, sd):
    """text_projection = Sequential(LayerNorm, Dropout(0.1), Linear) in eval mode (pacl.py:75-79)."""
    return layer_norm(x, sd["0.weight"], sd["0.bias"]) @ sd["2.weight"].T + sd["2.bias"]


# ----------------------------------------------------------------------------
# Eval protocol accounting  (PACL/eval_pacl.py)   -- SURVEY §8f rank 2
# ----------------------------------------------------------------------------
_RELS = ("left", "right", "on", "under", "in-front", "behind")


def whatsup_accounting(scores, set_id, rel_id):
    """eval_pacl.py:53-104 on precomputed diagonal scores [items,K]: `correct` is the strict comparison of caption 0
    against every other caption (:56, :133); eval_dict[(object pair)][relation] = correct (:59-60, later items
    overwrite); individual / pair / set counts (:63-85).  Returns (counts [ind_lr, ind_ou, ind_fb, pair_lr, pair_ou,
    pair_fb, sets, total], correct list).  Parity: pinned by golden G11 -- the reference's own `eval` / `eval_4`
    executed unmodified on planted scores (oracle/make_golden_protocols.py, tests/golden/goldens_protocols.json)."""
    items, K = scores.shape
    correct = [int(all(bool(scores[i, 0] > scores[i, k]) for k in range(1, K))) for i in range(items)]
    eval_dict = {}
    for i in range(items):
        eval_dict.setdefault(int(set_id[i]), {r: 0 for r in _RELS})[_RELS[int(rel_id[i])]] = correct[i]
    ind = [0, 0, 0]
    pair = [0, 0, 0]
    sets = 0
    for d in eval_dict.values():
        for g, (a, b) in enumerate((("left", "right"), ("under", "on"), ("behind", "in-front"))):
            if d[a] and d[b]:
                pair[g] += 1
            ind[g] += d[a] + d[b]
        if sum(d.values()) == 4:
            sets += 1
    return ind + pair + [sets, items], correct


def mmvp_accounting(s1, s2, gt, pairs_per_cat=0, ncat=1):
    """eval_pacl.py:303-335 on precomputed diagonal scores s1, s2 [pairs,2] (image 1 / image 2 against the two
    statements): logits_per_text rows = statements, softmax over the two images (fp32), pred = img1 iff prob > 0.5;
    returns ([[pair, single] per category], pred [pairs,2])."""
    pairs = s1.shape[0]
    counts = [[0, 0] for _ in range(ncat)]
    pred = torch.zeros(pairs, 2, dtype=torch.int32)
    for i in range(pairs):
        logits_per_text = torch.tensor([[s1[i, 0], s1[i, 1]], [s2[i, 0], s2[i, 1]]]).float().T
        ok = 0
        for t in range(2):
            p = int(logits_per_text[t].softmax(dim=-1)[0] > 0.5)
            pred[i, t] = p
            ok += int(p == int(gt[i, t] != 0))
        cat = min(i // pairs_per_cat, ncat - 1) if pairs_per_cat > 0 else 0
        counts[cat][0] += int(ok == 2)
        counts[cat][1] += ok
    return counts, pred


def apply_rope(embeddings):
    """apply_rope, pacl.py:147-181 (restated line by line; pinned by golden G8 in goldens_heads.pt)."""
    _, seq_length, dim = embeddings.shape
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
    angles = torch.arange(seq_length, dtype=torch.float).unsqueeze(1) * inv_freq
    s, c = torch.sin(angles).unsqueeze(0), torch.cos(angles).unsqueeze(0)
    x1, x2 = embeddings[..., 0::2], embeddings[..., 1::2]
    return torch.cat([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1)


# ----------------------------------------------------------------------------
# Other InfoNCE users  -- SURVEY §8f rank 4
# ----------------------------------------------------------------------------
def clip_metrics(image_features, text_features, logit_scale):
    """get_clip_metrics, open_clip/src/open_clip_train/train.py:360-377: descending argsort of the logits, position
    of the ground truth, mean / median rank (+1), R@1/5/10.  Also returns the 0-based positions per direction.
    Pinned by golden G9 (the reference's own function, run by oracle/make_golden.py)."""
    import numpy as np
    logits_per_image = logit_scale * image_features @ text_features.t()
    out, preds_all = {}, {}
    gt = torch.arange(len(text_features)).view(-1, 1)
    for name, logit in (("image_to_text", logits_per_image), ("text_to_image", logits_per_image.t())):
        ranking = torch.argsort(logit, descending=True)
        preds = torch.where(ranking == gt)[1].numpy()
        preds_all[name] = preds
        out[f"{name}_mean_rank"] = preds.mean() + 1
        out[f"{name}_median_rank"] = np.floor(np.median(preds)) + 1
        for k in (1, 5, 10):
            out[f"{name}_R@{k}"] = np.mean(preds < k)
    return out, preds_all


def simple_contrastive_loss(x, y, temperature=0.02, target=None, reduction="mean"):
    """SimpleContrastiveLoss.__call__, VLM2Vec/src/loss.py:11-19 (pinned by golden G10)."""
    if target is None:
        tpq = y.size(0) // x.size(0)
        target = torch.arange(0, x.size(0) * tpq, tpq, dtype=torch.long)
    return F.cross_entropy(x @ y.t() / temperature, target, reduction=reduction)
