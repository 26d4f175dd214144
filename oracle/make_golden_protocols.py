"""Golden G11: the reference's eval-protocol bookkeeping, executed UNMODIFIED.

    python oracle/make_golden_protocols.py        (build container only: reads /root/reference)

`eval`, `eval_4` and `eval_MMVP` of Patch-Aligned-Contrastive-Learning/eval_pacl.py (:26-104, :106-186, :236-349) are
extracted from the reference source with `ast` (the module itself cannot be imported: it loads models, datasets and a
CUDA device at import time) and executed with stubs for everything around the bookkeeping: a model that returns planted
features (so that `100.0 * image_features @ text_features.T` has a planted diagonal, exact ties included), `process`,
`Image.open`, `tqdm`, and small synthetic annotation lists / csv files in a temporary directory.  What they append to
`evaluation_results.txt` is parsed and stored with the planted scores in tests/golden/goldens_protocols.json; the oracle's
`whatsup_accounting` / `mmvp_accounting` are asserted against it here and on every CPU test run, the CUDA protocol
kernels on the GPU.  Test infrastructure: nothing under clip_embeds_b200/ imports this.
"""
import ast
import contextlib
import csv
import json
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_oracle as O  # noqa: E402

REF = "/root/reference/Patch-Aligned-Contrastive-Learning/eval_pacl.py"
OUT = os.path.join(ROOT, "tests", "golden", "goldens_protocols.json")
RELS = ["left", "right", "on", "under", "in-front", "behind"]
PREP = {"left": "to the left of", "right": "to the right of", "on": "on", "under": "under", "in-front": "in front of",
        "behind": "behind"}
OPP = {"left": "right", "right": "left", "on": "under", "under": "on", "in-front": "behind", "behind": "in-front"}


def _extract(names):
    src = open(REF).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
    assert set(out) == set(names), set(names) - set(out)
    return out


class _Img:
    def __init__(self, key):
        self.key = key

    def convert(self, _mode):
        return self


class _ImageMod:
    def __init__(self, by_path):
        self.by_path = by_path

    def open(self, path):
        return _Img(self.by_path[path])


class _Process:
    def preprocess_image(self, image):
        return torch.tensor([float(image.key)])

    def preprocess_text(self, options):
        return torch.zeros(len(options), 1)


def _model_from(scores):
    """scores: dict key -> list of K planted diagonal values.  Returns model(pixels, tokens) -> (img_feat [K,K], txt_feat
    [K,K]) with diag(100 * img @ txt.T)[k] = 100 * (scores[key][k] / 100); `effective` collects that value."""
    effective = {}

    def model(pixels, tokens):
        key = int(pixels.flatten()[0].item())
        s = torch.tensor(scores[key], dtype=torch.float32)
        K = tokens.shape[0]
        assert K == s.numel(), (K, s.numel())
        img = torch.diag(s / 100.0)
        txt = torch.eye(K)
        effective[key] = (100.0 * img @ txt.T).diagonal().tolist()
        return img, txt
    return model, effective


def _parse(path):
    vals = {}
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith("Pair:") and "Individual:" in ln:
            a, b = ln.split(",")
            vals["Pair"] = float(a.split(":")[1])
            vals["Individual"] = float(b.split(":")[1])
        elif ":" in ln:
            k, v = ln.rsplit(":", 1)
            vals[k.strip()] = float(v)
    return vals


def whatsup_case(fn_src, fname, four, seed):
    """Synthetic What'sUp-style annotation list: object pairs x 4 relations (subset A: left/right/on/under or subset B:
    left/right/in-front/behind), caption 0 = ground truth."""
    g = torch.Generator().manual_seed(seed)
    objs = [("mug", "knife"), ("book", "plate"), ("cup", "phone-case"), ("bowl", "fork"), ("can", "hat"), ("pen", "remote")]
    dataset, paths, scores, meta = [], {}, {}, []
    with tempfile.TemporaryDirectory() as td:
        key = 0
        for si, (o1, o2) in enumerate(objs):
            rels = ["left", "right", "on", "under"] if si % 2 == 0 else ["left", "right", "in-front", "behind"]
            for rel in rels:
                others = [r for r in rels if r != rel]
                caps = [f"A {o1.replace('-', ' ')} {PREP[rel]} a {o2.replace('-', ' ')}"]
                caps.append(f"A {o1.replace('-', ' ')} {PREP[OPP[rel]]} a {o2.replace('-', ' ')}")
                caps += [f"A {o1.replace('-', ' ')} {PREP[r]} a {o2.replace('-', ' ')}" for r in others if r != OPP[rel]]
                ipath = f"data/controlled_images/{o1}_{rel}_{o2}.jpeg"
                dataset.append({"image_path": ipath, "caption_options": caps})
                paths[os.path.join(td, ipath[5:])] = key
                K = 4 if four else 2
                s = (torch.randn(K, generator=g) * 3).round() / 4 + 20.0         # coarse grid: exact ties do occur
                if key % 7 == 3:
                    s[1] = s[0]                                                   # planted exact tie -> not correct (strict >)
                if key % 5 == 0:
                    s[0] = s.max() + 0.25                                         # clearly correct
                scores[key] = s.tolist()
                meta.append((si, RELS.index(rel)))
                key += 1
        model, eff = _model_from(scores)
        ns = {"torch": torch, "os": os, "tqdm": (lambda x: x), "Image": _ImageMod(paths), "process": _Process(),
              "device": torch.device("cpu"), "csv": csv, "json": json}
        # torch.cuda.amp.autocast() is a no-op context on the CPU build (it only warns); keep the reference's own call
        exec(fn_src, ns)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ns[fname](model, dataset, td, "controlled")
            ref = _parse(os.path.join(td, "evaluation_results.txt"))
        finally:
            os.chdir(cwd)
    eff_scores = [eff[k] for k in range(len(dataset))]
    return {"scores": eff_scores, "set_id": [m[0] for m in meta], "rel_id": [m[1] for m in meta], "reference": ref}


def mmvp_case(fn_src, mode, seed):
    g = torch.Generator().manual_seed(seed)
    npairs = 135 if mode == "mmvpvlm" else 20
    scores, paths = {}, {}
    with tempfile.TemporaryDirectory() as td:
        img_dir = os.path.join(td, "MLLM_VLM_Images" if mode == "mmvpvlm" else "MMVP_Images")
        rows = []
        for i in range(npairs):
            q1, q2 = 2 * i + 1, 2 * i + 2
            qtype = f"type{i // 15}"
            rows += [(q1, qtype, f"statement {q1}"), (q2, qtype, f"statement {q2}")]
            for q in (q1, q2):
                p = os.path.join(img_dir, qtype, f"{q}.jpg") if mode == "mmvpvlm" else os.path.join(img_dir, f"{q}.jpg")
                paths[p] = q
                s = (torch.randn(2, generator=g) * 2).round() / 2 + 18.0
                scores[q] = s.tolist()
            if i % 9 == 4:                       # planted tie: the two images score the same on statement 1 -> prob 0.5
                scores[q2][0] = scores[q1][0]
        main_csv = "Questions.csv" if mode == "mmvpvlm" else "Questions-clip.csv"
        with open(os.path.join(td, main_csv), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["qid", "type", "statement"])
            w.writerows(rows)
        other = "Questions-llava.csv" if mode == "mmvpvlm" else "Questions.csv"
        if not os.path.exists(os.path.join(td, other)):
            with open(os.path.join(td, other), "w", newline="") as f:
                csv.writer(f).writerow(["header"])
        model, eff = _model_from(scores)
        eff_log = {}

        def model2(pixels, tokens):           # the same image is scored once per pair: keep what the reference saw
            out = model(pixels, tokens)
            eff_log[int(pixels.flatten()[0].item())] = eff[int(pixels.flatten()[0].item())]
            return out
        ns = {"torch": torch, "os": os, "tqdm": (lambda x: x), "Image": _ImageMod(paths), "process": _Process(),
              "device": torch.device("cpu"), "csv": csv, "json": json}
        exec(fn_src, ns)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ns["eval_MMVP"](model2, td, mode)
            ref = _parse(os.path.join(td, "evaluation_results.txt"))
            preds = list(csv.reader(open(os.path.join(td, "output.csv"))))[1:]
        finally:
            os.chdir(cwd)
    s1 = [eff_log[2 * i + 1] for i in range(npairs)]
    s2 = [eff_log[2 * i + 2] for i in range(npairs)]
    gt = [[1, 0] for _ in range(npairs)]         # gt = img1 for odd qid (statement 1), img2 for even qid (statement 2)
    pred = [[int(r[2] == "img1"), int(r[3] == "img1")] for r in preds]
    return {"s1": s1, "s2": s2, "gt": gt, "pred": pred, "reference": ref,
            "pairs_per_cat": 15 if mode == "mmvpvlm" else 0, "ncat": 9 if mode == "mmvpvlm" else 1}


def check_oracle(G):
    """The oracle restatements against the reference outputs (also run by tests/test_oracle_protocols.py)."""
    for name in ("eval", "eval_4"):
        c = G[name]
        counts, _ = O.whatsup_accounting(torch.tensor(c["scores"]), torch.tensor(c["set_id"]), torch.tensor(c["rel_id"]))
        ind_lr, ind_ou, ind_fb, pair_lr, pair_ou, pair_fb, sets, total = counts
        ours = {
            "Individual accuracy": (ind_lr + ind_ou + ind_fb) * 100 / total,
            "Left Right Individual accuracy": ind_lr * 100 / (total / 2),
            "On Under Individual accuracy": ind_ou * 100 / (total / 2),
            "Front Back Individual accuracy": ind_fb * 100 / (total / 2),
            "Left Right Pair accuracy": pair_lr * 100 / (total / 4),
            "On Under Pair accuracy": pair_ou * 100 / (total / 4),
            "Front Back Pair accuracy": pair_fb * 100 / (total / 4),
            "Pair accuracy": (pair_lr + pair_ou + pair_fb) * 100 / (total / 2),
            "Set accuracy": sets * 100 / (total / 4),
        }
        for k, v in c["reference"].items():
            assert abs(ours[k] - v) < 1e-9, (name, k, ours[k], v)
    for name in ("mmvp", "mmvpvlm"):
        c = G[name]
        counts, pred = O.mmvp_accounting(torch.tensor(c["s1"]), torch.tensor(c["s2"]), torch.tensor(c["gt"]),
                                         c["pairs_per_cat"], c["ncat"])
        assert pred.tolist() == c["pred"], name
        pairs = len(c["s1"])
        assert abs(100 * sum(p for p, _ in counts) / pairs - c["reference"]["Pair"]) < 1e-9
        assert abs(100 * sum(s for _, s in counts) / pairs / 2 - c["reference"]["Individual"]) < 1e-9


def main():
    assert os.path.exists(REF), "reference not found; run in the build container"
    src = _extract(["eval", "eval_4", "eval_MMVP"])
    G = {"meta": {"torch": torch.__version__, "source": "Patch-Aligned-Contrastive-Learning/eval_pacl.py eval / eval_4 / eval_MMVP, "
                  "executed unmodified (ast-extracted) with a planted-feature model and stub datasets"},
         "eval": whatsup_case(src["eval"], "eval", False, 11), "eval_4": whatsup_case(src["eval_4"], "eval_4", True, 12),
         "mmvp": mmvp_case(src["eval_MMVP"], "mmvp", 13), "mmvpvlm": mmvp_case(src["eval_MMVP"], "mmvpvlm", 14)}
    check_oracle(G)
    with open(OUT, "w") as f:
        json.dump(G, f)
    for k in ("eval", "eval_4", "mmvp", "mmvpvlm"):
        print(k, {a: round(b, 3) for a, b in G[k]["reference"].items() if "ccuracy" in a or a in ("Pair", "Individual")})


if __name__ == "__main__":
    main()
