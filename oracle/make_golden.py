"""ORACLE tooling: generate the golden fixtures by running the REFERENCE ITSELF (CPU, fp32).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Imports ``/root/reference`` unmodified through the harness-side shims documented in SURVEY.md
section 8(c) / Appendix A, pre-seeds its module-level encoder + tokenizer with seeded random-init
models (no network: no pretrained weights, no real vocab), and records what the reference's own
``encode_sentence_bert`` (``src/data/pipelines/text/_text.py:143-208``) and aggregations
(``src/data/metrics/_group.py:176-334, 392-458, 488-544``) return. Outputs:

* ``lmms-owc_b200/assets/class_names.json`` - the ten real class-name lists (``metadata.csv``
  ``class_name`` column with each task's ``doc_to_target`` transform);
* ``tests/golden/minilm_caltech101.npz`` - MiniLM-L6 arch: strings, padded ids/mask of the first
  batch, embeddings from ``encode_sentence_bert``, paired scores from ``semantic_similarity``,
  MASS dict, concept-similarity rows (with the ``simple_concepts`` extractor standing in for spaCy);
* ``tests/golden/mpnet_synth.npz`` - MPNet-base arch pushed through the same operator;
* ``tests/golden/mpnet_c3.npz`` - BASELINE config 3 in miniature (``--only-c3`` regenerates just this one): MPNet-base
  arch, 1000 class names x 2048 predictions: paired scores from ``semantic_similarity`` and the top-5 / margins of
  ``S = Zp Zc^T`` on the reference's own embeddings.

Weights are NOT stored: both sides rebuild them with ``torch.manual_seed(1234)`` + HF default
init (``build_model`` below); a few weight checksums are stored so a silently different RNG stream
shows up as "weights drifted" rather than as a parity failure.
"""

from __future__ import annotations

import csv
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import owc_b200  # noqa: E402,F401
from owc_b200 import arch as A  # noqa: E402
from owc_b200 import synth  # noqa: E402
from oracle import metrics_ref  # noqa: E402
from oracle.hf_models import build_hf_tokenizer, build_model, weight_checksums  # noqa: E402

SEED = 1234  # eval_metrics.py:193


def install_reference_shims():
    """SURVEY.md Appendix A, verbatim in effect; no edit to the reference tree."""
    import datasets
    import pandas  # noqa: F401  (must precede the pytz stub)
    import torchvision.io

    if not hasattr(torchvision.io, "VideoReader"):
        torchvision.io.VideoReader = type("VideoReader", (), {})
    orig = datasets.Dataset.__getitem__

    def getitem(self, key):
        return orig(self, slice(None))[key] if isinstance(key, str) else orig(self, key)

    datasets.Dataset.__getitem__ = getitem
    datasets.disable_caching()

    class _Stub(types.ModuleType):
        __path__: list = []

        def __getattr__(self, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return type(n, (object,), {})

    for n in ("pytz", "gdown", "pytablewriter"):
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)
    sys.path.insert(0, REF)


def extract_class_names() -> dict[str, list[str]]:
    transforms = {
        "caltech101": lambda s: s.replace("_", " "),
        "dtd": lambda s: s,
        "eurosat": lambda s: s.lower(),
        "fgvc_aircraft": lambda s: s,
        "flowers102": lambda s: s,
        "food101": lambda s: s.replace("_", " "),
        "oxford_pets": lambda s: s.replace("_", " "),
        "stanford_cars": lambda s: s,
        "sun397": lambda s: s.replace("_", " "),
        "ucf101": lambda s: s.replace("_", " ").lower(),
    }
    out = {}
    for name, fn in transforms.items():
        path = f"{REF}/src/data/tasks/_classification/{name}/assets/metadata.csv"
        if not os.path.exists(path):
            continue
        with open(path, newline="", encoding="utf-8") as f:
            out[name] = [fn(row["class_name"]) for row in csv.DictReader(f)]
    return out


def run_reference(arch, vocab, class_names, preds, targets, concept_preds):
    import src.data.pipelines.text as TP
    import src.data.pipelines.text._text as T
    from src.data.metrics import get_metric_info

    model = build_model(arch)
    tok = build_hf_tokenizer(arch, vocab)
    T.sentence_bert_model = model
    T.sentence_bert_processor = tok

    refs = [class_names[t] for t in targets]
    # offline item shape: (target: str, filtered_resps: list[str])  - eval_metrics.py:63-70
    items = [(r, [p]) for r, p in zip(refs, preds)]

    out = {}
    enc = T.encode_sentence_bert({"text": list(class_names)}, input_column="text")
    out["class_embeds"] = np.asarray(enc["text_sentence_bert_embeds"], dtype=np.float32)
    pe = []
    for s in range(0, len(preds), 1024):
        e = T.encode_sentence_bert({"prediction": preds[s:s + 1024]}, input_column="prediction")
        pe.append(np.asarray(e["prediction_sentence_bert_embeds"], dtype=np.float32))
    out["pred_embeds"] = np.concatenate(pe)

    mi = get_metric_info("semantic_similarity")
    out["paired_scores"] = np.asarray(mi.group_fn(mi.builder_fn(items), reduce="none"), dtype=np.float32)
    out["paired_mean"] = np.float64(mi.group_fn(mi.builder_fn(items), reduce="mean"))
    mm = get_metric_info("mean_average_semantic_similarity")
    mass_none = mm.group_fn(mm.builder_fn(items), reduce="none")
    mass_mean = mm.group_fn(mm.builder_fn(items), reduce="mean")
    out["mass_keys"] = np.array(list(mass_none.keys()))
    out["mass_none"] = np.asarray([mass_none[k] for k in mass_none], dtype=np.float32)
    out["mass_mean"] = np.asarray([mass_mean[k] for k in mass_mean], dtype=np.float64)

    if concept_preds is not None:
        # spaCy is absent: stand in the extractor on the reference side, keep everything else.
        TP.concept_extraction_spacy = metrics_ref.simple_concepts
        mc = get_metric_info("concept_semantic_similarity")
        citems = [(r, [p]) for r, p in zip(refs[:len(concept_preds)], concept_preds)]
        rows = mc.group_fn(mc.builder_fn(citems), reduce="none")
        out["concept_json"] = np.array(json.dumps([[list(c), [float(x) for x in s]] for c, s in rows]))
        out["concept_preds"] = np.array(concept_preds)
        for red in ("max", "mean", "median", "min"):
            out[f"concept_{red}"] = np.float64(mc.group_fn(mc.builder_fn(citems), reduce=red))

    first = tok(preds[:64], padding=True, truncation=True, return_tensors="pt")
    out["first_ids"] = first["input_ids"].numpy().astype(np.int32)
    out["first_mask"] = first["attention_mask"].numpy().astype(np.int32)
    with torch.no_grad():
        hs = model(**first, output_hidden_states=True).hidden_states
    # layer-by-layer anchors for kernel bring-up: embeddings output, layer 1, last layer (row 0..3)
    out["first_hidden_emb"] = hs[0][:4].numpy()
    out["first_hidden_l1"] = hs[1][:4].numpy()
    out["first_hidden_last"] = hs[-1][:4].numpy()
    out["weight_checksums"] = weight_checksums(model)
    out["class_names"] = np.array(class_names)
    out["preds"] = np.array(preds)
    out["targets"] = np.asarray(targets, dtype=np.int32)
    return out


def run_reference_c3(n_classes: int = 1000, n_preds: int = 2048):
    """BASELINE config 3 in miniature: the mpnet-base arch against 1000 class names. Records what the reference's own
    operator + ``semantic_similarity`` return, and the top-5 / margins of ``S = Zp Zc^T`` on the reference's embeddings
    (SURVEY.md 8c definition of the north-star extension). Embeddings are kept for the class names' first rows only:
    the scores are what the parity test needs."""
    import src.data.pipelines.text._text as T
    from src.data.metrics import get_metric_info

    arch = A.MPNET_BASE
    vocab = synth.build_vocab("mpnet", arch.vocab)
    words = synth.vocab_words(vocab)
    cls = synth.make_class_names(n_classes, words, seed=SEED)
    preds, targets = synth.make_predictions(n_preds, cls, words, seed=SEED)
    model = build_model(arch)
    T.sentence_bert_model = model
    T.sentence_bert_processor = build_hf_tokenizer(arch, vocab)

    def encode(texts, col):
        parts = []
        for s in range(0, len(texts), 1024):
            e = T.encode_sentence_bert({col: texts[s:s + 1024]}, input_column=col)
            parts.append(np.asarray(e[f"{col}_sentence_bert_embeds"], dtype=np.float32))
        return np.concatenate(parts)

    zc = encode(list(cls), "text")
    zp = encode(list(preds), "prediction")
    items = [(cls[t], [p]) for t, p in zip(targets, preds)]
    mi = get_metric_info("semantic_similarity")
    paired = np.asarray(mi.group_fn(mi.builder_fn(items), reduce="none"), dtype=np.float32)
    s64 = torch.from_numpy(zp).double() @ torch.from_numpy(zc).double().T
    s32 = torch.from_numpy(zp) @ torch.from_numpy(zc).T
    order = torch.sort(s32, dim=1, descending=True, stable=True)
    top64 = torch.sort(s64, dim=1, descending=True, stable=True).values
    consistency = float((s32[torch.arange(len(preds)), torch.tensor(targets)] - torch.from_numpy(paired)).abs().max())
    assert consistency < 1e-5, consistency     # S[i, target_i] == semantic_similarity(items, "none")[i]  
    return {"n_classes": np.int32(n_classes), "n_preds": np.int32(n_preds), "seed": np.int32(SEED),
            "targets": np.asarray(targets, dtype=np.int32), "paired_scores": paired,
            "top5_scores": order.values[:, :5].numpy().astype(np.float32), "top5_idx": order.indices[:, :5].numpy().astype(np.int32),
            "margin": (top64[:, 0] - top64[:, 1]).numpy().astype(np.float64),
            "class_embeds_head": zc[:32], "pred_embeds_head": zp[:64],
            "class_name_head": np.array(cls[:8]), "pred_head": np.array(preds[:8]),
            "weight_checksums": weight_checksums(model)}


def main() -> None:
    if "--only-c3" in sys.argv:
        install_reference_shims()
        out = run_reference_c3()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mpnet_c3.npz"), **out)
        print("c3:", out["top5_scores"].shape, float(out["paired_scores"].mean()), "near-ties:", int((out["margin"] < 1e-3).sum()))
        return
    names = extract_class_names()
    os.makedirs(os.path.join(ROOT, "lmms-owc_b200", "assets"), exist_ok=True)
    with open(os.path.join(ROOT, "lmms-owc_b200", "assets", "class_names.json"), "w", encoding="utf-8") as f:
        json.dump(names, f, indent=0, ensure_ascii=False)
    print({k: len(v) for k, v in names.items()})

    install_reference_shims()
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)

    # config 1 in miniature: Caltech101's real class names, MiniLM-L6 arch.
    vocab = synth.build_vocab("bert")
    words = synth.vocab_words(vocab)
    cls = names["caltech101"]
    preds, targets = synth.make_predictions(288, cls, words, seed=SEED)
    longp, _ = synth.make_predictions(8, cls, words, seed=SEED + 1, long=True)
    preds = preds + longp + ["", "a"]                # edge rows: long, empty string, one char
    targets = targets + [0] * (len(longp) + 2)
    concept_preds = [f"a {cls[i]} and the {words[i]}, {words[i + 1]} {words[i + 2]}" for i in range(24)]
    concept_preds += ["this photo", "an image, it and two"]   # rows whose concepts are all skip words
    out = run_reference(A.MINILM_L6, vocab, cls, preds, targets, concept_preds)
    np.savez_compressed(os.path.join(gold, "minilm_caltech101.npz"), **out)
    print("minilm:", out["pred_embeds"].shape, float(out["paired_mean"]))

    # MPNet-base arch through the same operator.
    mvocab = synth.build_vocab("mpnet", A.MPNET_BASE.vocab)
    mwords = synth.vocab_words(mvocab)
    mcls = synth.make_class_names(24, mwords, seed=SEED)
    mpreds, mtargets = synth.make_predictions(72, mcls, mwords, seed=SEED)
    out = run_reference(A.MPNET_BASE, mvocab, mcls, mpreds, mtargets, None)
    np.savez_compressed(os.path.join(gold, "mpnet_synth.npz"), **out)
    print("mpnet:", out["pred_embeds"].shape, float(out["paired_mean"]))

    out = run_reference_c3()
    np.savez_compressed(os.path.join(gold, "mpnet_c3.npz"), **out)
    print("c3:", out["top5_scores"].shape, float(out["paired_scores"].mean()))


if __name__ == "__main__":
    main()
