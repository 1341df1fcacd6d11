"""Group-level similarity metrics: drop-ins for the reference's ``src/data/metrics/_group.py``.

Same names, signatures, ``reduce`` values, return types and errors as
``semantic_similarity`` (``_group.py:488-544``), ``mean_average_semantic_similarity`` (``:392-458``)
and ``concept_semantic_similarity`` (``:176-334``). What changes is underneath: each DISTINCT string
is tokenised on the host and encoded once on the GPU (the reference re-encodes the class name for
every row and round-trips embeddings through Arrow as Python lists), and the paired dot runs on the
device through ``owc_pair_dot``. ``semantic_matching`` is the additive north-star extension
(prediction-by-class cosine matrix with fused top-k); it has no reference counterpart.
"""

from __future__ import annotations

from typing import Literal

import numpy as np

from .. import text as _text
from ._api import register_aggregation

__all__ = ["GROUP_METRICS", "concept_semantic_similarity", "mean_average_semantic_similarity",
           "semantic_similarity", "semantic_matching"]

GROUP_METRICS = ["concept_semantic_similarity", "mean_average_semantic_similarity", "semantic_similarity"]

_THRESHOLDS = [0.5, 0.6, 0.7, 0.8, 0.9]


def _select(items: list) -> tuple[list[str], list[str]]:
    # ref[0] if the reference is a list; pred[-1] (last round wins) - _group.py:515-519
    if hasattr(items, "refs") and hasattr(items, "preds"):      # jsonl.PackedPairs: the scanner applied both rules already
        return items.refs, items.preds
    columns = list(zip(*items, strict=True))      # IndexError on an empty list, like the reference
    refs, preds = columns[0], columns[1]
    refs = [ref[0] if isinstance(ref, list) else ref for ref in refs]
    preds = [pred[-1] if isinstance(pred, list) else pred for pred in preds]
    return refs, preds


def _paired_scores(refs: list[str], preds: list[str]) -> np.ndarray:
    """``s_i = <z(ref_i), z(pred_i)>`` (the ``torch.bmm`` of _group.py:537-544), fp32 [N]."""
    engine, tok = _text.get_engine()
    if hasattr(refs, "blob") and hasattr(preds, "blob"):        # jsonl.PackedStrings: dedup + tokenisation without Python rows
        tables = []
        for col in (refs, preds):
            uniq, inverse = col.unique()
            ids, cu = tok.encode_packed_blob(uniq.blob[:-1], uniq.n)
            tables.append((engine.encode(ids, cu, out="device"), inverse))
        (zr, ridx), (zp, pidx) = tables
        return engine.pair_dot(zr, zp, a_index=ridx, b_index=pidx)
    # embeddings stay in device memory between the two calls: only token ids / indices go up and N scores come down
    zr, ridx = _text.encode_unique(refs, out="device")
    zp, pidx = _text.encode_unique(preds, out="device")
    return engine.pair_dot(zr, zp, a_index=ridx, b_index=pidx)


@register_aggregation("semantic_similarity")
def semantic_similarity(items: list, reduce: Literal["none", "mean"] = "mean") -> float | list[float]:
    """Calculate the semantic similarity on a list of documents.

    Args:
    ----
        items (list): List of documents, each ``(reference, prediction)``.
        reduce ("none" | "mean"): "none" returns the value per sample; "mean" the mean similarity.

    """
    if reduce not in ["none", "mean"]:
        raise ValueError(
            "Unknown `reduce` value for `semantic_similarity` metric."
            ' Expected "none" or "mean", but got "%s"',
            reduce,
        )
    refs, preds = _select(items)
    scores = _paired_scores(refs, preds)
    if reduce == "mean":
        return float(scores.mean(dtype=np.float32))
    return scores.tolist()


@register_aggregation("mean_average_semantic_similarity")
def mean_average_semantic_similarity(
    items: list, reduce: Literal["none", "mean"] = "mean"
) -> dict[str, float] | dict[str, list[float]]:
    """Calculate the mean average semantic similarity (thresholds 0.5 ... 0.9 and their mean).

    Args:
    ----
        items (list): List of documents.
        reduce ("none" | "mean"): "none" returns per-sample values; "mean" the means.

    """
    if reduce not in ["none", "mean"]:
        raise ValueError(
            "Unknown `reduce` value for `mean_average_semantic_similarity` metric."
            ' Expected "none" or "mean", but got "%s"',
            reduce,
        )
    refs, preds = _select(items)
    scores = _paired_scores(refs, preds)
    outputs: dict = {}
    if reduce == "mean":
        for threshold in _THRESHOLDS:
            outputs[f"semantic_similarity@{threshold}"] = float(
                (scores >= np.float32(threshold)).astype(np.float32).mean(dtype=np.float32))
        outputs["semantic_similarity@avg"] = float(np.asarray(list(outputs.values()), dtype=np.float32).mean())
        return outputs
    for threshold in _THRESHOLDS:
        outputs[f"semantic_similarity@{threshold}"] = (scores >= np.float32(threshold)).astype(np.int32).tolist()
    outputs["semantic_similarity@avg"] = (
        np.asarray(list(outputs.values()), dtype=np.float32).mean(axis=0, dtype=np.float32).tolist())
    return outputs


# _group.py:207-234
_SKIP_WORDS_GROUPS = {
    "numbers_digits": ["1", "2", "3", "4", "5", "6", "7", "8", "9", "10"],
    "numbers_words": ["one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten"],
    "symbols": ["*"],
    "articles": ["a", "the"],
    "generic_nouns": ["image", "object", "photo", "type", "this photo"],
    "personal_pronouns": ["it", "they", "them"],
    "demonstratives": ["that", "this", "those"],
    "wh_words": ["which", "who", "whom", "whose", "where", "when", "what", "why", "how"],
    "quantifiers": ["some"],
}


@register_aggregation("concept_semantic_similarity")
def concept_semantic_similarity(
    items: list, reduce: Literal["none", "max", "mean", "median", "min"] = "max"
) -> float | list[tuple[list[str], list[float]]]:
    """Calculate the semantic similarity of the response concepts on a list of documents.

    Args:
    ----
        items (list): List of documents.
        reduce ("none" | "max" | "mean" | "median" | "min"): per-sample reduction over the concepts
            (then the mean over samples); "none" returns ``(concepts, similarities)`` per sample.

    """
    if reduce not in ["none", "max", "mean", "median", "min"]:
        raise ValueError(
            "Unknown `reduce` value for `concept_semantic_similarity` metric."
            ' Expected "none", "max", "mean", "median", or "min", but got "%s"',
            reduce,
        )
    skip_words = [word for category in _SKIP_WORDS_GROUPS.values() for word in category]
    refs, preds = _select(items)

    # concepts of every prediction (1024-row batches like the reference's map), + the prediction itself
    concepts: list[list[str]] = []
    for s in range(0, len(preds), 1024):
        batch = _text.concept_extraction_spacy(
            {"prediction": list(preds[s:s + 1024])}, input_column="prediction",
            output_column="prediction_concepts", skip_words=skip_words, remove_prefix_words=True)
        concepts.extend(batch["prediction_concepts"])
    concepts = [list(c) + [p] for c, p in zip(concepts, preds, strict=True)]

    # unique (reference, concept) pairs -> one paired score each (:271-303), then scatter (:306-314)
    pair_index: dict[tuple[str, str], int] = {}
    rows: list[list[int]] = []
    for ref, cs in zip(refs, concepts, strict=True):
        row = []
        for c in cs:
            key = (ref, c)
            j = pair_index.get(key)
            if j is None:
                j = pair_index[key] = len(pair_index)
            row.append(j)
        rows.append(row)
    pairs = list(pair_index)
    sims = _paired_scores([p[0] for p in pairs], [p[1] for p in pairs]) if pairs else np.zeros(0, np.float32)
    per_row = [sims[np.asarray(r, dtype=np.int64)] for r in rows]

    if reduce == "max":
        return float(np.mean([r.max() for r in per_row], dtype=np.float32))
    if reduce == "mean":
        return float(np.mean([r.mean(dtype=np.float32) for r in per_row], dtype=np.float32))
    if reduce == "median":
        # torch.median returns the LOWER of the two middle values for even counts (_group.py:324-328)
        return float(np.mean([np.sort(r)[(r.size - 1) // 2] for r in per_row], dtype=np.float32))
    if reduce == "min":
        return float(np.mean([r.min() for r in per_row], dtype=np.float32))
    return list(zip(concepts, [r.tolist() for r in per_row], strict=True))


def semantic_matching(predictions: list[str], class_names: list[str], k: int = 1,
                      targets: list[int] | None = None, precision: str | None = None) -> dict:
    """North-star extension: match every prediction against every class name.

    Forms the prediction-by-class cosine matrix on the tensor cores with a fused row-wise top-k (the
    matrix is never written to memory) and returns ``{"scores": [N,k], "indices": [N,k],
    "paired": [N] | None}`` as numpy arrays; ``paired[i] = S[i, targets[i]]`` equals what
    ``semantic_similarity`` reports for ``(class_names[targets[i]], predictions[i])``.
    """
    import torch

    engine, tok = _text.get_engine()
    cids, ccu = tok.encode_packed(list(class_names))
    zc = engine.encode(cids, ccu, out="device", precision=precision)
    pids, pcu = tok.encode_packed(list(predictions))
    tg = None if targets is None else torch.as_tensor(np.asarray(targets, dtype=np.int32))
    res = engine.score(pids, pcu, zc, targets=None if tg is None else tg.numpy(), k=k, precision=precision)
    return {"scores": res["scores"], "indices": res["idx"], "paired": res["paired"]}
