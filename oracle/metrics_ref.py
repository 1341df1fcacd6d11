"""ORACLE (test infrastructure): CPU restatement of the three similarity aggregations.

Follows ``src/data/metrics/_group.py`` of the reference: ``semantic_similarity`` (:488-544),
``mean_average_semantic_similarity`` (:392-458) and ``concept_semantic_similarity`` (:176-334),
minus the Arrow round trip (``datasets.map``) which carries values but computes nothing. The
embedding step is delegated to ``oracle.encoder_ref.encode_strings``. Pinned against the
reference's own outputs by ``tests/test_oracle.py`` using ``tests/golden/*.npz``.
"""

from __future__ import annotations

import torch

from . import encoder_ref

__all__ = ["select_items", "semantic_similarity", "mean_average_semantic_similarity",
           "concept_semantic_similarity", "simple_concepts", "SKIP_WORDS"]

THRESHOLDS = [0.5, 0.6, 0.7, 0.8, 0.9]

# _group.py:207-234
SKIP_WORDS = (
    ["1", "2", "3", "4", "5", "6", "7", "8", "9", "10"]
    + ["one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten"]
    + ["*"] + ["a", "the"] + ["image", "object", "photo", "type", "this photo"]
    + ["it", "they", "them"] + ["that", "this", "those"]
    + ["which", "who", "whom", "whose", "where", "when", "what", "why", "how"] + ["some"]
)


def select_items(items: list) -> tuple[list[str], list[str]]:
    """``ref[0]`` if the reference is a list, ``pred[-1]`` if the prediction is (_group.py:515-519)."""
    refs = list(zip(*items, strict=True))[0]
    preds = list(zip(*items, strict=True))[1]
    refs = [ref[0] if isinstance(ref, list) else ref for ref in refs]
    preds = [pred[-1] if isinstance(pred, list) else pred for pred in preds]
    return refs, preds


def _paired_scores(encode, refs: list[str], preds: list[str]) -> torch.Tensor:
    # The reference encodes the reference column row by row (no dedup, F5) - same values.
    return encoder_ref.paired_dot(encode(refs), encode(preds))


def semantic_similarity(items: list, encode, reduce: str = "mean"):
    """_group.py:488-544. ``encode(list[str]) -> [N, H]`` is the operator (oracle or candidate)."""
    if reduce not in ["none", "mean"]:
        raise ValueError("Unknown `reduce` value for `semantic_similarity` metric.")
    refs, preds = select_items(items)
    s = _paired_scores(encode, refs, preds)
    return s.mean().item() if reduce == "mean" else s.tolist()


def mean_average_semantic_similarity(items: list, encode, reduce: str = "mean"):
    """_group.py:392-458."""
    if reduce not in ["none", "mean"]:
        raise ValueError("Unknown `reduce` value for `mean_average_semantic_similarity` metric.")
    refs, preds = select_items(items)
    s = _paired_scores(encode, refs, preds)
    outputs = {}
    if reduce == "mean":
        for t in THRESHOLDS:
            outputs[f"semantic_similarity@{t}"] = (s >= t).float().mean().item()
        outputs["semantic_similarity@avg"] = torch.Tensor(list(outputs.values())).mean().item()
        return outputs
    for t in THRESHOLDS:
        outputs[f"semantic_similarity@{t}"] = (s >= t).int().tolist()
    outputs["semantic_similarity@avg"] = torch.Tensor(list(outputs.values())).mean(dim=0).tolist()
    return outputs


def simple_concepts(batch: dict, **kwargs) -> dict:
    """Stand-in for ``concept_extraction_spacy`` (_text.py:18-140): spaCy is not installable offline.

    Same signature and kwargs, same prefix-stripping / skip-word / dedup rules; the noun-chunker is
    replaced by a split on commas and " and ". Used identically on the reference side (monkeypatched
    in ``make_golden.py``) and on the candidate side, so everything downstream of extraction
    (pair building, dedup, encode, dot, scatter, reduce) is still compared against the reference.
    """
    input_column = kwargs.pop("input_column", "text")
    output_column = kwargs.pop("output_column", f"{input_column}_concepts")
    skip_words = kwargs.pop("skip_words", [])
    remove_prefix_words = kwargs.pop("remove_prefix_words", False)
    if input_column not in batch:
        raise ValueError(f"{input_column} missing in dataset")
    all_concepts = []
    for text in batch[input_column]:
        concepts: list[str] = []
        for chunk in text.replace(" and ", ",").split(","):
            concept = chunk.strip().lower()
            if not concept:
                continue
            if remove_prefix_words:
                for term in ["a", "an", "the", "his", "her", "its", "their"]:
                    if concept.startswith(term + " "):
                        concept = concept[len(term) + 1:]
                        break
                if concept in skip_words:
                    continue
            if concept not in concepts:
                concepts.append(concept)
        all_concepts.append(concepts)
    batch[output_column] = all_concepts
    return batch


def concept_semantic_similarity(items: list, encode, extract=simple_concepts, reduce: str = "max"):
    """_group.py:176-334 with the concept extractor injected."""
    if reduce not in ["none", "max", "mean", "median", "min"]:
        raise ValueError("Unknown `reduce` value for `concept_semantic_similarity` metric.")
    refs, preds = select_items(items)
    batch = extract({"prediction": list(preds)}, input_column="prediction",
                    output_column="prediction_concepts", skip_words=list(SKIP_WORDS),
                    remove_prefix_words=True)
    concepts = [c + [p] for c, p in zip(batch["prediction_concepts"], preds, strict=True)]  # :260-262
    pairs = [[(r, c) for c in cs] for r, cs in zip(refs, concepts, strict=True)]
    unique = sorted(set(sum(pairs, [])))  # order is irrelevant to the result
    to_idx = {" | ".join(p): i for i, p in enumerate(unique)}
    sims = _paired_scores(encode, [p[0] for p in unique], [p[1] for p in unique])
    rows = [sims[torch.tensor([to_idx[" | ".join(p)] for p in ps])] for ps in pairs]
    if reduce == "max":
        return torch.stack([r.max() for r in rows]).mean().item()
    if reduce == "mean":
        return torch.stack([r.mean() for r in rows]).mean().item()
    if reduce == "median":
        return torch.stack([r.median() for r in rows]).mean().item()
    if reduce == "min":
        return torch.stack([r.min() for r in rows]).mean().item()
    return list(zip(concepts, [r.tolist() for r in rows], strict=True))


def ranking_game_scores(encode, matches: list[dict], threshold: float = 0.05):
    """``eval_ranking.py:192-237`` restated: three encodes, two paired dots (``torch.bmm`` on ``[N,1,D] x [N,D,1]``),
    difference thresholded at 0.05 -> 1 / 0.5 / 0. ``encode(list[str]) -> torch.Tensor [n, D]``.
    Returns ``(scores list[float], sim_a, sim_b)``. Test infrastructure only."""
    import torch

    refs_z = encode([m["reference"] for m in matches]).unsqueeze(1)
    a_z = encode([m["player_a_response"] for m in matches]).unsqueeze(2)
    b_z = encode([m["player_b_response"] for m in matches]).unsqueeze(2)
    a_scores = torch.bmm(refs_z, a_z).squeeze()
    b_scores = torch.bmm(refs_z, b_z).squeeze()
    diff = a_scores - b_scores
    scores = torch.zeros_like(diff) + 0.5
    scores[diff > threshold] = 1
    scores[diff < -threshold] = 0
    return scores.tolist(), a_scores, b_scores
