"""Win/draw/loss scoring of Elo games by semantic similarity.

Second consumer of the scoring path in the reference: ``eval_ranking.py:192-237`` (``--criterion
semantic_similarity``) encodes ``player_a_response``, ``player_b_response`` and ``reference`` of every sampled
game with ``encode_sentence_bert`` (three ``datasets.map`` calls, ``:196-213``), takes the two paired dots
against the reference (``torch.bmm``, ``:215-220``) and thresholds their difference at 0.05 (``:222-227``):
1 = player A wins, 0 = player A loses, 0.5 = draw. Here each DISTINCT string is encoded once (references repeat
across games of a task, responses across games of a model) and the two paired dots run through ``owc_pair_dot``
with index arrays; everything around it (game sampling, Elo updates, bootstrap) is the reference's own CPU code
and is not mirrored.
"""

from __future__ import annotations

import numpy as np

from . import text as _text

__all__ = ["WIN_THRESHOLD", "semantic_similarity_game_scores"]

WIN_THRESHOLD = 0.05   # eval_ranking.py:223


def semantic_similarity_game_scores(matches: list[dict], threshold: float = WIN_THRESHOLD,
                                    return_similarities: bool = False):
    """Score games ``{"reference", "player_a_response", "player_b_response", ...}`` -> ``list[float]`` of 1 / 0.5 / 0.

    Same values as the ``score`` column ``eval_ranking.py:229-236`` adds; with ``return_similarities`` also the two
    fp32 similarity arrays ``(scores, sim_a, sim_b)``.
    """
    for key in ("reference", "player_a_response", "player_b_response"):
        for m in matches:
            if key not in m:
                raise ValueError(f"{key} missing in dataset")   # what encode_sentence_bert raises, _text.py:172-173
    if not matches:
        return ([], np.zeros(0, np.float32), np.zeros(0, np.float32)) if return_similarities else []
    n = len(matches)
    texts = ([m["reference"] for m in matches] + [m["player_a_response"] for m in matches]
             + [m["player_b_response"] for m in matches])
    engine, _ = _text.get_engine()
    z, inv = _text.encode_unique(texts, out="device")          # one table for all three columns, kept on the device
    ref_i, a_i, b_i = inv[:n], inv[n:2 * n], inv[2 * n:]
    sim_a = np.asarray(engine.pair_dot(z, z, a_index=ref_i, b_index=a_i), dtype=np.float32)
    sim_b = np.asarray(engine.pair_dot(z, z, a_index=ref_i, b_index=b_i), dtype=np.float32)
    diff = sim_a - sim_b
    scores = np.full(n, 0.5, dtype=np.float32)   # default to draw (:224)
    scores[diff > threshold] = 1.0
    scores[diff < -threshold] = 0.0
    out = scores.tolist()
    return (out, sim_a, sim_b) if return_similarities else out
