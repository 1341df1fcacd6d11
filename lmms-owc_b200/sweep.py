"""Multi-task sweep: many (model, task) prediction files re-scored in one run (BASELINE config 5).

The reference re-scores a sweep by running ``eval_metrics.py`` once per samples file (``eval_metrics.py:36-133``):
every file re-loads the encoder path and re-encodes its task's class names once per ROW. Files of the
same task share one class set, so here each distinct class set is encoded ONCE, the predictions of every file that
uses it are concatenated into one row stream (one ``owc_score`` call, which keeps the fused encoder's launches full
instead of paying a partial wave per file) and the per-row results are cut back per file. Per-file outputs are
identical to scoring each file on its own: rows are independent given the class embeddings.
"""

from __future__ import annotations

import numpy as np

from .pipeline import score_strings_pipelined

__all__ = ["score_sweep"]


def score_sweep(engine, tokenizer, tasks: dict, k: int = 1, precision: str | None = None) -> dict:
    """Score a sweep.

    ``tasks``: ``{task_name: {"class_names": [str], "files": {file_key: {"predictions": [str], "targets": [int] | None}}}}``
    Returns ``{task_name: {file_key: {"scores" [n,k], "idx" [n,k], "paired" [n] | None}}}`` (numpy, row order kept).
    """
    out: dict = {}
    for task, spec in tasks.items():
        names = list(spec["class_names"])
        files = spec["files"]
        out[task] = {}
        if not files:
            continue
        if k > len(names):
            raise ValueError(f"task {task!r}: k={k} exceeds its {len(names)} class names")
        cids, ccu = tokenizer.encode_packed(names)
        zc = engine.encode(cids, ccu, out="device", precision=precision)
        keys = list(files)
        rows = [list(files[key]["predictions"]) for key in keys]
        have_t = [files[key].get("targets") is not None for key in keys]
        if any(have_t) and not all(have_t):
            raise ValueError(f"task {task!r}: either every file carries targets or none does")
        counts = [len(r) for r in rows]
        flat = [p for r in rows for p in r]
        if not flat:
            for key in keys:
                out[task][key] = {"scores": np.zeros((0, k), np.float32), "idx": np.zeros((0, k), np.int32),
                                  "paired": np.zeros(0, np.float32) if all(have_t) else None}
            continue
        tg = None
        if all(have_t):
            tg = np.concatenate([np.asarray(files[key]["targets"], dtype=np.int32) for key in keys])
            if tg.shape[0] != len(flat) or (tg.size and (tg.min() < 0 or tg.max() >= len(names))):
                raise ValueError(f"task {task!r}: targets do not match the predictions / class names")
        # tokenisation of the next row chunk overlaps the device work of the previous one
        res = score_strings_pipelined(engine, tokenizer, flat, zc, targets=tg, k=k, precision=precision)
        lo = 0
        for key, c in zip(keys, counts):
            out[task][key] = {"scores": res["scores"][lo:lo + c], "idx": res["idx"][lo:lo + c],
                              "paired": None if tg is None else res["paired"][lo:lo + c]}
            lo += c
    return out
