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

from .pipeline import RowParts, score_jobs_pipelined

__all__ = ["score_sweep"]


def score_sweep(engine, tokenizer, tasks: dict, k: int = 1, precision: str | None = None) -> dict:
    """Score a sweep.

    ``tasks``: ``{task_name: {"class_names": [str], "files": {file_key: {"predictions": [str], "targets": [int] | None}}}}``
    Returns ``{task_name: {file_key: {"scores" [n,k], "idx" [n,k], "paired" [n] | None}}}`` (numpy, row order kept).
    """
    out: dict = {}
    plan = []          # (task, keys, counts, with targets)
    jobs = []          # (flat predictions, class embeddings on the device, targets)
    for task, spec in tasks.items():
        names = list(spec["class_names"])
        files = spec["files"]
        out[task] = {}
        if not files:
            continue
        if k > len(names):
            raise ValueError(f"task {task!r}: k={k} exceeds its {len(names)} class names")
        keys = list(files)
        rows = [r if isinstance(r, list) else list(r) for r in (files[key]["predictions"] for key in keys)]
        have_t = [files[key].get("targets") is not None for key in keys]
        if any(have_t) and not all(have_t):
            raise ValueError(f"task {task!r}: either every file carries targets or none does")
        counts = [len(r) for r in rows]
        if sum(counts) == 0:
            for key in keys:
                out[task][key] = {"scores": np.zeros((0, k), np.float32), "idx": np.zeros((0, k), np.int32),
                                  "paired": np.zeros(0, np.float32) if all(have_t) else None}
            continue
        tg = None
        if all(have_t):
            tg = np.concatenate([np.asarray(files[key]["targets"], dtype=np.int32) for key in keys])
            if tg.shape[0] != sum(counts) or (tg.size and (tg.min() < 0 or tg.max() >= len(names))):
                raise ValueError(f"task {task!r}: targets do not match the predictions / class names")
        cids, ccu = tokenizer.encode_packed(names)
        zc = engine.encode(cids, ccu, out="device", precision=precision)   # each class set once (tiny; queued ahead of the rows)
        plan.append((task, keys, counts, tg is not None))
        jobs.append((RowParts(rows), zc, tg))    # concatenated by the pipeline's worker thread
    # one pipeline over every task: the tokenisation of the next row chunk - of this task or the next one - overlaps the
    # device work of the previous one, and the device is drained once, at the end
    results = score_jobs_pipelined(engine, tokenizer, jobs, k=k, precision=precision)
    for (task, keys, counts, with_t), res in zip(plan, results):
        lo = 0
        for key, c in zip(keys, counts):
            out[task][key] = {"scores": res["scores"][lo:lo + c], "idx": res["idx"][lo:lo + c],
                              "paired": res["paired"][lo:lo + c] if with_t else None}
            lo += c
    return out
