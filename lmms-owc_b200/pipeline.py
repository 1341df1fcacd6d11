"""Strings -> results with host tokenisation overlapped with device scoring.

Tokenisation stays on the host (north star); once the device scores ~4.5 M rows/s the tokenizer's ~5 M rows/s is
the same order, so running the two back to back halves the end-to-end rate. Here the prediction list is cut into
row chunks: a worker thread tokenises chunk i+1 (``owc_tokenize_packed`` releases the GIL, and so does
``owc_score``) while the calling thread scores chunk i. Results are identical to one un-chunked ``Engine.score``
call: rows are independent given the class embeddings.
"""

from __future__ import annotations

import queue
import threading

import numpy as np

__all__ = ["RowParts", "score_jobs_pipelined", "score_strings_pipelined"]


class RowParts(list):
    """A job's predictions as a list of row lists (one per file): the pipeline's worker thread concatenates them when it
    reaches the job, so flattening a million strings is not serial host work in front of the device."""


def score_jobs_pipelined(engine, tokenizer, jobs: list, k: int = 1, chunk_rows: int = 20_000, precision: str | None = None,
                         first_chunk_rows: int | None = None) -> list:
    """Several (predictions, class embeddings, targets) jobs through ONE tokenise -> score pipeline.

    ``jobs``: ``[(predictions: list[str], zc_device, targets | None), ...]``. The worker thread walks the row chunks of all
    jobs in order, so the only tokenisation the device waits for is the first chunk of the FIRST job; a sweep over many
    class sets (`sweep.score_sweep`) no longer restarts the pipeline - and waits for the device to drain - once per set.
    Returns one ``{"scores" [n,k], "idx" [n,k], "paired" [n] or None}`` per job (numpy, original row order).
    """
    chunk_rows = max(1, int(chunk_rows))
    outs, tgs, bounds = [], [], []
    for j, (predictions, _, targets) in enumerate(jobs):
        if not isinstance(predictions, list):
            raise NotImplementedError
        n = sum(len(r) for r in predictions) if isinstance(predictions, RowParts) else len(predictions)
        tg = None if targets is None else np.ascontiguousarray(targets, dtype=np.int32)
        if tg is not None and tg.shape[0] != n:
            raise ValueError("targets must have one entry per prediction")
        tgs.append(tg)
        outs.append({"scores": np.empty((n, k), np.float32), "idx": np.empty((n, k), np.int32),
                     "paired": None if tg is None else np.empty(n, np.float32)})
        if n == 0:
            continue
        first = chunk_rows
        if not bounds and first_chunk_rows is not None:
            first = max(1, int(first_chunk_rows))
        first = min(n, first)
        bounds += [(j, 0, first)] + [(j, lo, min(lo + chunk_rows, n)) for lo in range(first, n, chunk_rows)]
    if not bounds:
        return outs
    ready: queue.Queue = queue.Queue(maxsize=2)   # at most two tokenised chunks ahead of the device

    def producer() -> None:
        try:
            cur, flat = -1, None
            for j, lo, hi in bounds:
                if j != cur:
                    cur, flat = j, jobs[j][0]
                    if isinstance(flat, RowParts):
                        flat = [p for r in flat for p in r]
                ready.put((j, lo, hi, tokenizer.encode_packed(flat[lo:hi])))
        except BaseException as exc:          # surfaced by the consumer
            ready.put(exc)

    worker = threading.Thread(target=producer, daemon=True)
    worker.start()
    parts = []
    try:
        for _ in bounds:
            item = ready.get()
            if isinstance(item, BaseException):
                raise item
            j, lo, hi, (ids, cu) = item
            tg = tgs[j]
            # results stay on the device: the call returns as soon as the launches are queued, so the device never
            # waits for the host between chunks; one copy at the end brings everything back
            parts.append((j, lo, hi, engine.score(ids, cu, jobs[j][1], targets=None if tg is None else tg[lo:hi], k=k,
                                                  precision=precision, out="device")))
    finally:
        worker.join(timeout=60)
    if hasattr(engine, "sync"):
        engine.sync()   # also raises if a chunk carried an out-of-vocabulary id

    def host(x):
        return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)

    for j, lo, hi, res in parts:
        out = outs[j]
        out["scores"][lo:hi] = host(res["scores"])
        out["idx"][lo:hi] = host(res["idx"])
        if tgs[j] is not None:
            out["paired"][lo:hi] = host(res["paired"])
    return outs


def score_strings_pipelined(engine, tokenizer, predictions: list[str], zc_device, targets=None, k: int = 1,
                            chunk_rows: int = 20_000, precision: str | None = None,
                            first_chunk_rows: int | None = None) -> dict:
    """Tokenise + encode + match ``predictions`` against class embeddings ``zc_device`` (CUDA fp32 ``[C, H]``).

    ``first_chunk_rows`` (default: ``chunk_rows``) sizes the first chunk separately: its tokenisation is the only one
    the device has to wait for, so a short first chunk followed by long ones starts the device early and keeps the
    per-chunk launch overhead low. Returns numpy ``{"scores" [n,k], "idx" [n,k], "paired" [n] or None}`` in the
    original row order.
    """
    if not isinstance(predictions, list):
        raise NotImplementedError
    return score_jobs_pipelined(engine, tokenizer, [(predictions, zc_device, targets)], k=k, chunk_rows=chunk_rows,
                                precision=precision, first_chunk_rows=first_chunk_rows)[0]
