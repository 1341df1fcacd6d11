"""Row-sharded scoring across the GPUs of one box: one process per GPU, ``torch.distributed``.

Every prediction row is independent given the class embeddings, so the path
shards by rows with exactly two exchange steps: the class-name embeddings are computed once (rank 0)
and broadcast, and the per-shard results are gathered on rank 0 at the end. There is no collective
inside the encoder or the matcher. The functions only touch tensors the engine returns, so the same
code runs over NCCL with the CUDA engine and over gloo with an oracle-backed engine in CPU tests.
"""

from __future__ import annotations

import numpy as np

__all__ = ["split_rows_balanced", "broadcast_class_embeddings", "gather_rows", "score_rows_sharded"]


def _dist():
    import torch.distributed as dist

    return dist


def split_rows_balanced(weights, world: int) -> list[tuple[int, int]]:
    """Contiguous row blocks ``[(begin, end)] * world`` with near-equal total weight (tokens/chars).

    Deterministic, so every rank computes the same partition without communicating.
    """
    w = np.asarray(weights, dtype=np.float64)
    n = int(w.size)
    if world <= 0:
        raise ValueError("world must be positive")
    if n == 0:
        return [(0, 0)] * world
    csum = np.concatenate([[0.0], np.cumsum(np.maximum(w, 1e-9))])
    targets = csum[-1] * np.arange(1, world) / world
    cuts = [int(np.searchsorted(csum, t, side="left")) for t in targets]
    bounds = [0] + [min(max(c, 0), n) for c in cuts] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def broadcast_class_embeddings(engine, class_ids, class_cu, group=None, precision: str | None = None, src: int = 0):
    """Rank ``src`` encodes the class names; everyone receives the ``[C, H]`` fp32 embeddings."""
    import torch

    dist = _dist()
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    n_cls = int(len(class_cu)) - 1
    if rank == src or not distributed:
        zc = engine.encode(class_ids, class_cu, out="device", precision=precision)
    else:
        ref = engine.encode(np.zeros(0, np.int32), np.zeros(1, np.int32), out="device")  # dtype/device probe
        zc = torch.empty(n_cls, engine.arch.hidden, dtype=torch.float32, device=ref.device)
    if distributed:
        dist.broadcast(zc, src=src, group=group)
    return zc


def gather_rows(t, counts: list[int], group=None, dst: int = 0):
    """Gather per-rank row blocks of unequal length onto ``dst`` (padded to the longest, then trimmed)."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    longest = max(max(counts), 1)
    pad = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def score_rows_sharded(engine, tokenizer, class_names: list[str], predictions: list[str],
                       targets: list[int] | None = None, k: int = 1, group=None, precision: str | None = None) -> dict | None:
    """Whole path over all ranks. Every rank passes the same full lists; each tokenises and scores
    only its own contiguous block (balanced by string length, a proxy for token count). Rank 0
    returns ``{"scores", "idx", "paired"}`` as numpy arrays in the original row order; other ranks
    return ``None``."""
    dist = _dist()
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0

    cids, ccu = tokenizer.encode_packed(list(class_names)) if rank == 0 or not distributed else (
        np.zeros(0, np.int32), np.zeros(len(class_names) + 1, np.int32))
    zc = broadcast_class_embeddings(engine, cids, ccu, group=group, precision=precision)

    blocks = split_rows_balanced([len(p) + 2 for p in predictions], world)
    b0, b1 = blocks[rank]
    pids, pcu = tokenizer.encode_packed(list(predictions[b0:b1]))
    tg = None if targets is None else np.asarray(targets[b0:b1], dtype=np.int32)
    res = engine.score(pids, pcu, zc, targets=tg, k=k, precision=precision, out="device")
    if not distributed:
        return {key: (None if v is None else v.cpu().numpy()) for key, v in res.items() if key != "embeds"}
    counts = [e - b for b, e in blocks]
    out = {}
    for key in ("scores", "idx", "paired"):
        if res[key] is None:
            out[key] = None
            continue
        g = gather_rows(res[key], counts, group=group)
        out[key] = None if g is None else g.cpu().numpy()
    return out if rank == 0 else None
