"""Row-sharded scoring across the GPUs of one box: one process per GPU, ``torch.distributed``.

Every prediction row is independent given the class embeddings, so the path shards by rows
with exactly two exchange steps and no collective inside the encoder or the matcher:

* class-name embeddings: a small class set is encoded once on rank 0 and broadcast; a vocabulary-scale one (config 4:
  117 k names) is split into equal slices, every rank encodes its slice and one ``all_gather`` assembles ``[C, H]`` on
  every rank. Either way the collective is launched asynchronously and each rank encodes its prediction shard while
  the class embeddings are still in flight - they are only needed by the matcher (``Engine.match``);
* results: one packed ``[rows, 2k+1]`` fp32 buffer per rank (scores | idx bit-cast | paired) in ONE ``gather`` into a
  preallocated ``[world, longest, 2k+1]`` tensor on rank 0 - no per-step allocation, no concatenation on the device.

Rank 0 does extra work before its shard (the class encode), so ``split_rows_balanced`` takes a per-rank head start
and hands rank 0 correspondingly fewer rows. The functions only touch tensors the engine returns, so the same code
runs over NCCL with the CUDA engine and over gloo with an oracle-backed engine in CPU tests.
"""

from __future__ import annotations

import numpy as np

__all__ = ["split_rows_balanced", "ClassEmbeddings", "class_embeddings", "broadcast_class_embeddings", "gather_rows",
           "ResultGatherer", "score_shard", "score_rows_sharded", "SHARD_CLASSES_FROM"]

SHARD_CLASSES_FROM = 8192   # class sets at least this large are encoded in slices by all ranks (config 4), smaller ones by rank 0


def _dist():
    import torch.distributed as dist

    return dist


def _world(group=None) -> tuple[bool, int, int]:
    dist = _dist()
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    return distributed, (dist.get_world_size(group) if distributed else 1), (dist.get_rank(group) if distributed else 0)


def split_rows_balanced(weights, world: int, head_start=None) -> list[tuple[int, int]]:
    """Contiguous row blocks ``[(begin, end)] * world`` with near-equal total weight (tokens/chars).

    ``head_start[r]`` (same unit as the weights) is work rank ``r`` has to do before its rows - rank 0's class-name
    encode - and is taken off its share. Deterministic, so every rank computes the same partition without communicating.
    """
    w = np.asarray(weights, dtype=np.float64)
    n = int(w.size)
    if world <= 0:
        raise ValueError("world must be positive")
    if n == 0:
        return [(0, 0)] * world
    pre = np.zeros(world) if head_start is None else np.asarray(head_start, dtype=np.float64)
    if pre.shape != (world,) or (pre < 0).any():
        raise ValueError("head_start needs one non-negative entry per rank")
    csum = np.concatenate([[0.0], np.cumsum(np.maximum(w, 1e-9))])
    share = np.maximum((csum[-1] + pre.sum()) / world - pre, 0.0)      # weight of rows each rank should take
    share *= csum[-1] / max(share.sum(), 1e-30)
    targets = np.cumsum(share)[:-1]
    cuts = [int(np.searchsorted(csum, t, side="left")) for t in targets]
    bounds = [0] + [min(max(c, 0), n) for c in cuts] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


class ClassEmbeddings:
    """Class embeddings that may still be in flight. ``wait()`` orders the CURRENT stream behind the collective (no host
    blocking) and returns the ``[C, H]`` fp32 tensor."""

    def __init__(self, tensor, work=None, n_classes: int | None = None) -> None:
        self._t = tensor
        self._work = work
        self._n = n_classes

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._t if self._n is None else self._t[: self._n]


def class_embeddings(engine, class_ids, class_cu, group=None, precision: str | None = None, src: int = 0,
                     shard_from: int = SHARD_CLASSES_FROM, scratch: dict | None = None) -> ClassEmbeddings:
    """Start computing the ``[C, H]`` class embeddings on every rank (see the module docstring) and return a handle.

    Every rank passes the same ``class_ids`` / ``class_cu`` (the full class list, tokenised). ``scratch`` (a dict the
    caller keeps) lets repeated calls reuse the receive buffers.
    """
    import torch

    dist = _dist()
    distributed, world, rank = _world(group)
    cu = np.ascontiguousarray(class_cu, dtype=np.int64)
    n_cls = int(cu.size) - 1
    h = engine.arch.hidden
    if not distributed:
        return ClassEmbeddings(engine.encode(class_ids, np.asarray(class_cu, np.int32), out="device", precision=precision))
    scratch = scratch if scratch is not None else {}
    if n_cls >= max(shard_from, world):
        per = -(-n_cls // world)                       # equal slices: rank r takes names [r*per, min((r+1)*per, C))
        lo, hi = min(rank * per, n_cls), min((rank + 1) * per, n_cls)
        ids = np.asarray(class_ids)[int(cu[lo]):int(cu[hi])]
        z = engine.encode(ids, (cu[lo:hi + 1] - cu[lo]).astype(np.int32), out="device", precision=precision)
        key = ("zc_all", world, per, h)
        if key not in scratch:
            scratch[key] = (torch.empty(world * per, h, dtype=torch.float32, device=z.device),
                            torch.zeros(per, h, dtype=torch.float32, device=z.device))
        out, mine = scratch[key]
        mine[: hi - lo].copy_(z)
        work = dist.all_gather_into_tensor(out, mine, group=group, async_op=True)
        return ClassEmbeddings(out, work, n_cls)
    if rank == src:
        zc = engine.encode(class_ids, np.asarray(class_cu, np.int32), out="device", precision=precision)
    else:
        key = ("zc_bcast", n_cls, h)
        if key not in scratch:
            dev = getattr(engine, "torch_device", None) or "cpu"
            scratch[key] = torch.empty(n_cls, h, dtype=torch.float32, device=dev)
        zc = scratch[key]
    work = dist.broadcast(zc, src=src, group=group, async_op=True)
    return ClassEmbeddings(zc, work)


def broadcast_class_embeddings(engine, class_ids, class_cu, group=None, precision: str | None = None, src: int = 0):
    """Blocking form of :func:`class_embeddings` (kept for callers that need the tensor at once)."""
    return class_embeddings(engine, class_ids, class_cu, group=group, precision=precision, src=src).wait()


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


class ResultGatherer:
    """Per-shard ``(scores [n,k], idx [n,k], paired [n])`` -> rank ``dst`` in one collective.

    The three results of a row are packed into one fp32 row ``[scores | idx bit-cast | paired]``; the send buffer and
    the ``[world, longest, 2k+1]`` receive tensor are allocated once. ``gather`` returns the receive tensor on ``dst``
    (``None`` elsewhere); ``unpack`` cuts it back into arrays in the original row order.
    """

    def __init__(self, counts: list[int], k: int, device, group=None, dst: int = 0, with_paired: bool = True) -> None:
        import torch

        dist = _dist()
        self.group, self.dst, self.k, self.counts = group, dst, int(k), [int(c) for c in counts]
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.width = 2 * self.k + (1 if with_paired else 0)
        self.with_paired = with_paired
        self.longest = max(max(self.counts), 1)
        self.send = torch.zeros(self.longest, self.width, dtype=torch.float32, device=device)
        self.recv = (torch.empty(self.world, self.longest, self.width, dtype=torch.float32, device=device)
                     if self.rank == dst else None)
        self._views = list(self.recv.unbind(0)) if self.recv is not None else None
        self._host = None

    def gather(self, res: dict):
        import torch

        n, k = res["scores"].shape[0], self.k
        if n != self.counts[self.rank]:
            raise ValueError("shard size does not match the partition")
        self.send[:n, :k] = res["scores"]
        self.send[:n, k:2 * k] = res["idx"].view(torch.float32)
        if self.with_paired:
            self.send[:n, 2 * k] = res["paired"]
        _dist().gather(self.send, self._views, dst=self.dst, group=self.group)
        return self.recv

    def unpack(self, recv=None, to_host: bool = True) -> dict:
        """Rows of all ranks in order. ``to_host``: numpy arrays through one pinned staging buffer (one D2H copy per rank)."""
        import torch

        recv = self.recv if recv is None else recv
        total, k = sum(self.counts), self.k
        if to_host:
            if self._host is None:
                self._host = torch.empty(total, self.width, dtype=torch.float32,
                                         pin_memory=recv.is_cuda)
            lo = 0
            for r, c in enumerate(self.counts):
                self._host[lo:lo + c].copy_(recv[r, :c], non_blocking=True)
                lo += c
            if recv.is_cuda:
                torch.cuda.current_stream(recv.device).synchronize()
            flat = self._host.numpy()
        else:
            flat = torch.cat([recv[r, :c] for r, c in enumerate(self.counts)], dim=0)
        out = {"scores": flat[:, :k], "idx": flat[:, k:2 * k].view(np.int32 if to_host else torch.int32)}
        out["paired"] = flat[:, 2 * k] if self.with_paired else None
        return out


def score_shard(engine, ids, cu, classes: ClassEmbeddings, targets=None, k: int = 1, precision: str | None = None) -> dict:
    """Encode one shard, THEN wait for the class embeddings, then match (results stay on the device)."""
    zp = engine.encode(ids, cu, out="device", precision=precision)
    zc = classes.wait()
    return engine.match(zp, zc, targets=targets, k=k, precision=precision, out="device")


def score_rows_sharded(engine, tokenizer, class_names: list[str], predictions: list[str],
                       targets: list[int] | None = None, k: int = 1, group=None, precision: str | None = None,
                       head_start: float = 0.0) -> dict | None:
    """Whole path over all ranks. Every rank passes the same full lists; each tokenises and scores only its own
    contiguous block (balanced by string length, a proxy for token count; ``head_start`` in the same unit is taken off
    rank 0's share for its class encode). Rank 0 returns ``{"scores", "idx", "paired"}`` as numpy arrays in the
    original row order; other ranks return ``None``."""
    distributed, world, rank = _world(group)
    cids, ccu = tokenizer.encode_packed(list(class_names))
    classes = class_embeddings(engine, cids, ccu, group=group, precision=precision)

    blocks = split_rows_balanced([len(p) + 2 for p in predictions], world,
                                 head_start=[head_start] + [0.0] * (world - 1) if world > 1 else None)
    b0, b1 = blocks[rank]
    pids, pcu = tokenizer.encode_packed(list(predictions[b0:b1]))
    tg = None if targets is None else np.asarray(targets[b0:b1], dtype=np.int32)
    if not distributed:
        res = engine.score(pids, pcu, classes.wait(), targets=tg, k=k, precision=precision, out="device")
        return {key: (None if v is None else v.cpu().numpy()) for key, v in res.items() if key != "embeds"}
    res = score_shard(engine, pids, pcu, classes, targets=tg, k=k, precision=precision)
    g = ResultGatherer([e - b for b, e in blocks], k, res["scores"].device, group=group, with_paired=tg is not None)
    recv = g.gather(res)
    if rank != 0:
        return None
    out = g.unpack(recv)
    return {key: (None if v is None else np.array(v)) for key, v in out.items()}
