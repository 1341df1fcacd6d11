"""Row sharding across ranks on CPU: world_size-2 gloo processes with an oracle-backed
engine standing in for the CUDA engine (tests may use the oracle; the product may not). Checks the
partition, the class-embedding broadcast, the padded gather and that the sharded result equals the
single-process result row for row."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import owc_b200  # noqa: F401
from owc_b200 import sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_rows_balanced_properties():
    rng = np.random.default_rng(0)
    for n, world in [(0, 4), (1, 4), (7, 2), (1000, 8), (1000, 3)]:
        w = rng.integers(3, 40, n)
        blocks = sharded.split_rows_balanced(w, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(b[1] == nb[0] for b, nb in zip(blocks, blocks[1:])) and all(b0 <= b1 for b0, b1 in blocks)
        if n >= 100:
            loads = [w[b0:b1].sum() for b0, b1 in blocks]
            assert max(loads) - min(loads) <= 2 * w.max()      # balanced by weight, not by row count
    assert sharded.split_rows_balanced([5, 5, 5, 5], 2) == [(0, 2), (2, 4)]
    # a head start (rank 0's class-name encode) comes off that rank's share of the rows
    w = np.full(1000, 10.0)
    blocks = sharded.split_rows_balanced(w, 4, head_start=[400.0, 0, 0, 0])
    loads = [w[b0:b1].sum() for b0, b1 in blocks]
    assert blocks[0][0] == 0 and blocks[-1][1] == 1000 and abs(loads[0] + 400 - loads[1]) <= 10 and abs(loads[1] - loads[3]) <= 10
    assert sharded.split_rows_balanced(w, 2, head_start=[1e9, 0]) == [(0, 0), (0, 1000)]   # never negative
    with pytest.raises(ValueError):
        sharded.split_rows_balanced(w, 2, head_start=[1.0])
    with pytest.raises(ValueError):
        sharded.split_rows_balanced([1], 0)


class _OracleEngine:
    """CPU stand-in with the Engine methods the sharding layer uses (encode/score, out='device')."""

    def __init__(self, arch, sd):
        self.arch, self.sd = arch, sd
        self.encode_calls = 0

    def _pad(self, ids, cu):
        n = len(cu) - 1
        lens = np.diff(cu)
        t = max(int(lens.max()) if n else 1, 1)
        ids_p = np.full((n, t), self.arch.pad_id, dtype=np.int64)
        mask = np.zeros((n, t), dtype=np.int64)
        for i in range(n):
            ids_p[i, :lens[i]] = ids[cu[i]:cu[i + 1]]
            mask[i, :lens[i]] = 1
        return torch.from_numpy(ids_p), torch.from_numpy(mask)

    def encode(self, ids, cu, out="host", precision=None):
        from oracle import encoder_ref

        self.encode_calls += 1
        n = len(cu) - 1
        if n == 0:
            return torch.zeros(0, self.arch.hidden)
        z = encoder_ref.encode_padded(self.sd, self.arch, *self._pad(np.asarray(ids), np.asarray(cu)))
        return z if out == "device" else z.numpy()

    def score(self, ids, cu, zc, targets=None, k=1, precision=None, out="host", want_embeds=False):
        from oracle import encoder_ref

        zp = self.encode(ids, cu, out="device")
        if zp.shape[0] == 0:
            return {"scores": torch.zeros(0, k), "idx": torch.zeros(0, k, dtype=torch.int32),
                    "paired": torch.zeros(0) if targets is not None else None, "embeds": None}
        s, i, _ = encoder_ref.sim_topk(zp, zc, k)
        paired = None
        if targets is not None:
            paired = (zp * zc[torch.from_numpy(np.asarray(targets)).long()]).sum(1)
        return {"scores": s.contiguous(), "idx": i.to(torch.int32).contiguous(), "paired": paired, "embeds": None}

    def match(self, zp, zc, targets=None, k=1, precision=None, out="host"):
        from oracle import encoder_ref

        if zp.shape[0] == 0:
            return {"scores": torch.zeros(0, k), "idx": torch.zeros(0, k, dtype=torch.int32),
                    "paired": torch.zeros(0) if targets is not None else None}
        s, i, _ = encoder_ref.sim_topk(zp, zc, k)
        paired = None
        if targets is not None:
            paired = (zp * zc[torch.from_numpy(np.asarray(targets)).long()]).sum(1)
        return {"scores": s.contiguous(), "idx": i.to(torch.int32).contiguous(), "paired": paired}


def _small_setup():
    from owc_b200 import arch as A
    from owc_b200 import synth
    from owc_b200.tokenizer import HostTokenizer
    from owc_b200.weights import random_state_dict

    arch = A.EncoderArch("tiny-bert", A.KIND_BERT, layers=1, hidden=256, heads=8, inter=512, vocab=2048, max_pos=64,
                         type_vocab=2, ln_eps=1e-12)
    vocab = synth.build_vocab("bert", arch.vocab)
    words = synth.vocab_words(vocab)
    classes = synth.make_class_names(12, words)
    preds, targets = synth.make_predictions(61, classes, words)
    return arch, random_state_dict(arch), HostTokenizer(vocab, "bert"), classes, preds, targets


def _worker(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        arch, sd, tok, classes, preds, targets = _small_setup()
        eng = _OracleEngine(arch, sd)
        out = sharded.score_rows_sharded(eng, tok, classes, preds, targets, k=3, head_start=40.0)
        calls = eng.encode_calls
        # vocabulary-scale class sets: every rank encodes an equal slice, one all_gather assembles [C, H] everywhere
        cids, ccu = tok.encode_packed(classes)
        zc_sliced = sharded.class_embeddings(eng, cids, ccu, shard_from=4).wait()
        zc_full = _OracleEngine(arch, sd).encode(cids, ccu, out="device")
        slice_err = float((zc_sliced - zc_full).abs().max())
        slice_calls = eng.encode_calls - calls
        # packed gather of unequal shards, without targets
        counts = [5, 3]
        g = sharded.ResultGatherer(counts, 2, "cpu", with_paired=False)
        n = counts[rank]
        recv = g.gather({"scores": torch.full((n, 2), float(rank)), "idx": torch.full((n, 2), 7 + rank, dtype=torch.int32)})
        if rank == 0:
            un = g.unpack(recv)
            assert un["paired"] is None and un["scores"].shape == (8, 2)
            assert (un["scores"][:5] == 0).all() and (un["scores"][5:] == 1).all()
            assert (un["idx"][:5] == 7).all() and (un["idx"][5:] == 8).all()
            queue.put({"scores": out["scores"], "idx": out["idx"], "paired": out["paired"], "calls": calls,
                       "slice_err": slice_err, "slice_calls": slice_calls})
        else:
            assert out is None and recv is None
            queue.put({"calls": calls, "rank": rank, "slice_err": slice_err, "slice_calls": slice_calls})
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_scoring_world2_matches_single_process():
    arch, sd, tok, classes, preds, targets = _small_setup()
    single = sharded.score_rows_sharded(_OracleEngine(arch, sd), tok, classes, preds, targets, k=3)
    assert single["scores"].shape == (61, 3)

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = next(r for r in results if "scores" in r)
    r1 = next(r for r in results if "rank" in r)
    assert np.array_equal(r0["idx"], single["idx"])
    assert np.abs(r0["scores"] - single["scores"]).max() < 1e-6
    assert np.abs(r0["paired"] - single["paired"]).max() < 1e-6
    # rank 0 encodes the class names (1) + its rows (1); rank 1 only encodes its rows: the class embeddings reach it
    # by broadcast, not by re-encoding
    assert r0["calls"] == 2 and r1["calls"] == 1
    # sliced class encode: one encode call per rank (its slice), identical embeddings on both ranks
    assert r0["slice_calls"] == 1 and r1["slice_calls"] == 1 and r0["slice_err"] < 1e-6 and r1["slice_err"] < 1e-6


# ---- the CLI mirror under a multi-process launcher: files dealt to ranks, one report on rank 0 -------------------
class _PairEngine(_OracleEngine):
    """Adds the paired dot the metric bodies call."""

    def pair_dot(self, a, b, a_index=None, b_index=None):
        a, b = torch.as_tensor(a), torch.as_tensor(b)
        if a_index is not None:
            a = a[torch.from_numpy(np.asarray(a_index)).long()]
        if b_index is not None:
            b = b[torch.from_numpy(np.asarray(b_index)).long()]
        return (a * b).sum(1).numpy()

    def close(self):
        pass


def _write_cli_logs(root, classes, preds, targets):
    import json

    sizes = [25, 9, 17, 10]                                    # different sizes: the deal is by bytes on disk
    lo = 0
    for m, n in enumerate(sizes):
        run = os.path.join(root, "task", f"model-{m}")
        os.makedirs(run)
        with open(os.path.join(run, "r_samples_task.jsonl"), "w", encoding="utf-8") as f:
            for i in range(lo, lo + n):
                f.write(json.dumps({"doc_id": i, "target": classes[targets[i]], "filtered_resps": [preds[i]]}) + "\n")
        lo += n


def _cli_worker(rank, world, port, root, queue):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    torch.set_num_threads(1)
    from owc_b200 import eval_metrics as EM
    from owc_b200 import text as T

    arch, sd, tok, *_ = _small_setup()
    T.sentence_bert_model, T.sentence_bert_processor = _PairEngine(arch, sd), tok
    try:
        out = EM.main(EM.build_parser().parse_args(["-i", root, "-m", "semantic_similarity", "--io", "native",
                                                    "--log-level", "WARNING"]))
        queue.put({"rank": rank, "out": out, "default_rank": T.default_rank})
    finally:
        T.sentence_bert_model = T.sentence_bert_processor = None
        if dist.is_initialized():
            dist.destroy_process_group()


def test_cli_deals_files_to_ranks_world2(tmp_path, monkeypatch):
    import shutil

    from owc_b200 import eval_metrics as EM
    from owc_b200 import text as T

    arch, sd, tok, classes, preds, targets = _small_setup()
    one, two = str(tmp_path / "one" / "logs"), str(tmp_path / "two" / "logs")
    _write_cli_logs(one, classes, preds, targets)
    shutil.copytree(one, two)
    files = EM.find_input_files(two)
    deal = EM.split_files_balanced(files, 2)
    assert sorted(deal[0] + deal[1]) == files and all(deal) and EM.split_files_balanced(files, 1) == [files]
    loads = [sum(os.path.getsize(f) for f in part) for part in deal]
    assert abs(loads[0] - loads[1]) <= max(os.path.getsize(f) for f in files)
    assert EM.split_files_balanced(files, 2) == deal and len(EM.split_files_balanced(files[:1], 3)) == 3

    monkeypatch.setattr(T, "sentence_bert_model", _PairEngine(arch, sd))
    monkeypatch.setattr(T, "sentence_bert_processor", tok)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    single = EM.main(EM.build_parser().parse_args(["-i", one, "-m", "semantic_similarity", "--io", "native",
                                                   "--log-level", "WARNING"]))
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cli_worker, args=(r, 2, port, two, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = {r["rank"]: r for r in (queue.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0]["out"] == single and results[1]["out"] == {}            # one report, on rank 0
    assert results[0]["default_rank"] == 0 and results[1]["default_rank"] == 1  # each rank on its own GPU
    for f in EM.find_input_files(one):
        assert open(f, "rb").read() == open(f.replace(one, two), "rb").read()  # every file scored exactly once


def _cli_failing_worker(rank, world, port, root, queue):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    torch.set_num_threads(1)
    from owc_b200 import eval_metrics as EM
    from owc_b200 import text as T

    arch, sd, tok, *_ = _small_setup()
    T.sentence_bert_model, T.sentence_bert_processor = _PairEngine(arch, sd), tok
    try:
        EM.main(EM.build_parser().parse_args(["-i", root, "-m", "semantic_similarity", "--io", "native", "--log-level", "ERROR"]))
        queue.put({"rank": rank, "error": None})
    except BaseException as exc:   # noqa: BLE001
        queue.put({"rank": rank, "error": f"{type(exc).__name__}: {exc}", "group_alive": dist.is_initialized()})
    finally:
        T.sentence_bert_model = T.sentence_bert_processor = None


def test_cli_rank_failure_reaches_rank0_instead_of_hanging(tmp_path):
    """One rank's file is broken: that rank raises, and rank 0 (which is waiting in the gather of the per-file outputs)
    gets the error and raises too, instead of blocking forever; the process group is torn down on both."""
    from owc_b200 import eval_metrics as EM

    arch, sd, tok, classes, preds, targets = _small_setup()
    root = str(tmp_path / "logs")
    _write_cli_logs(root, classes, preds, targets)
    files = EM.find_input_files(root)
    deal = EM.split_files_balanced(files, 2)
    with open(deal[1][0], "a", encoding="utf-8") as f:          # a file that rank 1 owns
        f.write('{"doc_id": 1, "target": "x", "filtered_resps": ["y"]} trailing junk\n')
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cli_failing_worker, args=(r, 2, port, root, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = {r["rank"]: r for r in (queue.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert "trailing data" in results[1]["error"] or "Expected object or value" in results[1]["error"] or "ValueError" in results[1]["error"]
    assert results[0]["error"] is not None and "rank 1" in results[0]["error"]
    assert not results[0]["group_alive"] and not results[1]["group_alive"]
