"""Host-side logic on CPU: tokenizer/packing, weight blob, synthetic data, plugin registry and the
metric bodies driven by an oracle-backed fake engine (tests may use the oracle; the product may not)."""

import json

import numpy as np
import pytest
import torch

import owc_b200  # noqa: F401
from oracle import encoder_ref
from owc_b200 import arch as A
from owc_b200 import metrics, synth
from owc_b200 import text as T
from owc_b200.tokenizer import HostTokenizer, pack_ids


class OracleEngine:
    """Stands in for owc_b200.engine.Engine in CPU tests: same methods, oracle arithmetic."""

    def __init__(self, bundle):
        self.b = bundle
        self.arch = bundle.arch

    def encode(self, ids, cu, out="host", precision=None):
        n = len(cu) - 1
        lens = np.diff(cu)
        t = max(int(lens.max()) if n else 1, 1)
        pad = self.b.arch.pad_id
        ids_p = np.full((n, t), pad, dtype=np.int64)
        mask = np.zeros((n, t), dtype=np.int64)
        for i in range(n):
            ids_p[i, :lens[i]] = ids[cu[i]:cu[i + 1]]
            mask[i, :lens[i]] = 1
        z = encoder_ref.encode_padded(self.b.sd, self.b.arch, torch.from_numpy(ids_p), torch.from_numpy(mask))
        return z.numpy()

    def pair_dot(self, a, b, a_index=None, b_index=None):
        a = torch.from_numpy(np.asarray(a))
        b = torch.from_numpy(np.asarray(b))
        if a_index is not None:
            a = a[torch.from_numpy(np.asarray(a_index)).long()]
        if b_index is not None:
            b = b[torch.from_numpy(np.asarray(b_index)).long()]
        return encoder_ref.paired_dot(a, b).numpy()

    def close(self):
        pass


@pytest.fixture()
def oracle_backed(minilm, monkeypatch):
    monkeypatch.setattr(T, "sentence_bert_model", OracleEngine(minilm))
    monkeypatch.setattr(T, "sentence_bert_processor", HostTokenizer(minilm.vocab, "bert"))
    yield minilm


def test_vocab_is_deterministic_and_real_sized():
    v = synth.build_vocab("bert")
    assert len(v) == 30522 and v[0] == "[PAD]" and v[100] == "[UNK]" and v[101] == "[CLS]" and v[102] == "[SEP]"
    assert v == synth.build_vocab("bert")
    m = synth.build_vocab("mpnet", 30527)
    assert m[:4] == ["<s>", "<pad>", "</s>", "<unk>"] and m[-1] == "<mask>"


def test_real_class_lists_present():
    assert len(synth.real_class_names("caltech101")) == 100
    assert len(synth.real_class_names("flowers102")) == 102
    assert len(synth.real_class_names("food101")) == 101
    assert "apple pie" in synth.real_class_names("food101")


def test_predictions_distinct_and_seeded():
    words = synth.vocab_words(synth.build_vocab("bert"))
    cls = synth.make_class_names(50, words)
    p1, t1 = synth.make_predictions(2000, cls, words)
    p2, t2 = synth.make_predictions(2000, cls, words)
    assert p1 == p2 and t1 == t2 and len(set(p1)) == 2000 and len(set(cls)) == 50


@pytest.mark.parametrize("name,kind", [("minilm", "bert"), ("mpnet", "mpnet")])
def test_host_tokenizer_matches_hf(name, kind, request):
    b = request.getfixturevalue(name)
    tok = HostTokenizer(b.vocab, kind)
    texts = [str(s) for s in b.gold["preds"]] + [str(s) for s in b.gold["class_names"]] + ["", "  ", "Ünï cödé ÀB", "x" * 150]
    ours = tok.encode_ragged(texts)
    theirs = b.hf_tok(texts, padding=False, truncation=True)["input_ids"]
    assert ours == theirs
    long = " ".join(["ab"] * 700)
    assert len(tok.encode_ragged([long])[0]) == 512 == len(b.hf_tok([long], truncation=True)["input_ids"][0])


def test_pack_ids_layout():
    ids, cu = pack_ids([[1, 2, 3], [], [4]])
    assert ids.dtype == np.int32 and cu.dtype == np.int32
    assert ids.tolist() == [1, 2, 3, 4] and cu.tolist() == [0, 3, 3, 4]
    ids, cu = pack_ids([])
    assert ids.size == 0 and cu.tolist() == [0]


@pytest.mark.parametrize("name", ["minilm", "mpnet"])
def test_weight_blob_layout(name, request):
    b = request.getfixturevalue(name)
    blob = A.pack_weights(b.arch, b.sd)
    assert blob.dtype == np.float32 and blob.size == A.blob_num_floats(b.arch)
    h = b.arch.hidden
    w = b.sd["embeddings.word_embeddings.weight"].numpy()
    assert np.array_equal(blob[:w.size].reshape(w.shape), w)
    last = b.sd[f"encoder.layer.{b.arch.layers - 1}.output.LayerNorm.bias"].numpy()
    assert np.array_equal(blob[-h:], last)
    bad = dict(b.sd)
    bad["embeddings.LayerNorm.weight"] = torch.zeros(h + 1)
    with pytest.raises(ValueError):
        A.pack_weights(b.arch, bad)


def test_mpnet_lut_matches_table(mpnet):
    b = mpnet
    blob = A.pack_weights(b.arch, b.sd)
    h, p = b.arch.hidden, b.arch.max_pos
    off = b.arch.vocab * h + p * h + 2 * h
    lut = blob[off:off + b.arch.heads * (2 * p - 1)].reshape(b.arch.heads, 2 * p - 1)
    table = b.sd["encoder.relative_attention_bias.weight"]
    rel = torch.arange(-(p - 1), p)
    want = table[encoder_ref.mpnet_bucket(rel)].T.numpy()
    assert np.array_equal(lut, want)


def test_registry_surface():
    for name in ("semantic_similarity", "mean_average_semantic_similarity", "concept_semantic_similarity"):
        info = metrics.get_metric_info(name)
        assert info.name == name and info.higher_is_better is True and info.output_types == ["generate_until"]
        assert info.group_fn is metrics.get_aggregation_builder(name)
        items = [("a", ["b"])]
        assert info.builder_fn(items) is items          # pass-through (_instance.py:445-462)
    with pytest.raises(KeyError):
        metrics.get_metric_info("nope")


def test_reduce_validation_raises_before_any_work():
    for fn in ("semantic_similarity", "mean_average_semantic_similarity", "concept_semantic_similarity"):
        with pytest.raises(ValueError):
            metrics.get_aggregation_builder(fn)([("a", ["b"])], reduce="sum")


def test_operator_contract(oracle_backed):
    b = oracle_backed
    batch = {"text": ["accordion", "a photo of a cat"]}
    out = T.encode_sentence_bert(batch, input_column="text")
    assert out is batch and "text_sentence_bert_embeds" in batch
    e = batch["text_sentence_bert_embeds"]
    assert isinstance(e, list) and isinstance(e[0], list) and isinstance(e[0][0], float) and len(e[0]) == 384
    with pytest.raises(ValueError):
        T.encode_sentence_bert({"other": ["x"]}, input_column="text")
    with pytest.raises(NotImplementedError):
        T.encode_sentence_bert({"text": "not a list"}, input_column="text")
    out = T.encode_sentence_bert({"p": ["x"]}, input_column="p", output_column="z")
    assert "z" in out
    gold = b.gold["class_embeds"][3]
    assert np.abs(np.asarray(e[0]) - gold).max() < 2e-6        # "accordion" is Caltech101 class 3


def test_metric_bodies_match_reference_goldens(oracle_backed):
    b = oracle_backed
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]]
    items = [(cls[t], [p]) for t, p in zip(b.gold["targets"], preds)]
    ss = metrics.get_metric_info("semantic_similarity")
    got = ss.group_fn(ss.builder_fn(items), reduce="none")
    assert isinstance(got, list) and isinstance(got[0], float)
    assert np.abs(np.asarray(got, np.float32) - b.gold["paired_scores"]).max() < 2e-6
    assert abs(ss.group_fn(items) - float(b.gold["paired_mean"])) < 1e-6
    # online item shape [[gold], [resp...]] and multi-response "last wins" (_group.py:518-519)
    online = [[[cls[t]], ["ignored", p]] for t, p in zip(b.gold["targets"][:16], preds[:16])]
    got2 = ss.group_fn(online, reduce="none")
    assert np.abs(np.asarray(got2, np.float32) - b.gold["paired_scores"][:16]).max() < 2e-6

    mass = metrics.get_aggregation_builder("mean_average_semantic_similarity")
    none = mass(items, reduce="none")
    assert list(none) == [str(k) for k in b.gold["mass_keys"]]
    arr = np.asarray([none[k] for k in none], np.float32)
    assert np.array_equal(arr[:5], b.gold["mass_none"][:5]) and np.abs(arr[5] - b.gold["mass_none"][5]).max() < 1e-6
    assert all(isinstance(v, int) for v in none["semantic_similarity@0.5"][:4])
    mean = mass(items, reduce="mean")
    assert np.abs(np.asarray(list(mean.values())) - b.gold["mass_mean"]).max() < 1e-6


def test_concept_metric_matches_reference_golden(oracle_backed, monkeypatch):
    from oracle import metrics_ref

    b = oracle_backed
    monkeypatch.setattr(T, "concept_extractor", metrics_ref.simple_concepts)
    cls = [str(s) for s in b.gold["class_names"]]
    cpreds = [str(s) for s in b.gold["concept_preds"]]
    items = [(cls[t], [p]) for t, p in zip(b.gold["targets"][:len(cpreds)], cpreds)]
    fn = metrics.get_aggregation_builder("concept_semantic_similarity")
    rows = fn(items, reduce="none")
    gold = json.loads(str(b.gold["concept_json"]))
    for (c, s), (gc, gs) in zip(rows, gold, strict=True):
        assert list(c) == gc and np.abs(np.asarray(s) - np.asarray(gs)).max() < 2e-6
    for red in ("max", "mean", "median", "min"):
        assert abs(fn(items, reduce=red) - float(b.gold[f"concept_{red}"])) < 1e-6


def test_encode_unique_dedups(oracle_backed):
    z, inv = T.encode_unique(["a b", "c", "a b", "c", "d"])
    assert z.shape == (3, 384) and inv.tolist() == [0, 1, 0, 1, 2]


def test_eval_metrics_cli_roundtrip(oracle_backed, tmp_path, capsys):
    """a1: file discovery, item building, JSONL write-back columns and stdout block of eval_metrics.py."""
    import pandas as pd

    from owc_b200 import eval_metrics as EM

    b = oracle_backed
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]][:40]
    tg = b.gold["targets"][:40]
    run = tmp_path / "logs" / "caltech101" / "model-a"
    run.mkdir(parents=True)
    rows = [{"doc_id": i, "target": cls[t], "filtered_resps": [p], "resps": [[p]]} for i, (t, p) in enumerate(zip(tg, preds))]
    f = run / "run_samples_caltech101.jsonl"
    pd.DataFrame(rows).to_json(f, lines=True, orient="records")
    # a multi-round file: [[r1, r2]] per row, last round scored
    run2 = tmp_path / "logs" / "caltech101" / "model-b"
    run2.mkdir(parents=True)
    rows2 = [{"doc_id": i, "target": cls[t], "filtered_resps": [["first round", p]]} for i, (t, p) in enumerate(zip(tg, preds))]
    pd.DataFrame(rows2).to_json(run2 / "x_samples_y.jsonl", lines=True, orient="records")

    args = EM.build_parser().parse_args(["-i", str(tmp_path / "logs"), "-m", "semantic_similarity,mean_average_semantic_similarity"])
    out = EM.main(args)
    assert set(out) == {"caltech101"} and set(out["caltech101"]) == {"model-a", "model-b"}
    gold = b.gold["paired_scores"][:40]
    assert abs(out["caltech101"]["model-a"]["semantic_similarity"] - float(gold.mean())) < 1e-6
    assert abs(out["caltech101"]["model-b"]["semantic_similarity"] - float(gold.mean())) < 1e-6
    df = pd.read_json(f, lines=True)
    assert np.abs(df["semantic_similarity"].to_numpy(np.float32) - gold).max() < 2e-6
    for col in ("mean_average_semantic_similarity", "semantic_similarity@0.5", "semantic_similarity@0.9"):
        assert col in df.columns
    assert df["semantic_similarity@0.9"].tolist() == (gold >= np.float32(0.9)).astype(int).tolist()
    text = capsys.readouterr().out
    assert "Semantic similarity on caltech101:" in text
    assert f"{'model-a':<29}: {float(gold.mean()):.3f}" in text
    with pytest.raises(KeyError):
        EM.score_file(str(f), ["bleu"])


def _games(bundle, n, seed=3):
    import random

    rng = random.Random(seed)
    cls = [str(s) for s in bundle.gold["class_names"]]
    preds = [str(s) for s in bundle.gold["preds"]]
    players = ["model_a", "model_b", "model_c"]
    out = []
    for _ in range(n):
        ref = rng.choice(cls)
        a, b = rng.sample(players, 2)
        out.append({"reference": ref, "player_a_name": a, "player_b_name": b,
                    "player_a_response": rng.choice([ref, rng.choice(preds), ref + " " + rng.choice(bundle.words)]),
                    "player_b_response": rng.choice([ref, rng.choice(preds)])})
    return out


def test_ranking_criterion_matches_oracle(oracle_backed):
    """eval_ranking.py:192-237 (semantic_similarity criterion) through the host mirror vs the oracle restatement."""
    from oracle import metrics_ref
    from owc_b200.ranking import semantic_similarity_game_scores

    b = oracle_backed
    games = _games(b, 300)
    got, sa, sb = semantic_similarity_game_scores(games, return_similarities=True)
    want, wa, wb = metrics_ref.ranking_game_scores(
        lambda texts: encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, list(texts)), games)
    assert np.abs(sa - wa.numpy()).max() < 1e-5 and np.abs(sb - wb.numpy()).max() < 1e-5
    far = np.abs(np.abs(wa.numpy() - wb.numpy()) - 0.05) > 1e-4       # decisions away from the threshold are identical
    assert np.array_equal(np.asarray(got)[far], np.asarray(want)[far])
    assert set(got) <= {0.0, 0.5, 1.0} and {0.5} <= set(got)            # identical responses draw
    assert semantic_similarity_game_scores([]) == []
    with pytest.raises(ValueError):
        semantic_similarity_game_scores([{"reference": "x", "player_a_response": "y"}])


def test_sweep_groups_files_by_class_set(oracle_backed):
    """score_sweep's grouping / cutting logic on CPU (oracle-backed engine: encode + a torch top-k stand-in)."""
    from owc_b200.sweep import score_sweep

    b = oracle_backed

    class Eng(OracleEngine):
        def encode(self, ids, cu, out="host", precision=None):
            z = super().encode(ids, cu)
            return torch.from_numpy(z) if out == "device" else z

        def score(self, ids, cu, zc, targets=None, k=1, precision=None, out="host", want_embeds=False):
            zp = torch.from_numpy(OracleEngine.encode(self, ids, cu))
            s, i, _ = encoder_ref.sim_topk(zp, zc, k)
            paired = None if targets is None else (zp * zc[torch.from_numpy(np.asarray(targets)).long()]).sum(1).numpy()
            return {"scores": s.numpy(), "idx": i.numpy().astype(np.int32), "paired": paired, "embeds": None}

    eng = Eng(b)
    tok = HostTokenizer(b.vocab, "bert")
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]]
    tg = b.gold["targets"].tolist()
    tasks = {"t1": {"class_names": cls, "files": {"m0": {"predictions": preds[:40], "targets": tg[:40]},
                                                   "m1": {"predictions": preds[40:100], "targets": tg[40:100]}}},
             "t2": {"class_names": cls[:10], "files": {"m0": {"predictions": preds[:25], "targets": None}}}}
    got = score_sweep(eng, tok, tasks, k=3)
    assert got["t1"]["m0"]["idx"].shape == (40, 3) and got["t1"]["m1"]["idx"].shape == (60, 3)
    assert got["t2"]["m0"]["paired"] is None and got["t2"]["m0"]["scores"].shape == (25, 3)
    assert np.abs(got["t1"]["m1"]["paired"] - b.gold["paired_scores"][40:100]).max() < 1e-5   # the reference's own scores
    with pytest.raises(ValueError):
        score_sweep(eng, tok, {"t": {"class_names": cls[:2], "files": {"m": {"predictions": ["x"], "targets": None}}}}, k=3)


def test_eval_metrics_parallel_io_is_byte_identical(oracle_backed, tmp_path):
    """N2: `--jobs` parses / rewrites files in worker processes (scoring stays here); files equal the serial path's
    byte for byte, and `--skip-scored` leaves scored files alone."""
    import shutil

    import pandas as pd

    from owc_b200 import eval_metrics as EM

    b = oracle_backed
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]]
    tg = b.gold["targets"]
    root = tmp_path / "logs"
    for m in range(3):
        run = root / "caltech101" / f"model-{m}"
        run.mkdir(parents=True)
        rows = [{"doc_id": i, "target": cls[tg[i]], "filtered_resps": [preds[i]], "note": "a/b é"} for i in range(20 * m, 20 * m + 30)]
        pd.DataFrame(rows).to_json(run / "r_samples_caltech101.jsonl", lines=True, orient="records")
    twin = tmp_path / "twin"
    shutil.copytree(root, twin / "logs")
    metrics = "semantic_similarity,mean_average_semantic_similarity"
    serial = EM.main(EM.build_parser().parse_args(["-i", str(root), "-m", metrics]))
    par = EM.main(EM.build_parser().parse_args(["-i", str(twin / "logs"), "-m", metrics, "--jobs", "2"]))
    assert serial == par
    for f in sorted(root.rglob("*.jsonl")):
        g = twin / "logs" / f.relative_to(root)
        assert f.read_bytes() == g.read_bytes()
    before = {f: f.stat().st_mtime_ns for f in root.rglob("*.jsonl")}
    again = EM.main(EM.build_parser().parse_args(["-i", str(root), "-m", metrics, "--skip-scored"]))
    assert before == {f: f.stat().st_mtime_ns for f in root.rglob("*.jsonl")}            # scored files are left alone ...
    # ... but stay in the report: their outputs are recomputed from the stored columns (10 decimals on disk)
    assert again.keys() == serial.keys()
    for task in serial:
        assert again[task].keys() == serial[task].keys()
        for model, out in serial[task].items():
            assert again[task][model].keys() == out.keys() and again[task][model]["_num_samples"] == out["_num_samples"]
            assert all(abs(again[task][model][k] - v) < 1e-9 for k, v in out.items())


def test_jobs_pipeline_keeps_rows_in_order_across_jobs():
    """pipeline.score_jobs_pipelined: chunks of several jobs run through one worker thread; every job gets its own rows
    back in order (ragged chunking, an empty job, a job given as per-file row lists, targets on some jobs only)."""
    from owc_b200.pipeline import RowParts, score_jobs_pipelined, score_strings_pipelined

    class Tok:
        def encode_packed(self, strings):
            ids = np.asarray([int(s) for s in strings], np.int32)      # one "token" per row: the row's own number
            return ids, np.arange(len(strings) + 1, dtype=np.int32)

    class Eng:
        def __init__(self):
            self.calls = []

        def score(self, ids, cu, zc, targets=None, k=1, precision=None, out="host"):
            self.calls.append(len(ids))
            s = np.repeat(ids[:, None].astype(np.float32), k, 1) + zc          # zc: the job's marker
            return {"scores": s, "idx": np.repeat(ids[:, None], k, 1).astype(np.int32),
                    "paired": None if targets is None else ids.astype(np.float32) * 2 + np.asarray(targets)}

    a = [str(i) for i in range(53)]
    b = RowParts([[str(i) for i in range(100, 117)], [], [str(i) for i in range(200, 230)]])
    jobs = [(a, 1000.0, None), ([], 2000.0, None), (b, 3000.0, np.arange(47, dtype=np.int32))]
    eng = Eng()
    ra, re_, rb = score_jobs_pipelined(eng, Tok(), jobs, k=2, chunk_rows=20, first_chunk_rows=5)
    assert eng.calls == [5, 20, 20, 8, 20, 20, 7]
    assert np.array_equal(ra["idx"][:, 0], np.arange(53)) and np.array_equal(ra["scores"][:, 1], np.arange(53) + 1000.0) and ra["paired"] is None
    assert re_["scores"].shape == (0, 2)
    want_b = np.concatenate([np.arange(100, 117), np.arange(200, 230)])
    assert np.array_equal(rb["idx"][:, 1], want_b) and np.array_equal(rb["paired"], want_b * 2.0 + np.arange(47))
    one = score_strings_pipelined(Eng(), Tok(), a, 7.0, k=1, chunk_rows=16)
    assert np.array_equal(one["scores"][:, 0], np.arange(53) + 7.0)
    with pytest.raises(ValueError):
        score_jobs_pipelined(Eng(), Tok(), [(a, 0.0, np.zeros(3, np.int32))])
    with pytest.raises(NotImplementedError):
        score_strings_pipelined(Eng(), Tok(), "not a list", 0.0)
