"""GPU parity of the whole path (encode -> pool -> normalise -> match) against the reference's golden
vectors and the oracle, through the C-ABI and through the plugin API (pytest -m gpu)."""

import json

import numpy as np
import pytest
import torch

import owc_b200  # noqa: F401
from oracle import encoder_ref, metrics_ref
from owc_b200 import metrics, synth
from owc_b200 import text as T
from owc_b200._cabi import OwcError
from owc_b200.engine import Engine
from owc_b200.tokenizer import HostTokenizer

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}   # north star: |dcos| <= 1e-4 in check mode, 1e-2 in bf16


@pytest.fixture(scope="module")
def eng_minilm(minilm):
    e = Engine(minilm.arch, minilm.sd, device=0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_mpnet(mpnet):
    e = Engine(mpnet.arch, mpnet.sd, device=0)
    yield e
    e.close()


def _tok(b):
    return HostTokenizer(b.vocab, "bert" if b.arch.kind == 0 else "mpnet")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["minilm", "mpnet"])
def test_embeddings_match_reference_golden(name, precision, request):
    b = request.getfixturevalue(name)
    eng = request.getfixturevalue("eng_" + name)
    tok = _tok(b)
    for col, gold in (("class_names", "class_embeds"), ("preds", "pred_embeds")):
        ids, cu = tok.encode_packed([str(s) for s in b.gold[col]])
        z = eng.encode(ids, cu, precision=precision)
        err = np.abs(z - b.gold[gold]).max()
        # per-element error of a unit vector bounds the cosine error; stricter than required
        assert err < TOL[precision] / (10 if precision == "bf16" else 1), f"{name}/{col}/{precision}: {err}"
        assert np.abs(np.linalg.norm(z, axis=1) - 1).max() < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_score_matches_oracle_and_paired_golden(minilm, eng_minilm, precision):
    b, eng = minilm, eng_minilm
    tok = _tok(b)
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]]
    cids, ccu = tok.encode_packed(cls)
    zc = eng.encode(cids, ccu, out="device", precision=precision)
    pids, pcu = tok.encode_packed(preds)
    res = eng.score(pids, pcu, zc, targets=b.gold["targets"], k=5, precision=precision, want_embeds=True)
    zp_ref, zc_ref = torch.from_numpy(b.gold["pred_embeds"]), torch.from_numpy(b.gold["class_embeds"])
    s_ref, i_ref, margin = encoder_ref.sim_topk(zp_ref, zc_ref, 5)
    assert np.abs(res["paired"] - b.gold["paired_scores"]).max() < TOL[precision]
    assert np.abs(res["scores"] - s_ref.numpy()).max() < TOL[precision]
    clear = margin.numpy() >= 1e-3
    agree = res["idx"][:, 0] == i_ref[:, 0].numpy()
    assert agree[clear].all(), f"argmax differs off near-ties at rows {np.nonzero(~agree & clear)[0][:10]}"
    assert agree.mean() >= 0.90
    # device-resident variant returns the same numbers
    dev = eng.score(torch.from_numpy(pids).cuda(), pcu, zc, targets=torch.from_numpy(b.gold["targets"]).cuda(), k=5,
                    precision=precision, out="device")
    torch.cuda.synchronize()
    assert np.array_equal(dev["idx"].cpu().numpy(), res["idx"])
    assert np.abs(dev["paired"].cpu().numpy() - res["paired"]).max() == 0


def test_chunking_is_invisible(minilm):
    b = minilm
    tok = _tok(b)
    words = b.words
    preds, _ = synth.make_predictions(3000, [str(s) for s in b.gold["class_names"]], words, seed=7)
    ids, cu = tok.encode_packed(preds)
    big = Engine(b.arch, b.sd, device=0)
    small = Engine(b.arch, b.sd, device=0, chunk_tokens=1024)
    try:
        for prec in ("bf16", "fp32"):
            z1 = big.encode(ids, cu, precision=prec)
            z2 = small.encode(ids, cu, precision=prec)
            assert np.abs(z1 - z2).max() < 1e-6
        st = small.stats()
        assert st["tokens_encoded"] == 2 * int(cu[-1]) and st["sequences_encoded"] == 2 * len(preds)
    finally:
        big.close()
        small.close()


@pytest.mark.parametrize("name", ["minilm", "mpnet"])
def test_long_ragged_and_edge_sequences(name, request):
    b = request.getfixturevalue(name)
    eng = request.getfixturevalue("eng_" + name)
    tok = _tok(b)
    w = b.words
    texts = ["", "a", " ".join(w[:30]), " ".join(w[:62]), " ".join(w[:63]), " ".join(w[:100]), " ".join(w[:300]),
             " ".join(w[:700]), w[5]]
    ids, cu = tok.encode_packed(texts)
    assert np.diff(cu).max() == 512 and np.diff(cu).min() == 2
    want = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, texts).numpy()
    for prec in ("fp32", "bf16"):
        z = eng.encode(ids, cu, precision=prec)
        assert np.abs(z - want).max() < TOL[prec] / (10 if prec == "bf16" else 1), prec
    # padding-invariance: the same string alone or next to a 512-token neighbour
    ids1, cu1 = tok.encode_packed([texts[2]])
    z1 = eng.encode(ids1, cu1, precision="fp32")
    z_all = eng.encode(ids, cu, precision="fp32")
    assert np.abs(z1[0] - z_all[2]).max() < 1e-6


def test_empty_and_invalid_inputs(minilm, eng_minilm):
    eng = eng_minilm
    assert eng.encode(np.zeros(0, np.int32), np.zeros(1, np.int32)).shape == (0, 384)
    with pytest.raises(OwcError):
        eng.encode(np.zeros(600, np.int32), np.array([0, 600], np.int32))          # longer than the position table
    with pytest.raises(OwcError):
        eng.encode(np.array([101, 40000, 102], np.int32), np.array([0, 3], np.int32))  # id outside the vocab
    z = eng.encode(np.array([101, 102], np.int32), np.array([0, 2], np.int32))          # engine still usable
    assert np.isfinite(z).all()
    with pytest.raises(OwcError):
        eng.encode(np.array([101, 102], np.int32), np.array([1, 2], np.int32))          # cu[0] != 0
    zn = eng.encode(np.array([101, 102], np.int32), np.array([0, 0, 2], np.int32))      # empty sequence -> NaN row
    assert np.isnan(zn[0]).all() and np.isfinite(zn[1]).all()


def test_plugin_api_end_to_end(minilm, monkeypatch):
    """The reference-facing calls (eval_metrics.py:81-90) on the GPU vs the reference's own outputs."""
    b = minilm
    T.configure(b.arch, b.sd, b.vocab, device=0)
    try:
        cls = [str(s) for s in b.gold["class_names"]]
        preds = [str(s) for s in b.gold["preds"]]
        items = [(cls[t], [p]) for t, p in zip(b.gold["targets"], preds)]
        info = metrics.get_metric_info("semantic_similarity")
        got = np.asarray(info.group_fn(info.builder_fn(items), reduce="none"), np.float32)
        assert np.abs(got - b.gold["paired_scores"]).max() < 1e-2
        assert abs(info.group_fn(items) - float(b.gold["paired_mean"])) < 1e-3
        mass = metrics.get_aggregation_builder("mean_average_semantic_similarity")(items, reduce="none")
        gold = b.gold["mass_none"]
        far = np.min(np.abs(b.gold["paired_scores"][None, :] - np.array([.5, .6, .7, .8, .9])[:, None]), axis=0) > 1e-2
        got5 = np.asarray([mass[k] for k in list(mass)[:5]], np.float32)
        assert np.array_equal(got5[:, far], gold[:5][:, far])
        batch = T.encode_sentence_bert({"text": cls[:8]}, input_column="text")
        assert np.abs(np.asarray(batch["text_sentence_bert_embeds"]) - b.gold["class_embeds"][:8]).max() < 1e-3
        monkeypatch.setattr(T, "concept_extractor", metrics_ref.simple_concepts)
        cpreds = [str(s) for s in b.gold["concept_preds"]]
        citems = [(cls[t], [p]) for t, p in zip(b.gold["targets"][:len(cpreds)], cpreds)]
        rows = metrics.get_aggregation_builder("concept_semantic_similarity")(citems, reduce="none")
        for (c, s), (gc, gs) in zip(rows, json.loads(str(b.gold["concept_json"])), strict=True):
            assert list(c) == gc and np.abs(np.asarray(s) - np.asarray(gs)).max() < 1e-2
        m = metrics.semantic_matching(preds, cls, k=5, targets=b.gold["targets"].tolist())
        assert np.abs(m["paired"] - b.gold["paired_scores"]).max() < 1e-2
        assert m["indices"].shape == (len(preds), 5)
    finally:
        T.reset()


def test_full_size_properties_c2(minilm, eng_minilm):
    """BASELINE config 2 shape (100k predictions x Food101's 101 names): size-independent checks."""
    b, eng = minilm, eng_minilm
    tok = _tok(b)
    cls = synth.real_class_names("food101")
    preds, targets = synth.make_predictions(100_000, cls, b.words, seed=1234)
    cids, ccu = tok.encode_packed(cls)
    zc = eng.encode(cids, ccu, out="device")
    pids, pcu = tok.encode_packed(preds)
    t = np.asarray(targets, np.int32)
    res = eng.score(pids, pcu, zc, targets=t, k=5, want_embeds=True)
    z = res["embeds"]
    assert np.abs(np.linalg.norm(z, axis=1) - 1).max() < 1e-5                 # unit rows
    assert (np.diff(res["scores"], axis=1) <= 0).all()                         # sorted
    assert (res["scores"][:, 0] >= res["paired"] - 1e-6).all()                 # top-1 dominates the paired score
    hit = res["idx"][:, 0] == t
    assert np.abs(res["scores"][hit, 0] - res["paired"][hit]).max() < 1e-6     # and equals it when target wins
    assert all(len(set(r)) == 5 for r in res["idx"][:2000].tolist())           # distinct classes per row
    # permutation equivariance: shuffling the rows shuffles the results
    perm = np.random.default_rng(0).permutation(len(preds))[:5000]
    ids2, cu2 = tok.encode_packed([preds[i] for i in perm])
    res2 = eng.score(ids2, cu2, zc, targets=t[perm], k=5)
    assert np.abs(res2["paired"] - res["paired"][perm]).max() < 1e-6
    assert np.array_equal(res2["idx"], res["idx"][perm])
    # a 2k-row sample against the oracle
    sample = perm[:2000]
    zp_ref = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, [preds[i] for i in sample])
    zc_ref = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, cls)
    s_ref, i_ref, margin = encoder_ref.sim_topk(zp_ref, zc_ref, 5)
    assert np.abs(res["scores"][sample] - s_ref.numpy()).max() < 1e-2
    clear = margin.numpy() >= 1e-3
    agree = res["idx"][sample, 0] == i_ref[:, 0].numpy()
    assert agree[clear].all() and agree.mean() >= 0.9


def test_ranking_criterion_on_gpu(minilm):
    """eval_ranking.py:192-237 through the host mirror on the GPU vs the oracle restatement of the same lines."""
    import random

    from owc_b200.ranking import semantic_similarity_game_scores

    b = minilm
    T.configure(b.arch, b.sd, b.vocab, device=0)
    try:
        rng = random.Random(3)
        cls = [str(s) for s in b.gold["class_names"]]
        preds = [str(s) for s in b.gold["preds"]]
        games = [{"reference": (ref := rng.choice(cls)), "player_a_response": rng.choice([ref, rng.choice(preds)]),
                  "player_b_response": rng.choice([ref, rng.choice(preds), ref + " " + rng.choice(b.words)])}
                 for _ in range(2000)]
        got, sa, sb = semantic_similarity_game_scores(games, return_similarities=True)
        want, wa, wb = metrics_ref.ranking_game_scores(
            lambda texts: encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, list(texts)), games)
        assert np.abs(sa - wa.numpy()).max() < 1e-2 and np.abs(sb - wb.numpy()).max() < 1e-2
        far = np.abs(np.abs(wa.numpy() - wb.numpy()) - 0.05) > 2e-3
        assert far.mean() > 0.9 and np.array_equal(np.asarray(got)[far], np.asarray(want)[far])
    finally:
        T.reset()
