"""Pin the oracle: CPU restatement vs outputs of the reference itself (tests/golden) and vs HF."""

import json

import numpy as np
import pytest
import torch

from oracle import encoder_ref, metrics_ref


def _encode_fn(b):
    return lambda texts: encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, list(texts))


@pytest.mark.parametrize("name", ["minilm", "mpnet"])
def test_oracle_matches_hf_forward(name, request):
    b = request.getfixturevalue(name)
    ids = torch.from_numpy(b.gold["first_ids"]).long()
    mask = torch.from_numpy(b.gold["first_mask"]).long()
    hs: list = []
    with torch.no_grad():
        out = encoder_ref.encoder_forward(b.sd, b.arch, ids, mask, hidden_states=hs)
        hf = b.model(input_ids=ids, attention_mask=mask, output_hidden_states=True)
    m = mask.bool()
    for mine, theirs in zip(hs, hf.hidden_states, strict=True):
        assert (mine - theirs)[m].abs().max() < 5e-5      # padded positions are don't-care
    assert (out - hf.last_hidden_state)[m].abs().max() < 5e-5
    # anchors recorded at golden time (reference process) agree with this process
    assert np.abs(hs[0][:4].numpy() - b.gold["first_hidden_emb"])[m[:4].numpy()].max() < 5e-5
    assert np.abs(hs[1][:4].numpy() - b.gold["first_hidden_l1"])[m[:4].numpy()].max() < 5e-5
    assert np.abs(out[:4].numpy() - b.gold["first_hidden_last"])[m[:4].numpy()].max() < 5e-5


@pytest.mark.parametrize("name", ["minilm", "mpnet"])
def test_oracle_matches_reference_embeddings(name, request):
    """encode_sentence_bert of the reference (golden) == oracle.encode_strings, to fp32 rounding."""
    b = request.getfixturevalue(name)
    zc = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, [str(s) for s in b.gold["class_names"]])
    zp = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, [str(s) for s in b.gold["preds"]])
    assert np.abs(zc.numpy() - b.gold["class_embeds"]).max() < 2e-6
    assert np.abs(zp.numpy() - b.gold["pred_embeds"]).max() < 2e-6
    assert np.abs(np.linalg.norm(zp.numpy(), axis=1) - 1).max() < 1e-5


@pytest.mark.parametrize("name", ["minilm", "mpnet"])
def test_oracle_matches_reference_metrics(name, request):
    b = request.getfixturevalue(name)
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]]
    items = [(cls[t], [p]) for t, p in zip(b.gold["targets"], preds)]
    enc = _encode_fn(b)
    s = metrics_ref.semantic_similarity(items, enc, reduce="none")
    assert np.abs(np.asarray(s, dtype=np.float32) - b.gold["paired_scores"]).max() < 2e-6
    assert abs(metrics_ref.semantic_similarity(items, enc, reduce="mean") - float(b.gold["paired_mean"])) < 1e-6
    mass = metrics_ref.mean_average_semantic_similarity(items, enc, reduce="none")
    assert list(mass.keys()) == [str(k) for k in b.gold["mass_keys"]]
    got = np.asarray([mass[k] for k in mass], dtype=np.float32)
    # a threshold flip needs a paired score within 2e-6 of 0.5/0.6/...: none in the fixture
    assert np.array_equal(got[:5], b.gold["mass_none"][:5])
    assert np.abs(got[5] - b.gold["mass_none"][5]).max() < 1e-6
    mass_mean = metrics_ref.mean_average_semantic_similarity(items, enc, reduce="mean")
    assert np.abs(np.asarray(list(mass_mean.values())) - b.gold["mass_mean"]).max() < 1e-6


def test_oracle_diag_of_matrix_is_paired_score(minilm):
    """S[i, target_i] must reproduce the reference's paired score (the link back)."""
    b = minilm
    zc = torch.from_numpy(b.gold["class_embeds"])
    zp = torch.from_numpy(b.gold["pred_embeds"])
    s = zp @ zc.T
    t = torch.from_numpy(b.gold["targets"]).long()
    diag = s[torch.arange(len(t)), t]
    assert (diag - torch.from_numpy(b.gold["paired_scores"])).abs().max() < 1e-6
    scores, idx, margin = encoder_ref.sim_topk(zp, zc, 5)
    assert scores.shape == (len(t), 5) and (scores[:, :-1] >= scores[:, 1:]).all()
    assert (s.gather(1, idx) == scores).all()
    assert torch.equal(idx[:, 0], s.argmax(1)) or (margin[idx[:, 0] != s.argmax(1)] == 0).all()


def test_oracle_concepts_match_reference(minilm):
    b = minilm
    cls = [str(s) for s in b.gold["class_names"]]
    cpreds = [str(s) for s in b.gold["concept_preds"]]
    items = [(cls[t], [p]) for t, p in zip(b.gold["targets"][:len(cpreds)], cpreds)]
    enc = _encode_fn(b)
    rows = metrics_ref.concept_semantic_similarity(items, enc, reduce="none")
    gold = json.loads(str(b.gold["concept_json"]))
    assert len(rows) == len(gold)
    for (c, s), (gc, gs) in zip(rows, gold):
        assert list(c) == gc
        assert np.abs(np.asarray(s) - np.asarray(gs)).max() < 2e-6
    for red in ("max", "mean", "median", "min"):
        v = metrics_ref.concept_semantic_similarity(items, enc, reduce=red)
        assert abs(v - float(b.gold[f"concept_{red}"])) < 1e-6


def test_oracle_reduce_validation(minilm):
    with pytest.raises(ValueError):
        metrics_ref.semantic_similarity([("a", ["b"])], _encode_fn(minilm), reduce="sum")


def test_mpnet_bucket_table_matches_reference_expression():
    from owc_b200.arch import mpnet_relative_bucket

    rel = torch.arange(-513, 514)
    want = encoder_ref.mpnet_bucket(rel)
    got = torch.tensor([mpnet_relative_bucket(int(r)) for r in rel])
    assert torch.equal(want, got)


def test_oracle_matches_reference_c3_golden(mpnet):
    """BASELINE config 3 in miniature (mpnet-base arch x 1000 class names), produced by the reference itself
    (oracle/make_golden.py --only-c3): paired scores of its semantic_similarity and the top-5 of S on its embeddings.
    A 256-row slice keeps the CPU suite short; the whole golden is held against the CUDA path in tests/test_gpu_configs.py."""
    import os

    from owc_b200 import synth

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpnet_c3.npz"), allow_pickle=False))
    b = mpnet
    assert np.abs(g["weight_checksums"] - b.gold["weight_checksums"]).max() < 1e-6       # same seeded weights as the other fixture
    cls = synth.make_class_names(int(g["n_classes"]), b.words, seed=int(g["seed"]))
    preds, targets = synth.make_predictions(int(g["n_preds"]), cls, b.words, seed=int(g["seed"]))
    assert [str(x) for x in g["pred_head"]] == preds[:8] and np.array_equal(np.asarray(targets, np.int32), g["targets"])
    n = 256
    zc = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, cls)
    zp = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, preds[:n])
    assert np.abs(zc[:32].numpy() - g["class_embeds_head"]).max() < 2e-6
    assert np.abs(zp[:64].numpy() - g["pred_embeds_head"]).max() < 2e-6
    scores, idx, margin = encoder_ref.sim_topk(zp, zc, 5)
    assert np.abs(scores.numpy() - g["top5_scores"][:n]).max() < 2e-6
    clear = g["margin"][:n] >= 1e-5
    assert np.array_equal(idx.numpy()[clear], g["top5_idx"][:n][clear].astype(np.int64))
    paired = encoder_ref.paired_dot(zc[torch.from_numpy(g["targets"][:n]).long()], zp)
    assert np.abs(paired.numpy() - g["paired_scores"][:n]).max() < 2e-6
