"""GPU parity of the block-resident fused encoder kernel (csrc/enc_fused.cuh) through the C-ABI: against the
per-GEMM kernel path of the same library, against the oracle, and on the edge cases of its row-block packing
(pytest -m gpu). The fused kernel serves MiniLM-width BERT encoders in bf16 mode when no sequence exceeds 64
tokens; everything else takes the per-GEMM path, which these tests use as the second opinion."""

import numpy as np
import pytest

import owc_b200  # noqa: F401
from oracle import encoder_ref
from owc_b200 import synth
from owc_b200.engine import Engine
from owc_b200.tokenizer import HostTokenizer

pytestmark = pytest.mark.gpu


def _engine(b, monkeypatch, fused: int, waves: int | None = None):
    monkeypatch.setenv("OWC_FUSED", str(fused))
    if waves is not None:
        monkeypatch.setenv("OWC_FUSED_WAVES", str(waves))
    else:
        monkeypatch.delenv("OWC_FUSED_WAVES", raising=False)
    return Engine(b.arch, b.sd, device=0)


@pytest.fixture()
def engines(minilm, monkeypatch):
    fused = _engine(minilm, monkeypatch, 1)
    plain = _engine(minilm, monkeypatch, 0)
    yield fused, plain
    fused.close()
    plain.close()


def _ragged_texts(words, n, seed, lo=1, hi=24):
    rng = np.random.default_rng(seed)
    return [" ".join(rng.choice(words, size=int(k))) for k in rng.integers(lo, hi, size=n)]


def _launches(eng, fn):
    a = eng.stats()["kernel_launches"]
    out = fn()
    return out, eng.stats()["kernel_launches"] - a


def test_fused_path_is_taken_and_matches_per_gemm_path(minilm, engines):
    fused, plain = engines
    tok = HostTokenizer(minilm.vocab, "bert")
    texts = _ragged_texts(minilm.words, 3000, 1)
    ids, cu = tok.encode_packed(texts)
    assert np.diff(cu).max() <= 64
    zf, nf = _launches(fused, lambda: fused.encode(ids, cu))
    zp, np_ = _launches(plain, lambda: plain.encode(ids, cu))
    # one embed + one fused + one pool launch per chunk, against ~27 launches per chunk on the per-GEMM path
    assert nf < np_ / 3, (nf, np_)
    assert np.abs(zf - zp).max() < 1e-3
    assert np.abs(np.linalg.norm(zf, axis=1) - 1).max() < 1e-5
    want = encoder_ref.encode_strings(minilm.sd, minilm.arch, minilm.hf_tok, texts[:400]).numpy()
    assert np.abs(zf[:400] - want).max() < 1e-3   # north star: 1e-2 on the cosine in bf16 mode


def test_fused_matches_reference_golden(minilm, engines):
    fused, _ = engines
    tok = HostTokenizer(minilm.vocab, "bert")
    for col, gold in (("class_names", "class_embeds"), ("preds", "pred_embeds")):
        ids, cu = tok.encode_packed([str(s) for s in minilm.gold[col]])
        assert np.diff(cu).max() <= 64
        z = fused.encode(ids, cu)
        assert np.abs(z - minilm.gold[gold]).max() < 1e-3


@pytest.mark.parametrize("lens", [
    [1], [64], [64, 64], [64, 64, 1], [63, 2, 63, 2], [3] * 43, [2] * 200, [1] * 300,
    [64] * 5 + [1] * 7 + [33] * 9, [5, 0, 7, 0, 0, 64, 0], [0, 0, 3], [128 // 3] * 3 + [2],
])
def test_row_block_packing_edges(minilm, engines, lens):
    """Sequences exactly filling a 128-row block, blocks closed by the 128-sequence cap, empty sequences, a single
    row, odd block counts (the pair's second CTA has no block)."""
    fused, plain = engines
    rng = np.random.default_rng(len(lens) * 1000 + sum(lens))
    ids = rng.integers(5, minilm.arch.vocab, size=sum(lens)).astype(np.int32)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    zf = fused.encode(ids, cu)
    zp = plain.encode(ids, cu)
    empty = np.asarray(lens) == 0
    assert np.isnan(zf[empty]).all() and np.isnan(zp[empty]).all()       # _text.py:202: no epsilon
    assert np.isfinite(zf[~empty]).all()
    assert np.abs(zf[~empty] - zp[~empty]).max() < 1e-3


def test_position_in_block_and_launch_size_are_invisible(minilm, monkeypatch):
    """A sequence's embedding does not depend on its neighbours, its row inside the block, or how many block
    pairs one launch carries."""
    tok = HostTokenizer(minilm.vocab, "bert")
    texts = _ragged_texts(minilm.words, 5000, 3)
    ids, cu = tok.encode_packed(texts)
    e4 = _engine(minilm, monkeypatch, 1)
    e1 = _engine(minilm, monkeypatch, 1, waves=1)
    try:
        z4 = e4.encode(ids, cu)
        z1, n1 = _launches(e1, lambda: e1.encode(ids, cu))
        assert n1 > 3          # several launches at one wave per launch
        assert np.array_equal(z1, z4)
        perm = np.random.default_rng(0).permutation(len(texts))
        ids2, cu2 = tok.encode_packed([texts[i] for i in perm])
        z2 = e4.encode(ids2, cu2)
        assert np.abs(z2 - z4[perm]).max() < 1e-6
        ids3, cu3 = tok.encode_packed([texts[17]])
        assert np.abs(e4.encode(ids3, cu3)[0] - z4[17]).max() < 1e-6
    finally:
        e4.close()
        e1.close()


def test_longer_sequences_fall_back_to_the_per_gemm_path(minilm, engines):
    fused, plain = engines
    tok = HostTokenizer(minilm.vocab, "bert")
    texts = _ragged_texts(minilm.words, 200, 5) + [" ".join(minilm.words[:90])]
    ids, cu = tok.encode_packed(texts)
    assert np.diff(cu).max() > 64
    zf = fused.encode(ids, cu)
    zp = plain.encode(ids, cu)
    assert np.array_equal(zf, zp)


def test_score_through_fused_encoder(minilm, engines):
    fused, plain = engines
    tok = HostTokenizer(minilm.vocab, "bert")
    cls = synth.real_class_names("food101")
    preds, targets = synth.make_predictions(4000, cls, minilm.words, seed=11)
    cids, ccu = tok.encode_packed(cls)
    pids, pcu = tok.encode_packed(preds)
    t = np.asarray(targets, np.int32)
    out = []
    for e in (fused, plain):
        zc = e.encode(cids, ccu, out="device")
        out.append(e.score(pids, pcu, zc, targets=t, k=5))
    f, p = out
    assert np.abs(f["paired"] - p["paired"]).max() < 1e-3
    assert np.abs(f["scores"] - p["scores"]).max() < 1e-3
    margin = p["scores"][:, 0] - p["scores"][:, 1]
    clear = margin >= 1e-3
    assert (f["idx"][clear, 0] == p["idx"][clear, 0]).all()


def test_pipelined_strings_to_results_equals_one_call(minilm, engines):
    from owc_b200.pipeline import score_strings_pipelined

    fused, _ = engines
    tok = HostTokenizer(minilm.vocab, "bert")
    cls = synth.real_class_names("food101")
    preds, targets = synth.make_predictions(9000, cls, minilm.words, seed=21)
    preds[17] = "café " + preds[17]            # a row the native tokenizer hands to the backend
    cids, ccu = tok.encode_packed(cls)
    zc = fused.encode(cids, ccu, out="device")
    t = np.asarray(targets, np.int32)
    pids, pcu = tok.encode_packed(preds)
    one = fused.score(pids, pcu, zc, targets=t, k=5)
    many = score_strings_pipelined(fused, tok, preds, zc, targets=t, k=5, chunk_rows=2000)
    assert np.array_equal(one["idx"], many["idx"])
    assert np.abs(one["scores"] - many["scores"]).max() < 1e-6 and np.abs(one["paired"] - many["paired"]).max() < 1e-6
    empty = score_strings_pipelined(fused, tok, [], zc, k=5)
    assert empty["scores"].shape == (0, 5) and empty["paired"] is None


def test_config4_scale_properties(minilm, engines):
    """BASELINE config 4 shape: a 117k-name vocabulary encoded by the fused kernel, predictions matched with the fused
    top-5; checked through size-independent properties and against an exact fp32 product of the same embeddings."""
    import torch

    fused, _ = engines
    tok = HostTokenizer(minilm.vocab, "bert")
    names = synth.make_class_names(117_000, minilm.words, seed=1234)
    assert len(set(names)) == len(names)
    preds, targets = synth.make_predictions(6000, names, minilm.words, seed=9)
    cids, ccu = tok.encode_packed(names)
    zc = fused.encode(cids, ccu, out="device")
    assert zc.shape == (117_000, 384) and torch.isfinite(zc).all()
    assert (zc.norm(dim=1) - 1).abs().max().item() < 1e-5
    pids, pcu = tok.encode_packed(preds)
    t = np.asarray(targets, np.int32)
    res = fused.score(pids, pcu, zc, targets=t, k=5, want_embeds=True)
    assert (np.diff(res["scores"], axis=1) <= 0).all()
    assert (res["scores"][:, 0] >= res["paired"] - 1e-6).all()
    zp = torch.from_numpy(res["embeds"]).cuda()
    exact = zp @ zc.T                                   # 6000 x 117000 fp32 on the device, test only
    top_s, top_i = torch.topk(exact, 2, dim=1)
    assert (torch.from_numpy(res["scores"][:, 0]).cuda() - top_s[:, 0]).abs().max().item() < 1e-4
    clear = ((top_s[:, 0] - top_s[:, 1]) >= 1e-3).cpu().numpy()
    assert (res["idx"][clear, 0] == top_i[:, 0].cpu().numpy()[clear]).all()
    paired = exact[torch.arange(len(preds), device="cuda"), torch.from_numpy(t).long().cuda()].cpu().numpy()
    assert np.abs(res["paired"] - paired).max() < 1e-4


@pytest.mark.parametrize("seed,lo,hi", [(0, 1, 65), (1, 1, 9), (2, 30, 65), (3, 60, 65), (4, 1, 3), (5, 10, 40)])
def test_random_length_mixes_are_deterministic_and_match(minilm, engines, seed, lo, hi):
    """Blocks with very few or very many sequences put different numbers of warps to work in the kernel's two
    attention groups; results must not depend on warp timing (two runs are bit-identical) and match the per-GEMM path."""
    fused, plain = engines
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi, size=int(40000 / ((lo + hi) / 2)))
    ids = rng.integers(5, minilm.arch.vocab, size=int(lens.sum())).astype(np.int32)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    z1 = fused.encode(ids, cu)
    z2 = fused.encode(ids, cu)
    assert np.array_equal(z1, z2)
    assert np.isfinite(z1).all()
    zp = plain.encode(ids, cu)
    assert np.abs(z1 - zp).max() < 1e-3


def test_other_depth_and_ffn_width(monkeypatch):
    """The fused kernel is generic in depth and FFN width (multiples of 128): 3 layers, FFN 1152 (an ODD number of
    128-wide slices, which exercises the alternating accumulator barriers) against the per-GEMM path and the oracle."""
    from oracle.hf_models import build_hf_tokenizer, build_model
    from owc_b200 import arch as A

    arch = A.EncoderArch("bert-3x1152", A.KIND_BERT, layers=3, hidden=384, heads=12, inter=1152, vocab=4096,
                         max_pos=512, type_vocab=2, ln_eps=1e-12)
    vocab = synth.build_vocab("bert", arch.vocab)
    words = synth.vocab_words(vocab)
    model = build_model(arch)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    tok = HostTokenizer(vocab, "bert")
    texts = _ragged_texts(words, 2500, 13, lo=1, hi=20)
    ids, cu = tok.encode_packed(texts)
    assert np.diff(cu).max() <= 64
    monkeypatch.setenv("OWC_FUSED", "1")
    fused = Engine(arch, sd, device=0)
    monkeypatch.setenv("OWC_FUSED", "0")
    plain = Engine(arch, sd, device=0)
    try:
        zf, nf = _launches(fused, lambda: fused.encode(ids, cu))
        zp = plain.encode(ids, cu)
        assert nf <= 3                                   # embed + fused + pool
        assert np.abs(zf - zp).max() < 1e-3
        want = encoder_ref.encode_strings(sd, arch, build_hf_tokenizer(arch, vocab), texts[:300]).numpy()
        assert np.abs(zf[:300] - want).max() < 1e-3
    finally:
        fused.close()
        plain.close()


def test_multi_task_sweep_equals_per_file_scoring(minilm, engines):
    """BASELINE config 5 in miniature: several models' prediction files over several tasks; the sweep (one class
    encode per task, files concatenated) returns what scoring every file on its own returns."""
    from owc_b200.sweep import score_sweep

    fused, _ = engines
    tok = HostTokenizer(minilm.vocab, "bert")
    tasks = {}
    for ti, ds in enumerate(("caltech101", "dtd", "eurosat", "oxford_pets")):
        names = synth.real_class_names(ds)
        files = {}
        for mi in range(3):
            preds, targets = synth.make_predictions(700 + 150 * mi, names, minilm.words, seed=100 * ti + mi)
            files[f"model{mi}"] = {"predictions": preds, "targets": targets}
        tasks[ds] = {"class_names": names, "files": files}
    tasks["empty_task"] = {"class_names": ["a", "b", "c", "d", "e"], "files": {}}
    got = score_sweep(fused, tok, tasks, k=5)
    assert set(got) == set(tasks) and got["empty_task"] == {}
    for ds, spec in tasks.items():
        if not spec["files"]:
            continue
        cids, ccu = tok.encode_packed(spec["class_names"])
        zc = fused.encode(cids, ccu, out="device")
        for key, f in spec["files"].items():
            pids, pcu = tok.encode_packed(f["predictions"])
            one = fused.score(pids, pcu, zc, targets=np.asarray(f["targets"], np.int32), k=5)
            assert np.array_equal(one["idx"], got[ds][key]["idx"]), (ds, key)
            assert np.abs(one["scores"] - got[ds][key]["scores"]).max() < 1e-6
            assert np.abs(one["paired"] - got[ds][key]["paired"]).max() < 1e-6
