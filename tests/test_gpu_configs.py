"""GPU parity at the BASELINE configurations the round-1 tests left to self-comparisons (pytest -m gpu):

* c3 - mpnet-base arch x 1000 class names: scores / paired / argmax against the golden produced by the REFERENCE itself
  (tests/golden/mpnet_c3.npz, oracle/make_golden.py --only-c3), bf16 and fp32 check mode, and the wide path
  (persistent GEMMs, deferred LayerNorm) against the per-GEMM path;
* c4 - a sample of the 117 k-name vocabulary encoded by the CUDA path against the oracle, top-5 against the oracle's S;
* c5 - the multi-task sweep against the oracle on a sample per task;
* GELU at pretrained-like activation ranges; owc_match; target validation.
"""

import os

import numpy as np
import pytest
import torch

import owc_b200  # noqa: F401
from oracle import encoder_ref
from oracle.hf_models import build_hf_tokenizer, build_model
from owc_b200 import arch as A
from owc_b200 import synth
from owc_b200._cabi import OwcError
from owc_b200.engine import Engine
from owc_b200.tokenizer import HostTokenizer

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"fp32": 1e-4, "bf16": 1e-2}


@pytest.fixture(scope="module")
def c3(mpnet):
    g = dict(np.load(os.path.join(GOLDEN, "mpnet_c3.npz"), allow_pickle=False))
    cls = synth.make_class_names(int(g["n_classes"]), mpnet.words, seed=int(g["seed"]))
    preds, targets = synth.make_predictions(int(g["n_preds"]), cls, mpnet.words, seed=int(g["seed"]))
    assert [str(s) for s in g["class_name_head"]] == cls[:8] and [str(s) for s in g["pred_head"]] == preds[:8]
    assert np.array_equal(np.asarray(targets, np.int32), g["targets"])
    return g, cls, preds


@pytest.fixture(scope="module")
def eng_mpnet(mpnet):
    e = Engine(mpnet.arch, mpnet.sd, device=0)
    yield e
    e.close()


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_c3_scores_against_reference_golden(mpnet, eng_mpnet, c3, precision):
    g, cls, preds = c3
    eng, tok = eng_mpnet, HostTokenizer(mpnet.vocab, "mpnet")
    cids, ccu = tok.encode_packed(cls)
    zc = eng.encode(cids, ccu, out="device", precision=precision)
    assert np.abs(zc[:32].cpu().numpy() - g["class_embeds_head"]).max() < TOL[precision] / (10 if precision == "bf16" else 1)
    pids, pcu = tok.encode_packed(preds)
    res = eng.score(pids, pcu, zc, targets=g["targets"], k=5, precision=precision, want_embeds=True)
    assert np.abs(res["embeds"][:64] - g["pred_embeds_head"]).max() < TOL[precision] / (10 if precision == "bf16" else 1)
    assert np.abs(res["paired"] - g["paired_scores"]).max() < TOL[precision]          # semantic_similarity of the reference
    assert np.abs(res["scores"] - g["top5_scores"]).max() < TOL[precision]
    clear = g["margin"] >= 1e-3
    agree = res["idx"][:, 0] == g["top5_idx"][:, 0]
    assert clear.sum() > 1500
    assert agree[clear].all(), f"argmax differs off near-ties at rows {np.nonzero(~agree & clear)[0][:10]}"
    assert agree.mean() >= 0.90


def test_c3_wide_path_equals_per_gemm_path(mpnet, c3, monkeypatch):
    """Deferred-LayerNorm persistent GEMMs (gemm_wide.cuh) against the per-GEMM path with explicit LayerNorm kernels:
    same bf16 operand precision, so the embeddings agree far inside the tolerance; chunking is invisible; few launches."""
    _, cls, preds = c3
    tok = HostTokenizer(mpnet.vocab, "mpnet")
    ids, cu = tok.encode_packed(preds + cls)
    wide = Engine(mpnet.arch, mpnet.sd, device=0)
    small = Engine(mpnet.arch, mpnet.sd, device=0, chunk_tokens=2048)      # many chunks, column splits inside each
    monkeypatch.setenv("OWC_WIDE", "0")
    plain = Engine(mpnet.arch, mpnet.sd, device=0)
    try:
        s0 = wide.stats()["kernel_launches"]
        zw = wide.encode(ids, cu)
        n_launch = wide.stats()["kernel_launches"] - s0
        zs = small.encode(ids, cu)
        zp = plain.encode(ids, cu)
        assert np.isfinite(zw).all() and np.abs(np.linalg.norm(zw, axis=1) - 1).max() < 1e-5
        assert np.abs(zw - zp).max() < 2e-3
        # another chunking regroups the fp32 partial row statistics (column splits), and a last-bit change there flips
        # bf16 roundings downstream: same tolerance as against the other path, not bit-identity
        assert np.abs(zw - zs).max() < 2e-3
        assert n_launch <= 2 + 5 * mpnet.arch.layers          # embed + 12 x (QKV, attention, O, FFN1, FFN2) + pool: one chunk
        assert np.array_equal(wide.encode(ids, cu), zw)       # no atomics anywhere: run-to-run bit-identical
        assert np.array_equal(small.encode(ids, cu), zs)
    finally:
        wide.close()
        small.close()
        plain.close()


def test_c3_full_size_properties(mpnet, eng_mpnet, c3):
    """c3 at a size several chunks deep (100 k rows): size-independent properties + a 1 k-row oracle sample."""
    _, cls, _ = c3
    eng, tok = eng_mpnet, HostTokenizer(mpnet.vocab, "mpnet")
    preds, targets = synth.make_predictions(100_000, cls, mpnet.words, seed=4321)
    t = np.asarray(targets, np.int32)
    cids, ccu = tok.encode_packed(cls)
    zc = eng.encode(cids, ccu, out="device")
    pids, pcu = tok.encode_packed(preds)
    res = eng.score(pids, pcu, zc, targets=t, k=5, want_embeds=True)
    assert np.abs(np.linalg.norm(res["embeds"], axis=1) - 1).max() < 1e-5
    assert (np.diff(res["scores"], axis=1) <= 0).all()
    assert (res["scores"][:, 0] >= res["paired"] - 1e-6).all()
    hit = res["idx"][:, 0] == t
    assert hit.sum() > 50 and np.abs(res["scores"][hit, 0] - res["paired"][hit]).max() < 1e-6   # targets are uniform-random: ~0.1 % hits
    perm = np.random.default_rng(1).permutation(len(preds))[:4000]
    ids2, cu2 = tok.encode_packed([preds[i] for i in perm])
    res2 = eng.score(ids2, cu2, zc, targets=t[perm], k=5)
    assert np.abs(res2["paired"] - res["paired"][perm]).max() < 1e-3         # another chunk layout, same rows
    sample = perm[:1000]
    zp_ref = encoder_ref.encode_strings(mpnet.sd, mpnet.arch, mpnet.hf_tok, [preds[i] for i in sample])
    zc_ref = encoder_ref.encode_strings(mpnet.sd, mpnet.arch, mpnet.hf_tok, cls)
    s_ref, i_ref, margin = encoder_ref.sim_topk(zp_ref, zc_ref, 5)
    assert np.abs(res["scores"][sample] - s_ref.numpy()).max() < 1e-2
    clear = margin.numpy() >= 1e-3
    assert (res["idx"][sample, 0] == i_ref[:, 0].numpy())[clear].all()


def test_c4_vocabulary_sample_against_oracle(minilm):
    """c4: 117 k names. A 2048-name sample of the CUDA-encoded vocabulary against the oracle's embeddings, and the fused
    top-5 of 1500 predictions over those 2048 names against the oracle's S (the full-vocabulary matcher is compared with
    an exact product of the same embeddings in test_gpu_fused.py::test_config4_scale_properties)."""
    b = minilm
    tok = HostTokenizer(b.vocab, "bert")
    names = synth.make_class_names(117_000, b.words, seed=1234)
    eng = Engine(b.arch, b.sd, device=0)
    try:
        cids, ccu = tok.encode_packed(names)
        zc_all = eng.encode(cids, ccu, out="device")
        pick = np.random.default_rng(5).choice(len(names), 2048, replace=False)
        pick.sort()
        sub = [names[i] for i in pick]
        zc_ref = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, sub)
        assert (zc_all[torch.from_numpy(pick).cuda()].cpu() - zc_ref).abs().max().item() < 1e-3
        preds, targets = synth.make_predictions(1500, sub, b.words, seed=11)
        zp_ref = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, preds)
        s_ref, i_ref, margin = encoder_ref.sim_topk(zp_ref, zc_ref, 5)
        pids, pcu = tok.encode_packed(preds)
        zc_sub = zc_all[torch.from_numpy(pick).cuda()].contiguous()
        res = eng.score(pids, pcu, zc_sub, targets=np.asarray(targets, np.int32), k=5)
        assert np.abs(res["scores"] - s_ref.numpy()).max() < 1e-2
        clear = margin.numpy() >= 1e-3
        agree = res["idx"][:, 0] == i_ref[:, 0].numpy()
        assert agree[clear].all() and agree.mean() >= 0.9
        paired_ref = (zp_ref * zc_ref[torch.tensor(targets)]).sum(1).numpy()
        assert np.abs(res["paired"] - paired_ref).max() < 1e-2
    finally:
        eng.close()


def test_c5_sweep_against_oracle(minilm):
    """c5 in miniature: the sweep's per-file results against the ORACLE (not against the same engine) on every task."""
    from owc_b200.sweep import score_sweep

    b = minilm
    tok = HostTokenizer(b.vocab, "bert")
    eng = Engine(b.arch, b.sd, device=0)
    try:
        tasks = {}
        for ti, ds in enumerate(synth.REAL_CLASS_SETS):
            names = synth.real_class_names(ds)
            files = {}
            for mi in range(2):
                preds, targets = synth.make_predictions(150 + 30 * mi, names, b.words, seed=500 * ti + mi)
                files[f"m{mi}"] = {"predictions": preds, "targets": targets}
            tasks[ds] = {"class_names": names, "files": files}
        got = score_sweep(eng, tok, tasks, k=5)
        for ds, spec in tasks.items():
            zc_ref = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, spec["class_names"])
            for key, f in spec["files"].items():
                zp_ref = encoder_ref.encode_strings(b.sd, b.arch, b.hf_tok, f["predictions"])
                s_ref, i_ref, margin = encoder_ref.sim_topk(zp_ref, zc_ref, 5)
                r = got[ds][key]
                assert np.abs(r["scores"] - s_ref.numpy()).max() < 1e-2, (ds, key)
                clear = margin.numpy() >= 1e-3
                assert (r["idx"][:, 0] == i_ref[:, 0].numpy())[clear].all(), (ds, key)
                paired_ref = (zp_ref * zc_ref[torch.tensor(f["targets"])]).sum(1).numpy()
                assert np.abs(r["paired"] - paired_ref).max() < 1e-2, (ds, key)
    finally:
        eng.close()


@pytest.mark.parametrize("kind", ["minilm", "mpnet"])
def test_gelu_at_pretrained_like_activation_ranges(kind):
    """With initializer_range 0.02 the FFN pre-activations are tiny and any GELU looks right. Scale the FFN1 weights so
    |pre-activation| reaches 1-4 (where the tanh and erf forms differ most, 4.8e-4) and hold the production kernels to
    the oracle's exact-erf GELU (modeling_bert.py:330-342): cosine error <= 1e-2, argmax identical off near-ties."""
    arch = (A.EncoderArch("bert-gelu", A.KIND_BERT, layers=6, hidden=384, heads=12, inter=1536, vocab=4096, max_pos=512,
                          type_vocab=2, ln_eps=1e-12) if kind == "minilm" else
            A.EncoderArch("mpnet-gelu", A.KIND_MPNET, layers=4, hidden=768, heads=12, inter=3072, vocab=4096, max_pos=514,
                          type_vocab=0, ln_eps=1e-5, rel_buckets=32, rel_max_distance=128, pad_id=1))
    vkind = "bert" if kind == "minilm" else "mpnet"
    vocab = synth.build_vocab(vkind, arch.vocab)
    words = synth.vocab_words(vocab)
    model = build_model(arch)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k in sd:
        if k.endswith("intermediate.dense.weight"):
            sd[k] = sd[k] * 5.0                                  # pre-activation std ~2: most elements inside |x| in [1, 4]
        if k.endswith("intermediate.dense.bias"):
            sd[k] = sd[k] + 0.5 * torch.randn_like(sd[k])
    hf_tok = build_hf_tokenizer(arch, vocab)
    tok = HostTokenizer(vocab, vkind)
    cls = synth.make_class_names(64, words, seed=3)
    preds, _ = synth.make_predictions(600, cls, words, seed=4)
    # the stress is real: FFN1 pre-activations on a LayerNorm output of the oracle (layer 0's output stands in for layer
    # 1's attention-block output: both are unit-variance LayerNorm rows)
    enc = hf_tok(preds[:64], padding=True, return_tensors="pt")
    hs: list = []
    with torch.no_grad():
        encoder_ref.encoder_forward(sd, arch, enc["input_ids"], enc["attention_mask"], hidden_states=hs)
        pre = hs[1] @ sd["encoder.layer.1.intermediate.dense.weight"].T + sd["encoder.layer.1.intermediate.dense.bias"]
    assert pre.abs().gt(1.0).float().mean() > 0.3
    zc_ref = encoder_ref.encode_strings(sd, arch, hf_tok, cls)
    zp_ref = encoder_ref.encode_strings(sd, arch, hf_tok, preds)
    s_ref, i_ref, margin = encoder_ref.sim_topk(zp_ref, zc_ref, 5)
    eng = Engine(arch, sd, device=0)
    try:
        cids, ccu = tok.encode_packed(cls)
        zc = eng.encode(cids, ccu, out="device")
        pids, pcu = tok.encode_packed(preds)
        res = eng.score(pids, pcu, zc, k=5, want_embeds=True)
        cos = (torch.from_numpy(res["embeds"]) * zp_ref).sum(1)
        assert (1 - cos).abs().max().item() < 1e-2
        assert np.abs(res["scores"] - s_ref.numpy()).max() < 1e-2
        clear = margin.numpy() >= 1e-3
        agree = res["idx"][:, 0] == i_ref[:, 0].numpy()
        assert agree[clear].all() and agree.mean() >= 0.9
    finally:
        eng.close()


def test_match_equals_score_and_targets_are_validated(minilm):
    b = minilm
    tok = HostTokenizer(b.vocab, "bert")
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]]
    t = b.gold["targets"]
    eng = Engine(b.arch, b.sd, device=0)
    try:
        cids, ccu = tok.encode_packed(cls)
        zc = eng.encode(cids, ccu, out="device")
        pids, pcu = tok.encode_packed(preds)
        one = eng.score(pids, pcu, zc, targets=t, k=5)
        zp = eng.encode(pids, pcu, out="device")
        for prec in ("bf16", "fp32"):
            ref = one if prec == "bf16" else eng.score(pids, pcu, zc, targets=t, k=5, precision="fp32")
            zpp = zp if prec == "bf16" else eng.encode(pids, pcu, out="device", precision="fp32")
            two = eng.match(zpp, zc, targets=t, k=5, precision=prec)
            assert np.array_equal(two["idx"], ref["idx"])
            assert np.abs(two["scores"] - ref["scores"]).max() < 1e-6 and np.abs(two["paired"] - ref["paired"]).max() < 1e-6
        dev = eng.match(zp, zc, targets=torch.from_numpy(t).cuda(), k=5, out="device")
        torch.cuda.synchronize()
        assert np.array_equal(dev["idx"].cpu().numpy(), one["idx"])
        assert eng.match(zp, zc, k=1)["paired"] is None
        # out-of-range targets: host arrays are rejected up front, device tensors by the kernel (clamped, call fails)
        bad = t.copy()
        bad[3] = len(cls)
        with pytest.raises(IndexError):
            eng.score(pids, pcu, zc, targets=bad, k=5)
        with pytest.raises(IndexError):
            eng.match(zp, zc, targets=-np.ones_like(t), k=5)
        eng.score(torch.from_numpy(pids).cuda(), pcu, zc, targets=torch.from_numpy(bad).cuda(), k=5, out="device")
        with pytest.raises(OwcError, match="out of range"):
            eng.sync()
        again = eng.score(pids, pcu, zc, targets=t, k=5)          # the engine is still usable
        assert np.array_equal(again["idx"], one["idx"])
    finally:
        eng.close()


def test_two_contexts_in_one_process_share_nothing(minilm):
    """The shared-memory opt-in of the tcgen05 kernels is tracked per context (a second context must set it again on
    its own device; on a one-GPU box the second context lands on the same device and must simply work)."""
    b = minilm
    tok = HostTokenizer(b.vocab, "bert")
    ids, cu = tok.encode_packed([str(s) for s in b.gold["preds"]])
    dev2 = 1 if torch.cuda.device_count() > 1 else 0
    e1, e2 = Engine(b.arch, b.sd, device=0), Engine(b.arch, b.sd, device=dev2)
    try:
        z1, z2 = e1.encode(ids, cu), e2.encode(ids, cu)
        assert np.array_equal(z1, z2)
        long_ids = np.concatenate([ids[:2]] + [ids[1:2]] * 200 + [ids[-1:]]).astype(np.int32)   # > 64 tokens: SIMT attention path
        zl1 = e1.encode(long_ids, np.array([0, len(long_ids)], np.int32))
        zl2 = e2.encode(long_ids, np.array([0, len(long_ids)], np.int32))
        assert np.array_equal(zl1, zl2)
    finally:
        e1.close()
        e2.close()


@pytest.mark.parametrize("seed,lo,hi", [(0, 1, 65), (1, 1, 9), (2, 30, 65), (3, 60, 65), (4, 1, 3), (5, 14, 20)])
def test_qkv_attention_kernel_on_ragged_lengths(mpnet, monkeypatch, seed, lo, hi):
    """The QKV projection with the attention in its epilogue (qkv_attn.cuh): row blocks of whole sequences, several
    short sequences per 16 x 16 tile, 32- and 64-row tiles for longer ones, block boundaries anywhere. Against the
    separate-kernel path (same bf16 operands: tight) and the oracle; bit-identical run to run."""
    b = mpnet
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi, size=int(30000 / ((lo + hi) / 2)))
    ids = rng.integers(5, b.arch.vocab - 1, size=int(lens.sum())).astype(np.int32)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    monkeypatch.setenv("OWC_QKV_ATTN", "1")
    fused = Engine(b.arch, b.sd, device=0)
    monkeypatch.setenv("OWC_QKV_ATTN", "0")
    split = Engine(b.arch, b.sd, device=0)
    try:
        s0 = fused.stats()["kernel_launches"]
        z1 = fused.encode(ids, cu)
        assert fused.stats()["kernel_launches"] - s0 == 2 + 4 * b.arch.layers       # embed + 12 x (QKV+attn, O, FFN1, FFN2) + pool
        assert np.isfinite(z1).all() and np.array_equal(z1, fused.encode(ids, cu))
        z0 = split.encode(ids, cu)
        assert np.abs(z1 - z0).max() < 2e-3
        n = 96
        pad = np.full((n, int(lens[:n].max())), b.arch.pad_id, np.int64)
        mask = np.zeros_like(pad)
        for i in range(n):
            pad[i, :lens[i]] = ids[cu[i]:cu[i + 1]]
            mask[i, :lens[i]] = 1
        want = encoder_ref.encode_padded(b.sd, b.arch, torch.from_numpy(pad), torch.from_numpy(mask)).numpy()
        assert np.abs(z1[:n] - want).max() < 1e-3
    finally:
        fused.close()
        split.close()


def test_qkv_attention_early_release_is_stable_at_scale(mpnet, monkeypatch):
    """k_qkv_attn gives the q|k|v rows of a head back to the drain warps as soon as every attention warp holds its unit in
    registers (one shared buffer). A hole in that hand-over would show as run-to-run differences on a launch that keeps all
    148 SMs busy for many heads: several chunks of label-length sequences, six runs, every run bit-identical, and equal to
    the separate-kernel path within bf16 noise."""
    b = mpnet
    rng = np.random.default_rng(11)
    lens = rng.integers(1, 17, size=48_000)                      # ~400 k tokens: three chunks of 1 184 row blocks
    ids = rng.integers(5, b.arch.vocab - 1, size=int(lens.sum())).astype(np.int32)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    monkeypatch.setenv("OWC_QKV_ATTN", "1")
    fused = Engine(b.arch, b.sd, device=0)
    monkeypatch.setenv("OWC_QKV_ATTN", "0")
    split = Engine(b.arch, b.sd, device=0)
    try:
        z0 = fused.encode(ids, cu)
        assert np.isfinite(z0).all()
        for _ in range(5):
            assert np.array_equal(z0, fused.encode(ids, cu))
        assert np.abs(z0 - split.encode(ids, cu)).max() < 2e-3
    finally:
        fused.close()
        split.close()


@pytest.mark.parametrize("case", ["empties", "one_token_rows", "full_blocks", "single_row", "odd_blocks"])
def test_qkv_attention_kernel_block_edges(mpnet, monkeypatch, case):
    """Row-block packing edges of k_qkv_attn on the wide path: empty sequences between labels (NaN rows, as `_text.py:202`
    has no epsilon), blocks closed by the 128-sequence cap (one-token rows: eight units of sixteen sequences), sequences
    that fill blocks exactly, a single row, an odd number of blocks (the pair's second CTA idles). Against the
    separate-kernel path."""
    b = mpnet
    rng = np.random.default_rng(5)
    if case == "empties":
        lens = rng.integers(0, 9, size=3000)
        lens[::7] = 0
    elif case == "one_token_rows":
        lens = np.ones(1000, np.int64)
        lens[500:] = rng.integers(1, 4, size=500)
    elif case == "full_blocks":
        lens = np.asarray([16] * 64 + [8] * 32 + [64, 64] * 3 + [5] * 40, np.int64)
    elif case == "single_row":
        lens = np.asarray([7], np.int64)
    else:
        lens = np.asarray([16] * 8 * 3, np.int64)          # exactly three blocks
    ids = rng.integers(5, b.arch.vocab - 1, size=int(lens.sum())).astype(np.int32)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    monkeypatch.setenv("OWC_QKV_ATTN", "1")
    fused = Engine(b.arch, b.sd, device=0)
    monkeypatch.setenv("OWC_QKV_ATTN", "0")
    split = Engine(b.arch, b.sd, device=0)
    try:
        z1, z0 = fused.encode(ids, cu), split.encode(ids, cu)
        empty = lens == 0
        assert np.isnan(z1[empty]).all() and np.isnan(z0[empty]).all()
        assert np.isfinite(z1[~empty]).all()
        assert np.abs(z1[~empty] - z0[~empty]).max() < 2e-3
        assert np.array_equal(z1[~empty], fused.encode(ids, cu)[~empty])
    finally:
        fused.close()
        split.close()
