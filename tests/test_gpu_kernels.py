"""GPU parity of the individual kernels through the C-ABI (run on the B200 box: pytest -m gpu)."""

import numpy as np
import pytest
import torch

import owc_b200  # noqa: F401
from oracle import encoder_ref
from owc_b200 import arch as A
from owc_b200.engine import Engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny_engine():
    from oracle.hf_models import build_model

    arch = A.EncoderArch("tiny-bert", A.KIND_BERT, layers=1, hidden=256, heads=8, inter=512, vocab=512,
                         max_pos=64, type_vocab=2, ln_eps=1e-12)
    model = build_model(arch)
    eng = Engine(arch, {k: v.detach() for k, v in model.state_dict().items()}, device=0)
    yield eng
    eng.close()


def _bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


# tile_n: column-tile width; +1000 = CTA-pair kernel (cta_group::2); +2000 = CTA pair with resident A block
@pytest.mark.parametrize("tile_n", [128, 192, 256, 1128, 1192, 1256, 2128, 2192, 2256])
@pytest.mark.parametrize("shape", [(128, 192, 64), (300, 384, 384), (1000, 1152, 384), (257, 1536, 384),
                                   (4096, 384, 1536), (129, 32, 128), (20000, 768, 768)])
def test_tcgen05_gemm_matches_torch(tiny_engine, shape, tile_n):
    m, n, k = shape
    if tile_n >= 2000 and k > 384:
        pytest.skip("resident-A variant holds at most K = 384")
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.05
    bias = torch.randn(n, device="cuda", generator=g)
    got = tiny_engine.debug_gemm(a, w, bias, precision="bf16", tile_n=tile_n)
    want = _bf16_round(a).double() @ _bf16_round(w).double().T + bias.double()
    torch.cuda.synchronize()
    err = (got.double() - want).abs().max().item()
    assert err < 2e-4 * max(1.0, want.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("shape", [(70, 100, 64), (513, 384, 384), (64, 1536, 384)])
def test_fp32_gemm_matches_torch(tiny_engine, shape):
    m, n, k = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.05
    bias = torch.randn(n, device="cuda", generator=g)
    got = tiny_engine.debug_gemm(a, w, bias, precision="fp32")
    want = a.double() @ w.double().T + bias.double()
    assert (got.double() - want).abs().max().item() < 1e-4


def test_pair_dot_with_and_without_index(tiny_engine):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((37, 384)).astype(np.float32)
    b = rng.standard_normal((1000, 384)).astype(np.float32)
    ai = rng.integers(0, 37, 1000).astype(np.int32)
    got = tiny_engine.pair_dot(a, b, a_index=ai)
    want = encoder_ref.paired_dot(torch.from_numpy(a)[torch.from_numpy(ai).long()], torch.from_numpy(b)).numpy()
    assert np.abs(got - want).max() < 1e-4
    got = tiny_engine.pair_dot(b, b)
    assert np.abs(got - (b.astype(np.float64) ** 2).sum(1)).max() < 1e-3
    bi = rng.integers(0, 1000, 5).astype(np.int32)
    got = tiny_engine.pair_dot(a[:5], b, b_index=bi)
    assert np.abs(got - (a[:5] * b[bi]).sum(1)).max() < 1e-4
    assert tiny_engine.pair_dot(a[:0], b[:0]).shape == (0,)
    with pytest.raises(IndexError):
        tiny_engine.pair_dot(a, b, a_index=np.full(1000, 37, np.int32))
    # device-resident tables (what the metric bodies pass): bit-identical to the host-buffer call
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    assert np.array_equal(tiny_engine.pair_dot(ad, bd, a_index=ai), tiny_engine.pair_dot(a, b, a_index=ai))
    assert np.array_equal(tiny_engine.pair_dot(bd, bd), tiny_engine.pair_dot(b, b))
    assert np.array_equal(tiny_engine.pair_dot(ad, bd, a_index=ai[:5], b_index=bi), tiny_engine.pair_dot(a, b, a_index=ai[:5], b_index=bi))
    with pytest.raises(ValueError):
        tiny_engine.pair_dot(ad, b, a_index=ai)


def _unit_rows(n, d, seed, spread=0.3):
    """Anisotropic unit rows (common component + noise) like random-init sentence embeddings."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1, d, generator=torch.Generator().manual_seed(99))
    z = base + spread * torch.randn(n, d, generator=g)
    return z / z.norm(dim=1, keepdim=True)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("n,c,d,k", [(300, 100, 384, 1), (1000, 101, 384, 5), (257, 5, 384, 5), (129, 1000, 768, 5),
                                     (64, 3000, 384, 16), (5000, 700, 384, 5), (1, 9, 384, 1)])
def test_sim_topk_matches_oracle(tiny_engine, n, c, d, k, precision):
    zp, zc = _unit_rows(n, d, 1), _unit_rows(c, d, 2)
    s_ref, i_ref, margin = encoder_ref.sim_topk(zp.double(), zc.double(), k)
    scores, idx = tiny_engine.sim_topk(zp.numpy(), zc.numpy(), k, precision=precision)
    assert scores.shape == (n, k) and idx.dtype == np.int32
    # scores are exact fp32 dot products of the chosen classes in both modes
    assert np.abs(scores - s_ref.numpy()).max() < 2e-6
    assert (np.diff(scores, axis=1) <= 0).all()
    exact = (zp.double() @ zc.double().T).numpy()
    assert np.abs(np.take_along_axis(exact, idx.astype(np.int64), 1) - scores).max() < 2e-6
    if precision == "fp32":
        clear = margin.numpy() >= 1e-6
    else:
        clear = margin.numpy() >= 1e-3
    assert (idx[clear, 0] == i_ref[clear, 0].numpy()).all()


def test_sim_topk_device_tensors_and_ties(tiny_engine):
    zc = _unit_rows(40, 384, 3)
    zc[17] = zc[5]                      # exact duplicate class rows: lower index must win
    zp = zc[[5, 17, 30]].clone()
    for prec in ("bf16", "fp32"):
        scores, idx = tiny_engine.sim_topk(zp.cuda(), zc.cuda(), 2, precision=prec)
        torch.cuda.synchronize()
        assert idx[:, 0].tolist() == [5, 5, 30] and idx[:2, 1].tolist() == [17, 17]
        assert torch.allclose(scores[:, 0].cpu(), torch.ones(3), atol=1e-5)


def test_sim_topk_rejects_bad_k(tiny_engine):
    from owc_b200._cabi import OwcError

    z = _unit_rows(4, 384, 1).numpy()
    with pytest.raises(OwcError):
        tiny_engine.sim_topk(z, z, 5)       # k > c
    with pytest.raises(OwcError):
        tiny_engine.sim_topk(z, z, 0)


def test_sim_kernel_bench_hook_runs(tiny_engine):
    ms = tiny_engine.bench_sim_kernel(4096, 1000, 384, 5, 3)
    assert ms > 0


def test_sim_topk_at_the_117k_vocabulary(tiny_engine):
    """BASELINE config 4 scale: 117 000 class rows (WordNet-noun scale), fused top-5, against exact fp64 scores
    computed in row slabs (the n x c matrix is never held in one piece here either)."""
    n, c, d, k = 1536, 117_000, 384, 5
    zp, zc = _unit_rows(n, d, 11, spread=0.5), _unit_rows(c, d, 12, spread=0.5)
    scores, idx = tiny_engine.sim_topk(zp.cuda(), zc.cuda(), k, precision="bf16")
    torch.cuda.synchronize()
    scores, idx = scores.cpu().numpy(), idx.cpu().numpy()
    assert (np.diff(scores, axis=1) <= 0).all() and idx.min() >= 0 and idx.max() < c
    zc64 = zc.double()
    for lo in range(0, n, 256):
        s = zp[lo:lo + 256].double() @ zc64.T
        top_s, top_i = torch.topk(s, k + 1, dim=1)
        got_s, got_i = scores[lo:lo + 256], idx[lo:lo + 256]
        # returned scores are exact dot products of the returned classes
        assert np.abs(np.take_along_axis(s.numpy(), got_i.astype(np.int64), 1) - got_s).max() < 2e-6
        # wherever rank j and rank j+1 are separated by more than the bf16 candidate noise, rank j is identical
        gap = (top_s[:, :k] - top_s[:, 1:]).numpy()
        clear = gap >= 1e-3
        assert (got_i[clear] == top_i[:, :k].numpy()[clear]).all()
        assert np.abs(got_s[:, 0] - top_s[:, 0].numpy()).max() < 1e-3
