"""Test wiring: ``gpu`` marker, repo-root import path, shared fixtures (seeded models, golden data)."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Build libowc_b200.so if it is missing (a fresh checkout has no binary: *.so is git-ignored). nvcc
    cross-compiles without a GPU; without nvcc the tests that need the library fail loudly."""
    import shutil

    import __graft_entry__ as entry

    if os.path.exists(entry.LIB) or shutil.which(os.environ.get("NVCC", "nvcc")) is None:
        return
    entry.build()


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Bundle:
    """Seeded random-init HF model + tokenizers + golden fixture of one architecture."""

    def __init__(self, arch, vocab_kind, golden_file):
        import owc_b200  # noqa: F401
        from oracle.hf_models import build_hf_tokenizer, build_model, weight_checksums
        from owc_b200 import synth

        self.arch = arch
        self.vocab = synth.build_vocab(vocab_kind, arch.vocab)
        self.words = synth.vocab_words(self.vocab)
        self.model = build_model(arch)
        self.sd = {k: v.detach() for k, v in self.model.state_dict().items()}
        self.hf_tok = build_hf_tokenizer(arch, self.vocab)
        self.gold = dict(np.load(os.path.join(GOLDEN, golden_file), allow_pickle=False))
        drift = np.abs(weight_checksums(self.model) - self.gold["weight_checksums"]).max()
        assert drift < 1e-6, f"weights drifted from the golden run (RNG/torch version?): {drift}"


@pytest.fixture(scope="session")
def minilm():
    import owc_b200  # noqa: F401
    from owc_b200 import arch as A

    return Bundle(A.MINILM_L6, "bert", "minilm_caltech101.npz")


@pytest.fixture(scope="session")
def mpnet():
    import owc_b200  # noqa: F401
    from owc_b200 import arch as A

    return Bundle(A.MPNET_BASE, "mpnet", "mpnet_synth.npz")
