"""ORACLE (test infrastructure): seeded random-init HF models and tokenizers.

"Identical random-init weights" (north star) are produced the same way on both sides:
``torch.manual_seed(1234)`` (the reference's default seed, ``eval_metrics.py:190-195``) followed by
the HF constructor with its default ``initializer_range=0.02``. ``transformers`` is
the third-party library the reference itself calls (``_text.py:159-163``); it is installed on the
GPU box, so tests can rebuild the weights there instead of shipping them.
"""

from __future__ import annotations

import numpy as np
import torch

from owc_b200 import arch as A

SEED = 1234


def build_model(arch: A.EncoderArch, seed: int = SEED):
    """Seeded random-init HF model of ``arch`` (SURVEY.md 8c "Weight init"); shared with the product's
    offline weight source so both sides hold bit-identical weights."""
    from owc_b200.weights import hf_random_init

    return hf_random_init(arch, seed)


def build_hf_tokenizer(arch: A.EncoderArch, vocab: list[str]):
    from transformers import BertTokenizer, MPNetTokenizer

    table = {w: i for i, w in enumerate(vocab)}
    cls = BertTokenizer if arch.kind == A.KIND_BERT else MPNetTokenizer
    tok = cls(vocab=table, do_lower_case=True, model_max_length=512)
    assert len(tok) == len(vocab), "vocab dict was not picked up (SURVEY H8)"
    return tok


def weight_checksums(model) -> np.ndarray:
    sd = model.state_dict()
    keys = sorted(k for k in sd if k.endswith("weight"))[:8]
    return np.array([float(sd[k].double().sum()) for k in keys], dtype=np.float64)
