"""Encoder weights for runs without network access: seeded random initialisation.

The north star asks for "identical random-init weights" on both sides (there is no network, hence
no pretrained checkpoint). They come from the HF constructors with their default
``initializer_range=0.02`` after ``torch.manual_seed(seed)`` - the reference's default seed is 1234
(``eval_metrics.py:190-195``). This is initialisation only; no arithmetic of the scoring path
happens here.
"""

from __future__ import annotations

from . import arch as A

__all__ = ["hf_random_init", "random_state_dict", "DEFAULT_SEED"]

DEFAULT_SEED = 1234


def hf_random_init(arch: A.EncoderArch, seed: int = DEFAULT_SEED):
    """Seeded random-init ``transformers`` ``BertModel`` / ``MPNetModel`` of ``arch`` in eval mode."""
    import torch
    from transformers import BertConfig, BertModel, MPNetConfig, MPNetModel

    torch.manual_seed(seed)
    if arch.kind == A.KIND_BERT:
        cfg = BertConfig(vocab_size=arch.vocab, hidden_size=arch.hidden, num_hidden_layers=arch.layers,
                         num_attention_heads=arch.heads, intermediate_size=arch.inter,
                         max_position_embeddings=arch.max_pos, type_vocab_size=arch.type_vocab,
                         layer_norm_eps=arch.ln_eps, hidden_act="gelu")
        return BertModel(cfg).eval()
    cfg = MPNetConfig(vocab_size=arch.vocab, hidden_size=arch.hidden, num_hidden_layers=arch.layers,
                      num_attention_heads=arch.heads, intermediate_size=arch.inter,
                      max_position_embeddings=arch.max_pos, layer_norm_eps=arch.ln_eps,
                      relative_attention_num_buckets=arch.rel_buckets)
    return MPNetModel(cfg).eval()


def random_state_dict(arch: A.EncoderArch, seed: int = DEFAULT_SEED) -> dict:
    return {k: v.detach() for k, v in hf_random_init(arch, seed).state_dict().items()}
