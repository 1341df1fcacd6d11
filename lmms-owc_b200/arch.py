"""Encoder architectures on the scoring path and the packed-weight blob handed to the C-ABI.

The reference hard-codes ``sentence-transformers/all-MiniLM-L6-v2`` (a 6-layer BERT,
``src/data/pipelines/text/_text.py:161``); BASELINE.json's config 3 adds the ``all-mpnet-base-v2``
architecture. Only the architecture constants live here; there are no pretrained weights offline,
so weights are always a ``state_dict`` supplied by the caller (random-init in tests and bench).

Blob layout (all fp32, little-endian, in this order; ``owc_encoder_load`` in
``include/owc_b200.h`` documents the same):

    word_emb [vocab, H] | pos_emb [max_pos, H] | type_emb [type_vocab, H] (BERT only)
    emb_ln_gamma [H] | emb_ln_beta [H]
    rel_bias_lut [heads, 2*max_pos-1] (MPNet only; index = (k - q) + max_pos - 1)
    per layer: Wqkv [3H, H] | bqkv [3H] | Wo [H, H] | bo [H] | ln1_gamma [H] | ln1_beta [H]
               W1 [I, H] | b1 [I] | W2 [H, I] | b2 [H] | ln2_gamma [H] | ln2_beta [H]

Every weight matrix keeps ``nn.Linear``'s ``[out, in]`` row-major layout, i.e. it is K-major for
``C[M, N] = A[M, K] . W[N, K]^T`` - exactly what the TMA descriptors of the GEMM kernels want.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

__all__ = ["EncoderArch", "MINILM_L6", "MPNET_BASE", "ARCHS", "pack_weights", "blob_num_floats",
           "mpnet_relative_bucket"]

KIND_BERT = 0
KIND_MPNET = 1


@dataclass(frozen=True)
class EncoderArch:
    """Architecture constants of one sentence encoder (mirrors ``owc_arch`` in the C header)."""

    name: str
    kind: int  # KIND_BERT | KIND_MPNET
    layers: int
    hidden: int
    heads: int
    inter: int
    vocab: int
    max_pos: int
    type_vocab: int
    ln_eps: float
    rel_buckets: int = 0
    rel_max_distance: int = 128
    pad_id: int = 0

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def max_tokens(self) -> int:
        """Longest sequence (incl. specials) the position table supports after truncation."""
        return self.max_pos if self.kind == KIND_BERT else self.max_pos - 2

    def with_vocab(self, vocab: int) -> "EncoderArch":
        return EncoderArch(**{**self.__dict__, "vocab": vocab})

    def flops_per_sequence(self, tokens: int) -> float:
        """Algorithmic encoder FLOPs for one ``tokens``-long sequence."""
        h = self.hidden
        return self.layers * (2.0 * (4 * h * h + 2 * h * self.inter) * tokens + 4.0 * h * tokens * tokens)


# all-MiniLM-L6-v2 architecture (reference: _text.py:161).
MINILM_L6 = EncoderArch("minilm-l6", KIND_BERT, layers=6, hidden=384, heads=12, inter=1536,
                        vocab=30522, max_pos=512, type_vocab=2, ln_eps=1e-12, pad_id=0)
# all-mpnet-base-v2 architecture (BASELINE.json config 3).
MPNET_BASE = EncoderArch("mpnet-base", KIND_MPNET, layers=12, hidden=768, heads=12, inter=3072,
                         vocab=30527, max_pos=514, type_vocab=0, ln_eps=1e-5, rel_buckets=32,
                         rel_max_distance=128, pad_id=1)
ARCHS = {a.name: a for a in (MINILM_L6, MPNET_BASE)}


def mpnet_relative_bucket(rel: int, num_buckets: int = 32, max_distance: int = 128) -> int:
    """T5-style bucket of ``rel = key_pos - query_pos`` (transformers ``modeling_mpnet.py:343-362``).

    Evaluated in float32 exactly like the reference's tensor expression so the truncation to an
    integer bucket lands on the same side for every distance.
    """
    n = -rel
    half = num_buckets // 2
    ret = half if n < 0 else 0
    n = abs(n)
    max_exact = half // 2
    if n < max_exact:
        return ret + n
    ratio = np.float32(np.log(np.float32(n) / np.float32(max_exact))) / np.float32(
        math.log(max_distance / max_exact))
    val = max_exact + int(np.float32(ratio * np.float32(half - max_exact)))
    return ret + min(val, half - 1)


def blob_num_floats(arch: EncoderArch) -> int:
    h, i = arch.hidden, arch.inter
    n = arch.vocab * h + arch.max_pos * h + arch.type_vocab * h + 2 * h
    if arch.kind == KIND_MPNET:
        n += arch.heads * (2 * arch.max_pos - 1)
    per_layer = 3 * h * h + 3 * h + h * h + h + 2 * h + i * h + i + h * i + h + 2 * h
    return n + arch.layers * per_layer


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def pack_weights(arch: EncoderArch, state_dict: dict) -> np.ndarray:
    """Flatten a HF ``BertModel`` / ``MPNetModel`` ``state_dict`` into the fp32 blob above."""
    sd = {k: v for k, v in state_dict.items()}
    parts: list[np.ndarray] = []

    def take(key: str, shape: tuple[int, ...]) -> None:
        a = _np(sd[key])
        if a.shape != shape:
            raise ValueError(f"{key}: expected shape {shape}, got {a.shape}")
        parts.append(a.reshape(-1))

    h, i = arch.hidden, arch.inter
    take("embeddings.word_embeddings.weight", (arch.vocab, h))
    take("embeddings.position_embeddings.weight", (arch.max_pos, h))
    if arch.kind == KIND_BERT:
        take("embeddings.token_type_embeddings.weight", (arch.type_vocab, h))
    take("embeddings.LayerNorm.weight", (h,))
    take("embeddings.LayerNorm.bias", (h,))
    if arch.kind == KIND_MPNET:
        table = _np(sd["encoder.relative_attention_bias.weight"])  # [buckets, heads]
        if table.shape != (arch.rel_buckets, arch.heads):
            raise ValueError(f"relative_attention_bias: got {table.shape}")
        p = arch.max_pos
        buckets = np.array([mpnet_relative_bucket(r, arch.rel_buckets, arch.rel_max_distance)
                            for r in range(-(p - 1), p)], dtype=np.int64)
        parts.append(np.ascontiguousarray(table[buckets].T).reshape(-1))  # [heads, 2p-1]
    for l in range(arch.layers):
        pre = f"encoder.layer.{l}."
        if arch.kind == KIND_BERT:
            q, k, v = (pre + f"attention.self.{n}" for n in ("query", "key", "value"))
            o, ln1 = pre + "attention.output.dense", pre + "attention.output.LayerNorm"
        else:
            q, k, v = (pre + f"attention.attn.{n}" for n in ("q", "k", "v"))
            o, ln1 = pre + "attention.attn.o", pre + "attention.LayerNorm"
        for n in (q, k, v):
            take(n + ".weight", (h, h))
        for n in (q, k, v):
            take(n + ".bias", (h,))
        take(o + ".weight", (h, h))
        take(o + ".bias", (h,))
        take(ln1 + ".weight", (h,))
        take(ln1 + ".bias", (h,))
        take(pre + "intermediate.dense.weight", (i, h))
        take(pre + "intermediate.dense.bias", (i,))
        take(pre + "output.dense.weight", (h, i))
        take(pre + "output.dense.bias", (h,))
        take(pre + "output.LayerNorm.weight", (h,))
        take(pre + "output.LayerNorm.bias", (h,))
    blob = np.concatenate(parts)
    assert blob.size == blob_num_floats(arch), (blob.size, blob_num_floats(arch))
    return blob
