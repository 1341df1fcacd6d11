"""Host-side tokenisation and packing (tokenisation stays on the host, per the north star).

Mirrors what the reference gets from ``AutoTokenizer`` at ``src/data/pipelines/text/_text.py:193-195``
(``padding=True, truncation=True``): BERT normaliser (lower-case, clean text, CJK spacing), BERT
pre-tokeniser, WordPiece with ``##`` continuation and ``[UNK]``, ``[CLS] A [SEP]`` template (or
``<s> A </s>`` for MPNet) and truncation to the model's maximum length. Padding is NOT produced:
the device path takes ragged sequences packed back to back with a ``cu_seqlens`` offset array, and
every kernel is padding-invariant, so pad-to-longest would only add work.
"""

from __future__ import annotations

import numpy as np

__all__ = ["HostTokenizer", "pack_ids"]


def pack_ids(seqs: list[list[int]]) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate ragged id lists into ``(ids_packed int32 [M], cu_seqlens int32 [n+1])``."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    cu = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=cu[1:])
    if cu[-1] >= 2**31:
        raise ValueError("packed batch exceeds 2^31 tokens; split it")
    ids = np.fromiter((t for s in seqs for t in s), dtype=np.int32, count=int(cu[-1]))
    return ids, cu.astype(np.int32)


class HostTokenizer:
    """WordPiece tokenizer over an in-memory vocabulary (``tokenizers`` Rust backend)."""

    def __init__(self, vocab: list[str], kind: str = "bert", max_length: int = 512) -> None:
        from tokenizers import Tokenizer, normalizers, pre_tokenizers, processors
        from tokenizers.models import WordPiece

        self.kind = kind
        self.max_length = max_length
        table = {w: i for i, w in enumerate(vocab)}
        if kind == "bert":
            cls, sep, pad = "[CLS]", "[SEP]", "[PAD]"
            specials = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
        elif kind == "mpnet":
            cls, sep, pad = "<s>", "</s>", "<pad>"
            specials = ["<s>", "<pad>", "</s>", "[UNK]", "<mask>"]
        else:
            raise ValueError(f"unknown tokenizer kind {kind!r}")
        for s in (cls, sep, pad, "[UNK]"):
            if s not in table:
                raise ValueError(f"vocabulary lacks the special token {s}")
        tok = Tokenizer(WordPiece(vocab=table, unk_token="[UNK]", continuing_subword_prefix="##",
                                  max_input_chars_per_word=100))
        tok.add_special_tokens([s for s in specials if s in table])
        tok.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True,
                                                    strip_accents=None, lowercase=True)
        tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
        tok.post_processor = processors.TemplateProcessing(
            single=f"{cls} $A {sep}", special_tokens=[(cls, table[cls]), (sep, table[sep])])
        tok.enable_truncation(max_length=max_length)
        self._tok = tok
        self.pad_id = table[pad]
        self.cls_id = table[cls]
        self.sep_id = table[sep]
        self.vocab_size = len(vocab)

    def encode_ragged(self, texts: list[str]) -> list[list[int]]:
        if not isinstance(texts, list):
            raise NotImplementedError
        return [e.ids for e in self._tok.encode_batch(texts)]

    def encode_packed(self, texts: list[str]) -> tuple[np.ndarray, np.ndarray]:
        """Strings -> ``(ids_packed, cu_seqlens)`` ready for ``owc_encode``."""
        return pack_ids(self.encode_ragged(texts))
