"""Host-side tokenisation and packing (tokenisation stays on the host, per the north star).

Mirrors what the reference gets from ``AutoTokenizer`` at ``src/data/pipelines/text/_text.py:193-195``
(``padding=True, truncation=True``): BERT normaliser (lower-case, clean text, CJK spacing), BERT
pre-tokeniser, WordPiece with ``##`` continuation and ``[UNK]``, ``[CLS] A [SEP]`` template (or
``<s> A </s>`` for MPNet) and truncation to the model's maximum length. Padding is NOT produced:
the device path takes ragged sequences packed back to back with a ``cu_seqlens`` offset array, and
every kernel is padding-invariant, so pad-to-longest would only add work.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["HostTokenizer", "NativeWordPiece", "pack_ids"]


def pack_ids(seqs: list[list[int]]) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate ragged id lists into ``(ids_packed int32 [M], cu_seqlens int32 [n+1])``."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    cu = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=cu[1:])
    if cu[-1] >= 2**31:
        raise ValueError("packed batch exceeds 2^31 tokens; split it")
    ids = np.fromiter((t for s in seqs for t in s), dtype=np.int32, count=int(cu[-1]))
    return ids, cu.astype(np.int32)


class NativeWordPiece:
    """The library's multi-threaded C++ WordPiece (``csrc/tokenizer.cpp``; ``owc_tokenize_packed``).

    Handles ASCII rows exactly as the BERT tokenizer does and writes the packed layout directly; rows it
    flags (non-ASCII bytes, literal special tokens) are left to the caller's fallback tokenizer."""

    def __init__(self, vocab: list[str], kind: str, max_length: int) -> None:
        from . import _cabi

        self._lib = _cabi.lib()
        enc = [w.encode("utf-8") for w in vocab]
        offs = np.zeros(len(enc) + 1, dtype=np.int64)
        np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.int64, count=len(enc)), out=offs[1:])
        blob = b"".join(enc)
        self._h = C.c_void_p()
        rc = self._lib.owc_tokenizer_create(blob, offs.ctypes.data, len(enc), 0 if kind == "bert" else 1, int(max_length),
                                            C.byref(self._h))
        if rc != 0:
            raise _cabi.OwcError(self._lib.owc_tokenizer_last_error().decode("utf-8", "replace"))

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            if self._h:
                self._lib.owc_tokenizer_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def tokenize(self, texts: list[str], threads: int = 0) -> tuple[np.ndarray, np.ndarray]:
        """``(ids of the unflagged rows back to back, lens int32 [n] with -1 for flagged rows)``."""
        from . import _cabi

        n = len(texts)
        joined = "\x00".join(texts)
        if joined.count("\x00") == n - 1:       # no NUL inside a row: one join + one encode instead of n of each
            blob = joined.encode("utf-8")
            offs_ptr, total = None, len(blob)
        else:
            enc = [t.encode("utf-8") for t in texts]
            offs = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.int64, count=n), out=offs[1:])
            blob = b"".join(enc)
            offs_ptr, total = offs.ctypes.data, int(offs[-1])
        cap = total + 2 * n + 1
        ids = np.empty(cap, dtype=np.int32)
        lens = np.empty(n, dtype=np.int32)
        n_ids = C.c_int64(0)
        rc = self._lib.owc_tokenize_packed(self._h, blob, offs_ptr, n, int(threads), ids.ctypes.data, cap,
                                           lens.ctypes.data, C.byref(n_ids))
        if rc != 0:
            raise _cabi.OwcError(self._lib.owc_tokenizer_last_error().decode("utf-8", "replace"))
        return ids[: n_ids.value], lens


    def tokenize_blob(self, blob: bytes, n: int, threads: int = 0) -> tuple[np.ndarray, np.ndarray]:
        """Same as :meth:`tokenize` for ``n`` rows already joined by NUL bytes (no trailing NUL, no NUL inside a row):
        the bytes go to the library as they are, no Python string is made."""
        from . import _cabi

        cap = len(blob) + 2 * n + 1
        ids = np.empty(cap, dtype=np.int32)
        lens = np.empty(n, dtype=np.int32)
        n_ids = C.c_int64(0)
        rc = self._lib.owc_tokenize_packed(self._h, blob, None, n, int(threads), ids.ctypes.data, cap, lens.ctypes.data,
                                           C.byref(n_ids))
        if rc != 0:
            raise _cabi.OwcError(self._lib.owc_tokenizer_last_error().decode("utf-8", "replace"))
        return ids[: n_ids.value], lens


class HostTokenizer:
    """WordPiece tokenizer over an in-memory vocabulary: the library's native C++ path for ASCII rows
    (``NativeWordPiece``), the ``tokenizers`` backend for everything else and as the cross-check."""

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
        self._native = None
        if os.environ.get("OWC_NATIVE_TOK", "1") != "0":
            try:
                self._native = NativeWordPiece(vocab, kind, max_length)
            except Exception:       # library not built yet (CPU-only checkout): the backend below does all rows
                self._native = None

    def encode_ragged(self, texts: list[str]) -> list[list[int]]:
        if not isinstance(texts, list):
            raise NotImplementedError
        return [e.ids for e in self._tok.encode_batch(texts)]

    def encode_packed(self, texts: list[str], native: bool | None = None) -> tuple[np.ndarray, np.ndarray]:
        """Strings -> ``(ids_packed, cu_seqlens)`` ready for ``owc_encode``."""
        if not isinstance(texts, list):
            raise NotImplementedError
        use_native = self._native is not None if native is None else (native and self._native is not None)
        if not use_native or not texts:
            return pack_ids(self.encode_ragged(texts))
        # 0 = every core; several ranks on one host share the cores (OWC_TOKENIZER_THREADS, set by the multi-rank CLI)
        ids, lens = self._native.tokenize(texts, threads=int(os.environ.get("OWC_TOKENIZER_THREADS", "0")))
        return self._finish_native(ids, lens, lambda rows: [texts[i] for i in rows])

    def encode_packed_blob(self, blob: bytes, n: int) -> tuple[np.ndarray, np.ndarray]:
        """:meth:`encode_packed` for ``n`` UTF-8 rows joined by NUL bytes (the JSONL scanner's output minus its last
        terminator): same ids, but only the rows the native path flags ever become Python strings."""
        if n == 0:
            return pack_ids([])
        if self._native is None:
            return self.encode_packed(blob.decode("utf-8").split("\0"))
        ids, lens = self._native.tokenize_blob(blob, n, threads=int(os.environ.get("OWC_TOKENIZER_THREADS", "0")))

        def rows_of(rows):
            parts = blob.split(b"\0")
            return [parts[i].decode("utf-8") for i in rows]

        return self._finish_native(ids, lens, rows_of)

    def _finish_native(self, ids, lens, texts_of) -> tuple[np.ndarray, np.ndarray]:
        """cu_seqlens from the native lengths; rows it flagged (-1) are tokenised by the backend and spliced in."""
        n = len(lens)
        flagged = np.nonzero(lens < 0)[0]
        if flagged.size == 0:
            cu = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(lens, out=cu[1:])
            if cu[-1] >= 2**31:
                raise ValueError("packed batch exceeds 2^31 tokens; split it")
            return ids, cu.astype(np.int32)
        # splice the rows the native path left to the backend (non-ASCII, literal special tokens) back in
        extra = self.encode_ragged(texts_of(flagged))
        lens = lens.astype(np.int64)
        lens[flagged] = [len(e) for e in extra]
        cu = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=cu[1:])
        if cu[-1] >= 2**31:
            raise ValueError("packed batch exceeds 2^31 tokens; split it")
        out = np.empty(int(cu[-1]), dtype=np.int32)
        is_extra = np.zeros(n, dtype=bool)
        is_extra[flagged] = True
        tok_extra = np.repeat(is_extra, lens)
        out[~tok_extra] = ids
        out[tok_extra] = np.fromiter((t for e in extra for t in e), dtype=np.int32, count=int(tok_extra.sum()))
        return out, cu.astype(np.int32)
