"""Synthetic workloads for the scoring path (no network: no real vocab, no real predictions).

Follows the measurement plan in SURVEY.md section 8(d): a deterministic WordPiece vocabulary of the
real size (so the embedding table has its real footprint), the reference's real class-name lists
where the config names a dataset, synthetic class names elsewhere, and label-like prediction
strings that are all distinct (so host-side dedup cannot inflate a throughput figure).
"""

from __future__ import annotations

import json
import os
import random
import string

__all__ = [
    "build_vocab",
    "vocab_words",
    "real_class_names",
    "make_class_names",
    "make_predictions",
    "REAL_CLASS_SETS",
]

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

# fgvc_aircraft (the tenth task) reads its 100 variant names from the downloaded dataset, not from the repository
REAL_CLASS_SETS = (
    "caltech101",
    "dtd",
    "eurosat",
    "flowers102",
    "food101",
    "oxford_pets",
    "stanford_cars",
    "sun397",
    "ucf101",
)


def _random_words(rng: random.Random, count: int, taken: set[str]) -> list[str]:
    words: list[str] = []
    letters = string.ascii_lowercase
    while len(words) < count:
        w = "".join(rng.choice(letters) for _ in range(rng.randint(2, 7)))
        if w not in taken:
            taken.add(w)
            words.append(w)
    return words


def build_vocab(kind: str = "bert", size: int = 30522, seed: int = 0) -> list[str]:
    """Deterministic WordPiece vocabulary (token at index i has id i).

    ``bert``: ``[PAD]``=0, ``[unused0..98]``, ``[UNK]``=100, ``[CLS]``=101, ``[SEP]``=102,
    ``[MASK]``=103 (the ids of the real BERT/MiniLM vocab), printable ASCII characters,
    ``##``+alphanumerics, then random lowercase words of 2-7 letters.
    ``mpnet``: ``<s>``=0, ``<pad>``=1, ``</s>``=2, ``<unk>``=3, then ``[UNK]`` (the tokenizer's
    default unknown token), the same character/word body, and ``<mask>`` last.
    """
    rng = random.Random(seed)
    chars = [chr(c) for c in range(33, 127)]
    cont = ["##" + c for c in string.ascii_lowercase + string.digits]
    if kind == "bert":
        head = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
        tail: list[str] = []
    elif kind == "mpnet":
        head = ["<s>", "<pad>", "</s>", "<unk>", "[UNK]"]
        tail = ["<mask>"]
    else:
        raise ValueError(f"unknown vocab kind {kind!r}")
    body = chars + cont
    taken = set(head) | set(body) | set(tail)
    n_words = size - len(head) - len(body) - len(tail)
    if n_words < 0:
        raise ValueError(f"vocab size {size} too small")
    words = _random_words(rng, n_words, taken)
    vocab = head + body + words + tail
    assert len(vocab) == size and len(set(vocab)) == size
    return vocab


def vocab_words(vocab: list[str]) -> list[str]:
    """The whole-word entries of a vocabulary (what the generators draw from)."""
    return [w for w in vocab if len(w) >= 2 and w.isalpha() and w.islower()]


def real_class_names(dataset: str) -> list[str]:
    """Class names of one of the reference's ten classification tasks, target-transformed.

    Extracted from the reference's ``src/data/tasks/_classification/<dataset>/assets/metadata.csv``
    (``class_name`` column, then that task's ``doc_to_target`` transform) by
    ``oracle/make_golden.py``; stored in ``assets/class_names.json``.
    """
    with open(os.path.join(_ASSETS, "class_names.json"), encoding="utf-8") as f:
        return list(json.load(f)[dataset])


def make_class_names(count: int, words: list[str], seed: int = 1234) -> list[str]:
    """``count`` distinct synthetic class names of 1-3 vocabulary words each."""
    rng = random.Random(seed)
    names: list[str] = []
    seen: set[str] = set()
    while len(names) < count:
        name = " ".join(rng.choice(words) for _ in range(rng.randint(1, 3)))
        if name not in seen:
            seen.add(name)
            names.append(name)
    return names


def make_predictions(
    count: int,
    class_names: list[str],
    words: list[str],
    seed: int = 1234,
    long: bool = False,
) -> tuple[list[str], list[int]]:
    """Distinct label-like prediction strings plus a uniform-random target class per row.

    With probability 0.5 a prediction is ``"<class name> <random word>"`` (a near hit), otherwise
    1-6 random vocabulary words. ``long=True`` draws 24-48 words instead (the 64-token bucket; the
    tokenizer truncates whatever exceeds the encoder's limit).
    """
    rng = random.Random(seed)
    preds: list[str] = []
    seen: set[str] = set()
    n_cls = len(class_names)
    while len(preds) < count:
        if long:
            p = " ".join(rng.choice(words) for _ in range(rng.randint(24, 48)))
        elif rng.random() < 0.5:
            p = class_names[rng.randrange(n_cls)] + " " + rng.choice(words)
        else:
            p = " ".join(rng.choice(words) for _ in range(rng.randint(1, 6)))
        if p not in seen:
            seen.add(p)
            preds.append(p)
    targets = [rng.randrange(n_cls) for _ in range(count)]
    return preds, targets
