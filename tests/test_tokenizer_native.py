"""The library's native C++ WordPiece (csrc/tokenizer.cpp, `owc_tokenize_packed`) against the `tokenizers` backend,
which is what the reference's AutoTokenizer call (_text.py:193-195) runs. CPU only: no CUDA is involved."""

import string

import numpy as np
import pytest

import owc_b200  # noqa: F401
from owc_b200 import synth
from owc_b200.tokenizer import HostTokenizer


@pytest.fixture(scope="module", params=["bert", "mpnet"])
def tok(request):
    vocab = synth.build_vocab(request.param, 30522 if request.param == "bert" else 30527)
    t = HostTokenizer(vocab, request.param)
    if t._native is None:
        pytest.skip("libowc_b200.so not built")
    t.words = synth.vocab_words(vocab)
    return t


def _same(tok, texts):
    a = tok.encode_packed(texts, native=False)
    b = tok.encode_packed(texts, native=True)
    assert np.array_equal(a[1], b[1]), "cu_seqlens differ"
    assert np.array_equal(a[0], b[0]), "ids differ"
    return b


def test_workload_strings_are_identical(tok):
    classes = synth.real_class_names("food101") + synth.real_class_names("caltech101") + synth.real_class_names("sun397")
    preds, _ = synth.make_predictions(20000, classes, tok.words, seed=5)   # > 4096 rows: the threaded path
    ids, cu = _same(tok, classes + preds)
    assert ids.dtype == np.int32 and cu.dtype == np.int32 and cu[0] == 0 and cu[-1] == ids.size


def test_ascii_edge_cases_are_identical(tok):
    rng = np.random.default_rng(0)
    alpha = list(string.ascii_letters + string.digits + string.punctuation + " \t\n\r\x0b\x0c\x00\x1f\x7f" + "   ")
    texts = ["".join(rng.choice(alpha, size=int(k))) for k in rng.integers(0, 90, size=4000)]
    texts += ["", " ", "\t\n", "a" * 100, "a" * 101, "x" * 300 + " y", "don't stop-me_now!", "A.B,C;D", "word " * 600,
              "UPPER lower MiXeD 123abc", "##ab ##", "e-mail@host.com", "[unk] [mask] <S>"]
    _same(tok, texts)


def test_flagged_rows_fall_back_and_splice(tok):
    specials = ["[CLS] hi", "hello [SEP]", "x [UNK] y", "<s>", "a </s> b", "[PAD]"]
    texts = ["plain row", "café au lait", "中文 text", "naïve", "zürich 12", *specials, "last plain row"] * 3
    ids, cu = _same(tok, texts)
    _, lens = tok._native.tokenize(texts)
    assert (lens[[1, 2, 3, 4]] == -1).all() and lens[0] > 0      # non-ASCII rows are handed to the backend
    assert (lens < 0).sum() >= 3 * 6                              # and so is every row with a literal special token


def test_truncation_and_specials(tok):
    texts = [" ".join(tok.words[:700]), "a"]
    ids, cu = _same(tok, texts)
    assert cu[1] - cu[0] == 512 and ids[0] == tok.cls_id and ids[cu[1] - 1] == tok.sep_id
    assert list(ids[cu[1]:cu[2]][[0, -1]]) == [tok.cls_id, tok.sep_id]


def test_empty_batch_and_bad_input(tok):
    ids, cu = tok.encode_packed([])
    assert ids.size == 0 and list(cu) == [0]
    with pytest.raises(NotImplementedError):
        tok.encode_packed("not a list")
