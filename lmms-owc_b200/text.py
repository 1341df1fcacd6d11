"""Text pipeline operator: drop-in for the reference's ``src/data/pipelines/text/_text.py``.

``encode_sentence_bert(batch, rank=None, **kwargs)`` keeps the reference's signature, kwargs,
mutation-and-return convention, output schema (list of lists of Python floats) and errors
(``_text.py:143-208``); underneath, tokenisation runs on the host and everything else in
``libowc_b200.so`` on the GPU. Like the reference it keeps the model and tokenizer as lazily
created module-level singletons (``_text.py:10-15, 156-163``).
"""

from __future__ import annotations

import numpy as np

from . import arch as _arch
from .engine import Engine
from .tokenizer import HostTokenizer

__all__ = ["encode_sentence_bert", "concept_extraction_spacy", "configure", "reset", "get_engine",
           "encode_unique", "PreExtractedConcepts", "MODEL_NAME"]

MODEL_NAME = "sentence-transformers/all-MiniLM-L6-v2"  # _text.py:161

sentence_bert_model: Engine | None = None          # same global names as the reference module
sentence_bert_processor: HostTokenizer | None = None
concept_extractor = None                           # callable(batch, **kwargs) -> batch; default spaCy
default_rank: int | None = None                    # GPU chosen when a caller passes rank=None (set by multi-process CLIs)


def configure(arch: _arch.EncoderArch, state_dict: dict, vocab: list[str], device: int | None = None,
              precision: str = "bf16", chunk_tokens: int = 0) -> Engine:
    """Install an encoder explicitly (weights + vocabulary in memory; used offline and by tests)."""
    global sentence_bert_model, sentence_bert_processor
    reset()
    kind = "bert" if arch.kind == _arch.KIND_BERT else "mpnet"
    sentence_bert_processor = HostTokenizer(vocab, kind=kind, max_length=min(512, arch.max_tokens))
    sentence_bert_model = Engine(arch, state_dict, device=device or 0, precision=precision,
                                 chunk_tokens=chunk_tokens)
    return sentence_bert_model


def reset() -> None:
    global sentence_bert_model, sentence_bert_processor
    if sentence_bert_model is not None:
        sentence_bert_model.close()
    sentence_bert_model = None
    sentence_bert_processor = None


def _lazy_load(rank: int | None) -> None:
    """First-call load, as in ``_text.py:158-163``: pretrained all-MiniLM-L6-v2 from the HF hub/cache."""
    import torch
    from transformers import AutoModel, AutoTokenizer

    tok = AutoTokenizer.from_pretrained(MODEL_NAME)
    model = AutoModel.from_pretrained(MODEL_NAME)
    vocab = [w for w, _ in sorted(tok.get_vocab().items(), key=lambda kv: kv[1])]
    cfg = model.config
    arch = _arch.EncoderArch(MODEL_NAME, _arch.KIND_BERT, cfg.num_hidden_layers, cfg.hidden_size,
                             cfg.num_attention_heads, cfg.intermediate_size, cfg.vocab_size,
                             cfg.max_position_embeddings, cfg.type_vocab_size, cfg.layer_norm_eps, pad_id=0)
    device = (rank or 0) % max(torch.cuda.device_count(), 1)   # _text.py:165-166
    configure(arch, model.state_dict(), vocab, device=device)


def get_engine(rank: int | None = None) -> tuple[Engine, HostTokenizer]:
    if sentence_bert_model is None:
        _lazy_load(default_rank if rank is None else rank)
    return sentence_bert_model, sentence_bert_processor


def encode_unique(texts: list[str], rank: int | None = None, out: str = "host"):
    """Embed each DISTINCT string once: returns ``(embeds [n_unique, H], inverse int32 [len(texts)])``.

    The reference encodes the class-name column once per row; the values are
    identical, so the host dedups and the paired score indexes into the unique table.
    """
    engine, tok = get_engine(rank)
    index = {t: i for i, t in enumerate(dict.fromkeys(texts))}     # first occurrence order
    inverse = np.fromiter(map(index.__getitem__, texts), dtype=np.int32, count=len(texts))
    ids, cu = tok.encode_packed(list(index))
    return engine.encode(ids, cu, out=out), inverse


def encode_sentence_bert(batch: dict, rank: int | None = None, **kwargs) -> dict:
    """Encode text with the Sentence-BERT model (drop-in for ``_text.py:143-208``).

    Args:
    ----
        batch (dict): The batch of data.
        rank (int, optional): The rank of the process (selects ``cuda:{rank % n}``). Defaults to None.
        **kwargs: ``input_column`` (default ``"text"``), ``output_column``
            (default ``f"{input_column}_sentence_bert_embeds"``).

    """
    input_column = kwargs.pop("input_column", "text")
    output_column = kwargs.pop("output_column", f"{input_column}_sentence_bert_embeds")
    engine, tok = get_engine(rank)

    if input_column not in batch:
        raise ValueError(f"{input_column} missing in dataset")

    if isinstance(batch[input_column], list):
        ids, cu = tok.encode_packed(batch[input_column])
        batch[output_column] = engine.encode(ids, cu, out="host").tolist()
    else:
        raise NotImplementedError

    return batch


class PreExtractedConcepts:
    """Concept extractor backed by concepts that were extracted elsewhere (SURVEY.md section 8(f), N4 fallback).

    spaCy + ``en_core_web_lg`` cannot be installed offline, so ``concept_semantic_similarity`` can take the concepts of
    every prediction from the samples file instead: a ``prediction_concepts`` column (what the reference's
    ``concept_extraction_spacy(..., skip_words=..., remove_prefix_words=True)`` returns for the scored prediction,
    ``_group.py:247-257``), or the ``last_resp_concepts`` column the reference itself writes back
    (``eval_metrics.py:93-104``: the same list with the prediction appended). Install an instance in
    ``owc_b200.text.concept_extractor``; it has the extractor's call signature and fails loudly for a prediction it has
    no concepts for.
    """

    def __init__(self, predictions, concepts) -> None:
        if len(predictions) != len(concepts):
            raise ValueError("one concept list per prediction")
        self.table: dict[str, list[str]] = {}
        for p, c in zip(predictions, concepts):
            c = [str(x) for x in c]
            if p in self.table and self.table[p] != c:
                raise ValueError(f"two different concept lists for the prediction {p!r}")
            self.table[p] = c

    def __call__(self, batch: dict, **kwargs) -> dict:
        input_column = kwargs.pop("input_column", "text")
        output_column = kwargs.pop("output_column", f"{input_column}_concepts")
        if input_column not in batch:
            raise ValueError(f"{input_column} missing in dataset")
        texts = batch[input_column]
        try:
            batch[output_column] = ([list(self.table[t]) for t in texts] if isinstance(texts, list)
                                    else list(self.table[texts]))
        except KeyError as exc:
            raise KeyError(f"no pre-extracted concepts for the prediction {exc.args[0]!r}") from None
        return batch


def concept_extraction_spacy(batch: dict, **kwargs) -> dict:
    """Concept extraction hook (reference: ``_text.py:18-140``, spaCy ``en_core_web_lg``).

    Noun-chunking is CPU NLP outside this build's scope. If a custom
    extractor was installed in ``concept_extractor`` it is used; otherwise spaCy is required and
    its absence raises (no silent substitute).
    """
    if concept_extractor is not None:
        return concept_extractor(batch, **kwargs)
    try:
        import spacy  # noqa: F401
    except ImportError as exc:  # pragma: no cover - environment dependent
        raise ImportError(
            "concept extraction needs spaCy + en_core_web_lg (as in the reference) or an extractor "
            "installed in owc_b200.text.concept_extractor") from exc
    return _spacy_extract(batch, **kwargs)


_spacy_model = None


def _spacy_extract(batch: dict, **kwargs) -> dict:  # pragma: no cover - spaCy absent offline
    import spacy

    global _spacy_model
    input_column = kwargs.pop("input_column", "text")
    output_column = kwargs.pop("output_column", f"{input_column}_concepts")
    skip_words = kwargs.pop("skip_words", [])
    remove_prefix_words = kwargs.pop("remove_prefix_words", False)
    if _spacy_model is None:
        _spacy_model = spacy.load("en_core_web_lg")
    if input_column not in batch:
        raise ValueError(f"{input_column} missing in dataset")
    prefixes = ["a", "an", "the", "his", "her", "its", "their"]

    def clean(text: str) -> str | None:
        concept = text.lower()
        if remove_prefix_words:
            for term in prefixes:
                if concept.startswith(term + " "):
                    concept = concept[len(term) + 1:]
                    break
            if concept in skip_words:
                return None
        return concept

    texts = batch[input_column]
    single = not isinstance(texts, list)
    docs = [_spacy_model(texts)] if single else _spacy_model.pipe(texts, batch_size=max(len(texts), 1))
    all_concepts = []
    for doc in docs:
        concepts: list[str] = []
        for chunk in doc.noun_chunks:
            c = clean(chunk.text)
            # the reference appends noun chunks only inside the remove_prefix_words branch
            if c is not None and remove_prefix_words:
                concepts.append(c)
        for ent in doc.ents:
            c = clean(ent.text)
            if c is not None and c not in concepts:
                concepts.append(c)
        all_concepts.append(concepts)
    batch[output_column] = all_concepts[0] if single else all_concepts
    return batch
