"""N2: the native samples-JSONL scanner against pandas / ``json`` - the reader must hand the
metrics the strings the reference's ``pd.read_json`` + ``ref[0]`` / ``pred[-1]`` selection hands them, and the writer's
files must parse to the records of the reference's ``df.to_json(..., lines=True, orient="records")`` rewrite. CPU only."""

import io
import json

import numpy as np
import pytest

import owc_b200  # noqa: F401

WEIRD = ["plain", "a/b", "é ü 日本", 'quote " inside', "back\\slash", "new\nline\ttab", "😀 emoji", "", " lead",
         "x" * 300, "}{][,:", '\\"', "  sep", "\x7f\x01 ctl"]


def _rows(n=300):
    rows = []
    for i in range(n):
        w = WEIRD[i % len(WEIRD)] + str(i)
        resps = [w] if i % 3 else ["first", w]
        if i % 7 == 0:
            resps = [["r0", w], ["later round"]]
        rows.append({"doc_id": i, "doc": {"image": f"img_{i}.jpg", "nested": {"a": [1, 2, {"b": '}]"'}]}},
                     "target": "class/é %d" % (i % 100),
                     "arguments": {"0": 'What is "this"?', "1": {"max_new_tokens": 32, "temperature": 0.123456789012345}},
                     "resps": [["a photo of thing %d" % i]], "filtered_resps": resps, "doc_hash": "x" * 64,
                     "input": "What?", "f": 1 / 3, "n": None, "t": True})
    return rows


def _expected_pairs(raw: bytes):
    """The reference's item building (eval_metrics.py:62-70) + selection (_group.py:518-519), through pandas."""
    import pandas as pd

    df = pd.read_json(io.BytesIO(raw), lines=True)
    preds, refs = df["filtered_resps"].tolist(), df["target"].tolist()
    if isinstance(preds[0], list) and isinstance(preds[0][0], list):
        preds = [p[0] for p in preds]
    refs = [r[0] if isinstance(r, list) else r for r in refs]
    preds = [p[-1] if isinstance(p, list) else p for p in preds]
    return list(zip(refs, preds, strict=True)), df


@pytest.mark.parametrize("ascii_only", [False, True])
@pytest.mark.parametrize("sep", ["\n", "\n\n", "\r\n"])
def test_reader_matches_pandas_selection(ascii_only, sep):
    from owc_b200 import jsonl

    rows = [r for r in _rows() if not isinstance(r["filtered_resps"][0], list)]
    raw = sep.join(json.dumps(r, ensure_ascii=ascii_only) for r in rows).encode() + sep.encode()
    exp, _ = _expected_pairs(raw)
    assert jsonl.read_pairs(raw) == exp
    assert jsonl.count_rows(raw) == len(rows)


def test_reader_multi_round_and_list_targets():
    from owc_b200 import jsonl

    def dump(rows):
        return ("\n".join(json.dumps(r) for r in rows)).encode()            # no trailing newline

    # multi-round is decided once per file, on the first row (eval_metrics.py:66-68), then `prediction[0]` for all
    multi = [{"target": ["cat", "feline"], "filtered_resps": [["round one", "a cat"], ["round two"]]},
             {"target": "dog", "filtered_resps": [["only"]]},
             {"target": "cow", "filtered_resps": ["flat in a multi-round file", "ignored"]}]
    assert jsonl.read_pairs(dump(multi)) == _expected_pairs(dump(multi))[0] == [
        ("cat", "a cat"), ("dog", "only"), ("cow", "flat in a multi-round file")]
    flat = [{"filtered_resps": "bare string", "target": "x", "filtered_resps_extra": ["no"]},
            {"target": "dup", "filtered_resps": ["old"], "pad": 1, "filtered_resps ": ["key with space"]},
            {"target": "mix", "filtered_resps": [["nested but not first row"], "last wins"]}]
    assert jsonl.read_pairs(dump(flat)) == _expected_pairs(dump(flat))[0] == [
        ("x", "bare string"), ("dup", "old"), ("mix", "last wins")]
    with pytest.raises(jsonl.Unsupported):                 # the reference would hand a list to the tokenizer
        jsonl.read_pairs(dump(flat + [{"target": "t", "filtered_resps": [["a", "b"]]}]))
    with pytest.raises(jsonl.Unsupported):                 # ... or index a string by character
        jsonl.read_pairs(dump(multi + [{"target": "t", "filtered_resps": "bare"}]))
    # a repeated member: the last one wins, as in json.loads
    assert jsonl.read_pairs(b'{"target":"a","filtered_resps":["p"],"target":"b"}') == [("b", "p")]
    assert jsonl.read_pairs(b"") == [] and jsonl.read_pairs(b"\n \n") == []


@pytest.mark.parametrize("line", [
    b'{"target":"a","filtered_resps":[]}',                     # pred[-1] of an empty list: IndexError in the reference
    b'{"target":"a"}',                                          # member missing
    b'{"target":3,"filtered_resps":["p"]}',                    # not a string
    b'{"target":"a","filtered_resps":["p",4]}',
    b'{"target":"a","filtered_resps":[[["deep"]]]}',
    b'{"target":"a\\u0000b","filtered_resps":["p"]}',          # NUL cannot travel in the NUL-separated blob
    b'{"target":"a","filtered_resps":["\\ud83d"]}',            # lone surrogate
])
def test_reader_reports_unsupported_shapes(line):
    from owc_b200 import jsonl

    with pytest.raises(jsonl.Unsupported):
        jsonl.read_pairs(b'{"target":"ok","filtered_resps":["fine"]}\n' + line + b"\n")


@pytest.mark.parametrize("line,where", [
    (b'{"target":"a","filtered_resps":["p"]', "line 2"),
    (b'["not","an","object"]', "line 2"),
    (b'{"target":"a" "filtered_resps":["p"]}', "line 2"),
    (b'{"target":"unterminated}', "line 2"),
    (b'{"target":"a","filtered_resps":["\\q"]}', "line 2"),
])
def test_reader_rejects_malformed_lines_with_the_line_number(line, where):
    from owc_b200 import jsonl

    with pytest.raises(jsonl.JsonlError, match=where):
        jsonl.read_pairs(b'{"target":"ok","filtered_resps":["fine"]}\n' + line + b"\n")


def test_single_member_reader_agrees_with_the_pair_reader():
    from owc_b200 import jsonl

    rows = [r for r in _rows(60) if not isinstance(r["filtered_resps"][0], list)]
    raw = "\n".join(json.dumps(r) for r in rows).encode()
    pairs = jsonl.read_pairs(raw)
    assert jsonl.read_strings(raw, "target", jsonl.FIRST) == [p[0] for p in pairs]
    assert jsonl.read_strings(raw, "filtered_resps", jsonl.LAST) == [p[1] for p in pairs]
    assert jsonl.read_strings(raw, "input", jsonl.FIRST) == ["What?"] * len(rows)
    with pytest.raises(jsonl.Unsupported):
        jsonl.read_strings(raw, "doc_id", jsonl.FIRST)            # a number, not a string
    with pytest.raises(jsonl.JsonlError):
        jsonl.read_strings(raw, "target", 7)


def test_reader_decodes_escapes_like_json():
    from owc_b200 import jsonl

    texts = ["é中\U0001f600", "tab\there", 'q"q', "sl/ash", "\\u0041 literal", "\b\f\r", "mixed \U0001f680 ü x"]
    for ensure_ascii in (True, False):
        raw = "\n".join(json.dumps({"target": t, "filtered_resps": [t[::-1]]}, ensure_ascii=ensure_ascii) for t in texts).encode()
        assert jsonl.read_pairs(raw) == [(t, t[::-1]) for t in texts]
    assert jsonl.read_pairs(b'{"target":"\\/\\u00E9","filtered_resps":["\\uD83D\\uDE00"]}') == [("/é", "\U0001f600")]


def test_writer_matches_pandas_rewrite():
    """Same records as `df[col] = values; df.to_json(lines=True, orient="records")`; the score members are written
    with pandas' number format; every other member keeps its bytes."""
    import pandas as pd

    from owc_b200 import jsonl

    rows = [r for r in _rows() if not isinstance(r["filtered_resps"][0], list)]
    raw = "\n".join(json.dumps(r, ensure_ascii=False) for r in rows).encode() + b"\n"
    n = len(rows)
    rng = np.random.default_rng(0)
    scores = rng.uniform(-1, 1, n).astype(np.float32)
    scores[:6] = [0.0, 1.0, -1.0, 0.5, 1e-11, -1e-12]
    cols = {"semantic_similarity": scores.tolist(), "semantic_similarity@0.5": (scores >= 0.5).astype(np.int32).tolist(),
            "mean_average_semantic_similarity": (scores.astype(np.float64) / 3).tolist()}
    out = jsonl.render_columns(raw, cols)
    df = pd.read_json(io.BytesIO(raw), lines=True)
    for k, v in cols.items():
        df[k] = v
    ref_lines = df.to_json(lines=True, orient="records").splitlines()
    got_lines = out.decode().split("\n")[:-1]                  # not splitlines(): U+2028 inside a string is data
    assert len(got_lines) == len(ref_lines) == n
    src_lines = raw.decode().split("\n")[:-1]
    for src, got, ref in zip(src_lines, got_lines, ref_lines, strict=True):
        a, b = json.loads(got), json.loads(ref)
        assert list(a) == list(b)
        tail_a, tail_b = got[got.index('"semantic_similarity":'):], ref[ref.index('"semantic_similarity":'):]
        assert tail_a == tail_b                                  # pandas' "%.10f"-trimmed floats, plain ints
        assert got.startswith(src[:-1] + ",")                   # untouched members: verbatim
        for k in a:
            if k not in cols and k != "arguments":               # pandas rounds nested floats to 10 decimals
                assert a[k] == pytest.approx(b[k]) if isinstance(a[k], float) else a[k] == b[k]
    assert jsonl.render_columns(out, cols) == out               # idempotent: members are replaced, not duplicated
    assert jsonl.has_members(out, list(cols)) and not jsonl.has_members(raw, list(cols))
    assert not jsonl.has_members(out, ["semantic_similarity", "absent"])


def test_writer_replaces_existing_members_and_handles_nan():
    from owc_b200 import jsonl

    raw = b'{"a":1,"semantic_similarity":0.25,"b":"x"}\n\n{"semantic_similarity":null , "a":2,"b":"y"}\n{}\n'
    out = jsonl.render_columns(raw, {"semantic_similarity": [0.75, float("nan"), 2.0]})
    assert out == b'{"a":1,"b":"x","semantic_similarity":0.75}\n{ "a":2,"b":"y","semantic_similarity":null}\n{"semantic_similarity":2.0}\n'
    with pytest.raises(jsonl.JsonlError, match="more rows than values"):
        jsonl._cabi.lib()  # make sure the library is loaded before calling into it with bad sizes
        import ctypes as C

        names = (C.c_char_p * 1)(b"s")
        kinds = (C.c_int32 * 1)(0)
        vals = np.zeros(2)
        buf = C.create_string_buffer(1024)
        n = C.c_size_t()
        rc = jsonl._cabi.lib().owc_jsonl_set_columns(raw, len(raw), names, kinds, vals.ctypes.data_as(C.c_void_p), 1, 2,
                                                     buf, 1024, C.byref(n))
        assert rc == jsonl.ERROR
        jsonl._raise(rc)
    with pytest.raises(ValueError):
        jsonl.render_columns(raw, {"s": [1.0, 2.0]})
    with pytest.raises(jsonl.Unsupported):
        jsonl.render_columns(raw, {"concepts": [["a"], ["b"], ["c"]]})


def test_fuzz_against_json():
    """Seeded random records (nested containers, escapes, odd spacing): reader == json.loads + selection; writer's
    lines parse to the source record plus the new members."""
    import random

    from owc_b200 import jsonl

    rnd = random.Random(7)
    alphabet = list('ab {}[]",:\\/\n\té中\U0001f600')

    def rstr():
        return "".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 12)))

    def rval(depth=0):
        k = rnd.randrange(7 if depth < 3 else 5)
        if k == 0:
            return rstr()
        if k == 1:
            return rnd.randrange(-1000, 1000)
        if k == 2:
            return rnd.uniform(-5, 5)
        if k == 3:
            return rnd.choice([True, False, None])
        if k == 4:
            return rstr()
        if k == 5:
            return [rval(depth + 1) for _ in range(rnd.randrange(0, 4))]
        return {rstr(): rval(depth + 1) for _ in range(rnd.randrange(0, 4))}

    recs, lines = [], []
    for i in range(400):
        rec = {rstr() + "k": rval() for _ in range(rnd.randrange(0, 5))}
        pred = [rstr() for _ in range(rnd.randrange(1, 4))]
        rec["filtered_resps"] = [pred, ["x"]] if i % 5 == 0 else pred
        rec["target"] = [rstr(), rstr()] if i % 11 == 0 else rstr()
        items = list(rec.items())
        rnd.shuffle(items)
        rec = dict(items)
        recs.append(rec)
        seps = rnd.choice([(",", ":"), (", ", ": "), (" ,  ", " :\t")])
        lines.append(" " * rnd.randrange(3) + json.dumps(rec, ensure_ascii=bool(i % 2), separators=seps) + " " * rnd.randrange(3))
    raw = "\n".join(lines).encode()
    multi = isinstance(recs[0]["filtered_resps"][0], list)             # decided on the first row, applied to all
    exp = []
    for r in recs:
        pred = r["filtered_resps"][0] if multi else r["filtered_resps"]
        exp.append((r["target"][0] if isinstance(r["target"], list) else r["target"],
                    pred[-1] if isinstance(pred, list) else pred))
    assert multi
    assert jsonl.read_pairs(raw) == exp
    vals = [rnd.uniform(-1, 1) for _ in recs]
    out = jsonl.render_columns(raw, {"score": vals, "hit": [int(v > 0) for v in vals]})
    for line, rec, v in zip(out.decode().split("\n")[:-1], recs, vals, strict=True):
        got = json.loads(line)
        assert got.pop("score") == pytest.approx(v, abs=1e-10) and got.pop("hit") == int(v > 0)
        assert got == rec


def test_cli_native_io_scores_like_the_pandas_path(minilm, monkeypatch, tmp_path):
    """`--io native` (serial and threaded) returns the pandas path's outputs and writes files that parse to the same
    records; a file the scanner does not cover silently takes the pandas path."""
    import shutil

    import pandas as pd

    from owc_b200 import eval_metrics as EM
    from owc_b200 import text as T
    from owc_b200.tokenizer import HostTokenizer
    from test_host import OracleEngine

    monkeypatch.setattr(T, "sentence_bert_model", OracleEngine(minilm))
    monkeypatch.setattr(T, "sentence_bert_processor", HostTokenizer(minilm.vocab, "bert"))
    b = minilm
    cls = [str(s) for s in b.gold["class_names"]]
    preds = [str(s) for s in b.gold["preds"]]
    tg = b.gold["targets"]
    root = tmp_path / "a" / "logs"
    for m in range(3):
        run = root / "caltech101" / f"model-{m}"
        run.mkdir(parents=True)
        with open(run / "r_samples_caltech101.jsonl", "w", encoding="utf-8") as f:   # the tracker's writer: json.dumps per line
            for i in range(20 * m, 20 * m + 30):
                resps = [[("first round"), preds[i]]] if m == 1 else [preds[i]]
                f.write(json.dumps({"doc_id": i, "target": cls[tg[i]], "filtered_resps": resps, "note": "a/b é",
                                    "arguments": {"0": "q"}}, ensure_ascii=False) + "\n")
    odd = root / "caltech101" / "model-odd"
    odd.mkdir(parents=True)
    with open(odd / "r_samples_caltech101.jsonl", "w", encoding="utf-8") as f:
        for i in range(10):
            f.write(json.dumps({"doc_id": i, "target": cls[tg[i]], "filtered_resps": [preds[i]], "x\\u": "\x00"}) + "\n")
        f.write(json.dumps({"doc_id": 10, "target": "nul \x00 inside", "filtered_resps": [preds[10]]}) + "\n")
    for twin in ("b", "c"):
        shutil.copytree(tmp_path / "a", tmp_path / twin)
    metrics = "semantic_similarity,mean_average_semantic_similarity"
    ref = EM.main(EM.build_parser().parse_args(["-i", str(root), "-m", metrics]))
    nat = EM.main(EM.build_parser().parse_args(["-i", str(tmp_path / "b" / "logs"), "-m", metrics, "--io", "native"]))
    thr = EM.main(EM.build_parser().parse_args(["-i", str(tmp_path / "c" / "logs"), "-m", metrics, "--io", "native",
                                                "--jobs", "3"]))
    assert ref == nat == thr
    for f in sorted(root.rglob("*.jsonl")):
        want = pd.read_json(f, lines=True)
        for twin in ("b", "c"):
            g = tmp_path / twin / "logs" / f.relative_to(root)
            got = pd.read_json(g, lines=True)
            assert sorted(got.columns) == sorted(want.columns)
            pd.testing.assert_frame_equal(got[want.columns], want, check_exact=False, atol=1e-9, rtol=0)
            if "odd" not in str(f):                                 # native rewrite: source bytes kept
                assert b'"note": "a/b \xc3\xa9"' in g.read_bytes()
    # the concept metric is eligible too: with a pre-extracted concepts column it stays native (list columns are written
    # as raw fragments), without one the file takes the pandas path (test_cli_concepts_from_the_samples_file)
    assert EM._native_eligible(["concept_semantic_similarity"])
    with pytest.raises(ValueError):
        EM.score_files([], ["semantic_similarity"], io="arrow")


def test_cli_concepts_from_the_samples_file(minilm, monkeypatch, tmp_path):
    """N4 fallback: `concept_semantic_similarity` end to end WITHOUT a concept extractor, from a
    `prediction_concepts` column of the samples file, and re-scoring from the `last_resp_concepts` column the metric
    writes back - pandas and native I/O write the same records, equal to the run with the extractor installed (which is
    pinned to the reference's own output in tests/test_host.py)."""
    import shutil

    import pandas as pd

    from oracle import metrics_ref
    from owc_b200 import eval_metrics as EM
    from owc_b200 import text as T
    from owc_b200.tokenizer import HostTokenizer
    from test_host import OracleEngine

    monkeypatch.setattr(T, "sentence_bert_model", OracleEngine(minilm))
    monkeypatch.setattr(T, "sentence_bert_processor", HostTokenizer(minilm.vocab, "bert"))
    b = minilm
    cls = [str(s) for s in b.gold["class_names"]]
    cpreds = [str(s) for s in b.gold["concept_preds"]]
    tg = b.gold["targets"]
    skip = [w for cat in __import__("owc_b200.metrics._group", fromlist=["x"])._SKIP_WORDS_GROUPS.values() for w in cat]
    extracted = metrics_ref.simple_concepts({"prediction": cpreds}, input_column="prediction", output_column="c",
                                            skip_words=skip, remove_prefix_words=True)["c"]

    def write(root, with_column):
        run = root / "caltech101" / "model-0"
        run.mkdir(parents=True)
        with open(run / "r_samples_caltech101.jsonl", "w", encoding="utf-8") as f:
            for i, p in enumerate(cpreds):
                row = {"doc_id": i, "target": cls[tg[i]], "filtered_resps": [p], "note": "a/b é"}
                if with_column:
                    row["prediction_concepts"] = extracted[i]
                f.write(json.dumps(row, ensure_ascii=False) + "\n")
        return run / "r_samples_caltech101.jsonl"

    args = lambda root, *extra: EM.build_parser().parse_args(["-i", str(root), "-m", "concept_semantic_similarity", *extra])  # noqa: E731
    # 1. the extractor installed, no column (what the existing tests pin to the reference)
    f_ext = write(tmp_path / "ext" / "logs", False)
    monkeypatch.setattr(T, "concept_extractor", metrics_ref.simple_concepts)
    want_out = EM.main(args(tmp_path / "ext" / "logs"))
    want = pd.read_json(f_ext, lines=True)
    monkeypatch.setattr(T, "concept_extractor", None)
    # 2. no extractor at all: the column feeds the metric, through pandas and through the native scanner
    f_pd = write(tmp_path / "pd" / "logs", True)
    f_nat = write(tmp_path / "nat" / "logs", True)
    assert EM.main(args(tmp_path / "pd" / "logs")) == want_out
    assert EM.main(args(tmp_path / "nat" / "logs", "--io", "native")) == want_out
    assert T.concept_extractor is None                              # restored after every file
    cols = ["concept_semantic_similarity", "last_resp_concepts", "last_resp_concepts_similarities"]
    for f in (f_pd, f_nat):
        got = pd.read_json(f, lines=True)
        assert got["last_resp_concepts"].tolist() == want["last_resp_concepts"].tolist()
        assert np.abs(got["concept_semantic_similarity"] - want["concept_semantic_similarity"]).max() < 1e-9
        for x, y in zip(got["last_resp_concepts_similarities"], want["last_resp_concepts_similarities"], strict=True):
            assert np.abs(np.asarray(x) - np.asarray(y)).max() < 1e-9
        assert got["prediction_concepts"].tolist() == extracted
    assert b'"note": "a/b \xc3\xa9"' in f_nat.read_bytes()          # native rewrite keeps untouched members' bytes
    # byte-level: the native writer formats the list columns exactly as pandas does
    line_pd = f_pd.read_bytes().splitlines()[0].decode()
    line_nat = f_nat.read_bytes().splitlines()[0].decode()
    for c in cols:
        frag = line_pd[line_pd.index(f'"{c}":'):]
        frag = frag[:frag.index(',"', 1) if ',"' in frag[1:] else len(frag) - 1] if c != cols[-1] else frag.rstrip("}")
        assert frag in line_nat, c
    # 3. re-scoring a file the metric (or the reference) has already written: `last_resp_concepts` is the source
    f_re = tmp_path / "re" / "logs" / "caltech101" / "model-0"
    f_re.mkdir(parents=True)
    df = want.drop(columns=["concept_semantic_similarity", "last_resp_concepts_similarities"])
    df.to_json(f_re / "r_samples_caltech101.jsonl", lines=True, orient="records")
    shutil.copytree(tmp_path / "re", tmp_path / "re_nat")
    assert EM.main(args(tmp_path / "re" / "logs")) == want_out
    assert EM.main(args(tmp_path / "re_nat" / "logs", "--io", "native")) == want_out
    again = pd.read_json(f_re / "r_samples_caltech101.jsonl", lines=True)
    assert again["last_resp_concepts"].tolist() == want["last_resp_concepts"].tolist()
    # 4. nothing to extract with and nothing extracted: the reference's requirement (spaCy) surfaces, nothing is made up
    write(tmp_path / "none" / "logs", False)
    with pytest.raises(ImportError):
        EM.main(args(tmp_path / "none" / "logs"))
    with pytest.raises(KeyError):
        EM.main(args(tmp_path / "none" / "logs", "--concepts-column", "my_concepts"))
    bad = T.PreExtractedConcepts(["a"], [["x"]])
    with pytest.raises(KeyError, match="no pre-extracted concepts"):
        bad({"prediction": ["b"]}, input_column="prediction")
    with pytest.raises(ValueError):
        T.PreExtractedConcepts(["a", "a"], [["x"], ["y"]])


def test_scanner_is_strict_about_control_bytes_and_trailing_data():
    """A raw control byte inside a string the metric uses, or bytes after the closing brace, are not JSON: pandas raises
    on both, so the native path must not score such a file (a raw NUL used to split a row in two)."""
    import pandas as pd
    from io import BytesIO

    from owc_b200 import jsonl

    good = b'{"target": "a", "filtered_resps": ["b"]}\n'
    assert jsonl.read_pairs(good) == [("a", "b")]
    for ctl in (b"\x00", b"\x01", b"\x1f", b"\n"[:0] + b"\t"):
        raw = b'{"target": "a' + ctl + b'b", "filtered_resps": ["c"]}\n' + good
        with pytest.raises(jsonl.Unsupported):
            jsonl.read_pairs(raw)                      # -> the CLI falls back to pandas ...
        with pytest.raises(jsonl.Unsupported):
            jsonl.read_packed_pairs(raw)
        try:                                           # ... whose verdict (an error for NUL, leniency for some others)
            pd.read_json(BytesIO(raw), lines=True)     # is then exactly the reference's verdict
        except ValueError:
            pass
    for junk in (b'{"target": "a", "filtered_resps": ["b"]} junk\n', b'{"target": "a", "filtered_resps": ["b"]}{}\n', b"{} x\n"):
        with pytest.raises(jsonl.JsonlError, match="trailing data"):
            jsonl.read_pairs(junk + good)
        with pytest.raises(jsonl.JsonlError, match="trailing data"):
            jsonl.render_columns(junk, {"s": [1.0]})
        with pytest.raises(ValueError):
            pd.read_json(BytesIO(junk), lines=True)
    assert jsonl.read_pairs(b'  {"target": "a", "filtered_resps": ["b"]}  \r\n') == [("a", "b")]   # surrounding whitespace is fine


def test_packed_route_equals_the_string_route():
    """The CLI's packed columns (scanner blob -> native dedup -> native tokenizer) carry exactly what the string route
    (read_pairs -> text.encode_unique -> encode_packed) carries: same distinct strings, same inverse, same ids."""
    from owc_b200 import jsonl, synth
    from owc_b200.tokenizer import HostTokenizer

    vocab = synth.build_vocab("bert", 30522)
    tok = HostTokenizer(vocab, "bert")
    words = synth.vocab_words(vocab)
    texts = [" ".join(words[(7 * i + j) % 997] for j in range(1 + i % 5)) for i in range(400)]
    texts += ["", "café au lait", "a [SEP] b", "dup", "dup", "日本 語", "", "x" * 700]      # flagged rows, repeats, empties
    rows = [{"target": texts[(3 * i) % 17], "filtered_resps": [texts[i]]} for i in range(len(texts))]
    raw = "\n".join(json.dumps(r, ensure_ascii=bool(i % 2)) for i, r in enumerate(rows)).encode()
    packed = jsonl.read_packed_pairs(raw)
    assert len(packed) == len(rows) and packed.tolist() == jsonl.read_pairs(raw)
    for col, strings in ((packed.refs, [r["target"] for r in rows]), (packed.preds, texts)):
        uniq, inverse = col.unique()
        first_seen = list(dict.fromkeys(strings))
        assert uniq.tolist() == first_seen and uniq.n == len(first_seen)
        assert [first_seen[i] for i in inverse] == strings
        ids_a, cu_a = tok.encode_packed_blob(uniq.blob[:-1], uniq.n)
        ids_b, cu_b = tok.encode_packed(first_seen)
        assert np.array_equal(cu_a, cu_b) and np.array_equal(ids_a, ids_b)
    empty = jsonl.PackedStrings(b"", 0)
    assert empty.tolist() == [] and empty.unique()[0].n == 0 and tok.encode_packed_blob(b"", 0)[1].tolist() == [0]
