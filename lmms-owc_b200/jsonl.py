"""Native samples-JSONL reader / writer for the scoring CLI.

The reference parses each ``*_samples_*.jsonl`` with ``pd.read_json(..., lines=True)`` and rewrites the whole frame with
``df.to_json(..., lines=True, orient="records")`` once per metric (``eval_metrics.py:62``, ``:123``). At sweep scale
(130 files, 1.09 M rows) that is ~15 s of pandas work around ~0.5 s of device work. This module does the same job with
``owc_jsonl_*`` of the C ABI (``csrc/jsonl.cpp``): one pass that pulls the reference string and the scored prediction
string out of every line, and one pass that copies every line through verbatim while replacing / appending the score
members.

Contract of the rewrite: every line parses to the same record the reference's pandas rewrite parses to (scores equal
to the 10 decimals pandas writes), members the metric does not touch keep their exact bytes (pandas re-escapes ``/`` and
non-ASCII text and rounds every float of the row to 10 decimals; this writer is lossless there), new members are
appended in the order pandas appends columns. Files whose rows do not have the shapes below are reported as
unsupported and the caller uses the pandas path:

* ``target``: a string, or a list whose first element is a string (``ref[0]``, ``_group.py:518``);
* ``filtered_resps``: a string, a non-empty list ending in a string (``pred[-1]``, ``_group.py:519``), or a list of
  such lists (multi-round generation: the first inner list is taken, ``eval_metrics.py:66-68``).

Only the CLI uses this; the metric functions themselves are unchanged.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi

__all__ = ["JsonlError", "Unsupported", "PackedPairs", "PackedStrings", "read_packed_pairs", "read_pairs", "read_strings",
           "read_raw", "read_json_column", "render_columns", "write_columns", "has_members", "count_rows"]

OK, ERROR, UNSUPPORTED = 0, 1, 2
FIRST, LAST = 0, 1
FLOAT, INT, RAW = 0, 1, 2


class JsonlError(ValueError):
    """The file is not valid JSONL (the message names the line)."""


class Unsupported(Exception):
    """Valid JSONL, but with row shapes the native path does not cover: use the pandas path."""


def _raise(rc: int) -> None:
    msg = _cabi.lib().owc_jsonl_last_error().decode("utf-8", "replace")
    if rc == UNSUPPORTED:
        raise Unsupported(msg)
    raise JsonlError(msg)


def read_strings(raw: bytes, key: str, mode: int) -> list[str]:
    """One string per row from top-level member ``key`` (``mode``: ``FIRST`` = reference rule, ``LAST`` = scored
    prediction rule); ``owc_jsonl_extract``."""
    lib = _cabi.lib()
    out = np.empty(len(raw) + 1, np.uint8)
    out_len, n_rows = C.c_size_t(), C.c_size_t()
    rc = lib.owc_jsonl_extract(raw, len(raw), key.encode(), mode, out.ctypes.data, len(raw) + 1, C.byref(out_len),
                               C.byref(n_rows))
    if rc != OK:
        _raise(rc)
    if n_rows.value == 0:
        return []
    # one NUL after every row (rows containing NUL were reported as unsupported)
    return out[: out_len.value - 1].tobytes().decode("utf-8").split("\0")


def read_raw(raw: bytes, key: str) -> list[str]:
    """The raw JSON text of top-level member ``key`` of every row (``owc_jsonl_extract_raw``); :class:`Unsupported` when
    a row lacks it."""
    lib = _cabi.lib()
    cap = len(raw) + count_rows(raw) + 1
    out = np.empty(cap, np.uint8)
    out_len, n_rows = C.c_size_t(), C.c_size_t()
    rc = lib.owc_jsonl_extract_raw(raw, len(raw), key.encode(), out.ctypes.data, cap, C.byref(out_len), C.byref(n_rows))
    if rc != OK:
        _raise(rc)
    if n_rows.value == 0:
        return []
    return out[: out_len.value - 1].tobytes().decode("utf-8").split("\0")


def read_json_column(raw: bytes, key: str) -> list:
    """Member ``key`` of every row, parsed (list-valued columns such as pre-extracted ``prediction_concepts``)."""
    import json

    return [json.loads(t) for t in read_raw(raw, key)]


def count_rows(raw: bytes) -> int:
    n = C.c_size_t()
    rc = _cabi.lib().owc_jsonl_count_rows(raw, len(raw), C.byref(n))
    if rc != OK:
        _raise(rc)
    return n.value


def read_pairs(raw: bytes) -> list[tuple[str, str]]:
    """``[(reference string, scored prediction string)]`` per row: what ``_select`` of ``metrics/_group.py`` makes of
    the ``(target, filtered_resps)`` items the reference builds (``eval_metrics.py:62-70``). One pass over the bytes."""
    lib = _cabi.lib()
    cap = len(raw) + 1
    refs, preds = np.empty(cap, np.uint8), np.empty(cap, np.uint8)       # not zero-filled: pages are touched on write
    refs_len, preds_len, n_rows = C.c_size_t(), C.c_size_t(), C.c_size_t()
    rc = lib.owc_jsonl_extract_pairs(raw, len(raw), b"target", b"filtered_resps", refs.ctypes.data, cap, C.byref(refs_len),
                                     preds.ctypes.data, cap, C.byref(preds_len), C.byref(n_rows))
    if rc != OK:
        _raise(rc)
    if n_rows.value == 0:
        return []
    # one NUL after every row (rows containing NUL were reported as unsupported)
    return list(zip(refs[: refs_len.value - 1].tobytes().decode("utf-8").split("\0"),
                    preds[: preds_len.value - 1].tobytes().decode("utf-8").split("\0"), strict=True))


class PackedStrings:
    """``n`` strings as one UTF-8 blob, each followed by a NUL byte (the scanner's own output format)."""

    __slots__ = ("blob", "n")

    def __init__(self, blob: bytes, n: int) -> None:
        self.blob, self.n = blob, n

    def __len__(self) -> int:
        return self.n

    def tolist(self) -> list[str]:
        return self.blob[:-1].decode("utf-8").split("\0") if self.n else []

    def unique(self) -> tuple["PackedStrings", np.ndarray]:
        """``(distinct strings in first-occurrence order, inverse int32 [n])`` - ``owc_blob_unique``."""
        inverse = np.empty(self.n, dtype=np.int32)
        if self.n == 0:
            return PackedStrings(b"", 0), inverse
        out = np.empty(len(self.blob) + 1, np.uint8)
        out_len, n_unique = C.c_size_t(), C.c_size_t()
        rc = _cabi.lib().owc_blob_unique(self.blob, len(self.blob), self.n, inverse.ctypes.data, out.ctypes.data,
                                         len(self.blob) + 1, C.byref(out_len), C.byref(n_unique))
        if rc != OK:
            _raise(rc)
        return PackedStrings(out[: out_len.value].tobytes(), n_unique.value), inverse


class PackedPairs:
    """The ``(reference, prediction)`` items of one samples file in packed form. Stands in for the list of tuples the
    reference builds (``eval_metrics.py:70``) on the CLI's native path only: ``len()`` works, the group metrics pick
    the two columns up without touching a row, ``tolist()`` gives the ordinary items."""

    __slots__ = ("refs", "preds")

    def __init__(self, refs: PackedStrings, preds: PackedStrings) -> None:
        if len(refs) != len(preds):
            raise ValueError("one reference per prediction")
        self.refs, self.preds = refs, preds

    def __len__(self) -> int:
        return len(self.refs)

    def tolist(self) -> list[tuple[str, str]]:
        return list(zip(self.refs.tolist(), self.preds.tolist(), strict=True))


def read_packed_pairs(raw: bytes) -> PackedPairs:
    """:func:`read_pairs` without making a Python string per row."""
    lib = _cabi.lib()
    cap = len(raw) + 1
    refs, preds = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
    refs_len, preds_len, n_rows = C.c_size_t(), C.c_size_t(), C.c_size_t()
    rc = lib.owc_jsonl_extract_pairs(raw, len(raw), b"target", b"filtered_resps", refs.ctypes.data, cap, C.byref(refs_len),
                                     preds.ctypes.data, cap, C.byref(preds_len), C.byref(n_rows))
    if rc != OK:
        _raise(rc)
    n = n_rows.value
    return PackedPairs(PackedStrings(refs[: refs_len.value].tobytes(), n), PackedStrings(preds[: preds_len.value].tobytes(), n))


def has_members(raw: bytes, names: list[str]) -> bool:
    """True if every row already carries every member of ``names``."""
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    flag = C.c_int()
    rc = _cabi.lib().owc_jsonl_has_members(raw, len(raw), arr, len(names), C.byref(flag))
    if rc != OK:
        _raise(rc)
    return bool(flag.value)


def _is_list_column(values) -> bool:
    return len(values) > 0 and all(isinstance(v, (list, tuple)) for v in values)


def _column_kind(values) -> int:
    if _is_list_column(values):
        return RAW
    a = np.asarray(values)
    if a.dtype.kind in "iub":
        return INT
    if a.dtype.kind == "f":
        return FLOAT
    raise Unsupported(f"column of dtype {a.dtype} (numeric per-row columns and lists of strings / numbers are written natively)")


def _pandas_number(v: float) -> str:
    """A double as pandas' ``to_json`` writes it: fixed notation, 10 decimals, trailing zeros trimmed, NaN / inf -> null."""
    v = float(v)
    if not np.isfinite(v):
        return "null"
    if v == 0.0:
        return "0.0"
    t = "%.10f" % v
    t = t.rstrip("0")
    return t + "0" if t.endswith(".") else t


def _pandas_fragment(v) -> str:
    """One JSON value the way ``df.to_json`` renders it (ASCII-escaped strings with ``/`` escaped, numbers as above)."""
    import json

    if isinstance(v, (list, tuple)):
        return "[" + ",".join(_pandas_fragment(x) for x in v) + "]"
    if isinstance(v, str):
        return json.dumps(v, ensure_ascii=True).replace("/", "\\/")
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if v is None:
        return "null"
    return _pandas_number(v)


def _column_args(raw: bytes, columns: dict):
    names = list(columns)
    kind_list = [_column_kind(columns[n]) for n in names]
    kinds = (C.c_int32 * len(names))(*kind_list)
    n_rows = len(columns[names[0]]) if names else count_rows(raw)
    vals, blobs = [], []
    for n, k in zip(names, kind_list):
        if len(columns[n]) != n_rows:
            raise ValueError(f"columns of different lengths: {[len(columns[m]) for m in names]}")
        if k == RAW:
            vals.append(np.zeros(n_rows))
            blobs.append(("\0".join(_pandas_fragment(v) for v in columns[n]) + "\0").encode("utf-8"))
        else:
            vals.append(np.asarray(columns[n], dtype=np.float64).reshape(-1))
            blobs.append(None)
    table = np.ascontiguousarray(np.stack(vals)) if vals else np.zeros((0, n_rows))
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    raws = (C.c_char_p * len(names))(*blobs) if any(b is not None for b in blobs) else None
    return names, arr, kinds, table, n_rows, raws


def render_columns(raw: bytes, columns: dict) -> bytes:
    """The file's bytes with ``columns`` (name -> one number per row) replaced / appended on every line."""
    names, arr, kinds, table, n_rows, raws = _column_args(raw, columns)
    if raws is not None:
        raise Unsupported("list-valued columns are only written to files (write_columns)")
    if n_rows != count_rows(raw):
        raise ValueError(f"{count_rows(raw)} rows in the file, {n_rows} values per column")
    cap = len(raw) + n_rows * (sum(len(n.encode()) + 4 for n in names) + 40 * len(names) + 4) + 64
    out = np.empty(cap, np.uint8)
    out_len = C.c_size_t()
    rc = _cabi.lib().owc_jsonl_set_columns(raw, len(raw), arr, kinds, table.ctypes.data_as(C.c_void_p), len(names),
                                           n_rows, out.ctypes.data, cap, C.byref(out_len))
    if rc != OK:
        _raise(rc)
    return out[: out_len.value].tobytes()


def write_columns(path: str, raw: bytes, columns: dict) -> None:
    """Rewrite ``path`` with the columns added: the library streams the new text into a temporary file, which then
    replaces the samples file (a crash cannot truncate it)."""
    if not columns:
        return
    names, arr, kinds, table, n_rows, raws = _column_args(raw, columns)
    tmp = path + ".tmp"
    if raws is None:
        rc = _cabi.lib().owc_jsonl_write_columns(raw, len(raw), arr, kinds, table.ctypes.data_as(C.c_void_p), len(names),
                                                 n_rows, os.fsencode(tmp))
    else:
        rc = _cabi.lib().owc_jsonl_write_columns_raw(raw, len(raw), arr, kinds, table.ctypes.data_as(C.c_void_p), raws,
                                                     len(names), n_rows, os.fsencode(tmp))
    if rc != OK:
        _raise(rc)
    os.replace(tmp, path)
