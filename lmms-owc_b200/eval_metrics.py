"""Offline metric re-scoring CLI: drop-in for the reference's ``eval_metrics.py`` on the similarity path.

Same flags (``-i/--input``, ``-m/--metrics``, ``--seed``, ``--log-level``), same file discovery
(``**/*_samples_*.jsonl`` under ``logs/.../{task_name}/{model_name}/``), same item building
(``(target, filtered_resps)``, multi-round ``[[...]]`` unwrapped), same JSONL columns written back and
the same stdout blocks as ``eval_metrics.py:19-171``. Only the three similarity metrics are available
(the rest of the reference's metric zoo is outside this build); unknown metric names raise ``KeyError``
exactly like the reference's registry lookup. Differences, all additive:

* the sample file is rewritten atomically (temp file + ``os.replace``) instead of in place;
* ``--jobs N`` parses / rewrites files in N workers, ``--skip-scored`` skips files that already carry the columns;
* launched under ``torchrun`` (``WORLD_SIZE`` > 1) the files are dealt out to the ranks by size, every rank scores its
  own files on its own GPU and rank 0 prints the report (the reference has no multi-GPU mode for this path: its
  ``datasets.map`` calls at ``_group.py:521-535`` pass neither ``with_rank`` nor ``num_proc``, so it runs on one device);
* ``--io native`` reads and rewrites the files with the library's streaming JSONL scanner (``jsonl.py``) instead of
  pandas: same records, untouched members keep their bytes, ~8x less host time per file;
* ``concept_semantic_similarity`` needs spaCy's ``en_core_web_lg`` in the reference (``_text.py:18-140``). When the
  samples file carries pre-extracted concepts - a ``prediction_concepts`` column, or the ``last_resp_concepts`` column
  the reference itself writes back - they are used instead (``--concepts-column``), so the
  metric runs end to end on a machine without spaCy.

    python -m owc_b200.eval_metrics -i logs/schedule -m semantic_similarity
"""

from __future__ import annotations

import logging
import os
import random
from argparse import ArgumentParser, Namespace
from pathlib import Path

import numpy as np

from .metrics import get_metric_info

log = logging.getLogger("owc_b200.eval_metrics")

METRICS_SAVING_INTERMEDIATE_VALUES = [
    "concept_semantic_similarity",
    "mean_average_semantic_similarity",
    "semantic_similarity",
]


def find_input_files(pattern: str) -> list[str]:
    """``eval_metrics.py:36-44``: resolve a glob / directory / file into sorted sample files."""
    paths = sorted(Path().glob(pattern)) if "*" in pattern else [Path(pattern)]
    per_path = [list(p.glob("**/*_samples_*.jsonl")) if p.is_dir() else [p] for p in paths]
    return sorted(map(str, sum(per_path, [])))


def _load_items(input_file: str):
    """Parse one samples file into ``(df, items)`` (``eval_metrics.py:62-70``)."""
    import pandas as pd

    df = pd.read_json(input_file, lines=True)
    predictions = df["filtered_resps"].tolist()
    references = df["target"].tolist()
    # multi-turn generation: drop the inner list (eval_metrics.py:66-68)
    if predictions and isinstance(predictions[0], list) and predictions[0] and isinstance(predictions[0][0], list):
        predictions = [p[0] for p in predictions]
    return df, list(zip(references, predictions, strict=True))


# ---- pre-extracted concepts (SURVEY.md section 8(f), N4 fallback) ---------------------------------------------------
CONCEPT_COLUMNS = ("prediction_concepts", "last_resp_concepts")


def _scored_prediction(pred):
    return pred[-1] if isinstance(pred, list) else pred          # _group.py:519


def _concept_source(available, wanted: str) -> str | None:
    """Which column of the file holds pre-extracted concepts: ``wanted`` = "auto" (the first of CONCEPT_COLUMNS that is
    present), "none" (always run the extractor) or a column name (which must exist)."""
    if wanted == "none":
        return None
    if wanted == "auto":
        return next((c for c in CONCEPT_COLUMNS if c in available), None)
    if wanted not in available:
        raise KeyError(f"--concepts-column {wanted!r}: the samples file has no such column")
    return wanted


def _concept_extractor_from(column: str, items: list, concept_lists: list):
    """A :class:`owc_b200.text.PreExtractedConcepts` for one file. ``last_resp_concepts`` is what the reference writes
    back (``eval_metrics.py:93-104``): the extractor's concepts with the scored prediction appended (``_group.py:260-262``),
    so the last element is dropped again."""
    from .text import PreExtractedConcepts

    preds = [_scored_prediction(p) for _, p in items]
    if column == "last_resp_concepts":
        for p, c in zip(preds, concept_lists, strict=True):
            if not c or c[-1] != p:
                raise ValueError(f"`last_resp_concepts` of the prediction {p!r} does not end with the prediction itself")
        concept_lists = [list(c[:-1]) for c in concept_lists]
    return PreExtractedConcepts(preds, concept_lists)


class _use_extractor:
    """Context manager: install a concept extractor in ``owc_b200.text`` for the duration of one file."""

    def __init__(self, extractor) -> None:
        self.extractor = extractor

    def __enter__(self):
        from . import text as _text

        self.saved = _text.concept_extractor
        if self.extractor is not None:
            _text.concept_extractor = self.extractor
        return self

    def __exit__(self, *exc):
        from . import text as _text

        _text.concept_extractor = self.saved
        return False


def _compute(items: list, metric_names: list[str]) -> tuple[dict, dict]:
    """Run the metrics on the items: ``(per-file outputs, {column: per-row values} to write back)``
    (``eval_metrics.py:72-133``). This is the only stage that touches the device."""
    outputs: dict = {"_num_samples": len(items)}
    columns: dict = {}
    for name in metric_names:
        info = get_metric_info(name)
        if info.name in METRICS_SAVING_INTERMEDIATE_VALUES:
            out = info.group_fn(info.builder_fn(items), reduce="none")
            extra = {}
            if info.name == "concept_semantic_similarity":
                concepts = list(zip(*out, strict=True))[0] if out else []
                sims = list(zip(*out, strict=True))[1] if out else []
                out = [np.max(row) for row in sims]
                extra["last_resp_concepts"] = concepts
                extra["last_resp_concepts_similarities"] = sims
            elif info.name == "mean_average_semantic_similarity":
                avg = out.pop("semantic_similarity@avg")
                extra.update(out)
                out = avg
            columns[info.name] = out
            columns.update(extra)
            value = float(np.mean(out)) if len(out) else float("nan")
        else:
            value = info.group_fn(info.builder_fn(items))
        if isinstance(value, dict):
            outputs.update(value)
        else:
            outputs[name] = value
    return outputs, columns


def _write_columns(input_file: str, df, columns: dict) -> None:
    """Add the per-row columns and rewrite the file the way the reference does (``df.to_json(..., lines=True,
    orient="records")``, ``eval_metrics.py:123``) - through a temporary file, so a crash cannot truncate the samples."""
    if not columns:
        return
    for key, val in columns.items():
        df[key] = val
    tmp = input_file + ".tmp"
    df.to_json(tmp, lines=True, orient="records")
    os.replace(tmp, input_file)


# ---- native JSONL path -----------------------------------------------------------------
# The two paired-similarity metrics only ever look at ref[0] / pred[-1] of each item and only write numeric columns,
# which is exactly what csrc/jsonl.cpp extracts and writes. Anything else (the concept metric's list columns, files
# with other row shapes, empty files) takes the pandas path, so `--io native` never changes what is computed.
_NATIVE_METRICS = {"semantic_similarity", "mean_average_semantic_similarity"}
_NATIVE_WITH_CONCEPTS = _NATIVE_METRICS | {"concept_semantic_similarity"}   # needs a pre-extracted concepts column


def _native_eligible(metric_names: list[str]) -> bool:
    return all(get_metric_info(name).name in _NATIVE_WITH_CONCEPTS for name in metric_names)


def _needs_concepts(metric_names: list[str]) -> bool:
    return any(get_metric_info(name).name == "concept_semantic_similarity" for name in metric_names)


def _load_items_native(input_file: str, metric_names: list[str] | None = None, concepts_column: str = "auto"):
    """``(raw bytes, items, extractor)`` with the items as ``jsonl.PackedPairs`` (the ``ref[0]`` / ``pred[-1]`` strings
    of every row in two packed columns), or ``None`` when the file needs the pandas path. With the concept metric the
    items are ordinary tuples and ``extractor`` serves the file's pre-extracted concepts (no such column: pandas path,
    where the spaCy extractor runs)."""
    from . import jsonl

    with open(input_file, "rb") as f:
        raw = f.read()
    try:
        items = jsonl.read_packed_pairs(raw)     # packed columns: no Python object per row anywhere on this path
        extractor = None
        if metric_names and _needs_concepts(metric_names):
            column = None
            if concepts_column != "none":
                names = CONCEPT_COLUMNS if concepts_column == "auto" else (concepts_column,)
                column = next((c for c in names if len(items) and jsonl.has_members(raw, [c])), None)
            if column is None:
                log.info("%s: no pre-extracted concepts column; using the pandas path", input_file)
                return None
            items = items.tolist()
            extractor = _concept_extractor_from(column, items, jsonl.read_json_column(raw, column))
    except jsonl.Unsupported as exc:
        log.info("%s: %s; using the pandas path", input_file, exc)
        return None
    return (raw, items, extractor) if len(items) else None


def _write_columns_native(input_file: str, raw: bytes, columns: dict) -> None:
    from . import jsonl

    jsonl.write_columns(input_file, raw, columns)


def score_file(input_file: str, metric_names: list[str], io: str = "pandas", concepts_column: str = "auto") -> dict:
    """Score one ``*_samples_*.jsonl`` file and write the per-row columns back (``eval_metrics.py:56-133``)."""
    if io == "native" and _native_eligible(metric_names):
        loaded = _load_items_native(input_file, metric_names, concepts_column)
        if loaded is not None:
            raw, items, extractor = loaded
            with _use_extractor(extractor):
                outputs, columns = _compute(items, metric_names)
            _write_columns_native(input_file, raw, columns)
            return outputs
    df, items = _load_items(input_file)
    extractor = None
    if _needs_concepts(metric_names):
        column = _concept_source(df.columns, concepts_column)
        if column is not None:
            extractor = _concept_extractor_from(column, items, df[column].tolist())
    with _use_extractor(extractor):
        outputs, columns = _compute(items, metric_names)
    _write_columns(input_file, df, columns)
    return outputs


# ---- parallel file I/O -----------------------------------------------------------------
# At sweep scale (130 files, 1.09 M rows) the device needs ~0.5 s for all predictions while pandas needs ~0.2 s per
# file to parse and re-serialise it. The JSON work is pure CPU and independent per file, so `--jobs N` moves it to a
# process pool: workers parse (and later rewrite) files, the parent process owns the device and does all scoring.
# Files are byte-identical to the serial path: the same pandas calls produce them.
def _worker_load(input_file: str):
    _, items = _load_items(input_file)
    return input_file, items


def _worker_write(job) -> str:
    input_file, columns = job
    df, _ = _load_items(input_file)
    _write_columns(input_file, df, columns)
    return input_file


def _already_scored(input_file: str, metric_names: list[str]) -> bool:
    """True if the first row already carries every per-row metric column (cheap: one line is parsed)."""
    import json

    wanted = [m for m in metric_names if m in METRICS_SAVING_INTERMEDIATE_VALUES]
    if not wanted or len(wanted) != len(metric_names):
        return False
    try:
        with open(input_file, encoding="utf-8") as f:
            first = json.loads(f.readline())
    except (OSError, ValueError):
        return False
    return all(m in first for m in wanted)


def _score_files_native(todo: list[str], metric_names: list[str], jobs: int, concepts_column: str = "auto") -> dict:
    """Native I/O on worker THREADS (the scanner and the file writer run outside the GIL): a few files are parsed
    ahead of the one being scored, files are scored one after the other in this thread, and each is rewritten in
    the background while the next is being scored."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor

    results: dict = {}
    # the scanner does ~1 GB/s per thread, so a few workers keep up with the device; more only fight the scoring
    # thread for the GIL and the tokenizer's threads for cores (8-core host: 3 workers 0.75 s, 16 workers 1.45 s)
    workers = max(2, min(jobs, 4))
    ahead = 2                               # parsed files waiting for the device
    with ThreadPoolExecutor(max_workers=workers) as pool:
        loads: deque = deque()
        writes = []
        nxt = 0
        for input_file in todo:
            while nxt < len(todo) and len(loads) < ahead:
                loads.append(pool.submit(_load_items_native, todo[nxt], metric_names, concepts_column))
                nxt += 1
            loaded = loads.popleft().result()
            if loaded is None:
                results[input_file] = score_file(input_file, metric_names, concepts_column=concepts_column)   # pandas path
                continue
            raw, items, extractor = loaded
            with _use_extractor(extractor):
                outputs, columns = _compute(items, metric_names)
            results[input_file] = outputs
            if columns:
                writes.append(pool.submit(_write_columns_native, input_file, raw, columns))
            del raw, items, loaded
        for w in writes:
            w.result()
    return results


def _stored_outputs(input_file: str, metric_names: list[str]) -> dict:
    """The per-file outputs of an already scored file, recomputed from its stored per-row columns (``--skip-scored``):
    the means the reference prints are means of exactly these columns (``eval_metrics.py:125``), so a skipped file stays
    in the report and in the "keep the run with more samples" rule."""
    import pandas as pd

    df = pd.read_json(input_file, lines=True)
    outputs: dict = {"_num_samples": len(df)}
    for name in metric_names:
        outputs[name] = float(np.mean(df[name].to_numpy(dtype=np.float64))) if len(df) else float("nan")
    return outputs


def score_files(input_files: list[str], metric_names: list[str], jobs: int = 1, skip_scored: bool = False,
                io: str = "pandas", concepts_column: str = "auto") -> dict:
    """``{file: outputs}`` for many files; with ``jobs > 1`` parsing and rewriting run in workers (threads for the
    native scanner, a process pool for pandas)."""
    if io not in ("pandas", "native"):
        raise ValueError(f"unknown io mode {io!r}: expected 'pandas' or 'native'")
    todo = [f for f in input_files if not (skip_scored and _already_scored(f, metric_names))]
    results: dict = {}
    for f in input_files:
        if f not in todo:
            log.info("Skipping %s: already scored (its stored columns feed the report)", f)
            results[f] = _stored_outputs(f, metric_names)
    if io == "native" and _native_eligible(metric_names) and jobs > 1 and len(todo) > 1:
        results.update(_score_files_native(todo, metric_names, jobs, concepts_column))
        return results
    if jobs <= 1 or len(todo) <= 1:
        for f in todo:
            results[f] = score_file(f, metric_names, io=io, concepts_column=concepts_column)
        return results
    if _needs_concepts(metric_names):       # the extractor of a file is built from its frame: keep it in one process
        for f in todo:
            results[f] = score_file(f, metric_names, io=io, concepts_column=concepts_column)
        return results
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    with ProcessPoolExecutor(max_workers=jobs, mp_context=mp.get_context("spawn")) as pool:
        writes = []
        for input_file, items in pool.map(_worker_load, todo):
            outputs, columns = _compute(items, metric_names)          # device work, parent process only
            results[input_file] = outputs
            if columns:
                writes.append(pool.submit(_worker_write, (input_file, columns)))
        for w in writes:
            w.result()
    return results


# ---- one process per GPU: files are independent units, so ranks share nothing but the final report -----------------
def split_files_balanced(input_files: list[str], world: int) -> list[list[str]]:
    """Deal files to ``world`` ranks, largest first onto the lightest rank (bytes on disk stand in for rows); every
    rank computes the same deal from the same sorted list. Order inside a rank follows ``input_files``."""
    if world < 1:
        raise ValueError("world must be >= 1")
    sizes = {f: (os.path.getsize(f) if os.path.exists(f) else 0) for f in input_files}
    load = [0] * world
    owner = {}
    for f in sorted(input_files, key=lambda f: (-sizes[f], f)):
        r = min(range(world), key=lambda i: (load[i], i))
        owner[f] = r
        load[r] += sizes[f] + 1
    return [[f for f in input_files if owner[f] == r] for r in range(world)]


def _dist_context():
    """``(rank, world, local_rank)`` from the launcher's environment (torchrun); initialises a gloo group for the one
    object gather the CLI needs (results are a few floats per file; the data path has no collective)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group("gloo")
    return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))


def format_report(tasks_outputs: dict) -> str:
    """``eval_metrics.py:155-171``: one block per (task, metric), one ``{model:<29}: {value:.3f}`` line per model."""
    blocks = []
    for task_name, task_outputs in tasks_outputs.items():
        names = sorted({k for m in task_outputs for k in task_outputs[m]})
        for metric_name in names:
            if metric_name.startswith("_"):
                continue
            text = f"{metric_name.capitalize().replace('_', ' ')} on {task_name}:\n"
            for model_name in task_outputs:
                text += f"{model_name:<29}: {task_outputs[model_name][metric_name]:.3f}\n"
            blocks.append(text)
    return "\n".join(blocks)


def main(args: Namespace) -> dict:
    log.setLevel(args.log_level)
    if args.seed:
        import torch

        random.seed(args.seed)
        np.random.seed(args.seed)
        torch.manual_seed(args.seed)
        torch.cuda.manual_seed_all(args.seed)
    input_files = find_input_files(args.input)
    log.info("Found %d `jsonl` files to process", len(input_files))
    tasks_outputs: dict = {}
    rank, world, local_rank = _dist_context()
    mine = input_files
    if world > 1:
        from . import text as _text

        _text.default_rank = local_rank
        try:                                   # torch calls without a device argument stay on this rank's GPU
            import torch

            if torch.cuda.is_available():
                torch.cuda.set_device(local_rank % torch.cuda.device_count())
        except ImportError:  # pragma: no cover
            pass
        os.environ.setdefault("OWC_TOKENIZER_THREADS", str(max(1, (os.cpu_count() or 1) // world)))   # share the host
        mine = split_files_balanced(input_files, world)[rank]
        log.info("Rank %d/%d scores %d of the files", rank, world, len(mine))
    failure: BaseException | None = None
    per_file: dict = {}
    try:
        per_file = score_files(mine, args.metrics.split(","), jobs=getattr(args, "jobs", 1),
                               skip_scored=getattr(args, "skip_scored", False), io=getattr(args, "io", "pandas"),
                               concepts_column=getattr(args, "concepts_column", "auto"))
    except BaseException as exc:   # noqa: BLE001 - under torchrun the other ranks must not wait forever in the gather
        if world == 1:
            raise
        failure = exc
    if world > 1:
        import torch.distributed as dist

        try:
            # a few floats per file (the data path has no collective); a failed rank sends its error instead, so every
            # rank reaches the gather and rank 0 re-raises
            payload = (failure is None, per_file if failure is None else f"{type(failure).__name__}: {failure}")
            parts = [None] * world if rank == 0 else None
            dist.gather_object(payload, parts, dst=0)
        finally:
            dist.destroy_process_group()
        if failure is not None:
            raise failure
        if rank != 0:
            return {}
        bad = [f"rank {r}: {msg}" for r, (ok, msg) in enumerate(parts) if not ok]
        if bad:
            raise RuntimeError("scoring failed on " + "; ".join(bad))
        per_file = {f: out for _, part in parts for f, out in part.items()}
    for input_file in input_files:
        if input_file not in per_file:
            continue
        task_name = Path(input_file).parent.parent.name
        model_name = Path(input_file).parent.name
        outputs = per_file[input_file]
        models = tasks_outputs.setdefault(task_name, {})
        # duplicate runs: keep the one with more samples, the older if even (eval_metrics.py:138-152)
        if model_name not in models or outputs["_num_samples"] > models[model_name]["_num_samples"]:
            models[model_name] = outputs
    report = format_report(tasks_outputs)
    if report:
        print(report)
    return tasks_outputs


def build_parser() -> ArgumentParser:
    parser = ArgumentParser()
    parser.add_argument("-i", "--input", required=True, type=str,
                        help="Path to the folder/file containing the samples to process")
    parser.add_argument("-m", "--metrics", required=True, type=str,
                        help="List of comma-separated metrics to evaluate on the data")
    parser.add_argument("--seed", type=int, default=1234, help="Random seed for reproducibility (default: 1234)")
    parser.add_argument("--log-level", type=str, default="INFO", help="Logging level (default: INFO)")
    # additions of this build (the reference's flags above are unchanged)
    parser.add_argument("--jobs", type=int, default=1,
                        help="Worker processes for JSONL parsing / rewriting (scoring stays in this process; default: 1)")
    parser.add_argument("--skip-scored", action="store_true",
                        help="Skip files whose rows already carry the requested per-row metric columns")
    parser.add_argument("--concepts-column", type=str, default="auto",
                        help="concept_semantic_similarity: column of the samples file with pre-extracted concepts "
                             "(`prediction_concepts`, or `last_resp_concepts` as written by the reference); 'auto' uses "
                             "one if present, 'none' always runs the spaCy extractor (default: auto)")
    parser.add_argument("--io", choices=["pandas", "native"], default="pandas",
                        help="JSONL reader/writer: pandas (byte-identical to the reference's rewrite; default) or the "
                             "library's streaming scanner (same records, untouched members keep their bytes)")
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
