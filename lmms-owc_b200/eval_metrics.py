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
  own files on its own GPU (the reference spreads ``datasets.map`` over the GPUs with ``num_proc``, ``_group.py:521-535``)
  and rank 0 prints the report;
* ``--io native`` reads and rewrites the files with the library's streaming JSONL scanner (``jsonl.py``) instead of
  pandas: same records, untouched members keep their bytes, ~8x less host time per file.

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


def _native_eligible(metric_names: list[str]) -> bool:
    return all(get_metric_info(name).name in _NATIVE_METRICS for name in metric_names)


def _load_items_native(input_file: str):
    """``(raw bytes, items)`` with the items as ``jsonl.PackedPairs`` (the ``ref[0]`` / ``pred[-1]`` strings of every
    row in two packed columns), or ``None`` when the file needs the pandas path."""
    from . import jsonl

    with open(input_file, "rb") as f:
        raw = f.read()
    try:
        items = jsonl.read_packed_pairs(raw)     # packed columns: no Python object per row anywhere on this path
    except jsonl.Unsupported as exc:
        log.info("%s: %s; using the pandas path", input_file, exc)
        return None
    return (raw, items) if len(items) else None


def _write_columns_native(input_file: str, raw: bytes, columns: dict) -> None:
    from . import jsonl

    jsonl.write_columns(input_file, raw, columns)


def score_file(input_file: str, metric_names: list[str], io: str = "pandas") -> dict:
    """Score one ``*_samples_*.jsonl`` file and write the per-row columns back (``eval_metrics.py:56-133``)."""
    if io == "native" and _native_eligible(metric_names):
        loaded = _load_items_native(input_file)
        if loaded is not None:
            raw, items = loaded
            outputs, columns = _compute(items, metric_names)
            _write_columns_native(input_file, raw, columns)
            return outputs
    df, items = _load_items(input_file)
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


def _score_files_native(todo: list[str], metric_names: list[str], jobs: int) -> dict:
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
                loads.append(pool.submit(_load_items_native, todo[nxt]))
                nxt += 1
            loaded = loads.popleft().result()
            if loaded is None:
                results[input_file] = score_file(input_file, metric_names)          # pandas path
                continue
            raw, items = loaded
            outputs, columns = _compute(items, metric_names)
            results[input_file] = outputs
            if columns:
                writes.append(pool.submit(_write_columns_native, input_file, raw, columns))
            del raw, items, loaded
        for w in writes:
            w.result()
    return results


def score_files(input_files: list[str], metric_names: list[str], jobs: int = 1, skip_scored: bool = False,
                io: str = "pandas") -> dict:
    """``{file: outputs}`` for many files; with ``jobs > 1`` parsing and rewriting run in workers (threads for the
    native scanner, a process pool for pandas)."""
    if io not in ("pandas", "native"):
        raise ValueError(f"unknown io mode {io!r}: expected 'pandas' or 'native'")
    todo = [f for f in input_files if not (skip_scored and _already_scored(f, metric_names))]
    for f in input_files:
        if f not in todo:
            log.info("Skipping %s: already scored", f)
    results: dict = {}
    if io == "native" and _native_eligible(metric_names) and jobs > 1 and len(todo) > 1:
        return _score_files_native(todo, metric_names, jobs)
    if jobs <= 1 or len(todo) <= 1:
        for f in todo:
            results[f] = score_file(f, metric_names, io=io)
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
    per_file = score_files(mine, args.metrics.split(","), jobs=getattr(args, "jobs", 1),
                           skip_scored=getattr(args, "skip_scored", False), io=getattr(args, "io", "pandas"))
    if world > 1:
        import torch.distributed as dist

        parts = [None] * world if rank == 0 else None
        dist.gather_object(per_file, parts, dst=0)          # a few floats per file; the data path has no collective
        if rank != 0:
            return {}
        per_file = {f: out for part in parts for f, out in part.items()}
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
    parser.add_argument("--io", choices=["pandas", "native"], default="pandas",
                        help="JSONL reader/writer: pandas (byte-identical to the reference's rewrite; default) or the "
                             "library's streaming scanner (same records, untouched members keep their bytes)")
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
