"""Offline metric re-scoring CLI: drop-in for the reference's ``eval_metrics.py`` on the similarity path.

Same flags (``-i/--input``, ``-m/--metrics``, ``--seed``, ``--log-level``), same file discovery
(``**/*_samples_*.jsonl`` under ``logs/.../{task_name}/{model_name}/``), same item building
(``(target, filtered_resps)``, multi-round ``[[...]]`` unwrapped), same JSONL columns written back and
the same stdout blocks as ``eval_metrics.py:19-171``. Only the three similarity metrics are available
(the rest of the reference's metric zoo is outside this build); unknown metric names raise ``KeyError``
exactly like the reference's registry lookup. Differences, all additive:

* the sample file is rewritten atomically (temp file + ``os.replace``) instead of in place;
* ``--rank`` selects the GPU like the operator's ``rank`` argument.

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


def score_file(input_file: str, metric_names: list[str]) -> dict:
    """Score one ``*_samples_*.jsonl`` file and write the per-row columns back (``eval_metrics.py:56-133``)."""
    import pandas as pd

    df = pd.read_json(input_file, lines=True)
    predictions = df["filtered_resps"].tolist()
    references = df["target"].tolist()
    # multi-turn generation: drop the inner list (eval_metrics.py:66-68)
    if predictions and isinstance(predictions[0], list) and predictions[0] and isinstance(predictions[0][0], list):
        predictions = [p[0] for p in predictions]
    items = list(zip(references, predictions, strict=True))

    outputs: dict = {"_num_samples": len(items)}
    dirty = False
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
            df[info.name] = out
            for key, val in extra.items():
                df[key] = val
            dirty = True
            value = float(np.mean(out)) if len(out) else float("nan")
        else:
            value = info.group_fn(info.builder_fn(items))
        if isinstance(value, dict):
            outputs.update(value)
        else:
            outputs[name] = value
    if dirty:
        tmp = input_file + ".tmp"
        df.to_json(tmp, lines=True, orient="records")
        os.replace(tmp, input_file)
    return outputs


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
    for input_file in input_files:
        task_name = Path(input_file).parent.parent.name
        model_name = Path(input_file).parent.name
        outputs = score_file(input_file, args.metrics.split(","))
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
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
