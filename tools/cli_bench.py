"""The CLI mirror end to end at BASELINE config 5 scale: `eval_metrics -m semantic_similarity` over 13 models x 10
tasks of `*_samples_*.jsonl` files (the tracker's row format, `src/engine/_tracker.py:314-341`; real class names,
test-split sizes, synthetic predictions) - files on disk -> scores -> files rewritten - on one GPU.
Usage: python tools/cli_bench.py [n_models] [io modes, comma separated: native,pandas]"""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import owc_b200  # noqa: E402,F401
from owc_b200 import arch as A, eval_metrics as EM, synth, text as T  # noqa: E402
from owc_b200.weights import random_state_dict  # noqa: E402

TEST_ROWS = {"caltech101": 2465, "dtd": 1692, "eurosat": 8100, "fgvc_aircraft": 3333, "flowers102": 2463,
             "food101": 30300, "oxford_pets": 3669, "stanford_cars": 8041, "sun397": 19850, "ucf101": 3783}


def write_logs(root: str, n_models: int, words) -> int:
    total = 0
    for ti, (ds, rows) in enumerate(TEST_ROWS.items()):
        names = synth.make_class_names(100, words, seed=77) if ds == "fgvc_aircraft" else synth.real_class_names(ds)
        for m in range(n_models):
            preds, targets = synth.make_predictions(rows, names, words, seed=1000 * ti + m)
            run = os.path.join(root, ds, f"model{m:02d}")
            os.makedirs(run)
            with open(os.path.join(run, f"20250101_samples_{ds}.jsonl"), "w", encoding="utf-8") as f:
                for i, (p, t) in enumerate(zip(preds, targets)):
                    f.write(json.dumps({
                        "doc_id": i, "doc": {"image": f"<image {i}>", "label": int(t)}, "target": names[t],
                        "arguments": {"0": "What type of object is in this photo?", "1": {"max_new_tokens": 32, "temperature": 0,
                                                                                         "until": ["\n\n"]}},
                        "resps": [[p]], "filtered_resps": [p], "doc_hash": "d" * 64, "prompt_hash": "p" * 64,
                        "target_hash": "t" * 64, "input": "What type of object is in this photo?"}, ensure_ascii=False) + "\n")
            total += rows
    return total


if __name__ == "__main__":
    n_models = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    modes = (sys.argv[2] if len(sys.argv) > 2 else "native,pandas").split(",")
    # under torchrun: one process per GPU, files dealt to the ranks by eval_metrics.main; rank 0 writes the logs
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)      # torch.cuda.synchronize() below must not create a context on GPU 0 in every rank
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo")
    arch = A.MINILM_L6
    vocab = synth.build_vocab("bert", arch.vocab)
    words = synth.vocab_words(vocab)
    T.configure(arch, random_state_dict(arch), vocab, device=local_rank)
    base = tempfile.mkdtemp() if world == 1 else os.path.join(tempfile.gettempdir(), f"owc_cli_bench_{os.environ.get('MASTER_PORT', '0')}")
    src = os.path.join(base, "src", "logs")
    warm = os.path.join(base, f"warm{rank}", "logs")
    total = size = 0
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
        t0 = time.perf_counter()
        total = write_logs(src, n_models, words)
        size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(src) for f in fs)
        print(f"wrote {n_models * len(TEST_ROWS)} files, {total} rows, {size / 1e6:.0f} MB in {time.perf_counter() - t0:.1f}s", file=sys.stderr)
        for mode in modes:
            shutil.copytree(src, os.path.join(base, mode, "logs"))
    if world > 1:
        dist.barrier()
    # warm-up on this rank's GPU: one small task end to end, outside the multi-rank dealing
    shutil.copytree(os.path.join(src, "dtd"), os.path.join(warm, "dtd"))
    saved = os.environ.pop("WORLD_SIZE", None)
    import contextlib
    import io as _io
    with contextlib.redirect_stdout(_io.StringIO()):
        EM.main(EM.build_parser().parse_args(["-i", warm, "-m", "semantic_similarity", "--io", "native", "--log-level", "WARNING"]))
    if saved is not None:
        os.environ["WORLD_SIZE"] = saved
    out = {"workload": f"c5 through the CLI mirror: {n_models} models x 10 tasks = {n_models * len(TEST_ROWS)} files, {total} rows, "
                       f"{size / 1e6:.0f} MB of JSONL; eval_metrics -m semantic_similarity (files -> paired scores -> files rewritten), "
                       f"{world} GPU(s), one process each",
           "rows": total, "n_gpus": world}
    results = {}
    T0 = [time.perf_counter()]
    _score_files = EM.score_files

    def timed_score_files(files, *a, **k):          # per-rank view: own files only, before the gather
        t = time.perf_counter()
        try:
            return _score_files(files, *a, **k)
        finally:
            torch.cuda.synchronize()
            print(f"rank {rank}: {len(files)} files scored in {time.perf_counter() - t:.2f}s", file=sys.stderr, flush=True)

    EM.score_files = timed_score_files
    if world > 1 and os.environ.get("OWC_CLI_TIMING"):           # where a multi-rank run spends the rest
        def timed(mod, name):
            fn = getattr(mod, name)

            def wrapper(*a, **k):
                t = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    print(f"rank {rank}: {name} {time.perf_counter() - t:.3f}s at {time.perf_counter() - T0[0]:.3f}s", file=sys.stderr, flush=True)

            setattr(mod, name, wrapper)

        for _name in ("format_report", "split_files_balanced", "_dist_context"):
            timed(EM, _name)
        timed(torch.cuda, "manual_seed_all")
        timed(torch, "manual_seed")
        _find = EM.find_input_files

        def timed_find(*a, **k):
            t = time.perf_counter()
            try:
                return _find(*a, **k)
            finally:
                print(f"rank {rank}: find_input_files {time.perf_counter() - t:.3f}s", file=sys.stderr, flush=True)

        EM.find_input_files = timed_find
    for mode in modes:
        work = os.path.join(base, mode, "logs")
        argv = ["-i", work, "-m", "semantic_similarity", "--io", mode, "--jobs", "16", "--log-level", "WARNING"]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T0[0] = t0
        with contextlib.redirect_stdout(_io.StringIO()):
            results[mode] = EM.main(EM.build_parser().parse_args(argv))      # gathers on rank 0 when world > 1
        t_main = time.perf_counter() - t0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if os.environ.get("OWC_CLI_TIMING"):
            print(f"rank {rank}: main {t_main:.3f}s, with final sync {dt:.3f}s", file=sys.stderr, flush=True)
        if rank == 0:
            out[mode] = {"seconds": dt, "rows_per_s": total / dt, "args": " ".join(argv[2:])}
    if rank == 0:
        if len(results) == 2:
            a, b = (results[m] for m in modes)
            out["same_outputs"] = a == b
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
    T.reset()
