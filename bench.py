"""bench.py - predictions scored/sec (encode + match) on the B200-native scoring path.

    python bench.py --gpus N --steps K --warmup W            # this build (one rank per GPU via torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload at N=1 = BASELINE.json configs[1]: Food101 fine-grained matching, 100k synthetic predictions
against the 101 real Food101 class names, all-MiniLM-L6-v2 architecture (random-init weights, synthetic
WordPiece vocabulary - there is no network), top-5 + paired target score. A "step" is one pass of the
hot path over that batch. Multi-GPU: every rank scores its own 100k-row shard (weak scaling), class
embeddings encoded on rank 0 and broadcast, per-shard results gathered on rank 0.

One JSON line on stdout (rank 0). `value`: token ids already resident in HBM; `e2e`: token ids in pinned
host memory -> results in host memory through the C-ABI call `owc_score` (copies inside the timed region).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS_PER_GPU = 100_000
TOPK = 5
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks() -> tuple[dict, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path, encoding="utf-8") as f:
            return json.load(f), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 20 ms while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index = index
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = 0
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self) -> None:
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def finish(self) -> dict:
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz or None,
                "reasons": sorted(self.reasons), "samples": len(s)}


WORKLOADS = {
    # configs[0]: the reference's own CPU-runnable case (for the record: `--workload c1`, 1000 rows unless --rows is given)
    "c1": {"arch": "minilm-l6", "classes": "caltech101", "rows": 1000,
           "desc": "c1: Caltech101 100 class names x 1k synthetic predictions, MiniLM-L6 arch, top-5 + paired score"},
    # BASELINE.json configs[1] - the headline single-GPU configuration
    "c2": {"arch": "minilm-l6", "classes": "food101", "desc": "c2: Food101 101 class names x 100k synthetic predictions per GPU, MiniLM-L6 arch, top-5 + paired score"},
    # configs[2]: ImageNet-1k-scale matching with the mpnet-base architecture (parity-tested; bench on request)
    "c3": {"arch": "mpnet-base", "classes": 1000, "desc": "c3: 1000 synthetic class names x synthetic predictions, mpnet-base arch, top-5 + paired score"},
    # configs[3]: concept matching against a WordNet-noun-scale vocabulary
    "c4": {"arch": "minilm-l6", "classes": 117_000, "desc": "c4: 117k synthetic names x synthetic predictions, MiniLM-L6 arch, fused top-5"},
}


def build_workload(rank: int, rows: int, workload: str = "c2"):
    """Tokenised synthetic workload of one rank: class names + distinct predictions."""
    import numpy as np

    import owc_b200  # noqa: F401
    from owc_b200 import arch as A
    from owc_b200 import synth
    from owc_b200.tokenizer import HostTokenizer

    spec = WORKLOADS[workload]
    arch = A.ARCHS[spec["arch"]]
    kind = "bert" if arch.kind == A.KIND_BERT else "mpnet"
    vocab = synth.build_vocab(kind, arch.vocab)
    words = synth.vocab_words(vocab)
    classes = (synth.real_class_names(spec["classes"]) if isinstance(spec["classes"], str)
               else synth.make_class_names(spec["classes"], words, seed=1234))
    preds, targets = synth.make_predictions(rows, classes, words, seed=1234 + rank)
    tok = HostTokenizer(vocab, kind)
    cids, ccu = tok.encode_packed(classes)
    pids, pcu = tok.encode_packed(preds)
    return {"vocab": vocab, "classes": classes, "preds": preds, "targets": np.asarray(targets, np.int32),
            "cids": cids, "ccu": ccu, "pids": pids, "pcu": pcu, "tok": tok, "arch": arch, "desc": spec["desc"]}


def cpu_reference_rate(sample_rows: int, steps: int, warmup: int, wl=None) -> dict:
    """The reference's own CPU path for this metric, restated by the oracle (kind "port"): HF-style
    tokenise (pad to longest, 1024-row batches), encode the reference column per row AND the prediction
    column (no dedup, as `_group.py:521-535` does), paired dot; plus the N x C matrix + top-k the
    north star adds (torch restatement). All host threads torch can use."""
    import torch

    import owc_b200  # noqa: F401
    from oracle import encoder_ref, metrics_ref
    from oracle.hf_models import build_hf_tokenizer
    from owc_b200 import arch as A
    from owc_b200.weights import random_state_dict

    torch.set_num_threads(os.cpu_count() or 1)
    wl = wl or build_workload(0, max(sample_rows, 1024))
    arch = wl["arch"]
    sd = random_state_dict(arch)
    hf_tok = build_hf_tokenizer(arch, wl["vocab"])
    classes = wl["classes"]

    def encode(texts):
        return encoder_ref.encode_strings(sd, arch, hf_tok, list(texts))

    def one_step(lo: int) -> None:
        preds = wl["preds"][lo:lo + sample_rows]
        tg = wl["targets"][lo:lo + sample_rows]
        items = [(classes[t], [p]) for t, p in zip(tg, preds)]
        refs, prs = metrics_ref.select_items(items)
        zr = encode(refs)                        # the reference encodes the class name once per ROW (F5)
        zp = encode(prs)
        encoder_ref.paired_dot(zr, zp)           # _group.py:544
        zc = encode(classes)                     # + the north-star matrix / top-k on the same embeddings
        encoder_ref.sim_topk(zp, zc, TOPK)

    for _ in range(warmup):
        one_step(0)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step(0)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": sample_rows / dt, "unit": "predictions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_rows} of the workload's rows per step: reference semantic_similarity body (refs re-encoded "
                      f"per row, preds, paired dot) + NxC matrix/top-{TOPK}; fp32, torch CPU, {steps} timed steps",
            "ms_per_step": dt * 1e3}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2048
    r = cpu_reference_rate(sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "predictions scored/sec (encode+match)", "value": r["value"],
        "unit": "predictions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: Food101 101 class names x 100k synthetic predictions, MiniLM-L6 arch, top-5 + paired score",
                   "sample_rows_per_step": sample},
        "cpu_baseline": {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import owc_b200  # noqa: F401
    from owc_b200 import arch as A
    from owc_b200 import sharded
    from owc_b200.engine import Engine
    from owc_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner on stdout when the communicator is created; stdout carries exactly one
        # JSON line, so route fd 1 to stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rows = args.rows
    wl = build_workload(rank, rows, args.workload)
    arch = wl["arch"]
    eng = Engine(arch, random_state_dict(arch), device=local_rank, chunk_tokens=args.chunk_tokens)
    n_tok = int(wl["pcu"][-1])
    counts = [rows] * world

    ids_pin = torch.from_numpy(wl["pids"]).pin_memory()
    tg_pin = torch.from_numpy(wl["targets"]).pin_memory()
    ids_dev = ids_pin.to(dev)
    tg_dev = tg_pin.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        zc = sharded.broadcast_class_embeddings(eng, wl["cids"], wl["ccu"])
        res = eng.score(ids_dev, wl["pcu"], zc, targets=tg_dev, k=TOPK, out="device")
        if world > 1:
            for key in ("scores", "idx", "paired"):
                sharded.gather_rows(res[key], counts)
        return res

    def step_host():
        zc = sharded.broadcast_class_embeddings(eng, wl["cids"], wl["ccu"])
        res = eng.score(ids_pin, wl["pcu"], zc, targets=tg_pin, k=TOPK, out="host")
        if world > 1:   # results are on the host already; gather the per-shard means like eval_metrics prints
            t = torch.tensor([float(res["paired"].mean())], device=dev)
            dist.all_reduce(t)
        return res

    for _ in range(args.warmup):
        step_device()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    stats0 = eng.stats()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.zero_()                                   # evict L2 between timed steps (not timed)
        a.record(torch.cuda.current_stream(dev))
        step_device()
        b.record(torch.cuda.current_stream(dev))
    barrier()
    stats1 = eng.stats()
    clocks = sampler.finish()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = rows * world / (ms_per_step / 1e3)
    launches = (stats1["kernel_launches"] - stats0["kernel_launches"]) // args.steps

    # ---- e2e through the C-ABI with host buffers ----
    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    s1 = eng.stats()
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": rows * world * args.steps / e2e_s, "unit": "predictions/s",
           "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // args.steps,
           "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // args.steps,
           "ms_per_step": e2e_s / args.steps * 1e3,
           "path": "owc_score(ids, targets in pinned host memory -> scores/idx/paired in host memory)"}

    # ---- strings -> results (host tokenisation included; SURVEY 8d reports it beside the headline) ----
    e2e_strings = None
    if rank == 0 and world == 1:
        from owc_b200.pipeline import score_strings_pipelined

        tok = wl["tok"]
        tok.encode_packed(wl["preds"][:4096])
        t0 = time.perf_counter()
        tok.encode_packed(wl["preds"])
        t_tok = time.perf_counter() - t0
        zc = sharded.broadcast_class_embeddings(eng, wl["cids"], wl["ccu"])
        score_strings_pipelined(eng, tok, wl["preds"][:30000], zc, targets=wl["targets"][:30000], k=TOPK)   # warm
        t1 = time.perf_counter()
        score_strings_pipelined(eng, tok, wl["preds"], zc, targets=wl["targets"], k=TOPK)
        t_all = time.perf_counter() - t1
        e2e_strings = {"value": rows / t_all, "unit": "predictions/s", "ms_per_step": t_all * 1e3, "tokenize_alone_ms": t_tok * 1e3,
                       "tokenizer": "native C++ WordPiece (owc_tokenize_packed)" if tok._native is not None else "tokenizers backend",
                       "path": "strings -> WordPiece ids on a host thread, overlapped with owc_score of the previous 20k-row chunk "
                               "-> results in host memory (owc_b200.pipeline.score_strings_pipelined), one rank"}

    # ---- roofline leg: every launch bracketed by CUDA events on the engine's stream ----
    peaks, peak_src = load_peaks()
    eng.profile_begin()
    step_device()
    prof = eng.profile_end()
    h, inter = arch.hidden, arch.inter
    lens = np.diff(wl["pcu"]).astype(np.float64)
    gemm_flops = arch.layers * 2.0 * (4 * h * h + 2 * h * inter) * n_tok           # encoder contractions, real tokens
    attn_flops = arch.layers * 4.0 * h * float((lens * lens).sum())                 # QK^T + PV (4*H*T^2 per sequence-layer)
    fused = arch.kind == A.KIND_BERT and h == 384 and arch.heads == 12 and inter % 128 == 0 and lens.max() <= 64 \
        and os.environ.get("OWC_FUSED", "1") != "0"
    gemm_ms = prof["gemm"]["ms"]
    # the fused kernel runs the attention inside the same launch: its algorithmic work is the whole encoder stack
    kernel_flops = gemm_flops + (attn_flops if fused else 0.0)
    achieved = kernel_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
    total_prof_ms = sum(v["ms"] for v in prof.values())
    launches_k = max(prof["gemm"]["launches"], 1)
    traffic, tinfo = None, {}
    tpath = os.path.join(ROOT, "profiles", "r01_fused_traffic.json")
    if fused and os.path.exists(tpath):     # dram bytes per launch of the same kernel from the committed ncu --set full capture
        with open(tpath, encoding="utf-8") as f:
            tinfo = json.load(f)
        traffic = tinfo.get("bytes_per_launch")
    kname = ("k_encoder_fused (all encoder layers of a 256-row block pair per cluster: QKV/O/FFN contractions on tcgen05, "
             "attention, GELU, LayerNorm in one launch)") if fused else "k_gemm_tc (encoder QKV/O/FFN contractions, tcgen05)"
    roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_source": (tinfo.get("launch") if traffic is not None else None),
                "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "algorithmic_flops_per_launch": kernel_flops / launches_k,
                "launches_per_step": prof["gemm"]["launches"],
                "avg_launch_ms": gemm_ms / launches_k,
                "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None,
                "per_kernel_ms": {k: round(v["ms"], 4) for k, v in prof.items()}}

    line = {
        "metric": "predictions scored/sec (encode+match)", "value": value, "unit": "predictions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["desc"],
                   "rows_per_gpu": rows, "tokens_per_gpu": n_tok, "classes": len(wl["classes"]), "topk": TOPK,
                   "l2": "256 MB buffer written between timed steps; per-step activations (~GBs) also exceed the 126 MB L2",
                   "parallelism": f"rows sharded x{world}; class embeddings broadcast; results gathered on rank 0"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }
    if e2e_strings is not None:
        line["e2e_strings"] = e2e_strings

    if rank == 0 and not args.no_sim_leg:
        # north-star second metric: the fused similarity GEMM + top-k kernel alone at a 117k-name vocabulary
        n_sim, c_sim = 148 * 128 * 2, 117_000
        ms = eng.bench_sim_kernel(n_sim, c_sim, h, TOPK, 5)
        tf = 2.0 * n_sim * c_sim * h / (ms / 1e3) / 1e12
        burst = float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"]))
        line["sim_gemm_topk"] = {"rows": n_sim, "classes": c_sim, "dim": h, "k": TOPK, "ms_per_launch": ms, "tflops": tf,
                                 "peak": burst, "frac": tf / burst, "peak_source": f"{peak_src} bf16_tflops (kernel timed alone)",
                                 "algorithmic_bytes": 2 * h * (n_sim + c_sim) + 8 * 8 * n_sim}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_rate(min(2048, rows), 2, 1, wl)
        line["cpu_baseline"] = {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="prediction rows per GPU (default: the workload's own size)")
    ap.add_argument("--chunk-tokens", type=int, default=0)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 is the headline; c3/c4 for the record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sim-leg", action="store_true")
    args = ap.parse_args()
    if args.rows <= 0:
        args.rows = WORKLOADS[args.workload].get("rows", ROWS_PER_GPU)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
