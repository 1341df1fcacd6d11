"""bench.py - predictions scored/sec (encode + match) on the B200-native scoring path.

    python bench.py --gpus N --steps K --warmup W            # this build (one rank per GPU via torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Headline workload = BASELINE.json configs[2] (c3), the configuration the metric "at 1/2/4/8 B200" is quoted on:
ImageNet-1k-scale matching, ONE fixed set of 1 M synthetic predictions against 1000 synthetic class names with the
all-mpnet-base-v2 architecture (random-init weights, synthetic WordPiece vocabulary - there is no network), top-5 +
paired target score. STRONG scaling: the 1 M rows are split N ways by token count (`sharded.split_rows_balanced`);
a step = class names encoded (rank 0, broadcast over NCCL, asynchronously) + every rank's shard encoded and matched +
ONE packed gather of (scores, idx, paired) to rank 0. `value`: token ids already resident in HBM; `e2e`: ids/targets in
pinned host memory -> results of ALL rows in rank 0's host memory (H2D and D2H inside the timed region).
The other BASELINE configs ride along under `sub` (c2: 100 k x Food101 MiniLM, c4: 1 M x 117 k-name vocabulary,
c5: the 13-model x 10-task sweep), each with its own value / e2e / roofline.

One JSON line on stdout (rank 0).
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

TOPK = 5
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
METRIC = "predictions scored/sec (encode+match)"

WORKLOADS = {
    # configs[0]: the reference's own CPU-runnable case
    "c1": {"arch": "minilm-l6", "classes": "caltech101", "rows": 1000,
           "desc": "c1: Caltech101 100 class names x 1k synthetic predictions, MiniLM-L6 arch, top-5 + paired score"},
    "c2": {"arch": "minilm-l6", "classes": "food101", "rows": 100_000,
           "desc": "c2: Food101 101 class names x 100k synthetic predictions, MiniLM-L6 arch, top-5 + paired score"},
    # configs[2]: the configuration the metric is quoted on ("at 1/2/4/8 B200") - the headline
    "c3": {"arch": "mpnet-base", "classes": 1000, "rows": 1_000_000,
           "desc": "c3: 1M synthetic predictions x 1000 synthetic class names, all-mpnet-base-v2 arch, top-5 + paired score"},
    "c4": {"arch": "minilm-l6", "classes": 117_000, "rows": 1_000_000,
           "desc": "c4: 1M synthetic predictions x 117k-name synthetic vocabulary (WordNet-noun scale), MiniLM-L6 arch, fused top-5 + paired score"},
}
C5_TEST_ROWS = {"caltech101": 2465, "dtd": 1692, "eurosat": 8100, "fgvc_aircraft": 3333, "flowers102": 2463,
                "food101": 30300, "oxford_pets": 3669, "stanford_cars": 8041, "sun397": 19850, "ucf101": 3783}
C5_DESC = "c5: 13 models x 10 benchmarks (130 sample sets, 1.09M predictions, real class names) re-scored in one run, MiniLM-L6 arch"


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


# ------------------------------------------------------------------------------------------------------------------
# workloads (strings only: shared by both arms; nothing here loads the CUDA library)
# ------------------------------------------------------------------------------------------------------------------
def make_strings(workload: str, rows: int):
    """Class names, `rows` distinct predictions and a uniform-random target per row, seed 1234."""
    import owc_b200  # noqa: F401
    from owc_b200 import arch as A
    from owc_b200 import synth

    spec = WORKLOADS[workload]
    arch = A.ARCHS[spec["arch"]]
    kind = "bert" if arch.kind == A.KIND_BERT else "mpnet"
    vocab = synth.build_vocab(kind, arch.vocab)
    words = synth.vocab_words(vocab)
    classes = (synth.real_class_names(spec["classes"]) if isinstance(spec["classes"], str)
               else synth.make_class_names(spec["classes"], words, seed=1234))
    preds, targets = synth.make_predictions(rows, classes, words, seed=1234)
    return {"arch": arch, "kind": kind, "vocab": vocab, "classes": classes, "preds": preds, "targets": targets}


def config_for(workload: str, rows: int) -> dict:
    """The `config` object - identical in both arms."""
    spec = WORKLOADS[workload]
    n_cls = spec["classes"] if isinstance(spec["classes"], int) else {"caltech101": 100, "food101": 101}[spec["classes"]]
    return {"workload": spec["desc"], "arch": spec["arch"], "rows": rows, "classes": n_cls, "topk": TOPK,
            "partitioning": "one fixed set of prediction rows split across the GPUs (strong scaling)",
            "l2": "256 MB buffer written between timed steps; a step's activations (GBs) also exceed the 126 MB L2"}


def make_sweep(n_models: int = 13):
    """c5: the reference's ten classification tasks (real class names; test-split sizes from its split files) x models."""
    import owc_b200  # noqa: F401
    from owc_b200 import arch as A
    from owc_b200 import synth

    arch = A.MINILM_L6
    vocab = synth.build_vocab("bert", arch.vocab)
    words = synth.vocab_words(vocab)
    tasks = {}
    for ti, (ds, rows) in enumerate(C5_TEST_ROWS.items()):
        # FGVC-Aircraft's 100 variant names live in the downloaded dataset (variants.txt), not in the reference repo
        names = synth.make_class_names(100, words, seed=77) if ds == "fgvc_aircraft" else synth.real_class_names(ds)
        files = {}
        for m in range(n_models):
            preds, targets = synth.make_predictions(rows, names, words, seed=1000 * ti + m)
            files[f"model{m:02d}"] = {"predictions": preds, "targets": targets}
        tasks[ds] = {"class_names": names, "files": files}
    return {"arch": arch, "vocab": vocab, "tasks": tasks}


# ------------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port)
# ------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's own CPU path for this metric, restated by the oracle (kind "port"): HF tokenizer (pad to longest,
    1024-row batches), the reference column encoded per ROW and the prediction column (no dedup, as
    `_group.py:521-535` does), paired dot; plus the N x C matrix + top-k the north star adds (torch restatement).
    All host threads torch can use. It omits the reference's `datasets.map` / Arrow / `tolist()` overhead,
    so the GPU/CPU ratio computed from it is a lower bound."""

    def __init__(self, workload: str, max_rows: int) -> None:
        import torch

        from oracle.hf_models import build_hf_tokenizer
        from owc_b200.weights import random_state_dict

        torch.set_num_threads(os.cpu_count() or 1)
        self.torch = torch
        self.wl = make_strings(workload, max_rows)
        self.arch = self.wl["arch"]
        self.sd = random_state_dict(self.arch)
        self.hf_tok = build_hf_tokenizer(self.arch, self.wl["vocab"])
        self.workload = workload

    def encode_classes(self) -> float:
        from oracle import encoder_ref

        t0 = time.perf_counter()
        self.zc = encoder_ref.encode_strings(self.sd, self.arch, self.hf_tok, list(self.wl["classes"]))
        return time.perf_counter() - t0

    def step(self, rows: int) -> float:
        from oracle import encoder_ref, metrics_ref

        wl = self.wl
        t0 = time.perf_counter()
        preds, tg = wl["preds"][:rows], wl["targets"][:rows]
        items = [(wl["classes"][t], [p]) for t, p in zip(tg, preds)]
        refs, prs = metrics_ref.select_items(items)
        zr = encoder_ref.encode_strings(self.sd, self.arch, self.hf_tok, list(refs))   # class name re-encoded per ROW (F5)
        zp = encoder_ref.encode_strings(self.sd, self.arch, self.hf_tok, list(prs))
        encoder_ref.paired_dot(zr, zp)                                                 # _group.py:544
        encoder_ref.sim_topk(zp, self.zc, TOPK)                                        # the north-star matrix / top-k
        return time.perf_counter() - t0

    def rate(self, steps: int, warmup: int, budget_s: float, max_rows: int, rows_total: int) -> dict:
        """`steps` timed steps of a bounded row sample sized so that warm-up + timed steps fit `budget_s` of CPU time.
        The class names are encoded once and their time is charged pro rata (sample rows / workload rows), i.e. the
        per-row rate of the sample is extrapolated linearly to the full workload."""
        t_cls = self.encode_classes()
        probe_rows = min(256, max_rows)
        self.step(min(64, max_rows))                  # first-touch costs (thread pool, allocator)
        t_probe = self.step(probe_rows)
        per_row = t_probe / probe_rows
        n_steps = max(steps + max(warmup - 1, 0), 1)
        rows = int(max(probe_rows, min(max_rows, (budget_s - t_cls) / n_steps / max(per_row, 1e-9))))
        for _ in range(max(warmup - 1, 0)):
            self.step(rows)
        t = [self.step(rows) for _ in range(max(steps, 1))]
        dt = sum(t) / len(t) + t_cls * rows / max(rows_total, 1)
        return {"value": rows / dt, "unit": "predictions/s", "cores": self.torch.get_num_threads(), "kind": "port",
                "sample": f"{rows} of the workload's {rows_total} rows per step ({WORKLOADS[self.workload]['arch']} arch): reference "
                          f"semantic_similarity body (class name re-encoded per row, predictions, paired dot) + NxC matrix/top-{TOPK} "
                          f"against the {len(self.wl['classes'])} class embeddings (encoded once, {t_cls:.1f} s, charged pro rata = "
                          f"linear extrapolation to the full row count); fp32 torch CPU, HF tokenizer, {len(t)} timed steps after "
                          f"{warmup} warm-up; omits the reference's datasets.map/Arrow/tolist overhead, so GPU/CPU ratios from it "
                          f"are lower bounds",
                "sample_rows": rows, "ms_per_step": dt * 1e3}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows_total = args.rows or WORKLOADS[args.workload]["rows"]
    ref = CpuReference(args.workload, min(rows_total, 10_000))
    r = ref.rate(args.steps, args.warmup, budget_s=float(args.ref_budget_s), max_rows=min(rows_total, 10_000), rows_total=rows_total)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "predictions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for(args.workload, rows_total),
        "cpu_baseline": {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# this build
# ------------------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self) -> None:
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL prints its version banner on stdout when the communicator is created; stdout carries exactly one
            # JSON line, so route fd 1 to stderr until the communicator exists
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)

    def barrier(self) -> None:
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def bcast(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.broadcast(t, src=0)
        return float(t.item())

    def close(self) -> None:
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class RowWorkload:
    """One of c1-c4 on this rank: the fixed row set split by token count, this rank's shard tokenised, engine loaded."""

    def __init__(self, name: str, rows: int, d: Dist, chunk_tokens: int = 0) -> None:
        import numpy as np

        from owc_b200 import sharded
        from owc_b200.engine import Engine
        from owc_b200.tokenizer import HostTokenizer
        from owc_b200.weights import random_state_dict

        self.np, self.sharded, self.d, self.name, self.rows = np, sharded, d, name, rows
        wl = make_strings(name, rows)
        self.arch, self.classes, self.preds = wl["arch"], wl["classes"], wl["preds"]
        self.targets = np.asarray(wl["targets"], np.int32)
        self.tok = HostTokenizer(wl["vocab"], wl["kind"])
        self.cids, self.ccu = self.tok.encode_packed(self.classes)
        # the whole row set is tokenised once (untimed set-up): shards are balanced by TOKEN count, the unit the encoder's
        # work is proportional to (string length is only a proxy: 0.5 % off on this workload)
        self.all_ids, self.all_cu = self.tok.encode_packed(self.preds)
        self.eng = Engine(self.arch, random_state_dict(self.arch), device=d.local_rank, chunk_tokens=chunk_tokens)
        self.sharded_classes = d.world > 1 and len(self.classes) >= sharded.SHARD_CLASSES_FROM
        self.scratch: dict = {}
        self.head_start = 0.0
        self._shard()

    def _shard(self) -> None:
        torch, d, np = self.d.torch, self.d, self.np
        weights = np.diff(self.all_cu)
        self.blocks = self.sharded.split_rows_balanced(weights, d.world,
                                                       head_start=[self.head_start] + [0.0] * (d.world - 1) if d.world > 1 else None)
        b0, b1 = self.blocks[d.rank]
        t0, t1 = int(self.all_cu[b0]), int(self.all_cu[b1])
        self.pids = np.ascontiguousarray(self.all_ids[t0:t1])
        self.pcu = (self.all_cu[b0:b1 + 1] - self.all_cu[b0]).astype(np.int32)
        self.n_rows, self.n_tok = b1 - b0, t1 - t0
        self.weight_own = float(self.n_tok)
        self.ids_pin = torch.from_numpy(self.pids).pin_memory()
        self.tg_pin = torch.from_numpy(np.ascontiguousarray(self.targets[b0:b1])).pin_memory()
        self.ids_dev, self.tg_dev = self.ids_pin.to(d.dev), self.tg_pin.to(d.dev)
        self.gatherer = (self.sharded.ResultGatherer([e - b for b, e in self.blocks], TOPK, d.dev) if d.world > 1 else None)

    def step(self, host: bool):
        """One pass. host=False: ids resident in HBM, results gathered on rank 0's device. host=True: ids/targets in
        pinned host memory, results of all rows in rank 0's host memory."""
        d, eng = self.d, self.eng
        ids, tg = (self.ids_pin, self.tg_pin) if host else (self.ids_dev, self.tg_dev)
        if d.world == 1:
            zc = eng.encode(self.cids, self.ccu, out="device")
            return eng.score(ids, self.pcu, zc, targets=tg, k=TOPK, out="host" if host else "device")
        cls = self.sharded.class_embeddings(eng, self.cids, self.ccu, scratch=self.scratch)
        res = self.sharded.score_shard(eng, ids, self.pcu, cls, targets=tg, k=TOPK)
        recv = self.gatherer.gather(res)
        if host and d.rank == 0:
            return self.gatherer.unpack(recv)
        return res

    def calibrate_head_start(self) -> None:
        """Rank 0 encodes the class names before its shard: measure that head start in units of the split weights and
        take it off rank 0's share (strong scaling would otherwise wait for rank 0 every step)."""
        d, torch = self.d, self.d.torch
        if d.world == 1 or self.sharded_classes:
            return
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        d.barrier()
        ev[0].record()
        if d.rank == 0:
            self.eng.encode(self.cids, self.ccu, out="device")
        ev[1].record()
        self.eng.encode(self.ids_dev, self.pcu, out="device")
        ev[2].record()
        torch.cuda.synchronize()
        t_cls, t_rows = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        head = t_cls / max(t_rows, 1e-6) * self.weight_own if d.rank == 0 else 0.0
        self.head_start = d.bcast(head)
        self._shard()

    def close(self) -> None:
        self.eng.close()


def time_steps(d: Dist, fn, steps: int, flush) -> float:
    """ms per step: CUDA events on the current stream around each step, L2 flushed (untimed) in between, max over ranks."""
    torch = d.torch
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    d.barrier()
    for a, b in evs:
        flush.zero_()
        a.record(torch.cuda.current_stream(d.dev))
        fn()
        b.record(torch.cuda.current_stream(d.dev))
    d.barrier()
    return d.max(sum(a.elapsed_time(b) for a, b in evs)) / steps


def encoder_flops(arch, lens) -> tuple[float, float]:
    """Algorithmic FLOPs on real tokens: contractions 2*(4H^2 + 2HI) per token-layer, attention 4*H*T^2 per sequence-layer."""
    h, inter = arch.hidden, arch.inter
    n_tok = float(lens.sum())
    return arch.layers * 2.0 * (4 * h * h + 2 * h * inter) * n_tok, arch.layers * 4.0 * h * float((lens * lens).sum())


def measure_rows(name: str, rows: int, d: Dist, steps: int, warmup: int, flush, peaks: dict, peak_src: str,
                 chunk_tokens: int = 0, strings_leg: bool = True) -> dict:
    import numpy as np

    from owc_b200 import arch as A

    wl = RowWorkload(name, rows, d, chunk_tokens)
    eng, arch = wl.eng, wl.arch
    for _ in range(2):
        wl.step(False)
    wl.calibrate_head_start()
    for _ in range(max(warmup - 2, 1)):
        wl.step(False)
    sampler = ClockSampler(d.local_rank)
    sampler.start()
    s0 = eng.stats()
    ms = time_steps(d, lambda: wl.step(False), steps, flush)
    s1 = eng.stats()
    clocks = sampler.finish()
    launches = (s1["kernel_launches"] - s0["kernel_launches"]) // steps

    # ---- e2e: host buffers in, all rows' results in rank 0's host memory ----
    for _ in range(2):
        wl.step(True)
    d.barrier()
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(steps):
        wl.step(True)
    d.torch.cuda.synchronize()
    e2e_s = d.max(time.perf_counter() - t0)
    s1 = eng.stats()
    width = 2 * TOPK + 1
    d2h = (s1["d2h_bytes"] - s0["d2h_bytes"]) // steps + (rows * width * 4 if d.world > 1 and d.rank == 0 else 0)
    e2e = {"value": rows * steps / e2e_s, "unit": "predictions/s",
           "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // steps, "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_s / steps * 1e3,
           "path": ("owc_score(ids, targets in pinned host memory -> scores/idx/paired in host memory)" if d.world == 1 else
                    "per rank owc_encode(ids in pinned host memory) + owc_match(targets in pinned host memory), one packed NCCL "
                    "gather, rank 0 copies all rows' scores/idx/paired to pinned host memory (byte counts are rank 0's)")}

    # ---- strings -> results (host tokenisation included), one rank ----
    e2e_strings = None
    if strings_leg and d.world == 1:
        from owc_b200.pipeline import score_strings_pipelined

        tok = wl.tok
        tok.encode_packed(wl.preds[:4096])
        t0 = time.perf_counter()
        tok.encode_packed(wl.preds)
        t_tok = time.perf_counter() - t0
        zc = eng.encode(wl.cids, wl.ccu, out="device")
        warm = min(rows, 60_000)
        score_strings_pipelined(eng, tok, wl.preds[:warm], zc, targets=wl.targets[:warm], k=TOPK)
        runs = []
        for _ in range(3):                      # host threads make single runs noisy (+/- 40 % seen): median of three
            t1 = time.perf_counter()
            score_strings_pipelined(eng, tok, wl.preds, zc, targets=wl.targets, k=TOPK)
            runs.append(time.perf_counter() - t1)
        t_all = sorted(runs)[1]
        e2e_strings = {"value": rows / t_all, "unit": "predictions/s", "ms_per_step": t_all * 1e3, "tokenize_alone_ms": t_tok * 1e3,
                       "runs_ms": [round(t * 1e3, 1) for t in runs],
                       "path": "strings -> WordPiece ids on a host thread (native C++), overlapped with owc_score of the previous "
                               "row chunk -> results in host memory (owc_b200.pipeline.score_strings_pipelined), one rank"}

    # ---- roofline leg: every launch bracketed by CUDA events on the engine's stream; class names and predictions apart ----
    d.barrier()
    eng.profile_begin()
    if d.world == 1 or d.rank == 0 or wl.sharded_classes:
        cls_ids, cls_cu = wl.cids, wl.ccu
        if wl.sharded_classes:
            per = -(-len(wl.classes) // d.world)
            lo, hi = min(d.rank * per, len(wl.classes)), min((d.rank + 1) * per, len(wl.classes))
            cls_ids, cls_cu = wl.cids[int(wl.ccu[lo]):int(wl.ccu[hi])], (wl.ccu[lo:hi + 1] - wl.ccu[lo]).astype(np.int32)
        eng.encode(cls_ids, cls_cu, out="device")
    prof_cls = eng.profile_end()
    zc = wl.sharded.class_embeddings(eng, wl.cids, wl.ccu, scratch=wl.scratch).wait()
    eng.profile_begin()
    eng.score(wl.ids_dev, wl.pcu, zc, targets=wl.tg_dev, k=TOPK, out="device")
    prof = eng.profile_end()
    lens = np.diff(wl.pcu).astype(np.float64)
    gemm_flops, attn_flops = encoder_flops(arch, lens)
    fused = arch.kind == A.KIND_BERT and arch.hidden == 384 and arch.heads == 12 and arch.inter % 128 == 0 and lens.max() <= 64 \
        and os.environ.get("OWC_FUSED", "1") != "0"
    wide = arch.hidden >= 768 and os.environ.get("OWC_WIDE", "1") != "0"
    gemm_ms = prof["gemm"]["ms"]
    kernel_flops = gemm_flops + (attn_flops if fused else 0.0)   # the fused kernel runs the attention inside the same launch
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
    shapes = ("gemm_qkv", "gemm_o", "gemm_ffn1", "gemm_ffn2")     # breakdown of "gemm" on the wide path: not added to totals
    total_ms = sum(v["ms"] for k, v in prof.items() if k not in shapes)
    n_launch = max(prof["gemm"]["launches"], 1)
    # Wide path with the attention in the QKV projection's epilogue (k_qkv_attn, no attention launches): two kernels share
    # the "gemm" bucket. The dominant one stays k_gemm_wide (O-projection, FFN1, FFN2); k_qkv_attn is reported beside it.
    h_, i_ = arch.hidden, arch.inter
    qkv_share = 3.0 * h_ * h_ / (4.0 * h_ * h_ + 2.0 * h_ * i_)
    qkv_attn_fused = wide and prof["attn"]["launches"] == 0 and prof.get("gemm_qkv", {}).get("launches", 0) > 0
    secondary = None
    if qkv_attn_fused:
        qkv_ms = prof["gemm_qkv"]["ms"]
        q_flops = gemm_flops * qkv_share + attn_flops
        secondary = {"kernel": "k_qkv_attn (QKV projection on tcgen05 with the deferred LayerNorm and the self-attention of the row block in its "
                               "epilogue: q|k|v never reach HBM; the warp-level MMAs of the attention share the tensor pipe with the UMMAs)",
                     "achieved": q_flops / (qkv_ms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": q_flops / (qkv_ms / 1e3) / 1e12 / peak if peak else None,
                     "algorithmic_flops_per_launch": q_flops / max(prof["gemm_qkv"]["launches"], 1),
                     "launches_per_step": prof["gemm_qkv"]["launches"], "avg_launch_ms": qkv_ms / max(prof["gemm_qkv"]["launches"], 1),
                     "share_of_step": qkv_ms / total_ms if total_ms else None}
        kernel_flops = gemm_flops * (1.0 - qkv_share)
        gemm_ms -= qkv_ms
        n_launch = max(n_launch - prof["gemm_qkv"]["launches"], 1)
    achieved = kernel_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic, tinfo = None, {}
    tpath = os.path.join(ROOT, "profiles", "r02_wide_traffic.json" if wide else "r01_fused_traffic.json")
    if (fused or wide) and os.path.exists(tpath):   # dram bytes per launch of the same kernel from the committed ncu --set full capture
        with open(tpath, encoding="utf-8") as f:
            tinfo = json.load(f)
        traffic = tinfo.get("bytes_per_launch")
    kname = ("k_encoder_fused (all encoder layers of a 256-row block pair per cluster in one launch: tcgen05 contractions, attention, GELU, LayerNorm)"
             if fused else
             "k_gemm_wide (persistent CTA-pair tcgen05 contractions O-projection / FFN1 / FFN2 with the LayerNorms folded into their epilogues; "
             "the QKV projection + attention kernel is under `secondary`)"
             if qkv_attn_fused else
             "k_gemm_wide (persistent CTA-pair tcgen05 contractions QKV / O / FFN1 / FFN2 with the LayerNorms folded into their epilogues)"
             if wide else "k_gemm_tc (encoder QKV/O/FFN contractions, tcgen05)")
    roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_source": (tinfo.get("launch") if traffic is not None else None),
                "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "scope": f"this rank's prediction shard ({wl.n_rows} rows, {wl.n_tok} tokens); the class-name launches are timed apart",
                "algorithmic_flops_per_launch": kernel_flops / n_launch, "launches_per_step": n_launch,
                "avg_launch_ms": gemm_ms / n_launch, "share_of_step": gemm_ms / total_ms if total_ms else None,
                "whole_step_tflops": (gemm_flops + attn_flops + 2.0 * wl.n_rows * len(wl.classes) * arch.hidden) / (total_ms / 1e3) / 1e12 if total_ms else None,
                "per_kernel_ms": {k: round(v["ms"], 4) for k, v in prof.items() if k not in shapes or v["launches"]},
                "class_names": {"launches": sum(v["launches"] for k, v in prof_cls.items() if k not in shapes),
                                "ms": round(sum(v["ms"] for k, v in prof_cls.items() if k not in shapes), 4), "tokens": int(wl.ccu[-1])}}
    if secondary is not None:
        roofline["secondary"] = secondary
        roofline["tensor_kernels_together"] = {"achieved": (gemm_flops + attn_flops) / (prof["gemm"]["ms"] / 1e3) / 1e12,
                                               "frac": (gemm_flops + attn_flops) / (prof["gemm"]["ms"] / 1e3) / 1e12 / peak if peak else None,
                                               "share_of_step": prof["gemm"]["ms"] / total_ms if total_ms else None}
    sim_ms = prof["sim_topk"]["ms"]
    sim_in_step = None
    if sim_ms > 0:
        tf = 2.0 * wl.n_rows * len(wl.classes) * arch.hidden / (sim_ms / 1e3) / 1e12
        sim_in_step = {"tflops": tf, "ms": sim_ms, "frac_of_sustained_peak": tf / peak if peak else None,
                       "rows": wl.n_rows, "classes": len(wl.classes), "dim": arch.hidden}
    out = {"value": rows / (ms / 1e3), "unit": "predictions/s", "ms_per_step": ms, "gpu_launches": int(launches),
           "config": config_for(name, rows), "clocks": clocks, "e2e": e2e, "roofline": roofline,
           "detail": {"rows_this_rank": wl.n_rows, "tokens_this_rank": wl.n_tok, "rank0_head_start_tokens": wl.head_start,
                      "class_embeddings": ("equal slices encoded by every rank + one all_gather" if wl.sharded_classes else
                                           "encoded on rank 0, asynchronous broadcast" if d.world > 1 else "encoded once per step"),
                      "results": "one packed [rows, 2k+1] gather to rank 0" if d.world > 1 else "single rank"}}
    if sim_in_step is not None:
        out["sim_gemm_topk_in_step"] = sim_in_step
    if e2e_strings is not None:
        out["e2e_strings"] = e2e_strings
    out["_engine_dim"] = arch.hidden
    wl.close()
    return out


def measure_sweep(d: Dist, steps: int, warmup: int) -> dict:
    """c5: strings of 130 sample sets -> top-5 + paired score per set. The sets are dealt to the ranks by row count; each
    class set is encoded once per rank that needs it; tokenisation overlaps scoring (owc_b200.sweep.score_sweep)."""
    from owc_b200.engine import Engine
    from owc_b200.sweep import score_sweep
    from owc_b200.tokenizer import HostTokenizer
    from owc_b200.weights import random_state_dict

    sw = make_sweep()
    arch, tasks = sw["arch"], sw["tasks"]
    if d.world > 1:   # the ranks share the host's cores: each tokenizer takes its share instead of all of them (as the CLI does)
        os.environ.setdefault("OWC_TOKENIZER_THREADS", str(max(1, (os.cpu_count() or 1) // d.world)))
    tok = HostTokenizer(sw["vocab"], "bert")
    eng = Engine(arch, random_state_dict(arch), device=d.local_rank)
    flat = [(ds, key, len(f["predictions"])) for ds, t in tasks.items() for key, f in t["files"].items()]
    total = sum(n for _, _, n in flat)
    # deal whole sample sets to ranks: largest first onto the least-loaded rank (deterministic, same on every rank)
    loads = [0] * d.world
    mine: dict = {}
    for ds, key, n in sorted(flat, key=lambda x: (-x[2], x[0], x[1])):
        r = loads.index(min(loads))
        loads[r] += n
        if r == d.rank:
            mine.setdefault(ds, {"class_names": tasks[ds]["class_names"], "files": {}})["files"][key] = tasks[ds]["files"][key]
    small = {ds: {"class_names": t["class_names"],
                  "files": {"w": {"predictions": next(iter(t["files"].values()))["predictions"][:512],
                                  "targets": next(iter(t["files"].values()))["targets"][:512]}}} for ds, t in mine.items()}
    score_sweep(eng, tok, small, k=TOPK)            # every class-set size once (kernel variants load on first launch)
    for _ in range(max(warmup - 1, 1)):
        score_sweep(eng, tok, mine, k=TOPK)
    d.barrier()
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = score_sweep(eng, tok, mine, k=TOPK)
        if d.world > 1:   # the report needs every set's mean paired score on rank 0
            means = {f"{ds}/{k}": float(v["paired"].mean()) for ds, t in res.items() for k, v in t.items()}
            gathered = [None] * d.world if d.rank == 0 else None
            d.dist.gather_object(means, gathered, dst=0)
    d.torch.cuda.synchronize()
    dt = d.max(time.perf_counter() - t0) / steps
    s1 = eng.stats()
    eng.close()
    return {"value": total / dt, "unit": "predictions/s", "ms_per_step": dt * 1e3,
            "gpu_launches": int((s1["kernel_launches"] - s0["kernel_launches"]) // steps),
            "config": {"workload": C5_DESC, "arch": "minilm-l6", "rows": total, "sample_sets": len(flat), "topk": TOPK,
                       "partitioning": "whole sample sets dealt to the GPUs by row count (largest first); per-set means gathered on rank 0"},
            "e2e": {"value": total / dt, "unit": "predictions/s",
                    "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // steps,
                    "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // steps,
                    "path": "strings in host memory -> WordPiece ids (host thread, overlapped) -> owc_score per class set -> results in host "
                            "memory; `value` of this record is this same end-to-end figure (the sweep is host-driven by construction)"},
            "detail": {"rows_this_rank": loads[d.rank], "rows_per_rank": loads}}


def run_ours(args) -> None:
    d = Dist()
    torch = d.torch
    peaks, peak_src = load_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=d.dev)
    rows = args.rows or WORKLOADS[args.workload]["rows"]

    head = measure_rows(args.workload, rows, d, args.steps, args.warmup, flush, peaks, peak_src, args.chunk_tokens)
    dim = head.pop("_engine_dim")
    line = {
        "metric": METRIC, "value": head["value"], "unit": "predictions/s", "n_gpus": d.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": head["config"], "clocks": head["clocks"],
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "detail": head["detail"],
    }
    for key in ("e2e_strings", "sim_gemm_topk_in_step"):
        if key in head:
            line[key] = head[key]

    sub = {}
    sub_steps = max(1, min(args.steps, 5))
    for name in [s for s in args.sub.split(",") if s]:
        if name == args.workload or (name == "c2" and d.world > 1):   # c2 is BASELINE's single-GPU configuration
            continue
        try:
            if name == "c5":
                sub[name] = measure_sweep(d, min(sub_steps, 3), 2)
            else:
                r = measure_rows(name, WORKLOADS[name]["rows"], d, sub_steps, 3, flush, peaks, peak_src)
                r.pop("_engine_dim", None)
                sub[name] = r
            sub[name]["steps"] = min(sub_steps, 3) if name == "c5" else sub_steps
        except Exception as exc:   # a sub-record must never take the headline down
            sub[name] = {"error": f"{type(exc).__name__}: {exc}"}
            d.barrier()
    if sub:
        line["sub"] = sub

    if d.rank == 0 and not args.no_sim_leg:
        # north-star second metric: the fused similarity GEMM + top-k kernel alone at a 117k-name vocabulary
        from owc_b200 import arch as A
        from owc_b200.engine import Engine
        from owc_b200.weights import random_state_dict

        eng = Engine(A.MINILM_L6, random_state_dict(A.MINILM_L6), device=d.local_rank)
        burst = float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"]))
        legs = {}
        for h in sorted({384, dim}):
            n_sim, c_sim = 148 * 128 * 2, 117_000
            ms = eng.bench_sim_kernel(n_sim, c_sim, h, TOPK, 5)
            tf = 2.0 * n_sim * c_sim * h / (ms / 1e3) / 1e12
            legs[f"d{h}"] = {"rows": n_sim, "classes": c_sim, "dim": h, "k": TOPK, "ms_per_launch": ms, "tflops": tf, "peak": burst,
                             "frac": tf / burst, "peak_source": f"{peak_src} bf16_tflops (kernel timed alone)",
                             "algorithmic_bytes": 2 * h * (n_sim + c_sim) + 8 * 8 * n_sim}
        eng.close()
        line["sim_gemm_topk"] = legs["d384"]
        if len(legs) > 1:
            line["sim_gemm_topk_wide"] = legs[f"d{dim}"]
    if d.rank == 0 and d.world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(args.workload, min(rows, 10_000))
        r = ref.rate(1, 1, budget_s=25.0, max_rows=min(rows, 10_000), rows_total=rows)
        line["cpu_baseline"] = {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    if d.rank == 0:
        print(json.dumps(line), flush=True)
    d.close()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="TOTAL prediction rows of the headline workload (default: the workload's own size)")
    ap.add_argument("--chunk-tokens", type=int, default=0)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS), help="headline workload; c3 is the configuration the metric is quoted on")
    ap.add_argument("--sub", default="c2,c4,c5", help="comma-separated configs measured beside the headline ('' for none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sim-leg", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: CPU seconds for warm-up + timed steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
