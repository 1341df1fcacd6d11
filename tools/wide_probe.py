"""Where the wide encoder path spends its time: per-shape event times of the persistent GEMMs, the in-kernel wait counters
of the instrumented build (OWC_WIDE_DBG=1) and the attention / embedding / pooling kernels, on c3-shaped rows.
Usage: python tools/wide_probe.py [rows] [waves,waves,...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
waves_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8]

import torch  # noqa: E402

import owc_b200  # noqa: E402,F401
from owc_b200 import arch as A, synth  # noqa: E402
from owc_b200.engine import Engine  # noqa: E402
from owc_b200.tokenizer import HostTokenizer  # noqa: E402
from owc_b200.weights import random_state_dict  # noqa: E402

arch = A.MPNET_BASE
vocab = synth.build_vocab("mpnet", arch.vocab)
words = synth.vocab_words(vocab)
cls = synth.make_class_names(1000, words, seed=1234)
preds, _ = synth.make_predictions(rows, cls, words, seed=1234)
tok = HostTokenizer(vocab, "mpnet")
ids, cu = tok.encode_packed(preds)
ids_d = torch.from_numpy(ids).cuda()
n_tok = int(cu[-1])
sd = random_state_dict(arch)
h, inter = arch.hidden, arch.inter
shape_flops = {"gemm_qkv": 2.0 * h * 3 * h, "gemm_o": 2.0 * h * h, "gemm_ffn1": 2.0 * h * inter, "gemm_ffn2": 2.0 * h * inter}
names = ["fold (QKV)", "fold+gelu (FFN1)", "res+stats (O)", "res+stats deep (FFN2)"]
for waves in waves_list:
    for dbg in (0, 1):
        os.environ["OWC_WIDE_WAVES"] = str(waves)
        os.environ["OWC_WIDE_DBG"] = str(dbg)
        eng = Engine(arch, sd, device=0)
        for _ in range(2):
            eng.encode(ids_d, cu, out="device")
        torch.cuda.synchronize()
        if dbg:
            eng.wide_counters(reset=True)
            eng.profile_begin()
            eng.encode(ids_d, cu, out="device")
            prof = eng.profile_end()
            c = eng.wide_counters()
            print(f"  instrumented build, OWC_WIDE_EXP={os.environ.get('OWC_WIDE_EXP', '0')}: " + ", ".join(
                f"{k} {v['ms']:.2f} ms ({arch.layers * shape_flops[k] * n_tok / v['ms'] / 1e9:.0f} TF/s)" for k, v in prof.items() if k in shape_flops))
            if c[4][2]:
                ring, acc, tot, dtf, dqf, dtot, n = (int(x) for x in c[4][:7])
                aqf, atbl, atot = (int(x) for x in c[5][:3])
                units = max(int(c[4][7]), 1)
                phs = [int(x) for x in c[5][3:8]]
                print(f"  k_qkv_attn attention warp 0: {units} units, clk per unit: fragment loads + release {phs[0] / units:.0f}, QK {phs[1] / units:.0f}, softmax {phs[2] / units:.0f}, "
                      f"PV {phs[3] / units:.0f}, output stores {phs[4] / units:.0f}, loop {(atot - aqf - atbl - sum(phs)) / units:.0f}", flush=True)
                tiles = n_tok / 100.0   # only for scale
                print(f"  k_qkv_attn: MMA warp ring wait {ring / tot:6.1%}  accumulator wait {acc / tot:6.1%}  issue {1 - (ring + acc) / tot:6.1%} | "
                      f"drain warp 0: accumulator wait {dtf / dtot:6.1%}  q|k|v buffer wait {dqf / dtot:6.1%}  work {1 - (dtf + dqf) / dtot:6.1%} | "
                      f"attention warp 0: rows wait {aqf / atot:6.1%}  tables {atbl / atot:6.1%}  work {1 - (aqf + atbl) / atot:6.1%} | "
                      f"clk per leader CTA and launch {tot / max(n, 1):.0f} (leader CTAs x launches {n})", flush=True)
            for k, nm in enumerate(names):
                ring, acc, tot, etf, etot, n = c[k][:6]
                if tot:
                    print(f"  waves={waves} {nm:22s} MMA warp: ring wait {ring / tot:6.1%}  accumulator wait {acc / tot:6.1%}  "
                          f"issue {1 - (ring + acc) / tot:6.1%} | epilogue warp waits for accumulators {etf / max(etot, 1):6.1%} "
                          f"(leader CTAs x launches {n})", flush=True)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                eng.encode(ids_d, cu, out="device")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            eng.profile_begin()
            eng.encode(ids_d, cu, out="device")
            prof = eng.profile_end()
            line = {"waves": waves, "rows": rows, "tokens": n_tok, "encode_ms": round(ms, 3),
                    "tflops_whole_encode": round(arch.layers * sum(shape_flops.values()) * n_tok / ms / 1e9, 1)}
            for k, v in prof.items():
                if v["launches"]:
                    line[k] = {"ms": round(v["ms"], 3), "n": v["launches"]}
                    if k in shape_flops:
                        line[k]["tflops"] = round(arch.layers * shape_flops[k] * n_tok / v["ms"] / 1e9, 1)
            print(json.dumps(line), flush=True)
        eng.close()
