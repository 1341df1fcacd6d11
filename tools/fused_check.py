"""GPU check of the block-resident fused encoder (enc_fused.cuh) against the per-GEMM path and the oracle,
plus a quick timing and the in-kernel timeline. Usage: python tools/fused_check.py [--rows N] [--no-oracle]"""

import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_engine(arch, sd, fused: int, dbg: int = 0):
    from owc_b200.engine import Engine

    os.environ["OWC_FUSED"] = str(fused)
    os.environ["OWC_FUSED_DBG"] = str(dbg)
    return Engine(arch, sd, device=0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20000)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--ablate", type=int, default=0)
    args = ap.parse_args()

    import torch

    import owc_b200  # noqa: F401
    from owc_b200 import arch as A
    from owc_b200 import synth
    from owc_b200.tokenizer import HostTokenizer
    from owc_b200.weights import random_state_dict

    arch = A.MINILM_L6
    if args.layers != 6:
        arch = A.EncoderArch("minilm-short", A.KIND_BERT, layers=args.layers, hidden=384, heads=12, inter=1536,
                             vocab=arch.vocab, max_pos=512, type_vocab=2, ln_eps=1e-12)
    sd = random_state_dict(arch)
    vocab = synth.build_vocab("bert", arch.vocab)
    words = synth.vocab_words(vocab)
    classes = synth.real_class_names("food101")
    preds, _ = synth.make_predictions(args.rows, classes, words, seed=1234)
    # ragged extras: long (up to 64 tokens) and tiny strings
    rng = np.random.default_rng(0)
    extra = [" ".join(rng.choice(words, size=int(n))) for n in rng.integers(1, 24, size=300)]
    tok = HostTokenizer(vocab, "bert")
    texts = classes + extra + preds
    ids, cu = tok.encode_packed(texts)
    lens = np.diff(cu)
    keep = lens <= 64
    if not keep.all():
        texts = [t for t, k in zip(texts, keep) if k]
        ids, cu = tok.encode_packed(texts)
        lens = np.diff(cu)
    print(f"{len(texts)} sequences, {int(cu[-1])} tokens, len min/mean/max = {lens.min()}/{lens.mean():.2f}/{lens.max()}", flush=True)

    e0 = make_engine(arch, sd, 0)
    z0 = e0.encode(ids, cu)
    os.environ["OWC_DEBUG_HOOKS"] = "1"
    os.environ["OWC_FUSED_ABLATE"] = str(args.ablate)
    e1 = make_engine(arch, sd, 1, dbg=1)
    t0 = time.perf_counter()
    z1 = e1.encode(ids, cu)
    print(f"fused first call {time.perf_counter() - t0:.3f} s", flush=True)
    d = np.abs(z1 - z0)
    print(f"fused vs per-GEMM path: max|dz| = {d.max():.3e}  mean = {d.mean():.3e}  nan = {int(np.isnan(z1).sum())}", flush=True)
    bad = np.nonzero(d.max(axis=1) > 5e-3)[0]
    if bad.size:
        print("rows off:", bad[:20], "lens", lens[bad[:20]], "of", bad.size)
    if not args.no_oracle:
        from oracle import encoder_ref
        from oracle.hf_models import build_hf_tokenizer

        hf_tok = build_hf_tokenizer(arch, vocab)
        n = min(len(texts), 600)
        zr = encoder_ref.encode_strings(sd, arch, hf_tok, texts[:n]).numpy()
        print(f"vs oracle ({n} rows): fused max|dz| = {np.abs(z1[:n] - zr).max():.3e}   per-GEMM max|dz| = {np.abs(z0[:n] - zr).max():.3e}", flush=True)

    # timing, ids resident on the device
    ids_d = torch.from_numpy(ids).cuda()
    for name, e in (("per-GEMM", e0), ("fused", e1)):
        for _ in range(2):
            e.encode(ids_d, cu, out="device")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            e.encode(ids_d, cu, out="device")
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{name}: {ms:.3f} ms per encode of {len(texts)} rows -> {len(texts) / ms * 1e3:.0f} rows/s", flush=True)
    e1.profile_begin()
    e1.encode(ids_d, cu, out="device")
    print("fused profile:", {k: (round(v["ms"], 3), v["launches"]) for k, v in e1.profile_end().items() if v["launches"]})
    st = (C.c_int64 * 2048)()
    from owc_b200 import _cabi

    _cabi.check(e1._lib.owc_debug_fused_stamps(e1._enc, st))
    s = np.array(st[:], dtype=np.int64)
    n = 2 + 4 * arch.layers
    if s[0] > 0:
        rel = (s[:n] - s[0])
        print("timeline (clk since block start): load", rel[1])
        names = ["attn", "ln1", "ffn", "ln2"]
        for l in range(arch.layers):
            seg = s[2 + 4 * l: 6 + 4 * l] - s[1 + 4 * l: 5 + 4 * l]
            print(f"  layer {l}: " + "  ".join(f"{nm} {int(v)}" for nm, v in zip(names, seg)))
        print("  block total clk", int(rel[n - 1]))
        if s[31] > s[30]:
            print(f"  globaltimer span of the stamped region: {int(s[31] - s[30])} ns")
        simt = ["wait t1_full(attn)", "E_Q ld", "E_Q cvt/store", "bar1", "attention", "bar2", "LN wait accx", "LN pass A",
                "LN stats", "LN pass B", "FFN wait t1_full", "FFN ld", "FFN gelu", "FFN wait f_free", "FFN store", "other"]
        mma = ["wait xb_ready", "wait t1_free", "wait ring full", "issue", "wait ao_full", "wait f_full"]
        print("SIMT thread 64 (clk, all blocks of CTA 0):")
        for nm, v in zip(simt, s[32:48]):
            print(f"    {nm:22s} {int(v):>10d}")
        print("MMA thread (clk):")
        for nm, v in zip(mma, s[48:54]):
            print(f"    {nm:22s} {int(v):>10d}")
        if s[55] > 0:
            print(f"ring: {int(s[58])} slots consumed; {int(s[55])} had to be waited for: TMA issue->landed latency mean {int(s[54] / s[55])} clk, "
                  f"max {int(s[56])}; mean age of a slot when consumed {int(s[57] / max(s[58], 1))} clk")
        print("phase sums of other observers (clk): columns = leader w0, w15, w7, peer w0, w15, w7")
        for i, nm in enumerate(simt):
            print(f"    {nm:22s} " + " ".join(f"{int(s[1024 + k * 16 + i]):>9d}" for k in range(6)))
        # merged event trace of layer 1 (SIMT thread 64 ids < 100, MMA lane ids >= 100)
        ev = []
        for base in (64, 512):
            for i in range(220):
                eid, clk = int(s[base + 2 * i]), int(s[base + 2 * i + 1])
                if clk > 0:
                    ev.append((clk, eid))
        ev.sort()
        if ev:
            names = {0: "S t1_full", 1: "S drained", 2: "S qkv ready", 3: "S attn done", 4: "S head done", 10: "S t1_full(F)",
                     11: "S drained(F)", 12: "S gelu done", 13: "S fc stored", 20: "S accx_full", 21: "S LN done",
                     100: "  M t1_free seen", 101: "  M T1 issued", 102: "  M ao_full seen", 103: "  M X kblock issued",
                     104: "  M f_full seen"}
            t0 = ev[0][0]
            prev = t0
            print("event trace, layer 1 (clk since first event, delta):")
            for clk, eid in ev[:(170 if os.environ.get('OWC_TRACE') else 0)]:
                print(f"    {clk - t0:>8d} {clk - prev:>6d}  {names.get(eid, eid)}")
                prev = clk
    e0.close()
    e1.close()


if __name__ == "__main__":
    main()
