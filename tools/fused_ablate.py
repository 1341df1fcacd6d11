"""Ablation timing of the fused encoder kernel: which role bounds it. Results are WRONG when a bit is set."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import owc_b200  # noqa
from owc_b200 import arch as A, synth
from owc_b200.engine import Engine
from owc_b200.tokenizer import HostTokenizer
from owc_b200.weights import random_state_dict

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 74 * 2 * 15 * 4
arch = A.MINILM_L6
sd = random_state_dict(arch)
vocab = synth.build_vocab("bert", arch.vocab)
words = synth.vocab_words(vocab)
classes = synth.real_class_names("food101")
preds, _ = synth.make_predictions(rows, classes, words, seed=1234)
tok = HostTokenizer(vocab, "bert")
ids, cu = tok.encode_packed(preds)
ids_d = torch.from_numpy(ids).cuda()
ntok = int(cu[-1])
flops = arch.layers * 2.0 * (4 * 384 * 384 + 2 * 384 * 1536) * ntok
print(f"{rows} rows, {ntok} tokens, {flops/1e12:.3f} TFLOP of contractions", flush=True)
os.environ["OWC_FUSED"] = "1"
e = Engine(arch, sd, device=0)
for ab in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,4,8,3,7,15".split(","))]:
    os.environ["OWC_DEBUG_HOOKS"] = "1"
    os.environ["OWC_FUSED_ABLATE"] = str(ab)
    for _ in range(2):
        e.encode(ids_d, cu, out="device")
    torch.cuda.synchronize()
    e.profile_begin()
    for _ in range(3):
        e.encode(ids_d, cu, out="device")
    p = e.profile_end()
    ms = p["gemm"]["ms"] / 3
    print(f"ablate={ab:2d}: fused kernel {ms:.3f} ms  -> {flops / ms / 1e9:.0f} TFLOP/s  ({p['gemm']['launches'] // 3} launches; embed {p['embed_ln']['ms']/3:.3f} pool {p['pool_l2']['ms']/3:.3f})", flush=True)
e.close()
