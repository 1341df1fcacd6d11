"""Chunk schedule of the strings -> results pipeline (`pipeline.score_strings_pipelined`) on the c2 workload:
ms per 100 k strings for several (first_chunk_rows, chunk_rows). Usage: python tools/pipeline_tune.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import owc_b200  # noqa: E402,F401
from owc_b200 import arch as A, synth  # noqa: E402
from owc_b200.engine import Engine  # noqa: E402
from owc_b200.pipeline import score_strings_pipelined  # noqa: E402
from owc_b200.tokenizer import HostTokenizer  # noqa: E402
from owc_b200.weights import random_state_dict  # noqa: E402

arch = A.MINILM_L6
vocab = synth.build_vocab("bert", arch.vocab)
words = synth.vocab_words(vocab)
tok = HostTokenizer(vocab, "bert")
eng = Engine(arch, random_state_dict(arch), device=0)
names = synth.real_class_names("food101")
preds, targets = synth.make_predictions(100_000, names, words, seed=1)
ids, cu = tok.encode_packed(names)
zc = eng.encode(ids, cu, out="device")
res = {}
for first, chunk in [(20000, 20000), (8000, 30000), (10000, 36000), (6000, 18000), (5000, 24000), (12000, 44000), (4000, 32000)]:
    score_strings_pipelined(eng, tok, preds[:30000], zc, targets=targets[:30000], k=5, chunk_rows=chunk, first_chunk_rows=first)
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        score_strings_pipelined(eng, tok, preds, zc, targets=targets, k=5, chunk_rows=chunk, first_chunk_rows=first)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    res[f"{first}/{chunk}"] = round(best, 2)
print(json.dumps(res))
eng.close()
