"""Chunk schedule of the strings -> results pipeline (`pipeline.score_strings_pipelined`): ms per run for several
(first_chunk_rows, chunk_rows) on the c2 workload (100 k strings, MiniLM) or, with `c3`, on the c3 workload (1 M strings,
mpnet-base arch, 1000 names). Usage: python tools/pipeline_tune.py [c2|c3]"""
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

wide = len(sys.argv) > 1 and sys.argv[1] == "c3"
arch = A.MPNET_BASE if wide else A.MINILM_L6
kind = "mpnet" if wide else "bert"
vocab = synth.build_vocab(kind, arch.vocab)
words = synth.vocab_words(vocab)
tok = HostTokenizer(vocab, kind)
eng = Engine(arch, random_state_dict(arch), device=0)
names = synth.make_class_names(1000, words, seed=1234) if wide else synth.real_class_names("food101")
preds, targets = synth.make_predictions(1_000_000 if wide else 100_000, names, words, seed=1)
ids, cu = tok.encode_packed(names)
zc = eng.encode(ids, cu, out="device")
res = {}
grid = ([(20000, 20000), (20000, 50000), (20000, 100000), (10000, 60000), (30000, 150000), (None, None)] if wide else
        [(20000, 20000), (8000, 30000), (10000, 36000), (6000, 18000), (5000, 24000), (12000, 44000), (4000, 32000), (None, None)])
for first, chunk in grid:
    kw = {} if chunk is None else {"chunk_rows": chunk, "first_chunk_rows": first}   # (None, None): the library's default schedule
    score_strings_pipelined(eng, tok, preds[:30000], zc, targets=targets[:30000], k=5, **kw)
    times = []
    for _ in range(3 if wide else 5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        score_strings_pipelined(eng, tok, preds, zc, targets=targets, k=5, **kw)
        times.append(round((time.perf_counter() - t0) * 1e3, 1))
    res[f"{first}/{chunk}"] = times
print(json.dumps(res))
eng.close()
