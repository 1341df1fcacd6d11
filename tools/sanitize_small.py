"""Small fused-encoder + scoring run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import owc_b200  # noqa
from owc_b200 import arch as A, synth
from owc_b200.engine import Engine
from owc_b200.tokenizer import HostTokenizer
from owc_b200.weights import random_state_dict

arch = A.EncoderArch("bert-2", A.KIND_BERT, layers=2, hidden=384, heads=12, inter=1536, vocab=2048, max_pos=512, type_vocab=2, ln_eps=1e-12)
vocab = synth.build_vocab("bert", arch.vocab)
words = synth.vocab_words(vocab)
cls = synth.make_class_names(40, words)
preds, targets = synth.make_predictions(700, cls, words)
tok = HostTokenizer(vocab, "bert")
eng = Engine(arch, random_state_dict(arch), device=0)
cids, ccu = tok.encode_packed(cls)
zc = eng.encode(cids, ccu, out="device")
pids, pcu = tok.encode_packed(preds + [" ".join(words[:40]), "a"])
res = eng.score(pids, pcu, zc, targets=np.asarray(targets + [0, 1], np.int32), k=5)
print("ok", res["scores"].shape, float(res["paired"].mean()))
eng.close()
