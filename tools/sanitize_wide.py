"""Small wide-path run (mpnet-base width: persistent deferred-LayerNorm GEMMs, QKV projection + attention kernel, matcher)
for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import owc_b200  # noqa
from owc_b200 import arch as A, synth
from owc_b200.engine import Engine
from owc_b200.tokenizer import HostTokenizer
from owc_b200.weights import random_state_dict

arch = A.EncoderArch("mpnet-2", A.KIND_MPNET, layers=2, hidden=768, heads=12, inter=3072, vocab=2048, max_pos=514, type_vocab=0,
                     ln_eps=1e-5, rel_buckets=32, rel_max_distance=128, pad_id=1)
vocab = synth.build_vocab("mpnet", arch.vocab)
words = synth.vocab_words(vocab)
cls = synth.make_class_names(40, words)
preds, targets = synth.make_predictions(500, cls, words)
tok = HostTokenizer(vocab, "mpnet")
eng = Engine(arch, random_state_dict(arch), device=0)
cids, ccu = tok.encode_packed(cls)
zc = eng.encode(cids, ccu, out="device")
# short labels + a 40-token and a 60-token row (32- and 64-row attention tiles: the kernel variant with out-of-line long tiles)
# + a one-word row; with OWC_SANITIZE_SHORT=1 only short rows (the variant every label set of the benchmark runs)
extra = ["a b", "c", "a"] if os.environ.get("OWC_SANITIZE_SHORT") == "1" else [" ".join(words[:30]), " ".join(words[:50]), "a"]
pids, pcu = tok.encode_packed(preds + extra)
res = eng.score(pids, pcu, zc, targets=np.asarray(targets + [0, 1, 2], np.int32), k=5)
print("ok", res["scores"].shape, float(res["paired"].mean()), "launches", eng.stats()["kernel_launches"])
eng.close()
