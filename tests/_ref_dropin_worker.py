"""Worker of tests/test_reference_dropin.py (run in its own process: it patches ``datasets`` and ``sys.modules``).

Imports the REFERENCE's own ``eval_metrics.py`` and ``src.data.metrics`` from /root/reference (shims of SURVEY.md
Appendix A), then runs the reference's ``main`` twice on identical sample files:
  A. untouched: the reference's metric functions over its own ``encode_sentence_bert`` (seeded random-init MiniLM);
  B. after ``owc_b200.metrics.install_into_reference(src.data.metrics)``: this build's metric functions, whose engine
     is an oracle-backed stand-in (no GPU in the build container; the CUDA engine sits behind the same two methods).
Prints one JSON object with what each run wrote and printed.
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

from oracle.make_golden import REF, install_reference_shims  # noqa: E402

install_reference_shims()
spec = importlib.util.spec_from_file_location("ref_eval_metrics", os.path.join(REF, "eval_metrics.py"))
ref_cli = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_cli)
import src.data.metrics as ref_metrics  # noqa: E402
import src.data.pipelines.text._text as ref_text  # noqa: E402

import owc_b200  # noqa: E402,F401
from oracle.hf_models import build_hf_tokenizer, build_model  # noqa: E402
from owc_b200 import arch as A, metrics as our_metrics, synth, text as our_text  # noqa: E402
from owc_b200.tokenizer import HostTokenizer  # noqa: E402
from test_host import OracleEngine  # noqa: E402

work = sys.argv[1]
arch = A.MINILM_L6
vocab = synth.build_vocab("bert", arch.vocab)
words = synth.vocab_words(vocab)
cls = synth.real_class_names("caltech101")
preds, targets = synth.make_predictions(96, cls, words, seed=99)
model = build_model(arch)


def write_logs(root):
    for m, (lo, hi) in enumerate([(0, 60), (60, 96)]):
        run = os.path.join(root, "caltech101", f"model-{m}")
        os.makedirs(run)
        with open(os.path.join(run, "r_samples_caltech101.jsonl"), "w", encoding="utf-8") as f:
            for i in range(lo, hi):
                resps = [preds[i]] if m == 0 else [[preds[(i + 1) % 96], preds[i]]]     # model-1 is a multi-round file
                f.write(json.dumps({"doc_id": i, "target": cls[targets[i]], "filtered_resps": resps, "doc": {"k": [1, 2]}}) + "\n")


def run(root):
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        ref_cli.main(Namespace(input=root, metrics="semantic_similarity,mean_average_semantic_similarity", seed=1234,
                               log_level="ERROR"))
    files = {}
    for m in range(2):
        df = pd.read_json(os.path.join(root, "caltech101", f"model-{m}", "r_samples_caltech101.jsonl"), lines=True)
        files[f"model-{m}"] = {c: df[c].tolist() for c in df.columns if c.startswith("semantic") or c.startswith("mean_")}
    return {"stdout": out.getvalue(), "files": files}


a_root, b_root = os.path.join(work, "a", "logs"), os.path.join(work, "b", "logs")
write_logs(a_root)
write_logs(b_root)

# A: the reference, untouched (its lazy from_pretrained is skipped by pre-seeding its module globals, SURVEY 8c shim 4)
ref_text.sentence_bert_model = model
ref_text.sentence_bert_processor = build_hf_tokenizer(arch, vocab)
before = {k: ref_metrics.METRICS[k].group_fn for k in ("semantic_similarity", "mean_average_semantic_similarity")}
res_a = run(a_root)

# B: this build's functions swapped into the reference's registry
patched = our_metrics.install_into_reference(ref_metrics)


class _B:   # what OracleEngine needs of a test bundle
    pass


b = _B()
b.arch, b.sd = arch, {k: v.detach() for k, v in model.state_dict().items()}
our_text.sentence_bert_model = OracleEngine(b)
our_text.sentence_bert_processor = HostTokenizer(vocab, "bert")
ref_text.sentence_bert_model = None      # the reference's own operator must not be reachable any more
ref_text.sentence_bert_processor = None
res_b = run(b_root)
swapped = all(ref_metrics.METRICS[k].group_fn is not before[k] and ref_metrics.AGGREGATIONS[k].builder_fn is ref_metrics.METRICS[k].group_fn
              for k in before)
print("RESULT " + json.dumps({"patched": patched, "swapped": swapped, "a": res_a, "b": res_b,
                              "modules": [ref_metrics.METRICS["semantic_similarity"].group_fn.__module__]},
                             default=lambda o: float(o) if isinstance(o, np.floating) else str(o)))
