"""Boundary proof: this build's metric functions installed INTO the reference's own registry
(`src.data.metrics`, /root/reference, unmodified) and driven by the reference's own `eval_metrics.main`.

The same sample files are scored twice by the reference's CLI - once untouched, once after
`install_into_reference` - and what the two runs write back and print must agree. Needs /root/reference, so it only
runs in the build container (skipped on the GPU box); no GPU is involved: the engine behind the swapped-in functions is
the oracle-backed stand-in the other host tests use."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/data/metrics"), reason="the reference tree is not on this machine")
def test_reference_cli_with_this_builds_metrics_installed(tmp_path):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_dropin_worker.py"), str(tmp_path)],
                          capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-3000:]
    line = next(ln for ln in proc.stdout.splitlines() if ln.startswith("RESULT "))
    r = json.loads(line[len("RESULT "):])
    assert set(r["patched"]) == {"semantic_similarity", "mean_average_semantic_similarity", "concept_semantic_similarity"}
    assert r["swapped"] and r["modules"] == ["owc_b200.metrics._group"]      # both dicts of the reference's registry
    a, b = r["a"], r["b"]
    for model in ("model-0", "model-1"):
        fa, fb = a["files"][model], b["files"][model]
        assert sorted(fa) == sorted(fb) == sorted(["semantic_similarity", "mean_average_semantic_similarity"] +
                                                  [f"semantic_similarity@{t}" for t in (0.5, 0.6, 0.7, 0.8, 0.9)])
        sa, sb = np.asarray(fa["semantic_similarity"]), np.asarray(fb["semantic_similarity"])
        assert sa.shape == sb.shape and np.abs(sa - sb).max() < 5e-6
        far = np.min(np.abs(sa[None, :] - np.array([.5, .6, .7, .8, .9])[:, None]), axis=0) > 1e-5
        for t in (0.5, 0.6, 0.7, 0.8, 0.9):
            ca, cb = np.asarray(fa[f"semantic_similarity@{t}"]), np.asarray(fb[f"semantic_similarity@{t}"])
            assert np.array_equal(ca[far], cb[far])
        assert np.abs(np.asarray(fa["mean_average_semantic_similarity"])[far] - np.asarray(fb["mean_average_semantic_similarity"])[far]).max() < 1e-6
    # the printed report: same blocks, same numbers to the printed 3 decimals
    assert a["stdout"].strip() and a["stdout"] == b["stdout"]
