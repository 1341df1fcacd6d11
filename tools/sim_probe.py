"""Run the fused similarity+top-k kernel alone (for ncu / quick timing). Usage: sim_probe.py rows classes dim k iters"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import owc_b200  # noqa
from owc_b200 import arch as A
from owc_b200.engine import Engine
from owc_b200.weights import random_state_dict
rows, classes, dim, k, iters = (int(x) for x in (sys.argv[1:6] + ["18944", "20000", "384", "5", "3"][len(sys.argv) - 1:]))
arch = A.EncoderArch("tiny", A.KIND_BERT, 1, 256, 8, 512, 512, 64, 2, 1e-12)
eng = Engine(arch, random_state_dict(arch), device=0)
ms = eng.bench_sim_kernel(rows, classes, dim, k, iters)
print(f"rows={rows} classes={classes} dim={dim} k={k}: {ms:.4f} ms/launch, {2.0*rows*classes*dim/ms/1e9:.1f} TFLOP/s")
eng.close()
