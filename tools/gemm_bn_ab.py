"""A/B of the tcgen05 GEMM column-tile width on the mpnet-base shapes (M tokens x N x K), through owc_debug_gemm
(fp32 epilogue) with the GEMM kernel alone bracketed by the library's event profile. Usage: python tools/gemm_bn_ab.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import owc_b200  # noqa: E402,F401
from owc_b200 import arch as A  # noqa: E402
from owc_b200.engine import Engine  # noqa: E402
from owc_b200.weights import random_state_dict  # noqa: E402

arch = A.EncoderArch("tiny", A.KIND_BERT, layers=1, hidden=384, heads=12, inter=1536, vocab=1024, max_pos=64, type_vocab=2, ln_eps=1e-12)
eng = Engine(arch, random_state_dict(arch), device=0)
M = 37888
out = {}
g = torch.Generator(device="cuda").manual_seed(0)
for (n, k) in [(2304, 768), (768, 768), (3072, 768), (768, 3072), (1152, 384), (1536, 384), (384, 1536)]:
    a = torch.randn(M, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.05
    b = torch.randn(n, device="cuda", generator=g)
    for tile in (192, 1192, 1256, 256, 1128):
        if n % (tile % 1000) != 0:
            continue
        for _ in range(2):
            eng.debug_gemm(a, w, b, precision="bf16", tile_n=tile)
        eng.profile_begin()
        for _ in range(3):
            eng.debug_gemm(a, w, b, precision="bf16", tile_n=tile)
        ms = eng.profile_end()["gemm"]["ms"] / 3
        out[f"{M}x{n}x{k} tile {tile}"] = round(2 * M * n * k / ms / 1e9, 1)
print(json.dumps(out, indent=0))
eng.close()
