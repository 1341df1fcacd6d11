"""Print the in-kernel timeline of the 1-CTA tcgen05 GEMM (debug stamps). Usage: gemm_timeline.py m n k [gelu]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import owc_b200  # noqa
from owc_b200 import _cabi, arch as A
from owc_b200.engine import Engine
from owc_b200.weights import random_state_dict
m, n, k = (int(x) for x in sys.argv[1:4])
gelu = int(sys.argv[4]) if len(sys.argv) > 4 else 0
arch = A.EncoderArch("tiny", A.KIND_BERT, 1, 256, 8, 512, 512, 64, 2, 1e-12)
eng = Engine(arch, random_state_dict(arch), device=0)
ctas = (m + 127) // 128
st = np.zeros((ctas, 8), dtype=np.int64)
_cabi.check(eng._lib.owc_debug_gemm_timeline(eng._ctx, m, n, k, gelu, st.ctypes.data, ctas))
rel = st[:, 1:7] - st[:, [0]]
names = ["setup done", "first operands", "last MMA issued", "first acc ready", "epilogue done", "all warps done"]
print(f"m={m} n={n} k={k} gelu={gelu}: {ctas} CTAs; cycles since CTA entry (min / median / max)")
for i, nm in enumerate(names):
    col = rel[:, i]
    print(f"  {nm:18s} {col.min():8d} {int(np.median(col)):8d} {col.max():8d}")
ms = C.c_float(0)
_cabi.check(eng._lib.owc_debug_gemm_loop(eng._ctx, m, n, k, gelu, 50, C.byref(ms)))
print(f"  50 back-to-back launches: {ms.value * 1e3:.2f} us per launch (in-kernel {int(np.median(rel[:, 4])) / 1.965e3:.2f} us at 1.965 GHz)")
gt = st[:, 7]
print(f"  CTA start skew (globaltimer): {int(gt.max() - gt.min())} ns")
eng.close()
