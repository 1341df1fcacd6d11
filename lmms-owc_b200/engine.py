"""Python handle on the C-ABI library: one :class:`Engine` = one ``owc_ctx`` + one ``owc_encoder``.

PyTorch is used only as plumbing here: it owns device tensors that outlive a call (class
embeddings, sharded results) and the CUDA stream the library enqueues on, so ``torch.cuda.Event``
and ``torch.distributed`` see the work. All arithmetic is in ``libowc_b200.so``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .arch import EncoderArch, pack_weights

__all__ = ["Engine", "PRECISIONS"]

PRECISIONS = {"bf16": _cabi.PREC_BF16, "fp32": _cabi.PREC_FP32}


def _torch():
    import torch

    return torch


def _as_arg(x, dtype, device_index: int):
    """(pointer, space, keepalive) for a numpy array or torch tensor of the given numpy dtype."""
    torch = _torch()
    if isinstance(x, torch.Tensor):
        want = {np.int32: torch.int32, np.float32: torch.float32}[dtype]
        if x.dtype != want or not x.is_contiguous():
            x = x.to(want).contiguous()
        if x.is_cuda:
            if x.device.index != device_index:
                raise _cabi.OwcError(f"tensor on {x.device}, engine on cuda:{device_index}")
            return x.data_ptr(), _cabi.MEM_DEVICE, x
        return x.data_ptr(), _cabi.MEM_HOST, x
    a = np.ascontiguousarray(x, dtype=dtype)
    return a.ctypes.data, _cabi.MEM_HOST, a


class Engine:
    """Encoder + scorer on one GPU. Not thread-safe (one call at a time, like the reference op)."""

    def __init__(self, arch: EncoderArch, state_dict: dict, device: int = 0, chunk_tokens: int = 0,
                 precision: str = "bf16") -> None:
        torch = _torch()
        self._lib = _cabi.lib()
        if not torch.cuda.is_available():
            raise _cabi.OwcError("no CUDA device: the scoring path has no CPU fallback")
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.arch = arch
        self.precision = precision
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        self.stream = torch.cuda.Stream(device=self.torch_device)
        self._ctx = C.c_void_p()
        self._enc = C.c_void_p()
        _cabi.check(self._lib.owc_ctx_create(self.device, C.c_void_p(self.stream.cuda_stream), int(chunk_tokens),
                                             C.byref(self._ctx)))
        blob = pack_weights(arch, state_dict)
        carch = _cabi.OwcArch(arch.kind, arch.layers, arch.hidden, arch.heads, arch.inter, arch.vocab, arch.max_pos,
                              arch.type_vocab, arch.pad_id, arch.ln_eps)
        _cabi.check(self._lib.owc_encoder_load(self._ctx, C.byref(carch), blob.ctypes.data, blob.size,
                                               C.byref(self._enc)))

    # ---- lifetime ----
    def close(self) -> None:
        if getattr(self, "_enc", None) is not None and self._enc:
            self._lib.owc_encoder_destroy(self._enc)
            self._enc = C.c_void_p()
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.owc_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # ---- stream plumbing ----
    def _enter(self) -> None:
        torch = _torch()
        self.stream.wait_stream(torch.cuda.current_stream(self.torch_device))

    def _leave(self) -> None:
        torch = _torch()
        torch.cuda.current_stream(self.torch_device).wait_stream(self.stream)

    def sync(self) -> None:
        _cabi.check(self._lib.owc_ctx_sync(self._ctx))

    def stats(self) -> dict:
        s = _cabi.OwcStats()
        _cabi.check(self._lib.owc_ctx_stats(self._ctx, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in _cabi.OwcStats._fields_}

    def profile_begin(self) -> None:
        """Start bracketing every kernel launch with CUDA events (roofline leg only)."""
        _cabi.check(self._lib.owc_ctx_profile_begin(self._ctx))

    def profile_end(self) -> dict:
        """Stop profiling; ``{category: {"ms": summed device time, "launches": n}}``."""
        p = _cabi.OwcProfile()
        _cabi.check(self._lib.owc_ctx_profile_end(self._ctx, C.byref(p)))
        return {name: {"ms": float(p.ms[i]), "launches": int(p.launches[i])} for i, name in enumerate(_cabi.KCATS)}

    def _prec(self, precision: str | None) -> int:
        return PRECISIONS[precision or self.precision]

    # ---- operations ----
    def encode(self, ids, cu_seqlens, out: str = "host", precision: str | None = None):
        """Packed token ids -> unit-norm embeddings ``[n_seq, H]`` (numpy if ``out='host'``, a CUDA
        tensor if ``out='device'``). ``owc_encode``; reference: ``_text.py:196-202``."""
        torch = _torch()
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n_seq = int(cu.size) - 1
        h = self.arch.hidden
        if n_seq <= 0:
            return np.zeros((0, h), np.float32) if out == "host" else torch.zeros(0, h, device=self.torch_device)
        p_ids, sp_ids, keep = _as_arg(ids, np.int32, self.device)
        self._enter()
        if out == "host":
            res = np.empty((n_seq, h), dtype=np.float32)
            _cabi.check(self._lib.owc_encode(self._enc, p_ids, sp_ids, cu.ctypes.data, n_seq, self._prec(precision),
                                             res.ctypes.data, _cabi.MEM_HOST))
        else:
            res = torch.empty(n_seq, h, dtype=torch.float32, device=self.torch_device)
            _cabi.check(self._lib.owc_encode(self._enc, p_ids, sp_ids, cu.ctypes.data, n_seq, self._prec(precision),
                                             res.data_ptr(), _cabi.MEM_DEVICE))
            self._leave()
        del keep
        return res

    def pair_dot(self, a, b, a_index=None, b_index=None) -> np.ndarray:
        """``out[i] = <a[a_index[i]], b[b_index[i]]>`` (identity index when ``None``), fp32 numpy ``[n]``.

        ``a`` / ``b`` are host arrays, or both CUDA tensors straight from ``encode(..., out='device')`` - then only the
        index vectors go up and the ``n`` scores come down. ``owc_pair_dot``; reference: the ``torch.bmm`` at
        ``_group.py:541-544``."""
        torch = _torch()
        on_dev = isinstance(a, torch.Tensor) and a.is_cuda
        if on_dev != (isinstance(b, torch.Tensor) and b.is_cuda):
            raise ValueError("pair_dot: a and b must live in the same memory space")
        if not on_dev:
            a = np.ascontiguousarray(a, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
        if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
            raise ValueError("pair_dot needs [na, dim] and [nb, dim] arrays")
        ai = None if a_index is None else np.ascontiguousarray(a_index, dtype=np.int32)
        bi = None if b_index is None else np.ascontiguousarray(b_index, dtype=np.int32)
        n = ai.size if ai is not None else a.shape[0]
        if (bi.size if bi is not None else b.shape[0]) != n:
            raise ValueError("pair_dot: both sides must describe the same number of rows")
        if ai is not None and n and (ai.min() < 0 or ai.max() >= a.shape[0]):
            raise IndexError("a_index out of range")
        if bi is not None and n and (bi.min() < 0 or bi.max() >= b.shape[0]):
            raise IndexError("b_index out of range")
        if not n:
            return np.empty(0, dtype=np.float32)
        if on_dev:
            p_a, _, keep_a = _as_arg(a, np.float32, self.device)
            p_b, _, keep_b = _as_arg(b, np.float32, self.device)
            dev = self.torch_device
            ai_d = None if ai is None else torch.from_numpy(ai).to(dev)
            bi_d = None if bi is None else torch.from_numpy(bi).to(dev)
            out_d = torch.empty(n, dtype=torch.float32, device=dev)
            self._enter()
            _cabi.check(self._lib.owc_pair_dot(
                self._ctx, p_a, keep_a.shape[0], ai_d.data_ptr() if ai_d is not None else None, p_b, keep_b.shape[0],
                bi_d.data_ptr() if bi_d is not None else None, n, keep_a.shape[1], _cabi.MEM_DEVICE, out_d.data_ptr()))
            self._leave()
            return out_d.cpu().numpy()
        out = np.empty(n, dtype=np.float32)
        _cabi.check(self._lib.owc_pair_dot(
            self._ctx, a.ctypes.data, a.shape[0], ai.ctypes.data if ai is not None else None, b.ctypes.data,
            b.shape[0], bi.ctypes.data if bi is not None else None, n, a.shape[1], _cabi.MEM_HOST,
            out.ctypes.data))
        return out

    def sim_topk(self, zp, zc, k: int, precision: str | None = None):
        """Row-wise top-k of ``zp @ zc.T`` without materialising it. ``owc_sim_topk``.

        Host arrays in -> numpy ``(scores [n,k], idx [n,k])`` out; CUDA tensors in -> CUDA tensors out.
        """
        torch = _torch()
        on_dev = isinstance(zp, torch.Tensor) and zp.is_cuda
        p_zp, sp, keep1 = _as_arg(zp, np.float32, self.device)
        p_zc, sc, keep2 = _as_arg(zc, np.float32, self.device)
        if sp != sc:
            raise ValueError("zp and zc must live in the same memory space")
        n, dim = keep1.shape
        c = keep2.shape[0]
        if on_dev:
            scores = torch.empty(n, k, dtype=torch.float32, device=self.torch_device)
            idx = torch.empty(n, k, dtype=torch.int32, device=self.torch_device)
            ps, pi = scores.data_ptr(), idx.data_ptr()
        else:
            scores = np.empty((n, k), dtype=np.float32)
            idx = np.empty((n, k), dtype=np.int32)
            ps, pi = scores.ctypes.data, idx.ctypes.data
        self._enter()
        _cabi.check(self._lib.owc_sim_topk(self._ctx, p_zp, n, p_zc, c, dim, k, self._prec(precision), sp, ps, pi))
        if on_dev:
            self._leave()
        return scores, idx

    @staticmethod
    def _check_targets(targets, c: int) -> None:
        """Host-side range check of numpy / host targets (device tensors are checked by the kernel, which clamps,
        flags and fails the call at its next synchronisation)."""
        if targets is None:
            return
        torch = _torch()
        if isinstance(targets, torch.Tensor):
            if targets.is_cuda or targets.numel() == 0:
                return
            lo, hi = int(targets.min()), int(targets.max())
        else:
            t = np.asarray(targets)
            if t.size == 0:
                return
            lo, hi = int(t.min()), int(t.max())
        if lo < 0 or hi >= c:
            raise IndexError(f"targets must lie in [0, {c}); got [{lo}, {hi}]")

    def match(self, zp_device, zc_device, targets=None, k: int = 1, precision: str | None = None, out: str = "host") -> dict:
        """Top-k of ``zp @ zc.T`` + paired target score on DEVICE-resident embeddings (``owc_match``): the second half of
        :meth:`score`, so a multi-GPU caller can encode its shard while the class embeddings are still in flight."""
        torch = _torch()
        for t in (zp_device, zc_device):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("match needs contiguous fp32 CUDA tensors")
        n, dim = zp_device.shape
        c = zc_device.shape[0]
        self._check_targets(targets, c)
        p_t, sp_t, keep = None, _cabi.MEM_HOST, None
        if targets is not None:
            p_t, sp_t, keep = _as_arg(targets, np.int32, self.device)
        res: dict = {}
        if out == "host":
            res["scores"] = np.empty((n, k), np.float32)
            res["idx"] = np.empty((n, k), np.int32)
            res["paired"] = np.empty(n, np.float32) if targets is not None else None
            ptr = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
            space = _cabi.MEM_HOST
        else:
            dev = self.torch_device
            res["scores"] = torch.empty(n, k, dtype=torch.float32, device=dev)
            res["idx"] = torch.empty(n, k, dtype=torch.int32, device=dev)
            res["paired"] = torch.empty(n, dtype=torch.float32, device=dev) if targets is not None else None
            ptr = lambda a: a.data_ptr() if a is not None else None  # noqa: E731
            space = _cabi.MEM_DEVICE
        if n == 0:
            return res
        self._enter()
        _cabi.check(self._lib.owc_match(self._ctx, zp_device.data_ptr(), n, zc_device.data_ptr(), c, dim, p_t, sp_t, k,
                                        self._prec(precision), ptr(res["scores"]), ptr(res["idx"]), ptr(res["paired"]), space))
        if out != "host":
            self._leave()
        del keep
        return res

    def score(self, ids, cu_seqlens, zc_device, targets=None, k: int = 1, precision: str | None = None,
              out: str = "host", want_embeds: bool = False) -> dict:
        """Whole path for one shard: encode + match + paired score. ``owc_score``."""
        torch = _torch()
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = int(cu.size) - 1
        h = self.arch.hidden
        if not (isinstance(zc_device, torch.Tensor) and zc_device.is_cuda and zc_device.dtype == torch.float32
                and zc_device.is_contiguous()):
            raise ValueError("zc_device must be a contiguous fp32 CUDA tensor [c, H]")
        c = zc_device.shape[0]
        self._check_targets(targets, c)
        p_ids, sp_ids, keep1 = _as_arg(ids, np.int32, self.device)
        p_t, keep2 = None, None
        if targets is not None:
            p_t, sp_t, keep2 = _as_arg(targets, np.int32, self.device)
            if sp_t != sp_ids:
                raise ValueError("targets must live in the same memory space as ids")
        res: dict = {}
        if out == "host":
            res["scores"] = np.empty((n, k), np.float32)
            res["idx"] = np.empty((n, k), np.int32)
            res["paired"] = np.empty(n, np.float32) if targets is not None else None
            res["embeds"] = np.empty((n, h), np.float32) if want_embeds else None
            ptr = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
            space = _cabi.MEM_HOST
        else:
            dev = self.torch_device
            res["scores"] = torch.empty(n, k, dtype=torch.float32, device=dev)
            res["idx"] = torch.empty(n, k, dtype=torch.int32, device=dev)
            res["paired"] = torch.empty(n, dtype=torch.float32, device=dev) if targets is not None else None
            res["embeds"] = torch.empty(n, h, dtype=torch.float32, device=dev) if want_embeds else None
            ptr = lambda a: a.data_ptr() if a is not None else None  # noqa: E731
            space = _cabi.MEM_DEVICE
        if n == 0:
            return res
        self._enter()
        _cabi.check(self._lib.owc_score(self._enc, p_ids, sp_ids, cu.ctypes.data, n, zc_device.data_ptr(), c, p_t, k,
                                        self._prec(precision), ptr(res["scores"]), ptr(res["idx"]), ptr(res["paired"]),
                                        ptr(res["embeds"]), space))
        if out != "host":
            self._leave()
        del keep1, keep2
        return res

    # ---- kernel-level hooks (tests, bench roofline leg) ----
    def debug_gemm(self, a, w, bias=None, precision: str = "bf16", tile_n: int = 0):
        torch = _torch()
        m, kdim = a.shape
        n = w.shape[0]
        out = torch.empty(m, n, dtype=torch.float32, device=self.torch_device)
        self._enter()
        _cabi.check(self._lib.owc_debug_gemm(self._ctx, a.data_ptr(), w.data_ptr(),
                                             bias.data_ptr() if bias is not None else None, m, n, kdim,
                                             PRECISIONS[precision], tile_n, out.data_ptr()))
        self._leave()
        return out

    def wide_counters(self, reset: bool = True) -> np.ndarray:
        """Cycle counters of the instrumented wide GEMM (``OWC_WIDE_DBG=1``): ``[6, 8]`` int64 (rows 0-3: the four
        epilogue kinds of ``k_gemm_wide``; rows 4-5: ``k_qkv_attn``), see owc_b200_debug.h."""
        out = np.zeros(48, dtype=np.int64)
        _cabi.check(self._lib.owc_debug_wide_counters(self._ctx, out.ctypes.data, int(reset)))
        return out.reshape(6, 8)

    def bench_sim_kernel(self, n: int, c: int, dim: int, k: int, iters: int) -> float:
        ms = C.c_float(0.0)
        _cabi.check(self._lib.owc_bench_sim_kernel(self._ctx, n, c, dim, k, iters, C.byref(ms)))
        return float(ms.value)
