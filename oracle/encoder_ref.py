"""ORACLE (test infrastructure, not product code): CPU restatement of the scoring path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker or the reported CPU baseline.
Nothing under ``lmms-owc_b200/`` imports it: the product path has no CPU fallback.

What it restates (reference = /root/reference = altndrr/lmms-owc, file:line cited per function):

* the operator ``encode_sentence_bert`` - ``src/data/pipelines/text/_text.py:143-208``;
* the third-party arithmetic that operator calls and that is NOT in the reference tree:
  ``transformers`` (pinned 4.47.0 in the reference's ``uv.lock``; 5.5.0 in this image)
  ``BertModel.forward`` (``models/bert/modeling_bert.py``: embeddings :53-112, attention :114-207,
  self-output :287-298, intermediate :330-342, output :345-356) and ``MPNetModel.forward``
  (``models/mpnet/modeling_mpnet.py``: embeddings :57-96, attention :116-186, encoder and
  relative-position bias :276-362, position ids :889-897), restated here from their published
  algorithm in plain torch CPU ops (matmul / softmax / erf-GELU / LayerNorm);
* the paired score ``torch.bmm`` of ``src/data/metrics/_group.py:537-544``;
* the north-star extension: the prediction-by-class cosine matrix with row-wise top-k, defined in
  SURVEY.md section 8(c) as ``S = Zp @ Zc.T``, ``topk(S, k)`` on the operator's embeddings.

Pinning: the reference ships no test or golden vector for this path, so the oracle
is pinned against outputs of the reference itself run in the build container:
``oracle/make_golden.py`` imports ``/root/reference``, runs its ``encode_sentence_bert`` /
``semantic_similarity`` / ``mean_average_semantic_similarity`` on seeded inputs with seeded
random-init weights and commits the results under ``tests/golden/``;
``tests/test_oracle.py`` checks this file against them (and against ``transformers``'
``BertModel`` / ``MPNetModel`` directly, which are installed on the GPU box too).
"""

from __future__ import annotations

import math

import torch

__all__ = [
    "bert_forward",
    "mpnet_forward",
    "encoder_forward",
    "mean_pool",
    "l2_normalize",
    "encode_padded",
    "encode_strings",
    "paired_dot",
    "sim_topk",
    "mpnet_bucket",
]


def _layer_norm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    # nn.LayerNorm over the last dim, biased variance.
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def _gelu_erf(x: torch.Tensor) -> torch.Tensor:
    # hidden_act="gelu" is the exact erf form (modeling_bert.py:330-342 via ACT2FN["gelu"]).
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _linear(x: torch.Tensor, sd: dict, name: str, dt, operand_round=None) -> torch.Tensor:
    w = sd[name + ".weight"].to(dt)
    if operand_round is not None:  # numerics emulation of a reduced-precision MMA operand
        x = operand_round(x)
        w = operand_round(w)
    return x @ w.T + sd[name + ".bias"].to(dt)


def _attention(q, k, v, heads: int, add_mask: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    b, t, h = q.shape
    d = h // heads
    q = q.view(b, t, heads, d).transpose(1, 2)
    k = k.view(b, t, heads, d).transpose(1, 2)
    v = v.view(b, t, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if bias is not None:
        s = s + bias
    s = s + add_mask
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(b, t, h)


def _extended_mask(attention_mask: torch.Tensor, dt) -> torch.Tensor:
    # get_extended_attention_mask: (1 - mask) * finfo.min, broadcast over heads and queries.
    m = attention_mask.to(dt)[:, None, None, :]
    return (1.0 - m) * torch.finfo(dt).min


def bert_forward(sd: dict, arch, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                 dtype=torch.float32, operand_round=None, hidden_states: list | None = None) -> torch.Tensor:
    """``BertModel(**inputs).last_hidden_state`` (modeling_bert.py:53-112, :114-207, :287-356)."""
    dt = dtype
    b, t = input_ids.shape
    x = (sd["embeddings.word_embeddings.weight"].to(dt)[input_ids]
         + sd["embeddings.token_type_embeddings.weight"].to(dt)[0]
         + sd["embeddings.position_embeddings.weight"].to(dt)[:t][None])
    x = _layer_norm(x, sd["embeddings.LayerNorm.weight"].to(dt), sd["embeddings.LayerNorm.bias"].to(dt), arch.ln_eps)
    add_mask = _extended_mask(attention_mask, dt)
    if hidden_states is not None:
        hidden_states.append(x)
    for l in range(arch.layers):
        p = f"encoder.layer.{l}."
        q = _linear(x, sd, p + "attention.self.query", dt, operand_round)
        k = _linear(x, sd, p + "attention.self.key", dt, operand_round)
        v = _linear(x, sd, p + "attention.self.value", dt, operand_round)
        a = _attention(q, k, v, arch.heads, add_mask, None)
        a = _linear(a, sd, p + "attention.output.dense", dt, operand_round)
        x = _layer_norm(a + x, sd[p + "attention.output.LayerNorm.weight"].to(dt),
                        sd[p + "attention.output.LayerNorm.bias"].to(dt), arch.ln_eps)
        f = _gelu_erf(_linear(x, sd, p + "intermediate.dense", dt, operand_round))
        f = _linear(f, sd, p + "output.dense", dt, operand_round)
        x = _layer_norm(f + x, sd[p + "output.LayerNorm.weight"].to(dt),
                        sd[p + "output.LayerNorm.bias"].to(dt), arch.ln_eps)
        if hidden_states is not None:
            hidden_states.append(x)
    return x


def mpnet_bucket(relative_position: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """``MPNetEncoder.relative_position_bucket`` (modeling_mpnet.py:343-362)."""
    ret = 0
    n = -relative_position
    num_buckets //= 2
    ret = ret + (n < 0).to(torch.long) * num_buckets
    n = torch.abs(n)
    max_exact = num_buckets // 2
    is_small = n < max_exact
    val_if_large = max_exact + (
        torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    val_if_large = torch.min(val_if_large, torch.full_like(val_if_large, num_buckets - 1))
    return ret + torch.where(is_small, n, val_if_large)


def mpnet_forward(sd: dict, arch, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                  dtype=torch.float32, operand_round=None, hidden_states: list | None = None) -> torch.Tensor:
    """``MPNetModel(**inputs).last_hidden_state`` (modeling_mpnet.py:57-96, :116-186, :276-362, :889-897)."""
    dt = dtype
    b, t = input_ids.shape
    not_pad = input_ids.ne(arch.pad_id).int()
    position_ids = (torch.cumsum(not_pad, dim=1).type_as(not_pad) * not_pad).long() + arch.pad_id
    x = (sd["embeddings.word_embeddings.weight"].to(dt)[input_ids]
         + sd["embeddings.position_embeddings.weight"].to(dt)[position_ids])
    x = _layer_norm(x, sd["embeddings.LayerNorm.weight"].to(dt), sd["embeddings.LayerNorm.bias"].to(dt), arch.ln_eps)
    add_mask = _extended_mask(attention_mask, dt)
    ctx = torch.arange(t)[:, None]
    mem = torch.arange(t)[None, :]
    bucket = mpnet_bucket(mem - ctx, arch.rel_buckets, arch.rel_max_distance)
    bias = sd["encoder.relative_attention_bias.weight"].to(dt)[bucket].permute(2, 0, 1)[None]
    if hidden_states is not None:
        hidden_states.append(x)
    for l in range(arch.layers):
        p = f"encoder.layer.{l}."
        q = _linear(x, sd, p + "attention.attn.q", dt, operand_round)
        k = _linear(x, sd, p + "attention.attn.k", dt, operand_round)
        v = _linear(x, sd, p + "attention.attn.v", dt, operand_round)
        a = _attention(q, k, v, arch.heads, add_mask, bias)
        a = _linear(a, sd, p + "attention.attn.o", dt, operand_round)
        x = _layer_norm(a + x, sd[p + "attention.LayerNorm.weight"].to(dt),
                        sd[p + "attention.LayerNorm.bias"].to(dt), arch.ln_eps)
        f = _gelu_erf(_linear(x, sd, p + "intermediate.dense", dt, operand_round))
        f = _linear(f, sd, p + "output.dense", dt, operand_round)
        x = _layer_norm(f + x, sd[p + "output.LayerNorm.weight"].to(dt),
                        sd[p + "output.LayerNorm.bias"].to(dt), arch.ln_eps)
        if hidden_states is not None:
            hidden_states.append(x)
    return x


def encoder_forward(sd: dict, arch, input_ids, attention_mask, **kw) -> torch.Tensor:
    return (bert_forward if arch.kind == 0 else mpnet_forward)(sd, arch, input_ids, attention_mask, **kw)


def mean_pool(hidden: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """``_mean_pooling`` - _text.py:175-189 (mask cast to float32, clamp(min=1e-9))."""
    m = attention_mask.unsqueeze(-1).expand(hidden.size()).float().to(hidden.dtype)
    return torch.sum(hidden * m, 1) / torch.clamp(m.sum(1), min=1e-9)


def l2_normalize(z: torch.Tensor) -> torch.Tensor:
    """_text.py:202 - no epsilon (a zero vector yields NaN, as in the reference)."""
    return z / z.norm(p=2, dim=-1, keepdim=True)


@torch.no_grad()
def encode_padded(sd: dict, arch, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                  dtype=torch.float32, operand_round=None) -> torch.Tensor:
    """Forward + mean-pool + L2 on an already tokenised, padded batch (_text.py:196-202)."""
    h = encoder_forward(sd, arch, input_ids, attention_mask, dtype=dtype, operand_round=operand_round)
    return l2_normalize(mean_pool(h, attention_mask))


@torch.no_grad()
def encode_strings(sd: dict, arch, tokenizer, texts: list[str], batch_size: int = 1024,
                   dtype=torch.float32, operand_round=None) -> torch.Tensor:
    """``encode_sentence_bert`` over a whole column the way ``datasets.map(batched, 1024)`` drives it.

    Tokenises each 1024-row batch with ``padding=True, truncation=True`` (_text.py:193-195,
    batch size from _group.py:527) and returns the stacked ``[N, H]`` embeddings.
    """
    out = []
    for s in range(0, len(texts), batch_size):
        enc = tokenizer(texts[s:s + batch_size], padding=True, truncation=True, return_tensors="pt")
        out.append(encode_padded(sd, arch, enc["input_ids"], enc["attention_mask"], dtype, operand_round))
    return torch.cat(out) if out else torch.zeros(0, arch.hidden, dtype=dtype)


def paired_dot(refs_z: torch.Tensor, preds_z: torch.Tensor) -> torch.Tensor:
    """``torch.bmm(refs_z.unsqueeze(1), preds_z.unsqueeze(2)).squeeze()`` - _group.py:537-544."""
    return torch.bmm(refs_z.unsqueeze(1), preds_z.unsqueeze(2)).reshape(-1)


def sim_topk(zp: torch.Tensor, zc: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """North-star extension: ``S = Zp @ Zc.T``; row-wise top-k and top1-top2 margin.

    Ties are broken towards the lower class index (stable descending sort), which is the rule the
    CUDA epilogue implements. Returns ``(scores [N,k], idx [N,k] int64, margin [N])``.
    """
    s = zp @ zc.T
    order = torch.sort(s, dim=1, descending=True, stable=True)
    kk = min(k, s.shape[1])
    scores, idx = order.values[:, :kk], order.indices[:, :kk]
    if s.shape[1] > 1:
        margin = order.values[:, 0] - order.values[:, 1]
    else:
        margin = torch.full((s.shape[0],), float("inf"), dtype=s.dtype)
    return scores, idx, margin
