"""CPU numerics emulation of the deferred-LayerNorm ("LN fold") scheme of the wide encoder path, run before the
kernels existed: the GEMM that follows a LayerNorm consumes bf16(y) of the PRE-LayerNorm sum y and applies the
normalisation in its epilogue,
    LN(y) W^T + b = rstd * (bf16(y) . bf16(W*gamma)^T  -  mu * rowsum(bf16(W*gamma))) + (W beta + b),
against (a) the fp32 oracle and (b) the straightforward scheme (bf16(LN(y)) . bf16(W)^T).
Usage: python tools/lnfold_emulate.py [n_strings] [gamma_spread] [beta_spread] [init_range]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import owc_b200  # noqa: E402,F401
from oracle import encoder_ref as R  # noqa: E402
from oracle.hf_models import build_hf_tokenizer  # noqa: E402
from owc_b200 import arch as A, synth  # noqa: E402
from owc_b200.weights import random_state_dict  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g_spread = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
b_spread = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
init = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
arch = A.ARCHS["mpnet-base"]
sd = random_state_dict(arch)
torch.manual_seed(7)
for k in list(sd):
    if init != 0.02 and k.endswith(".weight") and sd[k].dim() == 2 and "embeddings" not in k:
        sd[k] = sd[k] * (init / 0.02)
    if "LayerNorm.weight" in k and g_spread:
        sd[k] = 1.0 + g_spread * torch.randn_like(sd[k])
    if "LayerNorm.bias" in k and b_spread:
        sd[k] = b_spread * torch.randn_like(sd[k])
vocab = synth.build_vocab("mpnet", arch.vocab)
words = synth.vocab_words(vocab)
classes = synth.make_class_names(200, words)
preds, _ = synth.make_predictions(n, classes, words)
tok = build_hf_tokenizer(arch, vocab)
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)  # noqa: E731


def stats(y, eps):
    mu = y.mean(-1, keepdim=True)
    var = (y * y).mean(-1, keepdim=True) - mu * mu
    return mu, torch.rsqrt(var.clamp_min(0) + eps)


def folded_linear(y, mu, rstd, g, b, W, bias):
    Wf = bf(W * g[None, :])
    acc = bf(y) @ Wf.T
    return rstd * (acc - mu * Wf.sum(1)[None]) + (W @ b + bias)[None]


@torch.no_grad()
def forward_fold(ids, mask):
    dt = torch.float32
    not_pad = ids.ne(arch.pad_id).int()
    pos = (torch.cumsum(not_pad, 1).type_as(not_pad) * not_pad).long() + arch.pad_id
    x = sd["embeddings.word_embeddings.weight"][ids] + sd["embeddings.position_embeddings.weight"][pos]
    x = R._layer_norm(x, sd["embeddings.LayerNorm.weight"], sd["embeddings.LayerNorm.bias"], arch.ln_eps)
    t = ids.shape[1]
    add_mask = R._extended_mask(mask, dt)
    bucket = R.mpnet_bucket(torch.arange(t)[None, :] - torch.arange(t)[:, None], arch.rel_buckets, arch.rel_max_distance)
    bias = sd["encoder.relative_attention_bias.weight"][bucket].permute(2, 0, 1)[None]
    # residual stream is carried as (y, mu, rstd, gamma, beta): x = (y - mu) * rstd * gamma + beta
    H = arch.hidden
    y, mu, rstd = x, torch.zeros_like(x[..., :1]), torch.ones_like(x[..., :1])
    g, b = torch.ones(H), torch.zeros(H)
    for l in range(arch.layers):
        p = f"encoder.layer.{l}."
        lin = lambda name: (sd[p + name + ".weight"], sd[p + name + ".bias"])  # noqa: E731
        q = bf(folded_linear(y, mu, rstd, g, b, *lin("attention.attn.q")))
        k = bf(folded_linear(y, mu, rstd, g, b, *lin("attention.attn.k")))
        v = bf(folded_linear(y, mu, rstd, g, b, *lin("attention.attn.v")))
        a = bf(R._attention(q, k, v, arch.heads, add_mask, bias))
        Wo, bo = lin("attention.attn.o")
        xres = (y - mu) * rstd * g + b
        y1 = a @ bf(Wo).T + bo + xres
        mu1, rstd1 = stats(y1, arch.ln_eps)
        g1, b1 = sd[p + "attention.LayerNorm.weight"], sd[p + "attention.LayerNorm.bias"]
        f = bf(R._gelu_erf(folded_linear(y1, mu1, rstd1, g1, b1, *lin("intermediate.dense"))))
        W2, b2 = lin("output.dense")
        y = f @ bf(W2).T + b2 + ((y1 - mu1) * rstd1 * g1 + b1)
        mu, rstd = stats(y, arch.ln_eps)
        g, b = sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"]
    return (y - mu) * rstd * g + b


@torch.no_grad()
def encode(fn, texts):
    enc = tok(texts, padding=True, truncation=True, return_tensors="pt")
    h = fn(enc["input_ids"], enc["attention_mask"])
    return R.l2_normalize(R.mean_pool(h, enc["attention_mask"]))


ref = lambda i, m: R.mpnet_forward(sd, arch, i, m)  # noqa: E731
plain = lambda i, m: R.mpnet_forward(sd, arch, i, m, operand_round=bf)  # noqa: E731
zc = {name: encode(fn, classes) for name, fn in (("ref", ref), ("plain", plain), ("fold", forward_fold))}
zp = {name: encode(fn, preds) for name, fn in (("ref", ref), ("plain", plain), ("fold", forward_fold))}
S = {k: zp[k] @ zc[k].T for k in zp}
top = S["ref"].topk(2, dim=1)
margin = top.values[:, 0] - top.values[:, 1]
clear = margin >= 1e-3
for name in ("plain", "fold"):
    dz = (zp[name] - zp["ref"]).abs().max().item()
    ds = (S[name] - S["ref"]).abs().max().item()
    agree = (S[name].argmax(1) == S["ref"].argmax(1))
    print(f"{name:6s} max|dz|={dz:.3e} max|dS|={ds:.3e} argmax agree={agree.float().mean():.4f} off-near-ties={agree[clear].float().mean():.4f} (n clear {int(clear.sum())})")
