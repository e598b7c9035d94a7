#!/usr/bin/env python
"""CPU emulation of operand-precision policies for the ViT-B/16 forward (torch, fp32 arithmetic with the GEMM operands
rounded to the candidate type; products and accumulation exact fp32 -- SURVEY.md App. E's method) on the seed-42
synthetic weights and seed-7 images the tests and the bench use.  Answers, before any kernel is written: which
policies meet the stated 2e-2 + 1e-2 |ref| logit tolerance, and what an FP8 (kind::f8f6f4, E4M3) MLP would cost.
    python tools/precision_emulation.py [n_images] > profiles/r2_precision_emulation.txt"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import vit_hostio as H

torch.set_num_threads(8)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
W = [torch.from_numpy(a) for a in H.synth_weights(224, 42)]
IMGS = torch.from_numpy(H.synth_images(N, 224, 7))


def rnd(x, kind, per_row=False):
    if kind == "fp32":
        return x
    if kind == "bf16":
        return x.to(torch.bfloat16).float()
    if kind == "fp16":
        return x.to(torch.float16).float()
    if kind == "tf32t":   # truncation to a 10-bit mantissa (what kind::tf32 keeps of an fp32 operand)
        return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)
    if kind == "e4m3":    # scaled to the type's range (448): per tensor, or per row for activations
        amax = x.abs().amax(dim=-1, keepdim=True) if per_row else x.abs().max()
        s = 448.0 / amax.clamp_min(1e-30)
        return (x * s).to(torch.float8_e4m3fn).float() / s
    raise ValueError(kind)


def forward(policy):
    """policy: dict op -> (activation kind, weight kind[, per_row]) for op in conv, qkv, attn, out, fc1, fc2."""
    def lin(op, x, w, b):
        a_kind, w_kind, *rest = policy[op]
        return rnd(x, a_kind, bool(rest and rest[0])) @ rnd(w, w_kind).t() + b

    B = IMGS.shape[0]
    p = IMGS.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, 196, 768)
    x = lin("conv", p, W[1].reshape(768, 768), W[2])
    x = torch.cat([W[0].reshape(1, 1, 768).expand(B, 1, 768), x], 1) + W[3].reshape(1, 197, 768)
    ln = torch.nn.functional.layer_norm
    for l in range(12):
        w = W[4 + 12 * l: 16 + 12 * l]
        h = ln(x, (768,), w[0], w[1], 1e-6)
        qkv = lin("qkv", h, w[2].reshape(2304, 768), w[3]).reshape(B, 197, 3, 12, 64).permute(2, 0, 3, 1, 4)
        ak = policy["attn"]
        q, k, v = rnd(qkv[0], ak[0]), rnd(qkv[1], ak[0]), rnd(qkv[2], ak[1])
        s = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
        o = (rnd(s, ak[1]) @ v).permute(0, 2, 1, 3).reshape(B, 197, 768)
        x = x + lin("out", o, w[4].reshape(768, 768), w[5])
        h = ln(x, (768,), w[6], w[7], 1e-6)
        h = torch.nn.functional.gelu(lin("fc1", h, w[8].reshape(3072, 768), w[9]))
        x = x + lin("fc2", h, w[10].reshape(768, 3072), w[11])
    c = ln(x[:, 0], (768,), W[148], W[149], 1e-6)
    return c @ W[150].reshape(1000, 768).t() + W[151]


def uniform(kind, attn=None, conv=None):
    d = {op: (kind, kind) for op in ("conv", "qkv", "out", "fc1", "fc2")}
    d["attn"] = attn or (kind, "bf16" if kind != "fp32" else "fp32")
    if conv:
        d["conv"] = conv
    return d


with torch.no_grad():
    ref = forward(uniform("fp32"))
    rows = [("fp32 (reference arithmetic)", uniform("fp32")),
            ("BF16 operands everywhere (VIT_PREC_BF16)", uniform("bf16", conv=("tf32t", "fp32"))),
            ("FP16 operands, P and V in BF16, conv_proj tf32 (VIT_PREC_FP16 = the default policy's fast path)", uniform("fp16", conv=("tf32t", "fp32"))),
            ("BF16 activations x FP16 weights", {**{op: ("bf16", "fp16") for op in ("qkv", "out", "fc1", "fc2")}, "conv": ("tf32t", "fp32"), "attn": ("bf16", "bf16")}),
            ("FP16, but mlp_3 (hidden x W2) in E4M3, per-tensor scales", {**uniform("fp16", conv=("tf32t", "fp32")), "fc2": ("e4m3", "e4m3")}),
            ("FP16, but mlp_3 in E4M3, per-row activation scales", {**uniform("fp16", conv=("tf32t", "fp32")), "fc2": ("e4m3", "e4m3", True)}),
            ("FP16, but mlp_0 and mlp_3 in E4M3, per-row activation scales", {**uniform("fp16", conv=("tf32t", "fp32")), "fc1": ("e4m3", "e4m3", True), "fc2": ("e4m3", "e4m3", True)})]
    print(f"# operand-precision policies emulated on the CPU: {N} seed-7 images, seed-42 weights, logit std {ref.std():.3f}; fp32 torch reference")
    print(f"# stated tolerance: |dlogit| <= 2e-2 + 1e-2 |ref| on every logit, top-1 identical")
    print(f"{'policy':100s} {'max|dlogit|':>12s} {'mean':>9s} {'outside tol':>12s} {'top-1 same':>10s}")
    for name, pol in rows:
        got = forward(pol)
        err = (got - ref).abs()
        bad = int((err > 2e-2 + 1e-2 * ref.abs()).sum())
        print(f"{name:100s} {err.max():12.4f} {err.mean():9.5f} {bad:7d}/{err.numel():<5d} {int((got.argmax(1) == ref.argmax(1)).sum()):>5d}/{N}")
