#!/usr/bin/env python
"""GPU bring-up diagnostics: runs each kernel-level parity case in its own process (a device
trap kills the CUDA context) and prints error statistics instead of stopping at the first
failure.  Usage on the GPU box:  python tools/gpu_diag.py            (all cases)
                                 python tools/gpu_diag.py --case linear_small"""
import argparse
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np


def stats(name, got, ref, note=""):
    err = np.abs(got - ref)
    i = err.argmax()
    rel = err.max() / (np.abs(ref).max() + 1e-30)
    print(f"[{name}] max_abs_err={err.max():.4e} rel_to_max={rel:.3e} mean_err={err.mean():.3e} "
          f"ref_absmax={np.abs(ref).max():.3f} worst_idx={np.unravel_index(i, err.shape)} got={got.flat[i]:.5f} "
          f"ref={ref.flat[i]:.5f} nan={np.isnan(got).sum()} {note}", flush=True)


def rand(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def case_linear(m, n, k, epi, prec):
    import vit_b200 as V, oracle_py as O
    from conftest import round_operand
    x = round_operand(rand((m, k), 1), prec)
    W = round_operand(rand((n, k), 2, 0.03), prec)
    b = rand((n,), 3, 0.1)
    r = rand((m, n), 4) if epi == 2 else None
    t = time.time()
    got = V.op_linear(x, W, b, residual=r, epilogue=epi, precision=prec)
    dt = time.time() - t
    ref = O.linear(x, W, b)
    if epi == 1:
        ref = O.gelu(ref)
    if epi == 2:
        ref = ref + r
    stats(f"linear m={m} n={n} k={k} epi={epi} prec={prec}", got, ref, f"t={dt:.2f}s")
    # structured diagnosis: per 128-row / 256-col tile error
    err = np.abs(got - ref)
    if err.max() > 0.05:
        tm, tn = (m + 127) // 128, n // 256
        tile = np.zeros((tm, tn))
        for i in range(tm):
            for j in range(tn):
                tile[i, j] = err[i * 128:(i + 1) * 128, j * 256:(j + 1) * 256].max()
        print("  per-tile max err:\n", np.array2string(tile, precision=3, max_line_width=200))
        print("  got[0,:8]", got[0, :8], "\n  ref[0,:8]", ref[0, :8])
        print("  got[1,:8]", got[1, :8], "\n  ref[1,:8]", ref[1, :8])
        colerr = err[:128, :256].max(0)
        rowerr = err[:128, :256].max(1)
        print("  tile0 col err (first 64):", np.array2string(colerr[:64], precision=2, max_line_width=250))
        print("  tile0 row err (first 64):", np.array2string(rowerr[:64], precision=2, max_line_width=250))


def case_layernorm(rows, prec):
    import vit_b200 as V, oracle_py as O
    x = rand((rows, 768), 11, 2.0) + 0.5
    w = 1.0 + rand((768,), 12, 0.1)
    b = rand((768,), 13, 0.1)
    stats(f"layernorm rows={rows} prec={prec}", V.op_layernorm(x, w, b, precision=prec), O.layer_norm(x, w, b))


def round_qkv(qkv, prec):
    from conftest import round_operand
    out = round_operand(qkv, prec)
    out[:, 1536:] = round_operand(qkv[:, 1536:], 0)
    return out


def case_attention(batch, tokens, prec, peaky=False):
    import vit_b200 as V, oracle_py as O
    from conftest import round_operand
    qkv = round_qkv(rand((batch * tokens, 2304), 14 + tokens), prec)
    if peaky:  # one late key dominates: exercises the exact power-of-two repair of the single-pass softmax
        qkv[:, :768] *= 4.0
        qkv[150::tokens, 768:1536] = 6.0 * qkv[3::tokens, :768]
        qkv = round_qkv(qkv, prec)
    got = V.op_attention(qkv, batch, tokens, precision=prec)
    ref = np.empty((batch * tokens, 768), dtype=np.float32)
    for i in range(batch):
        blk = qkv[i * tokens:(i + 1) * tokens]
        ref[i * tokens:(i + 1) * tokens] = O.attention_core(np.ascontiguousarray(blk[:, :768]),
                                                             np.ascontiguousarray(blk[:, 768:1536]),
                                                             np.ascontiguousarray(blk[:, 1536:]))
    stats(f"attention batch={batch} tokens={tokens} prec={prec}", got, ref)
    err = np.abs(got - ref)
    if err.max() > 0.05:
        print("  per-head max err:", np.array2string(err.reshape(batch * tokens, 12, 64).max((0, 2)), precision=3))
        print("  per-row-block(32) max err:", np.array2string(
            np.array([err[i:i + 32].max() for i in range(0, batch * tokens, 32)]), precision=3, max_line_width=200))
        print("  got[0,:8]", got[0, :8], "\n  ref[0,:8]", ref[0, :8])


def case_embed(prec):
    import vit_b200 as V, oracle_py as O
    from conftest import round_operand
    w = V.synth_weights(224, 42)
    imgs = round_operand(V.synth_images(3, 224, 21), prec)
    conv_w = round_operand(w[1], prec)
    got = V.op_embed(imgs, w[0], conv_w, w[2], w[3], precision=prec)
    ref = np.concatenate([O.embed(imgs[i], w[0], conv_w, w[2], w[3]) for i in range(3)])
    stats(f"embed prec={prec}", got, ref)


def case_head():
    import vit_b200 as V, oracle_py as O
    w = V.synth_weights(224, 42)
    x = rand((5 * 197, 768), 31, 1.5)
    got = V.op_head(x, w[148], w[149], w[150], w[151], 5, 197)
    ref = O.linear(O.layer_norm(np.ascontiguousarray(x[::197]), w[148], w[149]), w[150].reshape(1000, 768), w[151])
    stats("head", got, ref)


def case_model(n, prec, max_batch, img=224):
    import vit_b200 as V, oracle_py as O
    w = V.synth_weights(img, 42)
    imgs = V.synth_images(n, img, 7)
    t = time.time()
    ref = O.forward(w, imgs, img)
    t_or = time.time() - t
    with V.Engine(w, img, max_batch=max_batch, precision=prec) as eng:
        print("engine info", eng.info(), flush=True)
        t = time.time()
        got, top1 = eng.forward(imgs, want_top1=True)
        t1 = time.time() - t
        t = time.time()
        got2 = eng.forward(imgs)
        t2 = time.time() - t
    stats(f"model n={n} prec={prec}", got, ref, f"oracle {t_or:.1f}s gpu first {t1:.3f}s second {t2:.3f}s")
    viol = (np.abs(got - ref) > 2e-2 + 1e-2 * np.abs(ref)).sum()
    print(f"  top1 equal: {np.array_equal(top1, ref.argmax(1))}  tolerance violations: {viol}/{got.size}  "
          f"deterministic: {np.array_equal(got, got2)}  logit std {ref.std():.3f}", flush=True)
    srt = np.sort(ref, axis=1)
    print("  top1-top2 margins:", np.array2string(srt[:, -1] - srt[:, -2], precision=3, max_line_width=200))


CASES = {
    "linear_small": lambda: case_linear(128, 256, 64, 0, 0),
    "linear_k768": lambda: case_linear(128, 256, 768, 0, 0),
    "linear_qkv": lambda: case_linear(197, 2304, 768, 0, 0),
    "linear_qkv_fp16": lambda: case_linear(197, 2304, 768, 0, 1),
    "linear_gelu": lambda: case_linear(197, 3072, 768, 1, 0),
    "linear_resid": lambda: case_linear(394, 768, 3072, 2, 0),
    "linear_big": lambda: case_linear(4000, 768, 768, 0, 0),
    "layernorm": lambda: case_layernorm(197, 0),
    "attention_64": lambda: case_attention(1, 64, 0),
    "attention_197": lambda: case_attention(2, 197, 0),
    "attention_197_fp16": lambda: case_attention(2, 197, 1),
    "attention_224": lambda: case_attention(1, 224, 0),
    "attention_256": lambda: case_attention(1, 256, 0),
    "attention_300": lambda: case_attention(2, 300, 0),
    "attention_577": lambda: case_attention(2, 577, 0),
    "attention_577_fp16": lambda: case_attention(2, 577, 1),
    "attention_640": lambda: case_attention(1, 640, 0),
    "attention_577_many": lambda: case_attention(30, 577, 0),
    "model384_bf16": lambda: case_model(2, 0, 2, 384),
    "attention_many": lambda: case_attention(40, 197, 0),
    "attention_peaky": lambda: case_attention(2, 197, 0, True),
    "attention_peaky_fp16": lambda: case_attention(2, 197, 1, True),
    "attention_many_small": lambda: case_attention(70, 100, 0),
    "embed": lambda: case_embed(0),
    "head": case_head,
    "model_bf16": lambda: case_model(8, 0, 8),
    "model_fp16": lambda: case_model(8, 1, 8),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", nargs="*")
    a = ap.parse_args()
    if a.case:
        CASES[a.case]()
        sys.exit(0)
    for name in (a.only or CASES):
        t = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=300)
            out = (p.stdout + p.stderr).strip()
            tail = "\n".join(out.splitlines()[-30:])
            print(f"=== {name}: rc={p.returncode} ({time.time() - t:.1f}s)\n{tail}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"=== {name}: TIMEOUT", flush=True)
