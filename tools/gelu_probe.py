#!/usr/bin/env python
"""Accuracy of the GELU the mlp_0 epilogue computes, on the GPU: every FP16 value in [-8, 8] goes through
vit_cuda_op_linear (identity weights, zero bias, FP16 operands, EPI_BIAS_GELU) and the 16-bit results are compared with the
fp64 definition 0.5 x (1 + erf(x / sqrt 2)) (ViT_seq.c:231-233) and with that definition rounded to FP16 (the best any
epilogue can store)."""
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V

bits = np.arange(0, 1 << 16, dtype=np.uint16)
vals = bits.view(np.float16).astype(np.float32)
vals = vals[np.isfinite(vals) & (np.abs(vals) <= 8.0)]
K = 768
rows = (vals.size + K - 1) // K
x = np.zeros(rows * K, np.float32)
x[:vals.size] = vals
x = np.ascontiguousarray(x.reshape(rows, K))
W = np.ascontiguousarray(np.eye(K, dtype=np.float32))
b = np.zeros(K, np.float32)
got = V.op_linear(x, W, b, epilogue=V.EPI_BIAS_GELU, precision=V.PREC_FP16).reshape(-1)[:vals.size].astype(np.float64)
xv = vals.astype(np.float64)
exact = 0.5 * xv * (1.0 + np.vectorize(math.erf)(xv / math.sqrt(2.0)))
best = exact.astype(np.float16).astype(np.float64)
err, floor = np.abs(got - exact), np.abs(best - exact)
i = int(err.argmax())
print(f"values {vals.size}  max |gpu - exact| {err.max():.3e} at x = {xv[i]:.4f} (rounding alone there: {floor[i]:.3e})  "
      f"mean {err.mean():.3e}  | rounding alone: max {floor.max():.3e} mean {floor.mean():.3e}")
for lo, hi in [(-8, -4), (-4, -2), (-2, -1), (-1, 0), (0, 1), (1, 2), (2, 4), (4, 8)]:
    m = (xv >= lo) & (xv < hi)
    print(f"  x in [{lo:3d}, {hi:3d}): max err {err[m].max():.3e}  mean {err[m].mean():.3e}   rounding alone: max {floor[m].max():.3e} mean {floor[m].mean():.3e}")
print(f"results that differ from the correctly rounded FP16 value: {int((got != best).sum())} of {vals.size}")
