#!/usr/bin/env python
"""Runs the same forward many times and reports where the logits differ run to run or from a small-batch forward.
    python tools/determinism_probe.py [runs=60] [batch=1024] [precision=auto|bf16|fp16] [img=224] [n_gpus=1]
(batch = images per call over all GPUs; the per-GPU workspace is batch / n_gpus)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V
n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
prec = {"fp16": V.PREC_FP16, "bf16": V.PREC_BF16}.get(sys.argv[3] if len(sys.argv) > 3 else "auto", V.PREC_AUTO)
S = int(sys.argv[4]) if len(sys.argv) > 4 else 224
G = int(sys.argv[5]) if len(sys.argv) > 5 else 1
w = V.synth_weights(S, 42)
base = V.synth_images(64, S, 7)
big = np.ascontiguousarray(np.tile(base, ((B + 63) // 64, 1, 1, 1))[:B])
with V.Engine(w, S, max_batch=(B + G - 1) // G, precision=prec, n_gpus=G) as eng:
    small = eng.forward(base)
    ref = np.tile(small, ((B + 63) // 64, 1))[:B]
    n_bad = 0
    for i in range(n_runs):
        o = eng.forward(big)
        d = np.abs(o - ref)
        bad = np.flatnonzero(d.max(1) > 0)
        if len(bad):
            n_bad += 1
            print(f"run {i}: {len(bad)} images differ from the 64-image forward; max |d| {d.max():.3e}; first bad images {bad[:12].tolist()}")
print(f"batch {B} precision {prec} img {S} gpus {G}: {n_bad} bad runs of {n_runs}")
