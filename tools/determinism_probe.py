#!/usr/bin/env python
"""Runs the same forward many times and reports where the logits differ run to run or from a small-batch forward.
    python tools/determinism_probe.py [runs=60] [batch=1024] [precision=bf16|fp16] [img=224]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V
n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
prec = V.PREC_FP16 if len(sys.argv) > 3 and sys.argv[3] == "fp16" else V.PREC_BF16
S = int(sys.argv[4]) if len(sys.argv) > 4 else 224
w = V.synth_weights(S, 42)
base = V.synth_images(64, S, 7)
big = np.ascontiguousarray(np.tile(base, ((B + 63) // 64, 1, 1, 1))[:B])
with V.Engine(w, S, max_batch=B, precision=prec) as eng:
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
print(f"batch {B} precision {prec} img {S}: {n_bad} bad runs of {n_runs}")
