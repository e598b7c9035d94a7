#!/usr/bin/env python
"""Runs the same 1024-image forward several times and reports where the logits differ run to run."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V
n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w = V.synth_weights(224, 42)
base = V.synth_images(64, 224, 7)
big = np.ascontiguousarray(np.tile(base, (16, 1, 1, 1)))
with V.Engine(w, 224, max_batch=1024) as eng:
    small = eng.forward(base)
    outs = [eng.forward(big) for _ in range(n_runs)]
ref = np.broadcast_to(small, (16, 64, 1000)).reshape(1024, 1000)
for i, o in enumerate(outs):
    d = np.abs(o - ref)
    bad = np.flatnonzero(d.max(1) > 0)
    print(f"run {i}: {len(bad)} images differ from the 64-image forward; max |d| {d.max():.3e}; first bad images {bad[:12].tolist()}")
