#!/usr/bin/env python
"""Per-category device time of a small-batch forward (default 1 image), averaged over 50 runs."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = V.synth_weights(224, 42)
eng = V.Engine(w, 224, max_batch=max(B, 8))
imgs = V.synth_images(B, 224, 7)
d_imgs, d_logits = V.dev_alloc(0, imgs.nbytes), V.dev_alloc(0, B * 4000)
V.dev_upload(0, d_imgs, imgs)
for _ in range(10):
    eng.enqueue_device(d_imgs, B, d_logits)
eng.sync()
eng.profile_enable(True)
N = 50
for _ in range(N):
    eng.enqueue_device(d_imgs, B, d_logits)
eng.sync()
prof = eng.profile_read()
eng.profile_enable(False)
tot = 0
for k, v in prof.items():
    print(f"{k:12s} {v['ms'] / N * 1e3:8.1f} us  ({v['launches'] // N} launches, {v['ms'] / max(v['launches'], 1) * 1e3:6.1f} us each)")
    tot += v["ms"] / N
print(f"sum of kernels {tot * 1e3:.1f} us")
ms = []
for _ in range(100):
    eng.timer_start(); eng.enqueue_device(d_imgs, B, d_logits); ms.append(eng.timer_stop())
print(f"graph replay median {np.median(ms) * 1e3:.1f} us")
eng.close()
