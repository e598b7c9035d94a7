#!/usr/bin/env python
"""Per-step device time of N consecutive forwards (B = 1024) right after engine start: shows how the clock /
power state settles."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V
N = int(sys.argv[1]) if len(sys.argv) > 1 else 80
B = 1024
w = V.synth_weights(224, 42)
eng = V.Engine(w, 224, max_batch=B)
imgs = V.synth_images(64, 224, 7)
imgs = np.ascontiguousarray(np.tile(imgs, (16, 1, 1, 1)))
d_imgs, d_logits = V.dev_alloc(0, imgs.nbytes), V.dev_alloc(0, B * 4000)
V.dev_upload(0, d_imgs, imgs)
ms = []
for i in range(N):
    eng.timer_start(); eng.enqueue_device(d_imgs, B, d_logits); ms.append(eng.timer_stop())
print(" ".join(f"{m:.2f}" for m in ms))
print("first 5 mean %.2f, steps 10-20 %.2f, last 20 mean %.2f" % (np.mean(ms[:5]), np.mean(ms[10:20]), np.mean(ms[-20:])))
eng.close()
