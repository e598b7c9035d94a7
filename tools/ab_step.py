#!/usr/bin/env python
"""Device-resident step time of the headline configuration (ViT-B/16 224x224, batch 1024, all rows in the last layer) for quick
A/B runs of two builds of the library: argv: [steps [repeats]] -> one line of ms per step per repeat."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, S = 1024, 224
eng = V.Engine(V.synth_weights(S, 42), S, max_batch=B)
eng.set_class_row_pruning(False)
imgs = V.synth_images(64, S, 7)
imgs = np.ascontiguousarray(np.tile(imgs, (B // 64, 1, 1, 1)))
d_imgs, d_logits = V.dev_alloc(0, imgs.nbytes), V.dev_alloc(0, B * 1000 * 4)
V.dev_upload(0, d_imgs, imgs)
for _ in range(5):
    eng.enqueue_device(d_imgs, B, d_logits)
eng.sync()
out = []
for _ in range(reps):
    eng.timer_start()
    for _ in range(steps):
        eng.enqueue_device(d_imgs, B, d_logits)
    out.append(eng.timer_stop() / steps)
print(" ".join(f"{x:.3f}" for x in out))
eng.close()
