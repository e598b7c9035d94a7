#!/usr/bin/env python
"""One forward pass of B images (default 1024) through the device-resident path, for ncu:
launch order is patchify, cls_rows, gemm(embed), then per layer
ln, gemm(qkv), attention, gemm(out), ln, gemm(fc1), gemm(fc2); finally head_ln, head_gemm."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = int(sys.argv[3]) if len(sys.argv) > 3 else 224
w = V.synth_weights(S, 42)
eng = V.Engine(w, S, max_batch=B)
imgs = V.synth_images(min(B, 64), S, 7)
imgs = np.ascontiguousarray(np.tile(imgs, ((B + imgs.shape[0] - 1) // imgs.shape[0], 1, 1, 1))[:B])
d_imgs = V.dev_alloc(0, imgs.nbytes)
d_logits = V.dev_alloc(0, B * 1000 * 4)
V.dev_upload(0, d_imgs, imgs)
for _ in range(passes):
    eng.enqueue_device(d_imgs, B, d_logits)
eng.sync()
print("done", V.launch_count())
eng.close()
