#!/usr/bin/env python
"""One forward pass of B images (default 1024) through the device-resident path, for ncu:
launch order is cls_rows, gemm(embed: conv_proj), then per layer gemm(qkv), attention, gemm(out), gemm(fc1), gemm(fc2);
finally head_ln, head_gemm: 64 launches per pass.  argv: [batch [passes [img_size [pruned]]]]"""
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
eng.set_class_row_pruning(len(sys.argv) > 4 and sys.argv[4] == "pruned")   # default: all rows in the last layer, as bench.py's headline
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
