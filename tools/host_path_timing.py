#!/usr/bin/env python
"""Host-to-host throughput of the three input forms of the C ABI at 1024 images: pinned contiguous, pageable contiguous,
and one allocation per image (the reference loader's form, what ViT_cuda() receives)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V
B = 1024
w = V.synth_weights(224, 42)
imgs = V.synth_images(B, 224, 7)
parts = [imgs[i].copy() for i in range(B)]
h_imgs, h_ptr = V.pinned_empty(imgs.shape)
h_imgs[...] = imgs
h_log, l_ptr = V.pinned_empty((B, 1000))
with V.Engine(w, 224, max_batch=B) as eng:
    def timeit(f, n=6):
        f(); f()
        t0 = time.perf_counter()
        for _ in range(n):
            f()
        return B * n / (time.perf_counter() - t0)
    a = timeit(lambda: eng.forward_raw(h_ptr, B, l_ptr))
    b = timeit(lambda: eng.forward(imgs))
    c = timeit(lambda: eng.forward_scattered(parts))
    print(f"pinned contiguous {a:9.0f} images/s | pageable contiguous {b:9.0f} | one allocation per image {c:9.0f}")
