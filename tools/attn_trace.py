#!/usr/bin/env python
"""Timeline of the persistent attention kernel: runs it with the debug trace (CTA 0, first 16 items)
and prints, per item, when each pipeline event happened (SM clock cycles relative to the first event).
    python tools/attn_trace.py [batch=128] [tokens=197]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
import numpy as np
import vit_b200 as V

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 160
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 197
qkv = (np.random.default_rng(1).standard_normal((batch * tokens, 2304))).astype(np.float32)
V.attention_trace(qkv, batch, tokens)  # warm-up (module load, L2)
tr = V.attention_trace(qkv, batch, tokens).astype(np.int64)
t0 = tr[tr > 0].min()
rel = np.where(tr > 0, tr - t0, -1)
n_items = min(16, (batch * 12 + 147) // 148)
names = {w: ["start", "S_rdy", "P_done", "O_rdy", "O_ld", "end"] for w in range(16)}
print(f"batch={batch} tokens={tokens}: items traced {n_items}; cycles relative to first event")
for it in range(n_items):
    print(f"--- item {it}")
    print(f"  producer TMA issue {rel[16, it, 0]:7d} | MMA: S0 {rel[17, it, 0]:7d} S1 {rel[18, it, 0]:7d} PV0 {rel[17, it, 1]:7d} PV1 {rel[18, it, 1]:7d}")
    for w in (0, 3, 4, 8, 12, 14):
        e = rel[w, it]
        print(f"  warp {w:2d} (tile {w >> 3} half {"AB"[(w >> 2) & 1]}): " + " ".join(f"{n}={e[i]:7d}" for i, n in enumerate(names[w])) +
              f" | wait_S {e[1] - e[0]:6d} softmax {e[2] - e[1]:6d} (pass1 {e[6] - e[1]:5d} xchg {e[7] - e[6]:5d} pass2 {e[2] - e[7]:5d}) wait_O {e[3] - e[2]:6d} epi {e[5] - e[3]:6d}")
per = np.diff(rel[0, 1:n_items, 0])
print("tile-0 period per item:", per.tolist(), " tile-1:", np.diff(rel[8, 1:n_items, 0]).tolist())
