#!/usr/bin/env python
"""Timeline of the attention kernel: runs it with the debug trace (CTA 0, first 16 units) and prints when
each pipeline event happened (SM clock cycles relative to the first event).
    python tools/attn_trace.py [batch=160] [tokens=197]
Streaming kernel (default): a unit is one 128-row query tile.  (Round 1 also had a two-slot persistent kernel; it is gone.)  See
git history of this tool."""
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
print(f"batch={batch} tokens={tokens}; cycles relative to first event")
for u in range(16):
    print(f"--- unit {u}: producer(item) {rel[16, u, 0]:7d} | issuer: S {rel[17, u, 0]:7d} PV {rel[17, u, 1]:7d}")
    for w in (0, 3, 4, 8, 10):
        e = rel[w, u]
        print(f"  exp warp {w:2d} (part {w >> 2} q{w & 3}): start={e[0]:7d} S_rdy={e[1]:7d} ref={e[6]:7d} P_done={e[2]:7d} | wait_S {e[1]-e[0]:6d} ref {e[6]-e[1]:5d} exp {e[2]-e[6]:6d}")
    for w in (12, 14):
        e = rel[w, u]
        print(f"  out warp {w:2d}: O_rdy={e[3]:7d} O_ld={e[4]:7d} stored={e[5]:7d} | ld {e[4]-e[3]:5d} store {e[5]-e[4]:5d}")
print("exp warp 0 period per unit:", np.diff(rel[0, 1:16, 0]).tolist())
