"""Bring-up probe for the im2col-free conv_proj (run on the GPU box): with identity weights, out[patch][n] is the pixel
that the kernel placed at K slot n of that patch's operand row.  Five images whose pixel value is one coordinate
(kx, ky, gx, gy, channel -- all exactly representable in tf32) decode the A operand's layout completely."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
sys.path.insert(0, str(ROOT / "oracle"))


def probe(S=224):
    import numpy as np
    import vit_b200 as V
    import oracle_py as O
    G = S // 16
    T = G * G + 1
    eye = np.eye(768, dtype=np.float32).reshape(-1).copy()
    zero768 = np.zeros(768, dtype=np.float32)
    pos0 = np.zeros(T * 768, dtype=np.float32)
    c_i, y_i, x_i = np.meshgrid(np.arange(3), np.arange(S), np.arange(S), indexing="ij")
    coords = {"kx": x_i % 16, "ky": y_i % 16, "gx": x_i // 16, "gy": y_i // 16, "c": c_i}
    n = np.arange(768)
    want_k = {"kx": n % 16, "ky": (n // 16) % 16, "c": n // 256}
    for name, val in coords.items():
        img = np.ascontiguousarray(val[None].astype(np.float32))
        out = V.op_embed(img, zero768, eye, zero768, pos0, precision=1)[1:]          # [patch][n]
        p = np.arange(G * G)
        if name in want_k:
            want = np.broadcast_to(want_k[name][None, :], out.shape)
        elif name == "gx":
            want = np.broadcast_to((p % G)[:, None], out.shape)
        else:
            want = np.broadcast_to((p // G)[:, None], out.shape)
        ok = out == want
        print(f"  {name}: {ok.mean() * 100:6.2f} % of (patch, k) slots right; rows fully right: {int(ok.all(1).sum())} / {G * G}")
        if not ok.all():
            bad = np.argwhere(~ok)[:6]
            print("     first mismatches (patch, k): got / want:", [(int(a), int(b), float(out[a, b]), int(want[a, b])) for a, b in bad])
            print("     patch 0, k 0..39:", out[0, :40].tolist())
            print("     patch 1, k 0..19:", out[1, :20].tolist(), "| patch", G, "k 0..7:", out[G, :8].tolist())
    rng = np.random.default_rng(0)
    w = V.synth_weights(S, 42) if S == 224 else None
    if w is not None:
        imgs = (V.synth_images(2, S, 21).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        cw = ((w[1].view(np.uint32).astype(np.uint64) + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)
        got = V.op_embed(imgs, w[0], cw, w[2], w[3], precision=1)
        ref = np.concatenate([O.embed(imgs[i], w[0], cw, w[2], w[3]) for i in range(2)])
        err = np.abs(got - ref)
        print(f"  random data: max err {err.max():.3e}; rows within 1e-4: {int((err.max(1) < 1e-4).sum())} / {len(err)}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        probe(int(sys.argv[1]))
    else:
        for S in ("224", "384", "32"):
            print(f"== {S}x{S}", flush=True)
            subprocess.run([sys.executable, __file__, S])
