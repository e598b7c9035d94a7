#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference's OWN code: oracle/_ref/libvit_ref.so is
ViT_seq.c / Network.c compiled unmodified from /root/reference (oracle/Makefile).  Run in the
build container (where /root/reference exists):   python tests/golden/make_golden.py

Fixtures (all inputs are the seeded synthetic assets of host/synth.c, so they can be regenerated
bit-exactly on any machine and are not stored):
  vit_seq_probs.npz      ViT_seq() softmax probabilities for 3 images (seed 7) with synthetic
                         weights (seed 42): top-1 label, top-1 prob, the 8 largest probs + indices,
                         and a float64 checksum of each 1000-vector
  ops.npz                outputs of the reference's layer_norm / linear_layer / gelu / Softmax /
                         multihead_attn on small seeded inputs (full tensors, they are small)
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
import oracle_py as O
import vit_b200 as V

OUT = Path(__file__).resolve().parent
f32p = C.POINTER(C.c_float)


def P(a):
    return a.ctypes.data_as(f32p)


def net(a):
    n = O._RefNetwork()
    n.data, n.size = P(a), a.size
    return n


def main():
    assert O.ref_available(), "build oracle/_ref first (make -C oracle)"
    r = O.ref_lib()
    w = V.synth_weights(224, 42)
    imgs = V.synth_images(3, 224, 7)
    probs = O.ref_vit_seq(w, imgs)
    order = np.argsort(-probs, axis=1)[:, :8]
    np.savez(OUT / "vit_seq_probs.npz", top_idx=order.astype(np.int32), top_prob=np.take_along_axis(probs, order, 1),
             checksum=probs.astype(np.float64).sum(1), l2=np.sqrt((probs.astype(np.float64) ** 2).sum(1)),
             weights_seed=42, images_seed=7)

    rng = np.random.default_rng(123)
    T, D = 197, 768
    x = rng.standard_normal((T, D)).astype(np.float32) * 1.7 + 0.3
    lw = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    lb = (0.1 * rng.standard_normal(D)).astype(np.float32)
    ln = np.zeros_like(x)
    r.layer_norm.argtypes = [f32p, f32p, O._RefNetwork, O._RefNetwork]
    r.layer_norm(P(x), P(ln), net(lw), net(lb))

    W = (rng.standard_normal((96, D)) * 0.05).astype(np.float32)
    b = (rng.standard_normal(96) * 0.1).astype(np.float32)
    lin = np.zeros((T, 96), dtype=np.float32)
    r.linear_layer.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, O._RefNetwork, O._RefNetwork]
    r.linear_layer(P(x), P(lin), T, D, 96, net(W), net(b))

    g_in = np.linspace(-6, 6, 4001).astype(np.float32)
    g_out = np.zeros_like(g_in)
    r.gelu_activation.argtypes = [f32p, f32p, C.c_int]
    r.gelu_activation(P(g_in), P(g_out), g_in.size)

    sm_in = (rng.standard_normal(1000) * 3).astype(np.float32)
    sm_out = np.zeros_like(sm_in)
    r.Softmax.argtypes = [f32p, f32p, C.c_int]
    r.Softmax(P(sm_in), P(sm_out), 1000)

    in_w = (rng.standard_normal((3 * D, D)) * 0.03).astype(np.float32)
    in_b = (rng.standard_normal(3 * D) * 0.02).astype(np.float32)
    out_w = (rng.standard_normal((D, D)) * 0.02).astype(np.float32)
    out_b = (rng.standard_normal(D) * 0.02).astype(np.float32)
    xa = ln.copy()
    mha = np.zeros_like(xa)
    r.multihead_attn.argtypes = [f32p, f32p] + [O._RefNetwork] * 4
    r.multihead_attn(P(xa), P(mha), net(in_w), net(in_b), net(out_w), net(out_b))
    # inputs are regenerated from the seed in the test; store outputs (fp16-size-reduced tensors kept exact: float32)
    np.savez_compressed(OUT / "ops.npz", ln=ln[:8], lin=lin[:8], gelu=g_out, softmax=sm_out, mha=mha[:4], seed=123)
    print("wrote", [p.name for p in OUT.glob("*.npz")])


if __name__ == "__main__":
    main()
