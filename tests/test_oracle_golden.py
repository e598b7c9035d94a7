"""The oracle (oracle/vit_oracle.c) against golden vectors produced by the reference's own
compiled code (tests/golden/make_golden.py).  Bit-exact: the oracle restates ViT_seq.c with the
same fp32 operation order."""
from pathlib import Path

import numpy as np

G = Path(__file__).resolve().parent / "golden"


def _ops_inputs():
    rng = np.random.default_rng(123)
    T, D = 197, 768
    x = rng.standard_normal((T, D)).astype(np.float32) * 1.7 + 0.3
    lw = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    lb = (0.1 * rng.standard_normal(D)).astype(np.float32)
    W = (rng.standard_normal((96, D)) * 0.05).astype(np.float32)
    b = (rng.standard_normal(96) * 0.1).astype(np.float32)
    g_in = np.linspace(-6, 6, 4001).astype(np.float32)
    sm_in = (rng.standard_normal(1000) * 3).astype(np.float32)
    in_w = (rng.standard_normal((3 * D, D)) * 0.03).astype(np.float32)
    in_b = (rng.standard_normal(3 * D) * 0.02).astype(np.float32)
    out_w = (rng.standard_normal((D, D)) * 0.02).astype(np.float32)
    out_b = (rng.standard_normal(D) * 0.02).astype(np.float32)
    return x, lw, lb, W, b, g_in, sm_in, in_w, in_b, out_w, out_b


def test_ops_match_reference_outputs(oracle):
    gold = np.load(G / "ops.npz")
    x, lw, lb, W, b, g_in, sm_in, in_w, in_b, out_w, out_b = _ops_inputs()
    ln = oracle.layer_norm(x, lw, lb)
    assert np.array_equal(ln[:8], gold["ln"])                       # layer_norm, ViT_seq.c:103-121
    assert np.array_equal(oracle.linear(x, W, b)[:8], gold["lin"])  # linear_layer, ViT_seq.c:240-250
    assert np.array_equal(oracle.gelu(g_in), gold["gelu"])          # gelu, ViT_seq.c:231-233
    assert np.array_equal(oracle.softmax(sm_in[None])[0], gold["softmax"])  # Softmax, ViT_seq.c:304-324
    import ctypes as C
    mha = np.empty_like(ln)
    oracle.lib.oracle_multihead_attn(oracle._p(ln), oracle._p(mha), 197, oracle._p(in_w), oracle._p(in_b),
                                     oracle._p(out_w), oracle._p(out_b))
    assert np.array_equal(mha[:4], gold["mha"])                     # multihead_attn, ViT_seq.c:123-229


def test_whole_model_matches_reference_probabilities(vit, oracle, weights224):
    gold = np.load(G / "vit_seq_probs.npz")
    imgs = vit.synth_images(3, 224, int(gold["images_seed"]))
    logits, probs = oracle.forward(weights224, imgs, 224, want_probs=True)
    order = np.argsort(-probs, axis=1)[:, :8]
    assert np.array_equal(order, gold["top_idx"])
    assert np.array_equal(np.take_along_axis(probs, order, 1), gold["top_prob"])
    assert np.array_equal(probs.astype(np.float64).sum(1), gold["checksum"])
    assert np.array_equal(np.sqrt((probs.astype(np.float64) ** 2).sum(1)), gold["l2"])
    # logits are the oracle's one extension over ViT_seq (which discards them): consistent with probs
    assert np.array_equal(oracle.softmax(logits), probs)


def test_threading_does_not_change_results(vit, oracle, weights224):
    imgs = vit.synth_images(2, 224, 99)
    a = oracle.forward(weights224, imgs, 224, n_threads=1)
    b = oracle.forward(weights224, imgs, 224, n_threads=0)
    assert np.array_equal(a, b)
