"""Whole-model parity through the C ABI (vit_cuda_init / vit_cuda_forward and the
reference-signature adaptor ViT_cuda) against the oracle, which is pinned bit-exactly to the
reference's ViT_seq (tests/test_oracle_vs_reference.py).

Stated tolerance (BASELINE.json north_star): top-1 identical; logits within
2e-2 absolute + 1e-2 relative.  The engine's DEFAULT precision policy (VIT_PREC_AUTO: FP16 operands / FP32
accumulate, BF16 operand set as the overflow fallback) is held to exactly that, every logit, strict top-1.
Explicit BF16 operands are a non-default variant with their own, wider, stated bound (BF16_ATOL)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATOL, RTOL = 2e-2, 1e-2
BF16_ATOL = 3.5e-2     # non-default VIT_PREC_BF16: 8-bit-mantissa operands through 12 layers, logit std ~1 (0.027-0.030)
N_IMAGES = 16


@pytest.fixture(scope="module")
def ref16(vit, oracle, weights224):
    imgs = vit.synth_images(N_IMAGES, 224, 7)
    logits = oracle.forward(weights224, imgs, 224)
    return imgs, logits


def _assert_top1(top1, ref, what, atol=ATOL):
    """Top-1 against the oracle for the NON-default BF16 variant: identical wherever the oracle's margin exceeds twice
    that variant's tolerance, and never a class outside the tolerance of the oracle's maximum (random-init logits,
    std ~1, have a few near-ties closer than 8-bit-mantissa operand noise).  The default policy is checked with plain
    equality (_assert_strict)."""
    best = ref.max(1)
    slack = 2 * (atol + RTOL * np.abs(best))
    picked = ref[np.arange(len(top1)), top1]
    decisive = np.sort(ref, 1)[:, -1] - np.sort(ref, 1)[:, -2] > slack
    assert np.array_equal(top1[decisive], ref.argmax(1)[decisive]), f"{what}: top-1 differs on a decisive image"
    assert np.all(picked >= best - slack), f"{what}: picked class outside the tolerance of the oracle's maximum"
    return int(decisive.sum())


def _assert_strict(got, top1, ref, what):
    """The stated tolerance, as stated: every logit within 2e-2 + 1e-2 |ref|, top-1 identical on every image."""
    err = np.abs(got - ref)
    assert np.all(err <= ATOL + RTOL * np.abs(ref)), f"{what}: {_report(got, ref)}"
    assert np.array_equal(top1, ref.argmax(1)), f"{what}: top-1 {top1.tolist()} vs oracle {ref.argmax(1).tolist()}"


def _assert_bf16_variant(got, top1, ref, what):
    err = np.abs(got - ref)
    _assert_top1(top1, ref, what, BF16_ATOL)
    assert np.all(err <= BF16_ATOL + RTOL * np.abs(ref)), f"{what}: {_report(got, ref)}"
    assert (err <= ATOL + RTOL * np.abs(ref)).mean() >= 0.99, f"{what}: {_report(got, ref)}"


def _report(got, ref):
    err = np.abs(got - ref)
    return (f"max|dlogit| {err.max():.4f}, mean {err.mean():.5f}, logit std {ref.std():.3f}, "
            f"violations {(err > ATOL + RTOL * np.abs(ref)).sum()} / {err.size}")


def test_forward_default_policy_meets_stated_tolerance(vit, weights224, ref16):
    """vit_cuda_init's default (VIT_PREC_AUTO): every logit within the stated 2e-2 + 1e-2*|ref| of the oracle (== the
    reference's ViT_seq, bit for bit) and top-1 identical on every image; it ran on FP16 operands without a fallback."""
    imgs, ref = ref16
    with vit.Engine(weights224, 224, max_batch=8) as eng:  # 2 passes of 8
        got, top1 = eng.forward(imgs, want_top1=True)
        info = eng.info()
    print("default policy", info, _report(got, ref))
    assert info["precision_policy"] == "auto" and info["precision"] == "fp16" and info["precision_fallbacks"] == 0
    _assert_strict(got, top1, ref, "default policy")


def test_forward_fp16_operands_meets_stated_tolerance(vit, weights224, ref16):
    """Explicit FP16 operands / FP32 accumulate: the same bits as the default policy."""
    imgs, ref = ref16
    with vit.Engine(weights224, 224, max_batch=8, precision=vit.PREC_FP16) as eng:
        got, top1 = eng.forward(imgs, want_top1=True)
    with vit.Engine(weights224, 224, max_batch=8) as eng:
        auto = eng.forward(imgs)
    print("fp16", _report(got, ref))
    _assert_strict(got, top1, ref, "fp16")
    assert np.array_equal(got, auto)


def test_forward_bf16_operands(vit, weights224, ref16):
    """BF16 operands / FP32 accumulate, the NON-default variant (also the overflow fallback of the default policy).  On
    these random-init weights (logit std 1.04, i.e. no confident class) 8-bit-mantissa operand rounding through 12
    layers leaves a fraction of a percent of the logits just outside 2e-2 (max |dlogit| ~0.03 -- what SURVEY.md F8 /
    App. E measured for this policy), which is why it is not the default; its own bound is BF16_ATOL + 1e-2 |ref| on
    every logit, >= 99 % inside the default's tolerance, top-1 identical wherever the oracle is decisive."""
    imgs, ref = ref16
    with vit.Engine(weights224, 224, max_batch=8, precision=vit.PREC_BF16) as eng:
        got, top1 = eng.forward(imgs, want_top1=True)
        assert eng.info()["precision"] == "bf16"
    print("bf16", _report(got, ref))
    _assert_bf16_variant(got, top1, ref, "bf16")


def test_forward_384_key_blocked_attention(vit, oracle):
    """BASELINE.json configs[4]: 384x384 images, 577 tokens (random-init pos_embedding [577,768]), which runs
    the key-blocked attention kernel.  Both operand precisions against the oracle parameterised on
    img_size (the reference hard-codes 224 as a #define, ViT_seq.c:10)."""
    w = vit.synth_weights(384, 42)
    imgs = vit.synth_images(3, 384, 7)
    ref = oracle.forward(w, imgs, 384)
    for prec, name in ((vit.PREC_AUTO, "auto"), (vit.PREC_BF16, "bf16")):
        with vit.Engine(w, 384, max_batch=2, precision=prec) as eng:   # passes of 2 + 1
            got, top1 = eng.forward(imgs, want_top1=True)
        print(name, "384", _report(got, ref))
        if prec == vit.PREC_AUTO:
            _assert_strict(got, top1, ref, "default policy 384")
        else:
            _assert_bf16_variant(got, top1, ref, "bf16 384")


@pytest.mark.parametrize("img_size,n", [(32, 5), (64, 3), (208, 2), (240, 2)])
def test_forward_odd_image_sizes(vit, oracle, img_size, n):
    """Edge geometries: 32x32 (5 tokens: one 16-key chunk, one row quarter), 64x64 (17 tokens), 208x208 (170 tokens,
    a second query tile of 42 rows) and 240x240 (226 tokens: the smallest size that takes the key-blocked attention
    kernel, with a one-chunk last key block).  Default policy against the oracle within the stated tolerance."""
    w = vit.synth_weights(img_size, 42)
    imgs = vit.synth_images(n, img_size, 7)
    ref = oracle.forward(w, imgs, img_size)
    with vit.Engine(w, img_size, max_batch=2) as eng:
        got, top1 = eng.forward(imgs, want_top1=True)
        assert eng.info()["attention_fallbacks"] == 0 and eng.info()["precision_fallbacks"] == 0
    print(img_size, _report(got, ref))
    _assert_strict(got, top1, ref, f"default policy {img_size}")


def test_empty_and_oversized_requests(vit, weights224):
    with vit.Engine(weights224, 224, max_batch=1) as eng:      # 70 passes of one image: any number of passes must work
        imgs = vit.synth_images(70, 224, 7)
        many = eng.forward(imgs)
    with vit.Engine(weights224, 224, max_batch=64) as eng:
        assert np.array_equal(many, eng.forward(imgs))
    with vit.Engine(weights224, 224, max_batch=2) as eng:
        out = eng.forward(np.empty((0, 3, 224, 224), dtype=np.float32))
        assert out.shape == (0, 1000)
        d = vit.dev_alloc(0, 3 * 3 * 224 * 224 * 4)
        with pytest.raises(vit.VitCudaError):
            eng.enqueue_device(d, 3, d)          # more than max_batch on the device-resident path
        vit.dev_free(0, d)
    w640 = None
    with pytest.raises(vit.VitCudaError) as ei:     # 26x26 patches + 1 = 677 tokens > 640
        bad = vit.synth_weights(416, 42)
        with vit.Engine(bad, 416, max_batch=1) as eng:
            eng.forward(vit.synth_images(1, 416, 7))
    assert "tokens" in str(ei.value)


def test_against_committed_reference_golden_vectors(vit, weights224):
    """The CUDA path against tests/golden/vit_seq_probs.npz -- outputs of the reference's OWN compiled ViT_seq()
    (tests/golden/make_golden.py), without the oracle in between: same top-1, top-8 probabilities within the
    reference comparator's 0.01 (comparator.c:70) and within what the stated logit tolerance allows."""
    from pathlib import Path
    gold = np.load(Path(__file__).resolve().parent / "golden" / "vit_seq_probs.npz")
    imgs = vit.synth_images(3, 224, int(gold["images_seed"]))
    for prec in (vit.PREC_FP16, vit.PREC_BF16):
        with vit.Engine(weights224, 224, max_batch=4, precision=prec) as eng:
            logits = eng.forward(imgs)
        probs = np.empty_like(logits)
        for i in range(3):
            vit.lib.vit_softmax(vit.fptr(logits[i]), vit.fptr(probs[i]), 1000)
        got = np.take_along_axis(probs, gold["top_idx"].astype(np.int64), 1)
        assert np.abs(got - gold["top_prob"]).max() <= 0.01
        assert np.allclose(got, gold["top_prob"], rtol=0.06, atol=1e-5)        # |dlogit| <= ~0.03 => |dp|/p <= ~0.06
        assert np.array_equal(probs.argmax(1), gold["top_idx"][:, 0]) or np.all(
            gold["top_prob"][:, 0] - gold["top_prob"][:, 1] < 0.06 * gold["top_prob"][:, 0])


def test_full_size_batch_1024_properties(vit, weights224):
    """BASELINE.json configs[2] size (1024 images, 201 728 token rows, every kernel at its full grid) through
    size-independent properties: the batch is 64 distinct images repeated 16 times, so (1) all 16 copies of an
    image must give bit-identical logits wherever they sit in the batch, (2) they must equal the logits of the
    same image in a 64-image forward (different pass sizes, tile boundaries and row positions), and (3) the
    checksum of the per-image checksums is reproducible run to run."""
    base = vit.synth_images(64, 224, 7)
    big = np.ascontiguousarray(np.tile(base, (16, 1, 1, 1)))
    with vit.Engine(weights224, 224, max_batch=1024) as eng:
        small = eng.forward(base)
        out1 = eng.forward(big)
        out2 = eng.forward(big)
        # run-to-run determinism under load: an unordered shared-memory exchange in the attention kernel once showed up
        # as ONE corrupted image in ~70 forwards (tools/determinism_probe.py) -- every forward must reproduce the bits
        for _ in range(12):
            assert np.array_equal(eng.forward(big), out1)
        assert eng.info()["attention_fallbacks"] == 0
    assert np.isfinite(out1).all()
    assert np.array_equal(out1, out2)
    assert np.array_equal(out1.reshape(16, 64, 1000), np.broadcast_to(small, (16, 64, 1000)))
    assert float(out1.astype(np.float64).sum(1).sum()) == float(out2.astype(np.float64).sum(1).sum())


def test_class_row_pruning_of_the_last_layer_keeps_the_logits(vit, oracle, weights224, ref16):
    """Default engine: the last layer runs out_proj / LayerNorm / MLP only for the class rows (the only ones the head
    reads).  Against the all-rows computation the logits differ only by the class-row attention's fp32 softmax
    (instead of bf16 P on the tensor cores), i.e. far less than either differs from the oracle."""
    imgs, ref = ref16
    for prec in (vit.PREC_FP16, vit.PREC_BF16):
        with vit.Engine(weights224, 224, max_batch=16, precision=prec) as eng:
            assert eng.info()["class_row_pruning"] == 1
            pruned = eng.forward(imgs)
            eng.set_class_row_pruning(False)
            full = eng.forward(imgs)
            one = eng.forward(np.ascontiguousarray(imgs[3:4]))
            eng.set_class_row_pruning(True)
            one_p = eng.forward(np.ascontiguousarray(imgs[3:4]))
        print("pruned vs full", np.abs(pruned - full).max(), "| pruned", _report(pruned, ref), "| full", _report(full, ref))
        assert np.abs(pruned - full).max() < 5e-3
        assert np.abs(pruned - ref).mean() <= 1.05 * np.abs(full - ref).mean() + 1e-5
        assert np.array_equal(full[3:4], one) and np.array_equal(pruned[3:4], one_p)   # position independence in both modes
    # 384x384 (key-blocked attention in the other layers)
    w = vit.synth_weights(384, 42)
    im = vit.synth_images(2, 384, 7)
    r = oracle.forward(w, im, 384)
    with vit.Engine(w, 384, max_batch=2) as eng:
        got = eng.forward(im)
    assert np.all(np.abs(got - r) <= ATOL + RTOL * np.abs(r)), _report(got, r)


def test_forward_with_trained_like_layernorm_parameters(vit, oracle, weights224):
    """The LayerNorm tensors the reference ships (Network/Weight_*_ln_*.bin; the large GEMM weights are missing from the
    mount) look nothing like random init: gains with mean 0.03 (layer 0 ln_1) ... 0.57 (layer 10 ln_2) and a
    spread as large as the mean, some NEGATIVE, biases up to +-0.3, final gain 0.69.  Those scales go straight into the
    folded weights W' = ln_w (.) W and the bias vector c = b + W ln_b of the LayerNorm-folded GEMMs, so the whole model is
    checked against the oracle with LayerNorm parameters drawn from the measured per-layer statistics."""
    rng = np.random.default_rng(2024)
    w = [a.copy() for a in weights224]
    ln1_mean = [0.03, 0.08, 0.13, 0.16, 0.19, 0.20, 0.23, 0.23, 0.25, 0.22, 0.23, 0.26]
    ln1_std = [0.07, 0.09, 0.13, 0.15, 0.18, 0.17, 0.17, 0.15, 0.12, 0.10, 0.09, 0.10]
    ln2_mean = [0.14, 0.19, 0.21, 0.30, 0.35, 0.38, 0.42, 0.48, 0.53, 0.57, 0.58, 0.55]
    ln2_std = [0.19, 0.20, 0.21, 0.25, 0.26, 0.26, 0.23, 0.23, 0.19, 0.13, 0.10, 0.06]
    for l in range(12):
        b = 4 + 12 * l
        w[b + 0] = (ln1_mean[l] + ln1_std[l] * rng.standard_normal(768)).astype(np.float32)
        w[b + 1] = (0.035 * rng.standard_normal(768)).astype(np.float32)
        w[b + 6] = (ln2_mean[l] + ln2_std[l] * rng.standard_normal(768)).astype(np.float32)
        w[b + 7] = (0.09 * rng.standard_normal(768)).astype(np.float32)
    w[148] = (0.69 + 0.125 * rng.standard_normal(768)).astype(np.float32)
    w[149] = (0.032 * rng.standard_normal(768)).astype(np.float32)
    w = [np.ascontiguousarray(np.round(a.astype(np.float64) * 1e6) / 1e6, dtype=np.float32) for a in w]   # the loader's rounding
    imgs = vit.synth_images(8, 224, 11)
    ref = oracle.forward(w, imgs, 224)
    for prec, name in ((vit.PREC_AUTO, "auto"), (vit.PREC_BF16, "bf16")):
        with vit.Engine(w, 224, max_batch=8, precision=prec) as eng:
            got, top1 = eng.forward(imgs, want_top1=True)
            eng.set_class_row_pruning(False)
            full = eng.forward(imgs)
        print(name, "trained-like LN", _report(got, ref), "| logit std", ref.std())
        if prec == vit.PREC_AUTO:
            _assert_strict(got, top1, ref, "default policy, trained-like LN")
            assert np.all(np.abs(full - ref) <= ATOL + RTOL * np.abs(ref))
        else:
            _assert_bf16_variant(got, top1, ref, "bf16, trained-like LN")


def test_parity_on_the_reference_shipped_tensors(vit, oracle, shipped224):
    """BASELINE.json configs[1]: the forward pass on the weight tensors the reference itself ships (116 of 152:
    conv_proj, class token, pos_embedding, all LayerNorms and biases, out_proj, the head -- Network/Weight_*.bin,
    Network.c:119-194; the 36 large GEMM weights missing from the mount come from the seed-42 synthetic set) against
    the oracle on the same tensors: default policy, strict stated tolerance, strict top-1; all rows and pruned."""
    w, n_shipped = shipped224
    assert n_shipped >= 100
    imgs = vit.synth_images(8, 224, 7)
    ref = oracle.forward(w, imgs, 224)
    with vit.Engine(w, 224, max_batch=8) as eng:
        got, top1 = eng.forward(imgs, want_top1=True)
        eng.set_class_row_pruning(False)
        full, top1_full = eng.forward(imgs, want_top1=True)
        one = eng.forward(np.ascontiguousarray(imgs[2:3]))
        info = eng.info()
    print(f"shipped tensors ({n_shipped} of 152):", info["precision"], _report(got, ref), "| logit std", float(ref.std()),
          "| top-2 margin min", float(np.min(np.sort(ref, 1)[:, -1] - np.sort(ref, 1)[:, -2])))
    assert info["precision"] == "fp16" and info["precision_fallbacks"] == 0
    _assert_strict(got, top1, ref, "shipped tensors")
    _assert_strict(full, top1_full, ref, "shipped tensors, all rows")
    assert np.array_equal(full[2:3], one)
    with vit.Engine(w, 224, max_batch=8, precision=vit.PREC_BF16) as eng:
        bf, top1_bf = eng.forward(imgs, want_top1=True)
    print("shipped tensors, bf16 variant:", _report(bf, ref))
    _assert_bf16_variant(bf, top1_bf, ref, "shipped tensors, bf16")


def test_precision_auto_falls_back_to_bf16_on_fp16_overflow(vit, oracle, weights224, ref16):
    """VIT_PREC_AUTO: FP16 operands until something overflows.  One mlp_0 bias of layer 5 is set to 1e5, so that hidden unit
    is ~1e5 for every token -- beyond FP16's 65504 (inf in the hidden activation, NaN in the residual stream after mlp_3,
    NaN logits), fine in BF16 and in the fp32 oracle.  The classifier kernel flags the non-finite logits and
    vit_cuda_forward repeats the call on the BF16 operand set: same bits as an explicit BF16 engine, finite, inside the
    BF16 variant's bound of the oracle.  Explicit FP16 reports VIT_E_RANGE; the device-resident API reports it once and
    has switched the engine over.  A WEIGHT beyond the FP16 range selects BF16 at init (AUTO) or fails init (FP16)."""
    imgs, _ = ref16
    imgs = np.ascontiguousarray(imgs[:4])
    w = [a.copy() for a in weights224]
    w[4 + 12 * 5 + 9][7] = 1.0e5                      # layer 5 mlp_0.bias[7]
    ref = oracle.forward(w, imgs, 224)
    assert np.isfinite(ref).all()
    with vit.Engine(w, 224, max_batch=4, precision=vit.PREC_BF16) as eng:
        want, top1 = eng.forward(imgs, want_top1=True)
    # (a 1e5 activation swamps every token row from layer 5 on, so this is not a precision test: the BF16 result only has
    # to be a sane answer to the same question the fp32 oracle answers)
    print("bf16 with a 1e5 hidden unit", _report(want, ref))
    assert np.isfinite(want).all() and np.abs(want - ref).max() < 0.25
    with vit.Engine(w, 224, max_batch=4) as eng:
        got = eng.forward(imgs)
        info = eng.info()
        assert info["precision_policy"] == "auto" and info["precision_fallbacks"] == 1 and info["precision"] == "fp16", info
        assert info["attention_fallbacks"] == 0 and info["attention_exact"] == 0, info    # the softmax was not the cause
        for _ in range(2):
            assert np.array_equal(eng.forward(imgs), want)
        assert eng.info()["precision_fallbacks"] == 3 and eng.info()["precision"] == "bf16"   # needed three times: stays on BF16
        assert np.array_equal(eng.forward(imgs), want) and eng.info()["precision_fallbacks"] == 3
    assert np.isfinite(got).all() and np.array_equal(got, want)
    with vit.Engine(w, 224, max_batch=4, precision=vit.PREC_FP16) as eng:
        with pytest.raises(vit.VitCudaError) as ei:
            eng.forward(imgs)
        assert ei.value.code == -6 and "FP16" in str(ei.value)
    with vit.Engine(w, 224, max_batch=4) as eng:      # device-resident API
        d_imgs, d_logits = vit.dev_alloc(0, imgs.nbytes), vit.dev_alloc(0, 4 * 1000 * 4)
        vit.dev_upload(0, d_imgs, imgs)
        # NaN rows also trip the softmax's range check, so the engine may first try the exact softmax (VIT_E_RANGE #1), find
        # the logits still non-finite, take that switch back and move to BF16 (VIT_E_RANGE #2); then the pass stands
        messages = []
        for _ in range(4):
            eng.enqueue_device(d_imgs, 4, d_logits)
            try:
                eng.sync()
                break
            except vit.VitCudaError as ex:
                assert ex.code == -6
                messages.append(str(ex))
        assert 1 <= len(messages) <= 2 and "BF16" in messages[-1], messages
        out = np.empty((4, 1000), dtype=np.float32)
        vit.dev_download(0, out, d_logits)
        vit.dev_free(0, d_imgs)
        vit.dev_free(0, d_logits)
        info = eng.info()
        assert info["precision"] == "bf16" and info["attention_exact"] == 0 and info["attention_fallbacks"] == 0, info
    assert np.array_equal(out, want)
    big = [a.copy() for a in weights224]
    big[4 + 12 * 3 + 4][5] = 1.0e6                     # layer 3 out_proj.weight: not representable in FP16
    with pytest.raises(vit.VitCudaError) as ei:
        vit.Engine(big, 224, max_batch=2, precision=vit.PREC_FP16)
    assert ei.value.code == -6
    with vit.Engine(big, 224, max_batch=2) as eng:
        assert eng.info()["precision"] == "bf16" and eng.info()["precision_policy"] == "auto"
        assert np.isfinite(eng.forward(imgs[:2].copy())).all()


def test_engine_options_unfused_layernorm_pdl_graphs(vit, weights224, ref16):
    """The run-time switches of vit_cuda_set_option.  Launch mechanics must not change a bit: programmatic dependent
    launch off, CUDA-graph replay of small passes off (and the first, un-captured run of a small pass against its
    replays).  Separate warp-per-row LayerNorm kernels instead of the folded form (VIT_OPT_LN_FUSED = 0: the
    unfolded GEMM variants and layernorm_kernel inside the whole model) change the rounding points, so that
    configuration is held to the stated tolerance against the oracle instead."""
    imgs, ref = ref16
    imgs8 = np.ascontiguousarray(imgs[:8])
    with vit.Engine(weights224, 224, max_batch=8) as eng:
        base, top1 = eng.forward(imgs8, want_top1=True)
        small = [eng.forward(np.ascontiguousarray(imgs[3:5])) for _ in range(3)]     # plain launches, capture, replay
        assert np.array_equal(small[0], small[1]) and np.array_equal(small[0], small[2]) and np.array_equal(small[0], base[3:5])
        for opt in (vit.OPT_PDL, vit.OPT_GRAPHS, vit.OPT_WAVE_PASSES):
            assert eng.get_option(opt) == 1
            eng.set_option(opt, 0)
            assert np.array_equal(eng.forward(imgs8), base), f"option {opt} changed the result"
            assert np.array_equal(eng.forward(np.ascontiguousarray(imgs[3:5])), base[3:5])
            eng.set_option(opt, 1)
        eng.set_option(vit.OPT_LN_FUSED, 0)
        unfused, top1_u = eng.forward(imgs8, want_top1=True)
        eng.set_class_row_pruning(False)
        unfused_full = eng.forward(imgs8)
        eng.set_option(vit.OPT_LN_FUSED, 1)
        eng.set_class_row_pruning(True)
        assert np.array_equal(eng.forward(imgs8), base)
        with pytest.raises(vit.VitCudaError):
            eng.set_option(99, 1)
    print("unfused LayerNorm", _report(unfused, ref[:8]), "| vs folded", float(np.abs(unfused - base).max()))
    _assert_strict(base, top1, ref[:8], "folded LayerNorm")
    _assert_strict(unfused, top1_u, ref[:8], "separate LayerNorm kernels")
    assert np.array_equal(unfused, unfused_full)      # without the folded form there is no class-row pruning: same path
    vit.lib.vit_cuda_set_option(vit.OPT_LN_FUSED, 1)


def test_fp16_residual_stream_option(vit, oracle, weights224, shipped224, ref16):
    """VIT_OPT_RESIDUAL16 (default on): with FP16 operands the patch rows' residual stream is held in FP16 (the rows out_proj /
    mlp_3 update in place are the next GEMM's operand) and each image's class-token row keeps an fp32 master copy.  Both
    settings are held to the stated tolerance against the oracle (synthetic and shipped tensors, pruned and all rows); the
    default must stay independent of batch position and pass size, switching it off and on again must give the same bits,
    and BF16 operands must ignore it."""
    imgs, ref = ref16
    with vit.Engine(weights224, 224, max_batch=16) as eng:
        assert eng.get_option(vit.OPT_RESIDUAL16) == 1
        got, top1 = eng.forward(imgs, want_top1=True)
        one = eng.forward(np.ascontiguousarray(imgs[5:6]))
        rev = eng.forward(np.ascontiguousarray(imgs[::-1]))[::-1]
        eng.set_class_row_pruning(False)
        full, top1_full = eng.forward(imgs, want_top1=True)
        xb = np.ascontiguousarray(np.random.default_rng(5).standard_normal((2 * 197, 768)).astype(np.float32))
        blk = vit.op_encoder_block(xb, 2, 3)
        eng.set_class_row_pruning(True)
        try:
            eng.set_option(vit.OPT_RESIDUAL16, 0)
            wide, top1_wide = eng.forward(imgs, want_top1=True)
            blk_wide = vit.op_encoder_block(xb, 2, 3)
        finally:
            eng.set_option(vit.OPT_RESIDUAL16, 1)
        assert np.array_equal(eng.forward(imgs), got)
    print("fp16 residual", _report(got, ref), "| all rows", _report(full, ref), "| fp32 residual", _report(wide, ref))
    assert not np.array_equal(wide, got)              # the switch selects a different set of kernels
    # one block in isolation: the patch rows of the result are FP16 values (half an ulp at 4..8 is 2e-3) of the oracle's block
    blk_ref = np.concatenate([oracle.encoder_block(np.ascontiguousarray(xb[i * 197:(i + 1) * 197]), weights224[4 + 12 * 3: 16 + 12 * 3]) for i in range(2)])
    err, err_wide = np.abs(blk - blk_ref), np.abs(blk_wide - blk_ref)
    print("block 3 alone: max err fp16 stream", float(err.max()), "class rows", float(err[::197].max()), "| fp32 stream", float(err_wide.max()))
    assert np.all(err <= 6e-3 + 1.5e-3 * np.abs(blk_ref))
    assert np.all(err_wide <= 5e-3 + 1e-3 * np.abs(blk_ref))
    assert np.all(err[::197] <= 5e-3 + 1e-3 * np.abs(blk_ref[::197]))      # the class rows are not rounded to FP16
    _assert_strict(got, top1, ref, "fp16 residual stream")
    _assert_strict(full, top1_full, ref, "fp16 residual stream, all rows")
    _assert_strict(wide, top1_wide, ref, "fp32 residual stream")
    assert np.array_equal(got[5:6], one) and np.array_equal(got, rev)
    w, _ = shipped224
    im8 = vit.synth_images(8, 224, 7)
    ref8 = oracle.forward(w, im8, 224)
    with vit.Engine(w, 224, max_batch=8) as eng:
        got8, top8 = eng.forward(im8, want_top1=True)
        try:
            eng.set_option(vit.OPT_RESIDUAL16, 0)
            wide8, topw8 = eng.forward(im8, want_top1=True)
        finally:
            eng.set_option(vit.OPT_RESIDUAL16, 1)
    print("shipped tensors: fp16 residual", _report(got8, ref8), "| fp32 residual", _report(wide8, ref8))
    _assert_strict(got8, top8, ref8, "fp16 residual stream, shipped tensors")
    _assert_strict(wide8, topw8, ref8, "fp32 residual stream, shipped tensors")
    with vit.Engine(weights224, 224, max_batch=16, precision=vit.PREC_BF16) as eng:
        a = eng.forward(imgs)
        try:
            eng.set_option(vit.OPT_RESIDUAL16, 0)
            b = eng.forward(imgs)
        finally:
            eng.set_option(vit.OPT_RESIDUAL16, 1)
    assert np.array_equal(a, b)


def test_operand_weight_cache_round_trip(vit, weights224, ref16, tmp_path):
    """vit_cuda_save_weight_cache / vit_cuda_init_from_cache: an engine started from the cache file (one read, one
    host-to-device copy, no conversion, no LayerNorm folding) returns the same bits; a truncated or corrupted file and a
    foreign file are refused."""
    import time
    imgs, _ = ref16
    imgs = np.ascontiguousarray(imgs[:4])
    path = tmp_path / "vit_b16_224.operands"
    t0 = time.perf_counter()
    with vit.Engine(weights224, 224, max_batch=4) as eng:
        t_init = time.perf_counter() - t0
        want = eng.forward(imgs)
        eng.save_weight_cache(str(path))
        mib = eng.info()["weights_mib"]
    assert abs(path.stat().st_size / 2**20 - mib) < 2
    t0 = time.perf_counter()
    with vit.Engine(None, max_batch=4, cache=str(path)) as eng:
        t_cache = time.perf_counter() - t0
        info = eng.info()
        got = eng.forward(imgs)
    print(f"init from 152 fp32 tensors {t_init:.3f} s, from the operand cache {t_cache:.3f} s ({mib} MiB)")
    assert info["precision_policy"] == "auto" and info["tokens"] == 197
    assert np.array_equal(got, want)
    raw = bytearray(path.read_bytes())
    raw[len(raw) // 2] ^= 0x40
    bad = tmp_path / "corrupt.operands"
    bad.write_bytes(bytes(raw))
    short = tmp_path / "short.operands"
    short.write_bytes(bytes(raw[:len(raw) // 3]))
    foreign = tmp_path / "foreign.operands"
    foreign.write_bytes(b"not a cache" * 100)
    for f in (bad, short, foreign, tmp_path / "missing.operands"):
        with pytest.raises(vit.VitCudaError):
            vit.Engine(None, max_batch=4, cache=str(f))
    with vit.Engine(weights224, 224, max_batch=4, precision=vit.PREC_BF16) as eng:   # a single-precision cache
        want_bf = eng.forward(imgs)
        eng.save_weight_cache(str(path))
    with vit.Engine(None, max_batch=4, cache=str(path)) as eng:
        assert eng.info()["precision_policy"] == "bf16" and np.array_equal(eng.forward(imgs), want_bf)


@pytest.mark.parametrize("layer", [0, 7, 11])
def test_encoder_block_against_the_oracle(vit, oracle, weights224, shipped224, layer):
    """Encoder (ViT_seq.c:271-302; pre-LN block: r = x + MHA(LN1 x), out = r + MLP(LN2 r)) as the forward pass runs it --
    in_proj with LayerNorm folded in, fused attention, out_proj + residual, mlp_0 with LayerNorm folded in + GELU, mlp_3 +
    residual -- on one block in isolation, against the oracle's restatement on the same fp32 input.  Synthetic weights and
    the reference's shipped tensors (real LayerNorm gains / biases, out_proj); folded and unfolded LayerNorm; both operand
    types.  The block's output carries the input through the residual path, so the bound is on the block's own
    contribution (|y - x| up to ~4.5): FP16 5e-3 + 1e-3 |ref| (measured max 0.0021), BF16 3e-2 + 6e-3 |ref| (measured 0.015)."""
    rng = np.random.default_rng(100 + layer)
    batch, tokens = 3, 197
    x = (rng.standard_normal((batch * tokens, 768)) * 1.3 + 0.2 * rng.standard_normal((1, 768))).astype(np.float32)
    for w, name in ((weights224, "synthetic"), (shipped224[0], "shipped")):
        lw = w[4 + 12 * layer: 16 + 12 * layer]
        ref = np.concatenate([oracle.encoder_block(np.ascontiguousarray(x[i * tokens:(i + 1) * tokens]), lw) for i in range(batch)])
        for prec, atol, rtol in ((vit.PREC_FP16, 5e-3, 1e-3), (vit.PREC_BF16, 3e-2, 6e-3)):
            with vit.Engine(w, 224, max_batch=4, precision=prec) as eng:
                got = vit.op_encoder_block(x, batch, layer)
                eng.set_option(vit.OPT_LN_FUSED, 0)
                unfused = vit.op_encoder_block(x, batch, layer)
                eng.set_option(vit.OPT_LN_FUSED, 1)
            for out, what in ((got, "folded"), (unfused, "separate LayerNorm")):
                err = np.abs(out - ref)
                print(f"layer {layer} {name} prec {prec} {what}: max err {err.max():.4f}, |delta| of the block {np.abs(ref - x).max():.2f}")
                assert np.all(err <= atol + rtol * np.abs(ref)), f"layer {layer} {name} prec {prec} {what}: max err {err.max():.4f}"


def test_batch_position_independence(vit, weights224, ref16):
    """An image's logits must not depend on its position in the batch or on the pass size
    (needed for bit-identical results across GPU counts)."""
    imgs, _ = ref16
    with vit.Engine(weights224, 224, max_batch=16) as eng:
        a = eng.forward(imgs)
        b = eng.forward(np.ascontiguousarray(imgs[::-1]))[::-1]
        c = eng.forward(np.ascontiguousarray(imgs[5:6]))
    assert np.array_equal(a, b)
    assert np.array_equal(a[5:6], c)


def test_single_pass_softmax_matches_exact_and_falls_back(vit, weights224, ref16):
    """The default single-pass softmax (exponent offset from the first 16 scores) against the exact
    two-pass kernel: same logits up to fp32 round-off of the shifted exponentials on ordinary data, and --
    with in_proj Q/K weights scaled so that logit gaps are in the hundreds -- the kernel must flag the
    range violation and vit_cuda_forward must transparently return the exact kernel's result."""
    imgs, ref = ref16
    imgs = np.ascontiguousarray(imgs[:6])
    with vit.Engine(weights224, 224, max_batch=8) as eng:
        fast = eng.forward(imgs)
        assert eng.info()["attention_fallbacks"] == 0 and eng.info()["attention_exact"] == 0
        eng.set_attention_exact(True)
        exact = eng.forward(imgs)
    # a different exponent offset means a different bf16 rounding of every P, so the two agree only as
    # well as either agrees with the oracle (the BF16 noise floor), not bit for bit
    ref6 = ref[:6]
    e_fast, e_exact = np.abs(fast - ref6), np.abs(exact - ref6)
    print("single-pass vs exact", np.abs(fast - exact).max(), "fast vs oracle", _report(fast, ref6), "| exact vs oracle", _report(exact, ref6))
    assert np.abs(fast - exact).max() < 2 * ATOL
    assert e_fast.mean() < 1.25 * e_exact.mean() + 1e-4 and np.all(e_fast <= 2 * ATOL + RTOL * np.abs(ref6))
    wild = [w.copy() for w in weights224]
    wild[6][:1536 * 768] *= 16.0   # layer 0 in_proj (flat [2304][768]): Q and K rows -> scores x 256
    with vit.Engine(wild, 224, max_batch=8) as eng:
        got = eng.forward(imgs)
        assert eng.info()["attention_fallbacks"] == 1 and eng.info()["precision_fallbacks"] == 0 and eng.info()["precision"] == "fp16"
        eng.set_attention_exact(True)
        want = eng.forward(imgs)
    assert np.isfinite(got).all() and np.array_equal(got, want)
    # device-resident API: the pass is reported invalid and the engine switches itself over
    with vit.Engine(wild, 224, max_batch=8) as eng:
        d_imgs, d_logits = vit.dev_alloc(0, imgs.nbytes), vit.dev_alloc(0, 6 * 1000 * 4)
        vit.dev_upload(0, d_imgs, imgs)
        eng.enqueue_device(d_imgs, 6, d_logits)
        with pytest.raises(vit.VitCudaError) as ei:
            eng.sync()
        assert "exact" in str(ei.value)
        eng.enqueue_device(d_imgs, 6, d_logits)
        eng.sync()
        out = np.empty((6, 1000), dtype=np.float32)
        vit.dev_download(0, out, d_logits)
        vit.dev_free(0, d_imgs)
        vit.dev_free(0, d_logits)
    assert np.array_equal(out, want)


def test_scattered_images_and_pinned_logits_paths(vit, weights224, ref16):
    """vit_cuda_forward_scattered (one allocation per image, as the reference's loader produces them) and the
    pinned / pageable result paths of vit_cuda_forward all return the same bits."""
    imgs, _ = ref16
    with vit.Engine(weights224, 224, max_batch=8) as eng:     # 16 images: passes of 8
        base = eng.forward(imgs)                                # pageable logits -> pinned staging inside the engine
        parts = [np.ascontiguousarray(imgs[i]).copy() for i in range(len(imgs))]
        scat, top1 = eng.forward_scattered(parts, want_top1=True)
        h_logits, h_ptr = vit.pinned_empty((len(imgs), 1000))
        h_imgs, h_imgs_ptr = vit.pinned_empty(imgs.shape)
        h_imgs[...] = imgs
        eng.forward_raw(h_imgs_ptr, len(imgs), h_ptr)           # pinned in, pinned out: direct copies
        pinned = h_logits.copy()
        vit.pinned_free(h_ptr)
        vit.pinned_free(h_imgs_ptr)
    assert np.array_equal(base, scat) and np.array_equal(base, pinned)
    assert np.array_equal(top1, base.argmax(1))


def test_reference_signature_adaptor(vit, weights224, ref16, oracle, tmp_path):
    """ViT_cuda(ImageData*, Network*, float**) + result file + comparator, the Main.c flow."""
    imgs, ref = ref16
    n = 4
    net = vit.as_network(weights224)
    arr = (vit.ImageData * n)()
    for i in range(n):
        arr[i].n, arr[i].c, arr[i].h, arr[i].w = n, 3, 224, 224
        arr[i].data = vit.fptr(imgs[i])
    probs = np.zeros((n, 1000), dtype=np.float32)
    rows = (C.POINTER(C.c_float) * n)(*[vit.fptr(probs[i]) for i in range(n)])
    assert vit.lib.initialize_cuda() == 0
    vit.lib.ViT_cuda(arr, net, rows)
    assert vit.lib.ViT_cuda_status() == 0, vit.lib.vit_cuda_last_error()
    vit.lib.Release_cuda()
    ref_probs = oracle.softmax(ref[:n])
    assert np.array_equal(probs.argmax(1), ref_probs.argmax(1))
    # the reference's acceptance rule (comparator.c:64-74) on files in the reference's format
    res, ans = tmp_path / "cuda_result.txt", tmp_path / "answer_result.txt"
    rows_ref = (C.POINTER(C.c_float) * n)(*[vit.fptr(ref_probs[i]) for i in range(n)])
    assert vit.lib.write_results(str(res).encode(), rows, n) == 0
    assert vit.lib.write_results(str(ans).encode(), rows_ref, n) == 0
    assert vit.lib.comparator_files(str(res).encode(), str(ans).encode(), n) == 0


def test_adaptor_notices_weights_reloaded_into_the_same_array(vit, weights224, ref16):
    """The reference hands ONE static Network[152] array to every call (Main.c:29,57), so the array's address says nothing
    about its contents.  ViT_cuda() fingerprints the tensors (pointers, sizes, 64 sampled elements each) and uploads again
    when they changed; vit_host_invalidate_weights() forces it for an edit the sample cannot see."""
    imgs, _ = ref16
    n = 2
    w = [a.copy() for a in weights224]
    net = vit.as_network(w)
    arr = (vit.ImageData * n)()
    for i in range(n):
        arr[i].n, arr[i].c, arr[i].h, arr[i].w = n, 3, 224, 224
        arr[i].data = vit.fptr(imgs[i])
    probs = np.zeros((n, 1000), dtype=np.float32)
    rows = (C.POINTER(C.c_float) * n)(*[vit.fptr(probs[i]) for i in range(n)])
    vit.lib.vit_host_invalidate_weights.restype = None
    assert vit.lib.initialize_cuda() == 0
    vit.lib.ViT_cuda(arr, net, rows)
    assert vit.lib.ViT_cuda_status() == 0
    first = probs.copy()
    w[151][:] = 0.0
    w[151][::2] = 30.0                  # a reload into the same buffers: every even class gets +30 on its logit
    vit.lib.ViT_cuda(arr, net, rows)
    assert vit.lib.ViT_cuda_status() == 0
    assert np.all(probs.argmax(1) % 2 == 0) and not np.allclose(probs, first)
    second = probs.copy()
    w[151][7] = 100.0                   # one element between the sampled ones: not seen ...
    vit.lib.ViT_cuda(arr, net, rows)
    assert np.array_equal(probs, second)
    vit.lib.vit_host_invalidate_weights()   # ... until the caller says so
    vit.lib.ViT_cuda(arr, net, rows)
    assert vit.lib.ViT_cuda_status() == 0 and np.all(probs.argmax(1) == 7)
    # failure paths fill prb with NaN and set the status, never exit()
    arr[0].c = 4
    vit.lib.ViT_cuda(arr, net, rows)
    assert vit.lib.ViT_cuda_status() != 0 and np.isnan(probs).all()
    vit.lib.Release_cuda()


def test_errors_are_reported_not_fatal(vit, weights224):
    bad = [w for w in weights224]
    bad[6] = bad[6][:-1].copy()
    with pytest.raises(vit.VitCudaError) as ei:
        vit.Engine(bad, 224, max_batch=2)
    assert "tensor 6" in str(ei.value)


def test_plain_c_driver_on_100_images_through_the_loader(vit, oracle, shipped224, tmp_path):
    """Config 1 of BASELINE.json with a stand-in for the missing Data/input-100.bin: the FIRST 100 images of the seeded
    stream (no selection) and the reference's shipped weight tensors (+ synthetic for the 36 missing ones) are written in
    the reference's file formats (Network.c:36-58, Weight_<idx>_<name>.bin); the plain-C driver (host/vit_main.c, the
    Main.c flow) loads them with load_image_data / load_weights, runs ViT_cuda(), writes the result file in the Main.c:71
    format and applies the comparator rule (label exact, |dprob| <= 0.01, 100 lines; comparator.c:64-74) against answers
    produced by the oracle.  Labels must be identical on all 100 lines for the default policy.  (Random-init GEMM weights
    give near-uniform logits, so a near-tie closer than the FP16 noise of ~5e-3 can exist among 100 images; if the oracle
    itself has such a tie the test says so and requires the comparator's difference count to equal exactly those lines.)
    Also through --stream (chunked reader) and from the operand cache: same result file, byte for byte."""
    import subprocess
    from pathlib import Path
    exe = Path(vit.PKG_DIR) / "bin" / "vit_main"
    assert exe.exists(), "build with make -C vision-transformer-opencl_b200"
    n = 100
    w, _ = shipped224
    imgs = vit.synth_images(n, 224, 7)
    logits = oracle.forward(w, imgs, 224)
    probs = oracle.softmax(logits)
    srt = np.sort(logits, 1)
    ties = np.flatnonzero(srt[:, -1] - srt[:, -2] < 1.0e-2)       # two classes closer than twice the FP16 noise
    print(f"oracle top-2 margins: min {float((srt[:, -1] - srt[:, -2]).min()):.4f}, near-ties (< 0.01): {ties.tolist()}")
    img_file, wdir = tmp_path / "input-100.bin", tmp_path / "Network"
    assert vit.lib.save_image_data(str(img_file).encode(), vit.fptr(imgs), n, 3, 224, 224) == 0
    assert vit.lib.save_weights(str(wdir).encode(), vit.as_network(w), 152, 224) == 0
    rows = (C.POINTER(C.c_float) * n)(*[vit.fptr(probs[i]) for i in range(n)])
    ans, res = tmp_path / "answer_result.txt", tmp_path / "cuda_result.txt"
    assert vit.lib.write_results(str(ans).encode(), rows, n) == 0
    base = [str(exe), "--images", str(img_file), "--weights", str(wdir), "--max-batch", "64", "--answer", str(ans)]
    out = subprocess.run(base + ["--result", str(res), "--timing", "--operand-cache", str(tmp_path / "w.operands")],
                         capture_output=True, text=True, timeout=600)
    print(out.stdout[-1500:])
    if len(ties) == 0:
        assert out.returncode == 0, out.stdout + out.stderr
        assert f"match on {n} lines" in out.stdout
    else:
        got_labels = [int(l.split("label:")[1].split("/")[0]) for l in res.read_text().splitlines()]
        diff = [i for i in range(n) if got_labels[i] != int(logits[i].argmax())]
        assert set(diff) <= set(ties.tolist()), f"label differs on decisive images {sorted(set(diff) - set(ties.tolist()))}"
        assert (f"match on {n} lines" in out.stdout) if not diff else (f"Comparator: {len(diff)} differences" in out.stdout)
    assert "mlp_0 + GELU" in out.stdout and "wrote operand cache" in out.stdout
    first = res.read_bytes()
    res2 = tmp_path / "cuda_result_stream.txt"
    out = subprocess.run(base + ["--result", str(res2), "--stream", "40"], capture_output=True, text=True, timeout=600)
    assert "streamed in chunks of 40" in out.stdout, out.stdout + out.stderr
    assert res2.read_bytes() == first
    res3 = tmp_path / "cuda_result_cache.txt"
    out = subprocess.run(base + ["--result", str(res3), "--operand-cache", str(tmp_path / "w.operands")], capture_output=True, text=True, timeout=600)
    assert "engine up from operand cache" in out.stdout, out.stdout + out.stderr
    assert res3.read_bytes() == first
    res4 = tmp_path / "cuda_result_bf16.txt"
    out = subprocess.run(base + ["--result", str(res4), "--precision", "bf16"], capture_output=True, text=True, timeout=600)
    assert "operands bf16 (policy bf16" in out.stdout, out.stdout + out.stderr


def test_multi_gpu_replicas_are_bit_identical(vit, weights224, ref16):
    """Data-parallel replicas inside one process (vit_cuda_init n_gpus = 2): same logits as one GPU."""
    import subprocess
    n_dev = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU ")
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    imgs, _ = ref16
    with vit.Engine(weights224, 224, max_batch=8, n_gpus=1) as eng:
        one = eng.forward(imgs)
    with vit.Engine(weights224, 224, max_batch=8, n_gpus=2) as eng:
        two = eng.forward(imgs)                               # one feeding thread per GPU
        odd = eng.forward(np.ascontiguousarray(imgs[:5]))     # ragged shards 3 + 2
        scat = eng.forward_scattered([np.ascontiguousarray(imgs[i]).copy() for i in range(len(imgs))])
        eng.set_option(vit.OPT_HOST_THREADS, 0)               # one thread issuing for both GPUs in turn
        serial = eng.forward(imgs)
        eng.set_option(vit.OPT_HOST_THREADS, 1)
    assert np.array_equal(one, two) and np.array_equal(one, scat) and np.array_equal(one, serial)
    assert np.array_equal(one[:5], odd)
