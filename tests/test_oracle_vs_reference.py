"""The oracle against the reference's own objects (ViT_seq.c / Network.c / comparator.c compiled
unmodified into oracle/_ref/libvit_ref.so by oracle/Makefile) run live on fresh inputs.
Skipped -- loudly -- where that library is absent."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libvit_ref.so not built (needs /root/reference at build time)")


def test_vit_seq_bit_exact_on_a_fresh_image(vit, weights224, capfd):
    img = vit.synth_images(1, 224, seed=2024)
    ref = O.ref_vit_seq(weights224, img)           # reference ViT_seq(), ~13 s on one core
    capfd.readouterr()                              # it prints a timing line (ViT_seq.c:371)
    _, probs = O.forward(weights224, img, 224, want_probs=True)
    assert np.array_equal(ref, probs)


def test_reference_loader_agrees_with_ours(vit, tmp_path):
    """load_weights (incl. the 1e-6 rounding, Network.c:185-187) and load_image_data: the reference's
    functions and the POSIX re-implementation read the same files to the same values."""
    rng = np.random.default_rng(5)
    net = (vit.Tensor * 152)()
    keep = []
    for i in (0, 2, 5, 151):
        a = (rng.standard_normal(vit.tensor_numel(i)) * 0.1234567).astype(np.float32)
        keep.append(a)
        net[i].data, net[i].size = vit.fptr(a), a.size
    wdir = tmp_path / "Network"
    assert vit.lib.save_weights(str(wdir).encode(), net, 152, 224) == 0
    ours = (vit.Tensor * 152)()
    assert vit.lib.load_weights(str(wdir).encode(), ours, 152) == 4
    r = O.ref_lib()
    theirs = (O._RefNetwork * 152)()
    r.load_weights(str(wdir).encode(), theirs, 152)
    for i in range(152):
        assert ours[i].size == theirs[i].size
        if ours[i].size:
            a = np.ctypeslib.as_array(ours[i].data, (ours[i].size,))
            b = np.ctypeslib.as_array(theirs[i].data, (theirs[i].size,))
            assert np.array_equal(a, b)
    vit.lib.free_weights(ours, 152)

    imgs = rng.standard_normal((3, 3, 224, 224)).astype(np.float32)
    path = tmp_path / "input-3.bin"
    assert vit.lib.save_image_data(str(path).encode(), vit.fptr(imgs), 3, 3, 224, 224) == 0
    a = vit.lib.load_image_data(str(path).encode())
    b = r.load_image_data(str(path).encode())
    for i in range(3):
        assert (a[i].n, a[i].c, a[i].h, a[i].w) == (b[i].n, b[i].c, b[i].h, b[i].w) == (3, 3, 224, 224)
        assert np.array_equal(np.ctypeslib.as_array(a[i].data, (3 * 224 * 224,)), np.ctypeslib.as_array(b[i].data, (3 * 224 * 224,)))
        assert np.array_equal(np.ctypeslib.as_array(a[i].data, (3, 224, 224)), imgs[i])
    vit.lib.free_image_data(a)


def test_reference_comparator_accepts_our_result_file(vit, tmp_path):
    """comparator() (comparator.c:23-80) reads ./Data/opencl_result.txt and ./Data/answer_result.txt
    relative to the CWD; write_results must produce a file it parses and accepts."""
    probs = np.full((1, 1000), 1e-4, dtype=np.float32)
    probs[0, 65] = 0.919345
    rows = (C.POINTER(C.c_float) * 1)(vit.fptr(probs[0]))
    (tmp_path / "Data").mkdir()
    assert vit.lib.write_results(str(tmp_path / "Data" / "opencl_result.txt").encode(), rows, 1) == 0
    (tmp_path / "Data" / "answer_result.txt").write_text("[0] label: 65 / prob: 0.919345\n")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert O.ref_lib().comparator() == 0
        (tmp_path / "Data" / "answer_result.txt").write_text("[0] label: 66 / prob: 0.919345\n")
        assert O.ref_lib().comparator() == 1
    finally:
        os.chdir(cwd)
