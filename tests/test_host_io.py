"""Host-side C code (no GPU): loaders, result writer, comparator rule, tensor table, synthetic
asset generator."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

REF_NET = Path("/root/reference/Network")


def test_tensor_table_matches_state_dict_layout(vit):
    sizes = [vit.tensor_numel(i) for i in range(152)]
    assert sum(sizes) == 86_567_656  # ViT-B/16 parameter count
    assert vit.tensor_numel(3, 384) == 577 * 768 and vit.tensor_numel(3) == 197 * 768
    assert vit.tensor_name(0) == "class_token" and vit.tensor_name(151) == "heads_head_bias"
    assert vit.tensor_name(4 + 12 * 7 + 8) == "encoder_layers_encoder_layer_7_mlp_0_weight"
    assert vit.tensor_numel(-1) == 0 and vit.tensor_numel(152) == 0


@pytest.mark.skipif(not REF_NET.exists(), reason="reference Network/ directory not mounted")
def test_tensor_table_matches_shipped_weight_files(vit):
    """Every Weight_<idx>_<name>.bin the reference ships has the name and size our table predicts."""
    files = sorted(REF_NET.glob("Weight_*.bin"))
    assert len(files) >= 100
    for f in files:
        idx = int(f.name.split("_")[1])
        assert f.name == f"Weight_{idx}_{vit.tensor_name(idx)}.bin"
        assert f.stat().st_size == 4 * vit.tensor_numel(idx)


@pytest.mark.skipif(not REF_NET.exists(), reason="reference Network/ directory not mounted")
def test_load_shipped_weights_reports_the_missing_ones(vit, capfd):
    net = (vit.Tensor * 152)()
    n = vit.lib.load_weights(str(REF_NET).encode(), net, 152)
    assert n == len(list(REF_NET.glob("Weight_*.bin"))) == 116
    rc = vit.lib.vit_validate_weights(net, 152, 224)
    assert rc == -(6 + 1)  # first absent tensor: layer-0 in_proj_weight
    assert "missing" in capfd.readouterr().err
    w = np.ctypeslib.as_array(net[1].data, (net[1].size,))
    # values are rounded to 6 decimals in fp32 (Network.c:185-187): re-rounding is the identity
    assert np.array_equal(np.float32(np.round(w * np.float32(1e6))) / np.float32(1e6), w)
    vit.lib.free_weights(net, 152)


def test_load_weights_errors(vit, tmp_path, capfd):
    net = (vit.Tensor * 152)()
    assert vit.lib.load_weights(str(tmp_path / "nope").encode(), net, 152) == -1
    (tmp_path / "Weight_abc_x.bin").write_bytes(b"\0" * 8)      # unparsable index: ignored
    (tmp_path / "Weight_999_x.bin").write_bytes(b"\0" * 8)      # out of range: ignored
    (tmp_path / "Weight_5_x.txt").write_bytes(b"\0" * 8)        # wrong extension: ignored
    (tmp_path / "Weight_7_empty.bin").write_bytes(b"")          # empty: ignored
    assert vit.lib.load_weights(str(tmp_path).encode(), net, 152) == 0
    assert vit.lib.load_image_data(str(tmp_path / "nope.bin").encode()) is None or not vit.lib.load_image_data(str(tmp_path / "nope.bin").encode())
    (tmp_path / "short.bin").write_bytes(np.array([2, 3, 4, 4], dtype=np.int32).tobytes() + b"\0" * 10)
    assert not vit.lib.load_image_data(str(tmp_path / "short.bin").encode())
    capfd.readouterr()


def test_result_file_format_and_comparator_rule(vit, tmp_path):
    """Main.c:71 line format; comparator.c:43-74 rule with the line count as an argument."""
    probs = np.full((3, 1000), 1e-4, dtype=np.float32)
    probs[0, 65], probs[1, 795], probs[2, 0] = 0.919345, 0.824735, 0.5
    rows = (C.POINTER(C.c_float) * 3)(*[vit.fptr(probs[i]) for i in range(3)])
    res = tmp_path / "res.txt"
    assert vit.lib.write_results(str(res).encode(), rows, 3) == 0
    assert res.read_text() == "[0] label: 65 / prob: 0.919345\n[1] label: 795 / prob: 0.824735\n[2] label: 0 / prob: 0.500000\n"
    ans = tmp_path / "ans.txt"
    cmp = lambda n: vit.lib.comparator_files(str(res).encode(), str(ans).encode(), n)
    ans.write_text("[0] label: 65 / prob: 0.925000\n[1] label: 795 / prob: 0.824735\n[2] label: 0 / prob: 0.5\n")
    assert cmp(3) == 0                                   # |dprob| = 0.0057 <= 0.01
    ans.write_text("[0] label: 65 / prob: 0.930000\n[1] label: 794 / prob: 0.824735\n[2] label: 1 / prob: 0.6\n")
    assert cmp(3) == 4                                   # prob; label; label + prob
    assert cmp(1) == 1
    ans.write_text("[0] label: 65 / prob: 0.919345\n")
    assert cmp(3) == 1                                   # short file: +1 and stop
    ans.write_text("garbage\n[1] label: 795 / prob: 0.824735\n")
    assert cmp(2) == 1                                   # unparsable line: +1, continue
    assert vit.lib.comparator_files(b"/nonexistent", str(ans).encode(), 1) == 1


def test_softmax_and_argmax(vit):
    x = np.array([1.0, 3.0, 3.0, -2.0], dtype=np.float32)
    p = np.empty_like(x)
    vit.lib.vit_softmax(vit.fptr(x), vit.fptr(p), 4)
    e = np.exp(x - 3.0)
    assert np.allclose(p, e / e.sum(), rtol=1e-6)
    assert vit.lib.vit_argmax(vit.fptr(x), 4) == 1       # lowest index wins ties


def test_synthetic_assets_are_deterministic(vit):
    a = vit.synth_weights(224, 42)
    b = vit.synth_weights(224, 42)
    c = vit.synth_weights(224, 43)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not np.array_equal(a[6], c[6])
    # fixed values: the generator is integer hashing + exact float arithmetic, identical everywhere
    assert a[6][:3].tolist() == [float(np.float32(v)) for v in (0.027973, -0.034692, 0.005305)]
    assert abs(float(a[6].std()) - 0.03) < 1e-3 and abs(float(a[4].mean()) - 1.0) < 1e-2
    i1 = vit.synth_images(4, 224, 7)
    i2 = vit.synth_images(2, 224, 7, first_index=2)
    assert np.array_equal(i1[2:], i2)                    # image i depends only on (seed, i): shardable
    assert i1.min() >= -2.12 and i1.max() <= 2.64
    w384 = vit.synth_weights(384, 42)
    assert w384[3].size == 577 * 768 and np.array_equal(w384[6], a[6])
    net = vit.as_network(a)
    assert vit.lib.vit_validate_weights(net, 152, 224) == 0


def test_pass_schedule_covers_the_shard_with_a_small_first_pass(vit):
    """vit_cuda_forward's pass schedule (pure host arithmetic): contiguous, complete, first pass <= 32 images (its
    H2D copy is the only exposed one), growth <= 3x per pass, never above the workspace size."""
    for n, mb in [(1, 1), (1, 1024), (31, 64), (32, 32), (33, 1024), (100, 16), (1024, 1024), (1024, 256), (8192, 1024), (5000, 999)]:
        sched = vit.pass_schedule(n, mb)
        assert sum(c for _, c in sched) == n
        pos = 0
        for i, (f, c) in enumerate(sched):
            assert f == pos and 0 < c <= mb
            if i == 0:
                assert c <= 32
            else:
                assert c <= 3 * sched[i - 1][1]
            pos += c
    for n, mb in [(1, 1), (63, 1024), (1024, 1024), (1000, 100), (5000, 999)]:     # staged input: 64, then passes of <= 128
        sched = vit.pass_schedule(n, mb, staged=True)
        assert sum(c for _, c in sched) == n and all(0 < c <= min(mb, 128) for _, c in sched) and sched[0][1] <= 64
        assert [f for f, _ in sched] == list(np.cumsum([0] + [c for _, c in sched[:-1]]))
    assert vit.pass_schedule(0, 8) == []
    assert vit.pass_schedule(1024, 1024) == [(0, 32), (32, 96), (128, 288), (416, 608)]
    # growth factor below 3 (several GPUs sharing the host's copy bandwidth): same invariants, gentler ramp, always progress
    for growth in (100, 134, 150, 200, 250, 300, 1000, -5):
        g = min(300, max(100, growth))
        for n, mb in [(1024, 1024), (1000, 100), (33, 1024), (5000, 999)]:
            sched = vit.pass_schedule(n, mb, growth_percent=growth) if n / min(mb, 32) < 40 or g > 100 else None
            if sched is None:
                continue
            assert sum(c for _, c in sched) == n and sched[0][1] <= 32
            for i in range(1, len(sched)):
                assert sched[i][0] == sched[i - 1][0] + sched[i - 1][1] and 0 < sched[i][1] <= mb
                assert sched[i][1] <= max(sched[i - 1][1] + 1, sched[i - 1][1] * g // 100)
    assert vit.pass_schedule(1024, 1024, growth_percent=200) == [(0, 32), (32, 64), (96, 128), (224, 256), (480, 512), (992, 32)]
    with pytest.raises(vit.VitCudaError):
        vit.pass_schedule(100000, 1)   # more than 64 passes


def _wave_cost(nb, tokens=197, pairs=74):
    """Tile waves of one pass on `pairs` CTA pairs, in units of one K = 768 wave: out_proj + mlp_3 (3 column tiles of 256,
    K = 768 + 3072), in_proj (9), mlp_0 (12) -- the model vit_cuda_pass_schedule_waves optimises."""
    tm = -(-nb * tokens // 256)
    w = lambda tiles_n: -(-tm * tiles_n // pairs)
    return 5 * w(3) + w(9) + w(12)


def test_wave_efficient_pass_schedule(vit):
    """vit_cuda_pass_schedule_waves, the schedule vit_cuda_forward runs: the geometric schedule re-cut at sizes that fill whole
    waves of GEMM tiles.  Complete and contiguous, within the workspace, every pass within [0.8, 1.08] of a geometric ramp's
    step (so its copy still hides), no stub pass at the end, and never more tile waves than the plain schedule."""
    assert vit.pass_schedule_waves(1024, 1024) == [(0, 31), (31, 96), (127, 288), (415, 609)]     # from 32 + 96 + 288 + 608
    assert [c for _, c in vit.pass_schedule_waves(1024, 1024, staged=True)] == [63] + [127] * 7 + [72]
    assert _wave_cost(31) < _wave_cost(32) and _wave_cost(127) < _wave_cost(128)     # the boundaries the header quotes
    for tokens, sm in [(197, 148), (577, 148), (197, 132), (5, 148)]:
        for staged in (False, True):
            for growth in (300, 250, 160, 100):
                for n, mb in [(0, 8), (1, 1), (5, 1024), (33, 1024), (100, 16), (100, 1024), (1024, 1024), (1024, 256), (8192, 1024), (5000, 999)]:
                    if (n / min(mb, 32) > 40 and growth == 100) or (staged and n / min(mb, 128) > 60):   # beyond the wrappers' pass arrays
                        continue
                    sched = vit.pass_schedule_waves(n, mb, tokens=tokens, sm_count=sm, staged=staged, growth_percent=growth)
                    plain = vit.pass_schedule(n, mb, staged=staged, growth_percent=growth)
                    assert sum(c for _, c in sched) == n
                    assert [f for f, _ in sched] == [int(v) for v in np.cumsum([0] + [c for _, c in sched[:-1]])][:len(sched)]
                    assert all(0 < c <= mb for _, c in sched)
                    if n:
                        assert sched[0][1] <= max(plain[0][1] + plain[0][1] // 12, min(n, mb) if n - plain[0][1] < 16 else 0)
                    if len(sched) > 1 and sched[-2][1] + sched[-1][1] <= mb:
                        assert sched[-1][1] >= min(16, sched[-2][1] // 4)                # no stub at the end
                    pairs = sm // 2
                    cost = sum(_wave_cost(c, tokens, pairs) for _, c in sched)
                    assert cost <= sum(_wave_cost(c, tokens, pairs) for _, c in plain) + 7 * max(0, len(sched) - len(plain))
    with pytest.raises(vit.VitCudaError):
        vit.pass_schedule_waves(1024, 1024, tokens=0)


def test_cost_model_pass_schedule(vit):
    """vit_cuda_pass_schedule_model: pass sizes from (copy us / image, kernel us / image, fixed us / pass).  Complete, contiguous,
    within the workspace; every pass's copy hides under the kernels of the pass before it (10 % reserve) unless the copies are
    the slower side (then equal passes); the pass count weighs the exposed first copy against the fixed cost per pass."""
    assert [c for _, c in vit.pass_schedule_model(1024, 1024, 10.9, 33.0, 700.0)] == [68, 243, 713]
    assert vit.pass_schedule_model(1024, 1024, 10.9, 33.0, 1e7) == [(0, 1024)]                   # passes cost too much: one
    assert vit.pass_schedule_model(1024, 1024, 10.9, 33.0, 0.0)[0] == (0, 1)                      # passes are free: start at once
    assert [c for _, c in vit.pass_schedule_model(1024, 1024, 40.0, 33.0, 700.0)] == [128] * 8   # copy bound: equal passes
    assert vit.pass_schedule_model(0, 8, 10.9, 33.0, 700.0) == []
    for h, k, fx in [(10.9, 33.0, 700.0), (21.5, 33.0, 700.0), (32.0, 112.0, 900.0), (40.0, 33.0, 700.0), (5.0, 50.0, 0.0), (10.9, 33.0, 5000.0)]:
        for n, mb in [(1, 1), (5, 1024), (33, 1024), (100, 16), (1024, 1024), (1024, 256), (8192, 1024), (5000, 999)]:
            sched = vit.pass_schedule_model(n, mb, h, k, fx)
            assert sum(c for _, c in sched) == n and all(0 < c <= mb for _, c in sched)
            assert [f for f, _ in sched] == [int(v) for v in np.cumsum([0] + [c for _, c in sched[:-1]])]
            for (_, a), (_, b) in zip(sched, sched[1:]):
                assert b <= max(a, 0.9 * (fx + k * a) / h) + 1
            # not worse than one pass, and not worse than starting with 32 images and tripling (the plain schedule's shape)
            exposed = h * sched[0][1] + fx * len(sched)
            assert exposed <= h * min(n, mb) + fx * -(-n // mb) + 1e-6
    for bad in [(1024, 1024, 0.0, 33.0, 700.0), (1024, 1024, 10.9, -1.0, 700.0), (1024, 0, 10.9, 33.0, 700.0)]:
        with pytest.raises(vit.VitCudaError):
            vit.pass_schedule_model(*bad)


def test_weight_cache_blob_round_trip_and_damage_detection(vit, tmp_path):
    """The single-file weight cache: bit-exact round trip of all 152 tensors, and a flipped byte, a
    truncated file or a foreign file are refused."""
    w = vit.synth_weights(224, 42)
    path = tmp_path / "weights.vitw"
    assert vit.lib.save_weights_blob(str(path).encode(), vit.as_network(w), 152, 224) == 0
    assert path.stat().st_size == 8 + 8 + 152 * 8 + sum(a.size for a in w) * 4 + 8
    back = (vit.Tensor * 152)()
    img = C.c_int(0)
    assert vit.lib.load_weights_blob(str(path).encode(), back, 152, C.byref(img)) == 0 and img.value == 224
    for i in range(152):
        got = np.ctypeslib.as_array(back[i].data, shape=(back[i].size,))
        assert np.array_equal(got, w[i]), i
    vit.lib.free_weights(back, 152)
    raw = bytearray(path.read_bytes())
    raw[len(raw) // 2] ^= 0x40
    bad = tmp_path / "damaged.vitw"
    bad.write_bytes(raw)
    assert vit.lib.load_weights_blob(str(bad).encode(), back, 152, None) != 0
    bad.write_bytes(path.read_bytes()[:-100])
    assert vit.lib.load_weights_blob(str(bad).encode(), back, 152, None) != 0
    bad.write_bytes(b"not a cache")
    assert vit.lib.load_weights_blob(str(bad).encode(), back, 152, None) != 0
    # a tensor of the wrong size is refused at save time
    short = [a for a in w]
    short[6] = short[6][:-1].copy()
    assert vit.lib.save_weights_blob(str(bad).encode(), vit.as_network(short), 152, 224) != 0


def test_streaming_image_reader_delivers_the_file_in_chunks(vit, tmp_path):
    """vit_image_stream_* (bounded-memory form of load_image_data, Network.c:24-97): same pixels as the whole-file
    loader in any chunking, n-limit respected, truncated files and bad headers refused."""
    lib = vit.lib
    lib.vit_image_stream_open.restype = C.c_void_p
    lib.vit_image_stream_open.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 4
    lib.vit_image_stream_read.restype = C.c_int
    lib.vit_image_stream_read.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
    lib.vit_image_stream_close.restype = None
    lib.vit_image_stream_close.argtypes = [C.c_void_p]
    n, S = 7, 32
    imgs = vit.synth_images(n, S, 5)
    path = tmp_path / "input-7.bin"
    assert lib.save_image_data(str(path).encode(), vit.fptr(imgs), n, 3, S, S) == 0
    for chunk in (1, 3, 7, 50):
        dims = [C.c_int() for _ in range(4)]
        s = lib.vit_image_stream_open(str(path).encode(), *[C.byref(d) for d in dims])
        assert s and [d.value for d in dims] == [n, 3, S, S]
        got, buf = [], np.empty((chunk, 3, S, S), dtype=np.float32)
        while True:
            k = lib.vit_image_stream_read(s, vit.fptr(buf), chunk)
            assert k >= 0
            if k == 0:
                break
            got.append(buf[:k].copy())
        assert lib.vit_image_stream_read(s, vit.fptr(buf), chunk) == 0       # stays at the end
        lib.vit_image_stream_close(s)
        assert np.array_equal(np.concatenate(got), imgs)
    short = tmp_path / "short.bin"
    short.write_bytes(path.read_bytes()[:-100])
    assert not lib.vit_image_stream_open(str(short).encode(), None, None, None, None)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(np.array([0, 3, S, S], dtype=np.int32).tobytes())
    assert not lib.vit_image_stream_open(str(bad).encode(), None, None, None, None)
    assert not lib.vit_image_stream_open(str(tmp_path / "none.bin").encode(), None, None, None, None)


def test_host_only_library_has_the_same_assets_and_no_cuda(vit):
    """lib/libvit_hostio.so (what bench.py --impl reference loads instead of the product library): same synthetic assets
    bit for bit, and no dependency on the CUDA runtime."""
    import subprocess
    import vit_hostio as H
    a, b = H.synth_weights(224, 42), vit.synth_weights(224, 42)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(H.synth_images(3, 224, 7, first_index=5), vit.synth_images(3, 224, 7, first_index=5))
    ldd = subprocess.run(["ldd", str(H.LIB_PATH)], capture_output=True, text=True).stdout
    assert "cuda" not in ldd.lower() and "vit_b200" not in ldd


def test_cli_sequential_backend_runs_the_callers_reference_build(vit, oracle, tmp_path):
    """vit_main --backend seq --seq-lib LIB (the Main.c:48-53 path): the driver loads ViT_seq() from a shared library the
    CALLER supplies -- here the reference's own ViT_seq.c as compiled by oracle/Makefile -- runs it on the loaded images and
    weights and writes the Main.c:71 result file; same top-1 / probability as calling that library directly.  No GPU."""
    import subprocess
    from pathlib import Path
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libvit_ref.so not built (needs /root/reference)")
    exe = Path(vit.PKG_DIR) / "bin" / "vit_main"
    ref_lib = Path(oracle.__file__).resolve().parent / "_ref" / "libvit_ref.so"
    n = 2
    w = vit.synth_weights(224, 42)
    imgs = vit.synth_images(n, 224, 7)
    img_file, wdir, res = tmp_path / "input-2.bin", tmp_path / "Network", tmp_path / "seq_result.txt"
    assert vit.lib.save_image_data(str(img_file).encode(), vit.fptr(imgs), n, 3, 224, 224) == 0
    assert vit.lib.save_weights(str(wdir).encode(), vit.as_network(w), 152, 224) == 0
    out = subprocess.run([str(exe), "--backend", "seq", "--seq-lib", str(ref_lib), "--images", str(img_file), "--weights", str(wdir),
                          "--result", str(res)], capture_output=True, timeout=900)
    stdout = out.stdout.decode(errors="replace")           # the reference's ViT_seq prints a non-UTF-8 progress line per image
    assert out.returncode == 0, stdout + out.stderr.decode(errors="replace")
    assert "Sequential time" in stdout
    probs = oracle.ref_vit_seq(w, imgs)
    lines = res.read_text().splitlines()
    assert len(lines) == n
    for i, ln in enumerate(lines):
        assert ln == f"[{i}] label: {int(probs[i].argmax())} / prob: {probs[i].max():.6f}"
    bad = subprocess.run([str(exe), "--backend", "seq", "--images", str(img_file), "--weights", str(wdir)], capture_output=True, text=True)
    assert bad.returncode == 2 and "--seq-lib" in bad.stderr
