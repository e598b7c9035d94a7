"""The C-ABI shared library loads and exports every function include/*.h declares; without a
GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import re
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
DECL = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?[\s\*]+(\w+)\s*\(", re.M)


def _declared(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    text = re.sub(r"typedef\s+struct[^;]*?\{.*?\}[^;]*;", "", text, flags=re.S)
    text = re.sub(r"enum\s*\{.*?\};", "", text, flags=re.S)
    names = set()
    for stmt in text.split(";"):
        m = DECL.search(stmt + "(") if "(" in stmt else None
        if m and m.group(1) not in ("defined",):
            names.add(m.group(1))
    return names


def test_every_declared_function_is_exported(vit):
    declared = _declared(ROOT / "include" / "vit_cuda.h") | _declared(ROOT / "include" / "vit_host.h")
    assert {"vit_cuda_init", "vit_cuda_forward", "vit_cuda_free", "ViT_cuda", "load_weights", "load_image_data",
            "comparator_files", "vit_cuda_op_linear", "vit_cuda_op_attention"} <= declared
    missing = [n for n in sorted(declared) if not hasattr(vit.lib, n)]
    assert not missing, f"declared in include/*.h but not exported by libvit_b200.so: {missing}"


def test_library_contains_sm100a_tcgen05_code():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    lib = ROOT / "vision-transformer-opencl_b200" / "lib" / "libvit_b200.so"
    elf = subprocess.run([cuobjdump, "-lelf", str(lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", str(lib)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM"):  # tcgen05.mma, TMA load/store, tcgen05.ld/st
        assert mnemonic in sass, mnemonic


def test_shard_ranges(vit):
    for n, g in [(8192, 8), (100, 8), (7, 4), (1, 8), (0, 2)]:
        spans = [vit.shard_range(n, g, i) for i in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) == -(-n // g)
    with pytest.raises(vit.VitCudaError):
        vit.shard_range(4, 2, 2)


def _have_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU ") > 0
    except OSError:
        return False


@pytest.mark.skipif(_have_gpu(), reason="this test documents the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(vit, weights224):
    with pytest.raises(vit.VitCudaError) as ei:
        vit.Engine(weights224, 224, max_batch=1)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    x = np.zeros((4, 64), dtype=np.float32)
    with pytest.raises(vit.VitCudaError):
        vit.op_linear(x, np.zeros((256, 64), dtype=np.float32), np.zeros(256, dtype=np.float32))
    assert vit.lib.initialize_cuda() == 0  # device selection is deferred; ViT_cuda reports the failure
