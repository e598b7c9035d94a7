import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


def round_operand(a: np.ndarray, precision: int) -> np.ndarray:
    """Round fp32 to the GEMM operand precision (0 = bf16, 1 = fp16), round-to-nearest-even,
    exactly as the device conversion does."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if precision == 1:
        return a.astype(np.float16).astype(np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32).reshape(a.shape)


@pytest.fixture(scope="session")
def vit():
    import vit_b200
    return vit_b200


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    return oracle_py


@pytest.fixture(scope="session")
def weights224(vit):
    return vit.synth_weights(224, 42)
