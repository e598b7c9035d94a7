"""ctypes binding of lib/libvit_hostio.so -- the host-only part of the package (file formats of the reference's loader,
seeded synthetic assets; include/vit_host.h), without the CUDA engine.  For CPU-side tools: the reference arm of bench.py
uses this so that it never maps the product library libvit_b200.so."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libvit_hostio.so"
NUM_TENSORS = 152

if not LIB_PATH.exists():
    raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {PKG_DIR}`")
lib = C.CDLL(str(LIB_PATH))


class Tensor(C.Structure):  # == reference `Network` (Network.h:18-21)
    _fields_ = [("data", C.POINTER(C.c_float)), ("size", C.c_size_t)]


_f32p = C.POINTER(C.c_float)
lib.vit_tensor_numel.restype = C.c_size_t
lib.vit_tensor_numel.argtypes = [C.c_int, C.c_int]
lib.vit_synth_weights.restype = C.c_int
lib.vit_synth_weights.argtypes = [C.POINTER(Tensor), C.c_int, C.c_int, C.c_uint64]
lib.vit_synth_images.restype = None
lib.vit_synth_images.argtypes = [_f32p, C.c_int, C.c_int, C.c_uint64, C.c_int]
lib.free_weights.restype = None
lib.free_weights.argtypes = [C.POINTER(Tensor), C.c_int]
lib.load_weights.restype = C.c_int
lib.load_weights.argtypes = [C.c_char_p, C.POINTER(Tensor), C.c_int]


def synth_weights(img_size: int = 224, seed: int = 42) -> list[np.ndarray]:
    """152 fp32 arrays with the loader's 1e-6 rounding already applied (same bits as vit_b200.synth_weights)."""
    tmp = (Tensor * NUM_TENSORS)()
    if lib.vit_synth_weights(tmp, NUM_TENSORS, img_size, seed) != 0:
        raise MemoryError("vit_synth_weights failed")
    out = []
    for i in range(NUM_TENSORS):
        n = int(lib.vit_tensor_numel(i, img_size))
        a = np.empty(n, dtype=np.float32)
        C.memmove(a.ctypes.data, tmp[i].data, n * 4)
        out.append(a)
    lib.free_weights(tmp, NUM_TENSORS)
    return out


def synth_images(n: int, img_size: int = 224, seed: int = 7, first_index: int = 0) -> np.ndarray:
    out = np.empty((n, 3, img_size, img_size), dtype=np.float32)
    lib.vit_synth_images(out.ctypes.data_as(_f32p), n, img_size, seed, first_index)
    return out


def load_weights_dir(directory: str) -> list[np.ndarray | None]:
    """load_weights (Network.c:119-194 on POSIX, with its 1e-6 rounding): one array per slot, None where the file is missing."""
    tmp = (Tensor * NUM_TENSORS)()
    if lib.load_weights(str(directory).encode(), tmp, NUM_TENSORS) < 0:
        raise FileNotFoundError(directory)
    out: list[np.ndarray | None] = []
    for i in range(NUM_TENSORS):
        if not tmp[i].data:
            out.append(None)
            continue
        a = np.empty(tmp[i].size, dtype=np.float32)
        C.memmove(a.ctypes.data, tmp[i].data, tmp[i].size * 4)
        out.append(a)
    lib.free_weights(tmp, NUM_TENSORS)
    return out
