"""ctypes loader for the CPU oracle (oracle/liboracle.so) and, when present, the compiled
reference (oracle/_ref/libvit_ref.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by the
product package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
LIB = ORACLE_DIR / "liboracle.so"
REF_LIB = ORACLE_DIR / "_ref" / "libvit_ref.so"
_f32p = C.POINTER(C.c_float)


def build(quiet: bool = True):
    subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load():
    if not LIB.exists():
        build()
    lib = C.CDLL(str(LIB))
    lib.vit_oracle_forward.restype = C.c_int
    lib.vit_oracle_forward.argtypes = [C.POINTER(_f32p), C.c_int, _f32p, C.c_int, _f32p, _f32p, C.c_int]
    lib.oracle_linear.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p]
    lib.oracle_layer_norm.argtypes = [_f32p, _f32p, C.c_int, _f32p, _f32p]
    lib.oracle_gelu_inplace.argtypes = [_f32p, C.c_size_t]
    lib.oracle_attention_core.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int]
    lib.oracle_multihead_attn.argtypes = [_f32p, _f32p, C.c_int, _f32p, _f32p, _f32p, _f32p]
    lib.oracle_embed.argtypes = [_f32p, _f32p, C.c_int, _f32p, _f32p, _f32p, _f32p]
    lib.oracle_encoder_block.argtypes = [_f32p, _f32p, C.c_int, C.POINTER(_f32p)]
    lib.oracle_softmax.argtypes = [_f32p, _f32p, C.c_int]
    lib.oracle_round_weights.argtypes = [_f32p, C.c_size_t]
    lib.oracle_tensor_numel.restype = C.c_size_t
    lib.oracle_tensor_numel.argtypes = [C.c_int, C.c_int]
    lib.oracle_max_threads.restype = C.c_int
    return lib


lib = _load()


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_f32p)


def _wptrs(weights):
    arr = (_f32p * len(weights))()
    for i, w in enumerate(weights):
        arr[i] = _p(w)
    return arr


def max_threads() -> int:
    return int(lib.oracle_max_threads())


def forward(weights, images: np.ndarray, img_size: int = 224, n_threads: int = 0, want_probs: bool = False):
    n = images.shape[0]
    logits = np.empty((n, 1000), dtype=np.float32)
    probs = np.empty((n, 1000), dtype=np.float32) if want_probs else None
    rc = lib.vit_oracle_forward(_wptrs(weights), img_size, _p(images), n, _p(logits),
                                _p(probs) if want_probs else None, n_threads)
    if rc != 0:
        raise RuntimeError("vit_oracle_forward failed")
    return (logits, probs) if want_probs else logits


def linear(x, W, b):
    assert x.ndim == 2 and W.ndim == 2 and W.shape[1] == x.shape[1] and b.size == W.shape[0]
    y = np.empty((x.shape[0], W.shape[0]), dtype=np.float32)
    lib.oracle_linear(_p(x), _p(y), x.shape[0], x.shape[1], W.shape[0], _p(W), _p(b))
    return y


def layer_norm(x, w, b):
    y = np.empty_like(x)
    lib.oracle_layer_norm(_p(x), _p(y), x.shape[0], _p(w), _p(b))
    return y


def gelu(x):
    y = np.ascontiguousarray(x, dtype=np.float32).copy()
    lib.oracle_gelu_inplace(_p(y), y.size)
    return y


def attention_core(q, k, v):
    out = np.empty_like(q)
    lib.oracle_attention_core(_p(q), _p(k), _p(v), _p(out), q.shape[0])
    return out


def embed(image, cls, conv_w, conv_b, pos):
    s = image.shape[-1]
    t = (s // 16) ** 2 + 1
    out = np.empty((t, 768), dtype=np.float32)
    lib.oracle_embed(_p(image), _p(out), s, _p(cls), _p(conv_w), _p(conv_b), _p(pos))
    return out


def encoder_block(x, layer_weights):
    y = np.empty_like(x)
    lib.oracle_encoder_block(_p(x), _p(y), x.shape[0], _wptrs(layer_weights))
    return y


def softmax(logits):
    out = np.empty_like(logits)
    for i in range(logits.shape[0]):
        lib.oracle_softmax(_p(logits[i]), _p(out[i]), logits.shape[1])
    return out


# ----------------------------------------------------------------------------- compiled reference
class _RefNetwork(C.Structure):
    _fields_ = [("data", _f32p), ("size", C.c_size_t)]


class _RefImage(C.Structure):
    _fields_ = [("n", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int), ("data", _f32p)]


def ref_available() -> bool:
    return REF_LIB.exists()


def ref_lib():
    r = C.CDLL(str(REF_LIB))
    r.ViT_seq.argtypes = [C.POINTER(_RefImage), C.POINTER(_RefNetwork), C.POINTER(_f32p)]
    r.ViT_seq.restype = None
    r.load_weights.argtypes = [C.c_char_p, C.POINTER(_RefNetwork), C.c_int]
    r.load_image_data.argtypes = [C.c_char_p]
    r.load_image_data.restype = C.POINTER(_RefImage)
    r.comparator.restype = C.c_int
    return r


def ref_vit_seq(weights, images: np.ndarray) -> np.ndarray:
    """Run the reference's own ViT_seq() (224x224 only) -> softmax probabilities [n][1000]."""
    r = ref_lib()
    n = images.shape[0]
    net = (_RefNetwork * 152)()
    for i, w in enumerate(weights):
        net[i].data = _p(w)
        net[i].size = w.size
    imgs = (_RefImage * n)()
    for i in range(n):
        imgs[i].n, imgs[i].c, imgs[i].h, imgs[i].w = n, 3, 224, 224
        imgs[i].data = _p(images[i])
    probs = np.zeros((n, 1000), dtype=np.float32)
    rows = (_f32p * n)()
    for i in range(n):
        rows[i] = _p(probs[i])
    r.ViT_seq(imgs, net, rows)
    return probs
