"""ctypes binding of lib/libvit_b200.so -- the C ABI declared in include/vit_cuda.h and
include/vit_host.h.  No torch, no numpy-side compute: arrays only carry bytes to and from the
C library.  Importing fails loudly if the library has not been built; every compute call
raises VitCudaError if the CUDA engine cannot run (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libvit_b200.so"

NUM_TENSORS = 152
NUM_CLASSES = 1000
PREC_BF16, PREC_FP16, PREC_AUTO = 0, 1, 2
PREC_NAMES = {0: "bf16", 1: "fp16", 2: "auto"}
OPT_ATTENTION_EXACT, OPT_CLASS_ROW_PRUNING, OPT_LN_FUSED, OPT_PDL, OPT_GRAPHS, OPT_HOST_THREADS, OPT_RESIDUAL16, OPT_WAVE_PASSES = range(8)
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2
PROF_CATEGORIES = ["class_rows", "embed_gemm", "layernorm", "qkv_gemm", "attention", "out_gemm", "fc1_gemm", "fc2_gemm", "head"]


class VitCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vit_cuda error {code}: {msg}")
        self.code = code


class Tensor(C.Structure):  # == reference `Network` (Network.h:18-21)
    _fields_ = [("data", C.POINTER(C.c_float)), ("size", C.c_size_t)]


class ImageData(C.Structure):  # reference `ImageData` (Network.h:7-13)
    _fields_ = [("n", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int), ("data", C.POINTER(C.c_float))]


if not LIB_PATH.exists():
    raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {PKG_DIR}` "
                      f"(or python -c 'import __graft_entry__ as g; g.build()')")
lib = C.CDLL(str(LIB_PATH))

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


def _sig(name, res, *args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = list(args)
    return fn


# ---- include/vit_cuda.h
_sig("vit_cuda_init", C.c_int, C.POINTER(Tensor), C.c_int, C.c_int, C.c_int, C.c_int)
_sig("vit_cuda_init_ex", C.c_int, C.POINTER(Tensor), C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int)
_sig("vit_cuda_forward", C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)
_sig("vit_cuda_shard_range", C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p)
_sig("vit_cuda_forward_scattered", C.c_int, C.POINTER(_f32p), C.c_int, C.c_void_p, C.c_void_p)
_sig("vit_cuda_pass_schedule", C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int)
_sig("vit_cuda_pass_schedule_ex", C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int)
_sig("vit_cuda_pass_schedule_growth", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int)
_sig("vit_cuda_pass_schedule_model", C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _i32p, _i32p, C.c_int)
_sig("vit_cuda_pass_schedule_waves", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int)
_sig("vit_cuda_forward_device", C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p)
_sig("vit_cuda_enqueue_device", C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p)
_sig("vit_cuda_sync", C.c_int, C.c_int)
_sig("vit_cuda_stream", C.c_void_p, C.c_int)
_sig("vit_cuda_free", None)
_sig("vit_cuda_last_error", C.c_char_p)
_sig("vit_cuda_launch_count", C.c_longlong)
_sig("vit_cuda_info", C.c_int, C.POINTER(C.c_longlong), C.c_int)
_sig("vit_cuda_set_option", C.c_int, C.c_int, C.c_int)
_sig("vit_cuda_get_option", C.c_int, C.c_int, _i32p)
_sig("vit_cuda_save_weight_cache", C.c_int, C.c_char_p)
_sig("vit_cuda_init_from_cache", C.c_int, C.c_char_p, C.c_int, C.c_int, _i32p)
_sig("vit_cuda_set_attention_exact", C.c_int, C.c_int)
_sig("vit_cuda_set_class_row_pruning", C.c_int, C.c_int)
_sig("vit_cuda_timer_start", C.c_int, C.c_int)
_sig("vit_cuda_timer_stop", C.c_int, C.c_int, _f32p)
_sig("vit_cuda_profile_enable", C.c_int, C.c_int)
_sig("vit_cuda_profile_read", C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int)
_sig("vit_cuda_dev_alloc", C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p))
_sig("vit_cuda_dev_free", C.c_int, C.c_int, C.c_void_p)
_sig("vit_cuda_dev_upload", C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
_sig("vit_cuda_dev_download", C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
_sig("vit_cuda_host_alloc_pinned", C.c_int, C.c_size_t, C.POINTER(C.c_void_p))
_sig("vit_cuda_host_free_pinned", C.c_int, C.c_void_p)
_sig("vit_cuda_op_linear", C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_sig("vit_cuda_op_layernorm", C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int)
_sig("vit_cuda_op_ln_linear", C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int)
_sig("vit_cuda_op_linear_residual_stats", C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int)
_sig("vit_cuda_op_attention", C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int)
_sig("vit_cuda_debug_attention_trace", C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_int)
_sig("vit_cuda_op_embed", C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int)
_sig("vit_cuda_op_encoder_block", C.c_int, _f32p, _f32p, C.c_int, C.c_int)
_sig("vit_cuda_op_head", C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int)
# ---- include/vit_host.h
_sig("load_image_data", C.POINTER(ImageData), C.c_char_p)
_sig("free_image_data", None, C.POINTER(ImageData))
_sig("load_weights", C.c_int, C.c_char_p, C.POINTER(Tensor), C.c_int)
_sig("free_weights", None, C.POINTER(Tensor), C.c_int)
_sig("vit_tensor_numel", C.c_size_t, C.c_int, C.c_int)
_sig("vit_tensor_name", C.c_char_p, C.c_int, C.c_char_p, C.c_size_t)
_sig("vit_validate_weights", C.c_int, C.POINTER(Tensor), C.c_int, C.c_int)
_sig("save_image_data", C.c_int, C.c_char_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int)
_sig("save_weights", C.c_int, C.c_char_p, C.POINTER(Tensor), C.c_int, C.c_int)
_sig("save_weights_blob", C.c_int, C.c_char_p, C.POINTER(Tensor), C.c_int, C.c_int)
_sig("load_weights_blob", C.c_int, C.c_char_p, C.POINTER(Tensor), C.c_int, _i32p)
_sig("initialize_cuda", C.c_int)
_sig("ViT_cuda", None, C.POINTER(ImageData), C.POINTER(Tensor), C.POINTER(_f32p))
_sig("ViT_cuda_status", C.c_int)
_sig("Release_cuda", None)
_sig("vit_softmax", None, _f32p, _f32p, C.c_int)
_sig("vit_argmax", C.c_int, _f32p, C.c_int)
_sig("write_results", C.c_int, C.c_char_p, C.POINTER(_f32p), C.c_int)
_sig("comparator_files", C.c_int, C.c_char_p, C.c_char_p, C.c_int)
_sig("vit_synth_fill", None, _f32p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_float)
_sig("vit_synth_weights", C.c_int, C.POINTER(Tensor), C.c_int, C.c_int, C.c_uint64)
_sig("vit_synth_images", None, _f32p, C.c_int, C.c_int, C.c_uint64, C.c_int)


def _check(rc: int):
    if rc != 0:
        raise VitCudaError(rc, lib.vit_cuda_last_error().decode(errors="replace"))


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def tokens_for(img_size: int) -> int:
    return (img_size // 16) ** 2 + 1


# ---------------------------------------------------------------------------- assets
def tensor_numel(idx: int, img_size: int = 224) -> int:
    return int(lib.vit_tensor_numel(idx, img_size))


def tensor_name(idx: int) -> str:
    buf = C.create_string_buffer(128)
    return lib.vit_tensor_name(idx, buf, 128).decode()


def synth_weights(img_size: int = 224, seed: int = 42) -> list[np.ndarray]:
    """152 fp32 arrays with the loader's 1e-6 rounding already applied."""
    out = []
    sizes = [tensor_numel(i, img_size) for i in range(NUM_TENSORS)]
    arr = (Tensor * NUM_TENSORS)()
    bufs = [np.empty(n, dtype=np.float32) for n in sizes]
    # generate through the C routine tensor by tensor so numpy owns the memory
    tmp = (Tensor * NUM_TENSORS)()
    rc = lib.vit_synth_weights(tmp, NUM_TENSORS, img_size, seed)
    if rc != 0:
        raise MemoryError("vit_synth_weights failed")
    for i in range(NUM_TENSORS):
        C.memmove(bufs[i].ctypes.data, tmp[i].data, sizes[i] * 4)
        out.append(bufs[i])
    lib.free_weights(tmp, NUM_TENSORS)
    del arr
    return out


def synth_images(n: int, img_size: int = 224, seed: int = 7, first_index: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.empty((n, 3, img_size, img_size), dtype=np.float32)
    lib.vit_synth_images(fptr(out), n, img_size, seed, first_index)
    return out


def as_network(weights: list[np.ndarray]):
    """ctypes `Network[152]` viewing the numpy arrays (which must stay alive)."""
    arr = (Tensor * NUM_TENSORS)()
    for i, w in enumerate(weights):
        assert w.dtype == np.float32 and w.flags["C_CONTIGUOUS"]
        arr[i].data = w.ctypes.data_as(_f32p)
        arr[i].size = w.size
    return arr


# ---------------------------------------------------------------------------- engine
class Engine:
    """vit_cuda_init / vit_cuda_forward / vit_cuda_free as a context manager."""

    def __init__(self, weights: list[np.ndarray] | None, img_size: int = 224, max_batch: int = 64, n_gpus: int = 1,
                 device_ids: list[int] | None = None, precision: int = PREC_AUTO, cache: str | None = None):
        """weights: the 152 fp32 tensors (vit_cuda_init_ex), or None with cache = path of a file written by
        save_weight_cache (vit_cuda_init_from_cache; image size and precision policy come from the file)."""
        self.img_size, self.max_batch, self.n_gpus = img_size, max_batch, n_gpus
        ids = None
        if device_ids is not None:
            ids = (C.c_int * n_gpus)(*device_ids)
        if weights is None:
            _check(lib.vit_cuda_init_from_cache(os.fsencode(cache), max_batch, n_gpus, ids))
        else:
            net = as_network(weights)
            _check(lib.vit_cuda_init_ex(net, NUM_TENSORS, img_size, max_batch, n_gpus, ids, precision))
        self._up = True

    def save_weight_cache(self, path: str):
        _check(lib.vit_cuda_save_weight_cache(os.fsencode(path)))

    def set_option(self, option: int, value: int):
        _check(lib.vit_cuda_set_option(option, value))

    def get_option(self, option: int) -> int:
        v = C.c_int()
        _check(lib.vit_cuda_get_option(option, C.byref(v)))
        return v.value

    def forward(self, images: np.ndarray, want_top1: bool = False):
        n = images.shape[0]
        assert images.dtype == np.float32 and images.flags["C_CONTIGUOUS"]
        logits = np.empty((n, NUM_CLASSES), dtype=np.float32)
        top1 = np.empty(n, dtype=np.int32) if want_top1 else None
        _check(lib.vit_cuda_forward(images.ctypes.data, n, logits.ctypes.data, top1.ctypes.data if want_top1 else None))
        return (logits, top1) if want_top1 else logits

    def forward_scattered(self, images: list[np.ndarray], want_top1: bool = False):
        """vit_cuda_forward_scattered: one separately allocated [3][S][S] array per image (the reference loader's form)."""
        n = len(images)
        ptrs = (_f32p * n)(*[fptr(a) for a in images])
        logits = np.empty((n, NUM_CLASSES), dtype=np.float32)
        top1 = np.empty(n, dtype=np.int32)
        _check(lib.vit_cuda_forward_scattered(ptrs, n, logits.ctypes.data, top1.ctypes.data if want_top1 else None))
        return (logits, top1) if want_top1 else logits

    def forward_raw(self, images_ptr: int, n: int, logits_ptr: int):
        _check(lib.vit_cuda_forward(images_ptr, n, logits_ptr, None))

    # ---- device-resident path + timing (bench.py)
    def enqueue_device(self, d_images: int, n: int, d_logits: int, slot: int = 0):
        _check(lib.vit_cuda_enqueue_device(slot, d_images, n, d_logits))

    def sync(self, slot: int = 0):
        _check(lib.vit_cuda_sync(slot))

    def timer_start(self, slot: int = 0):
        _check(lib.vit_cuda_timer_start(slot))

    def timer_stop(self, slot: int = 0) -> float:
        ms = C.c_float()
        _check(lib.vit_cuda_timer_stop(slot, C.byref(ms)))
        return float(ms.value)

    def profile_enable(self, on: bool):
        _check(lib.vit_cuda_profile_enable(1 if on else 0))

    def profile_read(self, slot: int = 0) -> dict:
        ms = (C.c_double * len(PROF_CATEGORIES))()
        cnt = (C.c_longlong * len(PROF_CATEGORIES))()
        _check(lib.vit_cuda_profile_read(slot, ms, cnt, len(PROF_CATEGORIES)))
        return {k: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, k in enumerate(PROF_CATEGORIES)}

    def info(self) -> dict:
        v = (C.c_longlong * 18)()
        _check(lib.vit_cuda_info(v, 18))
        keys = ["sm_count", "cc_major", "cc_minor", "max_batch", "tokens", "precision", "n_gpus", "workspace_mib",
                "attention_exact", "attention_fallbacks", "class_row_pruning", "precision_policy", "precision_fallbacks", "weights_mib",
                "pass_growth_percent", "h2d_mb_per_s", "pass_fixed_us", "pass_ns_per_image"]
        d = dict(zip(keys, [int(x) for x in v]))
        d["precision"] = PREC_NAMES[d["precision"]]            # the operand type the next pass runs in
        d["precision_policy"] = PREC_NAMES[d["precision_policy"]]
        return d

    def set_class_row_pruning(self, on: bool):
        """Last layer: everything behind the attention for the class rows only (default on)."""
        _check(lib.vit_cuda_set_class_row_pruning(1 if on else 0))

    def set_attention_exact(self, on: bool):
        """Two-pass (exact row maximum) softmax instead of the default single-pass one."""
        _check(lib.vit_cuda_set_attention_exact(1 if on else 0))

    def close(self):
        if self._up:
            lib.vit_cuda_free()
            self._up = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def shard_range(n: int, n_gpus: int, g: int) -> tuple[int, int]:
    lo, hi = C.c_int(), C.c_int()
    _check(lib.vit_cuda_shard_range(n, n_gpus, g, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def pass_schedule(n_images: int, max_batch: int, staged: bool = False, growth_percent: int = 300) -> list[tuple[int, int]]:
    first, count = (C.c_int * 64)(), (C.c_int * 64)()
    n = lib.vit_cuda_pass_schedule_growth(n_images, max_batch, 1 if staged else 0, growth_percent, first, count, 64)
    if n < 0:
        _check(n)
    return [(first[i], count[i]) for i in range(n)]


def pass_schedule_waves(n_images: int, max_batch: int, tokens: int = 197, sm_count: int = 148, staged: bool = False,
                        growth_percent: int = 300) -> list[tuple[int, int]]:
    """The wave-efficient schedule vit_cuda_forward runs (vit_cuda_pass_schedule_waves)."""
    first, count = (C.c_int * 256)(), (C.c_int * 256)()
    n = lib.vit_cuda_pass_schedule_waves(n_images, max_batch, 1 if staged else 0, growth_percent, tokens, sm_count, first, count, 256)
    if n < 0:
        _check(n)
    return [(first[i], count[i]) for i in range(n)]


def pass_schedule_model(n_images: int, max_batch: int, copy_us: float, kernel_us: float, fixed_us: float) -> list[tuple[int, int]]:
    """vit_cuda_pass_schedule_model: pass sizes from the pipeline's cost model (before the wave-efficient re-cut)."""
    first, count = (C.c_int * 256)(), (C.c_int * 256)()
    n = lib.vit_cuda_pass_schedule_model(n_images, max_batch, copy_us, kernel_us, fixed_us, first, count, 256)
    if n < 0:
        _check(n)
    return [(first[i], count[i]) for i in range(n)]


def dev_alloc(slot: int, nbytes: int) -> int:
    p = C.c_void_p()
    _check(lib.vit_cuda_dev_alloc(slot, nbytes, C.byref(p)))
    return p.value


def dev_free(slot: int, ptr: int):
    _check(lib.vit_cuda_dev_free(slot, ptr))


def dev_upload(slot: int, dptr: int, a: np.ndarray):
    _check(lib.vit_cuda_dev_upload(slot, dptr, a.ctypes.data, a.nbytes))


def dev_download(slot: int, a: np.ndarray, dptr: int):
    _check(lib.vit_cuda_dev_download(slot, a.ctypes.data, dptr, a.nbytes))


def pinned_empty(shape, dtype=np.float32) -> tuple[np.ndarray, int]:
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _check(lib.vit_cuda_host_alloc_pinned(nbytes, C.byref(p)))
    buf = (C.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p.value


def pinned_free(ptr: int):
    lib.vit_cuda_host_free_pinned(ptr)


# ---------------------------------------------------------------------------- single operators
def op_linear(x, W, b, residual=None, epilogue=EPI_BIAS, precision=PREC_BF16):
    m, k = x.shape
    n = W.shape[0]
    y = np.empty((m, n), dtype=np.float32)
    _check(lib.vit_cuda_op_linear(fptr(x), fptr(W), fptr(b), fptr(residual) if residual is not None else None,
                                  fptr(y), m, n, k, epilogue, precision))
    return y


def op_ln_linear(x, ln_w, ln_b, W, b, epilogue=EPI_BIAS, precision=PREC_BF16):
    """epilogue(LN(x) W^T + b) with the LayerNorm folded into the GEMM (in_proj / mlp_0 of the forward pass)."""
    m, n = x.shape[0], W.shape[0]
    y = np.empty((m, n), dtype=np.float32)
    _check(lib.vit_cuda_op_ln_linear(fptr(x), fptr(ln_w), fptr(ln_b), fptr(W), fptr(b), fptr(y), m, n, epilogue, precision))
    return y


def op_linear_residual_stats(x, W, b, residual, precision=PREC_BF16):
    """residual + x W^T + b (fp32) plus its operand-precision copy and per-row (sum, sum of squares)."""
    m, k = x.shape
    y, yc = np.empty((m, 768), dtype=np.float32), np.empty((m, 768), dtype=np.float32)
    s1, s2 = np.empty(m, dtype=np.float32), np.empty(m, dtype=np.float32)
    _check(lib.vit_cuda_op_linear_residual_stats(fptr(x), fptr(W), fptr(b), fptr(residual), fptr(y), fptr(yc), fptr(s1), fptr(s2), m, k, precision))
    return y, yc, s1, s2


def op_layernorm(x, w, b, precision=PREC_BF16):
    y = np.empty_like(x)
    _check(lib.vit_cuda_op_layernorm(fptr(x), fptr(w), fptr(b), fptr(y), x.shape[0], precision))
    return y


def op_attention(qkv, batch, tokens, precision=PREC_BF16):
    out = np.empty((batch * tokens, 768), dtype=np.float32)
    _check(lib.vit_cuda_op_attention(fptr(qkv), fptr(out), batch, tokens, precision))
    return out


def attention_trace(qkv, batch, tokens, precision=PREC_BF16):
    """SM-clock timestamps [warp 19][item 16][event 8] of CTA 0 of the attention kernel (debug)."""
    tr = np.zeros((19, 16, 8), dtype=np.uint64)
    _check(lib.vit_cuda_debug_attention_trace(fptr(qkv), batch, tokens, precision, tr.ctypes.data_as(C.POINTER(C.c_uint64)), tr.size))
    return tr


def op_embed(images, cls, conv_w, conv_b, pos, precision=PREC_BF16, want_cast=False):
    """conv_proj (tf32, straight from the fp32 image) + class token + position embedding; want_cast: also the
    operand-precision copy of the rows that the kernel emits for the first LayerNorm-folded GEMM."""
    batch, _, s, _ = images.shape
    out = np.empty((batch * tokens_for(s), 768), dtype=np.float32)
    cast = np.empty_like(out) if want_cast else None
    _check(lib.vit_cuda_op_embed(fptr(images), fptr(cls), fptr(conv_w), fptr(conv_b), fptr(pos), fptr(out),
                                 fptr(cast) if want_cast else None, batch, s, precision))
    return (out, cast) if want_cast else out


def op_encoder_block(x, batch, layer):
    """One encoder block of the initialised engine (vit_cuda_op_encoder_block): x [batch*tokens][768] fp32 -> same shape."""
    y = np.empty_like(x)
    _check(lib.vit_cuda_op_encoder_block(fptr(x), fptr(y), batch, layer))
    return y


def op_head(x, ln_w, ln_b, head_w, head_b, batch, tokens):
    logits = np.empty((batch, NUM_CLASSES), dtype=np.float32)
    _check(lib.vit_cuda_op_head(fptr(x), fptr(ln_w), fptr(ln_b), fptr(head_w), fptr(head_b), fptr(logits), batch, tokens))
    return logits


def softmax_rows(logits: np.ndarray) -> np.ndarray:
    out = np.empty_like(logits)
    for i in range(logits.shape[0]):
        lib.vit_softmax(fptr(logits[i]), fptr(out[i]), logits.shape[1])
    return out


def launch_count() -> int:
    return int(lib.vit_cuda_launch_count())
