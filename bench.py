#!/usr/bin/env python
"""bench.py -- ViT-B/16 224x224 images/sec on B200 (BASELINE.json metric), device-resident and
end-to-end through the C ABI, with the dominant kernel's roofline and the CPU reference beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision auto|fp16|bf16]
    python bench.py --impl reference ...      # the reference's own ViT_seq on the host cores

A "step" is one forward pass of the hot path over one batch of B synthetic images per GPU
(default 1024 = BASELINE.json configs[2]; for N > 1 this is configs[3], 8192 images over 8 GPUs,
weights replicated, no collective on the data path).  Under torchrun (N > 1) every rank owns one
GPU; torch.distributed is used only for the barrier and the max-over-ranks reduction.
"""
import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
faulthandler.enable()      # a native crash prints the Python stack to stderr instead of dying silently

import numpy as np

PARITY_IMAGES = 64     # images of the timed batch checked against the oracle in the same run (bench line key "parity")
FLOP_PER_IMAGE_224 = 35_127_656_448  # matmul-only, 2 FLOP/MAC, un-padded 197 tokens
FLOP_PER_IMAGE = {224: FLOP_PER_IMAGE_224, 384: 110_968_700_928}  # 384: BASELINE.json configs[4] (577 tokens)
# per-launch algorithmic FLOPs of one GEMM over `rows` token rows
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one launch at B = 1024, FP16 operands, from the ncu --set full
# capture committed under profiles/ (r2_ncu_layer_res16.txt: the default configuration, patch rows' residual stream in FP16;
# r2_ncu_layer.txt holds the fp32-stream capture: out_proj 1.813 GB, mlp_3 3.055 GB)
NCU_TRAFFIC_BYTES = {"qkv_gemm": 1_197_473_000, "out_gemm": 903_505_000, "fc1_gemm": 1_508_485_000, "fc2_gemm": 1_882_722_000}
GEMM_FLOP_PER_ROW = {"qkv_gemm": 2 * 768 * 2304, "out_gemm": 2 * 768 * 768, "fc1_gemm": 2 * 768 * 3072, "fc2_gemm": 2 * 3072 * 768}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load = those drawing at least 70 % of the highest power seen (the sampler also catches the idle
        # moments before and after the timed region)
        pmax = max(power)
        load = [c for c, w in zip(sm, power) if w >= 0.7 * pmax] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "power_w_max": pmax, "samples": len(sm), "samples_under_load": len(load),
                "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref = its ViT_seq.c compiled unmodified) on
    the host cores: one image per thread, all threads busy, each step a bounded sample."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py as O
    import vit_hostio as H      # host-only library: this arm never maps the product's libvit_b200.so
    cores = os.cpu_count() or 1
    if args.cpu_threads:
        cores = args.cpu_threads
    w = H.synth_weights(224, 42)
    kind = "reference" if O.ref_available() else "port"
    n = cores  # images per step, one per thread
    imgs = H.synth_images(n, 224, 7)

    def step():
        if kind == "reference":
            ths = [threading.Thread(target=O.ref_vit_seq, args=(w, imgs[i:i + 1])) for i in range(n)]
            [t.start() for t in ths]
            [t.join() for t in ths]
        else:
            O.forward(w, imgs, 224, n_threads=cores)

    devnull = os.open(os.devnull, os.O_WRONLY)  # ViT_seq prints a timing line per image
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1)
    value = n * args.steps / dt
    sample = f"{n} images per step ({'reference ViT_seq(), one image per thread' if kind == 'reference' else 'oracle port, OpenMP'}), same synthetic batch (seed 7) truncated"
    print(json.dumps({
        "impl": "reference", "metric": "ViT-B/16 224x224 inference throughput", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ViT-B/16 224x224 synthetic batch, CPU ViT_seq", "images_per_step": n},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(gpu_logits=None):
    """The reference's ViT_seq on the host cores (one image per thread), timed; and -- BASELINE.md section 4 -- the GPU's
    answers for the same images checked against it in the same run (top-1, reference comparator rule on the top-1
    probability, comparator.c:64-74)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py as O
    import vit_b200 as V
    cores = os.cpu_count() or 1
    w = V.synth_weights(224, 42)
    imgs = V.synth_images(cores, 224, 7)
    kind = "reference" if O.ref_available() else "port"
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    probs = np.zeros((cores, 1000), dtype=np.float32)
    t0 = time.perf_counter()
    try:
        if kind == "reference":
            def one(i):
                probs[i] = O.ref_vit_seq(w, imgs[i:i + 1])[0]
            ths = [threading.Thread(target=one, args=(i,)) for i in range(cores)]
            [t.start() for t in ths]
            [t.join() for t in ths]
        else:
            probs = O.softmax(O.forward(w, imgs, 224, n_threads=cores))
    finally:
        os.dup2(saved, 1)
    dt = time.perf_counter() - t0
    out = {"value": cores / dt, "unit": "images/s", "cores": cores, "kind": kind,
           "sample": f"{cores} images of the same synthetic batch (seed 7), one image per host thread through "
                     f"{'the reference ViT_seq() compiled from its own source (oracle/_ref)' if kind == 'reference' else 'the oracle port'}, {dt:.1f} s"}
    if gpu_logits is not None and len(gpu_logits) >= cores:
        g = O.softmax(np.ascontiguousarray(gpu_logits[:cores]))
        top_c, top_g = probs.argmax(1), g.argmax(1)
        srt = np.sort(probs, 1)
        decisive = srt[:, -1] - srt[:, -2] > 0.06 * srt[:, -1]          # beyond what the stated logit tolerance can flip
        out["gpu_vs_cpu"] = {"images": int(cores), "top1_equal": int((top_c == top_g).sum()), "decisive_images": int(decisive.sum()),
                             "top1_equal_on_decisive": bool((top_c[decisive] == top_g[decisive]).all()),
                             "max_abs_dprob_top1": float(np.abs(g[np.arange(cores), top_c] - probs[np.arange(cores), top_c]).max()),
                             "comparator_rule_0.01": bool((np.abs(g[np.arange(cores), top_c] - probs[np.arange(cores), top_c]) <= 0.01).all())}
    return out


_ORACLE_LOGITS = {}


def parity_block(gpu_logits, n_images, weights, images, dtype_name):
    """Same-run parity of the benchmarked precision: the first n_images of the timed batch through the oracle port
    (bit-identical to the reference's ViT_seq, tests/test_oracle_vs_reference.py) on all host cores, against the logits
    the timed engine produced for them -- the stated tolerance 2e-2 + 1e-2 |ref| on every logit, top-1 on every image."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py as O
    t0 = time.perf_counter()
    if n_images not in _ORACLE_LOGITS:    # (a variant measured later in the same run is checked against the same answers)
        _ORACLE_LOGITS[n_images] = O.forward(weights, np.ascontiguousarray(images[:n_images]), 224, n_threads=os.cpu_count() or 1)
    ref = _ORACLE_LOGITS[n_images]
    dt = time.perf_counter() - t0
    got = np.ascontiguousarray(gpu_logits[:n_images])
    err = np.abs(got - ref)
    tol = 2e-2 + 1e-2 * np.abs(ref)
    srt = np.sort(ref, 1)
    margin = srt[:, -1] - srt[:, -2]
    top_ref, top_gpu = ref.argmax(1), got.argmax(1)
    return {"images": int(n_images), "dtype": dtype_name, "tolerance": "|dlogit| <= 2e-2 + 1e-2*|ref| on every logit; top-1 identical",
            "max_abs_dlogit": float(err.max()), "mean_abs_dlogit": float(err.mean()), "logits_outside_tolerance": int((err > tol).sum()),
            "logits_checked": int(err.size), "top1_equal": int((top_ref == top_gpu).sum()), "top1_all_equal": bool((top_ref == top_gpu).all()),
            "oracle_min_top2_margin": float(margin.min()), "logit_std": float(ref.std()), "passes": bool((err <= tol).all() and (top_ref == top_gpu).all()),
            "oracle": "oracle/vit_oracle.c (C restatement of ViT_seq.c, OpenMP over images)", "oracle_seconds": dt}


def abi_inproc(V, weights, S, B, G, steps, prec, per_process_e2e):
    """vit_cuda_init(n_gpus = G) in ONE process: G x B pinned images through vit_cuda_forward (the ABI shards them
    contiguously, one feeding host thread per GPU), then the same images as G x B separate allocations through
    vit_cuda_forward_scattered (the form ViT_cuda() receives: every pass is gathered into pinned staging by host
    threads, so this leg is bound by host memory bandwidth), wall clock, host buffers in and out."""
    n = G * B
    res = {"n_gpus": G, "images_per_call": n}
    with V.Engine(weights, S, max_batch=B, n_gpus=G, device_ids=list(range(G)), precision=prec) as eng:
        eng.set_class_row_pruning(False)
        h_imgs, ip = V.pinned_empty((n, 3, S, S))
        h_log, lp = V.pinned_empty((n, 1000))
        for g in range(G):
            V.synth_images(B, S, 7, first_index=g * B, out=h_imgs[g * B:(g + 1) * B])
        for _ in range(4):       # the pass schedule adapts to the copy rate it measures in its first calls
            eng.forward_raw(ip, n, lp)
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.forward_raw(ip, n, lp)
        dt = time.perf_counter() - t0
        info = eng.info()
        res["pinned"] = {"value": n * steps / dt, "unit": "images/s", "ms_per_call": dt / steps * 1e3,
                         "vs_one_process_per_gpu_e2e": n * steps / dt / per_process_e2e,
                         "pass_growth_percent_slot0": info["pass_growth_percent"], "h2d_mb_per_s_slot0": info["h2d_mb_per_s"],
                         "pass_fixed_us_slot0": info["pass_fixed_us"], "pass_ns_per_image_slot0": info["pass_ns_per_image"]}
        checksum = int(h_log.argmax(1).sum())
        eng.set_option(V.OPT_HOST_THREADS, 0)
        eng.forward_raw(ip, n, lp)
        t0 = time.perf_counter()
        for _ in range(max(steps // 3, 2)):
            eng.forward_raw(ip, n, lp)
        dt1 = (time.perf_counter() - t0) / max(steps // 3, 2)
        eng.set_option(V.OPT_HOST_THREADS, 1)
        res["pinned_single_feeding_thread"] = {"value": n / dt1, "unit": "images/s", "ms_per_call": dt1 * 1e3}
        parts = [np.array(h_imgs[i]) for i in range(n)]       # n separate pageable allocations (Network.c:75-93)
        eng.forward_scattered(parts)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            lg = eng.forward_scattered(parts)
        dts = (time.perf_counter() - t0) / reps
        res["scattered"] = {"value": n / dts, "unit": "images/s", "ms_per_call": dts * 1e3, "host_gather_gbs": n * 3 * S * S * 4 / dts / 1e9,
                            "bound": "host memcpy into pinned staging (602 KB per image, read + write) shared by all GPUs' gather threads"}
        res["results_equal"] = bool(int(lg.argmax(1).sum()) == checksum)
        V.pinned_free(ip)
        V.pinned_free(lp)
    return res


def cublas_reference(seconds=2.0):
    """Context for the roofline fraction, measured in the same run on the same GPU: cuBLAS (torch.matmul) 8192^3 with BF16 and
    with FP16 operands, back to back for `seconds` each after a warm-up -- the way MEASURED_PEAKS.json's sustained figure was
    taken.  The engine's default policy multiplies FP16 operands, whose multipliers draw more power than BF16's; the step is
    power-bound, so the library GEMM shows the same gap.  Library code, used here as a yardstick only."""
    try:
        import torch
    except Exception as ex:  # pragma: no cover
        return {"unavailable": repr(ex)}
    out = {"shape": "8192 x 8192 x 8192", "seconds_each": seconds, "how": "torch.matmul (cuBLAS), CUDA events over a back-to-back loop after 1 s of warm-up"}
    n = 8192
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        a = torch.randn(n, n, device="cuda", dtype=dt)
        b = torch.randn(n, n, device="cuda", dtype=dt)
        c = torch.empty(n, n, device="cuda", dtype=dt)
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            for _ in range(10):
                torch.matmul(a, b, out=c)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters, t_end = 0, time.perf_counter() + seconds
        e0.record()
        while time.perf_counter() < t_end:
            for _ in range(20):
                torch.matmul(a, b, out=c)
            iters += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        out[f"{name}_tflops_sustained"] = 2.0 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b, c
    torch.cuda.empty_cache()
    return out


def variants(V, args, prec, peaks, parity_images=None):
    """Driver-visible sub-records: BASELINE.json configs[4] (384x384, 577 tokens, batch 512: the key-blocked attention
    kernel), the non-default BF16 operand set at the headline configuration, and configs[1] (batch-1 latency) on the
    weight tensors the reference ships."""
    out = []

    def measure(S, B, precision, weights, steps, label, residual16=True):
        T = (S // 16) ** 2 + 1
        extra = {}
        with V.Engine(weights, S, max_batch=B, precision=precision) as eng:
            eng.set_class_row_pruning(False)
            eng.set_option(V.OPT_RESIDUAL16, 1 if residual16 else 0)
            imgs = V.synth_images(B, S, 7)
            d_imgs, d_logits = V.dev_alloc(0, imgs.nbytes), V.dev_alloc(0, B * 1000 * 4)
            V.dev_upload(0, d_imgs, imgs)
            for _ in range(3):
                eng.enqueue_device(d_imgs, B, d_logits)
            eng.sync()
            eng.timer_start()
            for _ in range(steps):
                eng.enqueue_device(d_imgs, B, d_logits)
            ms = eng.timer_stop() / steps
            eng.sync()
            eng.profile_enable(True)
            for _ in range(steps):
                eng.enqueue_device(d_imgs, B, d_logits)
            prof = eng.profile_read()
            eng.profile_enable(False)
            name = eng.info()["precision"]
            if not residual16 and parity_images is not None:
                extra["parity"] = parity_block(eng.forward(parity_images), len(parity_images), weights, parity_images, name + ", fp32 residual stream")
            eng.set_option(V.OPT_RESIDUAL16, 1)
            V.dev_free(0, d_imgs)
            V.dev_free(0, d_logits)
        value = B / (ms * 1e-3)
        flop = FLOP_PER_IMAGE[S]
        return {"variant": label, "metric": f"ViT-B/16 {S}x{S} inference throughput", "value": value, "unit": "images/s", "ms_per_step": ms,
                "steps": steps, "warmup": 3, "dtype": name, "config": {"workload": f"ViT-B/16 {S}x{S} synthetic batch {B}, {T} tokens, all rows in the last layer"},
                "model_frac_of_peak": {"burst": flop * value / 1e12 / peaks["bf16_tflops"], "sustained": flop * value / 1e12 / peaks["bf16_tflops_sustained"]},
                "attention_ms_per_step": prof["attention"]["ms"] / steps, "step_breakdown_ms": {k: v["ms"] / steps for k, v in prof.items()}, **extra}

    steps = max(args.steps // 2, 3)
    out.append(measure(384, 512, prec, V.synth_weights(384, 42), steps, "configs[4]: 384x384, batch 512, default precision policy"))
    w224 = V.synth_weights(224, 42)
    other = V.PREC_BF16 if prec != V.PREC_BF16 else V.PREC_FP16
    out.append(measure(224, 1024, other, w224, steps, "headline configuration with the non-default operand set"))
    if prec != V.PREC_BF16:
        out.append(measure(224, 1024, prec, w224, steps, "headline configuration with VIT_OPT_RESIDUAL16 = 0: FP16 operands and an fp32 residual stream for every row "
                           "(the default keeps the patch rows' stream in FP16 and only the class rows' in fp32)", residual16=False))
    # batch-1 latency on the reference's shipped tensors (116 of 152; the 36 missing GEMM weights synthetic, seed 42)
    shipped_dir = ROOT / "baseline" / "_ref" / "Network"
    if shipped_dir.is_dir():
        import vit_hostio as H
        shipped = H.load_weights_dir(str(shipped_dir))
        n_shipped = sum(a is not None for a in shipped)
        w = [np.ascontiguousarray(a) if a is not None and a.size == b.size else b for a, b in zip(shipped, w224)]
        with V.Engine(w, 224, max_batch=8, precision=prec) as eng:
            img = V.synth_images(1, 224, 7)
            h_img, ip = V.pinned_empty((1, 3, 224, 224))
            h_img[...] = img
            h_log, lp = V.pinned_empty((1, 1000))
            d_img, d_log = V.dev_alloc(0, img.nbytes), V.dev_alloc(0, 4000)
            V.dev_upload(0, d_img, img)
            dev_ms, host_ms = [], []
            for i in range(50 + 1000):
                eng.timer_start()
                eng.enqueue_device(d_img, 1, d_log)
                ms = eng.timer_stop()
                if i >= 50:
                    dev_ms.append(ms)
            for i in range(50 + 1000):
                t1 = time.perf_counter()
                eng.forward_raw(ip, 1, lp)
                if i >= 50:
                    host_ms.append((time.perf_counter() - t1) * 1e3)
            info = eng.info()
            top1_shipped = int(h_log.argmax())     # before the pinned buffer goes away
            V.dev_free(0, d_img)
            V.dev_free(0, d_log)
            V.pinned_free(ip)
            V.pinned_free(lp)
        out.append({"variant": "configs[1]: batch-1 latency on the reference's shipped Network/ tensors", "shipped_tensors": n_shipped, "synthetic_tensors": 152 - n_shipped,
                    "dtype": info["precision"], "precision_fallbacks": info["precision_fallbacks"], "runs": 1000, "warmup": 50,
                    "device_ms_median": float(np.median(dev_ms)), "device_ms_p99": float(np.percentile(dev_ms, 99)),
                    "host_to_host_ms_median": float(np.median(host_ms)), "host_to_host_ms_p99": float(np.percentile(host_ms, 99)),
                    "h2d_bytes": int(img.nbytes), "d2h_bytes": 4000, "top1": top1_shipped})
    else:
        out.append({"variant": "configs[1]: batch-1 latency on the reference's shipped Network/ tensors", "unavailable": "baseline/_ref/Network absent (run __graft_entry__.build() where /root/reference is mounted)"})
    return out


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import vit_b200 as V
    rank, local_rank, world = dist_env()
    # The contract is ONE JSON line on stdout.  NCCL prints its version banner there (NCCL_DEBUG=VERSION in this image),
    # other libraries may too: everything but the final line is routed to stderr.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        wait_group = dist.new_group(backend="gloo")   # CPU-side barrier for phases in which one rank drives several GPUs
    n_gpus = world
    prec = {"auto": V.PREC_AUTO, "fp16": V.PREC_FP16, "bf16": V.PREC_BF16}[args.precision]
    B = args.batch
    peaks = load_peaks()
    S = args.img_size
    T = (S // 16) ** 2 + 1
    flop_per_image = FLOP_PER_IMAGE[S]
    weights = V.synth_weights(S, 42)
    eng = V.Engine(weights, S, max_batch=B, n_gpus=1, device_ids=[local_rank], precision=prec)
    info = eng.info()
    dtype_name = info["precision"]      # the operand type the passes run in ("fp16" for the default policy)

    # synthetic batch (seed 7), distinct images per rank; pinned host copy for the end-to-end leg
    h_imgs, h_imgs_ptr = V.pinned_empty((B, 3, S, S))
    V.synth_images(B, S, 7, first_index=rank * B, out=h_imgs)
    h_logits, h_logits_ptr = V.pinned_empty((B, 1000))
    d_imgs = V.dev_alloc(0, h_imgs.nbytes)
    d_logits = V.dev_alloc(0, h_logits.nbytes)
    V.dev_upload(0, d_imgs, h_imgs)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # The headline numbers (value, e2e, roofline, step breakdown) are measured with the FULL last layer -- every
    # multiply-add of the reference's forward (FLOP_PER_IMAGE).  The engine's default (class-row pruning of the last
    # layer, same logits, 6.3 % fewer multiply-adds) is measured separately below and reported under its own key.
    eng.set_class_row_pruning(False)

    # ---- device-resident throughput (inputs in HBM, 617 MB per batch >> 126 MB L2)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # nvidia-smi needs ~0.2 s to deliver its first sample: start it under the warm-up
    for _ in range(max(args.warmup, 3)):
        eng.enqueue_device(d_imgs, B, d_logits)
    eng.sync()
    barrier()
    launches0 = V.launch_count()
    eng.timer_start()
    for _ in range(args.steps):
        eng.enqueue_device(d_imgs, B, d_logits)
    ms_total = eng.timer_stop()
    eng.sync()
    barrier()
    launches = V.launch_count() - launches0
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = n_gpus * B * args.steps / (ms_total * 1e-3)
    # Second region of the same K steps with an event pair around EVERY launch (per-kernel times for the roofline and the
    # step breakdown).  Kept out of the headline region: 1300 event records per step cost ~1-2 %, and an event between two
    # kernels also disables their programmatic dependent launch overlap.
    eng.profile_enable(True)
    eng.timer_start()
    for _ in range(args.steps):
        eng.enqueue_device(d_imgs, B, d_logits)
    ms_profiled = eng.timer_stop() / args.steps
    eng.sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    prof = eng.profile_read()
    eng.profile_enable(False)

    # ---- end to end through vit_cuda_forward: pinned host images in, host logits out, every step
    for _ in range(4):       # the pass schedule adapts to the copy rate it measures in its first calls
        eng.forward_raw(h_imgs_ptr, B, h_logits_ptr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.forward_raw(h_imgs_ptr, B, h_logits_ptr)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_info = eng.info()
    e2e_value = n_gpus * B * args.steps / e2e_s
    top1 = h_logits.argmax(1)
    h_logits_full = h_logits.copy()     # logits of the timed (all-rows) configuration; rank 0's first images go to the parity block

    # ---- the same through the form the reference's own signature hands over (ViT_cuda -> vit_cuda_forward_scattered): one
    #      separately allocated, pageable buffer per image (Network.c:75-93), pageable logits; rank 0 at N = 1 only
    scattered = None
    if world == 1:
        parts = [np.array(h_imgs[i]) for i in range(B)]
        eng.forward_scattered(parts)
        t0 = time.perf_counter()
        reps = max(args.steps // 2, 2)
        for _ in range(reps):
            lg = eng.forward_scattered(parts)
        dt = (time.perf_counter() - t0) / reps
        scattered = {"value": B / dt, "unit": "images/s", "ms_per_step": dt * 1e3, "host_gather_gbs": B * 3 * S * S * 4 / dt / 1e9,
                     "what": "vit_cuda_forward_scattered on B separately allocated pageable images (what ViT_cuda(ImageData*, Network*, float**) passes on), pageable logits out",
                     "equal_to_pinned_path": bool(np.array_equal(lg, h_logits_full))}
        del parts

    # ---- the engine's default configuration: last layer pruned to the class rows (same logits)
    eng.set_class_row_pruning(True)
    for _ in range(3):
        eng.enqueue_device(d_imgs, B, d_logits)
    eng.sync()
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        eng.enqueue_device(d_imgs, B, d_logits)
    ms_pruned = max_over_ranks(eng.timer_stop())
    eng.sync()
    for _ in range(2):
        eng.forward_raw(h_imgs_ptr, B, h_logits_ptr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.forward_raw(h_imgs_ptr, B, h_logits_ptr)
    e2e_pruned_s = max_over_ranks(time.perf_counter() - t0)
    top1_pruned = h_logits.argmax(1)
    eng.set_class_row_pruning(False)

    # ---- optional logit gather over NCCL (the path itself needs no collective: this only shows what gathering
    #      the [B, 1000] fp32 logits of every rank onto every rank costs over NVLink)
    gather_ms = None
    if world > 1:
        t_log = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
        t_all = torch.empty((world * B, 1000), dtype=torch.float32, device="cuda")
        t_log.copy_(torch.from_numpy(h_logits))
        for _ in range(3):
            dist.all_gather_into_tensor(t_all, t_log)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_gather_into_tensor(t_all, t_log)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = max_over_ranks(e0.elapsed_time(e1) / 10)
        assert torch.equal(t_all[rank * B:(rank + 1) * B], t_log)

    # ---- batch-1 latency (BASELINE.json configs[1]), device resident and host-to-host
    lat = None
    if rank == 0 and not args.no_latency:
        dev_ms, host_ms = [], []
        for i in range(20 + 200):
            eng.timer_start()
            eng.enqueue_device(d_imgs, 1, d_logits)
            ms = eng.timer_stop()
            if i >= 20:
                dev_ms.append(ms)
        for i in range(20 + 200):
            t1 = time.perf_counter()
            eng.forward_raw(h_imgs_ptr, 1, h_logits_ptr)
            if i >= 20:
                host_ms.append((time.perf_counter() - t1) * 1e3)
        lat = {"device_ms_median": float(np.median(dev_ms)), "device_ms_p99": float(np.percentile(dev_ms, 99)),
               "host_to_host_ms_median": float(np.median(host_ms)), "host_to_host_ms_p99": float(np.percentile(host_ms, 99)), "runs": 200}

    out = None
    if rank == 0:
        rows = B * T
        step_ms_by_cat = {k: v["ms"] / args.steps for k, v in prof.items()}
        dom = max(GEMM_FLOP_PER_ROW, key=lambda k: step_ms_by_cat[k])  # dominant kernel of the step
        dom_ms = prof[dom]["ms"] / max(prof[dom]["launches"], 1)
        dom_tflops = GEMM_FLOP_PER_ROW[dom] * rows / (dom_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]  # kernel timed inside a long step
        out = {
            "metric": f"ViT-B/16 {S}x{S} inference throughput", "value": value, "unit": "images/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": f"ViT-B/16 {S}x{S} synthetic batch {B} per GPU ({n_gpus * B} total), {T} tokens, 12 layers, random-init weights (seed 42), fp32 residual stream",
                       "batch_per_gpu": B, "parallelism": f"dp{n_gpus}", "last_layer": "all rows (class-row pruning off)",
                       "precision_policy": info["precision_policy"], "operands": f"{dtype_name} x {dtype_name} -> fp32 accumulate (conv_proj: tf32 from the fp32 image)", "l2": f"inputs ({h_imgs.nbytes // 1000000} MB/batch) and activations larger than the 126 MB L2"},
            "model_tflops": flop_per_image * value / 1e12,
            "model_frac_of_peak": {"burst": flop_per_image * value / n_gpus / 1e12 / peaks["bf16_tflops"],
                                   "sustained": flop_per_image * value / n_gpus / 1e12 / peaks["bf16_tflops_sustained"], "peaks": peaks["source"]},
            "roofline": {"kernel": f"gemm_sm100_staged_kernel ({dom})", "bound": "tensor", "achieved": dom_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": dom_tflops / peak, "traffic": NCU_TRAFFIC_BYTES.get(dom) if (B, S) == (1024, 224) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r2_ncu_layer_res16.txt" if (B, S) == (1024, 224) else None,
                         "algorithmic_flop_per_launch": GEMM_FLOP_PER_ROW[dom] * rows,
                         "peak_kind": f"bf16_tflops_sustained ({peaks['source']})",
                         "ms_per_launch": dom_ms, "launches": prof[dom]["launches"]},
            "step_breakdown_ms": step_ms_by_cat, "ms_per_step_with_launch_events": ms_profiled,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(h_imgs.nbytes), "d2h_bytes_per_step": int(h_logits.nbytes),
                    "ms_per_step": e2e_s / args.steps * 1e3, "pass_growth_percent": e2e_info["pass_growth_percent"],
                    "pass_cost_model": {"fixed_us_per_pass": e2e_info["pass_fixed_us"], "kernel_ns_per_image": e2e_info["pass_ns_per_image"]},
                    "h2d_mb_per_s": e2e_info["h2d_mb_per_s"], "reference_signature_form": scattered},
            "gpu_launches": int(launches), "clocks": clocks, "engine": eng.info(), "top1_checksum": int(top1.sum()),
        }
        out["class_row_pruning"] = {
            "note": "engine default: last encoder layer computes out_proj / LayerNorm / MLP for the class rows only (the head reads nothing else); "
                    "same logits, 6.3 % of the multiply-adds not executed; NOT used for value / e2e / roofline above",
            "value": n_gpus * B * args.steps / (ms_pruned * 1e-3), "unit": "images/s", "ms_per_step": ms_pruned / args.steps,
            "e2e_value": n_gpus * B * args.steps / e2e_pruned_s, "executed_flop_per_image": flop_per_image - (2_199_515_136 if S == 224 else 0) if S == 224 else None,
            "top1_equal_to_full": bool((top1_pruned == top1).all())}
        if gather_ms is not None:
            out["nccl_logit_allgather_ms"] = gather_ms
        if lat:
            out["batch1_latency"] = lat
    # ---- done with the per-rank engine
    first_logits = h_logits_full[:PARITY_IMAGES].copy() if rank == 0 else None
    first_images = h_imgs[:PARITY_IMAGES].copy() if rank == 0 else None
    V.dev_free(0, d_imgs)
    V.dev_free(0, d_logits)
    eng.close()
    V.pinned_free(h_imgs_ptr)
    V.pinned_free(h_logits_ptr)

    # ---- N > 1: the C ABI's OWN multi-GPU path -- one process, vit_cuda_init(n_gpus = N), vit_cuda_forward feeding every
    #      GPU from its own host thread -- driven by rank 0 while the other ranks (their engines freed) wait on the CPU
    if world > 1:
        dist.barrier(group=wait_group)
        if rank == 0 and not args.no_inproc:
            out["abi_inproc"] = abi_inproc(V, weights, S, B, world, args.steps, prec, out["e2e"]["value"])
        dist.barrier(group=wait_group)

    if rank == 0:
        if n_gpus == 1 and S == 224 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(first_logits)
            out["parity"] = parity_block(first_logits, PARITY_IMAGES, weights, first_images, dtype_name)
        if n_gpus == 1 and S == 224 and B == 1024 and not args.no_variants:
            out["variants"] = variants(V, args, prec, peaks, first_images if not args.no_cpu_baseline else None)
            out["cublas_reference"] = cublas_reference()
            ref = out["cublas_reference"].get(f"{dtype_name}_tflops_sustained")
            if ref:
                out["roofline"]["frac_of_same_dtype_cublas_sustained_this_run"] = out["roofline"]["achieved"] / ref
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step")
    ap.add_argument("--img-size", type=int, choices=[224, 384], default=224, help="384: BASELINE.json configs[4] (use --batch 512)")
    ap.add_argument("--precision", choices=["auto", "bf16", "fp16"], default="auto", help="auto = the engine's default policy (FP16 operands, BF16 fallback on overflow)")
    ap.add_argument("--no-variants", action="store_true", help="skip the sub-records (384x384, non-default precision, shipped-tensor latency)")
    ap.add_argument("--no-inproc", action="store_true", help="N > 1: skip the single-process vit_cuda_init(n_gpus = N) leg")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    _, _, world = dist_env()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29531", __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
