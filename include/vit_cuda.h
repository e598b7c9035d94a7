/*
 * vit_cuda.h -- C ABI of the B200 (sm_100a) ViT-B/16 inference engine.
 *
 * This is the drop-in boundary for the reference's GPU backend.  It replaces
 *     void initialize_opencl(void);                                   ViT_opencl.h:21, ViT_opencl.c:74
 *     void ViT_opencl(ImageData*, Network*, float **prb);             ViT_opencl.h:18, ViT_opencl.c:785
 *     void Release_opencl(void);                                      ViT_opencl.h:22, ViT_opencl.c:103
 * (sole caller: Main.c:19,57,86).  The reference-signature adaptor ViT_cuda() that a
 * Main.c-style driver calls lives in vit_host.h; it is a thin plain-C wrapper over the
 * three functions below.
 *
 * Plain C: pointers and sizes only, no C++/torch types, no exceptions and no exit()
 * across the boundary.  Every function returning int returns 0 on success or a negative
 * VIT_E_* code; vit_cuda_last_error() then describes the failure.  There is no CPU
 * fallback: without a usable sm_100 device every compute entry point fails.
 *
 * Threading: like the reference (global g_opencl, ViT_opencl.c:33) the engine is one
 * process-wide singleton and is not re-entrant; calls must be serialised by the caller.
 */
#ifndef VIT_CUDA_H
#define VIT_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as the reference's `Network` (Network.h:18-21): one fp32 tensor in host
 * memory, `size` = number of floats.  vit_host.h typedefs `Network` to this. */
typedef struct vit_tensor {
    float* data;
    size_t size;
} vit_tensor;

enum {
    VIT_OK            = 0,
    VIT_E_ARG         = -1,  /* bad argument / tensor size mismatch / not initialised */
    VIT_E_NODEVICE    = -2,  /* no CUDA device, or device is not sm_100 */
    VIT_E_CUDA        = -3,  /* CUDA runtime/driver error (see last_error) */
    VIT_E_NOMEM       = -4,
    VIT_E_DEVICE_TRAP = -5,  /* a kernel watchdog fired (pipeline dead-lock guard) */
    VIT_E_RANGE       = -6   /* vit_cuda_sync: the single-pass softmax flagged a row, or FP16 operands overflowed
                                under VIT_PREC_AUTO -- the engine has switched to the exact softmax / to BF16 and
                                EVERY pass enqueued on that slot since its last vit_cuda_sync must be enqueued again.
                                vit_cuda_init_ex with VIT_PREC_FP16: a weight does not fit FP16. */
};

/* GEMM-operand storage type.  Accumulation, LayerNorm statistics, softmax and the classifier head are always fp32,
 * and so is the residual stream of the rows the head reads (see VIT_OPT_RESIDUAL16 for the other rows). */
enum {
    VIT_PREC_BF16 = 0,       /* BF16 operands, kind::f16 tcgen05, FP32 accumulate.  max |dlogit| vs ViT_seq ~0.03 on
                                random-init weights: outside the stated 2e-2 + 1e-2 |ref| on ~0.3 % of the logits. */
    VIT_PREC_FP16 = 1,       /* FP16 operands (same tensor-core rate, 3 more mantissa bits; meets the stated tolerance
                                with a 3x margin: max |dlogit| 0.006; ~5 % slower than BF16 because its multipliers
                                draw more power in a power-bound step).  Values beyond 65504 overflow: init fails with VIT_E_RANGE if a
                                weight does, and a forward whose activations do fails with VIT_E_RANGE. */
    VIT_PREC_AUTO = 2        /* default (vit_cuda_init, ViT_cuda): FP16 operands, with the BF16 operand set resident
                                beside them (+0.17 GB).  An FP16 overflow anywhere makes a logit non-finite, which the
                                classifier kernel detects; vit_cuda_forward then transparently repeats the call with
                                BF16 operands (after three such calls the engine stays on BF16); a weight that does
                                not fit FP16 selects BF16 at init.  vit_cuda_info reports policy, active type and
                                the number of fallbacks. */
};

#define VIT_NUM_TENSORS 152  /* torchvision vit_b_16 state_dict order, SURVEY.md App. A */
#define VIT_NUM_CLASSES 1000

/* Validate the 152 weight tensors (sizes for img_size 224 or 384), convert them once to
 * the operand precision, replicate them on n_gpus devices (devices 0..n_gpus-1) and
 * allocate per-device workspaces for up to max_batch_per_gpu images per pass.
 * The host weight arrays are not referenced after return (the reference re-uploads them
 * on every op, ViT_opencl.c:115-124).  Operand precision: VIT_PREC_AUTO. */
int vit_cuda_init(const vit_tensor* networks, int n_tensors, int img_size,
                  int max_batch_per_gpu, int n_gpus);

/* As vit_cuda_init, with explicit CUDA device ordinals and operand precision.
 * device_ids == NULL means 0..n_gpus-1. */
int vit_cuda_init_ex(const vit_tensor* networks, int n_tensors, int img_size,
                     int max_batch_per_gpu, int n_gpus, const int* device_ids, int precision);

/* Forward n images.  images_nchw: host, contiguous [n][3][S][S] fp32 (pageable or pinned).
 * logits_out: host [n][1000] fp32 pre-softmax logits.  top1_out: nullable, [n] argmax
 * (lowest index wins ties).  Images are sharded contiguously over the GPUs; each shard is
 * processed in passes of at most max_batch_per_gpu images with H2D / compute / D2H
 * overlapped.  Synchronous: everything is on the host when the call returns. */
int vit_cuda_forward(const float* images_nchw, int n, float* logits_out, int* top1_out);

/* Same for n separately allocated images (images[i] -> [3][S][S] fp32), the form the reference's
 * loader produces (one malloc per image, Network.c:75-93) and ViT_opencl() receives: every pass is
 * gathered into pinned staging buffers owned by the engine while the GPU works on the previous
 * pass.  ViT_cuda() goes through this. */
int vit_cuda_forward_scattered(const float* const* images, int n, float* logits_out, int* top1_out);

/* The contiguous shard [*lo, *hi) of n images that GPU slot g of n_gpus processes in
 * vit_cuda_forward: ceil(n / n_gpus) images per slot, the last ones possibly fewer or none.
 * Pure host arithmetic (usable without a device); returns VIT_E_ARG on bad arguments. */
int vit_cuda_shard_range(int n, int n_gpus, int g, int* lo, int* hi);

/* The pass schedule vit_cuda_forward uses for a shard of n_images on one GPU: pass i covers images
 * [first[i], first[i] + count[i]) of the shard.  The first pass is small (32 images) because its
 * host-to-device copy is the only one not hidden under kernels; each later pass may be three times
 * the previous one, up to max_batch.  Pure host arithmetic.  Returns the number of passes (<= cap)
 * or a negative status. */
int vit_cuda_pass_schedule(int n_images, int max_batch, int* first, int* count, int cap);
/* staged != 0: the schedule for input that has to be gathered into pinned staging first (pageable memory,
 * vit_cuda_forward_scattered): 64 images, then passes of 128 -- the host-side gather runs at about the rate
 * the GPU consumes images, so equal passes keep every gather hidden under the previous pass's kernels. */
int vit_cuda_pass_schedule_ex(int n_images, int max_batch, int staged, int* first, int* count, int cap);
/* The general form: every pass after the first may be growth_percent / 100 times the previous one (clamped to 100..300).
 * vit_cuda_forward derives the factor per GPU from the copy and kernel times it measured in its previous calls (how much
 * faster a pass's host-to-device copy is than its kernels, minus a 15 % reserve): 300 for a lone GPU on PCIe Gen5, less when
 * several GPUs pull from the same host memory at once -- a pass whose copy is slower than the previous pass's kernels would
 * otherwise wait for it (measured at 8 GPUs in one process: 47.8 ms per 8192 images with the fixed factor 3). */
int vit_cuda_pass_schedule_growth(int n_images, int max_batch, int staged, int growth_percent, int* first, int* count, int cap);
/* The schedule vit_cuda_forward uses for pinned input (VIT_OPT_WAVE_PASSES, default on), from a cost model of the pipeline: the
 * kernels of a pass of n images take fixed_us_per_pass + kernel_us_per_image * n (measured on a B200: ~0.7 ms + 33 us * n at
 * 224x224 -- some sixty kernels, each with its own fill and drain), its host-to-device copy copy_us_per_image * n (11 us at
 * 55 GB/s), the copies run back to back.  Pass i + 1 is the largest whose copy still hides under pass i's kernels (10 % reserve);
 * among the pass counts P the one with the smallest exposed cost -- the first pass's copy + P fixed costs -- is taken, with the
 * smallest first pass that reaches n_images in P passes.  (1024, 1024, 10.9, 33, 700) -> 68 + 243 + 713: three passes instead of
 * the geometric schedule's four.  vit_cuda_forward fits the three numbers per GPU to the pass times of its previous calls. */
int vit_cuda_pass_schedule_model(int n_images, int max_batch, double copy_us_per_image, double kernel_us_per_image, double fixed_us_per_pass,
                                 int* first, int* count, int cap);
/* ... and every schedule is then re-cut at wave-efficient sizes (VIT_OPT_WAVE_PASSES; shown here for the geometric one).
 * The GEMMs are persistent over sm_count / 2 CTA pairs with 256-row tiles, so a pass costs whole waves of tiles: 32 images of
 * 197 tokens are 25 row tiles = 75 out_proj / mlp_3 tiles = TWO waves on a 148-SM part, 31 images are one; 128 images need
 * 5 / 13 / 17 waves (N = 768 / 2304 / 3072), 127 need 4 / 12 / 16.  Every pass but the last becomes the size in
 * [0.8, 1.0] x its scheduled size ([0.8, 1.08] below 128 images) with the fewest tile waves per image; what that leaves over
 * moves to the later passes.
 * tokens = (img_size / 16)^2 + 1.  Pure host arithmetic, like the functions above: (1024, 1024, pinned, 300 %, 197, 148) ->
 * 31 + 96 + 288 + 609. */
int vit_cuda_pass_schedule_waves(int n_images, int max_batch, int staged, int growth_percent, int tokens, int sm_count,
                                 int* first, int* count, int cap);

/* Device-resident variant for one GPU slot (0 <= gpu_slot < n_gpus): d_images and d_logits
 * are device pointers on that GPU, n <= max_batch_per_gpu.  Work is enqueued on the
 * engine's stream for that slot and the call returns after it has completed. */
int vit_cuda_forward_device(int gpu_slot, const float* d_images, int n, float* d_logits);

/* Enqueue-only form of the above (no synchronisation), for timing with events, plus the
 * matching synchronise.  vit_cuda_stream() returns the cudaStream_t of a slot as void*. */
int   vit_cuda_enqueue_device(int gpu_slot, const float* d_images, int n, float* d_logits);
int   vit_cuda_sync(int gpu_slot);
void* vit_cuda_stream(int gpu_slot);

/* Operand-precision weight cache.  vit_cuda_save_weight_cache writes slot 0's device weight arena -- the converted
 * GEMM operands of every resident precision, the LayerNorm-folded copies of in_proj / mlp_0 with their column sums
 * and constant vectors, the tf32-rounded conv_proj weight and the small fp32 tensors -- as ONE checksummed file;
 * vit_cuda_init_from_cache brings an engine up from it with one read and one host-to-device copy per GPU: no 152
 * file reads, no fp32 upload, no conversion, no folding (the reference re-reads and re-rounds 330 MB of fp32 per
 * start and re-uploads weights on every op, Network.c:119-194, ViT_opencl.c:115-124).  Image size and precision
 * policy are those the cache was written with.  VIT_E_ARG on a bad / corrupt / mismatching file. */
int vit_cuda_save_weight_cache(const char* path);
int vit_cuda_init_from_cache(const char* path, int max_batch_per_gpu, int n_gpus, const int* device_ids);

/* Release all device memory, streams and pinned staging.  Safe to call when not
 * initialised. */
void vit_cuda_free(void);

/* Thread-local, never NULL; valid until the next failing call on this thread. */
const char* vit_cuda_last_error(void);

/* Number of kernels the engine launched since init (all slots); used by bench.py to
 * report gpu_launches. */
long long vit_cuda_launch_count(void);

/* Softmax variant of the fused attention kernel.  Default (0): ONE pass over the
 * scores with the exponent offset taken from 8 of the row's scores -- exact unless some logit
 * exceeds those by more than ~110, which the kernel detects; vit_cuda_forward then transparently
 * repeats the call with the exact variant, vit_cuda_sync returns VIT_E_RANGE and switches the
 * engine over.  1: always the exact two-pass softmax (row maximum first, ViT_seq.c:178-191).
 * The environment variable VIT_ATTN_EXACT=1 selects it at init. */
int vit_cuda_set_attention_exact(int on);

/* Last-layer pruning (default on; VIT_PRUNE_LAST=0 at init or this call with 0 turns it off).  The
 * head reads only the class token of the last encoder layer (the reference computes all 197 rows and
 * keeps row 0, ViT_seq.c:429-433), and after the last attention no token reads another one: with
 * pruning the last layer runs in_proj for all rows (keys and values), attention for the class query
 * only, and out_proj / LayerNorm / MLP on the [n][768] class rows.  The logits are the same
 * function of the input; 6.3 % of the model's multiply-adds are never executed. */
int vit_cuda_set_class_row_pruning(int on);

/* Run-time switches (all also readable).  Each has an environment variable of the same meaning that is read ONCE,
 * at vit_cuda_init*: VIT_ATTN_EXACT, VIT_PRUNE_LAST, VIT_LN_FUSED, VIT_PDL, VIT_GRAPHS, VIT_HOST_THREADS, VIT_RESIDUAL16, VIT_WAVE_PASSES. */
enum {
    VIT_OPT_ATTENTION_EXACT   = 0,  /* 1: always the exact two-pass softmax (default 0, see vit_cuda_set_attention_exact) */
    VIT_OPT_CLASS_ROW_PRUNING = 1,  /* default 1, see vit_cuda_set_class_row_pruning */
    VIT_OPT_LN_FUSED          = 2,  /* default 1: LayerNorm folded into the GEMMs; 0: separate warp-per-row LayerNorm kernels */
    VIT_OPT_PDL               = 3,  /* default 1: programmatic dependent launch between the kernels of a pass */
    VIT_OPT_GRAPHS            = 4,  /* default 1: passes of <= 8 images replay a captured CUDA graph */
    VIT_OPT_HOST_THREADS      = 5,  /* default 1: vit_cuda_forward feeds every GPU from its own host thread (n_gpus > 1);
                                       0: one thread issues for all GPUs in turn */
    VIT_OPT_RESIDUAL16        = 6,  /* default 1: with FP16 operands and folded LayerNorm the patch rows' residual stream is kept in
                                       FP16 (the rows out_proj / mlp_3 update ARE the next GEMM's operand; no fp32 row beside
                                       them: -21 % HBM traffic per step), while each image's class-token row -- the one row the
                                       head reads -- keeps an fp32 master copy that the same epilogues update.  Measured against
                                       ViT_seq this is as close as the fp32 stream (max |dlogit| 0.0053 vs 0.0060 on the bench's
                                       parity block).  Ignored for BF16 operands.  0: fp32 residual stream for every row. */
    VIT_OPT_WAVE_PASSES       = 7   /* default 1: vit_cuda_forward sizes its passes by the measured cost model
                                       (vit_cuda_pass_schedule_model) and cuts them at wave-efficient sizes
                                       (vit_cuda_pass_schedule_waves); 0: the plain geometric schedule */
};
int vit_cuda_set_option(int option, int value);
int vit_cuda_get_option(int option, int* value);

/* Facts about the engine/device, for logs: fills up to n entries of
 * {sm_count, cc_major, cc_minor, max_batch, tokens, active operand precision (VIT_PREC_BF16 / FP16), n_gpus,
 *  ws_bytes>>20, attention_exact, attention_fallbacks, class_row_pruning, precision policy (VIT_PREC_*),
 *  precision_fallbacks, weight_bytes>>20, pass-schedule growth percent of slot 0 (see vit_cuda_pass_schedule_growth),
 *  measured H2D copy rate of slot 0 in MB/s (0 until measured), fitted fixed cost of a pass on slot 0 in microseconds and
 *  kernel time per image in nanoseconds (vit_cuda_pass_schedule_model; 0 until a call with two passes has been timed)}. */
int vit_cuda_info(long long* out, int n);

/* CUDA-event stopwatch on a slot's stream: start records an event, stop records a second one,
 * waits for it and returns the elapsed device time in milliseconds. */
int vit_cuda_timer_start(int gpu_slot);
int vit_cuda_timer_stop(int gpu_slot, float* ms);

/* Per-kernel-category device timing.  While enabled, every launch of the forward is bracketed
 * by an event pair on the slot's stream; vit_cuda_profile_read synchronises, returns the summed
 * milliseconds and launch count per category (VIT_PROF_*) since the last read, and resets. */
enum {
    VIT_PROF_CLASS_ROWS = 0,   /* class-token rows (there is no patch extraction: conv_proj reads the image) */
    VIT_PROF_EMBED_GEMM,     /* conv_proj GEMM (tf32, straight from the fp32 image) */
    VIT_PROF_LAYERNORM,      /* ln_1 + ln_2 */
    VIT_PROF_QKV_GEMM,       /* in_proj */
    VIT_PROF_ATTENTION,      /* fused softmax(QK^T)V */
    VIT_PROF_OUT_GEMM,       /* out_proj + residual */
    VIT_PROF_FC1_GEMM,       /* mlp_0 + GELU */
    VIT_PROF_FC2_GEMM,       /* mlp_3 + residual */
    VIT_PROF_HEAD,           /* final LN + classifier */
    VIT_PROF_NCAT
};
int vit_cuda_profile_enable(int on);
int vit_cuda_profile_read(int gpu_slot, double* total_ms, long long* launches, int ncat);

/* Device memory helpers so that a host program without a CUDA toolchain (plain C driver,
 * ctypes) can stage device-resident inputs for vit_cuda_forward_device. */
int vit_cuda_dev_alloc(int gpu_slot, size_t bytes, void** d_ptr);
int vit_cuda_dev_free(int gpu_slot, void* d_ptr);
int vit_cuda_dev_upload(int gpu_slot, void* d_dst, const void* h_src, size_t bytes);
int vit_cuda_dev_download(int gpu_slot, void* h_dst, const void* d_src, size_t bytes);
int vit_cuda_host_alloc_pinned(size_t bytes, void** h_ptr);
int vit_cuda_host_free_pinned(void* h_ptr);

/* ------------------------------------------------------------------------------------------
 * Single-operator entry points.  Each runs ONE kernel of the hot path on device 0 with host
 * inputs/outputs (upload, launch, download); they exist so that every kernel can be checked
 * against the oracle's restatement of the corresponding reference function.  They use the
 * same kernels, tile shapes and epilogues as vit_cuda_forward.  `precision` is VIT_PREC_*.
 * ------------------------------------------------------------------------------------------ */

/* epilogue selectors for vit_cuda_op_linear */
enum {
    VIT_EPI_BIAS          = 0,  /* y = x W^T + b                 (in_proj, ViT_seq.c:134-147)  */
    VIT_EPI_BIAS_GELU     = 1,  /* y = gelu(x W^T + b)           (mlp_0,   ViT_seq.c:260-264)  */
    VIT_EPI_BIAS_RESIDUAL = 2   /* y = r + x W^T + b, fp32 out   (out_proj/mlp_3 + residual,
                                                                  ViT_seq.c:219-227,286-288,266,297-299) */
};

/* x: [m][k] fp32 (rounded to the operand precision on upload), W: [n][k] fp32 (ditto),
 * b: [n] fp32, residual: [m][n] fp32 or NULL, y: [m][n] fp32 (for the two operand-precision
 * outputs the device result is widened back to fp32).  Replaces linear_layer,
 * ViT_seq.c:240-250 / linear_forward_kernel, fc1_kernel, fc2_kernel in kernel.cl. */
int vit_cuda_op_linear(const float* x, const float* W, const float* b, const float* residual,
                       float* y, int m, int n, int k, int epilogue, int precision);

/* LayerNorm folded into the following linear layer, as the forward pass runs in_proj and mlp_0:
 * y = epilogue(LN(x; ln_w, ln_b) W^T + b) for x [m][768] fp32, W [n][768], epilogue VIT_EPI_BIAS or
 * VIT_EPI_BIAS_GELU.  The GEMM multiplies the operand-precision copy of the RAW rows by the folded
 * weights ln_w (.) W and applies the row statistics in its epilogue (csrc/gemm_sm100.cuh).
 * Replaces layer_norm + linear_layer, ViT_seq.c:103-121 + 240-250 (called at :281-283, :291-294). */
int vit_cuda_op_ln_linear(const float* x, const float* ln_w, const float* ln_b, const float* W,
                          const float* b, float* y, int m, int n, int epilogue, int precision);

/* The residual GEMM in its LayerNorm-producer form (out_proj / mlp_3 of the forward pass):
 * y = residual + x W^T + b  (fp32, [m][768]; x [m][k], W [768][k]), plus what the next folded GEMM
 * needs: y_cast = y rounded to the operand precision (widened back to fp32 here) and the per-row
 * sum and sum of squares of y (added up in the consumer's order). */
int vit_cuda_op_linear_residual_stats(const float* x, const float* W, const float* b,
                                      const float* residual, float* y, float* y_cast,
                                      float* row_sum, float* row_sumsq, int m, int k, int precision);

/* LayerNorm rows of x [rows][768] (fp32) -> y [rows][768] (operand precision widened to
 * fp32).  Replaces layer_norm, ViT_seq.c:103-121 / layer_norm_kernel, kernel.cl:6-80
 * (with the oracle's eps 1e-6, which the OpenCL kernel drops). */
int vit_cuda_op_layernorm(const float* x, const float* w, const float* b, float* y,
                          int rows, int precision);

/* Fused softmax(QK^T/8)V for `batch` images of `tokens` tokens, 12 heads of 64.
 * qkv: [batch*tokens][2304] fp32 (Q|K|V, rounded to operand precision on upload),
 * out: [batch*tokens][768] fp32.  Replaces ViT_seq.c:156-215 / the per-head
 * MHA_gemm_kernel + softmax_reduction_kernel loop, ViT_opencl.c:546-564. */
int vit_cuda_op_attention(const float* qkv, float* out, int batch, int tokens, int precision);

/* Debug: runs the attention kernel as above and returns SM-clock timestamps of the pipeline events
 * of CTA 0 (19 warps x 16 items x 8 events of uint64, layout in csrc/attention_sm100.cuh) instead of
 * the result.  Used by tools/attn_trace.py to read the kernel's timeline; not part of the hot path. */
int vit_cuda_debug_attention_trace(const float* qkv, int batch, int tokens, int precision,
                                   unsigned long long* trace, int trace_len);

/* Patch embedding for `batch` images [batch][3][S][S]: conv_proj + class token + position
 * embedding -> out [batch*tokens][768] fp32.  Replaces Conv2d/flatten_transpose/
 * class_token/pos_emb, ViT_seq.c:25-101 / Conv2d_Kernel, kernel.cl:120-175.  The GEMM reads the fp32
 * image directly (5-D TMA view, no im2col buffer) as kind::tf32: pixels are used with their 13 low
 * mantissa bits ignored, conv_w rounded to tf32; `precision` only selects the type of the
 * operand-precision copy of the rows that the kernel emits beside the fp32 result.
 * out_cast (nullable): that copy widened to fp32. */
int vit_cuda_op_embed(const float* images, const float* cls, const float* conv_w,
                      const float* conv_b, const float* pos, float* out, float* out_cast,
                      int batch, int img_size, int precision);

/* Final LayerNorm of the class rows + classifier: x [batch*tokens][768] fp32 ->
 * logits [batch][1000] fp32 (all fp32 arithmetic).  Replaces ViT_seq.c:429-435. */
int vit_cuda_op_head(const float* x, const float* ln_w, const float* ln_b,
                     const float* head_w, const float* head_b, float* logits,
                     int batch, int tokens);

/* One whole encoder block of the INITIALISED engine (its weights of block `layer`, its image size, its current precision
 * and options) on slot 0: x [batch*tokens][768] fp32 -> y, the same five kernels in the same configuration as the forward
 * pass runs them, all rows.  Replaces Encoder, ViT_seq.c:271-302 / Encoder_opencl, ViT_opencl.c:732-782. */
int vit_cuda_op_encoder_block(const float* x, float* y, int batch, int layer);

#ifdef __cplusplus
}
#endif
#endif /* VIT_CUDA_H */
