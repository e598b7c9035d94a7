/*
 * vit_main.c -- plain-C driver with the reference Main.c flow (Main.c:18-88): load images ->
 * load weights -> forward -> argmax + result file -> comparator, with ViT_opencl() replaced by
 * ViT_cuda().  Paths and the image count are arguments instead of being hard-coded
 * (Main.c:22,30,40,45; comparator.c:8).
 *
 *   vit_main --images Data/input-100.bin --weights Network [--n N] [--gpus G] [--max-batch B]
 *            [--precision auto|fp16|bf16] [--result Data/cuda_result.txt] [--answer Data/answer_result.txt]
 *            [--timing]                 per-stage device times, as Encoder_opencl prints them (ViT_opencl.c:745-779)
 *            [--stream CHUNK]           read the image file CHUNK images at a time (bounded memory, reader thread
 *                                       overlapped with the GPU) instead of load_image_data's whole-file read
 *            [--operand-cache FILE]     start from / write the engine's operand-precision weight cache
 *   vit_main --synthetic N [--img 224] ...     seeded synthetic images + weights (no files needed)
 *
 *   vit_main --backend seq --seq-lib libvit_ref.so ...    the reference's CPU path (Main.c:48-53): ViT_seq() is taken from
 *                                       a shared library the caller supplies -- the reference's own ViT_seq.c compiled by them
 *                                       (same ImageData / Network layouts).  This program contains no CPU implementation of
 *                                       the forward pass and never falls back to one.
 */
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "vit_host.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static int parse_precision(const char* v) {
    if (!strcmp(v, "fp16")) return VIT_PREC_FP16;
    if (!strcmp(v, "bf16")) return VIT_PREC_BF16;
    return VIT_PREC_AUTO;
}

static const char* const kStage[VIT_PROF_NCAT] = {"class_token rows", "conv_proj (tf32 GEMM)", "layer_norm (unfused only)", "in_proj (LN folded)",
                                                  "attention", "out_proj + residual", "mlp_0 + GELU (LN folded)", "mlp_3 + residual", "final LN + head"};

/* Per-stage device time of one more pass over the same images, every launch bracketed by CUDA events. */
static void print_stage_times(const float* const* ptrs, int n, float* logits) {
    double ms[VIT_PROF_NCAT];
    long long cnt[VIT_PROF_NCAT];
    if (vit_cuda_profile_enable(1) != 0) return;
    if (vit_cuda_forward_scattered(ptrs, n, logits, NULL) == 0 && vit_cuda_profile_read(0, ms, cnt, VIT_PROF_NCAT) == 0) {
        double total = 0;
        for (int k = 0; k < VIT_PROF_NCAT; ++k) total += ms[k];
        printf("per-stage device time, GPU slot 0 (%d images over all GPUs):\n", n);
        for (int k = 0; k < VIT_PROF_NCAT; ++k)
            if (cnt[k]) printf("  %-28s %9.3f ms  %5.1f %%  (%lld launches)\n", kStage[k], ms[k], 100.0 * ms[k] / total, cnt[k]);
        printf("  %-28s %9.3f ms\n", "sum", total);
    }
    vit_cuda_profile_enable(0);
}

/* ---- --stream: a reader thread fills one pinned buffer while the GPU works on the other */
typedef struct {
    vit_image_stream* s;
    float* dst;
    int max_images, got;
} read_job;
static void* read_chunk(void* arg) {
    read_job* j = (read_job*)arg;
    j->got = vit_image_stream_read(j->s, j->dst, j->max_images);
    return NULL;
}

int main(int argc, char** argv) {
    const char *images_path = NULL, *weights_dir = NULL, *result_path = "cuda_result.txt", *answer_path = NULL, *cache_out = NULL,
               *operand_cache = NULL, *backend = "cuda", *seq_lib = NULL;
    int n_limit = 0, synthetic = 0, img = 224, timing = 0, stream_chunk = 0;
    vit_host_config cfg = {1, 256, VIT_PREC_AUTO};
    for (int i = 1; i < argc; ++i) {
        const char* a = argv[i];
        const char* v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if (!strcmp(a, "--images") && v) images_path = v, ++i;
        else if (!strcmp(a, "--weights") && v) weights_dir = v, ++i;
        else if (!strcmp(a, "--save-weight-cache") && v) cache_out = v, ++i;
        else if (!strcmp(a, "--operand-cache") && v) operand_cache = v, ++i;
        else if (!strcmp(a, "--result") && v) result_path = v, ++i;
        else if (!strcmp(a, "--answer") && v) answer_path = v, ++i;
        else if (!strcmp(a, "--n") && v) n_limit = atoi(v), ++i;
        else if (!strcmp(a, "--gpus") && v) cfg.n_gpus = atoi(v), ++i;
        else if (!strcmp(a, "--max-batch") && v) cfg.max_batch_per_gpu = atoi(v), ++i;
        else if (!strcmp(a, "--synthetic") && v) synthetic = atoi(v), ++i;
        else if (!strcmp(a, "--img") && v) img = atoi(v), ++i;
        else if (!strcmp(a, "--precision") && v) cfg.precision = parse_precision(v), ++i;
        else if (!strcmp(a, "--stream") && v) stream_chunk = atoi(v), ++i;
        else if (!strcmp(a, "--timing")) timing = 1;
        else if (!strcmp(a, "--backend") && v) backend = v, ++i;
        else if (!strcmp(a, "--seq-lib") && v) seq_lib = v, ++i;
        else {
            fprintf(stderr, "usage: %s (--images FILE --weights DIR|CACHEFILE | --synthetic N [--img S]) [--n N] [--gpus G] "
                            "[--max-batch B] [--precision auto|fp16|bf16] [--result FILE] [--answer FILE] [--timing] [--stream CHUNK] "
                            "[--operand-cache FILE] [--save-weight-cache FILE] [--backend cuda|seq --seq-lib LIB]\n", argv[0]);
            return 2;
        }
    }
    if (!synthetic && (!images_path || !weights_dir)) {
        fprintf(stderr, "need --images and --weights, or --synthetic N\n");
        return 2;
    }
    const int use_seq = !strcmp(backend, "seq");
    if (!use_seq && strcmp(backend, "cuda")) {
        fprintf(stderr, "--backend must be cuda or seq\n");
        return 2;
    }
    void (*ref_vit_seq)(ImageData*, Network*, float**) = NULL;
    if (use_seq) {
        if (!seq_lib || stream_chunk > 0 || operand_cache) {
            fprintf(stderr, "--backend seq needs --seq-lib LIB (a build of the reference's ViT_seq.c) and neither --stream nor --operand-cache\n");
            return 2;
        }
        void* h = dlopen(seq_lib, RTLD_NOW | RTLD_LOCAL);
        if (!h || !(*(void**)(&ref_vit_seq) = dlsym(h, "ViT_seq"))) {
            fprintf(stderr, "--seq-lib %s: %s\n", seq_lib, dlerror());
            return 1;
        }
    }
    vit_host_set_config(&cfg);
    if (!use_seq && initialize_cuda() != 0) return 1;

    /* ---- weights: the operand cache if it exists, else the reference's Network/ directory (or an fp32 blob) */
    static Network network[VIT_NUM_TENSORS];
    int have_network = 0, engine_from_cache = 0;
    struct stat sb;
    double tw0 = now_s();
    if (operand_cache && stat(operand_cache, &sb) == 0 && S_ISREG(sb.st_mode)) {
        if (vit_cuda_init_from_cache(operand_cache, cfg.max_batch_per_gpu, cfg.n_gpus, NULL) != 0) {
            fprintf(stderr, "operand cache: %s\n", vit_cuda_last_error());
            return 1;
        }
        engine_from_cache = 1;
        printf("engine up from operand cache %s in %.3f s (no fp32 weights read, nothing converted)\n", operand_cache, now_s() - tw0);
    } else if (synthetic) {
        if (vit_synth_weights(network, VIT_NUM_TENSORS, img, 42) != 0) return 1;
        have_network = 1;
    } else if (stat(weights_dir, &sb) == 0 && S_ISREG(sb.st_mode)) {   /* an fp32 weight blob written by --save-weight-cache */
        int cached_img = 0;
        if (load_weights_blob(weights_dir, network, VIT_NUM_TENSORS, &cached_img) != 0) return 1;
        printf("loaded %d weight tensors (img_size %d) from blob %s\n", VIT_NUM_TENSORS, cached_img, weights_dir);
        have_network = 1;
    } else {
        const int loaded = load_weights(weights_dir, network, VIT_NUM_TENSORS);
        if (loaded < 0) return 1;
        printf("loaded %d / %d weight tensors from %s\n", loaded, VIT_NUM_TENSORS, weights_dir);
        have_network = 1;
    }

    /* ---- images */
    ImageData* images = NULL;
    vit_image_stream* stream = NULL;
    int n = 0;
    if (synthetic) {
        const size_t per = (size_t)3 * img * img;
        images = (ImageData*)calloc((size_t)synthetic, sizeof(ImageData));
        for (int i = 0; i < synthetic; ++i) {
            images[i].n = synthetic; images[i].c = 3; images[i].h = img; images[i].w = img;
            images[i].data = (float*)malloc(per * sizeof(float));
            vit_synth_images(images[i].data, 1, img, 7, i);
        }
        n = synthetic;
    } else if (stream_chunk > 0) {
        int c = 0, h = 0, w = 0;
        stream = vit_image_stream_open(images_path, &n, &c, &h, &w);
        if (!stream) return 1;
        if (c != 3 || h != w) {
            fprintf(stderr, "unsupported image shape %d x %d x %d\n", c, h, w);
            return 1;
        }
        img = h;
    } else {
        images = load_image_data(images_path);
        if (!images) return 1;
        img = images[0].h;
        n = images[0].n;
    }
    if (have_network && vit_validate_weights(network, VIT_NUM_TENSORS, img) != 0) return 1;
    if (cache_out && have_network) {
        if (save_weights_blob(cache_out, network, VIT_NUM_TENSORS, img) != 0) return 1;
        printf("wrote fp32 weight blob %s\n", cache_out);
    }
    if (n_limit > 0 && n_limit < n) {
        n = n_limit;
        if (images) images[0].n = n; /* the callee reads the loop bound here, as in Main.c:45-46 */
    }
    float** probabilities = (float**)malloc(sizeof(float*) * (size_t)n);
    for (int i = 0; i < n; ++i) probabilities[i] = (float*)malloc(sizeof(float) * VIT_NUM_CLASSES);
    float* logits = (float*)malloc((size_t)n * VIT_NUM_CLASSES * sizeof(float));

    printf("=====================Start========================\n");
    if (use_seq) {
        /* Main.c:48-53: the sequential CPU path, from the caller's build of the reference */
        if (img != 224) {
            fprintf(stderr, "--backend seq: the reference's ViT_seq hard-codes 224x224 (ViT_seq.c:10)\n");
            return 1;
        }
        const double t0 = now_s();
        ref_vit_seq(images, network, probabilities);
        printf("Sequential time: %f sec, %d images, %.3f images/s\n", now_s() - t0, n, n / (now_s() - t0));
        if (write_results(result_path, probabilities, n) != 0) return 1;
        int cmp_rc = 0;
        if (answer_path) {
            const int cmp = comparator_files(result_path, answer_path, n);
            if (cmp == 0) printf("Comparator: the two files match on %d lines.\n", n);
            else printf("Comparator: %d differences.\n", cmp);
            cmp_rc = cmp == 0 ? 0 : 3;
        }
        return cmp_rc;
    }
    if (stream || engine_from_cache) {
        /* these two modes talk to the engine directly (the reference signature has no way to pass a file or a cache) */
        if (!engine_from_cache) {
            tw0 = now_s();
            if (vit_cuda_init_ex(network, VIT_NUM_TENSORS, img, cfg.max_batch_per_gpu, cfg.n_gpus, NULL, cfg.precision) != 0) {
                fprintf(stderr, "init: %s\n", vit_cuda_last_error());
                return 1;
            }
            printf("engine up from %d fp32 tensors in %.3f s (upload + conversion + LayerNorm folding)\n", VIT_NUM_TENSORS, now_s() - tw0);
        }
        const double t0 = now_s();
        if (stream) {
            const size_t per = (size_t)3 * img * img;
            float* buf[2] = {NULL, NULL};
            for (int b = 0; b < 2; ++b)
                if (vit_cuda_host_alloc_pinned((size_t)stream_chunk * per * sizeof(float), (void**)&buf[b]) != 0) {
                    fprintf(stderr, "pinned buffer: %s\n", vit_cuda_last_error());
                    return 1;
                }
            read_job job = {stream, buf[0], stream_chunk < n ? stream_chunk : n, 0};
            read_chunk(&job);
            int done = 0, cur = 0;
            while (done < n && job.got > 0) {
                const int have = job.got;
                pthread_t th;
                read_job next = {stream, buf[cur ^ 1], (n - done - have) < stream_chunk ? (n - done - have) : stream_chunk, 0};
                const int more = next.max_images > 0;
                if (more && pthread_create(&th, NULL, read_chunk, &next) != 0) return 1;
                if (vit_cuda_forward(buf[cur], have, logits + (size_t)done * VIT_NUM_CLASSES, NULL) != 0) {
                    fprintf(stderr, "forward: %s\n", vit_cuda_last_error());
                    return 1;
                }
                if (more) pthread_join(th, NULL);
                done += have;
                job = next;
                cur ^= 1;
                if (!more) break;
            }
            if (done != n) {
                fprintf(stderr, "stream: %d of %d images read\n", done, n);
                return 1;
            }
            printf("CUDA time: %f sec for %d images streamed in chunks of %d (file read overlapped), %.1f images/s\n", now_s() - t0, n, stream_chunk,
                   n / (now_s() - t0));
            vit_cuda_host_free_pinned(buf[0]);
            vit_cuda_host_free_pinned(buf[1]);
        } else {
            const float** ptrs = (const float**)malloc((size_t)n * sizeof(*ptrs));
            for (int i = 0; i < n; ++i) ptrs[i] = images[i].data;
            if (vit_cuda_forward_scattered(ptrs, n, logits, NULL) != 0) {
                fprintf(stderr, "forward: %s\n", vit_cuda_last_error());
                return 1;
            }
            printf("CUDA time: %f sec, %d images, %.1f images/s\n", now_s() - t0, n, n / (now_s() - t0));
            if (timing) print_stage_times(ptrs, n, logits);
            free(ptrs);
        }
        for (int i = 0; i < n; ++i) vit_softmax(logits + (size_t)i * VIT_NUM_CLASSES, probabilities[i], VIT_NUM_CLASSES);
    } else {
        /* the reference's flow, ViT_opencl() -> ViT_cuda() */
        const double t0 = now_s();
        ViT_cuda(images, network, probabilities); /* first call uploads the weights */
        const double t1 = now_s();
        if (ViT_cuda_status() != 0) return 1;
        ViT_cuda(images, network, probabilities);
        const double t2 = now_s();
        if (ViT_cuda_status() != 0) return 1;
        printf("CUDA time: %f sec (first call incl. weight upload), %f sec (steady), %d images, %.1f images/s\n",
               t1 - t0, t2 - t1, n, n / (t2 - t1));
        if (timing) {
            const float** ptrs = (const float**)malloc((size_t)n * sizeof(*ptrs));
            for (int i = 0; i < n; ++i) ptrs[i] = images[i].data;
            print_stage_times(ptrs, n, logits);
            free(ptrs);
        }
    }
    if (operand_cache && !engine_from_cache) {
        if (vit_cuda_save_weight_cache(operand_cache) != 0) {
            fprintf(stderr, "operand cache: %s\n", vit_cuda_last_error());
            return 1;
        }
        printf("wrote operand cache %s\n", operand_cache);
    }
    {
        long long info[14] = {0};
        if (vit_cuda_info(info, 14) == 0)
            printf("engine: %lld GPU(s), operands %s (policy %s, %lld fallback(s)), weights %lld MiB, workspace %lld MiB per GPU\n", info[6],
                   info[5] == VIT_PREC_FP16 ? "fp16" : "bf16", info[11] == VIT_PREC_AUTO ? "auto" : (info[11] == VIT_PREC_FP16 ? "fp16" : "bf16"),
                   info[12], info[13], info[7]);
    }

    if (write_results(result_path, probabilities, n) != 0) return 1;
    int rc = 0;
    if (answer_path) {
        const int cmp = comparator_files(result_path, answer_path, n);
        if (cmp == 0) printf("Comparator: the two files match on %d lines.\n", n);
        else printf("Comparator: %d differences.\n", cmp);
        rc = cmp == 0 ? 0 : 3;
    }
    if (stream) vit_image_stream_close(stream);
    Release_cuda();
    return rc;
}
