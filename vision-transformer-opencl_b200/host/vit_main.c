/*
 * vit_main.c -- plain-C driver with the reference Main.c flow (Main.c:18-88): load images ->
 * load weights -> forward -> argmax + result file -> comparator, with ViT_opencl() replaced by
 * ViT_cuda().  Paths and the image count are arguments instead of being hard-coded
 * (Main.c:22,30,40,45; comparator.c:8).
 *
 *   vit_main --images Data/input-100.bin --weights Network [--n N] [--gpus G] [--max-batch B]
 *            [--precision bf16|fp16] [--result Data/cuda_result.txt] [--answer Data/answer_result.txt]
 *   vit_main --synthetic N [--img 224] ...     seeded synthetic images + weights (no files needed)
 */
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

int main(int argc, char** argv) {
    const char *images_path = NULL, *weights_dir = NULL, *result_path = "cuda_result.txt", *answer_path = NULL, *cache_out = NULL;
    int n_limit = 0, synthetic = 0, img = 224;
    vit_host_config cfg = {1, 256, VIT_PREC_BF16};
    for (int i = 1; i < argc; ++i) {
        const char* a = argv[i];
        const char* v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if (!strcmp(a, "--images") && v) images_path = v, ++i;
        else if (!strcmp(a, "--weights") && v) weights_dir = v, ++i;
        else if (!strcmp(a, "--save-weight-cache") && v) cache_out = v, ++i;
        else if (!strcmp(a, "--result") && v) result_path = v, ++i;
        else if (!strcmp(a, "--answer") && v) answer_path = v, ++i;
        else if (!strcmp(a, "--n") && v) n_limit = atoi(v), ++i;
        else if (!strcmp(a, "--gpus") && v) cfg.n_gpus = atoi(v), ++i;
        else if (!strcmp(a, "--max-batch") && v) cfg.max_batch_per_gpu = atoi(v), ++i;
        else if (!strcmp(a, "--synthetic") && v) synthetic = atoi(v), ++i;
        else if (!strcmp(a, "--img") && v) img = atoi(v), ++i;
        else if (!strcmp(a, "--precision") && v) cfg.precision = strcmp(v, "fp16") ? VIT_PREC_BF16 : VIT_PREC_FP16, ++i;
        else {
            fprintf(stderr, "usage: %s (--images FILE --weights DIR|CACHEFILE | --synthetic N [--img S]) [--n N] [--gpus G] "
                            "[--max-batch B] [--precision bf16|fp16] [--result FILE] [--answer FILE] [--save-weight-cache FILE]\n", argv[0]);
            return 2;
        }
    }
    if (!synthetic && (!images_path || !weights_dir)) {
        fprintf(stderr, "need --images and --weights, or --synthetic N\n");
        return 2;
    }
    vit_host_set_config(&cfg);
    if (initialize_cuda() != 0) return 1;

    ImageData* images = NULL;
    static Network network[VIT_NUM_TENSORS];
    if (synthetic) {
        const size_t per = (size_t)3 * img * img;
        images = (ImageData*)calloc((size_t)synthetic, sizeof(ImageData));
        for (int i = 0; i < synthetic; ++i) {
            images[i].n = synthetic; images[i].c = 3; images[i].h = img; images[i].w = img;
            images[i].data = (float*)malloc(per * sizeof(float));
            vit_synth_images(images[i].data, 1, img, 7, i);
        }
        if (vit_synth_weights(network, VIT_NUM_TENSORS, img, 42) != 0) return 1;
    } else {
        images = load_image_data(images_path);
        if (!images) return 1;
        img = images[0].h;
        struct stat sb;
        if (stat(weights_dir, &sb) == 0 && S_ISREG(sb.st_mode)) {   /* a weight cache written by --save-weight-cache */
            int cached_img = 0;
            if (load_weights_blob(weights_dir, network, VIT_NUM_TENSORS, &cached_img) != 0) return 1;
            printf("loaded %d weight tensors (img_size %d) from cache %s\n", VIT_NUM_TENSORS, cached_img, weights_dir);
        } else {
            const int loaded = load_weights(weights_dir, network, VIT_NUM_TENSORS);
            if (loaded < 0) return 1;
            printf("loaded %d / %d weight tensors from %s\n", loaded, VIT_NUM_TENSORS, weights_dir);
        }
    }
    if (vit_validate_weights(network, VIT_NUM_TENSORS, img) != 0) return 1;
    if (cache_out) {
        if (save_weights_blob(cache_out, network, VIT_NUM_TENSORS, img) != 0) return 1;
        printf("wrote weight cache %s\n", cache_out);
    }

    int n = images[0].n;
    if (n_limit > 0 && n_limit < n) {
        n = n_limit;
        images[0].n = n; /* the callee reads the loop bound here, as in Main.c:45-46 */
    }
    float** probabilities = (float**)malloc(sizeof(float*) * (size_t)n);
    for (int i = 0; i < n; ++i) probabilities[i] = (float*)malloc(sizeof(float) * VIT_NUM_CLASSES);

    printf("=====================Start========================\n");
    double t0 = now_s();
    ViT_cuda(images, network, probabilities); /* first call uploads the weights */
    double t1 = now_s();
    if (ViT_cuda_status() != 0) return 1;
    ViT_cuda(images, network, probabilities);
    double t2 = now_s();
    printf("CUDA time: %f sec (first call incl. weight upload), %f sec (steady), %d images, %.1f images/s\n",
           t1 - t0, t2 - t1, n, n / (t2 - t1));

    if (write_results(result_path, probabilities, n) != 0) return 1;
    int rc = 0;
    if (answer_path) {
        const int cmp = comparator_files(result_path, answer_path, n);
        if (cmp == 0) printf("Comparator: the two files match on %d lines.\n", n);
        else printf("Comparator: %d differences.\n", cmp);
        rc = cmp == 0 ? 0 : 3;
    }
    Release_cuda();
    return rc;
}
