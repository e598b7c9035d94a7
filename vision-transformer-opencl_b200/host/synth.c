/*
 * synth.c -- seeded synthetic weights and images.  The reference mount lacks
 * Data/input-100.bin and 36 of the 152 weight tensors and the GPU box has
 * no reference tree at all, so parity tests and the bench run on tensors generated here.
 * Counter based (value i depends only on seed, stream, i) and built from integer hashing
 * plus exactly-representable float arithmetic, hence bit-identical on every machine.
 */
#include "vit_host.h"

#include <math.h>
#include <stdlib.h>

static inline uint64_t mix64(uint64_t x) { /* splitmix64 finaliser */
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline float unit_normal(uint64_t seed, uint64_t stream, uint64_t i) {
    const uint64_t h = mix64(mix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + i);
    /* Irwin-Hall(4): sum of four 16-bit uniforms, variance 4/12 -> scale by sqrt(3) */
    const uint32_t s = (uint32_t)(h & 0xFFFF) + (uint32_t)((h >> 16) & 0xFFFF) +
                       (uint32_t)((h >> 32) & 0xFFFF) + (uint32_t)(h >> 48);
    return ((float)s - 131070.0f) * (1.7320508f / 65536.0f);
}

void vit_synth_fill(float* dst, size_t n, uint64_t seed, uint64_t stream,
                    float mean, float sigma, float lo, float hi) {
    for (size_t i = 0; i < n; ++i) {
        float v = mean + sigma * unit_normal(seed, stream, i);
        dst[i] = v < lo ? lo : (v > hi ? hi : v);
    }
}

int vit_synth_weights(Network network[], int count, int img_size, uint64_t seed) {
    if (count != VIT_NUM_TENSORS) return -1;
    for (int i = 0; i < count; ++i) {
        const size_t n = vit_tensor_numel(i, img_size);
        float* d = (float*)malloc(n * sizeof(float));
        if (!d) return -1;
        /* scales near the shipped tensors' (SURVEY.md App. A): conv 0.0092, out_proj 0.0196,
         * head 0.0374, pos_emb 0.0508, class_token 0.0161; 0.03 for the three absent kinds */
        float mean = 0.0f, sigma = 0.02f;
        if (i == 0) sigma = 0.0161f;
        else if (i == 1) sigma = 0.0092f;
        else if (i == 2) sigma = 0.05f;
        else if (i == 3) sigma = 0.0508f;
        else if (i == 148) { mean = 1.0f; sigma = 0.1f; }
        else if (i == 149) sigma = 0.05f;
        else if (i == 150) sigma = 0.0374f;
        else if (i == 151) sigma = 0.02f;
        else switch ((i - 4) % 12) {
            case 0: case 6: mean = 1.0f; sigma = 0.1f; break;   /* ln weights */
            case 1: case 7: sigma = 0.05f; break;               /* ln biases */
            case 2: sigma = 0.03f; break;                       /* in_proj_weight */
            case 4: sigma = 0.0196f; break;                     /* out_proj.weight */
            case 8: case 10: sigma = 0.03f; break;              /* mlp weights */
            default: sigma = 0.02f;                             /* biases */
        }
        vit_synth_fill(d, n, seed, (uint64_t)i + 1, mean, sigma, -1e30f, 1e30f);
        for (size_t k = 0; k < n; ++k) d[k] = roundf(d[k] * 1000000.0f) / 1000000.0f;
        network[i].data = d;
        network[i].size = n;
    }
    return 0;
}

void vit_synth_images(float* nchw, int n, int img_size, uint64_t seed, int first_index) {
    const size_t per = (size_t)3 * img_size * img_size;
    for (int i = 0; i < n; ++i)
        vit_synth_fill(nchw + (size_t)i * per, per, seed, 0x100000ull + (uint64_t)(first_index + i),
                       0.0f, 1.0f, -2.12f, 2.64f);
}
