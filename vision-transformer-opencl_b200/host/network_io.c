/*
 * network_io.c -- POSIX implementation of the reference's asset loaders
 * (load_image_data: Network.c:24-97, load_weights: Network.c:119-194) plus writers and
 * shape validation the reference lacks.  Plain C, no CUDA.
 */
#include "vit_host.h"

#include <dirent.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#define DIM 768
#define HID 3072
#define PATCH 16
#define CHANS 3

ImageData* load_image_data(const char* filename) {
    FILE* f = fopen(filename, "rb");
    if (!f) {
        fprintf(stderr, "load_image_data: cannot open %s: %s\n", filename, strerror(errno));
        return NULL;
    }
    int32_t header[4];
    if (fread(header, sizeof(int32_t), 4, f) != 4) {
        fprintf(stderr, "load_image_data: %s: short header\n", filename);
        fclose(f);
        return NULL;
    }
    const int n = header[0], c = header[1], h = header[2], w = header[3];
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0) {
        fprintf(stderr, "load_image_data: %s: bad header %d,%d,%d,%d\n", filename, n, c, h, w);
        fclose(f);
        return NULL;
    }
    const size_t image_size = (size_t)c * h * w;
    ImageData* images = (ImageData*)calloc((size_t)n, sizeof(ImageData));
    if (!images) {
        fclose(f);
        return NULL;
    }
    for (int i = 0; i < n; ++i) {
        /* every descriptor carries the batch size; each image owns its buffer */
        images[i].n = n;
        images[i].c = c;
        images[i].h = h;
        images[i].w = w;
        images[i].data = (float*)malloc(image_size * sizeof(float));
        if (!images[i].data || fread(images[i].data, sizeof(float), image_size, f) != image_size) {
            fprintf(stderr, "load_image_data: %s: image %d unreadable\n", filename, i);
            fclose(f);
            free_image_data(images);
            return NULL;
        }
    }
    fclose(f);
    return images;
}

/* ---- streaming reader for the same file format: bounded memory, caller-provided (e.g. pinned) buffers ---- */
struct vit_image_stream {
    FILE* f;
    int n, c, h, w, next;
};

vit_image_stream* vit_image_stream_open(const char* filename, int* n, int* c, int* h, int* w) {
    FILE* f = fopen(filename, "rb");
    if (!f) {
        fprintf(stderr, "vit_image_stream_open: cannot open %s: %s\n", filename, strerror(errno));
        return NULL;
    }
    int32_t header[4];
    if (fread(header, sizeof(int32_t), 4, f) != 4 || header[0] <= 0 || header[1] <= 0 || header[2] <= 0 || header[3] <= 0) {
        fprintf(stderr, "vit_image_stream_open: %s: bad header\n", filename);
        fclose(f);
        return NULL;
    }
    /* the header's image count must be backed by the file (the reference trusts it, Network.c:60-64) */
    struct stat sb;
    const size_t need = 16 + (size_t)header[0] * header[1] * header[2] * header[3] * sizeof(float);
    if (fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode) && (size_t)sb.st_size < need) {
        fprintf(stderr, "vit_image_stream_open: %s: %lld bytes, header promises %zu\n", filename, (long long)sb.st_size, need);
        fclose(f);
        return NULL;
    }
    vit_image_stream* s = (vit_image_stream*)calloc(1, sizeof(*s));
    if (!s) {
        fclose(f);
        return NULL;
    }
    s->f = f;
    s->n = header[0], s->c = header[1], s->h = header[2], s->w = header[3];
    if (n) *n = s->n;
    if (c) *c = s->c;
    if (h) *h = s->h;
    if (w) *w = s->w;
    return s;
}

int vit_image_stream_read(vit_image_stream* s, float* dst, int max_images) {
    if (!s || !dst || max_images < 0) return -1;
    const int left = s->n - s->next;
    const int take = left < max_images ? left : max_images;
    const size_t per = (size_t)s->c * s->h * s->w;
    if (take > 0 && fread(dst, sizeof(float), per * take, s->f) != per * take) {
        fprintf(stderr, "vit_image_stream_read: short read at image %d\n", s->next);
        return -1;
    }
    s->next += take;
    return take;
}

void vit_image_stream_close(vit_image_stream* s) {
    if (!s) return;
    fclose(s->f);
    free(s);
}

void free_image_data(ImageData* images) {
    if (!images) return;
    const int n = images[0].n;
    for (int i = 0; i < n; ++i) free(images[i].data);
    free(images);
}

int save_image_data(const char* filename, const float* nchw, int n, int c, int h, int w) {
    FILE* f = fopen(filename, "wb");
    if (!f) return -1;
    int32_t header[4] = {n, c, h, w};
    const size_t total = (size_t)n * c * h * w;
    int ok = fwrite(header, sizeof(int32_t), 4, f) == 4 && fwrite(nchw, sizeof(float), total, f) == total;
    fclose(f);
    return ok ? 0 : -1;
}

/* digits between "Weight_" and the next '_' (Network.c:99-117) */
static int index_from_filename(const char* name) {
    if (strncmp(name, "Weight_", 7) != 0) return -1;
    const char* start = name + 7;
    const char* end = strchr(start, '_');
    if (!end || end == start || end - start > 9) return -1;
    int idx = 0;
    for (const char* p = start; p < end; ++p) {
        if (*p < '0' || *p > '9') return -1;
        idx = idx * 10 + (*p - '0');
    }
    return idx;
}

int load_weights(const char* directory, Network network[], int count) {
    DIR* dir = opendir(directory);
    if (!dir) {
        fprintf(stderr, "load_weights: cannot open %s: %s\n", directory, strerror(errno));
        return -1;
    }
    for (int i = 0; i < count; ++i) {
        network[i].data = NULL;
        network[i].size = 0;
    }
    int loaded = 0;
    struct dirent* entry;
    while ((entry = readdir(dir)) != NULL) {
        const char* ext = strrchr(entry->d_name, '.');
        if (!ext || strcmp(ext, ".bin") != 0) continue;
        const int idx = index_from_filename(entry->d_name);
        if (idx < 0 || idx >= count) continue;
        char path[1024];
        snprintf(path, sizeof(path), "%s/%s", directory, entry->d_name);
        FILE* fp = fopen(path, "rb");
        if (!fp) { /* the reference dereferences NULL here (Network.c:151-158) */
            fprintf(stderr, "load_weights: cannot open %s: %s\n", path, strerror(errno));
            continue;
        }
        fseek(fp, 0, SEEK_END);
        const long file_size = ftell(fp);
        rewind(fp);
        if (file_size <= 0) {
            fclose(fp);
            continue;
        }
        const size_t num_floats = (size_t)file_size / sizeof(float);
        float* buffer = (float*)malloc(num_floats * sizeof(float));
        if (!buffer || fread(buffer, sizeof(float), num_floats, fp) != num_floats) {
            fprintf(stderr, "load_weights: %s unreadable\n", path);
            free(buffer);
            fclose(fp);
            continue;
        }
        fclose(fp);
        /* the values the model sees are rounded to 6 decimals, in fp32 (Network.c:185-187) */
        for (size_t i = 0; i < num_floats; ++i) buffer[i] = roundf(buffer[i] * 1000000.0f) / 1000000.0f;
        if (network[idx].data) free(network[idx].data); else ++loaded;
        network[idx].data = buffer;
        network[idx].size = num_floats;
    }
    closedir(dir);
    return loaded;
}

void free_weights(Network network[], int count) {
    for (int i = 0; i < count; ++i) {
        free(network[i].data);
        network[i].data = NULL;
        network[i].size = 0;
    }
}

size_t vit_tensor_numel(int idx, int img_size) {
    const size_t g = (size_t)(img_size / PATCH), tokens = g * g + 1;
    if (idx < 0 || idx >= VIT_NUM_TENSORS) return 0;
    switch (idx) {
        case 0: return DIM;
        case 1: return (size_t)DIM * CHANS * PATCH * PATCH;
        case 2: return DIM;
        case 3: return tokens * DIM;
        case 148: case 149: return DIM;
        case 150: return (size_t)VIT_NUM_CLASSES * DIM;
        case 151: return VIT_NUM_CLASSES;
    }
    switch ((idx - 4) % 12) {
        case 2: return (size_t)3 * DIM * DIM;
        case 3: return 3 * DIM;
        case 4: return (size_t)DIM * DIM;
        case 8: return (size_t)HID * DIM;
        case 9: return HID;
        case 10: return (size_t)DIM * HID;
        default: return DIM; /* ln_1/ln_2 weight+bias, out_proj bias, mlp_3 bias */
    }
}

const char* vit_tensor_name(int idx, char* buf, size_t buflen) {
    static const char* const per_layer[12] = {
        "ln_1_weight", "ln_1_bias", "self_attention_in_proj_weight", "self_attention_in_proj_bias",
        "self_attention_out_proj_weight", "self_attention_out_proj_bias", "ln_2_weight", "ln_2_bias",
        "mlp_0_weight", "mlp_0_bias", "mlp_3_weight", "mlp_3_bias"};
    switch (idx) {
        case 0: snprintf(buf, buflen, "class_token"); break;
        case 1: snprintf(buf, buflen, "conv_proj_weight"); break;
        case 2: snprintf(buf, buflen, "conv_proj_bias"); break;
        case 3: snprintf(buf, buflen, "encoder_pos_embedding"); break;
        case 148: snprintf(buf, buflen, "encoder_ln_weight"); break;
        case 149: snprintf(buf, buflen, "encoder_ln_bias"); break;
        case 150: snprintf(buf, buflen, "heads_head_weight"); break;
        case 151: snprintf(buf, buflen, "heads_head_bias"); break;
        default:
            if (idx < 0 || idx >= VIT_NUM_TENSORS) snprintf(buf, buflen, "invalid");
            else snprintf(buf, buflen, "encoder_layers_encoder_layer_%d_%s", (idx - 4) / 12, per_layer[(idx - 4) % 12]);
    }
    return buf;
}

int vit_validate_weights(const Network network[], int count, int img_size) {
    if (count != VIT_NUM_TENSORS) {
        fprintf(stderr, "vit_validate_weights: expected %d tensors, got %d\n", VIT_NUM_TENSORS, count);
        return -1;
    }
    for (int i = 0; i < count; ++i) {
        const size_t want = vit_tensor_numel(i, img_size);
        if (!network[i].data || network[i].size != want) {
            char name[128];
            fprintf(stderr, "vit_validate_weights: tensor %d (%s): have %zu floats%s, need %zu\n", i,
                    vit_tensor_name(i, name, sizeof(name)), network[i].size,
                    network[i].data ? "" : " (missing)", want);
            return -(i + 1);
        }
    }
    return 0;
}

int save_weights(const char* directory, const Network network[], int count, int img_size) {
    (void)img_size;
    mkdir(directory, 0777);
    for (int i = 0; i < count; ++i) {
        if (!network[i].data) continue;
        char name[128], path[1024];
        snprintf(path, sizeof(path), "%s/Weight_%d_%s.bin", directory, i, vit_tensor_name(i, name, sizeof(name)));
        FILE* f = fopen(path, "wb");
        if (!f) return -1;
        const int ok = fwrite(network[i].data, sizeof(float), network[i].size, f) == network[i].size;
        fclose(f);
        if (!ok) return -1;
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Weight cache: the 152 tensors AFTER load_weights' 1e-6 rounding in ONE file,
 * so a restart is a single sequential read instead of 152 opens + a rounding pass over 330 MB
 * (Network.c:119-194).  Layout, little endian:
 *     char magic[8] = "VITW0001";  uint32 img_size, count;  uint64 numel[count];
 *     float data[sum numel];       uint64 fnv1a64(over everything before it)
 * load_weights_blob verifies magic, tensor sizes for img_size and the checksum before handing
 * anything out.
 */
static uint64_t fnv1a64(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) {
        h ^= b[i];
        h *= 0x100000001b3ull;
    }
    return h;
}

int save_weights_blob(const char* path, const Network network[], int count, int img_size) {
    if (!path || !network || count != VIT_NUM_TENSORS) return -1;
    if (vit_validate_weights(network, count, img_size) != 0) return -1;
    FILE* f = fopen(path, "wb");
    if (!f) {
        fprintf(stderr, "save_weights_blob: cannot create %s: %s\n", path, strerror(errno));
        return -1;
    }
    uint64_t h = 0xcbf29ce484222325ull;
    const char magic[8] = {'V', 'I', 'T', 'W', '0', '0', '0', '1'};
    const uint32_t head[2] = {(uint32_t)img_size, (uint32_t)count};
    int ok = fwrite(magic, 1, 8, f) == 8 && fwrite(head, sizeof(uint32_t), 2, f) == 2;
    h = fnv1a64(fnv1a64(h, magic, 8), head, sizeof(head));
    for (int i = 0; ok && i < count; ++i) {
        const uint64_t n = network[i].size;
        ok = fwrite(&n, sizeof(n), 1, f) == 1;
        h = fnv1a64(h, &n, sizeof(n));
    }
    for (int i = 0; ok && i < count; ++i) {
        ok = fwrite(network[i].data, sizeof(float), network[i].size, f) == network[i].size;
        h = fnv1a64(h, network[i].data, network[i].size * sizeof(float));
    }
    ok = ok && fwrite(&h, sizeof(h), 1, f) == 1;
    if (fclose(f) != 0) ok = 0;
    if (!ok) fprintf(stderr, "save_weights_blob: short write to %s\n", path);
    return ok ? 0 : -1;
}

int load_weights_blob(const char* path, Network network[], int count, int* img_size_out) {
    if (!path || !network || count != VIT_NUM_TENSORS) return -1;
    for (int i = 0; i < count; ++i) {
        network[i].data = NULL;
        network[i].size = 0;
    }
    FILE* f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "load_weights_blob: cannot open %s: %s\n", path, strerror(errno));
        return -1;
    }
    char magic[8];
    uint32_t head[2];
    uint64_t h = 0xcbf29ce484222325ull, stored = 0;
    int ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "VITW0001", 8) == 0 && fread(head, sizeof(uint32_t), 2, f) == 2 &&
             head[1] == (uint32_t)count && head[0] % PATCH == 0 && head[0] > 0 && head[0] <= 1024;
    if (!ok) fprintf(stderr, "load_weights_blob: %s is not a VITW0001 weight cache for %d tensors\n", path, count);
    if (ok) h = fnv1a64(fnv1a64(h, magic, 8), head, sizeof(head));
    for (int i = 0; ok && i < count; ++i) {
        uint64_t n = 0;
        ok = fread(&n, sizeof(n), 1, f) == 1 && n == vit_tensor_numel(i, (int)head[0]);
        if (!ok) fprintf(stderr, "load_weights_blob: %s: tensor %d has the wrong size for img_size %u\n", path, i, head[0]);
        h = fnv1a64(h, &n, sizeof(n));
        network[i].size = (size_t)n;
    }
    for (int i = 0; ok && i < count; ++i) {
        network[i].data = (float*)malloc(network[i].size * sizeof(float));
        ok = network[i].data && fread(network[i].data, sizeof(float), network[i].size, f) == network[i].size;
        if (!ok) fprintf(stderr, "load_weights_blob: %s: tensor %d unreadable\n", path, i);
        else h = fnv1a64(h, network[i].data, network[i].size * sizeof(float));
    }
    if (ok) {
        ok = fread(&stored, sizeof(stored), 1, f) == 1 && stored == h;
        if (!ok) fprintf(stderr, "load_weights_blob: %s: checksum mismatch (file damaged or truncated)\n", path);
    }
    fclose(f);
    if (!ok) {
        free_weights(network, count);
        return -1;
    }
    if (img_size_out) *img_size_out = (int)head[0];
    return 0;
}
