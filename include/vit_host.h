/*
 * vit_host.h -- plain-C host side: the reference's loader / driver contract on POSIX, plus
 * the reference-signature adaptor over the CUDA engine (vit_cuda.h).
 *
 * Mirrors, name for name where the reference has one:
 *     ImageData, load_image_data()          Network.h:7-15,  Network.c:24-97
 *     Network,   load_weights()             Network.h:18-21,34, Network.c:119-194
 *     ViT_opencl()  -> ViT_cuda()           ViT_opencl.h:18, ViT_opencl.c:785-883
 *     initialize_opencl() -> initialize_cuda(), Release_opencl() -> Release_cuda()
 *     result line format                    Main.c:62-72
 *     comparator()                          comparator.c:23-80 (IMAGE_COUNT becomes an argument)
 * A maintainer of the reference swaps `#include "ViT_opencl.h"` for this header and
 * ViT_opencl(...) for ViT_cuda(...) in Main.c; see INTEGRATION.md.
 */
#ifndef VIT_HOST_H
#define VIT_HOST_H

#include <stddef.h>
#include <stdint.h>
#include "vit_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One image descriptor; load_image_data returns an array of n of them, each with its own
 * malloc'd [c][h][w] fp32 buffer, and images[0].n is the loop bound (Network.c:66-93,
 * ViT_seq.c:354). */
typedef struct {
    int n, c, h, w;
    float* data;
} ImageData;

typedef vit_tensor Network;

/* File format (Network.c:36-58): int32 n,c,h,w then n*c*h*w fp32, NCHW, native endian.
 * Returns NULL (with a message on stderr) on any I/O or format error. */
ImageData* load_image_data(const char* filename);
void       free_image_data(ImageData* images);

/* Streaming form of load_image_data for files that should not be held in memory as a whole (the reference reads the
 * complete file into n separately malloc'd images, Network.c:66-93): open validates the header against the file size,
 * read copies up to max_images further images, contiguous NCHW, into a caller-provided buffer (pinned memory makes
 * the engine's host-to-device copy direct) and returns how many it delivered (0 at the end, -1 on error). */
typedef struct vit_image_stream vit_image_stream;
vit_image_stream* vit_image_stream_open(const char* filename, int* n, int* c, int* h, int* w);
int  vit_image_stream_read(vit_image_stream* s, float* dst, int max_images);
void vit_image_stream_close(vit_image_stream* s);

/* Scans `directory` for Weight_<idx>_*.bin, idx in [0,count); reads each as raw fp32 and
 * rounds every value to 6 decimals exactly as Network.c:185-187 does.  Unlike the
 * reference it does not exit(): returns the number of tensors loaded, or -1 if the
 * directory cannot be opened.  Missing slots are left {NULL,0} (as in the reference). */
int  load_weights(const char* directory, Network network[], int count);
void free_weights(Network network[], int count);

/* Expected float count of tensor idx for img_size (224 or 384); 0 for a bad index. */
size_t vit_tensor_numel(int idx, int img_size);
/* state_dict name of tensor idx with '.' -> '_' (the Weight_<idx>_<name>.bin stem part). */
const char* vit_tensor_name(int idx, char* buf, size_t buflen);
/* 0 if all `count` tensors are present with the size the model needs, else -(idx+1) of the
 * first offender (message on stderr). */
int vit_validate_weights(const Network network[], int count, int img_size);

/* Writers for the two file formats (used by the synthetic-asset tools and the tests). */
int save_image_data(const char* filename, const float* nchw, int n, int c, int h, int w);
int save_weights(const char* directory, const Network network[], int count, int img_size);
/* Weight cache: all tensors (as load_weights returns them, i.e. after its 1e-6 rounding) in ONE
 * checksummed file, so a restart reads 330 MB sequentially instead of opening 152 files and
 * re-rounding them (Network.c:119-194).  load_weights_blob validates magic, the tensor sizes for
 * the stored image size and the checksum; on any failure nothing is returned (-1).  0 = ok. */
int save_weights_blob(const char* path, const Network network[], int count, int img_size);
int load_weights_blob(const char* path, Network network[], int count, int* img_size_out);

/* ---- engine lifecycle with the reference's shape --------------------------------------- */

/* Configuration picked up by initialize_cuda()/ViT_cuda(); also settable through the
 * environment: VIT_GPUS, VIT_MAX_BATCH, VIT_PRECISION=auto|fp16|bf16. */
typedef struct {
    int n_gpus;             /* default 1 */
    int max_batch_per_gpu;  /* default 256 */
    int precision;          /* VIT_PREC_*, default VIT_PREC_AUTO */
} vit_host_config;
void vit_host_set_config(const vit_host_config* cfg);

/* Replaces initialize_opencl(): selects the device(s); the weights are uploaded on the first
 * ViT_cuda() call (the reference signature only hands them over there).  Returns 0 / <0. */
int  initialize_cuda(void);
/* Replaces ViT_opencl(): image[0].n images -> prb[i][0..999] softmax probabilities
 * (host Softmax of the engine's logits, ViT_seq.c:304-324).  On engine failure prints
 * vit_cuda_last_error() to stderr and fills prb with NaN; never calls exit(). */
void ViT_cuda(ImageData* image, Network* networks, float** prb);
/* Status of the last ViT_cuda() call (0 ok). */
int  ViT_cuda_status(void);
/* ViT_cuda() uploads the weights when it first sees them and again whenever the Network array's address, any tensor's
 * data pointer or size, or a sample of 64 elements per tensor has changed since (the reference's usage is ONE static
 * Network[152] array, so reloading other weights into it keeps the address).  A caller that edits weights in place
 * in a way the sample may miss calls this to force the next ViT_cuda() to upload again. */
void vit_host_invalidate_weights(void);
/* Replaces Release_opencl(). */
void Release_cuda(void);

/* ---- results ------------------------------------------------------------------------------ */

void vit_softmax(const float* logits, float* probs, int length);          /* ViT_seq.c:304-324 */
/* argmax with a fresh scan per image, lowest index wins ties (the reference carries pred_idx
 * across images, Main.c:62 -- a quirk, not replicated; identical whenever n == 1 or the
 * top-1 is not class 0). */
int  vit_argmax(const float* v, int length);
/* Writes "[%d] label: %d / prob: %.6f\n" per image (Main.c:71).  Returns 0 / -1. */
int  write_results(const char* filename, float* const* prb, int n);
/* comparator.c:23-80 with explicit paths and line count: label must match, |dprob| <= 0.01f;
 * short file -> +1 and stop; unparsable line -> +1; unopenable file -> returns 1. */
int  comparator_files(const char* result_path, const char* answer_path, int image_count);

/* ---- seeded synthetic assets (counter based, bit-identical on every machine) ------------- */

/* dst[i] = clamp(mean + sigma * z(seed, stream, i), lo, hi) with z ~ approx N(0,1)
 * (sum of four 16-bit uniforms, exactly representable; |z| <= 3.47). */
void vit_synth_fill(float* dst, size_t n, uint64_t seed, uint64_t stream,
                    float mean, float sigma, float lo, float hi);
/* Fills all `count` tensors of a ViT-B/16 for img_size with trained-model-like scales
 * (SURVEY.md 8(d) config 2), mallocs the buffers, applies the loader's 1e-6 rounding. */
int  vit_synth_weights(Network network[], int count, int img_size, uint64_t seed);
/* [n][3][S][S] images ~ N(0,1) clipped to the ImageNet-normalised range [-2.12, 2.64]. */
void vit_synth_images(float* nchw, int n, int img_size, uint64_t seed, int first_index);

#ifdef __cplusplus
}
#endif
#endif /* VIT_HOST_H */
