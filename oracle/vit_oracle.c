/*
 * vit_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see vit_oracle.h).
 *
 * A restatement of the reference's sequential fp32 ViT-B/16 forward
 * (reference ViT_seq.c).  Every output element is produced by the same sequence of
 * fp32 operations as in the reference (same summation order, bias first, separate
 * multiply and add -- build with -ffp-contract=off and without -ffast-math), so the
 * results are bit-identical to ViT_seq() compiled with gcc -O2 on x86-64.
 * Speed comes only from computing many *independent* outputs side by side (SIMD
 * across output features / tokens, OpenMP across tokens, heads and images), which
 * does not change any individual result.
 */
#include "vit_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define D   VIT_O_DIM
#define NH  VIT_O_HEADS
#define DH  VIT_O_HEAD_DIM
#define HID VIT_O_HIDDEN

#if defined(__x86_64__) && defined(__GNUC__) && !defined(VIT_ORACLE_NO_CLONES)
#define ORACLE_CLONES __attribute__((target_clones("default", "avx2", "avx512f")))
#else
#define ORACLE_CLONES
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- linear ----------------------------------------------------------------------
 * ViT_seq.c:240-250:  sum = bias[o]; for i ascending: sum += in[t*in+i] * W[o*in+i].
 * Here 4 tokens x 16 output features are accumulated side by side from a transposed
 * copy of W; each accumulator sees exactly the reference's operation sequence. */
#define LIN_TB 4
#define LIN_OB 16

ORACLE_CLONES
static void linear_block(const float* x, float* y, int t0, int tn, int in_f, int out_f,
                         const float* Wt /* [in_f][out_pad] */, int out_pad,
                         const float* bp /* [out_pad], zero padded */) {
    /* rows past the end of a ragged last block alias the last valid row; their results
     * are computed and dropped, so no load is ever conditional */
    const float* xr[LIN_TB];
    for (int tt = 0; tt < LIN_TB; ++tt) xr[tt] = x + (size_t)(t0 + (tt < tn ? tt : tn - 1)) * in_f;
    for (int o0 = 0; o0 < out_pad; o0 += LIN_OB) {
        float acc[LIN_TB][LIN_OB];
        for (int tt = 0; tt < LIN_TB; ++tt)
            for (int j = 0; j < LIN_OB; ++j) acc[tt][j] = bp[o0 + j];
        for (int i = 0; i < in_f; ++i) {
            const float* wrow = Wt + (size_t)i * out_pad + o0;
            for (int tt = 0; tt < LIN_TB; ++tt) {
                const float xv = xr[tt][i];
                for (int j = 0; j < LIN_OB; ++j) {
                    float p = xv * wrow[j];
                    acc[tt][j] = acc[tt][j] + p;
                }
            }
        }
        const int jn = out_f - o0 < LIN_OB ? out_f - o0 : LIN_OB;
        for (int tt = 0; tt < tn; ++tt)
            for (int j = 0; j < jn; ++j) y[(size_t)(t0 + tt) * out_f + o0 + j] = acc[tt][j];
    }
}

void oracle_linear(const float* x, float* y, int tokens, int in_f, int out_f,
                   const float* W, const float* b) {
    if (tokens < LIN_TB) { /* not worth a transpose: plain reference loop */
        for (int t = 0; t < tokens; ++t)
            for (int o = 0; o < out_f; ++o) {
                float sum = b[o];
                const float* w = W + (size_t)o * in_f;
                const float* xi = x + (size_t)t * in_f;
                for (int i = 0; i < in_f; ++i) {
                    float p = xi[i] * w[i];
                    sum = sum + p;
                }
                y[(size_t)t * out_f + o] = sum;
            }
        return;
    }
    const int out_pad = (out_f + LIN_OB - 1) / LIN_OB * LIN_OB;
    float* Wt = (float*)calloc((size_t)in_f * out_pad + out_pad, sizeof(float));
    float* bp = Wt + (size_t)in_f * out_pad;
    memcpy(bp, b, sizeof(float) * out_f);
    for (int o = 0; o < out_f; ++o)
        for (int i = 0; i < in_f; ++i) Wt[(size_t)i * out_pad + o] = W[(size_t)o * in_f + i];
    const int nblk = (tokens + LIN_TB - 1) / LIN_TB;
#pragma omp parallel for schedule(static)
    for (int blk = 0; blk < nblk; ++blk) {
        int t0 = blk * LIN_TB;
        int tn = tokens - t0 < LIN_TB ? tokens - t0 : LIN_TB;
        linear_block(x, y, t0, tn, in_f, out_f, Wt, out_pad, bp);
    }
    free(Wt);
}

/* ---- layer norm --------------------------------------------------------------------
 * ViT_seq.c:103-121.  float sum/sum_sq accumulated in index order; mean = sum/768;
 * var = sum_sq/768 - mean*mean; inv_std = 1.0f / sqrtf(var + 1e-6) where the eps add
 * is done in double (eps is a double literal) and narrowed by sqrtf's float parameter;
 * y = (x - mean) * inv_std * w + b evaluated left to right. */
void oracle_layer_norm(const float* x, float* y, int tokens, const float* w, const float* b) {
    const double eps = 1e-6;
    for (int t = 0; t < tokens; ++t) {
        float sum = 0.0f, sum_sq = 0.0f;
        for (int i = 0; i < D; ++i) {
            float val = x[(size_t)t * D + i];
            sum = sum + val;
            float sq = val * val;
            sum_sq = sum_sq + sq;
        }
        float mean = sum / D;
        float msq = mean * mean;
        float var = sum_sq / D - msq;
        float inv_std = 1.0f / sqrtf((float)((double)var + eps));
        for (int i = 0; i < D; ++i) {
            size_t idx = (size_t)t * D + i;
            float c = x[idx] - mean;
            float n = c * inv_std;
            float s = n * w[i];
            y[idx] = s + b[i];
        }
    }
}

/* ---- GELU (ViT_seq.c:231-233) ------------------------------------------------------- */
float oracle_gelu(float x) { return 0.5f * x * (1.0f + erff(x / sqrtf(2.0f))); }

void oracle_gelu_inplace(float* x, size_t n) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) x[i] = oracle_gelu(x[i]);
}

/* ---- attention core (ViT_seq.c:156-215) ----------------------------------------------
 * per head: score[i][j] = (sum_d q*k, d ascending from 0.0f) / sqrtf(64);
 * row softmax with max subtraction, expf, ascending sum, division;
 * out[i][d] = sum_j p[i][j] * V[j][d], j ascending from 0.0f. */
ORACLE_CLONES
static void attention_head(const float* Q, const float* K, const float* V, float* out,
                           int tokens, int h, float* scores /* [tokens] scratch */,
                           float* Kt /* [DH][tokens] scratch */) {
    const int off = h * DH;
    const float scale = sqrtf((float)DH);
    for (int j = 0; j < tokens; ++j)
        for (int d = 0; d < DH; ++d) Kt[(size_t)d * tokens + j] = K[(size_t)j * D + off + d];
    for (int i = 0; i < tokens; ++i) {
        const float* q = Q + (size_t)i * D + off;
        for (int j = 0; j < tokens; ++j) scores[j] = 0.0f;
        for (int d = 0; d < DH; ++d) {
            const float qd = q[d];
            const float* kr = Kt + (size_t)d * tokens;
            for (int j = 0; j < tokens; ++j) {
                float p = qd * kr[j];
                scores[j] = scores[j] + p;
            }
        }
        for (int j = 0; j < tokens; ++j) scores[j] = scores[j] / scale;
        float max_val = scores[0];
        for (int j = 1; j < tokens; ++j)
            if (scores[j] > max_val) max_val = scores[j];
        float sum_exp = 0.0f;
        for (int j = 0; j < tokens; ++j) {
            scores[j] = expf(scores[j] - max_val);
            sum_exp = sum_exp + scores[j];
        }
        for (int j = 0; j < tokens; ++j) scores[j] = scores[j] / sum_exp;
        float acc[DH];
        for (int d = 0; d < DH; ++d) acc[d] = 0.0f;
        for (int j = 0; j < tokens; ++j) {
            const float pj = scores[j];
            const float* vr = V + (size_t)j * D + off;
            for (int d = 0; d < DH; ++d) {
                float p = pj * vr[d];
                acc[d] = acc[d] + p;
            }
        }
        for (int d = 0; d < DH; ++d) out[(size_t)i * D + off + d] = acc[d];
    }
}

void oracle_attention_core(const float* Q, const float* K, const float* V, float* out, int tokens) {
#pragma omp parallel
    {
        float* scores = (float*)malloc(sizeof(float) * tokens);
        float* Kt = (float*)malloc(sizeof(float) * tokens * DH);
#pragma omp for schedule(static)
        for (int h = 0; h < NH; ++h) attention_head(Q, K, V, out, tokens, h, scores, Kt);
        free(scores);
        free(Kt);
    }
}

/* ---- multi-head attention (ViT_seq.c:123-229) ------------------------------------------
 * Q/K/V rows of in_proj_weight are [0,768) / [768,1536) / [1536,2304); each of Q,K,V is
 * bias-first, j ascending -- i.e. three linears. */
void oracle_multihead_attn(const float* x, float* y, int tokens,
                           const float* in_w, const float* in_b,
                           const float* out_w, const float* out_b) {
    size_t n = (size_t)tokens * D;
    float* Q = (float*)malloc(sizeof(float) * n * 4);
    float *K = Q + n, *V = K + n, *A = V + n;
    oracle_linear(x, Q, tokens, D, D, in_w, in_b);
    oracle_linear(x, K, tokens, D, D, in_w + (size_t)D * D, in_b + D);
    oracle_linear(x, V, tokens, D, D, in_w + (size_t)2 * D * D, in_b + 2 * D);
    oracle_attention_core(Q, K, V, A, tokens);
    oracle_linear(A, y, tokens, D, D, out_w, out_b);
    free(Q);
}

/* ---- patch embedding (ViT_seq.c:25-101) ------------------------------------------------
 * Conv2d accumulates bias first, then (ic, kh, kw) ascending -- a linear over the
 * (ic,kh,kw)-flattened patch.  flatten_transpose puts patch oh*G+ow in row, class_token
 * prepends networks[0], pos_emb adds networks[3] elementwise. */
void oracle_embed(const float* image, float* out, int img_size,
                  const float* cls, const float* conv_w, const float* conv_b, const float* pos) {
    const int G = img_size / VIT_O_PATCH, P = G * G, KK = VIT_O_CHANS * VIT_O_PATCH * VIT_O_PATCH;
    float* patches = (float*)malloc(sizeof(float) * (size_t)P * KK);
    for (int oh = 0; oh < G; ++oh)
        for (int ow = 0; ow < G; ++ow) {
            float* dst = patches + (size_t)(oh * G + ow) * KK;
            for (int ic = 0; ic < VIT_O_CHANS; ++ic)
                for (int kh = 0; kh < VIT_O_PATCH; ++kh)
                    for (int kw = 0; kw < VIT_O_PATCH; ++kw)
                        dst[(ic * VIT_O_PATCH + kh) * VIT_O_PATCH + kw] =
                            image[((size_t)ic * img_size + oh * VIT_O_PATCH + kh) * img_size +
                                  ow * VIT_O_PATCH + kw];
        }
    oracle_linear(patches, out + D, P, KK, D, conv_w, conv_b);
    free(patches);
    for (int j = 0; j < D; ++j) out[j] = cls[j];
    const size_t total = (size_t)(P + 1) * D;
    for (size_t i = 0; i < total; ++i) out[i] = out[i] + pos[i];
}

/* ---- encoder block (ViT_seq.c:271-302) -------------------------------------------------- */
void oracle_encoder_block(const float* x, float* y, int tokens, const float* const* w) {
    size_t n = (size_t)tokens * D;
    float* ln = (float*)malloc(sizeof(float) * (n * 3 + (size_t)tokens * HID));
    float *attn = ln + n, *resid = attn + n, *hid = resid + n;
    oracle_layer_norm(x, ln, tokens, w[0], w[1]);
    oracle_multihead_attn(ln, attn, tokens, w[2], w[3], w[4], w[5]);
    for (size_t i = 0; i < n; ++i) resid[i] = x[i] + attn[i];
    oracle_layer_norm(resid, ln, tokens, w[6], w[7]);
    oracle_linear(ln, hid, tokens, D, HID, w[8], w[9]);            /* mlp_block, ViT_seq.c:251-268 */
    oracle_gelu_inplace(hid, (size_t)tokens * HID);
    oracle_linear(hid, attn, tokens, HID, D, w[10], w[11]);
    for (size_t i = 0; i < n; ++i) y[i] = resid[i] + attn[i];
    free(ln);
}

/* ---- softmax (ViT_seq.c:304-324) ---------------------------------------------------------- */
void oracle_softmax(const float* logits, float* probs, int n) {
    float max_val = logits[0];
    for (int i = 1; i < n; ++i)
        if (logits[i] > max_val) max_val = logits[i];
    float sum_exp = 0.0f;
    for (int i = 0; i < n; ++i) {
        probs[i] = expf(logits[i] - max_val);
        sum_exp = sum_exp + probs[i];
    }
    for (int i = 0; i < n; ++i) probs[i] = probs[i] / sum_exp;
}

/* ---- whole model (ViT_seq.c:337-439) --------------------------------------------------------
 * The reference runs the final layer_norm on all rows and keeps row 0; only row 0 is
 * computed here (rows are independent, so row 0 is identical). */
static void forward_one(const float* const* w, int img_size, const float* image,
                        float* logits, float* probs) {
    const int G = img_size / VIT_O_PATCH, T = G * G + 1;
    size_t n = (size_t)T * D;
    float* a = (float*)malloc(sizeof(float) * n * 2);
    float* b = a + n;
    oracle_embed(image, a, img_size, w[0], w[1], w[2], w[3]);
    for (int l = 0; l < VIT_O_DEPTH; ++l) {
        oracle_encoder_block(a, b, T, w + 4 + 12 * l);
        float* t = a; a = b; b = t;
    }
    float cls[D], lg[VIT_O_CLASSES];
    oracle_layer_norm(a, cls, 1, w[148], w[149]);
    oracle_linear(cls, lg, 1, D, VIT_O_CLASSES, w[150], w[151]);
    if (logits) memcpy(logits, lg, sizeof(lg));
    if (probs) oracle_softmax(lg, probs, VIT_O_CLASSES);
    free(a < b ? a : b);
}

int vit_oracle_forward(const float* const* weights, int img_size,
                       const float* images, int n,
                       float* logits, float* probs, int n_threads) {
    if (!weights || !images || n < 0 || img_size <= 0 || img_size % VIT_O_PATCH) return -1;
    for (int i = 0; i < VIT_O_NTENSORS; ++i)
        if (!weights[i]) return -1;
    const size_t img_elems = (size_t)VIT_O_CHANS * img_size * img_size;
    int nt = n_threads > 0 ? n_threads : oracle_max_threads();
#ifdef _OPENMP
    omp_set_dynamic(0);
    if (n >= nt && nt > 1) {
        /* images in parallel; the ops' own parallel regions are then serialised */
        omp_set_max_active_levels(1);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
        for (int i = 0; i < n; ++i)
            forward_one(weights, img_size, images + i * img_elems,
                        logits ? logits + (size_t)i * VIT_O_CLASSES : NULL,
                        probs ? probs + (size_t)i * VIT_O_CLASSES : NULL);
        return 0;
    }
    omp_set_num_threads(nt);
#endif
    for (int i = 0; i < n; ++i)
        forward_one(weights, img_size, images + i * img_elems,
                    logits ? logits + (size_t)i * VIT_O_CLASSES : NULL,
                    probs ? probs + (size_t)i * VIT_O_CLASSES : NULL);
    return 0;
}

/* ---- loader rounding (Network.c:185-187) -------------------------------------------------- */
void oracle_round_weights(float* w, size_t n) {
    for (size_t i = 0; i < n; ++i) w[i] = roundf(w[i] * 1000000.0f) / 1000000.0f;
}

/* ---- tensor sizes ------------------------------------------------------- */
size_t oracle_tensor_numel(int idx, int img_size) {
    const size_t G = (size_t)(img_size / VIT_O_PATCH), T = G * G + 1;
    if (idx < 0 || idx >= VIT_O_NTENSORS) return 0;
    if (idx == 0) return D;
    if (idx == 1) return (size_t)D * VIT_O_CHANS * VIT_O_PATCH * VIT_O_PATCH;
    if (idx == 2) return D;
    if (idx == 3) return T * D;
    if (idx == 148 || idx == 149) return D;
    if (idx == 150) return (size_t)VIT_O_CLASSES * D;
    if (idx == 151) return VIT_O_CLASSES;
    switch ((idx - 4) % 12) {
        case 0: case 1: case 6: case 7: case 5: case 11: return D;
        case 2: return (size_t)3 * D * D;
        case 3: return 3 * D;
        case 4: return (size_t)D * D;
        case 8: return (size_t)HID * D;
        case 9: return HID;
        case 10: return (size_t)D * HID;
    }
    return 0;
}
