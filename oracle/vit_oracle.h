/*
 * vit_oracle.h -- CPU oracle for the ViT-B/16 forward path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker (or as the CPU baseline being
 * timed), never as a fallback for the CUDA engine.
 *
 * The functions restate, in portable C, the arithmetic of the reference's sequential
 * implementation (reference ViT_seq.c) operation for operation: same accumulation
 * order, fp32 throughout, bias-first sums, no FMA contraction, libm expf/erff/sqrtf.
 * Each function cites the reference lines it follows.  Two deliberate extensions:
 * the token count is a run-time parameter (img_size 224 -> 197 tokens, 384 -> 577)
 * and the pre-softmax logits are returned (the reference discards them,
 * ViT_seq.c:432-437).
 *
 * Parity status: PINNED against the reference's own ViT_seq()/load_weights()/
 * load_image_data() objects compiled from /root/reference into oracle/_ref/
 * (tests/test_oracle_vs_reference.py, bit-exact probabilities), and against golden
 * vectors generated from that build (tests/golden/).  The shipped 100-image golden
 * file Data/answer_result.txt cannot be exercised: Data/input-100.bin and 36 weight
 * tensors are absent from the reference mount.
 */
#ifndef VIT_ORACLE_H
#define VIT_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIT_O_DIM      768
#define VIT_O_HEADS    12
#define VIT_O_HEAD_DIM 64
#define VIT_O_HIDDEN   3072
#define VIT_O_DEPTH    12
#define VIT_O_CLASSES  1000
#define VIT_O_PATCH    16
#define VIT_O_CHANS    3
#define VIT_O_NTENSORS 152

/* y[t][o] = b[o] + sum_i x[t][i] * W[o][i]   (ViT_seq.c:240-250) */
void oracle_linear(const float* x, float* y, int tokens, int in_f, int out_f,
                   const float* W, const float* b);

/* per-row LayerNorm, single-pass variance, eps 1e-6 added in double (ViT_seq.c:103-121) */
void oracle_layer_norm(const float* x, float* y, int tokens, const float* w, const float* b);

/* 0.5*x*(1+erff(x/sqrtf(2)))  (ViT_seq.c:231-233) */
float oracle_gelu(float x);
void  oracle_gelu_inplace(float* x, size_t n);

/* softmax(Q K^T / sqrtf(64)) V for all 12 heads; qkv are [tokens][768] each
 * (ViT_seq.c:156-215).  out is [tokens][768], heads interleaved along the row. */
void oracle_attention_core(const float* Q, const float* K, const float* V, float* out, int tokens);

/* in_proj -> attention core -> out_proj  (ViT_seq.c:123-229) */
void oracle_multihead_attn(const float* x, float* y, int tokens,
                           const float* in_w, const float* in_b,
                           const float* out_w, const float* out_b);

/* conv_proj + flatten_transpose + class_token + pos_emb (ViT_seq.c:25-101).
 * image is [3][S][S]; out is [tokens][768]. */
void oracle_embed(const float* image, float* out, int img_size,
                  const float* cls, const float* conv_w, const float* conv_b, const float* pos);

/* one pre-LN encoder block; w points at the 12 tensors of the layer in state_dict
 * order (ViT_seq.c:271-302). x -> y, both [tokens][768]. */
void oracle_encoder_block(const float* x, float* y, int tokens, const float* const* w);

/* stable softmax (ViT_seq.c:304-324) */
void oracle_softmax(const float* logits, float* probs, int n);

/* Whole model (ViT_seq.c:337-439) for n images laid out contiguously [n][3][S][S].
 * weights: 152 pointers in torchvision state_dict order.
 * logits and probs are [n][1000]; either may be NULL.  n_threads <= 0 means "all".
 * Images are independent; threading is over images first, then inside the ops.
 * Returns 0, or -1 on bad arguments. */
int vit_oracle_forward(const float* const* weights, int img_size,
                       const float* images, int n,
                       float* logits, float* probs, int n_threads);

/* Reference loader rounding (Network.c:185-187): w = roundf(w*1e6f)/1e6f, in place. */
void oracle_round_weights(float* w, size_t n);

/* Element count of tensor idx for a given img_size; 0 if idx invalid. */
size_t oracle_tensor_numel(int idx, int img_size);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
