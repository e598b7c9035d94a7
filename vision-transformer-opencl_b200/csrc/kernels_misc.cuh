// kernels_misc.cuh -- the HBM-bound kernels around the GEMMs: LayerNorm,
// class-token rows, the fp32 classifier head and operand conversion.
#pragma once

#include "ptx.cuh"

namespace vit {

constexpr int kDim = 768;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over 768 features, one warp per token row (layer_norm, ViT_seq.c:103-121; replaces
// layer_norm_kernel, kernel.cl:6-80).  fp32 in, eps = 1e-6 as in the CPU reference (the OpenCL
// kernel drops it), statistics in fp32 with a centred second pass over the registers, output
// in the GEMM operand precision.  Each lane owns 6 float4 (128-bit loads, fully coalesced).
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, T* __restrict__ y, int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * kDim);
    float4 v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = __ldcs(xr + lane + 32 * i);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / kDim);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float inv_std = rsqrtf(warp_sum(q) * (1.0f / kDim) + 1e-6f);
    const float4* wr = reinterpret_cast<const float4*>(w);
    const float4* br = reinterpret_cast<const float4*>(b);
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * kDim);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float4 g = __ldg(wr + lane + 32 * i), be = __ldg(br + lane + 32 * i);
        uint2 o;
        o.x = pack2<T>(fmaf(v[i].x * inv_std, g.x, be.x), fmaf(v[i].y * inv_std, g.y, be.y));
        o.y = pack2<T>(fmaf(v[i].z * inv_std, g.z, be.z), fmaf(v[i].w * inv_std, g.w, be.w));
        yr[lane + 32 * i] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm folded into the GEMMs (gemm_sm100_staged_kernel, LN = true): the two helpers around it.
//
// rowstats_cast (operator tests; the forward gets the same from the conv_proj epilogue and the class-row kernel) -- the
// operand-precision copy of the raw fp32 row plus its (sum, sum of squares), i.e. what the residual
// GEMMs emit for every later LayerNorm.  One warp per row, 128-bit loads.
template <typename T>
__global__ void __launch_bounds__(256) rowstats_cast_kernel(const float* __restrict__ x, T* __restrict__ y,
                                                            float2* __restrict__ stats, int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * kDim);
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * kDim);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float4 v = xr[lane + 32 * i];
        s += (v.x + v.y) + (v.z + v.w);
        q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
        uint2 o;
        o.x = pack2<T>(v.x, v.y);
        o.y = pack2<T>(v.z, v.w);
        yr[lane + 32 * i] = o;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) stats[row] = make_float2(s, q);
}

// fold_ln_weights (once, at init): W'[n][k] = round_T(ln_w[k] * W[n][k]),  colsum[n] = sum_k W'[n][k]
// (of the ROUNDED values, so that mean * colsum cancels exactly what the tensor cores accumulate),
// cvec[n] = bias[n] + sum_k ln_b[k] * W[n][k].  One warp per output feature n.
template <typename T>
__global__ void __launch_bounds__(256) fold_ln_weights_kernel(const float* __restrict__ W, const float* __restrict__ ln_w,
                                                              const float* __restrict__ ln_b, const float* __restrict__ bias,
                                                              T* __restrict__ Wp, float* __restrict__ colsum,
                                                              float* __restrict__ cvec, int N) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f, c = 0.f;
    for (int k = lane; k < kDim; k += 32) {
        const float w = W[static_cast<size_t>(n) * kDim + k];
        const T wp = from_float<T>(ln_w[k] * w);
        Wp[static_cast<size_t>(n) * kDim + k] = wp;
        if (!isfinite(to_float<T>(wp)) && isfinite(ln_w[k] * w)) atomicOr(&g_status_flags, VIT_FLAG_WEIGHT_RANGE);
        s += to_float<T>(wp);
        c = fmaf(ln_b[k], w, c);
    }
    s = warp_sum(s);
    c = warp_sum(c);
    if (lane == 0) {
        colsum[n] = s;
        cvec[n] = bias[n] + c;
    }
}

// ---------------------------------------------------------------------------------------------
// Class-token rows: X[b*tokens][:] = class_token + pos_embedding[0]  (class_token + pos_emb,
// ViT_seq.c:72-101).  The patch rows are written by the conv_proj GEMM epilogue.
__global__ void cls_rows_kernel(float* __restrict__ X, const float* __restrict__ cls, const float* __restrict__ pos,
                                int batch, int tokens) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * kDim) return;
    const int b = i / kDim, d = i - b * kDim;
    X[static_cast<size_t>(b) * tokens * kDim + d] = cls[d] + pos[d];
}

// Same for the LayerNorm-folded forward: also the operand-precision copy of the row and its (sum, sum of
// squares) as statistics part 0 (parts 1..5 zero), like the conv_proj epilogue emits for the patch rows.
// One warp per image.
template <typename T>
__global__ void __launch_bounds__(256) cls_rows_ln_kernel(float* __restrict__ X, T* __restrict__ Xc, float2* __restrict__ stats,
                                                          int stats_rows, const float* __restrict__ cls,
                                                          const float* __restrict__ pos, int batch, int tokens,
                                                          float* __restrict__ cls_rows32 = nullptr /* RES16: compact fp32 class rows */) {
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= batch) return;
    const int lane = threadIdx.x & 31;
    const size_t r = static_cast<size_t>(img) * tokens;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float4 a = reinterpret_cast<const float4*>(cls)[lane + 32 * i], b = reinterpret_cast<const float4*>(pos)[lane + 32 * i];
        const float4 v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        reinterpret_cast<float4*>(X + r * kDim)[lane + 32 * i] = v;
        if (cls_rows32) reinterpret_cast<float4*>(cls_rows32 + static_cast<size_t>(img) * kDim)[lane + 32 * i] = v;
        uint2 o;
        o.x = pack2<T>(v.x, v.y);
        o.y = pack2<T>(v.z, v.w);
        reinterpret_cast<uint2*>(Xc + r * kDim)[lane + 32 * i] = o;
        s += (v.x + v.y) + (v.z + v.w);
        q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane < 6) stats[static_cast<size_t>(lane) * stats_rows + r] = lane == 0 ? make_float2(s, q) : make_float2(0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------
// Class-row attention for the LAST encoder layer.  Only the class token of the last layer reaches the
// head (ViT_seq.c:429-435 normalises all rows and keeps row 0), and after the last attention no token reads
// another one, so everything behind it -- out_proj, LayerNorm, the MLP -- is needed for ONE row per image.  The
// query is that row; keys and values are still all of the image's tokens.
// One block per image, one warp per head:  s_j = q . k_j / 8 (lanes over keys), softmax in fp32 (exact maximum,
// as ViT_seq.c:178-191), o = sum_j p_j v_j (lanes over the 64 output columns).  Also gathers the image's fp32
// class row of the residual stream into the compact [batch][768] buffer the pruned layer tail works on.
// X16 != nullptr: the residual stream is held in the operand type (16-bit rows, RES16 forward); the class row is widened.
template <typename T>
__global__ void __launch_bounds__(384) cls_attention_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ X, const T* __restrict__ X16,
                                                           T* __restrict__ ao_c, float* __restrict__ x_c, int tokens) {
    __shared__ float sp[12][640];           // one row of scores / probabilities per head
    const int img = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t row0 = static_cast<size_t>(img) * tokens;
    if (threadIdx.x < kDim / 4 && (X16 || X)) {   // (neither: x_c already holds the fp32 class rows, RES16 forward)
        float4 v;
        if (X16) {
            const uint2 u = reinterpret_cast<const uint2*>(X16 + row0 * kDim)[threadIdx.x];
            const float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y);
            v = make_float4(a.x, a.y, b.x, b.y);
        } else {
            v = reinterpret_cast<const float4*>(X + row0 * kDim)[threadIdx.x];
        }
        reinterpret_cast<float4*>(x_c + static_cast<size_t>(img) * kDim)[threadIdx.x] = v;
    }
    const uint16_t* qrow = qkv + row0 * 2304 + head * 64;
    float q[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) {       // the whole query in every lane (uniform loads, 8 x 16 B)
        const uint4 v = *reinterpret_cast<const uint4*>(qrow + 8 * i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            q[8 * i + 2 * k] = to_float<T>(reinterpret_cast<const T*>(&w[k])[0]);
            q[8 * i + 2 * k + 1] = to_float<T>(reinterpret_cast<const T*>(&w[k])[1]);
        }
    }
    float* p = sp[head];
    float mx = -INFINITY;
    for (int j = lane; j < tokens; j += 32) {
        const uint16_t* krow = qkv + (row0 + j) * 2304 + 768 + head * 64;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(krow + 8 * i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc = fmaf(q[8 * i + 2 * k], to_float<T>(reinterpret_cast<const T*>(&w[k])[0]), acc);
                acc = fmaf(q[8 * i + 2 * k + 1], to_float<T>(reinterpret_cast<const T*>(&w[k])[1]), acc);
            }
        }
        acc *= 0.125f;
        p[j] = acc;
        mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < tokens; j += 32) {
        const float e = __expf(p[j] - mx);
        p[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    // o[d] for d = 2 lane, 2 lane + 1: V rows are always bf16 (the in_proj epilogue stores them so)
    float o0 = 0.f, o1 = 0.f;
    const __nv_bfloat162* vcol = reinterpret_cast<const __nv_bfloat162*>(qkv + row0 * 2304 + 1536 + head * 64) + lane;
    for (int j = 0; j < tokens; ++j) {
        const float2 v = __bfloat1622float2(vcol[static_cast<size_t>(j) * (2304 / 2)]);
        o0 = fmaf(p[j], v.x, o0);
        o1 = fmaf(p[j], v.y, o1);
    }
    const float inv = 1.0f / sum;
    reinterpret_cast<uint32_t*>(ao_c + static_cast<size_t>(img) * kDim + head * 64)[lane] = pack2<T>(o0 * inv, o1 * inv);
}

// ---------------------------------------------------------------------------------------------
// Final LayerNorm on the class rows only (the reference normalises all rows and keeps row 0,
// ViT_seq.c:429-433), fp32 out.  One warp per image.
// TIn = float (fp32 residual rows) or the operand type (RES16 forward: 16-bit residual rows).
template <typename TIn>
__global__ void __launch_bounds__(256) head_ln_kernel(const TIn* __restrict__ X, const float* __restrict__ w,
                                                      const float* __restrict__ b, float* __restrict__ out,
                                                      int batch, int tokens) {
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= batch) return;
    const int lane = threadIdx.x & 31;
    float4 v[6];
    if constexpr (sizeof(TIn) == 4) {
        const float4* xr = reinterpret_cast<const float4*>(X + static_cast<size_t>(img) * tokens * kDim);
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = xr[lane + 32 * i];
    } else {
        const uint2* xr = reinterpret_cast<const uint2*>(X + static_cast<size_t>(img) * tokens * kDim);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const uint2 u = xr[lane + 32 * i];
            const float2 a = unpack2<TIn>(u.x), c2 = unpack2<TIn>(u.y);
            v[i] = make_float4(a.x, a.y, c2.x, c2.y);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / kDim);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float inv_std = rsqrtf(warp_sum(q) * (1.0f / kDim) + 1e-6f);
    float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(img) * kDim);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float4 g = reinterpret_cast<const float4*>(w)[lane + 32 * i];
        const float4 be = reinterpret_cast<const float4*>(b)[lane + 32 * i];
        o[lane + 32 * i] = make_float4(fmaf(v[i].x * inv_std, g.x, be.x), fmaf(v[i].y * inv_std, g.y, be.y),
                                       fmaf(v[i].z * inv_std, g.z, be.z), fmaf(v[i].w * inv_std, g.w, be.w));
    }
}

// Classifier head in fp32 on the CUDA cores: logits[b][c] = bias[c] + sum_k xn[b][k] * W[c][k]
// (linear_layer with tokens = 1, ViT_seq.c:435).  0.002 % of the model's FLOPs; fp32 keeps the logits free of
// operand rounding.  One warp per class keeps its W row in registers (24 values per lane) and walks the images,
// which a block of 8 warps stages through shared memory 8 at a time.  Per output the summation order is fixed
// (lane-strided partial sums in ascending k, then a butterfly), so a logit does not depend on the batch size or
// the image's position in it; and a batch of ONE image still spreads over 63 blocks (the previous 64x64-tiled
// kernel ran 16 blocks for 134 us at batch 1).
constexpr int HEAD_IMGS = 16;   // 48 KB of shared memory per block
constexpr int HEAD_CLASSES = 16;   // per block: two per warp, so that every staged image row is read once per two classes
__global__ void __launch_bounds__(256) head_gemm_kernel(const float* __restrict__ xn, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ logits,
                                                        int batch, int classes) {
    __shared__ float4 sx[HEAD_IMGS][kDim / 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * HEAD_CLASSES + 2 * warp, c1 = c0 + 1;
    float4 w0[6], w1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        w0[i] = c0 < classes ? __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(c0) * kDim) + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        w1[i] = c1 < classes ? __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(c1) * kDim) + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float b0v = c0 < classes ? bias[c0] : 0.f, b1v = c1 < classes ? bias[c1] : 0.f;
    for (int b0 = blockIdx.y * HEAD_IMGS; b0 < batch; b0 += gridDim.y * HEAD_IMGS) {
        const int nb = min(HEAD_IMGS, batch - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * (kDim / 4); i += 256)
            sx[i / (kDim / 4)][i % (kDim / 4)] = reinterpret_cast<const float4*>(xn + static_cast<size_t>(b0) * kDim)[i];
        __syncthreads();
        for (int j = 0; j < nb; ++j) {
            float a0 = 0.f, a1 = 0.f;   // per output: lane-strided partial sums in ascending k, then a butterfly
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const float4 x = sx[j][lane + 32 * i];
                a0 = fmaf(x.x, w0[i].x, a0); a1 = fmaf(x.x, w1[i].x, a1);
                a0 = fmaf(x.y, w0[i].y, a0); a1 = fmaf(x.y, w1[i].y, a1);
                a0 = fmaf(x.z, w0[i].z, a0); a1 = fmaf(x.z, w1[i].z, a1);
                a0 = fmaf(x.w, w0[i].w, a0); a1 = fmaf(x.w, w1[i].w, a1);
            }
            a0 = warp_sum(a0);
            a1 = warp_sum(a1);
            if (lane == 0) {
                if (c0 < classes) logits[static_cast<size_t>(b0 + j) * classes + c0] = a0 + b0v;
                if (c1 < classes) logits[static_cast<size_t>(b0 + j) * classes + c1] = a1 + b1v;
                // An FP16 operand that overflowed anywhere upstream is inf, turns the fp32 residual row it feeds into inf / NaN
                // for good, and reaches the class row through the next attention: it always ends up here.
                if (!(fabsf(a0 + b0v) <= 3.0e38f) || !(fabsf(a1 + b1v) <= 3.0e38f)) atomicOr(&g_status_flags, VIT_FLAG_NONFINITE);
            }
        }
    }
}

// fp32 -> operand precision (weights at init, test inputs), and back (test outputs).  A finite value that does not
// fit the operand type (FP16: |w| > 65504) raises VIT_FLAG_WEIGHT_RANGE.
template <typename T>
__global__ void convert_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const T v = from_float<T>(src[i]);
        dst[i] = v;
        if (!isfinite(to_float<T>(v)) && isfinite(src[i])) atomicOr(&g_status_flags, VIT_FLAG_WEIGHT_RANGE);
    }
}
// fp32 -> tf32 (round to nearest, ties away: cvt.rna), still stored as fp32: conv_proj.weight for the kind::tf32 MMA,
// which ignores the 13 low mantissa bits of its operands.
__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(src[i]));
        dst[i] = __uint_as_float(r);
    }
}
// packed QKV activation [rows][2304] for the attention operator test: Q, K columns in T, V columns
// (>= 1536) in bf16 -- the layout the in_proj epilogue produces
template <typename T>
__global__ void convert_qkv_from_f32_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        if (i % 2304 >= 1536) {
            const __nv_bfloat16 v = __float2bfloat16_rn(src[i]);
            dst[i] = *reinterpret_cast<const uint16_t*>(&v);
        } else {
            const T v = from_float<T>(src[i]);
            dst[i] = *reinterpret_cast<const uint16_t*>(&v);
        }
    }
}
template <typename T>
__global__ void convert_to_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = to_float<T>(src[i]);
}

}  // namespace vit
