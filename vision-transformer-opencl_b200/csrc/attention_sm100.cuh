// attention_sm100.cuh -- fused multi-head softmax(Q K^T / sqrt(64)) V for sm_100a.
//
// Replaces the reference's per-head loop (ViT_seq.c:156-215; OpenCL: MHA_gemm_kernel +
// softmax_reduction_kernel + MHA_gemm_kernel per head, ViT_opencl.c:546-564).  One CTA per
// (image, head).  The [tokens x tokens] score matrix lives only in TMEM / registers / smem.
//
//   warp 8 (1 thread)  TMA: Q (2 x 128 rows), K, V tiles of the head straight out of the packed
//                      QKV activation [rows][2304]; tcgen05.mma S_t = Q_t K^T (K-major x K-major)
//                      and O_t = P_t V (V used as an MN-major B operand, no transpose pass)
//   warps 0-3 / 4-7    softmax warpgroup for query tile 0 / 1: one thread per query row reads its
//                      S row from TMEM (two passes: max, then exp2 + sum), writes P (operand
//                      precision) into 128B-swizzled smem, later scales O by 1/sum and stores it
//
// Keys are padded to a multiple of 16 (197 -> 208); padded columns are masked to -inf before
// the max, padded/foreign V rows meet P == 0.  This variant keeps the whole key range in one
// block, so tokens <= 256.
#pragma once

#include "ptx.cuh"

namespace vit {

struct AttnParams {
    int batch;
    int tokens;
    int kpad;          // tokens rounded up to 16
    void* out;         // [batch*tokens][768], operand precision
    float scale_log2;  // (1/sqrt(64)) * log2(e)
};

constexpr int ATTN_THREADS = 288;
constexpr int ATTN_DIM = 768;
constexpr int ATTN_DH = 64;
constexpr int ATTN_Q_TILE_BYTES = 128 * 128;  // 128 rows x 64 x 2 B

__host__ __device__ inline int attn_kv_bytes(int kpad) { return kpad * 128; }
__host__ __device__ inline int attn_p_tile_bytes(int kpad) { return ((kpad + 63) / 64) * ATTN_Q_TILE_BYTES; }
__host__ inline int attn_smem_bytes(int kpad) {
    return 2 * ATTN_Q_TILE_BYTES + 2 * attn_kv_bytes(kpad) + 2 * attn_p_tile_bytes(kpad) + 128 + 1024;
}

template <typename T>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attention_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                       const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kv_bytes = attn_kv_bytes(p.kpad);
    const int p_bytes = attn_p_tile_bytes(p.kpad);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + 2 * ATTN_Q_TILE_BYTES;
    uint8_t* sV = sK + kv_bytes;
    uint8_t* sP = sV + kv_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * p_bytes);
    uint64_t* bar_load = bars;       // Q,K,V landed
    uint64_t* bar_s = bars + 1;      // [2] S_t ready in TMEM
    uint64_t* bar_p = bars + 3;      // [2] P_t written (128 arrivals)
    uint64_t* bar_o = bars + 5;      // [2] O_t ready in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int img = blockIdx.x / 12;
    const int head = blockIdx.x - img * 12;
    const int row0 = img * p.tokens;  // first activation row of this image
    const int nqt = p.tokens > 128 ? 2 : 1;

    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q);
            tma_prefetch_desc(&tmap_kv);
            mbar_init(bar_load, 1);
            for (int t = 0; t < 2; ++t) {
                mbar_init(&bar_s[t], 1);
                mbar_init(&bar_p[t], 128);
                mbar_init(&bar_o[t], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_load, 2 * ATTN_Q_TILE_BYTES + 2 * kv_bytes);
            tma_load_2d(sQ, &tmap_q, bar_load, head * ATTN_DH, row0);
            tma_load_2d(sK, &tmap_kv, bar_load, ATTN_DIM + head * ATTN_DH, row0);
            tma_load_2d(sV, &tmap_kv, bar_load, 2 * ATTN_DIM + head * ATTN_DH, row0);
            mbar_wait(bar_load, 0);
            tc_fence_after();
            const uint32_t idesc_s = make_idesc<T>(128, static_cast<uint32_t>(p.kpad), 0, 0);
            const uint32_t idesc_o = make_idesc<T>(128, ATTN_DH, 0, 1);
            const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
            for (int t = 0; t < nqt; ++t) {
#pragma unroll
                for (int k = 0; k < ATTN_DH / 16; ++k)
                    umma_f16(tmem_base + t * 256, desc_kmajor_sw128(q_addr + t * ATTN_Q_TILE_BYTES, k),
                             desc_kmajor_sw128(k_addr, k), idesc_s, k != 0);
                umma_commit(&bar_s[t]);
            }
            const int ksteps = p.kpad / 16;
            for (int t = 0; t < nqt; ++t) {
                mbar_wait(&bar_p[t], 0);
                tc_fence_after();
                for (int ks = 0; ks < ksteps; ++ks)
                    umma_f16(tmem_base + t * 256,
                             desc_kmajor_sw128(p_addr + t * p_bytes + (ks >> 2) * ATTN_Q_TILE_BYTES, ks & 3),
                             desc_mnmajor_sw128(v_addr, ks), idesc_o, ks != 0);
                umma_commit(&bar_o[t]);
            }
        }
    } else {
        const int t = warp >> 2;        // query tile of this warpgroup
        const int quarter = warp & 3;   // TMEM lane quarter
        const int qrow = t * 128 + quarter * 32 + lane;  // query index inside the image
        const bool warp_active = t < nqt && (t * 128 + quarter * 32) < p.tokens;
        float inv_sum = 0.f;
        if (warp_active) {
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
            mbar_wait(&bar_s[t], 0);
            tc_fence_after();
            const int nch = p.kpad / 16;
            float mx = -INFINITY;
            for (int ch = 0; ch < nch; ++ch) {
                uint32_t r[16];
                tmem_ld_x16(taddr + ch * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = (ch * 16 + j < p.tokens) ? __uint_as_float(r[j]) : -INFINITY;
                    mx = fmaxf(mx, v);
                }
            }
            const float moff = -mx * p.scale_log2;
            float sum = 0.f;
            uint8_t* prow = sP + t * p_bytes + (quarter * 32 + lane) * 128;
            const int sw = lane & 7;  // (row % 8) of the 128B swizzle; row = quarter*32+lane
            for (int ch = 0; ch < nch; ++ch) {
                uint32_t r[16];
                tmem_ld_x16(taddr + ch * 16, r);
                tmem_ld_wait();
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c0 = ch * 16 + 2 * j;
                    float e0 = fast_exp2(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, moff));
                    float e1 = fast_exp2(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, moff));
                    e0 = (c0 < p.tokens) ? e0 : 0.f;
                    e1 = (c0 + 1 < p.tokens) ? e1 : 0.f;
                    sum += e0 + e1;
                    packed[j] = pack2<T>(e0, e1);
                }
                // keys [ch*16, ch*16+16) = 16B chunks 2ch, 2ch+1 of the row; K-block = chunk / 8
                const int c8 = ch * 2;
                uint8_t* blk = prow + (c8 >> 3) * ATTN_Q_TILE_BYTES;
                *reinterpret_cast<uint4*>(blk + (((c8 & 7) ^ sw) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                *reinterpret_cast<uint4*>(blk + ((((c8 + 1) & 7) ^ sw) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
            inv_sum = 1.0f / sum;
            fence_proxy_async_smem();
            tc_fence_before();
        }
        if (t < 2) mbar_arrive(&bar_p[t]);
        if (warp_active) {
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
            mbar_wait(&bar_o[t], 0);
            tc_fence_after();
            T* orow = static_cast<T*>(p.out) + static_cast<size_t>(row0 + qrow) * ATTN_DIM + head * ATTN_DH;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tmem_ld_x32(taddr + half * 32, r);
                tmem_ld_wait();
                if (qrow < p.tokens) {
                    uint4* dst = reinterpret_cast<uint4*>(orow + half * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            w[q] = pack2<T>(__uint_as_float(r[8 * j + 2 * q]) * inv_sum,
                                            __uint_as_float(r[8 * j + 2 * q + 1]) * inv_sum);
                        dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vit
