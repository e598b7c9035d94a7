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
            const uint32_t idesc_o = make_idesc<__nv_bfloat16>(128, ATTN_DH, 0, 1);  // P, V are always bf16
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
                    packed[j] = pack2<__nv_bfloat16>(e0, e1);
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

// =============================================================================================
// Persistent, software-pipelined variant (the one the engine uses).
//
// One CTA per SM loops over (image, head) items.  Per item the two 128-row query tiles own one
// 256-column TMEM region each:   S_t fp32 [0,kpad)  ->  P_t (operand precision, packed two per
// column, written back in place by the softmax threads) [0,kpad/2)  ->  O_t fp32 [128,192).
// P never touches shared memory: the second MMA takes its A operand from TMEM.  Shared memory
// holds only the double-buffered Q/K/V tiles of the current and the next item, so the TMA loads
// of item i+1 run under the softmax of item i, and while one warpgroup is in its softmax the
// tensor core works for the other one.
//
//   warp 8  TMA producer          warp 9  MMA issuer + TMEM owner
//   warps 0-3 / 4-7  softmax + output warpgroups for query tile 0 / 1
constexpr int ATTN2_THREADS = 320;
// TMEM column plan (512 columns).  With kpad <= 224 (ViT-B/16 at 224^2: kpad = 208):
//   S_0/P_0 [0,kpad)   S_1/P_1 [kpad,2 kpad)   O_1 inside its S region at +128   O_0 [2 kpad, 2 kpad+64)
// so S_0 of the next item can be issued without waiting for O_0 to be drained.  Otherwise
//   S_0 [0,256)  S_1 [256,512)  O_t inside its own S region at +128.
struct AttnTmemPlan {
    uint32_t s1, o0, o1;  // column of S_1, O_0, O_1 (S_0 is at column 0)
    bool spare;
    __device__ __forceinline__ uint32_t s_col(int t) const { return t ? s1 : 0u; }
    __device__ __forceinline__ uint32_t o_col(int t) const { return t ? o1 : o0; }
};
__device__ __forceinline__ AttnTmemPlan attn2_tmem_plan(int kpad) {
    AttnTmemPlan pl;
    pl.spare = 2 * kpad + 64 <= 512;
    pl.s1 = pl.spare ? kpad : 256;
    pl.o1 = pl.s1 + 128;
    pl.o0 = pl.spare ? 2 * kpad : 128;
    return pl;
}
__host__ __device__ inline int attn2_stage_bytes(int kpad) { return 2 * ATTN_Q_TILE_BYTES + 2 * attn_kv_bytes(kpad); }
__host__ inline int attn2_smem_bytes(int kpad) { return 2 * attn2_stage_bytes(kpad) + 256 + 1024; }

template <typename T>
__global__ void __launch_bounds__(ATTN2_THREADS, 1)
attention_sm100_persistent_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                                  const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kv_bytes = attn_kv_bytes(p.kpad);
    const int stage_bytes = attn2_stage_bytes(p.kpad);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
    uint64_t* kv_full = bars;        // [2 stages] Q,K,V of an item landed (tx)
    uint64_t* stage_free = bars + 2; // [2 stages] every MMA reading the stage has completed
    uint64_t* s_full = bars + 4;     // [2 tiles]  S_t in TMEM
    uint64_t* p_full = bars + 6;     // [2 tiles]  P_t written back (128 arrivals)
    uint64_t* o_full = bars + 8;     // [2 tiles]  O_t in TMEM
    uint64_t* o_free = bars + 10;    // [2 tiles]  O_t drained, region reusable (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.batch * 12;
    const int nqt = p.tokens > 128 ? 2 : 1;
    const AttnTmemPlan plan = attn2_tmem_plan(p.kpad);

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&stage_free[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_free[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int s = it & 1;
                const int img = item / 12, head = item - img * 12;
                const int row0 = img * p.tokens;
                uint8_t* sQ = smem + s * stage_bytes;
                uint8_t* sK = sQ + 2 * ATTN_Q_TILE_BYTES;
                uint8_t* sV = sK + kv_bytes;
                mbar_wait(&stage_free[s], ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
                tma_load_2d(sQ, &tmap_q, &kv_full[s], head * ATTN_DH, row0);
                tma_load_2d(sK, &tmap_kv, &kv_full[s], ATTN_DIM + head * ATTN_DH, row0);
                tma_load_2d(sV, &tmap_kv, &kv_full[s], 2 * ATTN_DIM + head * ATTN_DH, row0);
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp runs the (warp-uniform) state machine so that descriptors and barrier
        // addresses live in uniform registers; one elected lane issues the tcgen05 instructions.
        {
            const uint32_t idesc_s = make_idesc<T>(128, static_cast<uint32_t>(p.kpad), 0, 0);
            const uint32_t idesc_o = make_idesc<__nv_bfloat16>(128, ATTN_DH, 0, 1);  // P, V are always bf16
            const int ksteps = p.kpad / 16;
            auto ready = [&](uint64_t* bar, uint32_t parity) {  // non-blocking, warp-uniform
                return __shfl_sync(0xffffffffu, mbar_test_wait(bar, parity) ? 1 : 0, 0) != 0;
            };
            auto issue_s = [&](int t, int stage) {  // S_t = Q_t K^T of the item staged in `stage`
                if (elect_one()) {
                    const uint32_t q_addr = smem_u32(smem + stage * stage_bytes);
                    const uint32_t k_addr = q_addr + 2 * ATTN_Q_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < ATTN_DH / 16; ++k)
                        umma_f16(tmem_base + plan.s_col(t), desc_kmajor_sw128(q_addr + t * ATTN_Q_TILE_BYTES, k),
                                 desc_kmajor_sw128(k_addr, k), idesc_s, k != 0);
                    umma_commit(&s_full[t]);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int t, int stage) {  // O_t = P_t V, P_t read from TMEM
                if (elect_one()) {
                    const uint32_t v_addr = smem_u32(smem + stage * stage_bytes) + 2 * ATTN_Q_TILE_BYTES + kv_bytes;
                    for (int ks = 0; ks < ksteps; ++ks)
                        umma_f16_ts(tmem_base + plan.o_col(t), tmem_base + plan.s_col(t) + ks * 8,
                                    desc_mnmajor_sw128(v_addr, ks), idesc_o, ks != 0);
                    umma_commit(&o_full[t]);
                }
                __syncwarp();
            };
            // Event-driven issue: each query tile is a small state machine
            //     need P_t(i)  -> issue PV_t(i)            [t = 0 with spare columns: also O_0(i-1) drained]
            //     need K(i+1)  -> issue S_t(i+1)           [O_t inside the S region: also O_t(i) drained]
            // polled round robin, so the two softmax warpgroups run out of phase instead of being
            // re-synchronised by a fixed issue order, and the tensor pipe serves whichever is ready.
            // The tensor pipe executes in issue order, which protects P_t(i) from S_t(i+1).
            const int my_items = blockIdx.x < n_items ? (n_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
            if (my_items > 0) {
                mbar_wait(&kv_full[0], 0);
                tc_fence_after();
                for (int t = 0; t < nqt; ++t) issue_s(t, 0);
            }
            int it_t[2] = {0, 0};       // item each tile is working on
            int phase_t[2] = {0, 0};    // 0: waiting for P (issue PV), 1: waiting to issue S of the next item
            int pv_issued[2] = {0, 0};  // number of items whose PV_t has been issued
            int stage_committed = 0;    // items whose smem stage has been handed back to the producer
            int active = (my_items > 0) ? nqt : 0;
            uint32_t spins = 0;
            uint64_t t_start = 0;
            while (active > 0) {
                if ((++spins & 0xfff) == 0) {  // watchdog: a protocol bug must trap, not hang the GPU
                    const uint64_t now = global_timer_ns();
                    if (t_start == 0) t_start = now;
                    else if (now - t_start > VIT_WATCHDOG_NS) { atomicExch(&g_watchdog_flag, 2u); __trap(); }
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (t >= nqt || it_t[t] >= my_items) continue;
                    const int it = it_t[t];
                    if (phase_t[t] == 0) {
                        if (!ready(&p_full[t], it & 1)) continue;
                        if (t == 0 && plan.spare && !ready(&o_free[0], (it & 1) ^ 1)) continue;
                        tc_fence_after();
                        issue_pv(t, it & 1);
                        pv_issued[t] = it + 1;
                        // both tiles' PV of item `stage_committed` issued: its Q/K/V stage can be refilled
                        while (stage_committed < pv_issued[0] && (nqt == 1 || stage_committed < pv_issued[1])) {
                            if (elect_one()) umma_commit(&stage_free[stage_committed & 1]);
                            __syncwarp();
                            ++stage_committed;
                        }
                        if (it + 1 >= my_items) {
                            it_t[t] = my_items;
                            --active;
                        } else {
                            phase_t[t] = 1;
                        }
                    } else {
                        if (!ready(&kv_full[(it + 1) & 1], ((it + 1) >> 1) & 1)) continue;
                        if (!(t == 0 && plan.spare) && !ready(&o_free[t], it & 1)) continue;
                        tc_fence_after();
                        issue_s(t, (it + 1) & 1);
                        it_t[t] = it + 1;
                        phase_t[t] = 0;
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax / output warpgroups
        const int t = warp >> 2;
        const int quarter = warp & 3;
        const int qrow = t * 128 + quarter * 32 + lane;
        const bool warp_active = t < nqt && (t * 128 + quarter * 32) < p.tokens;
        const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t taddr = tmem_base + lane_bits + plan.s_col(t & 1);
        const uint32_t oaddr = tmem_base + lane_bits + plan.o_col(t & 1);
        const int nch = p.kpad / 16;
        int it = 0;
        // A warp whose rows are all padding still walks the barriers in lockstep (an mbarrier cannot
        // take arrivals for a future phase); a whole unused warpgroup (tokens <= 128) does nothing.
        for (int item = blockIdx.x; t < nqt && item < n_items; item += gridDim.x, ++it) {
            const int img = item / 12, head = item - img * 12;
            float inv_sum = 0.f;
            mbar_wait(&s_full[t], it & 1);
            tc_fence_after();
            if (warp_active) {
                // ONE pass over the S row (TMEM reads, 64 B/clk/SM, are this kernel's scarcest resource).
                // Softmax is shift invariant, so the exponent offset need not be the row maximum (which
                // would cost a first full pass): it starts as the maximum of the first 32 scores and is
                // raised lazily.  Before the exponentials of each 32-column step the step maximum is
                // compared with the offset; only if it exceeds it by more than 2^kLazy is everything
                // written so far (P in TMEM, the row sum) rescaled by an exact integer power of two and
                // the offset moved -- the online-softmax recurrence with the rescale made rare.  P is
                // bf16 (fp32 exponent range), so values up to 2^kLazy are harmless, and no exponential
                // is ever evaluated above that bound, whatever the scores are.
                // Two x16 TMEM loads per round trip, the next step's loads in flight while the current
                // one is processed.  P columns [16 st, 16 st + 16) are written in place and never overlap
                // S columns not yet read.
                constexpr float kLazy = 24.f;
                const int nsteps = (nch + 1) >> 1;
                uint32_t ra[32], rb[32];
                auto load_step = [&](uint32_t* buf, int st) {
                    tmem_ld_x16p(taddr + st * 32, buf);
                    if (2 * st + 1 < nch) tmem_ld_x16p(taddr + st * 32 + 16, buf + 16);
                };
                float sum4[4] = {0.f, 0.f, 0.f, 0.f};  // independent chains (ILP)
                float moff = 0.f;                      // -(offset) * scale * log2(e)
                auto exp_step = [&](const uint32_t* v, int st) {
                    const int base = st * 32;
                    const bool full = base + 32 <= p.tokens;
                    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            m4[j & 3] = fmaxf(m4[j & 3], fmaxf(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (base + j < p.tokens) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
                    }
                    const float smax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                    if (st == 0) {
                        moff = -smax * p.scale_log2;
                    } else {
                        const float over = fmaf(smax, p.scale_log2, moff);  // log2 of this step's largest p
                        if (__any_sync(0xffffffffu, over > kLazy)) {
                            // exact repair: 2^-sh on P[0, 16 st) and on the sum, offset raised by sh
                            const int sh = over > kLazy ? static_cast<int>(ceilf(fminf(over, 1.0e6f))) : 0;
                            const float f = sh > 126 ? 0.f : __int_as_float((127 - sh) << 23);
                            tmem_st_wait();
                            for (int c8 = 0; c8 < 2 * st; ++c8) {
                                uint32_t w[8];
                                tmem_ld_x8p(taddr + c8 * 8, w);
                                tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    w[j] = pack2<__nv_bfloat16>(__uint_as_float(w[j] << 16) * f, __uint_as_float(w[j] & 0xffff0000u) * f);
                                tmem_st_x8p(taddr + c8 * 8, w);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) sum4[j] *= f;
                            moff -= static_cast<float>(sh);
                        }
                    }
                    uint32_t packed[16];
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float e0 = fast_exp2(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, moff));
                            const float e1 = fast_exp2(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, moff));
                            sum4[j & 3] += e0 + e1;
                            packed[j] = pack2<__nv_bfloat16>(e0, e1);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c0 = base + 2 * j;
                            float e0 = 0.f, e1 = 0.f;
                            if (c0 < p.tokens) e0 = fast_exp2(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, moff));
                            if (c0 + 1 < p.tokens) e1 = fast_exp2(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, moff));
                            sum4[j & 3] += e0 + e1;
                            packed[j] = pack2<__nv_bfloat16>(e0, e1);
                        }
                    }
                    tmem_st_x8p(taddr + st * 16, packed);
                    if (2 * st + 1 < nch) tmem_st_x8p(taddr + st * 16 + 8, packed + 8);
                };
                load_step(ra, 0);
                for (int st = 0; st < nsteps; st += 2) {
                    tmem_ld_wait();
                    if (st + 1 < nsteps) load_step(rb, st + 1);
                    exp_step(ra, st);
                    if (st + 1 < nsteps) {
                        tmem_ld_wait();
                        if (st + 2 < nsteps) load_step(ra, st + 2);
                        exp_step(rb, st + 1);
                    }
                }
                const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
                inv_sum = fast_rcp(sum);
                tmem_st_wait();
                tc_fence_before();
            }
            mbar_arrive(&p_full[t]);
            mbar_wait(&o_full[t], it & 1);
            tc_fence_after();
            if (warp_active) {
                uint32_t r0[32], r1[32];
                tmem_ld_x32(oaddr, r0);
                tmem_ld_x32(oaddr + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&o_free[t]);
                if (qrow < p.tokens) {
                    T* orow = static_cast<T*>(p.out) + (static_cast<size_t>(img) * p.tokens + qrow) * ATTN_DIM + head * ATTN_DH;
                    uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            w[q] = pack2<T>(__uint_as_float(r0[8 * j + 2 * q]) * inv_sum, __uint_as_float(r0[8 * j + 2 * q + 1]) * inv_sum);
                        dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            w[q] = pack2<T>(__uint_as_float(r1[8 * j + 2 * q]) * inv_sum, __uint_as_float(r1[8 * j + 2 * q + 1]) * inv_sum);
                        dst[4 + j] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            } else {
                mbar_arrive(&o_free[t]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vit
