// attention_sm100.cuh -- fused multi-head softmax(Q K^T / sqrt(64)) V for sm_100a.
//
// Replaces the reference's per-head loop (ViT_seq.c:156-215; OpenCL: MHA_gemm_kernel +
// softmax_reduction_kernel + MHA_gemm_kernel per head, ViT_opencl.c:546-564).  The [tokens x tokens]
// score matrix lives only in TMEM and registers.  Keys are padded to a multiple of 16 (197 -> 208);
// padded columns get P == 0, so the foreign / zero-filled V rows behind them never contribute.
// The whole key range of an image is one block, so tokens <= 256 here; longer sequences (384^2:
// 577 tokens) use the key-blocked kernel below.
#pragma once

#include "ptx.cuh"

namespace vit {

struct AttnParams {
    int batch;
    int tokens;
    int kpad;          // tokens rounded up to 16
    void* out;         // [batch*tokens][768], operand precision
    float scale_log2;  // (1/sqrt(64)) * log2(e)
    unsigned long long* trace;  // optional (debug): SM-clock timestamps of CTA 0's pipeline events, see ATTN_TRACE
};

// Pipeline trace of CTA 0 (vit_cuda_debug_attention_trace): trace[(warp * ITEMS + item) * EVENTS + event] = clock64().
// softmax warps 0-15: 0 item start, 1 S ready, 2 P written, 3 O ready, 4 O in registers, 5 O stored
// warp 16 (producer): 0 stage free / TMA issued        warps 17, 18 (MMA issuers): 0 S_t issued, 1 PV_t issued
constexpr int ATTN_TRACE_ITEMS = 16, ATTN_TRACE_EVENTS = 8, ATTN_TRACE_WARPS = 19;
#define ATTN_TRACE(warp_, it_, ev_)                                                                          \
    do {                                                                                                     \
        if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (it_) < ATTN_TRACE_ITEMS)    \
            p.trace[((warp_) * ATTN_TRACE_ITEMS + (it_)) * ATTN_TRACE_EVENTS + (ev_)] = clock64();           \
    } while (0)

constexpr int ATTN_DIM = 768;
constexpr int ATTN_DH = 64;
constexpr int ATTN_Q_TILE_BYTES = 128 * 128;  // 128 rows x 64 x 2 B

__host__ __device__ inline int attn_kv_bytes(int kpad) { return kpad * 128; }

// Shared memory of the single-block kernels: two stages of {Q, K, V} x kpad rows x 128 B (Q only needs `tokens`
// rows; the second query tile's descriptor runs on into the K rows behind it, whose S rows nobody reads).
__host__ __device__ inline int attn2_stage_bytes(int kpad) { return 3 * attn_kv_bytes(kpad); }
constexpr int ATTN2_OSTAGE_BYTES = 128 * 128;   // one 128-row output staging tile (key-blocked two-pass kernel)

template <int NTHREADS>
__device__ __forceinline__ void attn_bar_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTHREADS) : "memory");
}

// The single-pass softmax raises VIT_FLAG_ATTN_RANGE in g_status_flags (ptx.cuh) when a row's exponent range left
// its safe window (see the kernels): the host then repeats the work with the exact two-pass variant.
constexpr float ATTN_FAST_SHIFT = 64.f;      // exponent head-room of the single-pass softmax (log2 units)
constexpr float ATTN_FAST_SUM_MAX = 1.0e30f; // ~2^100: a larger row sum means the window was left

// EXACT = true : two passes over S (exact row maximum first), P <= 1 as in the reference.
// EXACT = false: ONE pass.  Softmax is shift invariant, so the exponent offset need not be the row
//   maximum: it is m_ref = max of a few of the row's scores, lowered by 2^-64:  P_j = 2^((s_j - m_ref) c - 64).
//   The true maximum is >= m_ref, so the largest P is >= 2^-64 (nothing relevant underflows: bf16 and fp32
//   share the 8-bit exponent), and nothing overflows unless some score exceeds m_ref by more than ~160 / c
//   (a logit gap of > 110 between a key and the best of the reference keys).  That case is DETECTED (row sum
//   beyond 2^100 or not finite -> VIT_FLAG_ATTN_RANGE) and the host reruns with EXACT = true.
//   TMEM reads (~64 B/clk/SM) bound these kernels, and this halves the reads of S.

// =============================================================================================
// Streaming kernel (the one the engine uses for tokens <= 224).
//
// (Round 1's first kernel kept the two query tiles of an item in two TMEM slots, each a strictly serial chain
// S MMA -> softmax -> PV MMA -> output epilogue -> next S MMA; its timeline, profiles/r1_attention_trace.md,
// showed the softmax threads waiting or doing epilogue work for more than half of every period.  It is gone.)
//
// Here the unit of work is ONE 128-row query tile (an item is two consecutive units), the S accumulator
// is DOUBLE BUFFERED across units, and the roles are split so that nobody who exponentiates ever waits:
//
//   TMEM     S[0] [0,kpad)   S[1] [kpad,2 kpad)   O [2 kpad, 2 kpad + 64)          (kpad <= 224)
//   warps 0-11   exponentials: warp = part * 4 + lane quarter.  The 16-key chunks of a row are split into
//                three contiguous ranges (13 chunks: 5 + 4 + 4), one per part, so THREE threads share a row.
//                P (bf16) is written in place over the first half of the thread's own range.  When a unit
//                is done they go straight on to the next one, whose S is already waiting in the other buffer.
//   warps 12-15  output: one per lane quarter; O / row sum -> staging -> one TMA store of 32 rows, while
//                the exponential warps are already in the next unit.
//   warp 16      TMA producer (Q, K, V of an item, two stages)     warp 17  MMA issuer, TMEM owner
//
// Issue order of the tensor pipe:  S(0) | S(u+1), PV(u) | ...  -- S(u+1) overwrites the buffer that held
// P(u-1), whose PV(u-1) was issued before it (the pipe executes in order); PV(u) waits for P(u) and for the
// output warps to have drained O(u-1).
//
// Softmax: EXACT (two passes, row maximum exchanged between the three parts through shared memory),
// otherwise single pass with the exponent offset m_ref = max of TWO scores near the end of part 0's range --
// columns that never receive P (a part's P fills only the first half of its range), so all three parts read
// the same two scores whatever their relative progress.
constexpr int ATTN3_THREADS = 18 * 32;
constexpr int ATTN3_EXP_WARPS = 12, ATTN3_W_OUT = 12, ATTN3_W_PRODUCER = 16, ATTN3_W_ISSUER = 17;
constexpr int ATTN3_OSTAGE_BYTES = 4 * 2 * 4096;                       // [output warp][2] 32 rows x 128 B
constexpr int ATTN3_XCH_BYTES = (3 * 3 * 128 + 2 * 3 * 128) * 4;       // sum [unit % 3][part][row], max [unit parity][part][row]
__host__ inline int attn3_smem_bytes(int kpad) {
    return 2 * attn2_stage_bytes(kpad) + ATTN3_OSTAGE_BYTES + ATTN3_XCH_BYTES + 256 + 1024;
}

template <typename T, bool EXACT>
__global__ void __launch_bounds__(ATTN3_THREADS, 1)
attention_sm100_stream_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out32,
                              const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kv_bytes = attn_kv_bytes(p.kpad);
    const int stage_bytes = attn2_stage_bytes(p.kpad);
    uint8_t* sO = smem + 2 * stage_bytes;
    // Row sums are TRIPLE buffered: the exponential warps may write the sums of unit u + 2 before the output warps
    // have read those of unit u (nothing orders the two), but never those of unit u + 3 -- S(u + 3) is issued after
    // PV(u + 1), which waits for o_free(u), which the output warps give only after reading the sums of unit u.
    float* xsum = reinterpret_cast<float*>(sO + ATTN3_OSTAGE_BYTES);  // [unit % 3][part][128]
    float* xmax = xsum + 3 * 3 * 128;                                 // [unit parity][part][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xsum) + ATTN3_XCH_BYTES);
    uint64_t* kv_full = bars;        // [2 stages] Q, K, V of an item landed (tx)
    uint64_t* stage_free = bars + 2; // [2 stages] every MMA reading the stage has completed
    uint64_t* s_full = bars + 4;     // [2 buffers] S in TMEM
    uint64_t* p_full = bars + 6;     // [2 buffers] P written back (one arrival per exponential warp)
    uint64_t* o_full = bars + 8;     // O in TMEM
    uint64_t* o_free = bars + 9;     // O drained (one arrival per output warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.batch * 12;
    const int nqt = p.tokens > 128 ? 2 : 1;
    const int my_items = blockIdx.x < n_items ? (n_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    const int n_units = my_items * nqt;
    const int nch = p.kpad / 16;
    // contiguous chunk ranges of the three parts: sizes differ by at most one, larger ones first
    const int sz = nch / 3, rem = nch - 3 * sz;
    const int c0_1 = sz + (rem > 0), c0_2 = c0_1 + sz + (rem > 1);   // part 0: [0, c0_1)  part 1: [c0_1, c0_2)  part 2: [c0_2, nch)
    const uint32_t o_col = 2 * p.kpad;

    if (warp == ATTN3_W_PRODUCER && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_out32);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&stage_free[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], ATTN3_EXP_WARPS);
        }
        mbar_init(o_full, 1);
        mbar_init(o_free, 4);
        fence_barrier_init();
    }
    if (warp == ATTN3_W_ISSUER) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();     // the set-up above overlaps the previous kernel's tail (programmatic dependent launch)
    griddep_launch();

    if (warp == ATTN3_W_PRODUCER) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const int half_rows = p.kpad >> 1, half_bytes = half_rows * 128;
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int s = it & 1;
                const int img = item / 12, head = item - img * 12;
                const int row0 = img * p.tokens;
                uint8_t* sQ = smem + s * stage_bytes;
                uint8_t* sK = sQ + kv_bytes;
                uint8_t* sV = sK + kv_bytes;
                mbar_wait(&stage_free[s], ((it >> 1) & 1) ^ 1);
                ATTN_TRACE(warp, it, 0);
                mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    tma_load_2d(sQ + h * half_bytes, &tmap_qkv, &kv_full[s], head * ATTN_DH, row0 + h * half_rows);
                    tma_load_2d(sK + h * half_bytes, &tmap_qkv, &kv_full[s], ATTN_DIM + head * ATTN_DH, row0 + h * half_rows);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(sV + h * half_bytes, &tmap_qkv, &kv_full[s], 2 * ATTN_DIM + head * ATTN_DH, row0 + h * half_rows);
                const int ahead = item + 2 * static_cast<int>(gridDim.x);
                if (ahead < n_items) {
                    const int img2 = ahead / 12, head2 = ahead - img2 * 12;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        tma_prefetch_l2_2d(&tmap_qkv, head2 * ATTN_DH, img2 * p.tokens + h * half_rows);
                        tma_prefetch_l2_2d(&tmap_qkv, ATTN_DIM + head2 * ATTN_DH, img2 * p.tokens + h * half_rows);
                        tma_prefetch_l2_2d(&tmap_qkv, 2 * ATTN_DIM + head2 * ATTN_DH, img2 * p.tokens + h * half_rows);
                    }
                }
            }
        }
    } else if (warp == ATTN3_W_ISSUER) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc_s = make_idesc<T>(128, static_cast<uint32_t>(p.kpad), 0, 0);
        const uint32_t idesc_o = make_idesc<__nv_bfloat16>(128, ATTN_DH, 0, 1);  // P, V are always bf16
        auto issue_s = [&](int u) {  // S(u) = Q_t K^T into buffer u & 1
            const int it = u / nqt, t = u - it * nqt, stage = it & 1;
            if (t == 0) {
                mbar_wait(&kv_full[stage], (it >> 1) & 1);
                tc_fence_after();
            }
            ATTN_TRACE(warp, u, 0);
            if (elect_one()) {
                const uint32_t q_addr = smem_u32(smem + stage * stage_bytes);
                const uint32_t k_addr = q_addr + kv_bytes;
#pragma unroll
                for (int k = 0; k < ATTN_DH / 16; ++k)
                    umma_f16(tmem_base + (u & 1) * p.kpad, desc_kmajor_sw128(q_addr + t * ATTN_Q_TILE_BYTES, k),
                             desc_kmajor_sw128(k_addr, k), idesc_s, k != 0);
                umma_commit(&s_full[u & 1]);
            }
            __syncwarp();
        };
        if (n_units > 0) issue_s(0);
        for (int u = 0; u < n_units; ++u) {
            if (u + 1 < n_units) issue_s(u + 1);
            mbar_wait(&p_full[u & 1], (u >> 1) & 1);
            if (u > 0) mbar_wait(o_free, (u - 1) & 1);
            tc_fence_after();
            ATTN_TRACE(warp, u, 1);
            const int it = u / nqt, t = u - it * nqt, stage = it & 1;
            if (elect_one()) {
                const uint32_t v_addr = smem_u32(smem + stage * stage_bytes) + 2 * kv_bytes;
                const uint32_t pbase = tmem_base + (u & 1) * p.kpad;
                for (int ks = 0; ks < nch; ++ks) {  // keys [16 ks, 16 ks + 16): P sits in the first half of its part's range
                    const int c0 = ks >= c0_2 ? c0_2 : (ks >= c0_1 ? c0_1 : 0);
                    umma_f16_ts(tmem_base + o_col, pbase + 16 * c0 + 8 * (ks - c0), desc_mnmajor_sw128(v_addr, ks), idesc_o, ks != 0);
                }
                umma_commit(o_full);
                if (t == nqt - 1) umma_commit(&stage_free[stage]);  // the item's last unit: Q, K, V no longer needed
            }
            __syncwarp();
        }
    } else if (warp < ATTN3_EXP_WARPS) {
        // ------------------------------------------------------------ exponential warps
        const int part = warp >> 2;
        const int quarter = warp & 3;       // TMEM lane quarter (warp % 4)
        const int row = quarter * 32 + lane;
        const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
        const int ch0 = part == 0 ? 0 : (part == 1 ? c0_1 : c0_2);
        const int ch1 = part == 0 ? c0_1 : (part == 1 ? c0_2 : nch);
        const int nsteps = ch1 - ch0;
        // reference columns of the single-pass softmax: two of the last eight of part 0's range (never written with P)
        const int mcol = (nch == 1 && p.tokens <= 8) ? 0 : 16 * c0_1 - 8;
        // The stream over this part's chunks:  tcgen05.ld -> FFMA2 -> ex2 -> FADD2 / pack -> tcgen05.st, the next
        // chunk's load in flight while the current one is processed; `first` already holds (or is about to
        // receive) chunk ch0.  P is written in place over the first half of the part's own range.  `before_last`
        // runs just before the last chunk is processed (the other buffer is free by then).  Returns the row sum.
        auto exp_stream = [&](uint32_t taddr, float moff, uint32_t* first, uint32_t* second, auto before_last) -> float {
            const float2 sc2 = splat2(p.scale_log2), mo2 = splat2(moff);
            float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};  // independent chains, packed adds
            const uint32_t pbase = taddr + 16 * ch0;
            auto exp_step = [&](const uint32_t* v, int i) {   // chunk ch0 + i of this part, all 16 keys valid
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 a = fma2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc2, mo2);
                    const float2 e = make_float2(fast_exp2(a.x), fast_exp2(a.y));
                    acc2[j & 1] = add2(acc2[j & 1], e);
                    packed[j] = pack2<__nv_bfloat16>(e.x, e.y);
                }
                tmem_st_x8p(pbase + 8 * i, packed);   // in place: columns of this part that it has already read
            };
            auto exp_step_ragged = [&](const uint32_t* v, int i) {   // the row's last chunk (at 197 tokens: 5 valid keys)
                const int base = (ch0 + i) * 16;
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k0 = base + 2 * j;
                    float e0 = 0.f, e1 = 0.f;
                    if (k0 < p.tokens) e0 = fast_exp2(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, moff));
                    if (k0 + 1 < p.tokens) e1 = fast_exp2(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, moff));
                    acc2[j & 1] = add2(acc2[j & 1], make_float2(e0, e1));
                    packed[j] = pack2<__nv_bfloat16>(e0, e1);
                }
                tmem_st_x8p(pbase + 8 * i, packed);
            };
            // only the row's very last chunk can be ragged: it is peeled off the unrolled stream (one or two
            // instances of the masked code instead of one per step -- the stream must stay instruction-cache friendly)
            const int nfull = nsteps - ((ch1 == nch && (p.tokens & 15) != 0) ? 1 : 0);
#pragma unroll
            for (int i = 0; i < 6; i += 2) {
                if (i < nfull) {
                    tmem_ld_wait();
                    if (i + 1 < nsteps) tmem_ld_x16p(pbase + (i + 1) * 16, second);
                    else before_last();
                    exp_step(first, i);
                }
                if (i + 1 < nfull) {
                    tmem_ld_wait();
                    if (i + 2 < nsteps) tmem_ld_x16p(pbase + (i + 2) * 16, first);
                    else before_last();
                    exp_step(second, i + 1);
                }
            }
            if (nfull < nsteps) {
                tmem_ld_wait();
                before_last();
                if (nfull & 1) exp_step_ragged(second, nfull);
                else exp_step_ragged(first, nfull);
            }
            return (acc2[0].x + acc2[0].y) + (acc2[1].x + acc2[1].y);
        };
        auto unit_active = [&](int u) { return ((nqt == 2 ? (u & 1) : 0) * 128 + quarter * 32) < p.tokens && nsteps > 0; };
        if constexpr (EXACT) {
            for (int u = 0; u < n_units; ++u) {
                const uint32_t taddr = tmem_base + lane_bits + (u & 1) * p.kpad;
                float* my_sum = xsum + ((u % 3) * 3 + part) * 128 + row;
                const bool quarter_active = ((nqt == 2 ? (u & 1) : 0) * 128 + quarter * 32) < p.tokens;
                ATTN_TRACE(warp, u, 0);
                mbar_wait(&s_full[u & 1], (u >> 1) & 1);
                tc_fence_after();
                ATTN_TRACE(warp, u, 1);
                if (quarter_active) {   // a part without chunks still takes part in the maximum exchange
                    uint32_t ra[16], rb[16];
                    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    for (int c = ch0; c < ch1; ++c) {
                        uint32_t v[16];
                        tmem_ld_x16p(taddr + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c * 16 + j < p.tokens) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
                    }
                    if (nsteps > 0) tmem_ld_x16p(taddr + ch0 * 16, ra);  // pass 2's first load flies during the exchange
                    float* xm = xmax + (u & 1) * 3 * 128 + row;
                    xm[part * 128] = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                    attn_bar_sync<96>(1 + quarter);
                    const float mx = fmaxf(fmaxf(xm[0], xm[128]), xm[256]);
                    ATTN_TRACE(warp, u, 6);
                    *my_sum = exp_stream(taddr, -mx * p.scale_log2, ra, rb, [] {});
                    tmem_st_wait();
                    tc_fence_before();
                }
                ATTN_TRACE(warp, u, 2);
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[u & 1]);
            }
        } else {
            // Single pass.  (Requesting the next unit's first loads during the last chunk of this one -- software
            // pipelining across units -- was tried and LOST 5 %: the kernel is bound by TMEM read bandwidth,
            // ~64 B/clk/SM for the 142 KB a unit reads, not by the latency of the first round trip.)
            for (int u = 0; u < n_units; ++u) {
                ATTN_TRACE(warp, u, 0);
                mbar_wait(&s_full[u & 1], (u >> 1) & 1);
                tc_fence_after();
                ATTN_TRACE(warp, u, 1);
                if (unit_active(u)) {
                    uint32_t ra[16], rb[16];
                    const uint32_t taddr = tmem_base + lane_bits + (u & 1) * p.kpad;
                    tmem_ld_x2p(taddr + mcol, rb);
                    tmem_ld_x16p(taddr + ch0 * 16, ra);
                    tmem_ld_wait();
                    // reference exponent: the larger of TWO scores (columns mcol, mcol + 1; mcol < tokens always).  Any score of
                    // the row works as long as the row's maximum is within ~110 logits of it (checked through the row sum);
                    // eight reference columns per thread, as in round 1, cost 8 % more TMEM reads in a TMEM-read-bound kernel.
                    const float mx = (mcol + 1 < p.tokens) ? fmaxf(__uint_as_float(rb[0]), __uint_as_float(rb[1])) : __uint_as_float(rb[0]);
                    ATTN_TRACE(warp, u, 6);
                    xsum[((u % 3) * 3 + part) * 128 + row] = exp_stream(taddr, fmaf(-mx, p.scale_log2, -ATTN_FAST_SHIFT), ra, rb, [] {});
                    tmem_st_wait();
                    tc_fence_before();
                } else if (((nqt == 2 ? (u & 1) : 0) * 128 + quarter * 32) < p.tokens) {
                    xsum[((u % 3) * 3 + part) * 128 + row] = 0.f;   // a part without chunks in a live quarter
                }
                ATTN_TRACE(warp, u, 2);
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[u & 1]);
            }
        }
    } else if (warp < ATTN3_W_PRODUCER) {
        // ------------------------------------------------------------ output warps
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t oaddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + o_col;
        uint8_t* my_stage = sO + (warp - ATTN3_W_OUT) * 2 * 4096;
        uint32_t n_stores = 0;
        int u = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int img = item / 12, head = item - img * 12;
            for (int t = 0; t < nqt; ++t, ++u) {
                const bool warp_active = (t * 128 + quarter * 32) < p.tokens;
                mbar_wait(o_full, u & 1);
                tc_fence_after();
                ATTN_TRACE(warp, u, 3);
                if (!warp_active) {
                    if (lane == 0) mbar_arrive(o_free);
                    continue;
                }
                // all three parts' row sums were written before their p_full arrivals, which the PV MMA behind
                // o_full waited for; read before o_free is released, so unit u + 3 cannot overwrite them earlier
                const float* ps = xsum + (u % 3) * 3 * 128 + row;
                const float row_sum = (ps[0] + ps[128]) + ps[256];
                if constexpr (!EXACT) {
                    if (t * 128 + row < p.tokens && !(row_sum < ATTN_FAST_SUM_MAX)) atomicOr(&g_status_flags, VIT_FLAG_ATTN_RANGE);  // also inf / NaN
                }
                const float inv_sum = fast_rcp(row_sum);
                uint32_t r0[32], r1[32];
                tmem_ld_x32(oaddr, r0);
                tmem_ld_x32(oaddr + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(o_free);
                    tma_store_wait_read<1>();   // the store that last used this staging buffer (two units ago) has read it
                }
                __syncwarp();
                ATTN_TRACE(warp, u, 4);
                uint8_t* sb = my_stage + (n_stores & 1) * 4096;
                uint8_t* srow = sb + lane * 128;
                const uint32_t sw = lane & 7;
                const float2 inv2 = splat2(inv_sum);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t x[4], y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 a = mul2(make_float2(__uint_as_float(r0[8 * j + 2 * q]), __uint_as_float(r0[8 * j + 2 * q + 1])), inv2);
                        const float2 b = mul2(make_float2(__uint_as_float(r1[8 * j + 2 * q]), __uint_as_float(r1[8 * j + 2 * q + 1])), inv2);
                        x[q] = pack2<T>(a.x, a.y);
                        y[q] = pack2<T>(b.x, b.y);
                    }
                    *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<uint4*>(srow + (((4 + j) ^ sw) << 4)) = make_uint4(y[0], y[1], y[2], y[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {   // rows past the image's last token are clipped by the 3-D map
                    tma_store_3d(&tmap_out32, sb, head * ATTN_DH, t * 128 + quarter * 32, img);
                    tma_store_commit();
                }
                ++n_stores;
                ATTN_TRACE(warp, u, 5);
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == ATTN3_W_ISSUER) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vit

namespace vit {

// =============================================================================================
// Key-blocked kernel for longer sequences (224 < tokens <= 640; ViT-B/16 at 384^2: 577 tokens).
//
// One CTA per SM loops over (image, head) units.  The head's whole K and V (ceil(tokens/64) blocks of
// 64 keys x 128 B, <= 80 KB each) stay in shared memory for all of the head's 128-row query tiles,
// which the CTA's two softmax warpgroups take alternately (warpgroup w: tiles w, w+2, ...), each
// with its own Q buffer, MMA-issuer warp and TMEM columns:
//     S block buffers [0,64) [64,128) (double buffered: the MMA of block g+1 runs under the softmax
//     of block g), O [128,192); warpgroup 1 at +192.
// A row of S (up to 640 fp32) does not fit in TMEM next to O, so keys go in blocks of 64, and to
// keep O free of rescaling the tile makes TWO passes over its key blocks:
//     pass A   S_j = Q K_j^T  ->  running row maximum (exact)
//     pass B   S_j again      ->  P_j = exp2((S_j - max) / 8 * log2 e) written in place (bf16, 32 columns)
//                             ->  O += P_j V_j   (A operand from TMEM, V_j MN-major from shared memory)
// i.e. the reference's max / exp / sum / divide (ViT_seq.c:178-191) block by block, the division
// applied once to O.  QK^T is computed twice -- the tensor pipe has the slack (TMEM reads at
// ~64 B/clk/SM bound this kernel, profiles/r1_attention_trace.md); a single-pass variant with a
// lazy power-of-two rescale of O is the planned successor.
//
//   warps 0-3 / 4-7  softmax + output warpgroups 0 / 1 (one thread per query row)
//   warp 8  TMA producer      warps 9, 10  MMA issuers of warpgroup 0 / 1 (warp 9 owns TMEM)
constexpr int ATTNL_THREADS = 11 * 32;
constexpr int ATTNL_MAX_TOKENS = 640;
constexpr int ATTNL_KB = 64;                        // keys per block
constexpr int ATTNL_BLOCK_BYTES = ATTNL_KB * 128;   // one K or V block in shared memory
__host__ __device__ inline int attnl_blocks(int tokens) { return (tokens + ATTNL_KB - 1) / ATTNL_KB; }
__host__ inline int attnl_smem_bytes(int tokens) {
    return 2 * attnl_blocks(tokens) * ATTNL_BLOCK_BYTES + 2 * ATTN_Q_TILE_BYTES + 2 * ATTN2_OSTAGE_BYTES + 256 + 1024;
}

template <typename T>
__global__ void __launch_bounds__(ATTNL_THREADS, 1)
attention_sm100_blocked_kernel(const __grid_constant__ CUtensorMap tmap_qkv /* box {64, 64 rows} */,
                               const __grid_constant__ CUtensorMap tmap_out /* 3-D, box {64, 128, 1} */, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nb = attnl_blocks(p.tokens);          // key blocks per head
    const int nq = (p.tokens + 127) / 128;          // query tiles per head
    const int last_cols = ((p.tokens - (nb - 1) * ATTNL_KB) + 15) & ~15;  // S columns of the last key block
    uint8_t* sK = smem;
    uint8_t* sV = sK + nb * ATTNL_BLOCK_BYTES;
    uint8_t* sQ = sV + nb * ATTNL_BLOCK_BYTES;      // [2 warpgroups] 128 x 128 B
    uint8_t* sO = sQ + 2 * ATTN_Q_TILE_BYTES;       // [2 warpgroups] output staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * ATTN2_OSTAGE_BYTES);
    uint64_t* kv_full = bars;          // K, V of a head landed (tx)
    uint64_t* kv_free = bars + 1;      // both issuers are through with the head (2 arrivals)
    uint64_t* q_full = bars + 2;       // [wg]      Q tile landed (tx)
    uint64_t* q_free = bars + 4;       // [wg]      every MMA reading the Q tile has completed
    uint64_t* s_full = bars + 6;       // [wg][buf] S block in TMEM
    uint64_t* s_free = bars + 10;      // [wg][buf] pass A has read the block (128 arrivals)
    uint64_t* p_full = bars + 14;      // [wg][buf] pass B has written P over the block (128 arrivals)
    uint64_t* o_full = bars + 18;      // [wg]      O complete in TMEM
    uint64_t* o_free = bars + 20;      // [wg]      O drained (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_units = p.batch * 12;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_out);
        mbar_init(kv_full, 1);
        mbar_init(kv_free, 2);
        for (int w = 0; w < 2; ++w) {
            mbar_init(&q_full[w], 1);
            mbar_init(&q_free[w], 1);
            mbar_init(&o_full[w], 1);
            mbar_init(&o_free[w], 128);
            for (int b = 0; b < 2; ++b) {
                mbar_init(&s_full[w * 2 + b], 1);
                mbar_init(&s_free[w * 2 + b], 128);
                mbar_init(&p_full[w * 2 + b], 128);
            }
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
            uint32_t heads = 0, qcnt[2] = {0, 0};
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++heads) {
                const int img = unit / 12, head = unit - img * 12;
                const int row0 = img * p.tokens;
                mbar_wait(kv_free, (heads & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full, 2 * nb * ATTNL_BLOCK_BYTES);
                for (int j = 0; j < nb; ++j)
                    tma_load_2d(sK + j * ATTNL_BLOCK_BYTES, &tmap_qkv, kv_full, ATTN_DIM + head * ATTN_DH, row0 + j * ATTNL_KB);
                for (int j = 0; j < nb; ++j)
                    tma_load_2d(sV + j * ATTNL_BLOCK_BYTES, &tmap_qkv, kv_full, 2 * ATTN_DIM + head * ATTN_DH, row0 + j * ATTNL_KB);
                for (int t = 0; t < nq; ++t) {
                    const int w = t & 1;
                    mbar_wait(&q_free[w], (qcnt[w] & 1) ^ 1);
                    ++qcnt[w];
                    mbar_arrive_expect_tx(&q_full[w], ATTN_Q_TILE_BYTES);
                    tma_load_2d(sQ + w * ATTN_Q_TILE_BYTES, &tmap_qkv, &q_full[w], head * ATTN_DH, row0 + t * 128);
                    tma_load_2d(sQ + w * ATTN_Q_TILE_BYTES + ATTNL_BLOCK_BYTES, &tmap_qkv, &q_full[w], head * ATTN_DH, row0 + t * 128 + 64);
                }
                const int next = unit + static_cast<int>(gridDim.x);  // pull the next head's K, V towards L2
                if (next < n_units) {
                    const int img2 = next / 12, head2 = next - img2 * 12;
                    for (int j = 0; j < nb; ++j) {
                        tma_prefetch_l2_2d(&tmap_qkv, ATTN_DIM + head2 * ATTN_DH, img2 * p.tokens + j * ATTNL_KB);
                        tma_prefetch_l2_2d(&tmap_qkv, 2 * ATTN_DIM + head2 * ATTN_DH, img2 * p.tokens + j * ATTNL_KB);
                    }
                }
            }
        }
    } else if (warp >= 9) {
        // ------------------------------------------------------------ MMA issuer of warpgroup w
        const int w = warp - 9;
        const uint32_t tw = tmem_base + w * 192;
        const uint32_t idesc_o = make_idesc<__nv_bfloat16>(128, ATTN_DH, 0, 1);  // P, V are always bf16
        const uint32_t q_addr = smem_u32(sQ + w * ATTN_Q_TILE_BYTES), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
        uint32_t n_sfree[2] = {0, 0}, n_pfull[2] = {0, 0}, n_qfull = 0, n_ofree = 0, heads = 0;
        auto cols_of = [&](int j) { return j == nb - 1 ? last_cols : ATTNL_KB; };
        auto issue_s = [&](int g) {  // S block g of the tile's stream (g < nb: pass A, else pass B), key block g % nb
            const int j = g < nb ? g : g - nb, b = g & 1;
            if (elect_one()) {
                const uint32_t idesc_s = make_idesc<T>(128, static_cast<uint32_t>(cols_of(j)), 0, 0);
#pragma unroll
                for (int k = 0; k < ATTN_DH / 16; ++k)
                    umma_f16(tw + b * 64, desc_kmajor_sw128(q_addr, k), desc_kmajor_sw128(k_addr + j * ATTNL_BLOCK_BYTES, k), idesc_s, k != 0);
                umma_commit(&s_full[w * 2 + b]);
            }
            __syncwarp();
        };
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++heads) {
            mbar_wait(kv_full, heads & 1);
            for (int t = w; t < nq; t += 2) {
                mbar_wait(&q_full[w], n_qfull & 1);
                ++n_qfull;
                tc_fence_after();
                // the stream of 2 nb S blocks alternates between the two buffers; block g may be issued once
                // the previous user of its buffer is done: a pass-A read (s_free), or a pass-B P block, whose PV
                // MMA is issued just before it below (the tensor pipe keeps issue order)
                issue_s(0);
                issue_s(1);
                for (int g = 2; g < nb + 2; ++g) {
                    const int b = g & 1;
                    mbar_wait(&s_free[w * 2 + b], n_sfree[b] & 1);
                    ++n_sfree[b];
                    tc_fence_after();
                    issue_s(g);
                }
                for (int j = 0; j < nb; ++j) {
                    const int g = nb + j, b = g & 1;
                    mbar_wait(&p_full[w * 2 + b], n_pfull[b] & 1);
                    ++n_pfull[b];
                    if (j == 0) {  // O of this warpgroup's previous tile has been drained
                        mbar_wait(&o_free[w], (n_ofree & 1) ^ 1);
                        ++n_ofree;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const int ks_n = cols_of(j) / 16;
                        for (int ks = 0; ks < ks_n; ++ks)
                            umma_f16_ts(tw + 128, tw + b * 64 + ks * 8, desc_mnmajor_sw128(v_addr + j * ATTNL_BLOCK_BYTES, ks), idesc_o,
                                        (j | ks) != 0);
                    }
                    __syncwarp();
                    if (g + 2 < 2 * nb) issue_s(g + 2);
                }
                if (elect_one()) {
                    umma_commit(&o_full[w]);
                    umma_commit(&q_free[w]);
                    if (t + 2 >= nq) umma_commit(kv_free);  // this warpgroup's last tile of the head
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------ softmax / output warpgroups
        const int w = warp >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tw = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + w * 192;
        const bool storer = quarter == 0 && lane == 0;
        uint32_t n_sfull[2] = {0, 0}, n_ofull = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int img = unit / 12, head = unit - img * 12;
            for (int t = w; t < nq; t += 2) {
                // ---- pass A: exact row maximum
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                for (int g = 0; g < nb; ++g) {
                    const int b = g & 1, key0 = g * ATTNL_KB;
                    const int cols = g == nb - 1 ? last_cols : ATTNL_KB;
                    mbar_wait(&s_full[w * 2 + b], n_sfull[b] & 1);
                    ++n_sfull[b];
                    tc_fence_after();
                    for (int c = 0; c < cols; c += 16) {
                        uint32_t v[16];
                        tmem_ld_x16p(tw + b * 64 + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (key0 + c + i < p.tokens) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
                    }
                    tc_fence_before();
                    mbar_arrive(&s_free[w * 2 + b]);
                }
                const float moff = -fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
                // ---- pass B: exponentials, P in place, row sum
                float sum4[4] = {0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < nb; ++j) {
                    const int g = nb + j, b = g & 1, key0 = j * ATTNL_KB;
                    const int cols = j == nb - 1 ? last_cols : ATTNL_KB;
                    mbar_wait(&s_full[w * 2 + b], n_sfull[b] & 1);
                    ++n_sfull[b];
                    tc_fence_after();
                    for (int c = 0; c < cols; c += 16) {  // P chunk (8 columns) lands on S columns already read
                        uint32_t v[16], packed[8];
                        tmem_ld_x16p(tw + b * 64 + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int k0 = key0 + c + 2 * i;
                            float e0 = 0.f, e1 = 0.f;
                            if (k0 < p.tokens) e0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, moff));
                            if (k0 + 1 < p.tokens) e1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, moff));
                            sum4[i & 3] += e0 + e1;
                            packed[i] = pack2<__nv_bfloat16>(e0, e1);
                        }
                        tmem_st_x8p(tw + b * 64 + (c >> 1), packed);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&p_full[w * 2 + b]);
                }
                const float inv_sum = fast_rcp((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
                // ---- output: O / sum -> staging tile -> one TMA store (clipped at the image's last token)
                if (storer) tma_store_wait_read<0>();   // the previous tile's store has left the staging tile
                mbar_wait(&o_full[w], n_ofull & 1);
                ++n_ofull;
                tc_fence_after();
                uint32_t r0[32], r1[32];
                tmem_ld_x32(tw + 128, r0);
                tmem_ld_x32(tw + 160, r1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&o_free[w]);
                attn_bar_sync<128>(1 + w);              // ... as everybody in the warpgroup now knows
                uint8_t* srow = sO + w * ATTN2_OSTAGE_BYTES + row * 128;
                const uint32_t sw = lane & 7;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t x[4], y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        x[q] = pack2<T>(__uint_as_float(r0[8 * j + 2 * q]) * inv_sum, __uint_as_float(r0[8 * j + 2 * q + 1]) * inv_sum);
                        y[q] = pack2<T>(__uint_as_float(r1[8 * j + 2 * q]) * inv_sum, __uint_as_float(r1[8 * j + 2 * q + 1]) * inv_sum);
                    }
                    *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<uint4*>(srow + (((4 + j) ^ sw) << 4)) = make_uint4(y[0], y[1], y[2], y[3]);
                }
                fence_proxy_async_smem();
                attn_bar_sync<128>(1 + w);
                if (storer) {
                    tma_store_3d(&tmap_out, sO + w * ATTN2_OSTAGE_BYTES, head * ATTN_DH, t * 128, img);
                    tma_store_commit();
                }
            }
        }
        if (storer) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vit

namespace vit {

// =============================================================================================
// Streaming kernel for longer sequences (224 < tokens <= 640; ViT-B/16 at 384^2: 577 tokens), the
// key-blocked sibling of attention_sm100_stream_kernel and the engine's default for that range.
//
// With the single-pass softmax the exponent offset of a row is FIXED once it has been read from the row's
// first key block (eight scores of columns that never receive P), so the row's P blocks need no rescaling
// and O simply accumulates over the key blocks: one pass over the keys, Q K^T computed once -- the exact
// two-pass kernel above computes it twice and reads every S block twice.  Same roles as the single-block
// streaming kernel: 12 exponential warps (three threads per row, contiguous chunk ranges inside each
// key block, P in place), 4 output warps, TMA producer, one MMA issuer; S is double buffered per KEY
// BLOCK (2 x up to 224 columns), O (64 columns) accumulates over a unit's blocks.  The head's whole K and V stay in
// shared memory for all of its query tiles.
//
//   issue order:  S(g+1) | PV(g)  per key block g of the unit stream (S(g+1) overwrites the buffer whose P
//   was consumed by PV(g-1), already issued).  K and V are single buffered but released separately: K once the
//   head's last S has completed (its reload for the next head runs under the head's last exponentials and PVs),
//   V once its last PV has (its reload runs under the next head's first block of exponentials).
//
// Rows whose scores leave the exponent window raise VIT_FLAG_ATTN_RANGE exactly as in the single-block kernel;
// the host then repeats the call with attention_sm100_blocked_kernel (exact).
constexpr int ATTN4_THREADS = 18 * 32;
// keys per S block: the row is cut into the fewest blocks of at most 224 keys (2 S buffers + O fit the 512 TMEM
// columns), equally sized up to the 16-key chunk (577 tokens: 208 + 208 + 161)
// -- and K, V of the head (rounded up to the 64-row TMA boxes) must fit 640 rows of shared memory each
// (640 tokens: four blocks of 160 rather than three of 224)
__host__ __device__ inline int attn4_kb_for(int tokens, int nkb) { return ((tokens + nkb - 1) / nkb + 15) / 16 * 16; }
__host__ __device__ inline int attn4_blocks(int tokens) {
    int nkb = (tokens + 223) / 224;
    while ((nkb * attn4_kb_for(tokens, nkb) + 63) / 64 * 64 > 640) ++nkb;
    return nkb;
}
__host__ __device__ inline int attn4_kb(int tokens) { return attn4_kb_for(tokens, attn4_blocks(tokens)); }
constexpr int ATTN4_OSTAGE_BYTES = 4 * 4096;          // [output warp] 32 rows x 128 B
constexpr int ATTN4_XCH_BYTES = 2 * 3 * 128 * 4;      // row sums [unit parity][part][row]
__host__ __device__ inline int attn4_kv_rows(int tokens) { return (attn4_blocks(tokens) * attn4_kb(tokens) + 63) / 64 * 64; }  // 64-row TMA boxes
__host__ inline int attn4_smem_bytes(int tokens) {
    return 2 * attn4_kv_rows(tokens) * 128 + 2 * ATTN_Q_TILE_BYTES + ATTN4_OSTAGE_BYTES + ATTN4_XCH_BYTES + 256 + 1024;
}

template <typename T>
__global__ void __launch_bounds__(ATTN4_THREADS, 1)
attention_sm100_stream_blocked_kernel(const __grid_constant__ CUtensorMap tmap_qkv /* box {64, 64 rows} */,
                                      const __grid_constant__ CUtensorMap tmap_out32 /* 3-D, box {64, 32, 1} */, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkb = attn4_blocks(p.tokens);          // key blocks per head
    const int nq = (p.tokens + 127) / 128;           // query tiles (units) per head
    const int kb = attn4_kb(p.tokens);               // keys per block
    const int last_cols = ((p.tokens - (nkb - 1) * kb) + 15) & ~15;  // S columns of the last key block
    const int kv_boxes = attn4_kv_rows(p.tokens) / 64;
    const int kv_bytes = kv_boxes * 64 * 128;
    uint8_t* sK = smem;
    uint8_t* sV = sK + kv_bytes;
    uint8_t* sQ = sV + kv_bytes;                     // [2] 128 x 128 B
    uint8_t* sO = sQ + 2 * ATTN_Q_TILE_BYTES;        // [output warp] staging
    float* xsum = reinterpret_cast<float*>(sO + ATTN4_OSTAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xsum) + ATTN4_XCH_BYTES);
    uint64_t* k_full = bars;         // K of a head landed (tx)
    uint64_t* k_free = bars + 1;     // every S MMA of the head has completed
    uint64_t* q_full = bars + 2;     // [2] Q tile landed (tx)
    uint64_t* q_free = bars + 4;     // [2] every S MMA of the unit has completed
    uint64_t* s_full = bars + 6;     // [2] S block in TMEM
    uint64_t* p_full = bars + 8;     // [2] P written back (one arrival per exponential warp)
    uint64_t* o_full = bars + 10;    // O of a unit complete
    uint64_t* o_free = bars + 11;    // O drained (one arrival per output warp)
    uint64_t* v_full = bars + 12;    // V of a head landed (tx)
    uint64_t* v_free = bars + 13;    // every PV MMA of the head has completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.batch * 12;
    const int my_items = blockIdx.x < n_items ? (n_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    const int n_units = my_items * nq;
    const uint32_t o_col = 2 * kb;
    // chunk ranges of the three parts inside a key block of `nch` 16-key chunks: sizes differ by at most one
    auto part_lo = [](int nch, int part) {
        const int sz = nch / 3, rem = nch - 3 * sz;
        return part == 0 ? 0 : (part == 1 ? sz + (rem > 0) : 2 * sz + (rem > 0) + (rem > 1));
    };

    if (warp == ATTN3_W_PRODUCER && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_out32);
        mbar_init(k_full, 1);
        mbar_init(k_free, 1);
        mbar_init(v_full, 1);
        mbar_init(v_free, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_free[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], ATTN3_EXP_WARPS);
        }
        mbar_init(o_full, 1);
        mbar_init(o_free, 4);
        fence_barrier_init();
    }
    if (warp == ATTN3_W_ISSUER) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();
    griddep_launch();

    if (warp == ATTN3_W_PRODUCER) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0, uq = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int img = item / 12, head = item - img * 12;
                const int row0 = img * p.tokens;
                auto load_q = [&](int t) {
                    const int b = uq & 1;
                    mbar_wait(&q_free[b], ((uq >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&q_full[b], ATTN_Q_TILE_BYTES);
                    tma_load_2d(sQ + b * ATTN_Q_TILE_BYTES, &tmap_qkv, &q_full[b], head * ATTN_DH, row0 + t * 128);
                    tma_load_2d(sQ + b * ATTN_Q_TILE_BYTES + 64 * 128, &tmap_qkv, &q_full[b], head * ATTN_DH, row0 + t * 128 + 64);
                    ++uq;
                };
                // K first (free since the previous head's last S), then the first Q tile, so that the head's first S can be
                // issued without waiting for V, whose buffer the previous head's last PVs are still reading
                mbar_wait(k_free, (it & 1) ^ 1);
                mbar_arrive_expect_tx(k_full, kv_bytes);
                for (int j = 0; j < kv_boxes; ++j)
                    tma_load_2d(sK + j * 64 * 128, &tmap_qkv, k_full, ATTN_DIM + head * ATTN_DH, row0 + j * 64);
                load_q(0);
                mbar_wait(v_free, (it & 1) ^ 1);
                mbar_arrive_expect_tx(v_full, kv_bytes);
                for (int j = 0; j < kv_boxes; ++j)
                    tma_load_2d(sV + j * 64 * 128, &tmap_qkv, v_full, 2 * ATTN_DIM + head * ATTN_DH, row0 + j * 64);
                for (int t = 1; t < nq; ++t) load_q(t);
                const int next = item + static_cast<int>(gridDim.x);  // pull the next head's K, V towards L2
                if (next < n_items) {
                    const int img2 = next / 12, head2 = next - img2 * 12;
                    for (int j = 0; j < kv_boxes; ++j) {
                        tma_prefetch_l2_2d(&tmap_qkv, ATTN_DIM + head2 * ATTN_DH, img2 * p.tokens + j * 64);
                        tma_prefetch_l2_2d(&tmap_qkv, 2 * ATTN_DIM + head2 * ATTN_DH, img2 * p.tokens + j * 64);
                    }
                }
            }
        }
    } else if (warp == ATTN3_W_ISSUER) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc_o = make_idesc<__nv_bfloat16>(128, ATTN_DH, 0, 1);  // P, V are always bf16
        const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
        const int n_blocks = n_units * nkb;
        auto issue_s = [&](int g) {   // S block g of the stream: unit g / nkb, key block g % nkb, buffer g & 1
            const int u = g / nkb, j = g - u * nkb;
            if (j == 0) {
                if (u % nq == 0) mbar_wait(k_full, (u / nq) & 1);
                mbar_wait(&q_full[u & 1], (u >> 1) & 1);
                tc_fence_after();
            }
            if (elect_one()) {
                const uint32_t idesc_s = make_idesc<T>(128, static_cast<uint32_t>(j == nkb - 1 ? last_cols : kb), 0, 0);
                const uint32_t q_addr = smem_u32(sQ + (u & 1) * ATTN_Q_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATTN_DH / 16; ++k)
                    umma_f16(tmem_base + (g & 1) * kb, desc_kmajor_sw128(q_addr, k), desc_kmajor_sw128(k_addr + j * kb * 128, k), idesc_s, k != 0);
                umma_commit(&s_full[g & 1]);
                if (j == nkb - 1) {
                    umma_commit(&q_free[u & 1]);                 // the unit's last S: its Q tile may be replaced
                    if (u % nq == nq - 1) umma_commit(k_free);   // ... and the head's last: so may K
                }
            }
            __syncwarp();
        };
        if (n_blocks > 0) issue_s(0);
        for (int g = 0; g < n_blocks; ++g) {
            const int u = g / nkb, j = g - u * nkb;
            const bool head_ends = j == nkb - 1 && (u % nq) == nq - 1;
            if (g + 1 < n_blocks) issue_s(g + 1);
            mbar_wait(&p_full[g & 1], (g >> 1) & 1);
            if (j == 0) {
                if (u > 0) mbar_wait(o_free, (u - 1) & 1);
                if (u % nq == 0) mbar_wait(v_full, (u / nq) & 1);
            }
            tc_fence_after();
            if (elect_one()) {
                const int nch = (j == nkb - 1 ? last_cols : kb) / 16;
                const int lo1 = part_lo(nch, 1), lo2 = part_lo(nch, 2);
                const uint32_t pbase = tmem_base + (g & 1) * kb;
                for (int ks = 0; ks < nch; ++ks) {
                    const int c0 = ks >= lo2 ? lo2 : (ks >= lo1 ? lo1 : 0);
                    umma_f16_ts(tmem_base + o_col, pbase + 16 * c0 + 8 * (ks - c0), desc_mnmajor_sw128(v_addr + j * kb * 128, ks), idesc_o,
                                (j | ks) != 0);
                }
                if (j == nkb - 1) {
                    umma_commit(o_full);
                    if (head_ends) umma_commit(v_free);
                }
            }
            __syncwarp();
        }
    } else if (warp < ATTN3_EXP_WARPS) {
        // ------------------------------------------------------------ exponential warps
        const int part = warp >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
        const int mcol = 16 * part_lo(nkb > 1 ? kb / 16 : last_cols / 16, 1) - 8;  // last eight columns of part 0's range in block 0
        int g = 0;
        for (int u = 0; u < n_units; ++u) {
            const int t = u % nq;
            const bool warp_active = (t * 128 + quarter * 32) < p.tokens;
            float moff = 0.f;
            float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            for (int j = 0; j < nkb; ++j, ++g) {
                const int nch = (j == nkb - 1 ? last_cols : kb) / 16;
                const int ch0 = part_lo(nch, part), ch1 = part == 2 ? nch : part_lo(nch, part + 1);
                const int nsteps = ch1 - ch0;
                const int key0 = j * kb;
                const uint32_t taddr = tmem_base + lane_bits + (g & 1) * kb;
                mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc_fence_after();
                if (warp_active) {
                    uint32_t ra[16], rb[16];
                    if (j == 0) {
                        tmem_ld_x8p(taddr + mcol, rb);
                        if (nsteps > 0) tmem_ld_x16p(taddr + ch0 * 16, ra);
                        tmem_ld_wait();
                        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (mcol + i < p.tokens) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(rb[i]));
                        moff = fmaf(-fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), p.scale_log2, -ATTN_FAST_SHIFT);
                    } else if (nsteps > 0) {
                        tmem_ld_x16p(taddr + ch0 * 16, ra);
                    }
                    const float2 sc2 = splat2(p.scale_log2), mo2 = splat2(moff);
                    const uint32_t pbase = taddr + 16 * ch0;
                    auto exp_step = [&](const uint32_t* v, int i) {
                        uint32_t packed[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float2 a = fma2(make_float2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), sc2, mo2);
                            const float2 e = make_float2(fast_exp2(a.x), fast_exp2(a.y));
                            acc2[q & 1] = add2(acc2[q & 1], e);
                            packed[q] = pack2<__nv_bfloat16>(e.x, e.y);
                        }
                        tmem_st_x8p(pbase + 8 * i, packed);
                    };
                    auto exp_step_ragged = [&](const uint32_t* v, int i) {   // the row's last chunk: keys past `tokens` get P = 0
                        const int base = key0 + (ch0 + i) * 16;
                        uint32_t packed[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int k0 = base + 2 * q;
                            float e0 = 0.f, e1 = 0.f;
                            if (k0 < p.tokens) e0 = fast_exp2(fmaf(__uint_as_float(v[2 * q]), p.scale_log2, moff));
                            if (k0 + 1 < p.tokens) e1 = fast_exp2(fmaf(__uint_as_float(v[2 * q + 1]), p.scale_log2, moff));
                            acc2[q & 1] = add2(acc2[q & 1], make_float2(e0, e1));
                            packed[q] = pack2<__nv_bfloat16>(e0, e1);
                        }
                        tmem_st_x8p(pbase + 8 * i, packed);
                    };
                    const int nfull = nsteps - ((j == nkb - 1 && ch1 == nch && (p.tokens & 15) != 0 && nsteps > 0) ? 1 : 0);
#pragma unroll
                    for (int i = 0; i < 6; i += 2) {
                        if (i < nfull) {
                            tmem_ld_wait();
                            if (i + 1 < nsteps) tmem_ld_x16p(pbase + (i + 1) * 16, rb);
                            exp_step(ra, i);
                        }
                        if (i + 1 < nfull) {
                            tmem_ld_wait();
                            if (i + 2 < nsteps) tmem_ld_x16p(pbase + (i + 2) * 16, ra);
                            exp_step(rb, i + 1);
                        }
                    }
                    if (nfull < nsteps) {
                        tmem_ld_wait();
                        if (nfull & 1) exp_step_ragged(rb, nfull);
                        else exp_step_ragged(ra, nfull);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                }
                if (j == nkb - 1 && warp_active) xsum[((u & 1) * 3 + part) * 128 + row] = (acc2[0].x + acc2[0].y) + (acc2[1].x + acc2[1].y);
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g & 1]);
            }
        }
    } else if (warp < ATTN3_W_PRODUCER) {
        // ------------------------------------------------------------ output warps
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t oaddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + o_col;
        uint8_t* sb = sO + (warp - ATTN3_W_OUT) * 4096;
        int u = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int img = item / 12, head = item - img * 12;
            for (int t = 0; t < nq; ++t, ++u) {
                const bool warp_active = (t * 128 + quarter * 32) < p.tokens;
                mbar_wait(o_full, u & 1);
                tc_fence_after();
                if (!warp_active) {
                    if (lane == 0) mbar_arrive(o_free);
                    continue;
                }
                const float* ps = xsum + (u & 1) * 3 * 128 + row;
                const float row_sum = (ps[0] + ps[128]) + ps[256];
                if (t * 128 + row < p.tokens && !(row_sum < ATTN_FAST_SUM_MAX)) atomicOr(&g_status_flags, VIT_FLAG_ATTN_RANGE);  // also inf / NaN
                const float inv_sum = fast_rcp(row_sum);
                uint32_t r0[32], r1[32];
                tmem_ld_x32(oaddr, r0);
                tmem_ld_x32(oaddr + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(o_free);
                    tma_store_wait_read<0>();   // the previous store has left the staging tile
                }
                __syncwarp();
                uint8_t* srow = sb + lane * 128;
                const uint32_t sw = lane & 7;
                const float2 inv2 = splat2(inv_sum);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint32_t x[4], y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 a = mul2(make_float2(__uint_as_float(r0[8 * jj + 2 * q]), __uint_as_float(r0[8 * jj + 2 * q + 1])), inv2);
                        const float2 b = mul2(make_float2(__uint_as_float(r1[8 * jj + 2 * q]), __uint_as_float(r1[8 * jj + 2 * q + 1])), inv2);
                        x[q] = pack2<T>(a.x, a.y);
                        y[q] = pack2<T>(b.x, b.y);
                    }
                    *reinterpret_cast<uint4*>(srow + ((jj ^ sw) << 4)) = make_uint4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<uint4*>(srow + (((4 + jj) ^ sw) << 4)) = make_uint4(y[0], y[1], y[2], y[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tmap_out32, sb, head * ATTN_DH, t * 128 + quarter * 32, img);
                    tma_store_commit();
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == ATTN3_W_ISSUER) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vit
