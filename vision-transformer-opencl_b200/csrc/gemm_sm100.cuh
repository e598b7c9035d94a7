// gemm_sm100.cuh -- C[M,N] = A[M,K] * W[N,K]^T with fused epilogues, for sm_100a.
//
// Replaces the reference's linear_layer (ViT_seq.c:240-250) and its OpenCL twins
// (linear_forward_kernel / fc1_kernel / fc2_kernel / MHA_gemm_kernel in kernel.cl) for the
// token-flattened batch: M = images * tokens.  W is the PyTorch [out,in] layout, i.e. already
// the K-major "B^T" operand.
//
// One kernel: gemm_sm100_staged_kernel -- CTA pair (cta_group::2, 256 x 256 tiles), staged TMA-store epilogue, fused bias /
// exact-erf GELU / fp32 residual, LayerNorm folded in (producer + consumer forms), per-image EMBED addressing for conv_proj.
// (Round 1's single-CTA and register-epilogue pair kernels were superseded by it and have been removed.)
// Structure (one persistent CTA per SM, warp specialised): a TMA producer (A tile 128x64 and W tile per
// stage, 128B swizzle), ONE thread issuing tcgen05.mma (kind::f16, fp32 accumulators in TMEM, two accumulator
// stages), tcgen05.commit releasing shared-memory stages and publishing accumulators, and epilogue warps reading
// the accumulators with tcgen05.ld while the next tile's main loop runs in the other TMEM stage.
#pragma once

#include "ptx.cuh"

namespace vit {

enum : int {
    EPI_BIAS = 0,           // out[T]   = acc + bias
    EPI_BIAS_GELU = 1,      // out[T]   = gelu(acc + bias)
    EPI_BIAS_RESIDUAL = 2   // out[f32] = residual + acc + bias   (residual may alias out)
};

struct GemmParams {
    int M, N, K;
    const float* bias;      // [N]
    void* out;              // row-major, leading dimension N
    const float* residual;  // EPI_BIAS_RESIDUAL: [M][N] fp32
    int patches;            // EMBED (conv_proj): patches per image (196 / 576)
    int tokens;             // EMBED (conv_proj): tokens per image  (197 / 577)
    int grid_w;             // EMBED (conv_proj): patches per image row (14 / 24)
    int bf16_from_col;      // staged EPI_BIAS: output columns >= this are stored as bf16 whatever T is
                            // (the V block of in_proj: attention keeps P and V in bf16); <= 0: never
    // ---- LayerNorm folded into the GEMMs (staged kernel, LN = true), see gemm_sm100_staged_kernel
    const float2* stats_in;   // consumer: per-row partial (sum, sum of squares) of the fp32 residual row, [parts][stats_rows]
    float2* stats_out;        // producer (EPI_BIAS_RESIDUAL): partials of the rows it writes, part = 2 * n_tile + column half
    int stats_parts;          // consumer: number of partials per row to add up (6 after a residual GEMM, 1 after rowstats_cast)
    int stats_rows;           // row capacity of the stats arrays (stride between parts)
    const float* colsum;      // consumer: s[n] = sum_k W'[n][k] of the folded, rounded weights W' = ln_w (.) W
    float* cls_rows32;        // RES16: [M / tokens][N] fp32 master copy of the class-token rows (row % tokens == 0), updated in place
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 x 16-bit = one 128-byte swizzle row
constexpr int GEMM_NON_EPI_WARPS = 4;

// =============================================================================================
// CTA-pair GEMM with a staged epilogue (the kernel the engine uses for conv_proj, in_proj, out_proj,
// mlp_0 and mlp_3).  Main loop: two CTAs of a cluster share one 256 x 256 output tile
// (tcgen05.mma.cta_group::2, UMMA M = 256); each CTA TMA-loads only its own 128 rows of A and its own
// 128 of the 256 W rows per K block (32 KB per stage), which cuts the L2 -> SM operand traffic per FLOP
// by a third against a single-CTA 128 x 256 tile.  The leader CTA issues all MMAs; completion is
// multicast to both CTAs' barriers; the peer's epilogue releases accumulator stages by remote
// mbarrier arrives.  The epilogue never touches global memory from registers (one thread per row =>
// 32 different cache lines per warp instruction, a latency-bound trickle, ncu profiles/r1_*): the
// 8 epilogue warps walk the 128 x 256 accumulator
// in column chunks that are exactly one 128-byte-per-row, 128B-swizzled shared-memory slot
// (64 operand-precision columns, or 32 fp32 columns), and
//   * results leave through TMA bulk stores issued by one thread per chunk;
//   * for the residual epilogue the fp32 residual chunk is TMA-loaded into the slot ahead of time
//     by a dedicated loader warp, updated in place, and stored back -- so both directions of the
//     residual stream move as full 128-byte rows, asynchronously, under the next tile's MMAs.
// Slots form a ring shared by loader, epilogue warps and the storing thread.
constexpr int GEMM_SLOT_BYTES = 128 * 128;
constexpr int EMBED_SUBTILE_BYTES = 128 * 64;   // conv_proj A operand: 128 rows of 16 tf32 (one kernel row of a patch)

// PRE_BYTES: per-tile parameters staged one tile ahead by the loader warp of the non-residual kernels
// (bias, and for the LayerNorm consumer the column sums and the 128 rows' (rstd, -rstd * mean)), two buffers.
template <int STAGES, int SLOTS, int CAST_BUFS = 0, int PRE_BYTES = 0>
struct GemmStagedSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = 128 * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SLOT_OFF = STAGES * STAGE_BYTES;
    static constexpr int CAST_OFF = SLOT_OFF + SLOTS * GEMM_SLOT_BYTES;   // operand-precision copy of the output (LN producer)
    static constexpr int BIAS_OFF = CAST_OFF + CAST_BUFS * GEMM_SLOT_BYTES;
    static constexpr int BIAS_BYTES = PRE_BYTES > 0 ? 2 * PRE_BYTES : 2 * 256 * 4;  // unstaged: [bias | colsum][256] of the current tile
    static constexpr int BAR_OFF = BIAS_OFF + BIAS_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4 + 3 * SLOTS + 4;
    static constexpr int DYN_BYTES = BAR_OFF + NUM_BARS * 8 + 16;  // base must be 1024-aligned (checked)
};

template <int NTHREADS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// -DVIT_GEMM_TRACE (tools/gemm_trace.sh; never in the product build): the first CTA's roles add up the SM cycles they spend
// in each kind of wait and print one line per launch.
#ifdef VIT_GEMM_TRACE
#define TR_DECL(n) long long tr_##n = 0
#define TR_WAIT(n, ...) do { const long long t0__ = clock64(); __VA_ARGS__; tr_##n += clock64() - t0__; } while (0)
#else
#define TR_DECL(n)
#define TR_WAIT(n, ...) __VA_ARGS__
#endif

// LayerNorm folded into the GEMMs (LN = true).  The reference normalises every token row before in_proj
// and mlp_0 (layer_norm, ViT_seq.c:103-121, called at :281 and :291).  Here no normalised activation is
// ever written: with W' = ln_w (.) W (folded and rounded once at init), s[n] = sum_k W'[n][k] and
// c[n] = bias[n] + sum_k ln_b[k] W[n][k],
//     LN(x) W^T + bias  =  rstd * (x W'^T - mean * s) + c ,
// so the CONSUMER GEMM (EPI_BIAS / EPI_BIAS_GELU, LN = true) multiplies the operand-precision copy of the
// raw residual row and applies (mean, rstd) per row in its epilogue, and the PRODUCER GEMM
// (EPI_BIAS_RESIDUAL, LN = true), which writes the fp32 residual row anyway, also emits that copy
// (tmap_cast) and the row's partial (sum, sum of squares) over its 256 columns.  Partials are plain stores
// to stats_out[2 * n_tile + column half][row]: no atomics, fixed summation order, results independent of
// the batch position.  This removes the LayerNorm kernels (read 3 KB + write 1.5 KB per token, twice
// per layer) from the forward pass.
//
// EPI_WARPS: 8 (two warps per TMEM lane quarter) or 16 (four per quarter; for the GELU epilogue, whose
// ~13 dependent FP32 ops + 2 MUFU per element are latency bound with only two warps per scheduler).
template <typename T, int STAGES, int SLOTS, int EPI, int EPI_WARPS, bool LN = false, int CAST_BUFS = (LN && EPI == EPI_BIAS_RESIDUAL) ? 2 : 0,
          bool STAGED = (EPI != EPI_BIAS_RESIDUAL), bool EMBED = false, bool RES16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((GEMM_NON_EPI_WARPS + EPI_WARPS) * 32, 1)
gemm_sm100_staged_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_cast,
                         const __grid_constant__ CUtensorMap tmap_res, const GemmParams p) {
    // EMBED (conv_proj, EPI_BIAS_RESIDUAL): the patch embedding WITHOUT an im2col buffer (Conv2d + flatten_transpose +
    // class_token + pos_emb, ViT_seq.c:25-101).  The A operand is the fp32 image itself, seen through a 5-D TMA map as
    // [plane = image * 3 + channel][gy][gx][ky][kx]: a box {kx 16, ky 1, gx G, gy R/G, 1 plane} lands in shared memory
    // as R = (128 / G) * G rows (one per patch, whole patch rows gy of the image) of 16 fp32 = 64 bytes, 64B-swizzled --
    // a K-major operand sub-tile of 16 tf32 K elements.  Two such boxes (kernel rows ky, ky + 1) make one K block of 32,
    // matching the 128-byte rows of the weight tile (K order (channel, ky, kx) = conv_proj.weight's row order); the A and B
    // descriptors of an MMA carry their own swizzle modes.  (Measured on the hardware, tools/embed_probe.py: ONE box
    // {16, 2, G, R/G} under the 128B swizzle does NOT give 128-byte rows -- a box whose inner extent is 64 bytes keeps a
    // 128-byte row pitch and leaves the second half of every row unwritten.)  The MMA is kind::tf32 on the raw pixels
    // (weights rounded to tf32 once at init).  A CTA's rows are consecutive patches [gy0 * G, gy0 * G + R) of ONE
    // image: row tiles never straddle images, the "residual" is pos_embedding rows 1 + patch (tmap_res, 2-D, shared by all
    // images), and the output goes to token row 1 + patch of the image through 3-D maps [image][token][768] with R-row
    // boxes that clip at the image's last token.  Rows R..127 of the tile hold stale shared memory; MMA rows are
    // independent and those rows are never stored.
    // p.M = images * tiles per image * 256; T is the type of the operand-precision copy (LN producer).
    static_assert(!EMBED || EPI == EPI_BIAS_RESIDUAL, "EMBED is a residual-epilogue variant");
    constexpr int BK_ELEMS = EMBED ? 32 : GEMM_BK;               // K elements per 128-byte operand row: tf32 / 16-bit
    const int embed_gyc = EMBED ? 128 / p.grid_w : 1;            // patch rows (gy) per CTA
    const int embed_rows = EMBED ? embed_gyc * p.grid_w : 128;   // R: live rows of this CTA's tile
    const int embed_tpi = EMBED ? (p.grid_w + 2 * embed_gyc - 1) / (2 * embed_gyc) : 1;  // pair tiles per image
    const uint32_t embed_rank = cluster_ctarank();
    auto embed_img = [&](int mt) { return mt / embed_tpi; };
    auto embed_gy0 = [&](int mt) { return ((mt % embed_tpi) * 2 + static_cast<int>(embed_rank)) * embed_gyc; };  // this CTA's first patch row
    auto embed_patch0 = [&](int mt) { return embed_gy0(mt) * p.grid_w; };                                         // ... and first patch
    // RES16 (EPI_BIAS_RESIDUAL, LN): the residual stream itself is held in the operand type T -- tmap_out is the 16-bit row
    // buffer, loaded, updated and stored in place in 64-column chunks; there is no fp32 row and no separate operand copy (the
    // row IS what the next GEMM multiplies), and the row statistics are those of the rounded values.  Moves 0.93 instead of
    // 1.86 GB per launch at 201 728 rows.
    static_assert(!RES16 || (LN && EPI == EPI_BIAS_RESIDUAL && !EMBED), "RES16 is a variant of the LayerNorm-producer residual epilogue");
    static_assert((CAST_BUFS > 0) == (LN && EPI == EPI_BIAS_RESIDUAL && !RES16) && CAST_BUFS <= 2, "staging tiles of the operand-precision copy");
    // STAGED: the tile's parameters are put into shared memory one tile ahead by the loader warp (two buffers);
    // otherwise the epilogue threads load them themselves, from lines they prefetched into L1 a tile earlier.
    static_assert(!(STAGED && EPI == EPI_BIAS_RESIDUAL), "the residual kernel's loader warp is busy");
    constexpr int PRE_FLOATS = !STAGED ? 0 : (LN ? 256 + 256 + 2 * 128 : 256);  // bias | colsum | (rstd, nm) per row
    using L = GemmStagedSmem<STAGES, SLOTS, CAST_BUFS, PRE_FLOATS * 4>;
    static_assert(EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RESIDUAL, "staged epilogues");
    constexpr bool kResidual = EPI == EPI_BIAS_RESIDUAL;
    constexpr int BN = 256;
    constexpr bool kResidual32 = kResidual && !RES16;    // fp32 residual rows: 32 columns per 128-byte row segment
    constexpr int CHUNK_COLS = kResidual32 ? 32 : 64;    // one 128-byte row segment per chunk
    constexpr int NCHUNK = BN / CHUNK_COLS;
    constexpr int PARTS = EPI_WARPS / 4;                 // warps sharing a TMEM lane quarter split the chunk's columns
    constexpr int COLS_PER_WARP = CHUNK_COLS / PARTS;
    constexpr int PIECES = COLS_PER_WARP * (kResidual32 ? 4 : 2) / 16;  // 16-byte pieces of the 128-byte row per thread
    constexpr int EPI_THREADS = EPI_WARPS * 32;
    static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "epilogue warps");
    static_assert(!(kResidual && EPI_WARPS == 16), "residual epilogue uses 8 warps");

    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_bias = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* res_full = tempty_bar + 2;     // [SLOTS] residual chunk landed in the slot
    uint64_t* slot_free = res_full + SLOTS;  // [SLOTS] the store out of the slot has drained
    uint64_t* chunk_ready = slot_free + SLOTS;  // [SLOTS] RES16: every epilogue warp has updated its part of the slot (EPI_WARPS arrivals)
    uint64_t* pre_full = chunk_ready + SLOTS;   // [2] the tile's staged parameters are in shared memory (32 arrivals)
    uint64_t* pre_free = pre_full + 2;       // [2] every epilogue warp is through with them
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pre_free + 2);

    // Warp roles.  The epilogue warps come FIRST: the SM's warp scheduler favours higher warp ids
    // among eligible warps, and the single-thread TMA producer / MMA issuer must never be starved
    // by eight warps of GELU arithmetic (measured: mlp_0 tensor-pipe activity 64 % -> see profiles/).
    constexpr int W_PRODUCER = EPI_WARPS, W_MMA = EPI_WARPS + 1, W_TMEM = EPI_WARPS + 2, W_LOADER = EPI_WARPS + 3;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int tiles_n = p.N / BN;
    const int tiles_m = (p.M + 255) / 256;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = p.K / BK_ELEMS;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // swizzle atoms need 1 KB alignment
    if (warp == W_PRODUCER && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_out);
        if constexpr (CAST_BUFS > 0 || !kResidual) tma_prefetch_desc(&tmap_cast);
        if constexpr (EMBED) tma_prefetch_desc(&tmap_res);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 2 * EPI_WARPS);  // one arrival per epilogue warp of both CTAs (used in the leader)
        }
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&slot_free[s], 1);
            mbar_init(&chunk_ready[s], EPI_WARPS);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&pre_full[s], kResidual ? 1 : 32);   // residual kernel with ONE cast staging tile: [0] = "tile drained"
            mbar_init(&pre_free[s], EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == W_TMEM) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();     // everything above overlaps the previous kernel's tail; nothing below may
    griddep_launch();

    if (warp == W_PRODUCER) {
        // ------------------------------------------------------------ operand producer (both CTAs)
        // The whole warp walks the loop (warp-uniform control flow keeps addresses and barrier
        // operands in uniform registers); one elected lane issues.
        int stage = 0;
        uint32_t phase = 0;
        TR_DECL(empty);
        TR_DECL(total);
#ifdef VIT_GEMM_TRACE
        tr_total = -clock64();
#endif
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int m0 = (tile / tiles_n) * 256 + rank * 128;
            const int n0 = (tile % tiles_n) * BN + rank * 128;
            for (int kb = 0; kb < num_kb; ++kb) {
                TR_WAIT(empty, mbar_wait(&empty_bar[stage], phase ^ 1));
                if (elect_one()) {
                    uint8_t* sa = smem + stage * L::STAGE_BYTES;
                    // (a TMA box counts its full size towards the barrier, zero-filled out-of-bounds parts included)
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (EMBED ? embed_rows * 128 + L::B_BYTES : L::STAGE_BYTES));
                    if constexpr (EMBED) {
                        // K block kb = channel kb / 8, kernel rows ky = 2 (kb % 8) + {0, 1}, all 16 kx: one sub-tile per kernel row
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            tma_load_5d_pair(sa + h * EMBED_SUBTILE_BYTES, &tmap_a, &full_bar[stage], 0, (kb & 7) * 2 + h, 0, embed_gy0(tile / tiles_n),
                                             embed_img(tile / tiles_n) * 3 + (kb >> 3));
                    } else tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * BK_ELEMS, m0);
                    tma_load_2d_pair(sa + L::A_BYTES, &tmap_b, &full_bar[stage], kb * BK_ELEMS, n0);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
#ifdef VIT_GEMM_TRACE
        tr_total += clock64();
        if (blockIdx.x == 0 && lane == 0)
            printf("gemm_trace producer K=%d N=%d epi=%d res16=%d tiles=%d total=%lld wait_empty=%lld\n", p.K, p.N, EPI, int(RES16),
                   (num_tiles - pair + num_pairs - 1) / num_pairs, tr_total, tr_empty);
#endif
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = EMBED ? make_idesc_tf32(256, BN) : make_idesc<T>(256, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            TR_DECL(tempty);
            TR_DECL(full);
            TR_DECL(total);
#ifdef VIT_GEMM_TRACE
            tr_total = -clock64();
#endif
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                TR_WAIT(tempty, mbar_wait(&tempty_bar[acc], acc_phase ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    TR_WAIT(full, mbar_wait(&full_bar[stage], phase));
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
                        const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {   // four K steps of 32 bytes per 128-byte operand row
                            if constexpr (EMBED)   // A: K steps 0, 1 in the first 64B-swizzled sub-tile, 2, 3 in the second; B: 128B-swizzled rows
                                umma_tf32_pair(d_tmem, desc_kmajor_sw64(a_addr + (k >> 1) * EMBED_SUBTILE_BYTES, k & 1), desc_kmajor_sw128(b_addr, k), idesc, (kb | k) != 0);
                            else umma_f16_pair(d_tmem, desc_kmajor_sw128(a_addr, k), desc_kmajor_sw128(b_addr, k), idesc, (kb | k) != 0);
                        }
                        umma_commit_pair(&empty_bar[stage], 0x3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit_pair(&tfull_bar[acc], 0x3);
                __syncwarp();
                if ((acc ^= 1) == 0) acc_phase ^= 1;
            }
#ifdef VIT_GEMM_TRACE
            tr_total += clock64();
            if (blockIdx.x == 0 && lane == 0)
                printf("gemm_trace mma      K=%d N=%d epi=%d res16=%d total=%lld wait_tempty=%lld wait_full=%lld\n", p.K, p.N, EPI, int(RES16), tr_total, tr_tempty, tr_full);
#endif
        }
    } else if (warp == W_LOADER && STAGED) {
        // ------------------------------------------------------------ parameter stager (EPI_BIAS / EPI_BIAS_GELU)
        // One tile ahead of the epilogue: the tile's 256 bias values and, for the LayerNorm consumer, the
        // column sums of the folded weights and the finished row statistics (the six partial sums of each
        // of the 128 rows -> rstd and -rstd * mean).  The epilogue warps would otherwise sit through a
        // global-memory round trip (or several) at the start of every tile.
        int j = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++j) {
            const int m0 = (tile / tiles_n) * 256 + rank * 128;
            const int n0 = (tile % tiles_n) * BN;
            const int b = j & 1;
            float* pb = s_bias + b * PRE_FLOATS;
            float4 bv[2];
            [[maybe_unused]] float4 cv[2];
            [[maybe_unused]] float2 part[4][6];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                bv[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + lane + 32 * i);
                if constexpr (LN) cv[i] = __ldg(reinterpret_cast<const float4*>(p.colsum + n0) + lane + 32 * i);
            }
            if constexpr (LN) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = m0 + lane + 32 * i;
#pragma unroll
                    for (int q = 0; q < 6; ++q)
                        part[i][q] = (r < p.M && q < p.stats_parts) ? __ldg(p.stats_in + static_cast<size_t>(q) * p.stats_rows + r)
                                                                    : make_float2(0.f, 0.f);
                }
            }
            mbar_wait(&pre_free[b], ((j >> 1) & 1) ^ 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                reinterpret_cast<float4*>(pb)[lane + 32 * i] = bv[i];
                if constexpr (LN) reinterpret_cast<float4*>(pb + 256)[lane + 32 * i] = cv[i];
            }
            if constexpr (LN) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int q = 0; q < 6; ++q) {   // fixed order: parts 0..5
                        s1 += part[i][q].x;
                        s2 += part[i][q].y;
                    }
                    const float mean = s1 * (1.0f / 768.0f);
                    const float var = fmaxf(s2 * (1.0f / 768.0f) - mean * mean, 0.f);  // ViT_seq.c:115 (single pass)
                    const float rstd = rsqrtf(var + 1e-6f);
                    reinterpret_cast<float2*>(pb + 512)[lane + 32 * i] = make_float2(rstd, -rstd * mean);
                }
            }
            mbar_arrive(&pre_full[b]);
        }
    } else if (warp == W_LOADER) {
        // ------------------------------------------------------------ residual loader (EPI_BIAS_RESIDUAL)
        if (kResidual && lane == 0) {
            uint32_t k = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m0 = (tile / tiles_n) * 256 + rank * 128;
                const int n0 = (tile % tiles_n) * BN;
                for (int c = 0; c < NCHUNK; ++c, ++k) {
                    const uint32_t slot = k % SLOTS, ph = (k / SLOTS) & 1;
                    mbar_wait(&slot_free[slot], ph ^ 1);
                    mbar_arrive_expect_tx(&res_full[slot], GEMM_SLOT_BYTES);
                    if constexpr (EMBED)
                        tma_load_2d(smem + L::SLOT_OFF + slot * GEMM_SLOT_BYTES, &tmap_res, &res_full[slot], n0 + c * CHUNK_COLS, 1 + embed_patch0(tile / tiles_n));
                    else
                        tma_load_2d(smem + L::SLOT_OFF + slot * GEMM_SLOT_BYTES, &tmap_out, &res_full[slot], n0 + c * CHUNK_COLS, m0);
                }
            }
        }
    } else if (warp == W_TMEM) {
        // ------------------------------------------------------------ chunk storer (RES16)
        // The epilogue warps of the 16-bit residual kernel never wait for each other: a warp that has updated its part of
        // a slot arrives on chunk_ready[slot] and goes on to the next chunk; this otherwise idle warp issues the chunk's
        // TMA store once all EPI_WARPS parts are in, and hands the slot of the PREVIOUS chunk back to the loader when that
        // chunk's store has drained.  (With the store issued by epilogue thread 0 behind a CTA-wide barrier per chunk,
        // out_proj -- K = 768, one 3.2 us MMA per tile -- was bound by its epilogue: profiles/r2_gemm_trace.txt.)
        if (RES16 && lane == 0) {
            uint32_t k = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m0 = (tile / tiles_n) * 256 + rank * 128;
                const int n0 = (tile % tiles_n) * BN;
                for (int c = 0; c < NCHUNK; ++c, ++k) {
                    const uint32_t slot = k % SLOTS;
                    mbar_wait(&chunk_ready[slot], (k / SLOTS) & 1);
                    tma_store_2d(&tmap_out, smem + L::SLOT_OFF + slot * GEMM_SLOT_BYTES, n0 + c * CHUNK_COLS, m0);
                    tma_store_commit();
                    if (k >= 1) {
                        tma_store_wait_read<1>();
                        mbar_arrive(&slot_free[(k - 1) % SLOTS]);
                    }
                }
            }
            tma_store_wait_all<0>();
        }
    } else if (warp < EPI_WARPS) {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int ew = warp;
        const int quarter = warp & 3;
        const int half = ew >> 2;                 // which part of the chunk's columns
        const int etid = threadIdx.x;
        const bool storer = etid == 0;
        const int row = quarter * 32 + lane;      // row inside the CTA's 128-row block
        const uint32_t row_off = row * 128, sw = row & 7;
        const uint32_t tempty_leader = mapa_shared(smem_u32(&tempty_bar[0]), 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t k = 0;  // running chunk counter (ring position)
        [[maybe_unused]] int tile_par = 0;
        [[maybe_unused]] uint32_t pre_phase = 0;
        TR_DECL(tfull);
        TR_DECL(drain);
        TR_DECL(res);
        TR_DECL(bar);
        TR_DECL(store);
        TR_DECL(pre);
        TR_DECL(comp);
        TR_DECL(total);
#ifdef VIT_GEMM_TRACE
        tr_total = -clock64();
#endif
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int m0 = (tile / tiles_n) * 256 + rank * 128;
            const int n0 = (tile % tiles_n) * BN;
            float* sb = s_bias;
            if constexpr (RES16) sb = s_bias + tile_par * 256;   // two bias buffers: the warps drift apart by up to a tile
            float ln_rstd = 1.f, ln_nm = 0.f;  // LN consumer: this thread's row: rstd and -rstd * mean
            if constexpr (!STAGED) {
                // Loaded by the epilogue threads themselves.  One buffer is enough: in the fp32 residual kernel every
                // chunk iteration below ends with a barrier after its last read of it, so nobody still reads the
                // previous tile's values here (the other kernels regroup explicitly; RES16 alternates two buffers,
                // whose previous use lies before the previous tile's barrier).  The global-memory round trip is short because each thread asked for the NEXT tile's lines
                // (prefetch.global.L1) one tile ago; TMA traffic bypasses L1, so they are still there.
                [[maybe_unused]] float2 part[6];
                if constexpr (LN && !kResidual) {
                    const bool row_ok = m0 + row < p.M;
#pragma unroll
                    for (int q = 0; q < 6; ++q)
                        part[q] = (row_ok && q < p.stats_parts) ? __ldg(p.stats_in + static_cast<size_t>(q) * p.stats_rows + m0 + row)
                                                                : make_float2(0.f, 0.f);
                }
                if constexpr (!kResidual) epi_bar_sync<EPI_THREADS>();  // the quarters run free in the chunk loop: regroup before the buffer is rewritten
                if (etid < 256) {
                    const float bv = __ldg(p.bias + n0 + etid);
                    [[maybe_unused]] float cv = 0.f;
                    if constexpr (LN && !kResidual) cv = __ldg(p.colsum + n0 + etid);
                    sb[etid] = bv;
                    if constexpr (LN && !kResidual) sb[256 + etid] = cv;
                }
                if constexpr (LN && !kResidual) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int q = 0; q < 6; ++q) {   // fixed order: parts 0..5
                        s1 += part[q].x;
                        s2 += part[q].y;
                    }
                    const float mean = s1 * (1.0f / 768.0f);
                    const float var = fmaxf(s2 * (1.0f / 768.0f) - mean * mean, 0.f);  // ViT_seq.c:115 (single pass)
                    ln_rstd = rsqrtf(var + 1e-6f);
                    ln_nm = -ln_rstd * mean;
                }
                const int next = tile + num_pairs;
                if (next < num_tiles) {
                    const int m1 = (next / tiles_n) * 256 + rank * 128, n1 = (next % tiles_n) * BN;
                    if (etid < 256) {
                        prefetch_l1(p.bias + n1 + etid);
                        if constexpr (LN && !kResidual) prefetch_l1(p.colsum + n1 + etid);
                    }
                    if constexpr (LN && !kResidual) {
                        if (m1 + row < p.M) {
#pragma unroll
                            for (int q = 0; q < 6; ++q)
                                if (q < p.stats_parts) prefetch_l1(p.stats_in + static_cast<size_t>(q) * p.stats_rows + m1 + row);
                        }
                    }
                }
                TR_WAIT(bar, epi_bar_sync<EPI_THREADS>());
            } else {
                // staged one tile ahead by the loader warp (bias | column sums | row statistics)
                sb = s_bias + tile_par * PRE_FLOATS;
                TR_WAIT(pre, mbar_wait(&pre_full[tile_par], pre_phase));
                if constexpr (LN) {
                    const float2 st = reinterpret_cast<const float2*>(sb + 512)[row];
                    ln_rstd = st.x;
                    ln_nm = st.y;
                }
            }
            float st_sum = 0.f, st_sq = 0.f;   // LN producer: partial row statistics over this thread's 128 columns
            [[maybe_unused]] float* cls32 = nullptr;   // RES16: this thread's row is a class-token row -> its fp32 master row
            if constexpr (RES16) {
                const int grow = m0 + row;
                if (p.cls_rows32 && grow < p.M && grow % p.tokens == 0) cls32 = p.cls_rows32 + static_cast<size_t>(grow / p.tokens) * p.N;
            }

            TR_WAIT(tfull, mbar_wait(&tfull_bar[acc], acc_phase));
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * COLS_PER_WARP;
            // Early accumulator release: the thread's whole share of the accumulator stage (NCHUNK x
            // COLS_PER_WARP = 128 fp32) is pulled into registers first and the TMEM stage is handed back
            // to the MMA issuer at once, so the next-but-one tile's MMAs never wait for this tile's
            // GELU arithmetic or stores -- with only two accumulator stages, releasing the stage at the
            // end of the epilogue couples the MMA and epilogue periods (measured: MMA issuer spinning
            // on the stage barrier while the epilogue warps wait for the next accumulator).
            uint32_t racc[NCHUNK][COLS_PER_WARP];
#ifdef VIT_GEMM_TRACE
            tr_drain -= clock64();
#endif
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                tmem_ld_x16p(taddr + c * CHUNK_COLS, racc[c]);
                if constexpr (COLS_PER_WARP == 32) tmem_ld_x16p(taddr + c * CHUNK_COLS + 16, racc[c] + 16);
            }
            tmem_ld_wait();
#ifdef VIT_GEMM_TRACE
            tr_drain += clock64();
#endif
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + acc * 8);
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c, ++k) {
                const uint32_t slot = k % SLOTS;
                uint8_t* srow = smem + L::SLOT_OFF + slot * GEMM_SLOT_BYTES + row_off;
                const float* bcol = sb + c * CHUNK_COLS + half * COLS_PER_WARP;
                const uint32_t* r = racc[c];
                if constexpr (RES16) {
                    TR_WAIT(res, mbar_wait(&res_full[slot], (k / SLOTS) & 1));
#ifdef VIT_GEMM_TRACE
                    tr_comp -= clock64();
#endif
#pragma unroll
                    for (int j = 0; j < PIECES; ++j) {   // this thread's 32 columns of the 64-column chunk: four 16-byte pieces
                        uint4* q = reinterpret_cast<uint4*>(srow + (((half * PIECES + j) ^ sw) << 4));
                        const uint4 u = *q;
                        uint32_t w[4] = {u.x, u.y, u.z, u.w};
                        // A class-token row keeps an fp32 master copy: it is the one row the head reads, its residual adds
                        // go straight into the logits (the other rows reach it only through the attention's averages).
                        float2* m32 = cls32 ? reinterpret_cast<float2*>(cls32 + n0 + c * CHUNK_COLS + half * COLS_PER_WARP + 8 * j) : nullptr;
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            float2 xr = unpack2<T>(w[h]);
                            if (m32) xr = m32[h];
                            xr.x += __uint_as_float(r[8 * j + 2 * h]) + bcol[8 * j + 2 * h];
                            xr.y += __uint_as_float(r[8 * j + 2 * h + 1]) + bcol[8 * j + 2 * h + 1];
                            if (m32) m32[h] = xr;
                            w[h] = pack2<T>(xr.x, xr.y);
                            const float2 v = unpack2<T>(w[h]);   // statistics of the row as stored
                            st_sum += v.x + v.y;
                            st_sq = fmaf(v.x, v.x, fmaf(v.y, v.y, st_sq));
                        }
                        *q = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                } else if constexpr (kResidual) {
                    TR_WAIT(res, mbar_wait(&res_full[slot], (k / SLOTS) & 1));
                    [[maybe_unused]] uint32_t cast[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4* q = reinterpret_cast<float4*>(srow + (((half * 4 + j) ^ sw) << 4));
                        float4 v = *q;
                        v.x += __uint_as_float(r[4 * j + 0]) + bcol[4 * j + 0];
                        v.y += __uint_as_float(r[4 * j + 1]) + bcol[4 * j + 1];
                        v.z += __uint_as_float(r[4 * j + 2]) + bcol[4 * j + 2];
                        v.w += __uint_as_float(r[4 * j + 3]) + bcol[4 * j + 3];
                        *q = v;
                        if constexpr (LN) {
                            st_sum += (v.x + v.y) + (v.z + v.w);
                            st_sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, st_sq))));
                            cast[2 * j] = pack2<T>(v.x, v.y);
                            cast[2 * j + 1] = pack2<T>(v.z, v.w);
                        }
                    }
                    if constexpr (LN) {
                        // operand-precision copy: two fp32 chunks (2 x 32 columns) fill one 64-column staging
                        // tile; this thread's 16 columns are two 16-byte pieces of the 128-byte row
                        uint8_t* crow = smem + L::CAST_OFF + ((c >> 1) & (CAST_BUFS - 1)) * GEMM_SLOT_BYTES + row_off;
                        const uint32_t piece = (c & 1) * 4 + half * 2;
                        if constexpr (CAST_BUFS == 1) {
                            // single staging tile: the store of the previous 64-column block (issued by the storer
                            // after the last barrier) must have read it before anyone writes the next block
                            if ((c & 1) == 0 && k >= 2) mbar_wait(&pre_full[0], ((k >> 1) - 1) & 1);
                        }
                        *reinterpret_cast<uint4*>(crow + (((piece + 0) ^ sw) << 4)) = make_uint4(cast[0], cast[1], cast[2], cast[3]);
                        *reinterpret_cast<uint4*>(crow + (((piece + 1) ^ sw) << 4)) = make_uint4(cast[4], cast[5], cast[6], cast[7]);
                    }
                } else {
                    uint32_t packed[COLS_PER_WARP / 2];
                    const bool as_bf16 = EPI == EPI_BIAS && p.bf16_from_col > 0 && n0 + c * CHUNK_COLS >= p.bf16_from_col;
                    [[maybe_unused]] const float* scol = bcol + 256;
#pragma unroll
                    for (int j = 0; j < COLS_PER_WARP / 2; ++j) {
                        // two adjacent columns per step, fp32 arithmetic issued in pairs (FFMA2 / FMUL2)
                        const float2 a = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                        const float2 bc = *reinterpret_cast<const float2*>(bcol + 2 * j);
                        float2 v;
                        if constexpr (LN) v = fma2(splat2(ln_nm), *reinterpret_cast<const float2*>(scol + 2 * j), fma2(a, splat2(ln_rstd), bc));
                        else v = add2(a, bc);
                        if constexpr (EPI == EPI_BIAS_GELU) v = gelu_erf2(v);
                        packed[j] = as_bf16 ? pack2<__nv_bfloat16>(v.x, v.y) : pack2<T>(v.x, v.y);
                    }
                    // the slot's previous store (chunk k - SLOTS) has drained: the storer checked before
                    // the barrier that ended chunk k - 1
#pragma unroll
                    for (int j = 0; j < PIECES; ++j)
                        *reinterpret_cast<uint4*>(srow + (((half * PIECES + j) ^ sw) << 4)) =
                            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
                fence_proxy_async_smem();
                if constexpr (RES16) {   // hand this warp's part of the chunk to the storer warp and move on
#ifdef VIT_GEMM_TRACE
                    tr_comp += clock64();
#endif
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&chunk_ready[slot]);
                    continue;
                }
                if constexpr (!kResidual) {
                    // Each TMEM lane quarter (two warps, 32 rows) stores its own 32 x 64 piece of the chunk:
                    // the quarters only ever meet their own partner warp, so they drift apart and the MUFU
                    // pipe (two ops per GELU: the limiter of the mlp_0 epilogue) is not left idle while all
                    // eight warps gather at a chunk barrier.
                    const bool qstorer = half == 0 && lane == 0;
                    if (qstorer) TR_WAIT(store, tma_store_wait_read<SLOTS - 2>());  // frees this quarter's part of the slot of chunk k + 1
                    TR_WAIT(bar, asm volatile("bar.sync %0, %1;" ::"r"(2 + quarter), "n"(PARTS * 32) : "memory"));
                    if (qstorer) {
                        tma_store_2d(&tmap_cast /* 32-row boxes of the output */, smem + L::SLOT_OFF + slot * GEMM_SLOT_BYTES + quarter * 4096,
                                     n0 + c * CHUNK_COLS, m0 + quarter * 32);
                        tma_store_commit();
                    }
                    continue;
                }
                TR_WAIT(bar, epi_bar_sync<EPI_THREADS>());
                if (storer) {
                    if constexpr (EMBED)
                        tma_store_3d(&tmap_out, smem + L::SLOT_OFF + slot * GEMM_SLOT_BYTES, n0 + c * CHUNK_COLS, 1 + embed_patch0(tile / tiles_n), embed_img(tile / tiles_n));
                    else
                        tma_store_2d(&tmap_out, smem + L::SLOT_OFF + slot * GEMM_SLOT_BYTES, n0 + c * CHUNK_COLS, m0);
                    if constexpr (kResidual && LN && !RES16) {
                        // Same bulk group as the fp32 chunk, so the wait below also covers the staging tile:
                        // it is rewritten two 64-column blocks later, i.e. after the barriers of chunks c + 1
                        // and c + 2, which this thread only joins after wait_read<1> has seen this group through.
                        // (With a single staging tile the storer simply waits for this group before moving on.)
                        if (c & 1) {
                            const uint8_t* csrc = smem + L::CAST_OFF + ((c >> 1) & (CAST_BUFS - 1)) * GEMM_SLOT_BYTES;
                            if constexpr (EMBED) tma_store_3d(&tmap_cast, csrc, n0 + (c >> 1) * 64, 1 + embed_patch0(tile / tiles_n), embed_img(tile / tiles_n));
                            else tma_store_2d(&tmap_cast, csrc, n0 + (c >> 1) * 64, m0);
                        }
                    }
                    tma_store_commit();
                    if constexpr (CAST_BUFS == 1) {
                        if (c & 1) {
                            tma_store_wait_read<0>();
                            mbar_arrive(&pre_full[0]);
                        }
                    }
                    if constexpr (kResidual) {
                        if (k >= 1) {  // the previous chunk's store has finished reading its slot: hand it to the loader
                            TR_WAIT(store, tma_store_wait_read<1>());
                            mbar_arrive(&slot_free[(k - 1) % SLOTS]);
                        }
                    }
                }
            }
            if constexpr (kResidual && LN) {
                if constexpr (EMBED) {
                    const int patch = embed_patch0(tile / tiles_n) + row;
                    if (row < embed_rows && patch < p.patches)
                        p.stats_out[static_cast<size_t>(2 * (tile % tiles_n) + half) * p.stats_rows +
                                    static_cast<size_t>(embed_img(tile / tiles_n)) * p.tokens + 1 + patch] = make_float2(st_sum, st_sq);
                } else if (m0 + row < p.M)
                    p.stats_out[static_cast<size_t>(2 * (tile % tiles_n) + half) * p.stats_rows + m0 + row] = make_float2(st_sum, st_sq);
            }
            if constexpr (STAGED) {   // all reads of the staged parameters precede the last chunk's barrier
                __syncwarp();
                if (lane == 0) mbar_arrive(&pre_free[tile_par]);
                if ((tile_par ^= 1) == 0) pre_phase ^= 1;
            }
            if constexpr (RES16) tile_par ^= 1;
            if ((acc ^= 1) == 0) acc_phase ^= 1;
        }
#ifdef VIT_GEMM_TRACE
        tr_total += clock64();
        if (blockIdx.x == 0 && etid == 0)
            printf("gemm_trace epilogue K=%d N=%d epi=%d res16=%d total=%lld wait_tfull=%lld drain=%lld wait_residual=%lld barriers=%lld wait_store_read=%lld wait_params=%lld chunk_math=%lld\n",
                   p.K, p.N, EPI, int(RES16), tr_total, tr_tfull, tr_drain, tr_res, tr_bar, tr_store, tr_pre, tr_comp);
#endif
        if (kResidual32 ? storer : (!kResidual && half == 0 && lane == 0)) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == W_TMEM) {
        tc_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

}  // namespace vit
