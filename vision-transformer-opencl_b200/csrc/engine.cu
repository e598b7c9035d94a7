// engine.cu -- the C ABI of include/vit_cuda.h: weight arena (conversion, LayerNorm folding, cache file),
// workspaces, TMA descriptors, the ViT-B/16 forward schedule (ViT_seq.c:337-439 / ViT_opencl.c:785-883
// re-expressed as a batch of token-flattened kernels), the host pipeline (one feeding thread per GPU) and the
// single-operator test entry points (op_entry.inc).
//
// There is deliberately no CPU fallback: every entry point fails with VIT_E_NODEVICE unless a
// compute-capability-10.x device is present.
#include "../../include/vit_cuda.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

#include "attention_sm100.cuh"
#include "gemm_sm100.cuh"
#include "kernels_misc.cuh"

namespace {

using namespace vit;

constexpr int kHeads = 12, kHidden = 3072, kDepth = 12, kClasses = VIT_NUM_CLASSES, kPatch = 16;

thread_local char t_err[512] = "";
std::atomic<long long> g_launches{0};

int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                               \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return set_err(e__ == cudaErrorMemoryAllocation ? VIT_E_NOMEM : VIT_E_CUDA,            \
                           "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));  \
    } while (0)
#define VIT_TRY(expr)          \
    do {                       \
        int r__ = (expr);      \
        if (r__ != 0) return r__; \
    } while (0)

// ------------------------------------------------------------------------------------ driver API
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
std::atomic<PFN_encodeTiled> g_encode{nullptr};

int load_driver() {
    if (g_encode.load()) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return set_err(VIT_E_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode.store(reinterpret_cast<PFN_encodeTiled>(fn));
    return 0;
}

CUtensorMapDataType operand_dtype(int prec) {
    return prec == VIT_PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

// Dense tensor of `rank` dimensions (dims / box innermost first), 128B swizzle: the innermost box extent times the
// element size is <= 128 bytes.  strides_bytes[i] = stride of dimension i + 1.
int encode_tmap(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    VIT_TRY(load_driver());
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = g_encode.load()(m, dt, rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_err(VIT_E_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d, dims %llu x %llu ..., box %u x %u ...", (int)r, rank,
                       (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return 0;
}

// 2-D row-major [rows][cols] tensor, box {box_cols, box_rows} (box_cols * elem = 128 B).
int make_tmap_raw(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t cols, uint64_t rows,
                  uint32_t box_cols, uint32_t box_rows) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * elem_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    return encode_tmap(m, dt, 2, base, dims, strides, box);
}
// operand-precision (16-bit) tensor
int make_tmap(CUtensorMap* m, int prec, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows) {
    return make_tmap_raw(m, operand_dtype(prec), 2, base, cols, rows, box_cols, box_rows);
}
// fp32 tensor (residual stream, tf32 conv_proj weight): 32 columns = 128 bytes per box row
int make_tmap_f32(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    return make_tmap_raw(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, cols, rows, 32, box_rows);
}
// 3-D [d2][d1][d0] tensor of 16-bit elements, dense, box {64, box_d1, 1}: per-image views whose rows
// past d1 are clipped on store / zero-filled on load.
int make_tmap_3d(CUtensorMap* m, int prec, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box_d1) {
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    const cuuint32_t box[3] = {64, box_d1, 1};
    return encode_tmap(m, operand_dtype(prec), 3, base, dims, strides, box);
}
// 3-D [d2][d1][d0] fp32 tensor, box {32, box_d1, 1} (128 bytes per box row)
int make_tmap_3d_f32(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box_d1) {
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
    const cuuint32_t box[3] = {32, box_d1, 1};
    return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box);
}
// The fp32 images [nb][3][S][S] as the A operand of conv_proj (no im2col buffer): dimensions, innermost first,
// kx (16, contiguous), ky (16, stride S), gx (G, stride 16), gy (G, stride 16 S), plane = image * 3 + channel
// (stride S * S).  A box {16, 1, G, 128 / G, 1} is one kernel row (16 pixels = 64 bytes, 64B swizzle) of (128 / G) * G
// patches; the kernel loads two of them per K block (gemm_sm100_staged_kernel, EMBED).
int make_tmap_image5d(CUtensorMap* m, const float* images, int S, int nb) {
    const uint64_t G = S / kPatch, gyc = 128 / G;
    const cuuint64_t dims[5] = {16, 16, G, G, static_cast<cuuint64_t>(3) * nb};
    const cuuint64_t strides[4] = {static_cast<cuuint64_t>(S) * 4, 16 * 4, static_cast<cuuint64_t>(16) * S * 4, static_cast<cuuint64_t>(S) * S * 4};
    const cuuint32_t box[5] = {16, 1, static_cast<cuuint32_t>(G), static_cast<cuuint32_t>(gyc), 1};
    return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, images, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
}
// rows of a conv_proj CTA tile = box rows of its 3-D store maps (gemm_sm100_staged_kernel, EMBED)
int embed_rows_per_cta(int S) { return (128 / (S / kPatch)) * (S / kPatch); }
int embed_tiles_per_image(int S) {
    const int G = S / kPatch, gyc = 128 / G;
    return (G + 2 * gyc - 1) / (2 * gyc);
}

// ------------------------------------------------------------------------------------ run-time switches
struct Options {
    std::atomic<int> attn_exact{0}, prune_last{1}, ln_fused{1}, pdl{1}, graphs{1}, host_threads{1}, residual16{1}, wave_passes{1};
};
Options g_opt;
int snap_schedule(std::vector<int>& first, std::vector<int>& count, int n_pass, int n_images, int max_batch, int tokens, int pairs);

int env_flag(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? (atoi(s) != 0) : dflt;
}

// ------------------------------------------------------------------------------------ launches
constexpr int kGemmBN = 256;

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(VIT_E_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// Launch with programmatic stream serialization (the kernel must call griddep_wait() before it touches anything
// its predecessor wrote).  VIT_OPT_PDL = 0: ordinary launches.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_opt.pdl.load(std::memory_order_relaxed) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): set once per pair, from whichever host
// thread gets there first (one feeding thread per GPU, so the mask is atomic).
template <typename K>
int configure_smem(K kern, std::atomic<unsigned>& done_mask, int bytes) {
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (!(done_mask.load(std::memory_order_acquire) & (1u << dev))) {
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done_mask.fetch_or(1u << dev, std::memory_order_release);
    }
    return 0;
}

constexpr int kStagedStages = 5, kStagedSlots = 4;
template <typename T, int EPI, bool LN, int STAGES, int SLOTS, int CAST, bool PSTAGED = (EPI != EPI_BIAS_RESIDUAL), bool EMBED = false, int kEpiWarps = 8, bool RES16 = false>
int launch_gemm_staged_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tcast,
                           const GemmParams& p, int sm_count, cudaStream_t st, const CUtensorMap* tres = nullptr) {
    constexpr int kPreFloats = !PSTAGED ? 0 : (LN ? 256 + 256 + 2 * 128 : 256);
    using L = GemmStagedSmem<STAGES, SLOTS, CAST, kPreFloats * 4>;
    static_assert(L::DYN_BYTES <= 232448, "shared memory budget");
    auto kern = gemm_sm100_staged_kernel<T, STAGES, SLOTS, EPI, kEpiWarps, LN, CAST, PSTAGED, EMBED, RES16>;
    static std::atomic<unsigned> configured{0};
    VIT_TRY(configure_smem(kern, configured, L::DYN_BYTES));
    if (p.N % kGemmBN || p.K % GEMM_BK || p.M <= 0)
        return set_err(VIT_E_ARG, "gemm shape M=%d N=%d K=%d unsupported (N%%256, K%%64)", p.M, p.N, p.K);
    if (LN && (EPI == EPI_BIAS_RESIDUAL ? (p.N != kDim || !p.stats_out) : (p.K != kDim || !p.stats_in || !p.colsum || p.stats_parts <= 0)))
        return set_err(VIT_E_ARG, "LayerNorm-folded gemm: bad statistics arguments (N=%d K=%d)", p.N, p.K);
    if (EMBED && (p.grid_w <= 0 || p.grid_w > 64 || p.patches != p.grid_w * p.grid_w))
        return set_err(VIT_E_ARG, "conv_proj: patch grid %d x %d not supported (1..64)", p.grid_w, p.grid_w);
    const int tiles = ((p.M + 255) / 256) * (p.N / 256);
    const int grid = 2 * std::min(tiles, sm_count / 2);
    CU_TRY(launch_pdl(kern, dim3(grid), dim3((GEMM_NON_EPI_WARPS + kEpiWarps) * 32), L::DYN_BYTES, st, ta, tb, tout, tcast, tres ? *tres : tout, p));
    return check_launch("gemm_staged");
}
// LN = true: LayerNorm folded into the GEMM (consumer for EPI_BIAS / EPI_BIAS_GELU, producer for
// EPI_BIAS_RESIDUAL, see gemm_sm100.cuh).  The producer pays for the staging tiles of the
// operand-precision copy with an operand stage or a residual slot.
template <typename T, int EPI, bool LN>
int launch_gemm_staged_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tcast,
                         const GemmParams& p, int sm_count, cudaStream_t st) {
    if constexpr (LN && EPI == EPI_BIAS_RESIDUAL) {
        // Measured (B = 1024): mlp_3 (K = 3072, tensor bound) needs the fifth operand stage (4 stages: +9 %) and
        // is indifferent to the slot count; out_proj (K = 768, HBM bound) needs the four residual slots to keep
        // enough chunk loads in flight (3 slots: +23 %) and is indifferent to the stage count.
        if (p.K >= 2048) return launch_gemm_staged_cfg<T, EPI, LN, 5, 3, 1>(ta, tb, tout, tcast, p, sm_count, st);
        return launch_gemm_staged_cfg<T, EPI, LN, 4, 4, 2>(ta, tb, tout, tcast, p, sm_count, st);
    } else if constexpr (EPI == EPI_BIAS_RESIDUAL) {
        return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, kStagedSlots, 0>(ta, tb, tout, tcast, p, sm_count, st);
    } else if constexpr (LN) {
        // 6 KB of staged parameters cost one output slot (loading them in the epilogue threads instead, with 4 slots,
        // measured 7 % slower on mlp_0; 16 instead of 8 epilogue warps for the GELU epilogue: no change)
        return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, 3, 0, true>(ta, tb, tout, tcast, p, sm_count, st);
    } else {
        return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, kStagedSlots, 0, true>(ta, tb, tout, tcast, p, sm_count, st);
    }
}
// conv_proj as the EMBED variant of the residual kernel (tf32 MMA straight from the fp32 image): ta 5-D image map
// (make_tmap_image5d), tb fp32 weight map, tout 3-D fp32 token map, tcast 3-D operand-precision token map (ln only) --
// both with embed_rows_per_cta() rows per box --, tpos 2-D pos_embedding map.  p.M = images * tiles per image * 256.
template <typename T>
int launch_gemm_embed_t(bool ln, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tcast,
                        const CUtensorMap& tpos, const GemmParams& p, int sm_count, cudaStream_t st) {
    if (ln) return launch_gemm_staged_cfg<T, EPI_BIAS_RESIDUAL, true, 4, 4, 2, false, true>(ta, tb, tout, tcast, p, sm_count, st, &tpos);
    return launch_gemm_staged_cfg<T, EPI_BIAS_RESIDUAL, false, kStagedStages, kStagedSlots, 0, false, true>(ta, tb, tout, tcast, p, sm_count, st, &tpos);
}
int launch_gemm_embed(int prec, bool ln, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tcast,
                      const CUtensorMap& tpos, const GemmParams& p, int sm_count, cudaStream_t st) {
    return prec == VIT_PREC_FP16 ? launch_gemm_embed_t<__half>(ln, ta, tb, tout, tcast, tpos, p, sm_count, st)
                                 : launch_gemm_embed_t<__nv_bfloat16>(ln, ta, tb, tout, tcast, tpos, p, sm_count, st);
}

// tout: store map of the output, 128-row boxes (for the residual epilogue also the load map of the residual, in
// place).  taux: EPI_BIAS / EPI_BIAS_GELU: store map of the same output with 32-row boxes (one per TMEM lane quarter);
// EPI_BIAS_RESIDUAL: unused.
template <int EPI>
int launch_gemm_staged(int prec, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& taux,
                       const GemmParams& p, int sm_count, cudaStream_t st) {
    return prec == VIT_PREC_FP16 ? launch_gemm_staged_t<__half, EPI, false>(ta, tb, tout, taux, p, sm_count, st)
                                 : launch_gemm_staged_t<__nv_bfloat16, EPI, false>(ta, tb, tout, taux, p, sm_count, st);
}
// LayerNorm-folded variants; taux as above, for the producer (EPI_BIAS_RESIDUAL): store map of the operand-precision copy
template <int EPI>
int launch_gemm_staged_ln(int prec, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& taux,
                          const GemmParams& p, int sm_count, cudaStream_t st) {
    return prec == VIT_PREC_FP16 ? launch_gemm_staged_t<__half, EPI, true>(ta, tb, tout, taux, p, sm_count, st)
                                 : launch_gemm_staged_t<__nv_bfloat16, EPI, true>(ta, tb, tout, taux, p, sm_count, st);
}

// The residual GEMM on a 16-bit residual stream (RES16): trow = load / store map of the rows in the operand type (64-column,
// 128-row boxes: the same map the next GEMM loads its A operand through), updated in place; p.stats_out as for the fp32 form.
int launch_gemm_residual16(int prec, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& trow, const GemmParams& p, int sm_count,
                           cudaStream_t st) {
    if (prec == VIT_PREC_FP16)
        return launch_gemm_staged_cfg<__half, EPI_BIAS_RESIDUAL, true, 5, 4, 0, false, false, 8, true>(ta, tb, trow, trow, p, sm_count, st);
    return launch_gemm_staged_cfg<__nv_bfloat16, EPI_BIAS_RESIDUAL, true, 5, 4, 0, false, false, 8, true>(ta, tb, trow, trow, p, sm_count, st);
}

template <typename T, bool EXACT>
int launch_attention_stream_t(const CUtensorMap& tqkv, const CUtensorMap& tout32, const AttnParams& p, int sm_count, cudaStream_t st) {
    auto kern = attention_sm100_stream_kernel<T, EXACT>;
    const int smem = attn3_smem_bytes(224);   // the largest single-block case: one attribute value serves every token count
    static std::atomic<unsigned> configured{0};
    VIT_TRY(configure_smem(kern, configured, smem));
    CU_TRY(launch_pdl(kern, dim3(std::min(p.batch * kHeads, sm_count)), dim3(ATTN3_THREADS), attn3_smem_bytes(p.kpad), st, tqkv, tout32, p));
    return check_launch("attention_stream");
}
template <typename T>
int launch_attention_stream_blocked_t(const CUtensorMap& tqkv, const CUtensorMap& tout32, const AttnParams& p, int sm_count, cudaStream_t st) {
    auto kern = attention_sm100_stream_blocked_kernel<T>;
    static std::atomic<unsigned> configured{0};
    VIT_TRY(configure_smem(kern, configured, attn4_smem_bytes(ATTNL_MAX_TOKENS)));
    CU_TRY(launch_pdl(kern, dim3(std::min(p.batch * kHeads, sm_count)), dim3(ATTN4_THREADS), attn4_smem_bytes(p.tokens), st, tqkv, tout32, p));
    return check_launch("attention_stream_blocked");
}
template <typename T>
int launch_attention_blocked_t(const CUtensorMap& tqkv, const CUtensorMap& tout, const AttnParams& p, int sm_count, cudaStream_t st) {
    auto kern = attention_sm100_blocked_kernel<T>;
    static std::atomic<unsigned> configured{0};
    VIT_TRY(configure_smem(kern, configured, attnl_smem_bytes(ATTNL_MAX_TOKENS)));
    kern<<<std::min(p.batch * kHeads, sm_count), ATTNL_THREADS, attnl_smem_bytes(p.tokens), st>>>(tqkv, tout, p);
    return check_launch("attention_blocked");
}
constexpr int kAttnSingleBlockMaxTokens = 224;  // up to here the whole key range of an image is one S block in TMEM
// rows per box of the Q/K/V load map (make_tmap over the packed QKV activation) that launch_attention expects
int attention_load_box_rows(int tokens) {
    return tokens <= kAttnSingleBlockMaxTokens ? ((tokens + 15) / 16 * 16) / 2 : ATTNL_KB;
}
// tqkv: load map of the packed QKV activation with attention_load_box_rows(tokens) rows per box.
// exact: two-pass softmax (exact row maximum); otherwise the single-pass variant, which raises
// VIT_FLAG_ATTN_RANGE when a row left its exponent window (the caller then repeats with exact = true).
// tout: 3-D store map of the output with 128-row boxes (two-pass key-blocked kernel), tout32: the same
// with 32-row boxes (streaming kernels, one store per output warp).
int launch_attention(int prec, const CUtensorMap& tqkv, const CUtensorMap& tout, const CUtensorMap& tout32, const AttnParams& p,
                     int sm_count, cudaStream_t st, bool exact) {
    if (p.tokens > ATTNL_MAX_TOKENS) return set_err(VIT_E_ARG, "attention: tokens=%d > %d not supported", p.tokens, ATTNL_MAX_TOKENS);
    const bool h = prec == VIT_PREC_FP16;
    if (p.tokens > kAttnSingleBlockMaxTokens) {
        // key-blocked: single-pass streaming kernel unless the exact softmax is asked for (or after a range flag)
        if (!exact)
            return h ? launch_attention_stream_blocked_t<__half>(tqkv, tout32, p, sm_count, st)
                     : launch_attention_stream_blocked_t<__nv_bfloat16>(tqkv, tout32, p, sm_count, st);
        return h ? launch_attention_blocked_t<__half>(tqkv, tout, p, sm_count, st) : launch_attention_blocked_t<__nv_bfloat16>(tqkv, tout, p, sm_count, st);
    }
    if (exact) return h ? launch_attention_stream_t<__half, true>(tqkv, tout32, p, sm_count, st)
                        : launch_attention_stream_t<__nv_bfloat16, true>(tqkv, tout32, p, sm_count, st);
    return h ? launch_attention_stream_t<__half, false>(tqkv, tout32, p, sm_count, st)
             : launch_attention_stream_t<__nv_bfloat16, false>(tqkv, tout32, p, sm_count, st);
}

// Reads and clears the current device's status word (synchronous; operator tests and init).
int take_status_flags(unsigned int* flags) {
    unsigned int v = 0;
    CU_TRY(cudaMemcpyFromSymbol(&v, g_status_flags, sizeof(v)));
    *flags = v;
    if (v) {
        v = 0;
        CU_TRY(cudaMemcpyToSymbol(g_status_flags, &v, sizeof(v)));
    }
    return 0;
}

int launch_layernorm(int prec, const float* x, const float* w, const float* b, void* y, int rows, cudaStream_t st) {
    const int grid = (rows + 7) / 8;
    if (prec == VIT_PREC_FP16) layernorm_kernel<__half><<<grid, 256, 0, st>>>(x, w, b, static_cast<__half*>(y), rows);
    else layernorm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, w, b, static_cast<__nv_bfloat16*>(y), rows);
    return check_launch("layernorm");
}

int launch_rowstats_cast(int prec, const float* x, void* y, float2* stats, int rows, cudaStream_t st) {
    const int grid = (rows + 7) / 8;
    if (prec == VIT_PREC_FP16) rowstats_cast_kernel<__half><<<grid, 256, 0, st>>>(x, static_cast<__half*>(y), stats, rows);
    else rowstats_cast_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(y), stats, rows);
    return check_launch("rowstats_cast");
}

int launch_fold_ln(int prec, const float* W, const float* ln_w, const float* ln_b, const float* bias, void* Wp, float* colsum,
                   float* cvec, int N, cudaStream_t st) {
    const int grid = (N + 7) / 8;
    if (prec == VIT_PREC_FP16)
        fold_ln_weights_kernel<__half><<<grid, 256, 0, st>>>(W, ln_w, ln_b, bias, static_cast<__half*>(Wp), colsum, cvec, N);
    else
        fold_ln_weights_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(W, ln_w, ln_b, bias, static_cast<__nv_bfloat16*>(Wp), colsum, cvec, N);
    return check_launch("fold_ln_weights");
}

int launch_cls_rows(int prec, bool ln, float* x, void* xn, float2* pstats, int stats_rows, const float* cls, const float* pos, int nb,
                    int tokens, cudaStream_t st, float* cls_rows32 = nullptr) {
    if (!ln) cls_rows_kernel<<<(nb * kDim + 255) / 256, 256, 0, st>>>(x, cls, pos, nb, tokens);
    else if (prec == VIT_PREC_FP16)
        cls_rows_ln_kernel<__half><<<(nb + 7) / 8, 256, 0, st>>>(x, static_cast<__half*>(xn), pstats, stats_rows, cls, pos, nb, tokens, cls_rows32);
    else
        cls_rows_ln_kernel<__nv_bfloat16><<<(nb + 7) / 8, 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(xn), pstats, stats_rows, cls, pos, nb, tokens, cls_rows32);
    return check_launch("cls_rows");
}

int launch_convert_from_f32(int prec, const float* src, void* dst, size_t n, cudaStream_t st) {
    const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
    if (prec == VIT_PREC_FP16) convert_from_f32_kernel<__half><<<grid, 256, 0, st>>>(src, static_cast<__half*>(dst), n);
    else convert_from_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    return check_launch("convert");
}
int launch_convert_to_f32(int prec, const void* src, float* dst, size_t n, cudaStream_t st) {
    const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
    if (prec == VIT_PREC_FP16) convert_to_f32_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(src), dst, n);
    else convert_to_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), dst, n);
    return check_launch("convert");
}
int launch_round_tf32(const float* src, float* dst, size_t n, cudaStream_t st) {
    round_tf32_kernel<<<static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16)), 256, 0, st>>>(src, dst, n);
    return check_launch("round_tf32");
}

// ------------------------------------------------------------------------------------ device check
int check_device(int dev, int* sm_count) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return set_err(VIT_E_NODEVICE, "no CUDA device available (this engine has no CPU fallback)");
    }
    if (dev < 0 || dev >= count) return set_err(VIT_E_NODEVICE, "device %d requested, %d present", dev, count);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return set_err(VIT_E_NODEVICE, "device %d (%s) is sm_%d%d; this engine is built for sm_100a only", dev, prop.name,
                       prop.major, prop.minor);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return 0;
}

int watchdog_or_cuda_error(cudaError_t e, const char* what) {
    // After a trap the context is dead (the watchdog flag cannot be read back); report it from the error code.
    if (e == cudaErrorLaunchFailure || e == cudaErrorIllegalInstruction || e == cudaErrorAssert)
        return set_err(VIT_E_DEVICE_TRAP, "%s: device trap (%s) -- kernel watchdog or fault", what, cudaGetErrorString(e));
    return set_err(VIT_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------ engine state
struct timerEvents_t {
    cudaEvent_t start = nullptr, stop = nullptr;
};

// One operand precision's GEMM weights of an encoder layer
struct OperandW {
    void *qkv_w = nullptr, *out_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr;
    // LayerNorm folded into in_proj / mlp_0: W' = ln_w (.) W (rounded), column sums of the rounded W'
    void *qkv_wf = nullptr, *fc1_wf = nullptr;
    float *qkv_s = nullptr, *fc1_s = nullptr;
    CUtensorMap tm_qkv_w, tm_out_w, tm_fc1_w, tm_fc2_w, tm_qkv_wf, tm_fc1_wf;
};
struct LayerW {
    float *ln1_w, *ln1_b, *qkv_b, *out_b, *ln2_w, *ln2_b, *fc1_b, *fc2_b;  // fp32, verbatim
    float *qkv_c, *fc1_c;                                                   // c = bias + W ln_b (fp32, from the unrounded W)
    OperandW op[2];                                                         // [VIT_PREC_BF16], [VIT_PREC_FP16]
};

// Bump allocator over the weight arena.  A dry run (base == nullptr) only adds up the size.
struct Arena {
    uint8_t* base = nullptr;
    size_t off = 0;
    template <typename P>
    void take(P** p, size_t bytes) {
        *p = reinterpret_cast<P*>(base + off);
        off += (bytes + 255) & ~static_cast<size_t>(255);
    }
};

struct DeviceCtx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    std::vector<void*> allocs;
    // ---- weights: ONE device allocation (the arena), laid out by layout_weights(); this is what the cache file holds
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    float *cls = nullptr, *conv_b = nullptr, *pos = nullptr, *lnf_w = nullptr, *lnf_b = nullptr, *head_w = nullptr,
          *head_b = nullptr;
    float* conv_w = nullptr;   // fp32 rounded to tf32 (conv_proj runs as kind::tf32 on the raw image)
    CUtensorMap tm_conv_w, tm_pos;
    LayerW layer[kDepth];
    // ---- workspace (capacity = max_batch images)
    float* images[2] = {nullptr, nullptr};
    void *xn = nullptr, *qkv = nullptr, *ao = nullptr, *hid = nullptr;
    float *x = nullptr, *cls_ln = nullptr, *logits = nullptr;
    float* h_stage[2] = {nullptr, nullptr};  // pinned staging for scattered host images (vit_cuda_forward_scattered), lazily allocated
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};   // the H2D copy out of h_stage[i] has completed
    float* h_logits[2] = {nullptr, nullptr};        // pinned staging for logits when the caller's buffer is pageable
    cudaEvent_t ev_logits[2] = {nullptr, nullptr};
    int pend_first[2] = {0, 0}, pend_count[2] = {0, 0};  // logits waiting in h_logits[i] for their copy to the caller
    // compact [batch]-row buffers of the pruned last-layer tail (class rows only)
    void *ao_c = nullptr, *xn_c = nullptr, *hid_c = nullptr;
    float* x_c = nullptr;
    float2* pstats_c = nullptr;
    size_t stats_rows_c = 0;
    float2* pstats = nullptr;  // [6][max rows] partial (sum, sum of squares) of the residual rows (LN folding)
    size_t stats_rows = 0;
    // activation tensor maps, rebuilt when the pass size or the operand precision changes: row extent = rows
    // actually in use, so TMA zero-fills loads and clips stores past the last image
    int maps_nb = -1, maps_prec = -1;
    CUtensorMap tm_ao_c, tm_x_c, tm_xn_c, tm_hid_c, tm_hid_c32;
    CUtensorMap tm_x3, tm_xn3 /* per-image 3-D views for conv_proj */, tm_xn, tm_ao, tm_hid, tm_qkv_st, tm_hid32, tm_qkv_st32 /* 32-row store boxes */, tm_x, tm_q /* attention loads */, tm_kv /* attention store */, tm_kv32 /* attention store, 32-row boxes */;
    // 5-D views of the image buffers conv_proj has been asked to read (small cache: the two engine buffers, a caller's own)
    struct ImgMap {
        const float* ptr = nullptr;
        int nb = 0;
        CUtensorMap map;
    };
    ImgMap img_maps[4];
    int img_map_next = 0;
    size_t ws_bytes = 0;
    // device status word (g_status_flags of this device) and its pinned host mirror, read back after a pass
    unsigned int* d_flags = nullptr;
    unsigned int* h_flags = nullptr;
    bool pending_fast_softmax = false, pending_fp16 = false;   // work enqueued since the flags were last read
    // copy rate vs kernel rate of the passes of the last host calls (pinned input), for the pass schedule's growth factor:
    // timing events around every pass's H2D copy and kernels, read after the call; exponential averages in ms per image
    std::vector<cudaEvent_t> tm_events;      // 4 per pass: copy start / stop, kernels start / stop
    double h2d_ms_per_image = 0, kernel_ms_per_image = 0;
    double pass_fixed_ms = 0, pass_ms_per_image = 0;
    cudaEvent_t img_release = nullptr;   // set by run_shard: recorded right behind conv_proj, the only kernel that reads the image buffer   // kernels of a pass ~ pass_fixed_ms + pass_ms_per_image * images (fit over a call's passes)
    char err[512] = "";   // failure text of this slot's feeding thread
    // optional per-kernel-category timing (vit_cuda_profile_*): event pairs around launches
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[VIT_PROF_NCAT];
    size_t prof_used[VIT_PROF_NCAT] = {};
    timerEvents_t timer;
    struct Graph {   // captured launch sequence of one small pass
        int nb = 0, prec = 0;
        const float* images = nullptr;
        float* logits = nullptr;
        bool attn_exact = false, ln_fused = true, prune_last = true, pdl = true, res16 = false;
        long long launches = 0;
        cudaGraphExec_t exec = nullptr;
    };
    std::vector<Graph> graphs;
};

struct Engine {
    bool up = false;
    bool profiling = false;
    long long attn_fallbacks = 0;   // forwards repeated with the exact softmax
    long long prec_fallbacks = 0;   // forwards repeated with BF16 operands (VIT_PREC_AUTO)
    int img = 0, grid = 0, patches = 0, tokens = 0, max_batch = 0;
    int policy = VIT_PREC_AUTO;     // what the caller asked for
    int prec = VIT_PREC_FP16;       // operand precision of the next pass (VIT_PREC_BF16 / VIT_PREC_FP16)
    bool resident[2] = {false, false};   // which operand sets the arena holds
    bool exact_on_trial = false;         // vit_cuda_sync switched to the exact softmax when BOTH flags were up (see decide_retry)
    std::vector<DeviceCtx> ctx;
};
Engine g_eng;

int dev_alloc(DeviceCtx& c, void** p, size_t bytes, bool zero) {
    CU_TRY(cudaMalloc(p, bytes));
    c.allocs.push_back(*p);
    c.ws_bytes += bytes;
    if (zero) CU_TRY(cudaMemsetAsync(*p, 0, bytes, c.stream));
    return 0;
}

size_t tensor_numel(int idx, int img) {
    const size_t g = img / kPatch, tokens = g * g + 1;
    switch (idx) {
        case 0: case 2: case 148: case 149: return kDim;
        case 1: return static_cast<size_t>(kDim) * 3 * kPatch * kPatch;
        case 3: return tokens * kDim;
        case 150: return static_cast<size_t>(kClasses) * kDim;
        case 151: return kClasses;
    }
    switch ((idx - 4) % 12) {
        case 2: return static_cast<size_t>(3) * kDim * kDim;
        case 3: return 3 * kDim;
        case 4: return static_cast<size_t>(kDim) * kDim;
        case 8: case 10: return static_cast<size_t>(kHidden) * kDim;
        case 9: return kHidden;
        default: return kDim;
    }
}

// Assigns every weight pointer of the context inside the arena, in a fixed order that depends only on
// (image size, resident precisions) -- the cache file is the arena's bytes.
void layout_weights(DeviceCtx& c, Arena& a, const Engine& e) {
    const size_t f = sizeof(float);
    a.take(&c.cls, kDim * f);
    a.take(&c.conv_w, static_cast<size_t>(kDim) * kDim * f);
    a.take(&c.conv_b, kDim * f);
    a.take(&c.pos, static_cast<size_t>(e.tokens) * kDim * f);
    for (int l = 0; l < kDepth; ++l) {
        LayerW& L = c.layer[l];
        a.take(&L.ln1_w, kDim * f);
        a.take(&L.ln1_b, kDim * f);
        a.take(&L.qkv_b, 3 * kDim * f);
        a.take(&L.out_b, kDim * f);
        a.take(&L.ln2_w, kDim * f);
        a.take(&L.ln2_b, kDim * f);
        a.take(&L.fc1_b, kHidden * f);
        a.take(&L.fc2_b, kDim * f);
        a.take(&L.qkv_c, 3 * kDim * f);
        a.take(&L.fc1_c, kHidden * f);
        for (int pr = 0; pr < 2; ++pr) {
            if (!e.resident[pr]) continue;
            OperandW& o = L.op[pr];
            a.take(&o.qkv_w, static_cast<size_t>(3) * kDim * kDim * 2);
            a.take(&o.out_w, static_cast<size_t>(kDim) * kDim * 2);
            a.take(&o.fc1_w, static_cast<size_t>(kHidden) * kDim * 2);
            a.take(&o.fc2_w, static_cast<size_t>(kHidden) * kDim * 2);
            a.take(&o.qkv_wf, static_cast<size_t>(3) * kDim * kDim * 2);
            a.take(&o.fc1_wf, static_cast<size_t>(kHidden) * kDim * 2);
            a.take(&o.qkv_s, 3 * kDim * f);
            a.take(&o.fc1_s, kHidden * f);
        }
    }
    a.take(&c.lnf_w, kDim * f);
    a.take(&c.lnf_b, kDim * f);
    a.take(&c.head_w, static_cast<size_t>(kClasses) * kDim * f);
    a.take(&c.head_b, kClasses * f);
}

int make_weight_maps(DeviceCtx& c, const Engine& e) {
    for (int l = 0; l < kDepth; ++l)
        for (int pr = 0; pr < 2; ++pr) {
            if (!e.resident[pr]) continue;
            OperandW& o = c.layer[l].op[pr];
            VIT_TRY(make_tmap(&o.tm_qkv_w, pr, o.qkv_w, kDim, 3 * kDim, GEMM_BK, 128));
            VIT_TRY(make_tmap(&o.tm_out_w, pr, o.out_w, kDim, kDim, GEMM_BK, 128));
            VIT_TRY(make_tmap(&o.tm_fc1_w, pr, o.fc1_w, kDim, kHidden, GEMM_BK, 128));
            VIT_TRY(make_tmap(&o.tm_fc2_w, pr, o.fc2_w, kHidden, kDim, GEMM_BK, 128));
            VIT_TRY(make_tmap(&o.tm_qkv_wf, pr, o.qkv_wf, kDim, 3 * kDim, GEMM_BK, 128));
            VIT_TRY(make_tmap(&o.tm_fc1_wf, pr, o.fc1_wf, kDim, kHidden, GEMM_BK, 128));
        }
    VIT_TRY(make_tmap_f32(&c.tm_conv_w, c.conv_w, kDim, kDim, 128));
    VIT_TRY(make_tmap_f32(&c.tm_pos, c.pos, kDim, e.tokens, GEMM_BM));
    return 0;
}

void destroy_ctx(DeviceCtx& c) {
    if (c.device < 0) return;
    cudaSetDevice(c.device);
    if (c.stream) cudaStreamSynchronize(c.stream);
    if (c.copy_stream) cudaStreamSynchronize(c.copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (c.h_stage[i]) cudaFreeHost(c.h_stage[i]);
        if (c.ev_stage[i]) cudaEventDestroy(c.ev_stage[i]);
        if (c.h_logits[i]) cudaFreeHost(c.h_logits[i]);
        if (c.ev_logits[i]) cudaEventDestroy(c.ev_logits[i]);
    }
    if (c.h_flags) cudaFreeHost(c.h_flags);
    for (cudaEvent_t ev : c.tm_events) cudaEventDestroy(ev);
    for (auto& g : c.graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c.graphs.clear();
    for (void* p : c.allocs) cudaFree(p);
    c.allocs.clear();
    if (c.arena) cudaFree(c.arena);
    for (int i = 0; i < 2; ++i) {
        if (c.ev_h2d[i]) cudaEventDestroy(c.ev_h2d[i]);
        if (c.ev_done[i]) cudaEventDestroy(c.ev_done[i]);
    }
    for (auto& v : c.prof_events)
        for (auto& pr : v) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    if (c.timer.start) cudaEventDestroy(c.timer.start);
    if (c.timer.stop) cudaEventDestroy(c.timer.stop);
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    c = DeviceCtx();
}

// Streams, events, status word and the (empty) weight arena of one slot.
int open_ctx(DeviceCtx& c, int device, const Engine& e) {
    c.device = device;
    VIT_TRY(check_device(device, &c.sm_count));
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&c.ev_h2d[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c.ev_done[i], cudaEventDisableTiming));
    }
    CU_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&c.d_flags), g_status_flags));
    CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c.h_flags), sizeof(unsigned int), cudaHostAllocPortable));
    *c.h_flags = 0;
    CU_TRY(cudaMemsetAsync(c.d_flags, 0, sizeof(unsigned int), c.stream));
    Arena dry;
    layout_weights(c, dry, e);
    c.arena_bytes = dry.off;
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&c.arena), c.arena_bytes));
    Arena a;
    a.base = c.arena;
    layout_weights(c, a, e);
    return 0;
}

// Fills the arena from the 152 fp32 host tensors: small tensors verbatim, GEMM operands converted once per resident
// precision, in_proj / mlp_0 additionally folded with their LayerNorm, conv_proj.weight rounded to tf32.
int fill_arena_from_tensors(DeviceCtx& c, const vit_tensor* w, const Engine& e) {
    cudaStream_t st = c.stream;
    float* scratch = nullptr;
    CU_TRY(cudaMalloc(&scratch, static_cast<size_t>(kHidden) * kDim * sizeof(float)));
    auto put_f32 = [&](float* dst, const vit_tensor& t) -> int {
        CU_TRY(cudaMemcpyAsync(dst, t.data, t.size * sizeof(float), cudaMemcpyHostToDevice, st));
        return 0;
    };
    // fp32 tensor -> scratch, then one conversion per resident precision; `fold`: also W' = ln_w (.) W etc.
    auto put_operand = [&](const vit_tensor& t, void* OperandW::*plain, LayerW& L, void* OperandW::*folded, float* OperandW::*colsum,
                           const float* ln_w, const float* ln_b, const float* bias, float* cvec, int N) -> int {
        CU_TRY(cudaMemcpyAsync(scratch, t.data, t.size * sizeof(float), cudaMemcpyHostToDevice, st));
        for (int pr = 0; pr < 2; ++pr) {
            if (!e.resident[pr]) continue;
            OperandW& o = L.op[pr];
            VIT_TRY(launch_convert_from_f32(pr, scratch, o.*plain, t.size, st));
            if (folded) VIT_TRY(launch_fold_ln(pr, scratch, ln_w, ln_b, bias, o.*folded, o.*colsum, cvec, N, st));
        }
        return 0;
    };
    int rc = 0;
    do {
        if ((rc = put_f32(c.cls, w[0]))) break;
        CU_TRY(cudaMemcpyAsync(scratch, w[1].data, w[1].size * sizeof(float), cudaMemcpyHostToDevice, st));
        if ((rc = launch_round_tf32(scratch, c.conv_w, w[1].size, st))) break;
        if ((rc = put_f32(c.conv_b, w[2]))) break;
        if ((rc = put_f32(c.pos, w[3]))) break;
        for (int l = 0; l < kDepth && !rc; ++l) {
            const vit_tensor* lw = w + 4 + 12 * l;
            LayerW& L = c.layer[l];
            if ((rc = put_f32(L.ln1_w, lw[0]))) break;
            if ((rc = put_f32(L.ln1_b, lw[1]))) break;
            if ((rc = put_f32(L.qkv_b, lw[3]))) break;
            if ((rc = put_f32(L.out_b, lw[5]))) break;
            if ((rc = put_f32(L.ln2_w, lw[6]))) break;
            if ((rc = put_f32(L.ln2_b, lw[7]))) break;
            if ((rc = put_f32(L.fc1_b, lw[9]))) break;
            if ((rc = put_f32(L.fc2_b, lw[11]))) break;
            if ((rc = put_operand(lw[2], &OperandW::qkv_w, L, &OperandW::qkv_wf, &OperandW::qkv_s, L.ln1_w, L.ln1_b, L.qkv_b, L.qkv_c, 3 * kDim))) break;
            if ((rc = put_operand(lw[4], &OperandW::out_w, L, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0))) break;
            if ((rc = put_operand(lw[8], &OperandW::fc1_w, L, &OperandW::fc1_wf, &OperandW::fc1_s, L.ln2_w, L.ln2_b, L.fc1_b, L.fc1_c, kHidden))) break;
            if ((rc = put_operand(lw[10], &OperandW::fc2_w, L, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0))) break;
        }
        if (rc) break;
        if ((rc = put_f32(c.lnf_w, w[148]))) break;
        if ((rc = put_f32(c.lnf_b, w[149]))) break;
        if ((rc = put_f32(c.head_w, w[150]))) break;
        if ((rc = put_f32(c.head_b, w[151]))) break;
    } while (0);
    const cudaError_t se = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (rc) return rc;
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, "weight upload");
    return 0;
}

// Workspaces.  Zero-filled once: attention may read (and multiply by P == 0) rows past the last image of a
// pass, which must therefore always hold finite values.
int alloc_workspace(DeviceCtx& c, const Engine& e) {
    CU_TRY(cudaSetDevice(c.device));
    const size_t B = e.max_batch, rows = B * e.tokens;
    const size_t img_elems = static_cast<size_t>(3) * e.img * e.img;
    for (int i = 0; i < 2; ++i) VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.images[i]), B * img_elems * 4, false));
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.x), rows * kDim * 4, true));
    VIT_TRY(dev_alloc(c, &c.xn, rows * kDim * 2, true));
    VIT_TRY(dev_alloc(c, &c.qkv, rows * 3 * kDim * 2, true));
    VIT_TRY(dev_alloc(c, &c.ao, rows * kDim * 2, true));
    VIT_TRY(dev_alloc(c, &c.hid, rows * kHidden * 2, true));
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.cls_ln), B * kDim * 4, true));
    VIT_TRY(dev_alloc(c, &c.ao_c, B * kDim * 2, true));
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.x_c), B * kDim * 4, true));
    VIT_TRY(dev_alloc(c, &c.xn_c, B * kDim * 2, true));
    VIT_TRY(dev_alloc(c, &c.hid_c, B * kHidden * 2, true));
    c.stats_rows_c = (B + 255) / 256 * 256;
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.pstats_c), 6 * c.stats_rows_c * sizeof(float2), true));
    c.stats_rows = (rows + 255) / 256 * 256;
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.pstats), 6 * c.stats_rows * sizeof(float2), true));
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.logits), B * kClasses * 4, true));
    CU_TRY(cudaStreamSynchronize(c.stream));
    return 0;
}

int ensure_maps(DeviceCtx& c, const Engine& e, int nb, int prec) {
    if (c.maps_nb == nb && c.maps_prec == prec) return 0;
    const uint64_t rows = static_cast<uint64_t>(nb) * e.tokens;
    const uint32_t erows = embed_rows_per_cta(e.img);
    VIT_TRY(make_tmap(&c.tm_xn, prec, c.xn, kDim, rows, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_ao, prec, c.ao, kDim, rows, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_hid, prec, c.hid, kHidden, rows, GEMM_BK, GEMM_BM));     // mlp_0 store + mlp_3 A load
    VIT_TRY(make_tmap(&c.tm_qkv_st, prec, c.qkv, 3 * kDim, rows, GEMM_BK, GEMM_BM)); // in_proj store
    VIT_TRY(make_tmap_3d_f32(&c.tm_x3, c.x, kDim, e.tokens, nb, erows));
    VIT_TRY(make_tmap_3d(&c.tm_xn3, prec, c.xn, kDim, e.tokens, nb, erows));
    VIT_TRY(make_tmap(&c.tm_ao_c, prec, c.ao_c, kDim, nb, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap_f32(&c.tm_x_c, c.x_c, kDim, nb, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_xn_c, prec, c.xn_c, kDim, nb, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_hid_c, prec, c.hid_c, kHidden, nb, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_hid_c32, prec, c.hid_c, kHidden, nb, GEMM_BK, 32));
    VIT_TRY(make_tmap(&c.tm_hid32, prec, c.hid, kHidden, rows, GEMM_BK, 32));
    VIT_TRY(make_tmap(&c.tm_qkv_st32, prec, c.qkv, 3 * kDim, rows, GEMM_BK, 32));
    VIT_TRY(make_tmap_f32(&c.tm_x, c.x, kDim, rows, GEMM_BM));                       // residual load + store
    VIT_TRY(make_tmap(&c.tm_q, prec, c.qkv, 3 * kDim, rows, ATTN_DH, attention_load_box_rows(e.tokens)));  // Q/K/V boxes
    VIT_TRY(make_tmap_3d(&c.tm_kv, prec, c.ao, kDim, e.tokens, nb, 128));                                   // per-image output tiles
    VIT_TRY(make_tmap_3d(&c.tm_kv32, prec, c.ao, kDim, e.tokens, nb, 32));
    c.maps_nb = nb;
    c.maps_prec = prec;
    return 0;
}

int image_map(DeviceCtx& c, const Engine& e, const float* d_images, int nb, const CUtensorMap** out) {
    for (auto& m : c.img_maps)
        if (m.ptr == d_images && m.nb == nb) {
            *out = &m.map;
            return 0;
        }
    if (reinterpret_cast<uintptr_t>(d_images) & 15) return set_err(VIT_E_ARG, "device images must be 16-byte aligned");
    DeviceCtx::ImgMap& m = c.img_maps[c.img_map_next];
    c.img_map_next = (c.img_map_next + 1) % 4;
    VIT_TRY(make_tmap_image5d(&m.map, d_images, e.img, nb));
    m.ptr = d_images;
    m.nb = nb;
    *out = &m.map;
    return 0;
}

// RAII marker: when profiling is on, brackets one launch with an event pair of its category.
struct ProfScope {
    DeviceCtx& c;
    int cat;
    bool on;
    ProfScope(DeviceCtx& c_, bool on_, int cat_) : c(c_), cat(cat_), on(on_) {
        if (!on) return;
        auto& v = c.prof_events[cat];
        if (c.prof_used[cat] == v.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            v.emplace_back(a, b);
        }
        cudaEventRecord(v[c.prof_used[cat]].first, c.stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(c.prof_events[cat][c.prof_used[cat]].second, c.stream);
        ++c.prof_used[cat];
    }
};

// The switches a pass is enqueued with (snapshotted once per host call, so that a pass is self-consistent)
struct PassMode {
    int prec;
    bool attn_exact, ln_fused, prune_last, pdl;
    bool res16;   // residual stream in the operand type, class rows with an fp32 master copy (FP16 operands + folded LayerNorm only), VIT_OPT_RESIDUAL16
};
PassMode current_mode(const Engine& e) {
    const bool fused = g_opt.ln_fused.load() != 0;
    return PassMode{e.prec, g_opt.attn_exact.load() != 0, fused, g_opt.prune_last.load() != 0, g_opt.pdl.load() != 0,
                    g_opt.residual16.load() != 0 && fused && e.prec == VIT_PREC_FP16};
}

// One encoder block (Encoder, ViT_seq.c:271-302) for nb images on c.stream: in_proj (LayerNorm 1 folded in) -> attention ->
// out_proj + residual -> mlp_0 (LayerNorm 2 folded in, GELU) -> mlp_3 + residual, on c.x in place.  `tail_for_head`: this is the
// last block of a forward -- with class-row pruning everything behind its attention runs on the class rows only (returns 1:
// the result is in c.x_c), and its mlp_3 need not emit the copy / statistics for a next block.  Returns 0 normally, < 0 on error.
int enqueue_encoder_layer(DeviceCtx& c, const Engine& e, const PassMode& m, int l, int nb, bool tail_for_head) {
    cudaStream_t st = c.stream;
    const int prec = m.prec;
    const int rows = nb * e.tokens;
    const bool pf = e.profiling;
    const bool fused = m.ln_fused;
    const int stats_rows = static_cast<int>(c.stats_rows);
    AttnParams ap{nb, e.tokens, (e.tokens + 15) / 16 * 16, c.ao, 0.125f * 1.4426950408889634f, nullptr};
    const LayerW& L = c.layer[l];
    const OperandW& W = L.op[prec];
    if (!fused) {
        ProfScope ps(c, pf, VIT_PROF_LAYERNORM);
        VIT_TRY(launch_layernorm(prec, c.x, L.ln1_w, L.ln1_b, c.xn, rows, st));
    }
    {
        ProfScope ps(c, pf, VIT_PROF_QKV_GEMM);
        GemmParams p{rows, 3 * kDim, kDim, L.qkv_b, c.qkv, nullptr, 0, 0, 0, 2 * kDim};  // V block stored as bf16
        if (fused) {
            p.bias = L.qkv_c;
            p.colsum = W.qkv_s;
            p.stats_in = c.pstats;
            p.stats_parts = 6;
            p.stats_rows = stats_rows;
            VIT_TRY(launch_gemm_staged_ln<EPI_BIAS>(prec, c.tm_xn, W.tm_qkv_wf, c.tm_qkv_st, c.tm_qkv_st32, p, c.sm_count, st));
        } else VIT_TRY(launch_gemm_staged<EPI_BIAS>(prec, c.tm_xn, W.tm_qkv_w, c.tm_qkv_st, c.tm_qkv_st32, p, c.sm_count, st));
    }
    if (tail_for_head && fused && m.prune_last) {
        // Last layer: only the class token reaches the head, and no token reads another one after the attention.
        // Attention for the class query alone (all keys and values), then out_proj / LayerNorm / MLP on the compact
        // [nb][768] class rows: 1/197 of the rows of the other layers.
        {
            ProfScope ps(c, pf, VIT_PROF_ATTENTION);
            if (prec == VIT_PREC_FP16)
                cls_attention_kernel<__half><<<nb, 384, 0, st>>>(static_cast<const uint16_t*>(c.qkv), m.res16 ? nullptr : c.x, nullptr,   // RES16: c.x_c already
                                                                 static_cast<__half*>(c.ao_c), c.x_c, e.tokens);                       // holds the fp32 class rows
            else
                cls_attention_kernel<__nv_bfloat16><<<nb, 384, 0, st>>>(static_cast<const uint16_t*>(c.qkv), c.x, nullptr, static_cast<__nv_bfloat16*>(c.ao_c), c.x_c, e.tokens);
            VIT_TRY(check_launch("cls_attention"));
        }
        const int srows_c = static_cast<int>(c.stats_rows_c);
        {
            ProfScope ps(c, pf, VIT_PROF_OUT_GEMM);
            GemmParams p{nb, kDim, kDim, L.out_b, c.x_c, c.x_c};
            p.stats_out = c.pstats_c;
            p.stats_rows = srows_c;
            VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_RESIDUAL>(prec, c.tm_ao_c, W.tm_out_w, c.tm_x_c, c.tm_xn_c, p, c.sm_count, st));
        }
        {
            ProfScope ps(c, pf, VIT_PROF_FC1_GEMM);
            GemmParams p{nb, kHidden, kDim, L.fc1_c, c.hid_c, nullptr};
            p.colsum = W.fc1_s;
            p.stats_in = c.pstats_c;
            p.stats_parts = 6;
            p.stats_rows = srows_c;
            VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_GELU>(prec, c.tm_xn_c, W.tm_fc1_wf, c.tm_hid_c, c.tm_hid_c32, p, c.sm_count, st));
        }
        {
            ProfScope ps(c, pf, VIT_PROF_FC2_GEMM);
            GemmParams p{nb, kDim, kHidden, L.fc2_b, c.x_c, c.x_c};
            VIT_TRY(launch_gemm_staged<EPI_BIAS_RESIDUAL>(prec, c.tm_hid_c, W.tm_fc2_w, c.tm_x_c, c.tm_x_c, p, c.sm_count, st));
        }
        return 1;
    }
    {
        ProfScope ps(c, pf, VIT_PROF_ATTENTION);
        VIT_TRY(launch_attention(prec, c.tm_q, c.tm_kv, c.tm_kv32, ap, c.sm_count, st, m.attn_exact));
    }
    {
        ProfScope ps(c, pf, VIT_PROF_OUT_GEMM);
        GemmParams p{rows, kDim, kDim, L.out_b, c.x, c.x};
        if (m.res16) {
            p.stats_out = c.pstats;
            p.stats_rows = stats_rows;
            p.tokens = e.tokens;
            p.cls_rows32 = c.x_c;
            VIT_TRY(launch_gemm_residual16(prec, c.tm_ao, W.tm_out_w, c.tm_xn, p, c.sm_count, st));
        } else if (fused) {
            p.stats_out = c.pstats;
            p.stats_rows = stats_rows;
            VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_RESIDUAL>(prec, c.tm_ao, W.tm_out_w, c.tm_x, c.tm_xn, p, c.sm_count, st));
        } else VIT_TRY(launch_gemm_staged<EPI_BIAS_RESIDUAL>(prec, c.tm_ao, W.tm_out_w, c.tm_x, c.tm_x, p, c.sm_count, st));
    }
    if (!fused) {
        ProfScope ps(c, pf, VIT_PROF_LAYERNORM);
        VIT_TRY(launch_layernorm(prec, c.x, L.ln2_w, L.ln2_b, c.xn, rows, st));
    }
    {
        ProfScope ps(c, pf, VIT_PROF_FC1_GEMM);
        GemmParams p{rows, kHidden, kDim, L.fc1_b, c.hid, nullptr};
        if (fused) {
            p.bias = L.fc1_c;
            p.colsum = W.fc1_s;
            p.stats_in = c.pstats;
            p.stats_parts = 6;
            p.stats_rows = stats_rows;
            VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_GELU>(prec, c.tm_xn, W.tm_fc1_wf, c.tm_hid, c.tm_hid32, p, c.sm_count, st));
        } else VIT_TRY(launch_gemm_staged<EPI_BIAS_GELU>(prec, c.tm_xn, W.tm_fc1_w, c.tm_hid, c.tm_hid32, p, c.sm_count, st));
    }
    {
        ProfScope ps(c, pf, VIT_PROF_FC2_GEMM);
        GemmParams p{rows, kDim, kHidden, L.fc2_b, c.x, c.x};
        if (m.res16) {   // (the last block's statistics are not needed, but one kernel variant serves all blocks)
            p.stats_out = c.pstats;
            p.stats_rows = stats_rows;
            p.tokens = e.tokens;
            p.cls_rows32 = c.x_c;
            VIT_TRY(launch_gemm_residual16(prec, c.tm_hid, W.tm_fc2_w, c.tm_xn, p, c.sm_count, st));
        } else if (fused && !tail_for_head) {   // the last layer's output only feeds the class-row LayerNorm of the head
            p.stats_out = c.pstats;
            p.stats_rows = stats_rows;
            VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_RESIDUAL>(prec, c.tm_hid, W.tm_fc2_w, c.tm_x, c.tm_xn, p, c.sm_count, st));
        } else VIT_TRY(launch_gemm_staged<EPI_BIAS_RESIDUAL>(prec, c.tm_hid, W.tm_fc2_w, c.tm_x, c.tm_x, p, c.sm_count, st));
    }
    return 0;
}

// Enqueue the whole forward for nb images already resident in d_images (fp32 NCHW) on c.stream.
int enqueue_forward_kernels(DeviceCtx& c, const Engine& e, const PassMode& m, const float* d_images, int nb, float* d_logits) {
    cudaStream_t st = c.stream;
    const int prec = m.prec;
    VIT_TRY(ensure_maps(c, e, nb, prec));
    const CUtensorMap* tm_img = nullptr;
    VIT_TRY(image_map(c, e, d_images, nb, &tm_img));
    const bool pf = e.profiling;
    // LayerNorm folded into the GEMMs (default): in_proj / mlp_0 read the operand-precision copy of the raw
    // residual rows (c.xn) with the per-row statistics (c.pstats) that the previous residual GEMM -- for
    // layer 0 conv_proj and the class-row kernel -- left behind.  Otherwise: a LayerNorm kernel before each.
    const bool fused = m.ln_fused;
    const int stats_rows = static_cast<int>(c.stats_rows);
    {   // class-token rows (class_token + pos_emb, ViT_seq.c:72-101)
        ProfScope ps(c, pf, VIT_PROF_CLASS_ROWS);
        // (RES16: the class rows additionally start their fp32 master copy in the compact buffer c.x_c)
        VIT_TRY(launch_cls_rows(prec, fused, c.x, c.xn, c.pstats, stats_rows, c.cls, c.pos, nb, e.tokens, st, m.res16 ? c.x_c : nullptr));
    }
    {   // conv_proj: tf32 GEMM straight from the fp32 image; class_token offset / pos_emb / token layout are TMA addressing
        ProfScope ps(c, pf, VIT_PROF_EMBED_GEMM);
        GemmParams p{nb * embed_tiles_per_image(e.img) * 256, kDim, kDim, c.conv_b, c.x, nullptr, e.patches, e.tokens, e.grid};
        if (fused) {
            p.stats_out = c.pstats;
            p.stats_rows = stats_rows;
        }
        VIT_TRY(launch_gemm_embed(prec, fused, *tm_img, c.tm_conv_w, c.tm_x3, c.tm_xn3, c.tm_pos, p, c.sm_count, st));
        if (c.img_release) {   // the image buffer is free for the copy of the pass after next (never set while capturing a graph)
            CU_TRY(cudaEventRecord(c.img_release, st));
            c.img_release = nullptr;
        }
    }
    bool pruned_tail = false;
    for (int l = 0; l < kDepth; ++l) {
        const int r = enqueue_encoder_layer(c, e, m, l, nb, l == kDepth - 1);
        if (r < 0) return r;
        pruned_tail = r == 1;
    }
    ProfScope ps(c, pf, VIT_PROF_HEAD);
    if (pruned_tail) head_ln_kernel<float><<<(nb + 7) / 8, 256, 0, st>>>(c.x_c, c.lnf_w, c.lnf_b, c.cls_ln, nb, 1);
    else if (m.res16) head_ln_kernel<float><<<(nb + 7) / 8, 256, 0, st>>>(c.x_c, c.lnf_w, c.lnf_b, c.cls_ln, nb, 1);   // the fp32 master copies of the class rows
    else head_ln_kernel<float><<<(nb + 7) / 8, 256, 0, st>>>(c.x, c.lnf_w, c.lnf_b, c.cls_ln, nb, e.tokens);
    VIT_TRY(check_launch("head_ln"));
    head_gemm_kernel<<<dim3((kClasses + HEAD_CLASSES - 1) / HEAD_CLASSES, std::min((nb + HEAD_IMGS - 1) / HEAD_IMGS, 32)), 256, 0, st>>>(c.cls_ln, c.head_w, c.head_b, d_logits, nb,
                                                                                kClasses);
    return check_launch("head_gemm");
}

// Small passes (batch-1 latency, BASELINE.json configs[1]) are launch bound: ~65 kernels of 5-20 us each.  Their
// launch sequence is captured once per (pass size, buffers, mode) into a CUDA graph and replayed with a
// single launch.  VIT_OPT_GRAPHS = 0 disables this.
constexpr int kGraphMaxBatch = 8;
int enqueue_forward(DeviceCtx& c, const Engine& e, const PassMode& m, const float* d_images, int nb, float* d_logits) {
    if (!m.attn_exact) c.pending_fast_softmax = true;
    if (m.prec == VIT_PREC_FP16) c.pending_fp16 = true;
    if (nb > kGraphMaxBatch || e.profiling || !g_opt.graphs.load())
        return enqueue_forward_kernels(c, e, m, d_images, nb, d_logits);
    for (auto& g : c.graphs)
        if (g.nb == nb && g.prec == m.prec && g.images == d_images && g.logits == d_logits && g.attn_exact == m.attn_exact &&
            g.ln_fused == m.ln_fused && g.prune_last == m.prune_last && g.pdl == m.pdl && g.res16 == m.res16) {
            CU_TRY(cudaGraphLaunch(g.exec, c.stream));
            g_launches.fetch_add(g.launches, std::memory_order_relaxed);
            return 0;
        }
    // first time: run it once the ordinary way (sets the kernels' attributes, builds the tensor maps), then capture
    VIT_TRY(enqueue_forward_kernels(c, e, m, d_images, nb, d_logits));
    if (c.graphs.size() >= 16) return 0;   // a caller cycling through many buffers: stay with plain launches
    cudaGraph_t graph = nullptr;
    const long long before = g_launches.load();
    CU_TRY(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_forward_kernels(c, e, m, d_images, nb, d_logits);
    const cudaError_t ce = cudaStreamEndCapture(c.stream, &graph);
    const long long captured = g_launches.load() - before;
    g_launches.fetch_sub(captured, std::memory_order_relaxed);   // captured, not launched
    if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (ce != cudaSuccess) return set_err(VIT_E_CUDA, "graph capture: %s", cudaGetErrorString(ce));
    DeviceCtx::Graph g;
    g.nb = nb;
    g.prec = m.prec;
    g.images = d_images;
    g.logits = d_logits;
    g.attn_exact = m.attn_exact;
    g.ln_fused = m.ln_fused;
    g.prune_last = m.prune_last;
    g.pdl = m.pdl;
    g.res16 = m.res16;
    g.launches = captured;
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return set_err(VIT_E_CUDA, "graph instantiate: %s", cudaGetErrorString(ie));
    c.graphs.push_back(g);
    return 0;
}

// After the slot's stream has been synchronised behind a read_flags_async(): what the kernels reported.  Clears the
// device word when set, and the slot's pending marks.  Returns the flags that apply to the work that was pending.
int read_flags_async(DeviceCtx& c) {
    CU_TRY(cudaMemcpyAsync(c.h_flags, c.d_flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
    return 0;
}
int collect_flags(DeviceCtx& c, unsigned int* flags) {
    unsigned int v = *c.h_flags;
    if (v) CU_TRY(cudaMemsetAsync(c.d_flags, 0, sizeof(unsigned int), c.stream));
    *c.h_flags = 0;
    if (!c.pending_fast_softmax) v &= ~VIT_FLAG_ATTN_RANGE;
    if (!c.pending_fp16) v &= ~VIT_FLAG_NONFINITE;
    c.pending_fast_softmax = c.pending_fp16 = false;
    *flags = v & (VIT_FLAG_ATTN_RANGE | VIT_FLAG_NONFINITE);
    return 0;
}

// What to do about the flags a pass (enqueued in mode `used`) came back with.  Both causes produce both symptoms: a softmax
// that overflowed makes logits non-finite, and an FP16 operand that overflowed puts NaN rows into the attention, whose row
// sums then trip the softmax's range check.  So when both flags are up the exact softmax is tried FIRST, on trial; if the
// logits are still non-finite afterwards, it was the FP16 range: the softmax switch is taken back and the BF16 operand set is
// used instead.  Returns 1 if the work must be repeated (mode switched), 0 if the results stand, < 0 on a hard error.
int decide_retry(Engine& e, unsigned int flags, const PassMode& used, bool* exact_on_trial) {
    const bool attn = (flags & VIT_FLAG_ATTN_RANGE) && !used.attn_exact;
    const bool fp16_nonfinite = (flags & VIT_FLAG_NONFINITE) && used.prec == VIT_PREC_FP16;
    if (attn) {
        g_opt.attn_exact = 1;
        *exact_on_trial = fp16_nonfinite;
        return 1;
    }
    if (fp16_nonfinite) {
        if (*exact_on_trial) g_opt.attn_exact = 0;   // the exact softmax did not cure it
        *exact_on_trial = false;
        if (e.policy == VIT_PREC_FP16)
            return set_err(VIT_E_RANGE, "non-finite logits with FP16 operands (overflow): use VIT_PREC_AUTO or VIT_PREC_BF16");
        e.prec = VIT_PREC_BF16;
        return 1;
    }
    *exact_on_trial = false;
    return 0;
}

void read_env_options() {
    g_opt.attn_exact = env_flag("VIT_ATTN_EXACT", 0);
    g_opt.prune_last = env_flag("VIT_PRUNE_LAST", 1);
    g_opt.ln_fused = env_flag("VIT_LN_FUSED", 1);
    g_opt.pdl = env_flag("VIT_PDL", 1);
    g_opt.graphs = env_flag("VIT_GRAPHS", 1);
    g_opt.host_threads = env_flag("VIT_HOST_THREADS", 1);
    g_opt.residual16 = env_flag("VIT_RESIDUAL16", 1);
    g_opt.wave_passes = env_flag("VIT_WAVE_PASSES", 1);
}

int configure_engine(Engine& e, int img_size, int max_batch_per_gpu, int n_gpus, const int* device_ids, int precision) {
    if (img_size <= 0 || img_size % kPatch || img_size > 1024) return set_err(VIT_E_ARG, "img_size %d must be a positive multiple of 16 (<= 1024)", img_size);
    if (max_batch_per_gpu <= 0 || n_gpus <= 0 || n_gpus > 32) return set_err(VIT_E_ARG, "max_batch_per_gpu and n_gpus must be positive (n_gpus <= 32)");
    if (precision != VIT_PREC_BF16 && precision != VIT_PREC_FP16 && precision != VIT_PREC_AUTO) return set_err(VIT_E_ARG, "unknown precision %d", precision);
    for (int g = 0; g < n_gpus; ++g)
        for (int h = 0; h < g; ++h)
            if ((device_ids ? device_ids[g] : g) == (device_ids ? device_ids[h] : h)) return set_err(VIT_E_ARG, "device %d listed twice", device_ids[g]);
    e.img = img_size;
    e.grid = img_size / kPatch;
    e.patches = e.grid * e.grid;
    e.tokens = e.patches + 1;
    e.max_batch = max_batch_per_gpu;
    e.policy = precision;
    e.prec = precision == VIT_PREC_BF16 ? VIT_PREC_BF16 : VIT_PREC_FP16;
    e.resident[VIT_PREC_BF16] = precision != VIT_PREC_FP16;
    e.resident[VIT_PREC_FP16] = precision != VIT_PREC_BF16;
    e.attn_fallbacks = e.prec_fallbacks = 0;
    e.profiling = false;
    read_env_options();
    return 0;
}

void fail_init(int) {
    char keep[sizeof(t_err)];
    memcpy(keep, t_err, sizeof(keep));
    vit_cuda_free();
    memcpy(t_err, keep, sizeof(keep));
}

// ---- weight cache file: header + the arena's bytes
struct CacheHeader {
    char magic[8];            // "VITB200W"
    uint32_t version, img_size, policy, resident_mask;
    uint64_t arena_bytes, checksum;   // FNV-1a 64 over the arena, 8 bytes at a time
};
constexpr uint32_t kCacheVersion = 2;
uint64_t fnv1a64_words(const uint8_t* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x100000001b3ull;
    }
    for (; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

const char* vit_cuda_last_error(void) { return t_err; }
long long vit_cuda_launch_count(void) { return g_launches.load(); }

void vit_cuda_free(void) {
    for (auto& c : g_eng.ctx) destroy_ctx(c);
    g_eng.ctx.clear();
    g_eng.up = false;
}

int vit_cuda_init_ex(const vit_tensor* networks, int n_tensors, int img_size, int max_batch_per_gpu, int n_gpus,
                     const int* device_ids, int precision) {
    if (g_eng.up) vit_cuda_free();
    if (!networks || n_tensors != VIT_NUM_TENSORS) return set_err(VIT_E_ARG, "expected %d weight tensors, got %d", VIT_NUM_TENSORS, n_tensors);
    Engine& e = g_eng;
    VIT_TRY(configure_engine(e, img_size, max_batch_per_gpu, n_gpus, device_ids, precision));
    for (int i = 0; i < n_tensors; ++i)
        if (!networks[i].data || networks[i].size != tensor_numel(i, img_size))
            return set_err(VIT_E_ARG, "weight tensor %d: have %zu floats%s, need %zu for img_size %d", i, networks[i].size,
                           networks[i].data ? "" : " (missing)", tensor_numel(i, img_size), img_size);
    e.ctx.assign(n_gpus, DeviceCtx());
    for (int g = 0; g < n_gpus; ++g) {
        DeviceCtx& c = e.ctx[g];
        int rc = open_ctx(c, device_ids ? device_ids[g] : g, e);
        if (!rc) rc = fill_arena_from_tensors(c, networks, e);
        unsigned int flags = 0;
        if (!rc) rc = take_status_flags(&flags);
        if (!rc && (flags & VIT_FLAG_WEIGHT_RANGE)) {
            // a finite fp32 weight (or ln_w (.) W) does not fit FP16
            if (e.policy == VIT_PREC_FP16) rc = set_err(VIT_E_RANGE, "a weight exceeds the FP16 range (65504): use VIT_PREC_AUTO or VIT_PREC_BF16");
            else e.prec = VIT_PREC_BF16;   // AUTO: run on the BF16 set from the start
        }
        if (!rc) rc = make_weight_maps(c, e);
        if (!rc) rc = alloc_workspace(c, e);
        if (rc) {
            fail_init(rc);
            return rc;
        }
    }
    e.up = true;
    return 0;
}

int vit_cuda_init(const vit_tensor* networks, int n_tensors, int img_size, int max_batch_per_gpu, int n_gpus) {
    return vit_cuda_init_ex(networks, n_tensors, img_size, max_batch_per_gpu, n_gpus, nullptr, VIT_PREC_AUTO);
}

int vit_cuda_save_weight_cache(const char* path) {
    Engine& e = g_eng;
    if (!e.up || !path) return set_err(VIT_E_ARG, "engine not initialised");
    DeviceCtx& c = e.ctx[0];
    CU_TRY(cudaSetDevice(c.device));
    std::vector<uint8_t> host(c.arena_bytes);
    CU_TRY(cudaMemcpy(host.data(), c.arena, c.arena_bytes, cudaMemcpyDeviceToHost));
    CacheHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "VITB200W", 8);
    h.version = kCacheVersion;
    h.img_size = static_cast<uint32_t>(e.img);
    h.policy = static_cast<uint32_t>(e.policy);
    h.resident_mask = (e.resident[0] ? 1u : 0u) | (e.resident[1] ? 2u : 0u);
    h.arena_bytes = c.arena_bytes;
    h.checksum = fnv1a64_words(host.data(), host.size());
    FILE* f = fopen(path, "wb");
    if (!f) return set_err(VIT_E_ARG, "cannot open %s for writing", path);
    const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(host.data(), 1, host.size(), f) == host.size();
    if (fclose(f) != 0 || !ok) return set_err(VIT_E_ARG, "short write to %s", path);
    return 0;
}

int vit_cuda_init_from_cache(const char* path, int max_batch_per_gpu, int n_gpus, const int* device_ids) {
    if (g_eng.up) vit_cuda_free();
    if (!path) return set_err(VIT_E_ARG, "no cache path");
    VIT_TRY(check_device(device_ids ? device_ids[0] : 0, nullptr));
    FILE* f = fopen(path, "rb");
    if (!f) return set_err(VIT_E_ARG, "cannot open weight cache %s", path);
    CacheHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "VITB200W", 8) != 0 || h.version != kCacheVersion) {
        fclose(f);
        return set_err(VIT_E_ARG, "%s is not a weight cache of this engine (version %u)", path, kCacheVersion);
    }
    Engine& e = g_eng;
    int rc = configure_engine(e, static_cast<int>(h.img_size), max_batch_per_gpu, n_gpus, device_ids, static_cast<int>(h.policy));
    if (rc) {
        fclose(f);
        return rc;
    }
    uint8_t* host = nullptr;   // pinned: the host-to-device copies run at full PCIe rate
    if (cudaHostAlloc(reinterpret_cast<void**>(&host), h.arena_bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        fclose(f);
        return set_err(VIT_E_NOMEM, "cannot allocate %llu bytes of pinned memory for the weight cache", (unsigned long long)h.arena_bytes);
    }
    const bool read_ok = fread(host, 1, h.arena_bytes, f) == h.arena_bytes;
    fclose(f);
    if (!read_ok || fnv1a64_words(host, h.arena_bytes) != h.checksum) {
        cudaFreeHost(host);
        return set_err(VIT_E_ARG, "weight cache %s is truncated or corrupt (checksum)", path);
    }
    e.ctx.assign(n_gpus, DeviceCtx());
    for (int g = 0; g < n_gpus && !rc; ++g) {
        DeviceCtx& c = e.ctx[g];
        rc = open_ctx(c, device_ids ? device_ids[g] : g, e);
        if (!rc && (c.arena_bytes != h.arena_bytes || h.resident_mask != ((e.resident[0] ? 1u : 0u) | (e.resident[1] ? 2u : 0u))))
            rc = set_err(VIT_E_ARG, "weight cache %s: arena of %llu bytes, this build lays out %zu", path, (unsigned long long)h.arena_bytes, c.arena_bytes);
        if (!rc && cudaMemcpyAsync(c.arena, host, c.arena_bytes, cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
            rc = set_err(VIT_E_CUDA, "weight cache upload: %s", cudaGetErrorString(cudaGetLastError()));
        if (!rc) rc = make_weight_maps(c, e);
        if (!rc) rc = alloc_workspace(c, e);   // ends with a stream synchronise: the upload is through
    }
    cudaFreeHost(host);
    if (rc) {
        fail_init(rc);
        return rc;
    }
    e.up = true;
    return 0;
}

int vit_cuda_set_option(int option, int value) {
    const int v = value != 0;
    switch (option) {
        case VIT_OPT_ATTENTION_EXACT: g_opt.attn_exact = v; break;
        case VIT_OPT_CLASS_ROW_PRUNING: g_opt.prune_last = v; break;
        case VIT_OPT_LN_FUSED: g_opt.ln_fused = v; break;
        case VIT_OPT_PDL: g_opt.pdl = v; break;
        case VIT_OPT_GRAPHS: g_opt.graphs = v; break;
        case VIT_OPT_HOST_THREADS: g_opt.host_threads = v; break;
        case VIT_OPT_RESIDUAL16: g_opt.residual16 = v; break;
        case VIT_OPT_WAVE_PASSES: g_opt.wave_passes = v; break;
        default: return set_err(VIT_E_ARG, "unknown option %d", option);
    }
    return 0;
}
int vit_cuda_get_option(int option, int* value) {
    if (!value) return set_err(VIT_E_ARG, "bad arguments");
    switch (option) {
        case VIT_OPT_ATTENTION_EXACT: *value = g_opt.attn_exact; break;
        case VIT_OPT_CLASS_ROW_PRUNING: *value = g_opt.prune_last; break;
        case VIT_OPT_LN_FUSED: *value = g_opt.ln_fused; break;
        case VIT_OPT_PDL: *value = g_opt.pdl; break;
        case VIT_OPT_GRAPHS: *value = g_opt.graphs; break;
        case VIT_OPT_HOST_THREADS: *value = g_opt.host_threads; break;
        case VIT_OPT_RESIDUAL16: *value = g_opt.residual16; break;
        case VIT_OPT_WAVE_PASSES: *value = g_opt.wave_passes; break;
        default: return set_err(VIT_E_ARG, "unknown option %d", option);
    }
    return 0;
}
int vit_cuda_set_class_row_pruning(int on) {
    if (!g_eng.up) return set_err(VIT_E_ARG, "engine not initialised");
    return vit_cuda_set_option(VIT_OPT_CLASS_ROW_PRUNING, on);
}
int vit_cuda_set_attention_exact(int on) {
    if (!g_eng.up) return set_err(VIT_E_ARG, "engine not initialised");
    return vit_cuda_set_option(VIT_OPT_ATTENTION_EXACT, on);
}

int vit_cuda_enqueue_device(int gpu_slot, const float* d_images, int n, float* d_logits) {
    Engine& e = g_eng;
    if (!e.up) return set_err(VIT_E_ARG, "engine not initialised");
    if (gpu_slot < 0 || gpu_slot >= (int)e.ctx.size()) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
    if (n <= 0 || n > e.max_batch) return set_err(VIT_E_ARG, "n=%d outside (0, max_batch=%d]", n, e.max_batch);
    if (e.tokens > ATTNL_MAX_TOKENS) return set_err(VIT_E_ARG, "img_size %d (%d tokens): at most %d tokens are supported", e.img, e.tokens, ATTNL_MAX_TOKENS);
    DeviceCtx& c = e.ctx[gpu_slot];
    CU_TRY(cudaSetDevice(c.device));
    return enqueue_forward(c, e, current_mode(e), d_images, n, d_logits);
}

int vit_cuda_sync(int gpu_slot) {
    Engine& e = g_eng;
    if (!e.up || gpu_slot < 0 || gpu_slot >= (int)e.ctx.size()) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
    DeviceCtx& c = e.ctx[gpu_slot];
    CU_TRY(cudaSetDevice(c.device));
    const bool check = c.pending_fast_softmax || c.pending_fp16;
    if (check) VIT_TRY(read_flags_async(c));
    const cudaError_t se = cudaStreamSynchronize(c.stream);
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, "forward");
    if (!check) return 0;
    unsigned int flags = 0;
    VIT_TRY(collect_flags(c, &flags));
    // the mode the flagged work was enqueued in: the engine's current one (a caller that changes options between enqueue
    // and sync gets the conservative reading)
    const PassMode used = current_mode(e);
    const bool exact_before = used.attn_exact;
    const int prec_before = e.prec;
    const int again = decide_retry(e, flags, used, &e.exact_on_trial);
    if (again < 0) return again;
    if (again == 0) return 0;
    if (!exact_before && g_opt.attn_exact.load()) {
        ++e.attn_fallbacks;   // device-resident callers re-enqueue; the engine stays on the exact softmax
        return set_err(VIT_E_RANGE, "attention: a row's scores left the single-pass softmax's exponent window; every pass enqueued on slot %d "
                                    "since its last sync is invalid.  The engine has switched to the exact two-pass softmax: enqueue them again", gpu_slot);
    }
    if (prec_before == VIT_PREC_FP16 && e.prec == VIT_PREC_BF16) {
        if (exact_before && !g_opt.attn_exact.load()) --e.attn_fallbacks;   // the softmax switch was on trial and is taken back
        ++e.prec_fallbacks;
        return set_err(VIT_E_RANGE, "non-finite logits with FP16 operands (overflow): every pass enqueued on slot %d since its last sync is "
                                    "invalid.  The engine has switched to BF16 operands: enqueue them again", gpu_slot);
    }
    return 0;
}

void* vit_cuda_stream(int gpu_slot) {
    if (!g_eng.up || gpu_slot < 0 || gpu_slot >= (int)g_eng.ctx.size()) return nullptr;
    return g_eng.ctx[gpu_slot].stream;
}

int vit_cuda_forward_device(int gpu_slot, const float* d_images, int n, float* d_logits) {
    VIT_TRY(vit_cuda_enqueue_device(gpu_slot, d_images, n, d_logits));
    return vit_cuda_sync(gpu_slot);
}

int vit_cuda_pass_schedule_growth(int n_images, int max_batch, int staged, int growth_percent, int* first, int* count, int cap) {
    if (n_images < 0 || max_batch <= 0 || !first || !count || cap <= 0) return set_err(VIT_E_ARG, "bad schedule arguments");
    const int growth = std::min(300, std::max(100, growth_percent));
    int n = 0;
    // pinned input: the copy of a pass is `growth` times faster than its kernels -> every pass may be that much larger than
    // the one whose kernels hide its copy (300 %: a lone GPU on PCIe Gen5; less when several GPUs share the host's memory
    // and root complexes).  Staged input (pageable or one allocation per image): the host-side gather runs at about the rate
    // the GPU consumes images, so the passes after the first stay at 128 images -- each gather hides under the kernels of the
    // pass before it.
    for (int done = 0, sz = staged ? 64 : 32; done < n_images;
         sz = staged ? std::min(max_batch, 128) : std::min(max_batch, std::max(sz + 1, static_cast<int>(static_cast<long long>(sz) * growth / 100)))) {
        const int nb = std::min(std::min(sz, max_batch), n_images - done);
        if (n == cap) return set_err(VIT_E_ARG, "pass schedule of %d images with max_batch %d needs more than %d passes", n_images, max_batch, cap);
        first[n] = done;
        count[n] = nb;
        done += nb;
        ++n;
    }
    return n;
}

int vit_cuda_pass_schedule_model(int n_images, int max_batch, double copy_us_per_image, double kernel_us_per_image, double fixed_us_per_pass,
                                 int* first, int* count, int cap) {
    if (n_images < 0 || max_batch <= 0 || !first || !count || cap <= 0 || !(copy_us_per_image > 0) || !(kernel_us_per_image > 0) || !(fixed_us_per_pass >= 0))
        return set_err(VIT_E_ARG, "bad schedule arguments");
    if (n_images == 0) return 0;
    // the largest pass whose copy hides under a pass of `prev` images (10 % reserve), never smaller than that pass
    auto next_size = [&](int prev) {
        const double fit = 0.9 * (fixed_us_per_pass + kernel_us_per_image * prev) / copy_us_per_image;
        return static_cast<int>(std::min<double>(max_batch, std::max<double>(prev, fit)));
    };
    auto covered = [&](int n0, int passes) {
        long long sum = 0;
        for (int i = 0, sz = n0; i < passes && sum < n_images; ++i, sz = next_size(sz)) sum += sz;
        return sum;
    };
    int best_p = 0, best_n0 = 0;
    double best_cost = 0;
    const int n0_max = std::min(n_images, max_batch);
    for (int p = 1; p <= cap; ++p) {
        if (covered(n0_max, p) < n_images) continue;
        int lo = 1, hi = n0_max;   // the smallest first pass with which p passes reach n_images
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (covered(mid, p) >= n_images) hi = mid;
            else lo = mid + 1;
        }
        const double cost = copy_us_per_image * lo + fixed_us_per_pass * p;   // what the pipeline does not hide
        if (best_p == 0 || cost < best_cost) {
            best_p = p;
            best_n0 = lo;
            best_cost = cost;
        }
        if (lo == 1) break;   // more passes cannot make the first one smaller
    }
    if (best_p == 0) return set_err(VIT_E_ARG, "pass schedule of %d images with max_batch %d needs more than %d passes", n_images, max_batch, cap);
    int n = 0;
    for (int done = 0, sz = best_n0; done < n_images; sz = next_size(sz), ++n) {
        first[n] = done;
        count[n] = std::min(sz, n_images - done);
        done += count[n];
    }
    return n;
}

int vit_cuda_pass_schedule_waves(int n_images, int max_batch, int staged, int growth_percent, int tokens, int sm_count, int* first, int* count, int cap) {
    if (tokens <= 0 || sm_count < 2) return set_err(VIT_E_ARG, "bad schedule arguments");
    const int worst = std::max(cap, n_images / std::min(std::max(max_batch, 1), 32) + 40);
    std::vector<int> f(worst), c(worst);
    int n = vit_cuda_pass_schedule_growth(n_images, max_batch, staged, growth_percent, f.data(), c.data(), worst);
    if (n < 0) return n;
    n = snap_schedule(f, c, n, n_images, max_batch, tokens, sm_count / 2);
    if (n > cap) return set_err(VIT_E_ARG, "pass schedule of %d images with max_batch %d needs more than %d passes", n_images, max_batch, cap);
    std::copy(f.begin(), f.begin() + n, first);
    std::copy(c.begin(), c.begin() + n, count);
    return n;
}

int vit_cuda_pass_schedule_ex(int n_images, int max_batch, int staged, int* first, int* count, int cap) {
    return vit_cuda_pass_schedule_growth(n_images, max_batch, staged, 300, first, count, cap);
}

int vit_cuda_pass_schedule(int n_images, int max_batch, int* first, int* count, int cap) {
    return vit_cuda_pass_schedule_ex(n_images, max_batch, 0, first, count, cap);
}

int vit_cuda_shard_range(int n, int n_gpus, int g, int* lo, int* hi) {
    if (n < 0 || n_gpus <= 0 || g < 0 || g >= n_gpus || !lo || !hi) return set_err(VIT_E_ARG, "bad shard arguments");
    const int per_gpu = (n + n_gpus - 1) / n_gpus;
    *lo = std::min(n, g * per_gpu);
    *hi = std::min(n, *lo + per_gpu);
    return 0;
}

}  // extern "C"

namespace {

// hand the logits parked in pinned staging buffer `buf` (if any) over to the caller's (pageable) array
int drain_logits(DeviceCtx& c, int buf, float* logits_out) {
    if (c.pend_count[buf] == 0) return 0;
    CU_TRY(cudaEventSynchronize(c.ev_logits[buf]));
    memcpy(logits_out + static_cast<size_t>(c.pend_first[buf]) * kClasses, c.h_logits[buf], static_cast<size_t>(c.pend_count[buf]) * kClasses * sizeof(float));
    c.pend_count[buf] = 0;
    return 0;
}

// What one host call hands to every slot's feeding thread
struct HostJob {
    const float* images_nchw;          // contiguous images, or
    const float* const* image_ptrs;    // one pointer per image (the reference loader's form)
    int n;
    float* logits_out;
    bool images_pinned, logits_pinned;
    PassMode mode;
    int gather_threads;                // host threads a slot may use to gather a staged pass
    int per_gpu;                       // images of a full shard
};

// Growth factor (percent) of the pass schedule for pinned input on this slot: how much faster a pass's H2D copy is than its
// kernels, from the previous calls' measurements, with a 15 % reserve; 300 (a lone GPU on PCIe Gen5) until measured.
int schedule_growth(const DeviceCtx& c) {
    if (c.h2d_ms_per_image <= 0 || c.kernel_ms_per_image <= 0) return 300;
    return std::min(300, std::max(100, static_cast<int>(85.0 * c.kernel_ms_per_image / c.h2d_ms_per_image)));
}

// One slot's shard of a host call: its passes, H2D copies on the copy stream into alternating image buffers under the
// kernels of the previous pass, logits back per pass; ends with the slot's stream synchronised and its status flags read.
// Wave-aware pass sizes.  The GEMMs are persistent over sm_count / 2 CTA pairs with 256-row tiles, so a pass costs whole waves:
// 32 images are 25 row tiles = 75 out_proj / mlp_3 tiles = TWO waves on 74 pairs, 31 images are 24 row tiles = one; 128 images
// need 5 / 13 / 17 waves (N = 768 / 2304 / 3072) where 127 need 4 / 12 / 16.  Cost of a pass in units of one K = 768 tile wave:
// out_proj + mlp_3 (K = 768 + 3072, 3 column tiles), in_proj (9), mlp_0 (12).
long long pass_wave_cost(int nb, int tokens, int pairs) {
    const long long tm = (static_cast<long long>(nb) * tokens + 255) / 256;
    auto waves = [&](int tiles_n) { return (tm * tiles_n + pairs - 1) / pairs; };
    return 5 * waves(3) + waves(9) + waves(12);
}
// The most wave-efficient size in [0.8, 1.0] x target ([0.8, 1.08] below 128 images; not above limit); ties go to the larger pass.
int snap_pass_count(int target, int limit, int tokens, int pairs) {
    if (target >= limit) return limit;   // the shard's last pass takes what is left
    int best = target;
    double best_eff = static_cast<double>(target) / pass_wave_cost(target, tokens, pairs);
    // (upwards only for small passes, where 8 % are a few images: a larger pass's copy must still hide under the pass before it)
    for (int nb = std::max(1, target * 4 / 5); nb <= std::min(limit, target < 128 ? target + target / 12 : target); ++nb) {
        const double eff = static_cast<double>(nb) / pass_wave_cost(nb, tokens, pairs);
        if (eff > best_eff * (1 + 1e-9) || (eff > best_eff * (1 - 1e-9) && nb > best)) {
            best = nb;
            best_eff = eff;
        }
    }
    return best;
}
// Re-cut a shard's schedule (vit_cuda_pass_schedule_growth: sizes from the copy / kernel rates) at wave-efficient sizes.
int snap_schedule(std::vector<int>& first, std::vector<int>& count, int n_pass, int n_images, int max_batch, int tokens, int pairs) {
    std::vector<int> target(count.begin(), count.begin() + n_pass);
    int n = 0;
    for (int done = 0; done < n_images; ++n) {
        if (n == static_cast<int>(first.size())) {
            first.push_back(0);
            count.push_back(0);
        }
        const int left = n_images - done;
        const int want = n < n_pass ? target[n] : std::min(max_batch, left);
        first[n] = done;
        count[n] = snap_pass_count(std::min(want, left), std::min(max_batch, left), tokens, pairs);
        const int rest = left - count[n];   // no stub of a pass at the end: a pass has a fixed cost of ~60 launches
        if (rest > 0 && rest < std::max(16, count[n] / 4) && left <= max_batch) count[n] = left;
        done += count[n];
    }
    return n;
}

// Runs on the slot's own host thread when the engine has several GPUs: a gather of pass i + 1 for one GPU
// must not hold up the enqueueing for another.
int run_shard(Engine& e, int g, const HostJob& job, unsigned int* flags_out) {
    DeviceCtx& c = e.ctx[g];
    const int G = static_cast<int>(e.ctx.size());
    const size_t img_elems = static_cast<size_t>(3) * e.img * e.img;
    *flags_out = 0;
    CU_TRY(cudaSetDevice(c.device));
    c.pend_count[0] = c.pend_count[1] = 0;   // nothing parked from an earlier call that failed half-way
    int lo, hi;
    vit_cuda_shard_range(job.n, G, g, &lo, &hi);
    const bool staged = job.image_ptrs || !job.images_pinned;
    // Pass schedule of the shard: the H2D copy of pass i+1 (copy stream, second image buffer) hides under the kernels of
    // pass i, so only the FIRST pass's copy is exposed -- it is kept small (32 images, 19 MB) -- and every later pass may be
    // `growth` times the previous one, up to the workspace size.  A lone GPU on PCIe Gen5 copies images ~3.4x faster than the
    // kernels consume them (1024 images: 32 + 96 + 288 + 608); eight GPUs pulling from one host at once do not (measured
    // ~25 GB/s each: the 608-image pass then waits 10 ms for its copy), so the factor follows the measured rates.
    const int worst = job.per_gpu / std::min(e.max_batch, 32) + 40;
    std::vector<int> pass_first(worst), pass_count(worst);
    const bool tuned = g_opt.wave_passes.load() != 0;
    int n_pass;
    if (tuned && !staged) {
        // pinned input: sizes from the cost model of this GPU's pipeline (measured in the previous calls; before that, what a
        // lone B200 on PCIe Gen5 shows: 55 GB/s of copies, 34 us of kernels per 197-token image, 0.7 ms per pass)
        const double img_bytes = static_cast<double>(img_elems) * sizeof(float);
        const double copy_us = c.h2d_ms_per_image > 0 ? 1e3 * c.h2d_ms_per_image : img_bytes / 55e3;
        const double kern_us = c.pass_ms_per_image > 0 ? 1e3 * c.pass_ms_per_image : 34.3 * std::pow(e.tokens / 197.0, 1.1);
        const double fixed_us = c.pass_ms_per_image > 0 ? std::max(200.0, 1e3 * c.pass_fixed_ms) : 700.0;   // (a fit that lost the fixed cost
                                                                                                             // would ask for a string of tiny passes)
        n_pass = vit_cuda_pass_schedule_model(std::min(job.per_gpu, hi - lo), e.max_batch, copy_us, kern_us, fixed_us, pass_first.data(), pass_count.data(), worst);
    } else {
        n_pass = vit_cuda_pass_schedule_growth(job.per_gpu, e.max_batch, staged ? 1 : 0, schedule_growth(c), pass_first.data(), pass_count.data(), worst);
    }
    if (n_pass < 0) return n_pass;
    if (tuned) n_pass = snap_schedule(pass_first, pass_count, n_pass, std::min(job.per_gpu, hi - lo), e.max_batch, e.tokens, c.sm_count / 2);
    const bool timing = !staged && !e.profiling;
    if (timing)
        while (c.tm_events.size() < static_cast<size_t>(4 * n_pass)) {
            cudaEvent_t ev;
            CU_TRY(cudaEventCreate(&ev));
            c.tm_events.push_back(ev);
        }
    int n_timed = 0;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int first = lo + pass_first[pass];
        if (first >= hi) break;
        const int nb = std::min(pass_count[pass], hi - first);
        const int buf = pass & 1;
        // H2D of this pass overlaps the previous pass's compute (other image buffer)
        if (pass >= 2) CU_TRY(cudaStreamWaitEvent(c.copy_stream, c.ev_done[buf], 0));
        const float* src = nullptr;
        if (staged) {
            // separately allocated images (the reference's loader, Network.c:75-93) or pageable memory: gather this
            // pass into the slot's pinned staging buffer while the GPU works on the previous pass
            if (!c.h_stage[buf]) {
                CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c.h_stage[buf]), static_cast<size_t>(e.max_batch) * img_elems * sizeof(float), cudaHostAllocPortable));
                CU_TRY(cudaEventCreateWithFlags(&c.ev_stage[buf], cudaEventDisableTiming));
            } else {
                CU_TRY(cudaEventSynchronize(c.ev_stage[buf]));   // its previous copy has left the buffer
            }
            // One host thread copies ~10 GB/s: 1024 images (617 MB) would take longer than the GPU needs for them.
            float* dst = c.h_stage[buf];
            const float* const* ptrs = job.image_ptrs;
            const float* flat = job.images_nchw;
            auto gather = [=](int i0, int i1) {
                for (int i = i0; i < i1; ++i)
                    memcpy(dst + static_cast<size_t>(i) * img_elems, ptrs ? ptrs[first + i] : flat + static_cast<size_t>(first + i) * img_elems, img_elems * sizeof(float));
            };
            const int n_thr = std::min(job.gather_threads, (nb + 15) / 16);
            if (n_thr <= 1) {
                gather(0, nb);
            } else {
                std::vector<std::thread> pool;
                for (int t = 1; t < n_thr; ++t) pool.emplace_back(gather, nb * t / n_thr, nb * (t + 1) / n_thr);
                gather(0, nb / n_thr);
                for (auto& th : pool) th.join();
            }
            src = dst;
        } else {
            src = job.images_nchw + static_cast<size_t>(first) * img_elems;
        }
        if (timing) CU_TRY(cudaEventRecord(c.tm_events[4 * pass], c.copy_stream));
        CU_TRY(cudaMemcpyAsync(c.images[buf], src, static_cast<size_t>(nb) * img_elems * sizeof(float), cudaMemcpyHostToDevice, c.copy_stream));
        if (timing) CU_TRY(cudaEventRecord(c.tm_events[4 * pass + 1], c.copy_stream));
        if (staged) CU_TRY(cudaEventRecord(c.ev_stage[buf], c.copy_stream));
        CU_TRY(cudaEventRecord(c.ev_h2d[buf], c.copy_stream));
        CU_TRY(cudaStreamWaitEvent(c.stream, c.ev_h2d[buf], 0));
        if (timing) CU_TRY(cudaEventRecord(c.tm_events[4 * pass + 2], c.stream));
        c.img_release = c.ev_done[buf];
        const int frc = enqueue_forward(c, e, job.mode, c.images[buf], nb, c.logits);
        if (c.img_release) {   // not taken (graph replay of a small pass, or an error): the buffer is free when the pass is through
            c.img_release = nullptr;
            if (!frc) CU_TRY(cudaEventRecord(c.ev_done[buf], c.stream));
        }
        VIT_TRY(frc);
        if (timing) CU_TRY(cudaEventRecord(c.tm_events[4 * pass + 3], c.stream));
        n_timed = pass + 1;
        if (job.logits_pinned) {
            CU_TRY(cudaMemcpyAsync(job.logits_out + static_cast<size_t>(first) * kClasses, c.logits,
                                   static_cast<size_t>(nb) * kClasses * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        } else {
            // a device-to-pageable copy would block the host until this pass is through, and with it the
            // enqueueing (and gathering) of the next one: go through pinned staging, hand over later
            if (!c.h_logits[buf]) {
                CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c.h_logits[buf]), static_cast<size_t>(e.max_batch) * kClasses * sizeof(float), cudaHostAllocPortable));
                CU_TRY(cudaEventCreateWithFlags(&c.ev_logits[buf], cudaEventDisableTiming));
            }
            VIT_TRY(drain_logits(c, buf, job.logits_out));
            CU_TRY(cudaMemcpyAsync(c.h_logits[buf], c.logits, static_cast<size_t>(nb) * kClasses * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
            CU_TRY(cudaEventRecord(c.ev_logits[buf], c.stream));
            c.pend_first[buf] = first;
            c.pend_count[buf] = nb;
        }
    }
    VIT_TRY(read_flags_async(c));
    const cudaError_t se = cudaStreamSynchronize(c.stream);
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, "forward");
    for (int buf = 0; buf < 2; ++buf) VIT_TRY(drain_logits(c, buf, job.logits_out));
    if (timing) {
        // rates of this call's passes of >= 64 images (small passes are latency, not rate), folded into the running averages
        double h2d_ms = 0, k_ms = 0;
        long long imgs = 0;
        for (int pass = 0; pass < n_timed; ++pass) {
            const int nb = std::min(pass_count[pass], hi - lo - pass_first[pass]);
            if (nb < 64) continue;
            float a = 0, b = 0;
            if (cudaEventElapsedTime(&a, c.tm_events[4 * pass], c.tm_events[4 * pass + 1]) != cudaSuccess ||
                cudaEventElapsedTime(&b, c.tm_events[4 * pass + 2], c.tm_events[4 * pass + 3]) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            h2d_ms += a;
            k_ms += b;
            imgs += nb;
        }
        static const int pass_trace = env_flag("VIT_PASS_TRACE", 0);   // diagnosis: the call's timeline on stderr
        if (pass_trace) {
            for (int pass = 0; pass < n_timed; ++pass) {
                float t[4] = {0, 0, 0, 0};
                for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], c.tm_events[0], c.tm_events[4 * pass + i]);
                fprintf(stderr, "pass_trace gpu %d pass %d images %d: copy %.3f..%.3f ms, kernels %.3f..%.3f ms\n", g, pass, pass_count[pass], t[0], t[1], t[2], t[3]);
            }
            cudaGetLastError();
        }
        // kernels of a pass = fixed + per-image * images: least squares of the RELATIVE error over this call's passes (weights
        // 1 / images^2: the per-image rate falls a little with the size of a pass, and it is the small passes whose duration
        // the schedule has to predict -- the next pass's copy must hide under them); needs two different sizes
        {
            double sw = 0, sn = 0, st = 0, snn = 0, snt = 0;
            int m = 0;
            for (int pass = 0; pass < n_timed; ++pass) {
                float b = 0;
                if (cudaEventElapsedTime(&b, c.tm_events[4 * pass + 2], c.tm_events[4 * pass + 3]) != cudaSuccess) {
                    cudaGetLastError();
                    continue;
                }
                const double nb = pass_count[pass], w = 1.0 / (nb * nb);
                sw += w, sn += w * nb, st += w * b, snn += w * nb * nb, snt += w * nb * b;
                ++m;
            }
            const double det = sw * snn - sn * sn;
            if (m >= 2 && det > 1e-12 * sw * snn) {
                const double slope = (sw * snt - sn * st) / det, icpt = (st - slope * sn) / sw;
                if (slope > 0 && icpt >= 0 && icpt < 5.0) {
                    const double w = c.pass_ms_per_image > 0 ? 0.5 : 1.0;
                    c.pass_ms_per_image = (1 - w) * c.pass_ms_per_image + w * slope;
                    c.pass_fixed_ms = (1 - w) * c.pass_fixed_ms + w * icpt;
                }
            }
        }
        if (imgs > 0 && h2d_ms > 0 && k_ms > 0) {
            const double w = c.h2d_ms_per_image > 0 ? 0.5 : 1.0;
            c.h2d_ms_per_image = (1 - w) * c.h2d_ms_per_image + w * h2d_ms / imgs;
            c.kernel_ms_per_image = (1 - w) * c.kernel_ms_per_image + w * k_ms / imgs;
        }
    }
    return collect_flags(c, flags_out);
}

// After a failed shard: let everything this call put on the slot's streams finish, so that the next call starts clean.
void quiesce(DeviceCtx& c) {
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.copy_stream);
    cudaStreamSynchronize(c.stream);
    c.pend_count[0] = c.pend_count[1] = 0;
    cudaGetLastError();
}

int forward_host_once(const float* images_nchw, const float* const* image_ptrs, int n, float* logits_out, unsigned int* flags) {
    Engine& e = g_eng;
    *flags = 0;
    if (e.tokens > ATTNL_MAX_TOKENS) return set_err(VIT_E_ARG, "img_size %d (%d tokens): at most %d tokens are supported", e.img, e.tokens, ATTNL_MAX_TOKENS);
    const int G = static_cast<int>(e.ctx.size());
    const int per_gpu = (n + G - 1) / G;  // contiguous shards
    HostJob job;
    job.images_nchw = images_nchw;
    job.image_ptrs = image_ptrs;
    job.n = n;
    job.logits_out = logits_out;
    cudaPointerAttributes pa;
    job.logits_pinned = cudaPointerGetAttributes(&pa, logits_out) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    job.images_pinned = images_nchw && cudaPointerGetAttributes(&pa, images_nchw) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    job.mode = current_mode(e);
    job.per_gpu = per_gpu;
    // gathering threads: up to eight per slot; with several slots all of the host's cores (the slots' feeding threads sit in
    // the gather themselves), with one slot half of them
    const int hw = std::max(2, static_cast<int>(std::thread::hardware_concurrency()));
    job.gather_threads = std::max(1, std::min(8, G > 1 ? hw / G : hw / 2));

    std::vector<int> rc(G, 0);
    std::vector<unsigned int> fl(G, 0);
    auto work = [&](int g) {
        rc[g] = run_shard(e, g, job, &fl[g]);
        if (rc[g]) {
            memcpy(e.ctx[g].err, t_err, sizeof(t_err));
            quiesce(e.ctx[g]);
        }
    };
    if (G == 1 || !g_opt.host_threads.load()) {
        // single GPU, or VIT_OPT_HOST_THREADS = 0: this thread serves the slots one after the other
        for (int g = 0; g < G; ++g) work(g);
    } else {
        std::vector<std::thread> feeders;
        for (int g = 1; g < G; ++g) feeders.emplace_back(work, g);
        work(0);
        for (auto& th : feeders) th.join();
    }
    for (int g = 0; g < G; ++g) {
        if (rc[g]) {
            memcpy(t_err, e.ctx[g].err, sizeof(t_err));
            return rc[g];
        }
        *flags |= fl[g];
    }
    return 0;
}

int forward_host(const float* images_nchw, const float* const* image_ptrs, int n, float* logits_out, int* top1_out) {
    Engine& e = g_eng;
    if (!e.up) return set_err(VIT_E_ARG, "engine not initialised");
    if (!logits_out || n < 0) return set_err(VIT_E_ARG, "bad arguments");
    if (n == 0) return 0;
    // A flagged call is repeated: with the exact two-pass softmax (a row left the single-pass softmax's exponent window)
    // and / or, under VIT_PREC_AUTO, with the BF16 operand set (an FP16 operand overflowed -- 8-bit exponent instead of 5);
    // decide_retry picks the order.  A switch that was needed becomes permanent once it has been needed in three calls (the
    // data evidently does it regularly); until then the next call starts from the fast configuration again.
    const bool exact0 = g_opt.attn_exact.load() != 0;
    const int prec0 = e.prec;
    bool trial = false;
    for (int attempt = 0;; ++attempt) {
        unsigned int flags = 0;
        const PassMode used = current_mode(e);
        VIT_TRY(forward_host_once(images_nchw, image_ptrs, n, logits_out, &flags));
        const int again = decide_retry(e, flags, used, &trial);
        if (again < 0) {
            g_opt.attn_exact = exact0 ? 1 : 0;
            return again;
        }
        if (again == 0 || attempt == 3) break;
    }
    if (!exact0 && g_opt.attn_exact.load()) {
        if (++e.attn_fallbacks < 3) g_opt.attn_exact = 0;
    }
    if (prec0 == VIT_PREC_FP16 && e.prec == VIT_PREC_BF16) {
        if (++e.prec_fallbacks < 3) e.prec = VIT_PREC_FP16;
    }
    if (top1_out)
        for (int i = 0; i < n; ++i) {
            const float* row = logits_out + static_cast<size_t>(i) * kClasses;
            int best = 0;
            for (int j = 1; j < kClasses; ++j)
                if (row[j] > row[best]) best = j;
            top1_out[i] = best;
        }
    return 0;
}

}  // namespace

extern "C" {

int vit_cuda_forward(const float* images_nchw, int n, float* logits_out, int* top1_out) {
    if (n == 0 && g_eng.up) return 0;   // nothing to do, whatever the pointers are
    if (!images_nchw) return set_err(VIT_E_ARG, "bad arguments");
    return forward_host(images_nchw, nullptr, n, logits_out, top1_out);
}

int vit_cuda_forward_scattered(const float* const* images, int n, float* logits_out, int* top1_out) {
    if (n == 0 && g_eng.up) return 0;
    if (!images) return set_err(VIT_E_ARG, "bad arguments");
    for (int i = 0; i < n; ++i)
        if (!images[i]) return set_err(VIT_E_ARG, "image %d is NULL", i);
    return forward_host(nullptr, images, n, logits_out, top1_out);
}

int vit_cuda_info(long long* out, int n) {
    if (!g_eng.up || !out) return set_err(VIT_E_ARG, "engine not initialised");
    const DeviceCtx& c = g_eng.ctx[0];
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, c.device));
    const double img_mb = 3.0 * g_eng.img * g_eng.img * 4 / 1e6;
    const long long v[18] = {c.sm_count, prop.major, prop.minor, g_eng.max_batch, g_eng.tokens, g_eng.prec,
                             (long long)g_eng.ctx.size(), (long long)(c.ws_bytes >> 20), g_opt.attn_exact.load() ? 1 : 0,
                             g_eng.attn_fallbacks, g_opt.prune_last.load() ? 1 : 0, g_eng.policy, g_eng.prec_fallbacks,
                             (long long)(c.arena_bytes >> 20), schedule_growth(c),
                             c.h2d_ms_per_image > 0 ? (long long)(img_mb / c.h2d_ms_per_image * 1e3) : 0,
                             (long long)(1e3 * c.pass_fixed_ms + 0.5), (long long)(1e6 * c.pass_ms_per_image + 0.5)};
    for (int i = 0; i < n && i < 18; ++i) out[i] = v[i];
    return 0;
}

// ------------------------------------------------------------------------------------ timing helpers
int vit_cuda_timer_start(int gpu_slot) {
    if (!g_eng.up || gpu_slot < 0 || gpu_slot >= (int)g_eng.ctx.size()) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
    DeviceCtx& c = g_eng.ctx[gpu_slot];
    CU_TRY(cudaSetDevice(c.device));
    if (!c.timer.start) {
        CU_TRY(cudaEventCreate(&c.timer.start));
        CU_TRY(cudaEventCreate(&c.timer.stop));
    }
    CU_TRY(cudaEventRecord(c.timer.start, c.stream));
    return 0;
}
int vit_cuda_timer_stop(int gpu_slot, float* ms) {
    if (!g_eng.up || gpu_slot < 0 || gpu_slot >= (int)g_eng.ctx.size() || !ms) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
    DeviceCtx& c = g_eng.ctx[gpu_slot];
    if (!c.timer.start) return set_err(VIT_E_ARG, "timer not started");
    CU_TRY(cudaSetDevice(c.device));
    CU_TRY(cudaEventRecord(c.timer.stop, c.stream));
    const cudaError_t se = cudaEventSynchronize(c.timer.stop);
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, "timer_stop");
    CU_TRY(cudaEventElapsedTime(ms, c.timer.start, c.timer.stop));
    return 0;
}
int vit_cuda_profile_enable(int on) {
    if (!g_eng.up) return set_err(VIT_E_ARG, "engine not initialised");
    g_eng.profiling = on != 0;
    for (auto& c : g_eng.ctx)
        for (auto& u : c.prof_used) u = 0;
    return 0;
}
int vit_cuda_profile_read(int gpu_slot, double* total_ms, long long* launches, int ncat) {
    if (!g_eng.up || gpu_slot < 0 || gpu_slot >= (int)g_eng.ctx.size()) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
    DeviceCtx& c = g_eng.ctx[gpu_slot];
    CU_TRY(cudaSetDevice(c.device));
    const cudaError_t se = cudaStreamSynchronize(c.stream);
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, "profile_read");
    for (int k = 0; k < ncat && k < VIT_PROF_NCAT; ++k) {
        double acc = 0;
        for (size_t i = 0; i < c.prof_used[k]; ++i) {
            float ms = 0;
            CU_TRY(cudaEventElapsedTime(&ms, c.prof_events[k][i].first, c.prof_events[k][i].second));
            acc += ms;
        }
        total_ms[k] = acc;
        launches[k] = (long long)c.prof_used[k];
        c.prof_used[k] = 0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------ memory helpers
static int slot_device(int gpu_slot, int* dev) {
    if (g_eng.up) {
        if (gpu_slot < 0 || gpu_slot >= (int)g_eng.ctx.size()) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
        *dev = g_eng.ctx[gpu_slot].device;
    } else {
        *dev = gpu_slot;
        VIT_TRY(check_device(gpu_slot, nullptr));
    }
    return 0;
}
int vit_cuda_dev_alloc(int gpu_slot, size_t bytes, void** d_ptr) {
    int dev = 0;
    VIT_TRY(slot_device(gpu_slot, &dev));
    CU_TRY(cudaSetDevice(dev));
    CU_TRY(cudaMalloc(d_ptr, bytes));
    return 0;
}
int vit_cuda_dev_free(int gpu_slot, void* d_ptr) {
    int dev = 0;
    VIT_TRY(slot_device(gpu_slot, &dev));
    CU_TRY(cudaSetDevice(dev));
    CU_TRY(cudaFree(d_ptr));
    return 0;
}
int vit_cuda_dev_upload(int gpu_slot, void* d_dst, const void* h_src, size_t bytes) {
    int dev = 0;
    VIT_TRY(slot_device(gpu_slot, &dev));
    CU_TRY(cudaSetDevice(dev));
    CU_TRY(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int vit_cuda_dev_download(int gpu_slot, void* h_dst, const void* d_src, size_t bytes) {
    int dev = 0;
    VIT_TRY(slot_device(gpu_slot, &dev));
    CU_TRY(cudaSetDevice(dev));
    CU_TRY(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int vit_cuda_host_alloc_pinned(size_t bytes, void** h_ptr) {
    VIT_TRY(check_device(0, nullptr));
    CU_TRY(cudaHostAlloc(h_ptr, bytes, cudaHostAllocPortable));
    return 0;
}
int vit_cuda_host_free_pinned(void* h_ptr) {
    CU_TRY(cudaFreeHost(h_ptr));
    return 0;
}

#include "op_entry.inc"

}  // extern "C"
