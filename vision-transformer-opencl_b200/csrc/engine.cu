// engine.cu -- the C ABI of include/vit_cuda.h: weight upload, workspaces, TMA descriptors,
// the ViT-B/16 forward schedule (ViT_seq.c:337-439 / ViT_opencl.c:785-883 re-expressed as a
// batch of token-flattened kernels) and the single-operator test entry points.
//
// There is deliberately no CPU fallback: every entry point fails with VIT_E_NODEVICE unless a
// compute-capability-10.x device is present.
#include "../../include/vit_cuda.h"

#include <algorithm>
#include <atomic>
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
PFN_encodeTiled g_encode = nullptr;

int load_driver() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return set_err(VIT_E_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

// 2-D row-major [rows][cols] tensor, box {box_cols, box_rows}, 128B swizzle (box_cols * elem = 128 B).
int make_tmap_raw(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t cols, uint64_t rows,
                  uint32_t box_cols, uint32_t box_rows) {
    VIT_TRY(load_driver());
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * elem_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_err(VIT_E_CUDA, "cuTensorMapEncodeTiled failed (%d) cols=%llu rows=%llu box=%ux%u", (int)r,
                       (unsigned long long)cols, (unsigned long long)rows, box_cols, box_rows);
    return 0;
}
// operand-precision (16-bit) tensor
int make_tmap(CUtensorMap* m, int prec, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
              uint32_t box_rows) {
    return make_tmap_raw(m, prec == VIT_PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base,
                         cols, rows, box_cols, box_rows);
}
// fp32 tensor (residual stream): 32 columns = 128 bytes per box row
int make_tmap_f32(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    return make_tmap_raw(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, cols, rows, 32, box_rows);
}

// 3-D [d2][d1][d0] tensor of 16-bit elements, dense, box {64, box_d1, 1}, 128B swizzle: per-image views
// whose rows past d1 are clipped on store / zero-filled on load.
int make_tmap_3d(CUtensorMap* m, int prec, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box_d1) {
    VIT_TRY(load_driver());
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    const cuuint32_t box[3] = {64, box_d1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(m, prec == VIT_PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                                const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(VIT_E_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d) dims=%llu,%llu,%llu", (int)r,
                                          (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
    return 0;
}

// 3-D [d2][d1][d0] fp32 tensor, box {32, box_d1, 1} (128 bytes per box row), 128B swizzle
int make_tmap_3d_f32(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box_d1) {
    VIT_TRY(load_driver());
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
    const cuuint32_t box[3] = {32, box_d1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(VIT_E_CUDA, "cuTensorMapEncodeTiled(3d f32) failed (%d) dims=%llu,%llu,%llu", (int)r,
                                          (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
    return 0;
}

// ------------------------------------------------------------------------------------ launches
constexpr int kGemmBN = 256, kGemmStages = 4, kGemmEpiWG = 2;
constexpr int kGemmThreads = (GEMM_NON_EPI_WARPS + 4 * kGemmEpiWG) * 32;

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(VIT_E_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// Launch with programmatic stream serialization (the kernel must call griddep_wait() before it touches anything
// its predecessor wrote).  VIT_PDL=0: ordinary launches.
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("VIT_PDL");
        v = (s && atoi(s) == 0) ? 0 : 1;
    }
    return v == 1;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

int attn_no_pingpong() {  // VIT_ATTN_NO_PINGPONG=1: tuning switch of the attention kernel (A/B testing)
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("VIT_ATTN_NO_PINGPONG");
        v = (s && atoi(s) != 0) ? 1 : 0;
    }
    return v;
}

constexpr int kPairStages = 6;
int gemm_impl() {  // VIT_GEMM_IMPL=1 selects the single-CTA 128x256 kernel (A/B testing)
    static int impl = -1;
    if (impl < 0) {
        const char* s = getenv("VIT_GEMM_IMPL");
        impl = (s && atoi(s) == 1) ? 1 : 2;
    }
    return impl;
}

template <typename T, int EPI>
int launch_gemm_pair_t(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int sm_count, cudaStream_t st) {
    using L = GemmPairSmem<kPairStages>;
    auto kern = gemm_sm100_pair_kernel<T, kPairStages, kGemmEpiWG, EPI>;
    static int configured_dev_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured_dev_mask & (1 << dev))) {
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
        configured_dev_mask |= 1 << dev;
    }
    const int tiles = ((p.M + 255) / 256) * (p.N / 256);
    const int grid = 2 * std::min(tiles, sm_count / 2);
    kern<<<grid, kGemmThreads, L::DYN_BYTES, st>>>(ta, tb, p);
    return check_launch("gemm_pair");
}

constexpr int kStagedStages = 5, kStagedSlots = 4;
// LN = true: LayerNorm folded into the GEMM (consumer for EPI_BIAS / EPI_BIAS_GELU, producer for
// EPI_BIAS_RESIDUAL, see gemm_sm100.cuh).  The producer gives one operand stage up for the two
// staging tiles of the operand-precision copy.
template <typename T, int EPI, bool LN, int STAGES, int SLOTS, int CAST, bool PSTAGED = (EPI != EPI_BIAS_RESIDUAL), bool EMBED = false, int kEpiWarps = 8>
int launch_gemm_staged_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tcast,
                           const GemmParams& p, int sm_count, cudaStream_t st, const CUtensorMap* tres = nullptr) {
    constexpr int kPreFloats = !PSTAGED ? 0 : (LN ? 256 + 256 + 2 * 128 : 256);
    using L = GemmStagedSmem<STAGES, SLOTS, CAST, kPreFloats * 4>;
    static_assert(L::DYN_BYTES <= 232448, "shared memory budget");
    auto kern = gemm_sm100_staged_kernel<T, STAGES, SLOTS, EPI, kEpiWarps, LN, CAST, PSTAGED, EMBED>;
    static int configured_dev_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured_dev_mask & (1 << dev))) {
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
        configured_dev_mask |= 1 << dev;
    }
    if (p.N % kGemmBN || p.K % GEMM_BK || p.M <= 0)
        return set_err(VIT_E_ARG, "gemm shape M=%d N=%d K=%d unsupported (N%%256, K%%64)", p.M, p.N, p.K);
    if (LN && (EPI == EPI_BIAS_RESIDUAL ? (p.N != kDim || !p.stats_out) : (p.K != kDim || !p.stats_in || !p.colsum || p.stats_parts <= 0)))
        return set_err(VIT_E_ARG, "LayerNorm-folded gemm: bad statistics arguments (N=%d K=%d)", p.N, p.K);
    const int tiles = ((p.M + 255) / 256) * (p.N / 256);
    const int grid = 2 * std::min(tiles, sm_count / 2);
    CU_TRY(launch_pdl(kern, dim3(grid), dim3((GEMM_NON_EPI_WARPS + kEpiWarps) * 32), L::DYN_BYTES, st, ta, tb, tout, tcast, tres ? *tres : tout, p));
    return check_launch("gemm_staged");
}
int res_cfg() {  // VIT_RES_CFG: A/B switch of the residual GEMM's shared-memory split (tuning)
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("VIT_RES_CFG");
        v = s ? atoi(s) : 0;
    }
    return v;
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
        if ((p.K >= 2048) != (res_cfg() == 1)) return launch_gemm_staged_cfg<T, EPI, LN, 5, 3, 1>(ta, tb, tout, tcast, p, sm_count, st);
        return launch_gemm_staged_cfg<T, EPI, LN, 4, 4, 2>(ta, tb, tout, tcast, p, sm_count, st);
    } else if constexpr (EPI == EPI_BIAS_RESIDUAL) {
        return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, kStagedSlots, 0>(ta, tb, tout, tcast, p, sm_count, st);
    } else if constexpr (LN) {
        // 6 KB of staged parameters cost one output slot; loading them in the epilogue threads instead (4 slots)
        // measured 7 % slower on mlp_0
        if (res_cfg() == 4) return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, kStagedSlots, 0, false>(ta, tb, tout, tcast, p, sm_count, st);
        if constexpr (EPI == EPI_BIAS_GELU) {
            if (res_cfg() == 6) return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, 3, 0, true, false, 16>(ta, tb, tout, tcast, p, sm_count, st);
        }
        return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, 3, 0, true>(ta, tb, tout, tcast, p, sm_count, st);
    } else {
        if (res_cfg() == 5) return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, kStagedSlots, 0, false>(ta, tb, tout, tcast, p, sm_count, st);
        return launch_gemm_staged_cfg<T, EPI, LN, kStagedStages, kStagedSlots, 0, true>(ta, tb, tout, tcast, p, sm_count, st);
    }
}
// conv_proj as the EMBED variant of the residual kernel (one CTA pair per image): ta 3-D patch map, tout 3-D fp32
// token map, tcast 3-D operand-precision token map (ln only), tpos 2-D pos_embedding map.  p.M = images * row tiles per image * 256.
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

template <typename T, int EPI>
int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int sm_count, cudaStream_t st) {
    if (p.N % kGemmBN || p.K % GEMM_BK || p.N > GEMM_MAX_N || p.M <= 0)
        return set_err(VIT_E_ARG, "gemm shape M=%d N=%d K=%d unsupported (N%%256, K%%64, N<=3072)", p.M, p.N, p.K);
    if (gemm_impl() == 2) return launch_gemm_pair_t<T, EPI>(ta, tb, p, sm_count, st);
    using L = GemmSmem<kGemmBN, kGemmStages>;
    auto kern = gemm_sm100_kernel<T, kGemmBN, kGemmStages, kGemmEpiWG, EPI>;
    static bool configured = false;  // per instantiation; attribute is per device but identical on all
    static int configured_dev_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured || !(configured_dev_mask & (1 << dev))) {
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
        configured = true;
        configured_dev_mask |= 1 << dev;
    }
    if (p.N % kGemmBN || p.K % GEMM_BK || p.N > GEMM_MAX_N || p.M <= 0)
        return set_err(VIT_E_ARG, "gemm shape M=%d N=%d K=%d unsupported (N%%256, K%%64, N<=3072)", p.M, p.N, p.K);
    const int tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * (p.N / kGemmBN);
    const int grid = std::min(tiles, sm_count);
    kern<<<grid, kGemmThreads, L::DYN_BYTES, st>>>(ta, tb, p);
    return check_launch("gemm");
}

template <int EPI>
int launch_gemm(int prec, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int sm_count,
                cudaStream_t st) {
    return prec == VIT_PREC_FP16 ? launch_gemm_t<__half, EPI>(ta, tb, p, sm_count, st)
                                 : launch_gemm_t<__nv_bfloat16, EPI>(ta, tb, p, sm_count, st);
}

template <typename T, bool EXACT>
int launch_attention_t(const CUtensorMap& tqkv, const CUtensorMap& tout, const AttnParams& p, int sm_count, cudaStream_t st) {
    auto kern = attention_sm100_persistent_kernel<T, EXACT>;
    const int smem = attn2_smem_bytes(p.kpad);
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<std::min(p.batch * kHeads, sm_count), ATTN2_THREADS, smem, st>>>(tqkv, tout, p);
    return check_launch("attention");
}
template <typename T, bool EXACT>
int launch_attention_stream_t(const CUtensorMap& tqkv, const CUtensorMap& tout32, const AttnParams& p, int sm_count, cudaStream_t st) {
    auto kern = attention_sm100_stream_kernel<T, EXACT>;
    const int smem = attn3_smem_bytes(p.kpad);
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU_TRY(launch_pdl(kern, dim3(std::min(p.batch * kHeads, sm_count)), dim3(ATTN3_THREADS), smem, st, tqkv, tout32, p));
    return check_launch("attention_stream");
}
template <typename T>
int launch_attention_stream_blocked_t(const CUtensorMap& tqkv, const CUtensorMap& tout32, const AttnParams& p, int sm_count, cudaStream_t st) {
    auto kern = attention_sm100_stream_blocked_kernel<T>;
    const int smem = attn4_smem_bytes(p.tokens);
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU_TRY(launch_pdl(kern, dim3(std::min(p.batch * kHeads, sm_count)), dim3(ATTN4_THREADS), smem, st, tqkv, tout32, p));
    return check_launch("attention_stream_blocked");
}
int attn_impl() {  // VIT_ATTN_IMPL=2: the two-slot persistent kernel instead of the streaming one (A/B testing)
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("VIT_ATTN_IMPL");
        v = (s && atoi(s) == 2) ? 2 : 3;
    }
    return v;
}
template <typename T>
int launch_attention_blocked_t(const CUtensorMap& tqkv, const CUtensorMap& tout, const AttnParams& p, int sm_count, cudaStream_t st) {
    auto kern = attention_sm100_blocked_kernel<T>;
    const int smem = attnl_smem_bytes(p.tokens);
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<std::min(p.batch * kHeads, sm_count), ATTNL_THREADS, smem, st>>>(tqkv, tout, p);
    return check_launch("attention_blocked");
}
constexpr int kAttnSingleBlockMaxTokens = 224;  // up to here the whole key range of an image is one S block in TMEM
// rows per box of the Q/K/V load map (make_tmap over the packed QKV activation) that launch_attention expects
int attention_load_box_rows(int tokens) {
    return tokens <= kAttnSingleBlockMaxTokens ? ((tokens + 15) / 16 * 16) / 2 : ATTNL_KB;
}
// tqkv: load map of the packed QKV activation with attention_load_box_rows(tokens) rows per box,
// tout: 3-D store map of the output (make_tmap_3d, 128 rows per box)
// exact: two-pass softmax (exact row maximum); otherwise the single-pass variant, which raises
// g_attn_range_flag when a row left its exponent window (the caller then repeats with exact = true).
// tout: 3-D store map of the output with 128-row boxes (persistent and key-blocked kernels), tout32: the same
// with 32-row boxes (streaming kernel, one store per output warp).
int launch_attention(int prec, const CUtensorMap& tqkv, const CUtensorMap& tout, const CUtensorMap& tout32, const AttnParams& p,
                     int sm_count, cudaStream_t st, bool exact) {
    if (p.tokens > ATTNL_MAX_TOKENS) return set_err(VIT_E_ARG, "attention: tokens=%d > %d not supported", p.tokens, ATTNL_MAX_TOKENS);
    const bool h = prec == VIT_PREC_FP16;
    if (p.tokens > kAttnSingleBlockMaxTokens) {
        // key-blocked: single-pass streaming kernel unless the exact softmax is asked for (or after a range flag)
        if (!exact && attn_impl() == 3)
            return h ? launch_attention_stream_blocked_t<__half>(tqkv, tout32, p, sm_count, st)
                     : launch_attention_stream_blocked_t<__nv_bfloat16>(tqkv, tout32, p, sm_count, st);
        return h ? launch_attention_blocked_t<__half>(tqkv, tout, p, sm_count, st) : launch_attention_blocked_t<__nv_bfloat16>(tqkv, tout, p, sm_count, st);
    }
    if (attn_impl() == 3) {
        if (exact) return h ? launch_attention_stream_t<__half, true>(tqkv, tout32, p, sm_count, st)
                            : launch_attention_stream_t<__nv_bfloat16, true>(tqkv, tout32, p, sm_count, st);
        return h ? launch_attention_stream_t<__half, false>(tqkv, tout32, p, sm_count, st)
                 : launch_attention_stream_t<__nv_bfloat16, false>(tqkv, tout32, p, sm_count, st);
    }
    if (exact) return h ? launch_attention_t<__half, true>(tqkv, tout, p, sm_count, st) : launch_attention_t<__nv_bfloat16, true>(tqkv, tout, p, sm_count, st);
    return h ? launch_attention_t<__half, false>(tqkv, tout, p, sm_count, st) : launch_attention_t<__nv_bfloat16, false>(tqkv, tout, p, sm_count, st);
}
// Reads and clears the current device's range flag (after the stream has been synchronised).
int take_attn_range_flag(bool* was_set) {
    unsigned int v = 0;
    CU_TRY(cudaMemcpyFromSymbol(&v, g_attn_range_flag, sizeof(v)));
    *was_set = v != 0;
    if (v) {
        v = 0;
        CU_TRY(cudaMemcpyToSymbol(g_attn_range_flag, &v, sizeof(v)));
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

template <typename T>
void launch_patchify_t(const float* img, T* patches, int batch, int S, cudaStream_t st) {
    const dim3 grid((S * (S / 4) + 511) / 512, 3 * batch);
    if (S == 224) patchify_kernel<T, 224><<<grid, 256, 0, st>>>(img, patches, batch, S);
    else if (S == 384) patchify_kernel<T, 384><<<grid, 256, 0, st>>>(img, patches, batch, S);
    else patchify_kernel<T, 0><<<grid, 256, 0, st>>>(img, patches, batch, S);
}
int launch_patchify(int prec, const float* img, void* patches, int batch, int S, int sm_count, cudaStream_t st) {
    (void)sm_count;
    if (prec == VIT_PREC_FP16) launch_patchify_t(img, static_cast<__half*>(patches), batch, S, st);
    else launch_patchify_t(img, static_cast<__nv_bfloat16*>(patches), batch, S, st);
    return check_launch("patchify");
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

struct LayerW {
    float *ln1_w, *ln1_b, *qkv_b, *out_b, *ln2_w, *ln2_b, *fc1_b, *fc2_b;  // fp32
    void *qkv_w, *out_w, *fc1_w, *fc2_w;                                    // operand precision
    CUtensorMap tm_qkv_w, tm_out_w, tm_fc1_w, tm_fc2_w;
    // LayerNorm folded into in_proj / mlp_0: W' = ln_w (.) W, column sums of W', c = bias + W ln_b
    void *qkv_wf = nullptr, *fc1_wf = nullptr;
    float *qkv_s = nullptr, *qkv_c = nullptr, *fc1_s = nullptr, *fc1_c = nullptr;
    CUtensorMap tm_qkv_wf, tm_fc1_wf;
};

struct DeviceCtx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    std::vector<void*> allocs;
    // weights
    float *cls = nullptr, *conv_b = nullptr, *pos = nullptr, *lnf_w = nullptr, *lnf_b = nullptr, *head_w = nullptr,
          *head_b = nullptr;
    void* conv_w = nullptr;
    CUtensorMap tm_conv_w;
    LayerW layer[kDepth];
    // workspace (capacity = max_batch images)
    float* images[2] = {nullptr, nullptr};
    void *patches = nullptr, *xn = nullptr, *qkv = nullptr, *ao = nullptr, *hid = nullptr;
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
    CUtensorMap tm_ao_c, tm_x_c, tm_xn_c, tm_hid_c, tm_hid_c32;
    float2* pstats = nullptr;  // [6][max rows] partial (sum, sum of squares) of the residual rows (LN folding)
    size_t stats_rows = 0;
    // activation tensor maps, rebuilt when the pass size changes: row extent = rows actually in
    // use, so TMA zero-fills loads and clips stores past the last image
    int maps_nb = -1;
    CUtensorMap tm_patches3, tm_x3, tm_xn3 /* per-image 3-D views for conv_proj */, tm_pos, tm_patches, tm_xn, tm_ao, tm_hid, tm_qkv_st, tm_hid32, tm_qkv_st32 /* 32-row store boxes */, tm_x, tm_q /* attention loads */, tm_kv /* attention store */, tm_kv32 /* attention store, 32-row boxes */;
    size_t ws_bytes = 0;
    // optional per-kernel-category timing (vit_cuda_profile_*): event pairs around launches
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[VIT_PROF_NCAT];
    size_t prof_used[VIT_PROF_NCAT] = {};
    timerEvents_t timer;
    struct Graph {   // captured launch sequence of one small pass
        int nb = 0;
        const float* images = nullptr;
        float* logits = nullptr;
        bool attn_exact = false, ln_fused = true, prune_last = true;
        long long launches = 0;
        cudaGraphExec_t exec = nullptr;
    };
    std::vector<Graph> graphs;
};

struct Engine {
    bool up = false;
    bool profiling = false;
    bool prune_last = true;       // last layer: everything behind the attention for the class rows only (VIT_PRUNE_LAST=0: all rows)
    bool ln_fused = true;         // LayerNorm folded into the GEMMs (VIT_LN_FUSED=0: separate LayerNorm kernels)
    bool attn_exact = false;      // two-pass softmax (VIT_ATTN_EXACT=1, vit_cuda_set_attention_exact, or after a range flag)
    long long attn_fallbacks = 0; // forwards repeated with the exact softmax
    int img = 0, grid = 0, patches = 0, tokens = 0, max_batch = 0, prec = 0;
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

int upload_f32(DeviceCtx& c, float** dst, const vit_tensor& t) {
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(dst), t.size * sizeof(float), false));
    CU_TRY(cudaMemcpyAsync(*dst, t.data, t.size * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    return 0;
}

int upload_operand(DeviceCtx& c, void** dst, const vit_tensor& t, float* scratch, int prec) {
    VIT_TRY(dev_alloc(c, dst, t.size * 2, false));
    CU_TRY(cudaMemcpyAsync(scratch, t.data, t.size * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    return launch_convert_from_f32(prec, scratch, *dst, t.size, c.stream);
}

// Folded copy of a [N][768] weight whose fp32 values are still in `scratch` (just uploaded by
// upload_operand): W' = ln_w (.) W in the operand precision, its column sums and c = bias + W ln_b.
int upload_folded(DeviceCtx& c, void** wf, float** colsum, float** cvec, const float* scratch, const float* ln_w,
                  const float* ln_b, const float* bias, int N, int prec) {
    VIT_TRY(dev_alloc(c, wf, static_cast<size_t>(N) * kDim * 2, false));
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(colsum), static_cast<size_t>(N) * 4, false));
    VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(cvec), static_cast<size_t>(N) * 4, false));
    return launch_fold_ln(prec, scratch, ln_w, ln_b, bias, *wf, *colsum, *cvec, N, c.stream);
}

void destroy_ctx(DeviceCtx& c) {
    if (c.device < 0) return;
    cudaSetDevice(c.device);
    if (c.stream) cudaStreamSynchronize(c.stream);
    for (int i = 0; i < 2; ++i) {
        if (c.h_stage[i]) cudaFreeHost(c.h_stage[i]);
        if (c.ev_stage[i]) cudaEventDestroy(c.ev_stage[i]);
        if (c.h_logits[i]) cudaFreeHost(c.h_logits[i]);
        if (c.ev_logits[i]) cudaEventDestroy(c.ev_logits[i]);
    }
    for (auto& g : c.graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c.graphs.clear();
    for (void* p : c.allocs) cudaFree(p);
    c.allocs.clear();
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

int init_ctx(DeviceCtx& c, int device, const vit_tensor* w, const Engine& e) {
    c.device = device;
    VIT_TRY(check_device(device, &c.sm_count));
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&c.ev_h2d[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c.ev_done[i], cudaEventDisableTiming));
    }
    const int prec = e.prec;
    // ---- weights: fp32 small tensors verbatim, GEMM operands converted once
    float* scratch = nullptr;
    CU_TRY(cudaMalloc(&scratch, static_cast<size_t>(kHidden) * kDim * sizeof(float)));
    int rc = 0;
    do {
        if ((rc = upload_f32(c, &c.cls, w[0]))) break;
        if ((rc = upload_operand(c, &c.conv_w, w[1], scratch, prec))) break;
        if ((rc = upload_f32(c, &c.conv_b, w[2]))) break;
        if ((rc = upload_f32(c, &c.pos, w[3]))) break;
        for (int l = 0; l < kDepth && !rc; ++l) {
            const vit_tensor* lw = w + 4 + 12 * l;
            LayerW& L = c.layer[l];
            if ((rc = upload_f32(c, &L.ln1_w, lw[0]))) break;
            if ((rc = upload_f32(c, &L.ln1_b, lw[1]))) break;
            if ((rc = upload_operand(c, &L.qkv_w, lw[2], scratch, prec))) break;
            if ((rc = upload_f32(c, &L.qkv_b, lw[3]))) break;
            if ((rc = upload_folded(c, &L.qkv_wf, &L.qkv_s, &L.qkv_c, scratch, L.ln1_w, L.ln1_b, L.qkv_b, 3 * kDim, prec))) break;
            if ((rc = upload_operand(c, &L.out_w, lw[4], scratch, prec))) break;
            if ((rc = upload_f32(c, &L.out_b, lw[5]))) break;
            if ((rc = upload_f32(c, &L.ln2_w, lw[6]))) break;
            if ((rc = upload_f32(c, &L.ln2_b, lw[7]))) break;
            if ((rc = upload_operand(c, &L.fc1_w, lw[8], scratch, prec))) break;
            if ((rc = upload_f32(c, &L.fc1_b, lw[9]))) break;
            if ((rc = upload_folded(c, &L.fc1_wf, &L.fc1_s, &L.fc1_c, scratch, L.ln2_w, L.ln2_b, L.fc1_b, kHidden, prec))) break;
            if ((rc = upload_operand(c, &L.fc2_w, lw[10], scratch, prec))) break;
            if ((rc = upload_f32(c, &L.fc2_b, lw[11]))) break;
            if ((rc = make_tmap(&L.tm_qkv_w, prec, L.qkv_w, kDim, 3 * kDim, GEMM_BK, 128))) break;
            if ((rc = make_tmap(&L.tm_out_w, prec, L.out_w, kDim, kDim, GEMM_BK, 128))) break;
            if ((rc = make_tmap(&L.tm_fc1_w, prec, L.fc1_w, kDim, kHidden, GEMM_BK, 128))) break;
            if ((rc = make_tmap(&L.tm_fc2_w, prec, L.fc2_w, kHidden, kDim, GEMM_BK, 128))) break;
            if ((rc = make_tmap(&L.tm_qkv_wf, prec, L.qkv_wf, kDim, 3 * kDim, GEMM_BK, 128))) break;
            if ((rc = make_tmap(&L.tm_fc1_wf, prec, L.fc1_wf, kDim, kHidden, GEMM_BK, 128))) break;
        }
        if (rc) break;
        if ((rc = upload_f32(c, &c.lnf_w, w[148]))) break;
        if ((rc = upload_f32(c, &c.lnf_b, w[149]))) break;
        if ((rc = upload_f32(c, &c.head_w, w[150]))) break;
        if ((rc = upload_f32(c, &c.head_b, w[151]))) break;
        if ((rc = make_tmap(&c.tm_conv_w, prec, c.conv_w, kDim, kDim, GEMM_BK, 128))) break;
    } while (0);
    cudaError_t se = cudaStreamSynchronize(c.stream);
    cudaFree(scratch);
    if (rc) return rc;
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, "weight upload");

    // ---- workspaces.  Zero-filled once: attention may read (and multiply by P == 0) rows
    // past the last image of a pass, which must therefore always hold finite values.
    const size_t B = e.max_batch, rows = B * e.tokens, prow = B * e.patches;
    const size_t img_elems = static_cast<size_t>(3) * e.img * e.img;
    for (int i = 0; i < 2; ++i) VIT_TRY(dev_alloc(c, reinterpret_cast<void**>(&c.images[i]), B * img_elems * 4, false));
    VIT_TRY(dev_alloc(c, &c.patches, prow * kDim * 2, true));
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

int ensure_maps(DeviceCtx& c, const Engine& e, int nb) {
    if (c.maps_nb == nb) return 0;
    const int prec = e.prec;
    const uint64_t rows = static_cast<uint64_t>(nb) * e.tokens, prow = static_cast<uint64_t>(nb) * e.patches;
    VIT_TRY(make_tmap(&c.tm_patches, prec, c.patches, kDim, prow, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_xn, prec, c.xn, kDim, rows, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_ao, prec, c.ao, kDim, rows, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&c.tm_hid, prec, c.hid, kHidden, rows, GEMM_BK, GEMM_BM));     // mlp_0 store + mlp_3 A load
    VIT_TRY(make_tmap(&c.tm_qkv_st, prec, c.qkv, 3 * kDim, rows, GEMM_BK, GEMM_BM)); // in_proj store
    VIT_TRY(make_tmap_3d(&c.tm_patches3, prec, c.patches, kDim, e.patches, nb, GEMM_BM));
    VIT_TRY(make_tmap_3d_f32(&c.tm_x3, c.x, kDim, e.tokens, nb, GEMM_BM));
    VIT_TRY(make_tmap_3d(&c.tm_xn3, prec, c.xn, kDim, e.tokens, nb, GEMM_BM));
    VIT_TRY(make_tmap_f32(&c.tm_pos, c.pos, kDim, e.tokens, GEMM_BM));
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

// Enqueue the whole forward for nb images already resident in d_images (fp32 NCHW) on c.stream.
int enqueue_forward_kernels(DeviceCtx& c, const Engine& e, const float* d_images, int nb, float* d_logits) {
    cudaStream_t st = c.stream;
    const int prec = e.prec;
    const int rows = nb * e.tokens;
    VIT_TRY(ensure_maps(c, e, nb));
    const bool staged = gemm_impl() == 2;
    // conv_proj: patch rows -> GEMM with (+bias, +pos_embedding, row remap) epilogue; class rows aside
    const bool pf = e.profiling;
    // LayerNorm folded into the GEMMs (default): in_proj / mlp_0 read the operand-precision copy of the raw
    // residual rows (c.xn) with the per-row statistics (c.pstats) that the previous residual GEMM -- for
    // layer 0 conv_proj and the class-row kernel -- left behind.  Otherwise: a LayerNorm kernel before each.
    const bool fused = staged && e.ln_fused;
    const int stats_rows = static_cast<int>(c.stats_rows);
    {
        ProfScope ps(c, pf, VIT_PROF_PATCHIFY);
        VIT_TRY(launch_patchify(prec, d_images, c.patches, nb, e.img, c.sm_count, st));
        if (fused) {
            if (prec == VIT_PREC_FP16)
                cls_rows_ln_kernel<__half><<<(nb + 7) / 8, 256, 0, st>>>(c.x, static_cast<__half*>(c.xn), c.pstats, stats_rows, c.cls, c.pos, nb, e.tokens);
            else
                cls_rows_ln_kernel<__nv_bfloat16><<<(nb + 7) / 8, 256, 0, st>>>(c.x, static_cast<__nv_bfloat16*>(c.xn), c.pstats, stats_rows, c.cls, c.pos, nb, e.tokens);
        } else {
            cls_rows_kernel<<<(nb * kDim + 255) / 256, 256, 0, st>>>(c.x, c.cls, c.pos, nb, e.tokens);
        }
        VIT_TRY(check_launch("cls_rows"));
    }
    {
        ProfScope ps(c, pf, VIT_PROF_EMBED_GEMM);
        if (staged) {   // one CTA pair per image; class_token / pos_emb / token layout are pure TMA addressing
            GemmParams p{nb * ((e.patches + 255) / 256) * 256, kDim, kDim, c.conv_b, c.x, nullptr, e.patches, e.tokens};
            if (fused) {
                p.stats_out = c.pstats;
                p.stats_rows = stats_rows;
            }
            VIT_TRY(launch_gemm_embed(prec, fused, c.tm_patches3, c.tm_conv_w, c.tm_x3, c.tm_xn3, c.tm_pos, p, c.sm_count, st));
        } else {
            GemmParams p{nb * e.patches, kDim, kDim, c.conv_b, c.x, c.pos, e.patches, e.tokens};
            VIT_TRY(launch_gemm<EPI_PATCH_EMBED>(prec, c.tm_patches, c.tm_conv_w, p, c.sm_count, st));
        }
    }
    AttnParams ap{nb, e.tokens, (e.tokens + 15) / 16 * 16, c.ao, 0.125f * 1.4426950408889634f, attn_no_pingpong(), nullptr};
    bool pruned_tail = false;
    for (int l = 0; l < kDepth; ++l) {
        const LayerW& L = c.layer[l];
        if (!fused) {
            ProfScope ps(c, pf, VIT_PROF_LAYERNORM);
            VIT_TRY(launch_layernorm(prec, c.x, L.ln1_w, L.ln1_b, c.xn, rows, st));
        }
        {
            ProfScope ps(c, pf, VIT_PROF_QKV_GEMM);
            GemmParams p{rows, 3 * kDim, kDim, L.qkv_b, c.qkv, nullptr, 0, 0, 2 * kDim};  // V block stored as bf16
            if (fused) {
                p.bias = L.qkv_c;
                p.colsum = L.qkv_s;
                p.stats_in = c.pstats;
                p.stats_parts = 6;
                p.stats_rows = stats_rows;
                VIT_TRY(launch_gemm_staged_ln<EPI_BIAS>(prec, c.tm_xn, L.tm_qkv_wf, c.tm_qkv_st, c.tm_qkv_st32, p, c.sm_count, st));
            } else if (staged) VIT_TRY(launch_gemm_staged<EPI_BIAS>(prec, c.tm_xn, L.tm_qkv_w, c.tm_qkv_st, c.tm_qkv_st32, p, c.sm_count, st));
            else VIT_TRY(launch_gemm<EPI_BIAS>(prec, c.tm_xn, L.tm_qkv_w, p, c.sm_count, st));
        }
        if (l == kDepth - 1 && fused && e.prune_last) {
            // Last layer: only the class token reaches the head, and no token reads another one after the attention.
            // Attention for the class query alone (all keys and values), then out_proj / LayerNorm / MLP on the compact
            // [nb][768] class rows: 1/197 of the rows of the other layers.
            {
                ProfScope ps(c, pf, VIT_PROF_ATTENTION);
                if (prec == VIT_PREC_FP16)
                    cls_attention_kernel<__half><<<nb, 384, 0, st>>>(static_cast<const uint16_t*>(c.qkv), c.x, static_cast<__half*>(c.ao_c), c.x_c, e.tokens);
                else
                    cls_attention_kernel<__nv_bfloat16><<<nb, 384, 0, st>>>(static_cast<const uint16_t*>(c.qkv), c.x, static_cast<__nv_bfloat16*>(c.ao_c), c.x_c, e.tokens);
                VIT_TRY(check_launch("cls_attention"));
            }
            const int srows_c = static_cast<int>(c.stats_rows_c);
            {
                ProfScope ps(c, pf, VIT_PROF_OUT_GEMM);
                GemmParams p{nb, kDim, kDim, L.out_b, c.x_c, c.x_c, 0, 0};
                p.stats_out = c.pstats_c;
                p.stats_rows = srows_c;
                VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_RESIDUAL>(prec, c.tm_ao_c, L.tm_out_w, c.tm_x_c, c.tm_xn_c, p, c.sm_count, st));
            }
            {
                ProfScope ps(c, pf, VIT_PROF_FC1_GEMM);
                GemmParams p{nb, kHidden, kDim, L.fc1_c, c.hid_c, nullptr, 0, 0};
                p.colsum = L.fc1_s;
                p.stats_in = c.pstats_c;
                p.stats_parts = 6;
                p.stats_rows = srows_c;
                VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_GELU>(prec, c.tm_xn_c, L.tm_fc1_wf, c.tm_hid_c, c.tm_hid_c32, p, c.sm_count, st));
            }
            {
                ProfScope ps(c, pf, VIT_PROF_FC2_GEMM);
                GemmParams p{nb, kDim, kHidden, L.fc2_b, c.x_c, c.x_c, 0, 0};
                VIT_TRY(launch_gemm_staged<EPI_BIAS_RESIDUAL>(prec, c.tm_hid_c, L.tm_fc2_w, c.tm_x_c, c.tm_x_c, p, c.sm_count, st));
            }
            pruned_tail = true;
            break;
        }
        {
            ProfScope ps(c, pf, VIT_PROF_ATTENTION);
            VIT_TRY(launch_attention(prec, c.tm_q, c.tm_kv, c.tm_kv32, ap, c.sm_count, st, e.attn_exact));
        }
        {
            ProfScope ps(c, pf, VIT_PROF_OUT_GEMM);
            GemmParams p{rows, kDim, kDim, L.out_b, c.x, c.x, 0, 0};
            if (fused) {
                p.stats_out = c.pstats;
                p.stats_rows = stats_rows;
                VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_RESIDUAL>(prec, c.tm_ao, L.tm_out_w, c.tm_x, c.tm_xn, p, c.sm_count, st));
            } else if (staged) VIT_TRY(launch_gemm_staged<EPI_BIAS_RESIDUAL>(prec, c.tm_ao, L.tm_out_w, c.tm_x, c.tm_x, p, c.sm_count, st));
            else VIT_TRY(launch_gemm<EPI_BIAS_RESIDUAL>(prec, c.tm_ao, L.tm_out_w, p, c.sm_count, st));
        }
        if (!fused) {
            ProfScope ps(c, pf, VIT_PROF_LAYERNORM);
            VIT_TRY(launch_layernorm(prec, c.x, L.ln2_w, L.ln2_b, c.xn, rows, st));
        }
        {
            ProfScope ps(c, pf, VIT_PROF_FC1_GEMM);
            GemmParams p{rows, kHidden, kDim, L.fc1_b, c.hid, nullptr, 0, 0};
            if (fused) {
                p.bias = L.fc1_c;
                p.colsum = L.fc1_s;
                p.stats_in = c.pstats;
                p.stats_parts = 6;
                p.stats_rows = stats_rows;
                VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_GELU>(prec, c.tm_xn, L.tm_fc1_wf, c.tm_hid, c.tm_hid32, p, c.sm_count, st));
            } else if (staged) VIT_TRY(launch_gemm_staged<EPI_BIAS_GELU>(prec, c.tm_xn, L.tm_fc1_w, c.tm_hid, c.tm_hid32, p, c.sm_count, st));
            else VIT_TRY(launch_gemm<EPI_BIAS_GELU>(prec, c.tm_xn, L.tm_fc1_w, p, c.sm_count, st));
        }
        {
            ProfScope ps(c, pf, VIT_PROF_FC2_GEMM);
            GemmParams p{rows, kDim, kHidden, L.fc2_b, c.x, c.x, 0, 0};
            if (fused && l + 1 < kDepth) {   // the last layer's output only feeds the class-row LayerNorm of the head
                p.stats_out = c.pstats;
                p.stats_rows = stats_rows;
                VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_RESIDUAL>(prec, c.tm_hid, L.tm_fc2_w, c.tm_x, c.tm_xn, p, c.sm_count, st));
            } else if (staged) VIT_TRY(launch_gemm_staged<EPI_BIAS_RESIDUAL>(prec, c.tm_hid, L.tm_fc2_w, c.tm_x, c.tm_x, p, c.sm_count, st));
            else VIT_TRY(launch_gemm<EPI_BIAS_RESIDUAL>(prec, c.tm_hid, L.tm_fc2_w, p, c.sm_count, st));
        }
    }
    ProfScope ps(c, pf, VIT_PROF_HEAD);
    if (pruned_tail) head_ln_kernel<<<(nb + 7) / 8, 256, 0, st>>>(c.x_c, c.lnf_w, c.lnf_b, c.cls_ln, nb, 1);
    else head_ln_kernel<<<(nb + 7) / 8, 256, 0, st>>>(c.x, c.lnf_w, c.lnf_b, c.cls_ln, nb, e.tokens);
    VIT_TRY(check_launch("head_ln"));
    head_gemm_kernel<<<dim3((kClasses + HEAD_CLASSES - 1) / HEAD_CLASSES, std::min((nb + HEAD_IMGS - 1) / HEAD_IMGS, 32)), 256, 0, st>>>(c.cls_ln, c.head_w, c.head_b, d_logits, nb,
                                                                                kClasses);
    return check_launch("head_gemm");
}

// Small passes (batch-1 latency, BASELINE.json configs[1]) are launch bound: ~65 kernels of 5-20 us each.  Their
// launch sequence is captured once per (pass size, buffers, softmax mode) into a CUDA graph and replayed with a
// single launch.  VIT_GRAPHS=0 disables this.
constexpr int kGraphMaxBatch = 8;
bool graphs_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("VIT_GRAPHS");
        v = (s && atoi(s) == 0) ? 0 : 1;
    }
    return v == 1;
}
int enqueue_forward(DeviceCtx& c, const Engine& e, const float* d_images, int nb, float* d_logits) {
    if (nb > kGraphMaxBatch || e.profiling || !graphs_enabled() || gemm_impl() != 2)
        return enqueue_forward_kernels(c, e, d_images, nb, d_logits);
    for (auto& g : c.graphs)
        if (g.nb == nb && g.images == d_images && g.logits == d_logits && g.attn_exact == e.attn_exact && g.ln_fused == e.ln_fused && g.prune_last == e.prune_last) {
            CU_TRY(cudaGraphLaunch(g.exec, c.stream));
            g_launches.fetch_add(g.launches, std::memory_order_relaxed);
            return 0;
        }
    // first time: run it once the ordinary way (sets the kernels' attributes, builds the tensor maps), then capture
    VIT_TRY(enqueue_forward_kernels(c, e, d_images, nb, d_logits));
    if (c.graphs.size() >= 16) return 0;   // a caller cycling through many buffers: stay with plain launches
    cudaGraph_t graph = nullptr;
    const long long before = g_launches.load();
    CU_TRY(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_forward_kernels(c, e, d_images, nb, d_logits);
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
    g.images = d_images;
    g.logits = d_logits;
    g.attn_exact = e.attn_exact;
    g.ln_fused = e.ln_fused;
    g.prune_last = e.prune_last;
    g.launches = captured;
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return set_err(VIT_E_CUDA, "graph instantiate: %s", cudaGetErrorString(ie));
    c.graphs.push_back(g);
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
    if (img_size <= 0 || img_size % kPatch || img_size > 1024) return set_err(VIT_E_ARG, "img_size %d must be a positive multiple of 16", img_size);
    if (max_batch_per_gpu <= 0 || n_gpus <= 0) return set_err(VIT_E_ARG, "max_batch_per_gpu and n_gpus must be positive");
    if (precision != VIT_PREC_BF16 && precision != VIT_PREC_FP16) return set_err(VIT_E_ARG, "unknown precision %d", precision);
    for (int i = 0; i < n_tensors; ++i)
        if (!networks[i].data || networks[i].size != tensor_numel(i, img_size))
            return set_err(VIT_E_ARG, "weight tensor %d: have %zu floats%s, need %zu for img_size %d", i, networks[i].size,
                           networks[i].data ? "" : " (missing)", tensor_numel(i, img_size), img_size);
    Engine& e = g_eng;
    e.img = img_size;
    e.grid = img_size / kPatch;
    e.patches = e.grid * e.grid;
    e.tokens = e.patches + 1;
    e.max_batch = max_batch_per_gpu;
    e.prec = precision;
    {
        const char* ex = getenv("VIT_ATTN_EXACT");
        e.attn_exact = ex && atoi(ex) != 0;
        e.attn_fallbacks = 0;
        const char* lf = getenv("VIT_LN_FUSED");
        e.ln_fused = !(lf && atoi(lf) == 0);
        const char* pl = getenv("VIT_PRUNE_LAST");
        e.prune_last = !(pl && atoi(pl) == 0);
    }
    e.ctx.assign(n_gpus, DeviceCtx());
    for (int g = 0; g < n_gpus; ++g) {
        const int rc = init_ctx(e.ctx[g], device_ids ? device_ids[g] : g, networks, e);
        if (rc) {
            char keep[sizeof(t_err)];
            memcpy(keep, t_err, sizeof(keep));
            vit_cuda_free();
            memcpy(t_err, keep, sizeof(keep));
            return rc;
        }
    }
    e.up = true;
    return 0;
}

int vit_cuda_init(const vit_tensor* networks, int n_tensors, int img_size, int max_batch_per_gpu, int n_gpus) {
    return vit_cuda_init_ex(networks, n_tensors, img_size, max_batch_per_gpu, n_gpus, nullptr, VIT_PREC_BF16);
}

int vit_cuda_enqueue_device(int gpu_slot, const float* d_images, int n, float* d_logits) {
    Engine& e = g_eng;
    if (!e.up) return set_err(VIT_E_ARG, "engine not initialised");
    if (gpu_slot < 0 || gpu_slot >= (int)e.ctx.size()) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
    if (n <= 0 || n > e.max_batch) return set_err(VIT_E_ARG, "n=%d outside (0, max_batch=%d]", n, e.max_batch);
    if (e.tokens > ATTNL_MAX_TOKENS) return set_err(VIT_E_ARG, "img_size %d (%d tokens): at most %d tokens are supported", e.img, e.tokens, ATTNL_MAX_TOKENS);
    if (gemm_impl() != 2 && e.prec != VIT_PREC_BF16) return set_err(VIT_E_ARG, "VIT_GEMM_IMPL=1 (A/B test kernels) supports bf16 only");
    DeviceCtx& c = e.ctx[gpu_slot];
    CU_TRY(cudaSetDevice(c.device));
    return enqueue_forward(c, e, d_images, n, d_logits);
}

int vit_cuda_sync(int gpu_slot) {
    Engine& e = g_eng;
    if (!e.up || gpu_slot < 0 || gpu_slot >= (int)e.ctx.size()) return set_err(VIT_E_ARG, "bad gpu slot %d", gpu_slot);
    CU_TRY(cudaSetDevice(e.ctx[gpu_slot].device));
    const cudaError_t se = cudaStreamSynchronize(e.ctx[gpu_slot].stream);
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, "forward");
    if (!e.attn_exact) {
        bool flagged = false;
        VIT_TRY(take_attn_range_flag(&flagged));
        if (flagged) {
            e.attn_exact = true;  // device-resident callers re-enqueue; the engine stays on the exact softmax
            ++e.attn_fallbacks;
            return set_err(VIT_E_RANGE, "attention: a row's scores left the single-pass softmax's exponent window; the results of "
                                        "this pass are invalid.  The engine has switched to the exact two-pass softmax: enqueue again");
        }
    }
    return 0;
}

int vit_cuda_set_class_row_pruning(int on) {
    if (!g_eng.up) return set_err(VIT_E_ARG, "engine not initialised");
    g_eng.prune_last = on != 0;
    return 0;
}

int vit_cuda_set_attention_exact(int on) {
    if (!g_eng.up) return set_err(VIT_E_ARG, "engine not initialised");
    g_eng.attn_exact = on != 0;
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

int vit_cuda_pass_schedule_ex(int n_images, int max_batch, int staged, int* first, int* count, int cap) {
    if (n_images < 0 || max_batch <= 0 || !first || !count || cap <= 0) return set_err(VIT_E_ARG, "bad schedule arguments");
    int n = 0;
    // pinned input: the copy of a pass is ~3.4x faster than its kernels -> passes may triple.  Staged input (pageable or
    // one allocation per image): the host-side gather runs at about the rate the GPU consumes images, so the passes
    // after the first stay at 128 images -- each gather hides under the kernels of the pass before it.
    for (int done = 0, sz = staged ? 64 : 32; done < n_images; sz = staged ? std::min(max_batch, 128) : std::min(max_batch, 3 * sz)) {
        const int nb = std::min(std::min(sz, max_batch), n_images - done);
        if (n == cap) return set_err(VIT_E_ARG, "pass schedule of %d images with max_batch %d needs more than %d passes", n_images, max_batch, cap);
        first[n] = done;
        count[n] = nb;
        done += nb;
        ++n;
    }
    return n;
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

static int forward_host_once(const float* images_nchw, const float* const* image_ptrs, int n, float* logits_out, bool* range_flag);

static int forward_host(const float* images_nchw, const float* const* image_ptrs, int n, float* logits_out, int* top1_out);
// hand the logits parked in pinned staging buffer `buf` (if any) over to the caller's (pageable) array
static int drain_logits(DeviceCtx& c, int buf, float* logits_out) {
    if (c.pend_count[buf] == 0) return 0;
    CU_TRY(cudaEventSynchronize(c.ev_logits[buf]));
    memcpy(logits_out + static_cast<size_t>(c.pend_first[buf]) * kClasses, c.h_logits[buf], static_cast<size_t>(c.pend_count[buf]) * kClasses * sizeof(float));
    c.pend_count[buf] = 0;
    return 0;
}

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

static int forward_host(const float* images_nchw, const float* const* image_ptrs, int n, float* logits_out, int* top1_out) {
    Engine& e = g_eng;
    if (!e.up) return set_err(VIT_E_ARG, "engine not initialised");
    if (!logits_out || n < 0) return set_err(VIT_E_ARG, "bad arguments");
    if (n == 0) return 0;
    bool flagged = false;
    VIT_TRY(forward_host_once(images_nchw, image_ptrs, n, logits_out, &flagged));
    if (flagged) {
        // some row left the single-pass softmax's exponent window: repeat with the exact two-pass softmax
        // (and stay there once this has happened three times -- the data evidently does it regularly)
        ++e.attn_fallbacks;
        e.attn_exact = true;
        const int rc = forward_host_once(images_nchw, image_ptrs, n, logits_out, &flagged);
        if (e.attn_fallbacks < 3) e.attn_exact = false;
        VIT_TRY(rc);
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

static int forward_host_once(const float* images_nchw, const float* const* image_ptrs, int n, float* logits_out, bool* range_flag) {
    Engine& e = g_eng;
    *range_flag = false;
    if (e.tokens > ATTNL_MAX_TOKENS) return set_err(VIT_E_ARG, "img_size %d (%d tokens): at most %d tokens are supported", e.img, e.tokens, ATTNL_MAX_TOKENS);
    const int G = static_cast<int>(e.ctx.size());
    const size_t img_elems = static_cast<size_t>(3) * e.img * e.img;
    const int per_gpu = (n + G - 1) / G;  // contiguous shards
    // Pass schedule of a shard: the H2D copy of pass i+1 (copy stream, second image buffer) hides under the
    // kernels of pass i, so only the FIRST pass's copy is exposed -- it is kept small (32 images, 19 MB), and
    // every later pass may be three times the previous one (PCIe Gen5 moves images ~3.4x faster than
    // the kernels consume them) up to the workspace size.  1024 images: 32 + 96 + 288 + 608.
    cudaPointerAttributes pa;
    const bool logits_pinned = cudaPointerGetAttributes(&pa, logits_out) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const bool images_pinned = images_nchw && cudaPointerGetAttributes(&pa, images_nchw) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    // any number of passes (a tiny max_batch with a large n): size the schedule arrays for the worst case
    const int worst = per_gpu / std::min(e.max_batch, 32) + 8;
    std::vector<int> pass_first(worst), pass_count(worst);
    const int n_sched = vit_cuda_pass_schedule_ex(per_gpu, e.max_batch, (image_ptrs || !images_pinned) ? 1 : 0, pass_first.data(), pass_count.data(), worst);
    if (n_sched < 0) return n_sched;
    pass_first.resize(n_sched);
    pass_count.resize(n_sched);
    const int max_passes = static_cast<int>(pass_first.size());
    // pass-major issue order so that all GPUs are fed before any host-side wait
    for (int pass = 0; pass < max_passes; ++pass) {
        for (int g = 0; g < G; ++g) {
            DeviceCtx& c = e.ctx[g];
            int lo, hi;
            vit_cuda_shard_range(n, G, g, &lo, &hi);
            const int first = lo + pass_first[pass];
            if (first >= hi) continue;
            const int nb = std::min(pass_count[pass], hi - first);
            const int buf = pass & 1;
            CU_TRY(cudaSetDevice(c.device));
            // H2D of this pass overlaps the previous pass's compute (other image buffer)
            if (pass >= 2) CU_TRY(cudaStreamWaitEvent(c.copy_stream, c.ev_done[buf], 0));
            const float* src = nullptr;
            if (image_ptrs || !images_pinned) {
                // separately allocated images (the reference's loader, Network.c:75-93) or pageable memory: gather this
                // pass into the slot's pinned staging buffer while the GPU works on the previous pass
                if (!c.h_stage[buf]) {
                    CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c.h_stage[buf]), static_cast<size_t>(e.max_batch) * img_elems * sizeof(float), cudaHostAllocPortable));
                    CU_TRY(cudaEventCreateWithFlags(&c.ev_stage[buf], cudaEventDisableTiming));
                } else {
                    CU_TRY(cudaEventSynchronize(c.ev_stage[buf]));   // its previous copy has left the buffer
                }
                // One host thread copies ~10 GB/s: 1024 images (617 MB) would take longer than the GPU needs for them.
                // Up to eight threads (half the host cores) share a pass.
                float* dst = c.h_stage[buf];
                auto gather = [=](int i0, int i1) {
                    for (int i = i0; i < i1; ++i)
                        memcpy(dst + static_cast<size_t>(i) * img_elems,
                               image_ptrs ? image_ptrs[first + i] : images_nchw + static_cast<size_t>(first + i) * img_elems, img_elems * sizeof(float));
                };
                const int n_thr = std::min(std::max(1, static_cast<int>(std::thread::hardware_concurrency()) / 2), std::min(8, (nb + 15) / 16));
                if (n_thr <= 1) {
                    gather(0, nb);
                } else {
                    std::vector<std::thread> pool;
                    for (int t = 1; t < n_thr; ++t) pool.emplace_back(gather, nb * t / n_thr, nb * (t + 1) / n_thr);
                    gather(0, nb / n_thr);
                    for (auto& th : pool) th.join();
                }
                src = c.h_stage[buf];
            } else {
                src = images_nchw + static_cast<size_t>(first) * img_elems;
            }
            CU_TRY(cudaMemcpyAsync(c.images[buf], src, static_cast<size_t>(nb) * img_elems * sizeof(float), cudaMemcpyHostToDevice, c.copy_stream));
            if (src == c.h_stage[buf]) CU_TRY(cudaEventRecord(c.ev_stage[buf], c.copy_stream));
            CU_TRY(cudaEventRecord(c.ev_h2d[buf], c.copy_stream));
            CU_TRY(cudaStreamWaitEvent(c.stream, c.ev_h2d[buf], 0));
            VIT_TRY(enqueue_forward(c, e, c.images[buf], nb, c.logits));
            CU_TRY(cudaEventRecord(c.ev_done[buf], c.stream));
            if (logits_pinned) {
                CU_TRY(cudaMemcpyAsync(logits_out + static_cast<size_t>(first) * kClasses, c.logits,
                                       static_cast<size_t>(nb) * kClasses * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
            } else {
                // a device-to-pageable copy would block the host until this pass is through, and with it the
                // enqueueing (and gathering) of the next one: go through pinned staging, hand over later
                if (!c.h_logits[buf]) {
                    CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c.h_logits[buf]), static_cast<size_t>(e.max_batch) * kClasses * sizeof(float), cudaHostAllocPortable));
                    CU_TRY(cudaEventCreateWithFlags(&c.ev_logits[buf], cudaEventDisableTiming));
                }
                VIT_TRY(drain_logits(c, buf, logits_out));
                CU_TRY(cudaMemcpyAsync(c.h_logits[buf], c.logits, static_cast<size_t>(nb) * kClasses * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
                CU_TRY(cudaEventRecord(c.ev_logits[buf], c.stream));
                c.pend_first[buf] = first;
                c.pend_count[buf] = nb;
            }
        }
    }
    for (int g = 0; g < G; ++g) {
        DeviceCtx& c = e.ctx[g];
        CU_TRY(cudaSetDevice(c.device));
        const cudaError_t se = cudaStreamSynchronize(c.stream);
        if (se != cudaSuccess) return watchdog_or_cuda_error(se, "forward");
        for (int buf = 0; buf < 2; ++buf) VIT_TRY(drain_logits(c, buf, logits_out));
        if (!e.attn_exact) {
            bool f = false;
            VIT_TRY(take_attn_range_flag(&f));
            *range_flag |= f;
        }
    }
    return 0;
}

int vit_cuda_info(long long* out, int n) {
    if (!g_eng.up || !out) return set_err(VIT_E_ARG, "engine not initialised");
    const DeviceCtx& c = g_eng.ctx[0];
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, c.device));
    const long long v[11] = {c.sm_count, prop.major, prop.minor, g_eng.max_batch, g_eng.tokens, g_eng.prec,
                             (long long)g_eng.ctx.size(), (long long)(c.ws_bytes >> 20), g_eng.attn_exact ? 1 : 0,
                             g_eng.attn_fallbacks, g_eng.prune_last ? 1 : 0};
    for (int i = 0; i < n && i < 11; ++i) out[i] = v[i];
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

// ------------------------------------------------------------------------------------ single-op entry points
namespace {
struct Scratch {  // RAII device buffers for the op tests
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    int alloc(void** p, size_t bytes, bool zero = false) {
        CU_TRY(cudaMalloc(p, bytes));
        ptrs.push_back(*p);
        if (zero) CU_TRY(cudaMemset(*p, 0, bytes));
        return 0;
    }
    int upload_f32(float** p, const float* h, size_t n) {
        VIT_TRY(alloc(reinterpret_cast<void**>(p), n * 4));
        CU_TRY(cudaMemcpy(*p, h, n * 4, cudaMemcpyHostToDevice));
        return 0;
    }
    // fp32 host -> operand precision device (rows padded with zeros up to pad_elems)
    int upload_operand(void** p, const float* h, size_t n, int prec, size_t pad_elems = 0) {
        float* tmp = nullptr;
        VIT_TRY(upload_f32(&tmp, h, n));
        VIT_TRY(alloc(p, std::max(n, pad_elems) * 2, true));
        VIT_TRY(launch_convert_from_f32(prec, tmp, *p, n, nullptr));
        CU_TRY(cudaDeviceSynchronize());
        return 0;
    }
    int download_operand(float* h, const void* d, size_t n, int prec) {
        float* tmp = nullptr;
        VIT_TRY(alloc(reinterpret_cast<void**>(&tmp), n * 4));
        VIT_TRY(launch_convert_to_f32(prec, d, tmp, n, nullptr));
        CU_TRY(cudaMemcpy(h, tmp, n * 4, cudaMemcpyDeviceToHost));
        return 0;
    }
};
int op_begin(int* sm_count) {
    VIT_TRY(check_device(0, sm_count));
    CU_TRY(cudaSetDevice(0));
    return 0;
}
int op_end(const char* what) {
    const cudaError_t se = cudaDeviceSynchronize();
    if (se != cudaSuccess) return watchdog_or_cuda_error(se, what);
    return 0;
}
}  // namespace

int vit_cuda_op_linear(const float* x, const float* W, const float* b, const float* residual, float* y, int m, int n,
                       int k, int epilogue, int precision) {
    int sms = 0;
    VIT_TRY(op_begin(&sms));
    if (!x || !W || !b || !y || m <= 0) return set_err(VIT_E_ARG, "bad arguments");
    if (epilogue == VIT_EPI_BIAS_RESIDUAL && !residual) return set_err(VIT_E_ARG, "residual epilogue needs a residual");
    Scratch s;
    void *dx, *dw;
    float* db;
    VIT_TRY(s.upload_operand(&dx, x, (size_t)m * k, precision));
    VIT_TRY(s.upload_operand(&dw, W, (size_t)n * k, precision));
    VIT_TRY(s.upload_f32(&db, b, n));
    CUtensorMap ta, tb;
    VIT_TRY(make_tmap(&ta, precision, dx, k, m, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&tb, precision, dw, k, n, GEMM_BK, 128));
    const bool staged = gemm_impl() == 2;
    CUtensorMap tout;
    if (epilogue == VIT_EPI_BIAS_RESIDUAL) {
        float* dy;
        VIT_TRY(s.upload_f32(&dy, residual, (size_t)m * n));
        GemmParams p{m, n, k, db, dy, dy, 0, 0};
        VIT_TRY(make_tmap_f32(&tout, dy, n, m, GEMM_BM));
        if (staged) VIT_TRY(launch_gemm_staged<EPI_BIAS_RESIDUAL>(precision, ta, tb, tout, tout, p, sms, nullptr));
        else VIT_TRY(launch_gemm<EPI_BIAS_RESIDUAL>(precision, ta, tb, p, sms, nullptr));
        VIT_TRY(op_end("op_linear"));
        CU_TRY(cudaMemcpy(y, dy, (size_t)m * n * 4, cudaMemcpyDeviceToHost));
    } else {
        void* dy;
        VIT_TRY(s.alloc(&dy, (size_t)m * n * 2, true));
        GemmParams p{m, n, k, db, dy, nullptr, 0, 0};
        CUtensorMap tout32;
        VIT_TRY(make_tmap(&tout, precision, dy, n, m, GEMM_BK, GEMM_BM));
        VIT_TRY(make_tmap(&tout32, precision, dy, n, m, GEMM_BK, 32));
        if (epilogue == VIT_EPI_BIAS_GELU) {
            if (staged) VIT_TRY(launch_gemm_staged<EPI_BIAS_GELU>(precision, ta, tb, tout, tout32, p, sms, nullptr));
            else VIT_TRY(launch_gemm<EPI_BIAS_GELU>(precision, ta, tb, p, sms, nullptr));
        } else if (epilogue == VIT_EPI_BIAS) {
            if (staged) VIT_TRY(launch_gemm_staged<EPI_BIAS>(precision, ta, tb, tout, tout32, p, sms, nullptr));
            else VIT_TRY(launch_gemm<EPI_BIAS>(precision, ta, tb, p, sms, nullptr));
        } else return set_err(VIT_E_ARG, "unknown epilogue %d", epilogue);
        VIT_TRY(op_end("op_linear"));
        VIT_TRY(s.download_operand(y, dy, (size_t)m * n, precision));
    }
    return op_end("op_linear");
}

int vit_cuda_op_ln_linear(const float* x, const float* ln_w, const float* ln_b, const float* W, const float* b, float* y, int m,
                          int n, int epilogue, int precision) {
    int sms = 0;
    VIT_TRY(op_begin(&sms));
    if (!x || !ln_w || !ln_b || !W || !b || !y || m <= 0 || n <= 0) return set_err(VIT_E_ARG, "bad arguments");
    if (epilogue != VIT_EPI_BIAS && epilogue != VIT_EPI_BIAS_GELU) return set_err(VIT_E_ARG, "unknown epilogue %d", epilogue);
    Scratch s;
    float *dx, *dlw, *dlb, *dW, *db, *dcs, *dcv;
    void *dxc, *dwf, *dy;
    float2* dst;
    const size_t srows = (static_cast<size_t>(m) + 255) / 256 * 256;
    VIT_TRY(s.upload_f32(&dx, x, (size_t)m * kDim));
    VIT_TRY(s.upload_f32(&dlw, ln_w, kDim));
    VIT_TRY(s.upload_f32(&dlb, ln_b, kDim));
    VIT_TRY(s.upload_f32(&dW, W, (size_t)n * kDim));
    VIT_TRY(s.upload_f32(&db, b, n));
    VIT_TRY(s.alloc(&dxc, (size_t)m * kDim * 2, true));
    VIT_TRY(s.alloc(&dwf, (size_t)n * kDim * 2, true));
    VIT_TRY(s.alloc(reinterpret_cast<void**>(&dcs), (size_t)n * 4));
    VIT_TRY(s.alloc(reinterpret_cast<void**>(&dcv), (size_t)n * 4));
    VIT_TRY(s.alloc(reinterpret_cast<void**>(&dst), srows * sizeof(float2), true));
    VIT_TRY(s.alloc(&dy, (size_t)m * n * 2, true));
    VIT_TRY(launch_fold_ln(precision, dW, dlw, dlb, db, dwf, dcs, dcv, n, nullptr));
    VIT_TRY(launch_rowstats_cast(precision, dx, dxc, dst, m, nullptr));
    CUtensorMap ta, tb, tout;
    VIT_TRY(make_tmap(&ta, precision, dxc, kDim, m, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&tb, precision, dwf, kDim, n, GEMM_BK, 128));
    CUtensorMap tout32;
    VIT_TRY(make_tmap(&tout, precision, dy, n, m, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&tout32, precision, dy, n, m, GEMM_BK, 32));
    GemmParams p{m, n, kDim, dcv, dy, nullptr, 0, 0};
    p.colsum = dcs;
    p.stats_in = dst;
    p.stats_parts = 1;
    p.stats_rows = static_cast<int>(srows);
    if (epilogue == VIT_EPI_BIAS_GELU) VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_GELU>(precision, ta, tb, tout, tout32, p, sms, nullptr));
    else VIT_TRY(launch_gemm_staged_ln<EPI_BIAS>(precision, ta, tb, tout, tout32, p, sms, nullptr));
    VIT_TRY(op_end("op_ln_linear"));
    VIT_TRY(s.download_operand(y, dy, (size_t)m * n, precision));
    return op_end("op_ln_linear");
}

int vit_cuda_op_linear_residual_stats(const float* x, const float* W, const float* b, const float* residual, float* y,
                                      float* y_cast, float* row_sum, float* row_sumsq, int m, int k, int precision) {
    int sms = 0;
    VIT_TRY(op_begin(&sms));
    if (!x || !W || !b || !residual || !y || !y_cast || !row_sum || !row_sumsq || m <= 0 || k <= 0) return set_err(VIT_E_ARG, "bad arguments");
    Scratch s;
    void *dx, *dw, *dyc;
    float *db, *dy;
    float2* dst;
    const size_t srows = (static_cast<size_t>(m) + 255) / 256 * 256;
    VIT_TRY(s.upload_operand(&dx, x, (size_t)m * k, precision));
    VIT_TRY(s.upload_operand(&dw, W, (size_t)kDim * k, precision));
    VIT_TRY(s.upload_f32(&db, b, kDim));
    VIT_TRY(s.upload_f32(&dy, residual, (size_t)m * kDim));
    VIT_TRY(s.alloc(&dyc, (size_t)m * kDim * 2, true));
    VIT_TRY(s.alloc(reinterpret_cast<void**>(&dst), 6 * srows * sizeof(float2), true));
    CUtensorMap ta, tb, tout, tcast;
    VIT_TRY(make_tmap(&ta, precision, dx, k, m, GEMM_BK, GEMM_BM));
    VIT_TRY(make_tmap(&tb, precision, dw, k, kDim, GEMM_BK, 128));
    VIT_TRY(make_tmap_f32(&tout, dy, kDim, m, GEMM_BM));
    VIT_TRY(make_tmap(&tcast, precision, dyc, kDim, m, GEMM_BK, GEMM_BM));
    GemmParams p{m, kDim, k, db, dy, dy, 0, 0};
    p.stats_out = dst;
    p.stats_rows = static_cast<int>(srows);
    VIT_TRY(launch_gemm_staged_ln<EPI_BIAS_RESIDUAL>(precision, ta, tb, tout, tcast, p, sms, nullptr));
    VIT_TRY(op_end("op_linear_residual_stats"));
    CU_TRY(cudaMemcpy(y, dy, (size_t)m * kDim * 4, cudaMemcpyDeviceToHost));
    VIT_TRY(s.download_operand(y_cast, dyc, (size_t)m * kDim, precision));
    std::vector<float2> h(6 * srows);
    CU_TRY(cudaMemcpy(h.data(), dst, h.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    for (int r = 0; r < m; ++r) {   // the consumer's summation order (gemm_sm100_staged_kernel, LN consumer)
        float s1 = 0.f, s2 = 0.f;
        for (int q = 0; q < 6; ++q) {
            s1 += h[q * srows + r].x;
            s2 += h[q * srows + r].y;
        }
        row_sum[r] = s1;
        row_sumsq[r] = s2;
    }
    return op_end("op_linear_residual_stats");
}

int vit_cuda_op_layernorm(const float* x, const float* w, const float* b, float* y, int rows, int precision) {
    VIT_TRY(op_begin(nullptr));
    if (!x || !w || !b || !y || rows <= 0) return set_err(VIT_E_ARG, "bad arguments");
    Scratch s;
    float *dx, *dw, *db;
    void* dy;
    VIT_TRY(s.upload_f32(&dx, x, (size_t)rows * kDim));
    VIT_TRY(s.upload_f32(&dw, w, kDim));
    VIT_TRY(s.upload_f32(&db, b, kDim));
    VIT_TRY(s.alloc(&dy, (size_t)rows * kDim * 2));
    VIT_TRY(launch_layernorm(precision, dx, dw, db, dy, rows, nullptr));
    VIT_TRY(op_end("op_layernorm"));
    VIT_TRY(s.download_operand(y, dy, (size_t)rows * kDim, precision));
    return op_end("op_layernorm");
}

static int op_attention_impl(const float* qkv, float* out, int batch, int tokens, int precision,
                             unsigned long long* trace_out, int trace_len) {
    int sms = 0;
    VIT_TRY(op_begin(&sms));
    if (!qkv || batch <= 0 || tokens <= 0) return set_err(VIT_E_ARG, "bad arguments");
    if (tokens > ATTNL_MAX_TOKENS) return set_err(VIT_E_ARG, "attention: tokens=%d > %d not supported", tokens, ATTNL_MAX_TOKENS);
    Scratch s;
    const size_t rows = (size_t)batch * tokens;
    const int kpad = (tokens + 15) / 16 * 16;
    void *dqkv, *dout;
    {   // Q, K in the operand precision; V in bf16 (as the in_proj epilogue stores it)
        float* tmp = nullptr;
        VIT_TRY(s.upload_f32(&tmp, qkv, rows * 3 * kDim));
        VIT_TRY(s.alloc(&dqkv, rows * 3 * kDim * 2, true));
        const int grid = static_cast<int>(std::min<size_t>((rows * 3 * kDim + 255) / 256, 148 * 16));
        if (precision == VIT_PREC_FP16) convert_qkv_from_f32_kernel<__half><<<grid, 256>>>(tmp, static_cast<uint16_t*>(dqkv), rows * 3 * kDim);
        else convert_qkv_from_f32_kernel<__nv_bfloat16><<<grid, 256>>>(tmp, static_cast<uint16_t*>(dqkv), rows * 3 * kDim);
        VIT_TRY(check_launch("convert_qkv"));
        CU_TRY(cudaDeviceSynchronize());
    }
    VIT_TRY(s.alloc(&dout, rows * kDim * 2, true));
    CUtensorMap tq, tkv, tkv32;
    VIT_TRY(make_tmap(&tq, precision, dqkv, 3 * kDim, rows, ATTN_DH, attention_load_box_rows(tokens)));
    VIT_TRY(make_tmap_3d(&tkv, precision, dout, kDim, tokens, batch, 128));
    VIT_TRY(make_tmap_3d(&tkv32, precision, dout, kDim, tokens, batch, 32));
    AttnParams p{batch, tokens, kpad, dout, 0.125f * 1.4426950408889634f, attn_no_pingpong(), nullptr};
    constexpr size_t kTraceLen = static_cast<size_t>(ATTN_TRACE_WARPS) * ATTN_TRACE_ITEMS * ATTN_TRACE_EVENTS;
    if (trace_out) VIT_TRY(s.alloc(reinterpret_cast<void**>(&p.trace), kTraceLen * 8, true));
    const char* ex = getenv("VIT_ATTN_EXACT");
    bool exact = ex && atoi(ex) != 0;
    VIT_TRY(launch_attention(precision, tq, tkv, tkv32, p, sms, nullptr, exact));
    VIT_TRY(op_end("op_attention"));
    if (!exact && !trace_out) {  // same contract as vit_cuda_forward: repeat with the exact softmax when flagged
        bool flagged = false;
        VIT_TRY(take_attn_range_flag(&flagged));
        if (flagged) {
            VIT_TRY(launch_attention(precision, tq, tkv, tkv32, p, sms, nullptr, true));
            VIT_TRY(op_end("op_attention"));
        }
    }
    if (trace_out)
        CU_TRY(cudaMemcpy(trace_out, p.trace, std::min(kTraceLen, static_cast<size_t>(std::max(trace_len, 0))) * 8,
                          cudaMemcpyDeviceToHost));
    if (out) VIT_TRY(s.download_operand(out, dout, rows * kDim, precision));
    return op_end("op_attention");
}

int vit_cuda_op_attention(const float* qkv, float* out, int batch, int tokens, int precision) {
    if (!out) return set_err(VIT_E_ARG, "bad arguments");
    return op_attention_impl(qkv, out, batch, tokens, precision, nullptr, 0);
}

int vit_cuda_debug_attention_trace(const float* qkv, int batch, int tokens, int precision, unsigned long long* trace,
                                   int trace_len) {
    if (!trace || trace_len <= 0) return set_err(VIT_E_ARG, "bad arguments");
    return op_attention_impl(qkv, nullptr, batch, tokens, precision, trace, trace_len);
}

int vit_cuda_op_embed(const float* images, const float* cls, const float* conv_w, const float* conv_b, const float* pos,
                      float* out, int batch, int img_size, int precision) {
    int sms = 0;
    VIT_TRY(op_begin(&sms));
    if (!images || !cls || !conv_w || !conv_b || !pos || !out || batch <= 0 || img_size % kPatch) return set_err(VIT_E_ARG, "bad arguments");
    Scratch s;
    const int g = img_size / kPatch, patches = g * g, tokens = patches + 1;
    const size_t img_elems = (size_t)3 * img_size * img_size;
    float *dimg, *dcls, *dcb, *dpos, *dx;
    void *dw, *dpatch;
    VIT_TRY(s.upload_f32(&dimg, images, batch * img_elems));
    VIT_TRY(s.upload_f32(&dcls, cls, kDim));
    VIT_TRY(s.upload_f32(&dcb, conv_b, kDim));
    VIT_TRY(s.upload_f32(&dpos, pos, (size_t)tokens * kDim));
    VIT_TRY(s.upload_operand(&dw, conv_w, (size_t)kDim * kDim, precision));
    VIT_TRY(s.alloc(&dpatch, (size_t)batch * patches * kDim * 2, true));
    VIT_TRY(s.alloc(reinterpret_cast<void**>(&dx), (size_t)batch * tokens * kDim * 4, true));
    CUtensorMap ta, tb;
    VIT_TRY(make_tmap(&tb, precision, dw, kDim, kDim, GEMM_BK, 128));
    VIT_TRY(launch_patchify(precision, dimg, dpatch, batch, img_size, sms, nullptr));
    if (gemm_impl() == 2) {
        // the forward pass's path: conv_proj as the per-image EMBED kernel in its LayerNorm-producer form,
        // class rows (with their copy and statistics) from cls_rows_ln_kernel
        void* dxc;
        float2* dst;
        const size_t srows = ((size_t)batch * tokens + 255) / 256 * 256;
        VIT_TRY(s.alloc(&dxc, (size_t)batch * tokens * kDim * 2, true));
        VIT_TRY(s.alloc(reinterpret_cast<void**>(&dst), 6 * srows * sizeof(float2), true));
        CUtensorMap tx3, txc3, tpos;
        VIT_TRY(make_tmap_3d(&ta, precision, dpatch, kDim, patches, batch, GEMM_BM));
        VIT_TRY(make_tmap_3d_f32(&tx3, dx, kDim, tokens, batch, GEMM_BM));
        VIT_TRY(make_tmap_3d(&txc3, precision, dxc, kDim, tokens, batch, GEMM_BM));
        VIT_TRY(make_tmap_f32(&tpos, dpos, kDim, tokens, GEMM_BM));
        if (precision == VIT_PREC_FP16)
            cls_rows_ln_kernel<__half><<<(batch + 7) / 8, 256>>>(dx, static_cast<__half*>(dxc), dst, (int)srows, dcls, dpos, batch, tokens);
        else
            cls_rows_ln_kernel<__nv_bfloat16><<<(batch + 7) / 8, 256>>>(dx, static_cast<__nv_bfloat16*>(dxc), dst, (int)srows, dcls, dpos, batch, tokens);
        VIT_TRY(check_launch("cls_rows"));
        GemmParams p{batch * ((patches + 255) / 256) * 256, kDim, kDim, dcb, dx, nullptr, patches, tokens};
        p.stats_out = dst;
        p.stats_rows = (int)srows;
        VIT_TRY(launch_gemm_embed(precision, true, ta, tb, tx3, txc3, tpos, p, sms, nullptr));
    } else {
        VIT_TRY(make_tmap(&ta, precision, dpatch, kDim, (uint64_t)batch * patches, GEMM_BK, GEMM_BM));
        cls_rows_kernel<<<(batch * kDim + 255) / 256, 256>>>(dx, dcls, dpos, batch, tokens);
        VIT_TRY(check_launch("cls_rows"));
        GemmParams p{batch * patches, kDim, kDim, dcb, dx, dpos, patches, tokens};
        VIT_TRY(launch_gemm<EPI_PATCH_EMBED>(precision, ta, tb, p, sms, nullptr));
    }
    VIT_TRY(op_end("op_embed"));
    CU_TRY(cudaMemcpy(out, dx, (size_t)batch * tokens * kDim * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int vit_cuda_op_head(const float* x, const float* ln_w, const float* ln_b, const float* head_w, const float* head_b,
                     float* logits, int batch, int tokens) {
    VIT_TRY(op_begin(nullptr));
    if (!x || !ln_w || !ln_b || !head_w || !head_b || !logits || batch <= 0 || tokens <= 0) return set_err(VIT_E_ARG, "bad arguments");
    Scratch s;
    float *dx, *dlw, *dlb, *dhw, *dhb, *dcls, *dlog;
    VIT_TRY(s.upload_f32(&dx, x, (size_t)batch * tokens * kDim));
    VIT_TRY(s.upload_f32(&dlw, ln_w, kDim));
    VIT_TRY(s.upload_f32(&dlb, ln_b, kDim));
    VIT_TRY(s.upload_f32(&dhw, head_w, (size_t)kClasses * kDim));
    VIT_TRY(s.upload_f32(&dhb, head_b, kClasses));
    VIT_TRY(s.alloc(reinterpret_cast<void**>(&dcls), (size_t)batch * kDim * 4));
    VIT_TRY(s.alloc(reinterpret_cast<void**>(&dlog), (size_t)batch * kClasses * 4));
    head_ln_kernel<<<(batch + 7) / 8, 256>>>(dx, dlw, dlb, dcls, batch, tokens);
    VIT_TRY(check_launch("head_ln"));
    head_gemm_kernel<<<dim3((kClasses + HEAD_CLASSES - 1) / HEAD_CLASSES, std::min((batch + HEAD_IMGS - 1) / HEAD_IMGS, 32)), 256>>>(dcls, dhw, dhb, dlog, batch, kClasses);
    VIT_TRY(check_launch("head_gemm"));
    VIT_TRY(op_end("op_head"));
    CU_TRY(cudaMemcpy(logits, dlog, (size_t)batch * kClasses * 4, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
