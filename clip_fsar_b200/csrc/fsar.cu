// libfsar_sm100.so — host side of the C ABI declared in include/fsar.h.
//
// Owns: the packed weights (fp32 masters + 16-bit tensor-core operands), the activation workspace, the TMA
// tensor maps, the kernel launch sequence of the CLIP ViT frame encoder, the temporal prototype modulator
// and the cosine/OTAM head. There is no CPU path: every entry point needs an sm_100 device.
//
// Reference semantics replaced here (all in /root/reference/models/base/few_shot.py):
//   VisionTransformer.forward 671-688, ResidualAttentionBlock 633-640, Transformer_v1.forward 990-999,
//   CNN_OTAM_CLIPFSAR.forward eval branch 2932-2990.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/fsar.h"
#include "attention_tcgen05.cuh"
#include "gemm_pair_tcgen05.cuh"
#include "gemm_tcgen05.cuh"
#include "head_kernels.cuh"
#include "text_kernels.cuh"
#include "vit_kernels.cuh"
#ifdef FSAR_PROBES
#include "probes_attention_mma.cuh"   // round-1 mma.sync attention, A/B baseline of tools/gemm_probe.py only
#endif

using namespace fsar;

#ifdef FSAR_BF16
typedef __nv_bfloat16 T16;
static const int kOperandDtype = 1;
#else
typedef __half T16;
static const int kOperandDtype = 0;
#endif

namespace {

thread_local std::string g_create_error;

struct Weight {
    std::string name;
    int64_t numel = 0;
    float* d32 = nullptr;  // fp32 master (device); may point into a shared allocation (owns32 == false)
    T16* d16 = nullptr;    // 16-bit [rows, kp] operand for the tensor-core GEMMs, or nullptr
    int rows = 0, cols = 0, kp = 0;
    bool owns32 = true;
    bool set = false;
    bool required = true;  // false: text_features_{train,test} (needed by episodes only) and the optional text tower
    bool tower = false;    // true: a weight of the CLIP text tower ("clip.*", registered by fsar_text_configure)
};

// Workspace of one episode's head (modulator, prototypes, cos/OTAM). One per episode slot of a batched call, so the
// heads of the episodes of one fsar_episodes_* call can run concurrently on side streams.
struct HeadWs {
    float *seq = nullptr, *mod_ln = nullptr, *mod_qkvbuf = nullptr, *mod_att = nullptr, *mod_y = nullptr, *mod_h = nullptr,
          *mod_out = nullptr, *mod_tmp = nullptr, *mod_part = nullptr;
    float *protos = nullptr, *dists = nullptr, *cum = nullptr;
    int *cls = nullptr, *counts = nullptr;
};

// Device pointers of the weights, resolved once (fsar_create / fsar_text_configure): the allocations never move, so the
// launch sequence does no string building or map lookups (~90 per ViT pass before).
struct BlockW {   // one ResidualAttentionBlock (few_shot.py:619-640), frame encoder or text tower
    const T16 *w_in, *w_out, *w_fc, *w_proj;
    const float *b_in, *b_out, *b_fc, *b_proj, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};
struct ModW {     // one Transformer_v1 layer (few_shot.py:979-999)
    const float *norm_g, *norm_b, *qkv, *w_out, *b_out, *w_fc, *b_fc, *w_proj, *b_proj;
};
struct VitW {
    const T16* conv1;
    const float *cls_emb, *pos, *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b, *proj, *scale, *text_train, *text_test;
};
struct TextW {
    const float *tok, *pos, *lnf_g, *lnf_b, *proj;
};

struct ProfRec {
    int cls;
    cudaEvent_t a, b;
    double flops, bytes;
};

struct HostSlot {
    uint8_t* raw_dev = nullptr;    // uint8 staging of the _u8 entry point (grown on demand)
    size_t raw_bytes = 0;
    float* frames_dev = nullptr;   // [max_videos * max_tokens, 3, S, S]: support first, then target
    float* labels_dev = nullptr;   // [2 * max_videos]: support_labels | real_support_labels
    float* logits_dev = nullptr;   // [max_videos * max_videos]
    float* clogits_dev = nullptr;  // [max_videos * max_classes]
    float* logits_pin = nullptr;   // pinned host mirrors
    float* clogits_pin = nullptr;
    cudaEvent_t copied = nullptr, done = nullptr;
    size_t n_logits = 0, n_clogits = 0;
    int busy = 0;
};

}  // namespace

struct fsar_handle {
    fsar_config cfg;
    std::string err;
    int tokens = 0, grid = 0, patch_k = 0, patch_kp = 0, sms = 148;
    std::vector<Weight> w;
    std::unordered_map<std::string, int> widx;
    std::vector<std::string> missing_cache;
    int n_text_train = 0, n_text_test = 0;
    fsar_text_config text_cfg = {0, 0, 0, 0, 0};   // width == 0: no text tower configured
    // fused modulator QKV weights: one [3 * inner, E] allocation per layer
    std::vector<float*> mod_qkv;
    // resolved device pointers (see BlockW)
    VitW vw{};
    TextW tw{};
    std::vector<BlockW> vit_blocks, text_blocks;
    std::vector<ModW> mod_layers;
    // ---- ViT workspace (capacity cfg.max_frames)
    T16 *patches16 = nullptr, *ln16 = nullptr, *qkv16 = nullptr, *att16 = nullptr, *h16 = nullptr;
    T16 *cls_q16 = nullptr, *cls_att16 = nullptr, *cls_ln16 = nullptr, *cls_h16 = nullptr;   // last block: CLS rows only
    float* x32 = nullptr;
    // ---- head workspace
    float *feats = nullptr;  // [max_videos * max_tokens, E] support rows then target rows
    std::vector<HeadWs> ws;                       // [max_batch]
    std::vector<cudaStream_t> head_streams;       // [max_batch] side streams for concurrent heads
    std::vector<cudaEvent_t> head_done;           // [max_batch]
    cudaEvent_t head_fork = nullptr;
    int last_ws = 0;
    // last episode geometry (for fsar_peek)
    int last_S = 0, last_Q = 0, last_T = 0, last_way = 0, last_rows = 0;
    const float *last_sup = nullptr, *last_tgt = nullptr;
    // ---- TMA
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn encode = nullptr;
    std::map<std::tuple<const void*, int, int, int, long long>, CUtensorMap> tmaps;
    std::set<const void*> smem_opt_in;   // kernels whose dynamic smem limit has been raised on this handle's device
    // non-null while a uint8 entry point is enqueueing: the frame pointers of the segments address RAW uint8 [H, W, 3]
    // frames of this geometry, and the patch gather does the resize / crop / normalise itself (8f-1)
    const PreprocParams* u8_src = nullptr;
    // ---- host-buffer path
    HostSlot slot[2];
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    // ---- device-side input validation: kernels that index text_features_test with caller labels clamp
    // the index and raise bits of this flag, which lives in mapped pinned host memory: the host reads it without a
    // synchronisation of its own -- exactly after the event wait in *_collect_host, lazily (at the next call) on the
    // stream-ordered device entry points.
    int* status_host = nullptr;   // [0]: real_support_labels outside [0, n_text_test); [1]: more distinct support_labels
    int* status_dev = nullptr;    //      than `way`
    // ---- instrumentation
    int64_t launches = 0;
    bool profiling = false;
    size_t l2_persist_bytes = 0, l2_window_max = 0;   // L2 set-aside for the residual stream (0 = disabled)
    // Verification knobs of the product library (read once at fsar_create):
    bool cls_last_block = true;     // FSAR_FULL_LAST_BLOCK=1: the last block also computes the token rows nobody reads
                                    // (tests/test_gpu_episode.py proves both give the same frame features)
    bool pdl = true;                // FSAR_NO_PDL=1: no programmatic dependent launch between the frame-encoder kernels
    bool mod_fused = true;          // single-episode calls run a modulator layer as ONE cooperative kernel (probes build:
                                    // FSAR_NO_MOD_FUSED=1 keeps the six launches for A/B)
    int mod_fused_ctas_per_sm = 0;  // co-resident CTAs per SM of modulator_fused_kernel (0: cooperative launch unavailable)
    size_t mod_fused_smem_max = 0;  // dynamic shared memory the kernel was sized (and opted in) for at fsar_create
    // A/B switches that only exist in the -DFSAR_PROBES build (libfsar_sm100_probes.so, tools/gemm_probe.py); in the
    // product library they are compile-time constants and the alternative code paths are not compiled in:
    bool alternate_rows = true;     // FSAR_NO_ALTERNATE=1: every kernel walks rows first-to-last
    bool single_cta_gemm = false;   // FSAR_GEMM_SINGLE=1: one CTA per 128 x 256 tile instead of CTA pairs
    int gemm_debug = 0;             // FSAR_GEMM_DEBUG: bottleneck probes (results are WRONG when set)
    int att_debug = 0;              // FSAR_ATT_DEBUG: same for the attention core
    int small_m = 128;              // FSAR_SMALL_M: GEMMs with at most this many rows use 64-column single-CTA tiles
    bool legacy_attention = false;  // FSAR_LEGACY_ATTENTION=1: round-1 mma.sync attention core
    std::vector<ProfRec> prof;
};

namespace {

int fail(fsar_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

// Every entry point runs on the handle's device and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(const fsar_handle* h);
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
};

#define CU_OK(h, expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(h, FSAR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                        \
    } while (0)

#define RET_IF(expr)          \
    do {                      \
        int r__ = (expr);     \
        if (r__ != 0) return r__; \
    } while (0)

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

DeviceGuard::DeviceGuard(const fsar_handle* h) {
    if (h == nullptr) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != h->cfg.device) changed = (cudaSetDevice(h->cfg.device) == cudaSuccess);
}

// Report (and clear) what the label-checking kernels of EARLIER episodes flagged.
int check_status(fsar_handle* h) {
    if (h->status_host == nullptr) return 0;
    volatile int* f = reinterpret_cast<volatile int*>(h->status_host);
    const int st = (f[0] ? 1 : 0) | (f[1] ? 2 : 0);
    if (st == 0) return 0;
    f[0] = 0;
    f[1] = 0;
    if (st & 1)
        return fail(h, FSAR_E_INVALID, "an episode carried real_support_labels outside [0, %d) (rows of text_features_test); "
                    "its logits are invalid (the reference raises IndexError at few_shot.py:2946)", h->n_text_test);
    return fail(h, FSAR_E_INVALID, "an episode carried more distinct support_labels than its `way`; its logits are invalid");
}

// ---------------------------------------------------------------- profiling helpers
struct Scope {
    fsar_handle* h;
    cudaStream_t st;
    int cls;
    double flops, bytes;
    cudaEvent_t a = nullptr, b = nullptr;
    Scope(fsar_handle* h_, cudaStream_t st_, int cls_, double flops_, double bytes_)
        : h(h_), st(st_), cls(cls_), flops(flops_), bytes(bytes_) {
        if (h->profiling) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    ~Scope() {
        h->launches += 1;
        if (h->profiling) {
            cudaEventRecord(b, st);
            h->prof.push_back(ProfRec{cls, a, b, flops, bytes});
        }
    }
};

int check_launch(fsar_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, FSAR_E_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return 0;
}

// Launch of a frame-encoder kernel with programmatic dependent launch: the kernel may become resident while its
// predecessor in `st` drains (its prologue runs early, its body waits in griddepcontrol.wait, see ptx.cuh).
// FSAR_NO_PDL=1 launches it as an ordinary stream-ordered kernel.
template <typename... KArgs, typename... Args>
void launch_pdl(fsar_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // errors surface in check_launch()
}

// ---------------------------------------------------------------- TMA tensor maps
// Row-major matrix [rows, cols] (cols contiguous) of 16-bit operands (f32 == 0) or fp32 (f32 == 1),
// box = [box_rows, box_cols] with box_cols * elem_size == 128 bytes, 128-byte swizzle.
// `pitch` = distance between rows in elements (0: cols): a pitch of tokens * width over the token matrix selects one row
// per frame, e.g. the CLS rows.
int get_tmap(fsar_handle* h, const void* ptr, int rows, int cols, int box_rows, int box_cols, int f32, CUtensorMap* out,
             long long pitch = 0) {
    if (pitch == 0) pitch = cols;
    auto key = std::make_tuple(ptr, rows, cols, box_rows * 4 + f32 * 2 + (box_cols == 64 ? 1 : 0), pitch);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
        *out = it->second;
        return 0;
    }
    const int esz = f32 ? 4 : 2;
    if ((pitch * esz) % 16 != 0) return fail(h, FSAR_E_INVALID, "TMA needs a 16-byte row pitch, got %lld x %d bytes", pitch, esz);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * esz};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (kOperandDtype ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    CUresult r = h->encode(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FSAR_E_CUDA, "cuTensorMapEncodeTiled failed with %d (rows %d cols %d)", (int)r, rows, cols);
    if (h->tmaps.size() > 4096) h->tmaps.clear();
    h->tmaps[key] = m;
    *out = m;
    return 0;
}

// [n_frames][L][D] view of a 16-bit [n_frames * L, D] matrix, box = [1][32 tokens][64 columns], 128-byte swizzle:
// the attention core stores 32-row tiles through it and tokens >= L of a frame are clipped.
int get_tmap_tokens3d(fsar_handle* h, const void* ptr, int n_frames, int L, int D, CUtensorMap* out) {
    auto key = std::make_tuple(ptr, n_frames * 4096 + L, D, -3, 0LL);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
        *out = it->second;
        return 0;
    }
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)n_frames};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)L * D * 2};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->encode(&m, kOperandDtype ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                           const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FSAR_E_CUDA, "cuTensorMapEncodeTiled (3D) failed with %d (%d x %d x %d)", (int)r, n_frames, L, D);
    if (h->tmaps.size() > 4096) h->tmaps.clear();
    h->tmaps[key] = m;
    *out = m;
    return 0;
}

// Opt a kernel into more than 48 KB of dynamic shared memory, once per handle (the attribute is per device, and two
// handles of one process may sit on different devices).
template <typename Kernel>
int ensure_smem(fsar_handle* h, Kernel kern, int bytes) {
    const void* key = reinterpret_cast<const void*>(kern);
    if (h->smem_opt_in.count(key)) return 0;
    CU_OK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    h->smem_opt_in.insert(key);
    return 0;
}

// ---------------------------------------------------------------- GEMM dispatch
template <int BN, int EPI>
int launch_gemm_inst(fsar_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                     const GemmParams& p, cudaStream_t st) {
    typedef GemmCfg<BN> Cfg;
    auto kern = gemm_tn_tcgen05_kernel<BN, EPI, T16>;
    RET_IF(ensure_smem(h, kern, Cfg::SMEM_BYTES));
    const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM, n_tiles = (p.N + BN - 1) / BN;
    const int tiles = m_tiles * n_tiles;
    const int grid = tiles < h->sms ? tiles : h->sms;
    launch_pdl(h, kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, ta, tb, tc, p);
    return check_launch(h, "gemm_tn_tcgen05_kernel");
}

template <int EPI>
int launch_gemm_pair_inst(fsar_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                          const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_tn_tcgen05_pair_kernel<EPI, T16>;
    RET_IF(ensure_smem(h, kern, GEMM2_SMEM_BYTES));
    const int tiles = ((p.M + 255) / 256) * ((p.N + GEMM2_BN - 1) / GEMM2_BN);
    const int pairs = tiles < h->sms / 2 ? tiles : h->sms / 2;
    launch_pdl(h, kern, dim3(2 * pairs), dim3(GEMM_THREADS), GEMM2_SMEM_BYTES, st, ta, tb, tc, p);   // __cluster_dims__(2, 1, 1)
    return check_launch(h, "gemm_tn_tcgen05_pair_kernel");
}

int launch_gemm_pair(fsar_handle* h, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                     const GemmParams& p, cudaStream_t st) {
    switch (epi) {
        case EPI_STORE16: return launch_gemm_pair_inst<EPI_STORE16>(h, ta, tb, tc, p, st);
        case EPI_QGELU16: return launch_gemm_pair_inst<EPI_QGELU16>(h, ta, tb, tc, p, st);
        case EPI_RESID32: return launch_gemm_pair_inst<EPI_RESID32>(h, ta, tb, tc, p, st);
        case EPI_STORE32: return launch_gemm_pair_inst<EPI_STORE32>(h, ta, tb, tc, p, st);
    }
    return fail(h, FSAR_E_INVALID, "unknown GEMM epilogue %d", epi);
}

template <int BN>
int launch_gemm_bn(fsar_handle* h, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                   const GemmParams& p, cudaStream_t st) {
    switch (epi) {
        case EPI_STORE16: return launch_gemm_inst<BN, EPI_STORE16>(h, ta, tb, tc, p, st);
        case EPI_QGELU16: return launch_gemm_inst<BN, EPI_QGELU16>(h, ta, tb, tc, p, st);
        case EPI_RESID32: return launch_gemm_inst<BN, EPI_RESID32>(h, ta, tb, tc, p, st);
        case EPI_STORE32: return launch_gemm_inst<BN, EPI_STORE32>(h, ta, tb, tc, p, st);
    }
    return fail(h, FSAR_E_INVALID, "unknown GEMM epilogue %d", epi);
}

// out[M,N] (epilogue) = A16[M,K] W16[N,K]^T (+ bias); W rows are K apart; A rows lda apart (0: K), out rows ldc (0: N).
int gemm(fsar_handle* h, int cls, const T16* a, const T16* w, const float* bias, void* out, int M, int N, int K, int epi,
         cudaStream_t st, int reverse = 0, long long lda = 0, long long ldc = 0) {
    if (M <= 0 || N <= 0 || K <= 0 || (N % 8) != 0 || (K % 8) != 0)
        return fail(h, FSAR_E_INVALID, "gemm: unsupported shape M=%d N=%d K=%d (need N %% 8 == 0, K %% 8 == 0)", M, N, K);
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15)
        return fail(h, FSAR_E_INVALID, "gemm: operands must be 16-byte aligned");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = bias; p.reverse = reverse; p.debug = h->gemm_debug;
    // a handful of rows (the CLS rows of the last block): 64-column tiles on single CTAs, so that N / 64 SMs pull the
    // weights instead of N / 256 CTA pairs -- these launches are bound by what one SM can ingest
    const int bn = (M <= h->small_m) ? 64 : ((N >= 256) ? 256 : ((N >= 128) ? 128 : 64));
    const bool out16 = (epi == EPI_STORE16 || epi == EPI_QGELU16);
    CUtensorMap ta, tb, tc;
    const bool pair = (bn == 256) && !h->single_cta_gemm;   // CTA pairs (cta_group::2, 256 x 256 tiles)
    RET_IF(get_tmap(h, a, M, K, GEMM_BM, GEMM_BK, 0, &ta, lda));
    RET_IF(get_tmap(h, w, N, K, pair ? GEMM2_BN / 2 : bn, GEMM_BK, 0, &tb));
    RET_IF(get_tmap(h, out, M, N, 32, out16 ? 64 : 32, out16 ? 0 : 1, &tc, ldc));
    // algorithmic bytes: every unique tensor once (A, W, the output tile; the residual tile is the output itself)
    Scope s(h, st, cls, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K) + (double)M * N * (out16 ? 2.0 : 4.0));
    if (pair) return launch_gemm_pair(h, epi, ta, tb, tc, p, st);
#ifdef FSAR_PROBES
    if (bn == 256) return launch_gemm_bn<256>(h, epi, ta, tb, tc, p, st);   // FSAR_GEMM_SINGLE=1 A/B only
#endif
    if (bn == 128) return launch_gemm_bn<128>(h, epi, ta, tb, tc, p, st);
    return launch_gemm_bn<64>(h, epi, ta, tb, tc, p, st);
}

// ---------------------------------------------------------------- small launch wrappers
__global__ void f32_to_16_kernel(const float* __restrict__ src, T16* __restrict__ dst, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        uint2 w;
        w.x = pack2<T16>(v.x, v.y);
        w.y = pack2<T16>(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + i) = w;
    } else {
        for (; i < n; ++i) dst[i] = T16(src[i]);
    }
}

// weight [rows, cols] fp32 -> [rows, kp] 16-bit, zero padded columns
__global__ void pack_weight_kernel(const float* __restrict__ src, T16* __restrict__ dst, int rows, int cols, int kp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * kp) return;
    const int r = int(i / kp), c = int(i - (long long)r * kp);
    dst[i] = (c < cols) ? T16(src[(size_t)r * cols + c]) : T16(0.f);
}

// in_pitch: distance between input rows in elements (0: D); the output is always dense.
int layernorm(fsar_handle* h, const float* x, void* out, const float* g, const float* b, int rows, int D, bool out16,
              bool embed, int tokens, const float* cls_emb, const float* pos, cudaStream_t st, int cls, int reverse = 0,
              long long in_pitch = 0, T16* out2 = nullptr, const float* g2 = nullptr, const float* b2 = nullptr) {
    if (in_pitch == 0) in_pitch = D;
    if ((D % 128) != 0 || D > 1024) return fail(h, FSAR_E_INVALID, "layernorm: dim %d must be a multiple of 128 and <= 1024", D);
    if (out16 && !embed && in_pitch == D && (D == 512 || D == 768 || D == 1024)) {
        // dense rows -> 16-bit, the CLIP widths: the instance with the width compiled in (vit_kernels.cuh)
        Scope s(h, st, cls, 0.0, (double)rows * D * 6.0);
        T16* o16 = reinterpret_cast<T16*>(out);
        const int grid16 = (rows + 7) / 8;
        if (D == 768) launch_pdl(h, layernorm16_kernel<T16, 6>, dim3(grid16), dim3(256), 0, st, x, o16, g, b, rows, 1e-5f, reverse);
        else if (D == 1024) launch_pdl(h, layernorm16_kernel<T16, 8>, dim3(grid16), dim3(256), 0, st, x, o16, g, b, rows, 1e-5f, reverse);
        else launch_pdl(h, layernorm16_kernel<T16, 4>, dim3(grid16), dim3(256), 0, st, x, o16, g, b, rows, 1e-5f, reverse);
        return check_launch(h, "layernorm16_kernel");
    }
    const int wpb = 8;
    const int grid = (rows + wpb - 1) / wpb;
    Scope s(h, st, cls, 0.0, (double)rows * D * (4.0 + (out16 ? 2.0 : 4.0) + (out2 ? 2.0 : 0.0)));
    const float* np = nullptr;
    T16* n16 = nullptr;
    if (embed)
        launch_pdl(h, layernorm_kernel<T16, false, true>, dim3(grid), dim3(256), 0, st, x, out, g, b, rows, D, 1e-5f, tokens, cls_emb, pos, reverse, in_pitch, out2, g2, b2);
    else if (out16)
        launch_pdl(h, layernorm_kernel<T16, true, false>, dim3(grid), dim3(256), 0, st, x, out, g, b, rows, D, 1e-5f, tokens, np, np, reverse, in_pitch, n16, np, np);
    else
        launch_pdl(h, layernorm_kernel<T16, false, false>, dim3(grid), dim3(256), 0, st, x, out, g, b, rows, D, 1e-5f, tokens, np, np, reverse, in_pitch, n16, np, np);
    return check_launch(h, "layernorm_kernel");
}

int attention(fsar_handle* h, const T16* qkv, int n_frames, int L, int heads, T16* out, cudaStream_t st, int reverse = 0,
              int causal = 0) {
    const int D = heads * ATT_HD;
    const float scale_log2e = 0.125f * 1.4426950408889634f;  // head_dim ** -0.5 * log2(e)
    Scope s(h, st, FSAR_K_ATTENTION, 4.0 * n_frames * heads * (double)L * L * ATT_HD,
            (double)n_frames * L * D * 2.0 * 4.0);
    if (causal && L > ATT5_MAX_KEYS)
        return fail(h, FSAR_E_INVALID, "causal attention supports at most %d tokens, got %d", ATT5_MAX_KEYS, L);
#ifdef FSAR_PROBES
    if (h->legacy_attention && !causal && L <= 272) {
        const dim3 grid((L + ATT_QROWS - 1) / ATT_QROWS, heads, n_frames);
        if (L <= 208) {
            RET_IF(ensure_smem(h, attention_mma_kernel<T16, 13>, att_smem_bytes<13>()));
            launch_pdl(h, attention_mma_kernel<T16, 13>, grid, dim3(128), att_smem_bytes<13>(), st, qkv, out, L, D, scale_log2e);
        } else {
            RET_IF(ensure_smem(h, attention_mma_kernel<T16, 17>, att_smem_bytes<17>()));
            launch_pdl(h, attention_mma_kernel<T16, 17>, grid, dim3(128), att_smem_bytes<17>(), st, qkv, out, L, D, scale_log2e);
        }
        return check_launch(h, "attention_mma_kernel");
    }
#endif
    if (L > ATT5_MAX_TOKENS)
        return fail(h, FSAR_E_INVALID, "attention: %d tokens per frame exceeds the supported %d (ViT-L/14 at 224 x 224)", L,
                    ATT5_MAX_TOKENS);
    // tcgen05 / TMEM path: S and O accumulate in tensor memory, one softmax thread per query row.
    // L <= 208: MAXK = 208 instance (TMA-staged output); 208 < L <= 257: MAXK = 256 instance (ViT-L/14: 256 tokens
    // on the tensor cores + one scalar token).
    const bool big = L > ATT5_MAX_KEYS;
    RET_IF(ensure_smem(h, attention_tcgen05_kernel<T16, false, 208>, Att5Cfg<208>::SMEM_BYTES));
    RET_IF(ensure_smem(h, attention_tcgen05_kernel<T16, true, 208>, Att5Cfg<208>::SMEM_BYTES));
    RET_IF(ensure_smem(h, attention_tcgen05_kernel<T16, false, 256>, Att5Cfg<256>::SMEM_BYTES));
    Att5Params ap{};
    ap.n_frames = n_frames; ap.L = L; ap.heads = heads; ap.D = D;
    ap.Lm = L < 256 ? L : 256; ap.extra = L - ap.Lm;
    ap.LK = round_up(ap.Lm, 16); ap.n_mtiles = (ap.Lm + 127) / 128; ap.scale_log2e = scale_log2e; ap.out = out; ap.reverse = reverse;
    ap.debug = h->att_debug;
    CUtensorMap tq, tkv, tx, to;
    RET_IF(get_tmap(h, qkv, n_frames * L, 3 * D, 128, 64, 0, &tq));
    RET_IF(get_tmap(h, qkv, n_frames * L, 3 * D, ap.LK, 64, 0, &tkv));
    RET_IF(get_tmap(h, qkv, n_frames * L, 3 * D, 8, 64, 0, &tx));
    RET_IF(get_tmap_tokens3d(h, out, n_frames, L, D, &to));
    const int items = n_frames * heads;
    const int grid5 = items < h->sms ? items : h->sms;
    if (causal)
        launch_pdl(h, attention_tcgen05_kernel<T16, true, 208>, dim3(grid5), dim3(ATT5_THREADS), Att5Cfg<208>::SMEM_BYTES, st, tq, tkv, tx, to, ap);
    else if (big)
        launch_pdl(h, attention_tcgen05_kernel<T16, false, 256>, dim3(grid5), dim3(Att5Cfg<256>::THREADS), Att5Cfg<256>::SMEM_BYTES, st, tq, tkv, tx, to, ap);
    else
        launch_pdl(h, attention_tcgen05_kernel<T16, false, 208>, dim3(grid5), dim3(ATT5_THREADS), Att5Cfg<208>::SMEM_BYTES, st, tq, tkv, tx, to, ap);
    return check_launch(h, "attention_tcgen05_kernel");
}

template <int ACT>
int linear_f32(fsar_handle* h, const float* A, const float* W, const float* bias, const float* residual, float* C, int R,
               int N, int K, cudaStream_t st) {
    const dim3 grid((N + LIN_BN - 1) / LIN_BN, (R + LIN_BM - 1) / LIN_BM);
    Scope s(h, st, FSAR_K_MODULATOR, 2.0 * R * N * K, 4.0 * ((double)N * K + (double)R * K + (double)R * N));
    if ((K % (LIN_BK * LIN_WARPS)) != 0) return fail(h, FSAR_E_INVALID, "linear: K=%d must be a multiple of %d", K, LIN_BK * LIN_WARPS);
    launch_pdl(h, linear_f32_kernel<ACT>, dim3(grid), dim3(LIN_THREADS), 0, st, A, W, bias, residual, C, R, N, K);
    return check_launch(h, "linear_f32_kernel");
}

// ---------------------------------------------------------------- weights
Weight* find_w(fsar_handle* h, const std::string& name) {
    auto it = h->widx.find(name);
    return it == h->widx.end() ? nullptr : &h->w[it->second];
}
const float* W32(fsar_handle* h, const std::string& name) { return find_w(h, name)->d32; }
const T16* W16(fsar_handle* h, const std::string& name) { return find_w(h, name)->d16; }

void add_w(fsar_handle* h, const std::string& name, int64_t numel, int rows = 0, int cols = 0, int kp = 0,
           bool required = true) {
    Weight w;
    w.name = name; w.numel = numel; w.rows = rows; w.cols = cols; w.kp = kp; w.required = required;
    h->widx[name] = (int)h->w.size();
    h->w.push_back(w);
}

int alloc_weight_storage(fsar_handle* h);
void resolve_weights(fsar_handle* h);

int alloc_weights(fsar_handle* h) {
    const fsar_config& c = h->cfg;
    const int D = c.width, E = c.embed_dim, P = c.patch_size;
    const int inner = c.mod_heads * c.mod_dim_head, F = c.mod_mlp_dim;
    add_w(h, "backbone.conv1.weight", (int64_t)D * 3 * P * P, D, 3 * P * P, h->patch_kp);
    add_w(h, "backbone.class_embedding", D);
    add_w(h, "backbone.positional_embedding", (int64_t)h->tokens * D);
    add_w(h, "backbone.ln_pre.weight", D);
    add_w(h, "backbone.ln_pre.bias", D);
    add_w(h, "backbone.ln_post.weight", D);
    add_w(h, "backbone.ln_post.bias", D);
    add_w(h, "backbone.proj", (int64_t)D * E);
    for (int i = 0; i < c.layers; ++i) {
        const std::string p = "backbone.transformer.resblocks." + std::to_string(i) + ".";
        add_w(h, p + "attn.in_proj_weight", (int64_t)3 * D * D, 3 * D, D, D);
        add_w(h, p + "attn.in_proj_bias", 3 * D);
        add_w(h, p + "attn.out_proj.weight", (int64_t)D * D, D, D, D);
        add_w(h, p + "attn.out_proj.bias", D);
        add_w(h, p + "ln_1.weight", D);
        add_w(h, p + "ln_1.bias", D);
        add_w(h, p + "ln_2.weight", D);
        add_w(h, p + "ln_2.bias", D);
        add_w(h, p + "mlp.c_fc.weight", (int64_t)4 * D * D, 4 * D, D, D);
        add_w(h, p + "mlp.c_fc.bias", 4 * D);
        add_w(h, p + "mlp.c_proj.weight", (int64_t)4 * D * D, D, 4 * D, 4 * D);
        add_w(h, p + "mlp.c_proj.bias", D);
    }
    add_w(h, "scale", 1);
    for (int l = 0; l < c.mod_depth; ++l) {
        const std::string p = "context2.layers." + std::to_string(l) + ".";
        add_w(h, p + "0.norm.weight", E);
        add_w(h, p + "0.norm.bias", E);
        add_w(h, p + "0.fn.to_q.weight", (int64_t)inner * E);
        add_w(h, p + "0.fn.to_k.weight", (int64_t)inner * E);
        add_w(h, p + "0.fn.to_v.weight", (int64_t)inner * E);
        add_w(h, p + "0.fn.to_out.0.weight", (int64_t)E * inner);
        add_w(h, p + "0.fn.to_out.0.bias", E);
        add_w(h, p + "1.net.0.weight", (int64_t)F * E);
        add_w(h, p + "1.net.0.bias", F);
        add_w(h, p + "1.net.3.weight", (int64_t)E * F);
        add_w(h, p + "1.net.3.bias", E);
    }
    add_w(h, "text_features_train", (int64_t)c.max_classes * E, 0, 0, 0, false);
    add_w(h, "text_features_test", (int64_t)c.max_classes * E, 0, 0, 0, false);

    h->mod_qkv.assign(c.mod_depth, nullptr);
    for (int l = 0; l < c.mod_depth; ++l)
        CU_OK(h, cudaMalloc(&h->mod_qkv[l], sizeof(float) * 3 * (size_t)inner * E));
    RET_IF(alloc_weight_storage(h));
    resolve_weights(h);
    return 0;
}

int alloc_weight_storage(fsar_handle* h) {
    const fsar_config& c = h->cfg;
    const int E = c.embed_dim, inner = c.mod_heads * c.mod_dim_head;
    for (auto& w : h->w) {
        if (w.d32 != nullptr) continue;   // already allocated (fsar_text_configure calls this again)
        // the three modulator projections of a layer live in one [3 * inner, E] buffer (one fused launch)
        size_t pos;
        int which = -1;
        if ((pos = w.name.find(".0.fn.to_q.weight")) != std::string::npos) which = 0;
        else if ((pos = w.name.find(".0.fn.to_k.weight")) != std::string::npos) which = 1;
        else if ((pos = w.name.find(".0.fn.to_v.weight")) != std::string::npos) which = 2;
        if (which >= 0) {
            const int l = atoi(w.name.c_str() + strlen("context2.layers."));
            w.d32 = h->mod_qkv[l] + (size_t)which * inner * E;
            w.owns32 = false;
        } else {
            CU_OK(h, cudaMalloc(&w.d32, sizeof(float) * (size_t)w.numel));
        }
        if (w.rows > 0) CU_OK(h, cudaMalloc(&w.d16, sizeof(T16) * (size_t)w.rows * w.kp));
    }
    return 0;
}

void resolve_blocks(fsar_handle* h, const std::string& root, int layers, std::vector<BlockW>* out) {
    out->resize(layers);
    for (int i = 0; i < layers; ++i) {
        const std::string p = root + "transformer.resblocks." + std::to_string(i) + ".";
        BlockW& b = (*out)[i];
        b.w_in = W16(h, p + "attn.in_proj_weight");   b.b_in = W32(h, p + "attn.in_proj_bias");
        b.w_out = W16(h, p + "attn.out_proj.weight"); b.b_out = W32(h, p + "attn.out_proj.bias");
        b.w_fc = W16(h, p + "mlp.c_fc.weight");       b.b_fc = W32(h, p + "mlp.c_fc.bias");
        b.w_proj = W16(h, p + "mlp.c_proj.weight");   b.b_proj = W32(h, p + "mlp.c_proj.bias");
        b.ln1_g = W32(h, p + "ln_1.weight"); b.ln1_b = W32(h, p + "ln_1.bias");
        b.ln2_g = W32(h, p + "ln_2.weight"); b.ln2_b = W32(h, p + "ln_2.bias");
    }
}

void resolve_weights(fsar_handle* h) {
    const fsar_config& c = h->cfg;
    VitW& v = h->vw;
    v.conv1 = W16(h, "backbone.conv1.weight");
    v.cls_emb = W32(h, "backbone.class_embedding");
    v.pos = W32(h, "backbone.positional_embedding");
    v.ln_pre_g = W32(h, "backbone.ln_pre.weight");   v.ln_pre_b = W32(h, "backbone.ln_pre.bias");
    v.ln_post_g = W32(h, "backbone.ln_post.weight"); v.ln_post_b = W32(h, "backbone.ln_post.bias");
    v.proj = W32(h, "backbone.proj");
    v.scale = W32(h, "scale");
    v.text_train = W32(h, "text_features_train");
    v.text_test = W32(h, "text_features_test");
    resolve_blocks(h, "backbone.", c.layers, &h->vit_blocks);
    h->mod_layers.resize(c.mod_depth);
    for (int l = 0; l < c.mod_depth; ++l) {
        const std::string p = "context2.layers." + std::to_string(l) + ".";
        ModW& m = h->mod_layers[l];
        m.norm_g = W32(h, p + "0.norm.weight"); m.norm_b = W32(h, p + "0.norm.bias");
        m.qkv = h->mod_qkv[l];
        m.w_out = W32(h, p + "0.fn.to_out.0.weight"); m.b_out = W32(h, p + "0.fn.to_out.0.bias");
        m.w_fc = W32(h, p + "1.net.0.weight");        m.b_fc = W32(h, p + "1.net.0.bias");
        m.w_proj = W32(h, p + "1.net.3.weight");      m.b_proj = W32(h, p + "1.net.3.bias");
    }
}

template <typename T>
int dalloc(fsar_handle* h, T** p, size_t n) {
    CU_OK(h, cudaMalloc(p, sizeof(T) * (n ? n : 1)));
    return 0;
}

int alloc_workspace(fsar_handle* h) {
    const fsar_config& c = h->cfg;
    const size_t D = c.width, E = c.embed_dim;
    const size_t M = (size_t)c.max_frames * h->tokens;
    const size_t Mp = (size_t)c.max_frames * h->grid * h->grid;
    RET_IF(dalloc(h, &h->patches16, Mp * h->patch_kp));
    CU_OK(h, cudaMemset(h->patches16, 0, sizeof(T16) * Mp * h->patch_kp));
    RET_IF(dalloc(h, &h->x32, M * D));
    RET_IF(dalloc(h, &h->ln16, M * D));
    RET_IF(dalloc(h, &h->qkv16, M * 3 * D));
    RET_IF(dalloc(h, &h->att16, M * D));
    RET_IF(dalloc(h, &h->h16, M * 4 * D));
    RET_IF(dalloc(h, &h->cls_q16, (size_t)c.max_frames * D));
    RET_IF(dalloc(h, &h->cls_att16, (size_t)c.max_frames * D));
    RET_IF(dalloc(h, &h->cls_ln16, (size_t)c.max_frames * D));
    RET_IF(dalloc(h, &h->cls_h16, (size_t)c.max_frames * 4 * D));
    const size_t V = c.max_videos, T = c.max_tokens;
    const size_t rows = V * (T + 1);
    const size_t inner = (size_t)c.mod_heads * c.mod_dim_head;
    RET_IF(dalloc(h, &h->feats, (size_t)c.max_batch * V * T * E));
    h->ws.resize(c.max_batch);
    for (HeadWs& w : h->ws) {
        RET_IF(dalloc(h, &w.seq, rows * E));
        RET_IF(dalloc(h, &w.mod_ln, rows * E));
        RET_IF(dalloc(h, &w.mod_qkvbuf, rows * 3 * inner));
        RET_IF(dalloc(h, &w.mod_att, rows * inner));
        RET_IF(dalloc(h, &w.mod_y, rows * E));
        RET_IF(dalloc(h, &w.mod_h, rows * (size_t)c.mod_mlp_dim));
        RET_IF(dalloc(h, &w.mod_out, rows * E));
        RET_IF(dalloc(h, &w.mod_tmp, rows * E));
        RET_IF(dalloc(h, &w.mod_part, MODF_KSPLIT * rows * E));
        RET_IF(dalloc(h, &w.protos, V * T * E));
        RET_IF(dalloc(h, &w.dists, V * V * T * T));
        RET_IF(dalloc(h, &w.cum, V * V));
        RET_IF(dalloc(h, &w.cls, V));
        RET_IF(dalloc(h, &w.counts, V));
    }
    if (c.max_batch > 1) {
        h->head_streams.assign(c.max_batch, nullptr);
        h->head_done.assign(c.max_batch, nullptr);
        for (int i = 0; i < c.max_batch; ++i) {
            CU_OK(h, cudaStreamCreateWithFlags(&h->head_streams[i], cudaStreamNonBlocking));
            CU_OK(h, cudaEventCreateWithFlags(&h->head_done[i], cudaEventDisableTiming));
        }
        CU_OK(h, cudaEventCreateWithFlags(&h->head_fork, cudaEventDisableTiming));
    }
    return 0;
}

bool ready(fsar_handle* h, bool need_text) {
    for (auto& w : h->w)
        if (!w.set && !w.tower && (w.required || need_text)) return false;
    return true;
}

// ---------------------------------------------------------------- the ViT frame encoder
// Encodes `n` frames whose patches are already gathered into h->patches16 rows [0, n * G * G).
// The fp32 residual stream (48 MB for an 80-frame pass) is touched four times per layer (LN1, out-proj reduce-add,
// LN2, c_proj reduce-add) with > 100 MB of other traffic in between, so by default it round-trips to HBM every time.
// An L2 access-policy window marks it persisting (and everything else keeps the normal policy) for the kernels of
// this pass; the window is removed again before returning so the caller's stream is left as it was.
void set_l2_window(fsar_handle* h, cudaStream_t st, void* ptr, size_t bytes) {
    if (h->l2_persist_bytes == 0) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (ptr != nullptr && bytes > 0) {
        if (bytes > h->l2_window_max) bytes = h->l2_window_max;
        attr.accessPolicyWindow.base_ptr = ptr;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = bytes <= h->l2_persist_bytes ? 1.0f : float(double(h->l2_persist_bytes) / double(bytes));
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

int vit_encode_gathered(fsar_handle* h, int n, float* feats_out, cudaStream_t st) {
    const fsar_config& c = h->cfg;
    const int D = c.width, L = h->tokens, G2 = h->grid * h->grid;
    const int M = n * L;
    set_l2_window(h, st, h->x32, sizeof(float) * (size_t)M * D);
    // conv1 as a GEMM into a scratch [n * G * G, D] fp32 (the MLP hidden buffer is free at this point) ...
    float* patch32 = reinterpret_cast<float*>(h->h16);
    RET_IF(gemm(h, FSAR_K_GEMM_PATCH, h->patches16, h->vw.conv1, nullptr, patch32, n * G2, D,
                h->patch_kp, EPI_STORE32, st));
    // ... and ln_pre assembles [CLS | patches] + positional embedding on the fly (few_shot.py:675-677)
    // ... and applies ln_1 of the first block to the row while it is still in registers (x32 and ln16 in one pass)
    RET_IF(layernorm(h, patch32, h->x32, h->vw.ln_pre_g, h->vw.ln_pre_b, M, D, false,
                     true, L, h->vw.cls_emb, h->vw.pos, st,
                     FSAR_K_LAYERNORM, 0, 0, h->ln16, h->vit_blocks[0].ln1_g, h->vit_blocks[0].ln1_b));
    // Row direction alternates from kernel to kernel (dir ^= 1): every consumer walks the rows in the opposite order
    // of the producer of its big input (x32 58 MB, qkv16 87 MB, h16 116 MB at 96 frames — together more than the
    // 126 MB L2), so under LRU it starts on the rows that were written last and are still cache-resident.
    int dir = h->alternate_rows ? 1 : 0;
    const int flip = h->alternate_rows ? 1 : 0;
    for (int i = 0; i < c.layers; ++i) {
        const BlockW& bw = h->vit_blocks[i];
        if (i > 0) {   // block 0's ln_1 came out of the ln_pre kernel above
            RET_IF(layernorm(h, h->x32, h->ln16, bw.ln1_g, bw.ln1_b, M, D, true, false, L,
                             nullptr, nullptr, st, FSAR_K_LAYERNORM, dir));
        }
        dir ^= flip;
        if (i == c.layers - 1 && h->cls_last_block && L <= CLS_ATT_MAX_L) {
            // Last block: only x[:, 0, :] survives the transformer (ln_post(x[:, 0, :]) @ proj, few_shot.py:683-686), so
            // all tokens feed K and V, but Q, the attention output, out_proj, ln_2 and the MLP are evaluated for the
            // CLS row of every frame only -- the same numbers the full block would leave in those rows. Row pitches of
            // L * D address the CLS rows of ln16 / x32 in place.
            const long long cls_pitch = (long long)L * D;
            const T16* w_in = bw.w_in;
            const float* b_in = bw.b_in;
            T16* kv16 = h->qkv16;   // [M, 2 D]
            RET_IF(gemm(h, FSAR_K_GEMM_QKV, h->ln16, w_in + (size_t)D * D, b_in + D, kv16, M, 2 * D, D, EPI_STORE16, st, dir));
            RET_IF(gemm(h, FSAR_K_LAST_BLOCK_CLS, h->ln16, w_in, b_in, h->cls_q16, n, D, D, EPI_STORE16, st, 0, cls_pitch));
            {
                Scope s(h, st, FSAR_K_LAST_BLOCK_CLS, 4.0 * n * c.heads * (double)L * ATT_HD, (double)M * 2 * D * 2.0);
                launch_pdl(h, cls_attention_kernel<T16>, dim3(c.heads, n), dim3(128), 0, st, (const T16*)h->cls_q16,
                           (const T16*)kv16, h->cls_att16, L, D, 0.125f * 1.4426950408889634f);
                RET_IF(check_launch(h, "cls_attention_kernel"));
            }
            RET_IF(gemm(h, FSAR_K_LAST_BLOCK_CLS, h->cls_att16, bw.w_out, bw.b_out, h->x32, n, D, D, EPI_RESID32, st, 0, 0, cls_pitch));
            RET_IF(layernorm(h, h->x32, h->cls_ln16, bw.ln2_g, bw.ln2_b, n, D, true, false,
                             L, nullptr, nullptr, st, FSAR_K_LAST_BLOCK_CLS, 0, cls_pitch));
            RET_IF(gemm(h, FSAR_K_LAST_BLOCK_CLS, h->cls_ln16, bw.w_fc, bw.b_fc,
                        h->cls_h16, n, 4 * D, D, EPI_QGELU16, st));
            RET_IF(gemm(h, FSAR_K_LAST_BLOCK_CLS, h->cls_h16, bw.w_proj, bw.b_proj,
                        h->x32, n, D, 4 * D, EPI_RESID32, st, 0, 0, cls_pitch));
            break;
        }
        RET_IF(gemm(h, FSAR_K_GEMM_QKV, h->ln16, bw.w_in, bw.b_in,
                    h->qkv16, M, 3 * D, D, EPI_STORE16, st, dir));
        dir ^= flip;
        RET_IF(attention(h, h->qkv16, n, L, c.heads, h->att16, st, dir));
        dir ^= flip;
        RET_IF(gemm(h, FSAR_K_GEMM_OUT, h->att16, bw.w_out, bw.b_out,
                    h->x32, M, D, D, EPI_RESID32, st, dir));
        dir ^= flip;
        RET_IF(layernorm(h, h->x32, h->ln16, bw.ln2_g, bw.ln2_b, M, D, true, false, L,
                         nullptr, nullptr, st, FSAR_K_LAYERNORM, dir));
        dir ^= flip;
        RET_IF(gemm(h, FSAR_K_GEMM_FC1, h->ln16, bw.w_fc, bw.b_fc, h->h16, M,
                    4 * D, D, EPI_QGELU16, st, dir));
        dir ^= flip;
        RET_IF(gemm(h, FSAR_K_GEMM_FC2, h->h16, bw.w_proj, bw.b_proj, h->x32,
                    M, D, 4 * D, EPI_RESID32, st, dir));
        dir ^= flip;
    }
    {
        const dim3 grid((n + FINAL_FPC - 1) / FINAL_FPC, (c.embed_dim + FINAL_COLS - 1) / FINAL_COLS);
        Scope s(h, st, FSAR_K_FINAL_PROJ, 2.0 * n * D * c.embed_dim, 4.0 * ((double)D * c.embed_dim + (double)n * D));
        launch_pdl(h, final_proj_kernel, grid, dim3(256), sizeof(float) * FINAL_FPC * D, st,
                   (const float*)h->x32, h->vw.ln_post_g, h->vw.ln_post_b, h->vw.proj,
                   feats_out, n, L, D, c.embed_dim, 1e-5f);
        RET_IF(check_launch(h, "final_proj_kernel"));
    }
    set_l2_window(h, st, nullptr, 0);
    return 0;
}

int patch_gather(fsar_handle* h, const float* frames, int n, int row_frame_offset, cudaStream_t st) {
    const fsar_config& c = h->cfg;
    const int S = c.image_size, P = c.patch_size, G = h->grid;
    if (h->u8_src != nullptr) {
        PreprocParams pp = *h->u8_src;
        pp.n = n;
        const long long total = (long long)n * S * (S / 2);
        Scope s(h, st, FSAR_K_PATCH_GATHER, 0.0, (double)n * (3.0 * pp.H * pp.W + 6.0 * S * S));
        launch_pdl(h, preprocess_patches_u8_kernel<T16>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st,
                   reinterpret_cast<const uint8_t*>(frames), h->patches16 + (size_t)row_frame_offset * G * G * h->patch_kp, pp,
                   P, h->patch_kp);
        return check_launch(h, "preprocess_patches_u8_kernel");
    }
    const long long total = (long long)n * 3 * S * G;
    const int grid = (int)((total + 255) / 256);
    Scope s(h, st, FSAR_K_PATCH_GATHER, 0.0, (double)n * 3 * S * S * 6.0);
    launch_pdl(h, patch_gather_kernel<T16>, dim3(grid), dim3(256), 0, st, frames,
               h->patches16 + (size_t)row_frame_offset * G * G * h->patch_kp, n, S, P, h->patch_kp);
    return check_launch(h, "patch_gather_kernel");
}

// Encode the frames of a list of device buffers (segment i holds counts[i] frames) into consecutive feats rows, in
// passes of at most cfg.max_frames frames that ignore segment (video set / episode) boundaries.
int vit_encode_segments(fsar_handle* h, const float* const* ptrs, const int* counts, int nseg, float* feats,
                        cudaStream_t st) {
    const fsar_config& c = h->cfg;
    // distance between consecutive frames of a segment in units of float (the pointer type): fp32 crops, or raw uint8
    // frames addressed through the same pointers (u8_src; H * W * 3 bytes is a multiple of 4 only by luck, so step in bytes)
    const size_t frame_bytes = h->u8_src ? (size_t)h->u8_src->H * h->u8_src->W * 3 : sizeof(float) * 3 * c.image_size * c.image_size;
    int total = 0;
    for (int i = 0; i < nseg; ++i) total += counts[i];
    int seg = 0, seg_off = 0, done = 0;
    while (done < total) {
        const int n = (total - done < c.max_frames) ? total - done : c.max_frames;
        int filled = 0;
        while (filled < n) {
            while (seg_off == counts[seg]) { ++seg; seg_off = 0; }
            const int avail = counts[seg] - seg_off;
            const int take = (avail < n - filled) ? avail : n - filled;
            RET_IF(patch_gather(h, reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(ptrs[seg]) + (size_t)seg_off * frame_bytes),
                                take, filled, st));
            filled += take;
            seg_off += take;
        }
        RET_IF(vit_encode_gathered(h, n, feats + (size_t)done * c.embed_dim, st));
        done += n;
    }
    return 0;
}

// ---------------------------------------------------------------- temporal prototype modulator
// x [rows, E]: n_q sequences of T tokens followed by n_s sequences of T + 1 tokens -> out [rows, E]
size_t mod_fused_smem(const fsar_config& c, int T) {
    const size_t nmax = (size_t)T + 1;
    const size_t att = sizeof(float) * (3 * nmax * c.mod_dim_head + nmax * (nmax + 1));
    return att > sizeof(LinSmem) ? att : sizeof(LinSmem);
}

// `fused`: one cooperative launch per layer (modulator_fused_kernel) instead of six. Used when the call holds ONE episode
// (the nn.Module path, fsar_modulate): with several episodes per call their heads run concurrently on side streams, where
// six small launches per head overlap better than cooperative grids that each want every SM.
int modulate_rows(fsar_handle* h, HeadWs& w, const float* x, int n_q, int n_s, int T, float* out, cudaStream_t st,
                  bool fused = false) {
    const fsar_config& c = h->cfg;
    const int E = c.embed_dim, inner = c.mod_heads * c.mod_dim_head, F = c.mod_mlp_dim;
    const int rows = n_q * T + n_s * (T + 1);
    if (T + 1 > MOD_MAX_TOK) return fail(h, FSAR_E_INVALID, "modulator: %d tokens per sequence exceeds %d", T + 1, MOD_MAX_TOK);
    const float* cur = x;
    fused = fused && h->mod_fused && h->mod_fused_ctas_per_sm > 0 && mod_fused_smem(c, T) <= h->mod_fused_smem_max &&
            (F % (MODF_KSPLIT * LIN_BK * LIN_WARPS)) == 0 &&
            (E % (LIN_BK * LIN_WARPS)) == 0 && (inner % (LIN_BK * LIN_WARPS)) == 0;
    for (int l = 0; l < c.mod_depth; ++l) {
        const ModW& mw = h->mod_layers[l];
        // depth > 1: intermediate layers ping-pong between mod_tmp and seq (seq is dead once layer 0 consumed it)
        float* dst = (l == c.mod_depth - 1) ? out : ((l & 1) ? w.seq : w.mod_tmp);
        if (fused) {
            ModFusedParams mp{};
            mp.x = cur; mp.out = dst;
            mp.n_q = n_q; mp.n_s = n_s; mp.T = T; mp.E = E; mp.inner = inner; mp.F = F; mp.heads = c.mod_heads; mp.dh = c.mod_dim_head;
            mp.scale = 1.0f / sqrtf((float)c.mod_dim_head); mp.eps = 1e-5f;
            mp.norm_g = mw.norm_g; mp.norm_b = mw.norm_b; mp.w_qkv = mw.qkv; mp.w_out = mw.w_out; mp.b_out = mw.b_out;
            mp.w_fc = mw.w_fc; mp.b_fc = mw.b_fc; mp.w_proj = mw.w_proj; mp.b_proj = mw.b_proj;
            mp.ln = w.mod_ln; mp.qkv = w.mod_qkvbuf; mp.att = w.mod_att; mp.y = w.mod_y; mp.hid = w.mod_h; mp.part = w.mod_part;
            const size_t smem = mod_fused_smem(c, T);
            const int per_sm = h->mod_fused_ctas_per_sm < 3 ? h->mod_fused_ctas_per_sm : 3;   // measured: 1 -> 105, 2 -> 101, 3 -> 92, 4 -> 114 us
            Scope s(h, st, FSAR_K_MODULATOR, 2.0 * rows * ((double)3 * inner * E + (double)E * inner + 2.0 * E * F) + 4.0 * rows * (T + 1) * inner,
                    4.0 * ((double)3 * inner * E + (double)E * inner + 2.0 * E * F));
            void* args[] = {&mp};
            cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(modulator_fused_kernel), dim3(per_sm * h->sms),
                                                        dim3(LIN_THREADS), args, smem, st);
            if (e != cudaSuccess) return fail(h, FSAR_E_CUDA, "cooperative launch of modulator_fused_kernel failed: %s", cudaGetErrorString(e));
            cur = dst;
            continue;
        }
        RET_IF(layernorm(h, cur, w.mod_ln, mw.norm_g, mw.norm_b, rows, E, false, false,
                         1, nullptr, nullptr, st, FSAR_K_MODULATOR));
        RET_IF(linear_f32<LIN_NONE>(h, w.mod_ln, mw.qkv, nullptr, nullptr, w.mod_qkvbuf, rows, 3 * inner, E, st));
        {
            const int nmax = T + 1, dh = c.mod_dim_head;
            const size_t smem = sizeof(float) * ((size_t)3 * nmax * dh + (size_t)nmax * (nmax + 1));
            Scope s(h, st, FSAR_K_MODULATOR, 4.0 * rows * nmax * inner, 4.0 * 4.0 * rows * inner);
            launch_pdl(h, modulator_attention_kernel, dim3(n_q + n_s, c.mod_heads), dim3(128), smem, st, 
                w.mod_qkvbuf, w.mod_qkvbuf + inner, w.mod_qkvbuf + 2 * inner, w.mod_att, n_q, T, 3 * inner, inner, dh,
                1.0f / sqrtf((float)dh));
            RET_IF(check_launch(h, "modulator_attention_kernel"));
        }
        RET_IF(linear_f32<LIN_NONE>(h, w.mod_att, mw.w_out, mw.b_out, cur,
                                    w.mod_y, rows, E, inner, st));
        RET_IF(linear_f32<LIN_GELU>(h, w.mod_y, mw.w_fc, mw.b_fc, nullptr, w.mod_h,
                                    rows, F, E, st));
        RET_IF(linear_f32<LIN_NONE>(h, w.mod_h, mw.w_proj, mw.b_proj, w.mod_y, dst,
                                    rows, E, F, st));
        cur = dst;
    }
    return 0;
}

int otam_logits(fsar_handle* h, const float* q, const float* protos, int Q, int way, int T, int single_direct,
                float* logits, float* dists, float* cum, cudaStream_t st) {
    if (T > OTAM_MAX_T || T < 1) return fail(h, FSAR_E_INVALID, "otam: T=%d outside [1, %d]", T, OTAM_MAX_T);
    Scope s(h, st, FSAR_K_COS_OTAM, 2.0 * Q * way * T * T * h->cfg.embed_dim,
            4.0 * ((double)(Q + way) * T * h->cfg.embed_dim + (double)Q * way));
    launch_pdl(h, cos_otam_kernel, dim3(Q, way), dim3(256), 0, st, q, protos, T, h->cfg.embed_dim, way, h->cfg.otam_lambda, single_direct,
                                                  logits, dists, cum);
    return check_launch(h, "cos_otam_kernel");
}

// Everything after the frame encoder: frame features of one episode (support rows, then target rows) -> logits,
// class_logits.
int head_forward(fsar_handle* h, HeadWs& w, const float* sup, const float* tgt, const float* support_labels,
                 const float* real_support_labels, int S, int Q, int T, int way, int merge_before, int single_direct,
                 int text_mode, float text_coff, float* logits, float* class_logits, cudaStream_t st, bool fused_modulator) {
    const fsar_config& c = h->cfg;
    const int E = c.embed_dim;
    if (text_mode < 0 || text_mode > 2) return fail(h, FSAR_E_INVALID, "episode: text_mode %d not in {0, 1, 2}", text_mode);
    if (text_mode != 0) {
        if (way > TEXT_MAX_WAY) return fail(h, FSAR_E_INVALID, "text branches support at most %d classes", TEXT_MAX_WAY);
        class_logits = nullptr;   // the reference returns class_logits = None in these branches (few_shot.py:2852, 2930)
    }
    {
        Scope s(h, st, FSAR_K_HEAD_MISC, 0.0, 8.0 * S);
        launch_pdl(h, class_index_kernel, dim3(1), dim3(128), 0, st, support_labels, S, w.cls, w.counts, way,
                   real_support_labels, h->n_text_test, h->status_dev);
        RET_IF(check_launch(h, "class_index_kernel"));
    }
    if (text_mode == 1) {   // TRAIN.EVAL_TEXT: text probabilities only, the modulator / OTAM are not evaluated
        Scope s(h, st, FSAR_K_HEAD_MISC, 2.0 * Q * way * E, 4.0 * ((double)Q * T * E + (double)way * E));
        launch_pdl(h, text_fusion_kernel, dim3(Q), dim3(256), sizeof(float) * E, st, tgt, h->vw.text_test, real_support_labels, w.cls,
                                                              w.counts, S, T, E, way, h->n_text_test, h->vw.scale, 1, 0.f, nullptr, logits);
        RET_IF(check_launch(h, "text_fusion_kernel"));
        h->last_S = S; h->last_Q = Q; h->last_T = T; h->last_way = way; h->last_rows = 0;
        h->last_sup = sup; h->last_tgt = tgt;
        return 0;
    }
    if (class_logits != nullptr) {
        if (!find_w(h, "text_features_train")->set) return fail(h, FSAR_E_STATE, "text_features_train has not been set");
        Scope s(h, st, FSAR_K_HEAD_MISC, 2.0 * (S + Q) * h->n_text_train * E,
                4.0 * ((double)(S + Q) * T * E + (double)h->n_text_train * E));
        launch_pdl(h, class_text_logits_kernel, dim3(S + Q), dim3(256), sizeof(float) * E, st, sup, S, tgt, Q, T, E, h->vw.text_train,
                                                                        h->n_text_train, h->vw.scale, class_logits);
        RET_IF(check_launch(h, "class_text_logits_kernel"));
    }
    const int n_sup_seq = merge_before ? way : S;
    const int rows = Q * T + n_sup_seq * (T + 1);
    {
        Scope s(h, st, FSAR_K_HEAD_MISC, 0.0, 8.0 * rows * E);
        launch_pdl(h, build_sequences_kernel, dim3(rows), dim3(128), 0, st, sup, tgt, h->vw.text_test, real_support_labels, w.cls,
                                                     w.counts, S, Q, T, E, way, merge_before, h->n_text_test, w.seq);
        RET_IF(check_launch(h, "build_sequences_kernel"));
    }
    // depth > 1 uses w.seq as a ping-pong buffer, so the first layer must not read it after layer 2 wrote it:
    // layer l reads `cur` and writes dst != cur, and w.seq is only overwritten at l = 1 (after l = 0 consumed it).
    RET_IF(modulate_rows(h, w, w.seq, Q, n_sup_seq, T, w.mod_out, st, fused_modulator));
    {
        Scope s(h, st, FSAR_K_HEAD_MISC, 0.0, 8.0 * way * T * E);
        launch_pdl(h, prototype_kernel, dim3(way * T), dim3(128), 0, st, w.mod_out, Q * T, n_sup_seq, T, E, w.cls, w.counts, merge_before,
                                                  w.protos);
        RET_IF(check_launch(h, "prototype_kernel"));
    }
    RET_IF(otam_logits(h, w.mod_out, w.protos, Q, way, T, single_direct, logits, w.dists, w.cum, st));
    if (text_mode == 2) {   // TRAIN.COMBINE: geometric fusion of text and visual probabilities overwrites the logits
        Scope s(h, st, FSAR_K_HEAD_MISC, 2.0 * Q * way * E, 4.0 * ((double)Q * T * E + (double)way * E));
        launch_pdl(h, text_fusion_kernel, dim3(Q), dim3(256), sizeof(float) * E, st, tgt, h->vw.text_test, real_support_labels, w.cls,
                                                              w.counts, S, T, E, way, h->n_text_test, h->vw.scale, 2, text_coff, w.cum,
                                                              logits);
        RET_IF(check_launch(h, "text_fusion_kernel"));
    }
    h->last_S = S; h->last_Q = Q; h->last_T = T; h->last_way = way; h->last_rows = rows;
    h->last_sup = sup; h->last_tgt = tgt;
    return 0;
}

int check_episode(fsar_handle* h, const fsar_episode* ep) {
    const fsar_config& c = h->cfg;
    if (ep == nullptr) return fail(h, FSAR_E_INVALID, "episode is NULL");
    if (ep->n_support < 1 || ep->n_target < 1 || ep->n_frames < 1 || ep->way < 1 || ep->way > ep->n_support)
        return fail(h, FSAR_E_INVALID, "episode: bad geometry S=%d Q=%d T=%d way=%d", ep->n_support, ep->n_target,
                    ep->n_frames, ep->way);
    if (ep->n_support + ep->n_target > c.max_videos || ep->n_frames > c.max_tokens)
        return fail(h, FSAR_E_STATE, "episode exceeds capacity: %d videos (max %d), %d frames (max %d)",
                    ep->n_support + ep->n_target, c.max_videos, ep->n_frames, c.max_tokens);
    if (!ready(h, true)) return fail(h, FSAR_E_STATE, "%d weights have not been set", fsar_missing_weights(h));
    return 0;
}

int check_batch(fsar_handle* h, const fsar_episode* eps, int n) {
    if (eps == nullptr || n < 1) return fail(h, FSAR_E_INVALID, "episodes: NULL array or n_episodes < 1");
    if (n > h->cfg.max_batch) return fail(h, FSAR_E_STATE, "%d episodes exceed max_batch %d", n, h->cfg.max_batch);
    for (int i = 0; i < n; ++i) RET_IF(check_episode(h, &eps[i]));
    return 0;
}

// n episodes with DEVICE pointers: one frame-encoder sweep over all their frames, then the head per episode.
// logits: concatenated [n_target_i * way_i]; class_logits (or NULL): concatenated [(S_i + Q_i) * n_train].
int episodes_forward_dev(fsar_handle* h, const fsar_episode* eps, int n, float* logits, float* class_logits,
                         cudaStream_t st) {
    const int E = h->cfg.embed_dim;
    std::vector<const float*> ptrs;
    std::vector<int> counts;
    for (int i = 0; i < n; ++i) {
        ptrs.push_back(eps[i].support_frames); counts.push_back(eps[i].n_support * eps[i].n_frames);
        ptrs.push_back(eps[i].target_frames);  counts.push_back(eps[i].n_target * eps[i].n_frames);
    }
    RET_IF(vit_encode_segments(h, ptrs.data(), counts.data(), (int)ptrs.size(), h->feats, st));
    // heads: tiny latency-bound kernels (~15 launches per episode). With several episodes in the call they run
    // concurrently, one side stream per episode (fork after the encoder sweep, join before returning to `st`).
    const bool fork = n > 1 && !h->head_streams.empty() && !h->profiling;
    if (fork) CU_OK(h, cudaEventRecord(h->head_fork, st));
    size_t f_off = 0, l_off = 0, c_off = 0;
    for (int i = 0; i < n; ++i) {
        const fsar_episode& ep = eps[i];
        const int T = ep.n_frames;
        const float* sup = h->feats + f_off;
        const float* tgt = sup + (size_t)ep.n_support * T * E;
        cudaStream_t hs = fork ? h->head_streams[i] : st;
        if (fork) CU_OK(h, cudaStreamWaitEvent(hs, h->head_fork, 0));
        if (class_logits != nullptr && ep.text_mode != 0 && h->n_text_train > 0)   // the reference returns None there:
            CU_OK(h, cudaMemsetAsync(class_logits + c_off, 0,                     // the slice reads as zeros, never as stale memory
                                     sizeof(float) * (size_t)(ep.n_support + ep.n_target) * h->n_text_train, hs));
        RET_IF(head_forward(h, h->ws[i], sup, tgt, ep.support_labels, ep.real_support_labels, ep.n_support, ep.n_target, T,
                            ep.way, ep.merge_before, ep.single_direct, ep.text_mode, ep.text_coff, logits + l_off,
                            class_logits ? class_logits + c_off : nullptr, hs, /*fused_modulator=*/n == 1));
        if (fork) {
            CU_OK(h, cudaEventRecord(h->head_done[i], hs));
            CU_OK(h, cudaStreamWaitEvent(st, h->head_done[i], 0));
        }
        h->last_ws = i;
        f_off += (size_t)(ep.n_support + ep.n_target) * T * E;
        l_off += (size_t)ep.n_target * ep.way;
        c_off += (size_t)(ep.n_support + ep.n_target) * h->n_text_train;
    }
    return 0;
}

int ensure_host_path(fsar_handle* h) {
    if (h->copy_stream != nullptr) return 0;
    const fsar_config& c = h->cfg;
    const size_t frame_elems = (size_t)3 * c.image_size * c.image_size;
    const size_t V = c.max_videos, B = c.max_batch;
    CU_OK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CU_OK(h, cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        HostSlot& s = h->slot[i];
        RET_IF(dalloc(h, &s.frames_dev, B * V * c.max_tokens * frame_elems));
        RET_IF(dalloc(h, &s.labels_dev, B * 2 * V));
        RET_IF(dalloc(h, &s.logits_dev, B * V * V));
        RET_IF(dalloc(h, &s.clogits_dev, B * V * (size_t)c.max_classes));
        CU_OK(h, cudaMallocHost(&s.logits_pin, sizeof(float) * B * V * V));
        CU_OK(h, cudaMallocHost(&s.clogits_pin, sizeof(float) * B * V * c.max_classes));
        CU_OK(h, cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        CU_OK(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    return 0;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int fsar_version(void) { return FSAR_VERSION; }

const char* fsar_class_name(int k) {
    static const char* names[FSAR_PROF_CLASSES] = {"patch_gather", "gemm_patch", "layernorm", "gemm_qkv", "attention",
                                                   "gemm_out", "gemm_fc1", "gemm_fc2", "final_proj", "head_misc",
                                                   "modulator", "cos_otam", "last_block_cls"};
    return (k >= 0 && k < FSAR_PROF_CLASSES) ? names[k] : "?";
}

int fsar_operand_dtype(void) { return kOperandDtype; }

const char* fsar_last_error(const fsar_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int fsar_create(const fsar_config* cfg, fsar_handle** out) {
    if (cfg == nullptr || out == nullptr) return fail(nullptr, FSAR_E_INVALID, "fsar_create: NULL argument");
    *out = nullptr;
    const fsar_config& c = *cfg;
    if (c.patch_size <= 0 || c.image_size % c.patch_size != 0 || (c.patch_size & 1))
        return fail(nullptr, FSAR_E_INVALID, "image_size %d must be a multiple of an even patch_size %d", c.image_size, c.patch_size);
    if (c.width % 128 != 0 || c.width > 1024 || c.heads * 64 != c.width)
        return fail(nullptr, FSAR_E_INVALID, "width %d must be a multiple of 128, <= 1024 and equal heads * 64 (heads %d)", c.width, c.heads);
    if (c.embed_dim % 128 != 0 || c.embed_dim > 1024 || c.mod_heads * c.mod_dim_head <= 0 || c.mod_depth < 1 || c.layers < 1)
        return fail(nullptr, FSAR_E_INVALID, "unsupported head geometry (embed_dim %d, mod heads %d x %d, depth %d)", c.embed_dim, c.mod_heads, c.mod_dim_head, c.mod_depth);
    if (c.max_batch < 1 || c.max_batch > 64)
        return fail(nullptr, FSAR_E_INVALID, "max_batch %d must be in [1, 64]", c.max_batch);
    if (c.max_frames < 1 || c.max_videos < 2 || c.max_tokens < 1 || c.max_tokens > OTAM_MAX_T || c.max_classes < 1)
        return fail(nullptr, FSAR_E_INVALID, "bad capacities (max_frames %d, max_videos %d, max_tokens %d <= %d, max_classes %d)", c.max_frames, c.max_videos, c.max_tokens, OTAM_MAX_T, c.max_classes);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= c.device || c.device < 0) {
        cudaGetLastError();
        return fail(nullptr, FSAR_E_CUDA, "no CUDA device %d (found %d); libfsar_sm100 has no CPU path", c.device, ndev);
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c.device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, FSAR_E_CUDA, "device %d is sm_%d%d; libfsar_sm100 only runs on sm_100 (B200)", c.device, prop.major, prop.minor);
    fsar_handle* h = new fsar_handle();
    h->cfg = c;
    h->sms = prop.multiProcessorCount;
    h->grid = c.image_size / c.patch_size;
    h->tokens = h->grid * h->grid + 1;
    h->patch_k = 3 * c.patch_size * c.patch_size;
    h->patch_kp = round_up(h->patch_k, GEMM_BK);
    {
        const char* e = getenv("FSAR_FULL_LAST_BLOCK");
        h->cls_last_block = !(e != nullptr && e[0] == '1');
        e = getenv("FSAR_NO_PDL");
        h->pdl = !(e != nullptr && e[0] == '1');
#ifdef FSAR_PROBES
        e = getenv("FSAR_NO_MOD_FUSED");
        h->mod_fused = !(e != nullptr && e[0] == '1');
        e = getenv("FSAR_LEGACY_ATTENTION");
        h->legacy_attention = (e != nullptr && e[0] == '1');
        e = getenv("FSAR_NO_ALTERNATE");
        h->alternate_rows = !(e != nullptr && e[0] == '1');
        e = getenv("FSAR_GEMM_SINGLE");
        h->single_cta_gemm = (e != nullptr && e[0] == '1');
        e = getenv("FSAR_GEMM_DEBUG");
        h->gemm_debug = e != nullptr ? atoi(e) : 0;
        e = getenv("FSAR_ATT_DEBUG");
        h->att_debug = e != nullptr ? atoi(e) : 0;
        e = getenv("FSAR_SMALL_M");
        if (e != nullptr) h->small_m = atoi(e);
#endif
    }
    int rc = 0;
    DeviceGuard guard(h);
    do {
        if (cudaSetDevice(c.device) != cudaSuccess) { rc = fail(nullptr, FSAR_E_CUDA, "cudaSetDevice(%d) failed", c.device); break; }
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
            rc = fail(nullptr, FSAR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
            break;
        }
        h->encode = reinterpret_cast<fsar_handle::EncodeFn>(fn);
        {   // L2 set-aside for the residual stream (FSAR_NO_L2_PERSIST=1 disables it)
            const char* e = getenv("FSAR_NO_L2_PERSIST");
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c.device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c.device);
            if (!(e != nullptr && e[0] == '1') && max_persist > 0 && max_window > 0) {
                size_t want = sizeof(float) * (size_t)c.max_frames * h->tokens * c.width;
                if (want > (size_t)max_persist) want = (size_t)max_persist;
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                    h->l2_persist_bytes = want;
                    h->l2_window_max = (size_t)max_window;
                } else {
                    cudaGetLastError();
                }
            }
        }
        if ((rc = alloc_weights(h)) != 0) break;
        if ((rc = alloc_workspace(h)) != 0) break;
        {   // the fused modulator layer needs a cooperative launch: how many of its CTAs are co-resident per SM?
            int coop = 0, per_sm = 0;
            const size_t smem = mod_fused_smem(c, c.max_tokens);
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c.device);
            if (coop && (smem <= 48 * 1024 || cudaFuncSetAttribute(modulator_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                   (int)smem) == cudaSuccess) &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, modulator_fused_kernel, LIN_THREADS, smem) == cudaSuccess) {
                h->mod_fused_ctas_per_sm = per_sm;
                h->mod_fused_smem_max = smem;
            }
            else
                cudaGetLastError();
        }
        if (cudaHostAlloc(reinterpret_cast<void**>(&h->status_host), 64, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->status_dev), h->status_host, 0) != cudaSuccess) {
            rc = fail(nullptr, FSAR_E_NOMEM, "cannot allocate the mapped status word");
            break;
        }
        memset(h->status_host, 0, 64);
    } while (0);
    if (rc != 0) {
        if (!h->err.empty()) g_create_error = h->err;
        fsar_destroy(h);
        return rc;
    }
    *out = h;
    return 0;
}

void fsar_destroy(fsar_handle* h) {
    if (h == nullptr) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (auto& w : h->w) {
        if (w.owns32 && w.d32) cudaFree(w.d32);
        if (w.d16) cudaFree(w.d16);
    }
    for (float* p : h->mod_qkv) if (p) cudaFree(p);
    void* bufs[] = {h->patches16, h->ln16, h->qkv16, h->att16, h->h16, h->x32, h->feats,
                    h->cls_q16, h->cls_att16, h->cls_ln16, h->cls_h16};
    for (void* p : bufs) if (p) cudaFree(p);
    for (HeadWs& w : h->ws) {
        void* wb[] = {w.seq, w.mod_ln, w.mod_qkvbuf, w.mod_att, w.mod_y, w.mod_h, w.mod_out, w.mod_tmp, w.mod_part, w.protos, w.dists,
                      w.cum, w.cls, w.counts};
        for (void* p : wb) if (p) cudaFree(p);
    }
    for (cudaStream_t st : h->head_streams) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t ev : h->head_done) if (ev) cudaEventDestroy(ev);
    if (h->head_fork) cudaEventDestroy(h->head_fork);
    for (int i = 0; i < 2; ++i) {
        HostSlot& s = h->slot[i];
        if (s.frames_dev) cudaFree(s.frames_dev);
        if (s.raw_dev) cudaFree(s.raw_dev);
        if (s.labels_dev) cudaFree(s.labels_dev);
        if (s.logits_dev) cudaFree(s.logits_dev);
        if (s.clogits_dev) cudaFree(s.clogits_dev);
        if (s.logits_pin) cudaFreeHost(s.logits_pin);
        if (s.clogits_pin) cudaFreeHost(s.clogits_pin);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.done) cudaEventDestroy(s.done);
    }
    if (h->status_host) cudaFreeHost(h->status_host);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
    for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    delete h;
    if (prev >= 0) cudaSetDevice(prev);
}

int fsar_set_weight(fsar_handle* h, const char* name, const float* data, int64_t numel, int on_device) {
    if (h == nullptr || name == nullptr || data == nullptr) return fail(h, FSAR_E_INVALID, "fsar_set_weight: NULL argument");
    Weight* w = find_w(h, name);
    if (w == nullptr) return fail(h, FSAR_E_NAME, "unknown weight '%s'", name);
    const int E = h->cfg.embed_dim;
    const bool is_text = !w->required && !w->tower;   // text_features_{train,test}: any number of rows up to max_classes
    if (is_text) {
        if (numel <= 0 || numel % E != 0 || numel > w->numel)
            return fail(h, FSAR_E_INVALID, "%s: numel %lld must be a multiple of embed_dim %d and <= %lld", name,
                        (long long)numel, E, (long long)w->numel);
        if (w->name == "text_features_train") h->n_text_train = (int)(numel / E); else h->n_text_test = (int)(numel / E);
    } else if (numel != w->numel) {
        return fail(h, FSAR_E_INVALID, "%s: expected %lld elements, got %lld", name, (long long)w->numel, (long long)numel);
    }
    DeviceGuard guard(h);
    // Ordering contract (fsar.h): the copy runs on the legacy default stream, which does not order against the library's
    // non-blocking compute / copy / head streams nor against a caller's own streams. So: wait for everything in flight
    // on the device (no forward may still be reading the old weights), copy + re-pack, and wait again (the next
    // forward on any stream sees the new ones). Weight updates are rare; forwards never synchronise.
    CU_OK(h, cudaDeviceSynchronize());
    CU_OK(h, cudaMemcpy(w->d32, data, sizeof(float) * (size_t)numel, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    if (w->d16 != nullptr) {
        const long long n = (long long)w->rows * w->kp;
        pack_weight_kernel<<<(int)((n + 255) / 256), 256>>>(w->d32, w->d16, w->rows, w->cols, w->kp);
        RET_IF(check_launch(h, "pack_weight_kernel"));
    }
    CU_OK(h, cudaDeviceSynchronize());
    w->set = true;
    return 0;
}

int fsar_missing_weights(const fsar_handle* h) {
    if (h == nullptr) return -1;
    int n = 0;
    for (auto& w : h->w) n += w.set ? 0 : 1;
    return n;
}

const char* fsar_missing_weight(const fsar_handle* h, int i) {
    if (h == nullptr) return nullptr;
    for (auto& w : h->w)
        if (!w.set && i-- == 0) return w.name.c_str();
    return nullptr;
}

int fsar_vit_forward(fsar_handle* h, const float* frames_dev, int n_frames, float* feats_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || frames_dev == nullptr || feats_dev == nullptr || n_frames < 1)
        return fail(h, FSAR_E_INVALID, "fsar_vit_forward: bad argument");
    if (!ready(h, false)) return fail(h, FSAR_E_STATE, "%d weights have not been set (first: %s)", fsar_missing_weights(h), fsar_missing_weight(h, 0));
    return vit_encode_segments(h, &frames_dev, &n_frames, 1, feats_dev, (cudaStream_t)stream);
}

int fsar_modulate(fsar_handle* h, const float* x_dev, int n_seq, int n_tok, float* out_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || x_dev == nullptr || out_dev == nullptr || n_seq < 1 || n_tok < 1)
        return fail(h, FSAR_E_INVALID, "fsar_modulate: bad argument");
    if (!ready(h, false)) return fail(h, FSAR_E_STATE, "%d weights have not been set", fsar_missing_weights(h));
    if ((size_t)n_seq * n_tok > (size_t)h->cfg.max_videos * (h->cfg.max_tokens + 1))
        return fail(h, FSAR_E_STATE, "fsar_modulate: %d x %d rows exceed the workspace", n_seq, n_tok);
    // all sequences have n_tok tokens: express as n_seq "query" sequences of T = n_tok
    return modulate_rows(h, h->ws[0], x_dev, n_seq, 0, n_tok, out_dev, (cudaStream_t)stream, /*fused=*/true);
}

int fsar_otam_logits(fsar_handle* h, const float* q_dev, const float* protos_dev, int Q, int way, int T, int single_direct,
                     float* logits_dev, float* dists_dev, float* cum_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || q_dev == nullptr || protos_dev == nullptr || logits_dev == nullptr || Q < 1 || way < 1)
        return fail(h, FSAR_E_INVALID, "fsar_otam_logits: bad argument");
    return otam_logits(h, q_dev, protos_dev, Q, way, T, single_direct, logits_dev, dists_dev, cum_dev, (cudaStream_t)stream);
}

int fsar_episodes_forward(fsar_handle* h, const fsar_episode* eps, int n_episodes, float* logits_dev,
                          float* class_logits_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || logits_dev == nullptr) return fail(h, FSAR_E_INVALID, "fsar_episodes_forward: NULL argument");
    RET_IF(check_status(h));   // labels of EARLIER stream-ordered episodes (this call does not synchronise either)
    RET_IF(check_batch(h, eps, n_episodes));
    return episodes_forward_dev(h, eps, n_episodes, logits_dev, class_logits_dev, (cudaStream_t)stream);
}

int fsar_episode_forward(fsar_handle* h, const fsar_episode* ep, float* logits_dev, float* class_logits_dev, void* stream) {
    return fsar_episodes_forward(h, ep, 1, logits_dev, class_logits_dev, stream);
}

int fsar_episodes_submit_host(fsar_handle* h, int slot, const fsar_episode* eps, int n_episodes) {
    DeviceGuard guard(h);
    if (h == nullptr || slot < 0 || slot > 1) return fail(h, FSAR_E_INVALID, "fsar_episodes_submit_host: bad argument");
    RET_IF(check_batch(h, eps, n_episodes));
    RET_IF(ensure_host_path(h));
    HostSlot& s = h->slot[slot];
    if (s.busy) return fail(h, FSAR_E_STATE, "slot %d still holds uncollected episodes", slot);
    const fsar_config& c = h->cfg;
    const size_t frame_elems = (size_t)3 * c.image_size * c.image_size;
    // the slot's buffers are free: its previous batch was collected (done event waited in collect)
    std::vector<fsar_episode> dev(eps, eps + n_episodes);
    size_t f_off = 0, n_logits = 0, n_clogits = 0;
    for (int i = 0; i < n_episodes; ++i) {
        const fsar_episode& ep = eps[i];
        const size_t ns = (size_t)ep.n_support * ep.n_frames * frame_elems, nt = (size_t)ep.n_target * ep.n_frames * frame_elems;
        float* lab = s.labels_dev + (size_t)i * 2 * c.max_videos;
        CU_OK(h, cudaMemcpyAsync(s.frames_dev + f_off, ep.support_frames, sizeof(float) * ns, cudaMemcpyHostToDevice, h->copy_stream));
        CU_OK(h, cudaMemcpyAsync(s.frames_dev + f_off + ns, ep.target_frames, sizeof(float) * nt, cudaMemcpyHostToDevice, h->copy_stream));
        CU_OK(h, cudaMemcpyAsync(lab, ep.support_labels, sizeof(float) * ep.n_support, cudaMemcpyHostToDevice, h->copy_stream));
        CU_OK(h, cudaMemcpyAsync(lab + c.max_videos, ep.real_support_labels, sizeof(float) * ep.n_support, cudaMemcpyHostToDevice, h->copy_stream));
        dev[i].support_frames = s.frames_dev + f_off;
        dev[i].target_frames = s.frames_dev + f_off + ns;
        dev[i].support_labels = lab;
        dev[i].real_support_labels = lab + c.max_videos;
        f_off += ns + nt;
        n_logits += (size_t)ep.n_target * ep.way;
        n_clogits += (size_t)(ep.n_support + ep.n_target) * h->n_text_train;
    }
    CU_OK(h, cudaEventRecord(s.copied, h->copy_stream));
    CU_OK(h, cudaStreamWaitEvent(h->compute_stream, s.copied, 0));
    RET_IF(episodes_forward_dev(h, dev.data(), n_episodes, s.logits_dev, s.clogits_dev, h->compute_stream));
    CU_OK(h, cudaMemcpyAsync(s.logits_pin, s.logits_dev, sizeof(float) * n_logits, cudaMemcpyDeviceToHost, h->compute_stream));
    CU_OK(h, cudaMemcpyAsync(s.clogits_pin, s.clogits_dev, sizeof(float) * n_clogits, cudaMemcpyDeviceToHost, h->compute_stream));
    CU_OK(h, cudaEventRecord(s.done, h->compute_stream));
    s.n_logits = n_logits; s.n_clogits = n_clogits; s.busy = 1;
    return 0;
}

int fsar_episodes_collect_host(fsar_handle* h, int slot, float* logits_host, float* class_logits_host) {
    DeviceGuard guard(h);
    if (h == nullptr || slot < 0 || slot > 1) return fail(h, FSAR_E_INVALID, "fsar_episodes_collect_host: bad argument");
    HostSlot& s = h->slot[slot];
    if (!s.busy) return fail(h, FSAR_E_STATE, "slot %d holds no submitted episode", slot);
    CU_OK(h, cudaEventSynchronize(s.done));
    s.busy = 0;
    RET_IF(check_status(h));   // the slot's kernels have run: a bad label of THIS batch is reported here
    if (logits_host) memcpy(logits_host, s.logits_pin, sizeof(float) * s.n_logits);
    if (class_logits_host) memcpy(class_logits_host, s.clogits_pin, sizeof(float) * s.n_clogits);
    return 0;
}

int fsar_episode_submit_host(fsar_handle* h, int slot, const fsar_episode* ep) { return fsar_episodes_submit_host(h, slot, ep, 1); }

int fsar_episode_collect_host(fsar_handle* h, int slot, float* logits_host, float* class_logits_host) {
    return fsar_episodes_collect_host(h, slot, logits_host, class_logits_host);
}

int fsar_episode_forward_host(fsar_handle* h, const fsar_episode* ep, float* logits_host, float* class_logits_host) {
    RET_IF(fsar_episodes_submit_host(h, 0, ep, 1));
    return fsar_episodes_collect_host(h, 0, logits_host, class_logits_host);
}

static int preproc_params(fsar_handle* h, int n, int H, int W, int rh, int rw, const float* mean, const float* std,
                          PreprocParams* out) {
    const int S = h->cfg.image_size;
    if (n < 1 || H < 1 || W < 1 || rh < S || rw < S || mean == nullptr || std == nullptr)
        return fail(h, FSAR_E_INVALID, "preprocess: bad geometry (n %d, source %dx%d, resize %dx%d, crop %d)", n, H, W, rh, rw, S);
    PreprocParams pp;
    pp.n = n; pp.H = H; pp.W = W; pp.RH = rh; pp.RW = rw; pp.S = S;
    for (int c = 0; c < 3; ++c) { pp.mean[c] = mean[c]; pp.std[c] = std[c]; }
    *out = pp;
    return 0;
}

static int preprocess_u8(fsar_handle* h, const uint8_t* src, int n, int H, int W, int rh, int rw, const float* mean,
                         const float* std, float* dst, cudaStream_t st) {
    const int S = h->cfg.image_size;
    PreprocParams pp;
    RET_IF(preproc_params(h, n, H, W, rh, rw, mean, std, &pp));
    const long long total = (long long)n * S * S;
    Scope s(h, st, FSAR_K_PATCH_GATHER, 0.0, (double)n * (3.0 * H * W + 12.0 * S * S));
    preprocess_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, dst, pp);
    return check_launch(h, "preprocess_u8_kernel");
}

int fsar_preprocess_u8(fsar_handle* h, const uint8_t* frames_u8_dev, int n_frames, int H, int W, int resize_h, int resize_w,
                       const float mean[3], const float std[3], float* out_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || frames_u8_dev == nullptr || out_dev == nullptr) return fail(h, FSAR_E_INVALID, "fsar_preprocess_u8: NULL argument");
    return preprocess_u8(h, frames_u8_dev, n_frames, H, W, resize_h, resize_w, mean, std, out_dev, (cudaStream_t)stream);
}

int fsar_vit_forward_u8(fsar_handle* h, const uint8_t* frames_u8_dev, int n_frames, int H, int W, int resize_h, int resize_w,
                        const float mean[3], const float std[3], float* feats_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || frames_u8_dev == nullptr || feats_dev == nullptr || n_frames < 1)
        return fail(h, FSAR_E_INVALID, "fsar_vit_forward_u8: bad argument");
    if (!ready(h, false)) return fail(h, FSAR_E_STATE, "%d weights have not been set (first: %s)", fsar_missing_weights(h), fsar_missing_weight(h, 0));
    PreprocParams pp;
    RET_IF(preproc_params(h, n_frames, H, W, resize_h, resize_w, mean, std, &pp));
    const float* ptr = reinterpret_cast<const float*>(frames_u8_dev);
    h->u8_src = &pp;
    const int rc = vit_encode_segments(h, &ptr, &n_frames, 1, feats_dev, (cudaStream_t)stream);
    h->u8_src = nullptr;
    return rc;
}

int fsar_episodes_submit_host_u8(fsar_handle* h, int slot, const fsar_episode* eps, int n_episodes, int H, int W,
                                 int resize_h, int resize_w, const float mean[3], const float std[3]) {
    DeviceGuard guard(h);
    if (h == nullptr || slot < 0 || slot > 1) return fail(h, FSAR_E_INVALID, "fsar_episodes_submit_host_u8: bad argument");
    RET_IF(check_batch(h, eps, n_episodes));
    RET_IF(ensure_host_path(h));
    HostSlot& s = h->slot[slot];
    if (s.busy) return fail(h, FSAR_E_STATE, "slot %d still holds uncollected episodes", slot);
    const fsar_config& c = h->cfg;
    const size_t raw_frame = (size_t)H * W * 3;
    PreprocParams pp;
    RET_IF(preproc_params(h, 1, H, W, resize_h, resize_w, mean, std, &pp));
    size_t total_frames = 0;
    for (int i = 0; i < n_episodes; ++i) total_frames += (size_t)(eps[i].n_support + eps[i].n_target) * eps[i].n_frames;
    if (total_frames * raw_frame > s.raw_bytes) {
        if (s.raw_dev) { CU_OK(h, cudaStreamSynchronize(h->compute_stream)); cudaFree(s.raw_dev); s.raw_dev = nullptr; }
        CU_OK(h, cudaMalloc(&s.raw_dev, total_frames * raw_frame));
        s.raw_bytes = total_frames * raw_frame;
    }
    std::vector<fsar_episode> dev(eps, eps + n_episodes);
    size_t f_off = 0, n_logits = 0, n_clogits = 0;   // f_off in frames
    for (int i = 0; i < n_episodes; ++i) {
        const fsar_episode& ep = eps[i];
        const size_t ns = (size_t)ep.n_support * ep.n_frames, nt = (size_t)ep.n_target * ep.n_frames;
        float* lab = s.labels_dev + (size_t)i * 2 * c.max_videos;
        CU_OK(h, cudaMemcpyAsync(s.raw_dev + f_off * raw_frame, ep.support_frames, ns * raw_frame, cudaMemcpyHostToDevice, h->copy_stream));
        CU_OK(h, cudaMemcpyAsync(s.raw_dev + (f_off + ns) * raw_frame, ep.target_frames, nt * raw_frame, cudaMemcpyHostToDevice, h->copy_stream));
        CU_OK(h, cudaMemcpyAsync(lab, ep.support_labels, sizeof(float) * ep.n_support, cudaMemcpyHostToDevice, h->copy_stream));
        CU_OK(h, cudaMemcpyAsync(lab + c.max_videos, ep.real_support_labels, sizeof(float) * ep.n_support, cudaMemcpyHostToDevice, h->copy_stream));
        // the episode's frame pointers address the RAW bytes: the patch gather resizes / crops / normalises (u8_src)
        dev[i].support_frames = reinterpret_cast<const float*>(s.raw_dev + f_off * raw_frame);
        dev[i].target_frames = reinterpret_cast<const float*>(s.raw_dev + (f_off + ns) * raw_frame);
        dev[i].support_labels = lab;
        dev[i].real_support_labels = lab + c.max_videos;
        f_off += ns + nt;
        n_logits += (size_t)ep.n_target * ep.way;
        n_clogits += (size_t)(ep.n_support + ep.n_target) * h->n_text_train;
    }
    CU_OK(h, cudaEventRecord(s.copied, h->copy_stream));
    CU_OK(h, cudaStreamWaitEvent(h->compute_stream, s.copied, 0));
    h->u8_src = &pp;
    const int rc = episodes_forward_dev(h, dev.data(), n_episodes, s.logits_dev, s.clogits_dev, h->compute_stream);
    h->u8_src = nullptr;
    RET_IF(rc);
    CU_OK(h, cudaMemcpyAsync(s.logits_pin, s.logits_dev, sizeof(float) * n_logits, cudaMemcpyDeviceToHost, h->compute_stream));
    CU_OK(h, cudaMemcpyAsync(s.clogits_pin, s.clogits_dev, sizeof(float) * n_clogits, cudaMemcpyDeviceToHost, h->compute_stream));
    CU_OK(h, cudaEventRecord(s.done, h->compute_stream));
    s.n_logits = n_logits; s.n_clogits = n_clogits; s.busy = 1;
    return 0;
}

int fsar_text_configure(fsar_handle* h, const fsar_text_config* tc) {
    DeviceGuard guard(h);
    if (h == nullptr || tc == nullptr) return fail(h, FSAR_E_INVALID, "fsar_text_configure: NULL argument");
    if (h->text_cfg.width != 0) return fail(h, FSAR_E_STATE, "the text tower is already configured");
    if (tc->width % 128 != 0 || tc->width > 1024 || tc->heads * ATT_HD != tc->width || tc->layers < 1 ||
        tc->context_length < 1 || tc->context_length > ATT5_MAX_KEYS || tc->vocab_size < 1)
        return fail(h, FSAR_E_INVALID, "unsupported text tower (width %d, heads %d, layers %d, context %d, vocab %d)",
                    tc->width, tc->heads, tc->layers, tc->context_length, tc->vocab_size);
    const int W = tc->width, E = h->cfg.embed_dim;
    const size_t first = h->w.size();
    add_w(h, "clip.token_embedding.weight", (int64_t)tc->vocab_size * W, 0, 0, 0, false);
    add_w(h, "clip.positional_embedding", (int64_t)tc->context_length * W, 0, 0, 0, false);
    add_w(h, "clip.ln_final.weight", W, 0, 0, 0, false);
    add_w(h, "clip.ln_final.bias", W, 0, 0, 0, false);
    add_w(h, "clip.text_projection", (int64_t)W * E, 0, 0, 0, false);
    for (int i = 0; i < tc->layers; ++i) {
        const std::string p = "clip.transformer.resblocks." + std::to_string(i) + ".";
        add_w(h, p + "attn.in_proj_weight", (int64_t)3 * W * W, 3 * W, W, W, false);
        add_w(h, p + "attn.in_proj_bias", 3 * W, 0, 0, 0, false);
        add_w(h, p + "attn.out_proj.weight", (int64_t)W * W, W, W, W, false);
        add_w(h, p + "attn.out_proj.bias", W, 0, 0, 0, false);
        add_w(h, p + "ln_1.weight", W, 0, 0, 0, false);
        add_w(h, p + "ln_1.bias", W, 0, 0, 0, false);
        add_w(h, p + "ln_2.weight", W, 0, 0, 0, false);
        add_w(h, p + "ln_2.bias", W, 0, 0, 0, false);
        add_w(h, p + "mlp.c_fc.weight", (int64_t)4 * W * W, 4 * W, W, W, false);
        add_w(h, p + "mlp.c_fc.bias", 4 * W, 0, 0, 0, false);
        add_w(h, p + "mlp.c_proj.weight", (int64_t)4 * W * W, W, 4 * W, 4 * W, false);
        add_w(h, p + "mlp.c_proj.bias", W, 0, 0, 0, false);
    }
    for (size_t i = first; i < h->w.size(); ++i) h->w[i].tower = true;
    RET_IF(alloc_weight_storage(h));
    h->text_cfg = *tc;
    resolve_blocks(h, "clip.", tc->layers, &h->text_blocks);
    h->tw.tok = W32(h, "clip.token_embedding.weight");
    h->tw.pos = W32(h, "clip.positional_embedding");
    h->tw.lnf_g = W32(h, "clip.ln_final.weight");
    h->tw.lnf_b = W32(h, "clip.ln_final.bias");
    h->tw.proj = W32(h, "clip.text_projection");
    return 0;
}

int fsar_text_encode(fsar_handle* h, const int32_t* tokens_dev, int n_texts, float* out_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || tokens_dev == nullptr || out_dev == nullptr || n_texts < 1)
        return fail(h, FSAR_E_INVALID, "fsar_text_encode: bad argument");
    const fsar_text_config& t = h->text_cfg;
    if (t.width == 0) return fail(h, FSAR_E_STATE, "fsar_text_encode: call fsar_text_configure first");
    for (auto& w : h->w)
        if (w.tower && !w.set) return fail(h, FSAR_E_STATE, "text tower weight '%s' has not been set", w.name.c_str());
    cudaStream_t st = (cudaStream_t)stream;
    const int W = t.width, C = t.context_length, E = h->cfg.embed_dim;
    // the blocks run in the frame encoder's workspace (x32 / ln16 / qkv16 / att16 / h16 hold max_frames * tokens rows of
    // `width` columns); texts are encoded in chunks that fit
    const size_t cap_rows = (size_t)h->cfg.max_frames * h->tokens * h->cfg.width / W;
    const int chunk = (int)(cap_rows / C);
    if (chunk < 1) return fail(h, FSAR_E_STATE, "workspace too small for one text of %d tokens (raise max_frames)", C);
    for (int done = 0; done < n_texts; done += chunk) {
        const int n = n_texts - done < chunk ? n_texts - done : chunk;
        const int M = n * C;
        const int* tok = tokens_dev + (size_t)done * C;
        {
            Scope s(h, st, FSAR_K_HEAD_MISC, 0.0, 12.0 * M * W);
            text_embed_kernel<<<(M + 7) / 8, 256, 0, st>>>(tok, h->tw.tok, h->tw.pos, h->x32, M, C, W, t.vocab_size);
            RET_IF(check_launch(h, "text_embed_kernel"));
        }
        for (int i = 0; i < t.layers; ++i) {
            const BlockW& bw = h->text_blocks[i];
            RET_IF(layernorm(h, h->x32, h->ln16, bw.ln1_g, bw.ln1_b, M, W, true, false, C,
                             nullptr, nullptr, st, FSAR_K_LAYERNORM));
            RET_IF(gemm(h, FSAR_K_GEMM_QKV, h->ln16, bw.w_in, bw.b_in,
                        h->qkv16, M, 3 * W, W, EPI_STORE16, st));
            RET_IF(attention(h, h->qkv16, n, C, t.heads, h->att16, st, 0, /*causal=*/1));
            RET_IF(gemm(h, FSAR_K_GEMM_OUT, h->att16, bw.w_out, bw.b_out,
                        h->x32, M, W, W, EPI_RESID32, st));
            RET_IF(layernorm(h, h->x32, h->ln16, bw.ln2_g, bw.ln2_b, M, W, true, false, C,
                             nullptr, nullptr, st, FSAR_K_LAYERNORM));
            RET_IF(gemm(h, FSAR_K_GEMM_FC1, h->ln16, bw.w_fc, bw.b_fc, h->h16, M,
                        4 * W, W, EPI_QGELU16, st));
            RET_IF(gemm(h, FSAR_K_GEMM_FC2, h->h16, bw.w_proj, bw.b_proj, h->x32,
                        M, W, 4 * W, EPI_RESID32, st));
        }
        {
            Scope s(h, st, FSAR_K_FINAL_PROJ, 2.0 * n * W * E, 4.0 * ((double)W * E + (double)n * W));
            text_final_kernel<<<n, 256, sizeof(float) * W, st>>>(h->x32, tok, h->tw.lnf_g,
                                                                 h->tw.lnf_b, h->tw.proj,
                                                                 out_dev + (size_t)done * E, C, W, E, 1e-5f);
            RET_IF(check_launch(h, "text_final_kernel"));
        }
    }
    return 0;
}

int fsar_metrics_update(fsar_handle* h, const float* logits_dev, const float* target_labels_dev, int Q, int way,
                        int64_t* counters_dev, int64_t* per_class_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || logits_dev == nullptr || target_labels_dev == nullptr || counters_dev == nullptr || Q < 1 || way < 1)
        return fail(h, FSAR_E_INVALID, "fsar_metrics_update: bad argument");
    Scope s(h, (cudaStream_t)stream, FSAR_K_HEAD_MISC, 0.0, 4.0 * Q * (way + 1));
    metrics_kernel<<<(Q + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        logits_dev, target_labels_dev, Q, way, reinterpret_cast<unsigned long long*>(counters_dev),
        reinterpret_cast<unsigned long long*>(per_class_dev));
    return check_launch(h, "metrics_kernel");
}

int64_t fsar_peek(fsar_handle* h, const char* name, void* dst_host, int64_t numel, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || name == nullptr || dst_host == nullptr) return fail(h, FSAR_E_INVALID, "fsar_peek: NULL argument");
    const int E = h->cfg.embed_dim, S = h->last_S, Q = h->last_Q, T = h->last_T, way = h->last_way;
    const void* src = nullptr;
    int64_t avail = 0;
    size_t esz = sizeof(float);
    const std::string n(name);
    if (n == "support_feats") { src = h->last_sup; avail = (int64_t)S * T * E; }
    else if (n == "target_feats") { src = h->last_tgt; avail = (int64_t)Q * T * E; }
    const HeadWs& w = h->ws[h->last_ws];
    if (src != nullptr) {}
    else if (n == "mod_out") { src = w.mod_out; avail = (int64_t)h->last_rows * E; }
    else if (n == "protos") { src = w.protos; avail = (int64_t)way * T * E; }
    else if (n == "dists") { src = w.dists; avail = (int64_t)Q * way * T * T; }
    else if (n == "cum_dists") { src = w.cum; avail = (int64_t)Q * way; }
    else if (n == "class_index") { src = w.cls; avail = S; esz = sizeof(int); }
    else return fail(h, FSAR_E_NAME, "unknown tap '%s'", name);
    if (numel < avail) avail = numel;
    CU_OK(h, cudaStreamSynchronize((cudaStream_t)stream));
    if (h->compute_stream) CU_OK(h, cudaStreamSynchronize(h->compute_stream));
    CU_OK(h, cudaMemcpy(dst_host, src, esz * (size_t)avail, cudaMemcpyDeviceToHost));
    return avail;
}

int fsar_op_layernorm(fsar_handle* h, const float* x_dev, const float* gamma_dev, const float* beta_dev, int rows, int dim,
                      int out16, void* out_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || x_dev == nullptr || out_dev == nullptr || rows < 1) return fail(h, FSAR_E_INVALID, "fsar_op_layernorm: bad argument");
    return layernorm(h, x_dev, out_dev, gamma_dev, beta_dev, rows, dim, out16 != 0, false, 1, nullptr, nullptr,
                     (cudaStream_t)stream, FSAR_K_LAYERNORM);
}

int fsar_op_gemm(fsar_handle* h, const void* a16_dev, const void* w16_dev, const float* bias_dev, int M, int N, int K, int epi,
                 void* out_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || a16_dev == nullptr || w16_dev == nullptr || out_dev == nullptr) return fail(h, FSAR_E_INVALID, "fsar_op_gemm: NULL argument");
    return gemm(h, FSAR_K_GEMM_QKV, (const T16*)a16_dev, (const T16*)w16_dev, bias_dev, out_dev, M, N, K, epi,
                (cudaStream_t)stream);
}

int fsar_op_attention(fsar_handle* h, const void* qkv16_dev, int n_frames, int L, int heads, void* out16_dev, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || qkv16_dev == nullptr || out16_dev == nullptr || n_frames < 1 || L < 1 || heads < 1)
        return fail(h, FSAR_E_INVALID, "fsar_op_attention: bad argument");
    return attention(h, (const T16*)qkv16_dev, n_frames, L, heads, (T16*)out16_dev, (cudaStream_t)stream);
}

int fsar_op_f32_to_16(fsar_handle* h, const float* src_dev, void* dst16_dev, int64_t numel, void* stream) {
    DeviceGuard guard(h);
    if (h == nullptr || src_dev == nullptr || dst16_dev == nullptr || numel < 1) return fail(h, FSAR_E_INVALID, "fsar_op_f32_to_16: bad argument");
    const long long groups = (numel + 3) / 4;
    Scope s(h, (cudaStream_t)stream, FSAR_K_HEAD_MISC, 0.0, 6.0 * numel);
    f32_to_16_kernel<<<(int)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src_dev, (T16*)dst16_dev, numel);
    return check_launch(h, "f32_to_16_kernel");
}

int64_t fsar_launch_count(const fsar_handle* h) { return h ? h->launches : -1; }

int fsar_profile_begin(fsar_handle* h) {
    if (h == nullptr) return FSAR_E_INVALID;
    for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    h->prof.clear();
    h->profiling = true;
    return 0;
}

int fsar_profile_end(fsar_handle* h, fsar_profile* out) {
    DeviceGuard guard(h);
    if (h == nullptr || out == nullptr) return fail(h, FSAR_E_INVALID, "fsar_profile_end: NULL argument");
    h->profiling = false;
    CU_OK(h, cudaDeviceSynchronize());
    memset(out, 0, sizeof(*out));
    for (auto& r : h->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        out->ms[r.cls] += ms;
        out->launches[r.cls] += 1;
        out->flops[r.cls] += r.flops;
        out->bytes[r.cls] += r.bytes;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->prof.clear();
    return 0;
}

}  // extern "C"
