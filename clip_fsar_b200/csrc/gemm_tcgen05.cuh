// Persistent, warp-specialised TN GEMM on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   C[M, N] = A[M, K] * W[N, K]^T   (both operands K-major 16-bit, fp32 accumulation in TMEM)
//
// This single kernel replaces the dense contractions of the reference ViT
// (/root/reference/models/base/few_shot.py:672 conv1 as a patch GEMM, :635 MHA in/out projections,
//  :626-628 MLP c_fc / c_proj) with the element-wise tails fused into the epilogue:
//   EPI_STORE16 : out16 = acc + bias                      (QKV projection)
//   EPI_QGELU16 : out16 = quick_gelu(acc + bias)          (c_fc + QuickGELU, few_shot.py:616)
//   EPI_RESID32 : x32  += acc + bias                      (attn out-proj / c_proj + residual, :638-639)
//   EPI_STORE32 : out32 = acc + bias                      (conv1 patch GEMM, generic)
//
// Structure (one CTA per SM, 384 threads):
//   warp 0   TMA producer  : cp.async.bulk.tensor 2D tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1   MMA issuer    : one lane issues tcgen05.mma 128 x BN x 16, accumulators double-buffered in TMEM
//   warp 2   TMEM allocator
//   warps 4-11 epilogue    : two warps per TMEM lane quarter (alternating column chunks, so every SM sub-partition
//                            has two epilogue warps to hide MUFU / TMEM-load latency behind each other):
//                            tcgen05.ld (32 lanes x 32 columns) -> registers -> fused tail -> 128B-swizzled smem
//                            staging tile -> TMA store (or TMA reduce-add for the fp32 residual: x += tile is done
//                            in L2, the SM never reads the residual stream)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue),
//            bulk async-groups (one epilogue staging buffer per warp).
#pragma once
#include <type_traits>
#include "ptx.cuh"

namespace fsar {

enum GemmEpilogue : int { EPI_STORE16 = 0, EPI_QGELU16 = 1, EPI_RESID32 = 2, EPI_STORE32 = 4 };

struct GemmParams {
    int M, N, K;        // logical sizes; K is covered in blocks of 64 (TMA zero-fills the tail)
    const float* bias;  // [N] or nullptr
    int reverse;        // walk the row blocks last-to-first: a consumer that reads its producer's output in the opposite
                        // order finds the most recently written rows still in L2 (LRU), see fsar.cu
    int debug;          // only read by the -DFSAR_PROBES build (tools/gemm_probe.py; results are WRONG when set):
                        // 1 = epilogue drains TMEM but stores nothing, 2 = epilogue releases the accumulator without
                        // reading it, 4 = smem staging without the TMA store, 8 = every tile is stored over row block 0
};
#ifdef FSAR_PROBES
#define FSAR_PROBE(flags, bit) (((flags) & (bit)) != 0)
#else
#define FSAR_PROBE(flags, bit) false
#endif

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 x 2 B = 128 B = one swizzle-128B row
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_STAGE_TILE_BYTES = 32 * 128;  // epilogue staging tile: 32 rows x 128 B
constexpr int GEMM_STAGING_BYTES = GEMM_EPI_WARPS * GEMM_STAGE_TILE_BYTES;  // one buffer per epilogue warp

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : ((2 * BN <= 64) ? 64 : ((2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512)));
    static constexpr int BAR_BYTES = 256;  // barriers + tmem pointer
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_STAGING_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment
};

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// QuickGELU x * sigmoid(1.702 x) (few_shot.py:616) of two values, returned as a packed 16-bit pair.
// sigmoid(z) = 0.5 + 0.5 tanh(z / 2)  =>  x * sigmoid(1.702 x) = h + h * tanh(0.851 x), h = 0.5 x, evaluated in packed
// 16-bit arithmetic: ONE MUFU (tanh.approx.f16x2) per TWO elements instead of ex2 + rcp per element. The result is
// stored as a 16-bit GEMM operand anyway; emulating this rounding in the CPU oracle moves the ViT feature error from
// 5.57e-4 to 5.93e-4 (default-init) / 2.32e-3 to 2.30e-3 (stress weights), i.e. below the operand-rounding noise.
template <typename T>
__device__ __forceinline__ uint32_t quick_gelu_pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t quick_gelu_pack2<__half>(float a, float b) {
    const uint32_t x = pack2<__half>(a, b);
    uint32_t t, hx, y;
    asm("{\n\t"
        ".reg .b32 z;\n\t"
        "mul.f16x2 z, %3, %4;\n\t"        // 0.851 x
        "tanh.approx.f16x2 %0, z;\n\t"
        "mul.f16x2 %1, %3, %5;\n\t"       // h = 0.5 x
        "fma.rn.f16x2 %2, %1, %0, %1;\n\t" // h * t + h
        "}\n"
        : "=r"(t), "=r"(hx), "=r"(y)
        : "r"(x), "r"(0x3ACF3ACFu), "r"(0x38003800u));   // 0.851 -> 0x3ACF, 0.5 -> 0x3800 (fp16)
    return y;
}
template <>
__device__ __forceinline__ uint32_t quick_gelu_pack2<__nv_bfloat16>(float a, float b) {
    const uint32_t x = pack2<__nv_bfloat16>(a, b);
    uint32_t t, hx, y;
    asm("{\n\t"
        ".reg .b32 z;\n\t"
        "mul.bf16x2 z, %3, %4;\n\t"
        "tanh.approx.bf16x2 %0, z;\n\t"
        "mul.bf16x2 %1, %3, %5;\n\t"
        "fma.rn.bf16x2 %2, %1, %0, %1;\n\t"
        "}\n"
        : "=r"(t), "=r"(hx), "=r"(y)
        : "r"(x), "r"(0x3F5A3F5Au), "r"(0x3F003F00u));   // 0.851 -> 0x3F5A, 0.5 -> 0x3F00 (bf16)
    return y;
}

// Epilogue of one 128-row x BN-column accumulator tile for ONE warp: lanes [32 q, 32 q + 32) of TMEM (t_base already
// points at them), the column chunks c_first, c_first + c_step, ... (two warps per lane quarter take the chunks of even /
// odd index: (half, 2); one warp takes them all: (0, 1)). tcgen05.ld -> fused tail -> 128B-swizzled smem staging tile ->
// TMA store / reduce-add. `release()` hands the accumulator back to the MMA warp as soon as this warp has read its share.
template <int BN, int EPI, typename T16, typename Release>
__device__ __forceinline__ void gemm_epilogue_tile(uint32_t t_base, int row0, int col_base, const GemmParams& p,
                                                   const CUtensorMap* tmC, uint8_t* stage_ptr, uint32_t row_addr,
                                                   uint32_t sw, int c_first, int c_step, int lane, Release release) {
    constexpr bool kOut16 = (EPI == EPI_STORE16 || EPI == EPI_QGELU16);
    constexpr int CHUNK = kOut16 ? 64 : 32;
    constexpr int NCH = BN / CHUNK;
    bool released = false;
    if (FSAR_PROBE(p.debug, 2)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release();
        return;
    }

#pragma unroll 1
    for (int c = c_first; c < NCH; c += c_step) {
        const int col0 = col_base + c * CHUNK;
        uint32_t w[32];  // the staging row of this thread: 128 B
        if (kOut16) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(t_base + c * 64, r0);
            tmem_ld_32x32b_x32(t_base + c * 64 + 32, r1);
            tc_wait_ld();
            if (c + c_step >= NCH) {  // this warp's share of the accumulator is read: hand it back early
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release();
                released = true;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (p.bias != nullptr) {
                    if (col0 + j < p.N) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                    if (col0 + 32 + j < p.N) b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 32 + j));
                }
                float v0 = __uint_as_float(r0[j]) + b0.x, v1 = __uint_as_float(r0[j + 1]) + b0.y;
                float v2 = __uint_as_float(r0[j + 2]) + b0.z, v3 = __uint_as_float(r0[j + 3]) + b0.w;
                float u0 = __uint_as_float(r1[j]) + b1.x, u1 = __uint_as_float(r1[j + 1]) + b1.y;
                float u2 = __uint_as_float(r1[j + 2]) + b1.z, u3 = __uint_as_float(r1[j + 3]) + b1.w;
                if (EPI == EPI_QGELU16) {
                    w[j / 2] = quick_gelu_pack2<T16>(v0, v1);
                    w[j / 2 + 1] = quick_gelu_pack2<T16>(v2, v3);
                    w[16 + j / 2] = quick_gelu_pack2<T16>(u0, u1);
                    w[16 + j / 2 + 1] = quick_gelu_pack2<T16>(u2, u3);
                } else {
                    w[j / 2] = pack2<T16>(v0, v1);
                    w[j / 2 + 1] = pack2<T16>(v2, v3);
                    w[16 + j / 2] = pack2<T16>(u0, u1);
                    w[16 + j / 2 + 1] = pack2<T16>(u2, u3);
                }
            }
        } else {
            tmem_ld_32x32b_x32(t_base + c * 32, w);
            tc_wait_ld();
            if (c + c_step >= NCH) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release();
                released = true;
            }
            if (p.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (col0 + j < p.N) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                        w[j] = __float_as_uint(__uint_as_float(w[j]) + b.x);
                        w[j + 1] = __float_as_uint(__uint_as_float(w[j + 1]) + b.y);
                        w[j + 2] = __float_as_uint(__uint_as_float(w[j + 2]) + b.z);
                        w[j + 3] = __float_as_uint(__uint_as_float(w[j + 3]) + b.w);
                    }
                }
            }
        }
        if (FSAR_PROBE(p.debug, 1)) continue;
        // the staging buffer must have been drained by the previous TMA store of this warp
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            st_shared_v4(row_addr + ((uint32_t(j) ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < p.M && col0 < p.N && !FSAR_PROBE(p.debug, 4)) {
            const int r_st = FSAR_PROBE(p.debug, 8) ? (row0 & 255) : row0;
            if (EPI == EPI_RESID32) tma_reduce_add_2d(tmC, stage_ptr, col0, r_st);
            else tma_store_2d(tmC, stage_ptr, col0, r_st);
            tma_store_commit();
        }
    }
    if (!released) {  // no chunk for this warp in such a narrow tile: still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release();
    }
}

template <int BN, int EPI, typename T16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool kBf16 = std::is_same<T16, __nv_bfloat16>::value;
    constexpr bool kOut16 = (EPI == EPI_STORE16 || EPI == EPI_QGELU16);
    constexpr int CHUNK = kOut16 ? 64 : 32;  // output columns per staging tile (128 B per row)
    static_assert(BN % CHUNK == 0, "tile width must be a multiple of the staging chunk");

    extern __shared__ uint8_t smem_raw[];
    // swizzle-128B tiles need 1024 B alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                // STAGES x [128][64] 16-bit
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;        // STAGES x [BN][64] 16-bit
    uint8_t* smem_stage = smem + STAGES * Cfg::STAGE_BYTES;  // 8 warps x [32 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + GEMM_STAGING_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;

    const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        // The whole warp walks the loop (warp-uniform control flow keeps addresses / coordinates in uniform
        // registers); one elected lane issues.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            // m-major tile order: the n-tiles of one row block run concurrently on neighbouring CTAs, so an A tile is
            // fetched from HBM once and re-used out of L2 (the weights are a few MB and stay L2-resident anyway)
            const int m_lin = tile / n_tiles;
            const int n_blk = tile - m_lin * n_tiles;
            const int m_blk = p.reverse ? m_tiles - 1 - m_lin : m_lin;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
                    tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane)
        constexpr uint32_t idesc = umma_idesc_f16(GEMM_BM, BN, kBf16, false, false);
        constexpr uint64_t desc_hi = umma_smem_desc_hi(0, 1024, UMMA_LAYOUT_SW128);  // SBO = 8 rows x 128 B
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
                const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the 128 B swizzle atom
                        const uint64_t a_desc = umma_smem_desc(a_addr + k * 32, desc_hi);
                        const uint64_t b_desc = umma_smem_desc(b_addr + k * 32, desc_hi);
                        umma_f16_ss(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;          // TMEM lane quarter this warp may access (== warp % 4)
        const int half = (warp - 4) >> 2;  // which of the two warps of this quarter: even or odd column chunks
        uint8_t* stage_ptr = smem_stage + (warp - 4) * GEMM_STAGE_TILE_BYTES;
        const uint32_t row_addr = smem_u32(stage_ptr) + lane * 128;  // this thread's row inside the staging tile
        const uint32_t sw = uint32_t(lane & 7);                      // 128B swizzle: 16-byte chunk index ^= row % 8
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_lin = tile / n_tiles;
            const int n_blk = tile - m_lin * n_tiles;
            const int m_blk = p.reverse ? m_tiles - 1 - m_lin : m_lin;
            const int row0 = m_blk * GEMM_BM + q * 32;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
            gemm_epilogue_tile<BN, EPI, T16>(t_base, row0, n_blk * BN, p, &tmC, stage_ptr, row_addr, sw, half, 2, lane,
                                             [&]() { mbar_arrive(&tempty_bar[acc]); });
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (lane == 0) tma_store_wait<0>();  // all output tiles are globally written before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace fsar
