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
//   EPI_PATCH32 : x32[frame, 1 + patch] = acc + pos[1 + patch]   (conv1 + positional embedding, :672-676)
//   EPI_STORE32 : out32 = acc + bias                      (generic, used by the self tests)
//
// Structure (one CTA per SM, 256 threads):
//   warp 0   TMA producer  : cp.async.bulk.tensor 2D tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1   MMA issuer    : one lane issues tcgen05.mma 128 x BN x 16, accumulators double-buffered in TMEM
//   warp 2   TMEM allocator
//   warps 4-7 epilogue     : tcgen05.ld 32 lanes x 32 columns -> registers -> fused tail -> global
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#pragma once
#include <type_traits>
#include "ptx.cuh"

namespace fsar {

enum GemmEpilogue : int { EPI_STORE16 = 0, EPI_QGELU16 = 1, EPI_RESID32 = 2, EPI_PATCH32 = 3, EPI_STORE32 = 4 };

struct GemmParams {
    int M, N, K;            // logical sizes; K is covered in blocks of 64 (TMA zero-fills the tail)
    const float* bias;      // [N] or nullptr
    void* out;              // fp16/bf16 [M, ldo] or fp32 [*, ldo]
    int ldo;                // row pitch of out in elements
    const float* pos;       // EPI_PATCH32: positional embedding [(P + 1), N]
    int patches_per_frame;  // EPI_PATCH32: P (196 for 224/16)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 x 2 B = 128 B = one swizzle-128B row
constexpr int GEMM_THREADS = 256;

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : ((2 * BN <= 64) ? 64 : ((2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512)));
    static constexpr int BAR_BYTES = 256;  // barriers + tmem pointer
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 for manual 1 KB alignment
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

__device__ __forceinline__ float quick_gelu(float v) {
    // x * sigmoid(1.702 x), few_shot.py:616
    return v / (1.0f + __expf(-1.702f * v));
}

template <int BN, int EPI, typename T16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool kBf16 = std::is_same<T16, __nv_bfloat16>::value;

    extern __shared__ uint8_t smem_raw[];
    // swizzle-128B tiles need 1024 B alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                // STAGES x [128][64] 16-bit
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;        // STAGES x [BN][64] 16-bit
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
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

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_tiles;
                const int n_blk = tile - m_blk * n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
                    tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
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
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the 128 B swizzle atom
                        const uint64_t a_desc = umma_smem_desc(a_addr + k * 32, desc_hi);
                        const uint64_t b_desc = umma_smem_desc(b_addr + k * 32, desc_hi);
                        umma_f16_ss(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int q = warp - 4;  // TMEM lane quarter this warp may access (== warp % 4)
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / n_tiles;
            const int n_blk = tile - m_blk * n_tiles;
            const int row = m_blk * GEMM_BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * BN + (uint32_t(q * 32) << 16);

            size_t out_row_off;
            const float* pos_row = nullptr;
            if (EPI == EPI_PATCH32) {
                const int frame = row / p.patches_per_frame;
                const int patch = row - frame * p.patches_per_frame;
                out_row_off = (size_t(frame) * (p.patches_per_frame + 1) + 1 + patch) * size_t(p.ldo);
                pos_row = p.pos + size_t(1 + patch) * p.N;
            } else {
                out_row_off = size_t(row) * size_t(p.ldo);
            }

#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n_blk * BN + c * 32;
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_base + c * 32, r);
                tc_wait_ld();
                if (row_ok && col0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < p.N) {  // N % 8 == 0 is enforced on the host
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                        }
                    }
                    if (EPI == EPI_STORE16 || EPI == EPI_QGELU16) {
                        if (EPI == EPI_QGELU16) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
                        }
                        T16* o = reinterpret_cast<T16*>(p.out) + out_row_off + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 w;
                            w.x = pack2<T16>(v[j], v[j + 1]);
                            w.y = pack2<T16>(v[j + 2], v[j + 3]);
                            w.z = pack2<T16>(v[j + 4], v[j + 5]);
                            w.w = pack2<T16>(v[j + 6], v[j + 7]);
                            if (col0 + j < p.N) *reinterpret_cast<uint4*>(o + j) = w;
                        }
                    } else {
                        float* o = reinterpret_cast<float*>(p.out) + out_row_off + col0;
                        if (EPI == EPI_RESID32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (col0 + j < p.N) {
                                    const float4 x = *reinterpret_cast<const float4*>(o + j);
                                    v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
                                }
                            }
                        } else if (EPI == EPI_PATCH32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (col0 + j < p.N) {
                                    const float4 x = __ldg(reinterpret_cast<const float4*>(pos_row + col0 + j));
                                    v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < p.N)
                                *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace fsar
