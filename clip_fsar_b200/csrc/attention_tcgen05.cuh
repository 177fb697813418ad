// Attention core of the CLIP ViT on the 5th-generation tensor cores (tcgen05 + TMEM), for L <= 208 tokens.
//
//   out[f, :, h] = softmax(Q K^T / sqrt(64)) V      per (frame f, head h); no mask, no dropout
//   (nn.MultiheadAttention in eval mode, /root/reference/models/base/few_shot.py:623, 635)
//
//   qkv16 [n_frames * L, 3 D]  (row = frame * L + token; Q | K | V column blocks, head h at h * 64), 16-bit
//   out16 [n_frames * L, D]
//
// One persistent CTA per SM walks over (frame, head) items. Per item:
//   TMA     : Q (up to 2 x 128 rows), K and V (LK = ceil16(L) rows) -> 128B-swizzled smem, double buffered
//   MMA     : S_g = Q_g K^T  (tcgen05.mma 128 x LK x 16, 4 k-steps, fp32 S in TMEM columns [0, LK) of group g)
//   softmax : ONE THREAD PER QUERY ROW (TMEM lane == row, so row max / row sum are thread-local, no shuffles):
//             pass 1 tcgen05.ld -> running max; pass 2 tcgen05.ld -> exp2 -> fp32 row sum, fp16 pack ->
//             tcgen05.st P into TMEM columns [0, LK/2) (aliasing the S columns already consumed)
//   MMA     : O_g = P_g V    (A operand from TMEM, B = V tile as an MN-major smem operand, LK/16 k-steps,
//             fp32 O in TMEM columns [128, 192) of group g, dead S columns by then)
//   epilogue: tcgen05.ld O -> * 1/rowsum -> fp16 -> one 128-byte row per thread into a 128B-swizzled smem tile ->
//             TMA store of the warp's 32 rows x 128 B through a [frame][token][D] tensor map (rows >= L are clipped, so
//             a padded query tile never touches the next frame); full 128-byte lines instead of 16-byte fragments
// Two row groups g (query rows 0-127 and 128-255) own TMEM columns [0,256) and [256,512) and 4 warps each, so the
// tensor pipe works for one group while the other is in its softmax.
// Warp roles (384 threads): 0 TMA producer, 1 / 3 MMA issuers of group 0 / 1, 2 TMEM allocator, 4-7 softmax group 0,
// 8-11 softmax group 1.
#pragma once
#include "gemm_tcgen05.cuh"  // pack2<>
#include "ptx.cuh"

namespace fsar {

constexpr int ATT5_THREADS = 384;
constexpr int ATT5_MAX_KEYS = 208;                       // 13 x 16; N of one tcgen05.mma must be <= 256
constexpr int ATT5_Q_BYTES = 2 * 128 * 128;              // two 128-row query tiles, 128 B (64 x 16-bit) per row
constexpr int ATT5_KV_BYTES = ATT5_MAX_KEYS * 128;       // 26 KB, multiple of 1024
constexpr int ATT5_STAGE_BYTES = ATT5_Q_BYTES + 2 * ATT5_KV_BYTES;
constexpr int ATT5_OUT_TILE_BYTES = 32 * 128;             // output staging tile of one softmax warp: 32 rows x 128 B
constexpr int ATT5_SMEM_BYTES = 2 * ATT5_STAGE_BYTES + 8 * ATT5_OUT_TILE_BYTES + 256 + 1024;
constexpr uint32_t ATT5_O_COL = 128;                     // O accumulator columns inside a group's 256-column region

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Att5Params {
    int n_frames, L, heads, D;
    int LK;          // keys padded to a multiple of 16
    int n_mtiles;    // 1 or 2 query tiles of 128 rows
    float scale_log2e;
    void* out;       // [n_frames * L, D] 16-bit (written through tmO)
    int reverse;     // walk the (frame, head) items last-to-first (L2 reuse of the QKV rows written last)
    int debug;       // only read by the -DFSAR_PROBES build (tools/gemm_probe.py, results WRONG): 1 skip the max pass,
                     // 2 no exp2, 4 no stores
};

// CAUSAL: query token i attends to keys 0..i only (the additive -inf upper-triangular mask of the CLIP text transformer,
// few_shot.py:777-783); the frame encoder uses CAUSAL = false.
template <typename T16, bool CAUSAL>
__global__ void __launch_bounds__(ATT5_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         const __grid_constant__ CUtensorMap tmO, const Att5Params p) {
    constexpr bool kBf16 = std::is_same<T16, __nv_bfloat16>::value;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_out = smem + 2 * ATT5_STAGE_BYTES;      // 8 softmax warps x [32 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + 8 * ATT5_OUT_TILE_BYTES);
    uint64_t* full_bar = bars;          // [2] TMA -> MMA
    uint64_t* empty_bar = bars + 2;     // [2] MMA -> TMA
    uint64_t* s_full = bars + 4;        // [2] per group: S ready
    uint64_t* p_full = bars + 6;        // [2] per group: P written (4 warp arrivals)
    uint64_t* o_full = bars + 8;        // [2] per group: O ready
    uint64_t* o_empty = bars + 10;      // [2] per group: O read back (4 warp arrivals)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int n_items = p.n_frames * p.heads;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        tma_prefetch_desc(&tmO);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], p.n_mtiles);   // one commit per row group
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (warp-uniform loop, elected lane)
        const uint32_t bytes = uint32_t(p.n_mtiles) * 128 * 128 + 2u * uint32_t(p.LK) * 128;
        int i = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
            const int s = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const int it = p.reverse ? n_items - 1 - item : item;
            const int frame = it / p.heads, head = it - frame * p.heads;
            uint8_t* st = smem + s * ATT5_STAGE_BYTES;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                for (int g = 0; g < p.n_mtiles; ++g)
                    tma_load_2d(st + g * 128 * 128, &tmQ, &full_bar[s], head * 64, frame * p.L + g * 128);
                tma_load_2d(st + ATT5_Q_BYTES, &tmKV, &full_bar[s], p.D + head * 64, frame * p.L);
                tma_load_2d(st + ATT5_Q_BYTES + ATT5_KV_BYTES, &tmKV, &full_bar[s], 2 * p.D + head * 64, frame * p.L);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 3) {
        // ------------------------------------------------------------ MMA issuers: warp 1 drives row group 0, warp 3
        // row group 1, so the S -> softmax -> PV chain of one group never waits behind the other group's barriers
        // (warp-uniform loops, one elected lane issues)
        const int g = (warp == 1) ? 0 : 1;
        if (g < p.n_mtiles) {
            const uint32_t idesc_s = umma_idesc_f16(128, p.LK, kBf16, false, false);   // S = Q K^T, both K-major
            const uint32_t idesc_o = umma_idesc_f16(128, 64, kBf16, false, true);      // O = P V, V is MN-major
            constexpr uint64_t desc_hi = umma_smem_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
            const int ksteps_o = p.LK / 16;
            const uint32_t d_s = tmem_base + g * 256;                 // S (fp32) and P (16-bit pairs) columns
            const uint32_t d_o = tmem_base + g * 256 + ATT5_O_COL;
            int i = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1, ip = i & 1;
                uint8_t* st = smem + s * ATT5_STAGE_BYTES;
                const uint32_t q_addr = smem_u32(st) + g * 128 * 128, k_addr = smem_u32(st + ATT5_Q_BYTES),
                               v_addr = smem_u32(st + ATT5_Q_BYTES + ATT5_KV_BYTES);
                mbar_wait(&full_bar[s], ph);
                mbar_wait(&o_empty[g], ip ^ 1);   // previous item's O (aliases S columns) has been read
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(d_s, umma_smem_desc(q_addr + k * 32, desc_hi), umma_smem_desc(k_addr + k * 32, desc_hi),
                                    idesc_s, k != 0 ? 1u : 0u);
                    umma_commit(&s_full[g]);
                }
                __syncwarp();
                mbar_wait(&p_full[g], ip);
                tc_fence_after();
                if (elect_one()) {
                    for (int kk = 0; kk < ksteps_o; ++kk)
                        umma_f16_ts(d_o, d_s + kk * 8, umma_smem_desc(v_addr + kk * 2048, desc_hi), idesc_o,
                                    kk != 0 ? 1u : 0u);
                    umma_commit(&o_full[g]);
                    umma_commit(&empty_bar[s]);   // this group no longer reads Q/K/V of the stage
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ softmax + output: one thread per query row
        const int g = (warp - 4) >> 2;        // row group
        const int wq = warp & 3;              // TMEM lane quarter
        if (g < p.n_mtiles) {
            const int row = g * 128 + wq * 32 + lane;     // query token of this thread
            const bool warp_valid = (g * 128 + wq * 32) < p.L;
            const int n_keys = (CAUSAL && row + 1 < p.L) ? row + 1 : p.L;   // keys this query row may attend to
            const uint32_t t_row = tmem_base + g * 256 + (uint32_t(wq * 32) << 16);
            const int n32 = p.LK / 32;             // full 32-column chunks of the score row
            const bool tail16 = (p.LK & 16) != 0;  // plus one 16-column chunk
            uint8_t* out_tile = smem_out + (warp - 4) * ATT5_OUT_TILE_BYTES;
            const uint32_t out_row = smem_u32(out_tile) + lane * 128;
            const uint32_t sw = uint32_t(lane & 7);     // 128B swizzle: 16-byte chunk index ^= row % 8
            int i = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
                const uint32_t ip = i & 1;
                const int it = p.reverse ? n_items - 1 - item : item;
            const int frame = it / p.heads, head = it - frame * p.heads;
                float sum = 0.f;
                mbar_wait(&s_full[g], ip);
                tc_fence_after();
                if (warp_valid) {
                    // ---- pass 1: row maximum (only the last chunk can contain padded keys >= L)
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                    if (FSAR_PROBE(p.debug, 1)) mx0 = 0.f;
                    for (int c = 0; c < (FSAR_PROBE(p.debug, 1) ? 0 : n32); ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(t_row + c * 32, r);
                        tc_wait_ld();
                        const int lim = n_keys - c * 32;
                        if (lim >= 32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                mx0 = fmaxf(mx0, __uint_as_float(r[j]));
                                mx1 = fmaxf(mx1, __uint_as_float(r[j + 1]));
                                mx2 = fmaxf(mx2, __uint_as_float(r[j + 2]));
                                mx3 = fmaxf(mx3, __uint_as_float(r[j + 3]));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < lim) mx0 = fmaxf(mx0, __uint_as_float(r[j]));
                        }
                    }
                    if (tail16 && !FSAR_PROBE(p.debug, 1)) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(t_row + n32 * 32, r);
                        tc_wait_ld();
                        const int lim = n_keys - n32 * 32;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < lim) mx1 = fmaxf(mx1, __uint_as_float(r[j]));
                    }
                    const float m_scaled = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2e;
                    // ---- pass 2: p = 2^(s * scale * log2e - max), fp32 row sum (4 partial sums), 16-bit P -> TMEM
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    for (int c = 0; c < n32; ++c) {
                        uint32_t r[32], w[16];
                        tmem_ld_32x32b_x32(t_row + c * 32, r);
                        tc_wait_ld();
                        const int lim = n_keys - c * 32;
                        if (FSAR_PROBE(p.debug, 2)) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) w[j] = r[2 * j];
                        } else if (lim >= 32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -m_scaled));
                                const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -m_scaled));
                                const float e2 = ex2_approx(fmaf(__uint_as_float(r[j + 2]), p.scale_log2e, -m_scaled));
                                const float e3 = ex2_approx(fmaf(__uint_as_float(r[j + 3]), p.scale_log2e, -m_scaled));
                                s0 += e0; s1 += e1; s2 += e2; s3 += e3;
                                w[j / 2] = pack2<T16>(e0, e1);
                                w[j / 2 + 1] = pack2<T16>(e2, e3);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float e0 = (j < lim) ? ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -m_scaled)) : 0.f;
                                const float e1 = (j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -m_scaled)) : 0.f;
                                s0 += e0; s1 += e1;
                                w[j / 2] = pack2<T16>(e0, e1);
                            }
                        }
                        tmem_st_32x32b_x16(t_row + c * 16, w);   // P columns [16c, 16c+16) alias S columns already consumed
                    }
                    if (tail16) {
                        uint32_t r[16], w[8];
                        tmem_ld_32x32b_x16(t_row + n32 * 32, r);
                        tc_wait_ld();
                        const int lim = n_keys - n32 * 32;
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float e0 = (j < lim) ? ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -m_scaled)) : 0.f;
                            const float e1 = (j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -m_scaled)) : 0.f;
                            s2 += e0; s3 += e1;
                            w[j / 2] = pack2<T16>(e0, e1);
                        }
                        tmem_st_32x32b_x8(t_row + n32 * 16, w);
                    }
                    sum = (s0 + s1) + (s2 + s3);
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g]);

                mbar_wait(&o_full[g], ip);
                tc_fence_after();
                uint32_t o0[32], o1[32];
                if (warp_valid) {
                    tmem_ld_32x32b_x32(t_row + ATT5_O_COL, o0);
                    tmem_ld_32x32b_x32(t_row + ATT5_O_COL + 32, o1);
                    tc_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[g]);
                if (warp_valid && !FSAR_PROBE(p.debug, 4)) {
                    const float inv = 1.0f / sum;      // rows >= L: garbage, clipped by the TMA store
                    if (lane == 0) tma_store_wait_read<0>();   // the previous item's store has drained the tile
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4(out_row + ((uint32_t(j) ^ sw) << 4),
                                     pack2<T16>(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv),
                                     pack2<T16>(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv),
                                     pack2<T16>(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv),
                                     pack2<T16>(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv));
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4(out_row + ((uint32_t(4 + j) ^ sw) << 4),
                                     pack2<T16>(__uint_as_float(o1[8 * j]) * inv, __uint_as_float(o1[8 * j + 1]) * inv),
                                     pack2<T16>(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv),
                                     pack2<T16>(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv),
                                     pack2<T16>(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv));
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&tmO, out_tile, head * 64, g * 128 + wq * 32, frame);
                        tma_store_commit();
                    }
                }
            }
            if (lane == 0) tma_store_wait<0>();   // every output tile is globally written before the CTA retires
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace fsar
