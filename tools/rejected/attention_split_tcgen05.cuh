// REJECTED EXPERIMENT (round 2, kept for the record; not compiled into the library -- it lived in clip_fsar_b200/csrc/ and
// served L <= 208, non-causal). Parity-green (attention operator tests for L in {1, 5, 64, 65, 129, 197, 208}, all episode
// fixtures), but SLOWER in the episode: headline bench, same box, alternating runs (ms per episode, CUDA events per launch):
//   one softmax thread per row (attention_tcgen05_kernel<T16, false, 208>)   0.3835 / 0.3846   -> 314.7 / 315.5 episodes/s
//   two threads per row (this file, 640 threads, 3 pair barriers per item)    0.4177 / 0.4174   -> 311.8 / 311.8
// The round-1 probe that bounded the gain at -22 % halved the per-thread work WITHOUT the second warp set: with it, the
// same number of tcgen05.ld / MUFU instructions is issued per item by twice the warps, the three 64-thread barriers and
// the smem exchange sit on the per-item critical path, and the register file caps 640 threads at 102 registers.
//
// Attention core for L <= 208 tokens (ViT-B/16: 197, ViT-B/32: 50) with TWO softmax threads per query row.
//
// Same data flow as attention_tcgen05_kernel<T16, false, 208> (attention_tcgen05.cuh: TMA-staged Q/K/V, S = Q K^T and
// O = P V accumulate in tensor memory, P goes back to TMEM as 16-bit pairs, output rows leave through swizzled staging
// tiles and a [frame][token][D] TMA store), but the softmax of a row -- the latency chain of that kernel: ~208 x
// (tcgen05.ld, FFMA, MUFU.EX2, FADD, pack) per thread -- is split between two threads that own the same TMEM lane:
//   warps 4-11  (half 0): score columns [0, 16 * ceil(n16 / 2))         -> P pairs at TMEM columns [0, 56)
//   warps 12-19 (half 1): score columns [16 * ceil(n16 / 2), LK)        -> P pairs at TMEM columns [208, 256)
// (n16 = LK / 16). Half 1's P pairs live in the 48 columns the 208-key score tile leaves free in the group's 256, so a
// thread never overwrites score columns its partner still has to read. Row maxima and row sums are exchanged through
// shared memory under a 64-thread named barrier per warp pair; each thread normalises and stages 32 of the 64 output
// columns. Round-1 probe (profiles/README.md): halving the per-thread softmax work bounds the gain at -22 %.
#pragma once
#include "attention_tcgen05.cuh"

namespace fsar {

constexpr int ATT6_THREADS = 128 + 16 * 32;
constexpr int ATT6_KV_BYTES = ATT5_MAX_KEYS * 128;
constexpr int ATT6_STAGE_BYTES = ATT5_Q_BYTES + 2 * ATT6_KV_BYTES;
constexpr int ATT6_OUT_BYTES = 8 * ATT5_OUT_TILE_BYTES;
constexpr int ATT6_XCH_BYTES = 2 * 2 * 2 * 128 * 4;      // [max | sum][group][half][row] fp32
constexpr int ATT6_SMEM_BYTES = 2 * ATT6_STAGE_BYTES + ATT6_OUT_BYTES + ATT6_XCH_BYTES + 256 + 1024;
constexpr uint32_t ATT6_P1_COL = 208;                    // half 1's P pairs

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <typename T16>
__global__ void __launch_bounds__(ATT6_THREADS, 1)
attention_split_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                               const __grid_constant__ CUtensorMap tmO, const Att5Params p) {
    constexpr bool kBf16 = std::is_same<T16, __nv_bfloat16>::value;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_out = smem + 2 * ATT6_STAGE_BYTES;          // 8 (group, lane quarter) tiles x [32 rows][128 B]
    float* xch_max = reinterpret_cast<float*>(smem_out + ATT6_OUT_BYTES);   // [group][half][128]
    float* xch_sum = xch_max + 2 * 2 * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + ATT6_OUT_BYTES + ATT6_XCH_BYTES);
    uint64_t* full_bar = bars;          // [2] TMA -> MMA
    uint64_t* empty_bar = bars + 2;     // [2] MMA -> TMA
    uint64_t* s_full = bars + 4;        // [2] per group: S ready
    uint64_t* p_full = bars + 6;        // [2] per group: P written (8 warp arrivals)
    uint64_t* o_full = bars + 8;        // [2] per group: O ready
    uint64_t* o_empty = bars + 10;      // [2] per group: O read back (8 warp arrivals)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int n_items = p.n_frames * p.heads;
    const int n16 = p.LK / 16;
    const int n_u0 = (n16 + 1) / 2;     // 16-column units of half 0

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        tma_prefetch_desc(&tmO);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], p.n_mtiles);   // a stage is refilled when every row group's P V has been committed
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 8);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 8);
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
            uint8_t* st = smem + s * ATT6_STAGE_BYTES;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                for (int g = 0; g < p.n_mtiles; ++g)
                    tma_load_2d(st + g * 128 * 128, &tmQ, &full_bar[s], head * 64, frame * p.L + g * 128);
                tma_load_2d(st + ATT5_Q_BYTES, &tmKV, &full_bar[s], p.D + head * 64, frame * p.L);
                tma_load_2d(st + ATT5_Q_BYTES + ATT6_KV_BYTES, &tmKV, &full_bar[s], 2 * p.D + head * 64, frame * p.L);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 3) {
        // ------------------------------------------------------------ MMA issuers: warp 1 drives row group 0, warp 3 group 1
        const int g = (warp == 1) ? 0 : 1;
        if (g < p.n_mtiles) {
            const uint32_t idesc_s = umma_idesc_f16(128, p.LK, kBf16, false, false);   // S = Q K^T, both K-major
            const uint32_t idesc_o = umma_idesc_f16(128, 64, kBf16, false, true);      // O = P V, V is MN-major
            constexpr uint64_t desc_hi = umma_smem_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
            const uint32_t d_s = tmem_base + g * 256;
            const uint32_t d_o = tmem_base + g * 256 + ATT5_O_COL;
            int i = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1, ip = i & 1;
                uint8_t* st = smem + s * ATT6_STAGE_BYTES;
                const uint32_t q_addr = smem_u32(st) + g * 128 * 128, k_addr = smem_u32(st + ATT5_Q_BYTES),
                               v_addr = smem_u32(st + ATT5_Q_BYTES + ATT6_KV_BYTES);
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
                    for (int kk = 0; kk < n16; ++kk) {   // 16 keys = 8 TMEM columns of P pairs per k-step
                        const uint32_t a_tmem = (kk < n_u0) ? d_s + kk * 8 : d_s + ATT6_P1_COL + (kk - n_u0) * 8;
                        umma_f16_ts(d_o, a_tmem, umma_smem_desc(v_addr + kk * 2048, desc_hi), idesc_o, kk != 0 ? 1u : 0u);
                    }
                    umma_commit(&o_full[g]);
                    umma_commit(&empty_bar[s]);   // this group no longer reads Q/K/V of the stage
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ softmax + output: two threads per query row
        const int idx = (warp - 4) >> 2;
        const int g = idx & 1;                // row group
        const int hf = idx >> 1;              // column half of the score row / of the output row
        const int wq = warp & 3;              // TMEM lane quarter
        if (g < p.n_mtiles) {
            const bool warp_valid = (g * 128 + wq * 32) < p.Lm;
            const int bar_id = 1 + g * 4 + wq;
            const int u0 = hf ? n_u0 : 0, u1 = hf ? n16 : n_u0;
            const int col_begin = u0 * 16, ncols = (u1 - u0) * 16;
            const int n32 = ncols / 32;
            const bool tail16 = (ncols & 16) != 0;
            const uint32_t t_row = tmem_base + g * 256 + (uint32_t(wq * 32) << 16);
            const uint32_t t_s = t_row + col_begin;                         // this thread's score columns
            const uint32_t t_p = t_row + (hf ? ATT6_P1_COL : 0u);           // where its P pairs go
            float* my_max = xch_max + (g * 2 + hf) * 128 + wq * 32 + lane;
            const float* other_max = xch_max + (g * 2 + (hf ^ 1)) * 128 + wq * 32 + lane;
            float* my_sum = xch_sum + (g * 2 + hf) * 128 + wq * 32 + lane;
            const float* other_sum = xch_sum + (g * 2 + (hf ^ 1)) * 128 + wq * 32 + lane;
            uint8_t* out_tile = smem_out + (g * 4 + wq) * ATT5_OUT_TILE_BYTES;
            const uint32_t out_row = smem_u32(out_tile) + lane * 128;
            const uint32_t sw = uint32_t(lane & 7);
            int i = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
                const uint32_t ip = i & 1;
                const int it = p.reverse ? n_items - 1 - item : item;
                const int frame = it / p.heads, head = it - frame * p.heads;
                mbar_wait(&s_full[g], ip);
                tc_fence_after();
                if (warp_valid) {
                    // ---- pass 1: maximum over this thread's columns (only the last chunk can hold padded keys >= L)
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                    for (int c = 0; c < n32; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(t_s + c * 32, r);
                        tc_wait_ld();
                        const int lim = p.Lm - (col_begin + c * 32);
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
                    if (tail16) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(t_s + n32 * 32, r);
                        tc_wait_ld();
                        const int lim = p.Lm - (col_begin + n32 * 32);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < lim) mx1 = fmaxf(mx1, __uint_as_float(r[j]));
                    }
                    const float mine = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                    *my_max = mine;
                    pair_barrier(bar_id);                                   // B1: both halves' maxima are in smem
                    const float m_scaled = fmaxf(mine, *other_max) * p.scale_log2e;
                    // ---- pass 2: p = 2^(s * scale * log2e - max), fp32 partial row sum, 16-bit P pairs -> TMEM
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    for (int c = 0; c < n32; ++c) {
                        uint32_t r[32], w[16];
                        tmem_ld_32x32b_x32(t_s + c * 32, r);
                        tc_wait_ld();
                        const int lim = p.Lm - (col_begin + c * 32);
                        if (lim >= 32) {
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
                        tmem_st_32x32b_x16(t_p + c * 16, w);
                    }
                    if (tail16) {
                        uint32_t r[16], w[8];
                        tmem_ld_32x32b_x16(t_s + n32 * 32, r);
                        tc_wait_ld();
                        const int lim = p.Lm - (col_begin + n32 * 32);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float e0 = (j < lim) ? ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -m_scaled)) : 0.f;
                            const float e1 = (j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -m_scaled)) : 0.f;
                            s2 += e0; s3 += e1;
                            w[j / 2] = pack2<T16>(e0, e1);
                        }
                        tmem_st_32x32b_x8(t_p + n32 * 16, w);
                    }
                    *my_sum = (s0 + s1) + (s2 + s3);
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g]);

                mbar_wait(&o_full[g], ip);
                tc_fence_after();
                uint32_t o[32];
                if (warp_valid) {
                    tmem_ld_32x32b_x32(t_row + ATT5_O_COL + hf * 32, o);    // this thread's 32 of the 64 output columns
                    tc_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[g]);
                if (warp_valid) {
                    if (hf == 0 && lane == 0) tma_store_wait_read<0>();     // the previous item's store has drained the tile
                    __syncwarp();
                    pair_barrier(bar_id);                                   // B2: partial sums visible, staging tile free
                    const float inv = 1.0f / (*my_sum + *other_sum);        // rows >= L: garbage, clipped by the TMA store
                    uint4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j].x = pack2<T16>(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
                        v[j].y = pack2<T16>(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
                        v[j].z = pack2<T16>(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
                        v[j].w = pack2<T16>(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4(out_row + ((uint32_t(hf * 4 + j) ^ sw) << 4), v[j].x, v[j].y, v[j].z, v[j].w);
                    fence_proxy_async();
                    pair_barrier(bar_id);                                   // B3: both halves of the 128-byte rows staged
                    if (hf == 0 && lane == 0) {
                        tma_store_3d(&tmO, out_tile, head * 64, g * 128 + wq * 32, frame);
                        tma_store_commit();
                    }
                }
            }
            if (hf == 0 && lane == 0) tma_store_wait<0>();   // output tiles are globally written before the CTA retires
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
