// CTA-pair (cta_group::2) variant of the persistent TN GEMM: two CTAs on the two SMs of a TPC cooperate on one
// 256 x 256 output tile with UMMA M = 256.
//
// Why: with one CTA per tile (gemm_tcgen05.cuh) every 64-deep k-block stages 16 KB of A and 32 KB of B per SM and the
// tensor core reads them back: 96 B/clk of TMA writes + 96 B/clk of operand reads against the 128 B/clk shared-memory
// port — ncu shows sm__pipe_tensor_cycles_active ~65 %. In a pair each SM stages its 128 rows of A and only HALF of
// the B tile (the tensor cores exchange the halves), 64 + 64 B/clk, and the L2 -> SM operand traffic per FLOP drops
// by a third.
//
// Roles per CTA (384 threads) as in the single-CTA kernel; differences:
//   * both producers issue TMA for their own halves, the bytes are credited to the LEADER's (rank 0) full barrier;
//   * only the leader's MMA thread issues tcgen05.mma.cta_group::2; its tcgen05.commit multicasts the "slot free" and
//     "accumulator ready" arrivals to the barriers of both CTAs;
//   * the epilogue warps of both CTAs drain their own 128 accumulator rows and arrive (remotely for rank 1) on the
//     leader's "accumulator empty" barrier.
#pragma once
#include "gemm_tcgen05.cuh"

namespace fsar {

constexpr int GEMM2_BN = 256;
constexpr int GEMM2_STAGES = 6;
constexpr int GEMM2_A_BYTES = 128 * GEMM_BK * 2;            // this CTA's 128 rows of the 256-row A tile
constexpr int GEMM2_B_BYTES = (GEMM2_BN / 2) * GEMM_BK * 2;  // this CTA's half of the 256 B rows
constexpr int GEMM2_STAGE_BYTES = GEMM2_A_BYTES + GEMM2_B_BYTES;
constexpr int GEMM2_SMEM_BYTES = GEMM2_STAGES * GEMM2_STAGE_BYTES + GEMM_STAGING_BYTES + 256 + 1024;

template <int EPI, typename T16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    constexpr int STAGES = GEMM2_STAGES;
    constexpr int BN = GEMM2_BN;
    constexpr bool kBf16 = std::is_same<T16, __nv_bfloat16>::value;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * GEMM2_A_BYTES;
    uint8_t* smem_stage = smem + STAGES * GEMM2_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + GEMM_STAGING_BYTES);
    uint64_t* full_bar = bars;                     // [STAGES]  used in the leader only
    uint64_t* empty_bar = bars + STAGES;           // [STAGES]  one per CTA (multicast commit)
    uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]       one per CTA (multicast commit)
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]       used in the leader only (16 warp arrivals)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();       // 0 = leader
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    const int m_tiles = (p.M + 255) / 256;
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
            mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS);  // epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_ptr_smem, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();   // barrier inits of both CTAs are visible before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_trigger();   // the next kernel of the stream may set itself up while this grid runs ...
    pdl_wait();      // ... and this one touches global memory only after its predecessor has completed

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs; warp-uniform loop)
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < num_tiles; tile += n_pairs) {
            const int m_lin = tile / n_tiles;
            const int n_blk = tile - m_lin * n_tiles;
            const int m_blk = p.reverse ? m_tiles - 1 - m_lin : m_lin;
            const int a_row = m_blk * 256 + int(rank) * 128;
            const int b_row = n_blk * BN + int(rank) * (BN / 2);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const uint32_t leader_full = map_to_cta(smem_u32(&full_bar[stage]), 0);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);
                    tma_load_2d_pair(smem_a + stage * GEMM2_A_BYTES, &tmA, leader_full, kb * GEMM_BK, a_row);
                    tma_load_2d_pair(smem_b + stage * GEMM2_B_BYTES, &tmB, leader_full, kb * GEMM_BK, b_row);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only; warp-uniform loop)
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(256, BN, kBf16, false, false);
            constexpr uint64_t desc_hi = umma_smem_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += n_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * GEMM2_A_BYTES);
                    const uint32_t b_addr = smem_u32(smem_b + stage * GEMM2_B_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k)
                            umma_f16_ss_pair(d_tmem, umma_smem_desc(a_addr + k * 32, desc_hi),
                                             umma_smem_desc(b_addr + k * 32, desc_hi), idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair(&empty_bar[stage], 0x3);                          // slot free in both CTAs
                        if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[acc], 0x3);      // accumulator ready in both
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
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        uint8_t* stage_ptr = smem_stage + (warp - 4) * GEMM_STAGE_TILE_BYTES;
        const uint32_t row_addr = smem_u32(stage_ptr) + lane * 128;
        const uint32_t sw = uint32_t(lane & 7);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < num_tiles; tile += n_pairs) {
            const int m_lin = tile / n_tiles;
            const int n_blk = tile - m_lin * n_tiles;
            const int m_blk = p.reverse ? m_tiles - 1 - m_lin : m_lin;
            const int row0 = m_blk * 256 + int(rank) * 128 + q * 32;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
            const uint32_t leader_tempty = map_to_cta(smem_u32(&tempty_bar[acc]), 0);
            gemm_epilogue_tile<BN, EPI, T16>(t_base, row0, n_blk * BN, p, &tmC, stage_ptr, row_addr, sw, half, 2, lane,
                                             [&]() { mbar_arrive_cluster(leader_tempty); });
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }

    __syncwarp();
    tc_fence_before();
    cluster_sync_all();   // nobody frees TMEM / exits while the peer may still address this CTA
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace fsar
