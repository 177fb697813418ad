// REJECTED EXPERIMENT (round 1, kept for the record; not compiled into the library — it lived in clip_fsar_b200/csrc/
// next to gemm_pair_tcgen05.cuh and was selected by FSAR_GEMM_QUAD=1). Parity-green (45 operator tests) but not faster:
//   M = 18912            pair kernel     cluster-of-4 with W multicast
//   QKV  2304 x 768      1004 TFLOP/s    1012
//   out   768 x 768       846             781
//   fc1  3072 x 768       994             983
//   fc2   768 x 3072     1050             984
// Only 33 clusters of four fit on a B200 (132 of 148 SMs; every GPC has an odd TPC count), and halving the W bytes each SM
// requests from L2 does not raise the main-loop rate: the bound is what each SM can ingest (32 KB per k-block arrive in
// its shared memory either way), not L2 read volume. See profiles/README.md.
//
// Cluster-of-4 variant of the CTA-pair GEMM (gemm_pair_tcgen05.cuh): two CTA pairs that work on the SAME 256 weight
// rows (n-block) and on two ADJACENT 256-row blocks of A share the weight tile through TMA multicast.
//
// Why: tools/gemm_probe.py shows the pair kernel is bound by L2 -> SM delivery, not by the tensor pipe: with 256 x 256
// tiles every SM pulls 32 KB per 64-deep k-block (16 KB of A, 16 KB of W) and the chip saturates at 9-10 TB/s of L2
// reads (tensor pipe 60 % busy for K = 768; 80 % with the epilogue stores removed, which ride the same path).
// Here each CTA fetches its 16 KB of A but only 8 KB of W (64 rows), multicast to the CTA of the same rank in the other
// pair: 24 KB per SM per k-block, 175 instead of 131 FLOP per L2 byte.
//
// Cluster = 4 CTAs: rank R, pair pr = R >> 1 (row block 2 * band + pr), r = R & 1 (rank inside the pair, 0 = leader).
//   producer (warp 0, every CTA): A rows of its own half tile -> own smem; W rows [r * 128 + pr * 64, +64) of the
//       n-block -> smem offset pr * 8 KB of BOTH CTAs {r, 2 + r}; all bytes are credited to the full barrier of the
//       leader of the receiving pair (cta_group::2 TMA, 64 KB per stage and pair as in the pair kernel).
//   a stage may be refilled only when BOTH pairs have consumed it (the other pair's producer writes into it too):
//       the empty barriers count two arrivals, every leader's tcgen05.commit multicasts to all four CTAs.
//   MMA issuer / epilogue: per pair, as in the pair kernel.
// An odd number of row blocks leaves the last band with one real block; its partner runs on out-of-range rows (TMA
// zero fill, stores clipped).
#pragma once
#include "gemm_pair_tcgen05.cuh"

namespace fsar {

// W piece load with multicast: data to the same smem offset in every CTA of `cta_mask`, completion bytes to the
// barrier at the offset of `bar_cluster_addr` in the leader (even rank) of each destination CTA's pair.
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                    int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
          "h"(cta_mask)
        : "memory");
}

template <int EPI, typename T16>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05_quad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB64,
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
    uint64_t* full_bar = bars;                     // [STAGES]  used in the pair leaders
    uint64_t* empty_bar = bars + STAGES;           // [STAGES]  every CTA, two arrivals (one commit per pair)
    uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]       every CTA (commit multicast inside the pair)
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]       pair leaders (16 warp arrivals)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pr = rank >> 1, r = rank & 1, leader = rank & ~1u;
    const int cid = blockIdx.x >> 2;
    const int n_cl = gridDim.x >> 2;

    const int m_tiles = (p.M + 255) / 256;
    const int bands = (m_tiles + 1) / 2;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int items = bands * n_tiles;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB64);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 2);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_ptr_smem, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (all four CTAs)
        const uint16_t mc_mask = uint16_t((1u << r) | (1u << (2 + r)));
        int stage = 0;
        uint32_t phase = 0;
        for (int item = cid; item < items; item += n_cl) {
            const int band_lin = item / n_tiles;
            const int n_blk = item - band_lin * n_tiles;
            const int band = p.reverse ? bands - 1 - band_lin : band_lin;
            const int a_row = (2 * band + int(pr)) * 256 + int(r) * 128;
            const int b_row = n_blk * BN + int(r) * (BN / 2) + int(pr) * (BN / 4);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const uint32_t leader_full = map_to_cta(smem_u32(&full_bar[stage]), leader);
                if (elect_one()) {
                    if (r == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);
                    tma_load_2d_pair(smem_a + stage * GEMM2_A_BYTES, &tmA, leader_full, kb * GEMM_BK, a_row);
                    tma_load_2d_pair_mc(smem_b + stage * GEMM2_B_BYTES + pr * (GEMM2_B_BYTES / 2), &tmB64, leader_full,
                                        kb * GEMM_BK, b_row, mc_mask);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (the leader of each pair)
        if (r == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(256, BN, kBf16, false, false);
            constexpr uint64_t desc_hi = umma_smem_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
            const uint16_t pair_mask = uint16_t(0x3u << (2 * pr));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = cid; item < items; item += n_cl) {
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
                        umma_commit_pair(&empty_bar[stage], 0xF);                                // all four CTAs
                        if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[acc], pair_mask);      // own pair
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
        // ------------------------------------------------------------ epilogue (every CTA, own 128 rows)
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        uint8_t* stage_ptr = smem_stage + (warp - 4) * GEMM_STAGE_TILE_BYTES;
        const uint32_t row_addr = smem_u32(stage_ptr) + lane * 128;
        const uint32_t sw = uint32_t(lane & 7);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = cid; item < items; item += n_cl) {
            const int band_lin = item / n_tiles;
            const int n_blk = item - band_lin * n_tiles;
            const int band = p.reverse ? bands - 1 - band_lin : band_lin;
            const int row0 = (2 * band + int(pr)) * 256 + int(r) * 128 + q * 32;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
            const uint32_t leader_tempty = map_to_cta(smem_u32(&tempty_bar[acc]), leader);
            gemm_epilogue_tile<BN, EPI, T16>(t_base, row0, n_blk * BN, p, &tmC, stage_ptr, row_addr, sw, half, lane,
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
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace fsar
