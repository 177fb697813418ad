// REJECTED EXPERIMENT (round 2, kept for the record; not compiled into the library -- it lived in clip_fsar_b200/csrc/ and was
// selected for the 16-bit-output epilogues with N >= 512). Parity-green (GEMM operator tests incl. ragged column blocks),
// but slower. M = 18912 (96 frames), sustained 1 s loops, tools/gemm_probe.py (us per launch / TFLOP/s):
//                                      pair kernel 256 x 256        wide 256 x 512 (this file)
//   QKV 2304 x 768   full kernel        65.7 us  1019               75.5 us   887   (v1: 4 stages, 1 staging tile: 884)
//                    no TMA store       53.0                         65.1
//                    drain, no staging  50.9                         62.9
//                    TMEM released unread 47.3  1416                 50.3    1331
//   fc1 3072 x 768   full kernel        88.5 us  1009               98.8 us   903
// Reading: (1) the bare main loop is NOT faster with a quarter fewer L2 bytes per FLOP (50.3 vs 47.3 us): at the 1 kW power
// cap the loop is bound by the tensor pipe's power, not by L2 delivery; (2) the single-buffered accumulator exposes the
// drain (12.6 us per launch vs 3.6 us hidden behind the next tile in the pair kernel), a second staging tile per warp does
// not change that; (3) the TMA stores cost the same 10-13 us per launch in both kernels (they are additive, not
// overlapped -- the one remaining lever on the K = 768 GEMMs).
//
// Wide-tile variant of the CTA-pair GEMM: one 256 x 512 output tile per CTA pair, issued as TWO UMMA 256 x 256 x 16
// per k-step that share the A operand.
//
// Why: the pair kernel (gemm_pair_tcgen05.cuh) is bound by what L2 delivers, not by the tensor pipe: with 256 x 256
// tiles every SM pulls 16 KB of A + 16 KB of W per 64-deep k-block for 512 tensor-core clocks of work = 64 B/clk/SM,
// 9.5 KB/clk chip-wide against an L2 slice throughput of ~6.3 KB/clk (B300_MICROARCH.md, "LTS throughput cap"; the
// epilogue's stores ride the same path). Here a k-block is 16 KB of A + 32 KB of W for 1024 clocks: 48 B/clk/SM, a
// quarter fewer L2 bytes per FLOP.
//
// Price: the fp32 accumulator of a 256 x 512 tile is 512 TMEM columns per SM, i.e. ALL of tensor memory, so it cannot
// be double-buffered. The accumulator is instead handed back in two column HALVES: epilogue warps 4-7 drain columns
// [0, 256), warps 8-11 columns [256, 512) (all four lane quarters each), and the MMA issuer may start the next tile's
// left sub-tile as soon as the left half is free. Used for the 16-bit-output epilogues with N >= 512 (QKV, c_fc): their
// K = 768 main loop is short enough for the operand traffic to matter most, and a 16-bit epilogue drains quickly.
//
// A tile whose right half lies outside N (N = 2304 = 4.5 x 512) skips the right sub-tile's loads and MMAs.
#pragma once
#include "gemm_pair_tcgen05.cuh"

namespace fsar {

constexpr int GEMMW_BN = 512;
constexpr int GEMMW_STAGES = 3;
constexpr int GEMMW_A_BYTES = 128 * GEMM_BK * 2;              // this CTA's 128 rows of the 256-row A tile
constexpr int GEMMW_BSUB_BYTES = 128 * GEMM_BK * 2;           // this CTA's half of one 256-row W sub-tile
constexpr int GEMMW_STAGE_BYTES = GEMMW_A_BYTES + 2 * GEMMW_BSUB_BYTES;   // 48 KB
// The accumulator is single-buffered, so the drain is exposed: every epilogue warp gets TWO staging tiles and keeps one
// TMA store in flight while it converts the next chunk (with one tile the 4 chunks of a warp serialise on the store's
// read of shared memory: measured 884 vs 1037 TFLOP/s for the 256-wide kernel on QKV, call 2 of round 2).
constexpr int GEMMW_STAGING_BYTES = 2 * GEMM_STAGING_BYTES;
constexpr int GEMMW_SMEM_BYTES = GEMMW_STAGES * GEMMW_STAGE_BYTES + GEMMW_STAGING_BYTES + 256 + 1024;

// One warp's share of a 256-column half: lanes [32 q, +32), four 64-column chunks, two alternating staging tiles.
template <int EPI, typename T16, typename Release>
__device__ __forceinline__ void gemm_wide_epilogue(uint32_t t_base, int row0, int col_base, const GemmParams& p,
                                                   const CUtensorMap* tmC, uint8_t* stage_ptr, int lane, Release release) {
    const uint32_t sw = uint32_t(lane & 7);
    if (FSAR_PROBE(p.debug, 2)) {   // probe: hand the columns back unread (main-loop rate of the wide tiles)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release();
        return;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        const int col0 = col_base + c * 64;
        uint32_t r0[32], r1[32], w[32];
        tmem_ld_32x32b_x32(t_base + c * 64, r0);
        tmem_ld_32x32b_x32(t_base + c * 64 + 32, r1);
        tc_wait_ld();
        if (c == 3) {   // this warp's share of the half is in registers: hand the columns back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release();
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (p.bias != nullptr) {
                if (col0 + j < p.N) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                if (col0 + 32 + j < p.N) b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 32 + j));
            }
            const float v0 = __uint_as_float(r0[j]) + b0.x, v1 = __uint_as_float(r0[j + 1]) + b0.y;
            const float v2 = __uint_as_float(r0[j + 2]) + b0.z, v3 = __uint_as_float(r0[j + 3]) + b0.w;
            const float u0 = __uint_as_float(r1[j]) + b1.x, u1 = __uint_as_float(r1[j + 1]) + b1.y;
            const float u2 = __uint_as_float(r1[j + 2]) + b1.z, u3 = __uint_as_float(r1[j + 3]) + b1.w;
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
        if (FSAR_PROBE(p.debug, 1)) continue;   // probe: drain + convert, no staging, no store
        uint8_t* tile = stage_ptr + (c & 1) * GEMM_STAGE_TILE_BYTES;
        const uint32_t row_addr = smem_u32(tile) + lane * 128;
        // the tile written two chunks ago must have been read by its store; the previous chunk's store may still run
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            st_shared_v4(row_addr + ((uint32_t(j) ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < p.M && col0 < p.N && !FSAR_PROBE(p.debug, 4)) {
            tma_store_2d(tmC, tile, col0, row0);
            tma_store_commit();
        }
    }
}

template <int EPI, typename T16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05_wide_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    static_assert(EPI == EPI_STORE16 || EPI == EPI_QGELU16, "wide tiles serve the 16-bit-output epilogues");
    constexpr int STAGES = GEMMW_STAGES;
    constexpr bool kBf16 = std::is_same<T16, __nv_bfloat16>::value;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                   // STAGES x [128][64]
    uint8_t* smem_b = smem + STAGES * GEMMW_A_BYTES;          // STAGES x [2 sub-tiles][128][64]
    uint8_t* smem_stage = smem + STAGES * GEMMW_STAGE_BYTES;  // 8 warps x 2 x [32 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + GEMMW_STAGING_BYTES);
    uint64_t* full_bar = bars;                     // [STAGES]  used in the leader only
    uint64_t* empty_bar = bars + STAGES;           // [STAGES]  one per CTA (multicast commit)
    uint64_t* tfull_bar = bars + 2 * STAGES;       // [1]       one per CTA (multicast commit): the whole tile is ready
    uint64_t* tempty_bar = bars + 2 * STAGES + 1;  // [2]       leader only: column half j drained (2 x 4 warp arrivals)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 3);

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();       // 0 = leader
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    const int m_tiles = (p.M + 255) / 256;
    const int n_tiles = (p.N + GEMMW_BN - 1) / GEMMW_BN;
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
        mbar_init(&tfull_bar[0], 1);
        mbar_init(&tempty_bar[0], 2 * (GEMM_EPI_WARPS / 2));   // the four "left half" warps of both CTAs
        mbar_init(&tempty_bar[1], 2 * (GEMM_EPI_WARPS / 2));
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
        // ------------------------------------------------------------ TMA producer (both CTAs; warp-uniform loop)
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < num_tiles; tile += n_pairs) {
            const int m_lin = tile / n_tiles;
            const int n_blk = tile - m_lin * n_tiles;
            const int m_blk = p.reverse ? m_tiles - 1 - m_lin : m_lin;
            const int a_row = m_blk * 256 + int(rank) * 128;
            const int b_row = n_blk * GEMMW_BN + int(rank) * 128;           // sub-tile 1: + 256
            const int n_sub = (p.N - n_blk * GEMMW_BN > 256) ? 2 : 1;
            const uint32_t bytes = 2u * uint32_t(GEMMW_A_BYTES + n_sub * GEMMW_BSUB_BYTES);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const uint32_t leader_full = map_to_cta(smem_u32(&full_bar[stage]), 0);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], bytes);
                    tma_load_2d_pair(smem_a + stage * GEMMW_A_BYTES, &tmA, leader_full, kb * GEMM_BK, a_row);
                    uint8_t* b_dst = smem_b + stage * 2 * GEMMW_BSUB_BYTES;
                    tma_load_2d_pair(b_dst, &tmB, leader_full, kb * GEMM_BK, b_row);
                    if (n_sub == 2) tma_load_2d_pair(b_dst + GEMMW_BSUB_BYTES, &tmB, leader_full, kb * GEMM_BK, b_row + 256);
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
            constexpr uint32_t idesc = umma_idesc_f16(256, 256, kBf16, false, false);
            constexpr uint64_t desc_hi = umma_smem_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int tile = pair; tile < num_tiles; tile += n_pairs) {
                const int n_blk = tile % n_tiles;
                const int n_sub = (p.N - n_blk * GEMMW_BN > 256) ? 2 : 1;
                // the previous tile's columns must have been drained: the left half gates the first MMA, the right half
                // the second one of the first k-block (every epilogue warp arrives, also for a skipped sub-tile)
                mbar_wait(&tempty_bar[0], tphase ^ 1);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    if (kb == 0) mbar_wait(&tempty_bar[1], tphase ^ 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * GEMMW_A_BYTES);
                    const uint32_t b_addr = smem_u32(smem_b + stage * 2 * GEMMW_BSUB_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            const uint64_t a_desc = umma_smem_desc(a_addr + k * 32, desc_hi);
                            umma_f16_ss_pair(tmem_base, a_desc, umma_smem_desc(b_addr + k * 32, desc_hi), idesc,
                                             (kb | k) != 0 ? 1u : 0u);
                            if (n_sub == 2)
                                umma_f16_ss_pair(tmem_base + 256, a_desc,
                                                 umma_smem_desc(b_addr + GEMMW_BSUB_BYTES + k * 32, desc_hi), idesc,
                                                 (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_pair(&empty_bar[stage], 0x3);                       // slot free in both CTAs
                        if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[0], 0x3);     // the tile is ready in both
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows): warps 4-7 own
        // the left 256 columns, warps 8-11 the right 256; each drains all four 64-column chunks of its lane quarter
        const int q = warp & 3;
        const int sub = (warp - 4) >> 2;
        uint8_t* stage_ptr = smem_stage + (warp - 4) * 2 * GEMM_STAGE_TILE_BYTES;
        const uint32_t leader_tempty = map_to_cta(smem_u32(&tempty_bar[sub]), 0);
        uint32_t tphase = 0;
        for (int tile = pair; tile < num_tiles; tile += n_pairs) {
            const int m_lin = tile / n_tiles;
            const int n_blk = tile - m_lin * n_tiles;
            const int m_blk = p.reverse ? m_tiles - 1 - m_lin : m_lin;
            const int row0 = m_blk * 256 + int(rank) * 128 + q * 32;
            const int col_base = n_blk * GEMMW_BN + sub * 256;
            mbar_wait(&tfull_bar[0], tphase);
            tc_fence_after();
            if (col_base < p.N) {
                const uint32_t t_base = tmem_base + sub * 256 + (uint32_t(q * 32) << 16);
                gemm_wide_epilogue<EPI, T16>(t_base, row0, col_base, p, &tmC, stage_ptr, lane,
                                             [&]() { mbar_arrive_cluster(leader_tempty); });
            } else {   // the right sub-tile of the last column block does not exist: nothing to read, hand it back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_tempty);
            }
            tphase ^= 1;
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
