// Attention core of the CLIP ViT on the 5th-generation tensor cores (tcgen05 + TMEM), for L <= 257 tokens
// (197 = ViT-B/16, 257 = ViT-L/14 at 224 x 224, 77 = the causal text transformer).
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
//
// Two instances. MAXK = 208 (L <= 208): as above. MAXK = 256 (208 < L <= 257): one tcgen05.mma has N <= 256 and a row
// group owns 256 TMEM columns, so keys / query rows 0..255 run on the tensor cores exactly as above and the ONE token
// beyond (ViT-L/14: 16 x 16 patches + CLS = 257) is folded in with scalar code:
//   * extra key 256: every softmax thread adds s_x = q_row . k_256 (64 FMAs from smem) to its row max / row sum and
//     p_x * v_256 to its O row before normalising;
//   * extra query row 256: four additional warps (12-15, one per SM sub-partition) evaluate that single row from the
//     staged K / V tiles, 64 keys each (lane = key for the scores, lane = two output columns for P V), combine their
//     max / sum / partial outputs through shared memory and write the row's 128 bytes.
// That instance needs 2 x 99 KB of operand stages, so its softmax warps stage the O rows in their (by then dead) rows of
// the Q tile and release the stage after the TMA store has read them.
#pragma once
#include "gemm_tcgen05.cuh"  // pack2<>
#include "ptx.cuh"

namespace fsar {

constexpr int ATT5_THREADS = 384;
constexpr int ATT5_MAX_KEYS = 208;                       // 13 x 16: the instance whose stages leave room for output staging
constexpr int ATT5_MAX_TOKENS = 257;                     // 256 keys on the tensor cores + one scalar token (MAXK = 256)
constexpr int ATT5_Q_BYTES = 2 * 128 * 128;              // two 128-row query tiles, 128 B (64 x 16-bit) per row
constexpr int ATT5_OUT_TILE_BYTES = 32 * 128;            // output staging tile of one softmax warp: 32 rows x 128 B
constexpr uint32_t ATT5_O_COL = 128;                     // O accumulator columns inside a group's 256-column region
template <int MAXK>
struct Att5Cfg {
    static constexpr int KV_BYTES = MAXK * 128;                       // 26 KB / 32 KB, multiples of 1024
    static constexpr int X_BYTES = (MAXK == 256) ? 3 * 1024 : 0;      // 8-row boxes of Q, K, V starting at token 256
    static constexpr int STAGE_BYTES = ATT5_Q_BYTES + 2 * KV_BYTES + X_BYTES;
    static constexpr bool STAGED_OUT = (MAXK == 208);
    static constexpr int OUT_BYTES = STAGED_OUT ? 8 * ATT5_OUT_TILE_BYTES : 0;
    static constexpr int X_WARPS = (MAXK == 256) ? 4 : 0;             // warps 12-15: the scalar query row, 64 keys each
    static constexpr int THREADS = ATT5_THREADS + 32 * X_WARPS;
    static constexpr int SCRATCH_BYTES = (MAXK == 256) ? 2048 : 0;    // reductions of the scalar-row warps
    static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + OUT_BYTES + SCRATCH_BYTES + 256 + 1024;
};

// 16-bit row `row` of a 128B-swizzled [rows][64] tile -> 64 floats (chunk c of row r sits at ((c ^ (r & 7)) << 4))
template <typename T16>
__device__ __forceinline__ void att5_load_row(const uint8_t* tile, int row, float (&dst)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + row * 128 + ((c ^ (row & 7)) << 4));
        const T16* e = reinterpret_cast<const T16*>(&v);
#pragma unroll
        for (int t = 0; t < 8; ++t) dst[c * 8 + t] = float(e[t]);
    }
}
template <typename T16>
__device__ __forceinline__ float att5_dot_row(const uint8_t* tile, int row, const float (&q)[64]) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + row * 128 + ((c ^ (row & 7)) << 4));
        const T16* e = reinterpret_cast<const T16*>(&v);
#pragma unroll
        for (int t = 0; t < 8; ++t) s = fmaf(q[c * 8 + t], float(e[t]), s);
    }
    return s;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Att5Params {
    int n_frames, L, heads, D;
    int Lm;          // tokens handled on the tensor cores: min(L, 256)
    int extra;       // L - Lm (0 or 1): the scalar token of the MAXK = 256 instance
    int LK;          // Lm padded to a multiple of 16
    int n_mtiles;    // 1 or 2 query tiles of 128 rows
    float scale_log2e;
    void* out;       // [n_frames * L, D] 16-bit (written through tmO)
    int reverse;     // walk the (frame, head) items last-to-first (L2 reuse of the QKV rows written last)
    int debug;       // only read by the -DFSAR_PROBES build (tools/gemm_probe.py, results WRONG): 1 skip the max pass,
                     // 2 no exp2, 4 no stores, 8 softmax threads process half of their row, 16 no scalar query row
};

// CAUSAL: query token i attends to keys 0..i only (the additive -inf upper-triangular mask of the CLIP text transformer,
// few_shot.py:777-783); the frame encoder uses CAUSAL = false.
template <typename T16, bool CAUSAL, int MAXK>
__global__ void __launch_bounds__(Att5Cfg<MAXK>::THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                         const Att5Params p) {
    using Cfg = Att5Cfg<MAXK>;
    constexpr int ATT5_STAGE_BYTES = Cfg::STAGE_BYTES, ATT5_KV_BYTES = Cfg::KV_BYTES;
    constexpr int X_OFF = ATT5_Q_BYTES + 2 * ATT5_KV_BYTES;      // Q | K | V rows of token 256 (1 KB each)
    constexpr bool kExtra = (MAXK == 256);
    static_assert(!(kExtra && CAUSAL), "the scalar extra token is not implemented for the causal instance");
    constexpr bool kBf16 = std::is_same<T16, __nv_bfloat16>::value;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_out = smem + 2 * ATT5_STAGE_BYTES;      // 8 softmax warps x [32 rows][128 B] (STAGED_OUT only)
    float* x_scratch = reinterpret_cast<float*>(smem_out + Cfg::OUT_BYTES);   // [4] max | [4] sum | [4][64] partial O
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + Cfg::OUT_BYTES + Cfg::SCRATCH_BYTES);
    uint64_t* full_bar = bars;          // [2] TMA -> MMA
    uint64_t* empty_bar = bars + 2;     // [2] MMA -> TMA
    uint64_t* s_full = bars + 4;        // [2] per group: S ready
    uint64_t* p_full = bars + 6;        // [2] per group: P written (4 warp arrivals)
    uint64_t* o_full = bars + 8;        // [2] per group: O ready
    uint64_t* o_empty = bars + 10;      // [2] per group: O read back (4 warp arrivals)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int n_items = p.n_frames * p.heads;
    const bool extra = kExtra && p.extra != 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        if (kExtra) tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmO);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_bar[i], 1);
            // a stage is refilled when every row group's P V has been committed (+ the scalar-row warp); the MAXK = 256
            // instance also waits for the 4 softmax warps per group, which stage their output rows in the Q tile
            mbar_init(&empty_bar[i], p.n_mtiles * (Cfg::STAGED_OUT ? 1 : 5) + (extra ? Cfg::X_WARPS : 0));
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
        const uint32_t bytes = uint32_t(p.n_mtiles) * 128 * 128 + 2u * uint32_t(p.LK) * 128 + (extra ? 3u * 1024u : 0u);
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
                if (extra) {   // token 256 (+ 7 rows nobody reads): its Q, K and V rows
                    tma_load_2d(st + X_OFF, &tmX, &full_bar[s], head * 64, frame * p.L + 256);
                    tma_load_2d(st + X_OFF + 1024, &tmX, &full_bar[s], p.D + head * 64, frame * p.L + 256);
                    tma_load_2d(st + X_OFF + 2048, &tmX, &full_bar[s], 2 * p.D + head * 64, frame * p.L + 256);
                }
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
    } else if (warp >= 4 && warp < 12) {
        // ------------------------------------------------------------ softmax + output: one thread per query row
        const int g = (warp - 4) >> 2;        // row group
        const int wq = warp & 3;              // TMEM lane quarter
        if (g < p.n_mtiles) {
            const int row = g * 128 + wq * 32 + lane;     // query token of this thread
            const bool warp_valid = (g * 128 + wq * 32) < p.Lm;
            const int n_keys = (CAUSAL && row + 1 < p.Lm) ? row + 1 : p.Lm;   // (tensor-core) keys of this query row
            const uint32_t t_row = tmem_base + g * 256 + (uint32_t(wq * 32) << 16);
            const int n32 = FSAR_PROBE(p.debug, 8) ? p.LK / 64 : p.LK / 32;   // full 32-column chunks of the score row
            const bool tail16 = (p.LK & 16) != 0 && !FSAR_PROBE(p.debug, 8);  // plus one 16-column chunk
            // (probe 8: every softmax thread handles half of its row = the per-thread work of a two-threads-per-row split)
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
                // scalar token 256 as a KEY: its score for this query row and its V row, straight from the staged tiles
                // (they stay valid until this group's P V has been committed, i.e. until after the p_full arrival below)
                float s_x = -INFINITY, p_x = 0.f;
                if (extra && warp_valid) {
                    const uint8_t* stg = smem + (i & 1) * ATT5_STAGE_BYTES;
                    const uint8_t* q_tile = stg + g * 128 * 128;
                    const int qr = wq * 32 + lane;
                    s_x = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 qa = *reinterpret_cast<const uint4*>(q_tile + qr * 128 + ((c ^ (qr & 7)) << 4));
                        const uint4 kb = *reinterpret_cast<const uint4*>(stg + X_OFF + 1024 + (c << 4));   // row 0: no swizzle
                        const T16* qe = reinterpret_cast<const T16*>(&qa);
                        const T16* ke = reinterpret_cast<const T16*>(&kb);
#pragma unroll
                        for (int t = 0; t < 8; ++t) s_x = fmaf(float(qe[t]), float(ke[t]), s_x);
                    }
                }
                if (warp_valid) {
                    // ---- pass 1: row maximum (only the last chunk can contain padded keys >= L)
                    float mx0 = s_x, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
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
                    if (extra) {
                        const float e_x = ex2_approx(fmaf(s_x, p.scale_log2e, -m_scaled));
                        s3 += e_x;
                        p_x = float(T16(e_x));      // rounded like the P operand of the tensor-core keys
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
                if (extra && warp_valid) {   // O row += p_x * v_256 (the stage is still ours: it is released below)
                    const uint8_t* vx = smem + (i & 1) * ATT5_STAGE_BYTES + X_OFF + 2048;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 va = *reinterpret_cast<const uint4*>(vx + (c << 4));
                        const uint4 vb = *reinterpret_cast<const uint4*>(vx + ((4 + c) << 4));
                        const T16* v0 = reinterpret_cast<const T16*>(&va);
                        const T16* v1 = reinterpret_cast<const T16*>(&vb);
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            o0[8 * c + t] = __float_as_uint(fmaf(p_x, float(v0[t]), __uint_as_float(o0[8 * c + t])));
                            o1[8 * c + t] = __float_as_uint(fmaf(p_x, float(v1[t]), __uint_as_float(o1[8 * c + t])));
                        }
                    }
                }
                if (warp_valid && !FSAR_PROBE(p.debug, 4)) {
                    const float inv = 1.0f / sum;      // rows >= L: garbage, clipped by the TMA store / skipped below
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j].x = pack2<T16>(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
                        v[j].y = pack2<T16>(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
                        v[j].z = pack2<T16>(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
                        v[j].w = pack2<T16>(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
                        v[4 + j].x = pack2<T16>(__uint_as_float(o1[8 * j]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
                        v[4 + j].y = pack2<T16>(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
                        v[4 + j].z = pack2<T16>(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
                        v[4 + j].w = pack2<T16>(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
                    }
                    if (Cfg::STAGED_OUT) {
                        if (lane == 0) tma_store_wait_read<0>();   // the previous item's store has drained the tile
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st_shared_v4(out_row + ((uint32_t(j) ^ sw) << 4), v[j].x, v[j].y, v[j].z, v[j].w);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&tmO, out_tile, head * 64, g * 128 + wq * 32, frame);
                            tma_store_commit();
                        }
                    } else {
                        // no room for separate staging tiles next to two 99 KB stages: this warp's 32 rows of the Q tile
                        // (same 128B-swizzled [row][64] layout, dead since S = Q K^T completed) are the staging tile;
                        // the stage is handed back to the producer only after the store has read them
                        uint8_t* q_rows = smem + (i & 1) * ATT5_STAGE_BYTES + g * 128 * 128 + wq * ATT5_OUT_TILE_BYTES;
                        const uint32_t q_row = smem_u32(q_rows) + lane * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st_shared_v4(q_row + ((uint32_t(j) ^ sw) << 4), v[j].x, v[j].y, v[j].z, v[j].w);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&tmO, q_rows, head * 64, g * 128 + wq * 32, frame);
                            tma_store_commit();
                            tma_store_wait_read<0>();
                        }
                    }
                }
                if (!Cfg::STAGED_OUT) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[i & 1]);
                }
            }
            if (lane == 0) tma_store_wait<0>();   // output tiles are globally written before the CTA retires
        }
    } else if (warp >= 12 && extra) {
        // ------------------------------------------------------------ scalar token 256 as a QUERY row: warps 12-15, 64
        // of the 256 staged keys each (warp 12 also takes key 256). Scores: lane = key (2 per lane); P V: lane = two
        // output columns; row max, row sum and the partial O rows are combined through x_scratch.
        const int xw = warp - 12;
        float* x_max = x_scratch;          // [4]
        float* x_sum = x_scratch + 4;      // [4]
        float* x_part = x_scratch + 8;     // [4][64]
        T16* out = reinterpret_cast<T16*>(p.out);
        int i = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
            const int s = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const int it = p.reverse ? n_items - 1 - item : item;
            const int frame = it / p.heads, head = it - frame * p.heads;
            const uint8_t* stg = smem + s * ATT5_STAGE_BYTES;
            const uint8_t* k_tile = stg + ATT5_Q_BYTES;
            const uint8_t* v_tile = k_tile + ATT5_KV_BYTES;
            mbar_wait(&full_bar[s], ph);
            if (FSAR_PROBE(p.debug, 16)) {   // probe: the scalar row costs nothing
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                continue;
            }
            float q[64];
            att5_load_row<T16>(stg + X_OFF, 0, q);
            float sc[2];
            sc[0] = att5_dot_row<T16>(k_tile, 64 * xw + lane, q);
            sc[1] = att5_dot_row<T16>(k_tile, 64 * xw + 32 + lane, q);
            const float s_x = (xw == 0) ? att5_dot_row<T16>(stg + X_OFF + 1024, 0, q) : -INFINITY;   // key 256
            float mx = fmaxf(fmaxf(sc[0], sc[1]), s_x);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) x_max[xw] = mx;
            asm volatile("bar.sync 9, 128;" ::: "memory");
            const float m_scaled = fmaxf(fmaxf(x_max[0], x_max[1]), fmaxf(x_max[2], x_max[3])) * p.scale_log2e;
            float sum = 0.f;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float e = ex2_approx(fmaf(sc[t], p.scale_log2e, -m_scaled));
                sum += e;
                sc[t] = float(T16(e));      // 16-bit like the P operand of the tensor-core rows
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const uint32_t col = uint32_t(lane >> 2), sub = uint32_t(lane & 3) * 4;   // 16-byte chunk / byte offset inside it
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll 8
                for (int src = 0; src < 32; ++src) {
                    const int j = 64 * xw + 32 * t + src;
                    const float pj = __shfl_sync(0xffffffffu, sc[t], src);
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(v_tile + j * 128 + ((col ^ uint32_t(j & 7)) << 4) + sub);
                    const T16* e = reinterpret_cast<const T16*>(&w);
                    a0 = fmaf(pj, float(e[0]), a0);
                    a1 = fmaf(pj, float(e[1]), a1);
                }
            }
            if (xw == 0) {
                const float e_x = ex2_approx(fmaf(s_x, p.scale_log2e, -m_scaled));
                sum += e_x;
                const float p_x = float(T16(e_x));
                const uint32_t w = *reinterpret_cast<const uint32_t*>(stg + X_OFF + 2048 + (col << 4) + sub);
                const T16* e = reinterpret_cast<const T16*>(&w);
                a0 = fmaf(p_x, float(e[0]), a0);
                a1 = fmaf(p_x, float(e[1]), a1);
            }
            x_part[xw * 64 + 2 * lane] = a0;
            x_part[xw * 64 + 2 * lane + 1] = a1;
            if (lane == 0) x_sum[xw] = sum;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp no longer reads the stage
            asm volatile("bar.sync 9, 128;" ::: "memory");
            if (xw == 0) {
                const float inv = 1.0f / ((x_sum[0] + x_sum[1]) + (x_sum[2] + x_sum[3]));
                const float o0 = (x_part[2 * lane] + x_part[64 + 2 * lane]) + (x_part[128 + 2 * lane] + x_part[192 + 2 * lane]);
                const float o1 = (x_part[2 * lane + 1] + x_part[64 + 2 * lane + 1]) +
                                 (x_part[128 + 2 * lane + 1] + x_part[192 + 2 * lane + 1]);
                *reinterpret_cast<uint32_t*>(out + ((size_t)frame * p.L + 256) * p.D + head * 64 + 2 * lane) =
                    pack2<T16>(o0 * inv, o1 * inv);
            }
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
