// REJECTED EXPERIMENT (round 2, kept for the record; not compiled into the library -- it lived in vit_kernels.cuh and served
// the 16-bit-output LayerNorm launches with >= 2048 dense rows). Parity-green (tests/test_gpu_ops.py::test_layernorm_stream_kernel:
// 6 shapes incl. partial last tile and widths 128 / 512 / 768 / 1024; all episode fixtures), but SLOWER in the episode:
// headline bench, same box, alternating runs (ms per episode over 19.2 launches, CUDA events per launch):
//   warp-per-row kernel (layernorm_kernel, 4 CTAs x 8 warps per SM)          0.3741 / 0.3745  (4.59 TB/s)  -> 311.1 episodes/s
//   this kernel (2 persistent CTAs per SM, 4 x 24 KB cp.async.bulk stages)    0.4836 / 0.4856  (3.55 TB/s)  -> 306.5
// 16 consumer warps per SM that each walk one row at a time through two shuffle reductions do not drain the stages as fast
// as 32 independent warps issue their own loads; the bulk copies were never the limiter.
#pragma once
#include "../../clip_fsar_b200/csrc/vit_kernels.cuh"

namespace fsar {

// ------------------------------------------------------------------------------------------------
// The same LayerNorm (dense fp32 rows in, 16-bit rows out) as a PERSISTENT, bulk-copy-fed stream: the 23 full-size launches
// of a ViT pass (ln_1 / ln_2 over all token rows).
//
// Why: layernorm_kernel is a burst machine -- ncu (profiles/r2_ncu_full.txt): 4 waves of CTAs per SM, every warp issues its
// 6 loads, waits out the DRAM latency, reduces, stores, exits; 6.6 resident warps per scheduler with 1.1 eligible, 44 % of
// DRAM peak, 17.3 us for 87 MB. Here one producer lane keeps STAGES x 8 rows (24 KB per stage at D = 768) of cp.async.bulk
// copies in flight per CTA, two CTAs per SM, so ~190 KB per SM are always outstanding and the eight consumer warps only
// ever touch shared memory: the kernel runs at the memory system's pace instead of at one latency per wave.
//   rows % 1 == any; D % 128 == 0, D <= 1024; x rows are D apart (dense); out rows D apart.
constexpr int LNS_ROWS = 8;              // rows per stage = consumer warps
constexpr int LNS_THREADS = 32 * (LNS_ROWS + 1);   // + the producer warp
template <typename T16>
__global__ void __launch_bounds__(LNS_THREADS)
layernorm_stream_kernel(const float* __restrict__ x, T16* __restrict__ out, const float* __restrict__ gamma,
                        const float* __restrict__ beta, int rows, int D, float eps, int reverse, int stages) {
    extern __shared__ __align__(128) uint8_t lns_smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(lns_smem);          // [stages]
    uint64_t* empty_bar = full_bar + 8;                                  // [stages] (stages <= 8)
    float* tiles = reinterpret_cast<float*>(lns_smem + 128);             // stages x [LNS_ROWS][D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (rows + LNS_ROWS - 1) / LNS_ROWS;
    const int nv = D >> 7;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], LNS_ROWS);
        }
        fence_barrier_init();
    }
    __syncthreads();
    pdl_trigger();
    pdl_wait();
    if (warp == LNS_ROWS) {
        // ---- producer: one lane streams whole tiles (8 consecutive rows are one contiguous span of the dense matrix)
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int tile = reverse ? n_tiles - 1 - t : t;
                const int r0 = tile * LNS_ROWS;
                const int nr = rows - r0 < LNS_ROWS ? rows - r0 : LNS_ROWS;
                const uint32_t bytes = uint32_t(nr) * uint32_t(D) * 4u;
                mbar_wait(&empty_bar[st], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[st], bytes);
                bulk_load_1d(tiles + (size_t)st * LNS_ROWS * D, x + (size_t)r0 * D, bytes, &full_bar[st]);
                if (++st == stages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // ---- consumers: warp w owns row w of every tile
        int st = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int tile = reverse ? n_tiles - 1 - t : t;
            const int row = tile * LNS_ROWS + warp;
            mbar_wait(&full_bar[st], ph);
            if (row < rows) {
                const float4* xr = reinterpret_cast<const float4*>(tiles + ((size_t)st * LNS_ROWS + warp) * D);
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < nv) v[i] = xr[lane + 32 * i];
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean = s / float(D);
                float q = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < nv) {
                        const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                        q += (a * a + bb * bb) + (c * c + d * d);
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                const float rstd = 1.0f / sqrtf(q / float(D) + eps);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[st]);      // the row is in registers: the stage may be refilled
                uint2* orow = reinterpret_cast<uint2*>(out + (size_t)row * D);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < nv) {
                        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);   // L1-resident
                        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
                        uint2 w;
                        w.x = pack2<T16>((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
                        w.y = pack2<T16>((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
                        orow[lane + 32 * i] = w;
                    }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[st]);
            }
            if (++st == stages) { st = 0; ph ^= 1; }
        }
    }
}

}  // namespace fsar
