// HBM-bound kernels of the CLIP ViT frame encoder (everything that is not a GEMM):
//   patch gather (im2col + fp32->16-bit), LayerNorm (ln_pre / ln_1 / ln_2), ln_post + projection,
//   and the CLS-query attention of the last block (the full attention core is attention_tcgen05.cuh).
// Reference semantics: /root/reference/models/base/few_shot.py:605-611 (LayerNorm in fp32, eps 1e-5),
// :671-688 (VisionTransformer.forward), :633-640 (ResidualAttentionBlock).
#pragma once
#include "ptx.cuh"
#include "gemm_tcgen05.cuh"  // pack2<>

namespace fsar {

constexpr int ATT_HD = 64;        // head dim of every CLIP tower (ViT-B/16, ViT-B/32, ViT-L/14, text)

// ------------------------------------------------------------------------------------------------
// Test-time frame pre-processing of the reference loader, fused: ToTensorVideo (uint8 THWC -> float / 255),
// KineticsResizedCropFewshot (bilinear resize to RH x RW with align_corners = False, then the centre S x S crop;
// datasets/utils/transformations.py:676-716 with idx = TEST_CENTER_CROP, one spatial crop) and NormalizeVideo
// ((x - mean) / std; datasets/base/ssv2_few_shot.py:633-642). uint8 [n, H, W, 3] -> fp32 [n, 3, S, S], i.e. exactly
// the `support_set` / `target_set` tensors of the task dict, so raw frames can cross PCIe as bytes (4x fewer than fp32
// crops, more for larger source frames). One thread per output pixel, the three channels together.
struct PreprocParams {
    int n, H, W, RH, RW, S;
    float mean[3], std[3];
};
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, const PreprocParams p) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.n * p.S * p.S;
    if (idx >= total) return;
    const int x = int(idx % p.S);
    const int y = int((idx / p.S) % p.S);
    const int f = int(idx / ((long long)p.S * p.S));
    const int ry = y + (p.RH - p.S) / 2, rx = x + (p.RW - p.S) / 2;      // position in the resized frame
    // torch upsample_bilinear2d, align_corners = False: src = scale * (dst + 0.5) - 0.5, clamped at 0
    const float sh = float(p.H) / float(p.RH), sw = float(p.W) / float(p.RW);
    float fy = sh * (float(ry) + 0.5f) - 0.5f, fx = sw * (float(rx) + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = int(fy), x0 = int(fx);
    const int y1 = y0 + (y0 < p.H - 1 ? 1 : 0), x1 = x0 + (x0 < p.W - 1 ? 1 : 0);
    const float ly1 = fy - float(y0), lx1 = fx - float(x0);
    const float ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
    const uint8_t* base = src + (size_t)f * p.H * p.W * 3;
    const uint8_t* p00 = base + ((size_t)y0 * p.W + x0) * 3;
    const uint8_t* p01 = base + ((size_t)y0 * p.W + x1) * 3;
    const uint8_t* p10 = base + ((size_t)y1 * p.W + x0) * 3;
    const uint8_t* p11 = base + ((size_t)y1 * p.W + x1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v00 = float(p00[c]) / 255.0f, v01 = float(p01[c]) / 255.0f;
        const float v10 = float(p10[c]) / 255.0f, v11 = float(p11[c]) / 255.0f;
        const float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
        dst[(((size_t)f * 3 + c) * p.S + y) * p.S + x] = (v - p.mean[c]) / p.std[c];
    }
}

// The same transform fused INTO the patch gather ("feeding K1 directly"): uint8 [n, H, W, 3] -> the
// 16-bit im2col rows [n * G * G, Kp] of the conv1 GEMM, k = c * P * P + ky * P + kx. The fp32 NCHW crop (48 MB per
// 80-frame episode, written once and read once) never exists. Arithmetic is preprocess_u8_kernel's, rounded to the operand
// type exactly as patch_gather_kernel rounds the fp32 crop, so both routes give bit-identical patch rows.
// One thread per (frame, y, pair of x): two horizontally adjacent pixels x 3 channels, three 4-byte stores.
template <typename T16>
__global__ void __launch_bounds__(256)
preprocess_patches_u8_kernel(const uint8_t* __restrict__ src, T16* __restrict__ out, const PreprocParams p, int P, int Kp) {
    pdl_trigger();
    pdl_wait();
    const int half_w = p.S >> 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.n * p.S * half_w;
    if (idx >= total) return;
    const int x0 = int(idx % half_w) * 2;
    const int y = int((idx / half_w) % p.S);
    const int f = int(idx / ((long long)half_w * p.S));
    const int G = p.S / P;
    const float sh = float(p.H) / float(p.RH), sw = float(p.W) / float(p.RW);
    const int ry = y + (p.RH - p.S) / 2;
    float fy = sh * (float(ry) + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    const int sy0 = int(fy);
    const int sy1 = sy0 + (sy0 < p.H - 1 ? 1 : 0);
    const float ly1 = fy - float(sy0), ly0 = 1.0f - ly1;
    const uint8_t* base = src + (size_t)f * p.H * p.W * 3;
    float v[2][3];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int rx = x0 + i + (p.RW - p.S) / 2;
        float fx = sw * (float(rx) + 0.5f) - 0.5f;
        fx = fx < 0.f ? 0.f : fx;
        const int sx0 = int(fx);
        const int sx1 = sx0 + (sx0 < p.W - 1 ? 1 : 0);
        const float lx1 = fx - float(sx0), lx0 = 1.0f - lx1;
        const uint8_t* p00 = base + ((size_t)sy0 * p.W + sx0) * 3;
        const uint8_t* p01 = base + ((size_t)sy0 * p.W + sx1) * 3;
        const uint8_t* p10 = base + ((size_t)sy1 * p.W + sx0) * 3;
        const uint8_t* p11 = base + ((size_t)sy1 * p.W + sx1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v00 = float(p00[c]) / 255.0f, v01 = float(p01[c]) / 255.0f;
            const float v10 = float(p10[c]) / 255.0f, v11 = float(p11[c]) / 255.0f;
            const float t = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
            v[i][c] = (t - p.mean[c]) / p.std[c];
        }
    }
    const int py = y / P, ky = y - py * P;
    const int px = x0 / P, kx = x0 - px * P;       // P is even: both pixels fall into the same patch
    T16* dst = out + ((size_t)f * G * G + (size_t)py * G + px) * Kp + (size_t)ky * P + kx;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        *reinterpret_cast<uint32_t*>(dst + (size_t)c * P * P) = pack2<T16>(v[0][c], v[1][c]);
}

// ------------------------------------------------------------------------------------------------
// Patch gather: frames NCHW fp32 [n, 3, S, S] -> A16 [n * G * G, Kp], k = c * P * P + ky * P + kx
// (the flattening of conv1.weight [width, 3, P, P], few_shot.py:659,672). One thread per (frame, c, y, px):
// it reads P contiguous floats and writes P contiguous 16-bit values. Columns [3 P P, Kp) are never
// written; the buffer is zeroed once at creation.
template <typename T16>
__global__ void patch_gather_kernel(const float* __restrict__ frames, T16* __restrict__ out, int n_frames, int S,
                                    int P, int Kp) {
    pdl_trigger();
    pdl_wait();
    const int G = S / P;
    const long long total = (long long)n_frames * 3 * S * G;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int px = int(idx % G);
    long long t = idx / G;
    const int y = int(t % S);
    t /= S;
    const int c = int(t % 3);
    const int n = int(t / 3);
    const int py = y / P, ky = y - py * P;
    const float* src = frames + (((size_t)n * 3 + c) * S + y) * S + (size_t)px * P;
    T16* dst = out + ((size_t)n * G * G + (size_t)py * G + px) * Kp + (size_t)c * P * P + (size_t)ky * P;
    if (P == 16) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(src) + 2);
        const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 3);
        uint4 w0, w1;
        w0.x = pack2<T16>(a.x, a.y); w0.y = pack2<T16>(a.z, a.w); w0.z = pack2<T16>(b.x, b.y); w0.w = pack2<T16>(b.z, b.w);
        w1.x = pack2<T16>(c4.x, c4.y); w1.y = pack2<T16>(c4.z, c4.w); w1.z = pack2<T16>(d.x, d.y); w1.w = pack2<T16>(d.z, d.w);
        reinterpret_cast<uint4*>(dst)[0] = w0;
        reinterpret_cast<uint4*>(dst)[1] = w1;
    } else {
        // P even (14): 8-byte loads, 4-byte stores
        for (int kx = 0; kx < P; kx += 2) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(src + kx));
            *reinterpret_cast<uint32_t*>(dst + kx) = pack2<T16>(a.x, a.y);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension, one warp per row, statistics in fp32 (two-pass in registers).
//   D % 128 == 0, D <= 1024.  OUT16: write T16, else write fp32 (may alias the input: in-place).
//   Input rows are in_pitch elements apart (in_pitch = tokens * D picks the CLS row of every frame), output rows D.
//   EMBED (ln_pre, few_shot.py:675-677): the input row of (frame f, token t) is assembled on the fly as
//     t == 0 : class_embedding + positional_embedding[0]
//     t >= 1 : patch_out[f * (tokens - 1) + t - 1] + positional_embedding[t]     (patch_out = conv1 GEMM output)
//   so the concatenated / position-embedded sequence is never materialised before the LayerNorm.
//   out2 != nullptr (EMBED only): the row just normalised is normalised AGAIN with (gamma2, beta2) while it is still in
//   registers and written as 16-bit -- ln_1 of the first block applied to ln_pre's output (few_shot.py:677 then 637),
//   which saves the first block's LayerNorm launch and its re-read of the residual stream.
template <typename T16, bool OUT16, bool EMBED>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* x, void* out, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int rows, int D, float eps, int tokens, const float* __restrict__ cls_emb,
                 const float* __restrict__ pos, int reverse, long long in_pitch, T16* __restrict__ out2 = nullptr,
                 const float* __restrict__ gamma2 = nullptr, const float* __restrict__ beta2 = nullptr) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= rows) return;
    if (reverse) row = rows - 1 - row;   // blocks are scheduled in index order: last rows first
    const int nv = D >> 7;  // float4 per lane
    float4 v[8];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * in_pitch);
    const float4* pr = nullptr;
    bool is_cls = false;
    if (EMBED) {
        const int f = row / tokens, t = row - f * tokens;
        is_cls = (t == 0);
        pr = reinterpret_cast<const float4*>(pos + (size_t)t * D);
        xr = is_cls ? reinterpret_cast<const float4*>(cls_emb)
                    : reinterpret_cast<const float4*>(x + ((size_t)f * (tokens - 1) + (t - 1)) * D);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nv) {
            v[i] = xr[lane + 32 * i];
            if (EMBED) {
                const float4 b = __ldg(pr + lane + 32 * i);
                v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
            }
        }
    }
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
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / float(D) + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
            float4 y;
            y.x = (v[i].x - mean) * rstd * g.x + b.x;
            y.y = (v[i].y - mean) * rstd * g.y + b.y;
            y.z = (v[i].z - mean) * rstd * g.z + b.z;
            y.w = (v[i].w - mean) * rstd * g.w + b.w;
            if (OUT16) {
                uint2 w;
                w.x = pack2<T16>(y.x, y.y);
                w.y = pack2<T16>(y.z, y.w);
                reinterpret_cast<uint2*>(reinterpret_cast<T16*>(out) + (size_t)row * D)[lane + 32 * i] = w;
            } else {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)row * D)[lane + 32 * i] = y;
            }
            if (EMBED) v[i] = y;
        }
    }
    if (EMBED && out2 != nullptr) {
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nv) s2 += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        const float mean2 = s2 / float(D);
        float q2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nv) {
                const float a = v[i].x - mean2, b = v[i].y - mean2, c = v[i].z - mean2, d = v[i].w - mean2;
                q2 += (a * a + b * b) + (c * c + d * d);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
        const float rstd2 = 1.0f / sqrtf(q2 / float(D) + eps);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nv) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma2) + lane + 32 * i);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta2) + lane + 32 * i);
                uint2 w;
                w.x = pack2<T16>((v[i].x - mean2) * rstd2 * g.x + b.x, (v[i].y - mean2) * rstd2 * g.y + b.y);
                w.y = pack2<T16>((v[i].z - mean2) * rstd2 * g.z + b.z, (v[i].w - mean2) * rstd2 * g.w + b.w);
                reinterpret_cast<uint2*>(out2 + (size_t)row * D)[lane + 32 * i] = w;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The full-size 16-bit-output LayerNorm (ln_1 / ln_2 over all token rows: 23 launches per ViT pass) with the width compiled
// in (NV float4 per lane: 4 / 6 / 8 = width 512 / 768 / 1024) and the residual row read around L1 (ld.global.cg: every row is
// read exactly once). Same arithmetic, same order of operations as layernorm_kernel<T16, true, false>: bit-identical output.
// Measured in the episode (same box, alternating runs): 0.368 vs 0.378 ms, 316.9 vs 314.0 episodes/s. Two and four rows per
// warp (all loads issued before the first reduction) were measured too: no faster (0.379) / slower (0.452).
template <typename T16, int NV>
__global__ void __launch_bounds__(256)
layernorm16_kernel(const float* __restrict__ x, T16* __restrict__ out, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int rows, float eps, int reverse) {
    pdl_trigger();
    pdl_wait();
    constexpr int D = NV * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= rows) return;
    if (reverse) row = rows - 1 - row;
    float4 v[NV];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldcg(xr + lane + 32 * i);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / float(D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / float(D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
        uint2 w;
        w.x = pack2<T16>((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
        w.y = pack2<T16>((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        reinterpret_cast<uint2*>(out + (size_t)row * D)[lane + 32 * i] = w;
    }
}

// ------------------------------------------------------------------------------------------------
// ln_post(x[:, 0, :]) @ proj  (few_shot.py:683-686). proj is [D, E] fp32.
// grid = (ceil(frames / FPC), E / FINAL_COLS); 256 threads = FINAL_COLS columns x FINAL_KSPLIT slices of the D reduction.
constexpr int FINAL_FPC = 4;
constexpr int FINAL_COLS = 32;                    // 32 columns x 8 slices of the D reduction: short dependent-load chains
constexpr int FINAL_KSPLIT = 256 / FINAL_COLS;    // (the kernel is latency-bound: 1.5 MB of weights, a few MFLOP per CTA)
__global__ void __launch_bounds__(256)
final_proj_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ proj, float* __restrict__ out, int n_frames, int tokens, int D, int E,
                  float eps) {
    extern __shared__ float sm[];  // [FPC][D]
    __shared__ float red[2][FINAL_FPC][8];
    pdl_trigger();
    pdl_wait();
    const int f0 = blockIdx.x * FINAL_FPC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // load CLS rows
    for (int f = 0; f < FINAL_FPC; ++f) {
        const int fr = f0 + f;
        for (int d = threadIdx.x; d < D; d += blockDim.x)
            sm[f * D + d] = (fr < n_frames) ? x[(size_t)fr * tokens * D + d] : 0.f;
    }
    __syncthreads();
    // mean
    for (int f = 0; f < FINAL_FPC; ++f) {
        float s = 0.f;
        for (int d = threadIdx.x; d < D; d += blockDim.x) s += sm[f * D + d];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[0][f][warp] = s;
    }
    __syncthreads();
    float mean[FINAL_FPC], rstd[FINAL_FPC];
    for (int f = 0; f < FINAL_FPC; ++f) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[0][f][w];
        mean[f] = s / float(D);
    }
    for (int f = 0; f < FINAL_FPC; ++f) {
        float q = 0.f;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const float a = sm[f * D + d] - mean[f];
            q += a * a;
        }
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) red[1][f][warp] = q;
    }
    __syncthreads();
    for (int f = 0; f < FINAL_FPC; ++f) {
        float q = 0.f;
        for (int w = 0; w < 8; ++w) q += red[1][f][w];
        rstd[f] = 1.0f / sqrtf(q / float(D) + eps);
    }
    for (int f = 0; f < FINAL_FPC; ++f)
        for (int d = threadIdx.x; d < D; d += blockDim.x)
            sm[f * D + d] = (sm[f * D + d] - mean[f]) * rstd[f] * gamma[d] + beta[d];
    __syncthreads();
    // projection: column e of this CTA's slice, slice `kh` of the D reduction; slices are combined through smem
    __shared__ float part[FINAL_KSPLIT][FINAL_FPC][FINAL_COLS];
    const int ce = threadIdx.x & (FINAL_COLS - 1);
    const int e = blockIdx.y * FINAL_COLS + ce;
    const int kh = threadIdx.x / FINAL_COLS;
    const int dspan = D / FINAL_KSPLIT;
    const int dlo = kh * dspan, dhi = dlo + dspan;
    float acc[FINAL_FPC];
#pragma unroll
    for (int f = 0; f < FINAL_FPC; ++f) acc[f] = 0.f;
    if (e < E) {
#pragma unroll 16
        for (int d = dlo; d < dhi; ++d) {
            const float w = __ldg(proj + (size_t)d * E + e);
#pragma unroll
            for (int f = 0; f < FINAL_FPC; ++f) acc[f] = fmaf(sm[f * D + d], w, acc[f]);
        }
    }
#pragma unroll
    for (int f = 0; f < FINAL_FPC; ++f) part[kh][f][ce] = acc[f];
    __syncthreads();
    if (kh == 0 && e < E) {
#pragma unroll
        for (int f = 0; f < FINAL_FPC; ++f) {
            float v = part[0][f][ce];
#pragma unroll
            for (int k = 1; k < FINAL_KSPLIT; ++k) v += part[k][f][ce];
            if (f0 + f < n_frames) out[(size_t)(f0 + f) * E + e] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Attention of the CLS query only, for the LAST block of the frame encoder: VisionTransformer.forward keeps only
// x[:, 0, :] after the transformer (few_shot.py:683), so in the last ResidualAttentionBlock every token still feeds K
// and V but only the CLS row of Q, of the attention output and of the MLP is ever used.
//   q16  [n_frames, D]            (CLS rows of Q, bias included)
//   kv16 [n_frames * L, 2 D]      (K | V column blocks, head h at h * 64)
//   out16 [n_frames, D]
// One CTA per (head, frame), 128 threads. Numerics mirror the tensor-core path: fp32 scores and row sum, un-normalised
// probabilities rounded to the 16-bit operand type before P V, fp32 accumulation, one division at the end.
constexpr int CLS_ATT_MAX_L = 272;
template <typename T16>
__global__ void __launch_bounds__(128)
cls_attention_kernel(const T16* __restrict__ q16, const T16* __restrict__ kv16, T16* __restrict__ out16, int L, int D,
                     float scale_log2e) {
    __shared__ float qs[64];
    __shared__ float sc[CLS_ATT_MAX_L];
    __shared__ float red[4];
    __shared__ float part[64];
    pdl_trigger();
    pdl_wait();
    const int head = blockIdx.x, frame = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 64) qs[tid] = float(q16[(size_t)frame * D + head * 64 + tid]);
    __syncthreads();
    const T16* kbase = kv16 + (size_t)frame * L * 2 * D + head * 64;
    float mx = -INFINITY;
    for (int j = tid; j < L; j += 128) {
        const uint4* kr = reinterpret_cast<const uint4*>(kbase + (size_t)j * 2 * D);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 v = __ldg(kr + c);
            const T16* e = reinterpret_cast<const T16*>(&v);
#pragma unroll
            for (int t = 0; t < 8; ++t) s = fmaf(qs[c * 8 + t], float(e[t]), s);
        }
        sc[j] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * scale_log2e;
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < L; j += 128) {
        const float e = exp2f(fmaf(sc[j], scale_log2e, -mx));
        sum += e;
        sc[j] = float(T16(e));          // the P operand is 16-bit in the tensor-core path
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = (red[0] + red[1]) + (red[2] + red[3]);
    // O[d] = sum_j p_j V[j][d]: threads 0-63 take the even keys, 64-127 the odd keys of column d = tid % 64
    const int d = tid & 63, par = tid >> 6;
    const T16* vbase = kbase + D + d;
    float acc = 0.f;
    for (int j = par; j < L; j += 2) acc = fmaf(sc[j], float(vbase[(size_t)j * 2 * D]), acc);
    if (par == 1) part[d] = acc;
    __syncthreads();
    if (par == 0) out16[(size_t)frame * D + head * 64 + d] = T16((acc + part[d]) / sum);
}

}  // namespace fsar
