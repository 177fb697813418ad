// fp32 kernels of the few-shot head: class/text logits, temporal prototype modulator (Transformer_v1),
// per-class prototype means, cosine distance matrix and the OTAM soft-DTW dynamic program.
// Reference semantics (all fp32): /root/reference/models/base/few_shot.py
//   cos_sim 1115-1124, extract_class_indices 1127-1136, Transformer_v1 979-999, PreNormattention_qkv 971-977,
//   Attention_qkv 1035-1073, FeedForward 1643-1654, OTAM_cum_dist_v2 2657-2687,
//   CNN_OTAM_CLIPFSAR.forward eval branch 2932-2990.
#pragma once
#include <cooperative_groups.h>
#include "ptx.cuh"

namespace fsar {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Class index of every support video: cls[s] = rank of labels[s] among the sorted distinct labels
// (torch.unique(support_labels) is sorted, few_shot.py:2950/2960/2965); counts[c] = shots of class c.
// Single CTA; S is small (way * shot).
// It also validates the labels the later kernels index with (the reference fails with an IndexError at
// text_features_test[support_real_class.long()], few_shot.py:2946): a real label outside [0, n_text) -- negative, NaN,
// or beyond the rows that were set -- sets status[0] (mapped host memory, read by the host after the next
// event wait), more distinct labels than `way` sets status[1]. The indexing kernels clamp, so nothing reads out of bounds.
__device__ __forceinline__ int text_row(float label, int n_text) {
    const long long v = (long long)label;   // .long() truncation
    return v < 0 ? 0 : (v >= n_text ? n_text - 1 : (int)v);
}
__global__ void class_index_kernel(const float* __restrict__ labels, int S, int* __restrict__ cls,
                                   int* __restrict__ counts, int way, const float* __restrict__ real_labels, int n_text,
                                   int* status) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    for (int c = threadIdx.x; c < way; c += blockDim.x) counts[c] = 0;
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        if (real_labels != nullptr && status != nullptr) {
            const float rl = real_labels[s];
            if (!(rl >= 0.f && rl < float(n_text))) {
                reinterpret_cast<volatile int*>(status)[0] = 1;
                __threadfence_system();
            }
        }
        const long long ls = (long long)labels[s];  // .long() truncation
        int rank = 0;
        for (int a = 0; a < S; ++a) {
            const long long la = (long long)labels[a];
            if (la < ls) {
                bool first = true;
                for (int b = 0; b < a; ++b)
                    if ((long long)labels[b] == la) { first = false; break; }
                rank += first ? 1 : 0;
            }
        }
        cls[s] = rank;
        if (rank < way) {
            atomicAdd(&counts[rank], 1);
        } else if (status != nullptr) {
            reinterpret_cast<volatile int*>(status)[1] = 1;
            __threadfence_system();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// class_text_logits = cos_sim(mean_T(cat[support, target]), text_features_train) * scale  (few_shot.py:2937-2939)
// One CTA per video.
__global__ void __launch_bounds__(256)
class_text_logits_kernel(const float* __restrict__ sup, int S, const float* __restrict__ tgt, int Q, int T, int E,
                         const float* __restrict__ text, int C, const float* __restrict__ scale,
                         float* __restrict__ out) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    extern __shared__ float sm[];  // [E]
    __shared__ float red[8];
    const int vid = blockIdx.x;
    const float* src = (vid < S) ? sup + (size_t)vid * T * E : tgt + (size_t)(vid - S) * T * E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float sq = 0.f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float a = 0.f;
        for (int t = 0; t < T; ++t) a += src[(size_t)t * E + e];
        a /= float(T);
        sm[e] = a;
        sq += a * a;
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float xn = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) xn += red[w];
    xn = sqrtf(xn);
    const float sc = scale[0];
    for (int c = warp; c < C; c += (blockDim.x >> 5)) {
        const float* tr = text + (size_t)c * E;
        float dot = 0.f, tn = 0.f;
        for (int e = lane; e < E; e += 32) {
            const float tv = tr[e];
            dot = fmaf(sm[e], tv, dot);
            tn = fmaf(tv, tv, tn);
        }
        dot = warp_sum(dot);
        tn = sqrtf(warp_sum(tn));
        if (lane == 0) out[(size_t)vid * C + c] = dot / (xn * tn + 0.01f) * sc;
    }
}

// ------------------------------------------------------------------------------------------------
// Modulator input rows. Row layout of `seq` (fp32, [rows, E]):
//   rows [0, Q*T)                          : query sequences (target features, T tokens each)
//   rows [Q*T, Q*T + n_sup_seq * (T + 1))  : support sequences, T frame tokens + 1 text token
// merge_before (TRAIN.MERGE_BEFORE, few_shot.py:2949-2954): n_sup_seq = way, tokens are per-class means over shots
// (frames and text token alike); else n_sup_seq = S and the text token is text_test[real_support_labels[s]] (2946).
__global__ void __launch_bounds__(128)
build_sequences_kernel(const float* __restrict__ sup, const float* __restrict__ tgt, const float* __restrict__ text_test,
                       const float* __restrict__ real_labels, const int* __restrict__ cls, const int* __restrict__ counts,
                       int S, int Q, int T, int E, int way, int merge_before, int n_text, float* __restrict__ seq) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    const int row = blockIdx.x;
    const int qrows = Q * T;
    float* dst = seq + (size_t)row * E;
    if (row < qrows) {
        for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = tgt[(size_t)row * E + e];
        return;
    }
    const int r = row - qrows;
    const int sq = r / (T + 1), tok = r - sq * (T + 1);
    if (!merge_before) {
        const float* src = (tok < T) ? sup + ((size_t)sq * T + tok) * E
                                     : text_test + (size_t)text_row(real_labels[sq], n_text) * E;
        for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = src[e];
    } else {
        const float inv = 1.0f / float(counts[sq]);
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            float a = 0.f;
            for (int s = 0; s < S; ++s) {
                if (cls[s] == sq) {
                    a += (tok < T) ? sup[((size_t)s * T + tok) * E + e]
                                   : text_test[(size_t)text_row(real_labels[s], n_text) * E + e];
                }
            }
            dst[e] = a * inv;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 linear layer  C[R, N] = A[R, K] * W[N, K]^T (+ bias) (GELU) (+ residual).  SIMT fp32 FMA: the modulator is
// 3.1 M parameters / 0.5 GFLOP per episode, i.e. bound by latency and by how many SMs stream the weights, not by
// math; fp32 keeps parity with the reference at ~1e-6.
// CTA tile 16 rows x 32 columns; the K dimension is split over the CTA's 4 warps (each warp owns K/4 and private
// smem slabs, so the main loop has no block-wide barrier and 12 float4 loads per lane are in flight while the
// previous slab is being multiplied); the four partial tiles are summed through smem in a fixed order
// (deterministic). K % 128 == 0 (512 / 2048 / 768 / 128 / 256 here).
enum LinAct : int { LIN_NONE = 0, LIN_GELU = 1 };
constexpr int LIN_BM = 16, LIN_BN = 32, LIN_BK = 32, LIN_WARPS = 4, LIN_THREADS = 32 * LIN_WARPS;

// One 16 x 32 output tile at (r0, n0) over the K range [kofs, kofs + klen) (klen % 128 == 0), by one CTA of 128 threads.
// `partial` != nullptr: write the raw partial sums there (split-K over CTAs; bias / activation / residual are applied by
// whoever combines the partials); else finish the tile: + bias, activation, + residual -> C.
struct LinSmem {
    float a[LIN_WARPS][LIN_BK][LIN_BM + 4];   // [warp][k][row]
    float w[LIN_WARPS][LIN_BK][LIN_BN + 4];   // [warp][k][col]
};
template <int ACT, bool PREFETCH_ALL>
__device__ __forceinline__ void linear_f32_tile(LinSmem& sm, const float* __restrict__ A, const float* __restrict__ W,
                                                const float* __restrict__ bias, const float* residual, float* C,
                                                float* partial, int R, int N, int K, int r0, int n0, int kofs, int klen) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ty = lane >> 3;  // rows ty * 4 .. + 3
    const int tx = lane & 7;   // cols tx * 4 .. + 3
    const int kspan = klen / LIN_WARPS;
    const int kbeg = kofs + warp * kspan, kend = kbeg + kspan;
    // loader mapping: one float4 along K per (row, lane & 7); 4 rows of 8 float4 per pass
    const int lrow = lane >> 3, lk = (lane & 7) * 4;
    float4 ra[4], rw[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + lrow + 4 * i;
            ra[i] = (r < R) ? *reinterpret_cast<const float4*>(A + (size_t)r * K + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = n0 + lrow + 4 * i;
            rw[i] = (n < N) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    float acc[4][4] = {};
    float (*a_s)[LIN_BM + 4] = sm.a[warp];
    float (*w_s)[LIN_BN + 4] = sm.w[warp];
    auto stage = [&](const float4 (&pa)[4], const float4 (&pw)[8]) {   // registers -> this warp's k-major slabs
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lrow + 4 * i;
            a_s[lk][r] = pa[i].x; a_s[lk + 1][r] = pa[i].y; a_s[lk + 2][r] = pa[i].z; a_s[lk + 3][r] = pa[i].w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = lrow + 4 * i;
            w_s[lk][n] = pw[i].x; w_s[lk + 1][n] = pw[i].y; w_s[lk + 2][n] = pw[i].z; w_s[lk + 3][n] = pw[i].w;
        }
    };
    auto multiply = [&]() {
#pragma unroll
        for (int k = 0; k < LIN_BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&a_s[k][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&w_s[k][tx * 4]);
            acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
            acc[2][0] = fmaf(a.z, w.x, acc[2][0]); acc[2][1] = fmaf(a.z, w.y, acc[2][1]);
            acc[2][2] = fmaf(a.z, w.z, acc[2][2]); acc[2][3] = fmaf(a.z, w.w, acc[2][3]);
            acc[3][0] = fmaf(a.w, w.x, acc[3][0]); acc[3][1] = fmaf(a.w, w.y, acc[3][1]);
            acc[3][2] = fmaf(a.w, w.z, acc[3][2]); acc[3][3] = fmaf(a.w, w.w, acc[3][3]);
        }
    };
    if (PREFETCH_ALL && kspan == 4 * LIN_BK) {
        // K range of 512 (every modulator linear at embed_dim 512: K = 512, or c_proj split four ways): ALL four slabs of
        // this warp are requested before the first FMA -- 48 float4 per lane in flight -- because the weights come cold
        // from HBM and a one-slab-ahead pipeline pays the latency four times (same order of FMAs: same numbers). Measured:
        // the six-launch modulator 0.1155 -> 0.0966 ms per episode. The fused cooperative kernel keeps the one-slab
        // pipeline (PREFETCH_ALL = false): with 224 registers only two of its CTAs fit per SM and it ran 0.132 vs 0.093 ms
        float4 pa[4][4], pw[4][8];
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
            const int k0 = kbeg + sl * LIN_BK;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + lrow + 4 * i;
                pa[sl][i] = (r < R) ? *reinterpret_cast<const float4*>(A + (size_t)r * K + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = n0 + lrow + 4 * i;
                pw[sl][i] = (n < N) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
            stage(pa[sl], pw[sl]);
            __syncwarp();
            multiply();
            __syncwarp();
        }
    } else {
        fetch(kbeg);
        for (int k0 = kbeg; k0 < kend; k0 += LIN_BK) {
            stage(ra, rw);
            __syncwarp();
            if (k0 + LIN_BK < kend) fetch(k0 + LIN_BK);   // in flight while this slab is multiplied
            multiply();
            __syncwarp();
        }
    }
    // combine the four K-partials in a fixed order: red[warp][row][col] aliases the (now dead) weight slabs
    __syncthreads();
    float* red = &sm.w[0][0][0];   // 4 * 16 * 32 floats = 8 KB <= sizeof(sm.w)
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(red + ((warp * LIN_BM + ty * 4 + i) * LIN_BN + tx * 4)) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    // 512 outputs, 128 threads: thread t finishes row t / 8, cols (t % 8) * 4 .. + 3
    const int orow = threadIdx.x >> 3, ocol = (threadIdx.x & 7) * 4;
    float4 v = *reinterpret_cast<const float4*>(red + (orow * LIN_BN + ocol));
#pragma unroll
    for (int wv = 1; wv < LIN_WARPS; ++wv) {
        const float4 t = *reinterpret_cast<const float4*>(red + ((wv * LIN_BM + orow) * LIN_BN + ocol));
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    __syncthreads();               // the slabs are reused by the caller's next tile
    const int r = r0 + orow;
    if (r >= R) return;
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + ocol + j;
        if (n >= N) continue;
        float y = o[j];
        if (partial != nullptr) {
            partial[(size_t)r * N + n] = y;
            continue;
        }
        if (bias != nullptr) y += bias[n];
        if (ACT == LIN_GELU) y = 0.5f * y * (1.0f + erff(y * 0.70710678118654752440f));  // nn.GELU() exact
        if (residual != nullptr) y += residual[(size_t)r * N + n];
        C[(size_t)r * N + n] = y;
    }
}

template <int ACT>
__global__ void __launch_bounds__(LIN_THREADS)
linear_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                  const float* residual, float* C, int R, int N, int K) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    __shared__ __align__(16) LinSmem sm;
    linear_f32_tile<ACT, true>(sm, A, W, bias, residual, C, nullptr, R, N, K, blockIdx.y * LIN_BM, blockIdx.x * LIN_BN, 0, K);
}

// ------------------------------------------------------------------------------------------------
// Modulator attention (Attention_qkv.forward, few_shot.py:1055-1073): per (sequence, head)
// softmax(q k^T * dim_head^-0.5) v over n_tok <= 33 tokens. q/k/v: [rows, inner] fp32, head h at h * dh.
// `pitch` is the row pitch of q/k/v (they are column blocks of one fused [rows, 3 * inner] projection), o is [rows, inner].
// Sequences: the first n_q have T tokens starting at row i * T; the following have T + 1 tokens.
constexpr int MOD_MAX_TOK = 40;
__global__ void __launch_bounds__(128)
modulator_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                           float* __restrict__ o, int n_q_seq, int T, int pitch, int inner, int dh, float scale) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    extern __shared__ float sm[];  // q[n][dh], k[n][dh], v[n][dh], p[n][n+1]
    const int seq = blockIdx.x, head = blockIdx.y;
    int row0, n;
    if (seq < n_q_seq) { row0 = seq * T; n = T; }
    else { row0 = n_q_seq * T + (seq - n_q_seq) * (T + 1); n = T + 1; }
    float* sq = sm;
    float* sk = sq + n * dh;
    float* sv = sk + n * dh;
    float* sp = sv + n * dh;
    for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
        const int t = i / dh, d = i - t * dh;
        const size_t g = (size_t)(row0 + t) * pitch + head * dh + d;
        sq[i] = q[g]; sk[i] = k[g]; sv[i] = v[g];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
        const int a = i / n, b = i - a * n;
        float dot = 0.f;
        for (int d = 0; d < dh; ++d) dot = fmaf(sq[a * dh + d], sk[b * dh + d], dot);
        sp[a * (n + 1) + b] = dot * scale;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        float mx = -INFINITY;
        for (int b = 0; b < n; ++b) mx = fmaxf(mx, sp[a * (n + 1) + b]);
        float s = 0.f;
        for (int b = 0; b < n; ++b) { const float e = expf(sp[a * (n + 1) + b] - mx); sp[a * (n + 1) + b] = e; s += e; }
        const float inv = 1.0f / s;
        for (int b = 0; b < n; ++b) sp[a * (n + 1) + b] *= inv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
        const int a = i / dh, d = i - a * dh;
        float acc = 0.f;
        for (int b = 0; b < n; ++b) acc = fmaf(sp[a * (n + 1) + b], sv[b * dh + d], acc);
        o[(size_t)(row0 + a) * inner + head * dh + d] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// One Transformer_v1 layer as ONE persistent cooperative kernel (few_shot.py:990-999):
//   LN -> fused QKV projection -> 8-head attention in shared memory -> out-proj + bias + residual -> FFN (Linear + exact
//   GELU, Linear split-K over CTAs) + bias + residual,
// seven phases separated by grid barriers instead of six launches. Every phase is spread over the whole grid (the modulator
// is 0.54 GFLOP of fp32 FMAs over 12.6 MB of weights: it needs all SMs, not one cluster); the linears run the tile routine
// of linear_f32_kernel (same numbers), the last one split four ways along K so that 384 CTAs stream c_proj instead of 96
// walking 16 dependent slabs each; the partials are combined in a fixed order (deterministic).
// Launched with cudaLaunchCooperativeKernel (co-residency of the grid is guaranteed by the launch, no hand-made barrier).
struct ModFusedParams {
    const float* x;      // [rows, E]: n_q sequences of T tokens, then n_s sequences of T + 1 tokens
    float* out;          // [rows, E]
    int n_q, n_s, T, E, inner, F, heads, dh;
    float scale, eps;
    const float *norm_g, *norm_b, *w_qkv, *w_out, *b_out, *w_fc, *b_fc, *w_proj, *b_proj;
    float *ln, *qkv, *att, *y, *hid, *part;   // scratch: [rows,E] [rows,3 inner] [rows,inner] [rows,E] [rows,F] [4][rows,E]
};
constexpr int MODF_KSPLIT = 4;

__global__ void __launch_bounds__(LIN_THREADS)
modulator_fused_kernel(const ModFusedParams p) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) uint8_t modf_smem[];     // LinSmem (linear phases) | q, k, v, p tiles (attention phase)
    LinSmem& sm = *reinterpret_cast<LinSmem*>(modf_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = p.n_q * p.T + p.n_s * (p.T + 1);
    const int mt = (rows + LIN_BM - 1) / LIN_BM;
    const int E = p.E;

    // ---- phase 1: the shared LayerNorm of q / k / v (PreNormattention_qkv, few_shot.py:974-977), one warp per row
    {
        const int nv = E >> 7;
        for (int row = blockIdx.x * LIN_WARPS + warp; row < rows; row += gridDim.x * LIN_WARPS) {
            float4 v[8];
            const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)row * E);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) v[i] = xr[lane + 32 * i];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            s = warp_sum(s);
            const float mean = s / float(E);
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) {
                    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                    q += (a * a + b * b) + (c * c + d * d);
                }
            q = warp_sum(q);
            const float rstd = 1.0f / sqrtf(q / float(E) + p.eps);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(p.norm_g) + lane + 32 * i);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.norm_b) + lane + 32 * i);
                    float4 y;
                    y.x = (v[i].x - mean) * rstd * g.x + b.x;
                    y.y = (v[i].y - mean) * rstd * g.y + b.y;
                    y.z = (v[i].z - mean) * rstd * g.z + b.z;
                    y.w = (v[i].w - mean) * rstd * g.w + b.w;
                    reinterpret_cast<float4*>(p.ln + (size_t)row * E)[lane + 32 * i] = y;
                }
        }
    }
    grid.sync();
    // ---- phase 2: to_q | to_k | to_v in one [3 inner, E] projection, no bias (few_shot.py:1046-1048)
    {
        const int N = 3 * p.inner, nt = (N + LIN_BN - 1) / LIN_BN;
        for (int t = blockIdx.x; t < mt * nt; t += gridDim.x)
            linear_f32_tile<LIN_NONE, false>(sm, p.ln, p.w_qkv, nullptr, nullptr, p.qkv, nullptr, rows, N, E, (t / nt) * LIN_BM,
                                      (t % nt) * LIN_BN, 0, E);
    }
    grid.sync();
    // ---- phase 3: softmax(q k^T dh^-1/2) v per (sequence, head) in shared memory (Attention_qkv.forward, 1055-1073)
    {
        float* fsm = reinterpret_cast<float*>(modf_smem);
        const int n_seq = p.n_q + p.n_s, pitch = 3 * p.inner, dh = p.dh;
        for (int it = blockIdx.x; it < n_seq * p.heads; it += gridDim.x) {
            const int seq = it / p.heads, head = it - seq * p.heads;
            int row0, n;
            if (seq < p.n_q) { row0 = seq * p.T; n = p.T; }
            else { row0 = p.n_q * p.T + (seq - p.n_q) * (p.T + 1); n = p.T + 1; }
            float* sq = fsm;
            float* sk = sq + n * dh;
            float* sv = sk + n * dh;
            float* sp = sv + n * dh;
            for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
                const int t = i / dh, d = i - t * dh;
                const size_t g = (size_t)(row0 + t) * pitch + head * dh + d;
                sq[i] = p.qkv[g]; sk[i] = p.qkv[g + p.inner]; sv[i] = p.qkv[g + 2 * p.inner];
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
                const int a = i / n, b = i - a * n;
                float dot = 0.f;
                for (int d = 0; d < dh; ++d) dot = fmaf(sq[a * dh + d], sk[b * dh + d], dot);
                sp[a * (n + 1) + b] = dot * p.scale;
            }
            __syncthreads();
            for (int a = threadIdx.x; a < n; a += blockDim.x) {
                float mx = -INFINITY;
                for (int b = 0; b < n; ++b) mx = fmaxf(mx, sp[a * (n + 1) + b]);
                float sum = 0.f;
                for (int b = 0; b < n; ++b) { const float e = expf(sp[a * (n + 1) + b] - mx); sp[a * (n + 1) + b] = e; sum += e; }
                const float inv = 1.0f / sum;
                for (int b = 0; b < n; ++b) sp[a * (n + 1) + b] *= inv;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
                const int a = i / dh, d = i - a * dh;
                float acc = 0.f;
                for (int b = 0; b < n; ++b) acc = fmaf(sp[a * (n + 1) + b], sv[b * dh + d], acc);
                p.att[(size_t)(row0 + a) * p.inner + head * dh + d] = acc;
            }
            __syncthreads();   // the tiles are refilled by the next item
        }
    }
    grid.sync();
    // ---- phase 4: to_out + bias + residual (few_shot.py:1050-1053, 992): y = att W_o^T + b_o + x
    {
        const int nt = (E + LIN_BN - 1) / LIN_BN;
        for (int t = blockIdx.x; t < mt * nt; t += gridDim.x)
            linear_f32_tile<LIN_NONE, false>(sm, p.att, p.w_out, p.b_out, p.x, p.y, nullptr, rows, E, p.inner, (t / nt) * LIN_BM,
                                      (t % nt) * LIN_BN, 0, p.inner);
    }
    grid.sync();
    // ---- phase 5: FeedForward.net.0 + exact GELU (few_shot.py:1646-1648)
    {
        const int nt = (p.F + LIN_BN - 1) / LIN_BN;
        for (int t = blockIdx.x; t < mt * nt; t += gridDim.x)
            linear_f32_tile<LIN_GELU, false>(sm, p.y, p.w_fc, p.b_fc, nullptr, p.hid, nullptr, rows, p.F, E, (t / nt) * LIN_BM,
                                      (t % nt) * LIN_BN, 0, E);
    }
    grid.sync();
    // ---- phase 6: FeedForward.net.3, split four ways along K: partial sums
    {
        const int nt = (E + LIN_BN - 1) / LIN_BN, klen = p.F / MODF_KSPLIT;
        for (int it = blockIdx.x; it < mt * nt * MODF_KSPLIT; it += gridDim.x) {
            const int ks = it % MODF_KSPLIT, t = it / MODF_KSPLIT;
            linear_f32_tile<LIN_NONE, false>(sm, p.hid, p.w_proj, nullptr, nullptr, nullptr, p.part + (size_t)ks * rows * E, rows, E, p.F,
                                      (t / nt) * LIN_BM, (t % nt) * LIN_BN, ks * klen, klen);
        }
    }
    grid.sync();
    // ---- phase 7: out = ((p0 + p1) + (p2 + p3)) + b_proj + y   (few_shot.py:1654, 993)
    {
        const size_t n4 = (size_t)rows * E / 4, stride = (size_t)rows * E / 4;
        const float4* part = reinterpret_cast<const float4*>(p.part);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 a = part[i], b = part[i + stride], c = part[i + 2 * stride], d = part[i + 3 * stride];
            const float4 bias = __ldg(reinterpret_cast<const float4*>(p.b_proj) + (i % (size_t)(E / 4)));
            const float4 r = reinterpret_cast<const float4*>(p.y)[i];
            float4 o;
            o.x = ((a.x + b.x) + (c.x + d.x)) + bias.x + r.x;
            o.y = ((a.y + b.y) + (c.y + d.y)) + bias.y + r.y;
            o.z = ((a.z + b.z) + (c.z + d.z)) + bias.z + r.z;
            o.w = ((a.w + b.w) + (c.w + d.w)) + bias.w + r.w;
            reinterpret_cast<float4*>(p.out)[i] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Prototypes [way, T, E] from the modulated support sequences (first T tokens of each, few_shot.py:2956);
// without MERGE_BEFORE the per-class mean over shots happens here (2959-2962).
__global__ void __launch_bounds__(128)
prototype_kernel(const float* __restrict__ mod_out, int q_rows, int n_sup_seq, int T, int E, const int* __restrict__ cls,
                 const int* __restrict__ counts, int merge_before, float* __restrict__ protos) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    const int c = blockIdx.x / T, t = blockIdx.x - c * T;
    float* dst = protos + ((size_t)c * T + t) * E;
    if (merge_before) {
        const float* src = mod_out + ((size_t)q_rows + (size_t)c * (T + 1) + t) * E;
        for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = src[e];
    } else {
        const float inv = 1.0f / float(counts[c]);
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            float a = 0.f;
            for (int s = 0; s < n_sup_seq; ++s)
                if (cls[s] == c) a += mod_out[((size_t)q_rows + (size_t)s * (T + 1) + t) * E + e];
            dst[e] = a * inv;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Caller-side metrics of the episodic eval loop (runs/test_net_few_shot.py:111, 147-160; utils/metrics.py:100-138) as one
// device kernel with deferred host read: per query top-1 correctness (first maximum wins, like torch.topk) and the
// cross-entropy against target_labels, accumulated into int64 counters
//   counters[0] += n_correct, counters[1] += Q, counters[2] += sum_q round(CE_q * 1e6)
// plus optional per-class hit / count tables. Integer atomics: the result does not depend on the order of arrival.
__global__ void metrics_kernel(const float* __restrict__ logits /*[Q,way]*/, const float* __restrict__ target_labels,
                               int Q, int way, unsigned long long* __restrict__ counters,
                               unsigned long long* __restrict__ per_class /*[2*way] hits | counts, or null*/) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const float* row = logits + (size_t)q * way;
    float mx = row[0];
    int arg = 0;
    for (int c = 1; c < way; ++c)
        if (row[c] > mx) { mx = row[c]; arg = c; }
    float sum = 0.f;
    for (int c = 0; c < way; ++c) sum += expf(row[c] - mx);
    const int tgt = (int)((long long)target_labels[q]);
    const bool in_range = tgt >= 0 && tgt < way;
    const float ce = in_range ? (logf(sum) + mx - row[tgt]) : 0.f;
    const bool hit = in_range && arg == tgt;
    atomicAdd(&counters[0], hit ? 1ull : 0ull);
    atomicAdd(&counters[1], 1ull);
    atomicAdd(&counters[2], (unsigned long long)llrintf(ce * 1.0e6f));
    if (per_class != nullptr && in_range) {
        atomicAdd(&per_class[tgt], hit ? 1ull : 0ull);
        atomicAdd(&per_class[way + tgt], 1ull);
    }
}

// ------------------------------------------------------------------------------------------------
// Text branches of the eval forward (few_shot.py:2835-2930). One CTA per query video.
//   p_text[q, c] = softmax_c( scale * <img_q / |img_q|, txt_c / |txt_c|> )                      (2841-2849 / 2862-2870)
//     img_q = mean_T(target_features[q]) (raw ViT features), txt_c = mean over the shots of class c of
//     text_features_test[real_support_labels[s]]
//   mode 1 (TRAIN.EVAL_TEXT): logits = p_text                                                   (2851, 2989)
//   mode 2 (TRAIN.COMBINE)  : logits = p_text^a * softmax_c((8 - cum_dists_visual) / 8)^(1 - a)  (2921-2928), a = TEXT_COFF
constexpr int TEXT_MAX_WAY = 64;
__global__ void __launch_bounds__(256)
text_fusion_kernel(const float* __restrict__ tgt /*[Q,T,E]*/, const float* __restrict__ text_test,
                   const float* __restrict__ real_labels, const int* __restrict__ cls, const int* __restrict__ counts,
                   int S, int T, int E, int way, int n_text, const float* __restrict__ scale, int mode, float text_coff,
                   const float* __restrict__ cum_visual /*[Q,way] or null*/, float* __restrict__ logits /*[Q,way]*/) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    extern __shared__ float sm[];   // img[E]
    __shared__ float red[8];
    __shared__ float lg[TEXT_MAX_WAY];
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* src = tgt + (size_t)q * T * E;
    float sq = 0.f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float a = 0.f;
        for (int t = 0; t < T; ++t) a += src[(size_t)t * E + e];
        a /= float(T);
        sm[e] = a;
        sq += a * a;
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float xn = 0.f;
    for (int w = 0; w < nw; ++w) xn += red[w];
    xn = sqrtf(xn);
    for (int c = warp; c < way; c += nw) {
        const float inv = 1.0f / float(counts[c]);
        float dot = 0.f, tn = 0.f;
        for (int e = lane; e < E; e += 32) {
            float tv = 0.f;
            for (int s = 0; s < S; ++s)
                if (cls[s] == c) tv += text_test[(size_t)text_row(real_labels[s], n_text) * E + e];
            tv *= inv;
            dot = fmaf(sm[e], tv, dot);
            tn = fmaf(tv, tv, tn);
        }
        dot = warp_sum(dot);
        tn = sqrtf(warp_sum(tn));
        if (lane == 0) lg[c] = scale[0] * (dot / xn / tn);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = -INFINITY;
        for (int c = 0; c < way; ++c) mx = fmaxf(mx, lg[c]);
        float sum = 0.f;
        for (int c = 0; c < way; ++c) { lg[c] = expf(lg[c] - mx); sum += lg[c]; }
        float vmx = -INFINITY, vsum = 0.f;
        if (mode == 2) {
            for (int c = 0; c < way; ++c) vmx = fmaxf(vmx, (8.0f - cum_visual[(size_t)q * way + c]) / 8.0f);
            for (int c = 0; c < way; ++c) vsum += expf((8.0f - cum_visual[(size_t)q * way + c]) / 8.0f - vmx);
        }
        for (int c = 0; c < way; ++c) {
            const float pt = lg[c] / sum;
            float out = pt;
            if (mode == 2) {
                const float pv = expf((8.0f - cum_visual[(size_t)q * way + c]) / 8.0f - vmx) / vsum;
                out = powf(pt, text_coff) * powf(pv, 1.0f - text_coff);
            }
            logits[(size_t)q * way + c] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cosine distances + OTAM.  One CTA per (query, class):
//   dists[i][j] = 1 - q_i . p_j / (|q_i| |p_j| + 0.01)                      (cos_sim 1115-1124, 2973-2976)
//   cum = OTAM(dists) (+ OTAM(dists^T) unless single_direct)                 (2657-2687, 2979-2982)
//   logits[q][c] = -cum                                                      (2986-2989)
// The DP runs as a wavefront: lane l owns row l; at step t it fills padded column m = t - l.
constexpr int OTAM_MAX_T = 32;

__device__ __forceinline__ float otam_wavefront(const float* d /*[T][T] smem, this direction: d[l*sl + m*sm]*/,
                                                int sl, int sm_, int T, float lbda, float* c /*[T][T+2] smem*/,
                                                int lane) {
    const int W = T + 2;
    if (lane < T) c[lane * W] = 0.f;  // column 0 stays 0
    __syncwarp();
    for (int t = 1; t <= (T - 1) + (T + 1); ++t) {
        const int l = lane, m = t - l;
        if (l < T && m >= 1 && m <= T + 1) {
            const float dv = (m <= T) ? d[l * sl + (m - 1) * sm_] : 0.f;
            float val;
            if (l == 0) {
                val = dv + c[m - 1];
            } else {
                const float* up = c + (l - 1) * W;
                const float* cur = c + l * W;
                float sum;
                if (m == 1 || m == T + 1)
                    sum = expf(-up[m - 1] / lbda) + expf(-up[m] / lbda) + expf(-cur[m - 1] / lbda);
                else
                    sum = expf(-up[m - 1] / lbda) + expf(-cur[m - 1] / lbda);
                val = dv - lbda * logf(sum);
            }
            c[l * W + m] = val;
        }
        __syncwarp();
    }
    return c[(T - 1) * W + T + 1];
}

__global__ void __launch_bounds__(256)
cos_otam_kernel(const float* __restrict__ qf /*[Q,T,E]*/, const float* __restrict__ pf /*[way,T,E]*/, int T, int E,
                int way, float lbda, int single_direct, float* __restrict__ logits /*[Q,way]*/,
                float* __restrict__ dists_out /*[Q,way,T,T] or null*/, float* __restrict__ cum_out /*[Q,way] or null*/) {
    pdl_trigger();   // programmatic dependent launch: see ptx.cuh
    pdl_wait();
    __shared__ float sd[OTAM_MAX_T * OTAM_MAX_T];
    __shared__ float sqn[OTAM_MAX_T], spn[OTAM_MAX_T];
    __shared__ float sc[2][OTAM_MAX_T * (OTAM_MAX_T + 2)];
    __shared__ float sres[2];
    const int qi = blockIdx.x, ci = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* qb = qf + (size_t)qi * T * E;
    const float* pb = pf + (size_t)ci * T * E;
    // norms
    for (int i = warp; i < 2 * T; i += nw) {
        const float* r = (i < T) ? qb + (size_t)i * E : pb + (size_t)(i - T) * E;
        float s = 0.f;
        for (int e = lane; e < E; e += 32) s = fmaf(r[e], r[e], s);
        s = warp_sum(s);
        if (lane == 0) { if (i < T) sqn[i] = sqrtf(s); else spn[i - T] = sqrtf(s); }
    }
    __syncthreads();
    for (int ij = warp; ij < T * T; ij += nw) {
        const int i = ij / T, j = ij - i * T;
        const float* a = qb + (size_t)i * E;
        const float* b = pb + (size_t)j * E;
        float s = 0.f;
        for (int e = lane; e < E; e += 32) s = fmaf(a[e], b[e], s);
        s = warp_sum(s);
        if (lane == 0) {
            const float dist = 1.0f - s / (sqn[i] * spn[j] + 0.01f);
            sd[ij] = dist;
            if (dists_out != nullptr) dists_out[((size_t)qi * way + ci) * T * T + ij] = dist;
        }
    }
    __syncthreads();
    if (warp == 0) {
        const float r = otam_wavefront(sd, T, 1, T, lbda, sc[0], lane);   // dists[l][m]
        if (lane == 0) sres[0] = r;
    } else if (warp == 1 && !single_direct) {
        const float r = otam_wavefront(sd, 1, T, T, lbda, sc[1], lane);   // dists^T
        if (lane == 0) sres[1] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float cum = single_direct ? sres[0] : sres[0] + sres[1];
        if (cum_out != nullptr) cum_out[(size_t)qi * way + ci] = cum;
        logits[(size_t)qi * way + ci] = -cum;
    }
}

}  // namespace fsar
