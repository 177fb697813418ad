// Kernels of the CLIP text tower that are not shared with the frame encoder.
// Reference semantics: CLIP.encode_text, /root/reference/models/base/few_shot.py:793-806
//   x = token_embedding(text) + positional_embedding          (794-796)
//   x = transformer(x)  -- 12 ResidualAttentionBlocks with the causal mask of build_attention_mask (777-783)
//   x = ln_final(x)[arange(n), text.argmax(-1)] @ text_projection   (800-804)
// The transformer blocks themselves run on the frame encoder's kernels (LayerNorm, tcgen05 GEMMs, tcgen05 attention
// with CAUSAL = true).
#pragma once
#include "ptx.cuh"

namespace fsar {

// x[row, :] = token_embedding[tokens[row], :] + positional_embedding[row % C, :]; one warp per row, float4.
// A token id outside [0, vocab) (corrupt input) is clamped so the gather can never leave the table.
__global__ void __launch_bounds__(256)
text_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ tok_emb, const float* __restrict__ pos,
                  float* __restrict__ x, int rows, int C, int W, int vocab) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= rows) return;
    int tok = tokens[row];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    const float4* e = reinterpret_cast<const float4*>(tok_emb + (size_t)tok * W);
    const float4* p = reinterpret_cast<const float4*>(pos + (size_t)(row % C) * W);
    float4* o = reinterpret_cast<float4*>(x + (size_t)row * W);
    for (int i = lane; i < W / 4; i += 32) {
        const float4 a = __ldg(e + i), b = __ldg(p + i);
        o[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

// out[i, :] = ln_final(x[i, eot_i, :]) @ text_projection, eot_i = first position of the largest token id of text i
// (torch.argmax; the end-of-text token has the highest id). One CTA per text.
__global__ void __launch_bounds__(256)
text_final_kernel(const float* __restrict__ x, const int* __restrict__ tokens, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const float* __restrict__ proj /*[W, E]*/, float* __restrict__ out,
                  int C, int W, int E, float eps) {
    extern __shared__ float row[];   // [W]
    __shared__ float red[8];
    __shared__ int eot;
    const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        int best = -2147483647 - 1, pos = 0;
        for (int t = lane; t < C; t += 32) {
            const int v = tokens[(size_t)i * C + t];
            if (v > best) { best = v; pos = t; }      // ascending t per lane: keeps the first maximum
        }
        for (int o = 16; o > 0; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, pos, o);
            if (ob > best || (ob == best && op < pos)) { best = ob; pos = op; }
        }
        if (lane == 0) eot = pos;
    }
    __syncthreads();
    const float* xr = x + ((size_t)i * C + eot) * W;
    float s = 0.f;
    for (int d = threadIdx.x; d < W; d += blockDim.x) { row[d] = xr[d]; s += row[d]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < 8; ++w) mean += red[w];
    mean /= float(W);
    __syncthreads();
    float q = 0.f;
    for (int d = threadIdx.x; d < W; d += blockDim.x) { const float a = row[d] - mean; q += a * a; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    float var = 0.f;
    for (int w = 0; w < 8; ++w) var += red[w];
    const float rstd = 1.0f / sqrtf(var / float(W) + eps);
    for (int d = threadIdx.x; d < W; d += blockDim.x) row[d] = (row[d] - mean) * rstd * gamma[d] + beta[d];
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float acc = 0.f;
        for (int d = 0; d < W; ++d) acc = fmaf(row[d], __ldg(proj + (size_t)d * E + e), acc);
        out[(size_t)i * E + e] = acc;
    }
}

}  // namespace fsar
