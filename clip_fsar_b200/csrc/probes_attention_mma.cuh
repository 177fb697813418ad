// PROBES BUILD ONLY (-DFSAR_PROBES, libfsar_sm100_probes.so): the round-1 warp-level mma.sync attention core, kept as an
// A/B baseline for tools/gemm_probe.py. It is NOT part of the product library: libfsar_sm100.so contains only the
// tcgen05 / TMEM attention core (attention_tcgen05.cuh) and rejects L > 257.
#pragma once
#include "vit_kernels.cuh"

namespace fsar {

// ------------------------------------------------------------------------------------------------
// Attention core, register-resident softmax (warp-level mma.sync m16n8k16, fp32 accumulate).
//   qkv16 [n_frames * L, 3 D] (row = frame * L + token; Q | K | V column blocks, head h at h * 64)
//   out16 [n_frames * L, D]
// One CTA = (64 query rows, head, frame); 4 warps x 16 query rows. Whole K/V of the (frame, head)
// is staged in shared memory (L <= 16 * NKT keys), scores never touch HBM.
// softmax(QK^T / sqrt(64)) V, no mask, no dropout (nn.MultiheadAttention in eval, few_shot.py:623,635).
template <typename T16>
struct MmaOp;
template <>
struct MmaOp<__half> {
    __device__ static __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <>
struct MmaOp<__nv_bfloat16> {
    __device__ static __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

constexpr int ATT_PITCH = 72;     // smem row pitch in 16-bit elements (144 B: conflict-free ldmatrix)
constexpr int ATT_QROWS = 64;     // query rows per CTA

template <int NKT>
constexpr int att_smem_bytes() {
    return (2 * NKT * 16 + ATT_QROWS) * ATT_PITCH * 2;
}

template <typename T16, int NKT>  // NKT = number of 16-key tiles staged (L <= 16 * NKT)
__global__ void __launch_bounds__(128)
attention_mma_kernel(const T16* __restrict__ qkv, T16* __restrict__ out, int L, int D, float scale_log2e) {
    constexpr int LP = NKT * 16;
    extern __shared__ __align__(16) uint8_t att_smem[];
    T16* sK = reinterpret_cast<T16*>(att_smem);
    T16* sV = sK + LP * ATT_PITCH;
    T16* sQ = sV + LP * ATT_PITCH;

    pdl_trigger();
    pdl_wait();
    const int qblk = blockIdx.x, head = blockIdx.y, frame = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t row0 = (size_t)frame * L;
    const int ld = 3 * D;
    const T16* gq = qkv + row0 * ld + head * ATT_HD;
    const T16* gk = gq + D;
    const T16* gv = gq + 2 * D;

    // ---- stage K, V (all keys) and this CTA's 64 query rows; rows >= L are zero
    for (int i = threadIdx.x; i < LP * 8; i += blockDim.x) {
        const int r = i >> 3, c = (i & 7) * 8;
        uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (r < L) {
            kk = __ldg(reinterpret_cast<const uint4*>(gk + (size_t)r * ld + c));
            vv = __ldg(reinterpret_cast<const uint4*>(gv + (size_t)r * ld + c));
        }
        *reinterpret_cast<uint4*>(sK + r * ATT_PITCH + c) = kk;
        *reinterpret_cast<uint4*>(sV + r * ATT_PITCH + c) = vv;
    }
    for (int i = threadIdx.x; i < ATT_QROWS * 8; i += blockDim.x) {
        const int r = i >> 3, c = (i & 7) * 8;
        const int qr = qblk * ATT_QROWS + r;
        uint4 qq = make_uint4(0, 0, 0, 0);
        if (qr < L) qq = __ldg(reinterpret_cast<const uint4*>(gq + (size_t)qr * ld + c));
        *reinterpret_cast<uint4*>(sQ + r * ATT_PITCH + c) = qq;
    }
    __syncthreads();

    const int qrow_w = qblk * ATT_QROWS + warp * 16;  // first query row of this warp
    if (qrow_w >= L) return;

    // ---- Q fragments (A operand, 16 x 64): 4 k-steps x 4 regs
    uint32_t qf[4][4];
    {
        const int r = warp * 16 + (lane & 15);
        const int cofs = (lane >> 4) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(qf[kk], smem_u32(sQ + r * ATT_PITCH + kk * 16 + cofs));
    }

    // ---- S = Q K^T : NKT*2 n-tiles of 8 keys
    float s[NKT * 2][4];
#pragma unroll
    for (int j = 0; j < NKT * 2; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        uint32_t kb0[4], kb1[4];
        const int key = j * 8 + (lane & 7);
        const int cofs = (lane >> 3) * 8;
        ldmatrix_x4(kb0, smem_u32(sK + key * ATT_PITCH + cofs));        // dims 0..31
        ldmatrix_x4(kb1, smem_u32(sK + key * ATT_PITCH + 32 + cofs));   // dims 32..63
        MmaOp<T16>::mma(s[j], qf[0], kb0[0], kb0[1]);
        MmaOp<T16>::mma(s[j], qf[1], kb0[2], kb0[3]);
        MmaOp<T16>::mma(s[j], qf[2], kb1[0], kb1[1]);
        MmaOp<T16>::mma(s[j], qf[3], kb1[2], kb1[3]);
    }

    // ---- softmax over keys (rows g and g + 8 of this warp's 16), keys >= L masked
    const int t4 = lane & 3;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NKT * 2; ++j) {
        const int k0 = j * 8 + t4 * 2;
        if (k0 >= L) s[j][0] = s[j][2] = -INFINITY;
        if (k0 + 1 >= L) s[j][1] = s[j][3] = -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float m0 = mx0 * scale_log2e, m1 = mx1 * scale_log2e;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < NKT * 2; ++j) {
        s[j][0] = exp2f(s[j][0] * scale_log2e - m0);
        s[j][1] = exp2f(s[j][1] * scale_log2e - m0);
        s[j][2] = exp2f(s[j][2] * scale_log2e - m1);
        s[j][3] = exp2f(s[j][3] * scale_log2e - m1);
        sum0 += s[j][0] + s[j][1];
        sum1 += s[j][2] + s[j][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // ---- O = P V : P fragments come straight from the S accumulators
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NKT; ++kt) {
        uint32_t pa[4];
        pa[0] = pack2<T16>(s[2 * kt][0], s[2 * kt][1]);
        pa[1] = pack2<T16>(s[2 * kt][2], s[2 * kt][3]);
        pa[2] = pack2<T16>(s[2 * kt + 1][0], s[2 * kt + 1][1]);
        pa[3] = pack2<T16>(s[2 * kt + 1][2], s[2 * kt + 1][3]);
        const int key = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int cofs = (lane >> 4) * 8;
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-wide d tiles
            uint32_t vb[4];
            ldmatrix_x4_trans(vb, smem_u32(sV + key * ATT_PITCH + np * 16 + cofs));
            MmaOp<T16>::mma(o[2 * np], pa, vb[0], vb[1]);
            MmaOp<T16>::mma(o[2 * np + 1], pa, vb[2], vb[3]);
        }
    }

    // ---- normalise and store (row g: regs 0,1; row g + 8: regs 2,3)
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    const int g = lane >> 2;
    const int r0 = qrow_w + g, r1 = r0 + 8;
    T16* go = out + row0 * D + head * ATT_HD;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int col = n * 8 + t4 * 2;
        if (r0 < L) *reinterpret_cast<uint32_t*>(go + (size_t)r0 * D + col) = pack2<T16>(o[n][0] * inv0, o[n][1] * inv0);
        if (r1 < L) *reinterpret_cast<uint32_t*>(go + (size_t)r1 * D + col) = pack2<T16>(o[n][2] * inv1, o[n][3] * inv1);
    }
}

}  // namespace fsar
