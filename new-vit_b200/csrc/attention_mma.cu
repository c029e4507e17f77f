// bf16 encoder attention, head_dim 64, short sequences (N = 257 / 325 tokens): softmax(q k^T) v per
// (slice, head) with q pre-scaled (reference layers/attention.py:56-69).  One CTA per (slice, head): K and V
// of the head are staged once in shared memory (XOR-swizzled 16-byte chunks), each warp owns 16-row query
// blocks and runs a register-resident online-softmax loop over 32-key chunks on the warp-level tensor-core
// path (mma.sync m16n8k16, fp32 accumulate).  The 257x257 probability matrix is never written to HBM (the
// reference materialises [BD,6,257,257] fp32 per block).  Round-1 kernel; a tcgen05/TMEM version is the
// planned successor.
#include <math_constants.h>
#include "common.cuh"
#include "ptx.cuh"

namespace mst {

constexpr int ATT_WARPS = 6;
constexpr int ATT_THREADS = ATT_WARPS * 32;

__global__ void __launch_bounds__(ATT_THREADS, 3)
attention_bf16_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int N, int heads, int NKP) {
    extern __shared__ __align__(128) uint8_t smem_att[];
    uint8_t* Ks = smem_att;                 // [NKP][128 B]
    uint8_t* Vs = smem_att + NKP * 128;     // [NKP][128 B]
    const int s = blockIdx.x / heads, h = blockIdx.x % heads;
    const int E = heads * 64;
    const int64_t ld = 3 * E;
    const bf16* base = qkv + static_cast<int64_t>(s) * N * ld + h * 64;
    const uint32_t ks_u32 = smem_u32(Ks), vs_u32 = smem_u32(Vs);

    for (int idx = threadIdx.x; idx < NKP * 8; idx += ATT_THREADS) {
        const int r = idx >> 3, c = idx & 7;
        const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
        if (r < N) {
            const bf16* g = base + r * ld + c * 8;
            cp_async_16(ks_u32 + off, g + E);
            cp_async_16(vs_u32 + off, g + 2 * E);
        } else {
            *reinterpret_cast<uint4*>(Ks + off) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(Vs + off) = make_uint4(0, 0, 0, 0);
        }
    }
    cp_async_commit_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int n_rb = (N + 15) >> 4;
    const int n8_total = (N + 7) >> 3;
    constexpr float LOG2E = 1.4426950408889634f;

    for (int rb = warp; rb < n_rb; rb += ATT_WARPS) {
        // ---- Q fragments (A operand) straight from global: rows rb*16+g and +8 ----
        const int r0 = min(rb * 16 + g, N - 1), r1 = min(rb * 16 + g + 8, N - 1);
        uint32_t qa[4][4];
        {
            const uint32_t* q0 = reinterpret_cast<const uint32_t*>(base + r0 * ld);
            const uint32_t* q1 = reinterpret_cast<const uint32_t*>(base + r1 * ld);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qa[ks][0] = __ldg(q0 + ks * 8 + t);
                qa[ks][1] = __ldg(q1 + ks * 8 + t);
                qa[ks][2] = __ldg(q0 + ks * 8 + 4 + t);
                qa[ks][3] = __ldg(q1 + ks * 8 + 4 + t);
            }
        }
        float oacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f; }
        float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;

        for (int nb0 = 0; nb0 < n8_total; nb0 += 4) {
            const int nblk = min(4, n8_total - nb0);
            float sacc[4][4];
            // ---- S = Q K^T for up to 4 n8 key blocks ----
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                sacc[nb][0] = sacc[nb][1] = sacc[nb][2] = sacc[nb][3] = 0.f;
                if (nb < nblk) {
                    const int key = (nb0 + nb) * 8 + (lane & 7);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c = half * 4 + (lane >> 3);
                        uint32_t b[4];
                        ldmatrix_x4(b, ks_u32 + key * 128 + ((c ^ (key & 7)) << 4));
                        mma_bf16_16816(sacc[nb], qa[2 * half], b[0], b[1]);
                        mma_bf16_16816(sacc[nb], qa[2 * half + 1], b[2], b[3]);
                    }
                }
            }
            // ---- mask keys >= N (only the last n8 block can be partial), chunk max ----
            float cm0 = -CUDART_INF_F, cm1 = -CUDART_INF_F;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                if (nb < nblk) {
                    const int key = (nb0 + nb) * 8 + 2 * t;
                    if (key >= N) { sacc[nb][0] = -CUDART_INF_F; sacc[nb][2] = -CUDART_INF_F; }
                    if (key + 1 >= N) { sacc[nb][1] = -CUDART_INF_F; sacc[nb][3] = -CUDART_INF_F; }
                    cm0 = fmaxf(cm0, fmaxf(sacc[nb][0], sacc[nb][1]));
                    cm1 = fmaxf(cm1, fmaxf(sacc[nb][2], sacc[nb][3]));
                }
            }
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
            const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);  // finite: every chunk holds >= 1 valid key
            const float sc0 = exp2f((m0 - mn0) * LOG2E), sc1 = exp2f((m1 - mn1) * LOG2E);
            m0 = mn0; m1 = mn1;
            const float ms0 = mn0 * LOG2E, ms1 = mn1 * LOG2E;
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                if (nb < nblk) {
                    sacc[nb][0] = exp2f(fmaf(sacc[nb][0], LOG2E, -ms0));
                    sacc[nb][1] = exp2f(fmaf(sacc[nb][1], LOG2E, -ms0));
                    sacc[nb][2] = exp2f(fmaf(sacc[nb][2], LOG2E, -ms1));
                    sacc[nb][3] = exp2f(fmaf(sacc[nb][3], LOG2E, -ms1));
                    rs0 += sacc[nb][0] + sacc[nb][1];
                    rs1 += sacc[nb][2] + sacc[nb][3];
                }
            }
            l0 = fmaf(l0, sc0, rs0);
            l1 = fmaf(l1, sc1, rs1);
#pragma unroll
            for (int i = 0; i < 8; ++i) { oacc[i][0] *= sc0; oacc[i][1] *= sc0; oacc[i][2] *= sc1; oacc[i][3] *= sc1; }
            // ---- O += P V, 16 keys per step (sacc of unused blocks is exactly 0) ----
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                if (2 * kk < nblk) {
                    uint32_t pa[4];
                    pa[0] = pack_bf16x2(sacc[2 * kk][0], sacc[2 * kk][1]);
                    pa[1] = pack_bf16x2(sacc[2 * kk][2], sacc[2 * kk][3]);
                    pa[2] = pack_bf16x2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
                    pa[3] = pack_bf16x2(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
                    const int key = nb0 * 8 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
                    for (int dn = 0; dn < 8; dn += 2) {
                        const int c = dn + (lane >> 4);
                        uint32_t b[4];
                        ldmatrix_x4_trans(b, vs_u32 + key * 128 + ((c ^ (key & 7)) << 4));
                        mma_bf16_16816(oacc[dn], pa, b[0], b[1]);
                        mma_bf16_16816(oacc[dn + 1], pa, b[2], b[3]);
                    }
                }
            }
        }
        // ---- finalise: 1/l, bf16, store ----
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const int row0 = rb * 16 + g, row1 = row0 + 8;
        if (row0 < N) {
            uint32_t* o = reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(s) * N + row0) * E + h * 64);
#pragma unroll
            for (int dn = 0; dn < 8; ++dn) o[dn * 4 + t] = pack_bf16x2(oacc[dn][0] * i0, oacc[dn][1] * i0);
        }
        if (row1 < N) {
            uint32_t* o = reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(s) * N + row1) * E + h * 64);
#pragma unroll
            for (int dn = 0; dn < 8; ++dn) o[dn * 4 + t] = pack_bf16x2(oacc[dn][2] * i1, oacc[dn][3] * i1);
        }
    }
}

int launch_attention_bf16(const bf16* qkv, bf16* out, int BD, int N, int heads, cudaStream_t stream) {
    const int NKP = ((N + 31) / 32) * 32;  // a 32-key chunk may touch rows up to the next multiple of 16; keep 32
    const size_t smem = static_cast<size_t>(NKP) * 256;
    MST_REQUIRE(smem <= 227 * 1024, "attention: N=%d tokens do not fit shared memory", N);
    MST_SET_DYN_SMEM(attention_bf16_kernel, 227 * 1024);
    attention_bf16_kernel<<<BD * heads, ATT_THREADS, smem, stream>>>(qkv, out, N, heads, NKP);
    MST_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mst
