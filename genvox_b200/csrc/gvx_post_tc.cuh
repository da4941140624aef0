// Streaming post-passes of the attention BPTT for the fused persistent chains (bf16 mode, att_dim 128, 32 location filters,
// 31 taps): the two parameter gradients that are sums over EVERY (frame, row, token) of the step
//     d W_loc_dense[d, f] = sum_{t,b,n} d s[t,b,n,d] * conv[t,b,n,f]        (+ d v[d] = sum d e * tanh,  d pm[b,n,d] = sum_t d s)
//     d W_loc_conv[f,c,k] = sum_{t,b,n} d conv[t,b,n,f] * wcat[t,b,c,n+k-15]
// (/root/reference/models/tts/tacotron2.py:186-207: location_layer conv + dense, v; their autograd) are pure streams over the
// stashes - 3.0 GB and 0.55 GB at B = 64, N = 150, T = 800 - with ~30 flop per byte of tensor-core work, so the only thing
// that matters is keeping HBM busy:
//   * every global read is a 16-byte cp.async of a row that is CONTIGUOUS in the stash ([T][B][N][.] layouts: the 16 tokens
//     of a tile / the N tokens of a (frame, row) are adjacent), 3-4 stages deep per CTA, 4-8 CTAs per SM;
//   * the tanh stash is already bf16 with its 16-byte chunks swizzled by (token & 7) (k_att_chain_fwd), so the raw copy lands
//     conflict-free for ldmatrix.trans, which delivers the [dim][token] A fragments of mma.sync.m16n8k16 directly; d s is
//     formed ON the fragment registers (same elements every frame: the d pm sums are 16 registers per thread);
//   * d conv is written by k_att_chain_bwd as bf16 rows (64 B per token) for the same reason.
// Operands are rounded to bf16 like every contraction of bf16 mode (the cumulative attention weights, which can reach
// several units, are split hi + lo so that their rounding does not enter); accumulation, d pm and d v are fp32; partials are
// per CTA and reduced in a fixed order afterwards (deterministic).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "gvx_common.cuh"

namespace gvx {

__device__ __forceinline__ void pt_cp_async(void *smem_dst, const void *gsrc, bool valid, int bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (bytes == 8) {
        const int sz = valid ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
    } else {
        const int sz = valid ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
    }
}
__device__ __forceinline__ void pt_ldsm_x4_t(uint32_t (&r)[4], const void *p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
}
__device__ __forceinline__ void pt_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pt_pack(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ float pt_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float pt_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

// ---------------------------------------------------------------------------------------------------------------------------
// d W_loc_dense, d v, d pm.  Work item = (row b, tile of 16 tokens); the CTA walks the T frames of its tile.
constexpr int PT_TOK = 16, PT_DSTAGES = 4, PT_CLD = 36;      // conv rows padded to 36 floats: fragment reads conflict-free
struct PtDenseStage {
    uint8_t th[PT_TOK * 256];            // [token][16 chunks of 8 dims], chunk index ^ (token & 7)
    float conv[PT_TOK * PT_CLD];
    float de[PT_TOK];
};
__global__ void __launch_bounds__(128) k_post_dense_stream(const uint16_t *__restrict__ THB, const float *__restrict__ DE,
                                                           const float *__restrict__ CONVS, const float *__restrict__ v, int T, int B,
                                                           int N, float *__restrict__ DPM,
                                                           float *__restrict__ part /* [grid][128*32 + 128] */) {
    constexpr int D = 128, F = 32;
    __shared__ __align__(128) PtDenseStage st[PT_DSTAGES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    const int tiles = (N + PT_TOK - 1) / PT_TOK;
    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
    float vacc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    float vv[2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        vv[m][0] = v[32 * warp + 16 * m + g];
        vv[m][1] = v[32 * warp + 16 * m + g + 8];
    }
    // ldmatrix source of this lane: matrix q = lane >> 3 -> (token half q >> 1, chunk c0 + (q & 1)), row = lane & 7
    const int lm_tok = ((lane >> 4) << 3) + (lane & 7), lm_c = (lane >> 3) & 1;
    for (int item = blockIdx.x; item < B * tiles; item += gridDim.x) {
        const int b = item / tiles, n0 = (item - b * tiles) * PT_TOK;
        float pacc[2][4][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int r = 0; r < 4; ++r) pacc[m][r][0] = pacc[m][r][1] = 0.f;
        auto issue = [&](int t) {
            if (t < T) {
                PtDenseStage &s = st[t % PT_DSTAGES];
                const size_t row0 = ((size_t)t * B + b) * N + n0;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int i = tid + 128 * k, tok = i >> 4, ch = i & 15;
                    cp_async16(s.th + tok * 256 + ch * 16, THB + (row0 + tok) * D + ch * 8, n0 + tok < N);
                }
                {
                    const int tok = tid >> 3, ch = tid & 7;
                    cp_async16(s.conv + tok * PT_CLD + ch * 4, CONVS + (row0 + tok) * F + ch * 4, n0 + tok < N);
                }
                if (tid < PT_TOK) pt_cp_async(s.de + tid, DE + row0 + tid, n0 + tid < N, 4);
            }
            cp_async_commit();
        };
        __syncthreads();                 // the previous item's last stages are consumed
#pragma unroll
        for (int s = 0; s < PT_DSTAGES - 1; ++s) issue(s);
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PT_DSTAGES - 2>();
            __syncthreads();
            issue(t + PT_DSTAGES - 1);
            const PtDenseStage &s = st[t % PT_DSTAGES];
            const float2 dlo = *reinterpret_cast<const float2 *>(s.de + 2 * tig), dhi = *reinterpret_cast<const float2 *>(s.de + 2 * tig + 8);
            uint32_t bfr[4][2];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const float *c = s.conv + 2 * tig * PT_CLD + 8 * n + g;
                bfr[n][0] = pt_pack(c[0], c[PT_CLD]);
                bfr[n][1] = pt_pack(c[8 * PT_CLD], c[9 * PT_CLD]);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                uint32_t a[4];
                const int ch = 4 * warp + 2 * m + lm_c;
                pt_ldsm_x4_t(a, s.th + lm_tok * 256 + ((ch ^ (lm_tok & 7)) << 4));
                // a[0]: tokens (2 tig, 2 tig + 1), dim g;  a[1]: same tokens, dim g + 8;  a[2], a[3]: tokens + 8
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float2 de2 = r < 2 ? dlo : dhi;
                    const float vd = vv[m][r & 1];
                    const float t0 = pt_lo(a[r]), t1 = pt_hi(a[r]);
                    const float s0 = de2.x * vd * (1.f - t0 * t0), s1 = de2.y * vd * (1.f - t1 * t1);
                    pacc[m][r][0] += s0;
                    pacc[m][r][1] += s1;
                    vacc[m][r & 1] = fmaf(de2.x, t0, vacc[m][r & 1]);
                    vacc[m][r & 1] = fmaf(de2.y, t1, vacc[m][r & 1]);
                    a[r] = pt_pack(s0, s1);
                }
#pragma unroll
                for (int n = 0; n < 4; ++n) pt_mma(acc[m][n], a, bfr[n][0], bfr[n][1]);
            }
        }
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int d = 32 * warp + 16 * m + g + 8 * (r & 1), tok = 2 * tig + 8 * (r >> 1);
                if (n0 + tok < N) DPM[((size_t)b * N + n0 + tok) * D + d] = pacc[m][r][0];
                if (n0 + tok + 1 < N) DPM[((size_t)b * N + n0 + tok + 1) * D + d] = pacc[m][r][1];
            }
    }
    float *p = part + (size_t)blockIdx.x * (D * F + D);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int r0 = 32 * warp + 16 * m + g, c0 = 8 * n + 2 * tig;
            *reinterpret_cast<float2 *>(p + r0 * F + c0) = make_float2(acc[m][n][0], acc[m][n][1]);
            *reinterpret_cast<float2 *>(p + (r0 + 8) * F + c0) = make_float2(acc[m][n][2], acc[m][n][3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float x = vacc[m][h];
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            x += __shfl_xor_sync(0xffffffffu, x, 2);
            if (tig == 0) p[D * F + 32 * warp + 16 * m + g + 8 * h] = x;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// d W_loc_conv.  Work item = (frame t, row b): [32 filters] x [2 channels x 32 taps (31 used)] += d conv^T [32 x N] . Toeplitz
// window of (alignment of step t-1 | cumulative weights before step t) [N x 62]; the window operand is formed from the two
// zero-padded fp32 rows in shared memory (hi + lo bf16 split), the d conv operand comes through ldmatrix.trans.
constexpr int PT_CSTAGES = 3, PT_DCLD = 80;      // d conv rows padded 64 -> 80 bytes: ldmatrix rows hit 8 distinct bank groups
struct PtConvGeom {
    int NT, WCL, stage_bytes;
    __host__ __device__ explicit PtConvGeom(int N) {
        NT = (N + 15) & ~15;
        WCL = NT + 48;                   // index = token + 16; 16 zeros in front, >= 32 behind
        stage_bytes = NT * PT_DCLD + 2 * WCL * 4;
    }
};
__global__ void __launch_bounds__(128) k_post_conv_stream(const uint16_t *__restrict__ DCB /* [T][B][N][32] bf16 */,
                                                          const float *__restrict__ ALIGN, const float *__restrict__ CUMS, int T, int B,
                                                          int N, float *__restrict__ part /* [grid][32*2*31] */) {
    constexpr int F = 32, KS = 31;
    extern __shared__ __align__(128) uint8_t pt_sm[];
    const PtConvGeom G(N);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    for (int i = tid; i < PT_CSTAGES * G.stage_bytes / 4; i += 128) reinterpret_cast<uint32_t *>(pt_sm)[i] = 0u;
    __syncthreads();
    float acc[2][2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
    const int c = warp >> 1, kbase = 16 * (warp & 1) + g;           // columns of this warp: channel c, taps kbase, kbase + 8
    const int lm_tok = ((lane >> 4) << 3) + (lane & 7), lm_c = (lane >> 3) & 1;
    const int nitems = T * B;
    const int wb = (N & 1) ? 4 : 8, wchunks = (N & 1) ? N : N / 2;
    auto issue = [&](int k) {
        const int item = blockIdx.x + k * gridDim.x;
        if (item < nitems) {
            uint8_t *s = pt_sm + (size_t)(k % PT_CSTAGES) * G.stage_bytes;
            float *wc = reinterpret_cast<float *>(s + G.NT * PT_DCLD);
            const int t = item / B, b = item - t * B;
            const uint16_t *src = DCB + ((size_t)t * B + b) * N * F;
            for (int i = tid; i < N * 4; i += 128) cp_async16(s + (i >> 2) * PT_DCLD + (i & 3) * 16, src + i * 8, true);
            const float *al = ALIGN + ((size_t)b * T + (t > 0 ? t - 1 : 0)) * N, *cu = CUMS + ((size_t)b * T + t) * N;
            for (int i = tid; i < 2 * wchunks; i += 128) {
                const int cc = i >= wchunks, j = cc ? i - wchunks : i;
                const int e = j * (wb / 4);
                pt_cp_async(wc + cc * G.WCL + 16 + e, (cc ? cu : al) + e, cc || t > 0, wb);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < PT_CSTAGES - 1; ++k) issue(k);
    for (int k = 0; blockIdx.x + k * gridDim.x < nitems; ++k) {
        cp_async_wait<PT_CSTAGES - 2>();
        __syncthreads();
        issue(k + PT_CSTAGES - 1);
        const uint8_t *s = pt_sm + (size_t)(k % PT_CSTAGES) * G.stage_bytes;
        const float *wc = reinterpret_cast<const float *>(s + G.NT * PT_DCLD) + c * G.WCL;
        for (int ks = 0; ks < G.NT / 16; ++ks) {
            uint32_t a[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) pt_ldsm_x4_t(a[m], s + (16 * ks + lm_tok) * PT_DCLD + ((2 * m + lm_c) << 4));
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const float *x = wc + 16 * ks + 2 * tig + kbase + 8 * jj + 1;     // wcat[c][token + tap - 15], token = 16 ks + 2 tig
                const float x0 = x[0], x1 = x[1], x8 = x[8], x9 = x[9];
                const uint32_t b0 = pt_pack(x0, x1), b1 = pt_pack(x8, x9);
                const uint32_t l0 = pt_pack(x0 - pt_lo(b0), x1 - pt_hi(b0)), l1 = pt_pack(x8 - pt_lo(b1), x9 - pt_hi(b1));
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    pt_mma(acc[m][jj], a[m], b0, b1);
                    pt_mma(acc[m][jj], a[m], l0, l1);
                }
            }
        }
    }
    cp_async_wait<0>();
    float *p = part + (size_t)blockIdx.x * (F * 2 * KS);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = 16 * m + g + 8 * (i >> 1), k = 16 * (warp & 1) + 8 * jj + 2 * tig + (i & 1);
                if (k < KS) p[(f * 2 + c) * KS + k] = acc[m][jj][i];
            }
}

}  // namespace gvx
