// genvox_b200 — persistent recurrence kernels (bf16 mode): a whole LSTM chain in ONE launch.
//
// nn.LSTMCell's gate pre-activations (/root/reference/models/tts/tacotron2.py:340,:357) split into an input part and a
// recurrent part.  In teacher-forced training the input part of the decoder LSTM ([h_att_t | ctx_t] . W_ih^T + biases)
// does not depend on the decoder LSTM itself, so it is computed for all T frames by one time-batched GEMM and what
// remains on the sequential chain is   gates_t = pre_t + W_hh . h_{t-1}   followed by the cell.  These kernels run that
// chain for all T steps in a single launch:
//
//   * grid = H/8 CTAs (128 for H = 1024), one per SM, all co-resident; CTA j owns hidden units [8j, 8j+8);
//   * its slice of the recurrent weights (32 gate rows x H, bf16, 64 KB) is loaded ONCE into shared memory and stays
//     there for the whole sequence; the cell state (forward) / its gradient (backward) of the CTA's units lives in
//     registers across steps;
//   * per step the hidden state of all units ([64 rows][H] bf16 operand image, written by all CTAs) is streamed in by
//     TMA bulk copies and contracted on tcgen05 (UMMA 128 x 32 x 16, batch rows on the M side so that one TMEM lane =
//     one batch row and the cell math is thread-local), accumulators in TMEM;
//   * steps are separated by a grid-wide barrier (release/acquire counter in global memory).
//
// Backward (BPTT of the same chain): d h_{t-1} = W_hh^T . d gates_t has K = 4H; the K range is split over the 4 CTAs of
// a thread-block cluster (each keeps a [32 units x H] slice of W_hh^T resident), the four partial [64 x 32] tiles are
// exchanged through distributed shared memory and summed in rank order (deterministic), and each CTA finishes the cell
// backward of its own 8 units.
//
// Warp roles (128 threads): warps 0,1 = epilogue (TMEM lanes 0..63 = batch rows), warp 2 = TMA producer,
// warp 3 = MMA issuer (one elected lane each).  Every wait is bounded; on timeout an error code is recorded, the CTA's
// roles stop working (but keep the cluster barriers balanced) and the host raises.
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "gvx_bf16.cuh"
#include "gvx_gemm.cuh"
#include "gvx_tc.cuh"

namespace gvx {

constexpr int PC_ROWS = 64;                               // batch rows of an activation image (B <= 64)
constexpr int PC_N = 32;                                  // UMMA N: 8 units x 4 gates (fwd) / 32 output units (bwd)
constexpr int PC_CHUNK_BYTES = 64 * PC_ROWS * 2;          // one TMA chunk: 64 K elements = 8 k-chunks of [64][8] bf16
constexpr int PC_MAXRING = 16;
constexpr int PC_THREADS = 128;
constexpr long long PC_WAIT_CYCLES = 4000000000ll;        // ~2 s of SM clock

struct PcShared {
    uint64_t full[PC_MAXRING], empty[PC_MAXRING], tmem_full, wbar;
    uint32_t tmem_slot;
    volatile int dead;
};

__device__ __forceinline__ bool pc_mbar_wait(uint64_t *bar, uint32_t parity, volatile int *dead, int *err, int code) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    for (;;) {
        if (mbar_try_wait(bar, parity)) return true;
        if (*dead) return false;
        if (clock64() - t0 > PC_WAIT_CYCLES) {
            *dead = 1;
            atomicExch(err, code);
            return false;
        }
    }
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gbar_arrive(unsigned *ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ bool gbar_wait(const unsigned *ctr, unsigned target, volatile int *dead, int *err, int code) {
    if (ld_acquire_u32(ctr) >= target) return true;
    const long long t0 = clock64();
    for (;;) {
        if (ld_acquire_u32(ctr) >= target) return true;
        if (*dead) return false;
        if (clock64() - t0 > PC_WAIT_CYCLES) {
            *dead = 1;
            atomicExch(err, code);
            return false;
        }
    }
}
// generic-proxy global writes of other CTAs -> async-proxy (TMA) reads: fence on both sides of the barrier
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// row-major bf16 destination of the hidden state of frame t + toff (skipped when t + toff >= T)
struct PcOut {
    __nv_bfloat16 *p;
    int ld, koff, toff;
    long long tstride;
};

// ================================================================================================ forward
struct PcFwdArgs {
    const __nv_bfloat16 *Wimg;   // [H/8][H/8][32][8]   CTA, k-chunk, local gate row 4*lu+g, k in chunk  (W_hh)
    const float *pre;            // [T][B][4H]  unit-major columns 4u+g: input contribution (may alias gates_stash)
    const float *bias;           // [4H] unit-major b_ih + b_hh, or null
    __nv_bfloat16 *himg;         // [2][H/8][64][8]  ping-pong operand image of h; half 0 = h_{-1} (zeros), pad rows zero
    float *c_stash;              // [T+1][B][H]  row 0 = c_{-1} (caller), row t+1 written at step t
    float *gates_stash;          // [T][B][4H] gate activations (i,f,g,o per unit) or null
    PcOut out[2];
    unsigned *bar;               // grid barrier counter, zero at launch
    int *err;
    DropCfg drop;
    uint32_t site;
    int row_offset, B, T, H;
};

__global__ void __launch_bounds__(PC_THREADS, 1) k_lstm_chain_fwd(const PcFwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int H = a.H, T = a.T;
    const int img_bytes = H * 128;                         // one h image: H/8 k-chunks x 1 KB
    const int nchunk = (img_bytes + PC_CHUNK_BYTES - 1) / PC_CHUNK_BYTES;
    const int R = nchunk < PC_MAXRING ? nchunk : PC_MAXRING;
    const uint32_t wbytes = (uint32_t)H * 64;              // 32 rows x H x 2 B
    uint8_t *ring = smem;
    uint8_t *wsm = ring + (size_t)R * PC_CHUNK_BYTES;      // the M = 128 MMA over-reads 1 KB past a chunk: lands here
    PcShared *sh = (PcShared *)(wsm + wbytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, j = blockIdx.x;
    const unsigned ncta = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PC_MAXRING; ++s) { mbar_init(sh->full + s, 1); mbar_init(sh->empty + s, 1); }
        mbar_init(&sh->tmem_full, 1);
        mbar_init(&sh->wbar, 1);
        sh->dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_slot)), "n"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_slot;

    if (warp == 2) {
        // ------------------------------------------------ TMA producer
        if (elect_one()) {
            mbar_expect_tx(&sh->wbar, wbytes);
            const uint8_t *wsrc = (const uint8_t *)a.Wimg + (size_t)j * wbytes;
            for (uint32_t off = 0; off < wbytes; off += 16384) {
                const uint32_t n = wbytes - off < 16384 ? wbytes - off : 16384;
                tma_bulk_g2s(wsm + off, wsrc + off, n, &sh->wbar);
            }
            uint32_t g = 0;
            bool ok = true;
            for (int t = 0; t < T && ok; ++t) {
                if (t > 0) ok = gbar_wait(a.bar, ncta * (unsigned)t, &sh->dead, a.err, 11);
                if (!ok) break;
                fence_proxy_async_all();
                const uint8_t *src = (const uint8_t *)a.himg + (size_t)(t & 1) * img_bytes;
                for (int c = 0; c < nchunk; ++c, ++g) {
                    const int s = g % R;
                    const uint32_t ph = (g / R) & 1u;
                    if (!pc_mbar_wait(sh->empty + s, ph ^ 1u, &sh->dead, a.err, 12)) { ok = false; break; }
                    const int left = img_bytes - c * PC_CHUNK_BYTES;
                    const uint32_t n = left < PC_CHUNK_BYTES ? left : PC_CHUNK_BYTES;
                    mbar_expect_tx(sh->full + s, n);
                    tma_bulk_g2s(ring + (size_t)s * PC_CHUNK_BYTES, src + (size_t)c * PC_CHUNK_BYTES, n, sh->full + s);
                }
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ------------------------------------------------ MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, PC_N);
            const int nk16 = H / 16;
            bool ok = pc_mbar_wait(&sh->wbar, 0, &sh->dead, a.err, 13);
            uint32_t g = 0;
            const uint32_t ring_a = smem_u32(ring), w_a = smem_u32(wsm);
            for (int t = 0; t < T && ok; ++t) {
                for (int c = 0; c < nchunk; ++c, ++g) {
                    const int s = g % R;
                    const uint32_t ph = (g / R) & 1u;
                    if (!pc_mbar_wait(sh->full + s, ph, &sh->dead, a.err, 14)) { ok = false; break; }
                    tc_fence_after();
                    const int k0 = c * 4, k1 = k0 + 4 < nk16 ? k0 + 4 : nk16;
                    for (int q = k0; q < k1; ++q) {
                        // A: h image chunk [8 k-chunks][64][8]: LBO 1 KB (between the two k-chunks of a K=16 MMA), SBO 128 B
                        const uint64_t ad = umma_desc(ring_a + s * PC_CHUNK_BYTES + (q - k0) * 2048, 1024, 128);
                        // B: resident weights [H/8][32][8]: LBO 512 B, SBO 128 B
                        const uint64_t bd = umma_desc(w_a + q * 1024, 512, 128);
                        umma_bf16(tmem_base, ad, bd, idesc, q > 0 ? 1u : 0u);
                    }
                    umma_commit(sh->empty + s);
                }
                if (ok) umma_commit(&sh->tmem_full);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ epilogue: one thread = one batch row, 8 hidden units
        const int b = warp * 32 + lane;
        const bool valid = b < a.B;
        const int u0 = 8 * j;
        float c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = valid ? a.c_stash[(size_t)b * H + u0 + i] : 0.f;
        float4 bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bi[i] = a.bias ? *reinterpret_cast<const float4 *>(a.bias + 4 * (u0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        bool ok = true;
        for (int t = 0; t < T; ++t) {
            float4 pr[8];
            if (valid) {
                const float4 *pp = reinterpret_cast<const float4 *>(a.pre + ((size_t)t * a.B + b) * 4 * H + 4 * u0);
#pragma unroll
                for (int i = 0; i < 8; ++i) pr[i] = __ldcs(pp + i);
            }
            if (ok) ok = __all_sync(0xffffffffu, pc_mbar_wait(&sh->tmem_full, (uint32_t)t & 1u, &sh->dead, a.err, 15)) != 0;
            float acc[32];
            if (ok) {
                tc_fence_after();
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + 16u, acc + 16);
                tc_fence_before();
            }
            if (ok && valid) {
                float hv[8];
                float4 *gs = a.gates_stash ? reinterpret_cast<float4 *>(a.gates_stash + ((size_t)t * a.B + b) * 4 * H + 4 * u0) : nullptr;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float gi = sigmoidf_(acc[4 * i] + pr[i].x + bi[i].x), gf = sigmoidf_(acc[4 * i + 1] + pr[i].y + bi[i].y);
                    const float gg = tanhf(acc[4 * i + 2] + pr[i].z + bi[i].z), go = sigmoidf_(acc[4 * i + 3] + pr[i].w + bi[i].w);
                    const float cn = gf * c[i] + gi * gg;
                    c[i] = cn;
                    hv[i] = go * tanhf(cn) * drop_mult(a.drop, a.site, (uint32_t)t, (uint32_t)(b + a.row_offset), (uint32_t)(u0 + i));
                    if (gs) gs[i] = make_float4(gi, gf, gg, go);
                }
                float4 *cs = reinterpret_cast<float4 *>(a.c_stash + ((size_t)(t + 1) * a.B + b) * H + u0);
                cs[0] = make_float4(c[0], c[1], c[2], c[3]);
                cs[1] = make_float4(c[4], c[5], c[6], c[7]);
                const uint4 hp = make_uint4(pack_bf2(hv[0], hv[1]), pack_bf2(hv[2], hv[3]), pack_bf2(hv[4], hv[5]), pack_bf2(hv[6], hv[7]));
                *reinterpret_cast<uint4 *>(a.himg + (size_t)((t + 1) & 1) * H * 64 + ((size_t)j * PC_ROWS + b) * 8) = hp;
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const PcOut &d = a.out[o];
                    if (d.p && t + d.toff < T)
                        *reinterpret_cast<uint4 *>(d.p + (size_t)(t + d.toff) * d.tstride + (size_t)b * d.ld + d.koff + u0) = hp;
                }
            }
            __threadfence();
            fence_proxy_async_all();
            asm volatile("bar.sync 1, 64;" ::: "memory");
            if (threadIdx.x == 0) gbar_arrive(a.bar);
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
    }
}

// ================================================================================================ backward
struct PcBwdArgs {
    const __nv_bfloat16 *Wimg;   // [H/8][H/8][32][8]   CTA j = 4r+s: row n = output unit 32r+n, k = gate row s*H + kk (unit-major)
    __nv_bfloat16 *gimg;         // [2][4H/8][64][8]  ping-pong d-gates image (k = 4u+g); half 0 zero at launch
    const float *dh_ext;         // d h (dropped) of frame t from everything but the recurrence: dh_ext[t*dh_tstride + b*dh_ld + u]
    int dh_ld;
    long long dh_tstride;
    const float *gates_stash;    // [T][B][4H]
    const float *c_stash;        // [T+1][B][H]
    __nv_bfloat16 *dg_rm;        // [T][B][4H] row-major d gates (columns 4u+g) for the time-batched GEMMs
    unsigned *bar;
    int *err;
    DropCfg drop;
    uint32_t site;
    int row_offset, B, T, H;
};

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(PC_THREADS, 1) k_lstm_chain_bwd(const PcBwdArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int H = a.H, T = a.T;
    const int q_bytes = H * 128;                           // this CTA's K quarter of the d-gates image
    const int nchunk = (q_bytes + PC_CHUNK_BYTES - 1) / PC_CHUNK_BYTES;
    const int R = nchunk < PC_MAXRING ? nchunk : PC_MAXRING;
    const uint32_t wbytes = (uint32_t)H * 64;
    uint8_t *ring = smem;
    uint8_t *wsm = ring + (size_t)R * PC_CHUNK_BYTES;
    float *part = (float *)(wsm + wbytes);                 // [32 units][64 rows] fp32 partial of this K quarter
    PcShared *sh = (PcShared *)(part + PC_N * PC_ROWS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, j = blockIdx.x;
    const int s_rank = (int)cluster.block_rank();          // == j & 3
    const unsigned ncta = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PC_MAXRING; ++s) { mbar_init(sh->full + s, 1); mbar_init(sh->empty + s, 1); }
        mbar_init(&sh->tmem_full, 1);
        mbar_init(&sh->wbar, 1);
        sh->dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_slot)), "n"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_slot;
    cluster.sync();

    if (warp == 2) {
        // ------------------------------------------------ TMA producer
        const bool leader = elect_one();
        uint32_t g = 0;
        bool ok = true;
        if (leader) {
            mbar_expect_tx(&sh->wbar, wbytes);
            const uint8_t *wsrc = (const uint8_t *)a.Wimg + (size_t)j * wbytes;
            for (uint32_t off = 0; off < wbytes; off += 16384) {
                const uint32_t n = wbytes - off < 16384 ? wbytes - off : 16384;
                tma_bulk_g2s(wsm + off, wsrc + off, n, &sh->wbar);
            }
        }
        for (int i = 0; i < T; ++i) {
            if (leader && ok) {
                if (i > 0) ok = gbar_wait(a.bar, ncta * (unsigned)i, &sh->dead, a.err, 21);
                if (ok) {
                    fence_proxy_async_all();
                    const uint8_t *src = (const uint8_t *)a.gimg + (size_t)(i & 1) * 4 * q_bytes + (size_t)s_rank * q_bytes;
                    for (int c = 0; c < nchunk; ++c, ++g) {
                        const int s = g % R;
                        const uint32_t ph = (g / R) & 1u;
                        if (!pc_mbar_wait(sh->empty + s, ph ^ 1u, &sh->dead, a.err, 22)) { ok = false; break; }
                        const int left = q_bytes - c * PC_CHUNK_BYTES;
                        const uint32_t n = left < PC_CHUNK_BYTES ? left : PC_CHUNK_BYTES;
                        mbar_expect_tx(sh->full + s, n);
                        tma_bulk_g2s(ring + (size_t)s * PC_CHUNK_BYTES, src + (size_t)c * PC_CHUNK_BYTES, n, sh->full + s);
                    }
                }
            }
            __syncwarp();
            cluster.sync();
        }
    } else if (warp == 3) {
        // ------------------------------------------------ MMA issuer
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_bf16(128, PC_N);
        const int nk16 = H / 16;
        bool ok = true;
        if (leader) ok = pc_mbar_wait(&sh->wbar, 0, &sh->dead, a.err, 23);
        uint32_t g = 0;
        const uint32_t ring_a = smem_u32(ring), w_a = smem_u32(wsm);
        for (int i = 0; i < T; ++i) {
            if (leader && ok) {
                for (int c = 0; c < nchunk; ++c, ++g) {
                    const int s = g % R;
                    const uint32_t ph = (g / R) & 1u;
                    if (!pc_mbar_wait(sh->full + s, ph, &sh->dead, a.err, 24)) { ok = false; break; }
                    tc_fence_after();
                    const int k0 = c * 4, k1 = k0 + 4 < nk16 ? k0 + 4 : nk16;
                    for (int q = k0; q < k1; ++q) {
                        const uint64_t ad = umma_desc(ring_a + s * PC_CHUNK_BYTES + (q - k0) * 2048, 1024, 128);
                        const uint64_t bd = umma_desc(w_a + q * 1024, 512, 128);
                        umma_bf16(tmem_base, ad, bd, idesc, q > 0 ? 1u : 0u);
                    }
                    umma_commit(sh->empty + s);
                }
                if (ok) umma_commit(&sh->tmem_full);
            }
            __syncwarp();
            cluster.sync();
        }
    } else {
        // ------------------------------------------------ epilogue
        const int b = warp * 32 + lane;
        const bool valid = b < a.B;
        const int u0 = 8 * j;                              // == 32 * (j / 4) + 8 * s_rank
        const float *p0 = cluster.map_shared_rank(part, 0), *p1 = cluster.map_shared_rank(part, 1);
        const float *p2 = cluster.map_shared_rank(part, 2), *p3 = cluster.map_shared_rank(part, 3);
        float dc[8], c_new[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dc[i] = 0.f;
            c_new[i] = valid ? a.c_stash[((size_t)T * a.B + b) * H + u0 + i] : 0.f;
        }
        bool ok = true;
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            float4 ga[8];
            float c_prev[8], dhe[8];
            if (valid) {
                const float4 *gp = reinterpret_cast<const float4 *>(a.gates_stash + ((size_t)t * a.B + b) * 4 * H + 4 * u0);
#pragma unroll
                for (int k = 0; k < 8; ++k) ga[k] = __ldcs(gp + k);
                const float4 *cp = reinterpret_cast<const float4 *>(a.c_stash + ((size_t)t * a.B + b) * H + u0);
                const float4 c0 = __ldcs(cp), c1 = __ldcs(cp + 1);
                c_prev[0] = c0.x; c_prev[1] = c0.y; c_prev[2] = c0.z; c_prev[3] = c0.w;
                c_prev[4] = c1.x; c_prev[5] = c1.y; c_prev[6] = c1.z; c_prev[7] = c1.w;
                const float *dp = a.dh_ext + (size_t)t * a.dh_tstride + (size_t)b * a.dh_ld + u0;
#pragma unroll
                for (int k = 0; k < 8; ++k) dhe[k] = __ldcs(dp + k);
            }
            if (ok) ok = __all_sync(0xffffffffu, pc_mbar_wait(&sh->tmem_full, (uint32_t)i & 1u, &sh->dead, a.err, 25)) != 0;
            if (ok) {
                tc_fence_after();
                float acc[32];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + 16u, acc + 16);
                tc_fence_before();
#pragma unroll
                for (int n = 0; n < 32; ++n) part[n * PC_ROWS + b] = acc[n];
            }
            __syncwarp();
            cluster.sync();                                // the four K-quarter partials of this cluster are in shared memory
            if (ok && valid) {
                uint32_t dgp[16];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int o = (8 * s_rank + k) * PC_ROWS + b;
                    const float dh = dhe[k] + (((p0[o] + p1[o]) + p2[o]) + p3[o]);
                    const float mult = drop_mult(a.drop, a.site, (uint32_t)t, (uint32_t)(b + a.row_offset), (uint32_t)(u0 + k));
                    float dcp;
                    const float4 d4 = lstm_bwd_point(dh, mult, ga[k], c_prev[k], c_new[k], dc[k], dcp);
                    dc[k] = dcp;
                    c_new[k] = c_prev[k];
                    dgp[2 * k] = pack_bf2(d4.x, d4.y);
                    dgp[2 * k + 1] = pack_bf2(d4.z, d4.w);
                }
                // gate rows 4*u0 .. 4*u0+31 = k-chunks 4j .. 4j+3 of the image, 64 contiguous bytes of the row-major rows
                __nv_bfloat16 *img = a.gimg + (size_t)((i + 1) & 1) * 4 * H * 64;
                uint4 *rm = reinterpret_cast<uint4 *>(a.dg_rm + ((size_t)t * a.B + b) * 4 * H + 4 * u0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = make_uint4(dgp[4 * c], dgp[4 * c + 1], dgp[4 * c + 2], dgp[4 * c + 3]);
                    *reinterpret_cast<uint4 *>(img + ((size_t)(4 * j + c) * PC_ROWS + b) * 8) = v;
                    rm[c] = v;
                }
            }
            __threadfence();
            fence_proxy_async_all();
            asm volatile("bar.sync 1, 64;" ::: "memory");
            if (threadIdx.x == 0) gbar_arrive(a.bar);
        }
    }
    __syncthreads();
    cluster.sync();                                        // peers may still be reading this CTA's partial
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
    }
}

// ================================================================================================ host side
// recurrent-weight images.  w_hh: torch layout [4H, H] (rows g*H + u), ld = row stride.
// mode 0 (forward):  img[j][kc][4*lu+g][e] = w_hh[g*H + 8j+lu][kc*8+e]
// mode 1 (backward): img[j=4r+s][kc][n][e] = w_hh[g*H + u][32r+n],  4u+g = s*H + kc*8+e
__global__ void k_pc_pack_w(const float *__restrict__ w_hh, int ld, int H, int mode, __nv_bfloat16 *__restrict__ img) {
    const size_t total = (size_t)(H / 8) * (H / 8) * 32 * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(i & 7), r = (int)((i >> 3) & 31);
        const size_t rest = i >> 8;
        const int kc = (int)(rest % (H / 8)), j = (int)(rest / (H / 8));
        float v;
        if (mode == 0) {
            const int lu = r >> 2, g = r & 3;
            v = w_hh[(size_t)(g * H + 8 * j + lu) * ld + kc * 8 + e];
        } else {
            const int rr = j >> 2, s = j & 3;
            const int kg = s * H + kc * 8 + e, u = kg >> 2, g = kg & 3;
            v = w_hh[(size_t)(g * H + u) * ld + 32 * rr + r];
        }
        img[i] = __float2bfloat16(v);
    }
}

inline size_t pc_wimg_elems(int H) { return (size_t)(H / 8) * (H / 8) * 32 * 8; }   // == 4 * H * H
inline size_t pc_smem_bytes(int H, bool bwd) {
    const int img_bytes = H * 128;
    const int nchunk = (img_bytes + PC_CHUNK_BYTES - 1) / PC_CHUNK_BYTES;
    const int R = nchunk < PC_MAXRING ? nchunk : PC_MAXRING;
    return (size_t)R * PC_CHUNK_BYTES + (size_t)H * 64 + (bwd ? PC_N * PC_ROWS * 4 : 0) + sizeof(PcShared) + 1024 + 1024;
}

// can the persistent chain run for this shape on the current device?
inline bool pc_supported(int H, int B) {
    static int sms = -1;
    if (sms < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (H % 32 != 0 || H < 32 || B < 1 || B > PC_ROWS) return false;
    if (H / 8 > sms) return false;
    if (pc_smem_bytes(H, true) > 227 * 1024) return false;
    return true;
}
inline bool pc_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("GVX_PERSISTENT");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

inline int launch_lstm_chain_fwd(const PcFwdArgs &a, cudaStream_t st) {
    const size_t smem = pc_smem_bytes(a.H, false);
    static size_t configured = 0;
    if (configured < smem) {
        GVX_CUDA(cudaFuncSetAttribute(k_lstm_chain_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    GVX_CUDA(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st));
    k_lstm_chain_fwd<<<a.H / 8, PC_THREADS, smem, st>>>(a);
    GVX_LAUNCHED(1);
    GVX_CUDA(cudaGetLastError());
    return 0;
}
inline int launch_lstm_chain_bwd(const PcBwdArgs &a, cudaStream_t st) {
    const size_t smem = pc_smem_bytes(a.H, true);
    static size_t configured = 0;
    if (configured < smem) {
        GVX_CUDA(cudaFuncSetAttribute(k_lstm_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    GVX_CUDA(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st));
    k_lstm_chain_bwd<<<a.H / 8, PC_THREADS, smem, st>>>(a);
    GVX_LAUNCHED(1);
    GVX_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gvx
