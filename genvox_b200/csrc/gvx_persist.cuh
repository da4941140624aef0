// genvox_b200 — persistent recurrence kernels (bf16 mode): a whole LSTM chain in ONE launch.
//
// nn.LSTMCell's gate pre-activations (/root/reference/models/tts/tacotron2.py:340,:357) split into an input part and a
// recurrent part.  In teacher-forced training the input part of the decoder LSTM ([h_att_t | ctx_t] . W_ih^T + biases)
// does not depend on the decoder LSTM itself, so it is computed for all T frames by one time-batched GEMM and what
// remains on the sequential chain is   gates_t = pre_t + W_hh . h_{t-1}   followed by the cell.  These kernels run that
// chain for all T steps in a single launch:
//
//   * grid = H/8 CTAs (128 for H = 1024), one per SM, all co-resident;
//   * a CTA's slice of the recurrent weights is loaded ONCE into shared memory and stays there for the whole sequence
//     (forward: 64 gate rows x H = 128 KB; backward: 32 output units x a K quarter of 4H = 64 KB); the cell state (forward) /
//     its gradient (backward) of the CTA's (row, unit) pairs lives in registers across steps;
//   * per step the operand image written by all CTAs (h_{t-1}, resp. d gates_{t+1}; bf16, SWIZZLE_128B slabs of 64 K
//     columns) is copied into shared memory by ALL threads of the CTA with 16-byte cp.async and contracted on tcgen05 by one
//     unrolled block of UMMA 64 x 32 x 16 (accumulator in TMEM);
//   * steps are separated by a grid-wide barrier (release/acquire counter in global memory).
//
// Forward: CTA (jc, rh) owns 16 hidden units for the 32 batch rows of half rh (see k_lstm_chain_fwd_swap).
// Backward (BPTT of the same chain): d h_{t-1} = W_hh^T . d gates_t has K = 4H; the K range is split over the 4 CTAs of
// a thread-block cluster, the four partial [64 x 32] tiles are pushed to the owner rank through distributed shared memory
// (st.async + mbarrier complete_tx) and summed in rank order (deterministic), and each CTA finishes the cell backward of its
// own 8 units.  Every wait is bounded; on timeout an error code is recorded, the CTA's roles stop working and the host raises.
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "gvx_bf16.cuh"
#include "gvx_gemm.cuh"
#include "gvx_tc.cuh"

namespace gvx {

constexpr int PC_ROWS = 64;                               // batch rows of an activation image (B <= 64)
constexpr int PC_N = 32;                                  // UMMA N: 8 units x 4 gates (fwd) / 32 output units (bwd)
constexpr int PC_CHUNK_BYTES = 64 * PC_ROWS * 2;          // one TMA chunk = one 64-element K slab of [64 rows][128 B], SWIZZLE_128B
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

__device__ __forceinline__ void pc_stamp(long long *dbg, int cta, int t, int k) {
    if (dbg && cta == 0 && t < 1024) dbg[t * 32 + k] = clock64();
}

// row-major bf16 destination of the hidden state of frame t + toff (skipped when t + toff >= T)
struct PcOut {
    __nv_bfloat16 *p;
    int ld, koff, toff;
    long long tstride;
};

// ================================================================================================ forward
struct PcFwdArgs {
    const __nv_bfloat16 *Wimg;   // [H/16 blocks][K/64 slabs][64 rows][64] SWIZZLE_128B; row = local gate row 4*lu+g of the block's 16 units (W_hh, k_pc_pack_w mode 0)
    const float *pre;            // [T][B][4H]  unit-major columns 4u+g: input contribution (may alias gates_stash)
    const float *bias;           // [4H] unit-major b_ih + b_hh, or null
    __nv_bfloat16 *himg;         // [2][2 batch halves][K/64 slabs][32 rows][64] SWIZZLE_128B ping-pong operand image of h (private to the kernel); image 0 = h_{-1} (zeros)
    float *c_stash;              // [T+1][B][H]  row 0 = c_{-1} (caller), row t+1 written at step t
    float *gates_stash;          // [T][B][4H] gate activations (i,f,g,o per unit) or null
    PcOut out[2];
    unsigned *bar;               // two step-barrier counters (one per batch half, 128 B apart), zeroed by the launcher
    int *err;
    DropCfg drop;
    uint32_t site;
    int row_offset, B, T, H;
    long long *dbg;              // optional timeline of CTA 0 (gvx_debug_timeline): [t][8] clock64 stamps
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// SFU activations (ex2 / rcp, abs. error ~2e-7) for the cells on the sequential chains, as in the attention chains
__device__ __forceinline__ float pc_sigmoid(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return r;
}
__device__ __forceinline__ float pc_tanh(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}
__device__ __forceinline__ float4 pc_lstm_bwd_point(float dh_dropped, float mult, float4 ga, float c_prev, float c_new, float dc_in,
                                                    float &dc_prev) {
    const float dh = dh_dropped * mult;
    const float tc = pc_tanh(c_new);
    const float d_o = dh * tc;
    const float dc = dc_in + dh * ga.w * (1.f - tc * tc);
    const float d_i = dc * ga.z, d_g = dc * ga.x, d_f = dc * c_prev;
    dc_prev = dc * ga.y;
    return make_float4(d_i * ga.x * (1.f - ga.x), d_f * ga.y * (1.f - ga.y), d_g * (1.f - ga.z * ga.z), d_o * ga.w * (1.f - ga.w));
}

constexpr int PCF_THREADS = 512;           // backward chain
constexpr int PCF64_THREADS = 576;         // forward chain: 16 loader / epilogue warps + weight loader + MMA issuer

// Forward chain.  The two halves of the batch never meet in the recurrence, so CTA
// (jc, rh) = (j >> 1, j & 1) owns SIXTEEN hidden units (64 gate rows: the UMMA M side, 128 KB of W_hh resident) for the 32 batch
// rows of half rh (the N side): it needs only its half of the h image - 64 KB per step - and the step barrier only joins the 64
// CTAs of a half.  Two measured facts shape the step (profiles/chain_timeline.py):
//   * what bounded the TMA-fed version (one bulk copy + one mbarrier wait + 4 MMAs per 64-column slab, ~280 cycles per slab
//     whether the slab is 4, 8 or 16 KB, whatever the MMA shape, with 1 or 4 accumulators) was the single-thread issue loop
//     itself: ~45 dependent instructions per slab.  The 512 epilogue threads are idle at that moment, so THEY fetch the half image
//     with 16-byte cp.async (64 KB in ~1.4k cycles), and the issuer then fires all 64 MMAs from one unrolled block with
//     immediate descriptor offsets;
//   * the image is private to this kernel: [ping-pong][batch half][slab][32 rows][128 B], SWIZZLE_128B inside a slab.
//   UMMA 64 x 32 x 16: A = weight slab [64 gate rows][64 k] (resident), B = image slab [32 batch rows][64 k];
//   accumulator D[gate row m][batch row n]: m -> lane (m & 15) of TMEM quadrant m >> 4, n -> column.
//   Epilogue warp w reads quadrant w & 3, columns 8 (w >> 2) .. +7, and the [64 x 32] tile is transposed through shared
//   memory ([batch row][gate row], stride 68) so that the cell runs with thread = (batch row, unit), unit fastest.
__global__ void __launch_bounds__(PCF64_THREADS, 1) k_lstm_chain_fwd_swap(const PcFwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int H = a.H, T = a.T;
    const int nchunk = (H + 63) / 64;
    const int img_bytes = nchunk * PC_CHUNK_BYTES;         // one h image (both batch halves)
    const uint32_t wbytes = (uint32_t)nchunk * 8192;       // [slab][64 gate rows][128 B]
    constexpr uint32_t SLOT = 32 * 128;                    // one slab of a batch half: 32 rows x 128 B
    uint8_t *wsm = smem;
    uint8_t *ring = wsm + wbytes;
    PcShared *sh = (PcShared *)(ring + (size_t)nchunk * SLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, j = blockIdx.x;
    const int jc = j >> 1, rh = j & 1;
    const unsigned nhalf = gridDim.x >> 1;                 // CTAs that share a batch half = arrivals per barrier
    unsigned *bar = a.bar + 32 * rh;                       // one counter per batch half, 128 B apart

    if (threadIdx.x == 0) {
        mbar_init(sh->full + 0, 1);                        // "the first / second half of the K slabs of this step is in shared memory"
        mbar_init(sh->full + 1, 1);
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

    if (warp == 16) {
        // ------------------------------------------------ the resident weights, once
        if (elect_one()) {
            mbar_expect_tx(&sh->wbar, wbytes);
            const uint8_t *wsrc = (const uint8_t *)a.Wimg + (size_t)jc * wbytes;
            for (uint32_t off = 0; off < wbytes; off += 16384) {
                const uint32_t n = wbytes - off < 16384 ? wbytes - off : 16384;
                tma_bulk_g2s(wsm + off, wsrc + off, n, &sh->wbar);
            }
        }
        __syncwarp();
    } else if (warp == 17) {
        // ------------------------------------------------ MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(64, 32);
            bool ok = pc_mbar_wait(&sh->wbar, 0, &sh->dead, a.err, 13);
            const uint64_t a0 = umma_desc_sw128(smem_u32(wsm)), b0 = umma_desc_sw128(smem_u32(ring));
            for (int t = 0; t < T && ok; ++t) {
                // descriptors advance in 16-byte units: slab stride 8 KB (A) / 4 KB (B), 32 B per K = 16 step
                if (nchunk == 16) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (!pc_mbar_wait(sh->full + hf, (uint32_t)t & 1u, &sh->dead, a.err, 14)) { ok = false; break; }
                        if (hf == 0) pc_stamp(a.dbg, j, t, 8);
                        tc_fence_after();
#pragma unroll
                        for (int c = 8 * hf; c < 8 * hf + 8; ++c) {
                            const uint64_t ad = a0 + (uint64_t)(c * (8192 >> 4)), bd = b0 + (uint64_t)(c * (SLOT >> 4));
                            umma_bf16(tmem_base, ad, bd, idesc, c > 0 ? 1u : 0u);
                            umma_bf16(tmem_base, ad + 2, bd + 2, idesc, 1u);
                            umma_bf16(tmem_base, ad + 4, bd + 4, idesc, 1u);
                            umma_bf16(tmem_base, ad + 6, bd + 6, idesc, 1u);
                        }
                    }
                    if (!ok) break;
                } else {
                    if (!pc_mbar_wait(sh->full + 0, (uint32_t)t & 1u, &sh->dead, a.err, 14)) break;
                    if (!pc_mbar_wait(sh->full + 1, (uint32_t)t & 1u, &sh->dead, a.err, 14)) break;
                    tc_fence_after();
                    for (int c = 0; c < nchunk; ++c) {
                        const uint64_t ad = a0 + (uint64_t)(c * (8192 >> 4)), bd = b0 + (uint64_t)(c * (SLOT >> 4));
                        umma_bf16(tmem_base, ad, bd, idesc, c > 0 ? 1u : 0u);
                        umma_bf16(tmem_base, ad + 2, bd + 2, idesc, 1u);
                        umma_bf16(tmem_base, ad + 4, bd + 4, idesc, 1u);
                        umma_bf16(tmem_base, ad + 6, bd + 6, idesc, 1u);
                    }
                }
                umma_commit(&sh->tmem_full);
                pc_stamp(a.dbg, j, t, 2);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ image loaders + epilogue
        constexpr int GS = 68;                             // tile stride: [32 batch rows][64 gate rows + 4]
        float *gt = (float *)(((uintptr_t)(sh + 1) + 15) & ~(uintptr_t)15);
        const int e = threadIdx.x;                         // 0 .. 511
        const int bl = e >> 4, lu = e & 15, b = 32 * rh + bl, u = 16 * jc + lu;
        const bool valid = b < a.B;
        const int mrow = (warp & 3) * 16 + (lane & 15), cq = warp >> 2;       // TMEM side: gate row / column octet of this thread
        float c = valid ? a.c_stash[(size_t)b * H + u] : 0.f;
        const float4 bi = a.bias ? *reinterpret_cast<const float4 *>(a.bias + 4 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(8 * cq);
        // 2-byte slot of (row b, unit u) in the swizzled h image: slab u / 64, 16-byte chunk (u % 64) / 8
        const size_t himg_off = ((size_t)rh * nchunk + (jc >> 2)) * SLOT + bl * 128 + (((2 * (jc & 3) + (lu >> 3)) ^ (bl & 7)) << 4) + 2 * (lu & 7);
        const int nvec = nchunk * (int)SLOT / 16;
        bool ok = true;
        for (int t = 0; t < T; ++t) {
            float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
            float dm = 1.f;
            if (valid) {
                pr = __ldcs(reinterpret_cast<const float4 *>(a.pre + ((size_t)t * a.B + b) * 4 * H + 4 * u));
                dm = drop_mult(a.drop, a.site, (uint32_t)t, (uint32_t)(b + a.row_offset), (uint32_t)u);     // before the wait: off the chain
            }
            // ---- h_{t-1} of this batch half -> shared memory
            if (t > 0 && threadIdx.x == 0 && ok) gbar_wait(bar, nhalf * (unsigned)t, &sh->dead, a.err, 11);
            if (threadIdx.x == 0) pc_stamp(a.dbg, j, t, 0);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            {
                const uint8_t *src = (const uint8_t *)a.himg + (size_t)(t & 1) * img_bytes + (size_t)rh * nchunk * SLOT;
                // two commit groups (K slabs [0, n/2) and [n/2, n)): the issuer starts on the first half while the second is landing
                const int nv0 = (nchunk / 2) * (int)SLOT / 16;
                for (int i = e; i < nv0; i += 512) cp_async16(ring + (size_t)i * 16, src + (size_t)i * 16, true);
                cp_async_commit();
                for (int i = nv0 + e; i < nvec; i += 512) cp_async16(ring + (size_t)i * 16, src + (size_t)i * 16, true);
                cp_async_commit();
                cp_async_wait<1>();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> the MMA's async-proxy reads
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sh->full + 0)) : "memory");
            cp_async_wait<0>();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sh->full + 1)) : "memory");
                pc_stamp(a.dbg, j, t, 1);
            }
            if (ok) ok = __all_sync(0xffffffffu, pc_mbar_wait(&sh->tmem_full, (uint32_t)t & 1u, &sh->dead, a.err, 15)) != 0;
            if (threadIdx.x == 0) pc_stamp(a.dbg, j, t, 3);
            if (ok) {
                float acc[8];
                tc_fence_after();
                tmem_ld8(taddr, acc);
                tc_fence_before();
                if (lane < 16) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) gt[(8 * cq + i) * GS + mrow] = acc[i];
                }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
            float hv = 0.f;
            if (ok && valid) {
                const float4 g4 = *reinterpret_cast<const float4 *>(gt + bl * GS + 4 * lu);
                const float gi = pc_sigmoid(g4.x + pr.x + bi.x), gf = pc_sigmoid(g4.y + pr.y + bi.y);
                const float gg = pc_tanh(g4.z + pr.z + bi.z), go = pc_sigmoid(g4.w + pr.w + bi.w);
                c = gf * c + gi * gg;
                hv = go * pc_tanh(c) * dm;
                ga = make_float4(gi, gf, gg, go);
            }
            // two units per 4-byte store: the even lane of a unit pair writes both halves
            const float hv_hi = __shfl_down_sync(0xffffffffu, hv, 1);
            const uint32_t hp = pack_bf2(hv, hv_hi);
            const bool writer = ok && valid && (lu & 1) == 0;
            if (writer) *reinterpret_cast<uint32_t *>((uint8_t *)a.himg + (size_t)((t + 1) & 1) * img_bytes + himg_off) = hp;
            if (threadIdx.x == 0) pc_stamp(a.dbg, j, t, 4);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            // release at gpu scope is cumulative over the stores ordered before it by the CTA barrier
            if (threadIdx.x == 0) { gbar_arrive(bar); pc_stamp(a.dbg, j, t, 6); }
            if (ok && valid) {
                if (a.gates_stash) *reinterpret_cast<float4 *>(a.gates_stash + ((size_t)t * a.B + b) * 4 * H + 4 * u) = ga;
                a.c_stash[((size_t)(t + 1) * a.B + b) * H + u] = c;
                if (writer) {
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const PcOut &d = a.out[o];
                        if (d.p && t + d.toff < T)
                            *reinterpret_cast<uint32_t *>(d.p + (size_t)(t + d.toff) * d.tstride + (size_t)b * d.ld + d.koff + u) = hp;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
    }
}

// ================================================================================================ backward
struct PcBwdArgs {
    const __nv_bfloat16 *Wimg;   // [H/64 unit blocks][4 K quarters][H/64 slabs][64 rows][64] SWIZZLE_128B: row n = output unit 64 ro + n, k = gate row s*H + kk (unit-major)
    __nv_bfloat16 *gimg;         // [2][4 K-quarters][2 batch halves][H/64 slabs][32 rows][64] SWIZZLE_128B ping-pong d-gates image (k = 4u+g), private to the kernel
    const float *dh_ext;         // d h (dropped) of frame t from everything but the recurrence: dh_ext[t*dh_tstride + b*dh_ld + u]
    int dh_ld;
    long long dh_tstride;
    const float *gates_stash;    // [T][B][4H]
    const float *c_stash;        // [T+1][B][H]
    __nv_bfloat16 *dg_rm;        // [T][B][4H] row-major d gates (columns 4u+g) for the time-batched GEMMs
    unsigned *bar;               // two step-barrier counters (one per batch half, 128 B apart), zeroed by the launcher
    int *err;
    DropCfg drop;
    uint32_t site;
    int row_offset, B, T, H;
    long long *dbg;
};

// 512 threads, one loop.  Cluster c = j >> 2 = (output-unit block ro = c >> 1 of 64 units, batch half rh = c & 1); rank s = j & 3 owns
// the K quarter s of the 4H gate rows AND, for the cell backward, units 64 ro + 16 s .. + 15 of the 32 rows of its half:
// thread = (local row ub, unit uk), unit fastest, keeps d c of its (row, unit) in a register.  The two batch halves never meet
// in the recurrence: each half has its own step barrier (64 CTAs) and its own half of the d-gates image.  Per step
//   thread 0: step barrier; ALL threads: the CTA's [32 rows x K quarter] piece of the d-gates image -> shared memory with 16-byte
//             cp.async (64 KB, two commit groups); warp 1 (one elected lane): 64 x UMMA 64 x 32 x 16 from one unrolled block
//             (A = W_hh^T slab [64 output units][64 k], resident; B = image slab [32 rows][64 k]);
//   all warps: tcgen05.ld of the [64 units x 32 rows] partial tile (unit m in lane m & 15 of TMEM quadrant m >> 4 = the owner
//              rank), PUSHED into the shared memory of the owner (st.async + mbarrier complete_tx: no cluster barrier, no fence -
//              a cluster.sync per step costs a MEMBAR.ALL.GPU in every thread); each rank sums the four partials of its 16 units
//              in rank order.
constexpr int PCB_PLD = 16;        // pushed partials: [src rank][32 rows][16 units] - a push instruction (16 lanes = 16 units of one row) is one
                                   // contiguous 64-byte DSMEM write, and the reads by (row, unit) are conflict-free

struct PcbShared {
    uint64_t full[2], tmem_full, wbar, xb;
    uint32_t tmem_slot;
    volatile int dead;
};

__device__ __forceinline__ void pcb_st_async(uint32_t caddr, float v, uint32_t cmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(caddr), "f"(v), "r"(cmbar) : "memory");
}
__device__ __forceinline__ uint32_t pcb_try_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t pcb_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(PCF_THREADS, 1) k_lstm_chain_bwd(const PcBwdArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int H = a.H, T = a.T;
    const int nchunk = (H + 63) / 64;                      // K slabs of one quarter
    constexpr uint32_t SLOT = 32 * 128;                    // one slab of a batch half: 32 rows x 128 B
    const int qh_bytes = nchunk * (int)SLOT;               // [slab][32 rows][128 B]: one (K quarter, batch half) piece of the image
    const int img_bytes = 8 * qh_bytes;                    // one d-gates image: [4 K quarters][2 batch halves] pieces
    const uint32_t wbytes = (uint32_t)nchunk * 8192;       // [slab][64 output units][128 B]
    uint8_t *wsm = smem;
    uint8_t *ring = wsm + wbytes;
    float *pin = (float *)(ring + (size_t)nchunk * SLOT);  // [4 src ranks][32 rows][PCB_PLD] partial d h of the own units
    PcbShared *sh = (PcbShared *)(pin + 4 * 32 * PCB_PLD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, j = blockIdx.x;
    const int s_rank = (int)cluster.block_rank();          // == j & 3
    const int cl = j >> 2, ro = cl >> 1, rh = cl & 1;
    const unsigned nhalf = gridDim.x >> 1;                 // CTAs of one batch half = arrivals per step barrier
    unsigned *bar = a.bar + 32 * rh;
    const int ub = tid >> 4, uk = tid & 15, b = 32 * rh + ub, u = 64 * ro + 16 * s_rank + uk;
    const bool valid = b < a.B;

    if (tid == 0) {
        mbar_init(sh->full + 0, 1);
        mbar_init(sh->full + 1, 1);
        mbar_init(&sh->tmem_full, 1);
        mbar_init(&sh->wbar, 1);
        mbar_init(&sh->xb, 1);
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
    if (tid == 0) {
        mbar_expect_tx(&sh->wbar, wbytes);
        const uint8_t *wsrc = (const uint8_t *)a.Wimg + (size_t)(ro * 4 + s_rank) * wbytes;
        for (uint32_t off = 0; off < wbytes; off += 16384) {
            const uint32_t n = wbytes - off < 16384 ? wbytes - off : 16384;
            tma_bulk_g2s(wsm + off, wsrc + off, n, &sh->wbar);
        }
    }
    cluster.sync();                                        // every CTA's mbarriers exist before any remote push
    const bool okw = pc_mbar_wait(&sh->wbar, 0, &sh->dead, a.err, 23);

    // gate rows 4u .. 4u+3 of the 4H range = 8 bytes of the swizzled image piece (K quarter 4u / H, batch half rh)
    const int kg = 4 * u, quarter = kg / H, kk = kg - quarter * H;
    const size_t gimg_off = ((size_t)(quarter * 2 + rh) * nchunk + (kk >> 6)) * SLOT + ub * 128 + ((((kk & 63) >> 3) ^ (ub & 7)) << 4) + (kk & 7) * 2;
    const uint32_t my_pin = smem_u32(pin), my_xb = smem_u32(&sh->xb);
    const int tq = warp & 3, tcq = warp >> 2;              // TMEM side: lane quadrant (units 16 tq .. = owner rank tq) and row octet of this warp
    const uint32_t taddr = tmem_base + ((uint32_t)(tq * 32) << 16) + (uint32_t)(8 * tcq);
    const uint32_t dst_pin = pcb_mapa(my_pin, (uint32_t)tq) + 4u * (uint32_t)((s_rank * 32 + 8 * tcq) * PCB_PLD + (lane & 15));
    const uint32_t dst_xb = pcb_mapa(my_xb, (uint32_t)tq);
    const int nvec = qh_bytes / 16, nv0 = (nchunk / 2) * (int)SLOT / 16;
    float dc = 0.f;
    float c_new = valid ? a.c_stash[((size_t)T * a.B + b) * H + u] : 0.f;
    for (int i = 0; i < T; ++i) {
        const int t = T - 1 - i;
        float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
        float c_prev = 0.f, dhe = 0.f;
        if (valid) {
            ga = __ldcs(reinterpret_cast<const float4 *>(a.gates_stash + ((size_t)t * a.B + b) * 4 * H + 4 * u));
            c_prev = __ldcs(a.c_stash + ((size_t)t * a.B + b) * H + u);
            dhe = __ldcs(a.dh_ext + (size_t)t * a.dh_tstride + (size_t)b * a.dh_ld + u);
        }
        const float mult = valid ? drop_mult(a.drop, a.site, (uint32_t)t, (uint32_t)(b + a.row_offset), (uint32_t)u) : 1.f;
        float rec = 0.f;
        if (i > 0) {       // d h_t from the recurrence: W_hh^T . d gates_{t+1} (at t = T-1 there is none)
            const uint32_t par = (uint32_t)(i - 1) & 1u;
            if (tid == 0) mbar_expect_tx(&sh->xb, 4u * 16u * 32u * 4u);
            if (tid == 0 && okw) {
                gbar_wait(bar, nhalf * (unsigned)i, &sh->dead, a.err, 21);
                pc_stamp(a.dbg, j, i, 0);
            }
            __syncthreads();
            {
                const uint8_t *src = (const uint8_t *)a.gimg + (size_t)(i & 1) * img_bytes + (size_t)(s_rank * 2 + rh) * qh_bytes;
                for (int v = tid; v < nv0; v += PCF_THREADS) cp_async16(ring + (size_t)v * 16, src + (size_t)v * 16, true);
                cp_async_commit();
                for (int v = nv0 + tid; v < nvec; v += PCF_THREADS) cp_async16(ring + (size_t)v * 16, src + (size_t)v * 16, true);
                cp_async_commit();
                cp_async_wait<1>();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sh->full + 0)) : "memory");
            cp_async_wait<0>();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sh->full + 1)) : "memory");
            if (warp == 1) {
                if (elect_one()) {
                    constexpr uint32_t idesc = umma_idesc_bf16(64, 32);
                    const uint64_t a0 = umma_desc_sw128(smem_u32(wsm)), b0 = umma_desc_sw128(smem_u32(ring));
                    if (okw && !sh->dead) {
                        bool ok = true;
                        if (nchunk == 16) {
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                if (ok) ok = pc_mbar_wait(sh->full + hf, par, &sh->dead, a.err, 24);
                                tc_fence_after();
#pragma unroll
                                for (int c = 8 * hf; c < 8 * hf + 8; ++c) {
                                    const uint64_t ad = a0 + (uint64_t)(c * (8192 >> 4)), bd = b0 + (uint64_t)(c * (SLOT >> 4));
                                    umma_bf16(tmem_base, ad, bd, idesc, c > 0 ? 1u : 0u);
                                    umma_bf16(tmem_base, ad + 2, bd + 2, idesc, 1u);
                                    umma_bf16(tmem_base, ad + 4, bd + 4, idesc, 1u);
                                    umma_bf16(tmem_base, ad + 6, bd + 6, idesc, 1u);
                                }
                            }
                        } else {
                            ok = pc_mbar_wait(sh->full + 0, par, &sh->dead, a.err, 24) && pc_mbar_wait(sh->full + 1, par, &sh->dead, a.err, 24);
                            tc_fence_after();
                            for (int c = 0; c < nchunk; ++c) {
                                const uint64_t ad = a0 + (uint64_t)(c * (8192 >> 4)), bd = b0 + (uint64_t)(c * (SLOT >> 4));
                                umma_bf16(tmem_base, ad, bd, idesc, c > 0 ? 1u : 0u);
                                umma_bf16(tmem_base, ad + 2, bd + 2, idesc, 1u);
                                umma_bf16(tmem_base, ad + 4, bd + 4, idesc, 1u);
                                umma_bf16(tmem_base, ad + 6, bd + 6, idesc, 1u);
                            }
                        }
                        umma_commit(&sh->tmem_full);
                        pc_stamp(a.dbg, j, i, 2);
                        pc_mbar_wait(&sh->tmem_full, par, &sh->dead, a.err, 25);      // the only thread that waits for the accumulator
                    }
                }
                __syncwarp();
            }
            __syncthreads();
            if (tid == 0) pc_stamp(a.dbg, j, i, 3);
            if (!sh->dead) {
                float acc[8];
                tc_fence_after();
                tmem_ld8(taddr, acc);
                tc_fence_before();
                if (lane < 16) {
#pragma unroll
                    for (int n = 0; n < 8; ++n) pcb_st_async(dst_pin + 4u * (uint32_t)(n * PCB_PLD), acc[n], dst_xb);
                }
            }
            {   // lane 0 spins; then EVERY lane acquires the completed phase itself (pushed data is only guaranteed visible to
                // threads that have observed the barrier)
                int okx = 1;
                if (lane == 0) okx = pc_mbar_wait(&sh->xb, par, &sh->dead, a.err, 26) ? 1 : 0;
                okx = __shfl_sync(0xffffffffu, okx, 0);
                if (okx) {
                    while (!pcb_try_wait_cluster(&sh->xb, par)) {}
                }
            }
            const int o = ub * PCB_PLD + uk;
            rec = ((pin[o] + pin[32 * PCB_PLD + o]) + pin[64 * PCB_PLD + o]) + pin[96 * PCB_PLD + o];
            if (tid == 0) pc_stamp(a.dbg, j, i, 4);
        }
        uint2 dgp = make_uint2(0u, 0u);
        if (valid && !sh->dead) {
            float dcp;
            const float4 d4 = pc_lstm_bwd_point(dhe + rec, mult, ga, c_prev, c_new, dc, dcp);
            dc = dcp;
            c_new = c_prev;
            dgp = make_uint2(pack_bf2(d4.x, d4.y), pack_bf2(d4.z, d4.w));
            *reinterpret_cast<uint2 *>((uint8_t *)a.gimg + (size_t)((i + 1) & 1) * img_bytes + gimg_off) = dgp;
        }
        if (tid == 0) pc_stamp(a.dbg, j, i, 5);
        __syncthreads();
        if (tid == 0) { gbar_arrive(bar); pc_stamp(a.dbg, j, i, 6); }
        if (valid) *reinterpret_cast<uint2 *>(a.dg_rm + ((size_t)t * a.B + b) * 4 * H + 4 * u) = dgp;
    }
    __syncthreads();
    cluster.sync();                                        // peers may still be pushing into this CTA's shared memory
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
    }
}

// ================================================================================================ host side
// recurrent-weight images.  w_hh: torch layout [4H, H] (rows g*H + u), ld = row stride.  Per CTA j the image is
// [Kp/64 slabs][32 rows][128 B] in the SWIZZLE_128B K-major layout (16-byte chunk c of row r at position c ^ (r & 7)),
// K padded to a multiple of 64 with zeros.
// mode 0 (forward):  block jc = 16 units, [slab][64 rows m = 4*lu+g][128 B], k  ->  w_hh[g*H + 16jc+lu][k]
// mode 1 (backward): block (ro, s) = 64 output units x K quarter, [slab][64 rows n][128 B], k  ->  w_hh[g*H + u][64 ro + n] with 4u+g = s*H + k
__global__ void k_pc_pack_w(const float *__restrict__ w_hh, int ld, int H, int mode, __nv_bfloat16 *__restrict__ img) {
    const int nslab = (H + 63) / 64;
    const size_t per_blk = (size_t)nslab * 4096;              // one block: [slab][64 rows][64]
    const size_t total = (size_t)(H / 16) * per_blk;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int blk = (int)(i / per_blk), rm = (int)(i - (size_t)blk * per_blk);
        const int slab = rm >> 12, row = (rm >> 6) & 63, cpos = (rm >> 3) & 7, e = rm & 7;
        const int k = slab * 64 + ((cpos ^ (row & 7)) << 3) + e;
        float v = 0.f;
        if (k < H) {
            if (mode == 0) {          // forward: block = 16 units, row m = 4*lu+g
                v = w_hh[(size_t)((row & 3) * H + 16 * blk + (row >> 2)) * ld + k];
            } else {                  // backward: block = (ro, sq): row n = output unit 64 ro + n, k = index inside K quarter sq
                const int ro = blk >> 2, sq = blk & 3;
                const int gk = sq * H + k, u = gk >> 2, g = gk & 3;
                v = w_hh[(size_t)(g * H + u) * ld + 64 * ro + row];
            }
        }
        img[i] = __float2bfloat16(v);
    }
}

inline size_t pc_wimg_elems(int H) { return (size_t)(H / 8) * ((H + 63) / 64) * 2048; }   // == 4 * H * H when H % 64 == 0
inline size_t pc_himg_elems(int H) { return (size_t)2 * ((H + 63) / 64) * 64 * PC_ROWS; }     // ping-pong h image (bf16 elements)
inline size_t pc_gimg_elems(int H) { return 4 * pc_himg_elems(H); }                            // ping-pong d-gates image
inline size_t pc_smem_bytes(int H, bool bwd) {     // backward chain: resident 64-row W^T slabs + half-image piece + partial tiles
    (void)bwd;
    const int nchunk = (H + 63) / 64;
    return (size_t)nchunk * 8192 + (size_t)nchunk * 4096 + 4 * 32 * PCB_PLD * 4 + sizeof(PcbShared) + 256 + 1024;
}
inline size_t pc_smem_bytes_swap(int H) {       // k_lstm_chain_fwd_swap: resident 64-row weight slabs + half-image ring + the [32][68] tile
    const int nchunk = (H + 63) / 64;
    return (size_t)nchunk * 8192 + (size_t)nchunk * 4096 + 32 * 68 * 4 + sizeof(PcShared) + 256 + 1024;
}

// how many clusters of `cluster` CTAs (block size / dynamic shared memory given) the device keeps resident at once
template <class Kern>
inline int max_resident_clusters(Kern kern, int threads, size_t smem, int cluster, int grid) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// can the persistent chain run for this shape on the current device?  (every CTA of the grid must be resident at once: the
// kernels synchronise through grid-wide barriers)
inline bool pc_coresident(int H);
inline bool pc_supported(int H, int B) {
    static int sms = -1;
    if (sms < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (H % 64 != 0 || H < 64 || B < 1 || B > PC_ROWS) return false;      // 64-unit blocks (backward), 16-unit blocks (forward)
    if (H / 8 > sms || (H + 63) / 64 > PC_MAXRING) return false;
    if (pc_smem_bytes(H, true) > 227 * 1024) return false;
    return pc_coresident(H);
}
inline int &pc_mode() {        // -1 = not yet read from the environment, 0 = off, 1 = on
    static int on = -1;
    return on;
}
inline bool pc_enabled() {
    int &on = pc_mode();
    if (on < 0) {
        const char *e = getenv("GVX_PERSISTENT");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

inline long long *&pc_dbg_buffer() {
    static long long *p = nullptr;
    return p;
}

inline bool pc_coresident(int H) {
    static int cached_H = -1;
    static bool cached = false;
    if (cached_H != H) {
        const int grid = H / 8;
        const int fwd = max_resident_clusters(k_lstm_chain_fwd_swap, PCF64_THREADS, pc_smem_bytes_swap(H), 1, grid);
        const int bwd = max_resident_clusters(k_lstm_chain_bwd, PCF_THREADS, pc_smem_bytes(H, true), 4, grid);
        cached = fwd >= grid && 4 * bwd >= grid;
        cached_H = H;
    }
    return cached;
}

inline int launch_lstm_chain_fwd(const PcFwdArgs &a_in, cudaStream_t st) {
    PcFwdArgs a = a_in;
    a.dbg = pc_dbg_buffer();
    const size_t smem = pc_smem_bytes_swap(a.H);
    static size_t configured = 0;
    if (configured < smem) {
        GVX_CUDA(cudaFuncSetAttribute(k_lstm_chain_fwd_swap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    GVX_CUDA(cudaMemsetAsync(a.bar, 0, 64 * sizeof(unsigned), st));          // two counters (one per batch half), 128 B apart
    k_lstm_chain_fwd_swap<<<a.H / 8, PCF64_THREADS, smem, st>>>(a);
    GVX_LAUNCHED(1);
    GVX_CUDA(cudaGetLastError());
    return 0;
}
inline int launch_lstm_chain_bwd(const PcBwdArgs &a_in, cudaStream_t st) {
    PcBwdArgs a = a_in;
    a.dbg = pc_dbg_buffer() ? pc_dbg_buffer() + 32 * 1024 : nullptr;
    const size_t smem = pc_smem_bytes(a.H, true);
    static size_t configured = 0;
    if (configured < smem) {
        GVX_CUDA(cudaFuncSetAttribute(k_lstm_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    GVX_CUDA(cudaMemsetAsync(a.bar, 0, 64 * sizeof(unsigned), st));          // two counters (one per batch half), 128 B apart
    k_lstm_chain_bwd<<<a.H / 8, PCF_THREADS, smem, st>>>(a);
    GVX_LAUNCHED(1);
    GVX_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gvx
