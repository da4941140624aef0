// genvox_b200 — tcgen05 / TMEM / TMA path for the recurrent gate GEMMs (bf16 operands, fp32 accumulate).
//
//   P[ks][b][m] = sum_{k in split ks} W[m, k] * X[b, k]          m < Mtiles*128,  b < NPAD
//
// Used for the batch x (in + h) x 4h LSTM gate contractions of nn.LSTMCell
// (/root/reference/models/tts/tacotron2.py:340,:357), their BPTT transposes, the query projection (:98)
// and the mel/gate projections (:361-362) in bf16 mode.
//
// Operand images.  Both operands live in HBM already in the shared-memory image tcgen05 wants
// (K-major, SWIZZLE_128B: rows of 128 B = 64 bf16 along K, 8-row groups of 1 KB, 16-byte chunk c of row r at chunk
// position c ^ (r & 7)), so one 64-element k-block of an operand is ONE contiguous chunk and is fetched by one TMA
// bulk copy (cp.async.bulk ... mbarrier::complete_tx) issued by a single elected thread:
//   W image  [Mtiles][Kpad/64][128 rows][64] bf16    (row tile, k-block, row in tile, swizzled k in block)
//   X image  [Kpad/64][NPAD rows][64]        bf16    (k-block, batch row, swizzled k in block)
// UMMA descriptors: SWIZZLE_128B, stride byte offset 1 KB; a K = 16 step inside the block advances the start address by
// 32 B.  (The first version used the no-swizzle interleaved layout; in the persistent chain kernels the MMA issuer then
// advanced one 4-MMA slab every ~730 cycles against ~495 with SWIZZLE_128B in the same loop.)
//
// Mapping: the weights are the M side (UMMA_M = 128 rows per CTA, full TMEM lane use), the batch is
// the N side (UMMA_N = NPAD <= 256); grid = (Mtiles, KS): the K range is split over KS CTAs so that
// ~all 148 SMs stream a slice of the weights; fp32 partials go to HBM/L2 and the pointwise consumer
// kernels add them in a fixed order (deterministic).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each),
// warps 2..5 = epilogue (TMEM -> registers -> global), warp 2 also owns the TMEM allocation.
// Every mbarrier wait is bounded: on timeout an error code is written to `err` and the kernel runs to
// completion instead of hanging the device.
#pragma once
#include <cuda_bf16.h>

#include <cooperative_groups.h>

#include "gvx_common.cuh"
#include "gvx_io.cuh"

namespace gvx {

constexpr int TC_KB = 64;          // K elements per pipeline stage
constexpr int TC_M = 128;          // weight rows per CTA (UMMA M)
constexpr int TC_STAGES = 6;
constexpr int TC_THREADS = 192;
constexpr long long TC_WAIT_CYCLES = 1500000000ll;   // ~0.75 s of SM clock: a stuck pipeline reports instead of hanging

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// bounded wait; returns false (and records `code`) on timeout
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *err, int code) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (clock64() - t0 < TC_WAIT_CYCLES)
        if (mbar_try_wait(bar, parity)) return true;
    if (err) atomicExch(err, code);
    return false;
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version 1 (Blackwell)
    return d;                               // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// K-major SWIZZLE_128B descriptor: rows of 128 B (64 bf16 along K), 8-row groups of 1 KB (stride byte offset), the
// 16-byte chunk c of row r stored at chunk position c ^ (r & 7); the tile base is 1 KB aligned and a K = 16 step
// inside the 64-element slab advances the start address by 32 B.  The tensor core reads this layout at full
// shared-memory bandwidth (measured in the persistent chain kernels: one 4-MMA slab per ~495 cycles against ~730 with
// the no-swizzle layout above, same issue loop).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major (16 B)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                 // layout_type SWIZZLE_128B
    return d;
}
// MN-major SWIZZLE_128B operand: 128-byte rows = 64 elements along M/N, 8-row K groups 1 KB apart (stride byte offset),
// successive 64-element M/N columns `atom_bytes` apart (leading byte offset)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t saddr, uint32_t atom_bytes = 8192) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((atom_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcGemmArgs {
    const __nv_bfloat16 *Wimg;   // [Mtiles][Kpad/64][128][64]  SWIZZLE_128B
    const __nv_bfloat16 *Ximg;   // [Kpad/64][NPAD][64]         SWIZZLE_128B
    float *P;                    // [KS][B][ldp]   partial sums, m = tile*128 + row
    int Kpad;                    // multiple of 64
    int B;                       // valid batch rows (<= NPAD)
    int ldp;                     // >= Mtiles*128
    int KS;
    int *err;                    // device int, 0 = ok
};

template <int NPAD>
struct TcCfg {
    static constexpr int WB = TC_M * TC_KB * 2;
    static constexpr int XB = NPAD * TC_KB * 2;
    static constexpr int SB = WB + XB;
    static constexpr int BAR_OFF = TC_STAGES * SB;
    static constexpr size_t SMEM = (size_t)BAR_OFF + 256 + 1024;   // + barriers + alignment slack
    static constexpr int TMEM_COLS = NPAD < 32 ? 32 : NPAD;
};

template <int NPAD>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_gemm(const TcGemmArgs a) {
    using C = TcCfg<NPAD>;
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment computed as an OFFSET into the shared array: the compiler keeps the shared address space (LDS/STS,
    // not generic LD/ST) for everything derived from it
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = (uint64_t *)(smem + C::BAR_OFF);
    uint64_t *empty = full + TC_STAGES;
    uint64_t *tmem_full = empty + TC_STAGES;
    uint32_t *tmem_slot = (uint32_t *)(tmem_full + 1);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, ks = blockIdx.y;
    const int nkb_all = a.Kpad / TC_KB;
    const int kb0 = (int)((long long)ks * nkb_all / a.KS), kb1 = (int)((long long)(ks + 1) * nkb_all / a.KS);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer
        if (elect_one()) {
            const uint8_t *wsrc = (const uint8_t *)a.Wimg + ((size_t)tile * nkb_all + kb0) * C::WB;
            const uint8_t *xsrc = (const uint8_t *)a.Ximg + (size_t)kb0 * C::XB;
            // PDL prologue: the weight images do not depend on the previous kernel of the chain, so the first
            // ring of stages is filled with weights before griddepcontrol.wait; activations follow after it.
            const int npre = nkb < TC_STAGES ? nkb : TC_STAGES;
            for (int i = 0; i < npre; ++i) {
                mbar_expect_tx(full + i, (uint32_t)C::SB);
                tma_bulk_g2s(smem + (size_t)i * C::SB, wsrc + (size_t)i * C::WB, (uint32_t)C::WB, full + i);
            }
            pdl_wait();
            for (int i = 0; i < npre; ++i)
                tma_bulk_g2s(smem + (size_t)i * C::SB + C::WB, xsrc + (size_t)i * C::XB, (uint32_t)C::XB, full + i);
            for (int i = npre; i < nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
                if (!mbar_wait(empty + s, ph ^ 1u, a.err, 1)) break;
                mbar_expect_tx(full + s, (uint32_t)C::SB);
                tma_bulk_g2s(smem + (size_t)s * C::SB, wsrc + (size_t)i * C::WB, (uint32_t)C::WB, full + s);
                tma_bulk_g2s(smem + (size_t)s * C::SB + C::WB, xsrc + (size_t)i * C::XB, (uint32_t)C::XB, full + s);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(TC_M, NPAD < 16 ? 16 : NPAD);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
                if (!mbar_wait(full + s, ph, a.err, 2)) break;
                tc_fence_after();
                const uint32_t wbase = smem_u32(smem + (size_t)s * C::SB), xbase = wbase + C::WB;
                const uint64_t ad = umma_desc_sw128(wbase), bd = umma_desc_sw128(xbase);
#pragma unroll
                for (int kk = 0; kk < TC_KB / 16; ++kk)      // a K = 16 step = 32 B = 2 descriptor address units
                    umma_bf16(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                umma_commit(empty + s);        // smem slot is free once these MMAs have read it
            }
            umma_commit(tmem_full);            // accumulator complete
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> fp32 partials
        const int q = warp & 3;                // TMEM lane quarter this warp may access
        const bool ok = nkb == 0 || mbar_wait(tmem_full, 0, a.err, 3);
        tc_fence_after();
        const int row = tile * TC_M + q * 32 + lane;
        float *dst = a.P + (size_t)ks * a.B * a.ldp + row;
#pragma unroll 1
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
            float v[16];
            if (nkb > 0 && ok) {
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c0 + j < a.B) dst[(size_t)(c0 + j) * a.ldp] = v[j];
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

template <int NPAD>
inline int launch_tc_gemm_t(const TcGemmArgs &a, int Mtiles, cudaStream_t st) {
    using C = TcCfg<NPAD>;
    static bool configured = false;
    if (!configured) {
        GVX_CUDA(cudaFuncSetAttribute(k_tc_gemm<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    dim3 grid(Mtiles, a.KS);
    GVX_CUDA(launch_pdl(k_tc_gemm<NPAD>, grid, dim3(TC_THREADS), C::SMEM, st, a));
    GVX_LAUNCHED(1);
    GVX_CUDA(cudaGetLastError());
    return 0;
}

inline int tc_npad(int B) { return B <= 16 ? 16 : (B <= 32 ? 32 : (B <= 64 ? 64 : (B <= 128 ? 128 : 0))); }

inline int launch_tc_gemm(const TcGemmArgs &a, int Mtiles, cudaStream_t st) {
    GVX_CHECK(a.Kpad % TC_KB == 0 && a.Kpad > 0, "tc gemm: K must be padded to a multiple of 64");
    GVX_CHECK(a.KS >= 1 && a.KS <= a.Kpad / TC_KB, "tc gemm: bad split-K factor");
    GVX_CHECK(((uintptr_t)a.Wimg & 15) == 0 && ((uintptr_t)a.Ximg & 15) == 0, "tc gemm: operand images must be 16-byte aligned");
    switch (tc_npad(a.B)) {
        case 16: return launch_tc_gemm_t<16>(a, Mtiles, st);
        case 32: return launch_tc_gemm_t<32>(a, Mtiles, st);
        case 64: return launch_tc_gemm_t<64>(a, Mtiles, st);
        case 128: return launch_tc_gemm_t<128>(a, Mtiles, st);
        default: return fail("tc gemm: batch rows per call must be <= 128 in bf16 mode");
    }
}

// ------------------------------------------------------------------ gate GEMM + LSTM cell in one kernel
// Same pipeline as k_tc_gemm, but the 4 CTAs that split the K range of one 128-row tile form a thread-block
// cluster: each dumps its fp32 partial accumulator from TMEM to its own shared memory, the cluster syncs, and
// CTA j finishes hidden units [8j, 8j+8) of the tile — sums the 4 partials (own + 3 through distributed shared
// memory, always in rank order), adds the bias, applies nn.LSTMCell's pointwise part (tacotron2.py:340,:357) and the
// carried-state dropout (:341,:358), and writes h (bf16) into the next GEMMs' operand images.  No partials in HBM,
// one launch less per LSTM step.
struct TcLstmArgs {
    TcGemmArgs g;              // P unused; KS must be 4
    const float *bias;         // [4*HID] unit-major (b_ih + b_hh)
    const float *c_prev;       // [B, HID]
    float *c_out;              // [B, HID]
    float *gates_out;          // [B, 4*HID] unit-major activations or null
    BfDsts h_dst;
    DropCfg drop;
    uint32_t site, t;
    int row_offset, HID;
};

template <int NPAD>
__global__ void __cluster_dims__(1, 4, 1) __launch_bounds__(TC_THREADS, 1) k_tc_gemm_lstm(const TcLstmArgs p) {
    namespace cg = cooperative_groups;
    using C = TcCfg<NPAD>;
    const TcGemmArgs &a = p.g;
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment computed as an OFFSET into the shared array: the compiler keeps the shared address space (LDS/STS,
    // not generic LD/ST) for everything derived from it
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = (uint64_t *)(smem + C::BAR_OFF);
    uint64_t *empty = full + TC_STAGES;
    uint64_t *tmem_full = empty + TC_STAGES;
    uint32_t *tmem_slot = (uint32_t *)(tmem_full + 1);
    cg::cluster_group cluster = cg::this_cluster();

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, ks = blockIdx.y;          // ks == rank in the cluster (cluster spans gridDim.y = 4)
    const int nkb_all = a.Kpad / TC_KB;
    const int kb0 = (int)((long long)ks * nkb_all / 4), kb1 = (int)((long long)(ks + 1) * nkb_all / 4);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    float *red = reinterpret_cast<float *>(smem);          // [NPAD][128] fp32 partial, reuses the drained pipeline stages

    if (warp == 0) {
        if (elect_one()) {
            const uint8_t *wsrc = (const uint8_t *)a.Wimg + ((size_t)tile * nkb_all + kb0) * C::WB;
            const uint8_t *xsrc = (const uint8_t *)a.Ximg + (size_t)kb0 * C::XB;
            const int npre = nkb < TC_STAGES ? nkb : TC_STAGES;
            for (int i = 0; i < npre; ++i) {
                mbar_expect_tx(full + i, (uint32_t)C::SB);
                tma_bulk_g2s(smem + (size_t)i * C::SB, wsrc + (size_t)i * C::WB, (uint32_t)C::WB, full + i);
            }
            pdl_wait();
            for (int i = 0; i < npre; ++i)
                tma_bulk_g2s(smem + (size_t)i * C::SB + C::WB, xsrc + (size_t)i * C::XB, (uint32_t)C::XB, full + i);
            for (int i = npre; i < nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
                if (!mbar_wait(empty + s, ph ^ 1u, a.err, 1)) break;
                mbar_expect_tx(full + s, (uint32_t)C::SB);
                tma_bulk_g2s(smem + (size_t)s * C::SB, wsrc + (size_t)i * C::WB, (uint32_t)C::WB, full + s);
                tma_bulk_g2s(smem + (size_t)s * C::SB + C::WB, xsrc + (size_t)i * C::XB, (uint32_t)C::XB, full + s);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(TC_M, NPAD < 16 ? 16 : NPAD);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
                if (!mbar_wait(full + s, ph, a.err, 2)) break;
                tc_fence_after();
                const uint32_t wbase = smem_u32(smem + (size_t)s * C::SB), xbase = wbase + C::WB;
                const uint64_t ad = umma_desc_sw128(wbase), bd = umma_desc_sw128(xbase);
#pragma unroll
                for (int kk = 0; kk < TC_KB / 16; ++kk)      // a K = 16 step = 32 B = 2 descriptor address units
                    umma_bf16(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                umma_commit(empty + s);
            }
            umma_commit(tmem_full);
        }
    } else {
        // TMEM -> shared memory partial, transposed ([col][row]) so that the 32 lanes of a warp hit 32 banks
        const int q = warp & 3;
        const bool ok = nkb == 0 || mbar_wait(tmem_full, 0, a.err, 3);
        tc_fence_after();
        const int row = q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
            float v[16];
            if (nkb > 0 && ok) {
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) red[(c0 + j) * TC_M + row] = v[j];
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster.sync();                                        // all four partials are in shared memory

    {
        const float *r0 = cluster.map_shared_rank(red, 0), *r1 = cluster.map_shared_rank(red, 1);
        const float *r2 = cluster.map_shared_rank(red, 2), *r3 = cluster.map_shared_rank(red, 3);
        const int HID = p.HID;
        for (int idx = threadIdx.x; idx < 8 * a.B; idx += TC_THREADS) {
            const int l = ks * 8 + (idx & 7), b = idx >> 3;
            const int unit = tile * 32 + l;
            if (unit >= HID) continue;
            float pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int o = b * TC_M + g * 32 + l;
                pre[g] = ((r0[o] + r1[o]) + r2[o]) + r3[o];
            }
            const float4 bi = *reinterpret_cast<const float4 *>(p.bias + 4 * unit);
            const float gi = sigmoidf_(pre[0] + bi.x), gf = sigmoidf_(pre[1] + bi.y);
            const float gg = tanhf(pre[2] + bi.z), go = sigmoidf_(pre[3] + bi.w);
            const size_t ci = (size_t)b * HID + unit;
            const float cn = gf * p.c_prev[ci] + gi * gg;
            p.c_out[ci] = cn;
            const float h = go * tanhf(cn) * drop_mult(p.drop, p.site, p.t, (uint32_t)(b + p.row_offset), (uint32_t)unit);
            if (p.gates_out) *reinterpret_cast<float4 *>(p.gates_out + (size_t)b * 4 * HID + 4 * unit) = make_float4(gi, gf, gg, go);
            bf_store1(p.h_dst, b, unit, h);
        }
    }
    cluster.sync();                                        // peers may still be reading this CTA's partial
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

template <int NPAD>
inline int launch_tc_gemm_lstm_t(const TcLstmArgs &p, int Mtiles, cudaStream_t st) {
    using C = TcCfg<NPAD>;
    static bool configured = false;
    if (!configured) {
        GVX_CUDA(cudaFuncSetAttribute(k_tc_gemm_lstm<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    GVX_CUDA(launch_pdl(k_tc_gemm_lstm<NPAD>, dim3(Mtiles, 4), dim3(TC_THREADS), C::SMEM, st, p));
    GVX_LAUNCHED(1);
    return 0;
}

inline bool tc_fused_lstm_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("GVX_FUSED_LSTM");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

inline int launch_tc_gemm_lstm(const TcLstmArgs &p, int Mtiles, cudaStream_t st) {
    GVX_CHECK(p.g.Kpad % TC_KB == 0 && p.g.Kpad / TC_KB >= 4, "fused tc gemm: K must cover at least 4 k-blocks");
    switch (tc_npad(p.g.B)) {
        case 16: return launch_tc_gemm_lstm_t<16>(p, Mtiles, st);
        case 32: return launch_tc_gemm_lstm_t<32>(p, Mtiles, st);
        case 64: return launch_tc_gemm_lstm_t<64>(p, Mtiles, st);
        case 128: return launch_tc_gemm_lstm_t<128>(p, Mtiles, st);
        default: return fail("tc gemm: batch rows per call must be <= 128 in bf16 mode");
    }
}

// ------------------------------------------------------------------ operand image builders
// generic weight image: element (tile, k-block, r, swizzled k) <- src(m = tile*128 + r, k), zero outside
// mode 0: src[m*ld + k]                               (plain [Mtot, K])
// mode 1: LSTM forward: m = tile*128 + g*32 + l  -> torch row g*HID + 32*tile + l; k < Kih ? w_ih : w_hh
// mode 2: LSTM backward (transposed): m = input feature, k = 4*u + g -> torch row g*HID + u, column m
// mode 3: plain transposed: src[k*ld + m]
struct TcPackW {
    const float *s0;      // w / w_ih
    const float *s1;      // w_hh (modes 1, 2)
    int mode, Mtot, K, ld, HID, Kih;
};
__global__ void k_tc_pack_w(const TcPackW p, int Mtiles, int Kpad, __nv_bfloat16 *__restrict__ img) {
    const size_t total = (size_t)Mtiles * Kpad * TC_M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        // image index -> (tile, k-block, row, chunk position, element); the chunk position is the swizzled one
        const int j = (int)(i & 7), cpos = (int)((i >> 3) & 7);
        const int r = (int)((i >> 6) & (TC_M - 1));
        const size_t rest = i >> 13;                       // tile * (Kpad/64) + kb
        const int kb = (int)(rest % (Kpad / 64)), tile = (int)(rest / (Kpad / 64));
        const int m = tile * TC_M + r, k = kb * 64 + ((cpos ^ (r & 7)) << 3) + j;
        float v = 0.f;
        if (p.mode == 0) {
            if (m < p.Mtot && k < p.K) v = p.s0[(size_t)m * p.ld + k];
        } else if (p.mode == 3) {
            if (m < p.Mtot && k < p.K) v = p.s0[(size_t)k * p.ld + m];
        } else if (p.mode == 1) {
            const int g = r >> 5, l = r & 31, unit = tile * 32 + l;
            if (unit < p.HID && k < p.K) {
                const size_t row = (size_t)g * p.HID + unit;
                v = k < p.Kih ? p.s0[row * p.Kih + k] : p.s1[row * p.HID + (k - p.Kih)];
            }
        } else {
            const int u = k >> 2, g = k & 3;
            if (m < p.Mtot && u < p.HID) {
                const size_t row = (size_t)g * p.HID + u;
                v = m < p.Kih ? p.s0[row * p.Kih + m] : p.s1[row * p.HID + (m - p.Kih)];
            }
        }
        img[i] = __float2bfloat16(v);
    }
}

// activation image from a row-major fp32 matrix [B, K] (ld): element (k-block, b, swizzled k) <- X[b, k]
__global__ void k_tc_pack_x(const float *__restrict__ X, int B, int K, int ld, int NPAD, int Kpad, __nv_bfloat16 *__restrict__ img) {
    const size_t total = (size_t)Kpad * NPAD;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i & 7), cpos = (int)((i >> 3) & 7);
        const size_t rest = i >> 6;
        const int b = (int)(rest % NPAD), kb = (int)(rest / NPAD);
        const int k = kb * 64 + ((cpos ^ (b & 7)) << 3) + j;
        img[i] = __float2bfloat16((b < B && k < K) ? X[(size_t)b * ld + k] : 0.f);
    }
}

// out[b][m] = sum_ks P[ks][b][m]  (fixed order)
__global__ void k_tc_sum_partials(const float *__restrict__ P, int KS, int B, int ldp, int M, float *__restrict__ out, int ldo) {
    const size_t total = (size_t)B * M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M), b = (int)(i / M);
        float s = 0.f;
        for (int k = 0; k < KS; ++k) s += P[((size_t)k * B + b) * ldp + m];
        out[(size_t)b * ldo + m] = s;
    }
}

}  // namespace gvx
