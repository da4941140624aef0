// genvox_b200 — the time-batched contractions of training on tcgen05:   C[M, N] (fp32) = A[M, K] . B[N, K]^T   (bf16 operands)
//
// Everything of teacher-forced training that is NOT on the sequential chain is a plain contraction over all T*B frames
// (the reference runs them one decoder step at a time through nn.LSTMCell / nn.Linear and their autograd,
// /root/reference/models/tts/tacotron2.py:340,:357,:361-362 and :520):
//   forward : prenet part of the attention-LSTM gates, input part of the decoder-LSTM gates, mel / gate projections;
//   backward: d [h_att | ctx] through W_ih of the decoder LSTM, d prenet_out through W_ih of the attention LSTM,
//             and every weight gradient  d W = G^T . X  (K = T*B = 51200 at configs[2]: 1.8 TFLOP of the train step).
// One kernel serves all of them: the "NT" form (both operands K-major: weights that are needed transposed are packed transposed
// once, they are static) and the "TN" form for contractions over the frame axis (both operands MN-major, i.e. the frame-major
// activations / gradients exactly as they lie in HBM - tcgen05 takes MN-major shared-memory operands, so there is no transpose
// pass: that pass was 1.0 ms of the train step).
//
// Kernel: persistent, one CTA per SM, 128 x 256 output tiles, K blocks of 64.
//   warp 0  TMA producer  : cp.async.bulk.tensor.2d (tensor maps, SWIZZLE_128B) into a 4-stage ring, OOB rows / K tail zero-filled
//   warp 1  MMA issuer    : tcgen05.mma kind::f16 128 x 256 x 16, fp32 accumulators in TMEM, two accumulator buffers
//   warps 2..5 epilogue   : tcgen05.ld (each warp its 32-lane quadrant) -> fp32 rows of C; drains tile i while tile i+1 is contracted
// Every output element is accumulated in a fixed order (one CTA in k order; for the few small weight gradients whose tiles would
// leave most SMs idle the K blocks are split over CTAs and the partial tiles summed in split order) -> bit-exact run to run.
// Measured on B200 (profiles/nt_gemm_bench.py): 1.31 PFLOP/s on the decoder-LSTM weight gradient (4096 x 2560, K = 51200),
// 1.53 on d [h_att | ctx] (51200 x 1536, K = 4096).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "gvx_tc.cuh"

namespace gvx {

// 128 x 256 tiles (the largest single-CTA UMMA): 87 flop per operand byte pulled from L2, against 64 for 128 x 128.  Measured
// (profiles/nt_gemm_bench.py): 1.3-1.5 PFLOP/s on the long-K shapes with either tile, cuBLAS 1.8-2.0 - every SM ingests ~52 B/cycle
// here, the same per-SM L2 -> SM rate the chains see with cp.async, so the next step is not a bigger single-CTA tile but
// cta_group::2 (a CTA pair sharing its B halves: 131 flop per ingested byte).  Not built.
constexpr int NG_BM = 128, NG_BN = 256, NG_BK = 64;
constexpr int NG_STAGES = 4;
constexpr int NG_STAGE_BYTES = (NG_BM + NG_BN) * NG_BK * 2;       // 48 KB
constexpr int NG_THREADS = 192;

struct NgShared {
    uint64_t full[NG_STAGES], empty[NG_STAGES], tfull[2], tempty[2];
    uint32_t tmem_slot;
};
constexpr size_t NG_SMEM = (size_t)NG_STAGES * NG_STAGE_BYTES + sizeof(NgShared) + 1024;

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
        "%27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct NgArgs {
    float *C;
    int M, N, K, ldc, tiles_m, tiles_n;
    int splits;                  // > 1: the K blocks are dealt to `splits` work items per tile, partial tiles go to P[split][M][N]
    float *P;
    int *err;
};

// MN = false: both operands K-major in memory (A [M, K], B [N, K]);  MN = true: both MN-major (A^T [K, M], B^T [K, N], the frame-major
// activations / gradients of the weight-gradient contractions as they lie in HBM: no transpose pass).  MN-major tiles are loaded as
// 64 (MN elements = 128 B) x 64 (K rows) boxes - one SWIZZLE_128B atom column each, 8 KB apart (the descriptor's leading byte
// offset), 8-row K groups 1 KB apart (stride byte offset), a K = 16 MMA step = 2 KB - and the instruction descriptor carries the
// a_major / b_major bits.
template <bool MN>
__global__ void __launch_bounds__(NG_THREADS, 1) k_nt_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                           const NgArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    NgShared *sh = (NgShared *)(smem + (size_t)NG_STAGES * NG_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = a.tiles_m * a.tiles_n * a.splits, nkb_all = (a.K + NG_BK - 1) / NG_BK;
    // work item -> (output tile, K-block range); without split-K an item is a whole tile
    auto item_of = [&](int item, int &tm, int &tn, int &kb0, int &kb1, int &sp) {
        const int t = item / a.splits;
        sp = item - t * a.splits;
        tm = t / a.tiles_n;
        tn = t - tm * a.tiles_n;
        kb0 = (int)((long long)nkb_all * sp / a.splits);
        kb1 = (int)((long long)nkb_all * (sp + 1) / a.splits);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NG_STAGES; ++s) { mbar_init(sh->full + s, 1); mbar_init(sh->empty + s, 1); }
        for (int k = 0; k < 2; ++k) { mbar_init(sh->tfull + k, 1); mbar_init(sh->tempty + k, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_slot;

    if (warp == 0) {
        if (elect_one()) {      // ---- TMA producer
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // (tiles that share a row block of A run on neighbouring CTAs at the same time: L2 reuse of both operands)
                int tm, tn, kb0, kb1, sp;
                item_of(tile, tm, tn, kb0, kb1, sp);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!mbar_wait(sh->empty + s, ph ^ 1u, a.err, 71)) return;
                    uint8_t *dst = smem + (size_t)s * NG_STAGE_BYTES;
                    mbar_expect_tx(sh->full + s, NG_STAGE_BYTES);
                    if (MN) {
#pragma unroll
                        for (int q = 0; q < NG_BM / 64; ++q) tma_load_2d(dst + q * 8192, &tmA, tm * NG_BM + 64 * q, kb * NG_BK, sh->full + s);
#pragma unroll
                        for (int q = 0; q < NG_BN / 64; ++q)
                            tma_load_2d(dst + NG_BM * NG_BK * 2 + q * 8192, &tmB, tn * NG_BN + 64 * q, kb * NG_BK, sh->full + s);
                    } else {
                        tma_load_2d(dst, &tmA, kb * NG_BK, tm * NG_BM, sh->full + s);
                        tma_load_2d(dst + NG_BM * NG_BK * 2, &tmB, kb * NG_BK, tn * NG_BN, sh->full + s);
                    }
                    if (++s == NG_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {      // ---- MMA issuer
            constexpr uint32_t idesc = umma_idesc_bf16(NG_BM, NG_BN) | (MN ? (1u << 15) | (1u << 16) : 0u);
            constexpr int kstep = MN ? (2048 >> 4) : 2;                 // descriptor advance per K = 16 step, in 16-byte units
            const uint64_t ad0 = MN ? umma_desc_sw128_mn(smem_u32(smem)) : umma_desc_sw128(smem_u32(smem));
            int s = 0, it = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                // the epilogue warps have drained this accumulator buffer (first use of each buffer: immediately true)
                if (!mbar_wait(sh->tempty + acc, (uint32_t)((it >> 1) & 1) ^ 1u, a.err, 72)) return;
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * NG_BN);
                int tm, tn, kb0, kb1, sp;
                item_of(tile, tm, tn, kb0, kb1, sp);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!mbar_wait(sh->full + s, ph, a.err, 73)) return;
                    tc_fence_after();
                    const uint64_t ad = ad0 + (uint64_t)(s * (NG_STAGE_BYTES >> 4)), bd = ad + (uint64_t)((NG_BM * NG_BK * 2) >> 4);
#pragma unroll
                    for (int j = 0; j < NG_BK / 16; ++j) umma_bf16(tacc, ad + kstep * j, bd + kstep * j, idesc, (kb > kb0 || j > 0) ? 1u : 0u);
                    umma_commit(sh->empty + s);
                    if (++s == NG_STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(sh->tfull + acc);
            }
        }
    } else {
        // ---- epilogue: warp w reads TMEM lanes [32 (w & 3), +32) = rows of the tile; 4 x 32 columns per pass
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            int tm, tn, kb0, kb1, sp;
            item_of(tile, tm, tn, kb0, kb1, sp);
            if (!mbar_wait(sh->tfull + acc, (uint32_t)((it >> 1) & 1), a.err, 74)) return;
            tc_fence_after();
            const int m = tm * NG_BM + 32 * q + lane;
            const int ldc = a.splits > 1 ? a.N : a.ldc;
            float *crow = (a.splits > 1 ? a.P + (size_t)sp * a.M * a.N : a.C) + (size_t)m * ldc + tn * NG_BN;
#pragma unroll 1
            for (int c0 = 0; c0 < NG_BN; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * NG_BN + c0), v);
                if (m < a.M) {
                    const int nrem = a.N - (tn * NG_BN + c0);
                    if (nrem >= 32 && (ldc & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(crow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nrem) crow[c0 + i] = v[i];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_local(sh->tempty + acc);
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled ng_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}
// frame-major bf16 matrix [K rows, MN columns] with row stride ld: box = 64 MN-elements (128 B) x 64 K rows, SWIZZLE_128B
inline int ng_tensor_map_mn(CUtensorMap *tm, const __nv_bfloat16 *base, int K, int MNsize, int ld) {
    PFN_encodeTiled fn = ng_encode_fn();
    GVX_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    GVX_CHECK(((uintptr_t)base & 15) == 0 && ld % 8 == 0, "tn_gemm: operands must be 16-byte aligned with a row stride that is a multiple of 8");
    const cuuint64_t dims[2] = {(cuuint64_t)MNsize, (cuuint64_t)K};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)NG_BK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled (MN-major) failed: %d (K %d, MN %d, ld %d)", (int)r, K, MNsize, ld);
        return 1;
    }
    return 0;
}
// row-major bf16 matrix [rows, K] with row stride ld (elements): box = 64 K-elements x `box_rows` rows, SWIZZLE_128B
inline int ng_tensor_map(CUtensorMap *tm, const __nv_bfloat16 *base, int rows, int K, int ld, int box_rows) {
    PFN_encodeTiled fn = ng_encode_fn();
    GVX_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    GVX_CHECK(((uintptr_t)base & 15) == 0 && ld % 8 == 0, "nt_gemm: operands must be 16-byte aligned with a row stride that is a multiple of 8");
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)NG_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed: %d (rows %d, K %d, ld %d)", (int)r, rows, K, ld);
        return 1;
    }
    return 0;
}

// C[m][n] = sum over splits of P[s][m][n], ascending s (deterministic)
__global__ void __launch_bounds__(256) k_ng_reduce_splits(const float *__restrict__ P, int splits, int M, int N, float *__restrict__ C, int ldc) {
    const size_t total = (size_t)M * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += P[(size_t)k * total + i];
        C[(i / N) * ldc + (i % N)] = s;
    }
}

// C[M, N] (ldc) = A[M, K] (lda) . B[N, K]^T (ldb); K may be any multiple of 8 (the tail of the last K block is zero-filled by TMA).
// `ws` (optional, `ws_floats` floats): when the tiles alone would leave most SMs idle and K is long (the small weight gradients over
// all frames), the K blocks are split over several CTAs per tile; the partial tiles are summed in a fixed order afterwards.
inline int ng_gemm_bf16(bool mn, cudaStream_t st, int M, int N, int K, const __nv_bfloat16 *A, int lda, const __nv_bfloat16 *B, int ldb,
                       float *C, int ldc, int *err, float *ws, size_t ws_floats) {
    GVX_CHECK(M > 0 && N > 0 && K > 0 && (mn || K % 8 == 0), "nt_gemm: bad shape");
    CUtensorMap tmA, tmB;
    if (mn) {
        GVX_TRY(ng_tensor_map_mn(&tmA, A, K, M, lda));
        GVX_TRY(ng_tensor_map_mn(&tmB, B, K, N, ldb));
    } else {
        GVX_TRY(ng_tensor_map(&tmA, A, M, K, lda, NG_BM));
        GVX_TRY(ng_tensor_map(&tmB, B, N, K, ldb, NG_BN));
    }
    NgArgs a;
    a.C = C; a.M = M; a.N = N; a.K = K; a.ldc = ldc; a.err = err;
    a.tiles_m = (M + NG_BM - 1) / NG_BM; a.tiles_n = (N + NG_BN - 1) / NG_BN;
    a.splits = 1; a.P = nullptr;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        GVX_CUDA(cudaFuncSetAttribute(k_nt_gemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NG_SMEM));
        GVX_CUDA(cudaFuncSetAttribute(k_nt_gemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NG_SMEM));
    }
    int ntiles = a.tiles_m * a.tiles_n;
    const int nkb = (K + NG_BK - 1) / NG_BK;
    if (ws && nkb >= 128) {
        // split K when that fills the last wave of the persistent grid (tiles x splits just under a multiple of the SM count) or
        // when the tiles alone leave most SMs idle: each split keeps >= 64 K blocks, the partial tiles must fit the workspace
        auto eff = [&](int sp) {
            const int items = ntiles * sp, waves = (items + sms - 1) / sms;
            return (double)items / ((double)waves * sms);
        };
        int best = 1;
        double best_eff = eff(1);
        for (int sp = 2; sp <= 64 && nkb / sp >= 64; ++sp) {
            if ((size_t)sp * M * N > ws_floats) break;
            const double e = eff(sp);
            if (e > best_eff + 0.03) { best = sp; best_eff = e; }
        }
        if (best > 1) { a.splits = best; a.P = ws; ntiles *= best; }
    }
    if (mn) k_nt_gemm<true><<<ntiles < sms ? ntiles : sms, NG_THREADS, NG_SMEM, st>>>(tmA, tmB, a);
    else k_nt_gemm<false><<<ntiles < sms ? ntiles : sms, NG_THREADS, NG_SMEM, st>>>(tmA, tmB, a);
    GVX_LAUNCHED(1);
    GVX_CUDA(cudaGetLastError());
    if (a.splits > 1) {
        const size_t total = (size_t)M * N;
        k_ng_reduce_splits<<<(unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, st>>>(ws, a.splits, M, N, C, ldc);
        GVX_LAUNCHED(1);
        GVX_CUDA(cudaGetLastError());
    }
    return 0;
}

// C[M, N] = A[M, K] . B[N, K]^T   (both operands K-major)
inline int nt_gemm_bf16(cudaStream_t st, int M, int N, int K, const __nv_bfloat16 *A, int lda, const __nv_bfloat16 *B, int ldb, float *C,
                        int ldc, int *err, float *ws = nullptr, size_t ws_floats = 0) {
    return ng_gemm_bf16(false, st, M, N, K, A, lda, B, ldb, C, ldc, err, ws, ws_floats);
}
// C[M, N] = At[K, M]^T . Bt[K, N]   (both operands frame-major as they lie in HBM: the weight gradients d W = G^T . X)
inline int tn_gemm_bf16(cudaStream_t st, int M, int N, int K, const __nv_bfloat16 *At, int lda, const __nv_bfloat16 *Bt, int ldb, float *C,
                        int ldc, int *err, float *ws = nullptr, size_t ws_floats = 0) {
    return ng_gemm_bf16(true, st, M, N, K, At, lda, Bt, ldb, C, ldc, err, ws, ws_floats);
}

}  // namespace gvx
