// genvox_b200 — thin row-major wrappers over cuBLAS for the PLAIN time-batched GEMMs (weight gradients,
// projections over all frames).  The recurrent chain never calls cuBLAS.
#pragma once
#include <cublas_v2.h>

#include "gvx_common.cuh"

namespace gvx {

#define GVX_CUBLAS(expr)                                                                             \
    do {                                                                                             \
        cublasStatus_t s__ = (expr);                                                                 \
        if (s__ != CUBLAS_STATUS_SUCCESS) {                                                          \
            snprintf(gvx::g_err, sizeof(gvx::g_err), "%s:%d: %s -> cublas status %d", __FILE__, __LINE__, #expr, (int)s__); \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

inline int blas(cublasHandle_t *out, cudaStream_t st) {
    static thread_local cublasHandle_t h = nullptr;
    if (!h) {
        GVX_CUBLAS(cublasCreate(&h));
        GVX_CUBLAS(cublasSetMathMode(h, CUBLAS_DEFAULT_MATH));    // true fp32 sgemm (no TF32), like torch's default
        // a fixed workspace keeps cuBLAS from allocating while its GEMMs are being captured into a CUDA graph
        void *ws = nullptr;
        if (cudaMalloc(&ws, (size_t)64 << 20) == cudaSuccess) GVX_CUBLAS(cublasSetWorkspace(h, ws, (size_t)64 << 20));
    }
    GVX_CUBLAS(cublasSetStream(h, st));
    *out = h;
    return 0;
}

// row-major helpers: C[M,N] (ldc) = alpha * op(A) . op(B) + beta * C
// NN: A [M,K] (lda), B [K,N] (ldb)
inline int gemm_nn(cudaStream_t st, int M, int N, int K, const float *A, int lda, const float *B, int ldb, float *C, int ldc,
                   float beta) {
    cublasHandle_t h;
    GVX_TRY(blas(&h, st));
    const float alpha = 1.f;
    GVX_CUBLAS(cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, M, K, &alpha, B, ldb, A, lda, &beta, C, ldc));
    return 0;
}
// TN: A [K,M] (lda), B [K,N] (ldb):  C = A^T . B   (weight gradients: sum over the K = T*B rows)
inline int gemm_tn(cudaStream_t st, int M, int N, int K, const float *A, int lda, const float *B, int ldb, float *C, int ldc,
                   float beta) {
    cublasHandle_t h;
    GVX_TRY(blas(&h, st));
    const float alpha = 1.f;
    GVX_CUBLAS(cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, M, K, &alpha, B, ldb, A, lda, &beta, C, ldc));
    return 0;
}
// NT: A [M,K] (lda), W [N,K] (ldw):  C = A . W^T   (a linear layer over many rows)
inline int sgemm_nt(cudaStream_t st, int M, int N, int K, const float *A, int lda, const float *W, int ldw, float *C, int ldc) {
    cublasHandle_t h;
    GVX_TRY(blas(&h, st));
    const float alpha = 1.f, beta = 0.f;
    GVX_CUBLAS(cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &alpha, W, ldw, A, lda, &beta, C, ldc));
    return 0;
}
// column sums of a row-major X [rows, ncols] (ld): out[c] = sum_r X[r, c]
inline int colsum(cudaStream_t st, const float *X, int rows, int ncols, int ld, const float *ones, float *out) {
    cublasHandle_t h;
    GVX_TRY(blas(&h, st));
    const float alpha = 1.f, beta = 0.f;
    GVX_CUBLAS(cublasSgemv(h, CUBLAS_OP_N, ncols, rows, &alpha, X, ld, ones, 1, &beta, out, 1));
    return 0;
}

}  // namespace gvx
