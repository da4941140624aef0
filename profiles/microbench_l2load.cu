// Microbenchmark behind the persistent-kernel design: how fast can every SM pull the SAME 128-196 KB activation
// image (L2 resident, rewritten by other SMs between reads) into shared memory?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mb_l2load profiles/microbench_l2load.cu
// Modes: 0 = cp.async.bulk (one elected thread, chunk bytes given), 1 = ld.global.v4 + st.shared by all threads,
//        2 = cp.async 16 B (LDGSTS) by all threads.   Prints cycles per image load (median over CTAs, max over CTAs).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Args {
    uint8_t *img;        // [2][bytes] ping-pong image
    unsigned *bar;
    long long *out;      // [grid][iters] cycles of the load
    int bytes, chunk, mode, iters, rotate, rewrite;
};

__global__ void __launch_bounds__(512, 1) k_load(const Args a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mb;
    const int tid = threadIdx.x, j = blockIdx.x;
    const unsigned ncta = gridDim.x;
    if (tid == 0) {
        mbar_init(&mb, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nchunk = a.bytes / a.chunk;
    for (int it = 0; it < a.iters; ++it) {
        const uint8_t *src = a.img + (size_t)(it & 1) * a.bytes;
        // grid barrier (previous writes of every CTA visible)
        if (tid == 0 && it > 0) {
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory"); } while (v < ncta * (unsigned)it);
        }
        __syncthreads();
        const long long t0 = clock64();
        if (a.mode == 0) {
            if (tid == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");
                mbar_expect_tx(&mb, (uint32_t)a.bytes);
                for (int c = 0; c < nchunk; ++c) {
                    const int pc = a.rotate ? (c + j) % nchunk : c;
                    tma_bulk_g2s(sm + (size_t)pc * a.chunk, src + (size_t)pc * a.chunk, (uint32_t)a.chunk, &mb);
                }
            }
            while (!mbar_try_wait(&mb, (uint32_t)it & 1u)) {}
        } else if (a.mode == 1) {
            const int n16 = a.bytes / 16;
            const int off = a.rotate ? (j * 97) % n16 : 0;
#pragma unroll 8
            for (int i = tid; i < n16; i += 512) {
                int k = i + off;
                if (k >= n16) k -= n16;
                uint4 v;
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + (size_t)k * 16));
                *reinterpret_cast<uint4 *>(sm + (size_t)k * 16) = v;
            }
            __syncthreads();
        } else {
            const int n16 = a.bytes / 16;
            const int off = a.rotate ? (j * 97) % n16 : 0;
            for (int i = tid; i < n16; i += 512) {
                int k = i + off;
                if (k >= n16) k -= n16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + (size_t)k * 16)), "l"(src + (size_t)k * 16));
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
        const long long t1 = clock64();
        if (tid == 0) a.out[(size_t)j * a.iters + it] = t1 - t0;
        // every CTA rewrites its slice of the OTHER half (like the recurrence does), then arrives
        if (a.rewrite) {
            uint8_t *dst = a.img + (size_t)((it + 1) & 1) * a.bytes;
            const int per = a.bytes / (int)ncta;
            for (int i = tid * 16; i < per; i += 512 * 16) {
                uint4 v = *reinterpret_cast<uint4 *>(sm + (size_t)j * per + i);
                v.x += it;
                *reinterpret_cast<uint4 *>(dst + (size_t)j * per + i) = v;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
    }
}

int main() {
    const int grid = 128, iters = 60;
    uint8_t *img;
    unsigned *bar;
    long long *out;
    const int maxbytes = 196608;
    cudaMalloc(&img, 2 * maxbytes);
    cudaMemset(img, 1, 2 * maxbytes);
    cudaMalloc(&bar, 64);
    cudaMalloc(&out, sizeof(long long) * grid * iters);
    cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, maxbytes + 1024);
    std::vector<long long> h(grid * iters);
    struct Cfg { int bytes, chunk, mode, rotate, rewrite; };
    const Cfg cfgs[] = {
        {131072, 8192, 0, 0, 1},  {131072, 8192, 0, 1, 1},  {131072, 32768, 0, 0, 1}, {131072, 2048, 0, 0, 1},
        {131072, 8192, 0, 0, 0},  {131072, 16, 1, 0, 1},    {131072, 16, 1, 1, 1},    {131072, 16, 1, 0, 0},
        {131072, 16, 2, 0, 1},    {131072, 16, 2, 1, 1},    {65536, 8192, 0, 0, 1},   {65536, 16, 1, 0, 1},
        {65536, 16, 2, 0, 1},     {196608, 8192, 0, 0, 1},  {196608, 16, 2, 0, 1},    {16384, 16, 1, 0, 1},
        {16384, 8192, 0, 0, 1},
    };
    for (const Cfg &c : cfgs) {
        cudaMemset(bar, 0, 64);
        Args a{img, bar, out, c.bytes, c.chunk, c.mode, iters, c.rotate, c.rewrite};
        k_load<<<grid, 512, maxbytes + 1024>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), out, sizeof(long long) * grid * iters, cudaMemcpyDeviceToHost);
        std::vector<long long> med, mx;
        for (int it = 10; it < iters; ++it) {
            std::vector<long long> v;
            for (int j = 0; j < grid; ++j) v.push_back(h[(size_t)j * iters + it]);
            std::sort(v.begin(), v.end());
            med.push_back(v[grid / 2]);
            mx.push_back(v[grid - 1]);
        }
        std::sort(med.begin(), med.end());
        std::sort(mx.begin(), mx.end());
        const double m = (double)med[med.size() / 2], M = (double)mx[mx.size() / 2];
        printf("bytes %6d chunk %5d mode %d rotate %d rewrite %d : median CTA %7.0f cyc (%5.1f B/cyc/SM), slowest CTA %7.0f cyc (%6.0f B/cyc chip)\n",
               c.bytes, c.chunk, c.mode, c.rotate, c.rewrite, m, c.bytes / m, M, (double)c.bytes * grid / M);
    }
    return 0;
}
