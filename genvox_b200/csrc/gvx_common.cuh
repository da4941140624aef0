// genvox_b200 — shared device helpers: error plumbing, Philox4x32-10 dropout stream, math.
//
// The dropout stream replaces torch's RNG for F.dropout at
// /root/reference/models/tts/tacotron2.py:143 (prenet, always on), :341 and :358
// (carried LSTM state).  Host mirror used by the oracle: oracle/philox.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace gvx {

// ---------------------------------------------------------------- errors
extern thread_local char g_err[512];
inline int fail(const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return 1;
}
#define GVX_CHECK(cond, msg)                                                                   \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            snprintf(gvx::g_err, sizeof(gvx::g_err), "%s:%d: %s", __FILE__, __LINE__, (msg));  \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)
#define GVX_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            snprintf(gvx::g_err, sizeof(gvx::g_err), "%s:%d: %s -> %s", __FILE__, __LINE__,    \
                     #expr, cudaGetErrorString(e__));                                          \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)
#define GVX_TRY(expr)                                                                          \
    do {                                                                                       \
        int r__ = (expr);                                                                      \
        if (r__) return r__;                                                                   \
    } while (0)

// ---------------------------------------------------------------- device-side abort latch
// The persistent / tcgen05 kernels never hang: a wait that times out writes a code into the call's device error word and
// the grid drains.  Training calls are fully asynchronous (no host sync), so that word is latched - by a one-thread kernel
// at the end of every call - into a sticky word in mapped pinned host memory; the NEXT entry point (or an explicit
// gvx_device_error()) sees it without synchronising and fails loudly.  where: 1 train_fwd, 2 train_bwd, 3 infer.
inline int *&err_latch_host() { static int *p = nullptr; return p; }
inline int *&err_latch_dev() { static int *p = nullptr; return p; }
__global__ void k_latch_err(const int *__restrict__ err, int nwords, int *host_latch, int where) {
    for (int i = 0; i < nwords; ++i) {
        const int e = err[i];
        if (e != 0 && *reinterpret_cast<volatile int *>(host_latch) == 0) *reinterpret_cast<volatile int *>(host_latch) = where * 1000 + e;
    }
}
inline int latch_ready() {
    if (err_latch_host()) return 0;
    int *h = nullptr, *dv = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&dv, h, 0) != cudaSuccess)
        return fail("could not allocate the pinned error latch");
    *h = 0;
    err_latch_host() = h;
    err_latch_dev() = dv;
    return 0;
}
// refuse to enqueue more work once a previous call of this process aborted on the device
inline int latch_check() {
    if (latch_ready()) return 1;
    const int e = *reinterpret_cast<volatile int *>(err_latch_host());
    if (e != 0) {
        snprintf(g_err, sizeof(g_err), "a previous launch aborted on the device (entry %d, wait code %d: a persistent chain lost its "
                 "co-resident CTAs or a pipeline stalled); its outputs are invalid", e / 1000, e % 1000);
        return 1;
    }
    return 0;
}
inline int latch_record(const int *err_dev, int nwords, int where, cudaStream_t st) {
    if (latch_ready()) return 1;
    k_latch_err<<<1, 1, 0, st>>>(err_dev, nwords, err_latch_dev(), where);
    return cudaGetLastError() == cudaSuccess ? 0 : fail("latch kernel launch failed");
}

// ---------------------------------------------------------------- launch accounting / phase profiler
// g_launches counts every kernel this library launches (bench.py's `gpu_launches`).  The profiler,
// when enabled, brackets each phase launch with CUDA events on the launching stream so bench.py can
// report the average duration of the dominant kernel measured inside a real step.
inline unsigned long long g_launches = 0;
#define GVX_LAUNCHED(n) (gvx::g_launches += (n))

enum ProfSlot {
    PS_SETUP = 0, PS_PRENET, PS_ATT_LSTM, PS_QUERY, PS_ATTENTION, PS_DEC_LSTM, PS_PROJ, PS_OUTPUT,
    PS_BWD_DEC_POINT, PS_BWD_DEC_GEMM, PS_BWD_ATTENTION, PS_BWD_ATT_POINT, PS_BWD_ATT_GEMM, PS_BWD_BATCHED, PS_DEC_IN_GEMM, PS_NSLOT
};
inline const char *prof_slot_name(int s) {
    static const char *names[PS_NSLOT] = {"setup", "prenet", "att_lstm", "query", "attention", "dec_lstm", "proj", "output",
                                          "bwd_dec_pointwise", "bwd_dec_gemm", "bwd_attention", "bwd_att_pointwise_query",
                                          "bwd_att_gemm", "bwd_time_batched", "dec_lstm_input_gemm"};
    return (s >= 0 && s < PS_NSLOT) ? names[s] : nullptr;
}
struct ProfRec { int slot; cudaEvent_t a, b; };
struct ProfState {
    int on = 0;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    double total_ms[PS_NSLOT] = {0};
    long long count[PS_NSLOT] = {0};
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void collect() {          // synchronises on every recorded stop event
        for (auto &r : recs) {
            float ms = 0.f;
            if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                total_ms[r.slot] += ms;
                count[r.slot] += 1;
            }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
};
inline ProfState g_prof;
struct ProfScope {
    int active;
    ProfRec r;
    cudaStream_t st;
    ProfScope(int slot, cudaStream_t s) : active(g_prof.on), st(s) {
        if (active) { r.slot = slot; r.a = g_prof.get(); r.b = g_prof.get(); cudaEventRecord(r.a, st); }
    }
    ~ProfScope() {
        if (active) { cudaEventRecord(r.b, st); g_prof.recs.push_back(r); }
    }
};

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the per-step chain is launched with cudaLaunchAttributeProgrammaticStreamSerialization:
// it signals `launch_dependents` at once, so the next kernel's CTAs become resident early and run their
// prologue (barrier init, TMEM allocation, weight prefetch — nothing the previous kernel writes), then block
// in `griddepcontrol.wait` until the previous grid has completed and flushed.  GVX_NO_PDL=1 disables it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("GVX_NO_PDL");
        on = (e && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
}

// The kernels of a chain read, BEFORE griddepcontrol.wait, only data that was complete before the chain began
// (weights, processed memory, forward stashes).  Kernels ahead of the chain (prenet / processed-memory GEMMs) also
// trigger early, so the first kernel of every chain is launched with a full dependency: pdl_barrier_next().
inline thread_local bool g_pdl_skip_next = false;
inline void pdl_barrier_next() { g_pdl_skip_next = true; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_enabled() && !g_pdl_skip_next) ? 1 : 0;
    g_pdl_skip_next = false;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---------------------------------------------------------------- Philox4x32-10
enum : uint32_t { SITE_PRENET0 = 0, SITE_PRENET1 = 1, SITE_ATT = 2, SITE_DEC = 3 };

struct DropCfg {
    const uint32_t *kptr; // if non-null: {seed low, seed high} in device memory (lets a CUDA graph outlive the seed)
    uint32_t k0, k1;      // seed low / high word
    uint32_t threshold;   // keep iff word >= threshold ( = floor(p * 2^32) )
    float scale;          // 1 / (1 - p)
    int on;               // 0: identity
};

// device seed slot of the call being enqueued (set by the entry points around their launch sequence)
inline thread_local const uint32_t *g_seed_ptr = nullptr;

inline DropCfg make_drop(uint64_t seed, float p, int on, const uint32_t *kptr = nullptr) {
    DropCfg c;
    c.kptr = kptr ? kptr : g_seed_ptr;
    c.k0 = (uint32_t)(seed & 0xffffffffull);
    c.k1 = (uint32_t)(seed >> 32);
    double t = (double)p * 4294967296.0;
    c.threshold = (uint32_t)t;           // floor (p < 1)
    c.scale = (float)(1.0 / (1.0 - (double)p));
    c.on = (on && p > 0.f) ? 1 : 0;
    return c;
}

__device__ __forceinline__ uint32_t philox_word(uint32_t j, uint32_t row, uint32_t t, uint32_t site,
                                                uint32_t k0, uint32_t k1) {
    uint32_t c0 = j >> 2, c1 = row, c2 = t, c3 = site;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const uint32_t sel = j & 3u;
    return sel == 0 ? c0 : (sel == 1 ? c1 : (sel == 2 ? c2 : c3));
}

// multiplier applied to element (row, j) of dropout site `site` at index t: 0 or scale
__device__ __forceinline__ float drop_mult(const DropCfg &c, uint32_t site, uint32_t t, uint32_t row, uint32_t j) {
    if (!c.on) return 1.f;
    const uint32_t k0 = c.kptr ? c.kptr[0] : c.k0, k1 = c.kptr ? c.kptr[1] : c.k1;
    return philox_word(j, row, t, site, k0, k1) >= c.threshold ? c.scale : 0.f;
}

// ---------------------------------------------------------------- math
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide reductions; `scratch` holds >= 33 floats; every thread gets the result
__device__ __forceinline__ float block_sum(float v, float *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float x = lane < nw ? scratch[lane] : 0.f;
        x = warp_sum(x);
        if (lane == 0) scratch[32] = x;
    }
    __syncthreads();
    return scratch[32];
}
__device__ __forceinline__ float block_max(float v, float *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float x = lane < nw ? scratch[lane] : -INFINITY;
        x = warp_max(x);
        if (lane == 0) scratch[32] = x;
    }
    __syncthreads();
    return scratch[32];
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;       // src-size 0 => 16 bytes of zero fill, no global read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

}  // namespace gvx
