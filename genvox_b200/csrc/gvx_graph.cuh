// genvox_b200 — CUDA-graph cache for the per-call launch sequences.
//
// A teacher-forced step at BASELINE configs[2] is ~10^4 kernel launches; enqueueing them from the host costs
// about as much time as the GPU needs to run them (bench.py: host_enqueue_ms_per_step ~ ms_per_step).  The
// launch sequence of a call is a pure function of its arguments (shapes, pointers, flags) — the dropout seed
// is the only per-call scalar and lives in device memory — so the second time a call with the same key is seen
// it is captured into a CUDA graph (on a library-owned capture stream; the caller's stream may be the legacy
// default stream, which cannot be captured) and from then on replayed with one cudaGraphLaunch on the caller's
// stream.  The first sighting of a key runs eagerly, which also performs every one-time cudaFuncSetAttribute.
// OPT-IN (GVX_GRAPHS=1).  Measured on B200 (round 1): stream launches are NOT the bottleneck — the host merely
// runs ahead until the launch queue is full — and inside a captured graph the programmatic-dependent-launch
// overlap is lost (replay 128.8 ms/step = the GVX_NO_PDL figure, vs 83.9 ms/step for eager PDL launches), so the
// eager path stays the default.  The phase profiler (gvx_profile_enable) forces eager execution.
#pragma once
#include <functional>
#include <list>
#include <string>

#include "gvx_common.cuh"

namespace gvx {

struct GraphStats {
    unsigned long long eager = 0, captured = 0, replayed = 0, capture_failed = 0;
};
inline GraphStats g_graph_stats;

struct GraphEntry {
    std::string key;
    int sightings = 0;
    bool failed = false;
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches = 0;     // kernels of this library inside the graph (for gvx_launch_count)
};

inline std::list<GraphEntry> &graph_cache() {
    static std::list<GraphEntry> c;
    return c;
}

inline bool graphs_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("GVX_GRAPHS");
        on = (e && e[0] == '1') ? 1 : 0;      // opt-in: see the note at the top of this file
    }
    return on == 1;
}

inline cudaStream_t capture_stream() {
    static cudaStream_t cs = nullptr;
    if (!cs) cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    return cs;
}

struct KeyBuilder {
    std::string k;
    template <class T>
    KeyBuilder &add(const T &v) {
        k.append(reinterpret_cast<const char *>(&v), sizeof(T));
        return *this;
    }
};

// Run `body(stream)` — which only ENQUEUES work on the stream it is given — eagerly on `user`, or as a cached graph.
inline int run_cached(const std::string &key, cudaStream_t user, const std::function<int(cudaStream_t)> &body) {
    if (!graphs_enabled() || g_prof.on) {
        g_graph_stats.eager++;
        return body(user);
    }
    auto &cache = graph_cache();
    auto it = cache.begin();
    for (; it != cache.end(); ++it)
        if (it->key == key) break;
    if (it == cache.end()) {
        if (cache.size() >= 24) {
            if (cache.back().exec) cudaGraphExecDestroy(cache.back().exec);
            cache.pop_back();
        }
        cache.emplace_front();
        cache.front().key = key;
        it = cache.begin();
    } else if (it != cache.begin()) {
        cache.splice(cache.begin(), cache, it);      // most recently used first
        it = cache.begin();
    }
    GraphEntry &e = *it;
    e.sightings++;
    if (e.exec) {
        GVX_CUDA(cudaGraphLaunch(e.exec, user));
        g_launches += e.launches;
        g_graph_stats.replayed++;
        return 0;
    }
    if (e.sightings < 2 || e.failed) {
        g_graph_stats.eager++;
        return body(user);
    }
    // second sighting: capture
    cudaStream_t cs = capture_stream();
    const unsigned long long l0 = g_launches;
    cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        e.failed = true;
        g_graph_stats.capture_failed++;
        return body(user);
    }
    const int rc = body(cs);
    cudaGraph_t graph = nullptr;
    ce = cudaStreamEndCapture(cs, &graph);
    const unsigned long long captured_launches = g_launches - l0;
    g_launches = l0;
    if (rc != 0 || ce != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        e.failed = true;
        g_graph_stats.capture_failed++;
        if (rc != 0) return rc;          // a real argument / launch error: report it
        return body(user);
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess || !exec) {
        cudaGetLastError();
        e.failed = true;
        g_graph_stats.capture_failed++;
        return body(user);
    }
    e.exec = exec;
    e.launches = captured_launches;
    g_graph_stats.captured++;
    GVX_CUDA(cudaGraphLaunch(e.exec, user));
    g_launches += e.launches;
    return 0;
}

}  // namespace gvx
