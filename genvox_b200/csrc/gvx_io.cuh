// genvox_b200 — small I/O descriptors shared by the bf16-mode kernels and the attention kernels:
// SrcSum reads a value that may still be split-K partial sums, BfDsts scatters a bf16 activation to the
// tcgen05 operand image(s) and/or row-major rows.
#pragma once
#include <cuda_bf16.h>

#include "gvx_common.cuh"

namespace gvx {

// value(b, col) = sum_s p[s * split_stride + b * ld + col]   (split-K partials or a plain matrix)
struct SrcSum {
    const float *p;
    int ld;
    int nsplit;
    long long split_stride;
};
// All partial loads are issued BEFORE the first add: with the rolled loop the in-order pipeline stalled at every add
// until its load had returned, i.e. one L2 round trip per split (measured: 6300 cycles for a 10-way sum at the head of
// the BPTT attention kernel).  The summation order is unchanged (ascending split index).
constexpr int SRC_MAXSPLIT = 16;       // tc_pick_ks never splits K more than 16 ways
__device__ __forceinline__ float src_get(const SrcSum &s, int b, int col) {
    const float *q = s.p + (size_t)b * s.ld + col;
    if (s.nsplit == 1) return q[0];
    float part[SRC_MAXSPLIT];
#pragma unroll
    for (int i = 0; i < SRC_MAXSPLIT; ++i) part[i] = i < s.nsplit ? q[(size_t)i * s.split_stride] : 0.f;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < SRC_MAXSPLIT; ++i)
        if (i < s.nsplit) v += part[i];
    for (int i = SRC_MAXSPLIT; i < s.nsplit; ++i) v += q[(size_t)i * s.split_stride];
    return v;
}
inline SrcSum src_plain(const float *p, int ld) { return SrcSum{p, ld, p ? 1 : 0, 0}; }
inline SrcSum src_split(const float *p, int ld, int nsplit, long long stride) { return SrcSum{p, ld, p ? nsplit : 0, stride}; }
inline SrcSum src_none() { return SrcSum{nullptr, 0, 0, 0}; }

// destination of a bf16 activation vector: engine image or row-major rows
// element offset of (row b, column kk) inside an engine operand image: [K/64 slabs][npad rows][64] bf16, SWIZZLE_128B
// (the 16-byte chunk c of row b sits at chunk position c ^ (b & 7) of the row's 128 bytes; gvx_tc.cuh)
__device__ __forceinline__ size_t bf_img_off(int kk, int b, int npad) {
    return (size_t)(kk >> 6) * ((size_t)npad * 64) + (size_t)b * 64 + (size_t)((((kk >> 3) & 7) ^ (b & 7)) << 3) + (kk & 7);
}

struct BfDst {
    __nv_bfloat16 *p;
    int kind;    // 0 none, 1 engine image (see bf_img_off) starting at element column `koff`, 2 row-major (ld) starting at column `koff`
    int koff;    // multiple of 8
    int ld;      // image: NPAD; row-major: row stride in elements
    long long tstride;   // elements between consecutive frames (bf_store1_t only)
};
constexpr int BF_MAXDST = 5;
struct BfDsts {
    BfDst d[BF_MAXDST];
    int n;
};
inline void add_img(BfDsts &s, __nv_bfloat16 *p, int koff, int npad) {
    if (p) s.d[s.n++] = BfDst{p, 1, koff, npad, 0};
}
inline void add_rm(BfDsts &s, __nv_bfloat16 *p, int koff, int ld) {
    if (p) s.d[s.n++] = BfDst{p, 2, koff, ld, 0};
}
// 8 consecutive columns k..k+7 (k multiple of 8) of row b
__device__ __forceinline__ void bf_store8(const BfDsts &s, int b, int k, uint4 v) {
    for (int i = 0; i < s.n; ++i) {
        const BfDst &d = s.d[i];
        const int kk = d.koff + k;
        if (d.kind == 1) *reinterpret_cast<uint4 *>(d.p + bf_img_off(kk, b, d.ld)) = v;
        else *reinterpret_cast<uint4 *>(d.p + (size_t)b * d.ld + kk) = v;
    }
}
// 4 consecutive columns k..k+3 (k multiple of 4) of row b
__device__ __forceinline__ void bf_store4(const BfDsts &s, int b, int k, uint2 v) {
    for (int i = 0; i < s.n; ++i) {
        const BfDst &d = s.d[i];
        const int kk = d.koff + k;
        if (d.kind == 1) *reinterpret_cast<uint2 *>(d.p + bf_img_off(kk, b, d.ld)) = v;
        else *reinterpret_cast<uint2 *>(d.p + (size_t)b * d.ld + kk) = v;
    }
}
// one column k of row b
__device__ __forceinline__ void bf_store1(const BfDsts &s, int b, int k, float x) {
    const __nv_bfloat16 h = __float2bfloat16(x);
    for (int i = 0; i < s.n; ++i) {
        const BfDst &d = s.d[i];
        const int kk = d.koff + k;
        if (d.kind == 1) d.p[bf_img_off(kk, b, d.ld)] = h;
        else d.p[(size_t)b * d.ld + kk] = h;
    }
}
// one column k of row b of frame t (time-batched producers: prenet over all frames)
__device__ __forceinline__ void bf_store1_t(const BfDsts &s, int t, int b, int k, float x) {
    const __nv_bfloat16 h = __float2bfloat16(x);
    for (int i = 0; i < s.n; ++i) {
        const BfDst &d = s.d[i];
        const int kk = d.koff + k;
        __nv_bfloat16 *base = d.p + (size_t)t * d.tstride;
        if (d.kind == 1) base[bf_img_off(kk, b, d.ld)] = h;
        else base[(size_t)b * d.ld + kk] = h;
    }
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}

}  // namespace gvx
