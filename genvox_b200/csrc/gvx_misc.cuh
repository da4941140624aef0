// genvox_b200 — layout kernels: weight repack, frame pack / output unpack, gate-stop bookkeeping.
#pragma once
#include "gvx_common.cuh"

namespace gvx {

// nn.LSTMCell weights (tacotron2.py:286,:294; gate row order i,f,g,o) -> unit-major concatenated
// Wc[4u+g][k] = k < Kih ? w_ih[g*HID+u][k] : w_hh[g*HID+u][k-Kih];  WT = Wc^T;  bc = b_ih + b_hh
__global__ void k_pack_lstm(const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                            const float *__restrict__ b_ih, const float *__restrict__ b_hh, int HID, int Kih,
                            float *__restrict__ Wc, float *__restrict__ bc, float *__restrict__ WT) {
    const int K = Kih + HID, R = 4 * HID;
    const size_t total = (size_t)R * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int rp = (int)(i / K), k = (int)(i - (size_t)rp * K);
        const int u = rp >> 2, g = rp & 3, src = g * HID + u;
        const float v = k < Kih ? w_ih[(size_t)src * Kih + k] : w_hh[(size_t)src * HID + (k - Kih)];
        Wc[i] = v;
        WT[(size_t)k * R + rp] = v;
        if (k == 0) bc[rp] = b_ih[src] + b_hh[src];
    }
}

// inverse of k_pack_lstm for gradients: dWc [4HID, K] (packed) -> d w_ih, d w_hh (torch layout)
__global__ void k_unpack_lstm_grad(const float *__restrict__ dWc, int HID, int Kih, float *__restrict__ d_ih,
                                   float *__restrict__ d_hh) {
    const int K = Kih + HID, R = 4 * HID;
    const size_t total = (size_t)R * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int rp = (int)(i / K), k = (int)(i - (size_t)rp * K);
        const int u = rp >> 2, g = rp & 3, dst = g * HID + u;
        if (k < Kih) d_ih[(size_t)dst * Kih + k] = dWc[i];
        else d_hh[(size_t)dst * HID + (k - Kih)] = dWc[i];
    }
}

// out[c*R + r] = in[r*C + c]  (small matrices)
__global__ void k_transpose(const float *__restrict__ in, int R, int C, float *__restrict__ out) {
    const int total = R * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / C, c = i - r * C;
        out[(size_t)c * R + r] = in[i];
    }
}

// [mel rows | gate row] and biases concatenated (linear_projection + gate_layer, tacotron2.py:361-362)
__global__ void k_pack_proj(const float *__restrict__ proj_w, const float *__restrict__ proj_b,
                            const float *__restrict__ gate_w, const float *__restrict__ gate_b, int M, int Kp,
                            float *__restrict__ Wpg, float *__restrict__ bpg) {
    const int total = (M + 1) * Kp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / Kp, k = i - r * Kp;
        Wpg[i] = r < M ? proj_w[i] : gate_w[k];
        if (k == 0) bpg[r] = r < M ? proj_b[r] : gate_b[0];
    }
}

// parse_decoder_inputs + go frame (tacotron2.py:317-320,:370-372): FR[t][b][m] = t ? mel_in[b][m][t-1] : 0
__global__ void k_pack_frames(const float *__restrict__ mel_in, int B, int M, int T, float *__restrict__ FR) {
    const size_t total = (size_t)T * B * M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M);
        const size_t tb = i / M;
        const int b = (int)(tb % B), t = (int)(tb / B);
        FR[i] = t == 0 ? 0.f : mel_in[((size_t)b * M + m) * T + (t - 1)];
    }
}

// parse_decoder_outputs (tacotron2.py:322-331): OUT[t][b][OL] -> mel[b][m][t] (row stride Tal), gate[b][t]
__global__ void k_unpack_out(const float *__restrict__ OUT, int B, int M, int OL, int steps, int Tal,
                             float *__restrict__ mel, float *__restrict__ gate) {
    const size_t total = (size_t)B * (M + 1) * steps;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % steps);
        const size_t bm = i / steps;
        const int m = (int)(bm % (M + 1)), b = (int)(bm / (M + 1));
        const float v = OUT[((size_t)t * B + b) * OL + m];
        if (m < M) mel[((size_t)b * M + m) * Tal + t] = v;
        else gate[(size_t)b * Tal + t] = v;
    }
}

// inverse of k_unpack_out for the upstream gradients: DOUT[t][b][OL] <- d_mel[b][m][t], d_gate[b][t]
__global__ void k_pack_dout(const float *__restrict__ d_mel, const float *__restrict__ d_gate, int B, int M, int OL,
                            int T, float *__restrict__ DOUT) {
    const size_t total = (size_t)T * B * OL;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % OL);
        const size_t tb = i / OL;
        const int b = (int)(tb % B), t = (int)(tb / B);
        float v = 0.f;
        if (m < M) v = d_mel[((size_t)b * M + m) * T + t];
        else if (m == M) v = d_gate[(size_t)b * T + t];
        DOUT[i] = v;
    }
}

// stop test of Decoder.inference (tacotron2.py:405): sigmoid(gate) > threshold, strict, frame included.
// flags[0] = number of rows still running after this step.  One block.
__global__ void k_gate_check(const float *__restrict__ out_t, int B, int M, int OL, float thr, int t,
                             int32_t *__restrict__ n_frames, int *__restrict__ flags) {
    pdl_trigger();
    pdl_wait();
    __shared__ int running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int nf = n_frames[b];
        if (nf < 0) {
            if (sigmoidf_(out_t[(size_t)b * OL + M]) > thr) { nf = t + 1; n_frames[b] = nf; }
        }
        if (nf < 0) atomicAdd(&running, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) flags[0] = running;
}

__global__ void k_set_seed(uint32_t *p, uint32_t lo, uint32_t hi) {
    p[0] = lo;
    p[1] = hi;
}

__global__ void k_fill_i32(int32_t *p, int n, int32_t v, int only_negative) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!only_negative || p[i] < 0) p[i] = v;
}

__global__ void k_fill_f32(float *p, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

inline int grid_for(size_t n, int threads = 256, int cap = 148 * 16) {
    size_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > (size_t)cap) g = cap;
    return (int)g;
}

}  // namespace gvx
