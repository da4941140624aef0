"""Device time of the tcgen05 NT GEMM (csrc/gvx_nt_gemm.cuh) on the shapes of the train step, next to torch.matmul (cuBLAS) on the
same shapes.      python profiles/nt_gemm_bench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genvox_b200 import _native  # noqa: E402

lib = _native.load()
TB = 51200
shapes = [("dW dec-LSTM      (4096 x 2560, K = frames)", 4096, 2560, TB), ("dW att-LSTM      (4096 x 1792, K = frames)", 4096, 1792, TB),
          ("dW query         (128 x 1024, K = frames)", 128, 1024, TB), ("dW projections   (81 x 1536, K = frames)", 88, 1536, TB),
          ("dec gates input  (frames x 4096, K = 1536)", TB, 4096, 1536), ("att gates prenet (frames x 4096, K = 256)", TB, 4096, 256),
          ("d [h_att|ctx]    (frames x 1536, K = 4096)", TB, 1536, 4096), ("d prenet_out     (frames x 256, K = 4096)", TB, 256, 4096),
          ("projections      (frames x 81, K = 1536)", TB, 88, 1536)]
dev = torch.device("cuda:0")
for name, M, N, K in shapes:
    ms = C.c_float(0)
    _native.check(lib.gvx_bench_nt_gemm(M, N, K, 5, C.byref(ms)), "gvx_bench_nt_gemm")
    a = torch.zeros(M, K, dtype=torch.bfloat16, device=dev)
    b = torch.zeros(N, K, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        c = a @ b.t()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c = a @ b.t()
    e1.record()
    torch.cuda.synchronize()
    tms = e0.elapsed_time(e1) / 5
    fl = 2.0 * M * N * K
    print(f"{name:46s} own {ms.value:7.3f} ms = {fl / ms.value / 1e9:7.1f} TFLOP/s   cuBLAS(bf16 out) {tms:7.3f} ms = {fl / tms / 1e9:7.1f} TFLOP/s")
    del a, b, c
