"""Per-step timeline of the persistent LSTM-chain kernels (CTA 0, clock64 stamps; gvx_debug_timeline).
    python profiles/chain_timeline.py [T]
Prints the median cycle offsets of each stamp relative to the step's first stamp, and ms for the whole chain."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genvox_b200 import _native                    # noqa: E402
from genvox_b200.decoder import _ptr, _stream      # noqa: E402

lib = _native.load()
dev = torch.device("cuda:0")
B, T, H = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 400, 1024
g = torch.Generator().manual_seed(3)
w_hh = ((torch.rand(4 * H, H, generator=g) * 2 - 1) / 32).to(dev)
pre = (torch.rand(T, B, 4 * H, generator=g) * 2 - 1).to(dev)
dh = (torch.rand(T, B, H, generator=g) * 2 - 1).to(dev)
h_out = torch.empty(T, B, H, device=dev)
c_out = torch.empty(T + 1, B, H, device=dev)
gates = torch.empty(T, B, 4 * H, device=dev)
dgates = torch.empty(T, B, 4 * H, device=dev)
dbg = torch.zeros(2, 1024, 32, dtype=torch.int64, device=dev)
for it in range(2):
    lib.gvx_debug_timeline(_ptr(dbg) if it == 1 else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _native.check(lib.gvx_test_lstm_chain(_ptr(w_hh), _ptr(pre), B, T, H, 0.1, 77, 1, _ptr(h_out), _ptr(c_out), _ptr(gates), _ptr(dh),
                                          _ptr(dgates), _stream()), "gvx_test_lstm_chain")
    e1.record()
    torch.cuda.synchronize()
    print(f"run {it}: fwd+bwd chain incl. setup {e0.elapsed_time(e1):.3f} ms for T={T}")
lib.gvx_debug_timeline(None)
d = dbg.cpu().numpy()
names = {0: ["gbar_ok", "tma_issued", "mma_issued", "tmem_full", "cell_done", "fenced", "arrived"],
         1: ["gbar_ok", "-", "mma_issued", "tmem_full", "cluster_sync", "cell_done", "arrived"]}
for k, label in ((0, "forward"), (1, "backward")):
    x = d[k, 1:min(T, 1024)].astype(np.float64)
    step = np.diff(x[:, 0])
    print(f"{label}: step period median {np.median(step):.0f} cyc ({np.median(step) / 1.965e3:.2f} us), p90 {np.percentile(step, 90):.0f}")
    for i, n in enumerate(names[k]):
        if n == "-":
            continue
        off = x[:, i] - x[:, 0]
        print(f"   {n:14s} +{np.median(off):8.0f} cyc  ({np.median(off) / 1.965e3:6.2f} us)")
    if k == 0:
        for i in list(range(7, 8)) + list(range(24, 28)) + list(range(8, 24)):
            off = x[:, i] - x[:, 0]
            print(f"   slot {i:2d}        +{np.median(off):8.0f} cyc")
