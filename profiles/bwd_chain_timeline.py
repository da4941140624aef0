"""Phase timeline of the persistent attention-chain BPTT kernel (gvx_fused_bwd.cuh): clock64 stamps of CTA 0, one row per
reverse-time iteration.      python profiles/bwd_chain_timeline.py [B N T]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import genvox_b200                                 # noqa: E402
from genvox_b200 import _native                    # noqa: E402
from oracle import synth                           # noqa: E402
from test_cuda_parity import make_decoder          # noqa: E402

B, N, T = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 150, 48)
lib = _native.load()
dev = torch.device("cuda:0")
dims = synth.DecoderDims()
W = synth.make_decoder_weights(23, dims)
mem, mel, lens = synth.make_inputs(67, B, N, T, dims, ragged=False)
dec = make_decoder(dims, W, dev, True)
dec.precision = "bf16"
dbg = torch.zeros(4, 1024, 32, dtype=torch.int64, device=dev)
x_mem, x_mel, x_len = torch.from_numpy(mem).to(dev), torch.from_numpy(mel).to(dev), torch.from_numpy(lens).to(dev)


def step():
    dec.zero_grad(set_to_none=True)
    m, g, a = dec(x_mem.clone().requires_grad_(True), x_mel, x_len)
    (m.sum() + g.sum()).backward()
    torch.cuda.synchronize()
    genvox_b200.check_device_errors()


step()
lib.gvx_debug_timeline(dbg.data_ptr())
step()
lib.gvx_debug_timeline(None)
x = dbg[1, :T].cpu().numpy().astype(np.float64)
names = ["iter start", "d ctx assembled (poll, carry dots, tanh tile)", "d w / d e / d q partial published", "d conv mma + halo push done",
         "d q rows arrived, d h_q", "d gates written, barrier arrive", "tmem_full (TMA + MMA; conv transpose overlapped)",
         "partials pushed + summed, d ctx published"]
per = np.diff(x[2:-1, 0])
print(f"B={B} N={N} T={T}: iteration period median {np.median(per):.0f} cyc = {np.median(per) / 1.965e3:.2f} us (p10 {np.percentile(per, 10):.0f}, p90 {np.percentile(per, 90):.0f})")
for k, n in ((12, "  (d w partials on mma, sync)"), (13, "  (d e written, sync)"), (14, "  (cell inputs requested)"), (15, "  (warp 0: its d conv tile done)"), (8, "  (TMA: image barrier passed)"),
             (10, "  (MMA: first slab landed)"), (9, "  (TMA: last slab issued)"), (11, "  (MMA: last slab issued)")):
    print(f"    {n:46s} +{np.median(x[2:-1, k] - x[2:-1, 0]):7.0f} cyc")
prev = 0.0
for i, n in enumerate(names):
    d = np.median(x[2:-1, i] - x[2:-1, 0])
    print(f"  {i} {n:48s} +{d:7.0f} cyc  ({(d - prev) / 1.965e3:5.2f} us)")
    prev = d

# ---- per-CTA skew at reverse-time iteration 20 (%globaltimer, ns): when did each of the 128 CTAs reach each event?
g = dbg[2].reshape(-1)[: 128 * 16].cpu().numpy().astype(np.float64).reshape(128, 16)
ev = {0: "iter start", 1: "d ctx assembled", 2: "d q published", 3: "d conv done", 4: "d q rows arrived", 5: "image barrier arrive",
      8: "image barrier passed", 6: "tmem_full", 7: "d ctx published"}
t0 = g[:, 0].min()
print("per-CTA event times at iteration 20, ns after the first CTA started it: min / median / max (argmax CTA)")
for k in (0, 1, 2, 3, 4, 5, 8, 6, 7):
    col = g[:, k] - t0
    print(f"  {ev[k]:24s} {col.min():8.0f} {np.median(col):8.0f} {col.max():8.0f}   (CTA {int(col.argmax())}; CTA 0: {col[0]:.0f})")
