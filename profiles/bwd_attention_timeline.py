"""Phase timeline of the per-step BPTT attention kernel (k_attention_bwd_c2, CTA 0, clock64 stamps via gvx_debug_timeline).
    python profiles/bwd_attention_timeline.py [T]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from genvox_b200 import _native                    # noqa: E402
from oracle import synth                           # noqa: E402
from test_cuda_parity import make_decoder          # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B, N = 64, 150
lib = _native.load()
dev = torch.device("cuda:0")
dims = synth.DecoderDims()
W = synth.make_decoder_weights(23, dims)
mem, mel, lens = synth.make_inputs(67, B, N, T, dims, ragged=False)
dec = make_decoder(dims, W, dev, True)
dec.precision = "bf16"
dbg = torch.zeros(4, 1024, 32, dtype=torch.int64, device=dev)
for it in range(2):
    lib.gvx_debug_timeline(dbg.data_ptr() if it == 1 else None)
    dec.zero_grad(set_to_none=True)
    memory = torch.from_numpy(mem).to(dev).requires_grad_(True)
    m, g, a = dec(memory, torch.from_numpy(mel).to(dev), torch.from_numpy(lens).to(dev))
    (m.square().mean() + g.square().mean()).backward()
    torch.cuda.synchronize()
lib.gvx_debug_timeline(None)
x = dbg[1, 2:T - 2, :8].cpu().numpy().astype(np.float64)
names = ["entry", "pdl_wait passed", "inputs staged", "d w done", "softmax bwd done", "d s / d q / d conv done (cluster)",
         "halo exchanged", "conv^T + d h_q + carries done"]
print(f"k_attention_bwd_c2, B={B} N={N}: median cycles from kernel entry (CTA 0), {x.shape[0]} steps")
for i, n in enumerate(names):
    print(f"   {n:40s} +{np.median(x[:, i] - x[:, 0]):8.0f} cyc  ({np.median(x[:, i] - x[:, 0]) / 1.965e3:6.2f} us)")
step = np.diff(dbg[1, 2:T - 2, 0].cpu().numpy().astype(np.float64))
print(f"   entry-to-entry period (whole BPTT step, frames descend) median {np.median(np.abs(step)):8.0f} cyc ({np.median(np.abs(step)) / 1.965e3:.2f} us)")
