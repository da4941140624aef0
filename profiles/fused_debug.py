"""Post-mortem harness for the fused attention-chain kernel: runs one bf16 forward with the fused path, with the debug
buffer in pinned host memory so that a watchdog thread can print the per-CTA progress markers even if the launch hangs.
    python profiles/fused_debug.py [B N T]"""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from genvox_b200 import _native                    # noqa: E402
from oracle import synth                           # noqa: E402
from test_cuda_parity import make_decoder          # noqa: E402

B, N, T = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 150, 6)
lib = _native.load()
dev = torch.device("cuda:0")
dims = synth.DecoderDims()
W = synth.make_decoder_weights(23, dims)
mem, mel, lens = synth.make_inputs(67, B, N, T, dims, ragged=True)
dec = make_decoder(dims, W, dev, True)
dec.precision = "bf16"
dbg = torch.zeros(3, 1024, 32, dtype=torch.int64).pin_memory()
lib.gvx_debug_timeline(dbg.data_ptr())
lib.gvx_debug_option(b"fused", 1)
done = threading.Event()


def dump(tag):
    prog = dbg[2].view(torch.int32).reshape(-1)[: 3 * 128].numpy().reshape(3, 128)
    print(tag, "progress markers (workers 8t+phase / TMA 4t+2part+1 / MMA t+1):", flush=True)
    for r, name in enumerate(("workers", "tma", "mma")):
        print(f"  {name:8s}", " ".join(str(int(v)) for v in prog[r]), flush=True)


def watchdog():
    if not done.wait(25.0):
        dump("HUNG:")
        os._exit(3)


threading.Thread(target=watchdog, daemon=True).start()
try:
    dec.set_dropout_seed(5)
    with torch.no_grad():
        m, g, a = dec(torch.from_numpy(mem).to(dev), torch.from_numpy(mel).to(dev), torch.from_numpy(lens).to(dev))
    torch.cuda.synchronize()
    print("forward ok: finite =", bool(torch.isfinite(m).all()), "align row sums max dev =", float((a.sum(-1) - 1).abs().max()))
except Exception as exc:  # noqa: BLE001
    print("forward failed:", exc)
done.set()
dump("END:")
x = dbg[0, : min(T, 1024)].numpy().astype(np.float64)
if T > 2:
    per = np.diff(x[1:, 7])
    print(f"step period (gbar2 arrive to gbar2 arrive): median {np.median(per):.0f} cyc = {np.median(per) / 1.965e3:.2f} us")
    names = ["ctx_part_go(TMA)", "mma_issued", "tmem_full", "gbar1_arrive", "gbar1_pass", "q_ready", "peer_ready", "gbar2_arrive"]
    base = x[1:-1, 7]          # end of the previous step
    for i, n in enumerate(names):
        print(f"   {n:18s} +{np.median(x[2:, i] - base):8.0f} cyc")
