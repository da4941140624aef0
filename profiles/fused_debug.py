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
mem, mel, lens = synth.make_inputs(67, B, N, T, dims, ragged=len(sys.argv) > 4)
dec = make_decoder(dims, W, dev, True)
dec.precision = "bf16"
dbg = torch.zeros(4, 1024, 32, dtype=torch.int64).pin_memory()
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
x = dbg[3, : min(T, 1024)].numpy().astype(np.float64)
if T > 4:
    per = np.diff(x[2:, 7])
    print(f"step period (gbar2 arrive to gbar2 arrive): median {np.median(per):.0f} cyc = {np.median(per) / 1.965e3:.2f} us, "
          f"p10 {np.percentile(per, 10):.0f}, p90 {np.percentile(per, 90):.0f}, max {per.max():.0f}")
    names = ["ctx_part_go(TMA)", "mma_issued", "tmem_full", "gbar1_arrive", "gbar1_pass", "q_ready", "peer_ready", "gbar2_arrive"]
    for i, n in enumerate(names):
        print(f"   {n:18s} +{np.median(x[2:, i] - x[2:, 0]):8.0f} cyc")
    print(f"   loc_phase_done     +{np.median(x[2:, 12] - x[2:, 0]):8.0f} cyc   (started at the previous gbar2_arrive)")
    for i, n in ((13, "energies_done"), (14, "ctx_partial_done"), (15, "ctxp_sum_done"), (16, "next_hatt_first_mma"), (17, "next_hatt_mma_issued"), (18, "mma_done(MMA thr)"), (19, "loc_conv_done"), (20, "loc_tiles_done(w6)"), (21, "first_ctx_slab_mma")):
        print(f"   {n:18s} +{np.median(x[2:, i] - x[2:, 0]):8.0f} cyc")
    print("   next ctx_part_go   +%8.0f cyc" % np.median(x[3:, 0] - x[2:-1, 0]))

# ---- forward wall time on the device, fused vs per-step chain (no debug stamps)
lib.gvx_debug_timeline(None)
x_mem, x_mel, x_len = torch.from_numpy(mem).to(dev), torch.from_numpy(mel).to(dev), torch.from_numpy(lens).to(dev)
for fused in (1, 0, 1):
    lib.gvx_debug_option(b"fused", fused)
    ts = []
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        with torch.no_grad():
            dec(x_mem, x_mel, x_len)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"fused={fused}: forward {min(ts):.2f} ms (runs: {', '.join(f'{v:.2f}' for v in ts)})", flush=True)

# ---- does the debug buffer change the speed?  (stamps count SM cycles; events count wall time)
lib.gvx_debug_option(b"fused", 1)
for use_dbg in (1, 0, 1, 0):
    lib.gvx_debug_timeline(dbg.data_ptr() if use_dbg else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    with torch.no_grad():
        dec(x_mem, x_mel, x_len)
    e1.record()
    torch.cuda.synchronize()
    xx = dbg[3, :T].numpy().astype(np.float64)
    cyc = xx[T - 1, 7] - xx[2, 7]
    if use_dbg:
        g = dbg[3].numpy().astype(np.int64)
        t_entry, t_init, t_exit = g[0, 9], g[0, 10], g[0, 11]
        ends = g[:T, 8]
        print(f"   globaltimer: entry->init {(t_init - t_entry) / 1e3:.1f} us, init->step0 end {(ends[0] - t_init) / 1e3:.1f} us, "
              f"step0->step1 {(ends[1] - ends[0]) / 1e3:.1f} us, steps 1..T-1 {(ends[T - 1] - ends[1]) / 1e6:.3f} ms, "
              f"last step end->exit {(t_exit - ends[T - 1]) / 1e3:.1f} us, entry->exit {(t_exit - t_entry) / 1e6:.3f} ms; "
              f"slowest step {np.diff(ends).max() / 1e3:.1f} us at t={int(np.diff(ends).argmax()) + 1}")
    print(f"dbg={use_dbg}: forward {e0.elapsed_time(e1):.2f} ms; stamped chain cycles (last dbg run) {cyc:.0f} = {cyc / 1.965e6:.2f} ms @1.965 GHz", flush=True)
lib.gvx_debug_timeline(None)
