"""One train step at the other BASELINE.json training shapes (functional check + timing; the bench line is configs[2]):
configs[0] shape in fp32 mode (B=16, N=120, T=600) and configs[4] in bf16 mode (B=32, N=300, T=1600: N > 160 takes the
per-step attention chain, the decoder-LSTM chains stay persistent)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genvox_b200                                                       # noqa: E402
from bench import decoder_dims, synthetic_batch                          # noqa: E402
from genvox_b200.training import decoder_train_step, make_optimizer      # noqa: E402

dev = torch.device("cuda:0")
for name, (B, N, T), prec in (("configs[0] shape, fp32 mode", (16, 120, 600), "fp32"), ("configs[4], bf16 mode", (32, 300, 1600), "bf16")):
    torch.manual_seed(0)
    dec = genvox_b200.Decoder(**decoder_dims()).to(dev).train()
    dec.precision = prec
    opt = make_optimizer(dec)
    memory, mel, gate, lengths = (t.to(dev) for t in synthetic_batch(torch, B, N, T))
    for _ in range(2):
        loss, gn = decoder_train_step(dec, opt, memory, mel, gate, lengths)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        loss, gn = decoder_train_step(dec, opt, memory, mel, gate, lengths)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"{name}: B={B} N={N} T={T}: {ms:.1f} ms per train step = {B * T / ms * 1e3:,.0f} mel-frames/s, loss {float(loss):.4f}, "
          f"grad norm {float(gn):.3f}, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del dec, opt, memory, mel, gate
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
