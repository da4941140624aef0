"""Short driver for ncu: decoder train steps (fwd + BPTT) at the bench batch shape with few frames,
then a few inference steps.  Usage: python profiles/run_step.py [T] [infer_steps] [bf16|fp32]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import genvox_b200  # noqa: E402
from bench import decoder_dims, synthetic_batch  # noqa: E402
from genvox_b200.training import decoder_train_step, make_optimizer  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
PREC = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = genvox_b200.Decoder(**decoder_dims()).to(dev).train()
dec.precision = PREC
opt = make_optimizer(dec)
memory, mel, gate, lengths = (t.to(dev) for t in synthetic_batch(torch, 64, 150, T))
for _ in range(2):
    loss, _ = decoder_train_step(dec, opt, memory, mel, gate, lengths)
torch.cuda.synchronize()
if S > 0:
    dec.eval()
    dec.inference(memory, ignore_gate=True, max_decoder_steps=S)
    torch.cuda.synchronize()
print("ok", float(loss))
