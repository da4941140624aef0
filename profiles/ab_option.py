"""A/B of a gvx_debug_option inside ONE process (same box, same clocks): decoder train steps (forward + BPTT) at the bench batch
shape, the option alternated between 0 and 1.  Numbers from different gpurun boxes differ by several percent, so every kernel
variant is judged this way.    python profiles/ab_option.py <option> [T] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import genvox_b200  # noqa: E402
from bench import decoder_dims, synthetic_batch  # noqa: E402
from genvox_b200 import _native  # noqa: E402
from genvox_b200.training import decoder_train_step, make_optimizer  # noqa: E402

opt_name = sys.argv[1].encode()
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lib = _native.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = genvox_b200.Decoder(**decoder_dims()).to(dev).train()
dec.precision = "bf16"
opt = make_optimizer(dec)
memory, mel, gate, lengths = (t.to(dev) for t in synthetic_batch(torch, 64, 150, T))


def fwd_only():
    dec.set_dropout_seed(5)
    with torch.no_grad():
        dec(memory, mel, lengths)


def step():
    decoder_train_step(dec, opt, memory, mel, gate, lengths)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


res = {(v, k): [] for v in (0, 1) for k in ("fwd", "step")}
for r in range(reps):
    for v in (0, 1):
        lib.gvx_debug_option(opt_name, v)
        res[(v, "fwd")].append(timed(fwd_only))
        res[(v, "step")].append(timed(step))
lib.gvx_debug_option(opt_name, -1)
genvox_b200.check_device_errors()
for k in ("fwd", "step"):
    a, b = np.array(res[(0, k)]), np.array(res[(1, k)])
    print(f"{opt_name.decode()} T={T} {k}: off {np.median(a):.3f} ms (min {a.min():.3f})   on {np.median(b):.3f} ms (min {b.min():.3f})   "
          f"on/off {np.median(b) / np.median(a):.4f}")
