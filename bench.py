#!/usr/bin/env python
"""bench.py — throughput of the Tacotron2 decoder recurrence (the hot path) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[2] — data-parallel TRAINING, batch 64 per GPU,
150 tokens, 800x80 mel frames — one "step" is Tacotron2.train_step (tacotron2.py:515-522) restricted to
the decoder: zero_grad, Decoder.forward, loss, BPTT, [gradient all-reduce], clip_grad_norm_, Adam.
`value` = mel-frames/s with inputs resident in HBM, CUDA-event timed, max over ranks;
`e2e`   = the same step through genvox_b200.Decoder with HOST (pinned) inputs: H2D of memory/mel/gate
          and a D2H read of the loss inside the timed region;
`infer` = BASELINE.json configs[1] (batch 64, 1000 fixed decoder steps, fp32) in decoder steps/s;
`roofline` for the dominant kernel of the step, from CUDA events the library records around every
phase launch during one extra profiled step (gvx_profile_*), `cpu_baseline` = the oracle port
(torch CPU, all host threads) on a bounded sample of the same workload.
--impl reference times that CPU implementation alone (rank 0 only under torchrun).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRAIN = dict(B=64, N=150, T=int(os.environ.get("GVX_BENCH_T", "800")))     # BASELINE.json configs[2] (per GPU); the env
                                                                           # override is for debugging runs only
INFER = dict(B=64, N=150, steps=1000)     # BASELINE.json configs[1]
CPU_SAMPLE_T = 8                          # frames of the training workload the CPU legs run per step
METRIC = "Tacotron2 train mel-frames/s"
UNIT = "mel-frames/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def decoder_dims():
    return dict(n_mels=80, encoder_embedding_dim=512, decoder_rnn_dim=1024, prenet_dim=256, max_decoder_steps=1000,
                gate_threshold=0.5, p_attention_dropout=0.1, p_decoder_dropout=0.1, attention_rnn_dim=1024,
                attention_dim=128, attention_location_n_filters=32, attention_location_kernel_size=31)


def synthetic_batch(torch, B, N, T, rank=0):
    """Synthetic decoder inputs of the named shape: encoder outputs, teacher-forcing mels
    (input and target), gate target 1 at the last frame (models/tts/__init__.py:53), full lengths."""
    g = torch.Generator().manual_seed(1 + rank)
    memory = 0.5 * torch.randn(B, N, 512, generator=g)
    mel = torch.randn(B, 80, T, generator=g)
    gate = torch.zeros(B, T)
    gate[:, -1] = 1.0
    lengths = torch.full((B,), N, dtype=torch.int64)
    return memory, mel, gate, lengths


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU legs (oracle port)
def cpu_train_leg(steps, warmup, sample_T=CPU_SAMPLE_T):
    """The oracle port of the same train step on the host cores, on a bounded sample of the workload:
    the full batch (64 x 150 tokens) but `sample_T` teacher-forced frames per step."""
    import torch
    from oracle import decoder_oracle as O
    import genvox_b200
    from genvox_b200.training import decoder_loss
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, N = TRAIN["B"], TRAIN["N"]
    torch.manual_seed(0)
    dec = genvox_b200.Decoder(**decoder_dims())          # parameter container only: same init as the GPU arm
    P = {k: v.detach().clone().requires_grad_(True) for k, v in dec.named_parameters()}
    opt = torch.optim.Adam(list(P.values()), lr=1e-3, weight_decay=1e-6)
    memory, mel, gate, lengths = synthetic_batch(torch, B, N, sample_T)

    def step(i):
        opt.zero_grad(set_to_none=True)
        m, g, _ = O.forward_teacher(P, memory, mel, lengths, seed=123 + i, training=True)
        loss, _, _ = decoder_loss(m, g, mel, gate)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)
        opt.step()
        return float(loss.detach())

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=B * sample_T / dt, unit=UNIT, cores=cores, threads=torch.get_num_threads(), kind="port",
                sample=f"oracle/decoder_oracle.py train step (fwd+loss+BPTT+clip+Adam), B={B}, N={N}, {sample_T} of "
                       f"{TRAIN['T']} frames per step, {steps} timed steps after {warmup} warm-up",
                ms_per_step=dt * 1e3)


def torch_gpu_train_leg(steps, warmup, sample_T=100):
    """Optional (--impl reference --reference-device cuda; never part of the default run): the oracle port's torch ops
    executed on the GPU - the stock PyTorch eager path the reference itself takes on a CUDA device (nn.LSTMCell math,
    conv1d, bmm, autograd BPTT), with torch's own dropout instead of the Philox stream (mask values do not matter for
    timing) - on a bounded sample of the workload: the full batch, `sample_T` teacher-forced frames per step."""
    import torch
    import torch.nn.functional as F
    from oracle import decoder_oracle as O
    import genvox_b200
    from genvox_b200.training import decoder_loss
    dev = torch.device("cuda:0")
    B, N = TRAIN["B"], TRAIN["N"]
    torch.manual_seed(0)
    dec = genvox_b200.Decoder(**decoder_dims())          # parameter container only: same init as the other arms
    P = {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in dec.named_parameters()}
    opt = torch.optim.Adam(list(P.values()), lr=1e-3, weight_decay=1e-6)
    memory, mel, gate, lengths = synthetic_batch(torch, B, N, sample_T)
    memory, mel, gate = memory.to(dev), mel.to(dev), gate.to(dev)
    O.philox_dropout = lambda x, p, on, *a, **k: F.dropout(x, p, on)          # tacotron2.py:143,:341,:358
    O.get_mask_from_lengths = lambda lengths, max_len=None: (
        torch.arange(max_len if max_len is not None else int(lengths.max()), device=dev)[None, :] >= lengths.to(dev)[:, None])

    def step(i):
        opt.zero_grad(set_to_none=True)
        m, g, _ = O.forward_teacher(P, memory, mel, lengths, seed=123 + i, training=True)
        loss, _, _ = decoder_loss(m, g, mel, gate)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)
        opt.step()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=B * sample_T / dt, unit=UNIT, cores=os.cpu_count() or 1, kind="port",
                sample=f"oracle/decoder_oracle.py train step on cuda:0 (stock PyTorch eager ops, fp32), B={B}, N={N}, {sample_T} of "
                       f"{TRAIN['T']} frames per step, {steps} timed steps after {warmup} warm-up",
                ms_per_step=dt * 1e3)


def run_reference(args, rank):
    if rank != 0:
        return
    if args.reference_device == "cuda":
        leg = torch_gpu_train_leg(args.steps, args.warmup)
        print(json.dumps({"impl": "reference", "device": "cuda", "metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": 1,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
                          "dtype": "f32", "data": "synthetic", "config": {"workload": leg["sample"]}}), flush=True)
        return
    leg = cpu_train_leg(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"decoder train step, batch {TRAIN['B']}, {TRAIN['N']} tokens, {CPU_SAMPLE_T} of "
                                   f"{TRAIN['T']}x80 mel frames per step (bounded sample), CPU"},
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def _stage(rank, msg):
    if os.environ.get("GVX_BENCH_TRACE"):
        print(f"[rank {rank}] {time.strftime('%H:%M:%S')} {msg}", file=sys.stderr, flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import genvox_b200
    from genvox_b200 import _native
    from genvox_b200.training import decoder_train_step, make_optimizer

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU leg)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _native.load()
    B, N, T = TRAIN["B"], TRAIN["N"], TRAIN["T"]
    K, Wm = args.steps, max(args.warmup, 3)

    torch.manual_seed(0)                                   # identical init on every rank
    dec = genvox_b200.Decoder(**decoder_dims()).to(dev).train()
    dec.precision = args.precision
    dec.dropout_row_offset = rank * B                      # ranks draw the rows of one global batch
    opt = make_optimizer(dec)
    memory_h, mel_h, gate_h, lengths_h = (t.pin_memory() for t in synthetic_batch(torch, B, N, T, rank))
    memory, mel, gate, lengths = (t.to(dev) for t in (memory_h, mel_h, gate_h, lengths_h))
    group = dist.group.WORLD if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident(sync_gradients=True):
        return decoder_train_step(dec, opt, memory, mel, gate, lengths, group=group, sync_gradients=sync_gradients)

    loss_h = torch.zeros(1).pin_memory()

    def step_e2e():
        m = memory_h.to(dev, non_blocking=True)
        x = mel_h.to(dev, non_blocking=True)
        g = gate_h.to(dev, non_blocking=True)
        le = lengths_h.to(dev, non_blocking=True)
        loss, _ = decoder_train_step(dec, opt, m, x, g, le, group=group)
        loss_h.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_h[0])

    # ---- value: inputs resident in HBM
    _stage(rank, "setup done")
    for i in range(Wm):
        step_resident()
        if os.environ.get("GVX_BENCH_SYNC"):
            torch.cuda.synchronize()
        _stage(rank, f"warm-up step {i} enqueued")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    launches0 = lib.gvx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_t0 = time.perf_counter()
    for _ in range(K):
        loss, _ = step_resident()
    host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / K      # host time to ENQUEUE a step (no sync inside)
    e1.record()
    barrier()
    launches = lib.gvx_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    _stage(rank, "timed region done")
    ms = max_over_ranks(e0.elapsed_time(e1)) / K
    value = world * B * T / (ms * 1e-3)
    final_loss = float(loss)

    # ---- e2e: host buffers, H2D + D2H inside the timed region
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
    _stage(rank, "e2e done")
    h2d = sum(t.numel() * t.element_size() for t in (memory_h, mel_h, gate_h, lengths_h))
    e2e = {"value": world * B * T / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: decoder train step (fwd + loss + BPTT + allreduce + clip + Adam), "
                                   f"batch {B}/GPU, {N} tokens, {T}x80 mel frames, "
                                   + ("bf16 tcgen05 gate GEMMs with fp32 accumulation, fp32 pointwise/attention"
                                      if args.precision == "bf16" else "fp32 arithmetic"),
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2": "per-step working set (stash + workspace, several GB) is far larger than the 126 MB L2"},
            "e2e": e2e, "gpu_launches": int(launches), "final_loss": final_loss, "host_enqueue_ms_per_step": host_enqueue_ms}
    if clocks is not None:
        line["clocks"] = clocks
    gs = (C.c_ulonglong * 4)()
    lib.gvx_graph_stats(gs)
    line["cuda_graphs"] = {"eager_calls": int(gs[0]), "captured": int(gs[1]), "replays": int(gs[2]), "failed_captures": int(gs[3])}

    if rank == 0:
        # ---- one extra, profiled step: per-phase device time from CUDA events around every launch
        lib.gvx_profile_reset()
        lib.gvx_profile_enable(1)
        step_resident(sync_gradients=False)      # rank 0 alone: no collective in this extra step
        torch.cuda.synchronize()
        lib.gvx_profile_enable(0)
        phases, slot = {}, 0
        while True:
            name = lib.gvx_profile_slot_name(slot)
            if name is None:
                break
            tot, cnt = C.c_double(0), C.c_longlong(0)
            lib.gvx_profile_read(slot, C.byref(tot), C.byref(cnt))
            if cnt.value:
                phases[name.decode()] = {"ms": tot.value, "launches": cnt.value, "avg_us": 1e3 * tot.value / cnt.value}
            slot += 1
        line["phases_ms"] = {k: round(v["ms"], 3) for k, v in phases.items()}
        pk = peaks()
        H, Kd, Ka, E, D = 1024, 2560, 1792, 512, 128
        roofs = {}
        for name, Kc in (("dec_lstm", Kd), ("att_lstm", Ka), ("bwd_dec_gemm", Kd), ("bwd_att_gemm", Ka)):
            if name in phases:
                flops = 2.0 * B * Kc * 4 * H * T                # algorithmic FLOPs of the gate GEMMs of all T steps
                ach = flops / (phases[name]["ms"] * 1e-3) / 1e12     # (one launch per step, or one persistent launch)
                roofs[name] = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                               "frac": ach / pk["tf_sustained"], "traffic": None, "avg_us": phases[name]["avg_us"],
                               "us_per_step": 1e3 * phases[name]["ms"] / T, "launches": phases[name]["launches"],
                               "peak_source": pk["source"] + " (bf16 sustained)"
                               + ("" if args.precision == "bf16" else "; this kernel runs fp32 FFMA")}
        for name in ("attention", "bwd_attention"):
            if name in phases:
                fused = phases[name]["launches"] == 1          # the persistent attention chain: ONE launch for all T steps
                if name == "attention" and fused:
                    # SURVEY.md 8(d) attention path per step: memory (bf16 copy) + processed memory (fp32) read, alignment written
                    nbytes = B * N * (E * 2.0 + D * 4.0) + B * N * 4.0
                elif name == "bwd_attention" and phases.get("attention", {}).get("launches") == 1:
                    # after the fused forward chain the BPTT attention kernel reads the bf16 memory copy + the fp32 tanh stash
                    nbytes = B * N * (E * 2.0 + D * 4.0) + B * N * 4.0
                else:
                    nbytes = B * N * (E + D) * 4.0 + B * N * 4.0    # memory + processed memory (or stashed tanh) + weights row
                per_step_us = 1e3 * phases[name]["ms"] / T if fused else phases[name]["avg_us"]
                ach = nbytes / (per_step_us * 1e-6) / 1e9
                roofs[name] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": ach / pk["hbm_gbs"], "traffic": None, "avg_us": phases[name]["avg_us"],
                               "us_per_step": per_step_us, "launches": phases[name]["launches"],
                               "algorithmic_bytes_per_step": nbytes, "peak_source": pk["source"]}
                if name == "attention" and fused:
                    roofs[name]["note"] = ("one persistent launch = attention LSTM + query + attention of all T steps; "
                                           "latency-bound (2 grid barriers + 2 exchanges per step), operands L2 resident")
        tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.isfile(tpath):          # DRAM bytes per launch from the committed ncu --set full capture (cold cache)
            with open(tpath) as fh:
                traffic = json.load(fh)
            for k, r in roofs.items():
                if k in traffic:
                    r["traffic"] = traffic[k]["dram_bytes_per_launch"]
                    r["traffic_note"] = f'ncu {traffic[k]["kernel"]}, {traffic[k]["launch"]}, cold cache'
        if roofs:
            dominant = max(roofs, key=lambda k: phases[k]["ms"])
            line["roofline"] = dict(roofs[dominant], kernel=dominant)
            line["roofline_all"] = roofs
        # ---- inference, BASELINE configs[1] (fp32 as the config states; bf16 reported beside it)
        dec.eval()
        mem_i = memory[:INFER["B"]]
        infer = {"workload": f"BASELINE configs[1]: batch {INFER['B']}, {INFER['N']} tokens, {INFER['steps']} fixed "
                             "decoder steps (gate ignored)"}
        for prec in ("fp32", "bf16"):
            dec.precision = prec
            for _ in range(2):
                dec.inference(mem_i, ignore_gate=True, max_decoder_steps=INFER["steps"])
            torch.cuda.synchronize()
            e0.record()
            reps = 3
            for _ in range(reps):
                dec.inference(mem_i, ignore_gate=True, max_decoder_steps=INFER["steps"])
            e1.record()
            torch.cuda.synchronize()
            ims = e0.elapsed_time(e1) / reps
            infer[prec] = {"decoder_steps_per_s": INFER["steps"] / (ims * 1e-3),
                           "mel_frames_per_s": INFER["B"] * INFER["steps"] / (ims * 1e-3), "us_per_step": 1e3 * ims / INFER["steps"]}
        dec.precision = args.precision
        infer["decoder_steps_per_s"] = infer["fp32"]["decoder_steps_per_s"]       # headline: the config's own precision
        line["infer"] = infer
        dec.train()
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = {k: v for k, v in cpu_train_leg(2, 1).items() if k != "ms_per_step"}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--reference-device", choices=["cpu", "cuda"], default="cpu",
                    help="--impl reference only: 'cuda' times the same torch ops on the GPU (stock PyTorch eager path); "
                         "the contract's reference arm is the CPU one")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16",
                    help="arithmetic of the recurrent GEMMs (BASELINE configs[2] is bf16; fp32 = parity mode)")
    args = ap.parse_args()

    # watchdog: a run that has not finished after 10 minutes (the default run takes about one) dumps every thread's Python
    # stack and exits instead of hanging its caller; GVX_BENCH_TRACE=1 adds per-stage progress lines on stderr
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("GVX_BENCH_TRACE_AFTER", "600")), exit=True)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
