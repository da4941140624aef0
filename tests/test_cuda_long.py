"""GPU parity at the sequence lengths BASELINE.json actually runs (600 / 800 / 1600 frames), against the CPU oracle run on the
GPU box's host cores in the same test.

Rounding compounds over hundreds of recurrent steps, so these tests are what turns "parity at T <= 40" into parity of the
benchmarked workloads:
  * configs[0] shape (B=16, N=120, T=600), fp32 mode: outputs within the north_star bound (1e-4 of max|ref|), alignment argmax
    exact wherever the fp64-free oracle margin is clear, gradients within 2e-4 ... measured bound printed;
  * configs[2] shape (N=150, T=800), bf16 mode, fused persistent chains (forward + BPTT): the whole batch of 64 rows forward
    against the oracle with the same rounding points, 16 rows forward + backward (the oracle's autograd graph of 64 rows x
    800 frames is ~13 GB of host memory: bounded on purpose); fused vs per-step chain at full size; bit-exact run to run;
  * configs[4] token count (B=32, N=300) forward + backward at T=64.
The bounds asserted here are the ones DESIGN.md quotes.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import decoder_oracle as O
from oracle import synth
from test_cuda_parity import argmax_agrees, make_decoder

pytestmark = pytest.mark.gpu


def _u(seed, stream, shape):
    return torch.from_numpy((synth.uniform01(seed, stream, int(np.prod(shape))) - 0.5).astype(np.float32).reshape(shape))


def _gpu_run(dec, mem, mel, lens, r_mel, r_gate, seed, dev, backward=True):
    import genvox_b200
    dec.zero_grad(set_to_none=True)
    memory = torch.from_numpy(mem).to(dev).requires_grad_(backward)
    dec.set_dropout_seed(seed)
    m, g, a = dec(memory, torch.from_numpy(mel).to(dev), torch.from_numpy(lens).to(dev))
    grads = None
    if backward:
        ((m * r_mel.to(dev)).sum() + (g * r_gate.to(dev)).sum()).backward()
        grads = {k: p.grad.detach().cpu().clone() for k, p in dec.named_parameters()}
        grads["memory"] = memory.grad.detach().cpu().clone()
    torch.cuda.synchronize()
    genvox_b200.check_device_errors()
    return m.detach().cpu(), g.detach().cpu(), a.detach().cpu(), grads


def test_config0_shape_fp32_600_frames(cuda_device):
    """BASELINE configs[0]: B=16, N=120, T=600, fp32 mode, teacher-forced forward + BPTT."""
    dims = synth.DecoderDims()
    B, N, T, seed = 16, 120, 600, 777
    W = synth.make_decoder_weights(29, dims)
    mem, mel, lens = synth.make_inputs(83, B, N, T, dims, ragged=True)
    r_mel, r_gate = _u(83, 20, (B, dims.n_mels, T)), _u(83, 21, (B, T))
    (om, og, oa), ograds, omem = O.loss_and_grads(O.as_params(W), torch.from_numpy(mem), torch.from_numpy(mel), lens, r_mel, r_gate,
                                                  seed, True, dims.p_attention_dropout, dims.p_decoder_dropout)
    dec = make_decoder(dims, W, cuda_device, True)
    m, g, a, grads = _gpu_run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device)
    errs = {"mel": rel_err(m, om), "gate": rel_err(g, og), "align": rel_err(a, oa)}
    gerrs = {k: rel_err(grads[k], ograds[k]) for k in ograds}
    gerrs["memory"] = rel_err(grads["memory"], omem)
    print("T=600 fp32 vs fp32 oracle:", errs, "worst grad:", max(gerrs.items(), key=lambda kv: kv[1]))
    assert all(v < 1e-4 for v in errs.values()), errs                      # north_star: fp32 <= 1e-4 relative
    ok, frac = argmax_agrees(a.numpy(), oa.numpy().astype(np.float64), margin=1e-4)
    assert ok and frac > 0.5, frac                                           # alignment argmax exact where the margin is clear
    # ReLU ties: with 2.4 M prenet activations a pre-activation can land within fp32 summation noise of zero, and the two
    # implementations then pick different sides of relu'(0) (forward value ~1e-8: invisible; that unit's row of d W0 moves
    # by one sample's contribution).  Those units are identified from the fp64 pre-activations and excluded - nothing else.
    k0 = "prenet.layers.0.linear_layer.weight"
    frames = torch.cat((torch.zeros(1, B, dims.n_mels), torch.from_numpy(mel).permute(2, 0, 1)[: T - 1]), 0).double()
    z1 = frames @ torch.from_numpy(W[k0]).double().t()
    row_err = np.abs(grads[k0].numpy() - ograds[k0].numpy()).max(1) / np.abs(ograds[k0].numpy()).max()
    bad = [int(u) for u in np.nonzero(row_err > 5e-4)[0]]
    zmax = float(z1.abs().max())
    for u in bad:       # every deviating unit must have a (non-go-frame) pre-activation within fp32 summation noise of zero
        zu = z1[:, :, u].abs()
        assert float(zu[zu > 0].min()) < 1e-6 * zmax, (u, float(zu[zu > 0].min()), zmax)
    assert len(bad) <= 3, bad
    keep = [u for u in range(dims.prenet_dim) if u not in bad]
    gerrs[k0] = float(row_err[keep].max())
    print("prenet units with a ReLU tie (excluded from d W0):", bad, "d W0 err without them:", gerrs[k0])
    assert all(v < 5e-4 for v in gerrs.values()), gerrs


def test_config2_shape_bf16_800_frames(cuda_device):
    """BASELINE configs[2]: N=150, T=800, bf16 mode, the fused persistent chains (the benchmarked path)."""
    from genvox_b200 import _native
    lib = _native.load()
    dims = synth.DecoderDims()
    N, T, seed = 150, 800, 4242
    W = synth.make_decoder_weights(31, dims)
    P = O.as_params(W)

    # ---- all 64 rows, forward: fused vs the same-rounding oracle, fused vs per-step chain, bit-exact run to run
    B = 64
    mem, mel, lens = synth.make_inputs(89, B, N, T, dims, ragged=True)
    r_mel, r_gate = _u(89, 20, (B, dims.n_mels, T)), _u(89, 21, (B, T))
    dec = make_decoder(dims, W, cuda_device, True)
    dec.precision = "bf16"
    lib.gvx_debug_option(b"fused", 1)
    try:
        f1 = _gpu_run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device)
        f2 = _gpu_run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device)
        lib.gvx_debug_option(b"fused", 0)
        s1 = _gpu_run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device)
    finally:
        lib.gvx_debug_option(b"fused", -1)
    for x, y in zip(f1[:3], f2[:3]):
        assert bool(torch.isfinite(x).all()) and torch.equal(x, y)         # 800 steps of barriers / tagged exchanges: deterministic
    for k in f1[3]:
        assert torch.equal(f1[3][k], f2[3][k]), k
    fs = {"mel": rel_err(f1[0], s1[0]), "gate": rel_err(f1[1], s1[1]), "align": rel_err(f1[2], s1[2])}
    fsg = {k: rel_err(f1[3][k], s1[3][k]) for k in f1[3]}
    print("T=800 bf16 fused vs per-step chain:", fs, "worst grad:", max(fsg.items(), key=lambda kv: kv[1]))
    with torch.no_grad(), O.bf16_semantics(round_memory=True):
        om, og, oa = O.forward_teacher(P, torch.from_numpy(mem), torch.from_numpy(mel), lens, seed, True,
                                       dims.p_attention_dropout, dims.p_decoder_dropout)
    fo = {"mel": rel_err(f1[0], om), "gate": rel_err(f1[1], og), "align": rel_err(f1[2], oa)}
    print("T=800 bf16 fused vs same-rounding oracle (64 rows, forward):", fo)
    assert float((f1[2].sum(-1) - 1).abs().max()) < 1e-5
    for b in range(B):
        if int(lens[b]) < N:
            assert float(f1[2][b, :, int(lens[b]):].abs().max()) == 0.0
    # stated bf16 bound at the benchmarked length: 5e-3 of max|ref| on outputs against the same-rounding
    # oracle (measured 1.1e-3), 1e-2 against the per-step chain (measured 4.4e-3), 3e-2 on gradients (measured 7.8e-3)
    assert all(v < 5e-3 for v in fo.values()), fo
    assert all(v < 1e-2 for v in fs.values()), fs
    assert all(np.isfinite(v) and v < 3e-2 for v in fsg.values()), fsg

    # ---- 16 rows, forward + BPTT against the oracle's autograd
    B = 16
    mem, mel, lens = mem[:B], mel[:B], lens[:B]
    r_mel, r_gate = r_mel[:B], r_gate[:B]
    f = _gpu_run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device)
    with O.bf16_semantics(round_memory=True):
        (om, og, oa), ograds, omem = O.loss_and_grads(P, torch.from_numpy(mem), torch.from_numpy(mel), lens, r_mel, r_gate, seed, True,
                                                      dims.p_attention_dropout, dims.p_decoder_dropout)
    go = {k: rel_err(f[3][k], ograds[k]) for k in ograds}
    go["memory"] = rel_err(f[3]["memory"], omem)
    print("T=800 bf16 fused BPTT vs same-rounding oracle (16 rows):", {k: f"{v:.1e}" for k, v in go.items()})
    assert all(np.isfinite(v) and v < 3e-2 for v in go.values()), go          # measured: <= 6.8e-3


def test_config4_token_count_300(cuda_device):
    """BASELINE configs[4] token count: B=32, N=300, bf16 mode, forward + BPTT at T=64 against the same-rounding oracle."""
    dims = synth.DecoderDims()
    B, N, T, seed = 32, 300, 64, 99
    W = synth.make_decoder_weights(37, dims)
    mem, mel, lens = synth.make_inputs(97, B, N, T, dims, ragged=True)
    r_mel, r_gate = _u(97, 20, (B, dims.n_mels, T)), _u(97, 21, (B, T))
    dec = make_decoder(dims, W, cuda_device, True)
    dec.precision = "bf16"
    m, g, a, grads = _gpu_run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device)
    with O.bf16_semantics(round_memory=True):
        (om, og, oa), ograds, omem = O.loss_and_grads(O.as_params(W), torch.from_numpy(mem), torch.from_numpy(mel), lens, r_mel, r_gate,
                                                      seed, True, dims.p_attention_dropout, dims.p_decoder_dropout)
    errs = {"mel": rel_err(m, om), "gate": rel_err(g, og), "align": rel_err(a, oa)}
    gerrs = {k: rel_err(grads[k], ograds[k]) for k in ograds}
    gerrs["memory"] = rel_err(grads["memory"], omem)
    print("N=300 bf16 vs same-rounding oracle:", errs, "worst grad:", max(gerrs.items(), key=lambda kv: kv[1]))
    assert all(v < 1e-2 for v in errs.values()), errs
    assert all(np.isfinite(v) and v < 3e-2 for v in gerrs.values()), gerrs
