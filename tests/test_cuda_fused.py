"""GPU tests of the fused persistent attention chain (genvox_b200/csrc/gvx_fused_fwd.cuh): attention LSTM + query +
location-sensitive attention of ALL teacher-forced frames in one launch (tacotron2.py:338-353).

Yardstick 1: the per-step kernel chain of the same library on the same inputs (gvx_debug_option("fused", 0)) - the two
paths share every rounding point except that the fused path contracts the context with a bf16 copy of the encoder
memory (the context is rounded to bf16 right afterwards in both), so they agree to a few bf16 ulps.
Yardstick 2: the CPU oracle with the same bf16 rounding points, including the bf16 memory operand of the context
(oracle.decoder_oracle.bf16_semantics(round_memory=True)); bounds of tests/test_cuda_bf16.py."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import decoder_oracle as O
from oracle import synth
from test_cuda_parity import make_decoder

pytestmark = pytest.mark.gpu


def _run(dec, mem, mel, lens, r_mel, r_gate, seed, dev, fused):
    from genvox_b200 import _native
    lib = _native.load()
    lib.gvx_debug_option(b"fused", 1 if fused else 0)
    try:
        dec.zero_grad(set_to_none=True)
        memory = torch.from_numpy(mem).to(dev).requires_grad_(True)
        dec.set_dropout_seed(seed)
        m, g, a = dec(memory, torch.from_numpy(mel).to(dev), torch.from_numpy(lens).to(dev))
        ((m * r_mel.to(dev)).sum() + (g * r_gate.to(dev)).sum()).backward()
        torch.cuda.synchronize()
        import genvox_b200
        genvox_b200.check_device_errors()
        grads = {k: p.grad.detach().cpu().clone() for k, p in dec.named_parameters()}
        grads["memory"] = memory.grad.detach().cpu().clone()
        return m.detach().cpu(), g.detach().cpu(), a.detach().cpu(), grads
    finally:
        lib.gvx_debug_option(b"fused", -1)


@pytest.mark.parametrize("B,N,T,training,ragged", [(64, 150, 24, True, False), (64, 150, 9, False, True), (5, 37, 11, True, True),
                                                   (33, 160, 7, True, True), (1, 8, 5, True, False), (16, 120, 30, True, True),
                                                       # N > 160 at B <= 32: per-step forward chain + persistent BPTT chain (configs[4] path)
                                                       (16, 200, 9, True, True), (32, 320, 6, True, False)])
def test_fused_chain_matches_per_step_chain_and_oracle(cuda_device, B, N, T, training, ragged):
    dims = synth.DecoderDims()
    seed = 4242
    W = synth.make_decoder_weights(23, dims)
    mem, mel, lens = synth.make_inputs(67, B, N, T, dims, ragged=ragged)
    u = lambda s, shape: torch.from_numpy((synth.uniform01(67, s, int(np.prod(shape))) - 0.5).astype(np.float32).reshape(shape))
    r_mel, r_gate = u(20, (B, dims.n_mels, T)), u(21, (B, T))
    dec = make_decoder(dims, W, cuda_device, training)
    dec.precision = "bf16"
    fm, fg, fa, fgr = _run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device, True)
    sm, sg, sa, sgr = _run(dec, mem, mel, lens, r_mel, r_gate, seed, cuda_device, False)
    errs = {"mel": rel_err(fm, sm), "gate": rel_err(fg, sg), "align": rel_err(fa, sa)}
    gerrs = {k: rel_err(fgr[k], sgr[k]) for k in fgr}
    print("fused vs per-step:", errs, {k: f"{v:.1e}" for k, v in gerrs.items()})
    # (with a handful of gate logits a single bf16 rounding flip of h is a visible fraction of max|gate|)
    tol = {"mel": 5e-3, "gate": 5e-3 if B * T >= 64 else 2e-2, "align": 5e-3}
    assert all(np.isfinite(v) and v < tol[k] for k, v in errs.items()), errs
    assert all(np.isfinite(v) and v < 3e-2 for v in gerrs.values()), gerrs
    # masked tokens carry exactly zero weight; every row is a distribution
    for b in range(B):
        assert float(fa[b, :, int(lens[b]):].abs().max()) == 0.0 if int(lens[b]) < N else True
    assert float((fa.sum(-1) - 1).abs().max()) < 1e-5

    if T <= 12:
        with O.bf16_semantics(round_memory=True):
            (om, og, oa), ograds, omem = O.loss_and_grads(O.as_params(W), torch.from_numpy(mem), torch.from_numpy(mel), lens, r_mel,
                                                          r_gate, seed, training, dims.p_attention_dropout, dims.p_decoder_dropout)
        oerrs = {"mel": rel_err(fm, om), "gate": rel_err(fg, og), "align": rel_err(fa, oa)}
        print("fused vs same-rounding oracle:", oerrs)
        assert all(np.isfinite(v) and v < tol[k] for k, v in oerrs.items()), oerrs
        ogerrs = {k: rel_err(fgr[k], ograds[k]) for k in ograds}
        ogerrs["memory"] = rel_err(fgr["memory"], omem)
        assert all(np.isfinite(v) and v < 3e-2 for v in ogerrs.values()), ogerrs


def test_fused_chain_is_deterministic(cuda_device):
    dims = synth.DecoderDims()
    W = synth.make_decoder_weights(7, dims)
    B, N, T = 64, 150, 40
    mem, mel, lens = synth.make_inputs(71, B, N, T, dims, ragged=True)
    r_mel, r_gate = torch.ones(B, dims.n_mels, T), torch.ones(B, T)
    dec = make_decoder(dims, W, cuda_device, True)
    dec.precision = "bf16"
    a = _run(dec, mem, mel, lens, r_mel, r_gate, 5, cuda_device, True)
    b = _run(dec, mem, mel, lens, r_mel, r_gate, 5, cuda_device, True)
    for x, y in zip(a[:3], b[:3]):
        assert bool(torch.isfinite(x).all()) and torch.equal(x, y)
    for k in a[3]:
        assert torch.equal(a[3][k], b[3][k]), k
