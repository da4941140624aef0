"""GPU parity tests of bf16 mode (gate / query / projection GEMMs on tcgen05 with bf16 operands, fp32
accumulation; pointwise math, cell state and attention in fp32).

Two yardsticks:
  * the CPU oracle with the SAME rounding points (oracle.decoder_oracle.bf16_semantics): differences are
    accumulation order plus rare one-ulp bf16 rounding flips of an activation -> tight bounds;
  * the fp64 run of the unmodified reference (tests/golden): the stated bf16 bound of north_star,
    3e-2 of max|ref| on mel / gate / alignments for the short goldens.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import decoder_oracle as O
from oracle import synth
from test_cuda_parity import make_decoder

pytestmark = pytest.mark.gpu

SAME_ROUNDING_FWD_TOL = 5e-3
SAME_ROUNDING_GRAD_TOL = 3e-2
BF16_VS_FP64_TOL = 3e-2


@pytest.mark.parametrize("B,N,T,training,small", [(8, 50, 12, True, False), (64, 60, 4, True, False), (3, 33, 6, False, False),
                                                  (3, 11, 7, True, True)])
def test_bf16_forward_backward_match_same_rounding_oracle(cuda_device, B, N, T, training, small):
    dims = synth.SMALL_DIMS if small else synth.DecoderDims()     # small dims exercise the K / row padding of the operand images
    seed = 777
    W = synth.make_decoder_weights(17, dims)
    mem, mel, lens = synth.make_inputs(61, B, N, T, dims)
    u = lambda s, shape: torch.from_numpy((synth.uniform01(61, s, int(np.prod(shape))) - 0.5).astype(np.float32).reshape(shape))
    r_mel, r_gate = u(20, (B, dims.n_mels, T)), u(21, (B, T))
    with O.bf16_semantics():
        (om, og, oa), ograds, omem = O.loss_and_grads(O.as_params(W), torch.from_numpy(mem), torch.from_numpy(mel), lens, r_mel,
                                                      r_gate, seed, training, dims.p_attention_dropout, dims.p_decoder_dropout)
    dec = make_decoder(dims, W, cuda_device, training)
    dec.precision = "bf16"
    memory = torch.from_numpy(mem).to(cuda_device).requires_grad_(True)
    dec.set_dropout_seed(seed)
    m, g, a = dec(memory, torch.from_numpy(mel).to(cuda_device), torch.from_numpy(lens).to(cuda_device))
    errs = {"mel": rel_err(m.detach().cpu(), om), "gate": rel_err(g.detach().cpu(), og), "align": rel_err(a.detach().cpu(), oa)}
    print("bf16 fwd vs same-rounding oracle:", errs)
    assert all(np.isfinite(v) and v < SAME_ROUNDING_FWD_TOL for v in errs.values()), errs
    ((m * r_mel.to(cuda_device)).sum() + (g * r_gate.to(cuda_device)).sum()).backward()
    gerrs = {k: rel_err(p.grad.cpu(), ograds[k]) for k, p in dec.named_parameters()}
    gerrs["memory"] = rel_err(memory.grad.cpu(), omem)
    print("bf16 grads vs same-rounding oracle:", {k: f"{v:.1e}" for k, v in gerrs.items()})
    assert all(np.isfinite(v) and v < SAME_ROUNDING_GRAD_TOL for v in gerrs.values()), gerrs


@pytest.mark.parametrize("B,N,steps,masked", [(5, 40, 9, True), (64, 150, 6, False)])
def test_bf16_inference_matches_same_rounding_oracle(cuda_device, B, N, steps, masked):
    dims = synth.DecoderDims()
    W = synth.make_decoder_weights(19, dims)
    mem, _, lens = synth.make_inputs(63, B, N, 0, dims, ragged=masked)
    with O.bf16_semantics():
        rm, rg, ra, _ = O.inference(O.as_params(W), torch.from_numpy(mem), lens if masked else None, steps, 0.5, True, 4321)
    dec = make_decoder(dims, W, cuda_device, False)
    dec.precision = "bf16"
    dec.set_dropout_seed(4321)
    m, g, a = dec.inference(torch.from_numpy(mem).to(cuda_device),
                            memory_lengths=torch.from_numpy(lens).to(cuda_device) if masked else None,
                            ignore_gate=True, max_decoder_steps=steps)
    errs = {"mel": rel_err(m.cpu(), rm), "gate": rel_err(g.cpu(), rg), "align": rel_err(a.cpu(), ra)}
    print("bf16 inference vs same-rounding oracle:", errs)
    assert all(np.isfinite(v) and v < SAME_ROUNDING_FWD_TOL for v in errs.values()), errs


@pytest.mark.parametrize("tag", ["fwd_default_train", "fwd_default_eval_b1"])
def test_bf16_within_stated_bound_of_fp64_reference(cuda_device, tag):
    meta, z = load_golden(tag)
    dims = synth.DecoderDims(**meta["dims"])
    W = synth.make_decoder_weights(meta["weight_seed"], dims, meta["weight_scale"])
    B, N, T = meta["B"], meta["N"], meta["T"]
    mem, mel, lens = synth.make_inputs(meta["input_seed"], B, N, T, dims)
    dec = make_decoder(dims, W, cuda_device, meta["training"])
    dec.precision = "bf16"
    dec.set_dropout_seed(meta["dropout_seed"])
    with torch.no_grad():
        m, g, a = dec(torch.from_numpy(mem).to(cuda_device), torch.from_numpy(mel).to(cuda_device),
                      torch.from_numpy(lens).to(cuda_device))
    errs = {k: rel_err(v.cpu(), z[f"{k}_f64"]) for k, v in (("mel", m), ("gate", g), ("align", a))}
    print(tag, "bf16 vs fp64 reference:", errs)
    assert all(np.isfinite(v) and v < BF16_VS_FP64_TOL for v in errs.values()), errs


def test_bf16_full_size_is_deterministic_and_finite(cuda_device):
    dims = synth.DecoderDims()
    W = synth.make_decoder_weights(7, dims)
    mem, mel, lens = synth.make_inputs(71, 64, 150, 40, dims, ragged=False)
    dec = make_decoder(dims, W, cuda_device, True)
    dec.precision = "bf16"
    outs = []
    for _ in range(2):
        dec.zero_grad(set_to_none=True)
        memory = torch.from_numpy(mem).to(cuda_device).requires_grad_(True)
        dec.set_dropout_seed(5)
        m, g, a = dec(memory, torch.from_numpy(mel).to(cuda_device), torch.from_numpy(lens).to(cuda_device))
        (m.square().mean() + g.square().mean()).backward()
        outs.append((m.detach().clone(), a.detach().clone(), memory.grad.clone(),
                     dec.decoder_rnn.weight_hh.grad.clone(), dec.attention_rnn.weight_ih.grad.clone()))
    for x, y in zip(*outs):
        assert bool(torch.isfinite(x).all())
        assert torch.equal(x, y)
    assert float((outs[0][1].sum(-1) - 1).abs().max().cpu()) < 1e-5
