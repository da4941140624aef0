"""GPU tests of the persistent LSTM-chain kernels (genvox_b200/csrc/gvx_persist.cuh) on their own.

Forward: gates_t = pre_t + bf16(h_{t-1}) . bf16(W_hh)^T, nn.LSTMCell pointwise part
(/root/reference/models/tts/tacotron2.py:357), carried-state dropout (:358) from the Philox stream.
Backward: d loss / d pre_t for loss = sum_t <h_t, dh_ext_t> (BPTT of the same chain).
Reference: the same recurrence in torch fp64 with the same rounding points (h rounded to bf16 with a
straight-through gradient), autograd for the backward.  The kernel accumulates in fp32 and rounds the d-gates
to bf16 for the recurrent GEMM, so a bf16 rounding can flip: tolerances are a few bf16 ulps of max|ref|."""
import numpy as np
import pytest
import torch

from oracle import philox

pytestmark = pytest.mark.gpu


def _reference(w_hh, pre_um, dh_ext, B, T, H, p, seed, training):
    dev = w_hh.device
    Wb = w_hh.bfloat16().double()
    pre = pre_um.double().clone().requires_grad_(True)          # [T, B, 4H] columns 4u+g
    h = torch.zeros(B, H, dtype=torch.float64, device=dev)
    c = torch.zeros(B, H, dtype=torch.float64, device=dev)
    hs, cs, gs = [], [c], []
    for t in range(T):
        rec = (h @ Wb.t()).view(B, 4, H).permute(0, 2, 1)        # [B, H, 4] gate order i,f,g,o
        g = pre[t].view(B, H, 4) + rec
        gi, gf, gg, go = torch.sigmoid(g[..., 0]), torch.sigmoid(g[..., 1]), torch.tanh(g[..., 2]), torch.sigmoid(g[..., 3])
        c = gf * c + gi * gg
        hh = go * torch.tanh(c)
        if training and p > 0:
            keep = torch.from_numpy(philox.keep_mask(seed, philox.SITE_DEC, t, B, H, p)).to(dev)
            hh = hh * keep.double() * float(philox.dropout_scale(p))
        h = hh + (hh.bfloat16().double() - hh).detach()          # bf16 operand, straight-through gradient
        hs.append(h)
        cs.append(c)
        gs.append(torch.stack([gi, gf, gg, go], dim=-1).reshape(B, 4 * H))
    hs = torch.stack(hs)
    loss = (hs * dh_ext.double()).sum()
    (dpre,) = torch.autograd.grad(loss, pre)
    return hs.detach(), torch.stack(cs).detach(), torch.stack(gs).detach(), dpre


@pytest.mark.parametrize("B,T,H,p,training", [(64, 12, 1024, 0.1, True), (5, 7, 64, 0.0, False), (33, 20, 256, 0.1, True),
                                              (16, 9, 192, 0.25, True), (64, 40, 1024, 0.1, False)])
def test_persistent_lstm_chain_fwd_bwd(cuda_device, B, T, H, p, training):
    from genvox_b200 import _native
    from genvox_b200.decoder import _ptr, _stream
    lib = _native.load()
    g = torch.Generator().manual_seed(B * 131 + T * 7 + H)
    k = 1.0 / np.sqrt(H)
    w_hh = ((torch.rand(4 * H, H, generator=g) * 2 - 1) * k).to(cuda_device)
    pre = ((torch.rand(T, B, 4 * H, generator=g) * 2 - 1) * 1.5).to(cuda_device)
    dh = ((torch.rand(T, B, H, generator=g) * 2 - 1)).to(cuda_device)
    seed = 0x1234_5678_9ABC
    h_out = torch.full((T, B, H), float("nan"), device=cuda_device)
    c_out = torch.full((T + 1, B, H), float("nan"), device=cuda_device)
    gates = torch.full((T, B, 4 * H), float("nan"), device=cuda_device)
    dgates = torch.full((T, B, 4 * H), float("nan"), device=cuda_device)
    _native.check(lib.gvx_test_lstm_chain(_ptr(w_hh), _ptr(pre), B, T, H, p, seed, int(training), _ptr(h_out), _ptr(c_out),
                                          _ptr(gates), _ptr(dh), _ptr(dgates), _stream()), "gvx_test_lstm_chain")
    torch.cuda.synchronize()
    rh, rc, rg, rd = _reference(w_hh, pre, dh, B, T, H, p, seed, training)

    def err(a, b):
        return float((a.double() - b).abs().max() / b.abs().max())

    errs = {"h": err(h_out, rh), "c": err(c_out, rc), "gates": err(gates, rg), "dgates": err(dgates, rd)}
    print(B, T, H, errs)
    assert all(np.isfinite(v) for v in errs.values()), errs
    assert errs["h"] < 1e-2 and errs["c"] < 1e-2 and errs["gates"] < 1e-2, errs
    assert errs["dgates"] < 3e-2, errs
    # mean error is far below the flip-limited maximum
    assert float((h_out.double() - rh).abs().mean()) < 2e-4
    assert float((dgates.double() - rd).abs().mean() / rd.abs().mean()) < 1e-2


def test_persistent_chain_is_deterministic(cuda_device):
    from genvox_b200 import _native
    from genvox_b200.decoder import _ptr, _stream
    lib = _native.load()
    B, T, H = 64, 25, 1024
    g = torch.Generator().manual_seed(5)
    w_hh = ((torch.rand(4 * H, H, generator=g) * 2 - 1) / 32).to(cuda_device)
    pre = ((torch.rand(T, B, 4 * H, generator=g) * 2 - 1)).to(cuda_device)
    dh = ((torch.rand(T, B, H, generator=g) * 2 - 1)).to(cuda_device)
    outs = []
    for _ in range(2):
        h_out = torch.empty((T, B, H), device=cuda_device)
        c_out = torch.empty((T + 1, B, H), device=cuda_device)
        gates = torch.empty((T, B, 4 * H), device=cuda_device)
        dgates = torch.empty((T, B, 4 * H), device=cuda_device)
        _native.check(lib.gvx_test_lstm_chain(_ptr(w_hh), _ptr(pre), B, T, H, 0.1, 77, 1, _ptr(h_out), _ptr(c_out), _ptr(gates),
                                              _ptr(dh), _ptr(dgates), _stream()), "gvx_test_lstm_chain")
        torch.cuda.synchronize()
        outs.append((h_out, c_out, gates, dgates))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
