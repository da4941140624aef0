"""CPU tests of the host logic of genvox_b200.synthesis with a stand-in decoder: ordering, padding,
per-row trimming, batching and sharding.  (The real decoder needs a GPU: tests/test_cuda_synthesis.py.)"""
import torch
import torch.nn as nn

from genvox_b200 import synthesis


class _StubDecoder(nn.Module):
    """Decodes `n_tok + 2` frames for a row of n_tok valid tokens; frame t of a row is (sum of its valid memory) + t."""

    def __init__(self):
        super().__init__()
        self.calls = []
        self.last_n_frames = None

    def inference(self, memory, memory_lengths=None, ignore_gate=False, max_decoder_steps=None):
        B, N, _ = memory.shape
        self.calls.append((B, N, memory_lengths.tolist()))
        assert memory_lengths.tolist() == sorted(memory_lengths.tolist(), reverse=True)        # longest first
        for b in range(B):                                                                      # padding really is zero
            assert float(memory[b, int(memory_lengths[b]):].abs().sum()) == 0.0
        n_frames = memory_lengths.to(torch.int32) + 2
        T = int(n_frames.max())
        base = memory.sum(dim=(1, 2))
        mel = base[:, None, None] + torch.arange(T, dtype=torch.float32)[None, None, :].expand(B, 3, T)
        gate = torch.zeros(B, T)
        align = torch.zeros(B, T, N)
        self.last_n_frames = n_frames
        return mel, gate, align


class _StubModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.embedding = nn.Embedding(50, 4)
        self.encoder = type("E", (nn.Module,), {"inference": lambda self, x: x.transpose(1, 2) * 2.0})()
        self.decoder = _StubDecoder()
        self.postnet = lambda m: 0.5 * m


def test_batched_inference_orders_pads_and_trims():
    torch.manual_seed(0)
    m = _StubModel()
    rows = [[1, 2, 3], [4] * 9, [5, 6], [7] * 5, [8]]
    out = synthesis.batched_inference(m, rows, max_batch=2)
    assert [c[0] for c in m.decoder.calls] == [2, 2, 1]                     # 5 utterances in batches of 2, longest first
    assert m.decoder.calls[0][2] == [9, 5] and m.decoder.calls[1][2] == [3, 2] and m.decoder.calls[2][2] == [1]
    for i, r in enumerate(rows):
        o = out[i]
        n = len(r)
        assert o["mel_outputs"].shape == (1, 3, n + 2) and o["gate_outputs"].shape == (1, n + 2)
        assert o["alignments"].shape == (1, n + 2, n)                        # trimmed to the utterance's own tokens
        expect = (m.embedding(torch.tensor(r)) * 2.0).sum()                  # encoder ran on the unpadded utterance
        assert torch.allclose(o["mel_outputs"][0, 0, 0], expect, atol=1e-5)
        assert torch.allclose(o["mel_outputs_postnet"], 1.5 * o["mel_outputs"])   # mel + postnet(mel), tacotron2.py:490-491


def test_sharded_inference_partitions_the_utterances():
    m = _StubModel()
    rows = [[i + 1] * (i + 1) for i in range(7)]
    seen = []
    for rank in range(4):
        lo, outs = synthesis.sharded_inference(m, rows, rank, 4)
        seen += list(range(lo, lo + len(outs)))
        for j, o in enumerate(outs):
            assert o["alignments"].shape[2] == len(rows[lo + j])
    assert seen == list(range(7))
