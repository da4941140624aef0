"""TEST INFRASTRUCTURE — deterministic synthetic decoder weights and inputs.

Weights follow the reference initialisers' *distributions* (Xavier-uniform with the
gain of /root/reference/models/generic.py:15-18,49-52; nn.LSTMCell's
U(-1/sqrt(H), 1/sqrt(H)); nn.Linear bias U(-1/sqrt(in), 1/sqrt(in))) but are drawn
from the Philox stream in philox.py rather than torch's RNG, so the identical tensors
can be rebuilt on the GPU box where /root/reference does not exist.
"""
import math
from dataclasses import dataclass, asdict

import numpy as np

from .philox import uniform01


@dataclass(frozen=True)
class DecoderDims:
    """Constructor arguments of the reference Decoder (tacotron2.py:259-273); defaults
    are Tacotron2Config's (configs/models.py:10-33) and AudioConfig.n_mels = 80."""
    n_mels: int = 80
    encoder_embedding_dim: int = 512
    decoder_rnn_dim: int = 1024
    prenet_dim: int = 256
    max_decoder_steps: int = 1000
    gate_threshold: float = 0.5
    p_attention_dropout: float = 0.1
    p_decoder_dropout: float = 0.1
    attention_rnn_dim: int = 1024
    attention_dim: int = 128
    attention_location_n_filters: int = 32
    attention_location_kernel_size: int = 31

    def kwargs(self):
        return asdict(self)


SMALL_DIMS = DecoderDims(n_mels=8, encoder_embedding_dim=16, decoder_rnn_dim=32, prenet_dim=8,
                         max_decoder_steps=20, attention_rnn_dim=32, attention_dim=8,
                         attention_location_n_filters=4, attention_location_kernel_size=5)

_GAIN = {"linear": 1.0, "tanh": 5.0 / 3.0, "sigmoid": 1.0}


def param_shapes(d: DecoderDims):
    """name -> (shape, bound) in state_dict order."""
    E, H, A, P, M = d.encoder_embedding_dim, d.decoder_rnn_dim, d.attention_rnn_dim, d.prenet_dim, d.n_mels
    D, F, K = d.attention_dim, d.attention_location_n_filters, d.attention_location_kernel_size

    def xavier(out_f, in_f, gain="linear", rf=1):
        return _GAIN[gain] * math.sqrt(6.0 / ((in_f + out_f) * rf))

    return {
        "prenet.layers.0.linear_layer.weight": ((P, M), xavier(P, M)),
        "prenet.layers.1.linear_layer.weight": ((P, P), xavier(P, P)),
        "attention_rnn.weight_ih": ((4 * A, P + E), 1.0 / math.sqrt(A)),
        "attention_rnn.weight_hh": ((4 * A, A), 1.0 / math.sqrt(A)),
        "attention_rnn.bias_ih": ((4 * A,), 1.0 / math.sqrt(A)),
        "attention_rnn.bias_hh": ((4 * A,), 1.0 / math.sqrt(A)),
        "attention_layer.query_layer.linear_layer.weight": ((D, A), xavier(D, A, "tanh")),
        "attention_layer.memory_layer.linear_layer.weight": ((D, E), xavier(D, E, "tanh")),
        "attention_layer.v.linear_layer.weight": ((1, D), xavier(1, D)),
        "attention_layer.location_layer.location_conv.conv.weight": ((F, 2, K), xavier(F, 2, "linear", K)),
        "attention_layer.location_layer.location_dense.linear_layer.weight": ((D, F), xavier(D, F, "tanh")),
        "decoder_rnn.weight_ih": ((4 * H, A + E), 1.0 / math.sqrt(H)),
        "decoder_rnn.weight_hh": ((4 * H, H), 1.0 / math.sqrt(H)),
        "decoder_rnn.bias_ih": ((4 * H,), 1.0 / math.sqrt(H)),
        "decoder_rnn.bias_hh": ((4 * H,), 1.0 / math.sqrt(H)),
        "linear_projection.linear_layer.weight": ((M, H + E), xavier(M, H + E)),
        "linear_projection.linear_layer.bias": ((M,), 1.0 / math.sqrt(H + E)),
        "gate_layer.linear_layer.weight": ((1, H + E), xavier(1, H + E, "sigmoid")),
        "gate_layer.linear_layer.bias": ((1,), 1.0 / math.sqrt(H + E)),
    }


def make_decoder_weights(seed: int, d: DecoderDims = DecoderDims(), scale: float = 1.0):
    """dict name -> float32 ndarray, uniform in (-bound*scale, bound*scale)."""
    out = {}
    for stream, (name, (shape, bound)) in enumerate(param_shapes(d).items()):
        n = int(np.prod(shape))
        u = uniform01(seed, 1000 + stream, n)
        out[name] = ((2.0 * u - 1.0) * bound * scale).astype(np.float32).reshape(shape)
    return out


def _normalish(seed: int, stream: int, n: int) -> np.ndarray:
    """Irwin-Hall(4) approximation of N(0,1); exact reproducibility matters, shape does not."""
    u = uniform01(seed, stream, 4 * n).reshape(4, n)
    return (u.sum(0) - 2.0) * math.sqrt(3.0)


def make_inputs(seed: int, B: int, N: int, T: int, d: DecoderDims = DecoderDims(), ragged: bool = True):
    """memory [B,N,E] (encoder outputs, |x|<~1), mel_in [B,n_mels,T], memory_lengths [B]
    (sorted descending, max == N, as the collate guarantees, models/tts/__init__.py:32)."""
    E, M = d.encoder_embedding_dim, d.n_mels
    memory = (0.5 * _normalish(seed, 1, B * N * E)).astype(np.float32).reshape(B, N, E)
    mel_in = _normalish(seed, 2, B * M * max(T, 1)).astype(np.float32).reshape(B, M, max(T, 1))[:, :, :T]
    if ragged and B > 1:
        u = uniform01(seed, 3, B)
        lengths = np.sort((N // 2 + np.floor(u * (N - N // 2 + 1))).astype(np.int64))[::-1].copy()
        lengths = np.clip(lengths, 1, N)
        lengths[0] = N
    else:
        lengths = np.full([B], N, dtype=np.int64)
    return memory, np.ascontiguousarray(mel_in), lengths
