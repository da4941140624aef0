"""Parameter containers with the reference's names, shapes and initialisers.

The CUDA path never calls these modules' forward(); they exist so that a genvox_b200 Decoder has
exactly the `state_dict` of the reference Decoder and the same random init:
`linear_layer` / `conv` sub-module names and Xavier-uniform(gain) follow
/root/reference/models/generic.py:5-54, the module tree follows tacotron2.py:23-144,:259-301.
"""
import torch
from torch import nn


def _xavier(weight, gain_name):
    nn.init.xavier_uniform_(weight, gain=nn.init.calculate_gain(gain_name))


class LinearParams(nn.Module):
    """state_dict: linear_layer.weight [out, in] (+ linear_layer.bias)."""

    def __init__(self, in_dim, out_dim, bias=True, w_init_gain="linear"):
        super().__init__()
        self.linear_layer = nn.Linear(in_dim, out_dim, bias=bias)
        _xavier(self.linear_layer.weight, w_init_gain)


class ConvParams(nn.Module):
    """state_dict: conv.weight [out, in, k] (no bias on the decoder path)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=False, w_init_gain="linear"):
        super().__init__()
        assert kernel_size % 2 == 1
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, padding=(kernel_size - 1) // 2, bias=bias)
        _xavier(self.conv.weight, w_init_gain)


class PrenetParams(nn.Module):
    def __init__(self, in_dim, sizes):
        super().__init__()
        ins = [in_dim] + list(sizes[:-1])
        self.layers = nn.ModuleList([LinearParams(i, o, bias=False) for i, o in zip(ins, sizes)])


class LocationParams(nn.Module):
    def __init__(self, n_filters, kernel_size, attention_dim):
        super().__init__()
        self.location_conv = ConvParams(2, n_filters, kernel_size, bias=False)
        self.location_dense = LinearParams(n_filters, attention_dim, bias=False, w_init_gain="tanh")


class AttentionParams(nn.Module):
    def __init__(self, attention_rnn_dim, embedding_dim, attention_dim, n_filters, kernel_size):
        super().__init__()
        self.query_layer = LinearParams(attention_rnn_dim, attention_dim, bias=False, w_init_gain="tanh")
        self.memory_layer = LinearParams(embedding_dim, attention_dim, bias=False, w_init_gain="tanh")
        self.v = LinearParams(attention_dim, 1, bias=False)
        self.location_layer = LocationParams(n_filters, kernel_size, attention_dim)
