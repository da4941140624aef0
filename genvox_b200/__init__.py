"""genvox_b200 — B200-native (sm_100a) Tacotron2 decoder recurrence for saiakarsh193/GenVox.

Only what the hot path needs:
  csrc/         hand-written CUDA kernels + the C ABI (include/genvox_b200.h)
  _native.py    ctypes binding of that ABI
  decoder.py    drop-in `Decoder` (reference: models/tts/tacotron2.py:258-414) and `install(model)`
  containers.py parameter containers with the reference's state_dict names
  build.py      in-tree nvcc build
"""
from .decoder import Decoder, install, decoder_kwargs_from  # noqa: F401
from . import _native  # noqa: F401
from ._native import check_device_errors  # noqa: F401

__all__ = ["Decoder", "install", "decoder_kwargs_from", "check_device_errors"]
