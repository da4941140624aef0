"""TEST / MEASUREMENT INFRASTRUCTURE — import the UNMODIFIED reference.

In the build container the reference is imported from /root/reference (oracle/make_golden.py, the CPU tests that are
skipped when it is absent).  /root/reference does not exist on the GPU box: there the byte-identical copy staged by
oracle/build_ref.py under oracle/_ref (git-ignored, travels with the snapshot) is used - by bench.py's reference arm and
baselines only; the `-m gpu` tests and smoke() never need it.

The reference's import chain pulls four third-party modules that are not installed and
that the decoder never touches: yt_dlp, matplotlib, inflect, g2p_en.
They are replaced by empty stub modules.
"""
import contextlib
import os
import sys
import types

import torch

from . import philox

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("GENVOX_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "tts", "tacotron2.py")) and os.path.isdir(_STAGED):
    REFERENCE_ROOT = _STAGED


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "tts", "tacotron2.py"))


def _install_stubs():
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("yt_dlp")
    mpl = stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = stub("matplotlib.pyplot")
    stub("inflect", engine=lambda *a, **k: None)
    stub("g2p_en", G2p=object)
    try:
        import wandb  # noqa: F401
    except Exception:
        stub("wandb")


def import_reference():
    """Returns the reference's `models.tts.tacotron2` module (unmodified source)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.tts.tacotron2 as ref_t2  # type: ignore
    return ref_t2


def build_reference_decoder(dims, weights, dtype=torch.float32):
    """Reference Decoder(**dims) with `weights` (dict of numpy arrays, state_dict names) loaded."""
    ref = import_reference()
    dec = ref.Decoder(**dims.kwargs())
    sd = {k: torch.as_tensor(v) for k, v in weights.items()}
    missing, unexpected = dec.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return dec.to(dtype)


def build_reference_model(n_tokens=64, seed=0):
    """The reference's full Tacotron2 (tacotron2.py:416-448) with its default configs (configs/models.py, configs/__init__.py)."""
    ref = import_reference()
    import configs  # type: ignore
    from configs.models import Tacotron2Config  # type: ignore
    torch.manual_seed(seed)
    return ref.Tacotron2(Tacotron2Config(), configs.AudioConfig(), configs.TextConfig(n_tokens=n_tokens))


@contextlib.contextmanager
def philox_dropout_patch(seed: int, mode: str, row_offset: int = 0, skip: int = 0, count_inactive: bool = True):
    """Replace torch.nn.functional.dropout, while a reference Decoder method runs, by a mask
    provider that replays the shared Philox stream (oracle/philox.py).

    mode "forward":   call order inside Decoder.forward (tacotron2.py:373 then :341,:358 per step)
                      = prenet0[all frames], prenet1[all frames], (att_t, dec_t) for t = 0..T-1
    mode "inference": call order inside Decoder.inference (:398, :341, :358)
                      = (prenet0_t, prenet1_t, att_t, dec_t) for t = 0, 1, ...
    Around a whole Tacotron2 (whose encoder / postnet call F.dropout too): `skip` = number of leading calls that belong to the
    encoder and pass through unchanged; `count_inactive=False` ignores calls with training=False (an eval-mode encoder in
    front of a train-mode decoder) instead of counting them.
    """
    import torch.nn.functional as F
    orig = F.dropout
    calls = {"n": 0}

    def site_and_t(n):
        if mode == "forward":
            if n < 2:
                return (philox.SITE_PRENET0, philox.SITE_PRENET1)[n], None
            n -= 2
            return (philox.SITE_ATT, philox.SITE_DEC)[n % 2], n // 2
        return (philox.SITE_PRENET0, philox.SITE_PRENET1, philox.SITE_ATT, philox.SITE_DEC)[n % 4], n // 4

    def patched(x, p=0.5, training=True, inplace=False):
        if not training and not count_inactive:
            return x
        if calls.get("skipped", 0) < skip:
            calls["skipped"] = calls.get("skipped", 0) + 1
            return orig(x, p, training, inplace)
        site, t = site_and_t(calls["n"])
        calls["n"] += 1
        if not training or p == 0.0:
            return x
        scale = float(philox.dropout_scale(p))
        if x.dim() == 3:   # prenet over all frames: [F, B, W]
            assert t is None
            keep = torch.stack([torch.from_numpy(philox.keep_mask(seed, site, f, x.shape[1], x.shape[2], p, row_offset))
                                for f in range(x.shape[0])])
        else:
            assert t is not None
            keep = torch.from_numpy(philox.keep_mask(seed, site, t, x.shape[0], x.shape[1], p, row_offset))
        return x * (keep.to(x.dtype) * scale)

    F.dropout = patched
    try:
        yield calls
    finally:
        F.dropout = orig
