"""TEST INFRASTRUCTURE — numpy Philox4x32-10, the host mirror of the in-kernel dropout RNG.

The reference draws its dropout masks from torch's global RNG
(`F.dropout`, /root/reference/models/tts/tacotron2.py:143,341,358), which cannot be
reproduced inside a CUDA kernel.  The product therefore defines its own
counter-based stream and the oracle / the golden generator consume the *same*
stream by patching ``torch.nn.functional.dropout`` (see ref_import.py).

Stream definition (shared with genvox_b200/csrc/gvx_philox.cuh):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (j >> 2, row, t, site)            # j = feature index, row = batch row
    word    = philox4x32_10(counter, key)[j & 3]
    keep    = word >= threshold(p),  threshold(p) = floor(p * 2**32)
    y       = keep ? x * float32(1 / (1 - p)) : 0

sites: 0 = prenet layer 0, 1 = prenet layer 1 (p = 0.5, always on,
tacotron2.py:143), 2 = attention-LSTM hidden (tacotron2.py:341), 3 = decoder-LSTM
hidden (tacotron2.py:358).  ``t`` is the frame index for the prenet sites (frame 0 is
the all-zero go frame, tacotron2.py:370-373) and the decoder step for sites 2/3.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

SITE_PRENET0, SITE_PRENET1, SITE_ATT, SITE_DEC = 0, 1, 2, 3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counters are uint32 arrays (broadcastable), key two ints.

    Returns four uint32 arrays.
    """
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0          # 32x32 -> 64 bit products, no overflow in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def dropout_threshold(p: float) -> int:
    """uint32 threshold: an element is kept iff its random word >= threshold."""
    assert 0.0 <= p < 1.0
    return int(np.floor(float(p) * 4294967296.0))


def dropout_scale(p: float) -> np.float32:
    return np.float32(1.0 / (1.0 - float(p)))


def random_words(seed: int, site: int, t: int, rows: int, width: int, row_offset: int = 0) -> np.ndarray:
    """uint32 words [rows, width] of the stream for (seed, site, t)."""
    j = np.arange(width, dtype=np.uint64)[None, :]
    r = (np.arange(rows, dtype=np.uint64) + np.uint64(row_offset))[:, None]
    w = philox4x32_10(j >> np.uint64(2), r, np.uint64(t), np.uint64(site), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = (j & np.uint64(3)).astype(np.int64)
    sel = np.broadcast_to(sel, w[0].shape)
    out = np.where(sel == 0, w[0], np.where(sel == 1, w[1], np.where(sel == 2, w[2], w[3])))
    return out.astype(np.uint32)


def keep_mask(seed: int, site: int, t: int, rows: int, width: int, p: float, row_offset: int = 0) -> np.ndarray:
    """bool [rows, width]; True = element kept."""
    return random_words(seed, site, t, rows, width, row_offset) >= np.uint32(dropout_threshold(p))


def uniform01(seed: int, stream: int, n: int) -> np.ndarray:
    """float64 uniforms in [0, 1) for synthetic data: counter = (i>>2, 0, stream, 0xD47A)."""
    i = np.arange(n, dtype=np.uint64)
    w = philox4x32_10(i >> np.uint64(2), np.uint64(0), np.uint64(stream), np.uint64(0xD47A),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = (i & np.uint64(3)).astype(np.int64)
    words = np.where(sel == 0, w[0], np.where(sel == 1, w[1], np.where(sel == 2, w[2], w[3])))
    return words.astype(np.float64) / 4294967296.0
