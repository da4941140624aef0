"""Known-answer tests for the shared dropout stream (oracle/philox.py)."""
import numpy as np

from oracle import philox


def _kat(c, k):
    return [int(x) for x in philox.philox4x32_10(*[np.uint64(v) for v in c], k[0], k[1])]


def test_random123_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert _kat((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _kat((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _kat((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_thresholds_and_rates():
    assert philox.dropout_threshold(0.5) == 1 << 31
    assert philox.dropout_threshold(0.1) == 429496729
    assert float(philox.dropout_scale(0.5)) == 2.0
    m = philox.keep_mask(123, philox.SITE_ATT, 3, 64, 1024, 0.1)
    assert abs(m.mean() - 0.9) < 0.01
    m = philox.keep_mask(123, philox.SITE_PRENET0, 3, 64, 256, 0.5)
    assert abs(m.mean() - 0.5) < 0.02


def test_stream_is_keyed_by_every_coordinate():
    base = philox.random_words(1, 2, 3, 4, 16)
    assert base.shape == (4, 16)
    assert not np.array_equal(base, philox.random_words(2, 2, 3, 4, 16))
    assert not np.array_equal(base, philox.random_words(1, 3, 3, 4, 16))
    assert not np.array_equal(base, philox.random_words(1, 2, 4, 4, 16))
    # row_offset shifts rows (data-parallel ranks can draw the rows of a global batch)
    off = philox.random_words(1, 2, 3, 2, 16, row_offset=2)
    assert np.array_equal(base[2:], off)
    # 64-bit seed: the high word matters
    assert not np.array_equal(base, philox.random_words(1 + (1 << 32), 2, 3, 4, 16))
