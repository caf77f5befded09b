"""CPU: the RNG contract restated in oracle/ (Philox4x32-10 + lazy 32-bit uniforms)."""
import numpy as np


def test_philox_random123_kat(oracle):
    """known-answer vectors published with Random123 (kat_vectors, philox4x32-10)"""
    assert oracle.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniforms_range_and_resolution(oracle):
    n = 31 * 31 * 30
    u = oracle.ising_uniforms(42, 0, n)
    assert u.min() > 0.0 and u.max() <= 1.0           # (0, 1], cuRAND's convention
    k = u * 2.0 ** 32
    assert np.array_equal(k, np.round(k))             # exactly 32-bit resolution
    assert abs(u.mean() - 0.5) < 0.01
    # different draws / seeds / colours decorrelate
    v = oracle.ising_uniforms(42, 1, n)
    w = oracle.ising_uniforms(43, 0, n)
    assert abs(np.corrcoef(u, v)[0, 1]) < 0.03 and abs(np.corrcoef(u, w)[0, 1]) < 0.03
    assert abs(np.corrcoef(u[0::2], u[1::2])[0, 1]) < 0.03


def test_fast_generator_identical(oracle):
    for n in (3 * 3 * 2, 31 * 31 * 30, 101 * 100, 1001 * 1000):
        assert np.array_equal(oracle.ising_uniforms(7, 3, n), oracle.ising_uniforms_fast(7, 3, n))


def test_uniform_histogram_flat(oracle):
    u = oracle.ising_uniforms(1, 5, 1001 * 1000)
    h, _ = np.histogram(u, bins=128, range=(0, 1))    # the 7-bit first stage must be flat
    exp = u.size / 128
    chi2 = ((h - exp) ** 2 / exp).sum()
    assert chi2 < 200, chi2                           # 127 dof: mean 127, sd ~16
    lo = (u * 2 ** 32 - 1).astype(np.uint64) & 0x1FFFFFF   # 25-bit second stage
    h2, _ = np.histogram(lo, bins=64, range=(0, 2 ** 25))
    chi2 = ((h2 - u.size / 64) ** 2 / (u.size / 64)).sum()
    assert chi2 < 120, chi2
