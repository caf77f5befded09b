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


def _clk_pair(r, r2, e):
    """contract v2 (csrc/clock_word.cuh, oracle/rng_contract.c clk_uniform_pair), restated in Python"""
    half = lambda w, hs: (int(w) >> 16) if hs else (int(w) & 0xFFFF)
    hs = e >> 1
    a16, a16b = half(r[e & 1], hs), half(r2[e & 1], hs)
    p16, p16b = half(r[2 + (e & 1)], hs), half(r2[2 + (e & 1)], hs)
    return ((a16 & 0x7FFF) << 17) | ((a16 >> 15) << 16) | a16b, (p16 << 16) | p16b


def test_clock_contract_v2_assembly(oracle):
    """the periodic-clock and helical-clock uniform arrays, site by site, from the Philox blocks the contract names"""
    TAG_TORUS, TAG_CLOCK = 0x544F5253, 0x434C4F4B
    seed, draw, rep, nx, ny = 42, 3, 2, 72, 6          # nx/2 = 36: three vectors per row, the last one partial
    rn = oracle.torus_uniforms(seed, draw, rep, nx, ny).reshape(ny, nx, 2)
    nvr = (nx // 2 + 15) // 16
    for y0 in range(ny):
        for x0 in range(0, nx, 5):
            colour, xi = (x0 + y0) & 1, x0 >> 1
            j, blk = xi & 15, y0 * nvr + (xi >> 4)
            c3 = lambda sub: ((draw >> 32) & 0xFFFF) | (colour << 16) | (sub << 24)
            r = oracle.philox([blk, 0, draw, c3(j >> 2)], [seed, TAG_TORUS + rep])
            r2 = oracle.philox([blk, 0, draw, c3(4 + (j >> 2))], [seed, TAG_TORUS + rep])
            ua, up = _clk_pair(r, r2, j & 3)
            assert rn[y0, x0, 0] == (up + 1) * 2.0 ** -32 and rn[y0, x0, 1] == (ua + 1) * 2.0 ** -32
    n = 101 * 100
    L = (n // 2 + 15) // 16
    ra, rp = oracle.clock_uniforms(seed, draw, rep, n)
    for i in range(0, n, 37):
        colour, k = i & 1, i >> 1
        lane, p = k // L, k % L
        c3 = lambda sub: ((draw >> 32) & 0xFFFF) | (colour << 16) | (sub << 24)
        r = oracle.philox([p, 0, draw, c3(lane >> 2)], [seed, TAG_CLOCK + rep])
        r2 = oracle.philox([p, 0, draw, c3(4 + (lane >> 2))], [seed, TAG_CLOCK + rep])
        ua, up = _clk_pair(r, r2, lane & 3)
        assert ra[i] == (ua + 1) * 2.0 ** -32 and rp[i] == (up + 1) * 2.0 ** -32


def test_clock_contract_v2_statistics(oracle):
    """accept and proposal uniforms: (0, 1], 32-bit resolution, flat first-look fields, no correlation between the two or
    between the sites that share a Philox word"""
    rn = oracle.torus_uniforms(5, 1, 0, 512, 256).reshape(-1, 2)
    for col in (0, 1):
        u = rn[:, col]
        assert u.min() > 0.0 and u.max() <= 1.0
        k = u * 2.0 ** 32
        assert np.array_equal(k, np.round(k)) and abs(u.mean() - 0.5) < 0.005
        h, _ = np.histogram(u, bins=256, range=(0, 1))
        chi2 = ((h - u.size / 256) ** 2 / (u.size / 256)).sum()
        assert chi2 < 360, chi2                       # 255 dof: mean 255, sd ~23
        low = (k - 1).astype(np.uint64) & 0xFFFF      # the lazily evaluated low half
        h2, _ = np.histogram(low, bins=64, range=(0, 65536))
        assert ((h2 - u.size / 64) ** 2 / (u.size / 64)).sum() < 120
    assert abs(np.corrcoef(rn[:, 0], rn[:, 1])[0, 1]) < 0.01
    ua = rn[:, 1].reshape(256, 512)
    # sites e and e + 2 of a word (compact positions xi and xi + 2 -> x0 and x0 + 4) share a Philox word
    assert abs(np.corrcoef(ua[:, :-4].ravel(), ua[:, 4:].ravel())[0, 1]) < 0.01
