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


def _clk_vector(oracle, blk_lo, blk_hi, draw, colour, key, pm, periodic):
    """contract v3 (csrc/clock_word.cuh, oracle/rng_contract.c clk_vector_uniforms), restated in Python: the accept and
    proposal uniforms (as 32-bit integers U, u = (U + 1) 2^-32) of the 16 sites of a vector"""
    c3 = lambda sub: ((draw >> 32) & 0xFFFF) | (colour << 16) | (sub << 24)
    X = np.concatenate([oracle.philox([blk_lo, blk_hi, draw & 0xFFFFFFFF, c3(i)], key) for i in range(3)])
    Y = np.concatenate([oracle.philox([blk_lo, blk_hi, draw & 0xFFFFFFFF, c3(4 + i)], key) for i in range(3)])
    half = lambda w, hs: (int(w) >> 16) if hs else (int(w) & 0xFFFF)
    ua, up, digits = [], [], []
    for w in range(4):
        W = int(X[3 * w + 2])
        for e in range(4):
            a16, a16b = half(X[3 * w + (e & 1)], e >> 1), half(Y[3 * w + (e & 1)], e >> 1)
            ua.append(((a16 & 0x7FFF) << 17) | ((a16 >> 15) << 16) | a16b)
            d = (W * pm) >> 32
            # what the reference computes from u = (U + 1) 2^-32: ceiling(u pm) - 1 (periodic) or min(floor(u pm), pm - 1)
            cell = lambda U: (-(-((U + 1) * pm) // 2 ** 32) - 1) if periodic else min(((U + 1) * pm) >> 32, pm - 1)
            U = W if cell(W) == d else (W - 1) % 2 ** 32
            assert cell(U) == d
            up.append(U); digits.append(d)
            W = (W * pm) & 0xFFFFFFFF
    return ua, up, digits


def test_clock_contract_v3_assembly(oracle):
    """the periodic-clock and helical-clock uniform arrays, site by site, from the Philox blocks the contract names"""
    TAG_TORUS, TAG_CLOCK = 0x544F5253, 0x434C4F4B
    seed, draw, rep, nx, ny = 42, 3, 2, 72, 6          # nx/2 = 36: three vectors per row, the last one partial
    for q in (6, 5):
        rn = oracle.torus_uniforms(seed, draw, rep, nx, ny, q).reshape(ny, nx, 2)
        nvr = (nx // 2 + 15) // 16
        for y0 in range(ny):
            for v in range(nvr):
                for colour in range(2):
                    ua, up, dg = _clk_vector(oracle, y0 * nvr + v, rep, draw, colour, [seed, TAG_TORUS], q - 1, True)
                    for j in range(16):
                        xi = 16 * v + j
                        if xi >= nx // 2:
                            break
                        x0 = 2 * xi + ((y0 + colour) & 1)
                        assert rn[y0, x0, 0] == (up[j] + 1) * 2.0 ** -32 and rn[y0, x0, 1] == (ua[j] + 1) * 2.0 ** -32
                        assert int(np.ceil(rn[y0, x0, 0] * (q - 1))) == dg[j] + 1      # the reference's own expression (:142)
    n = 101 * 100
    L = (n // 2 + 15) // 16
    ra, rp = oracle.clock_uniforms(seed, draw, rep, n, 6)
    for p in list(range(0, L, 29)) + [L - 1]:
        for colour in range(2):
            ua, up, dg = _clk_vector(oracle, p, 0, draw, colour, [seed, TAG_CLOCK + rep], 6, False)
            for lane in range(16):
                k = lane * L + p
                if k >= n // 2:
                    continue
                i = 2 * k + colour
                assert ra[i] == (ua[lane] + 1) * 2.0 ** -32 and rp[i] == (up[lane] + 1) * 2.0 ** -32
                assert min(int(np.floor(rp[i] * 6)), 5) == dg[lane]                    # src/clock_gpu_m.f90:211


def test_clock_contract_v3_statistics(oracle):
    """accept uniforms: (0, 1], 32-bit resolution, flat first-look field and flat lazily evaluated low half; proposals: the
    digits are uniform, the digits of the four sites that share a proposal word are pairwise independent, and proposal and
    accept uniforms are uncorrelated"""
    nx, ny, q = 512, 256, 6
    rn = oracle.torus_uniforms(5, 1, 0, nx, ny, q).reshape(-1, 2)
    ua, upr = rn[:, 1], rn[:, 0]
    for u in (ua, upr):
        assert u.min() > 0.0 and u.max() <= 1.0
        k = u * 2.0 ** 32
        assert np.array_equal(k, np.round(k)) and abs(u.mean() - 0.5) < 0.005
    h, _ = np.histogram(ua, bins=256, range=(0, 1))
    chi2 = ((h - ua.size / 256) ** 2 / (ua.size / 256)).sum()
    assert chi2 < 360, chi2                           # 255 dof: mean 255, sd ~23
    low = (ua * 2.0 ** 32 - 1).astype(np.uint64) & 0xFFFF      # the lazily evaluated low half
    h2, _ = np.histogram(low, bins=64, range=(0, 65536))
    assert ((h2 - ua.size / 64) ** 2 / (ua.size / 64)).sum() < 120
    assert abs(np.corrcoef(ua, upr)[0, 1]) < 0.01
    # digits per site, on the colour-compact grid: sites xi = 4 g + e, e = 0..3 share one proposal word
    dig = (np.ceil(upr * (q - 1)).astype(np.int64) - 1).reshape(ny, nx)
    cnt = np.bincount(dig.ravel(), minlength=q - 1)
    assert cnt.size == q - 1 and (((cnt - dig.size / (q - 1)) ** 2) / (dig.size / (q - 1))).sum() < 25     # 4 dof
    # x0 = 2 xi + P: on even rows colour 0 has P = 0; take colour-0 sites of even rows
    d0 = dig[0::2, 0::2].reshape(ny // 2, nx // 8, 4)       # [row, word, e]
    for e1 in range(4):
        for e2 in range(e1 + 1, 4):
            joint = np.zeros((q - 1, q - 1))
            np.add.at(joint, (d0[:, :, e1].ravel(), d0[:, :, e2].ravel()), 1)
            exp = d0[:, :, 0].size / (q - 1) ** 2
            assert (((joint - exp) ** 2) / exp).sum() < 60, (e1, e2)          # 24 dof: mean 24, sd ~7


def test_xy_contract_v2_assembly_and_statistics(oracle):
    """XY periodic: 23 + 23 bits per site from three Philox blocks per 8 sites (csrc/xy.cu, oracle/rng_contract.c orc_xy_uniforms),
    restated here; both uniforms in (0, 1] with exactly 23-bit resolution, uncorrelated with each other and between the two
    rows of a pair that share the block of low halves"""
    TAG_XY = 0x58593244
    seed, draw, nx, ny = 9, 5, 44, 6            # nx/2 = 22: six groups per row, the last one partial
    r, c = oracle.xy_uniforms(seed, draw, nx, ny)
    r, c = r.reshape(ny, nx), c.reshape(ny, nx)
    gpr = (nx // 2 + 3) // 4
    for y0 in range(ny):
        for x0 in range(nx):
            colour, xi = (x0 + y0) & 1, x0 >> 1
            c3 = lambda sub: ((draw >> 32) & 0xFFFF) | (colour << 16) | (sub << 24)
            R = oracle.philox([y0 * gpr + (xi >> 2), 0, draw, c3(0)], [seed, TAG_XY])
            C = oracle.philox([(y0 & ~1) * gpr + (xi >> 2), 0, draw, c3(1)], [seed, TAG_XY])
            W, cw = int(R[xi & 3]), int(C[xi & 3])
            ur = (((W >> 24) & 0x7F) << 16) | ((cw >> 16) if (y0 & 1) else (cw & 0xFFFF))
            assert r[y0, x0] == (ur + 1) * 2.0 ** -23 and c[y0, x0] == ((W & 0x7FFFFF) + 1) * 2.0 ** -23
    r, c = oracle.xy_uniforms(1, 2, 512, 256)
    for u in (r, c):
        assert u.min() > 0.0 and u.max() <= 1.0 and abs(u.mean() - 0.5) < 0.005
        k = u * 2.0 ** 23
        assert np.array_equal(k, np.round(k))
        assert np.array_equal(u, u.astype(np.float32).astype(np.float64))       # exact in fp32: the GPU sees the same values
        h, _ = np.histogram(u, bins=256, range=(0, 1))
        assert ((h - u.size / 256) ** 2 / (u.size / 256)).sum() < 360
    assert abs(np.corrcoef(r, c)[0, 1]) < 0.01
    r2 = r.reshape(256, 512)
    assert abs(np.corrcoef(r2[0::2].ravel(), r2[1::2].ravel())[0, 1]) < 0.01
