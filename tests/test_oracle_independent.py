"""CPU: a second, independent restatement of the reference kernels -- vectorised numpy written straight from the
Fortran sources (periodic torus views / index arithmetic instead of the C oracle's site loops and halo cells) --
against the C oracle on the same uniforms.  The reference has no tests of its own, so the oracle
is pinned by agreement of two restatements that share no code.  Integer models: bit-exact; XY: 1e-12."""
import math

import numpy as np
import pytest

_exp = np.vectorize(math.exp)   # libm, like the host code of the reference and the C oracle (np.exp may differ in the last bit)
_cos = np.vectorize(math.cos)


def _ising2d_sweep(s, nx, ny, u, exparr):
    """src/ising2d_gpu_m.f90:133-162,191-196 on the ring of nall = nx * ny sites (0-based i = idx - 1): neighbours
    i +- 1, i +- nx (helical); odd idx (even i) first, then even idx; flip iff u(idx) <= exparr(dE)."""
    n = nx * ny
    for colour in (0, 1):                      # offset = 1: idx = 1, 3, ... -> i = 0, 2, ...
        i = np.arange(colour, n, 2)
        nb = s[(i + 1) % n] + s[(i - 1) % n] + s[(i + nx) % n] + s[(i - nx) % n]
        de = 2 * s[i] * nb                     # :195
        flip = u[i] <= exparr[de + 8]          # exparr(-8:8), :159
        s[i] = np.where(flip, -s[i], s[i])
    return s


def _ising3d_sweep(s, nx, ny, nz, u, ws):
    """src/ising3d_gpu_m.f90:174-206: s in {0, 1}, S = number of up neighbours among i +- 1, +- nx, +- nx ny;
    flip (s <- 1 - s) iff u(idx) <= ws(S, s)"""
    nxy, n = nx * ny, nx * ny * nz
    for colour in (0, 1):
        i = np.arange(colour, n, 2)
        S = s[(i + 1) % n] + s[(i - 1) % n] + s[(i + nx) % n] + s[(i - nx) % n] + s[(i + nxy) % n] + s[(i - nxy) % n]
        flip = u[i] <= ws[s[i], S]
        s[i] = np.where(flip, 1 - s[i], s[i])
    return s


@pytest.mark.parametrize("shape", [(5, 4), (33, 32), (101, 100)])
def test_ising2d_numpy_restatement(oracle, shape):
    nx, ny = shape
    n = nx * ny
    o = oracle.ising2d_gpu().init(nx, ny, 2.26918531421, 3)
    beta = o.beta()
    exparr = np.array([1.0 if d <= 0 else math.exp(-beta * d) for d in range(-8, 9)])     # :126-130
    assert np.array_equal(exparr, o.exparr)
    rng = np.random.default_rng(1)
    u0 = 1.0 - rng.random(n)
    o.set_random_spin(u0)
    s = np.where(u0 < 0.5, 1, -1).astype(np.int64)                                       # :83
    assert np.array_equal(o.spins()[nx:nx + n], s)
    for sweep in range(5):
        u = 1.0 - rng.random(n)
        u[rng.integers(0, n, 4)] = exparr[12]                                            # ties: u == exp(-4 beta) exactly
        o.update(randoms=u)
        s = _ising2d_sweep(s, nx, ny, u, exparr)
        full = o.spins()
        assert np.array_equal(full[nx:nx + n], s), sweep
        assert np.array_equal(full[:nx], s[n - nx:]) and np.array_equal(full[nx + n:], s[:nx])   # norishiro, :103-105
        assert o.calc_magne_sum() == int(s.sum())
        assert o.calc_energy_sum() == -int((s * (np.roll(s, -1) + np.roll(s, -nx))).sum())      # :209


@pytest.mark.parametrize("shape", [(3, 3, 2), (7, 5, 6), (15, 17, 16)])
def test_ising3d_numpy_restatement(oracle, shape):
    nx, ny, nz = shape
    nxy, n = nx * ny, nx * ny * nz
    o = oracle.ising3d_gpu().init(nx, ny, nz, 4.51152, 3)
    beta = o.beta()
    # ws(S, 0) = min(1, exp(-beta (e2 - e1))), ws(S, 1) = min(1, exp(-beta (e1 - e2))), e1 = 2S - 6, e2 = 6 - 2S, :153-170
    ws = np.empty((2, 7))
    for S in range(7):
        e1, e2 = 2 * S - 6, 6 - 2 * S
        ws[0, S] = min(1.0, math.exp(-beta * (e2 - e1)))
        ws[1, S] = min(1.0, math.exp(-beta * (e1 - e2)))
    assert np.allclose(ws, np.asarray(o.ws).reshape(2, 7), rtol=0, atol=0)
    rng = np.random.default_rng(2)
    u0 = 1.0 - rng.random(n)
    o.set_random_spin(u0)
    s = (u0 < 0.5).astype(np.int64)                                                      # :99
    for sweep in range(4):
        u = 1.0 - rng.random(n)
        o.update(randoms=u)
        s = _ising3d_sweep(s, nx, ny, nz, u, ws)
        assert np.array_equal(o.spins()[nxy:nxy + n], s), sweep
        fwd = np.roll(s, -1) + np.roll(s, -nx) + np.roll(s, -nxy)                        # :253
        e = int(np.where(s == 1, -(2 * fwd - 3), 2 * fwd - 3).sum())
        assert o.calc_energy_sum() == e and o.calc_magne_sum() == 2 * int(s.sum()) - n


def test_xy_periodic_numpy_restatement(oracle):
    """src/xy2d_periodic_gpu_m.f90:353-439,496-534 on torus views (np.roll) instead of the halo frame"""
    nx, ny, kbt = 16, 12, 0.89
    o = oracle.xy2d_gpu().init(nx, ny, kbt, 1)
    rng = np.random.default_rng(3)
    th = 2 * np.pi * (1.0 - rng.random((ny, nx)))
    o.set_angles(th)
    c, s = np.cos(th), np.sin(th)
    yy, xx = np.meshgrid(np.arange(1, ny + 1), np.arange(1, nx + 1), indexing="ij")       # 1-based like the source
    beta = 1 / kbt

    def field(a):   # s(x+1) + s(x-1) + s(y+1) + s(y-1), :394-396 (the order matters to the last bit)
        return np.roll(a, -1, 1) + np.roll(a, 1, 1) + np.roll(a, -1, 0) + np.roll(a, 1, 0)

    for sweep in range(3):
        r = 1.0 - rng.random((ny, nx))
        cand = 1.0 - rng.random((ny, nx))
        o.update(r, cand)
        cc, cs = np.cos(2 * np.pi * cand), np.sin(2 * np.pi * cand)
        for offset in (0, 1):                                    # (x + y + offset) even, :377-380
            m = ((xx + yy + offset) % 2) == 0
            hx, hy = field(c), field(s)
            de = -((cc - c) * hx + (cs - s) * hy)
            acc = m & (r <= np.exp(-beta * de))
            c, s = np.where(acc, cc, c), np.where(acc, cs, s)
        sp = o.spins()
        assert np.abs(sp[0, 1:-1, 1:-1] - c).max() < 1e-12 and np.abs(sp[1, 1:-1, 1:-1] - s).max() < 1e-12, sweep
        # over-relaxation: s <- 2 (h^ . s) h^ - s, renormalised, per colour (:418-439)
        o.update_over_relaxation(1)
        for offset in (0, 1):
            m = ((xx + yy + offset) % 2) == 0
            hx, hy = field(c), field(s)
            hn = np.hypot(hx, hy)
            ux, uy = hx / hn, hy / hn
            d = ux * c + uy * s
            nc_, ns_ = 2 * d * ux - c, 2 * d * uy - s
            nn = np.hypot(nc_, ns_)
            c, s = np.where(m, nc_ / nn, c), np.where(m, ns_ / nn, s)
        sp = o.spins()
        assert np.abs(sp[0, 1:-1, 1:-1] - c).max() < 1e-12 and np.abs(sp[1, 1:-1, 1:-1] - s).max() < 1e-12
        e = -(c * (np.roll(c, -1, 1) + np.roll(c, -1, 0)) + s * (np.roll(s, -1, 1) + np.roll(s, -1, 0))).sum()   # :505-506
        assert abs(o.calc_energy_sum() - e) < 1e-9 and abs(o.calc_magne_sum() - c.sum()) < 1e-9 and abs(o.calc_magne_y_sum() - s.sum()) < 1e-9


def test_clock_tableall_numpy_restatement(oracle):
    """src/clock/clock_tableall_gpu_m.f90:57-152: q^6 table from broadcasting, periodic neighbours from np.roll,
    proposal c + ceiling(r1 (q - 1)) mod q, accept iff r2 <= states_to_prob(c, new, r, u, l, d)"""
    nx, ny, q, kbt = 12, 8, 6, 0.91
    o = oracle.clock_tableall(nx, ny, kbt, q)
    psi = 2 * (4 * np.arctan(1.0)) / q
    k = np.arange(q)
    e3 = -_cos((k[None, :, None] - k[:, None, None]) * psi) - _cos((k[None, None, :] - k[:, None, None]) * psi)   # (c, u|l, r|d), :27-33
    C, N, R, U, L, D = np.meshgrid(k, k, k, k, k, k, indexing="ij")
    de = e3[N, R, U] - e3[C, R, U] + e3[N, L, D] - e3[C, L, D]                             # :72-75
    prob = np.where(de <= 0, 1.0, _exp(-(1 / kbt) * de))
    assert np.array_equal(prob.ravel(order="F"), o.prob)                                   # Fortran order: c fastest
    rng = np.random.default_rng(4)
    st = np.zeros((ny, nx), dtype=np.int64)                                                # sixclock(x, y) -> [y][x]
    yy, xx = np.meshgrid(np.arange(1, ny + 1), np.arange(1, nx + 1), indexing="ij")
    for sweep in range(6):
        rnds = 1.0 - rng.random((ny, nx, 2))                                               # rnds(2, nx, ny): [y][x][0:2]
        o.update_metropolis(rnds)
        for parity in (0, 1):                                                              # (x + y) & 1 == parity, :96-101,120
            m = ((xx + yy) % 2) == parity
            r_, l_ = np.roll(st, -1, 1), np.roll(st, 1, 1)
            u_, d_ = np.roll(st, -1, 0), np.roll(st, 1, 0)
            new = (st + np.ceil(rnds[..., 0] * (q - 1)).astype(np.int64)) % q              # :142
            acc = m & (rnds[..., 1] <= prob[st, new, r_, u_, l_, d_])                      # :146
            st = np.where(acc, new, st)
        assert np.array_equal(o.c.reshape(ny, nx), st), sweep
    assert abs(o.calc_magne() - np.cos(psi * st).sum() / (nx * ny)) < 1e-12


@pytest.mark.parametrize("multi", [False, True])
def test_clock_helical_numpy_restatement(oracle, multi):
    """src/clock_gpu_m.f90:105-146,183-216 (clock_gpu_multi_m.f90:215-236 for the strict comparator) on the ring of
    nall sites: ws(up, down, left, right, before, after) with up = s(i + nx), down = s(i - nx), left = s(i - 1),
    right = s(i + 1); nxt = floor(p q); accept iff r <= ws (multi: r < ws)"""
    nx, ny, q, kbt = 33, 32, 6, 0.8
    n = nx * ny
    o = oracle.clock_gpu().init(nx, ny, kbt, q, 5, 1 if multi else None)
    a = 2 * (4 * math.atan(1.0)) / q
    k = np.arange(q)
    etab = -(_cos(a * (k[:, None, None] - k[None, None, :])) + _cos(a * (k[None, :, None] - k[None, None, :])))   # Etab(i, j, c), :115
    I, J, K, L, CB, CA = np.meshgrid(k, k, k, k, k, k, indexing="ij")
    de = (etab[I, J, CA] + etab[K, L, CA]) - (etab[I, J, CB] + etab[K, L, CB])                                   # :127-131
    ws = np.where(de <= 0, 1.0, _exp(-(1 / kbt) * de))
    assert np.array_equal(ws.ravel(order="F"), o.ws)
    rng = np.random.default_rng(6)
    u0 = 1.0 - rng.random(n)
    o.set_random_spin(u0)
    s = np.minimum(np.floor(u0 * q).astype(np.int64), q - 1)                                # :103 (u == 1 clamped)
    for sweep in range(5):
        r, p = 1.0 - rng.random(n), 1.0 - rng.random(n)
        o.update(randoms=r, next_states=p)
        nxt = np.minimum(np.floor(p * q).astype(np.int64), q - 1)
        for colour in (0, 1):                                                               # idx odd (i even) first
            i = np.arange(colour, n, 2)
            w = ws[s[(i + nx) % n], s[(i - nx) % n], s[(i - 1) % n], s[(i + 1) % n], s[i], nxt[i]]
            acc = (r[i] < w) if multi else (r[i] <= w)
            s[i] = np.where(acc, nxt[i], s[i])
        got = o.spins()
        got = got[0] if multi else got
        assert np.array_equal(got[nx:nx + n], s), sweep
    e = etab[s[(np.arange(n) - nx) % n], s[(np.arange(n) - 1) % n], s].sum()                # :258
    eo = o.calc_energy_sum()
    assert abs((eo[0] if multi else eo) - e) < 1e-9


def test_xy_helical_numpy_restatement(oracle):
    """src/xy2d_gpu_m.f90:139-174,176-213,241-291 on the ring of nall sites: dE = -(cand - s) . (s(i-1) + s(i+1) +
    s(i+nx) + s(i-nx)), accept iff r <= exp(-beta dE); over-relaxation s <- 2 (h^ . s) h^ - s WITHOUT renormalisation;
    E = -sum s(i) . (s(i+1) + s(i+nx)), M = sum cos"""
    nx, ny, kbt = 9, 8, 0.9
    n = nx * ny
    o = oracle.xy2d_helical_gpu().init(nx, ny, kbt, 1)
    rng = np.random.default_rng(8)
    th = 2 * np.pi * (1.0 - rng.random(n))
    o.set_angles(th)
    c, s = np.cos(th), np.sin(th)
    i_all = np.arange(n)

    def field(a, i):
        return a[(i - 1) % n] + a[(i + 1) % n] + a[(i + nx) % n] + a[(i - nx) % n]          # :247, same order

    for sweep in range(3):
        r, cand = 1.0 - rng.random(n), 1.0 - rng.random(n)
        o.update(r, cand)
        cc, cs = np.cos(2 * np.pi * cand), np.sin(2 * np.pi * cand)
        for colour in (0, 1):                                                               # offset 1: idx odd = i even
            i = np.arange(colour, n, 2)
            de = -((cc[i] - c[i]) * field(c, i) + (cs[i] - s[i]) * field(s, i))
            acc = r[i] <= np.exp(-(1 / kbt) * de)
            c[i], s[i] = np.where(acc, cc[i], c[i]), np.where(acc, cs[i], s[i])
        sp = o.spins()
        assert np.abs(sp[0, nx:nx + n] - c).max() < 1e-12 and np.abs(sp[1, nx:nx + n] - s).max() < 1e-12, sweep
        o.update_over_relaxation(1)
        for colour in (0, 1):
            i = np.arange(colour, n, 2)
            hx, hy = field(c, i), field(s, i)
            inv = 1 / np.hypot(hx, hy)
            hx, hy = hx * inv, hy * inv
            d = 2 * (hx * c[i] + hy * s[i])
            c[i], s[i] = d * hx - c[i], d * hy - s[i]
        sp = o.spins()
        assert np.abs(sp[0, nx:nx + n] - c).max() < 1e-12 and np.abs(sp[1, nx:nx + n] - s).max() < 1e-12
        e = -(c * (c[(i_all + 1) % n] + c[(i_all + nx) % n])).sum() - (s * (s[(i_all + 1) % n] + s[(i_all + nx) % n])).sum()
        assert abs(o.calc_energy_sum() - e) < 1e-9 and abs(o.calc_magne_sum() - c.sum()) < 1e-9
