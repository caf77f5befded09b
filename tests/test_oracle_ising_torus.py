"""CPU: the periodic (torus) Ising restatement of oracle/oracle.c (orc_isingp_*) against an independent vectorised numpy
restatement (np.roll neighbours, boolean colour masks) written from the reference's update rule
(src/ising3d_gpu_m.f90:189-206,239-276; src/ising2d_gpu_m.f90:148-162,191-228): bit-identical trajectories and
observables on shared uniforms, known answers, and the RNG contract's (vector, lane) addressing."""
import numpy as np
import pytest


def _np_tables(ndim, beta):
    if ndim == 3:   # ws(S, s) = min(1, exp(-beta dE)), dE = 2 sigma (2S - 6)  (src/ising3d_gpu_m.f90:153-170)
        w = np.empty((2, 7))
        for s in (0, 1):
            for S in range(7):
                de = 2 * (2 * s - 1) * (2 * S - 6)
                w[s, S] = min(1.0, np.exp(-beta * float(de)))
        return w
    ex = np.ones(17)  # exparr(-8:8) (src/ising2d_gpu_m.f90:126-130)
    for d in range(1, 9):
        ex[d + 8] = np.exp(-beta * d)
    return ex


def _np_sweep(s, u, beta, method):
    """s: int array [nz][ny][nx] (3D, 0/1) or [ny][nx] (2D, -1/+1); u same shape"""
    ndim = s.ndim
    idx = np.indices(s.shape).sum(axis=0)
    for colour in (0, 1):
        nsum = sum(np.roll(s, sh, axis=ax) for ax in range(ndim) for sh in (1, -1))
        mask = (idx & 1) == colour
        if ndim == 3:
            if method == 0:
                w = _np_tables(3, beta)
                acc = u <= w[s, nsum]
                s = np.where(mask & acc, 1 - s, s)
            else:
                pup = 1.0 / (1.0 + np.exp(-2.0 * beta * (2 * nsum - 6).astype(np.float64)))
                s = np.where(mask, (u <= pup).astype(s.dtype), s)
        else:
            if method == 0:
                ex = _np_tables(2, beta)
                acc = u <= ex[2 * s * nsum + 8]
                s = np.where(mask & acc, -s, s)
            else:
                S = (nsum + 4) // 2
                pup = 1.0 / (1.0 + np.exp(-2.0 * beta * (2 * S - 4).astype(np.float64)))
                s = np.where(mask, np.where(u <= pup, 1, -1).astype(s.dtype), s)
    return s


def _np_em(s):
    ndim = s.ndim
    sig = 2 * s - 1 if ndim == 3 else s
    e = -sum((sig * np.roll(sig, -1, axis=ax)).sum() for ax in range(ndim))
    return int(e), int(sig.sum())


@pytest.mark.parametrize("shape", [(32, 4, 2), (64, 6, 4), (32, 2, 2), (64, 10, 0), (32, 2, 0)])
@pytest.mark.parametrize("method", [0, 1])
def test_oracle_torus_equals_numpy_restatement(oracle, shape, method):
    nx, ny, nz = shape
    kbt = 4.51152 if nz else 2.26918531421
    o = oracle.ising_periodic_gpu().init(nx, ny, nz, kbt, 11)
    o.set_random_spin()
    dims = (nz, ny, nx) if nz else (ny, nx)
    s = o.spins().reshape(dims).copy()
    rng = np.random.default_rng(3)
    tabvals = np.unique(np.concatenate([o.w, o.pup, [1.0]]))
    tabvals = tabvals[(tabvals > 0) & (tabvals <= 1)]
    for sweep in range(4):
        u = 1.0 - rng.random(s.size)
        pick = rng.random(s.size) < 0.25
        u[pick] = rng.choice(tabvals, size=int(pick.sum()))     # uniforms equal to table entries: the <= / > boundary
        (o.update_heatbath if method else o.update)(u)
        s = _np_sweep(s, u.reshape(dims), o.beta_, method)
        assert np.array_equal(o.spins().reshape(dims), s), sweep
        assert o.measure() == _np_em(s)


def test_oracle_torus_known_answers(oracle):
    o = oracle.ising_periodic_gpu().init(64, 6, 4, 4.5, 1)
    n = o.nall()
    assert o.measure() == (-3 * n, n)                      # all up: every bond aligned
    o.s[:] = 0
    assert o.measure() == (-3 * n, -n)
    x, y, z = np.meshgrid(np.arange(64), np.arange(6), np.arange(4), indexing="ij")
    o.s[:] = ((x + y + z) & 1).transpose(2, 1, 0).ravel()  # perfect antiferromagnet: every bond broken
    assert o.measure() == (3 * n, 0)
    o2 = oracle.ising_periodic_gpu().init(64, 6, 0, 2.2, 1)
    assert o2.measure() == (-2 * o2.nall(), o2.nall())
    # beta -> infinity from all up: nothing flips (dE > 0 everywhere, u in (0, 1]); beta = 0: every proposal accepted
    o3 = oracle.ising_periodic_gpu().init(64, 6, 4, 1e-9, 1)
    before = o3.spins()
    o3.update()
    assert np.array_equal(o3.spins(), before)
    o4 = oracle.ising_periodic_gpu().init(64, 6, 4, 1e12, 1)
    o4.update()
    assert np.array_equal(o4.spins(), 1 - before)


def test_torus_uniform_contract_addressing(oracle):
    """u(site) = the ring contract's uniform at (vector, lane) = (k / 16, k % 16), k the colour-compact row-major index:
    checked against the raw Philox function for a few sites"""
    import ctypes as C
    nx, ny, nz, seed, draw = 64, 6, 4, 42, 5
    u = oracle.isingp_uniforms(seed, draw, nx, ny, nz)
    assert u.shape == (nx * ny * nz,) and (u > 0).all() and (u <= 1).all()
    BYTEPOS = [0, 2, 4, 6, 1, 3, 5, 7, 8, 10, 12, 14, 9, 11, 13, 15]
    lib = oracle.lib()
    for (x, y, z) in [(0, 0, 0), (1, 0, 0), (33, 5, 3), (62, 2, 1), (63, 5, 3)]:
        colour = (x + y + z) & 1
        k = (z * ny + y) * (nx // 2) + (x >> 1)
        v, lane = k // 16, k % 16
        m = BYTEPOS[lane]
        key = (C.c_uint32 * 2)(seed, 0x49534E47)
        out = (C.c_uint32 * 4)()
        lib.orc_philox((C.c_uint32 * 4)(v, 0, draw, colour << 16), key, out)
        b7 = (out[m >> 2] >> (8 * (m & 3))) & 0x7F
        lib.orc_philox((C.c_uint32 * 4)(v, 0, draw, (colour << 16) | ((1 + (m >> 2)) << 24)), key, out)
        U = (b7 << 25) | (out[m & 3] & 0x1FFFFFF)
        assert u[x + nx * (y + ny * z)] == (U + 1) * 2.0 ** -32
