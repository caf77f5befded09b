"""CPU: the oracle (and the RNG contract) reproduces the committed golden vectors
(tests/golden/trajectories.json, written by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "trajectories.json")) as f:
        return json.load(f)


def test_oracle_reproduces_golden(oracle):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want, got = _golden(), mg.generate()
    assert got["ising"] == want["ising"]                      # integers + hashes: exact
    assert got["ising_torus"] == want["ising_torus"]
    for key in ("clock", "sixclock"):
        for a, b in zip(got[key], want[key]):
            assert a["hist"] == b["hist"]
            assert a.get("spins_sha256") == b.get("spins_sha256") and a.get("states_sha256") == b.get("states_sha256")
            assert np.allclose(a["energy"], b["energy"], rtol=1e-13, atol=1e-9)
            assert np.allclose(a["magne"], b["magne"], rtol=1e-13, atol=1e-9)
    for a, b in zip(got["xy"], want["xy"]):
        for k in ("start", "after_metropolis", "after_over_relaxation"):
            assert np.allclose(a[k], b[k], rtol=1e-12, atol=1e-9)
    assert got["rng"] == want["rng"]                          # uniforms: bit-exact doubles


def test_golden_uniforms_are_32bit_in_unit_interval():
    for name, v in _golden()["rng"].items():
        a = np.asarray(v, dtype=np.float64).ravel()
        assert a.min() > 0.0 and a.max() <= 1.0, name
        assert np.all(a * 2.0 ** 32 == np.round(a * 2.0 ** 32)), name   # u = (U + 1) 2^-32
