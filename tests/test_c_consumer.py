"""The C ABI from plain C: tests/c_consumer/relaxation_loop.c is compiled with gcc -std=c11 -Wall -Wextra -pedantic -Werror
against include/b200mc.h (nothing but nvcc had included that header before) and linked with libb200mc.so.
CPU: it builds, and without a device it fails loudly.  GPU: its E / M series equal the oracle's."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda_fortran_mc_simulation_spin_b200")
SRC = os.path.join(ROOT, "tests", "c_consumer", "relaxation_loop.c")


def _build(tmp_path):
    exe = str(tmp_path / "relaxation_loop")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", PKG, "-lb200mc", "-Wl,-rpath," + PKG]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_links(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run is covered by the gpu test")
    r = subprocess.run([exe, "3", "31", "31", "30", "4.51152", "42", "2", "allup"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,shape,kbt,start", [(3, (31, 31, 30), 4.51152, "allup"), (3, (63, 65, 64), 4.51152, "random"),
                                                 (2, (1001, 1000), 2.26918531421, "allup"), (2, (255, 256), 2.26918531421, "random")])
def test_c_driver_loop_equals_oracle(oracle, tmp_path, dim, shape, kbt, start):
    exe = _build(tmp_path)
    mcs = 6
    nx, ny = shape[0], shape[1]
    nz = shape[2] if dim == 3 else 0
    r = subprocess.run([exe, str(dim), str(nx), str(ny), str(nz), repr(kbt), "42", str(mcs), start], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == mcs + 1 and lines[-1].startswith("launches ") and int(lines[-1].split()[1]) > 0
    o = (oracle.ising3d_gpu() if dim == 3 else oracle.ising2d_gpu()).init(*shape, kbt, 42)
    if start == "random":
        o.set_random_spin()
    for i in range(mcs):
        o.update()
        assert [int(x) for x in lines[i].split()] == [i + 1, o.calc_energy_sum(), o.calc_magne_sum()]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,kbt,start", [((1024, 8, 4), 4.51152, "random"), ((64, 6, 4), 4.51152, "allup"), ((96, 10, 0), 2.26918531421, "random")])
def test_c_driver_loop_on_the_torus_equals_oracle(oracle, tmp_path, shape, kbt, start):
    exe = _build(tmp_path)
    mcs = 6
    dim = -3 if shape[2] else -2
    r = subprocess.run([exe, str(dim), str(shape[0]), str(shape[1]), str(shape[2]), repr(kbt), "42", str(mcs), start], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == mcs + 1 and lines[-1].startswith("launches ")
    o = oracle.ising_periodic_gpu().init(*shape, kbt, 42)
    if start == "random":
        o.set_random_spin()
    for i in range(mcs):
        o.update()
        assert [int(x) for x in lines[i].split()] == [i + 1, *o.measure()]
