"""CPU: the C++ drop-in headers (vectorizedadjoint_b200/include) compile, the tape recorder identifies the built-in
systems, and the example clients build -- including the reference's unmodified example sources where available."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vectorizedadjoint_b200")


def _ensure_lib():
    import vectorizedadjoint_b200 as va
    if not os.path.exists(va.LIB_PATH):
        va.build()


def test_tape_recorder_and_driver_surface():
    _ensure_lib()
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "tape_check")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(PKG, "include"), "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "tape_check.cpp"), "-o", exe, "-L", PKG, "-lva_engine", f"-Wl,-rpath,{PKG}"])
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "tape ok" in out.stdout
    assert "Must call setCostGradients() first!" in out.stdout  # reference lib/include/backpropagation.hpp:22-26


def test_example_clients_build():
    _ensure_lib()
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "examples")])
    for name in ("harmonic", "vanderpol", "lotka"):
        assert os.path.exists(os.path.join(PKG, "examples", "build", name))
    if os.path.isdir("/root/reference/examples"):
        for name in ("ref_harmonic", "ref_vanderpol", "ref_lotka"):  # unmodified reference sources against these headers
            assert os.path.exists(os.path.join(PKG, "examples", "build", name))
