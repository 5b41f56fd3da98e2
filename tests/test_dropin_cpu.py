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
        gen = os.path.join(d, "zoo_generated.h")
        out = subprocess.run([exe, gen], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "tape ok" in out.stdout
        assert "Must call setCostGradients() first!" in out.stdout  # reference lib/include/backpropagation.hpp:22-26
        # The generated rhs / vjp source of a functor that uses every recorded operation (elementary functions, comparisons,
        # iIf) is plain C++ apart from __device__: compile it on the host and check the vjp against central finite differences
        # of the rhs, on both sides of the recorded selects.
        drv = os.path.join(d, "zoo_check.cpp")
        open(drv, "w").write(r"""
#include <cmath>
#include <cstdio>
#define __device__
#include "zoo_generated.h"
int main() {
    const double pts[2][3] = {{0.31, 0.44, 0.52}, {0.95, 0.1, 0.85}};
    const double p0[2] = {0.8, 0.6}, w[3] = {0.7, -1.3, 0.45}, t = 0.25, h = 1e-6;
    double worst = 0.0;
    for (int c = 0; c < 2; ++c) {
        double x[3] = {pts[c][0], pts[c][1], pts[c][2]}, p[2] = {p0[0], p0[1]}, gx[3] = {0, 0, 0}, gp[2] = {0, 0};
        VaUserSys::vjp(x, p, t, w, gx, gp);
        for (int k = 0; k < 5; ++k) {
            double *q = k < 3 ? &x[k] : &p[k - 3];
            const double keep = *q;
            double fp[3], fm[3];
            *q = keep + h; VaUserSys::rhs(x, p, t, fp);
            *q = keep - h; VaUserSys::rhs(x, p, t, fm);
            *q = keep;
            double fd = 0.0;
            for (int i = 0; i < 3; ++i) fd += w[i] * (fp[i] - fm[i]) / (2 * h);
            const double an = k < 3 ? gx[k] : gp[k - 3];
            const double err = std::fabs(fd - an) / (std::fabs(fd) + 1e-3);
            if (err > worst) worst = err;
            std::printf("case %d var %d: vjp %.12g fd %.12g\n", c, k, an, fd);
        }
    }
    std::printf("worst %.3e\n", worst);
    return worst < 1e-7 ? 0 : 1;
}
""")
        zexe = os.path.join(d, "zoo_check")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", d, drv, "-o", zexe])
        z = subprocess.run([zexe], capture_output=True, text=True)
        assert z.returncode == 0, z.stdout + z.stderr


def test_example_clients_build():
    _ensure_lib()
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "examples")])
    for name in ("harmonic", "vanderpol", "lotka", "pendulum", "switched", "batch_lotka"):
        assert os.path.exists(os.path.join(PKG, "examples", "build", name))
    if os.path.isdir("/root/reference/examples"):
        for name in ("ref_harmonic", "ref_vanderpol", "ref_lotka"):  # unmodified reference sources against these headers
            assert os.path.exists(os.path.join(PKG, "examples", "build", name))
