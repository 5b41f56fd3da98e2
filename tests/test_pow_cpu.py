"""CPU: the controller's pow() (vectorizedadjoint_b200/csrc/va_pow.h, glibc's algorithm + tables restated for host and
device) must equal the system pow() bit for bit -- the reference's accept/reject sequence depends on it."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_va_pow_matches_system_pow_bitwise():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "pow_check")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "vectorizedadjoint_b200", "csrc"),
                               os.path.join(ROOT, "tests", "pow_check.c"), "-o", exe, "-lm"])
        out = subprocess.run([exe, "2000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "0 mismatches" in out.stdout


def test_pow_tables_match_this_libm(tmp_path):
    """The committed tables are the ones of the libm in this image (regenerate with tools/gen_pow_tables.py otherwise)."""
    out = tmp_path / "t.h"
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "gen_pow_tables.py"), "/lib/x86_64-linux-gnu/libm.so.6", str(out)],
                          stdout=subprocess.DEVNULL)
    assert out.read_text() == open(os.path.join(ROOT, "vectorizedadjoint_b200", "csrc", "va_pow_tables.h")).read()
