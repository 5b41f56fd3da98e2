"""CPU: the C-ABI library builds, loads and exports every symbol include/va_engine.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def va():
    import vectorizedadjoint_b200 as va
    if not os.path.exists(va.LIB_PATH):
        va.build()
    return va


def test_header_symbols_exported(va):
    hdr = open(os.path.join(ROOT, "include", "va_engine.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+)?(?:int|void|char)\s*\*?\s*(va_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert {"va_engine_create", "va_forward_adjoint_batch", "va_forward_batch", "va_adjoint_batch", "va_get_checkpoints",
            "va_engine_destroy", "va_last_error"} <= declared
    L = ctypes.CDLL(va.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/va_engine.h but not exported"
    assert set(va.EXPORTS) == declared


def test_struct_layouts_match_header(va):
    # sizes the C compiler gives the argument structs must equal the ctypes mirrors
    import subprocess
    import tempfile
    src = ('#include "va_engine.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(va_engine_desc), '
           'sizeof(va_batch_args), sizeof(va_engine_info));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(va._Desc), ctypes.sizeof(va._Args), ctypes.sizeof(va._Info)]


def test_no_cpu_fallback(va):
    """Without a CUDA device engine creation must fail loudly, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(va.EngineError):
        va.Engine(va.SYS_HARMONIC, 2, va.RK_RK4, False)


def test_product_never_touches_oracle():
    """The product package must not import, link or reference anything under oracle/."""
    pkg = os.path.join(ROOT, "vectorizedadjoint_b200")
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "va_oracle.h" not in txt and "libva_ref" not in txt and "import oracle" not in txt, os.path.join(base, f)


def test_torch_front_end_imports_and_fails_loudly_without_gpu():
    """The PyTorch front-end is plumbing over the C-ABI: importable anywhere, but without a CUDA device it must refuse to
    construct a solver (no CPU path)."""
    import torch
    import vectorizedadjoint_b200 as va
    from vectorizedadjoint_b200 import torch_api
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(va.EngineError):
        torch_api.OdeSolver(va.SYS_HARMONIC, 2, va.RK_RK4, False, ti=0.0, tf=1.0, dt0=0.01)


def test_headline_kernel_compiles_without_spills(va):
    """The GLV N = 64 kernel (va_glv_t8.cu, cash_karp54, adaptive, exact species count) is register-tuned to 252 registers and
    no spills; unrelated edits have broken that before (new fields in front of `coef` in VaGlvWideArgs moved constant-bank
    offsets: 104 bytes of spills, -4 % throughput). The ptxas log of the in-tree build is the witness."""
    log = os.path.join(ROOT, "vectorizedadjoint_b200", "csrc", "build", "va_glv_t8.ptxas.log")
    if not os.path.exists(log):
        va.build()
    if not os.path.exists(log):
        pytest.skip("no in-tree build log (library built elsewhere)")
    text = open(log).read()
    found = re.findall(r"Function properties for \S*k_glv_t8INS_7TabCK54ELb1ELb1ELb([01])E\S*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads",
                       text)
    assert {f[0] for f in found} == {"0", "1"}, "headline instantiations (parameters staged in shared memory / read from global) not found in the ptxas log"
    for staged, _, st, ld in found:
        assert (int(st), int(ld)) == (0, 0), f"k_glv_t8<TabCK54, adaptive, exact, staged={staged}> spills: {st} B stores, {ld} B loads"
