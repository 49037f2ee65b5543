"""The drop-in boundary without a GPU: the C-ABI library loads and exports every symbol include/*.h declares,
the .Call glue registers the reference's two routines (src/ExomeDepth_init.c:14-24), and — with no device —
the product path fails loudly instead of falling back to any CPU code."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "exomedepth_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(edb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from exomedepth_b200 import _lib
    L = _lib.load()
    names = declared_symbols()
    assert len(names) >= 17
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names, "exomedepth_b200/_lib.py EXPORTS and include/exomedepth_b200.h disagree"


def test_library_exports_nothing_else():
    from exomedepth_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(declared_symbols()), exported ^ set(declared_symbols())


def test_library_is_sm100a_only_and_has_the_kernels():
    from exomedepth_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    elfs = [l for l in out.stdout.splitlines() if "ELF file" in l]
    assert elfs and all("sm_100a" in l for l in elfs), elfs


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "exomedepth_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle_port" not in txt and "libexomedepth_ref" not in txt, f


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU every compute entry point must fail with ERR_CUDA and a message."""
    code = ("import numpy as np, exomedepth_b200 as e\n"
            "try:\n"
            "    e.get_loglike_matrix([0.01],[0.2],[100],[20],1.0)\n"
            "    print('COMPUTED')\n"
            "except e.EDB200Error as ex:\n"
            "    print('RAISED', ex)\n")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "RAISED" in out.stdout and "no CPU fallback" in out.stdout, out.stdout + out.stderr


def _glue():
    p = os.path.join(ROOT, "oracle", "_ref", "librglue_stub.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/librglue_stub.so not built (make -C oracle glue)")
    return C.CDLL(p)


def test_glue_registers_the_reference_call_entries():
    """R_init_ExomeDepth must register exactly C_hmm(6) and get_loglike_matrix(5) and switch dynamic lookup off."""
    g = _glue()

    class Entry(C.Structure):
        _fields_ = [("name", C.c_char_p), ("fun", C.c_void_p), ("numArgs", C.c_int)]

    class DllInfo(C.Structure):
        _fields_ = [("call_entries", C.POINTER(Entry)), ("use_dynamic_symbols", C.c_int)]

    info = DllInfo(None, 1)
    g.R_init_ExomeDepth.argtypes = [C.POINTER(DllInfo)]
    g.R_init_ExomeDepth.restype = None
    g.R_init_ExomeDepth(C.byref(info))
    got, i = [], 0
    while info.call_entries[i].name:
        got.append((info.call_entries[i].name.decode(), info.call_entries[i].numArgs))
        i += 1
    assert got == [("C_hmm", 6), ("get_loglike_matrix", 5)]
    assert info.use_dynamic_symbols == 0
    for name, _ in got:
        assert C.cast(getattr(g, name), C.c_void_p).value


def test_glue_rejects_other_state_counts_without_touching_the_gpu():
    """hmm.cpp:37-40: nstates != 3 prints and returns NULL — before any device work."""
    from oracle import sexp
    api = sexp.CallApi(_glue())
    api.quiet(True)
    api.rprintf_count(reset=True)
    import numpy as np
    assert api.c_hmm(np.full((5, 5), 0.2), np.zeros((4, 5)), np.arange(4), 1.0) is None
    assert api.rprintf_count() == 1
