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


def test_batch_struct_layout_matches_the_header(tmp_path):
    """exomedepth_b200/_lib.py:Batch is a hand-written ctypes image of include/exomedepth_b200.h:edb200_batch — size and
    every field offset must be what the C compiler lays out."""
    from exomedepth_b200 import _lib
    fields = [n for n, _ in _lib.Batch._fields_]
    src = tmp_path / "layout.c"
    lines = "\n".join(f'    printf("{n} %zu\\n", offsetof(edb200_batch, {n}));' for n in fields)
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "exomedepth_b200.h"\nint main(void) {\n'
                   '    printf("sizeof %zu\\n", sizeof(edb200_batch));\n' + lines + "\n    return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out.pop("sizeof")) == C.sizeof(_lib.Batch)
    for n in fields:
        assert int(out[n]) == getattr(_lib.Batch, n).offset, n
    # and the header has no field the ctypes image lacks
    body = re.search(r"typedef struct edb200_batch \{(.*?)\} edb200_batch;", re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S), flags=re.S).group(1)
    declared = re.findall(r"\b([a-z_0-9]+)\s*;", body)
    assert declared == fields, (declared, fields)


def test_ingestion_layouts_round_trip():
    """pack_counts (edb200_batch.observed16) and pack_counts12 (observed12): decoding the packed rows as the header
    describes them and applying the overflow list gives the counts back; sentinels themselves go to the list; odd bin
    counts and row strides as documented."""
    import numpy as np

    from exomedepth_b200 import pack_counts, pack_counts12
    rng = np.random.default_rng(5)
    for nb in (1, 2, 7, 8, 4097, 4098):
        obs = rng.integers(0, 5000, (5, nb)).astype(np.int32)
        obs[rng.integers(5), rng.integers(nb)] = 70000
        obs[0, 0], obs[4, nb - 1] = 4095, 65535
        u16, i16, v16 = pack_counts(obs)
        dec = u16.astype(np.int32)
        assert np.all(dec.ravel()[i16] == 65535) and np.all(v16 >= 65535)
        dec.ravel()[i16] = v16
        assert np.array_equal(dec, obs)
        u8, i12, v12 = pack_counts12(obs)
        pairs = (nb + 1) // 2
        assert u8.dtype == np.uint8 and u8.shape == (5, (3 * pairs + 3) // 4 * 4)
        trip = u8[:, :3 * pairs].reshape(5, pairs, 3).astype(np.int32)
        word = trip[:, :, 0] | trip[:, :, 1] << 8 | trip[:, :, 2] << 16               # v0 | v1 << 12
        dec = np.stack([word & 0xFFF, word >> 12], axis=2).reshape(5, 2 * pairs)[:, :nb].copy()
        assert np.array_equal(np.flatnonzero(dec.ravel() == 4095), i12) and np.all(v12 >= 4095)
        dec.ravel()[i12] = v12
        assert np.array_equal(dec, obs)
    with pytest.raises(ValueError):
        pack_counts12(np.array([[1, -1]]))
    with pytest.raises(ValueError):
        pack_counts(np.array([[1, -1]], np.int32))


def test_native_encoders_match_the_numpy_definition():
    """edb200_pack_counts16 / edb200_pack_counts12 (host threads, no GPU) against cohort.pack_counts_numpy / pack_counts12_numpy:
    a matrix large enough for several threads, more overflow entries than the first guess of the list capacity, a strided
    input, int64 input, a pinned-style preallocated output, and the C entry points' argument checks."""
    import numpy as np

    from exomedepth_b200 import _lib, pack_counts, pack_counts12
    from exomedepth_b200.cohort import pack_counts12_numpy, pack_counts_numpy
    rng = np.random.default_rng(9)
    big = rng.integers(0, 4200, (37, 40001)).astype(np.int32)                 # ~2.5 % at or beyond 4095: > 65,536 entries? no: ~37k
    big[:, ::3] = rng.integers(4095, 90000, big[:, ::3].shape)                 # a third of the bins overflow: ~493k entries
    wide = np.zeros((37, 40100), np.int32)
    wide[:, :40001] = big
    for obs in (big, wide[:, :40001], big.astype(np.int64), big[:1, :1], big[:3, :2]):
        for native, ref in ((pack_counts, pack_counts_numpy), (pack_counts12, pack_counts12_numpy)):
            got, want = native(obs), ref(obs)
            for g, w in zip(got, want):
                assert g.dtype == w.dtype and np.array_equal(g, w), (native.__name__, obs.shape)
    assert pack_counts12(big)[1].size > (1 << 16)
    out = np.full((37, (40001 + 1) // 2 * 3 + 3 & ~3), 0xEE, np.uint8)
    u8, idx, val = pack_counts12(big, out=out)
    assert u8 is out and np.array_equal(u8[:, :60003], pack_counts12_numpy(big)[0][:, :60003])
    L = _lib.load()
    i64, i32 = np.zeros(4, np.int64), np.zeros(4, np.int32)
    small = np.arange(6, dtype=np.int32).reshape(2, 3)
    o8 = np.zeros((2, 8), np.uint8)
    assert L.edb200_pack_counts12(small.ctypes.data, 3, 2, 3, o8.ctypes.data, 8, i64.ctypes.data, i32.ctypes.data, 4) == 0
    assert L.edb200_pack_counts12(small.ctypes.data, 3, 2, 3, o8.ctypes.data, 5, i64.ctypes.data, i32.ctypes.data, 4) == -1      # row of 6 bytes
    assert L.edb200_pack_counts12(small.ctypes.data, 2, 2, 3, o8.ctypes.data, 8, i64.ctypes.data, i32.ctypes.data, 4) == -1      # stride < n_bins
    assert L.edb200_pack_counts16(None, 3, 2, 3, o8.ctypes.data, 3, i64.ctypes.data, i32.ctypes.data, 4) == -1
    assert L.edb200_pack_counts16(small.ctypes.data, 3, 0, 3, None, 3, None, None, 0) == 0
    small[1, 1] = 70000
    assert L.edb200_pack_counts16(small.ctypes.data, 3, 2, 3, o8.ctypes.data, 3, None, None, 0) == 1           # counted, nothing written
