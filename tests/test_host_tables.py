"""The native host-side logic (exomedepth_b200/csrc/host_tables.cpp: sweep placement, CallCNVs transition matrix,
position framing, host-libm log-transition rows) checked on the CPU by a small C++ driver compiled here."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "exomedepth_b200", "csrc")


def test_host_tables_native(tmp_path):
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "host_tables_check")
    # the flags of csrc/Makefile for this file: no FMA contraction, so log / exp arguments carry the reference's bits
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-I", CSRC,
                    os.path.join(ROOT, "tests", "native", "host_tables_check.cpp"), os.path.join(CSRC, "host_tables.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


def test_structured_viterbi_step_native(tmp_path):
    """viterbi_step.h (the one-thread-per-chain sweep's step for CallCNVs-structured transition rows) against the plain
    scan of src/hmm.cpp:66-88, bit for bit, on adversarial inputs; and the structured rows against the general table."""
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "viterbi_step_check")
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-I", CSRC,
                    os.path.join(ROOT, "tests", "native", "viterbi_step_check.cpp"), os.path.join(CSRC, "host_tables.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-3000:] + r.stderr[-2000:]


def test_segmented_sweep_native(tmp_path):
    """The segmented Viterbi sweep's logic on the CPU (tests/native/segments_check.cpp): the piece cutter of host_tables.cpp,
    the lead-reporting steps of viterbi_step.h against the plain scan of src/hmm.cpp:66-88, and the certification scheme of
    viterbi_seam.h end to end — every certified decision of chains swept as pieces equals the sequential sweep's."""
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "segments_check")
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-I", CSRC,
                    os.path.join(ROOT, "tests", "native", "segments_check.cpp"), os.path.join(CSRC, "host_tables.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-3000:] + r.stderr[-2000:]
