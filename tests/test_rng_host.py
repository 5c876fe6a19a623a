"""CPU tests of the branch-free Box-Muller of K_A (orphics_b200/csrc/ox_rng.cuh): its constants are exactly what
tools/gen_rng_tables.py generates, and a host mirror of the device arithmetic (tools/ubench/rng_host_check.cpp)
reproduces the definition of the noise, sqrt(-2 ln u1) e^{2 pi i u2}, to 5e-15 over random and edge-case inputs
(5e-6 for the single-precision variant of the float32 pipeline)."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "orphics_b200", "csrc")


def test_tables_are_reproducible(tmp_path):
    pytest.importorskip("mpmath")
    work = tmp_path / "r"
    (work / "tools").mkdir(parents=True)
    (work / "orphics_b200" / "csrc").mkdir(parents=True)
    shutil.copy(os.path.join(ROOT, "tools", "gen_rng_tables.py"), work / "tools" / "gen_rng_tables.py")
    subprocess.run([sys.executable, str(work / "tools" / "gen_rng_tables.py")], check=True, capture_output=True)
    assert (work / "orphics_b200" / "csrc" / "ox_rng_tables.h").read_text() == open(os.path.join(CSRC, "ox_rng_tables.h")).read()


def test_host_mirror_matches_libm(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("g++ unavailable")
    exe = str(tmp_path / "rng_host_check")
    subprocess.run(["g++", "-O2", "-I" + CSRC, os.path.join(ROOT, "tools", "ubench", "rng_host_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "2000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "max abs err of normals" in out.stdout and "float32 variant" in out.stdout
