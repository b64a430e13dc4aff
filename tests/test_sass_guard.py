"""Build guard (no GPU): the hot loops of the static-table kernels must keep their warp-uniform coefficient operands.

Whether ptxas issues `DFMA R, R, UR, R` (two vector-register reads, full FP64 issue rate) or falls back to `LDC` into
vector registers (three reads, ~0.84 of the rate) depends on it proving the series loop warp-uniform,
and that has broken silently before (run-time transposed-image flags in the orbit kernel).  The check reads the SASS of the
built library with cuobjdump and applies tools/sass_dfma_model.py's operand model to the inner loop of each kernel."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

KERNELS = ["tquKernelILi4ELb1ELi2E", "tquOrbitKernelILi4ELi2ELi0ELb0ELb0E", "tquOrbitKernelILi4ELi2ELi8ELb0ELb0E", "tquOrbitKernelILi4ELi2ELi12ELb0ELb0E",
           "tquOrbitKernelILi4ELi2ELi0ELb1ELb0E", "tquOrbitKernelILi4ELi2ELi0ELb1ELb1E"]


@pytest.fixture(scope="module")
def sass(tmp_path_factory):
    lib = os.path.join(ROOT, "cosmopp_b200", "lib", "libcosmopp_b200.so")
    if not os.path.exists(lib) or not shutil.which("cuobjdump"):
        pytest.skip("library not built or cuobjdump missing")
    path = tmp_path_factory.mktemp("sass") / "lib.sass"
    with open(path, "w") as f:
        subprocess.run(["cuobjdump", "-sass", lib], stdout=f, check=True)
    return str(path)


@pytest.mark.parametrize("kernel", KERNELS)
def test_series_loop_reads_two_vector_registers_per_dfma(sass, kernel):
    import sass_dfma_model as m
    ins = m.parse(sass, kernel)
    assert ins, "kernel %s not found in the library" % kernel
    loops = m.loops(ins)
    inner = [lp for lp in loops if lp["dfma"] >= 600 and lp["dfma"] <= 700]
    assert inner, "no 16-step series loop (640 DFMA) found in %s" % kernel
    lp = inner[0]
    assert lp["hist"].get(3, 0) == 0, "%s: %d DFMAs of the series loop read three vector registers" % (kernel, lp["hist"][3])
    assert lp["other"] <= 140
