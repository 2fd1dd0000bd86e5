"""CPU suite: the rasteriser's exact row-span solver (drtk_b200/csrc/raster_core.cuh) selects exactly the samples the
reference's per-sample test selects.  The header is compiled for the host (g++, hardware FMA, FTZ/DAZ like the CUDA
build's --use_fast_math) together with tests/span_harness.cpp and run on random, degenerate and knife-edge triangles
with the reciprocal perturbed by up to +-2 ulp (MUFU.RCP is an approximation)."""
import os
import shutil
import subprocess

import pytest

from tests.util import ROOT

CSRC = os.path.join(ROOT, "drtk_b200", "csrc")


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return False


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None or not _cpu_has_fma():
        pytest.skip("needs g++ and a host CPU with FMA")
    exe = str(tmp_path_factory.mktemp("span") / "span_harness")
    subprocess.check_call([gxx, "-O2", "-mfma", "-ffp-contract=off", "-I", CSRC,
                           os.path.join(ROOT, "tests", "span_harness.cpp"), "-o", exe, "-lm"])
    return exe


@pytest.mark.parametrize("ntri,width,seed", [(600000, 2048, 1), (400000, 64, 2), (400000, 65000, 3)])
def test_exact_spans_equal_per_sample_test(harness, ntri, width, seed):
    out = subprocess.run([harness, str(ntri), str(width), str(seed)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = dict(kv.split("=") for kv in out.stdout.split())
    assert int(fields["rows"]) > 1_000_000 and int(fields["covered"]) > int(fields["rows"])
    assert int(fields["mismatches"]) == 0
