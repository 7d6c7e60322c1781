"""The reference's own C programs (unit_tests/*/problem.c), compiled UNMODIFIED against include/ and linked with the
product library in the build container (tests/ref_programs.py -> oracle/_ref/programs), run here on the GPU with the
synthetic ephemeris files standing in for the JPL ones (they look for "../../data/de440.bsp" etc.).  Only the
data-independent programs -- invariants (cache on/off, round trips, interpolation vs integration, variational vs
finite difference) and format handling -- can pass without the real files; the ones that compare with JPL Horizons
values need de440 itself and are not run."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

PROGRAMS = os.path.join(ROOT, "oracle", "_ref", "programs")
RUNNABLE = ["ephem_cache", "roundtrip_adaptive_spk", "roundtrip_adaptive_ascii", "onthefly_interpolation",
            "onthefly_backwards_interpolation", "variational_spk", "variational_ascii", "format_detection", "ascii_reject",
            "ascii_init_constants", "spk_join_masses", "spk_load_constants", "spk_planets_calc"]
# Not runnable on the synthetic files although they hold no Horizons value: roundtrip_spk / roundtrip_ascii integrate
# past JD 2465000.5, the end of the synthetic coverage (the real DE440 runs to 2650); spk_detection / spk_init assert
# the 14 targets of the real de440.bsp (the synthetic kernel holds the 12 that ASSIST reads).


@pytest.fixture(scope="module")
def rundir(paths, tmp_path_factory):
    base = tmp_path_factory.mktemp("refrun")
    data = base / "data"
    data.mkdir()
    os.symlink(paths["planets_bsp"], data / "de440.bsp")
    os.symlink(paths["asteroids_bsp"], data / "sb441-n16.bsp")
    os.symlink(paths["de440"], data / "linux_p1550p2650.440")
    cwd = base / "unit_tests" / "run"
    cwd.mkdir(parents=True)
    return str(cwd)


@pytest.mark.parametrize("name", RUNNABLE)
def test_reference_program_passes_on_the_gpu_library(name, rundir, lib):
    exe = os.path.join(PROGRAMS, name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/programs/%s not built (the build container compiles it from /root/reference)" % name)
    if lib.assist_gpu_device_count() < 1:
        pytest.fail("GPU tests need a CUDA device: assist-b200 has no CPU path")
    p = subprocess.run([exe], cwd=rundir, capture_output=True, text=True, timeout=180)
    assert p.returncode == 0, "%s exited with %d\n%s\n%s" % (name, p.returncode, p.stdout[-2000:], p.stderr[-2000:])
