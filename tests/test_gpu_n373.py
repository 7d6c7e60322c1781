"""Small-body kernels with more targets than sb441-n16 (sb441-n373, reference
docs/installation.md:44-52).  The first 16 targets live in the per-time body tables as before; the others act through
the direct term only (reference src/forces.c:281-344: asteroids first, in file order, then the planets) and are
evaluated inside it.  Against tests/golden/golden_n373.npz = the reference's own C code on the same files
(write_extended: N = 40 and N = 373 targets), through the C-ABI."""
import os

import numpy as np
import pytest

import cases
from conftest import ROOT
from assist_b200 import batch as ab
from assist_b200.synth import ephem_writer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden_n():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_n373.npz"))


@pytest.fixture(scope="module", params=cases.N373_SIZES)
def ephn(request):
    p = ephem_writer.write_extended(os.path.join(ROOT, "data"), request.param)
    return request.param, ab.EphemHandle(p["planets_bsp"], p["asteroids_bsp"])


def test_ephemeris_of_every_target(ephn, golden_n):
    N, eph = ephn
    g = golden_n["eph%d" % N]
    got, st = eph.eval(cases.n373_times())
    assert got.shape == (5, 11 + N, 10) and (st == 0).all()
    assert np.array_equal(got[:, :11, :], g[:, :11, :])
    assert np.array_equal(got[:, 11:, :4], g[:, 11:, :4])          # GM, x, y, z (the reference gives no velocities)
    assert (got[:, 11:, 0] > 0).all()


def test_force_evaluation_sees_every_asteroid(ephn, golden_n):
    """Direct term of 11 + N bodies on real and variational particles: bit-identical in strict math."""
    N, eph = ephn
    st = cases.n373_force_systems()
    acc = ab.eval_forces(eph, cases.T0 + 17.25, st, forces=0x7F)
    assert np.array_equal(acc, golden_n["acc%d" % N])
    # and it matters: the same evaluation on the 16-target file differs
    p16 = ephem_writer.write_all(os.path.join(ROOT, "data"))
    e16 = ab.EphemHandle(p16["planets_bsp"], p16["asteroids_bsp"])
    assert not np.array_equal(ab.eval_forces(e16, cases.T0 + 17.25, st, forces=0x7F), acc)


def test_per_particle_integration(ephn, golden_n):
    N, eph = ephn
    if N != cases.N373_SIZES[0]:
        pytest.skip("integrations are pinned at N = %d" % cases.N373_SIZES[0])
    st = cases.n373_pp_particles()
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + 400.0)
    got = b.get_state()
    s = b.stats()
    b.close()
    assert np.array_equal(got["state"], golden_n["pp_final"])
    assert np.array_equal(got["t"], golden_n["pp_t"]) and np.array_equal(got["dt"], golden_n["pp_dt"])
    assert [s["steps"], s["pc_iterations"], s["force_evals"], s["steps_rejected"]] == list(golden_n["pp_counts"])


def test_shared_step_integration_with_variational_particles(ephn, golden_n):
    N, eph = ephn
    if N != cases.N373_SIZES[0]:
        pytest.skip("integrations are pinned at N = %d" % cases.N373_SIZES[0])
    st = cases.n373_shared_systems()
    b = ab.Batch(eph, st.shape[0], 1, ab.SHARED_STEP, forces=0x7F)
    b.set_state(cases.T0, st)
    b.integrate(cases.T0 + 300.0)
    got = b.get_state()
    s = b.stats()
    b.close()
    assert np.array_equal(got["state"], golden_n["sh_final"])
    assert (got["t"][0], got["dt"][0], got["dt_last_done"][0]) == tuple(golden_n["sh_t_dt"])
    assert s["steps"] == golden_n["sh_counts"][0] and s["pc_iterations"] == golden_n["sh_counts"][1]
