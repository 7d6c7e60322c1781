"""Parity at the sizes BASELINE.json's north_star names (VERDICT r1 task 3), through the C-ABI, against
tests/golden/golden_large.npz = outputs of the reference's own C code (tests/golden/make_golden_large.py):

  C3  the fixed 1000-particle subsample of the 10^6 NEO+MBA bench population, 3652.5 d, min_dt 1e-3
  C4  256 systems x (1 + 6 variational), 1826.25 d
  C2  ONE shared-step simulation of 10^4 main-belt particles, 3652.5 d
  C5  256 comets with Marsden A1..A3, assist_integrate_or_interpolate at 1827 epochs, 50 yr backward
  a geocentric integration, and a batch with gr_eih_sources = 11

Strict math: bit-identical (states, t, dt, step / sweep / evaluation / rejection counts per particle); the Marsden
term goes through pow() (glibc vs CUDA, 3e-16 relative): those cases are held to the north star's 1e-12 AU.
Fast math: reported as the fraction of particles beyond 1e-12 AU of strict.  SPK planets file (the DE-binary
format is covered at small sizes by test_gpu_parity.py)."""
import os

import numpy as np
import pytest

import cases
from conftest import ROOT
from assist_b200 import batch as ab

pytestmark = pytest.mark.gpu
POS_TOL_AU = 1e-12


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_large.npz"))


@pytest.fixture(scope="module")
def eph(paths, lib):
    if lib.assist_gpu_device_count() < 1:
        pytest.fail("GPU tests need a CUDA device: assist-b200 has no CPU path")
    return ab.EphemHandle(paths["planets_bsp"], paths["asteroids_bsp"])


def _counts(b):
    c = b.counters()
    return np.stack([c["steps"], c["iters"], c["evals"], c["rejected"]], axis=1).astype(np.int64)


def test_c3_thousand_particle_subsample_ten_years(eph, G):
    st = cases.c3_subsample()
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F, min_dt=1e-3)
    b.set_state(cases.T0, st[:, None, :])
    b.snapshot()
    b.integrate(cases.T0 + 3652.5)
    got = b.get_state()
    assert (got["status"] == 0).all()
    assert np.array_equal(got["state"], G["c3_final"])
    assert np.array_equal(got["t"], G["c3_t"]) and np.array_equal(got["dt"], G["c3_dt"])
    assert np.array_equal(_counts(b), G["c3_counts"])
    assert G["c3_counts"][:, 0].max() > 2000 and G["c3_counts"][:, 0].min() < 250          # the case spans a 10x range of step counts (planet-crossing NEOs)
    # fast math against strict over the same 10 yr
    b.set_options(math=ab.MATH_FAST)
    b.restore()
    b.integrate(cases.T0 + 3652.5)
    d = np.linalg.norm(b.get_state()["state"][:, 0, :3] - G["c3_final"][:, 0, :3], axis=-1)
    frac = float((d > POS_TOL_AU).mean())
    print("C3 fast math vs reference after 10 yr: %.1f %% of 1000 particles beyond 1e-12 AU, max %.2e AU, median %.2e AU" % (
        100 * frac, d.max(), np.median(d)))
    assert frac <= 0.02 and np.median(d) <= 1e-13
    b.close()


def test_c4_variational_five_years(eph, G):
    st = cases.c4_subsample()
    b = ab.Batch(eph, st.shape[0], 6, ab.PER_PARTICLE, forces=0x7F, min_dt=0.0)
    b.set_state(cases.T0, st)
    b.integrate(cases.T0 + 1826.25)
    got = b.get_state()
    assert np.array_equal(got["state"], G["c4_final"])
    assert np.array_equal(got["t"], G["c4_t"]) and np.array_equal(got["dt"], G["c4_dt"])
    assert np.array_equal(_counts(b), G["c4_counts"])
    b.close()


def test_c2_ten_thousand_particles_one_shared_step_simulation(eph, G):
    st = cases.c2_population()
    b = ab.Batch(eph, st.shape[0], 0, ab.SHARED_STEP, forces=0x77)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + 3652.5)
    got = b.get_state()
    assert np.array_equal(got["state"][:, 0, :], G["c2_final"])
    assert [got["t"][0], got["dt"][0], got["dt_last_done"][0]] == list(G["c2_t_dt"])
    s = b.stats()
    assert [s["steps"], s["pc_iterations"], s["steps_rejected"]] == [G["c2_counts"][0], G["c2_counts"][1], G["c2_counts"][3]]
    b.close()


def test_c5_comets_dense_output_fifty_years_backward(eph, G):
    st, prm = cases.c5_subsample()
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F, min_dt=1e-3)
    b.set_state(cases.T0, st[:, None, :], params=prm[:, None, :])
    out = b.integrate_or_interpolate(cases.C5_EPOCHS)[:, :, 0, :]          # [epoch][comet][6]
    assert np.isfinite(out).all()
    d1 = np.linalg.norm(out[cases.C5_SPARSE][..., :3] - G["c5_sparse"][..., :3], axis=-1)
    d2 = np.linalg.norm(out[:, :8, :3] - G["c5_first8"][..., :3], axis=-1)
    print("C5 dense output vs reference: max %.2e AU over 256 comets x 63 epochs, %.2e AU over 8 comets x 1827 epochs" % (d1.max(), d2.max()))
    assert d1.max() <= POS_TOL_AU and d2.max() <= POS_TOL_AU
    b.close()


def test_geocentric_integration(eph, G):
    st = G["geo_init"]
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F, geocentric=1, min_dt=1e-3)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + 400.0)
    got = b.get_state()
    assert np.array_equal(got["state"], G["geo_final"])
    assert np.array_equal(got["t"], G["geo_t"]) and np.array_equal(got["dt"], G["geo_dt"])
    assert np.array_equal(_counts(b), G["geo_counts"])
    b.close()


def test_eleven_eih_sources_batch(eph, G):
    st = cases.eih11_case()
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F, gr_eih_sources=11, min_dt=1e-3)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + 1000.0)
    got = b.get_state()
    assert np.array_equal(got["state"], G["eih_final"])
    assert np.array_equal(got["t"], G["eih_t"]) and np.array_equal(got["dt"], G["eih_dt"])
    assert np.array_equal(_counts(b), G["eih_counts"])
    b.close()
