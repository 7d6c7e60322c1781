"""The reference's Python tests (assist/test/test_basic.py, test_forces.py, test_interpolate.py) against
assist_b200's mirror of that API, with the synthetic ephemeris files.  Where the reference tests compare with
JPL Horizons, these compare with the reference's C code (oracle/_ref) on the same files: bit-identical."""
import math

import numpy as np
import pytest

import assist_b200 as assist
import cases
import refharness as rh
from conftest import planets_path

pytestmark = pytest.mark.gpu

HOLMAN = dict(x=-2.724183384883979E+00, y=-3.523994546329214E-02, z=9.036596202793466E-02,
              vx=-1.374545432301129E-04, vy=-1.027075301472321E-02, vz=-4.195690627695180E-03)
AU2M = 149597870700


@pytest.fixture(scope="module")
def ephem(paths, fmt):
    return assist.Ephem(planets_path(paths, fmt), paths["asteroids_bsp"])


@pytest.fixture(scope="module")
def reph(ref, paths, fmt):
    return rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])


def holman_state():
    return np.array([[HOLMAN[k] for k in ("x", "y", "z", "vx", "vy", "vz")]])


def test_ephem_get_particle(ephem, ref, reph):
    """reference test_basic.py:18-27, 35-39: Sun, a planet, an asteroid; by id and by name."""
    for body, t in ((0, 0.0), (1, 100.0), (20, 200.0), (3, 8416.5), (4, -3000.25)):
        p = ephem.get_particle(body, t)
        want, st = rh.all_bodies(ref, reph, [t])
        assert st[0, body] == 0
        assert [p.x, p.y, p.z] == list(want[0, body, 1:4])
        assert p.m == want[0, body, 0]
        if body < 11:
            assert [p.vx, p.vy, p.vz] == list(want[0, body, 4:7])
    assert ephem.get_particle("Sun", 0).x == ephem.get_particle(0, 0).x
    assert ephem.get_particle("vesta", 10.0).y == ephem.get_particle(26, 10.0).y
    with pytest.raises(RuntimeError):
        ephem.get_particle(0, 1e6)                        # outside the files' coverage


def test_holman(ephem, ref, reph):
    """reference test_basic.py:44-84 (30 d of the Holman asteroid), checked against the reference C build."""
    sim = assist.Simulation()
    extras = assist.Extras(sim, ephem)
    sim.t = 8416.5
    sim.add(**HOLMAN)
    sim.integrate(8446.5)
    assert sim.t == 8446.5
    r = rh.Sim(ref, reph, 8416.5, holman_state())
    r.integrate(8446.5)
    p = sim.particles[0]
    assert [p.x, p.y, p.z, p.vx, p.vy, p.vz] == list(r.state()[0, 0])
    assert sim.steps_done == r.counters()["steps"] and sim.dt == r.dt
    r.close()
    del extras


def test_forces_gr_switch(ephem):
    """reference test_forces.py:25-57: ~100 m after 60 d with the EIH term switched off."""
    t = 8416.5
    sim = assist.Simulation()
    sim.t = t
    sim.add(**HOLMAN)
    sim2 = sim.copy()
    extras = assist.Extras(sim, ephem)
    extras2 = assist.Extras(sim2, ephem)
    forces = extras.forces
    n_forces = len(forces)
    forces.remove("GR_EIH")
    extras.forces = forces
    assert len(extras.forces) == n_forces - 1
    sim.integrate(t + 60.0)
    sim2.integrate(t + 60.0)
    d = sim.particles[0] - sim2.particles[0]
    assert math.fabs(d.x * AU2M - 100.0) < 20          # the reference's own bound
    del extras, extras2


def test_interpolate(ephem):
    """reference test_interpolate.py:10-40: integrate() vs integrate_or_interpolate() at 10 epochs."""
    t = 8416.5
    sim = assist.Simulation()
    sim.t = t
    sim.add(**HOLMAN)
    sim2 = sim.copy()
    extras = assist.Extras(sim, ephem)
    extras2 = assist.Extras(sim2, ephem)
    for _ in range(10):
        t += 40.0
        sim.integrate(t)
        extras2.integrate_or_interpolate(t)
        d = sim.particles[0] - sim2.particles[0]
        assert math.fabs(d.x * AU2M) < 0.02 and math.fabs(d.y * AU2M) < 0.01 and math.fabs(d.z * AU2M) < 0.01
    del extras, extras2


def test_variational_and_params(ephem, ref, reph):
    """A variational particle and Marsden parameters through the Python mirror (reference examples/variational,
    assist/test/test_apophis.py:38-45)."""
    st, prm = cases.comet_case()
    sim = assist.Simulation()
    extras = assist.Extras(sim, ephem)
    sim.t = cases.T0
    sim.add(x=st[0, 0], y=st[0, 1], z=st[0, 2], vx=st[0, 3], vy=st[0, 4], vz=st[0, 5])
    iv = sim.add_variation(testparticle=0)
    sim.particles[iv].x = 1.0
    params = np.zeros((2, 3))
    params[0] = prm[0]
    extras.particle_params = params.flatten()
    extras.gr_eih_sources = 11
    sim.ri_ias15.min_dt = 1e-3
    sim.integrate(cases.T0 + 200.0)
    full = np.zeros((1, 2, 6))
    full[0, 0] = st[0]
    full[0, 1, 0] = 1.0
    r = rh.Sim(ref, reph, cases.T0, full, params=params[None, :, :], gr_eih_sources=11, min_dt=1e-3)
    r.integrate(cases.T0 + 200.0)
    want = r.state()[0]
    got = np.array([[getattr(sim.particles[j], k) for k in ("x", "y", "z", "vx", "vy", "vz")] for j in (0, iv)])
    assert np.linalg.norm(got[0, :3] - want[0, :3]) <= 1e-12      # pow() in the Marsden term: not bit-identical
    assert np.linalg.norm(got[1, :3] - want[1, :3]) <= 1e-9 * max(1.0, np.linalg.norm(want[1, :3]))
    assert sim.t == r.t
    r.close()
    del extras


def test_snapshots_and_interpolated_simulation(ephem, tmp_path):
    """assist/tools.py:11-15 through the Python mirror: snapshots after every step, the simulation interpolated inside
    a step agrees with a direct integration to that time (reference unit_tests/interpolation_spk, 10 cm bound)."""
    t0 = 8416.5
    fn = tmp_path / "py_out.bin"
    sim = assist.Simulation()
    sim.t = t0
    sim.add(**HOLMAN)
    extras = assist.Extras(sim, ephem)
    sim.save_to_file(fn, step=1)
    sim.integrate(t0 + 583.5)
    sa = assist.SimulationArchive(fn)
    assert len(sa) == sim.steps_done + 1 and sa.t[0] == t0
    si = assist.assist_create_interpolated_simulation(sa, t0 + 30.0)
    assert si.t == pytest.approx(t0 + 30.0, abs=1e-9) and si.N == 1
    sim2 = assist.Simulation()
    sim2.t = t0
    sim2.add(**HOLMAN)
    extras2 = assist.Extras(sim2, ephem)
    sim2.integrate(t0 + 30.0)
    d = si.particles[0] - sim2.particles[0]
    assert math.fabs(d.x * AU2M) < 0.1 and math.fabs(d.y * AU2M) < 0.1 and math.fabs(d.z * AU2M) < 0.1
    conv = assist.simulation_convert_to_rebound(sim2, ephem, merge_moon=0)
    assert conv.N == 11 + 1 and conv.particles[11].x == sim2.particles[0].x
    del extras, extras2
