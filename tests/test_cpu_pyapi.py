"""Host side of the Python mirror (assist_b200.Ephem / Extras / Simulation): everything the reference's
assist/test/test_basic.py and test_forces.py check that needs no compute."""
import numpy as np
import pytest

import assist_b200 as assist
from conftest import planets_path


@pytest.fixture(scope="module")
def ephem(paths, fmt):
    return assist.Ephem(planets_path(paths, fmt), paths["asteroids_bsp"])


def test_ephem_constants_and_bounds(ephem):
    assert ephem.jd_ref == 2451545.0                      # reference test_basic.py:20
    assert ephem.AU == 149597870.7
    t_min, t_max = ephem.time_bounds()                    # reference test_basic.py:29-33 (synthetic coverage here)
    assert t_min == pytest.approx(-10544.5) and t_max == pytest.approx(13455.5)
    with pytest.raises(AttributeError):
        ephem.no_such_constant


def test_ephem_missing_file_raises():
    with pytest.raises(RuntimeError, match="not been found"):
        assist.Ephem("/nonexistent/planets.bsp", None)


def test_body_names(ephem):
    assert assist.ASSIST_BODY_IDS[0] == "Sun" and assist.ASSIST_BODY_IDS[26] == "Vesta" and len(assist.ASSIST_BODY_IDS) == 27
    with pytest.raises(ValueError, match="Cannot find body"):
        ephem.get_particle("Planet 9", 0)                 # reference test_basic.py:41-42
    with pytest.raises(ValueError, match="Expecting integer"):
        ephem.get_particle(1.5, 0)


def test_forces_property_errors(ephem):
    sim = assist.Simulation()
    extras = assist.Extras(sim, ephem)
    with pytest.raises(AttributeError):
        extras.forces = "no array"                        # reference test_forces.py:15-22
    with pytest.raises(AttributeError):
        extras.forces = ["Magic"]
    with pytest.raises(AttributeError):
        extras.forces = [1, 2, 3]
    forces = extras.forces
    assert forces == ["SUN", "PLANETS", "ASTEROIDS", "NON_GRAVITATIONAL", "EARTH_HARMONICS", "SUN_HARMONICS", "GR_EIH"]
    forces.remove("GR_EIH")
    extras.forces = forces
    assert len(extras.forces) == 6 and "GR_EIH" not in extras.forces
    assert extras.gr_eih_sources == 1 and extras.geocentric == 0
    extras.gr_eih_sources = 11
    assert extras.gr_eih_sources == 11
    assert (extras.alpha, extras.nk, extras.nm, extras.nn, extras.r0) == (1.0, 0.0, 2.0, 5.093, 1.0)
    with pytest.raises(AttributeError):
        extras.particle_params
    extras.particle_params = np.array([1e-9, 2e-10, 0.0])
    extras.detach(sim)


def test_simulation_particles(ephem):
    sim = assist.Simulation()
    extras = assist.Extras(sim, ephem)
    sim.t = 8416.5
    sim.add(x=-2.724183384883979, y=-3.523994546329214e-02, z=9.036596202793466e-02,
            vx=-1.374545432301129e-04, vy=-1.027075301472321e-02, vz=-4.195690627695180e-03)
    sim.add(assist.Particle(x=1.0, vy=0.017))
    assert sim.N == 2 and len(sim.particles) == 2 and sim.t == 8416.5
    assert sim.particles[0].x == -2.724183384883979 and sim.particles[-1].vy == 0.017
    sim.particles[1].x = 1.25                              # views write through to the C array
    assert sim.particles[1].x == 1.25
    d = sim.particles[1] - sim.particles[0]
    assert d.x == 1.25 + 2.724183384883979
    idx = sim.add_variation(testparticle=0)
    assert idx == 2 and sim.N == 3 and sim.N_var == 1
    with pytest.raises(IndexError):
        sim.particles[3]
    with pytest.raises(ValueError):
        sim.add(a=1.0, e=0.1)                              # orbital elements need the rebound package
    sim.ri_ias15.min_dt = 1e-3
    assert sim.ri_ias15.min_dt == 1e-3 and sim.ri_ias15.epsilon == 1e-9
    c = sim.copy()
    assert c.N == 3 and c.t == sim.t and c.particles[1].x == 1.25
    del extras


def test_compute_fails_loudly_without_gpu(ephem, have_gpu):
    if have_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="CUDA"):
        ephem.get_particle("Sun", 0.0)
    sim = assist.Simulation()
    extras = assist.Extras(sim, ephem)
    sim.t = 8416.5
    sim.add(x=-2.7, y=-0.03, z=0.09, vx=-1e-4, vy=-1e-2, vz=-4e-3)
    with pytest.raises(RuntimeError):
        sim.integrate(8420.0)
    del extras


def test_snapshot_file_through_the_python_mirror(ephem, tmp_path):
    """assist/tools.py's assist_create_interpolated_simulation with Simulation.save_to_file / SimulationArchive
    standing in for rebound's: host part (files, ranges, errors)."""
    sim = assist.Simulation()
    extras = assist.Extras(sim, ephem)
    sim.add(x=2.0, y=0.1, z=0.0, vx=0.0, vy=0.012, vz=0.001)
    fn = tmp_path / "py_snap.bin"
    for k in range(3):
        sim.t = 10.0 + k
        sim.save_to_file(fn)
    sa = assist.SimulationArchive(fn)
    assert len(sa) == 3 and sa.t == [10.0, 11.0, 12.0]
    s1 = sa[1]
    assert s1.t == 11.0 and s1.N == 1 and s1.particles[0].x == 2.0
    with pytest.raises(IndexError):
        sa[7]
    with pytest.raises(RuntimeError, match="outside the range"):
        assist.assist_create_interpolated_simulation(sa, 10.5)
    with pytest.raises(RuntimeError, match="cannot read"):
        assist.SimulationArchive(tmp_path / "missing.bin")
    del extras
