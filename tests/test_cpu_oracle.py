"""CPU tests of the ORACLE: the reference's own C code (oracle/_ref) driven by the IAS15
restatement (oracle/reb_shim.c), pinned against the committed golden vectors and against
the data-independent invariants of the reference's own test-suite."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import refharness as rh
from conftest import ROOT, planets_path, relerr
from assist_b200.cstructs import Particle
from assist_b200.synth import populations

AU_M = 149597870700.0


@pytest.fixture(scope="module")
def reph(ref, paths, fmt):
    return rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])


def test_ias15_constants_are_derived_and_pinned(tmp_path):
    """gen_ias15_constants.py asserts the published h[] and the SURVEY App. A known doubles."""
    out = tmp_path / "c.h"
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "gen_ias15_constants.py"), str(out), "T"], check=True)
    txt = out.read_text()
    assert txt.count("0x") == 78
    # the committed headers of the oracle and of the product hold the same numbers
    # (each side has its own generator: the product's build never touches oracle/)
    a = open(os.path.join(ROOT, "oracle", "ias15_constants.h")).read().replace("ORC_", "X_").split("\n", 1)[1]
    b = open(os.path.join(ROOT, "assist_b200", "csrc", "ias15_constants.h")).read().replace("AB_", "X_").split("\n", 1)[1]
    assert a == b
    out2 = tmp_path / "p.h"
    subprocess.run([sys.executable, os.path.join(ROOT, "assist_b200", "csrc", "gen_ias15_constants.py"), str(out2), "T"], check=True)
    assert out2.read_text().split("\n", 1)[1] == txt.split("\n", 1)[1]


def test_reference_reproduces_golden(ref, reph, fmt, golden):
    """The golden files are exactly what the reference build produces today (bit for bit)."""
    g = golden[fmt]
    out, st = rh.all_bodies(ref, reph, cases.EPHEM_TIMES)
    assert np.array_equal(np.nan_to_num(out, nan=-7), np.nan_to_num(g["ephem"], nan=-7))
    assert np.array_equal(st, g["ephem_status"])
    state, params = cases.force_case()
    for name, mask, src, geo in cases.FORCE_TERMS:
        a = rh.forces(ref, reph, cases.FORCE_T, state, params, forces=mask, gr_eih_sources=src, geocentric=geo)
        assert np.array_equal(a, g["force_" + name]), name
    fin, ts, _, cnt = rh.integrate_each(ref, reph, cases.T0, cases.pp_case(), cases.T0 + cases.PP_DAYS, forces=0x7F)
    assert np.array_equal(fin, g["pp_final"]) and np.array_equal(ts, g["pp_t"])
    assert cnt["steps"] == g["pp_counts"][0]


def test_reference_reproduces_golden_n373(ref):
    """tests/golden/golden_n373.npz (small-body kernels with 40 and 373 targets) is what the reference
    build produces today: N = 40 in full (ephemeris, one force evaluation, the per-particle integrations)."""
    from conftest import ROOT
    from assist_b200.synth import ephem_writer
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_n373.npz"))
    N = cases.N373_SIZES[0]
    p = ephem_writer.write_extended(os.path.join(ROOT, "data"), N)
    e = rh.open_ephem(ref, p["planets_bsp"], p["asteroids_bsp"])
    assert e.contents.spk_asteroids and ref.assist_gpu_ephem_nbodies(e) == 11 + N if hasattr(ref, "assist_gpu_ephem_nbodies") else True
    out, st = rh.all_bodies(ref, e, cases.n373_times(), nbodies=11 + N)
    assert (st == 0).all() and np.array_equal(np.nan_to_num(out, nan=-7), np.nan_to_num(g["eph%d" % N], nan=-7))
    assert np.array_equal(rh.forces(ref, e, cases.T0 + 17.25, cases.n373_force_systems(), forces=0x7F), g["acc%d" % N])
    fin, ts, dts, _ = rh.integrate_each(ref, e, cases.T0, cases.n373_pp_particles(), cases.T0 + 400.0, forces=0x7F)
    assert np.array_equal(fin, g["pp_final"]) and np.array_equal(ts, g["pp_t"]) and np.array_equal(dts, g["pp_dt"])


def test_ephemeris_formats_agree(ref, paths):
    """SPK and DE-binary providers hold the same Chebyshev data: same states (reference
    unit_tests/apophis_drift checks the same thing on real files)."""
    a = rh.open_ephem(ref, paths["planets_bsp"], paths["asteroids_bsp"])
    b = rh.open_ephem(ref, paths["de440"], paths["asteroids_bsp"])
    oa, _ = rh.all_bodies(ref, a, cases.EPHEM_TIMES[:4])
    ob, _ = rh.all_bodies(ref, b, cases.EPHEM_TIMES[:4])
    assert np.allclose(oa[:, :, :4], ob[:, :, :4], rtol=0, atol=1e-14)
    assert a.contents.AU == b.contents.AU and a.contents.EMRAT == b.contents.EMRAT
    assert a.contents.over_c_squared == b.contents.over_c_squared


def test_time_bounds_and_coverage_error(ref, reph):
    tb, te = ctypes.c_double(), ctypes.c_double()
    ref.assist_ephem_time_bounds(reph, tb, te)
    assert (tb.value, te.value) == (-10544.5, 13455.5)
    err = ctypes.c_int(0)
    ref.assist_get_particle_with_error(reph, 0, 20000.0, err)
    assert err.value == 5      # ASSIST_ERROR_COVERAGE


def test_invariant_cache_on_off_bit_identical(ref, reph):
    """reference unit_tests/ephem_cache/problem.c: the 7-slot cache must not change a single bit."""
    st = np.array([[3.3388753502614090e+00, -9.1765182678903168e-01, -5.0385906775843303e-01,
                    2.8056633153049852e-03, 7.5504086883996860e-03, 2.9800282074358684e-03]])
    for direction in (1, -1):
        res = []
        for use_cache in (False, True):
            s = rh.Sim(ref, reph, cases.T0, st)
            if not use_cache:
                s.ax.contents.ephem_cache = None
            s.integrate(cases.T0 + direction * 1000)
            assert s.t == cases.T0 + direction * 1000
            res.append(s.state())
            s.close()
        assert np.array_equal(res[0], res[1])


def test_invariant_roundtrip_fixed_step(ref, reph):
    """reference unit_tests/roundtrip_spk/problem.c: dt = 10, epsilon = 0, out and back."""
    x0 = np.array([3.3388753502614090e+00, -9.1765182678903168e-01, -5.0385906775843303e-01,
                   2.8056633153049852e-03, 7.5504086883996860e-03, 2.9800282074358684e-03])
    for trange, tol_m in ((100, 1e-4), (1000, 1e-3), (4000, 5e-3)):       # the synthetic files end 5000 d after T0
        s = rh.Sim(ref, reph, cases.T0, x0[None, :], epsilon=0.0, dt0=10.0)
        count = 0
        while s.t < cases.T0 + trange:
            ref.reb_simulation_step(s.r)
            count += 1
        s.r.contents.dt *= -1
        for _ in range(count):
            ref.reb_simulation_step(s.r)
        assert s.t == cases.T0
        d = np.linalg.norm(s.state()[0, 0, :3] - x0[:3]) * AU_M
        assert d < tol_m, (trange, d)
        s.close()


def test_invariant_roundtrip_adaptive(ref, reph):
    """reference unit_tests/roundtrip_adaptive_spk/problem.c: three particles, adaptive, out and back."""
    base = np.array([3.3388753502614090e+00, -9.1765182678903168e-01, -5.0385906775843303e-01,
                     5.22000300435103972568e-03, 6.56760310074497024591e-03, 2.44038701006633581073e-03])
    st = np.tile(base, (3, 1))
    st[:, 3] += np.arange(3) / 1e4
    for trange, tol_m in ((100, 1e-2), (1000, 5e-2), (4000, 1e-1)):
        s = rh.Sim(ref, reph, cases.T0, st)
        s.integrate(cases.T0 + trange)
        s.integrate(cases.T0)
        assert s.t == cases.T0
        d = np.mean(np.linalg.norm(s.state()[:, 0, :3] - st[:, :3], axis=1)) * AU_M
        assert d < tol_m, (trange, d)
        s.close()


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_invariant_onthefly_interpolation(ref, reph, sign):
    """reference unit_tests/onthefly_(backwards_)interpolation: dense output equals direct integration to 2e-13 AU."""
    st = np.array([[-2.724183384883979E+00, -3.523994546329214E-02, 9.036596202793466E-02,
                    -1.374545432301129E-04, -1.027075301472321E-02, -4.195690627695180E-03]])
    step = 40.0 if sign > 0 else 10.0
    ts = cases.T0 + sign * step * np.arange(1, 21)
    s1 = rh.Sim(ref, reph, cases.T0, st)
    direct = []
    for t in ts:
        s1.integrate(t)
        direct.append(s1.state()[0, 0, 0])
    s2 = rh.Sim(ref, reph, cases.T0, st)
    interp = [s2.integrate_or_interpolate(t)[0, 0, 0] for t in ts]
    assert np.max(np.abs(np.array(direct) - np.array(interp))) < 2e-13
    s1.close(); s2.close()


def test_invariant_variational_vs_finite_difference(ref, reph):
    """reference unit_tests/variational_spk/problem.c: variational particle vs 10 m shadow particle, 1e-5."""
    p0 = np.array([-2.724183384883979E+00, -3.523994546329214E-02, 9.036596202793466E-02,
                   -1.374545432301129E-04, -1.027075301472321E-02, -4.195690627695180E-03])
    dx = 10.0 / AU_M
    lib = ref
    r = lib.reb_simulation_create()
    ax = lib.assist_attach(r, reph)
    r.contents.t = cases.T0
    lib.reb_simulation_add(r, Particle(x=p0[0], y=p0[1], z=p0[2], vx=p0[3], vy=p0[4], vz=p0[5]))
    lib.reb_simulation_add(r, Particle(x=p0[0] + dx, y=p0[1], z=p0[2], vx=p0[3], vy=p0[4], vz=p0[5]))
    var = lib.reb_simulation_add_variation_1st_order(r, 0)
    r.contents.particles[var].x = 1.0
    for span in (1.0, 100.0):
        lib.reb_simulation_integrate(r, r.contents.t + span)
        P = r.contents.particles
        d1 = P[1].x - P[0].x
        d2 = P[var].x * dx
        assert abs((d1 - d2) / d1) < 1e-5
    lib.assist_free(ax)
    lib.reb_simulation_free(r)


def test_gr_moves_orbit_by_about_100m(ref, reph):
    """reference assist/test/test_forces.py:58: dropping GR_EIH moves a main-belt orbit ~100 m in 60 d."""
    st = np.array([[-2.724183384883979E+00, -3.523994546329214E-02, 9.036596202793466E-02,
                    -1.374545432301129E-04, -1.027075301472321E-02, -4.195690627695180E-03]])
    out = []
    for mask in (0x7F, 0x7F & ~0x40):
        s = rh.Sim(ref, reph, cases.T0, st, forces=mask)
        s.integrate(cases.T0 + 60.0)
        out.append(s.state()[0, 0, :3])
        s.close()
    d = np.linalg.norm(out[0] - out[1]) * AU_M
    assert 50.0 < d < 200.0, d


def test_min_dt_clamp_and_c1_counts(golden):
    """C1 (Apophis-like, 11 EIH sources, non-grav, min_dt 1e-3) completed and stayed finite."""
    for tag in ("bsp", "440"):
        g = golden[tag]
        assert np.isfinite(g["c1_final"]).all()
        assert g["c1_t"][0] == cases.T0 + 3652.5
        assert g["c1_counts"][0] > 100


# ---------------------------------------------------------------------------------------------
# the numpy restatement (oracle/np_oracle.py) against the reference build's outputs
# ---------------------------------------------------------------------------------------------
def _np_oracle():
    import importlib.util
    spec = importlib.util.spec_from_file_location("np_oracle", os.path.join(ROOT, "oracle", "np_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_numpy_restatement_ephemeris(paths, fmt, golden):
    npo = _np_oracle()
    eph = npo.Ephemeris(planets_path(paths, fmt), paths["asteroids_bsp"])
    g = golden[fmt]
    for it, t in enumerate(cases.EPHEM_TIMES):
        if g["ephem_status"][it].max() != 0:
            continue
        gm, pos, vel = eph.states(float(t))
        assert np.array_equal(gm, g["ephem"][it, :, 0])
        assert np.max(np.abs(pos - g["ephem"][it, :, 1:4])) < 1e-14
        assert np.max(np.abs(vel[:11] - g["ephem"][it, :11, 4:7])) < 1e-16 + 1e-13 * np.max(np.abs(g["ephem"][it, :11, 4:7]))


def test_numpy_restatement_force_terms(paths, fmt, golden):
    """Every force term of the structurally independent numpy restatement agrees with the reference
    build to ~1e-13 relative; the variational parts agree with central differences of it to 1e-6."""
    npo = _np_oracle()
    eph = npo.Ephemeris(planets_path(paths, fmt), paths["asteroids_bsp"])
    g = golden[fmt]
    state, params = cases.force_case()
    x, v = state[:, 0, :3], state[:, 0, 3:]
    for name, mask, src, geo in cases.FORCE_TERMS:
        if geo:
            continue
        a = npo.accelerations(eph, cases.FORCE_T, x, v, params[:, 0, :], forces=mask, gr_eih_sources=src)
        want = g["force_" + name][:, 0, :]
        assert relerr(a, want) < 2e-13, (name, relerr(a, want))
    # variational: the reference's Jacobian-times-variation vs numerical differentiation
    # ("sun", 0x07): the reference's variational direct term ignores the force mask (src/forces.c:359), so the
    # variational part of the Sun-only evaluation is the derivative of ALL direct terms
    for name, mask, src in (("sun", 0x07, 1), ("earth_harm", 0x10, 1), ("sun_harm", 0x20, 1), ("eih1", 0x40, 1),
                            ("gr_simple", 0x80, 1), ("gr_potential", 0x100, 1), ("nongrav", 0x08, 1)):
        for k in (1, 4):
            dx, dv = state[:, k, :3], state[:, k, 3:]
            da = npo.variational_by_differences(eph, cases.FORCE_T, x, v, dx, dv, params[:, 0, :], params[:, k, :],
                                                forces=mask, gr_eih_sources=src)
            want = g["force_" + name][:, k, :]
            if name == "nongrav":
                # particles with A1 = A2 = A3 = 0 are skipped by the reference, variational part included (src/forces.c:861)
                skipped = ~np.any(params[:, 0, :] != 0.0, axis=1)
                assert np.all(want[skipped] == 0.0)
                da[skipped] = 0.0
            scale = np.linalg.norm(want, axis=1).max()
            assert np.max(np.linalg.norm(da - want, axis=1)) < 2e-5 * scale, name


def test_ias15_restatement_on_the_kepler_problem(ref, paths, tmp_path):
    """An analytic pin of the IAS15 restatement (oracle/reb_shim.c), independent of any REBOUND source: with the
    planets' masses set to zero the synthetic Sun sits at the origin, and with only the SUN force a particle is on a
    Kepler ellipse.  After 170 periods (e = 0.6, ~1.3e4 steps) a 15th-order integrator at machine precision is back
    at its starting point to 2e-13 AU with a relative energy error of 7e-16 (Rein & Spiegel 2015 promise ~1e-16 sqrt(N)
    and phase errors ~N^1.5); a wrong constant, predictor or b/g update shows up many orders of magnitude above."""
    from assist_b200.synth import ephem_writer as ew
    model = ew.SolarSystemModel()
    for p in model.planets.values():
        p["gm"] = 0.0                                   # sun_bary() == 0: the Sun does not move
    planets = str(tmp_path / "kepler_planets.bsp")
    ew.write_planets_bsp(planets, model)
    reph = rh.open_ephem(ref, planets, paths["asteroids_bsp"])
    mu = ew.CONSTANTS["GMS"]
    a, e = 0.5, 0.6
    period = 2.0 * np.pi * np.sqrt(a ** 3 / mu)
    x0 = np.array([[a * (1 - e), 0.0, 0.0, 0.0, np.sqrt(mu / a * (1 + e) / (1 - e)), 0.0]])
    t0 = -10000.0
    s = rh.Sim(ref, reph, t0, x0, forces=0x01)
    def energy(st):
        return 0.5 * np.dot(st[3:], st[3:]) - mu / np.linalg.norm(st[:3])
    e0 = energy(x0[0])
    n_orbits = 170
    s.integrate(t0 + n_orbits * period)
    st = s.state()[0, 0]
    steps = s.counters()["steps"]
    s.close()
    assert 5000 < steps < 40000
    assert abs((energy(st) - e0) / e0) < 1e-14          # measured 7e-16
    # the end time is n * period only to the rounding of `period`: propagate the analytic orbit over the residual
    dt_res = (t0 + n_orbits * period) - t0 - n_orbits * period
    v0 = x0[0, 4]
    expect = np.array([x0[0, 0], v0 * dt_res, 0.0])
    assert np.linalg.norm(st[:3] - expect) < 1e-11      # measured 2e-13 AU after 23 462 steps
    # and the angular momentum, which every force evaluation of a central force conserves, to rounding
    h0 = x0[0, 0] * x0[0, 4]
    assert abs((st[0] * st[4] - st[1] * st[3]) / h0 - 1.0) < 1e-14


def test_ias15_variant_spread(ref, paths):
    """What a REBOUND binary could change (oracle/ias15_variants.py): the predictor written with
    the reference's s[0..8] coefficients instead of the nested form moves a 10-yr orbit by round-off only (bar: the
    north star's 1e-12 AU; measured 7e-13 AU max over 100 C3 particles), and not restoring the accelerations on a
    rejected step changes nothing where no step is rejected."""
    reph = rh.open_ephem(ref, paths["planets_bsp"], paths["asteroids_bsp"])
    st = populations.neo_mba_mix(1000000, seed=20261703)[::50000]        # 4 NEOs, 16 main-belt objects
    T0 = populations.T0
    try:
        base, _, _, cb = rh.integrate_each(ref, reph, T0, st, T0 + 3652.5, forces=0x7F, min_dt=1e-3)
        ref.reb_shim_set_variant(1, 1)
        alt, _, _, ca = rh.integrate_each(ref, reph, T0, st, T0 + 3652.5, forces=0x7F, min_dt=1e-3)
        ref.reb_shim_set_variant(0, 0)
        keep, _, _, ck = rh.integrate_each(ref, reph, T0, st, T0 + 3652.5, forces=0x7F, min_dt=1e-3)
    finally:
        ref.reb_shim_set_variant(0, 1)
    d = np.linalg.norm(alt[:, 0, :3] - base[:, 0, :3], axis=-1)
    assert 0.0 < d.max() <= 1e-12
    assert abs(ca["steps"] - cb["steps"]) <= 3
    assert cb["rejected"] == 0 and np.array_equal(keep, base) and ck == cb


def test_numpy_restatement_beyond_sixteen_asteroids():
    """The numpy restatement takes any number of small-body targets: ephemeris and one force evaluation on the
    40-target kernel against tests/golden/golden_n373.npz (the reference build's outputs)."""
    from conftest import ROOT
    from assist_b200.synth import ephem_writer
    npo = _np_oracle()
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_n373.npz"))
    N = cases.N373_SIZES[0]
    p = ephem_writer.write_extended(os.path.join(ROOT, "data"), N)
    eph = npo.Ephemeris(p["planets_bsp"], p["asteroids_bsp"])
    assert eph.nbodies == 11 + N
    for it, t in enumerate(cases.n373_times()):
        gm, pos, _ = eph.states(float(t))
        assert np.array_equal(gm, g["eph%d" % N][it, :, 0])
        assert np.max(np.abs(pos - g["eph%d" % N][it, :, 1:4])) < 1e-14
    st = cases.n373_force_systems()
    a = npo.accelerations(eph, cases.T0 + 17.25, st[:, 0, :3], st[:, 0, 3:], None, forces=0x7F)
    assert relerr(a, g["acc%d" % N][:, 0, :]) < 2e-13


def test_numpy_restatement_of_the_snapshot_interpolation(ref, reph):
    """oracle/np_oracle.interpolate_simulation (the IAS15 series written out) against the reference's
    assist_interpolate_simulation (src/assist.c:682-752) on two reference simulations one step apart."""
    npo = _np_oracle()
    st = cases.pp_case()[:5]
    s1 = rh.Sim(ref, reph, cases.T0, st)
    s2 = rh.Sim(ref, reph, cases.T0, st)
    ref.reb_simulation_steps(s1.r, 6)
    ref.reb_simulation_steps(s2.r, 7)
    n = st.shape[0]
    r1, r2 = s1.r.contents.ri_ias15, s2.r.contents.ri_ias15
    x0 = np.array([r1.x0[k] for k in range(3 * n)]); v0 = np.array([r1.v0[k] for k in range(3 * n)])
    a0 = np.array([r2.a0[k] for k in range(3 * n)])
    br = np.array([[getattr(r2.br, "p%d" % q)[k] for k in range(3 * n)] for q in range(7)])
    dt, t1 = s2.dt_last_done, s1.t
    for h in (0.0, 0.3, 0.77, 1.0):
        sa = rh.Sim(ref, reph, cases.T0, st)
        ref.reb_simulation_steps(sa.r, 6)
        assert ref.assist_interpolate_simulation(sa.r, s2.r, h) == 1
        want = sa.state()[:, 0, :]
        px, pv = npo.interpolate_simulation(x0, v0, a0, br, dt, h)
        assert np.max(np.abs(px.reshape(n, 3) - want[:, :3])) < 5e-15 * np.max(np.abs(want[:, :3]))
        assert np.max(np.abs(pv.reshape(n, 3) - want[:, 3:])) < 5e-15 * np.max(np.abs(want[:, 3:]))
        assert sa.t == t1 + dt * h
        sa.close()
    # h = 1 is the later snapshot itself (to rounding)
    assert np.max(np.abs(px.reshape(n, 3) - s2.state()[:, 0, :3])) < 1e-14
    s1.close()
    s2.close()
