"""GPU parity tests: the CUDA path, called through the C-ABI, against the oracle.

Oracle = the committed golden vectors (tests/golden/*.npz, generated from the reference's own
C code) and, when it travelled to the box, the reference build itself (oracle/_ref).

Tolerances (north star): every force term <= 1e-14 relative; integrated positions <= 1e-12 AU.
The strict math variant is additionally required to be BIT-IDENTICAL to the reference wherever
the reference does not call pow() (only the Marsden non-gravitational term does)."""
import ctypes
import os

import numpy as np
import pytest

import cases
import refharness as rh
from conftest import ROOT, planets_path, relerr
from assist_b200 import batch as ab
from assist_b200.cstructs import Particle
from assist_b200.synth import populations

pytestmark = pytest.mark.gpu

TERM_TOL = 1e-14
POS_TOL_AU = 1e-12


@pytest.fixture(scope="module")
def eph(paths, fmt, lib):
    if lib.assist_gpu_device_count() < 1:
        pytest.fail("GPU tests need a CUDA device: assist-b200 has no CPU path")
    return ab.EphemHandle(planets_path(paths, fmt), paths["asteroids_bsp"])


def same(a, b):
    return np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))


# ------------------------------------------------------------------ ephemeris
def test_ephemeris_bit_identical(eph, fmt, golden):
    g = golden[fmt]
    for math in (ab.MATH_STRICT, ab.MATH_FAST):
        out, st = eph.eval(cases.EPHEM_TIMES, math=math)
        assert np.array_equal(st, g["ephem_status"])
        ok = g["ephem_status"] == 0
        if math == ab.MATH_STRICT:
            assert same(out[ok], g["ephem"][ok])
        else:
            assert np.nanmax(np.abs(out[ok][:, 1:4] - g["ephem"][ok][:, 1:4])) < 1e-14


def test_ephemeris_against_reference_build(eph, fmt, ref, paths):
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    rng = np.random.default_rng(1)
    times = np.sort(rng.uniform(-10544.0, 13455.0, 200))
    out, st = eph.eval(times)
    rout, rst = rh.all_bodies(ref, reph, times)
    assert np.array_equal(st, rst) and same(out, rout)


# ------------------------------------------------------------------ force terms
@pytest.mark.parametrize("term", [t[0] for t in cases.FORCE_TERMS])
def test_force_terms(eph, fmt, golden, term):
    name, mask, src, geo = [t for t in cases.FORCE_TERMS if t[0] == term][0]
    state, params = cases.force_case()
    want = golden[fmt]["force_" + name]
    got = ab.eval_forces(eph, cases.FORCE_T, state, params, forces=mask, gr_eih_sources=src, geocentric=geo)
    assert relerr(got, want) <= TERM_TOL, name
    if not (mask & 0x08):
        assert np.array_equal(got, want), "strict math must reproduce the reference bit for bit (%s)" % name
    fast = ab.eval_forces(eph, cases.FORCE_T, state, params, forces=mask, gr_eih_sources=src, geocentric=geo, math=ab.MATH_FAST)
    assert relerr(fast, want) <= TERM_TOL, name


def test_forces_per_system_times_and_no_params(eph, fmt, ref, paths):
    """Each system at its own time (per-particle mode's evaluation), particle_params == NULL."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    st = populations.with_variations(populations.neo_mba_mix(12, seed=77), 6)
    times = cases.T0 + np.linspace(-900.0, 2500.0, 12)
    got = ab.eval_forces(eph, times, st, None, forces=0x7F, gr_eih_sources=3)
    for i in range(12):
        want = rh.forces(ref, reph, times[i], st[i:i + 1], None, forces=0x7F, gr_eih_sources=3)
        assert np.array_equal(got[i:i + 1], want)


# ------------------------------------------------------------------ per-particle IAS15
def test_per_particle_integration_bit_identical(eph, fmt, golden):
    g = golden[fmt]
    st = cases.pp_case()
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + cases.PP_DAYS)
    got = b.get_state()
    assert np.array_equal(got["state"], g["pp_final"])
    assert np.array_equal(got["t"], g["pp_t"]) and np.array_equal(got["dt"], g["pp_dt"])
    s = b.stats()
    assert [s["steps"], s["pc_iterations"], s["force_evals"], s["steps_rejected"]] == list(g["pp_counts"])
    assert (got["status"] == 0).all()


def test_per_particle_fast_math_within_tolerance(eph, fmt, golden):
    g = golden[fmt]
    st = cases.pp_case()
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F, math=ab.MATH_FAST)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + cases.PP_DAYS)
    d = np.linalg.norm(b.get_state()["state"][:, 0, :3] - g["pp_final"][:, 0, :3], axis=-1)
    assert d.max() <= POS_TOL_AU


def test_per_particle_backward(eph, fmt, golden):
    g = golden[fmt]
    st = cases.pp_case()[:6]
    b = ab.Batch(eph, 6, 0, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 - 300.0)
    got = b.get_state()
    assert np.array_equal(got["state"], g["ppback_final"]) and np.array_equal(got["t"], g["ppback_t"])


def test_per_particle_variational(eph, fmt, golden):
    g = golden[fmt]
    stv = cases.var_case()
    b = ab.Batch(eph, stv.shape[0], 6, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, stv)
    b.integrate(cases.T0 + cases.VAR_DAYS)
    got = b.get_state()
    assert np.array_equal(got["state"], g["var_final"]) and np.array_equal(got["t"], g["var_t"])


def test_c1_apophis_like_ten_years(eph, fmt, golden):
    """BASELINE config 1: all forces, 11 EIH sources, Marsden A1/A2, min_dt 1e-3, 10 yr (pow() in the path)."""
    g = golden[fmt]
    c1, prm = populations.apophis_like()
    b = ab.Batch(eph, 1, 0, ab.PER_PARTICLE, forces=0x7F, gr_eih_sources=11, min_dt=1e-3)
    b.set_state(cases.T0, c1[:, None, :], params=prm[:, None, :])
    b.integrate(cases.T0 + 3652.5)
    got = b.get_state()
    assert got["t"][0] == g["c1_t"][0]
    assert np.linalg.norm(got["state"][0, 0, :3] - g["c1_final"][0, 0, :3]) <= POS_TOL_AU


def test_schedulers_do_not_change_results(eph, fmt, golden, monkeypatch):
    """The default work-queue scheduler, and the alternative that pauses/resumes integrate() every few steps
    and packs the stragglers, are invisible in the results."""
    g = golden[fmt]
    st = cases.pp_case()
    for sched, cap, days in (("queue", "32", "0"), ("queue", "32", "7"), ("queue", "32", "32"), ("queue", "32", "1000"),
                             ("capped", "0", "0"), ("capped", "1", "0"), ("capped", "5", "0")):
        monkeypatch.setenv("ASSIST_B200_SCHED", sched)
        monkeypatch.setenv("ASSIST_B200_STEP_CAP", cap)
        monkeypatch.setenv("ASSIST_B200_SLICE_DAYS", days)      # time slices of the work queue
        b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F)
        b.set_state(cases.T0, st[:, None, :])
        b.integrate(cases.T0 + cases.PP_DAYS)
        assert np.array_equal(b.get_state()["state"], g["pp_final"]), (sched, cap, days)
        b.close()
    # more systems than working slots: every slot is reused many times
    monkeypatch.setenv("ASSIST_B200_SCHED", "queue")
    monkeypatch.setenv("ASSIST_B200_SLICE_DAYS", "10")
    n = 150000
    stn = populations.main_belt(n, seed=71)
    b = ab.Batch(eph, n, 0, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, stn[:, None, :])
    b.integrate(cases.T0 + 30.0)
    q = b.get_state()["state"].copy()
    b.close()
    monkeypatch.setenv("ASSIST_B200_SCHED", "capped")
    b = ab.Batch(eph, n, 0, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, stn[:, None, :])
    b.integrate(cases.T0 + 30.0)
    assert np.array_equal(b.get_state()["state"], q)
    b.close()


# ------------------------------------------------------------------ shared step
def test_shared_step_bit_identical(eph, fmt, golden):
    g = golden[fmt]
    st = cases.shared_case()
    b = ab.Batch(eph, st.shape[0], 0, ab.SHARED_STEP, forces=0x77)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + cases.SH_DAYS)
    got = b.get_state()
    assert np.array_equal(got["state"], g["sh_final"])
    assert [got["t"][0], got["dt"][0], got["dt_last_done"][0]] == list(g["sh_t_dt"])
    s = b.stats()
    assert [s["steps"], s["pc_iterations"], s["steps_rejected"]] == [g["sh_counts"][0], g["sh_counts"][1], g["sh_counts"][3]]


def test_shared_step_with_variational(eph, fmt, golden):
    g = golden[fmt]
    shv = cases.shared_var_case()
    b = ab.Batch(eph, 2, 2, ab.SHARED_STEP, forces=0x7F)
    b.set_state(cases.T0, shv)
    b.integrate(cases.T0 + 101.0)
    assert np.array_equal(b.get_state()["state"], g["shv_final"])


def test_shared_step_many_ctas_matches_reference(eph, fmt, ref, paths):
    """More systems than one CTA holds: grid-wide reductions must give the reference's dt sequence."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    st = populations.main_belt(700, seed=31)
    s = rh.Sim(ref, reph, cases.T0, st, forces=0x77)
    s.integrate(cases.T0 + 90.0)
    b = ab.Batch(eph, 700, 0, ab.SHARED_STEP, forces=0x77)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(cases.T0 + 90.0)
    got = b.get_state()
    assert np.array_equal(got["state"], s.state()) and got["t"][0] == s.t and got["dt"][0] == s.dt
    s.close()


# ------------------------------------------------------------------ dense output
def test_dense_output_comets(eph, fmt, golden):
    """BASELINE config 5 in small: Marsden comets, backward, assist_integrate_or_interpolate semantics."""
    g = golden[fmt]
    stc, prm = cases.comet_case()
    b = ab.Batch(eph, stc.shape[0], 0, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, stc[:, None, :], params=prm[:, None, :])
    out = b.integrate_or_interpolate(cases.DENSE_TIMES)
    assert np.nanmax(np.linalg.norm(out[..., :3] - g["dense"][..., :3], axis=-1)) <= POS_TOL_AU
    assert np.isfinite(out).all()


def test_dense_output_schedulers_agree(eph, fmt, monkeypatch):
    """Epoch output through the work queue (default) and through the one-thread-per-system kernel: same bits,
    same final time/step bookkeeping, also when the slots of the working batch are reused."""
    n = 70000
    st = populations.main_belt(n, seed=72)
    times = cases.T0 - np.array([0.0, 3.0, 10.0, 10.5, 11.0, 40.0, 41.0])
    res = {}
    for sched, days in (("queue", "0"), ("queue", "6"), ("queue", "128"), ("capped", "0")):
        monkeypatch.setenv("ASSIST_B200_SCHED", sched)
        monkeypatch.setenv("ASSIST_B200_SLICE_DAYS", days)
        b = ab.Batch(eph, n, 0, ab.PER_PARTICLE, forces=0x7F)
        b.set_state(cases.T0, st[:, None, :])
        out = b.integrate_or_interpolate(times)
        fin = b.get_state()
        res[sched + days] = (out.copy(), fin["state"].copy(), fin["t"].copy(), fin["dt"].copy(), fin["dt_last_done"].copy())
        b.close()
    for key in ("queue0", "queue6", "queue128"):
        for a, c in zip(res[key], res["capped0"]):
            assert np.array_equal(a, c, equal_nan=True), key
    assert np.isfinite(res["queue6"][0]).all()
    # variational systems too
    stv = populations.with_variations(populations.main_belt(300, seed=73), 6)
    outs = []
    for sched, days in (("queue", "9"), ("capped", "0")):
        monkeypatch.setenv("ASSIST_B200_SCHED", sched)
        monkeypatch.setenv("ASSIST_B200_SLICE_DAYS", days)
        b = ab.Batch(eph, 300, 6, ab.PER_PARTICLE, forces=0x7F)
        b.set_state(cases.T0, stv)
        outs.append(b.integrate_or_interpolate(cases.T0 + np.array([5.0, 20.0, 21.0, 60.0])).copy())
        # a second call continues from wherever the first one left the systems, then one goes back in time
        outs.append(b.integrate_or_interpolate(cases.T0 + np.array([61.0, 90.0])).copy())
        outs.append(b.integrate_or_interpolate(cases.T0 + np.array([80.0, 30.0, 31.0, 70.0])).copy())
        b.integrate(cases.T0 + 10.0)
        outs.append(b.get_state()["state"].copy())
        b.close()
    for a, c in zip(outs[:4], outs[4:]):
        assert np.array_equal(a, c, equal_nan=True)


# ------------------------------------------------------------------ the drop-in C API
def _product_sim(lib, ephptr, t0, state, forces=None):
    r = lib.reb_simulation_create()
    ax = lib.assist_attach(r, ephptr)
    r.contents.t = t0
    for s in state:
        lib.reb_simulation_add(r, Particle(x=s[0], y=s[1], z=s[2], vx=s[3], vy=s[4], vz=s[5]))
    if forces is not None:
        ax.contents.forces = forces
    return r, ax


def test_dropin_reb_simulation_integrate(eph, fmt, golden, lib):
    """reb_simulation_create / assist_attach / reb_simulation_add / reb_simulation_integrate on the PRODUCT library."""
    g = golden[fmt]
    st = cases.shared_case()
    r, ax = _product_sim(lib, eph.ptr, cases.T0, st, forces=0x77)
    status = lib.reb_simulation_integrate(r, cases.T0 + cases.SH_DAYS)
    s = r.contents
    assert status == 0 and s.t == g["sh_t_dt"][0] and s.dt == g["sh_t_dt"][1] and s.dt_last_done == g["sh_t_dt"][2]
    got = np.array([[s.particles[i].x, s.particles[i].y, s.particles[i].z, s.particles[i].vx, s.particles[i].vy, s.particles[i].vz]
                    for i in range(st.shape[0])])
    assert np.array_equal(got, g["sh_final"][:, 0, :])
    assert s.steps_done == g["sh_counts"][0]
    lib.assist_free(ax)
    lib.reb_simulation_free(r)


def test_dropin_variational_and_two_calls(eph, fmt, ref, paths, lib):
    """reference unit_tests/variational_spk layout run on both libraries, two consecutive integrate calls."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    p0 = np.array([-2.724183384883979E+00, -3.523994546329214E-02, 9.036596202793466E-02,
                   -1.374545432301129E-04, -1.027075301472321E-02, -4.195690627695180E-03])
    res = []
    for L, E in ((ref, reph), (lib, eph.ptr)):
        r = L.reb_simulation_create()
        ax = L.assist_attach(r, E)
        r.contents.t = cases.T0
        L.reb_simulation_add(r, Particle(x=p0[0], y=p0[1], z=p0[2], vx=p0[3], vy=p0[4], vz=p0[5]))
        L.reb_simulation_add(r, Particle(x=p0[0] + 1e-10, y=p0[1], z=p0[2], vx=p0[3], vy=p0[4], vz=p0[5]))
        var = L.reb_simulation_add_variation_1st_order(r, 0)
        r.contents.particles[var].x = 1.0
        L.reb_simulation_integrate(r, cases.T0 + 1.0)
        L.reb_simulation_integrate(r, r.contents.t + 100.0)
        P = r.contents.particles
        res.append([(P[i].x, P[i].y, P[i].z, P[i].vx, P[i].vy, P[i].vz) for i in range(3)] + [(r.contents.t, r.contents.dt, 0, 0, 0, 0)])
        L.assist_free(ax)
        L.reb_simulation_free(r)
    assert np.array_equal(np.array(res[0]), np.array(res[1]))


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_dropin_integrate_or_interpolate(eph, fmt, ref, paths, lib, sign):
    """assist_integrate_or_interpolate on the product library equals the reference at every epoch
    (reference unit_tests/onthefly_interpolation and onthefly_backwards_interpolation)."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    st = np.array([[-2.724183384883979E+00, -3.523994546329214E-02, 9.036596202793466E-02,
                    -1.374545432301129E-04, -1.027075301472321E-02, -4.195690627695180E-03]])
    ts = cases.T0 + sign * (40.0 if sign > 0 else 10.0) * np.arange(1, 21)
    s = rh.Sim(ref, reph, cases.T0, st)
    want = np.array([s.integrate_or_interpolate(t)[0, 0] for t in ts])
    r, ax = _product_sim(lib, eph.ptr, cases.T0, st)
    got = []
    for t in ts:
        lib.assist_integrate_or_interpolate(ax, float(t))
        p = r.contents.particles[0]
        got.append((p.x, p.y, p.z, p.vx, p.vy, p.vz))
    assert np.array_equal(np.array(got), want)
    lib.assist_free(ax)
    lib.reb_simulation_free(r)
    s.close()


def test_dropin_get_particle_and_update_acceleration(eph, fmt, golden, lib):
    g = golden[fmt]
    for body in (0, 3, 4, 10, 11, 26):
        err = ctypes.c_int(0)
        p = lib.assist_get_particle_with_error(eph.ptr, body, float(cases.EPHEM_TIMES[1]), err)
        assert err.value == 0
        want = g["ephem"][1, body]
        assert (p.m, p.x, p.y, p.z) == tuple(want[:4])
    err = ctypes.c_int(0)
    lib.assist_get_particle_with_error(eph.ptr, 0, 20000.0, err)
    assert err.value == 5
    # assist_additional_forces through reb_simulation_update_acceleration
    state, params = cases.force_case()
    got = rh.forces(lib, eph.ptr, cases.FORCE_T, state[:, :3], params[:, :3], forces=0x77, gr_eih_sources=1)
    want = ab.eval_forces(eph, cases.FORCE_T, state[:, :3], params[:, :3], forces=0x77, gr_eih_sources=1)
    assert np.array_equal(got, want)


def test_dropin_coverage_error_is_reported(eph, lib):
    """Leaving the ephemeris coverage sets REB_STATUS_GENERIC_ERROR (reference src/forces.c:317-321)
    instead of reading outside the tables."""
    st = cases.shared_case()[:3]
    r, ax = _product_sim(lib, eph.ptr, 13400.0, st)
    status = lib.reb_simulation_integrate(r, 13600.0)
    assert status == 1 and r.contents.t < 13455.5
    lib.assist_free(ax)
    lib.reb_simulation_free(r)


# ------------------------------------------------------------------ size-independent properties at scale
def test_properties_at_scale(eph, fmt):
    """20 000 particles: results do not depend on the batch a particle is integrated in, nor on its
    position in it; restore() reproduces a run exactly; out-and-back returns to the start."""
    n = 20000
    st = populations.neo_mba_mix(n, seed=99)
    b = ab.Batch(eph, n, 0, ab.PER_PARTICLE, forces=0x7F, min_dt=1e-3)
    b.set_state(cases.T0, st[:, None, :])
    b.snapshot()
    b.integrate(cases.T0 + 120.0)
    full = b.get_state()["state"].copy()
    assert np.isfinite(full).all()
    b.restore()
    b.integrate(cases.T0 + 120.0)
    assert np.array_equal(b.get_state()["state"], full)
    perm = np.random.default_rng(5).permutation(n)[:3000]
    b2 = ab.Batch(eph, perm.size, 0, ab.PER_PARTICLE, forces=0x7F, min_dt=1e-3)
    b2.set_state(cases.T0, st[perm][:, None, :])
    b2.integrate(cases.T0 + 120.0)
    assert np.array_equal(b2.get_state()["state"], full[perm])
    # out and back (reference unit_tests/roundtrip_adaptive: metres after thousands of days; here 120 d)
    b.integrate(cases.T0)
    back = b.get_state()
    assert (back["t"] == cases.T0).all()
    d = np.linalg.norm(back["state"][:, 0, :3] - st[:, :3], axis=-1)
    assert np.median(d) < 1e-13 and d.max() < 1e-9


def test_full_size_population(eph, fmt, ref, paths):
    """BASELINE config 3 at its full size (10^6 particles of the bench population, 60 d): every system finishes,
    a random subset integrated alone gives the same bits (so the 10^6-batch is 10^6 independent reference
    integrations), and a handful of them are checked against the reference C build itself."""
    if fmt != "bsp":
        pytest.skip("one file format is enough at this size")
    n = 1000000
    st = populations.neo_mba_mix(n, seed=20261703)
    t_end = cases.T0 + 60.0
    b = ab.Batch(eph, n, 0, ab.PER_PARTICLE, forces=0x7F, min_dt=1e-3)
    b.set_state(cases.T0, st[:, None, :])
    b.integrate(t_end)
    got = b.get_state()
    full = got["state"].copy()
    assert (got["status"] == 0).all() and (got["t"] == t_end).all() and np.isfinite(full).all()
    c = b.counters()
    assert c["steps"].min() >= 2 and int(c["evals"].sum()) == b.stats()["force_evals"]
    b.close()
    pick = np.sort(np.random.default_rng(11).choice(n, 5000, replace=False))
    b2 = ab.Batch(eph, pick.size, 0, ab.PER_PARTICLE, forces=0x7F, min_dt=1e-3)
    b2.set_state(cases.T0, st[pick][:, None, :])
    b2.integrate(t_end)
    assert np.array_equal(b2.get_state()["state"], full[pick])
    b2.close()
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    few = pick[::500]
    want, _, _, _ = rh.integrate_each(ref, reph, cases.T0, st[few], t_end, forces=0x7F, min_dt=1e-3)
    assert np.array_equal(full[few], want)


def test_kepler_orbit_analytic(tmp_path, paths, lib):
    """The analytic pin of tests/test_cpu_oracle.py on the CUDA path: Sun at rest (planet masses zeroed in the
    synthetic model), SUN force only, e = 0.6: back at the start after 170 periods."""
    from assist_b200.synth import ephem_writer as ew
    model = ew.SolarSystemModel()
    for p in model.planets.values():
        p["gm"] = 0.0
    planets = str(tmp_path / "kepler_planets.bsp")
    ew.write_planets_bsp(planets, model)
    e_ = ab.EphemHandle(planets, paths["asteroids_bsp"])
    mu = ew.CONSTANTS["GMS"]
    a, e = 0.5, 0.6
    period = 2.0 * np.pi * np.sqrt(a ** 3 / mu)
    x0 = np.array([[a * (1 - e), 0.0, 0.0, 0.0, np.sqrt(mu / a * (1 + e) / (1 - e)), 0.0]])
    t0 = -10000.0
    b = ab.Batch(e_, 1, 0, ab.PER_PARTICLE, forces=0x01)
    b.set_state(t0, x0[:, None, :])
    b.integrate(t0 + 170 * period)
    st = b.get_state()["state"][0, 0]
    b.close()
    def energy(q):
        return 0.5 * np.dot(q[3:], q[3:]) - mu / np.linalg.norm(q[:3])
    assert abs(energy(st) / energy(x0[0]) - 1.0) < 1e-14
    dt_res = (t0 + 170 * period) - t0 - 170 * period
    assert np.linalg.norm(st[:3] - np.array([x0[0, 0], x0[0, 4] * dt_res, 0.0])) < 1e-11


def test_branch_free_division_and_sqrt(lib):
    """fp_device.cuh: the branch-free quotient / square root used for groups of bodies equal the built-in IEEE
    operators bit for bit on 2^28 seeded operand pairs (random and structured mantissas, exponents over the whole
    range the callers' guard lets through)."""
    bad = (ctypes.c_ulonglong * 2)(7, 7)
    rc = lib.assist_gpu_selftest_fp(20261017, 1 << 28, bad)
    assert rc == 0, lib.assist_gpu_last_error()
    assert (bad[0], bad[1]) == (0, 0)


def test_one_population_over_several_devices(eph, fmt, lib):
    """assist_gpu_multi_*: the library deals a population out over the devices, runs them side by side on host threads
    and gathers the outputs in the caller's order -- the bits of a single-device batch.  On a one-GPU box the two
    sub-batches share device 0 (the deal, the threads and the gather are the same code)."""
    n = 3000
    st = populations.neo_mba_mix(n, seed=77)
    one = ab.Batch(eph, n, 0, ab.PER_PARTICLE, forces=0x7F, min_dt=1e-3)
    one.set_state(cases.T0, st[:, None, :])
    one.integrate(cases.T0 + 500.0)
    want, wc = one.get_state(), one.counters()
    devices = [0, 1, 2] if lib.assist_gpu_device_count() >= 3 else ([0, 1] if lib.assist_gpu_device_count() >= 2 else [0, 0])
    m = ab.MultiBatch(eph, n, 0, devices=devices, forces=0x7F, min_dt=1e-3)
    assert m.n_devices == len(devices)
    m.set_state(cases.T0, st[:, None, :])
    m.integrate(cases.T0 + 500.0)
    got, gc = m.get_state(), m.counters()
    for k in ("state", "t", "dt", "dt_last_done", "status"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("steps", "rejected", "iters", "evals"):
        assert np.array_equal(gc[k], wc[k]), k
    s = m.stats()
    assert s["steps"] == one.stats()["steps"] and len(s["kernel_ms_per_device"]) == len(devices)
    # epoch output through the same deal
    times = cases.T0 + 500.0 + 25.0 * np.arange(1, 9)
    assert same(m.integrate_or_interpolate(times), one.integrate_or_interpolate(times))
    m.close(); one.close()


@pytest.mark.parametrize("nvar", [0, 1])
def test_a_stuck_system_is_retired_not_looped_on(eph, fmt, monkeypatch, nvar):
    """a zero timestep or an exhausted budget of step attempts ends THAT system with an error status
    (1000 + message 7) and the launch returns; the other systems finish.  nvar = 0 runs pp_coop_kernel, nvar = 1
    the work-queue kernel."""
    st = populations.main_belt(40, seed=9)
    state = populations.with_variations(st, nvar) if nvar else st[:, None, :]
    b = ab.Batch(eph, 40, nvar, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, state, dt0=0.0)
    b.integrate(cases.T0 + 10.0)
    assert (b.get_state()["status"] == 1007).all()
    b.close()
    monkeypatch.setenv("ASSIST_B200_ATTEMPT_BUDGET", "5")
    b = ab.Batch(eph, 40, nvar, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, state)
    b.integrate(cases.T0 + 1000.0)
    got = b.get_state()
    c = b.counters()
    assert (got["status"] == 1007).all() and (got["t"] < cases.T0 + 1000.0).all()
    assert (c["steps"] + (c["rejected"] if nvar == 0 else 0) == 5).all()
    b.close()
    monkeypatch.delenv("ASSIST_B200_ATTEMPT_BUDGET")
    b = ab.Batch(eph, 40, nvar, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(cases.T0, state)
    b.integrate(cases.T0 + 100.0)
    assert (b.get_state()["status"] == 0).all()
    b.close()


@pytest.mark.parametrize("merge_moon", [0, 1])
def test_convert_to_rebound_matches_the_reference(eph, fmt, ref, paths, lib, merge_moon):
    """assist_simulation_convert_to_rebound (reference src/tools.c:35-70) on both libraries: the same particles in the
    same order (eleven ephemeris bodies, the Earth-Moon barycentre when asked for, then the test particles), N_active,
    t, dt -- the body states coming from the GPU ephemeris evaluation are the reference's bits."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    st = populations.main_belt(3, seed=5)
    out = []
    for L, E in ((ref, reph), (lib, eph.ptr)):
        L.assist_simulation_convert_to_rebound.restype = ctypes.POINTER(type(L.reb_simulation_create().contents))
        L.assist_simulation_convert_to_rebound.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        r = L.reb_simulation_create()
        ax = L.assist_attach(r, E)
        r.contents.t = cases.T0 + 12.5
        for s in st:
            L.reb_simulation_add(r, Particle(x=s[0], y=s[1], z=s[2], vx=s[3], vy=s[4], vz=s[5]))
        r2 = L.assist_simulation_convert_to_rebound(ctypes.cast(r, ctypes.c_void_p), ctypes.cast(E, ctypes.c_void_p), merge_moon)
        assert bool(r2)
        c = r2.contents
        rows = np.array([[c.particles[i].x, c.particles[i].y, c.particles[i].z, c.particles[i].vx, c.particles[i].vy, c.particles[i].vz,
                          c.particles[i].m] for i in range(c.N)])
        out.append((int(c.N), int(c.N_active), c.t, c.dt, rows))
        L.reb_simulation_free(r2)
        L.assist_free(ax)
        L.reb_simulation_free(r)
    assert out[0][0] == out[1][0] == 11 + merge_moon + 3 and out[0][1] == out[1][1] == 11 + merge_moon
    assert out[0][2] == out[1][2] and out[0][3] == out[1][3]
    assert np.array_equal(out[0][4], out[1][4])


def test_dropin_snapshot_file_and_interpolated_simulation(eph, fmt, ref, paths, lib, tmp_path):
    """reference unit_tests/interpolation_spk / interpolation_ascii on the synthetic files: a snapshot
    after every step (reb_simulation_save_to_file_step), the file read back, assist_create_interpolated_simulation at
    a time inside a step.  The snapshots carry the times of the reference's own steps bit for bit; the interpolated
    state equals the reference's assist_interpolate_simulation (src/assist.c:682-752) run on two reference
    simulations stepped to either side of that time, bit for bit; and it agrees with a direct integration."""
    t0 = cases.T0
    st = np.array([[-2.724183384883979E+00, -3.523994546329214E-02, 9.036596202793466E-02,
                    -1.374545432301129E-04, -1.027075301472321E-02, -4.195690627695180E-03],
                   [1.1, 0.3, 0.05, -0.004, 0.014, 0.002]])
    tq = t0 + 30.0
    fn = str(tmp_path / ("out_%s.bin" % fmt)).encode()
    r, ax = _product_sim(lib, eph.ptr, t0, st)
    lib.reb_simulation_save_to_file_step(r, fn, 1)
    assert lib.reb_simulation_integrate(r, t0 + 584.0) == 0
    steps = int(r.contents.steps_done)
    lib.assist_free(ax)
    lib.reb_simulation_free(r)
    sa = lib.reb_simulationarchive_create_from_file(fn)
    assert sa and sa.contents.nblobs == steps + 1 and sa.contents.t[0] == t0
    times = np.array([sa.contents.t[i] for i in range(sa.contents.nblobs)])
    blob = int(np.argmax(times[1:] >= tq)) + 1
    assert 1 < blob < steps
    ri = lib.assist_create_interpolated_simulation(sa, tq)
    assert ri and ri.contents.N == 2
    got = np.array([[ri.contents.particles[i].x, ri.contents.particles[i].y, ri.contents.particles[i].z,
                     ri.contents.particles[i].vx, ri.contents.particles[i].vy, ri.contents.particles[i].vz] for i in range(2)])
    got_t = ri.contents.t
    lib.reb_simulation_free(ri)
    # outside the stored range: NULL, as the reference
    assert not lib.assist_create_interpolated_simulation(sa, times[1]) and not lib.assist_create_interpolated_simulation(sa, times[-1])
    lib.reb_simulationarchive_free(sa)

    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    s1 = rh.Sim(ref, reph, t0, st)
    s2 = rh.Sim(ref, reph, t0, st)
    ref.reb_simulation_steps(s1.r, blob - 1)
    ref.reb_simulation_steps(s2.r, blob)
    assert s1.t == times[blob - 1] and s2.t == times[blob]
    h = (tq - s1.t) / s2.dt_last_done
    assert ref.assist_interpolate_simulation(s1.r, s2.r, h) == 1
    want = s1.state()[:, 0, :]
    assert np.array_equal(got, want) and got_t == s1.t
    s1.close()
    s2.close()

    b = ab.Batch(eph, 2, 0, ab.SHARED_STEP)
    b.set_state(t0, st[:, None, :])
    b.integrate(tq)
    direct = b.get_state()["state"][:, 0, :]
    b.close()
    assert np.abs(got[:, :3] - direct[:, :3]).max() < 5e-13


@pytest.mark.parametrize("direction", [1.0, -1.0])
def test_steps_across_the_segment_boundary(eph, fmt, ref, paths, direction):
    """Two SPK segments per target (boundary at t = 1455.5): steps whose nodes straddle it, forward and backward, on
    pp_coop_kernel (the staged fill leaves a series whose slots straddle segments to the global path) against one
    reference simulation per particle: bit-identical."""
    if fmt != "bsp":
        pytest.skip("segments are an SPK notion")
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    tb = 1455.5
    st = populations.neo_mba_mix(96, seed=4242)
    t0, t1 = tb - direction * 35.0, tb + direction * 70.0
    b = ab.Batch(eph, st.shape[0], 0, ab.PER_PARTICLE, forces=0x7F)
    b.set_state(t0, st[:, None, :])
    b.integrate(t1)
    got = b.get_state()
    b.close()
    want, wt, wdt, _ = rh.integrate_each(ref, reph, t0, st, t1, forces=0x7F)
    assert np.array_equal(got["state"], want) and np.array_equal(got["t"], wt) and np.array_equal(got["dt"], wdt)
