"""pp_coop_kernel's SOURCE run on the host (tests/emul, one OS thread per warp) against the golden vectors and the
reference build.  This checks the kernel's control flow (roles, barriers, queue, accept / reject, epochs) and the
arithmetic of its strict variant without a GPU; the CUDA build itself is checked by tests/test_gpu_parity.py."""
import numpy as np
import pytest

import cases
import coop_emul
import refharness as rh
from conftest import planets_path
from assist_b200 import batch as ab
from assist_b200.synth import populations


@pytest.fixture(scope="module")
def eph(paths, fmt, lib):
    return ab.EphemHandle(planets_path(paths, fmt), paths["asteroids_bsp"])


def same(a, b):
    return np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))


def test_emul_forward_and_backward_golden(eph, fmt, golden):
    g = golden[fmt]
    st = cases.pp_case()
    b = coop_emul.EmulBatch(eph, st.shape[0], forces=0x7F)
    b.set_state(cases.T0, st)
    b.integrate(cases.T0 + cases.PP_DAYS)
    got = b.get_state()
    assert np.array_equal(got["state"], g["pp_final"])
    assert np.array_equal(got["t"], g["pp_t"]) and np.array_equal(got["dt"], g["pp_dt"])
    c = got["counters"].sum(axis=0)
    assert [int(c[0]), int(c[2]), int(c[3]), int(c[1])] == [int(v) for v in g["pp_counts"]]
    assert (got["status"] == 0).all()
    b.close()
    b = coop_emul.EmulBatch(eph, 6, forces=0x7F)
    b.set_state(cases.T0, st[:6])
    b.integrate(cases.T0 - 300.0)
    got = b.get_state()
    assert np.array_equal(got["state"], g["ppback_final"]) and np.array_equal(got["t"], g["ppback_t"])
    b.close()


def test_emul_rejected_steps_many_blocks_two_calls(eph, fmt, ref, paths):
    """NEOs with min_dt and a first step far too long (rejected attempts), more systems than one CTA holds, three CTAs
    racing for the queue, and a second integrate() call that continues the first."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    n = 75
    st = populations.neo_mba_mix(n, seed=20261703)
    b = coop_emul.EmulBatch(eph, n, n_blocks=3, forces=0x7F, min_dt=1e-3)
    b.set_state(cases.T0, st, dt0=400.0)
    b.integrate(cases.T0 + 250.0)
    b.integrate(cases.T0 + 700.0)
    got = b.get_state()
    want = np.empty((n, 1, 6)); wt = np.empty(n); wdt = np.empty(n)
    tot = dict(steps=0, pc_iterations=0, force_evals=0, rejected=0)
    for i in range(n):
        s = rh.Sim(ref, reph, cases.T0, st[i:i + 1], forces=0x7F, min_dt=1e-3, dt0=400.0)
        s.integrate(cases.T0 + 250.0)
        s.integrate(cases.T0 + 700.0)
        want[i] = s.state()[0]; wt[i] = s.t; wdt[i] = s.dt
        c = s.counters()
        for k in tot:
            tot[k] += c[k]
        s.close()
    assert np.array_equal(got["state"], want) and np.array_equal(got["t"], wt) and np.array_equal(got["dt"], wdt)
    c = got["counters"].sum(axis=0)
    assert [int(c[0]), int(c[2]), int(c[3]), int(c[1])] == [tot["steps"], tot["pc_iterations"], tot["force_evals"], tot["rejected"]]
    assert tot["rejected"] > 0, "the case is meant to exercise the rejected-step path"
    b.close()


@pytest.mark.parametrize("mask", [0x01, 0x07, 0x37, 0x1F7, 0xB7])
def test_emul_force_masks(eph, fmt, ref, paths, mask):
    """Every term the kernel schedules as a task, switched on and off (0x100 potential GR, 0x80 simple GR)."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    st = populations.neo_mba_mix(5, seed=5)
    b = coop_emul.EmulBatch(eph, 5, forces=mask)
    b.set_state(cases.T0, st)
    b.integrate(cases.T0 + 120.0)
    want, wt, wdt, _ = rh.integrate_each(ref, reph, cases.T0, st, cases.T0 + 120.0, forces=mask)
    got = b.get_state()
    assert np.array_equal(got["state"], want) and np.array_equal(got["t"], wt) and np.array_equal(got["dt"], wdt)
    b.close()


def test_emul_dense_output(eph, fmt, ref, paths, golden):
    """assist_integrate_or_interpolate semantics: forward, backward and back-and-forth epoch lists without the Marsden
    term are bit-identical; the comet case (pow() in the path) within 1e-12 AU of the golden vectors."""
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    st = populations.main_belt(7, seed=72)
    for times in (cases.T0 + np.array([5.0, 20.0, 21.0, 60.0, 61.0, 200.0]),
                  cases.T0 - np.array([0.0, 3.0, 10.0, 10.5, 11.0, 40.0, 41.0]),
                  cases.T0 + np.array([80.0, 30.0, 31.0, 70.0])):
        b = coop_emul.EmulBatch(eph, 7, forces=0x7F)
        b.set_state(cases.T0, st)
        out = b.integrate_or_interpolate(times)
        want = rh.dense_each(ref, reph, cases.T0, st, times, forces=0x7F)
        assert same(out, want)
        b.close()
    stc, prm = cases.comet_case()
    b = coop_emul.EmulBatch(eph, stc.shape[0], forces=0x7F)
    b.set_state(cases.T0, stc, params=prm)
    out = b.integrate_or_interpolate(cases.DENSE_TIMES)
    assert np.nanmax(np.linalg.norm(out[..., :3] - golden[fmt]["dense"][..., :3], axis=-1)) <= 1e-12
    assert np.isfinite(out).all()
    b.close()


def test_emul_errors_retire_a_system(eph, fmt):
    """Leaving the ephemeris coverage, a zero timestep and an exhausted step budget end one system with an error
    status; the others finish (a stuck particle must not hold the kernel)."""
    st = populations.main_belt(3, seed=9)
    tb, te = eph.time_bounds()
    b = coop_emul.EmulBatch(eph, 3, forces=0x7F)
    b.set_state(te - 30.0, st)
    b.integrate(te + 50.0)
    got = b.get_state()
    assert (got["status"] == 1005).all() and (got["t"] < te).all()
    b.close()
    b = coop_emul.EmulBatch(eph, 3, forces=0x7F)
    b.set_state(cases.T0, st, dt0=0.0)
    b.integrate(cases.T0 + 10.0)
    assert (b.get_state()["status"] == 1007).all()
    b.close()
    b = coop_emul.EmulBatch(eph, 3, budget=5, forces=0x7F)
    b.set_state(cases.T0, st)
    b.integrate(cases.T0 + 1000.0)
    got = b.get_state()
    assert (got["status"] == 1007).all() and (got["counters"][:, 0] + got["counters"][:, 1] == 5).all()
    b.close()


def test_emul_c3_subsample_ten_years(eph, fmt):
    """The kernel's source on the host over the north star's span: every 25th particle of the fixed C3 subsample
    (8 NEOs, 32 main-belt objects), 3652.5 d, min_dt 1e-3, against the reference's own output in golden_large.npz --
    states, t, dt and the per-particle step / sweep / evaluation / rejection counts, bit for bit."""
    import os
    from conftest import ROOT
    if fmt != "bsp":
        pytest.skip("golden_large.npz holds the SPK planets file only")
    G = np.load(os.path.join(ROOT, "tests", "golden", "golden_large.npz"))
    pick = np.arange(0, 1000, 25)
    st = cases.c3_subsample()[pick]
    b = coop_emul.EmulBatch(eph, st.shape[0], n_blocks=1, forces=0x7F, min_dt=1e-3)
    b.set_state(cases.T0, st)
    b.integrate(cases.T0 + 3652.5)
    got = b.get_state()
    assert np.array_equal(got["state"], G["c3_final"][pick])
    assert np.array_equal(got["t"], G["c3_t"][pick]) and np.array_equal(got["dt"], G["c3_dt"][pick])
    c = got["counters"].astype(np.int64)          # steps, rejected, iters, evals
    assert np.array_equal(np.stack([c[:, 0], c[:, 2], c[:, 3], c[:, 1]], axis=1), G["c3_counts"][pick])
    b.close()


@pytest.mark.parametrize("direction", [1.0, -1.0])
def test_emul_steps_across_the_segment_boundary(eph, fmt, ref, paths, direction):
    """The synthetic kernels hold two segments per target (boundary at JD 2453000.5, t = 1455.5): steps whose eight nodes
    straddle it read their records from two segments (the staged fill then leaves such a series to the global path),
    forward and backward, against one reference simulation per particle."""
    if fmt != "bsp":
        pytest.skip("segments are an SPK notion")
    reph = rh.open_ephem(ref, planets_path(paths, fmt), paths["asteroids_bsp"])
    tb = 1455.5
    st = populations.neo_mba_mix(12, seed=4242)
    t0, t1 = tb - direction * 35.0, tb + direction * 70.0
    b = coop_emul.EmulBatch(eph, st.shape[0], forces=0x7F)
    b.set_state(t0, st)
    b.integrate(t1)
    got = b.get_state()
    b.close()
    want, wt, wdt, tot = rh.integrate_each(ref, reph, t0, st, t1, forces=0x7F)
    assert np.array_equal(got["state"], want) and np.array_equal(got["t"], wt) and np.array_equal(got["dt"], wdt)
    assert int(got["counters"][:, 0].sum()) == tot["steps"]
