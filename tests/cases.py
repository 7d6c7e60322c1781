"""Seeded test cases shared by the golden generator and the parity tests."""
import numpy as np

from assist_b200.synth import populations

T0 = populations.T0

# times relative to jd_ref: inside, near both ends of coverage, on record boundaries (8455.5+200.0 hits a
# 16-day and an 8-day boundary of the synthetic files), and one negative
EPHEM_TIMES = np.array([T0, T0 + 17.3, T0 - 1234.56789, T0 + 3000.25, -10000.0, T0 + 200.0, T0 + 199.99999, 13455.0, -10544.5])

FORCE_T = T0 + 3.7
# name, mask, gr_eih_sources, geocentric
FORCE_TERMS = [
    ("sun", 0x01, 1, 0), ("planets", 0x02, 1, 0), ("asteroids", 0x04, 1, 0), ("nongrav", 0x08, 1, 0),
    ("earth_harm", 0x10, 1, 0), ("sun_harm", 0x20, 1, 0), ("eih1", 0x40, 1, 0), ("eih11", 0x40, 11, 0),
    ("gr_simple", 0x80, 1, 0), ("gr_potential", 0x100, 1, 0), ("default", 0x7F, 1, 0), ("all11", 0x7F, 11, 0),
    ("geocentric", 0x7F, 1, 1),
]

PP_DAYS = 400.0
VAR_DAYS = 200.0
SH_DAYS = 300.0
DENSE_TIMES = T0 - 10.0 * np.arange(1, 40)


def force_case(n=64):
    """Systems with 6 variational particles (random variations) and non-zero A1..A3 / dA."""
    st6 = populations.neo_mba_mix(n, seed=5)
    state = populations.with_variations(st6, 6)
    rng = np.random.default_rng(3)
    state[:, 1:, :] += 0.1 * rng.standard_normal(state[:, 1:, :].shape)
    params = np.zeros((n, 7, 3))
    params[:, 0, :] = [1e-9, -2e-10, 3e-11]
    params[:, 1:, :] = rng.standard_normal((n, 6, 3))
    # a few particles without non-gravitational parameters (reference skips them, src/forces.c:861)
    params[::7, 0, :] = 0.0
    return state, params


def pp_case(n=16):
    return populations.neo_mba_mix(n, seed=11)


def var_case(n=8):
    return populations.with_variations(populations.main_belt(n, seed=13), 6)


def shared_case(n=40):
    return populations.main_belt(n, seed=17)


def shared_var_case():
    """Two real particles, the first with two variational particles, the second with one
    (layout of reference unit_tests/variational_spk/problem.c, extended)."""
    st = populations.main_belt(2, seed=23)
    out = np.zeros((2, 3, 6))
    out[:, 0, :] = st
    out[0, 1, 0] = 1.0
    out[0, 2, 4] = 1.0
    out[1, 1, 2] = 1.0
    return out


def comet_case(n=6):
    return populations.comets(n, seed=19)


# ---- north_star sizes (tests/golden/make_golden_large.py, tests/test_gpu_large.py) ----------------------------
# fixed subsamples of the populations bench.py runs (same generators, same seeds)
C5_EPOCHS = T0 - 10.0 * np.arange(1, 1828)          # 1827 epochs, every 10 d, 50 yr backward
C5_SPARSE = np.arange(0, 1827, 29)


def c3_subsample():
    """1000 of the 10^6 NEO+MBA particles of BASELINE config 3: every 1000th (200 NEOs, 800 main-belt)."""
    return populations.neo_mba_mix(1000000, seed=20261703)[::1000].copy()


def c4_subsample():
    """256 of the 10^5 systems of BASELINE config 4 (real particle + 6 variational)."""
    return populations.with_variations(populations.main_belt(100000, seed=20261704)[::390][:256], 6)


def c5_subsample():
    st, prm = populations.comets(100000, seed=20261705)
    return st[::390][:256].copy(), prm[::390][:256].copy()


def c2_population():
    return populations.main_belt(10000, seed=20261702)


def eih11_case():
    return populations.neo_mba_mix(64, seed=31)


def geocentric_case(earth_state):
    """12 near-Earth particles as GEOCENTRIC states: barycentric state minus Earth's at T0
    (earth_state(t) -> [x y z vx vy vz] of ASSIST body 3)."""
    st = populations.neo(12, seed=37)
    return st - np.asarray(earth_state(T0))[None, :]


# ---- small-body kernels with more than 16 targets (tests/golden/make_golden_n373.py) ----
N373_SIZES = (40, 373)


def n373_times():
    return np.array([T0, T0 + 17.3, T0 - 1234.56789, T0 + 3000.25, T0 + 199.99999])


def n373_force_systems():
    """6 systems, real particle + 2 variational particles (the direct term's Jacobians see every asteroid too)."""
    st = populations.neo_mba_mix(6, seed=373)
    return populations.with_variations(st, 2)


def n373_pp_particles():
    return populations.neo_mba_mix(8, seed=374)


def n373_shared_systems():
    return populations.with_variations(populations.neo_mba_mix(6, seed=375), 1)
