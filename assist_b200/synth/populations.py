"""Synthetic particle populations for the benchmark configurations.

All generators are seeded and return barycentric equatorial states [n][6]
(AU, AU/day) at t0 = JD 2460000.5 (t = 8455.5 relative to jd_ref = 2451545.0).
"""
from __future__ import annotations

import numpy as np

from .ephem_writer import CONSTANTS, JD_J2000, OBLIQUITY, SolarSystemModel, AU_KM

T0_JD = 2460000.5
T0 = T0_JD - JD_J2000          # 8455.5
GMS = CONSTANTS["GMS"]


def _kepler_E(M, e):
    M = np.mod(M + np.pi, 2 * np.pi) - np.pi
    E = np.where(e < 0.8, M + e * np.sin(M), np.pi * np.sign(M + 1e-300))
    E = np.where(np.abs(M) < 1e-12, M, E)
    for _ in range(100):
        dE = (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
        E = E - dE
        if np.max(np.abs(dE)) < 1e-14:
            break
    return E


def elements_to_state(a, e, inc, node, argp, M, gm=GMS):
    """Heliocentric ecliptic elements (rad) -> heliocentric EQUATORIAL state [n][6]."""
    a, e, inc, node, argp, M = [np.asarray(v, dtype=np.float64) for v in (a, e, inc, node, argp, M)]
    E = _kepler_E(M, e)
    cE, sE = np.cos(E), np.sin(E)
    b = a * np.sqrt(1 - e * e)
    n = np.sqrt(gm / a ** 3)
    xp = a * (cE - e)
    yp = b * sE
    Edot = n / (1 - e * cE)
    vxp = -a * sE * Edot
    vyp = b * cE * Edot
    cw, sw = np.cos(argp), np.sin(argp)
    cO, sO = np.cos(node), np.sin(node)
    ci, si = np.cos(inc), np.sin(inc)
    R11 = cO * cw - sO * sw * ci; R12 = -cO * sw - sO * cw * ci
    R21 = sO * cw + cO * sw * ci; R22 = -sO * sw + cO * cw * ci
    R31 = sw * si; R32 = cw * si
    x = R11 * xp + R12 * yp; y = R21 * xp + R22 * yp; z = R31 * xp + R32 * yp
    vx = R11 * vxp + R12 * vyp; vy = R21 * vxp + R22 * vyp; vz = R31 * vxp + R32 * vyp
    ce, se = np.cos(OBLIQUITY), np.sin(OBLIQUITY)
    return np.stack([x, ce * y - se * z, se * y + ce * z, vx, ce * vy - se * vz, se * vy + ce * vz], axis=-1)


def _sun_state(t0_jd=T0_JD):
    m = SolarSystemModel()
    p = m.sun_bary(np.array([t0_jd]))[:, 0] / AU_KM
    h = 0.01
    v = (m.sun_bary(np.array([t0_jd + h]))[:, 0] - m.sun_bary(np.array([t0_jd - h]))[:, 0]) / (2 * h) / AU_KM
    return np.concatenate([p, v])


def _to_bary(helio):
    return helio + _sun_state()[None, :]


def _planet_positions(t0_jd=T0_JD):
    """Barycentric positions (AU) and Hill radii (AU) of the nine planet barycentres at t0."""
    m = SolarSystemModel()
    pos, hill = [], []
    for name, el in m.planets.items():
        pos.append(m.position(name, np.array([t0_jd]))[:, 0] / AU_KM)
        hill.append(el["a"] * (el["gm"] / (3.0 * GMS)) ** (1.0 / 3.0))
    return np.array(pos), np.array(hill)


def _outside_hill_spheres(state, factor=3.0):
    """True for particles that start farther than `factor` Hill radii from every planet.

    Random orbital elements occasionally put a particle INSIDE a planet's Hill sphere with a
    small relative velocity, i.e. on a satellite orbit (period ~ hours).  Such an object is not
    a heliocentric small body; it is rejected at generation time (documented in DESIGN.md)."""
    pos, hill = _planet_positions()
    d = np.linalg.norm(state[:, None, :3] - pos[None, :, :], axis=-1)
    return np.all(d > factor * hill[None, :], axis=1)


def _filtered(gen, n, seed):
    """Draw from gen(m, seed) until n particles pass the Hill-sphere filter (deterministic)."""
    out = np.empty((0, 6))
    k = 0
    while out.shape[0] < n:
        cand = gen(n - out.shape[0] + 64, seed + 7919 * k)
        out = np.concatenate([out, cand[_outside_hill_spheres(cand)]], axis=0)
        k += 1
    return out[:n]


def main_belt(n, seed=20261702):
    """C2 / C4: a~U[2.1,3.3], e~Rayleigh(0.1) clipped 0.3, i~Rayleigh(8 deg)."""
    return _filtered(_main_belt_raw, n, seed)


def _main_belt_raw(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(2.1, 3.3, n)
    e = np.clip(rng.rayleigh(0.1, n), 0.0, 0.3)
    inc = np.deg2rad(np.clip(rng.rayleigh(8.0, n), 0.0, 60.0))
    node, argp, M = (rng.uniform(0, 2 * np.pi, n) for _ in range(3))
    return _to_bary(elements_to_state(a, e, inc, node, argp, M))


def neo(n, seed=20261703):
    """NEO part of C3: a~U[0.8,2.5], e~U[0.2,0.7], q<1.3, i~Rayleigh(12 deg)."""
    return _filtered(_neo_raw, n, seed)


def _neo_raw(n, seed):
    rng = np.random.default_rng(seed)
    a = np.empty(n); e = np.empty(n)
    filled = 0
    while filled < n:
        aa = rng.uniform(0.8, 2.5, 2 * (n - filled) + 16)
        ee = rng.uniform(0.2, 0.7, aa.size)
        ok = aa * (1 - ee) < 1.3
        k = min(n - filled, int(ok.sum()))
        a[filled:filled + k] = aa[ok][:k]; e[filled:filled + k] = ee[ok][:k]
        filled += k
    inc = np.deg2rad(np.clip(rng.rayleigh(12.0, n), 0.0, 80.0))
    node, argp, M = (rng.uniform(0, 2 * np.pi, n) for _ in range(3))
    return _to_bary(elements_to_state(a, e, inc, node, argp, M))


def neo_mba_mix(n, seed=20261703):
    """C3: 20% NEOs followed by 80% main-belt objects."""
    n_neo = n // 5
    return np.concatenate([neo(n_neo, seed), main_belt(n - n_neo, seed + 1000)], axis=0)


def comets(n, seed=20261705):
    """C5: q~U[0.5,3], e~U[0.6,0.98], isotropic, perihelion within +-10 yr of t0; Marsden A1..A3."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(0.5, 3.0, n)
    e = rng.uniform(0.6, 0.98, n)
    a = q / (1 - e)
    inc = np.arccos(rng.uniform(-1, 1, n))
    node, argp = (rng.uniform(0, 2 * np.pi, n) for _ in range(2))
    nmot = np.sqrt(GMS / a ** 3)
    M = nmot * rng.uniform(-3652.5, 3652.5, n)
    state = _to_bary(elements_to_state(a, e, inc, node, argp, M))
    A1 = 10 ** rng.uniform(-9, -8, n)
    A2 = rng.choice([-1.0, 1.0], n) * 10 ** rng.uniform(-10, -9, n)
    A3 = rng.choice([-1.0, 1.0], n) * 10 ** rng.uniform(-11, -10, n)
    return state, np.stack([A1, A2, A3], axis=-1)


def apophis_like():
    """C1: a=0.9224, e=0.1914, i=3.34, node=204.0, argp=126.7, M=142.9 deg; A1=5e-13, A2=-2.9e-14."""
    st = _to_bary(elements_to_state([0.9224], [0.1914], [np.deg2rad(3.34)], [np.deg2rad(204.0)],
                                    [np.deg2rad(126.7)], [np.deg2rad(142.9)]))
    return st, np.array([[5e-13, -2.9e-14, 0.0]])


def with_variations(state, n_var=6):
    """[n][6] -> [n][1+n_var][6] with the variational particles set to the 6x6 identity columns."""
    n = state.shape[0]
    out = np.zeros((n, 1 + n_var, 6))
    out[:, 0, :] = state
    for v in range(min(n_var, 6)):
        out[:, 1 + v, v] = 1.0
    return out
