"""Drive a libassist-compatible library (the oracle build oracle/_ref/libassist_ref.so, or
the product library through the very same entry points) from Python.

TEST INFRASTRUCTURE.  Everything here goes through the reference's public C API:
assist_ephem_create, assist_attach, reb_simulation_add, reb_simulation_integrate,
reb_simulation_update_acceleration, assist_integrate_or_interpolate, assist_all_ephem.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_double, c_int

import numpy as np

from assist_b200.cstructs import Extras, Particle, Simulation, bind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libassist_ref.so")

_ref = None


def have_ref():
    return os.path.exists(REF_LIB)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = bind(ctypes.CDLL(REF_LIB))
    return _ref


def open_ephem(lib, planets, asteroids):
    eph = lib.assist_ephem_create(str(planets).encode(), str(asteroids).encode() if asteroids else None)
    if not eph:
        raise RuntimeError("assist_ephem_create failed")
    return eph


def all_bodies(lib, eph, times, nbodies=27):
    """out[n_t][nbodies][10] (GM x y z vx vy vz ax ay az) through assist_all_ephem without cache."""
    times = np.atleast_1d(times)
    out = np.empty((times.size, nbodies, 10))
    st = np.zeros((times.size, nbodies), dtype=np.int32)
    v = [c_double() for _ in range(10)]
    for it, t in enumerate(times):
        for b in range(nbodies):
            for q in v:
                q.value = np.nan
            st[it, b] = lib.assist_all_ephem(eph, None, b, float(t), *[byref(q) for q in v])
            out[it, b, :] = [q.value for q in v]
    return out, st


class Sim:
    """A reb_simulation with ASSIST attached, systems laid out as [real..., variational...]."""

    def __init__(self, lib, eph, t0, state, params=None, forces=None, gr_eih_sources=1, geocentric=0,
                 epsilon=1e-9, min_dt=0.0, dt0=0.001, ng=None):
        self.lib = lib
        state = np.asarray(state, dtype=np.float64)
        if state.ndim == 2:
            state = state[:, None, :]
        n, K = state.shape[0], state.shape[1]
        self.n, self.K = n, K
        self.r = lib.reb_simulation_create()
        self.ax = lib.assist_attach(self.r, eph)
        r = self.r.contents
        r.t = float(t0)
        r.dt = float(dt0)
        r.ri_ias15.epsilon = epsilon
        r.ri_ias15.min_dt = min_dt
        for i in range(n):
            s = state[i, 0]
            lib.reb_simulation_add(self.r, Particle(x=s[0], y=s[1], z=s[2], vx=s[3], vy=s[4], vz=s[5]))
        self.var_index = np.full((n, max(K - 1, 0)), -1, dtype=np.int64)
        for i in range(n):
            for v in range(K - 1):
                idx = lib.reb_simulation_add_variation_1st_order(self.r, i)
                self.var_index[i, v] = idx
                p = self.r.contents.particles[idx]
                s = state[i, 1 + v]
                p.x, p.y, p.z, p.vx, p.vy, p.vz = [float(q) for q in s]
        ax = self.ax.contents
        if forces is not None:
            ax.forces = int(forces)
        ax.gr_eih_sources = gr_eih_sources
        ax.geocentric = geocentric
        if ng is not None:
            ax.alpha, ax.nk, ax.nm, ax.nn, ax.r0 = ng
        self._params = None
        if params is not None:
            params = np.asarray(params, dtype=np.float64).reshape(n, K, 3)
            flat = np.zeros((n + n * (K - 1), 3))
            flat[:n] = params[:, 0]
            # variational rows are indexed N_real + (var_config index) -- reference src/forces.c:1030-1032
            for i in range(n):
                for v in range(K - 1):
                    flat[n + i * (K - 1) + v] = params[i, 1 + v]
            self._params = np.ascontiguousarray(flat)
            ax.particle_params = self._params.ctypes.data_as(POINTER(c_double))

    def _pidx(self, i, j):
        return i if j == 0 else int(self.var_index[i, j - 1])

    def state(self):
        out = np.empty((self.n, self.K, 6))
        P = self.r.contents.particles
        for i in range(self.n):
            for j in range(self.K):
                p = P[self._pidx(i, j)]
                out[i, j] = (p.x, p.y, p.z, p.vx, p.vy, p.vz)
        return out

    def acc(self):
        out = np.empty((self.n, self.K, 3))
        P = self.r.contents.particles
        for i in range(self.n):
            for j in range(self.K):
                p = P[self._pidx(i, j)]
                out[i, j] = (p.ax, p.ay, p.az)
        return out

    def update_acceleration(self):
        self.lib.reb_simulation_update_acceleration(self.r)
        return self.acc()

    def integrate(self, t, exact_finish_time=1):
        self.r.contents.exact_finish_time = exact_finish_time
        return self.lib.reb_simulation_integrate(self.r, float(t))

    def integrate_or_interpolate(self, t):
        self.lib.assist_integrate_or_interpolate(self.ax, float(t))
        return self.state()

    @property
    def t(self):
        return self.r.contents.t

    @property
    def dt(self):
        return self.r.contents.dt

    @property
    def dt_last_done(self):
        return self.r.contents.dt_last_done

    def counters(self):
        r = self.r.contents
        return dict(steps=int(r.steps_done), pc_iterations=int(r.ri_ias15.b200_pc_iterations),
                    force_evals=int(r.ri_ias15.b200_force_evals), rejected=int(r.ri_ias15.b200_steps_rejected))

    def close(self):
        if self.r:
            self.lib.assist_free(self.ax)
            self.lib.reb_simulation_free(self.r)
            self.r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def forces(lib, eph, t, state, params=None, **kw):
    """One force evaluation for all systems at a common time: acc[n][K][3]."""
    s = Sim(lib, eph, t, state, params=params, **kw)
    a = s.update_acceleration()
    s.close()
    return a


def integrate_each(lib, eph, t0, state, t_end, params=None, exact_finish_time=1, **kw):
    """One simulation per system (per-particle-dt semantics). Returns final state [n][K][6], t, dt, counters."""
    state = np.asarray(state, dtype=np.float64)
    if state.ndim == 2:
        state = state[:, None, :]
    n, K = state.shape[:2]
    out = np.empty_like(state)
    ts = np.empty(n); dts = np.empty(n)
    tot = dict(steps=0, pc_iterations=0, force_evals=0, rejected=0)
    for i in range(n):
        p = None if params is None else np.asarray(params).reshape(n, K, 3)[i:i + 1]
        s = Sim(lib, eph, t0, state[i:i + 1], params=p, **kw)
        s.integrate(t_end, exact_finish_time)
        out[i] = s.state()[0]
        ts[i] = s.t; dts[i] = s.dt
        c = s.counters()
        for k in tot:
            tot[k] += c[k]
        s.close()
    return out, ts, dts, tot


def dense_each(lib, eph, t0, state, times, params=None, **kw):
    """assist_integrate_or_interpolate at every epoch, one simulation per system: out[n_t][n][K][6]."""
    state = np.asarray(state, dtype=np.float64)
    if state.ndim == 2:
        state = state[:, None, :]
    n, K = state.shape[:2]
    out = np.empty((len(times), n, K, 6))
    for i in range(n):
        p = None if params is None else np.asarray(params).reshape(n, K, 3)[i:i + 1]
        s = Sim(lib, eph, t0, state[i:i + 1], params=p, **kw)
        for e, t in enumerate(times):
            out[e, i] = s.integrate_or_interpolate(t)[0]
        s.close()
    return out
