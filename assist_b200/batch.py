"""Host-side Python mirror of the GPU batch API (include/assist_gpu.h).

`EphemHandle` wraps assist_ephem_create (reference src/assist.c:350-360; Python mirror
assist/ephem.py:38-58).  `Batch` drives whole populations through the CUDA stepper:
`integrate(t)` is reb_simulation_integrate for every system, and
`integrate_or_interpolate(times)` is assist_integrate_or_interpolate for a sorted list
of epochs (reference src/assist.c:642-680).  numpy is used for buffers only.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_int

import numpy as np

from . import _lib
from ._lib import GpuOptions, GpuStats
from .cstructs import ASSIST_FORCES

SHARED_STEP = 0
PER_PARTICLE = 1
MATH_STRICT = 0
MATH_FAST = 1


def _dp(a):
    return a.ctypes.data_as(POINTER(c_double))


def _check(lib, rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib.assist_gpu_last_error().decode()))


class EphemHandle:
    """Owns a `struct assist_ephem*`."""

    def __init__(self, planets_path, asteroids_path=None):
        self.lib = _lib.load()
        p = str(planets_path).encode()
        a = str(asteroids_path).encode() if asteroids_path is not None else None
        self.ptr = self.lib.assist_ephem_create(p, a)
        if not self.ptr:
            raise RuntimeError("assist_ephem_create failed for %s / %s" % (planets_path, asteroids_path))

    @property
    def struct(self):
        return self.ptr.contents

    @property
    def nbodies(self):
        return self.lib.assist_gpu_ephem_nbodies(self.ptr)

    def time_bounds(self):
        tb, te = c_double(), c_double()
        self.lib.assist_ephem_time_bounds(self.ptr, byref(tb), byref(te))
        return tb.value, te.value

    def eval(self, times, math=MATH_STRICT):
        """Body states at `times` (relative to jd_ref): (out[n_t][nbodies][10], status[n_t][nbodies])."""
        t = np.ascontiguousarray(np.atleast_1d(times), dtype=np.float64)
        nb = self.nbodies
        out = np.empty((t.size, nb, 10), dtype=np.float64)
        st = np.empty((t.size, nb), dtype=np.int32)
        rc = self.lib.assist_gpu_ephem_eval(self.ptr, math, _dp(t), t.size, _dp(out), st.ctypes.data_as(POINTER(c_int)))
        _check(self.lib, rc, "assist_gpu_ephem_eval")
        return out, st

    def close(self):
        if self.ptr:
            self.lib.assist_ephem_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_options(lib, forces=None, gr_eih_sources=1, geocentric=0, math=MATH_STRICT, epsilon=1e-9, min_dt=0.0,
                 alpha=1.0, nk=0.0, nm=2.0, nn=5.093, r0=1.0):
    opt = GpuOptions()
    lib.assist_gpu_default_options(byref(opt))
    if forces is not None:
        if isinstance(forces, (list, tuple)):
            m = 0
            for f in forces:
                m |= ASSIST_FORCES[f.upper()]
            forces = m
        opt.forces = int(forces)
    opt.gr_eih_sources = gr_eih_sources
    opt.geocentric = geocentric
    opt.math = math
    opt.epsilon = epsilon
    opt.min_dt = min_dt
    opt.alpha, opt.nk, opt.nm, opt.nn, opt.r0 = alpha, nk, nm, nn, r0
    return opt


def eval_forces(ephem, t, state, params=None, **opts):
    """One force evaluation per system.  state[n][K][6]; t scalar or [n]; returns acc[n][K][3]."""
    lib = ephem.lib
    state = np.ascontiguousarray(state, dtype=np.float64)
    n, K = state.shape[0], state.shape[1]
    tt = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
    per = 1 if tt.size == n and np.ndim(t) > 0 else 0
    acc = np.empty((n, K, 3), dtype=np.float64)
    st = np.empty(n, dtype=np.int32)
    opt = make_options(lib, **opts)
    prm = None
    if params is not None:
        prm = np.ascontiguousarray(params, dtype=np.float64)
        assert prm.shape == (n, K, 3)
    rc = lib.assist_gpu_eval_forces(ephem.ptr, byref(opt), n, K - 1, _dp(tt), per, _dp(state),
                                    _dp(prm) if prm is not None else None, _dp(acc),
                                    st.ctypes.data_as(POINTER(c_int)))
    _check(lib, rc, "assist_gpu_eval_forces")
    return acc


class Batch:
    """A population of systems (real particle + n_var variational particles) on one GPU."""

    def __init__(self, ephem, n_sys, n_var=0, mode=PER_PARTICLE, **opts):
        self.lib = ephem.lib
        self.ephem = ephem
        self.n, self.n_var, self.K, self.mode = int(n_sys), int(n_var), 1 + int(n_var), mode
        self.ptr = self.lib.assist_gpu_batch_create(ephem.ptr, self.n, self.n_var, mode)
        if not self.ptr:
            raise RuntimeError("assist_gpu_batch_create: " + self.lib.assist_gpu_last_error().decode())
        self.set_options(**opts)

    def set_options(self, **opts):
        self.opt = make_options(self.lib, **opts)
        _check(self.lib, self.lib.assist_gpu_batch_set_options(self.ptr, byref(self.opt)), "set_options")

    def set_state(self, t0, state, params=None, dt0=0.001, nvar_per_system=None):
        state = np.ascontiguousarray(state, dtype=np.float64).reshape(self.n, self.K, 6)
        prm = None
        if params is not None:
            prm = np.ascontiguousarray(params, dtype=np.float64).reshape(self.n, self.K, 3)
        nv = None
        if nvar_per_system is not None:
            nv = np.ascontiguousarray(nvar_per_system, dtype=np.int32)
        rc = self.lib.assist_gpu_batch_set_state(self.ptr, float(t0), float(dt0), _dp(state),
                                                 _dp(prm) if prm is not None else None,
                                                 nv.ctypes.data_as(POINTER(c_int)) if nv is not None else None)
        _check(self.lib, rc, "set_state")

    def snapshot(self):
        _check(self.lib, self.lib.assist_gpu_batch_snapshot(self.ptr), "snapshot")

    def restore(self):
        _check(self.lib, self.lib.assist_gpu_batch_restore(self.ptr), "restore")

    def integrate(self, t_end, exact_finish_time=1, max_steps=0):
        rc = self.lib.assist_gpu_batch_integrate(self.ptr, float(t_end), int(exact_finish_time), int(max_steps))
        _check(self.lib, rc, "integrate")

    def integrate_or_interpolate(self, times):
        times = np.ascontiguousarray(times, dtype=np.float64)
        out = np.empty((times.size, self.n, self.K, 6), dtype=np.float64)
        rc = self.lib.assist_gpu_batch_integrate_or_interpolate(self.ptr, _dp(times), times.size, _dp(out))
        _check(self.lib, rc, "integrate_or_interpolate")
        return out

    def get_state(self, with_acc=False):
        state = np.empty((self.n, self.K, 6), dtype=np.float64)
        acc = np.empty((self.n, self.K, 3), dtype=np.float64) if with_acc else None
        m = self.n if self.mode == PER_PARTICLE else 1
        t = np.empty(m); dt = np.empty(m); dtl = np.empty(m)
        st = np.empty(m, dtype=np.int32)
        rc = self.lib.assist_gpu_batch_get_state(self.ptr, _dp(state), _dp(acc) if with_acc else None, _dp(t), _dp(dt),
                                                 _dp(dtl), st.ctypes.data_as(POINTER(c_int)))
        _check(self.lib, rc, "get_state")
        res = dict(state=state, t=t, dt=dt, dt_last_done=dtl, status=st)
        if with_acc:
            res["acc"] = acc
        return res

    def stats(self):
        s = GpuStats()
        _check(self.lib, self.lib.assist_gpu_batch_get_stats(self.ptr, byref(s)), "get_stats")
        return dict(steps=s.steps, steps_rejected=s.steps_rejected, pc_iterations=s.pc_iterations,
                    force_evals=s.force_evals, kernel_launches=s.kernel_launches, last_kernel_ms=s.last_kernel_ms)

    def counters(self):
        """Per-system counters (per-particle batches): dict of uint64 arrays of length n."""
        out = {k: np.zeros(self.n, dtype=np.uint64) for k in ("steps", "rejected", "iters", "evals")}
        u64 = POINTER(ctypes.c_ulonglong)
        rc = self.lib.assist_gpu_batch_get_counters(self.ptr, *[out[k].ctypes.data_as(u64) for k in ("steps", "rejected", "iters", "evals")])
        _check(self.lib, rc, "get_counters")
        return out

    def close(self):
        if self.ptr:
            self.lib.assist_gpu_batch_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiBatch:
    """One per-particle population over several GPUs of a box (include/assist_gpu.h: assist_gpu_multi_*): one
    sub-batch and one host thread per device, no collective, a host gather of outputs.  devices=None: every visible
    device.  Results are those of a single-device Batch, bit for bit."""

    def __init__(self, ephem, n_sys, n_var=0, devices=None, **opts):
        self.lib = ephem.lib
        self.ephem = ephem
        self.n, self.n_var, self.K = int(n_sys), int(n_var), 1 + int(n_var)
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        self.ptr = self.lib.assist_gpu_multi_create(ephem.ptr, self.n, self.n_var,
                                                    dev.ctypes.data_as(POINTER(c_int)) if dev is not None else None,
                                                    0 if dev is None else dev.size)
        if not self.ptr:
            raise RuntimeError("assist_gpu_multi_create: " + self.lib.assist_gpu_multi_last_error().decode())
        self.n_devices = self.lib.assist_gpu_multi_device_count(self.ptr)
        self.opt = make_options(self.lib, **opts)
        self._check(self.lib.assist_gpu_multi_set_options(self.ptr, byref(self.opt)), "set_options")

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.assist_gpu_multi_last_error().decode()))

    def set_state(self, t0, state, params=None, dt0=0.001):
        state = np.ascontiguousarray(state, dtype=np.float64).reshape(self.n, self.K, 6)
        prm = None if params is None else np.ascontiguousarray(params, dtype=np.float64).reshape(self.n, self.K, 3)
        self._check(self.lib.assist_gpu_multi_set_state(self.ptr, float(t0), float(dt0), _dp(state),
                                                        _dp(prm) if prm is not None else None), "set_state")

    def integrate(self, t_end, exact_finish_time=1):
        self._check(self.lib.assist_gpu_multi_integrate(self.ptr, float(t_end), int(exact_finish_time)), "integrate")

    def integrate_or_interpolate(self, times):
        times = np.ascontiguousarray(times, dtype=np.float64)
        out = np.empty((times.size, self.n, self.K, 6), dtype=np.float64)
        self._check(self.lib.assist_gpu_multi_integrate_or_interpolate(self.ptr, _dp(times), times.size, _dp(out)), "integrate_or_interpolate")
        return out

    def get_state(self):
        state = np.empty((self.n, self.K, 6)); t = np.empty(self.n); dt = np.empty(self.n); dtl = np.empty(self.n)
        st = np.empty(self.n, dtype=np.int32)
        self._check(self.lib.assist_gpu_multi_get_state(self.ptr, _dp(state), _dp(t), _dp(dt), _dp(dtl),
                                                        st.ctypes.data_as(POINTER(c_int))), "get_state")
        return dict(state=state, t=t, dt=dt, dt_last_done=dtl, status=st)

    def counters(self):
        out = {k: np.zeros(self.n, dtype=np.uint64) for k in ("steps", "rejected", "iters", "evals")}
        u64 = POINTER(ctypes.c_ulonglong)
        self._check(self.lib.assist_gpu_multi_get_counters(self.ptr, *[out[k].ctypes.data_as(u64) for k in ("steps", "rejected", "iters", "evals")]), "get_counters")
        return out

    def stats(self):
        s = GpuStats()
        per_dev = np.zeros(self.n_devices)
        self._check(self.lib.assist_gpu_multi_get_stats(self.ptr, byref(s), _dp(per_dev)), "get_stats")
        return dict(steps=s.steps, steps_rejected=s.steps_rejected, pc_iterations=s.pc_iterations, force_evals=s.force_evals,
                    kernel_launches=s.kernel_launches, last_kernel_ms=s.last_kernel_ms, kernel_ms_per_device=per_dev)

    def close(self):
        if self.ptr:
            self.lib.assist_gpu_multi_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
