"""Driver for tests/emul/libcoop_emul.so: the source of pp_coop_kernel run on the host (TEST INFRASTRUCTURE).

Builds the shared object on demand (g++, -ffp-contract=off) and offers the same calls as assist_b200.batch.Batch
for a per-particle population without variational particles."""
import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_double, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

import numpy as np

from assist_b200 import batch as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
CSRC = os.path.join(ROOT, "assist_b200", "csrc")


def build():
    so = os.path.join(EMUL_DIR, "libcoop_emul.so")
    srcs = [os.path.join(EMUL_DIR, "coop_emul.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread",
                               "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                               os.path.join(EMUL_DIR, "coop_emul.cpp"), "-o", so])
    return so


_emul = None


def lib():
    global _emul
    if _emul is None:
        e = ctypes.CDLL(build())
        e.abc_emul_sizeof.restype = c_size_t
        e.abc_emul_create.restype = c_void_p
        e.abc_emul_create.argtypes = [c_int, c_int]
        e.abc_emul_free.argtypes = [c_void_p]
        e.abc_emul_set_state.argtypes = [c_void_p, c_double, c_double, POINTER(c_double), POINTER(c_double), c_double, c_double]
        e.abc_emul_get_state.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_double), POINTER(c_double),
                                         POINTER(c_int), POINTER(c_ulonglong)]
        e.abc_emul_run.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int, POINTER(c_double), c_int,
                                   POINTER(c_double), c_int]
        _emul = e
    return _emul


def _dp(a):
    return a.ctypes.data_as(POINTER(c_double))


class EmulBatch:
    def __init__(self, ephem, n, n_blocks=1, budget=0, **opts):
        self.e = lib()
        self.plib = ephem.lib
        self.n, self.n_blocks = int(n), int(n_blocks)
        self.ephem = ephem
        self.opt = ab.make_options(self.plib, **opts)
        self.budget = budget
        self.h = self.e.abc_emul_create(self.n, self.n_blocks)
        szE, szF, szP = (self.e.abc_emul_sizeof(k) for k in (0, 1, 3))
        self.E = ctypes.create_string_buffer(szE)
        self.F = ctypes.create_string_buffer(szF)
        self.P = ctypes.create_string_buffer(szP)
        f = self.plib.ab_gpu_build_ephem_host
        f.argtypes = [c_void_p, c_void_p, c_size_t]
        rc = f(ctypes.cast(ephem.ptr, c_void_p), self.E, szE)
        assert rc == 0, self.plib.assist_gpu_last_error()
        self.has_params = 0

    def _plan(self):
        f = self.plib.ab_gpu_build_force_opts_host
        f.argtypes = [c_void_p, c_int, c_void_p, c_size_t]
        assert f(byref(self.opt), self.has_params, self.F, len(self.F)) == 0
        g = self.plib.ab_gpu_build_coop_plan_host
        g.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_size_t]
        assert g(ctypes.cast(self.ephem.ptr, c_void_p), self.E, self.F, self.budget, self.P, len(self.P)) == 0

    def set_state(self, t0, state, params=None, dt0=0.001):
        state = np.ascontiguousarray(state, dtype=np.float64).reshape(self.n, 6)
        prm = None if params is None else np.ascontiguousarray(params, dtype=np.float64).reshape(self.n, 3)
        self.has_params = 0 if prm is None else 1
        self.e.abc_emul_set_state(self.h, float(t0), float(dt0), _dp(state), _dp(prm) if prm is not None else None,
                                  float(self.opt.epsilon), float(self.opt.min_dt))

    def integrate(self, t_end, exact_finish_time=1):
        self._plan()
        rc = self.e.abc_emul_run(self.h, self.E, self.F, self.P, float(t_end), int(exact_finish_time), None, 0, None, self.n_blocks)
        assert rc == 0

    def integrate_or_interpolate(self, times):
        self._plan()
        times = np.ascontiguousarray(times, dtype=np.float64)
        out = np.full((times.size, self.n, 1, 6), np.nan)
        rc = self.e.abc_emul_run(self.h, self.E, self.F, self.P, 0.0, 0, _dp(times), times.size, _dp(out), self.n_blocks)
        assert rc == 0
        return out

    def get_state(self):
        st = np.empty((self.n, 1, 6)); t = np.empty(self.n); dt = np.empty(self.n); dtl = np.empty(self.n)
        status = np.empty(self.n, dtype=np.int32)
        cnt = np.empty((self.n, 4), dtype=np.uint64)
        self.e.abc_emul_get_state(self.h, _dp(st), _dp(t), _dp(dt), _dp(dtl), status.ctypes.data_as(POINTER(c_int)),
                                  cnt.ctypes.data_as(POINTER(c_ulonglong)))
        return dict(state=st, t=t, dt=dt, dt_last_done=dtl, status=status, counters=cnt)

    def close(self):
        if self.h:
            self.e.abc_emul_free(self.h)
            self.h = None
