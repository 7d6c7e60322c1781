"""Loader for the in-tree libassist.so (same discovery rule as the reference's
assist/_libassist.py:34-41: $ASSIST_LIBASSIST_PATH overrides the default location).

No PyTorch, no numpy requirement here: plain ctypes.  Importing this module never
touches the GPU; the first compute call does, and fails loudly without a device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_long, c_ulonglong, c_void_p

from .cstructs import Ephem, bind

_here = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    override = os.environ.get("ASSIST_LIBASSIST_PATH")
    if override:
        if not os.path.isabs(override):
            override = os.path.join(os.path.dirname(_here), override)
        return override
    return os.path.join(_here, "libassist.so")


class GpuOptions(Structure):
    _fields_ = [("forces", c_int), ("gr_eih_sources", c_int), ("geocentric", c_int), ("math", c_int),
                ("alpha", c_double), ("nk", c_double), ("nm", c_double), ("nn", c_double), ("r0", c_double),
                ("epsilon", c_double), ("min_dt", c_double)]


class GpuStats(Structure):
    _fields_ = [("steps", c_ulonglong), ("steps_rejected", c_ulonglong), ("pc_iterations", c_ulonglong),
                ("force_evals", c_ulonglong), ("kernel_launches", c_ulonglong), ("last_kernel_ms", c_double)]


_lib = None


def load():
    """Load (once) and return the bound library.  Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "assist-b200: %s not found. Build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C assist_b200/csrc). "
            "There is no pure-Python / CPU fallback." % path)
    lib = bind(ctypes.CDLL(path))
    P = POINTER
    lib.assist_gpu_device_count.restype = c_int
    lib.assist_gpu_set_device.argtypes = [c_int]
    lib.assist_gpu_last_error.restype = c_char_p
    lib.assist_gpu_default_options.argtypes = [P(GpuOptions)]
    lib.assist_gpu_default_options.restype = None
    lib.assist_gpu_ephem_upload.argtypes = [P(Ephem)]
    lib.assist_gpu_ephem_nbodies.argtypes = [P(Ephem)]
    lib.assist_gpu_ephem_eval.argtypes = [P(Ephem), c_int, P(c_double), c_int, P(c_double), P(c_int)]
    lib.assist_gpu_eval_forces.argtypes = [P(Ephem), P(GpuOptions), c_int, c_int, P(c_double), c_int,
                                           P(c_double), P(c_double), P(c_double), P(c_int)]
    lib.assist_gpu_batch_create.restype = c_void_p
    lib.assist_gpu_batch_create.argtypes = [P(Ephem), c_int, c_int, c_int]
    lib.assist_gpu_batch_free.restype = None
    lib.assist_gpu_batch_free.argtypes = [c_void_p]
    lib.assist_gpu_batch_set_options.argtypes = [c_void_p, P(GpuOptions)]
    lib.assist_gpu_batch_set_state.argtypes = [c_void_p, c_double, c_double, P(c_double), P(c_double), P(c_int)]
    lib.assist_gpu_batch_update_particles.argtypes = [c_void_p, P(c_double)]
    lib.assist_gpu_batch_snapshot.argtypes = [c_void_p]
    lib.assist_gpu_batch_restore.argtypes = [c_void_p]
    lib.assist_gpu_batch_integrate.argtypes = [c_void_p, c_double, c_int, c_long]
    lib.assist_gpu_batch_integrate_or_interpolate.argtypes = [c_void_p, P(c_double), c_int, P(c_double)]
    lib.assist_gpu_batch_get_state.argtypes = [c_void_p, P(c_double), P(c_double), P(c_double), P(c_double),
                                               P(c_double), P(c_int)]
    lib.assist_gpu_batch_set_time.argtypes = [c_void_p, c_double, c_double]
    lib.assist_gpu_batch_interpolate.argtypes = [c_void_p, c_double, P(c_double)]
    lib.assist_gpu_batch_get_stats.argtypes = [c_void_p, P(GpuStats)]
    lib.assist_gpu_measure_fp64_peak.restype = c_double
    lib.assist_gpu_measure_fp64_peak.argtypes = [c_int]
    lib.assist_gpu_kernel_launches.restype = c_ulonglong
    lib.assist_gpu_kernel_launches.argtypes = []
    lib.assist_gpu_batch_get_counters.argtypes = [c_void_p, P(c_ulonglong), P(c_ulonglong), P(c_ulonglong), P(c_ulonglong)]
    lib.assist_gpu_host_alloc.restype = c_void_p
    lib.assist_gpu_host_alloc.argtypes = [ctypes.c_size_t]
    lib.assist_gpu_host_free.restype = None
    lib.assist_gpu_host_free.argtypes = [c_void_p]
    lib.assist_gpu_selftest_fp.argtypes = [c_ulonglong, ctypes.c_longlong, P(c_ulonglong)]
    # one population over several GPUs
    lib.assist_gpu_multi_create.restype = c_void_p
    lib.assist_gpu_multi_create.argtypes = [P(Ephem), c_int, c_int, P(c_int), c_int]
    lib.assist_gpu_multi_free.restype = None
    lib.assist_gpu_multi_free.argtypes = [c_void_p]
    lib.assist_gpu_multi_device_count.argtypes = [c_void_p]
    lib.assist_gpu_multi_last_error.restype = c_char_p
    lib.assist_gpu_multi_set_options.argtypes = [c_void_p, P(GpuOptions)]
    lib.assist_gpu_multi_set_state.argtypes = [c_void_p, c_double, c_double, P(c_double), P(c_double)]
    lib.assist_gpu_multi_integrate.argtypes = [c_void_p, c_double, c_int]
    lib.assist_gpu_multi_integrate_or_interpolate.argtypes = [c_void_p, P(c_double), c_int, P(c_double)]
    lib.assist_gpu_multi_get_state.argtypes = [c_void_p, P(c_double), P(c_double), P(c_double), P(c_double), P(c_int)]
    lib.assist_gpu_multi_get_counters.argtypes = [c_void_p, P(c_ulonglong), P(c_ulonglong), P(c_ulonglong), P(c_ulonglong)]
    lib.assist_gpu_multi_get_stats.argtypes = [c_void_p, P(GpuStats), P(c_double)]
    _lib = lib
    return lib
