"""ctypes mirrors of the C structs in include/rebound.h and include/assist.h.

`Ephem` and `Extras` follow, field by field, the reference's own Python mirrors
(reference assist/ephem.py:95-120 and assist/extras.py:84-100), which in turn
mirror `struct assist_ephem` / `struct assist_extras` (reference src/assist.h:125-195).
The REBOUND structs mirror include/rebound.h (our minimal REBOUND surface).

`bind(lib)` declares argument / return types on a loaded library.  The same
declarations serve the product library (libassist) and the test oracle
(oracle/_ref/libassist_ref.so), because both are compiled against the same headers.
"""
from __future__ import annotations

import ctypes
from ctypes import (CFUNCTYPE, POINTER, Structure, c_char_p, c_double, c_int, c_int64,
                    c_long, c_size_t, c_uint, c_uint32, c_uint64, c_ulonglong, c_void_p)


class Particle(Structure):
    _fields_ = [("x", c_double), ("y", c_double), ("z", c_double),
                ("vx", c_double), ("vy", c_double), ("vz", c_double),
                ("ax", c_double), ("ay", c_double), ("az", c_double),
                ("m", c_double), ("r", c_double), ("last_collision", c_double),
                ("c", c_void_p), ("hash", c_uint32), ("ap", c_void_p), ("sim", c_void_p)]


assert ctypes.sizeof(Particle) == 128


class DP7(Structure):
    _fields_ = [("p%d" % i, POINTER(c_double)) for i in range(7)]


class VarConfig(Structure):
    _fields_ = [("sim", c_void_p), ("order", c_int), ("index", c_int), ("testparticle", c_int),
                ("index_1st_order_a", c_int), ("index_1st_order_b", c_int), ("lrescale", c_double)]


class IAS15(Structure):
    _fields_ = [("epsilon", c_double), ("min_dt", c_double), ("adaptive_mode", c_uint),
                ("iterations_max_exceeded", c_uint64), ("N_allocated", c_uint),
                ("at", POINTER(c_double)), ("x0", POINTER(c_double)), ("v0", POINTER(c_double)),
                ("a0", POINTER(c_double)), ("csx", POINTER(c_double)), ("csv", POINTER(c_double)),
                ("csa0", POINTER(c_double)),
                ("g", DP7), ("b", DP7), ("csb", DP7), ("e", DP7), ("br", DP7), ("er", DP7),
                ("map", POINTER(c_int)), ("N_allocated_map", c_uint),
                ("b200_pc_iterations", c_uint64), ("b200_force_evals", c_uint64),
                ("b200_steps_rejected", c_uint64)]


class Simulation(Structure):
    pass


FORCE_FN = CFUNCTYPE(None, POINTER(Simulation))

Simulation._fields_ = [
    ("t", c_double), ("G", c_double), ("softening", c_double), ("dt", c_double),
    ("dt_last_done", c_double), ("steps_done", c_uint64), ("N", c_uint), ("N_var", c_int),
    ("N_var_config", c_uint), ("var_config", POINTER(VarConfig)), ("N_active", c_int),
    ("N_allocated", c_uint), ("particles", POINTER(Particle)), ("status", c_int),
    ("exact_finish_time", c_int), ("force_is_velocity_dependent", c_uint),
    ("integrator", c_int), ("gravity", c_int), ("ri_ias15", IAS15),
    ("additional_forces", FORCE_FN), ("pre_timestep_modifications", FORCE_FN),
    ("post_timestep_modifications", FORCE_FN), ("heartbeat", FORCE_FN),
    ("extras", c_void_p), ("extras_cleanup", FORCE_FN),
    ("messages", c_char_p), ("messages_waiting", c_int), ("b200_batch", c_void_p),
    ("simulationarchive_filename", c_char_p), ("simulationarchive_auto_step", c_uint64),
    ("simulationarchive_next_step", c_uint64)]


class SimulationArchive(Structure):
    """include/rebound.h: struct reb_simulationarchive (snapshot file of an attached simulation)."""
    _fields_ = [("r", POINTER(Simulation)), ("filename", c_char_p), ("nblobs", c_long), ("t", POINTER(c_double)),
                ("b200_offset", POINTER(c_long))]


class Ephem(Structure):
    _fields_ = [("jd_ref", c_double), ("spk_planets", c_void_p), ("spk_asteroids", c_void_p),
                ("ascii_planets", c_void_p), ("planets_source", c_int), ("planets_calc", c_void_p),
                ("spk_target_index", c_int * 11), ("spk_emb_index", c_int),
                ("AU", c_double), ("EMRAT", c_double), ("J2E", c_double), ("J3E", c_double),
                ("J4E", c_double), ("J2SUN", c_double), ("RE", c_double), ("CLIGHT", c_double),
                ("ASUN", c_double), ("Re_eq", c_double), ("Rs_eq", c_double),
                ("c_AU_per_day", c_double), ("c_squared", c_double), ("over_c_squared", c_double)]


assert ctypes.sizeof(Ephem) == 208


class Extras(Structure):
    _fields_ = [("sim", POINTER(Simulation)), ("ephem", POINTER(Ephem)), ("ephem_cache", c_void_p),
                ("extras_should_free_ephem", c_int), ("geocentric", c_int),
                ("last_state", POINTER(Particle)), ("current_state", POINTER(Particle)),
                ("particle_params", POINTER(c_double)), ("steps_done", c_int), ("forces", c_int),
                ("gr_eih_sources", c_int), ("alpha", c_double), ("nk", c_double), ("nm", c_double),
                ("nn", c_double), ("r0", c_double)]


assert ctypes.sizeof(Extras) == 112

ASSIST_FORCES = {
    "SUN": 0x01, "PLANETS": 0x02, "ASTEROIDS": 0x04, "NON_GRAVITATIONAL": 0x08,
    "EARTH_HARMONICS": 0x10, "SUN_HARMONICS": 0x20, "GR_EIH": 0x40, "GR_SIMPLE": 0x80,
    "GR_POTENTIAL": 0x100,
}


def bind(lib):
    """Declare the ASSIST + REBOUND-surface prototypes on a loaded CDLL."""
    P = POINTER
    lib.reb_simulation_create.restype = P(Simulation)
    lib.reb_simulation_create.argtypes = []
    lib.reb_simulation_free.restype = None
    lib.reb_simulation_free.argtypes = [P(Simulation)]
    lib.reb_simulation_copy.restype = P(Simulation)
    lib.reb_simulation_copy.argtypes = [P(Simulation)]
    lib.reb_simulation_add.restype = None
    lib.reb_simulation_add.argtypes = [P(Simulation), Particle]
    lib.reb_simulation_add_variation_1st_order.restype = c_int
    lib.reb_simulation_add_variation_1st_order.argtypes = [P(Simulation), c_int]
    lib.reb_simulation_integrate.restype = c_int
    lib.reb_simulation_integrate.argtypes = [P(Simulation), c_double]
    lib.reb_simulation_step.restype = None
    lib.reb_simulation_step.argtypes = [P(Simulation)]
    lib.reb_simulation_update_acceleration.restype = None
    lib.reb_simulation_update_acceleration.argtypes = [P(Simulation)]

    lib.assist_ephem_create.restype = P(Ephem)
    lib.assist_ephem_create.argtypes = [c_char_p, c_char_p]
    lib.assist_ephem_init.restype = c_int
    lib.assist_ephem_init.argtypes = [P(Ephem), c_char_p, c_char_p]
    lib.assist_ephem_free.restype = None
    lib.assist_ephem_free.argtypes = [P(Ephem)]
    lib.assist_ephem_free_pointers.restype = None
    lib.assist_ephem_free_pointers.argtypes = [P(Ephem)]
    lib.assist_attach.restype = P(Extras)
    lib.assist_attach.argtypes = [P(Simulation), P(Ephem)]
    lib.assist_init.restype = None
    lib.assist_init.argtypes = [P(Extras), P(Simulation), P(Ephem)]
    lib.assist_free.restype = None
    lib.assist_free.argtypes = [P(Extras)]
    lib.assist_free_pointers.restype = None
    lib.assist_free_pointers.argtypes = [P(Extras)]
    lib.assist_detach.restype = None
    lib.assist_detach.argtypes = [P(Simulation), P(Extras)]
    lib.assist_integrate_or_interpolate.restype = None
    lib.assist_integrate_or_interpolate.argtypes = [P(Extras), c_double]
    lib.assist_get_particle.restype = Particle
    lib.assist_get_particle.argtypes = [P(Ephem), c_int, c_double]
    lib.assist_get_particle_with_error.restype = Particle
    lib.assist_get_particle_with_error.argtypes = [P(Ephem), c_int, c_double, P(c_int)]
    lib.assist_ephem_time_bounds.restype = None
    lib.assist_ephem_time_bounds.argtypes = [P(Ephem), P(c_double), P(c_double)]
    lib.assist_additional_forces.restype = None
    lib.assist_additional_forces.argtypes = [P(Simulation)]
    lib.assist_all_ephem.restype = c_int
    lib.assist_all_ephem.argtypes = [P(Ephem), c_void_p, c_int, c_double] + [P(c_double)] * 10
    lib.assist_detect_ephemeris_file_format.restype = c_int
    lib.assist_detect_ephemeris_file_format.argtypes = [c_int]
    lib.reb_simulation_create_from_simulationarchive_with_messages.restype = None
    lib.reb_simulation_create_from_simulationarchive_with_messages.argtypes = [P(Simulation), P(SimulationArchive), c_int64, c_void_p]
    lib.reb_simulation_steps.restype = None
    lib.reb_simulation_steps.argtypes = [P(Simulation), c_uint]
    lib.assist_interpolate_simulation.restype = c_int
    lib.assist_interpolate_simulation.argtypes = [P(Simulation), P(Simulation), c_double]
    lib.assist_create_interpolated_simulation.restype = P(Simulation)
    lib.assist_create_interpolated_simulation.argtypes = [P(SimulationArchive), c_double]
    if hasattr(lib, "reb_simulation_save_to_file_step"):      # the snapshot file: product library only
        lib.reb_simulation_save_to_file_step.restype = None
        lib.reb_simulation_save_to_file_step.argtypes = [P(Simulation), c_char_p, c_ulonglong]
        lib.reb_simulation_save_to_file.restype = None
        lib.reb_simulation_save_to_file.argtypes = [P(Simulation), c_char_p]
        lib.reb_simulationarchive_create_from_file.restype = P(SimulationArchive)
        lib.reb_simulationarchive_create_from_file.argtypes = [c_char_p]
        lib.reb_simulationarchive_free.restype = None
        lib.reb_simulationarchive_free.argtypes = [P(SimulationArchive)]
    return lib
