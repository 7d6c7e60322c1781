"""Python mirror of the reference's `assist` package on top of libassist (assist-b200).

    import assist_b200 as assist
    ephem  = assist.Ephem("planets.bsp", "asteroids.bsp")       # reference assist/ephem.py:39-57
    sim    = assist.Simulation()                                 # stand-in for rebound.Simulation
    extras = assist.Extras(sim, ephem)                           # reference assist/extras.py:31-38
    sim.t = 8416.5
    sim.add(x=..., y=..., z=..., vx=..., vy=..., vz=...)
    sim.integrate(8446.5)                                        # runs on the GPU
    extras.integrate_or_interpolate(8450.0)

Same names, arguments and error behaviour as the reference (`Ephem.get_particle`, `Ephem.time_bounds`,
`Extras.forces`, `Extras.particle_params`, `Extras.detach`, `Extras.integrate_or_interpolate`,
`ASSIST_BODY_IDS`, `ASSIST_FORCES`, and assist/tools.py's `simulation_convert_to_rebound` and
`assist_create_interpolated_simulation` with `Simulation.save_to_file` / `SimulationArchive` standing in for rebound's).  The `rebound` Python package is not needed: `Simulation` and
`Particle` are thin ctypes views of the REBOUND surface in include/rebound.h -- the members ASSIST users
touch (t, dt, N, particles, add, add_variation, integrate, step, copy, ri_ias15.{epsilon,min_dt},
exact_finish_time, steps_done, status).  Many-particle work goes through `assist_b200.Batch`.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_char_p, c_double, c_int

import numpy as np

from . import _lib
from .cstructs import ASSIST_FORCES
from .cstructs import Particle as _CParticle

ASSIST_BODY_IDS = {
    0: "Sun", 1: "Mercury", 2: "Venus", 3: "Earth", 4: "Moon", 5: "Mars", 6: "Jupiter", 7: "Saturn",
    8: "Uranus", 9: "Neptune", 10: "Pluto", 11: "Camilla", 12: "Ceres", 13: "Cybele", 14: "Davida",
    15: "Eunomia", 16: "Euphrosyne", 17: "Europa", 18: "Hygiea", 19: "Interamnia", 20: "Iris",
    21: "Juno", 22: "Pallas", 23: "Psyche", 24: "Sylvia", 25: "Thisbe", 26: "Vesta",
}


def assist_error_messages(code: int) -> str:
    """Text of an ASSIST_STATUS code (reference assist/_libassist.py:46-67)."""
    lib = _lib.load()
    n = c_int.in_dll(lib, "assist_error_messages_N").value
    msgs = (c_char_p * n).in_dll(lib, "assist_error_messages")
    return msgs[code].decode("ascii") if 0 <= code < n else "Unknown ASSIST error %d" % code


class Particle(_CParticle):
    """struct reb_particle (128 bytes).  Cartesian initialisation only: ASSIST works in barycentric
    AU, AU/day (reference assist/test/test_basic.py:50-55)."""

    def __init__(self, x=0.0, y=0.0, z=0.0, vx=0.0, vy=0.0, vz=0.0, m=0.0, r=0.0):
        super().__init__()
        self.x, self.y, self.z, self.vx, self.vy, self.vz, self.m, self.r = x, y, z, vx, vy, vz, m, r

    def __sub__(self, other):
        return Particle(self.x - other.x, self.y - other.y, self.z - other.z,
                        self.vx - other.vx, self.vy - other.vy, self.vz - other.vz, self.m - other.m)

    @property
    def xyz(self):
        return [self.x, self.y, self.z]

    @property
    def vxyz(self):
        return [self.vx, self.vy, self.vz]

    def __repr__(self):
        return "<Particle x=%r y=%r z=%r vx=%r vy=%r vz=%r>" % (self.x, self.y, self.z, self.vx, self.vy, self.vz)


class _Particles:
    """sim.particles: indexable view of the simulation's particle array (live C memory)."""

    def __init__(self, sim):
        self._sim = sim

    def __len__(self):
        return int(self._sim._r.contents.N)

    def __getitem__(self, i):
        n = len(self)
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(n))]
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("Particle index out of range")
        p = Particle.from_address(ctypes.addressof(self._sim._r.contents.particles[i]))   # live view, same layout
        p._owner = self._sim            # the C array must outlive the view
        return p


class Simulation:
    """Stand-in for rebound.Simulation: IAS15, no self-gravity, forces from ASSIST.  An `Extras` must be
    attached before integrate()/step() (the library has no N-body gravity of its own)."""

    def __init__(self, _ptr=None):
        self._lib = _lib.load()
        self._r = _ptr if _ptr is not None else self._lib.reb_simulation_create()
        if not self._r:
            raise MemoryError("reb_simulation_create failed")
        self._extras_ref = None
        self.particles = _Particles(self)

    # scalar members a user of the reference touches
    def _get(name):
        return property(lambda self: getattr(self._r.contents, name),
                        lambda self, v: setattr(self._r.contents, name, v))

    t = _get("t")
    dt = _get("dt")
    dt_last_done = _get("dt_last_done")
    exact_finish_time = _get("exact_finish_time")
    status = _get("status")
    del _get

    @property
    def N(self):
        return int(self._r.contents.N)

    @property
    def N_var(self):
        return int(self._r.contents.N_var)

    @property
    def steps_done(self):
        return int(self._r.contents.steps_done)

    @property
    def ri_ias15(self):
        return self._r.contents.ri_ias15

    def add(self, particle=None, **kwargs):
        """sim.add(x=.., y=.., z=.., vx=.., vy=.., vz=..) or sim.add(Particle)."""
        if particle is None:
            bad = set(kwargs) - {"x", "y", "z", "vx", "vy", "vz", "m", "r"}
            if bad:
                raise ValueError("assist_b200.Simulation.add takes Cartesian coordinates only (got %s)" % ", ".join(sorted(bad)))
            particle = Particle(**kwargs)
        self._lib.reb_simulation_add(self._r, particle)

    def add_variation(self, testparticle=0):
        """First-order variational particle of real particle `testparticle`; returns its index in sim.particles
        (rebound's add_variation(testparticle=i) + .index, reference examples/variational)."""
        idx = self._lib.reb_simulation_add_variation_1st_order(self._r, int(testparticle))
        if idx < 0:
            raise RuntimeError("reb_simulation_add_variation_1st_order failed")
        return idx

    def integrate(self, tmax, exact_finish_time=1):
        self._r.contents.exact_finish_time = int(exact_finish_time)
        status = self._lib.reb_simulation_integrate(self._r, float(tmax))
        self._raise_messages()
        if status > 0:
            raise RuntimeError("reb_simulation_integrate stopped with status %d" % status)
        return status

    def step(self):
        self._lib.reb_simulation_step(self._r)
        self._raise_messages()

    def save_to_file(self, filename, step=None):
        """rebound's sim.save_to_file(filename, step=n): a snapshot before the first step of integrate() and after every
        n completed steps (what assist_create_interpolated_simulation reads back); step=None appends one snapshot now."""
        fn = str(filename).encode()
        if step is None:
            self._lib.reb_simulation_save_to_file(self._r, fn)
        else:
            self._lib.reb_simulation_save_to_file_step(self._r, fn, int(step))
        self._raise_messages()

    def copy(self):
        """Particles, time and step settings; ASSIST must be attached to the copy again (as in the reference)."""
        c = self._lib.reb_simulation_copy(self._r)
        if not c:
            raise MemoryError("reb_simulation_copy failed")
        return Simulation(_ptr=c)

    def _raise_messages(self):
        r = self._r.contents
        if r.messages_waiting and r.messages:
            msg = r.messages.decode("ascii", "replace")
            r.messages_waiting = 0
            if r.status > 0:
                raise RuntimeError(msg)

    def __del__(self):
        try:
            if self._extras_ref is not None:
                self._extras_ref._release()
            if self._r:
                self._lib.reb_simulation_free(self._r)
                self._r = None
        except Exception:
            pass


class SimulationArchive:
    """Stand-in for rebound.Simulationarchive: the snapshot file written by Simulation.save_to_file."""

    def __init__(self, filename):
        self._lib = _lib.load()
        self._sa = self._lib.reb_simulationarchive_create_from_file(str(filename).encode())
        if not self._sa:
            raise RuntimeError("cannot read the snapshot file %s" % filename)

    @property
    def nblobs(self):
        return int(self._sa.contents.nblobs)

    def __len__(self):
        return self.nblobs

    @property
    def t(self):
        return [self._sa.contents.t[i] for i in range(self.nblobs)]

    def __getitem__(self, i):
        """Snapshot i as a Simulation (time, particles, step data; ASSIST is not attached to it)."""
        r = self._lib.reb_simulation_create()
        self._lib.reb_simulation_create_from_simulationarchive_with_messages(r, self._sa, int(i), None)
        sim = Simulation(_ptr=r)
        if r.contents.messages_waiting:
            raise IndexError(r.contents.messages.decode("ascii", "replace"))
        return sim

    def __del__(self):
        try:
            if self._sa:
                self._lib.reb_simulationarchive_free(self._sa)
                self._sa = None
        except Exception:
            pass


def assist_create_interpolated_simulation(sa, t):
    """reference assist/tools.py:11-15: the simulation at time t, interpolated inside the step between two snapshots."""
    r = sa._lib.assist_create_interpolated_simulation(sa._sa, float(t))
    if not r:
        raise RuntimeError("requested time outside the range of the snapshot file")
    return Simulation(_ptr=r)


def simulation_convert_to_rebound(sim, ephem, merge_moon=1):
    """reference assist/tools.py:6-9: a plain simulation holding the ephemeris bodies at sim.t followed by sim's
    particles."""
    from .cstructs import Simulation as _CSimulation
    lib = sim._lib
    lib.assist_simulation_convert_to_rebound.restype = POINTER(_CSimulation)
    lib.assist_simulation_convert_to_rebound.argtypes = [POINTER(_CSimulation), ctypes.c_void_p, c_int]
    r = lib.assist_simulation_convert_to_rebound(sim._r, ctypes.cast(byref(ephem._c), ctypes.c_void_p), int(merge_moon))
    if not r:
        raise RuntimeError("assist_simulation_convert_to_rebound failed")
    return Simulation(_ptr=r)


class Ephem:
    """Main object for all ephemeris operations; not tied to a simulation (reference assist/ephem.py:33-120)."""

    def __init__(self, planets_path=None, asteroids_path=None):
        self._lib = _lib.load()
        from .cstructs import Ephem as _CEphem
        self._c = _CEphem()
        pp = None if planets_path is None else str(planets_path).encode("ascii")
        ap = None if asteroids_path is None else str(asteroids_path).encode("ascii")
        self._ok = False
        ret = self._lib.assist_ephem_init(byref(self._c), pp, ap)
        if ret != 0:
            raise RuntimeError(assist_error_messages(ret))
        self._ok = True

    def __getattr__(self, name):
        # jd_ref, AU, EMRAT, J2E, ..., over_c_squared: the fields of struct assist_ephem
        c = self.__dict__.get("_c")
        if c is not None and name in dict(c._fields_):
            return getattr(c, name)
        raise AttributeError(name)

    def get_particle(self, body, t):
        if isinstance(body, str):
            body_str = body.lower()
            body = -1
            for k, name in ASSIST_BODY_IDS.items():
                if body_str == name.lower():
                    body = k
            if body < 0:
                raise ValueError("Cannot find body '" + body_str + "'. Needs to be one of: " + ", ".join(ASSIST_BODY_IDS.values()) + ".")
        if not isinstance(body, (int, np.integer)) or isinstance(body, bool):
            raise ValueError("Expecting integer for body id.")
        e = c_int(0)
        p = self._lib.assist_get_particle_with_error(byref(self._c), int(body), float(t), byref(e))
        if e.value:
            raise RuntimeError(assist_error_messages(e.value))
        return Particle(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.m)

    def time_bounds(self):
        """(t_beg, t_end) of the ephemeris coverage, relative to jd_ref."""
        t_beg, t_end = c_double(0.0), c_double(0.0)
        self._lib.assist_ephem_time_bounds(byref(self._c), byref(t_beg), byref(t_end))
        return (t_beg.value, t_end.value)

    def __del__(self):
        try:
            if self._ok:
                self._lib.assist_ephem_free_pointers(byref(self._c))
                self._ok = False
        except Exception:
            pass


class Extras:
    """ASSIST attached to one simulation (reference assist/extras.py:26-100)."""

    def __init__(self, sim: Simulation, ephem: Ephem):
        self._lib = _lib.load()
        self._sim, self._ephem = sim, ephem           # keep both alive
        self._ax = self._lib.assist_attach(sim._r, byref(ephem._c))
        if not self._ax:
            raise RuntimeError("assist_attach failed")
        sim._extras_ref = self
        self._particle_params_reference = None

    def _release(self):
        if self._ax:
            self._lib.assist_free(self._ax)
            self._ax = None

    def __del__(self):
        try:
            if self._sim is not None and self._sim._extras_ref is self:
                self._sim._extras_ref = None
            self._release()
        except Exception:
            pass

    def detach(self, sim: Simulation):
        sim._extras_ref = None
        if self._ax:
            self._lib.assist_detach(sim._r, self._ax)

    def integrate_or_interpolate(self, t):
        self._lib.assist_integrate_or_interpolate(self._ax, float(t))
        self._sim._raise_messages()

    # struct assist_extras members
    def _field(name):
        return property(lambda self: getattr(self._ax.contents, name),
                        lambda self, v: setattr(self._ax.contents, name, v))

    geocentric = _field("geocentric")
    gr_eih_sources = _field("gr_eih_sources")
    alpha = _field("alpha")
    nk = _field("nk")
    nm = _field("nm")
    nn = _field("nn")
    r0 = _field("r0")
    del _field

    @property
    def forces(self):
        return [k for k, bit in ASSIST_FORCES.items() if self._ax.contents.forces & bit]

    @forces.setter
    def forces(self, value):
        if not isinstance(value, list):
            raise AttributeError("Forces need to be a list.")
        for elem in value:
            if not isinstance(elem, str):
                raise AttributeError("Each force needs to be a string.")
            if elem.upper() not in ASSIST_FORCES:
                raise AttributeError("Force '" + elem + "' not recognized. Needs to be one of the following: " + ", ".join(ASSIST_FORCES))
        v = 0
        for k, bit in ASSIST_FORCES.items():
            if k in value:
                v |= bit
        self._ax.contents.forces = v

    @property
    def particle_params(self):
        raise AttributeError("Cannot get particle_params. Only setting is supported.")

    @particle_params.setter
    def particle_params(self, value):
        self._particle_params_reference = np.array(value, dtype=np.float64).copy()     # keep the buffer alive
        self._ax.contents.particle_params = self._particle_params_reference.ctypes.data_as(POINTER(c_double))
