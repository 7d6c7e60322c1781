#!/usr/bin/env python3
"""Golden vectors for small-body kernels with more targets than the 16 of sb441-n16 (sb441-n373),
from the reference's own C code (oracle/_ref), which takes any number of targets.

    python tests/golden/make_golden_n373.py            (build container: needs /root/reference -> oracle/_ref)

Files: assist_b200.synth.ephem_writer.write_extended -- a small-body kernel with N targets (the 16 of sb441-n16
first, then seeded main-belt ellipses) and a planets kernel whose comment area carries all N masses.  Two sizes:
N = 40 (every case) and N = 373 (ephemeris and one force evaluation).  Writes tests/golden/golden_n373.npz:

  eph{N}      assist_all_ephem for all 11 + N bodies at 5 times (GM, x, y, z of the asteroids; everything of the planets)
  acc{N}      one force evaluation (mask 0x7F) of 6 systems with 2 variational particles each
  pp_*        N = 40: 8 NEO+MBA particles, per-particle dt, 400 d: final states, t, dt, step / sweep / evaluation counts
  sh_*        N = 40: ONE shared-step simulation of 6 particles + 1 variational particle each, 300 d
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refharness as rh
from assist_b200.synth import ephem_writer
import cases

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    lib = rh.ref_lib()
    out = {}
    for N in cases.N373_SIZES:
        paths = ephem_writer.write_extended(os.path.join(ROOT, "data"), N)
        eph = rh.open_ephem(lib, paths["planets_bsp"], paths["asteroids_bsp"])
        b, st = rh.all_bodies(lib, eph, cases.n373_times(), nbodies=11 + N)
        assert (st == 0).all()
        out["eph%d" % N] = b
        state = cases.n373_force_systems()
        out["acc%d" % N] = rh.forces(lib, eph, cases.T0 + 17.25, state, forces=0x7F)
        if N == cases.N373_SIZES[0]:
            pp = cases.n373_pp_particles()
            out["pp_final"], out["pp_t"], out["pp_dt"], tot = rh.integrate_each(lib, eph, cases.T0, pp, cases.T0 + 400.0, forces=0x7F)
            out["pp_counts"] = np.array([tot["steps"], tot["pc_iterations"], tot["force_evals"], tot["rejected"]], dtype=np.int64)
            sh = cases.n373_shared_systems()
            s = rh.Sim(lib, eph, cases.T0, sh, forces=0x7F)
            s.integrate(cases.T0 + 300.0)
            out["sh_final"] = s.state()
            out["sh_t_dt"] = np.array([s.t, s.dt, s.dt_last_done])
            c = s.counters()
            out["sh_counts"] = np.array([c["steps"], c["pc_iterations"], c["force_evals"], c["rejected"]], dtype=np.int64)
            s.close()
        print("N = %d done" % N, flush=True)
    np.savez_compressed(os.path.join(OUT, "golden_n373.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
