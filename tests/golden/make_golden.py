#!/usr/bin/env python3
"""Generate the committed golden vectors from the reference's own C code (oracle/_ref).

Run in the build container (needs oracle/_ref/libassist_ref.so, i.e. /root/reference):
    python tests/golden/make_golden.py
Everything is computed on the synthetic ephemeris files written by
assist_b200.synth.ephem_writer (deterministic), with seeded populations, through the
reference's public API only (tests/refharness.py).  The GPU tests compare the CUDA path
with these files when oracle/_ref is not present on the test box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refharness as rh
from assist_b200.synth import ephem_writer, populations
import cases

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    paths = ephem_writer.write_all(os.path.join(ROOT, "data"))
    lib = rh.ref_lib()
    for key in ("planets_bsp", "de440"):
        eph = rh.open_ephem(lib, paths[key], paths["asteroids_bsp"])
        tag = "bsp" if key == "planets_bsp" else "440"
        out = {}
        # ephemeris states
        out["ephem_times"] = cases.EPHEM_TIMES
        out["ephem"], out["ephem_status"] = rh.all_bodies(lib, eph, cases.EPHEM_TIMES)
        # force terms
        state, params = cases.force_case()
        for name, mask, src, geo in cases.FORCE_TERMS:
            out["force_" + name] = rh.forces(lib, eph, cases.FORCE_T, state, params, forces=mask, gr_eih_sources=src, geocentric=geo)
        # per-particle integration
        st = cases.pp_case()
        fin, ts, dts, cnt = rh.integrate_each(lib, eph, cases.T0, st, cases.T0 + cases.PP_DAYS, forces=0x7F, min_dt=0.0)
        out["pp_final"], out["pp_t"], out["pp_dt"] = fin, ts, dts
        out["pp_counts"] = np.array([cnt["steps"], cnt["pc_iterations"], cnt["force_evals"], cnt["rejected"]])
        # backward, exact finish
        fin, ts, dts, cnt = rh.integrate_each(lib, eph, cases.T0, st[:6], cases.T0 - 300.0, forces=0x7F)
        out["ppback_final"], out["ppback_t"] = fin, ts
        # variational
        stv = cases.var_case()
        fin, ts, dts, cnt = rh.integrate_each(lib, eph, cases.T0, stv, cases.T0 + cases.VAR_DAYS, forces=0x7F)
        out["var_final"], out["var_t"] = fin, ts
        # C1: Apophis-like, all forces, 11 EIH sources, non-grav, min_dt 1e-3, 10 yr
        c1, c1p = populations.apophis_like()
        fin, ts, dts, cnt = rh.integrate_each(lib, eph, cases.T0, c1, cases.T0 + 3652.5, params=c1p[:, None, :], forces=0x7F,
                                              gr_eih_sources=11, min_dt=1e-3)
        out["c1_final"], out["c1_t"] = fin, ts
        out["c1_counts"] = np.array([cnt["steps"], cnt["pc_iterations"], cnt["force_evals"], cnt["rejected"]])
        # shared step (C2-like, small)
        sh = cases.shared_case()
        s = rh.Sim(lib, eph, cases.T0, sh, forces=0x77)
        s.integrate(cases.T0 + cases.SH_DAYS)
        out["sh_final"] = s.state()
        out["sh_t_dt"] = np.array([s.t, s.dt, s.dt_last_done])
        c = s.counters()
        out["sh_counts"] = np.array([c["steps"], c["pc_iterations"], c["force_evals"], c["rejected"]])
        s.close()
        # shared step with variational particles and a second real particle (reference unit test layout)
        shv = cases.shared_var_case()
        s = rh.Sim(lib, eph, cases.T0, shv, forces=0x7F)
        s.integrate(cases.T0 + 101.0)
        out["shv_final"] = s.state()
        s.close()
        # dense output, comets with non-gravitational forces, backward
        stc, prm = cases.comet_case()
        out["dense"] = rh.dense_each(lib, eph, cases.T0, stc, cases.DENSE_TIMES, params=prm[:, None, :], forces=0x7F)
        np.savez_compressed(os.path.join(OUT, "golden_%s.npz" % tag), **out)
        print("wrote golden_%s.npz" % tag, {k: np.asarray(v).shape for k, v in out.items()})
        lib.assist_ephem_free(eph)

    # smoke() fixture
    eph = rh.open_ephem(lib, paths["planets_bsp"], paths["asteroids_bsp"])
    st = populations.neo_mba_mix(8, seed=42)
    stv = populations.with_variations(st, 6)
    acc = rh.forces(lib, eph, cases.T0 + 1.25, stv, None, forces=0x7F, gr_eih_sources=1)
    fin, ts, _, _ = rh.integrate_each(lib, eph, cases.T0, st, cases.T0 + 60.0, forces=0x7F, min_dt=1e-3)
    np.savez_compressed(os.path.join(OUT, "smoke.npz"), acc=acc, final=fin, t=ts)
    print("wrote smoke.npz")


if __name__ == "__main__":
    main()
