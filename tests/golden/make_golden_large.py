#!/usr/bin/env python3
"""Golden vectors at the sizes BASELINE.json's north_star states, from the reference's own C code (oracle/_ref).

    python tests/golden/make_golden_large.py            (build container: needs /root/reference -> oracle/_ref)

Writes tests/golden/golden_large.npz (SPK planets file; the DE-binary format is covered at small sizes by
golden_440.npz).  Cases (tests/cases.py holds the selections so that generator and tests cannot drift apart):

  c3   fixed 1000-particle subsample of the 10^6 NEO+MBA bench population, per-particle dt, forces 0x7F,
       min_dt 1e-3 d, 3652.5 d: final states, t, dt, per-particle step / sweep / evaluation / rejection counts
  c4   256 systems (real particle + 6 variational) of the C4 bench population, 1826.25 d
  c2   ONE shared-step simulation of the 10^4 main-belt bench population, forces 0x77, 3652.5 d: final states,
       t, dt, dt_last_done, step / sweep / evaluation / rejection counts
  c5   256 comets with Marsden A1..A3 of the C5 bench population, assist_integrate_or_interpolate at 1827 epochs
       (every 10 d, 50 yr BACKWARD): every 29th epoch of all comets + every epoch of the first 8
  geo  geocentric integration (extras->geocentric = 1) of 12 near-Earth particles, 400 d
  eih  64 NEO+MBA particles with gr_eih_sources = 11, 1000 d

One reference simulation per particle (per-particle semantics), spread over the host cores.
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refharness as rh
from assist_b200.synth import ephem_writer
import cases

OUT = os.path.dirname(os.path.abspath(__file__))
_P = {}


def _eph():
    if "eph" not in _P:
        paths = ephem_writer.write_all(os.path.join(ROOT, "data"))
        _P["lib"] = rh.ref_lib()
        _P["eph"] = rh.open_ephem(_P["lib"], paths["planets_bsp"], paths["asteroids_bsp"])
    return _P["lib"], _P["eph"]


def _one(job):
    kind, state, params, t_end, kw = job
    lib, eph = _eph()
    s = rh.Sim(lib, eph, cases.T0, state[None], params=None if params is None else params[None], **kw)
    if kind == "integrate":
        s.integrate(t_end)
        c = s.counters()
        res = (s.state()[0], s.t, s.dt, [c["steps"], c["pc_iterations"], c["force_evals"], c["rejected"]])
    else:
        res = np.stack([s.integrate_or_interpolate(t)[0] for t in t_end])
    s.close()
    return res


def _each(pool, state, t_end, params=None, **kw):
    if state.ndim == 2:
        state = state[:, None, :]
    n = state.shape[0]
    jobs = [("integrate", state[i], None if params is None else np.asarray(params).reshape(n, -1, 3)[i], t_end, kw) for i in range(n)]
    res = pool.map(_one, jobs, chunksize=max(1, n // 64))
    return (np.stack([r[0] for r in res]), np.array([r[1] for r in res]), np.array([r[2] for r in res]),
            np.array([r[3] for r in res], dtype=np.int64))


def main():
    t00 = time.time()
    out = {}
    with Pool(os.cpu_count()) as pool:
        st = cases.c3_subsample()
        out["c3_final"], out["c3_t"], out["c3_dt"], out["c3_counts"] = _each(pool, st, cases.T0 + 3652.5, forces=0x7F, min_dt=1e-3)
        print("c3 done %.0f s, steps %d" % (time.time() - t00, out["c3_counts"][:, 0].sum()), flush=True)

        st = cases.c4_subsample()
        out["c4_final"], out["c4_t"], out["c4_dt"], out["c4_counts"] = _each(pool, st, cases.T0 + 1826.25, forces=0x7F, min_dt=0.0)
        print("c4 done %.0f s, steps %d" % (time.time() - t00, out["c4_counts"][:, 0].sum()), flush=True)

        st, prm = cases.c5_subsample()
        n = st.shape[0]
        jobs = [("dense", st[i][None], prm[i][None], cases.C5_EPOCHS, dict(forces=0x7F, min_dt=1e-3)) for i in range(n)]
        res = pool.map(_one, jobs, chunksize=4)
        dense = np.stack(res, axis=1)[:, :, 0, :]                 # [epoch][comet][6]
        out["c5_sparse"] = dense[cases.C5_SPARSE]
        out["c5_first8"] = dense[:, :8]
        print("c5 done %.0f s" % (time.time() - t00), flush=True)

        st = cases.eih11_case()
        out["eih_final"], out["eih_t"], out["eih_dt"], out["eih_counts"] = _each(pool, st, cases.T0 + 1000.0, forces=0x7F, gr_eih_sources=11, min_dt=1e-3)
        print("eih11 done %.0f s" % (time.time() - t00), flush=True)

        lib, eph = _eph()
        st = cases.geocentric_case(lambda t: rh.all_bodies(lib, eph, [t])[0][0, 3, 1:7])
        out["geo_init"] = st
        out["geo_final"], out["geo_t"], out["geo_dt"], out["geo_counts"] = _each(pool, st, cases.T0 + 400.0, forces=0x7F, geocentric=1, min_dt=1e-3)
        print("geocentric done %.0f s" % (time.time() - t00), flush=True)

    # C2: one simulation, one core
    lib, eph = _eph()
    sh = cases.c2_population()
    s = rh.Sim(lib, eph, cases.T0, sh, forces=0x77)
    s.integrate(cases.T0 + 3652.5)
    out["c2_final"] = s.state()[:, 0, :]
    out["c2_t_dt"] = np.array([s.t, s.dt, s.dt_last_done])
    c = s.counters()
    out["c2_counts"] = np.array([c["steps"], c["pc_iterations"], c["force_evals"], c["rejected"]])
    s.close()
    print("c2 done %.0f s, steps %d" % (time.time() - t00, c["steps"]), flush=True)

    np.savez_compressed(os.path.join(OUT, "golden_large.npz"), **out)
    print("wrote golden_large.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
