#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: how far can a REBOUND 4.x binary be from this repo's restatement of IAS15?

REBOUND is absent offline, so the restatement oracle/reb_shim.c cannot be compared with it.
This script bounds the two places where the restatement is not certain to follow REBOUND's own text
(reb_shim_set_variant): the FORM of the predictor polynomial (nested Horner vs the s[0..8] coefficients of the
reference's own copy, src/assist.c:562-596 -- same value, different rounding) and whether a REJECTED step restores
the accelerations of the step start.  It integrates BASELINE config 1 (Apophis-like, 11 EIH sources, Marsden
terms) and a 100-particle sample of the C3 population over 10 yr with the reference's src/*.c under each variant and
prints the spread against the default variant.

    python oracle/ias15_variants.py          (build container; needs oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refharness as rh
from assist_b200.synth import ephem_writer, populations

VARIANTS = [("Horner, restore a (default)", 0, 1), ("s[] form, restore a", 1, 1), ("Horner, keep a", 0, 0), ("s[] form, keep a", 1, 0)]


def run(lib, eph, state, t_end, **kw):
    out = {}
    for name, form, restore in VARIANTS:
        lib.reb_shim_set_variant(form, restore)
        fin, ts, dts, cnt = rh.integrate_each(lib, eph, populations.T0, state, t_end, **kw)
        out[name] = (fin[:, 0, :3], cnt)
    lib.reb_shim_set_variant(0, 1)
    return out


def spread(out):
    base = out[VARIANTS[0][0]][0]
    rows = []
    for name, _, _ in VARIANTS[1:]:
        d = np.linalg.norm(out[name][0] - base, axis=-1)
        rows.append((name, float(d.max()), float(np.median(d)), out[name][1]))
    return rows


def main():
    paths = ephem_writer.write_all(os.path.join(ROOT, "data"))
    lib = rh.ref_lib()
    eph = rh.open_ephem(lib, paths["planets_bsp"], paths["asteroids_bsp"])
    T0 = populations.T0
    c1, c1p = populations.apophis_like()
    pop = populations.neo_mba_mix(1000000, seed=20261703)[::10000]
    cases = [
        ("C1 Apophis-like, 11 EIH sources, Marsden, min_dt 1e-3, 3652.5 d", c1, dict(params=c1p[:, None, :], forces=0x7F, gr_eih_sources=11, min_dt=1e-3)),
        ("C3 100-particle sample, forces 0x7F, min_dt 1e-3, 3652.5 d", pop, dict(forces=0x7F, min_dt=1e-3)),
        ("C3 100-particle sample, first step 400 d (every system starts with rejected attempts)", pop, dict(forces=0x7F, min_dt=1e-3, dt0=400.0)),
    ]
    for title, st, kw in cases:
        out = run(lib, eph, st, T0 + 3652.5, **kw)
        base_cnt = out[VARIANTS[0][0]][1]
        print(title)
        print("    default variant: %d steps, %d rejected" % (base_cnt["steps"], base_cnt["rejected"]))
        for name, dmax, dmed, cnt in spread(out):
            print("    %-22s max |dx| %.3e AU   median %.3e AU   (%d steps, %d rejected)" % (name, dmax, dmed, cnt["steps"], cnt["rejected"]))


if __name__ == "__main__":
    main()
