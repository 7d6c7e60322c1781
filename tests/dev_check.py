"""Developer parity check: CUDA path vs the reference build, printed as numbers.
Run on a GPU box: python tests/dev_check.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from assist_b200 import batch as ab
from assist_b200.synth import ephem_writer, populations
import refharness as rh


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    num = np.linalg.norm(a - b, axis=-1)
    den = np.linalg.norm(b, axis=-1)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(num / den))


def main():
    paths = ephem_writer.write_all(os.path.join(ROOT, "data"))
    ref = rh.ref_lib()
    T0 = populations.T0
    for planets in (paths["planets_bsp"], paths["de440"]):
        print("=== planets file:", os.path.basename(planets))
        eph = ab.EphemHandle(planets, paths["asteroids_bsp"])
        reph = rh.open_ephem(ref, planets, paths["asteroids_bsp"])
        # ephemeris
        times = np.array([T0, T0 + 17.3, T0 - 1234.56789, T0 + 3000.25, -10000.0])
        out, st = eph.eval(times)
        rout, rst = rh.all_bodies(ref, reph, times)
        same = np.array_equal(np.nan_to_num(out, nan=-7.0), np.nan_to_num(rout, nan=-7.0))
        print("ephem bit-identical:", same, " max abs diff:", np.nanmax(np.abs(out - rout)), "status", st.max(), rst.max())
        # forces, term by term
        n = 64
        st6 = populations.neo_mba_mix(n, seed=5)
        state = populations.with_variations(st6, 6)
        rng = np.random.default_rng(3)
        state[:, 1:, :] += 0.1 * rng.standard_normal(state[:, 1:, :].shape)
        params = np.zeros((n, 7, 3))
        params[:, 0, :] = [1e-9, -2e-10, 3e-11]
        params[:, 1:, :] = rng.standard_normal((n, 6, 3))
        for name, mask, src in [("SUN", 0x01, 1), ("PLANETS", 0x02, 1), ("ASTEROIDS", 0x04, 1), ("NONGRAV", 0x08, 1),
                                ("EARTH_HARM", 0x10, 1), ("SUN_HARM", 0x20, 1), ("GR_EIH(1)", 0x40, 1),
                                ("GR_EIH(11)", 0x40, 11), ("GR_SIMPLE", 0x80, 1), ("GR_POT", 0x100, 1), ("ALL", 0x7f, 11)]:
            a = ab.eval_forces(eph, T0 + 3.7, state, params, forces=mask, gr_eih_sources=src)
            ra = rh.forces(ref, reph, T0 + 3.7, state, params, forces=mask, gr_eih_sources=src)
            af = ab.eval_forces(eph, T0 + 3.7, state, params, forces=mask, gr_eih_sources=src, math=ab.MATH_FAST)
            print("%-11s strict: bit-identical=%s rel(real)=%.2e rel(var)=%.2e | fast: rel(real)=%.2e rel(var)=%.2e" % (
                name, np.array_equal(a, ra), rel(a[:, 0], ra[:, 0]), rel(a[:, 1:], ra[:, 1:]),
                rel(af[:, 0], ra[:, 0]), rel(af[:, 1:], ra[:, 1:])))
        # per-particle integration
        n = 16
        st6 = populations.neo_mba_mix(n, seed=11)
        tend = T0 + 400.0
        t0w = time.time()
        rfin, rts, rdts, rc = rh.integrate_each(ref, reph, T0, st6, tend)
        tref = time.time() - t0w
        for math in (ab.MATH_STRICT, ab.MATH_FAST):
            b = ab.Batch(eph, n, 0, ab.PER_PARTICLE, math=math)
            b.set_state(T0, st6[:, None, :])
            b.integrate(tend)
            g = b.get_state()
            s = b.stats()
            d = np.max(np.linalg.norm(g["state"][:, 0, :3] - rfin[:, 0, :3], axis=-1))
            print("pp integrate math=%d: bit-identical=%s max|dx|=%.3e AU  t ok=%s steps gpu/ref %d/%d iters %d/%d evals %d/%d  kernel %.2f ms ref %.2f s" % (
                math, np.array_equal(g["state"], rfin), d, np.array_equal(g["t"], rts), s["steps"], rc["steps"],
                s["pc_iterations"], rc["pc_iterations"], s["force_evals"], rc["force_evals"], s["last_kernel_ms"], tref))
            b.close()
        # variational per-particle
        n = 8
        stv = populations.with_variations(populations.main_belt(n, seed=13), 6)
        rfin, rts, rdts, rc = rh.integrate_each(ref, reph, T0, stv, T0 + 200.0)
        b = ab.Batch(eph, n, 6, ab.PER_PARTICLE)
        b.set_state(T0, stv)
        b.integrate(T0 + 200.0)
        g = b.get_state()
        print("pp variational: bit-identical=%s rel=%.2e" % (np.array_equal(g["state"], rfin), rel(g["state"].reshape(n, -1), rfin.reshape(n, -1))))
        b.close()
        # shared step
        n = 40
        st6 = populations.main_belt(n, seed=17)
        s = rh.Sim(ref, reph, T0, st6, forces=0x77)
        s.integrate(T0 + 300.0)
        rfin = s.state(); rcnt = s.counters()
        b = ab.Batch(eph, n, 0, ab.SHARED_STEP, forces=0x77)
        b.set_state(T0, st6[:, None, :])
        b.integrate(T0 + 300.0)
        g = b.get_state()
        sg = b.stats()
        print("shared-step: bit-identical=%s max|dx|=%.3e t=%r/%r dt=%r/%r steps %d/%d kernel %.2f ms" % (
            np.array_equal(g["state"], rfin), np.max(np.abs(g["state"] - rfin)), g["t"][0], s.t, g["dt"][0], s.dt,
            sg["steps"], rcnt["steps"], sg["last_kernel_ms"]))
        b.close(); s.close()
        # dense output
        n = 6
        stc, prm = populations.comets(n, seed=19)
        times = T0 - 10.0 * np.arange(1, 40)
        rd = rh.dense_each(ref, reph, T0, stc, times, params=prm[:, None, :])
        b = ab.Batch(eph, n, 0, ab.PER_PARTICLE)
        b.set_state(T0, stc[:, None, :], params=prm[:, None, :])
        gd = b.integrate_or_interpolate(times)
        print("dense (comets, backward): bit-identical=%s max|d|=%.3e" % (np.array_equal(gd, rd), np.nanmax(np.abs(gd - rd))))
        b.close()
        eph.close()
    print("fp64 peak (DFMA) TFLOP/s:", ab._lib.load().assist_gpu_measure_fp64_peak(2000))


if __name__ == "__main__":
    main()
