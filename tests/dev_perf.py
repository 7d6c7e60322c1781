"""Developer throughput probe: python tests/dev_perf.py [n] [days] [math] [nvar]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from assist_b200 import batch as ab
from assist_b200.synth import ephem_writer, populations


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    days = float(sys.argv[2]) if len(sys.argv) > 2 else 365.25
    math = {"strict": 0, "fast": 1}[sys.argv[3]] if len(sys.argv) > 3 else 0
    nvar = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    min_dt = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-3
    paths = ephem_writer.write_all(os.path.join(ROOT, "data"))
    eph = ab.EphemHandle(paths["planets_bsp"], paths["asteroids_bsp"])
    T0 = populations.T0
    st = populations.neo_mba_mix(n, seed=1)
    state = populations.with_variations(st, nvar) if nvar else st[:, None, :]
    b = ab.Batch(eph, n, nvar, ab.PER_PARTICLE, math=math, min_dt=min_dt)
    b.set_state(T0, state)
    b.snapshot()
    for rep in range(reps):
        b.restore()
        t0 = time.time()
        b.integrate(T0 + days)
        wall = time.time() - t0
        s = b.stats()
        ms = s["last_kernel_ms"]
        print("n=%d nvar=%d days=%g math=%d: kernel %.1f ms (wall %.1f ms) steps=%d rej=%d iters/step=%.2f evals=%d -> %.3e steps/s %.3e evals/s" % (
            n, nvar, days, math, ms, wall * 1e3, s["steps"], s["steps_rejected"], s["pc_iterations"] / max(s["steps"], 1),
            s["force_evals"], s["steps"] / (ms * 1e-3), s["force_evals"] / (ms * 1e-3)))
    c = b.counters()
    st_ = c["steps"].astype(np.int64)
    print("steps per particle: min %d median %d p99 %d max %d (argmax %d); evals max %d" % (
        st_.min(), np.median(st_), np.percentile(st_, 99), st_.max(), int(st_.argmax()), int(c["evals"].max())))
    b.close()


if __name__ == "__main__":
    main()
