#!/usr/bin/env python3
"""bench.py -- test-particle IAS15 steps/s through the assist-b200 GPU path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own C code on the host cores

Workload (BASELINE.json configs[2], "C3"): synthetic NEO+MBA population, per-particle
adaptive dt, all default forces (mask 0x7F, gr_eih_sources 1), 10 yr forward, synthetic
DE440-format planets .bsp + 16-asteroid .bsp.  One bench "step" = one full pass of the hot
path over the batch (every particle integrated over the whole span).  Weak scaling: every
GPU integrates its own --n-per-gpu particles; no collective on the data path.

`value`   accepted IAS15 particle-steps per second, whole job, inputs resident in HBM
          (device snapshot -> integrate), timed between barriers, max over ranks.
`e2e`     same metric with the step's inputs copied from pinned host memory and the final
          states copied back inside the timed region.
`roofline` FP64: algorithmic flops / kernel time / measured DFMA peak.
`cpu_baseline` the reference's src/*.c (oracle/_ref) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPAN_DAYS = 3652.5
FORCES = 0x7F
MIN_DT = 1e-3
METRIC = "test-particle IAS15 steps/sec (all forces)"
UNIT = "particle-steps/s"

# algorithmic flops, hand-counted from the reference source
F_FORCE_NV0 = 1053.0          # one force evaluation, Nv = 0, no non-grav
F_STEPPER_SUBSTEP = 300.0     # predictor + g/b update per substep (per body)
F_STEP_FINAL = 150.0          # dt control, advance, predict_next per step (per body)


def ephem_flops_per_eval(planet_P, ast_P):
    """Algorithmic flops of one per-particle body-table evaluation: position series for every
    body (T recurrence 3(P-2), three sums 6P, argument 10), velocity too for the Sun."""
    f = 0.0
    for P in planet_P:
        f += 3 * (P - 2) + 6 * P + 10 + 3
    f += 5 * (planet_P[0] - 2) + 9 * planet_P[0]       # Sun velocity (S recurrence + scaled sums)
    for P in ast_P:
        f += 3 * (P - 2) + 6 * P + 10 + 6
    f += 10 * 25                                        # EIH pair sums for one source (10 bodies)
    return f


# ---------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------
# CPU arm: the reference's own C code (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------
def _cpu_worker(job):
    core, planets, asteroids, t0, state, t_end = job
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness as rh
    lib = rh.ref_lib()
    eph = rh.open_ephem(lib, planets, asteroids)
    t_start = time.perf_counter()
    _, _, _, tot = rh.integrate_each(lib, eph, t0, state, t_end, forces=FORCES, min_dt=MIN_DT)
    return tot["steps"], tot["force_evals"], tot["pc_iterations"], time.perf_counter() - t_start


def cpu_reference_run(paths, state, t0, t_end, cores):
    """One simulation per particle, particles split over `cores` pinned processes.  Returns dict."""
    import multiprocessing as mp
    chunks = np.array_split(np.arange(state.shape[0]), cores)
    avail = sorted(os.sched_getaffinity(0))
    jobs = [(avail[c % len(avail)], paths["planets_bsp"], paths["asteroids_bsp"], t0, state[idx], t_end)
            for c, idx in enumerate(chunks) if idx.size]
    ctx = mp.get_context("fork")
    t_start = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t_start
    busy = max(r[3] for r in res)                       # max over processes, excluding pool start-up
    steps = sum(r[0] for r in res)
    return {"steps": steps, "evals": sum(r[1] for r in res), "iters": sum(r[2] for r in res), "wall_s": wall, "busy_s": busy,
            "steps_per_s": steps / busy}


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libassist_ref.so"))


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=1000000)
    ap.add_argument("--math", default="strict", choices=["strict", "fast"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the CPU baseline sample (0 = 512 per core)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --n-per-gpu particles on every GPU; strong: --n-per-gpu particles in total, sliced")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from assist_b200.synth import ephem_writer, populations
    T0 = populations.T0
    T_END = T0 + SPAN_DAYS
    data_dir = os.path.join(ROOT, "data")
    workload = ("C3 NEO+MBA population (20%% NEO / 80%% main belt), per-particle adaptive dt, forces 0x7F, "
                "gr_eih_sources 1, min_dt %g d, %.1f d forward" % (MIN_DT, SPAN_DAYS))

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not have_ref():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libassist_ref.so has not been built"}))
            return 0
        paths = ephem_writer.write_all(data_dir)
        cores = len(os.sched_getaffinity(0))
        n_sample = args.cpu_sample or 512 * cores
        st = populations.neo_mba_mix(args.n_per_gpu, seed=20261703)
        # spread the sample over the population so that it holds the same NEO/MBA mix
        pick = np.linspace(0, args.n_per_gpu - 1, n_sample).astype(np.int64)
        sample = st[pick]
        for _ in range(max(args.warmup, 0) and 1):
            cpu_reference_run(paths, sample[:cores * 2], T0, T_END, cores)
        tot_steps, tot_time = 0, 0.0
        for _ in range(args.steps):
            r = cpu_reference_run(paths, sample, T0, T_END, cores)
            tot_steps += r["steps"]; tot_time += r["busy_s"]
        value = tot_steps / tot_time
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * tot_time / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "n_per_gpu": args.n_per_gpu, "sample_particles": int(n_sample)},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                                 "sample": "%d particles spread over the %d-particle population, full %.1f d span, one simulation per particle, "
                                           "one pinned process per core; reference src/*.c + IAS15 restatement (oracle/_ref)" % (n_sample, args.n_per_gpu, SPAN_DAYS)},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
        sync_tensor = torch.zeros(1, device="cuda")

    def barrier():
        if dist is not None:
            dist.all_reduce(sync_tensor)
            import torch
            torch.cuda.synchronize()

    from assist_b200 import batch as ab
    lib = ab._lib.load()
    if lib.assist_gpu_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the assist-b200 path has no CPU fallback")
    lib.assist_gpu_set_device(local_rank)

    if rank == 0:
        paths = ephem_writer.write_all(data_dir)
    barrier()
    paths = ephem_writer.write_all(data_dir)

    from assist_b200 import sharding
    st = sharding.local_population(populations.neo_mba_mix, args.n_per_gpu, 20261703, world, rank, args.scaling)
    n = st.shape[0]
    math_mode = ab.MATH_FAST if args.math == "fast" else ab.MATH_STRICT
    eph = ab.EphemHandle(paths["planets_bsp"], paths["asteroids_bsp"])
    b = ab.Batch(eph, n, 0, ab.PER_PARTICLE, forces=FORCES, gr_eih_sources=1, min_dt=MIN_DT, math=math_mode)

    # pinned host buffers for the end-to-end leg
    nbytes = n * 6 * 8
    h_in = lib.assist_gpu_host_alloc(nbytes)
    h_out = lib.assist_gpu_host_alloc(nbytes)
    if not h_in or not h_out:
        raise SystemExit("pinned allocation failed: " + lib.assist_gpu_last_error().decode())
    in_arr = np.ctypeslib.as_array(ctypes.cast(h_in, ctypes.POINTER(ctypes.c_double)), shape=(n, 1, 6))
    out_arr = np.ctypeslib.as_array(ctypes.cast(h_out, ctypes.POINTER(ctypes.c_double)), shape=(n, 1, 6))
    in_arr[:, 0, :] = st

    def dptr(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def check(rc, what):
        if rc != 0:
            raise SystemExit("%s failed (%d): %s" % (what, rc, lib.assist_gpu_last_error().decode()))

    check(lib.assist_gpu_batch_set_state(b.ptr, T0, 0.001, dptr(in_arr), None, None), "set_state")
    b.snapshot()

    def device_step():
        b.restore()
        b.integrate(T_END)

    def e2e_step():
        check(lib.assist_gpu_batch_set_state(b.ptr, T0, 0.001, dptr(in_arr), None, None), "set_state")
        check(lib.assist_gpu_batch_integrate(b.ptr, T_END, 1, 0), "integrate")
        check(lib.assist_gpu_batch_get_state(b.ptr, dptr(out_arr), None, None, None, None, None), "get_state")

    for _ in range(args.warmup):
        device_step()
    s0 = b.stats()
    steps_per_pass = s0["steps"]
    evals_per_pass = s0["force_evals"]
    iters_per_pass = s0["pc_iterations"]
    launches_per_pass = s0["kernel_launches"] if args.warmup else None

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_start = time.perf_counter()
    kernel_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        device_step()
        sk = b.stats()
        kernel_ms += sk["last_kernel_ms"]
        launches += sk["kernel_launches"]
        steps_per_pass = sk["steps"]; evals_per_pass = sk["force_evals"]; iters_per_pass = sk["pc_iterations"]
    barrier()
    t_dev = time.perf_counter() - t_start
    clocks = sampler.stop()

    barrier()
    t_start = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t_start

    # parity guard inside the bench: the e2e result equals the device-resident result
    dev_state = b.get_state()["state"]
    if not np.array_equal(dev_state, out_arr):
        raise SystemExit("bench.py: e2e and device-resident paths disagree")

    # reduce over ranks: max time, sum steps
    (t_dev, t_e2e, kernel_ms), (tot_steps, tot_evals, tot_iters, tot_n) = sharding.reduce_max_sum(
        dist, [t_dev, t_e2e, kernel_ms], [float(steps_per_pass), float(evals_per_pass), float(iters_per_pass), float(n)],
        device="cuda" if dist is not None else None)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    K = args.steps
    value = tot_steps * K / t_dev
    e2e_value = tot_steps * K / t_e2e

    # roofline of the fused integrate kernel (this rank): algorithmic flops / kernel time
    peak = lib.assist_gpu_measure_fp64_peak(2000)
    sub_evals = evals_per_pass - steps_per_pass          # evaluations at Gauss-Radau nodes
    flops = evals_per_pass * F_FORCE_NV0 + sub_evals * F_STEPPER_SUBSTEP + steps_per_pass * F_STEP_FINAL
    planet_P = [11, 14, 10, 13, 13, 13, 11, 8, 7, 6, 6, 6]     # Sun, Mer, Ven, EMB, Earth, Moon, Mar..Plu (synthetic DE440 layout)
    f_eph = ephem_flops_per_eval(planet_P, [16] * 16)
    kernel_s = kernel_ms * 1e-3 / K
    achieved = flops / kernel_s / 1e12
    # the body tables are needed at 8 times per step attempt (start + 7 nodes); the reference's 7-slot time cache
    # gives it the same reuse, so the algorithmic ephemeris work is 8 evaluations per step, not one per force call
    achieved_eph = (flops + 8.0 * steps_per_pass * f_eph) / kernel_s / 1e12
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None,
                "kernel": "pp_integrate_kernel (fused ephemeris + forces + IAS15), %d launches per pass" % (launches // K),
                "flops_per_force_eval": F_FORCE_NV0, "achieved_incl_ephemeris": achieved_eph, "frac_incl_ephemeris": achieved_eph / peak,
                "ephemeris_flops_per_table": f_eph, "ephemeris_tables_per_step": 8, "force_evals_per_s": evals_per_pass / kernel_s,
                "pc_iterations_per_step": tot_iters / tot_steps,
                "peak_source": "register-resident DFMA loop measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}

    cpu_baseline = None
    if args.gpus == 1 and not args.no_cpu_baseline and have_ref():
        cores = len(os.sched_getaffinity(0))
        n_sample = args.cpu_sample or 512 * cores
        pick = np.linspace(0, n - 1, n_sample).astype(np.int64)
        r = cpu_reference_run(paths, st[pick], T0, T_END, cores)
        cpu_baseline = {"value": r["steps_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                        "sample": "%d particles spread over the population, full %.1f d span, one simulation per particle, one pinned "
                                  "process per core (%.1f s); reference src/*.c + IAS15 restatement (oracle/_ref)" % (n_sample, SPAN_DAYS, r["busy_s"])}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "n_per_gpu": n, "n_total": int(tot_n), "math": args.math,
                       "ephemeris": "synthetic DE440-layout planets .bsp + 16-asteroid .bsp (JD 2441000.5-2465000.5)",
                       "cache": "per-GPU state %.2f GB >> 126 MB L2, so every pass streams from HBM" % (n * 1.5e3 / 1e9),
                       "particle_steps_per_pass": tot_steps, "parity": "strict math is bit-identical to the reference C build (tests/)"},
            "kernel_ms_per_step": kernel_ms / K,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes)},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
