#!/usr/bin/env python3
"""bench.py -- test-particle IAS15 steps/s through the assist-b200 GPU path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own C code on the host cores

Default workload (BASELINE.json configs[2], "c3", the configuration the metric is quoted on):
synthetic NEO+MBA population, per-particle adaptive dt, all default forces (mask 0x7F,
gr_eih_sources 1), 10 yr forward, synthetic DE440-layout planets .bsp + 16-asteroid .bsp.
The other configs of BASELINE.json are available with --workload c2|c4|c5, the north star's target
configuration (10^6 particles WITH variational equations, 10 yr) with --workload target.
One bench "step" = one full pass of the hot path over the batch (every particle integrated over
the whole span).  Scaling: C3 is the configuration BASELINE.json states as "10^6 ... sharded 1/2/4/8", so its
default is STRONG scaling -- one population of --n particles in total, dealt out round-robin over the GPUs; at
N > 1 the line also carries a `weak` block (that many particles on EVERY GPU, fewer passes).  The other workloads
default to weak scaling.  No collective on the data path (torch.distributed only carries the barrier, a max/sum
of scalars and a gather of per-rank timings).

`value`   accepted IAS15 particle-steps per second, whole job, inputs resident in HBM
          (device snapshot -> integrate), timed between barriers, max over ranks.
`e2e`     same metric with the step's inputs copied from pinned host memory and the results
          copied back inside the timed region.
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

METRIC = "test-particle IAS15 steps/sec (all forces)"
UNIT = "particle-steps/s"

# algorithmic flops, hand-counted from the reference source
F_FORCE = {0: 1053.0, 6: 7540.0}      # one force evaluation of one system, default mask, one EIH source (Nv = 0 / 6)
F_MARSDEN = {0: 92.0, 6: 982.0}       # extra for the non-gravitational term
F_STEPPER_SUBSTEP = 300.0             # predictor + g/b update per substep, per body
F_STEP_FINAL = 150.0                  # dt control, advance, predict_next per step, per body


def ephem_flops_per_table(planet_P, ast_P):
    """Algorithmic flops of one body table (all bodies at one time): position series for every body
    (T recurrence 3(P-2), three sums 6P, argument 10), velocity too for the Sun, EIH pair sums."""
    f = 0.0
    for P in planet_P:
        f += 3 * (P - 2) + 6 * P + 10 + 3
    f += 5 * (planet_P[0] - 2) + 9 * planet_P[0]       # Sun velocity (S recurrence + scaled sums)
    for P in ast_P:
        f += 3 * (P - 2) + 6 * P + 10 + 6
    f += 10 * 25                                        # EIH pair sums for one source (10 bodies)
    return f


def workloads():
    """BASELINE.json configs as (population, batch options, span).  state generators return [n][K][6]."""
    from assist_b200.synth import populations as pop
    return {
        "c3": dict(desc="C3 NEO+MBA population (20% NEO / 80% main belt), per-particle adaptive dt, forces 0x7F, "
                        "gr_eih_sources 1, min_dt 0.001 d, 3652.5 d forward",
                   gen=lambda n, seed: pop.neo_mba_mix(n, seed=seed)[:, None, :], prm=None,
                   seed=20261703, n=1000000, nvar=0, mode="pp", forces=0x7F, min_dt=1e-3, span=3652.5, dense=None, cpu_per_core=512),
        "c2": dict(desc="C2 main-belt population in ONE simulation (shared step: global dt and convergence), forces 0x77, "
                        "gr_eih_sources 1, 3652.5 d forward",
                   gen=lambda n, seed: pop.main_belt(n, seed=seed)[:, None, :], prm=None,
                   seed=20261702, n=10000, nvar=0, mode="shared", forces=0x77, min_dt=0.0, span=3652.5, dense=None, cpu_per_core=300),
        "c4": dict(desc="C4 main-belt particles, each with 6 first-order variational particles (identity initial conditions), "
                        "per-particle dt, forces 0x7F, 1826.25 d forward",
                   gen=lambda n, seed: pop.with_variations(pop.main_belt(n, seed=seed), 6), prm=None,
                   seed=20261704, n=100000, nvar=6, mode="pp", forces=0x7F, min_dt=0.0, span=1826.25, dense=None, cpu_per_core=600),
        "target": dict(desc="north-star target: 10^6 main-belt particles, each with 6 first-order variational particles, per-particle dt, "
                            "forces 0x7F, min_dt 0.001 d, 3652.5 d forward (BASELINE.json north_star, last sentence)",
                       gen=lambda n, seed: pop.with_variations(pop.main_belt(n, seed=seed), 6), prm=None,
                       seed=20261706, n=1000000, nvar=6, mode="pp", forces=0x7F, min_dt=1e-3, span=3652.5, dense=None, cpu_per_core=300),
        "c5": dict(desc="C5 comets with Marsden A1/A2/A3, per-particle dt, forces 0x7F, min_dt 0.001 d, 18262.5 d BACKWARD, "
                        "dense output every 10 d (assist_integrate_or_interpolate semantics)",
                   gen=lambda n, seed: pop.comets(n, seed=seed)[0][:, None, :],
                   prm=lambda n, seed: pop.comets(n, seed=seed)[1][:, None, :],
                   seed=20261705, n=100000, nvar=0, mode="pp", forces=0x7F, min_dt=1e-3, span=-18262.5, dense=10.0, cpu_per_core=64),
    }


def dense_epochs(t0, wl):
    k = np.arange(1, int(abs(wl["span"]) / wl["dense"]) + 1)
    return np.ascontiguousarray(t0 + np.sign(wl["span"]) * wl["dense"] * k)


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
    core, planets, asteroids, t0, state, params, wl = job
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness as rh
    lib = rh.ref_lib()
    eph = rh.open_ephem(lib, planets, asteroids)
    n = state.shape[0]
    kw = dict(forces=wl["forces"], min_dt=wl["min_dt"])
    t_start = time.perf_counter()
    if wl["mode"] == "shared":
        s = rh.Sim(lib, eph, t0, state, **kw)
        s.integrate(t0 + wl["span"])
        c = s.counters()
        steps, evals = c["steps"] * n, c["force_evals"] * n
        s.close()
    elif wl["dense"]:
        times = dense_epochs(t0, wl)
        steps = evals = 0
        for i in range(n):
            s = rh.Sim(lib, eph, t0, state[i:i + 1], params=None if params is None else params[i:i + 1], **kw)
            for t in times:
                lib.assist_integrate_or_interpolate(s.ax, float(t))
            c = s.counters()
            steps += c["steps"]; evals += c["force_evals"]
            s.close()
    else:
        _, _, _, tot = rh.integrate_each(lib, eph, t0, state, t0 + wl["span"], params=params, **kw)
        steps, evals = tot["steps"], tot["force_evals"]
    return steps, evals, time.perf_counter() - t_start


def cpu_reference_run(paths, state, params, t0, wl, cores):
    """One simulation per particle (ONE simulation in all for the shared-step workload, which cannot be split
    without changing its dt sequence), particles split over `cores` pinned processes."""
    import multiprocessing as mp
    plain = {k: wl[k] for k in ("mode", "forces", "min_dt", "span", "dense")}
    if wl["mode"] == "shared":
        cores = 1
    # round-robin, not contiguous: the population is ordered (NEOs first, 6x the steps of a main-belt object),
    # and the reference arm is timed as the slowest process
    chunks = [np.arange(c, state.shape[0], cores) for c in range(cores)]
    avail = sorted(os.sched_getaffinity(0))
    jobs = [(avail[c % len(avail)], paths["planets_bsp"], paths["asteroids_bsp"], t0, state[idx],
             None if params is None else params[idx], plain) for c, idx in enumerate(chunks) if idx.size]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs)
    busy = max(r[2] for r in res)                       # max over processes, excluding pool start-up
    steps = sum(r[0] for r in res)
    return {"steps": steps, "evals": sum(r[1] for r in res), "busy_s": busy, "steps_per_s": steps / busy, "cores": len(jobs)}


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libassist_ref.so"))


def cpu_sample(wl, st, prm, cores, override):
    """Bounded sample of the workload for the CPU arm, spread over the population (same NEO/MBA mix)."""
    n_sample = override or (wl["cpu_per_core"] if wl["mode"] == "shared" else wl["cpu_per_core"] * cores)
    n_sample = min(n_sample, st.shape[0])
    pick = np.linspace(0, st.shape[0] - 1, n_sample).astype(np.int64)
    how = ("ONE %d-particle shared-step simulation on one core" % n_sample if wl["mode"] == "shared" else
           "%d particles spread over the %d-particle population, one simulation per particle, one pinned process per core"
           % (n_sample, st.shape[0]))
    return st[pick], (None if prm is None else prm[pick]), how + ", full span; reference src/*.c + IAS15 restatement (oracle/_ref)"


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4", "c5", "target"])
    ap.add_argument("--n-per-gpu", type=int, default=0, help="particles per GPU (0 = the workload's own size)")
    ap.add_argument("--math", default="strict", choices=["strict", "fast"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the CPU baseline sample (0 = sized for ~10-30 s)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: --n-per-gpu particles on every GPU; strong: that many in total, dealt out round-robin "
                         "(default: strong for c3, weak for the others)")
    ap.add_argument("--no-weak-leg", action="store_true", help="skip the extra weak-scaling measurement of a strong run at N > 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from assist_b200 import sharding
    from assist_b200.synth import ephem_writer, populations
    wl = workloads()[args.workload]
    if args.scaling is None:
        args.scaling = "strong" if args.workload in ("c3", "target") else "weak"
    n_req = args.n_per_gpu or wl["n"]
    T0 = populations.T0
    T_END = T0 + wl["span"]
    data_dir = os.path.join(ROOT, "data")

    def population(world_, rank_, scaling_=None):
        scaling_ = scaling_ or args.scaling
        st_ = sharding.local_population(lambda n, seed: wl["gen"](n, seed), n_req, wl["seed"], world_, rank_, scaling_)
        pr_ = None
        if wl["prm"]:
            pr_ = np.ascontiguousarray(sharding.local_population(lambda n, seed: wl["prm"](n, seed), n_req, wl["seed"], world_, rank_, scaling_))
        return np.ascontiguousarray(st_), pr_

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not have_ref():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libassist_ref.so has not been built"}))
            return 0
        paths = ephem_writer.write_all(data_dir)
        cores = len(os.sched_getaffinity(0))
        st, prm = population(1, 0)
        s_st, s_pr, how = cpu_sample(wl, st, prm, cores, args.cpu_sample)
        if args.warmup:
            k = max(2, cores)
            cpu_reference_run(paths, s_st[:k], None if s_pr is None else s_pr[:k], T0, wl, cores)
        tot_steps, tot_time, used = 0, 0.0, cores
        for _ in range(args.steps):
            r = cpu_reference_run(paths, s_st, s_pr, T0, wl, cores)
            tot_steps += r["steps"]; tot_time += r["busy_s"]; used = r["cores"]
        value = tot_steps / tot_time
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * tot_time / args.steps, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["desc"], "n_per_gpu": n_req, "sample_particles": int(s_st.shape[0])},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "reference", "sample": how},
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
        ephem_writer.write_all(data_dir)
    barrier()
    paths = ephem_writer.write_all(data_dir)

    st, prm = population(world, rank)
    n, K = st.shape[0], st.shape[1]
    shared = wl["mode"] == "shared"
    math_mode = ab.MATH_FAST if args.math == "fast" else ab.MATH_STRICT
    eph = ab.EphemHandle(paths["planets_bsp"], paths["asteroids_bsp"])
    b = ab.Batch(eph, n, wl["nvar"], ab.SHARED_STEP if shared else ab.PER_PARTICLE, forces=wl["forces"], gr_eih_sources=1,
                 min_dt=wl["min_dt"], math=math_mode)
    times = dense_epochs(T0, wl) if wl["dense"] else None

    def dptr(a):
        return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def check(rc, what):
        if rc != 0:
            raise SystemExit("%s failed (%d): %s" % (what, rc, lib.assist_gpu_last_error().decode()))

    def pinned(shape):
        nb = int(np.prod(shape)) * 8
        p = lib.assist_gpu_host_alloc(nb)
        if not p:
            raise SystemExit("pinned allocation failed: " + lib.assist_gpu_last_error().decode())
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=shape), nb

    # pinned host buffers for the end-to-end leg
    in_arr, in_bytes = pinned((n, K, 6))
    in_arr[...] = st
    prm_arr, prm_bytes = None, 0
    if prm is not None:
        prm_arr, prm_bytes = pinned((n, K, 3))
        prm_arr[...] = prm
    out_arr, out_bytes = pinned((n, K, 6) if times is None else (times.size, n, K, 6))

    def set_state():
        check(lib.assist_gpu_batch_set_state(b.ptr, T0, 0.001, dptr(in_arr), dptr(prm_arr), None), "set_state")

    def run(to_host):
        if times is None:
            check(lib.assist_gpu_batch_integrate(b.ptr, T_END, 1, 0), "integrate")
            if to_host:
                check(lib.assist_gpu_batch_get_state(b.ptr, dptr(out_arr), None, None, None, None, None), "get_state")
        else:
            # the epochs entry point returns its [epoch][particle] block to the host buffer in both legs
            check(lib.assist_gpu_batch_integrate_or_interpolate(b.ptr, dptr(times), times.size, dptr(out_arr)), "integrate_or_interpolate")

    def device_step():
        b.restore()
        run(False)

    def e2e_step():
        set_state()
        run(True)

    set_state()
    b.snapshot()
    for _ in range(args.warmup):
        device_step()

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_start = time.perf_counter()
    kernel_ms, launches = 0.0, 0
    launched_before = lib.assist_gpu_kernel_launches()      # every kernel of the library, counted at the launch sites
    sk = None
    for _ in range(args.steps):
        device_step()
        sk = b.stats()
        kernel_ms += sk["last_kernel_ms"]
        launches += sk["kernel_launches"]
    barrier()
    t_dev = time.perf_counter() - t_start
    all_launches = lib.assist_gpu_kernel_launches() - launched_before
    clocks = sampler.stop()
    iters_per_step = sk["pc_iterations"] / max(sk["steps"], 1)
    steps_per_pass = sk["steps"] * (n if shared else 1)    # a shared-step batch reports global steps; the metric counts particle-steps
    evals_per_pass = sk["force_evals"]

    dev_result = None if times is not None else b.get_state()["state"].copy()
    barrier()
    t_start = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t_start
    # guard inside the bench: the end-to-end leg produced the same numbers as the device-resident leg
    if dev_result is not None and not np.array_equal(dev_result, out_arr):
        raise SystemExit("bench.py: e2e and device-resident paths disagree")
    if times is not None and not np.isfinite(out_arr).all():
        raise SystemExit("bench.py: dense output holds non-finite values")

    # what bounds a rank's pass: its longest single system (a serial chain of steps) against the kernel time
    longest = None
    if not shared:
        cnt = b.counters()
        per_sys = (cnt["steps"] + cnt["rejected"]).astype(np.int64)
        longest = int(per_sys.max())
    my_kernel_ms = kernel_ms / args.steps

    # extra leg of a strong-scaled run: the same number of particles on EVERY GPU (weak scaling), fewer passes
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak_leg:
        b.close()
        st_w, prm_w = population(world, rank, "weak")
        bw = ab.Batch(eph, st_w.shape[0], wl["nvar"], ab.SHARED_STEP if shared else ab.PER_PARTICLE, forces=wl["forces"],
                      gr_eih_sources=1, min_dt=wl["min_dt"], math=math_mode)
        bw.set_state(T0, st_w, params=prm_w)
        bw.snapshot()
        passes = min(2, args.steps)
        w_kernel = 0.0

        def weak_pass():
            bw.restore()
            if times is None:
                bw.integrate(T_END)
            else:
                bw.integrate_or_interpolate(times)
        weak_pass()
        barrier()
        t_start = time.perf_counter()
        for _ in range(passes):
            weak_pass()
            w_kernel += bw.stats()["last_kernel_ms"]
        barrier()
        t_w = time.perf_counter() - t_start
        sw = bw.stats()
        (t_w, w_kernel), (w_steps, w_n) = sharding.reduce_max_sum(
            dist, [t_w, w_kernel], [float(sw["steps"] * (st_w.shape[0] if shared else 1)), float(st_w.shape[0])], device="cuda")
        weak = {"value": w_steps * passes / t_w, "unit": UNIT, "n_total": int(w_n), "steps": passes, "warmup": 1,
                "ms_per_step": 1e3 * t_w / passes, "kernel_ms_per_step_max": w_kernel / passes,
                "what": "the same workload with %d particles on EVERY GPU (population seeded per rank)" % st_w.shape[0]}
        bw.close()

    # reduce over ranks: max time, sum steps; per-rank table
    per_rank = sharding.gather_rows(dist, [my_kernel_ms, float(longest or 0), float(n), float(steps_per_pass)],
                                    device="cuda" if dist is not None else None)
    (t_dev, t_e2e, kernel_ms), (tot_steps, tot_n) = sharding.reduce_max_sum(
        dist, [t_dev, t_e2e, kernel_ms], [float(steps_per_pass), float(n)], device="cuda" if dist is not None else None)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    S = args.steps
    value = tot_steps * S / t_dev
    e2e_value = tot_steps * S / t_e2e

    # roofline of the fused integrate kernel (this rank): algorithmic flops / kernel time
    peak = lib.assist_gpu_measure_fp64_peak(2000)
    f_force = F_FORCE[wl["nvar"]] + (F_MARSDEN[wl["nvar"]] if prm is not None else 0.0)
    sub_evals = evals_per_pass - steps_per_pass          # evaluations at Gauss-Radau nodes
    flops = evals_per_pass * f_force + sub_evals * F_STEPPER_SUBSTEP * K + steps_per_pass * F_STEP_FINAL * K
    planet_P = [11, 14, 10, 13, 13, 13, 11, 8, 7, 6, 6, 6]     # Sun, Mer, Ven, EMB, Earth, Moon, Mar..Plu (synthetic DE440 layout)
    f_eph = ephem_flops_per_table(planet_P, [16] * 16)
    kernel_s = kernel_ms * 1e-3 / S
    achieved = flops / kernel_s / 1e12
    # body tables are needed at 8 times per step attempt (start + 7 nodes); the reference's 7-slot time cache gives it the
    # same reuse, so the algorithmic ephemeris work is 8 tables per step.  In a shared-step batch all particles share them.
    eph_flops = 8.0 * f_eph * (sk["steps"] if shared else steps_per_pass)
    achieved_eph = (flops + eph_flops) / kernel_s / 1e12
    # DRAM bytes and FP64-pipe activity of one launch of this kernel: NOT measured in this run -- read from the committed
    # ncu capture of the same workload / size / math (profiles/captures.json names the capture file of each entry); null
    # when no capture of this exact configuration is committed
    capture = None
    try:
        with open(os.path.join(ROOT, "profiles", "captures.json")) as fh:
            capture = json.load(fh).get("%s:%d:%s" % (args.workload, n, args.math))
    except OSError:
        pass
    traffic = capture["dram_bytes_per_launch"] if capture else None
    # the same launch against the HBM roof (MEASURED_PEAKS.json, driver-written): the kernel is far from both
    hbm = None
    if traffic:
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
        except (OSError, KeyError, ValueError):
            hbm_peak, hbm_src = 6650.0, "of fallback (B200_PROFILING.md)"
        hbm = {"achieved_gbs": traffic / kernel_s / 1e9, "peak_gbs": hbm_peak, "frac": traffic / kernel_s / 1e9 / hbm_peak,
               "peak_source": hbm_src}
    coop = (not shared) and wl["nvar"] == 0
    kname = ("sh_integrate_kernel" if shared else
             ("pp_coop_kernel" if coop else "pp_queue_kernel") + (", epoch output" if times is not None else ""))
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": (capture or {}).get("source"),
                "fp64_pipe_active_pct": (capture or {}).get("fp64_pipe_active_pct"), "hbm": hbm,
                "kernel": "fused ephemeris + forces + IAS15 integrate kernel (%s), %d launch(es) per pass" % (kname, launches // S),
                "flops_per_force_eval": f_force, "achieved_incl_ephemeris": achieved_eph, "frac_incl_ephemeris": achieved_eph / peak,
                "ephemeris_flops_per_table": f_eph, "ephemeris_tables_per_step": 8, "force_evals_per_s": evals_per_pass / kernel_s,
                "pc_iterations_per_step": iters_per_step,
                "ceiling": "strict math forbids FMA contraction (reference operation order): an add or a multiply fills a DFMA "
                           "slot with one flop, so `frac` cannot exceed 0.5 by construction; IEEE division and square root "
                           "expand to ~10 FP64 instructions each and count as one flop",
                "peak_source": "register-resident DFMA loop measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}

    cpu_baseline = None
    if args.gpus == 1 and not args.no_cpu_baseline and have_ref():
        cores = len(os.sched_getaffinity(0))
        s_st, s_pr, how = cpu_sample(wl, st, prm, cores, args.cpu_sample)
        r = cpu_reference_run(paths, s_st, s_pr, T0, wl, cores)
        cpu_baseline = {"value": r["steps_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                        "sample": how + " (%.1f s)" % r["busy_s"]}

    state_gb = n * K * (61 * 3 + 4) * 8 / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": S, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / S, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_per_gpu": n, "n_total": int(tot_n), "math": args.math,
                       "ephemeris": "synthetic DE440-layout planets .bsp + 16-asteroid .bsp (JD 2441000.5-2465000.5)",
                       "cache": ("per-GPU batch state %.2f GB >> 126 MB L2, so every pass streams from HBM" % state_gb) if state_gb > 0.5 else
                                ("per-GPU batch state %.3f GB; every pass restarts from a device snapshot and rewrites the whole "
                                 "state, nothing is reused between passes" % state_gb),
                       "particle_steps_per_pass": tot_steps,
                       "parity": "strict math: ephemeris, every force term and the integrated states are bit-identical to the reference's "
                                 "src/*.c driven by this repo's restatement of REBOUND's IAS15 (tests/); the IAS15 stepper itself is "
                                 "UNPINNED against a REBOUND binary (REBOUND is absent offline), see DESIGN.md section 2"},
            "kernel_ms_per_step": kernel_ms / S,
            "per_rank": [{"rank": r, "kernel_ms_per_step": row[0], "longest_system_attempts": int(row[1]), "particles": int(row[2]),
                          "particle_steps_per_pass": int(row[3]),
                          "ms_per_attempt_if_longest_system_bounds_the_pass": (row[0] / row[1]) if row[1] else None}
                         for r, row in enumerate(per_rank)],
            "weak": weak,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes + prm_bytes), "d2h_bytes_per_step": int(out_bytes)},
            "gpu_launches": int(all_launches), "integrate_kernel_launches": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
