/*
 * kernels.cu -- sm_100a kernels.  Compiled once per (math variant, translation unit):
 *   -DAB_STRICT=1 --fmad=false   namespace ab_strict : no FMA contraction, reference op order
 *   -DAB_STRICT=0 --fmad=true    namespace ab_fast   : FMA contraction allowed
 *   -DAB_TU=0  ephem_eval_kernel, force_eval_kernel     (assist_get_particle / assist_additional_forces / parity)
 *   -DAB_TU=1  pp_* kernels for systems without variational particles (state in registers)
 *   -DAB_TU=2  pp_* kernels for systems with up to AB_NVMAX variational particles
 *   -DAB_TU=3  sh_* kernels: shared-step IAS15 (persistent cooperative kernel) + dense output
 *   -DAB_TU=4  pp_coop_kernel: per-particle IAS15, a CTA steps 32 systems together with their working set on chip
 * The launchers at the bottom are what gpu_api.cu calls.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#ifndef AB_TU
#error "AB_TU must be defined"
#endif
#define AB_CAT3_(a, b, c) a##_##b##_tu##c
#define AB_CAT3(a, b, c) AB_CAT3_(a, b, c)
#define AB_CAT2_(a, b) a##_##b
#define AB_CAT2(a, b) AB_CAT2_(a, b)
#if AB_STRICT
#define AB_SFX strict
#else
#define AB_SFX fast
#endif
/* one namespace per (variant, translation unit): device helpers defined in headers stay private to it */
#define AB_NS AB_CAT3(ab, AB_SFX, AB_TU)

#include "device_types.h"
#include "ephem_device.cuh"
#include "forces_device.cuh"
#include "ias15_device.cuh"
#include "pp_common.cuh"
#if AB_TU == 4
#include "coop_roles.cuh"
#endif
#include "launchers.h"

namespace AB_NS {

/* Out-of-line copies keep one instance of the two big routines per kernel. */
__device__ __noinline__ void ab_body_states_ol(const AbEphem& E, const AbForceOpts& F, double t, AbBodies& B) {
    ab_body_states(E, F, t, B);
}
template <int KM, class BT>
__device__ __noinline__ void ab_forces_ol(const AbEphem& E, const AbForceOpts& F, const BT& B, AbSysT<KM>& S) {
    ab_forces<KM, BT>(E, F, B, S);
}

#if AB_TU == 0
/* ------------------------------------------------------------------------ */
/* ephemeris + force evaluation kernels                                     */
/* ------------------------------------------------------------------------ */

/* out[n_t][nbodies][10], status[n_t][nbodies]; one thread per (time, body). */
__global__ void ephem_eval_kernel(const __grid_constant__ AbEphem E, const double* __restrict__ t, int n_t,
                                  double* __restrict__ out, int* __restrict__ status) {
    const int nb = AB_NPLANETS + E.n_ast + E.n_ast_x;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_t * nb) return;
    const int it = (int)(gid / nb);
    const int body = (int)(gid % nb);
    const double tt = t[it];
    double GM = 0.0, x[3] = {0, 0, 0}, v[3], a[3];
    int flag;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (body < AB_NPLANETS) {
        flag = ab_planet<2>(E, body, tt, &GM, x, v, a);
    } else {
        /* asteroid: heliocentric SPK position + Sun; velocities are not provided (src/forces.c:208-224) */
        flag = ab_asteroid(E, body - AB_NPLANETS, tt, &GM, x);
        if (flag == AB_OK) {
            double GMs, xs[3], vs[3], as[3];
            flag = ab_planet<0>(E, 0, tt, &GMs, xs, vs, as);
            x[0] += xs[0]; x[1] += xs[1]; x[2] += xs[2];
        }
        v[0] = v[1] = v[2] = nan; a[0] = a[1] = a[2] = nan;
    }
    double* o = out + gid * 10;
    o[0] = GM; o[1] = x[0]; o[2] = x[1]; o[3] = x[2];
    o[4] = v[0]; o[5] = v[1]; o[6] = v[2]; o[7] = a[0]; o[8] = a[1]; o[9] = a[2];
    status[gid] = flag;
}

/* state[n][K][6], params[n][K][3] (or null), acc[n][K][3]; one thread per system. */
__global__ void force_eval_kernel(const __grid_constant__ AbEphem E, const __grid_constant__ AbForceOpts F,
                                  int n, int K, const double* __restrict__ t, int t_per_system,
                                  const double* __restrict__ state, const double* __restrict__ params,
                                  double* __restrict__ acc, int* __restrict__ status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    AbSysT<AB_KMAX> S;
    S.nv_ = K - 1;
    for (int j = 0; j < K; j++) {
        for (int c = 0; c < 3; c++) {
            S.x[j][c] = state[(i * K + j) * 6 + c];
            S.v[j][c] = state[(i * K + j) * 6 + 3 + c];
            S.prm[j][c] = params ? params[(i * K + j) * 3 + c] : 0.0;
        }
    }
    AbBodies B;
    ab_body_states_ol(E, F, t_per_system ? t[i] : t[0], B);
    ab_zero_acc(S);
    if (B.status == AB_OK) ab_forces_ol<AB_KMAX, AbBodies>(E, F, B, S);
    for (int j = 0; j < K; j++)
        for (int c = 0; c < 3; c++) acc[(i * K + j) * 3 + c] = S.a[j][c];
    status[i] = B.status;
}

/* One SPK target at one time, straight from the packed image (reference src/spk.c:492-547 assist_spk_target_pos,
 * :405-481 assist_spk_calc, :590-597 unit conversion).  mode 0: the file's units (km, km/s, km/s^2);
 * mode 1: AU, AU/day, AU/day^2 with the divisors u_d / u_rd, `emb` added first when has_emb (targets 399 / 301);
 * mode 2: position / 149597870.7 (the small-body convention).  out[9] = u[3] v[3] w[3]. */
__global__ void spk_target_kernel(const double* __restrict__ img, const AbSpkTarget tg, int has_emb, const AbSpkTarget emb,
                                  double jd_ref, double t, int mode, double ud0, double ud1, double ud2,
                                  double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double u[3], v[3], w[3];
    ab_spk_target_pos<2>(img, tg, jd_ref, t, u, v, w);
    if (has_emb) {
        double e[3], ev[3], ew[3];
        ab_spk_target_pos<2>(img, emb, jd_ref, t, e, ev, ew);
        for (int i = 0; i < 3; i++) { u[i] += e[i]; v[i] += ev[i]; w[i] += ew[i]; }
    }
    for (int i = 0; i < 3; i++) {
        if (mode == 1) { u[i] = u[i] / ud0; v[i] = v[i] / ud1; w[i] = w[i] / ud2; }
        else if (mode == 2) { u[i] = AB_DIVK(u[i], 149597870.7); }
        out[i] = u[i]; out[3 + i] = v[i]; out[6 + i] = w[i];
    }
}

/* One column of a DE-binary record given by the caller (reference src/ascii_ephem.c:27-65 assist_ascii_work):
 * P[niv][ncm][ncf] coefficients, t0 the fraction of the record, t1 its length in days.  out[3][ncm]. */
__global__ void ascii_work_kernel(const double* __restrict__ P, int ncm, int ncf, int niv, double t0, double t1,
                                  double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double T[32], S[32], U[32];
    const double t = t0 * (double)niv;
    const int b = (int)t;
    const double z = 2.0 * (t - (double)b) - 1.0;          /* == 2 fmod(t, 1) - 1 for t >= 0, exactly */
    const double c = (double)(niv * 2) / t1 / 86400.0;
    T[0] = 1.0; T[1] = z; S[0] = 0.0; S[1] = 1.0; U[0] = 0.0; U[1] = 0.0; U[2] = 4.0;
    for (int p = 2; p < ncf; p++) {
        T[p] = 2.0 * z * T[p - 1] - T[p - 2];
        S[p] = 2.0 * z * S[p - 1] + 2.0 * T[p - 1] - S[p - 2];
    }
    for (int p = 3; p < ncf; p++) U[p] = 2.0 * z * U[p - 1] + 4.0 * S[p - 1] - U[p - 2];
    for (int m = 0; m < ncm; m++) {
        double u = 0.0, v = 0.0, w = 0.0;
        const int n = ncf * (m + b * ncm);
        for (int p = 0; p < ncf; p++) {
            u += T[p] * P[n + p];
            v += S[p] * P[n + p] * c;
            w += U[p] * P[n + p] * c * c;
        }
        out[m] = u; out[ncm + m] = v; out[2 * ncm + m] = w;
    }
}
#endif  /* AB_TU == 0 */

#if AB_TU == 1 || AB_TU == 2
/* ------------------------------------------------------------------------ */
/* per-particle-dt IAS15                                                    */
/* ------------------------------------------------------------------------ */
#if AB_TU == 1
#define PP_KM 1
#else
#define PP_KM AB_KMAX
#endif

/* One reb_simulation_step of system i: force evaluation at the current state, then IAS15 attempts until one is
 * accepted.  Each thread owns its times, so the body tables are evaluated per thread, in one of two ways:
 *   NODES = false  one table per force evaluation (any configuration);
 *   NODES = true   the common configuration (one EIH source, barycentric): the body tables of the 8 times of an attempt
 *                  are evaluated once, side by side (ab_fill_nodes), and every predictor-corrector sweep reuses them --
 *                  what the reference's 7-slot time cache does.
 * One text for both: the control flow (sweeps, convergence, step-size control, rejection) is REBOUND's and must not
 * drift apart between the two. */
template <int KM, bool NODES>
__device__ __forceinline__ void pp_step_impl(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, long long i, PPState& P) {
    AbSysT<KM> S;
    AbBodies B;                              /* NODES = false */
    AbNode nodes[NODES ? AB_NT : 1];         /* NODES = true */
    double times[AB_NT];
    ab_load_sys(Bt, i, S);
    const int nv = S.nv();
    bool have_a0 = false;
    while (true) {   /* attempts */
        const double t_beginning = P.t;
        if constexpr (NODES) {
            times[0] = t_beginning;
            for (int nn = 1; nn < 8; nn++) times[nn] = t_beginning + P.dt * c_h[nn];
            const int flag = ab_fill_nodes(E, F, times, nodes);
            if (flag != AB_OK) { P.status = 1; Bt.status[i] = 1000 + flag; return; }
        }
        if (!have_a0) {      /* the force at the start of the step: once, a rejected attempt keeps it */
            ab_zero_acc(S);
            if constexpr (NODES) {
                ab_forces_ol<KM, AbNode>(E, F, nodes[0], S);
            } else {
                ab_body_states_ol(E, F, P.t, B);
                if (B.status != AB_OK) { P.status = 1; Bt.status[i] = 1000 + B.status; return; }
                ab_forces_ol<KM, AbBodies>(E, F, B, S);
            }
            P.evals++;
            ab_store_a0(Bt, i, S);
            have_a0 = true;
        }
        ab_attempt_begin(Bt, i, nv);
        double pc_err = 1e300, pc_err_last = 2;
        int iterations = 0;
        while (true) {
            if (pc_err < 1e-16) break;
            if (iterations > 2 && pc_err_last <= pc_err) break;
            if (iterations >= 12) break;
            pc_err_last = pc_err;
            pc_err = 0;
            iterations++;
            P.iters++;
            double maxak = 0.0, maxb6 = 0.0;
            for (int nn = 1; nn < 8; nn++) {
                ab_predict(Bt, i, nn, P.dt, S);
                if constexpr (NODES) {
                    ab_zero_acc(S);
                    ab_forces_ol<KM, AbNode>(E, F, nodes[nn], S);
                } else {
                    ab_body_states_ol(E, F, t_beginning + P.dt * c_h[nn], B);
                    if (B.status != AB_OK) { P.status = 1; Bt.status[i] = 1000 + B.status; return; }
                    ab_zero_acc(S);
                    ab_forces_ol<KM, AbBodies>(E, F, B, S);
                }
                P.evals++;
                ab_update_gb(Bt, i, nn, S, maxak, maxb6);
            }
            pc_err = maxb6 / maxak;
        }
        const double dt_done = P.dt;
        if (Bt.epsilon > 0) {
            double maxa = 0.0, maxj = 0.0;
            ab_dt_monitor(Bt, i, S, P.dt, maxa, maxj);
            double dt_new = ab_dt_new(Bt.epsilon, Bt.min_dt, maxa, maxj, dt_done);
            if (fabs(dt_new / dt_done) < 0.25) {
                ab_restore(Bt, i, nv);
                P.dt = dt_new;
                if (P.dt_last != 0.) ab_predict_next(Bt, i, nv, P.dt / P.dt_last, Bt.er, Bt.br);
                P.rejected++;
                continue;
            }
            if (fabs(dt_new / dt_done) > 1.0) {
                if (dt_new / dt_done > 1. / 0.25) dt_new = dt_done / 0.25;
            }
            P.dt = dt_new;
        }
        ab_advance(Bt, i, nv, dt_done);
        P.t += dt_done;
        P.dt_last = dt_done;
        ab_predict_next(Bt, i, nv, P.dt / dt_done, Bt.e, Bt.b);
        P.steps++;
        return;
    }
}
template <int KM>
__device__ __noinline__ void pp_step(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, long long i, PPState& P) {
    pp_step_impl<KM, false>(E, F, Bt, i, P);
}
template <int KM>
__device__ void pp_step_nodes(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, long long i, PPState& P) {
    pp_step_impl<KM, true>(E, F, Bt, i, P);
}

/* reb_simulation_integrate(tmax) for one system, at most `step_cap` steps in this call.
 * Returns true when integrate() has returned (status >= 0), false when it was paused. */
template <int KM>
__device__ bool pp_integrate_to(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, long long i, PPState& P,
                                double tmax, int exact_finish_time, bool resume, long long step_cap) {
    if (!resume) {
        if (tmax != P.t) P.dt = copysign(P.dt, (tmax > P.t) ? 1.0 : -1.0);
        P.last_full_dt = P.dt;
        P.dt_last = 0.;
        P.status = -1;
    }
    long long done = 0;
    while (true) {
        if (step_cap > 0 && done >= step_cap) return false;
        if (ab_check_exit(P.t, P.dt, P.dt_last, P.status, tmax, exact_finish_time, P.last_full_dt) >= 0) break;
        if (F.gr_eih_sources == 1 && !F.geocentric) pp_step_nodes<KM>(E, F, Bt, i, P);
        else pp_step<KM>(E, F, Bt, i, P);
        done++;
    }
    if (exact_finish_time == 1) P.dt = P.last_full_dt;
    return true;
}

#ifndef AB_PP_BLOCK
#define AB_PP_BLOCK 128
#endif
#ifndef AB_PP_LOCKSTEP
#define AB_PP_LOCKSTEP 1
#endif
/* Thread -> system through an optional list of still-running systems, so that a relaunch packs
 * the stragglers into full warps.  The warps of a CTA start every step together (one barrier per
 * step): the kernel is bound by instruction fetch (the hot code is ~200 KB; ncu: SM i-cache hit
 * rate 70 %, GPC instruction-cache requests at 86 % of peak), and warps that walk through the
 * same code at the same time share the fetched lines. */
__global__ void __launch_bounds__(AB_PP_BLOCK, AB_PP_MIN_BLOCKS)
pp_integrate_kernel(const __grid_constant__ AbEphem E, const __grid_constant__ AbForceOpts F,
                    const __grid_constant__ AbBatch Bt, double tmax, int exact_finish_time, int resume,
                    long long step_cap, const int* __restrict__ active, int n_active) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool running = false;
    long long i = 0;
    PPState P;
    if (tid < n_active) {
        i = active ? active[tid] : tid;
        pp_load(Bt, i, P);
        running = !(P.status >= 1000) && !(resume && P.status >= 0);
        if (running && !resume) {      /* entry of reb_simulation_integrate */
            if (tmax != P.t) P.dt = copysign(P.dt, (tmax > P.t) ? 1.0 : -1.0);
            P.last_full_dt = P.dt;
            P.dt_last = 0.;
            P.status = -1;
        }
    }
    const bool owns = running;
    for (long long done = 0;; done++) {
        if (step_cap > 0 && done >= step_cap) break;
#if AB_PP_LOCKSTEP
        if (step_cap > 0) __syncthreads();
        else if (!__syncthreads_or(running ? 1 : 0)) break;
#else
        if (!running) break;
#endif
        if (running) {
            if (ab_check_exit(P.t, P.dt, P.dt_last, P.status, tmax, exact_finish_time, P.last_full_dt) >= 0) {
                running = false;
                if (exact_finish_time == 1) P.dt = P.last_full_dt;
            } else if (F.gr_eih_sources == 1 && !F.geocentric) {
                pp_step_nodes<PP_KM>(E, F, Bt, i, P);
            } else {
                pp_step<PP_KM>(E, F, Bt, i, P);
            }
        }
    }
    if (owns) pp_store(Bt, i, P);
}

/* Time slices: the span of the call is cut into n_win windows [.., origin + (w + 1) * wlen) and the queue
 * hands out (window, system) pairs, window-major.  A thread takes a system through one window and puts
 * it back; the next window of that system is picked up by whichever thread gets to it (after `done[sys]`
 * says the previous one is stored).  Slicing never touches the step sequence of a system -- a step that
 * has started is finished, and the next window resumes exactly where integrate() was paused -- but it
 *   - keeps every lane busy to the end when systems differ 10x in step count or are few (comets), and
 *   - keeps the lanes of a warp within the same few Chebyshev records of the ephemeris. */

/* Work-queue scheduling: a fixed grid of resident threads; every thread takes the next (window, system)
 * pair from a global counter, moves the system's state into a working slot (slot-indexed arrays:
 * coalesced across the warp whatever the system indices are), runs reb_simulation_integrate for it up to
 * the end of the window, writes it back and takes the next pair.  All threads of a CTA do one step per
 * loop trip (one barrier per trip, see pp_integrate_kernel), so lanes only diverge inside a step.
 *
 * With `times` the kernel runs assist_integrate_or_interpolate for the n_times epochs of every system
 * (out[n_times][n][K][6]) instead of one integrate(tmax): the same loop, with the epochs that fall
 * inside the last completed step written out between steps. */
__global__ void __launch_bounds__(AB_PP_BLOCK, AB_PP_MIN_BLOCKS)
pp_queue_kernel(const __grid_constant__ AbEphem E, const __grid_constant__ AbForceOpts F,
                const __grid_constant__ AbBatch Bt, const __grid_constant__ AbBatch W,
                double tmax, int exact_finish_time, unsigned long long* __restrict__ queue_head,
                const __grid_constant__ AbSlices SL, const double* __restrict__ times, int n_times, double* __restrict__ out) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool dense = (times != nullptr);
    const double wsign = (SL.wlen < 0.0) ? -1.0 : 1.0;
    const unsigned long long n_items = (unsigned long long)SL.n_win * (unsigned long long)Bt.n;
    bool have = false, pending = false, exhausted = (slot >= W.n), integrating = false;
    long long sys = -1;
    int nv = 0, ep = 0, win = 0;
    double target = tmax, wend = 0.0;
    long long attempts = 0;
    PPState P;
    while (true) {
        /* Phase A, per lane: get to a point where a step is due.  A lane whose system pauses or finishes stores it
         * and takes the next one in the same trip (two attempts), so that it does not sit out a step of its CTA. */
        bool step_due = false;
        for (int attempt = 0; attempt < 2 && !step_due; attempt++) {
        while (!have && !exhausted) {
            if (!pending) {
                const unsigned long long q = atomicAdd(queue_head, 1ULL);
                if (q >= n_items) { exhausted = true; break; }
                win = (int)(q / (unsigned long long)Bt.n);
                sys = (long long)(q % (unsigned long long)Bt.n);
                if (SL.order) sys = SL.order[sys];
                pending = true;
            }
            /* the previous window of this system must have been stored (it was handed out earlier, so it is running or done) */
            if (win > 0 && *((volatile int*)(SL.done + sys)) < win) break;       /* try again after the next trip */
            __threadfence();
            pending = false;
            pp_load_cg(Bt, sys, P);
            wend = SL.origin + (double)(win + 1) * SL.wlen;
            const bool last = (win == SL.n_win - 1);
            const bool beyond = !last && wsign * P.t >= wsign * wend;            /* nothing to do in this window */
            bool skip = (P.status >= 1000);                                       /* failed earlier (ephemeris error) */
            if (!skip && !dense) {
                if (win == 0) {
                    pp_integrate_entry(P, tmax);                                  /* every system enters integrate() in window 0 */
                    if (beyond) { pp_store(Bt, sys, P); skip = true; }
                } else if (P.status >= 0 || beyond) {
                    skip = true;                                                  /* integrate() has returned / resumes later */
                }
                integrating = true;
            } else if (!skip) {
                ep = (win == 0) ? 0 : __ldcg(SL.epoch + sys);
                integrating = (win > 0 && (P.status == -1 || P.status == -2));    /* paused inside an integrate(times[ep]) */
                if (integrating) target = times[ep];
                if (beyond || (!integrating && ep >= n_times)) {
                    if (win == 0) SL.epoch[sys] = 0;
                    skip = true;
                }
            }
            if (skip) {
                if (SL.n_win > 1) { __threadfence(); *((volatile int*)(SL.done + sys)) = win + 1; }
                continue;
            }
            nv = Bt.nv[sys];
            W.nv[slot] = nv;
            W.status[slot] = 0;
            pp_copy_system<true>(Bt, sys, W, slot, nv);
            have = true;
            attempts = 0;
        }
        if (!have) break;
        /* bookkeeping until a step is due, the window ends or the system is done */
        const bool last = (win == SL.n_win - 1);
        while (true) {
            if (dense && !integrating) {
                /* epochs inside the last completed step are interpolated; the first one outside starts an integrate() */
                while (ep < n_times) {
                    target = times[ep];
                    const double dts = copysign(1., P.dt_last);
                    if (dts * (P.t - P.dt_last) > dts * target || dts * target > dts * P.t || P.dt_last == 0.0) {
                        pp_integrate_entry(P, target);
                        integrating = true;
                        break;
                    }
                    pp_emit(W, slot, nv, P, target, out + ((long long)ep * Bt.n + sys) * Bt.K * 6);
                    ep++;
                }
            }
            if (integrating && !last && wsign * P.t >= wsign * wend) break;      /* end of the window: pause integrate() */
            if (integrating && ab_check_exit(P.t, P.dt, P.dt_last, P.status, target, exact_finish_time, P.last_full_dt) < 0) {
                /* a step that cannot advance the time, or a system over its budget of attempts, is retired with an
                 * error instead of holding the resident grid forever (message 7 of assist_error_messages) */
                if (P.dt == 0.0 || (SL.attempt_budget > 0 && attempts >= SL.attempt_budget)) {
                    P.status = 1;
                    W.status[slot] = 1000 + 7;
                    continue;
                }
                step_due = true;
                break;
            }
            if (integrating) {           /* integrate() returns */
                if (exact_finish_time == 1) P.dt = P.last_full_dt;
                integrating = false;
                if (dense) {
                    pp_emit(W, slot, nv, P, target, out + ((long long)ep * Bt.n + sys) * Bt.K * 6);
                    ep++;
                    continue;
                }
            }
            break;                       /* finished */
        }
        if (!step_due) {
            /* paused at the end of the window, or finished: back to the population arrays */
            pp_copy_system<false>(W, slot, Bt, sys, nv);
            if (W.status[slot] >= 1000) Bt.status[sys] = W.status[slot];
            pp_store(Bt, sys, P);
            if (dense) SL.epoch[sys] = ep;
            if (SL.n_win > 1) { __threadfence(); *((volatile int*)(SL.done + sys)) = win + 1; }
            have = false;
        }
        }   /* attempts */
        /* Phase B, the CTA together: one step.  The barrier also brings back together the lanes that left phase A
         * through different exits -- otherwise the compiler runs the step (99 % of a trip) once per group. */
        if (!__syncthreads_or((have || pending) ? 1 : 0)) break;
        __syncwarp();
        if (step_due) {
            attempts++;
            if (F.gr_eih_sources == 1 && !F.geocentric) pp_step_nodes<PP_KM>(E, F, W, slot, P);
            else pp_step<PP_KM>(E, F, W, slot, P);
        }
    }
}

/* assist_integrate_or_interpolate(times[e]) for e = 0..n_times-1 per system
 * (reference src/assist.c:642-680); out[n_times][n][K][6]. */
__global__ void __launch_bounds__(AB_BLOCK, AB_PP_MIN_BLOCKS)
pp_dense_kernel(const __grid_constant__ AbEphem E, const __grid_constant__ AbForceOpts F,
                const __grid_constant__ AbBatch Bt, const double* __restrict__ times, int n_times,
                double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = Bt.n;
    if (i >= n) return;
    PPState P;
    pp_load(Bt, i, P);
    if (P.status >= 1000) return;
    const int K = Bt.K;
    const int nv = Bt.nv[i];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int e = 0; e < n_times; e++) {
        const double t = times[e];
        double* o = out + ((long long)e * n + i) * K * 6;
        const double dts = copysign(1., P.dt_last);
        if (dts * (P.t - P.dt_last) > dts * t || dts * t > dts * P.t || P.dt_last == 0.0) {
            pp_integrate_to<PP_KM>(E, F, Bt, i, P, t, 0, false, 0);
        }
        const double h = 1.0 - (P.t - t) / P.dt_last;
        if (P.status > 0) {
            for (int q = 0; q < 6 * (1 + nv); q++) o[q] = nan;
        } else if (P.t - t == 0.) {
            for (int j = 0; j <= nv; j++)
                for (int c = 0; c < 3; c++) {
                    o[6 * j + c] = AB1(Bt.pos, 3 * j + c);
                    o[6 * j + 3 + c] = AB1(Bt.vel, 3 * j + c);
                }
        } else if (h < 0.0 || h >= 1.0 || !ab_isnormal(h)) {
            for (int q = 0; q < 6 * (1 + nv); q++) o[q] = nan;
        } else {
            ab_interpolate(Bt, i, nv, P.dt_last, h, o);
        }
    }
    pp_store(Bt, i, P);
}
#endif  /* AB_TU == 1 || AB_TU == 2 */

#if AB_TU == 3
/* ------------------------------------------------------------------------ */
/* shared-step IAS15 (REBOUND semantics for one N-particle simulation)      */
/* ------------------------------------------------------------------------ */

__device__ __forceinline__ void grid_barrier(AbShared* sh, unsigned long long& phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        phase++;
        __threadfence();
        atomicAdd(&sh->barrier, 1ULL);
        const unsigned long long target = phase * gridDim.x;
        while (*((volatile unsigned long long*)&sh->barrier) < target) { }
        __threadfence();
    }
    __syncthreads();
}

/* Global max of two non-negative doubles over the whole grid. */
__device__ void grid_max2(AbShared* sh, unsigned long long& phase, unsigned long long& rcount, double& a, double& b) {
    __shared__ double s_a[AB_BLOCK / 32], s_b[AB_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_a[w] = a; s_b[w] = b; }
    __syncthreads();
    const int slot = (int)(rcount & 7ULL);
    if (threadIdx.x == 0) {
        double ma = s_a[0], mb = s_b[0];
        for (int q = 1; q < AB_BLOCK / 32; q++) { ma = fmax(ma, s_a[q]); mb = fmax(mb, s_b[q]); }
        atomicMax(&sh->red[slot][0], (unsigned long long)__double_as_longlong(ma));
        atomicMax(&sh->red[slot][1], (unsigned long long)__double_as_longlong(mb));
    }
    grid_barrier(sh, phase);
    a = __longlong_as_double((long long)*((volatile unsigned long long*)&sh->red[slot][0]));
    b = __longlong_as_double((long long)*((volatile unsigned long long*)&sh->red[slot][1]));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int z = (slot + 4) & 7;
        sh->red[z][0] = 0ULL; sh->red[z][1] = 0ULL;
    }
    rcount++;
}

/* Body tables for the start of the step (slot 0) or its 7 nodes (slots 1..7),
 * evaluated once per CTA; the threads of the CTA split the slots. */
__device__ void sh_fill_bodies(const AbEphem& E, const AbForceOpts& F, AbBodies* sb, double t0, double dt, int first, int last) {
    for (int s = first + (int)threadIdx.x; s <= last; s += blockDim.x) {
        const double ts = (s == 0) ? t0 : (t0 + dt * c_h[s]);
        ab_body_states_ol(E, F, ts, sb[s]);
    }
    __syncthreads();
}

template <int KM>
__device__ void sh_integrate_body(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, AbBodies* sb,
                                  double tmax, int exact_finish_time, long long max_steps, int flags) {
    AbShared* sh = Bt.sh;
    const long long n = Bt.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long phase = 0, rcount = 0;

    /* every thread carries the same copy of the control variables */
    double t = sh->t, dt = sh->dt, dt_last = sh->dt_last;
    int status = -1;
    unsigned long long steps = 0, rejected = 0, iters = 0, evals = 0;
    int err = 0;
    double last_full_dt = dt;
    const bool raw_step = (flags & 2) != 0;     /* exactly one reb_simulation_step, no exit logic */
    if (flags & 1) {
        /* resume an integrate() that was paused after max_steps */
        status = sh->status;
        last_full_dt = sh->last_full_dt;
    } else if (!raw_step) {
        if (tmax != t) dt = copysign(dt, (tmax > t) ? 1.0 : -1.0);
        last_full_dt = dt;
        dt_last = 0.;
    }

    while (raw_step || ab_check_exit(t, dt, dt_last, status, tmax, exact_finish_time, last_full_dt) < 0) {
        /* ---- reb_simulation_step ---- */
        sh_fill_bodies(E, F, sb, t, dt, 0, 0);
        if (sb[0].status != AB_OK) { err = sb[0].status; status = 1; break; }
        for (long long i = gtid; i < n; i += stride) {
            AbSysT<KM> S;
            ab_load_sys(Bt, i, S);
            ab_zero_acc(S);
            ab_forces_ol<KM, AbBodies>(E, F, sb[0], S);
            ab_store_a0(Bt, i, S);
        }
        evals++;
        bool accepted = false;
        while (!accepted) {
            sh_fill_bodies(E, F, sb, t, dt, 1, 7);
            for (int s = 1; s < 8; s++) if (sb[s].status != AB_OK) err = sb[s].status;
            if (err) break;
            for (long long i = gtid; i < n; i += stride) ab_attempt_begin(Bt, i, (KM == 1) ? 0 : Bt.nv[i]);
            double pc_err = 1e300, pc_err_last = 2;
            int iterations = 0;
            double dmaxa = 0.0, dmaxj = 0.0;
            while (true) {
                if (pc_err < 1e-16) break;
                if (iterations > 2 && pc_err_last <= pc_err) break;
                if (iterations >= 12) break;
                pc_err_last = pc_err;
                pc_err = 0;
                iterations++;
                iters++;
                double maxak = 0.0, maxb6 = 0.0;
                dmaxa = 0.0; dmaxj = 0.0;
                for (long long i = gtid; i < n; i += stride) {
                    AbSysT<KM> S;
                    S.nv_ = Bt.nv[i];
                    for (int j = 0; j <= S.nv(); j++)
                        for (int c = 0; c < 3; c++) S.prm[j][c] = Bt.has_params ? AB1(Bt.prm, 3 * j + c) : 0.0;
                    for (int nn = 1; nn < 8; nn++) {
                        ab_predict(Bt, i, nn, dt, S);
                        ab_zero_acc(S);
                        ab_forces_ol<KM, AbBodies>(E, F, sb[nn], S);
                        ab_update_gb(Bt, i, nn, S, maxak, maxb6);
                    }
                    /* step-size monitor uses the node-7 prediction of this (possibly last) sweep */
                    ab_dt_monitor(Bt, i, S, dt, dmaxa, dmaxj);
                }
                evals += 7;
                grid_max2(sh, phase, rcount, maxak, maxb6);
                pc_err = maxb6 / maxak;
            }
            const double dt_done = dt;
            if (Bt.epsilon > 0) {
                grid_max2(sh, phase, rcount, dmaxa, dmaxj);
                double dt_new = ab_dt_new(Bt.epsilon, Bt.min_dt, dmaxa, dmaxj, dt_done);
                if (fabs(dt_new / dt_done) < 0.25) {
                    dt = dt_new;
                    for (long long i = gtid; i < n; i += stride) {
                        const int nv = (KM == 1) ? 0 : Bt.nv[i];
                        ab_restore(Bt, i, nv);
                        if (dt_last != 0.) ab_predict_next(Bt, i, nv, dt / dt_last, Bt.er, Bt.br);
                    }
                    rejected++;
                    continue;
                }
                if (fabs(dt_new / dt_done) > 1.0) {
                    if (dt_new / dt_done > 1. / 0.25) dt_new = dt_done / 0.25;
                }
                dt = dt_new;
            }
            for (long long i = gtid; i < n; i += stride) {
                const int nv = (KM == 1) ? 0 : Bt.nv[i];
                ab_advance(Bt, i, nv, dt_done);
                ab_predict_next(Bt, i, nv, dt / dt_done, Bt.e, Bt.b);
            }
            t += dt_done;
            dt_last = dt_done;
            accepted = true;
        }
        if (err) { status = 1; break; }
        steps++;
        if (raw_step) break;
        if (max_steps > 0 && (long long)steps >= max_steps) break;
    }
    if (!raw_step && exact_finish_time == 1 && !err && status >= 0) dt = last_full_dt;

    if (gtid == 0) {
        sh->t = t; sh->dt = dt; sh->dt_last = dt_last; sh->last_full_dt = last_full_dt;
        sh->status = status;
        sh->steps += steps; sh->rejected += rejected; sh->iters += iters; sh->evals += evals;
        sh->err_status = err;
    }
}

__global__ void __launch_bounds__(AB_BLOCK)
sh_integrate_kernel_k1(const __grid_constant__ AbEphem E, const __grid_constant__ AbForceOpts F,
                       const __grid_constant__ AbBatch Bt, double tmax, int exact_finish_time, long long max_steps, int flags) {
    __shared__ AbBodies sb[8];
    sh_integrate_body<1>(E, F, Bt, sb, tmax, exact_finish_time, max_steps, flags);
}

__global__ void __launch_bounds__(AB_BLOCK)
sh_integrate_kernel_kv(const __grid_constant__ AbEphem E, const __grid_constant__ AbForceOpts F,
                       const __grid_constant__ AbBatch Bt, double tmax, int exact_finish_time, long long max_steps, int flags) {
    __shared__ AbBodies sb[8];
    sh_integrate_body<AB_KMAX>(E, F, Bt, sb, tmax, exact_finish_time, max_steps, flags);
}

/* out[n][K][6] */
__global__ void sh_interpolate_kernel(const __grid_constant__ AbBatch Bt, double dt_last_done, double h, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Bt.n) return;
    ab_interpolate(Bt, i, Bt.nv[i], dt_last_done, h, out + i * Bt.K * 6);
}
#endif  /* AB_TU == 3 */

#if AB_TU == 4
/* ------------------------------------------------------------------------ */
/* per-particle IAS15, one CTA per 32 systems (coop_device.cuh, coop_roles.cuh) */
/* ------------------------------------------------------------------------ */
__global__ void __launch_bounds__(ABC_THREADS, ABC_CTAS_PER_SM)
pp_coop_kernel(const __grid_constant__ AbEphem E, const __grid_constant__ AbForceOpts F, const __grid_constant__ AbcArgs A) {
    ABC_SM_HERE(sm);
    const int warp = (int)(threadIdx.x >> 5) % ABC_GWARPS;       /* role inside the group */
#if ABC_FILL_TMA
    abc_stage_init();
#endif
    if (warp < 3) abc_comp_main(E, F, A, sm, warp);
    else if (warp == ABC_CTRL_WARP) abc_control_main(E, F, A, sm);
    else abc_worker_main(E, F, A, sm, warp);
}

/* self-test of fp_device.cuh: branch-free division / square root against the built-in operators, bit for bit */
__global__ void fp_selftest_kernel(unsigned long long seed, int iters, unsigned long long* bad) {
    unsigned long long s = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long nd = 0, ns = 0;
    for (int it = 0; it < iters; it++) {
        double v[2];
        for (int k = 0; k < 2; k++) {
            s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
            const unsigned long long r = s * 0x2545F4914F6CDD1DULL;
            unsigned long long mant = r & 0xFFFFFFFFFFFFFULL;
            const int pat = (int)((r >> 52) & 7);
            if (pat == 0) mant = 0; else if (pat == 1) mant = 0xFFFFFFFFFFFFFULL; else if (pat == 2) mant &= 0xFFFFF00000000ULL;
            else if (pat == 3) mant |= 0xFFFFFFFF00000ULL;
            s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
            const unsigned long long r2 = s * 0x2545F4914F6CDD1DULL;
            const unsigned long long ex = 0x280ULL + (r2 % 0x300ULL);
            const unsigned long long sign = (k == 1) ? ((r2 >> 40) & 1ULL) : 0ULL;
            v[k] = __longlong_as_double((long long)((sign << 63) | (ex << 52) | mant));
        }
        if (ab_nb_ok(v[0]) && ab_nb_ok(v[1])) {
            const double q = ab_div_nb(v[1], v[0]), q0 = v[1] / v[0];
            if (__double_as_longlong(q) != __double_as_longlong(q0)) nd++;
            const double r = ab_sqrt_nb(v[0]), r0 = sqrt(v[0]);
            if (__double_as_longlong(r) != __double_as_longlong(r0)) ns++;
        } else {
            nd++; ns++;      /* the generator stays inside the range by construction */
        }
    }
    if (nd) atomicAdd(bad, nd);
    if (ns) atomicAdd(bad + 1, ns);
}
#endif  /* AB_TU == 4 */

}  // namespace AB_NS

/* ------------------------------------------------------------------------ */
/* launchers                                                                */
/* ------------------------------------------------------------------------ */
using namespace AB_NS;

/* every translation unit has its own copy of the IAS15 tables in constant memory */
cudaError_t AB_CAT3(ab_upload_constants, AB_SFX, AB_TU)() {
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_h, AB_H, sizeof(AB_H))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_rr, AB_RR, sizeof(AB_RR))) != cudaSuccess) return e;
    double rri[28];
    for (int k = 0; k < 28; k++) rri[k] = 1.0 / AB_RR[k];
    if ((e = cudaMemcpyToSymbol(c_rri, rri, sizeof(rri))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_c, AB_C, sizeof(AB_C))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_d, AB_D, sizeof(AB_D))) != cudaSuccess) return e;
    return cudaSuccess;
}

#if AB_TU == 0
cudaError_t AB_CAT2(ab_launch_ephem_eval, AB_SFX)(const AbEphem& E, const double* t, int n_t, double* out, int* status, cudaStream_t st) {
    const long long total = (long long)n_t * (AB_NPLANETS + E.n_ast + E.n_ast_x);
    const int grid = (int)((total + AB_BLOCK - 1) / AB_BLOCK);
    ephem_eval_kernel<<<grid, AB_BLOCK, 0, st>>>(E, t, n_t, out, status);
    return cudaGetLastError();
}

cudaError_t AB_CAT2(ab_launch_force_eval, AB_SFX)(const AbEphem& E, const AbForceOpts& F, int n, int K, const double* t, int t_per_system,
                                                  const double* state, const double* params, double* acc, int* status, cudaStream_t st) {
    const int grid = (n + AB_BLOCK - 1) / AB_BLOCK;
    force_eval_kernel<<<grid, AB_BLOCK, 0, st>>>(E, F, n, K, t, t_per_system, state, params, acc, status);
    return cudaGetLastError();
}
cudaError_t AB_CAT2(ab_launch_spk_target, AB_SFX)(const double* img, const AbSpkTarget& tg, int has_emb, const AbSpkTarget& emb, double jd_ref, double t,
                                                  int mode, const double* ud, double* out, cudaStream_t st) {
    spk_target_kernel<<<1, 32, 0, st>>>(img, tg, has_emb, emb, jd_ref, t, mode, ud[0], ud[1], ud[2], out);
    return cudaGetLastError();
}

cudaError_t AB_CAT2(ab_launch_ascii_work, AB_SFX)(const double* P, int ncm, int ncf, int niv, double t0, double t1, double* out, cudaStream_t st) {
    ascii_work_kernel<<<1, 32, 0, st>>>(P, ncm, ncf, niv, t0, t1, out);
    return cudaGetLastError();
}
#endif

#if AB_TU == 1 || AB_TU == 2
#if AB_TU == 1
#define PP_NAME(x) AB_CAT2(x##_k1, AB_SFX)
#else
#define PP_NAME(x) AB_CAT2(x##_kv, AB_SFX)
#endif
cudaError_t PP_NAME(ab_launch_pp_integrate)(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, double tmax, int exact, int resume,
                                            long long step_cap, const int* active, int n_active, cudaStream_t st) {
    const int grid = (n_active + AB_PP_BLOCK - 1) / AB_PP_BLOCK;
    if (grid < 1) return cudaSuccess;
    pp_integrate_kernel<<<grid, AB_PP_BLOCK, 0, st>>>(E, F, Bt, tmax, exact, resume, step_cap, active, n_active);
    return cudaGetLastError();
}

cudaError_t PP_NAME(ab_launch_pp_queue)(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, const AbBatch& W, double tmax, int exact,
                                        unsigned long long* queue_head, const AbSlices& SL, const double* times, int n_times, double* out,
                                        cudaStream_t st) {
    const int grid = (W.n + AB_PP_BLOCK - 1) / AB_PP_BLOCK;
    if (grid < 1) return cudaSuccess;
    pp_queue_kernel<<<grid, AB_PP_BLOCK, 0, st>>>(E, F, Bt, W, tmax, exact, queue_head, SL, times, n_times, out);
    return cudaGetLastError();
}

/* resident threads of the per-particle kernels on the current device (size of the working set) */
cudaError_t PP_NAME(ab_pp_resident_threads)(int* threads) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pp_queue_kernel, AB_PP_BLOCK, 0)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *threads = sms * per_sm * AB_PP_BLOCK;
    return cudaSuccess;
}

cudaError_t PP_NAME(ab_launch_pp_dense)(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, const double* times, int n_times, double* out, cudaStream_t st) {
    const int grid = (Bt.n + AB_BLOCK - 1) / AB_BLOCK;
    pp_dense_kernel<<<grid, AB_BLOCK, 0, st>>>(E, F, Bt, times, n_times, out);
    return cudaGetLastError();
}
#endif

#if AB_TU == 3
cudaError_t AB_CAT2(ab_launch_sh_integrate, AB_SFX)(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, double tmax, int exact, long long max_steps, int flags, cudaStream_t st) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e;
    void* kern = (Bt.K == 1) ? (void*)sh_integrate_kernel_k1 : (void*)sh_integrate_kernel_kv;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (Bt.K == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sh_integrate_kernel_k1, AB_BLOCK, 0);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sh_integrate_kernel_kv, AB_BLOCK, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    int grid = (Bt.n + AB_BLOCK - 1) / AB_BLOCK;
    const int max_grid = sms * per_sm;      /* every CTA must be resident: the kernel spins on a grid barrier */
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    long long ms = max_steps;
    void* args[] = {(void*)&E, (void*)&F, (void*)&Bt, (void*)&tmax, (void*)&exact, (void*)&ms, (void*)&flags};
    return cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(AB_BLOCK), args, 0, st);
}

cudaError_t AB_CAT2(ab_launch_sh_interpolate, AB_SFX)(const AbBatch& Bt, double dt_last_done, double h, double* out, cudaStream_t st) {
    const int grid = (Bt.n + AB_BLOCK - 1) / AB_BLOCK;
    sh_interpolate_kernel<<<grid, AB_BLOCK, 0, st>>>(Bt, dt_last_done, h, out);
    return cudaGetLastError();
}
#endif

#if AB_TU == 4
cudaError_t AB_CAT2(ab_launch_fp_selftest, AB_SFX)(unsigned long long seed, int blocks, int iters, unsigned long long* d_bad, cudaStream_t st) {
    fp_selftest_kernel<<<blocks, 256, 0, st>>>(seed, iters, d_bad);
    return cudaGetLastError();
}

cudaError_t AB_CAT2(ab_pp_coop_max_grid, AB_SFX)(int* max_grid) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(pp_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ABC_SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pp_coop_kernel, ABC_THREADS, ABC_SMEM_BYTES)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *max_grid = sms * per_sm;
    return cudaSuccess;
}

cudaError_t AB_CAT2(ab_launch_pp_coop, AB_SFX)(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, const AbBatch& W, double tmax, int exact,
                                               unsigned long long* queue_head, const AbSlices& SL, const double* times, int n_times, double* out,
                                               const void* plan, const AbSpkTarget* host_ast_tg, double* gtab, unsigned long long* timing, int grid, cudaStream_t st) {
    if (grid < 1) return cudaSuccess;
    {   /* launch-time constants of the fill routine (see coop_device.cuh) */
        cudaError_t ec;
        if ((ec = cudaMemcpyToSymbolAsync(c_abcE, &E, sizeof(AbEphem), 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return ec;
        if ((ec = cudaMemcpyToSymbolAsync(c_abcF, &F, sizeof(AbForceOpts), 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return ec;
        if ((ec = cudaMemcpyToSymbolAsync(c_abcP, plan, sizeof(AbcPlan), 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return ec;
        if (host_ast_tg && E.n_ast > 0 &&
            (ec = cudaMemcpyToSymbolAsync(c_abc_ast, host_ast_tg, sizeof(AbSpkTarget) * E.n_ast, 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return ec;
        if (E.n_ast > 0 && !host_ast_tg) return cudaErrorInvalidValue;
        static thread_local AbSpkTarget series[ABC_NSERIES];
        const int regular = abc_series_table(E, host_ast_tg, series);
        if ((ec = cudaMemcpyToSymbolAsync(c_abc_tg, series, sizeof(series), 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return ec;
        if ((ec = cudaMemcpyToSymbolAsync(c_abc_regular, &regular, sizeof(int), 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return ec;
    }
    AbcArgs A;
    A.Bt = Bt; A.W = W; A.tmax = tmax; A.exact_finish_time = exact; A.queue_head = queue_head; A.SL = SL;
    A.times = times; A.n_times = n_times; A.out = out;
    A.plan = *reinterpret_cast<const AbcPlan*>(plan);
    A.timing = timing;
    A.gtab = gtab;
    cudaError_t e = cudaFuncSetAttribute(pp_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ABC_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    pp_coop_kernel<<<grid, ABC_THREADS, ABC_SMEM_BYTES, st>>>(E, F, A);
    return cudaGetLastError();
}
#endif
