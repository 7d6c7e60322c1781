/*
 * coop_roles.cuh -- the three warp roles of pp_coop_kernel and the barrier protocol between them
 * (see coop_device.cuh for the mapping).  One trip of the CTA loop is one IAS15 step ATTEMPT of every slot of both
 * groups; the barriers are CTA-wide, "all" = the 8 warps of each group:
 *
 *   barrier  who            what
 *   -------  -------------  ----------------------------------------------------------------------
 *            control        bookkeeping per slot (queue, integrate() entry / exit, output epochs, windows)
 *   B1 (or)  all            anything left to do?  no -> the CTA retires
 *            all 8 warps    fill: thread = (slot, node): Chebyshev sums of every body at the 8 node times + the EIH pair
 *                           sums of the Sun; planets into shared memory, the rest into the CTA's global table
 *   B2
 *            control        ephemeris errors of the fill, sweep state of the attempt
 *   B4 (or)  all            does any slot need the force evaluation at the start of its step?
 *            components     state from the working batch into registers
 *   [ B5     workers        forces at node 0                     ]  only if B4 said so
 *   [ B6     components     a0, last_state                       ]
 *            components     attempt begin, prediction at node 1
 *   per node n = 1..7:
 *   B7       workers        force terms at node n
 *   B8       components     ordered sums, g/b update, prediction at node n + 1
 *   B9       control        convergence of the sweep
 *   B10 (or) all            another sweep?  yes -> components predict node 1, back to B7
 *            control        step-size control: accept / reject
 *   B11      components     advance + predict_next, or restore; registers back to the working batch
 *   B12
 */
#ifndef AB_COOP_ROLES_CUH
#define AB_COOP_ROLES_CUH

#include "coop_device.cuh"
#include "pp_common.cuh"

namespace AB_NS {

/* Optional phase timing (A.timing != NULL): the control warp passes every barrier, so the time between two barrier
 * exits is the length of a phase of the CTA.  Lane 0 adds the cycles to 16 global counters (scratch/phase_times.py). */
#ifdef AB_HOST_EMUL
#define ABC_TICK(slot_) do { } while (0)
#else
#define ABC_TICK(slot_) do { if (A.timing && threadIdx.x == 32 * ABC_CTRL_WARP) { const long long now_ = clock64(); tacc[slot_] += (unsigned long long)(now_ - tlast); tlast = now_; } } while (0)
#endif

#define ABC_ERR_BUDGET 7      /* index into assist_error_messages: step budget exhausted / dt == 0 */

struct AbcCtl {
    bool have, pending, exhausted, integrating, in_step, have_a0;
    long long sys;
    int nv, ep, win;
    double target, wend;
    PPState P;
    double pc_err, pc_err_last;
    int iterations;
    long long attempts;
};

/* ---- control warp ----------------------------------------------------------------------------- */

/* Bookkeeping of one slot up to the point where a step is due (port of phase A of pp_queue_kernel:
 * same decisions in the same order, so that the step sequence of a system is that of one uninterrupted
 * reb_simulation_integrate / assist_integrate_or_interpolate call).  Returns true when a step is due. */
__device__ bool abc_bookkeeping(const AbcArgs& A, AbcCtl& C, long long slot) {
    const AbBatch& Bt = A.Bt;
    const AbBatch& W = A.W;
    const AbSlices& SL = A.SL;
    const bool dense = (A.times != nullptr);
    const double wsign = (SL.wlen < 0.0) ? -1.0 : 1.0;
    const unsigned long long n_items = (unsigned long long)SL.n_win * (unsigned long long)Bt.n;
    bool step_due = false;
    for (int attempt = 0; attempt < 2 && !step_due; attempt++) {
        while (!C.have && !C.exhausted) {
            if (!C.pending) {
                const unsigned long long q = atomicAdd(A.queue_head, 1ULL);
                if (q >= n_items) { C.exhausted = true; break; }
                C.win = (int)(q / (unsigned long long)Bt.n);
                C.sys = (long long)(q % (unsigned long long)Bt.n);
                if (SL.order) C.sys = SL.order[C.sys];
                C.pending = true;
            }
            if (C.win > 0 && *((volatile int*)(SL.done + C.sys)) < C.win) break;       /* try again after the next trip */
            __threadfence();
            C.pending = false;
            pp_load_cg(Bt, C.sys, C.P);
            C.wend = SL.origin + (double)(C.win + 1) * SL.wlen;
            const bool last = (C.win == SL.n_win - 1);
            const bool beyond = !last && wsign * C.P.t >= wsign * C.wend;
            bool skip = (C.P.status >= 1000);
            if (!skip && !dense) {
                if (C.win == 0) {
                    pp_integrate_entry(C.P, A.tmax);
                    if (beyond) { pp_store(Bt, C.sys, C.P); skip = true; }
                } else if (C.P.status >= 0 || beyond) {
                    skip = true;
                }
                C.integrating = true;
                C.target = A.tmax;
            } else if (!skip) {
                C.ep = (C.win == 0) ? 0 : __ldcg(SL.epoch + C.sys);
                C.integrating = (C.win > 0 && (C.P.status == -1 || C.P.status == -2));
                if (C.integrating) C.target = A.times[C.ep];
                if (beyond || (!C.integrating && C.ep >= A.n_times)) {
                    if (C.win == 0) SL.epoch[C.sys] = 0;
                    skip = true;
                }
            }
            if (skip) {
                if (SL.n_win > 1) { __threadfence(); *((volatile int*)(SL.done + C.sys)) = C.win + 1; }
                continue;
            }
            C.nv = Bt.nv[C.sys];
            W.nv[slot] = C.nv;
            W.status[slot] = 0;
            pp_copy_system<true>(Bt, C.sys, W, slot, C.nv);
            C.have = true;
            C.have_a0 = false;
            C.attempts = 0;
        }
        if (!C.have) break;
        const bool last = (C.win == SL.n_win - 1);
        while (true) {
            if (dense && !C.integrating) {
                while (C.ep < A.n_times) {
                    C.target = A.times[C.ep];
                    const double dts = copysign(1., C.P.dt_last);
                    if (dts * (C.P.t - C.P.dt_last) > dts * C.target || dts * C.target > dts * C.P.t || C.P.dt_last == 0.0) {
                        pp_integrate_entry(C.P, C.target);
                        C.integrating = true;
                        break;
                    }
                    pp_emit(W, slot, C.nv, C.P, C.target, A.out + ((long long)C.ep * Bt.n + C.sys) * Bt.K * 6);
                    C.ep++;
                }
            }
            if (C.integrating && !last && wsign * C.P.t >= wsign * C.wend) break;      /* end of the window: pause integrate() */
            if (C.integrating && ab_check_exit(C.P.t, C.P.dt, C.P.dt_last, C.P.status, C.target, A.exact_finish_time, C.P.last_full_dt) < 0) {
                /* a step that cannot advance the time, or a system over its budget, is retired with an error
                 * instead of holding the CTA forever */
                if (C.P.dt == 0.0 || (A.plan.attempt_budget > 0 && C.attempts >= A.plan.attempt_budget)) {
                    C.P.status = 1;
                    W.status[slot] = 1000 + ABC_ERR_BUDGET;
                    continue;
                }
                step_due = true;
                break;
            }
            if (C.integrating) {           /* integrate() returns */
                if (A.exact_finish_time == 1) C.P.dt = C.P.last_full_dt;
                C.integrating = false;
                if (dense) {
                    pp_emit(W, slot, C.nv, C.P, C.target, A.out + ((long long)C.ep * Bt.n + C.sys) * Bt.K * 6);
                    C.ep++;
                    continue;
                }
            }
            break;                       /* finished */
        }
        if (!step_due) {
            pp_copy_system<false>(W, slot, Bt, C.sys, C.nv);
            if (W.status[slot] >= 1000) Bt.status[C.sys] = W.status[slot];
            pp_store(Bt, C.sys, C.P);
            if (dense) SL.epoch[C.sys] = C.ep;
            if (SL.n_win > 1) { __threadfence(); *((volatile int*)(SL.done + C.sys)) = C.win + 1; }
            C.have = false;
        }
    }
    return step_due;
}

__device__ void abc_control_main(ABC_CTXARG const AbEphem& E, const AbForceOpts& F, const AbcArgs& A, const AbcSmem& sm) {
    AbcCtl ctl[ABC_NL];
    const AbBatch& W = A.W;
#ifndef AB_HOST_EMUL
    unsigned long long tacc[16];
    for (int q = 0; q < 16; q++) tacc[q] = 0ULL;
    long long tlast = clock64();
#endif
    ABC_LANES(l) {
        AbcCtl& C = ctl[ABC_LI(l)];
        C.have = C.pending = C.exhausted = C.integrating = C.in_step = C.have_a0 = false;
        C.sys = -1; C.nv = 0; C.ep = 0; C.win = 0; C.target = A.tmax; C.wend = 0.0;
        C.pc_err = 0.0; C.pc_err_last = 0.0; C.iterations = 0; C.attempts = 0;
    }
    for (;;) {
        int alive = 0;
        ABC_LANES(l) {
            AbcCtl& C = ctl[ABC_LI(l)];
            const long long slot = (long long)ABC_BLOCK * ABC_SLOTS + l;
            bool step_due;
            if (C.in_step) step_due = true;                     /* the last attempt was rejected: try again */
            else step_due = abc_bookkeeping(A, C, slot);
            if (step_due) {
                const int flag = abc_coverage(E, C.P.t, C.P.t + C.P.dt * c_h[7]);
                if (flag != AB_OK) {
                    C.P.status = 1;
                    W.status[slot] = 1000 + flag;
                    C.in_step = false;
                    step_due = false;
                }
            }
            if (step_due && !C.in_step) {
                /* Marsden parameters of the slot's particle */
                const double a1 = A.Bt.has_params ? W.prm[0 * (long long)W.n + slot] : 0.0;
                const double a2 = A.Bt.has_params ? W.prm[1 * (long long)W.n + slot] : 0.0;
                const double a3 = A.Bt.has_params ? W.prm[2 * (long long)W.n + slot] : 0.0;
                sm.prm(0, l) = a1; sm.prm(1, l) = a2; sm.prm(2, l) = a3;
                sm.flag(ABC_SMI_NGON, l) = (F.has_params && !(a1 == 0. && a2 == 0. && a3 == 0.)) ? 1 : 0;
            }
            sm.flag(ABC_SMI_ACTIVE, l) = step_due ? 1 : 0;
            sm.flag(ABC_SMI_NEEDA0, l) = (step_due && !C.have_a0) ? 1 : 0;
            sm.flag(ABC_SMI_ERR, l) = AB_OK;
            sm.flag(ABC_SMI_DEC, l) = 0;
            sm.t0(l) = C.P.t;
            sm.dt(l) = C.P.dt;
            if (C.have || C.pending) alive = 1;
        }
        if (!ABC_SYNC_OR(alive)) {                                               /* B1 */
#ifndef AB_HOST_EMUL
            if (A.timing && threadIdx.x == 32 * ABC_CTRL_WARP) for (int q = 0; q < 16; q++) atomicAdd(A.timing + q, tacc[q]);
#endif
            return;
        }
        ABC_TICK(0);
        abc_fill_warp(ABC_CTXPASS A.gtab + (long long)ABC_BLOCK * ABC_GT_DOUBLES, ABC_CTRL_WARP);
        ABC_SYNC();                                                              /* B2 */
        ABC_TICK(1);
        int anya0 = 0;
        ABC_LANES(l) {
            AbcCtl& C = ctl[ABC_LI(l)];
            const long long slot = (long long)ABC_BLOCK * ABC_SLOTS + l;
            if (sm.flag(ABC_SMI_ACTIVE, l)) {
                const int err = sm.flag(ABC_SMI_ERR, l);
                if (err != AB_OK) {              /* rare kernel layouts only: the pre-check has covered coverage */
                    C.P.status = 1;
                    W.status[slot] = 1000 + err;
                    C.in_step = false;
                    sm.flag(ABC_SMI_ACTIVE, l) = 0;
                    sm.flag(ABC_SMI_NEEDA0, l) = 0;
                }
            }
            const int act = sm.flag(ABC_SMI_ACTIVE, l);
            sm.flag(ABC_SMI_SW, l) = act;
            if (act) {
                C.attempts++;
                if (sm.flag(ABC_SMI_NEEDA0, l)) { C.P.evals++; anya0 = 1; }
                /* first trip through the loop head of the predictor-corrector */
                C.pc_err = 0.0; C.pc_err_last = 1e300; C.iterations = 1;
                C.P.iters++;
            }
        }
        const bool a0_round = ABC_SYNC_OR(anya0);                                /* B4 */
        ABC_TICK(3);
        if (a0_round) {
            ABC_SYNC();                                                          /* B5 */
            ABC_TICK(6);
            ABC_SYNC();                                                          /* B6 */
            ABC_TICK(5);
#ifndef AB_HOST_EMUL
            tacc[11]++;
#endif
        }
#ifndef AB_HOST_EMUL
        tacc[10]++;
#endif
        for (;;) {
            for (int nn = 1; nn < 8; nn++) {
                ABC_SYNC();                                                      /* B7 */
                ABC_TICK(6);
                ABC_SYNC();                                                      /* B8 */
                ABC_TICK(5);
#ifndef AB_HOST_EMUL
                tacc[11]++;
#endif
            }
            ABC_SYNC();                                                          /* B9 */
            ABC_TICK(6);
            int more = 0;
            ABC_LANES(l) {
                AbcCtl& C = ctl[ABC_LI(l)];
                if (sm.flag(ABC_SMI_SW, l)) {
                    double maxak = 0.0, maxb6 = 0.0;
                    for (int c = 0; c < 3; c++) {
                        const double ak = sm.mon(c, l);
                        if (ab_isnormal(ak) && ak > maxak) maxak = ak;
                        const double b6ktmp = sm.mon(3 + c, l);
                        if (ab_isnormal(b6ktmp) && b6ktmp > maxb6) maxb6 = b6ktmp;
                    }
                    C.P.evals += 7;
                    C.pc_err = maxb6 / maxak;
                    bool cont = true;
                    if (C.pc_err < 1e-16) cont = false;
                    else if (C.iterations > 2 && C.pc_err_last <= C.pc_err) cont = false;
                    else if (C.iterations >= 12) cont = false;
                    if (cont) {
                        C.pc_err_last = C.pc_err;
                        C.pc_err = 0;
                        C.iterations++;
                        C.P.iters++;
                        more = 1;
                    }
                    sm.flag(ABC_SMI_SW, l) = cont ? 1 : 0;
                }
            }
            const bool again = ABC_SYNC_OR(more);                                /* B10 */
            ABC_TICK(7);
            if (!again) break;
        }
        /* step-size control (per system: adaptive_mode 1 over the real particle) */
        ABC_LANES(l) {
            AbcCtl& C = ctl[ABC_LI(l)];
            if (sm.flag(ABC_SMI_ACTIVE, l)) {
                PPState& P = C.P;
                const double dt_done = P.dt;
                int dec = 1;
                if (A.Bt.epsilon > 0) {
                    double maxa = 0.0, maxj = 0.0;
                    const double vx = sm.xv(3, l), vy = sm.xv(4, l), vz = sm.xv(5, l);
                    const double xx = sm.xv(0, l), xy = sm.xv(1, l), xz = sm.xv(2, l);
                    const double v2 = vx * vx + vy * vy + vz * vz;
                    const double x2 = xx * xx + xy * xy + xz * xz;
                    if (!(fabs(v2 * P.dt * P.dt / x2) < 1e-16)) {
                        for (int k = 0; k < 3; k++) {
                            const double ak = sm.mon(k, l);
                            if (ab_isnormal(ak) && ak > maxa) maxa = ak;
                            const double b6k = sm.mon(6 + k, l);
                            if (ab_isnormal(b6k) && b6k > maxj) maxj = b6k;
                        }
                    }
                    double dt_new = ab_dt_new(A.Bt.epsilon, A.Bt.min_dt, maxa, maxj, dt_done);
                    if (fabs(dt_new / dt_done) < 0.25) {
                        P.dt = dt_new;
                        if (P.dt_last != 0.) { sm.ratio(l) = P.dt / P.dt_last; dec = 2; }
                        else dec = 3;
                        P.rejected++;
                    } else {
                        if (fabs(dt_new / dt_done) > 1.0) {
                            if (dt_new / dt_done > 1. / 0.25) dt_new = dt_done / 0.25;
                        }
                        P.dt = dt_new;
                    }
                }
                if (dec == 1) {
                    sm.ratio(l) = P.dt / dt_done;
                    P.t += dt_done;
                    P.dt_last = dt_done;
                    P.steps++;
                    C.in_step = false;
                    C.have_a0 = false;
                } else {
                    C.in_step = true;
                    C.have_a0 = true;
                }
                sm.flag(ABC_SMI_DEC, l) = dec;
            }
        }
        ABC_SYNC();                                                              /* B11 */
        ABC_TICK(8);
        ABC_SYNC();                                                              /* B12 */
        ABC_TICK(9);
    }
}

/* ---- component warps (c = 0, 1, 2) --------------------------------------------------------------- */

__device__ void abc_comp_main(ABC_CTXARG const AbEphem& E, const AbForceOpts& F, const AbcArgs& A, const AbcSmem& sm, int c) {
    AbcComp st[ABC_NL];
    const AbBatch& W = A.W;
    const long long wn = W.n;
    for (;;) {
        if (!ABC_SYNC_OR(0)) return;                                             /* B1 */
        abc_fill_warp(ABC_CTXPASS A.gtab + (long long)ABC_BLOCK * ABC_GT_DOUBLES, c);
        ABC_SYNC();                                                              /* B2 */
        const bool a0_round = ABC_SYNC_OR(0);                                    /* B4 */
        ABC_LANES(l) {
            /* every lane loads (an idle slot reads its stale working copy and never uses it): the registers carry
             * nothing from one attempt to the next, so they are free during the fill */
            AbcComp& s = st[ABC_LI(l)];
            abc_comp_load(W, (long long)ABC_BLOCK * ABC_SLOTS + l, c, s);
            if (sm.flag(ABC_SMI_NEEDA0, l)) { sm.xv(c, l) = s.pos; sm.xv(3 + c, l) = s.vel; }
        }
        if (a0_round) {
            ABC_SYNC();                                                          /* B5 */
            ABC_SYNC();                                                          /* B6 */
            ABC_LANES(l) {
                if (sm.flag(ABC_SMI_NEEDA0, l)) {
                    AbcComp& s = st[ABC_LI(l)];
                    const long long ws = (long long)ABC_BLOCK * ABC_SLOTS + l;
                    const double a = abc_sum_forces(E, F, sm, c, l);
                    s.acc = a;
                    ABC_W1(W.acc, c) = a;
                    ABC_W1(W.ls_pos, c) = s.pos;
                    ABC_W1(W.ls_vel, c) = s.vel;
                    ABC_W1(W.ls_acc, c) = a;
                }
            }
        }
        ABC_LANES(l) {
            if (sm.flag(ABC_SMI_ACTIVE, l)) {
                AbcComp& s = st[ABC_LI(l)];
                abc_attempt_begin(s);
                double xk, vk;
                abc_predict(s, 1, sm.dt(l), xk, vk);
                sm.xv(c, l) = xk; sm.xv(3 + c, l) = vk;
            }
        }
        for (;;) {
            for (int nn = 1; nn < 8; nn++) {
                ABC_SYNC();                                                      /* B7 */
                if (nn < 6) {
                    /* while the workers evaluate node nn: the part of the prediction for node nn + 1 that the
                     * coming update cannot change (it touches b_0 .. b_{nn-1}) */
                    ABC_LANES(l) {
                        if (sm.flag(ABC_SMI_SW, l)) {
                            AbcComp& s = st[ABC_LI(l)];
                            s.hx = 0.0; s.hv = 0.0;
                            abc_predict_stages(s, c_h[nn + 1], 6, nn + 1, s.hx, s.hv);
                        }
                    }
                }
                ABC_SYNC();                                                      /* B8 */
                ABC_LANES(l) {
                    if (sm.flag(ABC_SMI_SW, l)) {
                        AbcComp& s = st[ABC_LI(l)];
                        const double at = abc_sum_forces(E, F, sm, c, l);
                        s.at = at;
                        const double db6 = abc_update_gb(s, nn, at);
                        if (nn < 7) {
                            double xk, vk;
                            const double h = c_h[nn + 1];
                            double px = s.hx, pv = s.hv;
                            abc_predict_stages(s, h, (nn < 6) ? nn : 6, 0, px, pv);
                            abc_predict_final(s, h, sm.dt(l), px, pv, xk, vk);
                            sm.xv(c, l) = xk; sm.xv(3 + c, l) = vk;
                        } else {
                            sm.mon(c, l) = fabs(at);
                            sm.mon(3 + c, l) = db6;
                            sm.mon(6 + c, l) = fabs(s.b[6]);
                        }
                    }
                }
            }
            ABC_SYNC();                                                          /* B9 */
            if (!ABC_SYNC_OR(0)) break;                                          /* B10 */
            ABC_LANES(l) {
                if (sm.flag(ABC_SMI_SW, l)) {
                    AbcComp& s = st[ABC_LI(l)];
                    double xk, vk;
                    abc_predict(s, 1, sm.dt(l), xk, vk);
                    sm.xv(c, l) = xk; sm.xv(3 + c, l) = vk;
                }
            }
        }
        ABC_SYNC();                                                              /* B11 */
        ABC_LANES(l) {
            if (sm.flag(ABC_SMI_ACTIVE, l)) {
                AbcComp& s = st[ABC_LI(l)];
                const long long ws = (long long)ABC_BLOCK * ABC_SLOTS + l;
                const int dec = sm.flag(ABC_SMI_DEC, l);
                if (dec == 1) {
                    abc_advance(s, sm.dt(l));
#pragma unroll
                    for (int j = 0; j < 7; j++) { ABC_W7(W.er, j, c) = s.e[j]; ABC_W7(W.br, j, c) = s.b[j]; }
                    abc_predict_next(s, sm.ratio(l), s.e[0], s.e[1], s.e[2], s.e[3], s.e[4], s.e[5], s.e[6],
                                     s.b[0], s.b[1], s.b[2], s.b[3], s.b[4], s.b[5], s.b[6]);
                } else {
                    s.pos = s.x0; s.vel = s.v0; s.acc = s.a0;
                    if (dec == 2) {
                        abc_predict_next(s, sm.ratio(l),
                                         ABC_W7(W.er, 0, c), ABC_W7(W.er, 1, c), ABC_W7(W.er, 2, c), ABC_W7(W.er, 3, c),
                                         ABC_W7(W.er, 4, c), ABC_W7(W.er, 5, c), ABC_W7(W.er, 6, c),
                                         ABC_W7(W.br, 0, c), ABC_W7(W.br, 1, c), ABC_W7(W.br, 2, c), ABC_W7(W.br, 3, c),
                                         ABC_W7(W.br, 4, c), ABC_W7(W.br, 5, c), ABC_W7(W.br, 6, c));
                    }
                }
                abc_comp_store(W, ws, c, s);
            }
        }
        ABC_SYNC();                                                              /* B12 */
    }
}

/* ---- worker warps --------------------------------------------------------------------------------- */

/* One copy in the kernel (the a0 round and the node rounds call it): see abc_task_all_bodies. */
__device__ __noinline__ void abc_worker_tasks(ABC_CTXARG const double* gt, int widx, int node, int which_flag) {
    const AbEphem& E = c_abcE;
    const AbForceOpts& F = c_abcF;
    ABC_SM_HERE(sm);
    ABC_LANES(l) {
        if (sm.flag(which_flag, l)) {
            const AbcWorkerPlan& wp = c_abcP.w[widx];
            const double* tb = sm.tab(node, l);
            const double* g = gt + node * ABC_GT_NODE + l;
            /* requests to the global table first: they travel while the planets are worked on */
            double ac[ABC_BN][3];
            if (wp.nast) abc_ast_fetch(wp, g, 0, ac);
            double ev[7];
            const bool eihsrc = (wp.scalar[0] == ABC_T_EIHSRC || wp.scalar[1] == ABC_T_EIHSRC);
            if (eihsrc) {
#pragma unroll
                for (int q = 0; q < 7; q++) ev[q] = __ldcg(g + (ABC_GT_SVEL(0) + q) * ABC_SLOTS);
            }
            if (wp.scalar[0] != ABC_T_NONE) abc_run_task(E, F, sm, tb, g, ev, l, wp.scalar[0]);
            if (wp.scalar[1] != ABC_T_NONE) abc_run_task(E, F, sm, tb, g, ev, l, wp.scalar[1]);
            if (wp.nbody + wp.nast) abc_task_all_bodies(E, F, sm, tb, g, wp, l, ac);
        }
    }
}

__device__ void abc_worker_main(ABC_CTXARG const AbEphem& E, const AbForceOpts& F, const AbcArgs& A, const AbcSmem& sm, int warp) {
    const int widx = warp - ABC_FIRST_WORKER;
    double* gt = A.gtab + (long long)ABC_BLOCK * ABC_GT_DOUBLES;
    for (;;) {
        if (!ABC_SYNC_OR(0)) return;                                             /* B1 */
        abc_fill_warp(ABC_CTXPASS gt, warp);
        ABC_SYNC();                                                              /* B2 */
        const bool a0_round = ABC_SYNC_OR(0);                                    /* B4 */
        if (a0_round) {
            ABC_SYNC();                                                          /* B5 */
            abc_worker_tasks(ABC_CTXPASS gt, widx, 0, ABC_SMI_NEEDA0);
            ABC_SYNC();                                                          /* B6 */
        }
        for (;;) {
            for (int nn = 1; nn < 8; nn++) {
                ABC_SYNC();                                                      /* B7 */
                abc_worker_tasks(ABC_CTXPASS gt, widx, nn, ABC_SMI_SW);
                ABC_SYNC();                                                      /* B8 */
            }
            ABC_SYNC();                                                          /* B9 */
            if (!ABC_SYNC_OR(0)) break;                                          /* B10 */
        }
        ABC_SYNC();                                                              /* B11 */
        ABC_SYNC();                                                              /* B12 */
    }
}

}  // namespace AB_NS
#endif
