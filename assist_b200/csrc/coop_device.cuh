/*
 * coop_device.cuh -- per-particle-dt IAS15 with a whole CTA stepping 32 systems together.
 *
 * Why: one thread per system (kernels.cu, pp_queue_kernel) keeps the 8 body tables of a step (6 KB) and the
 * IAS15 tables in thread-local / global memory and walks one dependent chain per system.  Here a system's
 * working set lives ON CHIP -- the body tables of its 8 node times in shared memory, its IAS15 state in the
 * registers of three "component" warps -- and the independent work of a step is spread over the CTA:
 *
 *     lane  = system slot (32 systems per CTA, structure-of-arrays in shared memory with stride 32:
 *             every access of a warp is 32 consecutive doubles, conflict free)
 *     warp  = task
 *         warps 0-2    component x / y / z of every system: predictor, ordered force sums, g/b update,
 *                      advance, predict_next  (b, g, e, csb, x0, v0, a0, cs* in registers)
 *         warp  3      control: queue, reb_simulation_integrate bookkeeping, output epochs, convergence and
 *                      step-size control (sqrt7)
 *         warps 4-15   workers: one body of the direct term (+ its share of the EIH sums) or one of the
 *                      single-body terms (Earth J2-J4, solar J2, EIH source block, Marsden, GR variants) per task
 *         all 16       the Chebyshev fill of the node tables: thread = (slot, node), so the eight nodes of a
 *                      system read the same one or two coefficient records (broadcast loads instead of 32 lanes
 *                      gathering 32 records)
 *
 * Values are those of the one-thread-per-system path, bit for bit in the strict build: every term is formed
 * by the same operations, and the sums that the reference accumulates in a fixed order (27 direct terms,
 * 11 EIH potential terms, the term sequence of the dispatcher) are added in that order by the component warps
 * from the per-body values left in shared memory.
 *   reference src/forces.c:49-173 (dispatcher order), :266-344 (direct), :1288-1501 (EIH, real particle),
 *   src/spk.c:405-547, src/ascii_ephem.c:27-65, 275-384 (Chebyshev), REBOUND IAS15 (see ias15_device.cuh).
 *
 * Scope: systems without variational particles, one EIH source, barycentric -- the configuration of the
 * node-table fast path.  Everything else runs on pp_queue_kernel.
 *
 * The code is written as warp-level role functions made of per-lane blocks (ABC_LANES) separated by CTA
 * barriers (ABC_SYNC).  Built with AB_HOST_EMUL the same source runs on the host, one OS thread per warp
 * (tests/emul): that is how the control flow was brought up without a GPU.
 */
#ifndef AB_COOP_DEVICE_CUH
#define AB_COOP_DEVICE_CUH

#include "device_types.h"
#include "ephem_device.cuh"
#include "forces_device.cuh"
#include "ias15_device.cuh"

/* node table: entries of one (node, slot) */
#define ABC_TAB_E 88
#define ABC_NODE_STRIDE (ABC_TAB_E * ABC_SLOTS + 4)    /* +4: the 8 nodes of a slot fall into different banks (fill stores) */
#define ABC_E_POS(i, c) ((i) * 3 + (c))
#define ABC_E_SVEL(c) (81 + (c))
#define ABC_E_TERM1 84
#define ABC_E_AR(c) (85 + (c))

/* contribution slots */
#define ABC_C_NG 0
#define ABC_C_EARTH 3
#define ABC_C_SUNJ2 6
#define ABC_C_M 9          /* (-prefacij) * d_ij */
#define ABC_C_T 12         /* term1 .. term6 */
#define ABC_C_E78 18       /* term7_sum * over_C2 + term8_sum * over_C2 */
#define ABC_C_GRPOT 21
#define ABC_C_GRSIMPLE 24
#define ABC_NCON 27

/* shared-memory carve-up (doubles first, then ints).  XV .. EXTRA are free while the node tables are being filled
 * and double as the staging area of the coefficient records (ABC_SM_STAGE); the total is the 227 KB a CTA can have. */
#define ABC_SM_TAB 0
#define ABC_SM_XV (ABC_SM_TAB + 8 * ABC_NODE_STRIDE)
#define ABC_SM_PROD (ABC_SM_XV + 6 * ABC_SLOTS)
#define ABC_SM_Q (ABC_SM_PROD + 81 * ABC_SLOTS)
#define ABC_SM_CON (ABC_SM_Q + 11 * ABC_SLOTS)
#define ABC_SM_MON (ABC_SM_CON + ABC_NCON * ABC_SLOTS)       /* |a| x3, |db6| x3, |b6| x3 */
#define ABC_SM_EXTRA (ABC_SM_MON + 9 * ABC_SLOTS)            /* staging only */
#define ABC_SM_PRM (ABC_SM_EXTRA + 1920)
#define ABC_SM_T0 (ABC_SM_PRM + 3 * ABC_SLOTS)               /* start time of the attempt */
#define ABC_SM_DT (ABC_SM_T0 + ABC_SLOTS)                    /* step of the attempt */
#define ABC_SM_RATIO (ABC_SM_DT + ABC_SLOTS)                 /* predict_next ratio */
#define ABC_SM_DOUBLES (ABC_SM_RATIO + ABC_SLOTS)
#define ABC_SM_STAGE ABC_SM_XV
#define ABC_STAGE_DOUBLES (ABC_SM_PRM - ABC_SM_XV)           /* 6208: 32 slots x (cap_p + cap_a), cap_p + cap_a <= 194 */
#define ABC_SMI_ACTIVE 0      /* slot takes part in this attempt */
#define ABC_SMI_NEEDA0 1      /* slot needs the force evaluation at the start of the step */
#define ABC_SMI_SW 2          /* slot is still sweeping */
#define ABC_SMI_DEC 3         /* 0 nothing, 1 accepted, 2 rejected, 3 rejected and no previous step */
#define ABC_SMI_NGON 4        /* Marsden term active for this slot */
#define ABC_SMI_ERR 5         /* ephemeris status of the fill */
#define ABC_SM_INTS (6 * ABC_SLOTS)
#define ABC_SMEM_BYTES ((size_t)ABC_SM_DOUBLES * 8 + (size_t)ABC_SM_INTS * 4)

#ifdef AB_HOST_EMUL
#define ABC_NL 32
#define ABC_LANES(l) for (int l = 0; l < 32; ++l)
#define ABC_LI(l) (l)
#define ABC_SYNC() abc_emul_sync(ctx)
#define ABC_SYNC_OR(p) abc_emul_sync_or(ctx, (p))
#define ABC_CTXARG AbcEmulCtx *ctx,
#define ABC_CTXPASS ctx,
#define ABC_BLOCK (ctx->block)
#else
#define ABC_NL 1
#define ABC_LANES(l) for (int l = (int)(threadIdx.x & 31), once_ = 1; once_; once_ = 0)
#define ABC_LI(l) 0
#define ABC_SYNC() __syncthreads()
#define ABC_SYNC_OR(p) (__syncthreads_or(p) != 0)
#define ABC_CTXARG
#define ABC_CTXPASS
#define ABC_BLOCK ((int)blockIdx.x)
#endif

namespace AB_NS {

/* Launch-time copies in CONSTANT memory for the out-of-line fill routine: a kernel parameter reached through a
 * reference is read with generic loads (hundreds of cycles each, one after the other in the record look-up); these
 * are read through the constant cache.  The asteroid descriptors, which AbEphem keeps in global memory, come along. */
static __constant__ AbEphem c_abcE;
static __constant__ AbForceOpts c_abcF;
static __constant__ AbSpkTarget c_abc_ast[AB_MAX_AST];
#ifndef AB_HOST_EMUL
extern __shared__ double abc_shared[];
#endif

/* What the force routines of forces_device.cuh see as "body table": one node of one slot in shared memory. */
struct AbcRow {
    const double* p;
    __device__ __forceinline__ double operator[](int c) const { return p[c * ABC_SLOTS]; }
};
struct AbcRows {
    const double* base;
    int stride;       /* doubles between consecutive rows */
    __device__ __forceinline__ AbcRow operator[](int i) const { return AbcRow{base + i * stride}; }
};
struct AbcScalars {
    const double* base;
    __device__ __forceinline__ double operator[](int) const { return *base; }
};
struct AbcTabView {
    const double* gm;
    AbcRows pos, vel, eih_ar, eih_av;
    AbcScalars eih_term1;
    double earth_acc[3];
    __device__ __forceinline__ AbcTabView(const double* gm_, const double* node_slot) : gm(gm_) {
        pos.base = node_slot; pos.stride = 3 * ABC_SLOTS;
        vel.base = node_slot + ABC_E_SVEL(0) * ABC_SLOTS; vel.stride = 0;
        eih_ar.base = node_slot + ABC_E_AR(0) * ABC_SLOTS; eih_ar.stride = 0;
        eih_av = eih_ar;
        eih_term1.base = node_slot + ABC_E_TERM1 * ABC_SLOTS;
        earth_acc[0] = earth_acc[1] = earth_acc[2] = 0.0;
    }
};

struct AbcSmem {
    double* d;
    int* i;
    __device__ __forceinline__ double* tab(int node, int slot) const { return d + ABC_SM_TAB + node * ABC_NODE_STRIDE + slot; }
    __device__ __forceinline__ double& xv(int e, int slot) const { return d[ABC_SM_XV + e * ABC_SLOTS + slot]; }
    __device__ __forceinline__ double& prod(int body, int c, int slot) const { return d[ABC_SM_PROD + (body * 3 + c) * ABC_SLOTS + slot]; }
    __device__ __forceinline__ double& q(int k, int slot) const { return d[ABC_SM_Q + k * ABC_SLOTS + slot]; }
    __device__ __forceinline__ double& con(int e, int slot) const { return d[ABC_SM_CON + e * ABC_SLOTS + slot]; }
    __device__ __forceinline__ double& mon(int e, int slot) const { return d[ABC_SM_MON + e * ABC_SLOTS + slot]; }
    __device__ __forceinline__ double& prm(int e, int slot) const { return d[ABC_SM_PRM + e * ABC_SLOTS + slot]; }
    __device__ __forceinline__ double& t0(int slot) const { return d[ABC_SM_T0 + slot]; }
    __device__ __forceinline__ double& dt(int slot) const { return d[ABC_SM_DT + slot]; }
    __device__ __forceinline__ double& ratio(int slot) const { return d[ABC_SM_RATIO + slot]; }
    __device__ __forceinline__ int& flag(int which, int slot) const { return i[which * ABC_SLOTS + slot]; }
};

/* everything a role function needs, by value */
struct AbcArgs {
    AbBatch Bt, W;
    double tmax;
    int exact_finish_time;
    unsigned long long* queue_head;
    AbSlices SL;
    const double* times;
    int n_times;
    double* out;
    AbcPlan plan;
    unsigned long long* timing;      /* optional: 16 cycle counters of the phases (coop_roles.cuh, ABC_TICK) */
};

/* ------------------------------------------------------------------------------------------ */
/* node-table fill                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* Coverage and presence checks of a step, as ab_fill_nodes makes them (the two end nodes decide). */
__device__ int abc_coverage(const AbEphem& E, double t_first, double t_last) {
    const double jd_ref = E.jd_ref;
    if (E.cov_simple) {     /* the common window, prepared on the host: same comparisons, folded */
        const double j0 = jd_ref + t_first, j1 = jd_ref + t_last;
        if (j0 >= E.cov_lo && j0 <= E.cov_hi && j1 >= E.cov_lo && j1 <= E.cov_hi) return AB_OK;
    }
    for (int k = 0; k < 2; k++) {
        const double jd = jd_ref + (k == 0 ? t_first : t_last);
        if (E.planets_source == AB_SRC_ASCII) {
            if (jd < E.a_beg || jd > E.a_end) return AB_ERR_COVERAGE;
        } else {
            for (int b = 0; b < AB_NPLANETS; b++) {
                const int idx = E.p_index[b];
                if (idx < 0) { if (b != 3) return AB_ERR_NEPHEM; continue; }
                if (jd < E.p_tgt[idx].beg || jd > E.p_tgt[idx].end) return AB_ERR_COVERAGE;
            }
        }
        for (int m = 0; m < E.n_ast; m++)
            if (jd < E.a_tgt[m].beg || jd > E.a_tgt[m].end) return AB_ERR_COVERAGE;
    }
    return AB_OK;
}

/* ---- the fill: thread = (slot, node); warps 0-7 take the planets, a few asteroids (plan.ast_split) and then the
 * particle-independent EIH sums of their node, warps 8-15 the other asteroids of the same slots.
 * Lane = 8 * (slot within the warp's four) + node.
 *
 * The kernel is bound by the latency of each warp's own dependent chain (16 warps per SM, FP64 results 8 cycles
 * apart), so a lane evaluates FOUR series side by side: four record look-ups, four MID loads and four Chebyshev
 * recurrences in flight at once, 16-byte coefficient loads from the packed image (the eight node-lanes of a slot
 * read the same one or two records: the loads of a warp touch a handful of lines).  Arithmetic per (series, time) is
 * that of ephem_device.cuh: same sums, same order.
 * (A variant that copied the records of a slot cooperatively into shared memory first measured slower: the copy and
 * its bookkeeping cost more instructions than the gather it avoided.) */
struct AbcSeriesRef {
    const double* img;
    const AbSpkTarget* tg;
    int kind;              /* 0 EMB, 2 planet `idx`, 3 asteroid `idx`, -1 none */
    int idx;
};

/* series s of a warp's list: EMB, Mercury .. Pluto (the Sun is evaluated on its own, with velocity), then asteroids */
__device__ __forceinline__ AbcSeriesRef abc_series_ref(int ast_split, bool planets_half, int s, int s_end) {
    const AbEphem& E = c_abcE;
    AbcSeriesRef r;
    if (s >= s_end) { r.img = E.spka_img; r.tg = &c_abc_ast[0]; r.kind = -1; r.idx = 0; return r; }
    if (planets_half && s < AB_NPLANETS) {
        r.img = E.spkp_img;
        if (s == 0) { r.kind = 0; r.idx = -1; r.tg = &E.p_tgt[E.emb_index]; }
        else { r.idx = s; r.kind = 2; r.tg = &E.p_tgt[E.p_index[s]]; }
    } else {
        const int m = planets_half ? (s - AB_NPLANETS) : (ast_split + s);
        r.img = E.spka_img; r.kind = 3; r.idx = m; r.tg = &c_abc_ast[m];
    }
    return r;
}

/* Position sums (file units) of four series at time t, the four recurrences side by side. */
__device__ __forceinline__ void abc_quad_eval(const AbcSeriesRef* R, double jd_ref, double t, double (*u)[3]) {
    const double2* q[4];
    int P[4];
    double z[4], T1[4], T2[4], a0[4], a1[4], a2[4];
    int Pmax = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if (R[s].kind < 0) { P[s] = 0; z[s] = 0.0; q[s] = nullptr; continue; }      /* past the end of the list */
        const AbSpkTarget& tg = *R[s].tg;
        const AbSpkSeg& sg = tg.seg[ab_spk_segment(tg, jd_ref, t)];
        double c;
        const double* cf = ab_spk_record_in(R[s].img, sg, jd_ref, t, &z[s], &c);
        q[s] = reinterpret_cast<const double2*>(cf);
        P[s] = sg.P;
        if (P[s] > Pmax) Pmax = P[s];
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        /* p = 0 (T = 1) and p = 1 (T = z) */
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (P[s] > 0) {
            const double2 A = __ldg(q[s]), B = __ldg(q[s] + 1), C = __ldg(q[s] + 2);      /* x0 y0 | z0 x1 | y1 z1 */
            s0 += A.x * 1.0; s1 += A.y * 1.0; s2 += B.x * 1.0;
            s0 += B.y * z[s]; s1 += C.x * z[s]; s2 += C.y * z[s];
        }
        a0[s] = s0; a1[s] = s1; a2[s] = s2;
        T2[s] = 1.0; T1[s] = z[s];
    }
#pragma unroll 1
    for (int p = 2; p < Pmax; p += 2) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (p < P[s]) {
                const double2* qq = q[s] + 3 * (p >> 1);
                const double2 A = __ldg(qq), B = __ldg(qq + 1);                          /* xp yp | zp xp+1 */
                const double Ta = 2.0 * z[s] * T1[s] - T2[s];
                a0[s] += A.x * Ta; a1[s] += A.y * Ta; a2[s] += B.x * Ta;
                if (p + 1 < P[s]) {      /* an odd number of terms: the second half of the last pair is padding */
                    const double2 C = __ldg(qq + 2);                                      /* yp+1 zp+1 */
                    const double Tb = 2.0 * z[s] * Ta - T1[s];
                    a0[s] += B.y * Tb; a1[s] += C.x * Tb; a2[s] += C.y * Tb;
                    T2[s] = Ta; T1[s] = Tb;
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; s++) { u[s][0] = a0[s]; u[s][1] = a1[s]; u[s][2] = a2[s]; }
}

/* what becomes of the sums of one series: EMB kept, planets and asteroids into the table */
__device__ __forceinline__ void abc_fill_store(const AbcSeriesRef& R, double* u, double* emb, double* tb) {
    const AbEphem& E = c_abcE;
    if (R.kind == 0) {
        emb[0] = u[0]; emb[1] = u[1]; emb[2] = u[2];
    } else if (R.kind == 3) {
        /* heliocentric position / 149597870.7 (reference src/spk.c:470); the Sun is added after the barrier */
        const int b = AB_NPLANETS + R.idx;
        tb[ABC_E_POS(b, 0) * ABC_SLOTS] = AB_DIVK(u[0], 149597870.7);
        tb[ABC_E_POS(b, 1) * ABC_SLOTS] = AB_DIVK(u[1], 149597870.7);
        tb[ABC_E_POS(b, 2) * ABC_SLOTS] = AB_DIVK(u[2], 149597870.7);
    } else if (R.kind == 2) {
        const int b = R.idx;
        if (b == 3 || b == 4) { u[0] += emb[0]; u[1] += emb[1]; u[2] += emb[2]; }    /* relative to the EMB (reference src/spk.c:572-587) */
        tb[ABC_E_POS(b, 0) * ABC_SLOTS] = ab_divc(u[0], E.u_d[0], E.u_rd[0]);
        tb[ABC_E_POS(b, 1) * ABC_SLOTS] = ab_divc(u[1], E.u_d[0], E.u_rd[0]);
        tb[ABC_E_POS(b, 2) * ABC_SLOTS] = ab_divc(u[2], E.u_d[0], E.u_rd[0]);
    }
}

#ifdef AB_HOST_EMUL
#define ABC_SM_HERE(sm) const AbcSmem& sm = *ctx->sm
#else
#define ABC_SM_HERE(sm) AbcSmem sm; sm.d = abc_shared; sm.i = reinterpret_cast<int*>(abc_shared + ABC_SM_DOUBLES)
#endif

struct AbcFillLane {
    double t;
    int active;
    double emb[3];
};

__device__ __noinline__ void abc_fill_warp(ABC_CTXARG int ast_split, int cap_p, int cap_a, int warp) {
    const AbEphem& E = c_abcE;
    const AbForceOpts& F = c_abcF;
    ABC_SM_HERE(sm);
    (void)cap_p; (void)cap_a;
    AbcFillLane L[ABC_NL];
    const double jd_ref = E.jd_ref;
    const bool planets_half = (warp < 8);
    ABC_LANES(l) {
        AbcFillLane& q = L[ABC_LI(l)];
        const int slot = 4 * (warp & 7) + (l >> 3);
        const int node = l & 7;
        q.active = sm.flag(ABC_SMI_ACTIVE, slot);
        const double t0 = sm.t0(slot);
        q.t = (node == 0) ? t0 : (t0 + sm.dt(slot) * c_h[node]);
        q.emb[0] = q.emb[1] = q.emb[2] = 0.0;
    }
    /* usual SPK layout: every body and the EMB have a target */
    bool spk_regular = (E.planets_source != AB_SRC_ASCII) && E.emb_index >= 0;
    for (int b = 0; b < AB_NPLANETS && spk_regular; b++) if (E.p_index[b] < 0) spk_regular = false;
    int s_first = 0;
    if (planets_half) {
        ABC_LANES(l) {
            AbcFillLane& q = L[ABC_LI(l)];
            if (q.active) {
                const int slot = 4 * (warp & 7) + (l >> 3);
                double* tb = sm.tab(l & 7, slot);
                int err = AB_OK;
                /* the Sun with its velocity; every planet when the layout is unusual (DE-binary planets, a kernel
                 * without Earth or EMB target): the one-time routines */
                for (int b = 0; b < (spk_regular ? 1 : AB_NPLANETS); b++) {
                    double GM, x[3], v[3], a[3];
                    int flag;
                    if (b == 0) {
                        flag = ab_planet<1>(E, 0, q.t, &GM, x, v, a);
                        tb[ABC_E_SVEL(0) * ABC_SLOTS] = v[0]; tb[ABC_E_SVEL(1) * ABC_SLOTS] = v[1]; tb[ABC_E_SVEL(2) * ABC_SLOTS] = v[2];
                    } else {
                        flag = ab_planet<0>(E, b, q.t, &GM, x, v, a);
                    }
                    if (flag != AB_OK && err == AB_OK) err = flag;
                    tb[ABC_E_POS(b, 0) * ABC_SLOTS] = x[0]; tb[ABC_E_POS(b, 1) * ABC_SLOTS] = x[1]; tb[ABC_E_POS(b, 2) * ABC_SLOTS] = x[2];
                }
                if (err != AB_OK) sm.flag(ABC_SMI_ERR, slot) = err;      /* the lanes of a slot may race: every value written is a valid code */
            }
        }
        if (!spk_regular) s_first = AB_NPLANETS;
    }
    const int s_end = planets_half ? (AB_NPLANETS + ast_split) : (E.n_ast - ast_split);
#pragma unroll 1
    for (int s = s_first; s < s_end; s += 4) {
        AbcSeriesRef R[4];
#pragma unroll
        for (int j = 0; j < 4; j++) R[j] = abc_series_ref(ast_split, planets_half, s + j, s_end);
        ABC_LANES(l) {
            AbcFillLane& q = L[ABC_LI(l)];
            if (q.active) {
                double u[4][3];
                abc_quad_eval(R, jd_ref, q.t, u);
                double* tb = sm.tab(l & 7, 4 * (warp & 7) + (l >> 3));
#pragma unroll
                for (int j = 0; j < 4; j++) abc_fill_store(R[j], u[j], q.emb, tb);
            }
        }
    }
    /* particle-independent EIH sums of the Sun at this node (ab_fill_nodes, same operations): the lane reads back
     * the eleven positions it has just written */
    if (planets_half && (F.forces & 0x40)) {
        ABC_LANES(l) {
            const AbcFillLane& q = L[ABC_LI(l)];
            if (q.active) {
                double* tb = sm.tab(l & 7, 4 * (warp & 7) + (l >> 3));
                const double sx = tb[ABC_E_POS(0, 0) * ABC_SLOTS], sy = tb[ABC_E_POS(0, 1) * ABC_SLOTS], sz = tb[ABC_E_POS(0, 2) * ABC_SLOTS];
                double term1 = 0.0, arx = 0.0, ary = 0.0, arz = 0.0;
#pragma unroll 1
                for (int k0 = 1; k0 < AB_NPLANETS; k0 += 5) {
                    /* five terms at a time: the square roots and divisions of a group are independent and overlap,
                     * then the group is added in order */
                    double t1[5], fx[5], fy[5], fz[5];
#pragma unroll
                    for (int j = 0; j < 5; j++) {
                        const int k = k0 + j;
                        const double GMk = E.gm[k];
                        const double dxjk = sx - tb[ABC_E_POS(k, 0) * ABC_SLOTS];
                        const double dyjk = sy - tb[ABC_E_POS(k, 1) * ABC_SLOTS];
                        const double dzjk = sz - tb[ABC_E_POS(k, 2) * ABC_SLOTS];
                        const double rjk2 = dxjk * dxjk + dyjk * dyjk + dzjk * dzjk;
                        const double _rjk = sqrt(rjk2);
                        t1[j] = GMk / _rjk;
                        const double fac = GMk / (rjk2 * _rjk);
                        fx[j] = fac * dxjk; fy[j] = fac * dyjk; fz[j] = fac * dzjk;
                    }
#pragma unroll
                    for (int j = 0; j < 5; j++) { term1 += t1[j]; arx -= fx[j]; ary -= fy[j]; arz -= fz[j]; }
                }
                tb[ABC_E_TERM1 * ABC_SLOTS] = term1;
                tb[ABC_E_AR(0) * ABC_SLOTS] = arx; tb[ABC_E_AR(1) * ABC_SLOTS] = ary; tb[ABC_E_AR(2) * ABC_SLOTS] = arz;
            }
        }
    }
}

/* after the barrier: asteroids heliocentric -> barycentric (reference src/forces.c:213-219); thread = (slot, node, half) */
__device__ void abc_fill_shift(const AbEphem& E, const AbcSmem& sm, int warp, int lane) {
    const int slot = 4 * (warp & 7) + (lane >> 3);
    const int node = lane & 7;
    if (!sm.flag(ABC_SMI_ACTIVE, slot)) return;
    double* tb = sm.tab(node, slot);
    const double sx = tb[ABC_E_POS(0, 0) * ABC_SLOTS], sy = tb[ABC_E_POS(0, 1) * ABC_SLOTS], sz = tb[ABC_E_POS(0, 2) * ABC_SLOTS];
    const int half = (E.n_ast + 1) / 2;
    const int m0 = (warp < 8) ? 0 : half, m1 = (warp < 8) ? half : E.n_ast;
    for (int m = m0; m < m1; m++) {
        const int b = AB_NPLANETS + m;
        tb[ABC_E_POS(b, 0) * ABC_SLOTS] = tb[ABC_E_POS(b, 0) * ABC_SLOTS] + sx;
        tb[ABC_E_POS(b, 1) * ABC_SLOTS] = tb[ABC_E_POS(b, 1) * ABC_SLOTS] + sy;
        tb[ABC_E_POS(b, 2) * ABC_SLOTS] = tb[ABC_E_POS(b, 2) * ABC_SLOTS] + sz;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* worker tasks (lane = slot)                                                                 */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ bool abc_body_on(int i, int fmask) {
    if (i == 0) return (fmask & 0x01) != 0;
    if (i < AB_NPLANETS) return (fmask & 0x02) != 0;
    return (fmask & 0x04) != 0;
}

/* N bodies of the direct term (reference src/forces.c:325-344) and, for planets, their terms of the EIH potential
 * sum (src/forces.c:1400-1416: same separation, same square root).  The bodies are independent until the component
 * warps add them up, so their chains (difference, square root, division) are laid side by side. */
template <int N, bool PLANETS>
__device__ __forceinline__ void abc_task_bodies(const AbEphem& E, const AbForceOpts& F, const AbcSmem& sm, const double* tb,
                                                const unsigned char* ids, int slot) {
    const double px = sm.xv(0, slot), py = sm.xv(1, slot), pz = sm.xv(2, slot);
    const double xo = 0.0, yo = 0.0, zo = 0.0;
    double dx[N], dy[N], dz[N], _r[N], GM[N];
#pragma unroll
    for (int n = 0; n < N; n++) {
        const int i = ids[n];
        const double cx = tb[ABC_E_POS(i, 0) * ABC_SLOTS], cy = tb[ABC_E_POS(i, 1) * ABC_SLOTS], cz = tb[ABC_E_POS(i, 2) * ABC_SLOTS];
        GM[n] = E.gm[i];
        dx[n] = px + (xo - cx);
        dy[n] = py + (yo - cy);
        dz[n] = pz + (zo - cz);
        const double r2 = dx[n] * dx[n] + dy[n] * dy[n] + dz[n] * dz[n];
        _r[n] = sqrt(r2);
    }
    const bool eih = PLANETS && (F.forces & 0x40);
#pragma unroll
    for (int n = 0; n < N; n++) {
        const int i = ids[n];
        const double prefac = GM[n] / (_r[n] * _r[n] * _r[n]);
        const double p0 = prefac * dx[n], p1 = prefac * dy[n], p2 = prefac * dz[n];
        if (abc_body_on(i, F.forces)) { sm.prod(i, 0, slot) = p0; sm.prod(i, 1, slot) = p1; sm.prod(i, 2, slot) = p2; }
        if (PLANETS && i < AB_NPLANETS && eih) sm.q(i, slot) = GM[n] / _r[n];
    }
}

/* The group of a worker warp, one body after the other through the same few hundred bytes of code: nine warps walk
 * through it at the same time, so it is fetched once per SM and not once per warp (the kernel's hot code has to fit
 * the instruction cache: unrolled per-warp variants of this loop made the force phase twice as long). */
__device__ __forceinline__ void abc_task_group(const AbEphem& E, const AbForceOpts& F, const AbcSmem& sm, const double* tb,
                                            const AbcWorkerPlan& wp, int slot) {
#pragma unroll 1
    for (int n = 0; n < wp.nbody; n++) abc_task_bodies<1, true>(E, F, sm, tb, wp.body + n, slot);
}

/* EIH source block of the Sun for the real particle (reference src/forces.c:1319-1501 with j = 0), everything
 * except the potential sum over the planets, which the component warps add in order.  Same expressions as
 * ab_force_eih. */
__device__ void abc_task_eih_source(const AbEphem& E, const AbcSmem& sm, const double* tb, int slot) {
    const double over_C2 = E.over_c_squared;
    const double beta = 1.0;
    const double gamma = 1.0;
    const double xo = 0.0, yo = 0.0, zo = 0.0, vxo = 0.0, vyo = 0.0, vzo = 0.0, axo = 0.0, ayo = 0.0, azo = 0.0;
    const double pix = sm.xv(0, slot), piy = sm.xv(1, slot), piz = sm.xv(2, slot);
    const double pivx = sm.xv(3, slot), pivy = sm.xv(4, slot), pivz = sm.xv(5, slot);
    double term7x_sum = 0.0, term7y_sum = 0.0, term7z_sum = 0.0;
    double term8x_sum = 0.0, term8y_sum = 0.0, term8z_sum = 0.0;
    {
        const double GMj = E.gm[0];
        const double xj = tb[ABC_E_POS(0, 0) * ABC_SLOTS], yj = tb[ABC_E_POS(0, 1) * ABC_SLOTS], zj = tb[ABC_E_POS(0, 2) * ABC_SLOTS];
        const double vxj = tb[ABC_E_SVEL(0) * ABC_SLOTS], vyj = tb[ABC_E_SVEL(1) * ABC_SLOTS], vzj = tb[ABC_E_SVEL(2) * ABC_SLOTS];

        const double dxij = pix + (xo - xj);
        const double dyij = piy + (yo - yj);
        const double dzij = piz + (zo - zj);
        const double rij2 = dxij * dxij + dyij * dyij + dzij * dzij;
        const double _rij = sqrt(rij2);
        const double prefacij = GMj / (rij2 * _rij);

        const double vi2 = pivx * pivx + pivy * pivy + pivz * pivz;
        const double term2 = gamma * over_C2 * vi2;
        const double vj2 = (vxj - vxo) * (vxj - vxo) + (vyj - vyo) * (vyj - vyo) + (vzj - vzo) * (vzj - vzo);
        const double term3 = (1 + gamma) * over_C2 * vj2;
        const double vidotvj = pivx * (vxj - vxo) + pivy * (vyj - vyo) + pivz * (vzj - vzo);
        const double term4 = -2 * (1 + gamma) * over_C2 * vidotvj;
        const double rijdotvj = dxij * (vxj - vxo) + dyij * (vyj - vyo) + dzij * (vzj - vzo);
        const double term5 = -1.5 * over_C2 * (rijdotvj * rijdotvj) / (_rij * _rij);

        const double fx = (2 + 2 * gamma) * pivx - (1 + 2 * gamma) * (vxj - vxo);
        const double fy = (2 + 2 * gamma) * pivy - (1 + 2 * gamma) * (vyj - vyo);
        const double fz = (2 + 2 * gamma) * pivz - (1 + 2 * gamma) * (vzj - vzo);
        const double f = dxij * fx + dyij * fy + dzij * fz;

        const double prefacij_f = prefacij * f;
        term7x_sum += prefacij_f * (pivx - (vxj - vxo));
        term7y_sum += prefacij_f * (pivy - (vyj - vyo));
        term7z_sum += prefacij_f * (pivz - (vzj - vzo));

        double term1 = tb[ABC_E_TERM1 * ABC_SLOTS];
        const double axj = tb[ABC_E_AR(0) * ABC_SLOTS], ayj = tb[ABC_E_AR(1) * ABC_SLOTS], azj = tb[ABC_E_AR(2) * ABC_SLOTS];
        term1 *= -(2 * beta - 1) * over_C2;

        const double rijdotaj = dxij * (axj - axo) + dyij * (ayj - ayo) + dzij * (azj - azo);
        const double term6 = -0.5 * over_C2 * rijdotaj;

        const double term8_fac = GMj / _rij * (3 + 4 * gamma) / 2;
        term8x_sum += term8_fac * axj;
        term8y_sum += term8_fac * ayj;
        term8z_sum += term8_fac * azj;

        sm.con(ABC_C_T + 0, slot) = term1;
        sm.con(ABC_C_T + 1, slot) = term2;
        sm.con(ABC_C_T + 2, slot) = term3;
        sm.con(ABC_C_T + 3, slot) = term4;
        sm.con(ABC_C_T + 4, slot) = term5;
        sm.con(ABC_C_T + 5, slot) = term6;
        sm.con(ABC_C_M + 0, slot) = -prefacij * dxij;
        sm.con(ABC_C_M + 1, slot) = -prefacij * dyij;
        sm.con(ABC_C_M + 2, slot) = -prefacij * dzij;
    }
    sm.con(ABC_C_E78 + 0, slot) = term7x_sum * over_C2 + term8x_sum * over_C2;
    sm.con(ABC_C_E78 + 1, slot) = term7y_sum * over_C2 + term8y_sum * over_C2;
    sm.con(ABC_C_E78 + 2, slot) = term7z_sum * over_C2 + term8z_sum * over_C2;
}

/* The single-body terms through the routines of forces_device.cuh: the particle of the slot as a one-body system,
 * accelerations start from zero, what the routine adds is the term. */
__device__ __forceinline__ void abc_sys_from_slot(const AbcSmem& sm, int slot, AbSysT<1>& S) {
    S.nv_ = 0;
    for (int c = 0; c < 3; c++) {
        S.x[0][c] = sm.xv(c, slot);
        S.v[0][c] = sm.xv(3 + c, slot);
        S.a[0][c] = 0.0;
        S.prm[0][c] = sm.prm(c, slot);
    }
}

__device__ void abc_run_task(const AbEphem& E, const AbForceOpts& F, const AbcSmem& sm, int node, int slot, int kind) {
    const double* tb = sm.tab(node, slot);
    if (kind == ABC_T_EIHSRC) { abc_task_eih_source(E, sm, tb, slot); return; }
    AbSysT<1> S;
    abc_sys_from_slot(sm, slot, S);
    const AbcTabView B(E.gm, tb);
    int dst;
    switch (kind) {
        case ABC_T_EARTHJ: ab_force_earth_harmonics<1, AbcTabView>(E, F, B, S, 0.0, 0.0, 0.0); dst = ABC_C_EARTH; break;
        case ABC_T_SUNJ2: ab_force_solar_j2<1, AbcTabView>(E, F, B, S, 0.0, 0.0, 0.0); dst = ABC_C_SUNJ2; break;
        case ABC_T_NG:
            if (!sm.flag(ABC_SMI_NGON, slot)) return;
            ab_force_nongrav<1, AbcTabView>(F, B, S, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0); dst = ABC_C_NG; break;
        case ABC_T_GRPOT: ab_force_potential_gr<1, AbcTabView>(E, B, S, 0.0, 0.0, 0.0); dst = ABC_C_GRPOT; break;
        case ABC_T_GRSIMPLE: ab_force_simple_gr<1, AbcTabView>(E, B, S, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0); dst = ABC_C_GRSIMPLE; break;
        default: return;
    }
    sm.con(dst + 0, slot) = S.a[0][0];
    sm.con(dst + 1, slot) = S.a[0][1];
    sm.con(dst + 2, slot) = S.a[0][2];
}

/* ------------------------------------------------------------------------------------------ */
/* component warps                                                                            */
/* ------------------------------------------------------------------------------------------ */

struct AbcComp {
    double pos, vel, acc, x0, v0, a0, csx, csv, at;
    double hx, hv;         /* head of the next prediction (abc_predict_stages) */
    double b[7], g[7], e[7], csb[7];
};

/* Acceleration component c of the slot's particle: the terms in the dispatcher's order, the direct terms in
 * the reference's body order (src/forces.c:120-147, 281-306). */
__device__ __forceinline__ double abc_sum_forces(const AbEphem& E, const AbForceOpts& F, const AbcSmem& sm, int c, int slot) {
    double a = 0.0;
    if ((F.forces & 0x08) && sm.flag(ABC_SMI_NGON, slot)) a += sm.con(ABC_C_NG + c, slot);
    if (F.forces & 0x10) a += sm.con(ABC_C_EARTH + c, slot);
    if (F.forces & 0x20) a += sm.con(ABC_C_SUNJ2 + c, slot);
    if (F.forces & 0x40) {
        const double over_C2 = E.over_c_squared;
        const double beta = 1.0;
        const double gamma = 1.0;
        double term0_sum = 0.0;
#pragma unroll
        for (int k = 0; k < AB_NPLANETS; k++) term0_sum += sm.q(k, slot);
        double term0 = term0_sum;
        term0 *= -2 * (beta + gamma) * over_C2;
        const double term1 = sm.con(ABC_C_T + 0, slot), term2 = sm.con(ABC_C_T + 1, slot), term3 = sm.con(ABC_C_T + 2, slot);
        const double term4 = sm.con(ABC_C_T + 3, slot), term5 = sm.con(ABC_C_T + 4, slot), term6 = sm.con(ABC_C_T + 5, slot);
        const double factor = term0 + term1 + term2 + term3 + term4 + term5 + term6;
        a += sm.con(ABC_C_M + c, slot) * factor;
        a += sm.con(ABC_C_E78 + c, slot);
    }
    if (F.forces & 0x100) a += sm.con(ABC_C_GRPOT + c, slot);
    if (F.forces & 0x80) a += sm.con(ABC_C_GRSIMPLE + c, slot);
    if (F.forces & (0x01 | 0x02 | 0x04)) {
        const int ast_num = E.n_ast;
        if (F.forces & 0x04)
            for (int k = 0; k < ast_num; k++) a -= sm.prod(AB_NPLANETS + k, c, slot);
        if (F.forces & 0x02) {
            a -= sm.prod(10, c, slot); a -= sm.prod(4, c, slot); a -= sm.prod(5, c, slot); a -= sm.prod(1, c, slot);
            a -= sm.prod(9, c, slot); a -= sm.prod(8, c, slot); a -= sm.prod(3, c, slot); a -= sm.prod(2, c, slot);
            a -= sm.prod(7, c, slot); a -= sm.prod(6, c, slot);
        }
        if (F.forces & 0x01) a -= sm.prod(0, c, slot);
    }
    return a;
}

/* ab_predict for one component held in registers, in stages: stage j = 6..0 multiplies the running sum into the
 * next power of h and adds b_{j-1} (a0 for j = 0); "final" adds v0 and x0.  An update at node n changes b_0..b_{n-1}
 * only, so the stages 6..n+1 of the prediction for node n + 1 are formed BEFORE the forces of node n have arrived
 * (abc_predict_head, while the workers are busy) and the rest after the update (abc_predict_tail).  Same operations
 * in the same order as the one-piece form (ias15_device.cuh: ab_predict). */
__device__ __forceinline__ void abc_predict_stages(const AbcComp& s, double h, int from, int to, double& px, double& pv) {
    if (from >= 6 && 6 >= to) { px = AB_DIVK(s.b[6] * 7. * h, 9.) + s.b[5];   pv = s.b[6] * 7. * h / 8. + s.b[5]; }
    if (from >= 5 && 5 >= to) { px = px * 3. * h / 4. + s.b[4];               pv = AB_DIVK(pv * 6. * h, 7.) + s.b[4]; }
    if (from >= 4 && 4 >= to) { px = AB_DIVK(px * 5. * h, 7.) + s.b[3];       pv = AB_DIVK(pv * 5. * h, 6.) + s.b[3]; }
    if (from >= 3 && 3 >= to) { px = AB_DIVK(px * 2. * h, 3.) + s.b[2];       pv = AB_DIVK(pv * 4. * h, 5.) + s.b[2]; }
    if (from >= 2 && 2 >= to) { px = AB_DIVK(px * 3. * h, 5.) + s.b[1];       pv = pv * 3. * h / 4. + s.b[1]; }
    if (from >= 1 && 1 >= to) { px = px * h / 2. + s.b[0];                    pv = AB_DIVK(pv * 2. * h, 3.) + s.b[0]; }
    if (from >= 0 && 0 >= to) { px = AB_DIVK(px * h, 3.) + s.a0;              pv = pv * h / 2. + s.a0; }
}

__device__ __forceinline__ void abc_predict_final(const AbcComp& s, double h, double dt, double px, double pv, double& xk_out, double& vk_out) {
    px = px * dt * h / 2. + s.v0;
    const double xk = -s.csx + px * dt * h;
    xk_out = xk + s.x0;
    const double vk = -s.csv + pv * dt * h;
    vk_out = vk + s.v0;
}

/* the whole prediction at node nn */
__device__ __forceinline__ void abc_predict(const AbcComp& s, int nn, double dt, double& xk_out, double& vk_out) {
    const double h = c_h[nn];
    double px = 0.0, pv = 0.0;
    abc_predict_stages(s, h, 6, 0, px, pv);
    abc_predict_final(s, h, dt, px, pv, xk_out, vk_out);
}

#define ABC_DIVRR(x, k) ab_divc((x), c_rr[k], c_rri[k])

/* ab_update_gb for one component held in registers; returns |change of b6| at node 7 */
__device__ __forceinline__ double abc_update_gb(AbcComp& s, int nn, double at) {
    const double gk = at + (-s.a0);
    double tmp = 0.0, gn;
    switch (nn) {
        case 1:
            tmp = s.g[0];
            gn = ABC_DIVRR(gk, 0);
            s.g[0] = gn;
            ab_add_cs(s.b[0], s.csb[0], gn - tmp);
            break;
        case 2:
            tmp = s.g[1];
            gn = ABC_DIVRR(ABC_DIVRR(gk, 1) - s.g[0], 2);
            s.g[1] = gn;
            tmp = gn - tmp;
            ab_add_cs(s.b[0], s.csb[0], tmp * c_c[0]);
            ab_add_cs(s.b[1], s.csb[1], tmp);
            break;
        case 3:
            tmp = s.g[2];
            gn = ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(gk, 3) - s.g[0], 4) - s.g[1], 5);
            s.g[2] = gn;
            tmp = gn - tmp;
            ab_add_cs(s.b[0], s.csb[0], tmp * c_c[1]);
            ab_add_cs(s.b[1], s.csb[1], tmp * c_c[2]);
            ab_add_cs(s.b[2], s.csb[2], tmp);
            break;
        case 4:
            tmp = s.g[3];
            gn = ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(gk, 6) - s.g[0], 7) - s.g[1], 8) - s.g[2], 9);
            s.g[3] = gn;
            tmp = gn - tmp;
            ab_add_cs(s.b[0], s.csb[0], tmp * c_c[3]);
            ab_add_cs(s.b[1], s.csb[1], tmp * c_c[4]);
            ab_add_cs(s.b[2], s.csb[2], tmp * c_c[5]);
            ab_add_cs(s.b[3], s.csb[3], tmp);
            break;
        case 5:
            tmp = s.g[4];
            gn = ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(gk, 10) - s.g[0], 11) - s.g[1], 12) - s.g[2], 13) - s.g[3], 14);
            s.g[4] = gn;
            tmp = gn - tmp;
            ab_add_cs(s.b[0], s.csb[0], tmp * c_c[6]);
            ab_add_cs(s.b[1], s.csb[1], tmp * c_c[7]);
            ab_add_cs(s.b[2], s.csb[2], tmp * c_c[8]);
            ab_add_cs(s.b[3], s.csb[3], tmp * c_c[9]);
            ab_add_cs(s.b[4], s.csb[4], tmp);
            break;
        case 6:
            tmp = s.g[5];
            gn = ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(gk, 15) - s.g[0], 16) - s.g[1], 17) - s.g[2], 18) - s.g[3], 19) - s.g[4], 20);
            s.g[5] = gn;
            tmp = gn - tmp;
            ab_add_cs(s.b[0], s.csb[0], tmp * c_c[10]);
            ab_add_cs(s.b[1], s.csb[1], tmp * c_c[11]);
            ab_add_cs(s.b[2], s.csb[2], tmp * c_c[12]);
            ab_add_cs(s.b[3], s.csb[3], tmp * c_c[13]);
            ab_add_cs(s.b[4], s.csb[4], tmp * c_c[14]);
            ab_add_cs(s.b[5], s.csb[5], tmp);
            break;
        default:
            tmp = s.g[6];
            gn = ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(ABC_DIVRR(gk, 21) - s.g[0], 22) - s.g[1], 23) - s.g[2], 24) - s.g[3], 25) - s.g[4], 26) - s.g[5], 27);
            s.g[6] = gn;
            tmp = gn - tmp;
            ab_add_cs(s.b[0], s.csb[0], tmp * c_c[15]);
            ab_add_cs(s.b[1], s.csb[1], tmp * c_c[16]);
            ab_add_cs(s.b[2], s.csb[2], tmp * c_c[17]);
            ab_add_cs(s.b[3], s.csb[3], tmp * c_c[18]);
            ab_add_cs(s.b[4], s.csb[4], tmp * c_c[19]);
            ab_add_cs(s.b[5], s.csb[5], tmp * c_c[20]);
            ab_add_cs(s.b[6], s.csb[6], tmp);
            break;
    }
    return fabs(tmp);
}

/* ab_attempt_begin for one component */
__device__ __forceinline__ void abc_attempt_begin(AbcComp& s) {
    s.x0 = s.pos; s.v0 = s.vel; s.a0 = s.acc;
    const double b0 = s.b[0], b1 = s.b[1], b2 = s.b[2], b3 = s.b[3], b4 = s.b[4], b5 = s.b[5], b6 = s.b[6];
    s.csb[0] = 0.; s.csb[1] = 0.; s.csb[2] = 0.; s.csb[3] = 0.; s.csb[4] = 0.; s.csb[5] = 0.; s.csb[6] = 0.;
    s.g[0] = b6 * c_d[15] + b5 * c_d[10] + b4 * c_d[6] + b3 * c_d[3] + b2 * c_d[1] + b1 * c_d[0] + b0;
    s.g[1] = b6 * c_d[16] + b5 * c_d[11] + b4 * c_d[7] + b3 * c_d[4] + b2 * c_d[2] + b1;
    s.g[2] = b6 * c_d[17] + b5 * c_d[12] + b4 * c_d[8] + b3 * c_d[5] + b2;
    s.g[3] = b6 * c_d[18] + b5 * c_d[13] + b4 * c_d[9] + b3;
    s.g[4] = b6 * c_d[19] + b5 * c_d[14] + b4;
    s.g[5] = b6 * c_d[20] + b5;
    s.g[6] = b6;
}

/* ab_predict_next for one component: (e_j, b_j given as scalars) -> s.e, s.b.  Scalars, not pointers: an array whose
 * address is taken would put the whole register state of the component warps into local memory. */
__device__ __forceinline__ void abc_predict_next(AbcComp& s, double ratio,
                                                 double se0, double se1, double se2, double se3, double se4, double se5, double se6,
                                                 double _b0, double _b1, double _b2, double _b3, double _b4, double _b5, double _b6) {
    if (ratio > 20.) {
        s.e[0] = 0.; s.e[1] = 0.; s.e[2] = 0.; s.e[3] = 0.; s.e[4] = 0.; s.e[5] = 0.; s.e[6] = 0.;
        s.b[0] = 0.; s.b[1] = 0.; s.b[2] = 0.; s.b[3] = 0.; s.b[4] = 0.; s.b[5] = 0.; s.b[6] = 0.;
        return;
    }
    const double q1 = ratio;
    const double q2 = q1 * q1;
    const double q3 = q1 * q2;
    const double q4 = q2 * q2;
    const double q5 = q2 * q3;
    const double q6 = q3 * q3;
    const double q7 = q3 * q4;
    const double be0 = _b0 - se0;
    const double be1 = _b1 - se1;
    const double be2 = _b2 - se2;
    const double be3 = _b3 - se3;
    const double be4 = _b4 - se4;
    const double be5 = _b5 - se5;
    const double be6 = _b6 - se6;
    const double e0 = q1 * (_b6 * 7.0 + _b5 * 6.0 + _b4 * 5.0 + _b3 * 4.0 + _b2 * 3.0 + _b1 * 2.0 + _b0);
    const double e1 = q2 * (_b6 * 21.0 + _b5 * 15.0 + _b4 * 10.0 + _b3 * 6.0 + _b2 * 3.0 + _b1);
    const double e2 = q3 * (_b6 * 35.0 + _b5 * 20.0 + _b4 * 10.0 + _b3 * 4.0 + _b2);
    const double e3 = q4 * (_b6 * 35.0 + _b5 * 15.0 + _b4 * 5.0 + _b3);
    const double e4 = q5 * (_b6 * 21.0 + _b5 * 6.0 + _b4);
    const double e5 = q6 * (_b6 * 7.0 + _b5);
    const double e6 = q7 * _b6;
    s.e[0] = e0; s.e[1] = e1; s.e[2] = e2; s.e[3] = e3; s.e[4] = e4; s.e[5] = e5; s.e[6] = e6;
    s.b[0] = e0 + be0; s.b[1] = e1 + be1; s.b[2] = e2 + be2; s.b[3] = e3 + be3;
    s.b[4] = e4 + be4; s.b[5] = e5 + be5; s.b[6] = e6 + be6;
}

/* ab_advance for one component (x0, v0 with compensated sums; particles <- x0, v0) */
__device__ __forceinline__ void abc_advance(AbcComp& s, double dt_done) {
    const double b0 = s.b[0], b1 = s.b[1], b2 = s.b[2], b3 = s.b[3], b4 = s.b[4], b5 = s.b[5], b6 = s.b[6];
    double x0 = s.x0, v0 = s.v0;
    const double a0 = s.a0;
    double csx = s.csx, csv = s.csv;
    ab_add_cs(x0, csx, AB_DIVK(b6, 72.) * dt_done * dt_done);
    ab_add_cs(x0, csx, AB_DIVK(b5, 56.) * dt_done * dt_done);
    ab_add_cs(x0, csx, AB_DIVK(b4, 42.) * dt_done * dt_done);
    ab_add_cs(x0, csx, AB_DIVK(b3, 30.) * dt_done * dt_done);
    ab_add_cs(x0, csx, AB_DIVK(b2, 20.) * dt_done * dt_done);
    ab_add_cs(x0, csx, AB_DIVK(b1, 12.) * dt_done * dt_done);
    ab_add_cs(x0, csx, AB_DIVK(b0, 6.) * dt_done * dt_done);
    ab_add_cs(x0, csx, a0 / 2. * dt_done * dt_done);
    ab_add_cs(x0, csx, v0 * dt_done);
    ab_add_cs(v0, csv, b6 / 8. * dt_done);
    ab_add_cs(v0, csv, AB_DIVK(b5, 7.) * dt_done);
    ab_add_cs(v0, csv, AB_DIVK(b4, 6.) * dt_done);
    ab_add_cs(v0, csv, AB_DIVK(b3, 5.) * dt_done);
    ab_add_cs(v0, csv, b2 / 4. * dt_done);
    ab_add_cs(v0, csv, AB_DIVK(b1, 3.) * dt_done);
    ab_add_cs(v0, csv, b0 / 2. * dt_done);
    ab_add_cs(v0, csv, a0 * dt_done);
    s.x0 = x0; s.v0 = v0; s.csx = csx; s.csv = csv;
    s.pos = x0; s.vel = v0;
}

/* working-batch accessors: component k of the slot's system (C = 3: no variational particles) */
#define ABC_W1(arr, k) (arr)[(long long)(k) * wn + ws]
#define ABC_W7(arr, j, k) (arr)[((long long)(j) * 3 + (k)) * wn + ws]

__device__ __forceinline__ void abc_comp_load(const AbBatch& W, long long ws, int c, AbcComp& s) {
    const long long wn = W.n;
    s.pos = ABC_W1(W.pos, c); s.vel = ABC_W1(W.vel, c); s.acc = ABC_W1(W.acc, c);
    s.x0 = ABC_W1(W.x0, c); s.v0 = ABC_W1(W.v0, c); s.a0 = ABC_W1(W.a0, c);
    s.csx = ABC_W1(W.csx, c); s.csv = ABC_W1(W.csv, c);
    s.at = 0.0;
#pragma unroll
    for (int j = 0; j < 7; j++) {
        s.b[j] = ABC_W7(W.b, j, c); s.g[j] = ABC_W7(W.g, j, c); s.e[j] = ABC_W7(W.e, j, c); s.csb[j] = ABC_W7(W.csb, j, c);
    }
}

__device__ __forceinline__ void abc_comp_store(const AbBatch& W, long long ws, int c, const AbcComp& s) {
    const long long wn = W.n;
    ABC_W1(W.pos, c) = s.pos; ABC_W1(W.vel, c) = s.vel; ABC_W1(W.acc, c) = s.acc;
    ABC_W1(W.x0, c) = s.x0; ABC_W1(W.v0, c) = s.v0; ABC_W1(W.a0, c) = s.a0;
    ABC_W1(W.csx, c) = s.csx; ABC_W1(W.csv, c) = s.csv;
#pragma unroll
    for (int j = 0; j < 7; j++) {
        ABC_W7(W.b, j, c) = s.b[j]; ABC_W7(W.g, j, c) = s.g[j]; ABC_W7(W.e, j, c) = s.e[j]; ABC_W7(W.csb, j, c) = s.csb[j];
    }
}

}  // namespace AB_NS
#endif
