/*
 * coop_device.cuh -- per-particle-dt IAS15 with a CTA stepping 2 x 32 systems together.
 *
 * Why: one thread per system (kernels.cu, pp_queue_kernel) keeps the 8 body tables of a step (6 KB) and the
 * IAS15 tables in thread-local / global memory and walks one dependent chain per system.  Here a system's
 * working set lives ON CHIP -- the planets at its 8 node times in shared memory, its IAS15 state in the
 * registers of three "component" warps -- and the independent work of a step is spread over warps:
 *
 *     group = 32 systems with 8 warps of their own, a 113.5 KB block of shared memory and a 110 KB table in global
 *             memory (L2) for what one task reads once per node (asteroid positions, Sun velocity, EIH pair sums).
 *             A CTA holds TWO groups that walk through the phases of a step attempt together (CTA-wide barriers):
 *             every latency-bound phase serves 64 systems, and all 16 warps of the SM run the same code at the
 *             same time (the hot code of the roles is larger than the instruction cache).
 *     lane  = system slot (structure-of-arrays in shared memory with stride 32: every access of a warp is
 *             32 consecutive doubles, conflict free)
 *     warp  = task (index inside the group)
 *         warps 0-2    component x / y / z of every system: predictor, ordered force sums, g/b update,
 *                      advance, predict_next  (b, g, e, csb, x0, v0, a0, cs* in registers)
 *         warp  3      control: queue, reb_simulation_integrate bookkeeping, output epochs, convergence and
 *                      step-size control (sqrt7)
 *         warps 4-7    workers: their share of the 27 bodies of the direct term (+ the planets' terms of the EIH
 *                      potential sum) and of the single-body terms (Earth J2-J4, solar J2, EIH source block,
 *                      Marsden, GR variants), from a launch-time plan (gpu_api.cu: build_coop_plan)
 *         all 8        the Chebyshev fill of the node tables: thread = (slot, node), so the eight nodes of a
 *                      system read the same one or two coefficient records -- which the TMA unit copies into
 *                      shared memory one series ahead (cp.async.bulk per slot, mbarrier completion) -- and
 *                      everything one (slot, node) needs comes from one thread
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
 * (tests/emul): that is how the control flow is checked without a GPU.
 */
#ifndef AB_COOP_DEVICE_CUH
#define AB_COOP_DEVICE_CUH

#include "device_types.h"
#include "ephem_device.cuh"
#include "forces_device.cuh"
#include "ias15_device.cuh"
#include <string.h>

/* node table in shared memory: the planets' positions of one (node, slot).  What a single task reads (asteroid
 * positions, the Sun's velocity, the particle-independent EIH sums) lives in the CTA's global table, device_types.h. */
#define ABC_TAB_E 33
#define ABC_NODE_STRIDE (ABC_TAB_E * ABC_SLOTS + 2)    /* +2: the (slot, node) lanes of a half-warp of the fill fall into 16 different banks */
#define ABC_E_POS(i, c) ((i) * 3 + (c))

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

/* shared-memory carve-up of a group (doubles first, then ints): 113.5 KB, two groups fill the 227 KB of an SM.
 * XV .. MON are scratch of the node rounds and dead while the node tables are filled; together with the extension
 * behind them they are the STAGING area of the fill (coefficient records copied in by the TMA unit, see below). */
#define ABC_SM_TAB 0
#define ABC_SM_PRM (ABC_SM_TAB + 8 * ABC_NODE_STRIDE)
#define ABC_SM_T0 (ABC_SM_PRM + 3 * ABC_SLOTS)               /* start time of the attempt */
#define ABC_SM_DT (ABC_SM_T0 + ABC_SLOTS)                    /* step of the attempt */
#define ABC_SM_RATIO (ABC_SM_DT + ABC_SLOTS)                 /* predict_next ratio */
#define ABC_SM_XV (ABC_SM_RATIO + ABC_SLOTS)
#define ABC_SM_PROD (ABC_SM_XV + 6 * ABC_SLOTS)
#define ABC_SM_Q (ABC_SM_PROD + 81 * ABC_SLOTS)
#define ABC_SM_CON (ABC_SM_Q + 11 * ABC_SLOTS)
#define ABC_SM_MON (ABC_SM_CON + ABC_NCON * ABC_SLOTS)       /* |a| x3, |db6| x3, |b6| x3 */
#define ABC_SM_EXT (ABC_SM_MON + 9 * ABC_SLOTS)
#define ABC_SM_EXT_DOUBLES 1488
#define ABC_SM_DOUBLES (ABC_SM_EXT + ABC_SM_EXT_DOUBLES)
/* staging area: per warp two stage buffers of ABC_ST_BUF doubles; behind them, outside the scratch that the node rounds
 * overwrite, per warp two mbarriers and one word with their phase bits */
#define ABC_ST_BASE ABC_SM_XV
#define ABC_ST_MBAR (ABC_SM_DOUBLES - 3 * ABC_GWARPS)
#define ABC_ST_WARP (((ABC_ST_MBAR - ABC_ST_BASE) / ABC_GWARPS) & ~1)
#define ABC_ST_BUF ((ABC_ST_WARP / 2) & ~1)
#define ABC_SMI_ACTIVE 0      /* slot takes part in this attempt */
#define ABC_SMI_NEEDA0 1      /* slot needs the force evaluation at the start of the step */
#define ABC_SMI_SW 2          /* slot is still sweeping */
#define ABC_SMI_DEC 3         /* 0 nothing, 1 accepted, 2 rejected, 3 rejected and no previous step */
#define ABC_SMI_NGON 4        /* Marsden term active for this slot */
#define ABC_SMI_ERR 5         /* ephemeris status of the fill */
#define ABC_SM_INTS (6 * ABC_SLOTS)
#define ABC_SMEM_GROUP_BYTES ((size_t)ABC_SM_DOUBLES * 8 + (size_t)ABC_SM_INTS * 4)
#define ABC_SMEM_BYTES (ABC_GROUPS * ABC_SMEM_GROUP_BYTES)
static_assert(ABC_SMEM_BYTES <= 232448, "two groups must fit the 227 KB of dynamic shared memory of an SM");
static_assert((ABC_ST_BASE & 1) == 0 && (ABC_SMEM_GROUP_BYTES % 16) == 0, "stage buffers are 16-byte aligned");

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
#define ABC_SYNC_OR(p) abc_sync_or(p)
#define ABC_CTXARG
#define ABC_CTXPASS
#define ABC_BLOCK ((int)blockIdx.x * ABC_GROUPS + (int)(threadIdx.x / (32 * ABC_GWARPS)))     /* index of the group */
#endif

namespace AB_NS {

#ifndef AB_HOST_EMUL
/* __syncthreads_or whose result passes through a volatile local: ptxas (12.9) otherwise REMATERIALISES the predicate
 * where it is live across a region of high register pressure -- by executing the barrier a second time (seen in the
 * SASS of the component warps: BAR.RED.OR twice, first result dropped), which desynchronises the CTA. */
__device__ __forceinline__ bool abc_sync_or(int p) {
    volatile int keep = __syncthreads_or(p);
    return keep != 0;
}
#endif

/* Launch-time copies in CONSTANT memory for the out-of-line fill routine: a kernel parameter reached through a
 * reference is read with generic loads (hundreds of cycles each, one after the other in the record look-up); these
 * are read through the constant cache.  The asteroid descriptors, which AbEphem keeps in global memory, come along. */
static __constant__ AbEphem c_abcE;
static __constant__ AbForceOpts c_abcF;
static __constant__ AbSpkTarget c_abc_ast[AB_MAX_AST];
static __constant__ AbcPlan c_abcP;
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
/* the Sun's velocity of one (node, slot): in the CTA's global table, read where a term needs it (Marsden, simple GR) */
struct AbcVelG {
    const double* g;
    __device__ __forceinline__ double operator[](int c) const { return __ldcg(g + c * ABC_SLOTS); }
};
struct AbcVelRows {
    AbcVelG s;
    __device__ __forceinline__ AbcVelG operator[](int) const { return s; }
};
struct AbcTabView {
    const double* gm;
    AbcRows pos;
    AbcVelRows vel;
    double earth_acc[3];
    __device__ __forceinline__ AbcTabView(const double* gm_, const double* node_slot, const double* gt_node_slot) : gm(gm_) {
        pos.base = node_slot; pos.stride = 3 * ABC_SLOTS;
        vel.s.g = gt_node_slot + ABC_GT_SVEL(0) * ABC_SLOTS;
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
    double* gtab;                    /* [grid][ABC_GT_DOUBLES]: the CTAs' global tables */
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

/* ---- the fill: thread = (slot, node), lane = 8 * (slot within the warp's four) + node.  A thread evaluates the Sun
 * (with its velocity), then the EMB, the planets and the asteroids, then the particle-independent EIH sums of its
 * node: everything one (slot, node) needs comes from one thread, so the fill has no barrier inside.
 *
 * The coefficient records are read straight from the packed image with 16-byte loads (the eight node-lanes of a slot
 * read the same one or two records: a warp's load touches a handful of lines).  With two CTAs' tables in shared memory
 * the L1 is a few KB, i.e. every record comes from L2: a thread runs ABC_FILL_NS series side by side and loads the
 * coefficients of the next pair of terms while it works on the current one.  Arithmetic per (series, time) is that of
 * ephem_device.cuh: same sums, same order. */
#ifndef ABC_FILL_NS
#define ABC_FILL_NS 2
#endif

struct AbcSeriesRef {
    const double* img;
    const AbSpkTarget* tg;
    int kind;              /* 0 EMB, 2 planet `idx`, 3 asteroid `idx`, -1 none */
    int idx;
};

/* series s of the list: EMB, Mercury .. Pluto (the Sun is evaluated on its own, with velocity), then the asteroids */
__device__ __forceinline__ AbcSeriesRef abc_series_ref(int s, int s_end) {
    const AbEphem& E = c_abcE;
    AbcSeriesRef r;
    if (s >= s_end) { r.img = E.spka_img; r.tg = &c_abc_ast[0]; r.kind = -1; r.idx = 0; return r; }
    if (s < AB_NPLANETS) {
        r.img = E.spkp_img;
        if (s == 0) { r.kind = 0; r.idx = -1; r.tg = &E.p_tgt[E.emb_index]; }
        else { r.idx = s; r.kind = 2; r.tg = &E.p_tgt[E.p_index[s]]; }
    } else {
        const int m = s - AB_NPLANETS;
        r.img = E.spka_img; r.kind = 3; r.idx = m; r.tg = &c_abc_ast[m];
    }
    return r;
}

/* Position sums (file units) of NS series at time t, the recurrences side by side, coefficients one pair of terms ahead. */
template <int NS>
__device__ __forceinline__ void abc_multi_eval(const AbcSeriesRef* R, double jd_ref, double t, double (*u)[3]) {
    const double2* q[NS];
    int P[NS];
    double z[NS], T1[NS], T2[NS], a0[NS], a1[NS], a2[NS];
    double2 cA[NS], cB[NS], cC[NS];
    int Pmax = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
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
    for (int s = 0; s < NS; s++) {
        /* p = 0 (T = 1) and p = 1 (T = z) */
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        cA[s] = cB[s] = cC[s] = double2{0.0, 0.0};
        if (P[s] > 0) {
            const double2 A = __ldg(q[s]), B = __ldg(q[s] + 1), C = __ldg(q[s] + 2);      /* x0 y0 | z0 x1 | y1 z1 */
            if (P[s] > 2) {
                cA[s] = __ldg(q[s] + 3); cB[s] = __ldg(q[s] + 4);
                if (P[s] > 3) cC[s] = __ldg(q[s] + 5);
            }
            s0 += A.x * 1.0; s1 += A.y * 1.0; s2 += B.x * 1.0;
            s0 += B.y * z[s]; s1 += C.x * z[s]; s2 += C.y * z[s];
        }
        a0[s] = s0; a1[s] = s1; a2[s] = s2;
        T2[s] = 1.0; T1[s] = z[s];
    }
#pragma unroll 1
    for (int p = 2; p < Pmax; p += 2) {
        double2 nA[NS], nB[NS], nC[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            nA[s] = nB[s] = nC[s] = double2{0.0, 0.0};
            if (p + 2 < P[s]) {
                const double2* qq = q[s] + 3 * ((p + 2) >> 1);
                nA[s] = __ldg(qq); nB[s] = __ldg(qq + 1);                                /* xp yp | zp xp+1 */
                if (p + 3 < P[s]) nC[s] = __ldg(qq + 2);                                  /* yp+1 zp+1 */
            }
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (p < P[s]) {
                const double2 A = cA[s], B = cB[s];
                const double Ta = 2.0 * z[s] * T1[s] - T2[s];
                a0[s] += A.x * Ta; a1[s] += A.y * Ta; a2[s] += B.x * Ta;
                if (p + 1 < P[s]) {      /* an odd number of terms: the second half of the last pair is padding */
                    const double2 C = cC[s];
                    const double Tb = 2.0 * z[s] * Ta - T1[s];
                    a0[s] += B.y * Tb; a1[s] += C.x * Tb; a2[s] += C.y * Tb;
                    T2[s] = Ta; T1[s] = Tb;
                }
            }
            cA[s] = nA[s]; cB[s] = nB[s]; cC[s] = nC[s];
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++) { u[s][0] = a0[s]; u[s][1] = a1[s]; u[s][2] = a2[s]; }
}

/* what becomes of the sums of one series: EMB kept, planets into shared memory, asteroids (heliocentric / 149597870.7,
 * reference src/spk.c:470, + Sun, reference src/forces.c:213-219) into the global table */
__device__ __forceinline__ void abc_fill_store(const AbcSeriesRef& R, double* u, double* emb, double* tb, double* g,
                                               double sx, double sy, double sz) {
    const AbEphem& E = c_abcE;
    if (R.kind == 0) {
        emb[0] = u[0]; emb[1] = u[1]; emb[2] = u[2];
    } else if (R.kind == 3) {
        const int m = R.idx;
        g[ABC_GT_AST(m, 0) * ABC_SLOTS] = AB_DIVK(u[0], 149597870.7) + sx;
        g[ABC_GT_AST(m, 1) * ABC_SLOTS] = AB_DIVK(u[1], 149597870.7) + sy;
        g[ABC_GT_AST(m, 2) * ABC_SLOTS] = AB_DIVK(u[2], 149597870.7) + sz;
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
#define ABC_SM_HERE(sm) AbcSmem sm; sm.d = abc_shared + (threadIdx.x / (32 * ABC_GWARPS)) * (ABC_SMEM_GROUP_BYTES / 8); sm.i = reinterpret_cast<int*>(sm.d + ABC_SM_DOUBLES)
#endif

/* particle-independent EIH sums of the Sun at one (slot, node) (ab_fill_nodes, same operations): the lane reads back
 * the eleven positions it has just written */
__device__ __forceinline__ void abc_fill_eih(const double* tb, double* g, double sx, double sy, double sz) {
    const AbEphem& E = c_abcE;
    const AbForceOpts& F = c_abcF;
    if (F.forces & 0x40) {
        double term1 = 0.0, arx = 0.0, ary = 0.0, arz = 0.0;
#pragma unroll 1
        for (int k0 = 1; k0 < AB_NPLANETS; k0 += 5) {
            /* five terms at a time: the square roots and divisions of a group are independent and overlap,
             * then the group is added in order */
            double t1[5], fx[5], fy[5], fz[5], GMk[5], dxjk[5], dyjk[5], dzjk[5], rjk2[5], _rjk[5], den[5];
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const int k = k0 + j;
                GMk[j] = E.gm[k];
                dxjk[j] = sx - tb[ABC_E_POS(k, 0) * ABC_SLOTS];
                dyjk[j] = sy - tb[ABC_E_POS(k, 1) * ABC_SLOTS];
                dzjk[j] = sz - tb[ABC_E_POS(k, 2) * ABC_SLOTS];
                rjk2[j] = dxjk[j] * dxjk[j] + dyjk[j] * dyjk[j] + dzjk[j] * dzjk[j];
                ok = ok && ab_nb_ok(rjk2[j]) && ab_nb_ok(GMk[j]);
            }
#if !AB_STRICT && !defined(AB_HOST_EMUL)
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const double ir = ab_rsqrt_fast(rjk2[j]);
                _rjk[j] = rjk2[j] * ir; den[j] = rjk2[j] * _rjk[j];
                t1[j] = GMk[j] * ir; fx[j] = t1[j] * (ir * ir);
            }
#else
#pragma unroll
            for (int j = 0; j < 5; j++) _rjk[j] = ab_sqrt_nb(rjk2[j]);
#pragma unroll
            for (int j = 0; j < 5; j++) { den[j] = rjk2[j] * _rjk[j]; ok = ok && ab_nb_ok(den[j]); }
#pragma unroll
            for (int j = 0; j < 5; j++) { t1[j] = ab_div_nb(GMk[j], _rjk[j]); fx[j] = ab_div_nb(GMk[j], den[j]); }
#endif
            if (!ok) {
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    _rjk[j] = sqrt(rjk2[j]);
                    t1[j] = GMk[j] / _rjk[j];
                    fx[j] = GMk[j] / (rjk2[j] * _rjk[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const double fac = fx[j];
                fx[j] = fac * dxjk[j]; fy[j] = fac * dyjk[j]; fz[j] = fac * dzjk[j];
            }
#pragma unroll
            for (int j = 0; j < 5; j++) { term1 += t1[j]; arx -= fx[j]; ary -= fy[j]; arz -= fz[j]; }
        }
        g[ABC_GT_TERM1 * ABC_SLOTS] = term1;
        g[ABC_GT_AR(0) * ABC_SLOTS] = arx; g[ABC_GT_AR(1) * ABC_SLOTS] = ary; g[ABC_GT_AR(2) * ABC_SLOTS] = arz;
    }

}

/* ---- TMA-staged fill (ABC_FILL_TMA, the default) ---------------------------------------------------------------
 * The eight node-lanes of a slot read the same one or two coefficient records (a record spans 4-32 days, a step a few
 * days to a few weeks).  Read with a 16-byte load per lane and pair of terms (the direct fill below, kept as a build
 * option), the fill ISSUES 750 instructions per series and warp for 170 FP64 ones -- address arithmetic, prefetch
 * rotation, predicates for odd term counts -- and every load waits for L2: measured 130 k cycles per attempt of a CTA.
 * Here the records are COPIED INTO SHARED MEMORY BY THE TMA UNIT (cp.async.bulk: one request per slot covering the
 * consecutive records its eight nodes fall into, completion on an mbarrier) one series ahead of the arithmetic, and
 * the lanes read the coefficients from the staged copy with broadcast shared-memory loads, four terms per trip (the
 * packed image pads every record with zero coefficients to a multiple of four terms, see pack_spk in gpu_api.cu).
 *
 * Per warp and series: every lane locates its record (integer arithmetic on the segment descriptor in constant
 * memory, no global access); a slot's records are consecutive in the image, so one bulk copy per slot brings them all;
 * the four slots share a stage buffer of ABC_ST_BUF doubles (7-13 records), handed out in slot order.  A lane whose
 * record did not fit (the Moon's 4-day records under a 20-day step), or whose slot straddles two segments, reads from
 * the image in global memory: same arithmetic, the values are those of the direct fill bit for bit.  Two stage
 * buffers per warp: series s + 1 is in flight while series s is evaluated. */
#ifndef ABC_FILL_TMA
#define ABC_FILL_TMA 1
#endif

#ifdef AB_HOST_EMUL
static inline int ab_ffs(unsigned v) { return __builtin_ffs((int)v); }
#else
__device__ __forceinline__ int ab_ffs(unsigned v) { return __ffs((int)v); }
#endif

/* series of the staged fill */
#define ABC_S_SUN 0
#define ABC_S_EMB 1
#define ABC_S_PLANET0 1                      /* series of planet b (1 .. 10) is ABC_S_PLANET0 + b */
#define ABC_S_AST0 (AB_NPLANETS + 1)
#define ABC_NSERIES (ABC_S_AST0 + AB_MAX_AST)
static __constant__ AbSpkTarget c_abc_tg[ABC_NSERIES];   /* the targets in series order, seg[].stage_cap filled in */
static __constant__ int c_abc_regular;                   /* usual SPK layout: every body and the EMB have a target */

/* Launch-time table of the series (host).  stage_cap: how many records of a segment fit a stage buffer. */
static inline int abc_series_table(const AbEphem& E, const AbSpkTarget* ast, AbSpkTarget* out) {
    memset(out, 0, sizeof(AbSpkTarget) * ABC_NSERIES);
    int regular = (E.planets_source != AB_SRC_ASCII) && E.emb_index >= 0;
    for (int b = 0; b < AB_NPLANETS && regular; b++) if (E.p_index[b] < 0) regular = 0;
    if (regular) {
        out[ABC_S_SUN] = E.p_tgt[E.p_index[0]];
        out[ABC_S_EMB] = E.p_tgt[E.emb_index];
        for (int b = 1; b < AB_NPLANETS; b++) out[ABC_S_PLANET0 + b] = E.p_tgt[E.p_index[b]];
    }
    for (int m = 0; m < E.n_ast && m < AB_MAX_AST; m++) out[ABC_S_AST0 + m] = ast[m];
    for (int s = 0; s < ABC_NSERIES; s++)
        for (int k = 0; k < AB_MAXSEG; k++) out[s].seg[k].stage_cap = out[s].seg[k].R > 0 ? ABC_ST_BUF / out[s].seg[k].R : 0;
    /* same_grid = 1: the series has the time grid of the series before it (same segments, same record boundaries, same record
     * size -- the 16 asteroids of sb441-n16, the outer planets): segment, record index and staging layout of a lane are
     * then the same numbers, formed from the same operands, and are not formed again */
    for (int s = 0; s < ABC_NSERIES; s++) {
        out[s].same_grid = 0;
        if (s == 0 || out[s].nseg < 1 || out[s].nseg != out[s - 1].nseg) continue;
        const AbSpkTarget &a = out[s], &b = out[s - 1];
        bool same = a.beg == b.beg && a.end == b.end && a.res == b.res && a.res_rd == b.res_rd;
        for (int k = 0; same && k < a.nseg; k++) {
            const AbSpkSeg &p = a.seg[k], &q = b.seg[k];
            same = p.R == q.R && p.nrec == q.nrec && p.jul_init == q.jul_init && p.intlen_d == q.intlen_d &&
                   p.intlen_rd == q.intlen_rd && p.stage_cap == q.stage_cap;
        }
        out[s].same_grid = same ? 1 : 0;
    }
    return regular;
}

#if ABC_FILL_TMA
#ifndef AB_HOST_EMUL
__device__ __forceinline__ unsigned abc_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void abc_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void abc_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void abc_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool abc_mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
/* global -> shared bulk copy by the TMA unit; `bytes` a multiple of 16, both addresses 16-byte aligned */
__device__ __forceinline__ void abc_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void abc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* once per kernel, by every warp for its own two barriers (the first CTA barrier of the role loops follows) */
__device__ __forceinline__ void abc_stage_init() {
    ABC_SM_HERE(sm);
    const int warp = (int)(threadIdx.x >> 5) % ABC_GWARPS;
    if ((threadIdx.x & 31) == 0) {
        double* mb = sm.d + ABC_ST_MBAR + 3 * warp;
        abc_mbar_init(abc_smem_addr(mb), 1);
        abc_mbar_init(abc_smem_addr(mb + 1), 1);
        *reinterpret_cast<unsigned*>(mb + 2) = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        abc_fence_proxy_async();
    }
    __syncwarp();
}
#endif

/* what a lane carries through the fill */
struct AbcFillLane {
    double t, sx, sy, sz, emb[3];
    double *tb, *g;
    int active;
    int seg, off;            /* series being evaluated: its segment, and where its record lies in the stage buffer (doubles; < 0: not staged) */
    int nseg, noff;          /* the same for the series after it */
    int lo, nfit, offr, uni; /* the slot's run of records as the last located series staged it: first record, records staged, offset in
                              * the stage buffer (doubles), and whether the warp staged at all -- reused by series on the same time grid */
};

/* Chebyshev argument (and derivative scale) of time t for the record at `rec` = [_jul(MID), RADIUS, ...]: the second
 * half of ab_spk_record_in */
__device__ __forceinline__ double abc_record_z(const double* rec, const AbSpkSeg& sg, double jd_ref, double t, double* c) {
    const double jul_mid = rec[0];
    if (sg.uniform) {
        *c = sg.radius_inv;
        return ab_divc((jd_ref - jul_mid) + t, sg.radius_d, sg.radius_rd);
    }
    const double radius = rec[1];
    *c = 1.0 / radius;
    return ((jd_ref - jul_mid) + t) / AB_DIVK(radius, 86400.0);
}

/* Position sums (file units) of one series from the record at `rec` ([_jul(MID), RADIUS, x0 y0 z0 x1 ...], staged or
 * in the image), four terms per trip: the packed copy holds zero coefficients behind the last term up to a multiple of
 * four, and adding 0 * T_p leaves a sum unchanged (a sum that starts from +0 is never -0).  The sums and their order
 * are those of ab_cheb3. */
__device__ __forceinline__ void abc_series_pos(const double* rec, const AbSpkSeg& sg, double jd_ref, double t, double* u) {
    double c;
    const double z = abc_record_z(rec, sg, jd_ref, t, &c);
    const int P4 = (sg.P + 3) & ~3;
    const double2* q = reinterpret_cast<const double2*>(rec + 2);
    const double z2 = 2.0 * z;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, T1, T2;
    {   /* x0 y0 | z0 x1 | y1 z1 | x2 y2 | z2 x3 | y3 z3 */
        const double2 c0 = q[0], c1 = q[1], c2 = q[2], c3 = q[3], c4 = q[4], c5 = q[5];
        a0 += c0.x * 1.0; a1 += c0.y * 1.0; a2 += c1.x * 1.0;
        a0 += c1.y * z; a1 += c2.x * z; a2 += c2.y * z;
        const double Ta = z2 * z - 1.0;
        a0 += c3.x * Ta; a1 += c3.y * Ta; a2 += c4.x * Ta;
        const double Tb = z2 * Ta - z;
        a0 += c4.y * Tb; a1 += c5.x * Tb; a2 += c5.y * Tb;
        T2 = Ta; T1 = Tb;
    }
#pragma unroll 1
    for (int p = 4; p < P4; p += 4) {
        const double2* qq = q + 3 * (p >> 1);
        const double2 c0 = qq[0], c1 = qq[1], c2 = qq[2], c3 = qq[3], c4 = qq[4], c5 = qq[5];
        const double Ta = z2 * T1 - T2;
        a0 += c0.x * Ta; a1 += c0.y * Ta; a2 += c1.x * Ta;
        const double Tb = z2 * Ta - T1;
        a0 += c1.y * Tb; a1 += c2.x * Tb; a2 += c2.y * Tb;
        const double Tc = z2 * Tb - Ta;
        a0 += c3.x * Tc; a1 += c3.y * Tc; a2 += c4.x * Tc;
        const double Td = z2 * Tc - Tb;
        a0 += c4.y * Td; a1 += c5.x * Td; a2 += c5.y * Td;
        T2 = Tc; T1 = Td;
    }
    u[0] = a0; u[1] = a1; u[2] = a2;
}

/* what becomes of the sums of series s: the Sun (evaluated with its velocity, ab_spk_planet<1>) gives the barycentric
 * offset of the asteroids, the EMB is kept for Earth and Moon, planets go into shared memory, asteroids into the
 * global table */
__device__ __forceinline__ void abc_series_finish(int s, const double* rec, const AbSpkSeg& sg, double jd_ref, AbcFillLane& X) {
    const AbEphem& E = c_abcE;
    if (s == ABC_S_SUN) {
        double c, u[3], uv[3] = {0, 0, 0}, uw[3] = {0, 0, 0};
        const double z = abc_record_z(rec, sg, jd_ref, X.t, &c);
        ab_cheb3<1, true, false>(rec + 2, sg.P, z, c, u, uv, uw);
        X.sx = ab_divc(u[0], E.u_d[0], E.u_rd[0]); X.sy = ab_divc(u[1], E.u_d[0], E.u_rd[0]); X.sz = ab_divc(u[2], E.u_d[0], E.u_rd[0]);
        X.g[ABC_GT_SVEL(0) * ABC_SLOTS] = ab_divc(uv[0], E.u_d[1], E.u_rd[1]);
        X.g[ABC_GT_SVEL(1) * ABC_SLOTS] = ab_divc(uv[1], E.u_d[1], E.u_rd[1]);
        X.g[ABC_GT_SVEL(2) * ABC_SLOTS] = ab_divc(uv[2], E.u_d[1], E.u_rd[1]);
        X.tb[ABC_E_POS(0, 0) * ABC_SLOTS] = X.sx; X.tb[ABC_E_POS(0, 1) * ABC_SLOTS] = X.sy; X.tb[ABC_E_POS(0, 2) * ABC_SLOTS] = X.sz;
        return;
    }
    double u[3];
    abc_series_pos(rec, sg, jd_ref, X.t, u);
    AbcSeriesRef R;
    R.img = nullptr; R.tg = nullptr;
    if (s == ABC_S_EMB) { R.kind = 0; R.idx = -1; }
    else if (s < ABC_S_AST0) { R.kind = 2; R.idx = s - ABC_S_PLANET0; }
    else { R.kind = 3; R.idx = s - ABC_S_AST0; }
    abc_fill_store(R, u, X.emb, X.tb, X.g, X.sx, X.sy, X.sz);
}

/* Locate the records of series `s` for every lane and request the warp's records into stage buffer `buf` (`bar`: the
 * shared-memory address of its mbarrier).  Returns true when every active lane reads from the stage buffer. */
__device__ __forceinline__ bool abc_fill_stage(AbcFillLane* L, int s, bool have_prev, bool prev_all, double jd_ref, unsigned actmask,
                                               double* buf, unsigned bar) {
    const AbSpkTarget& tg = c_abc_tg[s];
    const double* img = (s < ABC_S_AST0) ? c_abcE.spkp_img : c_abcE.spka_img;
    const int first = ab_ffs(actmask) - 1;
    const bool reuse = have_prev && tg.same_grid != 0;      /* same time grid as the series just promoted to "current" */
#ifdef AB_HOST_EMUL
    (void)bar;
    if (reuse) {
        for (int l = 0; l < 32; l++) {
            AbcFillLane& X = L[l];
            X.nseg = X.seg; X.noff = X.off;
            if (X.active && X.uni && (l & 7) == 0 && X.nfit > 0) {
                const AbSpkSeg& sg = tg.seg[X.seg];
                memcpy(buf + X.offr, img + (sg.one - 1) + (long long)X.lo * sg.R, (size_t)X.nfit * sg.R * 8);
            }
        }
        return prev_all;
    }
    int bb[32];
    for (int l = 0; l < 32; l++) {
        AbcFillLane& X = L[l];
        X.nseg = 0; X.noff = -1; bb[l] = 0;
        if (X.active) {
            X.nseg = ab_spk_segment(tg, jd_ref, X.t);
            bb[l] = ab_spk_record_index(tg.seg[X.nseg], jd_ref, X.t);
        }
    }
    const int nref = L[first].nseg;
    bool uni = true;
    for (int l = 0; l < 32; l++) if (L[l].active && L[l].nseg != nref) uni = false;
    bool all = true;
    for (int l = 0; l < 32; l++) { L[l].uni = uni ? 1 : 0; L[l].lo = 0; L[l].nfit = 0; L[l].offr = 0; }
    if (uni) {
        const AbSpkSeg& sg = tg.seg[nref];
        int off = 0;
        for (int k = 0; k < 4; k++) {
            const int f = 8 * k, e = 8 * k + 7;
            if (!L[f].active) continue;
            const int lo = bb[f] < bb[e] ? bb[f] : bb[e];
            const int span = (bb[f] < bb[e] ? bb[e] - bb[f] : bb[f] - bb[e]) + 1;
            int nfit = sg.stage_cap - off;
            if (nfit > span) nfit = span;
            if (nfit < 0) nfit = 0;
            if (nfit > 0) memcpy(buf + (size_t)off * sg.R, img + (sg.one - 1) + (long long)lo * sg.R, (size_t)nfit * sg.R * 8);
            for (int l = f; l <= e; l++) {
                L[l].lo = lo; L[l].nfit = nfit; L[l].offr = off * sg.R;
                if ((unsigned)(bb[l] - lo) < (unsigned)nfit) L[l].noff = (off + bb[l] - lo) * sg.R;
            }
            off += span;
        }
    }
    for (int l = 0; l < 32; l++) if (L[l].active && L[l].noff < 0) all = false;
    return all;
#else
    const int l = (int)(threadIdx.x & 31);
    AbcFillLane& X = L[0];
    if (reuse) {
        X.nseg = X.seg; X.noff = X.off;
        abc_fence_proxy_async();
        if (X.active && X.uni && (l & 7) == 0 && X.nfit > 0) {
            const AbSpkSeg& sg = tg.seg[X.seg];
            const unsigned bytes = (unsigned)(X.nfit * sg.R) * 8u;
            abc_mbar_expect_tx(bar, bytes);
            abc_bulk_g2s(abc_smem_addr(buf + X.offr), img + (sg.one - 1) + (long long)X.lo * sg.R, bytes, bar);
        }
        __syncwarp();
        if (l == 0) abc_mbar_arrive(bar);
        return prev_all;
    }
    int n = 0, b = 0;
    if (X.active) {
        n = ab_spk_segment(tg, jd_ref, X.t);
        b = ab_spk_record_index(tg.seg[n], jd_ref, X.t);
    }
    X.nseg = n; X.noff = -1;
    const int f = l & ~7, e = l | 7;
    const int bf = __shfl_sync(0xffffffffu, b, f), be = __shfl_sync(0xffffffffu, b, e);
    const int nref = __shfl_sync(0xffffffffu, n, first);
    const bool uni = __all_sync(0xffffffffu, !X.active || n == nref);
    const int lo = bf < be ? bf : be;
    const int span = X.active ? ((bf < be ? be - bf : bf - be) + 1) : 0;
    const int sp0 = __shfl_sync(0xffffffffu, span, 0), sp1 = __shfl_sync(0xffffffffu, span, 8), sp2 = __shfl_sync(0xffffffffu, span, 16);
    const int k = l >> 3;
    const int off = (k > 0 ? sp0 : 0) + (k > 1 ? sp1 : 0) + (k > 2 ? sp2 : 0);
    abc_fence_proxy_async();      /* the buffer was read through the generic proxy two series ago */
    X.uni = uni ? 1 : 0; X.lo = lo; X.nfit = 0; X.offr = 0;
    if (uni && X.active) {
        const AbSpkSeg& sg = tg.seg[nref];
        const int Rr = sg.R;
        int nfit = sg.stage_cap - off;
        if (nfit > span) nfit = span;
        if (nfit < 0) nfit = 0;
        X.nfit = nfit; X.offr = off * Rr;
        if ((unsigned)(b - lo) < (unsigned)nfit) X.noff = (off + b - lo) * Rr;
        if (l == f && nfit > 0) {
            const unsigned bytes = (unsigned)(nfit * Rr) * 8u;
            abc_mbar_expect_tx(bar, bytes);
            abc_bulk_g2s(abc_smem_addr(buf + off * Rr), img + (sg.one - 1) + (long long)lo * Rr, bytes, bar);
        }
    }
    __syncwarp();
    if (l == 0) abc_mbar_arrive(bar);      /* the phase ends when the requested bytes have landed (at once if none were) */
    return __all_sync(0xffffffffu, !X.active || X.noff >= 0);
#endif
}

__device__ __noinline__ void abc_fill_warp(ABC_CTXARG double* gt, int warp) {
    const AbEphem& E = c_abcE;
    ABC_SM_HERE(sm);
    const double jd_ref = E.jd_ref;
    const bool spk_regular = c_abc_regular != 0;
    const int s_end = ABC_S_AST0 + E.n_ast;
    const int s_begin = spk_regular ? 0 : ABC_S_AST0;
    AbcFillLane L[ABC_NL];
    unsigned actmask = 0u;
    ABC_LANES(l) {
        AbcFillLane& X = L[ABC_LI(l)];
        const int slot = 4 * warp + (l >> 3);
        const int node = l & 7;
        X.active = sm.flag(ABC_SMI_ACTIVE, slot);
#ifdef AB_HOST_EMUL
        if (X.active) actmask |= 1u << l;
#else
        actmask = __ballot_sync(0xffffffffu, X.active);
#endif
        const double t0 = sm.t0(slot);
        X.t = (node == 0) ? t0 : (t0 + sm.dt(slot) * c_h[node]);
        X.tb = sm.tab(node, slot);
        X.g = gt + node * ABC_GT_NODE + slot;
        X.sx = X.sy = X.sz = 0.0;
        X.emb[0] = X.emb[1] = X.emb[2] = 0.0;
        X.seg = X.nseg = 0; X.off = X.noff = -1;
        X.lo = X.nfit = X.offr = X.uni = 0;
        if (X.active && !spk_regular) {
            /* unusual layout (DE-binary planets, a kernel without Earth or EMB target): the one-time routines */
            int err = AB_OK;
            for (int b = 0; b < AB_NPLANETS; b++) {
                double GM, x[3], v[3], a[3];
                int flag;
                if (b == 0) {
                    flag = ab_planet<1>(E, 0, X.t, &GM, x, v, a);
                    X.g[ABC_GT_SVEL(0) * ABC_SLOTS] = v[0]; X.g[ABC_GT_SVEL(1) * ABC_SLOTS] = v[1]; X.g[ABC_GT_SVEL(2) * ABC_SLOTS] = v[2];
                    X.sx = x[0]; X.sy = x[1]; X.sz = x[2];
                } else {
                    flag = ab_planet<0>(E, b, X.t, &GM, x, v, a);
                }
                if (flag != AB_OK && err == AB_OK) err = flag;
                X.tb[ABC_E_POS(b, 0) * ABC_SLOTS] = x[0]; X.tb[ABC_E_POS(b, 1) * ABC_SLOTS] = x[1]; X.tb[ABC_E_POS(b, 2) * ABC_SLOTS] = x[2];
            }
            if (err != AB_OK) sm.flag(ABC_SMI_ERR, slot) = err;      /* the lanes of a slot may race: every value written is a valid code */
        }
    }
    if (actmask == 0u) return;       /* none of the warp's four slots steps in this attempt */
    if (s_begin < s_end) {
        double* stage = sm.d + ABC_ST_BASE + warp * ABC_ST_WARP;
#ifdef AB_HOST_EMUL
        const unsigned bar0 = 0u;
#else
        double* mb = sm.d + ABC_ST_MBAR + 3 * warp;
        const unsigned bar0 = abc_smem_addr(mb);
        unsigned phase = *reinterpret_cast<volatile unsigned*>(mb + 2);
#endif
        bool nall = abc_fill_stage(L, s_begin, false, false, jd_ref, actmask, stage, bar0);
#pragma unroll 1
        for (int s = s_begin; s < s_end; s++) {
            const int j = (s - s_begin) & 1;
            const bool all = nall;
            ABC_LANES(l) {
                AbcFillLane& X = L[ABC_LI(l)];
                X.seg = X.nseg; X.off = X.noff;
            }
            if (s + 1 < s_end) nall = abc_fill_stage(L, s + 1, true, all, jd_ref, actmask, stage + (j ^ 1) * ABC_ST_BUF, bar0 + 8u * (unsigned)(j ^ 1));
#ifndef AB_HOST_EMUL
            {   /* the records of series s have landed */
                const unsigned bar = bar0 + 8u * (unsigned)j, par = (phase >> j) & 1u;
                if (!abc_mbar_try_wait(bar, par)) {
                    const long long tw = clock64();
                    while (!abc_mbar_try_wait(bar, par))
                        if (clock64() - tw > (1LL << 34)) __trap();      /* ~9 s: a lost copy must not hang the grid */
                }
                phase ^= 1u << j;
            }
#endif
            const AbSpkTarget& tg = c_abc_tg[s];
            const double* sbuf = stage + j * ABC_ST_BUF;
            if (all) {      /* the usual case: every lane reads shared memory */
                ABC_LANES(l) {
                    AbcFillLane& X = L[ABC_LI(l)];
                    if (X.active) abc_series_finish(s, sbuf + X.off, tg.seg[X.seg], jd_ref, X);
                }
            } else {
                const double* img = (s < ABC_S_AST0) ? E.spkp_img : E.spka_img;
                ABC_LANES(l) {
                    AbcFillLane& X = L[ABC_LI(l)];
                    if (X.active) {
                        const AbSpkSeg& sg = tg.seg[X.seg];
                        const double* rec = (X.off >= 0) ? (sbuf + X.off) : ab_spk_record_ptr(img, sg, jd_ref, X.t);
                        abc_series_finish(s, rec, sg, jd_ref, X);
                    }
                }
            }
#ifndef AB_HOST_EMUL
            __syncwarp();      /* every lane is done with buffer j before series s + 2 is copied into it */
#endif
        }
#ifndef AB_HOST_EMUL
        if ((threadIdx.x & 31) == 0) *reinterpret_cast<volatile unsigned*>(mb + 2) = phase;
#endif
    }
    ABC_LANES(l) {
        AbcFillLane& X = L[ABC_LI(l)];
        if (X.active) abc_fill_eih(X.tb, X.g, X.sx, X.sy, X.sz);
    }
}
#else
__device__ __noinline__ void abc_fill_warp(ABC_CTXARG double* gt, int warp) {
    const AbEphem& E = c_abcE;
    const AbForceOpts& F = c_abcF;
    ABC_SM_HERE(sm);
    const double jd_ref = E.jd_ref;
    /* usual SPK layout: every body and the EMB have a target */
    bool spk_regular = (E.planets_source != AB_SRC_ASCII) && E.emb_index >= 0;
    for (int b = 0; b < AB_NPLANETS && spk_regular; b++) if (E.p_index[b] < 0) spk_regular = false;
    const int s_end = AB_NPLANETS + E.n_ast;
    ABC_LANES(l) {
        const int slot = 4 * warp + (l >> 3);
        const int node = l & 7;
        if (!sm.flag(ABC_SMI_ACTIVE, slot)) continue;
        const double t0 = sm.t0(slot);
        const double t = (node == 0) ? t0 : (t0 + sm.dt(slot) * c_h[node]);
        double* tb = sm.tab(node, slot);
        double* g = gt + node * ABC_GT_NODE + slot;
        int err = AB_OK;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        /* the Sun with its velocity; every planet when the layout is unusual (DE-binary planets, a kernel
         * without Earth or EMB target): the one-time routines */
        for (int b = 0; b < (spk_regular ? 1 : AB_NPLANETS); b++) {
            double GM, x[3], v[3], a[3];
            int flag;
            if (b == 0) {
                flag = ab_planet<1>(E, 0, t, &GM, x, v, a);
                g[ABC_GT_SVEL(0) * ABC_SLOTS] = v[0]; g[ABC_GT_SVEL(1) * ABC_SLOTS] = v[1]; g[ABC_GT_SVEL(2) * ABC_SLOTS] = v[2];
                sx = x[0]; sy = x[1]; sz = x[2];
            } else {
                flag = ab_planet<0>(E, b, t, &GM, x, v, a);
            }
            if (flag != AB_OK && err == AB_OK) err = flag;
            tb[ABC_E_POS(b, 0) * ABC_SLOTS] = x[0]; tb[ABC_E_POS(b, 1) * ABC_SLOTS] = x[1]; tb[ABC_E_POS(b, 2) * ABC_SLOTS] = x[2];
        }
        if (err != AB_OK) sm.flag(ABC_SMI_ERR, slot) = err;      /* the lanes of a slot may race: every value written is a valid code */
        double emb[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
        for (int s = spk_regular ? 0 : AB_NPLANETS; s < s_end; s += ABC_FILL_NS) {
            AbcSeriesRef R[ABC_FILL_NS];
#pragma unroll
            for (int j = 0; j < ABC_FILL_NS; j++) R[j] = abc_series_ref(s + j, s_end);
            double u[ABC_FILL_NS][3];
            abc_multi_eval<ABC_FILL_NS>(R, jd_ref, t, u);
#pragma unroll
            for (int j = 0; j < ABC_FILL_NS; j++) abc_fill_store(R[j], u[j], emb, tb, g, sx, sy, sz);
        }
        abc_fill_eih(tb, g, sx, sy, sz);
    }
}
#endif  /* ABC_FILL_TMA */

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
 * warps add them up, so their chains (difference, square root, division) are laid side by side.  ids[n] = 255: no body. */
template <int N>
__device__ __forceinline__ void abc_bodies(const AbEphem& E, const AbForceOpts& F, const AbcSmem& sm, const int* ids,
                                           const double (*c)[3], double px, double py, double pz, int slot) {
    const double xo = 0.0, yo = 0.0, zo = 0.0;
    double dx[N], dy[N], dz[N], r2[N], _r[N], r3[N], GM[N], prefac[N], q[N];
    bool ok = true;
#pragma unroll
    for (int n = 0; n < N; n++) {
        const int i = (ids[n] == 255) ? 0 : ids[n];
        GM[n] = E.gm[i];
        dx[n] = px + (xo - c[n][0]);
        dy[n] = py + (yo - c[n][1]);
        dz[n] = pz + (zo - c[n][2]);
        r2[n] = dx[n] * dx[n] + dy[n] * dy[n] + dz[n] * dz[n];
        ok = ok && ab_nb_ok(r2[n]) && ab_nb_ok(GM[n]);
    }
    /* square roots and quotients of the group side by side (fp_device.cuh); operands outside the branch-free range
     * (a particle inside a body, a zero mass) take the built-in operators */
#if !AB_STRICT && !defined(AB_HOST_EMUL)
    /* fast math: 1/r from a refined reciprocal square root, 1/r^3 as its cube (relative error ~3e-16 per term) */
#pragma unroll
    for (int n = 0; n < N; n++) {
        const double ir = ab_rsqrt_fast(r2[n]);
        _r[n] = r2[n] * ir;
        r3[n] = _r[n] * r2[n];
        q[n] = GM[n] * ir;
        prefac[n] = q[n] * (ir * ir);
    }
#else
#pragma unroll
    for (int n = 0; n < N; n++) _r[n] = ab_sqrt_nb(r2[n]);
#pragma unroll
    for (int n = 0; n < N; n++) { r3[n] = _r[n] * _r[n] * _r[n]; ok = ok && ab_nb_ok(r3[n]); }
#pragma unroll
    for (int n = 0; n < N; n++) { prefac[n] = ab_div_nb(GM[n], r3[n]); q[n] = ab_div_nb(GM[n], _r[n]); }
#endif
    if (!ok) {
#pragma unroll
        for (int n = 0; n < N; n++) {
            _r[n] = sqrt(r2[n]);
            prefac[n] = GM[n] / (_r[n] * _r[n] * _r[n]);
            q[n] = GM[n] / _r[n];
        }
    }
    const bool eih = (F.forces & 0x40) != 0;
#pragma unroll
    for (int n = 0; n < N; n++) {
        const int i = ids[n];
        if (i == 255) continue;
        const double p0 = prefac[n] * dx[n], p1 = prefac[n] * dy[n], p2 = prefac[n] * dz[n];
        if (abc_body_on(i, F.forces)) { sm.prod(i, 0, slot) = p0; sm.prod(i, 1, slot) = p1; sm.prod(i, 2, slot) = p2; }
        if (i < AB_NPLANETS && eih) sm.q(i, slot) = q[n];
    }
}

#ifndef ABC_BN
#define ABC_BN 2      /* bodies side by side */
#endif
__device__ __forceinline__ void abc_ast_fetch(const AbcWorkerPlan& wp, const double* g, int k0, double (*c)[3]) {
#pragma unroll
    for (int j = 0; j < ABC_BN; j++) {
        const int m = (k0 + j < wp.nast) ? wp.ast[k0 + j] : wp.ast[0];
#pragma unroll
        for (int q = 0; q < 3; q++) c[j][q] = (k0 + j < wp.nast) ? __ldcg(g + ABC_GT_AST(m, q) * ABC_SLOTS) : 0.0;
    }
}

/* planets first, then asteroids, through ONE copy of abc_bodies<ABC_BN>: the hot code of a node round has to stay
 * inside the instruction cache */
__device__ __forceinline__ void abc_task_all_bodies(const AbEphem& E, const AbForceOpts& F, const AbcSmem& sm, const double* tb,
                                                    const double* g, const AbcWorkerPlan& wp, int slot, double (*first)[3]) {
    const double px = sm.xv(0, slot), py = sm.xv(1, slot), pz = sm.xv(2, slot);
    double nxt[ABC_BN][3];
#pragma unroll
    for (int j = 0; j < ABC_BN; j++) { nxt[j][0] = first[j][0]; nxt[j][1] = first[j][1]; nxt[j][2] = first[j][2]; }
    const int np = (wp.nbody + ABC_BN - 1) / ABC_BN * ABC_BN;           /* planets in groups of ABC_BN, then asteroids */
    const int total = np + wp.nast;
#pragma unroll 1
    for (int n = 0; n < total; n += ABC_BN) {
        int ids[ABC_BN];
        double c[ABC_BN][3];
        if (n < np) {
#pragma unroll
            for (int j = 0; j < ABC_BN; j++) {
                ids[j] = (n + j < wp.nbody) ? wp.body[n + j] : 255;
                const int i = (ids[j] == 255) ? 0 : ids[j];
                c[j][0] = tb[ABC_E_POS(i, 0) * ABC_SLOTS]; c[j][1] = tb[ABC_E_POS(i, 1) * ABC_SLOTS]; c[j][2] = tb[ABC_E_POS(i, 2) * ABC_SLOTS];
            }
        } else {
            const int k0 = n - np;
#pragma unroll
            for (int j = 0; j < ABC_BN; j++) {
                ids[j] = (k0 + j < wp.nast) ? AB_NPLANETS + wp.ast[k0 + j] : 255;
                c[j][0] = nxt[j][0]; c[j][1] = nxt[j][1]; c[j][2] = nxt[j][2];
            }
            if (k0 + ABC_BN < wp.nast) abc_ast_fetch(wp, g, k0 + ABC_BN, nxt);
        }
        abc_bodies<ABC_BN>(E, F, sm, ids, c, px, py, pz, slot);
    }
}

/* EIH source block of the Sun for the real particle (reference src/forces.c:1319-1501 with j = 0), everything
 * except the potential sum over the planets, which the component warps add in order.  Same expressions as
 * ab_force_eih. */
__device__ void abc_task_eih_source(const AbEphem& E, const AbcSmem& sm, const double* tb, const double* ev, int slot) {
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
        const double vxj = ev[0], vyj = ev[1], vzj = ev[2];         /* Sun's velocity, EIH sums: fetched from the global table by the caller */

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

        double term1 = ev[3];
        const double axj = ev[4], ayj = ev[5], azj = ev[6];
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

__device__ void abc_run_task(const AbEphem& E, const AbForceOpts& F, const AbcSmem& sm, const double* tb, const double* g,
                             const double* ev, int slot, int kind) {
    if (kind == ABC_T_EIHSRC) { abc_task_eih_source(E, sm, tb, ev, slot); return; }
    AbSysT<1> S;
    abc_sys_from_slot(sm, slot, S);
    const AbcTabView B(E.gm, tb, g);
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
        double qk[AB_NPLANETS];
#pragma unroll
        for (int k = 0; k < AB_NPLANETS; k++) qk[k] = sm.q(k, slot);
#pragma unroll
        for (int k = 0; k < AB_NPLANETS; k++) term0_sum += qk[k];
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
        /* the 27 terms of the direct sum in the reference's order; the values are fetched from shared memory EIGHT AT A
         * TIME before they are subtracted, so that the chain of dependent subtractions does not wait for one load per
         * term (a missing asteroid is +0.0: a - 0.0 == a exactly, signed zeros included) */
        const int ast_num = E.n_ast;
        if (F.forces & 0x04) {
#pragma unroll 1
            for (int k0 = 0; k0 < ast_num; k0 += 8) {
                double pa[8];
#pragma unroll
                for (int j = 0; j < 8; j++) pa[j] = (k0 + j < ast_num) ? sm.prod(AB_NPLANETS + k0 + j, c, slot) : 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++) a -= pa[j];
            }
        }
        if (F.forces & 0x02) {
            const double p10 = sm.prod(10, c, slot), p4 = sm.prod(4, c, slot), p5 = sm.prod(5, c, slot), p1 = sm.prod(1, c, slot);
            const double p9 = sm.prod(9, c, slot), p8 = sm.prod(8, c, slot), p3 = sm.prod(3, c, slot), p2 = sm.prod(2, c, slot);
            const double p7 = sm.prod(7, c, slot), p6 = sm.prod(6, c, slot);
            a -= p10; a -= p4; a -= p5; a -= p1; a -= p9; a -= p8; a -= p3; a -= p2; a -= p7; a -= p6;
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
