/*
 * ephem_device.cuh -- Chebyshev ephemeris evaluation on the device.
 *
 * The coefficient tables are byte-exact images of the .bsp / .44x files in HBM
 * (they are a few MB, i.e. L2-resident); the small descriptors travel in the
 * __grid_constant__ AbEphem kernel parameter.  Arithmetic follows the reference
 * operation by operation so that, compiled without FMA contraction, a body state
 * has the same bits as the C build produces:
 *   SPK planets      reference src/spk.c:492-547 (target_pos), :550-610 (EMB shift, units)
 *   SPK asteroids    reference src/spk.c:405-481 (position only, literal 149597870.7)
 *   DE binary        reference src/ascii_ephem.c:27-65 (work), :275-384 (calc)
 *   body dispatch    reference src/forces.c:175-263 (assist_all_ephem: asteroid + Sun shift)
 * Differences in HOW (not in what is computed):
 *   - the T/S/U recurrences are carried in registers instead of 32-entry arrays, and
 *     velocity / acceleration sums are only formed when a caller needs them;
 *   - the per-segment trailer values (INIT, INTLEN, RSIZE, N) and every invariant
 *     divisor are prepared once at upload time, so an evaluation needs one dependent
 *     load (the record's MID) before the coefficient stream, and no true division
 *     (fp_device.cuh: exact quotients from precomputed reciprocals).
 */
#ifndef AB_EPHEM_DEVICE_CUH
#define AB_EPHEM_DEVICE_CUH

#include "device_types.h"
#include "fp_device.cuh"

#define AB_OK 0
#define AB_ERR_EPHEM_FILE 1
#define AB_ERR_AST_FILE 2
#define AB_ERR_NAST 3
#define AB_ERR_NEPHEM 4
#define AB_ERR_COVERAGE 5

namespace AB_NS {

__device__ __forceinline__ double ab_jul(double eph) { return 2451545.0 + AB_DIVK(eph, 86400.0); }

/* Chebyshev sums for one record: 3 components, P coefficients each, argument z,
 * derivative scale c.  LEVEL 0: position, 1: +velocity, 2: +acceleration. */
/* PACKED: the device copy of an SPK kernel holds the coefficients of a record as [p][x y z] (gpu_api.cu,
 * upload_packed_spk); DE-binary images keep the file's [x y z][p]. */
/* LDG: the coefficients are read-only global memory (non-coherent loads); false: any address space (the staged
 * copies of pp_coop_kernel live in shared memory). */
template <int LEVEL, bool PACKED, bool LDG = true>
__device__ __forceinline__ void ab_cheb3(const double* __restrict__ cf, int P, double z, double c,
                                         double u[3], double v[3], double w[3]) {
    double u0 = 0.0, u1 = 0.0, u2 = 0.0;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    double Tm1 = 0.0, Tm2 = 0.0, Sm1 = 0.0, Sm2 = 0.0, Um1 = 0.0, Um2 = 0.0;
    const int stride = PACKED ? 3 : 1;
    const double* __restrict__ cx = cf;
    const double* __restrict__ cy = cf + (PACKED ? 1 : P);
    const double* __restrict__ cz = cf + (PACKED ? 2 : 2 * P);
    for (int p = 0; p < P; p++) {
        double T, S = 0.0, U = 0.0;
        if (p == 0) { T = 1.0; S = 0.0; U = 0.0; }
        else if (p == 1) { T = z; S = 1.0; U = 0.0; }
        else {
            T = 2.0 * z * Tm1 - Tm2;
            if (LEVEL >= 1) S = 2.0 * z * Sm1 + 2.0 * Tm1 - Sm2;
            if (LEVEL >= 2) U = (p == 2) ? 4.0 : (2.0 * z * Um1 + 4.0 * Sm1 - Um2);
        }
        const double ax = LDG ? __ldg(cx + stride * p) : cx[stride * p];
        const double ay = LDG ? __ldg(cy + stride * p) : cy[stride * p];
        const double az = LDG ? __ldg(cz + stride * p) : cz[stride * p];
        u0 += ax * T; u1 += ay * T; u2 += az * T;
        if (LEVEL >= 1) { v0 += ax * S * c; v1 += ay * S * c; v2 += az * S * c; }
        if (LEVEL >= 2) { w0 += ax * U * c * c; w1 += ay * U * c * c; w2 += az * U * c * c; }
        Tm2 = Tm1; Tm1 = T;
        if (LEVEL >= 1) { Sm2 = Sm1; Sm1 = S; }
        if (LEVEL >= 2) { Um2 = Um1; Um1 = U; }
    }
    u[0] = u0; u[1] = u1; u[2] = u2;
    if (LEVEL >= 1) { v[0] = v0; v[1] = v1; v[2] = v2; }
    if (LEVEL >= 2) { w[0] = w0; w[1] = w1; w[2] = w2; }
}

/* Segment of `tg` that holds jd_ref + t (reference src/spk.c:501-502). */
__device__ __forceinline__ int ab_spk_segment(const AbSpkTarget& tg, double jd_ref, double t) {
    int n = (int)ab_divc(jd_ref + t - tg.beg, tg.res, tg.res_rd);
    if (n > tg.nseg - 1) n = tg.nseg - 1;       /* jd == end: the reference indexes one past; stay in the last segment */
    if (n < 0) n = 0;
    return n;
}

/* Locate the type-2 record of segment `sg` that holds jd_ref + t (reference src/spk.c:503-517). */
__device__ __forceinline__ int ab_spk_record_index(const AbSpkSeg& sg, double jd_ref, double t) {
    int b = (int)ab_divc((jd_ref - sg.jul_init) + t, sg.intlen_d, sg.intlen_rd);
    if (b > sg.nrec - 1) b = sg.nrec - 1;
    if (b < 0) b = 0;
    return b;
}
__device__ __forceinline__ const double* ab_spk_record_ptr(const double* __restrict__ img, const AbSpkSeg& sg, double jd_ref, double t) {
    /* the record in the packed copy: [_jul(MID), RADIUS, coefficients] */
    return img + (sg.one - 1) + (long long)ab_spk_record_index(sg, jd_ref, t) * sg.R;
}
__device__ __forceinline__ const double* ab_spk_record_in(const double* __restrict__ img, const AbSpkSeg& sg,
                                                          double jd_ref, double t, double* z, double* c) {
    const double* rec = ab_spk_record_ptr(img, sg, jd_ref, t);
    const double jul_mid = __ldg(rec);          /* the packed copy holds _jul(MID) = 2451545.0 + MID / 86400.0, formed at upload */
    if (sg.uniform) {
        *z = ab_divc((jd_ref - jul_mid) + t, sg.radius_d, sg.radius_rd);
        *c = sg.radius_inv;
    } else {
        const double radius = __ldg(rec + 1);
        *z = ((jd_ref - jul_mid) + t) / AB_DIVK(radius, 86400.0);
        *c = 1.0 / radius;
    }
    return rec + 2;
}

__device__ __forceinline__ const double* ab_spk_record(const double* __restrict__ img, const AbSpkTarget& tg,
                                                       double jd_ref, double t, int* P, double* z, double* c) {
    const AbSpkSeg& sg = tg.seg[ab_spk_segment(tg, jd_ref, t)];
    *P = sg.P;
    return ab_spk_record_in(img, sg, jd_ref, t, z, c);
}

template <int LEVEL>
__device__ __forceinline__ void ab_spk_target_pos(const double* __restrict__ img, const AbSpkTarget& tg,
                                                  double jd_ref, double t, double u[3], double v[3], double w[3]) {
    int P; double z, c;
    const double* cf = ab_spk_record(img, tg, jd_ref, t, &P, &z, &c);
    ab_cheb3<LEVEL, true>(cf, P, z, c, u, v, w);
}

/* Planet (body < 11) from an SPK kernel, reference src/spk.c:550-610, 632-693. */
template <int LEVEL>
__device__ int ab_spk_planet(const AbEphem& E, int body, double t, double* GM, double x[3], double v[3], double a[3]) {
    const double jd_ref = E.jd_ref;
    double u[3], uv[3] = {0, 0, 0}, uw[3] = {0, 0, 0};
    const int idx = E.p_index[body];
    if (idx >= 0) {
        const AbSpkTarget& tg = E.p_tgt[idx];
        if (jd_ref + t < tg.beg || jd_ref + t > tg.end) return AB_ERR_COVERAGE;
        *GM = tg.mass;
        ab_spk_target_pos<LEVEL>(E.spkp_img, tg, jd_ref, t, u, uv, uw);
        if (body == 3 || body == 4) {           /* NAIF 399 / 301 are given relative to the EMB */
            if (E.emb_index < 0) return AB_ERR_NEPHEM;
            double e[3], ev[3] = {0, 0, 0}, ew[3] = {0, 0, 0};
            ab_spk_target_pos<LEVEL>(E.spkp_img, E.p_tgt[E.emb_index], jd_ref, t, e, ev, ew);
            for (int i = 0; i < 3; i++) { u[i] += e[i]; uv[i] += ev[i]; uw[i] += ew[i]; }
        }
    } else if (body == 3 && E.emb_index >= 0 && E.p_index[4] >= 0) {
        /* Earth from EMB and Moon when 399 is absent (reference src/spk.c:667-691; GM = 0 there) */
        double e[3], ev[3] = {0, 0, 0}, ew[3] = {0, 0, 0}, m[3], mv[3] = {0, 0, 0}, mw[3] = {0, 0, 0};
        ab_spk_target_pos<LEVEL>(E.spkp_img, E.p_tgt[E.emb_index], jd_ref, t, e, ev, ew);
        ab_spk_target_pos<LEVEL>(E.spkp_img, E.p_tgt[E.p_index[4]], jd_ref, t, m, mv, mw);
        const double frac = 1.0 / (1.0 + E.EMRAT);
        for (int i = 0; i < 3; i++) {
            u[i] = -frac * m[i] + e[i]; uv[i] = -frac * mv[i] + ev[i]; uw[i] = -frac * mw[i] + ew[i];
        }
        *GM = 0.0;
    } else {
        return AB_ERR_NEPHEM;
    }
    /* km, km/s, km/s^2 -> AU, AU/day, AU/day^2: divisors au, au/86400, au/86400^2 */
    for (int i = 0; i < 3; i++) {
        x[i] = ab_divc(u[i], E.u_d[0], E.u_rd[0]);
        if (LEVEL >= 1) v[i] = ab_divc(uv[i], E.u_d[1], E.u_rd[1]);
        if (LEVEL >= 2) a[i] = ab_divc(uw[i], E.u_d[2], E.u_rd[2]);
    }
    return AB_OK;
}

/* One column of a DE-binary record, reference src/ascii_ephem.c:27-65. */
template <int LEVEL>
__device__ __forceinline__ void ab_ascii_work(const double* __restrict__ Pcol, int ncf, int niv, double c,
                                              double t0, double u[3], double v[3], double w[3]) {
    const double tt = t0 * (double)niv;
    int b = (int)tt;
    /* t0 == 1 only at the very end of the file's coverage (jd == end): the last sub-interval at z = +1.  The reference
     * indexes one sub-interval past the column there (src/ascii_ephem.c:37-41) and returns whatever follows. */
    if (b > niv - 1) b = niv - 1;
    const double frac = tt - (double)b;            /* == fmod(tt, 1.0) for 0 <= tt < niv, exactly */
    const double z = 2.0 * frac - 1.0;
    ab_cheb3<LEVEL, false>(Pcol + ncf * (b * 3), ncf, z, c, u, v, w);
}

/* Planet (body < 11) from a DE-binary file, reference src/ascii_ephem.c:275-384. */
template <int LEVEL>
__device__ int ab_ascii_planet(const AbEphem& E, int body, double t, double* GM, double x[3], double v[3], double a[3]) {
    const double jd_ref = E.jd_ref;
    *GM = E.a_mass[body];
    if (jd_ref + t < E.a_beg || jd_ref + t > E.a_end) return AB_ERR_COVERAGE;
    long long blk = (long long)(unsigned int)ab_divc(jd_ref + t - E.a_beg, E.a_inc, E.a_inc_rd);
    if (blk > E.a_nrec - 1) blk = E.a_nrec - 1;    /* jd == end */
    const double* z = E.ascii_img + (blk + 2) * E.a_rec_words;
    const double tr = ab_divc((jd_ref - E.a_beg - (double)blk * E.a_inc) + t, E.a_inc, E.a_inc_rd);
    /* ASCII_* column of each ASSIST body */
    const int col = (body == 0) ? 10 : (body <= 3 ? body - 1 : (body == 4 ? 2 : body - 2));
    double u[3], uv[3] = {0, 0, 0}, uw[3] = {0, 0, 0};
    ab_ascii_work<LEVEL>(z + E.a_off[col], E.a_ncf[col], E.a_niv[col], E.a_c[col], tr, u, uv, uw);
    if (body == 3 || body == 4) {
        double l[3], lv[3] = {0, 0, 0}, lw[3] = {0, 0, 0};
        ab_ascii_work<LEVEL>(z + E.a_off[9], E.a_ncf[9], E.a_niv[9], E.a_c[9], tr, l, lv, lw);
        const double f = (body == 3) ? E.a_f_earth : E.a_f_moon;
        for (int i = 0; i < 3; i++) { u[i] += l[i] * f; uv[i] += lv[i] * f; uw[i] += lw[i] * f; }
    }
    for (int i = 0; i < 3; i++) {
        x[i] = ab_divc(u[i], E.u_d[0], E.u_rd[0]);
        if (LEVEL >= 1) v[i] = ab_divc(uv[i], E.u_d[1], E.u_rd[1]);
        if (LEVEL >= 2) a[i] = ab_divc(uw[i], E.u_d[2], E.u_rd[2]);
    }
    return AB_OK;
}

template <int LEVEL>
__device__ __forceinline__ int ab_planet(const AbEphem& E, int body, double t, double* GM, double x[3], double v[3], double a[3]) {
    if (E.planets_source == AB_SRC_ASCII) return ab_ascii_planet<LEVEL>(E, body, t, GM, x, v, a);
    return ab_spk_planet<LEVEL>(E, body, t, GM, x, v, a);
}

/* Heliocentric asteroid position in AU, reference src/spk.c:405-481. */
__device__ __forceinline__ int ab_asteroid(const AbEphem& E, int m, double t, double* GM, double x[3]) {
    if (E.spka_img == nullptr) return AB_ERR_AST_FILE;
    if (m < 0 || m >= E.n_ast + E.n_ast_x) return AB_ERR_NAST;
    const AbSpkTarget& tg = E.a_tgt[m];
    const double jd_ref = E.jd_ref;
    if (jd_ref + t < tg.beg || jd_ref + t > tg.end) return AB_ERR_COVERAGE;
    *GM = tg.mass;
    int P; double z, c;
    const double* cf = ab_spk_record(E.spka_img, tg, jd_ref, t, &P, &z, &c);
    double u[3], dv[3], dw[3];
    ab_cheb3<0, true>(cf, P, z, c, u, dv, dw);
    x[0] = AB_DIVK(u[0], 149597870.7); x[1] = AB_DIVK(u[1], 149597870.7); x[2] = AB_DIVK(u[2], 149597870.7);
    return AB_OK;
}

/* An asteroid beyond the body table (sb441-n373), barycentric, for the direct term (reference src/forces.c:208-224 via
 * assist_all_ephem): evaluated where it is needed.  Out of line: the direct loop of the n16 runs stays as it is. */
__device__ __noinline__ void ab_extra_asteroid(const AbEphem& E, int m, double t, double sx, double sy, double sz, double* GM, double* c) {
    double x[3] = {0.0, 0.0, 0.0};
    *GM = 0.0;
    ab_asteroid(E, m, t, GM, x);        /* coverage has been checked for the whole step (ab_extra_coverage) */
    c[0] = x[0] + sx; c[1] = x[1] + sy; c[2] = x[2] + sz;
}
__device__ __forceinline__ int ab_extra_coverage(const AbEphem& E, double t) {
    const double jd = E.jd_ref + t;
    for (int m = E.n_ast; m < E.n_ast + E.n_ast_x; m++)
        if (jd < E.a_tgt[m].beg || jd > E.a_tgt[m].end) return AB_ERR_COVERAGE;
    return AB_OK;
}

/* All body states one force evaluation needs at time t.  The particle-independent
 * sums of the EIH term (reference src/forces.c:1400-1436 and :1733-1772: GM_k/r_jk
 * and the Newtonian acceleration a_j of each source) are formed here once per time
 * instead of once per particle; the operations and their order are unchanged. */
__device__ void ab_body_states(const AbEphem& E, const AbForceOpts& F, double t, AbBodies& B) {
    int status = AB_OK;
    const int ns = F.gr_eih_sources;
    const bool need_eih = (F.forces & 0x40) != 0;
    for (int i = 0; i < AB_NPLANETS; i++) {
        double acc[3];
        int flag;
        const bool need_vel = (need_eih && i < ns) || i == 0 || (F.geocentric && i == 3);
        if (F.geocentric && i == 3) {
            flag = ab_planet<2>(E, i, t, &B.gm[i], B.pos[i], B.vel[i], acc);
            B.earth_acc[0] = acc[0]; B.earth_acc[1] = acc[1]; B.earth_acc[2] = acc[2];
        } else if (need_vel) {
            flag = ab_planet<1>(E, i, t, &B.gm[i], B.pos[i], B.vel[i], acc);
        } else {
            flag = ab_planet<0>(E, i, t, &B.gm[i], B.pos[i], B.vel[i], acc);
        }
        if (flag != AB_OK && status == AB_OK) status = flag;
    }
    /* asteroids are always needed: the variational direct term ignores the force mask (src/forces.c:359) */
    for (int m = 0; m < E.n_ast; m++) {
        double x[3];
        int flag = ab_asteroid(E, m, t, &B.gm[AB_NPLANETS + m], x);
        if (flag != AB_OK && status == AB_OK) status = flag;
        /* heliocentric -> barycentric, reference src/forces.c:213-219 */
        B.pos[AB_NPLANETS + m][0] = x[0] + B.pos[0][0];
        B.pos[AB_NPLANETS + m][1] = x[1] + B.pos[0][1];
        B.pos[AB_NPLANETS + m][2] = x[2] + B.pos[0][2];
    }
    if (E.n_ast_x > 0 && status == AB_OK) status = ab_extra_coverage(E, t);
    B.t = t;
    if (need_eih && status == AB_OK) {
        for (int j = 0; j < ns; j++) {
            double term1 = 0.0;
            double arx = 0.0, ary = 0.0, arz = 0.0;
            double avx = 0.0, avy = 0.0, avz = 0.0;
            for (int k = 0; k < AB_NPLANETS; k++) {
                if (k == j) continue;
                const double GMk = B.gm[k];
                const double dxjk = B.pos[j][0] - B.pos[k][0];
                const double dyjk = B.pos[j][1] - B.pos[k][1];
                const double dzjk = B.pos[j][2] - B.pos[k][2];
                const double rjk2 = dxjk * dxjk + dyjk * dyjk + dzjk * dzjk;
                const double _rjk = sqrt(rjk2);
                term1 += GMk / _rjk;
                const double fac = GMk / (rjk2 * _rjk);
                arx -= fac * dxjk; ary -= fac * dyjk; arz -= fac * dzjk;
                const AbDivisor r3(_rjk * _rjk * _rjk);
                avx -= r3(GMk * dxjk);
                avy -= r3(GMk * dyjk);
                avz -= r3(GMk * dzjk);
            }
            B.eih_term1[j] = term1;
            B.eih_ar[j][0] = arx; B.eih_ar[j][1] = ary; B.eih_ar[j][2] = arz;
            B.eih_av[j][0] = avx; B.eih_av[j][1] = avy; B.eih_av[j][2] = avz;
        }
    }
    B.status = status;
}

/* ------------------------------------------------------------------------------------------
 * Node tables: the body states of one IAS15 step, evaluated once.
 *
 * Within a step the force routine is called at the same 8 times (start of step + 7 Gauss-Radau
 * nodes) in every predictor-corrector sweep.  The reference memoises them in a 7-slot cache
 * keyed on the time (src/forces.c:180-197, 227-261); here a thread fills its 8 node tables in
 * one go.  The Chebyshev recurrences of the 8 times are independent, so they are run side by
 * side (8-way instruction-level parallelism on a latency-bound FP64 chain), and a coefficient
 * record shared by several nodes is read through L1 once.  Values are the same as those of the
 * one-time routines above: each sum sees the same operands in the same order.
 * ------------------------------------------------------------------------------------------ */
#define AB_NT 8

/* position sums (km) of one series at AB_NT arguments; cf[k] points at the first coefficient of the record.
 * PACKED records ([p][x y z], 16-byte aligned) are read two terms at a time with three 16-byte loads. */
template <bool PACKED>
__device__ __forceinline__ void ab_cheb_pos_multi(const double* const* cf, int P, const double* z, double (*u)[3]) {
    double T1[AB_NT], T2[AB_NT], a0[AB_NT], a1[AB_NT], a2[AB_NT];
    if (PACKED) {
#pragma unroll
        for (int k = 0; k < AB_NT; k++) {
            /* p = 0 (T = 1) and p = 1 (T = z) */
            const double2* q = reinterpret_cast<const double2*>(cf[k]);
            const double2 A = __ldg(q), B = __ldg(q + 1), C = __ldg(q + 2);      /* x0 y0 | z0 x1 | y1 z1 */
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            s0 += A.x * 1.0; s1 += A.y * 1.0; s2 += B.x * 1.0;
            s0 += B.y * z[k]; s1 += C.x * z[k]; s2 += C.y * z[k];
            a0[k] = s0; a1[k] = s1; a2[k] = s2;
            T2[k] = 1.0; T1[k] = z[k];
        }
        int p = 2;
#pragma unroll 1
        for (; p + 1 < P; p += 2) {
#pragma unroll
            for (int k = 0; k < AB_NT; k++) {
                const double2* q = reinterpret_cast<const double2*>(cf[k]) + 3 * (p >> 1);
                const double2 A = __ldg(q), B = __ldg(q + 1), C = __ldg(q + 2);  /* xp yp | zp xp+1 | yp+1 zp+1 */
                const double Ta = 2.0 * z[k] * T1[k] - T2[k];
                a0[k] += A.x * Ta; a1[k] += A.y * Ta; a2[k] += B.x * Ta;
                const double Tb = 2.0 * z[k] * Ta - T1[k];
                a0[k] += B.y * Tb; a1[k] += C.x * Tb; a2[k] += C.y * Tb;
                T2[k] = Ta; T1[k] = Tb;
            }
        }
        if (p < P) {      /* odd number of terms: the second half of the last 16-byte pair is padding */
#pragma unroll
            for (int k = 0; k < AB_NT; k++) {
                const double2* q = reinterpret_cast<const double2*>(cf[k]) + 3 * (p >> 1);
                const double2 A = __ldg(q), B = __ldg(q + 1);
                const double Ta = 2.0 * z[k] * T1[k] - T2[k];
                a0[k] += A.x * Ta; a1[k] += A.y * Ta; a2[k] += B.x * Ta;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < AB_NT; k++) {
            /* p = 0 (T = 1) and p = 1 (T = z) */
            const double* c = cf[k];
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            s0 += __ldg(c) * 1.0; s1 += __ldg(c + P) * 1.0; s2 += __ldg(c + 2 * P) * 1.0;
            s0 += __ldg(c + 1) * z[k]; s1 += __ldg(c + P + 1) * z[k]; s2 += __ldg(c + 2 * P + 1) * z[k];
            a0[k] = s0; a1[k] = s1; a2[k] = s2;
            T2[k] = 1.0; T1[k] = z[k];
        }
#pragma unroll 1
        for (int p = 2; p < P; p++) {
#pragma unroll
            for (int k = 0; k < AB_NT; k++) {
                const double T = 2.0 * z[k] * T1[k] - T2[k];
                const double* c = cf[k] + p;
                a0[k] += __ldg(c) * T; a1[k] += __ldg(c + P) * T; a2[k] += __ldg(c + 2 * P) * T;
                T2[k] = T1[k]; T1[k] = T;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < AB_NT; k++) { u[k][0] = a0[k]; u[k][1] = a1[k]; u[k][2] = a2[k]; }
}

/* one SPK target at AB_NT times (positions in km) */
__device__ __noinline__ void ab_spk_pos_multi(const double* __restrict__ img, const AbSpkTarget& tg, double jd_ref,
                                              const double* t, double (*u)[3]) {
    const double* cf[AB_NT];
    double z[AB_NT];
    int P0 = 0;
    bool same = true;
    /* the node times run monotonically from t[0] to t[AB_NT - 1]: when the two ends fall into the same segment
     * (segments are years long) so do all of them, and its descriptor is looked up once */
    const int n_first = ab_spk_segment(tg, jd_ref, t[0]);
    if (n_first == ab_spk_segment(tg, jd_ref, t[AB_NT - 1])) {
        const AbSpkSeg& sg = tg.seg[n_first];
        P0 = sg.P;
#pragma unroll
        for (int k = 0; k < AB_NT; k++) {
            double c;
            cf[k] = ab_spk_record_in(img, sg, jd_ref, t[k], &z[k], &c);
        }
    } else {
#pragma unroll
        for (int k = 0; k < AB_NT; k++) {
            int P; double c;
            cf[k] = ab_spk_record(img, tg, jd_ref, t[k], &P, &z[k], &c);
            if (k == 0) P0 = P; else same = same && (P == P0);
        }
    }
    if (same) {
        ab_cheb_pos_multi<true>(cf, P0, z, u);
    } else {   /* nodes in segments with different record sizes: one at a time */
        for (int k = 0; k < AB_NT; k++) {
            double dv[3], dw[3];
            ab_spk_target_pos<0>(img, tg, jd_ref, t[k], u[k], dv, dw);
        }
    }
}

/* one DE-binary column at AB_NT times (positions in km) */
__device__ __noinline__ void ab_ascii_pos_multi(const AbEphem& E, int col, const double* t, double (*u)[3]) {
    const double* cf[AB_NT];
    double z[AB_NT];
    const int ncf = E.a_ncf[col], niv = E.a_niv[col];
#pragma unroll
    for (int k = 0; k < AB_NT; k++) {
        long long blk = (long long)(unsigned int)ab_divc(E.jd_ref + t[k] - E.a_beg, E.a_inc, E.a_inc_rd);
        if (blk > E.a_nrec - 1) blk = E.a_nrec - 1;
        const double* rec = E.ascii_img + (blk + 2) * E.a_rec_words;
        const double tr = ab_divc((E.jd_ref - E.a_beg - (double)blk * E.a_inc) + t[k], E.a_inc, E.a_inc_rd);
        const double tt = tr * (double)niv;
        int b = (int)tt;
        if (b > niv - 1) b = niv - 1;              /* jd == end of coverage, see ab_ascii_work */
        z[k] = 2.0 * (tt - (double)b) - 1.0;
        cf[k] = rec + E.a_off[col] + ncf * (b * 3);
    }
    ab_cheb_pos_multi<false>(cf, ncf, z, u);
}

/* Fill the AB_NT node tables of a step for the common configuration: one EIH source (the Sun),
 * barycentric.  Returns an ASSIST status. */
__device__ __noinline__ int ab_fill_nodes(const AbEphem& E, const AbForceOpts& F, const double* t, AbNode* nodes) {
    const double jd_ref = E.jd_ref;
    /* coverage (reference src/spk.c:563-565, 416-418; src/ascii_ephem.c:294-295): the node times run monotonically
     * from t[0] to t[AB_NT - 1], so the two ends decide for all of them */
    for (int k = 0; k < AB_NT; k += AB_NT - 1) {
        const double jd = jd_ref + t[k];
        if (E.planets_source == AB_SRC_ASCII) {
            if (jd < E.a_beg || jd > E.a_end) return AB_ERR_COVERAGE;
        } else {
            for (int b = 0; b < AB_NPLANETS; b++) {
                const int idx = E.p_index[b];
                if (idx < 0) { if (b != 3) return AB_ERR_NEPHEM; continue; }
                if (jd < E.p_tgt[idx].beg || jd > E.p_tgt[idx].end) return AB_ERR_COVERAGE;
            }
        }
        for (int m = 0; m < E.n_ast + E.n_ast_x; m++)
            if (jd < E.a_tgt[m].beg || jd > E.a_tgt[m].end) return AB_ERR_COVERAGE;
    }
    for (int k = 0; k < AB_NT; k++) nodes[k].t = t[k];
    double u[AB_NT][3];
    if (E.planets_source == AB_SRC_ASCII) {
        double emb[AB_NT][3], lun[AB_NT][3];
        ab_ascii_pos_multi(E, 2, t, emb);
        ab_ascii_pos_multi(E, 9, t, lun);
        for (int b = 0; b < AB_NPLANETS; b++) {
            if (b == 3 || b == 4) {
                const double f = (b == 3) ? E.a_f_earth : E.a_f_moon;
                for (int k = 0; k < AB_NT; k++)
                    for (int c = 0; c < 3; c++) nodes[k].pos[b][c] = ab_divc(emb[k][c] + lun[k][c] * f, E.u_d[0], E.u_rd[0]);
            } else {
                const int col = (b == 0) ? 10 : (b <= 2 ? b - 1 : b - 2);
                ab_ascii_pos_multi(E, col, t, u);
                for (int k = 0; k < AB_NT; k++)
                    for (int c = 0; c < 3; c++) nodes[k].pos[b][c] = ab_divc(u[k][c], E.u_d[0], E.u_rd[0]);
            }
        }
    } else {
        double emb[AB_NT][3];
        bool have_emb = false;
        for (int b = 0; b < AB_NPLANETS; b++) {
            const int idx = E.p_index[b];
            if (idx < 0 || ((b == 3 || b == 4) && E.emb_index < 0)) {
                /* rare layouts (no 399 target): the one-time routine knows the fallbacks */
                for (int k = 0; k < AB_NT; k++) {
                    double GM, v[3], a[3];
                    const int flag = ab_spk_planet<0>(E, b, t[k], &GM, nodes[k].pos[b], v, a);
                    if (flag != AB_OK) return flag;
                }
                continue;
            }
            ab_spk_pos_multi(E.spkp_img, E.p_tgt[idx], jd_ref, t, u);
            if (b == 3 || b == 4) {
                if (!have_emb) { ab_spk_pos_multi(E.spkp_img, E.p_tgt[E.emb_index], jd_ref, t, emb); have_emb = true; }
                for (int k = 0; k < AB_NT; k++)
                    for (int c = 0; c < 3; c++) u[k][c] += emb[k][c];
            }
            for (int k = 0; k < AB_NT; k++)
                for (int c = 0; c < 3; c++) nodes[k].pos[b][c] = ab_divc(u[k][c], E.u_d[0], E.u_rd[0]);
        }
    }
    /* right after the planets, while their table entries are still in L1 */
    /* particle-independent EIH sums for source 0 */
    const bool need_eih = (F.forces & 0x40) != 0;
    for (int k = 0; k < AB_NT; k++) {
        AbNode& N = nodes[k];
        N.gm = E.gm;
        if (!need_eih) continue;
        double term1 = 0.0, arx = 0.0, ary = 0.0, arz = 0.0, avx = 0.0, avy = 0.0, avz = 0.0;
        for (int q = 1; q < AB_NPLANETS; q++) {
            const double GMk = E.gm[q];
            const double dxjk = N.pos[0][0] - N.pos[q][0];
            const double dyjk = N.pos[0][1] - N.pos[q][1];
            const double dzjk = N.pos[0][2] - N.pos[q][2];
            const double rjk2 = dxjk * dxjk + dyjk * dyjk + dzjk * dzjk;
            const double _rjk = sqrt(rjk2);
            term1 += GMk / _rjk;
            const double fac = GMk / (rjk2 * _rjk);
            arx -= fac * dxjk; ary -= fac * dyjk; arz -= fac * dzjk;
            const AbDivisor r3(_rjk * _rjk * _rjk);
            avx -= r3(GMk * dxjk); avy -= r3(GMk * dyjk); avz -= r3(GMk * dzjk);
        }
        N.eih_term1[0] = term1;
        N.eih_ar[0][0] = arx; N.eih_ar[0][1] = ary; N.eih_ar[0][2] = arz;
        N.eih_av[0][0] = avx; N.eih_av[0][1] = avy; N.eih_av[0][2] = avz;
    }
    /* Sun velocity (non-grav, simple GR and the EIH source) */
    for (int k = 0; k < AB_NT; k++) {
        double GM, x[3], a[3];
        const int flag = ab_planet<1>(E, 0, t[k], &GM, x, nodes[k].vel[0], a);
        if (flag != AB_OK) return flag;
    }
    /* asteroids: heliocentric SPK position / 149597870.7 + Sun.  The Sun's 24 values are copied out of the
     * 8 tables once: by now their lines have left L1, and 16 asteroids would fetch them 16 times. */
    double sun[AB_NT][3];
    for (int k = 0; k < AB_NT; k++)
        for (int c = 0; c < 3; c++) sun[k][c] = nodes[k].pos[0][c];
    for (int m = 0; m < E.n_ast; m++) {
        ab_spk_pos_multi(E.spka_img, E.a_tgt[m], jd_ref, t, u);
        for (int k = 0; k < AB_NT; k++)
            for (int c = 0; c < 3; c++)
                nodes[k].pos[AB_NPLANETS + m][c] = AB_DIVK(u[k][c], 149597870.7) + sun[k][c];
    }
    return AB_OK;
}

}  // namespace AB_NS
#endif
