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
 * The T/S/U recurrences are carried in registers instead of 32-entry arrays, and
 * velocity / acceleration sums are only formed when a caller needs them (they do
 * not feed the position sums, so skipping them changes no result).
 */
#ifndef AB_EPHEM_DEVICE_CUH
#define AB_EPHEM_DEVICE_CUH

#include "device_types.h"

#define AB_OK 0
#define AB_ERR_EPHEM_FILE 1
#define AB_ERR_AST_FILE 2
#define AB_ERR_NAST 3
#define AB_ERR_NEPHEM 4
#define AB_ERR_COVERAGE 5

namespace AB_NS {

__device__ __forceinline__ double ab_jul(double eph) { return 2451545.0 + eph / 86400.0; }

/* Chebyshev sums for one record: NCM components, P coefficients each, argument z,
 * derivative scale c.  LEVEL 0: position, 1: +velocity, 2: +acceleration. */
template <int LEVEL>
__device__ __forceinline__ void ab_cheb3(const double* __restrict__ cf, int P, double z, double c,
                                         double u[3], double v[3], double w[3]) {
    double u0 = 0.0, u1 = 0.0, u2 = 0.0;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    /* T[p-1], T[p-2] etc. */
    double Tm1 = 0.0, Tm2 = 0.0, Sm1 = 0.0, Sm2 = 0.0, Um1 = 0.0, Um2 = 0.0;
    const double* __restrict__ cx = cf;
    const double* __restrict__ cy = cf + P;
    const double* __restrict__ cz = cf + 2 * P;
    for (int p = 0; p < P; p++) {
        double T, S = 0.0, U = 0.0;
        if (p == 0) { T = 1.0; S = 0.0; U = 0.0; }
        else if (p == 1) { T = z; S = 1.0; U = 0.0; }
        else {
            T = 2.0 * z * Tm1 - Tm2;
            if (LEVEL >= 1) S = 2.0 * z * Sm1 + 2.0 * Tm1 - Sm2;
            if (LEVEL >= 2) U = (p == 2) ? 4.0 : (2.0 * z * Um1 + 4.0 * Sm1 - Um2);
        }
        const double ax = __ldg(cx + p), ay = __ldg(cy + p), az = __ldg(cz + p);
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

/* Locate the type-2 record of `tg` that holds jd_ref + t (reference src/spk.c:501-517). */
__device__ __forceinline__ const double* ab_spk_record(const double* __restrict__ img, const AbSpkTarget& tg,
                                                       double jd_ref, double t, int* P, double* z, double* c) {
    int n = (int)((jd_ref + t - tg.beg) / tg.res);
    if (n > tg.nseg - 1) n = tg.nseg - 1;       /* jd == end: the reference indexes one past; stay in the last segment */
    if (n < 0) n = 0;
    const double* val = img + tg.two[n] - 1;
    const int R = (int)__ldg(val - 1);
    *P = (R - 2) / 3;
    int b = (int)(((jd_ref - ab_jul(__ldg(val - 3))) + t) / (__ldg(val - 2) / 86400.0));
    const int nrec = (int)__ldg(val);
    if (b > nrec - 1) b = nrec - 1;
    if (b < 0) b = 0;
    const double* rec = img + (tg.one[n] - 1) + (long long)b * R;
    const double radius = __ldg(rec + 1);
    *z = ((jd_ref - ab_jul(__ldg(rec))) + t) / (radius / 86400.0);
    *c = 1.0 / radius;
    return rec + 2;
}

template <int LEVEL>
__device__ __forceinline__ void ab_spk_target_pos(const double* __restrict__ img, const AbSpkTarget& tg,
                                                  double jd_ref, double t, double u[3], double v[3], double w[3]) {
    int P; double z, c;
    const double* cf = ab_spk_record(img, tg, jd_ref, t, &P, &z, &c);
    ab_cheb3<LEVEL>(cf, P, z, c, u, v, w);
}

/* Planet (body < 11) from an SPK kernel, reference src/spk.c:550-610, 632-693. */
template <int LEVEL>
__device__ int ab_spk_planet(const AbEphem& E, int body, double t, double* GM, double x[3], double v[3], double a[3]) {
    const int naif[AB_NPLANETS] = {10, 1, 2, 399, 301, 4, 5, 6, 7, 8, 9};
    const double jd_ref = E.jd_ref;
    double u[3], uv[3] = {0, 0, 0}, uw[3] = {0, 0, 0};
    const int idx = E.p_index[body];
    if (idx >= 0) {
        const AbSpkTarget& tg = E.p_tgt[idx];
        if (jd_ref + t < tg.beg || jd_ref + t > tg.end) return AB_ERR_COVERAGE;
        *GM = tg.mass;
        ab_spk_target_pos<LEVEL>(E.spkp_img, tg, jd_ref, t, u, uv, uw);
        const int code = naif[body];
        if (code == 301 || code == 399) {
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
    const double au = E.AU;
    const double seconds_per_day = 86400.;
    for (int i = 0; i < 3; i++) {
        x[i] = u[i] / au;
        if (LEVEL >= 1) v[i] = uv[i] / (au / seconds_per_day);
        if (LEVEL >= 2) a[i] = uw[i] / (au / (seconds_per_day * seconds_per_day));
    }
    return AB_OK;
}

/* One column of a DE-binary record, reference src/ascii_ephem.c:27-65. */
template <int LEVEL>
__device__ __forceinline__ void ab_ascii_work(const double* __restrict__ Pcol, int ncm, int ncf, int niv,
                                              double t0, double t1, double u[3], double v[3], double w[3]) {
    const double tt = t0 * (double)niv;
    const int b = (int)tt;
    const double frac = tt - (double)b;            /* == fmod(tt, 1.0) for tt >= 0, exactly */
    const double z = 2.0 * frac - 1.0;
    const double c = (double)(niv * 2) / t1 / 86400.0;
    ab_cheb3<LEVEL>(Pcol + ncf * (b * ncm), ncf, z, c, u, v, w);
}

/* Planet (body < 11) from a DE-binary file, reference src/ascii_ephem.c:275-384. */
template <int LEVEL>
__device__ int ab_ascii_planet(const AbEphem& E, int body, double t, double* GM, double x[3], double v[3], double a[3]) {
    const double jd_ref = E.jd_ref;
    *GM = E.a_mass[body];
    if (jd_ref + t < E.a_beg || jd_ref + t > E.a_end) return AB_ERR_COVERAGE;
    long long blk = (long long)(unsigned int)((jd_ref + t - E.a_beg) / E.a_inc);
    if (blk > E.a_nrec - 1) blk = E.a_nrec - 1;    /* jd == end */
    const double* z = E.ascii_img + (blk + 2) * E.a_rec_words;
    const double tr = ((jd_ref - E.a_beg - (double)blk * E.a_inc) + t) / E.a_inc;
    /* ASCII_* column of each ASSIST body */
    const int col_of[AB_NPLANETS] = {10, 0, 1, 2, 2, 3, 4, 5, 6, 7, 8};
    double u[3], uv[3] = {0, 0, 0}, uw[3] = {0, 0, 0};
    const int col = col_of[body];
    ab_ascii_work<LEVEL>(z + E.a_off[col], 3, E.a_ncf[col], E.a_niv[col], tr, E.a_inc, u, uv, uw);
    if (body == 3 || body == 4) {
        double l[3], lv[3] = {0, 0, 0}, lw[3] = {0, 0, 0};
        ab_ascii_work<LEVEL>(z + E.a_off[9], 3, E.a_ncf[9], E.a_niv[9], tr, E.a_inc, l, lv, lw);
        const double f = (body == 3) ? (-1.0 / (1.0 + E.a_cem)) : (E.a_cem / (1.0 + E.a_cem));
        for (int i = 0; i < 3; i++) { u[i] += l[i] * f; uv[i] += lv[i] * f; uw[i] += lw[i] * f; }
    }
    for (int i = 0; i < 3; i++) {
        x[i] = u[i] / E.a_cau;
        if (LEVEL >= 1) v[i] = uv[i] / (E.a_cau / 86400.);
        if (LEVEL >= 2) a[i] = uw[i] / (E.a_cau / (86400. * 86400.));
    }
    return AB_OK;
}

template <int LEVEL>
__device__ __forceinline__ int ab_planet(const AbEphem& E, int body, double t, double* GM, double x[3], double v[3], double a[3]) {
    if (E.planets_source == AB_SRC_ASCII) return ab_ascii_planet<LEVEL>(E, body, t, GM, x, v, a);
    return ab_spk_planet<LEVEL>(E, body, t, GM, x, v, a);
}

/* Heliocentric asteroid position in AU, reference src/spk.c:405-481. */
__device__ int ab_asteroid(const AbEphem& E, int m, double t, double* GM, double x[3]) {
    if (E.spka_img == nullptr) return AB_ERR_AST_FILE;
    if (m < 0 || m >= E.n_ast) return AB_ERR_NAST;
    const AbSpkTarget& tg = E.a_tgt[m];
    const double jd_ref = E.jd_ref;
    if (jd_ref + t < tg.beg || jd_ref + t > tg.end) return AB_ERR_COVERAGE;
    *GM = tg.mass;
    int P; double z, c;
    const double* cf = ab_spk_record(E.spka_img, tg, jd_ref, t, &P, &z, &c);
    double u[3], dv[3], dw[3];
    ab_cheb3<0>(cf, P, z, c, u, dv, dw);
    x[0] = u[0] / 149597870.7; x[1] = u[1] / 149597870.7; x[2] = u[2] / 149597870.7;
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
    const bool need_ast = (F.forces & 0x04) != 0 || true;   /* variational direct term ignores the mask (src/forces.c:359) */
    if (need_ast) {
        for (int m = 0; m < E.n_ast; m++) {
            double x[3];
            int flag = ab_asteroid(E, m, t, &B.gm[AB_NPLANETS + m], x);
            if (flag != AB_OK && status == AB_OK) status = flag;
            /* heliocentric -> barycentric, reference src/forces.c:213-219 */
            B.pos[AB_NPLANETS + m][0] = x[0] + B.pos[0][0];
            B.pos[AB_NPLANETS + m][1] = x[1] + B.pos[0][1];
            B.pos[AB_NPLANETS + m][2] = x[2] + B.pos[0][2];
        }
    }
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
                avx -= GMk * dxjk / (_rjk * _rjk * _rjk);
                avy -= GMk * dyjk / (_rjk * _rjk * _rjk);
                avz -= GMk * dzjk / (_rjk * _rjk * _rjk);
            }
            B.eih_term1[j] = term1;
            B.eih_ar[j][0] = arx; B.eih_ar[j][1] = ary; B.eih_ar[j][2] = arz;
            B.eih_av[j][0] = avx; B.eih_av[j][1] = avy; B.eih_av[j][2] = avz;
        }
    }
    B.status = status;
}

}  // namespace AB_NS
#endif
