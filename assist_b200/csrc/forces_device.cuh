/*
 * forces_device.cuh -- the ASSIST force model, fused, one thread per system.
 *
 * A system is a real test particle plus its nv first-order variational particles.
 * One call sums every enabled term for the real particle and the matching
 * Jacobian-times-variation products for its variational particles, in ONE pass
 * over the body table (the reference makes seven passes over the particle array,
 * reference src/forces.c:120-147).  Accumulation order is the reference's:
 *   non-grav -> Earth J2/J3/J4 -> solar J2 -> EIH -> potential GR -> simple GR -> direct
 *   direct: asteroids, then Pluto, Moon, Mars, Mercury, Neptune, Uranus, Earth,
 *           Venus, Saturn, Jupiter, Sun            (src/forces.c:281-306)
 * and every expression keeps the reference's operation order, so that a build
 * without FMA contraction reproduces the C build's doubles.  Quirks kept on
 * purpose: the variational direct term ignores the force mask (src/forces.c:359);
 * the variational EIH pass rounds its prefactors differently from the real pass
 * (src/forces.c:1591 vs :1349, :1766 vs :1429-1432).
 *
 *   direct         src/forces.c:266-433        solar J2      src/forces.c:644-772
 *   Earth J2-J4    src/forces.c:435-642        non-grav      src/forces.c:774-1057
 *   potential GR   src/forces.c:1059-1161      simple GR     src/forces.c:1163-1286
 *   EIH GR         src/forces.c:1288-1983
 */
#ifndef AB_FORCES_DEVICE_CUH
#define AB_FORCES_DEVICE_CUH

#include "device_types.h"
#include "fp_device.cuh"

namespace AB_NS {

/* Positions/velocities/accelerations of one system in registers/local memory. */
template <int KM>
struct AbSysT {
    double x[KM][3];
    double v[KM][3];
    double a[KM][3];
    double prm[KM][3];
    int nv_;
    /* KM == 1 means "no variational particles": the count is a compile-time zero so that
     * every array index is static and the state lives in registers */
    __device__ __forceinline__ int nv() const { return (KM == 1) ? 0 : nv_; }
};

/* 3x6 Jacobian applied to every variational particle of the system. */
#define AB_APPLY_J36(S, dxdx, dxdy, dxdz, dxdvx, dxdvy, dxdvz, dydx, dydy, dydz, dydvx, dydvy, dydvz, dzdx, dzdy, dzdz, dzdvx, dzdvy, dzdvz) \
    for (int vv = 1; vv <= (S).nv(); vv++) {                                                             \
        const double ddx = (S).x[vv][0], ddy = (S).x[vv][1], ddz = (S).x[vv][2];                     \
        const double ddvx = (S).v[vv][0], ddvy = (S).v[vv][1], ddvz = (S).v[vv][2];                  \
        const double dax = ddx * dxdx + ddy * dxdy + ddz * dxdz + ddvx * dxdvx + ddvy * dxdvy + ddvz * dxdvz; \
        const double day = ddx * dydx + ddy * dydy + ddz * dydz + ddvx * dydvx + ddvy * dydvy + ddvz * dydvz; \
        const double daz = ddx * dzdx + ddy * dzdy + ddz * dzdz + ddvx * dzdvx + ddvy * dzdvy + ddvz * dzdvz; \
        (S).a[vv][0] += dax; (S).a[vv][1] += day; (S).a[vv][2] += daz;                                 \
    }

/* ---- Marsden non-gravitational term, reference src/forces.c:774-1057 ------- */
template <int KM, class BT>
__device__ void ab_force_nongrav(const AbForceOpts& F, const BT& B, AbSysT<KM>& S,
                                 double xo, double yo, double zo, double vxo, double vyo, double vzo) {
    if (!F.has_params) return;
    const double A1 = S.prm[0][0], A2 = S.prm[0][1], A3 = S.prm[0][2];
    if (A1 == 0. && A2 == 0. && A3 == 0.) return;
    const double xr = B.pos[0][0], yr = B.pos[0][1], zr = B.pos[0][2];
    const double vxr = B.vel[0][0], vyr = B.vel[0][1], vzr = B.vel[0][2];
    const double alpha = F.alpha, nk = F.nk, nm = F.nm, nn = F.nn, r0 = F.r0;

    double dx = S.x[0][0] + (xo - xr);
    double dy = S.x[0][1] + (yo - yr);
    double dz = S.x[0][2] + (zo - zr);
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double r = sqrt(r2);
    const double g = alpha * pow(r / r0, -nm) * pow(1.0 + pow(r / r0, nn), -nk);

    double dvx = S.v[0][0] + (vxo - vxr);
    double dvy = S.v[0][1] + (vyo - vyr);
    double dvz = S.v[0][2] + (vzo - vzr);

    double hx = dy * dvz - dz * dvy;
    double hy = dz * dvx - dx * dvz;
    double hz = dx * dvy - dy * dvx;
    double h2 = hx * hx + hy * hy + hz * hz;
    double h = sqrt(h2);

    double tx = hy * dz - hz * dy;
    double ty = hz * dx - hx * dz;
    double tz = hx * dy - hy * dx;
    const double t2 = tx * tx + ty * ty + tz * tz;
    const double _t = sqrt(t2);

    S.a[0][0] += A1 * g * dx / r + A2 * g * tx / _t + A3 * g * hx / h;
    S.a[0][1] += A1 * g * dy / r + A2 * g * ty / _t + A3 * g * hy / h;
    S.a[0][2] += A1 * g * dz / r + A2 * g * tz / _t + A3 * g * hz / h;

    if (S.nv() == 0) return;

    const double r3 = r * r * r;
    const double v2 = dvx * dvx + dvy * dvy + dvz * dvz;
    const double rdotv = dx * dvx + dy * dvy + dz * dvz;
    const double vdott = dvx * tx + dvy * ty + dvz * tz;

    const double dgdr = (alpha / r0) * (-nm * pow(r / r0, -nm - 1) * pow(1.0 + pow(r / r0, nn), -nk)
                                        + pow(r / r0, -nm) * (-nk * nn) * pow(r / r0, nn - 1) * pow(1.0 + pow(r / r0, nn), -nk - 1));
    const double dgx = dgdr * dx / r;
    const double dgy = dgdr * dy / r;
    const double dgz = dgdr * dz / r;

    const double hxh3 = hx / (h * h * h);
    const double hyh3 = hy / (h * h * h);
    const double hzh3 = hz / (h * h * h);

    const double txt3 = tx / (_t * _t * _t);
    const double tyt3 = ty / (_t * _t * _t);
    const double tzt3 = tz / (_t * _t * _t);

    const double dxdA1 = g * dx / r, dydA1 = g * dy / r, dzdA1 = g * dz / r;
    const double dxdA2 = g * tx / _t, dydA2 = g * ty / _t, dzdA2 = g * tz / _t;
    const double dxdA3 = g * hx / h, dydA3 = g * hy / h, dzdA3 = g * hz / h;

    const double dxdx = A1 * (dgx * dx / r + g * (1. / r - dx * dx / r3))
        + A2 * (dgx * tx / _t + g * ((dx * dvx - rdotv) / _t - txt3 * (2. * dx * vdott - rdotv * tx)))
        + A3 * (dgx * hx / h + g * (-hxh3) * (v2 * dx - rdotv * dvx));
    const double dydy = A1 * (dgy * dy / r + g * (1. / r - dy * dy / r3))
        + A2 * (dgy * ty / _t + g * ((dy * dvy - rdotv) / _t - tyt3 * (2. * dy * vdott - rdotv * ty)))
        + A3 * (dgy * hy / h + g * (-hyh3) * (v2 * dy - rdotv * dvy));
    const double dzdz = A1 * (dgz * dz / r + g * (1. / r - dz * dz / r3))
        + A2 * (dgz * tz / _t + g * ((dz * dvz - rdotv) / _t - tzt3 * (2. * dz * vdott - rdotv * tz)))
        + A3 * (dgz * hz / h + g * (-hzh3) * (v2 * dz - rdotv * dvz));
    const double dxdy = A1 * (dgy * dx / r + g * (-dx * dy / r3))
        + A2 * (dgy * tx / _t + g * ((2 * dy * dvx - dx * dvy) / _t - txt3 * (2 * dy * vdott - rdotv * ty)))
        + A3 * (dgy * hx / h + g * (dvz / h - hxh3 * (v2 * dy - rdotv * dvy)));
    const double dydx = A1 * (dgx * dy / r + g * (-dy * dx / r3))
        + A2 * (dgx * ty / _t + g * ((2 * dx * dvy - dy * dvx) / _t - tyt3 * (2 * dx * vdott - rdotv * tx)))
        + A3 * (dgx * hy / h + g * (-dvz / h - hyh3 * (v2 * dx - rdotv * dvx)));
    const double dxdz = A1 * (dgz * dx / r + g * (-dx * dz / r3))
        + A2 * (dgz * tx / _t + g * ((2 * dz * dvx - dx * dvz) / _t - txt3 * (2 * dz * vdott - rdotv * tz)))
        + A3 * (dgz * hx / h + g * (-dvy / h - hxh3 * (v2 * dz - rdotv * dvz)));
    const double dzdx = A1 * (dgx * dz / r + g * (-dz * dx / r3))
        + A2 * (dgx * tz / _t + g * ((2 * dx * dvz - dz * dvx) / _t - tzt3 * (2 * dx * vdott - rdotv * tx)))
        + A3 * (dgx * hz / h + g * (dvy / h - hzh3 * (v2 * dx - rdotv * dvx)));
    const double dydz = A1 * (dgz * dy / r + g * (-dy * dz / r3))
        + A2 * (dgz * ty / _t + g * ((2 * dz * dvy - dy * dvz) / _t - tyt3 * (2 * dz * vdott - rdotv * tz)))
        + A3 * (dgz * hy / h + g * (dvx / h - hyh3 * (v2 * dz - rdotv * dvz)));
    const double dzdy = A1 * (dgy * dz / r + g * (-dz * dy / r3))
        + A2 * (dgy * tz / _t + g * ((2 * dy * dvz - dz * dvy) / _t - tzt3 * (2 * dy * vdott - rdotv * ty)))
        + A3 * (dgy * hz / h + g * (-dvx / h - hzh3 * (v2 * dy - rdotv * dvy)));

    const double dxdvx = A1 * (0.) + A2 * g * ((dy * dy + dz * dz) / _t - txt3 * r2 * tx) + A3 * g * (-hxh3 * (r2 * dvx - dx * rdotv));
    const double dydvy = A1 * (0.) + A2 * g * ((dx * dx + dz * dz) / _t - tyt3 * r2 * ty) + A3 * g * (-hyh3 * (r2 * dvy - dy * rdotv));
    const double dzdvz = A1 * (0.) + A2 * g * ((dx * dx + dy * dy) / _t - tzt3 * r2 * tz) + A3 * g * (-hzh3 * (r2 * dvz - dz * rdotv));
    const double dxdvy = A1 * (0.) + A2 * g * (-dy * dx / _t - tyt3 * r2 * tx) + A3 * g * (-dz / h - hxh3 * (r2 * dvy - dy * rdotv));
    const double dydvx = A1 * (0.) + A2 * g * (-dx * dy / _t - txt3 * r2 * ty) + A3 * g * (dz / h - hyh3 * (r2 * dvx - dx * rdotv));
    const double dxdvz = A1 * (0.) + A2 * g * (-dz * dx / _t - tzt3 * r2 * tx) + A3 * g * (dy / h - hxh3 * (r2 * dvz - dz * rdotv));
    const double dzdvx = A1 * (0.) + A2 * g * (-dx * dz / _t - txt3 * r2 * tz) + A3 * g * (-dy / h - hzh3 * (r2 * dvx - dx * rdotv));
    const double dydvz = A1 * (0.) + A2 * g * (-dz * dy / _t - tzt3 * r2 * ty) + A3 * g * (-dx / h - hyh3 * (r2 * dvz - dz * rdotv));
    const double dzdvy = A1 * (0.) + A2 * g * (-dy * dz / _t - tyt3 * r2 * tz) + A3 * g * (dx / h - hzh3 * (r2 * dvy - dy * rdotv));

    for (int vv = 1; vv <= S.nv(); vv++) {
        const double ddx = S.x[vv][0], ddy = S.x[vv][1], ddz = S.x[vv][2];
        const double ddvx = S.v[vv][0], ddvy = S.v[vv][1], ddvz = S.v[vv][2];
        const double dA1 = S.prm[vv][0], dA2 = S.prm[vv][1], dA3 = S.prm[vv][2];
        const double dax = ddx * dxdx + ddy * dxdy + ddz * dxdz
            + ddvx * dxdvx + ddvy * dxdvy + ddvz * dxdvz + dA1 * dxdA1 + dA2 * dxdA2 + dA3 * dxdA3;
        const double day = ddx * dydx + ddy * dydy + ddz * dydz
            + ddvx * dydvx + ddvy * dydvy + ddvz * dydvz + dA1 * dydA1 + dA2 * dydA2 + dA3 * dydA3;
        const double daz = ddx * dzdx + ddy * dzdy + ddz * dzdz
            + ddvx * dzdvx + ddvy * dzdvy + ddvz * dzdvz + dA1 * dzdA1 + dA2 * dzdA2 + dA3 * dzdA3;
        S.a[vv][0] += dax; S.a[vv][1] += day; S.a[vv][2] += daz;
    }
}

/* ---- Earth J2/J3/J4, reference src/forces.c:435-642 ------------------------ */
template <int KM, class BT>
__device__ void ab_force_earth_harmonics(const AbEphem& E, const AbForceOpts& F, const BT& B, AbSysT<KM>& S,
                                         double xo, double yo, double zo) {
    const double GMearth = B.gm[3];
    const double xr = B.pos[3][0], yr = B.pos[3][1], zr = B.pos[3][2];
    const double J2e = E.J2E, J3e = E.J3E, J4e = E.J4E, Re_eq = E.Re_eq;
    const double cosa = F.e_cosa, sina = F.e_sina, cosd = F.e_cosd, sind = F.e_sind;

    double dx = S.x[0][0] + (xo - xr);
    double dy = S.x[0][1] + (yo - yr);
    double dz = S.x[0][2] + (zo - zr);
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double r = sqrt(r2);
    const AbDivisor dr2(r2), dr(r);   /* every /r2 and /r below is an exact quotient from one reciprocal */

    double dxp = -dx * sina + dy * cosa;
    double dyp = -dx * cosa * sind - dy * sina * sind + dz * cosd;
    double dzp = dx * cosa * cosd + dy * sina * cosd + dz * sind;
    dx = dxp; dy = dyp; dz = dzp;

    const double costheta2 = dr2(dz * dz);
    const double J2e_prefac = dr(dr2(dr2(3. * J2e * Re_eq * Re_eq))) / 2.;
    const double J2e_fac = 5. * costheta2 - 1.;

    double resx = GMearth * J2e_prefac * J2e_fac * dx;
    double resy = GMearth * J2e_prefac * J2e_fac * dy;
    double resz = GMearth * J2e_prefac * (J2e_fac - 2.) * dz;

    const double J3e_prefac = dr(dr2(dr2(5. * J3e * Re_eq * Re_eq * Re_eq))) / 2.;
    const double J3e_fac = 3. - 7. * costheta2;

    resx += -GMearth * J3e_prefac * dr2(1.) * J3e_fac * dx * dz;
    resy += -GMearth * J3e_prefac * dr2(1.) * J3e_fac * dy * dz;
    resz += -GMearth * J3e_prefac * (6. * costheta2 - 7. * costheta2 * costheta2 - 0.6);

    const double J4e_prefac = dr(dr2(dr2(dr2(5. * J4e * Re_eq * Re_eq * Re_eq * Re_eq)))) / 8.;
    const double J4e_fac = 63. * costheta2 * costheta2 - 42. * costheta2 + 3.;

    resx += GMearth * J4e_prefac * J4e_fac * dx;
    resy += GMearth * J4e_prefac * J4e_fac * dy;
    resz += GMearth * J4e_prefac * (J4e_fac + 12. - 28. * costheta2) * dz;

    double resxp = -resx * sina - resy * cosa * sind + resz * cosa * cosd;
    double resyp = resx * cosa - resy * sina * sind + resz * sina * cosd;
    double reszp = +resy * cosd + resz * sind;

    S.a[0][0] += resxp; S.a[0][1] += resyp; S.a[0][2] += reszp;

    if (S.nv() == 0) return;

    const double J2e_fac2 = 7. * costheta2 - 1.;
    const double J2e_fac3 = 35. * costheta2 * costheta2 - 30. * costheta2 + 3.;

    const double dxdx = GMearth * J2e_prefac * (J2e_fac - dr2(5. * J2e_fac2 * dx * dx));
    const double dydy = GMearth * J2e_prefac * (J2e_fac - dr2(5. * J2e_fac2 * dy * dy));
    const double dzdz = GMearth * J2e_prefac * (-1.) * J2e_fac3;
    const double dxdy = dr2(GMearth * J2e_prefac * (-5.) * J2e_fac2 * dx * dy);
    const double dydz = dr2(GMearth * J2e_prefac * (-5.) * (J2e_fac2 - 2.) * dy * dz);
    const double dxdz = dr2(GMearth * J2e_prefac * (-5.) * (J2e_fac2 - 2.) * dx * dz);

    const double costheta = dr(dz);
    const double J3e_fac2 = dr2(21 * (-3. * costheta2 + 1.));
    const double J3e_fac3 = dr2(3 * (-21. * costheta2 * costheta2 + 14. * costheta2 - 1.));
    const double J3e_fac4 = dr((-63. * costheta2 * costheta2 + 70. * costheta2 - 15.) * costheta);

    const double dxdxJ3 = dr(GMearth * J3e_prefac * costheta * (J3e_fac2 * dx * dx - J3e_fac));
    const double dydyJ3 = dr(GMearth * J3e_prefac * costheta * (J3e_fac2 * dy * dy - J3e_fac));
    const double dzdzJ3 = GMearth * J3e_prefac * J3e_fac4;
    const double dxdyJ3 = dr(GMearth * J3e_prefac * J3e_fac2 * costheta * dx * dy);
    const double dydzJ3 = GMearth * J3e_prefac * J3e_fac3 * dy;
    const double dxdzJ3 = GMearth * J3e_prefac * J3e_fac3 * dx;

    const double J4e_fac2 = 33. * costheta2 * costheta2 - 18. * costheta2 + 1.;
    const double J4e_fac3 = 33. * costheta2 * costheta2 - 30. * costheta2 + 5.;
    const double J4e_fac4 = 231. * costheta2 * costheta2 * costheta2 - 315. * costheta2 * costheta2 + 105. * costheta2 - 5.;

    const double dxdxJ4 = GMearth * J4e_prefac * (J4e_fac - dr2(21. * J4e_fac2 * dx * dx));
    const double dydyJ4 = GMearth * J4e_prefac * (J4e_fac - dr2(21. * J4e_fac2 * dy * dy));
    const double dzdzJ4 = GMearth * J4e_prefac * (-3.) * J4e_fac4;
    const double dxdyJ4 = dr2(GMearth * J4e_prefac * (-21.) * J4e_fac2 * dx * dy);
    const double dydzJ4 = dr2(GMearth * J4e_prefac * (-21.) * J4e_fac3 * dy * dz);
    const double dxdzJ4 = dr2(GMearth * J4e_prefac * (-21.) * J4e_fac3 * dx * dz);

    for (int vv = 1; vv <= S.nv(); vv++) {
        const double ddx = S.x[vv][0], ddy = S.x[vv][1], ddz = S.x[vv][2];
        double ddxp = -ddx * sina + ddy * cosa;
        double ddyp = -ddx * cosa * sind - ddy * sina * sind + ddz * cosd;
        double ddzp = ddx * cosa * cosd + ddy * sina * cosd + ddz * sind;

        double dax = ddxp * dxdx + ddyp * dxdy + ddzp * dxdz;
        double day = ddxp * dxdy + ddyp * dydy + ddzp * dydz;
        double daz = ddxp * dxdz + ddyp * dydz + ddzp * dzdz;

        dax += ddxp * dxdxJ3 + ddyp * dxdyJ3 + ddzp * dxdzJ3;
        day += ddxp * dxdyJ3 + ddyp * dydyJ3 + ddzp * dydzJ3;
        daz += ddxp * dxdzJ3 + ddyp * dydzJ3 + ddzp * dzdzJ3;

        dax += ddxp * dxdxJ4 + ddyp * dxdyJ4 + ddzp * dxdzJ4;
        day += ddxp * dxdyJ4 + ddyp * dydyJ4 + ddzp * dydzJ4;
        daz += ddxp * dxdzJ4 + ddyp * dydzJ4 + ddzp * dzdzJ4;

        double daxp = -dax * sina - day * cosa * sind + daz * cosa * cosd;
        double dayp = dax * cosa - day * sina * sind + daz * sina * cosd;
        double dazp = +day * cosd + daz * sind;

        S.a[vv][0] += daxp; S.a[vv][1] += dayp; S.a[vv][2] += dazp;
    }
}

/* ---- solar J2, reference src/forces.c:644-772 ------------------------------ */
template <int KM, class BT>
__device__ void ab_force_solar_j2(const AbEphem& E, const AbForceOpts& F, const BT& B, AbSysT<KM>& S,
                                  double xo, double yo, double zo) {
    const double GMsun = B.gm[0];
    const double xr = B.pos[0][0], yr = B.pos[0][1], zr = B.pos[0][2];
    const double Rs_eq = E.Rs_eq, J2s = E.J2SUN;
    const double cosa = F.s_cosa, sina = F.s_sina, cosd = F.s_cosd, sind = F.s_sind;

    double dx = S.x[0][0] + (xo - xr);
    double dy = S.x[0][1] + (yo - yr);
    double dz = S.x[0][2] + (zo - zr);
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double r = sqrt(r2);
    const AbDivisor dr2(r2), dr(r);

    double dxp = -dx * sina + dy * cosa;
    double dyp = -dx * cosa * sind - dy * sina * sind + dz * cosd;
    double dzp = dx * cosa * cosd + dy * sina * cosd + dz * sind;
    dx = dxp; dy = dyp; dz = dzp;

    const double costheta2 = dr2(dz * dz);
    const double J2s_prefac = dr(dr2(dr2(3. * J2s * Rs_eq * Rs_eq))) / 2.;
    const double J2s_fac = 5. * costheta2 - 1.;
    const double J2s_fac2 = 7. * costheta2 - 1.;
    const double J2s_fac3 = 35. * costheta2 * costheta2 - 30. * costheta2 + 3.;

    double resx = GMsun * J2s_prefac * J2s_fac * dx;
    double resy = GMsun * J2s_prefac * J2s_fac * dy;
    double resz = GMsun * J2s_prefac * (J2s_fac - 2.) * dz;

    double resxp = -resx * sina - resy * cosa * sind + resz * cosa * cosd;
    double resyp = resx * cosa - resy * sina * sind + resz * sina * cosd;
    double reszp = +resy * cosd + resz * sind;

    S.a[0][0] += resxp; S.a[0][1] += resyp; S.a[0][2] += reszp;

    if (S.nv() == 0) return;

    const double dxdx = GMsun * J2s_prefac * (J2s_fac - dr2(5. * J2s_fac2 * dx * dx));
    const double dydy = GMsun * J2s_prefac * (J2s_fac - dr2(5. * J2s_fac2 * dy * dy));
    const double dzdz = GMsun * J2s_prefac * (-1.) * J2s_fac3;
    const double dxdy = dr2(GMsun * J2s_prefac * (-5.) * J2s_fac2 * dx * dy);
    const double dydz = dr2(GMsun * J2s_prefac * (-5.) * (J2s_fac2 - 2.) * dy * dz);
    const double dxdz = dr2(GMsun * J2s_prefac * (-5.) * (J2s_fac2 - 2.) * dx * dz);

    for (int vv = 1; vv <= S.nv(); vv++) {
        double ddx = S.x[vv][0], ddy = S.x[vv][1], ddz = S.x[vv][2];
        double ddxp = -ddx * sina + ddy * cosa;
        double ddyp = -ddx * cosa * sind - ddy * sina * sind + ddz * cosd;
        double ddzp = ddx * cosa * cosd + ddy * sina * cosd + ddz * sind;
        ddx = ddxp; ddy = ddyp; ddz = ddzp;

        double daxp = ddx * dxdx + ddy * dxdy + ddz * dxdz;
        double dayp = ddx * dxdy + ddy * dydy + ddz * dydz;
        double dazp = ddx * dxdz + ddy * dydz + ddz * dzdz;

        double dax = -daxp * sina - dayp * cosa * sind + dazp * cosa * cosd;
        double day = daxp * cosa - dayp * sina * sind + dazp * sina * cosd;
        double daz = +dayp * cosd + dazp * sind;

        S.a[vv][0] += dax; S.a[vv][1] += day; S.a[vv][2] += daz;
    }
}

/* ---- potential GR (Nobili & Roxburgh), reference src/forces.c:1059-1161 ---- */
template <int KM, class BT>
__device__ void ab_force_potential_gr(const AbEphem& E, const BT& B, AbSysT<KM>& S, double xo, double yo, double zo) {
    const double C2 = E.c_squared;
    const double GMsun = B.gm[0];
    const double px = S.x[0][0] + (xo - B.pos[0][0]);
    const double py = S.x[0][1] + (yo - B.pos[0][1]);
    const double pz = S.x[0][2] + (zo - B.pos[0][2]);
    const double r2 = px * px + py * py + pz * pz;
    const double r = sqrt(r2);
    const double prefac = -6.0 * GMsun * GMsun / (C2 * r2 * r2);
    S.a[0][0] += prefac * px; S.a[0][1] += prefac * py; S.a[0][2] += prefac * pz;
    if (S.nv() == 0) return;
    const double dxdx = prefac + -4.0 * prefac * (px / r) * (px / r);
    const double dxdy = -4.0 * prefac * (px / r) * (py / r);
    const double dxdz = -4.0 * prefac * (px / r) * (pz / r);
    const double dydx = -4.0 * prefac * (py / r) * (px / r);
    const double dydy = prefac + -4.0 * prefac * (py / r) * (py / r);
    const double dydz = -4.0 * prefac * (py / r) * (pz / r);
    const double dzdx = -4.0 * prefac * (pz / r) * (px / r);
    const double dzdy = -4.0 * prefac * (pz / r) * (py / r);
    const double dzdz = prefac + -4.0 * prefac * (pz / r) * (pz / r);
    for (int vv = 1; vv <= S.nv(); vv++) {
        const double ddx = S.x[vv][0], ddy = S.x[vv][1], ddz = S.x[vv][2];
        const double dax = ddx * dxdx + ddy * dxdy + ddz * dxdz;
        const double day = ddx * dydx + ddy * dydy + ddz * dydz;
        const double daz = ddx * dzdx + ddy * dzdy + ddz * dzdz;
        S.a[vv][0] += dax; S.a[vv][1] += day; S.a[vv][2] += daz;
    }
}

/* ---- simple GR (Damour & Deruelle), reference src/forces.c:1163-1286 ------- */
template <int KM, class BT>
__device__ void ab_force_simple_gr(const AbEphem& E, const BT& B, AbSysT<KM>& S,
                                   double xo, double yo, double zo, double vxo, double vyo, double vzo) {
    const double C2 = E.c_squared;
    const double GMsun = B.gm[0];
    const double px = S.x[0][0] + (xo - B.pos[0][0]);
    const double py = S.x[0][1] + (yo - B.pos[0][1]);
    const double pz = S.x[0][2] + (zo - B.pos[0][2]);
    const double pvx = S.v[0][0] + (vxo - B.vel[0][0]);
    const double pvy = S.v[0][1] + (vyo - B.vel[0][1]);
    const double pvz = S.v[0][2] + (vzo - B.vel[0][2]);

    const double v2 = pvx * pvx + pvy * pvy + pvz * pvz;
    const double r = sqrt(px * px + py * py + pz * pz);
    const double A = 4.0 * GMsun / r - v2;
    const double Bq = 4.0 * (px * pvx + py * pvy + pz * pvz);
    const double prefac = GMsun / (r * r * r * C2);

    S.a[0][0] += prefac * (A * px + Bq * pvx);
    S.a[0][1] += prefac * (A * py + Bq * pvy);
    S.a[0][2] += prefac * (A * pz + Bq * pvz);
    if (S.nv() == 0) return;

    const double dpdr = -3.0 * prefac / r;
    const double dxdx = dpdr * px / r * (A * px + Bq * pvx) + prefac * (A - px * (px / r) * 4.0 * GMsun / (r * r) + 4.0 * pvx * pvx);
    const double dxdy = dpdr * py / r * (A * px + Bq * pvx) + prefac * (-px * (py / r) * 4.0 * GMsun / (r * r) + 4.0 * pvy * pvx);
    const double dxdz = dpdr * pz / r * (A * px + Bq * pvx) + prefac * (-px * (pz / r) * 4.0 * GMsun / (r * r) + 4.0 * pvz * pvx);
    const double dxdvx = prefac * (-2.0 * pvx * px + 4.0 * px * pvx + Bq);
    const double dxdvy = prefac * (-2.0 * pvy * px + 4.0 * py * pvx);
    const double dxdvz = prefac * (-2.0 * pvz * px + 4.0 * pz * pvx);

    const double dydx = dpdr * px / r * (A * py + Bq * pvy) + prefac * (-py * (px / r) * 4.0 * GMsun / (r * r) + 4.0 * pvx * pvy);
    const double dydy = dpdr * py / r * (A * py + Bq * pvy) + prefac * (A - py * (py / r) * 4.0 * GMsun / (r * r) + 4.0 * pvy * pvy);
    const double dydz = dpdr * pz / r * (A * py + Bq * pvy) + prefac * (-py * (pz / r) * 4.0 * GMsun / (r * r) + 4.0 * pvz * pvy);
    const double dydvx = prefac * (-2.0 * pvx * py + 4.0 * px * pvy);
    const double dydvy = prefac * (-2.0 * pvy * py + 4.0 * py * pvy + Bq);
    const double dydvz = prefac * (-2.0 * pvz * py + 4.0 * pz * pvy);

    const double dzdx = dpdr * px / r * (A * pz + Bq * pvz) + prefac * (-pz * (px / r) * 4.0 * GMsun / (r * r) + 4.0 * pvx * pvz);
    const double dzdy = dpdr * py / r * (A * pz + Bq * pvz) + prefac * (-pz * (py / r) * 4.0 * GMsun / (r * r) + 4.0 * pvy * pvz);
    const double dzdz = dpdr * pz / r * (A * pz + Bq * pvz) + prefac * (A - pz * (pz / r) * 4.0 * GMsun / (r * r) + 4.0 * pvz * pvz);
    const double dzdvx = prefac * (-2.0 * pvx * pz + 4.0 * px * pvz);
    const double dzdvy = prefac * (-2.0 * pvy * pz + 4.0 * py * pvz);
    const double dzdvz = prefac * (-2.0 * pvz * pz + 4.0 * pz * pvz + Bq);

    AB_APPLY_J36(S, dxdx, dxdy, dxdz, dxdvx, dxdvy, dxdvz, dydx, dydy, dydz, dydvx, dydvy, dydvz,
                 dzdx, dzdy, dzdz, dzdvx, dzdvy, dzdvz)
}

/* ---- Einstein-Infeld-Hoffman, reference src/forces.c:1288-1983 ------------- */
template <int KM, class BT>
__device__ void ab_force_eih(const AbEphem& E, const AbForceOpts& F, const BT& B, AbSysT<KM>& S,
                             double xo, double yo, double zo, double vxo, double vyo, double vzo,
                             double axo, double ayo, double azo) {
    const double over_C2 = E.over_c_squared;
    const int ns = F.gr_eih_sources;
    const double beta = 1.0;
    const double gamma = 1.0;
    const double pix = S.x[0][0], piy = S.x[0][1], piz = S.x[0][2];
    const double pivx = S.v[0][0], pivy = S.v[0][1], pivz = S.v[0][2];

    /* sum over the 11 planets of GM_k / r_ik: identical for every source j (src/forces.c:1400-1416) */
    double term0_sum = 0.0;
    {   /* rolled, the next planet's table entries requested one trip ahead */
        const double* const gm = B.gm;
        double nx = B.pos[0][0], ny = B.pos[0][1], nz = B.pos[0][2], ngm = gm[0];
#pragma unroll 1
        for (int k = 0; k < AB_NPLANETS; k++) {
            const double cx = nx, cy = ny, cz = nz, GMk = ngm;
            if (k + 1 < AB_NPLANETS) { nx = B.pos[k + 1][0]; ny = B.pos[k + 1][1]; nz = B.pos[k + 1][2]; ngm = gm[k + 1]; }
            const double dxik = pix + (xo - cx);
            const double dyik = piy + (yo - cy);
            const double dzik = piz + (zo - cz);
            const double rik2 = dxik * dxik + dyik * dyik + dzik * dzik;
            const double _rik = sqrt(rik2);
            term0_sum += GMk / _rik;
        }
    }

    {   /* real particle, src/forces.c:1319-1501 */
        double term7x_sum = 0.0, term7y_sum = 0.0, term7z_sum = 0.0;
        double term8x_sum = 0.0, term8y_sum = 0.0, term8z_sum = 0.0;
        for (int j = 0; j < ns; j++) {
            const double GMj = B.gm[j];
            const double xj = B.pos[j][0], yj = B.pos[j][1], zj = B.pos[j][2];
            const double vxj = B.vel[j][0], vyj = B.vel[j][1], vzj = B.vel[j][2];

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

            double term0 = term0_sum;
            double term1 = B.eih_term1[j];
            const double axj = B.eih_ar[j][0], ayj = B.eih_ar[j][1], azj = B.eih_ar[j][2];

            term0 *= -2 * (beta + gamma) * over_C2;
            term1 *= -(2 * beta - 1) * over_C2;

            const double rijdotaj = dxij * (axj - axo) + dyij * (ayj - ayo) + dzij * (azj - azo);
            const double term6 = -0.5 * over_C2 * rijdotaj;

            const double term8_fac = GMj / _rij * (3 + 4 * gamma) / 2;
            term8x_sum += term8_fac * axj;
            term8y_sum += term8_fac * ayj;
            term8z_sum += term8_fac * azj;

            const double factor = term0 + term1 + term2 + term3 + term4 + term5 + term6;

            S.a[0][0] += -prefacij * dxij * factor;
            S.a[0][1] += -prefacij * dyij * factor;
            S.a[0][2] += -prefacij * dzij * factor;
        }
        S.a[0][0] += term7x_sum * over_C2 + term8x_sum * over_C2;
        S.a[0][1] += term7y_sum * over_C2 + term8y_sum * over_C2;
        S.a[0][2] += term7z_sum * over_C2 + term8z_sum * over_C2;
    }

    if (S.nv() == 0) return;

    /* variational particles, src/forces.c:1508-1982 */
    double dterm0dx_sum = 0.0, dterm0dy_sum = 0.0, dterm0dz_sum = 0.0;
    for (int k = 0; k < AB_NPLANETS; k++) {
        const double GMk = B.gm[k];
        const double dxik = pix + (xo - B.pos[k][0]);
        const double dyik = piy + (yo - B.pos[k][1]);
        const double dzik = piz + (zo - B.pos[k][2]);
        const double rik2 = dxik * dxik + dyik * dyik + dzik * dzik;
        const double _rik = sqrt(rik2);
        dterm0dx_sum -= GMk / (_rik * _rik * _rik) * dxik;
        dterm0dy_sum -= GMk / (_rik * _rik * _rik) * dyik;
        dterm0dz_sum -= GMk / (_rik * _rik * _rik) * dzik;
    }

    double dxdx = 0.0, dxdy = 0.0, dxdz = 0.0, dxdvx = 0.0, dxdvy = 0.0, dxdvz = 0.0;
    double dydx = 0.0, dydy = 0.0, dydz = 0.0, dydvx = 0.0, dydvy = 0.0, dydvz = 0.0;
    double dzdx = 0.0, dzdy = 0.0, dzdz = 0.0, dzdvx = 0.0, dzdvy = 0.0, dzdvz = 0.0;

    double dterm7x_sumdx = 0.0, dterm7x_sumdy = 0.0, dterm7x_sumdz = 0.0, dterm7x_sumdvx = 0.0, dterm7x_sumdvy = 0.0, dterm7x_sumdvz = 0.0;
    double dterm7y_sumdx = 0.0, dterm7y_sumdy = 0.0, dterm7y_sumdz = 0.0, dterm7y_sumdvx = 0.0, dterm7y_sumdvy = 0.0, dterm7y_sumdvz = 0.0;
    double dterm7z_sumdx = 0.0, dterm7z_sumdy = 0.0, dterm7z_sumdz = 0.0, dterm7z_sumdvx = 0.0, dterm7z_sumdvy = 0.0, dterm7z_sumdvz = 0.0;
    double dterm8x_sumdx = 0.0, dterm8x_sumdy = 0.0, dterm8x_sumdz = 0.0;
    double dterm8y_sumdx = 0.0, dterm8y_sumdy = 0.0, dterm8y_sumdz = 0.0;
    double dterm8z_sumdx = 0.0, dterm8z_sumdy = 0.0, dterm8z_sumdz = 0.0;

    for (int j = 0; j < ns; j++) {
        const double GMj = B.gm[j];
        const double xj = B.pos[j][0], yj = B.pos[j][1], zj = B.pos[j][2];
        const double vxj = B.vel[j][0], vyj = B.vel[j][1], vzj = B.vel[j][2];

        const double dxij = pix + (xo - xj);
        const double dyij = piy + (yo - yj);
        const double dzij = piz + (zo - zj);
        const double rij2 = dxij * dxij + dyij * dyij + dzij * dzij;
        const double _rij = sqrt(rij2);
        const double prefacij = GMj / (_rij * _rij * _rij);

        const double dprefacijdx = -3.0 * GMj / (_rij * _rij * _rij * _rij * _rij) * dxij;
        const double dprefacijdy = -3.0 * GMj / (_rij * _rij * _rij * _rij * _rij) * dyij;
        const double dprefacijdz = -3.0 * GMj / (_rij * _rij * _rij * _rij * _rij) * dzij;

        const double vi2 = pivx * pivx + pivy * pivy + pivz * pivz;
        const double term2 = gamma * over_C2 * vi2;
        const double dterm2dvx = 2.0 * gamma * over_C2 * pivx;
        const double dterm2dvy = 2.0 * gamma * over_C2 * pivy;
        const double dterm2dvz = 2.0 * gamma * over_C2 * pivz;

        const double vj2 = (vxj - vxo) * (vxj - vxo) + (vyj - vyo) * (vyj - vyo) + (vzj - vzo) * (vzj - vzo);
        const double term3 = (1 + gamma) * over_C2 * vj2;

        const double vidotvj = pivx * (vxj - vxo) + pivy * (vyj - vyo) + pivz * (vzj - vzo);
        const double term4 = -2 * (1 + gamma) * over_C2 * vidotvj;
        const double dterm4dvx = -2 * (1 + gamma) * over_C2 * (vxj - vxo);
        const double dterm4dvy = -2 * (1 + gamma) * over_C2 * (vyj - vyo);
        const double dterm4dvz = -2 * (1 + gamma) * over_C2 * (vzj - vzo);

        const double rijdotvj = dxij * (vxj - vxo) + dyij * (vyj - vyo) + dzij * (vzj - vzo);
        const double term5 = -1.5 * over_C2 * (rijdotvj * rijdotvj) / (_rij * _rij);
        const double term5_fac = 3.0 * over_C2 * rijdotvj / _rij;
        const double dterm5dx = -term5_fac * ((vxj - vxo) / _rij - rijdotvj * dxij / (_rij * _rij * _rij));
        const double dterm5dy = -term5_fac * ((vyj - vyo) / _rij - rijdotvj * dyij / (_rij * _rij * _rij));
        const double dterm5dz = -term5_fac * ((vzj - vzo) / _rij - rijdotvj * dzij / (_rij * _rij * _rij));

        double fx = (2 + 2 * gamma) * pivx - (1 + 2 * gamma) * (vxj - vxo);
        double fy = (2 + 2 * gamma) * pivy - (1 + 2 * gamma) * (vyj - vyo);
        double fz = (2 + 2 * gamma) * pivz - (1 + 2 * gamma) * (vzj - vzo);
        double f = dxij * fx + dyij * fy + dzij * fz;

        double dfdx = fx, dfdy = fy, dfdz = fz;
        double dfdvx = dxij * (2 + 2 * gamma);
        double dfdvy = dyij * (2 + 2 * gamma);
        double dfdvz = dzij * (2 + 2 * gamma);

        const double wx = pivx - (vxj - vxo);
        const double wy = pivy - (vyj - vyo);
        const double wz = pivz - (vzj - vzo);

        dterm7x_sumdx += dprefacijdx * f * wx + prefacij * dfdx * wx;
        dterm7x_sumdy += dprefacijdy * f * wx + prefacij * dfdy * wx;
        dterm7x_sumdz += dprefacijdz * f * wx + prefacij * dfdz * wx;
        dterm7x_sumdvx += prefacij * dfdvx * wx + prefacij * f;
        dterm7x_sumdvy += prefacij * dfdvy * wx;
        dterm7x_sumdvz += prefacij * dfdvz * wx;

        dterm7y_sumdx += dprefacijdx * f * wy + prefacij * dfdx * wy;
        dterm7y_sumdy += dprefacijdy * f * wy + prefacij * dfdy * wy;
        dterm7y_sumdz += dprefacijdz * f * wy + prefacij * dfdz * wy;
        dterm7y_sumdvx += prefacij * dfdvx * wy;
        dterm7y_sumdvy += prefacij * dfdvy * wy + prefacij * f;
        dterm7y_sumdvz += prefacij * dfdvz * wy;

        dterm7z_sumdx += dprefacijdx * f * wz + prefacij * dfdx * wz;
        dterm7z_sumdy += dprefacijdy * f * wz + prefacij * dfdy * wz;
        dterm7z_sumdz += dprefacijdz * f * wz + prefacij * dfdz * wz;
        dterm7z_sumdvx += prefacij * dfdvx * wz;
        dterm7z_sumdvy += prefacij * dfdvy * wz;
        dterm7z_sumdvz += prefacij * dfdvz * wz + prefacij * f;

        double term0 = term0_sum;
        double dterm0dx = dterm0dx_sum, dterm0dy = dterm0dy_sum, dterm0dz = dterm0dz_sum;
        double term1 = B.eih_term1[j];
        const double dterm1dx = 0.0, dterm1dy = 0.0, dterm1dz = 0.0;
        const double dterm1dvx = 0.0, dterm1dvy = 0.0, dterm1dvz = 0.0;
        const double axj = B.eih_av[j][0], ayj = B.eih_av[j][1], azj = B.eih_av[j][2];

        term0 *= -2 * (beta + gamma) * over_C2;
        dterm0dx *= -2 * (beta + gamma) * over_C2;
        dterm0dy *= -2 * (beta + gamma) * over_C2;
        dterm0dz *= -2 * (beta + gamma) * over_C2;
        term1 *= -(2 * beta - 1) * over_C2;

        const double rijdotaj = dxij * (axj - axo) + dyij * (ayj - ayo) + dzij * (azj - azo);
        const double term6 = -0.5 * over_C2 * rijdotaj;
        const double dterm6dx = -0.5 * over_C2 * (axj - axo);
        const double dterm6dy = -0.5 * over_C2 * (ayj - ayo);
        const double dterm6dz = -0.5 * over_C2 * (azj - azo);

        dterm8x_sumdx += -GMj * axj / (_rij * _rij * _rij) * dxij * (3 + 4 * gamma) / 2;
        dterm8x_sumdy += -GMj * axj / (_rij * _rij * _rij) * dyij * (3 + 4 * gamma) / 2;
        dterm8x_sumdz += -GMj * axj / (_rij * _rij * _rij) * dzij * (3 + 4 * gamma) / 2;
        dterm8y_sumdx += -GMj * ayj / (_rij * _rij * _rij) * dxij * (3 + 4 * gamma) / 2;
        dterm8y_sumdy += -GMj * ayj / (_rij * _rij * _rij) * dyij * (3 + 4 * gamma) / 2;
        dterm8y_sumdz += -GMj * ayj / (_rij * _rij * _rij) * dzij * (3 + 4 * gamma) / 2;
        dterm8z_sumdx += -GMj * azj / (_rij * _rij * _rij) * dxij * (3 + 4 * gamma) / 2;
        dterm8z_sumdy += -GMj * azj / (_rij * _rij * _rij) * dyij * (3 + 4 * gamma) / 2;
        dterm8z_sumdz += -GMj * azj / (_rij * _rij * _rij) * dzij * (3 + 4 * gamma) / 2;

        double factor = term0 + term1 + term2 + term3 + term4 + term5 + term6;
        double dfactordx = dterm0dx + dterm1dx + dterm5dx + dterm6dx;
        double dfactordy = dterm0dy + dterm1dy + dterm5dy + dterm6dy;
        double dfactordz = dterm0dz + dterm1dz + dterm5dz + dterm6dz;
        double dfactordvx = dterm1dvx + dterm2dvx + dterm4dvx;
        double dfactordvy = dterm1dvy + dterm2dvy + dterm4dvy;
        double dfactordvz = dterm1dvz + dterm2dvz + dterm4dvz;

        dxdx += -dprefacijdx * dxij * factor - prefacij * factor - prefacij * dxij * dfactordx;
        dxdy += -dprefacijdy * dxij * factor - prefacij * dxij * dfactordy;
        dxdz += -dprefacijdz * dxij * factor - prefacij * dxij * dfactordz;
        dxdvx += -prefacij * dxij * dfactordvx;
        dxdvy += -prefacij * dxij * dfactordvy;
        dxdvz += -prefacij * dxij * dfactordvz;

        dydx += -dprefacijdx * dyij * factor - prefacij * dyij * dfactordx;
        dydy += -dprefacijdy * dyij * factor - prefacij * factor - prefacij * dyij * dfactordy;
        dydz += -dprefacijdz * dyij * factor - prefacij * dyij * dfactordz;
        dydvx += -prefacij * dyij * dfactordvx;
        dydvy += -prefacij * dyij * dfactordvy;
        dydvz += -prefacij * dyij * dfactordvz;

        dzdx += -dprefacijdx * dzij * factor - prefacij * dzij * dfactordx;
        dzdy += -dprefacijdy * dzij * factor - prefacij * dzij * dfactordy;
        dzdz += -dprefacijdz * dzij * factor - prefacij * factor - prefacij * dzij * dfactordz;
        dzdvx += -prefacij * dzij * dfactordvx;
        dzdvy += -prefacij * dzij * dfactordvy;
        dzdvz += -prefacij * dzij * dfactordvz;
    }

    dxdx += dterm7x_sumdx * over_C2 + dterm8x_sumdx * over_C2;
    dxdy += dterm7x_sumdy * over_C2 + dterm8x_sumdy * over_C2;
    dxdz += dterm7x_sumdz * over_C2 + dterm8x_sumdz * over_C2;
    dxdvx += dterm7x_sumdvx * over_C2;
    dxdvy += dterm7x_sumdvy * over_C2;
    dxdvz += dterm7x_sumdvz * over_C2;

    dydx += dterm7y_sumdx * over_C2 + dterm8y_sumdx * over_C2;
    dydy += dterm7y_sumdy * over_C2 + dterm8y_sumdy * over_C2;
    dydz += dterm7y_sumdz * over_C2 + dterm8y_sumdz * over_C2;
    dydvx += dterm7y_sumdvx * over_C2;
    dydvy += dterm7y_sumdvy * over_C2;
    dydvz += dterm7y_sumdvz * over_C2;

    dzdx += dterm7z_sumdx * over_C2 + dterm8z_sumdx * over_C2;
    dzdy += dterm7z_sumdy * over_C2 + dterm8z_sumdy * over_C2;
    dzdz += dterm7z_sumdz * over_C2 + dterm8z_sumdz * over_C2;
    dzdvx += dterm7z_sumdvx * over_C2;
    dzdvy += dterm7z_sumdvy * over_C2;
    dzdvz += dterm7z_sumdvz * over_C2;

    AB_APPLY_J36(S, dxdx, dxdy, dxdz, dxdvx, dxdvy, dxdvz, dydx, dydy, dydz, dydvx, dydvy, dydvz,
                 dzdx, dzdy, dzdz, dzdvx, dzdvy, dzdvz)
}

/* ---- direct Newtonian terms, reference src/forces.c:266-433 ---------------- */
/* Body visited at position k of the reference's loop (src/forces.c:281-306): asteroids first, then
 * Pluto, Moon, Mars, Mercury, Neptune, Uranus, Earth, Venus, Saturn, Jupiter, Sun -- the planet
 * sequence packed four bits each, so the index is arithmetic, not a (dependent) table load. */
__device__ __forceinline__ int ab_direct_body(int k, int ast_num) {
    return (k >= ast_num) ? (int)((0x0672389154AULL >> (4 * (k - ast_num))) & 0xFULL) : (k + AB_NPLANETS);
}

template <int KM, class BT>
__device__ void ab_force_direct(const AbEphem& E, const AbForceOpts& F, const BT& B, AbSysT<KM>& S,
                                double xo, double yo, double zo) {
    /* all asteroids of the small-body kernel, in file order, come first in the reference's loop; those beyond the
     * body table (sb441-n373: indices AB_MAX_BODIES ...) are evaluated here, at the table's time */
    const int ast_num = E.n_ast + E.n_ast_x;
    const int nb = AB_NPLANETS + ast_num;
    const double px = S.x[0][0], py = S.x[0][1], pz = S.x[0][2];
    const int fmask = F.forces;
    /* the running sums stay in registers; the additions happen in the reference's sequence */
    double acx = S.a[0][0], acy = S.a[0][1], acz = S.a[0][2];
    /* the table entries of the next body are requested while the current body is being worked on */
    const double* const gm = B.gm;       /* read once: the table is not written while the forces are evaluated */
    int i_next = ab_direct_body(0, ast_num);
    double bx = B.pos[i_next][0], by = B.pos[i_next][1], bz = B.pos[i_next][2], bgm = gm[i_next];
#pragma unroll 1      /* rolled on purpose: unrolled x2 / x3 it is 3-6 % slower (instruction fetch), profiles/README.md */
    for (int k = 0; k < nb; k++) {
        const int i = i_next;
        const double GM = bgm;
        const double cx = bx, cy = by, cz = bz;
        if (k + 1 < nb) {
            i_next = ab_direct_body(k + 1, ast_num);
            if (i_next < AB_MAX_BODIES) {
                bx = B.pos[i_next][0]; by = B.pos[i_next][1]; bz = B.pos[i_next][2]; bgm = gm[i_next];
            } else {
                double c3[3];
                ab_extra_asteroid(E, i_next - AB_NPLANETS, B.t, B.pos[0][0], B.pos[0][1], B.pos[0][2], &bgm, c3);
                bx = c3[0]; by = c3[1]; bz = c3[2];
            }
        }
        const double dx = px + (xo - cx);
        const double dy = py + (yo - cy);
        const double dz = pz + (zo - cz);
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double _r = sqrt(r2);
        bool on = true;
        if (i == 0 && !(fmask & 0x01)) on = false;
        if (i > 0 && i < AB_NPLANETS && !(fmask & 0x02)) on = false;
        if (i >= AB_NPLANETS && !(fmask & 0x04)) on = false;
        if (on) {
            const double prefac = GM / (_r * _r * _r);
            acx -= prefac * dx;
            acy -= prefac * dy;
            acz -= prefac * dz;
        }
        if (S.nv() > 0) {
            /* no force-mask check here, as in the reference (src/forces.c:359) */
            const double r3inv = 1. / (r2 * _r);
            const double r5inv = 3. * r3inv / r2;
            const double dxdx = dx * dx * r5inv - r3inv;
            const double dydy = dy * dy * r5inv - r3inv;
            const double dzdz = dz * dz * r5inv - r3inv;
            const double dxdy = dx * dy * r5inv;
            const double dxdz = dx * dz * r5inv;
            const double dydz = dy * dz * r5inv;
            for (int vv = 1; vv <= S.nv(); vv++) {
                const double ddx = S.x[vv][0], ddy = S.x[vv][1], ddz = S.x[vv][2];
                const double dax = ddx * dxdx + ddy * dxdy + ddz * dxdz;
                const double day = ddx * dxdy + ddy * dydy + ddz * dydz;
                const double daz = ddx * dxdz + ddy * dydz + ddz * dzdz;
                S.a[vv][0] += GM * dax;
                S.a[vv][1] += GM * day;
                S.a[vv][2] += GM * daz;
            }
        }
    }
    S.a[0][0] = acx; S.a[0][1] = acy; S.a[0][2] = acz;
}

/* ---- dispatcher, reference src/forces.c:49-173 ----------------------------- */
/* S.a must be zero on entry (REBOUND zeroes accelerations before the plug-in runs). */
template <int KM, class BT>
__device__ void ab_forces(const AbEphem& E, const AbForceOpts& F, const BT& B, AbSysT<KM>& S) {
    double xo = 0.0, yo = 0.0, zo = 0.0, vxo = 0.0, vyo = 0.0, vzo = 0.0, axo = 0.0, ayo = 0.0, azo = 0.0;
    if (F.geocentric == 1) {
        xo = B.pos[3][0]; yo = B.pos[3][1]; zo = B.pos[3][2];
        vxo = B.vel[3][0]; vyo = B.vel[3][1]; vzo = B.vel[3][2];
        axo = B.earth_acc[0]; ayo = B.earth_acc[1]; azo = B.earth_acc[2];
    }
    if (F.forces & 0x08) ab_force_nongrav(F, B, S, xo, yo, zo, vxo, vyo, vzo);
    if (F.forces & 0x10) ab_force_earth_harmonics(E, F, B, S, xo, yo, zo);
    if (F.forces & 0x20) ab_force_solar_j2(E, F, B, S, xo, yo, zo);
    if (F.forces & 0x40) ab_force_eih(E, F, B, S, xo, yo, zo, vxo, vyo, vzo, axo, ayo, azo);
    if (F.forces & 0x100) ab_force_potential_gr(E, B, S, xo, yo, zo);
    if (F.forces & 0x80) ab_force_simple_gr(E, B, S, xo, yo, zo, vxo, vyo, vzo);
    if (F.forces & (0x01 | 0x02 | 0x04)) ab_force_direct(E, F, B, S, xo, yo, zo);
    if (F.geocentric == 1) {
        S.a[0][0] -= axo; S.a[0][1] -= ayo; S.a[0][2] -= azo;
    }
}

}  // namespace AB_NS
#endif
