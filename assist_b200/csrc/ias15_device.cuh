/*
 * ias15_device.cuh -- IAS15 (15th-order Gauss-Radau predictor-corrector) per system.
 *
 * Device restatement of the REBOUND 4.x stepper that ASSIST drives (REBOUND is an
 * un-vendored third-party dependency of the reference; see oracle/reb_shim.c for
 * the CPU restatement these kernels are checked against, and SURVEY.md App. A).
 * Anchors in the reference itself: adaptive_mode = 1 and velocity-dependent forces
 * (src/assist.c:445-446), the b-series divisors (src/assist.c:562-580), last_state
 * + br as the dense-output state (src/assist.c:674-677, 754-758).
 *
 * State lives in HBM as structure-of-arrays over systems (coalesced across the
 * threads of a warp); see AbBatch in device_types.h.  Expressions keep the CPU
 * restatement's operation order so a no-FMA build follows it bit for bit.
 */
#ifndef AB_IAS15_DEVICE_CUH
#define AB_IAS15_DEVICE_CUH

#include "device_types.h"
#include "ias15_constants.h"
#include "fp_device.cuh"
#include "forces_device.cuh"

namespace AB_NS {

static __constant__ double c_h[8];
static __constant__ double c_rr[28];
static __constant__ double c_rri[28];   /* RN(1/rr[k]) for exact division by rr[k] (fp_device.cuh) */
static __constant__ double c_c[21];
static __constant__ double c_d[21];

#define AB1(arr, k) (arr)[(long long)(k) * n + i]
#define AB7(arr, j, k) (arr)[((long long)(j) * C + (k)) * n + i]

__device__ __forceinline__ void ab_add_cs(double& p, double& csp, double inp) {
    const double y = inp - csp;
    const double t = p + y;
    csp = (t - p) - y;
    p = t;
}

__device__ __forceinline__ bool ab_isnormal(double x) {
    const double ax = fabs(x);
    return ax >= 2.2250738585072014e-308 && ax <= 1.7976931348623157e308;
}

/* REBOUND's machine independent seventh root. */
__device__ double ab_sqrt7(double a) {
    double scale = 1;
    while (a < 1e-7 && ab_isnormal(a)) { scale *= 0.1; a *= 1e7; }
    while (a > 1e2 && ab_isnormal(a)) { scale *= 10; a *= 1e-7; }
    double x = 1.;
    for (int k = 0; k < 20; k++) {
        double x6 = x * x * x * x * x * x;
        const double xn = x + AB_DIVK(a / x6 - x, 7.);
        if (xn == x) break;      /* a fixed point: the remaining iterations of REBOUND's 20 would reproduce it */
        x = xn;
    }
    return x * scale;
}

/* Load particles of system i into registers. */
template <int KM>
__device__ __forceinline__ void ab_load_sys(const AbBatch& Bt, long long i, AbSysT<KM>& S) {
    const long long n = Bt.n;
    S.nv_ = Bt.nv[i];
    for (int j = 0; j <= S.nv(); j++) {
        for (int c = 0; c < 3; c++) {
            S.x[j][c] = AB1(Bt.pos, 3 * j + c);
            S.v[j][c] = AB1(Bt.vel, 3 * j + c);
            S.prm[j][c] = Bt.has_params ? AB1(Bt.prm, 3 * j + c) : 0.0;
        }
    }
}

template <int KM>
__device__ __forceinline__ void ab_zero_acc(AbSysT<KM>& S) {
    for (int j = 0; j <= S.nv(); j++) { S.a[j][0] = 0.0; S.a[j][1] = 0.0; S.a[j][2] = 0.0; }
}

/* Start of a reb_simulation_step: accelerations at the current state become a0 and
 * the dense-output anchor (last_state) is taken. */
template <int KM>
__device__ __forceinline__ void ab_store_a0(const AbBatch& Bt, long long i, const AbSysT<KM>& S) {
    const long long n = Bt.n;
    for (int j = 0; j <= S.nv(); j++) {
        for (int c = 0; c < 3; c++) {
            const int k = 3 * j + c;
            AB1(Bt.acc, k) = S.a[j][c];
            AB1(Bt.ls_pos, k) = S.x[j][c];
            AB1(Bt.ls_vel, k) = S.v[j][c];
            AB1(Bt.ls_acc, k) = S.a[j][c];
        }
    }
}

/* Beginning of a step attempt: x0,v0,a0 from the particles, csb cleared, g from b. */
__device__ void ab_attempt_begin(const AbBatch& Bt, long long i, int nv) {
    const long long n = Bt.n;
    const int C = Bt.C;
    const int Ca = 3 * (1 + nv);
    for (int k = 0; k < Ca; k++) {
        AB1(Bt.x0, k) = AB1(Bt.pos, k);
        AB1(Bt.v0, k) = AB1(Bt.vel, k);
        AB1(Bt.a0, k) = AB1(Bt.acc, k);
        const double b0 = AB7(Bt.b, 0, k), b1 = AB7(Bt.b, 1, k), b2 = AB7(Bt.b, 2, k), b3 = AB7(Bt.b, 3, k);
        const double b4 = AB7(Bt.b, 4, k), b5 = AB7(Bt.b, 5, k), b6 = AB7(Bt.b, 6, k);
        for (int j = 0; j < 7; j++) AB7(Bt.csb, j, k) = 0.;
        AB7(Bt.g, 0, k) = b6 * c_d[15] + b5 * c_d[10] + b4 * c_d[6] + b3 * c_d[3] + b2 * c_d[1] + b1 * c_d[0] + b0;
        AB7(Bt.g, 1, k) = b6 * c_d[16] + b5 * c_d[11] + b4 * c_d[7] + b3 * c_d[4] + b2 * c_d[2] + b1;
        AB7(Bt.g, 2, k) = b6 * c_d[17] + b5 * c_d[12] + b4 * c_d[8] + b3 * c_d[5] + b2;
        AB7(Bt.g, 3, k) = b6 * c_d[18] + b5 * c_d[13] + b4 * c_d[9] + b3;
        AB7(Bt.g, 4, k) = b6 * c_d[19] + b5 * c_d[14] + b4;
        AB7(Bt.g, 5, k) = b6 * c_d[20] + b5;
        AB7(Bt.g, 6, k) = b6;
    }
}

/* Predict positions and velocities of every body of system i at node n. */
template <int KM>
__device__ void ab_predict(const AbBatch& Bt, long long i, int nn, double dt, AbSysT<KM>& S) {
    const long long n = Bt.n;
    const int C = Bt.C;
    const double h = c_h[nn];
    for (int j = 0; j <= S.nv(); j++) {
        for (int c = 0; c < 3; c++) {
            const int k = 3 * j + c;
            const double b0 = AB7(Bt.b, 0, k), b1 = AB7(Bt.b, 1, k), b2 = AB7(Bt.b, 2, k), b3 = AB7(Bt.b, 3, k);
            const double b4 = AB7(Bt.b, 4, k), b5 = AB7(Bt.b, 5, k), b6 = AB7(Bt.b, 6, k);
            const double x0 = AB1(Bt.x0, k), v0 = AB1(Bt.v0, k), a0 = AB1(Bt.a0, k);
            const double csx = AB1(Bt.csx, k), csv = AB1(Bt.csv, k);
            /* position series, nested exactly as REBOUND writes it:
             * ((((((((b6*7h/9 + b5)*3h/4 + b4)*5h/7 + b3)*2h/3 + b2)*3h/5 + b1)*h/2 + b0)*h/3 + a0)*dt*h/2 + v0)*dt*h */
            double px = AB_DIVK(b6 * 7. * h, 9.) + b5;
            px = px * 3. * h / 4. + b4;
            px = AB_DIVK(px * 5. * h, 7.) + b3;
            px = AB_DIVK(px * 2. * h, 3.) + b2;
            px = AB_DIVK(px * 3. * h, 5.) + b1;
            px = px * h / 2. + b0;
            px = AB_DIVK(px * h, 3.) + a0;
            px = px * dt * h / 2. + v0;
            const double xk = -csx + px * dt * h;
            S.x[j][c] = xk + x0;
            /* velocity series: (((((((b6*7h/8 + b5)*6h/7 + b4)*5h/6 + b3)*4h/5 + b2)*3h/4 + b1)*2h/3 + b0)*h/2 + a0)*dt*h */
            double pv = b6 * 7. * h / 8. + b5;
            pv = AB_DIVK(pv * 6. * h, 7.) + b4;
            pv = AB_DIVK(pv * 5. * h, 6.) + b3;
            pv = AB_DIVK(pv * 4. * h, 5.) + b2;
            pv = pv * 3. * h / 4. + b1;
            pv = AB_DIVK(pv * 2. * h, 3.) + b0;
            pv = pv * h / 2. + a0;
            const double vk = -csv + pv * dt * h;
            S.v[j][c] = vk + v0;
        }
    }
}

/* Improve g and b from the accelerations at node n.  At node 7 also returns the
 * largest |a| and |change of b6| over the components (convergence monitor). */
template <int KM>
__device__ void ab_update_gb(const AbBatch& Bt, long long i, int nn, const AbSysT<KM>& S, double& maxak, double& maxb6) {
    const long long n = Bt.n;
    const int C = Bt.C;
    for (int j = 0; j <= S.nv(); j++) {
        for (int c = 0; c < 3; c++) {
            const int k = 3 * j + c;
            const double at = S.a[j][c];
            /* add_cs(&gk, &gk_cs = 0, -a0) reduces to one rounded subtraction */
            const double gk = at + (-AB1(Bt.a0, k));
            double tmp;
            double g0, g1, g2, g3, g4, g5;
            switch (nn) {
                case 1: {
                    tmp = AB7(Bt.g, 0, k);
                    const double gn = ab_divc(gk, c_rr[0], c_rri[0]);
                    AB7(Bt.g, 0, k) = gn;
                    double b0 = AB7(Bt.b, 0, k), cs0 = AB7(Bt.csb, 0, k);
                    ab_add_cs(b0, cs0, gn - tmp);
                    AB7(Bt.b, 0, k) = b0; AB7(Bt.csb, 0, k) = cs0;
                } break;
                case 2: {
                    tmp = AB7(Bt.g, 1, k);
                    g0 = AB7(Bt.g, 0, k);
                    const double gn = ab_divc(ab_divc(gk, c_rr[1], c_rri[1]) - g0, c_rr[2], c_rri[2]);
                    AB7(Bt.g, 1, k) = gn;
                    tmp = gn - tmp;
                    double b0 = AB7(Bt.b, 0, k), cs0 = AB7(Bt.csb, 0, k);
                    double b1 = AB7(Bt.b, 1, k), cs1 = AB7(Bt.csb, 1, k);
                    ab_add_cs(b0, cs0, tmp * c_c[0]);
                    ab_add_cs(b1, cs1, tmp);
                    AB7(Bt.b, 0, k) = b0; AB7(Bt.csb, 0, k) = cs0;
                    AB7(Bt.b, 1, k) = b1; AB7(Bt.csb, 1, k) = cs1;
                } break;
                case 3: {
                    tmp = AB7(Bt.g, 2, k);
                    g0 = AB7(Bt.g, 0, k); g1 = AB7(Bt.g, 1, k);
                    const double gn = ab_divc(ab_divc(ab_divc(gk, c_rr[3], c_rri[3]) - g0, c_rr[4], c_rri[4]) - g1, c_rr[5], c_rri[5]);
                    AB7(Bt.g, 2, k) = gn;
                    tmp = gn - tmp;
                    double b0 = AB7(Bt.b, 0, k), cs0 = AB7(Bt.csb, 0, k);
                    double b1 = AB7(Bt.b, 1, k), cs1 = AB7(Bt.csb, 1, k);
                    double b2 = AB7(Bt.b, 2, k), cs2 = AB7(Bt.csb, 2, k);
                    ab_add_cs(b0, cs0, tmp * c_c[1]);
                    ab_add_cs(b1, cs1, tmp * c_c[2]);
                    ab_add_cs(b2, cs2, tmp);
                    AB7(Bt.b, 0, k) = b0; AB7(Bt.csb, 0, k) = cs0;
                    AB7(Bt.b, 1, k) = b1; AB7(Bt.csb, 1, k) = cs1;
                    AB7(Bt.b, 2, k) = b2; AB7(Bt.csb, 2, k) = cs2;
                } break;
                case 4: {
                    tmp = AB7(Bt.g, 3, k);
                    g0 = AB7(Bt.g, 0, k); g1 = AB7(Bt.g, 1, k); g2 = AB7(Bt.g, 2, k);
                    const double gn = ab_divc(ab_divc(ab_divc(ab_divc(gk, c_rr[6], c_rri[6]) - g0, c_rr[7], c_rri[7]) - g1, c_rr[8], c_rri[8]) - g2, c_rr[9], c_rri[9]);
                    AB7(Bt.g, 3, k) = gn;
                    tmp = gn - tmp;
                    double b0 = AB7(Bt.b, 0, k), cs0 = AB7(Bt.csb, 0, k);
                    double b1 = AB7(Bt.b, 1, k), cs1 = AB7(Bt.csb, 1, k);
                    double b2 = AB7(Bt.b, 2, k), cs2 = AB7(Bt.csb, 2, k);
                    double b3 = AB7(Bt.b, 3, k), cs3 = AB7(Bt.csb, 3, k);
                    ab_add_cs(b0, cs0, tmp * c_c[3]);
                    ab_add_cs(b1, cs1, tmp * c_c[4]);
                    ab_add_cs(b2, cs2, tmp * c_c[5]);
                    ab_add_cs(b3, cs3, tmp);
                    AB7(Bt.b, 0, k) = b0; AB7(Bt.csb, 0, k) = cs0;
                    AB7(Bt.b, 1, k) = b1; AB7(Bt.csb, 1, k) = cs1;
                    AB7(Bt.b, 2, k) = b2; AB7(Bt.csb, 2, k) = cs2;
                    AB7(Bt.b, 3, k) = b3; AB7(Bt.csb, 3, k) = cs3;
                } break;
                case 5: {
                    tmp = AB7(Bt.g, 4, k);
                    g0 = AB7(Bt.g, 0, k); g1 = AB7(Bt.g, 1, k); g2 = AB7(Bt.g, 2, k); g3 = AB7(Bt.g, 3, k);
                    const double gn = ab_divc(ab_divc(ab_divc(ab_divc(ab_divc(gk, c_rr[10], c_rri[10]) - g0, c_rr[11], c_rri[11]) - g1, c_rr[12], c_rri[12]) - g2, c_rr[13], c_rri[13]) - g3, c_rr[14], c_rri[14]);
                    AB7(Bt.g, 4, k) = gn;
                    tmp = gn - tmp;
                    double b0 = AB7(Bt.b, 0, k), cs0 = AB7(Bt.csb, 0, k);
                    double b1 = AB7(Bt.b, 1, k), cs1 = AB7(Bt.csb, 1, k);
                    double b2 = AB7(Bt.b, 2, k), cs2 = AB7(Bt.csb, 2, k);
                    double b3 = AB7(Bt.b, 3, k), cs3 = AB7(Bt.csb, 3, k);
                    double b4 = AB7(Bt.b, 4, k), cs4 = AB7(Bt.csb, 4, k);
                    ab_add_cs(b0, cs0, tmp * c_c[6]);
                    ab_add_cs(b1, cs1, tmp * c_c[7]);
                    ab_add_cs(b2, cs2, tmp * c_c[8]);
                    ab_add_cs(b3, cs3, tmp * c_c[9]);
                    ab_add_cs(b4, cs4, tmp);
                    AB7(Bt.b, 0, k) = b0; AB7(Bt.csb, 0, k) = cs0;
                    AB7(Bt.b, 1, k) = b1; AB7(Bt.csb, 1, k) = cs1;
                    AB7(Bt.b, 2, k) = b2; AB7(Bt.csb, 2, k) = cs2;
                    AB7(Bt.b, 3, k) = b3; AB7(Bt.csb, 3, k) = cs3;
                    AB7(Bt.b, 4, k) = b4; AB7(Bt.csb, 4, k) = cs4;
                } break;
                case 6: {
                    tmp = AB7(Bt.g, 5, k);
                    g0 = AB7(Bt.g, 0, k); g1 = AB7(Bt.g, 1, k); g2 = AB7(Bt.g, 2, k); g3 = AB7(Bt.g, 3, k); g4 = AB7(Bt.g, 4, k);
                    const double gn = ab_divc(ab_divc(ab_divc(ab_divc(ab_divc(ab_divc(gk, c_rr[15], c_rri[15]) - g0, c_rr[16], c_rri[16]) - g1, c_rr[17], c_rri[17]) - g2, c_rr[18], c_rri[18]) - g3, c_rr[19], c_rri[19]) - g4, c_rr[20], c_rri[20]);
                    AB7(Bt.g, 5, k) = gn;
                    tmp = gn - tmp;
                    double b0 = AB7(Bt.b, 0, k), cs0 = AB7(Bt.csb, 0, k);
                    double b1 = AB7(Bt.b, 1, k), cs1 = AB7(Bt.csb, 1, k);
                    double b2 = AB7(Bt.b, 2, k), cs2 = AB7(Bt.csb, 2, k);
                    double b3 = AB7(Bt.b, 3, k), cs3 = AB7(Bt.csb, 3, k);
                    double b4 = AB7(Bt.b, 4, k), cs4 = AB7(Bt.csb, 4, k);
                    double b5 = AB7(Bt.b, 5, k), cs5 = AB7(Bt.csb, 5, k);
                    ab_add_cs(b0, cs0, tmp * c_c[10]);
                    ab_add_cs(b1, cs1, tmp * c_c[11]);
                    ab_add_cs(b2, cs2, tmp * c_c[12]);
                    ab_add_cs(b3, cs3, tmp * c_c[13]);
                    ab_add_cs(b4, cs4, tmp * c_c[14]);
                    ab_add_cs(b5, cs5, tmp);
                    AB7(Bt.b, 0, k) = b0; AB7(Bt.csb, 0, k) = cs0;
                    AB7(Bt.b, 1, k) = b1; AB7(Bt.csb, 1, k) = cs1;
                    AB7(Bt.b, 2, k) = b2; AB7(Bt.csb, 2, k) = cs2;
                    AB7(Bt.b, 3, k) = b3; AB7(Bt.csb, 3, k) = cs3;
                    AB7(Bt.b, 4, k) = b4; AB7(Bt.csb, 4, k) = cs4;
                    AB7(Bt.b, 5, k) = b5; AB7(Bt.csb, 5, k) = cs5;
                } break;
                default: {   /* node 7 */
                    tmp = AB7(Bt.g, 6, k);
                    g0 = AB7(Bt.g, 0, k); g1 = AB7(Bt.g, 1, k); g2 = AB7(Bt.g, 2, k); g3 = AB7(Bt.g, 3, k);
                    g4 = AB7(Bt.g, 4, k); g5 = AB7(Bt.g, 5, k);
                    const double gn = ab_divc(ab_divc(ab_divc(ab_divc(ab_divc(ab_divc(ab_divc(gk, c_rr[21], c_rri[21]) - g0, c_rr[22], c_rri[22]) - g1, c_rr[23], c_rri[23]) - g2, c_rr[24], c_rri[24]) - g3, c_rr[25], c_rri[25]) - g4, c_rr[26], c_rri[26]) - g5, c_rr[27], c_rri[27]);
                    AB7(Bt.g, 6, k) = gn;
                    tmp = gn - tmp;
                    double b0 = AB7(Bt.b, 0, k), cs0 = AB7(Bt.csb, 0, k);
                    double b1 = AB7(Bt.b, 1, k), cs1 = AB7(Bt.csb, 1, k);
                    double b2 = AB7(Bt.b, 2, k), cs2 = AB7(Bt.csb, 2, k);
                    double b3 = AB7(Bt.b, 3, k), cs3 = AB7(Bt.csb, 3, k);
                    double b4 = AB7(Bt.b, 4, k), cs4 = AB7(Bt.csb, 4, k);
                    double b5 = AB7(Bt.b, 5, k), cs5 = AB7(Bt.csb, 5, k);
                    double b6 = AB7(Bt.b, 6, k), cs6 = AB7(Bt.csb, 6, k);
                    ab_add_cs(b0, cs0, tmp * c_c[15]);
                    ab_add_cs(b1, cs1, tmp * c_c[16]);
                    ab_add_cs(b2, cs2, tmp * c_c[17]);
                    ab_add_cs(b3, cs3, tmp * c_c[18]);
                    ab_add_cs(b4, cs4, tmp * c_c[19]);
                    ab_add_cs(b5, cs5, tmp * c_c[20]);
                    ab_add_cs(b6, cs6, tmp);
                    AB7(Bt.b, 0, k) = b0; AB7(Bt.csb, 0, k) = cs0;
                    AB7(Bt.b, 1, k) = b1; AB7(Bt.csb, 1, k) = cs1;
                    AB7(Bt.b, 2, k) = b2; AB7(Bt.csb, 2, k) = cs2;
                    AB7(Bt.b, 3, k) = b3; AB7(Bt.csb, 3, k) = cs3;
                    AB7(Bt.b, 4, k) = b4; AB7(Bt.csb, 4, k) = cs4;
                    AB7(Bt.b, 5, k) = b5; AB7(Bt.csb, 5, k) = cs5;
                    AB7(Bt.b, 6, k) = b6; AB7(Bt.csb, 6, k) = cs6;
                    const double ak = fabs(at);
                    if (ab_isnormal(ak) && ak > maxak) maxak = ak;
                    const double b6ktmp = fabs(tmp);
                    if (ab_isnormal(b6ktmp) && b6ktmp > maxb6) maxb6 = b6ktmp;
                } break;
            }
        }
    }
}

/* Step-size monitor of one system: largest |a| and |b6| of the REAL particle,
 * unless its acceleration is slowly varying.  S holds the node-7 prediction. */
template <int KM>
__device__ void ab_dt_monitor(const AbBatch& Bt, long long i, const AbSysT<KM>& S, double dt, double& maxa, double& maxj) {
    const long long n = Bt.n;
    const int C = Bt.C;
    const double v2 = S.v[0][0] * S.v[0][0] + S.v[0][1] * S.v[0][1] + S.v[0][2] * S.v[0][2];
    const double x2 = S.x[0][0] * S.x[0][0] + S.x[0][1] * S.x[0][1] + S.x[0][2] * S.x[0][2];
    if (fabs(v2 * dt * dt / x2) < 1e-16) return;
    for (int k = 0; k < 3; k++) {
        const double ak = fabs(S.a[0][k]);
        if (ab_isnormal(ak) && ak > maxa) maxa = ak;
        const double b6k = fabs(AB7(Bt.b, 6, k));
        if (ab_isnormal(b6k) && b6k > maxj) maxj = b6k;
    }
}

/* New step size from the error estimate (adaptive_mode 1). */
__device__ double ab_dt_new(double epsilon, double min_dt, double maxa, double maxj, double dt_done) {
    const double integrator_error = maxj / maxa;
    double dt_new;
    if (ab_isnormal(integrator_error)) dt_new = ab_sqrt7(epsilon / integrator_error) * dt_done;
    else dt_new = dt_done / 0.25;
    if (fabs(dt_new) < min_dt) dt_new = copysign(min_dt, dt_new);
    return dt_new;
}

/* b,e <- prediction for a step `ratio` times as long, from (src_e, src_b). */
__device__ void ab_predict_next(const AbBatch& Bt, long long i, int nv, double ratio, const double* src_e, const double* src_b) {
    const long long n = Bt.n;
    const int C = Bt.C;
    const int Ca = 3 * (1 + nv);
    if (ratio > 20.) {
        for (int k = 0; k < Ca; k++)
            for (int j = 0; j < 7; j++) { AB7(Bt.e, j, k) = 0.; AB7(Bt.b, j, k) = 0.; }
        return;
    }
    const double q1 = ratio;
    const double q2 = q1 * q1;
    const double q3 = q1 * q2;
    const double q4 = q2 * q2;
    const double q5 = q2 * q3;
    const double q6 = q3 * q3;
    const double q7 = q3 * q4;
    for (int k = 0; k < Ca; k++) {
        const double _b0 = AB7(src_b, 0, k), _b1 = AB7(src_b, 1, k), _b2 = AB7(src_b, 2, k), _b3 = AB7(src_b, 3, k);
        const double _b4 = AB7(src_b, 4, k), _b5 = AB7(src_b, 5, k), _b6 = AB7(src_b, 6, k);
        const double be0 = _b0 - AB7(src_e, 0, k);
        const double be1 = _b1 - AB7(src_e, 1, k);
        const double be2 = _b2 - AB7(src_e, 2, k);
        const double be3 = _b3 - AB7(src_e, 3, k);
        const double be4 = _b4 - AB7(src_e, 4, k);
        const double be5 = _b5 - AB7(src_e, 5, k);
        const double be6 = _b6 - AB7(src_e, 6, k);
        const double e0 = q1 * (_b6 * 7.0 + _b5 * 6.0 + _b4 * 5.0 + _b3 * 4.0 + _b2 * 3.0 + _b1 * 2.0 + _b0);
        const double e1 = q2 * (_b6 * 21.0 + _b5 * 15.0 + _b4 * 10.0 + _b3 * 6.0 + _b2 * 3.0 + _b1);
        const double e2 = q3 * (_b6 * 35.0 + _b5 * 20.0 + _b4 * 10.0 + _b3 * 4.0 + _b2);
        const double e3 = q4 * (_b6 * 35.0 + _b5 * 15.0 + _b4 * 5.0 + _b3);
        const double e4 = q5 * (_b6 * 21.0 + _b5 * 6.0 + _b4);
        const double e5 = q6 * (_b6 * 7.0 + _b5);
        const double e6 = q7 * _b6;
        AB7(Bt.e, 0, k) = e0; AB7(Bt.e, 1, k) = e1; AB7(Bt.e, 2, k) = e2; AB7(Bt.e, 3, k) = e3;
        AB7(Bt.e, 4, k) = e4; AB7(Bt.e, 5, k) = e5; AB7(Bt.e, 6, k) = e6;
        AB7(Bt.b, 0, k) = e0 + be0; AB7(Bt.b, 1, k) = e1 + be1; AB7(Bt.b, 2, k) = e2 + be2; AB7(Bt.b, 3, k) = e3 + be3;
        AB7(Bt.b, 4, k) = e4 + be4; AB7(Bt.b, 5, k) = e5 + be5; AB7(Bt.b, 6, k) = e6 + be6;
    }
}

/* Rejected attempt: particles back to the start of the step. */
__device__ void ab_restore(const AbBatch& Bt, long long i, int nv) {
    const long long n = Bt.n;
    const int Ca = 3 * (1 + nv);
    for (int k = 0; k < Ca; k++) {
        AB1(Bt.pos, k) = AB1(Bt.x0, k);
        AB1(Bt.vel, k) = AB1(Bt.v0, k);
        AB1(Bt.acc, k) = AB1(Bt.a0, k);
    }
}

/* Accepted attempt: advance x0,v0 with compensated sums, publish particles, er<-e, br<-b. */
__device__ void ab_advance(const AbBatch& Bt, long long i, int nv, double dt_done) {
    const long long n = Bt.n;
    const int C = Bt.C;
    const int Ca = 3 * (1 + nv);
    for (int k = 0; k < Ca; k++) {
        const double b0 = AB7(Bt.b, 0, k), b1 = AB7(Bt.b, 1, k), b2 = AB7(Bt.b, 2, k), b3 = AB7(Bt.b, 3, k);
        const double b4 = AB7(Bt.b, 4, k), b5 = AB7(Bt.b, 5, k), b6 = AB7(Bt.b, 6, k);
        double x0 = AB1(Bt.x0, k), v0 = AB1(Bt.v0, k);
        const double a0 = AB1(Bt.a0, k);
        double csx = AB1(Bt.csx, k), csv = AB1(Bt.csv, k);
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
        AB1(Bt.x0, k) = x0; AB1(Bt.v0, k) = v0; AB1(Bt.csx, k) = csx; AB1(Bt.csv, k) = csv;
        AB1(Bt.pos, k) = x0; AB1(Bt.vel, k) = v0;
        for (int j = 0; j < 7; j++) {
            AB7(Bt.er, j, k) = AB7(Bt.e, j, k);
            AB7(Bt.br, j, k) = AB7(Bt.b, j, k);
        }
    }
}

/* Dense output inside the last completed step, reference src/assist.c:556-597. */
__device__ void ab_interpolate(const AbBatch& Bt, long long i, int nv, double dt_last_done, double h,
                               double* __restrict__ out /* [K][6] for this system */) {
    const long long n = Bt.n;
    const int C = Bt.C;
    double s[9], sv[8];
    s[0] = dt_last_done * h;
    s[1] = s[0] * s[0] / 2.;
    s[2] = AB_DIVK(s[1] * h, 3.);
    s[3] = s[2] * h / 2.;
    s[4] = AB_DIVK(3. * s[3] * h, 5.);
    s[5] = AB_DIVK(2. * s[4] * h, 3.);
    s[6] = AB_DIVK(5. * s[5] * h, 7.);
    s[7] = 3. * s[6] * h / 4.;
    s[8] = AB_DIVK(7. * s[7] * h, 9.);
    sv[0] = dt_last_done * h;
    sv[1] = sv[0] * h / 2.;
    sv[2] = AB_DIVK(2. * sv[1] * h, 3.);
    sv[3] = 3. * sv[2] * h / 4.;
    sv[4] = AB_DIVK(4. * sv[3] * h, 5.);
    sv[5] = AB_DIVK(5. * sv[4] * h, 6.);
    sv[6] = AB_DIVK(6. * sv[5] * h, 7.);
    sv[7] = 7. * sv[6] * h / 8.;
    for (int j = 0; j <= nv; j++) {
        for (int c = 0; c < 3; c++) {
            const int k = 3 * j + c;
            const double b0 = AB7(Bt.br, 0, k), b1 = AB7(Bt.br, 1, k), b2 = AB7(Bt.br, 2, k), b3 = AB7(Bt.br, 3, k);
            const double b4 = AB7(Bt.br, 4, k), b5 = AB7(Bt.br, 5, k), b6 = AB7(Bt.br, 6, k);
            const double lx = AB1(Bt.ls_pos, k), lv = AB1(Bt.ls_vel, k), la = AB1(Bt.ls_acc, k);
            out[6 * j + c] = lx + (s[8] * b6 + s[7] * b5 + s[6] * b4 + s[5] * b3 + s[4] * b2 + s[3] * b1 + s[2] * b0 + s[1] * la + s[0] * lv);
            out[6 * j + 3 + c] = lv + sv[7] * b6 + sv[6] * b5 + sv[5] * b4 + sv[4] * b3 + sv[3] * b2 + sv[2] * b1 + sv[1] * b0 + sv[0] * la;
        }
    }
}

/* reb_check_exit: decides whether integrate() goes on; may shorten dt for the last step. */
__device__ int ab_check_exit(double t, double& dt, double dt_last, int& status, double tmax, int exact_finish_time, double& last_full_dt) {
    const double dtsign = copysign(1., dt);
    if (!(dt == dt) || !(t == t)) status = 1;      /* NaN step or time: stop instead of spinning forever */
    if (status >= 0) {
        /* exit now */
    } else if (exact_finish_time == 1) {
        if ((t + dt) * dtsign >= tmax * dtsign) {
            if (t == tmax) {
                status = 0;
            } else if (status == -2) {
                double tscale = 1e-12 * fabs(tmax);
                if (tscale < 1e-200) tscale = 1e-12;
                if (fabs(t - tmax) < tscale) status = 0;
                else dt = tmax - t;
            } else {
                status = -2;
                if (dt_last != 0.) last_full_dt = dt_last;
                dt = tmax - t;
            }
        } else {
            if (status == -2) status = -1;
        }
    } else {
        if (t * dtsign >= tmax * dtsign) status = 0;
    }
    return status;
}

}  // namespace AB_NS
#endif
