/*
 * pp_common.cuh -- per-system bookkeeping shared by the per-particle kernels (kernels.cu: pp_queue_kernel,
 * pp_integrate_kernel, pp_dense_kernel; coop_roles.cuh: pp_coop_kernel).
 */
#ifndef AB_PP_COMMON_CUH
#define AB_PP_COMMON_CUH

#include "device_types.h"
#include "ias15_device.cuh"

namespace AB_NS {

struct PPState {
    double t, dt, dt_last, last_full_dt;
    int status;
    unsigned long long steps, rejected, iters, evals;
};

__device__ __forceinline__ void pp_load(const AbBatch& Bt, long long i, PPState& P) {
    P.t = Bt.t[i]; P.dt = Bt.dt[i]; P.dt_last = Bt.dt_last[i]; P.last_full_dt = Bt.last_full_dt[i]; P.status = Bt.status[i];
    P.steps = Bt.steps[i]; P.rejected = Bt.rejected[i]; P.iters = Bt.iters[i]; P.evals = Bt.evals[i];
}
__device__ __forceinline__ void pp_store(const AbBatch& Bt, long long i, const PPState& P) {
    Bt.t[i] = P.t; Bt.dt[i] = P.dt; Bt.dt_last[i] = P.dt_last; Bt.last_full_dt[i] = P.last_full_dt;
    if (Bt.status[i] < 1000) Bt.status[i] = P.status;
    Bt.steps[i] = P.steps; Bt.rejected[i] = P.rejected; Bt.iters[i] = P.iters; Bt.evals[i] = P.evals;
}

/* Move one system between the population arrays (index `i`, stride src.n) and a working slot
 * (index `s`, stride dst.n): 54 * C doubles.  CG: the source is read past L1 (another SM may have
 * written it earlier in the same launch: time slices of one system run on whatever thread is free). */
template <bool CG>
__device__ __forceinline__ double pp_ld(const double* p) { return CG ? __ldcg(p) : *p; }

template <bool CG>
__device__ void pp_copy_system(const AbBatch& src, long long i, const AbBatch& dst, long long s, int nv) {
    const long long ns = src.n, nd = dst.n;
    const int C = src.C;
    const int Ca = 3 * (1 + nv);
    double* const s1[12] = {src.pos, src.vel, src.acc, src.x0, src.v0, src.a0, src.csx, src.csv, src.ls_pos, src.ls_vel, src.ls_acc, src.prm};
    double* const d1[12] = {dst.pos, dst.vel, dst.acc, dst.x0, dst.v0, dst.a0, dst.csx, dst.csv, dst.ls_pos, dst.ls_vel, dst.ls_acc, dst.prm};
    double* const s7[6] = {src.b, src.g, src.e, src.csb, src.br, src.er};
    double* const d7[6] = {dst.b, dst.g, dst.e, dst.csb, dst.br, dst.er};
    /* a body (3 components) at a time: 36 / 21 independent loads before the stores, so that the trips to L2 overlap */
    for (int k0 = 0; k0 < Ca; k0 += 3) {
        double tmp[12][3];
#pragma unroll
        for (int a = 0; a < 12; a++)
#pragma unroll
            for (int c = 0; c < 3; c++) tmp[a][c] = pp_ld<CG>(s1[a] + (long long)(k0 + c) * ns + i);
#pragma unroll
        for (int a = 0; a < 12; a++)
#pragma unroll
            for (int c = 0; c < 3; c++) d1[a][(long long)(k0 + c) * nd + s] = tmp[a][c];
    }
    for (int a = 0; a < 6; a++)
        for (int k0 = 0; k0 < Ca; k0 += 3) {
            double tmp[7][3];
#pragma unroll
            for (int j = 0; j < 7; j++)
#pragma unroll
                for (int c = 0; c < 3; c++) tmp[j][c] = pp_ld<CG>(s7[a] + ((long long)j * C + k0 + c) * ns + i);
#pragma unroll
            for (int j = 0; j < 7; j++)
#pragma unroll
                for (int c = 0; c < 3; c++) d7[a][((long long)j * C + k0 + c) * nd + s] = tmp[j][c];
        }
}

__device__ __forceinline__ void pp_load_cg(const AbBatch& Bt, long long i, PPState& P) {
    P.t = __ldcg(Bt.t + i); P.dt = __ldcg(Bt.dt + i); P.dt_last = __ldcg(Bt.dt_last + i); P.last_full_dt = __ldcg(Bt.last_full_dt + i);
    P.status = __ldcg(Bt.status + i);
    P.steps = __ldcg(Bt.steps + i); P.rejected = __ldcg(Bt.rejected + i); P.iters = __ldcg(Bt.iters + i); P.evals = __ldcg(Bt.evals + i);
}

/* State at one output epoch of assist_integrate_or_interpolate (reference src/assist.c:642-680, 556-597):
 * the system sits in slot `s` of W at time P.t, its last completed step was P.dt_last long. */
__device__ void pp_emit(const AbBatch& W, long long s, int nv, const PPState& P, double t, double* __restrict__ o) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double h = 1.0 - (P.t - t) / P.dt_last;
    if (P.status > 0) {
        for (int q = 0; q < 6 * (1 + nv); q++) o[q] = nan;
    } else if (P.t - t == 0.) {
        for (int j = 0; j <= nv; j++)
            for (int c = 0; c < 3; c++) {
                o[6 * j + c] = W.pos[(long long)(3 * j + c) * W.n + s];
                o[6 * j + 3 + c] = W.vel[(long long)(3 * j + c) * W.n + s];
            }
    } else if (h < 0.0 || h >= 1.0 || !ab_isnormal(h)) {
        for (int q = 0; q < 6 * (1 + nv); q++) o[q] = nan;
    } else {
        ab_interpolate(W, s, nv, P.dt_last, h, o);
    }
}

/* entry of reb_simulation_integrate(tmax) */
__device__ __forceinline__ void pp_integrate_entry(PPState& P, double tmax) {
    if (tmax != P.t) P.dt = copysign(P.dt, (tmax > P.t) ? 1.0 : -1.0);
    P.last_full_dt = P.dt;
    P.dt_last = 0.;
    P.status = -1;
}

}  // namespace AB_NS
#endif
