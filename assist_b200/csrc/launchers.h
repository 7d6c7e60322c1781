/* launchers.h -- host-callable kernel launchers, one set per math variant. */
#ifndef AB_LAUNCHERS_H
#define AB_LAUNCHERS_H

#include <cuda_runtime.h>
#include "device_types.h"

#define AB_DECLARE_LAUNCHERS(sfx)                                                                                   \
    cudaError_t ab_upload_constants_##sfx##_tu0();                                                                  \
    cudaError_t ab_upload_constants_##sfx##_tu1();                                                                  \
    cudaError_t ab_upload_constants_##sfx##_tu2();                                                                  \
    cudaError_t ab_upload_constants_##sfx##_tu3();                                                                  \
    cudaError_t ab_upload_constants_##sfx##_tu4();                                                                  \
    cudaError_t ab_pp_coop_max_grid_##sfx(int* max_grid);                                                           \
    cudaError_t ab_launch_fp_selftest_##sfx(unsigned long long seed, int blocks, int iters, unsigned long long* d_bad, cudaStream_t st); \
    cudaError_t ab_launch_pp_coop_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, const AbBatch& W, \
                                        double tmax, int exact, unsigned long long* queue_head, const AbSlices& SL,  \
                                        const double* times, int n_times, double* out, const void* plan,              \
                                        const AbSpkTarget* host_ast_tg, double* gtab, unsigned long long* timing, int grid, \
                                        cudaStream_t st);                                                             \
    cudaError_t ab_launch_ephem_eval_##sfx(const AbEphem& E, const double* t, int n_t, double* out, int* status,   \
                                           cudaStream_t st);                                                        \
    cudaError_t ab_launch_force_eval_##sfx(const AbEphem& E, const AbForceOpts& F, int n, int K, const double* t,  \
                                           int t_per_system, const double* state, const double* params,            \
                                           double* acc, int* status, cudaStream_t st);                             \
    cudaError_t ab_launch_spk_target_##sfx(const double* img, const AbSpkTarget& tg, int has_emb, const AbSpkTarget& emb, \
                                           double jd_ref, double t, int mode, const double* ud, double* out,         \
                                           cudaStream_t st);                                                        \
    cudaError_t ab_launch_ascii_work_##sfx(const double* P, int ncm, int ncf, int niv, double t0, double t1,          \
                                           double* out, cudaStream_t st);                                            \
    cudaError_t ab_launch_pp_integrate_k1_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt,         \
                                                double tmax, int exact, int resume, long long step_cap,            \
                                                const int* active, int n_active, cudaStream_t st);                 \
    cudaError_t ab_launch_pp_integrate_kv_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt,         \
                                                double tmax, int exact, int resume, long long step_cap,            \
                                                const int* active, int n_active, cudaStream_t st);                 \
    cudaError_t ab_launch_pp_queue_k1_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, const AbBatch& W,  \
                                            double tmax, int exact, unsigned long long* queue_head, const AbSlices& SL, \
                                            const double* times, int n_times, double* out, cudaStream_t st);        \
    cudaError_t ab_launch_pp_queue_kv_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt, const AbBatch& W,  \
                                            double tmax, int exact, unsigned long long* queue_head, const AbSlices& SL, \
                                            const double* times, int n_times, double* out, cudaStream_t st);        \
    cudaError_t ab_pp_resident_threads_k1_##sfx(int* threads);                                                      \
    cudaError_t ab_pp_resident_threads_kv_##sfx(int* threads);                                                      \
    cudaError_t ab_launch_pp_dense_k1_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt,             \
                                            const double* times, int n_times, double* out, cudaStream_t st);       \
    cudaError_t ab_launch_pp_dense_kv_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt,             \
                                            const double* times, int n_times, double* out, cudaStream_t st);       \
    cudaError_t ab_launch_sh_integrate_##sfx(const AbEphem& E, const AbForceOpts& F, const AbBatch& Bt,            \
                                             double tmax, int exact, long long max_steps, int flags,               \
                                             cudaStream_t st);                                                     \
    cudaError_t ab_launch_sh_interpolate_##sfx(const AbBatch& Bt, double dt_last_done, double h, double* out,      \
                                               cudaStream_t st);

AB_DECLARE_LAUNCHERS(strict)
AB_DECLARE_LAUNCHERS(fast)

#endif
