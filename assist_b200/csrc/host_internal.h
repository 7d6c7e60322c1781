/* host_internal.h -- helpers shared by the host-side translation units (not part of the public ABI). */
#ifndef AB_HOST_INTERNAL_H
#define AB_HOST_INTERNAL_H

#include "assist.h"
#include "assist_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* extra status code: no CUDA device (message index 6 in assist_error_messages) */
#define ASSIST_ERROR_GPU 6

void ab_host_drop_batch(struct reb_simulation* r);
void ab_host_fill_options(const struct reb_simulation* r, const struct assist_extras* ax, struct assist_gpu_options* opt);
int ab_host_interpolate(struct reb_simulation* r, double h, struct reb_particle* dest);
/* The hook assist_integrate_or_interpolate installs (reference src/assist.c:645, 754-758).  The GPU
 * stepper records the state at the start of every step itself, so this is only a marker. */
void ab_host_pre_timestep_marker(struct reb_simulation* r);

/* flags: 1 = resume an interrupted integrate (keep status / last_full_dt / dt_last_done),
 *        2 = exactly one reb_simulation_step, no exit logic (shared-step batches) */
int ab_gpu_batch_integrate_ex(assist_gpu_batch* b, double t_end, int exact_finish_time, long max_steps, int flags);
int ab_gpu_batch_update_params(assist_gpu_batch* b, const double* params);
int ab_gpu_batch_get_last_state(assist_gpu_batch* b, double* state, double* acc);
int ab_gpu_batch_get_br(assist_gpu_batch* b, double* br);      /* br[7][n][K][3] */
struct spk_s;
int ab_gpu_spk_target_eval(struct spk_s* file, int target_index, int emb_index, double jd_ref, double jd_rel,
                           int mode, const double* ud, double* out);
int ab_gpu_ascii_work(const double* P, int ncm, int ncf, int niv, double t0, double t1, double* out);
/* order[k] = the system with the k-th largest expected step count (a^-3/2 (1 - e)^-1 of the osculating orbit);
 * state[n][K][6], the real particle first.  Host only. */
void ab_gpu_cost_order_host(const double* state, int n, int K, int* order);
/* frees the cached descriptor block of an SPK file (struct spk_s::b200_host_desc) */
void ab_spk_desc_free(void* desc);

#ifdef __cplusplus
}
#endif
#endif
