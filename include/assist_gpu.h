/*
 * assist_gpu.h -- thin C-ABI over the sm_100a kernels (plain pointers and sizes only).
 *
 * These are the entry points a binding (ctypes, cgo, JNI ...) uses to drive the GPU
 * path directly with whole populations, without going through per-particle
 * `struct reb_particle` marshalling.  The ASSIST host API in assist.h
 * (assist_attach / reb_simulation_integrate / assist_integrate_or_interpolate ...)
 * is implemented on top of exactly these calls.
 *
 * What each call replaces in the reference:
 *   assist_gpu_ephem_eval            assist_all_ephem            src/forces.c:175-263
 *                                    assist_spk_target_pos       src/spk.c:492-547
 *                                    assist_spk_calc             src/spk.c:405-481
 *                                    assist_ascii_calc           src/ascii_ephem.c:275-384
 *   assist_gpu_eval_forces           assist_additional_forces    src/forces.c:49-173 (+ :266-1983)
 *   assist_gpu_batch_integrate       reb_simulation_integrate -> reb_integrator_ias15_step
 *                                    (REBOUND; call site src/assist.c:663)
 *   assist_gpu_batch_integrate_or_interpolate
 *                                    assist_integrate_or_interpolate + assist_interpolate
 *                                    src/assist.c:642-680, 556-597
 *
 * Layouts.  A "system" is one real particle followed by its n_var first-order
 * variational particles (K = 1 + n_var bodies).  Host buffers are plain row-major
 * doubles:
 *   state  [n_sys][K][6]   x y z vx vy vz          (AU, AU/day; barycentric)
 *   params [n_sys][K][3]   A1 A2 A3 for the real particle; dA1 dA2 dA3 for each
 *                          variational particle (reference src/forces.c:839-841, 1030-1032)
 *   acc    [n_sys][K][3]   ax ay az
 * All functions return 0 on success, an ASSIST_STATUS (>0) for ephemeris errors, or
 * a negative ASSIST_GPU_ERR_* code; assist_gpu_last_error() gives the message.
 */
#ifndef _ASSIST_B200_GPU_H
#define _ASSIST_B200_GPU_H

#include "assist.h"

#ifdef __cplusplus
extern "C" {
#endif

enum ASSIST_GPU_ERR {
    ASSIST_GPU_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: there is NO CPU fallback */
    ASSIST_GPU_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    ASSIST_GPU_ERR_ARG = -3,         /* bad argument */
    ASSIST_GPU_ERR_UNSUPPORTED = -4, /* configuration outside what the kernels implement */
};

enum ASSIST_GPU_MODE {
    ASSIST_GPU_SHARED_STEP = 0,      /* one global dt, global convergence: REBOUND semantics for an N-particle sim */
    ASSIST_GPU_PER_PARTICLE = 1,     /* every system steps on its own dt: one-simulation-per-particle semantics */
};

enum ASSIST_GPU_MATH {
    ASSIST_GPU_MATH_STRICT = 0,      /* no FMA contraction, reference operation order: bit-faithful to the C build */
    ASSIST_GPU_MATH_FAST = 1,        /* FMA contraction allowed (results agree to rounding, not bitwise) */
};

#define ASSIST_GPU_MAX_NVAR 6        /* variational particles per real particle the kernels are built for */

struct assist_gpu_options {
    int forces;                      /* ASSIST_FORCES bitmask, reference src/assist.h:76-87 */
    int gr_eih_sources;              /* 1..11, reference src/assist.c:415 */
    int geocentric;                  /* reference src/forces.c:57, 72-87 */
    int math;                        /* ASSIST_GPU_MATH */
    double alpha, nk, nm, nn, r0;    /* Marsden g(r), reference src/assist.c:434-438 */
    double epsilon;                  /* IAS15 accuracy parameter (REBOUND default 1e-9) */
    double min_dt;                   /* IAS15 minimum |dt| (REBOUND default 0) */
};

struct assist_gpu_stats {
    unsigned long long steps;        /* accepted IAS15 steps, summed over systems (per-particle) or global steps (shared) */
    unsigned long long steps_rejected;
    unsigned long long pc_iterations;/* predictor-corrector sweeps */
    unsigned long long force_evals;  /* force evaluations x systems */
    unsigned long long kernel_launches;
    double last_kernel_ms;           /* CUDA-event time of the last integrate call, on the launch stream */
};

typedef struct assist_gpu_batch assist_gpu_batch;

/* ---- device management --------------------------------------------------- */
int assist_gpu_device_count(void);
int assist_gpu_set_device(int device);
const char* assist_gpu_last_error(void);
void assist_gpu_default_options(struct assist_gpu_options* opt);
/* Self-test of the kernels' branch-free IEEE division / square root against the built-in operators on n_pairs
 * pseudo-random operand pairs (seeded); mismatches[0] = quotients, mismatches[1] = roots that differ in any bit.
 * New with the GPU build: the reference divides on the host (src/forces.c:331-333 and every other quotient). */
int assist_gpu_selftest_fp(unsigned long long seed, long long n_pairs, unsigned long long mismatches[2]);

/* ---- ephemeris ------------------------------------------------------------ */
/* Upload (once per device) the coefficient tables of an initialised ephemeris. */
int assist_gpu_ephem_upload(const struct assist_ephem* ephem);
/* Number of bodies: 11 + asteroids in the small-body file. */
int assist_gpu_ephem_nbodies(const struct assist_ephem* ephem);
/* out[n_t][nbodies][10] = GM x y z vx vy vz ax ay az; status[n_t][nbodies] = ASSIST_STATUS. */
int assist_gpu_ephem_eval(const struct assist_ephem* ephem, int math, const double* t, int n_t,
                          double* out, int* status);

/* ---- one force evaluation (term-by-term parity hook) ---------------------- */
/* t has n_sys entries when t_per_system != 0, else one entry shared by all systems. */
int assist_gpu_eval_forces(const struct assist_ephem* ephem, const struct assist_gpu_options* opt,
                           int n_sys, int n_var, const double* t, int t_per_system,
                           const double* state, const double* params, double* acc, int* status);

/* ---- batched IAS15 -------------------------------------------------------- */
assist_gpu_batch* assist_gpu_batch_create(const struct assist_ephem* ephem, int n_sys, int n_var, int mode);
void assist_gpu_batch_free(assist_gpu_batch* b);
int assist_gpu_batch_set_options(assist_gpu_batch* b, const struct assist_gpu_options* opt);
/* (Re)start: copies state/params from HOST memory, sets t, dt for every system, clears IAS15 history.
 * nvar_per_system may be NULL (every system uses n_var). params may be NULL (no non-gravitational forces). */
int assist_gpu_batch_set_state(assist_gpu_batch* b, double t0, double dt0, const double* state,
                               const double* params, const int* nvar_per_system);
/* Replace positions/velocities only, keeping the IAS15 history (what REBOUND does when a user
 * edits sim->particles between two integrate calls). */
int assist_gpu_batch_update_particles(assist_gpu_batch* b, const double* state);
/* Keep a device-resident copy of the current state as "initial conditions"; restart from it
 * without any host traffic (used to time the device path with inputs already in HBM). */
int assist_gpu_batch_snapshot(assist_gpu_batch* b);
int assist_gpu_batch_restore(assist_gpu_batch* b);
/* reb_simulation_integrate(tmax) for every system. exact_finish_time as in REBOUND. max_steps<=0: no limit
 * (shared-step mode only: stop after that many accepted steps). */
int assist_gpu_batch_integrate(assist_gpu_batch* b, double t_end, int exact_finish_time, long max_steps);
/* assist_integrate_or_interpolate(times[e]) for e = 0..n_times-1, per system (reference src/assist.c:642-680):
 * out[n_times][n_sys][K][6], per-particle batches only.  The epochs are visited in the given order, whatever it is
 * (an epoch inside the last completed step is interpolated, any other one integrates towards it, as the reference
 * does call by call); lists that run in one direction are additionally cut into time slices for load balance.
 * Rows that cannot be produced (a system that left the ephemeris coverage, variational slots a system does not use)
 * hold NaN. */
int assist_gpu_batch_integrate_or_interpolate(assist_gpu_batch* b, const double* times, int n_times, double* out);
/* state/acc[n_sys][K][6|3]; t, dt, dt_last_done: n_sys entries in per-particle mode, 1 in shared-step mode.
 * Any pointer may be NULL. */
int assist_gpu_batch_get_state(assist_gpu_batch* b, double* state, double* acc, double* t, double* dt,
                               double* dt_last_done, int* status);
int assist_gpu_batch_set_time(assist_gpu_batch* b, double t, double dt);
/* Dense output inside the last completed step (shared-step mode): out[n_sys][K][6]. */
int assist_gpu_batch_interpolate(assist_gpu_batch* b, double h, double* out);
int assist_gpu_batch_get_stats(assist_gpu_batch* b, struct assist_gpu_stats* stats);
/* Per-system counters of a per-particle batch (each array n_sys long; any may be NULL). */
int assist_gpu_batch_get_counters(assist_gpu_batch* b, unsigned long long* steps, unsigned long long* rejected,
                                  unsigned long long* iters, unsigned long long* evals);
/* ---- one population over several GPUs of one box (north star (4)) -------------------------------------------
 * New with the GPU build (the reference runs one simulation per process).  Per-particle semantics
 * only.  One sub-batch per device, one host thread per device for every call that launches work; state, outputs and
 * counters are scattered / gathered in the order of the caller's population.  devices == NULL: every visible device.
 * The systems are dealt out in the order of their expected step counts, so every device gets the same mix. */
typedef struct assist_gpu_multi assist_gpu_multi;
assist_gpu_multi* assist_gpu_multi_create(const struct assist_ephem* ephem, int n_sys, int n_var, const int* devices, int n_devices);
void assist_gpu_multi_free(assist_gpu_multi* m);
int assist_gpu_multi_device_count(const assist_gpu_multi* m);
const char* assist_gpu_multi_last_error(void);
int assist_gpu_multi_set_options(assist_gpu_multi* m, const struct assist_gpu_options* opt);
int assist_gpu_multi_set_state(assist_gpu_multi* m, double t0, double dt0, const double* state, const double* params);
int assist_gpu_multi_integrate(assist_gpu_multi* m, double t_end, int exact_finish_time);
int assist_gpu_multi_integrate_or_interpolate(assist_gpu_multi* m, const double* times, int n_times, double* out);
int assist_gpu_multi_get_state(assist_gpu_multi* m, double* state, double* t, double* dt, double* dt_last_done, int* status);
int assist_gpu_multi_get_counters(assist_gpu_multi* m, unsigned long long* steps, unsigned long long* rejected,
                                  unsigned long long* iters, unsigned long long* evals);
/* sums over the devices; last_kernel_ms = the slowest device; kernel_ms_per_device (may be NULL): n_devices entries */
int assist_gpu_multi_get_stats(assist_gpu_multi* m, struct assist_gpu_stats* stats, double* kernel_ms_per_device);

/* assist_interpolate_simulation (reference src/assist.c:682-752) for m = 3 N components: the state at fraction h of
 * the step that starts at (x0, v0, a0) and has the b coefficients br[7][m] and length dt_last_done.  Host arrays. */
int assist_gpu_interpolate_simulation(int m, const double* x0, const double* v0, const double* a0, const double* br,
                                      double dt_last_done, double h, double* pos, double* vel);

/* Page-locked host memory for state / output buffers (cudaHostAlloc). */
void* assist_gpu_host_alloc(size_t bytes);
void assist_gpu_host_free(void* p);
/* Peak FP64 FMA rate of the current device measured with a register-resident DFMA loop (TFLOP/s). */
double assist_gpu_measure_fp64_peak(int iters);
/* Number of CUDA kernels this library has launched in this process so far (all devices, all batches). */
unsigned long long assist_gpu_kernel_launches(void);
/* Host-only test hook: the re-packed copy of an SPK kernel that is uploaded to the device (type-2 records only,
 * each [_jul(MID), RADIUS, (x y z) of term 0, (x y z) of term 1, ...], zero terms up to a multiple of four).
 * *out is malloc'ed; seg_off[target * 4 + segment] = first word of that segment in the copy. */
struct spk_s;
int assist_gpu_spk_pack_host(const struct spk_s* file, double** out, size_t* words, long long* seg_off, int seg_off_len);

#ifdef __cplusplus
}
#endif
#endif
