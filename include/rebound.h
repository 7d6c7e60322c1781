/*
 * rebound.h -- minimal REBOUND-compatible surface used by assist-b200.
 *
 * REBOUND (PyPI `rebound>=4.4.11,<5`, CI tag 4.6.0; reference setup.py:177-179,
 * .github/workflows/c.yml:19-21) is a third-party dependency of ASSIST that is NOT
 * vendored in the reference tree.  ASSIST is a force plug-in for REBOUND's IAS15
 * integrator (reference src/assist.c:440-446), so a drop-in for ASSIST's hot path
 * needs the part of REBOUND's public surface that ASSIST, its unit tests and its
 * examples touch.  This header declares exactly that part, with REBOUND's names:
 *
 *   types      reb_particle, reb_dp7, reb_variational_configuration,
 *              reb_integrator_ias15, reb_simulation, reb_simulationarchive
 *   functions  reb_simulation_{create,free,add,add_fmt,copy,integrate,step,
 *              update_acceleration,error,add_variation_1st_order},
 *              reb_particle_{iadd,distance,com_of_pair}
 *
 * Field names and meanings follow REBOUND 4.x; the struct LAYOUT is our own (the
 * real header is not available offline).  `struct reb_particle` is laid out as
 * REBOUND's 128-byte record (9 state doubles, m, r, last_collision, c, hash, ap,
 * sim) because ASSIST memcpy's arrays of it (reference src/assist.c:649-653).
 *
 * Two implementations sit behind this header:
 *   - oracle/reb_shim.c       CPU restatement of REBOUND's IAS15 (test oracle only)
 *   - assist_b200/csrc/ (C++/CUDA)  product: the same calls drive the CUDA batch stepper
 */
#ifndef _REBOUND_B200_SURFACE_H
#define _REBOUND_B200_SURFACE_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <math.h>
#include <sys/time.h>      /* REBOUND's header brings struct timeval along (reference unit_tests/benchmark/problem.c relies on it) */

#ifdef __cplusplus
extern "C" {
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct reb_simulation;
struct reb_treecell;

struct reb_particle {
    double x, y, z;
    double vx, vy, vz;
    double ax, ay, az;
    double m;
    double r;
    double last_collision;
    struct reb_treecell* c;
    uint32_t hash;
    void* ap;
    struct reb_simulation* sim;
};

struct reb_dp7 {
    double* p0;
    double* p1;
    double* p2;
    double* p3;
    double* p4;
    double* p5;
    double* p6;
};

struct reb_variational_configuration {
    struct reb_simulation* sim;
    int order;              /* 1 = first order */
    int index;              /* index of the first variational particle */
    int testparticle;       /* index of the real particle being varied */
    int index_1st_order_a;
    int index_1st_order_b;
    double lrescale;
};

enum REB_STATUS {
    REB_STATUS_SINGLE_STEP = -10,
    REB_STATUS_SCREENSHOT_READY = -5,
    REB_STATUS_SCREENSHOT = -4,
    REB_STATUS_PAUSED = -3,
    REB_STATUS_LAST_STEP = -2,
    REB_STATUS_RUNNING = -1,
    REB_STATUS_SUCCESS = 0,
    REB_STATUS_GENERIC_ERROR = 1,
    REB_STATUS_NO_PARTICLES = 2,
    REB_STATUS_ENCOUNTER = 3,
    REB_STATUS_ESCAPE = 4,
    REB_STATUS_USER = 5,
    REB_STATUS_SIGINT = 6,
    REB_STATUS_COLLISION = 7,
};

enum REB_INTEGRATOR {
    REB_INTEGRATOR_IAS15 = 0,
    REB_INTEGRATOR_WHFAST = 1,
    REB_INTEGRATOR_NONE = 7,
};

enum REB_GRAVITY {
    REB_GRAVITY_NONE = 0,
    REB_GRAVITY_BASIC = 1,
    REB_GRAVITY_COMPENSATED = 2,
};

enum reb_simulation_binary_error_codes {
    REB_SIMULATION_BINARY_WARNING_NONE = 0,
};

struct reb_integrator_ias15 {
    double epsilon;                 /* accuracy control parameter, default 1e-9 */
    double min_dt;                  /* minimum |dt|, default 0 */
    unsigned int adaptive_mode;     /* 0 individual, 1 global (ASSIST), 2 PRS23 */
    uint64_t iterations_max_exceeded;
    unsigned int N_allocated;       /* particles the arrays below are sized for */
    double* at;
    double* x0;
    double* v0;
    double* a0;
    double* csx;
    double* csv;
    double* csa0;
    struct reb_dp7 g;
    struct reb_dp7 b;
    struct reb_dp7 csb;
    struct reb_dp7 e;
    struct reb_dp7 br;              /* b of the last completed step (dense output) */
    struct reb_dp7 er;
    int* map;
    unsigned int N_allocated_map;
    /* assist-b200 extension: number of predictor-corrector sweeps and force
     * evaluations done so far (REBOUND itself does not count these). */
    uint64_t b200_pc_iterations;
    uint64_t b200_force_evals;
    uint64_t b200_steps_rejected;
};

struct reb_simulation {
    double t;
    double G;
    double softening;
    double dt;
    double dt_last_done;
    uint64_t steps_done;
    unsigned int N;                 /* real + variational particles */
    int N_var;                      /* number of variational particles */
    unsigned int N_var_config;
    struct reb_variational_configuration* var_config;
    int N_active;
    unsigned int N_allocated;
    struct reb_particle* particles;
    enum REB_STATUS status;
    int exact_finish_time;
    unsigned int force_is_velocity_dependent;
    enum REB_INTEGRATOR integrator;
    enum REB_GRAVITY gravity;
    struct reb_integrator_ias15 ri_ias15;
    void (*additional_forces)(struct reb_simulation* const r);
    void (*pre_timestep_modifications)(struct reb_simulation* const r);
    void (*post_timestep_modifications)(struct reb_simulation* const r);
    void (*heartbeat)(struct reb_simulation* r);
    void* extras;
    void (*extras_cleanup)(struct reb_simulation* r);
    /* error/warning messages raised through reb_simulation_error() */
    char* messages;
    int messages_waiting;
    /* assist-b200 extension: opaque handle of the device batch mirroring this sim */
    void* b200_batch;
    /* snapshots during reb_simulation_integrate (reb_simulation_save_to_file_step) */
    char* simulationarchive_filename;
    uint64_t simulationarchive_auto_step;   /* a snapshot every this many completed steps; 0: none */
    uint64_t simulationarchive_next_step;
};

/* A file of snapshots of an attached simulation (what assist_create_interpolated_simulation reads back): t[] and
 * nblobs as in REBOUND; the layout of the file is assist-b200's own (reb_surface.cpp), not REBOUND's binary format. */
struct reb_simulationarchive {
    struct reb_simulation* r;
    char* filename;
    long nblobs;
    double* t;
    long* b200_offset;      /* file offset of every snapshot */
};

/* ---- simulation life cycle --------------------------------------------- */
struct reb_simulation* reb_simulation_create(void);
void reb_simulation_free(struct reb_simulation* const r);
struct reb_simulation* reb_simulation_copy(struct reb_simulation* r);
void reb_simulation_add(struct reb_simulation* const r, struct reb_particle pt);
/* Only the "x y z vx vy vz m r" keys (any subset, any order) are understood. */
void reb_simulation_add_fmt(struct reb_simulation* r, const char* fmt, ...);
/* Appends one zeroed first-order variational particle for real particle
 * `testparticle`; returns its index.  Real particles must be added first. */
int reb_simulation_add_variation_1st_order(struct reb_simulation* const r, int testparticle);

/* ---- time stepping ----------------------------------------------------- */
enum REB_STATUS reb_simulation_integrate(struct reb_simulation* const r, double tmax);
void reb_simulation_step(struct reb_simulation* const r);
void reb_simulation_steps(struct reb_simulation* const r, unsigned int N_steps);
void reb_simulation_update_acceleration(struct reb_simulation* r);
void reb_simulation_error(struct reb_simulation* const r, const char* const msg);
void reb_simulation_warning(struct reb_simulation* const r, const char* const msg);

/* ---- particle helpers -------------------------------------------------- */
void reb_particle_iadd(struct reb_particle* p1, struct reb_particle* p2);
void reb_particle_isub(struct reb_particle* p1, struct reb_particle* p2);
double reb_particle_distance(struct reb_particle* p1, struct reb_particle* p2);
struct reb_particle reb_particle_com_of_pair(struct reb_particle p1, struct reb_particle p2);

/* ---- SimulationArchive ------------------------------------------------- */
/* Snapshots of an ASSIST-attached simulation: time, step sizes, particles and what dense output needs of the last
 * completed step (its start acceleration and b coefficients).  One is appended before the first step of
 * reb_simulation_integrate and after every `step` completed steps. */
void reb_simulation_save_to_file_step(struct reb_simulation* const r, const char* filename, unsigned long long step);
void reb_simulation_save_to_file(struct reb_simulation* const r, const char* filename);
struct reb_simulationarchive* reb_simulationarchive_create_from_file(const char* filename);
void reb_simulationarchive_free(struct reb_simulationarchive* sa);
/* Fills `r` (a fresh reb_simulation_create()) from snapshot `snapshot`: t, dt, dt_last_done, particles, and
 * ri_ias15.x0 / v0 / a0 / br as REBOUND leaves them after a completed step.  ASSIST is not attached to it. */
void reb_simulation_create_from_simulationarchive_with_messages(
        struct reb_simulation* r, struct reb_simulationarchive* sa, int64_t snapshot,
        enum reb_simulation_binary_error_codes* warnings);

#ifdef __cplusplus
}
#endif
#endif
