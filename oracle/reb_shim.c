/*
 * reb_shim.c -- CPU restatement of the part of REBOUND that ASSIST drives.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (assist_b200/, the
 * C-ABI library) may include, link or call this file; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it.
 *
 * PARITY UNPINNED: REBOUND (rebound>=4.4.11,<5; CI tag 4.6.0 -- reference
 * setup.py:177-179, .github/workflows/c.yml:19-21) is a third-party dependency
 * that is not vendored in /root/reference and is not installed in this image.
 * This file restates the published IAS15 algorithm (Rein & Spiegel 2015, Everhart
 * 1985) as REBOUND 4.x implements it, anchored on the reference's own call sites:
 *   - integrator / gravity / adaptive_mode=1 / velocity-dependent forces:
 *       reference src/assist.c:440-446
 *   - b-series position and velocity polynomials (the /2,/6,/12,/20,... and
 *       /2../8 divisors): reference src/assist.c:562-580 (assist_interpolate)
 *   - br, x0, v0, a0 as the dense-output state: src/assist.c:674-677, 687-691
 *   - pre_timestep_modifications hook + update_acceleration: src/assist.c:645, 754-758
 *   - N_var / var_config bookkeeping: src/forces.c:59-60, 406-409
 *   - exact_finish_time=0 overshoot semantics: src/assist.c:646, 658-667
 *   - direction flip toward tmax: unit_tests/holman_reverse_spk/problem.c:23-30
 * The constants h/rr/c/d are derived, not remembered (gen_ias15_constants.py).
 * What IS checked offline: the reference's data-independent invariants
 *  run on this shim together with the reference's own,
 * verbatim-compiled force and ephemeris code (oracle/_ref).
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <stdio.h>
#include <math.h>
#include <float.h>
#include "rebound.h"
#include "ias15_constants.h"

#define hh ORC_H
#define rr ORC_RR
#define cc ORC_C
#define dd ORC_D

static const double safety_factor = 0.25;

/* ---- variants of the two places where this restatement could differ from a REBOUND 4.x binary at round-off or
 * worse ("What's weak" 1).  Defaults = what every golden vector was generated with.  The spread between
 * the variants bounds what the real REBOUND could change (tests/test_cpu_oracle.py::test_ias15_variant_spread,
 * DESIGN.md section 2).
 *   predictor_form 0: nested Horner form (the form of REBOUND's integrator_ias15.c as remembered)
 *                  1: precomputed s[0..8] / sv[0..7] coefficients, the form the reference's own copy of the same
 *                     polynomial uses (reference src/assist.c:562-596) -- same value, different rounding
 *   reject_restores_acc 1: a rejected step restores accelerations together with positions and velocities
 *                       0: positions and velocities only */
static int shim_predictor_form = 0;
static int shim_reject_restores_acc = 1;
void reb_shim_set_variant(int predictor_form, int reject_restores_acc){
    shim_predictor_form = predictor_form;
    shim_reject_restores_acc = reject_restores_acc;
}

/* ------------------------------------------------------------------------ */
/* simulation life cycle                                                    */
/* ------------------------------------------------------------------------ */

struct reb_simulation* reb_simulation_create(void){
    struct reb_simulation* r = calloc(1, sizeof(struct reb_simulation));
    r->t = 0.0;
    r->G = 1.0;
    r->dt = 0.001;
    r->dt_last_done = 0.0;
    r->N_active = -1;
    r->status = REB_STATUS_PAUSED;
    r->exact_finish_time = 1;
    r->integrator = REB_INTEGRATOR_IAS15;
    r->gravity = REB_GRAVITY_BASIC;
    r->ri_ias15.epsilon = 1e-9;
    r->ri_ias15.min_dt = 0.0;
    r->ri_ias15.adaptive_mode = 2;   /* REBOUND 4.x default (PRS23); ASSIST sets 1 */
    return r;
}

static void free_dp7(struct reb_dp7* d){
    free(d->p0); free(d->p1); free(d->p2); free(d->p3); free(d->p4); free(d->p5); free(d->p6);
    memset(d, 0, sizeof(*d));
}

static void ias15_free_arrays(struct reb_simulation* r){
    struct reb_integrator_ias15* ri = &r->ri_ias15;
    free(ri->at); free(ri->x0); free(ri->v0); free(ri->a0);
    free(ri->csx); free(ri->csv); free(ri->csa0); free(ri->map);
    ri->at = ri->x0 = ri->v0 = ri->a0 = ri->csx = ri->csv = ri->csa0 = NULL;
    ri->map = NULL;
    free_dp7(&ri->g); free_dp7(&ri->b); free_dp7(&ri->csb);
    free_dp7(&ri->e); free_dp7(&ri->br); free_dp7(&ri->er);
    ri->N_allocated = 0;
    ri->N_allocated_map = 0;
}

void reb_simulation_free(struct reb_simulation* const r){
    if (r == NULL) return;
    if (r->extras_cleanup) r->extras_cleanup(r);
    ias15_free_arrays(r);
    free(r->particles);
    free(r->var_config);
    free(r->messages);
    free(r);
}

struct reb_simulation* reb_simulation_copy(struct reb_simulation* r){
    struct reb_simulation* c = reb_simulation_create();
    c->t = r->t; c->G = r->G; c->dt = r->dt; c->dt_last_done = r->dt_last_done;
    c->exact_finish_time = r->exact_finish_time;
    c->force_is_velocity_dependent = r->force_is_velocity_dependent;
    c->integrator = r->integrator; c->gravity = r->gravity;
    c->ri_ias15.epsilon = r->ri_ias15.epsilon;
    c->ri_ias15.min_dt = r->ri_ias15.min_dt;
    c->ri_ias15.adaptive_mode = r->ri_ias15.adaptive_mode;
    c->N_active = r->N_active;
    for (unsigned int i = 0; i < r->N; i++){
        reb_simulation_add(c, r->particles[i]);
    }
    c->N_var = r->N_var;
    c->N_var_config = r->N_var_config;
    if (r->N_var_config){
        c->var_config = malloc(sizeof(struct reb_variational_configuration) * r->N_var_config);
        memcpy(c->var_config, r->var_config, sizeof(struct reb_variational_configuration) * r->N_var_config);
        for (unsigned int v = 0; v < c->N_var_config; v++) c->var_config[v].sim = c;
    }
    return c;
}

void reb_simulation_add(struct reb_simulation* const r, struct reb_particle pt){
    if (r->N >= r->N_allocated){
        r->N_allocated = r->N_allocated ? 2 * r->N_allocated : 128;
        r->particles = realloc(r->particles, sizeof(struct reb_particle) * r->N_allocated);
    }
    pt.sim = r;
    r->particles[r->N] = pt;
    r->N++;
}

void reb_simulation_add_fmt(struct reb_simulation* r, const char* fmt, ...){
    struct reb_particle p = {0};
    va_list args;
    va_start(args, fmt);
    char* copy = strdup(fmt);
    char* save = NULL;
    for (char* tok = strtok_r(copy, " ", &save); tok; tok = strtok_r(NULL, " ", &save)){
        double v = va_arg(args, double);
        if      (!strcmp(tok, "x"))  p.x = v;
        else if (!strcmp(tok, "y"))  p.y = v;
        else if (!strcmp(tok, "z"))  p.z = v;
        else if (!strcmp(tok, "vx")) p.vx = v;
        else if (!strcmp(tok, "vy")) p.vy = v;
        else if (!strcmp(tok, "vz")) p.vz = v;
        else if (!strcmp(tok, "m"))  p.m = v;
        else if (!strcmp(tok, "r"))  p.r = v;
        else reb_simulation_error(r, "reb_simulation_add_fmt: unsupported key.");
    }
    free(copy);
    va_end(args);
    reb_simulation_add(r, p);
}

int reb_simulation_add_variation_1st_order(struct reb_simulation* const r, int testparticle){
    r->N_var_config++;
    r->var_config = realloc(r->var_config, sizeof(struct reb_variational_configuration) * r->N_var_config);
    struct reb_variational_configuration* vc = &r->var_config[r->N_var_config - 1];
    memset(vc, 0, sizeof(*vc));
    vc->sim = r;
    vc->order = 1;
    vc->index = (int)r->N;
    vc->testparticle = testparticle;
    vc->lrescale = 0;
    struct reb_particle p0 = {0};
    if (testparticle >= 0){
        reb_simulation_add(r, p0);
        r->N_var++;
    }else{
        /* REBOUND: testparticle<0 varies all real particles.  Not used by ASSIST. */
        int N_real = (int)r->N - r->N_var;
        for (int i = 0; i < N_real; i++) reb_simulation_add(r, p0);
        r->N_var += N_real;
    }
    return vc->index;
}

void reb_simulation_error(struct reb_simulation* const r, const char* const msg){
    fprintf(stderr, "\n(REBOUND shim) Error: %s\n", msg);
    if (r){
        free(r->messages);
        r->messages = strdup(msg);
        r->messages_waiting = 1;
    }
}

void reb_simulation_warning(struct reb_simulation* const r, const char* const msg){
    (void)r;
    fprintf(stderr, "\n(REBOUND shim) Warning: %s\n", msg);
}

void reb_particle_iadd(struct reb_particle* p1, struct reb_particle* p2){
    p1->x += p2->x; p1->y += p2->y; p1->z += p2->z;
    p1->vx += p2->vx; p1->vy += p2->vy; p1->vz += p2->vz;
    p1->m += p2->m;
}

void reb_particle_isub(struct reb_particle* p1, struct reb_particle* p2){
    p1->x -= p2->x; p1->y -= p2->y; p1->z -= p2->z;
    p1->vx -= p2->vx; p1->vy -= p2->vy; p1->vz -= p2->vz;
    p1->m -= p2->m;
}

double reb_particle_distance(struct reb_particle* p1, struct reb_particle* p2){
    double dx = p1->x - p2->x, dy = p1->y - p2->y, dz = p1->z - p2->z;
    return sqrt(dx*dx + dy*dy + dz*dz);
}

struct reb_particle reb_particle_com_of_pair(struct reb_particle p1, struct reb_particle p2){
    p1.x = p1.x*p1.m + p2.x*p2.m; p1.y = p1.y*p1.m + p2.y*p2.m; p1.z = p1.z*p1.m + p2.z*p2.m;
    p1.vx = p1.vx*p1.m + p2.vx*p2.m; p1.vy = p1.vy*p1.m + p2.vy*p2.m; p1.vz = p1.vz*p1.m + p2.vz*p2.m;
    p1.ax = p1.ax*p1.m + p2.ax*p2.m; p1.ay = p1.ay*p1.m + p2.ay*p2.m; p1.az = p1.az*p1.m + p2.az*p2.m;
    p1.m += p2.m;
    if (p1.m > 0.){
        p1.x /= p1.m; p1.y /= p1.m; p1.z /= p1.m;
        p1.vx /= p1.m; p1.vy /= p1.m; p1.vz /= p1.m;
        p1.ax /= p1.m; p1.ay /= p1.m; p1.az /= p1.m;
    }
    return p1;
}

void reb_simulation_create_from_simulationarchive_with_messages(
        struct reb_simulation* r, struct reb_simulationarchive* sa, int64_t snapshot,
        enum reb_simulation_binary_error_codes* warnings){
    (void)sa; (void)snapshot; (void)warnings;
    reb_simulation_error(r, "SimulationArchive is not provided by the REBOUND shim.");
}

/* ------------------------------------------------------------------------ */
/* accelerations                                                            */
/* ------------------------------------------------------------------------ */

/* ASSIST runs REBOUND with gravity NONE (src/assist.c:441): REBOUND then only
 * zeroes the accelerations of ALL particles (real and variational -- the force
 * routines accumulate with += into both, src/forces.c:340-342, 425-427) and
 * calls the additional_forces plug-in. */
void reb_simulation_update_acceleration(struct reb_simulation* r){
    struct reb_particle* const particles = r->particles;
    const unsigned int N = r->N;
    for (unsigned int i = 0; i < N; i++){
        particles[i].ax = 0.; particles[i].ay = 0.; particles[i].az = 0.;
    }
    if (r->additional_forces){
        r->additional_forces(r);
        r->ri_ias15.b200_force_evals++;
    }
}

/* ------------------------------------------------------------------------ */
/* IAS15                                                                    */
/* ------------------------------------------------------------------------ */

static void alloc_dp7(struct reb_dp7* d, int N3){
    d->p0 = realloc(d->p0, sizeof(double) * N3); d->p1 = realloc(d->p1, sizeof(double) * N3);
    d->p2 = realloc(d->p2, sizeof(double) * N3); d->p3 = realloc(d->p3, sizeof(double) * N3);
    d->p4 = realloc(d->p4, sizeof(double) * N3); d->p5 = realloc(d->p5, sizeof(double) * N3);
    d->p6 = realloc(d->p6, sizeof(double) * N3);
    for (int k = 0; k < N3; k++){
        d->p0[k] = 0.; d->p1[k] = 0.; d->p2[k] = 0.; d->p3[k] = 0.; d->p4[k] = 0.; d->p5[k] = 0.; d->p6[k] = 0.;
    }
}

static double* realloc_zero(double* p, int n){
    p = realloc(p, sizeof(double) * n);
    for (int k = 0; k < n; k++) p[k] = 0.;
    return p;
}

/* REBOUND re-allocates (and zeroes b, e, csx, csv, ...) whenever N grows. */
static void ias15_alloc(struct reb_simulation* r){
    struct reb_integrator_ias15* ri = &r->ri_ias15;
    int N3 = 3 * (int)r->N;
    if (N3 > (int)(3 * ri->N_allocated)){
        alloc_dp7(&ri->g, N3); alloc_dp7(&ri->b, N3); alloc_dp7(&ri->csb, N3);
        alloc_dp7(&ri->e, N3); alloc_dp7(&ri->br, N3); alloc_dp7(&ri->er, N3);
        ri->at = realloc_zero(ri->at, N3); ri->x0 = realloc_zero(ri->x0, N3);
        ri->v0 = realloc_zero(ri->v0, N3); ri->a0 = realloc_zero(ri->a0, N3);
        ri->csx = realloc_zero(ri->csx, N3); ri->csv = realloc_zero(ri->csv, N3);
        ri->csa0 = realloc_zero(ri->csa0, N3);
        ri->N_allocated = r->N;
    }
}

/* Kahan-style compensated accumulation, REBOUND's add_cs. */
static inline void add_cs(double* p, double* csp, double inp){
    const double y = inp - *csp;
    const double t = *p + y;
    *csp = (t - *p) - y;
    *p = t;
}

/* Machine independent seventh root (REBOUND sqrt7). */
static double sqrt7(double a){
    double scale = 1;
    while (a < 1e-7 && isnormal(a)){ scale *= 0.1; a *= 1e7; }
    while (a > 1e2 && isnormal(a)){ scale *= 10; a *= 1e-7; }
    double x = 1.;
    for (int k = 0; k < 20; k++){
        double x6 = x*x*x*x*x*x;
        x += (a/x6 - x)/7.;
    }
    return x*scale;
}

static void predict_next_step(double ratio, int N3, const struct reb_dp7 _e, const struct reb_dp7 _b,
                              const struct reb_dp7 e, const struct reb_dp7 b){
    if (ratio > 20.){
        /* Do not predict if the step size increase is very large. */
        for (int k = 0; k < N3; ++k){
            e.p0[k] = 0.; e.p1[k] = 0.; e.p2[k] = 0.; e.p3[k] = 0.; e.p4[k] = 0.; e.p5[k] = 0.; e.p6[k] = 0.;
            b.p0[k] = 0.; b.p1[k] = 0.; b.p2[k] = 0.; b.p3[k] = 0.; b.p4[k] = 0.; b.p5[k] = 0.; b.p6[k] = 0.;
        }
    }else{
        const double q1 = ratio;
        const double q2 = q1 * q1;
        const double q3 = q1 * q2;
        const double q4 = q2 * q2;
        const double q5 = q2 * q3;
        const double q6 = q3 * q3;
        const double q7 = q3 * q4;
        for (int k = 0; k < N3; ++k){
            double be0 = _b.p0[k] - _e.p0[k];
            double be1 = _b.p1[k] - _e.p1[k];
            double be2 = _b.p2[k] - _e.p2[k];
            double be3 = _b.p3[k] - _e.p3[k];
            double be4 = _b.p4[k] - _e.p4[k];
            double be5 = _b.p5[k] - _e.p5[k];
            double be6 = _b.p6[k] - _e.p6[k];

            e.p0[k] = q1*(_b.p6[k]* 7.0 + _b.p5[k]* 6.0 + _b.p4[k]* 5.0 + _b.p3[k]* 4.0 + _b.p2[k]* 3.0 + _b.p1[k]*2.0 + _b.p0[k]);
            e.p1[k] = q2*(_b.p6[k]*21.0 + _b.p5[k]*15.0 + _b.p4[k]*10.0 + _b.p3[k]* 6.0 + _b.p2[k]* 3.0 + _b.p1[k]);
            e.p2[k] = q3*(_b.p6[k]*35.0 + _b.p5[k]*20.0 + _b.p4[k]*10.0 + _b.p3[k]* 4.0 + _b.p2[k]);
            e.p3[k] = q4*(_b.p6[k]*35.0 + _b.p5[k]*15.0 + _b.p4[k]* 5.0 + _b.p3[k]);
            e.p4[k] = q5*(_b.p6[k]*21.0 + _b.p5[k]* 6.0 + _b.p4[k]);
            e.p5[k] = q6*(_b.p6[k]* 7.0 + _b.p5[k]);
            e.p6[k] = q7* _b.p6[k];

            b.p0[k] = e.p0[k] + be0;
            b.p1[k] = e.p1[k] + be1;
            b.p2[k] = e.p2[k] + be2;
            b.p3[k] = e.p3[k] + be3;
            b.p4[k] = e.p4[k] + be4;
            b.p5[k] = e.p5[k] + be5;
            b.p6[k] = e.p6[k] + be6;
        }
    }
}

static void copybuffers(const struct reb_dp7 a, const struct reb_dp7 b, int N3){
    for (int k = 0; k < N3; k++){
        b.p0[k] = a.p0[k]; b.p1[k] = a.p1[k]; b.p2[k] = a.p2[k]; b.p3[k] = a.p3[k];
        b.p4[k] = a.p4[k]; b.p5[k] = a.p5[k]; b.p6[k] = a.p6[k];
    }
}

/* One IAS15 step attempt.  Returns 1 if accepted, 0 if rejected (retry with the
 * smaller r->dt).  Accelerations at the start of the step must already be in
 * particles[].a{x,y,z}. */
static int ias15_step(struct reb_simulation* r){
    ias15_alloc(r);
    struct reb_particle* const particles = r->particles;
    const int N = (int)r->N;
    const int N3 = 3*N;
    struct reb_integrator_ias15* ri = &r->ri_ias15;

    double* restrict const csx = ri->csx;
    double* restrict const csv = ri->csv;
    double* restrict const csa0 = ri->csa0;
    double* restrict const at = ri->at;
    double* restrict const x0 = ri->x0;
    double* restrict const v0 = ri->v0;
    double* restrict const a0 = ri->a0;
    const struct reb_dp7 g = ri->g;
    const struct reb_dp7 e = ri->e;
    const struct reb_dp7 b = ri->b;
    const struct reb_dp7 csb = ri->csb;
    const struct reb_dp7 er = ri->er;
    const struct reb_dp7 br = ri->br;

    for (int k = 0; k < N; k++){
        x0[3*k]   = particles[k].x;  x0[3*k+1] = particles[k].y;  x0[3*k+2] = particles[k].z;
        v0[3*k]   = particles[k].vx; v0[3*k+1] = particles[k].vy; v0[3*k+2] = particles[k].vz;
        a0[3*k]   = particles[k].ax; a0[3*k+1] = particles[k].ay; a0[3*k+2] = particles[k].az;
    }
    /* gravity is not COMPENSATED, so there is no compensation term for a0 */
    for (int k = 0; k < N3; k++) csa0[k] = 0;
    for (int k = 0; k < N3; k++){
        csb.p0[k] = 0.; csb.p1[k] = 0.; csb.p2[k] = 0.; csb.p3[k] = 0.;
        csb.p4[k] = 0.; csb.p5[k] = 0.; csb.p6[k] = 0.;
    }

    /* g from the b values predicted at the end of the previous step */
    for (int k = 0; k < N3; k++){
        g.p0[k] = b.p6[k]*dd[15] + b.p5[k]*dd[10] + b.p4[k]*dd[6] + b.p3[k]*dd[3]  + b.p2[k]*dd[1]  + b.p1[k]*dd[0]  + b.p0[k];
        g.p1[k] = b.p6[k]*dd[16] + b.p5[k]*dd[11] + b.p4[k]*dd[7] + b.p3[k]*dd[4]  + b.p2[k]*dd[2]  + b.p1[k];
        g.p2[k] = b.p6[k]*dd[17] + b.p5[k]*dd[12] + b.p4[k]*dd[8] + b.p3[k]*dd[5]  + b.p2[k];
        g.p3[k] = b.p6[k]*dd[18] + b.p5[k]*dd[13] + b.p4[k]*dd[9] + b.p3[k];
        g.p4[k] = b.p6[k]*dd[19] + b.p5[k]*dd[14] + b.p4[k];
        g.p5[k] = b.p6[k]*dd[20] + b.p5[k];
        g.p6[k] = b.p6[k];
    }

    double t_beginning = r->t;
    double predictor_corrector_error = 1e300;
    double predictor_corrector_error_last = 2;
    int iterations = 0;
    while (1){
        if (predictor_corrector_error < 1e-16) break;
        if (iterations > 2 && predictor_corrector_error_last <= predictor_corrector_error) break;
        if (iterations >= 12){
            ri->iterations_max_exceeded++;
            if (ri->iterations_max_exceeded == 10){
                reb_simulation_warning(r, "At least 10 predictor corrector loops in IAS15 did not converge. This is typically an indication of the timestep being too large.");
            }
            break;
        }
        predictor_corrector_error_last = predictor_corrector_error;
        predictor_corrector_error = 0;
        iterations++;
        ri->b200_pc_iterations++;

        for (int n = 1; n < 8; n++){
            r->t = t_beginning + r->dt * hh[n];

            if (shim_predictor_form == 1){
                /* s[] form: x = x0 + (s8 b6 + ... + s2 b0 + s1 a0 + s0 v0), compensated sums subtracted first */
                const double h = hh[n];
                double s[9], sv[8];
                s[0] = r->dt * h;
                s[1] = s[0] * s[0] / 2.;
                s[2] = s[1] * h / 3.;
                s[3] = s[2] * h / 2.;
                s[4] = 3. * s[3] * h / 5.;
                s[5] = 2. * s[4] * h / 3.;
                s[6] = 5. * s[5] * h / 7.;
                s[7] = 3. * s[6] * h / 4.;
                s[8] = 7. * s[7] * h / 9.;
                sv[0] = r->dt * h;
                sv[1] = sv[0] * h / 2.;
                sv[2] = 2. * sv[1] * h / 3.;
                sv[3] = 3. * sv[2] * h / 4.;
                sv[4] = 4. * sv[3] * h / 5.;
                sv[5] = 5. * sv[4] * h / 6.;
                sv[6] = 6. * sv[5] * h / 7.;
                sv[7] = 7. * sv[6] * h / 8.;
                for (int i = 0; i < N; i++){
                    double* const px[3] = {&particles[i].x, &particles[i].y, &particles[i].z};
                    double* const pv[3] = {&particles[i].vx, &particles[i].vy, &particles[i].vz};
                    for (int c = 0; c < 3; c++){
                        const int k = 3*i+c;
                        const double xk = -csx[k] + (s[8]*b.p6[k] + s[7]*b.p5[k] + s[6]*b.p4[k] + s[5]*b.p3[k] + s[4]*b.p2[k] + s[3]*b.p1[k] + s[2]*b.p0[k] + s[1]*a0[k] + s[0]*v0[k]);
                        *px[c] = xk + x0[k];
                        if (r->additional_forces && r->force_is_velocity_dependent){
                            const double vk = -csv[k] + (sv[7]*b.p6[k] + sv[6]*b.p5[k] + sv[5]*b.p4[k] + sv[4]*b.p3[k] + sv[3]*b.p2[k] + sv[2]*b.p1[k] + sv[1]*b.p0[k] + sv[0]*a0[k]);
                            *pv[c] = vk + v0[k];
                        }
                    }
                }
            } else {
            for (int i = 0; i < N; i++){
                const int k0 = 3*i+0, k1 = 3*i+1, k2 = 3*i+2;
                double xk0 = -csx[k0] + ((((((((b.p6[k0]*7.*hh[n]/9. + b.p5[k0])*3.*hh[n]/4. + b.p4[k0])*5.*hh[n]/7. + b.p3[k0])*2.*hh[n]/3. + b.p2[k0])*3.*hh[n]/5. + b.p1[k0])*hh[n]/2. + b.p0[k0])*hh[n]/3. + a0[k0])*r->dt*hh[n]/2. + v0[k0])*r->dt*hh[n];
                double xk1 = -csx[k1] + ((((((((b.p6[k1]*7.*hh[n]/9. + b.p5[k1])*3.*hh[n]/4. + b.p4[k1])*5.*hh[n]/7. + b.p3[k1])*2.*hh[n]/3. + b.p2[k1])*3.*hh[n]/5. + b.p1[k1])*hh[n]/2. + b.p0[k1])*hh[n]/3. + a0[k1])*r->dt*hh[n]/2. + v0[k1])*r->dt*hh[n];
                double xk2 = -csx[k2] + ((((((((b.p6[k2]*7.*hh[n]/9. + b.p5[k2])*3.*hh[n]/4. + b.p4[k2])*5.*hh[n]/7. + b.p3[k2])*2.*hh[n]/3. + b.p2[k2])*3.*hh[n]/5. + b.p1[k2])*hh[n]/2. + b.p0[k2])*hh[n]/3. + a0[k2])*r->dt*hh[n]/2. + v0[k2])*r->dt*hh[n];
                particles[i].x = xk0 + x0[k0];
                particles[i].y = xk1 + x0[k1];
                particles[i].z = xk2 + x0[k2];
            }
            if (r->additional_forces && r->force_is_velocity_dependent){
                for (int i = 0; i < N; i++){
                    const int k0 = 3*i+0, k1 = 3*i+1, k2 = 3*i+2;
                    double vk0 = -csv[k0] + (((((((b.p6[k0]*7.*hh[n]/8. + b.p5[k0])*6.*hh[n]/7. + b.p4[k0])*5.*hh[n]/6. + b.p3[k0])*4.*hh[n]/5. + b.p2[k0])*3.*hh[n]/4. + b.p1[k0])*2.*hh[n]/3. + b.p0[k0])*hh[n]/2. + a0[k0])*r->dt*hh[n];
                    double vk1 = -csv[k1] + (((((((b.p6[k1]*7.*hh[n]/8. + b.p5[k1])*6.*hh[n]/7. + b.p4[k1])*5.*hh[n]/6. + b.p3[k1])*4.*hh[n]/5. + b.p2[k1])*3.*hh[n]/4. + b.p1[k1])*2.*hh[n]/3. + b.p0[k1])*hh[n]/2. + a0[k1])*r->dt*hh[n];
                    double vk2 = -csv[k2] + (((((((b.p6[k2]*7.*hh[n]/8. + b.p5[k2])*6.*hh[n]/7. + b.p4[k2])*5.*hh[n]/6. + b.p3[k2])*4.*hh[n]/5. + b.p2[k2])*3.*hh[n]/4. + b.p1[k2])*2.*hh[n]/3. + b.p0[k2])*hh[n]/2. + a0[k2])*r->dt*hh[n];
                    particles[i].vx = vk0 + v0[k0];
                    particles[i].vy = vk1 + v0[k1];
                    particles[i].vz = vk2 + v0[k2];
                }
            }

            }
            reb_simulation_update_acceleration(r);

            for (int k = 0; k < N; ++k){
                at[3*k]   = particles[k].ax;
                at[3*k+1] = particles[k].ay;
                at[3*k+2] = particles[k].az;
            }
            switch (n){
                case 1:
                    for (int k = 0; k < N3; ++k){
                        double tmp = g.p0[k];
                        double gk = at[k];
                        double gk_cs = csa0[k];
                        add_cs(&gk, &gk_cs, -a0[k]);
                        g.p0[k] = gk/rr[0];
                        add_cs(&(b.p0[k]), &(csb.p0[k]), g.p0[k]-tmp);
                    } break;
                case 2:
                    for (int k = 0; k < N3; ++k){
                        double tmp = g.p1[k];
                        double gk = at[k];
                        double gk_cs = csa0[k];
                        add_cs(&gk, &gk_cs, -a0[k]);
                        g.p1[k] = (gk/rr[1] - g.p0[k])/rr[2];
                        tmp = g.p1[k] - tmp;
                        add_cs(&(b.p0[k]), &(csb.p0[k]), tmp * cc[0]);
                        add_cs(&(b.p1[k]), &(csb.p1[k]), tmp);
                    } break;
                case 3:
                    for (int k = 0; k < N3; ++k){
                        double tmp = g.p2[k];
                        double gk = at[k];
                        double gk_cs = csa0[k];
                        add_cs(&gk, &gk_cs, -a0[k]);
                        g.p2[k] = ((gk/rr[3] - g.p0[k])/rr[4] - g.p1[k])/rr[5];
                        tmp = g.p2[k] - tmp;
                        add_cs(&(b.p0[k]), &(csb.p0[k]), tmp * cc[1]);
                        add_cs(&(b.p1[k]), &(csb.p1[k]), tmp * cc[2]);
                        add_cs(&(b.p2[k]), &(csb.p2[k]), tmp);
                    } break;
                case 4:
                    for (int k = 0; k < N3; ++k){
                        double tmp = g.p3[k];
                        double gk = at[k];
                        double gk_cs = csa0[k];
                        add_cs(&gk, &gk_cs, -a0[k]);
                        g.p3[k] = (((gk/rr[6] - g.p0[k])/rr[7] - g.p1[k])/rr[8] - g.p2[k])/rr[9];
                        tmp = g.p3[k] - tmp;
                        add_cs(&(b.p0[k]), &(csb.p0[k]), tmp * cc[3]);
                        add_cs(&(b.p1[k]), &(csb.p1[k]), tmp * cc[4]);
                        add_cs(&(b.p2[k]), &(csb.p2[k]), tmp * cc[5]);
                        add_cs(&(b.p3[k]), &(csb.p3[k]), tmp);
                    } break;
                case 5:
                    for (int k = 0; k < N3; ++k){
                        double tmp = g.p4[k];
                        double gk = at[k];
                        double gk_cs = csa0[k];
                        add_cs(&gk, &gk_cs, -a0[k]);
                        g.p4[k] = ((((gk/rr[10] - g.p0[k])/rr[11] - g.p1[k])/rr[12] - g.p2[k])/rr[13] - g.p3[k])/rr[14];
                        tmp = g.p4[k] - tmp;
                        add_cs(&(b.p0[k]), &(csb.p0[k]), tmp * cc[6]);
                        add_cs(&(b.p1[k]), &(csb.p1[k]), tmp * cc[7]);
                        add_cs(&(b.p2[k]), &(csb.p2[k]), tmp * cc[8]);
                        add_cs(&(b.p3[k]), &(csb.p3[k]), tmp * cc[9]);
                        add_cs(&(b.p4[k]), &(csb.p4[k]), tmp);
                    } break;
                case 6:
                    for (int k = 0; k < N3; ++k){
                        double tmp = g.p5[k];
                        double gk = at[k];
                        double gk_cs = csa0[k];
                        add_cs(&gk, &gk_cs, -a0[k]);
                        g.p5[k] = (((((gk/rr[15] - g.p0[k])/rr[16] - g.p1[k])/rr[17] - g.p2[k])/rr[18] - g.p3[k])/rr[19] - g.p4[k])/rr[20];
                        tmp = g.p5[k] - tmp;
                        add_cs(&(b.p0[k]), &(csb.p0[k]), tmp * cc[10]);
                        add_cs(&(b.p1[k]), &(csb.p1[k]), tmp * cc[11]);
                        add_cs(&(b.p2[k]), &(csb.p2[k]), tmp * cc[12]);
                        add_cs(&(b.p3[k]), &(csb.p3[k]), tmp * cc[13]);
                        add_cs(&(b.p4[k]), &(csb.p4[k]), tmp * cc[14]);
                        add_cs(&(b.p5[k]), &(csb.p5[k]), tmp);
                    } break;
                case 7:
                {
                    double maxak = 0.0;
                    double maxb6ktmp = 0.0;
                    for (int k = 0; k < N3; ++k){
                        double tmp = g.p6[k];
                        double gk = at[k];
                        double gk_cs = csa0[k];
                        add_cs(&gk, &gk_cs, -a0[k]);
                        g.p6[k] = ((((((gk/rr[21] - g.p0[k])/rr[22] - g.p1[k])/rr[23] - g.p2[k])/rr[24] - g.p3[k])/rr[25] - g.p4[k])/rr[26] - g.p5[k])/rr[27];
                        tmp = g.p6[k] - tmp;
                        add_cs(&(b.p0[k]), &(csb.p0[k]), tmp * cc[15]);
                        add_cs(&(b.p1[k]), &(csb.p1[k]), tmp * cc[16]);
                        add_cs(&(b.p2[k]), &(csb.p2[k]), tmp * cc[17]);
                        add_cs(&(b.p3[k]), &(csb.p3[k]), tmp * cc[18]);
                        add_cs(&(b.p4[k]), &(csb.p4[k]), tmp * cc[19]);
                        add_cs(&(b.p5[k]), &(csb.p5[k]), tmp * cc[20]);
                        add_cs(&(b.p6[k]), &(csb.p6[k]), tmp);

                        /* convergence monitor: change of b6 relative to the acceleration */
                        if (ri->adaptive_mode != 0){
                            const double ak = fabs(at[k]);
                            if (isnormal(ak) && ak > maxak) maxak = ak;
                            const double b6ktmp = fabs(tmp);
                            if (isnormal(b6ktmp) && b6ktmp > maxb6ktmp) maxb6ktmp = b6ktmp;
                        }else{
                            const double errork = fabs(tmp/at[k]);
                            if (isnormal(errork) && errork > predictor_corrector_error) predictor_corrector_error = errork;
                        }
                    }
                    if (ri->adaptive_mode != 0){
                        predictor_corrector_error = maxb6ktmp/maxak;
                    }
                    break;
                }
            }
        }
    }
    r->t = t_beginning;
    const double dt_done = r->dt;

    if (ri->epsilon > 0){
        /* step size control from the last term of the series (real particles only) */
        const unsigned int Nreal = (unsigned int)(N - r->N_var);
        double integrator_error = 0.0;
        if (ri->adaptive_mode != 0){
            double maxa = 0.0;
            double maxj = 0.0;
            for (unsigned int i = 0; i < Nreal; i++){
                const double v2 = particles[i].vx*particles[i].vx + particles[i].vy*particles[i].vy + particles[i].vz*particles[i].vz;
                const double x2 = particles[i].x*particles[i].x + particles[i].y*particles[i].y + particles[i].z*particles[i].z;
                /* skip slowly varying accelerations */
                if (fabs(v2*r->dt*r->dt/x2) < 1e-16) continue;
                for (unsigned int k = 3*i; k < 3*(i+1); k++){
                    const double ak = fabs(at[k]);
                    if (isnormal(ak) && ak > maxa) maxa = ak;
                    const double b6k = fabs(b.p6[k]);
                    if (isnormal(b6k) && b6k > maxj) maxj = b6k;
                }
            }
            integrator_error = maxj/maxa;
        }else{
            for (int k = 0; k < N3; k++){
                const double errork = fabs(b.p6[k]/at[k]);
                if (isnormal(errork) && errork > integrator_error) integrator_error = errork;
            }
        }

        double dt_new;
        if (isnormal(integrator_error)){
            dt_new = sqrt7(ri->epsilon/integrator_error)*dt_done;
        }else{
            dt_new = dt_done/safety_factor;
        }
        if (fabs(dt_new) < ri->min_dt) dt_new = copysign(ri->min_dt, dt_new);

        if (fabs(dt_new/dt_done) < safety_factor){
            /* reject: restore and retry with the smaller step */
            for (int k = 0; k < N; ++k){
                particles[k].x = x0[3*k+0]; particles[k].y = x0[3*k+1]; particles[k].z = x0[3*k+2];
                particles[k].vx = v0[3*k+0]; particles[k].vy = v0[3*k+1]; particles[k].vz = v0[3*k+2];
                /* the retry starts from the accelerations at the beginning of the step */
                if (shim_reject_restores_acc){ particles[k].ax = a0[3*k+0]; particles[k].ay = a0[3*k+1]; particles[k].az = a0[3*k+2]; }
            }
            r->dt = dt_new;
            if (r->dt_last_done != 0.){
                double ratio = r->dt/r->dt_last_done;
                predict_next_step(ratio, N3, er, br, e, b);
            }
            ri->b200_steps_rejected++;
            return 0;
        }
        if (fabs(dt_new/dt_done) > 1.0){
            if (dt_new/dt_done > 1./safety_factor) dt_new = dt_done/safety_factor;
        }
        r->dt = dt_new;
    }

    /* new position and velocity at the end of the step */
    for (int k = 0; k < N3; ++k){
        add_cs(&(x0[k]), &(csx[k]), b.p6[k]/72.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), b.p5[k]/56.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), b.p4[k]/42.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), b.p3[k]/30.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), b.p2[k]/20.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), b.p1[k]/12.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), b.p0[k]/6.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), a0[k]/2.*dt_done*dt_done);
        add_cs(&(x0[k]), &(csx[k]), v0[k]*dt_done);
        add_cs(&(v0[k]), &(csv[k]), b.p6[k]/8.*dt_done);
        add_cs(&(v0[k]), &(csv[k]), b.p5[k]/7.*dt_done);
        add_cs(&(v0[k]), &(csv[k]), b.p4[k]/6.*dt_done);
        add_cs(&(v0[k]), &(csv[k]), b.p3[k]/5.*dt_done);
        add_cs(&(v0[k]), &(csv[k]), b.p2[k]/4.*dt_done);
        add_cs(&(v0[k]), &(csv[k]), b.p1[k]/3.*dt_done);
        add_cs(&(v0[k]), &(csv[k]), b.p0[k]/2.*dt_done);
        add_cs(&(v0[k]), &(csv[k]), a0[k]*dt_done);
    }

    r->t += dt_done;
    r->dt_last_done = dt_done;

    for (int k = 0; k < N; ++k){
        particles[k].x = x0[3*k+0]; particles[k].y = x0[3*k+1]; particles[k].z = x0[3*k+2];
        particles[k].vx = v0[3*k+0]; particles[k].vy = v0[3*k+1]; particles[k].vz = v0[3*k+2];
    }
    copybuffers(e, er, N3);
    copybuffers(b, br, N3);
    double ratio = r->dt/dt_done;
    predict_next_step(ratio, N3, e, b, e, b);
    return 1;
}

/* ------------------------------------------------------------------------ */
/* driver                                                                   */
/* ------------------------------------------------------------------------ */

void reb_simulation_step(struct reb_simulation* const r){
    if (r->pre_timestep_modifications){
        r->pre_timestep_modifications(r);
    }
    reb_simulation_update_acceleration(r);
    if (r->integrator == REB_INTEGRATOR_IAS15){
        while (!ias15_step(r)){
            if (r->status > 0 || r->messages_waiting) break;   /* e.g. ephemeris coverage error */
        }
    }
    if (r->post_timestep_modifications){
        r->post_timestep_modifications(r);
    }
    r->steps_done++;
}

void reb_simulation_steps(struct reb_simulation* const r, unsigned int N_steps){
    for (unsigned int i = 0; i < N_steps; i++) reb_simulation_step(r);
}

static int check_exit(struct reb_simulation* const r, const double tmax, double* last_full_dt){
    if (r->status <= REB_STATUS_SINGLE_STEP){
        if (r->status == REB_STATUS_SINGLE_STEP) r->status = REB_STATUS_PAUSED;
        else r->status++;
    }
    const double dtsign = copysign(1., r->dt);
    if (r->messages_waiting){
        r->status = REB_STATUS_GENERIC_ERROR;
    }
    if (r->status >= 0){
        /* exit now */
    }else if (tmax != INFINITY){
        if (r->exact_finish_time == 1){
            if ((r->t + r->dt)*dtsign >= tmax*dtsign){
                if (r->t == tmax){
                    r->status = REB_STATUS_SUCCESS;
                }else if (r->status == REB_STATUS_LAST_STEP){
                    double tscale = 1e-12*fabs(tmax);
                    if (tscale < 1e-200) tscale = 1e-12;
                    if (fabs(r->t - tmax) < tscale){
                        r->status = REB_STATUS_SUCCESS;
                    }else{
                        r->dt = tmax - r->t;
                    }
                }else{
                    r->status = REB_STATUS_LAST_STEP;
                    if (r->dt_last_done != 0.){
                        *last_full_dt = r->dt_last_done;
                    }
                    r->dt = tmax - r->t;
                }
            }else{
                if (r->status == REB_STATUS_LAST_STEP){
                    r->status = REB_STATUS_RUNNING;
                }
            }
        }else{
            if (r->t*dtsign >= tmax*dtsign){
                r->status = REB_STATUS_SUCCESS;
            }
        }
    }
    if (r->N <= 0){
        r->status = REB_STATUS_NO_PARTICLES;
    }
    return r->status;
}

enum REB_STATUS reb_simulation_integrate(struct reb_simulation* const r, double tmax){
    if (tmax != r->t){
        double dt_sign = (tmax > r->t) ? 1.0 : -1.0;
        r->dt = copysign(r->dt, dt_sign);
    }
    double last_full_dt = r->dt;
    r->dt_last_done = 0.;
    r->status = REB_STATUS_RUNNING;
    r->messages_waiting = 0;
    if (r->heartbeat) r->heartbeat(r);
    while (check_exit(r, tmax, &last_full_dt) < 0){
        reb_simulation_step(r);
        if (r->heartbeat) r->heartbeat(r);
    }
    if (r->exact_finish_time == 1){
        r->dt = last_full_dt;
    }
    return r->status;
}
