/*
 * reb_surface.cpp -- the REBOUND-compatible surface of include/rebound.h, product side.
 *
 * Bookkeeping (create / add / copy / variational configuration) is host code; every
 * function that computes -- reb_simulation_integrate, reb_simulation_step,
 * reb_simulation_update_acceleration -- marshals the simulation into a shared-step
 * GPU batch (assist_gpu.h) and runs the CUDA stepper.  Only simulations that ASSIST
 * has been attached to are supported (IAS15, gravity NONE, assist_additional_forces),
 * which is the configuration assist_init installs (reference src/assist.c:440-446).
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "assist.h"
#include "assist_gpu.h"
#include "host_internal.h"

static void ias15_free(struct reb_simulation* r);

/* ------------------------------------------------------------------------ */
/* bookkeeping                                                              */
/* ------------------------------------------------------------------------ */

extern "C" struct reb_simulation* reb_simulation_create(void) {
    struct reb_simulation* r = (struct reb_simulation*)calloc(1, sizeof(struct reb_simulation));
    r->G = 1.0;
    r->dt = 0.001;
    r->N_active = -1;
    r->status = REB_STATUS_PAUSED;
    r->exact_finish_time = 1;
    r->integrator = REB_INTEGRATOR_IAS15;
    r->gravity = REB_GRAVITY_BASIC;
    r->ri_ias15.epsilon = 1e-9;
    r->ri_ias15.min_dt = 0.0;
    r->ri_ias15.adaptive_mode = 2;
    return r;
}

extern "C" void reb_simulation_free(struct reb_simulation* const r) {
    if (r == NULL) return;
    if (r->extras_cleanup) r->extras_cleanup(r);
    ab_host_drop_batch(r);
    free(r->particles);
    free(r->var_config);
    free(r->messages);
    free(r->simulationarchive_filename);
    ias15_free(r);
    free(r);
}

extern "C" void reb_simulation_add(struct reb_simulation* const r, struct reb_particle pt) {
    if (r->N >= r->N_allocated) {
        r->N_allocated = r->N_allocated ? 2 * r->N_allocated : 128;
        r->particles = (struct reb_particle*)realloc(r->particles, sizeof(struct reb_particle) * r->N_allocated);
    }
    pt.sim = r;
    r->particles[r->N++] = pt;
}

extern "C" struct reb_simulation* reb_simulation_copy(struct reb_simulation* r) {
    struct reb_simulation* c = reb_simulation_create();
    c->t = r->t; c->G = r->G; c->dt = r->dt; c->dt_last_done = r->dt_last_done;
    c->exact_finish_time = r->exact_finish_time;
    c->force_is_velocity_dependent = r->force_is_velocity_dependent;
    c->integrator = r->integrator; c->gravity = r->gravity;
    c->ri_ias15.epsilon = r->ri_ias15.epsilon;
    c->ri_ias15.min_dt = r->ri_ias15.min_dt;
    c->ri_ias15.adaptive_mode = r->ri_ias15.adaptive_mode;
    c->N_active = r->N_active;
    for (unsigned int i = 0; i < r->N; i++) reb_simulation_add(c, r->particles[i]);
    c->N_var = r->N_var;
    c->N_var_config = r->N_var_config;
    if (r->N_var_config) {
        c->var_config = (struct reb_variational_configuration*)malloc(sizeof(struct reb_variational_configuration) * r->N_var_config);
        memcpy(c->var_config, r->var_config, sizeof(struct reb_variational_configuration) * r->N_var_config);
        for (unsigned int v = 0; v < c->N_var_config; v++) c->var_config[v].sim = c;
    }
    return c;
}

extern "C" void reb_simulation_add_fmt(struct reb_simulation* r, const char* fmt, ...) {
    struct reb_particle p;
    memset(&p, 0, sizeof(p));
    va_list args;
    va_start(args, fmt);
    char* copy = strdup(fmt);
    char* save = NULL;
    for (char* tok = strtok_r(copy, " ", &save); tok; tok = strtok_r(NULL, " ", &save)) {
        const double v = va_arg(args, double);
        if (!strcmp(tok, "x")) p.x = v;
        else if (!strcmp(tok, "y")) p.y = v;
        else if (!strcmp(tok, "z")) p.z = v;
        else if (!strcmp(tok, "vx")) p.vx = v;
        else if (!strcmp(tok, "vy")) p.vy = v;
        else if (!strcmp(tok, "vz")) p.vz = v;
        else if (!strcmp(tok, "m")) p.m = v;
        else if (!strcmp(tok, "r")) p.r = v;
        else reb_simulation_error(r, "reb_simulation_add_fmt: only the keys x y z vx vy vz m r are supported.");
    }
    free(copy);
    va_end(args);
    reb_simulation_add(r, p);
}

extern "C" int reb_simulation_add_variation_1st_order(struct reb_simulation* const r, int testparticle) {
    if (testparticle < 0) {
        reb_simulation_error(r, "Variations of all particles at once (testparticle<0) are not supported.");
        return -1;
    }
    r->N_var_config++;
    r->var_config = (struct reb_variational_configuration*)realloc(r->var_config, sizeof(struct reb_variational_configuration) * r->N_var_config);
    struct reb_variational_configuration* vc = &r->var_config[r->N_var_config - 1];
    memset(vc, 0, sizeof(*vc));
    vc->sim = r;
    vc->order = 1;
    vc->index = (int)r->N;
    vc->testparticle = testparticle;
    struct reb_particle p0;
    memset(&p0, 0, sizeof(p0));
    reb_simulation_add(r, p0);
    r->N_var++;
    return vc->index;
}

extern "C" void reb_simulation_error(struct reb_simulation* const r, const char* const msg) {
    fprintf(stderr, "\n(assist-b200) Error: %s\n", msg);
    if (r) {
        free(r->messages);
        r->messages = strdup(msg);
        r->messages_waiting = 1;
    }
}

extern "C" void reb_simulation_warning(struct reb_simulation* const r, const char* const msg) {
    (void)r;
    fprintf(stderr, "\n(assist-b200) Warning: %s\n", msg);
}

extern "C" void reb_particle_iadd(struct reb_particle* p1, struct reb_particle* p2) {
    p1->x += p2->x; p1->y += p2->y; p1->z += p2->z;
    p1->vx += p2->vx; p1->vy += p2->vy; p1->vz += p2->vz;
    p1->m += p2->m;
}

extern "C" void reb_particle_isub(struct reb_particle* p1, struct reb_particle* p2) {
    p1->x -= p2->x; p1->y -= p2->y; p1->z -= p2->z;
    p1->vx -= p2->vx; p1->vy -= p2->vy; p1->vz -= p2->vz;
    p1->m -= p2->m;
}

extern "C" double reb_particle_distance(struct reb_particle* p1, struct reb_particle* p2) {
    const double dx = p1->x - p2->x, dy = p1->y - p2->y, dz = p1->z - p2->z;
    return sqrt(dx * dx + dy * dy + dz * dz);
}

extern "C" struct reb_particle reb_particle_com_of_pair(struct reb_particle p1, struct reb_particle p2) {
    p1.x = p1.x * p1.m + p2.x * p2.m; p1.y = p1.y * p1.m + p2.y * p2.m; p1.z = p1.z * p1.m + p2.z * p2.m;
    p1.vx = p1.vx * p1.m + p2.vx * p2.m; p1.vy = p1.vy * p1.m + p2.vy * p2.m; p1.vz = p1.vz * p1.m + p2.vz * p2.m;
    p1.ax = p1.ax * p1.m + p2.ax * p2.m; p1.ay = p1.ay * p1.m + p2.ay * p2.m; p1.az = p1.az * p1.m + p2.az * p2.m;
    p1.m += p2.m;
    if (p1.m > 0.) {
        p1.x /= p1.m; p1.y /= p1.m; p1.z /= p1.m;
        p1.vx /= p1.m; p1.vy /= p1.m; p1.vz /= p1.m;
        p1.ax /= p1.m; p1.ay /= p1.m; p1.az /= p1.m;
    }
    return p1;
}

/* ------------------------------------------------------------------------ */
/* marshalling between struct reb_simulation and a shared-step GPU batch    */
/* ------------------------------------------------------------------------ */

struct AbHostBatch {
    assist_gpu_batch* gb;
    unsigned int N;
    int N_var;
    unsigned int N_var_config;
    int n_real, K;
    std::vector<int> nv;                       /* variational particles per real particle */
    std::vector<int> var_pidx;                 /* [n_real][K-1] particle index of each slot, -1 if empty */
    std::vector<int> var_cfg;                  /* [n_real][K-1] var_config index of each slot */
    std::vector<int> cfg_index, cfg_tp;        /* var_config snapshot to detect edits */
};

extern "C" void ab_host_drop_batch(struct reb_simulation* r) {
    AbHostBatch* hb = (AbHostBatch*)r->b200_batch;
    if (!hb) return;
    assist_gpu_batch_free(hb->gb);
    delete hb;
    r->b200_batch = NULL;
}

static int fail(struct reb_simulation* r, const char* msg) {
    reb_simulation_error(r, msg);
    r->status = REB_STATUS_GENERIC_ERROR;
    return -1;
}

static struct assist_extras* attached_extras(struct reb_simulation* r) {
    if (r->extras == NULL || r->additional_forces != assist_additional_forces) return NULL;
    return (struct assist_extras*)r->extras;
}

/* (Re)build the layout if the particle set changed; returns NULL on failure. */
static AbHostBatch* ensure_layout(struct reb_simulation* r, struct assist_extras* ax, bool* fresh) {
    AbHostBatch* hb = (AbHostBatch*)r->b200_batch;
    bool same = hb && hb->N == r->N && hb->N_var == r->N_var && hb->N_var_config == r->N_var_config;
    if (same) {
        for (unsigned int v = 0; v < r->N_var_config; v++)
            if (hb->cfg_index[v] != r->var_config[v].index || hb->cfg_tp[v] != r->var_config[v].testparticle) same = false;
    }
    *fresh = !same;
    if (same) return hb;
    ab_host_drop_batch(r);
    const int n_real = (int)r->N - r->N_var;
    if (n_real <= 0) { fail(r, "No real particles in simulation."); return NULL; }
    hb = new AbHostBatch();
    hb->N = r->N; hb->N_var = r->N_var; hb->N_var_config = r->N_var_config; hb->n_real = n_real;
    hb->nv.assign(n_real, 0);
    for (unsigned int v = 0; v < r->N_var_config; v++) {
        const struct reb_variational_configuration& vc = r->var_config[v];
        hb->cfg_index.push_back(vc.index);
        hb->cfg_tp.push_back(vc.testparticle);
        if (vc.order != 1 || vc.testparticle < 0 || vc.testparticle >= n_real || vc.index < n_real || vc.index >= (int)r->N) {
            delete hb;
            fail(r, "Unsupported variational configuration (only first-order variations of single test particles).");
            return NULL;
        }
        hb->nv[vc.testparticle]++;
    }
    int nvmax = 0;
    for (int i = 0; i < n_real; i++) if (hb->nv[i] > nvmax) nvmax = hb->nv[i];
    if (nvmax > ASSIST_GPU_MAX_NVAR) {
        delete hb;
        fail(r, "More variational particles per test particle than the GPU kernels are built for.");
        return NULL;
    }
    hb->K = 1 + nvmax;
    hb->var_pidx.assign((size_t)n_real * nvmax, -1);
    hb->var_cfg.assign((size_t)n_real * nvmax, -1);
    std::vector<int> fill(n_real, 0);
    for (unsigned int v = 0; v < r->N_var_config; v++) {
        const int tp = r->var_config[v].testparticle;
        const int slot = fill[tp]++;
        hb->var_pidx[(size_t)tp * nvmax + slot] = r->var_config[v].index;
        hb->var_cfg[(size_t)tp * nvmax + slot] = (int)v;
    }
    hb->gb = assist_gpu_batch_create(ax->ephem, n_real, nvmax, ASSIST_GPU_SHARED_STEP);
    if (!hb->gb) {
        delete hb;
        fail(r, assist_gpu_last_error());
        return NULL;
    }
    r->b200_batch = hb;
    return hb;
}

static void gather_state(const struct reb_simulation* r, const AbHostBatch* hb, std::vector<double>& st) {
    const int K = hb->K, nvmax = K - 1;
    st.assign((size_t)hb->n_real * K * 6, 0.0);
    for (int i = 0; i < hb->n_real; i++) {
        for (int j = 0; j < K; j++) {
            int pidx = (j == 0) ? i : hb->var_pidx[(size_t)i * nvmax + (j - 1)];
            if (pidx < 0) continue;
            const struct reb_particle& p = r->particles[pidx];
            double* o = &st[((size_t)i * K + j) * 6];
            o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.vx; o[4] = p.vy; o[5] = p.vz;
        }
    }
}

/* particle_params rows: real particle j -> row j; variational slot -> row N_real + var_config index
 * (reference src/forces.c:839-841, 1030-1032). */
static bool gather_params(const struct assist_extras* ax, const AbHostBatch* hb, std::vector<double>& prm) {
    if (ax->particle_params == NULL) return false;
    const int K = hb->K, nvmax = K - 1;
    prm.assign((size_t)hb->n_real * K * 3, 0.0);
    for (int i = 0; i < hb->n_real; i++) {
        for (int j = 0; j < K; j++) {
            int row;
            if (j == 0) row = i;
            else {
                const int cfg = hb->var_cfg[(size_t)i * nvmax + (j - 1)];
                if (cfg < 0) continue;
                row = hb->n_real + cfg;
            }
            for (int c = 0; c < 3; c++) prm[((size_t)i * K + j) * 3 + c] = ax->particle_params[3 * row + c];
        }
    }
    return true;
}

extern "C" void ab_host_fill_options(const struct reb_simulation* r, const struct assist_extras* ax, struct assist_gpu_options* opt) {
    assist_gpu_default_options(opt);
    opt->forces = ax->forces;
    opt->gr_eih_sources = ax->gr_eih_sources;
    opt->geocentric = ax->geocentric;
    opt->alpha = ax->alpha; opt->nk = ax->nk; opt->nm = ax->nm; opt->nn = ax->nn; opt->r0 = ax->r0;
    opt->epsilon = r->ri_ias15.epsilon;
    opt->min_dt = r->ri_ias15.min_dt;
    const char* m = getenv("ASSIST_B200_MATH");
    opt->math = (m && (!strcmp(m, "fast") || !strcmp(m, "FAST"))) ? ASSIST_GPU_MATH_FAST : ASSIST_GPU_MATH_STRICT;
}

static void scatter_state(struct reb_simulation* r, const AbHostBatch* hb, const std::vector<double>& st, const std::vector<double>* acc,
                          struct reb_particle* dest) {
    const int K = hb->K, nvmax = K - 1;
    for (int i = 0; i < hb->n_real; i++) {
        for (int j = 0; j < K; j++) {
            int pidx = (j == 0) ? i : hb->var_pidx[(size_t)i * nvmax + (j - 1)];
            if (pidx < 0) continue;
            struct reb_particle& p = dest[pidx];
            const double* o = &st[((size_t)i * K + j) * 6];
            p.x = o[0]; p.y = o[1]; p.z = o[2]; p.vx = o[3]; p.vy = o[4]; p.vz = o[5];
            if (acc) {
                const double* a = &(*acc)[((size_t)i * K + j) * 3];
                p.ax = a[0]; p.ay = a[1]; p.az = a[2];
            }
        }
    }
}

/* Push the simulation to the device; returns the batch or NULL (error already raised). */
static AbHostBatch* sync_to_device(struct reb_simulation* r) {
    struct assist_extras* ax = attached_extras(r);
    if (!ax) { fail(r, "assist-b200 integrates only simulations that ASSIST is attached to (assist_attach)."); return NULL; }
    if (r->integrator != REB_INTEGRATOR_IAS15 || r->gravity != REB_GRAVITY_NONE || r->ri_ias15.adaptive_mode != 1) {
        fail(r, "assist-b200 supports the configuration assist_attach installs: IAS15, gravity NONE, adaptive_mode 1.");
        return NULL;
    }
    if (r->pre_timestep_modifications && r->pre_timestep_modifications != ab_host_pre_timestep_marker) {
        fail(r, "User pre_timestep_modifications callbacks cannot run inside the GPU stepper.");
        return NULL;
    }
    if (r->post_timestep_modifications) { fail(r, "post_timestep_modifications callbacks cannot run inside the GPU stepper."); return NULL; }
    bool fresh = false;
    AbHostBatch* hb = ensure_layout(r, ax, &fresh);
    if (!hb) return NULL;
    struct assist_gpu_options opt;
    ab_host_fill_options(r, ax, &opt);
    std::vector<double> st, prm;
    gather_state(r, hb, st);
    const bool has_prm = gather_params(ax, hb, prm);
    int rc;
    if (fresh) {
        rc = assist_gpu_batch_set_state(hb->gb, r->t, r->dt, st.data(), has_prm ? prm.data() : NULL, hb->nv.data());
    } else {
        /* same particle set: keep the IAS15 history (b, e, compensation terms), take the user's particles */
        rc = assist_gpu_batch_update_particles(hb->gb, st.data());
        if (!rc) rc = assist_gpu_batch_set_time(hb->gb, r->t, r->dt);
        if (!rc) rc = ab_gpu_batch_update_params(hb->gb, has_prm ? prm.data() : NULL);
    }
    if (!rc) rc = assist_gpu_batch_set_options(hb->gb, &opt);
    if (rc) { fail(r, assist_gpu_last_error()); return NULL; }
    return hb;
}

static int sync_from_device(struct reb_simulation* r, AbHostBatch* hb) {
    std::vector<double> st((size_t)hb->n_real * hb->K * 6), acc((size_t)hb->n_real * hb->K * 3);
    double t, dt, dtl;
    int status;
    if (assist_gpu_batch_get_state(hb->gb, st.data(), acc.data(), &t, &dt, &dtl, &status)) return fail(r, assist_gpu_last_error());
    scatter_state(r, hb, st, &acc, r->particles);
    r->t = t; r->dt = dt; r->dt_last_done = dtl;
    r->status = (enum REB_STATUS)status;
    struct assist_gpu_stats s;
    if (!assist_gpu_batch_get_stats(hb->gb, &s)) {
        r->steps_done = s.steps;
        r->ri_ias15.b200_pc_iterations = s.pc_iterations;
        r->ri_ias15.b200_force_evals = s.force_evals / (unsigned long long)hb->n_real;
        r->ri_ias15.b200_steps_rejected = s.steps_rejected;
    }
    /* dense-output anchor for assist_integrate_or_interpolate (reference src/assist.c:754-758) */
    struct assist_extras* ax = attached_extras(r);
    if (ax && ax->last_state) {
        std::vector<double> ls((size_t)hb->n_real * hb->K * 6), la((size_t)hb->n_real * hb->K * 3);
        if (!ab_gpu_batch_get_last_state(hb->gb, ls.data(), la.data())) scatter_state(r, hb, ls, &la, ax->last_state);
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* snapshots of an attached simulation (SimulationArchive)                  */
/* ------------------------------------------------------------------------ */
/* The reference reads snapshots back through REBOUND's SimulationArchive (src/assist.c:599-633: two consecutive
 * snapshots give x0, v0 of a step's start and a0, br, dt_last_done of the step).  REBOUND's binary format is not
 * available here, so the file layout is assist-b200's own; the API (reb_simulation_save_to_file_step,
 * reb_simulationarchive_create_from_file, sa->t / sa->nblobs, reb_simulation_create_from_simulationarchive_with_messages)
 * is REBOUND's.  File: 16-byte magic, then per snapshot
 *     uint64 payload bytes | double t, dt, dt_last_done | uint64 steps_done | uint32 N, N_var, N_var_config, 0 |
 *     N_var_config x (int32 index, int32 testparticle) | N x reb_particle | a0[3N] | br[7][3N]
 * a0 = acceleration at the start of the last completed step, br = its b coefficients (zeros before the first step). */
static const char AB_SA_MAGIC[16] = {'A', 'S', 'S', 'I', 'S', 'T', '-', 'B', '2', '0', '0', ' ', 'S', 'A', '1', '\n'};

struct AbSnapHead {
    double t, dt, dt_last_done;
    uint64_t steps_done;
    uint32_t N, N_var, N_var_config, zero;
};

static void ias15_alloc(struct reb_simulation* r, unsigned int N) {
    struct reb_integrator_ias15* ri = &r->ri_ias15;
    if (ri->N_allocated >= N && ri->x0) return;
    const size_t m = 3 * (size_t)N;
    double** one[] = {&ri->x0, &ri->v0, &ri->a0};
    for (double** p : one) { free(*p); *p = (double*)calloc(m, sizeof(double)); }
    double** seven[] = {&ri->br.p0, &ri->br.p1, &ri->br.p2, &ri->br.p3, &ri->br.p4, &ri->br.p5, &ri->br.p6};
    for (double** p : seven) { free(*p); *p = (double*)calloc(m, sizeof(double)); }
    ri->N_allocated = N;
}

static void ias15_free(struct reb_simulation* r) {
    struct reb_integrator_ias15* ri = &r->ri_ias15;
    free(ri->x0); free(ri->v0); free(ri->a0);
    free(ri->br.p0); free(ri->br.p1); free(ri->br.p2); free(ri->br.p3); free(ri->br.p4); free(ri->br.p5); free(ri->br.p6);
    ri->x0 = ri->v0 = ri->a0 = NULL;
    ri->br.p0 = ri->br.p1 = ri->br.p2 = ri->br.p3 = ri->br.p4 = ri->br.p5 = ri->br.p6 = NULL;
    ri->N_allocated = 0;
}

extern "C" void reb_simulation_save_to_file_step(struct reb_simulation* const r, const char* filename, unsigned long long step) {
    if (!r || !filename) return;
    free(r->simulationarchive_filename);
    r->simulationarchive_filename = strdup(filename);
    if (r->simulationarchive_auto_step != step) {
        r->simulationarchive_auto_step = step;
        r->simulationarchive_next_step = r->steps_done;
    }
}

/* Appends a snapshot of `r` as it stands (the device batch, if any, supplies a0 and br of the last completed step). */
extern "C" void reb_simulation_save_to_file(struct reb_simulation* const r, const char* filename) {
    if (!r) return;
    if (!filename) filename = r->simulationarchive_filename;
    if (!filename) { reb_simulation_error(r, "reb_simulation_save_to_file: no file name."); return; }
    const size_t m = 3 * (size_t)r->N;
    std::vector<double> a0(m, 0.0), br(7 * m, 0.0);
    AbHostBatch* hb = (AbHostBatch*)r->b200_batch;
    if (!hb && r->ri_ias15.a0 && r->ri_ias15.br.p0 && r->ri_ias15.N_allocated >= r->N) {
        /* a simulation restored from a snapshot: it carries the step data itself */
        const struct reb_integrator_ias15* ri = &r->ri_ias15;
        const double* seven[7] = {ri->br.p0, ri->br.p1, ri->br.p2, ri->br.p3, ri->br.p4, ri->br.p5, ri->br.p6};
        memcpy(a0.data(), ri->a0, sizeof(double) * m);
        for (int q = 0; q < 7; q++) memcpy(&br[q * m], seven[q], sizeof(double) * m);
    }
    if (hb && r->steps_done > 0) {
        const size_t per = (size_t)hb->n_real * hb->K;
        std::vector<double> la(per * 3), bb(7 * per * 3);
        if (ab_gpu_batch_get_last_state(hb->gb, NULL, la.data()) || ab_gpu_batch_get_br(hb->gb, bb.data())) { fail(r, assist_gpu_last_error()); return; }
        const int nvmax = hb->K - 1;
        for (int i = 0; i < hb->n_real; i++)
            for (int j = 0; j < hb->K; j++) {
                const int pidx = (j == 0) ? i : hb->var_pidx[(size_t)i * nvmax + (j - 1)];
                if (pidx < 0) continue;
                for (int c = 0; c < 3; c++) {
                    a0[3 * (size_t)pidx + c] = la[((size_t)i * hb->K + j) * 3 + c];
                    for (int q = 0; q < 7; q++) br[q * m + 3 * (size_t)pidx + c] = bb[(size_t)q * per * 3 + ((size_t)i * hb->K + j) * 3 + c];
                }
            }
    }
    FILE* f = fopen(filename, "rb");
    const bool exists = f != NULL;
    if (f) fclose(f);
    f = fopen(filename, "ab");
    if (!f) { reb_simulation_error(r, "reb_simulation_save_to_file: cannot open the file."); return; }
    size_t bad = 0;      /* short writes (disk full): reported once at the end */
#define AB_WR(ptr, size, count) do { if (fwrite((ptr), (size), (count), f) != (size_t)(count)) bad++; } while (0)
    if (!exists) AB_WR(AB_SA_MAGIC, 1, sizeof(AB_SA_MAGIC));
    AbSnapHead h;
    h.t = r->t; h.dt = r->dt; h.dt_last_done = r->dt_last_done; h.steps_done = r->steps_done;
    h.N = r->N; h.N_var = (uint32_t)r->N_var; h.N_var_config = r->N_var_config; h.zero = 0;
    const uint64_t bytes = sizeof(h) + 8 * (uint64_t)r->N_var_config + sizeof(struct reb_particle) * (uint64_t)r->N + 8 * (uint64_t)(8 * m);
    AB_WR(&bytes, 8, 1);
    AB_WR(&h, sizeof(h), 1);
    for (unsigned int v = 0; v < r->N_var_config; v++) {
        const int32_t pair[2] = {r->var_config[v].index, r->var_config[v].testparticle};
        AB_WR(pair, 4, 2);
    }
    for (unsigned int i = 0; i < r->N; i++) {
        struct reb_particle p = r->particles[i];
        p.sim = NULL; p.c = NULL; p.ap = NULL;      /* pointers mean nothing in a file */
        AB_WR(&p, sizeof(p), 1);
    }
    AB_WR(a0.data(), 8, m);
    AB_WR(br.data(), 8, 7 * m);
#undef AB_WR
    if (fclose(f) != 0) bad++;
    if (bad) reb_simulation_error(r, "reb_simulation_save_to_file: the snapshot could not be written completely.");
}

/* what REBOUND's heartbeat does for the archive: before the first step and after every completed step */
static void archive_heartbeat(struct reb_simulation* r) {
    if (r->simulationarchive_auto_step && r->simulationarchive_next_step <= r->steps_done) {
        r->simulationarchive_next_step += r->simulationarchive_auto_step;
        reb_simulation_save_to_file(r, NULL);
    }
}

extern "C" struct reb_simulationarchive* reb_simulationarchive_create_from_file(const char* filename) {
    FILE* f = filename ? fopen(filename, "rb") : NULL;
    if (!f) return NULL;
    char magic[16];
    if (fread(magic, 1, 16, f) != 16 || memcmp(magic, AB_SA_MAGIC, 16)) {
        fprintf(stderr, "\n(assist-b200) Error: %s is not a snapshot file written by reb_simulation_save_to_file.\n", filename);
        fclose(f);
        return NULL;
    }
    std::vector<double> t;
    std::vector<long> off;
    for (;;) {
        const long here = ftell(f);
        uint64_t bytes;
        AbSnapHead h;
        if (fread(&bytes, 8, 1, f) != 1) break;
        if (bytes < sizeof(h) || fread(&h, sizeof(h), 1, f) != 1) break;
        if (fseek(f, (long)(bytes - sizeof(h)), SEEK_CUR)) break;
        if (ftell(f) != here + 8 + (long)bytes) break;
        /* a truncated last snapshot is dropped */
        {
            const long end_of_blob = ftell(f);
            fseek(f, 0, SEEK_END);
            const long end = ftell(f);
            if (end_of_blob > end) break;
            fseek(f, end_of_blob, SEEK_SET);
        }
        t.push_back(h.t);
        off.push_back(here);
    }
    fclose(f);
    struct reb_simulationarchive* sa = (struct reb_simulationarchive*)calloc(1, sizeof(*sa));
    sa->filename = strdup(filename);
    sa->nblobs = (long)t.size();
    sa->t = (double*)malloc(sizeof(double) * (t.size() + 1));
    sa->b200_offset = (long*)malloc(sizeof(long) * (off.size() + 1));
    for (size_t i = 0; i < t.size(); i++) { sa->t[i] = t[i]; sa->b200_offset[i] = off[i]; }
    return sa;
}

extern "C" void reb_simulationarchive_free(struct reb_simulationarchive* sa) {
    if (!sa) return;
    free(sa->filename); free(sa->t); free(sa->b200_offset);
    free(sa);
}

extern "C" void reb_simulation_create_from_simulationarchive_with_messages(
        struct reb_simulation* r, struct reb_simulationarchive* sa, int64_t snapshot,
        enum reb_simulation_binary_error_codes* warnings) {
    if (warnings) *warnings = REB_SIMULATION_BINARY_WARNING_NONE;
    if (!r || !sa) return;
    if (snapshot < 0) snapshot += sa->nblobs;
    if (snapshot < 0 || snapshot >= sa->nblobs) { reb_simulation_error(r, "Snapshot index out of range."); return; }
    FILE* f = fopen(sa->filename, "rb");
    if (!f) { reb_simulation_error(r, "Cannot open the snapshot file."); return; }
    uint64_t bytes;
    AbSnapHead h;
    bool ok = !fseek(f, sa->b200_offset[snapshot], SEEK_SET) && fread(&bytes, 8, 1, f) == 1 && fread(&h, sizeof(h), 1, f) == 1;
    if (ok) {
        r->t = h.t; r->dt = h.dt; r->dt_last_done = h.dt_last_done; r->steps_done = h.steps_done;
        r->integrator = REB_INTEGRATOR_IAS15; r->gravity = REB_GRAVITY_NONE; r->force_is_velocity_dependent = 1;
        r->ri_ias15.adaptive_mode = 1;
        r->N = 0; r->N_var = 0;
        free(r->var_config);
        r->var_config = NULL;
        r->N_var_config = h.N_var_config;
        if (h.N_var_config) r->var_config = (struct reb_variational_configuration*)calloc(h.N_var_config, sizeof(struct reb_variational_configuration));
        for (unsigned int v = 0; ok && v < h.N_var_config; v++) {
            int32_t pair[2];
            ok = fread(pair, 4, 2, f) == 2;
            r->var_config[v].sim = r; r->var_config[v].order = 1; r->var_config[v].index = pair[0]; r->var_config[v].testparticle = pair[1];
        }
        for (unsigned int i = 0; ok && i < h.N; i++) {
            struct reb_particle p;
            ok = fread(&p, sizeof(p), 1, f) == 1;
            if (ok) reb_simulation_add(r, p);
        }
        r->N_var = (int)h.N_var;
        const size_t m = 3 * (size_t)h.N;
        ias15_alloc(r, h.N);
        struct reb_integrator_ias15* ri = &r->ri_ias15;
        ok = ok && fread(ri->a0, 8, m, f) == m;
        double* seven[] = {ri->br.p0, ri->br.p1, ri->br.p2, ri->br.p3, ri->br.p4, ri->br.p5, ri->br.p6};
        for (double* p : seven) ok = ok && fread(p, 8, m, f) == m;
        /* as REBOUND leaves them after a completed step: x0, v0 = the particles' state */
        for (unsigned int i = 0; ok && i < h.N; i++) {
            ri->x0[3 * i] = r->particles[i].x; ri->x0[3 * i + 1] = r->particles[i].y; ri->x0[3 * i + 2] = r->particles[i].z;
            ri->v0[3 * i] = r->particles[i].vx; ri->v0[3 * i + 1] = r->particles[i].vy; ri->v0[3 * i + 2] = r->particles[i].vz;
        }
    }
    fclose(f);
    if (!ok) reb_simulation_error(r, "Snapshot file is truncated or corrupt.");
}

/* ------------------------------------------------------------------------ */
/* compute entry points                                                     */
/* ------------------------------------------------------------------------ */

extern "C" enum REB_STATUS reb_simulation_integrate(struct reb_simulation* const r, double tmax) {
    r->messages_waiting = 0;
    if (r->N == 0) { r->status = REB_STATUS_NO_PARTICLES; return r->status; }
    AbHostBatch* hb = sync_to_device(r);
    if (!hb) return r->status;
    int rc;
    if (r->heartbeat || r->simulationarchive_auto_step) {
        /* one accepted step per launch so the callback (and the snapshot file) sees every step */
        if (r->heartbeat) r->heartbeat(r);
        archive_heartbeat(r);
        int resume = 0;
        while (true) {
            rc = ab_gpu_batch_integrate_ex(hb->gb, tmax, r->exact_finish_time, 1, resume);
            if (rc) break;
            if (sync_from_device(r, hb)) return r->status;
            if (r->heartbeat) r->heartbeat(r);
            archive_heartbeat(r);
            if (r->status >= 0) break;
            resume = 1;
        }
    } else {
        rc = assist_gpu_batch_integrate(hb->gb, tmax, r->exact_finish_time, 0);
    }
    if (rc > 0) {   /* ephemeris error inside a force evaluation (reference src/forces.c:317-321) */
        reb_simulation_error(r, assist_error_messages[rc]);
        sync_from_device(r, hb);
        r->status = REB_STATUS_GENERIC_ERROR;
        return r->status;
    }
    if (rc < 0) { fail(r, assist_gpu_last_error()); return r->status; }
    sync_from_device(r, hb);
    return r->status;
}

extern "C" void reb_simulation_step(struct reb_simulation* const r) {
    if (r->N == 0) return;
    AbHostBatch* hb = sync_to_device(r);
    if (!hb) return;
    const enum REB_STATUS before = r->status;
    int rc = ab_gpu_batch_integrate_ex(hb->gb, 0.0, 0, 1, 2 /* one raw step, no exit logic */);
    if (rc > 0) { reb_simulation_error(r, assist_error_messages[rc]); r->status = REB_STATUS_GENERIC_ERROR; return; }
    if (rc < 0) { fail(r, assist_gpu_last_error()); return; }
    sync_from_device(r, hb);
    r->status = before;
}

extern "C" void reb_simulation_steps(struct reb_simulation* const r, unsigned int N_steps) {
    for (unsigned int i = 0; i < N_steps; i++) reb_simulation_step(r);
}

/* Zero the accelerations (gravity NONE) and run the force plug-in on the GPU. */
extern "C" void reb_simulation_update_acceleration(struct reb_simulation* r) {
    for (unsigned int i = 0; i < r->N; i++) { r->particles[i].ax = 0.; r->particles[i].ay = 0.; r->particles[i].az = 0.; }
    if (r->additional_forces) r->additional_forces(r);
}

/* assist_additional_forces: the reference's plug-in entry point (src/forces.c:49-173).
 * Adds the accelerations of every enabled force term to sim->particles[].a{x,y,z}. */
extern "C" void assist_additional_forces(struct reb_simulation* sim) {
    struct assist_extras* ax = (struct assist_extras*)sim->extras;
    if (!ax || !ax->ephem) { fail(sim, "assist_additional_forces: ASSIST is not attached."); return; }
    const int n_real = (int)sim->N - sim->N_var;
    if (n_real <= 0) return;
    /* layout without touching the cached integration batch */
    AbHostBatch hb;
    hb.gb = NULL; hb.n_real = n_real;
    hb.nv.assign(n_real, 0);
    for (unsigned int v = 0; v < sim->N_var_config; v++) {
        const int tp = sim->var_config[v].testparticle;
        if (tp < 0 || tp >= n_real) { fail(sim, "Unsupported variational configuration."); return; }
        hb.nv[tp]++;
    }
    int nvmax = 0;
    for (int i = 0; i < n_real; i++) if (hb.nv[i] > nvmax) nvmax = hb.nv[i];
    if (nvmax > ASSIST_GPU_MAX_NVAR) { fail(sim, "More variational particles per test particle than the GPU kernels are built for."); return; }
    hb.K = 1 + nvmax;
    hb.var_pidx.assign((size_t)n_real * nvmax, -1);
    hb.var_cfg.assign((size_t)n_real * nvmax, -1);
    std::vector<int> fill(n_real, 0);
    for (unsigned int v = 0; v < sim->N_var_config; v++) {
        const int tp = sim->var_config[v].testparticle;
        const int slot = fill[tp]++;
        hb.var_pidx[(size_t)tp * nvmax + slot] = sim->var_config[v].index;
        hb.var_cfg[(size_t)tp * nvmax + slot] = (int)v;
    }
    std::vector<double> st, prm, acc((size_t)n_real * hb.K * 3);
    gather_state(sim, &hb, st);
    const bool has_prm = gather_params(ax, &hb, prm);
    struct assist_gpu_options opt;
    ab_host_fill_options(sim, ax, &opt);
    const double t = sim->t;
    int rc = assist_gpu_eval_forces(ax->ephem, &opt, n_real, nvmax, &t, 0, st.data(), has_prm ? prm.data() : NULL, acc.data(), NULL);
    if (rc > 0) { reb_simulation_error(sim, assist_error_messages[rc]); sim->status = REB_STATUS_GENERIC_ERROR; return; }
    if (rc < 0) { fail(sim, assist_gpu_last_error()); return; }
    for (int i = 0; i < n_real; i++) {
        for (int j = 0; j < hb.K; j++) {
            int pidx = (j == 0) ? i : hb.var_pidx[(size_t)i * nvmax + (j - 1)];
            if (pidx < 0) continue;
            const double* a = &acc[((size_t)i * hb.K + j) * 3];
            sim->particles[pidx].ax += a[0]; sim->particles[pidx].ay += a[1]; sim->particles[pidx].az += a[2];
        }
    }
}

/* Dense output of the attached simulation inside its last completed step: fills dest[] (N particles). */
extern "C" int ab_host_interpolate(struct reb_simulation* r, double h, struct reb_particle* dest) {
    AbHostBatch* hb = (AbHostBatch*)r->b200_batch;
    if (!hb) return -1;
    std::vector<double> st((size_t)hb->n_real * hb->K * 6);
    if (assist_gpu_batch_interpolate(hb->gb, h, st.data())) return fail(r, assist_gpu_last_error());
    scatter_state(r, hb, st, NULL, dest);
    return 0;
}
