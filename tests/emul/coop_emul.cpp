/*
 * coop_emul.cpp -- TEST INFRASTRUCTURE: runs the source of pp_coop_kernel (assist_b200/csrc/coop_roles.cuh,
 * strict variant) on the host, one OS thread per warp and one pthread barrier per CTA, so that its control flow and
 * its arithmetic can be checked against the oracle without a GPU (tests/test_cpu_coop_emul.py).  Nothing in the
 * product links or loads this file.
 *
 * The CUDA keywords are defined away, the per-lane blocks of the kernel become loops over 32 lanes (AB_HOST_EMUL),
 * __syncthreads becomes the barrier.  IEEE double arithmetic, sqrt, division and fma() are the same on both sides
 * (the build uses -ffp-contract=off), so the strict build's bits are reproduced; only pow() (Marsden term) differs.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define AB_HOST_EMUL 1
#define AB_STRICT 1
#define AB_NS ab_emul

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __grid_constant__

struct double2 { double x, y; };
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence() { __sync_synchronize(); }

struct AbcEmulShared {
    pthread_barrier_t bar;
    int or_flag[2];
};
namespace ab_emul { struct AbcSmem; }
struct AbcEmulCtx {
    const ab_emul::AbcSmem* sm;
    int block;      /* index of the group (ABC_BLOCK) */
    int warp;       /* warp index inside the CTA */
    int phase;
    AbcEmulShared* sh;
};
static inline void abc_emul_sync(AbcEmulCtx* ctx) { pthread_barrier_wait(&ctx->sh->bar); }
static inline bool abc_emul_sync_or(AbcEmulCtx* ctx, int p) {
    int* f = &ctx->sh->or_flag[ctx->phase & 1];
    if (p) __atomic_fetch_or(f, 1, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&ctx->sh->bar);
    const int v = __atomic_load_n(f, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&ctx->sh->bar);
    if (ctx->warp == 0) __atomic_store_n(f, 0, __ATOMIC_SEQ_CST);
    ctx->phase++;
    return v != 0;
}

#include "coop_roles.cuh"

using namespace ab_emul;

struct WarpJob {
    AbcEmulCtx ctx;
    const AbEphem* E;
    const AbForceOpts* F;
    const AbcArgs* A;
    AbcSmem sm;
};

static void* warp_main(void* arg) {
    WarpJob* j = (WarpJob*)arg;
    const int w = j->ctx.warp % ABC_GWARPS;
    if (w < 3) abc_comp_main(&j->ctx, *j->E, *j->F, *j->A, j->sm, w);
    else if (w == ABC_CTRL_WARP) abc_control_main(&j->ctx, *j->E, *j->F, *j->A, j->sm);
    else abc_worker_main(&j->ctx, *j->E, *j->F, *j->A, j->sm, w);
    return NULL;
}

extern "C" size_t abc_emul_sizeof(int what) {
    switch (what) {
        case 0: return sizeof(AbEphem);
        case 1: return sizeof(AbForceOpts);
        case 2: return sizeof(AbBatch);
        case 3: return sizeof(AbcPlan);
        default: return 0;
    }
}

/* A population on the host with its own (arbitrary) layout: the kernel only sees the AbBatch pointers. */
struct EmulBatch {
    int n;
    std::vector<double> mem, wmem;
    std::vector<unsigned long long> cnt, wcnt;
    std::vector<int> ints, wints;
    AbBatch d, w;
    int w_slots;
};

static void layout(AbBatch& d, std::vector<double>& mem, std::vector<unsigned long long>& cnt, std::vector<int>& ints, size_t n) {
    const size_t C = 3, per = C * n;
    mem.assign(12 * per + 42 * per + 4 * n, 0.0);
    cnt.assign(4 * n, 0ULL);
    ints.assign(2 * n, 0);
    memset(&d, 0, sizeof(d));
    double* p = mem.data();
    d.n = (int)n; d.K = 1; d.C = 3; d.mode = 1;
    d.pos = p; p += per; d.vel = p; p += per; d.acc = p; p += per;
    d.x0 = p; p += per; d.v0 = p; p += per; d.a0 = p; p += per; d.csx = p; p += per; d.csv = p; p += per;
    d.ls_pos = p; p += per; d.ls_vel = p; p += per; d.ls_acc = p; p += per; d.prm = p; p += per;
    d.b = p; p += 7 * per; d.g = p; p += 7 * per; d.e = p; p += 7 * per;
    d.csb = p; p += 7 * per; d.br = p; p += 7 * per; d.er = p; p += 7 * per;
    d.t = p; p += n; d.dt = p; p += n; d.dt_last = p; p += n; d.last_full_dt = p; p += n;
    d.steps = cnt.data(); d.rejected = cnt.data() + n; d.iters = cnt.data() + 2 * n; d.evals = cnt.data() + 3 * n;
    d.nv = ints.data(); d.status = ints.data() + n;
}

extern "C" void* abc_emul_create(int n, int max_blocks) {
    EmulBatch* b = new EmulBatch();
    b->n = n;
    layout(b->d, b->mem, b->cnt, b->ints, (size_t)n);
    b->w_slots = max_blocks * ABC_GROUPS * ABC_SLOTS;
    layout(b->w, b->wmem, b->wcnt, b->wints, (size_t)b->w_slots);
    return b;
}
extern "C" void abc_emul_free(void* h) { delete (EmulBatch*)h; }

/* state[n][6], params[n][3] or NULL: fresh IAS15 history, as assist_gpu_batch_set_state */
extern "C" void abc_emul_set_state(void* h, double t0, double dt0, const double* state, const double* params, double epsilon, double min_dt) {
    EmulBatch* b = (EmulBatch*)h;
    const size_t n = (size_t)b->n;
    std::fill(b->mem.begin(), b->mem.end(), 0.0);
    std::fill(b->cnt.begin(), b->cnt.end(), 0ULL);
    for (size_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            b->d.pos[c * n + i] = state[i * 6 + c];
            b->d.vel[c * n + i] = state[i * 6 + 3 + c];
            if (params) b->d.prm[c * n + i] = params[i * 3 + c];
        }
        b->d.t[i] = t0; b->d.dt[i] = dt0; b->d.nv[i] = 0; b->d.status[i] = -3;
    }
    b->d.has_params = params ? 1 : 0;
    b->d.epsilon = epsilon; b->d.min_dt = min_dt;
}

extern "C" void abc_emul_get_state(void* h, double* state, double* t, double* dt, double* dt_last, int* status, unsigned long long* counters) {
    EmulBatch* b = (EmulBatch*)h;
    const size_t n = (size_t)b->n;
    for (size_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) { state[i * 6 + c] = b->d.pos[c * n + i]; state[i * 6 + 3 + c] = b->d.vel[c * n + i]; }
        t[i] = b->d.t[i]; dt[i] = b->d.dt[i]; dt_last[i] = b->d.dt_last[i]; status[i] = b->d.status[i];
        counters[i * 4 + 0] = b->d.steps[i]; counters[i * 4 + 1] = b->d.rejected[i];
        counters[i * 4 + 2] = b->d.iters[i]; counters[i * 4 + 3] = b->d.evals[i];
    }
}

/* times == NULL: reb_simulation_integrate(tmax) for every system; else assist_integrate_or_interpolate over the
 * epochs into out[n_times][n][6]. */
extern "C" int abc_emul_run(void* h, const void* E_, const void* F_, const void* plan_,
                            double tmax, int exact, const double* times, int n_times, double* out, int n_blocks) {
    EmulBatch* hb = (EmulBatch*)h;
    const AbEphem& E = *(const AbEphem*)E_;
    const AbForceOpts& F = *(const AbForceOpts*)F_;
    memcpy(c_h, AB_H, sizeof(AB_H));
    memcpy(c_rr, AB_RR, sizeof(AB_RR));
    for (int k = 0; k < 28; k++) c_rri[k] = 1.0 / AB_RR[k];
    memcpy(c_c, AB_C, sizeof(AB_C));
    memcpy(c_d, AB_D, sizeof(AB_D));
    c_abcE = E; c_abcF = F; c_abcP = *(const AbcPlan*)plan_;
    for (int m = 0; m < E.n_ast && m < AB_MAX_AST; m++) c_abc_ast[m] = E.a_tgt[m];
    c_abc_regular = abc_series_table(E, E.a_tgt, c_abc_tg);

    AbcArgs A;
    hb->w.epsilon = hb->d.epsilon; hb->w.min_dt = hb->d.min_dt; hb->w.has_params = hb->d.has_params;
    A.Bt = hb->d;
    A.W = hb->w;
    A.tmax = tmax; A.exact_finish_time = exact;
    unsigned long long queue = 0;
    A.queue_head = &queue;
    std::vector<int> done((size_t)A.Bt.n, 0), epoch((size_t)A.Bt.n, 0);
    A.SL.origin = 0.0; A.SL.wlen = 1.0; A.SL.n_win = 1; A.SL.done = done.data(); A.SL.epoch = epoch.data(); A.SL.order = nullptr; A.SL.attempt_budget = 0;
    A.times = times; A.n_times = n_times; A.out = out;
    A.plan = *(const AbcPlan*)plan_;
    A.timing = nullptr;
    std::vector<double> gtab((size_t)n_blocks * ABC_GROUPS * ABC_GT_DOUBLES, 0.0);
    A.gtab = gtab.data();
    if (A.W.n < n_blocks * ABC_GROUPS * ABC_SLOTS) return -1;

    std::vector<AbcEmulShared> shared((size_t)n_blocks);
    std::vector<std::vector<double>> smem((size_t)n_blocks * ABC_GROUPS);
    std::vector<WarpJob> jobs((size_t)n_blocks * ABC_WARPS);
    std::vector<pthread_t> th((size_t)n_blocks * ABC_WARPS);
    for (int b = 0; b < n_blocks; b++) {
        pthread_barrier_init(&shared[b].bar, NULL, ABC_WARPS);
        shared[b].or_flag[0] = shared[b].or_flag[1] = 0;
        for (int g = 0; g < ABC_GROUPS; g++) smem[(size_t)b * ABC_GROUPS + g].assign(ABC_SMEM_GROUP_BYTES / 8 + 1, 0.0);
        for (int w = 0; w < ABC_WARPS; w++) {
            WarpJob& j = jobs[(size_t)b * ABC_WARPS + w];
            const size_t unit = (size_t)b * ABC_GROUPS + w / ABC_GWARPS;
            j.ctx.block = (int)unit; j.ctx.warp = w; j.ctx.phase = 0; j.ctx.sh = &shared[b];
            j.E = &E; j.F = &F; j.A = &A;
            j.sm.d = smem[unit].data();
            j.sm.i = (int*)(smem[unit].data() + ABC_SM_DOUBLES);
            j.ctx.sm = &j.sm;
        }
    }
    for (size_t k = 0; k < jobs.size(); k++) pthread_create(&th[k], NULL, warp_main, &jobs[k]);
    for (size_t k = 0; k < jobs.size(); k++) pthread_join(th[k], NULL);
    for (int b = 0; b < n_blocks; b++) pthread_barrier_destroy(&shared[b].bar);
    return 0;
}
