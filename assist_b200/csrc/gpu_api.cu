/*
 * gpu_api.cu -- the thin C-ABI of include/assist_gpu.h: device memory, marshalling
 * and kernel launches.  No physics here; the kernels are in kernels.cu.
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <vector>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "assist_gpu.h"
#include "assist_ephem_files.h"
#include "device_types.h"
#include "launchers.h"
#include "host_internal.h"

/* ------------------------------------------------------------------------ */
/* errors / device                                                          */
/* ------------------------------------------------------------------------ */

static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return set_err(ASSIST_GPU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e));  \
    } while (0)

extern "C" const char* assist_gpu_last_error(void) { return g_err; }

/* every kernel this library launches, process-wide (assist_gpu_kernel_launches) */
static std::atomic<unsigned long long> g_kernel_launches{0};
#define AB_COUNT(k) g_kernel_launches.fetch_add((unsigned long long)(k), std::memory_order_relaxed)

extern "C" unsigned long long assist_gpu_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

extern "C" int assist_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static bool g_const_uploaded[ASSIST_B200_MAX_DEVICES] = {false};

/* Returns the current device after making sure a device exists and constants are loaded. */
static int ensure_device(int* dev_out) {
    if (assist_gpu_device_count() < 1)
        return set_err(ASSIST_GPU_ERR_NO_DEVICE,
                       "no CUDA device available: assist-b200 has no CPU compute path");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= ASSIST_B200_MAX_DEVICES) return set_err(ASSIST_GPU_ERR_ARG, "device index %d out of range", dev);
    if (!g_const_uploaded[dev]) {
        CU(ab_upload_constants_strict_tu0()); CU(ab_upload_constants_strict_tu1());
        CU(ab_upload_constants_strict_tu2()); CU(ab_upload_constants_strict_tu3());
        CU(ab_upload_constants_fast_tu0()); CU(ab_upload_constants_fast_tu1());
        CU(ab_upload_constants_fast_tu2()); CU(ab_upload_constants_fast_tu3());
        CU(ab_upload_constants_strict_tu4()); CU(ab_upload_constants_fast_tu4());
        g_const_uploaded[dev] = true;
    }
    *dev_out = dev;
    return 0;
}

extern "C" int assist_gpu_set_device(int device) {
    if (assist_gpu_device_count() < 1)
        return set_err(ASSIST_GPU_ERR_NO_DEVICE, "no CUDA device available: assist-b200 has no CPU compute path");
    CU(cudaSetDevice(device));
    return 0;
}

extern "C" int assist_gpu_selftest_fp(unsigned long long seed, long long n_pairs, unsigned long long mismatches[2]) {
    int dev = 0;
    int rc = ensure_device(&dev);
    if (rc) return rc;
    if (!mismatches || n_pairs < 1) return set_err(ASSIST_GPU_ERR_ARG, "bad argument");
    unsigned long long* d_bad = nullptr;
    CU(cudaMalloc((void**)&d_bad, 2 * sizeof(unsigned long long)));
    CU(cudaMemset(d_bad, 0, 2 * sizeof(unsigned long long)));
    const int blocks = 1184, iters = (int)((n_pairs + (long long)blocks * 256 - 1) / ((long long)blocks * 256));
    cudaError_t e = ab_launch_fp_selftest_strict(seed, blocks, iters, d_bad, 0);
    if (e == cudaSuccess) e = cudaMemcpy(mismatches, d_bad, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "fp self-test failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" void assist_gpu_default_options(struct assist_gpu_options* opt) {
    /* defaults of assist_init, reference src/assist.c:415-438, and REBOUND's IAS15 defaults */
    opt->forces = ASSIST_FORCE_SUN | ASSIST_FORCE_PLANETS | ASSIST_FORCE_ASTEROIDS | ASSIST_FORCE_NON_GRAVITATIONAL |
                  ASSIST_FORCE_EARTH_HARMONICS | ASSIST_FORCE_SUN_HARMONICS | ASSIST_FORCE_GR_EIH;
    opt->gr_eih_sources = 1;
    opt->geocentric = 0;
    opt->math = ASSIST_GPU_MATH_STRICT;
    opt->alpha = 1.0; opt->nk = 0.0; opt->nm = 2.0; opt->nn = 5.093; opt->r0 = 1.0;
    opt->epsilon = 1e-9;
    opt->min_dt = 0.0;
}

/* ------------------------------------------------------------------------ */
/* ephemeris upload                                                         */
/* ------------------------------------------------------------------------ */

static int upload_image(void** slot, const void* host, size_t len) {
    if (*slot) return 0;
    void* d = nullptr;
    CU(cudaMalloc(&d, len));
    CU(cudaMemcpy(d, host, len, cudaMemcpyHostToDevice));
    *slot = d;
    return 0;
}

/* Device copy of an SPK kernel: only the type-2 records, segment after segment, each record as
 * [MID, RADIUS, (x y z) of term 0, (x y z) of term 1, ...], the term list padded with ZERO coefficients to a multiple
 * of four terms: the coefficients of a record start on a 16-byte boundary, the three components of a term are
 * adjacent (two terms are three 16-byte loads) and the staged fill of pp_coop_kernel takes four terms per trip without
 * a test for the end of the list (a zero coefficient adds nothing to a sum).  Values are copied, never recomputed.
 * off[m * AB_MAXSEG + s] = first word of segment s of target m in that copy. */
static inline int packed_record_words(int R) {       /* R = 2 + 3 P words in the file */
    const int P = (R - 2) / 3;
    return 2 + 3 * ((P + 3) & ~3);
}
static int spk_layout(const struct spk_s* file, std::vector<long long>& off, size_t* total_words) {
    const double* img = (const double*)file->map;
    const size_t words = file->len / sizeof(double);
    off.assign((size_t)file->num * AB_MAXSEG, 0);
    long long cur = 0;
    for (int m = 0; m < file->num; m++) {
        const struct spk_target* t = &file->targets[m];
        const int nseg = t->ind + 1;
        if (nseg > AB_MAXSEG)
            return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "SPK target %d has %d segments (max %d)", t->code, nseg, AB_MAXSEG);
        for (int s = 0; s < nseg; s++) {
            if (t->two[s] < 4 || (size_t)t->two[s] > words || t->one[s] < 1)
                return set_err(ASSIST_GPU_ERR_ARG, "SPK target %d: segment addresses out of range", t->code);
            const double* val = img + t->two[s] - 1;
            const int R = (int)val[-1], nrec = (int)val[0];
            if (R < 8 || R > 98 || nrec < 1 || (size_t)(t->one[s] - 1) + (size_t)nrec * R > words)
                return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "SPK target %d: not a type-2 Chebyshev segment", t->code);
            off[(size_t)m * AB_MAXSEG + s] = cur;
            cur += (long long)nrec * packed_record_words(R);
        }
    }
    *total_words = (size_t)cur + 2;      /* the last 16-byte pair of an odd-length record may be read past its end */
    return 0;
}

static void pack_spk(const struct spk_s* file, const std::vector<long long>& off, std::vector<double>& buf);

static int upload_packed_spk(void** slot, const struct spk_s* file, const std::vector<long long>& off, size_t total_words) {
    if (*slot) return 0;
    std::vector<double> buf(total_words, 0.0);
    pack_spk(file, off, buf);
    void* d = nullptr;
    CU(cudaMalloc(&d, sizeof(double) * total_words));
    CU(cudaMemcpy(d, buf.data(), sizeof(double) * total_words, cudaMemcpyHostToDevice));
    *slot = d;
    return 0;
}

/* Host-only view of the packed copy (no device needed): what upload_packed_spk sends.  `*out` is malloc'ed
 * (free() it), seg_off receives AB_MAXSEG entries per target.  Used by tests/test_cpu_host.py. */
extern "C" int assist_gpu_spk_pack_host(const struct spk_s* file, double** out, size_t* words, long long* seg_off, int seg_off_len) {
    if (!file || !out || !words) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    std::vector<long long> off;
    size_t total = 0;
    int rc = spk_layout(file, off, &total);
    if (rc) return rc;
    std::vector<double> buf(total, 0.0);
    pack_spk(file, off, buf);
    double* o = (double*)malloc(sizeof(double) * total);
    if (!o) return set_err(ASSIST_GPU_ERR_ARG, "out of memory");
    memcpy(o, buf.data(), sizeof(double) * total);
    *out = o; *words = total;
    if (seg_off) for (int q = 0; q < seg_off_len && q < (int)off.size(); q++) seg_off[q] = off[q];
    return 0;
}

static void pack_spk(const struct spk_s* file, const std::vector<long long>& off, std::vector<double>& buf) {
    const double* img = (const double*)file->map;
    for (int m = 0; m < file->num; m++) {
        const struct spk_target* t = &file->targets[m];
        for (int s = 0; s <= t->ind; s++) {
            const double* val = img + t->two[s] - 1;
            const int R = (int)val[-1], nrec = (int)val[0], P = (R - 2) / 3, Rp = packed_record_words(R);
            for (int b = 0; b < nrec; b++) {
                const double* src = img + (t->one[s] - 1) + (size_t)b * R;
                double* dst = buf.data() + off[(size_t)m * AB_MAXSEG + s] + (size_t)b * Rp;
                dst[0] = 2451545.0 + src[0] / 86400.0;      /* _jul(MID), reference src/spk.c:60: one IEEE division and one addition, as there */
                dst[1] = src[1];
                for (int p = 0; p < P; p++)
                    for (int c = 0; c < 3; c++) dst[2 + 3 * p + c] = src[2 + c * P + p];
            }
        }
    }
}

static int fill_target(AbSpkTarget* d, const struct spk_target* t, const struct spk_s* file, const long long* seg_off) {
    memset(d, 0, sizeof(*d));
    d->beg = t->beg; d->end = t->end; d->res = t->res; d->res_rd = 1.0 / t->res; d->mass = t->mass;
    d->code = t->code; d->cen = t->cen; d->nseg = t->ind + 1;
    if (d->nseg > AB_MAXSEG)
        return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "SPK target %d has %d segments (max %d)", t->code, d->nseg, AB_MAXSEG);
    const double* img = (const double*)file->map;
    const size_t words = file->len / sizeof(double);
    for (int s = 0; s < d->nseg; s++) {
        /* segment trailer [INIT, INTLEN, RSIZE, N]; `two` addresses its last word (reference src/spk.c:503-510) */
        if (t->two[s] < 4 || (size_t)t->two[s] > words || t->one[s] < 1)
            return set_err(ASSIST_GPU_ERR_ARG, "SPK target %d: segment addresses out of range", t->code);
        const double* val = img + t->two[s] - 1;
        AbSpkSeg* sg = &d->seg[s];
        sg->one = t->one[s];
        sg->R = (int)val[-1];
        sg->P = (sg->R - 2) / 3;
        sg->nrec = (int)val[0];
        if (sg->P < 2 || sg->P >= 32 || sg->nrec < 1)
            return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "SPK target %d: not a type-2 Chebyshev segment", t->code);
        sg->jul_init = 2451545.0 + val[-3] / 86400.0;
        sg->intlen_d = val[-2] / 86400.0;
        sg->intlen_rd = 1.0 / sg->intlen_d;
        const double* rec0 = img + (t->one[s] - 1);
        const double radius = rec0[1];
        sg->radius_d = radius / 86400.0;
        sg->radius_rd = 1.0 / sg->radius_d;
        sg->radius_inv = 1.0 / radius;
        sg->uniform = 1;
        for (int b = 1; b < sg->nrec; b++)
            if (rec0[(size_t)b * sg->R + 1] != radius) { sg->uniform = 0; break; }
        /* from here on the descriptor addresses the packed device copy, not the file */
        sg->one = (int)(seg_off[s] + 1);
        sg->R = packed_record_words(sg->R);
    }
    return 0;
}

/* Per-file cache (struct spk_s::b200_host_desc): where each segment lies in the packed copy and the segment
 * descriptors.  Building them walks every record of the file once (the uniform-RADIUS check), so it is done
 * once per file, not once per call. */
struct AbSpkDesc {
    std::vector<long long> off;
    size_t words = 0;
    std::vector<AbSpkTarget> tg;                                   /* masses are refreshed by the caller */
    std::vector<double> host_copy;                                 /* packed copy on the host (emulation hook only) */
    std::vector<double> dev_mass[ASSIST_B200_MAX_DEVICES];         /* masses last sent with the device descriptors */
};

extern "C" void ab_spk_desc_free(void* desc) { delete (AbSpkDesc*)desc; }

static int spk_desc(struct spk_s* f, AbSpkDesc** out) {
    if (f->b200_host_desc) { *out = (AbSpkDesc*)f->b200_host_desc; return 0; }
    AbSpkDesc* d = new AbSpkDesc();
    int rc = spk_layout(f, d->off, &d->words);
    if (!rc) {
        d->tg.resize((size_t)f->num);
        for (int m = 0; m < f->num && !rc; m++) rc = fill_target(&d->tg[m], &f->targets[m], f, &d->off[(size_t)m * AB_MAXSEG]);
    }
    if (rc) { delete d; return rc; }
    f->b200_host_desc = d;
    *out = d;
    return 0;
}

/* Build the kernel-parameter view of an ephemeris: on the current device, or (host_mode) with HOST pointers
 * to the same packed tables -- what the CPU emulation of the kernels in tests/emul reads. */
static int build_ephem_impl(const struct assist_ephem* e, AbEphem* E, bool host_mode) {
    int dev = 0;
    int rc;
    if (!host_mode) { rc = ensure_device(&dev); if (rc) return rc; }
    if (e == NULL) return set_err(ASSIST_GPU_ERR_ARG, "ephem is NULL");
    memset(E, 0, sizeof(*E));
    E->jd_ref = e->jd_ref;
    E->planets_source = e->planets_source;
    E->AU = e->AU; E->EMRAT = e->EMRAT; E->J2E = e->J2E; E->J3E = e->J3E; E->J4E = e->J4E; E->J2SUN = e->J2SUN;
    E->Re_eq = e->Re_eq; E->Rs_eq = e->Rs_eq; E->c_squared = e->c_squared; E->over_c_squared = e->over_c_squared;
    for (int k = 0; k < AB_NPLANETS; k++) E->p_index[k] = -1;
    E->emb_index = -1;
    if (e->ascii_planets) {
        struct ascii_s* a = e->ascii_planets;
        if (host_mode) {
            E->ascii_img = (const double*)a->map;
        } else {
            if ((rc = upload_image(&a->b200_dev_image[dev], a->map, a->len))) return rc;
            E->ascii_img = (const double*)a->b200_dev_image[dev];
        }
        E->a_beg = a->beg; E->a_end = a->end; E->a_inc = a->inc; E->a_cau = a->cau; E->a_cem = a->cem;
        E->a_inc_rd = 1.0 / a->inc;
        E->a_f_earth = -1.0 / (1.0 + a->cem);
        E->a_f_moon = a->cem / (1.0 + a->cem);
        E->u_d[0] = a->cau; E->u_d[1] = a->cau / 86400.; E->u_d[2] = a->cau / (86400. * 86400.);
        E->a_rec_words = (long long)(a->rec / sizeof(double));
        E->a_nrec = (long long)(a->len / a->rec) - 2;
        for (int p = 0; p < 15; p++) {
            E->a_off[p] = a->off[p]; E->a_ncf[p] = a->ncf[p]; E->a_niv[p] = a->niv[p];
            E->a_c[p] = (double)(a->niv[p] * 2) / a->inc / 86400.0;
        }
        for (int k = 0; k < AB_NPLANETS; k++) E->a_mass[k] = a->mass[k];
    } else if (e->spk_planets) {
        struct spk_s* pl = e->spk_planets;
        if (pl->num > AB_MAX_PTGT)
            return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "planet kernel has %d targets (max %d)", pl->num, AB_MAX_PTGT);
        AbSpkDesc* pd = nullptr;
        if ((rc = spk_desc(pl, &pd))) return rc;
        if (host_mode) {
            if (pd->host_copy.empty()) { pd->host_copy.assign(pd->words, 0.0); pack_spk(pl, pd->off, pd->host_copy); }
            E->spkp_img = pd->host_copy.data();
        } else {
            if ((rc = upload_packed_spk(&pl->b200_dev_image[dev], pl, pd->off, pd->words))) return rc;
            E->spkp_img = (const double*)pl->b200_dev_image[dev];
        }
        {
            const double au = e->AU, seconds_per_day = 86400.;
            E->u_d[0] = au; E->u_d[1] = au / seconds_per_day; E->u_d[2] = au / (seconds_per_day * seconds_per_day);
        }
        E->n_ptgt = pl->num;
        for (int m = 0; m < pl->num; m++) {
            E->p_tgt[m] = pd->tg[m];
            E->p_tgt[m].mass = pl->targets[m].mass;        /* masses can be joined after the first build */
        }
        static const int naif_by_assist[AB_NPLANETS] = {10, 1, 2, 399, 301, 4, 5, 6, 7, 8, 9};
        for (int k = 0; k < AB_NPLANETS; k++) {
            /* precomputed index when it is consistent, else a search by NAIF code (reference src/spk.c:646-659) */
            int idx = e->spk_target_index[k];
            if (!(idx >= 0 && idx < pl->num && pl->targets[idx].code == naif_by_assist[k])) {
                idx = -1;
                for (int m = 0; m < pl->num; m++) if (pl->targets[m].code == naif_by_assist[k]) { idx = m; break; }
            }
            E->p_index[k] = idx;
        }
        int emb = e->spk_emb_index;
        if (!(emb >= 0 && emb < pl->num && pl->targets[emb].code == 3)) {
            emb = -1;
            for (int m = 0; m < pl->num; m++) if (pl->targets[m].code == 3) { emb = m; break; }
        }
        E->emb_index = emb;
    } else {
        return set_err(ASSIST_ERROR_EPHEM_FILE, "ephemeris has no planets provider");
    }
    for (int q = 0; q < 3; q++) E->u_rd[q] = 1.0 / E->u_d[q];
    if (e->spk_asteroids) {
        struct spk_s* sb = e->spk_asteroids;
        if (sb->num > AB_MAX_AST_ALL)
            return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "small-body kernel has %d targets (max %d in this build)", sb->num, AB_MAX_AST_ALL);
        AbSpkDesc* ad = nullptr;
        if ((rc = spk_desc(sb, &ad))) return rc;
        for (int m = 0; m < sb->num; m++) ad->tg[m].mass = sb->targets[m].mass;
        /* the first AB_MAX_AST targets (all of sb441-n16) live in the per-time body tables; the others (sb441-n373)
         * act through the direct term only and are evaluated inside it */
        E->n_ast = sb->num < AB_MAX_AST ? sb->num : AB_MAX_AST;
        E->n_ast_x = sb->num - E->n_ast;
        if (host_mode) {
            if (ad->host_copy.empty()) { ad->host_copy.assign(ad->words, 0.0); pack_spk(sb, ad->off, ad->host_copy); }
            E->spka_img = ad->host_copy.data();
            E->a_tgt = ad->tg.data();
        } else {
            if ((rc = upload_packed_spk(&sb->b200_dev_image[dev], sb, ad->off, ad->words))) return rc;
            E->spka_img = (const double*)sb->b200_dev_image[dev];
            /* the descriptors carry the (joinable) masses: sent again only when a mass has changed */
            std::vector<double>& sent = ad->dev_mass[dev];
            bool same = sb->b200_dev_targets[dev] != nullptr && sent.size() == (size_t)sb->num;
            for (int m = 0; same && m < sb->num; m++) same = (sent[m] == sb->targets[m].mass);
            if (!same) {
                if (!sb->b200_dev_targets[dev]) CU(cudaMalloc(&sb->b200_dev_targets[dev], sizeof(AbSpkTarget) * (sb->num > AB_MAX_AST ? sb->num : AB_MAX_AST)));
                CU(cudaMemcpy(sb->b200_dev_targets[dev], ad->tg.data(), sizeof(AbSpkTarget) * sb->num, cudaMemcpyHostToDevice));
                sent.resize((size_t)sb->num);
                for (int m = 0; m < sb->num; m++) sent[m] = sb->targets[m].mass;
            }
            E->a_tgt = (const AbSpkTarget*)sb->b200_dev_targets[dev];
        }
        for (int m = 0; m < E->n_ast; m++) E->gm[AB_NPLANETS + m] = sb->targets[m].mass;
    }
    /* common coverage window (the checks of ab_fill_nodes / abc_coverage, folded) */
    E->cov_lo = -1e300; E->cov_hi = 1e300; E->cov_simple = 1;
    if (e->ascii_planets) {
        E->cov_lo = E->a_beg; E->cov_hi = E->a_end;
    } else {
        for (int k = 0; k < AB_NPLANETS; k++) {
            const int idx = E->p_index[k];
            if (idx < 0) { if (k != 3) E->cov_simple = 0; continue; }
            if (E->p_tgt[idx].beg > E->cov_lo) E->cov_lo = E->p_tgt[idx].beg;
            if (E->p_tgt[idx].end < E->cov_hi) E->cov_hi = E->p_tgt[idx].end;
        }
    }
    if (e->spk_asteroids) {
        for (int m = 0; m < e->spk_asteroids->num; m++) {
            const struct spk_target* t = &e->spk_asteroids->targets[m];
            if (t->beg > E->cov_lo) E->cov_lo = t->beg;
            if (t->end < E->cov_hi) E->cov_hi = t->end;
        }
    }
    for (int k = 0; k < AB_NPLANETS; k++) {
        if (e->ascii_planets) E->gm[k] = e->ascii_planets->mass[k];
        else E->gm[k] = (E->p_index[k] >= 0) ? E->p_tgt[E->p_index[k]].mass : 0.0;   /* Earth-from-EMB fallback reports GM = 0 */
    }
    return 0;
}

static int build_ephem(const struct assist_ephem* e, AbEphem* E) { return build_ephem_impl(e, E, false); }

/* Host-only test hook (no device needed): the kernel-parameter view with host pointers, plus the force options
 * as the launches build them.  tests/emul runs the kernels' source on the CPU with these. */
extern "C" int ab_gpu_build_ephem_host(const struct assist_ephem* e, void* E_out, size_t E_bytes) {
    if (E_bytes != sizeof(AbEphem)) return set_err(ASSIST_GPU_ERR_ARG, "AbEphem is %zu bytes, caller expects %zu", sizeof(AbEphem), E_bytes);
    return build_ephem_impl(e, (AbEphem*)E_out, true);
}

static void build_force_opts(const struct assist_gpu_options* o, int has_params, AbForceOpts* F) {
    memset(F, 0, sizeof(*F));
    F->forces = o->forces;
    F->gr_eih_sources = o->gr_eih_sources < 0 ? 0 : (o->gr_eih_sources > AB_NPLANETS ? AB_NPLANETS : o->gr_eih_sources);
    F->geocentric = o->geocentric;
    F->has_params = has_params;
    F->alpha = o->alpha; F->nk = o->nk; F->nm = o->nm; F->nn = o->nn; F->r0 = o->r0;
    /* Earth pole reset to the J2000 equator, solar pole RA 286.13 Dec 63.87
     * (reference src/forces.c:484-490, 670-676); same libm calls as the reference. */
    const double RAe = 0.0 * M_PI / 180., Dece = 90.0 * M_PI / 180.;
    F->e_cosa = cos(RAe); F->e_sina = sin(RAe); F->e_cosd = cos(Dece); F->e_sind = sin(Dece);
    const double RAs = 286.13 * M_PI / 180., Decs = 63.87 * M_PI / 180.;
    F->s_cosa = cos(RAs); F->s_sina = sin(RAs); F->s_cosd = cos(Decs); F->s_sind = sin(Decs);
}

extern "C" int assist_gpu_ephem_upload(const struct assist_ephem* ephem) {
    AbEphem E;
    return build_ephem(ephem, &E);
}

extern "C" int assist_gpu_ephem_nbodies(const struct assist_ephem* ephem) {
    if (!ephem) return 0;
    return AB_NPLANETS + (ephem->spk_asteroids ? ephem->spk_asteroids->num : 0);
}

/* Device scratch of the synchronous single-call entry points (assist_gpu_ephem_eval, assist_gpu_eval_forces: what
 * assist_get_particle, reb_simulation_update_acceleration and the REBOUND-driven plug-in path go through): grow-only
 * buffers per host thread and device instead of four cudaMalloc / cudaFree pairs per call. */
struct AbScratch { void* p[5]; size_t cap[5]; };
static thread_local AbScratch g_scratch[ASSIST_B200_MAX_DEVICES];
static int scratch_get(int slot, size_t bytes, void** out) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= ASSIST_B200_MAX_DEVICES) return set_err(ASSIST_GPU_ERR_ARG, "device index %d out of range", dev);
    AbScratch& s = g_scratch[dev];
    if (s.cap[slot] < bytes) {
        if (s.p[slot]) cudaFree(s.p[slot]);
        s.p[slot] = nullptr; s.cap[slot] = 0;
        const size_t want = bytes < 4096 ? 4096 : bytes + bytes / 4;
        CU(cudaMalloc(&s.p[slot], want));
        s.cap[slot] = want;
    }
    *out = s.p[slot];
    return 0;
}
#define SCRATCH(slot, bytes, ptr) do { int rc_ = scratch_get((slot), (bytes), (void**)&(ptr)); if (rc_) return rc_; } while (0)

extern "C" int assist_gpu_ephem_eval(const struct assist_ephem* ephem, int math, const double* t, int n_t,
                                     double* out, int* status) {
    AbEphem E;
    int rc = build_ephem(ephem, &E);
    if (rc) return rc;
    if (n_t <= 0) return 0;
    const int nb = AB_NPLANETS + E.n_ast + E.n_ast_x;
    double *d_t = nullptr, *d_out = nullptr;
    int* d_st = nullptr;
    SCRATCH(0, sizeof(double) * n_t, d_t);
    SCRATCH(1, sizeof(double) * 10 * (size_t)n_t * nb, d_out);
    SCRATCH(2, sizeof(int) * (size_t)n_t * nb, d_st);
    CU(cudaMemcpy(d_t, t, sizeof(double) * n_t, cudaMemcpyHostToDevice));
    cudaError_t e = (math == ASSIST_GPU_MATH_FAST) ? ab_launch_ephem_eval_fast(E, d_t, n_t, d_out, d_st, 0)
                                                   : ab_launch_ephem_eval_strict(E, d_t, n_t, d_out, d_st, 0);
    AB_COUNT(1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(double) * 10 * (size_t)n_t * nb, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && status) e = cudaMemcpy(status, d_st, sizeof(int) * (size_t)n_t * nb, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "ephem_eval: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int assist_gpu_eval_forces(const struct assist_ephem* ephem, const struct assist_gpu_options* opt,
                                      int n_sys, int n_var, const double* t, int t_per_system,
                                      const double* state, const double* params, double* acc, int* status) {
    AbEphem E;
    int rc = build_ephem(ephem, &E);
    if (rc) return rc;
    if (n_var < 0 || n_var > AB_NVMAX) return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "n_var=%d (max %d)", n_var, AB_NVMAX);
    if (n_sys <= 0) return 0;
    AbForceOpts F;
    build_force_opts(opt, params != NULL, &F);
    const int K = 1 + n_var;
    const size_t nt = t_per_system ? n_sys : 1;
    double *d_t = nullptr, *d_state = nullptr, *d_prm = nullptr, *d_acc = nullptr;
    int* d_st = nullptr;
    SCRATCH(0, sizeof(double) * nt, d_t);
    SCRATCH(1, sizeof(double) * 6 * (size_t)n_sys * K, d_state);
    SCRATCH(2, sizeof(double) * 3 * (size_t)n_sys * K, d_acc);
    SCRATCH(3, sizeof(int) * (size_t)n_sys, d_st);
    if (params) {
        SCRATCH(4, sizeof(double) * 3 * (size_t)n_sys * K, d_prm);
        CU(cudaMemcpy(d_prm, params, sizeof(double) * 3 * (size_t)n_sys * K, cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(d_t, t, sizeof(double) * nt, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_state, state, sizeof(double) * 6 * (size_t)n_sys * K, cudaMemcpyHostToDevice));
    cudaError_t e = (opt->math == ASSIST_GPU_MATH_FAST)
        ? ab_launch_force_eval_fast(E, F, n_sys, K, d_t, t_per_system, d_state, d_prm, d_acc, d_st, 0)
        : ab_launch_force_eval_strict(E, F, n_sys, K, d_t, t_per_system, d_state, d_prm, d_acc, d_st, 0);
    AB_COUNT(1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(acc, d_acc, sizeof(double) * 3 * (size_t)n_sys * K, cudaMemcpyDeviceToHost);
    int first_err = 0;
    if (e == cudaSuccess) {
        int* hst = (int*)malloc(sizeof(int) * n_sys);
        e = cudaMemcpy(hst, d_st, sizeof(int) * n_sys, cudaMemcpyDeviceToHost);
        for (int i = 0; i < n_sys; i++) {
            if (status) status[i] = hst[i];
            if (hst[i] && !first_err) first_err = hst[i];
        }
        free(hst);
    }
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "eval_forces: %s", cudaGetErrorString(e));
    if (first_err) return set_err(first_err, "%s", assist_error_messages[first_err]);
    return 0;
}


/* ---- single-target evaluators behind the reference's spk.h / ascii_ephem.h entry points ------------------- */

/* out[9] = u v w of one SPK target (index into file->targets) at jd_ref + jd_rel.  mode as spk_target_kernel;
 * emb_index >= 0: that target is added first (Earth and Moon are given relative to the EMB). */
extern "C" int ab_gpu_spk_target_eval(struct spk_s* file, int target_index, int emb_index, double jd_ref, double jd_rel,
                                      int mode, const double* ud, double* out) {
    int dev;
    int rc = ensure_device(&dev);
    if (rc) return rc;
    if (!file || target_index < 0 || target_index >= file->num) return set_err(ASSIST_GPU_ERR_ARG, "bad SPK target");
    AbSpkDesc* d = nullptr;
    if ((rc = spk_desc(file, &d))) return rc;
    if ((rc = upload_packed_spk(&file->b200_dev_image[dev], file, d->off, d->words))) return rc;
    double* d_out = nullptr;
    CU(cudaMalloc((void**)&d_out, sizeof(double) * 9));
    const double one[3] = {1.0, 1.0, 1.0};
    const AbSpkTarget& tg = d->tg[target_index];
    const AbSpkTarget& emb = d->tg[emb_index >= 0 ? emb_index : target_index];
    cudaError_t e = ab_launch_spk_target_strict((const double*)file->b200_dev_image[dev], tg, emb_index >= 0 ? 1 : 0, emb,
                                                jd_ref, jd_rel, mode, ud ? ud : one, d_out, 0);
    AB_COUNT(1);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(double) * 9, cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "spk_target_eval: %s", cudaGetErrorString(e));
    return 0;
}

/* assist_ascii_work on the device: P holds niv * ncm * ncf coefficients (host memory); out[3][ncm]. */
extern "C" int ab_gpu_ascii_work(const double* P, int ncm, int ncf, int niv, double t0, double t1, double* out) {
    int dev;
    int rc = ensure_device(&dev);
    if (rc) return rc;
    if (!P || ncm < 1 || ncm > 3 || ncf < 2 || ncf > 32 || niv < 1) return set_err(ASSIST_GPU_ERR_ARG, "assist_ascii_work: bad shape");
    const size_t n = (size_t)ncm * ncf * niv;
    double *d_P = nullptr, *d_out = nullptr;
    CU(cudaMalloc((void**)&d_P, sizeof(double) * n));
    CU(cudaMalloc((void**)&d_out, sizeof(double) * 9));
    cudaError_t e = cudaMemcpy(d_P, P, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { e = ab_launch_ascii_work_strict(d_P, ncm, ncf, niv, t0, t1, d_out, 0); AB_COUNT(1); }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(double) * 3 * ncm, cudaMemcpyDeviceToHost);
    cudaFree(d_P); cudaFree(d_out);
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "ascii_work: %s", cudaGetErrorString(e));
    return 0;
}

/* ------------------------------------------------------------------------ */
/* batches                                                                  */
/* ------------------------------------------------------------------------ */

#ifndef AB_DEFAULT_SLICE_DAYS
#define AB_DEFAULT_SLICE_DAYS 256.0
#endif

struct assist_gpu_batch {
    const struct assist_ephem* ephem;
    int n, nvar, K, C, mode, device;
    struct assist_gpu_options opt;
    AbBatch d;
    char* block;            /* one device allocation holding every array */
    size_t block_bytes;
    char* snapshot;         /* device copy made by assist_gpu_batch_snapshot */
    double* d_stage;        /* [n][K][6] AoS staging on the device */
    double* d_stage_prm;    /* [n][K][3] */
    double* d_out;          /* dense output staging */
    size_t d_out_bytes;
    AbBatch w;              /* working slots of the work-queue scheduler (per-particle mode) */
    char* wblock;
    unsigned long long* d_queue;
    int* d_slice_done;      /* [n] time slices completed per system (work-queue scheduler) */
    int* d_slice_epoch;     /* [n] next output epoch per system */
    double* d_trange;       /* [2] min / max of the systems' times */
    double slice_days;      /* length of a time slice; 0: one slice */
    int sched_queue;        /* 1: work-queue kernel, 0: capped launches + straggler packing */
    int sched_coop;         /* 1 (default): pp_coop_kernel wherever it applies (no variational particles, one EIH source, barycentric) */
    AbBatch wc;             /* working slots of pp_coop_kernel: 32 per CTA */
    char* wcblock;
    int* d_order;               /* work queue: systems in the order of their expected step counts, longest first (or NULL) */
    double* d_gtab;             /* pp_coop_kernel: the CTAs' global tables (ABC_GT_DOUBLES each) */
    int coop_grid;
    long long attempt_budget; /* step attempts per system and call (pp_coop_kernel); <= 0: unlimited */
    int* d_active[2];       /* ping-pong lists of systems still integrating */
    int* d_count;           /* length of the list being built */
    long long step_cap;     /* accepted steps per system per launch (per-particle mode) */
    cudaEvent_t ev0, ev1;
    struct assist_gpu_stats stats;
};

/* Append the systems of `in` (or 0..n_in-1) whose integrate() has not returned yet to `out`. */
__global__ void compact_active_kernel(const int* __restrict__ status, const int* __restrict__ in, int n_in,
                                      int* __restrict__ out, int* __restrict__ count) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int i = -1;
    bool act = false;
    if (tid < n_in) {
        i = in ? in[tid] : tid;
        act = status[i] < 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (act) out[base + __popc(m & ((1u << lane) - 1u))] = i;
}

__global__ void aos_to_soa_kernel(const double* __restrict__ aos, int n, int K, int width, int off, int cnt, double* __restrict__ soa) {
    /* aos[i][j][width]; copies fields off..off+cnt-1 of body j to soa[(3*j + c) * n + i] */
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * K) return;
    const long long i = gid % n;
    const int j = (int)(gid / n);
    for (int c = 0; c < cnt; c++) soa[((long long)(3 * j + c)) * n + i] = aos[(i * K + j) * width + off + c];
}

__global__ void soa_to_aos_kernel(const double* __restrict__ soa, int n, int K, int width, int off, int cnt, double* __restrict__ aos) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * K) return;
    const long long i = gid % n;
    const int j = (int)(gid / n);
    for (int c = 0; c < cnt; c++) aos[(i * K + j) * width + off + c] = soa[((long long)(3 * j + c)) * n + i];
}

__global__ void fill_kernel(double* p, long long n, double v) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) p[gid] = v;
}

__global__ void fill_int_kernel(int* p, long long n, int v) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) p[gid] = v;
}

static inline int blocks_for(long long n) { return (int)((n + 255) / 256); }

/* Carve one device allocation into the arrays of an AbBatch for n systems of K bodies. */
static size_t batch_bytes(size_t n, size_t C) {
    size_t bytes = (12 * C * n + 42 * C * n + 4 * n) * sizeof(double);
    bytes += 4 * n * sizeof(unsigned long long);
    bytes += 2 * n * sizeof(int);
    bytes += sizeof(AbShared) + 64;
    return bytes;
}

static void layout_batch(AbBatch& d, char* block, size_t n, int K, int mode) {
    const size_t C = 3 * (size_t)K, per = C * n;
    double* p = (double*)block;
    d.n = (int)n; d.K = K; d.C = (int)C; d.mode = mode;
    d.pos = p; p += per; d.vel = p; p += per; d.acc = p; p += per;
    d.x0 = p; p += per; d.v0 = p; p += per; d.a0 = p; p += per; d.csx = p; p += per; d.csv = p; p += per;
    d.ls_pos = p; p += per; d.ls_vel = p; p += per; d.ls_acc = p; p += per; d.prm = p; p += per;
    d.b = p; p += 7 * per; d.g = p; p += 7 * per; d.e = p; p += 7 * per;
    d.csb = p; p += 7 * per; d.br = p; p += 7 * per; d.er = p; p += 7 * per;
    d.t = p; p += n; d.dt = p; p += n; d.dt_last = p; p += n; d.last_full_dt = p; p += n;
    unsigned long long* q = (unsigned long long*)p;
    d.steps = q; q += n; d.rejected = q; q += n; d.iters = q; q += n; d.evals = q; q += n;
    d.sh = (AbShared*)q;
    int* ip = (int*)((char*)q + ((sizeof(AbShared) + 63) / 64) * 64);
    d.nv = ip; ip += n; d.status = ip; ip += n;
}

extern "C" assist_gpu_batch* assist_gpu_batch_create(const struct assist_ephem* ephem, int n_sys, int n_var, int mode) {
    int dev;
    if (ensure_device(&dev)) return NULL;
    if (n_sys <= 0 || n_var < 0 || n_var > AB_NVMAX) {
        set_err(ASSIST_GPU_ERR_UNSUPPORTED, "batch of %d systems with %d variational particles each (max %d) is not supported", n_sys, n_var, AB_NVMAX);
        return NULL;
    }
    assist_gpu_batch* b = (assist_gpu_batch*)calloc(1, sizeof(assist_gpu_batch));
    b->ephem = ephem; b->n = n_sys; b->nvar = n_var; b->K = 1 + n_var; b->C = 3 * b->K; b->mode = mode; b->device = dev;
    assist_gpu_default_options(&b->opt);
    const size_t n = n_sys, C = b->C;
    const size_t bytes = batch_bytes(n, C);
    if (cudaMalloc((void**)&b->block, bytes) != cudaSuccess) {
        set_err(ASSIST_GPU_ERR_CUDA, "cudaMalloc of %zu bytes failed", bytes);
        free(b);
        return NULL;
    }
    cudaMemset(b->block, 0, bytes);
    b->block_bytes = bytes;
    AbBatch& d = b->d;
    layout_batch(d, b->block, n, b->K, mode);
    {
        const char* sc = getenv("ASSIST_B200_SCHED");
        b->sched_queue = !(sc && !strcmp(sc, "capped"));
        b->sched_coop = !(sc && (!strcmp(sc, "capped") || !strcmp(sc, "queue")));
        const char* ab = getenv("ASSIST_B200_ATTEMPT_BUDGET");
        b->attempt_budget = ab ? atoll(ab) : 50000000LL;
        const char* sd = getenv("ASSIST_B200_SLICE_DAYS");
        b->slice_days = sd ? atof(sd) : AB_DEFAULT_SLICE_DAYS;
        if (!(b->slice_days >= 0.0)) b->slice_days = 0.0;
    }
    d.epsilon = b->opt.epsilon; d.min_dt = b->opt.min_dt; d.has_params = 0;
    cudaMalloc((void**)&b->d_stage, sizeof(double) * 6 * n * b->K);
    cudaMalloc((void**)&b->d_stage_prm, sizeof(double) * 3 * n * b->K);
    cudaMalloc((void**)&b->d_active[0], sizeof(int) * n);
    cudaMalloc((void**)&b->d_active[1], sizeof(int) * n);
    cudaMalloc((void**)&b->d_count, sizeof(int));
    {
        const char* cap = getenv("ASSIST_B200_STEP_CAP");
        b->step_cap = cap ? atoll(cap) : 32;
    }
    cudaEventCreate(&b->ev0);
    cudaEventCreate(&b->ev1);
    fill_int_kernel<<<blocks_for(n), 256>>>(d.nv, n, n_var); AB_COUNT(1);
    fill_int_kernel<<<blocks_for(n), 256>>>(d.status, n, -3); AB_COUNT(1);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        set_err(ASSIST_GPU_ERR_CUDA, "batch initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        assist_gpu_batch_free(b);
        return NULL;
    }
    return b;
}

extern "C" void assist_gpu_batch_free(assist_gpu_batch* b) {
    if (!b) return;
    cudaFree(b->block); cudaFree(b->snapshot); cudaFree(b->d_stage); cudaFree(b->d_stage_prm); cudaFree(b->d_out);
    cudaFree(b->d_active[0]); cudaFree(b->d_active[1]); cudaFree(b->d_count);
    cudaFree(b->wcblock);
    cudaFree(b->d_gtab);
    cudaFree(b->d_order);
    cudaFree(b->wblock); cudaFree(b->d_queue); cudaFree(b->d_slice_done); cudaFree(b->d_slice_epoch); cudaFree(b->d_trange);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    free(b);
}

extern "C" int assist_gpu_batch_set_options(assist_gpu_batch* b, const struct assist_gpu_options* opt) {
    if (!b || !opt) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    b->opt = *opt;
    b->d.epsilon = opt->epsilon;
    b->d.min_dt = opt->min_dt;
    return 0;
}

static int upload_particles(assist_gpu_batch* b, const double* state) {
    const size_t n = b->n;
    CU(cudaMemcpy(b->d_stage, state, sizeof(double) * 6 * n * b->K, cudaMemcpyHostToDevice));
    aos_to_soa_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d_stage, b->n, b->K, 6, 0, 3, b->d.pos); AB_COUNT(1);
    aos_to_soa_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d_stage, b->n, b->K, 6, 3, 3, b->d.vel); AB_COUNT(1);
    CU(cudaGetLastError());
    return 0;
}

/* Queue order of a per-particle batch: a system is a serial chain of steps, so a launch ends with its longest
 * system; handing the long ones out FIRST keeps the tail of the launch short (strong scaling: VERDICT r1 item 5).
 * Expected step count ~ a^-3/2 (1 - e)^-1 from the osculating elements of the (nearly heliocentric) barycentric
 * state -- log-correlation 0.98 with the measured counts of the C3 population.  A counting sort into 1024 buckets,
 * O(n) on the host.  Results do not depend on the order (tests/test_gpu_parity.py::test_properties_at_scale). */
extern "C" void ab_gpu_cost_order_host(const double* state, int n_, int K, int* order) {
    const size_t n = (size_t)n_;
    const double gms = 2.959122082855911e-4;        /* GM of the Sun, AU^3 / day^2: only the ORDER of the costs matters */
    const int NB = 1024;
    std::vector<unsigned short> key(n);
    std::vector<int> count(NB + 1, 0);
    const size_t stride = (size_t)K * 6;
    for (size_t i = 0; i < n; i++) {
        const double* s = state + i * stride;
        const double r = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        const double v2 = s[3] * s[3] + s[4] * s[4] + s[5] * s[5];
        const double inv_a = 2.0 / r - v2 / gms;                 /* 1 / a */
        double cost = 1e30;                                      /* unbound or degenerate: first */
        if (inv_a > 0.0 && r > 0.0) {
            const double hx = s[1] * s[5] - s[2] * s[4], hy = s[2] * s[3] - s[0] * s[5], hz = s[0] * s[4] - s[1] * s[3];
            double e2 = 1.0 - (hx * hx + hy * hy + hz * hz) * inv_a / gms;
            if (e2 < 0.0) e2 = 0.0;
            const double ome = 1.0 - sqrt(e2);
            if (ome > 1e-6) cost = inv_a * sqrt(inv_a) / ome;
        }
        /* log2 scale: 2^-12 .. 2^20 over the buckets, longest expected = bucket 0 */
        double lg = log2(cost);
        if (!(lg == lg)) lg = 20.0;
        int k = (int)((20.0 - lg) * (NB / 32.0));
        if (k < 0) k = 0;
        if (k >= NB) k = NB - 1;
        key[i] = (unsigned short)k;
        count[k + 1]++;
    }
    for (int k = 0; k < NB; k++) count[k + 1] += count[k];
    for (size_t i = 0; i < n; i++) order[(size_t)count[key[i]]++] = (int)i;
}

static int build_queue_order(assist_gpu_batch* b, const double* state) {
    const size_t n = b->n;
    if (b->mode != ASSIST_GPU_PER_PARTICLE || n < 2 || (getenv("ASSIST_B200_QUEUE_ORDER") && atoi(getenv("ASSIST_B200_QUEUE_ORDER")) == 0)) {
        if (b->d_order) { cudaFree(b->d_order); b->d_order = nullptr; }
        return 0;
    }
    std::vector<int> order(n);
    ab_gpu_cost_order_host(state, (int)n, b->K, order.data());
    if (!b->d_order) CU(cudaMalloc((void**)&b->d_order, sizeof(int) * n));
    CU(cudaMemcpy(b->d_order, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int assist_gpu_batch_set_state(assist_gpu_batch* b, double t0, double dt0, const double* state,
                                          const double* params, const int* nvar_per_system) {
    if (!b || !state) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    const size_t n = b->n;
    /* fresh IAS15 history: everything zero (REBOUND zeroes b, e, csx, csv ... on allocation) */
    CU(cudaMemset(b->block, 0, b->block_bytes));
    int rc = upload_particles(b, state);
    if (rc) return rc;
    rc = build_queue_order(b, state);
    if (rc) return rc;
    if (params) {
        CU(cudaMemcpy(b->d_stage_prm, params, sizeof(double) * 3 * n * b->K, cudaMemcpyHostToDevice));
        aos_to_soa_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d_stage_prm, b->n, b->K, 3, 0, 3, b->d.prm); AB_COUNT(1);
        b->d.has_params = 1;
    } else {
        b->d.has_params = 0;
    }
    if (nvar_per_system) {
        for (size_t i = 0; i < n; i++)
            if (nvar_per_system[i] < 0 || nvar_per_system[i] > b->nvar) return set_err(ASSIST_GPU_ERR_ARG, "nvar_per_system[%zu] out of range", i);
        CU(cudaMemcpy(b->d.nv, nvar_per_system, sizeof(int) * n, cudaMemcpyHostToDevice));
    } else {
        fill_int_kernel<<<blocks_for(n), 256>>>(b->d.nv, n, b->nvar); AB_COUNT(1);
    }
    fill_kernel<<<blocks_for(n), 256>>>(b->d.t, n, t0); AB_COUNT(1);
    fill_kernel<<<blocks_for(n), 256>>>(b->d.dt, n, dt0); AB_COUNT(1);
    fill_int_kernel<<<blocks_for(n), 256>>>(b->d.status, n, -3); AB_COUNT(1);
    AbShared sh;
    memset(&sh, 0, sizeof(sh));
    sh.t = t0; sh.dt = dt0; sh.status = -3;
    CU(cudaMemcpy(b->d.sh, &sh, sizeof(sh), cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    memset(&b->stats, 0, sizeof(b->stats));
    return 0;
}

extern "C" int assist_gpu_batch_update_particles(assist_gpu_batch* b, const double* state) {
    if (!b || !state) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    int rc = upload_particles(b, state);
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    return 0;
}

extern "C" int assist_gpu_batch_set_time(assist_gpu_batch* b, double t, double dt) {
    if (!b) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    const size_t n = b->n;
    fill_kernel<<<blocks_for(n), 256>>>(b->d.t, n, t); AB_COUNT(1);
    fill_kernel<<<blocks_for(n), 256>>>(b->d.dt, n, dt); AB_COUNT(1);
    AbShared sh;
    CU(cudaMemcpy(&sh, b->d.sh, sizeof(sh), cudaMemcpyDeviceToHost));
    sh.t = t; sh.dt = dt;
    CU(cudaMemcpy(b->d.sh, &sh, sizeof(sh), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int assist_gpu_batch_snapshot(assist_gpu_batch* b) {
    if (!b) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    if (!b->snapshot) CU(cudaMalloc((void**)&b->snapshot, b->block_bytes));
    CU(cudaMemcpy(b->snapshot, b->block, b->block_bytes, cudaMemcpyDeviceToDevice));
    return 0;
}

extern "C" int assist_gpu_batch_restore(assist_gpu_batch* b) {
    if (!b || !b->snapshot) return set_err(ASSIST_GPU_ERR_ARG, "no snapshot");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(b->block, b->snapshot, b->block_bytes, cudaMemcpyDeviceToDevice, 0));
    memset(&b->stats, 0, sizeof(b->stats));
    return 0;
}

static int finish_launch(assist_gpu_batch* b, cudaError_t e, const char* what) {
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
    CU(cudaEventRecord(b->ev1, 0));
    cudaError_t s = cudaEventSynchronize(b->ev1);
    if (s != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    b->stats.last_kernel_ms = ms;
    b->stats.kernel_launches++;
    AB_COUNT(1);
    return 0;
}

extern "C" int assist_gpu_batch_integrate(assist_gpu_batch* b, double t_end, int exact_finish_time, long max_steps) {
    return ab_gpu_batch_integrate_ex(b, t_end, exact_finish_time, max_steps, 0);
}

/* A record boundary of the small-body kernel (else of the planets file), as a time relative to jd_ref. */
static double ab_slice_anchor(const struct assist_ephem* e) {
    if (e->spk_asteroids && e->spk_asteroids->num > 0) return e->spk_asteroids->targets[0].beg - e->jd_ref;
    if (e->spk_planets && e->spk_planets->num > 0) return e->spk_planets->targets[0].beg - e->jd_ref;
    return 0.0;
}

/* min / max of the times of the systems that can still run */
__global__ void trange_kernel(const double* __restrict__ t, const int* __restrict__ status, int n, double* __restrict__ out) {
    double lo = 1e300, hi = -1e300;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (status[i] >= 1000) continue;
        const double v = t[i];
        if (v == v) { lo = fmin(lo, v); hi = fmax(hi, v); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        /* doubles compare like their bit patterns once the sign is folded in */
        auto key = [](double x) { long long b = __double_as_longlong(x); return (unsigned long long)(b < 0 ? ~b : (b | 0x8000000000000000LL)); };
        atomicMin((unsigned long long*)&out[0], key(lo));
        atomicMax((unsigned long long*)&out[1], key(hi));
    }
}

static double unkey(unsigned long long k) {
    long long b = (k & 0x8000000000000000ULL) ? (long long)(k & 0x7fffffffffffffffULL) : ~(long long)k;
    double x;
    memcpy(&x, &b, sizeof(x));
    return x;
}

/* Time slices of one call (see AbSlices): windows of slice_days on the record grid of the ephemeris, from the
 * earliest system towards t_end.  One slice when slicing is off or the systems lie on both sides of t_end. */
static int build_slices(assist_gpu_batch* b, const AbEphem& E, double t_end, AbSlices* SL) {
    SL->origin = 0.0; SL->wlen = 1.0; SL->n_win = 1;
    SL->done = b->d_slice_done; SL->epoch = b->d_slice_epoch; SL->order = b->d_order;
    SL->attempt_budget = b->attempt_budget;
    if (b->slice_days > 0.0 && t_end == t_end) {
        unsigned long long init[2] = {~0ULL, 0ULL}, got[2];
        CU(cudaMemcpy(b->d_trange, init, sizeof(init), cudaMemcpyHostToDevice));
        trange_kernel<<<256, 256>>>(b->d.t, b->d.status, b->n, b->d_trange); AB_COUNT(1);
        CU(cudaMemcpy(got, b->d_trange, sizeof(got), cudaMemcpyDeviceToHost));
        if (got[0] != ~0ULL) {
            const double lo = unkey(got[0]), hi = unkey(got[1]);
            const double D = b->slice_days;
            /* a record boundary of the ephemeris, as a time relative to jd_ref */
            double g0 = 0.0;
            if (b->ephem) g0 = ab_slice_anchor(b->ephem);
            double origin = 0.0, wlen = 0.0, span = -1.0;
            if (hi <= t_end) { origin = g0 + floor((lo - g0) / D) * D; wlen = D; span = t_end - origin; }
            else if (lo >= t_end) { origin = g0 + ceil((hi - g0) / D) * D; wlen = -D; span = origin - t_end; }
            if (span > 0.0) {
                double nw = ceil(span / D);
                if (nw < 1.0) nw = 1.0;
                if (nw > 1048576.0) { nw = 1.0; }
                else { SL->origin = origin; SL->wlen = wlen; }
                SL->n_win = (int)nw;
            }
        }
    }
    if (SL->n_win > 1) CU(cudaMemsetAsync(b->d_slice_done, 0, sizeof(int) * (size_t)b->n, 0));
    return 0;
}

/* Working batch of the work-queue scheduler: one slot per resident thread of the per-particle kernels. */
static int ensure_working_batch(assist_gpu_batch* b) {
    if (!b->wblock) {
        const bool k1 = (b->K == 1);
        int threads = 0, t2 = 0;
        cudaError_t eo = k1 ? ab_pp_resident_threads_k1_strict(&threads) : ab_pp_resident_threads_kv_strict(&threads);
        if (eo != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "occupancy query failed: %s", cudaGetErrorString(eo));
        /* both math variants must fit: take the smaller resident count */
        eo = k1 ? ab_pp_resident_threads_k1_fast(&t2) : ab_pp_resident_threads_kv_fast(&t2);
        if (eo == cudaSuccess && t2 < threads) threads = t2;
        size_t slots = (size_t)threads;
        if (slots > (size_t)b->n) slots = ((size_t)b->n + 127) / 128 * 128;
        CU(cudaMalloc((void**)&b->wblock, batch_bytes(slots, b->C)));
        CU(cudaMemset(b->wblock, 0, batch_bytes(slots, b->C)));
        layout_batch(b->w, b->wblock, slots, b->K, b->mode);
        if (!b->d_queue) CU(cudaMalloc((void**)&b->d_queue, sizeof(unsigned long long)));
        if (!b->d_slice_done) CU(cudaMalloc((void**)&b->d_slice_done, sizeof(int) * (size_t)b->n));
        if (!b->d_slice_epoch) CU(cudaMalloc((void**)&b->d_slice_epoch, sizeof(int) * (size_t)b->n));
        if (!b->d_trange) CU(cudaMalloc((void**)&b->d_trange, sizeof(double) * 2));
    }
    b->w.epsilon = b->d.epsilon; b->w.min_dt = b->d.min_dt; b->w.has_params = b->d.has_params;
    return 0;
}


/* ---- pp_coop_kernel: who runs it, its working slots and the plan of its worker warps ---- */

static bool coop_applies(const assist_gpu_batch* b, const AbForceOpts& F) {
    return b->sched_coop && b->mode == ASSIST_GPU_PER_PARTICLE && b->K == 1 && F.gr_eih_sources == 1 && !F.geocentric &&
           !(b->ephem->spk_asteroids && b->ephem->spk_asteroids->num > AB_MAX_AST);      /* sb441-n373: the work-queue kernel */
}

/* Spread the force terms over the worker warps: longest task first onto the least loaded warp.  The weights are
 * rough lengths (cycles) of the dependent chains in the strict build: a phase ends with its slowest warp.  Planets run
 * two side by side, asteroids four (coop_device.cuh).  ASSIST_B200_COOP_COSTS="earth,eih,sunj2,planet,asteroid"
 * overrides the weights (tuning aid). */
static void build_coop_plan(const struct assist_ephem* e, const AbEphem& E, const AbForceOpts& F, long long budget, AbcPlan* plan) {
    (void)e;
    int c_earth = 500, c_eih = 400, c_sunj2 = 300, c_planet = 250, c_ast = 200;
    if (const char* cs = getenv("ASSIST_B200_COOP_COSTS")) sscanf(cs, "%d,%d,%d,%d,%d", &c_earth, &c_eih, &c_sunj2, &c_planet, &c_ast);
    struct Task { int kind; int cost; };      /* kind: ABC_T_* or body index 0..26 */
    std::vector<Task> tasks;
    if ((F.forces & 0x08) && F.has_params) tasks.push_back({ABC_T_NG, 3000});
    if (F.forces & 0x10) tasks.push_back({ABC_T_EARTHJ, c_earth});
    if (F.forces & 0x40) tasks.push_back({ABC_T_EIHSRC, c_eih});
    if (F.forces & 0x20) tasks.push_back({ABC_T_SUNJ2, c_sunj2});
    if (F.forces & 0x80) tasks.push_back({ABC_T_GRSIMPLE, 600});
    if (F.forces & 0x100) tasks.push_back({ABC_T_GRPOT, 500});
    /* a body is evaluated when its direct term is on, or (planets) when the EIH potential sum needs its separation */
    const bool eih = (F.forces & 0x40) != 0;
    for (int i = 0; i < AB_NPLANETS; i++) {
        const bool on = (i == 0) ? (F.forces & 0x01) != 0 : (F.forces & 0x02) != 0;
        if (on || eih) tasks.push_back({i, c_planet});
    }
    if (F.forces & 0x04) for (int m = 0; m < E.n_ast; m++) tasks.push_back({AB_NPLANETS + m, c_ast});
    std::stable_sort(tasks.begin(), tasks.end(), [](const Task& a, const Task& b) { return a.cost > b.cost; });
    memset(plan, 0, sizeof(*plan));
    int load[ABC_NWORK] = {0};
    for (int w = 0; w < ABC_NWORK; w++) { plan->w[w].scalar[0] = ABC_T_NONE; plan->w[w].scalar[1] = ABC_T_NONE; }
    for (const Task& t : tasks) {
        int best = -1;
        for (int w = 0; w < ABC_NWORK; w++) {
            const AbcWorkerPlan& wp = plan->w[w];
            const bool fits = (t.kind >= ABC_T_EARTHJ) ? (wp.scalar[1] == ABC_T_NONE)
                              : (t.kind < AB_NPLANETS ? wp.nbody < ABC_MAX_GROUP : wp.nast < ABC_MAX_GROUP);
            if (fits && (best < 0 || load[w] < load[best])) best = w;
        }
        if (best < 0) continue;      /* cannot happen: 6 single-body terms, 11 + 16 bodies, 4 x (2 + 16 + 16) places */
        AbcWorkerPlan& wp = plan->w[best];
        if (t.kind >= ABC_T_EARTHJ) { if (wp.scalar[0] == ABC_T_NONE) wp.scalar[0] = (unsigned char)t.kind; else wp.scalar[1] = (unsigned char)t.kind; }
        else if (t.kind < AB_NPLANETS) wp.body[wp.nbody++] = (unsigned char)t.kind;
        else wp.ast[wp.nast++] = (unsigned char)(t.kind - AB_NPLANETS);
        load[best] += t.cost;
    }
    plan->attempt_budget = budget;
}

/* host copy of the asteroid descriptors (masses refreshed by build_ephem) for the kernel's constant memory */
static const AbSpkTarget* coop_ast_tg(const struct assist_ephem* e) {
    if (!e || !e->spk_asteroids || !e->spk_asteroids->b200_host_desc) return nullptr;
    return ((const AbSpkDesc*)e->spk_asteroids->b200_host_desc)->tg.data();
}

/* ASSIST_B200_COOP_TIMING=1: 16 device counters of the kernel's phases, printed to stderr after every launch */
static unsigned long long* g_coop_timing = nullptr;
static unsigned long long* coop_timing(assist_gpu_batch* b) {
    (void)b;
    if (!getenv("ASSIST_B200_COOP_TIMING")) return nullptr;
    if (!g_coop_timing) { if (cudaMalloc((void**)&g_coop_timing, 16 * sizeof(unsigned long long)) != cudaSuccess) return nullptr; }
    cudaMemsetAsync(g_coop_timing, 0, 16 * sizeof(unsigned long long), 0);
    return g_coop_timing;
}
static void coop_timing_report(assist_gpu_batch* b) {
    if (!g_coop_timing || !getenv("ASSIST_B200_COOP_TIMING")) return;
    unsigned long long t[16];
    if (cudaMemcpy(t, g_coop_timing, sizeof(t), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    /* slots 5 and 6 (force terms / component warps of a node round) are printed as ONE line: the control warp reads the
     * clock as soon as it ARRIVES at a barrier (BAR.SYNC.DEFER_BLOCKING lets it run on to the next blocking
     * instruction), so the split between two back-to-back phases of other warps is not reliable, their sum is */
    static const char* nm[10] = {"bookkeeping", "fill", "-", "fill-check", "-", "node rounds", "-", "convergence", "dt-control", "advance+store"};
    t[5] += t[6]; t[6] = 0;
    double tot = 0; for (int q = 0; q < 10; q++) tot += (double)t[q];
    fprintf(stderr, "[assist-b200 coop timing] grid %d, attempts/CTA %.0f, node rounds/CTA %.0f, cycles/CTA %.3e\n", b->coop_grid,
            (double)t[10] / b->coop_grid, (double)t[11] / b->coop_grid, tot / b->coop_grid);
    for (int q = 0; q < 10; q++) if (nm[q][0] != '-') fprintf(stderr, "    %-14s %5.1f %%   %9.0f cycles per attempt\n", nm[q], 100.0 * t[q] / tot, (double)t[q] / (double)(t[10] ? t[10] : 1));
}

static int ensure_coop_batch(assist_gpu_batch* b, bool fast) {
    if (!b->wcblock) {
        int g1 = 0, g2 = 0;
        cudaError_t eo = ab_pp_coop_max_grid_strict(&g1);
        if (eo != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "pp_coop occupancy query failed: %s", cudaGetErrorString(eo));
        eo = ab_pp_coop_max_grid_fast(&g2);
        if (eo != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "pp_coop occupancy query failed: %s", cudaGetErrorString(eo));
        int grid = g1 < g2 ? g1 : g2;
        const int need = (b->n + ABC_GROUPS * ABC_SLOTS - 1) / (ABC_GROUPS * ABC_SLOTS);
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        b->coop_grid = grid;
        const size_t slots = (size_t)grid * ABC_GROUPS * ABC_SLOTS;
        CU(cudaMalloc((void**)&b->wcblock, batch_bytes(slots, b->C)));
        CU(cudaMemset(b->wcblock, 0, batch_bytes(slots, b->C)));
        layout_batch(b->wc, b->wcblock, slots, b->K, b->mode);
        CU(cudaMalloc((void**)&b->d_gtab, sizeof(double) * (size_t)grid * ABC_GROUPS * ABC_GT_DOUBLES));
        CU(cudaMemset(b->d_gtab, 0, sizeof(double) * (size_t)grid * ABC_GROUPS * ABC_GT_DOUBLES));
        if (!b->d_queue) CU(cudaMalloc((void**)&b->d_queue, sizeof(unsigned long long)));
        if (!b->d_slice_done) CU(cudaMalloc((void**)&b->d_slice_done, sizeof(int) * (size_t)b->n));
        if (!b->d_slice_epoch) CU(cudaMalloc((void**)&b->d_slice_epoch, sizeof(int) * (size_t)b->n));
        if (!b->d_trange) CU(cudaMalloc((void**)&b->d_trange, sizeof(double) * 2));
    }
    (void)fast;
    b->wc.epsilon = b->d.epsilon; b->wc.min_dt = b->d.min_dt; b->wc.has_params = b->d.has_params;
    return 0;
}


/* Host-only test hooks for tests/emul (no device needed): the launch-time structures exactly as the launches build them. */
extern "C" int ab_gpu_build_force_opts_host(const struct assist_gpu_options* o, int has_params, void* F_out, size_t F_bytes) {
    if (F_bytes != sizeof(AbForceOpts)) return set_err(ASSIST_GPU_ERR_ARG, "AbForceOpts size mismatch");
    build_force_opts(o, has_params, (AbForceOpts*)F_out);
    return 0;
}
extern "C" int ab_gpu_build_coop_plan_host(const struct assist_ephem* e, const void* E, const void* F, long long budget, void* plan_out, size_t plan_bytes) {
    if (plan_bytes != sizeof(AbcPlan)) return set_err(ASSIST_GPU_ERR_ARG, "AbcPlan size mismatch");
    build_coop_plan(e, *(const AbEphem*)E, *(const AbForceOpts*)F, budget, (AbcPlan*)plan_out);
    return 0;
}
extern "C" size_t ab_gpu_batch_bytes_host(size_t n, size_t C) { return batch_bytes(n, C); }
extern "C" int ab_gpu_layout_batch_host(void* d, size_t d_bytes, char* block, size_t n, int K, int mode) {
    if (d_bytes != sizeof(AbBatch)) return set_err(ASSIST_GPU_ERR_ARG, "AbBatch size mismatch");
    layout_batch(*(AbBatch*)d, block, n, K, mode);
    return 0;
}

extern "C" int ab_gpu_batch_integrate_ex(assist_gpu_batch* b, double t_end, int exact_finish_time, long max_steps, int flags) {
    if (!b) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    AbEphem E;
    int rc = build_ephem(b->ephem, &E);
    if (rc) return rc;
    AbForceOpts F;
    build_force_opts(&b->opt, b->d.has_params, &F);
    const bool fast = (b->opt.math == ASSIST_GPU_MATH_FAST);
    cudaError_t e;
    if (b->mode == ASSIST_GPU_PER_PARTICLE) {
        /* Every launch advances each running system by at most step_cap accepted steps; the systems
         * whose integrate() has not returned are then packed into a dense list and relaunched.  This
         * keeps the lanes of a warp busy although step counts differ by 10x between particles. */
        const bool k1 = (b->K == 1);
        if (coop_applies(b, F)) {
            /* a CTA per 32 systems, working set on chip (coop_device.cuh) */
            rc = ensure_coop_batch(b, fast);
            if (rc) return rc;
            AbSlices SL;
            const double keep = b->slice_days;
            if (!getenv("ASSIST_B200_SLICE_DAYS")) b->slice_days = 0.0;      /* whole systems: a system is ~30 us per step here, no tail to cut */
            rc = build_slices(b, E, t_end, &SL);
            b->slice_days = keep;
            if (rc) return rc;
            AbcPlan plan;
            build_coop_plan(b->ephem, E, F, b->attempt_budget, &plan);
            CU(cudaMemsetAsync(b->d_queue, 0, sizeof(unsigned long long), 0));
            CU(cudaEventRecord(b->ev0, 0));
            e = fast ? ab_launch_pp_coop_fast(E, F, b->d, b->wc, t_end, exact_finish_time, b->d_queue, SL, NULL, 0, NULL, &plan, coop_ast_tg(b->ephem), b->d_gtab, coop_timing(b), b->coop_grid, 0)
                     : ab_launch_pp_coop_strict(E, F, b->d, b->wc, t_end, exact_finish_time, b->d_queue, SL, NULL, 0, NULL, &plan, coop_ast_tg(b->ephem), b->d_gtab, coop_timing(b), b->coop_grid, 0);
            rc = finish_launch(b, e, "pp_coop");
            coop_timing_report(b);
            return rc;
        }
        if (b->sched_queue) {
            /* work-queue scheduling: a resident grid of threads pulls systems until the queue is empty */
            rc = ensure_working_batch(b);
            if (rc) return rc;
            AbSlices SL;
            rc = build_slices(b, E, t_end, &SL);
            if (rc) return rc;
            CU(cudaMemsetAsync(b->d_queue, 0, sizeof(unsigned long long), 0));
            CU(cudaEventRecord(b->ev0, 0));
            if (k1) e = fast ? ab_launch_pp_queue_k1_fast(E, F, b->d, b->w, t_end, exact_finish_time, b->d_queue, SL, NULL, 0, NULL, 0)
                             : ab_launch_pp_queue_k1_strict(E, F, b->d, b->w, t_end, exact_finish_time, b->d_queue, SL, NULL, 0, NULL, 0);
            else e = fast ? ab_launch_pp_queue_kv_fast(E, F, b->d, b->w, t_end, exact_finish_time, b->d_queue, SL, NULL, 0, NULL, 0)
                          : ab_launch_pp_queue_kv_strict(E, F, b->d, b->w, t_end, exact_finish_time, b->d_queue, SL, NULL, 0, NULL, 0);
            return finish_launch(b, e, "pp_queue");
        }
        CU(cudaEventRecord(b->ev0, 0));
        const int* list = NULL;
        int n_active = b->n, resume = 0, which = 0;
        unsigned long long launches = 0;
        const bool trace = getenv("ASSIST_B200_TRACE") != NULL;
        struct timespec ts0, ts1;
        clock_gettime(CLOCK_MONOTONIC, &ts0);
        while (true) {
            if (k1) e = fast ? ab_launch_pp_integrate_k1_fast(E, F, b->d, t_end, exact_finish_time, resume, b->step_cap, list, n_active, 0)
                             : ab_launch_pp_integrate_k1_strict(E, F, b->d, t_end, exact_finish_time, resume, b->step_cap, list, n_active, 0);
            else e = fast ? ab_launch_pp_integrate_kv_fast(E, F, b->d, t_end, exact_finish_time, resume, b->step_cap, list, n_active, 0)
                          : ab_launch_pp_integrate_kv_strict(E, F, b->d, t_end, exact_finish_time, resume, b->step_cap, list, n_active, 0);
            if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "pp_integrate launch failed: %s", cudaGetErrorString(e));
            launches++;
            if (b->step_cap <= 0) break;
            CU(cudaMemsetAsync(b->d_count, 0, sizeof(int), 0));
            compact_active_kernel<<<(n_active + 255) / 256, 256>>>(b->d.status, list, n_active, b->d_active[which], b->d_count); AB_COUNT(1);
            int count = 0;
            CU(cudaMemcpy(&count, b->d_count, sizeof(int), cudaMemcpyDeviceToHost));
            if (trace) {
                clock_gettime(CLOCK_MONOTONIC, &ts1);
                fprintf(stderr, "[assist-b200] launch %llu: %d systems in, %d still running, %.2f ms\n", launches, n_active, count,
                        (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
                ts0 = ts1;
            }
            if (count == 0) break;
            if (launches > 4000000ULL) return set_err(ASSIST_GPU_ERR_CUDA, "per-particle integration did not finish after %llu launches", launches);
            list = b->d_active[which];
            n_active = count;
            which ^= 1;
            resume = 1;
        }
        AB_COUNT(2 * launches - 2);
        b->stats.kernel_launches += 2 * launches - 2;   /* integrate launches + compaction kernels (finish_launch adds one) */
        return finish_launch(b, cudaSuccess, "pp_integrate");
    }
    /* shared step: barrier counter and reduction slots start from zero */
    CU(cudaMemsetAsync((char*)b->d.sh + offsetof(AbShared, barrier), 0, sizeof(AbShared) - offsetof(AbShared, barrier), 0));
    CU(cudaEventRecord(b->ev0, 0));
    e = fast ? ab_launch_sh_integrate_fast(E, F, b->d, t_end, exact_finish_time, (long long)max_steps, flags, 0)
             : ab_launch_sh_integrate_strict(E, F, b->d, t_end, exact_finish_time, (long long)max_steps, flags, 0);
    rc = finish_launch(b, e, "sh_integrate");
    if (rc) return rc;
    AbShared sh;
    CU(cudaMemcpy(&sh, b->d.sh, sizeof(sh), cudaMemcpyDeviceToHost));
    if (sh.err_status) return set_err(sh.err_status, "%s", assist_error_messages[sh.err_status]);
    return 0;
}

extern "C" int assist_gpu_batch_integrate_or_interpolate(assist_gpu_batch* b, const double* times, int n_times, double* out) {
    if (!b || !times || !out || n_times <= 0) return set_err(ASSIST_GPU_ERR_ARG, "bad argument");
    if (b->mode != ASSIST_GPU_PER_PARTICLE)
        return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "batched epochs need a per-particle batch; use assist_gpu_batch_integrate + _interpolate for shared-step batches");
    CU(cudaSetDevice(b->device));
    AbEphem E;
    int rc = build_ephem(b->ephem, &E);
    if (rc) return rc;
    AbForceOpts F;
    build_force_opts(&b->opt, b->d.has_params, &F);
    const size_t out_bytes = sizeof(double) * 6 * (size_t)b->n * b->K * n_times;
    if (b->d_out_bytes < out_bytes + sizeof(double) * n_times) {
        cudaFree(b->d_out);
        b->d_out = nullptr; b->d_out_bytes = 0;
        CU(cudaMalloc((void**)&b->d_out, out_bytes + sizeof(double) * n_times));
        b->d_out_bytes = out_bytes + sizeof(double) * n_times;
    }
    double* d_times = (double*)((char*)b->d_out + out_bytes);
    CU(cudaMemcpy(d_times, times, sizeof(double) * n_times, cudaMemcpyHostToDevice));
    /* rows the kernel does not write (variational slots a system does not use, systems that failed earlier) read NaN */
    CU(cudaMemsetAsync(b->d_out, 0xFF, out_bytes, 0));
    const bool fastm = (b->opt.math == ASSIST_GPU_MATH_FAST);
    cudaError_t e;
    if (coop_applies(b, F)) {
        rc = ensure_coop_batch(b, fastm);
        if (rc) return rc;
        bool mono = true;
        const double dir = (n_times > 1) ? (times[n_times - 1] - times[0]) : 0.0;
        for (int q = 1; q < n_times; q++)
            if ((times[q] - times[q - 1]) * dir < 0.0 || !(times[q] == times[q])) mono = false;
        AbSlices SL;
        const double keep = b->slice_days;
        if (!mono || !getenv("ASSIST_B200_SLICE_DAYS")) b->slice_days = 0.0;
        rc = build_slices(b, E, times[n_times - 1], &SL);
        b->slice_days = keep;
        if (rc) return rc;
        if (SL.n_win > 1 && n_times > 1 && SL.wlen * dir < 0.0) SL.n_win = 1;
        AbcPlan plan;
        build_coop_plan(b->ephem, E, F, b->attempt_budget, &plan);
        CU(cudaMemsetAsync(b->d_queue, 0, sizeof(unsigned long long), 0));
        CU(cudaEventRecord(b->ev0, 0));
        e = fastm ? ab_launch_pp_coop_fast(E, F, b->d, b->wc, 0.0, 0, b->d_queue, SL, d_times, n_times, b->d_out, &plan, coop_ast_tg(b->ephem), b->d_gtab, coop_timing(b), b->coop_grid, 0)
                  : ab_launch_pp_coop_strict(E, F, b->d, b->wc, 0.0, 0, b->d_queue, SL, d_times, n_times, b->d_out, &plan, coop_ast_tg(b->ephem), b->d_gtab, coop_timing(b), b->coop_grid, 0);
    } else if (b->sched_queue) {
        rc = ensure_working_batch(b);
        if (rc) return rc;
        /* slices need epochs that run in one direction; anything else is one slice */
        bool mono = true;
        const double dir = (n_times > 1) ? (times[n_times - 1] - times[0]) : 0.0;
        for (int q = 1; q < n_times; q++)
            if ((times[q] - times[q - 1]) * dir < 0.0 || !(times[q] == times[q])) mono = false;
        AbSlices SL;
        const double keep = b->slice_days;
        if (!mono) b->slice_days = 0.0;
        rc = build_slices(b, E, times[n_times - 1], &SL);
        b->slice_days = keep;
        if (rc) return rc;
        if (SL.n_win > 1 && n_times > 1 && SL.wlen * dir < 0.0) {      /* systems ahead of the first epoch: do not slice */
            SL.n_win = 1;
        }
        CU(cudaMemsetAsync(b->d_queue, 0, sizeof(unsigned long long), 0));
        CU(cudaEventRecord(b->ev0, 0));
        if (b->K == 1) e = fastm ? ab_launch_pp_queue_k1_fast(E, F, b->d, b->w, 0.0, 0, b->d_queue, SL, d_times, n_times, b->d_out, 0)
                                 : ab_launch_pp_queue_k1_strict(E, F, b->d, b->w, 0.0, 0, b->d_queue, SL, d_times, n_times, b->d_out, 0);
        else e = fastm ? ab_launch_pp_queue_kv_fast(E, F, b->d, b->w, 0.0, 0, b->d_queue, SL, d_times, n_times, b->d_out, 0)
                       : ab_launch_pp_queue_kv_strict(E, F, b->d, b->w, 0.0, 0, b->d_queue, SL, d_times, n_times, b->d_out, 0);
    } else {
        CU(cudaEventRecord(b->ev0, 0));
        if (b->K == 1) e = fastm ? ab_launch_pp_dense_k1_fast(E, F, b->d, d_times, n_times, b->d_out, 0)
                                 : ab_launch_pp_dense_k1_strict(E, F, b->d, d_times, n_times, b->d_out, 0);
        else e = fastm ? ab_launch_pp_dense_kv_fast(E, F, b->d, d_times, n_times, b->d_out, 0)
                       : ab_launch_pp_dense_kv_strict(E, F, b->d, d_times, n_times, b->d_out, 0);
    }
    rc = finish_launch(b, e, "pp_dense");
    if (rc) return rc;
    CU(cudaMemcpy(out, b->d_out, out_bytes, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int assist_gpu_batch_get_state(assist_gpu_batch* b, double* state, double* acc, double* t, double* dt,
                                          double* dt_last_done, int* status) {
    if (!b) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    const size_t n = b->n;
    if (state) {
        soa_to_aos_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d.pos, b->n, b->K, 6, 0, 3, b->d_stage); AB_COUNT(1);
        soa_to_aos_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d.vel, b->n, b->K, 6, 3, 3, b->d_stage); AB_COUNT(1);
        CU(cudaMemcpy(state, b->d_stage, sizeof(double) * 6 * n * b->K, cudaMemcpyDeviceToHost));
    }
    if (acc) {
        soa_to_aos_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d.acc, b->n, b->K, 3, 0, 3, b->d_stage_prm); AB_COUNT(1);
        CU(cudaMemcpy(acc, b->d_stage_prm, sizeof(double) * 3 * n * b->K, cudaMemcpyDeviceToHost));
    }
    if (b->mode == ASSIST_GPU_PER_PARTICLE) {
        if (t) CU(cudaMemcpy(t, b->d.t, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (dt) CU(cudaMemcpy(dt, b->d.dt, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (dt_last_done) CU(cudaMemcpy(dt_last_done, b->d.dt_last, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (status) CU(cudaMemcpy(status, b->d.status, sizeof(int) * n, cudaMemcpyDeviceToHost));
    } else {
        AbShared sh;
        CU(cudaMemcpy(&sh, b->d.sh, sizeof(sh), cudaMemcpyDeviceToHost));
        if (t) *t = sh.t;
        if (dt) *dt = sh.dt;
        if (dt_last_done) *dt_last_done = sh.dt_last;
        if (status) *status = sh.status;
    }
    return 0;
}

extern "C" int ab_gpu_batch_update_params(assist_gpu_batch* b, const double* params) {
    if (!b) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    const size_t n = b->n;
    if (params) {
        CU(cudaMemcpy(b->d_stage_prm, params, sizeof(double) * 3 * n * b->K, cudaMemcpyHostToDevice));
        aos_to_soa_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d_stage_prm, b->n, b->K, 3, 0, 3, b->d.prm); AB_COUNT(1);
        CU(cudaGetLastError());
        b->d.has_params = 1;
    } else {
        b->d.has_params = 0;
    }
    return 0;
}

/* state at the start of the last completed step (the reference's ax->last_state) */
extern "C" int ab_gpu_batch_get_last_state(assist_gpu_batch* b, double* state, double* acc) {
    if (!b) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    const size_t n = b->n;
    if (state) {
        soa_to_aos_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d.ls_pos, b->n, b->K, 6, 0, 3, b->d_stage); AB_COUNT(1);
        soa_to_aos_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d.ls_vel, b->n, b->K, 6, 3, 3, b->d_stage); AB_COUNT(1);
        CU(cudaMemcpy(state, b->d_stage, sizeof(double) * 6 * n * b->K, cudaMemcpyDeviceToHost));
    }
    if (acc) {
        soa_to_aos_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d.ls_acc, b->n, b->K, 3, 0, 3, b->d_stage_prm); AB_COUNT(1);
        CU(cudaMemcpy(acc, b->d_stage_prm, sizeof(double) * 3 * n * b->K, cudaMemcpyDeviceToHost));
    }
    return 0;
}

/* b coefficients of the last completed step (dense output), br[7][n][K][3]: what a snapshot of the simulation keeps */
extern "C" int ab_gpu_batch_get_br(assist_gpu_batch* b, double* br) {
    if (!b || !br) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    const size_t n = b->n, per = 3 * n * b->K;
    for (int j = 0; j < 7; j++) {
        soa_to_aos_kernel<<<blocks_for((long long)n * b->K), 256>>>(b->d.br + (size_t)j * per, b->n, b->K, 3, 0, 3, b->d_stage_prm); AB_COUNT(1);
        CU(cudaMemcpy(br + (size_t)j * per, b->d_stage_prm, sizeof(double) * per, cudaMemcpyDeviceToHost));
    }
    return 0;
}

/* assist_interpolate_simulation (reference src/assist.c:682-752): positions and velocities at fraction h of the step
 * that starts at (x0, v0, a0) with the b coefficients br[7][m], m = 3 N components.  One thread per component, the
 * reference's expressions operation by operation (round-to-nearest intrinsics: no contraction whatever the build). */
__global__ void interpolate_simulation_kernel(const double* __restrict__ x0, const double* __restrict__ v0, const double* __restrict__ a0,
                                              const double* __restrict__ br, int m, double dt_last_done, double h,
                                              double* __restrict__ pos, double* __restrict__ vel) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define DIV(a, b) __ddiv_rn((a), (b))
    double s[9];
    s[0] = MUL(dt_last_done, h);
    s[1] = DIV(MUL(s[0], s[0]), 2.);
    s[2] = DIV(MUL(s[1], h), 3.);
    s[3] = DIV(MUL(s[2], h), 2.);
    s[4] = DIV(MUL(MUL(3., s[3]), h), 5.);
    s[5] = DIV(MUL(MUL(2., s[4]), h), 3.);
    s[6] = DIV(MUL(MUL(5., s[5]), h), 7.);
    s[7] = DIV(MUL(MUL(3., s[6]), h), 4.);
    s[8] = DIV(MUL(MUL(7., s[7]), h), 9.);
    const double b0 = br[k], b1 = br[m + k], b2 = br[2 * m + k], b3 = br[3 * m + k], b4 = br[4 * m + k], b5 = br[5 * m + k], b6 = br[6 * m + k];
    double sum = MUL(s[8], b6);
    sum = ADD(sum, MUL(s[7], b5)); sum = ADD(sum, MUL(s[6], b4)); sum = ADD(sum, MUL(s[5], b3)); sum = ADD(sum, MUL(s[4], b2));
    sum = ADD(sum, MUL(s[3], b1)); sum = ADD(sum, MUL(s[2], b0)); sum = ADD(sum, MUL(s[1], a0[k])); sum = ADD(sum, MUL(s[0], v0[k]));
    pos[k] = ADD(x0[k], sum);
    s[0] = MUL(dt_last_done, h);
    s[1] = DIV(MUL(s[0], h), 2.);
    s[2] = DIV(MUL(MUL(2., s[1]), h), 3.);
    s[3] = DIV(MUL(MUL(3., s[2]), h), 4.);
    s[4] = DIV(MUL(MUL(4., s[3]), h), 5.);
    s[5] = DIV(MUL(MUL(5., s[4]), h), 6.);
    s[6] = DIV(MUL(MUL(6., s[5]), h), 7.);
    s[7] = DIV(MUL(MUL(7., s[6]), h), 8.);
    double v = ADD(v0[k], MUL(s[7], b6));
    v = ADD(v, MUL(s[6], b5)); v = ADD(v, MUL(s[5], b4)); v = ADD(v, MUL(s[4], b3)); v = ADD(v, MUL(s[3], b2));
    v = ADD(v, MUL(s[2], b1)); v = ADD(v, MUL(s[1], b0)); v = ADD(v, MUL(s[0], a0[k]));
    vel[k] = v;
#undef MUL
#undef ADD
#undef DIV
}

/* host arrays in, host arrays out: x0, v0, a0 [m], br [7][m], pos, vel [m]; *dt_step = dt_last_done * h as the device
 * formed it (the reference advances sim1->t by it) */
extern "C" int assist_gpu_interpolate_simulation(int m, const double* x0, const double* v0, const double* a0, const double* br,
                                                 double dt_last_done, double h, double* pos, double* vel) {
    if (m < 1 || !x0 || !v0 || !a0 || !br || !pos || !vel) return set_err(ASSIST_GPU_ERR_ARG, "interpolate_simulation: bad argument");
    int dev = 0;
    int rc = ensure_device(&dev);
    if (rc) return rc;
    double* d = nullptr;
    const size_t M = (size_t)m;
    CU(cudaMalloc((void**)&d, sizeof(double) * 12 * M));
    cudaError_t e = cudaMemcpy(d, x0, sizeof(double) * M, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + M, v0, sizeof(double) * M, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + 2 * M, a0, sizeof(double) * M, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + 3 * M, br, sizeof(double) * 7 * M, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        interpolate_simulation_kernel<<<(m + 127) / 128, 128>>>(d, d + M, d + 2 * M, d + 3 * M, m, dt_last_done, h, d + 10 * M, d + 11 * M); AB_COUNT(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(pos, d + 10 * M, sizeof(double) * M, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(vel, d + 11 * M, sizeof(double) * M, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "interpolate_simulation: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int assist_gpu_batch_interpolate(assist_gpu_batch* b, double h, double* out) {
    if (!b || !out) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    if (b->mode != ASSIST_GPU_SHARED_STEP) return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "interpolate needs a shared-step batch");
    CU(cudaSetDevice(b->device));
    AbShared sh;
    CU(cudaMemcpy(&sh, b->d.sh, sizeof(sh), cudaMemcpyDeviceToHost));
    cudaError_t e = (b->opt.math == ASSIST_GPU_MATH_FAST) ? ab_launch_sh_interpolate_fast(b->d, sh.dt_last, h, b->d_stage, 0)
                                                          : ab_launch_sh_interpolate_strict(b->d, sh.dt_last, h, b->d_stage, 0);
    AB_COUNT(1);
    if (e != cudaSuccess) return set_err(ASSIST_GPU_ERR_CUDA, "interpolate: %s", cudaGetErrorString(e));
    CU(cudaMemcpy(out, b->d_stage, sizeof(double) * 6 * (size_t)b->n * b->K, cudaMemcpyDeviceToHost));
    return 0;
}

__global__ void sum_counters_kernel(const unsigned long long* __restrict__ c, long long n, unsigned long long* out) {
    unsigned long long s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += c[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

extern "C" int assist_gpu_batch_get_stats(assist_gpu_batch* b, struct assist_gpu_stats* stats) {
    if (!b || !stats) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    if (b->mode == ASSIST_GPU_PER_PARTICLE) {
        unsigned long long* d_sum = nullptr;
        unsigned long long h[4] = {0, 0, 0, 0};
        CU(cudaMalloc((void**)&d_sum, sizeof(h)));
        CU(cudaMemset(d_sum, 0, sizeof(h)));
        const unsigned long long* src[4] = {b->d.steps, b->d.rejected, b->d.iters, b->d.evals};
        for (int q = 0; q < 4; q++) sum_counters_kernel<<<256, 256>>>(src[q], b->n, d_sum + q);
        AB_COUNT(4);
        AB_COUNT(4);
        CU(cudaMemcpy(h, d_sum, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(d_sum);
        b->stats.steps = h[0]; b->stats.steps_rejected = h[1]; b->stats.pc_iterations = h[2]; b->stats.force_evals = h[3];
    } else {
        AbShared sh;
        CU(cudaMemcpy(&sh, b->d.sh, sizeof(sh), cudaMemcpyDeviceToHost));
        b->stats.steps = sh.steps; b->stats.steps_rejected = sh.rejected; b->stats.pc_iterations = sh.iters;
        b->stats.force_evals = sh.evals * (unsigned long long)b->n;
    }
    *stats = b->stats;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* FP64 peak micro-benchmark (roofline denominator)                         */
/* ------------------------------------------------------------------------ */

__global__ void dfma_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" double assist_gpu_measure_fp64_peak(int iters) {
    int dev;
    if (ensure_device(&dev)) return -1.0;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256, grid = sms * 8;
    double* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(double) * (size_t)grid * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_peak_kernel<<<grid, threads>>>(d, 16); AB_COUNT(1);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, 0);
        dfma_peak_kernel<<<grid, threads>>>(d, iters); AB_COUNT(1);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * (double)grid * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

/* ------------------------------------------------------------------------ */
/* pinned host buffers (for callers that want asynchronous-speed H2D/D2H)    */
/* ------------------------------------------------------------------------ */

extern "C" void* assist_gpu_host_alloc(size_t bytes) {
    int dev;
    if (ensure_device(&dev)) return NULL;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        set_err(ASSIST_GPU_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes);
        return NULL;
    }
    return p;
}

extern "C" void assist_gpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

/* Per-system counters of a per-particle batch (each array n_sys long; any may be NULL). */
extern "C" int assist_gpu_batch_get_counters(assist_gpu_batch* b, unsigned long long* steps, unsigned long long* rejected,
                                             unsigned long long* iters, unsigned long long* evals) {
    if (!b) return set_err(ASSIST_GPU_ERR_ARG, "NULL argument");
    if (b->mode != ASSIST_GPU_PER_PARTICLE) return set_err(ASSIST_GPU_ERR_UNSUPPORTED, "per-system counters need a per-particle batch");
    CU(cudaSetDevice(b->device));
    const size_t bytes = sizeof(unsigned long long) * (size_t)b->n;
    if (steps) CU(cudaMemcpy(steps, b->d.steps, bytes, cudaMemcpyDeviceToHost));
    if (rejected) CU(cudaMemcpy(rejected, b->d.rejected, bytes, cudaMemcpyDeviceToHost));
    if (iters) CU(cudaMemcpy(iters, b->d.iters, bytes, cudaMemcpyDeviceToHost));
    if (evals) CU(cudaMemcpy(evals, b->d.evals, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
