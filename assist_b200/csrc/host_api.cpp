/*
 * host_api.cpp -- the ASSIST host API of include/assist.h ("front door").
 *
 * Same entry points, argument meaning and error behaviour as the reference
 * (reference src/assist.c); what differs is what happens behind them: file
 * providers only parse, and every evaluation / integration is a CUDA launch.
 */
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "assist.h"
#include "assist_ephem_files.h"
#include "spk.h"
#include "ascii_ephem.h"
#include "assist_gpu.h"
#include "host_internal.h"

static_assert(sizeof(struct assist_ephem) == 208, "struct assist_ephem is ABI (reference assist/ephem.py:95-120)");
static_assert(sizeof(struct assist_extras) == 112, "struct assist_extras is ABI (reference assist/extras.py:84-100)");
static_assert(sizeof(struct reb_particle) == 128, "struct reb_particle must be REBOUND's 128-byte record");

#define AB_STR2(s) #s
#define AB_STR(s) AB_STR2(s)
#ifndef ASSISTGITHASH
#define ASSISTGITHASH notavailable0000000000000000000000000001
#endif

extern "C" {

const char* assist_build_str = __DATE__ " " __TIME__;
const char* assist_version_str = "1.2.0";                 /* API level of the reference this library mirrors */
const char* assist_githash_str = AB_STR(ASSISTGITHASH);

/* indices follow enum ASSIST_STATUS (reference src/assist.c:52-59); index 6 is ours */
const char* assist_error_messages[] = {
    "No error has occured.",
    "The JPL planet ephemeris file has not been found.",
    "The JPL asteroid ephemeris file has not been found. Asteroid forces have been disabled.",
    "The requested asteroid ID has not been found.",
    "The requested planet ID has not been found.",
    "The requested time is outside the coverage provided by the ephemeris file.",
    "No usable CUDA device: assist-b200 evaluates everything on the GPU and has no CPU path.",
    "The integration cannot advance: the timestep is zero or the particle exceeded its step budget.",
};
const int assist_error_messages_N = 8;

/* ---- format detection and discovery (reference src/assist.c:66-154) ------ */

int assist_detect_ascii_bin_signature(int fd) {
    /* three 6-character constant names live at 0x00FC; at least two must look like names */
    char names[18];
    if (pread(fd, names, sizeof(names), 0x00FC) != (ssize_t)sizeof(names)) return 0;
    int plausible = 0;
    for (int i = 0; i < 3; i++) {
        const unsigned char* nm = (const unsigned char*)&names[6 * i];
        const bool starts_alpha = (nm[0] >= 'A' && nm[0] <= 'Z') || (nm[0] >= 'a' && nm[0] <= 'z');
        if (!starts_alpha) continue;
        int alnum = 0;
        bool ok = true;
        for (int j = 0; j < 6; j++) {
            const unsigned char c = nm[j];
            if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9')) alnum++;
            else if (c != ' ' && c != '\0') { ok = false; break; }
        }
        if (ok && alnum >= 2) plausible++;
    }
    return plausible >= 2;
}

ephemeris_file_format_t assist_detect_ephemeris_file_format(int fd) {
    char magic[8];
    const ssize_t got = pread(fd, magic, sizeof(magic), 0);
    if (got <= 0) return FILE_FORMAT_UNKNOWN;
    if (got >= 8 && memcmp(magic, "DAF/SPK ", 8) == 0) { lseek(fd, 0, SEEK_SET); return FILE_FORMAT_VALID_BSP; }
    if (assist_detect_ascii_bin_signature(fd)) { lseek(fd, 0, SEEK_SET); return FILE_FORMAT_ASCII_BIN; }
    return FILE_FORMAT_UNKNOWN;
}

int assist_discover_planets_path(char* out_path, size_t out_path_size, const char* assist_dir) {
    if (out_path == NULL || out_path_size == 0 || assist_dir == NULL) return 0;
    static const char* candidates[] = {"/data/de441.bsp", "/data/de440.bsp", "/data/linux_m13000p17000.441",
                                       "/data/linux_p1550p2650.440"};
    for (int c = 0; c < 4; c++) {
        snprintf(out_path, out_path_size, "%s%s", assist_dir, candidates[c]);
        if (access(out_path, R_OK) == 0) return 1;
    }
    return 0;
}

/* ---- ephemeris life cycle (reference src/assist.c:188-377) ---------------- */

int assist_ephem_init(struct assist_ephem* ephem, char* user_planets_path, char* user_asteroids_path) {
    ephem->jd_ref = 2451545.0;
    ephem->spk_planets = NULL;
    ephem->spk_asteroids = NULL;
    ephem->ascii_planets = NULL;
    const char* base = getenv("ASSIST_DIR");
    char planets_path[1024], asteroids_path[1024];

    if (user_planets_path == NULL) {
        if (base == NULL) return ASSIST_ERROR_EPHEM_FILE;
        if (!assist_discover_planets_path(planets_path, sizeof(planets_path), base))
            snprintf(planets_path, sizeof(planets_path), "%s/data/de441.bsp", base);
    } else {
        snprintf(planets_path, sizeof(planets_path), "%s", user_planets_path);
    }
    ephemeris_file_format_t fmt = FILE_FORMAT_UNKNOWN;
    {
        const int fd = open(planets_path, O_RDONLY);
        if (fd >= 0) { fmt = assist_detect_ephemeris_file_format(fd); close(fd); }
    }

    bool have_ast_path = true;
    if (user_asteroids_path == NULL) {
        if (base == NULL) have_ast_path = false;
        else snprintf(asteroids_path, sizeof(asteroids_path), "%s/data/sb441-n16.bsp", base);
    } else {
        snprintf(asteroids_path, sizeof(asteroids_path), "%s", user_asteroids_path);
    }
    bool ast_missing = !have_ast_path;
    if (have_ast_path) {
        ephem->spk_asteroids = assist_spk_init(asteroids_path);
        if (ephem->spk_asteroids == NULL) ast_missing = true;
    }

    if (fmt == FILE_FORMAT_VALID_BSP) {
        ephem->spk_planets = assist_spk_init(planets_path);
        if (ephem->spk_planets == NULL) return ASSIST_ERROR_EPHEM_FILE;
        ephem->planets_source = FILE_FORMAT_VALID_BSP;
        static const int naif_by_assist[] = {10, 1, 2, 399, 301, 4, 5, 6, 7, 8, 9};
        for (int k = 0; k < ASSIST_BODY_NPLANETS; k++) {
            struct spk_target* t = assist_spk_find_target(ephem->spk_planets, naif_by_assist[k]);
            ephem->spk_target_index[k] = t ? (int)(t - ephem->spk_planets->targets) : -1;
        }
        struct spk_target* emb = assist_spk_find_target(ephem->spk_planets, 3);
        ephem->spk_emb_index = emb ? (int)(emb - ephem->spk_planets->targets) : -1;
        struct spk_constants_and_masses data = assist_load_spk_constants_and_masses(planets_path);
        assist_apply_spk_constants(ephem, &data);
        assist_spk_join_masses(ephem->spk_planets, &data.masses, ephem->EMRAT);
        if (ephem->spk_asteroids) assist_spk_join_masses(ephem->spk_asteroids, &data.masses, ephem->EMRAT);
        assist_free_spk_constants_and_masses(&data);
        ephem->planets_calc = assist_spk_calc_planets_by_assist;
    } else if (fmt == FILE_FORMAT_ASCII_BIN) {
        ephem->ascii_planets = assist_ascii_init(planets_path);
        if (ephem->ascii_planets == NULL) return ASSIST_ERROR_EPHEM_FILE;
        ephem->planets_source = FILE_FORMAT_ASCII_BIN;
        const struct ascii_s* a = ephem->ascii_planets;
        ephem->J2E = a->J2E; ephem->J3E = a->J3E; ephem->J4E = a->J4E; ephem->J2SUN = a->J2SUN;
        ephem->AU = a->AU; ephem->RE = a->RE; ephem->CLIGHT = a->CLIGHT; ephem->ASUN = a->ASUN;
        ephem->EMRAT = a->cem;
        ephem->Re_eq = ephem->RE / ephem->AU;
        ephem->Rs_eq = ephem->ASUN / ephem->AU;
        ephem->c_AU_per_day = (ephem->CLIGHT / ephem->AU) * 86400.0;
        ephem->c_squared = ephem->c_AU_per_day * ephem->c_AU_per_day;
        ephem->over_c_squared = 1.0 / ephem->c_squared;
        ephem->planets_calc = assist_ascii_calc_from_ephem;
        if (ephem->spk_asteroids) {
            /* asteroid GMs come from the MAxxxx constants of the planets file (reference src/assist.c:312-333) */
            for (int n = 0; n < ephem->spk_asteroids->num; n++) {
                struct spk_target* tg = &ephem->spk_asteroids->targets[n];
                if (tg->code >= 2000000) {
                    char key[16];
                    snprintf(key, sizeof(key), "MA%04d", tg->code - 2000000);
                    double gm = 0.0;
                    if (assist_ascii_find_constant(a, key, &gm)) tg->mass = gm;
                } else if (tg->code == 399) tg->mass = a->mass[ASSIST_BODY_EARTH];
                else if (tg->code == 301) tg->mass = a->mass[ASSIST_BODY_MOON];
                else if (tg->code == 10) tg->mass = a->mass[ASSIST_BODY_SUN];
            }
        }
    } else {
        fprintf(stderr, "(ASSIST) Error: Failed to initialize planets ephemeris from '%s'.\n", planets_path);
        fprintf(stderr, "(ASSIST) Supported: NAIF SPK kernels (.bsp, e.g. de440.bsp) and JPL binary ephemerides (.440/.441).\n");
        return ASSIST_ERROR_EPHEM_FILE;
    }
    if (ast_missing) fprintf(stderr, "(ASSIST) %s\n", assist_error_messages[ASSIST_ERROR_AST_FILE]);
    return ASSIST_SUCCESS;
}

void assist_ephem_free_pointers(struct assist_ephem* ephem) {
    if (ephem->spk_planets) { assist_spk_free(ephem->spk_planets); ephem->spk_planets = NULL; }
    if (ephem->spk_asteroids) { assist_spk_free(ephem->spk_asteroids); ephem->spk_asteroids = NULL; }
    if (ephem->ascii_planets) { assist_ascii_free(ephem->ascii_planets); ephem->ascii_planets = NULL; }
}

void assist_ephem_free(struct assist_ephem* ephem) {
    if (!ephem) return;
    assist_ephem_free_pointers(ephem);
    free(ephem);
}

struct assist_ephem* assist_ephem_create(char* user_planets_path, char* user_asteroids_path) {
    struct assist_ephem* ephem = (struct assist_ephem*)calloc(1, sizeof(struct assist_ephem));
    const int error = assist_ephem_init(ephem, user_planets_path, user_asteroids_path);
    if (error != ASSIST_SUCCESS) {
        fprintf(stderr, "(ASSIST) An error occured while trying to initialize the ephemeris structure.\n");
        fprintf(stderr, "(ASSIST) %s\n", assist_error_messages[error]);
        assist_ephem_free(ephem);
        return NULL;
    }
    return ephem;
}

void assist_ephem_time_bounds(const struct assist_ephem* ephem, double* t_beg, double* t_end) {
    if (ephem == NULL) return;
    double beg = -INFINITY, end = INFINITY;
    const struct spk_s* files[2] = {ephem->spk_planets, ephem->spk_asteroids};
    for (int f = 0; f < 2; f++) {
        if (!files[f]) continue;
        for (int i = 0; i < files[f]->num; i++) {
            if (files[f]->targets[i].beg > beg) beg = files[f]->targets[i].beg;
            if (files[f]->targets[i].end < end) end = files[f]->targets[i].end;
        }
    }
    if (ephem->spk_planets == NULL && ephem->ascii_planets != NULL) {
        if (ephem->ascii_planets->beg > beg) beg = ephem->ascii_planets->beg;
        if (ephem->ascii_planets->end < end) end = ephem->ascii_planets->end;
    }
    if (t_beg) *t_beg = beg - ephem->jd_ref;
    if (t_end) *t_end = end - ephem->jd_ref;
}

/* ---- ephemeris queries (GPU-backed) --------------------------------------- */

static int map_gpu_error(int rc) {
    if (rc >= 0) return rc;
    fprintf(stderr, "(ASSIST) %s\n", assist_gpu_last_error());
    return ASSIST_ERROR_GPU;
}

/* All bodies at time t in one launch; out has nbodies*10 doubles. */
static int eval_all_bodies(const struct assist_ephem* ephem, double t, std::vector<double>& out, std::vector<int>& st) {
    const int nb = assist_gpu_ephem_nbodies(ephem);
    out.resize((size_t)nb * 10);
    st.resize(nb);
    return assist_gpu_ephem_eval(ephem, ASSIST_GPU_MATH_STRICT, &t, 1, out.data(), st.data());
}

int assist_all_ephem(const struct assist_ephem* ephem, struct assist_ephem_cache* cache, const int i, const double t,
                     double* const GM, double* const x, double* const y, double* const z,
                     double* const vx, double* const vy, double* const vz,
                     double* const ax, double* const ay, double* const az) {
    const int nb = assist_gpu_ephem_nbodies(ephem);
    if (i < 0) return ASSIST_ERROR_NEPHEM;
    if (i >= nb) return (i < ASSIST_BODY_NPLANETS) ? ASSIST_ERROR_NEPHEM : (ephem->spk_asteroids ? ASSIST_ERROR_NAST : ASSIST_ERROR_AST_FILE);
    struct assist_cache_item item;
    bool hit = false;
    if (cache) {
        /* seven slots per body keyed on the exact time (reference src/forces.c:180-197) */
        for (int s = 0; s < 7; s++) if (cache->t[7 * i + s] == t) { item = cache->items[7 * i + s]; hit = true; break; }
    }
    if (!hit) {
        /* per-thread buffers that keep their capacity: a cache miss does not allocate */
        static thread_local std::vector<double> out;
        static thread_local std::vector<int> st;
        const int rc = eval_all_bodies(ephem, t, out, st);
        if (rc) return map_gpu_error(rc);
        if (st[i] != ASSIST_SUCCESS) return st[i];
        if (cache) {
            /* one launch produced every body: refresh the slot of each (oldest in the direction of travel) */
            for (int b = 0; b < nb; b++) {
                if (st[b] != ASSIST_SUCCESS) continue;
                double* ct = cache->t + 7 * b;
                int os = 0;
                for (int s = 1; s < 7; s++) if ((cache->dt_sign > 0) ? (ct[s] < ct[os]) : (ct[s] > ct[os])) os = s;
                ct[os] = t;
                memcpy(&cache->items[7 * b + os], &out[(size_t)b * 10], sizeof(struct assist_cache_item));
            }
        }
        memcpy(&item, &out[(size_t)i * 10], sizeof(item));
    }
    *GM = item.GM; *x = item.x; *y = item.y; *z = item.z;
    *vx = item.vx; *vy = item.vy; *vz = item.vz; *ax = item.ax; *ay = item.ay; *az = item.az;
    return ASSIST_SUCCESS;
}

struct reb_particle assist_get_particle_with_error(const struct assist_ephem* ephem, const int particle_id, const double t, int* error) {
    struct reb_particle p;
    memset(&p, 0, sizeof(p));
    double GM = 0;
    const int flag = assist_all_ephem(ephem, NULL, particle_id, t, &GM, &p.x, &p.y, &p.z, &p.vx, &p.vy, &p.vz, &p.ax, &p.ay, &p.az);
    *error = flag;
    p.m = GM;    /* GM, not mass (reference src/assist.c:509) */
    return p;
}

struct reb_particle assist_get_particle(const struct assist_ephem* ephem, const int particle_id, const double t) {
    int error = 0;
    struct reb_particle p = assist_get_particle_with_error(ephem, particle_id, t, &error);
    if (error != ASSIST_SUCCESS) {
        fprintf(stderr, "(ASSIST) An error occured while trying to initialize particle from ephemeris data.\n");
        fprintf(stderr, "(ASSIST) %s\n", assist_error_messages[error < assist_error_messages_N ? error : 0]);
    }
    return p;
}

/* planets_calc providers: same signature as the reference's function pointer */
static enum ASSIST_STATUS planets_calc_gpu(const struct assist_ephem* ephem, double jd_ref, double jd_rel, int body,
                                           double* GM, double* x, double* y, double* z, double* vx, double* vy, double* vz,
                                           double* ax, double* ay, double* az) {
    if (body < 0 || body >= ASSIST_BODY_NPLANETS) return ASSIST_ERROR_NEPHEM;
    /* the kernels take times relative to ephem->jd_ref */
    const double t = (jd_ref - ephem->jd_ref) + jd_rel;
    return (enum ASSIST_STATUS)assist_all_ephem(ephem, NULL, body, t, GM, x, y, z, vx, vy, vz, ax, ay, az);
}

enum ASSIST_STATUS assist_spk_calc_planets_by_assist(const struct assist_ephem* ephem, double jd_ref, double jd_rel, int assist_body,
                                                     double* GM, double* x, double* y, double* z, double* vx, double* vy, double* vz,
                                                     double* ax, double* ay, double* az) {
    if (!ephem || !ephem->spk_planets) return ASSIST_ERROR_NEPHEM;
    return planets_calc_gpu(ephem, jd_ref, jd_rel, assist_body, GM, x, y, z, vx, vy, vz, ax, ay, az);
}

enum ASSIST_STATUS assist_ascii_calc_from_ephem(const struct assist_ephem* ephem, double jd_ref, double jd_rel, int body,
                                                double* const GM, double* const x, double* const y, double* const z,
                                                double* const vx, double* const vy, double* const vz,
                                                double* const ax, double* const ay, double* const az) {
    if (!ephem || !ephem->ascii_planets) return ASSIST_ERROR_EPHEM_FILE;
    return planets_calc_gpu(ephem, jd_ref, jd_rel, body, GM, x, y, z, vx, vy, vz, ax, ay, az);
}

/* ---- the evaluator entry points of the reference's spk.h / ascii_ephem.h (src/spk.h:113-122, src/ascii_ephem.h:15-22).
 * Same names, arguments and status codes; every value is computed by a CUDA launch. */

static enum ASSIST_STATUS gpu_status(int rc) { return (enum ASSIST_STATUS)map_gpu_error(rc); }

struct mpos_s assist_spk_target_pos(const struct spk_s* pl, const struct spk_target* target, double jd_ref, double jd_rel) {
    struct mpos_s pos;
    const double nan = NAN;
    for (int i = 0; i < 3; i++) pos.u[i] = pos.v[i] = pos.w[i] = nan;
    if (!pl || !target) return pos;
    double out[9];
    if (ab_gpu_spk_target_eval((struct spk_s*)pl, (int)(target - pl->targets), -1, jd_ref, jd_rel, 0, NULL, out)) {
        fprintf(stderr, "(ASSIST) %s\n", assist_gpu_last_error());
        return pos;
    }
    for (int i = 0; i < 3; i++) { pos.u[i] = out[i]; pos.v[i] = out[3 + i]; pos.w[i] = out[6 + i]; }
    return pos;
}

enum ASSIST_STATUS assist_spk_calc(const struct spk_s* pl, double jd_ref, double jd_rel, int m, double* GM,
                                   double* out_x, double* out_y, double* out_z) {
    if (pl == NULL) return ASSIST_ERROR_AST_FILE;
    if (m < 0 || m >= pl->num) return ASSIST_ERROR_NAST;
    const struct spk_target* target = &pl->targets[m];
    if (jd_ref + jd_rel < target->beg || jd_ref + jd_rel > target->end) return ASSIST_ERROR_COVERAGE;
    *GM = target->mass;
    double out[9];
    const int rc = ab_gpu_spk_target_eval((struct spk_s*)pl, m, -1, jd_ref, jd_rel, 2, NULL, out);
    if (rc) return gpu_status(rc);
    *out_x = out[0]; *out_y = out[1]; *out_z = out[2];
    return ASSIST_SUCCESS;
}

enum ASSIST_STATUS assist_spk_calc_planets(const struct assist_ephem* ephem, double jd_ref, double jd_rel, int code, double* GM,
                                           double* out_x, double* out_y, double* out_z, double* out_vx, double* out_vy, double* out_vz,
                                           double* out_ax, double* out_ay, double* out_az) {
    if (!ephem) return ASSIST_ERROR_NEPHEM;
    struct spk_s* pl = ephem->spk_planets;
    if (!pl) return ASSIST_ERROR_NEPHEM;
    const struct spk_target* target = assist_spk_find_target(pl, code);
    if (target == NULL) return ASSIST_ERROR_NEPHEM;
    if (jd_ref + jd_rel < target->beg || jd_ref + jd_rel > target->end) return ASSIST_ERROR_COVERAGE;
    *GM = target->mass;
    int emb_index = -1;
    if (code == 301 || code == 399) {       /* relative to the Earth-Moon barycentre (reference src/spk.c:572-587) */
        const struct spk_target* emb = NULL;
        if (ephem->spk_emb_index >= 0 && ephem->spk_emb_index < pl->num && pl->targets[ephem->spk_emb_index].code == 3) emb = &pl->targets[ephem->spk_emb_index];
        else emb = assist_spk_find_target(pl, 3);
        if (!emb) return ASSIST_ERROR_NEPHEM;
        emb_index = (int)(emb - pl->targets);
    }
    const double au = ephem->AU, seconds_per_day = 86400.;
    const double ud[3] = {au, au / seconds_per_day, au / (seconds_per_day * seconds_per_day)};
    double out[9];
    const int rc = ab_gpu_spk_target_eval(pl, (int)(target - pl->targets), emb_index, jd_ref, jd_rel, 1, ud, out);
    if (rc) return gpu_status(rc);
    *out_x = out[0]; *out_y = out[1]; *out_z = out[2];
    *out_vx = out[3]; *out_vy = out[4]; *out_vz = out[5];
    *out_ax = out[6]; *out_ay = out[7]; *out_az = out[8];
    return ASSIST_SUCCESS;
}

enum ASSIST_STATUS assist_ascii_calc(struct ascii_s* pl, double jd_ref, double jd_rel, int body, double* const GM,
                                     double* const x, double* const y, double* const z,
                                     double* const vx, double* const vy, double* const vz,
                                     double* const ax, double* const ay, double* const az) {
    if (!pl) return ASSIST_ERROR_EPHEM_FILE;
    /* a bare file handle: wrap it the way assist_ephem_init does for the planets provider */
    struct assist_ephem tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.ascii_planets = pl;
    tmp.planets_source = FILE_FORMAT_ASCII_BIN;
    tmp.jd_ref = 0.0;
    tmp.AU = pl->AU; tmp.EMRAT = pl->cem;
    for (int k = 0; k < ASSIST_BODY_NPLANETS; k++) tmp.spk_target_index[k] = -1;
    tmp.spk_emb_index = -1;
    return planets_calc_gpu(&tmp, jd_ref, jd_rel, body, GM, x, y, z, vx, vy, vz, ax, ay, az);
}

void assist_ascii_work(double* P, int ncm, int ncf, int niv, double t0, double t1, double* u, double* v, double* w) {
    double out[9];
    if (ab_gpu_ascii_work(P, ncm, ncf, niv, t0, t1, out)) {
        fprintf(stderr, "(ASSIST) %s\n", assist_gpu_last_error());
        for (int m = 0; m < ncm && m < 3; m++) u[m] = v[m] = w[m] = NAN;
        return;
    }
    for (int m = 0; m < ncm; m++) { u[m] = out[m]; v[m] = out[ncm + m]; w[m] = out[2 * ncm + m]; }
}

/* ---- attach / detach (reference src/assist.c:379-501) --------------------- */

static void assist_extras_cleanup(struct reb_simulation* sim) {
    struct assist_extras* assist = (struct assist_extras*)sim->extras;
    if (assist) assist->sim = NULL;
}

void ab_host_pre_timestep_marker(struct reb_simulation* r) { (void)r; }

void assist_init(struct assist_extras* assist, struct reb_simulation* sim, struct assist_ephem* ephem) {
    assist->sim = sim;
    assist->ephem_cache = (struct assist_ephem_cache*)calloc(1, sizeof(struct assist_ephem_cache));
    int N_total = ASSIST_BODY_NPLANETS;
    if (ephem->spk_asteroids) N_total += ephem->spk_asteroids->num;
    assist->gr_eih_sources = 1;
    assist->ephem_cache->items = (struct assist_cache_item*)calloc((size_t)N_total * 7, sizeof(struct assist_cache_item));
    assist->ephem_cache->t = (double*)malloc((size_t)N_total * 7 * sizeof(double));
    for (int i = 0; i < 7 * N_total; i++) assist->ephem_cache->t[i] = -1e306;
    assist->ephem_cache->dt_sign = 1.0;
    assist->ephem = ephem;
    assist->particle_params = NULL;
    assist->forces = ASSIST_FORCE_SUN | ASSIST_FORCE_PLANETS | ASSIST_FORCE_ASTEROIDS | ASSIST_FORCE_NON_GRAVITATIONAL |
                     ASSIST_FORCE_EARTH_HARMONICS | ASSIST_FORCE_SUN_HARMONICS | ASSIST_FORCE_GR_EIH;
    assist->last_state = NULL;
    assist->current_state = NULL;
    assist->alpha = 1.0; assist->nk = 0.0; assist->nm = 2.0; assist->nn = 5.093; assist->r0 = 1.0;

    sim->integrator = REB_INTEGRATOR_IAS15;
    sim->gravity = REB_GRAVITY_NONE;
    sim->extras = assist;
    sim->extras_cleanup = assist_extras_cleanup;
    sim->additional_forces = assist_additional_forces;
    sim->force_is_velocity_dependent = 1;
    sim->ri_ias15.adaptive_mode = 1;
}

struct assist_extras* assist_attach(struct reb_simulation* sim, struct assist_ephem* ephem) {
    if (sim == NULL) {
        fprintf(stderr, "(ASSIST) Error: Simulation pointer passed to assist_attach was NULL.\n");
        return NULL;
    }
    int should_free = 0;
    if (ephem == NULL) {
        ephem = assist_ephem_create(NULL, NULL);
        if (ephem == NULL) {
            fprintf(stderr, "(ASSIST) Error: Ephemeris pointer passed to assist_attach was NULL. Initialization with default path failed.\n");
            return NULL;
        }
        should_free = 1;
    }
    struct assist_extras* assist = (struct assist_extras*)calloc(1, sizeof(*assist));
    assist_init(assist, sim, ephem);
    assist->extras_should_free_ephem = should_free;
    return assist;
}

void assist_detach(struct reb_simulation* sim, struct assist_extras* assist) {
    if (assist->sim) {
        sim->extras = NULL;
        sim->extras_cleanup = NULL;
        sim->additional_forces = NULL;
        sim->pre_timestep_modifications = NULL;
        ab_host_drop_batch(sim);
    }
    assist->sim = NULL;
}

void assist_free_pointers(struct assist_extras* assist) {
    if (assist->sim) { assist_detach(assist->sim, assist); assist->sim = NULL; }
    free(assist->last_state); assist->last_state = NULL;
    free(assist->current_state); assist->current_state = NULL;
    if (assist->ephem_cache) {
        free(assist->ephem_cache->items);
        free(assist->ephem_cache->t);
        free(assist->ephem_cache);
        assist->ephem_cache = NULL;
    }
    if (assist->extras_should_free_ephem && assist->ephem) { assist_ephem_free(assist->ephem); assist->ephem = NULL; }
}

void assist_free(struct assist_extras* assist) {
    if (!assist) return;
    assist_free_pointers(assist);
    free(assist);
}

void assist_error(struct assist_extras* assist, const char* const msg) {
    if (assist->sim == NULL)
        fprintf(stderr, "(ASSIST) Error: A Simulation is no longer attached to the ASSIST extras instance. Most likely the Simulation has been freed.\n");
    else
        reb_simulation_error(assist->sim, msg);
}

/* ---- dense output (reference src/assist.c:635-680) ------------------------- */

static void swap_particles(struct reb_simulation* sim, struct assist_extras* ax) {
    struct reb_particle* p = sim->particles;
    sim->particles = ax->current_state;
    ax->current_state = p;
}

void assist_integrate_or_interpolate(struct assist_extras* ax, double t) {
    struct reb_simulation* sim = ax->sim;
    sim->pre_timestep_modifications = ab_host_pre_timestep_marker;
    sim->exact_finish_time = 0;

    if (ax->current_state == NULL) {
        ax->current_state = (struct reb_particle*)malloc(sizeof(struct reb_particle) * sim->N);
        ax->last_state = (struct reb_particle*)malloc(sizeof(struct reb_particle) * sim->N);
        memcpy(ax->current_state, sim->particles, sizeof(struct reb_particle) * sim->N);
        memcpy(ax->last_state, sim->particles, sizeof(struct reb_particle) * sim->N);
    } else {
        swap_particles(sim, ax);
    }

    const double dts = copysign(1., sim->dt_last_done);
    if (dts * (sim->t - sim->dt_last_done) > dts * t || dts * t > dts * sim->t || sim->dt_last_done == 0.0) {
        reb_simulation_integrate(sim, t);
    }

    const double h = 1.0 - (sim->t - t) / sim->dt_last_done;
    if (sim->status > 0) {
        printf("Error: simulation exited with status %d.\n", sim->status);
    } else if (sim->t - t == 0.) {
        memcpy(ax->current_state, sim->particles, sizeof(struct reb_particle) * sim->N);
    } else if (h < 0.0 || h >= 1.0 || !isnormal(h)) {
        printf("Error: cannot interpolate beyond timestep bounds (h=%e).\n", h);
    } else if (sim->b200_batch == NULL) {
        printf("Error: cannot interpolate before first timestep is complete (h=%e).\n", h);
    } else {
        ab_host_interpolate(sim, h, ax->current_state);
    }
    swap_particles(sim, ax);
}

/* reference src/assist.c:682-752: sim1 (a snapshot) moved to fraction h of the step that sim2 (the next snapshot) has
 * just completed: x0, v0 of sim1, a0, br and dt_last_done of sim2.  The polynomial is evaluated on the GPU
 * (assist_gpu_interpolate_simulation), operation by operation as the reference writes it. */
int assist_interpolate_simulation(struct reb_simulation* sim1, struct reb_simulation* sim2, double h) {
    if (!sim1 || !sim2 || sim1->N != sim2->N || sim1->N == 0) return 0;
    const struct reb_integrator_ias15* r1 = &sim1->ri_ias15;
    const struct reb_integrator_ias15* r2 = &sim2->ri_ias15;
    if (!r1->x0 || !r1->v0 || !r2->a0 || !r2->br.p0) {
        reb_simulation_error(sim1, "assist_interpolate_simulation: the simulations carry no IAS15 step data (not restored from a snapshot file).");
        return 0;
    }
    const int m = 3 * (int)sim1->N;
    std::vector<double> br((size_t)7 * m), pos(m), vel(m);
    const double* seven[7] = {r2->br.p0, r2->br.p1, r2->br.p2, r2->br.p3, r2->br.p4, r2->br.p5, r2->br.p6};
    for (int q = 0; q < 7; q++) memcpy(&br[(size_t)q * m], seven[q], sizeof(double) * m);
    if (assist_gpu_interpolate_simulation(m, r1->x0, r1->v0, r2->a0, br.data(), sim2->dt_last_done, h, pos.data(), vel.data())) {
        reb_simulation_error(sim1, assist_gpu_last_error());
        return 0;
    }
    for (unsigned int j = 0; j < sim1->N; j++) {
        struct reb_particle* p = &sim1->particles[j];
        p->x = pos[3 * j]; p->y = pos[3 * j + 1]; p->z = pos[3 * j + 2];
        p->vx = vel[3 * j]; p->vy = vel[3 * j + 1]; p->vz = vel[3 * j + 2];
    }
    sim1->t += sim2->dt_last_done * h;
    return 1;
}

/* reference src/assist.c:599-633 */
struct reb_simulation* assist_create_interpolated_simulation(struct reb_simulationarchive* sa, double t) {
    if (sa == NULL) return NULL;
    /* the first snapshot cannot be used: it precedes the first step (no accelerations, no b coefficients) */
    if (sa->nblobs < 2 || t <= sa->t[1]) {
        printf("Requested time outside range of SimulationArchive.\n");
        return NULL;
    }
    if (t >= sa->t[sa->nblobs - 1]) {
        printf("Requested time outside range of SimulationArchive.\n");
        return NULL;
    }
    long blob = 0;
    for (long i = 1; i < sa->nblobs; i++) {
        if (sa->t[i] >= t) { blob = i; break; }
    }
    enum reb_simulation_binary_error_codes warnings = REB_SIMULATION_BINARY_WARNING_NONE;
    struct reb_simulation* r2 = reb_simulation_create();
    reb_simulation_create_from_simulationarchive_with_messages(r2, sa, blob - 1, &warnings);
    struct reb_simulation* r3 = reb_simulation_create();
    reb_simulation_create_from_simulationarchive_with_messages(r3, sa, blob, &warnings);
    if (r2->messages_waiting || r3->messages_waiting) { reb_simulation_free(r2); reb_simulation_free(r3); return NULL; }
    const double h = (t - r2->t) / (r3->dt_last_done);
    const int ok = assist_interpolate_simulation(r2, r3, h);
    reb_simulation_free(r3);
    if (!ok) { reb_simulation_free(r2); return NULL; }
    return r2;
}

/* reference src/tools.c:35-70: a plain REBOUND simulation holding the ephemeris bodies at r->t as active particles
 * followed by r's own particles.  The body states come from the GPU ephemeris evaluation (assist_get_particle); the
 * returned simulation is data -- integrating it needs REBOUND's own N-body gravity, which this library does not
 * replace (reb_simulation_integrate on it fails with a message: no ASSIST extras attached).
 * As in the reference, all eleven bodies are added whatever merge_moon is (its filter `i != EARTH || i != MOON` is
 * always true) and merge_moon = 1 appends the Earth-Moon barycentre as a twelfth active particle. */
struct reb_simulation* assist_simulation_convert_to_rebound(const struct reb_simulation* r, const struct assist_ephem* ephem, int merge_moon) {
    if (!r || !ephem) return NULL;
    struct reb_simulation* r2 = reb_simulation_create();
    if (!r2) return NULL;
    r2->t = r->t;
    r2->dt = r->dt;
    r2->ri_ias15.epsilon = r->ri_ias15.epsilon;
    r2->ri_ias15.adaptive_mode = r->ri_ias15.adaptive_mode;
    for (int i = 0; i < 11; i++) {
        int error = 0;
        struct reb_particle p = assist_get_particle_with_error(ephem, i, r->t, &error);
        if (error != ASSIST_SUCCESS) fprintf(stderr, "(ASSIST) An error occured while trying to initialize particle from ephemeris data.\n");
        else reb_simulation_add(r2, p);
    }
    if (merge_moon) {
        int error = 0;
        struct reb_particle p1 = assist_get_particle_with_error(ephem, ASSIST_BODY_EARTH, r->t, &error);
        struct reb_particle p2 = assist_get_particle_with_error(ephem, ASSIST_BODY_MOON, r->t, &error);
        if (error != ASSIST_SUCCESS) fprintf(stderr, "(ASSIST) An error occured while trying to initialize particle from ephemeris data.\n");
        else reb_simulation_add(r2, reb_particle_com_of_pair(p1, p2));
    }
    r2->N_active = (int)r2->N;
    for (unsigned int i = 0; i < r->N; i++) reb_simulation_add(r2, r->particles[i]);
    return r2;
}

}  // extern "C"
