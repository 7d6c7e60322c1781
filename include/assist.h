/*
 * assist.h -- the ASSIST host API, as exported by assist-b200's libassist.
 *
 * Same symbol names, signatures, enum values and struct layouts as the reference
 * header (reference src/assist.h).  Layouts are ABI: the reference's Python
 * binding mirrors them field by field (reference assist/ephem.py:95-120,
 * assist/extras.py:84-100), so `struct assist_ephem` is 208 bytes and
 * `struct assist_extras` is 112 bytes here too (static-asserted in the library).
 * Behind these entry points every computation runs in CUDA kernels on the
 * current device; there is no CPU compute path (calls fail with an error message
 * and REB_STATUS_GENERIC_ERROR / a non-zero ASSIST_STATUS when no device exists).
 */
#ifndef _ASSIST_B200_ASSIST_H
#define _ASSIST_B200_ASSIST_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include "rebound.h"

#ifdef __cplusplus
extern "C" {
#endif

extern const char* assist_build_str;      /* reference src/assist.c:45 */
extern const char* assist_version_str;    /* reference src/assist.c:46 */
extern const char* assist_githash_str;    /* reference src/assist.c:47 */

/* reference src/assist.h:53-61 */
typedef enum {
    FILE_FORMAT_VALID_BSP = 0,
    FILE_FORMAT_ASCII_BIN = 2,
    FILE_FORMAT_UNKNOWN = 3
} ephemeris_file_format_t;
#define FILE_FORMAT_BINARY_LEGACY FILE_FORMAT_ASCII_BIN

int assist_detect_ascii_bin_signature(int fd);                          /* src/assist.h:65 */
ephemeris_file_format_t assist_detect_ephemeris_file_format(int fd);    /* src/assist.h:68 */
int assist_discover_planets_path(char* out_path, size_t out_path_size, const char* assist_dir); /* :72 */

/* reference src/assist.h:76-87 */
enum ASSIST_FORCES {
    ASSIST_FORCE_NONE               = 0,
    ASSIST_FORCE_SUN                = 0x01,
    ASSIST_FORCE_PLANETS            = 0x02,
    ASSIST_FORCE_ASTEROIDS          = 0x04,
    ASSIST_FORCE_NON_GRAVITATIONAL  = 0x08,
    ASSIST_FORCE_EARTH_HARMONICS    = 0x10,
    ASSIST_FORCE_SUN_HARMONICS      = 0x20,
    ASSIST_FORCE_GR_EIH             = 0x40,
    ASSIST_FORCE_GR_SIMPLE          = 0x80,
    ASSIST_FORCE_GR_POTENTIAL       = 0x100,
};

/* reference src/assist.h:90-98 */
enum ASSIST_STATUS {
    ASSIST_SUCCESS,
    ASSIST_ERROR_EPHEM_FILE,
    ASSIST_ERROR_AST_FILE,
    ASSIST_ERROR_NAST,
    ASSIST_ERROR_NEPHEM,
    ASSIST_ERROR_COVERAGE,
    ASSIST_ERROR_N,
};

extern const char* assist_error_messages[];
extern const int assist_error_messages_N;

/* reference src/assist.h:102-116 */
enum ASSIST_BODY {
    ASSIST_BODY_SUN = 0, ASSIST_BODY_MERCURY = 1, ASSIST_BODY_VENUS = 2, ASSIST_BODY_EARTH = 3,
    ASSIST_BODY_MOON = 4, ASSIST_BODY_MARS = 5, ASSIST_BODY_JUPITER = 6, ASSIST_BODY_SATURN = 7,
    ASSIST_BODY_URANUS = 8, ASSIST_BODY_NEPTUNE = 9, ASSIST_BODY_PLUTO = 10,
    ASSIST_BODY_NPLANETS = 11,
};

struct ascii_s;
struct spk_s;
struct spk_target;

/* reference src/assist.h:125-155 */
struct assist_ephem {
    double jd_ref;
    struct spk_s* spk_planets;
    struct spk_s* spk_asteroids;
    struct ascii_s* ascii_planets;
    int planets_source;
    enum ASSIST_STATUS (*planets_calc)(const struct assist_ephem*, double, double, int,
                                       double* const,
                                       double* const, double* const, double* const,
                                       double* const, double* const, double* const,
                                       double* const, double* const, double* const);
    int spk_target_index[ASSIST_BODY_NPLANETS];
    int spk_emb_index;
    double AU;
    double EMRAT;
    double J2E;
    double J3E;
    double J4E;
    double J2SUN;
    double RE;
    double CLIGHT;
    double ASUN;
    double Re_eq;
    double Rs_eq;
    double c_AU_per_day;
    double c_squared;
    double over_c_squared;
};

/* reference src/assist.h:157-175.  The reference memoises (body, time) lookups here;
 * on the GPU the memoisation lives in the kernels, so the host-side cache is only
 * allocated for layout compatibility. */
struct assist_cache_item {
    double GM;
    double x, y, z;
    double vx, vy, vz;
    double ax, ay, az;
};

struct assist_ephem_cache {
    double* t;
    double dt_sign;
    struct assist_cache_item* items;
};

/* reference src/assist.h:177-195 */
struct assist_extras {
    struct reb_simulation* sim;
    struct assist_ephem* ephem;
    struct assist_ephem_cache* ephem_cache;
    int extras_should_free_ephem;
    int geocentric;
    struct reb_particle* last_state;
    struct reb_particle* current_state;
    double* particle_params;
    int steps_done;
    int forces;
    int gr_eih_sources;
    double alpha;
    double nk;
    double nm;
    double nn;
    double r0;
};

struct assist_extras* assist_attach(struct reb_simulation* sim, struct assist_ephem* ephem);   /* :202 */
void assist_free(struct assist_extras* assist);                                                /* :209 */
void assist_ephem_free(struct assist_ephem* ephem);                                            /* :211 */
void assist_detach(struct reb_simulation* sim, struct assist_extras* assist);                  /* :218 */
void assist_error(struct assist_extras* assist, const char* const msg);                        /* :227 */

int assist_interpolate_simulation(struct reb_simulation* sim1, struct reb_simulation* sim2, double h);          /* :230 */
struct reb_simulation* assist_create_interpolated_simulation(struct reb_simulationarchive* sa, double t);      /* :231 */
void assist_integrate_or_interpolate(struct assist_extras* ax, double t);                                      /* :232 */

struct reb_particle assist_get_particle(const struct assist_ephem* ephem, const int particle_id, const double t);  /* :235 */
struct reb_particle assist_get_particle_with_error(const struct assist_ephem* ephem, const int particle_id, const double t, int* error); /* :245 */
void assist_ephem_time_bounds(const struct assist_ephem* ephem, double* t_beg, double* t_end);  /* :262 */

void assist_init(struct assist_extras* assist, struct reb_simulation* sim, struct assist_ephem* ephem); /* :265 */
void assist_free_pointers(struct assist_extras* assist);                                        /* :266 */
void assist_ephem_free_pointers(struct assist_ephem* ephem);                                    /* :267 */

struct assist_ephem* assist_ephem_create(char* planets_file_name, char* asteroids_file_name);   /* :274 */
int assist_ephem_init(struct assist_ephem* ephem, char* user_planets_path, char* user_asteroids_path); /* :290 */

struct reb_simulation* assist_simulation_convert_to_rebound(const struct reb_simulation* r, const struct assist_ephem* ephem, int merge_moon); /* :304 */

/* reference src/forces.h:30-37 */
void assist_additional_forces(struct reb_simulation* sim);
int assist_all_ephem(const struct assist_ephem* ephem, struct assist_ephem_cache* cache, const int i, const double t,
                     double* const GM,
                     double* const x, double* const y, double* const z,
                     double* const vx, double* const vy, double* const vz,
                     double* const ax, double* const ay, double* const az);

#ifdef __cplusplus
}
#endif
#endif
