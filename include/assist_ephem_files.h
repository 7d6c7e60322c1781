/*
 * assist_ephem_files.h -- ephemeris file providers (host side).
 *
 * Mirrors the loader interface of the reference (reference src/spk.h:44-66, :105-122
 * and src/ascii_ephem.h:13-30, :53-77): same function names, same visible struct
 * fields.  The reference evaluates Chebyshev records on the CPU straight out of the
 * mmap; here the loaders only PARSE (host) and keep a byte-exact image of the file
 * that is uploaded once per GPU; every evaluation happens in CUDA kernels
 * (assist_b200/csrc/ephem_device.cuh).
 */
#ifndef _ASSIST_B200_EPHEM_FILES_H
#define _ASSIST_B200_EPHEM_FILES_H

#include "assist.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ASSIST_B200_MAX_DEVICES 16

/* ---- SPK / DAF (.bsp), reference src/spk.h ------------------------------ */

struct mpos_s {
    double u[3];
    double v[3];
    double w[3];
};

struct mass_data {
    char** names;
    double* values;
    size_t count;
};

struct spk_constants_and_masses {
    double AU, EMRAT, J2E, J3E, J4E, J2SUN, RE, CLIGHT, ASUN;
    struct mass_data masses;
};

struct spk_target {
    int code;           /* NAIF target code */
    int cen;            /* centre */
    double mass;        /* GM, 0 if unknown */
    double beg;         /* first epoch covered (JD) */
    double end;         /* last epoch covered (JD) */
    double res;         /* span of one segment (JD) */
    int* one;           /* per segment: 1-based word address of the first record */
    int* two;           /* per segment: 1-based word address of the last word */
    int ind;            /* index of the last segment */
    int allocated_ind;
};

struct spk_s {
    struct spk_target* targets;
    int num;
    int allocated_num;
    void* map;          /* read-only mmap of the file */
    size_t len;
    /* assist-b200: device copies of the file image, one per CUDA device, made lazily */
    void* b200_dev_image[ASSIST_B200_MAX_DEVICES];
    void* b200_dev_targets[ASSIST_B200_MAX_DEVICES];
    /* assist-b200: layout + segment descriptors of the packed device copy, built once per file (gpu_api.cu) */
    void* b200_host_desc;
};

int assist_spk_free(struct spk_s* pl);
struct spk_s* assist_spk_init(const char* path);
struct spk_constants_and_masses assist_load_spk_constants_and_masses(const char* path);
void assist_apply_spk_constants(struct assist_ephem* ephem, const struct spk_constants_and_masses* data);
void assist_free_spk_constants_and_masses(struct spk_constants_and_masses* data);
void assist_spk_join_masses(struct spk_s* sp, const struct mass_data* masses, double emrat);
struct spk_target* assist_spk_find_target(const struct spk_s* pl, int code);
/* Evaluation entry points (GPU-backed; one kernel launch per call). */
enum ASSIST_STATUS assist_spk_calc_planets_by_assist(const struct assist_ephem* ephem, double jd_ref, double jd_rel,
                                   int assist_body, double* GM,
                                   double* x, double* y, double* z, double* vx, double* vy, double* vz,
                                   double* ax, double* ay, double* az);

/* ---- DE binary (.440/.441), reference src/ascii_ephem.h ----------------- */

enum {
    ASCII_MER, ASCII_VEN, ASCII_EMB, ASCII_MAR, ASCII_JUP, ASCII_SAT, ASCII_URA, ASCII_NEP,
    ASCII_PLU, ASCII_LUN, ASCII_SUN, ASCII_NUT, ASCII_LIB, ASCII_MAN, ASCII_TDB,
    ASCII_N,
};

struct ascii_s {
    double beg, end;
    double inc;
    double cau;
    double cem;
    int32_t num;
    int32_t ver;
    int32_t off[ASCII_N];
    int32_t ncf[ASCII_N];
    int32_t niv[ASCII_N];
    int32_t ncm[ASCII_N];
    double mass[ASCII_N];
    double J2E, J3E, J4E, J2SUN, AU, RE, CLIGHT, ASUN;
    size_t len, rec;
    void* map;
    double* con;
    char** str;
    void* b200_dev_image[ASSIST_B200_MAX_DEVICES];
};

struct ascii_s* assist_ascii_init(char* path);
void assist_ascii_free(struct ascii_s* ascii);
int assist_ascii_find_constant(const struct ascii_s* ascii, const char* name, double* out_value);
enum ASSIST_STATUS assist_ascii_calc_from_ephem(const struct assist_ephem* ephem, double jd_ref, double jd_rel, int body,
                                   double* const GM,
                                   double* const x, double* const y, double* const z,
                                   double* const vx, double* const vy, double* const vz,
                                   double* const ax, double* const ay, double* const az);

#ifdef __cplusplus
}
#endif
#endif
