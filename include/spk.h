/*
 * spk.h -- the reference's header name for the SPK provider (reference src/spk.h).
 * Programs written against ASSIST include "spk.h" next to "assist.h" (reference unit_tests/spk_init,
 * spk_join_masses, spk_load_constants, spk_planets_calc, ascii_reject); the declarations live in
 * assist_ephem_files.h, the evaluators below are backed by CUDA launches (assist_b200/csrc/host_api.cpp).
 */
#ifndef _SPK_H
#define _SPK_H

#include "assist_ephem_files.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/spk.h:113: heliocentric position of target m of a small-body kernel in AU (position / 149597870.7) */
enum ASSIST_STATUS assist_spk_calc(const struct spk_s* pl, double jd_ref, double jd_rel, int m, double* GM,
                                   double* x, double* y, double* z);
/* reference src/spk.h:114: barycentric state of NAIF target `code` of ephem->spk_planets in AU, AU/day, AU/day^2 */
enum ASSIST_STATUS assist_spk_calc_planets(const struct assist_ephem* ephem, double jd_ref, double jd_rel, int code, double* GM,
                                           double* x, double* y, double* z, double* vx, double* vy, double* vz,
                                           double* ax, double* ay, double* az);
/* reference src/spk.h:122: Chebyshev sums of one target in the file's units (km, km/s, km/s^2) */
struct mpos_s assist_spk_target_pos(const struct spk_s* pl, const struct spk_target* target, double jd_ref, double jd_rel);

#ifdef __cplusplus
}
#endif
#endif
