/*
 * ascii_ephem.h -- the reference's header name for the DE-binary (.440 / .441) provider (reference
 * src/ascii_ephem.h).  Declarations live in assist_ephem_files.h; the two evaluators below are backed by
 * CUDA launches (assist_b200/csrc/host_api.cpp).
 */
#ifndef _ASSIST_ASCII_EPHEM_H
#define _ASSIST_ASCII_EPHEM_H

#include "assist_ephem_files.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/ascii_ephem.h:15: Chebyshev sums of one column; P[niv][ncm][ncf], t0 = fraction of the record,
 * t1 = record length in days; u, v, w hold ncm values each */
void assist_ascii_work(double* P, int ncm, int ncf, int niv, double t0, double t1, double* u, double* v, double* w);
/* reference src/ascii_ephem.h:16: barycentric state of ASSIST body `body` from a bare file handle */
enum ASSIST_STATUS assist_ascii_calc(struct ascii_s* pl, double jd_ref, double jd_rel, int body, double* const GM,
                                     double* const x, double* const y, double* const z,
                                     double* const vx, double* const vy, double* const vz,
                                     double* const ax, double* const ay, double* const az);

#ifdef __cplusplus
}
#endif
#endif
