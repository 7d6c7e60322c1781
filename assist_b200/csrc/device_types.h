/* device_types.h -- plain structs shared by the host API and the CUDA kernels. */
#ifndef AB_DEVICE_TYPES_H
#define AB_DEVICE_TYPES_H

#include <stdint.h>

#define AB_NPLANETS 11
#define AB_MAX_AST 16            /* asteroids held in the per-time body table (sb441-n16) */
#define AB_MAX_AST_ALL 1024      /* targets of a small-body kernel: those beyond the table (sb441-n373) are evaluated where the direct term needs them */
#define AB_MAX_BODIES (AB_NPLANETS + AB_MAX_AST)
#define AB_MAXSEG 4              /* SPK segments per target */
#define AB_MAX_PTGT 16           /* SPK targets in a planets kernel */
#define AB_NVMAX 6               /* variational particles per system */
#define AB_KMAX (1 + AB_NVMAX)
#define AB_BLOCK 128
#ifndef AB_PP_MIN_BLOCKS
#define AB_PP_MIN_BLOCKS 3      /* resident CTAs per SM the per-particle kernels are register-budgeted for */
#endif

#define AB_SRC_SPK 0
#define AB_SRC_ASCII 2

/* One type-2 SPK segment: what the reference re-reads from the segment trailer on every
 * evaluation (reference src/spk.c:503-517), hoisted to upload time with the same operations. */
struct AbSpkSeg {
    int one;             /* 1-based word address of the first record */
    int R, P, nrec;      /* record size, coefficients per component, record count */
    double jul_init;     /* 2451545.0 + INIT / 86400.0 */
    double intlen_d;     /* INTLEN / 86400.0 */
    double intlen_rd;    /* 1 / intlen_d */
    double radius_d;     /* RADIUS / 86400.0 (all records of a type-2 segment share RADIUS) */
    double radius_rd;    /* 1 / radius_d */
    double radius_inv;   /* 1.0 / RADIUS */
    int uniform;         /* 1 if every record really has the same RADIUS, else evaluate from the record */
    int stage_cap;       /* pp_coop_kernel: records of this segment that fit a stage buffer of the fill (set at launch) */
};

struct AbSpkTarget {
    double beg, end, res, res_rd, mass;
    int code, cen, nseg;
    int same_grid;       /* pp_coop_kernel's series table: same segments, record boundaries and record size as the series before (set at launch) */
    AbSpkSeg seg[AB_MAXSEG];
};

/* Everything a kernel needs to evaluate the ephemeris; passed by value (__grid_constant__). */
struct AbEphem {
    double jd_ref;
    int planets_source;
    int n_ast;                   /* asteroids in the body tables: min(targets of the small-body kernel, AB_MAX_AST) */
    int n_ast_x;                 /* further asteroids (sb441-n373: 357), direct term only, evaluated on the fly */
    /* constants, reference src/assist.h:140-154 */
    double AU, EMRAT, J2E, J3E, J4E, J2SUN, Re_eq, Rs_eq, c_squared, over_c_squared;
    /* unit-conversion divisors and their reciprocals (position, velocity, acceleration) */
    double u_d[3], u_rd[3];
    /* DE binary planets */
    const double* ascii_img;
    double a_beg, a_end, a_inc, a_cau, a_cem;
    long long a_rec_words;
    long long a_nrec;
    double a_inc_rd;
    int a_off[15], a_ncf[15], a_niv[15];
    double a_c[15];              /* (niv*2)/inc/86400.0 per column (reference src/ascii_ephem.c:37) */
    double a_f_earth, a_f_moon;  /* -1/(1+EMRAT), EMRAT/(1+EMRAT) (reference src/ascii_ephem.c:319-341) */
    double a_mass[AB_NPLANETS];
    /* SPK planets */
    const double* spkp_img;
    int p_index[AB_NPLANETS];
    int emb_index;
    int n_ptgt;
    AbSpkTarget p_tgt[AB_MAX_PTGT];
    double gm[AB_MAX_BODIES];      /* GM of every body, as assist_all_ephem reports it */
    /* SPK asteroids: descriptors live in global memory */
    const double* spka_img;
    const AbSpkTarget* a_tgt;
    /* Julian dates every body the force model reads is covered for (intersection over the targets), so that the
     * per-step coverage check is two comparisons; cov_simple == 0: a target is missing, ask the full check */
    double cov_lo, cov_hi;
    int cov_simple;
};

/* Force-model switches, snapshot of struct assist_extras at launch (reference src/assist.h:177-195). */
struct AbForceOpts {
    int forces;
    int gr_eih_sources;
    int geocentric;
    int has_params;
    double alpha, nk, nm, nn, r0;
    /* pole orientation terms, computed on the host with libm exactly as the reference
     * does (src/forces.c:484-490, 670-676) so that device and reference share the bits */
    double e_cosa, e_sina, e_cosd, e_sind;
    double s_cosa, s_sina, s_cosd, s_sind;
};

/* Body states at one time (AU, AU/day, AU/day^2), barycentric. */
struct AbBodies {
    double gm[AB_MAX_BODIES];
    double pos[AB_MAX_BODIES][3];
    double vel[AB_NPLANETS][3];     /* valid for j < gr_eih_sources (and Earth when geocentric) */
    double earth_acc[3];            /* valid when geocentric */
    /* particle-independent parts of the EIH term (reference src/forces.c:1418-1434, 1755-1770) */
    double eih_term1[AB_NPLANETS];
    double eih_ar[AB_NPLANETS][3];  /* a_j as the real-particle pass rounds it */
    double eih_av[AB_NPLANETS][3];  /* a_j as the variational pass rounds it */
    double t;                       /* the time the table is for (the asteroids beyond the table are evaluated at it) */
    int status;
};

/* Body table of one Gauss-Radau node in the common configuration (one EIH source, barycentric):
 * same member names as AbBodies so the force routines take either. */
struct AbNode {
    const double* gm;               /* -> AbEphem::gm */
    double pos[AB_MAX_BODIES][3];
    double vel[1][3];               /* Sun */
    double earth_acc[3];            /* unused (barycentric) */
    double eih_term1[1];
    double eih_ar[1][3];
    double eih_av[1][3];
    double t;
};

/* time slices of the work-queue scheduler (kernels.cu, pp_queue_kernel) */
struct AbSlices {
    double origin, wlen;      /* wlen carries the direction of integration */
    int n_win;
    int* done;                /* [n] windows completed per system */
    int* epoch;               /* [n] next output epoch per system (epoch runs) */
    long long attempt_budget; /* step attempts per system and call before it is retired with an error (<= 0: unlimited) */
    const int* order;         /* [n] or NULL: the queue hands out system order[k] as its k-th item of a window (longest expected first) */
};

#define AB_NODE_DOUBLES 95   /* doubles of an AbNode after the gm pointer */

/* Device-side state of a batch.  Arrays are structure-of-arrays over systems:
 * element (component k, system i) lives at [k * n + i]; the seven-deep IAS15
 * tables at [(j * C + k) * n + i], C = 3 * K. */
struct AbBatch {
    int n;            /* systems */
    int K;            /* bodies per system (1 + max variational) */
    int C;            /* 3 * K */
    int mode;
    double *pos, *vel, *acc;                 /* [C][n]  current particles */
    double *x0, *v0, *a0, *csx, *csv;        /* [C][n] */
    double *b, *g, *e, *csb, *br, *er;       /* [7][C][n] */
    double *ls_pos, *ls_vel, *ls_acc;        /* [C][n] state at the start of the last completed step */
    double *prm;                             /* [C][n] A1 A2 A3 / dA1 dA2 dA3 */
    int *nv;                                 /* [n] variational particles in use */
    /* per-particle mode */
    double *t, *dt, *dt_last, *last_full_dt; /* [n] */
    int *status;                             /* [n] REB_STATUS */
    /* counters: [n] in per-particle mode, [1] in shared-step mode */
    unsigned long long *steps, *rejected, *iters, *evals;
    /* shared-step control block (device) */
    struct AbShared* sh;
    double epsilon, min_dt;
    int has_params;
};

struct AbShared {
    double t, dt, dt_last, last_full_dt;
    int status;
    int pad;
    unsigned long long steps, rejected, iters, evals;
    unsigned long long barrier;               /* monotone grid-barrier counter */
    unsigned long long red[8][2];             /* ring of max-reduction slots (bit patterns of non-negative doubles) */
    int err_status;                           /* first ASSIST_STATUS error seen */
};

/* ---- pp_coop_kernel (coop_device.cuh): CTA geometry and the launch-time plan of its worker warps ---- */
/* A CTA holds ABC_GROUPS independent groups of ABC_SLOTS systems; a group has its own ABC_GWARPS warps (3 component
 * warps, 1 control warp, 4 workers), its own shared-memory block and global table.  The groups walk through the
 * phases of a step attempt TOGETHER (CTA-wide barriers): twice the systems under every latency-bound phase, and all
 * warps of the SM execute the same code at the same time (the hot code is larger than the instruction cache). */
#define ABC_SLOTS 32
#define ABC_GROUPS 2
#define ABC_GWARPS 8
#define ABC_WARPS (ABC_GROUPS * ABC_GWARPS)
#define ABC_THREADS (ABC_WARPS * 32)
#define ABC_CTAS_PER_SM 1
#define ABC_CTRL_WARP 3          /* roles by warp index inside the group */
#define ABC_FIRST_WORKER 4
#define ABC_NWORK (ABC_GWARPS - ABC_FIRST_WORKER)

/* per-group table in GLOBAL memory (L2 resident): what only one task reads at a node -- the asteroid positions and the
 * Sun's velocity / particle-independent EIH sums -- as [node][entry][slot] */
#define ABC_GT_AST(m, c) ((m) * 3 + (c))
#define ABC_GT_SVEL(c) (3 * AB_MAX_AST + (c))
#define ABC_GT_TERM1 (3 * AB_MAX_AST + 3)
#define ABC_GT_AR(c) (3 * AB_MAX_AST + 4 + (c))
#define ABC_GT_E (3 * AB_MAX_AST + 7)
#define ABC_GT_NODE (ABC_GT_E * ABC_SLOTS)
#define ABC_GT_DOUBLES (8 * ABC_GT_NODE)

/* task kinds of the workers */
#define ABC_T_BODY0 0            /* 0..26: body index */
#define ABC_T_EARTHJ 27
#define ABC_T_SUNJ2 28
#define ABC_T_NG 29
#define ABC_T_GRPOT 30
#define ABC_T_GRSIMPLE 31
#define ABC_T_EIHSRC 32          /* EIH source block of the Sun (everything but the potential sum over the planets) */
#define ABC_T_NONE 255

/* launch-time plan of the worker warps: up to two of the single-body terms, a list of planets (positions in shared
 * memory) and a list of asteroids (positions in the global table, fetched four at a time ahead of their use) */
#define ABC_MAX_GROUP 16
struct AbcWorkerPlan {
    unsigned char scalar[2];          /* ABC_T_EARTHJ ..., ABC_T_NONE */
    unsigned char nbody;              /* planets (and the Sun) in body[] */
    unsigned char nast;               /* asteroids in ast[] (index m of the small-body kernel) */
    unsigned char body[ABC_MAX_GROUP];
    unsigned char ast[ABC_MAX_GROUP];
};
struct AbcPlan {
    AbcWorkerPlan w[ABC_NWORK];
    long long attempt_budget; /* step attempts per system and call before it is retired with an error; <= 0: none */
};

#endif
