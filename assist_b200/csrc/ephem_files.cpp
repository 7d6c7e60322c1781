/*
 * ephem_files.cpp -- host-side parsers for the two ephemeris containers.
 *
 * Parsing only: segment directories, constants and masses.  The Chebyshev records
 * themselves are never evaluated here; the whole file stays mmap'ed so that
 * gpu_api.cu can copy its image to HBM once per device.
 *
 *   SPK/DAF   behaviour of reference src/spk.c:214-402 (file + summary records),
 *             :49-120 and :696-789 (comment-area constants), :137-210 (mass join)
 *   DE binary behaviour of reference src/ascii_ephem.c:105-252
 */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "assist_ephem_files.h"
#include "host_internal.h"
#include "assist_gpu.h"

namespace {

const size_t kDafRecord = 1024;

#pragma pack(push, 1)
struct DafSummary {        /* nd = 2 doubles, ni = 6 ints: 40 bytes */
    double beg, end;
    int32_t tar, cen, ref, type, one, two;
};
#pragma pack(pop)
static_assert(sizeof(DafSummary) == 40, "DAF summary must be 40 bytes");

double jd_of_et(double et) { return 2451545.0 + et / 86400.0; }

bool read_at(int fd, off_t off, void* buf, size_t len) {
    return pread(fd, buf, len, off) == (ssize_t)len;
}

/* Reads and validates the DAF file record; returns fward (record number of the first summary) or -1. */
int daf_file_record(int fd) {
    unsigned char rec[kDafRecord];
    if (!read_at(fd, 0, rec, sizeof(rec))) {
        fprintf(stderr, "Incomplete read. Expected %zu bytes.\n", kDafRecord);
        return -1;
    }
    if (memcmp(rec, "DAF/SPK", 7) != 0) {
        fprintf(stderr, "Error parsing DAF/SPK file. Incorrect header.\n");
        return -1;
    }
    int32_t nd, ni, fward;
    memcpy(&nd, rec + 8, 4);
    memcpy(&ni, rec + 12, 4);
    memcpy(&fward, rec + 76, 4);
    if (8 * (nd + (ni + 1) / 2) != (int)sizeof(DafSummary)) {
        fprintf(stderr, "Error parsing DAF/SPK file. Wrong size of summary record.\n");
        return -1;
    }
    return fward;
}

/* The comment area (records 2 .. fward-1) as text, '\n' separated. */
std::string daf_comments(int fd, int fward) {
    std::string text;
    for (int r = 2; r < fward; r++) {
        char rec[kDafRecord];
        if (!read_at(fd, (off_t)(r - 1) * kDafRecord, rec, sizeof(rec))) break;
        size_t len = sizeof(rec);
        while (len > 0 && (rec[len - 1] == '\0' || rec[len - 1] == '\4')) len--;   /* padding / end-of-text */
        text.append(rec, len);
    }
    for (char& c : text) if (c == '\0') c = '\n';
    return text;
}

}  // namespace

extern "C" {

struct spk_target* assist_spk_find_target(const struct spk_s* pl, int code) {
    if (!pl) return NULL;
    for (int i = 0; i < pl->num; i++)
        if (pl->targets[i].code == code) return &pl->targets[i];
    return NULL;
}

int assist_spk_free(struct spk_s* pl) {
    if (pl == NULL) return -1;
    for (int d = 0; d < ASSIST_B200_MAX_DEVICES; d++) {
        if (pl->b200_dev_image[d] || pl->b200_dev_targets[d]) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(d);
            cudaFree(pl->b200_dev_image[d]);
            cudaFree(pl->b200_dev_targets[d]);
            cudaSetDevice(cur);
        }
    }
    ab_spk_desc_free(pl->b200_host_desc);
    pl->b200_host_desc = NULL;
    if (pl->targets) {
        for (int m = 0; m < pl->num; m++) { free(pl->targets[m].one); free(pl->targets[m].two); }
        free(pl->targets);
    }
    if (pl->map) munmap(pl->map, pl->len);
    free(pl);
    return 0;
}

struct spk_s* assist_spk_init(const char* path) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return NULL;
    const int fward = daf_file_record(fd);
    if (fward < 0) { close(fd); return NULL; }

    struct spk_s* pl = (struct spk_s*)calloc(1, sizeof(struct spk_s));
    long recno = fward;
    bool first = true;
    while (recno > 0) {
        unsigned char rec[kDafRecord];
        if (!read_at(fd, (off_t)(recno - 1) * kDafRecord, rec, sizeof(rec))) break;
        double next, prev, nsum;
        memcpy(&next, rec, 8); memcpy(&prev, rec + 8, 8); memcpy(&nsum, rec + 16, 8);
        if (first && rec[8] != 0) {     /* the first summary record has no predecessor */
            fprintf(stderr, "Error parsing DAF/SPL file. Cannot find summary block.\n");
            close(fd);
            assist_spk_free(pl);
            return NULL;
        }
        first = false;
        for (int s = 0; s < (int)nsum; s++) {
            DafSummary sum;
            memcpy(&sum, rec + 24 + s * sizeof(DafSummary), sizeof(sum));
            struct spk_target* tg = assist_spk_find_target(pl, sum.tar);
            if (tg == NULL) {
                if (pl->num >= pl->allocated_num) {
                    pl->allocated_num += 32;
                    pl->targets = (struct spk_target*)realloc(pl->targets, pl->allocated_num * sizeof(struct spk_target));
                }
                tg = &pl->targets[pl->num++];
                memset(tg, 0, sizeof(*tg));
                tg->code = sum.tar;
                tg->cen = sum.cen;
                tg->beg = jd_of_et(sum.beg);
                tg->res = jd_of_et(sum.end) - tg->beg;     /* span of one segment; all segments of a target are equal */
                tg->ind = -1;
            }
            const int next_seg = tg->ind + 1;
            if (next_seg >= tg->allocated_ind) {
                int cap = tg->allocated_ind ? tg->allocated_ind * 2 : 32;
                tg->one = (int*)realloc(tg->one, cap * sizeof(int));
                tg->two = (int*)realloc(tg->two, cap * sizeof(int));
                tg->allocated_ind = cap;
            }
            tg->ind = next_seg;
            tg->one[next_seg] = sum.one;
            tg->two[next_seg] = sum.two;
            tg->end = jd_of_et(sum.end);
        }
        recno = (long)next;
    }

    struct stat sb;
    if (fstat(fd, &sb) < 0) {
        fprintf(stderr, "Error calculating size for DAF/SPL file.\n");
        close(fd);
        assist_spk_free(pl);
        return NULL;
    }
    pl->len = sb.st_size;
    pl->map = mmap(NULL, pl->len, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (pl->map == MAP_FAILED) {
        pl->map = NULL;
        fprintf(stderr, "Error creating memory map.\n");
        assist_spk_free(pl);
        return NULL;
    }
    return pl;
}

struct spk_constants_and_masses assist_load_spk_constants_and_masses(const char* path) {
    struct spk_constants_and_masses data;
    memset(&data, 0, sizeof(data));
    int fd = open(path, O_RDONLY);
    if (fd < 0) return data;
    const int fward = daf_file_record(fd);
    if (fward < 0) { close(fd); return data; }
    std::string text = daf_comments(fd, fward);
    close(fd);

    bool in_constants = false;
    size_t pos = 0;
    while (pos <= text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string line = text.substr(pos, eol - pos);
        pos = eol + 1;
        if (line.find("Initial conditions and constants used for integration:") != std::string::npos) in_constants = true;
        if (!in_constants) continue;
        /* Fortran exponents: every D/d in the line becomes e (as the reference does, src/spk.c:114-120) */
        for (char& c : line) if (c == 'D' || c == 'd') c = 'e';
        char key[64], val[64];
        if (sscanf(line.c_str(), "%63s %63s", key, val) != 2) continue;
        const double v = strtod(val, NULL);
        if (!strcmp(key, "cau") || !strcmp(key, "AU")) data.AU = v;
        else if (!strcmp(key, "EMRAT")) data.EMRAT = v;
        else if (!strcmp(key, "J2E")) data.J2E = v;
        else if (!strcmp(key, "J3E")) data.J3E = v;
        else if (!strcmp(key, "J4E")) data.J4E = v;
        else if (!strcmp(key, "J2SUN")) data.J2SUN = v;
        else if (!strcmp(key, "RE")) data.RE = v;
        else if (!strcmp(key, "CLIGHT")) data.CLIGHT = v;
        else if (!strcmp(key, "ASUN")) data.ASUN = v;
        else if (!strncmp(key, "GM", 2) || !strncmp(key, "MA", 2)) {
            data.masses.names = (char**)realloc(data.masses.names, (data.masses.count + 1) * sizeof(char*));
            data.masses.values = (double*)realloc(data.masses.values, (data.masses.count + 1) * sizeof(double));
            data.masses.names[data.masses.count] = strdup(key);
            data.masses.values[data.masses.count] = v;
            data.masses.count++;
        }
    }
    return data;
}

void assist_apply_spk_constants(struct assist_ephem* ephem, const struct spk_constants_and_masses* data) {
    if (!ephem || !data) return;
    ephem->AU = data->AU; ephem->EMRAT = data->EMRAT;
    ephem->J2E = data->J2E; ephem->J3E = data->J3E; ephem->J4E = data->J4E; ephem->J2SUN = data->J2SUN;
    ephem->RE = data->RE; ephem->CLIGHT = data->CLIGHT; ephem->ASUN = data->ASUN;
    ephem->Re_eq = ephem->RE / ephem->AU;
    ephem->Rs_eq = ephem->ASUN / ephem->AU;
    ephem->c_AU_per_day = (ephem->CLIGHT / ephem->AU) * 86400;
    ephem->c_squared = ephem->c_AU_per_day * ephem->c_AU_per_day;
    ephem->over_c_squared = 1.0 / ephem->c_squared;
}

void assist_free_spk_constants_and_masses(struct spk_constants_and_masses* data) {
    if (!data || !data->masses.names) return;
    for (size_t i = 0; i < data->masses.count; i++) free(data->masses.names[i]);
    free(data->masses.names);
    free(data->masses.values);
    data->masses.names = NULL; data->masses.values = NULL; data->masses.count = 0;
}

void assist_spk_join_masses(struct spk_s* sp, const struct mass_data* masses, double emrat) {
    if (sp == NULL || masses == NULL || masses->names == NULL) return;
    static const struct { const char* name; int code; } planet_codes[] = {
        {"GMS", 10}, {"GM1", 1}, {"GM2", 2}, {"GMB", 399}, {"GMB", 3}, {"GMB", 301},
        {"GM4", 4}, {"GM5", 5}, {"GM6", 6}, {"GM7", 7}, {"GM8", 8}, {"GM9", 9}};
    for (int m = 0; m < sp->num; m++) {
        struct spk_target* tg = &sp->targets[m];
        if (tg->mass != 0) continue;
        char label[64] = "";
        for (size_t i = 0; i < sizeof(planet_codes) / sizeof(planet_codes[0]); i++)
            if (tg->code == planet_codes[i].code) { snprintf(label, sizeof(label), "%s", planet_codes[i].name); break; }
        if (label[0] == '\0') snprintf(label, sizeof(label), "MA%04d", tg->code - 2000000);
        for (size_t i = 0; i < masses->count; i++) {
            if (strcmp(masses->names[i], label) != 0) continue;
            if (tg->code == 399) tg->mass = masses->values[i] * (emrat / (1. + emrat));
            else if (tg->code == 301) tg->mass = masses->values[i] * (1. / (1. + emrat));
            else tg->mass = masses->values[i];
            break;
        }
        if (tg->mass == 0 && tg->code != 199 && tg->code != 299)
            printf("Mass not found for target code: %d\n", tg->code);
    }
}

/* ---- DE binary ----------------------------------------------------------- */

static double ascii_constant(const struct ascii_s* a, const char* name6) {
    for (int p = 0; p < a->num; p++)
        if (strncmp(name6, a->str[p], 6) == 0) return a->con[p];
    fprintf(stderr, "WARNING: Constant [%s] not found in ephemeris file.\n", name6);
    return 0;
}

int assist_ascii_find_constant(const struct ascii_s* ascii, const char* name, double* out_value) {
    if (out_value) *out_value = 0.0;
    if (!ascii || !ascii->str || !ascii->con || !name || !out_value) return 0;
    char key[6];
    memset(key, ' ', 6);
    for (int i = 0; i < 6 && name[i] != '\0'; i++) key[i] = name[i];
    for (int p = 0; p < ascii->num; p++)
        if (memcmp(ascii->str[p], key, 6) == 0) { *out_value = ascii->con[p]; return 1; }
    return 0;
}

struct ascii_s* assist_ascii_init(char* path) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return NULL;
    struct stat sb;
    if (fstat(fd, &sb) < 0) {
        close(fd);
        fprintf(stderr, "Error while trying to determine filesize.\n");
        return NULL;
    }
    /* fixed part of the header at 0x0A5C: 3 doubles, NCON, AU, EMRAT, 12 triplets, DENUM, 1 triplet */
    unsigned char hdr[44 + 12 * 12 + 4 + 12];
    if (!read_at(fd, 0x0A5C, hdr, sizeof(hdr))) {
        close(fd);
        fprintf(stderr, "Error while seeking to header.\n");
        return NULL;
    }
    struct ascii_s* a = (struct ascii_s*)calloc(1, sizeof(struct ascii_s));
    size_t o = 0;
    memcpy(&a->beg, hdr + o, 8); o += 8;
    memcpy(&a->end, hdr + o, 8); o += 8;
    memcpy(&a->inc, hdr + o, 8); o += 8;
    memcpy(&a->num, hdr + o, 4); o += 4;
    memcpy(&a->cau, hdr + o, 8); o += 8;
    memcpy(&a->cem, hdr + o, 8); o += 8;
    for (int p = 0; p < ASCII_N; p++) a->ncm[p] = 3;
    a->ncm[ASCII_NUT] = 2;
    a->ncm[ASCII_TDB] = 1;
    for (int p = 0; p < 12; p++) {
        memcpy(&a->off[p], hdr + o, 4); memcpy(&a->ncf[p], hdr + o + 4, 4); memcpy(&a->niv[p], hdr + o + 8, 4);
        o += 12;
    }
    memcpy(&a->ver, hdr + o, 4); o += 4;
    memcpy(&a->off[12], hdr + o, 4); memcpy(&a->ncf[12], hdr + o + 4, 4); memcpy(&a->niv[12], hdr + o + 8, 4);
    if (a->num < 400 || a->num > 100000 || !(a->inc > 0)) {
        fprintf(stderr, "Error: implausible header in DE binary file.\n");
        close(fd); free(a);
        return NULL;
    }
    /* constant names: 400 at 0x00FC, the rest at 0x0B28, then the last two column triplets */
    a->str = (char**)calloc(a->num, sizeof(char*));
    for (int p = 0; p < a->num; p++) {
        a->str[p] = (char*)calloc(8, 1);
        const off_t at = (p < 400) ? (0x00FC + 6 * (off_t)p) : (0x0B28 + 6 * (off_t)(p - 400));
        read_at(fd, at, a->str[p], 6);
    }
    {
        unsigned char tr[24];
        read_at(fd, 0x0B28 + 6 * (off_t)(a->num - 400), tr, sizeof(tr));
        for (int p = 13; p < 15; p++) {
            memcpy(&a->off[p], tr + 12 * (p - 13), 4);
            memcpy(&a->ncf[p], tr + 12 * (p - 13) + 4, 4);
            memcpy(&a->niv[p], tr + 12 * (p - 13) + 8, 4);
        }
    }
    for (int p = 0; p < ASCII_N; p++) a->off[p] -= 1;      /* zero based */
    a->len = sb.st_size;
    a->rec = sizeof(double) * 2;
    for (int p = 0; p < ASCII_N; p++) a->rec += sizeof(double) * a->ncf[p] * a->niv[p] * a->ncm[p];

    a->map = mmap(NULL, a->len, PROT_READ, MAP_SHARED, fd, 0);
    if (a->map == MAP_FAILED) {
        a->map = NULL;
        close(fd);
        assist_ascii_free(a);
        fprintf(stderr, "Error while calling mmap().\n");
        return NULL;
    }
    a->con = (double*)calloc(a->num, sizeof(double));
    read_at(fd, (off_t)a->rec, a->con, sizeof(double) * a->num);   /* record 1 holds the values */
    close(fd);

    a->mass[ASSIST_BODY_SUN] = ascii_constant(a, "GMS   ");
    a->mass[ASSIST_BODY_MERCURY] = ascii_constant(a, "GM1   ");
    a->mass[ASSIST_BODY_VENUS] = ascii_constant(a, "GM2   ");
    const double emrat = ascii_constant(a, "EMRAT ");
    const double gmb = ascii_constant(a, "GMB   ");
    a->mass[ASSIST_BODY_EARTH] = (emrat / (1. + emrat)) * gmb;
    a->mass[ASSIST_BODY_MOON] = 1. / (1 + emrat) * gmb;
    a->mass[ASSIST_BODY_MARS] = ascii_constant(a, "GM4   ");
    a->mass[ASSIST_BODY_JUPITER] = ascii_constant(a, "GM5   ");
    a->mass[ASSIST_BODY_SATURN] = ascii_constant(a, "GM6   ");
    a->mass[ASSIST_BODY_URANUS] = ascii_constant(a, "GM7   ");
    a->mass[ASSIST_BODY_NEPTUNE] = ascii_constant(a, "GM8   ");
    a->mass[ASSIST_BODY_PLUTO] = ascii_constant(a, "GM9   ");
    a->J2E = ascii_constant(a, "J2E   ");
    a->J3E = ascii_constant(a, "J3E   ");
    a->J4E = ascii_constant(a, "J4E   ");
    a->J2SUN = ascii_constant(a, "J2SUN ");
    a->AU = ascii_constant(a, "AU    ");
    a->RE = ascii_constant(a, "RE    ");
    a->CLIGHT = ascii_constant(a, "CLIGHT");
    a->ASUN = ascii_constant(a, "ASUN  ");
    return a;
}

void assist_ascii_free(struct ascii_s* ascii) {
    if (ascii == NULL) return;
    for (int d = 0; d < ASSIST_B200_MAX_DEVICES; d++) {
        if (ascii->b200_dev_image[d]) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(d);
            cudaFree(ascii->b200_dev_image[d]);
            cudaSetDevice(cur);
        }
    }
    if (ascii->map) munmap(ascii->map, ascii->len);
    if (ascii->str) for (int p = 0; p < ascii->num; p++) free(ascii->str[p]);
    free(ascii->str);
    free(ascii->con);
    free(ascii);
}

}  // extern "C"
