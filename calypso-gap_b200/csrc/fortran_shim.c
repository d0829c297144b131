/*
 * fortran_shim.c -- the five external symbols the f2py module libgap.libgap links
 * against, with the gfortran calling convention (lower case + trailing
 * underscore, every argument by reference, hidden string length last), for
 * environments without a Fortran compiler (this image).  Each one forwards to
 * the C ABI exactly as fortran/libgap_driver.f90 does.
 *
 * Error behaviour: message on stdout, then the process stops.  Where the reference itself STOPs
 * -- gap_parameters missing or malformed (gap_calc.f90:323-327, wacsf.f90:40-47) and more than 1000
 * neighbours (gap_calc.f90:107-111) -- the exit status is 0 like a Fortran STOP.  Every other failure
 * (CUDA error, no device, a limit of this library, bad arguments, a species absent from the file) is
 * not a reference condition: the process exits with status 1 (ERROR STOP 1) so that a shell or a
 * scheduler does not take it for a finished run.  Set GAPCU_ERRORS=return to return to the caller with
 * NaN outputs instead (the tests and any Python host that prefers an exception).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gapcu.h"

void gapcu_print_last_error(void) { printf(" %s\n", gapcu_last_error()); fflush(stdout); }

static int stop_or_return(int code) {
    gapcu_print_last_error();
    const char *m = getenv("GAPCU_ERRORS");
    if (m && !strcmp(m, "return")) return 1;
    exit((code == GAPCU_EFILE || code == GAPCU_ENEIGH) ? 0 /* Fortran STOP */ : 1 /* ERROR STOP 1 */);
}

void fgap_calc_(int *na, int *species, double *lat, double *pos, double *ene, double *force, double *stress,
                double *variance, int *nsparsex, int *des_len, double *theta, double *mm, double *qmm,
                double *coeff, double *rcut, int *lgrad) {
    const int rc = gapcu_calc(*na, species, lat, pos, *nsparsex, *des_len, theta, mm, qmm, coeff, *rcut, *lgrad != 0, ene,
                              force, stress, variance);
    if (rc != 0 && stop_or_return(rc)) {
        *ene = NAN; *variance = NAN;
        for (int i = 0; i < 3 * *na; i++) force[i] = NAN;
        for (int i = 0; i < 6; i++) stress[i] = NAN;
    }
}

void fgap_read_(int *nsparsex, int *des_len, double *theta, double *mm, double *invcmm, double *coeff) {
    /* fixed capacities of the reference: nsf_max = 100, nsparseX_max = 4000 (gap_calc.f90:306-307) */
    const int rc = gapcu_read("gap_parameters", nsparsex, des_len, theta, 100, mm, 4000, 100, invcmm, 4000, coeff, 4000);
    /* exceeding nsf_max / nsparseX_max is a STOP of the reference too (gap_calc.f90:330-341) */
    if (rc != 0 && stop_or_return(rc == GAPCU_ELIMIT ? GAPCU_EFILE : rc)) {
        *nsparsex = 0; *des_len = 0;
    }
}

void fget_bond_(int *na, double *lat, int *elements, double *pos, double *rcut, double *min_bond) {
    const int rc = gapcu_bond(*na, lat, elements, pos, *rcut, min_bond);
    if (rc != 0 && stop_or_return(rc)) *min_bond = NAN;
}

void car2acsf_(int *na, int *max_neighbor, int *nf, double *pos, double *neighbor, int *neighbor_count, double *xx,
               double *dxdy, double *strs, int *lgrad) {
    const int rc = gapcu_car2acsf_table(*na, *max_neighbor, *nf, pos, neighbor, neighbor_count, *lgrad != 0, xx, dxdy, strs);
    if (rc != 0 && stop_or_return(rc)) {
        for (long i = 0; i < (long)*nf * *na; i++) xx[i] = NAN;
    }
}

/* wacsf.f90:798-810: plain text dump, one row per line, F20.10 */
void write_array_2dim_(int *n, int *m, double *a, char *name, size_t name_len) {
    char path[4096];
    size_t b = 0, e = name_len;
    while (b < e && name[b] == ' ') b++;
    while (e > b && name[e - 1] == ' ') e--;
    if (e - b >= sizeof path) return;
    memcpy(path, name + b, e - b);
    path[e - b] = 0;
    FILE *f = fopen(path, "w");
    if (!f) return;
    for (int i = 0; i < *n; i++) {
        for (int j = 0; j < *m; j++) fprintf(f, "%20.10f", a[i + (size_t)*n * j]);
        fprintf(f, "\n");
    }
    fclose(f);
}
