/*
 * gap_oracle.c -- CPU restatement of CALYPSO-GAP's libgap E/F/stress path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under calypso-gap_b200/ may include, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this code against
 * the 11 frames of real libgap output shipped by the reference in
 * gappy/example/ASE-GAPPY/ase.traj (fixture tests/golden/ase_traj_frames.npz).
 * What the golden data does NOT pin (stress off-diagonal order, mixed species
 * weights, triclinic cells, self-images, lgrad=.false.) is pinned by code
 * reading only and cross-checked by finite differences (tests/test_oracle_consistency.py).
 *
 * The reference itself cannot be compiled here (Fortran 90, no Fortran compiler
 * in the image), so this is a restatement in C of the algorithm in
 *   gappy/libgap/gap_calc.f90   (FGAP_CALC :1-301, GET_COV :268-288, FGAP_READ :303-364)
 *   gappy/libgap/wacsf.f90      (CAR2ACSF :2-796)
 *   gappy/libgap/get_bond.f90   (FGET_BOND :4-116)
 * Two forms are provided:
 *   gapo_calc_dense  -- follows the reference loop for loop: O(N^2 images)
 *                       neighbour table in (j,n1,n2,n3) order, dense
 *                       dxdy(D,N,N,3) and strs(3,3,D,N), per-SF loops with the
 *                       k_neighbor > j_neighbor triplet rule, GET_COV, dedg,
 *                       O(N^2 D) force loop.  It is the arbiter and the timed
 *                       CPU baseline.
 *   gapo_calc_sparse -- the same arithmetic in O(N) memory (cell list, chain
 *                       rule applied per neighbour instead of through dxdy),
 *                       validated against the dense form; used where the dense
 *                       form cannot run (N >~ 2000).
 * Build twice from this one source (oracle/Makefile):
 *   parity: gcc -O2 -ffp-contract=off   timing: gcc -O3 -march=native
 *
 * Array conventions of THIS file's API: C order.  lat[3][3] rows are lattice
 * vectors (gap_calc.f90:98), pos[na][3], force[na][3],
 * stress[6] = xx yy zz xy yz xz in GPa (gap_calc.f90:221-226).
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* literal constants of the reference */
static const double PI_REF = 3.141592654;   /* wacsf.f90:31, gap_calc.f90:59, get_bond.f90:21 */
static const double GPA2EVPANG = 6.24219e-3; /* gap_calc.f90:9 */
static const double DELTA_K = 1.0;           /* gap_calc.f90:8 */
static const double RMIN_WARN = 0.5;         /* gap_calc.f90:67 (warning only) */
#define MAX_NEIGHBOR_REF 1000                /* gap_calc.f90:68 */

typedef struct {
    int nspecies;
    int *z;
    double *w;
    int nsf;
    int *ntype;
    double *alpha;
    double *cutoff;
    int nsparse, des_len;
    double *theta;  /* [des_len] */
    double *mm;     /* [nsparse][des_len] */
    double *coeff;  /* [nsparse] */
} gapo_params;

/* ------------------------------------------------------------------------ */
/* gap_parameters reader: Fortran list-directed READ semantics, record based */
/* ------------------------------------------------------------------------ */
typedef struct {
    char *buf;
    size_t len, pos; /* pos = start of the next unread record */
} recfile;

/* One list-directed READ of `want` numeric items: starts at a new record,
 * takes tokens (blank/comma separated) across as many records as needed and
 * discards the rest of the last record touched.  want == 0 skips a record. */
static int rec_read(recfile *f, int want, double *out) {
    int got = 0;
    do {
        if (f->pos >= f->len) return -1;
        size_t e = f->pos;
        while (e < f->len && f->buf[e] != '\n') e++;
        size_t p = f->pos;
        while (got < want && p < e) {
            while (p < e && (isspace((unsigned char)f->buf[p]) || f->buf[p] == ',')) p++;
            if (p >= e) break;
            char tok[128];
            size_t n = 0;
            while (p < e && !isspace((unsigned char)f->buf[p]) && f->buf[p] != ',' && n < sizeof(tok) - 1) {
                char c = f->buf[p++];
                if (c == 'd' || c == 'D') c = 'e'; /* Fortran D exponent */
                tok[n++] = c;
            }
            tok[n] = 0;
            char *endp;
            out[got] = strtod(tok, &endp);
            if (endp == tok) return -2;
            got++;
        }
        f->pos = (e < f->len) ? e + 1 : e;
    } while (got < want);
    return 0;
}

void gapo_free(gapo_params *p) {
    if (!p) return;
    free(p->z); free(p->w); free(p->ntype); free(p->alpha); free(p->cutoff);
    free(p->theta); free(p->mm); free(p->coeff); free(p);
}

/* gap_calc.f90:75-83 (species weights), wacsf.f90:46-55 (SF table),
 * gap_calc.f90:330-362 (GPR data). */
gapo_params *gapo_read(const char *path, char *err, int errlen) {
    FILE *fp = fopen(path, "rb");
    if (!fp) { snprintf(err, errlen, "gap_parameters file does not exist!"); return NULL; }
    recfile f; memset(&f, 0, sizeof f);
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    f.buf = (char *)malloc((size_t)sz + 1); f.len = fread(f.buf, 1, (size_t)sz, fp); fclose(fp);
    gapo_params *p = (gapo_params *)calloc(1, sizeof *p);
    double v[3];
#define FAIL(msg) do { snprintf(err, errlen, "%s", msg); free(f.buf); gapo_free(p); return NULL; } while (0)
    if (rec_read(&f, 1, v)) FAIL("bad nspecies record");
    p->nspecies = (int)v[0];
    p->z = (int *)calloc(p->nspecies > 0 ? p->nspecies : 1, sizeof(int));
    p->w = (double *)calloc(p->nspecies > 0 ? p->nspecies : 1, sizeof(double));
    for (int i = 0; i < p->nspecies; i++) {
        if (rec_read(&f, 2, v)) FAIL("bad species record");
        p->z[i] = (int)v[0]; p->w[i] = v[1];
    }
    if (rec_read(&f, 1, v)) FAIL("bad nsf record");
    p->nsf = (int)v[0];
    p->ntype = (int *)calloc(p->nsf > 0 ? p->nsf : 1, sizeof(int));
    p->alpha = (double *)calloc(p->nsf > 0 ? p->nsf : 1, sizeof(double));
    p->cutoff = (double *)calloc(p->nsf > 0 ? p->nsf : 1, sizeof(double));
    for (int i = 0; i < p->nsf; i++) {
        if (rec_read(&f, 3, v)) FAIL("bad symmetry-function record");
        p->ntype[i] = (int)v[0]; p->alpha[i] = v[1]; p->cutoff[i] = v[2];
    }
    if (rec_read(&f, 2, v)) FAIL("bad nsparseX/des_len record");
    p->nsparse = (int)v[0]; p->des_len = (int)v[1];
    if (p->nsparse < 0 || p->des_len < 0) FAIL("negative sizes");
    for (int i = 0; i < 3; i++) if (rec_read(&f, 0, v)) FAIL("missing skipped record");
    p->theta = (double *)calloc((size_t)p->des_len + 1, sizeof(double));
    p->mm = (double *)calloc((size_t)p->nsparse * p->des_len + 1, sizeof(double));
    p->coeff = (double *)calloc((size_t)p->nsparse + 1, sizeof(double));
    if (rec_read(&f, p->des_len, p->theta)) FAIL("bad theta record");
    for (int i = 0; i < p->nsparse; i++)
        if (rec_read(&f, p->des_len, p->mm + (size_t)i * p->des_len)) FAIL("bad MM record");
    if (rec_read(&f, p->nsparse, p->coeff)) FAIL("bad coeff record");
#undef FAIL
    free(f.buf);
    return p;
}

/* Build a parameter set from arrays (synthetic potentials in tests). */
gapo_params *gapo_make(int nspecies, const int *z, const double *w, int nsf, const int *ntype,
                       const double *alpha, const double *cutoff, int nsparse, int des_len,
                       const double *theta, const double *mm, const double *coeff) {
    gapo_params *p = (gapo_params *)calloc(1, sizeof *p);
    p->nspecies = nspecies; p->nsf = nsf; p->nsparse = nsparse; p->des_len = des_len;
    p->z = (int *)malloc(sizeof(int) * (nspecies + 1)); memcpy(p->z, z, sizeof(int) * nspecies);
    p->w = (double *)malloc(sizeof(double) * (nspecies + 1)); memcpy(p->w, w, sizeof(double) * nspecies);
    p->ntype = (int *)malloc(sizeof(int) * (nsf + 1)); memcpy(p->ntype, ntype, sizeof(int) * nsf);
    p->alpha = (double *)malloc(sizeof(double) * (nsf + 1)); memcpy(p->alpha, alpha, sizeof(double) * nsf);
    p->cutoff = (double *)malloc(sizeof(double) * (nsf + 1)); memcpy(p->cutoff, cutoff, sizeof(double) * nsf);
    p->theta = (double *)malloc(sizeof(double) * (des_len + 1)); memcpy(p->theta, theta, sizeof(double) * des_len);
    p->mm = (double *)malloc(sizeof(double) * ((size_t)nsparse * des_len + 1));
    memcpy(p->mm, mm, sizeof(double) * (size_t)nsparse * des_len);
    p->coeff = (double *)malloc(sizeof(double) * (nsparse + 1)); memcpy(p->coeff, coeff, sizeof(double) * nsparse);
    return p;
}

void gapo_info(const gapo_params *p, int *nspecies, int *nsf, int *nsparse, int *des_len) {
    *nspecies = p->nspecies; *nsf = p->nsf; *nsparse = p->nsparse; *des_len = p->des_len;
}

void gapo_get(const gapo_params *p, int *z, double *w, int *ntype, double *alpha, double *cutoff,
              double *theta, double *mm, double *coeff) {
    memcpy(z, p->z, sizeof(int) * p->nspecies); memcpy(w, p->w, sizeof(double) * p->nspecies);
    memcpy(ntype, p->ntype, sizeof(int) * p->nsf); memcpy(alpha, p->alpha, sizeof(double) * p->nsf);
    memcpy(cutoff, p->cutoff, sizeof(double) * p->nsf);
    memcpy(theta, p->theta, sizeof(double) * p->des_len);
    memcpy(mm, p->mm, sizeof(double) * (size_t)p->nsparse * p->des_len);
    memcpy(coeff, p->coeff, sizeof(double) * p->nsparse);
}

/* ------------------------------------------------------------------------ */
/* lattice helpers: gap_calc.f90:235-266 (vectorlength, recipvector, volume, */
/* crossp) and :290-299 (Det).  Note recipvector works on COLUMNS of lat and */
/* nabc reads ROWS of the result, which together give ceil(rcut/d_i) for the */
/* row-vector lattice.                                                        */
/* ------------------------------------------------------------------------ */
static void crossp(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

static void image_range(const double lat[9], double rcut, int nabc[3]) {
    double col[3][3], cr[3], rec[3][3]; /* rec[r][c] = recipvector(r+1,c+1) */
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) col[c][r] = lat[r * 3 + c];
    crossp(col[1], col[2], cr);
    double vol = fabs((col[0][0] * cr[0] + col[0][1] * cr[1]) + col[0][2] * cr[2]);
    double t[3];
    crossp(col[1], col[2], t); for (int r = 0; r < 3; r++) rec[r][0] = t[r];
    crossp(col[2], col[0], t); for (int r = 0; r < 3; r++) rec[r][1] = t[r];
    crossp(col[0], col[1], t); for (int r = 0; r < 3; r++) rec[r][2] = t[r];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rec[r][c] = rec[r][c] / vol * PI_REF * 2.0;
    for (int r = 0; r < 3; r++) {
        double len = sqrt((rec[r][0] * rec[r][0] + rec[r][1] * rec[r][1]) + rec[r][2] * rec[r][2]);
        nabc[r] = (int)ceil(rcut * len / PI_REF / 2);
    }
}

static double det3(const double m[9]) { /* gap_calc.f90:295-297, m[r*3+c] = Matrix(r+1,c+1) */
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
           m[2] * (m[3] * m[7] - m[6] * m[4]);
}

void gapo_image_range(const double *lat, double rcut, int *nabc) { image_range(lat, rcut, nabc); }

/* image of atom j and its distance from atom i, in the reference's operation
 * order (gap_calc.f90:98-100) */
static inline double image_dist(const double *pos, const double *lat, int i, int j, int n1, int n2,
                                int n3, double xyz[3]) {
    double dr[3];
    for (int c = 0; c < 3; c++) {
        xyz[c] = pos[j * 3 + c] + (double)n1 * lat[0 + c] + (double)n2 * lat[3 + c] + (double)n3 * lat[6 + c];
        dr[c] = pos[i * 3 + c] - xyz[c];
    }
    return sqrt(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
}

/* ------------------------------------------------------------------------ */
/* neighbour table, reference order (j, n1, n2, n3): gap_calc.f90:89-120      */
/* ------------------------------------------------------------------------ */
typedef struct {
    int na, cap;
    int *count;    /* [na] */
    double *xyz;   /* [na][cap][3] */
    double *dis;   /* [na][cap] */
    double *wgt;   /* [na][cap] */
    int *idx;      /* [na][cap] neighbour atom index (0-based) */
    int *shift;    /* [na][cap][3] */
    long nwarn;    /* pairs closer than 0.5 A (the reference only prints) */
} nbtab;

static void nbtab_free(nbtab *t) {
    free(t->count); free(t->xyz); free(t->dis); free(t->wgt); free(t->idx); free(t->shift);
}

/* returns 0, or -1 if an atom has more than `cap` neighbours (reference: stop) */
static int nbtab_build(nbtab *t, int na, const double *lat, const double *pos, const double *weights,
                       double rcut, int cap, int c0, int c1) {
    memset(t, 0, sizeof *t);
    t->na = na; t->cap = cap;
    size_t nc = (size_t)(c1 - c0);
    t->count = (int *)calloc(nc + 1, sizeof(int));
    t->xyz = (double *)malloc(sizeof(double) * 3 * nc * cap + 8);
    t->dis = (double *)malloc(sizeof(double) * nc * cap + 8);
    t->wgt = (double *)malloc(sizeof(double) * nc * cap + 8);
    t->idx = (int *)malloc(sizeof(int) * nc * cap + 8);
    t->shift = (int *)malloc(sizeof(int) * 3 * nc * cap + 8);
    int nabc[3];
    image_range(lat, rcut, nabc);
    for (int i = c0; i < c1; i++) {
        size_t row = (size_t)(i - c0) * cap;
        int cnt = 0;
        for (int j = 0; j < na; j++)
            for (int n1 = -nabc[0]; n1 <= nabc[0]; n1++)
                for (int n2 = -nabc[1]; n2 <= nabc[1]; n2++)
                    for (int n3 = -nabc[2]; n3 <= nabc[2]; n3++) {
                        if (n1 == 0 && n2 == 0 && n3 == 0 && i == j) continue;
                        double xyz[3];
                        double dis = image_dist(pos, lat, i, j, n1, n2, n3, xyz);
                        if (dis > rcut) continue;
                        if (dis < RMIN_WARN) t->nwarn++;
                        if (cnt >= cap) return -1;
                        size_t s = row + cnt;
                        t->xyz[3 * s] = xyz[0]; t->xyz[3 * s + 1] = xyz[1]; t->xyz[3 * s + 2] = xyz[2];
                        t->dis[s] = dis; t->wgt[s] = weights ? weights[j] : 0.0; t->idx[s] = j;
                        t->shift[3 * s] = n1; t->shift[3 * s + 1] = n2; t->shift[3 * s + 2] = n3;
                        cnt++;
                    }
        t->count[i - c0] = cnt;
    }
    return 0;
}

/* Export the neighbour table (tests: pair-set exactness).  Outputs are sized
 * [na][cap]; returns max count, or -1 on overflow. */
int gapo_neighbors(int na, const double *lat, const double *pos, double rcut, int cap, int *count,
                   int *idx, int *shift, double *dis) {
    nbtab t;
    if (nbtab_build(&t, na, lat, pos, NULL, rcut, cap, 0, na)) { nbtab_free(&t); return -1; }
    int mx = 0;
    for (int i = 0; i < na; i++) {
        count[i] = t.count[i];
        if (t.count[i] > mx) mx = t.count[i];
    }
    for (int i = 0; i < na; i++)
        for (int s = 0; s < t.count[i]; s++) {
            size_t q = (size_t)i * cap + s;
            idx[q] = t.idx[q]; dis[q] = t.dis[q];
            shift[3 * q] = t.shift[3 * q]; shift[3 * q + 1] = t.shift[3 * q + 1]; shift[3 * q + 2] = t.shift[3 * q + 2];
        }
    nbtab_free(&t);
    return mx;
}

/* get_bond.f90:30-78: global minimum image distance within rcut, initial 10.0 */
double gapo_get_bond(int na, const double *lat, const double *pos, double rcut) {
    int nabc[3];
    image_range(lat, rcut, nabc);
    double min_bond = 10.0;
    for (int i = 0; i < na; i++)
        for (int j = 0; j < na; j++)
            for (int n1 = -nabc[0]; n1 <= nabc[0]; n1++)
                for (int n2 = -nabc[1]; n2 <= nabc[1]; n2++)
                    for (int n3 = -nabc[2]; n3 <= nabc[2]; n3++) {
                        if (n1 == 0 && n2 == 0 && n3 == 0 && i == j) continue;
                        double xyz[3];
                        double dis = image_dist(pos, lat, i, j, n1, n2, n3, xyz);
                        if (dis > rcut) continue;
                        if (dis < min_bond) min_bond = dis;
                    }
    return min_bond;
}

static int species_weights(const gapo_params *p, int na, const int *species, double *weights) {
    /* gap_calc.f90:75-83.  A species absent from the file leaves the weight
     * uninitialised in the reference; here it is an error (-2). */
    for (int j = 0; j < na; j++) {
        int found = 0;
        for (int s = 0; s < p->nspecies; s++)
            if (species[j] == p->z[s]) { weights[j] = p->w[s]; found = 1; }
        if (!found) return -2;
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* CAR2ACSF, dense: wacsf.f90:65-795                                          */
/*   xx[(i-c0)*D + k], strs[((i-c0)*D + k)*9 + a*3 + b] = strs(a+1,b+1,k+1,i+1)  */
/*   dxdy[(((n-c0)*na + i)*3 + c)*D + k] = dxdy(k+1, n+1(centre), i+1(atom), c+1) */
/* ------------------------------------------------------------------------ */
typedef struct {
    int na, D, c0;
    double *xx, *dxdy, *strs;
} dense_out;

static inline double *DX(dense_out *o, int centre, int atom, int c) {
    return o->dxdy + ((((size_t)(centre - o->c0) * o->na + atom) * 3 + c) * o->D);
}
static inline double *ST(dense_out *o, int centre, int k) {
    return o->strs + (((size_t)(centre - o->c0) * o->D + k) * 9);
}

static void radial_dense(const gapo_params *p, int ii, int shifted, const double *pos, const nbtab *nb,
                         int c0, int c1, int lgrad, dense_out *o) {
    const int nnn = p->nsf;
    const double cutoff = p->cutoff[ii];
    const double alpha = shifted ? 4.0 : p->alpha[ii];  /* wacsf.f90:439 */
    const double rshift = shifted ? p->alpha[ii] : 0.0; /* wacsf.f90:438 */
    for (int i = c0; i < c1; i++) {
        size_t row = (size_t)(i - c0) * nb->cap;
        double *xx = o->xx + (size_t)(i - c0) * o->D;
        for (int s = 0; s < nb->count[i - c0]; s++) {
            double rij = nb->dis[row + s];
            if (rij > cutoff) continue;
            const double *xyz = nb->xyz + 3 * (row + s);
            double w = nb->wgt[row + s];
            int n = nb->idx[row + s];
            double fc = 0.5 * (cos(PI_REF * rij / cutoff) + 1.0);
            double ex = shifted ? exp(-1.0 * alpha * ((rij - rshift) * (rij - rshift)))
                                : exp(-1.0 * alpha * (rij * rij));
            xx[ii] += ex * fc;
            xx[ii + nnn] += ex * fc * w;
            if (!lgrad) continue;
            double tfc = 0.5 * (-sin(PI_REF * rij / cutoff)) * (PI_REF / cutoff);
            double delta[3];
            for (int d = 0; d < 3; d++) delta[d] = -1.0 * (pos[i * 3 + d] - xyz[d]);
            for (int d = 0; d < 3; d++) {
                double drdi = -1.0 * delta[d] / rij, drdj = -1.0 * drdi;
                double dfdi = tfc * drdi, dfdj = -1.0 * dfdi;
                double gi, gj;
                if (!shifted) { /* wacsf.f90:107-113 */
                    double t1 = -2.0 * alpha * rij * ex * fc, t2 = ex;
                    gi = drdi * t1 + t2 * dfdi;
                    gj = drdj * t1 + t2 * dfdj;
                } else { /* wacsf.f90:469-479 */
                    double t1 = -2.0 * alpha * (rij - rshift), t2 = ex;
                    gi = t1 * drdi * t2 * fc + t2 * dfdi;
                    gj = t1 * drdj * t2 * fc + t2 * dfdj;
                }
                double gjw = gj * w;
                DX(o, i, i, d)[ii] += gi;
                DX(o, i, i, d)[ii + nnn] += gi * w;
                DX(o, i, n, d)[ii] += gj;
                DX(o, i, n, d)[ii + nnn] += gjw;
                for (int a = 0; a < 3; a++) {
                    ST(o, i, ii)[a * 3 + d] += delta[a] * gj;
                    ST(o, i, ii + nnn)[a * 3 + d] += delta[a] * gjw;
                }
            }
        }
    }
}

/* types 2 (lambda=+1) and 4 (lambda=-1): wacsf.f90:169-432, 534-791 */
/* Export of the triplet set of one centre for one cutoff (tests: triplet-set exactness): the
 * (js, ks) slot pairs the loops of wacsf.f90:177-244 keep -- "k_neighbor > j_neighbor" in list order,
 * "rij .gt. cutoff" / "rik .gt. cutoff" / "rjk .gt. cutoff" cycle -- with the reference's arithmetic.
 * pairs: [cap][2]; returns the number of kept pairs (all of them are counted, the first cap stored),
 * or -1 when the neighbour table overflows. */
int gapo_triplets(int na, const double *lat, const double *pos, double rcut, int centre, double cutoff, int cap, int *pairs) {
    nbtab t;
    if (nbtab_build(&t, na, lat, pos, NULL, rcut, MAX_NEIGHBOR_REF, centre, centre + 1)) { nbtab_free(&t); return -1; }
    int cnt = t.count[0], n = 0;
    for (int js = 0; js < cnt; js++) {
        double rij = t.dis[js];
        if (rij > cutoff) continue;
        const double *xj = t.xyz + 3 * js;
        for (int ks = js + 1; ks < cnt; ks++) {
            double rik = t.dis[ks];
            if (rik > cutoff) continue;
            const double *xk = t.xyz + 3 * ks;
            double rjk = (xj[0] - xk[0]) * (xj[0] - xk[0]) + (xj[1] - xk[1]) * (xj[1] - xk[1]) +
                         (xj[2] - xk[2]) * (xj[2] - xk[2]);
            rjk = sqrt(rjk);
            if (rjk > cutoff) continue;
            if (n < cap) { pairs[2 * n] = js; pairs[2 * n + 1] = ks; }
            n++;
        }
    }
    nbtab_free(&t);
    return n;
}

static void angular_dense(const gapo_params *p, int ii, double lambda, const double *pos, const nbtab *nb,
                          int c0, int c1, int lgrad, dense_out *o, double *ntrip_out) {
    const int nnn = p->nsf;
    const double cutoff = p->cutoff[ii], alpha = p->alpha[ii];
    double ntrip = 0;
    for (int i = c0; i < c1; i++) {
        size_t row = (size_t)(i - c0) * nb->cap;
        double *xx = o->xx + (size_t)(i - c0) * o->D;
        int cnt = nb->count[i - c0];
        for (int js = 0; js < cnt; js++) {
            double rij = nb->dis[row + js];
            if (rij > cutoff) continue;
            const double *xj = nb->xyz + 3 * (row + js);
            double fij = 0.5 * (cos(PI_REF * rij / cutoff) + 1.0);
            double wj = nb->wgt[row + js];
            int n = nb->idx[row + js];
            double dj[3], rij_i[3], rij_j[3], fij_i[3], fij_j[3];
            if (lgrad) {
                double t = 0.5 * (-sin(PI_REF * rij / cutoff)) * (PI_REF / cutoff);
                for (int d = 0; d < 3; d++) {
                    dj[d] = -1.0 * (pos[i * 3 + d] - xj[d]);
                    rij_i[d] = -1.0 * dj[d] / rij; rij_j[d] = -1.0 * rij_i[d];
                    fij_i[d] = t * rij_i[d]; fij_j[d] = -1.0 * fij_i[d];
                }
            }
            for (int ks = js + 1; ks < cnt; ks++) { /* k_neighbor > j_neighbor, wacsf.f90:210 */
                double rik = nb->dis[row + ks];
                if (rik > cutoff) continue;
                const double *xk = nb->xyz + 3 * (row + ks);
                double wk = nb->wgt[row + ks];
                int m = nb->idx[row + ks];
                double fik = 0.5 * (cos(PI_REF * rik / cutoff) + 1.0);
                double rjk = (xj[0] - xk[0]) * (xj[0] - xk[0]) + (xj[1] - xk[1]) * (xj[1] - xk[1]) +
                             (xj[2] - xk[2]) * (xj[2] - xk[2]);
                rjk = sqrt(rjk);
                if (rjk > cutoff) continue; /* wacsf.f90:244 */
                ntrip += 1;
                double fjk = 0.5 * (cos(PI_REF * rjk / cutoff) + 1.0);
                double f = rjk * rjk - rij * rij - rik * rik;
                double g = -2.0 * rij * rik;
                double costheta = f / g;
                costheta = 1.0 + lambda * costheta;
                double expxyz = exp(-alpha * (rij * rij + rik * rik + rjk * rjk));
                double val = costheta * expxyz * fij * fik * fjk;
                xx[ii] += val;
                xx[ii + nnn] += val * wj * wk;
                if (!lgrad) continue;
                double tk = 0.5 * (-sin(PI_REF * rik / cutoff)) * (PI_REF / cutoff);
                double tjk = 0.5 * (-sin(PI_REF * rjk / cutoff)) * (PI_REF / cutoff);
                double ig2 = 1.0 / (g * g);
                double te = -alpha * 2.0 * expxyz;
                double dk[3];
                for (int d = 0; d < 3; d++) dk[d] = -1.0 * (pos[i * 3 + d] - xk[d]);
                for (int d = 0; d < 3; d++) {
                    double rik_i = -dk[d] / rik, rik_k = -1.0 * rik_i;
                    double fik_i = tk * rik_i, fik_k = -1.0 * fik_i;
                    double rjk_j = (xj[d] - xk[d]) / rjk, rjk_k = -1.0 * rjk_j;
                    double fjk_j = tjk * rjk_j, fjk_k = -1.0 * fjk_j;
                    /* f, g derivatives: wacsf.f90:278-300 */
                    double df_i = -2.0 * rij * rij_i[d] - 2.0 * rik * rik_i;
                    double df_j = 2.0 * rjk * rjk_j - 2.0 * rij * rij_j[d];
                    double df_k = 2.0 * rjk * rjk_k - 2.0 * rik * rik_k;
                    double dg_i = -2.0 * (rij_i[d] * rik + rij * rik_i);
                    double dg_j = -2.0 * rij_j[d] * rik;
                    double dg_k = -2.0 * rij * rik_k;
                    double dc_i = lambda * (df_i * g - f * dg_i) * ig2;
                    double dc_j = lambda * (df_j * g - f * dg_j) * ig2;
                    double dc_k = lambda * (df_k * g - f * dg_k) * ig2;
                    /* wacsf.f90:315-324 (the i/j/k zero terms are dropped) */
                    double de_i = (rij * rij_i[d] + rik * rik_i) * te;
                    double de_j = (rij * rij_j[d] + rjk * rjk_j) * te;
                    double de_k = (rik * rik_k + rjk * rjk_k) * te;
                    /* wacsf.f90:331-345 */
                    double t1 = dc_i * expxyz * fij * fik * fjk + costheta * de_i * fij * fik * fjk +
                                costheta * expxyz * fij_i[d] * fik * fjk + costheta * expxyz * fij * fik_i * fjk;
                    double t2 = dc_j * expxyz * fij * fik * fjk + costheta * de_j * fij * fik * fjk +
                                costheta * expxyz * fij_j[d] * fik * fjk + costheta * expxyz * fij * fik * fjk_j;
                    double t3 = dc_k * expxyz * fij * fik * fjk + costheta * de_k * fij * fik * fjk +
                                costheta * expxyz * fij * fik_k * fjk + costheta * expxyz * fij * fik * fjk_k;
                    double t4 = t1 * wj * wk, t5 = t2 * wj * wk, t6 = t3 * wj * wk;
                    DX(o, i, i, d)[ii] += t1; DX(o, i, n, d)[ii] += t2; DX(o, i, m, d)[ii] += t3;
                    DX(o, i, i, d)[ii + nnn] += t4; DX(o, i, n, d)[ii + nnn] += t5; DX(o, i, m, d)[ii + nnn] += t6;
                    for (int a = 0; a < 3; a++) {
                        ST(o, i, ii)[a * 3 + d] += dj[a] * t2 + dk[a] * t3;
                        ST(o, i, ii + nnn)[a * 3 + d] += dj[a] * t5 + dk[a] * t6;
                    }
                }
            }
        }
    }
    if (ntrip_out) *ntrip_out += ntrip;
}

/* GET_COV (gap_calc.f90:268-288), energy (:152-154), dE/dG (:160-166).
 * kk[(i)*D+k]; outputs e[i], dedg[i*D+k] */
static void gpr_block(const gapo_params *p, int n, const double *kk, double *eatom, double *dedg,
                      int lgrad) {
    const int M = p->nsparse, D = p->des_len;
    double *ckm = (double *)malloc(sizeof(double) * (size_t)M + 8);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < M; j++) {
            double temp = 0.0;
            for (int k = 0; k < D; k++) {
                double q = (kk[(size_t)i * D + k] - p->mm[(size_t)j * D + k]) / p->theta[k];
                temp = temp + q * q;
            }
            ckm[j] = DELTA_K * exp(-0.5 * temp);
        }
        double e = 0.0;
        for (int j = 0; j < M; j++) e += ckm[j] * p->coeff[j];
        eatom[i] = e;
        if (dedg) {
            for (int k = 0; k < D; k++) {
                double acc = 0.0;
                if (lgrad || 1) /* the reference computes dedg regardless of lgrad */
                    for (int j = 0; j < M; j++)
                        acc = acc - 1.0 * (kk[(size_t)i * D + k] - p->mm[(size_t)j * D + k]) /
                                        (p->theta[k] * p->theta[k]) * ckm[j] * p->coeff[j];
                dedg[(size_t)i * D + k] = acc;
            }
        }
    }
    free(ckm);
}

/*
 * Dense transliteration of FGAP_CALC (gap_calc.f90:67-226) restricted to the
 * centre atoms [c0,c1) (c0=0,c1=na is the reference).  With a sub-range the
 * energy is the sum over those centres and force/stress hold only their
 * contributions: used as a bounded timing sample of the reference algorithm.
 * Optional outputs (may be NULL): xx_out[(c1-c0)*D], dedg_out[(c1-c0)*D],
 * eatom_out[c1-c0], stats[8] = {sum P, sum kept triplet-SF evaluations, n warnings}.
 * Returns 0; -1 neighbour overflow (>1000, reference stops); -2 species missing
 * from the file; -3 des_len != 2*nsf; -4 out of memory.
 */
int gapo_calc_dense(const gapo_params *p, int na, const int *species, const double *lat, const double *pos,
                    double rcut, int lgrad, int c0, int c1, double *ene, double *force, double *stress,
                    double *xx_out, double *dedg_out, double *eatom_out, double *stats) {
    const int D = p->des_len;
    if (D != 2 * p->nsf) return -3;
    if (c0 < 0) c0 = 0;
    if (c1 > na || c1 <= 0) c1 = na;
    int nc = c1 - c0;
    double *weights = (double *)malloc(sizeof(double) * (size_t)na + 8);
    int rc = species_weights(p, na, species, weights);
    if (rc) { free(weights); return rc; }
    nbtab nb;
    if (nbtab_build(&nb, na, lat, pos, weights, rcut, MAX_NEIGHBOR_REF, c0, c1)) {
        nbtab_free(&nb); free(weights); return -1;
    }
    dense_out o; o.na = na; o.D = D; o.c0 = c0;
    o.xx = (double *)calloc((size_t)nc * D + 1, sizeof(double));
    o.dxdy = (double *)calloc((size_t)nc * na * 3 * D + 1, sizeof(double)); /* the O(N^2) array */
    o.strs = (double *)calloc((size_t)nc * D * 9 + 1, sizeof(double));
    if (!o.xx || !o.dxdy || !o.strs) { free(o.xx); free(o.dxdy); free(o.strs); nbtab_free(&nb); free(weights); return -4; }
    double ntrip = 0;
    for (int ii = 0; ii < p->nsf; ii++) {
        switch (p->ntype[ii]) {
            case 1: radial_dense(p, ii, 0, pos, &nb, c0, c1, lgrad, &o); break;
            case 2: angular_dense(p, ii, +1.0, pos, &nb, c0, c1, lgrad, &o, &ntrip); break;
            case 3: radial_dense(p, ii, 1, pos, &nb, c0, c1, lgrad, &o); break;
            case 4: angular_dense(p, ii, -1.0, pos, &nb, c0, c1, lgrad, &o, &ntrip); break;
            default: break; /* reference prints 'Unknown function type' and continues */
        }
    }
    double *eatom = (double *)malloc(sizeof(double) * (size_t)nc + 8);
    double *dedg = (double *)malloc(sizeof(double) * (size_t)nc * D + 8);
    gpr_block(p, nc, o.xx, eatom, dedg, lgrad);
    double e = 0.0;
    for (int i = 0; i < nc; i++) e += eatom[i];
    *ene = e;
    /* force: gap_calc.f90:177-185 */
    for (int i = 0; i < na; i++)
        for (int c = 0; c < 3; c++) {
            double acc = 0.0;
            if (lgrad)
                for (int n = c0; n < c1; n++) {
                    const double *dx = DX(&o, n, i, c);
                    const double *dg = dedg + (size_t)(n - c0) * D;
                    for (int k = 0; k < D; k++) acc = acc - dg[k] * dx[k];
                }
            force[i * 3 + c] = acc;
        }
    /* stress: gap_calc.f90:189-203, 221-226 */
    double s6[6];
    int k1 = 0;
    for (int a = 0; a < 3; a++)
        for (int b = a; b < 3; b++) {
            double acc = 0.0;
            for (int n = c0; n < c1; n++)
                for (int k = 0; k < D; k++) acc = acc - dedg[(size_t)(n - c0) * D + k] * ST(&o, n, k)[a * 3 + b];
            s6[k1++] = acc;
        }
    double volume = fabs(det3(lat));
    for (int q = 0; q < 6; q++) s6[q] = s6[q] * (1.0 / GPA2EVPANG) / volume;
    stress[0] = s6[0]; stress[1] = s6[3]; stress[2] = s6[5];
    stress[3] = s6[1]; stress[4] = s6[4]; stress[5] = s6[2];
    if (xx_out) memcpy(xx_out, o.xx, sizeof(double) * (size_t)nc * D);
    if (dedg_out) memcpy(dedg_out, dedg, sizeof(double) * (size_t)nc * D);
    if (eatom_out) memcpy(eatom_out, eatom, sizeof(double) * (size_t)nc);
    if (stats) {
        double sp = 0; for (int i = 0; i < nc; i++) sp += nb.count[i];
        stats[0] = sp; stats[1] = ntrip; stats[2] = (double)nb.nwarn;
    }
    free(eatom); free(dedg); free(o.xx); free(o.dxdy); free(o.strs); nbtab_free(&nb); free(weights);
    return 0;
}

/* Dense descriptor block as exposed by the reference's public car2acsf
 * (wacsf.f90:2-12): xx[na*D], dxdy (layout above with c0=0), strs. */
int gapo_car2acsf_dense(const gapo_params *p, int na, const int *species, const double *lat,
                        const double *pos, double rcut, int lgrad, double *xx, double *dxdy, double *strs) {
    const int D = p->des_len;
    if (D != 2 * p->nsf) return -3;
    double *weights = (double *)malloc(sizeof(double) * (size_t)na + 8);
    int rc = species_weights(p, na, species, weights);
    if (rc) { free(weights); return rc; }
    nbtab nb;
    if (nbtab_build(&nb, na, lat, pos, weights, rcut, MAX_NEIGHBOR_REF, 0, na)) { nbtab_free(&nb); free(weights); return -1; }
    dense_out o; o.na = na; o.D = D; o.c0 = 0; o.xx = xx; o.dxdy = dxdy; o.strs = strs;
    memset(xx, 0, sizeof(double) * (size_t)na * D);
    memset(dxdy, 0, sizeof(double) * (size_t)na * na * 3 * D);
    memset(strs, 0, sizeof(double) * (size_t)na * D * 9);
    for (int ii = 0; ii < p->nsf; ii++) {
        switch (p->ntype[ii]) {
            case 1: radial_dense(p, ii, 0, pos, &nb, 0, na, lgrad, &o); break;
            case 2: angular_dense(p, ii, +1.0, pos, &nb, 0, na, lgrad, &o, NULL); break;
            case 3: radial_dense(p, ii, 1, pos, &nb, 0, na, lgrad, &o); break;
            case 4: angular_dense(p, ii, -1.0, pos, &nb, 0, na, lgrad, &o, NULL); break;
            default: break;
        }
    }
    nbtab_free(&nb); free(weights);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Sparse O(N) form.  Same inclusion tests and the same per-term arithmetic;  */
/* the chain rule dE/dG * dG/dr is applied per neighbour    */
/* instead of through the dense dxdy array.  The neighbour list of each centre */
/* is found with a cell list but kept in the reference (j,n1,n2,n3) order.     */
/* ------------------------------------------------------------------------ */
typedef struct { int j, n1, n2, n3; double xyz[3], dis, w; } nbent;

static int nbent_cmp(const void *a, const void *b) {
    const nbent *x = (const nbent *)a, *y = (const nbent *)b;
    if (x->j != y->j) return x->j < y->j ? -1 : 1;
    if (x->n1 != y->n1) return x->n1 < y->n1 ? -1 : 1;
    if (x->n2 != y->n2) return x->n2 < y->n2 ? -1 : 1;
    if (x->n3 != y->n3) return x->n3 < y->n3 ? -1 : 1;
    return 0;
}

typedef struct {
    int nb[3];       /* bins per lattice direction */
    int *head, *next; /* linked cells over atoms */
    int *bin;         /* [na][3] bin coordinates */
    int *wrap;        /* [na][3] integer lattice offsets removed when wrapping */
    double inv[9];    /* inverse lattice: frac = pos * inv  (row-vector convention) */
} cellgrid;

static void invert3(const double m[9], double inv[9]) {
    double d = det3(m);
    inv[0] = (m[4] * m[8] - m[5] * m[7]) / d; inv[1] = (m[2] * m[7] - m[1] * m[8]) / d; inv[2] = (m[1] * m[5] - m[2] * m[4]) / d;
    inv[3] = (m[5] * m[6] - m[3] * m[8]) / d; inv[4] = (m[0] * m[8] - m[2] * m[6]) / d; inv[5] = (m[2] * m[3] - m[0] * m[5]) / d;
    inv[6] = (m[3] * m[7] - m[4] * m[6]) / d; inv[7] = (m[1] * m[6] - m[0] * m[7]) / d; inv[8] = (m[0] * m[4] - m[1] * m[3]) / d;
}

static void cellgrid_build(cellgrid *g, int na, const double *lat, const double *pos, double rcut) {
    invert3(lat, g->inv);
    /* interplanar spacing d_c = 1/|column c of inv| */
    for (int c = 0; c < 3; c++) {
        double len = sqrt(g->inv[0 + c] * g->inv[0 + c] + g->inv[3 + c] * g->inv[3 + c] + g->inv[6 + c] * g->inv[6 + c]);
        double dsp = 1.0 / len;
        int nbin = (int)floor(dsp / (rcut * (1.0 + 1e-9)));
        g->nb[c] = nbin < 1 ? 1 : nbin;
    }
    int ntot = g->nb[0] * g->nb[1] * g->nb[2];
    g->head = (int *)malloc(sizeof(int) * (size_t)ntot);
    for (int b = 0; b < ntot; b++) g->head[b] = -1;
    g->next = (int *)malloc(sizeof(int) * (size_t)na + 8);
    g->bin = (int *)malloc(sizeof(int) * 3 * (size_t)na + 8);
    g->wrap = (int *)malloc(sizeof(int) * 3 * (size_t)na + 8);
    for (int i = na - 1; i >= 0; i--) {
        int b[3];
        for (int c = 0; c < 3; c++) {
            double f = pos[i * 3] * g->inv[0 + c] + pos[i * 3 + 1] * g->inv[3 + c] + pos[i * 3 + 2] * g->inv[6 + c];
            double fl = floor(f);
            g->wrap[i * 3 + c] = (int)fl;
            int bc = (int)((f - fl) * g->nb[c]);
            if (bc >= g->nb[c]) bc = g->nb[c] - 1;
            if (bc < 0) bc = 0;
            b[c] = bc; g->bin[i * 3 + c] = bc;
        }
        int id = (b[0] * g->nb[1] + b[1]) * g->nb[2] + b[2];
        g->next[i] = g->head[id]; g->head[id] = i;
    }
}

static void cellgrid_free(cellgrid *g) { free(g->head); free(g->next); free(g->bin); free(g->wrap); }

static inline int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

/* neighbours of centre i, reference inclusion test and reference order */
static int sparse_neighbors(const cellgrid *g, int na, const double *lat, const double *pos,
                            const double *weights, double rcut, const int nabc[3], int i, nbent *out, int cap) {
    int cnt = 0, m[3];
    for (int c = 0; c < 3; c++) {
        /* bins are >= rcut thick when nb>1; with a single bin, scan nabc(+1) images */
        m[c] = (g->nb[c] > 1) ? 1 : nabc[c] + 1;
    }
    (void)na;
    for (int d0 = -m[0]; d0 <= m[0]; d0++)
        for (int d1 = -m[1]; d1 <= m[1]; d1++)
            for (int d2 = -m[2]; d2 <= m[2]; d2++) {
                int dd[3] = {d0, d1, d2}, bb[3], sh[3];
                for (int c = 0; c < 3; c++) {
                    int t = g->bin[i * 3 + c] + dd[c];
                    sh[c] = floordiv(t, g->nb[c]);
                    bb[c] = t - sh[c] * g->nb[c];
                }
                int id = (bb[0] * g->nb[1] + bb[1]) * g->nb[2] + bb[2];
                for (int j = g->head[id]; j >= 0; j = g->next[j]) {
                    /* shift in wrapped coordinates -> shift of the original coordinates */
                    int n1 = sh[0] - g->wrap[j * 3] + g->wrap[i * 3];
                    int n2 = sh[1] - g->wrap[j * 3 + 1] + g->wrap[i * 3 + 1];
                    int n3 = sh[2] - g->wrap[j * 3 + 2] + g->wrap[i * 3 + 2];
                    if (n1 == 0 && n2 == 0 && n3 == 0 && i == j) continue;
                    /* the reference only looks at |n| <= nabc (gap_calc.f90:94-96) */
                    if (abs(n1) > nabc[0] || abs(n2) > nabc[1] || abs(n3) > nabc[2]) continue;
                    double xyz[3];
                    double dis = image_dist(pos, lat, i, j, n1, n2, n3, xyz);
                    if (dis > rcut) continue;
                    if (cnt >= cap) return -1;
                    nbent *e = &out[cnt++];
                    e->j = j; e->n1 = n1; e->n2 = n2; e->n3 = n3;
                    e->xyz[0] = xyz[0]; e->xyz[1] = xyz[1]; e->xyz[2] = xyz[2];
                    e->dis = dis; e->w = weights ? weights[j] : 0.0;
                }
            }
    qsort(out, (size_t)cnt, sizeof(nbent), nbent_cmp);
    return cnt;
}

/* Neighbour sets through the cell list (tests: must equal gapo_neighbors). */
int gapo_neighbors_sparse(int na, const double *lat, const double *pos, double rcut, int cap, int *count,
                          int *idx, int *shift, double *dis) {
    int nabc[3];
    image_range(lat, rcut, nabc);
    cellgrid g;
    cellgrid_build(&g, na, lat, pos, rcut);
    nbent *buf = (nbent *)malloc(sizeof(nbent) * (size_t)cap + 8);
    int mx = 0;
    for (int i = 0; i < na; i++) {
        int cnt = sparse_neighbors(&g, na, lat, pos, NULL, rcut, nabc, i, buf, cap);
        if (cnt < 0) { mx = -1; break; }
        count[i] = cnt;
        if (cnt > mx) mx = cnt;
        for (int s = 0; s < cnt; s++) {
            size_t q = (size_t)i * cap + s;
            idx[q] = buf[s].j; shift[3 * q] = buf[s].n1; shift[3 * q + 1] = buf[s].n2; shift[3 * q + 2] = buf[s].n3;
            dis[q] = buf[s].dis;
        }
    }
    free(buf); cellgrid_free(&g);
    return mx;
}

/* forward descriptors of one centre from its neighbour list */
static void sparse_forward(const gapo_params *p, const nbent *nb, int cnt, double *xx, double *wc) {
    const int nnn = p->nsf;
    for (int k = 0; k < 2 * nnn; k++) xx[k] = 0.0;
    for (int ii = 0; ii < nnn; ii++) {
        const double cutoff = p->cutoff[ii];
        int t = p->ntype[ii];
        if (t == 1 || t == 3) {
            double alpha = (t == 3) ? 4.0 : p->alpha[ii], rs = (t == 3) ? p->alpha[ii] : 0.0;
            for (int s = 0; s < cnt; s++) {
                double rij = nb[s].dis;
                if (rij > cutoff) continue;
                double fc = 0.5 * (cos(PI_REF * rij / cutoff) + 1.0);
                double ex = (t == 3) ? exp(-1.0 * alpha * ((rij - rs) * (rij - rs))) : exp(-1.0 * alpha * (rij * rij));
                xx[ii] += ex * fc;
                xx[ii + nnn] += ex * fc * nb[s].w;
                if (wc) wc[1] += 1;
            }
        } else if (t == 2 || t == 4) {
            double lambda = (t == 2) ? 1.0 : -1.0, alpha = p->alpha[ii];
            for (int js = 0; js < cnt; js++) {
                double rij = nb[js].dis;
                if (rij > cutoff) continue;
                double fij = 0.5 * (cos(PI_REF * rij / cutoff) + 1.0);
                for (int ks = js + 1; ks < cnt; ks++) {
                    double rik = nb[ks].dis;
                    if (rik > cutoff) continue;
                    if (wc) wc[2] += 1;
                    const double *xj = nb[js].xyz, *xk = nb[ks].xyz;
                    double rjk = (xj[0] - xk[0]) * (xj[0] - xk[0]) + (xj[1] - xk[1]) * (xj[1] - xk[1]) +
                                 (xj[2] - xk[2]) * (xj[2] - xk[2]);
                    rjk = sqrt(rjk);
                    if (rjk > cutoff) continue;
                    if (wc) wc[3] += 1;
                    double fik = 0.5 * (cos(PI_REF * rik / cutoff) + 1.0);
                    double fjk = 0.5 * (cos(PI_REF * rjk / cutoff) + 1.0);
                    double f = rjk * rjk - rij * rij - rik * rik, g = -2.0 * rij * rik;
                    double ct = 1.0 + lambda * (f / g);
                    double ex = exp(-alpha * (rij * rij + rik * rik + rjk * rjk));
                    double val = ct * ex * fij * fik * fjk;
                    xx[ii] += val;
                    xx[ii + nnn] += val * nb[js].w * nb[ks].w;
                }
            }
        }
    }
}

/* backward of one centre: accumulates dE/dx of the centre (gi), of each
 * neighbour slot (gn[s][3]) and the centre's strs contraction vir[a][b]. */
static void sparse_backward(const gapo_params *p, const double *posi, const nbent *nb, int cnt,
                            const double *dedg, double *gi, double *gn, double *vir) {
    const int nnn = p->nsf;
    for (int ii = 0; ii < nnn; ii++) {
        const double cutoff = p->cutoff[ii];
        const double du = dedg[ii], dw = dedg[ii + nnn];
        int t = p->ntype[ii];
        if (t == 1 || t == 3) {
            double alpha = (t == 3) ? 4.0 : p->alpha[ii], rs = (t == 3) ? p->alpha[ii] : 0.0;
            for (int s = 0; s < cnt; s++) {
                double rij = nb[s].dis;
                if (rij > cutoff) continue;
                double fc = 0.5 * (cos(PI_REF * rij / cutoff) + 1.0);
                double dfc = 0.5 * (-sin(PI_REF * rij / cutoff)) * (PI_REF / cutoff);
                double ex, dg;
                if (t == 1) { ex = exp(-1.0 * alpha * (rij * rij)); dg = -2.0 * alpha * rij * ex * fc + ex * dfc; }
                else { ex = exp(-1.0 * alpha * ((rij - rs) * (rij - rs))); dg = -2.0 * alpha * (rij - rs) * ex * fc + ex * dfc; }
                double c = (du + nb[s].w * dw) * dg / rij; /* dE/dr_ij / r_ij */
                for (int d = 0; d < 3; d++) {
                    double del = nb[s].xyz[d] - posi[d];
                    double gj = c * del;
                    gn[3 * s + d] += gj; gi[d] -= gj;
                    for (int a = 0; a < 3; a++) vir[a * 3 + d] += (nb[s].xyz[a] - posi[a]) * gj;
                }
            }
        } else if (t == 2 || t == 4) {
            double lambda = (t == 2) ? 1.0 : -1.0, alpha = p->alpha[ii];
            for (int js = 0; js < cnt; js++) {
                double rij = nb[js].dis;
                if (rij > cutoff) continue;
                double fij = 0.5 * (cos(PI_REF * rij / cutoff) + 1.0);
                double dfij = 0.5 * (-sin(PI_REF * rij / cutoff)) * (PI_REF / cutoff);
                for (int ks = js + 1; ks < cnt; ks++) {
                    double rik = nb[ks].dis;
                    if (rik > cutoff) continue;
                    const double *xj = nb[js].xyz, *xk = nb[ks].xyz;
                    double rjk = (xj[0] - xk[0]) * (xj[0] - xk[0]) + (xj[1] - xk[1]) * (xj[1] - xk[1]) +
                                 (xj[2] - xk[2]) * (xj[2] - xk[2]);
                    rjk = sqrt(rjk);
                    if (rjk > cutoff) continue;
                    double fik = 0.5 * (cos(PI_REF * rik / cutoff) + 1.0);
                    double dfik = 0.5 * (-sin(PI_REF * rik / cutoff)) * (PI_REF / cutoff);
                    double fjk = 0.5 * (cos(PI_REF * rjk / cutoff) + 1.0);
                    double dfjk = 0.5 * (-sin(PI_REF * rjk / cutoff)) * (PI_REF / cutoff);
                    double f = rjk * rjk - rij * rij - rik * rik, g = -2.0 * rij * rik;
                    double cosv = f / g;
                    double A = 1.0 + lambda * cosv;
                    double ex = exp(-alpha * (rij * rij + rik * rik + rjk * rjk));
                    double phi = fij * fik * fjk;
                    double gam = du + nb[js].w * nb[ks].w * dw;
                    /* partials of v = A ex phi w.r.t. the three distances */
                    double vij = ex * (lambda * (1.0 / rik - cosv / rij) * phi - 2.0 * alpha * rij * A * phi + A * dfij * fik * fjk);
                    double vik = ex * (lambda * (1.0 / rij - cosv / rik) * phi - 2.0 * alpha * rik * A * phi + A * fij * dfik * fjk);
                    double vjk = ex * (lambda * (-rjk / (rij * rik)) * phi - 2.0 * alpha * rjk * A * phi + A * fij * fik * dfjk);
                    double cij = gam * vij / rij, cik = gam * vik / rik, cjk = gam * vjk / rjk;
                    for (int d = 0; d < 3; d++) {
                        double dj = xj[d] - posi[d], dk = xk[d] - posi[d], djk = xj[d] - xk[d];
                        double gj = cij * dj + cjk * djk;
                        double gk = cik * dk - cjk * djk;
                        gn[3 * js + d] += gj; gn[3 * ks + d] += gk; gi[d] -= gj + gk;
                        for (int a = 0; a < 3; a++)
                            vir[a * 3 + d] += (xj[a] - posi[a]) * gj + (xk[a] - posi[a]) * gk;
                    }
                }
            }
        }
    }
}

/*
 * Sparse evaluation.  stats (optional, [8]): {sum P, radial pair-SF evals,
 * candidate pair-SF tests, kept triplet-SF evals, 0...}.  Same return codes as
 * gapo_calc_dense; max_nb is the neighbour capacity per atom (reference: 1000).
 */
static int calc_sparse_range(const gapo_params *p, int na, const int *species, const double *lat, const double *pos,
                             double rcut, int lgrad, int max_nb, int c0, int c1, double *ene, double *force, double *stress,
                             double *xx_out, double *dedg_out, double *eatom_out, double *stats);

int gapo_calc_sparse(const gapo_params *p, int na, const int *species, const double *lat, const double *pos,
                     double rcut, int lgrad, int max_nb, double *ene, double *force, double *stress,
                     double *xx_out, double *dedg_out, double *eatom_out, double *stats) {
    return calc_sparse_range(p, na, species, lat, pos, rcut, lgrad, max_nb, 0, na, ene, force, stress, xx_out, dedg_out,
                             eatom_out, stats);
}

/* Centres [c0, c1) only: their energies, the forces they exert and their share of the stress (a
 * bounded sample of a structure too large to finish on the CPU in benchmark time: bench.py's
 * cpu_baseline on the 100k-atom cell).  Descriptor outputs are indexed by the full atom index. */
int gapo_calc_sparse_centres(const gapo_params *p, int na, const int *species, const double *lat, const double *pos,
                             double rcut, int lgrad, int max_nb, int c0, int c1, double *ene, double *force, double *stress) {
    if (c0 < 0) c0 = 0;
    if (c1 > na || c1 <= 0) c1 = na;
    return calc_sparse_range(p, na, species, lat, pos, rcut, lgrad, max_nb, c0, c1, ene, force, stress, NULL, NULL, NULL, NULL);
}

static int calc_sparse_range(const gapo_params *p, int na, const int *species, const double *lat, const double *pos,
                             double rcut, int lgrad, int max_nb, int c0, int c1, double *ene, double *force, double *stress,
                             double *xx_out, double *dedg_out, double *eatom_out, double *stats) {
    const int D = p->des_len;
    if (D != 2 * p->nsf) return -3;
    double *weights = (double *)malloc(sizeof(double) * (size_t)na + 8);
    int rc = species_weights(p, na, species, weights);
    if (rc) { free(weights); return rc; }
    int nabc[3];
    image_range(lat, rcut, nabc);
    cellgrid g;
    cellgrid_build(&g, na, lat, pos, rcut);
    nbent *nb = (nbent *)malloc(sizeof(nbent) * (size_t)max_nb + 8);
    double *xx = (double *)malloc(sizeof(double) * D + 8), *dedg = (double *)malloc(sizeof(double) * D + 8);
    double *gn = (double *)malloc(sizeof(double) * 3 * (size_t)max_nb + 8);
    double vir[9] = {0}, e = 0.0, wc[4] = {0, 0, 0, 0};
    for (int i = 0; i < 3 * na; i++) force[i] = 0.0;
    int ret = 0;
    for (int i = c0; i < c1; i++) {
        int cnt = sparse_neighbors(&g, na, lat, pos, weights, rcut, nabc, i, nb, max_nb);
        if (cnt < 0) { ret = -1; break; }
        wc[0] += cnt;
        sparse_forward(p, nb, cnt, xx, stats ? wc : NULL);
        double ei;
        gpr_block(p, 1, xx, &ei, dedg, lgrad);
        e += ei;
        if (xx_out) memcpy(xx_out + (size_t)i * D, xx, sizeof(double) * D);
        if (dedg_out) memcpy(dedg_out + (size_t)i * D, dedg, sizeof(double) * D);
        if (eatom_out) eatom_out[i] = ei;
        if (!lgrad) continue;
        double gi[3] = {0, 0, 0};
        for (int s = 0; s < 3 * cnt; s++) gn[s] = 0.0;
        sparse_backward(p, pos + 3 * i, nb, cnt, dedg, gi, gn, vir);
        for (int d = 0; d < 3; d++) force[i * 3 + d] -= gi[d];
        for (int s = 0; s < cnt; s++)
            for (int d = 0; d < 3; d++) force[nb[s].j * 3 + d] -= gn[3 * s + d];
    }
    *ene = e;
    double volume = fabs(det3(lat));
    double s6[6];
    int k1 = 0;
    for (int a = 0; a < 3; a++)
        for (int b = a; b < 3; b++) s6[k1++] = -vir[a * 3 + b] * (1.0 / GPA2EVPANG) / volume;
    stress[0] = s6[0]; stress[1] = s6[3]; stress[2] = s6[5];
    stress[3] = s6[1]; stress[4] = s6[4]; stress[5] = s6[2];
    if (stats) { stats[0] = wc[0]; stats[1] = wc[1]; stats[2] = wc[2]; stats[3] = wc[3]; }
    free(gn); free(xx); free(dedg); free(nb); cellgrid_free(&g); free(weights);
    return ret;
}
