// potential.cpp -- gap_parameters reader and evaluation plan (host only).
// Compile with -ffp-contract=off: make_cell() must reproduce the reference's
// image-window arithmetic (gap_calc.f90:85-88, :240-266) operation by operation.
#include "potential.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace gapcu {

namespace {

// Fortran list-directed input over records (lines): each READ statement starts
// on a fresh record, consumes blank/comma separated items over as many records
// as it needs and drops the remainder of the last one.
class ListReader {
  public:
    explicit ListReader(const std::string &path) {
        std::ifstream in(path, std::ios::binary);
        if (!in) throw std::runtime_error("gap_parameters file does not exist!");
        std::string line;
        while (std::getline(in, line)) lines_.push_back(line);
    }
    void read(int n, double *out, const char *what) {
        int got = 0;
        do {
            if (next_ >= lines_.size())
                throw std::runtime_error(std::string("gap_parameters: end of file while reading ") + what);
            std::string rec = lines_[next_++];
            for (char &ch : rec) {
                if (ch == ',') ch = ' ';
                if (ch == 'd' || ch == 'D') ch = 'e';
            }
            const char *p = rec.c_str();
            while (got < n) {
                char *end = nullptr;
                double v = std::strtod(p, &end);
                if (end == p) break;
                out[got++] = v;
                p = end;
            }
            if (got < n) {  // anything left on the record that is not a number is an error
                while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
                if (*p) throw std::runtime_error(std::string("gap_parameters: bad value in ") + what);
            }
        } while (got < n);
    }
    void skip(const char *what) {
        if (next_ >= lines_.size())
            throw std::runtime_error(std::string("gap_parameters: end of file while skipping ") + what);
        ++next_;
    }

  private:
    std::vector<std::string> lines_;
    size_t next_ = 0;
};

}  // namespace

PotentialFile read_gap_parameters(const std::string &path) {
    ListReader rd(path);
    PotentialFile pf;
    double v[3];
    rd.read(1, v, "nspecies");
    int nspecies = (int)v[0];
    if (nspecies < 0 || nspecies > 200) throw std::runtime_error("gap_parameters: bad nspecies");
    for (int i = 0; i < nspecies; i++) {
        rd.read(2, v, "species weight");
        pf.z.push_back((int)v[0]);
        pf.w.push_back(v[1]);
    }
    rd.read(1, v, "nsf");
    int nsf = (int)v[0];
    if (nsf < 0 || nsf > 100000) throw std::runtime_error("gap_parameters: bad nsf");
    for (int i = 0; i < nsf; i++) {
        rd.read(3, v, "symmetry function");
        pf.ntype.push_back((int)v[0]);
        pf.alpha.push_back(v[1]);
        pf.cutoff.push_back(v[2]);
    }
    rd.read(2, v, "nsparseX des_len");
    pf.nsparse = (int)v[0];
    pf.des_len = (int)v[1];
    if (pf.nsparse < 0 || pf.des_len < 0) throw std::runtime_error("gap_parameters: negative sizes");
    for (int i = 0; i < 3; i++) rd.skip("header records");  // gap_calc.f90:351-353
    pf.theta.resize(pf.des_len);
    pf.mm.resize((size_t)pf.nsparse * pf.des_len);
    pf.coeff.resize(pf.nsparse);
    rd.read(pf.des_len, pf.theta.data(), "theta");
    for (int i = 0; i < pf.nsparse; i++) rd.read(pf.des_len, pf.mm.data() + (size_t)i * pf.des_len, "MM");
    rd.read(pf.nsparse, pf.coeff.data(), "coeff");
    pf.has_gpr = true;
    return pf;
}

double sqrt_threshold(double c) {
    // sqrt is correctly rounded and monotone, so {x : sqrt(x) <= c} = [0, T].
    double t = c * c;
    while (std::sqrt(t) > c) t = std::nextafter(t, 0.0);
    for (;;) {
        double u = std::nextafter(t, INFINITY);
        if (std::sqrt(u) <= c) t = u; else break;
    }
    return t;
}

SfPlan make_plan(const std::vector<int> &ntype, const std::vector<double> &alpha,
                 const std::vector<double> &cutoff) {
    SfPlan pl;
    pl.nsf = (int)ntype.size();
    pl.D = 2 * pl.nsf;
    std::vector<double> rcs;
    for (int i = 0; i < pl.nsf; i++)
        if (ntype[i] >= 1 && ntype[i] <= 4) rcs.push_back(cutoff[i]);
    std::sort(rcs.begin(), rcs.end(), [](double a, double b) { return a > b; });
    rcs.erase(std::unique(rcs.begin(), rcs.end()), rcs.end());
    if ((int)rcs.size() > MAXC) throw std::length_error("more than 16 distinct symmetry-function cutoffs");
    pl.ncls = (int)rcs.size();
    for (int c = 0; c < pl.ncls; c++) {
        pl.rc[c] = rcs[c];
        pl.t2[c] = sqrt_threshold(rcs[c]);
        pl.pirc[c] = PI_REF / rcs[c];
    }
    auto cls_of = [&](double rc) { return (int)(std::find(rcs.begin(), rcs.end(), rc) - rcs.begin()); };
    for (int c = 0; c < pl.ncls; c++) {
        pl.grp_begin[c] = (int)pl.grp_alpha.size();
        const size_t first = pl.grp_alpha.size();
        for (int i = 0; i < pl.nsf; i++) {
            if (!(ntype[i] == 2 || ntype[i] == 4) || cutoff[i] != rcs[c]) continue;
            const bool plus = ntype[i] == 2;
            size_t g = first;
            for (; g < pl.grp_alpha.size(); g++)
                if (pl.grp_alpha[g] == alpha[i] && (plus ? pl.grp_iplus[g] : pl.grp_iminus[g]) < 0) break;
            if (g == pl.grp_alpha.size()) {
                pl.grp_alpha.push_back(alpha[i]);
                pl.grp_iplus.push_back(-1);
                pl.grp_iminus.push_back(-1);
            }
            (plus ? pl.grp_iplus[g] : pl.grp_iminus[g]) = i;
            pl.n_asf++;
        }
        if (pl.grp_alpha.size() > first)
            for (int b = c + 1; b <= pl.ncls; b++) pl.ang_prefix_mask |= (1u << b);
    }
    for (int c = pl.ncls; c <= MAXC; c++) pl.grp_begin[c] = (int)pl.grp_alpha.size();
    for (int i = 0; i < pl.nsf; i++) {
        if (ntype[i] == 1 || ntype[i] == 3) {
            pl.rad_ii.push_back(i);
            pl.rad_cls.push_back(cls_of(cutoff[i]));
            pl.rad_type.push_back(ntype[i]);
            pl.rad_p.push_back(alpha[i]);
        } else if (ntype[i] != 2 && ntype[i] != 4) {
            pl.n_unknown++;
        }
    }
    pl.n_rad = (int)pl.rad_ii.size();
    pl.n_grp = (int)pl.grp_alpha.size();
    auto puti = [&](const std::vector<int> &v) { int o = (int)pl.itab.size(); pl.itab.insert(pl.itab.end(), v.begin(), v.end()); return o; };
    auto putd = [&](const std::vector<double> &v) { int o = (int)pl.dtab.size(); pl.dtab.insert(pl.dtab.end(), v.begin(), v.end()); return o; };
    pl.o_rad_ii = puti(pl.rad_ii);
    pl.o_rad_cls = puti(pl.rad_cls);
    pl.o_rad_type = puti(pl.rad_type);
    pl.o_grp_iplus = puti(pl.grp_iplus);
    pl.o_grp_iminus = puti(pl.grp_iminus);
    pl.o_rad_p = putd(pl.rad_p);
    pl.o_grp_alpha = putd(pl.grp_alpha);
    return pl;
}

namespace {
inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
}  // namespace

CellInfo make_cell(const double *L, double rcut, double rbin) {
    if (!(rbin >= rcut)) rbin = rcut;   // the cell list serves the candidate radius rcut + skin; the image window stays the reference's
    CellInfo ci;
    std::memcpy(ci.lat, L, sizeof ci.lat);
    // --- image window exactly as gap_calc.f90:85-88 / :240-257 -------------
    // recipvector() crosses the COLUMNS of lat; nabc reads the ROWS of the result.
    double col[3][3];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) col[c][r] = L[r * 3 + c];
    double x23[3], x31[3], x12[3];
    cross3(col[1], col[2], x23);
    cross3(col[2], col[0], x31);
    cross3(col[0], col[1], x12);
    double vol = std::fabs((col[0][0] * x23[0] + col[0][1] * x23[1]) + col[0][2] * x23[2]);
    for (int r = 0; r < 3; r++) {
        double g[3] = {x23[r] / vol * PI_REF * 2.0, x31[r] / vol * PI_REF * 2.0, x12[r] / vol * PI_REF * 2.0};
        double len = std::sqrt((g[0] * g[0] + g[1] * g[1]) + g[2] * g[2]);
        ci.nabc[r] = (int)std::ceil(rcut * len / PI_REF / 2);
    }
    // --- determinant as Det() (gap_calc.f90:295-297) -------------------------
    double det = L[0] * (L[4] * L[8] - L[5] * L[7]) - L[1] * (L[3] * L[8] - L[5] * L[6]) +
                 L[2] * (L[3] * L[7] - L[6] * L[4]);
    ci.volume = std::fabs(det);
    // --- inverse (frac = pos * inv) and the cell-list grid -------------------
    const double *m = L;
    ci.inv[0] = (m[4] * m[8] - m[5] * m[7]) / det; ci.inv[1] = (m[2] * m[7] - m[1] * m[8]) / det; ci.inv[2] = (m[1] * m[5] - m[2] * m[4]) / det;
    ci.inv[3] = (m[5] * m[6] - m[3] * m[8]) / det; ci.inv[4] = (m[0] * m[8] - m[2] * m[6]) / det; ci.inv[5] = (m[2] * m[3] - m[0] * m[5]) / det;
    ci.inv[6] = (m[3] * m[7] - m[4] * m[6]) / det; ci.inv[7] = (m[1] * m[6] - m[0] * m[7]) / det; ci.inv[8] = (m[0] * m[4] - m[1] * m[3]) / det;
    for (int c = 0; c < 3; c++) {
        double len = std::sqrt(ci.inv[c] * ci.inv[c] + ci.inv[3 + c] * ci.inv[3 + c] + ci.inv[6 + c] * ci.inv[6 + c]);
        double spacing = 1.0 / len;                       // interplanar spacing along direction c
        // Bins at least rcut/2 thick, two neighbour layers: atoms whose bins differ by d layers are at
        // least (d-1) bin widths apart along this direction, so d <= 2 covers rcut; the 5x5x5 block of
        // half-size bins holds ~3.7 rcut-spheres of candidates instead of the ~6.4 of 3x3x3 full-size ones
        // (k_neigh tests every candidate with the reference's arithmetic, so candidates are its cost).
        const double half = 0.5 * rbin * (1.0 + 1e-9), full = rbin * (1.0 + 1e-9);
        int nb = (int)std::floor(spacing / half);
        if (nb < 1) nb = 1;
        if (nb > 1024) nb = 1024;
        ci.nbin[c] = nb;
        // a single thinner bin is scanned over every image the reference would visit (plus one for safety)
        const double w = spacing / nb;
        ci.mscan[c] = w >= full ? 1 : w >= half ? 2 : ci.nabc[c] + 1;
    }
    return ci;
}

}  // namespace gapcu
