// potential.hpp -- host side of a gap_parameters potential: the text reader
// (format: SURVEY.md 8(b); reference readers gap_calc.f90:75-83, :330-362,
// wacsf.f90:46-55) and the "evaluation plan" the kernels consume.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gapcu {

constexpr double PI_REF = 3.141592654;      // truncated pi of the reference (wacsf.f90:31)
constexpr double GPA2EVPANG = 6.24219e-3;   // gap_calc.f90:9
constexpr int MAX_NEIGHBOR_REF = 1000;      // gap_calc.f90:68
constexpr int MAXC = 16;                    // distinct symmetry-function cutoffs supported

struct PotentialFile {
    std::vector<int> z;
    std::vector<double> w;
    std::vector<int> ntype;
    std::vector<double> alpha, cutoff;
    int nsparse = 0, des_len = 0;
    std::vector<double> theta, mm, coeff;  // mm[nsparse][des_len] (C order)
    bool has_gpr = false;
};

// Parses the whole file.  Throws std::runtime_error with the reference's
// message when the file is missing.
PotentialFile read_gap_parameters(const std::string &path);

// Symmetry functions regrouped for the kernels.  Cutoff classes are the distinct
// cutoffs in DESCENDING order, so the classes a distance belongs to are always a
// prefix 0..nc-1 of the list.  Angular functions of a class are grouped by alpha;
// a group holds at most one lambda=+1 (type 2) and one lambda=-1 (type 4) function
// (a repeated (cutoff, alpha, lambda) opens a further group of the same alpha).
struct SfPlan {
    int nsf = 0, D = 0, ncls = 0;
    double rc[MAXC];        // class cutoff
    double t2[MAXC];        // largest x with sqrt_rn(x) <= rc  (exact squared test)
    double pirc[MAXC];      // PI_REF / rc
    int grp_begin[MAXC + 1];
    // radial functions (types 1 and 3)
    std::vector<int> rad_ii, rad_cls, rad_type;
    std::vector<double> rad_p;      // alpha (type 1) or r_shift (type 3)
    // angular groups
    std::vector<double> grp_alpha;  // [ngrp]
    std::vector<int> grp_iplus, grp_iminus;  // descriptor index of the type-2 / type-4 function, -1 if absent
    uint32_t ang_prefix_mask = 0;   // bit b: some class c < b holds angular functions
    int n_unknown = 0;              // functions of unknown type (reference: print and continue)
    // flat tables uploaded to the device: ints then doubles
    std::vector<int> itab;
    std::vector<double> dtab;
    int o_rad_ii, o_rad_cls, o_rad_type, o_grp_iplus, o_grp_iminus;  // offsets into itab
    int o_rad_p, o_grp_alpha;                                        // offsets into dtab
    int n_rad = 0, n_grp = 0, n_asf = 0;
};

SfPlan make_plan(const std::vector<int> &ntype, const std::vector<double> &alpha,
                 const std::vector<double> &cutoff);

// largest double x such that sqrt (round to nearest) of x is <= c
double sqrt_threshold(double c);

// Lattice quantities of one structure.  lat is C order, rows = lattice vectors.
struct CellInfo {
    double lat[9];
    double inv[9];     // frac = pos * inv  (row vector times matrix)
    int nabc[3];       // image window of the reference (gap_calc.f90:85-88)
    int nbin[3];       // cell-list bins per lattice direction
    int mscan[3];      // bins scanned either side of the centre's bin
    double volume;     // abs(det(lat))  (gap_calc.f90:201)
};
CellInfo make_cell(const double *lat_c_order, double rcut, double rbin = 0.0);   // rbin: radius the cell list must cover (>= rcut)

}  // namespace gapcu
