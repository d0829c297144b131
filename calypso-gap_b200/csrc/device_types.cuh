// device_types.cuh -- plain structs shared by the host orchestration and the kernels.
#pragma once
#include <cstdint>

namespace gapcu {

// One periodic structure of a batch (device copy of CellInfo + offsets).
struct StructDev {
    double lat[9];   // rows = lattice vectors
    double inv[9];   // frac = pos * inv
    double volume;
    int nabc[3];
    int nbin[3];
    int mscan[3];
    int atom_off, natoms, bin_off, nbins;
};

// Neighbour key: canonical (reference) order (j, n1, n2, n3) is plain integer order.
//   bits 63..32 atom index j (global within the batch), 29..20 n1+512, 19..10 n2+512, 9..0 n3+512
__host__ __device__ inline uint64_t nbr_key(int j, int n1, int n2, int n3) {
    return ((uint64_t)(uint32_t)j << 32) | (uint64_t)(((n1 + 512) << 20) | ((n2 + 512) << 10) | (n3 + 512));
}
__host__ __device__ inline void nbr_unkey(uint64_t k, int &j, int &n1, int &n2, int &n3) {
    j = (int)(k >> 32);
    uint32_t s = (uint32_t)k;
    n1 = (int)((s >> 20) & 1023) - 512;
    n2 = (int)((s >> 10) & 1023) - 512;
    n3 = (int)(s & 1023) - 512;
}

// Flags / counters written by the kernels, read back with the results.
struct DevFlags {
    int overflow;        // some atom has more neighbours than the list capacity
    int maxcount;        // largest neighbour count seen
    int too_many;        // some atom exceeds the reference's 1000-neighbour limit
    int close_pairs;     // pairs closer than 0.5 A (reference prints a warning)
    unsigned long long work[10];  // see gapcu_ctx_work_counters
};

// Symmetry-function tables on the device (flat int / double tables + offsets).
struct PlanDev {
    const int *itab;
    const double *dtab;
    int n_itab, n_dtab;
    int nsf, D, ncls, n_rad, n_grp, n_asf;
    int o_rad_ii, o_rad_cls, o_rad_type, o_cls_grp, o_grp_sf, o_asf_ii;
    int o_rc, o_t2, o_pirc, o_rad_p, o_grp_alpha, o_asf_lambda;
    uint32_t ang_prefix_mask;
};

// Everything the per-centre kernels need.
struct CentreArgs {
    PlanDev plan;
    const StructDev *structs;
    const int *sid;             // [NT] structure of each atom
    const double *pos;          // [3][NT] SoA
    const double *wgt;          // [NT] species weight
    const uint64_t *nbr_keys;   // [NT][cap]
    const int *nbr_cnt;         // [NT]
    int ntot, cap, pcap;        // pcap: shared-memory capacity (>= max count)
    double *G;                  // [NT][D]   forward out
    const double *dEdG;         // [NT][D]   backward in
    double *fpair;              // [NT][cap][3] dE_i/dx_(slot)  backward out
    double *gself;              // [NT][3]      dE_i/dx_i
    double *vir;                // [NT][6]      sum_slots delta_a * grad_b, (xx,xy,xz,yy,yz,zz)
    DevFlags *flags;
};

}  // namespace gapcu
