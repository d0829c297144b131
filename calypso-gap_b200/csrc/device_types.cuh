// device_types.cuh -- plain structs shared by the host orchestration and the kernels.
#pragma once
#include <cstdint>

namespace gapcu {

// One periodic structure of a batch (device copy of CellInfo + offsets).
struct StructDev {
    double lat[9];   // rows = lattice vectors
    double inv[9];   // frac = pos * inv
    double volume;
    double spacing[3];   // interplanar spacing of each lattice direction
    int nabc[3];
    int nbin[3];
    int mscan[3];
    int atom_off, natoms, bin_off, nbins;
    // binned region in fractional coordinates: the whole periodic cell (org = 0, wid = 1, open = 0) or,
    // for one rank of a decomposed run, its brick plus the ghost shell, not periodic (open = 1)
    double org[3], wid[3];
    int open;
    // blocks of cells for k_neigh_block: bs bins per block and direction, nblk blocks per direction,
    // blk_off = blocks of the structures before this one
    int bs[3], nblk[3], blk_off;
};

// Spatial decomposition of ONE large structure over ranks: the cell is cut
// into grid[0] x grid[1] x grid[2] bricks in fractional coordinates; a rank OWNS the atoms of
// its brick (it evaluates them as centres) and receives, from the owners, the GHOST images within
// rcut + skin of the brick (halo.cu).  margin = (rcut + skin) / interplanar spacing: the shell
// thickness in fractional units.
struct DomainDev {
    int enabled;
    int grid[3];
    int mine[3];
    double margin[3];
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
    int n_active;        // atoms with role >= 1 (decomposed runs)
    int n_centres;       // atoms with role == 2
    int n_gt[4];         // centres with more than 128 / 256 / 512 / 1024 neighbours: the capacity tiers' places in `order`
    int queue[16];       // work-queue heads of the persistent centre kernels (mode * 4 + tier)
    unsigned long long work[10];  // see gapcu_ctx_work_counters
    // persistent centre kernel, nanoseconds of %globaltimer: earliest start, earliest / latest
    // exit and the sum of the CTAs' busy times (load-balance diagnostics)
    unsigned long long t_start_min, t_exit_min, t_exit_max, t_busy_sum, n_ctas;
    int ticket;          // CTAs of k_neigh_direct that are done (the last one orders the centres)
    int maxskin;         // largest skin-list length seen (the lists' capacity must hold it)
    int stale;           // an atom moved more than skin/2 since the skin lists were built: rebuild
    // decomposed runs (halo.cu)
    int halo_overflow;   // a send list or the local point array outgrew its capacity
    int halo_count[27];  // records this rank sends in each direction (index 13 = centre, unused)
    int n_ghost, n_loc;  // ghosts received, owned + ghosts
    int halo_far;        // an owned atom drifted further from its brick than the exchange pattern covers
    int blk_overflow;    // a block of cells has more candidates than k_neigh_block holds: use k_neigh
    // GAPCU_VARIANT & 16: cycles thread 0 of every centre CTA spent per phase of the centre kernel
    // (0 stage, 1 radial fwd, 2 list build, 3 angular fwd, 4 reduce + GPR, 5 radial bwd, 6 angular bwd, 7 epilogue)
    unsigned long long phase_cycles[16];   // 8..15: sub-phases (8 pair tests, 9 bucket scan, 10 list scatter, 11 descriptor sums, 12 GPR distances, 13 GPR gradient, 14 backward batches, 15 accumulator merge); their parents hold the rest
};

constexpr int MAXC_DEV = 16;  // distinct cutoffs (classes) supported

// Symmetry-function tables on the device (flat int / double tables + offsets).
// Angular functions are grouped: class (cutoff) -> groups of equal alpha; a group
// holds at most one lambda=+1 (type 2) and one lambda=-1 (type 4) function.
struct PlanDev {
    const int *itab;
    const double *dtab;
    int n_itab, n_dtab;
    int nsf, D, ncls, n_rad, n_grp;
    int o_rad_ii, o_rad_cls, o_rad_type, o_grp_iplus, o_grp_iminus;  // offsets into itab
    int o_rad_p, o_grp_alpha;                                         // offsets into dtab
};

// Cutoff-class tables, passed by value in the kernel parameters (constant bank).
struct ClassTab {
    double rc[MAXC_DEV];    // class cutoff, descending
    double t2[MAXC_DEV];    // largest x with sqrt_rn(x) <= rc; -1 for absent classes
    double pirc[MAXC_DEV];  // PI_REF / rc
    int grp_begin[MAXC_DEV + 1];
    uint32_t angmask;       // bit b: some class c < b holds angular functions
};

// shared-memory layout of the centre kernel (byte offsets), computed on the host
struct SmemLayout {
    int t32, t2, galpha, gd, x, r, ir, w, fc, dfc, gw, sG, gx, sdu, xs, sW, acc, red, S, scratch, ctl, rad, nc, total;
};

// Everything the per-centre kernel needs.
struct CentreArgs {
    PlanDev plan;
    ClassTab cls;
    SmemLayout lay;
    const StructDev *structs;
    const int *sid;             // [NT] structure of each atom
    const double *pos;          // [3][NT] SoA
    const double *wgt;          // [NT] species weight
    const uint64_t *nbr_keys;   // [NT][cap]
    const int *nbr_cnt;         // [NT]
    const double *nbr_table;    // CAR2ACSF entry: the caller's neighbor(NA,ld,6) table (column-major) instead of keys
    int table_ld;
    const int *order;           // [NT] centres by descending neighbour count (null: natural order)
    const int *n_centres;       // device count of entries in `order` (null: ntot)
    // capacity tier served by this launch: entries [*q_begin, *q_end) of `order` (null: 0 / all), own queue head
    const int *q_begin, *q_end;
    int queue_slot;
    const double *exp2_table;   // [32] 2^(j/32), then SINCOS_TAB_N x (cos, sin)(k/16)
    int ntot, cap, pcap;        // ntot: stride of the SoA arrays; pcap: shared-memory neighbour capacity (>= max count)
    int ncentres_max;           // upper bound of the centres this launch serves (sizes the persistent grid)
    int lcap;                   // triplet-list capacity per chunk
    uint32_t *list_scratch;     // [ctas][list_scratch_chunks][lcap+32+512] sorted lists kept from forward to backward
    int list_scratch_chunks;
    int npa;                    // private accumulator sets in backward: NW, or 1 (= shared + atomics)
    int lgrad;
    int variant;                // experiment switches (environment GAPCU_VARIANT), 0 in production
    int cs;                     // CTAs per centre (thread-block cluster size of the fused kernel): 1, 2 or 4
    double *G;                  // [NT][D]   descriptors out (forward / fused; may be null in fused)
    const double *dEdG;         // [NT][D]   backward in (MODE_BWD)
    double *dEdG_out;           // [NT][D]   fused: dE/dG out (may be null)
    double *eatom;              // [NT]      fused: atomic energies
    double *fpair;              // [NT][cap][3] dE_i/dx_(slot)
    double *gself;              // [NT][3]      dE_i/dx_i
    double *vir;                // [NT][6]      sum_slots delta_a * grad_b, (xx,xy,xz,yy,yz,zz)
    // fused GPR (scaled, centred sparse set; see gpr.cu)
    int gpr_M, gpr_Mp, gpr_Dp;
    const double *gpr_Mt;       // [Mp][Dp]
    const double *gpr_MtT;      // [Dp][Mp]
    const double *gpr_coeff;    // [Mp]
    const double *gpr_cmean;    // [Dp]
    const double *gpr_itheta;   // [Dp]
    DevFlags *flags;
    // MODE_FUSED_SE: parked exponentials, [persistent CTAs][estash_stride] double2 (L2 resident)
    double2 *estash;
    int estash_stride;
    int share_exp;              // 1 or 2: every angular class carries the same one / two alphas; 0: not so
    int c_first;                // first class with angular functions
    // debug export of the kept neighbour pairs (triplets i-j-k) per centre: items = slot_j | slot_k << 10 | nclasses << 20
    uint32_t *trip_out;         // [ntot][trip_cap] or null
    int *trip_cnt;              // [ntot]
    int trip_cap;
};

}  // namespace gapcu
