// launch.cuh -- host launchers of the kernels (one per .cu file) and shared constants.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.cuh"

namespace gapcu {

constexpr int MAX_NEIGHBOR_REF_DEV = 1000;  // gap_calc.f90:68

// neigh.cu
struct NeighborBuild {
    const StructDev *structs;
    const int *sid;
    const double *pos;        // [3][ntot] SoA
    int ntot, nbins_total;
    double rcut, rskin;       // rskin = rcut + skin: radius of the candidate (skin) lists
    int cap;                  // row length of skin_keys / nbr_keys
    // cell list scratch
    int *arank_scratch;       // ints for the multi-CTA cell scan: nbins_total / 4096 + 2
    int4 *abin; int *arank; int *bin_count /* [2][nbins_total] */; int *bin_start; int *bin_atoms; int4 *sabin; double *spos;
    // outputs
    uint64_t *skin_keys; int *skin_cnt;   // candidates within rskin, reference order (null: not kept)
    uint64_t *nbr_keys; int *nbr_cnt;     // the reference's lists (dis <= rcut); nbr_keys null: counts / min distance only
    double *min_dis;
    DevFlags *flags;
    int *order;               // direct kernel only: centres by descending count
    bool direct;              // every structure is small: one kernel runs the reference's double loop
    // decomposed runs: local points = owned [0, n_own) then ghosts, *nloc in all, image shifts in sft
    const int4 *sft;          // null: periodic structures
    const int *nloc;
    int n_own;                // = ntot when not decomposed
    // block form of the list kernel (k_neigh_block): blocks of cells in all structures (0: one CTA per centre, k_neigh)
    int nblocks;
    const int *blk_struct;    // structure of every block (null: one structure)
    double t2skin, t2cut, t2close;   // max{x : sqrt_rn(x) <= rskin}, the same for rcut, max{x : sqrt_rn(x) < 0.5}
};
int neighbor_block_max_bins();        // candidate bins a block may have
int neighbor_block_max_candidates();  // candidates a block may have (beyond: flags->blk_overflow, fall back to k_neigh)
void launch_neighbor_build(cudaStream_t st, const NeighborBuild &b, long *launches);
// exact lists from kept skin lists and moved positions; flags->stale when an atom moved > skin/2
void launch_refilter(cudaStream_t st, const NeighborBuild &b, const double *pos_build, double skin, long *launches);
// direct = true needs natoms * (image window) <= neighbor_direct_max_candidates() for every structure and
// no decomposition; that kernel also fills `order` (no launch_order needed)
int neighbor_direct_max_candidates();

// hist: NB_MAXLIST + 2 ints of device scratch (null: always the single-CTA kernel)
void launch_order(cudaStream_t st, const int *nbr_cnt, int n_centres, int *order, DevFlags *flags, int *hist, long *launches);

// centre.cu  (mode: 0 forward, 1 backward, 2 fused forward + GPR + backward)
size_t centre_smem_bytes(const CentreArgs &a, int mode);
int centre_warps();
int centre_pcap_template(int pcap);   // capacity of the kernel instance that serves `pcap` neighbours: 128, 256, 512 or 1024
size_t centre_stash_words(const CentreArgs &a, int chunks, int ctas);
int launch_forward(cudaStream_t st, const CentreArgs &a, long *launches);
int launch_backward(cudaStream_t st, const CentreArgs &a, long *launches);
int launch_fused(cudaStream_t st, const CentreArgs &a, long *launches);

// gpr.cu
struct GprDev {
    int M, Mp, D, Dp;        // Mp: M padded to 16, Dp: D padded to a supported multiple of 8
    const double *Mt;        // [Mp][Dp]  (MM - cmean)/theta, zero padded
    const double *MtT;       // [Dp][Mp]  its transpose (in-CTA GPR of the fused kernel)
    const double *mn;        // [Mp]      |Mt row|^2
    const double *coeff;     // [Mp]      zero padded
    const double *cmean;     // [Dp]
    const double *itheta;    // [Dp]      1/theta, zero padded
};
int gpr_max_slices(int ntot, int Mp);
int launch_gpr(cudaStream_t st, const GprDev &g, const double *G, int ntot, double *eatom, double *dEdG,
               double *epart, double *accpart, int max_slices, const double *t32, long *launches);
void launch_gpr_prepare(cudaStream_t st, int M, int D, const double *mm_c_order, const double *theta,
                        const double *coeff, int Mp, int Dp, double *Mt, double *MtT, double *mn,
                        double *coeff_p, double *cmean, double *itheta);

// gather.cu
struct GatherArgs {
    const StructDev *structs;
    int nstruct;
    const int *sid;
    int ntot, cap;
    const uint64_t *skin_keys; const int *skin_cnt;   // walked by the gathering atom
    const uint64_t *nbr_keys; const int *nbr_cnt;     // searched for the mirror entry
    const double *fpair, *gself, *vir, *eatom;
    int lgrad;
    double *force_soa;        // [3][ntot] forces of the (owned) atoms
    double *out8;             // [nstruct][8]
    double *partial;          // [nstruct][finalize_chunks][8]
    int max_natoms;
    // decomposed runs: atoms >= n_own are ghosts; their sums go to ghost_grad[(i - n_own)][3] (gradient,
    // not force) and travel back to the owners; E and the strs sums cover the owned atoms only and are
    // NOT converted to stress here (raw partial sums in out8, combined over ranks by halo.cu)
    const int *nloc;
    int n_own;
    double *ghost_grad;       // [padded ghost slots][3]
    const int *gslot;         // ghost (local index - n_own) -> padded slot
    int decomposed;
    // part of the atoms served by this launch (decomposed runs split the gather so that the ghost part, which
    // the gradient return waits for, goes first): atoms [i_begin, i_begin + i_count); i_count = 0: all of them.
    // no_fin: leave the E / strs reduction to the other launch
    int i_begin, i_count, no_fin;
};
void launch_gather(cudaStream_t st, const GatherArgs &g, long *launches);
int finalize_chunks(int max_natoms);   // partial needs nstruct * finalize_chunks(max_natoms) * 8 doubles

// halo.cu (decomposed runs)
constexpr int HALO_HDR = 16;      // bytes: record count of the message, padding
constexpr int HALO_REC = 48;      // bytes per ghost record: x y z w (doubles), gid s1 s2 s3 (ints)
constexpr int HALO_RECLEN = 48;   // doubles per rank in the all-gathered record (E, strs sums, flags, counts)
struct HaloGeom {
    double inv[9];
    int grid[3], mine[3];
    double nu[3];       // (rcut + skin + drift) / brick width: shell thickness in brick units
    double drift[3];    // how far (brick units) an owned atom may sit outside its brick
    int wrap[3][3];     // floor((mine + delta) / grid) for delta = -1, 0, +1: cell crossings towards that neighbour
};
// Layout of the exchange buffers.  A direction's message is a header and cap[d] records.  In the SEND buffer
// the directions that go to the same rank are contiguous (one ncclSend per peer), in the RECEIVE buffer
// those that come from the same rank are (one ncclRecv per peer); within a peer the directions ascend, so
// both sides see the same sub-layout.  The gradient buffers use the same two orders in units of records:
// a rank's ghosts (and the gradients it returns) follow its receive order, the gradients it gets back
// follow its send order.
struct HaloBufs {
    unsigned char *send, *recv;
    const double *rgrad;          // [sum cap][3] gradients returned by the ranks that hold my atoms as ghosts (send order)
    int cap[27];
    unsigned soff[27], roff[27];  // byte offset of direction d in the send / receive buffer
    int ks[27], kr[27];           // record offset of direction d in send order / receive order
    signed char rorder[27];       // directions in receive order (-1: unused tail)
};
void launch_halo_select(cudaStream_t st, const HaloGeom &G, const double *pos, int stride, int n_own, int4 *sft, uint32_t *mask,
                        int *tile_cnt, int *tile_base, const HaloBufs &B, DevFlags *flags, long *launches);
void launch_halo_fill(cudaStream_t st, const HaloGeom &G, const double *pos, const double *wgt, const int4 *sft, int stride,
                      int n_own, const uint32_t *mask, const int *tile_base, const HaloBufs &B, const DevFlags *flags, long *launches);
void launch_halo_unpack(cudaStream_t st, const HaloBufs &B, int n_own, int stride, double *pos, double *wgt, int4 *sft, int *gslot,
                        DevFlags *flags, long *launches);
void launch_halo_add(cudaStream_t st, int n_own, int stride, const uint32_t *mask, const int *tile_base, const HaloBufs &B,
                     double *force, long *launches);
void launch_halo_rec(cudaStream_t st, const double *out8_raw, const DevFlags *flags, double *rec, long *launches);
void launch_halo_combine(cudaStream_t st, const double *rec_all, int nranks, double volume, double *out8, DevFlags *flags, long *launches);

void launch_onehot(cudaStream_t st, double *dEdG, int ntot, int D, int k);

// variance.cu
int launch_variance(cudaStream_t st, const StructDev *structs, int nstruct, int ntot, const double *G, int D, int M, int Mp,
                    int Dp, const double *Mt, const double *cmean, const double *itheta, const double *qmm, double *covf,
                    double *variance);

// microbench.cu
void launch_fp64_peaks(cudaStream_t st, double *dfma_tflops, double *dmma_tflops);

}  // namespace gapcu
