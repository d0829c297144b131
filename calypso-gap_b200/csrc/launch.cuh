// launch.cuh -- host launchers of the kernels (one per .cu file) and shared constants.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.cuh"

namespace gapcu {

constexpr int MAX_NEIGHBOR_REF_DEV = 1000;  // gap_calc.f90:68

// neigh.cu
void launch_neighbor_build(cudaStream_t st, const StructDev *structs, const int *sid, const double *pos,
                           int ntot, int nbins_total, double rcut, int cap, int4 *abin, int *arank,
                           int *bin_count, int *bin_start, int *bin_atoms, int4 *sabin, double *spos, uint64_t *nbr_keys,
                           int *nbr_cnt, double *min_dis, DevFlags *flags, const DomainDev &dom, unsigned char *role,
                           int *active, int *order, bool direct, long *launches);
// direct = true: every structure has natoms * (image window) <= neighbor_direct_max_candidates() and the
// cell is not decomposed; then one kernel builds the lists AND fills `order` (no launch_order needed)
int neighbor_direct_max_candidates();

void launch_order(cudaStream_t st, const int *nbr_cnt, int ntot, int *order, const unsigned char *role,
                  DevFlags *flags, long *launches);

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
void launch_gather(cudaStream_t st, const StructDev *structs, int nstruct, const int *sid, int ntot, int cap,
                   const uint64_t *nbr_keys, const int *nbr_cnt, const double *fpair, const double *gself,
                   const double *vir, const double *eatom, int lgrad, double *force_soa, double *out8,
                   const unsigned char *role, const int *active, const DevFlags *flags, double *partial, int max_natoms,
                   long *launches);
int finalize_chunks(int max_natoms);   // partial needs nstruct * finalize_chunks(max_natoms) * 8 doubles

void launch_onehot(cudaStream_t st, double *dEdG, int ntot, int D, int k);

// variance.cu
int launch_variance(cudaStream_t st, const StructDev *structs, int nstruct, int ntot, const double *G, int D, int M, int Mp,
                    int Dp, const double *Mt, const double *cmean, const double *itheta, const double *qmm, double *covf,
                    double *variance);

// microbench.cu
void launch_fp64_peaks(cudaStream_t st, double *dfma_tflops, double *dmma_tflops);

}  // namespace gapcu
