// gpr.cu -- K3: fused sparse-GPR prediction on the FP64 tensor cores (DMMA).
//
// Reference: GET_COV (gap_calc.f90:268-288), e = matmul(ckm, coeff) (:152-154) and
// the dE/dG triple loop (:160-166).  With x' = (G - c)/theta, m'_j = (MM_j - c)/theta
// (c = mean of the sparse points, a shift that cancels in every difference but
// keeps the norms small):
//     r2_ij  = |x'_i|^2 + |m'_j|^2 - 2 x'_i . m'_j          <- GEMM 1 (N x M x D)
//     W_ij   = exp(-r2_ij / 2) * coeff_j                     <- fused epilogue
//     E_i    = sum_j W_ij
//     dE/dG_ik = -( x'_ik E_i - sum_j W_ij m'_jk ) / theta_k <- GEMM 2 (N x D x M)
// so neither the covariance matrix ckm(N,M) nor W ever reaches memory: atomic
// energies and dE/dG leave the kernel directly (a "flash" structure: the M loop
// is the streamed dimension, the N x D accumulator stays in registers).
//
// Both contractions use mma.sync.aligned.m8n8k4 f64 (SASS: DMMA.8x8x4), the only
// FP64 tensor shape native to sm_100a (tcgen05 / wgmma have no f64 kind).  One
// warp owns 8 atoms.  GEMM 1's C fragment (row g, columns 2t, 2t+1 of an 8-sparse
// tile) is fed to GEMM 2 as the A operand without any shuffle by letting GEMM 2's
// k index enumerate the sparse points in the order (2t) then (2t+1).
#include <cstdint>

#include "launch.cuh"

namespace gapcu {

constexpr int GPR_WARPS = 4;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// NT = number of 8-wide descriptor tiles (Dp = 8*NT)
template <int NT>
__global__ void __launch_bounds__(32 * GPR_WARPS)
k_gpr(GprDev p, const double *__restrict__ G, int ntot, double *__restrict__ eatom, double *__restrict__ dEdG) {
    extern __shared__ __align__(16) double xs_all[];  // [GPR_WARPS][8][Dp+1]
    constexpr int Dp = 8 * NT;
    constexpr int LDX = Dp + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *xs = xs_all + (size_t)wid * 8 * LDX;
    const int row0 = (blockIdx.x * GPR_WARPS + wid) * 8;
    if (row0 >= ntot) return;
    // scaled, centred descriptors of this warp's 8 atoms
    for (int k = lane; k < Dp; k += 32) {
        const double it = p.itheta[k], cm = p.cmean[k];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int row = row0 + r;
            xs[r * LDX + k] = (row < ntot && k < p.D) ? (G[(size_t)row * p.D + k] - cm) * it : 0.0;
        }
    }
    __syncwarp();
    // |x'|^2 of row g (the 4 lanes of a quad split k, then combine)
    double xn = 0.0;
    for (int k = t; k < Dp; k += 4) { const double v = xs[g * LDX + k]; xn += v * v; }
    xn += __shfl_xor_sync(0xffffffffu, xn, 1);
    xn += __shfl_xor_sync(0xffffffffu, xn, 2);

    double acc[NT][2];
#pragma unroll
    for (int n = 0; n < NT; n++) acc[n][0] = acc[n][1] = 0.0;
    double esum = 0.0;
    const double *__restrict__ Mt = p.Mt;
    for (int sp0 = 0; sp0 < p.Mp; sp0 += 8) {
        // GEMM 1: S(8 atoms x 8 sparse) over k = descriptor index
        double c0 = 0.0, c1 = 0.0;
        const double *mrow = Mt + (size_t)(sp0 + g) * Dp + t;  // B(k=t, n=g) = Mt[sp0+g][4ks+t]
        const double *xrow = xs + g * LDX + t;                 // A(row=g, k=t)
#pragma unroll 4
        for (int ks = 0; ks < 2 * NT; ks++) dmma884(c0, c1, xrow[4 * ks], __ldg(mrow + 4 * ks));
        const int col = sp0 + 2 * t;
        const double w0 = exp(-0.5 * (xn + __ldg(p.mn + col) - 2.0 * c0)) * __ldg(p.coeff + col);
        const double w1 = exp(-0.5 * (xn + __ldg(p.mn + col + 1) - 2.0 * c1)) * __ldg(p.coeff + col + 1);
        esum += w0 + w1;
        // GEMM 2: acc(8 atoms x Dp) += W(8 x 8 sparse) * Mt(8 sparse x Dp);
        // k-step 0 enumerates sparse points 2t, k-step 1 the points 2t+1.
        const double *b0 = Mt + (size_t)(sp0 + 2 * t) * Dp + g;  // B(k=t, n=g) = Mt[sp0+2t][8n+g]
        const double *b1 = b0 + Dp;
#pragma unroll
        for (int n = 0; n < NT; n++) {
            dmma884(acc[n][0], acc[n][1], w0, __ldg(b0 + 8 * n));
            dmma884(acc[n][0], acc[n][1], w1, __ldg(b1 + 8 * n));
        }
    }
    esum += __shfl_xor_sync(0xffffffffu, esum, 1);
    esum += __shfl_xor_sync(0xffffffffu, esum, 2);
    const int row = row0 + g;
    if (row < ntot) {
        if (t == 0) eatom[row] = esum;
#pragma unroll
        for (int n = 0; n < NT; n++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int k = 8 * n + 2 * t + e;
                if (k < p.D) dEdG[(size_t)row * p.D + k] = -p.itheta[k] * (xs[g * LDX + k] * esum - acc[n][e]);
            }
        }
    }
}

// Scaled/centred sparse set (once per potential).  One thread per (row, column).
__global__ void k_gpr_prepare(int M, int D, const double *mm, const double *theta, const double *coeff, int Mp,
                              int Dp, double *Mt, double *MtT, double *mn, double *coeff_p, double *cmean, double *itheta) {
    // phase 1 (block 0 does the column means; tiny problem, run as <<<1, 256>>>)
    for (int k = threadIdx.x; k < Dp; k += blockDim.x) {
        double s = 0.0;
        if (k < D) { for (int j = 0; j < M; j++) s += mm[(size_t)j * D + k]; s /= (M > 0 ? M : 1); }
        cmean[k] = s;
        itheta[k] = (k < D) ? 1.0 / theta[k] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Mp; j += blockDim.x) {
        double nn = 0.0;
        for (int k = 0; k < Dp; k++) {
            double v = (j < M && k < D) ? (mm[(size_t)j * D + k] - cmean[k]) * itheta[k] : 0.0;
            Mt[(size_t)j * Dp + k] = v;
            MtT[(size_t)k * Mp + j] = v;
            nn += v * v;
        }
        mn[j] = nn;
        coeff_p[j] = (j < M) ? coeff[j] : 0.0;
    }
}

void launch_gpr_prepare(cudaStream_t st, int M, int D, const double *mm_c_order, const double *theta,
                        const double *coeff, int Mp, int Dp, double *Mt, double *MtT, double *mn,
                        double *coeff_p, double *cmean, double *itheta) {
    k_gpr_prepare<<<1, 256, 0, st>>>(M, D, mm_c_order, theta, coeff, Mp, Dp, Mt, MtT, mn, coeff_p, cmean, itheta);
}

template <int NT>
static int launch_gpr_nt(cudaStream_t st, const GprDev &g, const double *G, int ntot, double *eatom, double *dEdG) {
    const size_t sm = sizeof(double) * GPR_WARPS * 8 * (8 * NT + 1);
    if (sm > 48 * 1024)
        if (cudaFuncSetAttribute((const void *)k_gpr<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
            return -1;
    const int blocks = (ntot + 8 * GPR_WARPS - 1) / (8 * GPR_WARPS);
    k_gpr<NT><<<blocks, 32 * GPR_WARPS, sm, st>>>(g, G, ntot, eatom, dEdG);
    return 0;
}

int launch_gpr(cudaStream_t st, const GprDev &g, const double *G, int ntot, double *eatom, double *dEdG,
               long *launches) {
    if (launches) *launches += 1;
    switch (g.Dp / 8) {
#define CASE(n) case n: return launch_gpr_nt<n>(st, g, G, ntot, eatom, dEdG);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
        CASE(13) CASE(14) CASE(15) CASE(16) CASE(20) CASE(24) CASE(28) CASE(32)
#undef CASE
        default: return -1;
    }
}

// Dp choices the dispatcher above supports (host picks the smallest >= ceil(D/8)*8)
}  // namespace gapcu
