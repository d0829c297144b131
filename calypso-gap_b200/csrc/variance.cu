// variance.cu -- predictive variance of the sparse GPR.
//
// The reference carries the formula commented out (gap_calc.f90:205-210) and returns
// VARIANCE = 0; its shipped potential file holds no inverse covariance (gap_calc.f90:358-361).
//     covf(i)  = delta - ckm(i,:) . matmul(QMM, ckm(i,:))
//     VARIANCE = sum_i covf(i) / na
// with ckm(i,j) = delta * exp(-1/2 sum_k ((G(i,k) - MM(j,k)) / theta_k)^2)   (GET_COV, :268-288),
// delta = 1 (:8).  This is the additive entry point for callers that do have QMM; the drop-in
// entry points keep returning 0 as the reference does.
//
// One CTA per atom: the covariance row in shared memory (difference form on the centred, scaled
// descriptors the GPR kernels use), then thread j forms (QMM k)_j with coalesced reads of the
// Fortran-ordered QMM and the CTA sums k_j (QMM k)_j in a fixed order.  N*M^2 multiply-adds: a
// diagnostic pass, not on the E/F/stress path (for M ~ 10^4 it wants the DMMA treatment of gpr.cu).
#include <cstdint>

#include "device_types.cuh"
#include "launch.cuh"

namespace gapcu {

constexpr int VT = 256;

__global__ void __launch_bounds__(VT)
k_variance(const double *G, int D, int M, int Mp, int Dp, const double *Mt, const double *cmean, const double *itheta,
           const double *qmm, double *covf) {
    extern __shared__ double sm[];
    double *xs = sm;              // [D]   scaled, centred descriptor of this atom
    double *kv = sm + ((D + 1) & ~1);   // [M] covariance with every sparse point
    __shared__ double red[VT / 32];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = tid; k < D; k += VT) xs[k] = (G[(size_t)i * D + k] - cmean[k]) * itheta[k];
    __syncthreads();
    for (int j = tid; j < M; j += VT) {
        const double *row = Mt + (size_t)j * Dp;
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        for (; k + 1 < D; k += 2) {
            const double d0 = xs[k] - row[k], d1 = xs[k + 1] - row[k + 1];
            s0 = fma(d0, d0, s0); s1 = fma(d1, d1, s1);
        }
        if (k < D) { const double d0 = xs[k] - row[k]; s0 = fma(d0, d0, s0); }
        kv[j] = exp(-0.5 * (s0 + s1));   // delta = 1
    }
    __syncthreads();
    double acc = 0.0;
    for (int j = tid; j < M; j += VT) {
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
        const double *q = qmm + j;   // QMM(j, l) at qmm[j + M*l]
        int l = 0;
        for (; l + 3 < M; l += 4) {
            r0 = fma(q[(size_t)M * l], kv[l], r0); r1 = fma(q[(size_t)M * (l + 1)], kv[l + 1], r1);
            r2 = fma(q[(size_t)M * (l + 2)], kv[l + 2], r2); r3 = fma(q[(size_t)M * (l + 3)], kv[l + 3], r3);
        }
        for (; l < M; l++) r0 = fma(q[(size_t)M * l], kv[l], r0);
        acc = fma(kv[j], (r0 + r1) + (r2 + r3), acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < VT / 32; w++) t += red[w];
        covf[i] = 1.0 - t;
    }
}

// VARIANCE of every structure: mean of covf over its atoms, fixed order (one warp per structure)
__global__ void __launch_bounds__(32) k_variance_mean(const StructDev *structs, const double *covf, double *variance) {
    const StructDev &sd = structs[blockIdx.x];
    double s = 0.0;
    for (int t = threadIdx.x; t < sd.natoms; t += 32) s += covf[sd.atom_off + t];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) variance[blockIdx.x] = sd.natoms > 0 ? s / sd.natoms : 0.0;
}

int launch_variance(cudaStream_t st, const StructDev *structs, int nstruct, int ntot, const double *G, int D, int M, int Mp,
                    int Dp, const double *Mt, const double *cmean, const double *itheta, const double *qmm, double *covf,
                    double *variance) {
    const size_t sm = sizeof(double) * (((D + 1) & ~1) + (size_t)M);
    if (sm > 200 * 1024) return -1;
    if (cudaFuncSetAttribute((const void *)k_variance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -2;
    k_variance<<<ntot, VT, sm, st>>>(G, D, M, Mp, Dp, Mt, cmean, itheta, qmm, covf);
    k_variance_mean<<<nstruct, 32, 0, st>>>(structs, covf, variance);
    return 0;
}

}  // namespace gapcu
