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

#include "fastmath.cuh"
#include "launch.cuh"

namespace gapcu {

constexpr int GPR_WARPS = 8;   // 64 atoms per CTA share the staged tiles of the sparse set
constexpr int GPR_TS = 16;     // sparse points per staged tile

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Row stride (doubles) of the shared-memory matrices: == 4 (mod 16).  A half warp of a
// fragment load touches 4 rows x 4 consecutive doubles; with this stride those 16 doubles
// fall into 16 distinct bank pairs.
__host__ __device__ constexpr int gpr_ld(int Dp) { return Dp + ((20 - Dp % 16) % 16); }

// Column n of GEMM 1's 8-wide output tile is sparse row rowmap(n) = {0,1,2,3,5,4,7,6}[n] of
// the staged tile.  GEMM 2 reads the rows rowmap(2t+e), t = 0..3, in one instruction: with
// this permutation they are distinct mod 4 for e = 0 and e = 1 (as are rowmap(0..3) and
// rowmap(4..7) for GEMM 1), so every fragment load of both GEMMs is bank-conflict free.
__device__ __forceinline__ int gpr_rowmap(int n) { return n < 4 ? n : (n ^ 1); }

// NT = number of 8-wide descriptor tiles (Dp = 8*NT).  grid = (ceil(N/64), M slices).
// Every CTA handles 64 atoms x one slice of the sparse set and writes PARTIAL sums
// (E over the slice, W*m' over the slice); k_gpr_combine adds the slices in fixed order.
// Tiles of 16 sparse points are double buffered with cp.async: tile i+1 streams in while
// tile i is consumed by the DMMAs.
template <int NT>
__global__ void __launch_bounds__(32 * GPR_WARPS)
k_gpr(GprDev p, const double *__restrict__ G, int ntot, int mslice, double *__restrict__ epart,
      double *__restrict__ accpart, const double *__restrict__ t32g) {
    extern __shared__ __align__(16) double sm[];
    constexpr int Dp = 8 * NT;
    constexpr int LD = gpr_ld(Dp);
    double *xs_all = sm;                                  // [GPR_WARPS][8][LD]
    double *tiles = sm + GPR_WARPS * 8 * LD;              // [2][GPR_TS][LD]
    double *s_t32 = tiles + 2 * GPR_TS * LD;              // [32] exp table
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *xs = xs_all + (size_t)wid * 8 * LD;
    const int row0 = (blockIdx.x * GPR_WARPS + wid) * 8;
    const int sp_begin = blockIdx.y * mslice, sp_end = min(p.Mp, sp_begin + mslice);
    const int ntiles = (sp_end - sp_begin + GPR_TS - 1) / GPR_TS;
    auto stage = [&](int tile_idx, int buf) {
        const char *src = (const char *)(p.Mt + (size_t)(sp_begin + tile_idx * GPR_TS) * Dp);
        double *dst = tiles + (size_t)buf * GPR_TS * LD;
        for (int idx = tid; idx < GPR_TS * (Dp / 2); idx += 32 * GPR_WARPS) {
            const int r = idx / (Dp / 2), c2 = idx - r * (Dp / 2);
            cp_async16(dst + r * LD + 2 * c2, src + (size_t)idx * 16);
        }
        cp_async_commit();
    };
    if (ntiles > 0) stage(0, 0);
    if (tid < 32) s_t32[tid] = t32g[tid];
    // scaled, centred descriptors of this warp's 8 atoms
    for (int k = lane; k < Dp; k += 32) {
        const double it = p.itheta[k], cm = p.cmean[k];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int row = row0 + r;
            xs[r * LD + k] = (row < ntot && k < p.D) ? (G[(size_t)row * p.D + k] - cm) * it : 0.0;
        }
    }
    __syncwarp();
    double xn = 0.0;
    for (int k = t; k < Dp; k += 4) { const double v = xs[g * LD + k]; xn += v * v; }
    xn += __shfl_xor_sync(0xffffffffu, xn, 1);
    xn += __shfl_xor_sync(0xffffffffu, xn, 2);

    double acc[NT][2];
#pragma unroll
    for (int n = 0; n < NT; n++) acc[n][0] = acc[n][1] = 0.0;
    double esum = 0.0;
    const int rg = gpr_rowmap(g);                             // tile row feeding GEMM 1's column n = g
    const int r0 = gpr_rowmap(2 * t), r1 = gpr_rowmap(2 * t + 1);   // rows behind this lane's two C columns
    for (int i = 0; i < ntiles; i++) {
        if (i + 1 < ntiles) { stage(i + 1, (i + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();                                       // tile i is complete for every thread
        const double *tile = tiles + (size_t)(i & 1) * GPR_TS * LD;
        const int sp0 = sp_begin + i * GPR_TS;
        // GEMM 1: S(8 atoms x 16 sparse) over the descriptor index; one A fragment feeds both 8-wide halves
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
        const double *xrow = xs + g * LD + t;                  // A(row=g, k=t)
        const double *m0 = tile + rg * LD + t;                 // B(k=t, n=g) = tile[rowmap(g)][4ks+t]
        const double *m1 = m0 + 8 * LD;
#pragma unroll 4
        for (int ks = 0; ks < 2 * NT; ks++) {
            const double av = xrow[4 * ks];
            dmma884(c00, c01, av, m0[4 * ks]);
            dmma884(c10, c11, av, m1[4 * ks]);
        }
        const int ca = sp0 + r0, cb = sp0 + r1;                // sparse points of this lane's C columns
        const double w00 = exp_neg(-0.5 * (xn + __ldg(p.mn + ca) - 2.0 * c00), s_t32) * __ldg(p.coeff + ca);
        const double w01 = exp_neg(-0.5 * (xn + __ldg(p.mn + cb) - 2.0 * c01), s_t32) * __ldg(p.coeff + cb);
        const double w10 = exp_neg(-0.5 * (xn + __ldg(p.mn + ca + 8) - 2.0 * c10), s_t32) * __ldg(p.coeff + ca + 8);
        const double w11 = exp_neg(-0.5 * (xn + __ldg(p.mn + cb + 8) - 2.0 * c11), s_t32) * __ldg(p.coeff + cb + 8);
        esum += (w00 + w01) + (w10 + w11);
        // GEMM 2: acc(8 atoms x Dp) += W(8 x 16 sparse) * tile(16 sparse x Dp).  The C fragment of
        // GEMM 1 (row g, columns 2t, 2t+1) is used as the A operand as is: k-step e enumerates the
        // sparse rows rowmap(2t+e), so B(k=t, n=g) = tile[rowmap(2t+e)][8n+g].
        const double *b0 = tile + r0 * LD + g, *b1 = tile + r1 * LD + g;
        const double *b2 = b0 + 8 * LD, *b3 = b1 + 8 * LD;
        // k-step outer, descriptor tile inner: consecutive DMMAs write different accumulators (the four k-steps
        // of one accumulator in a row were a dependent chain on the tensor pipe's latency)
#pragma unroll
        for (int n = 0; n < NT; n++) dmma884(acc[n][0], acc[n][1], w00, b0[8 * n]);
#pragma unroll
        for (int n = 0; n < NT; n++) dmma884(acc[n][0], acc[n][1], w01, b1[8 * n]);
#pragma unroll
        for (int n = 0; n < NT; n++) dmma884(acc[n][0], acc[n][1], w10, b2[8 * n]);
#pragma unroll
        for (int n = 0; n < NT; n++) dmma884(acc[n][0], acc[n][1], w11, b3[8 * n]);
        __syncthreads();                                       // every warp is done with this buffer
    }
    esum += __shfl_xor_sync(0xffffffffu, esum, 1);
    esum += __shfl_xor_sync(0xffffffffu, esum, 2);
    const int row = row0 + g;
    if (row < ntot) {
        const size_t o = (size_t)blockIdx.y * ntot + row;
        if (t == 0) epart[o] = esum;
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const int k = 8 * n + 2 * t;
            *(double2 *)(accpart + o * Dp + k) = make_double2(acc[n][0], acc[n][1]);
        }
    }
}

// e_i = sum over slices; dE/dG_ik = -(x'_ik e_i - sum_slices acc_ik) / theta_k     (fixed order)
__global__ void k_gpr_combine(GprDev p, const double *__restrict__ G, int ntot, int nslice,
                              const double *__restrict__ epart, const double *__restrict__ accpart,
                              double *__restrict__ eatom, double *__restrict__ dEdG) {
    const int row = blockIdx.x;
    double e = 0.0;
    for (int s = 0; s < nslice; s++) e += epart[(size_t)s * ntot + row];
    if (threadIdx.x == 0) eatom[row] = e;
    for (int k = threadIdx.x; k < p.D; k += blockDim.x) {
        double a = 0.0;
        for (int s = 0; s < nslice; s++) a += accpart[((size_t)s * ntot + row) * p.Dp + k];
        const double it = p.itheta[k];
        const double xk = (G[(size_t)row * p.D + k] - p.cmean[k]) * it;
        dEdG[(size_t)row * p.D + k] = -it * (xk * e - a);
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

static int gpr_sms() {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    return sms;
}

// upper bound of the number of sparse-set slices launch_gpr may use (sizes the partial buffers)
int gpr_max_slices(int ntot, int Mp) {
    const int row_ctas = (ntot + 8 * GPR_WARPS - 1) / (8 * GPR_WARPS);
    const int tiles = (Mp + GPR_TS - 1) / GPR_TS;
    int want = (8 * gpr_sms() + row_ctas - 1) / row_ctas;
    return want < 1 ? 1 : (want > tiles ? tiles : want);
}

template <int NT>
static int launch_gpr_nt(cudaStream_t st, const GprDev &g, const double *G, int ntot, int max_slices, double *epart,
                         double *accpart, const double *t32, int *nslice_out) {
    constexpr int Dp = 8 * NT;
    const size_t sm = sizeof(double) * (GPR_WARPS * 8 * gpr_ld(Dp) + 2 * GPR_TS * gpr_ld(Dp) + 32);
    if (cudaFuncSetAttribute((const void *)k_gpr<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
        return -1;
    // slice the sparse set so that the grid is a whole number of waves of resident CTAs
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gpr<NT>, 32 * GPR_WARPS, sm) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int slots = gpr_sms() * per_sm;
    const int row_ctas = (ntot + 8 * GPR_WARPS - 1) / (8 * GPR_WARPS);
    const int tiles = (g.Mp + GPR_TS - 1) / GPR_TS;
    int want = slots / row_ctas;                     // one full wave
    if (want < 1) want = 1;
    if (want > tiles) want = tiles;
    if (want > max_slices) want = max_slices;
    const int tps = (tiles + want - 1) / want;       // tiles per slice
    const int nslice = (tiles + tps - 1) / tps;
    dim3 grid(row_ctas, nslice);
    k_gpr<NT><<<grid, 32 * GPR_WARPS, sm, st>>>(g, G, ntot, tps * GPR_TS, epart, accpart, t32);
    *nslice_out = nslice;
    return 0;
}

int launch_gpr(cudaStream_t st, const GprDev &g, const double *G, int ntot, double *eatom, double *dEdG,
               double *epart, double *accpart, int max_slices, const double *t32, long *launches) {
    if (launches) *launches += 2;
    int rc = -1, nslice = 1;
    switch (g.Dp / 8) {
#define CASE(n) case n: rc = launch_gpr_nt<n>(st, g, G, ntot, max_slices, epart, accpart, t32, &nslice); break;
        CASE(2) CASE(4) CASE(6) CASE(8) CASE(9) CASE(10) CASE(12) CASE(14) CASE(16) CASE(20) CASE(24) CASE(28) CASE(32)
#undef CASE
        default: return -1;
    }
    if (rc) return rc;
    k_gpr_combine<<<ntot, 128, 0, st>>>(g, G, ntot, nslice, epart, accpart, eatom, dEdG);
    return 0;
}

}  // namespace gapcu
