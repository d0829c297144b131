// gather.cu -- K5: force assembly and the per-structure reductions.
//
// The reference sums  force(i) = - sum_n sum_k dedg(n,k) * dxdy(k,n,i,:)  over ALL
// centres n through the dense dxdy array (gap_calc.f90:177-185).  Here every
// centre n left dE_n/dx_slot for each of its neighbour slots (centre_impl.cuh); atom i
// walks its own SKIN list (every image within rcut + skin, neigh.cu) and, for each
// candidate (n, shift), looks for the mirror entry (i, -shift) in n's sorted EXACT list
// by binary search; if n lists i, that slot's gradient is added.  Walking the superset
// makes the gather complete even when the two directed distances of a pair straddle rcut
// by an ulp (n lists i although i does not list n): no push path, no atomics, one plain
// store per force component, every sum in a fixed order -> bit-reproducible.
#include <cstdint>

#include "device_types.cuh"
#include "launch.cuh"

namespace gapcu {

constexpr double GPA2EVPANG = 6.24219e-3;  // gap_calc.f90:9

constexpr int GT = 128;  // threads per atom in k_gather: one mirror lookup per thread for P <= 128

// arguments of the per-structure reduction that rides along in k_gather's grid (see below)
struct FinArgs {
    const double *eatom, *vir;
    double *partial, *out8;
    int lgrad, nchunk, nstruct, n_own, raw;
};
__device__ void finalize_partial(const StructDev *structs, const FinArgs &f, int chunk, int st);

// Grid: nchunk * nstruct CTAs that reduce E and the strs contraction per structure (first, so
// that they do not form a tail), then ntot CTAs that gather the forces (one atom each).  Both
// parts read only what the centre kernel wrote, so they share one launch (a separate launch
// costs ~7 us on small inputs).
__global__ void __launch_bounds__(GT)
k_gather(const GatherArgs g, const int nfin, const FinArgs fin) {
    __shared__ double red[GT / 32][3];
    if ((int)blockIdx.x < nfin) {
        finalize_partial(g.structs, fin, blockIdx.x % fin.nchunk, blockIdx.x / fin.nchunk);
        return;
    }
    const int i = g.i_begin + blockIdx.x - nfin;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (g.nloc && i >= *g.nloc) return;
    const int ntot = g.ntot, cap = g.cap;
    const bool i_owned = i < g.n_own;
    const StructDev &sd = g.structs[g.sid[i]];
    const int il = i - sd.atom_off;
    const int P = min(g.skin_cnt[i], cap);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int s = tid; s < P; s += GT) {
        int jl, n1, n2, n3;
        nbr_unkey(g.skin_keys[(size_t)i * cap + s], jl, n1, n2, n3);
        const int nb = sd.atom_off + jl;
        if (nb >= g.n_own) continue;      // a ghost is nobody's centre here: its owner's rank returns that part
        const uint64_t want = nbr_key(il, -n1, -n2, -n3);
        const uint64_t *lst = g.nbr_keys + (size_t)nb * cap;
        const int cnt_nb = min(g.nbr_cnt[nb], cap);
        int lo = 0, hi = cnt_nb;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (lst[mid] < want) lo = mid + 1; else hi = mid;
        }
        if (lo < cnt_nb && lst[lo] == want) {
            const double *fp = g.fpair + ((size_t)nb * cap + lo) * 3;
            gx += fp[0]; gy += fp[1]; gz += fp[2];
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        gx += __shfl_xor_sync(0xffffffffu, gx, o);
        gy += __shfl_xor_sync(0xffffffffu, gy, o);
        gz += __shfl_xor_sync(0xffffffffu, gz, o);
    }
    if (lane == 0) { red[wid][0] = gx; red[wid][1] = gy; red[wid][2] = gz; }
    __syncthreads();
    if (tid < 3) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < GT / 32; w++) v += red[w][tid];
        if (i_owned) g.force_soa[(size_t)tid * ntot + i] = -(g.gself[(size_t)i * 3 + tid] + v);
        else g.ghost_grad[(size_t)g.gslot[i - g.n_own] * 3 + tid] = v;
    }
}

// E = sum e_i (gap_calc.f90:154) and the strs contraction (gap_calc.f90:189-203), in two fixed-order
// levels: CTA (chunk, structure) sums FIN_CHUNK atoms, then one warp per structure adds the
// chunks in order and writes the outputs in the order of gap_calc.f90:221-226.  A single CTA per
// structure (the first version) costs ~0.1 ms on a 100k-atom cell.
constexpr int FIN_CHUNK = 2048;

// s = (E, then (xx, xy, xz, yy, yz, zz) of sum delta_a * dE/dx_b); stress = -that / (6.24219e-3 * V)
__device__ __forceinline__ void write_out8(const double *s, const StructDev &sd, double *o) {
    o[0] = s[0];
    const double f = (1.0 / GPA2EVPANG) / sd.volume;
    o[1] = -s[1] * f;  // xx
    o[2] = -s[4] * f;  // yy
    o[3] = -s[6] * f;  // zz
    o[4] = -s[2] * f;  // xy
    o[5] = -s[5] * f;  // yz
    o[6] = -s[3] * f;  // xz
    o[7] = 0.0;        // variance (gap_calc.f90:206)
}

__device__ void finalize_partial(const StructDev *structs, const FinArgs &f, const int chunk, const int st) {
    __shared__ double red[GT / 32][7];
    __shared__ double tot[8];
    const StructDev &sd = structs[st];
    const double *eatom = f.eatom, *vir = f.vir;
    double *partial = f.partial, *out8 = f.out8;
    const int lgrad = f.lgrad, nchunk = f.nchunk;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t0 = chunk * FIN_CHUNK, t1 = min(min(sd.natoms, f.n_own), t0 + FIN_CHUNK);   // decomposed run: this rank's centres only
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int t = t0 + tid; t < t1; t += GT) {
        const int i = sd.atom_off + t;
        v[0] += eatom[i];
        if (lgrad)
#pragma unroll
            for (int q = 0; q < 6; q++) v[1 + q] += vir[(size_t)i * 6 + q];
    }
#pragma unroll
    for (int q = 0; q < 7; q++) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[wid][q] = x;
    }
    __syncthreads();
    if (tid < 7) {
        double x = 0.0;
        for (int w = 0; w < GT / 32; w++) x += red[w][tid];
        partial[((size_t)st * nchunk + chunk) * 8 + tid] = x;
        tot[tid] = x;
    }
    if (nchunk == 1) {   // small structures: this CTA already holds the totals, no second launch
        __syncthreads();
        if (tid == 0) {
            if (f.raw) { for (int q = 0; q < 7; q++) out8[(size_t)st * 8 + q] = tot[q]; out8[(size_t)st * 8 + 7] = 0.0; }
            else write_out8(tot, sd, out8 + (size_t)st * 8);
        }
    }
}

__global__ void __launch_bounds__(32)
k_finalize(const StructDev *structs, const double *partial, int nchunk, double *out8, int n_own, int raw) {
    __shared__ double s[8];
    const StructDev &sd = structs[blockIdx.x];
    const int lane = threadIdx.x;
    const int used = (min(sd.natoms, n_own) + FIN_CHUNK - 1) / FIN_CHUNK;   // chunks of this structure that hold atoms
    if (lane < 7) {
        const double *p = partial + (size_t)blockIdx.x * nchunk * 8 + lane;
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
        int c = 0;
        for (; c + 3 < used; c += 4) { x0 += p[8 * c]; x1 += p[8 * c + 8]; x2 += p[8 * c + 16]; x3 += p[8 * c + 24]; }
        for (; c < used; c++) x0 += p[8 * c];
        s[lane] = (x0 + x1) + (x2 + x3);
    }
    __syncwarp();
    if (lane == 0) {
        if (raw) { for (int q = 0; q < 7; q++) out8[(size_t)blockIdx.x * 8 + q] = s[q]; out8[(size_t)blockIdx.x * 8 + 7] = 0.0; }
        else write_out8(s, sd, out8 + (size_t)blockIdx.x * 8);
    }
}

// dE/dG := unit vector e_k for every atom (CAR2ACSF export: one backward pass per descriptor)
__global__ void k_onehot(double *dEdG, int ntot, int D, int k) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (size_t)ntot * D) dEdG[t] = ((int)(t % D) == k) ? 1.0 : 0.0;
}
void launch_onehot(cudaStream_t st, double *dEdG, int ntot, int D, int k) {
    const size_t n = (size_t)ntot * D;
    k_onehot<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dEdG, ntot, D, k);
}

int finalize_chunks(int max_natoms) { return max_natoms > 0 ? (max_natoms + FIN_CHUNK - 1) / FIN_CHUNK : 1; }

void launch_gather(cudaStream_t st, const GatherArgs &g, long *launches) {
    const int nchunk = finalize_chunks(g.max_natoms);
    FinArgs fin;
    fin.eatom = g.eatom; fin.vir = g.vir; fin.partial = g.partial; fin.out8 = g.out8; fin.lgrad = g.lgrad; fin.nchunk = nchunk;
    fin.nstruct = g.nstruct; fin.n_own = g.n_own; fin.raw = g.decomposed;
    const int natom = g.i_count > 0 ? g.i_count : g.ntot - g.i_begin;
    const int ngather = g.lgrad ? natom : 0;   // without gradients only the reduction CTAs run
    const int nfin = g.no_fin ? 0 : nchunk * g.nstruct;
    if (!g.lgrad && !g.no_fin) cudaMemsetAsync(g.force_soa, 0, sizeof(double) * 3 * (size_t)g.ntot, st);
    if (nfin + ngather > 0) k_gather<<<nfin + ngather, GT, 0, st>>>(g, nfin, fin);
    if (nfin && nchunk > 1) k_finalize<<<g.nstruct, 32, 0, st>>>(g.structs, g.partial, nchunk, g.out8, g.n_own, g.decomposed);
    if (launches) *launches += (nfin && nchunk > 1) ? 2 : 1;
}

}  // namespace gapcu
