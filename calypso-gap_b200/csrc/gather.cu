// gather.cu -- K5: force assembly and the per-structure reductions.
//
// The reference sums  force(i) = - sum_n sum_k dedg(n,k) * dxdy(k,n,i,:)  over ALL
// centres n through the dense dxdy array (gap_calc.f90:177-185).  Here every
// centre n left dE_n/dx_slot for each of its neighbour slots (desc.cu); atom i
// walks its OWN neighbour list and, for each entry (n, shift), finds the mirror
// entry (i, -shift) in n's sorted list by binary search and reads that gradient:
// a segmented, atomic-free gather in a fixed order.  The only atomics are one
// add per atom into the zeroed output and the "orphan" path: a pair kept in one
// direction only (distance within one ulp of rcut) is pushed by its producer.
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
    int lgrad, nchunk, nstruct;
};
__device__ void finalize_partial(const StructDev *structs, const FinArgs &f, const unsigned char *role, int chunk, int st);

// Grid: nchunk * nstruct CTAs that reduce E and the strs contraction per structure (first, so
// that they do not form a tail), then ntot CTAs that gather the forces (one atom each).  Both
// parts read only what the centre kernel wrote, so they share one launch (a separate launch
// costs ~7 us on small inputs).
__global__ void __launch_bounds__(GT)
k_gather(const StructDev *structs, const int *sid, int ntot, int cap, const uint64_t *nbr_keys,
         const int *nbr_cnt, const double *fpair, const double *gself, double *force,
         const unsigned char *role, const int *active, const DevFlags *flags, const int nfin, const FinArgs fin) {
    __shared__ double red[GT / 32][3];
    if ((int)blockIdx.x < nfin) {
        finalize_partial(structs, fin, role, blockIdx.x % fin.nchunk, blockIdx.x / fin.nchunk);
        return;
    }
    const int slot_i = blockIdx.x - nfin;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (slot_i >= (active ? flags->n_active : ntot)) return;
    const int i = active ? active[slot_i] : slot_i;
    const bool i_owned = !role || role[i] == 2;
    const StructDev &sd = structs[sid[i]];
    const int il = i - sd.atom_off;
    const int P = min(nbr_cnt[i], cap);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int s = tid; s < P; s += GT) {
        int jl, n1, n2, n3;
        nbr_unkey(nbr_keys[(size_t)i * cap + s], jl, n1, n2, n3);
        const int nb = sd.atom_off + jl;
        const bool nb_owned = !role || role[nb] == 2;      // only this rank's centres have gradients here
        if (!nb_owned && !i_owned) continue;
        const uint64_t want = nbr_key(il, -n1, -n2, -n3);
        const uint64_t *lst = nbr_keys + (size_t)nb * cap;
        const int cnt_nb = min(nbr_cnt[nb], cap);
        int lo = 0, hi = cnt_nb;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (lst[mid] < want) lo = mid + 1; else hi = mid;
        }
        if (lo < cnt_nb && lst[lo] == want) {
            if (nb_owned) {
                const double *fp = fpair + ((size_t)nb * cap + lo) * 3;
                gx += fp[0]; gy += fp[1]; gz += fp[2];
            }
        } else if (i_owned) {
            // nb does not list me, so nobody will gather what I exert on nb: push it
            const double *fp = fpair + ((size_t)i * cap + s) * 3;
            atomicAdd(&force[nb], -fp[0]);
            atomicAdd(&force[ntot + nb], -fp[1]);
            atomicAdd(&force[2 * ntot + nb], -fp[2]);
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
        const double self = i_owned ? gself[(size_t)i * 3 + tid] : 0.0;
        atomicAdd(&force[(size_t)tid * ntot + i], -(self + v));
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

__device__ void finalize_partial(const StructDev *structs, const FinArgs &f, const unsigned char *role, const int chunk, const int st) {
    __shared__ double red[GT / 32][7];
    __shared__ double tot[8];
    const StructDev &sd = structs[st];
    const double *eatom = f.eatom, *vir = f.vir;
    double *partial = f.partial, *out8 = f.out8;
    const int lgrad = f.lgrad, nchunk = f.nchunk;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t0 = chunk * FIN_CHUNK, t1 = min(sd.natoms, t0 + FIN_CHUNK);
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int t = t0 + tid; t < t1; t += GT) {
        const int i = sd.atom_off + t;
        if (role && role[i] != 2) continue;   // decomposed run: partial sums over this rank's centres
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
        if (tid == 0) write_out8(tot, sd, out8 + (size_t)st * 8);
    }
}

__global__ void __launch_bounds__(32)
k_finalize(const StructDev *structs, const double *partial, int nchunk, double *out8) {
    __shared__ double s[8];
    const StructDev &sd = structs[blockIdx.x];
    const int lane = threadIdx.x;
    const int used = (sd.natoms + FIN_CHUNK - 1) / FIN_CHUNK;   // chunks of this structure (the others were not written)
    if (lane < 7) {
        const double *p = partial + (size_t)blockIdx.x * nchunk * 8 + lane;
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
        int c = 0;
        for (; c + 3 < used; c += 4) { x0 += p[8 * c]; x1 += p[8 * c + 8]; x2 += p[8 * c + 16]; x3 += p[8 * c + 24]; }
        for (; c < used; c++) x0 += p[8 * c];
        s[lane] = (x0 + x1) + (x2 + x3);
    }
    __syncwarp();
    if (lane == 0) write_out8(s, sd, out8 + (size_t)blockIdx.x * 8);
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

void launch_gather(cudaStream_t st, const StructDev *structs, int nstruct, const int *sid, int ntot, int cap,
                   const uint64_t *nbr_keys, const int *nbr_cnt, const double *fpair, const double *gself,
                   const double *vir, const double *eatom, int lgrad, double *force_soa, double *out8,
                   const unsigned char *role, const int *active, const DevFlags *flags, double *partial, int max_natoms,
                   long *launches) {
    cudaMemsetAsync(force_soa, 0, sizeof(double) * 3 * (size_t)ntot, st);
    const int nchunk = finalize_chunks(max_natoms);
    FinArgs fin;
    fin.eatom = eatom; fin.vir = vir; fin.partial = partial; fin.out8 = out8; fin.lgrad = lgrad; fin.nchunk = nchunk; fin.nstruct = nstruct;
    const int ngather = lgrad ? ntot : 0;   // without gradients only the reduction CTAs run
    k_gather<<<nchunk * nstruct + ngather, GT, 0, st>>>(structs, sid, ntot, cap, nbr_keys, nbr_cnt, fpair, gself, force_soa, role,
                                                         active, flags, nchunk * nstruct, fin);
    if (nchunk > 1) k_finalize<<<nstruct, 32, 0, st>>>(structs, partial, nchunk, out8);
    if (launches) *launches += nchunk > 1 ? 2 : 1;
}

}  // namespace gapcu
