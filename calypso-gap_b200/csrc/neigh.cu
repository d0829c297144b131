// neigh.cu -- K1: cell-list neighbour build on the GPU.
//
// Replaces the O(N^2 * images) double loop of gap_calc.f90:89-120 (and of
// get_bond.f90:46-78).  The SET of pairs is the reference's, bit for bit:
//   * the image position and the distance are formed with the reference's
//     operation order and without FMA contraction (gap_calc.f90:98-100),
//   * the test is dis > rcut -> skip (gap_calc.f90:101), only the zero shift of
//     the atom itself is excluded (:97),
//   * only shifts inside the reference's +-nabc window are kept (:94-96), so
//     atoms the caller did not wrap into the cell lose the same neighbours.
// The cell list only proposes candidates.  Each list is sorted into the
// reference order (j, n1, n2, n3), which also makes the result deterministic.
//
// Two lists per atom.  The SKIN list holds every image within rcut + skin, sorted; the
// EXACT list is its subset with dis <= rcut (the reference's set), produced by re-testing the
// skin entries in order (refilter).  A molecular-dynamics caller keeps the skin list over
// several steps (gapcu_ctx_update_positions): each step only re-filters it with the reference
// arithmetic on the new positions, so the sets stay bit-exact, and the list is rebuilt when an
// atom has moved more than skin/2 since it was built.  The skin is never zero (the host adds
// a floor of ~1e-9 A): a pair whose two directed distances straddle rcut by an ulp is then
// still a candidate of BOTH atoms, which is what lets the force gather (gather.cu) find every
// contribution by walking the skin list, without a push path and without atomics.
#include <cstdint>

#include "device_types.cuh"
#include "geom.cuh"
#include "launch.cuh"

namespace gapcu {

constexpr int NB_THREADS = 128;   // threads per centre in k_neigh
constexpr int NB_MAXLIST = 1024;  // shared-memory list length (reference stops above 1000)
constexpr int NB_CELLS = 128;     // cells whose candidates are enumerated together

__device__ __forceinline__ int floordiv_i(int a, int b) {
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return q;
}

// fractional coordinates -> wrap offsets and bin; counts atoms per bin.
// Periodic structures: wrap offset = floor(f).  Open regions (decomposed runs): the point's image
// shift is given (sft.yzw), its "wrap offset" is minus that shift, and the bin is taken relative to
// the region's origin; owned points [0, n_own) and ghosts are counted separately so that the owned
// records come first inside every cell (bin_count: [2][nbins_total]).
__global__ void k_bin(const StructDev *structs, const int *sid, const double *pos, int ntot, const int *nloc, int n_own,
                      const int4 *sft, int4 *abin, int *arank, int *bin_count, int nbins_total) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (nloc ? *nloc : ntot)) return;
    const StructDev &s = structs[sid[i]];
    double x = pos[i], y = pos[ntot + i], z = pos[2 * ntot + i];
    int b[3], w[3];
    int sh[3] = {0, 0, 0};
    if (s.open) { const int4 q = sft[i]; sh[0] = q.y; sh[1] = q.z; sh[2] = q.w; }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double f = x * s.inv[c] + y * s.inv[3 + c] + z * s.inv[6 + c];
        double fw;
        if (s.open) { w[c] = -sh[c]; fw = (f + (double)sh[c] - s.org[c]) / s.wid[c]; }
        else { const double fl = floor(f); w[c] = (int)fl; fw = f - fl; }
        int bc = (int)floor(fw * s.nbin[c]);
        bc = min(max(bc, 0), s.nbin[c] - 1);
        b[c] = bc;
    }
    int id = (b[0] * s.nbin[1] + b[1]) * s.nbin[2] + b[2];
    abin[i] = make_int4(id, w[0], w[1], w[2]);
    arank[i] = atomicAdd(&bin_count[(i >= n_own ? nbins_total : 0) + s.bin_off + id], 1);
}

// exclusive scan of the per-cell totals (owned + ghost) -> bin_start[nb+1]; single CTA.
__global__ void k_scan_bins(const int *bin_count, int *bin_start, int nb) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < nb; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = (i < nb) ? bin_count[i] + bin_count[nb + i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = (lane < nw) ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        int prefix = carry + (wid ? warp_sums[wid - 1] : 0) + x - v;
        if (i < nb) bin_start[i] = prefix;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_start[nb] = carry;
}

// The same scan for many cells, in three launches.  sums: one int per CTA (+1).
constexpr int SCAN_PER_CTA = 4096;
__global__ void __launch_bounds__(1024) k_scan_bins_part(const int *bin_count, int *bin_start, int nb, int *sums) {
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int i0 = blockIdx.x * SCAN_PER_CTA + tid * 4;
    int v[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int i = i0 + k; v[k] = i < nb ? bin_count[i] + bin_count[nb + i] : 0; tot += v[k]; }
    int x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int s = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        wsum[lane] = s;
    }
    __syncthreads();
    int run = (wid ? wsum[wid - 1] : 0) + x - tot;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (i0 + k < nb) bin_start[i0 + k] = run; run += v[k]; }
    if (tid == 1023) sums[blockIdx.x] = run;
}
__global__ void __launch_bounds__(1024) k_scan_bins_top(int *sums, int n, int *bin_start, int nb) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? sums[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            wsum[lane] = s;
        }
        __syncthreads();
        const int excl = carry + (wid ? wsum[wid - 1] : 0) + x - v;
        if (i < n) sums[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) bin_start[nb] = carry;
}
__global__ void __launch_bounds__(1024) k_scan_bins_add(int *bin_start, int nb, const int *sums) {
    const int off = sums[blockIdx.x];
    const int i0 = blockIdx.x * SCAN_PER_CTA + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) if (i0 + k < nb) bin_start[i0 + k] += off;
}

// Besides the permutation bin_atoms, the atoms' records are copied into bin order (sabin: atom
// index + wrap offsets, spos: coordinates), so that k_neigh reads the candidates of a cell as
// contiguous, independent loads instead of a chain bin_atoms -> abin -> pos.
__global__ void k_fill_bins(const StructDev *structs, const int *sid, const double *pos, const int4 *abin, const int *arank,
                            const int *bin_start, const int *bin_count, int nbins_total, int ntot, const int *nloc, int n_own,
                            int *bin_atoms, int4 *sabin, double *spos) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (nloc ? *nloc : ntot)) return;
    const StructDev &s = structs[sid[i]];
    const int4 b = abin[i];
    const int slot = bin_start[s.bin_off + b.x] + (i >= n_own ? bin_count[s.bin_off + b.x] : 0) + arank[i];
    bin_atoms[slot] = i;
    sabin[slot] = make_int4(i, b.y, b.z, b.w);
    spos[slot] = pos[i]; spos[ntot + slot] = pos[ntot + i]; spos[2 * ntot + slot] = pos[2 * ntot + i];
}

// Small batches: bin, scan and fill in ONE single-CTA kernel (shared-memory counters) instead
// of a memset and three launches; same outputs as k_bin / k_scan_bins / k_fill_bins (periodic
// structures only).
constexpr int SMALL_NBINS = 4096;
__global__ void __launch_bounds__(1024)
k_bin_small(const StructDev *structs, const int *sid, const double *pos, int ntot, int nbins_total, int4 *abin,
            int *bin_start, int *bin_atoms, int4 *sabin, double *spos) {
    __shared__ int cnt[SMALL_NBINS + 1];
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int t = tid; t <= nbins_total; t += 1024) cnt[t] = 0;
    __syncthreads();
    for (int i = tid; i < ntot; i += 1024) {
        const StructDev &s = structs[sid[i]];
        const double x = pos[i], y = pos[ntot + i], z = pos[2 * ntot + i];
        int b[3], w[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double f = x * s.inv[c] + y * s.inv[3 + c] + z * s.inv[6 + c];
            const double fl = floor(f);
            w[c] = (int)fl;
            const double fw = f - fl;
            b[c] = min(max((int)(fw * s.nbin[c]), 0), s.nbin[c] - 1);
        }
        const int id = (b[0] * s.nbin[1] + b[1]) * s.nbin[2] + b[2];
        abin[i] = make_int4(id, w[0], w[1], w[2]);
        atomicAdd(&cnt[s.bin_off + id], 1);
    }
    __syncthreads();
    // exclusive scan of cnt[0..nbins_total) in chunks of 1024
    int carry = 0;
    for (int base = 0; base < nbins_total; base += 1024) {
        const int idx = base + tid;
        const int v = idx < nbins_total ? cnt[idx] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int sv = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, sv, o); if (lane >= o) sv += y; }
            wsum[lane] = sv;
        }
        __syncthreads();
        const int excl = carry + (wid ? wsum[wid - 1] : 0) + x - v;
        if (idx < nbins_total) { cnt[idx] = excl; bin_start[idx] = excl; }
        carry += wsum[31];
        __syncthreads();
    }
    if (tid == 0) bin_start[nbins_total] = carry;
    __syncthreads();
    for (int i = tid; i < ntot; i += 1024) {
        const StructDev &s = structs[sid[i]];
        const int4 b = abin[i];
        const int slot = atomicAdd(&cnt[s.bin_off + b.x], 1);
        bin_atoms[slot] = i;
        sabin[slot] = make_int4(i, b.y, b.z, b.w);
        spos[slot] = pos[i]; spos[ntot + slot] = pos[ntot + i]; spos[2 * ntot + slot] = pos[2 * ntot + i];
    }
}


// Exact list of centre i from its sorted candidate list `cand` (shared or global memory): the
// entries with dis <= rcut (gap_calc.f90:101, reference arithmetic), in order.  All NT threads of
// the CTA call it; wsum: NT/32 + 1 ints of shared memory.  Returns the count (every thread).
template <int NT>
__device__ int refilter_list(const uint64_t *cand, int ncand, int i, const double *pos, int ntot, const double *lat,
                             int aoff, double xi, double yi, double zi, double rcut, uint64_t *out, int *wsum,
                             int &nclose_out) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int outn = 0, nclose = 0;
    for (int base = 0; base < ncand; base += NT) {
        const int t = base + tid;
        bool keep = false;
        uint64_t key = 0;
        if (t < ncand) {
            key = cand[t];
            int jl, n1, n2, n3;
            nbr_unkey(key, jl, n1, n2, n3);
            double ox, oy, oz;
            const double dis = image_distance(pos, ntot, aoff + jl, lat, n1, n2, n3, xi, yi, zi, ox, oy, oz);
            keep = !(dis > rcut);
            nclose += keep && dis < 0.5;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncthreads();                 // wsum of the previous round is consumed
        if (lane == 0) wsum[wid] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) { const int c = wsum[w]; if (w < wid) before += c; total += c; }
        if (keep) out[outn + before + __popc(m & ((1u << lane) - 1u))] = key;
        outn += total;
    }
    nclose_out = nclose;
    return outn;
}

// One CTA per centre atom: gather candidates from the surrounding bins, apply the
// reference test against rcut + skin, sort into reference order, store the skin list, then
// re-filter it against rcut into the exact list (+ optional min distance).
// Decomposed runs: the atoms are this rank's LOCAL points -- owned atoms
// [0, n_own) followed by the ghost images received from the neighbour ranks, *nloc in all -- in
// an open (non-periodic) region of fractional space; the "wrap offsets" of a point are minus
// its image shift, so the shift arithmetic below is the same.  A ghost centre only lists OWNED
// candidates (the first bin_nown[cell] records of a cell): its list exists to collect the forces
// this rank's centres exert on it (gather.cu), not to evaluate it.
struct NeighArgs {
    const StructDev *structs;
    const int *sid;
    const double *pos;
    const int4 *abin;
    const int *bin_start;
    const int *bin_nown;    // decomposed: owned records per cell (they come first); else null
    const int4 *sabin;
    const double *spos;
    int ntot;               // stride of the SoA arrays
    const int *nloc;        // decomposed: device count of local points; else null (= ntot)
    int n_own;              // decomposed: centres >= n_own are ghosts; else ntot
    double rcut, rskin;
    int cap;
    uint64_t *skin_keys;    // [ntot][cap] or null (bond-length query)
    int *skin_cnt;
    uint64_t *nbr_keys;     // [ntot][cap]
    int *nbr_cnt;
    double *min_dis;
    DevFlags *flags;
};

__global__ void __launch_bounds__(NB_THREADS, 8) k_neigh(const NeighArgs A) {
    __shared__ uint64_t keys[NB_MAXLIST];
    __shared__ int nkeys;
    __shared__ double lat[9];
    __shared__ double wmin[NB_THREADS / 32];
    __shared__ int wsum[NB_THREADS / 32 + 1];
    __shared__ int c_start[NB_CELLS], c_off[NB_CELLS + 1];
    __shared__ int4 c_shift[NB_CELLS];
    const int i = blockIdx.x;
    if (A.nloc && i >= *A.nloc) return;
    const bool ghost = i >= A.n_own;
    const int ntot = A.ntot;
    const double *pos = A.pos;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const StructDev &s = A.structs[A.sid[i]];
    if (tid < 9) lat[tid] = s.lat[tid];
    if (tid == 0) nkeys = 0;
    __syncthreads();
    const double xi = pos[i], yi = pos[ntot + i], zi = pos[2 * ntot + i];
    const int4 bi = A.abin[i];
    const int nb0 = s.nbin[0], nb1 = s.nbin[1], nb2 = s.nbin[2];
    const int b0 = bi.x / (nb1 * nb2), b1 = (bi.x / nb2) % nb1, b2 = bi.x % nb2;
    const int m0 = s.mscan[0], m1 = s.mscan[1], m2 = s.mscan[2];
    const int na0 = s.nabc[0], na1 = s.nabc[1], na2 = s.nabc[2];
    const int w1 = 2 * m1 + 1, w2 = 2 * m2 + 1;
    const int nscan = (2 * m0 + 1) * w1 * w2;
    const int aoff = s.atom_off;
    const bool open = s.open != 0;
    const double rskin = A.rskin;
    // position of the centre in bin units along each lattice direction: a whole cell is skipped when
    // the slab gap to it along any direction already exceeds the search radius (prunes corner bins
    // and far images).  Periodic cells: org = 0, wid = 1 and the wrap offset is floor(f).
    const int wo[3] = {bi.y, bi.z, bi.w};
    const int nbv[3] = {nb0, nb1, nb2};
    double fb[3], bw[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double f = xi * s.inv[c] + yi * s.inv[3 + c] + zi * s.inv[6 + c];
        fb[c] = (f - (double)wo[c] - s.org[c]) * ((double)nbv[c] / s.wid[c]);
        bw[c] = s.spacing[c] * s.wid[c] / nbv[c];   // bin width in length units
    }
    const double fb0 = fb[0], fb1 = fb[1], fb2 = fb[2], bw0 = bw[0], bw1 = bw[1], bw2 = bw[2];
    const double rprune = rskin * (1.0 + 1e-9) + 1e-9;
    const bool prune = nscan > 27;
    double dmin = 1e300;
    int nclose = 0, nexact = 0;
    // Cells are handled NB_CELLS at a time: their (start, count, shift) are fetched by one
    // thread each, the counts are scanned, and then the candidates of all those cells form ONE
    // flat index space dealt over the threads.
    for (int c0 = 0; c0 < nscan; c0 += NB_CELLS) {
        const int nc = min(NB_CELLS, nscan - c0);
        __syncthreads();
        for (int t = tid; t < NB_CELLS; t += NB_THREADS) {
            int cnt = 0;
            if (t < nc) {
                const int cell = c0 + t;
                const int d0 = cell / (w1 * w2) - m0, d1 = (cell / w2) % w1 - m1, d2 = cell % w2 - m2;
                const int t0 = b0 + d0, t1 = b1 + d1, t2 = b2 + d2;
                int s0 = 0, s1 = 0, s2 = 0;
                bool inside = true;
                if (open) inside = t0 >= 0 && t0 < nb0 && t1 >= 0 && t1 < nb1 && t2 >= 0 && t2 < nb2;
                else { s0 = floordiv_i(t0, nb0); s1 = floordiv_i(t1, nb1); s2 = floordiv_i(t2, nb2); }
                if (inside) {
                    const int id = ((t0 - s0 * nb0) * nb1 + (t1 - s1 * nb1)) * nb2 + (t2 - s2 * nb2);
                    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
                    if (prune) {   // slab gaps (negative: inside); only worth it when many image cells are scanned
                        g0 = fmax((double)t0 - fb0, fb0 - (double)(t0 + 1)) * bw0;
                        g1 = fmax((double)t1 - fb1, fb1 - (double)(t1 + 1)) * bw1;
                        g2 = fmax((double)t2 - fb2, fb2 - (double)(t2 + 1)) * bw2;
                    }
                    const int start = A.bin_start[s.bin_off + id];
                    if (!(g0 > rprune || g1 > rprune || g2 > rprune))
                        cnt = ghost ? A.bin_nown[s.bin_off + id] : A.bin_start[s.bin_off + id + 1] - start;
                    c_start[t] = start;
                    c_shift[t] = make_int4(s0, s1, s2, 0);
                }
            }
            c_off[t] = cnt;
        }
        __syncthreads();
        if (wid == 0) {  // exclusive scan of the NB_CELLS counts by one warp (NB_CELLS / 32 per lane)
            int run = 0;
            for (int base = 0; base < NB_CELLS; base += 32) {
                const int v = c_off[base + lane];
                int x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                c_off[base + lane] = run + x - v;
                run += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) c_off[NB_CELLS] = run;
        }
        __syncthreads();
        const int total = c_off[NB_CELLS];
        // Every thread takes a contiguous run of the flat candidate index: one cell lookup per run
        // instead of one per candidate, then it walks (cells are crossed rarely); the record of
        // the next candidate is requested before the current one is tested.
        struct Cand { int4 bj; double x, y, z; };
        auto load = [&](int slot) {
            Cand r;
            r.bj = A.sabin[slot];                          // (atom, wrap offsets) and coordinates in bin order:
            r.x = A.spos[slot]; r.y = A.spos[ntot + slot]; r.z = A.spos[2 * ntot + slot];   // independent loads
            return r;
        };
        const int per = (total + NB_THREADS - 1) / NB_THREADS;
        int cand = tid * per;
        const int cend = min(total, cand + per);
        if (cand < cend) {
            int lo = 0, hi = nc;  // cell of the first candidate: last cell with c_off <= cand
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (c_off[mid] <= cand) lo = mid; else hi = mid; }
            int slot = c_start[lo] + (cand - c_off[lo]);
            int cell_end = c_off[lo + 1];
            int4 sh = c_shift[lo];
            Cand r = load(slot);
            for (; cand < cend; cand++) {
                const Cand cur = r;
                const int4 csh = sh;
                if (cand + 1 < cend) {
                    slot++;
                    if (cand + 1 >= cell_end) {   // next candidate lives in a later (non-empty) cell
                        do lo++; while (c_off[lo + 1] <= cand + 1);
                        slot = c_start[lo] + (cand + 1 - c_off[lo]);
                        cell_end = c_off[lo + 1];
                        sh = c_shift[lo];
                    }
                    r = load(slot);
                }
                const int j = cur.bj.x;
                // shift between wrapped coordinates -> shift of the caller's coordinates
                const int n1 = csh.x - cur.bj.y + bi.y, n2 = csh.y - cur.bj.z + bi.z, n3 = csh.z - cur.bj.w + bi.w;
                if (j == i && n1 == 0 && n2 == 0 && n3 == 0) continue;
                if (abs(n1) > na0 || abs(n2) > na1 || abs(n3) > na2) continue;
                double ox, oy, oz;
                const double dis = image_distance_xyz(cur.x, cur.y, cur.z, lat, n1, n2, n3, xi, yi, zi, ox, oy, oz);
                if (dis > rskin) continue;
                if (!(dis > A.rcut)) { dmin = fmin(dmin, dis); nclose += dis < 0.5; nexact++; }
                const int p = atomicAdd(&nkeys, 1);   // order of arrival is irrelevant: the list is sorted below
                if (p < NB_MAXLIST) keys[p] = nbr_key(j - aoff, n1, n2, n3);
            }
        }
    }
    if (A.min_dis) {
#pragma unroll
        for (int o = 16; o; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        if (lane == 0) wmin[wid] = dmin;
    }
    __syncthreads();
    const int count = nkeys;
    if (tid == 0) {
        if (A.skin_cnt) A.skin_cnt[i] = count;
        atomicMax(&A.flags->maxskin, count);
        if (count > A.cap || count > NB_MAXLIST) { atomicExch(&A.flags->overflow, 1); A.nbr_cnt[i] = 0; }
        if (A.min_dis) {
            double m = wmin[0];
            for (int w = 1; w < NB_THREADS / 32; w++) m = fmin(m, wmin[w]);
            A.min_dis[i] = m;
        }
    }
    if (count > A.cap || count > NB_MAXLIST) {
        // the list does not fit (the host enlarges it and runs again); the reference's own limit is still
        // reported from the number of neighbours within rcut (gap_calc.f90:107-111)
#pragma unroll
        for (int o = 16; o; o >>= 1) nexact += __shfl_xor_sync(0xffffffffu, nexact, o);
        if (lane == 0) wsum[wid] = nexact;
        __syncthreads();
        if (tid == 0 && !ghost) {
            int tot = 0;
            for (int w = 0; w < NB_THREADS / 32; w++) tot += wsum[w];
            atomicMax(&A.flags->maxcount, tot);
            if (tot > MAX_NEIGHBOR_REF_DEV) atomicExch(&A.flags->too_many, 1);
        }
        return;
    }
    // bitonic sort of the keys (padded to a power of two with +inf keys)
    int n2 = 1;
    while (n2 < count) n2 <<= 1;
    if (n2 <= NB_THREADS) {
        // one key per thread: compare-exchange distances below 32 go through warp shuffles,
        // only the few larger ones through shared memory
        uint64_t key = tid < count ? keys[tid] : ~0ull;
        for (int k = 2; k <= n2; k <<= 1)
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                uint64_t other;
                if (jj >= 32) {
                    __syncthreads();
                    keys[tid] = key;
                    __syncthreads();
                    other = keys[tid ^ jj];
                } else {
                    other = __shfl_xor_sync(0xffffffffu, key, jj);
                }
                const bool keep_min = ((tid & jj) == 0) == ((tid & k) == 0);
                key = keep_min ? (key < other ? key : other) : (key < other ? other : key);
            }
        __syncthreads();
        keys[tid] = key;
    } else {
        for (int p = count + tid; p < n2; p += NB_THREADS) keys[p] = ~0ull;
        __syncthreads();
        for (int k = 2; k <= n2; k <<= 1)
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int t = tid; t < n2; t += NB_THREADS) {
                    int p = t ^ jj;
                    if (p > t) {
                        uint64_t a = keys[t], b = keys[p];
                        bool up = ((t & k) == 0);
                        if ((a > b) == up) { keys[t] = b; keys[p] = a; }
                    }
                }
                __syncthreads();
            }
    }
    __syncthreads();
    if (A.skin_keys) for (int p = tid; p < count; p += NB_THREADS) A.skin_keys[(size_t)i * A.cap + p] = keys[p];
#pragma unroll
    for (int o = 16; o; o >>= 1) nclose += __shfl_xor_sync(0xffffffffu, nclose, o);
    if (lane == 0 && nclose && !ghost) atomicAdd(&A.flags->close_pairs, nclose);
    if (A.nbr_keys == nullptr) {   // bond-length query: counts and flags only
        if (tid == 0) { atomicMax(&A.flags->maxcount, count); }
        return;
    }
    int nclose2;
    const int exact = refilter_list<NB_THREADS>(keys, count, i, pos, ntot, lat, aoff, xi, yi, zi, A.rcut,
                                                A.nbr_keys + (size_t)i * A.cap, wsum, nclose2);
    if (tid == 0) {
        A.nbr_cnt[i] = exact;
        if (!ghost) {
            atomicMax(&A.flags->maxcount, exact);
            if (exact > MAX_NEIGHBOR_REF_DEV) atomicExch(&A.flags->too_many, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K1, block form: one CTA per BLOCK of cells (up to 2 x 2 x 2 bins, ~20 centres) instead of one per centre.
// The images in the bins around the block -- every candidate any centre of the block can have -- are
// collected ONCE, as keys (j, cell shift - wrap offset of j), and sorted ONCE.  Adding a centre's own wrap
// offset to the shift fields turns such a key into the centre's reference key (j, n1, n2, n3) without
// changing the order, so a warp that walks the sorted candidates and appends the survivors with ballots
// emits both lists of its centre already in reference order: no per-centre sort, no atomics, the
// candidates' coordinates staged once in shared memory.  Same tests, same arithmetic as k_neigh (which
// stays as the fallback for blocks with more candidates than fit and for exotic cell shapes).
constexpr int KB_THREADS = 256, KB_WARPS = KB_THREADS / 32, KB_CAP = 1024, KB_BINS = 512, KB_GROUP = 128;
struct NeighBlockArgs {
    NeighArgs n;
    const int *blk_struct;   // structure of every block (null: a single structure)
    double t2skin, t2cut, t2close;   // squared-distance thresholds equivalent to dis <= rskin, dis <= rcut, dis < 0.5
    float f2pre;             // single-precision pre-filter: (rskin + margin)^2
};
// dynamic shared memory of k_neigh_block
struct KbSmem {
    uint64_t ck[KB_CAP];                       // sorted candidate keys
    double cx[KB_CAP], cy[KB_CAP], cz[KB_CAP]; // their raw coordinates (exact test)
    float fx[KB_CAP], fy[KB_CAP], fz[KB_CAP];  // their image positions relative to the block, single precision (pre-filter)
    int cb_start[KB_BINS], cb_off[KB_BINS + 1], cb_shift[KB_BINS];
    unsigned short surv[KB_WARPS][KB_GROUP];   // per warp: candidates of the current group that pass the pre-filter
    double lat[9], org[3];
    int own_start[8], own_off[9];
    int flags[6];
    int ncand_owned;                           // candidates with an owned atom j (a prefix of the sorted keys)
};

__global__ void __launch_bounds__(KB_THREADS) k_neigh_block(const NeighBlockArgs B) {
    const NeighArgs &A = B.n;
    extern __shared__ __align__(16) unsigned char kb_raw[];
    KbSmem &S = *reinterpret_cast<KbSmem *>(kb_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const StructDev &s = A.structs[B.blk_struct ? B.blk_struct[blockIdx.x] : 0];
    const int lb = blockIdx.x - s.blk_off;
    const int nb0 = s.nbin[0], nb1 = s.nbin[1], nb2 = s.nbin[2];
    const int B0 = (lb / (s.nblk[1] * s.nblk[2])) * s.bs[0], B1 = ((lb / s.nblk[2]) % s.nblk[1]) * s.bs[1], B2 = (lb % s.nblk[2]) * s.bs[2];
    const int e0 = min(B0 + s.bs[0], nb0), e1 = min(B1 + s.bs[1], nb1), e2 = min(B2 + s.bs[2], nb2);   // block = bins [B, e)
    const int ntot = A.ntot;
    const bool open = s.open != 0;
    if (tid < 9) S.lat[tid] = s.lat[tid];
    if (tid < 3) {
        // a reference point near the block (its lower corner in the wrapped frame): single precision is only
        // ever applied to differences from it, so the pre-filter's error stays ~1e-6 A
        const double f0 = s.org[0] + s.wid[0] * B0 / nb0, f1 = s.org[1] + s.wid[1] * B1 / nb1, f2 = s.org[2] + s.wid[2] * B2 / nb2;
        S.org[tid] = f0 * s.lat[tid] + f1 * s.lat[3 + tid] + f2 * s.lat[6 + tid];
    }
    if (tid < 6) S.flags[tid] = 0;
    // ---- centres of the block
    const int o1 = e1 - B1, o2 = e2 - B2, nown = (e0 - B0) * o1 * o2;
    __syncwarp();
    if (tid < 8) {
        int cnt = 0, own = 0;
        if (tid < nown) {
            const int id = ((B0 + tid / (o1 * o2)) * nb1 + (B1 + (tid / o2) % o1)) * nb2 + (B2 + tid % o2);
            S.own_start[tid] = A.bin_start[s.bin_off + id];
            cnt = A.bin_start[s.bin_off + id + 1] - S.own_start[tid];
            own = A.bin_nown ? A.bin_nown[s.bin_off + id] : cnt;
        }
        S.own_off[tid] = cnt;
        if (own) atomicAdd(&S.flags[5], own);      // centres of the block that this rank owns (decomposed runs)
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < 8; k++) { const int v = S.own_off[k]; S.own_off[k] = run; run += v; }
        S.own_off[8] = run;
    }
    __syncthreads();
    const int ncentres = S.own_off[8];
    if (ncentres == 0) return;
    // A block of the ghost shell (no owned centre): its centres only list OWNED neighbours (what the ghost part
    // of the force gather needs), so only the owned records of the candidate bins -- they come first in every
    // cell -- are collected and sorted: most shell blocks then hold few candidates or none.
    const bool all_ghost = A.bin_nown != nullptr && S.flags[5] == 0;
    // ---- candidate bins: [B - m, e - 1 + m] per direction, wrapped (periodic) or clipped (open region)
    const int m0 = s.mscan[0], m1 = s.mscan[1], m2 = s.mscan[2];
    const int w0 = e0 - B0 + 2 * m0, w1 = e1 - B1 + 2 * m1, w2 = e2 - B2 + 2 * m2;
    const int ncb = w0 * w1 * w2;      // host guarantees <= KB_BINS
    for (int t = tid; t < KB_BINS; t += KB_THREADS) {
        int cnt = 0;
        if (t < ncb) {
            const int t0 = B0 - m0 + t / (w1 * w2), t1 = B1 - m1 + (t / w2) % w1, t2 = B2 - m2 + t % w2;
            int s0 = 0, s1 = 0, s2 = 0;
            bool inside = true;
            if (open) inside = t0 >= 0 && t0 < nb0 && t1 >= 0 && t1 < nb1 && t2 >= 0 && t2 < nb2;
            else { s0 = floordiv_i(t0, nb0); s1 = floordiv_i(t1, nb1); s2 = floordiv_i(t2, nb2); }
            if (inside) {
                const int id = ((t0 - s0 * nb0) * nb1 + (t1 - s1 * nb1)) * nb2 + (t2 - s2 * nb2);
                const int start = A.bin_start[s.bin_off + id];
                cnt = all_ghost ? A.bin_nown[s.bin_off + id] : A.bin_start[s.bin_off + id + 1] - start;
                S.cb_start[t] = start;
                S.cb_shift[t] = ((s0 + 512) << 20) | ((s1 + 512) << 10) | (s2 + 512);
            }
        }
        S.cb_off[t] = cnt;
    }
    __syncthreads();
    if (wid == 0) {  // exclusive scan of the bin counts by one warp
        int run = 0;
        for (int base = 0; base < KB_BINS; base += 32) {
            const int v = S.cb_off[base + lane];
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            S.cb_off[base + lane] = run + x - v;
            run += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) S.cb_off[KB_BINS] = run;
    }
    __syncthreads();
    const int ncand = S.cb_off[KB_BINS];
    if (ncand > KB_CAP) {       // too dense for this form: the host switches to k_neigh and runs again
        if (tid == 0) atomicExch(&A.flags->blk_overflow, 1);
        return;
    }
    // ---- candidate keys: (j, shift of the bin - wrap offset of j), then one sort for the whole block
    for (int c = tid; c < ncand; c += KB_THREADS) {
        int lo = 0, hi = ncb;    // bin of candidate c: last bin with cb_off <= c
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.cb_off[mid] <= c) lo = mid; else hi = mid; }
        const int4 rec = A.sabin[S.cb_start[lo] + (c - S.cb_off[lo])];      // (atom, wrap offsets)
        const int sh = S.cb_shift[lo];
        const int c0 = ((sh >> 20) & 1023) - rec.y, c1 = ((sh >> 10) & 1023) - rec.z, c2 = (sh & 1023) - rec.w;   // still offset by +512
        // a field outside 0..1023 cannot be a neighbour of anything inside the reference's image window
        const bool ok = (unsigned)c0 < 1024u && (unsigned)c1 < 1024u && (unsigned)c2 < 1024u;
        S.ck[c] = ok ? (((uint64_t)(uint32_t)rec.x << 32) | (uint64_t)((c0 << 20) | (c1 << 10) | c2)) : ~0ull;
    }
    int n2 = 32;
    while (n2 < ncand) n2 <<= 1;
    for (int c = ncand + tid; c < n2; c += KB_THREADS) S.ck[c] = ~0ull;
    __syncthreads();
    // bitonic sort, one compare-exchange per thread and trip (pair q -> elements t and t + jj, t = q with a 0 inserted at bit jj)
    for (int k = 2; k <= n2; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int q = tid; q < (n2 >> 1); q += KB_THREADS) {
                const int t = ((q & ~(jj - 1)) << 1) | (q & (jj - 1));
                const uint64_t a = S.ck[t], b = S.ck[t + jj];
                if ((a > b) == ((t & k) == 0)) { S.ck[t] = b; S.ck[t + jj] = a; }
            }
            __syncthreads();
        }
    for (int c = tid; c < ncand; c += KB_THREADS) {
        const uint64_t key = S.ck[c];
        float gx = 1e30f, gy = 1e30f, gz = 1e30f;
        if (key != ~0ull) {
            const int j = (int)(key >> 32);
            const uint32_t f = (uint32_t)key;
            const double x = A.pos[j], y = A.pos[ntot + j], z = A.pos[2 * ntot + j];
            S.cx[c] = x; S.cy[c] = y; S.cz[c] = z;
            const double q1 = (double)((int)((f >> 20) & 1023) - 512), q2 = (double)((int)((f >> 10) & 1023) - 512), q3 = (double)((int)(f & 1023) - 512);
            gx = (float)(x + q1 * S.lat[0] + q2 * S.lat[3] + q3 * S.lat[6] - S.org[0]);
            gy = (float)(y + q1 * S.lat[1] + q2 * S.lat[4] + q3 * S.lat[7] - S.org[1]);
            gz = (float)(z + q1 * S.lat[2] + q2 * S.lat[5] + q3 * S.lat[8] - S.org[2]);
        }
        S.fx[c] = gx; S.fy[c] = gy; S.fz[c] = gz;
    }
    if (tid == 0) {
        // the keys are sorted by atom index first and a rank's owned atoms come first: the candidates a GHOST
        // centre may list are a prefix of the array
        int lo = 0, hi = ncand;
        if (A.bin_nown) {
            const uint64_t bound = (uint64_t)(uint32_t)A.n_own << 32;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (S.ck[mid] < bound) lo = mid + 1; else hi = mid; }
        } else lo = ncand;
        S.ncand_owned = lo;
    }
    __syncthreads();
    // ---- every warp walks the sorted candidates for its centres: groups of 128 candidates pass a cheap
    // single-precision pre-filter (a strict superset of dis <= rskin), the survivors are compacted in order
    // and take the exact test with the reference's arithmetic
    const int na0 = s.nabc[0], na1 = s.nabc[1], na2 = s.nabc[2];
    const int aoff = s.atom_off, cap = A.cap;
    const unsigned ltmask = (1u << lane) - 1u;
    const double *lat = S.lat;
    unsigned short *sv = S.surv[wid];
    int w_maxskin = 0, w_maxcnt = 0, w_close = 0, w_many = 0, w_over = 0;
    for (int kc = wid; kc < ncentres; kc += KB_WARPS) {
        int ob = 0;
        while (S.own_off[ob + 1] <= kc) ob++;
        const int4 ci = A.sabin[S.own_start[ob] + (kc - S.own_off[ob])];
        const int i = ci.x;
        const bool ghost = i >= A.n_own;
        const double xi = A.pos[i], yi = A.pos[ntot + i], zi = A.pos[2 * ntot + i];
        // the centre in the same wrapped frame, relative to the block's reference point
        const float hx = (float)(xi - ((double)ci.y * lat[0] + (double)ci.z * lat[3] + (double)ci.w * lat[6]) - S.org[0]);
        const float hy = (float)(yi - ((double)ci.y * lat[1] + (double)ci.z * lat[4] + (double)ci.w * lat[7]) - S.org[1]);
        const float hz = (float)(zi - ((double)ci.y * lat[2] + (double)ci.z * lat[5] + (double)ci.w * lat[8]) - S.org[2]);
        uint64_t *skin = A.skin_keys ? A.skin_keys + (size_t)i * cap : nullptr;
        uint64_t *exact = A.nbr_keys ? A.nbr_keys + (size_t)i * cap : nullptr;
        int ns = 0, ne = 0, nclose = 0;
        double d2min = 1e300;
        const int ncand_i = ghost ? S.ncand_owned : ncand;
        for (int g0 = 0; g0 < ncand_i; g0 += KB_GROUP) {
            int nsv = 0;
#pragma unroll
            for (int u = 0; u < KB_GROUP / 32; u++) {
                const int c = g0 + u * 32 + lane;
                bool pass = false;
                if (c < ncand_i) {
                    const float dx = S.fx[c] - hx, dy = S.fy[c] - hy, dz = S.fz[c] - hz;
                    pass = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= B.f2pre;
                }
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                if (pass) sv[nsv + __popc(m & ltmask)] = (unsigned short)c;
                nsv += __popc(m);
            }
            __syncwarp();
            for (int base = 0; base < nsv; base += 32) {
                bool ks = false, ke = false;
                uint64_t key = 0;
                if (base + lane < nsv) {
                    const int c = sv[base + lane];
                    const uint64_t k0 = S.ck[c];
                    const int j = (int)(k0 >> 32);
                    const uint32_t f = (uint32_t)k0;
                    const int n1 = (int)((f >> 20) & 1023) - 512 + ci.y, n2s = (int)((f >> 10) & 1023) - 512 + ci.z, n3 = (int)(f & 1023) - 512 + ci.w;
                    bool ok = !(j == i && n1 == 0 && n2s == 0 && n3 == 0) && abs(n1) <= na0 && abs(n2s) <= na1 && abs(n3) <= na2;
                    if (ghost) ok = ok && j < A.n_own;
                    if (ok) {
                        const double d2 = image_dist2_xyz(S.cx[c], S.cy[c], S.cz[c], lat, n1, n2s, n3, xi, yi, zi);
                        ks = d2 <= B.t2skin;
                        ke = d2 <= B.t2cut;
                        if (ke) { d2min = fmin(d2min, d2); nclose += d2 <= B.t2close; }
                        key = nbr_key(j - aoff, n1, n2s, n3);
                    }
                }
                const unsigned ms = __ballot_sync(0xffffffffu, ks), me = __ballot_sync(0xffffffffu, ke);
                if (ks && skin) { const int p = ns + __popc(ms & ltmask); if (p < cap) skin[p] = key; }
                if (ke && exact) { const int p = ne + __popc(me & ltmask); if (p < cap) exact[p] = key; }
                ns += __popc(ms); ne += __popc(me);
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { nclose += __shfl_xor_sync(0xffffffffu, nclose, o); d2min = fmin(d2min, __shfl_xor_sync(0xffffffffu, d2min, o)); }
        if (lane == 0) {
            if (A.skin_cnt) A.skin_cnt[i] = ns;
            A.nbr_cnt[i] = ns > cap ? 0 : ne;
            if (A.min_dis) A.min_dis[i] = d2min < 1e299 ? __dsqrt_rn(d2min) : 1e300;   // sqrt is monotone: the root of the smallest d2
        }
        w_maxskin = max(w_maxskin, ns);
        if (ns > cap) w_over = 1;
        if (!ghost) {
            w_maxcnt = max(w_maxcnt, ne);
            w_close += nclose;
            if (ne > MAX_NEIGHBOR_REF_DEV) w_many = 1;
        }
    }
    if (lane == 0) {
        atomicMax(&S.flags[0], w_maxskin); atomicMax(&S.flags[1], w_maxcnt);
        if (w_many) S.flags[2] = 1;
        if (w_over) S.flags[3] = 1;
        if (w_close) atomicAdd(&S.flags[4], w_close);
    }
    __syncthreads();
    if (tid == 0) {
        atomicMax(&A.flags->maxskin, S.flags[0]);
        atomicMax(&A.flags->maxcount, S.flags[1]);
        if (S.flags[2]) atomicExch(&A.flags->too_many, 1);
        if (S.flags[3]) atomicExch(&A.flags->overflow, 1);
        if (S.flags[4]) atomicAdd(&A.flags->close_pairs, S.flags[4]);
    }
}

// Verlet reuse (gapcu_ctx_update_positions): the skin list is kept, only the exact list is
// re-derived from the new positions.  Thread 0 also checks how far the centre has moved since the
// skin list was built: beyond skin/2 a pair could have entered rcut without being listed, and the
// host rebuilds (flags->stale).
__global__ void __launch_bounds__(NB_THREADS, 8)
k_refilter(const StructDev *structs, const int *sid, const double *pos, const double *pos_build, int ntot, const int *nloc,
           int n_own, double rcut, double half_skin2, int cap, const uint64_t *skin_keys, const int *skin_cnt,
           uint64_t *nbr_keys, int *nbr_cnt, DevFlags *flags) {
    __shared__ double lat[9];
    __shared__ int wsum[NB_THREADS / 32 + 1];
    const int i = blockIdx.x;
    if (nloc && i >= *nloc) return;
    const bool ghost = i >= n_own;
    const int tid = threadIdx.x, lane = tid & 31;
    const StructDev &s = structs[sid[i]];
    if (tid < 9) lat[tid] = s.lat[tid];
    __syncthreads();
    const double xi = pos[i], yi = pos[ntot + i], zi = pos[2 * ntot + i];
    if (tid == 0 && !ghost) {
        const double dx = xi - pos_build[i], dy = yi - pos_build[ntot + i], dz = zi - pos_build[2 * ntot + i];
        if (dx * dx + dy * dy + dz * dz > half_skin2) atomicExch(&flags->stale, 1);
    }
    const int ncand = min(skin_cnt[i], cap);
    int nclose = 0;
    const int exact = refilter_list<NB_THREADS>(skin_keys + (size_t)i * cap, ncand, i, pos, ntot, lat, s.atom_off, xi, yi, zi,
                                                rcut, nbr_keys + (size_t)i * cap, wsum, nclose);
#pragma unroll
    for (int o = 16; o; o >>= 1) nclose += __shfl_xor_sync(0xffffffffu, nclose, o);
    if (lane == 0 && nclose && !ghost) atomicAdd(&flags->close_pairs, nclose);
    if (tid == 0) {
        nbr_cnt[i] = exact;
        if (!ghost) {
            atomicMax(&flags->maxcount, exact);
            if (exact > MAX_NEIGHBOR_REF_DEV) atomicExch(&flags->too_many, 1);
        }
        atomicMax(&flags->maxskin, ncand);
    }
}

// Centres ordered by descending neighbour count (counting sort; single CTA).  The order
// only schedules the persistent centre kernel; results do not depend on it.  The centres are the
// atoms [0, n) (decomposed runs: this rank's owned atoms).
// The body works for any CTA size NT that divides 1024: thread t owns the NT-th part of the 1024
// keys.  hist: NB_MAXLIST + 2 ints, wsum: 32 ints of shared memory.
template <int NT>
__device__ void order_body(const int *nbr_cnt, int n, int *order, DevFlags *flags, int *hist, int *wsum) {
    constexpr int PER = 1024 / NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int t = tid; t < NB_MAXLIST + 2; t += NT) hist[t] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) atomicAdd(&hist[NB_MAXLIST - min(max(nbr_cnt[i], 1), NB_MAXLIST)], 1);
    __syncthreads();
    // exclusive scan of the 1024 keys (PER consecutive keys per thread), key 0 = largest count
    int loc[PER], v = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { loc[k] = hist[tid * PER + k]; v += loc[k]; }
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int s = lane < NT / 32 ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        wsum[lane] = s;
    }
    __syncthreads();
    int excl = (wid ? wsum[wid - 1] : 0) + x - v;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int key = tid * PER + k;
        hist[key] = excl;
        // the owner of key(count c) knows the number of centres with more than c neighbours
        if (key == NB_MAXLIST - 128) flags->n_gt[0] = excl;
        if (key == NB_MAXLIST - 256) flags->n_gt[1] = excl;
        if (key == NB_MAXLIST - 512) flags->n_gt[2] = excl;
        if (key == 0) flags->n_gt[3] = 0;
        excl += loc[k];
        if (key == NB_MAXLIST - 1) flags->n_centres = excl;   // total number of centres
    }
    __syncthreads();
    // NOTE the places inside one key come from atomics: the order of equal-count centres varies from
    // run to run, which only changes which CTA evaluates which centre
    for (int i = tid; i < n; i += NT) {
        const int key = NB_MAXLIST - min(max(nbr_cnt[i], 1), NB_MAXLIST);   // 0..1023 (0 and 1 neighbours share a key)
        order[atomicAdd(&hist[key], 1)] = i;
    }
}

__global__ void __launch_bounds__(1024) k_order_by_count(const int *nbr_cnt, int n, int *order, DevFlags *flags) {
    __shared__ int hist[NB_MAXLIST + 2];
    __shared__ int wsum[32];
    order_body<1024>(nbr_cnt, n, order, flags, hist, wsum);
}

// Small cells (an MD cell of tens of atoms): the reference's own double loop, one CTA per centre.
// Candidate c = (j, n1, n2, n3) in the reference's loop order (gap_calc.f90:90-96: j, then n1, n2,
// n3, the last fastest) IS the sorted order of the keys, so every thread tests a contiguous run of
// at most 32 candidates, keeps a bit mask, and one block-wide prefix sum places the survivors:
// no cell list, no sort, no atomics on the list.  Raw (unwrapped) positions and the +-nabc window,
// exactly as in the reference.  The CTA that finishes last orders the centres by neighbour count
// for the centre kernel (a handful of atoms: an own launch would cost more than the work).
constexpr int DIRECT_MAX_CANDIDATES = 32 * NB_THREADS;
__global__ void __launch_bounds__(NB_THREADS)
k_neigh_direct(const StructDev *structs, const int *sid, const double *pos, int ntot, double rcut, double rskin, int cap,
               uint64_t *skin_keys, int *skin_cnt, uint64_t *nbr_keys, int *nbr_cnt, double *min_dis, DevFlags *flags,
               int *order) {
    __shared__ double lat[9];
    __shared__ int wcnt[NB_THREADS / 32], wskin[NB_THREADS / 32], wclose[NB_THREADS / 32];
    __shared__ double wmin[NB_THREADS / 32];
    __shared__ int hist[NB_MAXLIST + 2 + 32];
    __shared__ int s_last;
    const int i = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const StructDev &s = structs[sid[i]];
    if (tid < 9) lat[tid] = s.lat[tid];
    __syncthreads();
    const double xi = pos[i], yi = pos[ntot + i], zi = pos[2 * ntot + i];
    const int na0 = s.nabc[0], na1 = s.nabc[1], na2 = s.nabc[2];
    const int w2 = 2 * na2 + 1, w12 = (2 * na1 + 1) * w2, W = (2 * na0 + 1) * w12;
    const int aoff = s.atom_off, il = i - aoff;
    const int total = s.natoms * W;
    const int per = (total + NB_THREADS - 1) / NB_THREADS;   // <= 32 (host checks DIRECT_MAX_CANDIDATES)
    const int c0 = tid * per, c1 = min(total, c0 + per);
    unsigned mask = 0, smask = 0;   // exact (dis <= rcut) and skin (dis <= rskin) survivors of this thread's run
    int nclose = 0;
    double dmin = 1e300;
    if (c0 < c1) {
        int jl = c0 / W, w = c0 - jl * W;
        double xj = pos[aoff + jl], yj = pos[ntot + aoff + jl], zj = pos[2 * ntot + aoff + jl];
        for (int c = c0; c < c1; c++) {
            const int n1 = w / w12 - na0, r = w % w12, n2 = r / w2 - na1, n3 = r % w2 - na2;
            if (!(jl == il && n1 == 0 && n2 == 0 && n3 == 0)) {
                double ox, oy, oz;
                const double dis = image_distance_xyz(xj, yj, zj, lat, n1, n2, n3, xi, yi, zi, ox, oy, oz);
                if (!(dis > rskin)) smask |= 1u << (c - c0);
                if (!(dis > rcut)) {
                    mask |= 1u << (c - c0);
                    nclose += dis < 0.5;
                    dmin = fmin(dmin, dis);
                }
            }
            if (++w == W && c + 1 < c1) {
                w = 0; jl++;
                xj = pos[aoff + jl]; yj = pos[ntot + aoff + jl]; zj = pos[2 * ntot + aoff + jl];
            }
        }
    }
    // block-wide exclusive prefix sums of the kept counts (exact and skin)
    const int mine = __popc(mask), smine = __popc(smask);
    int x = mine, y = smine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, x, o), v = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o) { x += u; y += v; }
    }
    int ncl = nclose;
#pragma unroll
    for (int o = 16; o; o >>= 1) { ncl += __shfl_xor_sync(0xffffffffu, ncl, o); dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o)); }
    if (lane == 31) { wcnt[wid] = x; wskin[wid] = y; }
    if (lane == 0) { wclose[wid] = ncl; wmin[wid] = dmin; }
    __syncthreads();
    int base = x - mine, count = 0, sbase = y - smine, scount = 0;
#pragma unroll
    for (int w = 0; w < NB_THREADS / 32; w++) {
        if (w < wid) { base += wcnt[w]; sbase += wskin[w]; }
        count += wcnt[w]; scount += wskin[w];
    }
    if (tid == 0) {
        nbr_cnt[i] = count;
        if (skin_cnt) skin_cnt[i] = scount;
        atomicMax(&flags->maxcount, count);
        atomicMax(&flags->maxskin, scount);
        if (count > MAX_NEIGHBOR_REF_DEV) atomicExch(&flags->too_many, 1);
        if (scount > cap) atomicExch(&flags->overflow, 1);
        int nc = 0;
        double m = wmin[0];
        for (int w = 0; w < NB_THREADS / 32; w++) { nc += wclose[w]; m = fmin(m, wmin[w]); }
        if (nc) atomicAdd(&flags->close_pairs, nc);
        if (min_dis) min_dis[i] = m;
    }
    if (nbr_keys != nullptr && scount <= cap) {
        uint64_t *dst = nbr_keys + (size_t)i * cap + base, *sdst = skin_keys + (size_t)i * cap + sbase;
        for (unsigned m = smask; m; m &= m - 1) {
            const int bit = __ffs(m) - 1, c = c0 + bit;
            const int jl = c / W, w = c - jl * W;
            const int n1 = w / w12 - na0, r = w % w12, n2 = r / w2 - na1, n3 = r % w2 - na2;
            const uint64_t key = nbr_key(jl, n1, n2, n3);
            *sdst++ = key;
            if ((mask >> bit) & 1u) *dst++ = key;
        }
    }
    if (!order) return;
    __threadfence();          // this centre's count is visible before its ticket
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&flags->ticket, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    order_body<NB_THREADS>(nbr_cnt, ntot, order, flags, hist, hist + NB_MAXLIST + 2);
}

// The same ordering for many centres (a 10^5-atom cell): a single CTA would walk the counts alone for
// ~0.1 ms.  Histogram with per-CTA shared-memory counters flushed by global atomics, the 1024-key scan
// in one small CTA, then a grid-wide scatter.  hist: NB_MAXLIST + 2 ints of global memory.
constexpr int ORD_PER_CTA = 1024;
__global__ void __launch_bounds__(256) k_order_hist(const int *nbr_cnt, int n, int *hist) {
    __shared__ int h[NB_MAXLIST];
    for (int t = threadIdx.x; t < NB_MAXLIST; t += 256) h[t] = 0;
    __syncthreads();
    const int i0 = blockIdx.x * ORD_PER_CTA, i1 = min(n, i0 + ORD_PER_CTA);
    for (int i = i0 + threadIdx.x; i < i1; i += 256) atomicAdd(&h[NB_MAXLIST - min(max(nbr_cnt[i], 1), NB_MAXLIST)], 1);
    __syncthreads();
    for (int t = threadIdx.x; t < NB_MAXLIST; t += 256) if (h[t]) atomicAdd(&hist[t], h[t]);
}
__global__ void __launch_bounds__(1024) k_order_scan(int *hist, DevFlags *flags) {
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int v = hist[tid];
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int s = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        wsum[lane] = s;
    }
    __syncthreads();
    const int excl = (wid ? wsum[wid - 1] : 0) + x - v;
    hist[tid] = excl;
    if (tid == NB_MAXLIST - 128) flags->n_gt[0] = excl;
    if (tid == NB_MAXLIST - 256) flags->n_gt[1] = excl;
    if (tid == NB_MAXLIST - 512) flags->n_gt[2] = excl;
    if (tid == 0) flags->n_gt[3] = 0;
    if (tid == NB_MAXLIST - 1) flags->n_centres = excl + v;
}
__global__ void __launch_bounds__(256) k_order_scatter(const int *nbr_cnt, int n, int *hist, int *order) {
    // places inside one key come from atomics: the order of equal-count centres varies from run to run,
    // which only changes which CTA evaluates which centre.  Per-CTA shared counters keep the global
    // atomics to one per (CTA, key present).
    __shared__ int h[NB_MAXLIST], base[NB_MAXLIST];
    for (int t = threadIdx.x; t < NB_MAXLIST; t += 256) h[t] = 0;
    __syncthreads();
    const int i0 = blockIdx.x * ORD_PER_CTA, i1 = min(n, i0 + ORD_PER_CTA);
    int key[ORD_PER_CTA / 256], slot[ORD_PER_CTA / 256];
#pragma unroll
    for (int k = 0; k < ORD_PER_CTA / 256; k++) {
        const int i = i0 + threadIdx.x + k * 256;
        key[k] = -1;
        if (i < i1) { key[k] = NB_MAXLIST - min(max(nbr_cnt[i], 1), NB_MAXLIST); slot[k] = atomicAdd(&h[key[k]], 1); }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NB_MAXLIST; t += 256) if (h[t]) base[t] = atomicAdd(&hist[t], h[t]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ORD_PER_CTA / 256; k++)
        if (key[k] >= 0) order[base[key[k]] + slot[k]] = i0 + threadIdx.x + k * 256;
}

void launch_order(cudaStream_t st, const int *nbr_cnt, int n_centres, int *order, DevFlags *flags, int *hist, long *launches) {
    if (n_centres <= 8192 || !hist) {
        k_order_by_count<<<1, 1024, 0, st>>>(nbr_cnt, n_centres, order, flags);
        if (launches) *launches += 1;
        return;
    }
    const int g = (n_centres + ORD_PER_CTA - 1) / ORD_PER_CTA;
    cudaMemsetAsync(hist, 0, sizeof(int) * NB_MAXLIST, st);
    k_order_hist<<<g, 256, 0, st>>>(nbr_cnt, n_centres, hist);
    k_order_scan<<<1, 1024, 0, st>>>(hist, flags);
    k_order_scatter<<<g, 256, 0, st>>>(nbr_cnt, n_centres, hist, order);
    if (launches) *launches += 3;
}

// ---- host launchers ------------------------------------------------------
int neighbor_direct_max_candidates() { return DIRECT_MAX_CANDIDATES; }

void launch_neighbor_build(cudaStream_t st, const NeighborBuild &b, long *launches) {
    const int ntot = b.ntot;
    if (b.direct) {
        // small cells: one launch does what binning, cell scan, sort and centre ordering do otherwise
        k_neigh_direct<<<ntot, NB_THREADS, 0, st>>>(b.structs, b.sid, b.pos, ntot, b.rcut, b.rskin, b.cap, b.skin_keys, b.skin_cnt,
                                                    b.nbr_keys, b.nbr_cnt, b.min_dis, b.flags, b.order);
        if (launches) *launches += 1;
        return;
    }
    if (!b.sft && ntot <= 8192 && b.nbins_total <= SMALL_NBINS) {
        k_bin_small<<<1, 1024, 0, st>>>(b.structs, b.sid, b.pos, ntot, b.nbins_total, b.abin, b.bin_start, b.bin_atoms, b.sabin, b.spos);
        if (launches) *launches += 1;
    } else {
        cudaMemsetAsync(b.bin_count, 0, sizeof(int) * 2 * (size_t)b.nbins_total, st);
        int tb = 256, gb = (ntot + tb - 1) / tb;
        k_bin<<<gb, tb, 0, st>>>(b.structs, b.sid, b.pos, ntot, b.nloc, b.n_own, b.sft, b.abin, b.arank, b.bin_count, b.nbins_total);
        if (b.nbins_total <= 16384) {
            k_scan_bins<<<1, 1024, 0, st>>>(b.bin_count, b.bin_start, b.nbins_total);
        } else {   // many cells: per-CTA scans, a scan of the CTA totals, then the offsets are added
            const int gs = (b.nbins_total + SCAN_PER_CTA - 1) / SCAN_PER_CTA;
            k_scan_bins_part<<<gs, 1024, 0, st>>>(b.bin_count, b.bin_start, b.nbins_total, b.arank_scratch);
            k_scan_bins_top<<<1, 1024, 0, st>>>(b.arank_scratch, gs, b.bin_start, b.nbins_total);
            k_scan_bins_add<<<gs, 1024, 0, st>>>(b.bin_start, b.nbins_total, b.arank_scratch);
            if (launches) *launches += 2;
        }
        k_fill_bins<<<gb, tb, 0, st>>>(b.structs, b.sid, b.pos, b.abin, b.arank, b.bin_start, b.bin_count, b.nbins_total, ntot, b.nloc,
                                       b.n_own, b.bin_atoms, b.sabin, b.spos);
        if (launches) *launches += 3;
    }
    NeighArgs A;
    A.structs = b.structs; A.sid = b.sid; A.pos = b.pos; A.abin = b.abin; A.bin_start = b.bin_start;
    A.bin_nown = b.sft ? b.bin_count : nullptr;
    A.sabin = b.sabin; A.spos = b.spos; A.ntot = ntot; A.nloc = b.nloc; A.n_own = b.n_own;
    A.rcut = b.rcut; A.rskin = b.rskin; A.cap = b.cap; A.skin_keys = b.skin_keys; A.skin_cnt = b.skin_cnt;
    A.nbr_keys = b.nbr_keys; A.nbr_cnt = b.nbr_cnt; A.min_dis = b.min_dis; A.flags = b.flags;
    if (b.nblocks > 0) {
        NeighBlockArgs BA;
        BA.n = A; BA.blk_struct = b.blk_struct;
        BA.t2skin = b.t2skin; BA.t2cut = b.t2cut; BA.t2close = b.t2close;
        const float pre = (float)(b.rskin * 1.0001 + 0.02);     // far above the single-precision error of the pre-filter (~1e-5 A)
        BA.f2pre = pre * pre;
        static thread_local int attr_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != attr_dev) {
            cudaFuncSetAttribute((const void *)k_neigh_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbSmem));
            attr_dev = dev;
        }
        k_neigh_block<<<b.nblocks, KB_THREADS, sizeof(KbSmem), st>>>(BA);
    } else {
        k_neigh<<<ntot, NB_THREADS, 0, st>>>(A);
    }
    if (launches) *launches += 1;
}

int neighbor_block_max_bins() { return KB_BINS; }
int neighbor_block_max_candidates() { return KB_CAP; }

void launch_refilter(cudaStream_t st, const NeighborBuild &b, const double *pos_build, double skin, long *launches) {
    k_refilter<<<b.ntot, NB_THREADS, 0, st>>>(b.structs, b.sid, b.pos, pos_build, b.ntot, b.nloc, b.n_own, b.rcut,
                                               0.25 * skin * skin, b.cap, b.skin_keys, b.skin_cnt, b.nbr_keys, b.nbr_cnt, b.flags);
    if (launches) *launches += 1;
}

}  // namespace gapcu
