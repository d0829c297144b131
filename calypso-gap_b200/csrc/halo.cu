// halo.cu -- K6: ghost-atom halo exchange of a spatially decomposed supercell.
//
// The reference evaluates one structure in one process with an O(N^2 * images) neighbour loop and
// a 48 KB/atom table (gap_calc.f90:71, 92-120); it cannot hold a 10^5-atom cell at all.  Here the
// cell is cut into g0 x g1 x g2 bricks in fractional coordinates, one rank per brick.  A rank keeps
// only its OWNED atoms; before the neighbour build it receives, from the owners, every periodic
// image that lies within rcut + skin + drift of its brick (the GHOSTS), and after the force gather it
// returns to each owner the gradient its centres put on that owner's atoms.
//
// Exchange pattern: 26 directions delta in {-1,0,1}^3.  An owned atom at brick-relative coordinate
// t (brick units, 0 <= t < 1 up to the drift allowance) is needed by the brick at mine + delta iff
// for every axis: delta = 0, or delta = -1 and t <= nu, or delta = +1 and t >= 1 - nu, where
// nu = (rcut + skin + drift) / brick width <= 1.  Bricks wrap periodically, so with 1 or 2 bricks along
// an axis several directions lead to the same rank (or to the sender itself): they carry different
// images of the atom and stay separate messages.  The image is described by an integer shift s
// with  (fractional coordinate in the receiver's frame) = pos * inv + s ; the pair shift the
// reference would use for centre i and neighbour image p is then s_p - s_i (neigh.cu), and every
// distance is still formed from the caller's raw coordinates with the reference's arithmetic.
//
// Everything is deterministic: the send list of a direction is the ordered compaction of the owned
// atoms that qualify (tile counts -> scan -> ordered fill), ghosts are appended to the local arrays
// in direction order, and returned gradients are subtracted from an owner's force in direction
// order.  Message sizes are fixed capacities learned on the first pass (count in a header), so the
// steady state needs no host synchronisation; an overflow raises a flag that all ranks see.
#include <cstdint>

#include "device_types.cuh"
#include "launch.cuh"

namespace gapcu {

constexpr int HT = 256;   // threads per CTA; a TILE of the ordered compaction is one warp = 32 consecutive owned atoms

__device__ __forceinline__ int dir_index(int d0, int d1, int d2) { return (d0 + 1) * 9 + (d1 + 1) * 3 + (d2 + 1); }

// Brick-relative coordinates of the owned atoms -> 27-bit direction masks, the atoms' own image
// shifts, per-tile counts of every direction.
__global__ void __launch_bounds__(HT)
k_halo_mask(const HaloGeom G, const double *pos, int stride, int n_own, int4 *sft, uint32_t *mask, int *tile_cnt, int ntiles,
            DevFlags *flags) {
    const int a = blockIdx.x * HT + threadIdx.x;
    uint32_t m = 0;
    if (a < n_own) {
        const double x = pos[a], y = pos[stride + a], z = pos[2 * stride + a];
        int lo[3], hi[3], s[3];
        bool far = false;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double f = x * G.inv[c] + y * G.inv[3 + c] + z * G.inv[6 + c];
            const double fl = floor(f);
            const int g = G.grid[c];
            double t = (f - fl) * g - G.mine[c];                     // brick units, relative to my brick
            const double adj = floor((t - 0.5) / g + 0.5);           // nearest periodic image of the atom to my brick
            t -= adj * g;
            s[c] = -(int)fl - (int)adj;
            lo[c] = t <= G.nu[c];
            hi[c] = t >= 1.0 - G.nu[c];
            far = far || t < -G.drift[c] || t > 1.0 + G.drift[c];
        }
        if (far) atomicExch(&flags->halo_far, 1);
        sft[a] = make_int4(sft[a].x, s[0], s[1], s[2]);
#pragma unroll
        for (int d0 = -1; d0 <= 1; d0++)
#pragma unroll
            for (int d1 = -1; d1 <= 1; d1++)
#pragma unroll
                for (int d2 = -1; d2 <= 1; d2++) {
                    const bool ok = (d0 == 0 || (d0 < 0 ? lo[0] : hi[0])) && (d1 == 0 || (d1 < 0 ? lo[1] : hi[1])) &&
                                    (d2 == 0 || (d2 < 0 ? lo[2] : hi[2]));
                    if (ok && (d0 | d1 | d2)) m |= 1u << dir_index(d0, d1, d2);
                }
        mask[a] = m;
    }
    // a tile is one warp's 32 atoms: per direction, the number of its atoms that go that way
    const int lane = threadIdx.x & 31, tile = blockIdx.x * (HT / 32) + (threadIdx.x >> 5);
    int mine = 0;
#pragma unroll
    for (int d = 0; d < 27; d++) {
        const unsigned b = __ballot_sync(0xffffffffu, (m >> d) & 1u);
        if (lane == d) mine = __popc(b);
    }
    if (lane < 27 && tile < ntiles) tile_cnt[lane * ntiles + tile] = mine;
}

// one warp per direction: exclusive scan of the tile counts, total -> flags / header, capacity check
__global__ void __launch_bounds__(32)
k_halo_scan(const int *tile_cnt, int *tile_base, int ntiles, HaloBufs B, DevFlags *flags) {
    const int d = blockIdx.x, lane = threadIdx.x;
    int run = 0;
    for (int base = 0; base < ntiles; base += 32) {
        const int t = base + lane;
        const int v = t < ntiles ? tile_cnt[d * ntiles + t] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (t < ntiles) tile_base[d * ntiles + t] = run + x - v;
        run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) {
        flags->halo_count[d] = run;
        if (run > B.cap[d]) atomicExch(&flags->halo_overflow, 1);
    }
}

// position of this lane's atom in the send list of direction d (valid where its mask bit is set)
__device__ __forceinline__ int tile_rank(bool bit, int tile_base) {
    const unsigned b = __ballot_sync(0xffffffffu, bit);
    return tile_base + __popc(b & ((1u << (threadIdx.x & 31)) - 1u));
}

// ordered fill of the 26 send buffers: header (count), then records (x, y, z, w | gid, s1, s2, s3)
__global__ void __launch_bounds__(HT)
k_halo_fill(const HaloGeom G, const double *pos, const double *wgt, const int4 *sft, int stride, int n_own,
            const uint32_t *mask, const int *tile_base, int ntiles, HaloBufs B, const DevFlags *flags) {
    const int a = blockIdx.x * HT + threadIdx.x;
    const int tile = blockIdx.x * (HT / 32) + (threadIdx.x >> 5);
    uint32_t m = 0;
    double x = 0, y = 0, z = 0, w = 0;
    int4 q = make_int4(0, 0, 0, 0);
    if (a < n_own) { m = mask[a]; x = pos[a]; y = pos[stride + a]; z = pos[2 * stride + a]; w = wgt[a]; q = sft[a]; }
    if (blockIdx.x == 0 && threadIdx.x < 27 && threadIdx.x != 13 && B.cap[threadIdx.x] > 0) {
        int *hdr = (int *)(B.send + B.soff[threadIdx.x]);
        hdr[0] = min(flags->halo_count[threadIdx.x], B.cap[threadIdx.x]);
    }
    // only the directions some atom of this warp goes to
    for (unsigned todo = __reduce_or_sync(0xffffffffu, m) & ~(1u << 13); todo; todo &= todo - 1) {
        const int d = __ffs(todo) - 1;
        if (B.cap[d] == 0) continue;
        const bool bit = (m >> d) & 1u;
        const int k = tile_rank(bit, tile_base[d * ntiles + tile]);
        if (bit && k < B.cap[d]) {
            unsigned char *rec = B.send + B.soff[d] + HALO_HDR + (size_t)k * HALO_REC;
            double *rd = (double *)rec;
            rd[0] = x; rd[1] = y; rd[2] = z; rd[3] = w;
            // receiver's frame: its brick is mine + delta, wrapped into the grid
            const int d0 = d / 9 - 1, d1 = (d / 3) % 3 - 1, d2 = d % 3 - 1;
            int4 r;
            r.x = q.x;
            r.y = q.y - G.wrap[0][d0 + 1]; r.z = q.z - G.wrap[1][d1 + 1]; r.w = q.w - G.wrap[2][d2 + 1];
            *(int4 *)(rec + 32) = r;
        }
    }
}

// received ghosts -> local arrays behind the owned atoms, in receive order; slot map for the way back
__global__ void __launch_bounds__(HT)
k_halo_unpack(HaloBufs B, int n_own, int stride, double *pos, double *wgt, int4 *sft, int *gslot, DevFlags *flags) {
    __shared__ int cnt[27], cpre[28], kpre[28];
    if (threadIdx.x < 27) {
        const int d = B.rorder[threadIdx.x];
        cnt[threadIdx.x] = d < 0 ? 0 : min(*(const int *)(B.recv + B.roff[d]), B.cap[d]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0, k = 0;
        for (int p = 0; p < 27; p++) { cpre[p] = c; kpre[p] = k; c += cnt[p]; k += B.rorder[p] < 0 ? 0 : B.cap[B.rorder[p]]; }
        cpre[27] = c; kpre[27] = k;
        if (blockIdx.x == 0) {
            flags->n_ghost = c; flags->n_loc = n_own + c;
            if (n_own + c > stride) atomicExch(&flags->halo_overflow, 1);
        }
    }
    __syncthreads();
    const int idx = blockIdx.x * HT + threadIdx.x;   // padded slot (receive order)
    if (idx >= kpre[27]) return;
    int p = 0;
    while (idx >= kpre[p + 1]) p++;
    const int k = idx - kpre[p];
    if (k >= cnt[p]) return;
    const int dst = n_own + cpre[p] + k;
    if (dst >= stride) return;
    const unsigned char *rec = B.recv + B.roff[B.rorder[p]] + HALO_HDR + (size_t)k * HALO_REC;
    const double *rd = (const double *)rec;
    pos[dst] = rd[0]; pos[stride + dst] = rd[1]; pos[2 * stride + dst] = rd[2];
    wgt[dst] = rd[3];
    sft[dst] = *(const int4 *)(rec + 32);
    gslot[dst - n_own] = idx;
}

// owners subtract the returned gradients from their forces, direction by direction (fixed order)
__global__ void __launch_bounds__(HT)
k_halo_add(int n_own, int stride, const uint32_t *mask, const int *tile_base, int ntiles, HaloBufs B, double *force) {
    const int a = blockIdx.x * HT + threadIdx.x;
    const int tile = blockIdx.x * (HT / 32) + (threadIdx.x >> 5);
    const uint32_t m = a < n_own ? mask[a] : 0u;
    unsigned todo = __reduce_or_sync(0xffffffffu, m) & ~(1u << 13);
    if (!todo) return;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (m) { fx = force[a]; fy = force[stride + a]; fz = force[2 * stride + a]; }
    for (; todo; todo &= todo - 1) {          // ascending directions: a fixed order of subtraction
        const int d = __ffs(todo) - 1;
        if (B.cap[d] == 0) continue;
        const bool bit = (m >> d) & 1u;
        const int k = tile_rank(bit, tile_base[d * ntiles + tile]);
        if (bit && k < B.cap[d]) {
            const double *gr = B.rgrad + 3 * (size_t)(B.ks[d] + k);
            fx -= gr[0]; fy -= gr[1]; fz -= gr[2];
        }
    }
    if (m) { force[a] = fx; force[stride + a] = fy; force[2 * stride + a] = fz; }
}

// this rank's record for the all-gather: raw E and strs sums, flags, send counts
__global__ void k_halo_rec(const double *out8_raw, const DevFlags *flags, double *rec) {
    const int t = threadIdx.x;
    if (t < 7) rec[t] = out8_raw[t];
    if (t == 7) rec[7] = (double)flags->overflow;
    if (t == 8) rec[8] = (double)flags->too_many;
    if (t == 9) rec[9] = (double)flags->maxcount;
    if (t == 10) rec[10] = (double)flags->maxskin;
    if (t == 11) rec[11] = (double)flags->stale;
    if (t == 12) rec[12] = (double)flags->halo_overflow;
    if (t == 13) rec[13] = (double)flags->halo_far;
    if (t == 14) rec[14] = (double)flags->close_pairs;
    if (t == 15) rec[15] = (double)flags->n_loc;
    if (t >= 16 && t < 16 + 27) rec[t] = (double)flags->halo_count[t - 16];
    if (t == 43) rec[43] = (double)flags->blk_overflow;
}

// all ranks' records -> E and stress (sums in rank order: every rank gets the same bits) and merged flags
__global__ void k_halo_combine(const double *rec_all, int nranks, double volume, double *out8, DevFlags *flags) {
    const int t = threadIdx.x;
    __shared__ double s[8];
    if (t < 7) {
        double v = 0.0;
        for (int r = 0; r < nranks; r++) v += rec_all[(size_t)r * HALO_RECLEN + t];
        s[t] = v;
    }
    __syncthreads();
    if (t == 0) {
        const double f = (1.0 / 6.24219e-3) / volume;   // gap_calc.f90:9,203
        out8[0] = s[0];
        out8[1] = -s[1] * f; out8[2] = -s[4] * f; out8[3] = -s[6] * f;   // xx yy zz  (strs sums are xx xy xz yy yz zz)
        out8[4] = -s[2] * f; out8[5] = -s[5] * f; out8[6] = -s[3] * f;   // xy yz xz  (gap_calc.f90:221-226)
        out8[7] = 0.0;
    }
    if (t >= 7 && t < 16 + 28) {
        double mx = 0.0, sum = 0.0;
        for (int r = 0; r < nranks; r++) { const double v = rec_all[(size_t)r * HALO_RECLEN + t]; mx = fmax(mx, v); sum += v; }
        const int iv = (int)mx;
        if (t == 7) flags->overflow = iv;
        if (t == 8) flags->too_many = iv;
        if (t == 9) flags->maxcount = iv;
        if (t == 10) flags->maxskin = iv;
        if (t == 11) flags->stale = iv;
        if (t == 12) flags->halo_overflow = iv;
        if (t == 13) flags->halo_far = iv;
        if (t == 14) flags->close_pairs = (int)sum;
        if (t == 15) flags->n_loc = iv;               // largest local point count of any rank
        if (t >= 16 && t < 43) flags->halo_count[t - 16] = iv;  // largest send count per direction
        if (t == 43) flags->blk_overflow = iv;
    }
}

void launch_halo_select(cudaStream_t st, const HaloGeom &G, const double *pos, int stride, int n_own, int4 *sft, uint32_t *mask,
                        int *tile_cnt, int *tile_base, const HaloBufs &B, DevFlags *flags, long *launches) {
    const int ntiles = (n_own + 31) / 32;
    k_halo_mask<<<(n_own + HT - 1) / HT, HT, 0, st>>>(G, pos, stride, n_own, sft, mask, tile_cnt, ntiles, flags);
    k_halo_scan<<<27, 32, 0, st>>>(tile_cnt, tile_base, ntiles, B, flags);
    if (launches) *launches += 2;
}

void launch_halo_fill(cudaStream_t st, const HaloGeom &G, const double *pos, const double *wgt, const int4 *sft, int stride,
                      int n_own, const uint32_t *mask, const int *tile_base, const HaloBufs &B, const DevFlags *flags, long *launches) {
    const int ntiles = (n_own + 31) / 32;
    k_halo_fill<<<(n_own + HT - 1) / HT, HT, 0, st>>>(G, pos, wgt, sft, stride, n_own, mask, tile_base, ntiles, B, flags);
    if (launches) *launches += 1;
}

void launch_halo_unpack(cudaStream_t st, const HaloBufs &B, int n_own, int stride, double *pos, double *wgt, int4 *sft, int *gslot,
                        DevFlags *flags, long *launches) {
    int total = 0;
    for (int d = 0; d < 27; d++) total += B.cap[d];
    k_halo_unpack<<<(total + HT - 1) / HT + 1, HT, 0, st>>>(B, n_own, stride, pos, wgt, sft, gslot, flags);
    if (launches) *launches += 1;
}

void launch_halo_add(cudaStream_t st, int n_own, int stride, const uint32_t *mask, const int *tile_base, const HaloBufs &B,
                     double *force, long *launches) {
    const int ntiles = (n_own + 31) / 32;
    k_halo_add<<<(n_own + HT - 1) / HT, HT, 0, st>>>(n_own, stride, mask, tile_base, ntiles, B, force);
    if (launches) *launches += 1;
}

void launch_halo_rec(cudaStream_t st, const double *out8_raw, const DevFlags *flags, double *rec, long *launches) {
    k_halo_rec<<<1, 64, 0, st>>>(out8_raw, flags, rec);
    if (launches) *launches += 1;
}

void launch_halo_combine(cudaStream_t st, const double *rec_all, int nranks, double volume, double *out8, DevFlags *flags, long *launches) {
    k_halo_combine<<<1, 64, 0, st>>>(rec_all, nranks, volume, out8, flags);
    if (launches) *launches += 1;
}

}  // namespace gapcu
