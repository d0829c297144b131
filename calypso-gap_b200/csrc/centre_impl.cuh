// centre_impl.cuh -- the per-centre wACSF kernel (K2 forward, K4 backward, or both fused
// with an in-CTA GPR): one CTA per centre atom.  Template over the mode and over the
// neighbour capacity PCAP of the CTA; centre_p*.cu instantiate it, centre.cu dispatches.
//
// What is computed (reference loops wacsf.f90:65-795):
//   radial  type 1  G = sum_j exp(-a r^2) fc(r)            wacsf.f90:69-162
//           type 3  G = sum_j exp(-4 (r-rs)^2) fc(r)       wacsf.f90:436-527
//   angular type 2/4  G = sum_{j<k} (1 +- cos) exp(-a (rij^2+rik^2+rjk^2)) fc fc fc
//                                                          wacsf.f90:169-432, 534-791
//   each in an unweighted channel ii and a species-weighted channel ii+nsf, then
//   (fused mode) e_i and dE/dG from the sparse GPR (gap_calc.f90:143-166), then the
//   chain rule dE/dG * dG/dr (gap_calc.f90:177-203) WITHOUT the reference's dense
//   dxdy(D,N,N,3): the geometry is recomputed and contracted on the fly, leaving
//   per neighbour slot dE_i/dx_slot, plus dE_i/dx_i and the centre's strs sums.
//
// Structure per centre (all orders fixed -> results are bit-reproducible):
//   1. stage neighbours in shared memory as records (x, y, z, r, 1/r, w) -- absolute image
//      coordinates in the reference's arithmetic (geom.cuh).  Tiers with one neighbour per thread
//      (PCAP <= 256) stage them ORDERED by the number of cutoff classes they belong to (classes =
//      distinct cutoffs, descending, so the classes of a distance are a prefix): the members of
//      class c are the places 0 .. pc[c]-1, and a byte permutation maps places back to list slots.
//      These arrays sit at COMPILE-TIME offsets (template PCAP), so the hot loops address them as
//      [index register + immediate];
//   2. pair tests: all pairs q = (a < b) of the flat triangular index are tested ONCE with the
//      exact reference arithmetic (squared-distance thresholds equivalent to the reference's
//      sqrt(..) > cutoff); a 32-pair trip tests as many thresholds as its rows' class count asks
//      for (one for most trips); survivors carry their "bucket" = number of classes they belong to;
//   3. while ONE warp places the (warp, bucket) runs (a short scan), the other warps fill the
//      (fc, fc') tables of every (class, member); then the survivors move to their bucket's run of
//      the sorted list S (descending bucket: the items of class c are the prefix S[0 .. npre[c]))
//      and the radial functions (one warp task per two functions) run in the same slot;
//   4. forward: class-outer loop over that prefix, per-thread register accumulators
//      per (class, alpha) group [sum pe, sum pe*cos, and the two weighted sums; the
//      lambda=+-1 functions are (sum pe +- sum pe*cos)], one warp reduction per class;
//   5. (fused) GPR for this atom in the CTA: difference form, no cancellation; the sparse set is
//      read past L1 (ld.global.cg);
//   6. backward: warps take batches of 32 triplets of the SAME bucket; per class one
//      sincos and per (class, alpha) one exp (ONE per triplet when every class carries the same
//      two exponents, fetched from the forward pass's parking buffer in MODE_FUSED_SE): the sum over
//      symmetry functions is folded into four per-centre constants per group (sum du, sum dw,
//      sum lam*du, sum lam*dw), so there is no inner loop over functions; the three leg scalars
//      go to per-warp private accumulators (dE/dx_j = A_j d_j - V_j) with in-warp
//      conflict serialisation: no atomics anywhere.
//   The kernel runs at the 80-register cap of three CTAs per SM and leaves ~30 KB of L1: values that
//   would be spilled across the hot loops (coordinates, bucket constants, the parking buffer's
//   start) are re-read from shared memory instead -- tools/spill_lines.py shows what is left.
//
// Thread-block clusters (template CS = 1, 2 or 4 CTAs per centre): when a launch has fewer
// centres than the device has CTA slots (a 64-atom MD cell uses 64 of 444), CS CTAs of one
// cluster share a centre.  Every CTA stages the neighbours, then takes 1/CS of the pair
// range (own triplet list, own forward sums, own backward accumulators) and 1/CS of the
// radial functions; the partial descriptors are exchanged through distributed shared memory
// (each CTA adds the CS partial vectors in rank order, so all hold the same bits), the small
// GPR is evaluated by every CTA, and rank 0 adds the CS gradient accumulators in rank order
// and writes the centre's outputs.  Three cluster barriers per centre.
#pragma once
#include <cooperative_groups.h>

#include <cstdint>
#include <cstring>

#include "device_types.cuh"
#include "fastmath.cuh"
#include "geom.cuh"
#include "launch.cuh"

namespace gapcu {
namespace cg = cooperative_groups;

#define GRP_BEGIN(a, c) (a).cls.grp_begin[c]   // first group of class c / one past its last group

constexpr int CT = 256;  // threads per centre CTA (measured: 384 threads x 2 CTAs/SM is no faster, barrier stalls grow)
constexpr int NW = CT / 32;
constexpr int MAXG = 4;  // alpha groups of one class handled per forward pass (register accumulators)

// MODE_FUSED_SE: the fused kernel for potentials whose angular cutoff classes all carry the same one or two
// exponents alpha (the shipped table: 0.01 and 0.1 for every cutoff).  exp(-alpha s) of a triplet then does
// not depend on the class: it is evaluated ONCE, in the forward pass of the first angular class, parked in
// an L2-resident per-CTA buffer, and read back by the other classes' forward passes and by the backward
// pass (which otherwise repeats every exponential of the forward pass).
enum { MODE_FWD = 0, MODE_BWD = 1, MODE_FUSED = 2, MODE_FUSED_SE = 3 };

// Byte offsets of the per-neighbour arrays at the start of the dynamic shared memory.
template <int PCAP>
struct Hot {
    static constexpr int NB = 0;                 // [PCAP] records of 6 doubles: x y z r 1/r w
    static constexpr int ACC = NB + 48 * PCAP;   // [3][PCAP] dE_i/dx of every neighbour slot
    static constexpr int NC = ACC + 24 * PCAP;   // [PCAP] bytes: number of classes the neighbour belongs to
    static constexpr int PERM = NC + PCAP;       // [PCAP] bytes: list slot of the neighbour staged at this place (sorted staging)
    static constexpr int FCD = PERM + PCAP;      // [ncls][PCAP] (fc, fc') pairs
};
inline int hot_bytes(int pcap_t, int ncls) { return 48 * pcap_t + 24 * pcap_t + 2 * pcap_t + 16 * ncls * pcap_t; }

// control block in shared memory
struct __align__(16) Ctl {
    int cntw[NW];               // kept items per warp segment
    int hw[NW][MAXC_DEV + 1];   // kept items per (warp, bucket)
    int basew[NW][MAXC_DEV + 1];// running output position per (warp, bucket)
    int tot[MAXC_DEV + 1];      // items per bucket
    int npre[MAXC_DEV + 1];     // items with bucket > c  (prefix length of class c)
    // order slot o (= bucket ncls-o, heavy buckets first): x = batches of 32, y = items, z = start in S, w = number
    // of its first batch; one 128-bit load per batch of the backward pass
    int4 ob[MAXC_DEV + 2];
    int TB, nkept;
    int pc[MAXC_DEV + 1];       // neighbours that belong to class c (sorted staging: the places 0 .. pc[c]-1)
    int estb;                   // MODE_FUSED_SE: first entry of this CTA's (and chunk's) slice of the parked exponentials
};
static_assert(sizeof(Ctl) <= 4 * 512, "Ctl does not fit its slot");

__device__ __forceinline__ void st_shared_u32(unsigned addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum four per-lane values over the warp with 6 instead of 20 value shuffles: after the
// call the lanes with (lane>>3) == k hold the total of v_k.
__device__ __forceinline__ double warp_sum4(double v0, double v1, double v2, double v3, int lane) {
    const bool hi16 = lane & 16, hi8 = lane & 8;
    // lanes with bit4 = 0 keep (v0, v1), the others (v2, v3); each gets the partner's copy
    double k0 = hi16 ? v2 : v0, k1 = hi16 ? v3 : v1;
    const double s0 = hi16 ? v0 : v2, s1 = hi16 ? v1 : v3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    double k = hi8 ? k1 : k0;
    const double s = hi8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// pair index q -> (a < b), q = b(b-1)/2 + a
__device__ __forceinline__ void tri_decode(int q, int &a, int &b) {
    int bb = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)q)) * 0.5f);
    if (bb * (bb - 1) / 2 > q) bb--;
    if ((bb + 1) * bb / 2 <= q) bb++;
    b = bb;
    a = q - bb * (bb - 1) / 2;
}

// add a 3-vector to a [3][PCAP] accumulator at index idx: lanes of the calling warp that hit the same
// idx take turns in lane order (deterministic, no atomics).  The set is either private to the warp or,
// for very long neighbour lists, shared by the CTA -- then the caller lets one warp at a time in.
template <int PCAP>
__device__ __forceinline__ void scatter3(double *pa, int idx, double v0, double v1, double v2,
                                         unsigned amask, unsigned ltmask) {
    const unsigned peers = __match_any_sync(amask, idx);
    const int rank = __popc(peers & ltmask);
    const int maxr = __reduce_max_sync(amask, rank);
    for (int r = 0; r <= maxr; r++) {
        if (rank == r) {
            pa[idx] += v0;
            pa[PCAP + idx] += v1;
            pa[2 * PCAP + idx] += v2;
        }
        __syncwarp(amask);
    }
}

// The same for a warp-PRIVATE set laid out as (x, y) pairs [PCAP] followed by z [PCAP], through plain 32-bit
// shared addresses: two loads and two stores per turn instead of three and three, no address arithmetic inside
// the loop.  sa = shared address of the set.
template <int PCAP>
__device__ __forceinline__ void scatter3p(unsigned sa, int idx, double v0, double v1, double v2, unsigned amask, unsigned ltmask) {
    const unsigned peers = __match_any_sync(amask, idx);
    const int rank = __popc(peers & ltmask);
    const int maxr = __reduce_max_sync(amask, rank);
    const unsigned axy = sa + 16u * (unsigned)idx, az = sa + 16u * PCAP + 8u * (unsigned)idx;
    for (int r = 0; r <= maxr; r++) {
        if (rank == r) {
            double x, y, z;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(axy) : "memory");
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(z) : "r"(az) : "memory");
            x += v0; y += v1; z += v2;
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(axy), "d"(x), "d"(y) : "memory");
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(az), "d"(z) : "memory");
        }
        __syncwarp(amask);
    }
}

extern __shared__ __align__(16) unsigned char smem[];   // dynamic shared memory: Hot<PCAP> arrays, then SmemLayout
// work counters (gapcu_ctx_work_counters) of this CTA, flushed once at its exit: 29 global atomics per centre
// on ten addresses were a serial point of their own.  File scope: a fixed address, no pointer to carry.
__shared__ unsigned long long s_work[10];

// squared distance with fused multiply-adds (value only; the kept/dropped decision of a pair
// is always taken with pair_dist2, the reference's arithmetic)
__device__ __forceinline__ double dist2_fma(double dx, double dy, double dz) { return fma(dz, dz, fma(dy, dy, dx * dx)); }

template <int MODE, int PCAP, int CS>
__device__ __forceinline__ void process_centre(const CentreArgs &a, const int i, const bool first) {
    constexpr bool FWD = MODE != MODE_BWD, BWD = MODE != MODE_FWD, FUSED = MODE >= MODE_FUSED, SE = MODE == MODE_FUSED_SE;
    // rank of this CTA among the CS CTAs that share centre i; lead = the one that writes the outputs
    const int crank = CS > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const bool lead = crank == 0;
    using H = Hot<PCAP>;
    // neighbour record s: (x, y) (z, r) (1/r, w); class pair (c, s): (fc, fc')
#define NB2(s, k) (*(double2 *)(smem + H::NB + (s) * 48 + (k) * 16))
#define FCD2(c, s) (*(double2 *)(smem + H::FCD + ((c) * PCAP + (s)) * 16))
#define FCV(c, s) (*(double *)(smem + H::FCD + ((c) * PCAP + (s)) * 16))
#define NCB(s) (smem[H::NC + (s)])
#define PERM(s) (smem[H::PERM + (s)])
    // Tiers whose lists fit one neighbour per thread stage the neighbours ORDERED by the number of classes
    // they belong to (descending; ties in list order): the neighbours of class c are then the places
    // 0 .. pc[c]-1, and a pair (a < b) can only belong to the classes of b -- for two thirds of the pairs that
    // is the widest class alone, one threshold test.  PERM maps a place back to the list slot the outputs
    // (fpair) are indexed by.
#ifndef GAPCU_SORTED
#define GAPCU_SORTED 1
#endif
    constexpr bool SORTED = PCAP <= CT && GAPCU_SORTED;
    double *const s_acc = (double *)(smem + H::ACC);   // [3][PCAP]
    const PlanDev &pl = a.plan;
    const int ncls = pl.ncls, D = pl.D, nsf = pl.nsf;
    const int lcap = a.lcap;
    const SmemLayout &L = a.lay;
    double *s_t32 = (double *)(smem + L.t32);
    // the two tables of fastmath.cuh as plain shared addresses (exp_neg_s, sincos_tab_s)
    const unsigned t32_sa = (unsigned)__cvta_generic_to_shared(s_t32);
    double *s_t2 = (double *)(smem + L.t2);       // class thresholds padded with -1 (never passes)
    double *s_galpha = (double *)(smem + L.galpha);
    double *s_gd = (double *)(smem + L.gd);      // [n_grp][4]: DU, DW, DUL, DWL (backward)
    double *s_gw = (double *)(smem + L.gw);      // [NW][D] per-warp partial descriptors
    double *s_G = (double *)(smem + L.sG);
    double *s_gx = (double *)(smem + L.gx);      // this CTA's partial descriptors, read by its cluster peers (CS > 1)
    double *s_du = (double *)(smem + L.sdu);     // dE/dG of this centre
    double *s_xs = (double *)(smem + L.xs);
    double *s_W = (double *)(smem + L.sW);
    double *s_red = (double *)(smem + L.red);
    uint32_t *s_S = (uint32_t *)(smem + L.S);
    uint32_t *s_U = (uint32_t *)(smem + L.scratch);
    double *s_pa = (double *)(smem + L.scratch);  // [NW][3][PCAP] (aliases U; live only in backward phase B)
    Ctl *ctl = (Ctl *)(smem + L.ctl);
    int2 *s_radi = (int2 *)(smem + L.rad);        // per radial function: (ii, cls | type<<16)
    double *s_radp = (double *)(smem + L.rad + 8 * (pl.n_rad + 1));

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    // development aid: where a centre's time goes.  Compiled in only with -DGAPCU_PHASE_TIMING (the
    // bookkeeping costs registers the production kernel does not have); then GAPCU_VARIANT & 16 switches it on.
#ifdef GAPCU_PHASE_TIMING
    const bool ptime = (a.variant & 16) && tid == 0;
    long long pt_last = ptime ? clock64() : 0;
    auto phase_end = [&](int ph) {
        if (ptime) { const long long now = clock64(); atomicAdd(&a.flags->phase_cycles[ph], (unsigned long long)(now - pt_last)); pt_last = now; }
    };
#else
    auto phase_end = [](int) {};
#endif
    const int P = a.nbr_cnt[i];
    const StructDev &sd = a.structs[a.sid[i]];
    const int ntot = a.ntot;
    const int *g_iplus = pl.itab + pl.o_grp_iplus, *g_iminus = pl.itab + pl.o_grp_iminus;
    // descriptor indices of a group's lambda = +1 / -1 function: a shared copy for the forward sums (the global
    // table put a load from L1/L2 on the critical path of every class reduction)
    constexpr int GIDX_CAP = 64;
    __shared__ int2 s_gidx[GIDX_CAP];
    const bool gidx_ok = pl.n_grp <= GIDX_CAP;

    // ---- 0: tables (centre independent: loaded once per persistent CTA) ----------------
    __shared__ int s_nrad[MAXC_DEV], s_nf[MAXC_DEV];   // per class: radial functions, angular functions (work counters)
    __shared__ __align__(16) SinCosEntry s_trig[SINCOS_TAB_N];       // (cos, sin)(k/16) for sincos_tab
    const unsigned trig_sa = (unsigned)__cvta_generic_to_shared(s_trig);
    if (first) {
        if (tid < 2 * SINCOS_TAB_N) ((double *)s_trig)[tid] = a.exp2_table[32 + tid];
        if (tid < MAXC_DEV) {
            int nr = 0, nf = 0;
            for (int q = 0; q < pl.n_rad; q++) nr += pl.itab[pl.o_rad_cls + q] == tid;
            if (tid < ncls)
                for (int g = GRP_BEGIN(a, tid); g < GRP_BEGIN(a, tid + 1); g++) nf += (g_iplus[g] >= 0) + (g_iminus[g] >= 0);
            s_nrad[tid] = nr; s_nf[tid] = nf;
        }
        if (tid < 32) s_t32[tid] = a.exp2_table[tid];
        if (tid < GIDX_CAP && tid < pl.n_grp) s_gidx[tid] = make_int2(g_iplus[tid], g_iminus[tid]);
        if (tid < MAXC_DEV) s_t2[tid] = tid < ncls ? a.cls.t2[tid] : -1.0;
        for (int t = tid; t < pl.n_grp; t += CT) s_galpha[t] = pl.dtab[pl.o_grp_alpha + t];
        for (int t = tid; t < pl.n_rad; t += CT) {
            s_radi[t] = make_int2(pl.itab[pl.o_rad_ii + t], pl.itab[pl.o_rad_cls + t] | (pl.itab[pl.o_rad_type + t] << 16));
            s_radp[t] = pl.dtab[pl.o_rad_p + t];
        }
    }
    if (P > PCAP || P > a.cap) {  // host re-runs with a larger capacity (checked after the tables are in place: `first` is spent)
        if (tid == 0) atomicExch(&a.flags->overflow, 1);
        return;
    }
    if (FWD) for (int t = tid; t < NW * D; t += CT) s_gw[t] = 0.0;
    if (MODE == MODE_BWD) for (int t = tid; t < D; t += CT) s_du[t] = a.dEdG[(size_t)i * D + t];
    __shared__ double s_lat[9];
    __shared__ double s_ctr[3];   // the centre's position: read where needed, not carried in registers through the hot loops
    if (tid < 9) s_lat[tid] = sd.lat[tid];
    if (tid < 3) s_ctr[tid] = a.pos[tid * ntot + i];
    __syncthreads();

    const double xi = a.pos[i], yi = a.pos[ntot + i], zi = a.pos[2 * ntot + i];

    // ---- 1: stage neighbours (a: geometry per neighbour, b: fc/fc' per (neighbour, class)) ----
    auto neighbour_record = [&](int s, double &ox, double &oy, double &oz, double &dis, double &wj) {
        if (a.nbr_table) {
            // CAR2ACSF: image position, distance and weight as the caller tabulated them (wacsf.f90:80-84)
            const size_t NA = (size_t)ntot, ld = (size_t)a.table_ld;
            const double *T = a.nbr_table + i + NA * s;
            ox = T[0]; oy = T[NA * ld]; oz = T[2 * NA * ld]; dis = T[3 * NA * ld]; wj = T[4 * NA * ld];
        } else {
            int jl, n1, n2, n3;
            nbr_unkey(a.nbr_keys[(size_t)i * a.cap + s], jl, n1, n2, n3);
            const int j = sd.atom_off + jl;
            dis = image_distance(a.pos, ntot, j, s_lat, n1, n2, n3, xi, yi, zi, ox, oy, oz);
            wj = a.wgt[j];
        }
        int nc = 0;
        while (nc < ncls && !(dis > a.cls.rc[nc])) nc++;  // reference: "if (rij.gt.cutoff) cycle"
        return nc;
    };
    if constexpr (SORTED) {
        // one neighbour per thread; counting sort by class count (descending), list order within a count
        const int s = tid;
        double ox = 0.0, oy = 0.0, oz = 0.0, dis = 1.0, wj = 0.0;
        int nc = -1;
        if (s < P) nc = neighbour_record(s, ox, oy, oz, dis, wj);
        unsigned mym = 0;
        int cntv = 0;
        for (int v = 0; v <= ncls; v++) {
            const unsigned m = __ballot_sync(0xffffffffu, nc == v);
            if (lane == v) cntv = __popc(m);
            if (nc == v) mym = m;
        }
        if (lane <= ncls) ctl->hw[wid][lane] = cntv;
        __syncthreads();
        // lane v: neighbours with count v staged by the warps before this one / by all warps
        int before = 0, all = 0;
        if (lane <= ncls)
            for (int w = 0; w < NW; w++) { const int h = ctl->hw[w][lane]; before += w < wid ? h : 0; all += h; }
        int ge = all;   // -> neighbours with count >= lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_down_sync(0xffffffffu, ge, d); if (lane + d < 32) ge += y; }
        const int place = __shfl_sync(0xffffffffu, ge - all + before, nc < 0 ? 0 : nc) + __popc(mym & ltmask);
        if (s < P) {
            NB2(place, 0) = make_double2(ox, oy);
            NB2(place, 1) = make_double2(oz, dis);
            NB2(place, 2) = make_double2(1.0 / dis, wj);
            NCB(place) = (unsigned char)nc;
            PERM(place) = (unsigned char)s;
        }
        if (wid == 0 && lane <= ncls) ctl->pc[lane] = ge - all;   // count > lane: members of class `lane`
    } else {
        for (int s = tid; s < P; s += CT) {
            double ox, oy, oz, dis, wj;
            const int nc = neighbour_record(s, ox, oy, oz, dis, wj);
            NB2(s, 0) = make_double2(ox, oy);
            NB2(s, 1) = make_double2(oz, dis);
            NB2(s, 2) = make_double2(1.0 / dis, wj);
            NCB(s) = (unsigned char)nc;
        }
    }
    __syncthreads();
    if constexpr (!SORTED) {
        // members of every class (work counters; the sorted staging has them from its counting sort)
        for (int c = wid; c < ncls; c += NW) {
            int n = 0;
            for (int s = lane; s < P; s += 32) n += NCB(s) > c;
            n = __reduce_add_sync(0xffffffffu, n);
            if (lane == 0) ctl->pc[c] = n;
        }
    }
    const int P32 = (P + 31) & ~31;
    // fc / fc' of every (class, member) pair; threads t0, t0 + tstep, ... of the caller's group
    auto fc_tables = [&](int t0, int tstep) {
        for (int t = t0; t < ncls * P32; t += tstep) {
            const int c = t / P32, s = t - c * P32;
            if (s < P && c < NCB(s)) {
                const double pirc = a.cls.pirc[c];
                double sn, cs;
                sincos_tab_s(NB2(s, 1).y * pirc, trig_sa, &sn, &cs);
                FCD2(c, s) = make_double2(0.5 * (cs + 1.0), -0.5 * pirc * sn);
            }
        }
    };
    // ---- 2: radial forward: one warp task per function, lanes over neighbours ----------
    auto radial_forward = [&]() {
        // two functions per warp and pass (independent exponential chains, one transposed reduction for the four sums)
        for (int q = wid + NW * crank; q < pl.n_rad; q += 2 * NW * CS) {   // the cluster's CTAs share the functions
            const int q2 = q + NW * CS;
            const bool two = q2 < pl.n_rad;
            const int2 ri = s_radi[q], rj = s_radi[two ? q2 : q];
            const int c = ri.y & 0xffff, c2 = two ? (rj.y & 0xffff) : ncls;   // ncls: no neighbour belongs to it
            const double prm = s_radp[q], prm2 = s_radp[two ? q2 : q];
            const bool t1 = (ri.y >> 16) == 1, t2nd = (rj.y >> 16) == 1;
            double gu = 0.0, gwt = 0.0, hu = 0.0, hwt = 0.0;
            for (int s = lane; s < P; s += 32) {
                const int nc = NCB(s);
                const double dis = NB2(s, 1).y, wj = NB2(s, 2).y;
                if (c < nc) {
                    const double d = t1 ? dis : dis - prm;
                    const double g = exp_neg_s((t1 ? -prm : -4.0) * d * d, t32_sa) * FCV(c, s);
                    gu += g;
                    gwt = fma(g, wj, gwt);
                }
                if (c2 < nc) {
                    const double d = t2nd ? dis : dis - prm2;
                    const double g = exp_neg_s((t2nd ? -prm2 : -4.0) * d * d, t32_sa) * FCV(c2, s);
                    hu += g;
                    hwt = fma(g, wj, hwt);
                }
            }
            const double tot = warp_sum4(gu, gwt, hu, hwt, lane);   // lanes 0, 8, 16, 24 hold the four sums
            if ((lane & 7) == 0) {
                const int which = lane >> 3;
                if (which < 2) s_gw[wid * D + ri.x + (which ? nsf : 0)] += tot;
                else if (two) s_gw[wid * D + rj.x + ((which & 1) ? nsf : 0)] += tot;
            }
        }
    };
    // work counters of this centre: pairs per class, radial evaluations, candidate pairs
    // of the angular classes -- all functions of the classes' member counts.  One warp.
    auto count_centre = [&]() {
        const int pcv = lane < ncls ? ctl->pc[lane] : 0;
        const unsigned npc = __reduce_add_sync(0xffffffffu, (unsigned)pcv);
        const unsigned nrd = __reduce_add_sync(0xffffffffu, (unsigned)(lane < ncls ? pcv * s_nrad[lane] : 0));
        const unsigned ncq = __reduce_add_sync(0xffffffffu, (unsigned)((lane < ncls && s_nf[lane]) ? pcv * (pcv - 1) / 2 : 0));
        if (lane == 0) {
            s_work[0] += 1; s_work[1] += P; s_work[2] += npc; s_work[3] += (unsigned)(P * (P - 1) / 2);
            s_work[7] += nrd; s_work[8] += ncq;
        }
    };

    // ---- 3: triplet list builder (phase A + deterministic counting sort) ----------
    const uint32_t angmask = a.cls.angmask;
    // pair range of this CTA: the whole triangle, or one of CS contiguous parts of it
    const int Qall = P * (P - 1) / 2;
    int Qlo = 0, Qhi = Qall;
    if constexpr (CS > 1) {
        const int qpart = ((Qall + CS - 1) / CS + 31) & ~31;
        Qlo = min(Qall, crank * qpart); Qhi = min(Qall, Qlo + qpart);
    }
    const int Q = Qhi - Qlo;
    // (one chunk is the rule: no integer divisions on that path)
    const int nchunk = (angmask && Q > 0) ? (Q <= lcap ? 1 : (Q + lcap - 1) / lcap) : 0;
    const int qchunk = nchunk <= 1 ? ((Q + 31) & ~31) : (((Q + nchunk - 1) / nchunk + 31) & ~31);
    // parked exponentials: one double2 per kept pair, chunk after chunk of this centre's list
    constexpr bool se = SE;
    double2 *est = SE ? a.estash + (size_t)blockIdx.x * a.estash_stride : nullptr;

    // Part 1 (all warps): every pair of the flat triangular range [q0, q1) is tested once; survivors go to this
    // warp's part of the scratch list with their bucket (= number of classes they belong to), the warp's
    // bucket counts to ctl->hw.  Returns the warp's number of survivors.
    auto list_pairs = [&](int q0, int q1) {
        const int n = q1 - q0;
        const int R = (((n + NW - 1) / NW) + 31) & ~31;  // per-warp contiguous sub-range
        const int wq0 = q0 + wid * R, wq1 = min(q1, wq0 + R);
        uint32_t *seg = s_U + wid * R;
        int cnt = 0;
        const bool packed = ncls <= 8;       // per-lane bucket counters in one 64-bit word (8 bits each)
        unsigned long long hist = 0;
        if (!packed) {
            if (lane <= ncls) ctl->hw[wid][lane] = 0;
            __syncwarp();
        }
        const bool allang = (angmask >> 1) & 1u;   // every bucket >= 1 is wanted (bits above a set bit are set)
        const unsigned seg_sa = (unsigned)__cvta_generic_to_shared(seg);
        int n1 = 0;                                // bucket-1 survivors of the trips that can hold nothing else
        int ra = 0, rb = 1;
        if (wq0 + lane < wq1) tri_decode(wq0 + lane, ra, rb);
        for (int qb = wq0; qb < wq1; qb += 32) {
            int bk = 0, lim = 0;
            double rjk2 = 0.0;
            if (qb + lane < wq1) {
                // sorted staging: the place rb > ra belongs to no more classes than ra
                lim = SORTED ? (int)NCB(rb) : min((int)NCB(ra), (int)NCB(rb));
                const double2 axy = NB2(ra, 0), bxy = NB2(rb, 0);
                const double az = NB2(ra, 1).x, bz = NB2(rb, 1).x;
                rjk2 = pair_dist2(axy.x, axy.y, az, bxy.x, bxy.y, bz);
            }
            // thresholds descend: the classes with rjk2 <= t2[c] are a prefix, and only the first `lim` of them
            // matter.  A trip covers one or two rows rb, whose lim the sorted staging makes (nearly) equal and
            // mostly 1: test as many thresholds as the trip's largest lim asks for
            const int limmax = __reduce_max_sync(0xffffffffu, lim);
            int lo = rjk2 <= a.cls.t2[0];
#pragma unroll 1
            for (int c = 1; c < limmax; c++) lo += rjk2 <= s_t2[c];
            bk = min(lo, lim);
            if (!allang && !((angmask >> bk) & 1u)) bk = 0;
            const unsigned m = __ballot_sync(0xffffffffu, bk > 0);
            // (plain 32-bit shared address: the generic pointer form re-derived the window base on every trip)
            if (bk > 0) st_shared_u32(seg_sa + 4u * (unsigned)(cnt + __popc(m & ltmask)), (uint32_t)ra | ((uint32_t)rb << 10) | ((uint32_t)bk << 20));
            cnt += __popc(m);
            if (limmax <= 1) {
                n1 += __popc(m);                 // only bucket 1 can occur in this trip: no per-lane bookkeeping
            } else if (packed) {
                hist += (unsigned long long)(bk > 0) << (((bk - 1) & 7) * 8);
            } else if (bk > 0) {
                const unsigned peers = __match_any_sync(m, bk);
                if ((peers & ltmask) == 0) ctl->hw[wid][bk] += __popc(peers);  // one lane per bucket present
            }
            // next pair of this lane: q += 32
            ra += 32;
            while (ra >= rb) { ra -= rb; rb++; }
        }
        if (packed) {
            // a lane saw at most R/32 <= 255 pairs: byte fields cannot overflow; widen to 16 bits for the warp sum
            const uint32_t lo32 = (uint32_t)hist, hi32 = (uint32_t)(hist >> 32);
            const uint32_t e0 = __reduce_add_sync(0xffffffffu, lo32 & 0x00ff00ffu);          // buckets 1, 3
            const uint32_t o0 = __reduce_add_sync(0xffffffffu, (lo32 >> 8) & 0x00ff00ffu);   // buckets 2, 4
            const uint32_t e1 = __reduce_add_sync(0xffffffffu, hi32 & 0x00ff00ffu);          // buckets 5, 7
            const uint32_t o1 = __reduce_add_sync(0xffffffffu, (hi32 >> 8) & 0x00ff00ffu);   // buckets 6, 8
            if (lane < 8) {
                const uint32_t w = (lane & 4) ? ((lane & 1) ? o1 : e1) : ((lane & 1) ? o0 : e0);
                ctl->hw[wid][lane + 1] = ((lane & 2) ? (w >> 16) : (w & 0xffffu)) + (lane == 0 ? n1 : 0);
            }
            if (lane == 8) ctl->hw[wid][0] = 0;
        } else if (lane == 0) {
            ctl->hw[wid][1] += n1;
        }
        if (lane == 0) ctl->cntw[wid] = cnt;
        return cnt;
    };
    // Part 2 (ONE warp, after a block barrier): places of the (warp, bucket) runs in the sorted list S
    auto list_scan = [&](bool count_work) {
        // lane o <-> bucket v = ncls - o (heavy buckets first); totals over warps, then an
        // exclusive scan over the lanes gives every bucket its place in S
        const int o = lane, v = ncls - lane;
        int t = 0;
        if (o < ncls)
            for (int w = 0; w < NW; w++) { const int h = ctl->hw[w][v]; ctl->basew[w][v] = t; t += h; }
        const int nb = (t + 31) >> 5;
        int off = t, bp = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int yo = __shfl_up_sync(0xffffffffu, off, d), yb = __shfl_up_sync(0xffffffffu, bp, d);
            if (lane >= d) { off += yo; bp += yb; }
        }
        // off/bp are inclusive sums over order slots 0..o
        if (o < ncls) {
            ctl->tot[v] = t;
            ctl->ob[o] = make_int4(nb, t, off - t, bp - nb);
            for (int w = 0; w < NW; w++) ctl->basew[w][v] += off - t;
            ctl->npre[v - 1] = off;   // items with bucket > v-1, i.e. bucket >= v: slots 0..o
        }
        if (o == ncls - 1 || (ncls == 0 && o == 0)) { ctl->TB = bp; ctl->nkept = off; }
        __syncwarp();
        if (count_work) {
            const int tcv = (lane < ncls && s_nf[lane]) ? ctl->npre[lane] : 0;
            const unsigned tc = __reduce_add_sync(0xffffffffu, (unsigned)tcv);
            const unsigned tsf = __reduce_add_sync(0xffffffffu, (unsigned)(tcv * (lane < ncls ? s_nf[lane] : 0)));
            if (lane == 0) { s_work[4] += (unsigned)ctl->nkept; s_work[5] += tc; s_work[6] += tsf; }
        }
    };
    // Part 3 (all warps, after a block barrier): the warp's survivors move to their bucket's run in S
    auto list_scatter = [&](int q0, int q1, int cnt) {
        const int R = (((q1 - q0 + NW - 1) / NW) + 31) & ~31;
        const uint32_t *seg = s_U + wid * R;
        for (int t0 = 0; t0 < cnt; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < cnt;
            const unsigned am = __ballot_sync(0xffffffffu, act);
            if (act) {
                const uint32_t it = seg[t];
                const int v = it >> 20;
                const unsigned peers = __match_any_sync(am, v);
                const int rank = __popc(peers & ltmask);
                const int base = ctl->basew[wid][v];
                s_S[base + rank] = it;
                __syncwarp(am);
                if (rank == 0) ctl->basew[wid][v] = base + __popc(peers);
            }
            __syncwarp();
        }
    };
    // the three parts in sequence (list re-builds of the backward pass and of further chunks)
    auto build_list = [&](int q0, int q1, bool count_work) {
        const int cnt = list_pairs(q0, q1);
        __syncthreads();
        if (wid == 0) list_scan(count_work);
        __syncthreads();
        list_scatter(q0, q1, cnt);
        __syncthreads();
    };

    // ---- 4: forward over the sorted list ----------------------------------------------
    auto forward_list = [&]() {
        int rot = 0;  // items dealt so far: the next class continues the round robin where this one stopped
        for (int c = 0; c < ncls; c++) {
            const int g0 = GRP_BEGIN(a, c), g1 = GRP_BEGIN(a, c + 1);
            const int n_c = ctl->npre[c];
            if (g0 == g1 || n_c == 0) continue;
            const int myrank = (tid - rot + CT) % CT;   // this thread's position in the deal for class c
            rot = (rot + n_c) % CT;
            const double pirc = a.cls.pirc[c];
            const unsigned char *fcc = smem + H::FCD + c * (PCAP * 16);   // fc of class c: record stride 16 bytes
            // per item: geometry, fc product and the angle; shared by the loops below
            struct Item { double phi, cosv, ssum, ww; };
            auto prep = [&](int t) {
                const uint32_t it = s_S[t];
                const int ra = it & 1023, rb = (it >> 10) & 1023;
                const double2 axy = NB2(ra, 0), azr = NB2(ra, 1), aiw = NB2(ra, 2);
                const double2 bxy = NB2(rb, 0), bzr = NB2(rb, 1), biw = NB2(rb, 2);
                const double rjk2 = dist2_fma(axy.x - bxy.x, axy.y - bxy.y, azr.x - bzr.x);
                const double ra2 = azr.y * azr.y, rb2 = bzr.y * bzr.y;
                Item x;
                x.cosv = (ra2 + rb2 - rjk2) * 0.5 * aiw.x * biw.x;
                x.ssum = ra2 + rb2 + rjk2;
                x.ww = aiw.y * biw.y;
                const double rjk = rjk2 * rsqrt_pos(rjk2);   // two neighbours never coincide (the backward pass divides by rjk as the reference does)
                double sn, cs;
                sincos_tab_s(rjk * pirc, trig_sa, &sn, &cs);
                x.phi = *(const double *)(fcc + ra * 16) * *(const double *)(fcc + rb * 16) * (0.5 * (cs + 1.0));
                return x;
            };
            if (g1 - g0 <= 2) {
                // one or two exponents in this class (the usual case): TWO triplets per thread and trip,
                // i.e. four independent exponential chains in flight -- the kernel is latency bound
                const int ng = g1 - g0;
                const double al0 = s_galpha[g0], al1 = s_galpha[g1 - 1];
                double acc[2][4];
#pragma unroll
                for (int g = 0; g < 2; g++) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.0;
                auto add = [&](const Item &x, double e0, double e1) {
                    const double pe0 = x.phi * e0, pe1 = x.phi * e1, pw0 = pe0 * x.ww, pw1 = pe1 * x.ww;
                    acc[0][0] += pe0; acc[0][1] = fma(pe0, x.cosv, acc[0][1]); acc[0][2] += pw0; acc[0][3] = fma(pw0, x.cosv, acc[0][3]);
                    acc[1][0] += pe1; acc[1][1] = fma(pe1, x.cosv, acc[1][1]); acc[1][2] += pw1; acc[1][3] = fma(pw1, x.cosv, acc[1][3]);
                };
                int t = myrank;
                const bool reuse = se && c != a.c_first;   // this class reads the exponentials the first angular class parked
                for (; t + CT < n_c; t += 2 * CT) {
                    const Item x = prep(t), y = prep(t + CT);
                    double ex0, ey0, ex1, ey1;
                    if (reuse) {
                        const double2 sx = __ldcg(est + t), sy = __ldcg(est + t + CT);
                        ex0 = sx.x; ex1 = sx.y; ey0 = sy.x; ey1 = sy.y;
                    } else {
                        ex0 = exp_neg_s(-al0 * x.ssum, t32_sa); ey0 = exp_neg_s(-al0 * y.ssum, t32_sa);
                        ex1 = exp_neg_s(-al1 * x.ssum, t32_sa); ey1 = exp_neg_s(-al1 * y.ssum, t32_sa);
                        if (se) { __stcg(est + t, make_double2(ex0, ex1)); __stcg(est + t + CT, make_double2(ey0, ey1)); }
                    }
                    add(x, ex0, ex1);
                    add(y, ey0, ey1);
                }
                if (t < n_c) {
                    const Item x = prep(t);
                    double ex0, ex1;
                    if (reuse) { const double2 sx = __ldcg(est + t); ex0 = sx.x; ex1 = sx.y; }
                    else {
                        ex0 = exp_neg_s(-al0 * x.ssum, t32_sa); ex1 = exp_neg_s(-al1 * x.ssum, t32_sa);
                        if (se) __stcg(est + t, make_double2(ex0, ex1));
                    }
                    add(x, ex0, ex1);
                }
                if (__any_sync(0xffffffffu, myrank < n_c)) {
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        if (g < ng) {   // with a single exponent acc[1] duplicates acc[0] and is dropped
                            const double tot = warp_sum4(acc[g][0], acc[g][1], acc[g][2], acc[g][3], lane);
                            const double oth = __shfl_xor_sync(0xffffffffu, tot, 8);
                            if ((lane & 7) == 0) {
                                int ip, im;
                                if (gidx_ok) { const int2 gi = s_gidx[g0 + g]; ip = gi.x; im = gi.y; }
                                else { ip = g_iplus[g0 + g]; im = g_iminus[g0 + g]; }
                                double *gw = s_gw + wid * D + ((lane & 16) ? nsf : 0);
                                if (!(lane & 8)) { if (ip >= 0) gw[ip] += tot + oth; }
                                else { if (im >= 0) gw[im] += oth - tot; }
                            }
                        }
                    }
                }
                if (se && c == a.c_first) __syncthreads();   // the parked values are read by other threads from here on
                continue;
            }
            // (the parked-exponentials variant is only dispatched for <= 2 exponents per class: dead code there)
            for (int gb = g0; !SE && gb < g1; gb += MAXG) {
                const int ng = min(MAXG, g1 - gb);
                double acc[MAXG][4];
#pragma unroll
                for (int g = 0; g < MAXG; g++) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.0;
                for (int t = myrank; t < n_c; t += CT) {
                    const Item x = prep(t);
                    const double phi = x.phi, cosv = x.cosv, ssum = x.ssum, ww = x.ww;
                    // groups in pairs without a branch between their exponentials, so the two
                    // dependent chains interleave (same trick as in the backward loop)
                    auto one = [&](double (&ac)[4], int g) {
                        const double pe = phi * exp_neg_s(-s_galpha[gb + g] * ssum, t32_sa);
                        const double pw = pe * ww;
                        ac[0] += pe; ac[1] = fma(pe, cosv, ac[1]); ac[2] += pw; ac[3] = fma(pw, cosv, ac[3]);
                    };
                    auto two = [&](double (&a0)[4], double (&a1)[4], int g) {
                        const double e0 = exp_neg_s(-s_galpha[gb + g] * ssum, t32_sa);
                        const double e1 = exp_neg_s(-s_galpha[gb + g + 1] * ssum, t32_sa);
                        const double pe0 = phi * e0, pe1 = phi * e1, pw0 = pe0 * ww, pw1 = pe1 * ww;
                        a0[0] += pe0; a0[1] = fma(pe0, cosv, a0[1]); a0[2] += pw0; a0[3] = fma(pw0, cosv, a0[3]);
                        a1[0] += pe1; a1[1] = fma(pe1, cosv, a1[1]); a1[2] += pw1; a1[3] = fma(pw1, cosv, a1[3]);
                    };
                    if (ng >= 2) two(acc[0], acc[1], 0); else one(acc[0], 0);
                    if (ng >= 4) two(acc[2], acc[3], 2); else if (ng == 3) one(acc[2], 2);
                }
                if (__any_sync(0xffffffffu, myrank < n_c)) {
#pragma unroll
                    for (int g = 0; g < MAXG; g++) {
                        if (g < ng) {
                            // lanes 0,8,16,24 end up with sum pe, sum pe*cos, sum w*pe, sum w*pe*cos
                            const double tot = warp_sum4(acc[g][0], acc[g][1], acc[g][2], acc[g][3], lane);
                            const double oth = __shfl_xor_sync(0xffffffffu, tot, 8);  // the cos partner / the plain partner
                            if ((lane & 7) == 0) {
                                const int ip = g_iplus[gb + g], im = g_iminus[gb + g];
                                double *gw = s_gw + wid * D + ((lane & 16) ? nsf : 0);
                                // lane&8 == 0: tot = plain sum, oth = cos sum -> lambda=+1 ; else the lambda=-1 one
                                if (!(lane & 8)) { if (ip >= 0) gw[ip] += tot + oth; }
                                else { if (im >= 0) gw[im] += oth - tot; }
                            }
                        }
                    }
                }
            }
        }
    };

    // ---- 6: backward over the sorted list -----------------------------------------------
    auto backward_list = [&]() {
        const bool priv = a.npa > 1;
        double *pa = priv ? s_pa + wid * (3 * PCAP) : s_acc;
        const unsigned pa_sa = (unsigned)__cvta_generic_to_shared(s_pa + wid * (3 * PCAP));   // private set: (x, y)[PCAP] then z[PCAP]
        if (priv) {
            for (int t = lane; t < 3 * PCAP; t += 32) pa[t] = 0.0;
            __syncwarp();
        }
        // Batches of 32 triplets of one bucket (uniform class loop), heavy buckets first; batch number
        // obp[o] + q goes to warp (obp[o] + q) mod NW: a fixed round robin over all buckets, so every warp
        // gets its share of the expensive batches and the summation order is reproducible.  Everything
        // that depends on the bucket only is fetched once per bucket.
        // g2: every angular class carries the same two exponents (the shipped table) -- the exponentials of
        // a triplet are then evaluated (or fetched, MODE_FUSED_SE) once, not once per class, and the sum over
        // the groups is straight-line code.
        const bool g2 = a.share_exp == 2;
        const double al0 = s_galpha[GRP_BEGIN(a, a.c_first)], al1 = s_galpha[GRP_BEGIN(a, a.c_first) + (g2 ? 1 : 0)];
        const unsigned ob_sa = (unsigned)__cvta_generic_to_shared(&ctl->ob[0]);
        for (int o = 0; o < ncls; o++) {
            for (int qr = 0;; qr += NW) {
            // the bucket's constants are re-read from shared memory per batch (one 128-bit load through a plain
            // shared address): carried in registers across the batch they were spilled to local memory, whose
            // loads miss the small L1 this kernel leaves
            int Qb, n, base, p0;
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(Qb), "=r"(n), "=r"(base), "=r"(p0) : "r"(ob_sa + 16u * (unsigned)o) : "memory");
            if (qr >= Qb) break;
            const int v = ncls - o;
            const int q = qr + ((wid - p0) & (NW - 1));
            // shared accumulator set (very long lists, no room for private sets): the warps of a round add
            // their batches one after the other in warp order, so block barriers sit inside this loop and
            // every warp has to reach them -- a warp without a batch only skips the arithmetic
            bool act = false;
            unsigned am = 0;
            int ra = 0, rb = 0;
            double v0 = 0, v1 = 0, v2 = 0, w0 = 0, w1 = 0, w2 = 0;
            if (q < Qb) {
            __syncwarp();   // orders this batch's accumulator reads after the previous batch's writes for lanes that sat that one out
            const int idx = lane * Qb + q;  // lanes far apart in the list -> mostly distinct rows
            act = idx < n;
            am = __ballot_sync(0xffffffffu, act);
            if (act) {
            const int li = base + idx;
            const uint32_t it = s_S[li];
            double2 pk = make_double2(0.0, 0.0);
            if (se) pk = __ldcg(a.estash + (size_t)(ctl->estb + li));   // requested before the geometry below needs it; the slice's
                                                                       // start comes from shared memory (as a register it was spilled)
            ra = it & 1023; rb = (it >> 10) & 1023;
            const double2 axy = NB2(ra, 0), azr = NB2(ra, 1), aiw = NB2(ra, 2);
            const double2 bxy = NB2(rb, 0), bzr = NB2(rb, 1), biw = NB2(rb, 2);
            const double rja = azr.y, rkb = bzr.y, ira = aiw.x, irb = biw.x;
            const double rjk2 = dist2_fma(axy.x - bxy.x, axy.y - bxy.y, azr.x - bzr.x);
            const double irjk = rsqrt_pos(rjk2);
            const double rjk = rjk2 * irjk;
            const double ra2 = rja * rja, rb2 = rkb * rkb;
            const double cosv = (ra2 + rb2 - rjk2) * 0.5 * ira * irb;
            const double ssum = ra2 + rb2 + rjk2;
            const double ww = aiw.y * biw.y;
            const double u1 = irb - cosv * ira, u2 = ira - cosv * irb, u3 = -rjk * ira * irb;
            double cij = 0.0, cik = 0.0, cjk = 0.0;
            if (g2) {
                const double e0 = se ? pk.x : exp_neg_s(-al0 * ssum, t32_sa), e1 = se ? pk.y : exp_neg_s(-al1 * ssum, t32_sa);
                const double e0w = e0 * ww, e1w = e1 * ww;
#pragma unroll 2
                for (int c = 0; c < v; c++) {
                    const int g0 = GRP_BEGIN(a, c);
                    if (g0 == GRP_BEGIN(a, c + 1)) continue;
                    const double pirc = a.cls.pirc[c];
                    double sn, cs;
                    sincos_tab_s(rjk * pirc, trig_sa, &sn, &cs);
                    const double fjk = 0.5 * (cs + 1.0), dfjk = -0.5 * pirc * sn;
                    const double2 fda = FCD2(c, ra), fdb = FCD2(c, rb);
                    const double fa = fda.x, fb = fdb.x, dfa = fda.y, dfb = fdb.y;
                    const double fab = fa * fb, phi = fab * fjk;
                    const double4 gd0 = *(const double4 *)(s_gd + 4 * g0), gd1 = *(const double4 *)(s_gd + 4 * g0 + 4);   // DU, DW, DUL, DWL
                    const double t00 = fma(e0w, gd0.y, e0 * gd0.x), t01 = fma(e0w, gd0.w, e0 * gd0.z);
                    const double t10 = fma(e1w, gd1.y, e1 * gd1.x), t11 = fma(e1w, gd1.w, e1 * gd1.z);
                    const double T0 = t00 + t10, S1 = t01 + t11;
                    const double S3 = fma(al0, fma(cosv, t01, t00), al1 * fma(cosv, t11, t10));
                    const double S2 = fma(cosv, S1, T0);
                    const double pS1 = phi * S1, pS3 = 2.0 * phi * S3;
                    cij += pS1 * u1 - pS3 * rja + S2 * (dfa * fb * fjk);
                    cik += pS1 * u2 - pS3 * rkb + S2 * (fa * dfb * fjk);
                    cjk += pS1 * u3 - pS3 * rjk + S2 * (fab * dfjk);
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < v; c++) {
                const int g0 = GRP_BEGIN(a, c), g1 = GRP_BEGIN(a, c + 1);
                if (g0 == g1) continue;
                const double pirc = a.cls.pirc[c];
                double sn, cs;
                sincos_tab_s(rjk * pirc, trig_sa, &sn, &cs);
                const double fjk = 0.5 * (cs + 1.0), dfjk = -0.5 * pirc * sn;
                const double2 fda = FCD2(c, ra), fdb = FCD2(c, rb);
                const double fa = fda.x, fb = fdb.x, dfa = fda.y, dfb = fdb.y;
                const double fab = fa * fb, phi = fab * fjk;
                // S1 = sum gamma e lam, T0 = sum gamma e, S3 = sum alpha gamma e (1 + lam cos)
                double S1 = 0.0, T0 = 0.0, S3 = 0.0;
                // two groups per trip: their exponentials are independent chains (ILP; measured -7 %)
                int gg = g0;
                for (; gg + 1 < g1; gg += 2) {
                    const double bl0 = s_galpha[gg], bl1 = s_galpha[gg + 1];
                    const double e0 = se ? pk.x : exp_neg_s(-bl0 * ssum, t32_sa), e1 = se ? pk.y : exp_neg_s(-bl1 * ssum, t32_sa);
                    const double4 gd0 = *(const double4 *)(s_gd + 4 * gg), gd1 = *(const double4 *)(s_gd + 4 * gg + 4);
                    const double t00 = e0 * fma(ww, gd0.y, gd0.x), t01 = e0 * fma(ww, gd0.w, gd0.z);
                    const double t10 = e1 * fma(ww, gd1.y, gd1.x), t11 = e1 * fma(ww, gd1.w, gd1.z);
                    T0 += t00 + t10;
                    S1 += t01 + t11;
                    S3 = fma(bl0, fma(cosv, t01, t00), fma(bl1, fma(cosv, t11, t10), S3));
                }
                if (gg < g1) {
                    const double al = s_galpha[gg];
                    const double e = se ? pk.x : exp_neg_s(-al * ssum, t32_sa);
                    const double4 gd = *(const double4 *)(s_gd + 4 * gg);   // DU, DW, DUL, DWL
                    const double t0 = e * fma(ww, gd.y, gd.x), t1 = e * fma(ww, gd.w, gd.z);
                    T0 += t0;
                    S1 += t1;
                    S3 = fma(al, fma(cosv, t1, t0), S3);
                }
                const double S2 = fma(cosv, S1, T0);  // sum gamma e (1 + lam cos)
                const double pS1 = phi * S1, pS3 = 2.0 * phi * S3;
                cij += pS1 * u1 - pS3 * rja + S2 * (dfa * fb * fjk);
                cik += pS1 * u2 - pS3 * rkb + S2 * (fa * dfb * fjk);
                cjk += pS1 * u3 - pS3 * rjk + S2 * (fab * dfjk);
            }
            }
            // dE/dx_j = (gij+gjk) d_j - gjk d_k ; dE/dx_k = (gik+gjk) d_k - gjk d_j
            const double gij = cij * ira, gik = cik * irb, gjk = cjk * irjk;
            // coordinates again from shared memory: holding them across the class loop cost spills to local memory
            asm volatile("" ::: "memory");
            const double2 axy2 = NB2(ra, 0), bxy2 = NB2(rb, 0);
            const double az2 = NB2(ra, 1).x, bz2 = NB2(rb, 1).x;
            const double cx = s_ctr[0], cy = s_ctr[1], cz = s_ctr[2];
            const double dxa = axy2.x - cx, dya = axy2.y - cy, dza = az2 - cz, dxb = bxy2.x - cx, dyb = bxy2.y - cy, dzb = bz2 - cz;
            const double ga = gij + gjk, gb = gik + gjk;
            v0 = fma(ga, dxa, -gjk * dxb); v1 = fma(ga, dya, -gjk * dyb); v2 = fma(ga, dza, -gjk * dzb);
            w0 = fma(gb, dxb, -gjk * dxa); w1 = fma(gb, dyb, -gjk * dya); w2 = fma(gb, dzb, -gjk * dza);
            }
            }
            if (priv) {
                if (act) {
                    scatter3p<PCAP>(pa_sa, ra, v0, v1, v2, am, ltmask);
                    scatter3p<PCAP>(pa_sa, rb, w0, w1, w2, am, ltmask);
                }
            } else {
                for (int w = 0; w < NW; w++) {
                    if (w == wid && act) {
                        scatter3<PCAP>(pa, ra, v0, v1, v2, am, ltmask);
                        scatter3<PCAP>(pa, rb, w0, w1, w2, am, ltmask);
                    }
                    __syncthreads();
                }
            }
            }
        }
        __syncthreads();
        phase_end(14);
        if (priv) {
            for (int t = tid; t < PCAP; t += CT) {   // private sets: (x, y)[PCAP] then z[PCAP]; s_acc: [3][PCAP]
                double x = 0.0, y = 0.0, z = 0.0;
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const double2 q = ((const double2 *)(s_pa + w * (3 * PCAP)))[t];
                    x += q.x; y += q.y; z += s_pa[w * (3 * PCAP) + 2 * PCAP + t];
                }
                s_acc[t] += x; s_acc[PCAP + t] += y; s_acc[2 * PCAP + t] += z;
            }
            __syncthreads();
        }
    };

    // ---- drive the phases ---------------------------------------------------------------
    bool list_ready = false, list_stashed = false;
    if (!FWD || nchunk == 0) {
        fc_tables(tid, CT);
        __syncthreads();
    }
    phase_end(0);
    if (FWD) {
        const bool stash = FUSED && nchunk > 1 && a.list_scratch && nchunk <= a.list_scratch_chunks;
        int trip_base = 0;
        if (nchunk == 0) {
            radial_forward();
            if (lead && wid == 0) count_centre();
            __syncthreads();
        }
        phase_end(1);
        for (int ch = 0; ch < nchunk; ch++) {
            if (SE) {
                est = a.estash + (size_t)blockIdx.x * a.estash_stride + (size_t)ch * (lcap + 32);
                if (tid == 0) ctl->estb = (int)(blockIdx.x * a.estash_stride + ch * (lcap + 32));   // read after the barriers below
            }
            const int q0 = Qlo + ch * qchunk, q1 = min(Qhi, Qlo + (ch + 1) * qchunk);
            // The pair tests need the staged records only.  While ONE warp then places the buckets (a short
            // serial step), the other seven fill the fc tables; the list scatter and the radial functions
            // (which need those tables) share the next slot: two block barriers and a serial section less
            // than stage -> tables -> radial -> list.
            const int cnt = list_pairs(q0, q1);
            __syncthreads();
            phase_end(8);
            if (wid == 0) {
                list_scan(true);
                if (ch == 0 && lead) count_centre();
            } else if (ch == 0) {
                fc_tables(tid - 32, CT - 32);
            }
            __syncthreads();
            phase_end(9);
            list_scatter(q0, q1, cnt);
            if (ch == 0) radial_forward();
            __syncthreads();
            phase_end(10);
            if (a.trip_out) {
                // debug export (gapcu_ctx_debug_triplets): the kept pairs exactly as the passes below consume them
                const int n = ctl->nkept;
                for (int t = tid; t < n; t += CT)
                    if (trip_base + t < a.trip_cap) {
                        uint32_t it = s_S[t];
                        if constexpr (SORTED) {   // places -> list slots, smaller slot first
                            const uint32_t sa = PERM(it & 1023), sb = PERM((it >> 10) & 1023);
                            it = min(sa, sb) | (max(sa, sb) << 10) | (it & ~0xfffffu);
                        }
                        a.trip_out[(size_t)i * a.trip_cap + trip_base + t] = it;
                    }
                trip_base += n;
            }
            phase_end(2);
            if (stash) {
                // keep the sorted list of this chunk (L2 resident) for the backward pass
                uint32_t *dst = a.list_scratch + ((size_t)blockIdx.x * a.list_scratch_chunks + ch) * (size_t)(lcap + 32 + 512);
                const int n = ctl->nkept;
                for (int t = tid; t < n; t += CT) dst[t] = s_S[t];
                const int *cs = (const int *)ctl;
                for (int t = tid; t < (int)(sizeof(Ctl) / 4); t += CT) dst[lcap + 32 + t] = (uint32_t)cs[t];
            }
            forward_list();
            __syncthreads();
            phase_end(3);
        }
        if (a.trip_out && tid == 0) a.trip_cnt[i] = trip_base;
        list_ready = (nchunk == 1);
        list_stashed = stash;
        __syncthreads();
        // per-warp partial sums -> descriptors (fixed order)
        for (int k = tid; k < D; k += CT) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) v += s_gw[w * D + k];
            if constexpr (CS > 1) {
                s_gx[k] = v;
            } else {
                if (FUSED) s_G[k] = v;
                if (a.G) a.G[(size_t)i * D + k] = v;
            }
        }
        if constexpr (CS > 1) {
            // partial descriptors of the cluster's CTAs, added in rank order by every CTA
            cg::cluster_group cl = cg::this_cluster();
            cl.sync();
            for (int k = tid; k < D; k += CT) {
                double v = 0.0;
#pragma unroll
                for (int r = 0; r < CS; r++) v += cl.map_shared_rank(s_gx, r)[k];
                if (FUSED) s_G[k] = v;
                if (lead && a.G) a.G[(size_t)i * D + k] = v;
            }
        }
    }
    if (FUSED) {
        // ---- 5: sparse GPR of this atom (gap_calc.f90:143-166), difference form ----------
        __syncthreads();
        phase_end(11);
        const int M = a.gpr_M, Mp = a.gpr_Mp, Dp = a.gpr_Dp;
        for (int k = tid; k < D; k += CT) s_xs[k] = (s_G[k] - a.gpr_cmean[k]) * a.gpr_itheta[k];
        __syncthreads();
        // (the sparse set is read past L1 -- ld.global.cg: 2 x 72 KB per centre stream through an L1 of ~30 KB
        // that the spilled registers and the neighbour records of the co-resident CTAs live in)
        // squared distances: thread (sparse point j, slab h of the descriptor components); as many
        // slabs as the CTA has threads for, so that a thread runs a long loop instead of a short one
        double *part = (double *)(smem + L.scratch);          // [slabs <= NW][Mp] partial sums (gw/U are dead by now)
        const int nsA = max(1, min(CT / Mp, NW));
        {
            const int kslab = (D + nsA - 1) / nsA;
            for (int item = tid; item < Mp * nsA; item += CT) {
                const int h = item / Mp, j = item - h * Mp;
                const int k0 = h * kslab, k1 = min(D, k0 + kslab);
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                const double *col = a.gpr_MtT + (size_t)k0 * Mp + j;
                int k = k0;
                for (; k + 15 < k1; k += 16, col += 16 * (size_t)Mp) {   // sixteen loads in flight: this loop waits on L2, not on arithmetic
                    double m[16];
                    // (uniform base + 32-bit byte offset: one add per load instead of 64-bit index arithmetic)
                    const unsigned o0 = (unsigned)((const char *)col - (const char *)a.gpr_MtT), st = 8u * (unsigned)Mp;
#pragma unroll
                    for (int u = 0; u < 16; u++) m[u] = __ldcg((const double *)((const char *)a.gpr_MtT + (o0 + u * st)));
#pragma unroll
                    for (int u = 0; u < 16; u += 4) {
                        const double d0 = s_xs[k + u] - m[u], d1 = s_xs[k + u + 1] - m[u + 1], d2 = s_xs[k + u + 2] - m[u + 2], d3 = s_xs[k + u + 3] - m[u + 3];
                        s0 = fma(d0, d0, s0); s1 = fma(d1, d1, s1); s2 = fma(d2, d2, s2); s3 = fma(d3, d3, s3);
                    }
                }
                for (; k + 3 < k1; k += 4, col += 4 * (size_t)Mp) {   // four independent chains, loads issued together
                    const double m0 = __ldcg(col), m1 = __ldcg(col + Mp), m2 = __ldcg(col + 2 * (size_t)Mp), m3 = __ldcg(col + 3 * (size_t)Mp);
                    const double d0 = s_xs[k] - m0, d1 = s_xs[k + 1] - m1, d2 = s_xs[k + 2] - m2, d3 = s_xs[k + 3] - m3;
                    s0 = fma(d0, d0, s0); s1 = fma(d1, d1, s1); s2 = fma(d2, d2, s2); s3 = fma(d3, d3, s3);
                }
                for (; k < k1; k++, col += Mp) { const double d0 = s_xs[k] - __ldcg(col); s0 = fma(d0, d0, s0); }
                part[h * Mp + j] = (s0 + s1) + (s2 + s3);
            }
        }
        __syncthreads();
        phase_end(12);
        double esum = 0.0;
        for (int j = tid; j < Mp; j += CT) {
            double sacc = 0.0;
            for (int h = 0; h < nsA; h++) sacc += part[h * Mp + j];
            const double wv = (j < M) ? exp_neg_s(-0.5 * sacc, t32_sa) * a.gpr_coeff[j] : 0.0;
            s_W[j] = wv;
            esum += wv;
        }
        esum = warp_sum(esum);
        if (lane == 0) s_red[wid] = esum;
        __syncthreads();
        if (tid == 0) {
            double e = 0.0;
            for (int w = 0; w < NW; w++) e += s_red[w];
            if (lead) a.eatom[i] = e;
        }
        // dE/dG_k = -(1/theta_k) sum_j W_j (x'_k - m'_jk): thread (component k, slab h of the sparse points)
        const int nsB = max(1, min(CT / D, NW));
        {
            const int jslab = (M + nsB - 1) / nsB;
            for (int item = tid; item < D * nsB; item += CT) {
                const int h = item / D, k = item - h * D;
                const int j0 = h * jslab, j1 = min(M, j0 + jslab);
                const double xk = s_xs[k];
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                const double *rp = a.gpr_Mt + k + (size_t)j0 * Dp;
                int j = j0;
                for (; j + 15 < j1; j += 16, rp += 16 * (size_t)Dp) {
                    double m[16];
                    const unsigned o0 = (unsigned)((const char *)rp - (const char *)a.gpr_Mt), st = 8u * (unsigned)Dp;
#pragma unroll
                    for (int u = 0; u < 16; u++) m[u] = __ldcg((const double *)((const char *)a.gpr_Mt + (o0 + u * st)));
#pragma unroll
                    for (int u = 0; u < 16; u += 4) {
                        a0 = fma(s_W[j + u], xk - m[u], a0); a1 = fma(s_W[j + u + 1], xk - m[u + 1], a1);
                        a2 = fma(s_W[j + u + 2], xk - m[u + 2], a2); a3 = fma(s_W[j + u + 3], xk - m[u + 3], a3);
                    }
                }
                for (; j + 3 < j1; j += 4, rp += 4 * (size_t)Dp) {
                    const double m0 = __ldcg(rp), m1 = __ldcg(rp + Dp), m2 = __ldcg(rp + 2 * (size_t)Dp), m3 = __ldcg(rp + 3 * (size_t)Dp);
                    a0 = fma(s_W[j], xk - m0, a0); a1 = fma(s_W[j + 1], xk - m1, a1);
                    a2 = fma(s_W[j + 2], xk - m2, a2); a3 = fma(s_W[j + 3], xk - m3, a3);
                }
                for (; j < j1; j++, rp += Dp) a0 = fma(s_W[j], xk - __ldcg(rp), a0);
                part[h * D + k] = (a0 + a1) + (a2 + a3);   // part is reused with stride D
            }
        }
        __syncthreads();
        for (int k2 = tid; k2 < D; k2 += CT) {
            double acc = 0.0;
            for (int h = 0; h < nsB; h++) acc += part[h * D + k2];
            const double v = -a.gpr_itheta[k2] * acc;
            s_du[k2] = v;
            if (lead && a.dEdG_out) a.dEdG_out[(size_t)i * D + k2] = v;
        }
        __syncthreads();
    }
    phase_end(4);
    if (BWD) {
        if (!a.lgrad) return;
        // per (class, alpha) group: sums of dE/dG over its functions
        for (int g = tid; g < pl.n_grp; g += CT) {
            const int ip = g_iplus[g], im = g_iminus[g];
            const double dup = ip >= 0 ? s_du[ip] : 0.0, dwp = ip >= 0 ? s_du[ip + nsf] : 0.0;
            const double dum = im >= 0 ? s_du[im] : 0.0, dwm = im >= 0 ? s_du[im + nsf] : 0.0;
            s_gd[4 * g] = dup + dum; s_gd[4 * g + 1] = dwp + dwm; s_gd[4 * g + 2] = dup - dum; s_gd[4 * g + 3] = dwp - dwm;
        }
        // radial backward: thread (neighbour, part); part p takes the functions q = p, p+nparts, ...
        {
            const int nparts = P32 <= CT ? min(CT / P32, 8) : 1;
            const int stride = P32 <= CT ? P32 : CT;       // neighbours covered per sweep
            double *part = s_pa;  // [nparts][stride] scratch, <= CT doubles (free until backward_list)
            const int p = tid / stride, s0 = tid - p * stride;
            for (int sb = 0; sb < P32; sb += stride) {
                const int s = sb + s0;
                double cacc = 0.0;
                if (p < nparts && s < P) {
                    const double dis = NB2(s, 1).y, wj = NB2(s, 2).y;
                    const int nc = NCB(s);
                    for (int q = p + nparts * crank; q < pl.n_rad; q += nparts * CS) {
                        const int2 ri = s_radi[q];
                        const int c = ri.y & 0xffff;
                        if (c >= nc) continue;
                        double arg, dgf;
                        if ((ri.y >> 16) == 1) { const double al = s_radp[q]; arg = -al * dis * dis; dgf = -2.0 * al * dis; }
                        else { const double d = dis - s_radp[q]; arg = -4.0 * d * d; dgf = -8.0 * d; }
                        const double2 fd = FCD2(c, s);
                        const double dg = exp_neg_s(arg, t32_sa) * fma(dgf, fd.x, fd.y);
                        cacc = fma(s_du[ri.x] + wj * s_du[ri.x + nsf], dg, cacc);
                    }
                }
                if (p < nparts) part[p * stride + s0] = cacc;
                __syncthreads();
                if (tid < stride && sb + tid < PCAP) {
                    double v = 0.0;
                    for (int pp = 0; pp < nparts; pp++) v += part[pp * stride + tid];
                    const int s2 = sb + tid;
                    const bool in = s2 < P;
                    const double g = in ? v * NB2(s2, 2).x : 0.0;     // (dE/dr) / r
                    s_acc[s2] = in ? g * (NB2(s2, 0).x - s_ctr[0]) : 0.0;
                    s_acc[PCAP + s2] = in ? g * (NB2(s2, 0).y - s_ctr[1]) : 0.0;
                    s_acc[2 * PCAP + s2] = in ? g * (NB2(s2, 1).x - s_ctr[2]) : 0.0;
                }
                __syncthreads();
            }
            for (int s2 = P32 + tid; s2 < PCAP; s2 += CT) {
                s_acc[s2] = 0.0; s_acc[PCAP + s2] = 0.0; s_acc[2 * PCAP + s2] = 0.0;
            }
        }
        __syncthreads();
        phase_end(5);
        for (int ch = 0; ch < nchunk; ch++) {
            if (SE && nchunk > 1) {   // (one chunk: still set from the forward pass; several: barriers follow in both branches below)
                est = a.estash + (size_t)blockIdx.x * a.estash_stride + (size_t)ch * (lcap + 32);
                if (tid == 0) ctl->estb = (int)(blockIdx.x * a.estash_stride + ch * (lcap + 32));
            }
            if (list_stashed) {
                const uint32_t *src = a.list_scratch + ((size_t)blockIdx.x * a.list_scratch_chunks + ch) * (size_t)(lcap + 32 + 512);
                int *cs = (int *)ctl;
                for (int t = tid; t < (int)(sizeof(Ctl) / 4); t += CT) cs[t] = (int)src[lcap + 32 + t];
                __syncthreads();
                const int n = ctl->nkept;
                for (int t = tid; t < n; t += CT) s_S[t] = src[t];
                __syncthreads();
            } else if (!list_ready) {
                build_list(Qlo + ch * qchunk, min(Qhi, Qlo + (ch + 1) * qchunk), false);
            }
            backward_list();
        }
        phase_end(6);
        // ---- epilogue: per neighbour gradient, centre gradient, strs contraction ----------
        if constexpr (CS > 1) {
            cg::this_cluster().sync();   // every CTA's accumulator is final; the lead adds them in rank order
            if (!lead) return;
        }
        double acc9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // gself xyz, vir xx xy xz yy yz zz
        for (int s = tid; s < P; s += CT) {
            const double dx = NB2(s, 0).x - s_ctr[0], dy = NB2(s, 0).y - s_ctr[1], dz = NB2(s, 1).x - s_ctr[2];
            double gx = s_acc[s], gy = s_acc[PCAP + s], gz = s_acc[2 * PCAP + s];
            if constexpr (CS > 1) {
                cg::cluster_group cl = cg::this_cluster();
#pragma unroll
                for (int r = 1; r < CS; r++) {
                    const double *ra = cl.map_shared_rank(s_acc, r);
                    gx += ra[s]; gy += ra[PCAP + s]; gz += ra[2 * PCAP + s];
                }
            }
            double *fp = a.fpair + ((size_t)i * a.cap + (SORTED ? (int)PERM(s) : s)) * 3;
            fp[0] = gx; fp[1] = gy; fp[2] = gz;
            acc9[0] -= gx; acc9[1] -= gy; acc9[2] -= gz;
            acc9[3] += dx * gx; acc9[4] += dx * gy; acc9[5] += dx * gz;
            acc9[6] += dy * gy; acc9[7] += dy * gz; acc9[8] += dz * gz;
        }
#pragma unroll
        for (int q = 0; q < 9; q++) {
            const double v = warp_sum(acc9[q]);
            if (lane == 0) s_red[wid * 16 + q] = v;
        }
        __syncthreads();
        if (tid < 9) {
            double v = 0.0;
            for (int w = 0; w < NW; w++) v += s_red[w * 16 + tid];
            if (tid < 3) a.gself[(size_t)i * 3 + tid] = v;
            else a.vir[(size_t)i * 6 + (tid - 3)] = v;
        }
        phase_end(7);
    }
#undef NB2
#undef FCD2
#undef FCV
#undef NCB
#undef PERM
}

// Persistent CTAs: as many as fit the device, each pulling centre atoms from a queue
// ordered by descending neighbour count (longest first), so that 1000 centres on 444
// resident CTAs do not cost full waves and the heavy centres do not form the tail.
template <int MODE, int PCAP, int CS>
__global__ void __launch_bounds__(CT, PCAP <= 128 ? 3 : PCAP <= 256 ? 2 : 1) k_centre(const CentreArgs a) {   // residency the tier's shared memory allows: the larger tiers need not squeeze into 80 registers
    __shared__ int s_next;
    if (threadIdx.x < 10) s_work[threadIdx.x] = 0;
    bool first = true;
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    const int q_begin = a.q_begin ? *a.q_begin : 0;
    const int q_end = a.q_end ? *a.q_end : (a.n_centres ? *a.n_centres : a.ntot);
    for (;;) {
        // everybody (CS > 1: every CTA of the cluster, whose shared memory the lead reads) is done
        // with the previous centre; the lead then pulls the next one for the whole cluster
        if constexpr (CS > 1) {
            cg::cluster_group cl = cg::this_cluster();
            cl.sync();
            if (cl.block_rank() == 0 && threadIdx.x == 0) {
                const int nx = q_begin + atomicAdd(&a.flags->queue[a.queue_slot], 1);
#pragma unroll
                for (int r = 0; r < CS; r++) *cl.map_shared_rank(&s_next, r) = nx;
            }
            cl.sync();
        } else {
            __syncthreads();
            if (threadIdx.x == 0) s_next = q_begin + atomicAdd(&a.flags->queue[a.queue_slot], 1);
            __syncthreads();
        }
        const int n = s_next;
        if (n >= q_end) break;
        process_centre<MODE, PCAP, CS>(a, a.order ? a.order[n] : n, first);
        first = false;
    }
    __syncthreads();
    if (threadIdx.x < 10 && s_work[threadIdx.x]) atomicAdd(&a.flags->work[threadIdx.x], s_work[threadIdx.x]);
    if (threadIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        atomicMax(&a.flags->t_start_min, ~t0);   // flags start zeroed: minima are kept as maxima of the complement
        atomicMax(&a.flags->t_exit_min, ~t1);
        atomicMax(&a.flags->t_exit_max, t1);
        atomicAdd(&a.flags->t_busy_sum, t1 - t0);
        atomicAdd(&a.flags->n_ctas, 1ull);
    }
}

// launch k_centre<MODE, PCAP> on the persistent grid; a.lay must be the layout for (MODE, PCAP)
template <int MODE, int PCAP, int CS>
int launch_centre(cudaStream_t st, const CentreArgs &a) {
    const size_t sm = (size_t)a.lay.total;
    if (sm > 227 * 1024) return -1;
    // attribute and occupancy queries cost microseconds each and sit on the critical path of a small
    // call (the kernels before this one are short): remembered per (instance, device, footprint) and per
    // host thread -- contexts driven from different threads (one per GPU) never share or race on it
    static thread_local int c_dev = -1, c_sms = 0, c_fit = 0;
    static thread_local size_t c_sm = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != c_dev || sm != c_sm) {
        if (cudaFuncSetAttribute((const void *)k_centre<MODE, PCAP, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
            return -2;
        cudaDeviceGetAttribute(&c_sms, cudaDevAttrMultiProcessorCount, dev);
        c_fit = 0;
        if (CS == 1) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c_fit, k_centre<MODE, PCAP, CS>, CT, sm) != cudaSuccess || c_fit < 1) c_fit = 1;
            c_fit *= c_sms;   // resident CTAs of the device
        } else {
            cudaLaunchConfig_t q;
            memset(&q, 0, sizeof q);
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = CS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.gridDim = dim3(CS * c_sms, 1, 1); q.blockDim = dim3(CT, 1, 1); q.dynamicSmemBytes = sm; q.stream = st;
            q.attrs = qa; q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&c_fit, k_centre<MODE, PCAP, CS>, &q) != cudaSuccess || c_fit < 1) return -3;   // resident clusters
        }
        c_dev = dev; c_sm = sm;
    }
    if (CS == 1) {
        const int nmax = a.ncentres_max > 0 ? a.ncentres_max : a.ntot;
        const int grid = nmax < c_fit ? nmax : c_fit;
        k_centre<MODE, PCAP, CS><<<grid, CT, sm, st>>>(a);
        return 0;
    }
    // CS CTAs per centre: a persistent grid of whole clusters
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    const int nmax = a.ncentres_max > 0 ? a.ncentres_max : a.ntot;
    const int clusters = nmax < c_fit ? nmax : c_fit;
    cfg.gridDim = dim3(CS * clusters, 1, 1); cfg.blockDim = dim3(CT, 1, 1); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_centre<MODE, PCAP, CS>, a) == cudaSuccess ? 0 : -4;
}

// one translation unit per capacity (centre_p*.cu) defines these
int launch_centre_p128(cudaStream_t st, const CentreArgs &a, int mode);
int launch_centre_p256(cudaStream_t st, const CentreArgs &a, int mode);
int launch_centre_p512(cudaStream_t st, const CentreArgs &a, int mode);
int launch_centre_p1024(cudaStream_t st, const CentreArgs &a, int mode);

// the split pipeline (forward / backward launches) always runs one CTA per centre; the fused
// kernel also exists with 2 and 4 CTAs per centre (a.cs, chosen by the host for small launches)
template <int PC, bool ON>
int launch_centre_se(cudaStream_t st, const CentreArgs &a) {
    if constexpr (ON) return launch_centre<MODE_FUSED_SE, PC, 1>(st, a);
    else return launch_centre<MODE_FUSED, PC, 1>(st, a);
}
#define GAPCU_CENTRE_INSTANCE_(PC, SE_ON)                                                \
    namespace gapcu {                                                                     \
    int launch_centre_p##PC(cudaStream_t st, const CentreArgs &a, int mode) {             \
        if (mode == MODE_FWD) return launch_centre<MODE_FWD, PC, 1>(st, a);               \
        if (mode == MODE_BWD) return launch_centre<MODE_BWD, PC, 1>(st, a);               \
        /* a cluster launch the device cannot place falls back to one CTA per centre */   \
        if (a.cs == 4 && launch_centre<MODE_FUSED, PC, 4>(st, a) == 0) return 0;          \
        if (a.cs == 2 && launch_centre<MODE_FUSED, PC, 2>(st, a) == 0) return 0;          \
        if (a.cs > 1) cudaGetLastError();                                                 \
        if (mode == MODE_FUSED_SE) return launch_centre_se<PC, SE_ON>(st, a);             \
        return launch_centre<MODE_FUSED, PC, 1>(st, a);                                   \
    }                                                                                     \
    }
// instances with / without the parked-exponentials variant (only the two small capacity tiers carry it)
#define GAPCU_CENTRE_INSTANCE(PC) GAPCU_CENTRE_INSTANCE_(PC, false)
#define GAPCU_CENTRE_INSTANCE_SE(PC) GAPCU_CENTRE_INSTANCE_(PC, true)

}  // namespace gapcu
