// desc.cu -- K2 (forward) and K4 (backward) of the weighted atom-centred symmetry
// functions: one CTA per centre atom.
//
// What is computed (reference loops wacsf.f90:65-795):
//   radial  type 1  G = sum_j exp(-a r^2) fc(r)            wacsf.f90:69-162
//           type 3  G = sum_j exp(-4 (r-rs)^2) fc(r)       wacsf.f90:436-527
//   angular type 2/4  G = sum_{j<k} (1 +- cos) exp(-a (rij^2+rik^2+rjk^2)) fc fc fc
//                                                          wacsf.f90:169-432, 534-791
//   each in an unweighted channel ii and a species-weighted channel ii+nsf.
// The backward kernel never forms the reference's dense dxdy(D,N,N,3): it
// recomputes the geometry and contracts dG/dr with dE/dG on the fly, leaving per
// neighbour slot the gradient dE_i/dx_slot, plus dE_i/dx_i and the centre's
// strs contraction (gap_calc.f90:177-203 does the same sums through dxdy/strs).
//
// Work layout per centre:
//   1. stage the neighbour list in shared memory: absolute image coordinates,
//      distance, 1/r, species weight, and fc / fc' for every cutoff class the
//      neighbour belongs to (classes = distinct cutoffs, descending);
//   2. radial functions: one thread per neighbour;
//   3. angular functions, in row chunks of <= LCAP candidate pairs:
//        A. every pair (a<b) is tested ONCE with the exact reference arithmetic
//           (rjk from absolute coordinates, squared-distance thresholds that are
//           equivalent to the reference's  sqrt(..) > cutoff  test); survivors are
//           tagged with the number of classes they belong to ("bucket") and
//           compacted into a list,
//        B. the list is counting-sorted by bucket so a warp works on 32 triplets
//           that need the same classes: geometry once per triplet, one sincos per
//           (triplet, class), one exp per (triplet, class, distinct alpha);
//      forward: values are summed into lane-spread shared accumulators,
//      backward: the three leg scalars are scattered to per-neighbour accumulators
//                (dE/dx_j = A_j * d_j - V_j, see below).
#include <cstdint>

#include "device_types.cuh"
#include "geom.cuh"
#include "launch.cuh"

namespace gapcu {

constexpr int CT = 256;       // threads per centre CTA
constexpr int NW = CT / 32;
constexpr int LCAP = 4096;    // candidate pairs per chunk (list capacity)
constexpr int GS = 32;        // lane spread of the forward accumulators
constexpr int MAXC_D = 16;

struct SmemLayout {
    int dtab, x, r, ir, w, fc, dfc, gacc, du, acc, red, itab, U, S, ctl, nc, total;
};

__host__ __device__ inline SmemLayout make_layout(int n_dtab, int n_itab, int ncls, int D, int pcap, bool bwd) {
    SmemLayout L;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    L.dtab = take(8 * n_dtab);
    L.x = take(8 * 3 * pcap);
    L.r = take(8 * pcap);
    L.ir = take(8 * pcap);
    L.w = take(8 * pcap);
    L.fc = take(8 * ncls * pcap);
    L.dfc = bwd ? take(8 * ncls * pcap) : 0;
    L.gacc = bwd ? 0 : take(8 * D * GS);
    L.du = bwd ? take(8 * D) : 0;
    L.acc = bwd ? take(8 * 4 * pcap) : 0;
    L.red = take(8 * NW * 9);
    L.itab = take(4 * n_itab);
    L.U = take(4 * LCAP);
    L.S = take(4 * LCAP);
    L.ctl = take(4 * 8 * (MAXC_D + 2));
    L.nc = take(pcap);
    L.total = o;
    return L;
}

size_t centre_smem_bytes(const PlanDev &plan, int pcap, bool backward) {
    return (size_t)make_layout(plan.n_dtab, plan.n_itab, plan.ncls, plan.D, pcap, backward).total;
}

// control block (ints) in shared memory
struct Ctl {
    int nU;                  // items in the unsorted list
    int TB;                  // batches in this chunk
    int hist[MAXC_D + 2];    // items per bucket
    int cur[MAXC_D + 2];     // counting-sort cursors
    int base[MAXC_D + 2];    // start of each order slot in S   (order o <-> bucket ncls-o)
    int cnt[MAXC_D + 2];     // items of each order slot
    int bq[MAXC_D + 2];      // batches of each order slot
    int bp[MAXC_D + 2];      // first batch of each order slot
};
static_assert(sizeof(Ctl) <= 4 * 8 * (MAXC_D + 2), "Ctl does not fit its shared-memory slot");

template <bool BWD>
__global__ void __launch_bounds__(CT, 2) k_centre(CentreArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const PlanDev &pl = a.plan;
    const int ncls = pl.ncls, D = pl.D, nsf = pl.nsf, pcap = a.pcap;
    const SmemLayout L = make_layout(pl.n_dtab, pl.n_itab, ncls, D, pcap, BWD);
    double *s_dtab = (double *)(smem + L.dtab);
    double *s_x = (double *)(smem + L.x);  // [3][pcap]
    double *s_r = (double *)(smem + L.r);
    double *s_ir = (double *)(smem + L.ir);
    double *s_w = (double *)(smem + L.w);
    double *s_fc = (double *)(smem + L.fc);    // [ncls][pcap]
    double *s_dfc = (double *)(smem + L.dfc);  // [ncls][pcap] (backward)
    double *s_gacc = (double *)(smem + L.gacc);  // [D][GS]    (forward)
    double *s_du = (double *)(smem + L.du);      // [D]        (backward)
    double *s_acc = (double *)(smem + L.acc);    // [4][pcap]  (backward): A, Vx, Vy, Vz
    double *s_red = (double *)(smem + L.red);
    int *s_itab = (int *)(smem + L.itab);
    uint32_t *s_U = (uint32_t *)(smem + L.U);
    uint32_t *s_S = (uint32_t *)(smem + L.S);
    Ctl *ctl = (Ctl *)(smem + L.ctl);
    unsigned char *s_nc = smem + L.nc;

    const int i = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int P = a.nbr_cnt[i];
    if (P > pcap || P > a.cap) {  // host re-runs with a larger capacity
        if (tid == 0) atomicExch(&a.flags->overflow, 1);
        return;
    }
    const StructDev &sd = a.structs[a.sid[i]];
    const int ntot = a.ntot;

    for (int t = tid; t < pl.n_dtab; t += CT) s_dtab[t] = pl.dtab[t];
    for (int t = tid; t < pl.n_itab; t += CT) s_itab[t] = pl.itab[t];
    if (BWD) {
        for (int t = tid; t < D; t += CT) s_du[t] = a.dEdG[(size_t)i * D + t];
        for (int t = tid; t < 4 * pcap; t += CT) s_acc[t] = 0.0;
    } else {
        for (int t = tid; t < D * GS; t += CT) s_gacc[t] = 0.0;
    }
    __shared__ double s_lat[9];
    if (tid < 9) s_lat[tid] = sd.lat[tid];
    __syncthreads();

    const double *rc = s_dtab + pl.o_rc, *t2 = s_dtab + pl.o_t2, *pirc = s_dtab + pl.o_pirc;
    const double *rad_p = s_dtab + pl.o_rad_p, *grp_alpha = s_dtab + pl.o_grp_alpha;
    const double *asf_lambda = s_dtab + pl.o_asf_lambda;
    const int *rad_ii = s_itab + pl.o_rad_ii, *rad_cls = s_itab + pl.o_rad_cls, *rad_type = s_itab + pl.o_rad_type;
    const int *cls_grp = s_itab + pl.o_cls_grp, *grp_sf = s_itab + pl.o_grp_sf, *asf_ii = s_itab + pl.o_asf_ii;

    const double xi = a.pos[i], yi = a.pos[ntot + i], zi = a.pos[2 * ntot + i];
    unsigned long long wk_pc = 0, wk_rad = 0;  // work counters (forward only)

    // ---- 1+2: stage neighbours, radial functions --------------------------
    for (int s = tid; s < P; s += CT) {
        int jl, n1, n2, n3;
        nbr_unkey(a.nbr_keys[(size_t)i * a.cap + s], jl, n1, n2, n3);
        const int j = sd.atom_off + jl;
        double ox, oy, oz;
        const double dis = image_distance(a.pos, ntot, j, s_lat, n1, n2, n3, xi, yi, zi, ox, oy, oz);
        const double wj = a.wgt[j];
        s_x[s] = ox; s_x[pcap + s] = oy; s_x[2 * pcap + s] = oz;
        s_r[s] = dis;
        const double ir = 1.0 / dis;
        s_ir[s] = ir;
        s_w[s] = wj;
        int nc = 0;
        while (nc < ncls && !(dis > rc[nc])) nc++;  // reference: "if (rij.gt.cutoff) cycle"
        s_nc[s] = (unsigned char)nc;
        for (int c = 0; c < nc; c++) {
            double sn, cs;
            sincos(dis * pirc[c], &sn, &cs);
            s_fc[c * pcap + s] = 0.5 * (cs + 1.0);
            if (BWD) s_dfc[c * pcap + s] = -0.5 * pirc[c] * sn;
        }
        wk_pc += nc;
        double cacc = 0.0;
        for (int q = 0; q < pl.n_rad; q++) {
            const int c = rad_cls[q];
            if (c >= nc) continue;
            const int ii = rad_ii[q];
            const double fc = s_fc[c * pcap + s];
            double ex, dgf;  // value = ex*fc ; dgf = d(ln-ish) factor for the derivative
            if (rad_type[q] == 1) {
                const double al = rad_p[q];
                ex = exp(-al * dis * dis);
                dgf = -2.0 * al * dis;
            } else {
                const double d = dis - rad_p[q];
                ex = exp(-4.0 * d * d);
                dgf = -8.0 * d;
            }
            if (BWD) {
                const double dg = ex * (dgf * fc + s_dfc[c * pcap + s]);
                cacc += (s_du[ii] + wj * s_du[ii + nsf]) * dg;
            } else {
                const double g = ex * fc;
                atomicAdd(&s_gacc[ii * GS + lane], g);
                atomicAdd(&s_gacc[(ii + nsf) * GS + lane], g * wj);
                wk_rad++;
            }
        }
        if (BWD) s_acc[s] = cacc * ir;
    }
    __syncthreads();

    if (!BWD && tid < ncls && cls_grp[tid + 1] > cls_grp[tid]) {
        // sum_c Q_c of SURVEY.md 8(d): candidate pairs of every angular cutoff class
        unsigned long long pc = 0;
        for (int s = 0; s < P; s++) pc += (s_nc[s] > tid);
        atomicAdd(&a.flags->work[8], pc * (pc - 1) / 2);
    }
    // ---- 3: angular functions ---------------------------------------------
    unsigned long long wk_cand = 0, wk_trip = 0, wk_tc = 0, wk_tsf = 0;
    const uint32_t angmask = pl.ang_prefix_mask;
    int a0 = 0;
    while (angmask && a0 < P - 1) {
        int a1 = a0, tot = 0;
        while (a1 < P - 1) {
            const int len = P - 1 - a1;
            if (tot + len > LCAP && a1 > a0) break;
            tot += len;
            a1++;
        }
        if (tid < MAXC_D + 2) { ctl->hist[tid] = 0; ctl->cur[tid] = 0; }
        if (tid == 0) ctl->nU = 0;
        __syncthreads();
        // -- phase A: test every candidate pair once, compact the survivors
        for (int ra = a0 + wid; ra < a1; ra += NW) {
            const int nca = s_nc[ra];
            if (nca == 0) continue;
            const double xa = s_x[ra], ya = s_x[pcap + ra], za = s_x[2 * pcap + ra];
            for (int b0 = ra + 1; b0 < P; b0 += 32) {
                const int rb = b0 + lane;
                int bk = 0;
                if (rb < P) {
                    const int ncb = s_nc[rb];
                    if (ncb) {
                        const double rjk2 = pair_dist2(xa, ya, za, s_x[rb], s_x[pcap + rb], s_x[2 * pcap + rb]);
                        int ncj = 0;
                        const int lim = min(nca, ncb);
                        while (ncj < lim && rjk2 <= t2[ncj]) ncj++;
                        bk = ncj;
                        if (!((angmask >> bk) & 1u)) bk = 0;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, bk > 0);
                if (!BWD && lane == 0) wk_cand += min(32, P - b0);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&ctl->nU, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (bk > 0) {
                        s_U[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)ra | ((uint32_t)rb << 10) | ((uint32_t)bk << 20);
                        atomicAdd(&ctl->hist[bk], 1);
                    }
                }
            }
        }
        __syncthreads();
        // -- counting sort by bucket, heavy buckets first
        if (tid == 0) {
            int off = 0, bp = 0;
            for (int o = 0; o < ncls; o++) {
                const int v = ncls - o, n = ctl->hist[v];
                ctl->base[o] = off; ctl->cnt[o] = n; ctl->bq[o] = (n + 31) >> 5; ctl->bp[o] = bp;
                off += n; bp += (n + 31) >> 5;
            }
            ctl->bp[ncls] = bp;
            ctl->TB = bp;
            if (!BWD) {
                wk_trip += ctl->nU;
                for (int v = 1; v <= ncls; v++) {
                    int nac = 0, nas = 0;
                    for (int c = 0; c < v; c++) {
                        const int g0 = cls_grp[c], g1 = cls_grp[c + 1];
                        if (g1 > g0) { nac++; nas += grp_sf[g1] - grp_sf[g0]; }
                    }
                    wk_tc += (unsigned long long)ctl->hist[v] * nac;
                    wk_tsf += (unsigned long long)ctl->hist[v] * nas;
                }
            }
        }
        __syncthreads();
        const int nU = ctl->nU;
        for (int t = tid; t < nU; t += CT) {
            const uint32_t it = s_U[t];
            const int v = it >> 20;
            const int o = ncls - v;
            s_S[ctl->base[o] + atomicAdd(&ctl->cur[v], 1)] = it;
        }
        __syncthreads();
        // -- phase B: one warp per batch of 32 triplets of the same bucket
        const int TB = ctl->TB;
        for (int g = wid; g < TB; g += NW) {
            int o = 0;
            while (o + 1 < ncls && g >= ctl->bp[o + 1]) o++;
            const int v = ncls - o, q = g - ctl->bp[o], Q = ctl->bq[o], n = ctl->cnt[o];
            const int idx = lane * Q + q;  // lanes far apart in the (row-major) list -> distinct rows
            if (idx >= n) continue;
            const uint32_t it = s_S[ctl->base[o] + idx];
            const int ra = it & 1023, rb = (it >> 10) & 1023;
            const double xa = s_x[ra], ya = s_x[pcap + ra], za = s_x[2 * pcap + ra];
            const double xb = s_x[rb], yb = s_x[pcap + rb], zb = s_x[2 * pcap + rb];
            const double rja = s_r[ra], rkb = s_r[rb], ira = s_ir[ra], irb = s_ir[rb];
            const double rjk2 = pair_dist2(xa, ya, za, xb, yb, zb);
            const double rjk = sqrt(rjk2);
            const double ra2 = rja * rja, rb2 = rkb * rkb;
            const double cosv = (ra2 + rb2 - rjk2) * 0.5 * ira * irb;
            const double ssum = ra2 + rb2 + rjk2;
            const double ww = s_w[ra] * s_w[rb];
            if (!BWD) {
                for (int c = 0; c < v; c++) {
                    const int g0 = cls_grp[c], g1 = cls_grp[c + 1];
                    if (g0 == g1) continue;
                    const double fjk = 0.5 * (cos(rjk * pirc[c]) + 1.0);
                    const double phi = s_fc[c * pcap + ra] * s_fc[c * pcap + rb] * fjk;
                    for (int gg = g0; gg < g1; gg++) {
                        const double pe = phi * exp(-grp_alpha[gg] * ssum);
                        for (int sf = grp_sf[gg]; sf < grp_sf[gg + 1]; sf++) {
                            const int ii = asf_ii[sf];
                            const double val = (1.0 + asf_lambda[sf] * cosv) * pe;
                            atomicAdd(&s_gacc[ii * GS + lane], val);
                            atomicAdd(&s_gacc[(ii + nsf) * GS + lane], val * ww);
                        }
                    }
                }
            } else {
                const double irjk = 1.0 / rjk;
                const double u1 = irb - cosv * ira, u2 = ira - cosv * irb, u3 = -rjk * ira * irb;
                double cij = 0.0, cik = 0.0, cjk = 0.0;
                for (int c = 0; c < v; c++) {
                    const int g0 = cls_grp[c], g1 = cls_grp[c + 1];
                    if (g0 == g1) continue;
                    double sn, cs;
                    sincos(rjk * pirc[c], &sn, &cs);
                    const double fjk = 0.5 * (cs + 1.0), dfjk = -0.5 * pirc[c] * sn;
                    const double fa = s_fc[c * pcap + ra], fb = s_fc[c * pcap + rb];
                    const double dfa = s_dfc[c * pcap + ra], dfb = s_dfc[c * pcap + rb];
                    const double fab = fa * fb, phi = fab * fjk;
                    double S1 = 0.0, S2 = 0.0, S3 = 0.0;
                    for (int gg = g0; gg < g1; gg++) {
                        const double al = grp_alpha[gg];
                        const double e = exp(-al * ssum);
                        double S3g = 0.0;
                        for (int sf = grp_sf[gg]; sf < grp_sf[gg + 1]; sf++) {
                            const int ii = asf_ii[sf];
                            const double lam = asf_lambda[sf];
                            const double t = (s_du[ii] + ww * s_du[ii + nsf]) * e;  // gamma * e
                            const double tA = t * (1.0 + lam * cosv);
                            S1 += lam * t;
                            S2 += tA;
                            S3g += tA;
                        }
                        S3 += al * S3g;
                    }
                    const double pS1 = phi * S1, pS3 = 2.0 * phi * S3;
                    cij += pS1 * u1 - pS3 * rja + S2 * (dfa * fb * fjk);
                    cik += pS1 * u2 - pS3 * rkb + S2 * (fa * dfb * fjk);
                    cjk += pS1 * u3 - pS3 * rjk + S2 * (fab * dfjk);
                }
                // dE/dx_j = (gij+gjk) d_j - gjk d_k ; dE/dx_k = (gik+gjk) d_k - gjk d_j
                const double gij = cij * ira, gik = cik * irb, gjk = cjk * irjk;
                atomicAdd(&s_acc[ra], gij + gjk);
                atomicAdd(&s_acc[rb], gik + gjk);
                atomicAdd(&s_acc[pcap + ra], gjk * (xb - xi));
                atomicAdd(&s_acc[2 * pcap + ra], gjk * (yb - yi));
                atomicAdd(&s_acc[3 * pcap + ra], gjk * (zb - zi));
                atomicAdd(&s_acc[pcap + rb], gjk * (xa - xi));
                atomicAdd(&s_acc[2 * pcap + rb], gjk * (ya - yi));
                atomicAdd(&s_acc[3 * pcap + rb], gjk * (za - zi));
            }
        }
        __syncthreads();
        a0 = a1;
    }

    // ---- epilogue -------------------------------------------------------------
    if (!BWD) {
        for (int k = tid; k < D; k += CT) {
            double sum = 0.0;
#pragma unroll 8
            for (int l = 0; l < GS; l++) sum += s_gacc[k * GS + l];
            a.G[(size_t)i * D + k] = sum;
        }
        // work counters: warp-reduce the per-thread parts, one atomic per CTA each
        unsigned long long v0 = wk_pc, v1 = wk_rad, v2 = wk_cand;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if (lane == 0) {
            atomicAdd(&a.flags->work[2], v0);
            atomicAdd(&a.flags->work[7], v1);
            atomicAdd(&a.flags->work[3], v2);
        }
        if (tid == 0) {
            atomicAdd(&a.flags->work[0], 1ull);
            atomicAdd(&a.flags->work[1], (unsigned long long)P);
            atomicAdd(&a.flags->work[4], wk_trip);
            atomicAdd(&a.flags->work[5], wk_tc);
            atomicAdd(&a.flags->work[6], wk_tsf);
        }
    } else {
        // per neighbour gradient, centre gradient, strs contraction
        double acc9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // gself xyz, vir xx xy xz yy yz zz
        for (int s = tid; s < P; s += CT) {
            const double dx = s_x[s] - xi, dy = s_x[pcap + s] - yi, dz = s_x[2 * pcap + s] - zi;
            const double A = s_acc[s];
            const double gx = A * dx - s_acc[pcap + s], gy = A * dy - s_acc[2 * pcap + s], gz = A * dz - s_acc[3 * pcap + s];
            double *fp = a.fpair + ((size_t)i * a.cap + s) * 3;
            fp[0] = gx; fp[1] = gy; fp[2] = gz;
            acc9[0] -= gx; acc9[1] -= gy; acc9[2] -= gz;
            acc9[3] += dx * gx; acc9[4] += dx * gy; acc9[5] += dx * gz;
            acc9[6] += dy * gy; acc9[7] += dy * gz; acc9[8] += dz * gz;
        }
#pragma unroll
        for (int q = 0; q < 9; q++) {
            double v = acc9[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[wid * 9 + q] = v;
        }
        __syncthreads();
        if (tid < 9) {
            double v = 0.0;
            for (int w = 0; w < NW; w++) v += s_red[w * 9 + tid];
            if (tid < 3) a.gself[(size_t)i * 3 + tid] = v;
            else a.vir[(size_t)i * 6 + (tid - 3)] = v;
        }
    }
}

static int set_smem_attr(const void *fn, size_t bytes) {
    if (bytes > 227 * 1024) return -1;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : -2;
}

int launch_forward(cudaStream_t st, const CentreArgs &a, long *launches) {
    size_t sm = centre_smem_bytes(a.plan, a.pcap, false);
    if (set_smem_attr((const void *)k_centre<false>, sm)) return -1;
    k_centre<false><<<a.ntot, CT, sm, st>>>(a);
    if (launches) *launches += 1;
    return 0;
}

int launch_backward(cudaStream_t st, const CentreArgs &a, long *launches) {
    size_t sm = centre_smem_bytes(a.plan, a.pcap, true);
    if (set_smem_attr((const void *)k_centre<true>, sm)) return -1;
    k_centre<true><<<a.ntot, CT, sm, st>>>(a);
    if (launches) *launches += 1;
    return 0;
}

}  // namespace gapcu
