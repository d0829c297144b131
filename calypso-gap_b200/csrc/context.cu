// context.cu -- host orchestration and the C ABI (include/gapcu.h).
//
// Pipeline of one evaluation (all on ctx->stream, no host sync in between once
// the neighbour capacity is known):
//   K1  neigh.cu    bin -> scan -> fill -> per-centre list (reference order)
//   K2  desc.cu     forward descriptors G[N][D]
//   K3  gpr.cu      DMMA GPR: e_i and dE/dG
//   K4  desc.cu     backward: per-slot gradients, centre gradient, strs contraction
//   K5  gather.cu   force gather (mirror-pair lookup) + per-structure E / stress
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gapcu.h"
#include "device_types.cuh"
#include "launch.cuh"
#include "fastmath.cuh"
#include "potential.hpp"

using namespace gapcu;

struct gapcu_ctx;
static void domain_destroy(gapcu_ctx *c);   // domain_host.inc
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
extern "C" const char *gapcu_last_error(void) { return g_err.c_str(); }

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(GAPCU_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

namespace {

template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    bool view = false;   // p points into another allocation (the results block): not owned
    cudaError_t ensure(size_t want) {
        if (view) return want <= n ? cudaSuccess : cudaErrorInvalidValue;
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        size_t grow = want + want / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, grow * sizeof(T));
        if (e == cudaSuccess) n = grow;
        return e;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; n = 0; }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Every public entry point runs on its context's device and gives the caller's current
// device back on return (a host application such as PyTorch keeps its own per-thread device).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct gapcu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // decomposed runs: the gather of the owned atoms runs here, beside the gradient return on `stream`
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_centres = nullptr, ev_owned = nullptr;
    bool owned_gather_pending = false;
    // ---- potential
    bool have_sf = false, have_gpr = false;
    std::vector<int> z;
    std::vector<double> w;
    SfPlan plan;
    DBuf<int> d_itab;
    DBuf<double> d_dtab;
    int M = 0, D = 0, Mp = 0, Dp = 0;
    std::vector<double> h_theta, h_mm, h_coeff;  // cached GPR data (C order)
    DBuf<double> d_mm_raw, d_theta_raw, d_coeff_raw, d_Mt, d_MtT, d_mn, d_coeff, d_cmean, d_itheta, d_exp2;
    int pipeline = 0;  // 0 auto, 1 split (K2 -> DMMA K3 -> K4), 2 fused single centre kernel
    int cluster = 0;   // CTAs per centre of the fused kernel: 0 auto, 1, 2 or 4
    bool direct_ok = false;   // every structure is small enough for k_neigh_direct (set_structures)
    bool h2d_pending = false; // the pinned staging buffer feeds a copy nobody has waited for yet
    // ---- structures
    int nstruct = 0, ntot = 0, nbins = 0;
    int n_centres = 0;            // atoms evaluated as centres: all of them, or this rank's owned atoms (then ntot = owned + ghost capacity)
    double rcut = 0.0;
    std::vector<StructDev> h_structs;
    std::vector<int> h_natoms;
    DBuf<StructDev> d_structs;
    DBuf<int> d_hist, d_scan_sums;   // scratch of the multi-CTA centre ordering / cell scan
    int *d_blk = nullptr;            // structure of every block of cells (in the inputs block)
    int nblocks = 0;                 // blocks of cells for k_neigh_block; 0: some structure does not fit that form
    bool k1_legacy = false;          // a block overflowed once: stay with k_neigh for this context
    DBuf<int> d_sid, d_arank, d_bin_count, d_bin_start, d_bin_atoms, d_nbr_cnt, d_skin_cnt, d_order;
    DBuf<int4> d_abin, d_sabin;   // per atom / in bin order: (bin | atom, wrap offsets)
    DBuf<double> d_spos;          // coordinates in bin order
    DBuf<double> d_finpart;       // per (structure, chunk) partial sums of E and the strs contraction
    int max_natoms = 0;           // largest structure of the batch
    DBuf<double> d_pos, d_wgt, d_G, d_dEdG, d_eatom, d_fpair, d_gself, d_vir, d_force, d_out8, d_mindis, d_epart, d_accpart;
    DBuf<uint64_t> d_keys, d_skin_keys;   // exact lists (dis <= rcut) and candidate lists (dis <= rcut + skin)
    DBuf<double> d_pos_build;             // positions the skin lists were built from (Verlet reuse)
    double skin_user = 0.0;               // gapcu_ctx_set_skin; the effective skin never drops below SKIN_FLOOR
    bool lists_valid = false;             // skin lists + pos_build describe the resident structures
    bool reuse_next = false;              // the next compute may re-filter instead of rebuilding
    bool last_reuse = false;              // how the pass in flight was run (a stale flag then means: rebuild)
    uint32_t *dbg_trip = nullptr; int *dbg_trip_cnt = nullptr; int dbg_trip_cap = 0;   // gapcu_ctx_debug_triplets
    DBuf<uint32_t> d_stash;
    DBuf<double2> d_estash;       // parked exponentials of the fused kernel, per persistent CTA
    int sm_count = 0;
    DBuf<DevFlags> d_flags;
    // flags | out8 | force live in ONE allocation so that a call reads its results back with one copy
    DBuf<unsigned char> d_results;
    size_t res_o_out = 0, res_o_f = 0;   // byte offsets of out8 and force in d_results
    // likewise structs | sid | pos | wgt ("inputs block"): one host-to-device copy per call
    DBuf<unsigned char> d_inputs;
    // what the device copy of the inputs block still holds from the last set_structures (an MD or relaxation loop
    // sends the same structure sizes and species every step: only cell records and positions travel then)
    bool in_valid = false, in_wgt_ok = false;
    std::vector<int> in_natoms, in_species;
    size_t in_nblocks = 0;
    unsigned long in_sf_version = 0, sf_version = 0;
    DBuf<unsigned char> d_flush;
    // ---- spatial decomposition over ranks (domain_host.inc) + NCCL (loaded lazily with dlopen)
    DomainDev dom = {0, {1, 1, 1}, {0, 0, 0}, {0.0, 0.0, 0.0}};
    struct DomainState *ds = nullptr;
    void *nccl_comm = nullptr;
    int nccl_ranks = 1, nccl_rank = 0;
    int cap = 0, pcap = 0, last_ntot = -1;
    bool pcap_known = false;
    int last_lgrad = 1;
    bool computed = false;
    DevFlags h_flags;
    long launches = 0;
    // pinned staging
    void *h_pin = nullptr;
    size_t h_pin_bytes = 0;
    // timing instrumentation
    cudaEvent_t stage_ev[GAPCU_NSTAGE + 1];
    bool stage_ev_init = false;

    double skin() const { return std::max(skin_user, 1e-9 * std::max(1.0, rcut)); }   // SKIN_FLOOR: see neigh.cu header
    PlanDev plan_dev() const {
        PlanDev p;
        p.itab = d_itab.p; p.dtab = d_dtab.p;
        p.n_itab = (int)plan.itab.size(); p.n_dtab = (int)plan.dtab.size();
        p.nsf = plan.nsf; p.D = plan.D; p.ncls = plan.ncls; p.n_rad = plan.n_rad; p.n_grp = plan.n_grp;
        p.o_rad_ii = plan.o_rad_ii; p.o_rad_cls = plan.o_rad_cls; p.o_rad_type = plan.o_rad_type;
        p.o_grp_iplus = plan.o_grp_iplus; p.o_grp_iminus = plan.o_grp_iminus;
        p.o_rad_p = plan.o_rad_p; p.o_grp_alpha = plan.o_grp_alpha;
        return p;
    }
    ClassTab class_tab() const {
        ClassTab t;
        memset(&t, 0, sizeof t);
        for (int c = 0; c < plan.ncls; c++) { t.rc[c] = plan.rc[c]; t.t2[c] = plan.t2[c]; t.pirc[c] = plan.pirc[c]; }
        for (int c = plan.ncls; c < MAXC_DEV; c++) t.t2[c] = -1.0;   // no squared distance passes an absent class
        for (int c = 0; c <= MAXC_DEV; c++) t.grp_begin[c] = plan.grp_begin[c];
        t.angmask = plan.ang_prefix_mask;
        return t;
    }
    int pin(size_t bytes) {
        if (bytes <= h_pin_bytes) return 0;
        if (h_pin) { cudaStreamSynchronize(stream); cudaFreeHost(h_pin); }   // pending copies may still read the old buffer
        h_pin = nullptr; h_pin_bytes = 0;
        size_t grow = bytes + bytes / 4 + 4096;
        if (cudaMallocHost(&h_pin, grow) != cudaSuccess) return -1;
        h_pin_bytes = grow;
        return 0;
    }
};

static_assert(sizeof(DevFlags) <= 512, "DevFlags outgrew its slot in the results block");
static cudaError_t ensure_results(gapcu_ctx *c, size_t nstruct, size_t NT) {
    const size_t o_out = 512, o_f = o_out + ((64 * nstruct + 255) & ~(size_t)255), total = o_f + 24 * NT;
    cudaError_t e = c->d_results.ensure(total);
    if (e != cudaSuccess) return e;
    c->d_flags.view = c->d_out8.view = c->d_force.view = true;
    c->d_flags.p = (DevFlags *)c->d_results.p; c->d_flags.n = 1;
    c->d_out8.p = (double *)(c->d_results.p + o_out); c->d_out8.n = 8 * nstruct;
    c->d_force.p = (double *)(c->d_results.p + o_f); c->d_force.n = 3 * NT;
    c->res_o_out = o_out; c->res_o_f = o_f;
    return cudaSuccess;
}

// inputs block: same 256-byte aligned layout on the device as in the pinned staging buffer
struct InputLayout { size_t o_structs, o_sid, o_pos, o_wgt, o_blk, total; };
static InputLayout input_layout(size_t nstruct, size_t NT, size_t nblocks = 0) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    InputLayout L;
    L.o_structs = 0;
    L.o_sid = al(sizeof(StructDev) * nstruct);
    L.o_pos = al(L.o_sid + sizeof(int) * NT);
    L.o_wgt = al(L.o_pos + sizeof(double) * 3 * NT);
    L.o_blk = al(L.o_wgt + sizeof(double) * NT);
    L.total = al(L.o_blk + sizeof(int) * nblocks);
    return L;
}
static cudaError_t ensure_inputs(gapcu_ctx *c, size_t nstruct, size_t NT, size_t nblocks = 0) {
    const InputLayout L = input_layout(nstruct, NT, nblocks);
    c->in_valid = false;   // whoever lays the block out anew owns its contents (set_structures_impl re-validates)
    cudaError_t e = c->d_inputs.ensure(L.total);
    if (e != cudaSuccess) return e;
    c->d_structs.view = c->d_sid.view = c->d_pos.view = c->d_wgt.view = true;
    c->d_structs.p = (StructDev *)(c->d_inputs.p + L.o_structs); c->d_structs.n = nstruct;
    c->d_sid.p = (int *)(c->d_inputs.p + L.o_sid); c->d_sid.n = NT;
    c->d_pos.p = (double *)(c->d_inputs.p + L.o_pos); c->d_pos.n = 3 * NT;
    c->d_wgt.p = (double *)(c->d_inputs.p + L.o_wgt); c->d_wgt.n = NT;
    c->d_blk = (int *)(c->d_inputs.p + L.o_blk);
    return cudaSuccess;
}

// ---------------------------------------------------------------------------
// context life cycle
// ---------------------------------------------------------------------------
extern "C" int gapcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" gapcu_ctx *gapcu_ctx_create(int device) {
    int n = gapcu_device_count();
    if (n <= 0) { fail(GAPCU_ENODEV, "no CUDA device: gapcu has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= n) { fail(GAPCU_EARG, "bad device index"); return nullptr; }
    DeviceGuard dg_(device);
    {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != device) { fail(GAPCU_ECUDA, "cudaSetDevice failed"); return nullptr; }
    }
    gapcu_ctx *c = new gapcu_ctx();
    c->device = device;
    // the main stream at the highest priority, the side stream at the lowest: what the side stream runs (the owned
    // atoms' gather of a decomposed pass) then yields, CTA by CTA, to the gradient return on the main stream
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_centres, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_owned, cudaEventDisableTiming) != cudaSuccess) {
        fail(GAPCU_ECUDA, "cudaStreamCreate failed");
        delete c;
        return nullptr;
    }
    memset(&c->h_flags, 0, sizeof c->h_flags);
    // 2^(j/32) for exp_neg, then (cos, sin)(k/16) for sincos_tab (fastmath.cuh): one small device table
    double t32[32 + 2 * SINCOS_TAB_N];
    fill_exp2_table(t32);
    fill_sincos_table((SinCosEntry *)(t32 + 32));
    if (c->d_exp2.ensure(32 + 2 * SINCOS_TAB_N) != cudaSuccess ||
        cudaMemcpy(c->d_exp2.p, t32, sizeof t32, cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(GAPCU_ECUDA, "cudaMalloc failed");
        delete c;
        return nullptr;
    }
    if (const char *e = getenv("GAPCU_PIPELINE")) c->pipeline = !strcmp(e, "split") ? 1 : !strcmp(e, "fused") ? 2 : 0;
    if (const char *e = getenv("GAPCU_CLUSTER")) { const int v = atoi(e); c->cluster = (v == 1 || v == 2 || v == 4) ? v : 0; }
    return c;
}

extern "C" void gapcu_ctx_destroy(gapcu_ctx *c) {
    if (!c) return;
    DeviceGuard dg_(c->device);
    cudaStreamSynchronize(c->stream);
    c->d_itab.release(); c->d_dtab.release(); c->d_mm_raw.release(); c->d_theta_raw.release(); c->d_coeff_raw.release();
    c->d_Mt.release(); c->d_MtT.release(); c->d_exp2.release(); c->d_mn.release(); c->d_coeff.release(); c->d_cmean.release(); c->d_itheta.release();
    c->d_structs.release(); c->d_sid.release(); c->d_arank.release(); c->d_bin_count.release(); c->d_bin_start.release();
    c->d_bin_atoms.release(); c->d_nbr_cnt.release(); c->d_order.release(); c->d_abin.release(); c->d_sabin.release(); c->d_spos.release(); c->d_finpart.release(); c->d_pos.release(); c->d_wgt.release();
    c->d_G.release(); c->d_dEdG.release(); c->d_eatom.release(); c->d_fpair.release(); c->d_gself.release();
    c->d_vir.release(); c->d_force.release(); c->d_out8.release(); c->d_mindis.release(); c->d_keys.release();
    c->d_stash.release(); c->d_epart.release(); c->d_accpart.release(); c->d_flags.release(); c->d_results.release(); c->d_inputs.release(); c->d_flush.release();
    c->d_skin_keys.release(); c->d_skin_cnt.release(); c->d_pos_build.release(); c->d_hist.release(); c->d_scan_sums.release(); c->d_estash.release();
    domain_destroy(c);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->stage_ev_init) for (auto &e : c->stage_ev) cudaEventDestroy(e);
    if (c->ev_centres) cudaEventDestroy(c->ev_centres);
    if (c->ev_owned) cudaEventDestroy(c->ev_owned);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    cudaStreamDestroy(c->stream);
    delete c;
}

// ---------------------------------------------------------------------------
// potential
// ---------------------------------------------------------------------------
static int set_sf(gapcu_ctx *c, const std::vector<int> &z, const std::vector<double> &w,
                  const std::vector<int> &ntype, const std::vector<double> &alpha, const std::vector<double> &cutoff) {
    DeviceGuard dg_(c->device);
    try {
        c->plan = make_plan(ntype, alpha, cutoff);
    } catch (const std::exception &e) {
        return fail(GAPCU_ELIMIT, e.what());
    }
    if (c->plan.n_unknown)
        fprintf(stdout, " Unknown function type in gap_parameters (%d functions left at zero)\n", c->plan.n_unknown);
    c->z = z; c->w = w; c->sf_version++;
    // exponent arguments -alpha*(rij^2+rik^2+rjk^2), -alpha*r^2, -4 (r-rs)^2 must be <= 0 (exp_neg, fastmath.cuh;
    // arbitrarily negative ones are fine: the result saturates at ~2^-1021)
    for (size_t i = 0; i < ntype.size(); i++)
        if ((ntype[i] == 1 || ntype[i] == 2 || ntype[i] == 4) && alpha[i] < 0.0) return fail(GAPCU_ELIMIT, "negative symmetry-function alpha is not supported");
    CU(c->d_itab.ensure(c->plan.itab.size() + 1));
    CU(c->d_dtab.ensure(c->plan.dtab.size() + 1));
    CU(cudaMemcpyAsync(c->d_itab.p, c->plan.itab.data(), sizeof(int) * c->plan.itab.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_dtab.p, c->plan.dtab.data(), sizeof(double) * c->plan.dtab.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->have_sf = true;
    c->pcap_known = false;
    return 0;
}

static int pick_dp(int D) {
    const int nt = (D + 7) / 8;
    for (int cand : {2, 4, 6, 8, 9, 10, 12, 14, 16, 20, 24, 28, 32}) if (nt <= cand) return 8 * cand;  // k_gpr<NT> instances
    return -1;
}

// mm in C order [M][D]
static int set_gpr(gapcu_ctx *c, int M, int D, const double *theta, const double *mm, const double *coeff) {
    DeviceGuard dg_(c->device);
    if (M < 0 || D <= 0) return fail(GAPCU_EARG, "bad GPR sizes");
    int Dp = pick_dp(D);
    if (Dp < 0) return fail(GAPCU_ELIMIT, "des_len > 256 is beyond this build's DMMA tile set");
    if (c->have_gpr && c->M == M && c->D == D && !memcmp(c->h_theta.data(), theta, sizeof(double) * D) &&
        !memcmp(c->h_coeff.data(), coeff, sizeof(double) * M) && !memcmp(c->h_mm.data(), mm, sizeof(double) * (size_t)M * D))
        return 0;  // unchanged: keep the device copy
    for (int k = 0; k < D; k++)
        if (!(theta[k] != 0.0)) return fail(GAPCU_EARG, "theta contains a zero");
    c->h_theta.assign(theta, theta + D);
    c->h_coeff.assign(coeff, coeff + M);
    c->h_mm.assign(mm, mm + (size_t)M * D);
    c->M = M; c->D = D; c->Dp = Dp; c->Mp = round_up(std::max(M, 1), 16);
    CU(c->d_mm_raw.ensure((size_t)M * D + 1)); CU(c->d_theta_raw.ensure(D)); CU(c->d_coeff_raw.ensure(M + 1));
    CU(c->d_Mt.ensure((size_t)c->Mp * Dp)); CU(c->d_MtT.ensure((size_t)c->Mp * Dp)); CU(c->d_mn.ensure(c->Mp)); CU(c->d_coeff.ensure(c->Mp));
    CU(c->d_cmean.ensure(Dp)); CU(c->d_itheta.ensure(Dp));
    CU(cudaMemcpyAsync(c->d_mm_raw.p, mm, sizeof(double) * (size_t)M * D, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_theta_raw.p, theta, sizeof(double) * D, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_coeff_raw.p, coeff, sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    launch_gpr_prepare(c->stream, M, D, c->d_mm_raw.p, c->d_theta_raw.p, c->d_coeff_raw.p, c->Mp, Dp, c->d_Mt.p,
                       c->d_MtT.p, c->d_mn.p, c->d_coeff.p, c->d_cmean.p, c->d_itheta.p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->have_gpr = true;
    return 0;
}

extern "C" int gapcu_ctx_set_potential(gapcu_ctx *c, int nspecies, const int *z, const double *w, int nsf,
                                       const int *ntype, const double *alpha, const double *cutoff, int nsparse,
                                       int des_len, const double *theta, const double *mm, const double *coeff) {
    if (!c) return fail(GAPCU_EARG, "null context");
    if (des_len != 2 * nsf) return fail(GAPCU_EARG, "des_len must equal 2*nsf (wacsf.f90:59,87)");
    int rc = set_sf(c, std::vector<int>(z, z + nspecies), std::vector<double>(w, w + nspecies),
                    std::vector<int>(ntype, ntype + nsf), std::vector<double>(alpha, alpha + nsf),
                    std::vector<double>(cutoff, cutoff + nsf));
    if (rc) return rc;
    return set_gpr(c, nsparse, des_len, theta, mm, coeff);
}

extern "C" int gapcu_ctx_load_potential(gapcu_ctx *c, const char *path) {
    if (!c) return fail(GAPCU_EARG, "null context");
    PotentialFile pf;
    try {
        pf = read_gap_parameters(path);
    } catch (const std::exception &e) {
        return fail(GAPCU_EFILE, e.what());
    }
    return gapcu_ctx_set_potential(c, (int)pf.z.size(), pf.z.data(), pf.w.data(), (int)pf.ntype.size(), pf.ntype.data(),
                                   pf.alpha.data(), pf.cutoff.data(), pf.nsparse, pf.des_len, pf.theta.data(),
                                   pf.mm.data(), pf.coeff.data());
}

extern "C" int gapcu_ctx_set_cluster(gapcu_ctx *c, int ctas_per_centre) {
    if (!c || !(ctas_per_centre == 0 || ctas_per_centre == 1 || ctas_per_centre == 2 || ctas_per_centre == 4))
        return fail(GAPCU_EARG, "CTAs per centre must be 0 (automatic), 1, 2 or 4");
    c->cluster = ctas_per_centre;
    return 0;
}

extern "C" int gapcu_ctx_set_pipeline(gapcu_ctx *c, int mode) {
    if (!c || mode < 0 || mode > 2) return fail(GAPCU_EARG, "bad pipeline mode");
    c->pipeline = mode;
    return 0;
}

// ---------------------------------------------------------------------------
// structures
// ---------------------------------------------------------------------------
static int lookup_weight(const gapcu_ctx *c, int species, double *w) {
    // gap_calc.f90:75-83; a species missing from the file is an error here
    int f = -1;
    for (int q = 0; q < (int)c->z.size(); q++) if (species == c->z[q]) f = q;  // last match wins, as in the reference loop
    if (f < 0) return fail(GAPCU_ESPECIES, "species " + std::to_string(species) + " is not in gap_parameters");
    *w = c->w[f];
    return 0;
}

// cell record of one periodic structure: lattice, image window (reference), cell-list grid for the
// candidate radius rskin, bin count kept proportional to the atom count
static int fill_struct(StructDev &sd, const double *lat9, double rcut, double rskin, int natoms, bool *direct) {
    CellInfo ci = make_cell(lat9, rcut, rskin);
    if (!(ci.volume > 0.0)) return fail(GAPCU_EARG, "singular lattice");
    for (int d = 0; d < 3; d++)
        if (ci.nabc[d] > 500) return fail(GAPCU_ENEIGH, "cell far smaller than rcut: neighbour list would exceed max_neighbor");
    long cells = (long)ci.nbin[0] * ci.nbin[1] * ci.nbin[2];
    const long lim = std::max(8l, 2l * natoms);
    while (cells > lim) {
        int d = 0;
        for (int q = 1; q < 3; q++) if (ci.nbin[q] > ci.nbin[d]) d = q;
        if (ci.nbin[d] <= 1) break;
        ci.nbin[d]--;
        cells = (long)ci.nbin[0] * ci.nbin[1] * ci.nbin[2];
    }
    memset(&sd, 0, sizeof sd);
    memcpy(sd.lat, ci.lat, sizeof sd.lat);
    memcpy(sd.inv, ci.inv, sizeof sd.inv);
    sd.volume = ci.volume;
    for (int d = 0; d < 3; d++)
        sd.spacing[d] = 1.0 / std::sqrt(ci.inv[d] * ci.inv[d] + ci.inv[3 + d] * ci.inv[3 + d] + ci.inv[6 + d] * ci.inv[6 + d]);
    for (int d = 0; d < 3; d++) {
        sd.nabc[d] = ci.nabc[d];
        sd.nbin[d] = ci.nbin[d];
        // layers of bins scanned either side (potential.cpp:make_cell; the bin count may have been reduced above)
        const double w = sd.spacing[d] / ci.nbin[d], full = rskin * (1.0 + 1e-9);
        sd.mscan[d] = w >= full ? 1 : w >= 0.5 * full ? 2 : ci.nabc[d] + 1;
        sd.org[d] = 0.0; sd.wid[d] = 1.0;
    }
    sd.open = 0;
    sd.natoms = natoms; sd.nbins = (int)cells;
    if (direct && (long)natoms * (2 * ci.nabc[0] + 1) * (2 * ci.nabc[1] + 1) * (2 * ci.nabc[2] + 1) > neighbor_direct_max_candidates()) *direct = false;
    return 0;
}

// Blocks of cells for k_neigh_block: up to 2 x 2 x 2 bins, shrunk until the bins around a block (block
// plus the scanned layers) number at most neighbor_block_max_bins() and are expected to hold well under
// neighbor_block_max_candidates() atoms.  Returns the number of blocks, 0 when no block shape qualifies.
static int plan_blocks(StructDev &sd, double atoms_in_region) {
    int bs[3];
    for (int d = 0; d < 3; d++) bs[d] = std::min(2, sd.nbin[d]);
    const double per_bin = atoms_in_region / std::max(1, sd.nbin[0] * sd.nbin[1] * sd.nbin[2]);
    for (;;) {
        const long ncb = (long)(bs[0] + 2 * sd.mscan[0]) * (bs[1] + 2 * sd.mscan[1]) * (bs[2] + 2 * sd.mscan[2]);
        if (ncb <= neighbor_block_max_bins() && ncb * per_bin <= 0.72 * neighbor_block_max_candidates()) break;
        int d = -1;
        for (int q = 0; q < 3; q++) if (bs[q] > 1 && (d < 0 || sd.mscan[q] < sd.mscan[d])) d = q;   // shrinking a thin-layer direction saves most
        if (d < 0) return 0;
        bs[d] = 1;
    }
    int n = 1;
    for (int d = 0; d < 3; d++) { sd.bs[d] = bs[d]; sd.nblk[d] = (sd.nbin[d] + bs[d] - 1) / bs[d]; n *= sd.nblk[d]; }
    return n;
}

static int set_structures_domain(gapcu_ctx *c, int natoms, const int *species, const double *lat_c, const double *pos, bool pos_soa, double rcut);

// pos_soa: pos is [3][ntot_of_that_structure] per structure (Fortran pos(NA,3)); else C order [ntot][3].
// need_weights = false for the bond-length path (no potential involved).
static int set_structures_impl(gapcu_ctx *c, int nstruct, const int *natoms, const int *species, const double *lat_c,
                               const double *pos, bool pos_soa, double rcut, bool need_weights) {
    DeviceGuard dg_(c->device);
    if (nstruct <= 0) return fail(GAPCU_EARG, "nstruct must be positive");
    if (!(rcut > 0.0)) return fail(GAPCU_EARG, "rcut must be positive");
    if (need_weights && !c->have_sf) return fail(GAPCU_EARG, "no potential loaded");
    long ntot = 0;
    for (int s = 0; s < nstruct; s++) {
        if (natoms[s] <= 0) return fail(GAPCU_EARG, "structure without atoms");
        ntot += natoms[s];
    }
    if (ntot > (1l << 30)) return fail(GAPCU_ELIMIT, "too many atoms");
    c->lists_valid = false; c->reuse_next = false;
    if (c->dom.enabled && need_weights) {
        if (nstruct != 1) return fail(GAPCU_EARG, "spatial decomposition works on a single structure");
        c->rcut = rcut;
        return set_structures_domain(c, natoms[0], species, lat_c, pos, pos_soa, rcut);
    }
    c->rcut = rcut;
    const double rskin = rcut + c->skin();
    c->h_structs.resize(nstruct);
    c->h_natoms.assign(natoms, natoms + nstruct);
    int boff = 0, aoff = 0;
    double max_density = 0.0;
    bool direct = true;   // all structures small enough for the direct neighbour kernel
    bool blocks_ok = true;   // every structure fits the block form of the list kernel
    int nblk_total = 0;
    for (int s = 0; s < nstruct; s++) {
        StructDev &sd = c->h_structs[s];
        int rc = fill_struct(sd, lat_c + 9 * (size_t)s, rcut, rskin, natoms[s], &direct);
        if (rc) return rc;
        sd.atom_off = aoff; sd.bin_off = boff;
        aoff += natoms[s]; boff += sd.nbins;
        max_density = std::max(max_density, natoms[s] / sd.volume);
        if (blocks_ok) {
            const int nb = plan_blocks(sd, natoms[s]);
            sd.blk_off = nblk_total;
            if (nb == 0) blocks_ok = false; else nblk_total += nb;
        }
    }
    c->nblocks = blocks_ok ? nblk_total : 0;
    c->nstruct = nstruct; c->ntot = (int)ntot; c->n_centres = (int)ntot; c->nbins = boff;
    c->direct_ok = direct && !getenv("GAPCU_NO_DIRECT");   // GAPCU_NO_DIRECT: always the cell list (A/B and tests)
    const size_t NT = (size_t)ntot;
    // ---- pack host staging: structs | sid | pos SoA | wgt
    size_t b_structs = sizeof(StructDev) * nstruct, b_wgt = sizeof(double) * NT;
    const InputLayout IL = input_layout((size_t)nstruct, NT, (size_t)c->nblocks);
    const size_t o_structs = IL.o_structs, o_sid = IL.o_sid, o_pos = IL.o_pos, o_wgt = IL.o_wgt, total = IL.total;
    // the staging buffer may still feed the previous call's copy if the caller never waited for it
    if (c->h2d_pending) { CU(cudaStreamSynchronize(c->stream)); c->h2d_pending = false; }
    if (c->pin(total)) return fail(GAPCU_ECUDA, "cudaMallocHost failed");
    char *hp = (char *)c->h_pin;
    // ---- device buffers (first: the inputs block tells which of its parts are still good)
    const bool was_valid = c->in_valid;
    const unsigned char *old_block = c->d_inputs.p;
    CU(ensure_inputs(c, (size_t)nstruct, NT, (size_t)c->nblocks));
    const bool block_kept = was_valid && c->d_inputs.p == old_block && c->in_nblocks == (size_t)c->nblocks &&
                            c->in_natoms.size() == (size_t)nstruct && !memcmp(c->in_natoms.data(), natoms, sizeof(int) * nstruct);
    const bool keep_sid = block_kept;   // structure ids and block owners depend on the atom counts (and block plan) only
    const bool keep_wgt = block_kept && need_weights && c->in_wgt_ok && c->in_sf_version == c->sf_version &&
                          c->in_species.size() == NT && !memcmp(c->in_species.data(), species, sizeof(int) * NT);
    memcpy(hp + o_structs, c->h_structs.data(), b_structs);
    int *h_sid = (int *)(hp + o_sid);
    double *h_pos = (double *)(hp + o_pos), *h_wgt = (double *)(hp + o_wgt);
    for (int s = 0, a = 0; s < nstruct; s++) {
        const int n = natoms[s];
        if (!keep_sid) for (int t = 0; t < n; t++) h_sid[a + t] = s;
        if (pos_soa) {
            for (int d = 0; d < 3; d++) memcpy(h_pos + d * NT + a, pos + 3 * (size_t)a + (size_t)d * n, sizeof(double) * n);
        } else {
            for (int t = 0; t < n; t++)
                for (int d = 0; d < 3; d++) h_pos[d * NT + a + t] = pos[3 * (size_t)(a + t) + d];
        }
        a += n;
    }
    if (!keep_sid) {   // structure of every block of cells
        int *h_blk = (int *)(hp + IL.o_blk);
        for (int s = 0; s < nstruct && c->nblocks; s++) {
            const StructDev &sd = c->h_structs[s];
            const int nb = sd.nblk[0] * sd.nblk[1] * sd.nblk[2];
            for (int k = 0; k < nb; k++) h_blk[sd.blk_off + k] = s;
        }
    }
    if (!keep_wgt) {
        if (need_weights) {
            for (size_t t = 0; t < NT; t++) { int rc = lookup_weight(c, species[t], &h_wgt[t]); if (rc) return rc; }
        } else {
            memset(h_wgt, 0, b_wgt);
        }
    }
    CU(c->d_abin.ensure(NT)); CU(c->d_sabin.ensure(NT)); CU(c->d_spos.ensure(3 * NT)); CU(c->d_arank.ensure(NT)); CU(c->d_bin_count.ensure(2 * (size_t)c->nbins + 2));
    CU(c->d_bin_start.ensure(c->nbins + 2)); CU(c->d_bin_atoms.ensure(NT)); CU(c->d_nbr_cnt.ensure(NT)); CU(c->d_skin_cnt.ensure(NT)); CU(c->d_order.ensure(NT));
    CU(ensure_results(c, (size_t)nstruct, NT));
    CU(c->d_mindis.ensure(NT)); CU(c->d_scan_sums.ensure((size_t)c->nbins / 4096 + 4));
    if (!keep_sid && !keep_wgt) {
        CU(cudaMemcpyAsync(c->d_inputs.p, hp, total, cudaMemcpyHostToDevice, c->stream));   // structs | sid | pos | wgt | blk in one copy
    } else {
        // only what changed travels: cell records and positions, and whichever of the other parts is stale
        CU(cudaMemcpyAsync(c->d_inputs.p + o_structs, hp + o_structs, b_structs, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_inputs.p + o_pos, hp + o_pos, sizeof(double) * 3 * NT, cudaMemcpyHostToDevice, c->stream));
        if (!keep_sid) {
            CU(cudaMemcpyAsync(c->d_inputs.p + o_sid, hp + o_sid, sizeof(int) * NT, cudaMemcpyHostToDevice, c->stream));
            if (c->nblocks) CU(cudaMemcpyAsync(c->d_inputs.p + IL.o_blk, hp + IL.o_blk, sizeof(int) * (size_t)c->nblocks, cudaMemcpyHostToDevice, c->stream));
        }
        if (!keep_wgt) CU(cudaMemcpyAsync(c->d_inputs.p + o_wgt, hp + o_wgt, b_wgt, cudaMemcpyHostToDevice, c->stream));
    }
    c->h2d_pending = true;
    c->in_valid = true;
    c->in_natoms.assign(natoms, natoms + nstruct); c->in_nblocks = (size_t)c->nblocks;
    if (need_weights) {
        if (!keep_wgt) c->in_species.assign(species, species + NT);
        c->in_wgt_ok = true; c->in_sf_version = c->sf_version;
    } else {
        c->in_wgt_ok = false;
    }
    // ---- neighbour capacity estimate (grown on demand)
    int est = (int)(4.18879 * rskin * rskin * rskin * max_density * 1.25) + 32;
    est = std::min(1024, std::max(64, round_up(est, 32)));
    if (c->cap < est) { c->cap = est; c->pcap_known = false; }
    if (c->last_ntot != c->ntot) { c->pcap_known = false; c->last_ntot = c->ntot; }
    c->computed = false;
    return 0;
}

extern "C" int gapcu_ctx_set_structures(gapcu_ctx *c, int nstruct, const int *natoms, const int *species,
                                        const double *lat, const double *pos, double rcut) {
    if (!c) return fail(GAPCU_EARG, "null context");
    return set_structures_impl(c, nstruct, natoms, species, lat, pos, false, rcut, true);
}

extern "C" int gapcu_ctx_set_skin(gapcu_ctx *c, double skin) {
    if (!c || !(skin >= 0.0)) return fail(GAPCU_EARG, "skin must be >= 0");
    c->skin_user = skin;
    c->lists_valid = false;
    c->pcap_known = false;
    return 0;
}

// ---------------------------------------------------------------------------
// NCCL, resolved at run time (the library has no link-time dependency on it)
// ---------------------------------------------------------------------------
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
} g_nccl;
constexpr int NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_FLOAT64 = 8;

int load_nccl() {
    if (g_nccl.ok) return 0;
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) if ((h = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return fail(GAPCU_ECUDA, std::string("cannot load NCCL: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(NcclId *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd)
        return fail(GAPCU_ECUDA, "NCCL symbols missing");
    g_nccl.ok = true;
    return 0;
}
int nccl_fail(const char *what, int r) {
    return fail(GAPCU_ECUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
}
}  // namespace

extern "C" int gapcu_nccl_unique_id(char *out128) {
    int rc = load_nccl();
    if (rc) return rc;
    NcclId id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r) return fail(GAPCU_ECUDA, "ncclGetUniqueId failed");
    memcpy(out128, id.internal, 128);
    return 0;
}

extern "C" int gapcu_ctx_nccl_init(gapcu_ctx *c, int nranks, int rank, const char *id128) {
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return fail(GAPCU_EARG, "bad NCCL rank/size");
    int rc = load_nccl();
    if (rc) return rc;
    DeviceGuard dg_(c->device);
    NcclId id;
    memcpy(id.internal, id128, 128);
    void *comm = nullptr;
    const int r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r) return nccl_fail("ncclCommInitRank", r);
    c->nccl_comm = comm;
    c->nccl_ranks = nranks;
    c->nccl_rank = rank;
    return 0;
}

// ---------------------------------------------------------------------------
// compute
// ---------------------------------------------------------------------------
static int ensure_work_buffers(gapcu_ctx *c) {
    const size_t NT = (size_t)c->ntot, NC = (size_t)c->n_centres;
    CU(c->d_hist.ensure(1024 + 2)); CU(c->d_scan_sums.ensure((size_t)c->nbins / 4096 + 4));
    CU(c->d_keys.ensure(NT * c->cap)); CU(c->d_skin_keys.ensure(NT * c->cap));
    CU(c->d_G.ensure(NC * c->D)); CU(c->d_dEdG.ensure(NC * c->D)); CU(c->d_eatom.ensure(NC));
    CU(c->d_fpair.ensure(NC * c->cap * 3)); CU(c->d_gself.ensure(NC * 3)); CU(c->d_vir.ensure(NC * 6));
    c->max_natoms = 0;
    for (const StructDev &sd : c->h_structs) c->max_natoms = std::max(c->max_natoms, sd.natoms);
    CU(c->d_finpart.ensure((size_t)std::max(1, c->nstruct) * finalize_chunks(c->max_natoms) * 8));
    return 0;
}

// small undecomposed cells take the direct neighbour kernel, which also orders the centres
static bool neighbors_direct(const gapcu_ctx *c) { return c->direct_ok && !c->dom.enabled; }

static NeighborBuild neighbor_args(gapcu_ctx *c, bool with_keys, bool with_min, bool with_order);

static int run_neighbors(gapcu_ctx *c, bool with_keys, bool with_min, bool with_order = false) {
    const NeighborBuild b = neighbor_args(c, with_keys, with_min, with_order);
    launch_neighbor_build(c->stream, b, &c->launches);
    CU(cudaGetLastError());
    return 0;
}

static int read_flags(gapcu_ctx *c) {
    CU(cudaMemcpyAsync(&c->h_flags, c->d_flags.p, sizeof(DevFlags), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->h2d_pending = false;
    return 0;
}

// kernel arguments of the centre kernel for the context's current state; picks the pipeline
// and the shared-memory budget (triplet-list capacity, private accumulator sets)
static int make_centre_args(gapcu_ctx *c, int lgrad, int pcap, CentreArgs *out, bool *fused_out, int cs = 1) {
    CentreArgs a;
    memset(&a, 0, sizeof a);
    a.cs = cs;
    a.plan = c->plan_dev();
    a.cls = c->class_tab();
    a.structs = c->d_structs.p; a.sid = c->d_sid.p; a.pos = c->d_pos.p; a.wgt = c->d_wgt.p;
    a.nbr_keys = c->d_keys.p; a.nbr_cnt = c->d_nbr_cnt.p; a.order = c->d_order.p; a.n_centres = &c->d_flags.p->n_centres; a.exp2_table = c->d_exp2.p;
    a.ntot = c->ntot; a.ncentres_max = c->n_centres; a.cap = c->cap; a.pcap = pcap; a.lgrad = lgrad;
    { static int var = -1; if (var < 0) { const char *e = getenv("GAPCU_VARIANT"); var = e ? atoi(e) : 0; } a.variant = var; }
    a.G = c->d_G.p; a.dEdG = c->d_dEdG.p; a.dEdG_out = c->d_dEdG.p; a.eatom = c->d_eatom.p;
    a.fpair = c->d_fpair.p; a.gself = c->d_gself.p; a.vir = c->d_vir.p;
    a.gpr_M = c->M; a.gpr_Mp = c->Mp; a.gpr_Dp = c->Dp; a.gpr_Mt = c->d_Mt.p; a.gpr_MtT = c->d_MtT.p;
    a.gpr_coeff = c->d_coeff.p; a.gpr_cmean = c->d_cmean.p; a.gpr_itheta = c->d_itheta.p;
    a.flags = c->d_flags.p;
    a.trip_out = c->dbg_trip; a.trip_cnt = c->dbg_trip_cnt; a.trip_cap = c->dbg_trip_cap;
    {   // parked exponentials (centre_impl.cuh MODE_FUSED_SE): every angular class carries the same 1 or 2 alphas
        const SfPlan &pl = c->plan;
        int first = -1, ng0 = 0;
        bool same = true;
        for (int k = 0; k < pl.ncls && same; k++) {
            const int g0 = pl.grp_begin[k], g1 = pl.grp_begin[k + 1];
            if (g1 == g0) continue;
            if (first < 0) { first = k; ng0 = g1 - g0; same = ng0 <= 2; continue; }
            same = (g1 - g0) == ng0;
            for (int g = 0; g < ng0 && same; g++) same = pl.grp_alpha[g0 + g] == pl.grp_alpha[pl.grp_begin[first] + g];
        }
        static const bool off = getenv("GAPCU_NO_SHARE_EXP") != nullptr;   // A/B switch
        a.share_exp = (first >= 0 && same && !off) ? ng0 : 0;   // 1 or 2: the number of shared exponents
        a.c_first = first < 0 ? 0 : first;
    }
    // The in-CTA GPR re-reads the sparse set once per atom: worth it while that set is
    // small (it stays in L1/L2 and a separate GEMM launch would be latency bound);
    // large sets go through the tiled DMMA kernel.
    const bool fused = c->pipeline == 2 || (c->pipeline == 0 && (size_t)c->Mp * c->Dp <= 64 * 1024);
    // shared-memory budget: triplet-list capacity and private accumulator sets.  Prefer a
    // footprint that lets 3 CTAs share an SM (the kernel is latency bound: more resident
    // warps matter more than building the triplet list in one chunk), then 2, then 1.
    {
        const int q = pcap * (pcap - 1) / 2;
        const int want = std::min(8192, std::max(2048, round_up(q, 32)));
        const int mode = fused ? 2 : 1;
        const size_t targets[3] = {(size_t)((a.variant & 8) ? 112 * 1024 : 74752), 110 * 1024, 220 * 1024};   // 3, 2, 1 CTAs per SM (the kernel's static tables take another 1.7 KB, the system 1 KB per CTA: 3 x (73 + 1.7 + 1) KB fits the 228 KB of an SM)
        const int lmin[3] = {3584, 3072, 1024};
        bool ok = false;
        for (int t = 0; t < 3 && !ok; t++)
            for (int pass = 0; pass < 2 && !ok; pass++) {
                if (pass == 1 && t < 2) continue;          // the shared accumulator set only as a last resort
                a.npa = pass == 0 ? centre_warps() : 1;
                for (a.lcap = want; a.lcap >= std::min(want, lmin[t]); a.lcap -= 512)
                    if (centre_smem_bytes(a, mode) <= targets[t]) { ok = true; break; }
            }
        if (!ok) return fail(GAPCU_ELIMIT, "centre kernel needs more shared memory than an SM has");
        // several chunks per centre: the forward pass parks its sorted lists for the backward pass
        const int chunks = (q + a.lcap - 1) / a.lcap;
        if (fused && chunks > 1 && chunks <= 16) {
            if (!c->sm_count) cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
            const int ctas = c->sm_count * 4;   // upper bound of the persistent grid
            CU(c->d_stash.ensure(centre_stash_words(a, chunks, ctas)));
            a.list_scratch = c->d_stash.p; a.list_scratch_chunks = chunks;
        }
        if (fused && a.share_exp && pcap <= 256) {
            if (!c->sm_count) cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
            a.estash_stride = std::max(1, chunks) * (a.lcap + 32);     // every chunk of a centre's list has its own slice
            CU(c->d_estash.ensure((size_t)c->sm_count * 4 * a.estash_stride));
            a.estash = c->d_estash.p;
        }
    }
    *out = a;
    *fused_out = fused;
    return 0;
}

// order -> centre kernel tiers -> (split pipeline: GPR, backward) -> force gather.  The neighbour
// lists of the context's atoms are in place; ev (optional) records stage boundaries 1..5.
static int run_centres_and_gather(gapcu_ctx *c, int lgrad, cudaEvent_t *ev) {
    int rc;
    if (!neighbors_direct(c) || c->last_reuse)
        launch_order(c->stream, c->d_nbr_cnt.p, c->n_centres, c->d_order.p, c->d_flags.p, c->d_hist.p, &c->launches);
    CU(cudaGetLastError());
    if (ev) CU(cudaEventRecord(ev[1], c->stream));
    // Capacity tiers: `order` lists the centres by descending neighbour count, so the centres
    // that need the 1024 / 512 / 256-neighbour instance of the kernel are a prefix; each tier is
    // served by its own instance (own shared-memory footprint, hence own residency): a few crowded
    // atoms of a heterogeneous batch no longer dictate the footprint of all the others.
    struct Tier { CentreArgs a; };
    std::vector<Tier> tiers;
    bool fused = false;
    {
        const int top = centre_pcap_template(c->pcap);
        DevFlags *F = c->d_flags.p;
        for (int cap = top; cap >= 128; cap >>= 1) {
            const int idx = cap == 128 ? 0 : cap == 256 ? 1 : cap == 512 ? 2 : 3;
            Tier t;
            // CTAs per centre (thread-block cluster of the fused kernel): a launch with fewer centres
            // than CTA slots spreads every centre over 2 or 4 CTAs.  The host knows the number of
            // centres of a tier only as an upper bound (the number of atoms); gapcu_ctx_set_cluster /
            // GAPCU_CLUSTER=1|2|4 overrides.
            int cs = 1;
            {
                const int forced = c->cluster;
                if (!c->sm_count) cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
                const int slots = c->sm_count * (cap == 128 ? 3 : cap == 256 ? 2 : 1);   // resident CTAs of this tier's instance
                if (forced == 1 || forced == 2 || forced == 4) cs = forced;
                // no tier can hold more centres than there are atoms.  Measured on B200 (tools/cluster_probe.py):
                // the split pays while the CTAs of a launch do not have to share SMs more than two at a time
                else if (c->n_centres * 4 <= c->sm_count) cs = 4;
                else if (c->n_centres <= c->sm_count && c->n_centres * 2 <= slots) cs = 2;
            }
            if ((rc = make_centre_args(c, lgrad, cap == top ? c->pcap : cap, &t.a, &fused, cs))) return rc;
            t.a.q_begin = cap == top ? nullptr : &F->n_gt[idx];
            t.a.q_end = cap == 128 ? &F->n_centres : &F->n_gt[idx - 1];
            tiers.push_back(t);
            // a launch with at most one centre per SM gains nothing from the leaner instances: the top
            // one serves every centre and the (mostly empty) lower tiers are not launched
            if (cap == top && top <= 256 && c->n_centres <= c->sm_count) { tiers.back().a.q_end = &F->n_centres; break; }
        }
        for (Tier &t : tiers) {  // a later tier may have grown (moved) the shared parking buffers
            if (t.a.list_scratch) t.a.list_scratch = c->d_stash.p;
            if (t.a.estash) t.a.estash = c->d_estash.p;
        }
    }
    const char *too_big = "centre kernel needs more shared memory than an SM has";
    if (fused) {
        for (Tier &t : tiers) {
            if (launch_fused(c->stream, t.a, &c->launches)) return fail(GAPCU_ELIMIT, too_big);
            CU(cudaGetLastError());
        }
        if (ev) { CU(cudaEventRecord(ev[2], c->stream)); CU(cudaEventRecord(ev[3], c->stream)); CU(cudaEventRecord(ev[4], c->stream)); }
    } else {
        for (Tier &t : tiers) {
            if (launch_forward(c->stream, t.a, &c->launches)) return fail(GAPCU_ELIMIT, too_big);
            CU(cudaGetLastError());
        }
        if (ev) CU(cudaEventRecord(ev[2], c->stream));
        GprDev g;
        g.M = c->M; g.Mp = c->Mp; g.D = c->D; g.Dp = c->Dp; g.Mt = c->d_Mt.p; g.MtT = c->d_MtT.p; g.mn = c->d_mn.p;
        g.coeff = c->d_coeff.p; g.cmean = c->d_cmean.p; g.itheta = c->d_itheta.p;
        const int max_slices = gpr_max_slices(c->n_centres, c->Mp);
        CU(c->d_epart.ensure((size_t)max_slices * c->n_centres));
        CU(c->d_accpart.ensure((size_t)max_slices * c->n_centres * c->Dp));
        if (launch_gpr(c->stream, g, c->d_G.p, c->n_centres, c->d_eatom.p, c->d_dEdG.p, c->d_epart.p, c->d_accpart.p,
                       max_slices, c->d_exp2.p, &c->launches))
            return fail(GAPCU_ELIMIT, "unsupported descriptor length for the GPR kernel");
        CU(cudaGetLastError());
        if (ev) CU(cudaEventRecord(ev[3], c->stream));
        if (lgrad)
            for (Tier &t : tiers) {
                if (launch_backward(c->stream, t.a, &c->launches)) return fail(GAPCU_ELIMIT, too_big);
                CU(cudaGetLastError());
            }
        if (ev) CU(cudaEventRecord(ev[4], c->stream));
    }
    return 0;
}

static int enqueue_pass(gapcu_ctx *c, int lgrad, cudaEvent_t *ev, bool reuse);
#include "domain_host.inc"

static NeighborBuild neighbor_args(gapcu_ctx *c, bool with_keys, bool with_min, bool with_order) {
    NeighborBuild b;
    memset(&b, 0, sizeof b);
    b.structs = c->d_structs.p; b.sid = c->d_sid.p; b.pos = c->d_pos.p; b.ntot = c->ntot; b.nbins_total = c->nbins;
    b.rcut = c->rcut; b.rskin = with_keys ? c->rcut + c->skin() : c->rcut; b.cap = c->cap;
    b.arank_scratch = c->d_scan_sums.p;
    b.abin = c->d_abin.p; b.arank = c->d_arank.p; b.bin_count = c->d_bin_count.p; b.bin_start = c->d_bin_start.p;
    b.bin_atoms = c->d_bin_atoms.p; b.sabin = c->d_sabin.p; b.spos = c->d_spos.p;
    b.skin_keys = with_keys ? c->d_skin_keys.p : nullptr; b.skin_cnt = c->d_skin_cnt.p;
    b.nbr_keys = with_keys ? c->d_keys.p : nullptr; b.nbr_cnt = c->d_nbr_cnt.p;
    b.min_dis = with_min ? c->d_mindis.p : nullptr;
    b.flags = c->d_flags.p;
    b.order = with_order ? c->d_order.p : nullptr;
    b.direct = neighbors_direct(c);
    b.n_own = c->n_centres;
    if (c->dom.enabled && c->ds) { b.sft = c->ds->d_sft.p; b.nloc = &c->d_flags.p->n_loc; }
    static const bool legacy_env = getenv("GAPCU_K1_LEGACY") != nullptr;   // A/B switch: always one CTA per centre
    b.nblocks = (c->k1_legacy || legacy_env) ? 0 : c->nblocks;
    // the block form pays from a few blocks per SM on; below, one CTA per centre keeps more of the device busy
    if (!c->sm_count) cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
    if (b.nblocks < 2 * c->sm_count) b.nblocks = 0;
    b.t2skin = sqrt_threshold(b.rskin); b.t2cut = sqrt_threshold(b.rcut);
    b.t2close = std::nextafter(sqrt_threshold(0.5), 0.0);     // sqrt(x) < 0.5: one step below the largest x with sqrt(x) <= 0.5 ...
    while (std::sqrt(b.t2close) >= 0.5) b.t2close = std::nextafter(b.t2close, 0.0);   // ... made exact
    b.blk_struct = c->nstruct > 1 ? c->d_blk : nullptr;
    return b;
}

// enqueue one full pass; ev (optional) = GAPCU_NSTAGE+1 events recorded at stage boundaries.
// reuse: keep the skin lists, only re-filter them against the current positions (Verlet reuse).
static int enqueue_pass(gapcu_ctx *c, int lgrad, cudaEvent_t *ev, bool reuse) {
    if (!c->have_sf || !c->have_gpr) return fail(GAPCU_EARG, "no potential loaded");
    if (c->D != c->plan.D) return fail(GAPCU_EARG, "des_len of the GPR data does not match 2*nsf of the SF table");
    if (c->ntot <= 0) return fail(GAPCU_EARG, "no structures set");
    if (c->dom.enabled) return domain_enqueue_pass(c, lgrad, ev, reuse);
    reuse = reuse && c->lists_valid;
    int rc = ensure_work_buffers(c);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->d_flags.p, 0, sizeof(DevFlags), c->stream));
    if (ev) CU(cudaEventRecord(ev[0], c->stream));
    c->last_reuse = reuse;
    if (reuse) {
        launch_refilter(c->stream, neighbor_args(c, true, false, false), c->d_pos_build.p, c->skin(), &c->launches);
        CU(cudaGetLastError());
    } else {
        if ((rc = run_neighbors(c, true, false, true))) return rc;
        if (!c->pcap_known) {
            // first pass for this kind of input: learn the largest list lengths
            for (int attempt = 0; attempt < 4; attempt++) {
                if ((rc = read_flags(c))) return rc;
                if (c->h_flags.too_many)
                    return fail(GAPCU_ENEIGH, "Atoms neighbor: " + std::to_string(c->h_flags.maxcount) +
                                                  " large than max_neighbor 1000");
                if (c->h_flags.blk_overflow) c->k1_legacy = true;
                if (!c->h_flags.overflow && !c->h_flags.blk_overflow) break;
                if (c->h_flags.maxskin > 1024) return fail(GAPCU_ELIMIT, "more than 1024 atoms within rcut + skin: reduce the skin");
                if (c->h_flags.overflow) c->cap = std::min(1024, round_up(c->h_flags.maxskin + 16, 32));
                if ((rc = ensure_work_buffers(c))) return rc;
                CU(cudaMemsetAsync(c->d_flags.p, 0, sizeof(DevFlags), c->stream));
                if ((rc = run_neighbors(c, true, false, true))) return rc;
            }
            c->pcap = std::min(c->cap, std::max(32, round_up(c->h_flags.maxcount + 8, 32)));
            c->pcap_known = true;
        }
        if (c->skin_user > 0.0) {   // a caller that set a skin will come back with moved positions
            CU(c->d_pos_build.ensure(3 * (size_t)c->ntot));
            CU(cudaMemcpyAsync(c->d_pos_build.p, c->d_pos.p, sizeof(double) * 3 * (size_t)c->ntot, cudaMemcpyDeviceToDevice, c->stream));
            c->lists_valid = true;
        }
    }
    if ((rc = run_centres_and_gather(c, lgrad, ev))) return rc;
    GatherArgs g;
    memset(&g, 0, sizeof g);
    g.structs = c->d_structs.p; g.nstruct = c->nstruct; g.sid = c->d_sid.p; g.ntot = c->ntot; g.cap = c->cap;
    g.skin_keys = c->d_skin_keys.p; g.skin_cnt = c->d_skin_cnt.p; g.nbr_keys = c->d_keys.p; g.nbr_cnt = c->d_nbr_cnt.p;
    g.fpair = c->d_fpair.p; g.gself = c->d_gself.p; g.vir = c->d_vir.p; g.eatom = c->d_eatom.p; g.lgrad = lgrad;
    g.force_soa = c->d_force.p; g.out8 = c->d_out8.p; g.partial = c->d_finpart.p; g.max_natoms = c->max_natoms;
    g.n_own = c->ntot;
    launch_gather(c->stream, g, &c->launches);
    CU(cudaGetLastError());
    if (ev) CU(cudaEventRecord(ev[5], c->stream));
    c->last_lgrad = lgrad;
    c->computed = true;
    return 0;
}

extern "C" int gapcu_ctx_compute(gapcu_ctx *c, int lgrad) {
    if (!c) return fail(GAPCU_EARG, "null context");
    if (c->dom.enabled && c->ds && c->ds->group) return fail(GAPCU_EARG, "this context belongs to a group: use gapcu_group_compute");
    DeviceGuard dg_(c->device);
    const bool reuse = c->reuse_next;
    c->reuse_next = false;
    return enqueue_pass(c, lgrad, nullptr, reuse);
}

// New positions for the resident structures (same atoms, same cells): an MD or relaxation step.
// pos: C order [ntot][3] (decomposed runs: this rank's owned atoms in the order of gapcu_ctx_owned).
// reuse_lists != 0 and a skin > 0 set with gapcu_ctx_set_skin: the next compute keeps the skin lists and
// only re-filters them (rebuilding on its own when an atom has moved more than skin/2).
extern "C" int gapcu_ctx_update_positions(gapcu_ctx *c, const double *pos, int reuse_lists) {
    if (!c || !pos) return fail(GAPCU_EARG, "null argument");
    if (c->ntot <= 0) return fail(GAPCU_EARG, "no structures set");
    DeviceGuard dg_(c->device);
    if (c->dom.enabled) return domain_update_positions(c, pos, reuse_lists);
    const size_t NT = (size_t)c->ntot;
    if (c->h2d_pending) { CU(cudaStreamSynchronize(c->stream)); c->h2d_pending = false; }
    if (c->pin(sizeof(double) * 3 * NT)) return fail(GAPCU_ECUDA, "cudaMallocHost failed");
    double *hp = (double *)c->h_pin;
    for (size_t t = 0; t < NT; t++) { hp[t] = pos[3 * t]; hp[NT + t] = pos[3 * t + 1]; hp[2 * NT + t] = pos[3 * t + 2]; }
    CU(cudaMemcpyAsync(c->d_pos.p, hp, sizeof(double) * 3 * NT, cudaMemcpyHostToDevice, c->stream));
    c->h2d_pending = true;
    c->reuse_next = reuse_lists && c->skin_user > 0.0 && c->lists_valid;
    c->computed = false;
    return 0;
}

// After the results block came back: decide what the flags ask for.  Returns 0 = results are good,
// 1 = the pass was re-enqueued (read again), < 0 = error.  In a decomposed run the flags are the
// maxima over all ranks (halo.cu:k_halo_combine), so every rank takes the same branch here and the
// re-run's collectives pair up.
static int react_to_flags(gapcu_ctx *c, int attempt) {
    const DevFlags &f = c->h_flags;
    if (f.too_many)
        return fail(GAPCU_ENEIGH, "Atoms neighbor: " + std::to_string(f.maxcount) + " large than max_neighbor 1000");
    if (f.halo_far)
        return fail(GAPCU_EDOMAIN, "an owned atom left its brick by more than the drift allowance: set the structure again to re-partition");
    bool rerun = false, reuse = false;
    if (f.overflow) {
        if (f.maxskin > 1024) return fail(GAPCU_ELIMIT, "more than 1024 atoms within rcut + skin: reduce the skin");
        c->cap = std::max(c->cap, std::min(1024, round_up(f.maxskin + 16, 32)));
        c->pcap_known = false; c->lists_valid = false;
        rerun = true;
    }
    if (f.halo_overflow && c->ds) { domain_forget_caps(c); c->lists_valid = false; rerun = true; }
    if (f.blk_overflow) { c->k1_legacy = true; c->lists_valid = false; rerun = true; }
    if (f.stale) { c->lists_valid = false; rerun = true; }
    if (c->dom.enabled && !c->pcap_known && !rerun) {
        // first decomposed pass ran with the list capacity as kernel capacity: remember the real one
        c->pcap = std::min(c->cap, std::max(32, round_up(f.maxcount + 8, 32)));
        c->pcap_known = true;
    }
    if (!rerun) return 0;
    if (attempt >= 3) return fail(GAPCU_ECUDA, "capacities did not converge");
    int rc = enqueue_pass(c, c->last_lgrad, nullptr, reuse);
    return rc ? rc : 1;
}

// wait for the pass; if a flag asks for it (a list outgrew its capacity, the skin lists went stale)
// enlarge / rebuild and run again.
static int finish_pass(gapcu_ctx *c) {
    for (int attempt = 0;; attempt++) {
        int rc = read_flags(c);
        if (rc) return rc;
        rc = react_to_flags(c, attempt);
        if (rc <= 0) return rc;
    }
}

extern "C" int gapcu_ctx_fetch(gapcu_ctx *c, double *ene, double *force, double *stress) {
    if (!c) return fail(GAPCU_EARG, "null context");
    if (!c->computed) return fail(GAPCU_EARG, "nothing computed");
    DeviceGuard dg_(c->device);
    // flags, per-structure outputs and forces come back in one copy; if a flag asks for another pass
    // (capacity outgrown, stale skin lists) it is run and read again
    const size_t NT = (size_t)c->ntot, NF = (size_t)c->n_centres;   // NF: atoms whose forces this context returns
    const size_t b_f = sizeof(double) * 3 * NT;
    double *h_out = nullptr, *h_f = nullptr;
    for (int attempt = 0;; attempt++) {
        const size_t bytes = force ? c->res_o_f + b_f : c->res_o_f;
        if (c->pin(c->res_o_f + b_f)) return fail(GAPCU_ECUDA, "cudaMallocHost failed");
        if (force && NF < NT) {
            // decomposed run: the owned atoms are the first NF of NT local points in each SoA row
            CU(cudaMemcpyAsync(c->h_pin, c->d_results.p, c->res_o_f, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaMemcpy2DAsync((char *)c->h_pin + c->res_o_f, sizeof(double) * NF, c->d_force.p, sizeof(double) * NT,
                                 sizeof(double) * NF, 3, cudaMemcpyDeviceToHost, c->stream));
        } else {
            CU(cudaMemcpyAsync(c->h_pin, c->d_results.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));
        c->h2d_pending = false;
        c->h_flags = *(const DevFlags *)c->h_pin;
        h_out = (double *)((char *)c->h_pin + c->res_o_out); h_f = (double *)((char *)c->h_pin + c->res_o_f);
        const int rc = react_to_flags(c, attempt);
        if (rc < 0) return rc;
        if (rc == 0) break;
    }
    if (c->h_flags.close_pairs)
        fprintf(stdout, " Warning: The distance of two atoms is very small (%d pairs below 0.5)\n", c->h_flags.close_pairs);
    for (int s = 0; s < c->nstruct; s++) {
        if (ene) ene[s] = h_out[8 * s];
        if (stress) for (int q = 0; q < 6; q++) stress[6 * s + q] = h_out[8 * s + 1 + q];
    }
    if (force) {
        const size_t ld = NF < NT ? NF : NT;
        for (size_t t = 0; t < NF; t++)
            for (int d = 0; d < 3; d++) force[3 * t + d] = h_f[d * ld + t];
    }
    return 0;
}

extern "C" int gapcu_ctx_fetch_descriptors(gapcu_ctx *c, double *xx, double *dedg, double *eatom) {
    if (!c || !c->computed) return fail(GAPCU_EARG, "nothing computed");
    DeviceGuard dg_(c->device);
    int rc = finish_pass(c);
    if (rc) return rc;
    const size_t n = (size_t)c->ntot * c->D;
    if (xx) CU(cudaMemcpyAsync(xx, c->d_G.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (dedg) CU(cudaMemcpyAsync(dedg, c->d_dEdG.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (eatom) CU(cudaMemcpyAsync(eatom, c->d_eatom.p, sizeof(double) * c->ntot, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gapcu_ctx_variance(gapcu_ctx *c, const double *qmm, double *variance, double *covf) {
    if (!c || !c->computed) return fail(GAPCU_EARG, "nothing computed");
    if (!qmm || !variance) return fail(GAPCU_EARG, "qmm and variance are required");
    DeviceGuard dg_(c->device);
    int rc = finish_pass(c);
    if (rc) return rc;
    const size_t M = (size_t)c->M, NT = (size_t)c->ntot;
    DBuf<double> d_q, d_covf, d_var;
    auto bail = [&](int r) { d_q.release(); d_covf.release(); d_var.release(); return r; };
    if (d_q.ensure(M * M) != cudaSuccess || d_covf.ensure(NT) != cudaSuccess || d_var.ensure(c->nstruct) != cudaSuccess)
        return bail(fail(GAPCU_ECUDA, "out of device memory for the variance pass"));
    if (cudaMemcpyAsync(d_q.p, qmm, sizeof(double) * M * M, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        return bail(fail(GAPCU_ECUDA, "upload of QMM failed"));
    if (launch_variance(c->stream, c->d_structs.p, c->nstruct, c->ntot, c->d_G.p, c->D, c->M, c->Mp, c->Dp, c->d_Mt.p,
                        c->d_cmean.p, c->d_itheta.p, d_q.p, d_covf.p, d_var.p))
        return bail(fail(GAPCU_ELIMIT, "sparse set too large for the variance kernel's shared memory"));
    c->launches += 2;
    std::vector<double> h_var(c->nstruct);
    bool ok = cudaMemcpyAsync(h_var.data(), d_var.p, sizeof(double) * c->nstruct, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    if (ok && covf) ok = cudaMemcpyAsync(covf, d_covf.p, sizeof(double) * NT, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess && cudaGetLastError() == cudaSuccess;
    if (!ok) return bail(fail(GAPCU_ECUDA, "variance pass failed"));
    for (int s = 0; s < c->nstruct; s++) variance[s] = h_var[s];
    return bail(0);
}

extern "C" int gapcu_ctx_fetch_neighbors(gapcu_ctx *c, int cap, int *count, int *idx, int *shift, double *dis) {
    if (!c || !c->computed) return fail(GAPCU_EARG, "nothing computed");
    DeviceGuard dg_(c->device);
    int rc = finish_pass(c);
    if (rc) return rc;
    const size_t NT = (size_t)c->ntot;
    std::vector<int> h_cnt(NT);
    std::vector<uint64_t> h_keys(NT * c->cap);
    std::vector<double> h_pos(3 * NT);
    CU(cudaMemcpy(h_cnt.data(), c->d_nbr_cnt.p, sizeof(int) * NT, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_keys.data(), c->d_keys.p, sizeof(uint64_t) * NT * c->cap, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_pos.data(), c->d_pos.p, sizeof(double) * 3 * NT, cudaMemcpyDeviceToHost));
    int mx = 0, sidx = 0;
    for (size_t i = 0; i < NT; i++) {
        while (sidx + 1 < c->nstruct && (int)i >= c->h_structs[sidx + 1].atom_off) sidx++;
        const StructDev &sd = c->h_structs[sidx];
        count[i] = h_cnt[i];
        mx = std::max(mx, h_cnt[i]);
        if (h_cnt[i] > cap) return fail(GAPCU_EARG, "fetch_neighbors: cap too small");
        for (int s = 0; s < h_cnt[i]; s++) {
            int j, n1, n2, n3;
            nbr_unkey(h_keys[i * c->cap + s], j, n1, n2, n3);
            const size_t q = i * cap + s;
            idx[q] = j; shift[3 * q] = n1; shift[3 * q + 1] = n2; shift[3 * q + 2] = n3;
            if (dis) {  // the reference arithmetic again (host code is built with -ffp-contract=off)
                const size_t jg = (size_t)sd.atom_off + j;
                double d2 = 0.0;
                for (int d = 0; d < 3; d++) {
                    double x = h_pos[d * NT + jg] + (double)n1 * sd.lat[d];
                    x = x + (double)n2 * sd.lat[3 + d];
                    x = x + (double)n3 * sd.lat[6 + d];
                    const double dr = h_pos[d * NT + i] - x;
                    d2 = d2 + dr * dr;
                }
                dis[q] = std::sqrt(d2);
            }
        }
    }
    return mx;
}

// Debug export of the neighbour PAIRS the angular functions are summed over (the reference's
// "k_neighbor > j_neighbor" loops with their three cutoff tests, wacsf.f90:177-244): for every atom the
// kept pairs exactly as the centre kernel's passes consume them, items[atom][k] = slot_j | slot_k << 10 |
// nclasses << 20 (slots index the atom's neighbour list in reference order; the pair belongs to the
// cutoff classes 0 .. nclasses-1, classes = distinct SF cutoffs in descending order).  Runs one pass of
// the split pipeline.  count[ntot], items[ntot][cap]; returns the largest count.
extern "C" int gapcu_ctx_debug_triplets(gapcu_ctx *c, int cap, int *count, unsigned *items) {
    if (!c || cap <= 0 || !count || !items) return fail(GAPCU_EARG, "bad arguments");
    if (c->ntot <= 0 || c->dom.enabled) return fail(GAPCU_EARG, "set an undecomposed structure first");
    DeviceGuard dg_(c->device);
    const size_t NT = (size_t)c->ntot;
    DBuf<uint32_t> d_items; DBuf<int> d_cnt;
    auto bail = [&](int r) { d_items.release(); d_cnt.release(); c->dbg_trip = nullptr; c->dbg_trip_cnt = nullptr; c->dbg_trip_cap = 0; return r; };
    if (d_items.ensure(NT * cap) != cudaSuccess || d_cnt.ensure(NT) != cudaSuccess) return bail(fail(GAPCU_ECUDA, "out of device memory"));
    if (cudaMemsetAsync(d_cnt.p, 0, sizeof(int) * NT, c->stream) != cudaSuccess) return bail(fail(GAPCU_ECUDA, "memset failed"));
    const int pipe = c->pipeline, clus = c->cluster;
    c->pipeline = 1; c->cluster = 1;
    c->dbg_trip = d_items.p; c->dbg_trip_cnt = d_cnt.p; c->dbg_trip_cap = cap;
    int rc = enqueue_pass(c, 1, nullptr, false);
    if (!rc) rc = finish_pass(c);
    c->pipeline = pipe; c->cluster = clus;
    if (rc) return bail(rc);
    if (cudaMemcpy(count, d_cnt.p, sizeof(int) * NT, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(items, d_items.p, sizeof(uint32_t) * NT * cap, cudaMemcpyDeviceToHost) != cudaSuccess)
        return bail(fail(GAPCU_ECUDA, "download failed"));
    int mx = 0;
    for (size_t i = 0; i < NT; i++) mx = std::max(mx, count[i]);
    return bail(mx);
}

extern "C" int gapcu_ctx_balance(gapcu_ctx *c, double *out4) {
    if (!c || !c->computed) return fail(GAPCU_EARG, "nothing computed");
    DeviceGuard dg_(c->device);
    int rc = read_flags(c);
    if (rc) return rc;
    const DevFlags &f = c->h_flags;
    out4[0] = (double)f.n_ctas;
    const unsigned long long tstart = ~f.t_start_min, tfirst = ~f.t_exit_min;     // stored as complements
    out4[1] = (double)(f.t_exit_max - tstart) * 1e-3;                              // kernel span, us
    out4[2] = (double)(tfirst - tstart) * 1e-3;                                    // first CTA done, us
    out4[3] = f.n_ctas ? (double)f.t_busy_sum / ((double)f.n_ctas * (double)(f.t_exit_max - tstart)) : 0.0;
    return 0;
}

extern "C" int gapcu_ctx_work_counters(gapcu_ctx *c, double *out, int n) {
    if (!c || !c->computed) return fail(GAPCU_EARG, "nothing computed");
    DeviceGuard dg_(c->device);
    int rc = read_flags(c);
    if (rc) return rc;
    for (int q = 0; q < n && q < 10; q++) out[q] = (double)c->h_flags.work[q];
    for (int q = 10; q < n && q < 26; q++) out[q] = (double)c->h_flags.phase_cycles[q - 10];   // GAPCU_VARIANT & 16
    return 0;
}

// ---------------------------------------------------------------------------
// timing
// ---------------------------------------------------------------------------
extern "C" const char *gapcu_stage_name(int s) {
    static const char *names[GAPCU_NSTAGE] = {"neighbor_build", "descriptor_forward", "gpr_dmma", "descriptor_backward",
                                              "force_gather_reduce", "halo_forward", "halo_return", ""};
    return (s >= 0 && s < GAPCU_NSTAGE) ? names[s] : "";
}

extern "C" int gapcu_ctx_time_compute(gapcu_ctx *c, int lgrad, int steps, long l2_flush_bytes, double *ms_total,
                                      double *stage_ms, long *launches) {
    if (!c) return fail(GAPCU_EARG, "null context");
    DeviceGuard dg_(c->device);
    if (!c->stage_ev_init) {
        for (auto &e : c->stage_ev) CU(cudaEventCreate(&e));
        c->stage_ev_init = true;
    }
    if (l2_flush_bytes > 0) CU(c->d_flush.ensure((size_t)l2_flush_bytes));
    // after gapcu_ctx_update_positions(.., reuse_lists = 1) the passes of this call re-filter the kept
    // skin lists; otherwise every pass rebuilds the neighbour lists (and, decomposed, the halo lists)
    const bool reuse = c->reuse_next;
    c->reuse_next = false;
    // make sure capacities are settled before timing
    int rc = enqueue_pass(c, lgrad, nullptr, reuse);
    if (rc) return rc;
    if ((rc = finish_pass(c))) return rc;
    std::vector<cudaEvent_t> ev(2 * (size_t)steps);
    for (auto &e : ev) CU(cudaEventCreate(&e));
    const long l0 = c->launches;
    for (int s = 0; s < steps; s++) {
        if (l2_flush_bytes > 0) CU(cudaMemsetAsync(c->d_flush.p, s & 0xff, (size_t)l2_flush_bytes, c->stream));
        CU(cudaEventRecord(ev[2 * s], c->stream));
        if ((rc = enqueue_pass(c, lgrad, nullptr, reuse))) return rc;
        CU(cudaEventRecord(ev[2 * s + 1], c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    if (launches) *launches = c->launches - l0;
    double tot = 0.0;
    for (int s = 0; s < steps; s++) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ev[2 * s], ev[2 * s + 1]));
        tot += ms;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    if (ms_total) *ms_total = tot;
    if ((rc = finish_pass(c))) return rc;
    if (stage_ms) {
        for (int q = 0; q < GAPCU_NSTAGE; q++) stage_ms[q] = 0.0;
        for (int s = 0; s < steps; s++) {
            if (l2_flush_bytes > 0) CU(cudaMemsetAsync(c->d_flush.p, s & 0xff, (size_t)l2_flush_bytes, c->stream));
            if ((rc = enqueue_pass(c, lgrad, c->stage_ev, reuse))) return rc;
            CU(cudaStreamSynchronize(c->stream));
            for (int q = 0; q < 5; q++) {
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, c->stage_ev[q], c->stage_ev[q + 1]));
                stage_ms[q] += ms;
            }
            if (c->dom.enabled) {
                // decomposed run: the exchange steps are carved out of the first and the last stage
                float fwd = 0.f, ret = 0.f;
                CU(cudaEventElapsedTime(&fwd, c->stage_ev[0], c->stage_ev[6]));
                CU(cudaEventElapsedTime(&ret, c->stage_ev[7], c->stage_ev[5]));
                stage_ms[5] += fwd; stage_ms[6] += ret;
                stage_ms[0] -= fwd; stage_ms[4] -= ret;
            }
        }
    }
    return 0;
}

extern "C" int gapcu_fp64_peaks(gapcu_ctx *c, double *dfma, double *dmma) {
    if (!c) return fail(GAPCU_EARG, "null context");
    DeviceGuard dg_(c->device);
    launch_fp64_peaks(c->stream, dfma, dmma);
    CU(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// Fortran-layout entry points on a process-wide default context
// ---------------------------------------------------------------------------
namespace {
std::mutex g_mu;
gapcu_ctx *g_ctx = nullptr;
struct FileId { dev_t dev = 0; ino_t ino = 0; off_t size = -1; long mt_s = 0, mt_ns = 0; } g_file;

// devices of the drop-in entry points (gapcu_set_devices); empty: GAPCU_DEVICE or device 0
std::vector<int> g_devices;
// gapcu_calc_batch: one context per device, each with the whole ./gap_parameters loaded
struct BatchSlot { int device = -1; gapcu_ctx *ctx = nullptr; FileId file; };
std::vector<BatchSlot> g_batch;

int default_device() {
    if (!g_devices.empty()) return g_devices[0];
    if (const char *e = getenv("GAPCU_DEVICE")) return atoi(e);
    return 0;
}

int default_ctx(gapcu_ctx **out) {
    if (!g_ctx) {
        const int dev = default_device();
        g_ctx = gapcu_ctx_create(dev);
        if (!g_ctx) return g_err.find("no CUDA device") != std::string::npos ? GAPCU_ENODEV : GAPCU_ECUDA;
    }
    *out = g_ctx;
    return 0;
}

// (re)load species weights + SF table from ./gap_parameters when the file changed
int refresh_sf_from_cwd(gapcu_ctx *c) {
    struct stat st;
    if (stat("gap_parameters", &st) != 0) return fail(GAPCU_EFILE, "gap_parameters file does not exist!");
    if (c->have_sf && st.st_dev == g_file.dev && st.st_ino == g_file.ino && st.st_size == g_file.size &&
        st.st_mtim.tv_sec == g_file.mt_s && st.st_mtim.tv_nsec == g_file.mt_ns)
        return 0;
    PotentialFile pf;
    try {
        pf = read_gap_parameters("gap_parameters");
    } catch (const std::exception &e) {
        return fail(GAPCU_EFILE, e.what());
    }
    int rc = set_sf(c, pf.z, pf.w, pf.ntype, pf.alpha, pf.cutoff);
    if (rc) return rc;
    g_file.dev = st.st_dev; g_file.ino = st.st_ino; g_file.size = st.st_size;
    g_file.mt_s = st.st_mtim.tv_sec; g_file.mt_ns = st.st_mtim.tv_nsec;
    return 0;
}
}  // namespace

// FGAP_CALC on several devices: with more than one device set (gapcu_set_devices) a large structure is cut
// into one brick per device and evaluated by an in-process group (domain_host.inc) -- same arguments, same
// results, the caller does not change.  Below GAPCU_GROUP_MIN_ATOMS atoms (default 20000) a single device is
// faster and is used.
namespace {
gapcu_group *g_group = nullptr;
std::vector<int> g_group_devs;
FileId g_group_file;
}  // namespace

static int calc_on_group(int na, const int *species, const double *lat, const double *pos, int nsparsex, int des_len,
                         const double *theta, const double *mm, const double *coeff, double rcut, int lgrad, double *ene,
                         double *force, double *stress) {
    if (!g_group || g_group_devs != g_devices) {
        if (g_group) gapcu_group_destroy(g_group);
        g_group = gapcu_group_create((int)g_devices.size(), g_devices.data());
        if (!g_group) return GAPCU_ECUDA;
        g_group_devs = g_devices;
        g_group_file = FileId();
    }
    struct stat st;
    if (stat("gap_parameters", &st) != 0) return fail(GAPCU_EFILE, "gap_parameters file does not exist!");
    const bool same_file = st.st_dev == g_group_file.dev && st.st_ino == g_group_file.ino && st.st_size == g_group_file.size &&
                           st.st_mtim.tv_sec == g_group_file.mt_s && st.st_mtim.tv_nsec == g_group_file.mt_ns;
    int rc;
    std::vector<double> mm_c;
    for (gapcu_ctx *c : g_group->ctx) {
        DeviceGuard dg_(c->device);
        if (!same_file || !c->have_sf) {
            PotentialFile pf;
            try { pf = read_gap_parameters("gap_parameters"); } catch (const std::exception &e) { return fail(GAPCU_EFILE, e.what()); }
            if ((rc = set_sf(c, pf.z, pf.w, pf.ntype, pf.alpha, pf.cutoff))) return rc;
        }
        if (des_len != c->plan.D) return fail(GAPCU_EARG, "des_len does not equal 2*nsf of ./gap_parameters");
        if (mm_c.empty()) {
            mm_c.resize((size_t)nsparsex * des_len);
            for (int k = 0; k < des_len; k++)
                for (int s = 0; s < nsparsex; s++) mm_c[(size_t)s * des_len + k] = mm[s + (size_t)nsparsex * k];
        }
        if ((rc = set_gpr(c, nsparsex, des_len, theta, mm_c.data(), coeff))) return rc;   // keeps the device copy when unchanged
    }
    g_group_file.dev = st.st_dev; g_group_file.ino = st.st_ino; g_group_file.size = st.st_size;
    g_group_file.mt_s = st.st_mtim.tv_sec; g_group_file.mt_ns = st.st_mtim.tv_nsec;
    double lat_c[9];
    for (int r = 0; r < 3; r++) for (int col = 0; col < 3; col++) lat_c[r * 3 + col] = lat[r + 3 * col];
    std::vector<double> pos_c(3 * (size_t)na), f_c(3 * (size_t)na);
    for (int t = 0; t < na; t++) for (int d = 0; d < 3; d++) pos_c[3 * (size_t)t + d] = pos[t + (size_t)na * d];
    if ((rc = gapcu_group_set_structure(g_group, na, species, lat_c, pos_c.data(), rcut, nullptr))) return rc;
    if ((rc = gapcu_group_compute(g_group, lgrad ? 1 : 0))) return rc;
    if ((rc = gapcu_group_fetch(g_group, ene, f_c.data(), stress))) return rc;
    for (int t = 0; t < na; t++) for (int d = 0; d < 3; d++) force[t + (size_t)na * d] = f_c[3 * (size_t)t + d];
    return 0;
}

extern "C" int gapcu_calc(int na, const int *species, const double *lat, const double *pos, int nsparsex, int des_len,
                          const double *theta, const double *mm, const double *qmm, const double *coeff, double rcut,
                          int lgrad, double *ene, double *force, double *stress, double *variance) {
    (void)qmm;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_devices.size() > 1 && na > 0) {
        const char *e_min = getenv("GAPCU_GROUP_MIN_ATOMS");
        const int min_atoms = e_min ? atoi(e_min) : 20000;
        if (na >= min_atoms) {
            const int rc = calc_on_group(na, species, lat, pos, nsparsex, des_len, theta, mm, coeff, rcut, lgrad, ene, force, stress);
            // a cell too small to cut (GAPCU_EARG from the brick planner) falls through to one device
            if (rc != GAPCU_EARG) { if (!rc && variance) *variance = 0.0; return rc; }
        }
    }
    // GAPCU_TRACE=1: host-side time stamps of this call's phases on stderr (development aid)
    static const bool trace = getenv("GAPCU_TRACE") != nullptr;
    std::chrono::steady_clock::time_point tp[6];
    auto stamp = [&](int k) { if (trace) tp[k] = std::chrono::steady_clock::now(); };
    stamp(0);
    gapcu_ctx *c = nullptr;
    int rc = default_ctx(&c);
    if (rc) return rc;
    if (na <= 0) return fail(GAPCU_EARG, "NA must be positive");
    DeviceGuard dg_(c->device);
    if ((rc = refresh_sf_from_cwd(c))) return rc;
    if (des_len != c->plan.D) return fail(GAPCU_EARG, "des_len does not equal 2*nsf of ./gap_parameters");
    // GPR data: compare the caller's (Fortran-layout) arrays with the last call's; only a
    // changed set is transposed (mm(nsparseX,des_len) column-major -> C order) and uploaded.
    // Up to 1 MiB of MM (the shipped potential: 68 KB) the comparison is exact (memcmp, ~5 us).  A large
    // block (BASELINE config 5: 20 MB, 2 ms per call -- longer than the evaluation) is recognised by its
    // address, its sizes, theta and coeff compared in full, and a fingerprint of 8192 evenly spread
    // entries of MM plus its first and last one; a caller that edits single entries of a large MM in
    // place between calls must set GAPCU_POTENTIAL_CHECK=full (always memcmp).
    {
        const size_t nmm = (size_t)nsparsex * des_len;
        static std::vector<double> last_mm, last_theta, last_coeff;
        static const double *last_ptr = nullptr;
        static size_t last_n = 0;
        static uint64_t last_fp = 0;
        static const bool full_check = [] { const char *e = getenv("GAPCU_POTENTIAL_CHECK"); return e && !strcmp(e, "full"); }();
        const bool small = full_check || nmm * sizeof(double) <= (1u << 20);
        auto fingerprint = [&]() {
            uint64_t h = 1469598103934665603ull;
            const size_t step = std::max<size_t>(1, nmm / 8192);
            auto mix = [&](double v) { uint64_t b; memcpy(&b, &v, 8); h = (h ^ b) * 1099511628211ull; };
            for (size_t k = 0; k < nmm; k += step) mix(mm[k]);
            if (nmm) mix(mm[nmm - 1]);
            return h;
        };
        bool same = c->have_gpr && c->M == nsparsex && c->D == des_len && last_n == nmm &&
                    last_theta.size() == (size_t)des_len && last_coeff.size() == (size_t)nsparsex &&
                    !memcmp(last_theta.data(), theta, sizeof(double) * des_len) &&
                    !memcmp(last_coeff.data(), coeff, sizeof(double) * nsparsex);
        uint64_t fp = 0;
        if (same) {
            if (small) same = last_mm.size() == nmm && !memcmp(last_mm.data(), mm, sizeof(double) * nmm);
            else { fp = fingerprint(); same = mm == last_ptr && fp == last_fp; }
        }
        if (!same) {
            std::vector<double> mm_c(nmm);
            for (int k = 0; k < des_len; k++)
                for (int s = 0; s < nsparsex; s++) mm_c[(size_t)s * des_len + k] = mm[s + (size_t)nsparsex * k];
            c->have_gpr = false;
            if ((rc = set_gpr(c, nsparsex, des_len, theta, mm_c.data(), coeff))) return rc;
            if (small) last_mm.assign(mm, mm + nmm); else { last_mm.clear(); last_mm.shrink_to_fit(); }
            last_theta.assign(theta, theta + des_len); last_coeff.assign(coeff, coeff + nsparsex);
            last_ptr = mm; last_n = nmm; last_fp = small ? 0 : fingerprint();
        }
    }
    double lat_c[9];
    for (int r = 0; r < 3; r++) for (int col = 0; col < 3; col++) lat_c[r * 3 + col] = lat[r + 3 * col];
    stamp(1);
    if ((rc = set_structures_impl(c, 1, &na, species, lat_c, pos, true, rcut, true))) return rc;
    stamp(2);
    if ((rc = enqueue_pass(c, lgrad ? 1 : 0, nullptr, false))) return rc;
    stamp(3);
    // results: flags | out8 | force SoA (= Fortran FORCE(NA,3)) in one batch, one synchronisation
    size_t b_f = sizeof(double) * 3 * (size_t)na;
    if (c->pin(c->res_o_f + b_f)) return fail(GAPCU_ECUDA, "cudaMallocHost failed");
    DevFlags *h_fl = (DevFlags *)c->h_pin;
    double *h_out = (double *)((char *)c->h_pin + c->res_o_out), *h_f = (double *)((char *)c->h_pin + c->res_o_f);
    for (int attempt = 0;; attempt++) {
        CU(cudaMemcpyAsync(c->h_pin, c->d_results.p, c->res_o_f + b_f, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->h2d_pending = false;
        c->h_flags = *h_fl;
        // a list outgrew the learned capacity: react_to_flags enlarges and runs the pass again
        rc = react_to_flags(c, attempt);
        if (rc < 0) return rc;
        if (rc == 0) break;
    }
    if (c->h_flags.close_pairs)
        fprintf(stdout, " Warning: The distance of two atoms is very small (%d pairs below 0.5)\n", c->h_flags.close_pairs);
    stamp(4);
    *ene = h_out[0];
    for (int q = 0; q < 6; q++) stress[q] = h_out[1 + q];
    memcpy(force, h_f, b_f);
    if (variance) *variance = 0.0;  // gap_calc.f90:206
    stamp(5);
    if (trace) {
        auto us = [&](int a, int b) { return std::chrono::duration<double, std::micro>(tp[b] - tp[a]).count(); };
        fprintf(stderr, "gapcu_calc trace (us): potential check %.1f | set_structures + H2D %.1f | enqueue %.1f | D2H + wait %.1f | copy out %.1f | total %.1f\n",
                us(0, 1), us(1, 2), us(2, 3), us(3, 4), us(4, 5), us(0, 5));
    }
    return 0;
}

extern "C" int gapcu_set_devices(int n, const int *devices) {
    std::lock_guard<std::mutex> lk(g_mu);
    const int have = gapcu_device_count();
    if (have <= 0) return fail(GAPCU_ENODEV, "no CUDA device: gapcu has no CPU fallback");
    if (n < 0 || (n > 0 && !devices)) return fail(GAPCU_EARG, "bad device list");
    std::vector<int> want(devices, devices + n);
    for (size_t k = 0; k < want.size(); k++) {
        if (want[k] < 0 || want[k] >= have) return fail(GAPCU_EARG, "device " + std::to_string(want[k]) + " does not exist");
        // a device may be listed several times: that many contexts (batch shards / bricks) share it
    }
    g_devices = want;
    // contexts on devices that are no longer wanted go away; the others keep their state
    if (g_ctx && g_ctx->device != default_device()) { gapcu_ctx_destroy(g_ctx); g_ctx = nullptr; g_file = FileId(); }
    for (BatchSlot &b : g_batch) if (b.ctx) gapcu_ctx_destroy(b.ctx);
    g_batch.clear();
    if (g_group) { gapcu_group_destroy(g_group); g_group = nullptr; g_group_devs.clear(); }
    return 0;
}

// A CALYPSO-style batch through the drop-in side channel: the potential (SF table AND GPR
// block) is ./gap_parameters, as for FGAP_READ + FGAP_CALC; the structures are independent,
// so they are dealt to the devices by estimated cost and every device runs its shard as one
// batched launch sequence on its own context and stream, with no inter-device traffic.
extern "C" int gapcu_calc_batch(int nstruct, const int *natoms, const int *species, const double *lat, const double *pos,
                                double rcut, int lgrad, double *ene, double *force, double *stress) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (nstruct <= 0 || !natoms || !species || !lat || !pos) return fail(GAPCU_EARG, "bad batch arguments");
    for (int s = 0; s < nstruct; s++) if (natoms[s] <= 0) return fail(GAPCU_EARG, "a structure of the batch has no atoms");
    if (gapcu_device_count() <= 0) return fail(GAPCU_ENODEV, "no CUDA device: gapcu has no CPU fallback");
    std::vector<int> devs = g_devices;
    if (devs.empty()) devs.push_back(default_device());
    if (g_batch.size() != devs.size()) {
        for (BatchSlot &b : g_batch) if (b.ctx) gapcu_ctx_destroy(b.ctx);
        g_batch.assign(devs.size(), BatchSlot());
    }
    struct stat st;
    if (stat("gap_parameters", &st) != 0) return fail(GAPCU_EFILE, "gap_parameters file does not exist!");
    for (size_t d = 0; d < devs.size(); d++) {
        BatchSlot &b = g_batch[d];
        if (!b.ctx) { b.ctx = gapcu_ctx_create(devs[d]); b.device = devs[d]; b.file = FileId(); }
        if (!b.ctx) return GAPCU_ECUDA;
        const bool same = b.ctx->have_sf && b.ctx->have_gpr && st.st_dev == b.file.dev && st.st_ino == b.file.ino &&
                          st.st_size == b.file.size && st.st_mtim.tv_sec == b.file.mt_s && st.st_mtim.tv_nsec == b.file.mt_ns;
        if (!same) {
            int rc = gapcu_ctx_load_potential(b.ctx, "gap_parameters");
            if (rc) return rc;
            b.file.dev = st.st_dev; b.file.ino = st.st_ino; b.file.size = st.st_size;
            b.file.mt_s = st.st_mtim.tv_sec; b.file.mt_ns = st.st_mtim.tv_nsec;
        }
    }
    // greedy longest-processing-time deal by N * P^2 (P = mean neighbour count; the triplet work dominates)
    const int nd = (int)devs.size();
    std::vector<size_t> off(nstruct + 1, 0);
    for (int s = 0; s < nstruct; s++) off[s + 1] = off[s] + (size_t)natoms[s];
    std::vector<double> cost(nstruct);
    for (int s = 0; s < nstruct; s++) {
        const double *L = lat + 9 * (size_t)s;
        const double vol = std::fabs(L[0] * (L[4] * L[8] - L[5] * L[7]) - L[1] * (L[3] * L[8] - L[5] * L[6]) + L[2] * (L[3] * L[7] - L[4] * L[6]));
        const double p = 4.0 / 3.0 * 3.141592653589793 * rcut * rcut * rcut * natoms[s] / std::max(vol, 1e-30);
        cost[s] = natoms[s] * p * p;
    }
    std::vector<int> order(nstruct);
    for (int s = 0; s < nstruct; s++) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    std::vector<std::vector<int>> shard(nd);
    std::vector<double> load(nd, 0.0);
    for (int s : order) {
        int best = 0;
        for (int d = 1; d < nd; d++) if (load[d] < load[best]) best = d;
        shard[best].push_back(s);
        load[best] += cost[s];
    }
    struct Packed { std::vector<int> na, sp; std::vector<double> lat, pos, e, f, s; };
    std::vector<Packed> pk(nd);
    for (int d = 0; d < nd; d++) {
        std::sort(shard[d].begin(), shard[d].end());
        if (shard[d].empty()) continue;
        Packed &k = pk[d];
        for (int s : shard[d]) {
            k.na.push_back(natoms[s]);
            k.sp.insert(k.sp.end(), species + off[s], species + off[s + 1]);
            k.lat.insert(k.lat.end(), lat + 9 * (size_t)s, lat + 9 * (size_t)s + 9);
            k.pos.insert(k.pos.end(), pos + 3 * off[s], pos + 3 * off[s + 1]);
        }
        int rc = gapcu_ctx_set_structures(g_batch[d].ctx, (int)k.na.size(), k.na.data(), k.sp.data(), k.lat.data(), k.pos.data(), rcut);
        if (!rc) rc = gapcu_ctx_compute(g_batch[d].ctx, lgrad ? 1 : 0);   // asynchronous: the devices run side by side
        if (rc) return rc;
    }
    for (int d = 0; d < nd; d++) {
        if (shard[d].empty()) continue;
        Packed &k = pk[d];
        k.e.resize(k.na.size()); k.f.resize(k.pos.size()); k.s.resize(6 * k.na.size());
        int rc = gapcu_ctx_fetch(g_batch[d].ctx, k.e.data(), k.f.data(), k.s.data());
        if (rc) return rc;
        size_t fo = 0;
        for (size_t q = 0; q < shard[d].size(); q++) {
            const int s = shard[d][q];
            if (ene) ene[s] = k.e[q];
            if (stress) memcpy(stress + 6 * (size_t)s, k.s.data() + 6 * q, 6 * sizeof(double));
            if (force) memcpy(force + 3 * off[s], k.f.data() + fo, 3 * sizeof(double) * natoms[s]);
            fo += 3 * (size_t)natoms[s];
        }
    }
    return 0;
}

// Zero-fill of a large caller buffer (FGAP_READ's INVCMM = 0 on 4000 x 4000 doubles, gap_calc.f90:361,
// which the reference's ASE calculator pays on every MD step because it builds a fresh Calculator
// per step, gappy/ASE/gap_calc.py:41).  Writing 128 MB of zeros costs ~35 ms of page faults and
// bandwidth.  For whole pages of a PRIVATE ANONYMOUS mapping (what a large malloc / numpy allocation
// is) Linux documents a cheaper way to the same contents: madvise(MADV_DONTNEED) drops the pages and
// "subsequent accesses ... will result in zero-fill-on-demand pages" -- independent of whether the
// pages were present, swapped or never touched, so no page-map inspection is involved.  The range is
// checked against /proc/self/maps first (writable, private, no backing file); anything else, any
// failure, and the partial pages at both ends are cleared with memset.  GAPCU_ZERO=memset forces memset.
static void zero_fill(void *ptr, size_t bytes) {
    const size_t PAGE = (size_t)sysconf(_SC_PAGESIZE);
    static const bool force_memset = [] { const char *e = getenv("GAPCU_ZERO"); return e && !strcmp(e, "memset"); }();
    if (bytes < (8u << 20) || PAGE == 0 || force_memset) { memset(ptr, 0, bytes); return; }
    const uintptr_t a0 = (uintptr_t)ptr, a1 = a0 + bytes;
    const uintptr_t p0 = (a0 + PAGE - 1) / PAGE * PAGE, p1 = a1 / PAGE * PAGE;   // whole pages inside
    bool anon = false;
    if (FILE *mf = fopen("/proc/self/maps", "r")) {
        char line[512];
        uintptr_t cur = a0;
        while (fgets(line, sizeof line, mf)) {
            unsigned long lo = 0, hi = 0, inode = 0;
            char perms[8] = {0}, path[256] = {0};
            const int got = sscanf(line, "%lx-%lx %7s %*x %*s %lu %255s", &lo, &hi, perms, &inode, path);
            if (got < 4) continue;
            if (hi <= cur) continue;
            // a hole, a file mapping, a shared mapping or a special region ([stack], hugetlb, ...): not provably anonymous
            if (lo > cur || inode != 0 || perms[1] != 'w' || perms[3] != 'p' || (got >= 5 && path[0] && strcmp(path, "[heap]") != 0)) break;
            cur = hi;
            if (cur >= a1) { anon = true; break; }
        }
        fclose(mf);
    }
    if (!anon || p1 <= p0 || madvise((void *)p0, p1 - p0, MADV_DONTNEED) != 0) { memset(ptr, 0, bytes); return; }
    memset(ptr, 0, p0 - a0);
    memset((void *)p1, 0, a1 - p1);
}

extern "C" int gapcu_read(const char *path, int *nsparsex, int *des_len, double *theta, int theta_cap, double *mm,
                          int mm_ld, int mm_cols, double *invcmm, int invcmm_ld, double *coeff, int coeff_cap) {
    // the parse is cached by the file's identity (the reference's ASE calculator re-reads it every step)
    static std::mutex mu;
    static PotentialFile pf;
    static std::string pf_path;
    static FileId pf_id;
    std::lock_guard<std::mutex> lk(mu);
    const char *fname = path ? path : "gap_parameters";
    struct stat st;
    const bool have_stat = stat(fname, &st) == 0;
    char full[4096];
    const std::string key = (have_stat && realpath(fname, full)) ? std::string(full) : std::string();
    const bool same = have_stat && !key.empty() && key == pf_path && st.st_dev == pf_id.dev && st.st_ino == pf_id.ino &&
                      st.st_size == pf_id.size && st.st_mtim.tv_sec == pf_id.mt_s && st.st_mtim.tv_nsec == pf_id.mt_ns;
    if (!same) {
        pf_path.clear();
        try {
            pf = read_gap_parameters(fname);
        } catch (const std::exception &e) {
            return fail(GAPCU_EFILE, e.what());
        }
        if (have_stat && !key.empty()) {
            pf_path = key;
            pf_id.dev = st.st_dev; pf_id.ino = st.st_ino; pf_id.size = st.st_size;
            pf_id.mt_s = st.st_mtim.tv_sec; pf_id.mt_ns = st.st_mtim.tv_nsec;
        }
    }
    if (pf.nsparse > mm_ld || pf.nsparse > coeff_cap)
        return fail(GAPCU_ELIMIT, "The siez of sparse set large than nsparseX_max=" + std::to_string(mm_ld));
    if (pf.des_len > mm_cols || pf.des_len > theta_cap)
        return fail(GAPCU_ELIMIT, "The length of descriptors large than nsf_max=" + std::to_string(mm_cols));
    *nsparsex = pf.nsparse;
    *des_len = pf.des_len;
    for (int k = 0; k < pf.des_len; k++) theta[k] = pf.theta[k];
    for (int s = 0; s < pf.nsparse; s++)
        for (int k = 0; k < pf.des_len; k++) mm[s + (size_t)mm_ld * k] = pf.mm[(size_t)s * pf.des_len + k];
    if (invcmm) zero_fill(invcmm, sizeof(double) * (size_t)invcmm_ld * invcmm_ld);
    for (int s = 0; s < pf.nsparse; s++) coeff[s] = pf.coeff[s];
    return 0;
}

extern "C" int gapcu_bond(int na, const double *lat, const int *elements, const double *pos, double rcut,
                          double *min_bond) {
    (void)elements;
    std::lock_guard<std::mutex> lk(g_mu);
    gapcu_ctx *c = nullptr;
    int rc = default_ctx(&c);
    if (rc) return rc;
    if (na <= 0) return fail(GAPCU_EARG, "NA must be positive");
    DeviceGuard dg_(c->device);
    double lat_c[9];
    for (int r = 0; r < 3; r++) for (int col = 0; col < 3; col++) lat_c[r * 3 + col] = lat[r + 3 * col];
    if ((rc = set_structures_impl(c, 1, &na, nullptr, lat_c, pos, true, rcut, false))) return rc;
    CU(cudaMemsetAsync(c->d_flags.p, 0, sizeof(DevFlags), c->stream));
    if ((rc = run_neighbors(c, false, true))) return rc;
    std::vector<double> h((size_t)na);
    CU(cudaMemcpyAsync(h.data(), c->d_mindis.p, sizeof(double) * na, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = read_flags(c))) return rc;
    double m = 10.0;  // get_bond.f90:32
    for (int i = 0; i < na; i++) if (h[i] < m) m = h[i];
    if (c->h_flags.close_pairs)
        fprintf(stdout, " Warning: The distance of two atoms is very small (%d pairs below 0.5)\n", c->h_flags.close_pairs);
    *min_bond = m;
    c->computed = false;
    return 0;
}

extern "C" int gapcu_car2acsf_table(int na, int max_neighbor, int nf, const double *pos, const double *neighbor,
                                    const int *neighbor_count, int lgrad, double *xx, double *dxdy, double *strs) {
    std::lock_guard<std::mutex> lk(g_mu);
    gapcu_ctx *c = nullptr;
    int rc = default_ctx(&c);
    if (rc) return rc;
    if (na <= 0 || max_neighbor <= 0) return fail(GAPCU_EARG, "NA and max_neighbor must be positive");
    if ((rc = refresh_sf_from_cwd(c))) return rc;
    if (nf != c->plan.D) return fail(GAPCU_EARG, "nf does not equal 2*nsf of ./gap_parameters");
    const int D = nf;
    int maxcnt = 0;
    for (int i = 0; i < na; i++) {
        if (neighbor_count[i] < 0 || neighbor_count[i] > max_neighbor) return fail(GAPCU_EARG, "neighbor_count out of range");
        maxcnt = std::max(maxcnt, neighbor_count[i]);
    }
    if (maxcnt > 1023) return fail(GAPCU_ENEIGH, "more than 1023 neighbours");
    DeviceGuard dg_(c->device);
    // a single pseudo structure: the kernels only need the centre positions and the table
    const size_t NA = (size_t)na;
    c->h_structs.assign(1, StructDev());
    StructDev &sd = c->h_structs[0];
    memset(&sd, 0, sizeof sd);
    sd.lat[0] = sd.lat[4] = sd.lat[8] = 1.0; sd.inv[0] = sd.inv[4] = sd.inv[8] = 1.0; sd.volume = 1.0;
    sd.natoms = na; sd.nbins = 1; sd.nbin[0] = sd.nbin[1] = sd.nbin[2] = 1;
    c->h_natoms.assign(1, na);
    c->nstruct = 1; c->ntot = na; c->n_centres = na; c->nbins = 1; c->computed = false; c->pcap_known = false; c->last_ntot = -1; c->lists_valid = false;
    c->cap = std::max(32, round_up(maxcnt, 32)); c->pcap = c->cap;
    if (!c->have_gpr || c->D != D) {  // the GPR part is irrelevant here; give the kernels a consistent dummy
        std::vector<double> th(D, 1.0), m1(D, 0.0), c1(1, 0.0);
        if ((rc = set_gpr(c, 1, D, th.data(), m1.data(), c1.data()))) return rc;
    }
    DBuf<double> d_table;
    std::vector<int> h_sid(NA, 0);
    CU(ensure_inputs(c, 1, NA));
    CU(c->d_nbr_cnt.ensure(NA)); CU(ensure_results(c, 1, NA)); CU(c->d_order.ensure(NA));
    CU(d_table.ensure(NA * max_neighbor * 6));
    if ((rc = ensure_work_buffers(c))) { d_table.release(); return rc; }
    auto bail = [&](int code) { d_table.release(); return code; };
    if (cudaMemcpy(c->d_structs.p, &sd, sizeof sd, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_sid.p, h_sid.data(), sizeof(int) * NA, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_pos.p, pos, sizeof(double) * 3 * NA, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_nbr_cnt.p, neighbor_count, sizeof(int) * NA, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_table.p, neighbor, sizeof(double) * NA * max_neighbor * 6, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(c->d_wgt.p, 0, sizeof(double) * NA) != cudaSuccess)
        return bail(fail(GAPCU_ECUDA, "upload failed"));
    CentreArgs a;
    bool fused = false;
    if ((rc = make_centre_args(c, 1, c->pcap, &a, &fused))) return bail(rc);
    a.nbr_table = d_table.p; a.table_ld = max_neighbor; a.order = nullptr; a.n_centres = nullptr; a.lgrad = 1;
    auto run = [&](bool backward) -> int {
        if (cudaMemsetAsync(c->d_flags.p, 0, sizeof(DevFlags), c->stream) != cudaSuccess) return fail(GAPCU_ECUDA, "memset failed");
        int r = backward ? launch_backward(c->stream, a, &c->launches) : launch_forward(c->stream, a, &c->launches);
        if (r) return fail(GAPCU_ELIMIT, "centre kernel needs more shared memory than an SM has");
        if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(GAPCU_ECUDA, "centre kernel failed");
        return 0;
    };
    // descriptors: xx(nf,na) column-major == G[na][D] row-major
    if ((rc = run(false))) return bail(rc);
    if (cudaMemcpy(xx, c->d_G.p, sizeof(double) * NA * D, cudaMemcpyDeviceToHost) != cudaSuccess) return bail(fail(GAPCU_ECUDA, "download failed"));
    memset(dxdy, 0, sizeof(double) * (size_t)D * NA * NA * 3);
    memset(strs, 0, sizeof(double) * 9 * (size_t)D * NA);
    if (lgrad) {
        // one backward pass per descriptor with dE/dG = e_k: the per-slot gradients ARE dG_k/dr
        std::vector<double> h_fp(NA * c->cap * 3), h_gs(NA * 3), h_vir(NA * 6);
        for (int k = 0; k < D; k++) {
            launch_onehot(c->stream, c->d_dEdG.p, na, D, k);
            if ((rc = run(true))) return bail(rc);
            if (cudaMemcpy(h_fp.data(), c->d_fpair.p, sizeof(double) * h_fp.size(), cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(h_gs.data(), c->d_gself.p, sizeof(double) * h_gs.size(), cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(h_vir.data(), c->d_vir.p, sizeof(double) * h_vir.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
                return bail(fail(GAPCU_ECUDA, "download failed"));
            for (size_t n = 0; n < NA; n++) {
                for (int d = 0; d < 3; d++) dxdy[k + (size_t)D * (n + NA * (n + NA * d))] += h_gs[n * 3 + d];
                for (int s = 0; s < neighbor_count[n]; s++) {
                    const long j = (long)neighbor[n + NA * (s + (size_t)max_neighbor * 5)] - 1;  // real(j), 1-based (gap_calc.f90:115)
                    if (j < 0 || j >= na) return bail(fail(GAPCU_EARG, "neighbor index out of range"));
                    for (int d = 0; d < 3; d++)
                        dxdy[k + (size_t)D * (n + NA * ((size_t)j + NA * d))] += h_fp[(n * c->cap + s) * 3 + d];
                }
                // strs(a,b,k,n) = sum delta_a * dG/dx_b ; the kernel keeps the upper triangle (xx,xy,xz,yy,yz,zz)
                double *S = strs + 9 * ((size_t)k + (size_t)D * n);
                const double *v = &h_vir[n * 6];
                S[0] = v[0]; S[4] = v[3]; S[8] = v[5];
                S[3] = S[1] = v[1]; S[6] = S[2] = v[2]; S[7] = S[5] = v[4];
            }
        }
    }
    d_table.release();
    c->have_gpr = false;      // the dummy GPR set must not survive into the next fgap_calc
    c->pcap_known = false;
    return 0;
}
