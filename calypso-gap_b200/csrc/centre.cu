// centre.cu -- host side of the per-centre wACSF kernel: shared-memory layout and dispatch to
// the capacity-specialised instances (centre_impl.cuh, centre_p*.cu).
#include <cstdlib>

#include "centre_impl.cuh"

namespace gapcu {

// neighbour capacity of the instance that serves a runtime capacity
int centre_pcap_template(int pcap) { return pcap <= 128 ? 128 : pcap <= 256 ? 256 : pcap <= 512 ? 512 : 1024; }

static SmemLayout make_layout(const CentreArgs &a, int mode) {
    SmemLayout L;
    memset(&L, 0, sizeof L);
    const int pt = centre_pcap_template(a.pcap), D = a.plan.D;
    int o = hot_bytes(pt, a.plan.ncls);   // Hot<PCAP>: neighbour records, gradient accumulator, class counts, fc tables
    auto take = [&](long bytes) { int r = o; o += (int)((bytes + 15) & ~15l); return r; };
    take(0);
    const bool bwd = mode != MODE_FWD, fwd = mode != MODE_BWD, fused = mode >= MODE_FUSED;
    L.t32 = take(8 * 32);
    L.t2 = take(8 * MAXC_DEV);
    L.galpha = take(8 * (a.plan.n_grp + 1));
    L.gd = bwd ? take(8 * 4 * (a.plan.n_grp + 1)) : 0;
    L.sG = fused ? take(8 * D) : 0;
    L.gx = (fused && a.cs > 1) ? take(8 * D) : 0;
    L.sdu = bwd ? take(8 * D) : 0;
    L.red = take(8 * NW * 16);
    L.S = take(4 * (a.lcap + 32));
    // scratch region, three lives: [U | gw] while lists are built and the forward runs,
    // [part | xs | W] during the in-CTA GPR, [private accumulators] during the backward
    const long scr_u = ((4l * (a.lcap + NW * 64 + 32) + 15) & ~15l);   // per-warp parts: a fair share + up to 63 entries
    const long scr_fwd = scr_u + (fwd ? 8l * NW * D : 0);
    const long gpr_part = 8l * NW * (a.gpr_Mp > D ? a.gpr_Mp : D);
    const long scr_gpr = fused ? gpr_part + 8l * (D + a.gpr_Mp + 8) : 0;
    const long scr_pa = (bwd && a.npa > 1) ? 8l * a.npa * 3 * pt : 8l * CT;
    long scr = scr_fwd > scr_pa ? scr_fwd : scr_pa;
    if (scr_gpr > scr) scr = scr_gpr;
    L.scratch = take(scr);
    L.gw = L.scratch + (int)scr_u;
    L.xs = L.scratch + (int)gpr_part;
    L.sW = L.xs + 8 * ((D + 1) & ~1);
    L.ctl = take(4 * 512);
    L.rad = take(16 * (a.plan.n_rad + 1));
    L.total = o;
    return L;
}

size_t centre_smem_bytes(const CentreArgs &a, int mode) { return (size_t)make_layout(a, mode).total; }
int centre_warps() { return NW; }
size_t centre_stash_words(const CentreArgs &a, int chunks, int ctas) { return (size_t)ctas * chunks * (size_t)(a.lcap + 32 + 512); }

static int launch_mode(cudaStream_t st, const CentreArgs &a_in, int mode) {
    CentreArgs a = a_in;
    a.lay = make_layout(a, mode);
    const int pt = centre_pcap_template(a.pcap);
    a.queue_slot = mode * 4 + (pt == 128 ? 0 : pt == 256 ? 1 : pt == 512 ? 2 : 3);
    switch (centre_pcap_template(a.pcap)) {
        case 128: return launch_centre_p128(st, a, mode);
        case 256: return launch_centre_p256(st, a, mode);
        case 512: return launch_centre_p512(st, a, mode);
        default: return launch_centre_p1024(st, a, mode);
    }
}

int launch_forward(cudaStream_t st, const CentreArgs &a, long *launches) {
    if (launches) *launches += 1;
    return launch_mode(st, a, MODE_FWD);
}
int launch_backward(cudaStream_t st, const CentreArgs &a, long *launches) {
    if (launches) *launches += 1;
    return launch_mode(st, a, MODE_BWD);
}
int launch_fused(cudaStream_t st, const CentreArgs &a, long *launches) {
    if (launches) *launches += 1;
    // parked exponentials pay on large launches (+1 % at 27k centres); on a 1000-centre launch the extra L2
    // round trips cost 4 %
    static const int se_min = getenv("GAPCU_SE_MIN") ? atoi(getenv("GAPCU_SE_MIN")) : 4096;   // A/B switch
    const bool se = a.share_exp && a.estash && a.cs == 1 && centre_pcap_template(a.pcap) <= 256 && a.ncentres_max >= se_min;
    return launch_mode(st, a, se ? MODE_FUSED_SE : MODE_FUSED);
}

}  // namespace gapcu
