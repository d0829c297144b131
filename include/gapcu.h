/*
 * gapcu.h -- C ABI of the B200-native libgap energy/force/stress path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Every "fortran-layout" entry point below is what the reference's Fortran
 * subroutine of the same role would bind through ISO_C_BINDING (see
 * calypso-gap_b200/fortran/libgap_driver.f90 and INTEGRATION.md):
 *
 *   gapcu_calc      <- SUBROUTINE FGAP_CALC   gappy/libgap/gap_calc.f90:1-301
 *   gapcu_read      <- SUBROUTINE FGAP_READ   gappy/libgap/gap_calc.f90:303-364
 *   gapcu_bond      <- SUBROUTINE FGET_BOND   gappy/libgap/get_bond.f90:4-116
 *   gapcu_car2acsf_table <- SUBROUTINE CAR2ACSF gappy/libgap/wacsf.f90:2-796
 *
 * Arrays of those four use the Fortran column-major layout of the argument they
 * replace: pos(NA,3) -> pos[i + NA*c]; lat(3,3) -> lat[r + 3*c] with ROW r the
 * r-th lattice vector (gap_calc.f90:98); force like pos; mm(nsparseX,des_len)
 * -> mm[s + nsparseX*k]; stress(6) = xx yy zz xy yz xz in GPa
 * (gap_calc.f90:221-226).  LOGICAL arguments are int (0/1).
 *
 * Like the reference, gapcu_calc / gapcu_car2acsf_table take the species weights and
 * the symmetry-function table from the file ./gap_parameters in the current
 * working directory (gap_calc.f90:75-83, wacsf.f90:40-56); the parsed file is
 * cached and re-read only when its (device, inode, size, mtime) change.
 *
 * All functions return 0 on success or a negative GAPCU_E* code; the message is
 * available from gapcu_last_error() (thread local).  There is NO CPU fallback:
 * without a CUDA device every compute entry point returns GAPCU_ENODEV.
 */
#ifndef GAPCU_H
#define GAPCU_H

#ifdef __cplusplus
extern "C" {
#endif

#define GAPCU_OK 0
#define GAPCU_EFILE -1      /* gap_parameters missing or malformed (reference: print + stop) */
#define GAPCU_ENEIGH -2     /* an atom has > 1000 neighbours (gap_calc.f90:107-111: stop)     */
#define GAPCU_ESPECIES -3   /* species absent from gap_parameters (reference: uninitialised)   */
#define GAPCU_ELIMIT -4     /* potential exceeds a compiled limit of this library             */
#define GAPCU_ECUDA -5      /* CUDA runtime error                                             */
#define GAPCU_ENODEV -6     /* no CUDA device                                                 */
#define GAPCU_EARG -7       /* invalid argument                                               */
#define GAPCU_EDOMAIN -8    /* decomposed run: an atom left its brick, the caller must re-partition */

const char *gapcu_last_error(void);

/* ---- Fortran-layout entry points (the reference's f2py surface) ---------- */

/* FGAP_CALC.  qmm is accepted and ignored (the reference never reads it:
 * gap_calc.f90:205-210), variance is set to 0 (gap_calc.f90:206). */
int gapcu_calc(int na, const int *species, const double *lat, const double *pos,
               int nsparsex, int des_len, const double *theta, const double *mm,
               const double *qmm, const double *coeff, double rcut, int lgrad,
               double *ene, double *force, double *stress, double *variance);

/* FGAP_READ with explicit capacities.  theta[theta_cap], mm[mm_ld x mm_cols]
 * column-major, coeff[coeff_cap]; entries beyond the file's sizes are left
 * untouched.  invcmm (may be NULL) is zero-filled as invcmm_ld x invcmm_ld
 * (gap_calc.f90:361).  Host only, no GPU needed. */
int gapcu_read(const char *path, int *nsparsex, int *des_len, double *theta, int theta_cap,
               double *mm, int mm_ld, int mm_cols, double *invcmm, int invcmm_ld, double *coeff,
               int coeff_cap);

/* FGET_BOND: smallest image distance <= rcut (10.0 if none, get_bond.f90:32). */
int gapcu_bond(int na, const double *lat, const int *elements, const double *pos, double rcut,
               double *min_bond);

/* CAR2ACSF with the reference's own arguments: neighbor(NA,max_neighbor,6) holds
 * per neighbour the absolute image position (1:3), the distance (4), the species
 * weight (5) and real(j) (6) as built by gap_calc.f90:112-115; outputs
 * xx(nf,na), dxdy(nf,na,na,3), strs(3,3,nf,na), all column-major.  The SF table
 * comes from ./gap_parameters (wacsf.f90:40-56) and nf must equal 2*nsf. */
int gapcu_car2acsf_table(int na, int max_neighbor, int nf, const double *pos, const double *neighbor,
                         const int *neighbor_count, int lgrad, double *xx, double *dxdy, double *strs);

/* ---- additive drop-in entry points (same ./gap_parameters side channel) --- */

/* Devices used by the entry points of this section and the previous one (default: device
 * GAPCU_DEVICE or 0).  gapcu_calc / gapcu_bond / gapcu_car2acsf_table run on devices[0];
 * gapcu_calc_batch shards over all of them.  n = 0 restores the default. */
int gapcu_set_devices(int n, const int *devices);

/* A batch of independent structures, CALYPSO-search style (what gappy/tools/cgg2.py:111-115
 * does with one process per structure): nstruct structures in the C-order layout of
 * gapcu_ctx_set_structures, the potential -- SF table AND GPR block -- read from
 * ./gap_parameters (parsed once per device, re-read when the file changes).  Structures are
 * dealt to the devices by estimated cost, every device evaluates its shard as one batched
 * launch sequence, concurrently, with no inter-device communication.  ene[nstruct],
 * force[sum natoms][3], stress[nstruct][6] (xx yy zz xy yz xz, GPa); any may be NULL. */
int gapcu_calc_batch(int nstruct, const int *natoms, const int *species, const double *lat, const double *pos,
                     double rcut, int lgrad, double *ene, double *force, double *stress);

/* prints gapcu_last_error() on stdout the way the reference prints before STOP */
void gapcu_print_last_error(void);

/* ---- additive API: persistent context, batches, device-side timing ------- */

typedef struct gapcu_ctx gapcu_ctx;

int gapcu_device_count(void);
gapcu_ctx *gapcu_ctx_create(int device);          /* NULL on failure */
void gapcu_ctx_destroy(gapcu_ctx *ctx);

/* potential = SF table + species weights + GPR data, from a file ...          */
int gapcu_ctx_load_potential(gapcu_ctx *ctx, const char *path);
/* ... or from arrays (C order: mm[nsparse][des_len]).                          */
int gapcu_ctx_set_potential(gapcu_ctx *ctx, int nspecies, const int *z, const double *w, int nsf,
                            const int *ntype, const double *alpha, const double *cutoff,
                            int nsparse, int des_len, const double *theta, const double *mm,
                            const double *coeff);

/* Pipeline choice: 0 = automatic (default), 1 = split: forward kernel -> tiled DMMA GPR
 * kernel -> backward kernel, 2 = fused: one centre kernel with the GPR of each atom
 * done inside its CTA (chosen automatically while the sparse set is small).  Also
 * settable with the environment variable GAPCU_PIPELINE=split|fused. */
int gapcu_ctx_set_pipeline(gapcu_ctx *ctx, int mode);

/* CTAs per centre atom of the fused kernel (a thread-block cluster shares one centre: each CTA
 * takes a part of the neighbour-pair range, partial descriptors and gradients are combined
 * through distributed shared memory in rank order).  0 = automatic (default: 2 or 4 when the
 * launch has fewer centres than the device has SMs, e.g. the 64-atom cells of an MD run;
 * otherwise 1), or 1, 2, 4.  Also settable with GAPCU_CLUSTER=1|2|4.  Results differ from the
 * one-CTA evaluation only by summation order (~1e-15 relative). */
int gapcu_ctx_set_cluster(gapcu_ctx *ctx, int ctas_per_centre);

/* Spatial decomposition of ONE large structure over the ranks of a node (SURVEY.md 8(e),
 * BASELINE config 4).  The cell is cut into g0 x g1 x g2 bricks in fractional coordinates; rank
 * (m0*g1 + m1)*g2 + m2 evaluates the atoms of brick (m0,m1,m2).  gapcu_ctx_set_structures is given
 * the WHOLE structure on every rank (what FGAP_CALC's caller holds) and keeps only the atoms of its
 * brick; before every neighbour build the ranks exchange the ghost images within rcut + skin + drift
 * of each brick (grouped ncclSend/ncclRecv, 26 directions) and after the force gather the gradients
 * a rank's centres put on ghosts return to the owners; E, the stress sums and the status flags are
 * combined with one all-gather of a 48-double record.  gapcu_ctx_fetch then returns E and stress of
 * the whole structure and the forces of THIS RANK'S atoms, force[n_owned][3] in the order of
 * gapcu_ctx_owned.  g0*g1*g2 == 1 switches the decomposition off.
 * Every brick must be at least rcut + skin + 2*drift thick (GAPCU_EARG otherwise).  After
 * gapcu_ctx_update_positions an atom may have left its brick: up to `drift` Angstrom (default 0.25,
 * gapcu_ctx_set_drift) that is absorbed by the ghost shell; beyond, fetch returns GAPCU_EDOMAIN and
 * the caller sets the structure again, which re-partitions.
 * NCCL is loaded with dlopen at the first call; the unique id made by rank 0 with
 * gapcu_nccl_unique_id (128 bytes) must be sent to the other ranks by the caller.  compute and
 * fetch are collective: every rank calls them, and a re-run (a capacity outgrown, stale skin lists)
 * is decided from flags merged over all ranks, so all ranks re-run together. */
int gapcu_nccl_unique_id(char *out128);
int gapcu_ctx_nccl_init(gapcu_ctx *ctx, int nranks, int rank, const char *id128);
int gapcu_ctx_set_domain(gapcu_ctx *ctx, int g0, int g1, int g2, int m0, int m1, int m2);
int gapcu_ctx_set_drift(gapcu_ctx *ctx, double drift_angstrom);
/* atoms this context returns forces for: all of them, or the owned atoms of a decomposed run
 * (ids = indices into the structure given to gapcu_ctx_set_structures; ids may be NULL) */
int gapcu_ctx_owned(gapcu_ctx *ctx, int *n_owned, int *ids);

/* The same decomposition inside ONE process: nranks contexts on the listed devices (a device may
 * appear several times: several bricks on one GPU), ghost records and gradients moved by
 * device-to-device copies ordered with events instead of NCCL.  All arrays are those of the whole
 * structure (C order); grid3 = NULL picks the brick grid with the smallest ghost shell. */
typedef struct gapcu_group gapcu_group;
gapcu_group *gapcu_group_create(int nranks, const int *devices);
void gapcu_group_destroy(gapcu_group *g);
int gapcu_group_size(gapcu_group *g);
gapcu_ctx *gapcu_group_ctx(gapcu_group *g, int rank);
int gapcu_group_load_potential(gapcu_group *g, const char *path);
int gapcu_group_set_skin(gapcu_group *g, double skin);
int gapcu_group_set_structure(gapcu_group *g, int na, const int *species, const double *lat, const double *pos,
                              double rcut, const int *grid3);
int gapcu_group_update_positions(gapcu_group *g, const double *pos, int reuse_lists);
int gapcu_group_compute(gapcu_group *g, int lgrad);
int gapcu_group_fetch(gapcu_group *g, double *ene, double *force, double *stress);

/* Verlet-skin reuse of the neighbour lists across MD / relaxation steps (the
 * reference rebuilds its table on every call, gap_calc.f90:89-120).  With a skin > 0 the candidate
 * lists hold every image within rcut + skin; gapcu_ctx_update_positions(pos, reuse_lists = 1) uploads
 * moved positions of the SAME atoms and the next compute only re-tests the candidates against rcut
 * with the reference's arithmetic, so the neighbour sets stay exactly the reference's.  When an atom
 * has moved more than skin/2 since the lists were built the library rebuilds them by itself (at
 * fetch).  pos: [ntot][3], C order (decomposed runs: the owned atoms, [n_owned][3]).  reuse_lists = 0
 * rebuilds everything from the new positions. */
int gapcu_ctx_set_skin(gapcu_ctx *ctx, double skin_angstrom);
int gapcu_ctx_update_positions(gapcu_ctx *ctx, const double *pos, int reuse_lists);

/* A batch of nstruct independent periodic structures (C order this time:
 * natoms[nstruct]; species[sum natoms]; lat[nstruct][3][3] rows = lattice
 * vectors; pos[sum natoms][3]).  Copies host -> device. */
int gapcu_ctx_set_structures(gapcu_ctx *ctx, int nstruct, const int *natoms, const int *species,
                             const double *lat, const double *pos, double rcut);
/* Enqueue the whole E/F/stress pipeline on the context's stream (asynchronous). */
int gapcu_ctx_compute(gapcu_ctx *ctx, int lgrad);
/* Wait and copy device -> host.  ene[nstruct], force[sum natoms][3] (C order),
 * stress[nstruct][6]; any may be NULL. */
int gapcu_ctx_fetch(gapcu_ctx *ctx, double *ene, double *force, double *stress);
/* Descriptors / dE/dG / atomic energies of the last compute ([sum natoms][des_len], C order). */
int gapcu_ctx_fetch_descriptors(gapcu_ctx *ctx, double *xx, double *dedg, double *eatom);
/* Predictive variance of the last compute (additive).  The reference carries the
 * formula commented out and returns VARIANCE = 0 (gappy/libgap/gap_calc.f90:205-210), which the
 * drop-in entry points above keep doing; this entry point evaluates it for callers that hold QMM:
 *     covf(i) = delta - ckm(i,:) . matmul(QMM, ckm(i,:)),  VARIANCE = sum_i covf(i) / na,
 * ckm = GET_COV (gap_calc.f90:268-288), delta = 1 (gap_calc.f90:8).
 * qmm: QMM(nsparseX, nsparseX) in Fortran order; variance[nstruct]; covf[sum natoms] or NULL. */
int gapcu_ctx_variance(gapcu_ctx *ctx, const double *qmm, double *variance, double *covf);
/* Neighbour lists of the last compute in reference order (j, n1, n2, n3):
 * count[ntot], idx[ntot][cap] (index within the structure), shift[ntot][cap][3],
 * dis[ntot][cap].  Returns the largest count, or a negative code. */
int gapcu_ctx_fetch_neighbors(gapcu_ctx *ctx, int cap, int *count, int *idx, int *shift, double *dis);

/* Debug: the neighbour pairs (triplets centre-j-k) the angular symmetry functions are summed over,
 * as the kernel keeps them after the reference's three cutoff tests (wacsf.f90:177-244, "k_neighbor >
 * j_neighbor", "rjk .gt. cutoff").  items[atom][k] = slot_j | slot_k << 10 | nclasses << 20 with
 * slot_j < slot_k positions in the atom's neighbour list (reference order) and the pair a member of
 * the cutoff classes 0..nclasses-1 (distinct SF cutoffs, descending).  count[ntot], items[ntot][cap].
 * Returns the largest count (entries beyond cap are dropped) or a negative code. */
int gapcu_ctx_debug_triplets(gapcu_ctx *ctx, int cap, int *count, unsigned *items);

/* Device-side timing of `steps` back-to-back gapcu_ctx_compute passes with CUDA
 * events on the context's stream (inputs resident).  If l2_flush_bytes > 0 a
 * buffer of that size is overwritten between passes, outside the timed events.
 * ms_total = sum of per-pass event times.  stage_ms (may be NULL) receives the
 * summed time of each pipeline stage, GAPCU_NSTAGE entries, measured in a
 * separate instrumented run of the same passes (see gapcu_stage_name). */
#define GAPCU_NSTAGE 8
int gapcu_ctx_time_compute(gapcu_ctx *ctx, int lgrad, int steps, long l2_flush_bytes,
                           double *ms_total, double *stage_ms, long *launches);
const char *gapcu_stage_name(int stage);

/* Work counters of the last compute, for the roofline:
 * out[0]=atoms, [1]=sum P (pairs), [2]=sum_c P_c, [3]=candidate pairs tested,
 * [4]=kept triplets, [5]=sum_c T_c (triplet-class evaluations),
 * [6]=sum_c n_ang(c) T_c, [7]=sum_c n_rad(c) P_c, [8]=sum_c Q_c (candidate pairs
 * per angular cutoff class, P_c(P_c-1)/2). */
int gapcu_ctx_work_counters(gapcu_ctx *ctx, double *out, int n);

/* Load balance of the persistent centre kernel in the last compute: out4 = {CTAs, span of
 * the kernel in us, time until the first CTA ran out of work in us, mean busy fraction}. */
int gapcu_ctx_balance(gapcu_ctx *ctx, double *out4);

/* FP64 peak micro-benchmarks on the context's device: DFMA-chain (CUDA cores)
 * and mma.sync m8n8k4 f64 (DMMA).  TFLOP/s each. */
int gapcu_fp64_peaks(gapcu_ctx *ctx, double *dfma_tflops, double *dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* GAPCU_H */
