/* ucgb200.h — C-ABI of libucgb200.so: the B200 (sm_100a) implementation of the
 * LAMMPS UCG package's per-timestep hot path.
 *
 * This is the drop-in boundary.  Everything the reference's
 * LAMMPS styles do inside Pair::compute / Fix::initial_integrate /
 * Fix::post_force / Fix::final_integrate / AtomVec::force_clear and everything
 * stock LAMMPS does for them inside Neighbor::decide/build and Comm::forward/
 * reverse/borders is reachable from here with plain pointers and sizes.
 * The host-side style classes (lammps-ucg-dev_b200/host/) and the Python
 * ctypes binding (tests, bench) are the only callers.
 *
 * Conventions
 *   - every entry point returns int: 0 = ok, <0 = API misuse / CUDA failure
 *     (text via ucgb200_last_error), >0 = runtime condition mirrored from the
 *     reference's error->one() texts (see UCGB200_ERR_*).
 *   - one context <-> one GPU <-> one calling thread; calls are ordered on the
 *     context's stream and asynchronous until a *_download/_sync/_status call.
 *   - host buffers are caller-owned, device memory is context-owned; pointers
 *     named d_* are DEVICE pointers (for halo exchange plumbing), all others are
 *     HOST pointers.
 *   - reals are FP64, ids/types/states are int32, exactly as in the reference
 *     (atom.h:180-192 of the reference).
 *   - atom TYPES are "actual" types; "formal" types only index tables and
 *     chemical potentials (UCG/pair_table_ucgld.cpp:114,176,430).
 */
#ifndef UCGB200_H
#define UCGB200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ucgb200_ctx ucgb200_ctx;

/* runtime conditions (positive return codes / status word) */
#define UCGB200_OK 0
#define UCGB200_ERR_TABLE_INNER 1  /* "Pair distance < table inner cutoff"  pair_table_ucgld.cpp:223,279,354,437 */
#define UCGB200_ERR_TABLE_OUTER 2  /* "Pair distance > table outer cutoff"  pair_table_ucgld.cpp:228,... */
#define UCGB200_ERR_NEIGH_OVERFLOW 3 /* internal: neighbor row capacity exceeded (list is regrown and rebuilt) */
#define UCGB200_ERR_DENSITY_TYPE 4 /* rleucg/bethe_density: density CV defined for actual type 1 only (Q15) */
#define UCGB200_ERR_LOST_ATOMS 5   /* an atom left the box by more than one period */
#define UCGB200_ERR_BOX_TOO_SMALL 6 /* periodic box shorter than cut+skin */
#define UCGB200_ERR_PEER_TIMEOUT 7  /* multi-brick run: a peer brick's halo did not arrive within 2 s (peer-mapped forward exchange) */

/* table styles: enum{LOOKUP, LINEAR, SPLINE, BITMAP} pair_table_ucgld.h */
#define UCGB200_TAB_LOOKUP 0
#define UCGB200_TAB_LINEAR 1
#define UCGB200_TAB_SPLINE 2
#define UCGB200_TAB_BITMAP 3

/* field masks for atoms_upload / atoms_download (AtomVecUCG field lists,
 * UCG/atom_vec_ucg.cpp:48-90) */
#define UCGB200_F_X        (1u << 0)   /* double[3n] */
#define UCGB200_F_V        (1u << 1)   /* double[3n] */
#define UCGB200_F_F        (1u << 2)   /* double[3n] */
#define UCGB200_F_TYPE     (1u << 3)   /* int[n] */
#define UCGB200_F_MASK     (1u << 4)   /* int[n] */
#define UCGB200_F_TAG      (1u << 5)   /* int[n] */
#define UCGB200_F_MOLECULE (1u << 6)   /* int[n] */
#define UCGB200_F_UCGSTATE (1u << 7)   /* int[n] */
#define UCGB200_F_UCGL     (1u << 8)   /* double[n] */
#define UCGB200_F_UCGVL    (1u << 9)   /* double[n] */
#define UCGB200_F_UCGML    (1u << 10)  /* double[n] */
#define UCGB200_F_UCGP     (1u << 11)  /* double[n] */
#define UCGB200_F_UCGFORCE (1u << 12)  /* double[n] */
#define UCGB200_F_SCORES   (1u << 13)  /* double[2n]  ucgsoftmaxscores */
#define UCGB200_F_NUMSTATES (1u << 14) /* int[n]  num_ucgstates (download only) */
#define UCGB200_F_ALL 0x7fffu

/* host view of the per-atom arrays, AoS exactly as LAMMPS holds them
 * (double **x is a contiguous [n][3] block).  Unused members may be NULL. */
typedef struct ucgb200_atoms {
  double *x, *v, *f;
  int *type, *mask, *tag, *molecule;
  int *ucgstate;
  double *ucgl, *ucgvl, *ucgml, *ucgp, *ucgforce;
  double *ucgsoftmaxscores;
  int *num_ucgstates;
} ucgb200_atoms;

/* ---------------------------------------------------------------- lifecycle */
int ucgb200_create(int device, ucgb200_ctx **out);
int ucgb200_destroy(ucgb200_ctx *ctx);
const char *ucgb200_last_error(const ucgb200_ctx *ctx);
/* run on an externally owned stream (cudaStream_t as void*); NULL = own stream */
int ucgb200_set_stream(ucgb200_ctx *ctx, void *cuda_stream);
int ucgb200_sync(ucgb200_ctx *ctx);
/* number of kernels this context has launched so far (bench gpu_launches) */
long long ucgb200_launch_count(const ucgb200_ctx *ctx);

/* ------------------------------------------------------------ global set-up */
/* force->boltz, ftm2v, mvv2e of the deck's `units` [stock Force] */
int ucgb200_set_units(ucgb200_ctx *ctx, double boltz, double ftm2v, double mvv2e);
/* domain->boxlo/boxhi/periodicity [stock Domain]; orthogonal boxes only */
int ucgb200_set_box(ucgb200_ctx *ctx, const double lo[3], const double hi[3], const int periodic[3]);
/* page-locked host memory for callers that stage large downloads (dump rows / text) */
int ucgb200_pinned_alloc(size_t bytes, void **out);
int ucgb200_pinned_free(void *p);
/* page-lock / release caller-owned arrays in place (cudaHostRegister); unregister before the memory is freed or moved */
int ucgb200_host_register(void *p, size_t bytes);
int ucgb200_host_unregister(void *p);
int ucgb200_get_box(const ucgb200_ctx *ctx, double lo[3], double hi[3], int periodic[3]);
/* sub-domain owned by this context (multi-GPU brick); defaults to the box */
int ucgb200_set_subdomain(ucgb200_ctx *ctx, const double sublo[3], const double subhi[3]);
/* update->dt */
int ucgb200_set_timestep(ucgb200_ctx *ctx, double dt);
/* force->special_lj[0..3] */
int ucgb200_set_special_lj(ucgb200_ctx *ctx, const double special_lj[4]);

/* type maps read by read_state_settings (pair_table_ucgld.cpp:565-652):
 *   n_states[1..n_actual], formal_from_actual[(t)*2 + s] (0 = undefined),
 *   chem_pot[1..n_formal], mass[1..ntypes] (atom->mass, ntypes == n_formal).
 * Arrays are 1-based like LAMMPS (element 0 unused). */
int ucgb200_set_types(ucgb200_ctx *ctx, int n_actual, int n_formal, const int *n_states,
                      const int *formal_from_actual, const double *chem_pot, const double *mass);
/* kT = boltz * t_target of the first fix exporting "t_target"
 * (pair_table_ucgld.cpp:876-881, fix_ucgstate.cpp:148-156) */
int ucgb200_set_kT(ucgb200_ctx *ctx, double kT);

/* one tabulated potential as built by compute_table (pair_table_ucgld.cpp:1105-1344).
 * n = number of entries of e/f (tablength for LINEAR/SPLINE, tablength-1 for LOOKUP,
 * 1<<tablength for BITMAP).  e2/f2 only for SPLINE; rsq/drsq/de/df only for BITMAP
 * (for LINEAR de/df/rsq are recomputed on the device bit-exactly from e,f,innersq,delta).
 * Returns the table index (>= 0) in *index. */
int ucgb200_table_upload(ucgb200_ctx *ctx, int tabstyle, int tablength, int n, double innersq,
                         double delta, double invdelta, double deltasq6, double cut, int nmask,
                         int nshiftbits, const double *e, const double *f, const double *e2,
                         const double *f2, const double *rsq, const double *drsq, const double *de,
                         const double *df, int *index);
/* pair_style settings(): drops every table and the style-specific configuration (rleucg /
 * bethe_density settings); the next pair call needs fresh tables and maps */
int ucgb200_tables_clear(ucgb200_ctx *ctx);
/* tabindex[(n_formal+1)^2] (pair_table_ucgld.cpp:844,892) and
 * cutsq[(n_actual+1)^2] (init_one :886-895; only actual-type pairs are read, :213) */
int ucgb200_set_pair_maps(ucgb200_ctx *ctx, const int *tabindex, const double *cutsq);

/* ---------------------------------------------------------------- atom data */
/* AtomVecUCG arrays -> device.  nlocal owned atoms (ghosts are built on the device). */
int ucgb200_atoms_upload(ucgb200_ctx *ctx, int nlocal, const ucgb200_atoms *host, unsigned fields);
/* device -> host, in the ORIGINAL host order of the last full upload (the device
 * re-sorts atoms by cell at every neighbor rebuild; tags travel with them). */
int ucgb200_atoms_download(ucgb200_ctx *ctx, int nlocal_capacity, ucgb200_atoms *host, unsigned fields);
int ucgb200_natoms(const ucgb200_ctx *ctx, int *nlocal, int *nghost);
/* ONE timestep of the configured deck (ucgb200_deck_configure + ucgb200_setup) driven from HOST arrays: `in` (fields
 * in_fields, nlocal sites in host order) goes to the device, the step runs, `out` (fields out_fields) comes back —
 * what ucgb200_atoms_upload + ucgb200_run(ctx, 1) + ucgb200_atoms_download do, but with the device->host copies on a
 * second stream so that they overlap the kernels: x (and ucgl / ucgstate when no later stage of the deck writes them)
 * leave while the pair kernel runs, f while the fix stages run, the rest after them.  Page-lock the host arrays
 * (ucgb200_host_register / cudaHostRegister) or the copies serialise.  Returns when `out` is complete. */
int ucgb200_step_host(ucgb200_ctx *ctx, const ucgb200_atoms *in, unsigned in_fields, ucgb200_atoms *out, unsigned out_fields);
/* AtomVecUCG::force_clear (atom_vec_ucg.cpp:131-135) + stock f memset.  The pair
 * kernels overwrite f/ucgforce/scores, so this is only needed when no pair style runs. */
int ucgb200_force_clear(ucgb200_ctx *ctx);

/* ----------------------------------------------------------------- neighbor */
#define UCGB200_NEIGH_FULL 1 /* every neighbor of every owned atom (what the kernels walk) */
/* `neighbor <skin> bin`; cutneigh = sqrt(max cutsq) + skin.  cut_override > 0 forces the
 * list cutoff (fix cluster_switch requests its own list). */
int ucgb200_neigh_configure(ucgb200_ctx *ctx, double skin, double cut_override);
/* Neighbor::decide/check_distance [stock]: *rebuild = 1 iff any owned atom moved more
 * than skin/2 since the last build.  With several bricks this is the LOCAL flag; the host
 * layer reduces it with MAX over the ranks (stock LAMMPS: MPI_Allreduce in decide()). */
int ucgb200_neigh_decide(ucgb200_ctx *ctx, int *rebuild);
/* domain->pbc + cell sort + ghost construction (comm->borders) + list build */
int ucgb200_neigh_build(ucgb200_ctx *ctx);
/* comm->forward_comm(): refresh ghost x, ucgstate, ucgl, ucgp from their owners
 * (fields_comm, atom_vec_ucg.cpp:71) */
int ucgb200_ghosts_forward(ucgb200_ctx *ctx);
/* neighbor list download for parity tests: rows are in current DEVICE order;
 * tag_i[nlocal], numneigh[nlocal], offsets[nlocal+1], neigh_tags[total] (tags of the
 * neighbors), neigh_shift[total] (image code: 0 owned, 1..26 local periodic image, 32+ from
 * another brick).  Pass
 * NULL pointers with *total to query sizes. */
int ucgb200_neigh_download(ucgb200_ctx *ctx, int *nlocal, long long *total, int *tag_i, int *numneigh,
                           long long *offsets, int *neigh_tags, int *neigh_shift);
int ucgb200_neigh_stats(ucgb200_ctx *ctx, long long *total_pairs, int *max_row, int *nbuilds);

/* --------------------------------------------------------------- pair styles */
/* PairTable_UCGLD::compute (pair_table_ucgld.cpp:111-541). */
int ucgb200_pair_ucgld(ucgb200_ctx *ctx, int eflag, int vflag);
/* PairTable_UCG_Bethe::compute (pair_table_ucg_bethe.cpp:88-630).
 * method 0 = mean field, 1 = bethe; pseudo 1 = pseudo-likelihood scores, 0 = SCE;
 * prior 0 = chemical potential, 2 = ucgl (noise prior is stochastic: seed/noise). */
int ucgb200_pair_bethe(ucgb200_ctx *ctx, int eflag, int vflag, int method, int pseudo, int prior,
                       double noise_level, int seed);
/* PairTable_RLEUCG_INTERFACE::compute (pair_table_rleucg_interface.cpp:177-505).  In this style
 * atom types are STATE types (a 2-state site of base type t uses tables of types t, t+1).
 * Arrays are 1-based: actual_from_state[ntypes+1] (:649-652), n_states/use_entropy/cv_threshold/
 * threshold_radius[n_actual+1] (:618-640), chem_pot[ntypes+1] (:641-647), tabindex and
 * cutsq[(ntypes+1)^2] (coeff :736-741, init_one :802-809), mass[ntypes+1]; kT as for the
 * other styles (init_style :781-792).  Tables come from ucgb200_table_upload. */
int ucgb200_pair_rleucg_configure(ucgb200_ctx *ctx, int ntypes, const int *actual_from_state, int n_actual,
                                  const int *n_states, const int *use_entropy, const double *cv_threshold,
                                  const double *threshold_radius, const double *chem_pot, const int *tabindex,
                                  const double *cutsq, const double *mass, double kT);
int ucgb200_pair_rleucg(ucgb200_ctx *ctx, int eflag, int vflag);
/* substate_probability[i][0] and the CV force F_p*dp/drho of the last evaluation (host order) */
int ucgb200_pair_rleucg_probabilities(ucgb200_ctx *ctx, int cap, double *prob, double *cvforce);
/* PairTable_UCG_Bethe_Density::compute (pair_table_ucg_bethe_density.cpp:133-758), repaired
 * semantics (SURVEY Q9-Q12,Q14; oracle/repair_bethe_density.py lists every deviation).  Types,
 * formal-type maps, chemical potentials, tables, cutsq and kT come from ucgb200_set_types exactly as
 * for table_ucgld; this call adds the per-actual-type density settings read_state_settings parses
 * (:827-880), all arrays 1-based [n_actual+1]: use_density ("density" keyword), use_entropy
 * ("entropy" / "no_entropy"), cv_threshold and threshold_radius.  Requires newton off (init_style
 * :1151) — the host class enforces it.  The style writes atom->f and atom->ucgp. */
int ucgb200_pair_bethe_density_configure(ucgb200_ctx *ctx, int n_actual, const int *use_density,
                                         const int *use_entropy, const double *cv_threshold,
                                         const double *threshold_radius);
int ucgb200_pair_bethe_density(ucgb200_ctx *ctx, int eflag, int vflag);
/* prior_prob[i][0] and sum_s prior_prob_force*prior_prob_partial of the last evaluation (host order) */
int ucgb200_pair_bethe_density_priors(ucgb200_ctx *ctx, int cap, double *prob0, double *cvforce);
/* eng_vdwl, virial[6] (xx,yy,zz,xy,xz,yz) of the last pair call with eflag/vflag */
int ucgb200_pair_energy_virial(ucgb200_ctx *ctx, double *eng_vdwl, double virial[6]);
/* Per-atom energy and virial of the last pair call that asked for them: LAMMPS' flag bits are honoured, eflag & 2 =
 * ENERGY_ATOM, vflag & 4 = VIRIAL_ATOM (what compute pe/atom and compute stress/atom set through Pair::ev_setup).
 * eatom[i] = sum_j E_ij / 2, vatom[i][0..5] = sum_j (d x d) fpair / 2 in LAMMPS' xx yy zz xy xz yz order — each site's
 * own half of every pair it is in, which is what [stock] Pair::ev_tally (pair_table_ucgld.cpp:531-533) leaves on the
 * owners after the reverse communication of the compute.  Host order; either pointer may be NULL.
 * Implemented for all four pair calls.  The two density styles run with newton off: there [stock] ev_tally gives half
 * of every visit to the centre site and half to a LOCAL partner (pair_table_rleucg_interface.cpp:439, 488,
 * pair_table_ucg_bethe_density.cpp:407, 510, 647, 729), and the virial of the CV back-force sweep is added. */
int ucgb200_pair_peratom(ucgb200_ctx *ctx, int nlocal_capacity, double *eatom /*[n]*/, double *vatom /*[6n]*/);

/* --------------------------------------------------------------------- fixes */
/* FixNVE_UCGLD::initial_integrate (fix_nve_ucgld.cpp:44-101); wall != 0 adds the
 * ucgstate = (lambda < 0.5 ? 0 : 1) assignment of FixNVE_UCGLD_Wall_Hard
 * (fix_nve_ucgld_wall_hard.cpp:125-131). */
int ucgb200_fix_nve_initial(ucgb200_ctx *ctx, double dtv, double dtf, int groupbit, int wall);
/* ::final_integrate (fix_nve_ucgld.cpp:104-153); wall != 0 adds the hard-wall
 * reflection (fix_nve_ucgld_wall_hard.cpp:194-200). */
int ucgb200_fix_nve_final(ucgb200_ctx *ctx, double dtf, int groupbit, int wall);
/* FixNVE_UCGLD_Wall_Hard::post_force bias (fix_nve_ucgld_wall_hard.cpp:234-257) */
int ucgb200_fix_wall_bias(ucgb200_ctx *ctx, double barrier, int groupbit);
/* FixUCGState::post_force (fix_ucgstate.cpp:88-132). mode 0 = deterministic round,
 * 1 = ld (probabilities only), 2 = mc(seed, rate).  step = update->ntimestep (RNG counter). */
int ucgb200_fix_ucgstate(ucgb200_ctx *ctx, int mode, int seed, double rate, long long step);
/* Fix_UCGLD_Langevin::post_force (fix_ucgld_langevin.cpp:226-297) for the current
 * t_target; gamma1/gamma2 per ACTUAL type index 1..ntypes as computed in init()
 * (:164-171).  step feeds the counter-based RNG. */
int ucgb200_fix_langevin(ucgb200_ctx *ctx, const double *gfactor1, const double *gfactor2, int ntypes,
                         double tsqrt, int seed, long long step, int groupbit, int zero_v_skip);
/* Fix_UCGLD_Langevin::end_of_step lambda temperature (:303-312), globally reduced
 * on this context: returns sum 0.5*ml*vl^2*mvv2e and the group count. */
int ucgb200_lambda_ke(ucgb200_ctx *ctx, int groupbit, double *ke_sum, long long *count);
/* particle kinetic energy sum 0.5*m*v^2*mvv2e (thermo) */
int ucgb200_kinetic_energy(ucgb200_ctx *ctx, int groupbit, double *ke_sum, long long *count);

/* FixClusterSwitch (fix_cluster_switch.cpp).  _configure = the constructor (:36-170): arguments
 * of the fix line (mol_seed, mol_offset, cutoff, seed), rates file (probON; the switch types and
 * their ON / OFF atom types, :252-275) and contact file (ordered type pairs, :338-349); it scans the
 * atoms already on the device for maxmol, nSwitchPerMol and the initial mol_state / mol_restrict.
 * _check = check_cluster (:537-731) on the current FULL list: connected molecules (contact map,
 * rsq < cutoff^2, offset partners tied) get the smallest label of their cluster; molecules in the
 * cluster of mol_seed are restricted and set ON.  _switch = attempt_switch (:733-839): molecules
 * outside it flip ON<->OFF with probability probON / 1-probON, drawn from the same RanPark stream in
 * the same (ascending molecule id) order as the reference, then the atom types are rewritten.
 * Single brick only. */
int ucgb200_cluster_configure(ucgb200_ctx *ctx, int mol_seed, int mol_offset, double cutoff, int seed,
                              double prob_on, int n_switch_types, const int *type_on, const int *type_off,
                              int n_contacts, const int *contact_pairs /*[2*n_contacts]*/, int ntypes, int groupbit);
int ucgb200_cluster_check(ucgb200_ctx *ctx, int *n_cluster);
int ucgb200_cluster_switch(ucgb200_ctx *ctx, int *n_attempts, int *n_success);
/* compute_vector (:923-933): out[0..6] = nAttemptsTotal, nSuccessTotal, nAttemptsON, nAttemptsOFF,
 * nSuccessON, nSuccessOFF, nCluster; out[7] = labelling rounds of the last check */
int ucgb200_cluster_stats(ucgb200_ctx *ctx, double out[8]);
/* Fix::next_reneighbor of fix cluster_switch for the resident loop (fix_cluster_switch.cpp:71, 480) */
int ucgb200_cluster_next_reneighbor(ucgb200_ctx *ctx, long long step);
/* per-molecule arrays [0..maxmol] (cluster_assignment.log / state_assignment.log, :711-727) */
int ucgb200_cluster_get(ucgb200_ctx *ctx, int cap, int *mol_cluster, int *mol_state, int *mol_restrict,
                        int *mol_accept, int *max_mol);

/* ---------------------------------------------------- resident run (no host data) */
/* deck description for the resident integrator: which fixes are active, in
 * fix-definition order semantics of SURVEY §3.1. */
typedef struct ucgb200_deck {
  int pair_style;        /* 0 ucgld, 1 bethe, 2 rleucg, 3 bethe_density */
  int nve;               /* 1 = fix nve/ucgld, 2 = fix nve/ucgld/wall/hard */
  int nve_groupbit;
  int wall_bias;         /* bias_potential flag */
  double wall_barrier;
  int langevin;          /* fix ucgld/langevin present */
  double t_start, t_stop, t_period;
  int langevin_seed;
  int langevin_groupbit;
  int ucgstate;          /* 0 = absent, 1 = deterministic, 2 = ld, 3 = mc */
  int ucgstate_seed;
  double ucgstate_rate;
  int bethe_method, bethe_pseudo, bethe_prior;
  int thermo_every;      /* eflag/vflag on steps that are multiples of this (0 = never) */
  int cluster_freq;      /* fix cluster_switch rateFreq (0 = absent); configure it with ucgb200_cluster_configure */
  /* order in which the post_force stages act = the deck's fix definition order ([stock] Modify::post_force): decimal
   * digits, first stage in the highest place, 1 = ucgld/langevin, 2 = ucgstate, 3 = wall bias of nve/ucgld/wall/hard;
   * 0 = 123.  Matters when the wall fix (with bias_potential) is defined before a non-ld fix ucgstate, which
   * overwrites the ucgl the bias reads (fix_nve_ucgld_wall_hard.cpp:234-239, fix_ucgstate.cpp:130). */
  int post_force_order;
  /* table_ucg_bethe `prior chemical_potential noise <level> <seed>` (bethe_prior = 1, pair_table_ucg_bethe.cpp:189-194):
   * acts on the first evaluation only (sites whose ucgp is still -1); one Philox draw per site, keyed by (seed, tag) */
  double bethe_noise_level;
  int bethe_seed;
  /* fix ucgld/langevin with a bias temperature compute (fix_ucgld_langevin.cpp:174-177): the reference's Tp_BIAS
   * branch differs from the plain one in a single rule, no random force on a site whose lambda velocity is exactly
   * zero (:285; remove_bias / restore_bias are commented out there) */
  int langevin_bias;
  int reserved[2];
} ucgb200_deck;
int ucgb200_deck_configure(ucgb200_ctx *ctx, const ucgb200_deck *deck);
/* Verlet::setup [stock]: pbc, build, force_clear, pair->compute, fix setup() calls */
int ucgb200_setup(ucgb200_ctx *ctx);
/* Verlet::run(n) [stock] with every stage on the device; ntimestep continues from
 * the last call; beginstep/endstep of the run are (current, current+n). */
int ucgb200_run(ucgb200_ctx *ctx, int nsteps);
/* update->ntimestep at the start of a run (reset_timestep, read_dump, a host driver taking over) */
int ucgb200_set_ntimestep(ucgb200_ctx *ctx, long long ntimestep);
/* `run N start S stop E`: nsteps steps of a run spanning beginstep..endstep (the span only matters to ramps such
 * as the target temperature of fix ucgld/langevin): a run cut into pieces at dump steps stays bit-identical */
int ucgb200_run_between(ucgb200_ctx *ctx, int nsteps, long long beginstep, long long endstep);
/* thermo scalars of the last thermo step: out[0]=eng_vdwl, out[1..6]=virial,
 * out[7]=particle KE sum, out[8]=lambda KE sum, out[9]=lambda temperature,
 * out[10]=ntimestep, out[11]=rebuild count, out[12]=nlocal, out[13]=nghost */
int ucgb200_thermo(ucgb200_ctx *ctx, double out[16]);
/* sticky device error word: code (UCGB200_ERR_*), tags of the pair, rsq */
int ucgb200_status(ucgb200_ctx *ctx, int *code, int *tag_i, int *tag_j, double *rsq);
/* the same word WITHOUT clearing it (ucgb200_status clears a non-zero word once it has been read): lets a caller test
 * the code first and fetch the pair's tags and distance with ucgb200_status afterwards */
int ucgb200_status_peek(ucgb200_ctx *ctx, int *code, int *tag_i, int *tag_j, double *rsq);
/* per-stage CUDA-event timers of the resident loop (ms since the last reset): pair, neigh (decide + rebuild), comm
 * (ghost refresh / halo), modify (fix stages).  enable: -1 read only; 0 off; 1 on, blocking (one event synchronisation
 * per stage: the loop loses its speculative pair launch while it is being timed); 2 = 1 + reset; 3 on + reset,
 * NON-blocking: event pairs from a ring, resolved when the totals are read, so the loop keeps the schedule it has when
 * it is not being timed.  ucgb200_last_pair_ms (pair kernel alone) needs any of the "on" modes. */
int ucgb200_timers(ucgb200_ctx *ctx, int enable, double out_ms[4], long long out_launches[4]);
/* duration of the last pair kernel launch in ms (events on the context stream) */
int ucgb200_last_pair_ms(ucgb200_ctx *ctx, double *ms);

/* ---------------------------------------------------- multi-GPU halo plumbing */
/* Brick decomposition (what stock Comm does over MPI for the reference, with the payloads of
 * AtomVecUCG's field lists, UCG/atom_vec_ucg.cpp:66-82).  Rank r owns brick
 * (r % gx, (r/gx) % gy, r/(gx*gy)) of the box.  Ghosts within cut+skin of a brick face come
 * from the brick on the other side — or are local periodic images when that brick is the
 * rank itself.  Every exchange is ONE message per peer rank (NVSwitch makes all peers equal:
 * no staged x/y/z relay); records for rank k are contiguous and ordered by k in the send
 * buffer, the receive buffer is ordered by source rank.  The pack/unpack kernels work on
 * DEVICE buffers so that any transport can carry them (NCCL all-to-all driven by the host
 * layer, or peer-mapped stores).  The device list is full, so there is no reverse halo. */
int ucgb200_halo_configure(ucgb200_ctx *ctx, int rank, int nranks, const int procgrid[3]);
int ucgb200_halo_info(const ucgb200_ctx *ctx, int *rank, int *nranks);
/* record sizes in bytes: border (rebuild), forward (every step), migrate (rebuild) */
int ucgb200_halo_record_bytes(int *border, int *forward, int *migrate);
/* comm->exchange(): wrap, count the sites that left this brick per destination rank ... */
int ucgb200_migrate_prepare(ucgb200_ctx *ctx, int *counts /*[nranks]*/);
/* ... pack them (and compact the stayers) ... */
int ucgb200_migrate_pack(ucgb200_ctx *ctx, void *d_sendbuf);
/* ... append the arrivals. */
int ucgb200_migrate_unpack(ucgb200_ctx *ctx, const void *d_recvbuf, int nrecv);
/* rebuild, local part: cell sort + image lists; then the border records per destination */
int ucgb200_neigh_build_local(ucgb200_ctx *ctx);
int ucgb200_halo_send_counts(ucgb200_ctx *ctx, int *counts /*[nranks]*/);
int ucgb200_halo_pack_border(ucgb200_ctx *ctx, void *d_sendbuf);
int ucgb200_halo_unpack_border(ucgb200_ctx *ctx, const void *d_recvbuf, const int *recv_counts /*[nranks]*/);
/* rebuild, final part: bin ghosts, build rows */
int ucgb200_neigh_build_finish(ucgb200_ctx *ctx);
/* comm->forward_comm() across bricks, every non-rebuild step (same counts as the borders) */
int ucgb200_halo_pack_forward(ucgb200_ctx *ctx, void *d_sendbuf);
int ucgb200_halo_unpack_forward(ucgb200_ctx *ctx, const void *d_recvbuf);
/* device pointer of the rebuild flag (int32) so the host layer can all-reduce it (MAX) */
int ucgb200_neigh_flag_ptr(ucgb200_ctx *ctx, void **d_flag);

/* Resident multi-brick runs: with a communicator attached, ucgb200_setup / ucgb200_run drive the
 * exchanges above themselves — on the context stream: the per-step ghost refresh as direct stores into the peers'
 * mapped memory (fallback: NCCL send/recv groups plus a 4-byte ncclAllReduce(MAX) for the rebuild decision), the
 * rebuild (migration, borders) as NCCL send/recv groups — so a step needs no host-side orchestration and exactly one
 * host synchronisation.  Rank 0 creates the id (128 bytes, the
 * ncclUniqueId), the host layer broadcasts it by whatever means it has (MPI_Bcast inside LAMMPS,
 * torch.distributed in the tests), every rank calls _comm_init after ucgb200_halo_configure.
 * NCCL is loaded at run time (libnccl.so.2); _comm_unique_id returns -4 when it is absent. */
int ucgb200_comm_unique_id(char *id, int len);
int ucgb200_comm_init(ucgb200_ctx *ctx, const char *id, int len);
int ucgb200_comm_destroy(ucgb200_ctx *ctx);
int ucgb200_comm_stats(ucgb200_ctx *ctx, long long *bytes_forward, int *nrebuilds, int *send_records);
/* how the forward halo of the current lists travels: *peer_mapped = 1 when every brick stores its records straight
 * into its peers' ghost staging through CUDA-IPC-mapped memory over NVLink (the default on one NVSwitch node; the flag
 * and displacement reductions of Neighbor::decide ride in the same push), 0 for NCCL send/recv groups (UCGB200_P2P=0,
 * or the mapping could not be established).  *pushes counts the peer-mapped exchanges so far. */
int ucgb200_comm_transport(ucgb200_ctx *ctx, int *peer_mapped, long long *pushes);

/* ------------------------------------------------------ dump / read_dump taps */
/* Column codes: the `dump custom` keywords the reference's patched dump_custom.cpp adds
 * (ucgstate ucgl ucgp, dump_custom.cpp:1672-1687, 3552-3577), the stock per-atom keywords around
 * them, and the `compute property/atom` names of AtomVecUCG::property_atom
 * (UCG/atom_vec_ucg.cpp:172-181; those columns are zero outside the COMPUTE's group, :184-231).
 * The same codes name the fields of a read_dump snapshot (reader.h:24-26, read_dump.cpp:1326-1352). */
#define UCGB200_COL_ID 0
#define UCGB200_COL_MOL 1
#define UCGB200_COL_TYPE 2
#define UCGB200_COL_MASS 3
#define UCGB200_COL_X 4
#define UCGB200_COL_Y 5
#define UCGB200_COL_Z 6
#define UCGB200_COL_XS 7
#define UCGB200_COL_YS 8
#define UCGB200_COL_ZS 9
#define UCGB200_COL_VX 10
#define UCGB200_COL_VY 11
#define UCGB200_COL_VZ 12
#define UCGB200_COL_FX 13
#define UCGB200_COL_FY 14
#define UCGB200_COL_FZ 15
#define UCGB200_COL_UCGSTATE 16
#define UCGB200_COL_UCGL 17
#define UCGB200_COL_UCGP 18
#define UCGB200_COL_PROC 19
#define UCGB200_COL_Q 20            /* atom_style ucg carries q but no UCG style reads it: always 0 */
#define UCGB200_COL_P_UCGSTATE 21   /* compute property/atom ... (index 0..5 of property_atom) */
#define UCGB200_COL_P_UCGL 22
#define UCGB200_COL_P_UCGFORCE 23
#define UCGB200_COL_P_UCGVL 24
#define UCGB200_COL_P_UCGP 25
#define UCGB200_COL_P_UCGML 26
#define UCGB200_COL_COUNT 27
#define UCGB200_DUMP_MAXCOL 32
#define UCGB200_DUMP_MAXTHRESH 8
/* dump_modify thresh operators, enum{LT,LE,GT,GE,EQ,NEQ,XOR} dump_custom.cpp:50 */
#define UCGB200_THRESH_LT 0
#define UCGB200_THRESH_LE 1
#define UCGB200_THRESH_GT 2
#define UCGB200_THRESH_GE 3
#define UCGB200_THRESH_EQ 4
#define UCGB200_THRESH_NEQ 5
#define UCGB200_THRESH_XOR 6
/* row order: the host index order an unsorted serial dump has, or `dump_modify sort id` */
#define UCGB200_DUMP_ORDER_INDEX 0
#define UCGB200_DUMP_ORDER_ID 1

typedef struct ucgb200_dump_spec {
  int ncols;
  const int *cols;          /* [ncols] UCGB200_COL_* */
  const int *col_groupbit;  /* [ncols] group bit of the compute behind a P_* column (NULL = all) */
  int groupbit;             /* the dump's group */
  int nthresh;              /* dump_modify thresh: attribute (a non-P column code), operator, value */
  const int *thresh_col;
  const int *thresh_op;
  const double *thresh_value;
  int order;                /* UCGB200_DUMP_ORDER_* */
} ucgb200_dump_spec;

/* DumpCustom::count(): number of owned atoms in the group that pass every threshold (dump_custom.cpp:721-1368) */
int ucgb200_dump_count(ucgb200_ctx *ctx, const ucgb200_dump_spec *spec, long long *nrows);
/* DumpCustom::count() + pack() (+ Dump::sort): selection, ordering and the column gather run on the
 * device; buf[row * ncols + col] receives what DumpCustom::buf holds (ints as exactly converted doubles) */
int ucgb200_dump_pack(ucgb200_ctx *ctx, const ucgb200_dump_spec *spec, double *buf, long long capacity_rows,
                      long long *nrows);
/* The same rows as TEXT, formatted on the device exactly as DumpCustom::convert_string / write_lines
 * do with the default formats (dump_custom.cpp:1388-1468: "%d" for id mol type proc ucgstate, "%g" for the
 * rest, one blank between columns, none before the newline).  Only nbytes characters cross PCIe. */
int ucgb200_dump_text(ucgb200_ctx *ctx, const ucgb200_dump_spec *spec, char *text, long long capacity_bytes,
                      long long *nrows, long long *nbytes);
/* text == NULL above formats on the device and reports the size; this fetches the characters */
int ucgb200_dump_text_copy(ucgb200_ctx *ctx, char *text, long long capacity_bytes);
/* ReadDump::process_atoms in replace mode + the remap of migrate_atoms_by_coords (read_dump.cpp:797-935,
 * 1150-1163): fields[i * nfield + j] is the snapshot as Reader::read_atoms leaves it, fieldtype[0] must be
 * UCGB200_COL_ID; rows whose id matches an owned atom overwrite x y z (unscaled with the SNAPSHOT box when
 * `scaled`), vx vy vz, fx fy fz, ucgstate, ucgl, ucgp of that atom; positions are then wrapped into the
 * CURRENT box (ucgb200_set_box first for `box yes`).  fields == NULL takes the block ucgb200_snapshot_parse left on
 * the device.  updated[nlocal] (host index order, may be NULL)
 * receives the updateflag array `trim` works from; *nreplace counts the matches. */
int ucgb200_atoms_update_by_tag(ucgb200_ctx *ctx, int nnew, int nfield, const int *fieldtype, const double *fields,
                                int scaled, const double snap_lo[3], const double snap_hi[3], int *updated,
                                long long *nreplace);
/* ReaderNative::read_atoms (reader_native.cpp:486-500) on the device: `text` holds the nrows atom lines of one
 * snapshot (nwords columns each); column fieldindex[m] of every line becomes fields[row][m] of a block that stays
 * on the device for ucgb200_atoms_update_by_tag(..., fields = NULL, ...).  Decimal tokens that convert exactly in one
 * IEEE operation (<= 2^53 mantissa, |power of ten| <= 22: everything "%g" prints up to e+-22) are converted
 * there; rows holding any other token are reported (row index, byte offset of the line) so that the caller
 * converts them with strtod and sends them back with ucgb200_snapshot_patch.  Returns UCGB200_PARSE_ON_HOST when the
 * block should be converted by the caller instead (>= 2 GiB of text, > 64 columns, more than slow_cap such rows). */
#define UCGB200_PARSE_ON_HOST 7
int ucgb200_snapshot_parse(ucgb200_ctx *ctx, const char *text, long long nbytes, long long nrows, int nwords, int nfield,
                           const int *fieldindex, int slow_cap, int *slow_rows, int *slow_offsets, long long *nslow);
int ucgb200_snapshot_patch(ucgb200_ctx *ctx, int n, const int *rows, const double *values /* [n][nfield] */);
/* Domain::remap of every owned atom into the current periodic box, any number of periods
 * (ReadDump::migrate_atoms_by_coords, read_dump.cpp:1150-1163); invalidates the neighbor list */
int ucgb200_atoms_remap(ucgb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* UCGB200_H */
