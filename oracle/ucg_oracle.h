/* ucg_oracle.h — CPU oracle for the UCG hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's algorithms (KJAdams2000/LAMMPS-UCG-dev,
 * UCG/*.cpp) plus the stock-LAMMPS machinery they sit on (Verlet call order,
 * binned neighbor lists with periodic ghosts, skin check, ev_tally, RanMars).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libucgb200.so) never does.
 *
 * Pinning: the restatement is checked against the reference's own UCG/*.cpp compiled
 * verbatim against a minimal LAMMPS-API shim (oracle/_ref, see oracle/Makefile) by
 * tests/test_oracle_vs_ref.py and against the golden vectors that build produced
 * (tests/golden/).  The reference itself ships no tests or golden vectors
 * (SURVEY.md §4), and stock LAMMPS is not available here, so everything tagged
 * [stock] below is restated from the published LAMMPS algorithms.
 */
#ifndef UCG_ORACLE_H
#define UCG_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_sys orc_sys;

enum { ORC_LOOKUP = 0, ORC_LINEAR = 1, ORC_SPLINE = 2, ORC_BITMAP = 3 };
enum { ORC_RNONE = 0, ORC_RLINEAR = 1, ORC_RSQ = 2, ORC_BMP = 3 };

orc_sys *orc_create(void);
void orc_destroy(orc_sys *s);
const char *orc_error(orc_sys *s); /* last error->one/all text, "" if none */

void orc_set_units(orc_sys *s, double boltz, double ftm2v, double mvv2e);
void orc_set_box(orc_sys *s, const double lo[3], const double hi[3]);
void orc_set_dt(orc_sys *s, double dt);
void orc_set_special_lj(orc_sys *s, const double special_lj[4]);
void orc_set_newton(orc_sys *s, int newton_pair);

/* pair_style table_ucgld/table_ucg_bethe <style> <N> + state settings (1-based arrays) */
void orc_pair_style(orc_sys *s, int tabstyle, int tablength);
void orc_set_types(orc_sys *s, int n_actual, int n_formal, const int *n_states,
                   const int *formal_from_actual, const double *chem_pot, const double *mass);
/* read_table + coeff-time checks + spline_table + compute_table; returns table index */
int orc_table_add_file(orc_sys *s, const char *file, const char *keyword, double cut);
int orc_table_add_arrays(orc_sys *s, int ninput, int rflag, double rlo, double rhi, int fpflag,
                         double fplo, double fphi, const double *rfile, const double *efile,
                         const double *ffile, double cut);
int orc_table_len(orc_sys *s, int idx);
/* which: 0 rsq, 1 e, 2 f, 3 de, 4 df, 5 e2, 6 f2, 7 drsq; returns count copied */
int orc_table_get(orc_sys *s, int idx, int which, double *out);
/* params: innersq, delta, invdelta, deltasq6, cut, nmask, nshiftbits, match */
void orc_table_params(orc_sys *s, int idx, double out[8]);
/* pair_coeff I J Ns_i Ns_j tables[Ns_i*Ns_j] (coeff(), pair_table_ucgld.cpp:719-865) */
int orc_pair_coeff(orc_sys *s, int ilo, int ihi, int jlo, int jhi, int ns_i, int ns_j, const int *tables);
/* Pair::init [stock] -> init_one for all i<=j formal types */
int orc_pair_init(orc_sys *s);
void orc_get_pair_maps(orc_sys *s, int *tabindex, double *cutsq); /* (n_formal+1)^2 each */
void orc_set_kT(orc_sys *s, double kT);

/* atoms (AoS like LAMMPS) */
void orc_set_atoms(orc_sys *s, int n, const double *x, const double *v, const int *type,
                   const int *mask, const int *tag, const int *molecule, const int *ucgstate,
                   const double *ucgl, const double *ucgvl, const double *ucgml, const double *ucgp);
int orc_nlocal(orc_sys *s);
int orc_nghost(orc_sys *s);
/* any pointer may be NULL; arrays sized nlocal (x,v,f: 3n; scores: 2n) */
void orc_get_atoms(orc_sys *s, double *x, double *v, double *f, int *type, int *tag, int *ucgstate,
                   double *ucgl, double *ucgvl, double *ucgp, double *ucgforce, double *scores,
                   int *num_ucgstates);
void orc_get_ghosts(orc_sys *s, double *x, int *tag);

/* neighbor [stock]: `neighbor skin bin`, half (newton) or full list */
void orc_neigh_config(orc_sys *s, double skin, int full);
void orc_pbc(orc_sys *s);
void orc_borders(orc_sys *s);
void orc_neigh_build(orc_sys *s);
int orc_neigh_decide(orc_sys *s);
void orc_forward_comm(orc_sys *s);
void orc_reverse_comm(orc_sys *s);
long long orc_neigh_total(orc_sys *s);
/* pairs as (tag_i, tag_j) with i = row owner; returns count */
long long orc_neigh_pairs(orc_sys *s, int *tag_i, int *tag_j);
/* use an externally supplied list on owned+ghost local indices */
void orc_force_clear(orc_sys *s);

/* pair styles */
void orc_pair_ucgld(orc_sys *s, int eflag, int vflag);
/* method 0 mf / 1 bethe; pseudo; prior 0 chem-pot, 1 chem-pot+noise, 2 ucgl */
void orc_pair_bethe_config(orc_sys *s, int method, int pseudo, int prior, double noise, int seed);
void orc_pair_bethe(orc_sys *s, int eflag, int vflag);
double orc_eng_vdwl(orc_sys *s);
/* per-pair tally (what ev_tally would add with vflag_global), xx yy zz xy xz yz */
void orc_virial(orc_sys *s, double v[6]);

/* fixes, in definition order */
enum { ORC_FIX_TTARGET = 0, ORC_FIX_NVE = 1, ORC_FIX_NVE_WALL = 2, ORC_FIX_LANGEVIN = 3, ORC_FIX_UCGSTATE = 4 };
void orc_fix_clear(orc_sys *s);
void orc_fix_ttarget(orc_sys *s, double T); /* stub thermostat exporting t_target only */
void orc_fix_nve(orc_sys *s, int groupbit);
void orc_fix_nve_wall(orc_sys *s, int groupbit, int bias_flag, double barrier);
void orc_fix_langevin(orc_sys *s, int groupbit, double t_start, double t_stop, double t_period, int seed);
void orc_fix_ucgstate(orc_sys *s, int mode /*0 det,1 ld,2 mc*/, int seed, double rate);
double orc_lambda_temp(orc_sys *s);

/* single fix operations on the current arrays (teacher forcing) */
void orc_nve_initial(orc_sys *s, int groupbit, int wall);
void orc_nve_final(orc_sys *s, int groupbit, int wall);
void orc_wall_bias(orc_sys *s, int groupbit, double barrier);
void orc_ucgstate_post_force(orc_sys *s, int mode, double rate);
void orc_set_forces(orc_sys *s, const double *f, const double *ucgforce, const double *scores);

/* Verlet [stock] */
void orc_setup(orc_sys *s, int eflag, int vflag);
void orc_run(orc_sys *s, int nsteps, int thermo_every);
long long orc_ntimestep(orc_sys *s);
int orc_nbuilds(orc_sys *s);
/* timing breakdown of orc_run in seconds: pair, neigh, comm, modify */
void orc_timers(orc_sys *s, double out[4]);

/* RNG known-answer access */
void orc_ranmars_fill(int seed, int n, double *out);
void orc_ranpark_fill(int seed, int n, double *out);

#ifdef __cplusplus
}
#endif
#endif
