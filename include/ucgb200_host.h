/* ucgb200_host.h — host-side set-up helpers of libucgb200.so (no GPU needed).
 *
 * These are the setup-time halves of the reference's pair styles, kept on the host
 * exactly as the reference does: bit-identical tables are a
 * precondition of force parity.  The LAMMPS style classes in lammps-ucg-dev_b200/host/
 * call them from settings()/coeff()/init_one(); tests call them through ctypes.
 *
 *   table building   read_table / param_extract / spline_table / compute_table / spline /
 *                    splint            UCG/pair_table_ucgld.cpp:897-1017, 1047-1344, 1375-1428
 *   state settings   read_state_settings    UCG/pair_table_ucgld.cpp:565-652
 *   table->type map  coeff() / init_one()    UCG/pair_table_ucgld.cpp:753-865, 886-895
 */
#ifndef UCGB200_HOST_H
#define UCGB200_HOST_H
#include "ucgb200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ucgb200_table ucgb200_table;
typedef struct ucgb200_statemap ucgb200_statemap;

/* rflag of the table file header: NONE, R, RSQ, BITMAP (param_extract :1080-1086) */
#define UCGB200_R_NONE 0
#define UCGB200_R_LINEAR 1
#define UCGB200_R_RSQ 2
#define UCGB200_R_BMP 3

/* Read section `keyword` of a LAMMPS table file and build the run-time table for
 * (tabstyle, tablength, cut); cut < 0 selects the table's own upper end.  Error text (the reference's error->one/all message) goes
 * to errbuf.  Returns 0 or -1. */
int ucgb200_host_table_from_file(const char *file, const char *keyword, double cut, int tabstyle,
                                 int tablength, ucgb200_table **out, char *errbuf, int errlen);
int ucgb200_host_table_from_arrays(int ninput, int rflag, double rlo, double rhi, int fpflag, double fplo,
                                   double fphi, const double *rfile, const double *efile,
                                   const double *ffile, double cut, int tabstyle, int tablength,
                                   ucgb200_table **out, char *errbuf, int errlen);
/* params: innersq, delta, invdelta, deltasq6, cut, nmask, nshiftbits, match; *n = entries */
int ucgb200_host_table_info(const ucgb200_table *t, double params[8], int *n);
/* which: 0 rsq, 1 e, 2 f, 3 de, 4 df, 5 e2, 6 f2, 7 drsq; returns number of values copied */
int ucgb200_host_table_array(const ucgb200_table *t, int which, double *out, int cap);
/* Pair::single for one table (pair_table_ucgld.cpp:1474-1520): returns 0 or UCGB200_ERR_TABLE_* */
int ucgb200_host_table_single(const ucgb200_table *t, double rsq, double factor_lj, double *phi, double *fforce);
void ucgb200_host_table_free(ucgb200_table *t);
int ucgb200_host_table_upload(ucgb200_ctx *ctx, const ucgb200_table *t, int *index);

/* read_state_settings(): "n_actual n_formal max_states" then per actual type
 * "<type> <nstates>" [+ "<formal0> <formal1>" + "<mu0> <mu1>" for 2-state types] */
int ucgb200_host_statemap_from_file(const char *file, ucgb200_statemap **out, char *errbuf, int errlen);
int ucgb200_host_statemap_create(int n_actual, int n_formal, const int *n_states,
                                 const int *formal_from_actual, const double *chem_pot,
                                 ucgb200_statemap **out, char *errbuf, int errlen);
void ucgb200_host_statemap_free(ucgb200_statemap *m);
int ucgb200_host_statemap_sizes(const ucgb200_statemap *m, int *n_actual, int *n_formal);
/* pair_coeff ilo*ihi jlo*jhi Ns_i Ns_j {table cut}x(Ns_i*Ns_j): tables[] are indices as
 * returned by table_upload, cuts[] their cutoffs */
int ucgb200_host_statemap_coeff(ucgb200_statemap *m, int ilo, int ihi, int jlo, int jhi, int ns_i, int ns_j,
                                const int *tables, const double *cuts, char *errbuf, int errlen);
/* Pair::init: init_one(i,j) for all i<=j formal types ("All pair coeffs are not set") */
int ucgb200_host_statemap_init(ucgb200_statemap *m, char *errbuf, int errlen);
/* copies n_states[n_actual+1], formal_from_actual[2(n_actual+1)], chem_pot[n_formal+1],
 * tabindex[(n_formal+1)^2], cutsq[(n_formal+1)^2]; any pointer may be NULL */
int ucgb200_host_statemap_get(const ucgb200_statemap *m, int *n_states, int *formal_from_actual,
                              double *chem_pot, int *tabindex, double *cutsq);
/* set_types + set_pair_maps on a context; mass[1..n_formal] = atom->mass */
int ucgb200_host_statemap_apply(ucgb200_ctx *ctx, const ucgb200_statemap *m, const double *mass);

/* ------------------------------------------------- dump custom / read_dump / read_data
 * Host halves of the taps around the resident step, with the reference's grammar and file formats
 * (dump_custom.cpp, read_dump.cpp, reader_native.cpp — the patched stock files of the reference tree — and the
 * data-file columns of UCG/atom_vec_ucg.cpp:85-90).  The rows themselves are selected, ordered, gathered
 * and — with the default formats — turned into text on the device (ucgb200_dump_pack / _dump_text). */
typedef struct ucgb200_dump ucgb200_dump;
typedef struct ucgb200_data ucgb200_data;

/* `dump ID group custom N file col ...`: arg[0..narg) are the words after "dump"; groupbit is the bit of `group`.
 * Columns: id mol type mass x y z xs ys zs vx vy vz fx fy fz q proc ucgstate ucgl ucgp, c_ID / c_ID[k]. */
int ucgb200_host_dump_create(int narg, const char *const *arg, int groupbit, ucgb200_dump **out, char *errbuf, int errlen);
void ucgb200_host_dump_free(ucgb200_dump *d);
/* `compute ID group property/atom name ...` (names of AtomVecUCG::property_atom, atom_vec_ucg.cpp:172-181):
 * resolves the dump's c_ID / c_ID[k] columns, as DumpCustom::init_style does by compute ID */
int ucgb200_host_dump_bind_compute(ucgb200_dump *d, const char *id, int groupbit, int nvalues, const char *const *names,
                                   char *errbuf, int errlen);
/* `dump_modify ID keyword value ...` (words after the ID): append buffer every flush header pad time units,
 * sort off|id, format line|int|float|M|none, thresh attribute op value | none */
int ucgb200_host_dump_modify(ucgb200_dump *d, int narg, const char *const *arg, char *errbuf, int errlen);
/* Dump::write() of one snapshot of the context's resident atoms */
int ucgb200_host_dump_write(ucgb200_dump *d, ucgb200_ctx *ctx, long long ntimestep, double time, const char *unit_style,
                            char *errbuf, int errlen);
/* `run N` of a resident deck (ucgb200_deck_configure + ucgb200_setup done) with dumps, as [stock] Output schedules
 * them: every dump writes at the start of the run if the current step is a multiple of its N, then on every multiple;
 * between dump steps the steps run on the device without touching the host (ucgb200_run_between) */
int ucgb200_host_run(ucgb200_ctx *ctx, long long nsteps, int ndump, ucgb200_dump *const *dumps, double dt, const char *unit_style,
                     char *errbuf, int errlen);
int ucgb200_host_dump_stats(const ucgb200_dump *d, long long *rows, long long *bytes, int *nevery);

/* `read_dump file Nstep field ... keyword value ...` on the resident atoms (native text dump files):
 * fields x y z vx vy vz fx fy fz q ucgstate ucgl ucgp; keywords box timestep replace trim label scaled wrapped
 * format native (purge / add are not supported).  stats[7] = atoms before, in snapshot, purged, replaced,
 * trimmed, added, after — the lines the reference logs (read_dump.cpp:143-151). */
int ucgb200_host_read_dump(ucgb200_ctx *ctx, int narg, const char *const *arg, long long stats[7], char *errbuf, int errlen);

/* A data file of atom_style ucg (header, Masses, Atoms, Velocities) -> arrays, data_atom_post applied */
int ucgb200_host_data_read(const char *file, ucgb200_data **out, char *errbuf, int errlen);
void ucgb200_host_data_free(ucgb200_data *d);
int ucgb200_host_data_info(const ucgb200_data *d, long long *natoms, int *ntypes, double lo[3], double hi[3]);
int ucgb200_host_data_view(ucgb200_data *d, ucgb200_atoms *view, const double **q, const int **image, const double **mass);
int ucgb200_host_data_upload(ucgb200_ctx *ctx, ucgb200_data *d, const int periodic[3]);

#ifdef __cplusplus
}
#endif
#endif
