/* ucg_oracle.c — CPU oracle for the UCG hot path.  TEST INFRASTRUCTURE ONLY
 * (see ucg_oracle.h for the rules about who may load it and how it is pinned).
 *
 * Each function cites the reference file:line it restates (paths relative to the
 * reference repo KJAdams2000/LAMMPS-UCG-dev).  "[stock]" marks algorithms of
 * upstream LAMMPS that the reference relies on but does not ship.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: results must be the
 * plain IEEE sequence the reference's expressions denote).
 */
#include "ucg_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NEIGHMASK 0x1FFFFFFF /* [stock] lmptype.h */
#define SBBITS 30
#define MAXFIX 16

typedef struct {
  int ninput, rflag, fpflag, match, ntablebits, nshiftbits, nmask;
  double rlo, rhi, fplo, fphi, cut;
  double *rfile, *efile, *ffile, *e2file, *f2file;
  double innersq, delta, invdelta, deltasq6;
  double *rsq, *drsq, *e, *de, *f, *df, *e2, *f2;
  int n; /* entries in e/f */
} Table;

typedef struct {
  double u[98];
  int i97, j97;
  double c, cd, cm;
} RanMars;

typedef struct {
  int kind;
  int groupbit;
  /* ttarget / langevin */
  double t_start, t_stop, t_period, t_target, tsqrt;
  int seed;
  RanMars rng;
  double *gfactor1, *gfactor2;
  double lambda_temp;
  /* wall */
  int bias_flag;
  double barrier;
  /* ucgstate */
  int mode;
  double rate;
  double kT;
} FixDesc;

struct orc_sys {
  char err[512];
  double boltz, ftm2v, mvv2e, dt;
  double boxlo[3], boxhi[3], prd[3];
  double special_lj[4];
  int newton_pair;
  /* pair style */
  int tabstyle, tablength;
  int n_actual, n_formal, max_states;
  int *n_states;     /* [n_actual+1] */
  int *formal_from;  /* [(n_actual+1)*2] */
  int *actual_from;  /* [n_formal+1] */
  double *chem_pot;  /* [n_formal+1] */
  double *mass;      /* [n_formal+1] */
  Table *tables;
  int ntables;
  int *tabindex, *setflag; /* (n_formal+1)^2 */
  double *cutsq;
  double kT;
  int kT_set;
  /* bethe */
  int pair_kind; /* 0 ucgld, 1 bethe (what run() drives) */
  int b_method, b_pseudo, b_prior, b_seed;
  double b_noise;
  RanMars b_rng;
  double *prior_prob, *post_prob;
  int b_nmax;
  /* atoms */
  int nlocal, nghost, nmax;
  double *x, *v, *f;
  int *type, *mask, *tag, *molecule, *ucgstate, *num_ucgstates;
  double *ucgl, *ucgvl, *ucgml, *ucgp, *ucgforce, *scores;
  int *ghost_owner;
  double *ghost_shift; /* 3 per ghost, in units of prd */
  /* neighbor */
  double skin, cutneighmax;
  int full;
  double *xhold;
  int *numneigh;
  long long *firstneigh;
  int *neigh;
  long long neigh_cap, neigh_total;
  int ago, nbuilds;
  /* accumulators */
  double eng_vdwl, virial[6];
  int evflag, eflag, vflag;
  /* fixes */
  FixDesc fix[MAXFIX];
  int nfix;
  long long ntimestep, beginstep, endstep;
  double timers[4];
};

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------ RNGs */
/* [stock] RanMars (random_mars.cpp): Marsaglia's 97-lag subtract-with-borrow
 * generator; call sites fix_ucgld_langevin.cpp:87,280, fix_ucgstate.cpp:61,117 */
static double ranmars_uniform(RanMars *r) {
  double uni = r->u[r->i97] - r->u[r->j97];
  if (uni < 0.0) uni += 1.0;
  r->u[r->i97] = uni;
  r->i97--;
  if (r->i97 == 0) r->i97 = 97;
  r->j97--;
  if (r->j97 == 0) r->j97 = 97;
  r->c -= r->cd;
  if (r->c < 0.0) r->c += r->cm;
  uni -= r->c;
  if (uni < 0.0) uni += 1.0;
  return uni;
}
static void ranmars_init(RanMars *r, int seed) {
  int ij = (seed - 1) / 30082;
  int kl = (seed - 1) - 30082 * ij;
  int i = (ij / 177) % 177 + 2;
  int j = ij % 177 + 2;
  int k = (kl / 169) % 178 + 1;
  int l = kl % 169;
  memset(r->u, 0, sizeof(r->u));
  for (int ii = 1; ii <= 97; ii++) {
    double s = 0.0, t = 0.5;
    for (int jj = 1; jj <= 24; jj++) {
      int m = ((i * j) % 179) * k % 179;
      i = j;
      j = k;
      k = m;
      l = (53 * l + 1) % 169;
      if ((l * m) % 64 >= 32) s = s + t;
      t = 0.5 * t;
    }
    r->u[ii] = s;
  }
  r->c = 362436.0 / 16777216.0;
  r->cd = 7654321.0 / 16777216.0;
  r->cm = 16777213.0 / 16777216.0;
  r->i97 = 97;
  r->j97 = 33;
  ranmars_uniform(r);
}
void orc_ranmars_fill(int seed, int n, double *out) {
  RanMars r;
  ranmars_init(&r, seed);
  for (int i = 0; i < n; i++) out[i] = ranmars_uniform(&r);
}
/* [stock] RanPark (random_park.cpp): Park-Miller minimal standard;
 * call site fix_cluster_switch.cpp:915 */
void orc_ranpark_fill(int seed, int n, double *out) {
  const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
  const double AM = 1.0 / IM;
  for (int i = 0; i < n; i++) {
    int k = seed / IQ;
    seed = IA * (seed - k * IQ) - IR * k;
    if (seed < 0) seed += IM;
    out[i] = AM * seed;
  }
}

/* ---------------------------------------------------------------- lifecycle */
orc_sys *orc_create(void) {
  orc_sys *s = (orc_sys *)calloc(1, sizeof(orc_sys));
  s->boltz = s->ftm2v = s->mvv2e = 1.0;
  s->dt = 0.005;
  for (int k = 0; k < 4; k++) s->special_lj[k] = 1.0;
  s->newton_pair = 1;
  s->tabstyle = ORC_LINEAR;
  s->max_states = 2;
  return s;
}
static void free_table(Table *tb) {
  free(tb->rfile); free(tb->efile); free(tb->ffile); free(tb->e2file); free(tb->f2file);
  free(tb->rsq); free(tb->drsq); free(tb->e); free(tb->de); free(tb->f); free(tb->df);
  free(tb->e2); free(tb->f2);
}
void orc_destroy(orc_sys *s) {
  if (!s) return;
  for (int m = 0; m < s->ntables; m++) free_table(&s->tables[m]);
  free(s->tables);
  free(s->n_states); free(s->formal_from); free(s->actual_from); free(s->chem_pot); free(s->mass);
  free(s->tabindex); free(s->setflag); free(s->cutsq);
  free(s->x); free(s->v); free(s->f); free(s->type); free(s->mask); free(s->tag); free(s->molecule);
  free(s->ucgstate); free(s->num_ucgstates); free(s->ucgl); free(s->ucgvl); free(s->ucgml);
  free(s->ucgp); free(s->ucgforce); free(s->scores); free(s->ghost_owner); free(s->ghost_shift);
  free(s->xhold); free(s->numneigh); free(s->firstneigh); free(s->neigh);
  free(s->prior_prob); free(s->post_prob);
  for (int i = 0; i < s->nfix; i++) { free(s->fix[i].gfactor1); free(s->fix[i].gfactor2); }
  free(s);
}
const char *orc_error(orc_sys *s) { return s->err; }
static void set_err(orc_sys *s, const char *msg) {
  if (!s->err[0]) snprintf(s->err, sizeof(s->err), "%s", msg);
}

void orc_set_units(orc_sys *s, double boltz, double ftm2v, double mvv2e) {
  s->boltz = boltz; s->ftm2v = ftm2v; s->mvv2e = mvv2e;
}
void orc_set_box(orc_sys *s, const double lo[3], const double hi[3]) {
  for (int d = 0; d < 3; d++) { s->boxlo[d] = lo[d]; s->boxhi[d] = hi[d]; s->prd[d] = hi[d] - lo[d]; }
}
void orc_set_dt(orc_sys *s, double dt) { s->dt = dt; }
void orc_set_special_lj(orc_sys *s, const double sl[4]) { for (int k = 0; k < 4; k++) s->special_lj[k] = sl[k]; }
void orc_set_newton(orc_sys *s, int n) { s->newton_pair = n; }
void orc_set_kT(orc_sys *s, double kT) { s->kT = kT; s->kT_set = 1; }

/* ------------------------------------------------------------- pair set-up */
/* settings(): pair_table_ucgld.cpp:654-716 */
void orc_pair_style(orc_sys *s, int tabstyle, int tablength) {
  s->tabstyle = tabstyle;
  s->tablength = tablength;
  if (tablength < 2) set_err(s, "Illegal number of pair table entries");
}
/* read_state_settings(): pair_table_ucgld.cpp:565-652 (arrays instead of a file) */
void orc_set_types(orc_sys *s, int n_actual, int n_formal, const int *n_states,
                   const int *formal_from_actual, const double *chem_pot, const double *mass) {
  s->n_actual = n_actual; s->n_formal = n_formal;
  s->n_states = (int *)calloc(n_actual + 1, sizeof(int));
  s->formal_from = (int *)calloc((n_actual + 1) * 2, sizeof(int));
  s->actual_from = (int *)calloc(n_formal + 1, sizeof(int));
  s->chem_pot = (double *)calloc(n_formal + 1, sizeof(double));
  s->mass = (double *)calloc(n_formal + 1, sizeof(double));
  for (int i = 1; i <= n_actual; i++) {
    s->n_states[i] = n_states[i];
    if (n_states[i] < 1 || n_states[i] > 2) set_err(s, "Invalid number of states for atom type");
    for (int k = 0; k < 2; k++) {
      int ft = formal_from_actual[i * 2 + k];
      /* 1-state types keep formal_types_from_actual[i][*] = 0 in the reference
         (:644-646); the compute code then indexes tabindex[itype][...] directly. */
      s->formal_from[i * 2 + k] = (n_states[i] == 2) ? ft : 0;
      if (n_states[i] == 2 && ft > 0) s->actual_from[ft] = i;
    }
  }
  for (int i = 1; i <= n_formal; i++) { s->chem_pot[i] = chem_pot ? chem_pot[i] : 0.0; s->mass[i] = mass ? mass[i] : 1.0; }
  /* allocate(): pair_table_ucgld.cpp:93-109 */
  int nt = n_formal + 1;
  s->tabindex = (int *)calloc(nt * nt, sizeof(int));
  s->setflag = (int *)calloc(nt * nt, sizeof(int));
  s->cutsq = (double *)calloc(nt * nt, sizeof(double));
}

/* spline(): pair_table_ucgld.cpp:1375-1404 */
static void spline(const double *x, const double *y, int n, double yp1, double ypn, double *y2) {
  double p, qn, sig, un;
  double *u = (double *)malloc(n * sizeof(double));
  if (yp1 > 0.99e30)
    y2[0] = u[0] = 0.0;
  else {
    y2[0] = -0.5;
    u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1);
  }
  for (int i = 1; i < n - 1; i++) {
    sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    p = sig * y2[i - 1] + 2.0;
    y2[i] = (sig - 1.0) / p;
    u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
    u[i] = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
  }
  if (ypn > 0.99e30)
    qn = un = 0.0;
  else {
    qn = 0.5;
    un = (3.0 / (x[n - 1] - x[n - 2])) * (ypn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
  }
  y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
  for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
  free(u);
}
/* splint(): pair_table_ucgld.cpp:1408-1428 */
static double splint(const double *xa, const double *ya, const double *y2a, int n, double x) {
  int klo = 0, khi = n - 1, k;
  while (khi - klo > 1) {
    k = (khi + klo) >> 1;
    if (xa[k] > x) khi = k; else klo = k;
  }
  double h = xa[khi] - xa[klo];
  double a = (xa[khi] - x) / h;
  double b = (x - xa[klo]) / h;
  return a * ya[klo] + b * ya[khi] + ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (h * h) / 6.0;
}
/* spline_table(): pair_table_ucgld.cpp:1047-1065 */
static void spline_table(Table *tb) {
  int n = tb->ninput;
  tb->e2file = (double *)malloc(n * sizeof(double));
  tb->f2file = (double *)malloc(n * sizeof(double));
  double ep0 = -tb->ffile[0];
  double epn = -tb->ffile[n - 1];
  spline(tb->rfile, tb->efile, n, ep0, epn, tb->e2file);
  if (tb->fpflag == 0) {
    tb->fplo = (tb->ffile[1] - tb->ffile[0]) / (tb->rfile[1] - tb->rfile[0]);
    tb->fphi = (tb->ffile[n - 1] - tb->ffile[n - 2]) / (tb->rfile[n - 1] - tb->rfile[n - 2]);
  }
  spline(tb->rfile, tb->ffile, n, tb->fplo, tb->fphi, tb->f2file);
}

typedef union { int i; float f; } int_float_t;

/* [stock] Pair::init_bitmap (pair.cpp): bit masks for float-indexed tables */
static void init_bitmap(orc_sys *s, double inner, double outer, int ntablebits, int *masklo,
                        int *maskhi, int *nmask, int *nshiftbits) {
  if (ntablebits > (int)sizeof(float) * 8) set_err(s, "Too many total bits for bitmapped lookup table");
  if (inner >= outer) set_err(s, "Table inner cutoff >= outer cutoff");
  int nlowermin = 1;
  while (!((pow(2.0, (double)nlowermin) <= inner * inner) && (pow(2.0, (double)nlowermin + 1.0) > inner * inner))) {
    if (pow(2.0, (double)nlowermin) <= inner * inner) nlowermin++; else nlowermin--;
  }
  int nexpbits = 0;
  double required_range = outer * outer / pow(2.0, (double)nlowermin);
  double available_range = 2.0;
  while (available_range < required_range) {
    nexpbits++;
    available_range = pow(2.0, pow(2.0, (double)nexpbits));
  }
  int nmantbits = ntablebits - nexpbits;
  if (nexpbits > (int)sizeof(float) * 8 - 24) set_err(s, "Too many exponent bits for lookup table");
  if (nmantbits + 1 > 24) set_err(s, "Too many mantissa bits for lookup table");
  if (nmantbits < 3) set_err(s, "Too few bits for lookup table");
  *nshiftbits = 23 - nmantbits;
  *nmask = 1;
  for (int j = 0; j < ntablebits + *nshiftbits; j++) *nmask *= 2;
  *nmask -= 1;
  int_float_t rsq_lookup;
  rsq_lookup.f = (float)(outer * outer);
  *maskhi = rsq_lookup.i & ~(*nmask);
  rsq_lookup.f = (float)(inner * inner);
  *masklo = rsq_lookup.i & ~(*nmask);
}

/* compute_table(): pair_table_ucgld.cpp:1105-1344 */
static void compute_table(orc_sys *s, Table *tb) {
  int tablength = s->tablength, tabstyle = s->tabstyle;
  int tlm1 = tablength - 1;
  double inner = tb->rflag ? tb->rlo : tb->rfile[0];
  tb->innersq = inner * inner;
  tb->delta = (tb->cut * tb->cut - tb->innersq) / tlm1;
  tb->invdelta = 1.0 / tb->delta;

  if (tabstyle == ORC_LOOKUP) {
    tb->n = tlm1;
    tb->e = (double *)malloc(tlm1 * sizeof(double));
    tb->f = (double *)malloc(tlm1 * sizeof(double));
    for (int i = 0; i < tlm1; i++) {
      double rsq = tb->innersq + (i + 0.5) * tb->delta;
      double r = sqrt(rsq);
      tb->e[i] = splint(tb->rfile, tb->efile, tb->e2file, tb->ninput, r);
      tb->f[i] = splint(tb->rfile, tb->ffile, tb->f2file, tb->ninput, r) / r;
    }
  }
  if (tabstyle == ORC_LINEAR) {
    tb->n = tablength;
    tb->rsq = (double *)malloc(tablength * sizeof(double));
    tb->e = (double *)malloc(tablength * sizeof(double));
    tb->f = (double *)malloc(tablength * sizeof(double));
    tb->de = (double *)malloc(tablength * sizeof(double));
    tb->df = (double *)malloc(tablength * sizeof(double));
    for (int i = 0; i < tablength; i++) {
      double rsq = tb->innersq + i * tb->delta;
      double r = sqrt(rsq);
      tb->rsq[i] = rsq;
      if (tb->match) {
        tb->e[i] = tb->efile[i];
        tb->f[i] = tb->ffile[i] / r;
      } else {
        tb->e[i] = splint(tb->rfile, tb->efile, tb->e2file, tb->ninput, r);
        tb->f[i] = splint(tb->rfile, tb->ffile, tb->f2file, tb->ninput, r) / r;
      }
    }
    for (int i = 0; i < tlm1; i++) {
      tb->de[i] = tb->e[i + 1] - tb->e[i];
      tb->df[i] = tb->f[i + 1] - tb->f[i];
    }
    tb->de[tlm1] = tb->df[tlm1] = 0.0; /* reference arrays are tlm1 long */
  }
  if (tabstyle == ORC_SPLINE) {
    tb->n = tablength;
    tb->rsq = (double *)malloc(tablength * sizeof(double));
    tb->e = (double *)malloc(tablength * sizeof(double));
    tb->f = (double *)malloc(tablength * sizeof(double));
    tb->e2 = (double *)malloc(tablength * sizeof(double));
    tb->f2 = (double *)malloc(tablength * sizeof(double));
    tb->deltasq6 = tb->delta * tb->delta / 6.0;
    for (int i = 0; i < tablength; i++) {
      double rsq = tb->innersq + i * tb->delta;
      double r = sqrt(rsq);
      tb->rsq[i] = rsq;
      if (tb->match) {
        tb->e[i] = tb->efile[i];
        tb->f[i] = tb->ffile[i] / r;
      } else {
        tb->e[i] = splint(tb->rfile, tb->efile, tb->e2file, tb->ninput, r);
        tb->f[i] = splint(tb->rfile, tb->ffile, tb->f2file, tb->ninput, r);
      }
    }
    double ep0 = -tb->f[0] / (2.0 * sqrt(tb->innersq));
    double epn = -tb->f[tlm1] / (2.0 * tb->cut);
    spline(tb->rsq, tb->e, tablength, ep0, epn, tb->e2);
    double fp0, fpn;
    double secant_factor = 0.1;
    if (tb->fpflag)
      fp0 = (tb->fplo / sqrt(tb->innersq) - tb->f[0] / tb->innersq) / (2.0 * sqrt(tb->innersq));
    else {
      double rsq1 = tb->innersq;
      double rsq2 = rsq1 + secant_factor * tb->delta;
      fp0 = (splint(tb->rfile, tb->ffile, tb->f2file, tb->ninput, sqrt(rsq2)) / sqrt(rsq2) -
             tb->f[0] / sqrt(rsq1)) / (secant_factor * tb->delta);
    }
    if (tb->fpflag && tb->cut == tb->rfile[tb->ninput - 1])
      fpn = (tb->fphi / tb->cut - tb->f[tlm1] / (tb->cut * tb->cut)) / (2.0 * tb->cut);
    else {
      double rsq2 = tb->cut * tb->cut;
      double rsq1 = rsq2 - secant_factor * tb->delta;
      fpn = (tb->f[tlm1] / sqrt(rsq2) -
             splint(tb->rfile, tb->ffile, tb->f2file, tb->ninput, sqrt(rsq1)) / sqrt(rsq1)) /
          (secant_factor * tb->delta);
    }
    for (int i = 0; i < tablength; i++) tb->f[i] /= sqrt(tb->rsq[i]);
    spline(tb->rsq, tb->f, tablength, fp0, fpn, tb->f2);
  }
  if (tabstyle == ORC_BITMAP) {
    int masklo, maskhi;
    int_float_t rsq_lookup;
    init_bitmap(s, inner, tb->cut, tablength, &masklo, &maskhi, &tb->nmask, &tb->nshiftbits);
    int ntable = 1 << tablength;
    int ntablem1 = ntable - 1;
    tb->n = ntable;
    tb->rsq = (double *)malloc(ntable * sizeof(double));
    tb->e = (double *)malloc(ntable * sizeof(double));
    tb->f = (double *)malloc(ntable * sizeof(double));
    tb->de = (double *)malloc(ntable * sizeof(double));
    tb->df = (double *)malloc(ntable * sizeof(double));
    tb->drsq = (double *)malloc(ntable * sizeof(double));
    int_float_t minrsq_lookup;
    minrsq_lookup.i = 0 << tb->nshiftbits;
    minrsq_lookup.i |= maskhi;
    for (int i = 0; i < ntable; i++) {
      rsq_lookup.i = i << tb->nshiftbits;
      rsq_lookup.i |= masklo;
      if (rsq_lookup.f < tb->innersq) {
        rsq_lookup.i = i << tb->nshiftbits;
        rsq_lookup.i |= maskhi;
      }
      double r = sqrtf(rsq_lookup.f);
      tb->rsq[i] = rsq_lookup.f;
      if (tb->match) {
        tb->e[i] = tb->efile[i];
        tb->f[i] = tb->ffile[i] / r;
      } else {
        tb->e[i] = splint(tb->rfile, tb->efile, tb->e2file, tb->ninput, r);
        tb->f[i] = splint(tb->rfile, tb->ffile, tb->f2file, tb->ninput, r) / r;
      }
      minrsq_lookup.f = (minrsq_lookup.f < rsq_lookup.f) ? minrsq_lookup.f : rsq_lookup.f;
    }
    tb->innersq = minrsq_lookup.f;
    for (int i = 0; i < ntablem1; i++) {
      tb->de[i] = tb->e[i + 1] - tb->e[i];
      tb->df[i] = tb->f[i + 1] - tb->f[i];
      tb->drsq[i] = 1.0 / (tb->rsq[i + 1] - tb->rsq[i]);
    }
    tb->de[ntablem1] = tb->e[0] - tb->e[ntablem1];
    tb->df[ntablem1] = tb->f[0] - tb->f[ntablem1];
    tb->drsq[ntablem1] = 1.0 / (tb->rsq[0] - tb->rsq[ntablem1]);
    int itablemin = minrsq_lookup.i & tb->nmask;
    itablemin >>= tb->nshiftbits;
    int itablemax = itablemin - 1;
    if (itablemin == 0) itablemax = ntablem1;
    int itablemaxm1 = itablemax - 1;
    if (itablemax == 0) itablemaxm1 = ntablem1;
    rsq_lookup.i = itablemax << tb->nshiftbits;
    rsq_lookup.i |= maskhi;
    if (rsq_lookup.f < tb->cut * tb->cut) {
      if (tb->match) {
        tb->de[itablemax] = tb->de[itablemaxm1];
        tb->df[itablemax] = tb->df[itablemaxm1];
        tb->drsq[itablemax] = tb->drsq[itablemaxm1];
      } else {
        rsq_lookup.f = tb->cut * tb->cut;
        double r = sqrtf(rsq_lookup.f);
        double e_tmp = splint(tb->rfile, tb->efile, tb->e2file, tb->ninput, r);
        double f_tmp = splint(tb->rfile, tb->ffile, tb->f2file, tb->ninput, r) / r;
        tb->de[itablemax] = e_tmp - tb->e[itablemax];
        tb->df[itablemax] = f_tmp - tb->f[itablemax];
        tb->drsq[itablemax] = 1.0 / (rsq_lookup.f - tb->rsq[itablemax]);
      }
    }
  }
}

/* the per-table part of coeff(): pair_table_ucgld.cpp:789-829 */
static int finish_table(orc_sys *s, Table *tb, double cut) {
  tb->cut = cut;
  if (tb->ninput <= 1) { set_err(s, "Invalid pair table length"); return -1; }
  double rlo, rhi;
  if (tb->rflag == 0) { rlo = tb->rfile[0]; rhi = tb->rfile[tb->ninput - 1]; }
  else { rlo = tb->rlo; rhi = tb->rhi; }
  if (tb->cut <= rlo || tb->cut > rhi) { set_err(s, "Pair table cutoff outside of table"); return -1; }
  if (rlo <= 0.0) { set_err(s, "Invalid pair table lower boundary"); return -1; }
  tb->match = 0;
  if (s->tabstyle == ORC_LINEAR && tb->ninput == s->tablength && tb->rflag == ORC_RSQ && tb->rhi == tb->cut) tb->match = 1;
  if (s->tabstyle == ORC_BITMAP && tb->ninput == 1 << s->tablength && tb->rflag == ORC_BMP && tb->rhi == tb->cut) tb->match = 1;
  if (tb->rflag == ORC_BMP && tb->match == 0) { set_err(s, "Bitmapped table in file does not match requested table"); return -1; }
  if (tb->match == 0) spline_table(tb);
  compute_table(s, tb);
  return 0;
}

static Table *new_table(orc_sys *s) {
  s->tables = (Table *)realloc(s->tables, (s->ntables + 1) * sizeof(Table));
  Table *tb = &s->tables[s->ntables];
  memset(tb, 0, sizeof(Table));
  return tb;
}

/* r values as read_table() recomputes them: pair_table_ucgld.cpp:954-972 */
static void fill_rfile(orc_sys *s, Table *tb, const double *rfile_in) {
  int masklo = 0, maskhi = 0, nmask = 0, nshiftbits = 0;
  int_float_t rsq_lookup;
  if (tb->rflag == ORC_BMP) {
    tb->ntablebits = 0;
    while (1 << tb->ntablebits < tb->ninput) tb->ntablebits++;
    if (1 << tb->ntablebits != tb->ninput) set_err(s, "Bitmapped table is incorrect length in table file");
    init_bitmap(s, tb->rlo, tb->rhi, tb->ntablebits, &masklo, &maskhi, &nmask, &nshiftbits);
  }
  for (int i = 0; i < tb->ninput; i++) {
    double rnew = rfile_in ? rfile_in[i] : 0.0;
    if (tb->rflag == ORC_RLINEAR)
      rnew = tb->rlo + (tb->rhi - tb->rlo) * i / (tb->ninput - 1);
    else if (tb->rflag == ORC_RSQ) {
      rnew = tb->rlo * tb->rlo + (tb->rhi * tb->rhi - tb->rlo * tb->rlo) * i / (tb->ninput - 1);
      rnew = sqrt(rnew);
    } else if (tb->rflag == ORC_BMP) {
      rsq_lookup.i = i << nshiftbits;
      rsq_lookup.i |= masklo;
      if (rsq_lookup.f < tb->rlo * tb->rlo) {
        rsq_lookup.i = i << nshiftbits;
        rsq_lookup.i |= maskhi;
      }
      rnew = sqrtf(rsq_lookup.f);
    }
    tb->rfile[i] = rnew;
  }
}

int orc_table_add_arrays(orc_sys *s, int ninput, int rflag, double rlo, double rhi, int fpflag,
                         double fplo, double fphi, const double *rfile, const double *efile,
                         const double *ffile, double cut) {
  Table *tb = new_table(s);
  tb->ninput = ninput; tb->rflag = rflag; tb->rlo = rlo; tb->rhi = rhi;
  tb->fpflag = fpflag; tb->fplo = fplo; tb->fphi = fphi;
  tb->rfile = (double *)malloc(ninput * sizeof(double));
  tb->efile = (double *)malloc(ninput * sizeof(double));
  tb->ffile = (double *)malloc(ninput * sizeof(double));
  memcpy(tb->efile, efile, ninput * sizeof(double));
  memcpy(tb->ffile, ffile, ninput * sizeof(double));
  fill_rfile(s, tb, rfile);
  if (finish_table(s, tb, cut)) return -1;
  return s->ntables++;
}

/* read_table() + param_extract(): pair_table_ucgld.cpp:897-1017, 1067-1102.
 * [stock] TableFileReader: skip to the line whose first word is `keyword`, the next
 * non-blank line holds the parameters, then a blank line, then N data lines. */
int orc_table_add_file(orc_sys *s, const char *file, const char *keyword, double cut) {
  FILE *fp = fopen(file, "r");
  if (!fp) { set_err(s, "Cannot open table file"); return -1; }
  char line[1024];
  int found = 0;
  while (fgets(line, sizeof(line), fp)) {
    char word[256];
    if (line[0] == '#') continue;
    if (sscanf(line, "%255s", word) == 1 && strcmp(word, keyword) == 0) { found = 1; break; }
  }
  if (!found) { fclose(fp); set_err(s, "Did not find keyword in table file"); return -1; }
  if (!fgets(line, sizeof(line), fp)) { fclose(fp); set_err(s, "Unexpected end of table file"); return -1; }
  Table *tb = new_table(s);
  char *tok = strtok(line, " \t\n\r");
  while (tok) {
    if (strcmp(tok, "N") == 0) tb->ninput = atoi(strtok(NULL, " \t\n\r"));
    else if (!strcmp(tok, "R") || !strcmp(tok, "RSQ") || !strcmp(tok, "BITMAP")) {
      tb->rflag = !strcmp(tok, "R") ? ORC_RLINEAR : (!strcmp(tok, "RSQ") ? ORC_RSQ : ORC_BMP);
      tb->rlo = atof(strtok(NULL, " \t\n\r"));
      tb->rhi = atof(strtok(NULL, " \t\n\r"));
    } else if (strcmp(tok, "FPRIME") == 0) {
      tb->fpflag = 1;
      tb->fplo = atof(strtok(NULL, " \t\n\r"));
      tb->fphi = atof(strtok(NULL, " \t\n\r"));
    } else { fclose(fp); set_err(s, "Invalid keyword in pair table parameters"); return -1; }
    tok = strtok(NULL, " \t\n\r");
  }
  if (tb->ninput == 0) { fclose(fp); set_err(s, "Pair table parameters did not set N"); return -1; }
  int n = tb->ninput;
  tb->rfile = (double *)malloc(n * sizeof(double));
  tb->efile = (double *)malloc(n * sizeof(double));
  tb->ffile = (double *)malloc(n * sizeof(double));
  double *rin = (double *)malloc(n * sizeof(double));
  int got = 0;
  while (got < n && fgets(line, sizeof(line), fp)) {
    int idx;
    double r, e, f;
    if (sscanf(line, "%d %lg %lg %lg", &idx, &r, &e, &f) == 4) {
      rin[got] = r; tb->efile[got] = e; tb->ffile[got] = f; got++;
    }
  }
  fclose(fp);
  if (got < n) { free(rin); set_err(s, "Data missing when parsing pair table"); return -1; }
  fill_rfile(s, tb, rin);
  free(rin);
  if (finish_table(s, tb, cut)) return -1;
  return s->ntables++;
}

int orc_table_len(orc_sys *s, int idx) { return s->tables[idx].n; }
int orc_table_get(orc_sys *s, int idx, int which, double *out) {
  Table *tb = &s->tables[idx];
  double *src[8] = {tb->rsq, tb->e, tb->f, tb->de, tb->df, tb->e2, tb->f2, tb->drsq};
  if (!src[which]) return 0;
  int n = tb->n;
  if ((which == 3 || which == 4) && s->tabstyle == ORC_LINEAR) n = tb->n - 1;
  memcpy(out, src[which], n * sizeof(double));
  return n;
}
void orc_table_params(orc_sys *s, int idx, double out[8]) {
  Table *tb = &s->tables[idx];
  out[0] = tb->innersq; out[1] = tb->delta; out[2] = tb->invdelta; out[3] = tb->deltasq6;
  out[4] = tb->cut; out[5] = tb->nmask; out[6] = tb->nshiftbits; out[7] = tb->match;
}

/* table-to-type assignment of coeff(): pair_table_ucgld.cpp:833-851 */
int orc_pair_coeff(orc_sys *s, int ilo, int ihi, int jlo, int jhi, int ns_i, int ns_j, const int *tables) {
  int nt = s->n_formal + 1;
  int t = 0;
  for (int s_i = 0; s_i < ns_i; s_i++)
    for (int s_j = 0; s_j < ns_j; s_j++) {
      int count = 0;
      for (int i = ilo; i <= ihi; i++)
        for (int j = (jlo > i ? jlo : i); j <= jhi; j++) {
          /* 1-state types: formal_types_from_actual is 0 in the reference and the
             assignment would hit "Formal type not defined"; real decks therefore give
             1-state types formal == actual through the state file.  We accept either. */
          int fi = s->formal_from[i * 2 + s_i];
          int fj = s->formal_from[j * 2 + s_j];
          if (s->n_states[i] == 1) fi = i;
          if (s->n_states[j] == 1) fj = j;
          if (fi == 0 || fj == 0) { set_err(s, "Formal type not defined in pair_style command"); return -1; }
          s->tabindex[fi * nt + fj] = tables[t];
          s->setflag[fi * nt + fj] = 1;
          count++;
        }
      if (count == 0) { set_err(s, "Illegal pair_coeff command"); return -1; }
      t++;
    }
  return 0;
}
/* Pair::init [stock] loops i<=j over atom->ntypes (== n_formal) calling init_one():
 * pair_table_ucgld.cpp:886-895 */
int orc_pair_init(orc_sys *s) {
  int nt = s->n_formal + 1;
  for (int i = 1; i <= s->n_formal; i++)
    for (int j = i; j <= s->n_formal; j++) {
      if (s->setflag[i * nt + j] == 0) { set_err(s, "All pair coeffs are not set"); return -1; }
      s->tabindex[j * nt + i] = s->tabindex[i * nt + j];
      double cut = s->tables[s->tabindex[i * nt + j]].cut;
      s->cutsq[i * nt + j] = s->cutsq[j * nt + i] = cut * cut;
    }
  return 0;
}
void orc_get_pair_maps(orc_sys *s, int *tabindex, double *cutsq) {
  int nt = s->n_formal + 1;
  memcpy(tabindex, s->tabindex, nt * nt * sizeof(int));
  memcpy(cutsq, s->cutsq, nt * nt * sizeof(double));
}

/* -------------------------------------------------------------------- atoms */
static void grow_atoms(orc_sys *s, int nmax) {
  if (nmax <= s->nmax) return;
  s->nmax = nmax + nmax / 4 + 1024;
  int n = s->nmax;
#define GROW(p, T, k) s->p = (T *)realloc(s->p, (size_t)n * (k) * sizeof(T))
  GROW(x, double, 3); GROW(v, double, 3); GROW(f, double, 3);
  GROW(type, int, 1); GROW(mask, int, 1); GROW(tag, int, 1); GROW(molecule, int, 1);
  GROW(ucgstate, int, 1); GROW(num_ucgstates, int, 1);
  GROW(ucgl, double, 1); GROW(ucgvl, double, 1); GROW(ucgml, double, 1); GROW(ucgp, double, 1);
  GROW(ucgforce, double, 1); GROW(scores, double, 2);
  GROW(ghost_owner, int, 1); GROW(ghost_shift, double, 3);
#undef GROW
}
void orc_set_atoms(orc_sys *s, int n, const double *x, const double *v, const int *type,
                   const int *mask, const int *tag, const int *molecule, const int *ucgstate,
                   const double *ucgl, const double *ucgvl, const double *ucgml, const double *ucgp) {
  grow_atoms(s, n);
  s->nlocal = n; s->nghost = 0;
  memcpy(s->x, x, 3 * (size_t)n * sizeof(double));
  if (v) memcpy(s->v, v, 3 * (size_t)n * sizeof(double)); else memset(s->v, 0, 3 * (size_t)n * sizeof(double));
  memset(s->f, 0, 3 * (size_t)n * sizeof(double));
  for (int i = 0; i < n; i++) {
    s->type[i] = type[i];
    s->mask[i] = mask ? mask[i] : 1;
    s->tag[i] = tag ? tag[i] : i + 1;
    s->molecule[i] = molecule ? molecule[i] : 0;
    /* data_atom_post(): atom_vec_ucg.cpp:145-170 clamps */
    int st = ucgstate ? ucgstate[i] : 0;
    s->ucgstate[i] = st < 0 ? 0 : (st > 1 ? 1 : st);
    double l = ucgl ? ucgl[i] : 0.0;
    s->ucgl[i] = l < 0 ? 0. : (l > 1 ? 1. : l);
    s->ucgvl[i] = ucgvl ? ucgvl[i] : 0.0;
    s->ucgml[i] = ucgml ? ucgml[i] : 1.0;
    s->ucgp[i] = ucgp ? ucgp[i] : -1.0;
    s->ucgforce[i] = 0.0;
    s->scores[2 * i] = s->scores[2 * i + 1] = 0.0;
    s->num_ucgstates[i] = 0;
  }
}
int orc_nlocal(orc_sys *s) { return s->nlocal; }
int orc_nghost(orc_sys *s) { return s->nghost; }
void orc_get_atoms(orc_sys *s, double *x, double *v, double *f, int *type, int *tag, int *ucgstate,
                   double *ucgl, double *ucgvl, double *ucgp, double *ucgforce, double *scores,
                   int *num_ucgstates) {
  size_t n = s->nlocal;
  if (x) memcpy(x, s->x, 3 * n * sizeof(double));
  if (v) memcpy(v, s->v, 3 * n * sizeof(double));
  if (f) memcpy(f, s->f, 3 * n * sizeof(double));
  if (type) memcpy(type, s->type, n * sizeof(int));
  if (tag) memcpy(tag, s->tag, n * sizeof(int));
  if (ucgstate) memcpy(ucgstate, s->ucgstate, n * sizeof(int));
  if (ucgl) memcpy(ucgl, s->ucgl, n * sizeof(double));
  if (ucgvl) memcpy(ucgvl, s->ucgvl, n * sizeof(double));
  if (ucgp) memcpy(ucgp, s->ucgp, n * sizeof(double));
  if (ucgforce) memcpy(ucgforce, s->ucgforce, n * sizeof(double));
  if (scores) memcpy(scores, s->scores, 2 * n * sizeof(double));
  if (num_ucgstates) memcpy(num_ucgstates, s->num_ucgstates, n * sizeof(int));
}
void orc_get_ghosts(orc_sys *s, double *x, int *tag) {
  if (x) memcpy(x, s->x + 3 * (size_t)s->nlocal, 3 * (size_t)s->nghost * sizeof(double));
  if (tag) memcpy(tag, s->tag + s->nlocal, (size_t)s->nghost * sizeof(int));
}
void orc_set_forces(orc_sys *s, const double *f, const double *ucgforce, const double *scores) {
  size_t n = s->nlocal;
  if (f) memcpy(s->f, f, 3 * n * sizeof(double));
  if (ucgforce) memcpy(s->ucgforce, ucgforce, n * sizeof(double));
  if (scores) memcpy(s->scores, scores, 2 * n * sizeof(double));
}

/* ------------------------------------------------- domain / comm / neighbor */
void orc_neigh_config(orc_sys *s, double skin, int full) { s->skin = skin; s->full = full; }

static double max_cut(orc_sys *s) {
  double c = 0.0;
  for (int m = 0; m < s->ntables; m++) if (s->tables[m].cut > c) c = s->tables[m].cut;
  /* [stock] Neighbor::init: cutneighmax = max over type pairs of sqrt(cutsq)+skin */
  int nt = s->n_formal + 1;
  double c2 = 0.0;
  for (int i = 1; i <= s->n_formal; i++)
    for (int j = 1; j <= s->n_formal; j++) if (s->cutsq[i * nt + j] > c2) c2 = s->cutsq[i * nt + j];
  if (c2 > 0.0) c = sqrt(c2);
  return c;
}

/* [stock] Domain::pbc for an orthogonal, fully periodic box */
void orc_pbc(orc_sys *s) {
  for (int i = 0; i < s->nlocal; i++)
    for (int d = 0; d < 3; d++) {
      double *xi = &s->x[3 * i + d];
      if (*xi < s->boxlo[d]) *xi += s->prd[d];
      if (*xi >= s->boxhi[d]) {
        *xi -= s->prd[d];
        if (*xi < s->boxlo[d]) *xi = s->boxlo[d];
      }
    }
}

/* [stock] CommBrick::borders on one rank with periodic self-exchange: for each
 * dimension, atoms (owned + ghosts made by earlier dimensions) within cutghost of
 * the low face are copied to +prd, those within cutghost of the high face to -prd.
 * Payload per ghost = fields_border (atom_vec_ucg.cpp:66-67). */
void orc_borders(orc_sys *s) {
  double cutghost = max_cut(s) + s->skin;
  s->cutneighmax = cutghost;
  s->nghost = 0;
  for (int dim = 0; dim < 3; dim++) {
    int nlast = s->nlocal + s->nghost;
    for (int side = 0; side < 2; side++) {
      double lo = side == 0 ? -1e300 : s->boxhi[dim] - cutghost;
      double hi = side == 0 ? s->boxlo[dim] + cutghost : 1e300;
      double shift = side == 0 ? 1.0 : -1.0;
      for (int i = 0; i < nlast; i++) {
        double xi = s->x[3 * i + dim];
        if (xi >= lo && xi <= hi) {
          int g = s->nlocal + s->nghost;
          grow_atoms(s, g + 1);
          for (int d = 0; d < 3; d++) s->x[3 * g + d] = s->x[3 * i + d];
          s->x[3 * g + dim] = s->x[3 * i + dim] + shift * s->prd[dim];
          int owner = i < s->nlocal ? i : s->ghost_owner[i - s->nlocal];
          s->ghost_owner[g - s->nlocal] = owner;
          for (int d = 0; d < 3; d++)
            s->ghost_shift[3 * (g - s->nlocal) + d] = i < s->nlocal ? 0.0 : s->ghost_shift[3 * (i - s->nlocal) + d];
          s->ghost_shift[3 * (g - s->nlocal) + dim] = shift;
          s->type[g] = s->type[i]; s->mask[g] = s->mask[i]; s->tag[g] = s->tag[i];
          s->molecule[g] = s->molecule[i];
          s->ucgstate[g] = s->ucgstate[i]; s->num_ucgstates[g] = s->num_ucgstates[i];
          s->ucgl[g] = s->ucgl[i]; s->ucgp[g] = s->ucgp[i];
          s->nghost++;
        }
      }
    }
  }
}

/* [stock] Comm::forward_comm: x + fields_comm {ucgstate, ucgl, ucgp}
 * (atom_vec_ucg.cpp:71); ghost x = owner x + pbc shift, applied per dimension in the
 * order the ghost was created */
void orc_forward_comm(orc_sys *s) {
  for (int g = 0; g < s->nghost; g++) {
    int o = s->ghost_owner[g], k = s->nlocal + g;
    for (int d = 0; d < 3; d++) {
      double sh = s->ghost_shift[3 * g + d];
      s->x[3 * k + d] = (sh == 0.0) ? s->x[3 * o + d] : s->x[3 * o + d] + sh * s->prd[d];
    }
    s->ucgstate[k] = s->ucgstate[o];
    s->ucgl[k] = s->ucgl[o];
    s->ucgp[k] = s->ucgp[o];
  }
}
/* [stock] Comm::reverse_comm: f + fields_reverse {ucgforce, ucgsoftmaxscores}
 * (atom_vec_ucg.cpp:73) summed into the owner, last ghost first */
void orc_reverse_comm(orc_sys *s) {
  for (int g = s->nghost - 1; g >= 0; g--) {
    int o = s->ghost_owner[g], k = s->nlocal + g;
    for (int d = 0; d < 3; d++) s->f[3 * o + d] += s->f[3 * k + d];
    s->ucgforce[o] += s->ucgforce[k];
    s->scores[2 * o] += s->scores[2 * k];
    s->scores[2 * o + 1] += s->scores[2 * k + 1];
  }
}

/* [stock] AtomVec::force_clear + AtomVecUCG::force_clear (atom_vec_ucg.cpp:131-135);
 * with newton on the ghosts are cleared too */
void orc_force_clear(orc_sys *s) {
  size_t n = s->nlocal + (s->newton_pair ? s->nghost : 0);
  memset(s->f, 0, 3 * n * sizeof(double));
  memset(s->ucgforce, 0, n * sizeof(double));
  memset(s->scores, 0, 2 * n * sizeof(double));
}

/* [stock] NBinStandard + NStencil{Half,Full}Bin3d + NPair{Half,Full}BinNewton:
 * bins of side ~cutneighmax/2, atoms linked in ascending index (owned before ghost);
 * half/newton: same bin -> j after i in the list, ghosts only if "above/right" of i;
 * other bins -> only the upper half stencil.  Criterion rsq <= cutneighsq[itype][jtype]. */
void orc_neigh_build(orc_sys *s) {
  int nlocal = s->nlocal, nall = s->nlocal + s->nghost;
  double cutneigh = s->cutneighmax > 0 ? s->cutneighmax : max_cut(s) + s->skin;
  int nt = s->n_formal + 1;
  double binsize = 0.5 * cutneigh;
  double bboxlo[3], bboxhi[3];
  int nbin[3], mbinlo[3], mbin[3];
  double bininv[3];
  for (int d = 0; d < 3; d++) {
    bboxlo[d] = s->boxlo[d]; bboxhi[d] = s->boxhi[d];
    nbin[d] = (int)((bboxhi[d] - bboxlo[d]) / binsize);
    if (nbin[d] == 0) nbin[d] = 1;
    double bs = (bboxhi[d] - bboxlo[d]) / nbin[d];
    bininv[d] = 1.0 / bs;
    /* extend to cover ghosts */
    double lo = bboxlo[d] - cutneigh - 1e-6 * (bboxhi[d] - bboxlo[d]);
    double hi = bboxhi[d] + cutneigh + 1e-6 * (bboxhi[d] - bboxlo[d]);
    mbinlo[d] = (int)floor((lo - bboxlo[d]) * bininv[d]) - 1;
    int mbinhi = (int)floor((hi - bboxlo[d]) * bininv[d]) + 1;
    mbin[d] = mbinhi - mbinlo[d] + 1;
  }
  long long mbins = (long long)mbin[0] * mbin[1] * mbin[2];
  int *binhead = (int *)malloc(mbins * sizeof(int));
  int *bins = (int *)malloc((size_t)nall * sizeof(int));
  int *atom2bin = (int *)malloc((size_t)nall * sizeof(int));
  for (long long b = 0; b < mbins; b++) binhead[b] = -1;
#define COORD2BIN(xp, out)                                                         \
  do {                                                                             \
    int ib[3];                                                                     \
    for (int d = 0; d < 3; d++) {                                                  \
      ib[d] = (int)floor(((xp)[d] - bboxlo[d]) * bininv[d]) - mbinlo[d];           \
      if (ib[d] < 0) ib[d] = 0;                                                    \
      if (ib[d] >= mbin[d]) ib[d] = mbin[d] - 1;                                   \
    }                                                                              \
    out = (ib[2] * mbin[1] + ib[1]) * mbin[0] + ib[0];                             \
  } while (0)
  for (int i = nall - 1; i >= nlocal; i--) {
    int b; COORD2BIN(&s->x[3 * i], b);
    atom2bin[i] = b; bins[i] = binhead[b]; binhead[b] = i;
  }
  for (int i = nlocal - 1; i >= 0; i--) {
    int b; COORD2BIN(&s->x[3 * i], b);
    atom2bin[i] = b; bins[i] = binhead[b]; binhead[b] = i;
  }
  /* stencil */
  int sx = (int)(cutneigh * bininv[0]); if (sx * (1.0 / bininv[0]) < cutneigh) sx++;
  int sy = (int)(cutneigh * bininv[1]); if (sy * (1.0 / bininv[1]) < cutneigh) sy++;
  int sz = (int)(cutneigh * bininv[2]); if (sz * (1.0 / bininv[2]) < cutneigh) sz++;
  int maxst = (2 * sx + 1) * (2 * sy + 1) * (2 * sz + 1);
  int *stencil = (int *)malloc(maxst * sizeof(int));
  int nstencil = 0;
  double cutsqmax = cutneigh * cutneigh;
  for (int k = (s->full ? -sz : 0); k <= sz; k++)
    for (int j = -sy; j <= sy; j++)
      for (int i = -sx; i <= sx; i++) {
        if (!s->full && !(k > 0 || j > 0 || (j == 0 && i > 0))) continue;
        if (s->full && i == 0 && j == 0 && k == 0) continue;
        double dx = i > 0 ? (i - 1) / bininv[0] : (i == 0 ? 0.0 : (i + 1) / bininv[0]);
        double dy = j > 0 ? (j - 1) / bininv[1] : (j == 0 ? 0.0 : (j + 1) / bininv[1]);
        double dz = k > 0 ? (k - 1) / bininv[2] : (k == 0 ? 0.0 : (k + 1) / bininv[2]);
        if (dx * dx + dy * dy + dz * dz < cutsqmax) stencil[nstencil++] = (k * mbin[1] + j) * mbin[0] + i;
      }
  /* cutneighsq per ACTUAL type pair (the reference reads cutsq[itype][jtype] with
     actual types, pair_table_ucgld.cpp:213) */
  s->numneigh = (int *)realloc(s->numneigh, (size_t)nlocal * sizeof(int));
  s->firstneigh = (long long *)realloc(s->firstneigh, (size_t)(nlocal + 1) * sizeof(long long));
  long long total = 0;
  for (int i = 0; i < nlocal; i++) {
    int itype = s->type[i];
    double xt = s->x[3 * i], yt = s->x[3 * i + 1], zt = s->x[3 * i + 2];
    s->firstneigh[i] = total;
    int n = 0;
    if (total + 4096 > s->neigh_cap) {
      s->neigh_cap = (s->neigh_cap + 4096) * 2;
      s->neigh = (int *)realloc(s->neigh, (size_t)s->neigh_cap * sizeof(int));
    }
    int *row = s->neigh + total;
    /* own bin */
    for (int j = s->full ? binhead[atom2bin[i]] : bins[i]; j >= 0; j = bins[j]) {
      if (s->full) { if (j == i) continue; }
      else if (j >= nlocal) {
        if (s->x[3 * j + 2] < zt) continue;
        if (s->x[3 * j + 2] == zt) {
          if (s->x[3 * j + 1] < yt) continue;
          if (s->x[3 * j + 1] == yt && s->x[3 * j] < xt) continue;
        }
      }
      double delx = xt - s->x[3 * j], dely = yt - s->x[3 * j + 1], delz = zt - s->x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      double c = sqrt(s->cutsq[itype * nt + s->type[j]]) + s->skin;
      if (rsq <= c * c) row[n++] = j;
    }
    int ibin = atom2bin[i];
    for (int k = 0; k < nstencil; k++)
      for (int j = binhead[ibin + stencil[k]]; j >= 0; j = bins[j]) {
        double delx = xt - s->x[3 * j], dely = yt - s->x[3 * j + 1], delz = zt - s->x[3 * j + 2];
        double rsq = delx * delx + dely * dely + delz * delz;
        double c = sqrt(s->cutsq[itype * nt + s->type[j]]) + s->skin;
        if (rsq <= c * c) row[n++] = j;
      }
    s->numneigh[i] = n;
    total += n;
  }
  s->firstneigh[nlocal] = total;
  s->neigh_total = total;
  free(binhead); free(bins); free(atom2bin); free(stencil);
  /* [stock] Neighbor::build: remember positions for the skin check */
  s->xhold = (double *)realloc(s->xhold, 3 * (size_t)nlocal * sizeof(double));
  memcpy(s->xhold, s->x, 3 * (size_t)nlocal * sizeof(double));
  s->ago = 0;
  s->nbuilds++;
}

/* [stock] Neighbor::decide + check_distance with `neigh_modify delay 0 every 1 check yes` */
int orc_neigh_decide(orc_sys *s) {
  s->ago++;
  double deltasq = 0.25 * s->skin * s->skin; /* triggersq */
  for (int i = 0; i < s->nlocal; i++) {
    double delx = s->x[3 * i] - s->xhold[3 * i];
    double dely = s->x[3 * i + 1] - s->xhold[3 * i + 1];
    double delz = s->x[3 * i + 2] - s->xhold[3 * i + 2];
    double rsq = delx * delx + dely * dely + delz * delz;
    if (rsq > deltasq) return 1;
  }
  return 0;
}
long long orc_neigh_total(orc_sys *s) { return s->neigh_total; }
long long orc_neigh_pairs(orc_sys *s, int *tag_i, int *tag_j) {
  long long n = 0;
  for (int i = 0; i < s->nlocal; i++)
    for (int jj = 0; jj < s->numneigh[i]; jj++) {
      int j = s->neigh[s->firstneigh[i] + jj] & NEIGHMASK;
      tag_i[n] = s->tag[i]; tag_j[n] = s->tag[j]; n++;
    }
  return n;
}

/* ------------------------------------------------------------ pair: ucgld */
/* one table evaluation; the four copies in the reference are identical:
 * pair_table_ucgld.cpp:223-268, 279-324, 354-399, 437-482 */
static int table_eval(orc_sys *s, const Table *tb, double rsq, double factor_lj, double *evdwl, double *fpair) {
  int tlm1 = s->tablength - 1, itable;
  double fraction = 0, value, a = 0, b = 0;
  if (rsq < tb->innersq) { set_err(s, "Pair distance < table inner cutoff"); return 1; }
  if (s->tabstyle == ORC_LOOKUP) {
    itable = (int)((rsq - tb->innersq) * tb->invdelta);
    if (itable >= tlm1) { set_err(s, "Pair distance > table outer cutoff"); return 2; }
    *fpair = factor_lj * tb->f[itable];
  } else if (s->tabstyle == ORC_LINEAR) {
    itable = (int)((rsq - tb->innersq) * tb->invdelta);
    if (itable >= tlm1) { set_err(s, "Pair distance > table outer cutoff"); return 2; }
    fraction = (rsq - tb->rsq[itable]) * tb->invdelta;
    value = tb->f[itable] + fraction * tb->df[itable];
    *fpair = factor_lj * value;
  } else if (s->tabstyle == ORC_SPLINE) {
    itable = (int)((rsq - tb->innersq) * tb->invdelta);
    if (itable >= tlm1) { set_err(s, "Pair distance > table outer cutoff"); return 2; }
    b = (rsq - tb->rsq[itable]) * tb->invdelta;
    a = 1.0 - b;
    value = a * tb->f[itable] + b * tb->f[itable + 1] +
        ((a * a * a - a) * tb->f2[itable] + (b * b * b - b) * tb->f2[itable + 1]) * tb->deltasq6;
    *fpair = factor_lj * value;
  } else {
    int_float_t rsq_lookup;
    rsq_lookup.f = (float)rsq;
    itable = rsq_lookup.i & tb->nmask;
    itable >>= tb->nshiftbits;
    fraction = (rsq_lookup.f - tb->rsq[itable]) * tb->drsq[itable];
    value = tb->f[itable] + fraction * tb->df[itable];
    *fpair = factor_lj * value;
  }
  if (s->tabstyle == ORC_LOOKUP)
    *evdwl = tb->e[itable];
  else if (s->tabstyle == ORC_LINEAR || s->tabstyle == ORC_BITMAP)
    *evdwl = tb->e[itable] + fraction * tb->de[itable];
  else
    *evdwl = a * tb->e[itable] + b * tb->e[itable + 1] +
        ((a * a * a - a) * tb->e2[itable] + (b * b * b - b) * tb->e2[itable + 1]) * tb->deltasq6;
  *evdwl *= factor_lj;
  return 0;
}

/* [stock] Pair::ev_tally, global accumulators only.  As shipped the
 * reference's virial stays zero with newton on; we tally per pair so the physically
 * meaningful number can be compared. */
static void ev_tally(orc_sys *s, int i, int j, double evdwl, double fpair, double delx, double dely, double delz) {
  int nlocal = s->nlocal;
  if (s->eflag) {
    if (s->newton_pair) s->eng_vdwl += evdwl;
    else {
      if (i < nlocal) s->eng_vdwl += 0.5 * evdwl;
      if (j < nlocal) s->eng_vdwl += 0.5 * evdwl;
    }
  }
  if (s->vflag) {
    double v[6] = {delx * delx * fpair, dely * dely * fpair, delz * delz * fpair,
                   delx * dely * fpair, delx * delz * fpair, dely * delz * fpair};
    if (s->newton_pair) for (int k = 0; k < 6; k++) s->virial[k] += v[k];
    else {
      if (i < nlocal) for (int k = 0; k < 6; k++) s->virial[k] += 0.5 * v[k];
      if (j < nlocal) for (int k = 0; k < 6; k++) s->virial[k] += 0.5 * v[k];
    }
  }
}
static void ev_init(orc_sys *s, int eflag, int vflag) {
  s->eflag = eflag; s->vflag = vflag; s->evflag = eflag || vflag;
  s->eng_vdwl = 0.0;
  for (int k = 0; k < 6; k++) s->virial[k] = 0.0;
}
double orc_eng_vdwl(orc_sys *s) { return s->eng_vdwl; }
void orc_virial(orc_sys *s, double v[6]) { for (int k = 0; k < 6; k++) v[k] = s->virial[k]; }

/* formal type of actual type t in substate k (1-state types map to themselves) */
static inline int formal_of(orc_sys *s, int t, int k) {
  return s->n_states[t] == 1 ? t : s->formal_from[t * 2 + k];
}

/* PairTable_UCGLD::compute: pair_table_ucgld.cpp:111-541.
 * Scenario 2 follows the intended `sj` keying. */
void orc_pair_ucgld(orc_sys *s, int eflag, int vflag) {
  ev_init(s, eflag, vflag);
  int nlocal = s->nlocal, newton_pair = s->newton_pair;
  int nt = s->n_formal + 1;
  double kT = s->kT;
  double *x = s->x, *f = s->f, *ucgf = s->ucgforce, *ucgl = s->ucgl, *sc = s->scores;
  int *type = s->type, *ucgstate = s->ucgstate;

  /* chemical-potential pre-add: :170-180 */
  for (int i = 0; i < nlocal; i++) {
    int itype = type[i];
    s->num_ucgstates[i] = s->n_states[itype];
    if (s->n_states[itype] > 1) {
      double mui = s->chem_pot[s->formal_from[itype * 2 + 1]] - s->chem_pot[s->formal_from[itype * 2 + 0]];
      ucgf[i] -= mui;
      sc[2 * i + 1] -= mui / kT;
    }
  }
  /* pair loop: :184-539 */
  for (int i = 0; i < nlocal; i++) {
    int itype = type[i], istate = ucgstate[i];
    double ldi = ucgl[i];
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    const int *jlist = s->neigh + s->firstneigh[i];
    int jnum = s->numneigh[i];
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj];
      double factor_lj = s->special_lj[(j >> SBBITS) & 3];
      j &= NEIGHMASK;
      int jtype = type[j], jstate = ucgstate[j];
      double ldj = ucgl[j];
      double delx = xtmp - x[3 * j], dely = ytmp - x[3 * j + 1], delz = ztmp - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq < s->cutsq[itype * nt + jtype]) {
        double evdwl = 0, fpair = 0;
        double u00 = 0, u01 = 0, u10 = 0, u11 = 0, fp00 = 0, fp01 = 0, fp10 = 0, fp11 = 0;
        int ni = s->n_states[itype], nj = s->n_states[jtype];
        int jok = (j < nlocal || newton_pair);
        if (ni == 1 && nj == 1) { /* scenario 1: :219-270 */
          if (table_eval(s, &s->tables[s->tabindex[itype * nt + jtype]], rsq, factor_lj, &evdwl, &fpair)) return;
        } else if (ni == 1 && nj > 1) { /* scenario 2: :274-345 */
          for (int sj = 0; sj < nj; sj++) {
            int jt = s->formal_from[jtype * 2 + sj];
            if (table_eval(s, &s->tables[s->tabindex[itype * nt + jt]], rsq, factor_lj, &evdwl, &fpair)) return;
            if (sj == 0) { fp00 = fpair; u00 = evdwl; } else { fp01 = fpair; u01 = evdwl; }
            if (jok) sc[2 * j + sj] -= evdwl / kT;
          }
          fpair = (1. - ldj) * fp00 + ldj * fp01;
          evdwl = (1. - ldj) * u00 + ldj * u01;
          if (jok) ucgf[j] -= u01 - u00;
        } else if (ni > 1 && nj == 1) { /* scenario 3: :349-421 */
          for (int si = 0; si < ni; si++) {
            int it = s->formal_from[itype * 2 + si];
            if (table_eval(s, &s->tables[s->tabindex[it * nt + jtype]], rsq, factor_lj, &evdwl, &fpair)) return;
            if (si == 0) { fp00 = fpair; u00 = evdwl; } else { fp10 = fpair; u10 = evdwl; }
            sc[2 * i + si] -= evdwl / kT;
          }
          fpair = (1. - ldi) * fp00 + ldi * fp10;
          evdwl = (1. - ldi) * u00 + ldi * u10;
          ucgf[i] -= u10 - u00;
        } else { /* scenario 4: :424-519 */
          for (int si = 0; si < ni; si++) {
            int it = s->formal_from[itype * 2 + si];
            for (int sj = 0; sj < nj; sj++) {
              int jt = s->formal_from[jtype * 2 + sj];
              if (table_eval(s, &s->tables[s->tabindex[it * nt + jt]], rsq, factor_lj, &evdwl, &fpair)) return;
              if (si == 0 && sj == 0) { fp00 = fpair; u00 = evdwl; }
              else if (si == 1 && sj == 0) { fp10 = fpair; u10 = evdwl; }
              else if (si == 0 && sj == 1) { fp01 = fpair; u01 = evdwl; }
              else { fp11 = fpair; u11 = evdwl; }
              if (sj == jstate) sc[2 * i + si] -= evdwl / kT;
              if (si == istate && jok) sc[2 * j + sj] -= evdwl / kT;
            }
          }
          evdwl = (1. - ldi) * (1. - ldj) * u00 + (1. - ldi) * ldj * u01 + (1. - ldj) * ldi * u10 + ldi * ldj * u11;
          fpair = (1. - ldi) * (1. - ldj) * fp00 + (1. - ldi) * ldj * fp01 + (1. - ldj) * ldi * fp10 + ldi * ldj * fp11;
          ucgf[i] -= ldj * (u11 - u01) + (1. - ldj) * (u10 - u00);
          if (jok) ucgf[j] -= ldi * (u11 - u10) + (1. - ldi) * (u01 - u00);
        }
        f[3 * i] += delx * fpair; f[3 * i + 1] += dely * fpair; f[3 * i + 2] += delz * fpair;
        if (jok) { f[3 * j] -= delx * fpair; f[3 * j + 1] -= dely * fpair; f[3 * j + 2] -= delz * fpair; }
        if (s->evflag) ev_tally(s, i, j, evdwl, fpair, delx, dely, delz);
      }
    }
  }
}

/* ------------------------------------------------------------ pair: bethe */
void orc_pair_bethe_config(orc_sys *s, int method, int pseudo, int prior, double noise, int seed) {
  s->pair_kind = 1;
  s->b_method = method; s->b_pseudo = pseudo; s->b_prior = prior; s->b_noise = noise; s->b_seed = seed;
  if (prior == 1) ranmars_init(&s->b_rng, seed);
}
/* prior_prob_from_type built in init_style(): pair_table_ucg_bethe.cpp:1056-1076:
 * softmax of -mu/kT over the substates */
static void prior_from_type(orc_sys *s, int t, double p[2]) {
  if (s->n_states[t] == 1) { p[0] = 1.0; p[1] = 0.0; return; }
  double e0 = exp(-s->chem_pot[s->formal_from[t * 2]] / s->kT);
  double e1 = exp(-s->chem_pot[s->formal_from[t * 2 + 1]] / s->kT);
  p[0] = e0 / (e0 + e1);
  p[1] = e1 / (e0 + e1);
}

/* PairTable_UCG_Bethe::compute: pair_table_ucg_bethe.cpp:88-630 */
void orc_pair_bethe(orc_sys *s, int eflag, int vflag) {
  const double EPSILONE = 1.0e-6;
  ev_init(s, eflag, vflag);
  int nlocal = s->nlocal, newton_pair = s->newton_pair, nall = s->nlocal + s->nghost;
  int nt = s->n_formal + 1;
  double kT = s->kT;
  double *x = s->x, *f = s->f, *ucgl = s->ucgl, *ucgp = s->ucgp, *sc = s->scores;
  int *type = s->type, *ucgstate = s->ucgstate;
  if (nall > s->b_nmax) {
    s->b_nmax = nall;
    s->prior_prob = (double *)realloc(s->prior_prob, 2 * (size_t)nall * sizeof(double));
    s->post_prob = (double *)realloc(s->post_prob, 2 * (size_t)nall * sizeof(double));
  }
  double *pp = s->prior_prob;
  memset(pp, 0, 2 * (size_t)nall * sizeof(double));
  memset(s->post_prob, 0, 2 * (size_t)nall * sizeof(double));
  /* score initialisation: :162-170 */
  for (int i = 0; i < nlocal; i++) {
    int itype = type[i];
    s->num_ucgstates[i] = s->n_states[itype];
    for (int si = 0; si < s->n_states[itype]; si++)
      sc[2 * i + si] = -s->chem_pot[formal_of(s, itype, si)] / kT;
  }
  for (int i = 0; i < nlocal; i++) {
    int itype = type[i], istate = ucgstate[i];
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    const int *jlist = s->neigh + s->firstneigh[i];
    int jnum = s->numneigh[i];
    /* prior of i: :179-206 */
    if (ucgp[i] < -0.999) {
      if (s->n_states[itype] > 1) {
        if (s->b_prior == 0) prior_from_type(s, itype, &pp[2 * i]);
        else if (s->b_prior == 1) {
          double pt[2]; prior_from_type(s, itype, pt);
          double r = (ranmars_uniform(&s->b_rng) - 0.5) * 2;
          r *= s->b_noise;
          pp[2 * i] = fmin(0.999999, fmax(pt[0] + r, 0.0));
          pp[2 * i + 1] = 1. - pp[2 * i];
        } else { pp[2 * i] = 1.0 - ucgl[i]; pp[2 * i + 1] = ucgl[i]; }
      } else pp[2 * i] = 1.0;
    } else {
      if (s->n_states[itype] > 1) { pp[2 * i + 1] = ucgl[i]; pp[2 * i] = 1.0 - ucgl[i]; }
      else pp[2 * i] = 1.0;
    }
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj];
      double factor_lj = s->special_lj[(j >> SBBITS) & 3];
      j &= NEIGHMASK;
      int jtype = type[j], jstate = ucgstate[j];
      double delx = xtmp - x[3 * j], dely = ytmp - x[3 * j + 1], delz = ztmp - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      /* prior of j: :223-253 (Q6: itype used for the type prior; Q7: ucgp not ucgl) */
      if (ucgp[j] < -0.999) {
        if (s->n_states[jtype] > 1) {
          if (s->b_prior == 0) { double pt[2]; prior_from_type(s, itype, pt); for (int si = 0; si < s->n_states[itype]; si++) pp[2 * j + si] = pt[si]; }
          else if (s->b_prior == 1) {
            double pt[2]; prior_from_type(s, jtype, pt);
            double r = (ranmars_uniform(&s->b_rng) - 0.5) * 2;
            r *= s->b_noise;
            pp[2 * j] = fmin(0.999999, fmax(pt[0] + r, 0.0));
            pp[2 * j + 1] = 1. - pp[2 * j];
          } else { pp[2 * j] = 1.0 - ucgl[j]; pp[2 * j + 1] = ucgl[j]; }
        } else pp[2 * j] = 1.0;
      } else {
        if (s->n_states[jtype] > 1) { pp[2 * j + 1] = ucgp[j]; pp[2 * j] = 1.0 - ucgp[j]; }
        else pp[2 * j] = 1.0;
      }
      if (rsq < s->cutsq[itype * nt + jtype]) {
        double evdwl = 0, fpair = 0;
        double u00 = 0, u01 = 0, u10 = 0, u11 = 0, fp00 = 0, fp01 = 0, fp10 = 0, fp11 = 0;
        int ni = s->n_states[itype], nj = s->n_states[jtype];
        int jok = (j < nlocal || newton_pair);
        if (ni == 1 && nj == 1) {
          if (table_eval(s, &s->tables[s->tabindex[itype * nt + jtype]], rsq, factor_lj, &evdwl, &fpair)) return;
        } else if (ni == 1 && nj > 1) { /* :310-385 (sj keying, Q1) */
          for (int sj = 0; sj < nj; sj++) {
            int jt = s->formal_from[jtype * 2 + sj];
            if (table_eval(s, &s->tables[s->tabindex[itype * nt + jt]], rsq, factor_lj, &evdwl, &fpair)) return;
            if (sj == 0) { fp00 = fpair; u00 = evdwl; } else { fp01 = fpair; u01 = evdwl; }
            if (jok) sc[2 * j + sj] -= evdwl / kT;
          }
          fpair = pp[2 * j] * fp00 + pp[2 * j + 1] * fp01;
          evdwl = pp[2 * j] * u00 + pp[2 * j + 1] * u01;
        } else if (ni > 1 && nj == 1) { /* :389-460 */
          for (int si = 0; si < ni; si++) {
            int it = s->formal_from[itype * 2 + si];
            if (table_eval(s, &s->tables[s->tabindex[it * nt + jtype]], rsq, factor_lj, &evdwl, &fpair)) return;
            if (si == 0) { fp00 = fpair; u00 = evdwl; } else { fp10 = fpair; u10 = evdwl; }
            sc[2 * i + si] -= evdwl / kT;
          }
          fpair = pp[2 * i] * fp00 + pp[2 * i + 1] * fp10;
          evdwl = pp[2 * i] * u00 + pp[2 * i + 1] * u10;
        } else { /* scenario 4: :463-605 */
          for (int si = 0; si < ni; si++) {
            int it = s->formal_from[itype * 2 + si];
            for (int sj = 0; sj < nj; sj++) {
              int jt = s->formal_from[jtype * 2 + sj];
              if (table_eval(s, &s->tables[s->tabindex[it * nt + jt]], rsq, factor_lj, &evdwl, &fpair)) return;
              if (si == 0 && sj == 0) { fp00 = fpair; u00 = evdwl; }
              else if (si == 1 && sj == 0) { fp10 = fpair; u10 = evdwl; }
              else if (si == 0 && sj == 1) { fp01 = fpair; u01 = evdwl; }
              else { fp11 = fpair; u11 = evdwl; }
              if (s->b_pseudo == 0) { /* the reference's flag value 0 selects the pseudo-likelihood tally */
                if (sj == jstate) sc[2 * i + si] -= evdwl / kT;
                if (si == istate && jok) sc[2 * j + sj] -= evdwl / kT;
              }
            }
          }
          double Jij = u11 + u00 - u01 - u10;
          if (Jij / kT < -709.0) Jij = -700.0 * kT;
          double bij = exp(-Jij / kT), aij = expm1(-Jij / kT);
          double pi0 = pp[2 * i], pi1 = pp[2 * i + 1], pj0 = pp[2 * j], pj1 = pp[2 * j + 1];
          double Qij = (pi1 + pj1) * aij + 1.;
          double Dij = Qij * Qij - 4. * aij * bij * pi1 * pj1;
          Dij = fmax(Dij, 0.0);
          double pij11 = 0;
          if (s->b_method == 1) {
            if (fabs(aij) < EPSILONE) pij11 = pi1 * pj1;
            else if (Qij < 0.0) pij11 = (Qij - sqrt(Dij)) / (2. * aij);
            else pij11 = (2. * bij * pi1 * pj1) / (Qij + sqrt(Dij));
          } else pij11 = pi1 * pj1;
          double pij00 = 1. + pij11 - pi1 - pj1, pij10 = pi1 - pij11, pij01 = pj1 - pij11;
          if (s->b_pseudo == 1) { /* SCE conditionals, literal: :583-601 */
            double pj0i0 = pij00 / pi0, pj0i1 = pij01 / pi0, pj1i0 = pij10 / pi1, pj1i1 = pij11 / pi1;
            double pi0j0 = pij00 / pj0, pi0j1 = pij10 / pj0, pi1j0 = pij01 / pj1, pi1j1 = pij11 / pj1;
            sc[2 * i] -= (pj0i0 * u00 + pj1i0 * u01) / kT;
            sc[2 * i + 1] -= (pj0i1 * u10 + pj1i1 * u11) / kT;
            if (jok) {
              sc[2 * j] -= (pi0j0 * u00 + pi0j1 * u01) / kT;
              sc[2 * j + 1] -= (pi1j0 * u10 + pi1j1 * u11) / kT;
            }
          }
          evdwl = pij00 * u00 + pij01 * u01 + pij10 * u10 + pij11 * u11;
          fpair = pij00 * fp00 + pij01 * fp01 + pij10 * fp10 + pij11 * fp11;
        }
        f[3 * i] += delx * fpair; f[3 * i + 1] += dely * fpair; f[3 * i + 2] += delz * fpair;
        if (jok) { f[3 * j] -= delx * fpair; f[3 * j + 1] -= dely * fpair; f[3 * j + 2] -= delz * fpair; }
        if (s->evflag) ev_tally(s, i, j, evdwl, fpair, delx, dely, delz);
      }
    }
  }
}

/* -------------------------------------------------------------------- fixes */
void orc_fix_clear(orc_sys *s) {
  for (int i = 0; i < s->nfix; i++) { free(s->fix[i].gfactor1); free(s->fix[i].gfactor2); }
  memset(s->fix, 0, sizeof(s->fix));
  s->nfix = 0;
}
void orc_fix_ttarget(orc_sys *s, double T) {
  FixDesc *fx = &s->fix[s->nfix++];
  fx->kind = ORC_FIX_TTARGET; fx->t_target = T;
}
void orc_fix_nve(orc_sys *s, int groupbit) {
  FixDesc *fx = &s->fix[s->nfix++];
  fx->kind = ORC_FIX_NVE; fx->groupbit = groupbit;
}
void orc_fix_nve_wall(orc_sys *s, int groupbit, int bias_flag, double barrier) {
  FixDesc *fx = &s->fix[s->nfix++];
  fx->kind = ORC_FIX_NVE_WALL; fx->groupbit = groupbit; fx->bias_flag = bias_flag; fx->barrier = barrier;
}
/* Fix_UCGLD_Langevin ctor: fix_ucgld_langevin.cpp:55-122 */
void orc_fix_langevin(orc_sys *s, int groupbit, double t_start, double t_stop, double t_period, int seed) {
  FixDesc *fx = &s->fix[s->nfix++];
  fx->kind = ORC_FIX_LANGEVIN; fx->groupbit = groupbit;
  fx->t_start = t_start; fx->t_target = t_start; fx->t_stop = t_stop; fx->t_period = t_period; fx->seed = seed;
  ranmars_init(&fx->rng, seed + 0 /* comm->me */);
}
/* FixUCGState ctor: fix_ucgstate.cpp:34-71 */
void orc_fix_ucgstate(orc_sys *s, int mode, int seed, double rate) {
  FixDesc *fx = &s->fix[s->nfix++];
  fx->kind = ORC_FIX_UCGSTATE; fx->mode = mode; fx->seed = seed; fx->rate = rate;
  if (mode == 2) ranmars_init(&fx->rng, seed + 0);
}

/* first fix exporting "t_target": pair_table_ucgld.cpp:876-881, fix_ucgstate.cpp:148-156 */
static int find_ttarget(orc_sys *s, double *T) {
  for (int i = 0; i < s->nfix; i++)
    if (s->fix[i].kind == ORC_FIX_TTARGET || s->fix[i].kind == ORC_FIX_LANGEVIN) { *T = s->fix[i].t_target; return 1; }
  return 0;
}

/* FixNVE_UCGLD::initial_integrate (fix_nve_ucgld.cpp:83-99, per-type mass branch);
 * wall: fix_nve_ucgld_wall_hard.cpp:109-134 */
void orc_nve_initial(orc_sys *s, int groupbit, int wall) {
  double dtv = s->dt, dtf = 0.5 * s->dt * s->ftm2v;
  for (int i = 0; i < s->nlocal; i++) {
    if (s->mask[i] & groupbit) {
      double dtfm = dtf / s->mass[s->type[i]];
      s->v[3 * i] += dtfm * s->f[3 * i];
      s->v[3 * i + 1] += dtfm * s->f[3 * i + 1];
      s->v[3 * i + 2] += dtfm * s->f[3 * i + 2];
      s->x[3 * i] += dtv * s->v[3 * i];
      s->x[3 * i + 1] += dtv * s->v[3 * i + 1];
      s->x[3 * i + 2] += dtv * s->v[3 * i + 2];
      double dtflm = dtf / s->ucgml[i];
      s->ucgvl[i] += dtflm * s->ucgforce[i];
      s->ucgl[i] += dtv * s->ucgvl[i];
      if (wall) s->ucgstate[i] = (s->ucgl[i] < 0.5) ? 0 : 1;
    }
  }
}
/* ::final_integrate (fix_nve_ucgld.cpp:139-151); wall reflection
 * fix_nve_ucgld_wall_hard.cpp:182-202 */
void orc_nve_final(orc_sys *s, int groupbit, int wall) {
  double dtf = 0.5 * s->dt * s->ftm2v;
  for (int i = 0; i < s->nlocal; i++) {
    if (s->mask[i] & groupbit) {
      double dtfm = dtf / s->mass[s->type[i]];
      s->v[3 * i] += dtfm * s->f[3 * i];
      s->v[3 * i + 1] += dtfm * s->f[3 * i + 1];
      s->v[3 * i + 2] += dtfm * s->f[3 * i + 2];
      double dtflm = dtf / s->ucgml[i];
      s->ucgvl[i] += dtflm * s->ucgforce[i];
      if (wall) {
        if (s->ucgl[i] < 0.0) { s->ucgl[i] = -s->ucgl[i]; s->ucgvl[i] = -s->ucgvl[i]; }
        else if (s->ucgl[i] > 1.0) { s->ucgl[i] = 2.0 - s->ucgl[i]; s->ucgvl[i] = -s->ucgvl[i]; }
      }
    }
  }
}
/* bias_force + post_force: fix_nve_ucgld_wall_hard.cpp:234-257 */
void orc_wall_bias(orc_sys *s, int groupbit, double H) {
  for (int i = 0; i < s->nlocal; i++)
    if (s->mask[i] & groupbit) {
      double x = s->ucgl[i] - 0.5;
      s->ucgforce[i] += (-7980 * x * x * x * x * x * x * x * x * x + 2 * x) * 10 * H;
    }
}
/* FixUCGState::post_force: fix_ucgstate.cpp:88-132 */
static void ucgstate_post_force(orc_sys *s, int mode, double rate, RanMars *rng) {
  int ld_flag = mode == 1, mc_flag = mode == 2;
  for (int i = 0; i < s->nlocal; i++) {
    if (s->num_ucgstates[i] == 1) {
      if (!ld_flag) s->ucgstate[i] = 0;
      s->ucgp[i] = 1.0;
    } else {
      double denom = 0.0, e[2] = {1.0, 1.0};
      for (int si = 0; si < s->num_ucgstates[i]; si++) {
        e[si] = exp(fmin(s->scores[2 * i + si], 700.0));
        denom += e[si];
      }
      s->ucgp[i] = fmin(1.0 - 1e-6, fmax(1e-6, e[1] / denom));
      if (!ld_flag) {
        if (mc_flag) {
          double mc_factor;
          if (s->ucgstate[i] == 0) mc_factor = s->ucgp[i] / (1.0 - s->ucgp[i]);
          else mc_factor = (1.0 - s->ucgp[i]) / s->ucgp[i];
          mc_factor = fmin(mc_factor, 1.0) * rate;
          double mc_rand = ranmars_uniform(rng);
          s->ucgstate[i] = (mc_rand < mc_factor) ? 0 : 1;
        } else
          s->ucgstate[i] = (int)round(s->ucgp[i]);
      }
    }
    if (!ld_flag) s->ucgl[i] = s->ucgp[i];
  }
}
void orc_ucgstate_post_force(orc_sys *s, int mode, double rate) {
  RanMars r; ranmars_init(&r, 12345);
  ucgstate_post_force(s, mode, rate, &r);
}

/* Fix_UCGLD_Langevin::init gfactors (fix_ucgld_langevin.cpp:164-171, incl. the
 * ucgml[type index] quirk Q20) */
static void langevin_init(orc_sys *s, FixDesc *fx) {
  int nt = s->n_formal;
  fx->gfactor1 = (double *)realloc(fx->gfactor1, (nt + 1) * sizeof(double));
  fx->gfactor2 = (double *)realloc(fx->gfactor2, (nt + 1) * sizeof(double));
  for (int i = 1; i <= nt; i++) {
    double ml = (i < s->nlocal) ? s->ucgml[i] : s->ucgml[0];
    fx->gfactor1[i] = -ml / fx->t_period / s->ftm2v;
    fx->gfactor2[i] = sqrt(ml) / s->ftm2v;
    fx->gfactor2[i] *= sqrt(24.0 * s->boltz / fx->t_period / s->dt / s->mvv2e);
  }
}
/* compute_target (:318-353, CONSTANT) + post_force_templated<0> (:226-297) */
static void langevin_post_force(orc_sys *s, FixDesc *fx) {
  double delta = (double)(s->ntimestep - s->beginstep);
  if (delta != 0.0) delta /= (double)(s->endstep - s->beginstep);
  fx->t_target = fx->t_start + delta * (fx->t_stop - fx->t_start);
  fx->tsqrt = sqrt(fx->t_target);
  for (int i = 0; i < s->nlocal; i++)
    if (s->mask[i] & fx->groupbit) {
      double gamma1 = fx->gfactor1[s->type[i]];
      double gamma2 = fx->gfactor2[s->type[i]] * fx->tsqrt;
      double fran = gamma2 * (ranmars_uniform(&fx->rng) - 0.5);
      double fdrag = gamma1 * s->ucgvl[i];
      s->ucgforce[i] += fdrag + fran;
    }
}
/* end_of_step: :303-312 */
static void langevin_end_of_step(orc_sys *s, FixDesc *fx) {
  double ek = 0.0;
  for (int i = 0; i < s->nlocal; i++)
    if (s->mask[i] & fx->groupbit) ek += 0.5 * s->ucgml[i] * s->ucgvl[i] * s->ucgvl[i] * s->mvv2e;
  fx->lambda_temp = ek / (0.5 * s->boltz * s->nlocal);
}
double orc_lambda_temp(orc_sys *s) {
  for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == ORC_FIX_LANGEVIN) return s->fix[i].lambda_temp;
  return 0.0;
}

/* ------------------------------------------------------------------- Verlet */
static void pair_compute(orc_sys *s, int eflag, int vflag) {
  if (s->pair_kind == 1) orc_pair_bethe(s, eflag, vflag);
  else orc_pair_ucgld(s, eflag, vflag);
}

/* [stock] Verlet::setup */
void orc_setup(orc_sys *s, int eflag, int vflag) {
  /* init(): pair kT from the first t_target provider, fix init() */
  double T;
  if (!s->kT_set) { if (find_ttarget(s, &T)) s->kT = s->boltz * T; else set_err(s, "no fix exports t_target (Q2)"); }
  for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == ORC_FIX_LANGEVIN) langevin_init(s, &s->fix[i]);
  orc_pbc(s);
  orc_borders(s);
  orc_neigh_build(s);
  s->nbuilds = 0;
  orc_force_clear(s);
  pair_compute(s, eflag, vflag);
  if (s->newton_pair) orc_reverse_comm(s);
  /* modify->setup(): fix setup() in definition order */
  for (int i = 0; i < s->nfix; i++) {
    FixDesc *fx = &s->fix[i];
    if (fx->kind == ORC_FIX_LANGEVIN) langevin_post_force(s, fx);       /* setup -> post_force :187-197 */
    else if (fx->kind == ORC_FIX_UCGSTATE) {                             /* fix_ucgstate.cpp:142-171 */
      int found = 0;
      for (int k = 0; k < i; k++) if (s->fix[k].kind == ORC_FIX_TTARGET || s->fix[k].kind == ORC_FIX_LANGEVIN) found = 1;
      if (!found && !find_ttarget(s, &T)) set_err(s, "FixUCGState requires a thermostat fix BEFORE ITSELF");
      ucgstate_post_force(s, fx->mode, fx->rate, &fx->rng);
    }
    /* the wall fix has no setup(); [stock] Fix::setup is a no-op, so the bias is not
       applied at step 0 */
  }
}

/* [stock] Verlet::run */
void orc_run(orc_sys *s, int nsteps, int thermo_every) {
  s->beginstep = s->ntimestep;
  s->endstep = s->ntimestep + nsteps;
  for (int k = 0; k < 4; k++) s->timers[k] = 0.0;
  for (int n = 0; n < nsteps; n++) {
    s->ntimestep++;
    int ev = thermo_every > 0 && (s->ntimestep % thermo_every == 0);
    double t0 = now();
    for (int i = 0; i < s->nfix; i++) {
      FixDesc *fx = &s->fix[i];
      if (fx->kind == ORC_FIX_NVE) orc_nve_initial(s, fx->groupbit, 0);
      else if (fx->kind == ORC_FIX_NVE_WALL) orc_nve_initial(s, fx->groupbit, 1);
    }
    double t1 = now();
    s->timers[3] += t1 - t0;
    int nflag = orc_neigh_decide(s);
    double t2 = now();
    s->timers[1] += t2 - t1;
    if (nflag == 0) {
      orc_forward_comm(s);
      s->timers[2] += now() - t2;
    } else {
      orc_pbc(s);
      orc_borders(s);
      double t3 = now();
      s->timers[2] += t3 - t2;
      orc_neigh_build(s);
      s->timers[1] += now() - t3;
    }
    double t4 = now();
    orc_force_clear(s);
    pair_compute(s, ev, ev);
    double t5 = now();
    s->timers[0] += t5 - t4;
    if (s->newton_pair) orc_reverse_comm(s);
    double t6 = now();
    s->timers[2] += t6 - t5;
    for (int i = 0; i < s->nfix; i++) { /* post_force in definition order */
      FixDesc *fx = &s->fix[i];
      if (fx->kind == ORC_FIX_LANGEVIN) langevin_post_force(s, fx);
      else if (fx->kind == ORC_FIX_UCGSTATE) ucgstate_post_force(s, fx->mode, fx->rate, &fx->rng);
      else if (fx->kind == ORC_FIX_NVE_WALL && fx->bias_flag) orc_wall_bias(s, fx->groupbit, fx->barrier);
    }
    for (int i = 0; i < s->nfix; i++) {
      FixDesc *fx = &s->fix[i];
      if (fx->kind == ORC_FIX_NVE) orc_nve_final(s, fx->groupbit, 0);
      else if (fx->kind == ORC_FIX_NVE_WALL) orc_nve_final(s, fx->groupbit, 1);
    }
    for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == ORC_FIX_LANGEVIN) langevin_end_of_step(s, &s->fix[i]);
    s->timers[3] += now() - t6;
    if (s->err[0]) return;
  }
}
long long orc_ntimestep(orc_sys *s) { return s->ntimestep; }
int orc_nbuilds(orc_sys *s) { return s->nbuilds; }
void orc_timers(orc_sys *s, double out[4]) { for (int k = 0; k < 4; k++) out[k] = s->timers[k]; }
