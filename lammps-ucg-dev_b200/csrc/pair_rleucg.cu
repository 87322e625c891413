// pair_rleucg.cu — PairTable_RLEUCG_INTERFACE::compute (UCG/pair_table_rleucg_interface.cpp:177-505)
// for sm_100a: the original RLE-UCG mean-field local-density style (full list, newton off).
//
// Atom types are STATE types: a 2-state site of base type t interacts through tables
// tabindex[t+a][t'+b] weighted p_a p_b (:339,354,400).  Three sweeps over the full rows:
//   1. k_rle_density   rho_i = sum_j 1/2 (1 - tanh((r - r_th)/(0.1 r_th)))               (:232-276, :165-168)
//                      p_i = 1/2 + 1/2 tanh((rho - rho_th)/(0.1 rho_th)), dp/drho       (:88-99)
//      + ghost refresh of (p, dp/drho)                         == comm->forward_comm(this) (:278)
//   2. k_rle_pair      one-body probability force -kT ln(p/(1-p)) - mu                      (:296-322)
//                      f_i += d * sum_ab p_a p_b f_ab, E, virial, pair part of the
//                      probability force                                                   (:326-442)
//      + ghost refresh of cvf = F_p * dp/drho
//   3. k_rle_back      CV back-force: the reference scatters -fpair*d to j and reverse-
//                      communicates (:448-502); with the full symmetric list the same sum
//                      arrives at the centre site as
//                      f_i += sum_j (cvf_i g_i(r) + cvf_j g_j(r))/r * d,  g = |d prox/dr|   (:170-174)
// As-is quirks reproduced: the density->probability map exists for actual
// type 1 only (else UCGB200_ERR_DENSITY_TYPE); the pair part of the probability force and the
// full virial weight are only tallied for GHOST neighbors / halved for ghost neighbors exactly
// as the reference's branches do, so single-domain results equal the reference's single-rank
// run.  Deviation (Q16): the pair energy is always evaluated — the reference feeds a stale
// `evdwl` into the probability force on steps without eflag.
#include "pair_common.cuh"

#include <algorithm>
#include <cmath>

using namespace ucg;

namespace {

struct RleType {   // per STATE type
  int actual, nstates, entropy, pad;
  double mu;       // chemical_potentials[t]
  double cv_th;    // cv_thresholds[actual]
  double r_th;     // threshold_radii[actual]
};

struct RleArgs {
  const double4 *pos;
  const int *ts;
  const int *tag;
  int nlocal;
  const int *neigh;
  int stride;
  const int *numneigh;
  const PairInfo *pinfo;   // per state-type pair: cutsq, tab[0]
  const RleType *rt;
  const int *tabindex;     // [(nt)*(nt)]
  int nt;                  // ntypes + 1
  const TableDev *tables;
  double special_lj[4];
  double kT;
  double *prob, *partial, *cvf;   // [nall]
  double4 *frc;
  double *partials;
  ErrWord *err;
  // per-atom tallies ([stock] ev_tally with newton off: half of every visit's energy / virial to the centre site, half to
  // a LOCAL partner), nullptr: not asked for.  The visits are symmetric, so a site's share is what it accumulates anyway.
  double *eatom, *vatom;
  FastTable ft;   // shared-memory table path (SM = true): slot k of a row = table k
  GatherTex gt;   // neighbor gathers through the texture pipe (0 = plain loads)
};

__device__ __forceinline__ double prox(double r, double rth) {
  return 0.5 * (1.0 - tanh((r - rth) / (0.1 * rth)));
}
__device__ __forceinline__ double prox_der(double r, double rth) {
  const double t = tanh((r - rth) / (0.1 * rth));
  return 0.5 * (1.0 - t * t) / (0.1 * rth);
}

template <int LPA, int BS>
__global__ void __launch_bounds__(BS) k_rle_density(RleArgs p) {
  const int gid = (blockIdx.x * BS + threadIdx.x) / LPA;
  const int sub = threadIdx.x % LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : 0;
  const double4 ri = p.pos[i];
  const int ti = p.ts[i] & 0xffff;
  const RleType rti = p.rt[ti];
  const int jnum = (active && rti.nstates > 1) ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.nt;
  double rho = 0.0;
  // two-deep software pipeline (row index two entries ahead, record and type one ahead): the sweep was latency-bound
  // (long-scoreboard stalls 5.9 per issue); same visits in the same order
  {
    int jj = sub, j = -1, tsj = 0;
    double4 rj = ri;
    if (jj < jnum) { j = row[rowslot(jj)] & UCG_NEIGHMASK; rj = p.pos[j]; tsj = p.ts[j]; }
    int jnext = (jj + LPA < jnum) ? (row[rowslot(jj + LPA)] & UCG_NEIGHMASK) : -1;
    while (j >= 0) {
      const int jn = jnext;
      double4 rn = rj;
      int tn = 0;
      if (jn >= 0) { rn = p.pos[jn]; tn = p.ts[jn]; }
      jj += LPA;
      jnext = (jj + LPA < jnum) ? (row[rowslot(jj + LPA)] & UCG_NEIGHMASK) : -1;
      const int tj = tsj & 0xffff;
      const double rsq = rsq_exact(ri.x - rj.x, ri.y - rj.y, ri.z - rj.z);
      if (rsq < prow[tj].cutsq) rho += prox(sqrt(rsq), rti.r_th);
      j = jn; rj = rn; tsj = tn;
    }
  }
  rho = group_sum<LPA>(rho);
  if (active && sub == 0) {
    double pr = 1.0, pa = 0.0;
    if (rti.nstates > 1) {
      if (rti.actual != 1) report_error(p.err, UCGB200_ERR_DENSITY_TYPE, p.tag[i], 0, rho);
      const double t = tanh((rho - rti.cv_th) / (0.1 * rti.cv_th));
      pr = 0.5 + 0.5 * t;
      pa = 0.5 * (1.0 - t * t) / (0.1 * rti.cv_th);
    }
    p.prob[i] = pr;
    p.partial[i] = pa;
  }
}

// SM = true: every uploaded table is LINEAR on one rsq grid and there are at most 4 of them — their rows
// are interleaved in shared memory (slot = table index) and the CTAs are persistent; otherwise tables
// come through L1.
template <int LPA, int BS, bool SM>
__global__ void __launch_bounds__(BS) k_rle_pair(RleArgs p) {
  extern __shared__ double2 s_tab[];
  if (SM) {
    const int nwords = p.ft.tablen * p.ft.W;
    for (int k = threadIdx.x; k < nwords; k += BS) s_tab[k] = p.ft.table[k];
    __syncthreads();
  }
  const int sub = threadIdx.x % LPA;
  constexpr int GROUPS = BS / LPA;
  double evacc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int base = blockIdx.x * GROUPS; base < p.nlocal; base += gridDim.x * GROUPS) {
  const int gid = base + threadIdx.x / LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : p.nlocal - 1;
  const double4 ri = p.pos[i];
  const int ti = p.ts[i] & 0xffff;
  const RleType rti = p.rt[ti];
  const int ni = rti.nstates;
  const double pi0 = p.prob[i];
  const int jnum = active ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.nt;
  double fx = 0, fy = 0, fz = 0, pf = 0, eacc = 0;
  double vir[6] = {0, 0, 0, 0, 0, 0};
  RowWalk<LPA> rw(row, sub, jnum);
  for (int jj = sub; jj < jnum; jj += LPA, rw.advance()) {
    const int jraw = rw.raw(jj);
    const double factor_lj = p.special_lj[(jraw >> UCG_SBBITS) & 3];
    const int j = jraw & UCG_NEIGHMASK;
    const double4 rj = gather_pos(p.gt, p.pos, j);
    const int tj = gather_ts(p.gt, p.ts, j) & 0xffff;
    const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
    const double rsq = rsq_exact(dx, dy, dz);
    if (rsq < prow[tj].cutsq) {
      const int nj = p.rt[tj].nstates;
      const double pj0 = p.prob[j];
      const bool jlocal = j < p.nlocal;
      double elj = 0.0, pfv = 0.0;
      bool bad = false;
      int it = 0;
      double frac = 0.0;
      if (SM) {
        const int ec = fast_table_index(p.ft, rsq, it, frac);
        if (ec) { report_error(p.err, ec, p.tag[i], p.tag[j], rsq); bad = true; }
      }
      for (int a = 0; a < ni && !bad; a++) {
        const double pa = ni > 1 ? (a == 0 ? pi0 : 1.0 - pi0) : 1.0;
        for (int b = 0; b < nj; b++) {
          const double pb = nj > 1 ? (b == 0 ? pj0 : 1.0 - pj0) : 1.0;
          double e, f;
          const int tix = p.tabindex[(ti + a) * p.nt + (tj + b)];
          if (SM) fast_table_slot(s_tab, p.ft.W, it, frac, tix, e, f);
          else {
            const int ec = table_eval(p.tables[tix], rsq, e, f);
            if (ec) { report_error(p.err, ec, p.tag[i], p.tag[j], rsq); bad = true; break; }
          }
          e *= factor_lj;
          const double fp = factor_lj * f * pa * pb;
          fx += dx * fp; fy += dy * fp; fz += dz * fp;
          if (jlocal) { elj += e * pa * pb * 0.5; pfv += fp * 0.5; }
          else {
            elj += e * pa * pb; pfv += fp;
            if (ni > 1) pf += (a == 0) ? -pb * e : pb * e;   // (sic) ghost neighbors only, Q17
          }
        }
      }
      const double w = jlocal ? 1.0 : 0.5;   // ev_tally with newton off
      eacc += w * elj;
      vir[0] += w * dx * dx * pfv; vir[1] += w * dy * dy * pfv; vir[2] += w * dz * dz * pfv;
      vir[3] += w * dx * dy * pfv; vir[4] += w * dx * dz * pfv; vir[5] += w * dy * dz * pfv;
    }
  }
  fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
  pf = group_sum<LPA>(pf);
  eacc = group_sum<LPA>(eacc);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double v = group_sum<LPA>(vir[k]);
    if (active && sub == 0) { evacc[1 + k] += v; if (p.vatom) p.vatom[6 * (size_t)i + k] = v; }
  }
  if (active && sub == 0) {
    if (p.eatom) p.eatom[i] = eacc;
    double cvf = 0.0;
    if (ni > 1) {
      // one-body terms (:296-322)
      if (rti.entropy) pf -= p.kT * log(pi0);
      pf -= rti.mu;
      if (rti.entropy) pf += p.kT * log(1.0 - pi0);
      cvf = pf * p.partial[i];
    }
    p.cvf[i] = cvf;
    p.frc[i] = make_double4(fx, fy, fz, 0.0);
    evacc[0] += eacc;
  }
  }   // persistent loop over site groups
  block_reduce_store<7, BS>(evacc, p.partials);
}

// posc[j] = {x, y, z, c_j} for the uniform back-force sweep (pair_common.cuh): c_j = cvf_j of a 2-state site (owned or
// ghost: the CV forces were forwarded to the ghosts), else 0
__global__ void k_rle_posc(RleArgs p, int nall, double4 *__restrict__ posc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nall) return;
  double4 r = p.pos[j];
  r.w = p.rt[p.ts[j] & 0xffff].nstates > 1 ? p.cvf[j] : 0.0;
  posc[j] = r;
}

template <int LPA, int BS, bool PA>
__global__ void __launch_bounds__(BS) k_rle_back(RleArgs p) {
  const int gid = (blockIdx.x * BS + threadIdx.x) / LPA;
  const int sub = threadIdx.x % LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : 0;
  const double4 ri = p.pos[i];
  const int ti = p.ts[i] & 0xffff;
  const RleType rti = p.rt[ti];
  const double cvf_i = p.cvf[i];
  const int jnum = active ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.nt;
  double fx = 0, fy = 0, fz = 0;
  double vir[6] = {0, 0, 0, 0, 0, 0};
  double va[6] = {0, 0, 0, 0, 0, 0};   // PA: this site's share of the tallies (own visits + visits of local partners)
  for (int jj = sub; jj < jnum; jj += LPA) {
    const int j = row[rowslot(jj)] & UCG_NEIGHMASK;
    const double4 rj = p.pos[j];
    const int tj = p.ts[j] & 0xffff;
    const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
    const double rsq = rsq_exact(dx, dy, dz);
    if (rsq < prow[tj].cutsq) {
      const double r = sqrt(rsq);
      const RleType rtj = p.rt[tj];
      // g(r) of both sites: ONE tanh when they share the threshold radius (sites of one actual type: the usual case)
      const bool di = rti.nstates > 1, dj = rtj.nstates > 1;
      const double rth_a = di ? rti.r_th : rtj.r_th;
      const double g_a = (di || dj) ? prox_der(r, rth_a) : 0.0;
      const double g_j = (di && dj && rtj.r_th != rth_a) ? prox_der(r, rtj.r_th) : g_a;
      const double own = di ? cvf_i * g_a / r : 0.0;        // i's loop (:478-481)
      const double oth = dj ? p.cvf[j] * g_j / r : 0.0;     // what j's loop scatters to i
      const double fp = own + oth;
      fx += fp * dx; fy += fp * dy; fz += fp * dz;
      const double w = (j < p.nlocal ? 1.0 : 0.5) * own;   // ev_tally(i,j,...,fpair) in i's loop only (:488)
      vir[0] += w * dx * dx; vir[1] += w * dy * dy; vir[2] += w * dz * dz;
      vir[3] += w * dx * dy; vir[4] += w * dx * dz; vir[5] += w * dy * dz;
      if (PA) {
        const double wa = 0.5 * own + (j < p.nlocal ? 0.5 * oth : 0.0);
        va[0] += wa * dx * dx; va[1] += wa * dy * dy; va[2] += wa * dz * dz;
        va[3] += wa * dx * dy; va[4] += wa * dx * dz; va[5] += wa * dy * dz;
      }
    }
  }
  fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double v = group_sum<LPA>(vir[k]);
    if (active && sub == 0) ev[1 + k] = v;
    if (PA) {
      const double a = group_sum<LPA>(va[k]);
      if (active && sub == 0) p.vatom[6 * (size_t)i + k] += a;
    }
  }
  if (active && sub == 0) {
    double4 f = p.frc[i];
    f.x += fx; f.y += fy; f.z += fz;
    p.frc[i] = f;
  }
  block_reduce_store<7, BS>(ev, p.partials);
}

// forward_comm(this) to self: copy a per-site scalar from the owners to their local images
__global__ void k_ghost_scalar(double *__restrict__ a, double *__restrict__ b, int nlocal, int nlimg,
                               const int *__restrict__ owner, const int *__restrict__ slot_of_src) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nlimg) return;
  const int s = nlocal + slot_of_src[k], o = owner[k];
  a[s] = a[o];
  if (b) b[s] = b[o];
}

}  // namespace

extern "C" int ucgb200_pair_rleucg_configure(ucgb200_ctx *c, int ntypes, const int *actual_from_state, int n_actual,
                                             const int *n_states, const int *use_entropy, const double *cv_threshold,
                                             const double *threshold_radius, const double *chem_pot, const int *tabindex,
                                             const double *cutsq, const double *mass, double kT) {
  if (!c || ntypes < 1 || !actual_from_state || !n_states || !cv_threshold || !threshold_radius || !tabindex || !cutsq) return -1;
  if (!(kT > 0)) return fail(c, "pair_rleucg: kT must be positive (no fix exports t_target?)");
  auto &d = c->dens;
  d.n_types = ntypes;
  d.n_actual = n_actual;
  d.actual_from_state.assign(actual_from_state, actual_from_state + ntypes + 1);
  d.n_states_of_type.assign(n_states, n_states + n_actual + 1);
  d.use_entropy.assign(n_actual + 1, 0);
  if (use_entropy) d.use_entropy.assign(use_entropy, use_entropy + n_actual + 1);
  d.density_threshold.assign(cv_threshold, cv_threshold + n_actual + 1);
  d.threshold_radius.assign(threshold_radius, threshold_radius + n_actual + 1);
  d.chem_pot.assign(ntypes + 1, 0.0);
  if (chem_pot) d.chem_pot.assign(chem_pot, chem_pot + ntypes + 1);
  const int nt = ntypes + 1;
  d.tabindex.assign(tabindex, tabindex + nt * nt);
  d.cutsq.assign(cutsq, cutsq + nt * nt);
  d.mass.assign(nt, 1.0);
  if (mass) d.mass.assign(mass, mass + nt);
  d.T = kT;
  c->kT = kT;
  for (int t = 1; t <= ntypes; t++) {
    int a = d.actual_from_state[t];
    if (a < 1 || a > n_actual) return fail(c, "pair_rleucg: state type without an actual type");
  }
  d.set = true;
  c->maps_dirty = true;
  return 0;
}

// device maps for the RLE-UCG mode (called from rebuild_maps)
int ucg_rebuild_rle_maps(ucgb200_ctx *c) {
  auto &d = c->dens;
  const int nt = d.n_types + 1;
  const int ntab = (int)c->tables.size();
  std::vector<PairInfo> pi(nt * nt);
  std::vector<TypeInfo> ti(nt);
  std::vector<RleType> rt(nt);
  c->max_cut = 0.0;
  for (int t = 1; t < nt; t++) {
    const int a = d.actual_from_state[t];
    ti[t].nstates = 1; ti[t].mass = d.mass[t]; ti[t].mu0 = ti[t].mu1 = ti[t].dmu = 0.0;
    rt[t].actual = a; rt[t].nstates = d.n_states_of_type[a]; rt[t].entropy = d.use_entropy[a];
    rt[t].mu = d.chem_pot[t]; rt[t].cv_th = d.density_threshold[a]; rt[t].r_th = d.threshold_radius[a];
    // a 2-state base type t (the first state type of its actual type; sites always carry the
    // base type, pair_table_rleucg_interface.cpp:238-244) uses tables of types t and t+1
    const bool base = (t == 1) || d.actual_from_state[t - 1] != a;
    if (base && rt[t].nstates > 1 && t + rt[t].nstates - 1 > d.n_types) { c->err = "pair_rleucg: substates exceed ntypes"; return -1; }
  }
  for (int i = 1; i < nt; i++)
    for (int j = 1; j < nt; j++) {
      PairInfo &p = pi[i * nt + j];
      p.cutsq = d.cutsq[i * nt + j];
      double cut = std::sqrt(p.cutsq);
      if (c->cut_override > 0) cut = std::max(cut, c->cut_override);
      const double cn = cut + c->skin;
      p.cutneighsq = cn * cn;
      c->max_cut = std::max(c->max_cut, cut);
      p.ni = p.nj = 1;
      const int tix = d.tabindex[i * nt + j];
      if (p.cutsq > 0 && (tix < 0 || tix >= ntab)) { c->err = "pair_rleucg: tabindex refers to a table that was not uploaded"; return -1; }
      for (int k = 0; k < 4; k++) p.tab[k] = tix;
    }
  UCG_CHECK(c, c->d_typeinfo.ensure(nt));
  UCG_CHECK(c, cudaMemcpy(c->d_typeinfo.p, ti.data(), nt * sizeof(TypeInfo), cudaMemcpyHostToDevice));
  UCG_CHECK(c, c->d_pairinfo.ensure(nt * nt));
  UCG_CHECK(c, cudaMemcpy(c->d_pairinfo.p, pi.data(), nt * nt * sizeof(PairInfo), cudaMemcpyHostToDevice));
  c->h_pairinfo = pi;
  UCG_CHECK(c, c->d_tables.ensure(std::max(ntab, 1)));
  if (ntab) UCG_CHECK(c, cudaMemcpy(c->d_tables.p, c->tables.data(), ntab * sizeof(TableDev), cudaMemcpyHostToDevice));
  UCG_CHECK(c, d.d_rt.ensure(nt * sizeof(RleType)));
  UCG_CHECK(c, cudaMemcpy(d.d_rt.p, rt.data(), nt * sizeof(RleType), cudaMemcpyHostToDevice));
  UCG_CHECK(c, d.d_tabindex.ensure(nt * nt));
  UCG_CHECK(c, cudaMemcpy(d.d_tabindex.p, d.tabindex.data(), nt * nt * sizeof(int), cudaMemcpyHostToDevice));
  c->n_actual = d.n_types;   // the neighbor build indexes PairInfo by (state) type
  c->fast_uniform = false;
  // shared-memory table path: all tables LINEAR on one grid, at most 4 of them
  d.sm_tables = false;
  if (ntab >= 1 && ntab <= 4) {
    const TableDev &t0 = c->tables[0];
    bool ok = t0.style == UCGB200_TAB_LINEAR && (size_t)t0.n * ntab * sizeof(double2) <= 220 * 1024;
    for (int k = 1; k < ntab && ok; k++) {
      const TableDev &t = c->tables[k];
      ok = t.style == UCGB200_TAB_LINEAR && t.n == t0.n && t.innersq == t0.innersq && t.delta == t0.delta && t.invdelta == t0.invdelta;
    }
    if (ok) {
      const int n = t0.n;
      std::vector<double2> rows((size_t)n * ntab), tmp(n);
      for (int k = 0; k < ntab; k++) {
        UCG_CHECK(c, cudaMemcpy(tmp.data(), c->tables[k].ef, n * sizeof(double2), cudaMemcpyDeviceToHost));
        for (int r = 0; r < n; r++) rows[(size_t)r * ntab + k] = tmp[r];
      }
      UCG_CHECK(c, c->d_fast_table.ensure(rows.size()));
      UCG_CHECK(c, cudaMemcpy(c->d_fast_table.p, rows.data(), rows.size() * sizeof(double2), cudaMemcpyHostToDevice));
      c->fast_ntab = ntab; c->fast_len = n;
      d.sm_tables = true;
    }
  }
  c->maps_dirty = false;
  c->list_valid = false;
  return 0;
}

extern "C" int ucgb200_pair_rleucg(ucgb200_ctx *c, int eflag, int vflag) {
  if (!c) return -1;
  if (!c->dens.set) return fail(c, "pair_rleucg: not configured");
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (!c->list_valid) return fail(c, "pair_rleucg: neighbor list not built");
  c->ev_valid = false;
  if (c->nlocal == 0) {   // an empty brick still reports (zero) energy and virial
    UCG_CHECK(c, cudaMemsetAsync(c->d_ev.p, 0, 32 * sizeof(double), c->stream));
    c->ev_valid = true;
    c->ev_two_parts = false;
    return 0;
  }
  // the energy is always evaluated; bits 2 / 4 ask for the per-atom tallies (ucgb200_pair_peratom)
  const bool want_eatom = (eflag & 2) != 0, want_vatom = (vflag & 4) != 0;
  c->eatom_valid = c->vatom_valid = false;
  auto &d = c->dens;
  const int nall = c->nlocal + c->nghost;
  UCG_CHECK(c, d.d_prob.ensure(nall));
  UCG_CHECK(c, d.d_partial.ensure(nall));
  UCG_CHECK(c, d.d_cvforce.ensure(nall));
  constexpr int LPA = 8, BS = 256;
  const int nblk = nblocks((long long)c->nlocal * LPA, BS);
  UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 8 + 64));
  RleArgs a{};
  a.pos = c->pos.p; a.ts = c->ts.p; a.tag = c->tag.p; a.nlocal = c->nlocal;
  a.neigh = c->neigh.p; a.stride = c->neigh_stride; a.numneigh = c->numneigh.p;
  a.pinfo = c->d_pairinfo.p; a.rt = (const RleType *)d.d_rt.p; a.tabindex = d.d_tabindex.p; a.nt = d.n_types + 1;
  a.tables = c->d_tables.p;
  for (int k = 0; k < 4; k++) a.special_lj[k] = c->special_lj[k];
  a.kT = d.T;
  a.prob = d.d_prob.p; a.partial = d.d_partial.p; a.cvf = d.d_cvforce.p;
  a.frc = c->frc.p; a.partials = c->d_partials.p; a.err = c->d_err.p;
  if (want_eatom) { UCG_CHECK(c, c->d_eatom.ensure((size_t)c->nlocal + 8)); a.eatom = c->d_eatom.p; }
  if (want_vatom) { UCG_CHECK(c, c->d_vatom.ensure(6 * (size_t)c->nlocal + 8)); a.vatom = c->d_vatom.p; }
  const auto &h = c->halo;
  const int *lown = c->img_owner.p + h.nsend;
  if (c->timers_on) cudaEventRecord(c->ev_pair0, c->stream);
  k_rle_density<LPA, BS><<<nblk, BS, 0, c->stream>>>(a);
  UCG_LAUNCHED(c);
  if (h.nlimg) {
    k_ghost_scalar<<<nblocks(h.nlimg, 256), 256, 0, c->stream>>>(a.prob, a.partial, c->nlocal, h.nlimg, lown, c->slot_of_src.p);
    UCG_LAUNCHED(c);
  }
  if ((rc = ucg_mb_forward_scalars(c, a.prob, a.partial, nullptr))) return rc;   // ghosts owned by other bricks
  int nblk_pair = nblk;
  if (d.sm_tables && !(getenv("UCGB200_FORCE_GENERAL") && atoi(getenv("UCGB200_FORCE_GENERAL")))) {
    constexpr int FLPA = 4, FBS = 512;
    // a.gt stays 0: texture-pipe gathers were measured here and gave nothing (these sweeps are bound by FP64
    // transcendentals, not by the LSU data pipe); UCGB200_TEX_ALL=1 turns them on for experiments
    if (getenv("UCGB200_TEX_ALL") && atoi(getenv("UCGB200_TEX_ALL")) && (rc = ucg_bind_gather_textures(c, &a.gt.pos, &a.gt.ts))) return rc;
    const TableDev &t0 = c->tables[0];
    a.ft.table = c->d_fast_table.p; a.ft.tablen = c->fast_len; a.ft.W = c->fast_ntab;
    a.ft.innersq = t0.innersq; a.ft.delta = t0.delta; a.ft.invdelta = t0.invdelta;
    const size_t tab_bytes = (size_t)c->fast_len * c->fast_ntab * sizeof(double2);
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, c->device);
    nblk_pair = std::min(dev_sms, nblocks((long long)c->nlocal * FLPA, FBS));
    UCG_CHECK(c, c->d_partials.ensure((size_t)std::max(nblk, nblk_pair) * 8 + 64));
    a.partials = c->d_partials.p;
    auto kern = k_rle_pair<FLPA, FBS, true>;
    UCG_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
    kern<<<nblk_pair, FBS, tab_bytes, c->stream>>>(a);
  } else {
    k_rle_pair<LPA, BS, false><<<nblk, BS, 0, c->stream>>>(a);
  }
  UCG_LAUNCHED(c);
  if ((rc = reduce_partials(c, nblk_pair, 7, 0))) return rc;
  if (h.nlimg) {
    k_ghost_scalar<<<nblocks(h.nlimg, 256), 256, 0, c->stream>>>(a.cvf, nullptr, c->nlocal, h.nlimg, lown, c->slot_of_src.p);
    UCG_LAUNCHED(c);
  }
  if ((rc = ucg_mb_forward_scalars(c, a.cvf, nullptr, nullptr))) return rc;
  // uniform case (every 2-state type has the same threshold radius, every pair the same cutoff): the back-force sweep
  // gathers one packed record per neighbor (k_cv_back_fast, pair_common.cuh); UCGB200_CV_FAST=0 and per-atom virial
  // requests take the general sweep (same bits)
  double urth = -1.0, ucut = -1.0;
  bool uniform = !want_vatom && !(getenv("UCGB200_CV_FAST") && atoi(getenv("UCGB200_CV_FAST")) == 0);
  for (int t = 1; t < a.nt && uniform; t++) {
    const int act = d.actual_from_state[t];
    if (d.n_states_of_type[act] > 1) { const double r = d.threshold_radius[act]; if (urth < 0) urth = r; else if (r != urth) uniform = false; }
  }
  for (int i = 1; i < a.nt && uniform; i++)
    for (int j = 1; j < a.nt && uniform; j++) {
      const double cs = c->h_pairinfo[i * a.nt + j].cutsq;
      if (ucut < 0) ucut = cs; else if (cs != ucut) uniform = false;
    }
  if (uniform && urth > 0 && ucut > 0) {
    UCG_CHECK(c, c->posc.ensure((size_t)nall + 8));
    k_rle_posc<<<nblocks(nall, 256), 256, 0, c->stream>>>(a, nall, c->posc.p);
    UCG_LAUNCHED(c);
    CvBackArgs ba{c->posc.p, c->nlocal, a.neigh, a.stride, a.numneigh, ucut, urth, a.frc, a.partials};
    k_cv_back_fast<LPA, BS, 1><<<nblk, BS, 0, c->stream>>>(ba);
  } else if (want_vatom) k_rle_back<LPA, BS, true><<<nblk, BS, 0, c->stream>>>(a);
  else k_rle_back<LPA, BS, false><<<nblk, BS, 0, c->stream>>>(a);
  UCG_LAUNCHED(c);
  if ((rc = reduce_partials(c, nblk, 7, 16))) return rc;   // second virial part -> d_ev[16..22]
  if (c->timers_on) { cudaEventRecord(c->ev_pair1, c->stream); c->pair_timed = true; }
  // scores are not used by this style; ucgsoftmaxscores keep their cleared value
  UCG_CHECK(c, cudaMemsetAsync(c->scores.p, 0, (size_t)c->nlocal * sizeof(double2), c->stream));
  c->ev_valid = true;
  c->ev_two_parts = true;
  c->eatom_valid = want_eatom;
  c->vatom_valid = want_vatom;
  return 0;
}

// per-site substate probabilities of the last evaluation (diagnostics / tests), host order
extern "C" int ucgb200_pair_rleucg_probabilities(ucgb200_ctx *c, int cap, double *prob, double *cvforce) {
  if (!c || cap < c->nlocal) return -1;
  cudaSetDevice(c->device);
  auto &d = c->dens;
  const int n = c->nlocal;
  if (n == 0) return 0;
  std::vector<double> p(n), f(n);
  std::vector<int> orig(n);
  UCG_CHECK(c, cudaMemcpyAsync(p.data(), d.d_prob.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaMemcpyAsync(f.data(), d.d_cvforce.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaMemcpyAsync(orig.data(), c->orig.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  for (int s = 0; s < n; s++) {
    if (prob) prob[orig[s]] = p[s];
    if (cvforce) cvforce[orig[s]] = f[s];
  }
  return 0;
}
