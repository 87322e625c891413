// pair_bethe_density.cu — PairTable_UCG_Bethe_Density::compute
// (UCG/pair_table_ucg_bethe_density.cpp:133-758) for sm_100a: Bethe pair probabilities with
// density-dependent one-point priors.  Full list, newton off; types are ACTUAL types, formal
// types index tables / chemical potentials exactly as in table_ucgld (ucgb200_set_types).
//
// Three sweeps over the full rows, every result accumulated into the CENTRE site (no atomics):
//   1. k_bd_prior   rho_i = sum_j 1/2 (1 - tanh((r - r_th)/(0.1 r_th)))                    (:219-252)
//                   p_i0 = 1/2 + 1/2 tanh((rho - rho_th)/(0.1 rho_th)), dp/drho             (:103-113)
//                   non-density 2-state types: softmax(-mu/kT); 1-state: p0 = 1            (:255-273)
//      + ghost refresh of p_0            == comm->forward_comm(this)  (:280, repaired Q12)
//   2. k_bd_pair    one-body probability forces (:302-317), then per neighbor the 4 tables,
//                   J, b = exp(-J/kT), a = b - 1, Q, D, p11 (stable root, repaired Q14), E and
//                   fpair = sum p_ab f_ab, the pseudo-likelihood scores and the two-point
//                   probability forces -(u10-u00+kT ln(p10/p00)), -(u11-u01+kT ln(p11/p01)) (:604-655);
//                   posterior ucgp = softmax(scores)[1]                                     (:675-693)
//   3. k_bd_back    CV back-force with the proximity FUNCTION (sic, Q13): f_i += cvf_i g(r)/r d
//                   and the reaction -cvf_j g(r)/r d_ji of every LOCAL neighbor j           (:696-726)
// As-is behaviour kept (it is what the reference computes on one rank, periodic images being
// ghosts): with newton off the reference only updates f[j] for j < nlocal, so reactions onto
// ghost neighbors are dropped (pass 3, and the CG-CG / UCG-CG scenarios); the entropy term uses
// the LIST length numneigh[i] (1 - jnum, :294,309), not the in-cutoff count; the density map
// exists for actual type 1 only (Q15).  Every visit of a UCG-UCG pair tallies half of E and of
// the virial (:622-625 with ev_tally newton off), i.e. each pair once in total.
#include "pair_common.cuh"

#include <algorithm>
#include <cmath>

using namespace ucg;

namespace {

struct BdType {   // per ACTUAL type
  int use_density, entropy;
  double cv_th, r_th;
};

struct BdArgs {
  const double4 *pos;
  const int *ts;
  const int *tag;
  int nlocal;
  const int *neigh;
  int stride;
  const int *numneigh;
  const PairInfo *pinfo;
  const TypeInfo *tinfo;
  const BdType *bt;
  int na;
  const TableDev *tables;
  double special_lj[4];
  double kT, inv_kT;
  double *prob0, *partial0, *cvf;   // [nall] / [nlocal] / [nlocal]
  double4 *frc;
  double2 *scores;
  double *ucgp;
  double *partials;
  ErrWord *err;
  // per-atom tallies ([stock] ev_tally with newton off: half of every visit's energy / virial to the centre site, half to
  // a LOCAL partner), nullptr: not asked for
  double *eatom, *vatom;
  int logsum;     // 0 one logarithm per ratio, 1 (default) running product with exponent renormalisation, 2 one logarithm per 4 ratios
  FastTable ft;   // shared-memory table path (W > 0)
  GatherTex gt;   // neighbor gathers through the texture pipe (0 = plain loads)
};

__device__ __forceinline__ double bd_prox(double r, double rth) {
  return 0.5 * (1.0 - tanh((r - rth) / (0.1 * rth)));
}

template <int LPA, int BS>
__global__ void __launch_bounds__(BS) k_bd_prior(BdArgs p) {
  const int gid = (blockIdx.x * BS + threadIdx.x) / LPA;
  const int sub = threadIdx.x % LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : 0;
  const double4 ri = p.pos[i];
  const int ti = p.ts[i] & 0xffff;
  const TypeInfo tyi = p.tinfo[ti];
  const BdType bti = p.bt[ti];
  const bool dens = bti.use_density == 1 && tyi.nstates > 1;
  const int jnum = (active && dens) ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.na;
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
      if (rsq < prow[tj].cutsq) rho += bd_prox(sqrt(rsq), bti.r_th);
      j = jn; rj = rn; tsj = tn;
    }
  }
  rho = group_sum<LPA>(rho);
  if (active && sub == 0) {
    double p0 = 1.0, pa = 0.0;
    if (dens) {
      if (ti != 1) report_error(p.err, UCGB200_ERR_DENSITY_TYPE, p.tag[i], 0, rho);
      const double t = tanh((rho - bti.cv_th) / (0.1 * bti.cv_th));
      p0 = 0.5 + 0.5 * t;
      pa = 0.5 * (1.0 - t * t) / (0.1 * bti.cv_th);
    } else if (tyi.nstates > 1) {
      const double e0 = exp(-tyi.mu0 / p.kT), e1 = exp(-tyi.mu1 / p.kT);
      p0 = e0 / (e0 + e1);
    }
    p.prob0[i] = p0;
    p.partial0[i] = pa;
  }
}

// forward_comm(this): priors of the owners to their periodic images
__global__ void k_bd_ghost(double *__restrict__ a, int nlocal, int nlimg, const int *__restrict__ owner,
                           const int *__restrict__ slot_of_src) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nlimg) return;
  a[nlocal + slot_of_src[k]] = a[owner[k]];
}

// W = 0: tables through L1 (any type system, any table style); W = 3 / 4: the interleaved LINEAR tables
// of the single 2-state type staged in shared memory, persistent CTAs (one per SM)
template <int LPA, int BS, int W, bool PA>
__global__ void __launch_bounds__(BS) k_bd_pair(BdArgs p) {
  extern __shared__ double2 s_tab[];
  if (W) fast_table_stage<W, BS>(s_tab, p.ft);
  const int sub = threadIdx.x % LPA;
  constexpr int GROUPS = BS / LPA;
  double evacc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int base = blockIdx.x * GROUPS; base < p.nlocal; base += gridDim.x * GROUPS) {
  const int gid = base + threadIdx.x / LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : p.nlocal - 1;
  const double4 ri = p.pos[i];
  const int ti = p.ts[i] & 0xffff;
  const TypeInfo tyi = p.tinfo[ti];
  const BdType bti = p.bt[ti];
  const int ni = tyi.nstates;
  const double pi0 = p.prob0[i], pi1 = 1.0 - pi0;
  const int jnum = active ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.na;
  const bool dens_i = ni > 1 && bti.use_density == 1;

  double fx = 0, fy = 0, fz = 0, eacc = 0, S0 = 0, S1 = 0, pf0 = 0, pf1 = 0;
  double vir[6] = {0, 0, 0, 0, 0, 0};
  // PA: this site's share of the tallies.  Symmetric visits (UCG-UCG, CG-CG): what the site accumulates anyway (we);
  // a UCG centre keeps half of its visit of a CG neighbor, a CG centre receives the other half of the visit its LOCAL
  // UCG neighbor makes (the reference skips the CG centre's own visit, :411-420)
  double ea = 0;
  double va[6] = {0, 0, 0, 0, 0, 0};
  double R0 = 1.0, R1 = 1.0;
  int E0 = 0, E1 = 0, nfac = 0;
  auto renorm = [](double &r, int &e) {
    const int hi = __double2hiint(r), ex = (hi >> 20) & 0x7ff;
    if (ex != 0 && ex != 0x7ff) {
      r = __hiloint2double((hi & (int)0x800fffff) | (1023 << 20), __double2loint(r));
      e += ex - 1023;
    }
  };
  double lpi0 = 0.0, lpi1 = 0.0;                      // log of this site's priors, taken at the first CG neighbor
  bool have_lpi = false;

  RowWalk<LPA> rw(row, sub, jnum);
  for (int jj = sub; jj < jnum; jj += LPA, rw.advance()) {
    const int jraw = rw.raw(jj);
    const double factor_lj = p.special_lj[(jraw >> UCG_SBBITS) & 3];
    const int j = jraw & UCG_NEIGHMASK;
    const double4 rj = gather_pos(p.gt, p.pos, j);
    const int tsj = gather_ts(p.gt, p.ts, j);
    const int tj = tsj & 0xffff;
    const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
    const double rsq = rsq_exact(dx, dy, dz);
    const PairInfo pi = prow[tj];
    if (rsq < pi.cutsq) {
      const int nj = pi.nj;
      const bool jlocal = j < p.nlocal;
      double u[4] = {0, 0, 0, 0}, f[4] = {0, 0, 0, 0};
      int ec = 0;
      if (W) ec = fast_table_eval<W>(s_tab, p.ft, rsq, u, f);
      else
        for (int a = 0; a < ni; a++)
          for (int b = 0; b < nj; b++) {
            int e1 = table_eval(p.tables[pi.tab[a * 2 + b]], rsq, u[a * 2 + b], f[a * 2 + b]);
            if (e1 && !ec) ec = e1;
          }
      if (ec) {
        report_error(p.err, ec, p.tag[i], p.tag[j], rsq);
        continue;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) { u[k] *= factor_lj; f[k] *= factor_lj; }
      double e, fpair, wf, we;   // wf: weight of d*fpair in f_i; we: weight in the E / virial tally
      double e_a = 0.0, wa = 0.0; // PA: energy of the visit this site has a share in, and the share
      if (ni == 2 && nj == 2) {
        const double pj0 = p.prob0[j], pj1 = 1.0 - pj0;
        const double J = u[3] + u[0] - u[1] - u[2];
        const double bij = exp(-J / p.kT), aij = bij - 1.0;
        const double Q = (pi1 + pj1) * aij + 1.0;
        const double D = fmax(Q * Q - 4.0 * aij * bij * pi1 * pj1, 0.0);
        double p11;
        if (fabs(aij) < 1.0e-6) p11 = pi1 * pj1;
        else if (Q < 0.0) p11 = (Q - sqrt(D)) / (2.0 * aij);
        else p11 = (2.0 * bij * pi1 * pj1) / (Q + sqrt(D));
        const double p00 = 1.0 + p11 - pi1 - pj1, p10 = pi1 - p11, p01 = pj1 - p11;
        e = p00 * u[0] + p01 * u[1] + p10 * u[2] + p11 * u[3];
        fpair = p00 * f[0] + p01 * f[1] + p10 * f[2] + p11 * f[3];
        wf = 1.0; we = 0.5;
        e_a = e; wa = we;
        const int sj = (tsj >> 16) & 1;           // (:596-598) only the neighbor's current state is tallied
        S0 += sj ? u[1] : u[0];
        S1 += sj ? u[3] : u[2];
        if (bti.use_density == 1) {
          if (p.logsum == 0) {
            pf0 -= (u[2] - u[0] + p.kT * log(p10 / p00));
            pf1 -= (u[3] - u[1] + p.kT * log(p11 / p01));
          } else if (p.logsum == 1) {
            pf0 -= u[2] - u[0];
            pf1 -= u[3] - u[1];
            R0 *= p10 / p00; renorm(R0, E0);
            R1 *= p11 / p01; renorm(R1, E1);
          } else {
            pf0 -= u[2] - u[0];
            pf1 -= u[3] - u[1];
            R0 *= p10 / p00;
            R1 *= p11 / p01;
            if (++nfac == 4) { pf0 -= p.kT * log(R0); pf1 -= p.kT * log(R1); R0 = R1 = 1.0; nfac = 0; }
          }
        }
      } else if (ni == 2) {                        // centre UCG, neighbor CG (:423-517)
        e = pi0 * u[0] + pi1 * u[2];
        fpair = pi0 * f[0] + pi1 * f[2];
        wf = 1.0; we = jlocal ? 1.0 : 0.5;
        e_a = e; wa = 0.5;                         // the other half belongs to the CG neighbor
        S0 += u[0];
        S1 += u[2];
        if (bti.use_density == 1) {
          if (!have_lpi) { lpi0 = log(pi0); lpi1 = log(pi1); have_lpi = true; }
          pf0 -= u[0] + p.kT * lpi0;
          pf1 -= u[2] + p.kT * lpi1;
        }
      } else if (nj == 2) {                        // centre CG, neighbor UCG: the reference skips this visit
        // (:411-420) and lets the neighbor's own visit scatter -d*fpair here, which with newton off
        // happens for LOCAL neighbors only
        const double pj0 = p.prob0[j], pj1 = 1.0 - pj0;
        e = 0.0;
        fpair = pj0 * f[0] + pj1 * f[1];
        wf = jlocal ? 1.0 : 0.0; we = 0.0;
        e_a = pj0 * u[0] + pj1 * u[1]; wa = jlocal ? 0.5 : 0.0;   // half of the visit the local UCG neighbor makes
      } else {                                     // CG-CG (:331-405): half per visit, reaction to local j only
        e = u[0];
        fpair = 0.5 * f[0];
        wf = jlocal ? 2.0 : 1.0; we = jlocal ? 1.0 : 0.5;
        e_a = e; wa = we;
      }
      eacc += we * e;
      const double ff = wf * fpair;
      fx += dx * ff; fy += dy * ff; fz += dz * ff;
      const double fv = we * fpair;
      vir[0] += dx * dx * fv; vir[1] += dy * dy * fv; vir[2] += dz * dz * fv;
      vir[3] += dx * dy * fv; vir[4] += dx * dz * fv; vir[5] += dy * dz * fv;
      if (PA) {
        ea += wa * e_a;
        const double fa = wa * fpair;
        va[0] += dx * dx * fa; va[1] += dy * dy * fa; va[2] += dz * dz * fa;
        va[3] += dx * dy * fa; va[4] += dx * dz * fa; va[5] += dy * dz * fa;
      }
    }
  }
  if (p.logsum && dens_i) {
    pf0 -= p.kT * (log(R0) + (double)E0 * 0.69314718055994530942);
    pf1 -= p.kT * (log(R1) + (double)E1 * 0.69314718055994530942);
  }
  fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
  eacc = group_sum<LPA>(eacc);
  if (PA) {
    ea = group_sum<LPA>(ea);
    if (active && sub == 0 && p.eatom) p.eatom[i] = ea;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double a = group_sum<LPA>(va[k]);
      if (active && sub == 0 && p.vatom) p.vatom[6 * (size_t)i + k] = a;
    }
  }
  S0 = group_sum<LPA>(S0); S1 = group_sum<LPA>(S1);
  pf0 = group_sum<LPA>(pf0); pf1 = group_sum<LPA>(pf1);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double v = group_sum<LPA>(vir[k]);
    if (active && sub == 0) evacc[1 + k] += v;
  }
  if (active && sub == 0) {
    double s0 = -S0 * p.inv_kT, s1 = -S1 * p.inv_kT, cvf = 0.0;
    if (dens_i) {
      // one-body terms (:302-317); jnum_f = 1 - numneigh[i] is the LIST length (as-is)
      const double jnum_f = 1.0 - (double)jnum;
      if (bti.entropy) {
        pf0 -= p.kT * log(pi0) * jnum_f;
        pf1 -= p.kT * log(pi1) * jnum_f;
      }
      pf0 -= tyi.mu0; pf1 -= tyi.mu1;
      s0 -= tyi.mu0 / p.kT; s1 -= tyi.mu1 / p.kT;
      const double pa = p.partial0[i];
      cvf = pf0 * pa + pf1 * (-pa);               // sum over si of prior_prob_force*prior_prob_partial (:700)
    }
    p.cvf[i] = cvf;
    p.frc[i] = make_double4(fx, fy, fz, 0.0);
    // posterior (:675-693, type index per repaired Q11); the style keeps its scores private and
    // publishes only ucgp — atom->ucgsoftmaxscores stay at their cleared value
    p.ucgp[i] = ni > 1 ? exp(s1) / (exp(s0) + exp(s1)) : 1.0;
    p.scores[i] = make_double2(0.0, 0.0);
    evacc[0] += eacc;
  }
  }   // persistent loop over site groups
  block_reduce_store<7, BS>(evacc, p.partials);
}

// posc[j] = {x, y, z, c_j} for the uniform back-force sweep (pair_common.cuh): c_j = cvf_j of an OWNED density site, else 0
// (with newton off the reference's loop of a ghost never runs, and only density sites carry a CV force)
__global__ void k_bd_posc(BdArgs p, int nall, double4 *__restrict__ posc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nall) return;
  double4 r = p.pos[j];
  const int tj = p.ts[j] & 0xffff;
  r.w = (j < p.nlocal && p.bt[tj].use_density == 1 && p.tinfo[tj].nstates > 1) ? p.cvf[j] : 0.0;
  posc[j] = r;
}

template <int LPA, int BS, bool PA>
__global__ void __launch_bounds__(BS) k_bd_back(BdArgs p) {
  const int gid = (blockIdx.x * BS + threadIdx.x) / LPA;
  const int sub = threadIdx.x % LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : 0;
  const double4 ri = p.pos[i];
  const int ti = p.ts[i] & 0xffff;
  const BdType bti = p.bt[ti];
  const bool dens_i = bti.use_density == 1 && p.tinfo[ti].nstates > 1;
  const double cvf_i = p.cvf[i];
  const int jnum = active ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.na;
  double fx = 0, fy = 0, fz = 0;
  double vir[6] = {0, 0, 0, 0, 0, 0};
  double va[6] = {0, 0, 0, 0, 0, 0};
  for (int jj = sub; jj < jnum; jj += LPA) {
    const int j = row[rowslot(jj)] & UCG_NEIGHMASK;
    const double4 rj = p.pos[j];
    const int tj = p.ts[j] & 0xffff;
    const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
    const double rsq = rsq_exact(dx, dy, dz);
    if (rsq < prow[tj].cutsq) {
      const double r = sqrt(rsq);
      const bool jlocal = j < p.nlocal;
      // g(r) of both sites: ONE tanh when they share the threshold radius (sites of one actual type: the usual case)
      const BdType btj = p.bt[tj];
      const bool dens_j = jlocal && btj.use_density == 1 && p.tinfo[tj].nstates > 1;
      const double rth_a = dens_i ? bti.r_th : btj.r_th;
      const double g_a = (dens_i || dens_j) ? bd_prox(r, rth_a) : 0.0;
      const double g_j = (dens_i && dens_j && btj.r_th != rth_a) ? bd_prox(r, btj.r_th) : g_a;
      const double own = dens_i ? cvf_i * g_a / r : 0.0;    // i's loop (:713-716)
      const double oth = dens_j ? p.cvf[j] * g_j / r : 0.0; // j's loop scatters to i
      const double fp = own + oth;
      fx += fp * dx; fy += fp * dy; fz += fp * dz;
      const double w = (jlocal ? 1.0 : 0.5) * own;   // ev_tally(i,j,nlocal,newton=0,0,0,fpair,...) in i's loop (:722)
      vir[0] += w * dx * dx; vir[1] += w * dy * dy; vir[2] += w * dz * dz;
      vir[3] += w * dx * dy; vir[4] += w * dx * dz; vir[5] += w * dy * dz;
      if (PA) {   // half of this site's visit, half of the visit of a local partner (oth is zero for a ghost)
        const double wa = 0.5 * (own + oth);
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

}  // namespace

extern "C" int ucgb200_pair_bethe_density_configure(ucgb200_ctx *c, int n_actual, const int *use_density,
                                                    const int *use_entropy, const double *cv_threshold,
                                                    const double *threshold_radius) {
  if (!c || n_actual < 1 || !use_density || !use_entropy || !cv_threshold || !threshold_radius) return -1;
  auto &b = c->bdens;
  b.n_actual = n_actual;
  b.use_density.assign(use_density, use_density + n_actual + 1);
  b.use_entropy.assign(use_entropy, use_entropy + n_actual + 1);
  b.cv_th.assign(cv_threshold, cv_threshold + n_actual + 1);
  b.r_th.assign(threshold_radius, threshold_radius + n_actual + 1);
  for (int t = 1; t <= n_actual; t++)
    if (b.use_density[t] && !(b.cv_th[t] > 0.0 && b.r_th[t] > 0.0))
      return fail(c, "pair_bethe_density: density threshold and radius must be positive");
  b.set = true;
  b.dirty = true;
  return 0;
}

extern "C" int ucgb200_pair_bethe_density(ucgb200_ctx *c, int eflag, int vflag) {
  if (!c) return -1;
  auto &b = c->bdens;
  if (!b.set) return fail(c, "pair_bethe_density: not configured");
  if (c->dens.set) return fail(c, "pair_bethe_density: context is configured for table_rleucg_interface");
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (b.n_actual != c->n_actual) return fail(c, "pair_bethe_density: configured for a different number of actual types");
  if (!c->list_valid) return fail(c, "pair_bethe_density: neighbor list not built");
  c->ev_valid = false;
  if (c->nlocal == 0) {   // an empty brick still reports (zero) energy and virial
    UCG_CHECK(c, cudaMemsetAsync(c->d_ev.p, 0, 32 * sizeof(double), c->stream));
    c->ev_valid = true;
    c->ev_two_parts = false;
    return 0;
  }
  // energy and virial are always evaluated; bits 2 / 4 ask for the per-atom tallies (ucgb200_pair_peratom)
  const bool want_eatom = (eflag & 2) != 0, want_vatom = (vflag & 4) != 0;
  c->eatom_valid = c->vatom_valid = false;
  if (b.dirty) {
    std::vector<BdType> bt(b.n_actual + 1);
    for (int t = 0; t <= b.n_actual; t++) {
      bt[t].use_density = b.use_density[t]; bt[t].entropy = b.use_entropy[t];
      bt[t].cv_th = b.cv_th[t]; bt[t].r_th = b.r_th[t];
    }
    UCG_CHECK(c, b.d_bt.ensure(bt.size() * sizeof(BdType)));
    UCG_CHECK(c, cudaMemcpy(b.d_bt.p, bt.data(), bt.size() * sizeof(BdType), cudaMemcpyHostToDevice));
    b.dirty = false;
  }
  const int nall = c->nlocal + c->nghost;
  UCG_CHECK(c, b.d_prob.ensure(nall));
  UCG_CHECK(c, b.d_partial.ensure(c->nlocal));
  UCG_CHECK(c, b.d_cvf.ensure(c->nlocal));
  constexpr int LPA = 8, BS = 256;
  const int nblk = nblocks((long long)c->nlocal * LPA, BS);
  UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 8 + 64));
  BdArgs a{};
  a.pos = c->pos.p; a.ts = c->ts.p; a.tag = c->tag.p; a.nlocal = c->nlocal;
  a.neigh = c->neigh.p; a.stride = c->neigh_stride; a.numneigh = c->numneigh.p;
  a.pinfo = c->d_pairinfo.p; a.tinfo = c->d_typeinfo.p; a.bt = (const BdType *)b.d_bt.p; a.na = c->n_actual + 1;
  a.tables = c->d_tables.p;
  for (int k = 0; k < 4; k++) a.special_lj[k] = c->special_lj[k];
  a.kT = c->kT; a.inv_kT = 1.0 / c->kT;
  // sum_j kT log(p10/p00) and sum_j kT log(p11/p01) of the probability force (:632-641) as ONE logarithm per lane: the ratios
  // are multiplied up, the binary exponent moves into an integer after every factor (no overflow / underflow), and
  // log(mantissa) + exponent ln 2 is subtracted after the loop.  The two logarithms per UCG-UCG visit were 39 % of the
  // pair sweep's instructions (profiles/r02_k_bd_pair_lines.json); the sums differ from the per-visit ones by a few ulp
  // (scripts/dbg_bd_logsum.py).  UCGB200_BD_LOGSUM=0: one logarithm per ratio, as the reference writes it.
  a.logsum = getenv("UCGB200_BD_LOGSUM") ? atoi(getenv("UCGB200_BD_LOGSUM")) : 1;
  a.prob0 = b.d_prob.p; a.partial0 = b.d_partial.p; a.cvf = b.d_cvf.p;
  a.frc = c->frc.p; a.scores = c->scores.p; a.ucgp = c->ucgp.p; a.partials = c->d_partials.p; a.err = c->d_err.p;
  if (want_eatom) { UCG_CHECK(c, c->d_eatom.ensure((size_t)c->nlocal + 8)); a.eatom = c->d_eatom.p; }
  if (want_vatom) { UCG_CHECK(c, c->d_vatom.ensure(6 * (size_t)c->nlocal + 8)); a.vatom = c->d_vatom.p; }
  const bool pa = want_eatom || want_vatom;
  const auto &h = c->halo;
  if (c->timers_on) cudaEventRecord(c->ev_pair0, c->stream);
  k_bd_prior<LPA, BS><<<nblk, BS, 0, c->stream>>>(a);
  UCG_LAUNCHED(c);
  if (h.nlimg) {
    k_bd_ghost<<<nblocks(h.nlimg, 256), 256, 0, c->stream>>>(a.prob0, c->nlocal, h.nlimg, c->img_owner.p + h.nsend, c->slot_of_src.p);
    UCG_LAUNCHED(c);
  }
  if ((rc = ucg_mb_forward_scalars(c, a.prob0, nullptr, nullptr))) return rc;   // ghosts owned by other bricks
  int nblk_pair = nblk;
  const size_t tab_bytes = (size_t)c->fast_len * c->fast_ntab * sizeof(double2);
  if (!pa && c->fast_uniform && tab_bytes <= 220 * 1024 && !(getenv("UCGB200_FORCE_GENERAL") && atoi(getenv("UCGB200_FORCE_GENERAL")))) {
    // one 2-state type, LINEAR tables on one grid: interleaved rows in shared memory, persistent CTAs
    constexpr int FLPA = 4, FBS = 512;
    // a.gt stays 0: texture-pipe gathers were measured here and gave nothing (these sweeps are bound by FP64
    // transcendentals, not by the LSU data pipe); UCGB200_TEX_ALL=1 turns them on for experiments
    if (getenv("UCGB200_TEX_ALL") && atoi(getenv("UCGB200_TEX_ALL")) && (rc = ucg_bind_gather_textures(c, &a.gt.pos, &a.gt.ts))) return rc;
    const ucg::TableDev &t0 = c->tables[c->fast_tab[0]];
    a.ft.table = c->d_fast_table.p; a.ft.tablen = c->fast_len; a.ft.W = c->fast_ntab;
    a.ft.innersq = t0.innersq; a.ft.delta = t0.delta; a.ft.invdelta = t0.invdelta;
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, c->device);
    nblk_pair = std::min(dev_sms, nblocks((long long)c->nlocal * FLPA, FBS));
    UCG_CHECK(c, c->d_partials.ensure((size_t)std::max(nblk, nblk_pair) * 8 + 64));
    a.partials = c->d_partials.p;
    if (c->fast_ntab == 3) {
      auto kern = k_bd_pair<FLPA, FBS, 3, false>;
      UCG_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      kern<<<nblk_pair, FBS, tab_bytes, c->stream>>>(a);
    } else {
      auto kern = k_bd_pair<FLPA, FBS, 4, false>;
      UCG_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      kern<<<nblk_pair, FBS, tab_bytes, c->stream>>>(a);
    }
  } else if (pa) {
    // per-atom tallies are asked for on output steps only: the general kernel carries them
    k_bd_pair<LPA, BS, 0, true><<<nblk, BS, 0, c->stream>>>(a);
  } else {
    k_bd_pair<LPA, BS, 0, false><<<nblk, BS, 0, c->stream>>>(a);
  }
  UCG_LAUNCHED(c);
  if ((rc = reduce_partials(c, nblk_pair, 7, 0))) return rc;
  // uniform case (every density type has the same threshold radius, every pair the same cutoff — one 2-state type is):
  // the back-force sweep gathers one packed record per neighbor (k_cv_back_fast, pair_common.cuh); UCGB200_CV_FAST=0 and
  // per-atom virial requests take the general sweep (same bits)
  double urth = -1.0, ucut = -1.0;
  bool uniform = !want_vatom && !(getenv("UCGB200_CV_FAST") && atoi(getenv("UCGB200_CV_FAST")) == 0);
  for (int t = 1; t <= b.n_actual && uniform; t++)
    if (b.use_density[t] == 1) { if (urth < 0) urth = b.r_th[t]; else if (b.r_th[t] != urth) uniform = false; }
  for (int i = 1; i < a.na && uniform; i++)
    for (int j = 1; j < a.na && uniform; j++) {
      const double cs = c->h_pairinfo[i * a.na + j].cutsq;
      if (ucut < 0) ucut = cs; else if (cs != ucut) uniform = false;
    }
  if (uniform && urth > 0 && ucut > 0) {
    UCG_CHECK(c, c->posc.ensure((size_t)nall + 8));
    k_bd_posc<<<nblocks(nall, 256), 256, 0, c->stream>>>(a, nall, c->posc.p);
    UCG_LAUNCHED(c);
    CvBackArgs ba{c->posc.p, c->nlocal, a.neigh, a.stride, a.numneigh, ucut, urth, a.frc, a.partials};
    k_cv_back_fast<LPA, BS, 0><<<nblk, BS, 0, c->stream>>>(ba);
  } else if (want_vatom) k_bd_back<LPA, BS, true><<<nblk, BS, 0, c->stream>>>(a);
  else k_bd_back<LPA, BS, false><<<nblk, BS, 0, c->stream>>>(a);
  UCG_LAUNCHED(c);
  if ((rc = reduce_partials(c, nblk, 7, 16))) return rc;   // second virial part -> d_ev[16..22]
  if (c->timers_on) { cudaEventRecord(c->ev_pair1, c->stream); c->pair_timed = true; }
  c->ev_valid = true;
  c->ev_two_parts = true;
  c->eatom_valid = want_eatom;
  c->vatom_valid = want_vatom;
  return 0;
}

// prior probability of substate 0 and the CV force of the last evaluation (diagnostics / tests), host order
extern "C" int ucgb200_pair_bethe_density_priors(ucgb200_ctx *c, int cap, double *prob0, double *cvforce) {
  if (!c || cap < c->nlocal) return -1;
  cudaSetDevice(c->device);
  auto &b = c->bdens;
  const int n = c->nlocal;
  if (!b.set || b.d_prob.cap < (size_t)n) return fail(c, "pair_bethe_density: nothing evaluated yet");
  std::vector<double> tmp(n);
  std::vector<int> orig(n);
  UCG_CHECK(c, cudaMemcpyAsync(orig.data(), c->orig.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  for (int pass = 0; pass < 2; pass++) {
    double *dst = pass == 0 ? prob0 : cvforce;
    if (!dst) continue;
    UCG_CHECK(c, cudaMemcpyAsync(tmp.data(), pass == 0 ? b.d_prob.p : b.d_cvf.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    UCG_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int s = 0; s < n; s++) dst[orig[s]] = tmp[s];
  }
  return 0;
}
