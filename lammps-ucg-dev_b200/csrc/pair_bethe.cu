// pair_bethe.cu — PairTable_UCG_Bethe::compute (UCG/pair_table_ucg_bethe.cpp:88-630) for sm_100a.
//
// Same schedule as pair_ucgld.cu (LPA lanes per owned site over its FULL row, centre-site
// accumulation, no atomics, no reverse halo).  Per UCG-UCG pair (:544-604):
//   J = u11+u00-u01-u10 (clamped at -700 kT), b = exp(-J/kT), a = expm1(-J/kT)
//   Q = (pi1+pj1) a + 1,  D = max(Q^2 - 4ab pi1 pj1, 0)
//   p11 = pi1 pj1                       (method mf, or |a| < 1e-6)
//       = (Q - sqrt D)/(2a)             (Q < 0)
//       = 2 b pi1 pj1 / (Q + sqrt D)    (otherwise)
//   p00 = 1+p11-pi1-pj1, p10 = pi1-p11, p01 = pj1-p11
//   E = sum p_ab u_ab, fpair = sum p_ab f_ab
// Scores: pseudo-likelihood tally s_i[si] -= u[si][state_j]/kT (:526-539) when the reference's
// pseudo_flag is 0 ("pseudo yes"), SCE conditionals (:583-601) when it is 1.
//
// Priors (decision for quirks Q6/Q7): the reference takes the prior of the row
// owner i from ucgl[i] and that of the neighbor j from ucgp[j]; which site of a pair is "i"
// depends on the half-list order.  The device uses ONE rule for every site k:
//   ucgp[k] < -0.999 (never evaluated) -> by prior_flag: softmax(-mu/kT) of k's own type, that
//                                          plus noise, or (1-ucgl[k], ucgl[k]);
//   otherwise                          -> (1-ucgl[k], ucgl[k]).
// This equals the reference whenever ucgl == ucgp (fix ucgstate without `ld`, the way the
// style is meant to be used) and on the first evaluation of single-type systems.
// The SCE score formulas of the reference are not symmetric under i<->j (they depend on which
// site owns the row); the device applies the row-owner ("i") formula to the centre site of
// every visit — energies and forces are unaffected.
#include "pair_common.cuh"

#include <algorithm>

using namespace ucg;

namespace {

struct BetheArgs {
  const double4 *pos;
  const int *ts;
  const int *tag;
  const double *ucgp;
  int nlocal;
  const int *neigh;
  int stride;
  const int *numneigh;
  const PairInfo *pinfo;
  const TypeInfo *tinfo;
  int na;
  const TableDev *tables;
  double special_lj[4];
  double kT, inv_kT;
  int method, pseudo, prior;
  double noise;
  unsigned seed;
  double4 *frc;
  double2 *scores;
  double *partials;
  ErrWord *err;
  double *eatom, *vatom;   // per-atom tallies (EV launches only), nullptr: not asked for
  FastTable ft;   // shared-memory table path (W > 0)
  GatherTex gt;   // neighbor gathers through the texture pipe (0 = plain loads)
};

// prior probability of substate 1 of site k
__device__ __forceinline__ double prior1(const BetheArgs &p, const TypeInfo &ty, double ucgl, double ucgp, int tag) {
  if (ty.nstates == 1) return 0.0;
  if (ucgp < -0.999) {
    if (p.prior == 2) return ucgl;
    const double e0 = exp(-ty.mu0 / p.kT), e1 = exp(-ty.mu1 / p.kT);
    double p0 = e0 / (e0 + e1);
    if (p.prior == 1) {
      double r = (philox_uniform(p.seed, 0x42455448u /* "BETH" */, (unsigned)tag, 0ull) - 0.5) * 2.0 * p.noise;
      p0 = fmin(0.999999, fmax(p0 + r, 0.0));
    }
    return 1.0 - p0;
  }
  return ucgl;
}

// W = 0: tables through L1 (any type system / table style); W = 3 / 4: interleaved LINEAR tables of the
// single 2-state type in shared memory, persistent CTAs (one per SM)
template <int LPA, bool EV, int BS, int W>
__global__ void __launch_bounds__(BS) k_pair_bethe(BetheArgs p) {
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
  const int tsi = p.ts[i];
  const int ti = tsi & 0xffff;
  const TypeInfo tyi = p.tinfo[ti];
  const int ni = tyi.nstates;
  const double pi1 = prior1(p, tyi, ri.w, p.ucgp[i], p.tag[i]);
  const double pi0 = 1.0 - pi1;
  const int jnum = active ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.na;

  double fx = 0, fy = 0, fz = 0, eacc = 0, S0 = 0, S1 = 0;
  double vir[6] = {0, 0, 0, 0, 0, 0};

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
      double e, fpair;
      if (ni == 2 && nj == 2) {
        const TypeInfo tyj = p.tinfo[tj];
        const double pj1 = prior1(p, tyj, rj.w, p.ucgp[j], p.tag[j]);
        const double pj0 = 1.0 - pj1;
        double J = u[3] + u[0] - u[1] - u[2];
        if (J * p.inv_kT < -709.0) J = -700.0 * p.kT;
        // the reference takes exp AND expm1 of the same argument (:550-551); one evaluation serves both: expm1(x) + 1
        // is exp(x) to 1.1e-16 ABSOLUTE, and b enters p11 and D only through products with probabilities <= 1
        const double aij = expm1(-J * p.inv_kT), bij = aij + 1.0;
        const double Q = (pi1 + pj1) * aij + 1.0;
        const double D = fmax(Q * Q - 4.0 * aij * bij * pi1 * pj1, 0.0);
        double p11;
        if (p.method == 1 && fabs(aij) >= 1.0e-6) {
          if (Q < 0.0) p11 = (Q - sqrt(D)) / (2.0 * aij);
          else p11 = (2.0 * bij * pi1 * pj1) / (Q + sqrt(D));
        } else
          p11 = pi1 * pj1;
        const double p00 = 1.0 + p11 - pi1 - pj1, p10 = pi1 - p11, p01 = pj1 - p11;
        e = p00 * u[0] + p01 * u[1] + p10 * u[2] + p11 * u[3];
        fpair = p00 * f[0] + p01 * f[1] + p10 * f[2] + p11 * f[3];
        if (p.pseudo == 0) {
          const int sj = (tsj >> 16) & 1;
          S0 += sj ? u[1] : u[0];
          S1 += sj ? u[3] : u[2];
        } else {
          // row-owner formulas of :584-596, literally: pj0i0=p00/pi0, pj1i0=p10/pi1, pj0i1=p01/pi0, pj1i1=p11/pi1
          S0 += (p00 / pi0) * u[0] + (p10 / pi1) * u[1];
          S1 += (p01 / pi0) * u[2] + (p11 / pi1) * u[3];
        }
      } else if (ni == 2) {  // centre UCG, neighbor CG (:389-460)
        e = pi0 * u[0] + pi1 * u[2];
        fpair = pi0 * f[0] + pi1 * f[2];
        S0 += u[0];
        S1 += u[2];
      } else if (nj == 2) {  // centre CG, neighbor UCG (:310-385)
        const TypeInfo tyj = p.tinfo[tj];
        const double pj1 = prior1(p, tyj, rj.w, p.ucgp[j], p.tag[j]);
        e = (1.0 - pj1) * u[0] + pj1 * u[1];
        fpair = (1.0 - pj1) * f[0] + pj1 * f[1];
      } else {
        e = u[0];
        fpair = f[0];
      }
      eacc += e;
      fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
      if (EV) {
        vir[0] += dx * dx * fpair; vir[1] += dy * dy * fpair; vir[2] += dz * dz * fpair;
        vir[3] += dx * dy * fpair; vir[4] += dx * dz * fpair; vir[5] += dy * dz * fpair;
      }
    }
  }
  fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
  eacc = group_sum<LPA>(eacc);
  S0 = group_sum<LPA>(S0); S1 = group_sum<LPA>(S1);
  if (active && sub == 0) {
    p.frc[i] = make_double4(fx, fy, fz, 0.0);   // the Bethe style leaves ucgforce at its cleared value
    if (ni == 2)  // :162-170 initialise with -mu/kT, then the pair tallies
      p.scores[i] = make_double2(-tyi.mu0 * p.inv_kT - S0 * p.inv_kT, -tyi.mu1 * p.inv_kT - S1 * p.inv_kT);
    else
      p.scores[i] = make_double2(-tyi.mu0 * p.inv_kT, 0.0);
    if (EV) evacc[0] += 0.5 * eacc;
    if (EV && p.eatom) p.eatom[i] = 0.5 * eacc;          // this site's half of every pair it is in (ev_tally, eflag_atom)
  }
  if (EV) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double v = group_sum<LPA>(vir[k]);
      if (active && sub == 0) { evacc[1 + k] += 0.5 * v; if (p.vatom) p.vatom[6 * (size_t)i + k] = 0.5 * v; }
    }
  }
  }   // persistent loop over site groups
  if (EV) block_reduce_store<7, BS>(evacc, p.partials);
}

}  // namespace

extern "C" int ucgb200_pair_bethe(ucgb200_ctx *c, int eflag, int vflag, int method, int pseudo, int prior,
                                  double noise_level, int seed) {
  if (!c) return -1;
  if (method < 0 || method > 1 || pseudo < 0 || pseudo > 1 || prior < 0 || prior > 2)
    return fail(c, "pair_bethe: unknown method/pseudo/prior");
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (!c->list_valid) return fail(c, "pair_bethe: neighbor list not built");
  c->ev_valid = false;
  c->ev_two_parts = false;
  if (c->nlocal == 0) {   // an empty brick still reports (zero) energy and virial
    UCG_CHECK(c, cudaMemsetAsync(c->d_ev.p, 0, 32 * sizeof(double), c->stream));
    c->ev_valid = true;
    c->ev_two_parts = false;
    return 0;
  }
  const bool ev = eflag || vflag;
  // LAMMPS flag bits: eflag & 2 = ENERGY_ATOM, vflag & 4 = VIRIAL_ATOM (ucgb200_pair_peratom fetches them)
  const bool want_eatom = (eflag & 2) != 0, want_vatom = (vflag & 4) != 0;
  c->eatom_valid = c->vatom_valid = false;
  constexpr int LPA = 8, BS = 256;
  BetheArgs a{};
  if (want_eatom) { UCG_CHECK(c, c->d_eatom.ensure((size_t)c->nlocal + 8)); a.eatom = c->d_eatom.p; }
  if (want_vatom) { UCG_CHECK(c, c->d_vatom.ensure(6 * (size_t)c->nlocal + 8)); a.vatom = c->d_vatom.p; }
  a.pos = c->pos.p; a.ts = c->ts.p; a.tag = c->tag.p; a.ucgp = c->ucgp.p; a.nlocal = c->nlocal;
  a.neigh = c->neigh.p; a.stride = c->neigh_stride; a.numneigh = c->numneigh.p;
  a.pinfo = c->d_pairinfo.p; a.tinfo = c->d_typeinfo.p; a.na = c->n_actual + 1; a.tables = c->d_tables.p;
  for (int k = 0; k < 4; k++) a.special_lj[k] = c->special_lj[k];
  a.kT = c->kT; a.inv_kT = 1.0 / c->kT;
  a.method = method; a.pseudo = pseudo; a.prior = prior; a.noise = noise_level; a.seed = (unsigned)seed;
  a.frc = c->frc.p; a.scores = c->scores.p; a.err = c->d_err.p;
  int nblk = nblocks((long long)c->nlocal * LPA, BS);
  const size_t tab_bytes = (size_t)c->fast_len * c->fast_ntab * sizeof(double2);
  const bool fast = c->fast_uniform && tab_bytes <= 220 * 1024 &&
                    !(getenv("UCGB200_FORCE_GENERAL") && atoi(getenv("UCGB200_FORCE_GENERAL")));
  constexpr int FLPA = 4, FBS = 512;
  if (fast) {
    // a.gt stays 0: texture-pipe gathers were measured here and gave nothing (these sweeps are bound by FP64
    // transcendentals, not by the LSU data pipe); UCGB200_TEX_ALL=1 turns them on for experiments
    if (getenv("UCGB200_TEX_ALL") && atoi(getenv("UCGB200_TEX_ALL")) && (rc = ucg_bind_gather_textures(c, &a.gt.pos, &a.gt.ts))) return rc;
    const ucg::TableDev &t0 = c->tables[c->fast_tab[0]];
    a.ft.table = c->d_fast_table.p; a.ft.tablen = c->fast_len; a.ft.W = c->fast_ntab;
    a.ft.innersq = t0.innersq; a.ft.delta = t0.delta; a.ft.invdelta = t0.invdelta;
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, c->device);
    nblk = std::min(dev_sms, nblocks((long long)c->nlocal * FLPA, FBS));
  }
  if (ev) UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 8 + 64));
  a.partials = c->d_partials.p;
  if (c->timers_on) cudaEventRecord(c->ev_pair0, c->stream);
  if (fast) {
#define UCG_BETHE_FAST(EVV, WW)                                                                              \
    do {                                                                                                    \
      auto kern = k_pair_bethe<FLPA, EVV, FBS, WW>;                                                         \
      UCG_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes)); \
      kern<<<nblk, FBS, tab_bytes, c->stream>>>(a);                                                         \
    } while (0)
    if (c->fast_ntab == 3) { if (ev) UCG_BETHE_FAST(true, 3); else UCG_BETHE_FAST(false, 3); }
    else { if (ev) UCG_BETHE_FAST(true, 4); else UCG_BETHE_FAST(false, 4); }
#undef UCG_BETHE_FAST
  } else if (ev) k_pair_bethe<LPA, true, BS, 0><<<nblk, BS, 0, c->stream>>>(a);
  else k_pair_bethe<LPA, false, BS, 0><<<nblk, BS, 0, c->stream>>>(a);
  UCG_LAUNCHED(c);
  if (c->timers_on) { cudaEventRecord(c->ev_pair1, c->stream); c->pair_timed = true; }
  if (ev) {
    if ((rc = reduce_partials(c, nblk, 7, 0))) return rc;
    c->ev_valid = true;
    c->eatom_valid = want_eatom;
    c->vatom_valid = want_vatom;
  }
  return 0;
}
