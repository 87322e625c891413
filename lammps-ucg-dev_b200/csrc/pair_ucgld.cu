// pair_ucgld.cu — PairTable_UCGLD::compute (UCG/pair_table_ucgld.cpp:111-541) for sm_100a.
//
// Schedule: a group of LPA lanes per owned site walks that site's FULL neighbor row
// (coalesced row read, 32-byte gathers of {x,y,z,lambda} + 4-byte {type,state}); every
// contribution is accumulated into the CENTRE site only and reduced with warp shuffles,
// so f / ucgforce / ucgsoftmaxscores are written exactly once per site (force_clear,
// atom_vec_ucg.cpp:131-135, and the chemical-potential pre-add, :170-180, are folded in)
// and there are neither FP64 atomics nor a reverse halo.  A pair (i,j) is therefore
// evaluated from both ends; by the symmetry of tabindex (init_one, :892) both ends see the
// same four potentials, so the result equals the reference's half-list/newton-on tally.
//
// Per visited pair inside the cutoff (scenario 4, :424-519), with a=1-li, b=1-lj:
//   A  = b*u00 + lj*u01      B  = b*u10 + lj*u11          (u_ss' from 4 table look-ups)
//   FA = b*f00 + lj*f01      FB = b*f10 + lj*f11
//   E_ij   = a*A + li*B      fpair = a*FA + li*FB          (:507, :509)
//   ucgf_i -= B - A          == lj(u11-u01) + (1-lj)(u10-u00)   (:514)
//   s_i[si] -= u[si][state_j]/kT                           (:492-498)
// Scenarios 1-3 (:219-421) are the same formulas with one-state sites having weight
// {1,0}; scenario 2 uses the intended `sj` keying.
#include "pair_common.cuh"

using namespace ucg;

namespace {

struct PairArgs {
  const double4 *pos;
  const int *ts;
  const int *tag;
  int nlocal;
  const int *neigh;
  int stride;
  const int *numneigh;
  const PairInfo *pinfo;
  const TypeInfo *tinfo;
  int na;
  const TableDev *tables;
  double special_lj[4];
  double inv_kT;
  double4 *frc;
  double2 *scores;
  double *partials;
  ErrWord *err;
  double *eatom, *vatom;   // per-atom tallies (EV launches only), nullptr: not asked for
};

// ---------------------------------------------------------------- general kernel
template <int LPA, bool EV, int BS>
__global__ void __launch_bounds__(BS) k_pair_ucgld(PairArgs p) {
  const int gid = (blockIdx.x * BS + threadIdx.x) / LPA;
  const int sub = threadIdx.x % LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : 0;
  const double4 ri = p.pos[i];
  const int tsi = p.ts[i];
  const int ti = tsi & 0xffff;
  const TypeInfo tyi = p.tinfo[ti];
  const int ni = tyi.nstates;
  const double li = ri.w, ai = 1.0 - li;
  const int jnum = active ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  const PairInfo *prow = p.pinfo + ti * p.na;

  double fx = 0, fy = 0, fz = 0, accA = 0, accB = 0, S0 = 0, S1 = 0;
  double vir[6] = {0, 0, 0, 0, 0, 0};

  for (int jj = sub; jj < jnum; jj += LPA) {
    int jraw = row[rowslot(jj)];
    const double factor_lj = p.special_lj[(jraw >> UCG_SBBITS) & 3];
    const int j = jraw & UCG_NEIGHMASK;
    const double4 rj = p.pos[j];
    const int tsj = p.ts[j];
    const int tj = tsj & 0xffff;
    const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
    const double rsq = rsq_exact(dx, dy, dz);
    const PairInfo pi = prow[tj];
    if (rsq < pi.cutsq) {
      const int nj = pi.nj;
      const int sj = (nj == 2) ? ((tsj >> 16) & 1) : 0;
      const double lj = rj.w;
      const double wj0 = (nj == 2) ? 1.0 - lj : 1.0, wj1 = (nj == 2) ? lj : 0.0;
      double u[4] = {0, 0, 0, 0}, f[4] = {0, 0, 0, 0};
      int ec = 0;
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
      const double A = wj0 * u[0] + wj1 * u[1];
      const double FA = wj0 * f[0] + wj1 * f[1];
      double fpair;
      if (ni == 2) {
        const double B = wj0 * u[2] + wj1 * u[3];
        const double FB = wj0 * f[2] + wj1 * f[3];
        accA += A; accB += B;
        fpair = ai * FA + li * FB;
        S0 += sj ? u[1] : u[0];
        S1 += sj ? u[3] : u[2];
      } else {
        accA += A;
        fpair = FA;
      }
      fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
      if (EV) {
        vir[0] += dx * dx * fpair; vir[1] += dy * dy * fpair; vir[2] += dz * dz * fpair;
        vir[3] += dx * dy * fpair; vir[4] += dx * dz * fpair; vir[5] += dy * dz * fpair;
      }
    }
  }
  fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
  accA = group_sum<LPA>(accA); accB = group_sum<LPA>(accB);
  S0 = group_sum<LPA>(S0); S1 = group_sum<LPA>(S1);
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  if (active && sub == 0) {
    double4 fo;
    fo.x = fx; fo.y = fy; fo.z = fz;
    double2 so;
    double e_i;
    if (ni == 2) {
      fo.w = -tyi.dmu - (accB - accA);               // :177, :514
      so.x = -S0 * p.inv_kT;
      so.y = -tyi.dmu * p.inv_kT - S1 * p.inv_kT;    // :178, :497
      e_i = ai * accA + li * accB;
    } else {
      fo.w = 0.0; so.x = 0.0; so.y = 0.0;
      e_i = accA;
    }
    p.frc[i] = fo;
    p.scores[i] = so;
    if (EV) ev[0] = 0.5 * e_i;
    if (EV && p.eatom) p.eatom[i] = 0.5 * e_i;          // this site's half of every pair it is in (ev_tally, eflag_atom)
  }
  if (EV) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double v = group_sum<LPA>(vir[k]);
      if (active && sub == 0) { ev[1 + k] = 0.5 * v; if (p.vatom) p.vatom[6 * (size_t)i + k] = 0.5 * v; }
    }
    block_reduce_store<7, BS>(ev, p.partials);
  }
}

// 256-bit gather of one {x,y,z,lambda} record (LDG.E.256, sm_100+): one L1 request per
// neighbor instead of two 128-bit ones.
__device__ __forceinline__ double4 ld256(const double4 *p) {
  double4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}

// ucgstate bits of all owned+ghost sites: the per-neighbor state gather then touches one
// 128-byte line per 4096 sites instead of one per 32.
__global__ void k_pack_statebits(const int *__restrict__ ts, int nall, unsigned *__restrict__ bits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool s = i < nall && ((ts[i] >> 16) & 1);
  unsigned m = __ballot_sync(0xffffffffu, s);
  if ((threadIdx.x & 31) == 0 && i < nall) bits[i >> 5] = m;
}

// ------------------------------------------------------------------- fast kernel
// One 2-state actual type, LINEAR tables on a common rsq grid (the benchmark liquids).
// The W (3 or 4) unique tables are interleaved row-wise and staged once per CTA in shared
// memory: row it = {e00,f00, e01,f01, [e10,f10,] e11,f11}; a pair reads rows it and it+1
// (2*W 16-byte words, contiguous).  CTAs are persistent (one per SM).
struct FastArgs {
  const double4 *pos;
  const int *ts;
  const unsigned *sbits;  // ucgstate of every owned+ghost site, 1 bit each
  const int *tag;
  int nlocal;
  const int *neigh;
  int stride;
  const int *numneigh;
  const double2 *table;  // [tablen][W]
  int tablen;
  double innersq, delta, invdelta, cutsq;
  double dmu, inv_kT;
  double4 *frc;
  double2 *scores;
  double *partials;
  ErrWord *err;
  int smem_table;  // 1: stage table in shared memory, 0: read it through L1/L2
  // exact skipping of skin entries that cannot have entered the cutoff yet (neighbor.cu): level counts of
  // every row, the largest squared displacement since the build, 8 / skin; levcnt == nullptr: visit everything
  const uint4 *levcnt;
  const unsigned long long *maxdisp;
  double inv_w;
  cudaTextureObject_t postex;   // {x,y,z,lambda} records as 2 int4 texels each (TEX = true)
  cudaTextureObject_t sbtex;    // state bits as 32-bit texels (TEX = 3)
  // multi-brick runs: walk only one part of the owned sites (site_list = interior sites, then boundary sites;
  // part_count[0] = number of interior sites); nullptr: every site
  const int *site_list;
  const int *part_count;
  int part;
  double *eatom, *vatom;        // per-atom tallies (EV launches only), nullptr: not asked for
};

// PF = 1 software-pipelines the neighbor gathers: index and {x,y,z,lambda}/state of the
// next neighbor are in flight while the current pair is evaluated.
// SM = true: table rows come from shared memory (LDS.128, not generic loads).
// gather of one site record through the texture pipe (two 16-byte texel fetches) instead of the LSU pipe
__device__ __forceinline__ double4 ldtex(cudaTextureObject_t t, int j) {
  const int4 a = tex1Dfetch<int4>(t, 2 * j), b = tex1Dfetch<int4>(t, 2 * j + 1);
  return make_double4(__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z));
}
// split: {x,y} through the texture pipe, {z,lambda} through the LSU pipe
__device__ __forceinline__ double4 ldtex_half(cudaTextureObject_t t, const double4 *pos, int j) {
  const int4 a = tex1Dfetch<int4>(t, 2 * j);
  double2 zl;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(zl.x), "=d"(zl.y) : "l"(reinterpret_cast<const double2 *>(pos + j) + 1));
  return make_double4(__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), zl.x, zl.y);
}

template <int LPA, bool EV, int W, int BS, int PF, bool SM, int TEX = 0>
__global__ void __launch_bounds__(BS) k_pair_ucgld_fast(FastArgs p) {
  extern __shared__ double2 s_tab[];
  if (SM) {
    const int nwords = p.tablen * W;
    for (int k = threadIdx.x; k < nwords; k += BS) s_tab[k] = p.table[k];
    __syncthreads();
  }
  const int sub = threadIdx.x % LPA;
  const int groups_per_block = BS / LPA;
  const int tlm1 = p.tablen - 1;
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  // displacement level: a skin pair that sat (r0 - cut) beyond the cutoff at build time can only be inside it
  // now if r0 - cut < |d_i| + |d_j| <= 2 sqrt(maxdisp2); rows hold their skin entries in ascending r0 and
  // levcnt[i] the number of leading entries with (r0 - cut) below each eighth of the skin
  int level = 7;
  if (p.levcnt) {
    const double md = sqrt(__longlong_as_double((long long)*p.maxdisp)) * (1.0 + 1e-12);
    level = min(7, (int)floor(2.0 * md * p.inv_w));
  }

  int first = 0, count = p.nlocal;
  if (p.site_list) {
    const int ninterior = p.part_count[0];
    first = p.part == 0 ? 0 : ninterior;
    count = p.part == 0 ? ninterior : p.nlocal - ninterior;
  }
  for (int base = blockIdx.x * groups_per_block; base < count; base += gridDim.x * groups_per_block) {
    const int gid = base + threadIdx.x / LPA;
    const bool active = gid < count;
    const int idx = active ? gid : count - 1;
    const int i = p.site_list ? p.site_list[first + idx] : idx;
    const double4 ri = p.pos[i];
    const double li = ri.w, ai = 1.0 - li;
    int jnum = active ? p.numneigh[i] : 0;
    if (p.levcnt && level < 7 && active) {
      const uint4 lc = p.levcnt[i];
      const unsigned w = level < 2 ? lc.x : (level < 4 ? lc.y : (level < 6 ? lc.z : lc.w));
      jnum = min(jnum, (int)((level & 1) ? (w >> 16) : (w & 0xffffu)));
    }
    const int *row = p.neigh + (size_t)i * p.stride;
    double fx = 0, fy = 0, fz = 0, accA = 0, accB = 0, S0 = 0, S1 = 0;
    double vir[6] = {0, 0, 0, 0, 0, 0};

    // row entries: with LPA == 4 each lane fetches its next four (logical sub, 4+sub, 8+sub, 12+sub of a
    // 16-entry block) with one 16-byte load from the transposed row storage (ucg_internal.cuh, rowslot);
    // the block after the current one is already in flight
    const int4 *rp = reinterpret_cast<const int4 *>(row) + sub;
    int4 q = make_int4(0, 0, 0, 0), qn = q;
    int qm = 0, qb = 0;
    if (LPA == 4) {
      if (sub < jnum) q = __ldg(rp);
      if (16 + sub < jnum) qn = __ldg(rp + 4);
    }
    auto entry = [&](int jj_) -> int {
      if (LPA == 4) return (qm == 0 ? q.x : (qm == 1 ? q.y : (qm == 2 ? q.z : q.w))) & UCG_NEIGHMASK;
      return row[rowslot(jj_)] & UCG_NEIGHMASK;
    };
    auto advance = [&]() {
      if (LPA == 4) {
        qm = (qm + 1) & 3;
        if (qm == 0) {
          qb++;
          q = qn;
          if (16 * (qb + 1) + sub < jnum) qn = __ldg(rp + 4 * (qb + 1));
        }
      }
    };
    int jj = sub;
    int j = -1, sj = 0;
    double4 rj = ri;
    if (jj < jnum) { j = entry(jj); rj = (TEX == 1 || TEX == 3) ? ldtex(p.postex, j) : (TEX == 2 ? ldtex_half(p.postex, p.pos, j) : ld256(p.pos + j)); sj = (TEX == 3 ? tex1Dfetch<unsigned>(p.sbtex, j >> 5) : p.sbits[j >> 5]) >> (j & 31); }
    while (j >= 0) {
      int jn = -1, sn = 0;
      double4 rn = rj;
      jj += LPA;
      advance();
      if (PF) {
        if (jj < jnum) { jn = entry(jj); rn = (TEX == 1 || TEX == 3) ? ldtex(p.postex, jn) : (TEX == 2 ? ldtex_half(p.postex, p.pos, jn) : ld256(p.pos + jn)); sn = (TEX == 3 ? tex1Dfetch<unsigned>(p.sbtex, jn >> 5) : p.sbits[jn >> 5]) >> (jn & 31); }
      }
      const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
      const double rsq = rsq_exact(dx, dy, dz);
      if (rsq < p.cutsq) {
        const int it = (int)__dmul_rn(__dadd_rn(rsq, -p.innersq), p.invdelta);
        if (rsq < p.innersq || it >= tlm1) {
          report_error(p.err, rsq < p.innersq ? UCGB200_ERR_TABLE_INNER : UCGB200_ERR_TABLE_OUTER, p.tag[i], p.tag[j], rsq);
        } else {
          const double rsq_it = __dadd_rn(p.innersq, __dmul_rn((double)it, p.delta));
          const double frac = (rsq - rsq_it) * p.invdelta;
          double2 a00, a01, a11, b00, b01, b11, a10, b10;
          if (SM) {
            const double2 *r0 = s_tab + it * W;
            a00 = r0[0]; a01 = r0[1]; a11 = r0[W - 1];
            b00 = r0[W]; b01 = r0[W + 1]; b11 = r0[2 * W - 1];
            if (W == 4) { a10 = r0[2]; b10 = r0[W + 2]; }
          } else {
            const double2 *r0 = p.table + it * W;
            a00 = __ldg(r0); a01 = __ldg(r0 + 1); a11 = __ldg(r0 + W - 1);
            b00 = __ldg(r0 + W); b01 = __ldg(r0 + W + 1); b11 = __ldg(r0 + 2 * W - 1);
            if (W == 4) { a10 = __ldg(r0 + 2); b10 = __ldg(r0 + W + 2); }
          }
          const double u00 = a00.x + frac * (b00.x - a00.x), f00 = a00.y + frac * (b00.y - a00.y);
          const double u01 = a01.x + frac * (b01.x - a01.x), f01 = a01.y + frac * (b01.y - a01.y);
          const double u11 = a11.x + frac * (b11.x - a11.x), f11 = a11.y + frac * (b11.y - a11.y);
          double u10, f10;
          if (W == 4) {
            u10 = a10.x + frac * (b10.x - a10.x);
            f10 = a10.y + frac * (b10.y - a10.y);
          } else { u10 = u01; f10 = f01; }
          const double lj = rj.w, bj = 1.0 - lj;
          const double A = bj * u00 + lj * u01, B = bj * u10 + lj * u11;
          const double FA = bj * f00 + lj * f01, FB = bj * f10 + lj * f11;
          accA += A; accB += B;
          const double fpair = ai * FA + li * FB;
          const bool s1 = sj & 1;
          S0 += s1 ? u01 : u00;
          S1 += s1 ? u11 : u10;
          fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
          if (EV) {
            vir[0] += dx * dx * fpair; vir[1] += dy * dy * fpair; vir[2] += dz * dz * fpair;
            vir[3] += dx * dy * fpair; vir[4] += dx * dz * fpair; vir[5] += dy * dz * fpair;
          }
        }
      }
      if (!PF) {
        if (jj < jnum) { jn = entry(jj); rn = (TEX == 1 || TEX == 3) ? ldtex(p.postex, jn) : (TEX == 2 ? ldtex_half(p.postex, p.pos, jn) : ld256(p.pos + jn)); sn = (TEX == 3 ? tex1Dfetch<unsigned>(p.sbtex, jn >> 5) : p.sbits[jn >> 5]) >> (jn & 31); }
      }
      j = jn; rj = rn; sj = sn;
    }
    fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
    accA = group_sum<LPA>(accA); accB = group_sum<LPA>(accB);
    S0 = group_sum<LPA>(S0); S1 = group_sum<LPA>(S1);
    if (active && sub == 0) {
      double4 fo;
      fo.x = fx; fo.y = fy; fo.z = fz;
      fo.w = -p.dmu - (accB - accA);
      p.frc[i] = fo;
      p.scores[i] = make_double2(-S0 * p.inv_kT, -p.dmu * p.inv_kT - S1 * p.inv_kT);
      if (EV) ev[0] += 0.5 * (ai * accA + li * accB);
      if (EV && p.eatom) p.eatom[i] = 0.5 * (ai * accA + li * accB);
    }
    if (EV) {
#pragma unroll
      for (int k = 0; k < 6; k++) {
        double v = group_sum<LPA>(vir[k]);
        if (active && sub == 0) { ev[1 + k] += 0.5 * v; if (p.vatom) p.vatom[6 * (size_t)i + k] = 0.5 * v; }
      }
    }
  }
  if (EV) block_reduce_store<7, BS>(ev, p.partials);
}

// ------------------------------------------------------------- Newton's-third-law variant (experiment, UCGB200_N3L=1)
// Every owned-owned pair is evaluated ONCE, by the site with the smaller index: the i side accumulates in registers
// as in the fast kernel, the j side (f[3], ucgforce, 2 scores: pair_table_ucgld.cpp:500-502, :516, :527-529) is
// scattered with one red.global.add.f64 per component.  Pairs with a ghost partner are evaluated from the owned side
// only (no reverse halo), exactly as in the full-list kernel.  The kernel walks the same full rows and skips the
// entries its partner owns, so rows, skin levels and ghosts need no second list.  Results are no longer bit-
// reproducible (atomic order), but stay inside the north-star tolerances; frc / scores must be zero on entry and every
// site adds its own sums atomically as well.  Measured against the full-list kernel in profiles/r02_pair_n3l.json.
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

template <bool EV, int W, int BS>
__global__ void __launch_bounds__(BS) k_pair_ucgld_n3l(FastArgs p) {
  constexpr int LPA = 4;
  extern __shared__ double2 s_tab[];
  {
    const int nwords = p.tablen * W;
    for (int k = threadIdx.x; k < nwords; k += BS) s_tab[k] = p.table[k];
    __syncthreads();
  }
  const int sub = threadIdx.x % LPA;
  const int groups_per_block = BS / LPA;
  const int tlm1 = p.tablen - 1;
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  int level = 7;
  if (p.levcnt) {
    const double md = sqrt(__longlong_as_double((long long)*p.maxdisp)) * (1.0 + 1e-12);
    level = min(7, (int)floor(2.0 * md * p.inv_w));
  }
  double *frc = reinterpret_cast<double *>(p.frc), *sco = reinterpret_cast<double *>(p.scores);
  for (int base = blockIdx.x * groups_per_block; base < p.nlocal; base += gridDim.x * groups_per_block) {
    const int gid = base + threadIdx.x / LPA;
    const bool active = gid < p.nlocal;
    const int i = active ? gid : p.nlocal - 1;
    const double4 ri = p.pos[i];
    const double li = ri.w, ai = 1.0 - li;
    const bool si1 = (p.sbits[i >> 5] >> (i & 31)) & 1;
    int jnum = active ? p.numneigh[i] : 0;
    if (p.levcnt && level < 7 && active) {
      const uint4 lc = p.levcnt[i];
      const unsigned w = level < 2 ? lc.x : (level < 4 ? lc.y : (level < 6 ? lc.z : lc.w));
      jnum = min(jnum, (int)((level & 1) ? (w >> 16) : (w & 0xffffu)));
    }
    const int *row = p.neigh + (size_t)i * p.stride;
    double fx = 0, fy = 0, fz = 0, accA = 0, accB = 0, S0 = 0, S1 = 0;
    double vir[6] = {0, 0, 0, 0, 0, 0};
    RowWalk<LPA> rw(row, sub, jnum);
    // next entry this lane has to evaluate: entries owned by the partner (j owned, j < i) are skipped before any gather
    auto next_entry = [&](int &jj) -> int {
      while (jj < jnum) {
        const int j = rw.raw(jj) & UCG_NEIGHMASK;
        if (j > i) return j;     // ghosts have j >= nlocal > i
        jj += LPA;
        rw.advance();
      }
      return -1;
    };
    int jj = sub;
    int j = next_entry(jj), sj = 0;
    double4 rj = ri;
    if (j >= 0) { rj = ldtex(p.postex, j); sj = tex1Dfetch<unsigned>(p.sbtex, j >> 5) >> (j & 31); }
    while (j >= 0) {
      jj += LPA;
      rw.advance();
      const int jn = next_entry(jj);
      int sn = 0;
      double4 rn = rj;
      if (jn >= 0) { rn = ldtex(p.postex, jn); sn = tex1Dfetch<unsigned>(p.sbtex, jn >> 5) >> (jn & 31); }
      const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
      const double rsq = rsq_exact(dx, dy, dz);
      if (rsq < p.cutsq) {
        const int it = (int)__dmul_rn(__dadd_rn(rsq, -p.innersq), p.invdelta);
        if (rsq < p.innersq || it >= tlm1) {
          report_error(p.err, rsq < p.innersq ? UCGB200_ERR_TABLE_INNER : UCGB200_ERR_TABLE_OUTER, p.tag[i], p.tag[j], rsq);
        } else {
          const double rsq_it = __dadd_rn(p.innersq, __dmul_rn((double)it, p.delta));
          const double frac = (rsq - rsq_it) * p.invdelta;
          const double2 *r0 = s_tab + it * W;
          const double2 a00 = r0[0], a01 = r0[1], a11 = r0[W - 1];
          const double2 b00 = r0[W], b01 = r0[W + 1], b11 = r0[2 * W - 1];
          const double u00 = a00.x + frac * (b00.x - a00.x), f00 = a00.y + frac * (b00.y - a00.y);
          const double u01 = a01.x + frac * (b01.x - a01.x), f01 = a01.y + frac * (b01.y - a01.y);
          const double u11 = a11.x + frac * (b11.x - a11.x), f11 = a11.y + frac * (b11.y - a11.y);
          double u10, f10;
          if (W == 4) {
            const double2 a10 = r0[2], b10 = r0[W + 2];
            u10 = a10.x + frac * (b10.x - a10.x);
            f10 = a10.y + frac * (b10.y - a10.y);
          } else { u10 = u01; f10 = f01; }
          const double lj = rj.w, bj = 1.0 - lj;
          const double A = bj * u00 + lj * u01, B = bj * u10 + lj * u11;
          const double FA = bj * f00 + lj * f01, FB = bj * f10 + lj * f11;
          accA += A; accB += B;
          const double fpair = ai * FA + li * FB;
          const bool s1 = sj & 1;
          S0 += s1 ? u01 : u00;
          S1 += s1 ? u11 : u10;
          const double px = dx * fpair, py = dy * fpair, pz = dz * fpair;
          fx += px; fy += py; fz += pz;
          const bool owned = j < p.nlocal;
          if (owned) {
            // the partner's side of the same pair (:500-502, :516, :527-529)
            red_add(frc + 4 * (size_t)j, -px); red_add(frc + 4 * (size_t)j + 1, -py); red_add(frc + 4 * (size_t)j + 2, -pz);
            red_add(frc + 4 * (size_t)j + 3, -(li * (u11 - u10) + ai * (u01 - u00)));
            red_add(sco + 2 * (size_t)j, -(si1 ? u10 : u00) * p.inv_kT);
            red_add(sco + 2 * (size_t)j + 1, -(si1 ? u11 : u01) * p.inv_kT);
          }
          if (EV) {
            // a pair with a ghost partner is seen again from the image's owner: half of it belongs here
            const double wgt = owned ? 1.0 : 0.5;
            const double e = (ai * A + li * B) * wgt, fw = fpair * wgt;
            ev[0] += e;
            vir[0] += dx * dx * fw; vir[1] += dy * dy * fw; vir[2] += dz * dz * fw;
            vir[3] += dx * dy * fw; vir[4] += dx * dz * fw; vir[5] += dy * dz * fw;
          }
        }
      }
      j = jn; rj = rn; sj = sn;
    }
    fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
    accA = group_sum<LPA>(accA); accB = group_sum<LPA>(accB);
    S0 = group_sum<LPA>(S0); S1 = group_sum<LPA>(S1);
    if (active && sub == 0) {
      red_add(frc + 4 * (size_t)i, fx); red_add(frc + 4 * (size_t)i + 1, fy); red_add(frc + 4 * (size_t)i + 2, fz);
      red_add(frc + 4 * (size_t)i + 3, -p.dmu - (accB - accA));
      red_add(sco + 2 * (size_t)i, -S0 * p.inv_kT);
      red_add(sco + 2 * (size_t)i + 1, -p.dmu * p.inv_kT - S1 * p.inv_kT);
    }
    if (EV) {
#pragma unroll
      for (int k = 0; k < 6; k++) ev[1 + k] += vir[k];
    }
  }
  if (EV) block_reduce_store<7, BS>(ev, p.partials);
}

// ---------------------------------------------- Newton's third law with TMA bulk reductions (UCGB200_N3L=2)
// Same pair ownership as k_pair_ucgld_n3l, but the partner's six sums {fx, fy, fz, ucgforce, s0, s1} leave the SM as
// ONE 48-byte cp.reduce.async.bulk.global.shared::cta.add.f64 (SASS UBLKRED) from a per-lane staging slot in shared
// memory instead of six red.global.add.f64: the reduction is carried out by the L2 slices, the SM only issues the
// descriptor.  Measured in isolation (csrc/microbench.cu, profiles/r02_pair_floors.json) 27.5 M such operations take
// 0.35 ms against 0.87 ms for the six REDs.  Accumulator layout: acc[j] = 6 contiguous doubles (48-byte records, zero on
// entry); k_merge_acc turns them into the frc / scores records.  The staging slots sit behind the 192 KB table:
// 48 B x 704 threads is what fits into 227 KB.
__device__ __forceinline__ void bulk_red_48(double *gdst, const double *slot) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 48;" ::"l"(gdst), "r"(sa) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <bool EV, int W, int BS>
__global__ void __launch_bounds__(BS) k_pair_ucgld_n3l_bulk(FastArgs p, double *__restrict__ acc) {
  constexpr int LPA = 4;
  extern __shared__ double2 s_tab[];
  double *slot = reinterpret_cast<double *>(s_tab + (size_t)p.tablen * W) + threadIdx.x * 6;
  {
    const int nwords = p.tablen * W;
    for (int k = threadIdx.x; k < nwords; k += BS) s_tab[k] = p.table[k];
    __syncthreads();
  }
  const int sub = threadIdx.x % LPA;
  const int groups_per_block = BS / LPA;
  const int tlm1 = p.tablen - 1;
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  int level = 7;
  if (p.levcnt) {
    const double md = sqrt(__longlong_as_double((long long)*p.maxdisp)) * (1.0 + 1e-12);
    level = min(7, (int)floor(2.0 * md * p.inv_w));
  }
  for (int base = blockIdx.x * groups_per_block; base < p.nlocal; base += gridDim.x * groups_per_block) {
    const int gid = base + threadIdx.x / LPA;
    const bool active = gid < p.nlocal;
    const int i = active ? gid : p.nlocal - 1;
    const double4 ri = p.pos[i];
    const double li = ri.w, ai = 1.0 - li;
    const bool si1 = (p.sbits[i >> 5] >> (i & 31)) & 1;
    int jnum = active ? p.numneigh[i] : 0;
    if (p.levcnt && level < 7 && active) {
      const uint4 lc = p.levcnt[i];
      const unsigned w = level < 2 ? lc.x : (level < 4 ? lc.y : (level < 6 ? lc.z : lc.w));
      jnum = min(jnum, (int)((level & 1) ? (w >> 16) : (w & 0xffffu)));
    }
    const int *row = p.neigh + (size_t)i * p.stride;
    double fx = 0, fy = 0, fz = 0, accA = 0, accB = 0, S0 = 0, S1 = 0;
    double vir[6] = {0, 0, 0, 0, 0, 0};
    RowWalk<LPA> rw(row, sub, jnum);
    auto next_entry = [&](int &jj) -> int {
      while (jj < jnum) {
        const int j = rw.raw(jj) & UCG_NEIGHMASK;
        if (j > i) return j;
        jj += LPA;
        rw.advance();
      }
      return -1;
    };
    int jj = sub;
    int j = next_entry(jj), sj = 0;
    double4 rj = ri;
    if (j >= 0) { rj = ldtex(p.postex, j); sj = tex1Dfetch<unsigned>(p.sbtex, j >> 5) >> (j & 31); }
    while (j >= 0) {
      jj += LPA;
      rw.advance();
      const int jn = next_entry(jj);
      int sn = 0;
      double4 rn = rj;
      if (jn >= 0) { rn = ldtex(p.postex, jn); sn = tex1Dfetch<unsigned>(p.sbtex, jn >> 5) >> (jn & 31); }
      const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
      const double rsq = rsq_exact(dx, dy, dz);
      if (rsq < p.cutsq) {
        const int it = (int)__dmul_rn(__dadd_rn(rsq, -p.innersq), p.invdelta);
        if (rsq < p.innersq || it >= tlm1) {
          report_error(p.err, rsq < p.innersq ? UCGB200_ERR_TABLE_INNER : UCGB200_ERR_TABLE_OUTER, p.tag[i], p.tag[j], rsq);
        } else {
          const double rsq_it = __dadd_rn(p.innersq, __dmul_rn((double)it, p.delta));
          const double frac = (rsq - rsq_it) * p.invdelta;
          const double2 *r0 = s_tab + it * W;
          const double2 a00 = r0[0], a01 = r0[1], a11 = r0[W - 1];
          const double2 b00 = r0[W], b01 = r0[W + 1], b11 = r0[2 * W - 1];
          const double u00 = a00.x + frac * (b00.x - a00.x), f00 = a00.y + frac * (b00.y - a00.y);
          const double u01 = a01.x + frac * (b01.x - a01.x), f01 = a01.y + frac * (b01.y - a01.y);
          const double u11 = a11.x + frac * (b11.x - a11.x), f11 = a11.y + frac * (b11.y - a11.y);
          double u10, f10;
          if (W == 4) {
            const double2 a10 = r0[2], b10 = r0[W + 2];
            u10 = a10.x + frac * (b10.x - a10.x);
            f10 = a10.y + frac * (b10.y - a10.y);
          } else { u10 = u01; f10 = f01; }
          const double lj = rj.w, bj = 1.0 - lj;
          const double A = bj * u00 + lj * u01, B = bj * u10 + lj * u11;
          const double FA = bj * f00 + lj * f01, FB = bj * f10 + lj * f11;
          accA += A; accB += B;
          const double fpair = ai * FA + li * FB;
          const bool s1 = sj & 1;
          S0 += s1 ? u01 : u00;
          S1 += s1 ? u11 : u10;
          const double px = dx * fpair, py = dy * fpair, pz = dz * fpair;
          fx += px; fy += py; fz += pz;
          const bool owned = j < p.nlocal;
          if (owned) {
            // the staging slot is free again once the previous reduction has READ it
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            slot[0] = -px; slot[1] = -py; slot[2] = -pz;
            slot[3] = -(li * (u11 - u10) + ai * (u01 - u00));
            slot[4] = -(si1 ? u10 : u00) * p.inv_kT;
            slot[5] = -(si1 ? u11 : u01) * p.inv_kT;
            bulk_red_48(acc + 6 * (size_t)j, slot);
          }
          if (EV) {
            const double wgt = owned ? 1.0 : 0.5;
            const double e = (ai * A + li * B) * wgt, fw = fpair * wgt;
            ev[0] += e;
            vir[0] += dx * dx * fw; vir[1] += dy * dy * fw; vir[2] += dz * dz * fw;
            vir[3] += dx * dy * fw; vir[4] += dx * dz * fw; vir[5] += dy * dz * fw;
          }
        }
      }
      j = jn; rj = rn; sj = sn;
    }
    fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
    accA = group_sum<LPA>(accA); accB = group_sum<LPA>(accB);
    S0 = group_sum<LPA>(S0); S1 = group_sum<LPA>(S1);
    if (active && sub == 0) {
      double *a = acc + 6 * (size_t)i;
      red_add(a, fx); red_add(a + 1, fy); red_add(a + 2, fz);
      red_add(a + 3, -p.dmu - (accB - accA));
      red_add(a + 4, -S0 * p.inv_kT);
      red_add(a + 5, -p.dmu * p.inv_kT - S1 * p.inv_kT);
    }
    if (EV) {
#pragma unroll
      for (int k = 0; k < 6; k++) ev[1 + k] += vir[k];
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (EV) block_reduce_store<7, BS>(ev, p.partials);
}

__global__ void k_merge_acc(const double *__restrict__ acc, int n, double4 *__restrict__ frc, double2 *__restrict__ scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 a = reinterpret_cast<const double2 *>(acc)[3 * (size_t)i], b = reinterpret_cast<const double2 *>(acc)[3 * (size_t)i + 1],
                c = reinterpret_cast<const double2 *>(acc)[3 * (size_t)i + 2];
  frc[i] = make_double4(a.x, a.y, b.x, b.y);
  scores[i] = c;
}

__global__ void k_reduce_partials(const double *__restrict__ partials, int nblocks, int nvals, double *__restrict__ out) {
  // one warp per value, fixed order => deterministic
  int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k >= nvals) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * nvals + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[k] = s;
}

}  // namespace

int ucg::reduce_partials(ucgb200_ctx *c, int nblocks, int nvals, int out_offset) {
  k_reduce_partials<<<1, 32 * nvals, 0, c->stream>>>(c->d_partials.p, nblocks, nvals, c->d_ev.p + out_offset);
  UCG_LAUNCHED(c);
  return 0;
}

static int g_sm_count = 0;
static int sm_count(int device) {
  if (!g_sm_count) cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, device);
  return g_sm_count ? g_sm_count : 148;
}

// tunables (set through the environment for experiments; defaults chosen by ncu)
static int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s ? atoi(s) : dflt;
}

template <int LPA, bool EV, int W, int BS, int PF>
static int launch_fast(ucgb200_ctx *c, FastArgs &a, int &nblk) {
  size_t smem = a.smem_table ? (size_t)a.tablen * W * sizeof(double2) : 0;
  auto kern = a.smem_table ? k_pair_ucgld_fast<LPA, EV, W, BS, PF, true> : k_pair_ucgld_fast<LPA, EV, W, BS, PF, false>;
  if (a.postex && a.smem_table && LPA == 4 && !EV && PF == 1)
    kern = env_int("UCGB200_TEX", 3) == 2 ? k_pair_ucgld_fast<LPA, EV, W, BS, PF, true, 2>
           : (env_int("UCGB200_TEX", 3) == 3 ? k_pair_ucgld_fast<LPA, EV, W, BS, PF, true, 3> : k_pair_ucgld_fast<LPA, EV, W, BS, PF, true, 1>);
  if (smem > 32 * 1024) UCG_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  if (!a.smem_table) per_sm = 2048 / BS > 3 ? 3 : 2048 / BS;
  else if (smem <= 100 * 1024 && BS <= 512) per_sm = 2;
  nblk = sm_count(c->device) * per_sm;
  int groups = BS / LPA;
  int need = (a.nlocal + groups - 1) / groups;
  if (nblk > need) nblk = need;
  if (EV) UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 8 + 64));
  a.partials = c->d_partials.p;
  kern<<<nblk, BS, smem, c->stream>>>(a);
  UCG_LAUNCHED(c);
  return 0;
}

template <bool EV, int W>
static int launch_n3l(ucgb200_ctx *c, FastArgs &a, int &nblk) {
  constexpr int BS = EV ? 512 : 768;   // the energy/virial variant spills at 768 threads (80 registers)
  const size_t smem = (size_t)a.tablen * W * sizeof(double2);
  auto kern = k_pair_ucgld_n3l<EV, W, BS>;
  UCG_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nblk = sm_count(c->device);
  const int groups = BS / 4, need = (a.nlocal + groups - 1) / groups;
  if (nblk > need) nblk = need;
  if (EV) UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 8 + 64));
  a.partials = c->d_partials.p;
  UCG_CHECK(c, cudaMemsetAsync(c->frc.p, 0, (size_t)a.nlocal * sizeof(double4), c->stream));
  UCG_CHECK(c, cudaMemsetAsync(c->scores.p, 0, (size_t)a.nlocal * sizeof(double2), c->stream));
  kern<<<nblk, BS, smem, c->stream>>>(a);
  UCG_LAUNCHED(c);
  return 0;
}

template <bool EV, int W>
static int launch_n3l_bulk(ucgb200_ctx *c, FastArgs &a, int &nblk) {
  constexpr int BS = EV ? 512 : 704;   // 48-byte staging slot per thread behind the table: 196608 + 48 * 704 <= 227 KB
  const size_t smem = (size_t)a.tablen * W * sizeof(double2) + (size_t)BS * 48;
  if (smem > 227 * 1024) return fail(c, "N3L bulk variant: table + staging slots exceed shared memory");
  auto kern = k_pair_ucgld_n3l_bulk<EV, W, BS>;
  UCG_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nblk = sm_count(c->device);
  const int groups = BS / 4, need = (a.nlocal + groups - 1) / groups;
  if (nblk > need) nblk = need;
  if (EV) UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 8 + 64));
  a.partials = c->d_partials.p;
  UCG_CHECK(c, c->pair_acc.ensure((size_t)a.nlocal * 6 + 8));
  UCG_CHECK(c, cudaMemsetAsync(c->pair_acc.p, 0, (size_t)a.nlocal * 6 * sizeof(double), c->stream));
  kern<<<nblk, BS, smem, c->stream>>>(a, c->pair_acc.p);
  UCG_LAUNCHED(c);
  k_merge_acc<<<nblocks(a.nlocal, 256), 256, 0, c->stream>>>(c->pair_acc.p, a.nlocal, c->frc.p, c->scores.p);
  UCG_LAUNCHED(c);
  return 0;
}

template <int LPA, int W>
static int dispatch_fast(ucgb200_ctx *c, FastArgs &a, int &nblk, bool ev, int bs, int pf) {
  // thermo steps are rare: one (BS, PF) variant is enough for EV
  if (ev) return launch_fast<LPA, true, W, 512, 1>(c, a, nblk);
  // 24 warps per SM (768 threads, 80 registers): with the gathers on the texture pipe neither data pipe is
  // saturated any more and the extra warps pay (0.64 -> 0.58 ms); 640 / 896 / 1024 threads were slower
  // (0.65 / 0.69 / 0.89 ms: the last two spill)
  if (LPA == 4 && pf && bs == 768) return launch_fast<LPA, false, W, 768, 1>(c, a, nblk);
  return pf ? launch_fast<LPA, false, W, 512, 1>(c, a, nblk) : launch_fast<LPA, false, W, 512, 0>(c, a, nblk);
}

template <int LPA, bool EV>
static int launch_general(ucgb200_ctx *c, PairArgs &a, int &nblk) {
  constexpr int BS = 256;
  long long nthreads = (long long)a.nlocal * LPA;
  nblk = nblocks(nthreads, BS);
  if (EV) UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 8 + 64));
  a.partials = c->d_partials.p;
  k_pair_ucgld<LPA, EV, BS><<<nblk, BS, 0, c->stream>>>(a);
  UCG_LAUNCHED(c);
  return 0;
}

// run.cu: may the next non-thermo evaluation be split into its interior and boundary parts (pair_part 0 / 1)?
int ucg_pair_can_split(ucgb200_ctx *c) {
  return c->fast_uniform && c->parts_valid && c->list_valid && !env_int("UCGB200_FORCE_GENERAL", 0) && !env_int("UCGB200_N3L", 0) &&
         env_int("UCGB200_LPA", 4) == 4 && env_int("UCGB200_TEX", 3) && env_int("UCGB200_SMEM_TABLE", 1) &&
         (size_t)c->fast_len * c->fast_ntab * sizeof(double2) <= 220 * 1024;
}

extern "C" int ucgb200_pair_ucgld(ucgb200_ctx *c, int eflag, int vflag) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (!c->list_valid) return fail(c, "pair_ucgld: neighbor list not built");
  c->ev_valid = false;
  c->ev_two_parts = false;
  if (c->nlocal == 0) {   // an empty brick still reports (zero) energy and virial
    UCG_CHECK(c, cudaMemsetAsync(c->d_ev.p, 0, 32 * sizeof(double), c->stream));
    c->ev_valid = true;
    c->ev_two_parts = false;
    return 0;
  }
  const bool ev = eflag || vflag;
  // LAMMPS flag bits: eflag & 2 = ENERGY_ATOM, vflag & 4 = VIRIAL_ATOM
  const bool want_eatom = (eflag & 2) != 0, want_vatom = (vflag & 4) != 0;
  c->eatom_valid = c->vatom_valid = false;
  double *d_eatom = nullptr, *d_vatom = nullptr;
  if (want_eatom) { UCG_CHECK(c, c->d_eatom.ensure((size_t)c->nlocal + 8)); d_eatom = c->d_eatom.p; }
  if (want_vatom) { UCG_CHECK(c, c->d_vatom.ensure(6 * (size_t)c->nlocal + 8)); d_vatom = c->d_vatom.p; }
  int nblk = 0;
  const bool timed = c->timers_on;
  if (timed) cudaEventRecord(c->ev_pair0, c->stream);
  const int force_general = env_int("UCGB200_FORCE_GENERAL", 0);
  const int lpa_fast = env_int("UCGB200_LPA", 4);
  const int smem_pref = env_int("UCGB200_SMEM_TABLE", 1);
  if (c->fast_uniform && !force_general) {
    FastArgs a{};
    {
      const int nall = c->nlocal + c->nghost;
      UCG_CHECK(c, c->statebits.ensure((size_t)nall / 32 + 8));
      k_pack_statebits<<<nblocks(nall, 256), 256, 0, c->stream>>>(c->ts.p, nall, c->statebits.p);
      UCG_LAUNCHED(c);
    }
    a.pos = c->pos.p; a.ts = c->ts.p; a.sbits = c->statebits.p; a.tag = c->tag.p; a.nlocal = c->nlocal;
    a.neigh = c->neigh.p; a.stride = c->neigh_stride; a.numneigh = c->numneigh.p;
    a.table = c->d_fast_table.p; a.tablen = c->fast_len;
    const ucg::TableDev &t0 = c->tables[c->fast_tab[0]];
    a.innersq = t0.innersq; a.delta = t0.delta; a.invdelta = t0.invdelta;
    a.cutsq = c->cutsq[1 * (c->n_formal + 1) + 1];
    a.dmu = c->chem_pot[c->formal_from[2 * 1 + 1]] - c->chem_pot[c->formal_from[2 * 1 + 0]];
    a.inv_kT = 1.0 / c->kT;
    a.frc = c->frc.p; a.scores = c->scores.p; a.err = c->d_err.p;
    size_t smem = (size_t)c->fast_len * c->fast_ntab * sizeof(double2);
    a.smem_table = (smem_pref && smem <= 220 * 1024) ? 1 : 0;
    a.levcnt = nullptr; a.maxdisp = c->d_maxdisp.p; a.inv_w = c->skin > 0.0 ? 8.0 / c->skin : 0.0;
    if (c->maxdisp_valid && c->skin > 0.0 && env_int("UCGB200_SKIN_LEVELS", 1)) a.levcnt = c->levcnt.p;
    a.postex = 0; a.sbtex = 0;
    a.site_list = nullptr; a.part_count = nullptr; a.part = -1;
    a.eatom = d_eatom; a.vatom = d_vatom;
    if (c->pair_part >= 0 && c->parts_valid && !ev && !env_int("UCGB200_N3L", 0)) {
      a.site_list = c->site_list.p; a.part_count = c->d_part.p; a.part = c->pair_part;
    }
    c->pair_part = -1;
    const int tex_mode = env_int("UCGB200_TEX", 3);   // 0: every gather through the LSU pipe
    if (tex_mode) {
      if ((rc = ucg_bind_gather_textures(c, &a.postex, nullptr))) return rc;
      if ((rc = ucg_bind_texture(c, c->tex_sbits, c->statebits.p, c->statebits.cap * sizeof(unsigned), cudaCreateChannelDesc<unsigned>()))) return rc;
      a.sbtex = c->tex_sbits.tex;
    }
    const int bs = env_int("UCGB200_BS", 768), pf = env_int("UCGB200_PF", 1);
    const int n3l = (want_eatom || want_vatom) ? 0 : env_int("UCGB200_N3L", 0);   // the experiments carry no per-atom tallies
    if (n3l == 2 && a.postex && a.sbtex && a.smem_table && c->fast_ntab == 3) {
      rc = ev ? launch_n3l_bulk<true, 3>(c, a, nblk) : launch_n3l_bulk<false, 3>(c, a, nblk);
    } else if (n3l && a.postex && a.sbtex && a.smem_table) {
      if (c->fast_ntab == 3) rc = ev ? launch_n3l<true, 3>(c, a, nblk) : launch_n3l<false, 3>(c, a, nblk);
      else rc = ev ? launch_n3l<true, 4>(c, a, nblk) : launch_n3l<false, 4>(c, a, nblk);
    } else if (c->fast_ntab == 3) {
      if (lpa_fast == 4) rc = dispatch_fast<4, 3>(c, a, nblk, ev, bs, pf);
      else if (lpa_fast == 16) rc = dispatch_fast<16, 3>(c, a, nblk, ev, bs, pf);
      else rc = dispatch_fast<8, 3>(c, a, nblk, ev, bs, pf);
    } else {
      if (lpa_fast == 4) rc = dispatch_fast<4, 4>(c, a, nblk, ev, bs, pf);
      else if (lpa_fast == 16) rc = dispatch_fast<16, 4>(c, a, nblk, ev, bs, pf);
      else rc = dispatch_fast<8, 4>(c, a, nblk, ev, bs, pf);
    }
  } else {
    PairArgs a{};
    a.pos = c->pos.p; a.ts = c->ts.p; a.tag = c->tag.p; a.nlocal = c->nlocal;
    a.neigh = c->neigh.p; a.stride = c->neigh_stride; a.numneigh = c->numneigh.p;
    a.pinfo = c->d_pairinfo.p; a.tinfo = c->d_typeinfo.p; a.na = c->n_actual + 1; a.tables = c->d_tables.p;
    for (int k = 0; k < 4; k++) a.special_lj[k] = c->special_lj[k];
    a.inv_kT = 1.0 / c->kT;
    a.frc = c->frc.p; a.scores = c->scores.p; a.err = c->d_err.p;
    a.eatom = d_eatom; a.vatom = d_vatom;
    if (ev) rc = launch_general<8, true>(c, a, nblk); else rc = launch_general<8, false>(c, a, nblk);
  }
  if (rc) return rc;
  if (timed) { cudaEventRecord(c->ev_pair1, c->stream); c->pair_timed = true; }
  if (ev) {
    if ((rc = reduce_partials(c, nblk, 7, 0))) return rc;
    c->ev_valid = true;
    c->eatom_valid = want_eatom;
    c->vatom_valid = want_vatom;
  }
  return 0;
}

namespace {
__global__ void k_peratom_to_host_order(const double *__restrict__ src, const int *__restrict__ orig, int n, int width, double *__restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const int i = t / width, k = t - i * width;
  dst[(size_t)orig[i] * width + k] = src[(size_t)i * width + k];
}
}  // namespace

extern "C" int ucgb200_pair_peratom(ucgb200_ctx *c, int cap, double *eatom, double *vatom) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  const int n = c->nlocal;
  if (cap < n) return fail(c, "pair_peratom: host capacity smaller than nlocal");
  if ((eatom && !c->eatom_valid) || (vatom && !c->vatom_valid))
    return fail(c, "pair_peratom: the last pair call did not tally per-atom values (eflag & 2 / vflag & 4), or the style does not support them");
  if (n == 0) return 0;
  UCG_CHECK(c, c->stage_d.ensure(16 * (size_t)n + 64));
  double *sd = c->stage_d.p;
  if (eatom) {
    k_peratom_to_host_order<<<nblocks(n, 256), 256, 0, c->stream>>>(c->d_eatom.p, c->orig.p, n, 1, sd);
    UCG_LAUNCHED(c);
    UCG_CHECK(c, cudaMemcpyAsync(eatom, sd, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (vatom) {
    k_peratom_to_host_order<<<nblocks((long long)n * 6, 256), 256, 0, c->stream>>>(c->d_vatom.p, c->orig.p, n, 6, sd + n);
    UCG_LAUNCHED(c);
    UCG_CHECK(c, cudaMemcpyAsync(vatom, sd + n, 6 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int ucgb200_pair_energy_virial(ucgb200_ctx *c, double *eng_vdwl, double virial[6]) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  if (!c->ev_valid) return fail(c, "no energy/virial available: last pair call had eflag=vflag=0");
  double h[7], h2[7] = {0, 0, 0, 0, 0, 0, 0};
  UCG_CHECK(c, cudaMemcpyAsync(h, c->d_ev.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (c->ev_two_parts) UCG_CHECK(c, cudaMemcpyAsync(h2, c->d_ev.p + 16, 7 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 7; k++) h[k] += h2[k];
  if (eng_vdwl) *eng_vdwl = h[0];
  if (virial) for (int k = 0; k < 6; k++) virial[k] = h[1 + k];
  return 0;
}

extern "C" int ucgb200_last_pair_ms(ucgb200_ctx *c, double *ms) {
  if (!c || !ms) return -1;
  if (!c->pair_timed) return fail(c, "pair timing not enabled (ucgb200_timers(ctx,1,...))");
  UCG_CHECK(c, cudaEventSynchronize(c->ev_pair1));
  float f = 0;
  UCG_CHECK(c, cudaEventElapsedTime(&f, c->ev_pair0, c->ev_pair1));
  *ms = f;
  return 0;
}
