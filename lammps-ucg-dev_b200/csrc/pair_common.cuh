// pair_common.cuh — table look-up and reduction helpers shared by the pair kernels.
#pragma once
#include "ucg_internal.cuh"

namespace ucg {

__device__ __forceinline__ void report_error(ErrWord *err, int code, int tag_i, int tag_j, double rsq) {
  if (atomicCAS(&err->code, 0, code) == 0) {
    err->tag_i = tag_i;
    err->tag_j = tag_j;
    err->rsq = rsq;
  }
}

// One tabulated potential at rsq: the "------- For one table -------" block that the
// reference repeats at pair_table_ucgld.cpp:223-268, 279-324, 354-399, 437-482 (and the
// clones in the other pair styles).  Returns 0, or UCGB200_ERR_TABLE_INNER/OUTER where the
// reference calls error->one().  e,f are NOT yet multiplied by factor_lj.
__device__ __forceinline__ int table_eval(const TableDev &tb, double rsq, double &e, double &f) {
  if (rsq < tb.innersq) return UCGB200_ERR_TABLE_INNER;
  const int tlm1 = tb.tablength - 1;
  if (tb.style == UCGB200_TAB_LINEAR) {
    // itable = static_cast<int>((rsq - innersq) * invdelta)            (:447)
    int it = (int)__dmul_rn(__dadd_rn(rsq, -tb.innersq), tb.invdelta);
    if (it >= tlm1) return UCGB200_ERR_TABLE_OUTER;
    // rsq[it] = innersq + it*delta, de = e[it+1]-e[it], df likewise    (compute_table :1157-1174)
    double rsq_it = __dadd_rn(tb.innersq, __dmul_rn((double)it, tb.delta));
    double frac = (rsq - rsq_it) * tb.invdelta;
    double2 a = __ldg(&tb.ef[it]), b = __ldg(&tb.ef[it + 1]);
    e = a.x + frac * (b.x - a.x);
    f = a.y + frac * (b.y - a.y);
  } else if (tb.style == UCGB200_TAB_SPLINE) {
    int it = (int)__dmul_rn(__dadd_rn(rsq, -tb.innersq), tb.invdelta);
    if (it >= tlm1) return UCGB200_ERR_TABLE_OUTER;
    double rsq_it = __dadd_rn(tb.innersq, __dmul_rn((double)it, tb.delta));
    double b = (rsq - rsq_it) * tb.invdelta;
    double a = 1.0 - b;
    double2 y0 = __ldg(&tb.ef[it]), y1 = __ldg(&tb.ef[it + 1]);
    double2 z0 = __ldg(&tb.ef2[it]), z1 = __ldg(&tb.ef2[it + 1]);
    double ca = a * a * a - a, cb = b * b * b - b;
    e = a * y0.x + b * y1.x + (ca * z0.x + cb * z1.x) * tb.deltasq6;
    f = a * y0.y + b * y1.y + (ca * z0.y + cb * z1.y) * tb.deltasq6;
  } else if (tb.style == UCGB200_TAB_LOOKUP) {
    int it = (int)__dmul_rn(__dadd_rn(rsq, -tb.innersq), tb.invdelta);
    if (it >= tlm1) return UCGB200_ERR_TABLE_OUTER;
    double2 a = __ldg(&tb.ef[it]);
    e = a.x; f = a.y;
  } else {  // BITMAP: index from the float bit pattern of rsq (:466-471)
    float rf = (float)rsq;
    int it = (__float_as_int(rf) & tb.nmask) >> tb.nshiftbits;
    double2 rd = __ldg(&tb.rd[it]);
    double frac = ((double)rf - rd.x) * rd.y;
    double2 a = __ldg(&tb.ef[it]), d = __ldg(&tb.dedf[it]);
    e = a.x + frac * d.x;
    f = a.y + frac * d.y;
  }
  return 0;
}

// ---- shared-memory table path (one 2-state actual type, LINEAR tables on one rsq grid; the layout
// built by rebuild_maps for pair_ucgld.cu): row it = {e00,f00, e01,f01, [e10,f10,] e11,f11}.
struct FastTable {
  const double2 *table;   // [tablen][W]
  int tablen, W;
  double innersq, delta, invdelta;
};
template <int W, int BS>
__device__ __forceinline__ void fast_table_stage(double2 *s_tab, const FastTable &ft) {
  const int nwords = ft.tablen * W;
  for (int k = threadIdx.x; k < nwords; k += BS) s_tab[k] = ft.table[k];
  __syncthreads();
}
// the four potentials u[a*2+b], f[a*2+b] at rsq; same index / fraction arithmetic as table_eval
template <int W>
__device__ __forceinline__ int fast_table_eval(const double2 *s_tab, const FastTable &ft, double rsq, double (&u)[4],
                                               double (&f)[4]) {
  if (rsq < ft.innersq) return UCGB200_ERR_TABLE_INNER;
  const int it = (int)__dmul_rn(__dadd_rn(rsq, -ft.innersq), ft.invdelta);
  if (it >= ft.tablen - 1) return UCGB200_ERR_TABLE_OUTER;
  const double rsq_it = __dadd_rn(ft.innersq, __dmul_rn((double)it, ft.delta));
  const double frac = (rsq - rsq_it) * ft.invdelta;
  const double2 *r0 = s_tab + it * W;
  const double2 a00 = r0[0], a01 = r0[1], a11 = r0[W - 1];
  const double2 b00 = r0[W], b01 = r0[W + 1], b11 = r0[2 * W - 1];
  u[0] = a00.x + frac * (b00.x - a00.x); f[0] = a00.y + frac * (b00.y - a00.y);
  u[1] = a01.x + frac * (b01.x - a01.x); f[1] = a01.y + frac * (b01.y - a01.y);
  u[3] = a11.x + frac * (b11.x - a11.x); f[3] = a11.y + frac * (b11.y - a11.y);
  if (W == 4) {
    const double2 a10 = r0[2], b10 = r0[W + 2];
    u[2] = a10.x + frac * (b10.x - a10.x); f[2] = a10.y + frac * (b10.y - a10.y);
  } else { u[2] = u[1]; f[2] = f[1]; }
  return 0;
}

// generic use of the same layout (any type system whose <= 4 uploaded tables are LINEAR on one grid;
// slot k of a row = table k): index / fraction once per pair, then one slot per (state, state) table
__device__ __forceinline__ int fast_table_index(const FastTable &ft, double rsq, int &it, double &frac) {
  if (rsq < ft.innersq) return UCGB200_ERR_TABLE_INNER;
  it = (int)__dmul_rn(__dadd_rn(rsq, -ft.innersq), ft.invdelta);
  if (it >= ft.tablen - 1) return UCGB200_ERR_TABLE_OUTER;
  const double rsq_it = __dadd_rn(ft.innersq, __dmul_rn((double)it, ft.delta));
  frac = (rsq - rsq_it) * ft.invdelta;
  return 0;
}
__device__ __forceinline__ void fast_table_slot(const double2 *s_tab, int W, int it, double frac, int slot, double &e, double &f) {
  const double2 a = s_tab[it * W + slot], b = s_tab[(it + 1) * W + slot];
  e = a.x + frac * (b.x - a.x);
  f = a.y + frac * (b.y - a.y);
}

// ---- row walking and texture-pipe gathers shared by the shared-memory-table kernels
// With LPA == 4 each lane fetches its next four row entries (logical sub, 4+sub, 8+sub, 12+sub of a 16-entry
// block of the transposed row storage, ucg_internal.cuh rowslot) with one 16-byte load; the block after the
// current one is already in flight.  Other LPA: plain rowslot() reads.
template <int LPA>
struct RowWalk {
  const int *row;
  const int4 *rp;
  int4 q, qn;
  int qm, qb, jnum, sub;
  __device__ __forceinline__ RowWalk(const int *row_, int sub_, int jnum_) : row(row_), qm(0), qb(0), jnum(jnum_), sub(sub_) {
    rp = reinterpret_cast<const int4 *>(row_) + sub_;
    q = make_int4(0, 0, 0, 0);
    qn = q;
    if (LPA == 4) {
      if (sub < jnum) q = __ldg(rp);
      if (16 + sub < jnum) qn = __ldg(rp + 4);
    }
  }
  __device__ __forceinline__ int raw(int jj) const {   // entry with its special-bond bits
    if (LPA == 4) return qm == 0 ? q.x : (qm == 1 ? q.y : (qm == 2 ? q.z : q.w));
    return row[rowslot(jj)];
  }
  __device__ __forceinline__ void advance() {
    if (LPA == 4) {
      qm = (qm + 1) & 3;
      if (qm == 0) {
        qb++;
        q = qn;
        if (16 * (qb + 1) + sub < jnum) qn = __ldg(rp + 4 * (qb + 1));
      }
    }
  }
};

// per-site gathers through the texture pipe, which works beside the (saturated) LSU data pipe; a zero
// texture object selects the plain load
struct GatherTex {
  cudaTextureObject_t pos;   // {x,y,z,lambda} records as two int4 texels
  cudaTextureObject_t ts;    // type | state << 16, one int texel per site
};
__device__ __forceinline__ double4 gather_pos(const GatherTex &g, const double4 *pos, int j) {
  if (g.pos) {
    const int4 a = tex1Dfetch<int4>(g.pos, 2 * j), b = tex1Dfetch<int4>(g.pos, 2 * j + 1);
    return make_double4(__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z));
  }
  return pos[j];
}
__device__ __forceinline__ int gather_ts(const GatherTex &g, const int *ts, int j) {
  return g.ts ? tex1Dfetch<int>(g.ts, j) : ts[j];
}

// sum `v` over the LPA lanes of a sub-warp group (result valid in every lane)
template <int LPA>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = LPA / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-level reduction of NV per-thread values into partials[blockIdx*NV + k];
// every thread of the block must call it.  Deterministic (fixed tree).
template <int NV, int BS>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *partials) {
  __shared__ double red[NV][BS / 32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[k][wid] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < BS / 32; w++) s += red[threadIdx.x][w];
    partials[(size_t)blockIdx.x * NV + threadIdx.x] = s;
  }
}

// ---- CV back-force sweep of the two density styles (pair_table_ucg_bethe_density.cpp:696-726,
// pair_table_rleucg_interface.cpp:466-502) when every density type shares ONE threshold radius and every pair ONE cutoff:
// the sweep then needs nothing of a neighbor but its position and "its CV force if it is a density site this sweep may react
// to, else 0" — one 32-byte gather of posc[j] = {x, y, z, c_j} (packed once per evaluation by the style) instead of the
// position, the type, the type's parameters and the CV force.  G = 0: g = 1/2 (1 - tanh y) (bethe_density, sic Q13);
// G = 1: g = 1/2 (1 - tanh^2 y) / (0.1 r_th) (rleucg).  Same expressions in the same order as the general sweeps: with
// LPA = 8 the sums are identical bit for bit.  Pairs whose two CV forces are both zero skip the special functions.
struct CvBackArgs {
  const double4 *posc;
  int nlocal;
  const int *neigh;
  int stride;
  const int *numneigh;
  double cutsq, rth;
  double4 *frc;
  double *partials;
};
template <int LPA, int BS, int G>
__global__ void __launch_bounds__(BS) k_cv_back_fast(CvBackArgs p) {
  const int gid = (blockIdx.x * BS + threadIdx.x) / LPA;
  const int sub = threadIdx.x % LPA;
  const bool active = gid < p.nlocal;
  const int i = active ? gid : 0;
  const double4 ri = p.posc[i];
  const double ci = ri.w;
  const int jnum = active ? p.numneigh[i] : 0;
  const int *row = p.neigh + (size_t)i * p.stride;
  double fx = 0, fy = 0, fz = 0;
  double vir[6] = {0, 0, 0, 0, 0, 0};
  // two-deep software pipeline: the row index is fetched two entries ahead and the record one entry ahead, so neither
  // load waits for the other while this pair's special functions run (the scalar index loads held 22 % of the stall samples)
  int jj = sub;
  int j = -1;
  double4 rj = ri;
  if (jj < jnum) { j = row[rowslot(jj)] & UCG_NEIGHMASK; rj = p.posc[j]; }
  int jnext = (jj + LPA < jnum) ? (row[rowslot(jj + LPA)] & UCG_NEIGHMASK) : -1;
  while (j >= 0) {
    const int jn = jnext;
    double4 rn = rj;
    if (jn >= 0) rn = p.posc[jn];
    jj += LPA;
    jnext = (jj + LPA < jnum) ? (row[rowslot(jj + LPA)] & UCG_NEIGHMASK) : -1;
    const double dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
    const double rsq = rsq_exact(dx, dy, dz);
    if (rsq < p.cutsq && (ci != 0.0 || rj.w != 0.0)) {
      const double r = sqrt(rsq);
      const double t = tanh((r - p.rth) / (0.1 * p.rth));
      const double g = G == 0 ? 0.5 * (1.0 - t) : 0.5 * (1.0 - t * t) / (0.1 * p.rth);
      const double own = ci * g / r;        // i's loop
      const double oth = rj.w * g / r;      // what j's loop scatters to i
      const double fp = own + oth;
      fx += fp * dx; fy += fp * dy; fz += fp * dz;
      const double w = (j < p.nlocal ? 1.0 : 0.5) * own;   // ev_tally(i, j, nlocal, newton = 0, ...) in i's loop only
      vir[0] += w * dx * dx; vir[1] += w * dy * dy; vir[2] += w * dz * dz;
      vir[3] += w * dx * dy; vir[4] += w * dx * dz; vir[5] += w * dy * dz;
    }
    j = jn; rj = rn;
  }
  fx = group_sum<LPA>(fx); fy = group_sum<LPA>(fy); fz = group_sum<LPA>(fz);
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double v = group_sum<LPA>(vir[k]);
    if (active && sub == 0) ev[1 + k] = v;
  }
  if (active && sub == 0) {
    double4 f = p.frc[i];
    f.x += fx; f.y += fy; f.z += fz;
    p.frc[i] = f;
  }
  block_reduce_store<7, BS>(ev, p.partials);
}

}  // namespace ucg
