// fixes.cu — the streaming per-site kernels of the UCG fixes:
//   FixNVE_UCGLD            UCG/fix_nve_ucgld.cpp:44-153
//   FixNVE_UCGLD_Wall_Hard  UCG/fix_nve_ucgld_wall_hard.cpp:62-257
//   FixUCGState             UCG/fix_ucgstate.cpp:88-132
//   Fix_UCGLD_Langevin      UCG/fix_ucgld_langevin.cpp:226-312
// One thread per owned site, 32-byte vector loads/stores of the {x,y,z,lambda},
// {v,vlambda}, {f,flambda} records; HBM-streaming bound.
#include "pair_common.cuh"

using namespace ucg;

namespace {

// v += dtf/m * f ; x += dtv * v ; vl += dtf/ml * fl ; l += dtv * vl      (fix_nve_ucgld.cpp:83-99)
// wall: ucgstate = (l < 0.5) ? 0 : 1                     (fix_nve_ucgld_wall_hard.cpp:125-131)
// PART: 0 the whole stage; 1 the {x, v} components only; 2 the {lambda, v_lambda} components (and the wall's state) only —
// the components are independent, so 1 followed by 2 is the stage bit for bit (ucgb200_step_host starts on {x, v} while
// lambda is still crossing PCIe)
template <bool WALL, int PART>
__global__ void k_nve_initial(double4 *__restrict__ pos, double4 *__restrict__ vel, const double4 *__restrict__ frc,
                              int *__restrict__ ts, const int *__restrict__ mask, const double *__restrict__ ucgml,
                              const TypeInfo *__restrict__ tinfo, int n, double dtv, double dtf, int groupbit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(mask[i] & groupbit)) return;
  double4 x = pos[i], v = vel[i];
  const double4 f = frc[i];
  int t = ts[i];
  if (PART != 2) {
    const double dtfm = dtf / tinfo[t & 0xffff].mass;
    v.x += dtfm * f.x; v.y += dtfm * f.y; v.z += dtfm * f.z;
    x.x += dtv * v.x; x.y += dtv * v.y; x.z += dtv * v.z;
  }
  if (PART != 1) {
    const double dtflm = dtf / ucgml[i];
    v.w += dtflm * f.w;
    x.w += dtv * v.w;
  }
  pos[i] = x; vel[i] = v;
  if (WALL && PART != 1) ts[i] = (t & 0xffff) | ((x.w < 0.5 ? 0 : 1) << 16);
}

// v += dtf/m * f ; vl += dtf/ml * fl                                   (fix_nve_ucgld.cpp:139-151)
// wall: l<0 -> -l, l>1 -> 2-l, vl -> -vl                 (fix_nve_ucgld_wall_hard.cpp:194-200)
template <bool WALL>
__global__ void k_nve_final(double4 *__restrict__ pos, double4 *__restrict__ vel, const double4 *__restrict__ frc,
                            const int *__restrict__ ts, const int *__restrict__ mask, const double *__restrict__ ucgml,
                            const TypeInfo *__restrict__ tinfo, int n, double dtf, int groupbit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(mask[i] & groupbit)) return;
  double4 v = vel[i];
  const double4 f = frc[i];
  const double dtfm = dtf / tinfo[ts[i] & 0xffff].mass;
  v.x += dtfm * f.x; v.y += dtfm * f.y; v.z += dtfm * f.z;
  const double dtflm = dtf / ucgml[i];
  v.w += dtflm * f.w;
  if (WALL) {
    double l = pos[i].w;
    if (l < 0.0) { pos[i].w = -l; v.w = -v.w; }
    else if (l > 1.0) { pos[i].w = 2.0 - l; v.w = -v.w; }
  }
  vel[i] = v;
}

// fl += (-7980 x^9 + 2x) * 10 * H, x = l - 1/2           (fix_nve_ucgld_wall_hard.cpp:234-257)
__global__ void k_wall_bias(const double4 *__restrict__ pos, double4 *__restrict__ frc, const int *__restrict__ mask,
                            int n, double H, int groupbit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(mask[i] & groupbit)) return;
  double x = pos[i].w - 0.5;
  frc[i].w += (-7980 * x * x * x * x * x * x * x * x * x + 2 * x) * 10 * H;
}

// FixUCGState::post_force (fix_ucgstate.cpp:99-131); the loop ignores the fix group (Q19).
// mode 0 deterministic, 1 ld, 2 mc.  The mc rule is reproduced literally (Q18).
__global__ void k_ucgstate(double4 *__restrict__ pos, int *__restrict__ ts, double *__restrict__ ucgp,
                           const double2 *__restrict__ scores, const int *__restrict__ tag,
                           const TypeInfo *__restrict__ tinfo, int n, int mode, unsigned seed, double rate,
                           unsigned long long step) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int t = ts[i];
  int state = (t >> 16) & 1;
  double p;
  if (tinfo[t & 0xffff].nstates == 1) {
    if (mode != 1) state = 0;
    p = 1.0;
  } else {
    double2 s = scores[i];
    double e0 = exp(fmin(s.x, 700.0)), e1 = exp(fmin(s.y, 700.0));
    p = fmin(1.0 - 1e-6, fmax(1e-6, e1 / (e0 + e1)));
    if (mode == 2) {
      double fac = state == 0 ? p / (1.0 - p) : (1.0 - p) / p;
      fac = fmin(fac, 1.0) * rate;
      double r = philox_uniform(seed, 0x55434753u /* "UCGS" */, (unsigned)tag[i], step);
      state = (r < fac) ? 0 : 1;
    } else if (mode == 0) {
      state = (int)round(p);  // half away from zero, like std::round (:125)
    }
  }
  ucgp[i] = p;
  if (mode != 1) {
    ts[i] = (t & 0xffff) | (state << 16);
    pos[i].w = p;  // ucgl = ucgp (:130)
  }
}

// fl += gamma1*vl + gamma2*(U-1/2)        (fix_ucgld_langevin.cpp:273-296, Tp_BIAS variants)
__global__ void k_langevin(const double4 *__restrict__ vel, double4 *__restrict__ frc, const int *__restrict__ ts,
                           const int *__restrict__ mask, const int *__restrict__ tag, const double *__restrict__ gfac,
                           int ntypes, int n, double tsqrt, unsigned seed, unsigned long long step, int groupbit,
                           int zero_v_skip) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(mask[i] & groupbit)) return;
  int type = ts[i] & 0xffff;
  double gamma1 = gfac[type];
  double gamma2 = gfac[ntypes + 1 + type] * tsqrt;
  double vl = vel[i].w;
  double fran = gamma2 * (philox_uniform(seed, 0x4c414e47u /* "LANG" */, (unsigned)tag[i], step) - 0.5);
  if (zero_v_skip && vl == 0.0) fran = 0.0;  // Tp_BIAS branch (:285)
  frc[i].w += gamma1 * vl + fran;
}

// ---- resident run loop: every per-site fix stage between two pair evaluations in ONE pass
// post_force fixes in the deck's definition order (default ucgld/langevin, ucgstate, wall bias) -> final_integrate of
// this step -> [initial_integrate of the next step -> Neighbor::check_distance].  The arithmetic of
// each stage is that of the single-stage kernels above, applied in the same order to the same
// values, so trajectories are identical; the site record is read and written once (~270 B/site)
// instead of five times.
struct TailArgs {
  double4 *pos, *vel, *frc;
  const double4 *xhold;
  int *ts;
  const int *mask, *tag;
  const double *ucgml, *gfac;
  const double2 *scores;
  double *ucgp;
  const TypeInfo *tinfo;
  int n, ntypes;
  // langevin
  int langevin, lgroupbit, lbias;
  double tsqrt;
  unsigned lseed;
  // ucgstate
  int ucgstate_mode;   // -1 absent, 0 deterministic, 1 ld, 2 mc
  unsigned useed;
  double urate;
  // integrator
  int groupbit, bias;
  double barrier, dtv, dtf;
  int order;           // post_force stages, one per byte, first stage in the lowest byte: 1 langevin, 2 ucgstate, 3 wall bias
  int fuse_next;       // also do the next step's initial_integrate + check_distance
  double triggersq;
  int *flags;
  unsigned long long *maxdisp;
  unsigned long long step;
};

template <bool WALL>
__global__ void __launch_bounds__(256) k_step_tail(TailArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int t = a.ts[i];
  const int type = t & 0xffff;
  const int m = a.mask[i];
  double4 x = a.pos[i], v = a.vel[i], f = a.frc[i];
  bool fdirty = false, xdirty = false;
  const bool ingroup = (m & a.groupbit) != 0;
  // post_force stages in the deck's fix definition order (TailArgs::order, [stock] Modify::post_force)
#pragma unroll
  for (int stage = 0; stage < 3; stage++) {
    const int which = (a.order >> (8 * stage)) & 0xff;
    if (which == 1) {
      if (a.langevin && (m & a.lgroupbit)) {                              // k_langevin
        const double gamma1 = a.gfac[type];
        const double gamma2 = a.gfac[a.ntypes + 1 + type] * a.tsqrt;
        double fran = gamma2 * (philox_uniform(a.lseed, 0x4c414e47u, (unsigned)a.tag[i], a.step) - 0.5);
        if (a.lbias && v.w == 0.0) fran = 0.0;                            // Tp_BIAS branch (:285)
        f.w += gamma1 * v.w + fran;
        fdirty = true;
      }
    } else if (which == 2) {
      if (a.ucgstate_mode >= 0) {                                         // k_ucgstate
        int state = (t >> 16) & 1;
        double p;
        if (a.tinfo[type].nstates == 1) {
          if (a.ucgstate_mode != 1) state = 0;
          p = 1.0;
        } else {
          const double2 s = a.scores[i];
          const double e0 = exp(fmin(s.x, 700.0)), e1 = exp(fmin(s.y, 700.0));
          p = fmin(1.0 - 1e-6, fmax(1e-6, e1 / (e0 + e1)));
          if (a.ucgstate_mode == 2) {
            double fac = state == 0 ? p / (1.0 - p) : (1.0 - p) / p;
            fac = fmin(fac, 1.0) * a.urate;
            const double r = philox_uniform(a.useed, 0x55434753u, (unsigned)a.tag[i], a.step);
            state = (r < fac) ? 0 : 1;
          } else if (a.ucgstate_mode == 0) {
            state = (int)round(p);
          }
        }
        a.ucgp[i] = p;
        if (a.ucgstate_mode != 1) {
          t = type | (state << 16);
          x.w = p;
          xdirty = true;
        }
      }
    } else {
      if (WALL && a.bias && ingroup) {                                    // k_wall_bias
        const double y = x.w - 0.5;
        f.w += (-7980 * y * y * y * y * y * y * y * y * y + 2 * y) * 10 * a.barrier;
        fdirty = true;
      }
    }
  }
  double dtfm = 0.0, dtflm = 0.0;
  if (ingroup) {                                                          // k_nve_final
    dtfm = a.dtf / a.tinfo[type].mass;
    dtflm = a.dtf / a.ucgml[i];
    v.x += dtfm * f.x; v.y += dtfm * f.y; v.z += dtfm * f.z;
    v.w += dtflm * f.w;
    if (WALL) {
      if (x.w < 0.0) { x.w = -x.w; v.w = -v.w; xdirty = true; }
      else if (x.w > 1.0) { x.w = 2.0 - x.w; v.w = -v.w; xdirty = true; }
    }
  }
  if (a.fuse_next) {
    if (ingroup) {                                                        // k_nve_initial of step + 1
      v.x += dtfm * f.x; v.y += dtfm * f.y; v.z += dtfm * f.z;
      x.x += a.dtv * v.x; x.y += a.dtv * v.y; x.z += a.dtv * v.z;
      v.w += dtflm * f.w;
      x.w += a.dtv * v.w;
      xdirty = true;
      if (WALL) t = type | ((x.w < 0.5 ? 0 : 1) << 16);
    }
    const double4 h = a.xhold[i];                                         // k_check_distance
    const double moved = rsq_exact(x.x - h.x, x.y - h.y, x.z - h.z);
    if (moved > a.triggersq) a.flags[0] = 1;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(moved);
    if (bits > *a.maxdisp) atomicMax(a.maxdisp, bits);
  }
  if (ingroup) a.vel[i] = v;
  if (fdirty) a.frc[i] = f;
  if (xdirty) a.pos[i] = x;
  a.ts[i] = t;
}

// sums: [0] sum 0.5*ml*vl^2*mvv2e, [1] sum 0.5*m*|v|^2*mvv2e, [2] count in group
template <int BS>
__global__ void __launch_bounds__(BS) k_kinetic(const double4 *__restrict__ vel, const int *__restrict__ ts,
                                                const int *__restrict__ mask, const double *__restrict__ ucgml,
                                                const TypeInfo *__restrict__ tinfo, int n, double mvv2e, int groupbit,
                                                double *__restrict__ partials) {
  double acc[3] = {0, 0, 0};
  for (int i = blockIdx.x * BS + threadIdx.x; i < n; i += gridDim.x * BS) {
    if (mask[i] & groupbit) {
      double4 v = vel[i];
      acc[0] += 0.5 * ucgml[i] * v.w * v.w * mvv2e;
      acc[1] += 0.5 * tinfo[ts[i] & 0xffff].mass * (v.x * v.x + v.y * v.y + v.z * v.z) * mvv2e;
      acc[2] += 1.0;
    }
  }
  block_reduce_store<3, BS>(acc, partials);
}

}  // namespace

#define GRID1(n) nblocks((n), 256), 256, 0, c->stream

int ucg_nve_initial_part(ucgb200_ctx *c, double dtv, double dtf, int groupbit, int wall, int part) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  c->maxdisp_valid = false;   // sites move: the displacement bound is stale until the next check_distance
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (c->nlocal == 0) return 0;
#define NVE_INITIAL(W, P) k_nve_initial<W, P><<<GRID1(c->nlocal)>>>(c->pos.p, c->vel.p, c->frc.p, c->ts.p, c->mask.p, c->ucgml.p, \
                                                                   c->d_typeinfo.p, c->nlocal, dtv, dtf, groupbit)
  if (wall) { if (part == 0) NVE_INITIAL(true, 0); else if (part == 1) NVE_INITIAL(true, 1); else NVE_INITIAL(true, 2); }
  else { if (part == 0) NVE_INITIAL(false, 0); else if (part == 1) NVE_INITIAL(false, 1); else NVE_INITIAL(false, 2); }
#undef NVE_INITIAL
  UCG_LAUNCHED(c);
  return 0;
}

extern "C" int ucgb200_fix_nve_initial(ucgb200_ctx *c, double dtv, double dtf, int groupbit, int wall) {
  return ucg_nve_initial_part(c, dtv, dtf, groupbit, wall, 0);
}

extern "C" int ucgb200_fix_nve_final(ucgb200_ctx *c, double dtf, int groupbit, int wall) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (c->nlocal == 0) return 0;
  if (wall)
    k_nve_final<true><<<GRID1(c->nlocal)>>>(c->pos.p, c->vel.p, c->frc.p, c->ts.p, c->mask.p, c->ucgml.p,
                                            c->d_typeinfo.p, c->nlocal, dtf, groupbit);
  else
    k_nve_final<false><<<GRID1(c->nlocal)>>>(c->pos.p, c->vel.p, c->frc.p, c->ts.p, c->mask.p, c->ucgml.p,
                                             c->d_typeinfo.p, c->nlocal, dtf, groupbit);
  UCG_LAUNCHED(c);
  return 0;
}

extern "C" int ucgb200_fix_wall_bias(ucgb200_ctx *c, double barrier, int groupbit) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  if (c->nlocal == 0) return 0;
  k_wall_bias<<<GRID1(c->nlocal)>>>(c->pos.p, c->frc.p, c->mask.p, c->nlocal, barrier, groupbit);
  UCG_LAUNCHED(c);
  return 0;
}

extern "C" int ucgb200_fix_ucgstate(ucgb200_ctx *c, int mode, int seed, double rate, long long step) {
  if (!c) return -1;
  if (mode < 0 || mode > 2) return fail(c, "Unknown argument for fix ucgstate");
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (c->nlocal == 0) return 0;
  k_ucgstate<<<GRID1(c->nlocal)>>>(c->pos.p, c->ts.p, c->ucgp.p, c->scores.p, c->tag.p, c->d_typeinfo.p, c->nlocal,
                                   mode, (unsigned)seed, rate, (unsigned long long)step);
  UCG_LAUNCHED(c);
  return 0;
}

extern "C" int ucgb200_fix_langevin(ucgb200_ctx *c, const double *gfactor1, const double *gfactor2, int ntypes,
                                    double tsqrt, int seed, long long step, int groupbit, int zero_v_skip) {
  if (!c || !gfactor1 || !gfactor2 || ntypes < 1) return -1;
  cudaSetDevice(c->device);
  if (c->nlocal == 0) return 0;
  std::vector<double> g(2 * (ntypes + 1));
  for (int t = 0; t <= ntypes; t++) { g[t] = gfactor1[t]; g[ntypes + 1 + t] = gfactor2[t]; }
  if (g != c->gfactor1) {
    UCG_CHECK(c, c->d_gfac.ensure(g.size()));
    UCG_CHECK(c, cudaMemcpyAsync(c->d_gfac.p, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    UCG_CHECK(c, cudaStreamSynchronize(c->stream));
    c->gfactor1 = g;
  }
  k_langevin<<<GRID1(c->nlocal)>>>(c->vel.p, c->frc.p, c->ts.p, c->mask.p, c->tag.p, c->d_gfac.p, ntypes, c->nlocal,
                                   tsqrt, (unsigned)seed, (unsigned long long)step, groupbit, zero_v_skip);
  UCG_LAUNCHED(c);
  return 0;
}

static int kinetic(ucgb200_ctx *c, int groupbit, double out[3]) {
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  out[0] = out[1] = out[2] = 0.0;
  if (c->nlocal == 0) return 0;
  constexpr int BS = 256;
  int nblk = std::min(nblocks(c->nlocal, BS), 148 * 8);
  UCG_CHECK(c, c->d_partials.ensure((size_t)nblk * 4 + 64));
  k_kinetic<BS><<<nblk, BS, 0, c->stream>>>(c->vel.p, c->ts.p, c->mask.p, c->ucgml.p, c->d_typeinfo.p, c->nlocal,
                                           c->mvv2e, groupbit, c->d_partials.p);
  UCG_LAUNCHED(c);
  if ((rc = reduce_partials(c, nblk, 3, 8))) return rc;
  UCG_CHECK(c, cudaMemcpyAsync(out, c->d_ev.p + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int ucgb200_lambda_ke(ucgb200_ctx *c, int groupbit, double *ke_sum, long long *count) {
  if (!c) return -1;
  double o[3];
  int rc = kinetic(c, groupbit, o);
  if (rc) return rc;
  if (ke_sum) *ke_sum = o[0];
  if (count) *count = (long long)(o[2] + 0.5);
  return 0;
}
extern "C" int ucgb200_kinetic_energy(ucgb200_ctx *c, int groupbit, double *ke_sum, long long *count) {
  if (!c) return -1;
  double o[3];
  int rc = kinetic(c, groupbit, o);
  if (rc) return rc;
  if (ke_sum) *ke_sum = o[1];
  if (count) *count = (long long)(o[2] + 0.5);
  return 0;
}

int ucg_post_force_order(const ucgb200_deck &d, int order[3]);   // run.cu
// Fused tail of a resident step (see k_step_tail).  gfac layout as uploaded by ucgb200_fix_langevin.
int ucg_step_tail(ucgb200_ctx *c, const ucgb200_deck &d, double tsqrt, int fuse_next) {
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (c->nlocal == 0) return 0;
  if (d.langevin) {
    const int nt = c->n_formal;
    std::vector<double> g(2 * (nt + 1));
    for (int t = 0; t <= nt; t++) { g[t] = c->lang_g1[t]; g[nt + 1 + t] = c->lang_g2[t]; }
    if (g != c->gfactor1) {
      UCG_CHECK(c, c->d_gfac.ensure(g.size()));
      UCG_CHECK(c, cudaMemcpyAsync(c->d_gfac.p, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      UCG_CHECK(c, cudaStreamSynchronize(c->stream));
      c->gfactor1 = g;
    }
  }
  TailArgs a{};
  a.pos = c->pos.p; a.vel = c->vel.p; a.frc = c->frc.p; a.xhold = c->xhold.p; a.ts = c->ts.p;
  a.mask = c->mask.p; a.tag = c->tag.p; a.ucgml = c->ucgml.p; a.gfac = c->d_gfac.p; a.scores = c->scores.p;
  a.ucgp = c->ucgp.p; a.tinfo = c->d_typeinfo.p; a.n = c->nlocal; a.ntypes = c->n_formal;
  a.langevin = d.langevin; a.lgroupbit = d.langevin_groupbit ? d.langevin_groupbit : 1; a.tsqrt = tsqrt; a.lbias = d.langevin_bias;
  a.lseed = (unsigned)d.langevin_seed;
  a.ucgstate_mode = d.ucgstate == 0 ? -1 : (d.ucgstate == 1 ? 0 : (d.ucgstate == 2 ? 1 : 2));
  a.useed = (unsigned)d.ucgstate_seed; a.urate = d.ucgstate_rate;
  a.groupbit = d.nve_groupbit ? d.nve_groupbit : 1; a.bias = d.wall_bias; a.barrier = d.wall_barrier;
  a.dtv = c->dt; a.dtf = 0.5 * c->dt * c->ftm2v;
  a.fuse_next = fuse_next; a.triggersq = 0.25 * c->skin * c->skin; a.flags = c->d_flags.p; a.maxdisp = c->d_maxdisp.p;
  a.step = (unsigned long long)c->ntimestep;
  {
    int ord[3];
    if (ucg_post_force_order(d, ord)) return fail(c, "deck: post_force_order must be a permutation of the digits 1 2 3");
    a.order = ord[0] | (ord[1] << 8) | (ord[2] << 16);
  }
  if (fuse_next) {
    UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p, 0, sizeof(int), c->stream));
    UCG_CHECK(c, cudaMemsetAsync(c->d_maxdisp.p, 0, sizeof(unsigned long long), c->stream));
  }
  c->maxdisp_valid = false;   // sites move (or not): valid again after the decide that follows
  if (d.nve == 2) k_step_tail<true><<<GRID1(c->nlocal)>>>(a);
  else k_step_tail<false><<<GRID1(c->nlocal)>>>(a);
  UCG_LAUNCHED(c);
  return 0;
}
