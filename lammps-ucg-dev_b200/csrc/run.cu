// run.cu — device-resident time stepping: the [stock] Verlet::setup / Verlet::run call
// order with every stage a kernel on the context stream.  The only
// host<->device traffic per step is the 4-byte rebuild flag (Neighbor::decide needs a
// host decision) and, on thermo steps, the 7-double energy/virial read-back.
#include "ucg_internal.cuh"

#include <cmath>

using namespace ucg;

static_assert(sizeof(ucgb200_deck) == 128, "ucgb200_deck is part of the C-ABI: new fields come out of reserved[]");
extern "C" int ucgb200_deck_configure(ucgb200_ctx *c, const ucgb200_deck *deck) {
  if (!c || !deck) return -1;
  if (deck->pair_style < 0 || deck->pair_style > 3) return fail(c, "unknown pair style");
  if (deck->langevin) {
    // ctor checks of Fix_UCGLD_Langevin (fix_ucgld_langevin.cpp:80-81)
    if (deck->t_period <= 0.0) return fail(c, "Fix langevin period must be > 0.0");
    if (deck->langevin_seed <= 0) return fail(c, "Illegal fix langevin command");
  }
  if (deck->ucgstate) {
    // FixUCGState::setup (fix_ucgstate.cpp:148-156): a fix exporting t_target must exist;
    // inside the package that is fix ucgld/langevin, otherwise the host layer supplies kT.
  }
  c->deck = *deck;
  c->deck_set = true;
  return 0;
}

// Fix_UCGLD_Langevin::init gfactors (fix_ucgld_langevin.cpp:164-171).  The reference
// indexes atom->ucgml with the TYPE index (quirk Q20); decks use a uniform lambda mass,
// for which this equals the per-site value.  ml_of_type[t] must be supplied by the caller
// of the low-level entry; the resident path reads site (t) like the reference does.
static int langevin_factors(ucgb200_ctx *c, std::vector<double> &g1, std::vector<double> &g2) {
  int nt = c->n_formal;
  g1.assign(nt + 1, 0.0);
  g2.assign(nt + 1, 0.0);
  std::vector<double> ml(nt + 1, 0.0);
  int nread = std::min(nt + 1, c->nlocal);
  if (nread > 0) {
    // ucgml[i] for i = 1..ntypes in the ORIGINAL host order == what the reference reads
    std::vector<double> all(c->nlocal);
    std::vector<int> orig(c->nlocal);
    UCG_CHECK(c, cudaMemcpyAsync(all.data(), c->ucgml.p, c->nlocal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    UCG_CHECK(c, cudaMemcpyAsync(orig.data(), c->orig.p, c->nlocal * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    UCG_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int s = 0; s < c->nlocal; s++)
      if (orig[s] <= nt) ml[orig[s]] = all[s];
  }
  for (int i = 1; i <= nt; i++) {
    double m = (i < c->nlocal) ? ml[i] : ml[0];
    g1[i] = -m / c->deck.t_period / c->ftm2v;
    g2[i] = std::sqrt(m) / c->ftm2v;
    g2[i] *= std::sqrt(24.0 * c->boltz / c->deck.t_period / c->dt / c->mvv2e);
  }
  return 0;
}

static double current_t_target(const ucgb200_ctx *c) {
  // compute_target, CONSTANT style (fix_ucgld_langevin.cpp:323-331)
  double delta = (double)(c->ntimestep - c->beginstep);
  if (delta != 0.0) delta /= (double)(c->endstep - c->beginstep);
  return c->deck.t_start + delta * (c->deck.t_stop - c->deck.t_start);
}

static int pair_compute(ucgb200_ctx *c, int ev) {
  const ucgb200_deck &d = c->deck;
  switch (d.pair_style) {
    case 0: return ucgb200_pair_ucgld(c, ev, ev);
    case 1: return ucgb200_pair_bethe(c, ev, ev, d.bethe_method, d.bethe_pseudo, d.bethe_prior, d.bethe_noise_level, d.bethe_seed ? d.bethe_seed : 1);
    case 2: return ucgb200_pair_rleucg(c, ev, ev);
    default: return ucgb200_pair_bethe_density(c, ev, ev);
  }
}

int ucg_step_tail(ucgb200_ctx *c, const ucgb200_deck &d, double tsqrt, int fuse_next);
int ucg_neigh_decide_prechecked(ucgb200_ctx *c, int *rebuild);
// comm.cu: the same three operations across bricks (NCCL)
int ucg_mb_rebuild(ucgb200_ctx *c);
int ucg_mb_forward(ucgb200_ctx *c);
int ucg_mb_decide(ucgb200_ctx *c, bool prechecked, int *rebuild);
int ucg_mb_forward_reduce(ucgb200_ctx *c, bool with_decide);   // forward halo + MAX of the bricks' rebuild flags / displacement bounds
int ucg_mb_forward_begin(ucgb200_ctx *c);   // peer-mapped transport: push + local images ...
int ucg_mb_forward_end(ucgb200_ctx *c);     // ... wait + flag reduction + unpack
int ucg_mb_p2p_active(ucgb200_ctx *c);
int ucg_pair_can_split(ucgb200_ctx *c);     // pair_ucgld.cu

static int do_build(ucgb200_ctx *c) { return c->halo.nranks > 1 ? ucg_mb_rebuild(c) : ucgb200_neigh_build(c); }
static int do_forward(ucgb200_ctx *c) { return c->halo.nranks > 1 ? ucg_mb_forward(c) : ucgb200_ghosts_forward(c); }
static int do_decide(ucgb200_ctx *c, bool prechecked, int *flag) {
  if (c->halo.nranks > 1) return ucg_mb_decide(c, prechecked, flag);
  return prechecked ? ucg_neigh_decide_prechecked(c, flag) : ucgb200_neigh_decide(c, flag);
}

// the three post_force stages in the deck's fix definition order (ucgb200_deck::post_force_order; 0 = 123)
int ucg_post_force_order(const ucgb200_deck &d, int order[3]) {
  int code = d.post_force_order ? d.post_force_order : 123;
  bool seen[4] = {false, false, false, false};
  for (int k = 2; k >= 0; k--) {
    const int s = code % 10;
    code /= 10;
    if (s < 1 || s > 3 || seen[s]) return -1;
    seen[s] = true;
    order[k] = s;
  }
  return code == 0 ? 0 : -1;
}

static int post_force(ucgb200_ctx *c, bool at_setup) {
  const ucgb200_deck &d = c->deck;
  int rc, order[3];
  if (ucg_post_force_order(d, order)) return fail(c, "deck: post_force_order must be a permutation of the digits 1 2 3");
  // fixes act in definition order ([stock] Modify::post_force): 1 ucgld/langevin, 2 ucgstate, 3 the wall bias of the
  // integrator fix.  The thermostat must precede ucgstate (fix_ucgstate.cpp:143-156) for t_target, not for post_force.
  for (int k = 0; k < 3; k++) {
    if (order[k] == 1 && d.langevin) {
      double tt = current_t_target(c);
      if (c->lang_g1.empty()) { if ((rc = langevin_factors(c, c->lang_g1, c->lang_g2))) return rc; }
      if ((rc = ucgb200_fix_langevin(c, c->lang_g1.data(), c->lang_g2.data(), c->n_formal, std::sqrt(tt),
                                     d.langevin_seed, c->ntimestep, d.langevin_groupbit ? d.langevin_groupbit : 1, d.langevin_bias)))
        return rc;
    }
    if (order[k] == 2 && d.ucgstate) {
      int mode = d.ucgstate == 1 ? 0 : (d.ucgstate == 2 ? 1 : 2);
      if ((rc = ucgb200_fix_ucgstate(c, mode, d.ucgstate_seed, d.ucgstate_rate, c->ntimestep))) return rc;
    }
    // [stock] Fix::setup() is a no-op for the wall fix, so no bias at step 0
    if (order[k] == 3 && !at_setup && d.nve == 2 && d.wall_bias) {
      if ((rc = ucgb200_fix_wall_bias(c, d.wall_barrier, d.nve_groupbit ? d.nve_groupbit : 1))) return rc;
    }
  }
  return 0;
}

// The pair kernels report "Pair distance < table inner cutoff" & co. (error->one in the reference,
// pair_table_ucgld.cpp:223-230) through the sticky device word and go on.  The resident loop reads the code with the
// rebuild flag that crosses to the host every step anyway (same pinned read-back, no extra synchronisation) and stops
// the run with the UCGB200_ERR_* code; ucgb200_status then yields the pair.
static int queue_error_readback(ucgb200_ctx *c) {
  UCG_CHECK(c, cudaMemcpyAsync(c->h_flags + 8, &c->d_err.p->code, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
static int device_error(ucgb200_ctx *c) {
  const int code = c->h_flags[8];
  if (!code) return 0;
  c->err = code == UCGB200_ERR_TABLE_INNER ? "Pair distance < table inner cutoff"
           : code == UCGB200_ERR_TABLE_OUTER ? "Pair distance > table outer cutoff"
           : code == UCGB200_ERR_DENSITY_TYPE ? "Declared type in RLEUCG does not exist." : "device error word set";
  return code;
}

extern "C" int ucgb200_setup(ucgb200_ctx *c) {
  if (!c) return -1;
  if (!c->deck_set) return fail(c, "deck not configured");
  cudaSetDevice(c->device);
  int rc;
  if (c->deck.langevin) {
    // pair styles and fix ucgstate take kT from the first fix exporting t_target
    c->kT = c->boltz * c->deck.t_start;
  }
  c->beginstep = c->endstep = c->ntimestep;
  if ((rc = do_build(c))) return rc;
  c->nbuilds = 0;
  if ((rc = pair_compute(c, 1))) return rc;
  c->lang_g1.clear();
  c->lang_g2.clear();
  if ((rc = post_force(c, true))) return rc;
  if ((rc = queue_error_readback(c))) return rc;
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return device_error(c);
}

extern "C" int ucgb200_run(ucgb200_ctx *c, int nsteps) {
  if (!c || nsteps < 0) return -1;
  return ucgb200_run_between(c, nsteps, c->ntimestep, c->ntimestep + nsteps);
}

// `run N start S stop E` [stock Run::command]: N steps of a run whose ramps (fix ucgld/langevin's target
// temperature, fix_ucgld_langevin.cpp:318-353) span S..E — lets a caller cut a run into pieces (dump steps) without
// changing a single bit of it
extern "C" int ucgb200_run_between(ucgb200_ctx *c, int nsteps, long long beginstep, long long endstep) {
  if (!c || nsteps < 0) return -1;
  if (!c->deck_set) return fail(c, "deck not configured");
  if (beginstep > c->ntimestep || endstep < c->ntimestep + nsteps) return fail(c, "run_between: the steps lie outside start..stop");
  cudaSetDevice(c->device);
  const ucgb200_deck &d = c->deck;
  const int gb = d.nve_groupbit ? d.nve_groupbit : 1;
  const double dtv = c->dt, dtf = 0.5 * c->dt * c->ftm2v;  // FixNVE_UCGLD::init (fix_nve_ucgld.cpp:35-41)
  c->beginstep = beginstep;
  c->endstep = endstep;
  int rc;
  // the per-site fix stages between two pair evaluations run as one fused kernel (fixes.cu,
  // k_step_tail) unless UCGB200_FUSED_TAIL=0; it needs an integrator fix and a single brick loop
  const bool fused = d.nve && !(getenv("UCGB200_FUSED_TAIL") && atoi(getenv("UCGB200_FUSED_TAIL")) == 0);
  bool pre_integrated = false;   // initial_integrate + check_distance of this step already done by the last tail
  // speculative launch of the next pair evaluation (fused tail); UCGB200_SPECULATE=0 turns it off.  Across bricks the
  // decision rests on the all-reduced displacement bound, which is the same number on every brick, so all bricks
  // speculate (or not) together.
  const bool speculative = fused && !(getenv("UCGB200_SPECULATE") && atoi(getenv("UCGB200_SPECULATE")) == 0);
  const bool bricks = c->halo.nranks > 1;
  c->last_maxdisp = -1.0;
  for (int n = 0; n < nsteps; n++) {
    c->ntimestep++;
    const int ev = d.thermo_every > 0 && (c->ntimestep % d.thermo_every == 0);
    if (!pre_integrated && !(n == 0 && c->skip_initial_once)) {   // ucgb200_step_host may have run it under its uploads
      StageTimer t(c, 3);
      if (d.nve) { if ((rc = ucgb200_fix_nve_initial(c, dtv, dtf, gb, d.nve == 2))) return rc; }
      t.stop();
    }
    c->skip_initial_once = false;
    int flag = 0;
    bool pair_in_flight = false, forward_done = false;
    const bool cluster_due = d.cluster_freq > 0 && c->cluster.set && c->cluster.next_reneighbor == c->ntimestep;
    if (speculative && pre_integrated && c->list_valid && !cluster_due && (!c->timers_on || c->timers_async)) {
      // Neighbor::decide without a pipeline bubble: the flag (and the largest squared displacement since the build) of
      // this step were computed by the previous step's fused tail.  Their read-back is queued, and — when the last known
      // displacement says a rebuild is still some steps away — so are the ghost refresh and the pair kernel of this
      // step, BEFORE the host waits for the flag: the device never idles while the host decides.  If the flag says
      // rebuild after all, the speculative results are simply overwritten by the rebuild + pair that follow.
      if (!c->ev_flag) UCG_CHECK(c, cudaEventCreateWithFlags(&c->ev_flag, cudaEventDisableTiming));
      const double quiet = 0.8 * 0.5 * c->skin;
      const bool calm = c->last_maxdisp >= 0.0 && c->last_maxdisp < quiet * quiet;
      // halo / compute overlap: the interior sites (no neighbor from another brick) are evaluated between the push of
      // this brick's records and the wait for the peers', the boundary sites after the unpack
      const bool split = bricks && calm && !ev && d.pair_style == 0 && ucg_mb_p2p_active(c) && ucg_pair_can_split(c);
      StageTimer tf(c, 2);
      if (split) {
        if ((rc = ucg_mb_forward_begin(c))) return rc;
        // interior rows only meet sites of this brick (and their local periodic images): the displacement bound of
        // THIS brick, left by the fused tail, is enough for the exact skin skipping of this part
        c->maxdisp_valid = true;
        c->pair_part = 0;
        if ((rc = pair_compute(c, 0))) return rc;
        if ((rc = ucg_mb_forward_end(c))) return rc;
        forward_done = true;
      } else if (bricks) {
        // the ghost refresh of this step carries every brick's flag and displacement bound with it (control words of
        // the peer-mapped push, or two small all-reduces behind the NCCL send/recv group): after it d_flags[0] and
        // d_maxdisp hold the MAX over all bricks.  A refresh that turns out to precede a rebuild is simply redone.
        if ((rc = ucg_mb_forward_reduce(c, true))) return rc;
        forward_done = true;
      }
      tf.stop();
      UCG_CHECK(c, cudaMemcpyAsync(c->h_flags, c->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      UCG_CHECK(c, cudaMemcpyAsync(c->h_flags + 6, c->d_maxdisp.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
      if ((rc = queue_error_readback(c))) return rc;
      UCG_CHECK(c, cudaEventRecord(c->ev_flag, c->stream));
      c->maxdisp_valid = true;
      if (calm) {
        if (!forward_done) {
          StageTimer t2(c, 2);
          if ((rc = do_forward(c))) return rc;
          t2.stop();
        }
        if (split) c->pair_part = 1;
        StageTimer t0(c, 0);
        if ((rc = pair_compute(c, ev))) return rc;
        t0.stop();
        pair_in_flight = true;
      }
      UCG_CHECK(c, cudaEventSynchronize(c->ev_flag));
      flag = c->h_flags[0] ? 1 : 0;
      unsigned long long bits;
      memcpy(&bits, c->h_flags + 6, sizeof bits);
      memcpy(&c->last_maxdisp, &bits, sizeof bits);
      if (flag) pair_in_flight = false;   // discarded
    } else {
      StageTimer t(c, 1);
      if ((rc = queue_error_readback(c))) return rc;     // rides on the synchronisation of the decision below
      if ((rc = do_decide(c, pre_integrated, &flag))) return rc;
      t.stop();
      c->last_maxdisp = -1.0;   // unknown until a fused tail has reported it
    }
    if ((rc = device_error(c))) return rc;   // set by a pair evaluation of an earlier step (or of setup)
    // fix cluster_switch: force_reneighbor at next_reneighbor; pre_exchange() rebuilds, labels the
    // clusters and switches types (fix_cluster_switch.cpp:464-481), then Verlet rebuilds again
    if (cluster_due) {
      StageTimer t(c, 3);
      if ((rc = do_build(c))) return rc;
      if ((rc = ucgb200_cluster_check(c, nullptr))) return rc;
      if ((rc = ucgb200_cluster_switch(c, nullptr, nullptr))) return rc;
      c->cluster.next_reneighbor = c->ntimestep + d.cluster_freq;
      flag = 1;
      t.stop();
    }
    if (flag) {
      if ((rc = do_build(c))) return rc;
      c->last_maxdisp = 0.0;
      c->host_out_done &= ~UCGB200_F_X;   // positions that left before the decision (ucgb200_step_host) leave again, wrapped
    }
    else if (!pair_in_flight && !forward_done) {
      StageTimer t(c, 2);
      if ((rc = do_forward(c))) return rc;
      t.stop();
    }
    if (c->host_out && !pair_in_flight) {
      // ucgb200_step_host: positions are final for this step once the list is valid; lambda and the state too unless a
      // later stage writes them (the wall's reflection, a non-ld fix ucgstate): their copy runs under the pair kernel
      unsigned early = UCGB200_F_X;
      if (d.nve != 2 && (d.ucgstate == 0 || d.ucgstate == 2)) early |= UCGB200_F_UCGL | UCGB200_F_UCGSTATE;
      if ((rc = ucg_host_out_queue(c, early))) return rc;
    }
    if (!pair_in_flight) {
      StageTimer t(c, 0);
      if ((rc = pair_compute(c, ev))) return rc;
      t.stop();
    }
    if (c->host_out && (rc = ucg_host_out_queue(c, UCGB200_F_F))) return rc;   // f is final after the pair kernel
    {
      StageTimer t(c, 3);
      if (fused) {
        if (d.langevin && c->lang_g1.empty()) { if ((rc = langevin_factors(c, c->lang_g1, c->lang_g2))) return rc; }
        const int fuse_next = (n + 1 < nsteps) ? 1 : 0;
        if ((rc = ucg_step_tail(c, d, d.langevin ? std::sqrt(current_t_target(c)) : 0.0, fuse_next))) return rc;
        pre_integrated = fuse_next != 0;
      } else {
        if ((rc = post_force(c, false))) return rc;
        if (d.nve) { if ((rc = ucgb200_fix_nve_final(c, dtf, gb, d.nve == 2))) return rc; }
      }
      t.stop();
    }
    if (ev) {
      double e, v[6];
      if ((rc = ucgb200_pair_energy_virial(c, &e, v))) return rc;
      c->thermo[0] = e;
      for (int k = 0; k < 6; k++) c->thermo[1 + k] = v[k];
    }
  }
  // the last pair evaluation(s) of this piece: one read-back per run_between call
  if (nsteps > 0) {
    if ((rc = queue_error_readback(c))) return rc;
    UCG_CHECK(c, cudaStreamSynchronize(c->stream));
    if ((rc = device_error(c))) return rc;
  }
  return 0;
}

extern "C" int ucgb200_set_ntimestep(ucgb200_ctx *c, long long ntimestep) {
  if (!c || ntimestep < 0) return -1;
  c->ntimestep = ntimestep;
  return 0;
}

extern "C" int ucgb200_thermo(ucgb200_ctx *c, double out[16]) {
  if (!c || !out) return -1;
  cudaSetDevice(c->device);
  int rc;
  double ke = 0, lke = 0;
  long long cnt = 0;
  if (c->nlocal) {
    if ((rc = ucgb200_kinetic_energy(c, 1, &ke, &cnt))) return rc;
    int lgb = c->deck.langevin_groupbit ? c->deck.langevin_groupbit : 1;
    if ((rc = ucgb200_lambda_ke(c, lgb, &lke, &cnt))) return rc;
  }
  if (c->ev_valid) {
    double e, v[6];
    if ((rc = ucgb200_pair_energy_virial(c, &e, v))) return rc;
    c->thermo[0] = e;
    for (int k = 0; k < 6; k++) c->thermo[1 + k] = v[k];
  }
  for (int k = 0; k < 16; k++) out[k] = 0.0;
  for (int k = 0; k < 7; k++) out[k] = c->thermo[k];
  out[7] = ke;
  out[8] = lke;
  // Fix_UCGLD_Langevin::end_of_step (:303-312): lambda temperature, divided by nlocal
  out[9] = c->nlocal ? lke / (0.5 * c->boltz * c->nlocal) : 0.0;
  out[10] = (double)c->ntimestep;
  out[11] = (double)c->nbuilds;
  out[12] = (double)c->nlocal;
  out[13] = (double)c->nghost;
  return 0;
}

extern "C" int ucgb200_timers(ucgb200_ctx *c, int enable, double out_ms[4], long long out_launches[4]) {
  if (!c) return -1;
  stage_resolve_all(c);
  if (out_ms) for (int k = 0; k < 4; k++) out_ms[k] = c->t_ms[k];
  if (out_launches) for (int k = 0; k < 4; k++) out_launches[k] = c->t_launch[k];
  if (enable >= 0) {
    if (enable != (int)c->timers_on || enable >= 2) {
      for (int k = 0; k < 4; k++) { c->t_ms[k] = 0; c->t_launch[k] = 0; }
    }
    c->timers_on = enable != 0;
    c->timers_async = enable == 3;
  }
  return 0;
}
