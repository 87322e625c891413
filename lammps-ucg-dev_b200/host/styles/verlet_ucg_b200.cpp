// run_style ucg/b200: see verlet_ucg_b200.h.  Mirrors [stock] Verlet::init/setup/run/cleanup and the call order of
// SURVEY.md section 3.1; the per-step work is ucgb200_run_between (csrc/run.cu).
#include "verlet_ucg_b200.h"

#include <cstring>

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "fix.h"
#include "force.h"
#include "modify.h"
#include "neighbor.h"
#include "output.h"
#include "pair.h"
#include "update.h"

#include "ucg_device.h"

using namespace LAMMPS_NS;

VerletUCGB200::VerletUCGB200(LAMMPS *lmp, int narg, char **arg) : Integrate(lmp, narg, arg), dev(nullptr) {
  memset(&deck, 0, sizeof deck);
}

void VerletUCGB200::collect_deck() {
  memset(&deck, 0, sizeof deck);
  parts.clear();
  auto *pair = dynamic_cast<UCGDeckPart *>(force->pair);
  if (!pair || !pair->ucg_deck(deck))
    error->all(FLERR, "run_style ucg/b200 needs one of the UCG pair styles; use run_style verlet");
  // the post_force stages act in fix definition order ([stock] Modify::post_force); the device loop is told which
  int order = 0, nstage = 0;
  bool have[4] = {false, false, false, false};
  for (int i = 0; i < modify->nfix; i++) {
    Fix *f = modify->fix[i];
    auto *part = dynamic_cast<UCGDeckPart *>(f);
    const ucgb200_deck before = deck;
    if (part && part->ucg_deck(deck)) {
      parts.push_back(part);
      const int stage = deck.langevin != before.langevin ? 1 : (deck.ucgstate != before.ucgstate ? 2 : (deck.nve == 2 && before.nve != 2 ? 3 : 0));
      if (stage && !have[stage]) { have[stage] = true; order = order * 10 + stage; nstage++; }
      continue;
    }
    if (part || modify->fmask[i])
      error->all(FLERR, "run_style ucg/b200: fix {} (style {}) has no part in the device loop; use run_style verlet", f->id, f->style);
  }
  if (comm->nprocs > 1) error->all(FLERR, "run_style ucg/b200 drives one context per process; multi-brick runs use the resident NCCL driver");
  for (int stage = 1; stage <= 3; stage++)
    if (!have[stage]) order = order * 10 + stage;   // absent stages do nothing, any place will do
  deck.post_force_order = order;
  deck.thermo_every = output ? output->thermo_every : 0;
}

void VerletUCGB200::init() {
  Integrate::init();
  if (atom->rmass) error->all(FLERR, "run_style ucg/b200: per-atom masses (rmass) are not supported");
  if (neighbor->every != 1 || neighbor->delay != 0 || !neighbor->dist_check)
    error->all(FLERR, "run_style ucg/b200 follows neigh_modify delay 0 every 1 check yes");
  dev = UCGDevice::get(lmp);
}

void VerletUCGB200::push() {
  dev->sync_globals(lmp);
  dev->static_uploaded = false;   // everything, including type / mask / tag / molecule / ucgml
  dev->upload(lmp, UCGB200_F_X | UCGB200_F_V | UCGB200_F_UCGL | UCGB200_F_UCGVL | UCGB200_F_UCGSTATE | UCGB200_F_UCGP);
  dev->check(lmp, ucgb200_set_ntimestep(dev->ctx, update->ntimestep), "set_ntimestep");
}

void VerletUCGB200::pull(bool thermo) {
  dev->download(lmp, UCGB200_F_X | UCGB200_F_V | UCGB200_F_F | UCGB200_F_UCGL | UCGB200_F_UCGVL | UCGB200_F_UCGSTATE | UCGB200_F_UCGP |
                         UCGB200_F_UCGFORCE | UCGB200_F_SCORES | UCGB200_F_NUMSTATES | UCGB200_F_TYPE);
  dev->list_ready = false;   // the offload-mode classes rebuild their list if they are used again
  if (thermo) {
    double th[16];
    dev->check(lmp, ucgb200_thermo(dev->ctx, th), "thermo");
    force->pair->eng_vdwl = th[0];
    for (int k = 0; k < 6; k++) force->pair->virial[k] = th[1 + k];
  }
}

// Verlet::setup: pbc, neighbor build, force_clear, pair->compute, fix setup() — ucgb200_setup
void VerletUCGB200::setup(int) {
  collect_deck();
  push();
  dev->check(lmp, ucgb200_deck_configure(dev->ctx, &deck), "deck_configure");
  dev->check(lmp, ucgb200_setup(dev->ctx), "setup");
  pull(true);
  if (output) output->setup();
}
void VerletUCGB200::setup_minimal(int flag) { setup(flag); }

// Verlet::run: the steps between two output steps never touch the host
void VerletUCGB200::run(int n) {
  const bigint last = update->ntimestep + n;
  while (update->ntimestep < last) {
    bigint next = last;
    if (output && output->next > update->ntimestep && output->next < next) next = output->next;
    for (UCGDeckPart *p : parts) {
      const long long stop = p->ucg_next_stop();
      if (stop > update->ntimestep && stop < next) next = stop;
    }
    const int rc = ucgb200_run_between(dev->ctx, (int) (next - update->ntimestep), update->beginstep, update->endstep);
    dev->check(lmp, rc, "run");
    update->ntimestep = next;
    for (UCGDeckPart *p : parts) if (p->ucg_next_stop() == next) p->ucg_after_step(next);
    if (output && output->next == next) {
      pull(true);
      output->write(next);
    }
  }
  pull(deck.thermo_every > 0 && update->ntimestep % deck.thermo_every == 0);
}

void VerletUCGB200::cleanup() {}
