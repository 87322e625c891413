// ucg_device.h — glue between the LAMMPS-facing style classes and the C-ABI (ucgb200_*).
// One device context per LAMMPS instance.  In this "offload" mode every style call uploads
// the AtomVecUCG arrays it reads and downloads the ones it writes, so an unmodified stock
// Verlet loop (which touches atom->x / atom->f directly) stays correct; the device keeps its
// own cell-sorted copy, ghosts and FULL neighbor list between calls.  The resident mode
// (ucgb200_deck_configure + ucgb200_run) avoids these transfers altogether.
#ifndef LMP_UCG_DEVICE_H
#define LMP_UCG_DEVICE_H

#include <map>
#include <string>

#include "atom.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "update.h"

#include "ucgb200.h"
#include "ucgb200_host.h"

namespace LAMMPS_NS {

// what a UCG style contributes to the resident integrator (run_style ucg/b200): each class writes its own
// part of the ucgb200_deck the device runs (include/ucgb200.h "resident run")
class UCGDeckPart {
 public:
  virtual ~UCGDeckPart() = default;
  virtual void ucg_deck(ucgb200_deck &deck) const = 0;
};

class UCGDevice {
 public:
  ucgb200_ctx *ctx = nullptr;
  int nlocal_dev = -1;
  bool static_uploaded = false;   // type, mask, tag, molecule, ucgml
  bool list_ready = false;

  static std::map<LAMMPS *, UCGDevice *> &instances() {
    static std::map<LAMMPS *, UCGDevice *> inst;
    return inst;
  }
  // to be called when the LAMMPS instance goes away (LAMMPS::destroy, or a driver's teardown): destroys the
  // context, so that a later instance at the same address starts from a fresh device
  static void drop(LAMMPS *lmp) {
    auto &inst = instances();
    auto it = inst.find(lmp);
    if (it == inst.end()) return;
    if (it->second->ctx) ucgb200_destroy(it->second->ctx);
    delete it->second;
    inst.erase(it);
  }
  static UCGDevice *get(LAMMPS *lmp) {
    auto &inst = instances();
    auto it = inst.find(lmp);
    if (it != inst.end()) return it->second;
    auto *d = new UCGDevice();
    int rc = ucgb200_create(0, &d->ctx);
    if (rc) lmp->error->all(FLERR, "ucg-b200: cannot create a CUDA context (rc {}); there is no CPU fallback", rc);
    inst[lmp] = d;
    return d;
  }

  // error->one / error->all with the reference's message texts
  void check(LAMMPS *lmp, int rc, const char *where) {
    if (rc == 0) return;
    if (rc == UCGB200_ERR_TABLE_INNER || rc == UCGB200_ERR_TABLE_OUTER) {
      int code, ti, tj; double rsq;
      ucgb200_status(ctx, &code, &ti, &tj, &rsq);
      lmp->error->one(FLERR, rc == UCGB200_ERR_TABLE_INNER ? "Pair distance < table inner cutoff: atoms {} {} dist {}"
                                                            : "Pair distance > table outer cutoff: atoms {} {} dist {}",
                      ti, tj, sqrt(rsq));
    }
    lmp->error->one(FLERR, "ucg-b200 {}: {} (rc {})", where, ucgb200_last_error(ctx), rc);
  }

  void sync_globals(LAMMPS *lmp) {
    Force *f = lmp->force;
    Domain *d = lmp->domain;
    check(lmp, ucgb200_set_units(ctx, f->boltz, f->ftm2v, f->mvv2e), "set_units");
    int per[3] = {d->xperiodic, d->yperiodic, d->zperiodic};
    check(lmp, ucgb200_set_box(ctx, d->boxlo, d->boxhi, per), "set_box");
    check(lmp, ucgb200_set_timestep(ctx, lmp->update->dt), "set_timestep");
    check(lmp, ucgb200_set_special_lj(ctx, f->special_lj), "set_special_lj");
  }

  ucgb200_atoms view(Atom *a) {
    ucgb200_atoms h{};
    h.x = a->x ? &a->x[0][0] : nullptr; h.v = a->v ? &a->v[0][0] : nullptr; h.f = a->f ? &a->f[0][0] : nullptr;
    h.type = a->type; h.mask = a->mask; h.tag = a->tag; h.molecule = a->molecule;
    h.ucgstate = a->ucgstate; h.ucgl = a->ucgl; h.ucgvl = a->ucgvl; h.ucgml = a->ucgml; h.ucgp = a->ucgp;
    h.ucgforce = a->ucgforce; h.ucgsoftmaxscores = a->ucgsoftmaxscores ? &a->ucgsoftmaxscores[0][0] : nullptr;
    h.num_ucgstates = a->num_ucgstates;
    return h;
  }

  void upload(LAMMPS *lmp, unsigned fields) {
    Atom *a = lmp->atom;
    if (a->nlocal != nlocal_dev) { static_uploaded = false; list_ready = false; }
    if (!static_uploaded) fields |= UCGB200_F_TYPE | UCGB200_F_MASK | UCGB200_F_TAG | UCGB200_F_MOLECULE | UCGB200_F_UCGML |
                                    UCGB200_F_X | UCGB200_F_V | UCGB200_F_UCGL | UCGB200_F_UCGVL | UCGB200_F_UCGSTATE | UCGB200_F_UCGP;
    ucgb200_atoms h = view(a);
    check(lmp, ucgb200_atoms_upload(ctx, a->nlocal, &h, fields), "atoms_upload");
    nlocal_dev = a->nlocal;
    static_uploaded = true;
  }
  void download(LAMMPS *lmp, unsigned fields) {
    Atom *a = lmp->atom;
    ucgb200_atoms h = view(a);
    check(lmp, ucgb200_atoms_download(ctx, a->nlocal, &h, fields), "atoms_download");
  }
  // the device list follows the same skin rule as Neighbor::decide
  void ensure_list(LAMMPS *lmp) {
    int flag = 1;
    if (list_ready) check(lmp, ucgb200_neigh_decide(ctx, &flag), "neigh_decide");
    if (flag) { check(lmp, ucgb200_neigh_build(ctx), "neigh_build"); list_ready = true; }
    else check(lmp, ucgb200_ghosts_forward(ctx), "ghosts_forward");
  }
};

}  // namespace LAMMPS_NS
#endif
