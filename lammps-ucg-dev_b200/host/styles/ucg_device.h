// ucg_device.h — glue between the LAMMPS-facing style classes and the C-ABI (ucgb200_*).
// One device context per LAMMPS instance.  In this "offload" mode an unmodified stock Verlet loop drives the styles;
// the device keeps its own cell-sorted copy of the AtomVecUCG arrays, ghosts and FULL neighbor list between calls.
// Which arrays cross PCIe is decided per field:
//  * eager (any deck): every style call uploads the arrays it reads and downloads the ones it writes — always right,
//    ~25 transfers per step;
//  * tracked (decks whose every force and fix style is a UCG style, inside `run`): a field is uploaded only when the
//    host copy is newer than the device's, and a field the device has written stays there until something on the host
//    needs it — positions after initial_integrate (stock Verlet's Neighbor::decide and Comm::forward_comm read them
//    every step), everything before an exchange (Fix::pre_exchange), on output steps (Fix::end_of_step) and when the
//    run ends (Fix::post_run).  An all-UCG deck then moves x out once per step and nothing in.
// The resident mode (run_style ucg/b200: ucgb200_deck_configure + ucgb200_run) avoids even that.
#ifndef LMP_UCG_DEVICE_H
#define LMP_UCG_DEVICE_H

#include <cstdlib>
#include <map>
#include <string>
#include <vector>

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
  // false: this style (or this set of its options) cannot run inside the device loop
  virtual bool ucg_deck(ucgb200_deck &deck) = 0;
  // the next step after which the device loop must hand control back to this style, and what it does then
  // true: inside a stock Verlet run this style reads and writes the per-atom arrays through UCGDevice::upload /
  // download only, so the tracked offload mode may leave them on the device between its calls
  virtual bool ucg_tracked_ok() const { return false; }
  virtual long long ucg_next_stop() const { return -1; }
  virtual void ucg_after_step(long long) {}
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
    it->second->unpin_all();
    if (it->second->scratch_p) ucgb200_pinned_free(it->second->scratch_p);
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
    int code = 0, ti = 0, tj = 0; double rsq = 0.0;
    if (rc > 0) ucgb200_status(ctx, &code, &ti, &tj, &rsq);   // one read fetches the pair and clears the sticky word
    if (rc == UCGB200_ERR_TABLE_INNER || rc == UCGB200_ERR_TABLE_OUTER) {
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

  // Offload mode moves the LAMMPS per-atom arrays themselves; page-locking them in place makes every transfer a DMA.
  // LAMMPS re-allocates per-atom arrays only when atom->nmax grows (AtomVec::grow), so the registrations are redone
  // whenever nmax or any of the pointers changed since the last call — before the arrays are touched again.
  // UCGB200_PIN_HOST=0 leaves the arrays pageable.
  struct Pinned { void *p; size_t bytes; };
  std::vector<Pinned> pinned;
  std::vector<const void *> pinned_key;
  int pinned_nmax = -1;
  void unpin_all() {
    for (auto &r : pinned) ucgb200_host_unregister(r.p);
    pinned.clear();
    pinned_key.clear();
    pinned_nmax = -1;
  }
  void pin_arrays(LAMMPS *lmp) {
    static const bool enabled = !(getenv("UCGB200_PIN_HOST") && atoi(getenv("UCGB200_PIN_HOST")) == 0);
    if (!enabled) return;
    Atom *a = lmp->atom;
    const size_t n = (size_t) a->nmax;
    const size_t D = sizeof(double), I = sizeof(int);
    const Pinned want[] = {{a->x ? &a->x[0][0] : nullptr, 3 * n * D}, {a->v ? &a->v[0][0] : nullptr, 3 * n * D},
                           {a->f ? &a->f[0][0] : nullptr, 3 * n * D}, {a->type, n * I}, {a->mask, n * I}, {a->tag, n * I},
                           {a->molecule, n * I}, {a->ucgstate, n * I}, {a->num_ucgstates, n * I}, {a->ucgl, n * D},
                           {a->ucgvl, n * D}, {a->ucgml, n * D}, {a->ucgp, n * D}, {a->ucgforce, n * D},
                           {a->ucgsoftmaxscores ? &a->ucgsoftmaxscores[0][0] : nullptr, 2 * n * D}};
    bool same = pinned_nmax == a->nmax && pinned_key.size() == sizeof(want) / sizeof(want[0]);
    for (size_t k = 0; same && k < pinned_key.size(); k++) same = pinned_key[k] == want[k].p;
    if (same) return;
    unpin_all();
    if (n == 0) return;
    for (const Pinned &w : want) {
      pinned_key.push_back(w.p);
      if (w.p && w.bytes && ucgb200_host_register(w.p, w.bytes) == 0) pinned.push_back(w);
    }
    pinned_nmax = a->nmax;
  }
  // page-locked scratch for results that are added to the host arrays (pair compute)
  void *scratch_p = nullptr;
  size_t scratch_bytes = 0;
  void *scratch(size_t bytes) {
    if (bytes <= scratch_bytes) return scratch_p;
    if (scratch_p) ucgb200_pinned_free(scratch_p);
    scratch_p = nullptr;
    scratch_bytes = 0;
    void *q = nullptr;
    if (ucgb200_pinned_alloc(bytes + bytes / 8, &q) || !q) return nullptr;
    scratch_p = q;
    scratch_bytes = bytes + bytes / 8;
    return scratch_p;
  }

  // ---- tracked mode (see the header of this file)
  bool tracked = false, tracking_pending = false;
  unsigned dev_valid = 0;    // fields whose device copy is what the host holds or newer: no upload needed
  unsigned dev_newer = 0;    // fields the device has written and the host has not seen yet
  long long bytes_up = 0, bytes_down = 0;   // per-site payload moved so far (diagnostics: ucg_traffic())
  static constexpr unsigned EAGER_OUT = UCGB200_F_X;   // what stock Verlet reads on the host every step
  static unsigned bytes_per_site(unsigned fields) {
    unsigned b = 0;
    if (fields & UCGB200_F_X) b += 24; if (fields & UCGB200_F_V) b += 24; if (fields & UCGB200_F_F) b += 24;
    if (fields & UCGB200_F_SCORES) b += 16;
    for (unsigned f : {UCGB200_F_UCGL, UCGB200_F_UCGVL, UCGB200_F_UCGML, UCGB200_F_UCGP, UCGB200_F_UCGFORCE}) if (fields & f) b += 8;
    for (unsigned f : {UCGB200_F_TYPE, UCGB200_F_MASK, UCGB200_F_TAG, UCGB200_F_MOLECULE, UCGB200_F_UCGSTATE, UCGB200_F_NUMSTATES}) if (fields & f) b += 4;
    return b;
  }
  // every force and fix style of the deck is a UCG style, one process, velocity Verlet: nothing else reads or writes
  // the per-atom arrays between two of our calls except stock Verlet itself (force_clear, decide, forward_comm)
  bool tracking_allowed(LAMMPS *lmp);
  void tracking_begin(LAMMPS *lmp) {
    tracked = tracking_allowed(lmp);
    dev_valid = dev_newer = 0;       // the host is the truth at the start of a run
  }
  void flush(LAMMPS *lmp) {          // the host is about to read (or reorder) the arrays
    if (!tracked || !dev_newer) return;
    const unsigned f = dev_newer;
    dev_newer = 0;
    raw_download(lmp, f);
  }
  void host_changed() { dev_valid = 0; static_uploaded = false; list_ready = false; }   // after a flush: host reordered / rewrote
  void tracking_end(LAMMPS *lmp) { flush(lmp); tracked = false; dev_valid = dev_newer = 0; }

  void raw_download(LAMMPS *lmp, unsigned fields) {
    Atom *a = lmp->atom;
    pin_arrays(lmp);
    ucgb200_atoms h = view(a);
    check(lmp, ucgb200_atoms_download(ctx, a->nlocal, &h, fields), "atoms_download");
    bytes_down += (long long) bytes_per_site(fields) * a->nlocal;
  }
  // a style is about to READ `fields` on the device
  void upload(LAMMPS *lmp, unsigned fields) {
    Atom *a = lmp->atom;
    pin_arrays(lmp);
    if (a->nlocal != nlocal_dev) { static_uploaded = false; list_ready = false; dev_valid = 0; dev_newer = 0; }
    if (tracked) fields &= ~dev_valid;
    if (!static_uploaded) fields |= (UCGB200_F_TYPE | UCGB200_F_MASK | UCGB200_F_TAG | UCGB200_F_MOLECULE | UCGB200_F_UCGML |
                                     UCGB200_F_X | UCGB200_F_V | UCGB200_F_UCGL | UCGB200_F_UCGVL | UCGB200_F_UCGSTATE | UCGB200_F_UCGP) &
                                    ~(tracked ? dev_newer : 0u);
    if (fields) {
      ucgb200_atoms h = view(a);
      check(lmp, ucgb200_atoms_upload(ctx, a->nlocal, &h, fields), "atoms_upload");
      bytes_up += (long long) bytes_per_site(fields) * a->nlocal;
    }
    nlocal_dev = a->nlocal;
    static_uploaded = true;
    if (tracked) dev_valid |= fields;
  }
  // a style has WRITTEN `fields` on the device
  void download(LAMMPS *lmp, unsigned fields) {
    if (tracked) {
      dev_valid |= fields;
      dev_newer |= fields & ~EAGER_OUT;
      fields &= EAGER_OUT;
      if (!fields) return;
    }
    raw_download(lmp, fields);
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
