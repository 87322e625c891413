// PairTable_UCG_Bethe_Density on the GPU: the LAMMPS-facing half (deck grammar and error texts of
// UCG/pair_table_ucg_bethe_density.cpp:778-958, 1103-1153); the three sweeps run in
// csrc/pair_bethe_density.cu behind ucgb200_pair_bethe_density.
#include "pair_table_ucg_bethe_density.h"

#include "atom.h"
#include "error.h"
#include "fix.h"
#include "force.h"
#include "memory.h"
#include "modify.h"
#include "neighbor.h"
#include "ucg_device.h"

#include <cstdio>
#include <cstring>

using namespace LAMMPS_NS;

#define MAXLINE 1024

PairTable_UCG_Bethe_Density::PairTable_UCG_Bethe_Density(LAMMPS *lmp) : PairTable_UCGLD(lmp), density_applied(false) {
  no_virial_fdotr = 1;   // newton off: the style tallies its virial pair by pair
}

// same file grammar as the reference; the type/formal-type part becomes a ucgb200_statemap
void PairTable_UCG_Bethe_Density::read_state_settings(const char *file) {
  char line[MAXLINE], state_type[MAXLINE], entropy_spec[MAXLINE];
  FILE *fp = fopen(file, "r");
  if (fp == nullptr) error->one(FLERR, "Cannot open file {}", file);
  int max_states = 0;
  if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of RLEUCG state settings file");
  if (sscanf(line, "%d %d %d", &n_actual, &n_formal, &max_states) != 3 || n_actual < 1 || n_formal < n_actual)
    error->one(FLERR, "Invalid first line in UCG state settings file");
  std::vector<int> ns(n_actual + 1, 0), ff(2 * (n_actual + 1), 0);
  std::vector<double> mu(n_formal + 1, 0.0);
  use_density.assign(n_actual + 1, 0);
  use_state_entropy.assign(n_actual + 1, 0);
  cv_thresholds.assign(n_actual + 1, 0.0);
  threshold_radii.assign(n_actual + 1, 0.0);
  for (int i = 1; i <= n_actual; i++) {
    if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of UCG state settings file");
    int this_type = 0;
    sscanf(line, "%d %d", &this_type, &ns[i]);
    if (ns[i] < 1 || ns[i] > 2)
      error->one(FLERR, "Invalid number of states for atom type {}: {}. Only 1 or 2 states are allowed.", i, ns[i]);
    if (this_type != i)
      error->one(FLERR, "Please write orderly. Invalid atom type {} in UCG state settings file. Expected {}.", this_type, i);
    if (ns[i] == 2) {
      if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of UCG state settings file");
      char *p1 = strtok(line, " \t\r\n");
      for (int j = 0; j < 2; j++) {
        if (p1 == nullptr) error->one(FLERR, "Not enough formal types specified for atom type {}.", i);
        ff[2 * i + j] = atoi(p1);
        if (ff[2 * i + j] < 1 || ff[2 * i + j] > n_formal) error->one(FLERR, "Formal type out of range for atom type {}.", i);
        p1 = strtok(nullptr, " \t\r\n");
      }
      if (p1 == nullptr) error->one(FLERR, "Missing state type for atom type {}.", i);
      strcpy(state_type, p1);
      p1 = strtok(nullptr, " \t\r\n");
      if (p1 == nullptr) error->one(FLERR, "Missing entropy specification for atom type {}.", i);
      strcpy(entropy_spec, p1);
      if (strcmp(entropy_spec, "entropy") == 0) use_state_entropy[i] = 1;
      else if (strcmp(entropy_spec, "no_entropy") == 0) use_state_entropy[i] = 0;
      else error->one(FLERR, "Unknown entropy specification: {}. Use 'entropy' or 'no_entropy'.", entropy_spec);
      if (strcmp(state_type, "density") == 0) {
        use_density[i] = 1;
        if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of RLEUCG state settings file");
        sscanf(line, "%lg %lg", &cv_thresholds[i], &threshold_radii[i]);
      }
      if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of UCG state settings file");
      char *p2 = strtok(line, " \t\r\n");
      for (int j = 0; j < 2; j++) {
        if (p2 == nullptr) error->one(FLERR, "Not enough chemical potentials for atom type {}.", i);
        mu[ff[2 * i + j]] = atof(p2);
        p2 = strtok(nullptr, " \t\r\n");
      }
    }
  }
  fclose(fp);
  char err[512] = "";
  if (smap) ucgb200_host_statemap_free(smap);
  smap = nullptr;
  if (ucgb200_host_statemap_create(n_actual, n_formal, ns.data(), ff.data(), mu.data(), &smap, err, sizeof(err)))
    error->one(FLERR, "{}", err);
}

void PairTable_UCG_Bethe_Density::settings(int narg, char **arg) {
  if (!atom->ucg_flag) error->all(FLERR, "This pair style requires atom style ucg.");
  if (narg < 2) utils::missing_cmd_args(FLERR, "pair_style table", error);
  if (strcmp(arg[0], "lookup") == 0) tabstyle = LOOKUP;
  else if (strcmp(arg[0], "linear") == 0) tabstyle = LINEAR;
  else if (strcmp(arg[0], "spline") == 0) tabstyle = SPLINE;
  else if (strcmp(arg[0], "bitmap") == 0) tabstyle = BITMAP;
  else error->all(FLERR, "Unknown table style in pair_style command: {}", arg[0]);
  tablength = utils::inumeric(FLERR, arg[1], false, lmp);
  if (tablength < 2) error->all(FLERR, "Illegal number of pair table entries: {}", tablength);
  if (narg < 3) utils::missing_cmd_args(FLERR, "pair_style table_ucg_bethe_density", error);
  for (int iarg = 3; iarg < narg; iarg++) {
    if (strcmp(arg[iarg], "ewald") == 0) ewaldflag = 1;
    else if (strcmp(arg[iarg], "pppm") == 0) pppmflag = 1;
    else if (strcmp(arg[iarg], "msm") == 0) msmflag = 1;
    else if (strcmp(arg[iarg], "dispersion") == 0) dispersionflag = 1;
    else if (strcmp(arg[iarg], "tip4p") == 0) tip4pflag = 1;
    else error->all(FLERR, "Unknown pair_style table keyword: {}", arg[iarg]);
  }
  read_state_settings(arg[2]);
  for (auto t : tables) ucgb200_host_table_free(t);
  tables.clear();
  tabcut.clear();
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
  allocated = 0;
  maps_applied = false;
  density_applied = false;
  dev = UCGDevice::get(lmp);
  dev->static_uploaded = false;   // a new pair_style: every per-site array is sent again
  dev->list_ready = false;
  dev->check(lmp, ucgb200_tables_clear(dev->ctx), "tables_clear");
}

void PairTable_UCG_Bethe_Density::init_style() {
  // no neighbor->add_request(): the device builds and owns the (full) list, following the same skin rule; LAMMPS
  // keeps its ghost exchange (cutghost = cutforce + skin) but spends nothing on a host-side pair list
  double *pT = nullptr;
  int pdim;
  kT_found = false;
  for (int ifix = 0; ifix < modify->nfix; ifix++) {
    pT = (double *) modify->fix[ifix]->extract("t_target", pdim);
    if (pT) { T = *pT; kT_found = true; break; }
  }
  if (!kT_found) error->all(FLERR, "pair_style table_ucg_bethe_density requires a fix that exports t_target");
  kT = force->boltz * T;
  if (force->newton_pair != 0)
    error->all(FLERR, "Newton pair is turned on. It has to be turned off in local density UCG simulation.");
  maps_applied = false;
  density_applied = false;
}

void PairTable_UCG_Bethe_Density::configure_device() {
  if (!maps_applied) apply_maps();
  if (!density_applied) {
    dev->check(lmp, ucgb200_pair_bethe_density_configure(dev->ctx, n_actual, use_density.data(), use_state_entropy.data(),
                                                         cv_thresholds.data(), threshold_radii.data()),
               "pair_bethe_density_configure");
    density_applied = true;
  }
}

bool PairTable_UCG_Bethe_Density::ucg_deck(ucgb200_deck &deck) {
  configure_device();
  deck.pair_style = 3;
  return true;
}

void PairTable_UCG_Bethe_Density::compute(int eflag, int vflag) {
  ev_init(eflag, vflag);
  configure_device();
  const int nlocal = atom->nlocal;
  dev->upload(lmp, UCGB200_F_X | UCGB200_F_UCGL | UCGB200_F_UCGSTATE | UCGB200_F_UCGP);
  dev->ensure_list(lmp);
  dev->check(lmp, ucgb200_pair_bethe_density(dev->ctx, 1 | (eflag_atom ? 2 : 0), 1 | (vflag_atom ? 4 : 0)), "pair_bethe_density");
  std::vector<double> f(3 * (size_t)nlocal);
  ucgb200_atoms h{};
  h.f = f.data(); h.ucgp = atom->ucgp;   // the style publishes its posterior in atom->ucgp (:689)
  dev->check(lmp, ucgb200_atoms_download(dev->ctx, nlocal, &h, UCGB200_F_F | UCGB200_F_UCGP), "atoms_download");
  int code;
  if ((code = ucgb200_status_peek(dev->ctx, nullptr, nullptr, nullptr, nullptr))) dev->check(lmp, code, "pair_bethe_density");
  double **fh = atom->f;
  for (int i = 0; i < nlocal; i++) { fh[i][0] += f[3 * i]; fh[i][1] += f[3 * i + 1]; fh[i][2] += f[3 * i + 2]; }
  if (eflag_atom || vflag_atom) {   // [stock] ev_tally's per-atom halves (compute pe/atom, stress/atom)
    std::vector<double> ea(eflag_atom ? (size_t)nlocal : 0), va(vflag_atom ? 6 * (size_t)nlocal : 0);
    dev->check(lmp, ucgb200_pair_peratom(dev->ctx, nlocal, eflag_atom ? ea.data() : nullptr, vflag_atom ? va.data() : nullptr), "pair_peratom");
    if (eflag_atom) for (int i = 0; i < nlocal; i++) eatom[i] += ea[i];
    if (vflag_atom) for (int i = 0; i < nlocal; i++) for (int k = 0; k < 6; k++) vatom[i][k] += va[6 * (size_t)i + k];
  }
  double e, v[6];
  dev->check(lmp, ucgb200_pair_energy_virial(dev->ctx, &e, v), "pair_energy_virial");
  if (eflag_global) eng_vdwl += e;
  for (int k = 0; k < 6; k++) {
#ifdef LAMMPS_UCG_SHIM
    virial_tally[k] = v[k];   // test-harness diagnostic, absent from stock Pair
#endif
    if (vflag_global) virial[k] += v[k];
  }
}
