// PairTable_RLEUCG_INTERFACE on the GPU: the LAMMPS-facing half.  Deck grammar and error texts follow
// UCG/pair_table_rleucg_interface.cpp (settings :526-573, read_state_settings :575-664, coeff :670-752,
// init_style :759-797, init_one :803-810); the three neighbor sweeps run in csrc/pair_rleucg.cu.
#include "pair_table_rleucg_interface.h"

#include "atom.h"
#include "comm.h"
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

PairTable_RLEUCG_INTERFACE::PairTable_RLEUCG_INTERFACE(LAMMPS *lmp)
    : Pair(lmp), tabstyle(LINEAR), tablength(0), T(0.0), kT(0.0), n_actual_types(0), n_total_states(0), tabindex(nullptr),
      configured(false), dev(nullptr) {
  no_virial_fdotr = 1;
  restartinfo = 1;
}

PairTable_RLEUCG_INTERFACE::~PairTable_RLEUCG_INTERFACE() {
  if (copymode) return;
  for (auto t : tables) ucgb200_host_table_free(t);
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
    memory->destroy(tabindex);
  }
}

void PairTable_RLEUCG_INTERFACE::allocate() {
  allocated = 1;
  const int nt = atom->ntypes + 1;
  memory->create(setflag, nt, nt, "pair:setflag");
  memory->create(cutsq, nt, nt, "pair:cutsq");
  memory->create(tabindex, nt, nt, "pair:tabindex");
  memset(&setflag[0][0], 0, nt * nt * sizeof(int));
  memset(&cutsq[0][0], 0, nt * nt * sizeof(double));
  memset(&tabindex[0][0], 0, nt * nt * sizeof(int));
}

void PairTable_RLEUCG_INTERFACE::settings(int narg, char **arg) {
  if (narg < 2) error->all(FLERR, "Illegal pair_style command");
  if (strcmp(arg[0], "lookup") == 0) tabstyle = LOOKUP;
  else if (strcmp(arg[0], "linear") == 0) tabstyle = LINEAR;
  else if (strcmp(arg[0], "spline") == 0) tabstyle = SPLINE;
  else if (strcmp(arg[0], "bitmap") == 0) tabstyle = BITMAP;
  else error->all(FLERR, "Unknown table style in pair_style command");
  tablength = utils::inumeric(FLERR, arg[1], false, lmp);
  if (tablength < 2) error->all(FLERR, "Illegal number of pair table entries");
  if (narg < 3) error->all(FLERR, "Illegal pair_style command");
  for (int iarg = 3; iarg < narg; iarg++) {
    if (strcmp(arg[iarg], "ewald") == 0) ewaldflag = 1;
    else if (strcmp(arg[iarg], "pppm") == 0) pppmflag = 1;
    else if (strcmp(arg[iarg], "msm") == 0) msmflag = 1;
    else if (strcmp(arg[iarg], "dispersion") == 0) dispersionflag = 1;
    else if (strcmp(arg[iarg], "tip4p") == 0) tip4pflag = 1;
    else error->all(FLERR, "Illegal pair_style command");
  }
  read_state_settings(arg[2]);
  for (auto t : tables) ucgb200_host_table_free(t);
  tables.clear();
  tabcut.clear();
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
    memory->destroy(tabindex);
  }
  allocated = 0;
  configured = false;
  dev = UCGDevice::get(lmp);
  dev->static_uploaded = false;   // a new pair_style: every per-site array is sent again
  dev->list_ready = false;
  dev->check(lmp, ucgb200_tables_clear(dev->ctx), "tables_clear");
}

void PairTable_RLEUCG_INTERFACE::read_state_settings(const char *file) {
  char line[MAXLINE], state_type[MAXLINE], entropy_spec[MAXLINE];
  FILE *fp = fopen(file, "r");
  if (fp == nullptr) error->one(FLERR, "Cannot open file {}", file);
  if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of RLEUCG state settings file");
  if (sscanf(line, "%d %d", &n_actual_types, &n_total_states) != 2 || n_actual_types < 1 || n_total_states < n_actual_types)
    error->one(FLERR, "Invalid first line in RLEUCG state settings file");
  n_states_per_type.assign(n_actual_types + 1, 0);
  use_state_entropy.assign(n_actual_types + 1, 0);
  cv_thresholds.assign(n_actual_types + 1, 0.0);
  threshold_radii.assign(n_actual_types + 1, 0.0);
  actual_types_from_state.assign(n_total_states + 1, 0);
  chemical_potentials.assign(n_total_states + 1, 0.0);
  int curr_state = 1;
  for (int i = 1; i <= n_actual_types; i++) {
    if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of RLEUCG state settings file");
    state_type[0] = entropy_spec[0] = 0;
    sscanf(line, "%d %s %s", &n_states_per_type[i], state_type, entropy_spec);
    if (n_states_per_type[i] < 1 || n_states_per_type[i] > 2)
      error->one(FLERR, "RLEUCG on the device supports 1 or 2 states per type");
    if (strcmp(entropy_spec, "use_entropy") == 0) use_state_entropy[i] = 1;
    else if (strcmp(entropy_spec, "no_entropy") == 0) use_state_entropy[i] = 0;
    if (n_states_per_type[i] > 1) {
      if (strcmp(state_type, "density") != 0) error->one(FLERR, "Unknown state assignment type for RLEUCG");
      if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of RLEUCG state settings file");
      sscanf(line, "%lg %lg", &cv_thresholds[i], &threshold_radii[i]);
      if (!fgets(line, MAXLINE, fp)) error->one(FLERR, "Unexpected end of RLEUCG state settings file");
      char *p = strtok(line, " \t\r\n");
      for (int j = 0; j < n_states_per_type[i] - 1; j++) {
        // (sic) indexed by actual type + substate, read back by state type + substate (:300 vs :651)
        if (p && i + j <= n_total_states) chemical_potentials[i + j] = atof(p);
        p = strtok(nullptr, " \t\r\n");
      }
    }
    for (int j = 0; j < n_states_per_type[i]; j++) {
      if (curr_state > n_total_states) error->one(FLERR, "More states than declared in RLEUCG state settings file");
      actual_types_from_state[curr_state++] = i;
    }
  }
  fclose(fp);
}

void PairTable_RLEUCG_INTERFACE::coeff(int narg, char **arg) {
  if (narg != 4 && narg != 5) error->all(FLERR, "Illegal pair_coeff command");
  if (!allocated) allocate();
  int ilo, ihi, jlo, jhi;
  utils::bounds(FLERR, arg[0], 1, atom->ntypes, ilo, ihi, error);
  utils::bounds(FLERR, arg[1], 1, atom->ntypes, jlo, jhi, error);
  // cut < 0: take the table's own upper end (rhi / last r of the file), as the reference does
  const double cut = narg == 5 ? utils::numeric(FLERR, arg[4], false, lmp) : -1.0;
  char err[512] = "";
  ucgb200_table *tb = nullptr;
  if (ucgb200_host_table_from_file(arg[2], arg[3], cut, tabstyle, tablength, &tb, err, sizeof(err))) error->all(FLERR, "{}", err);
  int index = -1;
  dev->check(lmp, ucgb200_host_table_upload(dev->ctx, tb, &index), "table_upload");
  if (index != (int) tables.size()) error->all(FLERR, "ucg-b200: table index mismatch");
  double info[8];
  int n = 0;
  ucgb200_host_table_info(tb, info, &n);
  tables.push_back(tb);
  tabcut.push_back(info[4]);
  int count = 0;
  for (int i = ilo; i <= ihi; i++)
    for (int j = MAX(jlo, i); j <= jhi; j++) {
      tabindex[i][j] = index;
      setflag[i][j] = 1;
      count++;
    }
  if (count == 0) error->all(FLERR, "Illegal pair_coeff command");
  configured = false;
}

void PairTable_RLEUCG_INTERFACE::init_style() {
  // no neighbor->add_request(): the device builds and owns the (full) list, following the same skin rule; LAMMPS
  // keeps its ghost exchange (cutghost = cutforce + skin) but spends nothing on a host-side pair list
  double *pT = nullptr;
  int pdim;
  bool found = false;
  for (int ifix = 0; ifix < modify->nfix; ifix++) {
    pT = (double *) modify->fix[ifix]->extract("t_target", pdim);
    if (pT) { T = *pT; found = true; break; }
  }
  if (!found) error->all(FLERR, "pair_style table_rleucg_interface requires a fix that exports t_target");
  kT = force->boltz * T;
  if (force->newton_pair != 0)
    error->all(FLERR, "Newton pair is turned on. It has to be turned off in local density UCG simulation.");
  configured = false;
}

double PairTable_RLEUCG_INTERFACE::init_one(int i, int j) {
  if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
  tabindex[j][i] = tabindex[i][j];
  return tabcut[tabindex[i][j]];
}

void PairTable_RLEUCG_INTERFACE::configure_device() {
  dev->sync_globals(lmp);
  const int nt = atom->ntypes;
  if (n_total_states != nt) error->all(FLERR, "RLEUCG state settings file declares {} states but the system has {} atom types", n_total_states, nt);
  std::vector<int> ti((nt + 1) * (nt + 1), 0);
  std::vector<double> cs((nt + 1) * (nt + 1), 0.0), mass(nt + 1, 1.0);
  for (int i = 1; i <= nt; i++) {
    mass[i] = atom->mass[i];
    for (int j = 1; j <= nt; j++) {
      ti[i * (nt + 1) + j] = tabindex[i][j];
      cs[i * (nt + 1) + j] = cutsq[i][j];
    }
  }
  dev->check(lmp, ucgb200_pair_rleucg_configure(dev->ctx, nt, actual_types_from_state.data(), n_actual_types,
                                                n_states_per_type.data(), use_state_entropy.data(), cv_thresholds.data(),
                                                threshold_radii.data(), chemical_potentials.data(), ti.data(), cs.data(),
                                                mass.data(), kT),
             "pair_rleucg_configure");
  dev->check(lmp, ucgb200_neigh_configure(dev->ctx, neighbor->skin, 0.0), "neigh_configure");
  dev->list_ready = false;
  configured = true;
}

bool PairTable_RLEUCG_INTERFACE::ucg_deck(ucgb200_deck &deck) {
  if (!configured) configure_device();
  deck.pair_style = 2;
  return true;
}

void PairTable_RLEUCG_INTERFACE::compute(int eflag, int vflag) {
  ev_init(eflag, vflag);
  if (!configured) configure_device();
  const int nlocal = atom->nlocal;
  dev->upload(lmp, UCGB200_F_X | UCGB200_F_UCGL | UCGB200_F_UCGSTATE | UCGB200_F_UCGP);
  dev->ensure_list(lmp);
  // the energy is always evaluated: the reference feeds a stale evdwl into the probability
  // force on steps without eflag
  dev->check(lmp, ucgb200_pair_rleucg(dev->ctx, 1 | (eflag_atom ? 2 : 0), 1 | (vflag_atom ? 4 : 0)), "pair_rleucg");
  std::vector<double> f(3 * (size_t) nlocal);
  ucgb200_atoms h{};
  h.f = f.data();
  dev->check(lmp, ucgb200_atoms_download(dev->ctx, nlocal, &h, UCGB200_F_F), "atoms_download");
  int code;
  if ((code = ucgb200_status_peek(dev->ctx, nullptr, nullptr, nullptr, nullptr))) {
    if (code == UCGB200_ERR_DENSITY_TYPE) error->one(FLERR, "Declared type in RLEUCG does not exist.");
    dev->check(lmp, code, "pair_rleucg");
  }
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

void PairTable_RLEUCG_INTERFACE::write_restart(FILE *fp) { write_restart_settings(fp); }

void PairTable_RLEUCG_INTERFACE::read_restart(FILE *fp) {
  read_restart_settings(fp);
  allocate();
}

void PairTable_RLEUCG_INTERFACE::write_restart_settings(FILE *fp) {
  fwrite(&tabstyle, sizeof(int), 1, fp);
  fwrite(&tablength, sizeof(int), 1, fp);
  fwrite(&ewaldflag, sizeof(int), 1, fp);
  fwrite(&pppmflag, sizeof(int), 1, fp);
  fwrite(&msmflag, sizeof(int), 1, fp);
  fwrite(&dispersionflag, sizeof(int), 1, fp);
  fwrite(&tip4pflag, sizeof(int), 1, fp);
}

void PairTable_RLEUCG_INTERFACE::read_restart_settings(FILE *fp) {
  int *vals[7] = {&tabstyle, &tablength, &ewaldflag, &pppmflag, &msmflag, &dispersionflag, &tip4pflag};
  for (auto v : vals) {
    if (comm->me == 0) utils::sfread(FLERR, v, sizeof(int), 1, fp, nullptr, error);
    MPI_Bcast(v, 1, MPI_INT, 0, world);
  }
}

double PairTable_RLEUCG_INTERFACE::single(int, int, int itype, int jtype, double rsq, double, double factor_lj, double &fforce) {
  if (!allocated || tables.empty()) error->all(FLERR, "Pair::single before pair_coeff");
  const ucgb200_table *tb = tables[tabindex[itype][jtype]];
  double phi = 0.0;
  int rc = ucgb200_host_table_single(tb, rsq, factor_lj, &phi, &fforce);
  if (rc == UCGB200_ERR_TABLE_INNER) error->one(FLERR, "Pair distance < table inner cutoff");
  if (rc == UCGB200_ERR_TABLE_OUTER) error->one(FLERR, "Pair distance > table outer cutoff");
  return phi;
}

void *PairTable_RLEUCG_INTERFACE::extract(const char *str, int &dim) {
  if (strcmp(str, "cut_coul") != 0) return nullptr;
  if (tables.empty()) error->all(FLERR, "All pair coeffs are not set");
  if (ewaldflag || pppmflag || msmflag || dispersionflag || tip4pflag) {
    for (size_t m = 1; m < tabcut.size(); m++)
      if (tabcut[m] != tabcut[0]) error->all(FLERR, "Pair table cutoffs must all be equal to use with KSpace");
    dim = 0;
    return &tabcut[0];
  }
  return nullptr;
}
