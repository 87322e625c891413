// PairTable_UCGLD on the GPU: the LAMMPS-facing half.  Deck parsing and error texts follow
// UCG/pair_table_ucgld.cpp (settings :654-716, coeff :719-865, init_style :867-884,
// init_one :886-895, restart :1431-1471, single :1474-1520, extract :1522-1541); the
// arithmetic lives behind the C-ABI.
#include "pair_table_ucgld.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "fix.h"
#include "force.h"
#include "memory.h"
#include "modify.h"
#include "neighbor.h"
#include "ucg_device.h"
#include "update.h"

#include <cstring>

using namespace LAMMPS_NS;

PairTable_UCGLD::PairTable_UCGLD(LAMMPS *lmp) : Pair(lmp), tabstyle(LINEAR), tablength(0), T(0.0), kT(0.0),
                                                kT_found(false), smap(nullptr), n_actual(0), n_formal(0),
                                                maps_applied(false), dev(nullptr) {
  // the device reports the per-pair virial tally itself (the shipped code leaves it 0, Q3)
  no_virial_fdotr = 1;
  restartinfo = 1;
}

PairTable_UCGLD::~PairTable_UCGLD() {
  if (copymode) return;
  for (auto t : tables) ucgb200_host_table_free(t);
  if (smap) ucgb200_host_statemap_free(smap);
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
}

void PairTable_UCGLD::allocate() {
  allocated = 1;
  const int nt = n_formal + 1;
  memory->create(setflag, nt, nt, "pair:setflag");
  memory->create(cutsq, nt, nt, "pair:cutsq");
  memset(&setflag[0][0], 0, nt * nt * sizeof(int));
  memset(&cutsq[0][0], 0, nt * nt * sizeof(double));
}

void PairTable_UCGLD::settings(int narg, char **arg) {
  if (!atom->ucg_flag) error->all(FLERR, "This pair style requires atom style ucg.");
  if (narg < 2) utils::missing_cmd_args(FLERR, "pair_style table_ucgld", error);
  if (strcmp(arg[0], "lookup") == 0) tabstyle = LOOKUP;
  else if (strcmp(arg[0], "linear") == 0) tabstyle = LINEAR;
  else if (strcmp(arg[0], "spline") == 0) tabstyle = SPLINE;
  else if (strcmp(arg[0], "bitmap") == 0) tabstyle = BITMAP;
  else error->all(FLERR, "Unknown table style in pair_style command: {}", arg[0]);
  tablength = utils::inumeric(FLERR, arg[1], false, lmp);
  if (tablength < 2) error->all(FLERR, "Illegal number of pair table entries: {}", tablength);
  if (narg < 3) utils::missing_cmd_args(FLERR, "pair_style table_ucgld", error);

  char err[512] = "";
  if (smap) ucgb200_host_statemap_free(smap);
  smap = nullptr;
  if (ucgb200_host_statemap_from_file(arg[2], &smap, err, sizeof(err))) error->one(FLERR, "{}", err);
  ucgb200_host_statemap_sizes(smap, &n_actual, &n_formal);

  for (int iarg = 3; iarg < narg; iarg++) {
    if (strcmp(arg[iarg], "ewald") == 0) ewaldflag = 1;
    else if (strcmp(arg[iarg], "pppm") == 0) pppmflag = 1;
    else if (strcmp(arg[iarg], "msm") == 0) msmflag = 1;
    else if (strcmp(arg[iarg], "dispersion") == 0) dispersionflag = 1;
    else if (strcmp(arg[iarg], "tip4p") == 0) tip4pflag = 1;
    else error->all(FLERR, "Unknown pair_style table keyword: {}", arg[iarg]);
  }
  for (auto t : tables) ucgb200_host_table_free(t);
  tables.clear();
  tabcut.clear();
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
  allocated = 0;
  maps_applied = false;
  dev = UCGDevice::get(lmp);
  dev->static_uploaded = false;   // a new pair_style: every per-site array is sent again
  dev->list_ready = false;
  dev->check(lmp, ucgb200_tables_clear(dev->ctx), "tables_clear");
}

void PairTable_UCGLD::coeff(int narg, char **arg) {
  if (narg < 7) {
    if (narg == 6) error->all(FLERR, "This pair style requires explicit definition of cutoff for each table.");
    error->all(FLERR, "Too few arguments.");
  }
  if (!smap) error->all(FLERR, "pair_coeff before pair_style");
  if (!allocated) allocate();
  int ilo, ihi, jlo, jhi;
  utils::bounds(FLERR, arg[0], 1, atom->ntypes, ilo, ihi, error);
  utils::bounds(FLERR, arg[1], 1, atom->ntypes, jlo, jhi, error);
  const int Ns_i = utils::inumeric(FLERR, arg[2], false, lmp);
  const int Ns_j = utils::inumeric(FLERR, arg[3], false, lmp);
  const int nt_this = Ns_i * Ns_j;
  if (narg != 4 + 3 * nt_this)
    error->all(FLERR, "Incorrect number of arguments for pair_coeff command. Expected 4 + 3 * n_states_i * n_states_j arguments.");
  std::vector<int> idx(nt_this);
  std::vector<double> cuts(nt_this);
  char err[512] = "";
  for (int k = 0; k < nt_this; k++) {
    const char *file = arg[4 + 3 * k], *keyword = arg[5 + 3 * k];
    const double cut = utils::numeric(FLERR, arg[6 + 3 * k], false, lmp);
    ucgb200_table *tb = nullptr;
    if (ucgb200_host_table_from_file(file, keyword, cut, tabstyle, tablength, &tb, err, sizeof(err))) error->all(FLERR, "{}", err);
    int index = -1;
    dev->check(lmp, ucgb200_host_table_upload(dev->ctx, tb, &index), "table_upload");
    if (index != (int)tables.size()) error->all(FLERR, "ucg-b200: table index mismatch");
    tables.push_back(tb);
    tabcut.push_back(cut);
    idx[k] = index;
    cuts[k] = cut;
  }
  if (ucgb200_host_statemap_coeff(smap, ilo, ihi, jlo, jhi, Ns_i, Ns_j, idx.data(), cuts.data(), err, sizeof(err)))
    error->all(FLERR, "{}", err);
  // mirror setflag for [stock] Pair::init / Info
  std::vector<int> ns(n_actual + 1), ff(2 * (n_actual + 1));
  ucgb200_host_statemap_get(smap, ns.data(), ff.data(), nullptr, nullptr, nullptr);
  for (int s_i = 0; s_i < Ns_i; s_i++)
    for (int s_j = 0; s_j < Ns_j; s_j++)
      for (int i = ilo; i <= ihi; i++)
        for (int j = MAX(jlo, i); j <= jhi; j++) {
          int fi = ns[i] == 1 ? i : ff[2 * i + s_i], fj = ns[j] == 1 ? j : ff[2 * j + s_j];
          setflag[fi][fj] = 1;
        }
  maps_applied = false;
}

void PairTable_UCGLD::init_style() {
  // no neighbor->add_request(): the device builds and owns the (full) list, following the same skin rule; LAMMPS
  // keeps its ghost exchange (cutghost = cutforce + skin) but spends nothing on a host-side pair list
  double *pT = nullptr;
  int pdim;
  kT_found = false;
  for (int ifix = 0; ifix < modify->nfix; ifix++) {
    pT = (double *) modify->fix[ifix]->extract("t_target", pdim);
    if (pT) { T = *pT; kT_found = true; break; }
  }
  // the reference silently uses an uninitialised T here (Q2); we refuse
  if (!kT_found) error->all(FLERR, "pair_style table_ucgld requires a fix that exports t_target (e.g. fix ucgld/langevin)");
  kT = force->boltz * T;
  if (force->newton_pair != 1)
    error->all(FLERR, "Newton pair is turned off. It has to be turned ON in non-CV UCG Bethe simulation.");
  maps_applied = false;
}

double PairTable_UCGLD::init_one(int i, int j) {
  if (setflag[i][j] == 0) error->all(FLERR, Error::NOLASTLINE, "All pair coeffs are not set");
  if (tabindex_flat.empty() || !maps_applied) {
    char err[512] = "";
    if (ucgb200_host_statemap_init(smap, err, sizeof(err))) error->all(FLERR, Error::NOLASTLINE, "{}", err);
    tabindex_flat.assign((n_formal + 1) * (n_formal + 1), 0);
    ucgb200_host_statemap_get(smap, nullptr, nullptr, nullptr, tabindex_flat.data(), nullptr);
  }
  return tabcut[tabindex_flat[i * (n_formal + 1) + j]];
}

void PairTable_UCGLD::apply_maps() {
  dev->sync_globals(lmp);
  std::vector<double> mass(n_formal + 1, 1.0);
  for (int t = 1; t <= n_formal && t <= atom->ntypes; t++) mass[t] = atom->mass[t];
  dev->check(lmp, ucgb200_host_statemap_apply(dev->ctx, smap, mass.data()), "statemap_apply");
  dev->check(lmp, ucgb200_set_kT(dev->ctx, kT), "set_kT");
  dev->check(lmp, ucgb200_neigh_configure(dev->ctx, neighbor->skin, 0.0), "neigh_configure");
  dev->list_ready = false;
  maps_applied = true;
}

void PairTable_UCGLD::device_compute(int eflag, int vflag) {
  dev->check(lmp, ucgb200_pair_ucgld(dev->ctx, eflag, vflag), "pair_ucgld");
}

void PairTable_UCGLD::compute(int eflag, int vflag) {
  ev_init(eflag, vflag);
  if (!maps_applied) apply_maps();
  const int nlocal = atom->nlocal;
  dev->upload(lmp, UCGB200_F_X | UCGB200_F_UCGL | UCGB200_F_UCGSTATE | UCGB200_F_UCGP);
  dev->ensure_list(lmp);
  const int ev = (eflag_either || vflag_either) ? 1 : 0;
  // per-atom tallies (compute pe/atom, stress/atom): LAMMPS' own flag bits go down to the device
  const bool want_peratom = eflag_atom || vflag_atom;
  if (want_peratom && !peratom_supported())
    error->all(FLERR, "ucg-b200: per-atom energy / virial is implemented for pair_style table_ucgld and table_ucg_bethe only");
  device_compute(ev | (eflag_atom ? 2 : 0), ev | (vflag_atom ? 4 : 0));
  if (want_peratom) {
    std::vector<double> ea(eflag_atom ? (size_t)nlocal : 0), va(vflag_atom ? 6 * (size_t)nlocal : 0);
    dev->check(lmp, ucgb200_pair_peratom(dev->ctx, nlocal, eflag_atom ? ea.data() : nullptr, vflag_atom ? va.data() : nullptr), "pair_peratom");
    if (eflag_atom) for (int i = 0; i < nlocal; i++) eatom[i] += ea[i];
    if (vflag_atom) for (int i = 0; i < nlocal; i++) for (int k = 0; k < 6; k++) vatom[i][k] += va[6 * (size_t)i + k];
  }
  if (dev->tracked) {
    // tracked offload mode: nobody on the host reads f / ucgforce / scores before the next flush, and after stock
    // Verlet's force_clear "added to zero" is "assigned": the results stay on the device
    dev->download(lmp, UCGB200_F_F | UCGB200_F_UCGFORCE | UCGB200_F_SCORES | UCGB200_F_NUMSTATES);
    int code;
    if ((code = ucgb200_status_peek(dev->ctx, nullptr, nullptr, nullptr, nullptr))) dev->check(lmp, code, "pair_ucgld");
    if (ev) {
      double e, v[6];
      dev->check(lmp, ucgb200_pair_energy_virial(dev->ctx, &e, v), "pair_energy_virial");
      if (eflag_global) eng_vdwl += e;
      for (int k = 0; k < 6; k++) {
#ifdef LAMMPS_UCG_SHIM
        virial_tally[k] = v[k];
#endif
        if (vflag_global) virial[k] += v[k];
      }
    }
    return;
  }
  // results are ADDED to the host arrays, like the reference's f[i] += ... after force_clear
  const size_t nl = (size_t)nlocal;
  char *buf = (char *)dev->scratch((6 * nl + 8) * sizeof(double) + (nl + 8) * sizeof(int));
  std::vector<char> pageable;
  if (!buf) { pageable.resize((6 * nl + 8) * sizeof(double) + (nl + 8) * sizeof(int)); buf = pageable.data(); }
  double *f = (double *)buf, *uf = f + 3 * nl, *sc = uf + nl;
  int *ns = (int *)(sc + 2 * nl);
  ucgb200_atoms h{};
  h.f = f; h.ucgforce = uf; h.ucgsoftmaxscores = sc; h.num_ucgstates = ns;
  dev->check(lmp, ucgb200_atoms_download(dev->ctx, nlocal, &h,
                                         UCGB200_F_F | UCGB200_F_UCGFORCE | UCGB200_F_SCORES | UCGB200_F_NUMSTATES),
             "atoms_download");
  int code;
  if ((code = ucgb200_status_peek(dev->ctx, nullptr, nullptr, nullptr, nullptr))) dev->check(lmp, code, "pair_ucgld");
  double **fh = atom->f, **sh = atom->ucgsoftmaxscores;
  for (int i = 0; i < nlocal; i++) {
    fh[i][0] += f[3 * i]; fh[i][1] += f[3 * i + 1]; fh[i][2] += f[3 * i + 2];
    atom->ucgforce[i] += uf[i];
    sh[i][0] += sc[2 * i]; sh[i][1] += sc[2 * i + 1];
    atom->num_ucgstates[i] = ns[i];
  }
  if (ev) {
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
}

void PairTable_UCGLD::write_restart(FILE *fp) { write_restart_settings(fp); }

void PairTable_UCGLD::read_restart(FILE *fp) {
  read_restart_settings(fp);
  allocate();
}

void PairTable_UCGLD::write_restart_settings(FILE *fp) {
  fwrite(&tabstyle, sizeof(int), 1, fp);
  fwrite(&tablength, sizeof(int), 1, fp);
  fwrite(&ewaldflag, sizeof(int), 1, fp);
  fwrite(&pppmflag, sizeof(int), 1, fp);
  fwrite(&msmflag, sizeof(int), 1, fp);
  fwrite(&dispersionflag, sizeof(int), 1, fp);
  fwrite(&tip4pflag, sizeof(int), 1, fp);
}

void PairTable_UCGLD::read_restart_settings(FILE *fp) {
  int *vals[7] = {&tabstyle, &tablength, &ewaldflag, &pppmflag, &msmflag, &dispersionflag, &tip4pflag};
  for (auto v : vals) {
    if (comm->me == 0) utils::sfread(FLERR, v, sizeof(int), 1, fp, nullptr, error);
    MPI_Bcast(v, 1, MPI_INT, 0, world);
  }
}

double PairTable_UCGLD::single(int, int, int itype, int jtype, double rsq, double, double factor_lj, double &fforce) {
  if (tabindex_flat.empty()) error->all(FLERR, "Pair::single before init");
  const ucgb200_table *tb = tables[tabindex_flat[itype * (n_formal + 1) + jtype]];
  double phi = 0.0;
  int rc = ucgb200_host_table_single(tb, rsq, factor_lj, &phi, &fforce);
  if (rc == UCGB200_ERR_TABLE_INNER) error->one(FLERR, "Pair distance < table inner cutoff");
  if (rc == UCGB200_ERR_TABLE_OUTER) error->one(FLERR, "Pair distance > table outer cutoff");
  return phi;
}

bool PairTable_UCGLD::ucg_deck(ucgb200_deck &deck) {
  if (!maps_applied) apply_maps();   // types, tables -> type maps, kT, skin: what the first compute() would send
  deck.pair_style = 0;
  return true;
}

void *PairTable_UCGLD::extract(const char *str, int &dim) {
  if (strcmp(str, "cut_coul") != 0) return nullptr;
  if (tables.empty()) error->all(FLERR, Error::NOLASTLINE, "All pair coeffs are not set");
  if (ewaldflag || pppmflag || msmflag || dispersionflag || tip4pflag) {
    for (size_t m = 1; m < tabcut.size(); m++)
      if (tabcut[m] != tabcut[0]) error->all(FLERR, Error::NOLASTLINE, "Pair table cutoffs must all be equal to use with KSpace");
    dim = 0;
    return &tabcut[0];
  }
  return nullptr;
}
