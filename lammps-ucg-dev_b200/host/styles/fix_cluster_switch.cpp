// FixClusterSwitch on the GPU: the LAMMPS-facing half.  Fix-line grammar, file formats, log files and
// error texts follow UCG/fix_cluster_switch.cpp (constructor :36-186, read_file :205-281,
// read_contacts :285-357, init :368-402, pre_exchange :464-481, compute_vector :923-933).
#include "fix_cluster_switch.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "neighbor.h"
#include "pair.h"
#include "ucg_device.h"
#include "update.h"

#include <cstdlib>
#include <cstring>

using namespace LAMMPS_NS;
using namespace FixConst;

#define MAXLINE 1024

FixClusterSwitch::FixClusterSwitch(LAMMPS *lmp, int narg, char **arg)
    : Fix(lmp, narg, arg), nSwitchTypes(0), nContactTypes(0), nAtomsPerContact(0), maxmol(-1), list(nullptr), fp1(nullptr),
      fp2(nullptr), dev(nullptr) {
  if (narg < 13) error->all(FLERR, "Illegal cluster_switch command");
  mol_seed = utils::inumeric(FLERR, arg[3], false, lmp);
  mol_offset = utils::inumeric(FLERR, arg[4], false, lmp);
  cutoff = utils::numeric(FLERR, arg[5], false, lmp);
  seed = utils::inumeric(FLERR, arg[6], false, lmp);
  switchFreq = utils::inumeric(FLERR, arg[8], false, lmp);
  if (seed <= 0) error->all(FLERR, "Invalid seed for Park random # generator");
  read_file(arg[10]);
  read_contacts(arg[12]);
  if (force->pair == nullptr) error->all(FLERR, "fix cluster_switch requires a pair style");
  if (force->pair->cutsq == nullptr) error->all(FLERR, "fix cluster_switch is incompatible with pair style");
  if (atom->molecule_flag == 0) error->all(FLERR, "fix cluster_switch requires that atoms have molecule attributes");

  force_reneighbor = 1;
  next_reneighbor = update->ntimestep + 1;
  comm_forward = 1;
  vector_flag = 1;
  size_vector = 7;
  global_freq = 1;
  extvector = 0;
  time_depend = 1;

  // the constructor scan over the atoms (maxmol, nSwitchPerMol, mol_state, mol_restrict) runs on the device
  dev = UCGDevice::get(lmp);
  dev->sync_globals(lmp);
  dev->static_uploaded = false;   // types / molecule ids / masks as they are NOW (the scan below reads them)
  dev->upload(lmp, 0);
  int rc = ucgb200_cluster_configure(dev->ctx, mol_seed, mol_offset, cutoff, seed, probON, nSwitchTypes, atomtypesON.data(),
                                     atomtypesOFF.data(), (int) contactPairs.size() / 2, contactPairs.data(), atom->ntypes,
                                     groupbit);
  if (rc) error->all(FLERR, "{}", ucgb200_last_error(dev->ctx));
  ucgb200_cluster_get(dev->ctx, 0, nullptr, nullptr, nullptr, nullptr, &maxmol);

  if (comm->me == 0) {
    fp1 = fopen("cluster_assignment.log", "w");
    if (fp1 == nullptr) error->one(FLERR, "File cluster_assignment.log in cluster_switch not open!\n");
    fp2 = fopen("state_assignment.log", "w");
    if (fp2 == nullptr) error->one(FLERR, "File state_assignment.log in cluster_switch not open!\n");
  }
}

FixClusterSwitch::~FixClusterSwitch() {
  if (fp1 && comm->me == 0) fclose(fp1);
  if (fp2 && comm->me == 0) fclose(fp2);
}

// one "data line" = a line that is not blank after stripping the '#' comment
static bool next_words(FILE *fp, char *line, std::vector<char *> &words, int &lineNum) {
  while (fgets(line, MAXLINE, fp)) {
    lineNum++;   // (sic) the reference counts every physical line, blank ones included (:238)
    char *ptr;
    if ((ptr = strchr(line, '#'))) *ptr = '\0';
    words.clear();
    for (char *w = strtok(line, " \t\n\r\f"); w; w = strtok(nullptr, " \t\n\r\f")) words.push_back(w);
    if (!words.empty()) return true;
  }
  return false;
}

void FixClusterSwitch::read_file(char *file) {
  FILE *fp = utils::open_potential(file, lmp, nullptr);
  if (fp == nullptr) error->one(FLERR, "Cannot open file {}: {}", file, utils::getsyserror());
  char line[MAXLINE];
  std::vector<char *> words;
  int lineNum = 0;
  while (next_words(fp, line, words, lineNum)) {
    if (lineNum == 1) {
      probON = atof(words[0]);
      if (probON > 1.0) error->one(FLERR, "Incorrect probability in rates.txt files (fix cluster_switch)");
      probOFF = 1.0 - probON;
    } else if (lineNum == 2) {
      nSwitchTypes = atoi(words[0]);
      if (nSwitchTypes > atom->ntypes) error->one(FLERR, "Incorrect number of atom switching types (fix cluster_switch)");
      atomtypesON.assign(nSwitchTypes, 0);
      atomtypesOFF.assign(nSwitchTypes, 0);
    } else if (lineNum == 3) {
      for (int i = 0; i < nSwitchTypes && i < (int) words.size(); i++) atomtypesON[i] = atoi(words[i]);
    } else if (lineNum == 4) {
      for (int i = 0; i < nSwitchTypes && i < (int) words.size(); i++) atomtypesOFF[i] = atoi(words[i]);
    }
  }
  fclose(fp);
  if (nSwitchTypes < 1) error->one(FLERR, "Incorrect number of atom switching types (fix cluster_switch)");
}

void FixClusterSwitch::read_contacts(char *file) {
  FILE *fp = utils::open_potential(file, lmp, nullptr);
  if (fp == nullptr) error->one(FLERR, "Cannot open file {}: {}", file, utils::getsyserror());
  char line[MAXLINE];
  std::vector<char *> words;
  int lineNum = 0;
  while (next_words(fp, line, words, lineNum)) {
    if (lineNum == 1) nContactTypes = words.size() > 1 ? atoi(words[1]) : 0;
    else if (lineNum == 2) nAtomsPerContact = words.size() > 1 ? atoi(words[1]) : 0;
    else if (words.size() >= 2 && (int) contactPairs.size() / 2 < nContactTypes * nAtomsPerContact) {
      contactPairs.push_back(atoi(words[0]));   // contactMap[i][j][0..1] (:342-346): an ordered (itype, jtype) pair
      contactPairs.push_back(atoi(words[1]));
    }
  }
  fclose(fp);
}

int FixClusterSwitch::setmask() { return PRE_EXCHANGE; }

void FixClusterSwitch::init() {
  // the reference asks for its own full list (:395); the device keeps one full list for every style
  // (no host-side list either: the labelling runs on the device list)
}

void FixClusterSwitch::init_list(int, NeighList *ptr) { list = ptr; }

void FixClusterSwitch::pre_exchange() {
  if (switchFreq == 0) return;
  if (next_reneighbor != update->ntimestep) return;
  // domain->pbc(); comm->exchange(); comm->borders(); neighbor->build(1) of the reference (:471-475) happen on
  // the device copy: positions and types go up, the device wraps, sorts and rebuilds its list
  dev->upload(lmp, UCGB200_F_X | UCGB200_F_TYPE | UCGB200_F_MASK | UCGB200_F_MOLECULE);
  dev->check(lmp, ucgb200_neigh_build(dev->ctx), "neigh_build");
  dev->list_ready = true;
  int ncl = 0, natt = 0, nsuc = 0;
  dev->check(lmp, ucgb200_cluster_check(dev->ctx, &ncl), "cluster_check");
  if (comm->me == 0) write_logs((int) update->ntimestep, false);   // written before the switch
  dev->check(lmp, ucgb200_cluster_switch(dev->ctx, &natt, &nsuc), "cluster_switch");
  dev->download(lmp, UCGB200_F_TYPE);
  comm->forward_comm(this);   // changed types to the ghosts (:825)
  next_reneighbor = update->ntimestep + switchFreq;
}

// cluster_assignment.log / state_assignment.log (:711-727): the labelling and the molecule states as they were BEFORE
// the switch of that step; after_switch undoes the accepted flips (state 0 <-> 1 where mol_accept == 1)
void FixClusterSwitch::write_logs(int now, bool after_switch) {
  std::vector<int> cl(maxmol + 1), st(maxmol + 1), acc(maxmol + 1, 0);
  dev->check(lmp, ucgb200_cluster_get(dev->ctx, maxmol + 1, cl.data(), st.data(), nullptr, after_switch ? acc.data() : nullptr, nullptr),
             "cluster_get");
  const int cid = cl[mol_seed];
  fprintf(fp1, "%d ", now);
  fprintf(fp2, "%d ", now);
  for (int i = 0; i <= maxmol; i++) {
    int state = st[i];
    if (after_switch && acc[i] == 1 && (state == 0 || state == 1)) state = 1 - state;
    fprintf(fp1, "%d ", cl[i] == cid ? 1 : 0);
    fprintf(fp2, "%d ", state);
  }
  fprintf(fp1, "\n");
  fprintf(fp2, "\n");
  fflush(fp1);
  fflush(fp2);
}

// run_style ucg/b200: the device loop rebuilds, labels and switches at next_reneighbor itself (csrc/run.cu); it returns
// right after that step so that the two log lines are written
bool FixClusterSwitch::ucg_deck(ucgb200_deck &deck) {
  deck.cluster_freq = switchFreq;
  if (switchFreq > 0) dev->check(lmp, ucgb200_cluster_next_reneighbor(dev->ctx, next_reneighbor), "cluster_next_reneighbor");
  return true;
}
long long FixClusterSwitch::ucg_next_stop() const { return switchFreq > 0 ? (long long) next_reneighbor : -1; }
void FixClusterSwitch::ucg_after_step(long long step) {
  if (switchFreq == 0 || step != next_reneighbor) return;
  if (comm->me == 0) write_logs((int) step, true);
  next_reneighbor = step + switchFreq;
}

int FixClusterSwitch::pack_forward_comm(int n, int *lst, double *buf, int, int *) {
  int m;
  for (m = 0; m < n; m++) buf[m] = atom->type[lst[m]];
  return m;
}

void FixClusterSwitch::unpack_forward_comm(int n, int first, double *buf) {
  for (int m = 0, i = first; m < n; m++, i++) atom->type[i] = static_cast<int>(buf[m]);
}

double FixClusterSwitch::compute_vector(int n) {
  double s[8];
  ucgb200_cluster_stats(dev->ctx, s);
  return (n >= 0 && n < 7) ? s[n] : 0.0;
}

double FixClusterSwitch::memory_usage() { return 0; }
