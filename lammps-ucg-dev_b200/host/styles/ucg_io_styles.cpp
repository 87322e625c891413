// dump custom/ucg/b200 and read_dump/ucg/b200: LAMMPS-side bindings of the dump / read_dump taps
// (include/ucgb200_host.h).  The classes own no formatting or parsing: they marshal the command words and the
// base-class dump_modify state into the host library and keep the host arrays and the device copy consistent.
// Mirrors: dump_custom.cpp (ctor :57-166, init_style :244, modify_param :1938), [stock] Dump::write,
// read_dump.cpp command :80-152 of the reference tree.
#include "dump_custom_ucg_b200.h"
#include "read_dump_ucg_b200.h"

#include <cstring>
#include <string>
#include <vector>

#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "group.h"
#include "memory.h"
#include "update.h"

#include "ucg_device.h"

using namespace LAMMPS_NS;

DumpCustomUCGB200::DumpCustomUCGB200(LAMMPS *lmp, int narg, char **arg) : Dump(lmp, narg, arg), handle(nullptr), ncolumns(narg - 5) {
  if (!atom->ucg_flag) error->all(FLERR, "Dump custom/ucg/b200 requires atom_style ucg");
  // the host library knows the style as `custom`: same words otherwise
  std::vector<const char *> words(arg, arg + narg);
  words[2] = "custom";
  char err[512];
  if (ucgb200_host_dump_create(narg, words.data(), groupbit, &handle, err, sizeof err)) error->all(FLERR, "{}", err);
  size_one = ncolumns;
  clearstep = 0;
  buffer_allow = 1;
  buffer_flag = 1;
}

DumpCustomUCGB200::~DumpCustomUCGB200() { ucgb200_host_dump_free(handle); }

void DumpCustomUCGB200::forward(int narg, const char *const *arg) {
  char err[512];
  if (ucgb200_host_dump_modify(handle, narg, arg, err, sizeof err)) error->all(FLERR, "{}", err);
}

void DumpCustomUCGB200::init_style() {
  // the keywords Dump::modify_params keeps in the base class travel to the writer before every run
  const std::string pad = std::to_string(padflag);
  std::vector<const char *> w = {"sort", sort_flag ? "id" : "off", "append", append_flag ? "yes" : "no", "header", header_flag ? "yes" : "no",
                                 "time", time_flag ? "yes" : "no", "units", unit_flag ? "yes" : "no", "flush", flush_flag ? "yes" : "no",
                                 "pad", pad.c_str()};
  if (sort_flag && sortcol != 0) error->all(FLERR, "Dump custom/ucg/b200 sorts by id only");
  forward((int)w.size(), w.data());
  if (format_line_user) { const char *f[] = {"format", "line", format_line_user}; forward(3, f); }
}

// dump_modify keywords the base class hands down: format int/float/M/none, thresh
int DumpCustomUCGB200::modify_param(int narg, char **arg) {
  int n = 0;
  if (strcmp(arg[0], "format") == 0) n = (narg > 1 && strcmp(arg[1], "none") == 0) ? 2 : 3;
  else if (strcmp(arg[0], "thresh") == 0) n = (narg > 1 && strcmp(arg[1], "none") == 0) ? 2 : 4;
  else return 0;
  if (narg < n) utils::missing_cmd_args(FLERR, std::string("dump_modify ") + arg[0], error);
  forward(n, arg);
  return n;
}

void DumpCustomUCGB200::write() {
  UCGDevice *dev = UCGDevice::get(lmp);
  dev->sync_globals(lmp);
  // offload mode: the host arrays are the truth between style calls; a resident run skips this upload
  dev->upload(lmp, UCGB200_F_X | UCGB200_F_V | UCGB200_F_F | UCGB200_F_UCGSTATE | UCGB200_F_UCGL | UCGB200_F_UCGVL | UCGB200_F_UCGP |
                       UCGB200_F_UCGFORCE | UCGB200_F_MASK | UCGB200_F_TYPE);
  char err[512];
  if (ucgb200_host_dump_write(handle, dev->ctx, update->ntimestep, compute_time(), update->unit_style, err, sizeof err))
    error->one(FLERR, "{}", err);
}

void ReadDumpUCGB200::command(int narg, char **arg) {
  if (domain->box_exist == 0) error->all(FLERR, "Read_dump command before simulation box is defined");
  if (narg < 2) utils::missing_cmd_args(FLERR, "read_dump", error);
  if (comm->nprocs > 1) error->all(FLERR, "read_dump/ucg/b200 runs on one process per GPU context");
  UCGDevice *dev = UCGDevice::get(lmp);
  dev->sync_globals(lmp);
  dev->upload(lmp, UCGB200_F_ALL & ~(UCGB200_F_NUMSTATES | UCGB200_F_SCORES));
  long long stats[7];
  char err[512];
  if (ucgb200_host_read_dump(dev->ctx, narg, arg, stats, err, sizeof err)) error->all(FLERR, "{}", err);
  // box yes: the snapshot box is now the simulation box
  int per[3];
  ucgb200_get_box(dev->ctx, domain->boxlo, domain->boxhi, per);
  domain->set_initial_box();
  domain->set_global_box();
  domain->set_local_box();
  // refresh the host arrays (trim may have removed atoms)
  atom->nlocal = (int) stats[6];
  atom->natoms = stats[6];
  atom->nghost = 0;
  dev->nlocal_dev = atom->nlocal;
  dev->list_ready = false;
  dev->download(lmp, UCGB200_F_ALL & ~(UCGB200_F_NUMSTATES | UCGB200_F_SCORES));
  // timestep yes (default): reset to the snapshot's step
  bool timestepflag = true;
  for (int i = 2; i + 1 < narg; i++) if (strcmp(arg[i], "timestep") == 0) timestepflag = utils::logical(FLERR, arg[i + 1], false, lmp);
  if (timestepflag) update->reset_timestep(utils::bnumeric(FLERR, arg[1], false, lmp), true);
  if (comm->me == 0)
    utils::logmesg(lmp, "  {} atoms before read\n  {} atoms in snapshot\n  {} atoms purged\n  {} atoms replaced\n  {} atoms trimmed\n"
                        "  {} atoms added\n  {} atoms after read\n", stats[0], stats[1], stats[2], stats[3], stats[4], stats[5], stats[6]);
}
