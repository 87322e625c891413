#ifdef COMMAND_CLASS
// clang-format off
CommandStyle(read_dump/ucg/b200, ReadDumpUCGB200);
// clang-format on
#else
#ifndef LMP_READ_DUMP_UCG_B200_H
#define LMP_READ_DUMP_UCG_B200_H

// read_dump/ucg/b200: `read_dump file Nstep fields keywords` (read_dump.cpp of the reference, incl. its ucgstate
// ucgl ucgp fields) applied to the device-resident atoms: rows are matched by id and scattered on the device, then
// the host arrays are refreshed from it.  replace / trim / box / scaled / wrapped / label; no purge / add.

#include "command.h"

namespace LAMMPS_NS {

class ReadDumpUCGB200 : public Command {
 public:
  ReadDumpUCGB200(class LAMMPS *lmp) : Command(lmp) {}
  void command(int, char **) override;
};

}  // namespace LAMMPS_NS
#endif
#endif
