#ifdef DUMP_CLASS
// clang-format off
DumpStyle(custom/ucg/b200, DumpCustomUCGB200);
// clang-format on
#else
#ifndef LMP_DUMP_CUSTOM_UCG_B200_H
#define LMP_DUMP_CUSTOM_UCG_B200_H

// dump custom/ucg/b200: the `dump custom` of the reference's patched dump_custom.cpp for the per-atom keywords
// of atom_style ucg (id mol type mass x y z xs ys zs vx vy vz fx fy fz q proc ucgstate ucgl ucgp, plus ucgforce
// ucgvl ucgml directly), with selection, ordering, packing and — for the default formats — the text conversion
// done on the device.  Same file format, same dump_modify keywords (sort id, thresh, format, append, header, time,
// units, pad, flush).  Columns that need a compute / fix / variable stay with the stock `dump custom`.

#include "dump.h"

struct ucgb200_dump;

namespace LAMMPS_NS {

class DumpCustomUCGB200 : public Dump {
 public:
  DumpCustomUCGB200(class LAMMPS *, int, char **);
  ~DumpCustomUCGB200() override;
  void write() override;

 protected:
  ucgb200_dump *handle;
  int ncolumns;

  void init_style() override;
  int modify_param(int, char **) override;
  void write_header(bigint) override {}
  void pack(tagint *) override {}
  void write_data(int, double *) override {}
  void forward(int narg, const char *const *arg);
};

}  // namespace LAMMPS_NS
#endif
#endif
