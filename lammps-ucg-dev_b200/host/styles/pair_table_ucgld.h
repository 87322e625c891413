#ifdef PAIR_CLASS
// clang-format off
PairStyle(table_ucgld, PairTable_UCGLD)
// clang-format on
#else
#ifndef LMP_PAIR_TABLE_UCGLD_H
#define LMP_PAIR_TABLE_UCGLD_H

// GPU-backed drop-in for the reference's PairTable_UCGLD (UCG/pair_table_ucgld.h:22-48):
// same style name, same deck grammar, same virtual-method set; compute() runs
// ucgb200_pair_ucgld on the device.

#include "pair.h"
#include "ucgb200_host.h"
#include "ucg_device.h"

#include <vector>

namespace LAMMPS_NS {

class PairTable_UCGLD : public Pair, public UCGDeckPart {
 public:
  PairTable_UCGLD(class LAMMPS *);
  ~PairTable_UCGLD() override;
  void compute(int, int) override;
  void settings(int, char **) override;
  void coeff(int, char **) override;
  void init_style() override;
  double init_one(int, int) override;
  void write_restart(FILE *) override;
  void read_restart(FILE *) override;
  void write_restart_settings(FILE *) override;
  void read_restart_settings(FILE *) override;
  double single(int, int, int, int, double, double, double, double &) override;
  void *extract(const char *, int &) override;
  bool ucg_deck(ucgb200_deck &deck) override;   // this style's part of the resident deck (run_style ucg/b200)
  bool ucg_tracked_ok() const override { return true; }
  virtual bool peratom_supported() const { return true; }   // eflag_atom / vflag_atom tallies on the device
  enum { LOOKUP, LINEAR, SPLINE, BITMAP };

 protected:
  int tabstyle, tablength;
  double T, kT;
  bool kT_found;
  ucgb200_statemap *smap;
  std::vector<ucgb200_table *> tables;   // host copies (Pair::single, restart)
  std::vector<double> tabcut;
  std::vector<int> tabindex_flat;         // (n_formal+1)^2 after init
  int n_actual, n_formal;
  bool maps_applied;
  class UCGDevice *dev;

  void allocate();
  void apply_maps();
  virtual void device_compute(int eflag, int vflag);
};

}  // namespace LAMMPS_NS
#endif
#endif
