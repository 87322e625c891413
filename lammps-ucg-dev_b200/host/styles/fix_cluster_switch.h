#ifdef FIX_CLASS
// clang-format off
FixStyle(cluster_switch, FixClusterSwitch);
// clang-format on
#else
#ifndef LMP_FIX_CLUSTER_SWITCH_H
#define LMP_FIX_CLUSTER_SWITCH_H

// GPU-backed drop-in for FixClusterSwitch (UCG/fix_cluster_switch.h:3-90):
//   fix ID group cluster_switch molID_seed mol_offset cutoff seed rateFreq N rateFile f contactFile f
// Cluster labelling and the Monte-Carlo type switching run in csrc/cluster_switch.cu; this class keeps
// the fix-line grammar, the two input-file formats, the log files and compute_vector.

#include "fix.h"
#include "ucg_device.h"

#include <cstdio>
#include <vector>

namespace LAMMPS_NS {

class FixClusterSwitch : public Fix, public UCGDeckPart {
 public:
  FixClusterSwitch(class LAMMPS *, int, char **);
  ~FixClusterSwitch() override;
  int setmask() override;
  void init() override;
  void init_list(int, class NeighList *) override;
  void pre_exchange() override;
  double compute_vector(int) override;
  double memory_usage() override;
  int pack_forward_comm(int, int *, double *, int, int *) override;
  void unpack_forward_comm(int, int, double *) override;
  bool ucg_deck(ucgb200_deck &deck) override;
  long long ucg_next_stop() const override;
  void ucg_after_step(long long) override;

 private:
  void write_logs(int now, bool after_switch);
  int mol_seed, mol_offset, seed, switchFreq;
  double cutoff, probON, probOFF;
  int nSwitchTypes, nContactTypes, nAtomsPerContact, maxmol;
  std::vector<int> atomtypesON, atomtypesOFF, contactPairs;
  class NeighList *list;
  FILE *fp1, *fp2;
  class UCGDevice *dev;

  void read_file(char *);
  void read_contacts(char *);
};

}  // namespace LAMMPS_NS
#endif
#endif
