#ifdef INTEGRATE_CLASS
// clang-format off
IntegrateStyle(ucg/b200, VerletUCGB200);
// clang-format on
#else
#ifndef LMP_VERLET_UCG_B200_H
#define LMP_VERLET_UCG_B200_H

// run_style ucg/b200: the [stock] Verlet loop of a UCG deck kept on the device.  With the default run_style the
// UCG classes work in offload mode (every style call moves the arrays it touches); with this one the deck — any of the
// four UCG pair styles, fix nve/ucgld(/wall/hard), ucgld/langevin, ucgstate, cluster_switch — is handed to ucgb200_setup /
// ucgb200_run_between once, and the host arrays are refreshed on output steps and at the end of the run only.
// Same kernels, same order, same random streams: the trajectory is the offload-mode one bit for bit.

#include <vector>

#include "integrate.h"
#include "ucgb200.h"

namespace LAMMPS_NS {

class VerletUCGB200 : public Integrate {
 public:
  VerletUCGB200(class LAMMPS *, int, char **);
  void init() override;
  void setup(int) override;
  void setup_minimal(int) override;
  void run(int) override;
  void cleanup() override;

 protected:
  class UCGDevice *dev;
  ucgb200_deck deck;
  std::vector<class UCGDeckPart *> parts;   // the fixes that take part in the device loop
  void collect_deck();
  void push();            // host arrays -> device (everything the run starts from)
  void pull(bool thermo); // device -> host arrays, pair energy / virial on thermo steps
};

}  // namespace LAMMPS_NS
#endif
#endif
