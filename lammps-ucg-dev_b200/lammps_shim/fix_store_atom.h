#include "lammps_shim_io.h"
