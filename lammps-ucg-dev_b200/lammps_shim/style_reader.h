#include "lammps_shim_io.h"
#include "reader_native.h"
