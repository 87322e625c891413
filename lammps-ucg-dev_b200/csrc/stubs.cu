// stubs.cu — entry points declared in include/ucgb200.h whose kernels are not built yet.
// They fail loudly (never silently fall back to a CPU path).
#include "ucg_internal.cuh"
using namespace ucg;
#define NOTYET(name) { if (!c) return -1; return fail(c, name ": not implemented in this build"); }

