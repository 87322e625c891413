// stubs.cu — entry points declared in include/ucgb200.h whose kernels are not built yet.
// They fail loudly (never silently fall back to a CPU path).
#include "ucg_internal.cuh"
using namespace ucg;
#define NOTYET(name) { if (!c) return -1; return fail(c, name ": not implemented in this build"); }

extern "C" int ucgb200_cluster_configure(ucgb200_ctx *c, int, int, double, int, const int *, const int *, const double *, const double *, int, const int *, int) NOTYET("cluster_configure")
extern "C" int ucgb200_cluster_check(ucgb200_ctx *c, int *, int *) NOTYET("cluster_check")
extern "C" int ucgb200_cluster_switch(ucgb200_ctx *c, int, long long, int *, int *) NOTYET("cluster_switch")
