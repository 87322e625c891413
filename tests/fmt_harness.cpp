// Host build of lammps-ucg-dev_b200/csrc/dump_format.cuh (the device "%g" / "%d" formatter of the dump tap)
// for tests/test_dump_format.py: every value is compared with the C library's snprintf, which is what
// DumpCustom::convert_string calls (dump_custom.cpp:1388-1421).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../lammps-ucg-dev_b200/csrc/dump_format.cuh"

static uint64_t sm64(uint64_t &s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

extern "C" {

int fmt_g(double v, char *out) { int n = ucgfmt::format_g(v, out); out[n] = 0; return n; }
int fmt_d(int v, char *out) { int n = ucgfmt::format_d(v, out); out[n] = 0; return n; }

static int check_one(double v, double *bad) {
  char a[64], b[64];
  int n = ucgfmt::format_g(v, a);
  a[n] = 0;
  snprintf(b, sizeof b, "%g", v);
  if (strcmp(a, b) != 0) { *bad = v; return 1; }
  return 0;
}

// mode 0: random bit patterns (all exponents); 1: uniform magnitudes 1e-6..1e7 (what a dump holds);
// 2: decimal strings with 7 significant digits ending in 5 (the nearest doubles sit on or next to a rounding
// boundary); 3: integers and half-integers scaled by powers of two (exact ties); 4: the double just below /
// at / just above every 6-digit boundary of the form (D + 0.5) * 10^k
long long fmt_check(int mode, unsigned long long seed, long long count, double *first_bad) {
  uint64_t s = seed;
  long long nbad = 0;
  double bad = 0;
  for (long long i = 0; i < count; i++) {
    double v;
    uint64_t r = sm64(s);
    if (mode == 0) {
      memcpy(&v, &r, sizeof v);
    } else if (mode == 1) {
      double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
      int ex = (int)(sm64(s) % 14) - 6;
      v = (u + 0.1) * pow(10.0, ex);
      if (r & 1) v = -v;
    } else if (mode == 2) {
      char txt[64];
      int ex = (int)(sm64(s) % 60) - 30;
      snprintf(txt, sizeof txt, "%llu5e%d", (unsigned long long)(100000 + r % 900000), ex);
      v = strtod(txt, nullptr);
      int nudge = (int)(sm64(s) % 3) - 1;
      uint64_t b;
      memcpy(&b, &v, sizeof b);
      b += nudge;
      memcpy(&v, &b, sizeof v);
    } else if (mode == 3) {
      long long k = (long long)(r % 4000001) ;
      int sh = (int)(sm64(s) % 40) - 20;
      v = ldexp((double)(2 * k + 1), sh - 1);
    } else {
      int ex = (int)(sm64(s) % 40) - 20;
      double D = (double)(100000 + r % 900000) + 0.5;
      v = D * pow(10.0, ex);
      int nudge = (int)(sm64(s) % 5) - 2;
      uint64_t b;
      memcpy(&b, &v, sizeof b);
      b += nudge;
      memcpy(&v, &b, sizeof v);
    }
    if (check_one(v, &bad)) { if (!nbad && first_bad) *first_bad = bad; nbad++; }
  }
  return nbad;
}
}
