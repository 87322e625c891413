// dump_format.cuh — "%d" and "%g" exactly as the C library prints them, usable on the device.
//
// DumpCustom::convert_string / write_lines (dump_custom.cpp:1388-1468) turn every packed double into text
// with snprintf and the default formats "%d" / "%g".  On a B200 the step takes 0.8 ms; formatting 1 M x 8
// columns with snprintf takes about a second of one host core, so a dump every thousand steps would double the
// run time.  Here every buffer element is formatted by its own thread and only the text crosses PCIe.
//
// "%g" = 6 significant digits, correctly rounded (ties to even, on the EXACT binary value, as glibc does),
// fixed notation for decimal exponents -4..5, otherwise d.ddddde[+-]XX, trailing zeros removed.
// Digits come from one scaled multiplication whenever the result is at least 1e-5 away from a rounding
// boundary (the scaled value carries an absolute error < 1e-9); otherwise — exact ties such as 0.5 -> "0.5" are
// not boundaries, 1000005 -> "1e+06" is — the decision is made in exact multi-word integer arithmetic
// (v = m * 2^e against D * 10^k), which is also what validates the fast path in tests/test_dump_format.py.
//
// The header compiles as plain C++ too (tests/fmt_harness.cpp runs it against snprintf on the host).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define UCGFMT_HD __host__ __device__
#else
#define UCGFMT_HD
#endif

namespace ucgfmt {

constexpr int BIG_WORDS = 36;   // 1152 bits: m * 5^330 and D << 800 both fit (see cmp_scaled)
struct Big {
  uint32_t w[BIG_WORDS];
  int n;
};

UCGFMT_HD inline void big_set(Big &b, uint64_t v) {
  b.w[0] = (uint32_t)v;
  b.w[1] = (uint32_t)(v >> 32);
  b.n = b.w[1] ? 2 : (b.w[0] ? 1 : 0);
}
UCGFMT_HD inline void big_mul_small(Big &b, uint32_t f) {
  uint64_t carry = 0;
  for (int i = 0; i < b.n; i++) {
    uint64_t t = (uint64_t)b.w[i] * f + carry;
    b.w[i] = (uint32_t)t;
    carry = t >> 32;
  }
  if (carry && b.n < BIG_WORDS) b.w[b.n++] = (uint32_t)carry;
}
UCGFMT_HD inline void big_mul_pow5(Big &b, int p) {
  while (p >= 13) { big_mul_small(b, 1220703125u); p -= 13; }   // 5^13 < 2^32
  uint32_t f = 1;
  for (int i = 0; i < p; i++) f *= 5u;
  if (f > 1) big_mul_small(b, f);
}
UCGFMT_HD inline void big_shl(Big &b, int s) {
  if (b.n == 0 || s == 0) return;
  const int ws = s >> 5, bs = s & 31;
  int top = b.n + ws + (bs ? 1 : 0);
  if (top > BIG_WORDS) top = BIG_WORDS;
  for (int i = top - 1; i >= 0; i--) {
    const int src = i - ws;
    uint32_t v = 0;
    if (src >= 0 && src < b.n) v = b.w[src] << bs;
    if (bs && src - 1 >= 0 && src - 1 < b.n) v |= b.w[src - 1] >> (32 - bs);
    b.w[i] = v;
  }
  b.n = top;
  while (b.n > 0 && b.w[b.n - 1] == 0) b.n--;
}
UCGFMT_HD inline int big_cmp(const Big &a, const Big &b) {
  if (a.n != b.n) return a.n > b.n ? 1 : -1;
  for (int i = a.n - 1; i >= 0; i--)
    if (a.w[i] != b.w[i]) return a.w[i] > b.w[i] ? 1 : -1;
  return 0;
}

// sign of  m * 2^e  -  D * 10^k,  exactly.  With 2^(e-k) moved to whichever side keeps it a left shift the
// operands stay below 830 bits for every finite double and the 6-digit D this file uses.
UCGFMT_HD inline int cmp_scaled(uint64_t m, int e, uint32_t D, int k) {
  Big L, R;
  big_set(L, m);
  big_set(R, D);
  if (k >= 0) big_mul_pow5(R, k); else big_mul_pow5(L, -k);
  const int a2 = e - k;
  if (a2 >= 0) big_shl(L, a2); else big_shl(R, -a2);
  return big_cmp(L, R);
}

UCGFMT_HD inline uint64_t double_bits(double v) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(v);
#else
  uint64_t b;
  memcpy(&b, &v, sizeof b);
  return b;
#endif
}
UCGFMT_HD inline double pow10_int(int n) {   // |n| <= 22: exact
  const double p[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17,
                        1e18, 1e19, 1e20, 1e21, 1e22};
  return p[n];
}
UCGFMT_HD inline double pow10_any(int n) {
#ifdef __CUDA_ARCH__
  return exp10((double)n);
#else
  return pow(10.0, (double)n);
#endif
}

// a / 10^k rounded to the nearest integer, ties to even, decided on the exact value of a = m * 2^e
UCGFMT_HD inline uint32_t round_scaled(double a, uint64_t m, int e, int k) {
  const int n = -k;
  double est;
  if (n >= 0 && n <= 22) est = a * pow10_int(n);
  else if (n < 0 && n >= -22) est = a / pow10_int(-n);
  else { const int n1 = n / 2; est = (a * pow10_any(n1)) * pow10_any(n - n1); }
  if (!(est < 4.0e9)) est = 4.0e9;
  const double fl = floor(est);
  if (fabs((est - fl) - 0.5) > 1e-5) return (uint32_t)rint(est);
  uint32_t q = (uint32_t)fl;
  while (q > 0 && cmp_scaled(m, e, q, k) < 0) q--;
  while (cmp_scaled(m, e, q + 1, k) >= 0) q++;
  const int c = cmp_scaled(m, e + 1, 2 * q + 1, k);
  if (c > 0 || (c == 0 && (q & 1))) q++;
  return q;
}

// snprintf(out, ., "%d", v): returns the length (no terminator)
UCGFMT_HD inline int format_d(int v, char *out) {
  char tmp[12];
  int len = 0, nd = 0;
  unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
  do { tmp[nd++] = (char)('0' + u % 10u); u /= 10u; } while (u);
  if (v < 0) out[len++] = '-';
  while (nd) out[len++] = tmp[--nd];
  return len;
}

// snprintf(out, ., "%g", v): returns the length (at most 13, no terminator)
UCGFMT_HD inline int format_g(double v, char *out) {
  int len = 0;
  uint64_t bits = double_bits(v);
  if (bits >> 63) out[len++] = '-';
  bits &= 0x7fffffffffffffffull;
  if (bits == 0) { out[len++] = '0'; return len; }
  const uint32_t ex = (uint32_t)(bits >> 52);
  const uint64_t frac = bits & 0xfffffffffffffull;
  if (ex == 0x7ff) {
    const char *s = frac ? "nan" : "inf";
    for (int i = 0; i < 3; i++) out[len++] = s[i];
    return len;
  }
  const uint64_t m = ex ? (frac | (1ull << 52)) : frac;
  const int e = ex ? (int)ex - 1075 : -1074;
  const double a = fabs(v);
  int X = (int)floor(log10(a));
  uint32_t q = 0;
  for (int it = 0; it < 4; it++) {
    q = round_scaled(a, m, e, X - 5);
    if (q < 100000u) { X--; continue; }
    if (q >= 1000000u) { X++; continue; }
    break;
  }
  char d[6];
  for (int i = 5; i >= 0; i--) { d[i] = (char)('0' + q % 10u); q /= 10u; }
  int nd = 6;
  while (nd > 1 && d[nd - 1] == '0') nd--;
  if (X >= -4 && X < 6) {
    if (X >= 0) {
      for (int i = 0; i <= X; i++) out[len++] = d[i];
      if (nd > X + 1) {
        out[len++] = '.';
        for (int i = X + 1; i < nd; i++) out[len++] = d[i];
      }
    } else {
      out[len++] = '0';
      out[len++] = '.';
      for (int i = 0; i < -X - 1; i++) out[len++] = '0';
      for (int i = 0; i < nd; i++) out[len++] = d[i];
    }
  } else {
    out[len++] = d[0];
    if (nd > 1) {
      out[len++] = '.';
      for (int i = 1; i < nd; i++) out[len++] = d[i];
    }
    out[len++] = 'e';
    int ax = X;
    if (ax < 0) { out[len++] = '-'; ax = -ax; } else out[len++] = '+';
    if (ax >= 100) { out[len++] = (char)('0' + ax / 100); ax %= 100; }
    out[len++] = (char)('0' + ax / 10);
    out[len++] = (char)('0' + ax % 10);
  }
  return len;
}

}  // namespace ucgfmt
