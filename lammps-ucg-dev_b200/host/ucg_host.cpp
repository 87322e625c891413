// ucg_host.cpp — host-side set-up logic of the UCG pair styles: tabulated-potential
// construction and the state-settings / pair_coeff bookkeeping.  Pure C++ (no CUDA calls
// except through the C-ABI upload entry points); results are bit-identical to the
// reference's compute_table because every expression keeps the reference's operand order
// (this file is compiled with -ffp-contract=off).
//
// Reference: UCG/pair_table_ucgld.cpp  read_table :897, param_extract :1067,
// spline_table :1047, compute_table :1105, spline :1375, splint :1408,
// read_state_settings :565, coeff :719, init_one :886, single :1474.
#include "ucg_host.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace ucg_host {

namespace {

union IntFloat {
  int i;
  float f;
};

// second derivatives of the natural/clamped cubic spline through (x,y)
std::vector<double> spline_second_derivs(const std::vector<double> &x, const std::vector<double> &y,
                                         double yp_first, double yp_last) {
  const int n = (int)x.size();
  std::vector<double> y2(n), u(n);
  if (yp_first > 0.99e30) {
    y2[0] = u[0] = 0.0;
  } else {
    y2[0] = -0.5;
    u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp_first);
  }
  for (int i = 1; i < n - 1; i++) {
    const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    const double p = sig * y2[i - 1] + 2.0;
    y2[i] = (sig - 1.0) / p;
    u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
    u[i] = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
  }
  double qn, un;
  if (yp_last > 0.99e30) {
    qn = un = 0.0;
  } else {
    qn = 0.5;
    un = (3.0 / (x[n - 1] - x[n - 2])) * (yp_last - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
  }
  y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
  for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
  return y2;
}

double spline_eval(const std::vector<double> &xa, const std::vector<double> &ya, const std::vector<double> &y2a,
                   double x) {
  int klo = 0, khi = (int)xa.size() - 1;
  while (khi - klo > 1) {
    const int k = (khi + klo) >> 1;
    if (xa[k] > x) khi = k; else klo = k;
  }
  const double h = xa[khi] - xa[klo];
  const double a = (xa[khi] - x) / h;
  const double b = (x - xa[klo]) / h;
  return a * ya[klo] + b * ya[khi] + ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (h * h) / 6.0;
}

// [stock] Pair::init_bitmap
struct BitmapMasks {
  int masklo, maskhi, nmask, nshiftbits;
};
BitmapMasks bitmap_masks(double inner, double outer, int ntablebits) {
  if (ntablebits > (int)sizeof(float) * 8) throw std::runtime_error("Too many total bits for bitmapped lookup table");
  if (inner >= outer) throw std::runtime_error("Table inner cutoff >= outer cutoff");
  int nlowermin = 1;
  while (!((std::pow(2.0, (double)nlowermin) <= inner * inner) &&
           (std::pow(2.0, (double)nlowermin + 1.0) > inner * inner))) {
    if (std::pow(2.0, (double)nlowermin) <= inner * inner) nlowermin++; else nlowermin--;
  }
  int nexpbits = 0;
  const double required_range = outer * outer / std::pow(2.0, (double)nlowermin);
  double available_range = 2.0;
  while (available_range < required_range) {
    nexpbits++;
    available_range = std::pow(2.0, std::pow(2.0, (double)nexpbits));
  }
  const int nmantbits = ntablebits - nexpbits;
  if (nexpbits > (int)sizeof(float) * 8 - 24) throw std::runtime_error("Too many exponent bits for lookup table");
  if (nmantbits + 1 > 24) throw std::runtime_error("Too many mantissa bits for lookup table");
  if (nmantbits < 3) throw std::runtime_error("Too few bits for lookup table");
  BitmapMasks m;
  m.nshiftbits = 23 - nmantbits;
  m.nmask = 1;
  for (int j = 0; j < ntablebits + m.nshiftbits; j++) m.nmask *= 2;
  m.nmask -= 1;
  IntFloat v;
  v.f = (float)(outer * outer);
  m.maskhi = v.i & ~(m.nmask);
  v.f = (float)(inner * inner);
  m.masklo = v.i & ~(m.nmask);
  return m;
}

}  // namespace

// r values the way read_table() regenerates them from the header (:954-972)
void TableInput::regenerate_r() {
  BitmapMasks bm{0, 0, 0, 0};
  if (rflag == UCGB200_R_BMP) {
    int bits = 0;
    while ((1 << bits) < ninput) bits++;
    if ((1 << bits) != ninput) throw std::runtime_error("Bitmapped table is incorrect length in table file");
    bm = bitmap_masks(rlo, rhi, bits);
  }
  for (int i = 0; i < ninput; i++) {
    double rnew = r[i];
    if (rflag == UCGB200_R_LINEAR) {
      rnew = rlo + (rhi - rlo) * i / (ninput - 1);
    } else if (rflag == UCGB200_R_RSQ) {
      rnew = rlo * rlo + (rhi * rhi - rlo * rlo) * i / (ninput - 1);
      rnew = std::sqrt(rnew);
    } else if (rflag == UCGB200_R_BMP) {
      IntFloat v;
      v.i = i << bm.nshiftbits;
      v.i |= bm.masklo;
      if (v.f < rlo * rlo) {
        v.i = i << bm.nshiftbits;
        v.i |= bm.maskhi;
      }
      rnew = sqrtf(v.f);
    }
    r[i] = rnew;
  }
}

TableInput TableInput::from_file(const std::string &file, const std::string &keyword) {
  std::ifstream in(file);
  if (!in) throw std::runtime_error("Cannot open table file " + file);
  std::string line;
  bool found = false;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string w;
    if (!(ls >> w) || w[0] == '#') continue;
    if (w == keyword) { found = true; break; }
  }
  if (!found) throw std::runtime_error("Did not find keyword " + keyword + " in table file");
  if (!std::getline(in, line)) throw std::runtime_error("Unexpected end of table file");
  TableInput t;
  {
    std::istringstream ls(line);
    std::string w;
    while (ls >> w) {
      if (w == "N") ls >> t.ninput;
      else if (w == "R" || w == "RSQ" || w == "BITMAP") {
        t.rflag = w == "R" ? UCGB200_R_LINEAR : (w == "RSQ" ? UCGB200_R_RSQ : UCGB200_R_BMP);
        ls >> t.rlo >> t.rhi;
      } else if (w == "FPRIME") {
        t.fpflag = 1;
        ls >> t.fplo >> t.fphi;
      } else
        throw std::runtime_error("Invalid keyword " + w + " in pair table parameters");
    }
  }
  if (t.ninput == 0) throw std::runtime_error("Pair table parameters did not set N");
  t.r.reserve(t.ninput); t.e.reserve(t.ninput); t.f.reserve(t.ninput);
  while ((int)t.r.size() < t.ninput && std::getline(in, line)) {
    std::istringstream ls(line);
    int idx;
    double r, e, f;
    if (ls >> idx >> r >> e >> f) { t.r.push_back(r); t.e.push_back(e); t.f.push_back(f); }
  }
  if ((int)t.r.size() < t.ninput)
    throw std::runtime_error("Data missing when parsing pair table '" + keyword + "' line " +
                             std::to_string(t.r.size() + 1) + " of " + std::to_string(t.ninput) + ".");
  t.regenerate_r();
  return t;
}

BuiltTable BuiltTable::build(TableInput in, double cut, int tabstyle, int tablength) {
  if (tablength < 2) throw std::runtime_error("Illegal number of pair table entries: " + std::to_string(tablength));
  // coeff()-time checks (:801-824)
  if (in.ninput <= 1) throw std::runtime_error("Invalid pair table length");
  const double rlo = in.rflag == UCGB200_R_NONE ? in.r.front() : in.rlo;
  const double rhi = in.rflag == UCGB200_R_NONE ? in.r.back() : in.rhi;
  // pair_coeff without a cutoff (pair_table_rleucg_interface.cpp:688-690): the table's upper end
  if (cut < 0.0) cut = rhi;
  if (cut <= rlo || cut > rhi) throw std::runtime_error("Pair table cutoff outside of table");
  if (rlo <= 0.0) throw std::runtime_error("Invalid pair table lower boundary");
  BuiltTable t;
  t.style = tabstyle; t.tablength = tablength; t.cut = cut;
  t.match = 0;
  if (tabstyle == UCGB200_TAB_LINEAR && in.ninput == tablength && in.rflag == UCGB200_R_RSQ && in.rhi == cut) t.match = 1;
  if (tabstyle == UCGB200_TAB_BITMAP && in.ninput == (1 << tablength) && in.rflag == UCGB200_R_BMP && in.rhi == cut) t.match = 1;
  if (in.rflag == UCGB200_R_BMP && t.match == 0)
    throw std::runtime_error("Bitmapped table in file does not match requested table");

  // spline_table (:1047-1065): only when the file values cannot be used directly
  std::vector<double> e2file, f2file;
  if (!t.match) {
    const int n = in.ninput;
    e2file = spline_second_derivs(in.r, in.e, -in.f[0], -in.f[n - 1]);
    if (!in.fpflag) {
      in.fplo = (in.f[1] - in.f[0]) / (in.r[1] - in.r[0]);
      in.fphi = (in.f[n - 1] - in.f[n - 2]) / (in.r[n - 1] - in.r[n - 2]);
    }
    f2file = spline_second_derivs(in.r, in.f, in.fplo, in.fphi);
  }
  auto e_at = [&](double r) { return spline_eval(in.r, in.e, e2file, r); };
  auto f_at = [&](double r) { return spline_eval(in.r, in.f, f2file, r); };

  const int tlm1 = tablength - 1;
  const double inner = in.rflag ? in.rlo : in.r[0];
  t.innersq = inner * inner;
  t.delta = (cut * cut - t.innersq) / tlm1;
  t.invdelta = 1.0 / t.delta;

  if (tabstyle == UCGB200_TAB_LOOKUP) {
    t.e.resize(tlm1); t.f.resize(tlm1);
    for (int i = 0; i < tlm1; i++) {
      const double rsq = t.innersq + (i + 0.5) * t.delta;
      const double r = std::sqrt(rsq);
      t.e[i] = e_at(r);
      t.f[i] = f_at(r) / r;
    }
  } else if (tabstyle == UCGB200_TAB_LINEAR) {
    t.rsq.resize(tablength); t.e.resize(tablength); t.f.resize(tablength);
    t.de.resize(tlm1); t.df.resize(tlm1);
    for (int i = 0; i < tablength; i++) {
      const double rsq = t.innersq + i * t.delta;
      const double r = std::sqrt(rsq);
      t.rsq[i] = rsq;
      if (t.match) { t.e[i] = in.e[i]; t.f[i] = in.f[i] / r; }
      else { t.e[i] = e_at(r); t.f[i] = f_at(r) / r; }
    }
    for (int i = 0; i < tlm1; i++) { t.de[i] = t.e[i + 1] - t.e[i]; t.df[i] = t.f[i + 1] - t.f[i]; }
  } else if (tabstyle == UCGB200_TAB_SPLINE) {
    t.rsq.resize(tablength); t.e.resize(tablength); t.f.resize(tablength);
    t.deltasq6 = t.delta * t.delta / 6.0;
    for (int i = 0; i < tablength; i++) {
      const double rsq = t.innersq + i * t.delta;
      const double r = std::sqrt(rsq);
      t.rsq[i] = rsq;
      if (t.match) { t.e[i] = in.e[i]; t.f[i] = in.f[i] / r; }
      else { t.e[i] = e_at(r); t.f[i] = f_at(r); }
    }
    // end-point slopes in the variable g = r^2 (:1208-1241)
    const double ep0 = -t.f[0] / (2.0 * std::sqrt(t.innersq));
    const double epn = -t.f[tlm1] / (2.0 * cut);
    t.e2 = spline_second_derivs(t.rsq, t.e, ep0, epn);
    double fp0, fpn;
    const double secant_factor = 0.1;
    if (in.fpflag)
      fp0 = (in.fplo / std::sqrt(t.innersq) - t.f[0] / t.innersq) / (2.0 * std::sqrt(t.innersq));
    else {
      const double rsq1 = t.innersq;
      const double rsq2 = rsq1 + secant_factor * t.delta;
      fp0 = (f_at(std::sqrt(rsq2)) / std::sqrt(rsq2) - t.f[0] / std::sqrt(rsq1)) / (secant_factor * t.delta);
    }
    if (in.fpflag && cut == in.r[in.ninput - 1])
      fpn = (in.fphi / cut - t.f[tlm1] / (cut * cut)) / (2.0 * cut);
    else {
      const double rsq2 = cut * cut;
      const double rsq1 = rsq2 - secant_factor * t.delta;
      fpn = (t.f[tlm1] / std::sqrt(rsq2) - f_at(std::sqrt(rsq1)) / std::sqrt(rsq1)) / (secant_factor * t.delta);
    }
    for (int i = 0; i < tablength; i++) t.f[i] /= std::sqrt(t.rsq[i]);
    t.f2 = spline_second_derivs(t.rsq, t.f, fp0, fpn);
  } else if (tabstyle == UCGB200_TAB_BITMAP) {
    const BitmapMasks bm = bitmap_masks(inner, cut, tablength);
    t.nmask = bm.nmask; t.nshiftbits = bm.nshiftbits;
    const int ntable = 1 << tablength, ntablem1 = ntable - 1;
    t.rsq.resize(ntable); t.e.resize(ntable); t.f.resize(ntable);
    t.de.resize(ntable); t.df.resize(ntable); t.drsq.resize(ntable);
    IntFloat v, vmin;
    vmin.i = 0 << t.nshiftbits;
    vmin.i |= bm.maskhi;
    for (int i = 0; i < ntable; i++) {
      v.i = i << t.nshiftbits;
      v.i |= bm.masklo;
      if (v.f < t.innersq) {
        v.i = i << t.nshiftbits;
        v.i |= bm.maskhi;
      }
      const double r = sqrtf(v.f);
      t.rsq[i] = v.f;
      if (t.match) { t.e[i] = in.e[i]; t.f[i] = in.f[i] / r; }
      else { t.e[i] = e_at(r); t.f[i] = f_at(r) / r; }
      vmin.f = std::min(vmin.f, v.f);
    }
    t.innersq = vmin.f;
    for (int i = 0; i < ntablem1; i++) {
      t.de[i] = t.e[i + 1] - t.e[i];
      t.df[i] = t.f[i + 1] - t.f[i];
      t.drsq[i] = 1.0 / (t.rsq[i + 1] - t.rsq[i]);
    }
    t.de[ntablem1] = t.e[0] - t.e[ntablem1];
    t.df[ntablem1] = t.f[0] - t.f[ntablem1];
    t.drsq[ntablem1] = 1.0 / (t.rsq[0] - t.rsq[ntablem1]);
    int itablemin = (vmin.i & t.nmask) >> t.nshiftbits;
    int itablemax = itablemin == 0 ? ntablem1 : itablemin - 1;
    int itablemaxm1 = itablemax == 0 ? ntablem1 : itablemax - 1;
    v.i = itablemax << t.nshiftbits;
    v.i |= bm.maskhi;
    if (v.f < cut * cut) {
      if (t.match) {
        t.de[itablemax] = t.de[itablemaxm1];
        t.df[itablemax] = t.df[itablemaxm1];
        t.drsq[itablemax] = t.drsq[itablemaxm1];
      } else {
        v.f = cut * cut;
        const double r = sqrtf(v.f);
        const double e_tmp = e_at(r), f_tmp = f_at(r) / r;
        t.de[itablemax] = e_tmp - t.e[itablemax];
        t.df[itablemax] = f_tmp - t.f[itablemax];
        t.drsq[itablemax] = 1.0 / (v.f - t.rsq[itablemax]);
      }
    }
  } else
    throw std::runtime_error("Unknown table style in pair_style command");
  return t;
}

// Pair::single (:1474-1520)
int BuiltTable::single(double rsq, double factor_lj, double &phi, double &fforce) const {
  const int tlm1 = tablength - 1;
  if (rsq < innersq) return UCGB200_ERR_TABLE_INNER;
  int it;
  double fraction = 0, a = 0, b = 0;
  if (style == UCGB200_TAB_BITMAP) {
    IntFloat v;
    v.f = (float)rsq;
    it = (v.i & nmask) >> nshiftbits;
    fraction = ((double)v.f - this->rsq[it]) * drsq[it];
    fforce = factor_lj * (f[it] + fraction * df[it]);
    phi = factor_lj * (e[it] + fraction * de[it]);
    return 0;
  }
  it = static_cast<int>((rsq - innersq) * invdelta);
  if (it >= tlm1) return UCGB200_ERR_TABLE_OUTER;
  if (style == UCGB200_TAB_LOOKUP) {
    fforce = factor_lj * f[it];
    phi = factor_lj * e[it];
  } else if (style == UCGB200_TAB_LINEAR) {
    fraction = (rsq - this->rsq[it]) * invdelta;
    fforce = factor_lj * (f[it] + fraction * df[it]);
    phi = factor_lj * (e[it] + fraction * de[it]);
  } else {
    b = (rsq - this->rsq[it]) * invdelta;
    a = 1.0 - b;
    fforce = factor_lj * (a * f[it] + b * f[it + 1] + ((a * a * a - a) * f2[it] + (b * b * b - b) * f2[it + 1]) * deltasq6);
    phi = factor_lj * (a * e[it] + b * e[it + 1] + ((a * a * a - a) * e2[it] + (b * b * b - b) * e2[it + 1]) * deltasq6);
  }
  return 0;
}

// ---------------------------------------------------------------- state map
StateMap StateMap::from_file(const std::string &file) {
  std::ifstream in(file);
  if (!in) throw std::runtime_error("Cannot open file " + file);
  std::string line;
  if (!std::getline(in, line)) throw std::runtime_error("Unexpected end of RLEUCG state settings file");
  StateMap m;
  int max_states = 0;
  {
    std::istringstream ls(line);
    // NOTE the order: actual, formal, max states (:582), whatever the comment at :550 says
    ls >> m.n_actual >> m.n_formal >> max_states;
  }
  if (m.n_actual < 1 || m.n_formal < m.n_actual) throw std::runtime_error("Invalid header of UCG state settings file");
  m.allocate();
  for (int i = 1; i <= m.n_actual; i++) {
    if (!std::getline(in, line)) throw std::runtime_error("Unexpected end of UCG state settings file");
    int this_type = 0;
    {
      std::istringstream ls(line);
      ls >> this_type >> m.n_states[i];
    }
    if (m.n_states[i] < 1 || m.n_states[i] > 2)
      throw std::runtime_error("Invalid number of states for atom type " + std::to_string(i) + ": " +
                               std::to_string(m.n_states[i]) + ". Only 1 or 2 states are allowed.");
    if (this_type != i)
      throw std::runtime_error("Please write orderly. Invalid atom type " + std::to_string(this_type) +
                               " in UCG state settings file. Expected " + std::to_string(i) + ".");
    if (m.n_states[i] == 2) {
      if (!std::getline(in, line)) throw std::runtime_error("Unexpected end of UCG state settings file");
      std::istringstream lf(line);
      for (int k = 0; k < 2; k++) {
        int ft;
        if (!(lf >> ft)) throw std::runtime_error("Not enough formal types specified for atom type " + std::to_string(i) + ".");
        if (ft < 1 || ft > m.n_formal) throw std::runtime_error("Formal type out of range in UCG state settings file");
        m.formal_from[2 * i + k] = ft;
      }
      if (!std::getline(in, line)) throw std::runtime_error("Unexpected end of UCG state settings file");
      std::istringstream lm(line);
      for (int k = 0; k < 2; k++) {
        double mu;
        if (!(lm >> mu)) throw std::runtime_error("Not enough formal types specified for atom type " + std::to_string(i) + ".");
        m.chem_pot[m.formal_from[2 * i + k]] = mu;
      }
    }
  }
  return m;
}

void StateMap::allocate() {
  n_states.assign(n_actual + 1, 0);
  formal_from.assign(2 * (n_actual + 1), 0);
  chem_pot.assign(n_formal + 1, 0.0);
  const int nt = n_formal + 1;
  tabindex.assign(nt * nt, 0);
  setflag.assign(nt * nt, 0);
  cutsq.assign(nt * nt, 0.0);
  tabcut.clear();
}

// the table -> formal-type assignment loop of coeff() (:833-851)
void StateMap::coeff(int ilo, int ihi, int jlo, int jhi, int ns_i, int ns_j, const int *tables, const double *cuts) {
  if (ilo < 1 || ihi > n_actual || jlo < 1 || jhi > n_actual || ilo > ihi || jlo > jhi)
    throw std::runtime_error("Illegal pair_coeff command: type range");
  for (int t = ilo; t < ihi; t++)  // (sic) the reference's check skips the upper bound (:766-775)
    if (ns_i != n_states[t]) throw std::runtime_error("Number of states for atom type does not match the number of states in the settings file.");
  for (int t = jlo; t < jhi; t++)
    if (ns_j != n_states[t]) throw std::runtime_error("Number of states for atom type does not match the number of states in the settings file.");
  const int nt = n_formal + 1;
  int k = 0;
  for (int s_i = 0; s_i < ns_i; s_i++)
    for (int s_j = 0; s_j < ns_j; s_j++, k++) {
      int count = 0;
      for (int i = ilo; i <= ihi; i++)
        for (int j = std::max(jlo, i); j <= jhi; j++) {
          // one-state types keep their own number as formal type (the reference leaves the
          // map entry 0 and then rejects it, which makes mixed CG/UCG decks unusable as
          // shipped)
          const int fi = n_states[i] == 1 ? i : formal_from[2 * i + s_i];
          const int fj = n_states[j] == 1 ? j : formal_from[2 * j + s_j];
          if (fi == 0) throw std::runtime_error("Formal type not defined in pair_style command for actual type " + std::to_string(i) + ", state " + std::to_string(s_i));
          if (fj == 0) throw std::runtime_error("Formal type not defined in pair_style command for actual type " + std::to_string(j) + ", state " + std::to_string(s_j));
          tabindex[fi * nt + fj] = tables[k];
          setflag[fi * nt + fj] = 1;
          count++;
        }
      if (count == 0) throw std::runtime_error("Illegal pair_coeff command");
      if ((int)tabcut.size() <= tables[k]) tabcut.resize(tables[k] + 1, 0.0);
      tabcut[tables[k]] = cuts[k];
    }
}

// Pair::init -> init_one(i,j) for i <= j (:886-895)
void StateMap::init() {
  const int nt = n_formal + 1;
  for (int i = 1; i <= n_formal; i++)
    for (int j = i; j <= n_formal; j++) {
      if (!setflag[i * nt + j]) throw std::runtime_error("All pair coeffs are not set");
      tabindex[j * nt + i] = tabindex[i * nt + j];
      const double cut = tabcut.at(tabindex[i * nt + j]);
      cutsq[i * nt + j] = cutsq[j * nt + i] = cut * cut;
    }
}

}  // namespace ucg_host

// ------------------------------------------------------------------ C-ABI
using namespace ucg_host;

struct ucgb200_table {
  BuiltTable t;
};
struct ucgb200_statemap {
  StateMap m;
};

static int put_err(char *errbuf, int errlen, const std::exception &e) {
  if (errbuf && errlen > 0) snprintf(errbuf, errlen, "%s", e.what());
  return -1;
}

extern "C" int ucgb200_host_table_from_file(const char *file, const char *keyword, double cut, int tabstyle,
                                            int tablength, ucgb200_table **out, char *errbuf, int errlen) {
  if (!file || !keyword || !out) return -1;
  try {
    auto *h = new ucgb200_table{BuiltTable::build(TableInput::from_file(file, keyword), cut, tabstyle, tablength)};
    *out = h;
    return 0;
  } catch (const std::exception &e) { return put_err(errbuf, errlen, e); }
}

extern "C" int ucgb200_host_table_from_arrays(int ninput, int rflag, double rlo, double rhi, int fpflag, double fplo,
                                              double fphi, const double *rfile, const double *efile,
                                              const double *ffile, double cut, int tabstyle, int tablength,
                                              ucgb200_table **out, char *errbuf, int errlen) {
  if (!efile || !ffile || !out || ninput < 0) return -1;
  try {
    TableInput in;
    in.ninput = ninput; in.rflag = rflag; in.rlo = rlo; in.rhi = rhi;
    in.fpflag = fpflag; in.fplo = fplo; in.fphi = fphi;
    in.r.assign(ninput, 0.0);
    if (rfile) in.r.assign(rfile, rfile + ninput);
    in.e.assign(efile, efile + ninput);
    in.f.assign(ffile, ffile + ninput);
    in.regenerate_r();
    *out = new ucgb200_table{BuiltTable::build(std::move(in), cut, tabstyle, tablength)};
    return 0;
  } catch (const std::exception &e) { return put_err(errbuf, errlen, e); }
}

extern "C" int ucgb200_host_table_info(const ucgb200_table *h, double params[8], int *n) {
  if (!h) return -1;
  const BuiltTable &t = h->t;
  if (params) {
    params[0] = t.innersq; params[1] = t.delta; params[2] = t.invdelta; params[3] = t.deltasq6;
    params[4] = t.cut; params[5] = t.nmask; params[6] = t.nshiftbits; params[7] = t.match;
  }
  if (n) *n = (int)t.e.size();
  return 0;
}

extern "C" int ucgb200_host_table_array(const ucgb200_table *h, int which, double *out, int cap) {
  if (!h || !out) return -1;
  const BuiltTable &t = h->t;
  const std::vector<double> *src[8] = {&t.rsq, &t.e, &t.f, &t.de, &t.df, &t.e2, &t.f2, &t.drsq};
  if (which < 0 || which > 7) return -1;
  const int n = (int)src[which]->size();
  if (cap < n) return -1;
  if (n) memcpy(out, src[which]->data(), n * sizeof(double));
  return n;
}

extern "C" int ucgb200_host_table_single(const ucgb200_table *h, double rsq, double factor_lj, double *phi, double *fforce) {
  if (!h || !phi || !fforce) return -1;
  return h->t.single(rsq, factor_lj, *phi, *fforce);
}

extern "C" void ucgb200_host_table_free(ucgb200_table *h) { delete h; }

extern "C" int ucgb200_host_table_upload(ucgb200_ctx *ctx, const ucgb200_table *h, int *index) {
  if (!ctx || !h) return -1;
  const BuiltTable &t = h->t;
  auto p = [](const std::vector<double> &v) { return v.empty() ? nullptr : v.data(); };
  // LINEAR de/df are recomputed on the device; BITMAP needs them (wrap-around entries)
  const bool bmp = t.style == UCGB200_TAB_BITMAP;
  return ucgb200_table_upload(ctx, t.style, t.tablength, (int)t.e.size(), t.innersq, t.delta, t.invdelta,
                              t.deltasq6, t.cut, t.nmask, t.nshiftbits, p(t.e), p(t.f), p(t.e2), p(t.f2),
                              bmp ? p(t.rsq) : nullptr, bmp ? p(t.drsq) : nullptr, bmp ? p(t.de) : nullptr,
                              bmp ? p(t.df) : nullptr, index);
}

extern "C" int ucgb200_host_statemap_from_file(const char *file, ucgb200_statemap **out, char *errbuf, int errlen) {
  if (!file || !out) return -1;
  try {
    *out = new ucgb200_statemap{StateMap::from_file(file)};
    return 0;
  } catch (const std::exception &e) { return put_err(errbuf, errlen, e); }
}

extern "C" int ucgb200_host_statemap_create(int n_actual, int n_formal, const int *n_states,
                                            const int *formal_from_actual, const double *chem_pot,
                                            ucgb200_statemap **out, char *errbuf, int errlen) {
  if (!n_states || !formal_from_actual || !out) return -1;
  try {
    if (n_actual < 1 || n_formal < n_actual) throw std::runtime_error("Invalid type counts");
    StateMap m;
    m.n_actual = n_actual; m.n_formal = n_formal;
    m.allocate();
    for (int i = 1; i <= n_actual; i++) {
      m.n_states[i] = n_states[i];
      if (n_states[i] < 1 || n_states[i] > 2)
        throw std::runtime_error("Invalid number of states for atom type " + std::to_string(i) + ": " +
                                 std::to_string(n_states[i]) + ". Only 1 or 2 states are allowed.");
      if (n_states[i] == 2)
        for (int k = 0; k < 2; k++) {
          int ft = formal_from_actual[2 * i + k];
          if (ft < 1 || ft > n_formal) throw std::runtime_error("Formal type out of range");
          m.formal_from[2 * i + k] = ft;
        }
    }
    for (int i = 1; i <= n_formal; i++) m.chem_pot[i] = chem_pot ? chem_pot[i] : 0.0;
    *out = new ucgb200_statemap{std::move(m)};
    return 0;
  } catch (const std::exception &e) { return put_err(errbuf, errlen, e); }
}

extern "C" void ucgb200_host_statemap_free(ucgb200_statemap *m) { delete m; }

extern "C" int ucgb200_host_statemap_sizes(const ucgb200_statemap *m, int *n_actual, int *n_formal) {
  if (!m) return -1;
  if (n_actual) *n_actual = m->m.n_actual;
  if (n_formal) *n_formal = m->m.n_formal;
  return 0;
}

extern "C" int ucgb200_host_statemap_coeff(ucgb200_statemap *m, int ilo, int ihi, int jlo, int jhi, int ns_i,
                                           int ns_j, const int *tables, const double *cuts, char *errbuf, int errlen) {
  if (!m || !tables || !cuts) return -1;
  try { m->m.coeff(ilo, ihi, jlo, jhi, ns_i, ns_j, tables, cuts); return 0; }
  catch (const std::exception &e) { return put_err(errbuf, errlen, e); }
}

extern "C" int ucgb200_host_statemap_init(ucgb200_statemap *m, char *errbuf, int errlen) {
  if (!m) return -1;
  try { m->m.init(); return 0; }
  catch (const std::exception &e) { return put_err(errbuf, errlen, e); }
}

extern "C" int ucgb200_host_statemap_get(const ucgb200_statemap *h, int *n_states, int *formal_from_actual,
                                         double *chem_pot, int *tabindex, double *cutsq) {
  if (!h) return -1;
  const StateMap &m = h->m;
  if (n_states) memcpy(n_states, m.n_states.data(), m.n_states.size() * sizeof(int));
  if (formal_from_actual) memcpy(formal_from_actual, m.formal_from.data(), m.formal_from.size() * sizeof(int));
  if (chem_pot) memcpy(chem_pot, m.chem_pot.data(), m.chem_pot.size() * sizeof(double));
  if (tabindex) memcpy(tabindex, m.tabindex.data(), m.tabindex.size() * sizeof(int));
  if (cutsq) memcpy(cutsq, m.cutsq.data(), m.cutsq.size() * sizeof(double));
  return 0;
}

extern "C" int ucgb200_host_statemap_apply(ucgb200_ctx *ctx, const ucgb200_statemap *h, const double *mass) {
  if (!ctx || !h) return -1;
  const StateMap &m = h->m;
  int rc = ucgb200_set_types(ctx, m.n_actual, m.n_formal, m.n_states.data(), m.formal_from.data(),
                             m.chem_pot.data(), mass);
  if (rc) return rc;
  return ucgb200_set_pair_maps(ctx, m.tabindex.data(), m.cutsq.data());
}
