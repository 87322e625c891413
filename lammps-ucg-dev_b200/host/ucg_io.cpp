// ucg_io.cpp — host side of the taps around the resident UCG step:
// `dump custom` / `dump_modify` / `read_dump` / `read_data` with the reference's grammar, header and
// line formats, sitting on the device entry points of csrc/dump.cu.  Pure C++ above the C-ABI.
//
// Reference (the patched stock files at the top of the reference tree):
//   dump_custom.cpp   ctor :57-166, init_style :244-330, header_item :651-672, count :721-1368, pack :1372-1384,
//                     convert_string :1388-1421, write_lines :1448-1468, parse_fields :1471-1843 (UCG :1672-1687),
//                     modify_param :1938-2335 (format :1967-2009, thresh :2012-2335, UCG attributes :2150-2155)
//   [stock] dump.cpp  Dump::write / openfile / modify_params (append buffer flush header pad sort time units format)
//   read_dump.cpp     command :80-152, header :443-567, atoms :573-666, process_atoms :797-935,
//                     fields_and_keywords :1169-1318, whichtype :1326-1352, xfield :1359-1380
//   reader_native.cpp read_time :54-103, skip :110-149, read_header :186-444 (UCG labels :423-433), read_atoms :452-501
//   UCG/atom_vec_ucg.cpp  fields_data_atom / fields_data_vel :85-90, data_atom_post :145-170, property_atom :172-181
#include <algorithm>
#include <cerrno>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ucgb200_host.h"

namespace {

struct IoError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

int fail(char *errbuf, int errlen, const std::string &msg) {
  if (errbuf && errlen > 0) snprintf(errbuf, errlen, "%s", msg.c_str());
  return -1;
}
std::string ctx_error(ucgb200_ctx *ctx) {
  const char *m = ucgb200_last_error(ctx);
  return m ? m : "";
}
std::vector<std::string> split_words(const std::string &s) {
  std::vector<std::string> out;
  std::istringstream is(s);
  std::string w;
  while (is >> w) out.push_back(w);
  return out;
}
double numeric(const std::string &s) {
  char *end = nullptr;
  double v = strtod(s.c_str(), &end);
  if (s.empty() || end == s.c_str() || *end != '\0')
    throw IoError("Expected floating point parameter instead of '" + s + "' in input script or data file");
  return v;
}
long long inumeric(const std::string &s) {
  char *end = nullptr;
  long long v = strtoll(s.c_str(), &end, 10);
  if (s.empty() || end == s.c_str() || *end != '\0')
    throw IoError("Expected integer parameter instead of '" + s + "' in input script or data file");
  return v;
}
int logical(const std::string &s) {
  if (s == "yes" || s == "on" || s == "true") return 1;
  if (s == "no" || s == "off" || s == "false") return 0;
  throw IoError("Expected boolean parameter instead of '" + s + "' in input script or data file");
}

// page-locked staging that grows geometrically (falls back to pageable memory if pinning fails)
struct Staging {
  char *p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  ~Staging() { release(); }
  void release() {
    if (p) { if (pinned) ucgb200_pinned_free(p); else free(p); }
    p = nullptr; cap = 0;
  }
  char *ensure(size_t n) {
    if (n <= cap) return p;
    release();
    size_t want = n + n / 4 + 4096;
    void *q = nullptr;
    if (ucgb200_pinned_alloc(want, &q) == 0 && q) { p = (char *)q; pinned = true; }
    else { p = (char *)malloc(want); pinned = false; if (!p) throw IoError("out of host memory"); }
    cap = want;
    return p;
  }
};

// [0, n) split into contiguous chunks over the host cores (UCGB200_IO_THREADS overrides the count)
template <class F>
void parallel_chunks(long long n, long long min_chunk, F &&fn) {
  int nt = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("UCGB200_IO_THREADS")) nt = atoi(e);
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if ((long long)nt * min_chunk > n) nt = (int)std::max<long long>(1, n / std::max<long long>(1, min_chunk));
  if (nt <= 1) { fn(0, 0LL, n); return; }
  std::vector<std::thread> th;
  std::vector<std::string> errs(nt);
  for (int t = 0; t < nt; t++) {
    const long long b = n * t / nt, e = n * (t + 1) / nt;
    th.emplace_back([&, t, b, e] { try { fn(t, b, e); } catch (const std::exception &x) { errs[t] = x.what(); } });
  }
  for (auto &x : th) x.join();
  for (auto &m : errs) if (!m.empty()) throw IoError(m);
}
int chunk_count(long long n, long long min_chunk) {
  int nt = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("UCGB200_IO_THREADS")) nt = atoi(e);
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if ((long long)nt * min_chunk > n) nt = (int)std::max<long long>(1, n / std::max<long long>(1, min_chunk));
  return nt;
}

// ------------------------------------------------------------------------------------ columns
struct Keyword { const char *name; int code; };
const Keyword DUMP_KEYWORDS[] = {
    {"id", UCGB200_COL_ID}, {"mol", UCGB200_COL_MOL}, {"type", UCGB200_COL_TYPE}, {"mass", UCGB200_COL_MASS},
    {"x", UCGB200_COL_X}, {"y", UCGB200_COL_Y}, {"z", UCGB200_COL_Z}, {"xs", UCGB200_COL_XS}, {"ys", UCGB200_COL_YS},
    {"zs", UCGB200_COL_ZS}, {"vx", UCGB200_COL_VX}, {"vy", UCGB200_COL_VY}, {"vz", UCGB200_COL_VZ},
    {"fx", UCGB200_COL_FX}, {"fy", UCGB200_COL_FY}, {"fz", UCGB200_COL_FZ}, {"q", UCGB200_COL_Q},
    {"proc", UCGB200_COL_PROC}, {"ucgstate", UCGB200_COL_UCGSTATE}, {"ucgl", UCGB200_COL_UCGL}, {"ucgp", UCGB200_COL_UCGP},
    // not dump keywords of the reference (there they need `compute property/atom`): the remaining UCG per-site arrays, directly
    {"ucgforce", UCGB200_COL_P_UCGFORCE}, {"ucgvl", UCGB200_COL_P_UCGVL}, {"ucgml", UCGB200_COL_P_UCGML}};
// AtomVecUCG::property_atom (atom_vec_ucg.cpp:172-181)
const Keyword PROPERTY_NAMES[] = {{"ucgstate", UCGB200_COL_P_UCGSTATE}, {"ucgl", UCGB200_COL_P_UCGL},
                                  {"ucgforce", UCGB200_COL_P_UCGFORCE}, {"ucgvl", UCGB200_COL_P_UCGVL},
                                  {"ucgp", UCGB200_COL_P_UCGP},         {"ucgml", UCGB200_COL_P_UCGML}};
int keyword_code(const std::string &w) {
  for (const Keyword &k : DUMP_KEYWORDS) if (w == k.name) return k.code;
  return -1;
}
bool code_is_int(int code) {   // vtype Dump::INT in parse_fields
  return code == UCGB200_COL_ID || code == UCGB200_COL_MOL || code == UCGB200_COL_TYPE || code == UCGB200_COL_PROC ||
         code == UCGB200_COL_UCGSTATE;
}

}  // namespace

// ------------------------------------------------------------------------------------ dump custom
struct ucgb200_dump {
  std::string id, filename;
  int groupbit = 1, nevery = 1;
  bool multifile = false, multiproc = false;
  // columns
  std::vector<std::string> colnames;         // as typed (ITEM: ATOMS line)
  std::vector<int> cols, colbit;             // device codes; -1 while a c_ID[k] column is unbound
  struct ComputeRef { std::string id; int index; };   // c_ID (index 0) or c_ID[k] (k >= 1)
  std::vector<ComputeRef> compute_ref;       // per column, id empty for plain keywords
  // dump_modify state (defaults of [stock] Dump::Dump and DumpCustom::DumpCustom)
  int header_flag = 1, append_flag = 0, flush_flag = 1, time_flag = 0, unit_flag = 0, unit_count = 0, padflag = 0,
      sort_flag = 0, buffer_flag = 1;
  std::string format_line_user, format_int_user, format_float_user;
  std::vector<std::string> format_column_user;
  std::vector<int> tcol, top;
  std::vector<double> tval;
  // run state
  FILE *fp = nullptr;
  bool opened = false;
  long long last_rows = 0, last_bytes = 0, last_step = -1;
  int device_format = 1;   // rows formatted on the device whenever every format is the default one
  Staging rows, text;   // page-locked: the D2H copies of the packed rows / the formatted text land here

  ~ucgb200_dump() { if (fp) fclose(fp); }

  bool default_formats() const {
    if (!format_line_user.empty() || !format_int_user.empty() || !format_float_user.empty()) return false;
    for (const std::string &f : format_column_user) if (!f.empty()) return false;
    return true;
  }
  // init_style(): per-column format = column > int/float > line > default, blank-terminated but the last
  std::vector<std::string> vformats() const {
    std::vector<std::string> words;
    if (!format_line_user.empty()) words = split_words(format_line_user);
    else for (int code : cols) words.push_back(code_is_int(code) ? "%d" : "%g");
    if (words.size() < cols.size()) throw IoError("Dump_modify format line is too short: " + format_line_user);
    std::vector<std::string> vf(cols.size());
    for (size_t i = 0; i < cols.size(); i++) {
      if (!format_column_user[i].empty()) vf[i] = format_column_user[i];
      else if (code_is_int(cols[i]) && !format_int_user.empty()) vf[i] = format_int_user;
      else if (!code_is_int(cols[i]) && !format_float_user.empty()) vf[i] = format_float_user;
      else vf[i] = words[i];
      if (i + 1 < cols.size()) vf[i] += " ";
    }
    return vf;
  }
};

extern "C" int ucgb200_host_dump_create(int narg, const char *const *arg, int groupbit, ucgb200_dump **out, char *errbuf,
                                        int errlen) {
  if (!out) return -1;
  try {
    // dump ID group style N file args
    if (narg < 5) throw IoError("Illegal dump command: missing argument(s)");
    if (std::string(arg[2]) != "custom") throw IoError(std::string("Unrecognized dump style '") + arg[2] + "'");
    if (narg == 5) throw IoError("No dump custom arguments specified");
    auto d = new ucgb200_dump();
    std::unique_ptr<ucgb200_dump> guard(d);
    d->id = arg[0];
    d->groupbit = groupbit;
    d->nevery = (int)inumeric(arg[3]);
    if (d->nevery <= 0) throw IoError("Illegal dump custom command: output frequency must be > 0");
    d->filename = arg[4];
    d->multiproc = d->filename.find('%') != std::string::npos;   // one file per brick, as `dump ... file.%` gives one per processor
    if (d->filename.size() > 4 && d->filename.compare(d->filename.size() - 4, 4, ".bin") == 0)
      throw IoError("Dump custom: binary files are not supported by the device writer");
    d->multifile = d->filename.find('*') != std::string::npos;
    if (narg - 5 > UCGB200_DUMP_MAXCOL) throw IoError("Dump custom: too many columns for the device writer");
    for (int k = 5; k < narg; k++) {
      std::string w = arg[k];
      ucgb200_dump::ComputeRef ref{"", 0};
      int code = keyword_code(w);
      if (code < 0) {
        if (w.size() > 2 && w[0] == 'c' && w[1] == '_') {   // c_ID or c_ID[k]: bound later, as init_style() looks computes up
          size_t b = w.find('[');
          ref.id = w.substr(2, b == std::string::npos ? std::string::npos : b - 2);
          if (b != std::string::npos) {
            size_t e = w.find(']', b);
            if (e == std::string::npos) throw IoError("Invalid attribute " + w + " in dump custom command");
            ref.index = (int)inumeric(w.substr(b + 1, e - b - 1));
            if (ref.index < 1) throw IoError("Invalid attribute " + w + " in dump custom command");
          }
        } else
          throw IoError("Invalid attribute " + w + " in dump custom command");
      }
      d->colnames.push_back(w);
      d->cols.push_back(code);
      d->colbit.push_back(~0);
      d->compute_ref.push_back(ref);
    }
    d->format_column_user.assign(d->cols.size(), "");
    const char *env = getenv("UCGB200_DUMP_DEVICE_FORMAT");
    if (env) d->device_format = atoi(env);
    *out = guard.release();
    return 0;
  } catch (const std::exception &e) {
    return fail(errbuf, errlen, e.what());
  }
}

extern "C" void ucgb200_host_dump_free(ucgb200_dump *d) { delete d; }

// `compute ID group property/atom name...`: binds the c_ID / c_ID[k] columns of the dump
extern "C" int ucgb200_host_dump_bind_compute(ucgb200_dump *d, const char *id, int groupbit, int nvalues,
                                              const char *const *names, char *errbuf, int errlen) {
  if (!d || !id || nvalues < 1) return -1;
  try {
    std::vector<int> codes;
    for (int k = 0; k < nvalues; k++) {
      int code = -1;
      for (const Keyword &p : PROPERTY_NAMES) if (std::string(names[k]) == p.name) code = p.code;
      if (code < 0) throw IoError(std::string("Invalid keyword ") + names[k] + " for atom style in compute property/atom command");
      codes.push_back(code);
    }
    for (size_t c = 0; c < d->cols.size(); c++) {
      const auto &ref = d->compute_ref[c];
      if (ref.id != id) continue;
      // DumpCustom::parse_fields: c_ID needs a per-atom vector, c_ID[k] a per-atom array with k <= columns
      if (ref.index == 0 && nvalues != 1) throw IoError("Dump custom compute " + ref.id + " does not calculate per-atom vector");
      if (ref.index > 0 && nvalues == 1) throw IoError("Dump custom compute " + ref.id + " does not calculate per-atom array");
      if (ref.index > nvalues) throw IoError("Dump custom compute " + ref.id + " vector is accessed out-of-range");
      d->cols[c] = codes[ref.index == 0 ? 0 : ref.index - 1];
      d->colbit[c] = groupbit;
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(errbuf, errlen, e.what());
  }
}

extern "C" int ucgb200_host_dump_modify(ucgb200_dump *d, int narg, const char *const *arg, char *errbuf, int errlen) {
  if (!d) return -1;
  try {
    int iarg = 0;
    auto need = [&](int n, const char *kw) {
      if (iarg + n > narg) throw IoError(std::string("Illegal dump_modify ") + kw + " command: missing argument(s)");
    };
    while (iarg < narg) {
      const std::string kw = arg[iarg];
      if (kw == "append") { need(2, "append"); d->append_flag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "buffer") { need(2, "buffer"); d->buffer_flag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "flush") { need(2, "flush"); d->flush_flag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "header") { need(2, "header"); d->header_flag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "pad") { need(2, "pad"); d->padflag = (int)inumeric(arg[iarg + 1]); iarg += 2; }
      else if (kw == "time") { need(2, "time"); d->time_flag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "units") { need(2, "units"); d->unit_flag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "every") { need(2, "every"); d->nevery = (int)inumeric(arg[iarg + 1]); if (d->nevery <= 0) throw IoError("Illegal dump_modify command"); iarg += 2; }
      else if (kw == "sort") {
        need(2, "sort");
        const std::string v = arg[iarg + 1];
        if (v == "off") d->sort_flag = 0;
        else if (v == "id") d->sort_flag = 1;
        else throw IoError("Dump_modify sort by column is not supported by the device writer (use 'id' or 'off')");
        iarg += 2;
      } else if (kw == "format") {
        need(2, "format");
        const std::string what = arg[iarg + 1];
        if (what == "none") {
          d->format_line_user.clear(); d->format_int_user.clear(); d->format_float_user.clear();
          d->format_column_user.assign(d->cols.size(), "");
          iarg += 2;
          continue;
        }
        need(3, "format");
        if (what == "line") d->format_line_user = arg[iarg + 2];
        else if (what == "int") {
          if (!strchr(arg[iarg + 2], 'd')) throw IoError("Dump_modify int format does not contain d character");
          d->format_int_user = arg[iarg + 2];
        } else if (what == "float") d->format_float_user = arg[iarg + 2];
        else {
          long long i = inumeric(what) - 1;
          if (i < 0 || i >= (long long)d->cols.size()) throw IoError("Unknown dump_modify format ID keyword: " + what);
          d->format_column_user[i] = arg[iarg + 2];
        }
        iarg += 3;
      } else if (kw == "thresh") {
        need(2, "thresh");
        if (std::string(arg[iarg + 1]) == "none") { d->tcol.clear(); d->top.clear(); d->tval.clear(); iarg += 2; continue; }
        need(4, "thresh");
        int code = keyword_code(arg[iarg + 1]);
        if (code < 0) throw IoError(std::string("Invalid dump_modify thresh attribute: ") + arg[iarg + 1]);
        static const char *ops[] = {"<", "<=", ">", ">=", "==", "!=", "|^"};
        int op = -1;
        for (int k = 0; k < 7; k++) if (std::string(arg[iarg + 2]) == ops[k]) op = k;
        if (op < 0) throw IoError("Invalid dump_modify thresh operator");
        if (std::string(arg[iarg + 3]) == "LAST") throw IoError("Dump_modify thresh LAST is not supported by the device writer");
        if ((int)d->tcol.size() >= UCGB200_DUMP_MAXTHRESH) throw IoError("Dump_modify: too many thresholds for the device writer");
        d->tcol.push_back(code); d->top.push_back(op); d->tval.push_back(numeric(arg[iarg + 3]));
        iarg += 4;
      } else
        throw IoError("Unknown dump_modify keyword: " + kw);
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(errbuf, errlen, e.what());
  }
}

namespace {

void dump_open(ucgb200_dump *d, ucgb200_ctx *ctx, long long ntimestep) {
  if (d->opened && !d->multifile) return;
  std::string name = d->filename;
  int rank = 0, nranks = 1;
  ucgb200_halo_info(ctx, &rank, &nranks);
  if (d->multiproc) name.replace(name.find('%'), 1, std::to_string(rank));
  else if (nranks > 1) throw IoError("Dump custom: a multi-brick run writes one file per brick (use '%' in the file name)");
  if (d->multifile) {   // utils::star_subst with dump_modify pad
    size_t star = name.find('*');
    char num[64];
    snprintf(num, sizeof num, "%0*lld", d->padflag, ntimestep);
    name = name.substr(0, star) + num + name.substr(star + 1);
  }
  d->fp = fopen(name.c_str(), d->append_flag ? "a" : "w");
  if (!d->fp) throw IoError("Cannot open dump file " + name + ": " + strerror(errno));
  d->opened = true;
}

// DumpCustom::header_item (dump_custom.cpp:651-672)
void dump_header(ucgb200_dump *d, ucgb200_ctx *ctx, long long ntimestep, double time, const char *unit_style, long long ndump) {
  double lo[3], hi[3];
  int per[3];
  if (ucgb200_get_box(ctx, lo, hi, per)) throw IoError("dump: no box");
  FILE *fp = d->fp;
  if (d->unit_flag && !d->unit_count) {
    ++d->unit_count;
    fprintf(fp, "ITEM: UNITS\n%s\n", unit_style ? unit_style : "lj");
  }
  if (d->time_flag) fprintf(fp, "ITEM: TIME\n%.16g\n", time);
  fprintf(fp, "ITEM: TIMESTEP\n%lld\nITEM: NUMBER OF ATOMS\n%lld\n", ntimestep, ndump);
  char bound[9];
  for (int k = 0; k < 3; k++) { bound[3 * k] = bound[3 * k + 1] = per[k] ? 'p' : 'f'; bound[3 * k + 2] = ' '; }
  bound[8] = '\0';
  fprintf(fp, "ITEM: BOX BOUNDS %s\n%1.16e %1.16e\n%1.16e %1.16e\n%1.16e %1.16e\n", bound, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
  std::string columns;
  for (size_t c = 0; c < d->colnames.size(); c++) columns += (c ? " " : "") + d->colnames[c];
  fprintf(fp, "ITEM: ATOMS %s\n", columns.c_str());
}

}  // namespace

// Dump::write() for one snapshot: header, then the rows — formatted on the device when every format is the
// default one, otherwise packed on the device and formatted here exactly as convert_string() does
extern "C" int ucgb200_host_dump_write(ucgb200_dump *d, ucgb200_ctx *ctx, long long ntimestep, double time,
                                       const char *unit_style, char *errbuf, int errlen) {
  if (!d || !ctx) return -1;
  try {
    for (size_t c = 0; c < d->cols.size(); c++)
      if (d->cols[c] < 0) throw IoError("Could not find dump custom compute ID: " + d->compute_ref[c].id);
    ucgb200_dump_spec sp;
    sp.ncols = (int)d->cols.size();
    sp.cols = d->cols.data();
    sp.col_groupbit = d->colbit.data();
    sp.groupbit = d->groupbit;
    sp.nthresh = (int)d->tcol.size();
    sp.thresh_col = d->tcol.data(); sp.thresh_op = d->top.data(); sp.thresh_value = d->tval.data();
    sp.order = d->sort_flag ? UCGB200_DUMP_ORDER_ID : UCGB200_DUMP_ORDER_INDEX;
    dump_open(d, ctx, ntimestep);
    long long nrows = 0, nbytes = 0;
    const bool on_device = d->device_format && d->default_formats();
    const char *body = nullptr;
    std::vector<std::string> parts;
    if (on_device) {
      if (ucgb200_dump_text(ctx, &sp, nullptr, 0, &nrows, &nbytes)) throw IoError("dump: " + ctx_error(ctx));
      if (d->header_flag) dump_header(d, ctx, ntimestep, time, unit_style, nrows);
      char *text = d->text.ensure((size_t)nbytes + 1);
      if (nbytes && ucgb200_dump_text_copy(ctx, text, (long long)d->text.cap)) throw IoError("dump: " + ctx_error(ctx));
      body = text;
    } else {
      int nl = 0;
      ucgb200_natoms(ctx, &nl, nullptr);
      const double *buf = (const double *)d->rows.ensure(((size_t)nl * sp.ncols + 1) * sizeof(double));
      if (ucgb200_dump_pack(ctx, &sp, (double *)d->rows.p, nl, &nrows)) throw IoError("dump: " + ctx_error(ctx));
      if (d->header_flag) dump_header(d, ctx, ntimestep, time, unit_style, nrows);
      const std::vector<std::string> vf = d->vformats();
      // convert_string(): snprintf per field, "\n" per row — rows are independent, so chunks of them go to the host cores
      const int nc = sp.ncols;
      parts.assign(chunk_count(nrows, 4096), std::string());
      parallel_chunks(nrows, 4096, [&](int t, long long b, long long e) {
        std::string &out = parts[t];
        out.reserve((size_t)(e - b) * nc * 12 + 64);
        char field[512];
        for (long long r = b; r < e; r++) {
          const double *row = buf + (size_t)r * nc;
          for (int c = 0; c < nc; c++) {
            int len = code_is_int(d->cols[c]) ? snprintf(field, sizeof field, vf[c].c_str(), static_cast<int>(row[c]))
                                              : snprintf(field, sizeof field, vf[c].c_str(), row[c]);
            out.append(field, (size_t)std::min<int>(len, (int)sizeof field - 1));
          }
          out.push_back('\n');
        }
      });
      for (const std::string &p : parts) nbytes += (long long)p.size();
    }
    if (body && nbytes) fwrite(body, 1, (size_t)nbytes, d->fp);
    for (const std::string &p : parts) if (!p.empty()) fwrite(p.data(), 1, p.size(), d->fp);
    if (d->flush_flag) fflush(d->fp);
    if (d->multifile) { fclose(d->fp); d->fp = nullptr; }
    d->last_rows = nrows;
    d->last_bytes = nbytes;
    return 0;
  } catch (const std::exception &e) {
    return fail(errbuf, errlen, e.what());
  }
}

extern "C" int ucgb200_host_run(ucgb200_ctx *ctx, long long nsteps, int ndump, ucgb200_dump *const *dumps, double dt,
                                const char *unit_style, char *errbuf, int errlen) {
  if (!ctx || nsteps < 0 || ndump < 0 || (ndump && !dumps)) return -1;
  try {
    double th[16];
    if (ucgb200_thermo(ctx, th)) throw IoError("run: " + ctx_error(ctx));
    long long step = (long long)th[10];
    const long long begin = step, end = step + nsteps;
    auto write_due = [&]() {
      for (int k = 0; k < ndump; k++) {
        if (step % dumps[k]->nevery) continue;
        if (dumps[k]->last_step == step && dumps[k]->opened) continue;   // already written at the end of the previous run
        if (ucgb200_host_dump_write(dumps[k], ctx, step, (double)step * dt, unit_style, errbuf, errlen)) throw IoError(errbuf);
        dumps[k]->last_step = step;
      }
    };
    write_due();
    while (step < end) {
      long long next = end;
      for (int k = 0; k < ndump; k++) next = std::min(next, (step / dumps[k]->nevery + 1) * dumps[k]->nevery);
      const int n = (int)(next - step);
      const int rc = ucgb200_run_between(ctx, n, begin, end);
      if (rc) throw IoError("run: " + ctx_error(ctx) + " (rc " + std::to_string(rc) + ")");
      step = next;
      write_due();
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(errbuf, errlen, e.what());
  }
}

extern "C" int ucgb200_host_dump_stats(const ucgb200_dump *d, long long *rows, long long *bytes, int *nevery) {
  if (!d) return -1;
  if (rows) *rows = d->last_rows;
  if (bytes) *bytes = d->last_bytes;
  if (nevery) *nevery = d->nevery;
  return 0;
}

// ------------------------------------------------------------------------------------ read_dump
namespace {

// Reader field types (reader.h:24-26) as device column codes; ix iy iz cannot be honoured (no image flags on the device)
int read_dump_fieldtype(const std::string &w) {
  static const Keyword f[] = {{"id", UCGB200_COL_ID}, {"type", UCGB200_COL_TYPE}, {"x", UCGB200_COL_X}, {"y", UCGB200_COL_Y},
                              {"z", UCGB200_COL_Z}, {"vx", UCGB200_COL_VX}, {"vy", UCGB200_COL_VY}, {"vz", UCGB200_COL_VZ},
                              {"q", UCGB200_COL_Q}, {"fx", UCGB200_COL_FX}, {"fy", UCGB200_COL_FY}, {"fz", UCGB200_COL_FZ},
                              {"ucgstate", UCGB200_COL_UCGSTATE}, {"ucgl", UCGB200_COL_UCGL}, {"ucgp", UCGB200_COL_UCGP}};
  for (const Keyword &k : f) if (w == k.name) return k.code;
  if (w == "ix" || w == "iy" || w == "iz") return -2;
  return -1;
}

enum { UNSET = 0, NOSCALE_NOWRAP, NOSCALE_WRAP, SCALE_NOWRAP, SCALE_WRAP };   // reader.h

struct Snapshot {
  long long ntimestep = 0, natoms = 0;
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  bool triclinic = false;
  std::vector<std::string> labels;
};

bool read_line(FILE *fp, std::string &line) {
  char buf[1024];
  line.clear();
  while (fgets(buf, sizeof buf, fp)) {
    line += buf;
    if (!line.empty() && line.back() == '\n') return true;
  }
  return !line.empty();
}
std::string trimmed(const std::string &s) {
  size_t b = s.find_first_not_of(" \t\r\n");
  if (b == std::string::npos) return "";
  return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
}
void need_line(FILE *fp, std::string &line) {
  if (!read_line(fp, line)) throw IoError("Unexpected end of dump file");
}

// ReaderNative::read_time, text: returns false at end of file
bool read_time(FILE *fp, Snapshot &s) {
  std::string line;
  if (!read_line(fp, line)) return false;
  if (trimmed(line) == "ITEM: UNITS") { need_line(fp, line); need_line(fp, line); }
  if (trimmed(line) == "ITEM: TIME") { need_line(fp, line); need_line(fp, line); }
  if (trimmed(line) != "ITEM: TIMESTEP") throw IoError("Dump file is incorrectly formatted");
  need_line(fp, line);
  s.ntimestep = inumeric(trimmed(line));
  return true;
}
// ReaderNative::read_header, text: atom count, box, column labels
void read_header(FILE *fp, Snapshot &s) {
  std::string line;
  need_line(fp, line);   // ITEM: NUMBER OF ATOMS
  need_line(fp, line);
  s.natoms = inumeric(trimmed(line));
  need_line(fp, line);
  if (line.find("ITEM: BOX BOUNDS") == std::string::npos) throw IoError("Dump file is incorrectly formatted");
  s.triclinic = line.find("xy") != std::string::npos || line.find("abc") != std::string::npos;
  for (int d = 0; d < 3; d++) {
    need_line(fp, line);
    std::vector<std::string> w = split_words(line);
    if (w.size() < 2) throw IoError("Dump file is incorrectly formatted");
    s.lo[d] = numeric(w[0]); s.hi[d] = numeric(w[1]);
  }
  need_line(fp, line);
  const std::string key = "ITEM: ATOMS";
  if (line.compare(0, key.size(), key) != 0) throw IoError("Dump file is incorrectly formatted");
  s.labels = split_words(line.substr(key.size()));
}
// ReaderNative::skip, text
void skip_snapshot(FILE *fp) {
  Snapshot s;
  read_header(fp, s);
  std::string line;
  for (long long i = 0; i < s.natoms; i++) need_line(fp, line);
}

}  // namespace

// read_dump file Nstep field ... keyword value ...   (serial, native text files)
// stats[7]: atoms before read, in snapshot, purged, replaced, trimmed, added, after read (the reference's log lines)
extern "C" int ucgb200_host_read_dump(ucgb200_ctx *ctx, int narg, const char *const *arg, long long stats[7], char *errbuf,
                                      int errlen) {
  if (!ctx) return -1;
  FILE *fp = nullptr;
  try {
    if (narg < 2) throw IoError("Illegal read_dump command: missing argument(s)");
    const std::string file = arg[0];
    const long long nstep = inumeric(arg[1]);
    // fields_and_keywords()
    std::vector<int> fieldtype{UCGB200_COL_ID};
    std::vector<std::string> fieldlabel{""};
    int iarg = 2;
    while (iarg < narg) {
      int t = read_dump_fieldtype(arg[iarg]);
      if (t == -2) throw IoError("read_dump ix/iy/iz: the device keeps no image flags");
      if (t < 0) break;
      fieldtype.push_back(t);
      fieldlabel.push_back("");
      iarg++;
    }
    if (fieldtype.size() == 1) throw IoError("Read_dump must use at least either 'id' or 'type' field");
    for (size_t i = 0; i < fieldtype.size(); i++)
      for (size_t j = i + 1; j < fieldtype.size(); j++)
        if (fieldtype[i] == fieldtype[j]) throw IoError("Duplicate fields in read_dump command");
    int boxflag = 1, replaceflag = 1, purgeflag = 0, trimflag = 0, scaleflag = 0, wrapflag = 1, timestepflag = 1;
    auto need = [&](int n, const char *kw) {
      if (iarg + n > narg) throw IoError(std::string("Illegal read_dump ") + kw + " command: missing argument(s)");
    };
    while (iarg < narg) {
      const std::string kw = arg[iarg];
      if (kw == "box") { need(2, "box"); boxflag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "timestep") { need(2, "timestep"); timestepflag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "replace") { need(2, "replace"); replaceflag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "purge") { need(2, "purge"); purgeflag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "trim") { need(2, "trim"); trimflag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "add") {
        need(2, "add");
        const std::string v = arg[iarg + 1];
        if (v == "yes" || v == "true" || v == "keep") throw IoError("read_dump add yes/keep is not supported (the reference notes the same for UCG, read_dump.cpp:954)");
        if (v != "no" && v != "false") throw IoError("Unknown read_dump add keyword " + v);
        iarg += 2;
      } else if (kw == "label") {
        need(3, "label");
        int t = read_dump_fieldtype(arg[iarg + 1]);
        size_t i = 0;
        for (; i < fieldtype.size(); i++) if (fieldtype[i] == t) break;
        if (t < 0 || i == fieldtype.size()) throw IoError("Illegal read_dump command");
        fieldlabel[i] = arg[iarg + 2];
        iarg += 3;
      } else if (kw == "scaled") { need(2, "scaled"); scaleflag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "wrapped") { need(2, "wrapped"); wrapflag = logical(arg[iarg + 1]); iarg += 2; }
      else if (kw == "format") {
        need(2, "format");
        if (std::string(arg[iarg + 1]) != "native") throw IoError(std::string("Unrecognized reader style '") + arg[iarg + 1] + "'");
        iarg += 2;
      } else
        throw IoError("Unknown read_dump keyword: " + kw);
    }
    (void)timestepflag;
    if (purgeflag && (replaceflag || trimflag)) throw IoError("If read_dump purges it cannot replace or trim");
    if (purgeflag) throw IoError("read_dump purge yes needs add yes, which is not supported");

    // seek(nstep, exact)
    fp = fopen(file.c_str(), "r");
    if (!fp) throw IoError("Cannot open file " + file + ": " + strerror(errno));
    Snapshot snap;
    bool found = false;
    while (read_time(fp, snap)) {
      if (snap.ntimestep == nstep) { found = true; break; }
      if (snap.ntimestep > nstep) break;
      skip_snapshot(fp);
    }
    if (!found) throw IoError("Dump file does not contain requested snapshot");
    read_header(fp, snap);
    if (boxflag && snap.triclinic) throw IoError("Read_dump triclinic status does not match simulation");

    // ReaderNative::read_header: match every field with a column; x y z fall back to the scaled / unwrapped variants,
    // the first one present in the file wins
    std::map<std::string, int> label;
    for (size_t k = 0; k < snap.labels.size(); k++) label[snap.labels[k]] = (int)k;
    auto find = [&](const std::string &l) { auto it = label.find(l); return it == label.end() ? -1 : it->second; };
    const int nwords = (int)snap.labels.size();
    std::vector<int> fieldindex(fieldtype.size(), -1);
    int xyzflag[3] = {UNSET, UNSET, UNSET};
    for (size_t i = 0; i < fieldtype.size(); i++) {
      const int t = fieldtype[i];
      const int dim = t == UCGB200_COL_X ? 0 : (t == UCGB200_COL_Y ? 1 : (t == UCGB200_COL_Z ? 2 : -1));
      if (!fieldlabel[i].empty()) {
        fieldindex[i] = find(fieldlabel[i]);
        if (dim >= 0) xyzflag[dim] = 2 * scaleflag + wrapflag + 1;
      } else if (dim >= 0) {
        const std::string base(1, "xyz"[dim]);
        fieldindex[i] = find(base);
        xyzflag[dim] = NOSCALE_WRAP;
        if (fieldindex[i] < 0) {
          fieldindex[i] = nwords;
          const int s = find(base + "s"), u = find(base + "u"), su = find(base + "su");
          if (s >= 0 && s < fieldindex[i]) { fieldindex[i] = s; xyzflag[dim] = SCALE_WRAP; }
          if (u >= 0 && u < fieldindex[i]) { fieldindex[i] = u; xyzflag[dim] = NOSCALE_NOWRAP; }
          if (su >= 0 && su < fieldindex[i]) { fieldindex[i] = su; xyzflag[dim] = SCALE_NOWRAP; }
        }
        if (fieldindex[i] == nwords) fieldindex[i] = -1;
      } else {
        static const Keyword names[] = {{"id", UCGB200_COL_ID}, {"type", UCGB200_COL_TYPE}, {"vx", UCGB200_COL_VX}, {"vy", UCGB200_COL_VY},
                                        {"vz", UCGB200_COL_VZ}, {"q", UCGB200_COL_Q}, {"fx", UCGB200_COL_FX}, {"fy", UCGB200_COL_FY},
                                        {"fz", UCGB200_COL_FZ}, {"ucgstate", UCGB200_COL_UCGSTATE}, {"ucgl", UCGB200_COL_UCGL},
                                        {"ucgp", UCGB200_COL_UCGP}};
        for (const Keyword &k : names) if (k.code == t) fieldindex[i] = find(k.name);
      }
    }
    for (int fi : fieldindex) if (fi < 0) throw IoError("One of the requested read_dump per-atom fields not found in dump file");
    int value = std::max(xyzflag[0], std::max(xyzflag[1], xyzflag[2]));
    for (int d = 0; d < 3; d++)
      if (xyzflag[d] != UNSET && xyzflag[d] != value) throw IoError("Read_dump xyz fields do not have consistent scaling/wrapping");
    const int scaled = (value == SCALE_NOWRAP || value == SCALE_WRAP) ? 1 : 0;

    // ReaderNative::read_atoms: the snapshot body is read in one block, lines are located, and chunks of lines are
    // tokenised and converted (strtod == std::stod) on the host cores
    // The snapshot body is read in one block (page-locked, so the upload is a DMA) and converted on the device
    // (ucgb200_snapshot_parse); rows holding a token the device does not convert exactly, or the whole block when the
    // device declines, are tokenised and converted (strtod == std::stod) on the host cores.
    const int nfield = (int)fieldtype.size();
    std::vector<double> fields;
    bool fields_on_device = false;
    {
      const long body0 = ftell(fp);
      fseek(fp, 0, SEEK_END);
      const long fend = ftell(fp);
      fseek(fp, body0, SEEK_SET);
      static thread_local Staging textbuf;   // page-locking 74 MB costs more than copying them: kept across calls
      char *text = textbuf.ensure((size_t)(fend - body0) + 1);
      const size_t got = fread(text, 1, (size_t)(fend - body0), fp);
      text[got] = '\0';
      auto parse_line = [&](const char *p, const char *end, double *out, std::vector<const char *> &tok) {
        tok.clear();
        while (p < end) {
          while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n' || *p == '\f')) p++;
          if (p >= end) break;
          tok.push_back(p);
          while (p < end && !(*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n' || *p == '\f')) p++;
        }
        if ((int)tok.size() < nwords) throw IoError("Insufficient columns in dump file");
        for (int m = 0; m < nfield; m++) out[m] = strtod(tok[fieldindex[m]], nullptr);
      };
      const char *env = getenv("UCGB200_READ_DUMP_DEVICE_PARSE");
      int rc = UCGB200_PARSE_ON_HOST;
      // only the lines of THIS snapshot: later snapshots may follow in the file
      size_t body_len = got;
      if (!(env && atoi(env) == 0)) {
        size_t pos = 0;
        for (long long i = 0; i < snap.natoms && pos < got; i++) {
          const char *nl = (const char *)memchr(text + pos, '\n', got - pos);
          pos = nl ? (size_t)(nl - text) + 1 : got;
          if (i + 1 == snap.natoms) body_len = pos;
        }
        const int cap = 1 << 16;
        std::vector<int> srow(cap), soff(cap);
        long long nslow = 0;
        rc = ucgb200_snapshot_parse(ctx, text, (long long)body_len, snap.natoms, nwords, nfield, fieldindex.data(), cap, srow.data(),
                                    soff.data(), &nslow);
        if (rc < 0) throw IoError(ctx_error(ctx));
        if (rc == 0) {
          if (nslow) {
            std::vector<double> vals((size_t)nslow * nfield);
            std::vector<const char *> tok;
            for (long long k = 0; k < nslow; k++) {
              const char *p = text + soff[k];
              const char *nl = (const char *)memchr(p, '\n', body_len - soff[k]);
              parse_line(p, nl ? nl + 1 : text + body_len, &vals[(size_t)k * nfield], tok);
            }
            if (ucgb200_snapshot_patch(ctx, (int)nslow, srow.data(), vals.data())) throw IoError(ctx_error(ctx));
          }
          fields_on_device = true;
        }
      }
      if (!fields_on_device) {
        fields.resize((size_t)snap.natoms * nfield);
        std::vector<size_t> start((size_t)snap.natoms + 1);
        size_t pos = 0;
        for (long long i = 0; i < snap.natoms; i++) {
          if (pos >= got) throw IoError("Unexpected end of dump file");
          start[i] = pos;
          const char *nl = (const char *)memchr(text + pos, '\n', got - pos);
          pos = nl ? (size_t)(nl - text) + 1 : got;
        }
        start[snap.natoms] = pos;
        parallel_chunks(snap.natoms, 2048, [&](int, long long b, long long e) {
          std::vector<const char *> tok;
          for (long long i = b; i < e; i++) parse_line(text + start[i], text + start[i + 1], &fields[(size_t)i * nfield], tok);
        });
      }
    }
    fclose(fp);
    fp = nullptr;

    int nbefore = 0;
    ucgb200_natoms(ctx, &nbefore, nullptr);
    // box yes: the snapshot box replaces the simulation box before the remap (read_dump.cpp:636-661)
    if (boxflag) {
      int per[3];
      ucgb200_get_box(ctx, nullptr, nullptr, per);
      if (ucgb200_set_box(ctx, snap.lo, snap.hi, per)) throw IoError("read_dump: " + ctx_error(ctx));
    }
    long long nreplace = 0, ntrim = 0;
    std::vector<int> updated((size_t)nbefore + 1, 0);
    if (replaceflag || trimflag) {
      // without `replace` only the update flags are needed: an id-only pass changes nothing
      // `replace no`: the same pass with every field but the id masked out marks the matches and changes nothing
      std::vector<int> ftype = fieldtype;
      if (!replaceflag) for (int j = 1; j < nfield; j++) ftype[j] = UCGB200_COL_Q;
      if (ucgb200_atoms_update_by_tag(ctx, (int)snap.natoms, nfield, ftype.data(), fields_on_device ? nullptr : fields.data(), scaled,
                                      snap.lo, snap.hi, updated.data(), &nreplace))
        throw IoError("read_dump: " + ctx_error(ctx));
      if (!replaceflag) nreplace = 0;
    }
    // migrate_atoms_by_coords: every atom, replaced or not, is wrapped into the (possibly new) box
    if (ucgb200_atoms_remap(ctx)) throw IoError("read_dump: " + ctx_error(ctx));
    int nafter = nbefore;
    if (trimflag) {
      // ReadDump::process_atoms :919-935: `avec->copy(nlocal-1,i)` into every hole, same resulting order
      std::vector<int> src((size_t)nbefore);
      for (int i = 0; i < nbefore; i++) src[i] = i;
      std::vector<int> flag(updated.begin(), updated.begin() + nbefore);
      int nlocal = nbefore, i = 0;
      while (i < nlocal) {
        if (!flag[i]) { src[i] = src[nlocal - 1]; flag[i] = flag[nlocal - 1]; nlocal--; ntrim++; }
        else i++;
      }
      if (ntrim) {
        const size_t n = (size_t)nbefore;
        std::vector<double> x(3 * n), v(3 * n), f(3 * n), ucgl(n), ucgvl(n), ucgml(n), ucgp(n), ucgforce(n), scores(2 * n);
        std::vector<int> type(n), mask(n), tag(n), mol(n), state(n);
        ucgb200_atoms h{x.data(), v.data(), f.data(), type.data(), mask.data(), tag.data(), mol.data(), state.data(), ucgl.data(),
                        ucgvl.data(), ucgml.data(), ucgp.data(), ucgforce.data(), scores.data(), nullptr};
        if (ucgb200_atoms_download(ctx, nbefore, &h, UCGB200_F_ALL & ~UCGB200_F_NUMSTATES)) throw IoError("read_dump: " + ctx_error(ctx));
        auto gather = [&](auto &a, int w) {
          auto b = a;
          for (int k = 0; k < nlocal; k++) for (int c = 0; c < w; c++) a[(size_t)k * w + c] = b[(size_t)src[k] * w + c];
        };
        gather(x, 3); gather(v, 3); gather(f, 3); gather(ucgl, 1); gather(ucgvl, 1); gather(ucgml, 1); gather(ucgp, 1);
        gather(ucgforce, 1); gather(scores, 2); gather(type, 1); gather(mask, 1); gather(tag, 1); gather(mol, 1); gather(state, 1);
        if (ucgb200_atoms_upload(ctx, nlocal, &h, UCGB200_F_ALL & ~UCGB200_F_NUMSTATES)) throw IoError("read_dump: " + ctx_error(ctx));
      }
      nafter = nlocal;
    }
    if (stats) {
      stats[0] = nbefore; stats[1] = snap.natoms; stats[2] = 0; stats[3] = nreplace; stats[4] = ntrim; stats[5] = 0; stats[6] = nafter;
    }
    return 0;
  } catch (const std::exception &e) {
    if (fp) fclose(fp);
    return fail(errbuf, errlen, e.what());
  }
}

// ------------------------------------------------------------------------------------ read_data
// A LAMMPS data file with `Atoms # ucg`: header (atoms, atom types, xlo xhi ...), Masses, Atoms, Velocities.
// Columns are AtomVecUCG's fields_data_atom / fields_data_vel (atom_vec_ucg.cpp:85-90):
//   Atoms       id molecule type q x y z ucgstate ucgl ucgml [ix iy iz]
//   Velocities  id vx vy vz ucgvl
// followed per atom by data_atom_post (:145-170): ucgl clamped to [0,1], ucgstate to {0,1}, ucgp = -1.
struct ucgb200_data {
  long long natoms = 0;
  int ntypes = 0;
  double lo[3] = {-0.5, -0.5, -0.5}, hi[3] = {0.5, 0.5, 0.5};
  std::vector<double> mass, x, v, q, ucgl, ucgvl, ucgml, ucgp;
  std::vector<int> tag, mol, type, state, image, mask;
};

extern "C" int ucgb200_host_data_read(const char *file, ucgb200_data **out, char *errbuf, int errlen) {
  if (!file || !out) return -1;
  FILE *fp = fopen(file, "r");
  if (!fp) return fail(errbuf, errlen, std::string("Cannot open file ") + file + ": " + strerror(errno));
  try {
    std::unique_ptr<ucgb200_data> d(new ucgb200_data());
    std::string line;
    read_line(fp, line);   // title
    auto strip = [](std::string s) { size_t h = s.find('#'); if (h != std::string::npos) s.erase(h); return trimmed(s); };
    // header: keyword lines until the first section name
    std::string section;
    while (read_line(fp, line)) {
      std::string t = strip(line);
      if (t.empty()) continue;
      std::vector<std::string> w = split_words(t);
      if (w.size() == 2 && w[1] == "atoms") d->natoms = inumeric(w[0]);
      else if (w.size() == 3 && w[1] == "atom" && w[2] == "types") d->ntypes = (int)inumeric(w[0]);
      else if (w.size() == 4 && w[2] == "xlo" && w[3] == "xhi") { d->lo[0] = numeric(w[0]); d->hi[0] = numeric(w[1]); }
      else if (w.size() == 4 && w[2] == "ylo" && w[3] == "yhi") { d->lo[1] = numeric(w[0]); d->hi[1] = numeric(w[1]); }
      else if (w.size() == 4 && w[2] == "zlo" && w[3] == "zhi") { d->lo[2] = numeric(w[0]); d->hi[2] = numeric(w[1]); }
      else if (w.size() >= 2 && (w.back() == "bonds" || w.back() == "angles" || w.back() == "dihedrals" || w.back() == "impropers" ||
                                 w.back() == "types")) {
        if (inumeric(w[0]) != 0 && w.back() != "types") throw IoError("read_data: bonded topology is outside the UCG hot path (SURVEY §8)");
      } else if (isalpha((unsigned char)w[0][0])) {
        section = w[0];
        break;
      } else
        throw IoError("Unknown identifier in data file: " + t);
    }
    if (d->natoms <= 0) throw IoError("No atoms in data file");
    if (d->ntypes <= 0) throw IoError("No atom types in data file");
    const size_t n = (size_t)d->natoms;
    d->mass.assign(d->ntypes + 1, 0.0);
    std::map<int, size_t> index_of;
    bool have_atoms = false;
    double prd[3];
    for (int k = 0; k < 3; k++) prd[k] = d->hi[k] - d->lo[k];
    while (!section.empty()) {
      read_line(fp, line);   // blank line after the section name
      if (section == "Masses") {
        for (int t = 0; t < d->ntypes; t++) {
          need_line(fp, line);
          std::vector<std::string> w = split_words(strip(line));
          if (w.size() < 2) throw IoError("Invalid format in Masses section of data file");
          int it = (int)inumeric(w[0]);
          if (it < 1 || it > d->ntypes) throw IoError("Invalid type for mass set");
          d->mass[it] = numeric(w[1]);
          if (d->mass[it] <= 0.0) throw IoError("Invalid mass value");
        }
      } else if (section == "Atoms") {
        d->x.resize(3 * n); d->v.assign(3 * n, 0.0); d->q.resize(n); d->ucgl.resize(n); d->ucgvl.assign(n, 0.0); d->ucgml.resize(n);
        d->ucgp.resize(n); d->tag.resize(n); d->mol.resize(n); d->type.resize(n); d->state.resize(n); d->image.resize(n); d->mask.assign(n, 1);
        for (size_t i = 0; i < n; i++) {
          need_line(fp, line);
          std::vector<std::string> w = split_words(strip(line));
          if (w.size() != 10 && w.size() != 13) throw IoError("Incorrect format in Atoms section of data file: " + trimmed(line));
          d->tag[i] = (int)inumeric(w[0]);
          d->mol[i] = (int)inumeric(w[1]);
          d->type[i] = (int)inumeric(w[2]);
          if (d->type[i] <= 0 || d->type[i] > d->ntypes) throw IoError("Invalid atom type in Atoms section of data file");
          d->q[i] = numeric(w[3]);
          int img[3] = {0, 0, 0};
          if (w.size() == 13) for (int k = 0; k < 3; k++) img[k] = (int)inumeric(w[10 + k]);
          for (int k = 0; k < 3; k++) {   // [stock] Atom::data_atoms: Domain::remap into the periodic box
            double c = numeric(w[4 + k]);
            while (c < d->lo[k]) { c += prd[k]; img[k]--; }
            while (c >= d->hi[k]) { c -= prd[k]; img[k]++; }
            c = std::max(c, d->lo[k]);
            d->x[3 * i + k] = c;
          }
          d->image[i] = ((img[0] + 512) & 1023) | (((img[1] + 512) & 1023) << 10) | (((img[2] + 512) & 1023) << 20);
          d->state[i] = (int)inumeric(w[7]);
          d->ucgl[i] = numeric(w[8]);
          d->ucgml[i] = numeric(w[9]);
          // data_atom_post
          if (d->ucgl[i] < 0) d->ucgl[i] = 0.; else if (d->ucgl[i] > 1) d->ucgl[i] = 1.;
          if (d->state[i] < 0) d->state[i] = 0; else if (d->state[i] > 1) d->state[i] = 1;
          d->ucgp[i] = -1.0;
          index_of[d->tag[i]] = i;
        }
        have_atoms = true;
      } else if (section == "Velocities") {
        if (!have_atoms) throw IoError("Must read Atoms before Velocities");
        for (size_t i = 0; i < n; i++) {
          need_line(fp, line);
          std::vector<std::string> w = split_words(strip(line));
          if (w.size() != 5) throw IoError("Incorrect format in Velocities section of data file: " + trimmed(line));
          auto it = index_of.find((int)inumeric(w[0]));
          if (it == index_of.end()) throw IoError("Invalid atom ID in Velocities section of data file");
          for (int k = 0; k < 3; k++) d->v[3 * it->second + k] = numeric(w[1 + k]);
          d->ucgvl[it->second] = numeric(w[4]);
        }
      } else
        throw IoError("Unknown section '" + section + "' in data file (atom_style ucg without bonded topology)");
      // next section name
      section.clear();
      while (read_line(fp, line)) {
        std::string t = strip(line);
        if (t.empty()) continue;
        section = split_words(t)[0];
        break;
      }
    }
    if (!have_atoms) throw IoError("No Atoms section in data file");
    fclose(fp);
    *out = d.release();
    return 0;
  } catch (const std::exception &e) {
    fclose(fp);
    return fail(errbuf, errlen, e.what());
  }
}
extern "C" void ucgb200_host_data_free(ucgb200_data *d) { delete d; }
extern "C" int ucgb200_host_data_info(const ucgb200_data *d, long long *natoms, int *ntypes, double lo[3], double hi[3]) {
  if (!d) return -1;
  if (natoms) *natoms = d->natoms;
  if (ntypes) *ntypes = d->ntypes;
  for (int k = 0; k < 3; k++) { if (lo) lo[k] = d->lo[k]; if (hi) hi[k] = d->hi[k]; }
  return 0;
}
// pointers into the parsed arrays (valid until ucgb200_host_data_free); q, image and mass[1..ntypes] separately
extern "C" int ucgb200_host_data_view(ucgb200_data *d, ucgb200_atoms *view, const double **q, const int **image, const double **mass) {
  if (!d || !view) return -1;
  memset(view, 0, sizeof *view);
  view->x = d->x.data(); view->v = d->v.data(); view->type = d->type.data(); view->mask = d->mask.data(); view->tag = d->tag.data();
  view->molecule = d->mol.data(); view->ucgstate = d->state.data(); view->ucgl = d->ucgl.data(); view->ucgvl = d->ucgvl.data();
  view->ucgml = d->ucgml.data(); view->ucgp = d->ucgp.data();
  if (q) *q = d->q.data();
  if (image) *image = d->image.data();
  if (mass) *mass = d->mass.data();
  return 0;
}
// read_data -> device: box (periodic in all dimensions unless told otherwise) + every data-file field
extern "C" int ucgb200_host_data_upload(ucgb200_ctx *ctx, ucgb200_data *d, const int periodic[3]) {
  if (!ctx || !d) return -1;
  int rc = ucgb200_set_box(ctx, d->lo, d->hi, periodic);
  if (rc) return rc;
  ucgb200_atoms view;
  ucgb200_host_data_view(d, &view, nullptr, nullptr, nullptr);
  return ucgb200_atoms_upload(ctx, (int)d->natoms, &view,
                              UCGB200_F_X | UCGB200_F_V | UCGB200_F_TYPE | UCGB200_F_MASK | UCGB200_F_TAG | UCGB200_F_MOLECULE |
                                  UCGB200_F_UCGSTATE | UCGB200_F_UCGL | UCGB200_F_UCGVL | UCGB200_F_UCGML | UCGB200_F_UCGP);
}
