// lammps_shim_io.h — the part of the minimal LAMMPS-API shim that the reference's patched
// top-level I/O files need (dump_custom.cpp, read_dump.cpp, reader.cpp, reader_native.cpp):
// class Dump (serial write() cycle), Command, Irregular, ArgInfo, FixStoreAtom, LabelMap,
// Region, Tokenizer, platform::, a few utils::.  Written from the public LAMMPS class
// interfaces; nothing is copied.  Test infrastructure only (oracle/_ref).
#pragma once
#include "lammps_shim.h"

#include <algorithm>
#include <numeric>

#define BIGINT_FORMAT "%" PRId64
#define TAGINT_FORMAT "%d"
#include <cinttypes>

// MPI calls only read_dump.cpp makes (one process: never reached or trivially satisfied)
typedef int MPI_Request;
typedef int MPI_Status;
#define MPI_COMM_NULL (-1)
#define MPI_LMP_BIGINT 4
static inline int MPI_Comm_split(MPI_Comm c, int, int, MPI_Comm *out) { *out = c; return 0; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out) { *out = c; return 0; }
static inline int MPI_Comm_free(MPI_Comm *) { return 0; }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
static inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return 0; }
static inline int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
static inline int MPI_Get_count(MPI_Status *, MPI_Datatype, int *n) { *n = 0; return 0; }

namespace fmt {
template <class... A>
inline std::string format(const std::string &f, const A &...a) { return LAMMPS_NS::shimfmt::format(f, a...); }
}  // namespace fmt

namespace LAMMPS_NS {

// ------------------------------------------------------------------------------ utils, platform
namespace utils {
inline std::vector<std::string> split_words(const std::string &text) {
  std::vector<std::string> out;
  std::istringstream is(text);
  std::string w;
  while (is >> w) out.push_back(w);
  return out;
}
template <class... A>
inline void print(FILE *fp, const std::string &f, const A &...a) {
  fputs(shimfmt::format(f, a...).c_str(), fp);
}
inline char *sfgets(const char *file, int line, char *s, int size, FILE *fp, const char *, Error *error) {
  char *rv = fgets(s, size, fp);
  if (rv == nullptr && error) error->one(file, line, "Unexpected end of file while reading");
  return rv;
}
inline int logical(const char *file, int line, const std::string &s, bool, LAMMPS *lmp) {
  if (s == "yes" || s == "on" || s == "true") return 1;
  if (s == "no" || s == "off" || s == "false") return 0;
  lmp->error->all(file, line, "Expected boolean parameter instead of '{}' in input script or data file", s);
  return 0;
}
inline int expand_args(const char *, int, int narg, char **arg, int, char **&earg, LAMMPS *) {
  earg = arg;   // no wildcard expansion in the shim (no compute/fix arrays to expand over)
  return narg;
}
inline bool is_integer(const std::string &s) {
  if (s.empty()) return false;
  size_t k = (s[0] == '-' || s[0] == '+') ? 1 : 0;
  if (k == s.size()) return false;
  for (; k < s.size(); k++) if (!isdigit((unsigned char)s[k])) return false;
  return true;
}
inline bool is_double(const std::string &s) {
  char *e = nullptr;
  strtod(s.c_str(), &e);
  return !s.empty() && e != s.c_str() && *e == '\0';
}
inline std::string star_subst(const std::string &name, bigint step, int pad) {
  size_t star = name.find('*');
  if (star == std::string::npos) return name;
  char num[64];
  snprintf(num, sizeof num, "%0*" PRId64, pad, step);
  return name.substr(0, star) + num + name.substr(star + 1);
}
inline std::string errorurl(int n) { return "\nFor more information see https://docs.lammps.org/err" + std::to_string(10000 + n).substr(1); }
inline std::string check_packages_for_style(const std::string &style, const std::string &name, LAMMPS *) { return "Unrecognized " + style + " style '" + name + "'"; }
inline std::string lowercase(std::string s) { for (auto &ch : s) ch = tolower(ch); return s; }
}  // namespace utils

namespace platform {
inline bool has_compress_extension(const std::string &) { return false; }
inline FILE *compressed_read(const std::string &) { return nullptr; }
inline FILE *compressed_write(const std::string &) { return nullptr; }
inline int pclose(FILE *fp) { return ::pclose(fp); }
inline int fseek(FILE *fp, bigint pos) { return ::fseek(fp, (long)pos, pos < 0 ? SEEK_END : SEEK_SET); }
inline bigint ftell(FILE *fp) { return ::ftell(fp); }
inline bool file_is_readable(const std::string &p) { FILE *f = fopen(p.c_str(), "r"); if (f) fclose(f); return f != nullptr; }
}  // namespace platform

class Tokenizer {
  std::vector<std::string> words;
  size_t pos = 0;

 public:
  explicit Tokenizer(const std::string &str, const std::string &sep = " \t\r\n\f") {
    size_t p = 0;
    while ((p = str.find_first_not_of(sep, p)) != std::string::npos) {
      size_t e = str.find_first_of(sep, p);
      words.push_back(str.substr(p, e == std::string::npos ? std::string::npos : e - p));
      if (e == std::string::npos) break;
      p = e;
    }
  }
  bool has_next() const { return pos < words.size(); }
  size_t count() const { return words.size(); }
  std::string next() {
    if (!has_next()) throw TokenizerException("No more tokens", "");
    return words[pos++];
  }
  std::vector<std::string> as_vector() const { return words; }
};

// ------------------------------------------------------------------------------ small stock classes
class Command : protected Pointers {
 public:
  explicit Command(LAMMPS *l) : Pointers(l) {}
  virtual void command(int, char **) = 0;
};

class Irregular : protected Pointers {   // serial: every atom already lives on its owner
 public:
  explicit Irregular(LAMMPS *l) : Pointers(l) {}
  void migrate_atoms(int = 0, int = 0, int * = nullptr) {}
  int create_data(int, int *, int = 0) { return 0; }
  void exchange_data(char *, int, char *) {}
  void destroy_data() {}
};

class ArgInfo {
 public:
  enum ArgTypes { NONE = 0, X = 1 << 0, V = 1 << 1, F = 1 << 2, COMPUTE = 1 << 3, FIX = 1 << 4, VARIABLE = 1 << 5,
                  KEYWORD = 1 << 6, TYPE = 1 << 7, MOLECULE = 1 << 8, DNAME = 1 << 9, INAME = 1 << 10, DENSITY_NUMBER = 1 << 11,
                  DENSITY_MASS = 1 << 12, MASS = 1 << 13, TEMPERATURE = 1 << 14, BIN1D = 1 << 15, BIN2D = 1 << 16, BIN3D = 1 << 17,
                  BINSPHERE = 1 << 18, BINCYLINDER = 1 << 19, UNKNOWN = 1 << 30 };
  ArgInfo(const std::string &arg, int allowed = COMPUTE | FIX | VARIABLE) : type(NONE), dim(0), index1(-1), index2(-1) {
    if (arg.size() > 2 && arg[1] == '_') {
      if (arg[0] == 'c' && (allowed & COMPUTE)) type = COMPUTE;
      else if (arg[0] == 'f' && (allowed & FIX)) type = FIX;
      else if (arg[0] == 'v' && (allowed & VARIABLE)) type = VARIABLE;
      else if (arg[0] == 'd' && (allowed & DNAME)) type = DNAME;
      else if (arg[0] == 'i' && (allowed & INAME)) type = INAME;
      else { index1 = 0; name = arg; return; }
      size_t b = arg.find('[', 2);
      if (b == std::string::npos) { index1 = 0; name = arg.substr(2); }
      else {
        name = arg.substr(2, b - 2);
        size_t e = arg.find(']', b);
        if (e == std::string::npos) { type = UNKNOWN; return; }
        index1 = atoi(arg.substr(b + 1, e - b - 1).c_str());
        dim = 1;
      }
    } else { index1 = 0; name = arg; }
  }
  int get_type() const { return type; }
  int get_dim() const { return dim; }
  int get_index1() const { return index1; }
  int get_index2() const { return index2; }
  const char *get_name() const { return name.c_str(); }
  char *copy_name() { return utils::strdup(name); }

 private:
  std::string name;
  int type, dim, index1, index2;
};

class FixStoreAtom : public Fix {
 public:
  double *vstore = nullptr;
  FixStoreAtom(LAMMPS *l, int n, char **a) : Fix(l, n, a) {}
  int setmask() override { return 0; }
};

class LabelMap {
 public:
  std::vector<std::string> typelabel;
};

class Region : protected Pointers {
 public:
  explicit Region(LAMMPS *l) : Pointers(l) {}
  void prematch() {}
  int match(double, double, double) { return 1; }
};

// ----------------------------------------------------------------------------------- Dump
// [stock] dump.h / dump.cpp, serial subset: one writer, text or binary, optional sort by id.
class Dump : protected Pointers {
  friend class Output;

 public:
  char *id = nullptr, *style = nullptr, *filename = nullptr;
  int igroup = 0, groupbit = 1;
  int first_flag = 0, clearstep = 0;
  int comm_forward = 0, comm_reverse = 0;

  Dump(LAMMPS *l, int narg, char **arg) : Pointers(l) {
    MPI_Comm_rank(world, &me);
    MPI_Comm_size(world, &nprocs);
    id = utils::strdup(arg[0]);
    igroup = l->group->find(arg[1]);
    if (igroup == -1) error->all(FLERR, "Could not find dump group ID {}", arg[1]);
    groupbit = l->group->bitmask[igroup];
    style = utils::strdup(arg[2]);
    filename = utils::strdup(arg[4]);
    format_line_user = format_float_user = format_int_user = format_bigint_user = nullptr;
    if (strchr(filename, '*')) multifile = 1;
    size_t fl = strlen(filename);
    if (fl > 4 && strcmp(filename + fl - 4, ".bin") == 0) binary = 1;
  }
  ~Dump() override {
    delete[] id; delete[] style; delete[] filename;
    delete[] format; delete[] format_default; delete[] format_line_user; delete[] format_float_user;
    delete[] format_int_user; delete[] format_bigint_user;
    memory->destroy(buf); memory->destroy(sbuf); memory->destroy(ids);
    if (fp && !multifile) fclose(fp);
  }
  void init() { init_style(); }
  virtual void write();
  void modify_params(int narg, char **arg);
  virtual double memory_usage() { return 0.0; }
  virtual int pack_forward_comm(int, int *, double *, int, int *) { return 0; }
  virtual void unpack_forward_comm(int, int, double *) {}

 protected:
  int me = 0, nprocs = 1;
  int filewriter = 1, multiproc = 0, nclusterprocs = 1;
  int compressed = 0, binary = 0, multifile = 0;
  int header_flag = 1, flush_flag = 1, sort_flag = 0, append_flag = 0, buffer_allow = 0, buffer_flag = 0, padflag = 0,
      pbcflag = 0, singlefile_opened = 0, sortcol = 0, sortcolm1 = 0, sortorder = 0, time_flag = 0, unit_flag = 0,
      unit_count = 0, delay_flag = 0, write_header_flag = 1, has_id = 1;
  bigint delaystep = 0;
  int refreshflag = 0, irefresh = 0, skipflag = 0, skipindex = 0;
  char *refresh = nullptr, *skipvar = nullptr, *idrefresh = nullptr;
  char boundstr[9] = {0};
  char *format = nullptr, *format_default = nullptr, *format_line_user, *format_float_user, *format_int_user,
       *format_bigint_user;
  char **format_column_user = nullptr;
  enum { INT, DOUBLE, STRING, STRING2, BIGINT };
  std::map<std::string, int> key2col;
  std::vector<std::string> keyword_user;
  FILE *fp = nullptr;
  int size_one = 0, nme = 0, nsme = 0;
  double boxxlo = 0, boxxhi = 0, boxylo = 0, boxyhi = 0, boxzlo = 0, boxzhi = 0, boxxy = 0, boxxz = 0, boxyz = 0;
  bigint ntotal = 0;
  int maxbuf = 0, maxids = 0, maxsbuf = 0;
  double *buf = nullptr;
  tagint *ids = nullptr;
  char *sbuf = nullptr;

  virtual void init_style() = 0;
  virtual void openfile();
  virtual int modify_param(int, char **) { return 0; }
  virtual void write_header(bigint) = 0;
  virtual int count();
  virtual void pack(tagint *) = 0;
  virtual int convert_string(int, double *) { return 0; }
  virtual void write_data(int, double *) = 0;
  virtual void write_footer() {}
  double compute_time() { return update->atime + (update->ntimestep - update->atimestep) * update->dt; }
};

inline int Dump::count() {
  if (igroup == 0) return atom->nlocal;
  int m = 0;
  for (int i = 0; i < atom->nlocal; i++) if (atom->mask[i] & groupbit) m++;
  return m;
}

inline void Dump::openfile() {
  if (singlefile_opened) return;
  if (multifile == 0) singlefile_opened = 1;
  std::string name = multifile ? utils::star_subst(filename, update->ntimestep, padflag) : std::string(filename);
  fp = fopen(name.c_str(), binary ? (append_flag ? "ab" : "wb") : (append_flag ? "a" : "w"));
  if (fp == nullptr) error->one(FLERR, "Cannot open dump file {}", name);
}

// [stock] Dump::write(), one process: box bounds, count, header, pack, optional sort, format, write
inline void Dump::write() {
  boxxlo = domain->boxlo[0]; boxxhi = domain->boxhi[0];
  boxylo = domain->boxlo[1]; boxyhi = domain->boxhi[1];
  boxzlo = domain->boxlo[2]; boxzhi = domain->boxhi[2];
  nme = count();
  ntotal = nme;
  if (multifile) openfile();
  if (write_header_flag && header_flag) write_header(ntotal);
  if (nme > maxbuf) {
    maxbuf = nme;
    memory->destroy(buf);
    memory->create(buf, (maxbuf ? maxbuf : 1) * size_one, "dump:buf");
  }
  if (sort_flag && sortcol == 0 && nme > maxids) {
    maxids = nme;
    memory->destroy(ids);
    memory->create(ids, maxids ? maxids : 1, "dump:ids");
  }
  if (sort_flag && sortcol == 0) pack(ids);
  else pack(nullptr);
  if (sort_flag) {   // Dump::sort(), serial: stable order by id (sortcol 0) or by a column, ascending / descending
    std::vector<int> idx(nme);
    std::iota(idx.begin(), idx.end(), 0);
    if (sortcol == 0) {
      if (sortorder == 0) std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ids[a] < ids[b]; });
      else std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ids[a] > ids[b]; });
    } else {
      const int col = sortcolm1, so = size_one;
      if (sortorder == 0) std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return buf[a * so + col] < buf[b * so + col]; });
      else std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return buf[a * so + col] > buf[b * so + col]; });
    }
    std::vector<double> tmp((size_t)nme * size_one);
    for (int r = 0; r < nme; r++) memcpy(&tmp[(size_t)r * size_one], &buf[(size_t)idx[r] * size_one], size_one * sizeof(double));
    if (nme) memcpy(buf, tmp.data(), tmp.size() * sizeof(double));
  }
  if (buffer_flag && !binary) {
    nsme = convert_string(nme, buf);
    write_data(nsme, (double *)sbuf);
  } else write_data(nme, buf);
  if (flush_flag && fp) fflush(fp);
  if (multifile && fp) { fclose(fp); fp = nullptr; }
}

// [stock] Dump::modify_params: the keywords of the base class, the rest goes to the style
inline void Dump::modify_params(int narg, char **arg) {
  int iarg = 0;
  while (iarg < narg) {
    auto need = [&](int n) { if (iarg + n > narg) error->all(FLERR, "Illegal dump_modify command"); };
    auto dupfmt = [&](char *&dst, const char *src) { delete[] dst; dst = utils::strdup(src); };
    if (strcmp(arg[iarg], "append") == 0) { need(2); append_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp); iarg += 2; }
    else if (strcmp(arg[iarg], "buffer") == 0) { need(2); buffer_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp);
      if (buffer_flag && buffer_allow == 0) error->all(FLERR, "Dump_modify buffer yes not allowed for this style"); iarg += 2; }
    else if (strcmp(arg[iarg], "flush") == 0) { need(2); flush_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp); iarg += 2; }
    else if (strcmp(arg[iarg], "header") == 0) { need(2); header_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp); iarg += 2; }
    else if (strcmp(arg[iarg], "pad") == 0) { need(2); padflag = utils::inumeric(FLERR, arg[iarg + 1], false, lmp); iarg += 2; }
    else if (strcmp(arg[iarg], "time") == 0) { need(2); time_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp); iarg += 2; }
    else if (strcmp(arg[iarg], "units") == 0) { need(2); unit_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp); iarg += 2; }
    else if (strcmp(arg[iarg], "sort") == 0) {
      need(2);
      if (strcmp(arg[iarg + 1], "off") == 0) sort_flag = 0;
      else if (strcmp(arg[iarg + 1], "id") == 0) { sort_flag = 1; sortcol = 0; sortorder = 0; }
      else { sort_flag = 1; sortcol = utils::inumeric(FLERR, arg[iarg + 1], false, lmp); sortorder = 0;
        if (sortcol == 0) error->all(FLERR, "Illegal dump_modify command");
        if (sortcol < 0) { sortorder = 1; sortcol = -sortcol; }
        sortcolm1 = sortcol - 1; }
      iarg += 2;
    }
    else if (strcmp(arg[iarg], "format") == 0) {
      need(2);
      if (strcmp(arg[iarg + 1], "none") == 0) {
        delete[] format_line_user; delete[] format_int_user; delete[] format_bigint_user; delete[] format_float_user;
        format_line_user = format_int_user = format_bigint_user = format_float_user = nullptr;
        for (int i = 0; i < size_one; i++) { delete[] format_column_user[i]; format_column_user[i] = nullptr; }
        iarg += 2;
        continue;
      }
      need(3);
      if (strcmp(arg[iarg + 1], "line") == 0) { dupfmt(format_line_user, arg[iarg + 2]); iarg += 3; }
      else {   // int / float / column index: handled by the style (DumpCustom::modify_param)
        int n = modify_param(narg - iarg, &arg[iarg]);
        if (n == 0) error->all(FLERR, "Illegal dump_modify command");
        iarg += n;
      }
    }
    else {
      int n = modify_param(narg - iarg, &arg[iarg]);
      if (n == 0) error->all(FLERR, "Unknown dump_modify keyword: {}", arg[iarg]);
      iarg += n;
    }
  }
}

}  // namespace LAMMPS_NS
