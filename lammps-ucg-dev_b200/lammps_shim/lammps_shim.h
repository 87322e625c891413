// lammps_shim.h — a minimal, hand-written stand-in for the parts of the stock LAMMPS API
// that the UCG package compiles against (no LAMMPS source exists in this environment).
//
// Two users:
//   * oracle/_ref: the reference's own UCG/*.cpp are compiled VERBATIM against these headers
//     plus a serial driver (oracle/ref_driver.cpp) to pin the CPU oracle;
//   * lammps-ucg-dev_b200/host/: the GPU-backed style classes are compile-checked and
//     exercised against the same API, so that they drop into a real LAMMPS tree unchanged.
//
// Only declarations the UCG sources actually touch are present (SURVEY.md §8c lists them).
// Everything here is written from the public LAMMPS class interfaces; nothing is copied.
#pragma once
#define LAMMPS_UCG_SHIM 1   // lets the style classes fill shim-only diagnostics (Pair::virial_tally)

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <regex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

// ----------------------------------------------------------------------- MPI (serial stubs)
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_CHAR 3
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
static inline size_t shim_mpi_size(MPI_Datatype t) { return (t == MPI_DOUBLE || t == 4 /* MPI_LMP_BIGINT */) ? 8 : (t == MPI_INT ? sizeof(int) : 1); }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if (in != out) memcpy(out, in, n * shim_mpi_size(t));
  return 0;
}

namespace LAMMPS_NS {

typedef int tagint;
typedef int64_t bigint;
typedef int imageint;
typedef int smallint;
#define MAXTAGINT 0x7FFFFFFF
#define MAXSMALLINT 0x7FFFFFFF
#define MAXBIGINT 0x7FFFFFFFFFFFFFFFLL
#define FLERR __FILE__, __LINE__
#ifndef MAX
#define MAX(A, B) ((A) > (B) ? (A) : (B))
#define MIN(A, B) ((A) < (B) ? (A) : (B))
#endif
#define SBBITS 30
#define IMGMASK 1023
#define IMGMAX 512
#define IMGBITS 10
#define IMG2BITS 20
#define NEIGHMASK 0x1FFFFFFF
#define BIG_SHIM 1.0e20

union union_int_float_t {
  int i;
  float f;
};

class LAMMPS;
class Memory;
class Error;
class Atom;
class AtomVec;
class Force;
class Update;
class Modify;
class Neighbor;
class NeighList;
class NeighRequest;
class Comm;
class Domain;
class Group;
class Input;
class Variable;
class Pair;
class Fix;
class Compute;
class Integrate;
class Bond;
class Output;

class LAMMPSException : public std::runtime_error {
 public:
  explicit LAMMPSException(const std::string &m) : std::runtime_error(m) {}
};

// ------------------------------------------------------------------------------ formatting
namespace shimfmt {
inline void put(std::ostringstream &os) { (void)os; }
template <class T>
inline void one(std::ostringstream &os, const T &v) { os << v; }
inline void one(std::ostringstream &os, const char *v) { os << (v ? v : "(null)"); }
inline void one(std::ostringstream &os, char *v) { os << (v ? v : "(null)"); }
// "{:>1.16e}" / "{:.16}" on a double: the fmt spec [align][width][.precision][type] as its printf twin
// (no type letter = general format); anything else falls back to the stream form
template <class T>
inline void one_spec(std::ostringstream &os, const std::string &, const T &v) { one(os, v); }
inline void one_spec(std::ostringstream &os, const std::string &spec, const double &v) {
  std::string p = spec;
  if (!p.empty() && (p[0] == '>' || p[0] == '<')) p = p.substr(1);
  char type = 'g';
  if (!p.empty() && isalpha((unsigned char)p.back())) { type = p.back(); p.pop_back(); }
  char buf[128];
  snprintf(buf, sizeof buf, ("%" + p + type).c_str(), v);
  os << buf;
}
inline std::string format(const std::string &f) { return f; }
template <class T, class... R>
inline std::string format(const std::string &f, const T &v, const R &...rest) {
  size_t p = f.find("{}");
  // printf-style texts in the reference ("%d") carry no {}: extra arguments are dropped
  if (p == std::string::npos) {
    size_t q = f.find("{:");
    if (q == std::string::npos) return f;
    size_t e = f.find('}', q);
    std::ostringstream os;
    os << f.substr(0, q);
    one_spec(os, f.substr(q + 2, e - q - 2), v);
    return os.str() + format(f.substr(e + 1), rest...);
  }
  std::ostringstream os;
  os.precision(15);
  os << f.substr(0, p);
  one(os, v);
  return os.str() + format(f.substr(p + 2), rest...);
}
}  // namespace shimfmt

// ----------------------------------------------------------------------------------- Error
class Error {
 public:
  static constexpr int NOLASTLINE = -2;
  static constexpr int NOPOINTER = -1;
  std::vector<std::string> warnings;
  template <class... A>
  [[noreturn]] void all(const std::string &file, int line, const std::string &f, const A &...a) {
    fail("ERROR: ", file, line, shimfmt::format(f, a...));
  }
  template <class... A>
  [[noreturn]] void all(const std::string &file, int line, int, const std::string &f, const A &...a) {
    fail("ERROR: ", file, line, shimfmt::format(f, a...));
  }
  template <class... A>
  [[noreturn]] void one(const std::string &file, int line, const std::string &f, const A &...a) {
    fail("ERROR on proc 0: ", file, line, shimfmt::format(f, a...));
  }
  template <class... A>
  [[noreturn]] void one(const std::string &file, int line, int, const std::string &f, const A &...a) {
    fail("ERROR on proc 0: ", file, line, shimfmt::format(f, a...));
  }
  template <class... A>
  void warning(const std::string &, int, const std::string &f, const A &...a) {
    warnings.push_back(shimfmt::format(f, a...));
  }
  template <class... A>
  void message(const std::string &, int, const std::string &, const A &...) {}

 private:
  [[noreturn]] static void fail(const char *pre, const std::string &file, int line, const std::string &msg) {
    size_t s = file.find_last_of('/');
    throw LAMMPSException(std::string(pre) + msg + " (" + (s == std::string::npos ? file : file.substr(s + 1)) +
                          ":" + std::to_string(line) + ")");
  }
};

// ---------------------------------------------------------------------------------- Memory
class Memory {
 public:
  void *smalloc(bigint n, const char *) { return n > 0 ? malloc((size_t)n) : nullptr; }
  void *srealloc(void *p, bigint n, const char *) {
    if (n == 0) { free(p); return nullptr; }
    return realloc(p, (size_t)n);
  }
  void sfree(void *p) { free(p); }
  template <class T>
  double usage(T *, int n1, int n2 = 1, int n3 = 1) { return (double)sizeof(T) * n1 * n2 * n3; }

  template <class T>
  T *create(T *&a, int n, const char *name) {
    a = (T *)smalloc((bigint)sizeof(T) * n, name);
    return a;
  }
  template <class T>
  T *grow(T *&a, int n, const char *name) {
    if (a == nullptr) return create(a, n, name);
    a = (T *)srealloc(a, (bigint)sizeof(T) * n, name);
    return a;
  }
  template <class T>
  void destroy(T *&a) { sfree(a); a = nullptr; }

  template <class T>
  T **create(T **&a, int n1, int n2, const char *name) {
    T *data = (T *)smalloc((bigint)sizeof(T) * n1 * n2, name);
    a = (T **)smalloc((bigint)sizeof(T *) * n1, name);
    bigint n = 0;
    for (int i = 0; i < n1; i++) { a[i] = &data[n]; n += n2; }
    return a;
  }
  template <class T>
  T **grow(T **&a, int n1, int n2, const char *name) {
    if (a == nullptr) return create(a, n1, n2, name);
    T *data = (T *)srealloc(a[0], (bigint)sizeof(T) * n1 * n2, name);
    a = (T **)srealloc(a, (bigint)sizeof(T *) * n1, name);
    bigint n = 0;
    for (int i = 0; i < n1; i++) { a[i] = &data[n]; n += n2; }
    return a;
  }
  template <class T>
  void destroy(T **&a) {
    if (a == nullptr) return;
    sfree(a[0]);
    sfree(a);
    a = nullptr;
  }
  template <class T>
  T ***create(T ***&a, int n1, int n2, int n3, const char *name) {
    T *data = (T *)smalloc((bigint)sizeof(T) * n1 * n2 * n3, name);
    T **plane = (T **)smalloc((bigint)sizeof(T *) * n1 * n2, name);
    a = (T ***)smalloc((bigint)sizeof(T **) * n1, name);
    bigint n = 0;
    for (int i = 0; i < n1; i++) {
      a[i] = &plane[(bigint)i * n2];
      for (int j = 0; j < n2; j++) { plane[(bigint)i * n2 + j] = &data[n]; n += n3; }
    }
    return a;
  }
  template <class T>
  void destroy(T ***&a) {
    if (a == nullptr) return;
    sfree(a[0][0]);
    sfree(a[0]);
    sfree(a);
    a = nullptr;
  }
};

// -------------------------------------------------------------------------------- Pointers
class LAMMPS {
 public:
  Memory *memory = nullptr;
  Error *error = nullptr;
  Atom *atom = nullptr;
  Force *force = nullptr;
  Update *update = nullptr;
  Modify *modify = nullptr;
  Neighbor *neighbor = nullptr;
  Comm *comm = nullptr;
  Domain *domain = nullptr;
  Group *group = nullptr;
  Input *input = nullptr;
  Output *output = nullptr;
  MPI_Comm world = 0;
  FILE *infile = nullptr, *screen = nullptr, *logfile = nullptr;
  const char *suffix = nullptr, *suffix2 = nullptr;
  int suffix_enable = 0;
};

class Pointers {
 public:
  Pointers(LAMMPS *ptr)
      : lmp(ptr), memory(ptr->memory), error(ptr->error), atom(ptr->atom), force(ptr->force), update(ptr->update),
        modify(ptr->modify), neighbor(ptr->neighbor), comm(ptr->comm), domain(ptr->domain), group(ptr->group),
        input(ptr->input), output(ptr->output), world(ptr->world), infile(ptr->infile), screen(ptr->screen),
        logfile(ptr->logfile) {}
  virtual ~Pointers() = default;

 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Atom *&atom;
  Force *&force;
  Update *&update;
  Modify *&modify;
  Neighbor *&neighbor;
  Comm *&comm;
  Domain *&domain;
  Group *&group;
  Input *&input;
  Output *&output;
  MPI_Comm &world;
  FILE *&infile;
  FILE *&screen;
  FILE *&logfile;
};

// ----------------------------------------------------------------------------------- utils
namespace utils {
enum { NOCONVERT = 0, METAL2REAL = 1, REAL2METAL = 1 << 1 };
enum { UNKNOWN = 0, ENERGY };
inline int get_supported_conversions(int) { return 0; }
inline double get_conversion_factor(int, int) { return 1.0; }
inline bool strmatch(const std::string &text, const std::string &pattern) {
  return std::regex_search(text, std::regex(pattern));
}
inline std::string strip_style_suffix(const std::string &style, LAMMPS *) { return style; }
inline double numeric(const char *file, int line, const std::string &s, bool, LAMMPS *lmp) {
  char *end = nullptr;
  double v = strtod(s.c_str(), &end);
  if (s.empty() || end == s.c_str() || *end != '\0')
    lmp->error->all(file, line, "Expected floating point parameter instead of '{}' in input script or data file", s);
  return v;
}
inline int inumeric(const char *file, int line, const std::string &s, bool, LAMMPS *lmp) {
  char *end = nullptr;
  long v = strtol(s.c_str(), &end, 10);
  if (s.empty() || end == s.c_str() || *end != '\0')
    lmp->error->all(file, line, "Expected integer parameter instead of '{}' in input script or data file", s);
  return (int)v;
}
inline bigint bnumeric(const char *file, int line, const std::string &s, bool b, LAMMPS *lmp) { return inumeric(file, line, s, b, lmp); }
template <class T>
inline void bounds(const char *file, int line, const std::string &str, bigint nmin, bigint nmax, T &nlo, T &nhi, Error *error) {
  size_t star = str.find('*');
  nlo = nhi = -1;
  try {
    if (star == std::string::npos) { nlo = nhi = (T)std::stol(str); }
    else if (str.size() == 1) { nlo = (T)nmin; nhi = (T)nmax; }
    else if (star == 0) { nlo = (T)nmin; nhi = (T)std::stol(str.substr(1)); }
    else if (star == str.size() - 1) { nlo = (T)std::stol(str.substr(0, star)); nhi = (T)nmax; }
    else { nlo = (T)std::stol(str.substr(0, star)); nhi = (T)std::stol(str.substr(star + 1)); }
  } catch (...) { nlo = nhi = -1; }
  if (error) {
    if (nlo < nmin || nhi > nmax || nlo > nhi)
      error->all(file, line, "Numeric index {} is out of bounds ({}-{})", str, nmin, nmax);
  }
}
inline void missing_cmd_args(const std::string &file, int line, const std::string &cmd, Error *error) {
  if (error) error->all(file, line, "Illegal {} command: missing argument(s)", cmd);
}
inline std::string getsyserror() { return std::string(strerror(errno)); }
inline FILE *open_potential(const std::string &name, LAMMPS *, int *) { return fopen(name.c_str(), "r"); }
inline void sfread(const char *file, int line, void *s, size_t size, size_t num, FILE *fp, const char *, Error *error) {
  size_t rv = fread(s, size, num, fp);
  if (rv != num && error) error->one(file, line, "Unexpected end of file while reading");
}
inline int trim_and_count_words(const std::string &text, const std::string &sep = " \t\r\n\f") {
  std::string t = text.substr(0, text.find('#'));
  int n = 0;
  size_t p = 0;
  while ((p = t.find_first_not_of(sep, p)) != std::string::npos) { n++; p = t.find_first_of(sep, p); if (p == std::string::npos) break; }
  return n;
}
inline char *strdup(const std::string &s) {
  char *r = new char[s.size() + 1];
  strcpy(r, s.c_str());
  return r;
}
template <class... A>
inline void logmesg(LAMMPS *lmp, const std::string &f, const A &...a) {
  if (lmp && lmp->screen) fputs(shimfmt::format(f, a...).c_str(), lmp->screen);
}
inline std::string trim(const std::string &s) {
  size_t b = s.find_first_not_of(" \t\r\n");
  if (b == std::string::npos) return "";
  return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
}
}  // namespace utils

// ------------------------------------------------------------------------------- tokenizer
class TokenizerException : public std::exception {
  std::string message;

 public:
  TokenizerException(const std::string &msg, const std::string &token)
      : message(token.empty() ? msg : msg + ": '" + token + "'") {}
  const char *what() const noexcept override { return message.c_str(); }
};
class InvalidIntegerException : public TokenizerException {
 public:
  explicit InvalidIntegerException(const std::string &t) : TokenizerException("Not a valid integer number", t) {}
};
class InvalidFloatException : public TokenizerException {
 public:
  explicit InvalidFloatException(const std::string &t) : TokenizerException("Not a valid floating-point number", t) {}
};

class ValueTokenizer {
  std::vector<std::string> words;
  size_t pos = 0;

 public:
  explicit ValueTokenizer(const std::string &str, const std::string &sep = " \t\r\n\f") {
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
  std::string next_string() {
    if (!has_next()) throw TokenizerException("No more tokens", "");
    return words[pos++];
  }
  int next_int() {
    std::string w = next_string();
    char *e = nullptr;
    long v = strtol(w.c_str(), &e, 10);
    if (e == w.c_str() || *e) throw InvalidIntegerException(w);
    return (int)v;
  }
  bigint next_bigint() { return next_int(); }
  tagint next_tagint() { return next_int(); }
  double next_double() {
    std::string w = next_string();
    char *e = nullptr;
    double v = strtod(w.c_str(), &e);
    if (e == w.c_str() || *e) throw InvalidFloatException(w);
    return v;
  }
  void skip(int n = 1) { pos += n; }
};

// [stock] PotentialFileReader/TableFileReader: comment-stripping line reader; a section
// starts at the line whose first word is the keyword.
class TableFileReader {
  LAMMPS *lmp;
  FILE *fp;
  std::string name;
  char line[1024];

  char *raw_line() {
    if (!fgets(line, sizeof(line), fp)) return nullptr;
    char *c = strchr(line, '#');
    if (c) *c = '\0';
    return line;
  }

 public:
  TableFileReader(LAMMPS *l, const std::string &filename, const std::string &type, const int = 0) : lmp(l), name(filename) {
    fp = fopen(filename.c_str(), "r");
    if (!fp) lmp->error->one(FLERR, "Cannot open {} table file {}: {}", type, filename, utils::getsyserror());
  }
  ~TableFileReader() { if (fp) fclose(fp); }
  int get_unit_convert() const { return 0; }
  char *find_section_start(const std::string &keyword) {
    while (char *l = raw_line()) {
      ValueTokenizer v(l);
      if (v.has_next() && v.next_string() == keyword) return l;
    }
    return nullptr;
  }
  // next non-blank line
  char *next_line(int = 0) {
    while (char *l = raw_line()) {
      if (utils::trim(l).size()) return l;
    }
    return nullptr;
  }
  void skip_line() { /* next_line() already skips the blank separator */ }
};

// ------------------------------------------------------------------------------------ RNGs
// [stock] RanMars / RanPark, restated from the published algorithms (Marsaglia; Park-Miller)
class RanMars : protected Pointers {
  double u[98];
  int i97, j97;
  double c, cd, cm;

 public:
  RanMars(LAMMPS *l, int seed) : Pointers(l) {
    if (seed <= 0 || seed > 900000000) error->one(FLERR, "Invalid seed for Marsaglia random # generator");
    int ij = (seed - 1) / 30082, kl = (seed - 1) - 30082 * ij;
    int i = (ij / 177) % 177 + 2, j = ij % 177 + 2, k = (kl / 169) % 178 + 1, l2 = kl % 169;
    memset(u, 0, sizeof(u));
    for (int ii = 1; ii <= 97; ii++) {
      double s = 0.0, t = 0.5;
      for (int jj = 1; jj <= 24; jj++) {
        int m = ((i * j) % 179) * k % 179;
        i = j; j = k; k = m;
        l2 = (53 * l2 + 1) % 169;
        if ((l2 * m) % 64 >= 32) s = s + t;
        t = 0.5 * t;
      }
      u[ii] = s;
    }
    c = 362436.0 / 16777216.0; cd = 7654321.0 / 16777216.0; cm = 16777213.0 / 16777216.0;
    i97 = 97; j97 = 33;
    uniform();
  }
  double uniform() {
    double uni = u[i97] - u[j97];
    if (uni < 0.0) uni += 1.0;
    u[i97] = uni;
    if (--i97 == 0) i97 = 97;
    if (--j97 == 0) j97 = 97;
    c -= cd;
    if (c < 0.0) c += cm;
    uni -= c;
    if (uni < 0.0) uni += 1.0;
    return uni;
  }
};
class RanPark : protected Pointers {
  int seed;

 public:
  RanPark(LAMMPS *l, int s) : Pointers(l), seed(s) {
    if (s <= 0) error->one(FLERR, "Invalid seed for Park random # generator");
  }
  double uniform() {
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
    int k = seed / IQ;
    seed = IA * (seed - k * IQ) - IR * k;
    if (seed < 0) seed += IM;
    return (1.0 / IM) * seed;
  }
  void reset(int s) { seed = s; }
};

// ------------------------------------------------------------------------------------ Atom
class Atom : protected Pointers {
 public:
  enum { DOUBLE, INT, BIGINT };
  enum { GROW = 0, RESTART = 1, BORDER = 2 };
  enum { ATOMIC = 0, MOLECULAR = 1, TEMPLATE = 2 };
  enum { ATOM = 0, BOND = 1, ANGLE = 2, DIHEDRAL = 3, IMPROPER = 4 };
  explicit Atom(LAMMPS *l) : Pointers(l) {}

  bigint natoms = 0;
  int nlocal = 0, nghost = 0, nmax = 0;
  int ntypes = 0;
  int nfirst = 0, firstgroup = -1;
  int tag_enable = 1, molecular = 0;
  int molecule_flag = 0, q_flag = 0, rmass_flag = 0;
  AtomVec *avec = nullptr;

  tagint *tag = nullptr;
  int *type = nullptr, *mask = nullptr;
  imageint *image = nullptr;
  double **x = nullptr, **v = nullptr, **f = nullptr;
  tagint *molecule = nullptr;
  double *q = nullptr, *rmass = nullptr, *mass = nullptr;
  int *mass_setflag = nullptr;
  int *num_bond = nullptr, *num_angle = nullptr, *num_dihedral = nullptr, *num_improper = nullptr;
  int **bond_type = nullptr, **angle_type = nullptr, **dihedral_type = nullptr, **improper_type = nullptr;
  int **nspecial = nullptr;
  tagint **special = nullptr;

  // UCG members exactly as the reference's patched atom.h:180-194, 244-247
  int *ucgstate = nullptr;
  double *ucgl = nullptr, *ucgvl = nullptr, *ucgml = nullptr, *ucgp = nullptr;
  double *ucgforce = nullptr;
  double **ucgsoftmaxscores = nullptr;
  int max_ucgstates = 0;
  int *num_ucgstates = nullptr;
  int ucg_flag = 0, msucg_flag = 0;

  int map_style = 0;
  int nextsort = 0, sortfreq = 0;

  // members only the reference's patched I/O files touch (dump_custom.cpp, read_dump.cpp); never allocated here
  enum { MAP_NONE = 0, MAP_ARRAY = 1, MAP_HASH = 2, MAP_YES = 3 };
  int mu_flag = 0, radius_flag = 0, omega_flag = 0, angmom_flag = 0, torque_flag = 0, heatflow_flag = 0, temperature_flag = 0;
  double **mu = nullptr, **omega = nullptr, **angmom = nullptr, **torque = nullptr;
  double *radius = nullptr, *heatflow = nullptr, *temperature = nullptr;
  int **ivector = nullptr, ***iarray = nullptr, *icols = nullptr, *dcols = nullptr;
  double **dvector = nullptr, ***darray = nullptr;
  class LabelMap *lmap = nullptr;
  double **msucgl = nullptr, **msucgp = nullptr;
  void data_fix_compute_variable(int, int) {}
  int find_custom(const char *, int &, int &) { return -1; }
  // global id -> local index map, [stock] Atom::map_* (serial: a hash over the owned atoms)
  tagint map_tag_max = -1;
  std::map<tagint, int> idmap;
  void map_init(int = 1) { idmap.clear(); map_tag_max = -1; for (int i = 0; i < nlocal; i++) map_tag_max = MAX(map_tag_max, tag[i]); }
  void map_clear() { idmap.clear(); }
  void map_set() { for (int i = nlocal - 1; i >= 0; i--) idmap[tag[i]] = i; }
  void map_delete() { idmap.clear(); }
  int map(tagint id) { auto it = idmap.find(id); return it == idmap.end() ? -1 : it->second; }
  void tag_check() {}
  void tag_extend() {}
};

// --------------------------------------------------------------------------------- AtomVec
class AtomVec : protected Pointers {
 public:
  enum { PER_ATOM = 0, PER_TYPE = 1 };
  int molecular = 0, bonds_allow = 0, angles_allow = 0, dihedrals_allow = 0, impropers_allow = 0;
  int mass_type = 0, dipole_type = 0, forceclearflag = 0;
  int comm_x_only = 1, comm_f_only = 1;
  int size_forward = 3, size_reverse = 3, size_border = 6, size_velocity = 3;
  std::vector<std::string> fields_grow, fields_copy, fields_comm, fields_comm_vel, fields_reverse, fields_border,
      fields_border_vel, fields_exchange, fields_restart, fields_create, fields_data_atom, fields_data_vel;

  explicit AtomVec(LAMMPS *l) : Pointers(l) {}
  virtual void grow_pointers() {}
  virtual void force_clear(int, size_t) {}
  virtual void data_atom_post(int) {}
  // implemented by the serial driver (it owns the per-atom storage)
  void (*copy_hook)(void *, int, int) = nullptr;
  void *hook_arg = nullptr;
  void copy(int i, int j, int = 0) { if (copy_hook) copy_hook(hook_arg, i, j); }
  virtual void create_atom(int, double *) {}
  virtual int property_atom(const std::string &) { return -1; }
  virtual void pack_property_atom(int, double *, int, int) {}
  void setup_fields() {
    // sizes as [stock] AtomVec::setup_fields derives them from the field lists
    auto cols = [](const std::string &f) { return f == "ucgsoftmaxscores" ? 2 : (f == "x" || f == "v" ? 3 : 1); };
    size_forward = 3; size_reverse = 3; size_border = 6;
    for (auto &f : fields_comm) size_forward += cols(f);
    for (auto &f : fields_reverse) size_reverse += cols(f);
    for (auto &f : fields_border) size_border += cols(f);
    comm_x_only = fields_comm.empty();
    comm_f_only = fields_reverse.empty();
  }
};

// ----------------------------------------------------------------- Force / Update / misc
class Bond;
class Force : protected Pointers {
 public:
  explicit Force(LAMMPS *l) : Pointers(l) {}
  double boltz = 1.0, hplanck = 1.0, mvv2e = 1.0, ftm2v = 1.0, mv2d = 1.0, nktv2p = 1.0, qqr2e = 1.0, qe2f = 1.0,
         vxmu2f = 1.0, xxt2kmu = 1.0, dielectric = 1.0, qqrd2e = 1.0, e_mass = 0.0, hhmrr2e = 0.0, mvh2r = 0.0,
         angstrom = 1.0, femtosecond = 1.0, qelectron = 1.0;
  int newton = 1, newton_pair = 1, newton_bond = 1;
  Pair *pair = nullptr;
  char *pair_style = nullptr;
  Bond *bond = nullptr;
  void *angle = nullptr, *dihedral = nullptr, *improper = nullptr, *kspace = nullptr;
  double special_lj[4] = {1.0, 0.0, 0.0, 0.0};
  double special_coul[4] = {1.0, 0.0, 0.0, 0.0};
  double numeric(const char *file, int line, char *str) { return utils::numeric(file, line, str, false, lmp); }
  int inumeric(const char *file, int line, char *str) { return utils::inumeric(file, line, str, false, lmp); }
};

class Integrate : protected Pointers {
 public:
  explicit Integrate(LAMMPS *l) : Pointers(l) {}
  Integrate(LAMMPS *l, int, char **) : Pointers(l) {}
  virtual ~Integrate() = default;
  virtual void init() {}
  virtual void setup(int) {}
  virtual void setup_minimal(int) {}
  virtual void run(int) {}
  virtual void force_clear() {}
  virtual void cleanup() {}
  virtual void reset_dt() {}
};
class Respa : public Integrate {
 public:
  explicit Respa(LAMMPS *l) : Integrate(l) {}
  int nlevels = 1;
  int level_inner = -1, level_middle = -1, level_outer = -1;
  double *step = nullptr;
  void copy_flevel_f(int) {}
  void copy_f_flevel(int) {}
};

class Update : protected Pointers {
 public:
  explicit Update(LAMMPS *l) : Pointers(l) {}
  double dt = 0.005;
  double atime = 0.0;
  bigint ntimestep = 0, firststep = 0, laststep = 0, beginstep = 0, endstep = 0;
  int nsteps = 0, whichflag = 0;
  int restrict_output = 0, setupflag = 0, multireplica = 0;
  bigint atimestep = 0;
  char *integrate_style = (char *)"verlet";
  Integrate *integrate = nullptr;
  char *unit_style = (char *)"lj";
  void reset_timestep(bigint n, bool = true) { ntimestep = n; }
};

class Compute : protected Pointers {
 public:
  explicit Compute(LAMMPS *l) : Pointers(l) {}
  char *id = nullptr, *style = nullptr;
  int igroup = 0, groupbit = 1, tempflag = 0, tempbias = 0;
  double scalar = 0.0;
  enum { INVOKED_NONE = 0, INVOKED_SCALAR = 1 << 0, INVOKED_VECTOR = 1 << 1, INVOKED_ARRAY = 1 << 2, INVOKED_PERATOM = 1 << 3 };
  int peratom_flag = 0, size_peratom_cols = 0, invoked_flag = 0, initialized_flag = 1;
  double *vector_atom = nullptr, **array_atom = nullptr;
  bool is_initialized() const { return initialized_flag == 1; }
  virtual void compute_peratom() {}
  virtual double compute_scalar() { return scalar; }
  virtual void remove_bias(int, double *) {}
  virtual void restore_bias(int, double *) {}
};

class Variable : protected Pointers {
 public:
  explicit Variable(LAMMPS *l) : Pointers(l) {}
  int find(const char *) { return -1; }
  int equalstyle(int) { return 0; }
  int atomstyle(int) { return 0; }
  double compute_equal(int) { return 0.0; }
  void compute_atom(int, int, double *, int, int) {}
};
class Input : protected Pointers {
 public:
  explicit Input(LAMMPS *l) : Pointers(l), variable(new Variable(l)) {}
  ~Input() override { delete variable; }
  Variable *variable;
  char **arg = nullptr;
  int narg = 0;
};
class Output : protected Pointers {
 public:
  explicit Output(LAMMPS *l) : Pointers(l) {}
  bigint next = MAXBIGINT;   // next step any output (thermo, dump, restart) is due
  int thermo_every = 0;
  // implemented by the serial driver
  void (*write_hook)(void *, bigint) = nullptr;
  void *hook_arg = nullptr;
  void setup(int = 1) {}
  void write(bigint step) { if (write_hook) write_hook(hook_arg, step); }
};

class Group : protected Pointers {
 public:
  explicit Group(LAMMPS *l) : Pointers(l) {
    names = new char *[32];
    bitmask = new int[32];
    for (int i = 0; i < 32; i++) { names[i] = nullptr; bitmask[i] = 1 << i; }
    names[0] = utils::strdup("all");
    ngroup = 1;
  }
  ~Group() override {
    for (int i = 0; i < 32; i++) delete[] names[i];
    delete[] names;
    delete[] bitmask;
  }
  int ngroup;
  char **names;
  int *bitmask;
  int find(const std::string &name) {
    for (int i = 0; i < 32; i++) if (names[i] && name == names[i]) return i;
    return -1;
  }
  bigint count(int igroup);
};

class Domain : protected Pointers {
 public:
  explicit Domain(LAMMPS *l) : Pointers(l) {}
  int dimension = 3, triclinic = 0;
  int xperiodic = 1, yperiodic = 1, zperiodic = 1;
  int periodicity[3] = {1, 1, 1};
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {1, 1, 1}, prd[3] = {1, 1, 1};
  double sublo[3] = {0, 0, 0}, subhi[3] = {1, 1, 1};
  double xprd = 1, yprd = 1, zprd = 1;
  // implemented by the serial driver
  void (*pbc_hook)(void *) = nullptr;
  void *hook_arg = nullptr;
  void pbc() { if (pbc_hook) pbc_hook(hook_arg); }

  // what the reference's patched I/O files read: orthogonal boxes only, the triclinic members stay zero
  int triclinic_general = 0, box_exist = 1;
  void print_box(const std::string &) {}
  int boundary[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  double xy = 0, xz = 0, yz = 0;
  double h[6] = {1, 1, 1, 0, 0, 0}, h_inv[6] = {1, 1, 1, 0, 0, 0};
  double boxlo_lamda[3] = {0, 0, 0}, boxhi_lamda[3] = {1, 1, 1}, boxlo_bound[3] = {0, 0, 0}, boxhi_bound[3] = {1, 1, 1};
  double avec[3] = {1, 0, 0}, bvec[3] = {0, 1, 0}, cvec[3] = {0, 0, 1};
  void restricted_to_general_vector(double *) {}
  void restricted_to_general_vector(double *, double *) {}
  void restricted_to_general_coords(double *) {}
  void restricted_to_general_coords(double *, double *) {}
  class Region *get_region_by_id(const std::string &) { return nullptr; }
  void boundary_string(char *str) {   // [stock] Domain::boundary_string: "pp pp pp" from boundary[][]
    int m = 0;
    for (int d = 0; d < 3; d++) {
      for (int s = 0; s < 2; s++) str[m++] = "pfsm"[boundary[d][s]];
      str[m++] = ' ';
    }
    str[8] = '\0';
  }
  // [stock] Domain::set_initial_box / set_global_box / set_local_box for an orthogonal box on one process
  void set_initial_box(int = 1) {}
  void set_global_box() {
    for (int d = 0; d < 3; d++) prd[d] = boxhi[d] - boxlo[d];
    xprd = prd[0]; yprd = prd[1]; zprd = prd[2];
    h[0] = xprd; h[1] = yprd; h[2] = zprd;
    h_inv[0] = 1.0 / h[0]; h_inv[1] = 1.0 / h[1]; h_inv[2] = 1.0 / h[2];
  }
  void set_local_box() { for (int d = 0; d < 3; d++) { sublo[d] = boxlo[d]; subhi[d] = boxhi[d]; } }
  void reset_box() {}
  void x2lamda(int) {}
  void lamda2x(int) {}
  // [stock] Domain::remap(x, image): wrap into the periodic box, counting the crossings in the image flags
  void remap(double *x, imageint &image) {
    const int per[3] = {xperiodic, yperiodic, zperiodic};
    const int shift[3] = {0, IMGBITS, IMG2BITS};
    for (int d = 0; d < 3; d++) {
      if (!per[d]) continue;
      imageint idim, otherdims;
      while (x[d] < boxlo[d]) {
        x[d] += prd[d];
        idim = (image >> shift[d]) & IMGMASK;
        otherdims = image ^ (idim << shift[d]);
        idim--; idim &= IMGMASK;
        image = otherdims | (idim << shift[d]);
      }
      while (x[d] >= boxhi[d]) {
        x[d] -= prd[d];
        idim = (image >> shift[d]) & IMGMASK;
        otherdims = image ^ (idim << shift[d]);
        idim++; idim &= IMGMASK;
        image = otherdims | (idim << shift[d]);
      }
      x[d] = MAX(x[d], boxlo[d]);
    }
  }
};

// -------------------------------------------------------------------------------- Neighbor
namespace NeighConst {
enum {
  REQ_DEFAULT = 0, REQ_FULL = 1 << 0, REQ_GHOST = 1 << 1, REQ_SIZE = 1 << 2, REQ_HISTORY = 1 << 3, REQ_OCCASIONAL = 1 << 4,
  REQ_RESPA_INOUT = 1 << 5, REQ_RESPA_ALL = 1 << 6, REQ_NEWTON_ON = 1 << 8, REQ_NEWTON_OFF = 1 << 9, REQ_SSA = 1 << 10
};
}
class NeighRequest {
 public:
  void *requestor = nullptr;
  int pair = 0, fix = 0, half = 1, full = 0, occasional = 0, id = 0;
  double cutoff = 0.0;
  int cut = 0;
  NeighList *list = nullptr;
  void set_id(int i) { id = i; }
  void set_cutoff(double c) { cut = 1; cutoff = c; }
  NeighRequest *apply_flags(int flags) {
    if (flags & NeighConst::REQ_FULL) { half = 0; full = 1; }
    if (flags & NeighConst::REQ_OCCASIONAL) occasional = 1;
    return this;
  }
};
class NeighList {
 public:
  int inum = 0, gnum = 0;
  int *ilist = nullptr, *numneigh = nullptr;
  int **firstneigh = nullptr;
  int full = 0;
  std::vector<int> store_ilist, store_num, store_neigh;
  std::vector<int *> store_first;
};
class Neighbor : protected Pointers {
 public:
  explicit Neighbor(LAMMPS *l) : Pointers(l) {}
  ~Neighbor() override { for (auto r : requests) { delete r->list; delete r; } }
  double skin = 0.3, cutneighmax = 0.0;
  int every = 1, delay = 0, dist_check = 1, ago = 0;
  bigint ncalls = 0, ndanger = 0, lastcall = 0;
  std::vector<NeighRequest *> requests;
  NeighRequest *add_request(Pair *p, int flags = 0);
  NeighRequest *add_request(Fix *f, int flags = 0);
  // implemented by the serial driver
  void (*build_hook)(void *, int) = nullptr;
  void (*build_one_hook)(void *, NeighList *) = nullptr;
  void *hook_arg = nullptr;
  void build(int topoflag = 1) { if (build_hook) build_hook(hook_arg, topoflag); }
  void build_one(NeighList *list, int = 0) { if (build_one_hook) build_one_hook(hook_arg, list); }
};

// ------------------------------------------------------------------------------------ Comm
class Comm : protected Pointers {
 public:
  explicit Comm(LAMMPS *l) : Pointers(l) {}
  int me = 0, nprocs = 1;
  double cutghostuser = 0.0;
  int ghost_velocity = 0;
  // implemented by the serial driver
  struct Hooks {
    void (*forward)(void *) = nullptr;
    void (*reverse)(void *) = nullptr;
    void (*forward_pair)(void *, Pair *) = nullptr;
    void (*reverse_pair)(void *, Pair *) = nullptr;
    void (*forward_fix)(void *, Fix *) = nullptr;
    void (*exchange)(void *) = nullptr;
    void (*borders)(void *) = nullptr;
    void *arg = nullptr;
  } hooks;
  void forward_comm(int = 0) { if (hooks.forward) hooks.forward(hooks.arg); }
  void reverse_comm() { if (hooks.reverse) hooks.reverse(hooks.arg); }
  void forward_comm(Pair *p, int = 0) { if (hooks.forward_pair) hooks.forward_pair(hooks.arg, p); }
  void reverse_comm(Pair *p, int = 0) { if (hooks.reverse_pair) hooks.reverse_pair(hooks.arg, p); }
  void forward_comm(Fix *f, int = 0) { if (hooks.forward_fix) hooks.forward_fix(hooks.arg, f); }
  void exchange() { if (hooks.exchange) hooks.exchange(hooks.arg); }
  void borders() { if (hooks.borders) hooks.borders(hooks.arg); }
  void setup() {}
  void set_proc_grid(int = 1) {}
};

// ------------------------------------------------------------------------------------- Fix
namespace FixConst {
enum {
  INITIAL_INTEGRATE = 1 << 0, POST_INTEGRATE = 1 << 1, PRE_EXCHANGE = 1 << 2, PRE_NEIGHBOR = 1 << 3, POST_NEIGHBOR = 1 << 4,
  PRE_FORCE = 1 << 5, PRE_REVERSE = 1 << 6, POST_FORCE = 1 << 7, FINAL_INTEGRATE = 1 << 8, END_OF_STEP = 1 << 9,
  POST_RUN = 1 << 10, INITIAL_INTEGRATE_RESPA = 1 << 11, POST_INTEGRATE_RESPA = 1 << 12, PRE_FORCE_RESPA = 1 << 13,
  POST_FORCE_RESPA = 1 << 14, FINAL_INTEGRATE_RESPA = 1 << 15, MIN_PRE_EXCHANGE = 1 << 16, MIN_PRE_NEIGHBOR = 1 << 17,
  MIN_POST_NEIGHBOR = 1 << 18, MIN_PRE_FORCE = 1 << 19, MIN_PRE_REVERSE = 1 << 20, MIN_POST_FORCE = 1 << 21,
  MIN_ENERGY = 1 << 22
};
}
class Fix : protected Pointers {
 public:
  char *id = nullptr, *style = nullptr;
  int igroup = 0, groupbit = 1;
  int restart_global = 0, restart_peratom = 0, restart_file = 0, force_reneighbor = 0;
  bigint next_reneighbor = 0;
  int box_change = 0, nevery = 1, thermo_energy = 0, thermo_virial = 0, energy_global_flag = 0, energy_peratom_flag = 0,
      virial_global_flag = 0, virial_peratom_flag = 0, ecouple_flag = 0, time_integrate = 0, rigid_flag = 0,
      no_change_box = 0, time_depend = 0, create_attribute = 0, restart_pbc = 0, wd_header = 0, wd_section = 0,
      dynamic_group_allow = 0, dynamic = 0, dof_flag = 0, special_alter_flag = 0, enforce2d_flag = 0, respa_level_support = 0,
      respa_level = 0, maxexchange = 0, maxexchange_dynamic = 0, pre_exchange_migrate = 0, stores_ids = 0;
  int scalar_flag = 0, vector_flag = 0, array_flag = 0, size_vector = 0, size_array_rows = 0, size_array_cols = 0,
      size_vector_variable = 0, size_array_rows_variable = 0, global_freq = 0, peratom_flag = 0, size_peratom_cols = 0,
      peratom_freq = 0, local_flag = 0, size_local_rows = 0, size_local_cols = 0, local_freq = 0, extscalar = 0,
      extvector = 0, *extlist = nullptr, extarray = 0;
  double *vector_atom = nullptr, **array_atom = nullptr, *vector_local = nullptr, **array_local = nullptr;
  int comm_forward = 0, comm_reverse = 0, comm_border = 0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  double *eatom = nullptr, **vatom = nullptr;
  int copymode = 0, kokkosable = 0;

  Fix(LAMMPS *l, int narg, char **arg) : Pointers(l) {
    if (narg < 3) error->all(FLERR, "Illegal fix command");
    id = utils::strdup(arg[0]);
    igroup = l->group->find(arg[1]);
    if (igroup == -1) error->all(FLERR, "Could not find fix group ID {}", arg[1]);
    groupbit = l->group->bitmask[igroup];
    style = utils::strdup(arg[2]);
  }
  ~Fix() override { delete[] id; delete[] style; }
  virtual int setmask() = 0;
  virtual void post_constructor() {}
  virtual void init() {}
  virtual void init_list(int, NeighList *) {}
  virtual void setup(int) {}
  virtual void setup_pre_exchange() {}
  virtual void setup_pre_neighbor() {}
  virtual void setup_post_neighbor() {}
  virtual void setup_pre_force(int) {}
  virtual void setup_pre_reverse(int, int) {}
  virtual void min_setup(int) {}
  virtual void initial_integrate(int) {}
  virtual void post_integrate() {}
  virtual void pre_exchange() {}
  virtual void pre_neighbor() {}
  virtual void post_neighbor() {}
  virtual void pre_force(int) {}
  virtual void pre_reverse(int, int) {}
  virtual void post_force(int) {}
  virtual void final_integrate() {}
  virtual void end_of_step() {}
  virtual void post_run() {}
  virtual void write_restart(FILE *) {}
  virtual void restart(char *) {}
  virtual void grow_arrays(int) {}
  virtual void copy_arrays(int, int, int) {}
  virtual void set_arrays(int) {}
  virtual int pack_exchange(int, double *) { return 0; }
  virtual int unpack_exchange(int, double *) { return 0; }
  virtual void initial_integrate_respa(int, int, int) {}
  virtual void post_integrate_respa(int, int) {}
  virtual void pre_force_respa(int, int, int) {}
  virtual void post_force_respa(int, int, int) {}
  virtual void final_integrate_respa(int, int) {}
  virtual void min_pre_force(int) {}
  virtual void min_post_force(int) {}
  virtual int pack_forward_comm(int, int *, double *, int, int *) { return 0; }
  virtual void unpack_forward_comm(int, int, double *) {}
  virtual int pack_reverse_comm(int, int, double *) { return 0; }
  virtual void unpack_reverse_comm(int, int *, double *) {}
  virtual double compute_scalar() { return 0.0; }
  virtual double compute_vector(int) { return 0.0; }
  virtual double compute_array(int, int) { return 0.0; }
  virtual void reset_target(double) {}
  virtual void reset_dt() {}
  virtual int modify_param(int, char **) { return 0; }
  virtual void *extract(const char *, int &) { return nullptr; }
  virtual double memory_usage() { return 0.0; }
};

class Modify : protected Pointers {
 public:
  explicit Modify(LAMMPS *l) : Pointers(l) {}
  int nfix = 0;
  Fix **fix = nullptr;
  int *fmask = nullptr;
  int n_pre_neighbor = 0;
  std::vector<Fix *> fixes;
  std::vector<Compute *> computes;
  void add(Fix *f) {
    fixes.push_back(f);
    masks.push_back(f->setmask());
    nfix = (int)fixes.size();
    fix = fixes.data();
    fmask = masks.data();
  }
  Compute *get_compute_by_id(const std::string &id) {
    for (auto c : computes) if (c->id && id == c->id) return c;
    return nullptr;
  }
  Fix *get_fix_by_id(const std::string &id) {
    for (auto f : fixes) if (f->id && id == f->id) return f;
    return nullptr;
  }
  Fix *add_fix(const std::string &, int = 1) { return nullptr; }
  void delete_fix(const std::string &) {}
  void clearstep_compute() {}
  void addstep_compute(bigint) {}
  void pre_neighbor() { for (int i = 0; i < nfix; i++) if (fmask[i] & FixConst::PRE_NEIGHBOR) fix[i]->pre_neighbor(); }

 private:
  std::vector<int> masks;
};

// ------------------------------------------------------------------------------------ Pair
class Pair : protected Pointers {
  friend class Neighbor;

 public:
  static int instance_total;
  double eng_vdwl = 0.0, eng_coul = 0.0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  double *eatom = nullptr, **vatom = nullptr, **cvatom = nullptr;
  double cutforce = 0.0;
  double **cutsq = nullptr;
  int **setflag = nullptr;
  int comm_forward = 0, comm_reverse = 0, comm_reverse_off = 0;
  int single_enable = 1, born_matrix_enable = 0, single_hessian_enable = 0, restartinfo = 1, respa_enable = 0,
      one_coeff = 0, manybody_flag = 0, unit_convert_flag = 0, no_virial_fdotr = 0, finitecutflag = 0, ghostneigh = 0,
      ewaldflag = 0, pppmflag = 0, msmflag = 0, dispersionflag = 0, tip4pflag = 0, dipoleflag = 0, spinflag = 0,
      reinitflag = 1, centroidstressflag = 0;
  int tail_flag = 0;
  double etail = 0, ptail = 0, etail_ij = 0, ptail_ij = 0;
  int trim_flag = 1;
  int evflag = 0, eflag_either = 0, eflag_global = 0, eflag_atom = 0, vflag_either = 0, vflag_global = 0, vflag_atom = 0,
      cvflag_atom = 0, vflag_fdotr = 0;
  int ncoultablebits = 12, ndisptablebits = 12;
  NeighList *list = nullptr, *listhalf = nullptr, *listfull = nullptr;
  int allocated = 0;
  int copymode = 0, kokkosable = 0;
  int mix_flag = 0;
  // what ev_tally would have added with vflag_global set (the reference's virial is left at
  // zero as shipped); kept separately so that both numbers can be inspected
  double virial_tally[6] = {0, 0, 0, 0, 0, 0};

  explicit Pair(LAMMPS *l) : Pointers(l) {}
  ~Pair() override = default;
  virtual void compute(int, int) = 0;
  virtual void settings(int, char **) = 0;
  virtual void coeff(int, char **) = 0;
  virtual void init_style() {}
  virtual void init_list(int, NeighList *ptr) { list = ptr; }
  virtual double init_one(int, int) { return 0.0; }
  virtual void write_restart(FILE *) {}
  virtual void read_restart(FILE *) {}
  virtual void write_restart_settings(FILE *) {}
  virtual void read_restart_settings(FILE *) {}
  virtual double single(int, int, int, int, double, double, double, double &fforce) { fforce = 0.0; return 0.0; }
  virtual void *extract(const char *, int &) { return nullptr; }
  virtual int pack_forward_comm(int, int *, double *, int, int *) { return 0; }
  virtual void unpack_forward_comm(int, int, double *) {}
  virtual int pack_reverse_comm(int, int, double *) { return 0; }
  virtual void unpack_reverse_comm(int, int *, double *) {}
  virtual double memory_usage() { return 0.0; }

  // [stock] Pair::init: init_style + init_one over i<=j with symmetric cutsq
  void init();
  // [stock] Pair::ev_setup flag logic (global accumulators only)
  void ev_init(int eflag, int vflag, int = 1) {
    if (eflag || vflag) ev_setup(eflag, vflag);
    else evflag = eflag_either = eflag_global = eflag_atom = vflag_either = vflag_global = vflag_atom = cvflag_atom = vflag_fdotr = 0;
  }
  void ev_setup(int eflag, int vflag, int = 1) {
    evflag = 1;
    // [stock] ENERGY_GLOBAL = 1, ENERGY_ATOM = 2; VIRIAL_PAIR = 1, VIRIAL_FDOTR = 2, VIRIAL_ATOM = 4
    eflag_either = eflag; eflag_global = eflag & 1; eflag_atom = eflag & 2;
    vflag_either = vflag; vflag_global = vflag & 3; vflag_atom = vflag & 4; cvflag_atom = 0;
    if (eflag_atom || vflag_atom) shim_peratom_alloc();
    // with newton_pair on and no_virial_fdotr unset, the global virial is left to
    // virial_fdotr_compute() and ev_tally skips it
    vflag_fdotr = 0;
    if (vflag_global && no_virial_fdotr == 0 && shim_newton_pair()) { vflag_fdotr = 1; vflag_global = 0; }
    if (eflag_global) eng_vdwl = eng_coul = 0.0;
    for (int i = 0; i < 6; i++) { virial[i] = 0.0; virial_tally[i] = 0.0; }
  }
  void ev_unset() { evflag = 0; }
  void ev_tally(int i, int j, int nlocal, int newton_pair, double evdwl, double ecoul, double fpair, double delx,
                double dely, double delz) {
    if (eflag_either && eflag_global) {
      if (newton_pair) { eng_vdwl += evdwl; eng_coul += ecoul; }
      else {
        if (i < nlocal) { eng_vdwl += 0.5 * evdwl; eng_coul += 0.5 * ecoul; }
        if (j < nlocal) { eng_vdwl += 0.5 * evdwl; eng_coul += 0.5 * ecoul; }
      }
    }
    if (eflag_either && eflag_atom) {
      const double epairhalf = 0.5 * (evdwl + ecoul);
      if (newton_pair || i < nlocal) eatom[i] += epairhalf;
      if (newton_pair || j < nlocal) eatom[j] += epairhalf;
    }
    if (vflag_either || vflag_fdotr) {
      double v[6] = {delx * delx * fpair, dely * dely * fpair, delz * delz * fpair,
                     delx * dely * fpair, delx * delz * fpair, dely * delz * fpair};
      double w = newton_pair ? 1.0 : 0.5 * ((i < nlocal) + (j < nlocal));
      for (int k = 0; k < 6; k++) {
        virial_tally[k] += w * v[k];
        if (vflag_global) virial[k] += w * v[k];
      }
      if (vflag_atom) {
        if (newton_pair || i < nlocal) for (int k = 0; k < 6; k++) vatom[i][k] += 0.5 * v[k];
        if (newton_pair || j < nlocal) for (int k = 0; k < 6; k++) vatom[j][k] += 0.5 * v[k];
      }
    }
  }
  std::vector<double> shim_eatom, shim_vatom;
  std::vector<double *> shim_vrows;
  void shim_peratom_alloc();   // eatom / vatom over nlocal + nghost, zeroed ([stock] ev_setup)
  void virial_fdotr_compute();
  void init_bitmap(double inner, double outer, int ntablebits, int &masklo, int &maskhi, int &nmask, int &nshiftbits);
  inline int sbmask(int j) const { return j >> SBBITS & 3; }

 protected:
  int instance_me = 0;
  int special_lj_flag = 0;
  int maxeatom = 0, maxvatom = 0;
  int shim_newton_pair();
};

class Bond : protected Pointers {
 public:
  explicit Bond(LAMMPS *l) : Pointers(l) {}
};

class Info {
 public:
  static std::string get_pair_coeff_status(const LAMMPS *) { return "(pair coeff status unavailable in the shim)\n"; }
};

namespace MathConst {
static constexpr double MY_PI = 3.14159265358979323846;
static constexpr double MY_2PI = 6.28318530717958647692;
static constexpr double MY_PI2 = 1.57079632679489661923;
}  // namespace MathConst

class AtomVecEllipsoid;

// ------------------------------------------------------------ out-of-line members (inline)
inline NeighRequest *Neighbor::add_request(Pair *p, int flags) {
  auto *r = new NeighRequest();
  r->requestor = p; r->pair = 1;
  r->apply_flags(flags);
  r->list = new NeighList();
  r->list->full = r->full;
  requests.push_back(r);
  return r;
}
inline NeighRequest *Neighbor::add_request(Fix *f, int flags) {
  auto *r = new NeighRequest();
  r->requestor = f; r->fix = 1;
  r->apply_flags(flags);
  r->list = new NeighList();
  r->list->full = r->full;
  requests.push_back(r);
  return r;
}
inline int Pair::shim_newton_pair() { return force->newton_pair; }
inline void Pair::shim_peratom_alloc() {
  const size_t n = (size_t)atom->nlocal + atom->nghost + 1;
  shim_eatom.assign(n, 0.0);
  shim_vatom.assign(6 * n, 0.0);
  shim_vrows.resize(n);
  for (size_t i = 0; i < n; i++) shim_vrows[i] = shim_vatom.data() + 6 * i;
  eatom = shim_eatom.data();
  vatom = shim_vrows.data();
}
inline bigint Group::count(int igroup) {
  bigint n = 0;
  for (int i = 0; i < atom->nlocal; i++) if (atom->mask[i] & bitmask[igroup]) n++;
  return n;
}
inline void Pair::init() {
  if (!allocated) error->all(FLERR, "All pair coeffs are not set");
  init_style();
  cutforce = 0.0;
  for (int i = 1; i <= atom->ntypes; i++)
    for (int j = i; j <= atom->ntypes; j++) {
      double cut = init_one(i, j);
      cutsq[i][j] = cutsq[j][i] = cut * cut;
      cutforce = MAX(cutforce, cut);
    }
}
inline void Pair::virial_fdotr_compute() {
  double **x = atom->x, **f = atom->f;
  int nall = atom->nlocal + atom->nghost;
  if (!force->newton_pair) nall = atom->nlocal + atom->nghost;
  for (int i = 0; i < nall; i++) {
    virial[0] += f[i][0] * x[i][0];
    virial[1] += f[i][1] * x[i][1];
    virial[2] += f[i][2] * x[i][2];
    virial[3] += f[i][1] * x[i][0];
    virial[4] += f[i][2] * x[i][0];
    virial[5] += f[i][2] * x[i][1];
  }
  vflag_fdotr = 0;
}
// [stock] Pair::init_bitmap
inline void Pair::init_bitmap(double inner, double outer, int ntablebits, int &masklo, int &maskhi, int &nmask,
                              int &nshiftbits) {
  if (sizeof(int) != sizeof(float)) error->all(FLERR, "Bitmapped lookup tables require int/float be same size");
  if (ntablebits > (int)sizeof(float) * 8) error->all(FLERR, "Too many total bits for bitmapped lookup table");
  if (inner >= outer) error->all(FLERR, "Table inner cutoff >= outer cutoff");
  int nlowermin = 1;
  while (!((pow(double(2), (double)nlowermin) <= inner * inner) && (pow(double(2), (double)nlowermin + 1.0) > inner * inner))) {
    if (pow(double(2), (double)nlowermin) <= inner * inner) nlowermin++; else nlowermin--;
  }
  int nexpbits = 0;
  double required_range = outer * outer / pow(double(2), (double)nlowermin);
  double available_range = 2.0;
  while (available_range < required_range) {
    nexpbits++;
    available_range = pow(double(2), pow(double(2), (double)nexpbits));
  }
  int nmantbits = ntablebits - nexpbits;
  if (nexpbits > (int)sizeof(float) * 8 - 24) error->all(FLERR, "Too many exponent bits for lookup table");
  if (nmantbits + 1 > 24) error->all(FLERR, "Too many mantissa bits for lookup table");
  if (nmantbits < 3) error->all(FLERR, "Too few bits for lookup table");
  nshiftbits = 23 - nmantbits;
  nmask = 1;
  for (int j = 0; j < ntablebits + nshiftbits; j++) nmask *= 2;
  nmask -= 1;
  union_int_float_t rsq_lookup;
  rsq_lookup.f = outer * outer;
  maskhi = rsq_lookup.i & ~(nmask);
  rsq_lookup.f = inner * inner;
  masklo = rsq_lookup.i & ~(nmask);
}

}  // namespace LAMMPS_NS
