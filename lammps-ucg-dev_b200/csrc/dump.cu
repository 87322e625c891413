// dump.cu — the taps either side of the hot path, on the device:
//
//   ucgb200_dump_pack          DumpCustom::count + pack + Dump::sort   dump_custom.cpp:721-1386, 3552-3577
//                              compute property/atom columns           UCG/atom_vec_ucg.cpp:172-234
//   ucgb200_atoms_update_by_tag  ReadDump::process_atoms + migrate_atoms_by_coords   read_dump.cpp:797-935, 1150-1163
//
// A dump never moves the per-site arrays to the host: the selection (group, thresholds), the ordering
// (host index order of an unsorted serial dump, or ascending id for `dump_modify sort id`) and the
// gather of the chosen columns into the row-major double buffer DumpCustom calls `buf` all run here;
// only nchoose x ncols doubles cross PCIe.  read_dump goes the other way: the parsed snapshot rows are
// matched to the resident sites by id (sorted ids + binary search replace Atom::map) and scattered
// into the records in place.
#include <cub/device/device_radix_sort.cuh>

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "dump_format.cuh"
#include "ucg_internal.cuh"

using namespace ucg;

namespace {

constexpr int MAXCOL = UCGB200_DUMP_MAXCOL;
constexpr int MAXTHRESH = UCGB200_DUMP_MAXTHRESH;
constexpr int MAXTYPE = 64;

struct DumpArgs {
  const double4 *pos, *vel, *frc;
  const double *ucgp, *ucgml;
  const int *ts, *mask, *tag, *mol, *orig;
  double boxlo[3], invprd[3];
  double mass[MAXTYPE + 1];
  int ntypes;
  int rank;
  int groupbit;
  int ncols;
  int cols[MAXCOL], colbit[MAXCOL];
  int nthresh;
  int tcol[MAXTHRESH], top[MAXTHRESH];
  double tval[MAXTHRESH];
};

// one column of one site, as the double DumpCustom::pack_*() stores (ints are converted exactly as
// `buf[n] = ucgstate[clist[i]]` does, dump_custom.cpp:3556)
__device__ __forceinline__ double column_value(const DumpArgs &a, int s, int col, int colbit) {
  switch (col) {
    case UCGB200_COL_ID: return (double)a.tag[s];
    case UCGB200_COL_MOL: return (double)a.mol[s];
    case UCGB200_COL_TYPE: return (double)(a.ts[s] & 0xffff);
    case UCGB200_COL_MASS: { int t = a.ts[s] & 0xffff; return t <= a.ntypes ? a.mass[t] : 0.0; }
    case UCGB200_COL_X: return a.pos[s].x;
    case UCGB200_COL_Y: return a.pos[s].y;
    case UCGB200_COL_Z: return a.pos[s].z;
    // pack_xs: (x[j][0] - boxxlo) * invxprd   dump_custom.cpp:2606-2649
    case UCGB200_COL_XS: return __dmul_rn(__dsub_rn(a.pos[s].x, a.boxlo[0]), a.invprd[0]);
    case UCGB200_COL_YS: return __dmul_rn(__dsub_rn(a.pos[s].y, a.boxlo[1]), a.invprd[1]);
    case UCGB200_COL_ZS: return __dmul_rn(__dsub_rn(a.pos[s].z, a.boxlo[2]), a.invprd[2]);
    case UCGB200_COL_VX: return a.vel[s].x;
    case UCGB200_COL_VY: return a.vel[s].y;
    case UCGB200_COL_VZ: return a.vel[s].z;
    case UCGB200_COL_FX: return a.frc[s].x;
    case UCGB200_COL_FY: return a.frc[s].y;
    case UCGB200_COL_FZ: return a.frc[s].z;
    case UCGB200_COL_UCGSTATE: return (double)((a.ts[s] >> 16) & 1);
    case UCGB200_COL_UCGL: return a.pos[s].w;
    case UCGB200_COL_UCGP: return a.ucgp[s];
    case UCGB200_COL_PROC: return (double)a.rank;
    case UCGB200_COL_Q: return 0.0;
    default: break;
  }
  // compute property/atom: zero outside the COMPUTE's group (atom_vec_ucg.cpp:184-231)
  if (!(a.mask[s] & colbit)) return 0.0;
  switch (col) {
    case UCGB200_COL_P_UCGSTATE: return (double)((a.ts[s] >> 16) & 1);
    case UCGB200_COL_P_UCGL: return a.pos[s].w;
    case UCGB200_COL_P_UCGFORCE: return a.frc[s].w;
    case UCGB200_COL_P_UCGVL: return a.vel[s].w;
    case UCGB200_COL_P_UCGP: return a.ucgp[s];
    case UCGB200_COL_P_UCGML: return a.ucgml[s];
    default: return 0.0;
  }
}

// `if (choose[i] && <negated op>) choose[i] = 0`   dump_custom.cpp:1289-1345
__device__ __forceinline__ bool thresh_keeps(int op, double v, double value) {
  switch (op) {
    case UCGB200_THRESH_LT: return !(v >= value);
    case UCGB200_THRESH_LE: return !(v > value);
    case UCGB200_THRESH_GT: return !(v <= value);
    case UCGB200_THRESH_GE: return !(v < value);
    case UCGB200_THRESH_EQ: return !(v != value);
    case UCGB200_THRESH_NEQ: return !(v == value);
    default: return !((v == 0.0 && value == 0.0) || (v != 0.0 && value != 0.0));   // XOR
  }
}

// DumpCustom::count(): group mask, then every threshold; chosen sites get their sort key (host index or
// id, either is unique), the others the sentinel that sorts them behind every chosen one
__global__ void k_dump_choose(DumpArgs a, int n, int by_id, unsigned *__restrict__ keys, int *__restrict__ vals,
                              int *__restrict__ nchoose) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (s < n) {
    keep = (a.mask[s] & a.groupbit) != 0;
    for (int t = 0; keep && t < a.nthresh; t++) keep = thresh_keeps(a.top[t], column_value(a, s, a.tcol[t], ~0), a.tval[t]);
    keys[s] = keep ? (unsigned)(by_id ? a.tag[s] : a.orig[s]) : 0xffffffffu;
    vals[s] = s;
  }
  unsigned m = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nchoose, __popc(m));
}

// DumpCustom::pack(): one thread per buffer element, so a warp writes 32 consecutive doubles
__global__ void k_dump_pack(DumpArgs a, const int *__restrict__ sites, long long nelem, double *__restrict__ buf) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  int row = (int)(e / a.ncols), c = (int)(e - (long long)row * a.ncols);
  buf[e] = column_value(a, sites[row], a.cols[c], a.colbit[c]);
}

int fill_args(ucgb200_ctx *c, const ucgb200_dump_spec *sp, DumpArgs &a) {
  if (sp->ncols < 1 || sp->ncols > MAXCOL) return fail(c, "dump: between 1 and UCGB200_DUMP_MAXCOL columns");
  if (sp->nthresh < 0 || sp->nthresh > MAXTHRESH) return fail(c, "dump: too many thresholds");
  a.pos = c->pos.p; a.vel = c->vel.p; a.frc = c->frc.p; a.ucgp = c->ucgp.p; a.ucgml = c->ucgml.p;
  a.ts = c->ts.p; a.mask = c->mask.p; a.tag = c->tag.p; a.mol = c->mol.p; a.orig = c->orig.p;
  for (int d = 0; d < 3; d++) { a.boxlo[d] = c->boxlo[d]; a.invprd[d] = 1.0 / c->prd[d]; }
  const std::vector<double> &mass = c->mass.empty() ? c->dens.mass : c->mass;
  a.ntypes = (int)mass.size() - 1 > MAXTYPE ? MAXTYPE : (int)mass.size() - 1;
  if (a.ntypes < 0) a.ntypes = 0;
  for (int t = 0; t <= MAXTYPE; t++) a.mass[t] = t <= a.ntypes && t < (int)mass.size() ? mass[t] : 0.0;
  a.rank = c->halo.rank;
  a.groupbit = sp->groupbit;
  a.ncols = sp->ncols;
  for (int k = 0; k < sp->ncols; k++) {
    if (sp->cols[k] < 0 || sp->cols[k] >= UCGB200_COL_COUNT) return fail(c, "dump: unknown column code");
    a.cols[k] = sp->cols[k];
    a.colbit[k] = sp->col_groupbit ? sp->col_groupbit[k] : ~0;
  }
  a.nthresh = sp->nthresh;
  for (int t = 0; t < sp->nthresh; t++) {
    if (sp->thresh_col[t] < 0 || sp->thresh_col[t] >= UCGB200_COL_COUNT) return fail(c, "dump: threshold on an unknown attribute");
    if (sp->thresh_op[t] < 0 || sp->thresh_op[t] > UCGB200_THRESH_XOR) return fail(c, "dump: unknown threshold operation");
    a.tcol[t] = sp->thresh_col[t]; a.top[t] = sp->thresh_op[t]; a.tval[t] = sp->thresh_value[t];
  }
  return 0;
}

// selection + ordering: leaves the chosen sites, in output order, in c->dump_sites.p[0 .. nchoose)
int choose_and_order(ucgb200_ctx *c, const ucgb200_dump_spec *sp, const DumpArgs &a, int *nchoose) {
  const int n = c->nlocal;
  cudaStream_t st = c->stream;
  *nchoose = 0;
  if (n == 0) return 0;
  UCG_CHECK(c, c->dump_keys.ensure(2 * (size_t)n + 4));
  UCG_CHECK(c, c->dump_sites.ensure(2 * (size_t)n + 4));
  unsigned *k_in = c->dump_keys.p, *k_out = c->dump_keys.p + n;
  int *v_in = c->dump_sites.p + n, *v_out = c->dump_sites.p;
  int *d_n = c->d_flags.p + 4;   // a spare word of the flag block
  UCG_CHECK(c, cudaMemsetAsync(d_n, 0, sizeof(int), st));
  k_dump_choose<<<nblocks(n, 256), 256, 0, st>>>(a, n, sp->order == UCGB200_DUMP_ORDER_ID, k_in, v_in, d_n);
  UCG_LAUNCHED(c);
  size_t tmp = 0;
  UCG_CHECK(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, n, 0, 32, st));
  UCG_CHECK(c, c->dump_tmp.ensure(tmp + 16));
  UCG_CHECK(c, cub::DeviceRadixSort::SortPairs(c->dump_tmp.p, tmp, k_in, k_out, v_in, v_out, n, 0, 32, st));
  c->launches += 7;   // histogram + 4 onesweep passes (+ scan) of the radix sort
  UCG_CHECK(c, cudaMemcpyAsync(nchoose, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  return 0;
}

}  // namespace

extern "C" int ucgb200_dump_count(ucgb200_ctx *c, const ucgb200_dump_spec *sp, long long *nrows) {
  if (!c || !sp || !nrows) return -1;
  cudaSetDevice(c->device);
  DumpArgs a;
  int rc = fill_args(c, sp, a);
  if (rc) return rc;
  int nch = 0;
  if ((rc = choose_and_order(c, sp, a, &nch))) return rc;
  *nrows = nch;
  return 0;
}

// device half of a dump: the packed rows stay in c->dump_buf (used by the device text formatter too)
int ucg_dump_pack_device(ucgb200_ctx *c, const ucgb200_dump_spec *sp, long long *nrows) {
  DumpArgs a;
  int rc = fill_args(c, sp, a);
  if (rc) return rc;
  int nch = 0;
  if ((rc = choose_and_order(c, sp, a, &nch))) return rc;
  *nrows = nch;
  if (nch == 0) return 0;
  const long long nelem = (long long)nch * sp->ncols;
  UCG_CHECK(c, c->dump_buf.ensure((size_t)nelem + 8));
  k_dump_pack<<<nblocks(nelem, 256), 256, 0, c->stream>>>(a, c->dump_sites.p, nelem, c->dump_buf.p);
  UCG_LAUNCHED(c);
  return 0;
}

extern "C" int ucgb200_dump_pack(ucgb200_ctx *c, const ucgb200_dump_spec *sp, double *buf, long long cap_rows,
                                 long long *nrows) {
  if (!c || !sp || !nrows) return -1;
  cudaSetDevice(c->device);
  int rc = ucg_dump_pack_device(c, sp, nrows);
  if (rc) return rc;
  if (*nrows == 0) return 0;
  if (!buf || cap_rows < *nrows) return fail(c, "dump_pack: host buffer smaller than the number of selected atoms");
  UCG_CHECK(c, cudaMemcpyAsync(buf, c->dump_buf.p, (size_t)*nrows * sp->ncols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------ text
namespace {

__host__ __device__ inline bool column_is_int(int col) {   // vtype[] of DumpCustom::parse_fields (dump_custom.cpp:1485-1687)
  return col == UCGB200_COL_ID || col == UCGB200_COL_MOL || col == UCGB200_COL_TYPE || col == UCGB200_COL_PROC ||
         col == UCGB200_COL_UCGSTATE;
}

struct ColTypes { int ncols; unsigned char is_int[MAXCOL]; };

// one 16-byte slot per buffer element: up to 13 characters, the length in the last byte
__global__ void k_dump_format(const double *__restrict__ buf, long long nelem, ColTypes ct, uint4 *__restrict__ slots) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  union { char ch[16]; uint4 v; } u;
  u.v = make_uint4(0, 0, 0, 0);
  const int c = (int)(e % ct.ncols);
  const double val = buf[e];
  int len = ct.is_int[c] ? ucgfmt::format_d((int)val, u.ch) : ucgfmt::format_g(val, u.ch);
  u.ch[15] = (char)len;
  slots[e] = u.v;
}
__global__ void k_dump_rowlen(const uint4 *__restrict__ slots, long long nrows, int ncols, long long *__restrict__ rowlen) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const unsigned char *p = (const unsigned char *)(slots + r * ncols);
  int len = ncols;   // one blank after every column but the last, one newline
  for (int c = 0; c < ncols; c++) len += p[16 * c + 15];
  rowlen[r] = len;
}
__global__ void k_dump_emit(const uint4 *__restrict__ slots, long long nelem, int ncols, const long long *__restrict__ rowoff,
                            char *__restrict__ text) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  const long long r = e / ncols;
  const int c = (int)(e - r * ncols);
  const unsigned char *row = (const unsigned char *)(slots + r * ncols);
  long long off = rowoff[r];
  for (int j = 0; j < c; j++) off += row[16 * j + 15] + 1;
  const unsigned char *me = row + 16 * c;
  const int len = me[15];
  for (int i = 0; i < len; i++) text[off + i] = (char)me[i];
  text[off + len] = c == ncols - 1 ? '\n' : ' ';
}

}  // namespace

extern "C" int ucgb200_dump_text(ucgb200_ctx *c, const ucgb200_dump_spec *sp, char *text, long long cap, long long *nrows,
                                 long long *nbytes) {
  if (!c || !sp || !nrows || !nbytes) return -1;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  *nbytes = 0;
  c->dump_text_bytes = 0;
  int rc = ucg_dump_pack_device(c, sp, nrows);
  if (rc) return rc;
  const long long nr = *nrows;
  if (nr == 0) return 0;
  const int nc = sp->ncols;
  const long long nelem = nr * nc;
  ColTypes ct;
  ct.ncols = nc;
  for (int k = 0; k < nc; k++) ct.is_int[k] = column_is_int(sp->cols[k]);
  // slots live behind the radix-sort scratch; row lengths / offsets in dump_off
  UCG_CHECK(c, c->dump_slots.ensure((size_t)nelem + 4));
  UCG_CHECK(c, c->dump_off.ensure(2 * (size_t)nr + 8));
  long long *rowlen = c->dump_off.p, *rowoff = c->dump_off.p + nr + 1;
  k_dump_format<<<nblocks(nelem, 256), 256, 0, st>>>(c->dump_buf.p, nelem, ct, c->dump_slots.p);
  UCG_LAUNCHED(c);
  k_dump_rowlen<<<nblocks(nr, 256), 256, 0, st>>>(c->dump_slots.p, nr, nc, rowlen);
  UCG_LAUNCHED(c);
  size_t tmp = 0;
  UCG_CHECK(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp, rowlen, rowoff, (int)nr, st));
  UCG_CHECK(c, c->dump_tmp.ensure(tmp + 16));
  UCG_CHECK(c, cub::DeviceScan::ExclusiveSum(c->dump_tmp.p, tmp, rowlen, rowoff, (int)nr, st));
  c->launches += 2;
  long long last[2];
  UCG_CHECK(c, cudaMemcpyAsync(&last[0], rowoff + nr - 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaMemcpyAsync(&last[1], rowlen + nr - 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  const long long total = last[0] + last[1];
  UCG_CHECK(c, c->dump_text.ensure((size_t)total + 16));
  k_dump_emit<<<nblocks(nelem, 256), 256, 0, st>>>(c->dump_slots.p, nelem, nc, rowoff, c->dump_text.p);
  UCG_LAUNCHED(c);
  c->dump_text_bytes = total;
  *nbytes = total;
  if (!text) return 0;   // size query: fetch with ucgb200_dump_text_copy
  return ucgb200_dump_text_copy(c, text, cap);
}

extern "C" int ucgb200_dump_text_copy(ucgb200_ctx *c, char *text, long long cap) {
  if (!c || !text) return -1;
  cudaSetDevice(c->device);
  if (cap < c->dump_text_bytes) return fail(c, "dump_text: host buffer smaller than the formatted text");
  if (c->dump_text_bytes == 0) return 0;
  UCG_CHECK(c, cudaMemcpyAsync(text, c->dump_text.p, (size_t)c->dump_text_bytes, cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------ read_dump
namespace {

struct UpdateArgs {
  double4 *pos, *vel, *frc;
  double *ucgp;
  int *ts;
  const int *orig;
  int nfield;
  int ftype[MAXCOL];
  int scaled;
  double snaplo[3], snapprd[3];   // snapshot box: xfield()/yfield()/zfield() unscaling, read_dump.cpp:1359-1380
  double boxlo[3], boxhi[3], prd[3];
  int periodic[3];
};

__device__ __forceinline__ int find_tag(const unsigned *__restrict__ sorted, int n, unsigned tag) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (sorted[mid] < tag) lo = mid + 1; else hi = mid;
  }
  return (lo < n && sorted[lo] == tag) ? lo : -1;
}

// one snapshot row per thread: `m = atom->map(mtag)`, then the field switch of read_dump.cpp:863-913 and the
// remap of migrate_atoms_by_coords (Domain::remap: as many periods as it takes, then clamp to the low face)
__global__ void k_update_by_tag(UpdateArgs a, const double *__restrict__ fields, int nnew, const unsigned *__restrict__ sorted_tags,
                                const int *__restrict__ sorted_sites, int nlocal, int *__restrict__ updated, int *__restrict__ nreplace) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnew) return;
  const double *row = fields + (size_t)i * a.nfield;
  int mtag = (int)row[0];
  if (mtag <= 0) return;
  int k = find_tag(sorted_tags, nlocal, (unsigned)mtag);
  if (k < 0) return;
  int s = sorted_sites[k];
  updated[a.orig[s]] = 1;
  atomicAdd(nreplace, 1);
  double4 r = a.pos[s], v = a.vel[s], f = a.frc[s];
  bool moved = false;
  for (int j = 1; j < a.nfield; j++) {
    double val = row[j];
    switch (a.ftype[j]) {
      case UCGB200_COL_X: r.x = a.scaled ? __dadd_rn(__dmul_rn(val, a.snapprd[0]), a.snaplo[0]) : val; moved = true; break;
      case UCGB200_COL_Y: r.y = a.scaled ? __dadd_rn(__dmul_rn(val, a.snapprd[1]), a.snaplo[1]) : val; moved = true; break;
      case UCGB200_COL_Z: r.z = a.scaled ? __dadd_rn(__dmul_rn(val, a.snapprd[2]), a.snaplo[2]) : val; moved = true; break;
      case UCGB200_COL_VX: v.x = val; break;
      case UCGB200_COL_VY: v.y = val; break;
      case UCGB200_COL_VZ: v.z = val; break;
      case UCGB200_COL_FX: f.x = val; break;
      case UCGB200_COL_FY: f.y = val; break;
      case UCGB200_COL_FZ: f.z = val; break;
      case UCGB200_COL_UCGSTATE: { int st = (int)val; st = st < 0 ? 0 : (st > 1 ? 1 : st); a.ts[s] = (a.ts[s] & 0xffff) | (st << 16); break; }
      case UCGB200_COL_UCGL: r.w = val; break;
      case UCGB200_COL_UCGP: a.ucgp[s] = val; break;
      default: break;   // id, type, q: not replaced (type only matters for `add`)
    }
  }
  if (moved) {
    double x[3] = {r.x, r.y, r.z};
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (!a.periodic[d]) continue;
      while (x[d] < a.boxlo[d]) x[d] += a.prd[d];
      while (x[d] >= a.boxhi[d]) x[d] -= a.prd[d];
      x[d] = fmax(x[d], a.boxlo[d]);
    }
    r.x = x[0]; r.y = x[1]; r.z = x[2];
  }
  a.pos[s] = r; a.vel[s] = v; a.frc[s] = f;
}

// ReadDump::migrate_atoms_by_coords (read_dump.cpp:1150-1163): Domain::remap of EVERY owned atom into the current box
__global__ void k_remap_all(double4 *__restrict__ pos, int n, UpdateArgs a) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 r = pos[s];
  double x[3] = {r.x, r.y, r.z};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (!a.periodic[d]) continue;
    while (x[d] < a.boxlo[d]) x[d] += a.prd[d];
    while (x[d] >= a.boxhi[d]) x[d] -= a.prd[d];
    x[d] = fmax(x[d], a.boxlo[d]);
  }
  r.x = x[0]; r.y = x[1]; r.z = x[2];
  pos[s] = r;
}

__global__ void k_tag_keys(const int *__restrict__ tag, int n, unsigned *__restrict__ keys, int *__restrict__ vals) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) { keys[s] = (unsigned)tag[s]; vals[s] = s; }
}

}  // namespace

extern "C" int ucgb200_atoms_update_by_tag(ucgb200_ctx *c, int nnew, int nfield, const int *fieldtype, const double *fields,
                                           int scaled, const double snap_lo[3], const double snap_hi[3], int *updated,
                                           long long *nreplace) {
  if (!c || nnew < 0 || nfield < 1 || nfield > MAXCOL || !fieldtype) return -1;
  if (fieldtype[0] != UCGB200_COL_ID) return fail(c, "atoms_update_by_tag: field 0 must be the atom id");
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  const int n = c->nlocal;
  if (nreplace) *nreplace = 0;
  if (updated) for (int i = 0; i < n; i++) updated[i] = 0;
  if (n == 0 || nnew == 0) return 0;
  if (!fields && (c->parsed_rows != nnew || c->parsed_fields != nfield))
    return fail(c, "atoms_update_by_tag: no host block given and no matching block parsed on the device");
  UpdateArgs a;
  a.pos = c->pos.p; a.vel = c->vel.p; a.frc = c->frc.p; a.ucgp = c->ucgp.p; a.ts = c->ts.p; a.orig = c->orig.p;
  a.nfield = nfield;
  bool moves = false;
  for (int j = 0; j < nfield; j++) {
    a.ftype[j] = fieldtype[j];
    if (fieldtype[j] == UCGB200_COL_X || fieldtype[j] == UCGB200_COL_Y || fieldtype[j] == UCGB200_COL_Z) moves = true;
  }
  a.scaled = scaled;
  if (scaled && (!snap_lo || !snap_hi)) return fail(c, "atoms_update_by_tag: scaled coordinates need the snapshot box");
  for (int d = 0; d < 3; d++) {
    a.snaplo[d] = snap_lo ? snap_lo[d] : 0.0;
    a.snapprd[d] = (snap_lo && snap_hi) ? snap_hi[d] - snap_lo[d] : 1.0;
    a.boxlo[d] = c->boxlo[d]; a.boxhi[d] = c->boxhi[d]; a.prd[d] = c->prd[d]; a.periodic[d] = c->periodic[d];
  }
  // id -> site: sorted ids of the resident sites
  UCG_CHECK(c, c->dump_keys.ensure(2 * (size_t)n + 4));
  UCG_CHECK(c, c->dump_sites.ensure(2 * (size_t)n + 4));
  unsigned *k_in = c->dump_keys.p, *k_out = c->dump_keys.p + n;
  int *v_in = c->dump_sites.p + n, *v_out = c->dump_sites.p;
  k_tag_keys<<<nblocks(n, 256), 256, 0, st>>>(c->tag.p, n, k_in, v_in);
  UCG_LAUNCHED(c);
  size_t tmp = 0;
  UCG_CHECK(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, n, 0, 32, st));
  UCG_CHECK(c, c->dump_tmp.ensure(tmp + 16));
  UCG_CHECK(c, cub::DeviceRadixSort::SortPairs(c->dump_tmp.p, tmp, k_in, k_out, v_in, v_out, n, 0, 32, st));
  c->launches += 7;
  // snapshot rows and the per-host-index update flags
  const size_t nval = (size_t)nnew * nfield;
  UCG_CHECK(c, c->stage_i.ensure((size_t)n + 8));
  if (fields) {
    UCG_CHECK(c, c->dump_buf.ensure(nval + 8));
    UCG_CHECK(c, cudaMemcpyAsync(c->dump_buf.p, fields, nval * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  c->parsed_rows = -1;
  UCG_CHECK(c, cudaMemsetAsync(c->stage_i.p, 0, ((size_t)n + 1) * sizeof(int), st));
  int *d_flag = c->stage_i.p, *d_cnt = c->stage_i.p + n;
  k_update_by_tag<<<nblocks(nnew, 256), 256, 0, st>>>(a, c->dump_buf.p, nnew, k_out, v_out, n, d_flag, d_cnt);
  UCG_LAUNCHED(c);
  int cnt = 0;
  UCG_CHECK(c, cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (updated) UCG_CHECK(c, cudaMemcpyAsync(updated, d_flag, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  if (nreplace) *nreplace = cnt;
  if (moves) { c->list_valid = false; c->maxdisp_valid = false; }
  return 0;
}

extern "C" int ucgb200_atoms_remap(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  if (c->nlocal == 0) return 0;
  UpdateArgs a;
  memset(&a, 0, sizeof a);
  for (int d = 0; d < 3; d++) { a.boxlo[d] = c->boxlo[d]; a.boxhi[d] = c->boxhi[d]; a.prd[d] = c->prd[d]; a.periodic[d] = c->periodic[d]; }
  k_remap_all<<<nblocks(c->nlocal, 256), 256, 0, c->stream>>>(c->pos.p, c->nlocal, a);
  UCG_LAUNCHED(c);
  c->list_valid = false;
  c->maxdisp_valid = false;
  return 0;
}

// ------------------------------------------------------------------------------------ snapshot text -> fields
// ReaderNative::read_atoms (reader_native.cpp:486-500) on the device: the snapshot body is uploaded as text, line starts
// are compacted, and one thread per line tokenises it and converts the requested columns.  A decimal token converts
// exactly when its digits fit 2^53 and its power of ten is at most 10^22 in magnitude (one correctly rounded IEEE
// multiplication or division of two exact doubles — every number a "%g" dump holds, unless its exponent is beyond e+-22);
// any other token (longer mantissas, large exponents, inf/nan, malformed text) marks the row, and the host converts
// those rows with strtod, as std::stod does in the reference.
namespace {

struct IsLineStart {
  const char *text;
  __device__ bool operator()(const int &i) const { return i == 0 || text[i - 1] == '\n'; }
};

constexpr int MAXWORDS = 64;
struct ParseArgs {
  int nwords, nfield, maxcol;
  int fieldindex[MAXCOL];
};

__device__ __forceinline__ bool is_blank(char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || ch == '\f'; }

// returns false when the token needs strtod
__device__ bool parse_token(const char *p, const char *end, double *out) {
  const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17,
                          1e18, 1e19, 1e20, 1e21, 1e22};
  bool neg = false;
  if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
  unsigned long long mant = 0;
  int e10 = 0;
  bool digits = false;
  while (p < end && *p >= '0' && *p <= '9') {
    if (mant >= 900000000000000000ull) return false;
    mant = mant * 10ull + (unsigned)(*p - '0');
    digits = true;
    p++;
  }
  if (p < end && *p == '.') {
    p++;
    while (p < end && *p >= '0' && *p <= '9') {
      if (mant >= 900000000000000000ull) return false;
      mant = mant * 10ull + (unsigned)(*p - '0');
      e10--;
      digits = true;
      p++;
    }
  }
  if (!digits) return false;
  if (p < end && (*p == 'e' || *p == 'E')) {
    p++;
    bool eneg = false;
    if (p < end && (*p == '-' || *p == '+')) { eneg = *p == '-'; p++; }
    if (!(p < end && *p >= '0' && *p <= '9')) return false;
    int ex = 0;
    while (p < end && *p >= '0' && *p <= '9') {
      if (ex > 10000) return false;
      ex = ex * 10 + (*p - '0');
      p++;
    }
    e10 += eneg ? -ex : ex;
  }
  if (p < end && !is_blank(*p)) return false;   // trailing characters: let strtod decide
  if (mant == 0) { *out = neg ? -0.0 : 0.0; return true; }
  while (e10 > 22 && mant < 900719925474099ull) { mant *= 10ull; e10--; }
  while (e10 < -22 && mant % 10ull == 0) { mant /= 10ull; e10++; }
  if (mant > 9007199254740992ull || e10 > 22 || e10 < -22) return false;
  const double m = (double)mant;
  const double v = e10 >= 0 ? __dmul_rn(m, P10[e10]) : __ddiv_rn(m, P10[-e10]);
  *out = neg ? -v : v;
  return true;
}

__global__ void k_parse_rows(const char *__restrict__ text, int nbytes, const int *__restrict__ start, int nrows, ParseArgs a,
                             double *__restrict__ fields, int *__restrict__ counters, int *__restrict__ slow_rows,
                             int *__restrict__ slow_off, int cap) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const char *p = text + start[r];
  const char *end = text + (r + 1 < nrows ? start[r + 1] : nbytes);
  int tok[MAXWORDS];
  int nt = 0;
  const char *q = p;
  while (q < end && nt <= a.maxcol) {
    while (q < end && is_blank(*q)) q++;
    if (q >= end) break;
    tok[nt++] = (int)(q - p);
    while (q < end && !is_blank(*q)) q++;
  }
  bool slow = false;
  if (nt <= a.maxcol) {
    // fewer words than the columns needed: is the line short of nwords as a whole ("Insufficient columns")?
    atomicAdd(&counters[1], 1);
    return;
  }
  for (int m = 0; m < a.nfield; m++) {
    double v = 0.0;
    if (!parse_token(p + tok[a.fieldindex[m]], end, &v)) slow = true;
    fields[(size_t)r * a.nfield + m] = v;
  }
  if (slow) {
    int k = atomicAdd(&counters[0], 1);
    if (k < cap) { slow_rows[k] = r; slow_off[k] = start[r]; }
  }
}

__global__ void k_patch_rows(double *__restrict__ fields, int nfield, const int *__restrict__ rows, const double *__restrict__ vals, int n) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * nfield) return;
  fields[(size_t)rows[e / nfield] * nfield + e % nfield] = vals[e];
}

}  // namespace

extern "C" int ucgb200_snapshot_parse(ucgb200_ctx *c, const char *text, long long nbytes, long long nrows, int nwords, int nfield,
                                      const int *fieldindex, int slow_cap, int *slow_rows, int *slow_offsets, long long *nslow) {
  if (!c || !text || !fieldindex || !nslow || nfield < 1 || nfield > MAXCOL) return -1;
  *nslow = 0;
  c->parsed_rows = -1;
  if (nbytes >= 0x7fffffffLL || nrows >= 0x7fffffffLL || nwords > MAXWORDS) return UCGB200_PARSE_ON_HOST;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  if (nrows == 0) { c->parsed_rows = 0; c->parsed_fields = nfield; return 0; }
  ParseArgs a;
  a.nwords = nwords; a.nfield = nfield; a.maxcol = 0;
  for (int m = 0; m < nfield; m++) {
    if (fieldindex[m] < 0 || fieldindex[m] >= nwords) return fail(c, "snapshot_parse: column index outside the line");
    a.fieldindex[m] = fieldindex[m];
    a.maxcol = fieldindex[m] > a.maxcol ? fieldindex[m] : a.maxcol;
  }
  const int nb = (int)nbytes, nr = (int)nrows;
  UCG_CHECK(c, c->dump_text.ensure((size_t)nb + 16));
  UCG_CHECK(c, cudaMemcpyAsync(c->dump_text.p, text, (size_t)nb, cudaMemcpyHostToDevice, st));
  // line starts: offsets i with i == 0 or text[i-1] == '\n'
  UCG_CHECK(c, c->dump_sites.ensure((size_t)nb / 2 + (size_t)nr + 2 * (size_t)slow_cap + 64));
  int *starts = c->dump_sites.p;
  int *d_cnt = c->d_flags.p + 4;   // [4] lines found / slow rows, [5] short lines
  cub::CountingInputIterator<int> idx(0);
  IsLineStart pred{c->dump_text.p};
  // a line is at least 2 bytes ("0\n"), so nb/2+1 offsets always fit
  size_t tmp = 0;
  UCG_CHECK(c, cub::DeviceSelect::If(nullptr, tmp, idx, starts, d_cnt, nb, pred, st));
  UCG_CHECK(c, c->dump_tmp.ensure(tmp + 16));
  UCG_CHECK(c, cub::DeviceSelect::If(c->dump_tmp.p, tmp, idx, starts, d_cnt, nb, pred, st));
  c->launches += 2;
  int nlines = 0;
  UCG_CHECK(c, cudaMemcpyAsync(&nlines, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  if (nlines < nr) return fail(c, "Unexpected end of dump file");
  UCG_CHECK(c, c->dump_buf.ensure((size_t)nr * nfield + 8));
  int *d_slow_rows = c->dump_sites.p + (size_t)nb / 2 + nr + 8, *d_slow_off = d_slow_rows + slow_cap;
  UCG_CHECK(c, cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), st));
  k_parse_rows<<<nblocks(nr, 128), 128, 0, st>>>(c->dump_text.p, nb, starts, nr, a, c->dump_buf.p, d_cnt, d_slow_rows, d_slow_off, slow_cap);
  UCG_LAUNCHED(c);
  int cnt[2] = {0, 0};
  UCG_CHECK(c, cudaMemcpyAsync(cnt, d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  if (cnt[1]) return fail(c, "Insufficient columns in dump file");
  if (cnt[0] > slow_cap) return UCGB200_PARSE_ON_HOST;   // mostly non-trivial numbers: the host cores convert the whole block
  if (cnt[0]) {
    UCG_CHECK(c, cudaMemcpyAsync(slow_rows, d_slow_rows, cnt[0] * sizeof(int), cudaMemcpyDeviceToHost, st));
    UCG_CHECK(c, cudaMemcpyAsync(slow_offsets, d_slow_off, cnt[0] * sizeof(int), cudaMemcpyDeviceToHost, st));
    UCG_CHECK(c, cudaStreamSynchronize(st));
  }
  *nslow = cnt[0];
  c->parsed_rows = nr;
  c->parsed_fields = nfield;
  return 0;
}

extern "C" int ucgb200_snapshot_patch(ucgb200_ctx *c, int n, const int *rows, const double *values) {
  if (!c || n < 0) return -1;
  if (n == 0) return 0;
  if (c->parsed_rows < 0 || !rows || !values) return fail(c, "snapshot_patch: nothing parsed");
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  const int nf = c->parsed_fields;
  UCG_CHECK(c, c->stage_i.ensure((size_t)n + 8));
  UCG_CHECK(c, c->stage_d.ensure((size_t)n * nf + 8));
  UCG_CHECK(c, cudaMemcpyAsync(c->stage_i.p, rows, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  UCG_CHECK(c, cudaMemcpyAsync(c->stage_d.p, values, (size_t)n * nf * sizeof(double), cudaMemcpyHostToDevice, st));
  k_patch_rows<<<nblocks((long long)n * nf, 256), 256, 0, st>>>(c->dump_buf.p, nf, c->stage_i.p, c->stage_d.p, n);
  UCG_LAUNCHED(c);
  UCG_CHECK(c, cudaStreamSynchronize(st));
  return 0;
}
