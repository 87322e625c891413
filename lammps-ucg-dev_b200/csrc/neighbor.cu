// neighbor.cu — on-GPU replacement for what stock LAMMPS does around the UCG styles on
// a rebuild step: Domain::pbc, spatial sort, Comm::borders (periodic ghost images),
// binned neighbor-list build, and the Neighbor::decide skin check; plus the per-step
// Comm::forward_comm ghost refresh (fields_comm, UCG/atom_vec_ucg.cpp:71).
//
// Requested by the reference at pair_table_ucgld.cpp:868 (half list, newton on),
// pair_table_rleucg_interface.cpp:779 and pair_table_ucg_bethe_density.cpp:1135 (full),
// fix_cluster_switch.cpp:396 (full).  The device list is always FULL (every neighbor of
// every owned site): the pair kernels accumulate into the centre site only, so no
// reverse communication and no FP64 atomics are needed; the reference's half list is
// the subset {j "above" i} of these rows and is what the parity tests compare.
//
// Membership test is bit-exact with the reference: rsq = dx*dx + dy*dy + dz*dz without
// FMA contraction, rsq <= (sqrt(cutsq[itype][jtype]) + skin)^2.
#include "ucg_internal.cuh"

#include <algorithm>
#include <cmath>

using namespace ucg;

// ------------------------------------------------------------- exclusive scan
namespace {

constexpr int SCAN_BS = 512;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_BS * SCAN_IPT;

__global__ void k_scan_tiles(const int *__restrict__ in, int *__restrict__ out, int n, int *__restrict__ tile_sums) {
  __shared__ int warp_sums[SCAN_BS / 32];
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
  int v[SCAN_IPT];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_IPT; k++) {
    int idx = base + k;
    v[k] = idx < n ? in[idx] : 0;
    sum += v[k];
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < SCAN_BS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < SCAN_BS / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  int excl = inc - sum + (wid ? warp_sums[wid - 1] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_IPT; k++) {
    int idx = base + k;
    if (idx < n) out[idx] = excl;
    excl += v[k];
  }
  if (threadIdx.x == SCAN_BS - 1) tile_sums[blockIdx.x] = excl;
}
__global__ void k_scan_add(int *__restrict__ out, int n, const int *__restrict__ tile_off) {
  int idx = blockIdx.x * SCAN_TILE + threadIdx.x;
  int off = tile_off[blockIdx.x];
  for (int k = 0; k < SCAN_IPT; k++, idx += SCAN_BS)
    if (idx < n) out[idx] += off;
}
__global__ void k_store_total(const int *scanned, const int *in, int n, int *total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total = n ? scanned[n - 1] + in[n - 1] : 0;
}

}  // namespace

// out[i] = sum_{k<i} in[k]; optional total into *d_total.  Multi-level, deterministic.
int ucg::exclusive_scan(ucgb200_ctx *c, const int *in, int *out, int n, int *d_total) {
  if (n <= 0) {
    if (d_total) UCG_CHECK(c, cudaMemsetAsync(d_total, 0, sizeof(int), c->stream));
    return 0;
  }
  // level sizes
  std::vector<int> sizes;
  int m = n;
  while (true) {
    int tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    sizes.push_back(tiles);
    if (tiles == 1) break;
    m = tiles;
  }
  size_t need = 0;
  for (int s : sizes) need += 2 * (size_t)s + 8;
  UCG_CHECK(c, c->scan_tmp.ensure(need));
  // level 0
  std::vector<int *> sums(sizes.size()), scanned(sizes.size());
  int *p = c->scan_tmp.p;
  for (size_t l = 0; l < sizes.size(); l++) { sums[l] = p; p += sizes[l] + 4; scanned[l] = p; p += sizes[l] + 4; }
  const int *cur_in = in;
  int *cur_out = out;
  int cur_n = n;
  for (size_t l = 0; l < sizes.size(); l++) {
    k_scan_tiles<<<sizes[l], SCAN_BS, 0, c->stream>>>(cur_in, cur_out, cur_n, sums[l]);
    UCG_LAUNCHED(c);
    cur_in = sums[l]; cur_out = scanned[l]; cur_n = sizes[l];
  }
  // top level has one tile: its scanned offsets are {0}; walk back down
  UCG_CHECK(c, cudaMemsetAsync(scanned[sizes.size() - 1], 0, sizeof(int), c->stream));
  for (int l = (int)sizes.size() - 2; l >= 0; l--) {
    // scanned[l] currently holds the within-tile scan of sums[l]; add offsets of level l+1
    k_scan_add<<<sizes[l + 1], SCAN_BS, 0, c->stream>>>(scanned[l], sizes[l], scanned[l + 1]);
    UCG_LAUNCHED(c);
  }
  if (sizes.size() > 1 || true) {
    k_scan_add<<<sizes[0], SCAN_BS, 0, c->stream>>>(out, n, scanned[0]);
    UCG_LAUNCHED(c);
  }
  if (d_total) {
    k_store_total<<<1, 32, 0, c->stream>>>(out, in, n, d_total);
    UCG_LAUNCHED(c);
  }
  return 0;
}

// ------------------------------------------------------------------- kernels
namespace {

struct BoxDev {
  double lo[3], hi[3], prd[3];
  double sublo[3], subhi[3];
  int periodic[3];
  double cutghost;
};

// [stock] Domain::pbc: wrap owned atoms back into the periodic box, then bin them.
__global__ void k_wrap_count(double4 *__restrict__ pos, int n, BoxDev box, Grid g, int *__restrict__ cell_of,
                             int *__restrict__ cell_count, int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 r = pos[i];
  double x[3] = {r.x, r.y, r.z};
  bool lost = false;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (box.periodic[d]) {
      if (x[d] < box.lo[d]) x[d] += box.prd[d];
      if (x[d] >= box.hi[d]) {
        x[d] -= box.prd[d];
        x[d] = fmax(x[d], box.lo[d]);
      }
    }
    if (!(x[d] >= box.lo[d] - box.prd[d]) || !(x[d] < box.hi[d] + box.prd[d])) lost = true;
  }
  if (lost) flags[3] = 1;
  r.x = x[0]; r.y = x[1]; r.z = x[2];
  pos[i] = r;
  int cid = cell_index(g, r.x, r.y, r.z, true);
  cell_of[i] = cid;
  atomicAdd(&cell_count[cid], 1);
}

__global__ void k_fill_order(const int *__restrict__ cell_of, int n, const int *__restrict__ cell_start,
                             int *__restrict__ cursor, int *__restrict__ order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cid = cell_of[i];
  int slot = cell_start[cid] + atomicAdd(&cursor[cid], 1);
  order[slot] = i;
}

// restore a deterministic order inside every cell (ascending previous index / key)
template <class T>
__global__ void k_sort_cells(T *__restrict__ keys, const int *__restrict__ start, int ncells) {
  int cid = blockIdx.x * blockDim.x + threadIdx.x;
  if (cid >= ncells) return;
  int b = start[cid], e = start[cid + 1];
  for (int a = b + 1; a < e; a++) {
    T key = keys[a];
    int k = a - 1;
    while (k >= b && keys[k] > key) { keys[k + 1] = keys[k]; k--; }
    keys[k + 1] = key;
  }
}

struct PermArgs {
  const double4 *pos, *vel, *frc;
  const double2 *scores;
  const double *ucgp, *ucgml;
  const int *ts, *mask, *tag, *mol, *orig;
  double4 *pos_o, *vel_o, *frc_o;
  double2 *scores_o;
  double *ucgp_o, *ucgml_o;
  int *ts_o, *mask_o, *tag_o, *mol_o, *orig_o;
};
__global__ void k_permute(PermArgs a, const int *__restrict__ order, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int i = order[s];
  a.pos_o[s] = a.pos[i]; a.vel_o[s] = a.vel[i]; a.frc_o[s] = a.frc[i];
  a.scores_o[s] = a.scores[i];
  a.ucgp_o[s] = a.ucgp[i]; a.ucgml_o[s] = a.ucgml[i];
  a.ts_o[s] = a.ts[i]; a.mask_o[s] = a.mask[i]; a.tag_o[s] = a.tag[i]; a.mol_o[s] = a.mol[i];
  a.orig_o[s] = a.orig[i];
}

// [stock] CommBrick::borders, periodic self-images: an owned atom within cutghost of a
// low face reappears at +prd, of a high face at -prd, in every combination of the three
// dimensions (the staged x,y,z exchange of LAMMPS yields exactly these images).
__device__ __forceinline__ int image_options(double x, double lo, double hi, double cut, int periodic, int opt[3]) {
  int n = 0;
  opt[n++] = 0;
  if (periodic) {
    if (x <= lo + cut) opt[n++] = 1;
    if (x >= hi - cut) opt[n++] = -1;
  }
  return n;
}
template <bool FILL>
__global__ void k_ghost_images(const double4 *__restrict__ pos, int n, BoxDev box, Grid g,
                               int *__restrict__ gcell_count, const int *__restrict__ gcell_start,
                               int *__restrict__ gcursor, long long *__restrict__ gkey) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 r = pos[i];
  int ox[3], oy[3], oz[3];
  int nx = image_options(r.x, box.sublo[0], box.subhi[0], box.cutghost, box.periodic[0], ox);
  int ny = image_options(r.y, box.sublo[1], box.subhi[1], box.cutghost, box.periodic[1], oy);
  int nz = image_options(r.z, box.sublo[2], box.subhi[2], box.cutghost, box.periodic[2], oz);
  if (nx * ny * nz == 1) return;
  for (int c = 0; c < nz; c++)
    for (int b = 0; b < ny; b++)
      for (int a = 0; a < nx; a++) {
        int sx = ox[a], sy = oy[b], sz = oz[c];
        if (sx == 0 && sy == 0 && sz == 0) continue;
        double gx = sx ? r.x + sx * box.prd[0] : r.x;
        double gy = sy ? r.y + sy * box.prd[1] : r.y;
        double gz = sz ? r.z + sz * box.prd[2] : r.z;
        int cid = cell_index(g, gx, gy, gz, false);
        if (!FILL) atomicAdd(&gcell_count[cid], 1);
        else {
          int slot = gcell_start[cid] + atomicAdd(&gcursor[cid], 1);
          int code = (sx + 1) + 3 * (sy + 1) + 9 * (sz + 1);
          gkey[slot] = (long long)i * 32 + code;
        }
      }
}

// ghost records from their owners: border payload (fields_border, atom_vec_ucg.cpp:66-67)
// when FULL, per-step forward payload (fields_comm, :71) otherwise.
template <bool FULL>
__global__ void k_ghost_fill(double4 *__restrict__ pos, int *__restrict__ ts, double *__restrict__ ucgp,
                             int *__restrict__ tag, int *__restrict__ mol, int nlocal, int nghost,
                             const long long *__restrict__ gkey, int *__restrict__ gowner, int *__restrict__ gcode,
                             double px, double py, double pz) {
  int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= nghost) return;
  int o, code;
  if (FULL) {
    long long key = gkey[gi];
    o = (int)(key >> 5); code = (int)(key & 31);
    gowner[gi] = o; gcode[gi] = code;
  } else { o = gowner[gi]; code = gcode[gi]; }
  int sx = code % 3 - 1, sy = (code / 3) % 3 - 1, sz = code / 9 - 1;
  double4 r = pos[o];
  if (sx) r.x = r.x + sx * px;
  if (sy) r.y = r.y + sy * py;
  if (sz) r.z = r.z + sz * pz;
  int k = nlocal + gi;
  pos[k] = r;
  ts[k] = ts[o];
  ucgp[k] = ucgp[o];
  if (FULL) { tag[k] = tag[o]; mol[k] = mol[o]; }
}

// One warp per owned site walks the 27-cell stencil.  Owned atoms and ghosts are both
// stored in cell order, so each (dy,dz) row of three x-adjacent cells is one contiguous
// index range of owned atoms plus one of ghosts; lanes test 32 candidates at a time and
// ballot-compact the hits, which keeps every row in a deterministic order.
// Rows are partitioned: neighbors inside the pair cutoff at build time come first (in
// cell order), the skin shell (cut < r <= cut+skin) after them.  The pair kernels test
// rsq < cutsq for every entry anyway; the partition only makes the outcome of that test
// coherent across a warp (lane efficiency 77 % -> ~93 %, profiles/).
__global__ void __launch_bounds__(256)
k_build_rows(const double4 *__restrict__ pos, const int *__restrict__ ts, int nlocal, Grid g,
             const int *__restrict__ ostart, const int *__restrict__ gstart, const PairInfo *__restrict__ pinfo,
             int na, int *__restrict__ neigh, int stride, int *__restrict__ numneigh, int *__restrict__ flags) {
  extern __shared__ int s_outer[];  // [warps per block][stride]
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= nlocal) return;
  int *outer = s_outer + (threadIdx.x >> 5) * stride;
  int i = warp;
  double4 ri = pos[i];
  int ti = ts[i] & 0xffff;
  int ix = min(max(cell_coord(ri.x, g.lo[0], g.inv[0], g.nc[0]), 1), g.ninner[0]);
  int iy = min(max(cell_coord(ri.y, g.lo[1], g.inv[1], g.nc[1]), 1), g.ninner[1]);
  int iz = min(max(cell_coord(ri.z, g.lo[2], g.inv[2], g.nc[2]), 1), g.ninner[2]);
  int *row = neigh + (size_t)i * stride;
  int cnt_in = 0, cnt_out = 0;
  const PairInfo *prow = pinfo + ti * na;
  const unsigned lt = (1u << lane) - 1;
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++) {
      int c0 = ((iz + dz) * g.nc[1] + (iy + dy)) * g.nc[0] + (ix - 1);
      for (int pass = 0; pass < 2; pass++) {
        int b, e, off;
        if (pass == 0) { b = ostart[c0]; e = ostart[c0 + 3]; off = 0; }
        else { b = gstart[c0]; e = gstart[c0 + 3]; off = nlocal; }
        for (int base = b; base < e; base += 32) {
          int jj = base + lane;
          bool hit = false, inner = false;
          int j = -1;
          if (jj < e) {
            j = jj + off;
            double4 rj = pos[j];
            int tj = ts[j] & 0xffff;
            double dx = ri.x - rj.x, dyv = ri.y - rj.y, dzv = ri.z - rj.z;
            double rsq = rsq_exact(dx, dyv, dzv);
            const PairInfo pi = prow[tj];
            hit = (j != i) && (rsq <= pi.cutneighsq);
            inner = hit && (rsq < pi.cutsq);
          }
          unsigned m_in = __ballot_sync(0xffffffffu, inner);
          unsigned m_out = __ballot_sync(0xffffffffu, hit && !inner);
          if (inner) {
            int p = cnt_in + __popc(m_in & lt);
            if (p < stride) row[p] = j;
          } else if (hit) {
            int p = cnt_out + __popc(m_out & lt);
            if (p < stride) outer[p] = j;
          }
          cnt_in += __popc(m_in);
          cnt_out += __popc(m_out);
        }
      }
    }
  __syncwarp();
  const int total = cnt_in + cnt_out;
  if (total <= stride)
    for (int k = lane; k < cnt_out; k += 32) row[cnt_in + k] = outer[k];
  if (lane == 0) {
    numneigh[i] = min(total, stride);
    if (total > stride) atomicMax(&flags[1], total);
  }
}

// [stock] Neighbor::check_distance: any owned atom moved > skin/2 since the last build
__global__ void k_check_distance(const double4 *__restrict__ pos, const double4 *__restrict__ xhold, int n,
                                 double triggersq, int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 a = pos[i], b = xhold[i];
  double rsq = rsq_exact(a.x - b.x, a.y - b.y, a.z - b.z);
  if (rsq > triggersq) flags[0] = 1;
}

}  // namespace

// ------------------------------------------------------------------ host side
extern "C" int ucgb200_neigh_configure(ucgb200_ctx *c, double skin, double cut_override) {
  if (!c || skin < 0) return -1;
  c->skin = skin;
  c->cut_override = cut_override;
  c->maps_dirty = true;
  c->list_valid = false;
  return 0;
}

static int setup_grid(ucgb200_ctx *c) {
  double cutneigh = c->max_cut + c->skin;
  c->cutneighmax = cutneigh;
  Grid &g = c->grid;
  long long ncells = 1;
  for (int d = 0; d < 3; d++) {
    double len = c->subhi[d] - c->sublo[d];
    if (c->periodic[d] && (c->boxhi[d] - c->boxlo[d]) < cutneigh) {
      c->err = "periodic box shorter than the neighbor cutoff (cut+skin)";
      return UCGB200_ERR_BOX_TOO_SMALL;
    }
    int n = (int)std::floor(len / cutneigh);
    if (n < 1) n = 1;
    if (n > 1 && len / n < cutneigh * (1.0 + 1e-9)) n--;
    g.ninner[d] = n;
    g.nc[d] = n + 2;
    g.lo[d] = c->sublo[d];
    g.inv[d] = n / len;
    ncells *= g.nc[d];
  }
  if (ncells > (1ll << 30)) return fail(c, "cell grid too large");
  g.ncells = (int)ncells;
  return 0;
}

static BoxDev make_box(const ucgb200_ctx *c) {
  BoxDev b;
  for (int d = 0; d < 3; d++) {
    b.lo[d] = c->boxlo[d]; b.hi[d] = c->boxhi[d]; b.prd[d] = c->prd[d];
    b.sublo[d] = c->sublo[d]; b.subhi[d] = c->subhi[d]; b.periodic[d] = c->periodic[d];
  }
  b.cutghost = c->cutneighmax;
  return b;
}

int ucg_ensure_atom_capacity(ucgb200_ctx *c, size_t nall, size_t nloc);

static int read_flags(ucgb200_ctx *c) {
  UCG_CHECK(c, cudaMemcpyAsync(c->h_flags, c->d_flags.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

static int build_rows(ucgb200_ctx *c) {
  int nlocal = c->nlocal;
  int na = c->n_actual + 1;
  while (true) {
    UCG_CHECK(c, c->neigh.ensure((size_t)nlocal * c->neigh_stride));
    UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p + 1, 0, sizeof(int), c->stream));
    long long nthreads = (long long)nlocal * 32;
    k_build_rows<<<nblocks(nthreads, 256), 256, 8 * c->neigh_stride * sizeof(int), c->stream>>>(
        c->pos.p, c->ts.p, nlocal, c->grid, c->cell_start.p, c->gcell_start.p, c->d_pairinfo.p, na,
        c->neigh.p, c->neigh_stride, c->numneigh.p, c->d_flags.p);
    UCG_LAUNCHED(c);
    int rc = read_flags(c);
    if (rc) return rc;
    if (c->h_flags[1] <= c->neigh_stride) break;
    // UCGB200_ERR_NEIGH_OVERFLOW handled internally: grow the row capacity and redo
    c->neigh_stride = ((c->h_flags[1] + c->h_flags[1] / 8 + 7) / 8) * 8;
  }
  return 0;
}

extern "C" int ucgb200_neigh_build(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if (c->nlocal == 0) { c->list_valid = true; return 0; }
  if ((rc = setup_grid(c))) return rc;
  cudaStream_t st = c->stream;
  int nlocal = c->nlocal;
  int ncells = c->grid.ncells;
  BoxDev box = make_box(c);
  cudaEvent_t t0 = c->ev_a, t1 = c->ev_b;
  if (c->timers_on) cudaEventRecord(t0, st);
  long long l0 = c->launches;

  UCG_CHECK(c, c->cell_count.ensure(ncells + 4));
  UCG_CHECK(c, c->cell_start.ensure(ncells + 4));
  UCG_CHECK(c, c->cell_cursor.ensure(ncells + 4));
  UCG_CHECK(c, c->gcell_count.ensure(ncells + 4));
  UCG_CHECK(c, c->gcell_start.ensure(ncells + 4));
  UCG_CHECK(c, c->order.ensure(nlocal));
  UCG_CHECK(c, c->cell_of.ensure(nlocal));
  UCG_CHECK(c, c->numneigh.ensure(nlocal));
  UCG_CHECK(c, cudaMemsetAsync(c->cell_count.p, 0, (ncells + 4) * sizeof(int), st));
  UCG_CHECK(c, cudaMemsetAsync(c->cell_cursor.p, 0, (ncells + 4) * sizeof(int), st));
  UCG_CHECK(c, cudaMemsetAsync(c->gcell_count.p, 0, (ncells + 4) * sizeof(int), st));
  UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p, 0, 8 * sizeof(int), st));

  // 1. pbc wrap + owned-cell histogram
  k_wrap_count<<<nblocks(nlocal, 256), 256, 0, st>>>(c->pos.p, nlocal, box, c->grid, c->cell_of.p,
                                                     c->cell_count.p, c->d_flags.p);
  UCG_LAUNCHED(c);
  // 2. cell offsets (ncells+3 entries so that start[c0+3] is always readable)
  if ((rc = exclusive_scan(c, c->cell_count.p, c->cell_start.p, ncells + 4, nullptr))) return rc;
  // 3. counting sort into cells, deterministic inside each cell
  k_fill_order<<<nblocks(nlocal, 256), 256, 0, st>>>(c->cell_of.p, nlocal, c->cell_start.p, c->cell_cursor.p, c->order.p);
  UCG_LAUNCHED(c);
  k_sort_cells<int><<<nblocks(ncells, 128), 128, 0, st>>>(c->order.p, c->cell_start.p, ncells);
  UCG_LAUNCHED(c);
  // 4. move every per-site array into cell order
  UCG_CHECK(c, c->pos_alt.ensure_exact(c->pos.cap)); UCG_CHECK(c, c->ts_alt.ensure_exact(c->ts.cap));
  UCG_CHECK(c, c->ucgp_alt.ensure_exact(c->ucgp.cap)); UCG_CHECK(c, c->tag_alt.ensure_exact(c->tag.cap));
  UCG_CHECK(c, c->mol_alt.ensure_exact(c->mol.cap)); UCG_CHECK(c, c->vel_alt.ensure_exact(c->vel.cap));
  UCG_CHECK(c, c->frc_alt.ensure_exact(c->frc.cap)); UCG_CHECK(c, c->scores_alt.ensure_exact(c->scores.cap));
  UCG_CHECK(c, c->ucgml_alt.ensure_exact(c->ucgml.cap)); UCG_CHECK(c, c->mask_alt.ensure_exact(c->mask.cap));
  UCG_CHECK(c, c->orig_alt.ensure_exact(c->orig.cap));
  PermArgs pa{c->pos.p, c->vel.p, c->frc.p, c->scores.p, c->ucgp.p, c->ucgml.p, c->ts.p, c->mask.p,
              c->tag.p, c->mol.p, c->orig.p, c->pos_alt.p, c->vel_alt.p, c->frc_alt.p, c->scores_alt.p,
              c->ucgp_alt.p, c->ucgml_alt.p, c->ts_alt.p, c->mask_alt.p, c->tag_alt.p, c->mol_alt.p,
              c->orig_alt.p};
  k_permute<<<nblocks(nlocal, 256), 256, 0, st>>>(pa, c->order.p, nlocal);
  UCG_LAUNCHED(c);
  std::swap(c->pos, c->pos_alt); std::swap(c->vel, c->vel_alt); std::swap(c->frc, c->frc_alt);
  std::swap(c->scores, c->scores_alt); std::swap(c->ucgp, c->ucgp_alt); std::swap(c->ucgml, c->ucgml_alt);
  std::swap(c->ts, c->ts_alt); std::swap(c->mask, c->mask_alt); std::swap(c->tag, c->tag_alt);
  std::swap(c->mol, c->mol_alt); std::swap(c->orig, c->orig_alt);

  // 5. periodic ghost images, stored in cell order after the owned atoms
  k_ghost_images<false><<<nblocks(nlocal, 256), 256, 0, st>>>(c->pos.p, nlocal, box, c->grid, c->gcell_count.p,
                                                              nullptr, nullptr, nullptr);
  UCG_LAUNCHED(c);
  if ((rc = exclusive_scan(c, c->gcell_count.p, c->gcell_start.p, ncells + 4, c->d_flags.p + 2))) return rc;
  if ((rc = read_flags(c))) return rc;
  if (c->h_flags[3]) { c->err = "atoms lost: position outside the periodic box by more than one period"; return UCGB200_ERR_LOST_ATOMS; }
  int nghost = c->h_flags[2];
  c->nghost = nghost;
  if ((rc = ucg_ensure_atom_capacity(c, (size_t)nlocal + nghost, nlocal))) return rc;
  if (nghost > 0) {
    UCG_CHECK(c, c->ghost_key.ensure(nghost));
    UCG_CHECK(c, c->ghost_owner.ensure(nghost));
    UCG_CHECK(c, c->ghost_code.ensure(nghost));
    UCG_CHECK(c, cudaMemsetAsync(c->cell_cursor.p, 0, (ncells + 4) * sizeof(int), st));
    k_ghost_images<true><<<nblocks(nlocal, 256), 256, 0, st>>>(c->pos.p, nlocal, box, c->grid, nullptr,
                                                               c->gcell_start.p, c->cell_cursor.p, c->ghost_key.p);
    UCG_LAUNCHED(c);
    k_sort_cells<long long><<<nblocks(ncells, 128), 128, 0, st>>>(c->ghost_key.p, c->gcell_start.p, ncells);
    UCG_LAUNCHED(c);
    k_ghost_fill<true><<<nblocks(nghost, 256), 256, 0, st>>>(c->pos.p, c->ts.p, c->ucgp.p, c->tag.p, c->mol.p,
                                                             nlocal, nghost, c->ghost_key.p, c->ghost_owner.p,
                                                             c->ghost_code.p, c->prd[0], c->prd[1], c->prd[2]);
    UCG_LAUNCHED(c);
  }
  // 6. neighbor rows
  if (c->neigh_stride == 0) {
    double vol = 1.0;
    for (int d = 0; d < 3; d++) vol *= (c->subhi[d] - c->sublo[d]);
    double est = 4.18879 * c->cutneighmax * c->cutneighmax * c->cutneighmax * nlocal / vol;
    int s = (int)(est * 1.35) + 24;
    c->neigh_stride = ((s + 7) / 8) * 8;
  }
  if ((rc = build_rows(c))) return rc;
  // 7. remember positions for the skin check
  UCG_CHECK(c, cudaMemcpyAsync(c->xhold.p, c->pos.p, (size_t)nlocal * sizeof(double4), cudaMemcpyDeviceToDevice, st));
  c->list_valid = true;
  c->nbuilds++;
  if (c->timers_on) {
    cudaEventRecord(t1, st);
    cudaEventSynchronize(t1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    c->t_ms[1] += ms;
    c->t_launch[1] += c->launches - l0;
  }
  return 0;
}

extern "C" int ucgb200_neigh_decide(ucgb200_ctx *c, int *rebuild) {
  if (!c || !rebuild) return -1;
  cudaSetDevice(c->device);
  if (!c->list_valid) { *rebuild = 1; return 0; }
  if (c->nlocal == 0) { *rebuild = 0; return 0; }
  double triggersq = 0.25 * c->skin * c->skin;
  UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p, 0, sizeof(int), c->stream));
  k_check_distance<<<nblocks(c->nlocal, 256), 256, 0, c->stream>>>(c->pos.p, c->xhold.p, c->nlocal, triggersq, c->d_flags.p);
  UCG_LAUNCHED(c);
  UCG_CHECK(c, cudaMemcpyAsync(c->h_flags, c->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  *rebuild = c->h_flags[0] ? 1 : 0;
  return 0;
}

extern "C" int ucgb200_ghosts_forward(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  if (c->nghost == 0) return 0;
  k_ghost_fill<false><<<nblocks(c->nghost, 256), 256, 0, c->stream>>>(
      c->pos.p, c->ts.p, c->ucgp.p, c->tag.p, c->mol.p, c->nlocal, c->nghost, nullptr, c->ghost_owner.p,
      c->ghost_code.p, c->prd[0], c->prd[1], c->prd[2]);
  UCG_LAUNCHED(c);
  return 0;
}

extern "C" int ucgb200_neigh_stats(ucgb200_ctx *c, long long *total_pairs, int *max_row, int *nbuilds) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  if (nbuilds) *nbuilds = c->nbuilds;
  if (!c->list_valid || c->nlocal == 0) { if (total_pairs) *total_pairs = 0; if (max_row) *max_row = 0; return 0; }
  std::vector<int> nn(c->nlocal);
  UCG_CHECK(c, cudaMemcpyAsync(nn.data(), c->numneigh.p, c->nlocal * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  long long t = 0; int m = 0;
  for (int v : nn) { t += v; m = std::max(m, v); }
  if (total_pairs) *total_pairs = t;
  if (max_row) *max_row = m;
  return 0;
}

extern "C" int ucgb200_neigh_download(ucgb200_ctx *c, int *nlocal_out, long long *total, int *tag_i, int *numneigh,
                                      long long *offsets, int *neigh_tags, int *neigh_shift) {
  if (!c || !total) return -1;
  cudaSetDevice(c->device);
  if (!c->list_valid) return fail(c, "neigh_download: no valid list");
  int nlocal = c->nlocal, nall = c->nlocal + c->nghost;
  std::vector<int> nn(nlocal);
  if (nlocal) UCG_CHECK(c, cudaMemcpy(nn.data(), c->numneigh.p, nlocal * sizeof(int), cudaMemcpyDeviceToHost));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  if (nlocal) UCG_CHECK(c, cudaMemcpy(nn.data(), c->numneigh.p, nlocal * sizeof(int), cudaMemcpyDeviceToHost));
  long long t = 0;
  for (int v : nn) t += v;
  if (nlocal_out) *nlocal_out = nlocal;
  if (!neigh_tags) { *total = t; return 0; }
  if (*total < t) return fail(c, "neigh_download: capacity too small");
  *total = t;
  std::vector<int> tags(nall), rows((size_t)nlocal * c->neigh_stride), gcode(std::max(c->nghost, 1));
  UCG_CHECK(c, cudaMemcpy(tags.data(), c->tag.p, nall * sizeof(int), cudaMemcpyDeviceToHost));
  UCG_CHECK(c, cudaMemcpy(rows.data(), c->neigh.p, rows.size() * sizeof(int), cudaMemcpyDeviceToHost));
  if (c->nghost) UCG_CHECK(c, cudaMemcpy(gcode.data(), c->ghost_code.p, c->nghost * sizeof(int), cudaMemcpyDeviceToHost));
  long long off = 0;
  for (int i = 0; i < nlocal; i++) {
    if (tag_i) tag_i[i] = tags[i];
    if (numneigh) numneigh[i] = nn[i];
    if (offsets) offsets[i] = off;
    for (int k = 0; k < nn[i]; k++) {
      int j = rows[(size_t)i * c->neigh_stride + k] & UCG_NEIGHMASK;
      neigh_tags[off + k] = tags[j];
      if (neigh_shift) neigh_shift[off + k] = j < nlocal ? 13 : gcode[j - nlocal];
    }
    off += nn[i];
  }
  if (offsets) offsets[nlocal] = off;
  return 0;
}

extern "C" int ucgb200_neigh_flag_ptr(ucgb200_ctx *c, void **d_flag) {
  if (!c || !d_flag) return -1;
  *d_flag = c->d_flags.p;
  return 0;
}
