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

// ---------------------------------------------------------------- ghosts / halo
// [stock] CommBrick::borders generalised to a brick decomposition over GPUs.  An owned
// atom within cutghost of a low ("L") or high ("H") face of its sub-domain has an image in
// the neighbouring brick of that direction; all combinations over the three dimensions
// exist (the staged x,y,z exchange of LAMMPS yields exactly these).  code = cx+3cy+9cz with
// c in {0 none, 1 L, 2 H}.  If the neighbouring brick is this rank itself (one brick in that
// dimension) the image is a local periodic ghost; otherwise it is a border record for the
// rank that owns that brick.  The SENDER applies the periodic shift.
struct ImageMap {
  int dest[27];        // destination rank of every code (own rank -> local image)
  double shift[27][3]; // coordinate shift applied by the sender
  int self;
  int nranks;
  int allow_lo[3], allow_hi[3];  // image exists in that direction (periodic or interior face)
};

__device__ __forceinline__ int face_options(double x, double lo, double hi, double cut, int allow_lo, int allow_hi, int opt[3]) {
  int n = 0;
  opt[n++] = 0;
  if (allow_lo && x <= lo + cut) opt[n++] = 1;
  if (allow_hi && x >= hi - cut) opt[n++] = 2;
  return n;
}

// counters/cursors layout: [0..nranks) remote destinations, [nranks] local images
template <bool FILL>
__global__ void k_images(const double4 *__restrict__ pos, int n, BoxDev box, ImageMap im, int *__restrict__ counters,
                         const int *__restrict__ offsets, int *__restrict__ list_owner, int *__restrict__ list_code) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  // 7 images for a site in a corner of the sub-domain, up to 26 when the box is barely wider than the ghost cutoff
  // (a site then lies within reach of both faces of a dimension): 5 bits per code, 26 codes in two 64-bit words + one
  unsigned long long packed[3] = {0ull, 0ull, 0ull};
  int ncode = 0;
  if (i < n) {
    const double4 r = pos[i];
    int ox[3], oy[3], oz[3];
    const int nx = face_options(r.x, box.sublo[0], box.subhi[0], box.cutghost, im.allow_lo[0], im.allow_hi[0], ox);
    const int ny = face_options(r.y, box.sublo[1], box.subhi[1], box.cutghost, im.allow_lo[1], im.allow_hi[1], oy);
    const int nz = face_options(r.z, box.sublo[2], box.subhi[2], box.cutghost, im.allow_lo[2], im.allow_hi[2], oz);
    for (int c = 0; c < nz; c++)
      for (int b = 0; b < ny; b++)
        for (int a = 0; a < nx; a++) {
          const int code = ox[a] + 3 * oy[b] + 9 * oz[c];
          if (code) { packed[ncode / 12] |= (unsigned long long)code << (5 * (ncode % 12)); ncode++; }
        }
  }
  // one atomicAdd per (warp, destination) instead of one per image: the counters are a handful of addresses that
  // ~18 % of all sites hit (the unaggregated kernel spent 0.15 ms per pass at 1 M sites in L2 atomic serialisation).
  // Slots inside a destination's list are handed out in lane order; the lists are sorted by (tag, image) later anyway.
  const int most = __reduce_max_sync(0xffffffffu, ncode);
  for (int t = 0; t < most; t++) {
    const bool has = t < ncode;
    const int code = has ? (int)((packed[t / 12] >> (5 * (t % 12))) & 31) : 0;
    const int d = has ? im.dest[code] : -1;
    const int idx = has ? ((d == im.self) ? im.nranks : d) : -1;
    const unsigned active = __ballot_sync(0xffffffffu, has);
    if (!has) continue;
    const unsigned peers = __match_any_sync(active, idx);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&counters[idx], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    if (FILL) {
      const int slot = offsets[idx] + base + __popc(peers & ((1u << lane) - 1));
      list_owner[slot] = i;
      list_code[slot] = code;
    }
  }
}

struct BorderRec {  // 64 B: fields_border of AtomVecUCG (atom_vec_ucg.cpp:66-67) + tag/type
  double4 pos;      // x,y,z (shifted), ucgl
  int ts, tag, mol, code;
  double ucgp;
  int mask, pad;    // group bits travel with the border record (fix cluster_switch masks by groupbit, fix_cluster_switch.cpp:102,138,569,578)
};
struct ForwardRec {  // 48 B: x + fields_comm {ucgstate, ucgl, ucgp} (atom_vec_ucg.cpp:71)
  double4 pos;
  int ts, pad;
  double ucgp;
};
struct MigrateRec {  // 96 B: the UCG part of fields_exchange (atom_vec_ucg.cpp:76-82)
  double4 pos, vel;
  int ts, mask, tag, mol;
  double ucgp, ucgml;
};

__global__ void k_pack_border(const double4 *__restrict__ pos, const int *__restrict__ ts, const int *__restrict__ tag,
                              const int *__restrict__ mol, const double *__restrict__ ucgp, const int *__restrict__ mask,
                              const int *__restrict__ owner, const int *__restrict__ code, int n, ImageMap im,
                              BorderRec *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int o = owner[k], cd = code[k];
  double4 r = pos[o];
  if (im.shift[cd][0] != 0.0) r.x = r.x + im.shift[cd][0];
  if (im.shift[cd][1] != 0.0) r.y = r.y + im.shift[cd][1];
  if (im.shift[cd][2] != 0.0) r.z = r.z + im.shift[cd][2];
  BorderRec b;
  b.pos = r; b.ts = ts[o]; b.tag = tag[o]; b.mol = mol[o]; b.code = cd; b.ucgp = ucgp[o]; b.mask = mask[o]; b.pad = 0;
  out[k] = b;
}
__global__ void k_pack_forward(const double4 *__restrict__ pos, const int *__restrict__ ts, const double *__restrict__ ucgp,
                               const int *__restrict__ owner, const int *__restrict__ code, int n, ImageMap im,
                               ForwardRec *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int o = owner[k], cd = code[k];
  double4 r = pos[o];
  if (im.shift[cd][0] != 0.0) r.x = r.x + im.shift[cd][0];
  if (im.shift[cd][1] != 0.0) r.y = r.y + im.shift[cd][1];
  if (im.shift[cd][2] != 0.0) r.z = r.z + im.shift[cd][2];
  ForwardRec f;
  f.pos = r; f.ts = ts[o]; f.pad = 0; f.ucgp = ucgp[o];
  out[k] = f;
}
__global__ void k_unpack_forward(double4 *__restrict__ pos, int *__restrict__ ts, double *__restrict__ ucgp, int nlocal,
                                 const ForwardRec *__restrict__ in, const int *__restrict__ slot_of_src, int nlimg, int n) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int s = nlocal + slot_of_src[nlimg + k];
  ForwardRec f = in[k];
  pos[s] = f.pos; ts[s] = f.ts; ucgp[s] = f.ucgp;
}

// Forward halo as direct stores into the peers' receive regions (NVLink / NVSwitch peer mapping): one kernel packs every
// record of the send list and stores it where the destination brick will read it; no send buffer, no NCCL kernel, no
// receive-side copy.  The last CTA to finish publishes this brick's rebuild flag and displacement bound and then the
// sequence number in every peer's control block, behind a system-scope fence.  comm->forward_comm() of the reference:
// the payload is AtomVecUCG's fields_comm (UCG/atom_vec_ucg.cpp:71).
__global__ void k_push_forward(const double4 *__restrict__ pos, const int *__restrict__ ts, const double *__restrict__ ucgp,
                               const int *__restrict__ owner, const int *__restrict__ code, int n, ImageMap im,
                               UcgPushTargets t) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    int r = 0;
#pragma unroll 1
    while (r + 1 < t.nranks && k >= t.send_off[r + 1]) r++;
    const int o = owner[k], cd = code[k];
    double4 p = pos[o];
    if (im.shift[cd][0] != 0.0) p.x = p.x + im.shift[cd][0];
    if (im.shift[cd][1] != 0.0) p.y = p.y + im.shift[cd][1];
    if (im.shift[cd][2] != 0.0) p.z = p.z + im.shift[cd][2];
    ForwardRec f;
    f.pos = p; f.ts = ts[o]; f.pad = 0; f.ucgp = ucgp[o];
    reinterpret_cast<ForwardRec *>(t.rec[r])[k - t.send_off[r]] = f;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(t.done, 1u);
    if (prev == gridDim.x - 1) {
      *t.done = 0u;
      const int flag = *t.local_flag;
      const unsigned long long md = *t.local_maxdisp;
      for (int r = 0; r < t.nranks; r++)
        if (r != t.self) {
          volatile UcgP2PCtl *w = t.ctl[r];
          w->maxdisp = md;
          w->flag = flag;
        }
      __threadfence_system();
      for (int r = 0; r < t.nranks; r++)
        if (r != t.self) {
          volatile UcgP2PCtl *w = t.ctl[r];
          w->seq = t.seq;
        }
      __threadfence_system();
    }
  }
}

// One thread per peer waits until that peer's push of this sequence number has landed (bounded: ~2 s of
// %globaltimer, then the error word is set and the run stops), then the rebuild flags and displacement bounds of all
// bricks are folded into this brick's d_flags[0] / d_maxdisp: Neighbor::decide's MPI_Allreduce without a collective.
__global__ void k_wait_reduce(const UcgP2PCtl *ctl, int nranks, int self, int seq, int *flag, unsigned long long *maxdisp,
                              ErrWord *err) {
  const int r = threadIdx.x;
  int f = 0;
  unsigned long long md = 0ull;
  if (r < nranks && r != self) {
    const volatile UcgP2PCtl *w = ctl + r;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    bool ok = true;
    while (w->seq != seq) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) { ok = false; break; }
      __nanosleep(64);
    }
    __threadfence_system();
    if (ok) { f = w->flag; md = w->maxdisp; }
    else if (atomicCAS(&err->code, 0, UCGB200_ERR_PEER_TIMEOUT) == 0) { err->tag_i = r; err->tag_j = seq; err->rsq = 0.0; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    f = max(f, __shfl_xor_sync(0xffffffffu, f, o));
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, md, o);
    md = other > md ? other : md;
  }
  if (threadIdx.x == 0) {
    if (f > *flag) *flag = f;
    if (md > *maxdisp) *maxdisp = md;
  }
}

// ghost sources = local images [0,nlimg) followed by received border records.
__device__ __forceinline__ double4 source_pos(int k, int nlimg, const double4 *pos, const int *lo, const int *lc,
                                              const ImageMap &im, const BorderRec *recv) {
  if (k < nlimg) {
    double4 r = pos[lo[k]];
    int cd = lc[k];
    if (im.shift[cd][0] != 0.0) r.x = r.x + im.shift[cd][0];
    if (im.shift[cd][1] != 0.0) r.y = r.y + im.shift[cd][1];
    if (im.shift[cd][2] != 0.0) r.z = r.z + im.shift[cd][2];
    return r;
  }
  return recv[k - nlimg].pos;
}
template <bool FILL>
__global__ void k_source_cells(const double4 *__restrict__ pos, const int *__restrict__ tag, int nsrc, int nlimg,
                               const int *__restrict__ lo, const int *__restrict__ lc, ImageMap im,
                               const BorderRec *__restrict__ recv, Grid g, int *__restrict__ gcell_count,
                               const int *__restrict__ gcell_start, int *__restrict__ cursor,
                               long long *__restrict__ gkey, int *__restrict__ gsrc) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsrc) return;
  double4 r = source_pos(k, nlimg, pos, lo, lc, im, recv);
  int cid = cell_index(g, r.x, r.y, r.z, false);
  if (!FILL) { atomicAdd(&gcell_count[cid], 1); return; }
  int slot = gcell_start[cid] + atomicAdd(&cursor[cid], 1);
  int t, cd;
  if (k < nlimg) { t = tag[lo[k]]; cd = lc[k]; }
  else { t = recv[k - nlimg].tag; cd = recv[k - nlimg].code; }
  gkey[slot] = ((long long)t << 6) | (long long)(cd + (k < nlimg ? 0 : 32));
  gsrc[slot] = k;
}
// deterministic order inside every ghost cell: ascending (tag, image code)
__global__ void k_sort_cells_kv(long long *__restrict__ keys, int *__restrict__ vals, const int *__restrict__ start, int ncells) {
  int cid = blockIdx.x * blockDim.x + threadIdx.x;
  if (cid >= ncells) return;
  int b = start[cid], e = start[cid + 1];
  for (int a = b + 1; a < e; a++) {
    long long key = keys[a];
    int v = vals[a];
    int k = a - 1;
    while (k >= b && keys[k] > key) { keys[k + 1] = keys[k]; vals[k + 1] = vals[k]; k--; }
    keys[k + 1] = key; vals[k + 1] = v;
  }
}
// fill ghost slots; FULL also writes tag/molecule and the slot_of_src map (rebuild), otherwise
// only the per-step payload of the LOCAL images is refreshed (forward_comm to self)
template <bool FULL>
__global__ void k_ghost_fill(double4 *__restrict__ pos, int *__restrict__ ts, double *__restrict__ ucgp,
                             int *__restrict__ tag, int *__restrict__ mol, int nlocal, int n, int nlimg,
                             const int *__restrict__ gsrc, int *__restrict__ slot_of_src, const int *__restrict__ lo,
                             const int *__restrict__ lc, ImageMap im, const BorderRec *__restrict__ recv,
                             int *__restrict__ gcode, const int *__restrict__ mask = nullptr, int *__restrict__ gmask = nullptr) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  int k, slot;
  if (FULL) { slot = idx; k = gsrc[slot]; slot_of_src[k] = slot; }
  else { k = idx; slot = slot_of_src[k]; }   // idx runs over local images only
  int s = nlocal + slot;
  if (k < nlimg) {
    int o = lo[k];
    pos[s] = source_pos(k, nlimg, pos, lo, lc, im, recv);
    ts[s] = ts[o];
    ucgp[s] = ucgp[o];
    if (FULL) { tag[s] = tag[o]; mol[s] = mol[o]; gcode[slot] = lc[k]; if (gmask) gmask[slot] = mask[o]; }
  } else if (FULL) {
    BorderRec b = recv[k - nlimg];
    pos[s] = b.pos; ts[s] = b.ts; ucgp[s] = b.ucgp; tag[s] = b.tag; mol[s] = b.mol; gcode[slot] = 32 + b.code;
    if (gmask) gmask[slot] = b.mask;
  }
}

// migration: atoms whose (wrapped) position left this brick
__global__ void k_migrate_classify(const double4 *__restrict__ pos, int n, BoxDev box, int gx, int gy, int gz, int self,
                                   int *__restrict__ dest, int *__restrict__ stay, int *__restrict__ counters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 r = pos[i];
  double x[3] = {r.x, r.y, r.z};
  int g[3] = {gx, gy, gz}, pc[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double w = box.prd[d] / g[d];
    int c = (int)floor((x[d] - box.lo[d]) / w);
    c = min(max(c, 0), g[d] - 1);
    // same plane arithmetic as ucgb200_halo_configure: split(c) = lo + c*w
    while (c > 0 && x[d] < box.lo[d] + c * w) c--;
    while (c < g[d] - 1 && x[d] >= box.lo[d] + (c + 1) * w) c++;
    pc[d] = c;
  }
  int r_dest = (pc[2] * gy + pc[1]) * gx + pc[0];
  dest[i] = r_dest;
  stay[i] = r_dest == self;
  if (r_dest != self) atomicAdd(&counters[r_dest], 1);
}
__global__ void k_migrate_pack(const double4 *__restrict__ pos, const double4 *__restrict__ vel, const int *__restrict__ ts,
                               const int *__restrict__ mask, const int *__restrict__ tag, const int *__restrict__ mol,
                               const double *__restrict__ ucgp, const double *__restrict__ ucgml, int n, int self,
                               const int *__restrict__ dest, const int *__restrict__ offsets, int *__restrict__ cursor,
                               MigrateRec *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = dest[i];
  if (d == self) return;
  int slot = offsets[d] + atomicAdd(&cursor[d], 1);
  MigrateRec m;
  m.pos = pos[i]; m.vel = vel[i]; m.ts = ts[i]; m.mask = mask[i]; m.tag = tag[i]; m.mol = mol[i];
  m.ucgp = ucgp[i]; m.ucgml = ucgml[i];
  out[slot] = m;
}
__global__ void k_stay_order(const int *__restrict__ stay, const int *__restrict__ stay_scan, int n, int *__restrict__ order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && stay[i]) order[stay_scan[i]] = i;
}
__global__ void k_migrate_unpack(double4 *__restrict__ pos, double4 *__restrict__ vel, double4 *__restrict__ frc,
                                 double2 *__restrict__ scores, int *__restrict__ ts, int *__restrict__ mask,
                                 int *__restrict__ tag, int *__restrict__ mol, double *__restrict__ ucgp,
                                 double *__restrict__ ucgml, int *__restrict__ orig, int base, int n,
                                 const MigrateRec *__restrict__ in) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  MigrateRec m = in[k];
  int s = base + k;
  pos[s] = m.pos; vel[s] = m.vel; frc[s] = make_double4(0, 0, 0, 0); scores[s] = make_double2(0, 0);
  ts[s] = m.ts; mask[s] = m.mask; tag[s] = m.tag; mol[s] = m.mol; ucgp[s] = m.ucgp; ucgml[s] = m.ucgml;
  orig[s] = s;
}
__global__ void k_iota2(int *p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
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
             int na, int *__restrict__ neigh, int stride, int *__restrict__ numneigh, int *__restrict__ flags,
             uint4 *__restrict__ levcnt) {
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
            if (p < stride) row[rowslot(p)] = j;
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
    for (int k = lane; k < cnt_out; k += 32) row[rowslot(cnt_in + k)] = outer[k];
  if (lane == 0) {
    numneigh[i] = min(total, stride);
    if (total > stride) atomicMax(&flags[1], total);
    const unsigned t2 = (unsigned)min(total, stride) * 0x10001u;   // no distance order here: every level visits everything
    levcnt[i] = make_uint4(t2, t2, t2, t2);
  }
}

// Cell-tiled variant of k_build_rows (same rows, same order): one CTA per owned cell.  The
// candidates of the 27-cell stencil (18 contiguous ranges: 9 x-rows of 3 cells, owned then ghost)
// are staged ONCE in shared memory — {x,y,z,.}, type and global index — and every site of the cell
// scans them from there, so each candidate record is read from L2/HBM once per cell instead of once
// per site (~18x less global traffic).  Cells whose stencil exceeds the staging capacity are
// processed in chunks.  Row order (dz, dy, owned|ghost, index) and the inner/skin partition are
// those of k_build_rows, so the lists are identical.
constexpr int TILE_CAP = 768;     // candidates per chunk: 768 * 40 B = 30 KB
constexpr int TILE_CAP_F32 = 1024;   // k_build_rows_tiled_f32: 1024 * 16 B = 16 KB
constexpr int TILE_BS = 128;
__global__ void __launch_bounds__(TILE_BS)
k_build_rows_tiled(const double4 *__restrict__ pos, const int *__restrict__ ts, int nlocal, Grid g,
                   const int *__restrict__ ostart, const int *__restrict__ gstart, const PairInfo *__restrict__ pinfo,
                   int na, int *__restrict__ neigh, int stride, int *__restrict__ numneigh, int *__restrict__ flags,
                   int cap, uint4 *__restrict__ levcnt, double skin, int defer_keys) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  // structure of arrays: a 16-byte {x,y} and an 8-byte z per candidate keep the warp-wide reads free of
  // bank conflicts (a 32-byte record read as two 16-byte halves is a 2-way conflict)
  double2 *s_xy = reinterpret_cast<double2 *>(s_raw);
  double *s_z = reinterpret_cast<double *>(s_xy + TILE_CAP);
  int *s_ts = reinterpret_cast<int *>(s_z + TILE_CAP);
  int *s_j = s_ts + TILE_CAP;
  int *s_outer = s_j + TILE_CAP;                 // [warps][stride] skin entries of the row being built ...
  unsigned *s_okey = reinterpret_cast<unsigned *>(s_outer + (TILE_BS / 32) * stride);   // ... their sort keys (first: the type of j) ...
  double *s_orsq = reinterpret_cast<double *>(s_okey + (TILE_BS / 32) * stride);         // ... and their squared distances
  __shared__ int s_rb[18], s_re[18], s_roff[18], s_pre[19];
  constexpr int NW = TILE_BS / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // owned cell of this CTA
  const int cell = blockIdx.x;
  const int ix = cell % g.ninner[0] + 1, iy = (cell / g.ninner[0]) % g.ninner[1] + 1, iz = cell / (g.ninner[0] * g.ninner[1]) + 1;
  const int c = (iz * g.nc[1] + iy) * g.nc[0] + ix;
  const int sb = ostart[c], se = ostart[c + 1];
  if (sb >= se) return;
  if (threadIdx.x < 18) {
    const int r = threadIdx.x, yz = r >> 1, pass = r & 1;
    const int dz = yz / 3 - 1, dy = yz % 3 - 1;
    const int c0 = ((iz + dz) * g.nc[1] + (iy + dy)) * g.nc[0] + (ix - 1);
    s_rb[r] = pass == 0 ? ostart[c0] : gstart[c0];
    s_re[r] = pass == 0 ? ostart[c0 + 3] : gstart[c0 + 3];
    s_roff[r] = pass == 0 ? 0 : nlocal;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < 18; r++) { s_pre[r] = acc; acc += s_re[r] - s_rb[r]; }
    s_pre[18] = acc;
  }
  __syncthreads();
  const int ncand = s_pre[18];
  int *outer = s_outer + wid * stride;
  unsigned *okey = s_okey + wid * stride;
  double *orsq = s_orsq + wid * stride;
  const unsigned lt = (1u << lane) - 1;
  const double inv_skin = skin > 0.0 ? 1.0 / skin : 0.0;
  // per-warp site cursor state lives in registers across chunks: one warp owns sites sb+wid, sb+wid+NW, ...
  // (counts are kept per site in numneigh scratch when several chunks are needed)
  const bool single = ncand <= cap;
  for (int cbase = 0; cbase < ncand; cbase += cap) {
    const int cn = min(cap, ncand - cbase);
    __syncthreads();
    for (int k = threadIdx.x; k < cn; k += TILE_BS) {
      const int q = cbase + k;
      int r = 0;
#pragma unroll
      for (int t = 1; t < 18; t++) r += (q >= s_pre[t]);
      const int j = s_rb[r] + (q - s_pre[r]) + s_roff[r];
      const double4 rj = pos[j];
      s_xy[k] = make_double2(rj.x, rj.y);
      s_z[k] = rj.z;
      s_ts[k] = ts[j] & 0xffff;
      s_j[k] = j;
    }
    __syncthreads();
    for (int i = sb + wid; i < se; i += NW) {
      const double4 ri = pos[i];
      const PairInfo *prow = pinfo + (ts[i] & 0xffff) * na;
      // one actual type (the benchmark liquids): the two cutoffs are loop invariants
      const bool one_type = na == 2;
      const double cn1 = prow[1].cutneighsq, cs1 = prow[1].cutsq;
      int *row = neigh + (size_t)i * stride;
      int cnt_in = 0, cnt_out = 0;
      if (!single && cbase > 0) {   // resume: counts parked in the row tail / numneigh
        cnt_in = numneigh[i] & 0xffff;
        cnt_out = numneigh[i] >> 16;
        for (int k = lane; k < min(cnt_out, stride); k += 32) outer[k] = row[rowslot(stride - 1 - k)];
        __syncwarp();
      }
      for (int base = 0; base < cn; base += 32) {
        const int k = base + lane;
        bool hit = false, inner = false;
        int j = -1;
        double rsq_k = 0.0, cs_k = cs1;
        if (k < cn) {
          j = s_j[k];
          const double2 rxy = s_xy[k];
          rsq_k = rsq_exact(ri.x - rxy.x, ri.y - rxy.y, ri.z - s_z[k]);
          double cn = cn1;
          if (!one_type) { const PairInfo pi = prow[s_ts[k]]; cn = pi.cutneighsq; cs_k = pi.cutsq; }
          hit = (j != i) && (rsq_k <= cn);
          inner = hit && (rsq_k < cs_k);
        }
        const unsigned m_in = __ballot_sync(0xffffffffu, inner);
        const unsigned m_out = __ballot_sync(0xffffffffu, hit && !inner);
        if (inner) {
          const int p = cnt_in + __popc(m_in & lt);
          if (p < stride) row[rowslot(p)] = j;
        } else if (hit) {
          const int p = cnt_out + __popc(m_out & lt);
          if (p < stride) {
            // only a few lanes take this branch in every pass over the candidates: the two square roots of the sort
            // key are taken later, once per row, with all lanes busy
            outer[p] = j;
            if (defer_keys) { orsq[p] = rsq_k; okey[p] = (unsigned)s_ts[k]; }
            else {
              const double beyond = (sqrt(rsq_k) - sqrt(cs_k)) * (1.0 - 1e-9) * inv_skin;
              okey[p] = (unsigned)fmin(fmax(beyond, 0.0) * 134217728.0, 134217727.0);
            }
          }
        }
        cnt_in += __popc(m_in);
        cnt_out += __popc(m_out);
      }
      __syncwarp();
      const bool last = cbase + cap >= ncand;
      const int total = cnt_in + cnt_out;
      if (last) {
        unsigned lc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // entries to visit at displacement level 0..7
        if (total <= stride) {
          if (single) {
            // key = how far beyond the cutoff the pair sits, in units of skin / 2^27, rounded DOWN with a
            // safety margin: key >> 24 is the displacement level (eighths of the skin) below which the
            // pair cannot have entered the cutoff yet
            for (int k = lane; defer_keys && k < cnt_out; k += 32) {
              const double cs_k = one_type ? cs1 : prow[okey[k]].cutsq;
              const double beyond = (sqrt(orsq[k]) - sqrt(cs_k)) * (1.0 - 1e-9) * inv_skin;
              okey[k] = (unsigned)fmin(fmax(beyond, 0.0) * 134217728.0, 134217727.0);
            }
            __syncwarp();
            // skin entries in ascending distance (rank sort inside the warp), and the level counts
            for (int k = lane; k < ((cnt_out + 31) & ~31); k += 32) {
              const bool have = k < cnt_out;
              const unsigned key = have ? okey[k] : 0xffffffffu;
              int rank = 0;
              if (have)
                for (int m = 0; m < cnt_out; m++) {
                  const unsigned km = okey[m];
                  rank += (km < key) || (km == key && m < k);
                }
              if (have) row[rowslot(cnt_in + rank)] = outer[k];
              const unsigned lev = have ? min(key >> 24, 7u) : 8u;
#pragma unroll
              for (int L = 0; L < 8; L++) lc[L] += __popc(__ballot_sync(0xffffffffu, lev <= (unsigned)L));
            }
#pragma unroll
            for (int L = 0; L < 8; L++) lc[L] += cnt_in;
          } else {
            for (int k = lane; k < cnt_out; k += 32) row[rowslot(cnt_in + k)] = outer[k];
#pragma unroll
            for (int L = 0; L < 8; L++) lc[L] = total;   // chunked build: keys of earlier chunks are gone, visit everything
          }
        }
        if (lane == 0) {
          numneigh[i] = min(total, stride);
          if (total > stride) atomicMax(&flags[1], total);
          levcnt[i] = make_uint4(lc[0] | (lc[1] << 16), lc[2] | (lc[3] << 16), lc[4] | (lc[5] << 16), lc[6] | (lc[7] << 16));
        }
      } else {
        // park the skin entries at the row's tail (reversed) until the next chunk; if the row is about
        // to overflow the final count reports it and the build is redone with a larger stride
        if (total <= stride)
          for (int k = lane; k < cnt_out; k += 32) row[rowslot(stride - 1 - k)] = outer[k];
        if (lane == 0) numneigh[i] = (min(cnt_in, 0xffff)) | (min(cnt_out, 0x7fff) << 16);
      }
      __syncwarp();
    }
  }
}

// displacement level of a skin entry from its squared distance: the number of thresholds ((cut + k skin / 8)^2, rounded
// up) it has reached, k = 1..7
struct LevelThresholds { float t[7]; };
__device__ __forceinline__ int level_of(float r2, const LevelThresholds &lv) {
  int n = 0;
#pragma unroll
  for (int k = 0; k < 7; k++) n += (r2 >= lv.t[k]);
  return n;
}

// Single-precision prefilter variant of k_build_rows_tiled for one actual type (one cutoff pair): the SAME rows in the
// SAME order, decided exactly.  Candidates are staged as float4 {x - ox, y - oy, z - oz, index} relative to the centre
// of the CTA's cell (16 bytes instead of 40, one LDS.128 per test), the squared distance is formed in FP32 and decides
// the test whenever it lies further than `band` from both thresholds (cutneighsq, cutsq); `band` bounds the FP32 error
// for every candidate within reach (host: 4 x [6 sqrt(3) E r_c + 4 r_c^2] 2^-24, E = largest |relative coordinate|,
// r_c = cut + skin).  The ~1e-5 of the candidates inside the band are re-tested with the reference's FP64 arithmetic on
// the records in global memory, so classification — and with it every row — is bit-identical to the FP64 kernel
// (tests/test_gpu_parity.py::test_neighbor_build_variants_give_identical_rows).  Exact squared distances are needed
// again only for the sort keys of the skin entries (~23 per row): they are recomputed in the dense per-row pass.
// ~30 instead of ~75 instructions per 32 candidates.
__global__ void __launch_bounds__(TILE_BS)
k_build_rows_tiled_f32(const double4 *__restrict__ pos, int nlocal, Grid g, const int *__restrict__ ostart,
                       const int *__restrict__ gstart, double cutneighsq, double cutsq, float band,
                       int *__restrict__ neigh, int stride, int *__restrict__ numneigh, int *__restrict__ flags,
                       int cap, uint4 *__restrict__ levcnt, double skin, int cull, LevelThresholds lvl) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4 *s_c = reinterpret_cast<float4 *>(s_raw);                       // [TILE_CAP_F32] staged candidates
  int *s_inner = reinterpret_cast<int *>(s_c + TILE_CAP_F32);            // [warps][stride] inner entries found in this chunk
  int *s_outer = s_inner + (TILE_BS / 32) * stride;                      // [warps][stride] skin entries of the row being built
  unsigned *s_okey = reinterpret_cast<unsigned *>(s_outer + (TILE_BS / 32) * stride);
  __shared__ int s_rb[18], s_re[18], s_roff[18], s_pre[19];
  constexpr int NW = TILE_BS / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cell = blockIdx.x;
  const int ix = cell % g.ninner[0] + 1, iy = (cell / g.ninner[0]) % g.ninner[1] + 1, iz = cell / (g.ninner[0] * g.ninner[1]) + 1;
  const int c = (iz * g.nc[1] + iy) * g.nc[0] + ix;
  const int sb = ostart[c], se = ostart[c + 1];
  if (sb >= se) return;
  if (threadIdx.x < 18) {
    const int r = threadIdx.x, yz = r >> 1, pass = r & 1;
    const int dz = yz / 3 - 1, dy = yz % 3 - 1;
    const int c0 = ((iz + dz) * g.nc[1] + (iy + dy)) * g.nc[0] + (ix - 1);
    s_rb[r] = pass == 0 ? ostart[c0] : gstart[c0];
    s_re[r] = pass == 0 ? ostart[c0 + 3] : gstart[c0 + 3];
    s_roff[r] = pass == 0 ? 0 : nlocal;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < 18; r++) { s_pre[r] = acc; acc += s_re[r] - s_rb[r]; }
    s_pre[18] = acc;
  }
  __syncthreads();
  const int ncand = s_pre[18];
  // centre of this cell: relative coordinates of everything in the 27-cell stencil stay below 1.5 cell edges
  const double ox = g.lo[0] + ((double)ix - 0.5) / g.inv[0], oy = g.lo[1] + ((double)iy - 0.5) / g.inv[1],
               oz = g.lo[2] + ((double)iz - 0.5) / g.inv[2];
  const float cnf = (float)cutneighsq, csf = (float)cutsq;
  // half edges of the cell, padded by the rounding of the owned sites' cell assignment at the faces
  // (cull == 0: a non-periodic dimension, where sites beyond the box face are binned into the edge cells — no culling)
  const float hx = cull ? (float)(0.5 / g.inv[0]) * 1.0001f : 1e30f, hy = cull ? (float)(0.5 / g.inv[1]) * 1.0001f : 1e30f,
              hz = cull ? (float)(0.5 / g.inv[2]) * 1.0001f : 1e30f;
  __shared__ int s_wcnt[TILE_BS / 32];
  int *inner_buf = s_inner + wid * stride;
  int *outer = s_outer + wid * stride;
  unsigned *okey = s_okey + wid * stride;
  const unsigned lt = (1u << lane) - 1;
  const double inv_skin = skin > 0.0 ? 1.0 / skin : 0.0;
  const double rcut = sqrt(cutsq);
  const bool single = ncand <= cap;
  for (int cbase = 0; cbase < ncand; cbase += cap) {
    const int cn_all = min(cap, ncand - cbase);
    __syncthreads();
    // Stage this chunk, dropping every candidate that lies further than cut + skin from the CELL (its box in
    // relative coordinates is [-h, h]^3): such a candidate is a neighbor of none of the cell's sites.  About a quarter
    // of the 27-cell stencil goes this way.  The compaction is stable, so rows keep their order.
    int cn = 0;
    for (int kb = 0; kb < cn_all; kb += TILE_BS) {
      const int k = kb + threadIdx.x;
      bool keep = false;
      float4 rec;
      if (k < cn_all) {
        const int q = cbase + k;
        int r = 0;
#pragma unroll
        for (int t = 1; t < 18; t++) r += (q >= s_pre[t]);
        const int j = s_rb[r] + (q - s_pre[r]) + s_roff[r];
        const double4 rj = pos[j];
        rec = make_float4((float)(rj.x - ox), (float)(rj.y - oy), (float)(rj.z - oz), __int_as_float(j));
        const float ex = fmaxf(fabsf(rec.x) - hx, 0.f), ey = fmaxf(fabsf(rec.y) - hy, 0.f), ez = fmaxf(fabsf(rec.z) - hz, 0.f);
        keep = fmaf(ez, ez, fmaf(ey, ey, ex * ex)) <= cnf + band;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_wcnt[wid] = __popc(m);
      __syncthreads();
      int off = cn;
      for (int w = 0; w < wid; w++) off += s_wcnt[w];
      if (keep) s_c[off + __popc(m & lt)] = rec;
      int tot = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) tot += s_wcnt[w];
      cn += tot;
      __syncthreads();
    }
    const int cn32 = (cn + 31) & ~31;       // cap is a multiple of 32: the padded tail stays inside the staging array
    for (int k = cn + threadIdx.x; k < cn32; k += TILE_BS)
      s_c[k] = make_float4(1e18f, 1e18f, 1e18f, __int_as_float(-1));   // never a hit, never inside the band
    __syncthreads();
    for (int i = sb + wid; i < se; i += NW) {
      const double4 ri = pos[i];
      const float xf = (float)(ri.x - ox), yf = (float)(ri.y - oy), zf = (float)(ri.z - oz);
      int *row = neigh + (size_t)i * stride;
      int cnt_in = 0, cnt_out = 0;
      if (!single && cbase > 0) {   // resume: counts parked in numneigh, skin entries at the row's tail
        cnt_in = numneigh[i] & 0xffff;
        cnt_out = numneigh[i] >> 16;
        for (int k = lane; k < min(cnt_out, stride); k += 32) outer[k] = row[rowslot(stride - 1 - k)];
        __syncwarp();
      }
      const int cnt_in0 = cnt_in;
      const unsigned sa_inner = (unsigned)__cvta_generic_to_shared(inner_buf) - 4u * (unsigned)cnt_in0;
      const unsigned sa_outer = (unsigned)__cvta_generic_to_shared(outer), sa_okey = (unsigned)__cvta_generic_to_shared(okey);
      for (int base = 0; base < cn32; base += 32) {
        const float4 cj = s_c[base + lane];
        const int j = __float_as_int(cj.w);
        const float dx = xf - cj.x, dy = yf - cj.y, dz = zf - cj.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        bool hit = r2 <= cnf, inner = r2 < csf;
        if ((fabsf(r2 - cnf) <= band) | (fabsf(r2 - csf) <= band)) {     // one branch (no short-circuit)
          // too close to a threshold for FP32: the reference's arithmetic on the FP64 records
          const double4 rj = pos[j];
          const double rsq = rsq_exact(ri.x - rj.x, ri.y - rj.y, ri.z - rj.z);
          hit = rsq <= cutneighsq;
          inner = rsq < cutsq;
        }
        hit = hit && (j != i);
        inner = inner && hit;
        const unsigned m_in = __ballot_sync(0xffffffffu, inner);
        const unsigned m_out = __ballot_sync(0xffffffffu, hit && !inner);
        if (hit) {
          // both kinds go to shared memory; the row is written in one dense pass below
          const int p = inner ? cnt_in + __popc(m_in & lt) : cnt_out + __popc(m_out & lt);
          if (p < stride) {
            // 32-bit shared-window addresses taken once per site: the generic-pointer form makes the compiler rebuild
            // the window base (S2UR SR_CgaCtaId, ULEA ...) inside this branch on every pass
            const unsigned a = (inner ? sa_inner : sa_outer) + 4u * (unsigned)p;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(j) : "memory");
            if (!inner) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa_okey + 4u * (unsigned)p), "r"(__float_as_uint(r2)) : "memory");   // for the displacement level of this skin entry
          }
        }
        cnt_in += __popc(m_in);
        cnt_out += __popc(m_out);
      }
      __syncwarp();
      const bool last = cbase + cap >= ncand;
      const int total = cnt_in + cnt_out;
      if (total <= stride)
        for (int k = cnt_in0 + lane; k < cnt_in; k += 32) row[rowslot(k)] = inner_buf[k - cnt_in0];
      if (last) {
        unsigned lc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (total <= stride) {
          if (single) {
            // The pair kernels only need the skin entries grouped by displacement level (the eighth of the skin the
            // pair sat in at build time), not sorted by distance: a counting sort over 8 levels, stable in candidate
            // order.  The level comes from the FP32 squared distance kept with the hit, pushed DOWN by the error band
            // (a lower level is visited earlier: always safe).  No FP64 gather, no square root, no rank sort.
            if (cnt_out <= 32) {
              // the usual case (~23 skin entries): one entry per lane, every ballot taken once
              const int lev = lane < cnt_out ? level_of(__uint_as_float(okey[lane]) - band, lvl) : 8;
              int running = cnt_in, mypos = 0;
#pragma unroll
              for (int L = 0; L < 8; L++) {
                const unsigned m = __ballot_sync(0xffffffffu, lev == L);
                if (lev == L) mypos = running + __popc(m & lt);
                running += __popc(m);
                lc[L] = running;
              }
              if (lev < 8) row[rowslot(mypos)] = outer[lane];
            } else {
            int cntL[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int kb = 0; kb < cnt_out; kb += 32) {
              const int k = kb + lane;
              const int lev = k < cnt_out ? level_of(__uint_as_float(okey[k]) - band, lvl) : 8;
#pragma unroll
              for (int L = 0; L < 8; L++) cntL[L] += __popc(__ballot_sync(0xffffffffu, lev == L));
            }
            int off[8], running = cnt_in;
#pragma unroll
            for (int L = 0; L < 8; L++) { off[L] = running; running += cntL[L]; lc[L] = running; }
            for (int kb = 0; kb < cnt_out; kb += 32) {
              const int k = kb + lane;
              const int lev = k < cnt_out ? level_of(__uint_as_float(okey[k]) - band, lvl) : 8;
#pragma unroll
              for (int L = 0; L < 8; L++) {
                const unsigned m = __ballot_sync(0xffffffffu, lev == L);
                if (lev == L) row[rowslot(off[L] + __popc(m & lt))] = outer[k];
                off[L] += __popc(m);
              }
            }
            }
          } else {
            for (int k = lane; k < cnt_out; k += 32) row[rowslot(cnt_in + k)] = outer[k];
#pragma unroll
            for (int L = 0; L < 8; L++) lc[L] = total;
          }
        }
        if (lane == 0) {
          numneigh[i] = min(total, stride);
          if (total > stride) atomicMax(&flags[1], total);
          levcnt[i] = make_uint4(lc[0] | (lc[1] << 16), lc[2] | (lc[3] << 16), lc[4] | (lc[5] << 16), lc[6] | (lc[7] << 16));
        }
      } else {
        if (total <= stride)
          for (int k = lane; k < cnt_out; k += 32) row[rowslot(stride - 1 - k)] = outer[k];
        if (lane == 0) numneigh[i] = (min(cnt_in, 0xffff)) | (min(cnt_out, 0x7fff) << 16);
      }
      __syncwarp();
    }
  }
}

// Lane-per-site variant of k_build_rows_tiled_f32 (experiment, UCGB200_BUILD_LANES=1): the SAME rows in the SAME order.
// One warp owns one cell.  The candidates of the 27-cell stencil are staged and culled exactly as above; then every
// LANE takes one of the cell's sites and walks the staged candidates one by one (the float4 is a shared-memory
// broadcast: one wavefront, no conflict), so a distance test costs ~11 warp instructions per 32 (site, candidate)
// pairs instead of ~57 per 32 with a warp per site (profiles/r02_build_rows_lines.json: ballots, ranks and the band
// test ran once per candidate AND per site there).  Pass A keeps the tile index of everything within
// cutneighsq + band in a 16-bit list per lane; pass B1 classifies only those ~80 entries per site (exact re-test
// inside the band, self exclusion, inner / displacement level) with every lane busy, pass B2 places them: inner
// entries first, skin entries grouped by level, both in candidate order.  Level counters live in one 64-bit word
// (8 x 8 bits; a multiply by 0x0101..01 gives the prefix sums).
constexpr int LANE_CAP = 512;                      // staged candidates per cell (after the cull); more: flags[2], generic kernel
__global__ void __launch_bounds__(32)
k_build_rows_lanes_f32(const double4 *__restrict__ pos, int nlocal, Grid g, const int *__restrict__ ostart,
                       const int *__restrict__ gstart, double cutneighsq, double cutsq, float band,
                       int *__restrict__ neigh, int stride, int *__restrict__ numneigh, int *__restrict__ flags,
                       uint4 *__restrict__ levcnt, int cull, LevelThresholds lvl) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4 *s_c = reinterpret_cast<float4 *>(s_raw);                                   // [LANE_CAP] staged candidates
  unsigned short *s_hit = reinterpret_cast<unsigned short *>(s_c + LANE_CAP);        // [32][HS] tile indices (+ class) per lane
  const int HS = stride + 2;                       // HS / 2 odd: the 32 lists start in 32 different banks
  const int lane = threadIdx.x;
  const unsigned lt = (1u << lane) - 1;
  const int cell = blockIdx.x;
  const int ix = cell % g.ninner[0] + 1, iy = (cell / g.ninner[0]) % g.ninner[1] + 1, iz = cell / (g.ninner[0] * g.ninner[1]) + 1;
  const int c = (iz * g.nc[1] + iy) * g.nc[0] + ix;
  const int sb = ostart[c], se = ostart[c + 1];
  if (sb >= se) return;
  // the 18 runs of the stencil (9 rows of three cells, owned sites then ghosts): lane r keeps run r
  int rb = 0, re = 0, roff = 0;
  if (lane < 18) {
    const int yz = lane >> 1, pass = lane & 1;
    const int dz = yz / 3 - 1, dy = yz % 3 - 1;
    const int c0 = ((iz + dz) * g.nc[1] + (iy + dy)) * g.nc[0] + (ix - 1);
    rb = pass == 0 ? ostart[c0] : gstart[c0];
    re = pass == 0 ? ostart[c0 + 3] : gstart[c0 + 3];
    roff = pass == 0 ? 0 : nlocal;
  }
  const double ox = g.lo[0] + ((double)ix - 0.5) / g.inv[0], oy = g.lo[1] + ((double)iy - 0.5) / g.inv[1],
               oz = g.lo[2] + ((double)iz - 0.5) / g.inv[2];
  const float cnf = (float)cutneighsq, csf = (float)cutsq, cnb = cnf + band;
  const float hx = cull ? (float)(0.5 / g.inv[0]) * 1.0001f : 1e30f, hy = cull ? (float)(0.5 / g.inv[1]) * 1.0001f : 1e30f,
              hz = cull ? (float)(0.5 / g.inv[2]) * 1.0001f : 1e30f;
  // stage + cull (stable): identical to k_build_rows_tiled_f32, run by run
  int cn = 0;
  for (int r = 0; r < 18; r++) {
    const int b = __shfl_sync(0xffffffffu, rb, r), e = __shfl_sync(0xffffffffu, re, r), off = __shfl_sync(0xffffffffu, roff, r);
    for (int k0 = b; k0 < e; k0 += 32) {
      const int k = k0 + lane;
      bool keep = false;
      float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < e) {
        const int j = k + off;
        const double4 rj = pos[j];
        rec = make_float4((float)(rj.x - ox), (float)(rj.y - oy), (float)(rj.z - oz), __int_as_float(j));
        const float ex = fmaxf(fabsf(rec.x) - hx, 0.f), ey = fmaxf(fabsf(rec.y) - hy, 0.f), ez = fmaxf(fabsf(rec.z) - hz, 0.f);
        keep = fmaf(ez, ez, fmaf(ey, ey, ex * ex)) <= cnb;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      const int at = cn + __popc(m & lt);
      if (keep && at < LANE_CAP) s_c[at] = rec;
      cn += __popc(m);
    }
  }
  if (cn > LANE_CAP) {                                  // a cell this crowded goes to the generic kernel (host re-launch)
    if (lane == 0) atomicMax(&flags[2], cn);
    return;
  }
  const int cn4 = (cn + 3) & ~3;                      // LANE_CAP is a multiple of 4: the padded tail stays inside the array
  if (lane < cn4 - cn) s_c[cn + lane] = make_float4(1e18f, 1e18f, 1e18f, __int_as_float(-1));   // never a hit
  __syncwarp();
  const unsigned sa_hit = (unsigned)__cvta_generic_to_shared(s_hit) + 2u * (unsigned)(lane * HS);
  for (int s0 = sb; s0 < se; s0 += 32) {
    const int i = s0 + lane;
    const bool have = i < se;
    double4 ri = make_double4(0, 0, 0, 0);
    float xf = 1e18f, yf = 1e18f, zf = 1e18f;        // an idle lane hits nothing
    if (have) {
      ri = pos[i];
      xf = (float)(ri.x - ox); yf = (float)(ri.y - oy); zf = (float)(ri.z - oz);
    }
    // ---- pass A: every staged candidate against every lane's site
    unsigned a = sa_hit;
    const unsigned a_end = sa_hit + 2u * (unsigned)HS;
    int nhit = 0;                                     // may run past HS: reported as an overflow
    for (int k0 = 0; k0 < cn4; k0 += 4) {              // four broadcast loads in flight, then the four tests
      float4 cj[4];
#pragma unroll
      for (int u = 0; u < 4; u++) cj[u] = s_c[k0 + u];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float dx = xf - cj[u].x, dy = yf - cj[u].y, dz = zf - cj[u].z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const bool h = r2 <= cnb;
        if (h && a < a_end) asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)(k0 + u)) : "memory");
        a += h ? 2u : 0u;
        nhit += h ? 1 : 0;
      }
    }
    const int nkeep = min(nhit, HS);
    // ---- pass B1: classify the kept entries: 0 inner, 1 + level for a skin entry, 15 not a neighbor
    int cnt_in = 0;
    unsigned long long lev64 = 0;
    const int nmax = __reduce_max_sync(0xffffffffu, nkeep);
    for (int e = 0; e < nmax; e++) {
      if (e < nkeep) {
        unsigned short kk;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(kk) : "r"(sa_hit + 2u * (unsigned)e));
        const float4 cj = s_c[kk];
        const int j = __float_as_int(cj.w);
        const float dx = xf - cj.x, dy = yf - cj.y, dz = zf - cj.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        bool hit = r2 <= cnf, inner = r2 < csf;
        if ((fabsf(r2 - cnf) <= band) | (fabsf(r2 - csf) <= band)) {
          const double4 rj = pos[j];
          const double rsq = rsq_exact(ri.x - rj.x, ri.y - rj.y, ri.z - rj.z);
          hit = rsq <= cutneighsq;
          inner = rsq < cutsq;
        }
        hit = hit && (j != i);
        int cls = 15;
        if (hit) {
          if (inner) { cls = 0; cnt_in++; }
          else {
            const int lev = level_of(r2 - band, lvl);
            cls = 1 + lev;
            lev64 += 1ull << (8 * lev);
          }
        }
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(sa_hit + 2u * (unsigned)e), "h"((unsigned short)(kk | (cls << 12))) : "memory");
      }
    }
    // ---- pass B2: place
    const unsigned long long incl = lev64 * 0x0101010101010101ull;      // byte L: skin entries of level <= L (cnt_out <= 255)
    int cnt_out = (int)(incl >> 56);
    bool levels = true;
    if (nhit > 255) {   // the byte counters may have wrapped: count again, keep candidate order, every level "visit all"
      levels = false;
      cnt_out = 0;
      for (int e = 0; e < nkeep; e++) { const int cls = s_hit[lane * HS + e] >> 12; cnt_out += (cls >= 1 && cls <= 8); }
    }
    const int total = (nhit > HS) ? nhit : cnt_in + cnt_out;   // an overfull list reports its (upper bound) length
    if (have) {
      int *row = neigh + (size_t)i * stride;
      if (total <= stride) {
        unsigned long long start = levels ? (incl << 8) : 0ull;         // byte L: entries of lower levels
        int in_pos = 0, out_pos = cnt_in;
        for (int e = 0; e < nkeep; e++) {
          unsigned short v;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(sa_hit + 2u * (unsigned)e));
          const int cls = v >> 12;
          if (cls == 15) continue;
          const int j = __float_as_int(s_c[v & 0xfff].w);
          int p;
          if (cls == 0) p = in_pos++;
          else if (levels) { const int sh = 8 * (cls - 1); p = cnt_in + (int)((start >> sh) & 0xffull); start += 1ull << sh; }
          else p = out_pos++;
          row[rowslot(p)] = j;
        }
      }
      unsigned lc[8];
#pragma unroll
      for (int L = 0; L < 8; L++) lc[L] = levels ? (unsigned)cnt_in + (unsigned)((incl >> (8 * L)) & 0xffull) : (unsigned)total;
      numneigh[i] = min(total, stride);
      if (total > stride) atomicMax(&flags[1], total);
      levcnt[i] = make_uint4(lc[0] | (lc[1] << 16), lc[2] | (lc[3] << 16), lc[4] | (lc[5] << 16), lc[6] | (lc[7] << 16));
    }
    __syncwarp();
  }
}

// [stock] Neighbor::check_distance: any owned atom moved > skin/2 since the last build
__global__ void k_check_distance(const double4 *__restrict__ pos, const double4 *__restrict__ xhold, int n,
                                 double triggersq, int *__restrict__ flags, unsigned long long *__restrict__ maxdisp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 a = pos[i], b = xhold[i];
  double rsq = rsq_exact(a.x - b.x, a.y - b.y, a.z - b.z);
  if (rsq > triggersq) flags[0] = 1;
  // largest squared displacement since the build (bits of a non-negative double order like integers)
  const unsigned long long bits = (unsigned long long)__double_as_longlong(rsq);
  if (bits > *maxdisp) atomicMax(maxdisp, bits);
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
    if (c->periodic[d] && len < cutneigh) {
      c->err = "periodic box (or brick) shorter than the neighbor cutoff (cut+skin)";
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

static ImageMap make_image_map(const ucgb200_ctx *c) {
  ImageMap im;
  const auto &h = c->halo;
  im.self = h.rank; im.nranks = h.nranks;
  for (int d = 0; d < 3; d++) {
    im.allow_lo[d] = c->periodic[d] || h.coord[d] > 0;
    im.allow_hi[d] = c->periodic[d] || h.coord[d] < h.grid[d] - 1;
  }
  for (int code = 0; code < 27; code++) {
    int cc[3] = {code % 3, (code / 3) % 3, code / 9};
    int pc[3];
    for (int d = 0; d < 3; d++) {
      im.shift[code][d] = 0.0;
      pc[d] = h.coord[d];
      if (cc[d] == 1) {  // towards the lower neighbour
        pc[d] = h.coord[d] - 1;
        if (pc[d] < 0) { pc[d] = h.grid[d] - 1; im.shift[code][d] = c->prd[d]; }
      } else if (cc[d] == 2) {
        pc[d] = h.coord[d] + 1;
        if (pc[d] >= h.grid[d]) { pc[d] = 0; im.shift[code][d] = -c->prd[d]; }
      }
    }
    im.dest[code] = (pc[2] * h.grid[1] + pc[1]) * h.grid[0] + pc[0];
  }
  return im;
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
  bool lanes_off = false;
  while (true) {
    UCG_CHECK(c, c->neigh.ensure((size_t)nlocal * c->neigh_stride));
    UCG_CHECK(c, c->levcnt.ensure(nlocal + 1));
    UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p + 1, 0, 2 * sizeof(int), c->stream));
    const int tiled = getenv("UCGB200_BUILD_TILED") ? atoi(getenv("UCGB200_BUILD_TILED")) : 1;
    int cap = getenv("UCGB200_TILE_CAP") ? atoi(getenv("UCGB200_TILE_CAP")) : TILE_CAP;   // small values exercise the chunked path
    cap = std::min(std::max(cap, 32), TILE_CAP);
    const int f32 = getenv("UCGB200_BUILD_F32") ? atoi(getenv("UCGB200_BUILD_F32")) : 1;
    if (tiled && f32 && na == 2 && c->h_pairinfo.size() == 4) {
      // one actual type: single-precision prefilter, exact re-test inside the error band
      // (k_build_rows_lanes_f32: a lane per site; k_build_rows_tiled_f32: a warp per site, any cell population)
      const int ncell_owned = c->grid.ninner[0] * c->grid.ninner[1] * c->grid.ninner[2];
      const size_t smem = (size_t)TILE_CAP_F32 * sizeof(float4) + (TILE_BS / 32) * (size_t)c->neigh_stride * (3 * sizeof(int));
      UCG_CHECK(c, cudaFuncSetAttribute(k_build_rows_tiled_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      if (smem > 96 * 1024) return fail(c, "neighbor rows too long for the tiled build");
      // measured (profiles/r02_build_rows_lanes_full.json): 0.89 G warp instructions instead of 1.41 G, but 16.5 KB of shared
      // memory per one-warp block leave 13 warps per SM (issue slots 56 % busy): 1.76 ms per rebuild against 1.66 ms -> off
      const int lanes_knob = getenv("UCGB200_BUILD_LANES") ? atoi(getenv("UCGB200_BUILD_LANES")) : 0;
      const size_t smem_lanes = (size_t)LANE_CAP * sizeof(float4) + 32 * (size_t)(c->neigh_stride + 2) * sizeof(unsigned short);
      const bool lanes = lanes_knob && !lanes_off && !getenv("UCGB200_TILE_CAP") && smem_lanes <= 96 * 1024 && c->nlocal + c->nghost < (1 << 30);
      const double cs = c->h_pairinfo[3].cutsq, cn = c->h_pairinfo[3].cutneighsq;   // the thresholds the FP64 kernels read
      const double rc = std::sqrt(cn);
      double edge = 0.0;
      for (int d = 0; d < 3; d++) edge = std::max(edge, 1.0 / c->grid.inv[d]);
      const double E = 1.5 * edge + 1e-6 * edge;
      const float band = (float)(4.0 * (6.0 * 1.7320508 * E * rc + 4.0 * rc * rc) * 5.9604644775390625e-08);
      // level k starts where the pair sits k eighths of the skin beyond the cutoff (the pair kernels visit an entry of
      // level l once 2 sqrt(maxdisp) has reached l eighths): thresholds rounded UP, so a level is never overstated
      LevelThresholds lvl;
      {
        const double rcut = std::sqrt(cs);
        for (int k = 1; k <= 7; k++) {
          const double r = rcut + k * c->skin / 8.0;
          lvl.t[k - 1] = std::nextafter((float)(r * r * (1.0 + 1e-6)), 1e30f);
        }
      }
      int cap32 = getenv("UCGB200_TILE_CAP") ? atoi(getenv("UCGB200_TILE_CAP")) : TILE_CAP_F32;
      cap32 = (std::min(std::max(cap32, 32), TILE_CAP_F32) / 32) * 32;   // whole passes of 32 candidates
      const int cull = (c->periodic[0] && c->periodic[1] && c->periodic[2] && !(getenv("UCGB200_BUILD_CULL") && atoi(getenv("UCGB200_BUILD_CULL")) == 0)) ? 1 : 0;
      if (lanes) {
        UCG_CHECK(c, cudaFuncSetAttribute(k_build_rows_lanes_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        k_build_rows_lanes_f32<<<ncell_owned, 32, smem_lanes, c->stream>>>(
            c->pos.p, nlocal, c->grid, c->cell_start.p, c->gcell_start.p, cn, cs, band, c->neigh.p, c->neigh_stride,
            c->numneigh.p, c->d_flags.p, c->levcnt.p, cull, lvl);
      } else {
        k_build_rows_tiled_f32<<<ncell_owned, TILE_BS, smem, c->stream>>>(
            c->pos.p, nlocal, c->grid, c->cell_start.p, c->gcell_start.p, cn, cs, band, c->neigh.p, c->neigh_stride,
            c->numneigh.p, c->d_flags.p, cap32, c->levcnt.p, c->skin, cull, lvl);
      }
    } else if (tiled) {
      const int ncell_owned = c->grid.ninner[0] * c->grid.ninner[1] * c->grid.ninner[2];
      const size_t smem = (size_t)TILE_CAP * (sizeof(double2) + sizeof(double) + 2 * sizeof(int)) +
                          (TILE_BS / 32) * (size_t)c->neigh_stride * (2 * sizeof(int) + sizeof(double));
      // the attribute is per device: set it on every call (a process may hold contexts on several GPUs)
      UCG_CHECK(c, cudaFuncSetAttribute(k_build_rows_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      if (smem > 96 * 1024) return fail(c, "neighbor rows too long for the tiled build");
      k_build_rows_tiled<<<ncell_owned, TILE_BS, smem, c->stream>>>(
          c->pos.p, c->ts.p, nlocal, c->grid, c->cell_start.p, c->gcell_start.p, c->d_pairinfo.p, na,
          c->neigh.p, c->neigh_stride, c->numneigh.p, c->d_flags.p, cap, c->levcnt.p, c->skin,
          getenv("UCGB200_BUILD_DEFER_KEYS") ? atoi(getenv("UCGB200_BUILD_DEFER_KEYS")) : 1);
    } else {
      long long nthreads = (long long)nlocal * 32;
      k_build_rows<<<nblocks(nthreads, 256), 256, 8 * c->neigh_stride * sizeof(int), c->stream>>>(
          c->pos.p, c->ts.p, nlocal, c->grid, c->cell_start.p, c->gcell_start.p, c->d_pairinfo.p, na,
          c->neigh.p, c->neigh_stride, c->numneigh.p, c->d_flags.p, c->levcnt.p);
    }
    UCG_LAUNCHED(c);
    int rc = read_flags(c);
    if (rc) return rc;
    if (c->h_flags[2]) { lanes_off = true; continue; }   // a cell beyond LANE_CAP candidates: this build goes to the warp-per-site kernel
    if (c->h_flags[1] <= c->neigh_stride) break;
    // UCGB200_ERR_NEIGH_OVERFLOW handled internally: grow the row capacity and redo
    c->neigh_stride = ((c->h_flags[1] + c->h_flags[1] / 8 + 15) / 16) * 16;
  }
  return 0;
}

// permute every per-site array by order[0..n) into the twin buffers and swap
static int permute_all(ucgb200_ctx *c, int n) {
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
  if (n > 0) {
    k_permute<<<nblocks(n, 256), 256, 0, c->stream>>>(pa, c->order.p, n);
    UCG_LAUNCHED(c);
  }
  std::swap(c->pos, c->pos_alt); std::swap(c->vel, c->vel_alt); std::swap(c->frc, c->frc_alt);
  std::swap(c->scores, c->scores_alt); std::swap(c->ucgp, c->ucgp_alt); std::swap(c->ucgml, c->ucgml_alt);
  std::swap(c->ts, c->ts_alt); std::swap(c->mask, c->mask_alt); std::swap(c->tag, c->tag_alt);
  std::swap(c->mol, c->mol_alt); std::swap(c->orig, c->orig_alt);
  return 0;
}

// UCGB200_BUILD_TRACE=1: wall-clock marks (with a stream synchronisation each) through a rebuild, to stderr
struct BuildTrace {
  ucgb200_ctx *c;
  bool on;
  double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit BuildTrace(ucgb200_ctx *ctx) : c(ctx), on(getenv("UCGB200_BUILD_TRACE") && atoi(getenv("UCGB200_BUILD_TRACE"))), t0(0) {
    if (on) { cudaStreamSynchronize(c->stream); t0 = now(); }
  }
  void mark(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(c->stream);
    const double t = now();
    fprintf(stderr, "[build] %-28s %8.3f ms\n", what, t - t0);
    t0 = t;
  }
};

struct BuildTimer {     // neighbor stage (slot 1) of ucgb200_timers: blocking or from the asynchronous event ring
  ucg::StageTimer t;
  ucgb200_ctx *c;
  long long l0;
  explicit BuildTimer(ucgb200_ctx *ctx) : t(ctx, 1), c(ctx), l0(ctx->launches) {}
  void stop() {
    if (c->timers_on) c->t_launch[1] += c->launches - l0;
    t.stop();
  }
};

// rebuild, local part: pbc wrap, cell sort of the owned sites, image lists (local periodic
// images + per-rank border send lists)
extern "C" int ucgb200_neigh_build_local(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if ((rc = setup_grid(c))) return rc;
  cudaStream_t st = c->stream;
  const int nlocal = c->nlocal;
  const int ncells = c->grid.ncells;
  BoxDev box = make_box(c);
  auto &h = c->halo;
  BuildTimer timer(c);
  c->list_valid = false;
  c->parts_valid = false;

  UCG_CHECK(c, c->cell_count.ensure(ncells + 4));
  UCG_CHECK(c, c->cell_start.ensure(ncells + 4));
  UCG_CHECK(c, c->cell_cursor.ensure(ncells + 4));
  UCG_CHECK(c, c->gcell_count.ensure(ncells + 4));
  UCG_CHECK(c, c->gcell_start.ensure(ncells + 4));
  UCG_CHECK(c, c->order.ensure(nlocal + 1));
  UCG_CHECK(c, c->cell_of.ensure(nlocal + 1));
  UCG_CHECK(c, c->numneigh.ensure(nlocal + 1));
  UCG_CHECK(c, cudaMemsetAsync(c->cell_count.p, 0, (ncells + 4) * sizeof(int), st));
  UCG_CHECK(c, cudaMemsetAsync(c->cell_cursor.p, 0, (ncells + 4) * sizeof(int), st));
  UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p, 0, 8 * sizeof(int), st));
  h.nlimg = 0; h.nsend = 0;
  std::fill(h.send_counts.begin(), h.send_counts.end(), 0);
  if (nlocal == 0) { timer.stop(); return 0; }

  // 1. pbc wrap + owned-cell histogram
  BuildTrace trace(c);
  k_wrap_count<<<nblocks(nlocal, 256), 256, 0, st>>>(c->pos.p, nlocal, box, c->grid, c->cell_of.p,
                                                     c->cell_count.p, c->d_flags.p);
  UCG_LAUNCHED(c);
  trace.mark("wrap_count");
  // 2. cell offsets (ncells+4 entries so that start[c0+3] is always readable)
  if ((rc = exclusive_scan(c, c->cell_count.p, c->cell_start.p, ncells + 4, nullptr))) return rc;
  // 3. counting sort into cells, deterministic inside each cell
  k_fill_order<<<nblocks(nlocal, 256), 256, 0, st>>>(c->cell_of.p, nlocal, c->cell_start.p, c->cell_cursor.p, c->order.p);
  UCG_LAUNCHED(c);
  k_sort_cells<int><<<nblocks(ncells, 128), 128, 0, st>>>(c->order.p, c->cell_start.p, ncells);
  UCG_LAUNCHED(c);
  // 4. move every per-site array into cell order
  trace.mark("scan+order+sort");
  if ((rc = permute_all(c, nlocal))) return rc;
  trace.mark("permute");

  // 5. image lists: counters [0..nranks) = border records per destination rank, [nranks] = local
  const int nr = h.nranks;
  ImageMap im = make_image_map(c);
  UCG_CHECK(c, c->img_counters.ensure(2 * (nr + 1) + 4));
  int *counters = c->img_counters.p, *offsets = c->img_counters.p + (nr + 1);
  UCG_CHECK(c, cudaMemsetAsync(counters, 0, (nr + 1) * sizeof(int), st));
  k_images<false><<<nblocks(nlocal, 256), 256, 0, st>>>(c->pos.p, nlocal, box, im, counters, nullptr, nullptr, nullptr);
  UCG_LAUNCHED(c);
  std::vector<int> hc(nr + 1), ho(nr + 2, 0);
  UCG_CHECK(c, cudaMemcpyAsync(hc.data(), counters, (nr + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  if ((rc = read_flags(c))) return rc;
  trace.mark("images count + readback");
  if (c->h_flags[3]) { c->err = "atoms lost: position outside the periodic box by more than one period"; return UCGB200_ERR_LOST_ATOMS; }
  for (int k = 0; k <= nr; k++) ho[k + 1] = ho[k] + hc[k];
  const int ntotal = ho[nr + 1];
  h.nsend = ho[nr];
  h.nlimg = hc[nr];
  h.send_counts.assign(hc.begin(), hc.begin() + nr);
  h.send_offsets.assign(ho.begin(), ho.begin() + nr + 1);
  UCG_CHECK(c, c->img_owner.ensure(ntotal + 1));
  UCG_CHECK(c, c->img_code.ensure(ntotal + 1));
  if (ntotal > 0) {
    UCG_CHECK(c, cudaMemcpyAsync(offsets, ho.data(), (nr + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    UCG_CHECK(c, cudaMemsetAsync(counters, 0, (nr + 1) * sizeof(int), st));
    k_images<true><<<nblocks(nlocal, 256), 256, 0, st>>>(c->pos.p, nlocal, box, im, counters, offsets, c->img_owner.p,
                                                         c->img_code.p);
    UCG_LAUNCHED(c);
  }
  trace.mark("images fill");
  timer.stop();
  return 0;
}

// rebuild, final part: bin the ghost sources (local images + received border records) into
// cells, materialise the ghost sites, build the neighbor rows, remember positions
extern "C" int ucgb200_neigh_build_finish(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  auto &h = c->halo;
  const int nlocal = c->nlocal;
  const int ncells = c->grid.ncells;
  BuildTimer timer(c);
  ImageMap im = make_image_map(c);
  const int nlimg = h.nlimg, nrecv = h.nrecv, nsrc = nlimg + nrecv;
  const int *lo = c->img_owner.p + h.nsend, *lc = c->img_code.p + h.nsend;  // local images follow the send lists
  const BorderRec *recv = (const BorderRec *)c->recv_border.p;
  int rc;
  UCG_CHECK(c, cudaMemsetAsync(c->gcell_count.p, 0, (ncells + 4) * sizeof(int), st));
  if (nsrc > 0) {
    k_source_cells<false><<<nblocks(nsrc, 256), 256, 0, st>>>(c->pos.p, c->tag.p, nsrc, nlimg, lo, lc, im, recv, c->grid,
                                                              c->gcell_count.p, nullptr, nullptr, nullptr, nullptr);
    UCG_LAUNCHED(c);
  }
  BuildTrace trace(c);
  if ((rc = exclusive_scan(c, c->gcell_count.p, c->gcell_start.p, ncells + 4, nullptr))) return rc;
  trace.mark("ghost cell scan");
  c->nghost = nsrc;
  if ((rc = ucg_ensure_atom_capacity(c, (size_t)nlocal + nsrc, nlocal))) return rc;
  if (nsrc > 0) {
    UCG_CHECK(c, c->ghost_key.ensure(nsrc));
    UCG_CHECK(c, c->ghost_src.ensure(nsrc));
    UCG_CHECK(c, c->slot_of_src.ensure(nsrc));
    UCG_CHECK(c, c->ghost_code.ensure(nsrc));
    UCG_CHECK(c, cudaMemsetAsync(c->cell_cursor.p, 0, (ncells + 4) * sizeof(int), st));
    k_source_cells<true><<<nblocks(nsrc, 256), 256, 0, st>>>(c->pos.p, c->tag.p, nsrc, nlimg, lo, lc, im, recv, c->grid,
                                                             nullptr, c->gcell_start.p, c->cell_cursor.p,
                                                             c->ghost_key.p, c->ghost_src.p);
    UCG_LAUNCHED(c);
    k_sort_cells_kv<<<nblocks(ncells, 128), 128, 0, st>>>(c->ghost_key.p, c->ghost_src.p, c->gcell_start.p, ncells);
    UCG_LAUNCHED(c);
    UCG_CHECK(c, c->ghost_mask.ensure(nsrc));
    k_ghost_fill<true><<<nblocks(nsrc, 256), 256, 0, st>>>(c->pos.p, c->ts.p, c->ucgp.p, c->tag.p, c->mol.p, nlocal, nsrc,
                                                           nlimg, c->ghost_src.p, c->slot_of_src.p, lo, lc, im, recv,
                                                           c->ghost_code.p, c->mask.p, c->ghost_mask.p);
    UCG_LAUNCHED(c);
  }
  if (nlocal > 0) {
    if (c->neigh_stride == 0) {
      double vol = 1.0;
      for (int d = 0; d < 3; d++) vol *= (c->subhi[d] - c->sublo[d]);
      double est = 4.18879 * c->cutneighmax * c->cutneighmax * c->cutneighmax * nlocal / vol;
      int s = (int)(est * 1.35) + 24;
      c->neigh_stride = ((s + 15) / 16) * 16;
    }
    trace.mark("ghost bin+sort+fill");
    if ((rc = build_rows(c))) return rc;
    trace.mark("build_rows (+ overflow check)");
    UCG_CHECK(c, c->xhold.ensure(nlocal));
    UCG_CHECK(c, cudaMemcpyAsync(c->xhold.p, c->pos.p, (size_t)nlocal * sizeof(double4), cudaMemcpyDeviceToDevice, st));
  }
  UCG_CHECK(c, cudaMemsetAsync(c->d_maxdisp.p, 0, sizeof(unsigned long long), st));   // nobody has moved since this build
  c->maxdisp_valid = true;
  c->list_valid = true;
  c->nbuilds++;
  timer.stop();
  return 0;
}

// single-GPU rebuild: every image is local
extern "C" int ucgb200_neigh_build(ucgb200_ctx *c) {
  if (!c) return -1;
  if (c->halo.nranks > 1) return fail(c, "neigh_build: this context is one brick of several; drive the rebuild through "
                                         "migrate_* / neigh_build_local / halo_* / neigh_build_finish");
  int rc = ucgb200_neigh_build_local(c);
  if (rc) return rc;
  c->halo.nrecv = 0;
  return ucgb200_neigh_build_finish(c);
}

extern "C" int ucgb200_neigh_decide(ucgb200_ctx *c, int *rebuild) {
  if (!c || !rebuild) return -1;
  cudaSetDevice(c->device);
  if (!c->list_valid) { *rebuild = 1; return 0; }
  *rebuild = 0;
  UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p, 0, sizeof(int), c->stream));
  UCG_CHECK(c, cudaMemsetAsync(c->d_maxdisp.p, 0, sizeof(unsigned long long), c->stream));
  c->maxdisp_valid = c->halo.nranks == 1;   // across bricks the maximum must be all-reduced (comm.cu)
  if (c->nlocal > 0) {
    double triggersq = 0.25 * c->skin * c->skin;
    k_check_distance<<<nblocks(c->nlocal, 256), 256, 0, c->stream>>>(c->pos.p, c->xhold.p, c->nlocal, triggersq, c->d_flags.p, c->d_maxdisp.p);
    UCG_LAUNCHED(c);
  }
  UCG_CHECK(c, cudaMemcpyAsync(c->h_flags, c->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  *rebuild = c->h_flags[0] ? 1 : 0;
  return 0;
}

int ucg::ucg_check_distance_launch(ucgb200_ctx *c) {
  const double triggersq = 0.25 * c->skin * c->skin;
  k_check_distance<<<nblocks(c->nlocal, 256), 256, 0, c->stream>>>(c->pos.p, c->xhold.p, c->nlocal, triggersq, c->d_flags.p, c->d_maxdisp.p);
  UCG_LAUNCHED(c);
  return 0;
}

// Neighbor::decide when k_check_distance already ran inside the fused step tail (fixes.cu): only the
// 4-byte flag travels
int ucg_neigh_decide_prechecked(ucgb200_ctx *c, int *rebuild) {
  cudaSetDevice(c->device);
  if (!c->list_valid) { *rebuild = 1; return 0; }
  c->maxdisp_valid = c->halo.nranks == 1;
  UCG_CHECK(c, cudaMemcpyAsync(c->h_flags, c->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  *rebuild = c->h_flags[0] ? 1 : 0;
  return 0;
}

// comm->forward_comm() to self: refresh the local periodic images
extern "C" int ucgb200_ghosts_forward(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  auto &h = c->halo;
  if (h.nlimg == 0) return 0;
  ImageMap im = make_image_map(c);
  k_ghost_fill<false><<<nblocks(h.nlimg, 256), 256, 0, c->stream>>>(
      c->pos.p, c->ts.p, c->ucgp.p, c->tag.p, c->mol.p, c->nlocal, h.nlimg, h.nlimg, nullptr, c->slot_of_src.p,
      c->img_owner.p + h.nsend, c->img_code.p + h.nsend, im, nullptr, nullptr);
  UCG_LAUNCHED(c);
  return 0;
}

// ----------------------------------------------------------------- multi-GPU halo
extern "C" int ucgb200_halo_configure(ucgb200_ctx *c, int rank, int nranks, const int procgrid[3]) {
  if (!c || !procgrid || nranks < 1 || rank < 0 || rank >= nranks) return -1;
  if (procgrid[0] * procgrid[1] * procgrid[2] != nranks) return fail(c, "halo_configure: procgrid does not multiply to nranks");
  auto &h = c->halo;
  h.rank = rank; h.nranks = nranks;
  for (int d = 0; d < 3; d++) h.grid[d] = procgrid[d];
  h.coord[0] = rank % procgrid[0];
  h.coord[1] = (rank / procgrid[0]) % procgrid[1];
  h.coord[2] = rank / (procgrid[0] * procgrid[1]);
  for (int d = 0; d < 3; d++) {
    double w = c->prd[d] / procgrid[d];
    c->sublo[d] = c->boxlo[d] + h.coord[d] * w;
    c->subhi[d] = (h.coord[d] == procgrid[d] - 1) ? c->boxhi[d] : c->boxlo[d] + (h.coord[d] + 1) * w;
  }
  c->sub_set = true;
  h.send_counts.assign(nranks, 0);
  h.send_offsets.assign(nranks + 1, 0);
  h.recv_counts.assign(nranks, 0);
  h.mig_counts.assign(nranks, 0);
  c->list_valid = false;
  return 0;
}

extern "C" int ucgb200_halo_record_bytes(int *border, int *forward, int *migrate) {
  if (border) *border = (int)sizeof(BorderRec);
  if (forward) *forward = (int)sizeof(ForwardRec);
  if (migrate) *migrate = (int)sizeof(MigrateRec);
  return 0;
}

// comm->exchange(), part 1: wrap, find the sites that left this brick; counts[nranks]
extern "C" int ucgb200_migrate_prepare(ucgb200_ctx *c, int *counts) {
  if (!c || !counts) return -1;
  cudaSetDevice(c->device);
  auto &h = c->halo;
  const int nr = h.nranks, n = c->nlocal;
  cudaStream_t st = c->stream;
  int rc = rebuild_maps(c);
  if (rc) return rc;
  if ((rc = setup_grid(c))) return rc;
  BoxDev box = make_box(c);
  UCG_CHECK(c, c->img_counters.ensure(2 * (nr + 1) + 4));
  UCG_CHECK(c, cudaMemsetAsync(c->img_counters.p, 0, (2 * (nr + 1) + 4) * sizeof(int), st));
  UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p, 0, 8 * sizeof(int), st));
  UCG_CHECK(c, c->cell_of.ensure(n + 1));
  UCG_CHECK(c, c->order.ensure(n + 1));
  UCG_CHECK(c, c->mig_dest.ensure(n + 1));
  UCG_CHECK(c, c->mig_stay.ensure(n + 8));
  UCG_CHECK(c, c->mig_scan.ensure(n + 8));
  UCG_CHECK(c, c->cell_count.ensure(c->grid.ncells + 4));
  if (n > 0) {
    UCG_CHECK(c, cudaMemsetAsync(c->cell_count.p, 0, (c->grid.ncells + 4) * sizeof(int), st));
    k_wrap_count<<<nblocks(n, 256), 256, 0, st>>>(c->pos.p, n, box, c->grid, c->cell_of.p, c->cell_count.p, c->d_flags.p);
    UCG_LAUNCHED(c);
    k_migrate_classify<<<nblocks(n, 256), 256, 0, st>>>(c->pos.p, n, box, h.grid[0], h.grid[1], h.grid[2], h.rank,
                                                        c->mig_dest.p, c->mig_stay.p, c->img_counters.p);
    UCG_LAUNCHED(c);
  }
  std::vector<int> hc(nr, 0);
  UCG_CHECK(c, cudaMemcpyAsync(hc.data(), c->img_counters.p, nr * sizeof(int), cudaMemcpyDeviceToHost, st));
  if ((rc = read_flags(c))) return rc;
  if (c->h_flags[3]) { c->err = "atoms lost: position outside the periodic box by more than one period"; return UCGB200_ERR_LOST_ATOMS; }
  h.mig_counts = hc;
  for (int k = 0; k < nr; k++) counts[k] = hc[k];
  return 0;
}

// part 2: pack the leavers (grouped by destination rank, ascending) and compact the stayers
extern "C" int ucgb200_migrate_pack(ucgb200_ctx *c, void *d_sendbuf) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  auto &h = c->halo;
  const int nr = h.nranks, n = c->nlocal;
  cudaStream_t st = c->stream;
  std::vector<int> off(nr + 1, 0);
  for (int k = 0; k < nr; k++) off[k + 1] = off[k] + h.mig_counts[k];
  const int nleave = off[nr];
  int rc;
  if (nleave > 0) {
    if (!d_sendbuf) return fail(c, "migrate_pack: null send buffer");
    int *cursor = c->img_counters.p, *offsets = c->img_counters.p + (nr + 1);
    UCG_CHECK(c, cudaMemsetAsync(cursor, 0, (nr + 1) * sizeof(int), st));
    UCG_CHECK(c, cudaMemcpyAsync(offsets, off.data(), (nr + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    k_migrate_pack<<<nblocks(n, 256), 256, 0, st>>>(c->pos.p, c->vel.p, c->ts.p, c->mask.p, c->tag.p, c->mol.p, c->ucgp.p,
                                                    c->ucgml.p, n, h.rank, c->mig_dest.p, offsets, cursor,
                                                    (MigrateRec *)d_sendbuf);
    UCG_LAUNCHED(c);
    // stable compaction of the stayers
    if ((rc = exclusive_scan(c, c->mig_stay.p, c->mig_scan.p, n, nullptr))) return rc;
    k_stay_order<<<nblocks(n, 256), 256, 0, st>>>(c->mig_stay.p, c->mig_scan.p, n, c->order.p);
    UCG_LAUNCHED(c);
    if ((rc = permute_all(c, n - nleave))) return rc;
    c->nlocal = n - nleave;
    UCG_CHECK(c, cudaStreamSynchronize(st));
  }
  c->list_valid = false;
  return 0;
}

// part 3: append the arrivals
extern "C" int ucgb200_migrate_unpack(ucgb200_ctx *c, const void *d_recvbuf, int nrecv) {
  if (!c || nrecv < 0) return -1;
  cudaSetDevice(c->device);
  int rc;
  const int base = c->nlocal;
  if (nrecv > 0) {
    if (!d_recvbuf) return fail(c, "migrate_unpack: null receive buffer");
    if ((rc = ucg_ensure_atom_capacity(c, (size_t)base + nrecv + c->nghost, (size_t)base + nrecv))) return rc;
    k_migrate_unpack<<<nblocks(nrecv, 256), 256, 0, c->stream>>>(c->pos.p, c->vel.p, c->frc.p, c->scores.p, c->ts.p,
                                                                 c->mask.p, c->tag.p, c->mol.p, c->ucgp.p, c->ucgml.p,
                                                                 c->orig.p, base, nrecv, (const MigrateRec *)d_recvbuf);
    UCG_LAUNCHED(c);
    c->nlocal = base + nrecv;
  }
  // device order is the only order once sites migrate: downloads are identified by tag
  if (c->nlocal > 0) {
    k_iota2<<<nblocks(c->nlocal, 256), 256, 0, c->stream>>>(c->orig.p, c->nlocal);
    UCG_LAUNCHED(c);
  }
  c->list_valid = false;
  return 0;
}

extern "C" int ucgb200_halo_send_counts(ucgb200_ctx *c, int *counts) {
  if (!c || !counts) return -1;
  for (int k = 0; k < c->halo.nranks; k++) counts[k] = c->halo.send_counts[k];
  return 0;
}

extern "C" int ucgb200_halo_pack_border(ucgb200_ctx *c, void *d_sendbuf) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  auto &h = c->halo;
  if (h.nsend == 0) return 0;
  if (!d_sendbuf) return fail(c, "halo_pack_border: null buffer");
  ImageMap im = make_image_map(c);
  k_pack_border<<<nblocks(h.nsend, 256), 256, 0, c->stream>>>(c->pos.p, c->ts.p, c->tag.p, c->mol.p, c->ucgp.p, c->mask.p,
                                                             c->img_owner.p, c->img_code.p, h.nsend, im,
                                                             (BorderRec *)d_sendbuf);
  UCG_LAUNCHED(c);
  return 0;
}

// received border records (grouped by source rank, ascending) are kept until the next rebuild
extern "C" int ucgb200_halo_unpack_border(ucgb200_ctx *c, const void *d_recvbuf, const int *recv_counts) {
  if (!c || !recv_counts) return -1;
  cudaSetDevice(c->device);
  auto &h = c->halo;
  int total = 0;
  for (int k = 0; k < h.nranks; k++) { h.recv_counts[k] = recv_counts[k]; total += recv_counts[k]; }
  h.nrecv = total;
  if (total > 0) {
    if (!d_recvbuf) return fail(c, "halo_unpack_border: null buffer");
    UCG_CHECK(c, c->recv_border.ensure((size_t)total * sizeof(BorderRec)));
    UCG_CHECK(c, cudaMemcpyAsync(c->recv_border.p, d_recvbuf, (size_t)total * sizeof(BorderRec), cudaMemcpyDeviceToDevice, c->stream));
  }
  return 0;
}

extern "C" int ucgb200_halo_pack_forward(ucgb200_ctx *c, void *d_sendbuf) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  auto &h = c->halo;
  if (h.nsend == 0) return 0;
  if (!d_sendbuf) return fail(c, "halo_pack_forward: null buffer");
  ImageMap im = make_image_map(c);
  k_pack_forward<<<nblocks(h.nsend, 256), 256, 0, c->stream>>>(c->pos.p, c->ts.p, c->ucgp.p, c->img_owner.p, c->img_code.p,
                                                              h.nsend, im, (ForwardRec *)d_sendbuf);
  UCG_LAUNCHED(c);
  return 0;
}

// ---- interior / boundary rows of a multi-brick list
// 8 lanes per row: does the row hold a ghost that came from another brick (source index >= nlimg)?
__global__ void k_row_has_remote(const int *__restrict__ neigh, int stride, const int *__restrict__ numneigh, int nlocal,
                                 const int *__restrict__ ghost_src, int nlimg, int *__restrict__ flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, sub = t & 7;
  bool hit = false;
  if (i < nlocal) {
    const int *row = neigh + (size_t)i * stride;
    const int n = numneigh[i];
    for (int k = sub; k < n; k += 8) {      // any order: storage slots 0..n-1 hold the row's n entries (rowslot permutes inside 16-blocks)
      const int j = row[rowslot(k)] & UCG_NEIGHMASK;
      if (j >= nlocal && ghost_src[j - nlocal] >= nlimg) hit = true;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if (i < nlocal && sub == 0) flag[i] = ((m >> ((threadIdx.x & 31) & ~7)) & 0xffu) ? 1 : 0;
}
__global__ void k_fill_site_list(const int *__restrict__ flag, const int *__restrict__ scan, int nlocal, const int *__restrict__ nboundary,
                                 int *__restrict__ list, int *__restrict__ part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) part[0] = nlocal - *nboundary;
  if (i >= nlocal) return;
  const int ninterior = nlocal - *nboundary;
  if (flag[i]) list[ninterior + scan[i]] = i;          // scan = boundary sites before i
  else list[i - scan[i]] = i;
}
int ucg_classify_rows(ucgb200_ctx *c) {
  cudaSetDevice(c->device);
  c->parts_valid = false;
  const int nlocal = c->nlocal;
  if (!c->list_valid || nlocal == 0) return 0;
  UCG_CHECK(c, c->row_flag.ensure(nlocal + 1));
  UCG_CHECK(c, c->row_scan.ensure(nlocal + 1));
  UCG_CHECK(c, c->site_list.ensure(nlocal + 1));
  UCG_CHECK(c, c->d_part.ensure(4));
  if (c->nghost == 0) {
    UCG_CHECK(c, cudaMemsetAsync(c->row_flag.p, 0, nlocal * sizeof(int), c->stream));
  } else {
    k_row_has_remote<<<nblocks((long long)nlocal * 8, 256), 256, 0, c->stream>>>(c->neigh.p, c->neigh_stride, c->numneigh.p, nlocal,
                                                                                  c->ghost_src.p, c->halo.nlimg, c->row_flag.p);
    UCG_LAUNCHED(c);
  }
  int rc = exclusive_scan(c, c->row_flag.p, c->row_scan.p, nlocal, c->d_part.p + 1);
  if (rc) return rc;
  k_fill_site_list<<<nblocks(nlocal, 256), 256, 0, c->stream>>>(c->row_flag.p, c->row_scan.p, nlocal, c->d_part.p + 1, c->site_list.p, c->d_part.p);
  UCG_LAUNCHED(c);
  c->parts_valid = true;
  return 0;
}

int ucg_halo_push_forward(ucgb200_ctx *c, const UcgPushTargets &t) {
  cudaSetDevice(c->device);
  auto &h = c->halo;
  ImageMap im = make_image_map(c);
  const int nb = std::max(1, nblocks(h.nsend, 256));     // at least one CTA: the control words go out even with an empty list
  k_push_forward<<<nb, 256, 0, c->stream>>>(c->pos.p, c->ts.p, c->ucgp.p, c->img_owner.p, c->img_code.p, h.nsend, im, t);
  UCG_LAUNCHED(c);
  return 0;
}
int ucg_halo_wait_reduce(ucgb200_ctx *c, const UcgP2PCtl *ctl_mine, int nranks, int self, int seq) {
  cudaSetDevice(c->device);
  k_wait_reduce<<<1, 32, 0, c->stream>>>(ctl_mine, nranks, self, seq, c->d_flags.p, c->d_maxdisp.p, c->d_err.p);
  UCG_LAUNCHED(c);
  return 0;
}

extern "C" int ucgb200_halo_unpack_forward(ucgb200_ctx *c, const void *d_recvbuf) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  auto &h = c->halo;
  if (h.nrecv == 0) return 0;
  if (!d_recvbuf) return fail(c, "halo_unpack_forward: null buffer");
  k_unpack_forward<<<nblocks(h.nrecv, 256), 256, 0, c->stream>>>(c->pos.p, c->ts.p, c->ucgp.p, c->nlocal,
                                                                (const ForwardRec *)d_recvbuf, c->slot_of_src.p, h.nlimg,
                                                                h.nrecv);
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
  std::vector<int> nn(std::max(nlocal, 1));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  if (nlocal) UCG_CHECK(c, cudaMemcpy(nn.data(), c->numneigh.p, nlocal * sizeof(int), cudaMemcpyDeviceToHost));
  long long t = 0;
  for (int i = 0; i < nlocal; i++) t += nn[i];
  if (nlocal_out) *nlocal_out = nlocal;
  if (!neigh_tags) { *total = t; return 0; }
  if (*total < t) return fail(c, "neigh_download: capacity too small");
  *total = t;
  std::vector<int> tags(std::max(nall, 1)), rows((size_t)nlocal * c->neigh_stride + 1), gcode(std::max(c->nghost, 1));
  if (nall) UCG_CHECK(c, cudaMemcpy(tags.data(), c->tag.p, nall * sizeof(int), cudaMemcpyDeviceToHost));
  if (nlocal) UCG_CHECK(c, cudaMemcpy(rows.data(), c->neigh.p, (size_t)nlocal * c->neigh_stride * sizeof(int), cudaMemcpyDeviceToHost));
  if (c->nghost) UCG_CHECK(c, cudaMemcpy(gcode.data(), c->ghost_code.p, c->nghost * sizeof(int), cudaMemcpyDeviceToHost));
  long long off = 0;
  for (int i = 0; i < nlocal; i++) {
    if (tag_i) tag_i[i] = tags[i];
    if (numneigh) numneigh[i] = nn[i];
    if (offsets) offsets[i] = off;
    for (int k = 0; k < nn[i]; k++) {
      int j = rows[(size_t)i * c->neigh_stride + rowslot(k)] & UCG_NEIGHMASK;
      neigh_tags[off + k] = tags[j];
      // image code of the neighbor: 0 = owned site, 1..26 local periodic image, 32+ = from another brick
      if (neigh_shift) neigh_shift[off + k] = j < nlocal ? 0 : gcode[j - nlocal];
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
