// cluster_switch.cu — FixClusterSwitch (UCG/fix_cluster_switch.cpp) for sm_100a.
//
// Every switchFreq steps the reference (pre_exchange :464-481) rebuilds the full list, labels the
// molecules connected to mol_seed (check_cluster :537-731) and Monte-Carlo flips the ON/OFF atom
// types of the molecules outside that cluster (attempt_switch :733-839).  Device version:
//
//   check_cluster   labels start as in :556-582 (label[m] = m for molecules with group atoms,
//                   label[m - mol_offset] = m for switchable m).  The reference then sweeps the
//                   list Gauss-Seidel fashion, merging on every contact the labels of both
//                   molecules AND of their offset partners to the minimum (:640-665).  Because a
//                   molecule and its partner start with one label and are always merged together,
//                   the fixed point is: label = min initial label over the connected component of
//                   the graph {contacts} U {(m, m - mol_offset)}.  That fixed point is computed
//                   here by k_cs_edges (one sweep over the rows -> contact edges between molecules)
//                   and min-label hooking + pointer jumping over the edges (k_cs_hook / k_cs_jump),
//                   O(log) rounds instead of O(diameter) sweeps.  Then :695-708 per molecule.
//   attempt_switch  confirm_molecule (:841-893) becomes one atomic tally per (atom, switch type);
//                   switch_flag (:896-921) draws RanPark uniforms in ascending molecule id (the
//                   std::map order of :741-775): the k-th value of that Lehmer stream is
//                   16807^k * seed mod (2^31-1), so every molecule computes its own draw from its
//                   rank among the drawing molecules (an exclusive scan) — the stream, and with it
//                   every accept decision, is the reference's bit for bit.  Types are then flipped
//                   per atom (:800-822) and mol_state per molecule.
// Decks in which a non-switchable molecule X has no switchable partner X + mol_offset make the
// reference read labels out of bounds / merge with -1 (:648-651); here such partners are ignored.
// A molecule with more switchable atoms than mol_seed has only its first nSwitchPerMol atoms (in
// local index order) recorded by the reference (:862-868); here all of them are tallied and flipped.
#include "ucg_internal.cuh"

#include <algorithm>

using namespace ucg;

namespace {

constexpr int CS_MAXSW = 8;       // switch types held in kernel arguments
constexpr unsigned long long RP_IA = 16807ULL, RP_IM = 2147483647ULL;   // [stock] RanPark

struct CsTypes {
  int n;
  int on[CS_MAXSW], off[CS_MAXSW];
};

// ---- setup (constructor :98-170): largest molecule id, switchable atoms of mol_seed and in total
__global__ void k_cs_scan_atoms(const int *__restrict__ ts, const int *__restrict__ mask, const int *__restrict__ mol,
                                int nlocal, int groupbit, CsTypes st, int mol_seed, int *__restrict__ out /*[3]*/) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal || !(mask[i] & groupbit)) return;
  const int t = ts[i] & 0xffff, m = mol[i];
  atomicMax(&out[0], m);
  int hits = 0;
  for (int k = 0; k < st.n; k++) hits += (t == st.on[k] || t == st.off[k]);
  if (hits) {
    atomicAdd(&out[1], hits);
    if (m == mol_seed) atomicAdd(&out[2], hits);
  }
}

// mol_state / mol_restrict from the atom types (:133-160).  The reference takes, per molecule, the
// first matching (atom, k) in local index order; for molecules whose switchable atoms agree that is
// any of them — here the state is the MAX over the atoms (ON wins in a mixed molecule), which is also
// what the reference's MPI_MAX does across ranks.
__global__ void k_cs_init_state(const int *__restrict__ ts, const int *__restrict__ mask, const int *__restrict__ mol,
                                int nlocal, int groupbit, CsTypes st, int *__restrict__ mol_state) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal || !(mask[i] & groupbit)) return;
  const int t = ts[i] & 0xffff;
  int s = -1;
  for (int k = 0; k < st.n; k++) {
    if (t == st.on[k]) s = max(s, 1);
    else if (t == st.off[k]) s = max(s, 0);
  }
  if (s >= 0) atomicMax(&mol_state[mol[i]], s);
}
__global__ void k_cs_init_restrict(const int *__restrict__ mol_state, int *__restrict__ mol_restrict, int nm, int mol_seed,
                                   int mol_offset) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nm) return;
  mol_restrict[m] = (mol_state[m] >= 0 && m != mol_seed && m != mol_seed - mol_offset) ? 1 : -1;
}

// ---- check_cluster
__global__ void k_cs_fill(int *__restrict__ a, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_cs_label_self(const int *__restrict__ mask, const int *__restrict__ mol, int nlocal, int groupbit,
                                int *__restrict__ label) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nlocal && (mask[i] & groupbit)) label[mol[i]] = mol[i];                      // :567-573
}
__global__ void k_cs_label_partner(const int *__restrict__ mask, const int *__restrict__ mol, int nlocal, int groupbit,
                                   const int *__restrict__ mol_state, int mol_offset, int nm, int *__restrict__ label) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal || !(mask[i] & groupbit)) return;
  const int m = mol[i], x = m - mol_offset;
  if (mol_state[m] >= 0 && x >= 0 && x < nm) label[x] = m;                             // :576-583
}
__global__ void k_cs_ghost_mask(const int *__restrict__ mask, int nlimg, const int *__restrict__ owner,
                                const int *__restrict__ slot_of_src, int *__restrict__ gmask) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nlimg) gmask[slot_of_src[k]] = mask[owner[k]];
}

// contact edges between molecules (:600-640): COUNT = true only counts
template <bool COUNT, int LPA, int BS>
__global__ void __launch_bounds__(BS) k_cs_edges(const double4 *__restrict__ pos, const int *__restrict__ ts,
                                                 const int *__restrict__ mask, const int *__restrict__ gmask,
                                                 const int *__restrict__ mol, int nlocal, const int *__restrict__ neigh,
                                                 int stride, const int *__restrict__ numneigh, int groupbit,
                                                 const unsigned char *__restrict__ contact, int ntp1, double cutsq,
                                                 int *__restrict__ counter, int2 *__restrict__ edges, int cap) {
  const int gid = (blockIdx.x * BS + threadIdx.x) / LPA;
  const int sub = threadIdx.x % LPA;
  if (gid >= nlocal) return;
  const int i = gid;
  if (!(mask[i] & groupbit)) return;
  const double4 ri = pos[i];
  const int ti = ts[i] & 0xffff, mi = mol[i];
  const int jnum = numneigh[i];
  const int *row = neigh + (size_t)i * stride;
  for (int jj = sub; jj < jnum; jj += LPA) {
    const int j = row[rowslot(jj)] & UCG_NEIGHMASK;
    const int mj = mol[j];
    if (mj == mi) continue;
    const int mk = j < nlocal ? mask[j] : gmask[j - nlocal];
    if (!(mk & groupbit)) continue;
    const int tj = ts[j] & 0xffff;
    if (!contact[ti * ntp1 + tj]) continue;
    const double4 rj = pos[j];
    const double rsq = rsq_exact(ri.x - rj.x, ri.y - rj.y, ri.z - rj.z);
    if (rsq < cutsq) {
      const int slot = atomicAdd(counter, 1);
      if (!COUNT && slot < cap) edges[slot] = make_int2(mi, mj);
    }
  }
}

__device__ __forceinline__ bool cs_merge(int *label, int a, int b) {
  const int la = label[a], lb = label[b];
  if (la == lb || la < 0 || lb < 0) return false;
  const int lo = min(la, lb), hi = max(la, lb);
  atomicMin(&label[hi], lo);   // hook the larger root under the smaller one
  atomicMin(&label[a], lo);
  atomicMin(&label[b], lo);
  return true;
}
__global__ void k_cs_hook(const int2 *__restrict__ edges, int nedges, int *__restrict__ label, int *__restrict__ changed) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nedges) return;
  const int2 ed = edges[e];
  if (cs_merge(label, ed.x, ed.y)) *changed = 1;
}
// pointer jumping + the (m, m - mol_offset) ties
__global__ void k_cs_jump(int *__restrict__ label, const int *__restrict__ mol_state, int nm, int mol_offset,
                          int *__restrict__ changed) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nm) return;
  bool ch = false;
  const int x = m - mol_offset;
  if (mol_state[m] >= 0 && x >= 0 && x < nm) ch = cs_merge(label, m, x);
  int l = label[m];
  if (l >= 0) {
    const int ll = label[l];
    if (ll >= 0 && ll < l) { atomicMin(&label[m], ll); ch = true; }
  }
  if (ch) *changed = 1;
}
// :695-708
__global__ void k_cs_restrict(const int *__restrict__ label, int nm, int mol_seed, int *__restrict__ mol_state,
                              int *__restrict__ mol_restrict, int *__restrict__ ncluster) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nm) return;
  const int cid = label[mol_seed];
  const int l = label[m];
  if (l == -1) return;
  if (mol_state[m] == 0 || mol_state[m] == 1) {
    if (l == cid) { mol_restrict[m] = -1; mol_state[m] = 1; }
    else mol_restrict[m] = 1;
  }
  if (l == cid) atomicAdd(ncluster, 1);
}

// ---- attempt_switch
__global__ void k_cs_tally(const int *__restrict__ ts, const int *__restrict__ mask, const int *__restrict__ mol, int nlocal,
                           int groupbit, CsTypes st, int *__restrict__ present, int *__restrict__ sum_state) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const int m = mol[i];
  if (mask[i] & groupbit) present[m] = 1;                                              // the std::map of :741-749
  const int t = ts[i] & 0xffff;
  int s = 0;
  for (int k = 0; k < st.n; k++) s += (t == st.on[k]) - (t == st.off[k]);            // confirm_molecule :852-886
  if (s) atomicAdd(&sum_state[m], s);
}
__global__ void k_cs_draws(const int *__restrict__ present, const int *__restrict__ sum_state,
                           const int *__restrict__ mol_restrict, int nm, double decision_buffer, int *__restrict__ draws) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nm) return;
  const double s = (double)sum_state[m];
  const bool confirm = s < -decision_buffer || s > decision_buffer;                    // :889-891
  draws[m] = (present[m] && mol_restrict[m] == 1 && confirm) ? 1 : 0;
}
__device__ __forceinline__ unsigned long long rp_powmod(unsigned long long k) {
  unsigned long long r = 1, b = RP_IA;
  while (k) {
    if (k & 1ULL) r = (r * b) % RP_IM;
    b = (b * b) % RP_IM;
    k >>= 1;
  }
  return r;
}
__global__ void k_cs_accept(const int *__restrict__ draws, const int *__restrict__ rank, const int *__restrict__ mol_state,
                            int nm, unsigned long long seed0, unsigned long long ndrawn, double prob_on, double prob_off,
                            int *__restrict__ accept) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nm) return;
  int a = -1;
  if (draws[m]) {
    // RanPark::uniform() number (ndrawn + rank + 1) of the stream started at seed0
    const unsigned long long s = (seed0 % RP_IM) * rp_powmod(ndrawn + (unsigned long long)rank[m] + 1ULL) % RP_IM;
    const double u = (1.0 / 2147483647.0) * (double)s;
    const double check = mol_state[m] == 0 ? prob_on : prob_off;                      // switch_flag :903-916
    a = u < check ? 1 : 0;
  }
  accept[m] = a;
}
// gather_statistics (:935-968): {attempts, attemptsON, attemptsOFF, success, successON, successOFF}
__global__ void k_cs_stats(const int *__restrict__ mol_restrict, const int *__restrict__ mol_state,
                           const int *__restrict__ accept, int nm, int *__restrict__ stats) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nm || mol_restrict[m] != 1) return;
  atomicAdd(&stats[0], 1);
  const int s = mol_state[m], ok = accept[m] == 1;
  if (s == 0) { atomicAdd(&stats[1], 1); if (ok) { atomicAdd(&stats[3], 1); atomicAdd(&stats[4], 1); } }
  else if (s == 1) { atomicAdd(&stats[2], 1); if (ok) { atomicAdd(&stats[3], 1); atomicAdd(&stats[5], 1); } }
}
__global__ void k_cs_flip_types(int *__restrict__ ts, const int *__restrict__ mol, int nlocal, CsTypes st,
                                const int *__restrict__ accept, const int *__restrict__ mol_state) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const int m = mol[i];
  if (accept[m] != 1) return;
  const int w = ts[i];
  int t = w & 0xffff;
  const int s = mol_state[m];
  for (int k = 0; k < st.n; k++) {                                                     // :804-817, k in order
    if (s == 0 && t == st.off[k]) t = st.on[k];
    else if (s == 1 && t == st.on[k]) t = st.off[k];
  }
  ts[i] = (w & ~0xffff) | t;
}
__global__ void k_cs_flip_state(int *__restrict__ mol_state, const int *__restrict__ accept, int nm) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nm || accept[m] != 1) return;
  if (mol_state[m] == 0) mol_state[m] = 1;
  else if (mol_state[m] == 1) mol_state[m] = 0;
}

CsTypes cs_types(const ucgb200_ctx *c) {
  CsTypes st{};
  st.n = c->cluster.n_switch;
  for (int k = 0; k < st.n; k++) { st.on[k] = c->cluster.type_on[k]; st.off[k] = c->cluster.type_off[k]; }
  return st;
}

}  // namespace

extern "C" int ucgb200_cluster_configure(ucgb200_ctx *c, int mol_seed, int mol_offset, double cutoff, int seed,
                                         double prob_on, int n_switch_types, const int *type_on, const int *type_off,
                                         int n_contacts, const int *contact_pairs, int ntypes, int groupbit) {
  if (!c || !type_on || !type_off || (n_contacts > 0 && !contact_pairs)) return -1;
  if (n_switch_types < 1 || n_switch_types > CS_MAXSW) return fail(c, "Incorrect number of atom switching types (fix cluster_switch)");
  if (prob_on > 1.0) return fail(c, "Incorrect probability in rates.txt files (fix cluster_switch)");
  if (seed <= 0) return fail(c, "Invalid seed for Park random # generator");
  if (c->nlocal <= 0) return fail(c, "fix cluster_switch: upload the atoms first (the constructor scans their molecule ids and types)");
  cudaSetDevice(c->device);
  auto &k = c->cluster;
  k.mol_seed = mol_seed; k.mol_offset = mol_offset; k.cutoff = cutoff; k.seed = seed; k.ndrawn = 0;
  k.prob_on = prob_on; k.prob_off = 1.0 - prob_on;
  k.n_switch = n_switch_types; k.groupbit = groupbit ? groupbit : 1; k.ntypes = ntypes;
  k.type_on.assign(type_on, type_on + n_switch_types);
  k.type_off.assign(type_off, type_off + n_switch_types);
  std::vector<unsigned char> cm((size_t)(ntypes + 1) * (ntypes + 1), 0);
  for (int p = 0; p < n_contacts; p++) {
    const int a = contact_pairs[2 * p], b = contact_pairs[2 * p + 1];
    if (a >= 1 && a <= ntypes && b >= 1 && b <= ntypes) cm[(size_t)a * (ntypes + 1) + b] = 1;
  }
  UCG_CHECK(c, k.d_contact.ensure(cm.size()));
  UCG_CHECK(c, cudaMemcpy(k.d_contact.p, cm.data(), cm.size(), cudaMemcpyHostToDevice));
  for (int i = 0; i < 7; i++) k.stats[i] = 0.0;
  // constructor scan (:98-127)
  UCG_CHECK(c, k.d_scratch.ensure(16));
  int h3[3] = {-1, 0, 0};
  UCG_CHECK(c, cudaMemcpyAsync(k.d_scratch.p, h3, sizeof(h3), cudaMemcpyHostToDevice, c->stream));
  const CsTypes st = cs_types(c);
  k_cs_scan_atoms<<<nblocks(c->nlocal, 256), 256, 0, c->stream>>>(c->ts.p, c->mask.p, c->mol.p, c->nlocal, k.groupbit, st,
                                                                  mol_seed, k.d_scratch.p);
  UCG_LAUNCHED(c);
  {  // MPI_Allreduce of :109-111
    int rc2;
    if ((rc2 = ucg_mb_allreduce_int(c, k.d_scratch.p, 1, 1))) return rc2;
    if ((rc2 = ucg_mb_allreduce_int(c, k.d_scratch.p + 1, 2, 0))) return rc2;
  }
  UCG_CHECK(c, cudaMemcpyAsync(h3, k.d_scratch.p, sizeof(h3), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  if (h3[0] < 0) return fail(c, "Selected group does not have any mols (fix cluster_switch)");
  if (h3[2] <= 0) return fail(c, "fix cluster_switch: molecule mol_seed has no switchable atoms");
  if (mol_seed > h3[0] || mol_seed - mol_offset < 0 || mol_seed - mol_offset > h3[0])
    return fail(c, "fix cluster_switch: mol_seed / mol_seed - mol_offset outside the molecule id range");
  k.max_mol = h3[0];
  k.n_switch_per_mol = h3[2];
  k.nmol = h3[1] / h3[2];
  const int nm = k.max_mol + 1;
  UCG_CHECK(c, k.d_state.ensure(nm)); UCG_CHECK(c, k.d_restrict.ensure(nm)); UCG_CHECK(c, k.d_label.ensure(nm));
  UCG_CHECK(c, k.d_accept.ensure(nm)); UCG_CHECK(c, k.d_present.ensure(nm)); UCG_CHECK(c, k.d_sum.ensure(nm));
  UCG_CHECK(c, k.d_draws.ensure(nm)); UCG_CHECK(c, k.d_rank.ensure(nm + 8));
  k_cs_fill<<<nblocks(nm, 256), 256, 0, c->stream>>>(k.d_state.p, nm, -1); UCG_LAUNCHED(c);
  k_cs_fill<<<nblocks(nm, 256), 256, 0, c->stream>>>(k.d_accept.p, nm, -1); UCG_LAUNCHED(c);
  k_cs_fill<<<nblocks(nm, 256), 256, 0, c->stream>>>(k.d_label.p, nm, -1); UCG_LAUNCHED(c);
  k_cs_init_state<<<nblocks(c->nlocal, 256), 256, 0, c->stream>>>(c->ts.p, c->mask.p, c->mol.p, c->nlocal, k.groupbit, st, k.d_state.p);
  UCG_LAUNCHED(c);
  { int rc2; if ((rc2 = ucg_mb_allreduce_int(c, k.d_state.p, nm, 1))) return rc2; }   // :160-161
  k_cs_init_restrict<<<nblocks(nm, 256), 256, 0, c->stream>>>(k.d_state.p, k.d_restrict.p, nm, mol_seed, mol_offset);
  UCG_LAUNCHED(c);
  k.next_reneighbor = c->ntimestep + 1;   // :71
  k.set = true;
  return 0;
}

extern "C" int ucgb200_cluster_check(ucgb200_ctx *c, int *n_cluster) {
  if (!c) return -1;
  auto &k = c->cluster;
  if (!k.set) return fail(c, "fix cluster_switch: not configured");
  if (!c->list_valid) return fail(c, "fix cluster_switch: neighbor list not built");
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  const int nm = k.max_mol + 1, nlocal = c->nlocal;
  // initial labels (:556-586)
  k_cs_fill<<<nblocks(nm, 256), 256, 0, st>>>(k.d_label.p, nm, -1); UCG_LAUNCHED(c);
  const int preset[2] = {k.mol_seed, k.mol_seed};
  UCG_CHECK(c, cudaMemcpyAsync(k.d_label.p + k.mol_seed, &preset[0], sizeof(int), cudaMemcpyHostToDevice, st));
  UCG_CHECK(c, cudaMemcpyAsync(k.d_label.p + (k.mol_seed - k.mol_offset), &preset[1], sizeof(int), cudaMemcpyHostToDevice, st));
  k_cs_label_self<<<nblocks(nlocal, 256), 256, 0, st>>>(c->mask.p, c->mol.p, nlocal, k.groupbit, k.d_label.p); UCG_LAUNCHED(c);
  k_cs_label_partner<<<nblocks(nlocal, 256), 256, 0, st>>>(c->mask.p, c->mol.p, nlocal, k.groupbit, k.d_state.p, k.mol_offset, nm,
                                                           k.d_label.p);
  UCG_LAUNCHED(c);
  { int rc2; if ((rc2 = ucg_mb_allreduce_int(c, k.d_label.p, nm, 1))) return rc2; }   // :586, MPI_MAX since labels start at -1
  // group bits of the ghosts: local periodic images from their owners, ghosts owned by other bricks from the border
  // records of the last rebuild (neighbor.cu, BorderRec::mask)
  UCG_CHECK(c, k.d_gmask.ensure(std::max(c->nghost, 1)));
  if (c->halo.nranks > 1 && c->nghost > 0) {
    if (c->ghost_mask.cap < (size_t)c->nghost) return fail(c, "cluster_check: ghost group bits missing (no rebuild since configure)");
    UCG_CHECK(c, cudaMemcpyAsync(k.d_gmask.p, c->ghost_mask.p, (size_t)c->nghost * sizeof(int), cudaMemcpyDeviceToDevice, st));
  }
  if (c->halo.nlimg) {
    k_cs_ghost_mask<<<nblocks(c->halo.nlimg, 256), 256, 0, st>>>(c->mask.p, c->halo.nlimg, c->img_owner.p + c->halo.nsend,
                                                                 c->slot_of_src.p, k.d_gmask.p);
    UCG_LAUNCHED(c);
  }
  // contact edges: count, then fill
  constexpr int LPA = 8, BS = 256;
  const int nblk = nblocks((long long)nlocal * LPA, BS);
  const double cutsq = k.cutoff * k.cutoff;
  int *counter = k.d_scratch.p;
  UCG_CHECK(c, cudaMemsetAsync(counter, 0, 4 * sizeof(int), st));
  k_cs_edges<true, LPA, BS><<<nblk, BS, 0, st>>>(c->pos.p, c->ts.p, c->mask.p, k.d_gmask.p, c->mol.p, nlocal, c->neigh.p,
                                                 c->neigh_stride, c->numneigh.p, k.groupbit, k.d_contact.p, k.ntypes + 1, cutsq,
                                                 counter, nullptr, 0);
  UCG_LAUNCHED(c);
  int nedges = 0;
  UCG_CHECK(c, cudaMemcpyAsync(&nedges, counter, sizeof(int), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  UCG_CHECK(c, k.d_edges.ensure(std::max(nedges, 1)));
  if (nedges) {
    UCG_CHECK(c, cudaMemsetAsync(counter, 0, sizeof(int), st));
    k_cs_edges<false, LPA, BS><<<nblk, BS, 0, st>>>(c->pos.p, c->ts.p, c->mask.p, k.d_gmask.p, c->mol.p, nlocal, c->neigh.p,
                                                    c->neigh_stride, c->numneigh.p, k.groupbit, k.d_contact.p, k.ntypes + 1,
                                                    cutsq, counter, k.d_edges.p, nedges);
    UCG_LAUNCHED(c);
  }
  // min-label hooking + pointer jumping until nothing changes
  int *changed = k.d_scratch.p + 1;
  k.rounds = 0;
  for (;;) {
    UCG_CHECK(c, cudaMemsetAsync(changed, 0, sizeof(int), st));
    for (int r = 0; r < 4; r++) {
      if (nedges) { k_cs_hook<<<nblocks(nedges, 256), 256, 0, st>>>(k.d_edges.p, nedges, k.d_label.p, changed); UCG_LAUNCHED(c); }
      k_cs_jump<<<nblocks(nm, 256), 256, 0, st>>>(k.d_label.p, k.d_state.p, nm, k.mol_offset, changed); UCG_LAUNCHED(c);
      k.rounds++;
      // :682-683: every brick's labels to the global minimum, "anychange" to the maximum
      int rc2;
      if ((rc2 = ucg_mb_allreduce_int(c, k.d_label.p, nm, 2))) return rc2;
    }
    { int rc2; if ((rc2 = ucg_mb_allreduce_int(c, changed, 1, 1))) return rc2; }
    int h = 0;
    UCG_CHECK(c, cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
    UCG_CHECK(c, cudaStreamSynchronize(st));
    if (!h) break;
    if (k.rounds > 4096) return fail(c, "fix cluster_switch: cluster labelling did not converge");
  }
  int *ncl = k.d_scratch.p + 2;
  UCG_CHECK(c, cudaMemsetAsync(ncl, 0, sizeof(int), st));
  k_cs_restrict<<<nblocks(nm, 256), 256, 0, st>>>(k.d_label.p, nm, k.mol_seed, k.d_state.p, k.d_restrict.p, ncl); UCG_LAUNCHED(c);
  int hn = 0;
  UCG_CHECK(c, cudaMemcpyAsync(&hn, ncl, sizeof(int), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  k.stats[6] = (double)hn;
  k.nedges = nedges;
  if (n_cluster) *n_cluster = hn;
  return 0;
}

extern "C" int ucgb200_cluster_switch(ucgb200_ctx *c, int *n_attempts, int *n_success) {
  if (!c) return -1;
  auto &k = c->cluster;
  if (!k.set) return fail(c, "fix cluster_switch: not configured");
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  const int nm = k.max_mol + 1, nlocal = c->nlocal;
  const CsTypes ty = cs_types(c);
  int rc0;
  UCG_CHECK(c, cudaMemsetAsync(k.d_present.p, 0, nm * sizeof(int), st));
  UCG_CHECK(c, cudaMemsetAsync(k.d_sum.p, 0, nm * sizeof(int), st));
  k_cs_tally<<<nblocks(nlocal, 256), 256, 0, st>>>(c->ts.p, c->mask.p, c->mol.p, nlocal, k.groupbit, ty, k.d_present.p, k.d_sum.p);
  UCG_LAUNCHED(c);
  // every brick then holds the same per-molecule tallies and takes the same decisions from the same RanPark
  // stream (the reference draws rank-local streams with identical seeds, Q22: there the outcome depends on
  // the decomposition; here it does not)
  if ((rc0 = ucg_mb_allreduce_int(c, k.d_present.p, nm, 1))) return rc0;
  if ((rc0 = ucg_mb_allreduce_int(c, k.d_sum.p, nm, 0))) return rc0;
  const double decision_buffer = (double)k.n_switch_per_mol / 2.0 - 1.0 + 0.01;        // :848
  k_cs_draws<<<nblocks(nm, 256), 256, 0, st>>>(k.d_present.p, k.d_sum.p, k.d_restrict.p, nm, decision_buffer, k.d_draws.p);
  UCG_LAUNCHED(c);
  int *total = k.d_scratch.p + 3;
  int rc = exclusive_scan(c, k.d_draws.p, k.d_rank.p, nm, total);
  if (rc) return rc;
  k_cs_accept<<<nblocks(nm, 256), 256, 0, st>>>(k.d_draws.p, k.d_rank.p, k.d_state.p, nm, (unsigned long long)k.seed, k.ndrawn,
                                                k.prob_on, k.prob_off, k.d_accept.p);
  UCG_LAUNCHED(c);
  int *dstats = k.d_scratch.p + 8;
  UCG_CHECK(c, cudaMemsetAsync(dstats, 0, 6 * sizeof(int), st));
  k_cs_stats<<<nblocks(nm, 256), 256, 0, st>>>(k.d_restrict.p, k.d_state.p, k.d_accept.p, nm, dstats); UCG_LAUNCHED(c);
  k_cs_flip_types<<<nblocks(nlocal, 256), 256, 0, st>>>(c->ts.p, c->mol.p, nlocal, ty, k.d_accept.p, k.d_state.p); UCG_LAUNCHED(c);
  k_cs_flip_state<<<nblocks(nm, 256), 256, 0, st>>>(k.d_state.p, k.d_accept.p, nm); UCG_LAUNCHED(c);
  int hs[6], ht = 0;
  UCG_CHECK(c, cudaMemcpyAsync(hs, dstats, sizeof(hs), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaMemcpyAsync(&ht, total, sizeof(int), cudaMemcpyDeviceToHost, st));
  UCG_CHECK(c, cudaStreamSynchronize(st));
  k.ndrawn += (unsigned long long)ht;
  // compute_vector order (:923-933): attempts, success, attemptsON, attemptsOFF, successON, successOFF
  k.stats[0] += hs[0]; k.stats[1] += hs[3]; k.stats[2] += hs[1]; k.stats[3] += hs[2]; k.stats[4] += hs[4]; k.stats[5] += hs[5];
  if (n_attempts) *n_attempts = hs[0];
  if (n_success) *n_success = hs[3];
  // comm->forward_comm(this) (:825): the changed types reach the ghosts
  if ((rc = ucgb200_ghosts_forward(c))) return rc;
  return 0;
}

// FixClusterSwitch::compute_vector (:923-933) and the per-molecule arrays (logs :711-727, tests)
extern "C" int ucgb200_cluster_next_reneighbor(ucgb200_ctx *c, long long step) {
  if (!c || !c->cluster.set) return -1;
  c->cluster.next_reneighbor = step;
  return 0;
}

extern "C" int ucgb200_cluster_stats(ucgb200_ctx *c, double out[8]) {
  if (!c || !out) return -1;
  for (int i = 0; i < 7; i++) out[i] = c->cluster.stats[i];
  out[7] = (double)c->cluster.rounds;
  return 0;
}
extern "C" int ucgb200_cluster_get(ucgb200_ctx *c, int cap, int *mol_cluster, int *mol_state, int *mol_restrict,
                                   int *mol_accept, int *max_mol) {
  if (!c) return -1;
  auto &k = c->cluster;
  if (!k.set) return fail(c, "fix cluster_switch: not configured");
  if (max_mol) *max_mol = k.max_mol;
  const int nm = k.max_mol + 1;
  if (cap < nm) return (mol_cluster || mol_state || mol_restrict || mol_accept) ? fail(c, "cluster_get: buffers too small") : 0;
  cudaSetDevice(c->device);
  const int *src[4] = {k.d_label.p, k.d_state.p, k.d_restrict.p, k.d_accept.p};
  int *dst[4] = {mol_cluster, mol_state, mol_restrict, mol_accept};
  for (int a = 0; a < 4; a++)
    if (dst[a]) UCG_CHECK(c, cudaMemcpyAsync(dst[a], src[a], nm * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}
