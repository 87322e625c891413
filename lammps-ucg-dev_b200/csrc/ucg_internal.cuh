// ucg_internal.cuh — device context and shared helpers of libucgb200.so (sm_100a only).
//
// HBM layout: per-site state is held as arrays of 32-byte vector
// records so that every gather in the pair kernels is one aligned sector:
//   pos[i]  = {x, y, z, ucgl}            owned + ghost   (the neighbor gather record)
//   ts[i]   = type | ucgstate << 16      owned + ghost
//   vel[i]  = {vx, vy, vz, ucgvl}        owned
//   frc[i]  = {fx, fy, fz, ucgforce}     owned
//   scores[i] = {s0, s1}                 owned            (ucgsoftmaxscores)
//   ucgp, ucgml, mask, tag, molecule     scalar arrays
// Reference schema: atom.h:180-192, atom.cpp:590-609, UCG/atom_vec_ucg.cpp:48-90.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ucgb200.h"

#define UCG_NEIGHMASK 0x1FFFFFFF  // [stock] lmptype.h NEIGHMASK
#define UCG_SBBITS 30

namespace ucg {

struct TableDev {
  int style, n, tablength;
  int nmask, nshiftbits;
  double innersq, delta, invdelta, deltasq6, cut;
  const double2 *ef;    // {e[i], f[i]}
  const double2 *ef2;   // SPLINE {e2[i], f2[i]}
  const double2 *dedf;  // BITMAP {de[i], df[i]}
  const double2 *rd;    // BITMAP {rsq[i], drsq[i]}
};

// everything the pair kernels need about an (actual type i, actual type j) pair
struct PairInfo {
  double cutsq;       // cutsq[itype][jtype]  (pair_table_ucgld.cpp:213)
  double cutneighsq;  // (sqrt(cutsq)+skin)^2  [stock Neighbor::init]
  int ni, nj;         // n_states_per_type
  int tab[4];         // tabindex[formal(i,a)][formal(j,b)] at [a*2+b]
};
struct TypeInfo {
  int nstates;
  int pad;
  double dmu;   // chem_pot[formal1]-chem_pot[formal0]  (pair_table_ucgld.cpp:176)
  double mu0, mu1;
  double mass;  // atom->mass[type]
};

struct ErrWord {
  int code;
  int tag_i, tag_j;
  int pad;
  double rsq;
};

struct Grid {
  double lo[3];      // origin of cell (1,1,1) == sub-domain low corner
  double inv[3];     // 1/cell size
  int ninner[3];     // inner cells per dim
  int nc[3];         // ninner + 2 (one halo layer each side)
  int ncells;
};

template <class T>
struct Buf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = n + n / 8 + 256;
    T *q = nullptr;
    cudaError_t e = cudaMalloc((void **)&q, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return e;
      cudaStreamSynchronize(st);
    }
    if (p) cudaFree(p);
    p = q;
    cap = ncap;
    return cudaSuccess;
  }
  // exactly n elements when growing (used for the permutation twins, which must not
  // out-grow their partner and trigger a realloc ping-pong on every rebuild)
  cudaError_t ensure_exact(size_t n) {
    if (n <= cap) return cudaSuccess;
    T *q = nullptr;
    cudaError_t e = cudaMalloc((void **)&q, n * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p) cudaFree(p);
    p = q;
    cap = n;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace ucg

struct ucgb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  long long launches = 0;

  // units / box / dt
  double boltz = 1, ftm2v = 1, mvv2e = 1, dt = 0.005;
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {1, 1, 1}, prd[3] = {1, 1, 1};
  int periodic[3] = {1, 1, 1};
  double sublo[3] = {0, 0, 0}, subhi[3] = {1, 1, 1};
  bool sub_set = false;
  double special_lj[4] = {1, 1, 1, 1};

  // types / tables
  int n_actual = 0, n_formal = 0;
  std::vector<int> n_states, formal_from, tabindex;
  std::vector<double> chem_pot, mass, cutsq;
  double kT = 1.0;
  bool maps_dirty = true;
  std::vector<ucg::TableDev> tables;
  std::vector<void *> table_allocs;
  ucg::Buf<ucg::TableDev> d_tables;
  ucg::Buf<ucg::PairInfo> d_pairinfo;
  std::vector<ucg::PairInfo> h_pairinfo;   // host copy of what d_pairinfo holds (the F32 neighbor build takes its two thresholds from it)
  ucg::Buf<ucg::TypeInfo> d_typeinfo;
  bool fast_uniform = false;   // one 2-state actual type, LINEAR tables on one rsq grid
  int fast_tab[4] = {0, 0, 0, 0};
  ucg::Buf<double2> d_fast_table;  // interleaved rows for the shared-memory pair kernel
  int fast_ntab = 0, fast_len = 0;
  double max_cut = 0.0;

  // atoms
  int nlocal = 0, nghost = 0;
  ucg::Buf<double4> pos, pos_alt, vel, vel_alt, frc, frc_alt, xhold;
  ucg::Buf<double2> scores, scores_alt;
  ucg::Buf<double> ucgp, ucgp_alt, ucgml, ucgml_alt;
  ucg::Buf<double> stage_d;   // host<->device staging of atoms_upload / atoms_download
  // dump.cu: sort keys / chosen sites / radix-sort scratch / packed rows (and formatted text) of a dump
  ucg::Buf<unsigned> dump_keys;
  ucg::Buf<int> dump_sites;
  ucg::Buf<char> dump_tmp, dump_text;
  ucg::Buf<double> dump_buf;
  ucg::Buf<long long> dump_off;
  ucg::Buf<uint4> dump_slots;
  long long dump_text_bytes = 0;
  int parsed_rows = -1, parsed_fields = 0;   // block left in dump_buf by ucgb200_snapshot_parse
  ucg::Buf<int> stage_i;
  ucg::Buf<int> ts, ts_alt, mask, mask_alt, tag, tag_alt, mol, mol_alt, orig, orig_alt;
  // ghosts: sources = local periodic images + border records received from other bricks
  ucg::Buf<int> ghost_owner, ghost_code, ghost_src, slot_of_src, ghost_mask;   // ghost_mask: group bits of every ghost slot
  ucg::Buf<long long> ghost_key;
  ucg::Buf<int> img_counters, img_owner, img_code;   // [send lists by dest rank | local images]
  ucg::Buf<char> recv_border;
  ucg::Buf<int> mig_dest, mig_stay, mig_scan;
  struct Halo {
    int rank = 0, nranks = 1;
    int grid[3] = {1, 1, 1}, coord[3] = {0, 0, 0};
    int nsend = 0, nrecv = 0, nlimg = 0;
    std::vector<int> send_counts{0}, send_offsets{0, 0}, recv_counts{0}, mig_counts{0};
  } halo;

  // neighbor
  double skin = 0.3, cut_override = 0.0, cutneighmax = 0.0;
  ucg::Grid grid;
  ucg::Buf<int> cell_count, cell_start, cell_cursor, gcell_count, gcell_start, order, cell_of;
  ucg::Buf<int> scan_tmp, ghost_cnt, ghost_off;
  ucg::Buf<int> neigh, numneigh;
  // skin entries of every row are sorted by their distance at build time; levcnt[i] holds, for 8 displacement
  // levels, how many leading row entries can possibly be inside the cutoff (8 x uint16), and d_maxdisp the
  // largest squared displacement of any site since the build (bits of a double; exact skipping, neighbor.cu)
  ucg::Buf<uint4> levcnt;
  ucg::Buf<unsigned long long> d_maxdisp;
  bool maxdisp_valid = false;
  ucg::Buf<unsigned> statebits;
  ucg::Buf<double> pair_acc;   // pair_ucgld.cu, N3L bulk variant: 6 doubles per owned site
  // multi-brick runs: owned sites whose row holds no ghost received from another brick ("interior", first in
  // site_list) and the rest ("boundary"): the interior part of a pair evaluation runs while the halo is in flight
  ucg::Buf<int> row_flag, row_scan, site_list, d_part;   // d_part[0] = number of interior sites
  bool parts_valid = false;
  int pair_part = -1;          // next ucgb200_pair_ucgld call: -1 all sites, 0 interior, 1 boundary (reset by the call)
  int neigh_stride = 0;
  bool list_valid = false;
  int nbuilds = 0;
  ucg::Buf<int> d_flags;  // [0] rebuild flag, [1] neighbor overflow (max row), [2] ghost total, [3] lost atoms
  int *h_flags = nullptr; // pinned mirror

  // results
  ucg::Buf<double> d_partials, d_ev;  // block partials, final {E, v[6], ke, lke, cnt...}
  ucg::Buf<ucg::ErrWord> d_err;
  double h_ev[16] = {0};
  bool ev_valid = false;
  ucg::Buf<double> d_eatom, d_vatom;   // per-atom energy [n] / virial [6n] of the last pair call that asked for them
  bool eatom_valid = false, vatom_valid = false;

  // deck / run state
  ucgb200_deck deck{};
  bool deck_set = false;
  long long ntimestep = 0, beginstep = 0, endstep = 0;
  std::vector<double> gfactor1, gfactor2;  // cache key of the last uploaded factors
  std::vector<double> lang_g1, lang_g2;    // deck langevin factors (init(), :164-171)
  ucg::Buf<double> d_gfac;
  double lambda_temp = 0.0;
  double thermo[16] = {0};

  // timers
  bool timers_on = false;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_pair0 = nullptr, ev_pair1 = nullptr;
  cudaEvent_t ev_flag = nullptr;   // run.cu: marks the read-back of the rebuild flag inside the step stream
  double last_maxdisp = -1.0;      // largest squared displacement since the build, as last read back (< 0: unknown)
  double t_ms[4] = {0, 0, 0, 0};
  long long t_launch[4] = {0, 0, 0, 0};
  // non-blocking stage timing (ucgb200_timers(ctx, 3, ...)): event pairs from a ring, resolved when the ring wraps or
  // when the totals are read, so that the loop being timed keeps its asynchronous (speculative) schedule
  bool timers_async = false;
  struct StageEv { cudaEvent_t a = nullptr, b = nullptr; int slot = -1; };
  std::vector<StageEv> stage_ring;
  size_t stage_next = 0;
  bool pair_timed = false;

  // fix cluster_switch (cluster_switch.cu): per-molecule arrays are indexed by molecule id
  struct Cluster {
    bool set = false;
    int mol_seed = 0, mol_offset = 0, max_mol = -1, n_switch = 0, n_switch_per_mol = 0, nmol = 0, ntypes = 0, groupbit = 1;
    int seed = 1, nedges = 0, rounds = 0;
    unsigned long long ndrawn = 0;     // RanPark uniforms consumed so far (random_unequal, :915)
    long long next_reneighbor = 0;     // :71, :480
    double cutoff = 0, prob_on = 0, prob_off = 1;
    double stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<int> type_on, type_off;
    ucg::Buf<int> d_state, d_restrict, d_label, d_accept, d_present, d_sum, d_draws, d_rank, d_gmask, d_scratch;
    ucg::Buf<int2> d_edges;
    ucg::Buf<unsigned char> d_contact;
  } cluster;

  // rleucg / bethe_density configuration (types are STATE types in these styles)
  struct Density {
    bool set = false, sm_tables = false;
    int n_types = 0, n_actual = 0;
    std::vector<int> actual_from_state, n_states_of_type, use_entropy, tabindex;
    std::vector<double> threshold_radius, density_threshold, chem_pot, cutsq, mass;
    double T = 1.0;
    ucg::Buf<double> d_prob, d_partial, d_pforce, d_cvforce;
    ucg::Buf<int> d_tabindex;
    ucg::Buf<double> d_cutsq;
    ucg::Buf<char> d_rt;
  } dens;
  // table_ucg_bethe_density: per-actual-type density settings + scratch (types/tables as for ucgld)
  struct BetheDensity {
    bool set = false, dirty = true;
    int n_actual = 0;
    std::vector<int> use_density, use_entropy;
    std::vector<double> cv_th, r_th;
    ucg::Buf<char> d_bt;
    ucg::Buf<double> d_prob, d_partial, d_cvf;
  } bdens;
  // texture objects for the gathers of the table_ucgld kernel (the TEX pipe works beside the LSU pipe)
  struct TexSlot { const void *ptr = nullptr; size_t bytes = 0; cudaTextureObject_t tex = 0; } tex_pos[2], tex_sbits, tex_ts[2];
  // ucgb200_step_host: results leave on a second stream while the step is still running
  ucgb200_atoms *host_out = nullptr;
  unsigned host_out_fields = 0, host_out_done = 0;
  ucg::Buf<double4> posc;          // density styles, uniform case: {x, y, z, CV force a neighbor's sweep may react to} (pair_common.cuh)
  bool skip_initial_once = false;  // ucgb200_step_host already ran this step's initial_integrate (in two parts, under the uploads)
  cudaStream_t stream_dl = nullptr;
  cudaEvent_t ev_dl[10] = {};      // one per result field: a field's copy starts as soon as its own gather has run
  // uploads: the host->device copies queue back to back on their own stream, every pack kernel (context stream) waits
  // for its own copy only, so the link never idles behind a pack kernel
  cudaStream_t stream_ul = nullptr;
  cudaEvent_t ev_ul_start = nullptr, ev_ul[16] = {};
  void *comm_state = nullptr;  // comm.cu: NCCL communicator + exchange buffers of a multi-brick run
  bool ev_two_parts = false;   // d_ev[16..22] holds a second virial part to be added (rleucg)
};

namespace ucg {

#define UCG_CHECK(ctx, expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                    \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)

#define UCG_LAUNCHED(ctx)                                                                 \
  do {                                                                                    \
    (ctx)->launches++;                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(_e) + " at " +     \
                   __FILE__ + ":" + std::to_string(__LINE__);                             \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)

// Stage timer shared by run.cu and neighbor.cu.  Blocking mode (timers_on, !timers_async): one event synchronisation
// per stage.  Async mode: the pair of events is taken from a ring and its elapsed time added later.
inline void stage_resolve(ucgb200_ctx *c, ucgb200_ctx::StageEv &e) {
  if (e.slot < 0) return;
  cudaEventSynchronize(e.b);
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) c->t_ms[e.slot] += ms;
  e.slot = -1;
}
inline void stage_resolve_all(ucgb200_ctx *c) {
  for (auto &e : c->stage_ring) stage_resolve(c, e);
}
struct StageTimer {
  ucgb200_ctx *c;
  int slot;
  long long l0;
  ucgb200_ctx::StageEv *ev = nullptr;
  StageTimer(ucgb200_ctx *ctx, int s) : c(ctx), slot(s), l0(ctx->launches) {
    if (!c->timers_on) return;
    if (c->timers_async) {
      if (c->stage_ring.empty()) c->stage_ring.resize(256);
      ev = &c->stage_ring[c->stage_next++ % c->stage_ring.size()];
      stage_resolve(c, *ev);
      if (!ev->a) { cudaEventCreate(&ev->a); cudaEventCreate(&ev->b); }
      cudaEventRecord(ev->a, c->stream);
    } else {
      cudaEventRecord(c->ev_a, c->stream);
    }
  }
  void stop() {
    if (!c->timers_on) return;
    if (slot != 1) c->t_launch[slot] += c->launches - l0;
    if (c->timers_async) {
      cudaEventRecord(ev->b, c->stream);
      ev->slot = slot;
      return;
    }
    cudaEventRecord(c->ev_b, c->stream);
    cudaEventSynchronize(c->ev_b);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev_a, c->ev_b);
    c->t_ms[slot] += ms;
  }
};

inline int fail(ucgb200_ctx *c, const char *msg) {
  c->err = msg;
  return -1;
}
inline int nblocks(long long n, int bs) { return (int)((n + bs - 1) / bs); }

// Neighbor rows are stored in blocks of 16 entries, transposed 4x4: memory slot 4*l + m of a block holds
// logical entry 4*m + l.  The 4 lanes of a site in the table_ucgld kernel then fetch their next four
// entries (logical l, 4+l, 8+l, 12+l) with ONE 16-byte load each; every other consumer goes through
// rowslot().  Row capacities (neigh_stride) are multiples of 16.
__host__ __device__ __forceinline__ int rowslot(int k) { return (k & ~15) | ((k & 3) << 2) | ((k >> 2) & 3); }

// exact (non-contracted) squared distance, evaluated in the reference's order
// delx*delx + dely*dely + delz*delz  (pair_table_ucgld.cpp:211, [stock] npair)
__device__ __forceinline__ double rsq_exact(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ int cell_coord(double x, double lo, double inv, int nc) {
  int c = (int)floor((x - lo) * inv) + 1;
  return c < 0 ? 0 : (c > nc - 1 ? nc - 1 : c);
}
__device__ __forceinline__ int cell_index(const Grid &g, double x, double y, double z, bool owned) {
  int ix = cell_coord(x, g.lo[0], g.inv[0], g.nc[0]);
  int iy = cell_coord(y, g.lo[1], g.inv[1], g.nc[1]);
  int iz = cell_coord(z, g.lo[2], g.inv[2], g.nc[2]);
  if (owned) {  // owned atoms live in the inner cells (rounding at the faces)
    ix = min(max(ix, 1), g.ninner[0]);
    iy = min(max(iy, 1), g.ninner[1]);
    iz = min(max(iz, 1), g.ninner[2]);
  }
  return (iz * g.nc[1] + iy) * g.nc[0] + ix;
}

// Philox4x32-10 counter-based RNG (Salmon et al. 2011): stream-independent draws keyed by
// (seed, purpose) and counted by (tag, timestep) so results do not depend on atom order
// or on the domain decomposition.  Replaces the sequential RanMars draws of
// fix_ucgld_langevin.cpp:280 and fix_ucgstate.cpp:117 (statistical parity only).
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                             uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ double philox_uniform(uint32_t seed, uint32_t purpose, uint32_t tag,
                                                 unsigned long long step) {
  uint32_t c0 = tag, c1 = (uint32_t)step, c2 = (uint32_t)(step >> 32), c3 = purpose;
  uint32_t k0 = seed, k1 = 0x5bd1e995u ^ purpose;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  // 53-bit uniform in [0,1)
  unsigned long long m = ((unsigned long long)(c0 >> 5) << 26) | (unsigned long long)(c1 >> 6);
  return (double)m * (1.0 / 9007199254740992.0);
}

// host-side launchers implemented across the .cu files
int rebuild_maps(ucgb200_ctx *c);
int exclusive_scan(ucgb200_ctx *c, const int *in, int *out, int n, int *d_total);
int reduce_partials(ucgb200_ctx *c, int nblocks, int nvals, int out_offset);
int ucg_check_distance_launch(ucgb200_ctx *c);
}  // namespace ucg
// Forward halo over peer-mapped memory (comm.cu drives it, the kernels live in neighbor.cu next to the record types).
// Every brick owns one buffer that all its peers have mapped: a control block (one 16-byte word per (parity, source
// rank): rebuild flag, largest squared displacement, sequence number) followed by two receive regions (parity = sequence
// number & 1) of forward records grouped by source rank.
#define UCG_P2P_MAX_RANKS 16
struct UcgP2PCtl {                 // written by the source rank, read by the owner of the buffer
  unsigned long long maxdisp;      // bits of a double >= 0
  int flag;                        // Neighbor::decide flag of the source rank
  int seq;                         // written LAST (after a system-scope fence): the records and the two words above are visible
};
struct UcgPushTargets {
  int nranks, self, seq;
  int send_off[UCG_P2P_MAX_RANKS + 1];   // send list of this brick, grouped by destination rank
  char *rec[UCG_P2P_MAX_RANKS];          // where this brick's block starts inside the destination's receive region (this parity)
  UcgP2PCtl *ctl[UCG_P2P_MAX_RANKS];     // this brick's control word in the destination's buffer (this parity)
  const int *local_flag;                 // d_flags[0]
  const unsigned long long *local_maxdisp;
  unsigned *done;                        // block counter of the push kernel (zero between launches)
};
int ucg_classify_rows(ucgb200_ctx *c);                                // neighbor.cu: fills site_list / d_part of the current list
int ucg_halo_push_forward(ucgb200_ctx *c, const UcgPushTargets &t);   // neighbor.cu
int ucg_halo_wait_reduce(ucgb200_ctx *c, const UcgP2PCtl *ctl_mine, int nranks, int self, int seq);   // neighbor.cu
int ucg_nve_initial_part(ucgb200_ctx *c, double dtv, double dtf, int groupbit, int wall, int part);   // fixes.cu: 1 {x,v}, 2 {lambda,v_lambda}
int ucg_host_out_queue(ucgb200_ctx *c, unsigned mask);   // context.cu: gather + D2H of the not yet delivered fields in mask
int ucg_dump_pack_device(ucgb200_ctx *c, const ucgb200_dump_spec *sp, long long *nrows);   // dump.cu
int ucg_mb_forward_scalars(ucgb200_ctx *c, double *a0, double *a1, double *a2);   // comm.cu
int ucg_mb_allreduce_int(ucgb200_ctx *c, int *d_buf, int n, int op);
// context.cu: texture objects over pos / ts of the current buffers (0 when UCGB200_TEX=0)
int ucg_bind_gather_textures(ucgb200_ctx *c, cudaTextureObject_t *pos, cudaTextureObject_t *ts);
int ucg_bind_texture(ucgb200_ctx *c, ucgb200_ctx::TexSlot &slot, const void *ptr, size_t bytes, cudaChannelFormatDesc desc);             // comm.cu: 0 sum, 1 max, 2 min
namespace ucg {   // neighbor.cu: k_check_distance into d_flags[0]

}  // namespace ucg
