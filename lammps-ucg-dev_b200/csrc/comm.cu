// comm.cu — brick-to-brick exchange for the resident run loop: what stock Comm does over MPI for
// the reference (exchange / borders / forward_comm with AtomVecUCG's payloads,
// UCG/atom_vec_ucg.cpp:66-82), here NCCL point-to-point groups over NVLink/NVSwitch issued on the
// context stream.  Every exchange is one grouped send/recv per peer pair (all peers addressed
// directly: corners travel in one hop) between the device pack/unpack kernels of neighbor.cu;
// the per-step rebuild decision is a 4-byte ncclAllReduce(MAX) on the device flag, so a step costs
// exactly one host synchronisation (the flag read-back that Neighbor::decide needs).
//
// NCCL is resolved at run time (dlopen libnccl.so.2): single-GPU users of libucgb200.so need no
// NCCL, and inside a PyTorch process the already loaded NCCL is the one that is used.
#include "ucg_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <numeric>

using namespace ucg;

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &nccl() {
  static NcclApi api;
  if (api.lib) return api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
#define UCG_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name))
  UCG_SYM(GetUniqueId, "ncclGetUniqueId"); UCG_SYM(CommInitRank, "ncclCommInitRank"); UCG_SYM(CommDestroy, "ncclCommDestroy");
  UCG_SYM(Send, "ncclSend"); UCG_SYM(Recv, "ncclRecv"); UCG_SYM(GroupStart, "ncclGroupStart"); UCG_SYM(GroupEnd, "ncclGroupEnd");
  UCG_SYM(AllReduce, "ncclAllReduce"); UCG_SYM(AllGather, "ncclAllGather"); UCG_SYM(GetErrorString, "ncclGetErrorString");
#undef UCG_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart && api.GroupEnd &&
           api.AllReduce && api.AllGather && api.GetErrorString;
  return api;
}

#define UCG_NCCL(ctx, expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      (ctx)->err = std::string(#expr) + ": " + nccl().GetErrorString(_r);                     \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)

struct CommState {
  ncclComm_t comm = nullptr;
  Buf<char> send, recv, fwd_send, fwd_recv;
  Buf<int> d_counts;                 // [nranks] mine, then [nranks*nranks] gathered
  std::vector<int> send_counts, recv_counts;   // border == forward counts of the current list
  std::vector<int> count_matrix;     // [src * nranks + dst] of the last exchange_counts
  long long bytes_forward = 0;
  int nrebuilds = 0;
  // forward halo through peer-mapped memory (UCGB200_P2P=0 keeps the NCCL send/recv path)
  struct P2P {
    bool tried = false, mapped = false, active = false;   // active: the current send lists fit the mapped regions
    char *mine = nullptr;                                 // control block + two receive regions
    char *peer[UCG_P2P_MAX_RANKS] = {nullptr};
    size_t region_records = 0;                            // capacity of ONE receive region
    unsigned *done = nullptr;
    int seq = 0;
    int remote_off[UCG_P2P_MAX_RANKS] = {0};              // my block's first record inside rank r's receive region
    long long pushes = 0;
  } p2p;
};
constexpr size_t P2P_CTL_BYTES = 2 * UCG_P2P_MAX_RANKS * sizeof(UcgP2PCtl);   // [parity][source rank]
constexpr size_t P2P_REC = 48;                                                // sizeof(ForwardRec)

CommState *state(ucgb200_ctx *c) { return static_cast<CommState *>(c->comm_state); }

// counts[k] = records this rank sends to rank k  ->  recv[k] = records rank k sends to this rank
int exchange_counts(ucgb200_ctx *c, const std::vector<int> &counts, std::vector<int> &recv) {
  CommState *s = state(c);
  const int nr = c->halo.nranks, me = c->halo.rank;
  UCG_CHECK(c, s->d_counts.ensure((size_t)nr * (nr + 1)));
  UCG_CHECK(c, cudaMemcpyAsync(s->d_counts.p, counts.data(), nr * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  UCG_NCCL(c, nccl().AllGather(s->d_counts.p, s->d_counts.p + nr, nr, ncclInt32, s->comm, c->stream));
  std::vector<int> all((size_t)nr * nr);
  UCG_CHECK(c, cudaMemcpyAsync(all.data(), s->d_counts.p + nr, all.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  recv.resize(nr);
  for (int k = 0; k < nr; k++) recv[k] = all[(size_t)k * nr + me];
  s->count_matrix = all;
  return 0;
}

// ---------------------------------------------------------------- peer-mapped forward halo
// One buffer per brick, mapped by every peer through CUDA IPC (the bricks are processes on one NVSwitch node).  Set up
// once, at the first rebuild, when the size of the ghost shell is known; the handles travel through an ncclAllGather.
// Every brick must reach the same verdict (mapped or not), hence the final MIN all-reduce.
int p2p_setup(ucgb200_ctx *c) {
  CommState *s = state(c);
  CommState::P2P &p = s->p2p;
  if (p.tried) return 0;
  p.tried = true;
  const int nr = c->halo.nranks, me = c->halo.rank;
  const char *env = getenv("UCGB200_P2P");
  int ok = (!env || atoi(env) != 0) && nr <= UCG_P2P_MAX_RANKS ? 1 : 0;
  // region capacity: twice the largest ghost shell of any brick at this first rebuild (the liquid's density fluctuates
  // by a few per cent; a list that ever outgrows it falls back to the NCCL path for that list)
  size_t largest = 0;
  for (int dst = 0; dst < nr; dst++) {
    size_t col = 0;
    for (int src = 0; src < nr; src++) col += (size_t)s->count_matrix[(size_t)src * nr + dst];
    largest = std::max(largest, col);
  }
  p.region_records = 2 * largest + 4096;
  const size_t bytes = P2P_CTL_BYTES + 2 * p.region_records * P2P_REC;
  cudaIpcMemHandle_t mine{};
  if (ok) {
    if (cudaMalloc((void **)&p.mine, bytes) != cudaSuccess || cudaMalloc((void **)&p.done, sizeof(unsigned)) != cudaSuccess) ok = 0;
    if (ok) {
      cudaMemsetAsync(p.mine, 0xff, P2P_CTL_BYTES, c->stream);     // sequence numbers start at -1
      cudaMemsetAsync(p.done, 0, sizeof(unsigned), c->stream);
      if (cudaIpcGetMemHandle(&mine, p.mine) != cudaSuccess) ok = 0;
    }
    cudaGetLastError();
  }
  // all-gather the handles (64 bytes each) through device memory
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  Buf<char> hb;
  UCG_CHECK(c, hb.ensure((size_t)(nr + 1) * 64));
  UCG_CHECK(c, cudaMemcpyAsync(hb.p, &mine, 64, cudaMemcpyHostToDevice, c->stream));
  UCG_NCCL(c, nccl().AllGather(hb.p, hb.p + 64, 64, ncclChar, s->comm, c->stream));
  std::vector<cudaIpcMemHandle_t> all(nr);
  UCG_CHECK(c, cudaMemcpyAsync(all.data(), hb.p + 64, (size_t)nr * 64, cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  hb.release();
  // every brick must have produced a handle before anyone opens one
  UCG_CHECK(c, s->d_counts.ensure((size_t)nr * (nr + 1)));
  UCG_CHECK(c, cudaMemcpyAsync(s->d_counts.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  UCG_NCCL(c, nccl().AllReduce(s->d_counts.p, s->d_counts.p, 1, ncclInt32, ncclMin, s->comm, c->stream));
  UCG_CHECK(c, cudaMemcpyAsync(&ok, s->d_counts.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  if (ok) {
    for (int r = 0; r < nr && ok; r++) {
      if (r == me) { p.peer[r] = p.mine; continue; }
      void *q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
      p.peer[r] = (char *)q;
    }
  }
  UCG_CHECK(c, cudaMemcpyAsync(s->d_counts.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  UCG_NCCL(c, nccl().AllReduce(s->d_counts.p, s->d_counts.p, 1, ncclInt32, ncclMin, s->comm, c->stream));
  UCG_CHECK(c, cudaMemcpyAsync(&ok, s->d_counts.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  p.mapped = ok != 0;
  if (getenv("UCGB200_COMM_TRACE")) fprintf(stderr, "[ucgb200 rank %d] peer-mapped forward halo: %s (%zu records per region)\n", me, p.mapped ? "on" : "off (NCCL send/recv)", p.region_records);
  return 0;
}

// after every rebuild: where my block starts inside each destination's receive region, and whether every brick's ghost
// shell still fits its region (the same matrix on every brick, hence the same verdict)
void p2p_plan(ucgb200_ctx *c) {
  CommState *s = state(c);
  CommState::P2P &p = s->p2p;
  p.active = false;
  if (!p.mapped) return;
  const int nr = c->halo.nranks, me = c->halo.rank;
  bool fits = true;
  for (int dst = 0; dst < nr; dst++) {
    size_t col = 0, before_me = 0;
    for (int src = 0; src < nr; src++) {
      if (src == me) before_me = col;
      col += (size_t)s->count_matrix[(size_t)src * nr + dst];
    }
    if (col > p.region_records) fits = false;
    p.remote_off[dst] = (int)before_me;
  }
  p.active = fits;
}

// first half: this brick's records and control words go out, the local periodic images are refreshed
int p2p_forward_begin(ucgb200_ctx *c) {
  CommState *s = state(c);
  CommState::P2P &p = s->p2p;
  const int nr = c->halo.nranks, me = c->halo.rank;
  const int seq = ++p.seq, parity = seq & 1;
  UcgPushTargets t{};
  t.nranks = nr; t.self = me; t.seq = seq;
  for (int r = 0; r <= nr; r++) t.send_off[r] = c->halo.send_offsets[r];
  for (int r = 0; r < nr; r++) {
    char *base = p.peer[r];
    t.rec[r] = base + P2P_CTL_BYTES + ((size_t)parity * p.region_records + (size_t)p.remote_off[r]) * P2P_REC;
    t.ctl[r] = reinterpret_cast<UcgP2PCtl *>(base) + parity * UCG_P2P_MAX_RANKS + me;
  }
  t.local_flag = c->d_flags.p;
  t.local_maxdisp = c->d_maxdisp.p;
  t.done = p.done;
  int rc;
  if ((rc = ucg_halo_push_forward(c, t))) return rc;
  if ((rc = ucgb200_ghosts_forward(c))) return rc;      // local periodic images meanwhile
  return 0;
}
// second half: wait for every peer's push of this sequence number (it has usually landed long ago when interior work
// ran in between), fold the bricks' flags, move the records into the ghost slots
int p2p_forward_end(ucgb200_ctx *c) {
  CommState *s = state(c);
  CommState::P2P &p = s->p2p;
  const int nr = c->halo.nranks, me = c->halo.rank;
  const int seq = p.seq, parity = seq & 1;
  int rc;
  if ((rc = ucg_halo_wait_reduce(c, reinterpret_cast<const UcgP2PCtl *>(p.mine) + parity * UCG_P2P_MAX_RANKS, nr, me, seq))) return rc;
  if ((rc = ucgb200_halo_unpack_forward(c, p.mine + P2P_CTL_BYTES + (size_t)parity * p.region_records * P2P_REC))) return rc;
  p.pushes++;
  return 0;
}
int p2p_forward(ucgb200_ctx *c) {
  int rc = p2p_forward_begin(c);
  return rc ? rc : p2p_forward_end(c);
}

// send buffer grouped by destination, receive buffer grouped by source (the layouts of neighbor.cu)
int all_to_all(ucgb200_ctx *c, const char *sendbuf, char *recvbuf, const std::vector<int> &scount, const std::vector<int> &rcount,
               size_t rec) {
  CommState *s = state(c);
  const int nr = c->halo.nranks;
  UCG_NCCL(c, nccl().GroupStart());
  size_t so = 0, ro = 0;
  for (int k = 0; k < nr; k++) {
    const size_t sb = (size_t)scount[k] * rec, rb = (size_t)rcount[k] * rec;
    if (sb) UCG_NCCL(c, nccl().Send(sendbuf + so, sb, ncclChar, k, s->comm, c->stream));
    if (rb) UCG_NCCL(c, nccl().Recv(recvbuf + ro, rb, ncclChar, k, s->comm, c->stream));
    so += sb; ro += rb;
  }
  UCG_NCCL(c, nccl().GroupEnd());
  return 0;
}

int total(const std::vector<int> &v) { return std::accumulate(v.begin(), v.end(), 0); }

// forward_comm(Pair*) payloads: up to 3 per-site doubles of the owners -> their ghosts on other bricks
__global__ void k_pack_scalars(const double *a0, const double *a1, const double *a2, int na, const int *__restrict__ owner,
                               int n, double *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int o = owner[k];
  out[(size_t)k * na] = a0[o];
  if (na > 1) out[(size_t)k * na + 1] = a1[o];
  if (na > 2) out[(size_t)k * na + 2] = a2[o];
}
__global__ void k_unpack_scalars(double *a0, double *a1, double *a2, int na, int nlocal, const int *__restrict__ slot_of_src,
                                 int nlimg, int n, const double *__restrict__ in) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int s = nlocal + slot_of_src[nlimg + k];
  a0[s] = in[(size_t)k * na];
  if (na > 1) a1[s] = in[(size_t)k * na + 1];
  if (na > 2) a2[s] = in[(size_t)k * na + 2];
}

}  // namespace

// comm->forward_comm(this) of a pair style across bricks (pair_table_rleucg_interface.cpp:141-160, the repaired
// forward comm of pair_table_ucg_bethe_density.cpp): arrays are indexed like pos (owned, then ghosts); the local
// periodic images are refreshed by the caller.  Same send lists and receive order as the border exchange.
int ucg_mb_forward_scalars(ucgb200_ctx *c, double *a0, double *a1, double *a2) {
  if (c->halo.nranks < 2) return 0;
  CommState *s = state(c);
  if (!s) return fail(c, "multi-brick pair style without ucgb200_comm_init (the density styles need the resident NCCL driver)");
  const int na = 1 + (a1 != nullptr) + (a2 != nullptr);
  const int ns = total(s->send_counts), nr = total(s->recv_counts);
  const size_t rec = (size_t)na * sizeof(double);
  UCG_CHECK(c, s->send.ensure((size_t)ns * rec + 64));
  UCG_CHECK(c, s->recv.ensure((size_t)nr * rec + 64));
  if (ns) {
    k_pack_scalars<<<nblocks(ns, 256), 256, 0, c->stream>>>(a0, a1, a2, na, c->img_owner.p, ns, reinterpret_cast<double *>(s->send.p));
    UCG_LAUNCHED(c);
  }
  int rc = all_to_all(c, s->send.p, s->recv.p, s->send_counts, s->recv_counts, rec);
  if (rc) return rc;
  if (nr) {
    k_unpack_scalars<<<nblocks(nr, 256), 256, 0, c->stream>>>(a0, a1, a2, na, c->nlocal, c->slot_of_src.p, c->halo.nlimg, nr,
                                                              reinterpret_cast<const double *>(s->recv.p));
    UCG_LAUNCHED(c);
  }
  return 0;
}

extern "C" int ucgb200_comm_unique_id(char *id, int len) {
  if (!id || len < (int)sizeof(ncclUniqueId)) return -1;
  if (!nccl().ok) return -4;
  ncclUniqueId u;
  if (nccl().GetUniqueId(&u) != ncclSuccess) return -2;
  memcpy(id, &u, sizeof(u));
  return (int)sizeof(u);
}

extern "C" int ucgb200_comm_init(ucgb200_ctx *c, const char *id, int len) {
  if (!c || !id || len < (int)sizeof(ncclUniqueId)) return -1;
  if (c->halo.nranks < 2) return fail(c, "comm_init: configure the brick grid first (ucgb200_halo_configure)");
  if (!nccl().ok) return fail(c, "comm_init: libnccl.so.2 could not be loaded");
  cudaSetDevice(c->device);
  if (c->comm_state) return fail(c, "comm_init: already initialised");
  auto *s = new CommState();
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclResult_t r = nccl().CommInitRank(&s->comm, c->halo.nranks, u, c->halo.rank);
  if (r != ncclSuccess) {
    c->err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r);
    delete s;
    return -2;
  }
  c->comm_state = s;
  return 0;
}

extern "C" int ucgb200_comm_destroy(ucgb200_ctx *c) {
  if (!c) return -1;
  CommState *s = state(c);
  if (!s) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int r = 0; r < UCG_P2P_MAX_RANKS; r++)
    if (s->p2p.peer[r] && s->p2p.peer[r] != s->p2p.mine) cudaIpcCloseMemHandle(s->p2p.peer[r]);
  if (s->p2p.mine) cudaFree(s->p2p.mine);
  if (s->p2p.done) cudaFree(s->p2p.done);
  if (s->comm) nccl().CommDestroy(s->comm);
  s->send.release(); s->recv.release(); s->fwd_send.release(); s->fwd_recv.release(); s->d_counts.release();
  delete s;
  c->comm_state = nullptr;
  return 0;
}

// comm->exchange() + comm->borders() + neighbor->build(): the rebuild of a multi-brick run
// UCGB200_COMM_TRACE=2: wall-clock marks (with a stream synchronisation each) through a multi-brick rebuild, rank 0, stderr
struct CommTrace {
  ucgb200_ctx *c;
  bool on;
  double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit CommTrace(ucgb200_ctx *ctx) : c(ctx), on(ctx->halo.rank == 0 && getenv("UCGB200_COMM_TRACE") && atoi(getenv("UCGB200_COMM_TRACE")) == 2), t0(0) {
    if (on) { cudaStreamSynchronize(c->stream); t0 = now(); }
  }
  void mark(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(c->stream);
    const double t = now();
    fprintf(stderr, "[mb_rebuild] %-30s %8.3f ms\n", what, t - t0);
    t0 = t;
  }
};

int ucg_mb_rebuild(ucgb200_ctx *c) {
  CommState *s = state(c);
  if (!s) return fail(c, "multi-brick run without ucgb200_comm_init");
  const int nr = c->halo.nranks;
  int rb = 0, rf = 0, rm = 0, rc;
  ucgb200_halo_record_bytes(&rb, &rf, &rm);
  CommTrace tr(c);
  // sites that left their brick
  std::vector<int> sc(nr), rcv;
  if ((rc = ucgb200_migrate_prepare(c, sc.data()))) return rc;
  tr.mark("migrate_prepare");
  if ((rc = exchange_counts(c, sc, rcv))) return rc;
  tr.mark("exchange_counts (migrate)");
  UCG_CHECK(c, s->send.ensure((size_t)total(sc) * rm + 64));
  UCG_CHECK(c, s->recv.ensure((size_t)total(rcv) * rm + 64));
  if ((rc = ucgb200_migrate_pack(c, s->send.p))) return rc;
  if ((rc = all_to_all(c, s->send.p, s->recv.p, sc, rcv, rm))) return rc;
  if ((rc = ucgb200_migrate_unpack(c, s->recv.p, total(rcv)))) return rc;
  tr.mark("migrate pack+a2a+unpack");
  // ghost shells
  if ((rc = ucgb200_neigh_build_local(c))) return rc;
  tr.mark("build_local");
  if ((rc = ucgb200_halo_send_counts(c, sc.data()))) return rc;
  if ((rc = exchange_counts(c, sc, rcv))) return rc;
  tr.mark("exchange_counts (borders)");
  UCG_CHECK(c, s->send.ensure((size_t)total(sc) * rb + 64));
  UCG_CHECK(c, s->recv.ensure((size_t)total(rcv) * rb + 64));
  if ((rc = ucgb200_halo_pack_border(c, s->send.p))) return rc;
  if ((rc = all_to_all(c, s->send.p, s->recv.p, sc, rcv, rb))) return rc;
  if ((rc = ucgb200_halo_unpack_border(c, s->recv.p, rcv.data()))) return rc;
  tr.mark("border pack+a2a+unpack");
  if ((rc = ucgb200_neigh_build_finish(c))) return rc;
  tr.mark("build_finish");
  s->send_counts = sc;
  s->recv_counts = rcv;
  if ((rc = p2p_setup(c))) return rc;     // first rebuild only
  p2p_plan(c);
  // UCGB200_OVERLAP=1: interior / boundary split of the pair evaluation around the halo (run.cu).  Off by default:
  // measured on 8 GPUs it LOSES 0.015 ms per step (two persistent launches stage the table twice and end in two
  // partial waves, against ~0.02 ms of exchange latency that the push already keeps off the critical path).
  if (s->p2p.active && getenv("UCGB200_OVERLAP") && atoi(getenv("UCGB200_OVERLAP")) == 1 && (rc = ucg_classify_rows(c))) return rc;
  UCG_CHECK(c, s->fwd_send.ensure((size_t)total(sc) * rf + 64));
  UCG_CHECK(c, s->fwd_recv.ensure((size_t)total(rcv) * rf + 64));
  s->nrebuilds++;
  return 0;
}

// comm->forward_comm(): refresh every ghost (records from other bricks + local periodic images)
int ucg_mb_forward(ucgb200_ctx *c);
// with_decide: the rebuild flag and displacement bound of every brick are reduced along the way (MAX), into d_flags[0]
// and d_maxdisp of every brick: the peer-mapped push carries them in its control words, the NCCL path adds the two
// small all-reduces of ucg_mb_decide.  Neither synchronises with the host.
int ucg_mb_forward_reduce(ucgb200_ctx *c, bool with_decide) {
  CommState *s = state(c);
  if (!s) return fail(c, "multi-brick run without ucgb200_comm_init");
  if (s->p2p.active) { s->bytes_forward += (long long)total(s->send_counts) * (long long)P2P_REC; return p2p_forward(c); }
  int rc = ucg_mb_forward(c);
  if (rc || !with_decide) return rc;
  UCG_NCCL(c, nccl().GroupStart());
  UCG_NCCL(c, nccl().AllReduce(c->d_flags.p, c->d_flags.p, 1, ncclInt32, ncclMax, s->comm, c->stream));
  UCG_NCCL(c, nccl().AllReduce(c->d_maxdisp.p, c->d_maxdisp.p, 1, ncclUint64, ncclMax, s->comm, c->stream));
  UCG_NCCL(c, nccl().GroupEnd());
  return 0;
}

int ucg_mb_forward(ucgb200_ctx *c) {
  CommState *s = state(c);
  if (!s) return fail(c, "multi-brick run without ucgb200_comm_init");
  if (s->p2p.active) return ucg_mb_forward_reduce(c, false);
  int rf = 0, rc;
  ucgb200_halo_record_bytes(nullptr, &rf, nullptr);
  if ((rc = ucgb200_halo_pack_forward(c, s->fwd_send.p))) return rc;
  if ((rc = ucgb200_ghosts_forward(c))) return rc;
  if ((rc = all_to_all(c, s->fwd_send.p, s->fwd_recv.p, s->send_counts, s->recv_counts, rf))) return rc;
  if ((rc = ucgb200_halo_unpack_forward(c, s->fwd_recv.p))) return rc;
  s->bytes_forward += (long long)total(s->send_counts) * rf;
  return 0;
}

// Neighbor::decide across bricks: MAX of the per-brick flags.  prechecked: k_check_distance already
// ran (fused step tail) and d_flags[0] holds this brick's flag.
int ucg_mb_decide(ucgb200_ctx *c, bool prechecked, int *rebuild) {
  CommState *s = state(c);
  if (!s) return fail(c, "multi-brick run without ucgb200_comm_init");
  cudaSetDevice(c->device);
  if (!prechecked) {
    UCG_CHECK(c, cudaMemsetAsync(c->d_flags.p, 0, sizeof(int), c->stream));
    UCG_CHECK(c, cudaMemsetAsync(c->d_maxdisp.p, 0, sizeof(unsigned long long), c->stream));
    if (!c->list_valid) {
      const int one = 1;
      UCG_CHECK(c, cudaMemcpyAsync(c->d_flags.p, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    } else if (c->nlocal > 0) {
      int rc = ucg_check_distance_launch(c);
      if (rc) return rc;
    }
  }
  // one fused launch: the rebuild flag and the largest squared displacement of any site of any brick (the
  // bound that lets the pair kernel skip skin entries must hold for ghosts owned by other bricks too)
  UCG_NCCL(c, nccl().GroupStart());
  UCG_NCCL(c, nccl().AllReduce(c->d_flags.p, c->d_flags.p, 1, ncclInt32, ncclMax, s->comm, c->stream));
  UCG_NCCL(c, nccl().AllReduce(c->d_maxdisp.p, c->d_maxdisp.p, 1, ncclUint64, ncclMax, s->comm, c->stream));
  UCG_NCCL(c, nccl().GroupEnd());
  c->maxdisp_valid = true;
  UCG_CHECK(c, cudaMemcpyAsync(c->h_flags, c->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  *rebuild = c->h_flags[0] ? 1 : 0;
  return 0;
}

// MPI_Allreduce of an int array in place (fix_cluster_switch.cpp:109-111,160-161,586,682-683,777);
// op: 0 sum, 1 max, 2 min.  No-op on a single brick.
int ucg_mb_allreduce_int(ucgb200_ctx *c, int *d_buf, int n, int op) {
  if (c->halo.nranks < 2 || n <= 0) return 0;
  CommState *s = state(c);
  if (!s) return fail(c, "multi-brick fix cluster_switch without ucgb200_comm_init (it needs the resident NCCL driver)");
  const ncclRedOp_t o = op == 0 ? ncclSum : (op == 1 ? ncclMax : ncclMin);
  UCG_NCCL(c, nccl().AllReduce(d_buf, d_buf, (size_t)n, ncclInt32, o, s->comm, c->stream));
  return 0;
}

// split form of ucg_mb_forward_reduce for the peer-mapped transport (run.cu puts the interior pair work in between)
int ucg_mb_forward_begin(ucgb200_ctx *c) {
  CommState *s = state(c);
  if (!s || !s->p2p.active) return fail(c, "forward_begin: peer-mapped transport not active");
  s->bytes_forward += (long long)total(s->send_counts) * (long long)P2P_REC;
  return p2p_forward_begin(c);
}
int ucg_mb_forward_end(ucgb200_ctx *c) { return p2p_forward_end(c); }

int ucg_mb_p2p_active(ucgb200_ctx *c) {
  CommState *s = state(c);
  return s && s->p2p.active ? 1 : 0;
}

extern "C" int ucgb200_comm_transport(ucgb200_ctx *c, int *peer_mapped, long long *pushes) {
  if (!c) return -1;
  CommState *s = state(c);
  if (!s) return fail(c, "comm not initialised");
  if (peer_mapped) *peer_mapped = s->p2p.active ? 1 : 0;
  if (pushes) *pushes = s->p2p.pushes;
  return 0;
}

extern "C" int ucgb200_comm_stats(ucgb200_ctx *c, long long *bytes_forward, int *nrebuilds, int *send_records) {
  if (!c) return -1;
  CommState *s = state(c);
  if (!s) return fail(c, "comm not initialised");
  if (bytes_forward) *bytes_forward = s->bytes_forward;
  if (nrebuilds) *nrebuilds = s->nrebuilds;
  if (send_records) *send_records = total(s->send_counts);
  return 0;
}
