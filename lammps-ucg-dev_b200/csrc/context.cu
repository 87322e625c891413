// context.cu — lifecycle, set-up calls, table upload and the AtomVecUCG <-> device
// record conversion of libucgb200.so.
#include "ucg_internal.cuh"

#include <algorithm>
#include <functional>
#include <cmath>

using namespace ucg;

// ------------------------------------------------------------------ lifecycle
extern "C" int ucgb200_create(int device, ucgb200_ctx **out) {
  if (!out) return -1;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return -3;  // no CUDA device: there is no CPU fallback
  if (device < 0 || device >= ndev) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -2;
  ucgb200_ctx *c = new ucgb200_ctx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return -2; }
  cudaEventCreate(&c->ev_a);
  cudaEventCreate(&c->ev_b);
  cudaEventCreate(&c->ev_pair0);
  cudaEventCreate(&c->ev_pair1);
  c->d_flags.ensure(8);
  cudaMemset(c->d_flags.p, 0, 8 * sizeof(int));
  cudaHostAlloc((void **)&c->h_flags, 16 * sizeof(int), cudaHostAllocDefault);
  if (c->h_flags) memset(c->h_flags, 0, 16 * sizeof(int));
  c->d_err.ensure(1);
  cudaMemset(c->d_err.p, 0, sizeof(ErrWord));
  c->d_ev.ensure(32);
  c->d_maxdisp.ensure(4);
  cudaMemset(c->d_maxdisp.p, 0, 4 * sizeof(unsigned long long));
  cudaMemset(c->d_ev.p, 0, 32 * sizeof(double));
  *out = c;
  return 0;
}

extern "C" int ucgb200_destroy(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  ucgb200_comm_destroy(c);
  for (auto &e : c->stage_ring) { if (e.a) cudaEventDestroy(e.a); if (e.b) cudaEventDestroy(e.b); }
  if (c->stream_dl) cudaStreamDestroy(c->stream_dl);
  if (c->stream_ul) cudaStreamDestroy(c->stream_ul);
  for (cudaEvent_t e : c->ev_dl) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_ul) if (e) cudaEventDestroy(e);
  if (c->ev_ul_start) cudaEventDestroy(c->ev_ul_start);
  for (auto *t : {&c->tex_pos[0], &c->tex_pos[1], &c->tex_sbits, &c->tex_ts[0], &c->tex_ts[1]}) if (t->tex) cudaDestroyTextureObject(t->tex);
  for (void *p : c->table_allocs) cudaFree(p);
  // Buf<> members are released explicitly (no destructors: buffers may be swapped)
  c->d_tables.release(); c->d_pairinfo.release(); c->d_typeinfo.release(); c->d_fast_table.release();
  c->posc.release();
  c->pos.release(); c->pos_alt.release(); c->vel.release(); c->vel_alt.release();
  c->frc.release(); c->frc_alt.release(); c->xhold.release();
  c->scores.release(); c->scores_alt.release(); c->ucgp.release(); c->ucgp_alt.release();
  c->ucgml.release(); c->ucgml_alt.release(); c->ts.release(); c->ts_alt.release();
  c->mask.release(); c->mask_alt.release(); c->tag.release(); c->tag_alt.release();
  c->mol.release(); c->mol_alt.release(); c->orig.release(); c->orig_alt.release();
  c->ghost_owner.release(); c->ghost_code.release(); c->ghost_key.release(); c->ghost_src.release(); c->slot_of_src.release();
  c->img_counters.release(); c->img_owner.release(); c->img_code.release(); c->recv_border.release();
  c->mig_dest.release(); c->mig_stay.release(); c->mig_scan.release();
  c->cell_count.release(); c->cell_start.release(); c->cell_cursor.release();
  c->gcell_count.release(); c->gcell_start.release(); c->order.release(); c->cell_of.release();
  c->scan_tmp.release(); c->ghost_cnt.release(); c->ghost_off.release();
  c->neigh.release(); c->numneigh.release(); c->statebits.release(); c->d_flags.release();
  c->dump_keys.release(); c->dump_sites.release(); c->dump_tmp.release(); c->dump_text.release(); c->dump_buf.release(); c->dump_off.release(); c->dump_slots.release();
  c->stage_d.release(); c->stage_i.release(); c->levcnt.release(); c->d_maxdisp.release();
  c->d_partials.release(); c->d_ev.release(); c->d_err.release(); c->d_gfac.release();
  {
    auto &k = c->cluster;
    k.d_state.release(); k.d_restrict.release(); k.d_label.release(); k.d_accept.release(); k.d_present.release();
    k.d_sum.release(); k.d_draws.release(); k.d_rank.release(); k.d_gmask.release(); k.d_scratch.release();
    k.d_edges.release(); k.d_contact.release();
    c->bdens.d_bt.release(); c->bdens.d_prob.release(); c->bdens.d_partial.release(); c->bdens.d_cvf.release();
  }
  c->dens.d_prob.release(); c->dens.d_partial.release(); c->dens.d_pforce.release();
  c->dens.d_cvforce.release(); c->dens.d_tabindex.release(); c->dens.d_cutsq.release(); c->dens.d_rt.release();
  if (c->h_flags) cudaFreeHost(c->h_flags);
  cudaEventDestroy(c->ev_a); cudaEventDestroy(c->ev_b);
  cudaEventDestroy(c->ev_pair0); cudaEventDestroy(c->ev_pair1);
  if (c->ev_flag) cudaEventDestroy(c->ev_flag);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" const char *ucgb200_last_error(const ucgb200_ctx *c) { return c ? c->err.c_str() : "null context"; }

extern "C" int ucgb200_set_stream(ucgb200_ctx *c, void *s) {
  if (!c) return -1;
  cudaStreamSynchronize(c->stream);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (s) { c->stream = (cudaStream_t)s; c->own_stream = false; }
  else { UCG_CHECK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  return 0;
}
extern "C" int ucgb200_sync(ucgb200_ctx *c) {
  if (!c) return -1;
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" long long ucgb200_launch_count(const ucgb200_ctx *c) { return c ? c->launches : -1; }

// --------------------------------------------------------------------- set-up
extern "C" int ucgb200_set_units(ucgb200_ctx *c, double boltz, double ftm2v, double mvv2e) {
  if (!c) return -1;
  c->boltz = boltz; c->ftm2v = ftm2v; c->mvv2e = mvv2e;
  return 0;
}
extern "C" int ucgb200_set_box(ucgb200_ctx *c, const double lo[3], const double hi[3], const int periodic[3]) {
  if (!c || !lo || !hi) return -1;
  for (int d = 0; d < 3; d++) {
    if (!(hi[d] > lo[d])) return fail(c, "set_box: hi <= lo");
    c->boxlo[d] = lo[d]; c->boxhi[d] = hi[d]; c->prd[d] = hi[d] - lo[d];
    c->periodic[d] = periodic ? periodic[d] : 1;
    if (!c->sub_set) { c->sublo[d] = lo[d]; c->subhi[d] = hi[d]; }
  }
  c->list_valid = false;
  return 0;
}
// page-locked host staging for the host layer above the C-ABI (dump text / packed rows): D2H copies into it run as DMA
extern "C" int ucgb200_pinned_alloc(size_t bytes, void **out) {
  if (!out) return -1;
  *out = nullptr;
  return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? 0 : -2;
}
extern "C" int ucgb200_pinned_free(void *p) {
  if (p) cudaFreeHost(p);
  return 0;
}
// page-lock caller-owned host arrays (the LAMMPS per-atom arrays of the offload-mode classes) so that uploads and
// downloads run as DMA; the caller must unregister before the memory is freed or moved
extern "C" int ucgb200_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return -1;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return -2; }
  return 0;
}
extern "C" int ucgb200_host_unregister(void *p) {
  if (!p) return -1;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { cudaGetLastError(); return -2; }
  return 0;
}
extern "C" int ucgb200_halo_info(const ucgb200_ctx *c, int *rank, int *nranks) {
  if (!c) return -1;
  if (rank) *rank = c->halo.rank;
  if (nranks) *nranks = c->halo.nranks;
  return 0;
}
extern "C" int ucgb200_get_box(const ucgb200_ctx *c, double lo[3], double hi[3], int periodic[3]) {
  if (!c) return -1;
  for (int d = 0; d < 3; d++) {
    if (lo) lo[d] = c->boxlo[d];
    if (hi) hi[d] = c->boxhi[d];
    if (periodic) periodic[d] = c->periodic[d];
  }
  return 0;
}
extern "C" int ucgb200_set_subdomain(ucgb200_ctx *c, const double sublo[3], const double subhi[3]) {
  if (!c || !sublo || !subhi) return -1;
  for (int d = 0; d < 3; d++) { c->sublo[d] = sublo[d]; c->subhi[d] = subhi[d]; }
  c->sub_set = true;
  c->list_valid = false;
  return 0;
}
extern "C" int ucgb200_set_timestep(ucgb200_ctx *c, double dt) {
  if (!c || !(dt > 0)) return -1;
  c->dt = dt;
  return 0;
}
extern "C" int ucgb200_set_special_lj(ucgb200_ctx *c, const double s[4]) {
  if (!c || !s) return -1;
  for (int k = 0; k < 4; k++) c->special_lj[k] = s[k];
  return 0;
}
extern "C" int ucgb200_set_types(ucgb200_ctx *c, int n_actual, int n_formal, const int *n_states,
                                 const int *formal_from_actual, const double *chem_pot, const double *mass) {
  if (!c || n_actual < 1 || n_formal < n_actual || !n_states || !formal_from_actual) return -1;
  c->n_actual = n_actual; c->n_formal = n_formal;
  c->n_states.assign(n_states, n_states + n_actual + 1);
  c->formal_from.assign(formal_from_actual, formal_from_actual + 2 * (n_actual + 1));
  c->chem_pot.assign(n_formal + 1, 0.0);
  c->mass.assign(n_formal + 1, 1.0);
  for (int i = 1; i <= n_formal; i++) {
    if (chem_pot) c->chem_pot[i] = chem_pot[i];
    if (mass) c->mass[i] = mass[i];
  }
  for (int t = 1; t <= n_actual; t++) {
    // "Invalid number of states for atom type ... Only 1 or 2 states are allowed."
    // (pair_table_ucgld.cpp:613-615)
    if (n_states[t] < 1 || n_states[t] > 2) return fail(c, "Invalid number of states for atom type: only 1 or 2 states are allowed");
    if (n_states[t] == 2)
      for (int k = 0; k < 2; k++) {
        int ft = formal_from_actual[2 * t + k];
        if (ft < 1 || ft > n_formal) return fail(c, "Formal type not defined in pair_style command");
      }
  }
  c->maps_dirty = true;
  return 0;
}
extern "C" int ucgb200_set_kT(ucgb200_ctx *c, double kT) {
  if (!c) return -1;
  if (!(kT > 0)) return fail(c, "kT must be positive (no fix exports t_target? see SURVEY Q2)");
  c->kT = kT;
  return 0;
}

// (re)create a linear texture object only when the buffer moved or grew
int ucg_bind_texture(ucgb200_ctx *c, ucgb200_ctx::TexSlot &slot, const void *ptr, size_t bytes, cudaChannelFormatDesc desc) {
  if (slot.tex && slot.ptr == ptr && slot.bytes == bytes) return 0;
  if (slot.tex) cudaDestroyTextureObject(slot.tex);
  slot.tex = 0;
  slot.ptr = nullptr;
  {
    // 1-D linear textures address at most cudaDevAttrMaxTexture1DLinearWidth texels (2^27..2^28): beyond that the
    // caller gets a zero object and takes its LSU-pipe gather path instead of failing
    int maxw = 0;
    cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxTexture1DLinearWidth, c->device);
    const size_t texel = (size_t)(desc.x + desc.y + desc.z + desc.w) / 8;
    if (maxw > 0 && texel > 0 && bytes / texel > (size_t)maxw) return 0;
  }
  cudaResourceDesc res{};
  res.resType = cudaResourceTypeLinear;
  res.res.linear.devPtr = const_cast<void *>(ptr);
  res.res.linear.desc = desc;
  res.res.linear.sizeInBytes = bytes;
  cudaTextureDesc td{};
  td.readMode = cudaReadModeElementType;
  UCG_CHECK(c, cudaCreateTextureObject(&slot.tex, &res, &td, nullptr));
  slot.ptr = ptr; slot.bytes = bytes;
  return 0;
}

// pos / ts swap with their twins at every rebuild: two slots each, the one already bound to the live buffer
// is reused, otherwise the slot that is not bound to the twin
template <class B>
static ucgb200_ctx::TexSlot &pick_slot(ucgb200_ctx::TexSlot (&slots)[2], const B &live, const B &twin) {
  int pick = slots[0].ptr == live.p ? 0 : (slots[1].ptr == live.p ? 1 : -1);
  if (pick < 0) pick = (slots[0].ptr == twin.p && slots[0].tex) ? 1 : 0;
  return slots[pick];
}

int ucg_bind_gather_textures(ucgb200_ctx *c, cudaTextureObject_t *pos, cudaTextureObject_t *ts) {
  if (pos) *pos = 0;
  if (ts) *ts = 0;
  const char *e = getenv("UCGB200_TEX");
  if (e && atoi(e) == 0) return 0;
  int rc;
  if (pos) {
    ucgb200_ctx::TexSlot &s = pick_slot(c->tex_pos, c->pos, c->pos_alt);
    if ((rc = ucg_bind_texture(c, s, c->pos.p, c->pos.cap * sizeof(double4), cudaCreateChannelDesc<int4>()))) return rc;
    *pos = s.tex;
  }
  if (ts) {
    ucgb200_ctx::TexSlot &s = pick_slot(c->tex_ts, c->ts, c->ts_alt);
    if ((rc = ucg_bind_texture(c, s, c->ts.p, c->ts.cap * sizeof(int), cudaCreateChannelDesc<int>()))) return rc;
    *ts = s.tex;
  }
  return 0;
}

extern "C" int ucgb200_tables_clear(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaStreamSynchronize(c->stream);
  for (void *p : c->table_allocs) cudaFree(p);
  c->table_allocs.clear();
  c->tables.clear();
  c->maps_dirty = true;
  // a new pair_style starts from scratch: forget the style-specific configuration
  c->dens.set = false;
  c->bdens.set = false;
  c->list_valid = false;
  return 0;
}

static int upload_interleaved(ucgb200_ctx *c, const double *a, const double *b, int n, const double2 **out) {
  std::vector<double2> h(n);
  for (int i = 0; i < n; i++) { h[i].x = a[i]; h[i].y = b ? b[i] : 0.0; }
  double2 *d = nullptr;
  UCG_CHECK(c, cudaMalloc((void **)&d, n * sizeof(double2)));
  UCG_CHECK(c, cudaMemcpy(d, h.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
  c->table_allocs.push_back(d);
  *out = d;
  return 0;
}

extern "C" int ucgb200_table_upload(ucgb200_ctx *c, int tabstyle, int tablength, int n, double innersq,
                                    double delta, double invdelta, double deltasq6, double cut, int nmask,
                                    int nshiftbits, const double *e, const double *f, const double *e2,
                                    const double *f2, const double *rsq, const double *drsq,
                                    const double *de, const double *df, int *index) {
  if (!c || !e || !f || n < 1) return -1;
  if (tabstyle < 0 || tabstyle > 3) return fail(c, "Unknown table style");
  if (tablength < 2) return fail(c, "Illegal number of pair table entries");
  if (tabstyle == UCGB200_TAB_SPLINE && (!e2 || !f2)) return fail(c, "SPLINE table needs e2/f2");
  if (tabstyle == UCGB200_TAB_BITMAP && (!rsq || !drsq || !de || !df)) return fail(c, "BITMAP table needs rsq/drsq/de/df");
  cudaSetDevice(c->device);
  TableDev t{};
  t.style = tabstyle; t.n = n; t.tablength = tablength; t.nmask = nmask; t.nshiftbits = nshiftbits;
  t.innersq = innersq; t.delta = delta; t.invdelta = invdelta; t.deltasq6 = deltasq6; t.cut = cut;
  int rc;
  if ((rc = upload_interleaved(c, e, f, n, &t.ef))) return rc;
  if (tabstyle == UCGB200_TAB_SPLINE && (rc = upload_interleaved(c, e2, f2, n, &t.ef2))) return rc;
  if (tabstyle == UCGB200_TAB_BITMAP) {
    if ((rc = upload_interleaved(c, de, df, n, &t.dedf))) return rc;
    if ((rc = upload_interleaved(c, rsq, drsq, n, &t.rd))) return rc;
  }
  c->tables.push_back(t);
  if (index) *index = (int)c->tables.size() - 1;
  c->maps_dirty = true;
  return 0;
}

extern "C" int ucgb200_set_pair_maps(ucgb200_ctx *c, const int *tabindex, const double *cutsq) {
  if (!c || !tabindex || !cutsq) return -1;
  if (c->n_formal < 1) return fail(c, "set_types must precede set_pair_maps");
  int nt = c->n_formal + 1;
  c->tabindex.assign(tabindex, tabindex + nt * nt);
  c->cutsq.assign(cutsq, cutsq + nt * nt);
  c->maps_dirty = true;
  return 0;
}

// Build the device-side type/pair maps and, when the deck qualifies, the interleaved
// table for the shared-memory pair kernel.
int ucg_rebuild_rle_maps(ucgb200_ctx *c);

int ucg::rebuild_maps(ucgb200_ctx *c) {
  if (!c->maps_dirty) return 0;
  if (c->dens.set) { cudaSetDevice(c->device); return ucg_rebuild_rle_maps(c); }
  if (c->n_actual < 1) return fail(c, "types not set");
  if (c->tabindex.empty()) return fail(c, "All pair coeffs are not set");
  cudaSetDevice(c->device);
  int na = c->n_actual + 1, nt = c->n_formal + 1;
  auto formal = [&](int t, int k) { return c->n_states[t] == 1 ? t : c->formal_from[2 * t + k]; };
  std::vector<TypeInfo> ti(na);
  for (int t = 1; t < na; t++) {
    ti[t].nstates = c->n_states[t];
    ti[t].mass = c->mass[t];
    ti[t].mu0 = c->chem_pot[formal(t, 0)];
    ti[t].mu1 = c->n_states[t] == 2 ? c->chem_pot[formal(t, 1)] : 0.0;
    ti[t].dmu = c->n_states[t] == 2 ? c->chem_pot[formal(t, 1)] - c->chem_pot[formal(t, 0)] : 0.0;
  }
  std::vector<PairInfo> pi(na * na);
  c->max_cut = 0.0;
  int ntab = (int)c->tables.size();
  for (int i = 1; i < na; i++)
    for (int j = 1; j < na; j++) {
      PairInfo &p = pi[i * na + j];
      p.cutsq = c->cutsq[i * nt + j];
      double cut = std::sqrt(p.cutsq);
      if (c->cut_override > 0) cut = std::max(cut, c->cut_override);
      double cn = cut + c->skin;
      p.cutneighsq = cn * cn;
      c->max_cut = std::max(c->max_cut, cut);
      p.ni = c->n_states[i]; p.nj = c->n_states[j];
      for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
          int fa = formal(i, a < p.ni ? a : 0), fb = formal(j, b < p.nj ? b : 0);
          int tix = c->tabindex[fa * nt + fb];
          if (tix < 0 || tix >= ntab) return fail(c, "tabindex refers to a table that was not uploaded");
          p.tab[a * 2 + b] = tix;
        }
    }
  UCG_CHECK(c, c->d_typeinfo.ensure(na));
  UCG_CHECK(c, cudaMemcpy(c->d_typeinfo.p, ti.data(), na * sizeof(TypeInfo), cudaMemcpyHostToDevice));
  UCG_CHECK(c, c->d_pairinfo.ensure(na * na));
  UCG_CHECK(c, cudaMemcpy(c->d_pairinfo.p, pi.data(), na * na * sizeof(PairInfo), cudaMemcpyHostToDevice));
  c->h_pairinfo = pi;
  UCG_CHECK(c, c->d_tables.ensure(std::max(ntab, 1)));
  if (ntab) UCG_CHECK(c, cudaMemcpy(c->d_tables.p, c->tables.data(), ntab * sizeof(TableDev), cudaMemcpyHostToDevice));

  // fast path: a single 2-state actual type whose LINEAR tables share one rsq grid
  c->fast_uniform = false;
  if (c->n_actual == 1 && c->n_states[1] == 2) {
    const PairInfo &p = pi[1 * na + 1];
    const TableDev &t0 = c->tables[p.tab[0]];
    bool ok = t0.style == UCGB200_TAB_LINEAR;
    for (int k = 1; k < 4 && ok; k++) {
      const TableDev &t = c->tables[p.tab[k]];
      ok = t.style == UCGB200_TAB_LINEAR && t.n == t0.n && t.innersq == t0.innersq && t.delta == t0.delta &&
           t.invdelta == t0.invdelta;
    }
    // the (0,1) and (1,0) tables coincide after init_one's symmetrisation
    // (pair_table_ucgld.cpp:892); the interleaved row then holds 3 tables
    if (ok) {
      for (int k = 0; k < 4; k++) c->fast_tab[k] = p.tab[k];
      c->fast_ntab = (p.tab[1] == p.tab[2]) ? 3 : 4;
      c->fast_len = t0.n;
      int n = t0.n, w = c->fast_ntab;
      std::vector<double2> rows((size_t)n * w);
      std::vector<double2> tmp(n);
      int order4[4] = {p.tab[0], p.tab[1], p.tab[2], p.tab[3]};
      int order3[3] = {p.tab[0], p.tab[1], p.tab[3]};
      for (int k = 0; k < w; k++) {
        int tix = (w == 3) ? order3[k] : order4[k];
        UCG_CHECK(c, cudaMemcpy(tmp.data(), c->tables[tix].ef, n * sizeof(double2), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) rows[(size_t)i * w + k] = tmp[i];
      }
      UCG_CHECK(c, c->d_fast_table.ensure(rows.size()));
      UCG_CHECK(c, cudaMemcpy(c->d_fast_table.p, rows.data(), rows.size() * sizeof(double2), cudaMemcpyHostToDevice));
      c->fast_uniform = true;
    }
  }
  c->maps_dirty = false;
  c->list_valid = false;
  return 0;
}

// ------------------------------------------------------------------ atom data
namespace {

__global__ void k_pack_vec3(double4 *dst, const double *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int h = orig[s];
  double4 r = dst[s];
  r.x = src[3 * h]; r.y = src[3 * h + 1]; r.z = src[3 * h + 2];
  dst[s] = r;
}
__global__ void k_pack_w(double4 *dst, const double *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  dst[s].w = src[orig[s]];
}
__global__ void k_pack_scalar_d(double *dst, const double *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) dst[s] = src[orig[s]];
}
__global__ void k_pack_scalar_i(int *dst, const int *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) dst[s] = src[orig[s]];
}
__global__ void k_pack_d2(double2 *dst, const double *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int h = orig[s];
  dst[s] = make_double2(src[2 * h], src[2 * h + 1]);
}
// which: 0 = type (low 16 bits), 1 = state (bit 16)
__global__ void k_pack_ts(int *ts, const int *src, const int *orig, int n, int which) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int v = src[orig[s]];
  int old = ts[s];
  if (which == 0) ts[s] = (old & ~0xffff) | (v & 0xffff);
  else {
    // data_atom_post clamps the state to {0,1} (atom_vec_ucg.cpp:162-163)
    v = v < 0 ? 0 : (v > 1 ? 1 : v);
    ts[s] = (old & 0xffff) | (v << 16);
  }
}
__global__ void k_iota(int *p, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) p[s] = s;
}

__global__ void k_unpack_vec3(double *dst, const double4 *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int h = orig[s];
  double4 r = src[s];
  dst[3 * h] = r.x; dst[3 * h + 1] = r.y; dst[3 * h + 2] = r.z;
}
__global__ void k_unpack_w(double *dst, const double4 *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) dst[orig[s]] = src[s].w;
}
__global__ void k_unpack_scalar_d(double *dst, const double *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) dst[orig[s]] = src[s];
}
__global__ void k_unpack_scalar_i(int *dst, const int *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) dst[orig[s]] = src[s];
}
__global__ void k_unpack_d2(double *dst, const double2 *src, const int *orig, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int h = orig[s];
  dst[2 * h] = src[s].x; dst[2 * h + 1] = src[s].y;
}
// which: 0 type, 1 state, 2 num_ucgstates (n_states_per_type[type], pair_table_ucgld.cpp:173)
__global__ void k_unpack_ts(int *dst, const int *ts, const int *orig, int n, int which, const TypeInfo *ti) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int v = ts[s];
  int out = which == 0 ? (v & 0xffff) : (which == 1 ? ((v >> 16) & 1) : ti[v & 0xffff].nstates);
  dst[orig[s]] = out;
}

}  // namespace

#define GRID1(n) nblocks((n), 256), 256, 0, c->stream

static int ensure_atom_capacity(ucgb200_ctx *c, size_t nall, size_t nloc) {
  cudaStream_t st = c->stream;
  UCG_CHECK(c, c->pos.ensure(nall, true, st));
  UCG_CHECK(c, c->ts.ensure(nall, true, st));
  UCG_CHECK(c, c->ucgp.ensure(nall, true, st));
  UCG_CHECK(c, c->tag.ensure(nall, true, st));
  UCG_CHECK(c, c->mol.ensure(nall, true, st));
  UCG_CHECK(c, c->vel.ensure(nloc, true, st));
  UCG_CHECK(c, c->frc.ensure(nloc, true, st));
  UCG_CHECK(c, c->scores.ensure(nloc, true, st));
  UCG_CHECK(c, c->ucgml.ensure(nloc, true, st));
  UCG_CHECK(c, c->mask.ensure(nloc, true, st));
  UCG_CHECK(c, c->orig.ensure(nloc, true, st));
  UCG_CHECK(c, c->xhold.ensure(nloc, true, st));
  return 0;
}
int ucg_ensure_atom_capacity(ucgb200_ctx *c, size_t nall, size_t nloc) { return ensure_atom_capacity(c, nall, nloc); }

// after_xv (ucgb200_step_host): called once the pack kernels of x and v are queued, before the other fields' packs
static int upload_impl(ucgb200_ctx *c, int nlocal, const ucgb200_atoms *h, unsigned fields, const std::function<int()> *after_xv) {
  if (!c || !h || nlocal < 0) return -1;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  bool fresh = (nlocal != c->nlocal);
  if (fresh) {
    // a new atom count resets the device ordering; every array must come along
    c->nlocal = nlocal; c->nghost = 0; c->list_valid = false; c->maxdisp_valid = false;
    int rc = ensure_atom_capacity(c, (size_t)nlocal + nlocal / 4 + 1024, nlocal);
    if (rc) return rc;
    k_iota<<<GRID1(nlocal)>>>(c->orig.p, nlocal); UCG_LAUNCHED(c);
    UCG_CHECK(c, cudaMemsetAsync(c->pos.p, 0, nlocal * sizeof(double4), st));
    UCG_CHECK(c, cudaMemsetAsync(c->vel.p, 0, nlocal * sizeof(double4), st));
    UCG_CHECK(c, cudaMemsetAsync(c->frc.p, 0, nlocal * sizeof(double4), st));
    UCG_CHECK(c, cudaMemsetAsync(c->scores.p, 0, nlocal * sizeof(double2), st));
    UCG_CHECK(c, cudaMemsetAsync(c->ts.p, 0, nlocal * sizeof(int), st));
    UCG_CHECK(c, cudaMemsetAsync(c->mol.p, 0, nlocal * sizeof(int), st));
    UCG_CHECK(c, cudaMemsetAsync(c->ucgp.p, 0, nlocal * sizeof(double), st));
    UCG_CHECK(c, cudaMemsetAsync(c->ucgml.p, 0, nlocal * sizeof(double), st));
    // default mask = group "all", tags 1..n
    std::vector<int> ones(nlocal, 1);
    UCG_CHECK(c, cudaMemcpyAsync(c->mask.p, ones.data(), nlocal * sizeof(int), cudaMemcpyHostToDevice, st));
    for (int i = 0; i < nlocal; i++) ones[i] = i + 1;
    UCG_CHECK(c, cudaMemcpyAsync(c->tag.p, ones.data(), nlocal * sizeof(int), cudaMemcpyHostToDevice, st));
    UCG_CHECK(c, cudaStreamSynchronize(st));
  }
  if (nlocal == 0) return 0;
  // device staging: every field has its own fixed slot (the same slots the result gathers of ucgb200_step_host use:
  // x 0, ucgl 3n, v 4n, ucgvl 7n, f 8n, ucgforce 11n, ucgp 12n, scores 13n, ucgml 15n doubles; ucgstate 0, type n,
  // mask 2n, tag 3n, molecule 4n ints), so the H2D copies queue back to back (true DMA when the host arrays are
  // pinned) and each pack kernel only waits for its own copy
  size_t n = nlocal;
  UCG_CHECK(c, c->stage_d.ensure(16 * n + 64));
  UCG_CHECK(c, c->stage_i.ensure(6 * n + 64));
  double *sd = c->stage_d.p;
  int *si = c->stage_i.p;
  const int *orig = c->orig.p;
  // the copies run on the upload stream, behind everything the context stream has queued so far (the staging slots
  // may still be read by earlier work); each pack kernel waits for the event of its own copy
  if (!c->stream_ul) UCG_CHECK(c, cudaStreamCreateWithFlags(&c->stream_ul, cudaStreamNonBlocking));
  if (!c->ev_ul_start) UCG_CHECK(c, cudaEventCreateWithFlags(&c->ev_ul_start, cudaEventDisableTiming));
  UCG_CHECK(c, cudaEventRecord(c->ev_ul_start, st));
  UCG_CHECK(c, cudaStreamWaitEvent(c->stream_ul, c->ev_ul_start, 0));
  int nev = 0;
  auto copied = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream_ul);
    if (e != cudaSuccess) return e;
    cudaEvent_t &ev = c->ev_ul[nev++];
    if (!ev && (e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ev, c->stream_ul)) != cudaSuccess) return e;
    return cudaStreamWaitEvent(st, ev, 0);
  };
#define UP_D(slot, ptr, cnt) UCG_CHECK(c, copied(sd + (slot), (ptr), (cnt) * sizeof(double)))
#define UP_I(slot, ptr, cnt) UCG_CHECK(c, copied(si + (slot), (ptr), (cnt) * sizeof(int)))
  if ((fields & UCGB200_F_X) && h->x) { c->maxdisp_valid = false; UP_D(0, h->x, 3 * n); k_pack_vec3<<<GRID1(nlocal)>>>(c->pos.p, sd, orig, nlocal); UCG_LAUNCHED(c); c->list_valid = c->list_valid && !fresh; }
  if ((fields & UCGB200_F_V) && h->v) { UP_D(4 * n, h->v, 3 * n); k_pack_vec3<<<GRID1(nlocal)>>>(c->vel.p, sd + 4 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if (after_xv) { int rc = (*after_xv)(); if (rc) return rc; }
  if ((fields & UCGB200_F_UCGL) && h->ucgl) { UP_D(3 * n, h->ucgl, n); k_pack_w<<<GRID1(nlocal)>>>(c->pos.p, sd + 3 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_UCGVL) && h->ucgvl) { UP_D(7 * n, h->ucgvl, n); k_pack_w<<<GRID1(nlocal)>>>(c->vel.p, sd + 7 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_F) && h->f) { UP_D(8 * n, h->f, 3 * n); k_pack_vec3<<<GRID1(nlocal)>>>(c->frc.p, sd + 8 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_UCGFORCE) && h->ucgforce) { UP_D(11 * n, h->ucgforce, n); k_pack_w<<<GRID1(nlocal)>>>(c->frc.p, sd + 11 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_SCORES) && h->ucgsoftmaxscores) { UP_D(13 * n, h->ucgsoftmaxscores, 2 * n); k_pack_d2<<<GRID1(nlocal)>>>(c->scores.p, sd + 13 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_UCGML) && h->ucgml) { UP_D(15 * n, h->ucgml, n); k_pack_scalar_d<<<GRID1(nlocal)>>>(c->ucgml.p, sd + 15 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_UCGP) && h->ucgp) { UP_D(12 * n, h->ucgp, n); k_pack_scalar_d<<<GRID1(nlocal)>>>(c->ucgp.p, sd + 12 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_TYPE) && h->type) { UP_I(n, h->type, n); k_pack_ts<<<GRID1(nlocal)>>>(c->ts.p, si + n, orig, nlocal, 0); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_UCGSTATE) && h->ucgstate) { UP_I(0, h->ucgstate, n); k_pack_ts<<<GRID1(nlocal)>>>(c->ts.p, si, orig, nlocal, 1); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_MASK) && h->mask) { UP_I(2 * n, h->mask, n); k_pack_scalar_i<<<GRID1(nlocal)>>>(c->mask.p, si + 2 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_TAG) && h->tag) { UP_I(3 * n, h->tag, n); k_pack_scalar_i<<<GRID1(nlocal)>>>(c->tag.p, si + 3 * n, orig, nlocal); UCG_LAUNCHED(c); }
  if ((fields & UCGB200_F_MOLECULE) && h->molecule) { UP_I(4 * n, h->molecule, n); k_pack_scalar_i<<<GRID1(nlocal)>>>(c->mol.p, si + 4 * n, orig, nlocal); UCG_LAUNCHED(c); }
#undef UP_D
#undef UP_I
  return 0;
}

extern "C" int ucgb200_atoms_upload(ucgb200_ctx *c, int nlocal, const ucgb200_atoms *h, unsigned fields) {
  return upload_impl(c, nlocal, h, fields, nullptr);
}

extern "C" int ucgb200_atoms_download(ucgb200_ctx *c, int cap, ucgb200_atoms *h, unsigned fields) {
  if (!c || !h) return -1;
  if (cap < c->nlocal) return fail(c, "atoms_download: host capacity smaller than nlocal");
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  int nlocal = c->nlocal;
  size_t n = nlocal;
  if (n == 0) return 0;
  if ((fields & UCGB200_F_NUMSTATES) && h->num_ucgstates) { int rc = rebuild_maps(c); if (rc) return rc; }
  // every field is gathered into its own slot of a device staging area in host order, all D2H
  // copies are queued behind the gathers, and the stream is synchronised once
  UCG_CHECK(c, c->stage_d.ensure(16 * n + 64));
  UCG_CHECK(c, c->stage_i.ensure(6 * n + 64));
  double *sd = c->stage_d.p;
  int *si = c->stage_i.p;
  const int *orig = c->orig.p;
  struct Copy { void *dst; const void *src; size_t bytes; };
  std::vector<Copy> copies;
#define DN_D(ptr, cnt) do { copies.push_back({(ptr), sd, (cnt) * sizeof(double)}); sd += (cnt); } while (0)
#define DN_I(ptr, cnt) do { copies.push_back({(ptr), si, (cnt) * sizeof(int)}); si += (cnt); } while (0)
  if ((fields & UCGB200_F_X) && h->x) { k_unpack_vec3<<<GRID1(nlocal)>>>(sd, c->pos.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->x, 3 * n); }
  if ((fields & UCGB200_F_UCGL) && h->ucgl) { k_unpack_w<<<GRID1(nlocal)>>>(sd, c->pos.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->ucgl, n); }
  if ((fields & UCGB200_F_V) && h->v) { k_unpack_vec3<<<GRID1(nlocal)>>>(sd, c->vel.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->v, 3 * n); }
  if ((fields & UCGB200_F_UCGVL) && h->ucgvl) { k_unpack_w<<<GRID1(nlocal)>>>(sd, c->vel.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->ucgvl, n); }
  if ((fields & UCGB200_F_F) && h->f) { k_unpack_vec3<<<GRID1(nlocal)>>>(sd, c->frc.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->f, 3 * n); }
  if ((fields & UCGB200_F_UCGFORCE) && h->ucgforce) { k_unpack_w<<<GRID1(nlocal)>>>(sd, c->frc.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->ucgforce, n); }
  if ((fields & UCGB200_F_SCORES) && h->ucgsoftmaxscores) { k_unpack_d2<<<GRID1(nlocal)>>>(sd, c->scores.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->ucgsoftmaxscores, 2 * n); }
  if ((fields & UCGB200_F_UCGML) && h->ucgml) { k_unpack_scalar_d<<<GRID1(nlocal)>>>(sd, c->ucgml.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->ucgml, n); }
  if ((fields & UCGB200_F_UCGP) && h->ucgp) { k_unpack_scalar_d<<<GRID1(nlocal)>>>(sd, c->ucgp.p, orig, nlocal); UCG_LAUNCHED(c); DN_D(h->ucgp, n); }
  if ((fields & UCGB200_F_TYPE) && h->type) { k_unpack_ts<<<GRID1(nlocal)>>>(si, c->ts.p, orig, nlocal, 0, nullptr); UCG_LAUNCHED(c); DN_I(h->type, n); }
  if ((fields & UCGB200_F_UCGSTATE) && h->ucgstate) { k_unpack_ts<<<GRID1(nlocal)>>>(si, c->ts.p, orig, nlocal, 1, nullptr); UCG_LAUNCHED(c); DN_I(h->ucgstate, n); }
  if ((fields & UCGB200_F_NUMSTATES) && h->num_ucgstates) { k_unpack_ts<<<GRID1(nlocal)>>>(si, c->ts.p, orig, nlocal, 2, c->d_typeinfo.p); UCG_LAUNCHED(c); DN_I(h->num_ucgstates, n); }
  if ((fields & UCGB200_F_MASK) && h->mask) { k_unpack_scalar_i<<<GRID1(nlocal)>>>(si, c->mask.p, orig, nlocal); UCG_LAUNCHED(c); DN_I(h->mask, n); }
  if ((fields & UCGB200_F_TAG) && h->tag) { k_unpack_scalar_i<<<GRID1(nlocal)>>>(si, c->tag.p, orig, nlocal); UCG_LAUNCHED(c); DN_I(h->tag, n); }
  if ((fields & UCGB200_F_MOLECULE) && h->molecule) { k_unpack_scalar_i<<<GRID1(nlocal)>>>(si, c->mol.p, orig, nlocal); UCG_LAUNCHED(c); DN_I(h->molecule, n); }
  for (const Copy &cp : copies) UCG_CHECK(c, cudaMemcpyAsync(cp.dst, cp.src, cp.bytes, cudaMemcpyDeviceToHost, st));
  if (!copies.empty()) UCG_CHECK(c, cudaStreamSynchronize(st));
#undef DN_D
#undef DN_I
  return 0;
}

// ucgb200_step_host: the fields of `mask` that the caller asked for and that have not left yet are gathered into
// host order on the context stream (fixed staging slot per field) and copied out on the download stream behind an event,
// so the kernels that follow on the context stream run while the copy engine works.
int ucg_host_out_queue(ucgb200_ctx *c, unsigned mask) {
  ucgb200_atoms *h = c->host_out;
  if (!h) return 0;
  const unsigned fields = c->host_out_fields & mask & ~c->host_out_done;
  if (!fields || c->nlocal == 0) return 0;
  cudaSetDevice(c->device);
  const int nlocal = c->nlocal;
  const size_t n = nlocal;
  if (!c->stream_dl) UCG_CHECK(c, cudaStreamCreateWithFlags(&c->stream_dl, cudaStreamNonBlocking));
  // the brick's population may have grown during the step (migration): cudaFree inside ensure() waits for copies in flight
  UCG_CHECK(c, c->stage_d.ensure(16 * n + 64));
  UCG_CHECK(c, c->stage_i.ensure(6 * n + 64));
  double *sd = c->stage_d.p;
  int *si = c->stage_i.p;
  const int *orig = c->orig.p;
  // slots: x 0, ucgl 3n, v 4n, ucgvl 7n, f 8n, ucgforce 11n, ucgp 12n, scores 13n (doubles); ucgstate 0 (ints).
  // Every field's copy waits for its own gather only.
  int nev = 0;
  auto leave = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
    cudaEvent_t &ev = c->ev_dl[nev++];
    cudaError_t e;
    if (!ev && (e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ev, c->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(c->stream_dl, ev, 0)) != cudaSuccess) return e;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream_dl);
  };
  if ((fields & UCGB200_F_X) && h->x) { k_unpack_vec3<<<GRID1(nlocal)>>>(sd, c->pos.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->x, sd, 3 * n * sizeof(double))); }
  if ((fields & UCGB200_F_UCGL) && h->ucgl) { k_unpack_w<<<GRID1(nlocal)>>>(sd + 3 * n, c->pos.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->ucgl, sd + 3 * n, n * sizeof(double))); }
  if ((fields & UCGB200_F_V) && h->v) { k_unpack_vec3<<<GRID1(nlocal)>>>(sd + 4 * n, c->vel.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->v, sd + 4 * n, 3 * n * sizeof(double))); }
  if ((fields & UCGB200_F_UCGVL) && h->ucgvl) { k_unpack_w<<<GRID1(nlocal)>>>(sd + 7 * n, c->vel.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->ucgvl, sd + 7 * n, n * sizeof(double))); }
  if ((fields & UCGB200_F_F) && h->f) { k_unpack_vec3<<<GRID1(nlocal)>>>(sd + 8 * n, c->frc.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->f, sd + 8 * n, 3 * n * sizeof(double))); }
  if ((fields & UCGB200_F_UCGFORCE) && h->ucgforce) { k_unpack_w<<<GRID1(nlocal)>>>(sd + 11 * n, c->frc.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->ucgforce, sd + 11 * n, n * sizeof(double))); }
  if ((fields & UCGB200_F_UCGP) && h->ucgp) { k_unpack_scalar_d<<<GRID1(nlocal)>>>(sd + 12 * n, c->ucgp.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->ucgp, sd + 12 * n, n * sizeof(double))); }
  if ((fields & UCGB200_F_SCORES) && h->ucgsoftmaxscores) { k_unpack_d2<<<GRID1(nlocal)>>>(sd + 13 * n, c->scores.p, orig, nlocal); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->ucgsoftmaxscores, sd + 13 * n, 2 * n * sizeof(double))); }
  if ((fields & UCGB200_F_UCGSTATE) && h->ucgstate) { k_unpack_ts<<<GRID1(nlocal)>>>(si, c->ts.p, orig, nlocal, 1, nullptr); UCG_LAUNCHED(c); UCG_CHECK(c, leave(h->ucgstate, si, n * sizeof(int))); }
  c->host_out_done |= fields;
  return 0;
}

extern "C" int ucgb200_step_host(ucgb200_ctx *c, const ucgb200_atoms *in, unsigned in_fields, ucgb200_atoms *out, unsigned out_fields) {
  if (!c || !in || !out) return -1;
  if (!c->deck_set) return fail(c, "step_host: deck not configured");
  if (!c->list_valid && c->nlocal == 0) return fail(c, "step_host: no atoms (upload them and call ucgb200_setup first)");
  const unsigned supported = UCGB200_F_X | UCGB200_F_V | UCGB200_F_F | UCGB200_F_UCGSTATE | UCGB200_F_UCGL | UCGB200_F_UCGVL |
                             UCGB200_F_UCGP | UCGB200_F_UCGFORCE | UCGB200_F_SCORES;
  if (out_fields & ~supported) return fail(c, "step_host: out_fields holds a field that a step does not produce");
  cudaSetDevice(c->device);
  // the staging slots of the results lie behind those of the inputs' pack kernels in stream order
  c->host_out = out;
  c->host_out_fields = out_fields;
  c->host_out_done = 0;
  // One brick with an integrator fix: the {x, v} half of this step's initial_integrate runs as soon as x and v have
  // landed, and the new positions start back to the host while lambda, v_lambda and the states are still crossing PCIe
  // in the other direction; the {lambda, v_lambda} half follows their pack kernels (the halves are independent
  // components: bit-identical to the one-kernel stage).  Positions that a rebuild then wraps into the box leave a second
  // time after the build (run.cu).  Opt-in (UCGB200_E2E_SPLIT=1): measured SLOWER at 1 M sites, 3.72 against 3.59 ms per
  // step — with both directions of the link busy each runs below its one-way rate, and the results' copies still end
  // with v, which exists only after the last stage.
  const ucgb200_deck &d = c->deck;
  const bool split = d.nve && c->halo.nranks <= 1 && c->nlocal > 0 && (in_fields & UCGB200_F_X) && in->x && (in_fields & UCGB200_F_V) && in->v &&
                     getenv("UCGB200_E2E_SPLIT") && atoi(getenv("UCGB200_E2E_SPLIT")) == 1;
  const int gb = d.nve_groupbit ? d.nve_groupbit : 1;
  const double dtv = c->dt, dtf = 0.5 * c->dt * c->ftm2v;
  int rc;
  if (split) {
    const std::function<int()> after_xv = [&]() -> int {
      int r = ucg_nve_initial_part(c, dtv, dtf, gb, d.nve == 2, 1);
      if (!r) r = ucg_host_out_queue(c, UCGB200_F_X);
      return r;
    };
    rc = upload_impl(c, c->nlocal, in, in_fields, &after_xv);
    if (!rc) rc = ucg_nve_initial_part(c, dtv, dtf, gb, d.nve == 2, 2);
    c->skip_initial_once = !rc;
  } else
    rc = ucgb200_atoms_upload(c, c->nlocal, in, in_fields);
  if (rc) { c->host_out = nullptr; c->skip_initial_once = false; return rc; }
  rc = ucgb200_run(c, 1);
  c->skip_initial_once = false;
  if (!rc) rc = ucg_host_out_queue(c, ~0u);     // everything that could not leave earlier
  c->host_out = nullptr;
  if (c->stream_dl) {
    cudaError_t e = cudaStreamSynchronize(c->stream_dl);
    if (!rc && e != cudaSuccess) { c->err = std::string("step_host download: ") + cudaGetErrorString(e); rc = -2; }
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (!rc && e != cudaSuccess) { c->err = std::string("step_host: ") + cudaGetErrorString(e); rc = -2; }
  return rc;
}

extern "C" int ucgb200_natoms(const ucgb200_ctx *c, int *nlocal, int *nghost) {
  if (!c) return -1;
  if (nlocal) *nlocal = c->nlocal;
  if (nghost) *nghost = c->nghost;
  return 0;
}

extern "C" int ucgb200_force_clear(ucgb200_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  if (c->nlocal == 0) return 0;
  UCG_CHECK(c, cudaMemsetAsync(c->frc.p, 0, (size_t)c->nlocal * sizeof(double4), c->stream));
  UCG_CHECK(c, cudaMemsetAsync(c->scores.p, 0, (size_t)c->nlocal * sizeof(double2), c->stream));
  return 0;
}

static int read_status(ucgb200_ctx *c, int *code, int *tag_i, int *tag_j, double *rsq, bool clear) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  ErrWord w;
  UCG_CHECK(c, cudaMemcpyAsync(&w, c->d_err.p, sizeof(ErrWord), cudaMemcpyDeviceToHost, c->stream));
  UCG_CHECK(c, cudaStreamSynchronize(c->stream));
  if (code) *code = w.code;
  if (tag_i) *tag_i = w.tag_i;
  if (tag_j) *tag_j = w.tag_j;
  if (rsq) *rsq = w.rsq;
  if (w.code && clear) {
    // sticky until read: clear so that a corrected run can continue
    UCG_CHECK(c, cudaMemsetAsync(c->d_err.p, 0, sizeof(ErrWord), c->stream));
  }
  return w.code;
}
extern "C" int ucgb200_status(ucgb200_ctx *c, int *code, int *tag_i, int *tag_j, double *rsq) {
  return read_status(c, code, tag_i, tag_j, rsq, true);
}
extern "C" int ucgb200_status_peek(ucgb200_ctx *c, int *code, int *tag_i, int *tag_j, double *rsq) {
  return read_status(c, code, tag_i, tag_j, rsq, false);
}
