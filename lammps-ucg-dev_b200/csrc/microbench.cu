// microbench.cu — floors of the table_ucgld pair kernel, measured in isolation (sm_100a).
//
// DESIGN.md §4 argues that the pair kernel is bound by the SM's shared-memory data pipe and, behind it, by the FP64
// pipe.  These micro-kernels replay ONE ingredient of a 1 M-site launch each, with the launch shape of the real
// kernel (148 persistent CTAs x 768 threads, the 192 KB interleaved table resident in shared memory), so that the
// floors are measurements and not estimates:
//   lds_random    the 6 x LDS.128 table reads of every in-cutoff visit (rows it, it+1 of the 3 interleaved tables),
//                 row indices drawn from the liquid's rsq distribution (density ~ sqrt(rsq) on [0.8, 6.25])
//   lds_ordered   the same reads with the 32 lanes of a warp on 32 consecutive rows (no bank-group conflicts): the
//                 pipe's ideal
//   fp64          the FP64 arithmetic of a launch (distance, index, 6 interpolations, lambda mixing, accumulation) on
//                 register operands, no memory
//   red_global    the j-side scatter a Newton's-third-law kernel needs: 6 red.global.add.f64 per pair into
//                 site records inside a window of 54 k sites (two z-planes of cells)
//   atoms_cas     the same 6 adds into a shared-memory accumulator tile: atomicAdd(double) on shared memory is a
//                 CAS loop on sm_100a (SASS: ATOMS.CAST.SPIN.64), also for 64-bit integers
//   bulk_red      one 48-byte cp.reduce.async.bulk.add.f64 (SASS: UBLKRED) per pair from a shared-memory staging slot
// Units replayed: 55 M in-cutoff visits (full list) or 27.5 M pairs (half list) — the counts of the 1 000 188-site
// liquid.  Prints one JSON line per kernel; scripts/microbench.sh stores them under profiles/.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int TABLEN = 4096, W = 3, BS = 768;
constexpr size_t TABLE_BYTES = (size_t)TABLEN * W * sizeof(double2);

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// ------------------------------------------------------------------ table reads
// its: precomputed row indices (uint16), n_per_thread consecutive-by-lane entries per thread
template <int MODE>
__global__ void __launch_bounds__(BS) k_lds(const double2 *__restrict__ table, const unsigned short *__restrict__ its,
                                            int per_thread, double *__restrict__ out) {
  extern __shared__ double2 s_tab[];
  for (int k = threadIdx.x; k < TABLEN * W; k += BS) s_tab[k] = table[k];
  __syncthreads();
  const size_t t = (size_t)blockIdx.x * BS + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * BS;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int lane = threadIdx.x & 31;
  for (int n = 0; n < per_thread; n++) {
    int it;
    if (MODE == 0) it = its[(size_t)n * nthreads + t];                       // coalesced 2-byte reads
    else if (MODE == 1) it = ((its[(size_t)n * nthreads + t - lane] + lane) & (TABLEN - 2));   // 32 consecutive rows
    else it = its[(size_t)n * nthreads + t - lane] & (TABLEN - 2);           // one row per warp (broadcast)
    const double2 *r0 = s_tab + it * W;
    const double2 x0 = r0[0], x1 = r0[1], x2 = r0[2], x3 = r0[3], x4 = r0[4], x5 = r0[5];
    a0 += x0.x + x3.y; a1 += x1.x + x4.y; a2 += x2.x + x5.y; a3 += x0.y + x1.y + x2.y + x3.x + x4.x + x5.x;
  }
  out[t] = a0 + a1 + a2 + a3;
}

// Conflict-free schedule of the same reads ("rejected on paper"; measured here).  In
// instruction k the l-th lane of a quarter-warp reads the slot of its 96-byte window (slots w .. w+5, w = 3*it) that
// lies in bank group (k + l) mod 8: eight instructions instead of six, two of them idle per lane, every wavefront
// conflict-free.  The loaded values arrive rotated by r = (w - l) mod 8 and are routed back to window order with a
// three-stage barrel rotation of the eight double2 registers (96 32-bit selects per visit).
__global__ void __launch_bounds__(BS) k_lds_rotated(const double2 *__restrict__ table, const unsigned short *__restrict__ its,
                                                    int per_thread, double *__restrict__ out) {
  extern __shared__ double2 s_tab[];
  for (int k = threadIdx.x; k < TABLEN * W; k += BS) s_tab[k] = table[k];
  __syncthreads();
  const size_t t = (size_t)blockIdx.x * BS + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * BS;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int l = threadIdx.x & 7;
  for (int n = 0; n < per_thread; n++) {
    const int it = its[(size_t)n * nthreads + t];
    const int w = it * W;
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int o = (k + l - w) & 7;                 // window offset whose bank group is (k + l) mod 8
      v[k] = o < 6 ? s_tab[w + o] : make_double2(0.0, 0.0);
    }
    // v[k] holds window offset (k - r) mod 8, r = (w - l) mod 8: rotate left by r so that x[o] = window offset o
    const int r = (w - l) & 7;
#pragma unroll
    for (int stage = 0; stage < 3; stage++) {
      const int sh = 1 << stage;
      const bool take = (r & sh) != 0;          // selects, not a divergent branch
      double2 u[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        u[k].x = take ? v[(k + sh) & 7].x : v[k].x;
        u[k].y = take ? v[(k + sh) & 7].y : v[k].y;
      }
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = u[k];
    }
    a0 += v[0].x + v[3].y; a1 += v[1].x + v[4].y; a2 += v[2].x + v[5].y; a3 += v[0].y + v[1].y + v[2].y + v[3].x + v[4].x + v[5].x;
  }
  out[t] = a0 + a1 + a2 + a3;
}

// ------------------------------------------------------------------ FP64 arithmetic of one launch
// per in-cutoff visit: dx,dy,dz (3 DADD), rsq (3 DMUL 2 DADD), index (1 DADD 1 DMUL, cvt), rsq_it (I2F, DMUL, DADD),
// frac (DADD DMUL), 6 x (DADD + DFMA), bj (DADD), A B FA FB (4 x (DMUL + DFMA)), accA accB (2 DADD),
// fpair (DMUL DFMA), S0 S1 (2 DADD + selects), f (3 DFMA)  ->  the instruction mix of k_pair_ucgld_fast
__global__ void __launch_bounds__(BS) k_fp64(int per_thread, double innersq, double invdelta, double delta, double *__restrict__ out) {
  const size_t t = (size_t)blockIdx.x * BS + threadIdx.x;
  unsigned h = hash32((unsigned)t);
  double rx = 1e-3 * (h & 1023), ry = 1e-3 * ((h >> 10) & 1023), rz = 1e-3 * ((h >> 20) & 1023);
  const double li = 0.3 + 1e-4 * (h & 255), ai = 1.0 - li;
  double fx = 0, fy = 0, fz = 0, accA = 0, accB = 0, S0 = 0, S1 = 0;
  // six "table" values that change every iteration (registers only)
  double e0 = 0.11, e1 = 0.23, e2 = 0.37, f0 = 1.1, f1 = 1.3, f2 = 1.7;
  double jx = 0.9, jy = 0.4, jz = 0.2, lj = 0.6;
  for (int n = 0; n < per_thread; n++) {
    jx += 1e-4; jy -= 2e-4; jz += 3e-4; lj += 1e-5;
    const double dx = rx - jx, dy = ry - jy, dz = rz - jz;
    const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const int it = (int)__dmul_rn(__dadd_rn(rsq, -innersq), invdelta);
    const double rsq_it = __dadd_rn(innersq, __dmul_rn((double)it, delta));
    const double frac = (rsq - rsq_it) * invdelta;
    // "loaded" rows: derived from the running values so nothing can be hoisted
    const double b0 = e0 + 1e-3 * frac, b1 = e1 - 1e-3 * frac, b2 = e2 + 2e-3 * frac;   // stand-ins for row it+1 (not counted)
    const double u00 = e0 + frac * (b0 - e0), f00 = f0 + frac * (b1 - f0);
    const double u01 = e1 + frac * (b1 - e1), f01 = f1 + frac * (b2 - f1);
    const double u11 = e2 + frac * (b2 - e2), f11 = f2 + frac * (b0 - f2);
    const double bj = 1.0 - lj;
    const double A = bj * u00 + lj * u01, B = bj * u01 + lj * u11;
    const double FA = bj * f00 + lj * f01, FB = bj * f01 + lj * f11;
    accA += A; accB += B;
    const double fpair = ai * FA + li * FB;
    const bool s1 = it & 1;
    S0 += s1 ? u01 : u00;
    S1 += s1 ? u11 : u01;
    fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
    e0 = u01 * 0.999; e1 = u11 * 1.001; e2 = u00; f0 = f01; f1 = f11 * 0.999; f2 = f00;
  }
  out[t] = fx + fy + fz + accA + accB + S0 + S1;
}

// ------------------------------------------------------------------ j-side scatter, three ways
// window: the half-list partners of the sites a CTA works on lie within ~2 z-planes of cells = 54 k sites
__global__ void __launch_bounds__(BS) k_red_global(double *__restrict__ acc, int nsites, int window, int per_thread, int ncomp) {
  const size_t t = (size_t)blockIdx.x * BS + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * BS;
  const int base0 = (int)((double)t / (double)nthreads * (double)(nsites - window));
  for (int n = 0; n < per_thread; n++) {
    const int j = base0 + (int)(hash32((unsigned)(t * 977u + n)) % (unsigned)window);
    double *p = acc + (size_t)j * 6;
    const double v = 1e-3 * n;
    for (int k = 0; k < ncomp; k++) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + k), "d"(v) : "memory");
  }
}

template <bool FIXED>
__global__ void __launch_bounds__(BS) k_atoms(double *__restrict__ out, int tile_sites, int per_thread) {
  extern __shared__ double s_acc[];
  for (int k = threadIdx.x; k < tile_sites * 6; k += BS) s_acc[k] = 0.0;
  __syncthreads();
  const size_t t = (size_t)blockIdx.x * BS + threadIdx.x;
  for (int n = 0; n < per_thread; n++) {
    const int j = (int)(hash32((unsigned)(t * 977u + n)) % (unsigned)tile_sites);
    const double v = 1e-3 * n;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      if (FIXED) atomicAdd(reinterpret_cast<unsigned long long *>(s_acc) + j * 6 + k, (unsigned long long)(long long)(v * 1099511627776.0));
      else atomicAdd(s_acc + j * 6 + k, v);
    }
  }
  __syncthreads();
  double s = 0;
  for (int k = threadIdx.x; k < tile_sites * 6; k += BS) s += s_acc[k];
  out[t] = s;
}

__global__ void __launch_bounds__(BS) k_bulk_red(double *__restrict__ acc, int nsites, int window, int per_thread) {
  extern __shared__ __align__(128) double s_stage[];      // 48 B per thread
  const size_t t = (size_t)blockIdx.x * BS + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * BS;
  const int base0 = (int)((double)t / (double)nthreads * (double)(nsites - window));
  double *mine = s_stage + threadIdx.x * 6;
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(mine);
  for (int n = 0; n < per_thread; n++) {
    const int j = base0 + (int)(hash32((unsigned)(t * 977u + n)) % (unsigned)window);
    // the previous reduction must have read the slot before it is overwritten
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    const double v = 1e-3 * n;
#pragma unroll
    for (int k = 0; k < 6; k++) mine[k] = v + k;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 48;" ::"l"(acc + (size_t)j * 6), "r"(saddr) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ host
template <class F>
static double time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f, sum = 0;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
    sum += ms;
  }
  CK(cudaGetLastError());
  (void)best;
  return sum / reps;
}

int main(int argc, char **argv) {
  const double visits_M = argc > 1 ? atof(argv[1]) : 55.0;     // in-cutoff visits of a full-list launch (1 000 188 sites)
  const double pairs_M = visits_M / 2;                         // in-cutoff pairs of a half-list launch
  int dev = 0, nsm = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int grid = nsm;
  const size_t nthreads = (size_t)grid * BS;
  const int per_visit = (int)std::ceil(visits_M * 1e6 / nthreads), per_pair = (int)std::ceil(pairs_M * 1e6 / nthreads);
  const double visits = (double)per_visit * nthreads, pairs = (double)per_pair * nthreads;

  // table + realistic row indices
  std::vector<double2> h_tab((size_t)TABLEN * W);
  for (size_t k = 0; k < h_tab.size(); k++) h_tab[k] = make_double2(1e-3 * (k % 977), 1e-3 * (k % 613));
  const double innersq = 0.25, outersq = 6.25, delta = (outersq - innersq) / (TABLEN - 1), invdelta = 1.0 / delta;
  std::vector<unsigned short> h_its((size_t)per_visit * nthreads);
  {
    uint64_t s = 88172645463325252ull;
    const double lo = std::pow(0.8, 1.5), hi = std::pow(6.25, 1.5);
    for (auto &v : h_its) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const double u = (double)(s >> 11) / 9007199254740992.0;
      const double rsq = std::pow(lo + u * (hi - lo), 2.0 / 3.0);
      int it = (int)((rsq - innersq) * invdelta);
      v = (unsigned short)(it < TABLEN - 1 ? it : TABLEN - 2);
    }
  }
  double2 *d_tab;
  unsigned short *d_its;
  double *d_out, *d_acc;
  const int nsites = 1000188, window = 54000;
  CK(cudaMalloc(&d_tab, TABLE_BYTES));
  CK(cudaMalloc(&d_its, h_its.size() * sizeof(unsigned short)));
  CK(cudaMalloc(&d_out, nthreads * sizeof(double)));
  CK(cudaMalloc(&d_acc, (size_t)nsites * 6 * sizeof(double)));
  CK(cudaMemset(d_acc, 0, (size_t)nsites * 6 * sizeof(double)));
  CK(cudaMemcpy(d_tab, h_tab.data(), TABLE_BYTES, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_its, h_its.data(), h_its.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));

  CK(cudaFuncSetAttribute(k_lds<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TABLE_BYTES));
  CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TABLE_BYTES));
  CK(cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TABLE_BYTES));
  const char *fmt = "{\"kernel\": \"%s\", \"ms\": %.4f, \"units\": %.0f, \"unit\": \"%s\", \"grid\": %d, \"block\": %d, \"note\": \"%s\"}\n";

  double ms = time_ms([&] { k_lds<0><<<grid, BS, TABLE_BYTES>>>(d_tab, d_its, per_visit, d_out); });
  printf(fmt, "lds_random", ms, visits, "in-cutoff visits", grid, BS, "6 x LDS.128 per visit, rows from the liquid's rsq distribution");
  ms = time_ms([&] { k_lds<1><<<grid, BS, TABLE_BYTES>>>(d_tab, d_its, per_visit, d_out); });
  printf(fmt, "lds_ordered", ms, visits, "in-cutoff visits", grid, BS, "same reads, lanes on consecutive rows: no bank-group conflicts");
  ms = time_ms([&] { k_lds<2><<<grid, BS, TABLE_BYTES>>>(d_tab, d_its, per_visit, d_out); });
  printf(fmt, "lds_broadcast", ms, visits, "in-cutoff visits", grid, BS, "same reads, one row per warp");
  CK(cudaFuncSetAttribute(k_lds_rotated, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TABLE_BYTES));
  ms = time_ms([&] { k_lds_rotated<<<grid, BS, TABLE_BYTES>>>(d_tab, d_its, per_visit, d_out); });
  printf(fmt, "lds_rotated", ms, visits, "in-cutoff visits", grid, BS, "conflict-free rotated schedule: 8 x LDS.128 + 3-stage register rotation (96 selects) per visit");
  ms = time_ms([&] { k_fp64<<<grid, BS>>>(per_visit, innersq, invdelta, delta, d_out); });
  printf(fmt, "fp64", ms, visits, "in-cutoff visits", grid, BS, "FP64 arithmetic of a visit on register operands");
  ms = time_ms([&] { k_fp64<<<grid, BS>>>(per_pair, innersq, invdelta, delta, d_out); });
  printf(fmt, "fp64_half", ms, pairs, "pairs", grid, BS, "the same for a half list");
  for (int ncomp : {6, 3, 1}) {
    ms = time_ms([&] { k_red_global<<<grid, BS>>>(d_acc, nsites, window, per_pair, ncomp); });
    char name[64];
    snprintf(name, sizeof name, "red_global_%d", ncomp);
    printf(fmt, name, ms, pairs, "pairs", grid, BS, "red.global.add.f64 per pair into 48-byte site records, 54 k-site window");
  }
  {
    const int tile = 600;
    const size_t smem = (size_t)tile * 6 * sizeof(double);
    ms = time_ms([&] { k_atoms<false><<<grid, BS, smem>>>(d_out, tile, per_pair); });
    printf(fmt, "atoms_cas_f64", ms, pairs, "pairs", grid, BS, "6 atomicAdd(double) on shared memory per pair (ATOMS.CAST.SPIN.64), 600-site tile");
    ms = time_ms([&] { k_atoms<true><<<grid, BS, smem>>>(d_out, tile, per_pair); });
    printf(fmt, "atoms_cas_u64", ms, pairs, "pairs", grid, BS, "6 atomicAdd(u64 fixed point) on shared memory per pair (also a CAS loop)");
  }
  {
    const size_t smem = (size_t)BS * 48;
    const int pp = per_pair > 64 ? 64 : per_pair;    // bounded: small bulk operations may be very slow
    ms = time_ms([&] { k_bulk_red<<<grid, BS, smem>>>(d_acc, nsites, window, pp); }, 3);
    printf(fmt, "bulk_red_48B", ms * (double)per_pair / pp, pairs, "pairs", grid, BS,
           "one 48-byte cp.reduce.async.bulk.add.f64 per pair (UBLKRED), time scaled from a bounded sample");
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
