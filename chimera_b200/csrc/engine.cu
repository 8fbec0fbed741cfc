// engine.cu -- device-resident PIC engine: particles (structure of arrays), grids, spectral state
// and operator tables live in HBM; one call runs the reference's step sequence on the device.
//
// Stage order and semantics follow the reference's Python driver, which stays the specification:
//   make_step      moduls/chimera_main.py:82-92      make_halfstep   :61-80
//   dep_curr/dep_dens/dep_bg  :153-248                project_fields  :130-151
//   fb_curr_in/fb_dens_in/FBGradDens/poiss_corr/maxwell_solver/G2B_FBRot/fb_fld_out
//                  moduls/solvers.py:407,422,517,301,281,536,450
//   chunk_and_damp moduls/species.py:351-398 (re-binning; here a device radix sort by
//                  (x-chunk, r-cell, x-cell) which refines the reference's chunk order)
#include <cub/device/device_radix_sort.cuh>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include "../../include/chimera_b200.h"
#include "fbops.cuh"

namespace chb {

extern long long g_h2d_bytes, g_d2h_bytes;  // api_host.cu: host<->device traffic counters

struct NamedArray {
  void* p = nullptr;
  size_t bytes = 0;  // the array proper (what upload / download move)
  size_t alloc = 0;  // allocated: J, Rho, EB, EB_slab carry spare columns for the column-block dataflow
};

struct Species {
  double *x = nullptr, *xh = nullptr, *p = nullptr, *w = nullptr;       // SoA, component stride = cap
  double *x2 = nullptr, *xh2 = nullptr, *p2 = nullptr, *w2 = nullptr;   // permutation targets
  i64 np = 0, cap = 0;
  double push_fact = 0.0;
  int still = 0;
  int* d_ind = nullptr;          // IndInChunk(0:nchnk); the last entry is the number of particles kept
  std::vector<int> h_ind;
  int* d_cta = nullptr;          // prefix of binned-deposit CTA counts per chunk (0:nchnk)
  int ncta = 0;
  std::vector<int> h_cta;        // host copy of d_cta
  int* d_cta_f = nullptr;        // the same for the fused kernel (kFusedNPB particles per CTA)
  int ncta_f = 0;
  i64 tile_w = 0;                // x cells per re-binning tile of the current particle order (0: unsorted)
  std::vector<DeviceSpec> devs;  // external-field devices of this species (species.py:55 Args['Devices'])
  std::vector<double*> dev_maps; // device copies of their maps (owned)
};

}  // namespace chb

using namespace chb;

constexpr size_t kColPad = 64;  // spare columns of J / Rho / EB (column blocks of up to 64 ranks)

struct chimera_engine {
  chimera_engine_config cfg;
  int col_world = 0;  // > 0: column-block dataflow set up for this many ranks (chimera_engine_set_colflow)
  int col_rank = 0;   // this rank's column block
  cudaStream_t st = nullptr;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // copy streams of chimera_engine_step_host
  std::vector<cudaEvent_t> host_evs;
  double *host_EG = nullptr, *host_G = nullptr, *host_mom = nullptr;  // between step_host_begin / _end
  int host_id = -1;
  int host_mid_done = 0;
  int fuse = 1;  // use the fused particle kernel inside multi-step calls (chimera_engine_set_fuse)
  // One fused step (particle kernel + the whole spectral update, ~50 launches) captured as a CUDA graph and replayed
  // for the steps between two re-binnings: on the demo-size grids a step is a few hundred microseconds and the launch
  // gaps are a fifth of it.  Two graphs: gradRho_fb_prv / _nxt swap every step (ph_fb_in_rho).
  struct StepGraph { cudaGraphExec_t exec; const void* nxt; long long launches; };
  std::vector<StepGraph> graphs;
  unsigned long long graph_sig = 0;
  unsigned long long graph_warm_gen = ~0ull;  // scratch generation under which a fused step last ran eagerly
  int graph_state = 0;  // 1: graphs in use, -1: a capture failed, graphs off for this engine
  int use_graph = 1;    // chimera_engine_set_graph
  std::vector<double> static_px;  // mean momentum per species for the next static_fields phase (multi-rank runs)
  double dev_time = 0.0;  // i_step * TimeStep seen by time-dependent devices in the next gather + push
  // the gather + push that closes the last step of a chimera_engine_step call is left pending, so that the next call
  // can run it inside its first fused kernel; every other entry point that reads or changes engine state completes it
  // first (ENG_ENTER)
  bool tail_pending = false;
  double tail_time = 0.0;  // dev_time of the pending gather + push
  int lazy_tail = 1;       // chimera_engine_set_lazy_tail
  // window that moves every step inside chimera_engine_step (chimera_engine_set_window): shift at stage 1 (before
  // push_coords) and at stage 2 (between dep_curr and dep_dens), chimera_main.py:286-302
  double win_s1 = 0.0, win_s2 = 0.0;
  double leftX_J = NAN;  // leftX the deposited J lives on, consumed by the next fb_in_J (NaN: cfg.leftX)
  double host_rho_from_bg = 1.0;  // 0 on the ranks that must not add BckGrndRho before an all-reduce
  bool own_stream = false;
  Scratch scr;
  FFTCache fft;
  std::map<std::string, NamedArray> arr;
  std::vector<Species> sp;
  PackedOps pInCurr, pOut, pDp, pDm;
  double* packed = nullptr;
  bool ops_dirty = true;
  // sort scratch
  unsigned *key_a = nullptr, *key_b = nullptr;
  int *idx_a = nullptr, *idx_b = nullptr;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  i64 sort_cap = 0;
  // profiling
  int profile = 0;
  double ms[CHB_NPHASES];
  long long calls[CHB_NPHASES];
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
  std::vector<cudaEvent_t> ev_pool;

  cd* A(const char* n) { return (cd*)arr[n].p; }
  double* D(const char* n) { return (double*)arr[n].p; }
};

namespace chb {

// ------------------------------------------------------------------------------------------
// layout conversion between the reference's (ncomp, np) Fortran arrays and the engine's SoA
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aos_to_soa_k(double* __restrict__ dst, const double* __restrict__ src, int ncomp,
                                                    i64 cap, i64 np) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= np * ncomp) return;
  const i64 ip = e / ncomp;
  const int c = (int)(e - ip * ncomp);
  dst[c * cap + ip] = src[e];
}
__global__ void __launch_bounds__(256) soa_to_aos_k(double* __restrict__ dst, const double* __restrict__ src, int ncomp,
                                                    i64 cap, i64 np) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= np * ncomp) return;
  const i64 ip = e / ncomp;
  const int c = (int)(e - ip * ncomp);
  dst[e] = src[c * cap + ip];
}

// ------------------------------------------------------------------------------------------
// re-binning keys: (x-chunk as particle_tools.f90:183, x-tile, r-cell, x-cell in tile); kdrop = leaves
// the domain.  Consecutive particles then share a cell (deposit_runs_k) and a CTA's particles sit in a
// compact (x, r) box (gather_push_tiled_k).
// ------------------------------------------------------------------------------------------
struct BinSpec {
  double x0, chunk_inv;       // Xgrid(0), 1 / chunk length
  double l0, l1, l2, l3;      // SimDom
  double leftX, dx_inv, r0, dr_inv;
  int nchnk;
  i64 nx, nrc;                // x cells, r cells
  i64 cs, tile_w, ntile;      // x cells per chunk, per x-tile, x-tiles per chunk
  unsigned kdrop;             // key of a particle that leaves the domain (sorts last)
};

__global__ void __launch_bounds__(256) bin_keys_k(const double* __restrict__ x, i64 cap, unsigned* __restrict__ key,
                                                  int* __restrict__ idx, BinSpec b, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const double xp = x[ip], yp = x[cap + ip], zp = x[2 * cap + ip];
  const double r2 = yp * yp + zp * zp;
  unsigned k = b.kdrop;
  if (xp >= b.l0 && xp <= b.l1 && r2 >= b.l2 && r2 <= b.l3) {
    i64 c = (i64)floor((xp - b.x0) * b.chunk_inv);
    c = c < 0 ? 0 : (c > b.nchnk - 1 ? b.nchnk - 1 : c);
    i64 ix = (i64)floor((xp - b.leftX) * b.dx_inv);
    ix = ix < 0 ? 0 : (ix > b.nx - 1 ? b.nx - 1 : ix);
    i64 ir = (i64)floor((sqrt(r2) - b.r0) * b.dr_inv);
    ir = ir < 0 ? 0 : (ir > b.nrc - 1 ? b.nrc - 1 : ir);
    i64 lx = ix - c * b.cs;  // the chunk id follows the reference's float rule; keep the cell inside it
    lx = lx < 0 ? 0 : (lx > b.cs - 1 ? b.cs - 1 : lx);
    const i64 xt = lx / b.tile_w;
    k = (unsigned)((((c * b.ntile + xt) * b.nrc + ir) * b.tile_w) + (lx - xt * b.tile_w));
  }
  key[ip] = k;
  idx[ip] = (int)ip;
}

// IndInChunk(c) = first sorted position whose key belongs to chunk >= c; IndInChunk(nchnk) = particles
// kept (dropped particles carry the key nchnk*per_chunk and sort last)
__global__ void chunk_bounds_k(const unsigned* __restrict__ key, int* __restrict__ ind, int nchnk, i64 per_chunk,
                               i64 np) {
  const int c = threadIdx.x;
  if (c > nchnk) return;
  const unsigned long long target = (unsigned long long)c * (unsigned long long)per_chunk;
  i64 lo = 0, hi = np;
  while (lo < hi) {
    const i64 mid = (lo + hi) >> 1;
    if ((unsigned long long)key[mid] < target) lo = mid + 1; else hi = mid;
  }
  ind[c] = (int)lo;
}

__global__ void __launch_bounds__(256) permute_soa_k(double* __restrict__ dx, double* __restrict__ dxh,
                                                     double* __restrict__ dp, double* __restrict__ dw,
                                                     const double* __restrict__ sx, const double* __restrict__ sxh,
                                                     const double* __restrict__ sp, const double* __restrict__ sw,
                                                     const int* __restrict__ idx, i64 cap, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const i64 j = idx[ip];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dx[c * cap + ip] = sx[c * cap + j];
    dxh[c * cap + ip] = sxh[c * cap + j];
    dp[c * cap + ip] = sp[c * cap + j];
  }
  dw[ip] = sw[j];
}

__global__ void __launch_bounds__(256) add_cd_k(cd* __restrict__ a, const cd* __restrict__ b, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) a[e] = cadd(a[e], b[e]);
}

}  // namespace chb

// ==========================================================================================
namespace {

#define ENG_CHECK(e) \
  if (!(e)) { set_error("null engine handle"); return 2; }
// entry points that observe or change the state a pending gather + push would touch complete it first
#define ENG_ENTER(e) \
  ENG_CHECK(e);      \
  if ((e)->tail_pending) CHB_TRY(flush_tail(e));
int flush_tail(chimera_engine* e);

int alloc_named(chimera_engine* e, const char* name, size_t bytes, bool zero = true, size_t spare = 0) {
  NamedArray a;
  a.bytes = bytes;
  a.alloc = bytes + spare;
  CHB_CUDA(cudaMalloc(&a.p, a.alloc ? a.alloc : 16));
  if (zero) CHB_CUDA(cudaMemsetAsync(a.p, 0, a.alloc ? a.alloc : 16, e->st));
  e->arr[name] = a;
  return 0;
}

GridGeom geom(chimera_engine* e) {
  const auto& c = e->cfg;
  GridGeom g;
  g.leftX = c.leftX; g.dx_inv = 1.0 / c.dx; g.dr_inv = 1.0 / c.dr; g.kx0 = c.kx0;
  g.r0 = -0.5 * c.dr;  // overwritten from the uploaded Rgrid in ensure_ops
  g.rmax = 0.0;
  g.Rgrid = e->D("Rgrid");
  g.nxn = c.nx; g.nrn = c.nrn; g.nm = c.nm;
  return g;
}

struct EngineHost {  // host copies of the few table entries the launch code needs
  double r0 = 0, rmax = 0;
};
std::map<chimera_engine*, EngineHost> g_host;

int ensure_ops(chimera_engine* e) {
  if (!e->ops_dirty) return 0;
  const auto& c = e->cfg;
  const i64 nr = c.nrn - 1;
  const int nd = (int)(c.env ? c.nm + 2 : c.nm + 1);
  const size_t b_in = packed_ops_bytes(nr, c.nkr, (int)c.nm), b_out = packed_ops_bytes(c.nkr, nr, (int)c.nm);
  const size_t b_d = packed_ops_bytes(c.nkr, c.nkr, nd);
  if (!e->packed) CHB_CUDA(cudaMalloc((void**)&e->packed, b_in + b_out + 2 * b_d));
  double* q = e->packed;
  CHB_TRY(pack_ops(e->st, e->pInCurr, q, e->D("InCurr"), nr, c.nkr, (int)c.nm)); q += b_in / sizeof(double);
  CHB_TRY(pack_ops(e->st, e->pOut, q, e->D("Out"), c.nkr, nr, (int)c.nm)); q += b_out / sizeof(double);
  CHB_TRY(pack_ops(e->st, e->pDp, q, e->D("DpS2S"), c.nkr, c.nkr, nd)); q += b_d / sizeof(double);
  CHB_TRY(pack_ops(e->st, e->pDm, q, e->D("DmS2S"), c.nkr, c.nkr, nd));
  double ends[2];
  CHB_CUDA(cudaMemcpyAsync(&ends[0], e->D("Rgrid"), sizeof(double), cudaMemcpyDeviceToHost, e->st));
  CHB_CUDA(cudaMemcpyAsync(&ends[1], e->D("Rgrid") + (c.nrn - 1), sizeof(double), cudaMemcpyDeviceToHost, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  g_host[e].r0 = ends[0];
  g_host[e].rmax = ends[1];
  e->ops_dirty = false;
  return 0;
}

GridGeom geom_ready(chimera_engine* e) {
  GridGeom g = geom(e);
  g.r0 = g_host[e].r0;
  g.rmax = g_host[e].rmax;
  return g;
}

FBCtx fbctx(chimera_engine* e) { return FBCtx{e->st, &e->scr, &e->fft}; }
// kx rows of the spectral state held by this engine: all of them, or one slab of mirror pairs when the
// spectral solve is sharded over the ranks (chimera_b200/sharding.py)
inline bool slab(const chimera_engine* e) { return e->cfg.nx_slab > 0 && e->cfg.nx_slab < e->cfg.nx; }
inline i64 nxs(const chimera_engine* e) { return slab(e) ? e->cfg.nx_slab : e->cfg.nx; }
FBMathDims mdims(chimera_engine* e) {
  return FBMathDims{nxs(e), e->cfg.nkr, e->cfg.nm, e->cfg.nkr, e->cfg.env, slab(e) ? e->cfg.mirror_shift : 0};
}

ChunkSpec chunkspec(chimera_engine* e, const Species& s) {
  if (!e->cfg.chunked) return ChunkSpec{0, nullptr, 1, 0, e->cfg.nx};
  return ChunkSpec{1, s.d_ind, e->cfg.nchnk, e->cfg.guards, e->cfg.nx / e->cfg.nchnk};
}

// CTA -> chunk table of the binned deposit (kernels.cuh SortedSpec), rebuilt whenever IndInChunk changes
int update_cta_table(chimera_engine* e, Species& s) {
  const int nchnk = (int)s.h_ind.size() - 1;
  std::vector<int> cta(nchnk + 1, 0);
  for (int c = 0; c < nchnk; ++c) {
    const int n = s.h_ind[c + 1] - s.h_ind[c];
    cta[c + 1] = cta[c] + (n > 0 ? (n + kDepNPB - 1) / kDepNPB : 0);
  }
  s.ncta = cta[nchnk];
  s.h_cta = cta;
  CHB_CUDA(cudaMemcpyAsync(s.d_cta, cta.data(), sizeof(int) * (nchnk + 1), cudaMemcpyHostToDevice, e->st));
  std::vector<int> ctaf(nchnk + 1, 0);
  for (int c = 0; c < nchnk; ++c) {
    const int n = s.h_ind[c + 1] - s.h_ind[c];
    ctaf[c + 1] = ctaf[c] + (n > 0 ? (n + kFusedNPB - 1) / kFusedNPB : 0);
  }
  s.ncta_f = ctaf[nchnk];
  CHB_CUDA(cudaMemcpyAsync(s.d_cta_f, ctaf.data(), sizeof(int) * (nchnk + 1), cudaMemcpyHostToDevice, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

SortedSpec sortedspec(chimera_engine* e, const Species& s) {
  const int nchnk = e->cfg.chunked ? e->cfg.nchnk : 1;
  return SortedSpec{s.d_ind, s.d_cta, nchnk, s.ncta, e->cfg.nx / nchnk, s.tile_w};
}

// ---- phases ----------------------------------------------------------------------------------
int ph_push_coords(chimera_engine* e) {
  for (auto& s : e->sp) {
    if (s.still || s.np == 0) continue;
    CHB_TRY(launch_push_coords(e->st, soa(s.x, s.cap), soa((const double*)s.p, s.cap), soa(s.xh, s.cap), e->cfg.dt, s.np));
  }
  return 0;
}

int ensure_sort_scratch(chimera_engine* e, i64 n) {
  if (n <= e->sort_cap) return 0;
  if (e->key_a) { cudaFree(e->key_a); cudaFree(e->key_b); cudaFree(e->idx_a); cudaFree(e->idx_b); cudaFree(e->cub_tmp); }
  CHB_CUDA(cudaMalloc((void**)&e->key_a, sizeof(unsigned) * n));
  CHB_CUDA(cudaMalloc((void**)&e->key_b, sizeof(unsigned) * n));
  CHB_CUDA(cudaMalloc((void**)&e->idx_a, sizeof(int) * n));
  CHB_CUDA(cudaMalloc((void**)&e->idx_b, sizeof(int) * n));
  size_t tb = 0;
  CHB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, e->key_a, e->key_b, e->idx_a, e->idx_b, (int)n, 0, 32, e->st));
  CHB_CUDA(cudaMalloc(&e->cub_tmp, tb + 16));
  e->cub_tmp_bytes = tb;
  e->sort_cap = n;
  return 0;
}

// left_margin: the absorbing layer of a moving window (species.py:373-376 SimDom[0] = leftX + AbsorbLayer dx)
// r2max > 0: radial cull limit of this call (the window's, species.py:373-376) instead of cfg.rcull2
int ph_sort(chimera_engine* e, int on_halfstep, double left_margin = 0.0, double r2max = 0.0) {
  const auto& c = e->cfg;
  const int nchnk = c.chunked ? c.nchnk : 1;
  for (auto& s : e->sp) {
    if (s.np == 0) continue;
    if (s.np > 0x7FFFFFF0LL) { set_error("re-binning supports up to 2^31 particles per GPU"); return 2; }
    CHB_TRY(ensure_sort_scratch(e, s.np));
    BinSpec b;
    const i64 cs = c.nx / nchnk;  // nodes per chunk
    b.x0 = c.leftX;
    b.chunk_inv = 1.0 / c.chunk_len;
    b.l0 = c.leftX + left_margin; b.l1 = c.rightX; b.l2 = 0.0; b.l3 = r2max > 0.0 ? r2max : c.rcull2;
    b.leftX = c.leftX; b.dx_inv = 1.0 / c.dx; b.r0 = g_host[e].r0; b.dr_inv = 1.0 / c.dr;
    b.nchnk = nchnk; b.nx = c.nx; b.nrc = c.nrn - 1;
    b.cs = cs; b.tile_w = cs < 32 ? cs : 32; b.ntile = (cs + b.tile_w - 1) / b.tile_w;
    const i64 per_chunk = b.ntile * b.nrc * b.tile_w;
    const unsigned long long kmax = (unsigned long long)nchnk * per_chunk;
    if (kmax >= 0xFFFFFFFFull) { set_error("re-binning key does not fit 32 bits"); return 2; }
    b.kdrop = (unsigned)kmax;
    int bits = 1;
    while (bits < 32 && (1ull << bits) < kmax + 1) ++bits;
    const double* src = on_halfstep ? s.xh : s.x;
    bin_keys_k<<<grid_for(s.np, 256), 256, 0, e->st>>>(src, s.cap, e->key_a, e->idx_a, b, s.np);
    CHB_LAUNCH_CHECK();
    size_t tb = e->cub_tmp_bytes;
    CHB_CUDA(cub::DeviceRadixSort::SortPairs(e->cub_tmp, tb, e->key_a, e->key_b, e->idx_a, e->idx_b, (int)s.np, 0, bits, e->st));
    chunk_bounds_k<<<1, 256, 0, e->st>>>(e->key_b, s.d_ind, nchnk, per_chunk, s.np);
    CHB_LAUNCH_CHECK();
    permute_soa_k<<<grid_for(s.np, 256), 256, 0, e->st>>>(s.x2, s.xh2, s.p2, s.w2, s.x, s.xh, s.p, s.w, e->idx_b, s.cap, s.np);
    CHB_LAUNCH_CHECK();
    std::swap(s.x, s.x2); std::swap(s.xh, s.xh2); std::swap(s.p, s.p2); std::swap(s.w, s.w2);
    s.h_ind.assign(nchnk + 1, 0);
    CHB_CUDA(cudaMemcpyAsync(s.h_ind.data(), s.d_ind, sizeof(int) * (nchnk + 1), cudaMemcpyDeviceToHost, e->st));
    CHB_CUDA(cudaStreamSynchronize(e->st));
    s.np = s.h_ind[nchnk];
    s.tile_w = b.tile_w;
    CHB_TRY(update_cta_table(e, s));
  }
  return 0;
}

int deposit_species(chimera_engine* e, int curr, cd* grid, bool still_only, bool use_half) {
  GridGeom g = geom_ready(e);
  for (auto& s : e->sp) {
    if (s.np == 0) continue;
    if (still_only != (s.still != 0)) continue;
    const double* xs = use_half ? s.xh : s.x;
    int rc = launch_deposit_binned(e->st, e->cfg.env, curr, xs, s.p, s.w, s.cap, grid, g, chunkspec(e, s), sortedspec(e, s));
    if (rc == -1)  // mode count without a binned instantiation
      rc = launch_deposit_runs(e->st, e->cfg.env, curr, soa(xs, s.cap), soa((const double*)s.p, s.cap), s.w, grid, g,
                               chunkspec(e, s), s.np);
    CHB_TRY(rc);
  }
  return 0;
}

int ph_deposit_j(chimera_engine* e) {
  const auto& c = e->cfg;
  const i64 n = c.nx * c.nrn * c.nm * 3;
  CHB_CUDA(cudaMemsetAsync(e->A("J"), 0, sizeof(cd) * n, e->st));
  CHB_TRY(deposit_species(e, 1, e->A("J"), false, true));
  e->leftX_J = c.leftX;
  return launch_ghost_fold(e->st, e->A("J"), c.nx, c.nrn, c.nm * 3);
}

// one stage of a window that moves every step (ChimeraRun.move_frame, chimera_main.py:286-290)
int ph_window(chimera_engine* e, int stage) {
  const double s = stage == 2 ? e->win_s2 : e->win_s1;
  e->cfg.leftX += s;
  e->cfg.rightX += s;
  return 0;
}

int ph_deposit_rho(chimera_engine* e, int from_bg) {
  const auto& c = e->cfg;
  const i64 n = c.nx * c.nrn * c.nm;
  if (from_bg) CHB_CUDA(cudaMemcpyAsync(e->A("Rho"), e->A("BckGrndRho"), sizeof(cd) * n, cudaMemcpyDeviceToDevice, e->st));
  else CHB_CUDA(cudaMemsetAsync(e->A("Rho"), 0, sizeof(cd) * n, e->st));
  CHB_TRY(deposit_species(e, 0, e->A("Rho"), false, c.static_kick != 0));  // chimera_main.py:186: coords_halfstep
  return launch_ghost_fold(e->st, e->A("Rho"), c.nx, c.nrn, c.nm);
}

int ph_deposit_bg(chimera_engine* e) {
  const auto& c = e->cfg;
  const i64 n = c.nx * c.nrn * c.nm;
  CHB_CUDA(cudaMemsetAsync(e->A("BckGrndRho"), 0, sizeof(cd) * n, e->st));
  CHB_TRY(deposit_species(e, 0, e->A("BckGrndRho"), true, false));
  return launch_ghost_fold(e->st, e->A("BckGrndRho"), c.nx, c.nrn, c.nm);
}

int ph_add_bg(chimera_engine* e) {
  const auto& c = e->cfg;
  const i64 n = c.nx * c.nrn * c.nm;
  add_cd_k<<<grid_for(n, 256), 256, 0, e->st>>>(e->A("Rho"), e->A("BckGrndRho"), n);
  CHB_LAUNCH_CHECK();
  return 0;
}

int ph_fb_in_j(chimera_engine* e) {
  const auto& c = e->cfg;
  FBCtx fb = fbctx(e);
  // the phase exp(-i kx leftX) belongs to the grid J was deposited on: with a 'Staged' window project_current runs
  // between the two half shifts (chimera_main.py:87-88)
  const double lx = std::isnan(e->leftX_J) ? c.leftX : e->leftX_J;
  e->leftX_J = NAN;
  if (slab(e))
    return fb_in_slab_dev(fb, e->A("J_fb"), e->A("J"), lx, e->D("kx_base"), e->pInCurr, e->D("DepFact"),
                          (const i64*)e->arr["slab_rows"].p, c.nx, nxs(e), c.nrn, c.nm, c.nkr, 3);
  return fb_in_dev(fb, e->A("J_fb"), e->A("J"), lx, e->D("kx_base"), e->pInCurr, e->D("DepFact"), c.nx, c.nrn, c.nm,
                   c.nkr, 3);
}

int ph_fb_in_rho(chimera_engine* e) {
  const auto& c = e->cfg;
  FBCtx fb = fbctx(e);
  std::swap(e->arr["gradRho_fb_prv"], e->arr["gradRho_fb_nxt"]);  // gradRho_fb_prv[:] = gradRho_fb_nxt
  if (slab(e))
    CHB_TRY(fb_in_slab_dev(fb, e->A("Rho_fb"), e->A("Rho"), c.leftX, e->D("kx_base"), e->pInCurr, e->D("DepFact"),
                           (const i64*)e->arr["slab_rows"].p, c.nx, nxs(e), c.nrn, c.nm, c.nkr, 1));
  else
    CHB_TRY(fb_in_dev(fb, e->A("Rho_fb"), e->A("Rho"), c.leftX, e->D("kx_base"), e->pInCurr, e->D("DepFact"), c.nx, c.nrn,
                      c.nm, c.nkr, 1));
  return fb_grad_dev(fb, e->A("gradRho_fb_nxt"), e->A("Rho_fb"), e->pDp, e->pDm, e->D("kx"), mdims(e));
}

int ph_poisson(chimera_engine* e) {
  const auto& c = e->cfg;
  const i64 P = nxs(e) * c.nkr * c.nm;
  FBCtx fb = fbctx(e);
  for (int it = 0; it < c.poisson_iters; ++it) {
    if (c.space_charge) {
      // grad div J_fb is consumed by the correction in the same pass; vec_fb (solvers.py:317-319) is not needed
      CHB_TRY(fb_poiss_iter_dev(fb, e->A("J_fb"), e->A("gradRho_fb_prv"), e->A("gradRho_fb_nxt"), 1.0 / c.dt,
                                e->D("PoissFact"), e->pDp, e->pDm, e->D("kx"), mdims(e)));
    } else {
      CHB_TRY(fb_graddiv_dev(fb, e->A("vec_fb"), e->A("J_fb"), e->pDp, e->pDm, e->D("kx"), mdims(e)));
      CHB_TRY(launch_mult_real(e->st, e->A("vec_fb"), e->D("PoissFact"), P, 3));
      CHB_TRY(launch_add(e->st, e->A("J_fb"), e->A("vec_fb"), P * 3));
    }
  }
  return 0;
}

int ph_maxwell(chimera_engine* e) {
  const auto& c = e->cfg;
  const i64 P = nxs(e) * c.nkr * c.nm;
  if (c.space_charge)
    return launch_maxwell_push(e->st, e->A("EG_fb"), e->A("J_fb"), e->A("gradRho_fb_prv"), e->A("gradRho_fb_nxt"),
                               e->arr["PSATD_E"].p, e->arr["PSATD_G"].p, 5, 0, P);
  return launch_maxwell_push(e->st, e->A("EG_fb"), e->A("J_fb"), nullptr, nullptr, e->arr["PSATD_E"].p, e->arr["PSATD_G"].p,
                             3, c.coef_complex, P);
}

// chimera_main.py:118-125 update_fields under 'StaticKick': the field is rebuilt from zero every step as the
// quasi-static field of each species' charge and current moving with its mean momentum
int ph_static_fields(chimera_engine* e) {
  const auto& c = e->cfg;
  if (!e->arr.count("w")) { set_error("StaticKick schedule: engine was created without static_kick"); return 2; }
  const i64 nxl = nxs(e);  // kx rows held here: all of them, or this rank's slab (every operation below is per kx row)
  const i64 P = nxl * c.nkr * c.nm;
  FBCtx fb = fbctx(e);
  CHB_CUDA(cudaMemsetAsync(e->A("EG_fb"), 0, sizeof(cd) * P * 6, e->st));
  for (size_t is = 0; is < e->sp.size(); ++is) {
    auto& s = e->sp[is];
    double px;
    if (is < e->static_px.size()) {
      // given from outside (chimera_engine_set_static_px): across ranks the 16 moments are all-reduced first
      px = e->static_px[is];
      if (std::isnan(px)) continue;  // the species is empty on every rank
    } else {
      if (s.np == 0) continue;
      // PXmean = sum(px w) / sum(w) (chimera_main.py:121-122): reduced on the device, 16 doubles to the host
      double* d_m = e->scr.take_n<double>(16);
      if (!d_m) return 6;
      double m[16];
      CHB_TRY(launch_beam_moments(e->st, s.xh, s.p, s.w, s.cap, s.np, d_m));
      CHB_CUDA(cudaMemcpyAsync(m, d_m, sizeof(m), cudaMemcpyDeviceToHost, e->st));
      CHB_CUDA(cudaStreamSynchronize(e->st));
      px = m[5] / m[0];
    }
    const double beta0 = px / sqrt(1.0 + px * px);
    if (c.poisson_iters > 0) {  // solvers.py:360-383 poiss_corr_stat (one pass, unless NoPoissonCorrection)
      cd* DT = e->scr.take_n<cd>(nxl);
      if (!DT) return 6;
      CHB_TRY(launch_dt_stat(e->st, DT, e->D("kx"), beta0, nxl));
      CHB_TRY(fb_graddiv_dev(fb, e->A("vec_fb"), e->A("J_fb"), e->pDp, e->pDm, e->D("kx"), mdims(e)));
      CHB_TRY(launch_poiss_corr_stat(e->st, e->A("J_fb"), e->A("vec_fb"), e->A("gradRho_fb_nxt"), DT, e->D("PoissFact"), nxl, P));
    }
    CHB_TRY(launch_maxwell_static_push(e->st, e->A("EG_fb"), e->A("J_fb"), e->A("gradRho_fb_nxt"), e->D("w"), e->D("kx"),
                                       beta0, nxl, P));
    CHB_TRY(launch_field_drift(e->st, e->A("EG_fb"), e->D("kx"), beta0, c.dt, nxl, c.nkr * c.nm * 6));
  }
  e->static_px.clear();
  return 0;
}

int ph_init_push(chimera_engine* e) {
  const auto& c = e->cfg;
  const i64 P = nxs(e) * c.nkr * c.nm;
  return launch_maxwell_init_push(e->st, e->A("EG_fb"), e->A("J_fb"), e->A("gradRho_fb_nxt"), e->A("CPSATD1"),
                                  e->A("CPSATD2"), P);
}

// ---- column-block dataflow of the multi-rank solve ------------------------------------------
// J (Nx, Nr, M, 3) is a column-major [Nx x C] matrix (x fastest).  Instead of all-reducing it and x-transforming all of it
// on every rank: reduce-scatter by column block -> x-FFT of the own block -> all-to-all (column block -> kx slab) -> DHT on
// the slab.  Backward: slab -> all-to-all (kx slab -> column block) -> inverse x-FFT of the own block -> all-gather of
// the blocks, which land in EB in their natural order.  C is padded to a multiple of the rank count (kColPad).
static i64 col_block(const chimera_engine* e, i64 ncols) { return (ncols + e->col_world - 1) / e->col_world; }

// which: 0 = J (3 components), 1 = Rho.  In: "<X>_blk" (reduce-scatter output).  Out: "<X>_send".
int ph_col_fwd(chimera_engine* e, int which) {
  const auto& c = e->cfg;
  if (!e->col_world) { set_error("column dataflow is not set up (chimera_engine_set_colflow)"); return 2; }
  FBCtx fb = fbctx(e);
  const i64 cb = col_block(e, c.nrn * c.nm * (which ? 1 : 3));
  return col_fwd_dev(fb, e->A(which ? "Rho_send" : "J_send"), e->A(which ? "Rho_blk" : "J_blk"),
                     (const i64*)e->arr["gather_map"].p, c.nx, nxs(e), cb);
}

// forward DHT of the slab that came in through the all-to-all ("J_in" / "Rho_in": (nx_slab, padded columns))
int ph_fb_in_col(chimera_engine* e, int which) {
  const auto& c = e->cfg;
  FBCtx fb = fbctx(e);
  if (!which) {
    const double lx = std::isnan(e->leftX_J) ? c.leftX : e->leftX_J;
    e->leftX_J = NAN;
    return fb_in_slab_post_dev(fb, e->A("J_fb"), e->A("J_in"), lx, e->D("kx_base"), e->pInCurr, e->D("DepFact"), nxs(e), c.nrn,
                               c.nm, c.nkr, 3);
  }
  std::swap(e->arr["gradRho_fb_prv"], e->arr["gradRho_fb_nxt"]);  // gradRho_fb_prv[:] = gradRho_fb_nxt
  CHB_TRY(fb_in_slab_post_dev(fb, e->A("Rho_fb"), e->A("Rho_in"), c.leftX, e->D("kx_base"), e->pInCurr, e->D("DepFact"), nxs(e),
                              c.nrn, c.nm, c.nkr, 1));
  return fb_grad_dev(fb, e->A("gradRho_fb_nxt"), e->A("Rho_fb"), e->pDp, e->pDm, e->D("kx"), mdims(e));
}

// backward: "EB_recv" ((rank, column, row) blocks from the all-to-all) -> "EB_blk": rows put in place, inverse x-FFT
int ph_col_bwd(chimera_engine* e) {
  const auto& c = e->cfg;
  if (!e->col_world) { set_error("column dataflow is not set up (chimera_engine_set_colflow)"); return 2; }
  FBCtx fb = fbctx(e);
  const i64 cb = col_block(e, c.nrn * c.nm * 6);
  return col_bwd_dev(fb, e->A("EB_blk"), e->A("EB_recv"), (const i64*)e->arr["gather_map"].p, c.nx, nxs(e), cb,
                     e->col_rank * cb, c.nrn, c.nm, c.env);
}

// after the all-gather of the blocks into EB: the ghost rows of eb_correction (grid_deps.f90:248-265; the normalisation
// went into col_bwd)
int ph_eb_finish(chimera_engine* e) {
  const auto& c = e->cfg;
  FBCtx fb = fbctx(e);
  return eb_ghost_dev(fb, e->A("EB"), c.nx, c.nrn, c.nm, c.env, 6);
}

// fields out, first half: B from G, backward DHT (+ phase).  Unsharded: also the inverse x-FFT and the
// normalisation, i.e. the whole of G2B_FBRot + fb_fld_out (solvers.py:536, 450).  kx-slab mode: stops at
// "EB_slab"; the caller all-gathers the slabs into "EB_gath" and runs the second half.
// part (kx-slab mode): 0 = E and B together; 1 = the E half only, 2 = the B half only -- so that the all-gather of one
// half can run while the other half is computed / finished (EB_slab and EB are (.., 6) with the component slowest:
// each half is contiguous; in split mode EB_gath holds [half][rank][(nx_slab, Nr, M, 3)])
int ph_fields_out_a(chimera_engine* e, bool whole, int part = 0) {
  const auto& c = e->cfg;
  const i64 P = nxs(e) * c.nkr * c.nm;
  FBCtx fb = fbctx(e);
  if (part != 1) {
    CHB_TRY(fb_rot_dev(fb, e->A("B_fb"), e->A("EG_fb") + P * 3, e->pDp, e->pDm, e->D("kx"), mdims(e)));
    CHB_TRY(launch_mult_real(e->st, e->A("B_fb"), e->D("PoissFact"), P, 3));
  }
  const cd* srcs[2] = {e->A("EG_fb"), e->A("B_fb")};
  if (slab(e)) {
    if (whole) { set_error("kx-slab engine: run fields_out_a, all-gather EB_slab into EB_gath, then fields_out_b"); return 2; }
    if (part == 0)
      return fb_out_slab_dev(fb, e->A("EB_slab"), srcs, 2, 3, c.leftX, e->D("kx_base"), e->pOut, nxs(e), c.nrn, c.nm, c.nkr);
    const i64 half = nxs(e) * c.nrn * c.nm * 3;
    return fb_out_slab_dev(fb, e->A("EB_slab") + (part - 1) * half, srcs + (part - 1), 1, 3, c.leftX, e->D("kx_base"), e->pOut,
                           nxs(e), c.nrn, c.nm, c.nkr);
  }
  if (part != 0) { set_error("fields_out halves are a kx-slab feature"); return 2; }
  CHB_TRY(fb_out_dev(fb, e->A("EB"), srcs, 2, 3, c.leftX, e->D("kx_base"), e->pOut, c.nx, c.nrn, c.nm, c.nkr));
  return launch_eb_correction(e->st, e->A("EB"), c.nx, c.nrn, c.nm, c.env);
}

int ph_fields_out_b(chimera_engine* e, int part = 0) {
  const auto& c = e->cfg;
  if (!slab(e)) return 0;  // everything was done by the first half
  FBCtx fb = fbctx(e);
  if (part == 0) {
    CHB_TRY(fb_out_finish_dev(fb, e->A("EB"), e->A("EB_gath"), (const i64*)e->arr["gather_map"].p, c.nx, nxs(e),
                              c.nrn * c.nm * 6));
    return launch_eb_correction(e->st, e->A("EB"), c.nx, c.nrn, c.nm, c.env);
  }
  const i64 half = c.nx * c.nrn * c.nm * 3;  // the same for EB and for the split EB_gath (world * nx_slab = nx)
  cd* eb = e->A("EB") + (part - 1) * half;
  CHB_TRY(fb_out_finish_dev(fb, eb, e->A("EB_gath") + (part - 1) * half, (const i64*)e->arr["gather_map"].p, c.nx, nxs(e),
                            c.nrn * c.nm * 3));
  return launch_eb_correction(e->st, eb, c.nx, c.nrn, c.nm, c.env, 3);
}

// the device list of a species at the engine's current device time (species.py:274-277)
static DeviceSet devset(const chimera_engine* e, const Species& s) {
  DeviceSet d;
  memset(&d, 0, sizeof(d));
  d.t = e->dev_time;
  d.n = (int)s.devs.size();
  for (int i = 0; i < d.n; ++i) d.d[i] = s.devs[i];
  return d;
}

int ph_gather_push(chimera_engine* e, double dt_frac) {
  const auto& c = e->cfg;
  GridGeom g = geom_ready(e);
  for (auto& s : e->sp) {
    if (s.still || s.np == 0) continue;
    const DeviceSet und = devset(e, s);
    int rc = -1;
    if (e->fuse)  // thread per particle straight from the grid (particles_fused.cu)
      rc = launch_gather_push_coords(e->st, c.env, s.x, s.xh, s.p, s.w, s.cap, e->A("EB"), g, s.push_fact * c.dt * dt_frac,
                                     c.dt, und, s.np, 0);
    if (rc == -1)
      rc = launch_gather_push_binned(e->st, c.env, s.x, s.w, e->A("EB"), s.p, s.cap, g, s.push_fact * c.dt * dt_frac, und,
                                     sortedspec(e, s));
    if (rc == -1)
      rc = launch_gather_push_tiled(e->st, c.env, soa((const double*)s.x, s.cap), s.w, e->A("EB"), soa(s.p, s.cap), g,
                                    s.push_fact * c.dt * dt_frac, und, s.np);
    CHB_TRY(rc);
  }
  return 0;
}

// The half of a re-binning step before its sort: gather + push_velocs of step k and push_coords of step k+1 in one
// streaming kernel (particles_fused.cu), no deposit.  The window stage 1 of step k+1 (chimera_main.py:83) comes after
// the gather and before anything that looks at the grid position, so it is applied here.
int ph_gather_push_coords(chimera_engine* e) {
  auto& c = e->cfg;
  GridGeom g = geom_ready(e);
  for (auto& s : e->sp) {
    if (s.still || s.np == 0) continue;
    const DeviceSet und = devset(e, s);
    int rc = launch_gather_push_coords(e->st, c.env, s.x, s.xh, s.p, s.w, s.cap, e->A("EB"), g, s.push_fact * c.dt, c.dt, und,
                                       s.np, 1);
    if (rc == -1) {
      rc = launch_gather_push_binned(e->st, c.env, s.x, s.w, e->A("EB"), s.p, s.cap, g, s.push_fact * c.dt, und, sortedspec(e, s));
      if (rc == -1)
        rc = launch_gather_push_tiled(e->st, c.env, soa((const double*)s.x, s.cap), s.w, e->A("EB"), soa(s.p, s.cap), g,
                                      s.push_fact * c.dt, und, s.np);
      CHB_TRY(rc);
      rc = launch_push_coords(e->st, soa(s.x, s.cap), soa((const double*)s.p, s.cap), soa(s.xh, s.cap), c.dt, s.np);
    }
    CHB_TRY(rc);
  }
  return ph_window(e, 1);
}

// dep_curr + dep_dens of a step from the stored x_half / x / p in ONE kernel (MODE 2 of the fused kernel): the half
// of a re-binning step after its sort, and the head of the first step of a step() call.  Applies window stage 2
// (between dep_curr and dep_dens, chimera_main.py:87) like ph_particles_fused does.
int ph_deposit_fused(chimera_engine* e, int rho_from_bg) {
  auto& c = e->cfg;
  GridGeom g = geom_ready(e);
  const double lJ = c.leftX, lR = lJ + e->win_s2;
  GridGeom gJ = g, gR = g;
  gR.leftX = lR;
  const bool rho = c.space_charge || c.static_kick;
  const i64 n = c.nx * c.nrn * c.nm;
  CHB_CUDA(cudaMemsetAsync(e->A("J"), 0, sizeof(cd) * n * 3, e->st));
  if (rho) {
    if (rho_from_bg) CHB_CUDA(cudaMemcpyAsync(e->A("Rho"), e->A("BckGrndRho"), sizeof(cd) * n, cudaMemcpyDeviceToDevice, e->st));
    else CHB_CUDA(cudaMemsetAsync(e->A("Rho"), 0, sizeof(cd) * n, e->st));
  }
  for (auto& s : e->sp) {
    if (s.still || s.np == 0) continue;
    const DeviceSet und = devset(e, s);
    SortedSpec spf = sortedspec(e, s);
    spf.cta = s.d_cta_f;
    spf.ncta = s.ncta_f;
    int rc = -1;
    if (!c.static_kick)  // 'StaticKick' deposits rho on coords_halfstep (chimera_main.py:186): separate kernels
      rc = launch_fused_particles(e->st, c.env, c.space_charge, s.x, s.xh, s.p, s.w, s.cap, e->A("EB"), e->A("J"),
                                  e->A("Rho"), g, chunkspec(e, s), s.push_fact * c.dt, c.dt, und, spf, lJ, lR, 1);
    if (rc == -1) {
      for (int curr = 1; curr >= (rho ? 0 : 1); --curr) {
        cd* grid = curr ? e->A("J") : e->A("Rho");
        const double* xs = (curr || c.static_kick) ? s.xh : s.x;
        const GridGeom& gd = curr ? gJ : gR;
        rc = launch_deposit_binned(e->st, c.env, curr, xs, s.p, s.w, s.cap, grid, gd, chunkspec(e, s), sortedspec(e, s));
        if (rc == -1)
          rc = launch_deposit_runs(e->st, c.env, curr, soa(xs, s.cap), soa((const double*)s.p, s.cap), s.w, grid, gd,
                                   chunkspec(e, s), s.np);
        CHB_TRY(rc);
      }
      rc = 0;
    }
    CHB_TRY(rc);
  }
  CHB_TRY(launch_ghost_fold(e->st, e->A("J"), c.nx, c.nrn, c.nm * 3));
  if (rho) CHB_TRY(launch_ghost_fold(e->st, e->A("Rho"), c.nx, c.nrn, c.nm));
  e->leftX_J = lJ;
  return ph_window(e, 2);
}

// gather + push_velocs of step k fused with push_coords + dep_curr + dep_dens of step k+1 (particles_fused.cu);
// species without a fused instantiation (or still ones) take the separate kernels
int ph_particles_fused(chimera_engine* e, int rho_from_bg) {
  auto& c = e->cfg;
  GridGeom g = geom_ready(e);  // the gather closes the previous step: window where that step left it
  // window stages of the step whose head is fused in: J on the grid after stage 1, rho after stage 2
  const double lJ = c.leftX + e->win_s1, lR = lJ + e->win_s2;
  GridGeom gJ = g, gR = g;
  gJ.leftX = lJ;
  gR.leftX = lR;
  const i64 n = c.nx * c.nrn * c.nm;
  CHB_CUDA(cudaMemsetAsync(e->A("J"), 0, sizeof(cd) * n * 3, e->st));
  if (c.space_charge) {
    if (rho_from_bg) CHB_CUDA(cudaMemcpyAsync(e->A("Rho"), e->A("BckGrndRho"), sizeof(cd) * n, cudaMemcpyDeviceToDevice, e->st));
    else CHB_CUDA(cudaMemsetAsync(e->A("Rho"), 0, sizeof(cd) * n, e->st));
  }
  for (auto& s : e->sp) {
    if (s.still || s.np == 0) continue;
    const DeviceSet und = devset(e, s);
    SortedSpec spf = sortedspec(e, s);
    spf.cta = s.d_cta_f;
    spf.ncta = s.ncta_f;
    int rc = launch_fused_particles(e->st, c.env, c.space_charge, s.x, s.xh, s.p, s.w, s.cap, e->A("EB"), e->A("J"),
                                    e->A("Rho"), g, chunkspec(e, s), s.push_fact * c.dt, c.dt, und, spf, lJ, lR);
    if (rc == -1) {  // no fused instantiation for this mode count: the separate kernels, same order of operations
      rc = launch_gather_push_binned(e->st, c.env, s.x, s.w, e->A("EB"), s.p, s.cap, g, s.push_fact * c.dt, und, sortedspec(e, s));
      if (rc == -1)
        rc = launch_gather_push_tiled(e->st, c.env, soa((const double*)s.x, s.cap), s.w, e->A("EB"), soa(s.p, s.cap), g,
                                      s.push_fact * c.dt, und, s.np);
      CHB_TRY(rc);
      CHB_TRY(launch_push_coords(e->st, soa(s.x, s.cap), soa((const double*)s.p, s.cap), soa(s.xh, s.cap), c.dt, s.np));
      for (int curr = 1; curr >= (c.space_charge ? 0 : 1); --curr) {
        cd* grid = curr ? e->A("J") : e->A("Rho");
        const double* xs = curr ? s.xh : s.x;
        const GridGeom& gd = curr ? gJ : gR;
        rc = launch_deposit_binned(e->st, c.env, curr, xs, s.p, s.w, s.cap, grid, gd, chunkspec(e, s), sortedspec(e, s));
        if (rc == -1)
          rc = launch_deposit_runs(e->st, c.env, curr, soa(xs, s.cap), soa((const double*)s.p, s.cap), s.w, grid, gd,
                                   chunkspec(e, s), s.np);
        CHB_TRY(rc);
      }
      rc = 0;
    }
    CHB_TRY(rc);
  }
  CHB_TRY(launch_ghost_fold(e->st, e->A("J"), c.nx, c.nrn, c.nm * 3));
  if (c.space_charge) CHB_TRY(launch_ghost_fold(e->st, e->A("Rho"), c.nx, c.nrn, c.nm));
  e->leftX_J = lJ;
  c.rightX += e->win_s1;  // same two roundings as the separate stages
  c.rightX += e->win_s2;
  c.leftX = lR;
  return 0;
}

cudaEvent_t get_event(chimera_engine* e) {
  if (!e->ev_pool.empty()) { cudaEvent_t v = e->ev_pool.back(); e->ev_pool.pop_back(); return v; }
  cudaEvent_t v;
  cudaEventCreate(&v);
  return v;
}

int collect_timings(chimera_engine* e) {
  if (e->pending.empty()) return 0;
  CHB_CUDA(cudaStreamSynchronize(e->st));
  for (auto& pe : e->pending) {
    float t = 0;
    cudaEventElapsedTime(&t, pe.second.first, pe.second.second);
    e->ms[pe.first] += t;
    e->calls[pe.first] += 1;
    e->ev_pool.push_back(pe.second.first);
    e->ev_pool.push_back(pe.second.second);
  }
  e->pending.clear();
  return 0;
}

int run_phase(chimera_engine* e, int phase, double arg) {
  CHB_TRY(ensure_ops(e));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (e->profile) {
    e0 = get_event(e); e1 = get_event(e);
    cudaEventRecord(e0, e->st);
  }
  int rc = 0;
  switch (phase) {
    case CHB_PUSH_COORDS: rc = ph_push_coords(e); break;
    case CHB_SORT: rc = ph_sort(e, arg != 0.0); break;
    case CHB_DEPOSIT_J: rc = ph_deposit_j(e); break;
    case CHB_DEPOSIT_RHO: rc = ph_deposit_rho(e, arg != 0.0); break;
    case CHB_DEPOSIT_BG: rc = ph_deposit_bg(e); break;
    case CHB_FB_IN_J: rc = ph_fb_in_j(e); break;
    case CHB_FB_IN_RHO: rc = ph_fb_in_rho(e); break;
    case CHB_POISSON: rc = ph_poisson(e); break;
    case CHB_MAXWELL: rc = ph_maxwell(e); break;
    case CHB_INIT_PUSH: rc = ph_init_push(e); break;
    case CHB_FIELDS_OUT: rc = ph_fields_out_a(e, true); break;
    case CHB_FIELDS_OUT_A: rc = ph_fields_out_a(e, false, arg == 1.0 ? 1 : (arg == 2.0 ? 2 : 0)); break;
    case CHB_FIELDS_OUT_B: rc = ph_fields_out_b(e, arg == 1.0 ? 1 : (arg == 2.0 ? 2 : 0)); break;
    case CHB_PARTICLES_FUSED: rc = ph_particles_fused(e, arg != 0.0); break;
    case CHB_GATHER_PUSH: rc = ph_gather_push(e, arg); break;
    case CHB_ADD_BG: rc = ph_add_bg(e); break;
    case CHB_STATIC_FIELDS: rc = ph_static_fields(e); break;
    case CHB_WINDOW: rc = ph_window(e, arg == 2.0 ? 2 : 1); break;
    case CHB_GATHER_PUSH_COORDS: rc = ph_gather_push_coords(e); break;
    case CHB_DEPOSIT_FUSED: rc = ph_deposit_fused(e, arg != 0.0); break;
    case CHB_COL_FWD: rc = ph_col_fwd(e, arg != 0.0); break;
    case CHB_FB_IN_COL: rc = ph_fb_in_col(e, arg != 0.0); break;
    case CHB_COL_BWD: rc = ph_col_bwd(e); break;
    case CHB_EB_FINISH: rc = ph_eb_finish(e); break;
    default: set_error("unknown engine phase %d", phase); rc = 2;
  }
  if (e->profile) {
    cudaEventRecord(e1, e->st);
    e->pending.push_back({phase, {e0, e1}});
    if (e->pending.size() > 4096) collect_timings(e);
  }
  // temporaries of this phase are dead once the stream reaches this point; later phases on the same
  // stream may reuse them
  e->scr.reset();
  return rc;
}

// the gather + push_velocs that closes the last step of the previous chimera_engine_step call
int flush_tail(chimera_engine* e) {
  if (!e->tail_pending) return 0;
  e->tail_pending = false;
  e->dev_time = e->tail_time;
  return run_phase(e, CHB_GATHER_PUSH, 1.0);
}

}  // namespace

extern "C" {

int chimera_engine_create(const chimera_engine_config* cfg, chimera_engine** out) {
  if (!cfg || !out) { set_error("engine_create: null argument"); return 2; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device available: libchimera_b200 has no CPU fallback");
    return 1;
  }
  const auto& c = *cfg;
  if (c.nx < 2 || c.nrn < 2 || c.nkr < 1 || c.nm < 1 || c.nm > 2 * kMaxModes) { set_error("engine_create: bad grid shape"); return 2; }
  if (c.env && (c.nm % 2) != 1) { set_error("engine_create: envelope solver needs an odd number of mode slots"); return 2; }
  if (c.chunked && (c.nchnk < 1 || c.nchnk > 127 || c.nx % c.nchnk)) { set_error("engine_create: bad chunking"); return 2; }
  chimera_engine* e = new chimera_engine();
  e->cfg = c;
  for (int i = 0; i < CHB_NPHASES; ++i) { e->ms[i] = 0; e->calls[i] = 0; }
  CHB_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
  e->own_stream = true;
  if (c.nx_slab < 0 || c.nx_slab > c.nx || (c.nx_slab & 1)) { set_error("engine_create: bad kx slab size"); chimera_engine_destroy(e); return 2; }
  const size_t nxf = (size_t)nxs(e);  // kx rows of the spectral state held here
  const size_t Pg = (size_t)c.nx * c.nrn * c.nm, Pf = nxf * c.nkr * c.nm, C = sizeof(cd);
  const i64 nr = c.nrn - 1;
  const int nd = (int)(c.env ? c.nm + 2 : c.nm + 1);
  const int ncoef = c.space_charge ? 5 : 3;
  struct { const char* n; size_t b; } spec[] = {
      // (J, Rho and EB get kColPad spare columns below: the column-block dataflow splits them into `world` equal blocks)
      {"J", Pg * 3 * C}, {"Rho", Pg * C}, {"BckGrndRho", Pg * C}, {"EB", Pg * 6 * C},
      {"EG_fb", Pf * 6 * C}, {"J_fb", Pf * 3 * C}, {"B_fb", Pf * 3 * C}, {"Rho_fb", Pf * C},
      {"gradRho_fb_prv", Pf * 3 * C}, {"gradRho_fb_nxt", Pf * 3 * C}, {"vec_fb", Pf * 3 * C},
      {"InCurr", sizeof(double) * nr * c.nkr * c.nm}, {"Out", sizeof(double) * nr * c.nkr * c.nm},
      {"DpS2S", sizeof(double) * c.nkr * c.nkr * nd}, {"DmS2S", sizeof(double) * c.nkr * c.nkr * nd},
      {"kx", sizeof(double) * nxf}, {"kx_base", sizeof(double) * nxf},
      {"DepFact", sizeof(double) * Pf}, {"PoissFact", sizeof(double) * Pf},
      {"PSATD_E", (c.coef_complex ? C : sizeof(double)) * Pf * ncoef},
      {"PSATD_G", (c.coef_complex ? C : sizeof(double)) * Pf * ncoef},
      {"CPSATD1", Pf * 2 * C}, {"CPSATD2", Pf * 2 * C}, {"Rgrid", sizeof(double) * c.nrn}};
  for (auto& s : spec) {
    const std::string nm(s.n);
    const bool padded = nm == "J" || nm == "Rho" || nm == "EB";
    int rc = alloc_named(e, s.n, s.b, true, padded ? kColPad * (size_t)c.nx * C : 0);
    if (rc) { chimera_engine_destroy(e); return rc; }
  }
  if (c.static_kick) {
    int rc = alloc_named(e, "w", sizeof(double) * Pf);
    if (rc) { chimera_engine_destroy(e); return rc; }
  }
  if (slab(e)) {
    struct { const char* n; size_t b; } extra[] = {
        {"slab_rows", sizeof(i64) * nxf}, {"gather_map", sizeof(i64) * (size_t)c.nx},
        {"EB_slab", nxf * c.nrn * c.nm * 6 * C}, {"EB_gath", Pg * 6 * C}};
    for (auto& s : extra) {
      int rc = alloc_named(e, s.n, s.b, true, std::string(s.n) == "EB_slab" ? kColPad * nxf * C : 0);
      if (rc) { chimera_engine_destroy(e); return rc; }
    }
  }
  *out = e;
  return 0;
}

int chimera_engine_destroy(chimera_engine* e) {
  if (!e) return 0;
  cudaStreamSynchronize(e->st);
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
  for (auto& kv : e->arr) cudaFree(kv.second.p);
  for (auto& s : e->sp) {
    cudaFree(s.x); cudaFree(s.xh); cudaFree(s.p); cudaFree(s.w);
    cudaFree(s.x2); cudaFree(s.xh2); cudaFree(s.p2); cudaFree(s.w2); cudaFree(s.d_ind); cudaFree(s.d_cta); cudaFree(s.d_cta_f);
    for (double* m : s.dev_maps) cudaFree(m);
  }
  cudaFree(e->packed);
  cudaFree(e->key_a); cudaFree(e->key_b); cudaFree(e->idx_a); cudaFree(e->idx_b); cudaFree(e->cub_tmp);
  e->scr.destroy();
  e->fft.destroy();
  for (auto& pe : e->pending) { cudaEventDestroy(pe.second.first); cudaEventDestroy(pe.second.second); }
  for (auto v : e->ev_pool) cudaEventDestroy(v);
  if (e->own_stream) cudaStreamDestroy(e->st);
  if (e->s_h2d) { cudaStreamDestroy(e->s_h2d); cudaStreamDestroy(e->s_d2h); }
  g_host.erase(e);
  delete e;
  return 0;
}

static int find_array(chimera_engine* e, const char* name, NamedArray** a) {
  auto it = e->arr.find(name ? name : "");
  if (it == e->arr.end()) { set_error("engine has no array named '%s'", name ? name : "(null)"); return 2; }
  *a = &it->second;
  return 0;
}

int chimera_engine_upload(chimera_engine* e, const char* name, const void* src, chb_i64 nbytes) {
  ENG_ENTER(e);
  NamedArray* a;
  CHB_TRY(find_array(e, name, &a));
  if ((size_t)nbytes != a->bytes) { set_error("upload '%s': %lld bytes given, %zu expected", name, nbytes, a->bytes); return 2; }
  CHB_CUDA(cudaMemcpyAsync(a->p, src, a->bytes, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  const std::string n(name);
  if (n == "InCurr" || n == "Out" || n == "DpS2S" || n == "DmS2S" || n == "Rgrid") e->ops_dirty = true;
  return 0;
}

int chimera_engine_download(chimera_engine* e, const char* name, void* dst, chb_i64 nbytes) {
  ENG_ENTER(e);
  NamedArray* a;
  CHB_TRY(find_array(e, name, &a));
  if ((size_t)nbytes != a->bytes) { set_error("download '%s': %lld bytes given, %zu expected", name, nbytes, a->bytes); return 2; }
  CHB_CUDA(cudaMemcpyAsync(dst, a->p, a->bytes, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

int chimera_engine_array(chimera_engine* e, const char* name, void** dev_ptr, chb_i64* nbytes) {
  ENG_ENTER(e);
  NamedArray* a;
  CHB_TRY(find_array(e, name, &a));
  *dev_ptr = a->p;
  *nbytes = (chb_i64)a->alloc;  // with the spare columns: the caller views what it needs
  return 0;
}

int chimera_engine_add_species(chimera_engine* e, const double* coords, const double* coords_half, const double* momenta,
                               const double* weights, chb_i64 np, double push_fact, int still, chb_i64 capacity,
                               int* id) {
  ENG_ENTER(e);
  if (np < 0) { set_error("add_species: np < 0"); return 2; }
  Species s;
  s.np = np;
  s.cap = capacity > np ? capacity : np;
  if (s.cap < 1) s.cap = 1;
  s.cap = (s.cap + 31) & ~31LL;  // keep every component row 256-byte aligned
  s.push_fact = push_fact;
  s.still = still;
  const size_t b3 = sizeof(double) * 3 * s.cap, b1 = sizeof(double) * s.cap;
  CHB_CUDA(cudaMalloc((void**)&s.x, b3)); CHB_CUDA(cudaMalloc((void**)&s.xh, b3)); CHB_CUDA(cudaMalloc((void**)&s.p, b3));
  CHB_CUDA(cudaMalloc((void**)&s.w, b1));
  CHB_CUDA(cudaMalloc((void**)&s.x2, b3)); CHB_CUDA(cudaMalloc((void**)&s.xh2, b3)); CHB_CUDA(cudaMalloc((void**)&s.p2, b3));
  CHB_CUDA(cudaMalloc((void**)&s.w2, b1));
  const int nchnk = e->cfg.chunked ? e->cfg.nchnk : 1;
  CHB_CUDA(cudaMalloc((void**)&s.d_ind, sizeof(int) * (nchnk + 1)));
  CHB_CUDA(cudaMalloc((void**)&s.d_cta, sizeof(int) * (nchnk + 1)));
  CHB_CUDA(cudaMalloc((void**)&s.d_cta_f, sizeof(int) * (nchnk + 1)));
  s.h_ind.assign(nchnk + 1, 0);
  // placeholder until the first re-binning (CHB_SORT must run before a chunked deposit, as the
  // reference's make_halfstep does, chimera_main.py:62-70)
  for (int c2 = 1; c2 <= nchnk; ++c2) s.h_ind[c2] = (int)np;
  CHB_CUDA(cudaMemcpyAsync(s.d_ind, s.h_ind.data(), sizeof(int) * (nchnk + 1), cudaMemcpyHostToDevice, e->st));
  if (np > 0) {
    // stage the (3,np) arrays through the permutation buffers, then transpose to SoA
    const double* srcs[3] = {coords, coords_half ? coords_half : coords, momenta};
    double* stage[3] = {s.x2, s.xh2, s.p2};
    double* dsts[3] = {s.x, s.xh, s.p};
    for (int k = 0; k < 3; ++k) {
      CHB_CUDA(cudaMemcpyAsync(stage[k], srcs[k], sizeof(double) * 3 * np, cudaMemcpyDefault, e->st));
      aos_to_soa_k<<<grid_for(3 * np, 256), 256, 0, e->st>>>(dsts[k], stage[k], 3, s.cap, np);
      CHB_LAUNCH_CHECK();
    }
    CHB_CUDA(cudaMemcpyAsync(s.w, weights, sizeof(double) * np, cudaMemcpyDefault, e->st));
  }
  CHB_CUDA(cudaStreamSynchronize(e->st));
  CHB_TRY(update_cta_table(e, s));
  if (e->cfg.undulator && !still) {  // configuration shortcut for the FEL setups: one analytic undulator on every beam
    const double prm[4] = {e->cfg.und_a0, e->cfg.und_lambda, e->cfg.und_X0, e->cfg.und_Lx};
    s.devs.push_back(one_device(DEV_UNDUL_ANALYTIC, 0.0, prm, 4, nullptr, 0, 0.0).d[0]);
  }
  e->sp.push_back(s);
  if (id) *id = (int)e->sp.size() - 1;
  return 0;
}

int chimera_engine_add_device(chimera_engine* e, int id, int kind, double a0, const double* params, int nparams,
                              const double* map, chb_i64 nx) {
  ENG_ENTER(e);
  if (id < -1 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  static const int need[7] = {0, 4, 5, 3, 5, 7, 8};
  if (kind < DEV_UNDUL_ANALYTIC || kind > DEV_GAUSSBEAM) { set_error("add_device: unknown device kind %d", kind); return 2; }
  if (!params || nparams != need[kind]) { set_error("add_device: kind %d takes %d parameters, got %d", kind, need[kind], nparams); return 2; }
  const bool mapped = kind == DEV_UNDUL_MAPPED || kind == DEV_UNDUL_MAPPED_TAP;
  if (mapped && (!map || nx < 3)) { set_error("add_device: mapped undulator needs a0(2,nx), nx >= 3"); return 2; }
  for (int k = 0; k < (int)e->sp.size(); ++k) {
    if (id >= 0 && k != id) continue;
    Species& s = e->sp[k];
    if (s.still) continue;  // species.py:272
    if ((int)s.devs.size() >= kMaxDevices) { set_error("add_device: at most %d devices per species", kMaxDevices); return 2; }
    double* dmap = nullptr;
    if (mapped) {
      CHB_CUDA(cudaMalloc(&dmap, sizeof(double) * 2 * nx));
      CHB_CUDA(cudaMemcpyAsync(dmap, map, sizeof(double) * 2 * nx, cudaMemcpyDefault, e->st));
      CHB_CUDA(cudaStreamSynchronize(e->st));
      s.dev_maps.push_back(dmap);
    }
    s.devs.push_back(one_device(kind, a0, params, nparams, dmap, (int)nx, 0.0).d[0]);
  }
  return 0;
}

int chimera_engine_set_time(chimera_engine* e, double t) {
  ENG_ENTER(e);
  e->dev_time = t;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Moving window on the device (chimera_main.py:250-304 frame_act stage 1; NEXT-2 row)
// ------------------------------------------------------------------------------------------
// solvers.py:619-633 damp_field(config, damp_b = False): x-space window on E and G
int chimera_engine_damp_field(chimera_engine* e, const double* filtr, chb_i64 nxfilt, int mode) {
  ENG_ENTER(e);
  const auto& c = e->cfg;
  if (slab(e)) { set_error("damp_field needs every kx row (x-FFT): not available on a kx-slab engine"); return 2; }
  if (!filtr || nxfilt < 1 || nxfilt > c.nx || mode < 0 || mode > 2) { set_error("damp_field: bad filter (%lld points, mode %d)", nxfilt, mode); return 2; }
  e->scr.reset();
  double* d_f = e->scr.take_n<double>(nxfilt);
  if (!d_f) return 6;
  CHB_CUDA(cudaMemcpyAsync(d_f, filtr, sizeof(double) * nxfilt, cudaMemcpyDefault, e->st));
  FBCtx fb = fbctx(e);
  const i64 half = c.nx * c.nkr * c.nm * 3;
  CHB_TRY(fb_filtr_dev(fb, e->A("EG_fb"), c.leftX, e->D("kx"), d_f, mode, c.nx, c.nkr, c.nm, nxfilt));
  CHB_TRY(fb_filtr_dev(fb, e->A("EG_fb") + half, c.leftX, e->D("kx"), d_f, mode, c.nx, c.nkr, c.nm, nxfilt));
  return 0;
}

// Solver.damp_field on a kx-slab engine.  The x-space window needs every kx row of a column, so the slabs of all ranks
// are brought together first: the caller all-gathers "EG_fb" into "EG_gath" ([rank][(nx_slab, nkr, nm, 6)], allocated by
// chimera_engine_damp_prepare together with the full-row scratch "EG_full" and filled with the full "kx_full");
// this call rebuilds the full rows, applies the same fb_filtr as the unsharded engine and keeps this rank's rows.
int chimera_engine_damp_prepare(chimera_engine* e) {
  ENG_ENTER(e);
  const auto& c = e->cfg;
  if (!slab(e)) { set_error("damp_prepare: only for kx-slab engines"); return 2; }
  const size_t full = sizeof(cd) * (size_t)c.nx * c.nkr * c.nm * 6;
  if (!e->arr.count("EG_gath")) CHB_TRY(alloc_named(e, "EG_gath", full));
  if (!e->arr.count("EG_full")) CHB_TRY(alloc_named(e, "EG_full", full));
  if (!e->arr.count("kx_full")) CHB_TRY(alloc_named(e, "kx_full", sizeof(double) * (size_t)c.nx));
  return 0;
}

int chimera_engine_damp_field_slab(chimera_engine* e, const double* filtr, chb_i64 nxfilt, int mode) {
  ENG_ENTER(e);
  const auto& c = e->cfg;
  if (!slab(e)) { set_error("damp_field_slab: only for kx-slab engines"); return 2; }
  if (!e->arr.count("EG_gath") || !e->arr.count("EG_full") || !e->arr.count("kx_full")) {
    set_error("damp_field_slab: call chimera_engine_damp_prepare, upload kx_full and all-gather EG_fb into EG_gath first");
    return 2;
  }
  if (!filtr || nxfilt < 1 || nxfilt > c.nx || mode < 0 || mode > 2) { set_error("damp_field: bad filter (%lld points, mode %d)", nxfilt, mode); return 2; }
  e->scr.reset();
  double* d_f = e->scr.take_n<double>(nxfilt);
  if (!d_f) return 6;
  CHB_CUDA(cudaMemcpyAsync(d_f, filtr, sizeof(double) * nxfilt, cudaMemcpyDefault, e->st));
  FBCtx fb = fbctx(e);
  const i64 ncols = c.nkr * c.nm * 6, half = c.nx * c.nkr * c.nm * 3;
  CHB_TRY(rows_scatter_dev(fb, e->A("EG_full"), e->A("EG_gath"), (const i64*)e->arr["gather_map"].p, c.nx, nxs(e), ncols));
  CHB_TRY(fb_filtr_dev(fb, e->A("EG_full"), c.leftX, e->D("kx_full"), d_f, mode, c.nx, c.nkr, c.nm, nxfilt));
  CHB_TRY(fb_filtr_dev(fb, e->A("EG_full") + half, c.leftX, e->D("kx_full"), d_f, mode, c.nx, c.nkr, c.nm, nxfilt));
  CHB_TRY(rows_take_dev(fb, e->A("EG_fb"), e->A("EG_full"), (const i64*)e->arr["slab_rows"].p, c.nx, nxs(e), ncols));
  return 0;
}

// chimera_main.py:286-290 move_frame: Xgrid += shiftX
int chimera_engine_set_window(chimera_engine* e, double shift_stage1, double shift_stage2) {
  ENG_ENTER(e);
  e->win_s1 = shift_stage1;
  e->win_s2 = shift_stage2;
  return 0;
}

int chimera_engine_move_window(chimera_engine* e, double shiftX) {
  ENG_ENTER(e);
  e->cfg.leftX += shiftX;
  e->cfg.rightX += shiftX;
  return 0;
}

static int species_reserve(chimera_engine* e, Species& s, i64 need) {
  if (need <= s.cap) return 0;
  i64 ncap = s.cap * 3 / 2;
  if (ncap < need) ncap = need;
  ncap = (ncap + 31) & ~31LL;
  const size_t D = sizeof(double);
  double** cur[4] = {&s.x, &s.xh, &s.p, &s.w};
  double** alt[4] = {&s.x2, &s.xh2, &s.p2, &s.w2};
  for (int k = 0; k < 4; ++k) {
    const int nc = k < 3 ? 3 : 1;
    double* nb = nullptr;
    CHB_CUDA(cudaMalloc((void**)&nb, D * nc * ncap));
    for (int cpt = 0; cpt < nc; ++cpt)
      if (s.np > 0) CHB_CUDA(cudaMemcpyAsync(nb + cpt * ncap, *cur[k] + cpt * s.cap, D * s.np, cudaMemcpyDeviceToDevice, e->st));
    CHB_CUDA(cudaStreamSynchronize(e->st));
    cudaFree(*cur[k]);
    *cur[k] = nb;
    cudaFree(*alt[k]);
    CHB_CUDA(cudaMalloc((void**)alt[k], D * nc * ncap));
  }
  s.cap = ncap;
  return 0;
}

// species.py:218-244 add_particles: new particles go to the end, coords_halfstep = coords.  The particle order
// is no longer binned afterwards: a re-binning (chimera_engine_sort) must follow before the next deposit.
int chimera_engine_append_particles(chimera_engine* e, int id, const double* coords, const double* momenta,
                                    const double* weights, chb_i64 n) {
  ENG_ENTER(e);
  if (id < 0 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  if (n < 0) { set_error("append_particles: n < 0"); return 2; }
  if (n == 0) return 0;
  if (!coords || !momenta || !weights) { set_error("append_particles: null buffer"); return 2; }
  Species& s = e->sp[id];
  CHB_TRY(species_reserve(e, s, s.np + n));
  const double* srcs[2] = {coords, momenta};
  for (int k = 0; k < 2; ++k) {  // stage the (3,n) arrays through the permutation buffer, transpose behind the old particles
    CHB_CUDA(cudaMemcpyAsync(s.x2, srcs[k], sizeof(double) * 3 * n, cudaMemcpyDefault, e->st));
    double* dst = k == 0 ? s.x : s.p;
    aos_to_soa_k<<<grid_for(3 * n, 256), 256, 0, e->st>>>(dst + s.np, s.x2, 3, s.cap, n);
    CHB_LAUNCH_CHECK();
    if (k == 0) {
      aos_to_soa_k<<<grid_for(3 * n, 256), 256, 0, e->st>>>(s.xh + s.np, s.x2, 3, s.cap, n);
      CHB_LAUNCH_CHECK();
    }
  }
  CHB_CUDA(cudaMemcpyAsync(s.w + s.np, weights, sizeof(double) * n, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  s.np += n;
  s.tile_w = 0;  // unsorted
  const int nchnk = e->cfg.chunked ? e->cfg.nchnk : 1;
  for (int c2 = 1; c2 <= nchnk; ++c2) s.h_ind[c2] = (int)s.np;
  CHB_CUDA(cudaMemcpyAsync(s.d_ind, s.h_ind.data(), sizeof(int) * (nchnk + 1), cudaMemcpyHostToDevice, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return update_cta_table(e, s);
}

// species.py:351-398 chunk_and_damp with an explicit absorbing layer: particles left of leftX + left_margin, right
// of rightX or beyond the radial limit are dropped, the rest re-binned (on coords_halfstep when on_halfstep != 0)
int chimera_engine_sort(chimera_engine* e, int on_halfstep, double left_margin) {
  ENG_ENTER(e);
  return ph_sort(e, on_halfstep, left_margin);
}
// the same with the radial limit of the call given explicitly: a window's damp_plasma culls at the SPECIES' upperR
// (species.py:92,376: its r grid has one node less than the solver's), not at the solver's like the per-step re-binning
int chimera_engine_sort_window(chimera_engine* e, int on_halfstep, double left_margin, double upper_r2) {
  ENG_ENTER(e);
  return ph_sort(e, on_halfstep, left_margin, upper_r2);
}

// ------------------------------------------------------------------------------------------
// Integrated diagnostics on the device (moduls/diagnostics.py; NEXT-3 row)
// ------------------------------------------------------------------------------------------
// diagnostics.py:109-124 nrg_out before its roll: out[kx] = sum_{kr,m} EnergyFact * sum_{c<3} |EG_fb[..,c]|^2.
// energy_fact (nx, nkr, nm) float64, host or device, is kept in HBM after the first call (pass NULL later).
int chimera_engine_field_energy(chimera_engine* e, const double* energy_fact, double* out) {
  ENG_ENTER(e);
  const auto& c = e->cfg;
  const i64 nkx = nxs(e), ncols = c.nkr * c.nm;
  if (!out) { set_error("field_energy: null output"); return 2; }
  if (energy_fact) {
    if (!e->arr.count("EnergyFact")) CHB_TRY(alloc_named(e, "EnergyFact", sizeof(double) * nkx * ncols, false));
    CHB_CUDA(cudaMemcpyAsync(e->D("EnergyFact"), energy_fact, sizeof(double) * nkx * ncols, cudaMemcpyDefault, e->st));
  } else if (!e->arr.count("EnergyFact")) { set_error("field_energy: the EnergyFact table was never supplied"); return 2; }
  e->scr.reset();
  double* d_out = e->scr.take_n<double>(nkx);
  if (!d_out) return 6;
  CHB_TRY(launch_field_energy(e->st, e->A("EG_fb"), e->D("EnergyFact"), d_out, nkx, ncols));
  CHB_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * nkx, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

// diagnostics.py:174-207 get_beam_envelops: the sums it is built from, on coords_halfstep and momenta
int chimera_engine_beam_moments(chimera_engine* e, int id, double* out16) {
  ENG_ENTER(e);
  if (id < 0 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  Species& s = e->sp[id];
  e->scr.reset();
  double* d = e->scr.take_n<double>(16);
  if (!d) return 6;
  CHB_TRY(launch_beam_moments(e->st, s.xh, s.p, s.w, s.cap, s.np, d));
  CHB_CUDA(cudaMemcpyAsync(out16, d, sizeof(double) * 16, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

int chimera_engine_spectrum(chimera_engine* e, int id, int quantity, double lo, double hi, chb_i64 nbins, double* hist) {
  ENG_ENTER(e);
  if (id < 0 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  if (nbins < 1 || nbins > 4096 || !(hi > lo) || quantity < 0 || quantity > 1) { set_error("spectrum: bad binning"); return 2; }
  Species& s = e->sp[id];
  e->scr.reset();
  double* d = e->scr.take_n<double>(nbins);
  if (!d) return 6;
  CHB_TRY(launch_spectrum(e->st, s.p, s.w, s.cap, s.np, quantity, lo, hi, (int)nbins, d));
  CHB_CUDA(cudaMemcpyAsync(hist, d, sizeof(double) * nbins, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

// out[ix] = A[ix, ir, m, l] of a complex grid array (e.g. "EB", ir = 0: the on-axis wake field)
int chimera_engine_lineout(chimera_engine* e, const char* name, chb_i64 ir, chb_i64 m, chb_i64 l, double* out) {
  ENG_ENTER(e);
  NamedArray* a;
  CHB_TRY(find_array(e, name, &a));
  const auto& c = e->cfg;
  const std::string n = name;
  const bool grid = n == "J" || n == "Rho" || n == "BckGrndRho" || n == "EB";
  const i64 nx = grid ? c.nx : nxs(e), nr = grid ? c.nrn : c.nkr;
  const i64 off = nx * (ir + nr * (m + c.nm * l));
  if (ir < 0 || ir >= nr || m < 0 || m >= c.nm || l < 0 || (size_t)(off + nx) * sizeof(cd) > a->bytes) {
    set_error("lineout: index (%lld, %lld, %lld) outside '%s'", ir, m, l, name);
    return 2;
  }
  e->scr.reset();
  cd* d = e->scr.take_n<cd>(nx);
  if (!d) return 6;
  CHB_TRY(launch_lineout(e->st, (const cd*)a->p, d, nx, off));
  CHB_CUDA(cudaMemcpyAsync(out, d, sizeof(cd) * nx, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

int chimera_engine_species_count(chimera_engine* e, int id, chb_i64* np) {
  ENG_ENTER(e);
  if (id < 0 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  *np = e->sp[id].np;
  return 0;
}

int chimera_engine_get_species(chimera_engine* e, int id, double* coords, double* coords_half, double* momenta,
                               double* weights) {
  ENG_ENTER(e);
  if (id < 0 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  Species& s = e->sp[id];
  if (s.np == 0) return 0;
  double* dsts[3] = {coords, coords_half, momenta};
  const double* srcs[3] = {s.x, s.xh, s.p};
  for (int k = 0; k < 3; ++k) {
    if (!dsts[k]) continue;
    soa_to_aos_k<<<grid_for(3 * s.np, 256), 256, 0, e->st>>>(s.x2, srcs[k], 3, s.cap, s.np);
    CHB_LAUNCH_CHECK();
    CHB_CUDA(cudaMemcpyAsync(dsts[k], s.x2, sizeof(double) * 3 * s.np, cudaMemcpyDefault, e->st));
    CHB_CUDA(cudaStreamSynchronize(e->st));
  }
  if (weights) CHB_CUDA(cudaMemcpyAsync(weights, s.w, sizeof(double) * s.np, cudaMemcpyDefault, e->st));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

int chimera_engine_get_chunks(chimera_engine* e, int id, int* ind) {
  ENG_ENTER(e);
  if (id < 0 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  const int nchnk = e->cfg.chunked ? e->cfg.nchnk : 1;
  for (int c = 0; c <= nchnk; ++c) ind[c] = e->sp[id].h_ind[c];
  return 0;
}

int chimera_engine_run(chimera_engine* e, int phase, double arg) {
  ENG_ENTER(e);
  return run_phase(e, phase, arg);
}

// ---- CUDA graph of one fused step --------------------------------------------------------------------------------
static int fused_step_phases(chimera_engine* e) {
  const auto& c = e->cfg;
  CHB_TRY(run_phase(e, CHB_PARTICLES_FUSED, 1));
  CHB_TRY(run_phase(e, CHB_FB_IN_J, 0));
  if (c.space_charge) CHB_TRY(run_phase(e, CHB_FB_IN_RHO, 0));
  CHB_TRY(run_phase(e, CHB_POISSON, 0));
  CHB_TRY(run_phase(e, CHB_MAXWELL, 0));
  return run_phase(e, CHB_FIELDS_OUT, 0);
}

static void drop_graphs(chimera_engine* e) {
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

// Kernel arguments are frozen in a graph: usable only while nothing they depend on changes from step to step -- no
// window that moves every step (leftX), no time-dependent device list, no per-phase event timing.
static bool graph_usable(chimera_engine* e) {
  if (!e->use_graph || e->graph_state < 0 || e->profile || gemm_profile_enabled() || slab(e)) return false;
  if (e->win_s1 != 0.0 || e->win_s2 != 0.0) return false;
  for (auto& s : e->sp)
    if (!s.devs.empty()) return false;
  return true;
}

static unsigned long long graph_signature(chimera_engine* e) {
  unsigned long long h = 1469598103934665603ull;
  auto mix = [&](unsigned long long v) { h = (h ^ v) * 1099511628211ull; };
  mix(e->scr.gen);
  mix((unsigned long long)(e->scr.blocks.empty() ? nullptr : e->scr.blocks[0].p));
  mix((unsigned long long)e->st);
  for (auto& s : e->sp) { mix((unsigned long long)s.np); mix((unsigned long long)s.ncta_f); mix((unsigned long long)s.x); }
  return h;
}

static int step_graphed(chimera_engine* e) {
  const unsigned long long sig = graph_signature(e);
  if (sig != e->graph_sig) {  // particle counts, buffers or the scratch block changed: the recorded arguments are stale
    drop_graphs(e);
    e->graph_sig = sig;
  }
  if (e->graph_warm_gen != e->scr.gen) {  // one eager step sizes the scratch and creates the FFT plans
    CHB_TRY(fused_step_phases(e));
    e->graph_warm_gen = e->scr.gen;
    return 0;
  }
  e->graph_state = 1;
  const void* nxt = e->cfg.space_charge ? (const void*)e->A("gradRho_fb_nxt") : nullptr;
  for (auto& g : e->graphs)
    if (g.nxt == nxt) {
      CHB_CUDA(cudaGraphLaunch(g.exec, e->st));
      // the host-side part of the phases
      if (e->cfg.space_charge) std::swap(e->arr["gradRho_fb_prv"], e->arr["gradRho_fb_nxt"]);
      e->leftX_J = NAN;
      g_launches += g.launches;
      return 0;
    }
  const long long l0 = g_launches;
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    e->graph_state = -1;
    return fused_step_phases(e);
  }
  const int rc = fused_step_phases(e);  // launches are recorded, nothing runs; the host-side state advances
  const cudaError_t ce = cudaStreamEndCapture(e->st, &graph);
  cudaGraphExec_t exec = nullptr;
  if (rc == 0 && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
    cudaGraphDestroy(graph);
    e->graphs.push_back({exec, nxt, g_launches - l0});
    CHB_CUDA(cudaGraphLaunch(exec, e->st));
    return 0;
  }
  // capture failed: undo the host-side swap, switch graphs off for this engine and run the step eagerly
  cudaGetLastError();
  if (graph) cudaGraphDestroy(graph);
  if (e->cfg.space_charge && (const void*)e->A("gradRho_fb_nxt") != nxt) std::swap(e->arr["gradRho_fb_prv"], e->arr["gradRho_fb_nxt"]);
  e->graph_state = -1;
  return fused_step_phases(e);
}

int chimera_engine_set_graph(chimera_engine* e, int on) {
  ENG_CHECK(e);
  e->use_graph = on ? 1 : 0;
  if (!on) { drop_graphs(e); e->graph_sig = 0; }
  else if (e->graph_state < 0) e->graph_state = 0;
  return 0;
}

// 'StaticKick' across ranks: PXmean of every species (chimera_main.py:121-122) from the all-reduced beam moments, used
// by the next CHB_STATIC_FIELDS phase instead of the local reduction (NaN: species empty everywhere)
int chimera_engine_set_static_px(chimera_engine* e, const double* px, int n) {
  ENG_ENTER(e);
  e->static_px.assign(px, px + (n > 0 ? n : 0));
  return 0;
}

// buffers of the column-block dataflow for `world` ranks (kx-slab engines only)
int chimera_engine_set_colflow(chimera_engine* e, int rank, int world) {
  ENG_ENTER(e);
  if (rank < 0 || rank >= world) { set_error("column dataflow: bad rank %d of %d", rank, world); return 2; }
  e->col_rank = rank;
  const auto& c = e->cfg;
  if (!slab(e)) { set_error("column dataflow: kx-slab engines only"); return 2; }
  if (world < 2 || (size_t)world > kColPad || nxs(e) * world != c.nx) { set_error("column dataflow: bad rank count %d", world); return 2; }
  e->col_world = world;
  const size_t C = sizeof(cd), nx = (size_t)c.nx, L = (size_t)nxs(e);
  const size_t cj = (size_t)col_block(e, c.nrn * c.nm * 3), cr = (size_t)col_block(e, c.nrn * c.nm), ce = (size_t)col_block(e, c.nrn * c.nm * 6);
  struct { const char* n; size_t b; } spec[] = {
      {"J_blk", nx * cj * C}, {"J_send", nx * cj * C}, {"J_in", L * cj * world * C},
      {"Rho_blk", nx * cr * C}, {"Rho_send", nx * cr * C}, {"Rho_in", L * cr * world * C},
      {"EB_recv", nx * ce * C}, {"EB_blk", nx * ce * C}};
  for (auto& s : spec)
    if (!e->arr.count(s.n)) CHB_TRY(alloc_named(e, s.n, s.b));
  return 0;
}

int chimera_engine_graph_info(chimera_engine* e, int* ngraphs, int* state) {
  ENG_CHECK(e);
  if (ngraphs) *ngraphs = (int)e->graphs.size();
  if (state) *state = e->graph_state;
  return 0;
}

int chimera_engine_step(chimera_engine* e, chb_i64 istep0, chb_i64 nsteps) {
  ENG_CHECK(e);
  const auto& c = e->cfg;
  if (slab(e)) { set_error("kx-slab engine: sequence the phases from the host (all-gather between fields_out_a / _b)"); return 2; }
  // The particle work between two field solves -- gather + push_velocs of step k, then push_coords + deposits
  // of step k+1 -- is independent per particle, so inside a multi-step call it runs as ONE kernel
  // (CHB_PARTICLES_FUSED); the first step's head, the last step's tail and re-binning steps (the sort sits
  // between push_coords and the deposits, chimera_main.py:82-86) use the separate phases.
  // The tail of the call's last step stays pending (tail_pending) and becomes part of the next call's first fused
  // kernel: a caller that steps one at a time (diagnostics every step) runs the same kernels as one long call.
  if (nsteps <= 0) return 0;
  bool gather_pending = e->tail_pending;
  e->tail_pending = false;
  for (i64 k = 0; k < nsteps; ++k) {
    const i64 istep = istep0 + k;
    const bool sort_now = c.sort_every > 0 && istep % c.sort_every == 0;
    // the pending gather + push closes the step before (make_device(istep - 1)); carried over from an earlier call it
    // keeps the time of that call's last step
    e->dev_time = (k == 0 && gather_pending) ? e->tail_time : (double)(istep - 1) * c.dt;
    if (gather_pending && !sort_now && e->fuse && !c.static_kick) {
      if (graph_usable(e)) {  // the whole step, spectral update included, as one graph launch
        CHB_TRY(step_graphed(e));
        continue;
      }
      CHB_TRY(run_phase(e, CHB_PARTICLES_FUSED, 1));
    } else if (e->fuse) {
      // re-binning step (the sort sits between push_coords and the deposits) or the first step of the call: one
      // streaming kernel before the sort, one deposit kernel after it; both apply their window stage
      if (gather_pending) CHB_TRY(run_phase(e, CHB_GATHER_PUSH_COORDS, 0));
      else {
        CHB_TRY(ph_window(e, 1));  // frame_act(istep) (chimera_main.py:83)
        CHB_TRY(run_phase(e, CHB_PUSH_COORDS, 0));
      }
      if (sort_now) CHB_TRY(run_phase(e, CHB_SORT, 1));
      CHB_TRY(run_phase(e, CHB_DEPOSIT_FUSED, 1));
    } else {
      if (gather_pending) CHB_TRY(run_phase(e, CHB_GATHER_PUSH, 1.0));
      const bool win = e->win_s1 != 0.0 || e->win_s2 != 0.0;
      if (win) CHB_TRY(ph_window(e, 1));  // frame_act(istep) (chimera_main.py:83)
      CHB_TRY(run_phase(e, CHB_PUSH_COORDS, 0));
      if (sort_now) CHB_TRY(run_phase(e, CHB_SORT, 1));
      CHB_TRY(run_phase(e, CHB_DEPOSIT_J, 0));
      if (win) CHB_TRY(ph_window(e, 2));  // frame_act(istep, 'stage2') (chimera_main.py:87)
      if (c.space_charge || c.static_kick) CHB_TRY(run_phase(e, CHB_DEPOSIT_RHO, 1));
    }
    CHB_TRY(run_phase(e, CHB_FB_IN_J, 0));
    if (c.space_charge || c.static_kick) CHB_TRY(run_phase(e, CHB_FB_IN_RHO, 0));
    if (c.static_kick) {
      CHB_TRY(run_phase(e, CHB_STATIC_FIELDS, 0));
    } else {
      CHB_TRY(run_phase(e, CHB_POISSON, 0));
      CHB_TRY(run_phase(e, CHB_MAXWELL, 0));
    }
    CHB_TRY(run_phase(e, CHB_FIELDS_OUT, 0));
    gather_pending = true;
  }
  e->dev_time = (double)(istep0 + nsteps - 1) * c.dt;
  if (e->lazy_tail && e->fuse) {
    e->tail_pending = true;
    e->tail_time = e->dev_time;
    return 0;
  }
  return run_phase(e, CHB_GATHER_PUSH, 1.0);
}

int chimera_engine_set_lazy_tail(chimera_engine* e, int on) {
  ENG_ENTER(e);
  e->lazy_tail = on ? 1 : 0;
  return 0;
}

// ------------------------------------------------------------------------------------------
// One make_step with the whole PIC state in HOST buffers (the reference's calling model: numpy owns
// every array, moduls/chimera_main.py:82-92), pipelined over three streams so that PCIe runs in both
// directions while the device computes:
//   h2d stream : coords / momenta / weights in NCHUNK pieces, then EG_fb and gradRho_fb_nxt
//   engine     : per piece transpose to SoA + push_coords (+ transpose back); then deposit .. gather+push
//   d2h stream : coords / coords_halfstep pieces as soon as they are pushed, EG_fb after the PSATD
//                advance, momenta after the Boris push
// ------------------------------------------------------------------------------------------
static int host_mark(chimera_engine* e, cudaStream_t on, cudaStream_t waiter) {
  cudaEvent_t ev;
  CHB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  e->host_evs.push_back(ev);
  CHB_CUDA(cudaEventRecord(ev, on));
  CHB_CUDA(cudaStreamWaitEvent(waiter, ev, 0));
  return 0;
}

int chimera_engine_step_host_begin(chimera_engine* e, int id, double* coords, double* coords_half, double* momenta,
                                   double* weights, chb_i64 np, double* EG_fb, double* gradRho_fb_nxt, chb_i64 istep,
                                   int rebin) {
  ENG_ENTER(e);
  if (id < 0 || id >= (int)e->sp.size()) { set_error("bad species id %d", id); return 2; }
  Species& s = e->sp[id];
  if (s.still) { set_error("step_host: species %d is still", id); return 2; }
  if (e->cfg.static_kick) { set_error("step_host: no 'StaticKick' schedule; use chimera_engine_step"); return 2; }
  if (e->win_s1 != 0.0 || e->win_s2 != 0.0) { set_error("step_host: no per-step window; use chimera_engine_step"); return 2; }
  if (np < 0 || np > s.cap) { set_error("step_host: np=%lld exceeds the species capacity %lld", np, s.cap); return 2; }
  // coords_half may be NULL: the centred positions are used inside the step only (deposit, re-binning) and a caller that
  // lets the engine do both has no use for them on the host -- 24 B per particle less on the device->host link
  if (!coords || !momenta || !weights) { set_error("step_host: null particle buffer"); return 2; }
  const auto& c = e->cfg;
  CHB_TRY(ensure_ops(e));
  if (!e->s_h2d) {
    CHB_CUDA(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
    CHB_CUDA(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
  }
  auto mark = [&](cudaStream_t on, cudaStream_t waiter) -> int { return host_mark(e, on, waiter); };
  const bool sort_now = rebin || (c.sort_every > 0 && istep % c.sort_every == 0);
  e->dev_time = (double)istep * c.dt;
  if (np != s.np && !sort_now) { set_error("step_host: particle count changed (%lld -> %lld) without rebin", s.np, np); return 2; }
  s.np = np;
  const size_t D = sizeof(double);
  // the copy streams start after whatever the engine stream did before
  CHB_TRY(mark(e->st, e->s_h2d));
  CHB_TRY(mark(e->st, e->s_d2h));
  constexpr int NCHUNK = 16;
  const i64 csz = ((np + NCHUNK - 1) / NCHUNK + 31) & ~31LL;
  for (i64 a = 0; a < np; a += csz) {
    const i64 n = (np - a < csz) ? np - a : csz;
    CHB_CUDA(cudaMemcpyAsync(s.x2 + 3 * a, coords + 3 * a, D * 3 * n, cudaMemcpyHostToDevice, e->s_h2d));
    CHB_CUDA(cudaMemcpyAsync(s.p2 + 3 * a, momenta + 3 * a, D * 3 * n, cudaMemcpyHostToDevice, e->s_h2d));
    CHB_CUDA(cudaMemcpyAsync(s.w + a, weights + a, D * n, cudaMemcpyHostToDevice, e->s_h2d));
    g_h2d_bytes += (long long)(D * 7 * n);
    CHB_TRY(mark(e->s_h2d, e->st));
    aos_to_soa_k<<<grid_for(3 * n, 256), 256, 0, e->st>>>(s.x + a, s.x2 + 3 * a, 3, s.cap, n);
    CHB_LAUNCH_CHECK();
    aos_to_soa_k<<<grid_for(3 * n, 256), 256, 0, e->st>>>(s.p + a, s.p2 + 3 * a, 3, s.cap, n);
    CHB_LAUNCH_CHECK();
    CHB_TRY(launch_push_coords(e->st, soa(s.x + a, s.cap), soa((const double*)(s.p + a), s.cap), soa(s.xh + a, s.cap), c.dt, n));
    if (!sort_now) {
      soa_to_aos_k<<<grid_for(3 * n, 256), 256, 0, e->st>>>(s.x2 + 3 * a, s.x + a, 3, s.cap, n);
      CHB_LAUNCH_CHECK();
      if (coords_half) {
        soa_to_aos_k<<<grid_for(3 * n, 256), 256, 0, e->st>>>(s.xh2 + 3 * a, s.xh + a, 3, s.cap, n);
        CHB_LAUNCH_CHECK();
      }
      CHB_TRY(mark(e->st, e->s_d2h));
      CHB_CUDA(cudaMemcpyAsync(coords + 3 * a, s.x2 + 3 * a, D * 3 * n, cudaMemcpyDeviceToHost, e->s_d2h));
      if (coords_half)
        CHB_CUDA(cudaMemcpyAsync(coords_half + 3 * a, s.xh2 + 3 * a, D * 3 * n, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)(D * (coords_half ? 6 : 3) * n);
    }
  }
  // spectral state of the solver: needed from the density transform on
  NamedArray *aEG = nullptr, *aG = nullptr;
  CHB_TRY(find_array(e, "EG_fb", &aEG));
  if (EG_fb) {
    CHB_CUDA(cudaMemcpyAsync(aEG->p, EG_fb, aEG->bytes, cudaMemcpyHostToDevice, e->s_h2d));
    g_h2d_bytes += (long long)aEG->bytes;
  }
  if (gradRho_fb_nxt && c.space_charge) {
    CHB_TRY(find_array(e, "gradRho_fb_nxt", &aG));
    CHB_CUDA(cudaMemcpyAsync(aG->p, gradRho_fb_nxt, aG->bytes, cudaMemcpyHostToDevice, e->s_h2d));
    g_h2d_bytes += (long long)aG->bytes;
  }
  if (sort_now) {
    CHB_TRY(run_phase(e, CHB_SORT, 1));  // synchronises the engine stream; s.np may shrink
    const i64 m = s.np;
    if (m > 0) {
      soa_to_aos_k<<<grid_for(3 * m, 256), 256, 0, e->st>>>(s.x2, s.x, 3, s.cap, m);
      CHB_LAUNCH_CHECK();
      if (coords_half) {
        soa_to_aos_k<<<grid_for(3 * m, 256), 256, 0, e->st>>>(s.xh2, s.xh, 3, s.cap, m);
        CHB_LAUNCH_CHECK();
      }
      CHB_TRY(mark(e->st, e->s_d2h));
      CHB_CUDA(cudaMemcpyAsync(coords, s.x2, D * 3 * m, cudaMemcpyDeviceToHost, e->s_d2h));
      if (coords_half) CHB_CUDA(cudaMemcpyAsync(coords_half, s.xh2, D * 3 * m, cudaMemcpyDeviceToHost, e->s_d2h));
      CHB_CUDA(cudaMemcpyAsync(weights, s.w, D * m, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)(D * (coords_half ? 7 : 4) * m);
    }
  }
  CHB_TRY(run_phase(e, CHB_DEPOSIT_J, 0));
  if (c.space_charge) CHB_TRY(run_phase(e, CHB_DEPOSIT_RHO, e->host_rho_from_bg));
  e->host_EG = EG_fb;
  e->host_G = (gradRho_fb_nxt && c.space_charge) ? gradRho_fb_nxt : nullptr;
  e->host_mom = momenta;
  e->host_id = id;
  return 0;
}

// kx-slab mode only: field update on this rank's slab and the first half of fields out (-> "EB_slab")
int chimera_engine_step_host_mid(chimera_engine* e) {
  ENG_ENTER(e);
  if (e->host_id < 0) { set_error("step_host_mid without step_host_begin"); return 2; }
  const auto& c = e->cfg;
  auto mark = [&](cudaStream_t on, cudaStream_t waiter) -> int { return host_mark(e, on, waiter); };
  NamedArray *aEG = nullptr, *aG = nullptr;
  CHB_TRY(find_array(e, "EG_fb", &aEG));
  if (e->host_G) CHB_TRY(find_array(e, "gradRho_fb_nxt", &aG));
  CHB_TRY(mark(e->s_h2d, e->st));  // EG_fb / gradRho_fb_nxt have landed
  CHB_TRY(run_phase(e, CHB_FB_IN_J, 0));
  if (c.space_charge) CHB_TRY(run_phase(e, CHB_FB_IN_RHO, 0));
  CHB_TRY(run_phase(e, CHB_POISSON, 0));
  CHB_TRY(run_phase(e, CHB_MAXWELL, 0));
  if (e->host_EG || aG) {
    CHB_TRY(mark(e->st, e->s_d2h));
    if (e->host_EG) {
      CHB_CUDA(cudaMemcpyAsync(e->host_EG, aEG->p, aEG->bytes, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)aEG->bytes;
    }
    if (aG) {
      CHB_CUDA(cudaMemcpyAsync(e->host_G, aG->p, aG->bytes, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)aG->bytes;
    }
  }
  CHB_TRY(run_phase(e, CHB_FIELDS_OUT_A, 0));
  e->host_mid_done = 1;
  return 0;
}

// second half: (the caller may all-reduce J / Rho on the engine stream in between) transforms, Poisson
// correction, PSATD advance, fields out, gather + push, copies out; synchronises everything
int chimera_engine_step_host_end(chimera_engine* e, chb_i64* np_out) {
  ENG_ENTER(e);
  if (e->host_id < 0) { set_error("step_host_end without step_host_begin"); return 2; }
  Species& s = e->sp[e->host_id];
  const auto& c = e->cfg;
  const size_t D = sizeof(double);
  auto mark = [&](cudaStream_t on, cudaStream_t waiter) -> int { return host_mark(e, on, waiter); };
  double* EG_fb = e->host_EG;
  double* gradRho_fb_nxt = e->host_G;
  double* momenta = e->host_mom;
  NamedArray *aEG = nullptr, *aG = nullptr;
  CHB_TRY(find_array(e, "EG_fb", &aEG));
  if (gradRho_fb_nxt) CHB_TRY(find_array(e, "gradRho_fb_nxt", &aG));
  if (e->host_mid_done) {  // chimera_engine_step_host_mid already ran the field update and the first half of fields out
    CHB_TRY(run_phase(e, CHB_FIELDS_OUT_B, 0));
  } else {
  if (slab(e)) { set_error("kx-slab engine: call step_host_mid, all-gather EB_slab into EB_gath, then step_host_end"); return 2; }
  CHB_TRY(mark(e->s_h2d, e->st));  // EG_fb / gradRho_fb_nxt have landed
  CHB_TRY(run_phase(e, CHB_FB_IN_J, 0));
  if (c.space_charge) CHB_TRY(run_phase(e, CHB_FB_IN_RHO, 0));
  CHB_TRY(run_phase(e, CHB_POISSON, 0));
  CHB_TRY(run_phase(e, CHB_MAXWELL, 0));
  if (EG_fb || aG) {
    CHB_TRY(mark(e->st, e->s_d2h));
    if (EG_fb) {
      CHB_CUDA(cudaMemcpyAsync(EG_fb, aEG->p, aEG->bytes, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)aEG->bytes;
    }
    if (aG) {
      CHB_CUDA(cudaMemcpyAsync(gradRho_fb_nxt, aG->p, aG->bytes, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)aG->bytes;
    }
  }
  CHB_TRY(run_phase(e, CHB_FIELDS_OUT, 0));
  }
  e->host_mid_done = 0;
  // gather + push in pieces of whole CTAs (consecutive CTAs own consecutive particles), each piece's momenta
  // copied out while the next piece is being pushed
  bool piecewise = false;
  if (s.np > 0 && s.ncta >= 2 && (c.nm == 1 || c.nm == 2 || c.nm == 3 || c.nm == 4 || c.nm == 5)) {
    piecewise = true;
    GridGeom g = geom_ready(e);
    const DeviceSet und = devset(e, s);
    const int nchnk = (int)s.h_ind.size() - 1;
    auto first_of = [&](int cta) -> i64 {  // first particle of CTA `cta` (== np for cta == ncta)
      if (cta >= s.ncta) return s.np;
      int ck = 0;
      while (ck + 1 < nchnk && cta >= s.h_cta[ck + 1]) ++ck;
      return (i64)s.h_ind[ck] + (i64)(cta - s.h_cta[ck]) * kDepNPB;
    };
    constexpr int NPIECE = 8;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e->profile) { e0 = get_event(e); e1 = get_event(e); cudaEventRecord(e0, e->st); }
    for (int k = 0; k < NPIECE; ++k) {
      const int c0 = (int)((i64)s.ncta * k / NPIECE), c1 = (int)((i64)s.ncta * (k + 1) / NPIECE);
      if (c1 <= c0) continue;
      SortedSpec spk = sortedspec(e, s);
      spk.cta_base = c0;
      spk.ncta = c1 - c0;
      CHB_TRY(launch_gather_push_binned(e->st, c.env, s.x, s.w, e->A("EB"), s.p, s.cap, g, s.push_fact * c.dt, und, spk));
      const i64 a = first_of(c0), b = first_of(c1), n = b - a;
      if (n <= 0) continue;
      soa_to_aos_k<<<grid_for(3 * n, 256), 256, 0, e->st>>>(s.p2 + 3 * a, s.p + a, 3, s.cap, n);
      CHB_LAUNCH_CHECK();
      CHB_TRY(mark(e->st, e->s_d2h));
      CHB_CUDA(cudaMemcpyAsync(momenta + 3 * a, s.p2 + 3 * a, D * 3 * n, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)(D * 3 * n);
    }
    if (e->profile) { cudaEventRecord(e1, e->st); e->pending.push_back({CHB_GATHER_PUSH, {e0, e1}}); }
  }
  if (!piecewise) {
    CHB_TRY(run_phase(e, CHB_GATHER_PUSH, 1.0));
    if (s.np > 0) {
      soa_to_aos_k<<<grid_for(3 * s.np, 256), 256, 0, e->st>>>(s.p2, s.p, 3, s.cap, s.np);
      CHB_LAUNCH_CHECK();
      CHB_TRY(mark(e->st, e->s_d2h));
      CHB_CUDA(cudaMemcpyAsync(momenta, s.p2, D * 3 * s.np, cudaMemcpyDeviceToHost, e->s_d2h));
      g_d2h_bytes += (long long)(D * 3 * s.np);
    }
  }
  CHB_CUDA(cudaStreamSynchronize(e->s_h2d));
  CHB_CUDA(cudaStreamSynchronize(e->s_d2h));
  CHB_CUDA(cudaStreamSynchronize(e->st));
  for (auto ev : e->host_evs) cudaEventDestroy(ev);
  e->host_evs.clear();
  e->host_id = -1;
  if (np_out) *np_out = s.np;
  return 0;
}

int chimera_engine_step_host(chimera_engine* e, int id, double* coords, double* coords_half, double* momenta,
                             double* weights, chb_i64 np, chb_i64* np_out, double* EG_fb, double* gradRho_fb_nxt,
                             chb_i64 istep, int rebin) {
  CHB_TRY(chimera_engine_step_host_begin(e, id, coords, coords_half, momenta, weights, np, EG_fb, gradRho_fb_nxt, istep, rebin));
  return chimera_engine_step_host_end(e, np_out);
}

int chimera_engine_set_fuse(chimera_engine* e, int on) {
  ENG_ENTER(e);
  e->fuse = on ? 1 : 0;
  return 0;
}

int chimera_engine_set_rho_from_bg(chimera_engine* e, int from_bg) {
  ENG_ENTER(e);
  e->host_rho_from_bg = from_bg ? 1.0 : 0.0;
  return 0;
}

int chimera_host_register(void* ptr, chb_i64 nbytes) {
  CHB_CUDA(cudaHostRegister(ptr, (size_t)nbytes, cudaHostRegisterPortable));
  return 0;
}
int chimera_host_unregister(void* ptr) {
  CHB_CUDA(cudaHostUnregister(ptr));
  return 0;
}

int chimera_engine_sync(chimera_engine* e) {
  ENG_ENTER(e);
  CHB_CUDA(cudaStreamSynchronize(e->st));
  return 0;
}

int chimera_engine_set_stream(chimera_engine* e, void* cuda_stream) {
  ENG_ENTER(e);
  CHB_CUDA(cudaStreamSynchronize(e->st));
  if (e->own_stream) cudaStreamDestroy(e->st);
  e->st = (cudaStream_t)cuda_stream;
  e->own_stream = false;
  return 0;
}

int chimera_engine_profile(chimera_engine* e, int on) {
  ENG_CHECK(e);
  CHB_TRY(collect_timings(e));
  e->profile = on;
  return 0;
}

int chimera_engine_timings(chimera_engine* e, double* ms, chb_i64* calls, int reset) {
  ENG_CHECK(e);
  CHB_TRY(collect_timings(e));
  for (int i = 0; i < CHB_NPHASES; ++i) {
    if (ms) ms[i] = e->ms[i];
    if (calls) calls[i] = e->calls[i];
    if (reset) { e->ms[i] = 0; e->calls[i] = 0; }
  }
  return 0;
}

}  // extern "C"
