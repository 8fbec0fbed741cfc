// particles_sorted.cu -- particle kernels of the resident engine that exploit the engine's particle
// order: structure-of-arrays storage, re-binned by (x-chunk, x-tile, r-cell, x-cell) (engine.cu).
//
//  * deposit_runs_k   charge/current deposition (grid_deps*.f90 dep_curr/dep_dens and their
//                     _chnk/_env variants).  Neighbouring particles of the sorted order share a cell,
//                     so a thread walks RUN consecutive particles, accumulates the per-mode complex
//                     node values of the current cell in registers and issues ONE red.global.add.f64
//                     per node value when the cell changes: ~10x fewer L2 atomics than one per
//                     particle (the L2 atomic units were the limiter: ncu, profiles/r01a_*).
//  * gather_push_tiled_k  proj_fld[_env] + external device + push_velocs fused: a CTA stages the
//                     (x, r) tile of all 6 x nm field planes its particles touch in shared memory
//                     (coalesced rows), the per-particle 4-node x 6 x nm complex reads then hit
//                     shared memory; particles that drifted out of the tile fall back to L2.
#include "common.cuh"
#include "kernels.cuh"
#include "particle_dev.cuh"

namespace chb {

namespace {
constexpr int RUN = 16;  // consecutive particles per thread in deposit_runs_k

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

template <int NM>
struct NodeAcc {
  cd a[2][2][NM];  // [x node][r node][mode slot]
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int m = 0; m < NM; ++m) a[i][k][m] = cmake(0.0, 0.0);
  }
};

template <int ENV, int CURR, int NM>
__global__ void __launch_bounds__(192)
deposit_runs_k(const double* __restrict__ x, const double* __restrict__ mom, const double* __restrict__ w, i64 cap,
               cd* __restrict__ grid, GridGeom g, ChunkSpec ch, i64 np) {
  constexpr int NC = CURR ? (ENV ? 1 : 3) : 1;  // components per particle run (Q1: env current has l=3 only)
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 run = t / NC;
  const int lc = (int)(t - run * NC);
  const int l = CURR ? (ENV ? 2 : lc) : 0;
  const i64 ip0 = run * RUN;
  if (ip0 >= np) return;
  constexpr int NKO = ENV ? (NM - 1) / 2 : NM - 1;
  const i64 plane = g.nxn * g.nrn;
  cd* const gl = grid + plane * g.nm * l;

  NodeAcc<NM> acc;
  acc.zero();
  i64 cix = -1, cir = -1;
  int cchunk = 0;
  i64 chunk_lo = 0, chunk_hi = ch.on ? 0 : np;  // particle index range of the chunk `cchunk`
  bool dirty = false;

  auto flush = [&]() {
    if (!dirty) return;
    bool keep[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const i64 gx = cix + i;
      keep[i] = (gx >= 0 && gx <= g.nxn - 1) && (ch.on ? chunk_keep(ch, cchunk, gx, g.nxn) : true);
    }
    cd* pl = gl + cix + g.nxn * cir;
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (!keep[i]) continue;
#pragma unroll
        for (int k = 0; k < 2; ++k) red_add(pl + plane * m + i + g.nxn * k, acc.a[i][k][m]);
      }
    acc.zero();
    dirty = false;
  };

#pragma unroll 1
  for (int j4 = 0; j4 < RUN; j4 += 4) {
    const i64 ipb = ip0 + j4;
    if (ipb >= np) break;
    // 4 particles per batch: one 32-byte sector per array and lane
    double xs[4], ys[4], zs[4], ws[4], ps[3][4];
    {
      double2 a = ldg2(x + ipb), b = ldg2(x + ipb + 2);
      xs[0] = a.x; xs[1] = a.y; xs[2] = b.x; xs[3] = b.y;
      a = ldg2(x + cap + ipb); b = ldg2(x + cap + ipb + 2);
      ys[0] = a.x; ys[1] = a.y; ys[2] = b.x; ys[3] = b.y;
      a = ldg2(x + 2 * cap + ipb); b = ldg2(x + 2 * cap + ipb + 2);
      zs[0] = a.x; zs[1] = a.y; zs[2] = b.x; zs[3] = b.y;
      a = ldg2(w + ipb); b = ldg2(w + ipb + 2);
      ws[0] = a.x; ws[1] = a.y; ws[2] = b.x; ws[3] = b.y;
      if (CURR) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          a = ldg2(mom + c * cap + ipb); b = ldg2(mom + c * cap + ipb + 2);
          ps[c][0] = a.x; ps[c][1] = a.y; ps[c][2] = b.x; ps[c][3] = b.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const i64 ip = ipb + j;
      if (ip >= np) break;
      const double wp = ws[j];
      if (wp == 0.0) continue;
      const double xp = xs[j], yp = ys[j], zp = zs[j];
      Shape s;
      if (!make_shape(g, xp, yp, zp, s)) continue;
      double v = 1.0;
      if (CURR) {
        const double p0 = ps[0][j], p1 = ps[1][j], p2 = ps[2][j];
        if (fabs(p0) + fabs(p1) + fabs(p2) == 0.0) continue;
        const double gp = sqrt(1.0 + p0 * p0 + p1 * p1 + p2 * p2);
        const double pl_ = (l == 0) ? p0 : (l == 1 ? p1 : p2);
        v = ENV ? pl_ / gp : pl_ * wp / gp;
      }
      if (ch.on && (ip < chunk_lo || ip >= chunk_hi)) {
        int lo = 0, hi = ch.nchnk;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (ip >= __ldg(ch.ind + mid)) lo = mid; else hi = mid;
        }
        if (ip >= __ldg(ch.ind + ch.nchnk)) break;  // beyond the last chunk: not deposited by the reference
        flush();
        cchunk = lo;
        chunk_lo = __ldg(ch.ind + lo);
        chunk_hi = __ldg(ch.ind + lo + 1);
      }
      if (s.ix != cix || s.ir != cir) {
        flush();
        cix = s.ix;
        cir = s.ir;
      }
      // complex weight of the particle: 1 | w (density) and the envelope carrier exp(-i kx0 x)
      cd base = CURR ? cmake(v, 0.0) : cmake(wp, 0.0);
      if (ENV) {
        double sn, cs;
        sincos(xp * g.kx0, &sn, &cs);
        const cd wpc = cmake(wp * cs, -wp * sn);
        base = CURR ? cscale(v, wpc) : cmul(wpc, wpc);  // Q2: the density weight enters twice
      }
      const cd ph1 = (s.rp > 0.0) ? cmake(yp / s.rp, -zp / s.rp) : cmake(0.0, 0.0);
      const double w00 = s.sx0 * s.sr0, w01 = s.sx0 * s.sr1, w10 = s.sx1 * s.sr0, w11 = s.sx1 * s.sr1;
      cd ph = cmake(1.0, 0.0);
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
        for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
          const cd phs = sgn ? cconj(ph) : ph;
          const int slot = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
          const cd f = cmul(base, phs);
          acc.a[0][0][slot].x += w00 * f.x; acc.a[0][0][slot].y += w00 * f.y;
          acc.a[0][1][slot].x += w01 * f.x; acc.a[0][1][slot].y += w01 * f.y;
          acc.a[1][0][slot].x += w10 * f.x; acc.a[1][0][slot].y += w10 * f.y;
          acc.a[1][1][slot].x += w11 * f.x; acc.a[1][1][slot].y += w11 * f.y;
        }
      }
      dirty = true;
    }
  }
  flush();
}

template <int ENV, int CURR>
int launch_runs_nm(cudaStream_t st, const double* x, const double* mom, const double* w, i64 cap, cd* grid,
                   const GridGeom& g, const ChunkSpec& ch, i64 np) {
  constexpr int NC = CURR ? (ENV ? 1 : 3) : 1;
  const i64 nruns = (np + RUN - 1) / RUN;
  const unsigned nb = grid_for(nruns * NC, 192);
#define CHB_RUNS(NMV)                                                                                   \
  case NMV:                                                                                             \
    deposit_runs_k<ENV, CURR, NMV><<<nb, 192, 0, st>>>(x, mom, w, cap, grid, g, ch, np);                \
    break;
  switch ((int)g.nm) {
    CHB_RUNS(1)
    CHB_RUNS(2)
    CHB_RUNS(3)
    CHB_RUNS(4)
    CHB_RUNS(5)
    default: return -1;  // caller falls back to the direct kernel
  }
#undef CHB_RUNS
  CHB_LAUNCH_CHECK();
  return 0;
}
}  // namespace

int launch_deposit_runs(cudaStream_t st, int env, int curr, CPView x, CPView mom, const double* w, cd* grid,
                        const GridGeom& g, const ChunkSpec& ch, i64 np) {
  if (np <= 0) return 0;
  const bool soa_ok = x.ps == 1 && (!curr || mom.ps == 1) && (x.cs % 2 == 0) && (!curr || mom.cs == x.cs) &&
                      (env ? (g.nm % 2 == 1) : true);
  int rc = -1;
  if (soa_ok) {
    if (env && curr) rc = launch_runs_nm<1, 1>(st, x.p, mom.p, w, x.cs, grid, g, ch, np);
    else if (env)    rc = launch_runs_nm<1, 0>(st, x.p, nullptr, w, x.cs, grid, g, ch, np);
    else if (curr)   rc = launch_runs_nm<0, 1>(st, x.p, mom.p, w, x.cs, grid, g, ch, np);
    else             rc = launch_runs_nm<0, 0>(st, x.p, nullptr, w, x.cs, grid, g, ch, np);
  }
  if (rc == -1) return launch_deposit_direct(st, env, curr, x, mom, w, grid, g, ch, np, false);
  return rc;
}

// ------------------------------------------------------------------------------------------------
namespace {
constexpr int GT_THREADS = 256, GT_PPT = 4, GT_NPB = GT_THREADS * GT_PPT;

template <int ENV>
__device__ __forceinline__ void gather_tile(const GridGeom& g, const cd* __restrict__ tile, int tx, int trx, int kx,
                                            int kr, const Shape& s, double xp, double yp, double zp, double F[6]) {
  const int nko = ENV ? (int)(g.nm - 1) / 2 : (int)g.nm - 1;
  const cd ph1 = (s.rp > 0.0) ? cmake(yp / s.rp, zp / s.rp) : (ENV ? cmake(1.0, 0.0) : cmake(0.0, 0.0));  // Q4
  cd car = cmake(1.0, 0.0);
  if (ENV) {
    double sn, cs;
    sincos(xp * g.kx0, &sn, &cs);
    car = cmake(cs, sn);
  }
  const double w00 = s.sr0 * s.sx0, w10 = s.sr0 * s.sx1, w01 = s.sr1 * s.sx0, w11 = s.sr1 * s.sx1;
  const int node = kx + tx * kr;
#pragma unroll
  for (int l = 0; l < 6; ++l) F[l] = 0.0;
  cd ph = cmake(1.0, 0.0);
  for (int iO = 0; iO <= nko; ++iO) {
    if (iO > 0) ph = cmul(ph, ph1);
    for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
      const cd phs = sgn ? cconj(ph) : ph;
      const int slot = ENV ? (nko + (sgn ? -iO : iO)) : iO;
      const cd p00 = cmul(cscale(w00, car), phs), p10 = cmul(cscale(w10, car), phs);
      const cd p01 = cmul(cscale(w01, car), phs), p11 = cmul(cscale(w11, car), phs);
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        const cd* pl = tile + trx * (slot + (int)g.nm * l) + node;
        const cd f00 = pl[0], f10 = pl[1], f01 = pl[tx], f11 = pl[tx + 1];
        double a = 0.0;
        a += p00.x * f00.x - p00.y * f00.y;
        a += p10.x * f10.x - p10.y * f10.y;
        a += p01.x * f01.x - p01.y * f01.y;
        a += p11.x * f11.x - p11.y * f11.y;
        F[l] += a;
      }
    }
  }
}

template <int ENV>
__global__ void __launch_bounds__(GT_THREADS, 2)
gather_push_tiled_k(const double* __restrict__ x, const double* __restrict__ w, const cd* __restrict__ Fld,
                    double* __restrict__ mom, i64 cap, GridGeom g, double dt_2, DeviceSet und, i64 np, int TX, int TR) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* tile = reinterpret_cast<cd*>(smem_raw);  // [6*nm][TR][TX]
  __shared__ int s_box[4];                     // min ix, min ir, max ix, max ir
  const int tid = threadIdx.x;
  const i64 base = (i64)blockIdx.x * GT_NPB;
  if (tid == 0) { s_box[0] = 0x7fffffff; s_box[1] = 0x7fffffff; s_box[2] = -1; s_box[3] = -1; }
  __syncthreads();

  double xp[GT_PPT], yp[GT_PPT], zp[GT_PPT];
  bool ok[GT_PPT];
  int mnx = 0x7fffffff, mnr = 0x7fffffff, mxx = -1, mxr = -1;
#pragma unroll
  for (int j = 0; j < GT_PPT; ++j) {
    const i64 ip = base + (i64)j * GT_THREADS + tid;
    ok[j] = false;
    if (ip < np) {
      xp[j] = __ldg(x + ip); yp[j] = __ldg(x + cap + ip); zp[j] = __ldg(x + 2 * cap + ip);
      Shape s;
      if (__ldg(w + ip) != 0.0 && make_shape(g, xp[j], yp[j], zp[j], s) && s.ix >= 0 && s.ix <= g.nxn - 2) {
        ok[j] = true;
        mnx = min(mnx, (int)s.ix); mxx = max(mxx, (int)s.ix);
        mnr = min(mnr, (int)s.ir); mxr = max(mxr, (int)s.ir);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mnr = min(mnr, __shfl_xor_sync(0xffffffffu, mnr, o));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxr = max(mxr, __shfl_xor_sync(0xffffffffu, mxr, o));
  }
  if ((tid & 31) == 0 && mxx >= 0) {
    atomicMin(&s_box[0], mnx); atomicMin(&s_box[1], mnr); atomicMax(&s_box[2], mxx); atomicMax(&s_box[3], mxr);
  }
  __syncthreads();
  const int ix0 = s_box[0], ir0 = s_box[1];
  const bool any = s_box[2] >= 0;
  // rows/columns of nodes actually needed, clipped to the tile capacity
  const int ncol = any ? min(TX, s_box[2] - ix0 + 2) : 0;
  const int nrow = any ? min(TR, s_box[3] - ir0 + 2) : 0;
  const int trx = TX * TR;
  const int nplanes = 6 * (int)g.nm;
  const i64 plane = g.nxn * g.nrn;
  // stage the tile: one (plane, row) strip per warp iteration, lanes along x (coalesced)
  for (int strip = tid >> 5; strip < nplanes * nrow; strip += GT_THREADS / 32) {
    const int q = strip / nrow, kr = strip - q * nrow;
    const i64 gr = ir0 + kr;
    const cd* src = Fld + plane * q + g.nxn * gr + ix0;
    cd* dst = tile + trx * q + TX * kr;
    for (int kx = tid & 31; kx < ncol; kx += 32) {
      const i64 gx = ix0 + kx;
      dst[kx] = (gx < g.nxn && gr < g.nrn) ? __ldg(src + kx) : cmake(0.0, 0.0);
    }
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < GT_PPT; ++j) {
    const i64 ip = base + (i64)j * GT_THREADS + tid;
    if (ip >= np) continue;
    double F[6] = {0, 0, 0, 0, 0, 0};
    if (ok[j]) {
      Shape s;
      make_shape(g, xp[j], yp[j], zp[j], s);  // recomputed rather than kept in registers across the barrier
      const int kx = (int)s.ix - ix0, kr = (int)s.ir - ir0;
      if (kx + 1 < ncol && kr + 1 < nrow) gather_tile<ENV>(g, tile, TX, trx, kx, kr, s, xp[j], yp[j], zp[j], F);
      else gather_one<ENV>(g, Fld, xp[j], yp[j], zp[j], F);
    }
    if (und.n) apply_devices(und, xp[j], yp[j], zp[j], F);
    double px = mom[ip], py = mom[cap + ip], pz = mom[2 * cap + ip];
    boris(px, py, pz, F[0], F[1], F[2], F[3], F[4], F[5], dt_2);
    mom[ip] = px; mom[cap + ip] = py; mom[2 * cap + ip] = pz;
  }
}
}  // namespace

int launch_gather_push_tiled(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView mom,
                             const GridGeom& g, double dt, const DeviceSet& und, i64 np) {
  if (np <= 0) return 0;
  if (x.ps != 1 || mom.ps != 1 || mom.cs != x.cs) return launch_gather_push(st, env, x, w, Fld, mom, g, dt, und, np);
  // tile capacity: x extent of one re-binning tile (32 cells) plus drift margin, rows by a 56 KB budget
  const int TX = 44;
  const int planes = 6 * (int)g.nm;
  int TR = (int)((56 * 1024) / ((size_t)planes * TX * sizeof(cd)));
  if (TR > 8) TR = 8;
  if (TR < 3) return launch_gather_push(st, env, x, w, Fld, mom, g, dt, und, np);
  const size_t smem = (size_t)planes * TX * TR * sizeof(cd);
  static bool attr_set = false;
  if (!attr_set) {
    CHB_CUDA(cudaFuncSetAttribute(gather_push_tiled_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CHB_CUDA(cudaFuncSetAttribute(gather_push_tiled_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  const unsigned nb = grid_for(np, GT_NPB);
  if (env) gather_push_tiled_k<1><<<nb, GT_THREADS, smem, st>>>(x.p, w, Fld, mom.p, x.cs, g, 0.5 * dt, und, np, TX, TR);
  else     gather_push_tiled_k<0><<<nb, GT_THREADS, smem, st>>>(x.p, w, Fld, mom.p, x.cs, g, 0.5 * dt, und, np, TX, TR);
  CHB_LAUNCH_CHECK();
  return 0;
}

// ================================================================================================
// CTA-binned particle kernels (see kernels.cuh).  A CTA takes kDepNPB consecutive particles of ONE
// x-chunk and
//  (A) computes every particle's cell inside a 64 x 32 cell box anchored at the CTA's first particle
//      and histograms the cells in shared memory,
//  (B) block-scans the histogram (cell -> first sorted slot) and cuts every cell's particles into
//      SEGMENTS of at most 16: a segment's particles share a cell, so whatever belongs to the cell is
//      held in registers for the whole segment,
//  (C) writes a per-particle record (shape fractions, azimuthal phase, amplitudes) to its sorted slot,
//  (D) runs one thread per (segment, component):
//        deposit: 4 nodes x nm modes of complex accumulators in registers, ONE red.global.add.f64
//                 per node value and segment (instead of one per particle);
//        gather : the 4 x nm node values of the component are loaded ONCE per segment, the
//                 per-particle field goes to shared memory,
//  (E) gather only: one thread per particle: external device, Boris push.
// Particles that drifted out of the box since the last re-binning take the direct path (L2 atomics
// / L2 loads); the result is identical, only slower.
// ================================================================================================
}  // namespace chb
#include <cub/block/block_scan.cuh>
namespace chb {
namespace {
constexpr int DB_THREADS = 192, DB_RUN = 16;
constexpr int DB_BX = 64, DB_BR = 32, DB_BINS = DB_BX * DB_BR;
constexpr int DB_PPT = (kDepNPB + DB_THREADS - 1) / DB_THREADS;
constexpr int DB_MAXTASK = kDepNPB / DB_RUN + kDepNPB;  // worst case: every particle alone in its cell

template <int THREADS>
struct BinShared {  // static shared memory of the binning stage
  int anchor[2];
  int total, ntask;
  typename cub::BlockScan<int, THREADS>::TempStorage scan;
};

struct CtaRange {
  int chunk;
  i64 first;
  int count;
};

__device__ __forceinline__ CtaRange cta_range(const SortedSpec& sp) {
  const int bid = (int)blockIdx.x + sp.cta_base;
  int lo = 0, hi = sp.nchnk;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (bid >= __ldg(sp.cta + mid)) lo = mid; else hi = mid;
  }
  CtaRange r;
  r.chunk = lo;
  r.first = (i64)__ldg(sp.ind + lo) + (i64)(bid - __ldg(sp.cta + lo)) * kDepNPB;
  const i64 n = (i64)__ldg(sp.ind + lo + 1) - r.first;
  r.count = (int)(n < kDepNPB ? n : kDepNPB);
  return r;
}

// anchor of the CTA's cell box: the re-binning tile of the first particle, centred in the 64-cell
// window, and 6 rows below the first particle's r-cell
__device__ __forceinline__ void cta_anchor(const double* __restrict__ x, i64 cap, const GridGeom& g, const SortedSpec& sp,
                                           const CtaRange& cr, int* anchor) {
  const double xp = __ldg(x + cr.first), yp = __ldg(x + cap + cr.first), zp = __ldg(x + 2 * cap + cr.first);
  const i64 ix = (i64)floor((xp - g.leftX) * g.dx_inv);
  const i64 ir = (i64)floor((sqrt(yp * yp + zp * zp) - g.r0) * g.dr_inv);
  i64 ax = ix - DB_BX / 2;
  if (sp.tile_w > 0) {
    i64 lx = ix - (i64)cr.chunk * sp.cs;
    lx = lx < 0 ? 0 : (lx > sp.cs - 1 ? sp.cs - 1 : lx);
    ax = (i64)cr.chunk * sp.cs + (lx / sp.tile_w) * sp.tile_w - (DB_BX - sp.tile_w) / 2;
  }
  anchor[0] = (int)ax;
  anchor[1] = (int)(ir - 6);
}

// stage (B): packed scan (low 16 bits: particles, high 16 bits: segments) and the segment table
//   task word = key | start << 11 | n << 22
template <int THREADS>
__device__ __forceinline__ void cta_scan_and_tasks(int* bins, int* tasks, BinShared<THREADS>& sh, int tid) {
  constexpr int DB_ITEMS = (DB_BINS + THREADS - 1) / THREADS;
  int items[DB_ITEMS], cnt[DB_ITEMS];
  int total = 0;
#pragma unroll
  for (int i = 0; i < DB_ITEMS; ++i) {
    const int b = tid * DB_ITEMS + i;
    cnt[i] = (b < DB_BINS) ? bins[b] : 0;
    items[i] = cnt[i] | (((cnt[i] + DB_RUN - 1) / DB_RUN) << 16);
  }
  cub::BlockScan<int, THREADS>(sh.scan).ExclusiveSum(items, items, total);
#pragma unroll
  for (int i = 0; i < DB_ITEMS; ++i) {
    const int b = tid * DB_ITEMS + i;
    if (b < DB_BINS) {
      bins[b] = items[i] & 0xFFFF;
      int start = items[i] & 0xFFFF, t = items[i] >> 16, left = cnt[i];
      while (left > 0) {
        const int n = left < DB_RUN ? left : DB_RUN;
        tasks[t++] = b | (start << 11) | (n << 22);
        start += n;
        left -= n;
      }
    }
  }
  if (tid == 0) { sh.total = total & 0xFFFF; sh.ntask = total >> 16; }
}

// x-node range a particle of chunk `c` may write (the chunk-edge rule of grid_deps_chnk.f90:95-115,
// identical to chunk_keep() for every node, evaluated once per CTA)
__device__ __forceinline__ void keep_range(const ChunkSpec& ch, int c, i64 nxn, i64& lo, i64& hi) {
  lo = 0;
  hi = nxn - 1;
  if (ch.on) {
    const i64 left = (i64)c * ch.cs;
    const i64 l2 = (left - ch.guards >= 0) ? left - ch.guards : left + 1;
    const i64 h2 = (left + ch.cs + ch.guards <= nxn - 1) ? left + ch.cs + ch.guards : left + ch.cs - 1;
    lo = l2 > lo ? l2 : lo;
    hi = h2 < hi ? h2 : hi;
  }
}

template <int ENV, int CURR>
struct DepLayout {
  static constexpr int NAMP = ENV ? 2 : (CURR ? 3 : 1);
  static constexpr int NF = 4 + NAMP;
  static constexpr size_t smem =
      sizeof(double) * NF * kDepNPB + sizeof(int) * (DB_BINS + DB_MAXTASK) + 2 * sizeof(unsigned short) * kDepNPB;
};

template <int ENV, int CURR, int NM>
__global__ void __launch_bounds__(DB_THREADS, 3)
deposit_binned_k(const double* __restrict__ x, const double* __restrict__ mom, const double* __restrict__ w, i64 cap,
                 cd* __restrict__ grid, GridGeom g, ChunkSpec ch, SortedSpec sp) {
  using L = DepLayout<ENV, CURR>;
  constexpr int NF = L::NF;
  constexpr int NC = CURR ? (ENV ? 1 : 3) : 1;   // components (Q1: the envelope current has l = 3 only)
  constexpr int SL = CURR ? NM : 1;              // mode slots per thread: all for J, one for rho
  constexpr int NS = NM / SL;                    // threads per (segment, component)
  constexpr int NKO = ENV ? (NM - 1) / 2 : NM - 1;
  constexpr int HB = DB_PPT / 2;                 // particles per load batch
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* rec = reinterpret_cast<double*>(smem_raw);            // [NF][kDepNPB], indexed by local particle id
  int* bins = reinterpret_cast<int*>(rec + NF * kDepNPB);       // [DB_BINS]
  int* tasks = bins + DB_BINS;                                  // [DB_MAXTASK]
  unsigned short* skey = reinterpret_cast<unsigned short*>(tasks + DB_MAXTASK);  // [kDepNPB] local id -> cell key
  unsigned short* order = skey + kDepNPB;                       // [kDepNPB] sorted slot -> local id
  __shared__ BinShared<DB_THREADS> sh;
  const int tid = threadIdx.x;

  const CtaRange cr = cta_range(sp);
  if (cr.count <= 0) return;
  for (int i = tid; i < DB_BINS; i += DB_THREADS) bins[i] = 0;
  if (tid == 0) cta_anchor(x, cap, g, sp, cr, sh.anchor);
  __syncthreads();
  const int ix0 = sh.anchor[0], ir0 = sh.anchor[1];

  // ---- (A) per-particle record, cell key, histogram.  Loads of a batch are issued together.
#pragma unroll 1
  for (int jb = 0; jb < DB_PPT; jb += HB) {
    double xs[HB], ys[HB], zs[HB], ws[HB], ps[3][HB];
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      const int li = tid + (jb + j) * DB_THREADS;
      const bool in = li < cr.count;
      const i64 ip = cr.first + (in ? li : 0);
      xs[j] = __ldg(x + ip); ys[j] = __ldg(x + cap + ip); zs[j] = __ldg(x + 2 * cap + ip);
      ws[j] = in ? __ldg(w + ip) : 0.0;
      if (CURR) { ps[0][j] = __ldg(mom + ip); ps[1][j] = __ldg(mom + cap + ip); ps[2][j] = __ldg(mom + 2 * cap + ip); }
    }
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      const int li = tid + (jb + j) * DB_THREADS;
      if (li >= kDepNPB) continue;
      unsigned short key = 0xFFFFu;
      const double wp = ws[j], xp = xs[j], yp = ys[j], zp = zs[j];
      Shape s;
      if (wp != 0.0 && make_shape(g, xp, yp, zp, s)) {
        const i64 kx = s.ix - ix0, kr = s.ir - ir0;
        if (kx >= 0 && kx < DB_BX && kr >= 0 && kr < DB_BR) {
          key = (unsigned short)(kr * DB_BX + kx);
          atomicAdd(&bins[key], 1);
          rec[li] = s.sx1;
          rec[kDepNPB + li] = s.sr1;
          const double rinv = (s.rp > 0.0) ? 1.0 / s.rp : 0.0;  // deposit phase exp(-i theta); 0 on the axis
          rec[2 * kDepNPB + li] = yp * rinv;
          rec[3 * kDepNPB + li] = -zp * rinv;
          double ginv = 1.0;
          if (CURR) ginv = 1.0 / sqrt(1.0 + ps[0][j] * ps[0][j] + ps[1][j] * ps[1][j] + ps[2][j] * ps[2][j]);
          if (ENV) {
            double sn, cs;
            sincos(xp * g.kx0, &sn, &cs);
            const cd wpc = cmake(wp * cs, -wp * sn);
            const cd base = CURR ? cscale(ps[2][j] * ginv, wpc) : cmul(wpc, wpc);  // Q1: l = 3 only; Q2: weight twice
            rec[4 * kDepNPB + li] = base.x;
            rec[5 * kDepNPB + li] = base.y;
          } else if (CURR) {
            const double wg = wp * ginv;
            rec[4 * kDepNPB + li] = ps[0][j] * wg;
            rec[5 * kDepNPB + li] = ps[1][j] * wg;
            rec[6 * kDepNPB + li] = ps[2][j] * wg;
          } else {
            rec[4 * kDepNPB + li] = wp;
          }
        } else {  // drifted out of the box: straight to the grid
          deposit_one<ENV, CURR>(g, ch, cr.chunk, grid, xp, yp, zp, CURR ? ps[0][j] : 0.0, CURR ? ps[1][j] : 0.0,
                                 CURR ? ps[2][j] : 0.0, wp);
        }
      }
      skey[li] = key;
    }
  }
  __syncthreads();
  // ---- (B)
  cta_scan_and_tasks(bins, tasks, sh, tid);
  __syncthreads();
  // ---- (C) local ids to sorted slots
#pragma unroll
  for (int j = 0; j < DB_PPT; ++j) {
    const int li = tid + j * DB_THREADS;
    if (li < kDepNPB) {
      const int key = skey[li];
      if (key != 0xFFFF) order[atomicAdd(&bins[key], 1)] = (unsigned short)li;
    }
  }
  __syncthreads();

  // ---- (D) one thread per (segment, component[, mode slot]): node accumulators in registers
  const i64 plane = g.nxn * g.nrn;
  i64 klo, khi;
  keep_range(ch, cr.chunk, g.nxn, klo, khi);
  const int ntask = sh.ntask;
#pragma unroll 1
  for (int t = tid; t < ntask * NC * NS; t += DB_THREADS) {
    const int task = t / (NC * NS), sub = t - task * (NC * NS);
    const int lc = sub / NS, s0 = (sub - lc * NS) * SL;  // first mode slot of this thread
    const int l = CURR ? (ENV ? 2 : lc) : 0;
    const int tw = tasks[task];
    const int key = tw & 0x7FF, start = (tw >> 11) & 0x7FF, n = tw >> 22;
    cd a[2][2][SL];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int m = 0; m < SL; ++m) a[i][k][m] = cmake(0.0, 0.0);
    const double* ra = rec + (4 + (ENV ? 0 : (CURR ? l : 0))) * kDepNPB;
#pragma unroll 2
    for (int q = 0; q < n; ++q) {
      const int li = order[start + q];
      const double fx = rec[li], fr = rec[kDepNPB + li];
      const cd ph1 = cmake(rec[2 * kDepNPB + li], rec[3 * kDepNPB + li]);
      const cd amp = ENV ? cmake(ra[li], ra[kDepNPB + li]) : cmake(ra[li], 0.0);
      const double w00 = (1.0 - fx) * (1.0 - fr), w01 = (1.0 - fx) * fr, w10 = fx * (1.0 - fr), w11 = fx * fr;
      cd phs[NM];  // exp(-i m theta) per mode slot
      cd ph = cmake(1.0, 0.0);
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
        if (ENV) { phs[NKO + iO] = ph; phs[NKO - iO] = cconj(ph); }
        else phs[iO] = ph;
      }
#pragma unroll
      for (int m = 0; m < SL; ++m) {
        cd pm = phs[0];
        if (SL == NM) pm = phs[m];
        else {
#pragma unroll
          for (int z = 1; z < NM; ++z) pm = (s0 == z) ? phs[z] : pm;
        }
        const cd f = ENV ? cmul(amp, pm) : cscale(amp.x, pm);
        a[0][0][m].x += w00 * f.x; a[0][0][m].y += w00 * f.y;
        a[0][1][m].x += w01 * f.x; a[0][1][m].y += w01 * f.y;
        a[1][0][m].x += w10 * f.x; a[1][0][m].y += w10 * f.y;
        a[1][1][m].x += w11 * f.x; a[1][1][m].y += w11 * f.y;
      }
    }
    const int kr = key / DB_BX, kx = key - kr * DB_BX;
    const i64 gx = (i64)ix0 + kx, gr = (i64)ir0 + kr;
    cd* pl = grid + plane * (g.nm * l + s0) + gx + g.nxn * gr;
    const bool k0 = gx >= klo && gx <= khi, k1 = gx + 1 >= klo && gx + 1 <= khi;
#pragma unroll
    for (int m = 0; m < SL; ++m) {
      if (k0) { red_add(pl + plane * m, a[0][0][m]); red_add(pl + plane * m + g.nxn, a[0][1][m]); }
      if (k1) { red_add(pl + plane * m + 1, a[1][0][m]); red_add(pl + plane * m + 1 + g.nxn, a[1][1][m]); }
    }
  }
}

template <int ENV, int CURR>
int launch_binned_nm(cudaStream_t st, const double* x, const double* mom, const double* w, i64 cap, cd* grid,
                     const GridGeom& g, const ChunkSpec& ch, const SortedSpec& sp) {
  const size_t smem = DepLayout<ENV, CURR>::smem;
#define CHB_BINNED(NMV)                                                                                          \
  case NMV: {                                                                                                    \
    static bool attr = false;                                                                                    \
    if (!attr) {                                                                                                 \
      CHB_CUDA(cudaFuncSetAttribute(deposit_binned_k<ENV, CURR, NMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                 \
      attr = true;                                                                                               \
    }                                                                                                            \
    deposit_binned_k<ENV, CURR, NMV><<<sp.ncta, DB_THREADS, smem, st>>>(x, mom, w, cap, grid, g, ch, sp);        \
  } break;
  switch ((int)g.nm) {
    CHB_BINNED(1)
    CHB_BINNED(2)
    CHB_BINNED(3)
    CHB_BINNED(4)
    CHB_BINNED(5)
    default: return -1;
  }
#undef CHB_BINNED
  CHB_LAUNCH_CHECK();
  return 0;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// Binned gather + push: stages (A)-(C) as above, (D) one thread per (segment, field component) keeps the
// 4 x nm node values of its cell and component in registers, (E) device + Boris push per particle.
namespace {
constexpr int GB_THREADS = 256, GB_PPT = kDepNPB / GB_THREADS;

template <int ENV>
struct GatLayout {
  static constexpr int NF = ENV ? 6 : 4;
  static constexpr size_t smem = sizeof(double) * (NF + 6) * kDepNPB + sizeof(int) * (DB_BINS + DB_MAXTASK) +
                                 2 * sizeof(unsigned short) * kDepNPB;
};

// OUT = 0: the gathered field goes into the Boris push (`mom` = momentum planes).  OUT = 1: it is added to the
// reference's per-particle array Fld_tot(6, np) (`mom` = that array, component fastest): proj_fld on its own, for the
// per-function entry points.
template <int ENV, int NM, int OUT>
__global__ void __launch_bounds__(GB_THREADS, 2)
gather_push_binned_k(const double* __restrict__ x, const double* __restrict__ w, const cd* __restrict__ Fld,
                     double* __restrict__ mom, i64 cap, GridGeom g, double dt_2, DeviceSet und, SortedSpec sp) {
  constexpr int NF = GatLayout<ENV>::NF;
  constexpr int NKO = ENV ? (NM - 1) / 2 : NM - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* rec = reinterpret_cast<double*>(smem_raw);            // [NF][kDepNPB] by local id
  double* fbuf = rec + NF * kDepNPB;                            // [6][kDepNPB] gathered field by local id
  int* bins = reinterpret_cast<int*>(fbuf + 6 * kDepNPB);
  int* tasks = bins + DB_BINS;
  unsigned short* skey = reinterpret_cast<unsigned short*>(tasks + DB_MAXTASK);
  unsigned short* order = skey + kDepNPB;
  __shared__ BinShared<GB_THREADS> sh;
  const int tid = threadIdx.x;

  const CtaRange cr = cta_range(sp);
  if (cr.count <= 0) return;
  for (int i = tid; i < DB_BINS; i += GB_THREADS) bins[i] = 0;
  if (tid == 0) cta_anchor(x, cap, g, sp, cr, sh.anchor);
  __syncthreads();
  const int ix0 = sh.anchor[0], ir0 = sh.anchor[1];

  // ---- (A)
  {
    double xs[GB_PPT], ys[GB_PPT], zs[GB_PPT], ws[GB_PPT];
#pragma unroll
    for (int j = 0; j < GB_PPT; ++j) {
      const int li = tid + j * GB_THREADS;
      const bool in = li < cr.count;
      const i64 ip = cr.first + (in ? li : 0);
      xs[j] = __ldg(x + ip); ys[j] = __ldg(x + cap + ip); zs[j] = __ldg(x + 2 * cap + ip);
      ws[j] = in ? __ldg(w + ip) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < GB_PPT; ++j) {
      const int li = tid + j * GB_THREADS;
      unsigned short key = 0xFFFFu;
      const double xp = xs[j], yp = ys[j], zp = zs[j];
      double F[6] = {0, 0, 0, 0, 0, 0};
      Shape s;
      if (ws[j] != 0.0 && make_shape(g, xp, yp, zp, s) && s.ix >= 0 && s.ix <= g.nxn - 2) {
        const i64 kx = s.ix - ix0, kr = s.ir - ir0;
        if (kx >= 0 && kx < DB_BX && kr >= 0 && kr < DB_BR) {
          key = (unsigned short)(kr * DB_BX + kx);
          atomicAdd(&bins[key], 1);
          rec[li] = s.sx1;
          rec[kDepNPB + li] = s.sr1;
          // gather phase exp(+i theta); on the axis 0 (real solver) or 1 (envelope solver), Q4
          const double rinv = (s.rp > 0.0) ? 1.0 / s.rp : 0.0;
          rec[2 * kDepNPB + li] = (s.rp > 0.0) ? yp * rinv : (ENV ? 1.0 : 0.0);
          rec[3 * kDepNPB + li] = zp * rinv;
          if (ENV) {
            double sn, cs;
            sincos(xp * g.kx0, &sn, &cs);
            rec[4 * kDepNPB + li] = cs;
            rec[5 * kDepNPB + li] = sn;
          }
        } else {
          gather_one<ENV>(g, Fld, xp, yp, zp, F);  // drifted out of the box
        }
      }
      if (key == 0xFFFFu) {
#pragma unroll
        for (int l = 0; l < 6; ++l) fbuf[l * kDepNPB + li] = F[l];
      }
      skey[li] = key;
    }
  }
  __syncthreads();
  cta_scan_and_tasks(bins, tasks, sh, tid);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < GB_PPT; ++j) {
    const int li = tid + j * GB_THREADS;
    const int key = skey[li];
    if (key != 0xFFFF) order[atomicAdd(&bins[key], 1)] = (unsigned short)li;
  }
  __syncthreads();

  // ---- (D)
  const i64 plane = g.nxn * g.nrn;
  const int ntask = sh.ntask;
#pragma unroll 1
  for (int t = tid; t < ntask * 6; t += GB_THREADS) {
    const int task = t / 6, l = t - task * 6;
    const int tw = tasks[task];
    const int key = tw & 0x7FF, start = (tw >> 11) & 0x7FF, n = tw >> 22;
    const int kr = key / DB_BX, kx = key - kr * DB_BX;
    const cd* pl = Fld + plane * g.nm * l + ((i64)ix0 + kx) + g.nxn * ((i64)ir0 + kr);
    cd N[NM][4];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      N[m][0] = __ldg(pl + plane * m); N[m][1] = __ldg(pl + plane * m + 1);
      N[m][2] = __ldg(pl + plane * m + g.nxn); N[m][3] = __ldg(pl + plane * m + g.nxn + 1);
    }
#pragma unroll 2
    for (int q = 0; q < n; ++q) {
      const int li = order[start + q];
      const double fx = rec[li], fr = rec[kDepNPB + li];
      const cd ph1 = cmake(rec[2 * kDepNPB + li], rec[3 * kDepNPB + li]);
      const double w00 = (1.0 - fr) * (1.0 - fx), w10 = (1.0 - fr) * fx, w01 = fr * (1.0 - fx), w11 = fr * fx;
      cd car = cmake(1.0, 0.0);
      if (ENV) car = cmake(rec[4 * kDepNPB + li], rec[5 * kDepNPB + li]);  // carrier exp(+i kx0 x)
      cd ph = cmake(1.0, 0.0);
      double F = 0.0;
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
        for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
          const int slot = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
          const cd pm = ENV ? cmul(car, sgn ? cconj(ph) : ph) : ph;
          const double sx = w00 * N[slot][0].x + w10 * N[slot][1].x + w01 * N[slot][2].x + w11 * N[slot][3].x;
          const double sy = w00 * N[slot][0].y + w10 * N[slot][1].y + w01 * N[slot][2].y + w11 * N[slot][3].y;
          F += pm.x * sx - pm.y * sy;
        }
      }
      fbuf[l * kDepNPB + li] = F;
    }
  }
  __syncthreads();

  // ---- (E) external device + Boris push
#pragma unroll
  for (int j = 0; j < GB_PPT; ++j) {
    const int li = tid + j * GB_THREADS;
    if (li >= cr.count) continue;
    const i64 ip = cr.first + li;
    double F[6];
#pragma unroll
    for (int l = 0; l < 6; ++l) F[l] = fbuf[l * kDepNPB + li];
    if (OUT) {
      if (skey[li] != 0xFFFFu || F[0] != 0.0 || F[1] != 0.0 || F[2] != 0.0 || F[3] != 0.0 || F[4] != 0.0 || F[5] != 0.0) {
#pragma unroll
        for (int l = 0; l < 6; ++l) mom[6 * ip + l] += F[l];
      }
      continue;
    }
    if (und.n) apply_devices(und, __ldg(x + ip), __ldg(x + cap + ip), __ldg(x + 2 * cap + ip), F);
    double px = mom[ip], py = mom[cap + ip], pz = mom[2 * cap + ip];
    boris(px, py, pz, F[0], F[1], F[2], F[3], F[4], F[5], dt_2);
    mom[ip] = px; mom[cap + ip] = py; mom[2 * cap + ip] = pz;
  }
}

template <int ENV, int OUT>
int launch_gather_binned_nm(cudaStream_t st, const double* x, const double* w, const cd* Fld, double* mom, i64 cap,
                            const GridGeom& g, double dt_2, const DeviceSet& und, const SortedSpec& sp) {
  const size_t smem = GatLayout<ENV>::smem;
#define CHB_GB(NMV)                                                                                               \
  case NMV: {                                                                                                     \
    static bool attr = false;                                                                                     \
    if (!attr) {                                                                                                  \
      CHB_CUDA(cudaFuncSetAttribute(gather_push_binned_k<ENV, NMV, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                  \
      attr = true;                                                                                                \
    }                                                                                                             \
    gather_push_binned_k<ENV, NMV, OUT><<<sp.ncta, GB_THREADS, smem, st>>>(x, w, Fld, mom, cap, g, dt_2, und, sp); \
  } break;
  switch ((int)g.nm) {
    CHB_GB(1)
    CHB_GB(2)
    CHB_GB(3)
    CHB_GB(4)
    CHB_GB(5)
    default: return -1;
  }
#undef CHB_GB
  CHB_LAUNCH_CHECK();
  return 0;
}
}  // namespace

int launch_gather_push_binned(cudaStream_t st, int env, const double* x, const double* w, const cd* Fld, double* mom,
                              i64 cap, const GridGeom& g, double dt, const DeviceSet& und, const SortedSpec& sp) {
  if (sp.ncta <= 0) return 0;
  if (env && (g.nm % 2) != 1) { set_error("envelope gather needs an odd number of mode slots"); return 2; }
  return env ? launch_gather_binned_nm<1, 0>(st, x, w, Fld, mom, cap, g, 0.5 * dt, und, sp)
             : launch_gather_binned_nm<0, 0>(st, x, w, Fld, mom, cap, g, 0.5 * dt, und, sp);
}

// proj_fld alone on binned particles: Fld_tot(6, np) += gathered field (-1: no instantiation for this mode count)
int launch_gather_binned_out(cudaStream_t st, int env, const double* x, const double* w, const cd* Fld, double* fld_tot,
                             i64 cap, const GridGeom& g, const SortedSpec& sp) {
  if (sp.ncta <= 0) return 0;
  if (env && (g.nm % 2) != 1) { set_error("envelope gather needs an odd number of mode slots"); return 2; }
  DeviceSet none;
  memset(&none, 0, sizeof(none));
  return env ? launch_gather_binned_nm<1, 1>(st, x, w, Fld, fld_tot, cap, g, 0.0, none, sp)
             : launch_gather_binned_nm<0, 1>(st, x, w, Fld, fld_tot, cap, g, 0.0, none, sp);
}

int launch_deposit_binned(cudaStream_t st, int env, int curr, const double* x, const double* mom, const double* w,
                          i64 cap, cd* grid, const GridGeom& g, const ChunkSpec& ch, const SortedSpec& sp) {
  if (sp.ncta <= 0) return 0;
  if (env && (g.nm % 2) != 1) { set_error("envelope deposit needs an odd number of mode slots"); return 2; }
  int rc;
  if (env && curr) rc = launch_binned_nm<1, 1>(st, x, mom, w, cap, grid, g, ch, sp);
  else if (env)    rc = launch_binned_nm<1, 0>(st, x, nullptr, w, cap, grid, g, ch, sp);
  else if (curr)   rc = launch_binned_nm<0, 1>(st, x, mom, w, cap, grid, g, ch, sp);
  else             rc = launch_binned_nm<0, 0>(st, x, nullptr, w, cap, grid, g, ch, sp);
  return rc;  // -1: mode count not instantiated, caller falls back
}

}  // namespace chb
