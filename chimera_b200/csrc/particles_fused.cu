// particles_fused.cu -- the whole particle half of a PIC step in ONE kernel.
//
// Between two field solves the reference does, per particle and with no dependence on other particles,
//   proj_fld (gather E,B at x_n)           grid_deps.f90:149 / grid_deps_env.f90:164   [end of make_step k]
//   device field (undul_analytic)           devices.f90:162
//   push_velocs (Boris)                     particle_tools.f90:18
//   push_coords (x_n -> x_n+1, x_n+1/2)     particle_tools.f90:58                        [start of make_step k+1]
//   dep_curr at x_n+1/2, dep_dens at x_n+1  grid_deps*.f90:18,89 (+ _chnk / _env variants)
// (moduls/chimera_main.py:82-92: the tail of one make_step followed by the head of the next).  Run as four
// kernels this moves 408 B per particle through HBM and bins the particles three times; fused it reads
// x, p, w once and writes x, x_half, p once (128 B per particle) and bins once:
//
//  (P) a CTA takes FNPB consecutive particles of the cell-sorted order (one x-chunk); one thread per particle:
//      gather straight from the grid (the lanes of a warp sit in the same one or two cells, so their 4 x NM x 6 node
//      loads fall into the same few L1 sectors), device field, Boris push, strict-IEEE position update, global
//      stores, then the deposit records at x_half (J) and x_new (rho) in shared memory, keyed by the J cell inside a
//      FBX x FBR cell window, and the cell histogram.  A particle whose rho cell is its J cell (the common case:
//      |v| dt/2 << dx) is "fast" for rho too; the others (and J cells outside the window) are queued;
//  (B) block scan -> segments of <= FRUN same-cell particles; (C) counting sort of the local ids;
//  (F) deposit: one thread per (segment, J component | rho) accumulates the 4 x NM node values of the
//      segment's fast particles in registers and issues one red.global.add.f64 per node value;
//  (G) the queued ones: one thread per (particle, component, node) -> red.global.add.f64.
// MODE 0: everything (inside multi-step calls).  MODE 2: deposit only, from the stored x_half / x / p -- the half
// after a re-binning step's sort.  The half before it (gather + push + position update, no deposit) is the plain
// streaming kernel gather_push_coords_k below.
// Arithmetic per particle is the same as in the separate kernels (same helpers), so the result differs
// from them only in the order of the floating-point sums.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cub/block/block_scan.cuh>
#include "common.cuh"
#include "kernels.cuh"
#include "particle_dev.cuh"

namespace chb {
namespace {
#ifndef CHB_FT
#define CHB_FT 256
#endif
#ifndef CHB_FRUN
#define CHB_FRUN 48
#endif
#ifndef CHB_FDSPLIT
#define CHB_FDSPLIT 2
#endif
#ifndef CHB_FSPLIT
#define CHB_FSPLIT 2
#endif
#ifndef CHB_FSYNC
#define CHB_FSYNC 0
#endif
#ifndef CHB_FMINB
#define CHB_FMINB 3
#endif
#ifndef CHB_FPREFETCH
#define CHB_FPREFETCH 0
#endif
#ifndef CHB_GPC_MINB
#define CHB_GPC_MINB 3
#endif
constexpr int FNPB = kFusedNPB, FT = CHB_FT, FRUN = CHB_FRUN;
constexpr int FDSPLIT = CHB_FDSPLIT;
constexpr int FSPLIT = CHB_FSPLIT;  // lanes that share a (segment, unit) in the deposit stage: 2 or 4
#ifndef CHB_FBX
#define CHB_FBX 40
#endif
#ifndef CHB_FBR
#define CHB_FBR 16
#endif
constexpr int FBX = CHB_FBX, FBR = CHB_FBR, FBINS = FBX * FBR;
constexpr int FPPT = (FNPB + FT - 1) / FT;
constexpr int FMAXTASK = FNPB / FRUN + (FBINS < FNPB ? FBINS : FNPB);
constexpr int FITEMS = (FBINS + FT - 1) / FT;
constexpr int FSTR = FNPB + 1;  // plane stride of the record area: the 6 component threads of a segment write the same
                                // particle slot of 6 planes at once -> odd stride keeps them in different banks
constexpr int FNF = 12;  // doubles per particle in the record area (gather: 4|6 + 6 field values; deposit: 7|6 + 5|6)

// a cell's particles are cut into runs of FRUN; a tail of up to FSLACK particles joins the last run instead of
// becoming a segment of its own (a segment costs 4 x NM node loads / red.global.adds whatever its length)
constexpr int FSLACK = FRUN / 3;
__device__ __forceinline__ int f_nseg(int cnt) {
  return cnt <= 0 ? 0 : 1 + (cnt > FRUN + FSLACK ? (cnt - FSLACK - 1) / FRUN : 0);
}

// per-stage clock profile (chimera_fused_profile): cycles summed over CTAs, read after each barrier by thread 0
__constant__ int c_fprof_on = 0;
__device__ unsigned long long g_fprof[8];
#define FPROF_MARK(slot)                                                   \
  if (c_fprof_on) {                                                        \
    __syncthreads();                                                       \
    if (tid == 0) {                                                        \
      const long long now = clock64();                                     \
      atomicAdd(&g_fprof[slot], (unsigned long long)(now - fprof_t));      \
      fprof_t = now;                                                       \
    }                                                                      \
  }

struct FShared {
  int anchor[2];
  int total, ntask;
  int nslow[2];
  typename cub::BlockScan<int, FT>::TempStorage scan;
};

struct FRange {
  int chunk;
  i64 first;
  int count;
};

__device__ __forceinline__ FRange f_range(const SortedSpec& sp, int block) {
  int lo = 0, hi = sp.nchnk;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (block >= __ldg(sp.cta + mid)) lo = mid; else hi = mid;
  }
  FRange r;
  r.chunk = lo;
  r.first = (i64)__ldg(sp.ind + lo) + (i64)(block - __ldg(sp.cta + lo)) * FNPB;
  const i64 n = (i64)__ldg(sp.ind + lo + 1) - r.first;
  r.count = (int)(n < FNPB ? n : FNPB);
  return r;
}

__device__ __forceinline__ void f_keep_range(const ChunkSpec& ch, int c, i64 nxn, i64& lo, i64& hi) {
  lo = 0;
  hi = nxn - 1;
  if (ch.on) {  // the chunk-edge rule of grid_deps_chnk.f90:95-115 as a node interval (== chunk_keep per node)
    const i64 left = (i64)c * ch.cs;
    const i64 l2 = (left - ch.guards >= 0) ? left - ch.guards : left + 1;
    const i64 h2 = (left + ch.cs + ch.guards <= nxn - 1) ? left + ch.cs + ch.guards : left + ch.cs - 1;
    lo = l2 > lo ? l2 : lo;
    hi = h2 < hi ? h2 : hi;
  }
}

// Per-particle gather straight from the grid (grid_deps.f90:149-217, grid_deps_env.f90:164-238): the node weights
// are multiplied into the mode phase once per mode, then every (component, mode) costs 4 complex loads and 8 FMAs
// (4 for mode 0 of the real solver, whose phase is 1 and of which only the real part enters).
template <int ENV, int NM>
__device__ __forceinline__ void gather_direct(const GridGeom& g, const cd* __restrict__ Fld, const Shape& s, double xp,
                                              double yp, double zp, double F[6]) {
  constexpr int NKO = ENV ? (NM - 1) / 2 : NM - 1;
  const double rinv = (s.rp > 0.0) ? rsqrt(s.rp * s.rp) : 0.0;
  // Q4: phase at r = 0 is 0 (real solver) or 1 (envelope solver)
  const cd ph1 = (s.rp > 0.0) ? cmake(yp * rinv, zp * rinv) : (ENV ? cmake(1.0, 0.0) : cmake(0.0, 0.0));
  cd car = cmake(1.0, 0.0);
  if (ENV) {
    double sn, cs;
    sincos(xp * g.kx0, &sn, &cs);
    car = cmake(cs, sn);
  }
  const double w00 = s.sx0 * s.sr0, w10 = s.sx1 * s.sr0, w01 = s.sx0 * s.sr1, w11 = s.sx1 * s.sr1;
  const i64 plane = g.nxn * g.nrn;
  const cd* base = Fld + s.ix + g.nxn * s.ir;
  cd ph = cmake(1.0, 0.0);
#pragma unroll
  for (int iO = 0; iO <= NKO; ++iO) {
    if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
    for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
      const int slot = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
      if (!ENV && iO == 0) {
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const cd* pl = base + plane * (slot + NM * l);
          const cd f00 = __ldg(pl), f10 = __ldg(pl + 1), f01 = __ldg(pl + g.nxn), f11 = __ldg(pl + g.nxn + 1);
          F[l] += fma(w11, f11.x, fma(w01, f01.x, fma(w10, f10.x, w00 * f00.x)));
        }
        continue;
      }
      const cd pm = ENV ? cmul(car, sgn ? cconj(ph) : ph) : ph;
      const cd c00 = cscale(w00, pm), c10 = cscale(w10, pm), c01 = cscale(w01, pm), c11 = cscale(w11, pm);
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        const cd* pl = base + plane * (slot + NM * l);
        const cd f00 = __ldg(pl), f10 = __ldg(pl + 1), f01 = __ldg(pl + g.nxn), f11 = __ldg(pl + g.nxn + 1);
        double acc = c00.x * f00.x;
        acc = fma(-c00.y, f00.y, acc);
        acc = fma(c10.x, f10.x, acc);
        acc = fma(-c10.y, f10.y, acc);
        acc = fma(c01.x, f01.x, acc);
        acc = fma(-c01.y, f01.y, acc);
        acc = fma(c11.x, f11.x, acc);
        acc = fma(-c11.y, f11.y, acc);
        F[l] += acc;
      }
    }
  }
}

// gather + device field + Boris push of one particle (the tail of make_step k): momenta updated in registers
template <int ENV, int NM>
__device__ __forceinline__ void gather_push_one(const GridGeom& g, const cd* __restrict__ Fld, const DeviceSet& und,
                                                double xp, double yp, double zp, double wp, double dt_2, double& px,
                                                double& py, double& pz) {
  double Fp[6] = {0, 0, 0, 0, 0, 0};
  Shape s;
  // proj_fld skips w = 0 and r >= rmax (grid_deps.f90:171-176); outside the x range there is no cell to read
  if (wp != 0.0 && make_shape(g, xp, yp, zp, s) && s.ix >= 0 && s.ix <= g.nxn - 2) gather_direct<ENV, NM>(g, Fld, s, xp, yp, zp, Fp);
  if (und.n) apply_devices(und, xp, yp, zp, Fp);
  boris(px, py, pz, Fp[0], Fp[1], Fp[2], Fp[3], Fp[4], Fp[5], dt_2);
}

// push_coords (particle_tools.f90:58-82) in strict IEEE arithmetic, see push_coords_k; returns dt / gamma
__device__ __forceinline__ double coords_one(const double x0[3], const double pp[3], double dt, double x1[3], double xc[3]) {
  const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(pp[0], pp[0]), __dmul_rn(pp[1], pp[1])), __dmul_rn(pp[2], pp[2]));
  const double dt_gp = __ddiv_rn(dt, __dsqrt_rn(__dadd_rn(1.0, p2)));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    x1[c] = __dadd_rn(x0[c], __dmul_rn(pp[c], dt_gp));
    xc[c] = __dmul_rn(0.5, __dadd_rn(x0[c], x1[c]));
  }
  return dt_gp;
}

// The half of a re-binning step that comes BEFORE the sort (chimera_main.py:82-86): gather + push of step k and
// push_coords of step k+1 in one streaming pass, no deposit (that follows the sort, MODE 2 of the kernel below).
// COORDS = 0: gather + push only (make_halfstep, and the tail of a step() call).
template <int ENV, int NM, int COORDS>
__global__ void __launch_bounds__(256, CHB_GPC_MINB)
gather_push_coords_k(double* __restrict__ x, double* __restrict__ xh, double* __restrict__ mom, const double* __restrict__ w,
                     i64 cap, const cd* __restrict__ Fld, GridGeom g, double dt_2, double dt, DeviceSet und, i64 np) {
  for (i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += (i64)gridDim.x * blockDim.x) {
    const double x0[3] = {__ldg(x + ip), __ldg(x + cap + ip), __ldg(x + 2 * cap + ip)};
    double px = mom[ip], py = mom[cap + ip], pz = mom[2 * cap + ip];
    gather_push_one<ENV, NM>(g, Fld, und, x0[0], x0[1], x0[2], __ldg(w + ip), dt_2, px, py, pz);
    mom[ip] = px; mom[cap + ip] = py; mom[2 * cap + ip] = pz;
    if (!COORDS) continue;
    const double pp[3] = {px, py, pz};
    double x1[3], xc[3];
    coords_one(x0, pp, dt, x1, xc);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      x[c * cap + ip] = x1[c];
      xh[c * cap + ip] = xc[c];
    }
  }
}

template <int ENV, int NM, int SC, int MODE>
__global__ void __launch_bounds__(FT, CHB_FMINB)
fused_pass_k(double* __restrict__ x, double* __restrict__ xh, double* __restrict__ mom, const double* __restrict__ w,
                  i64 cap, const cd* __restrict__ Fld, cd* __restrict__ J, cd* __restrict__ Rho, GridGeom g, ChunkSpec ch,
                  double dt_2, double dt, DeviceSet und, SortedSpec sp, double leftX_J, double leftX_R) {
  // leftX_J / leftX_R: node 0 of the grid the current / the charge is deposited on.  They equal g.leftX unless a
  // window moves every step (chimera_main.py:286-302: stage 1 before push_coords, stage 2 between dep_curr and
  // dep_dens for a 'Staged' frame); the gather always uses g.leftX, the window position of the closing step.
  constexpr int NKO = ENV ? (NM - 1) / 2 : NM - 1;
  constexpr int NCJ = ENV ? 1 : 3;  // Q1: the envelope current has l = 3 only
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* rec = reinterpret_cast<double*>(smem_raw);             // [FNF][FNPB], by local particle id
  int* bins = reinterpret_cast<int*>(rec + FNF * FSTR);          // [FBINS]
  int* tasks = bins + FBINS;                                     // [FMAXTASK]  key | start << 11 | n << 22
  unsigned short* skey = reinterpret_cast<unsigned short*>(tasks + FMAXTASK);  // [FNPB]
  unsigned short* order = skey + FNPB;                           // [FNPB]
  unsigned short* slowJ = order + FNPB;                          // [FNPB] local ids whose J cell is outside the window
  unsigned short* slowR = slowJ + FNPB;                          // [FNPB] local ids whose rho cell is not their J cell
  unsigned* cellJ = reinterpret_cast<unsigned*>(slowR + FNPB);   // [FNPB] (ix + 1) | ir << 20 of the J deposit cell
  unsigned* cellR = cellJ + FNPB;                                // [FNPB]
  unsigned char* fast = reinterpret_cast<unsigned char*>(cellR + FNPB);        // [FNPB] bit0: J, bit1: rho
  __shared__ FShared sh;
  const int tid = threadIdx.x;
  long long fprof_t = c_fprof_on ? clock64() : 0;

  const FRange cr = f_range(sp, (int)blockIdx.x + sp.cta_base);
  if (cr.count <= 0) return;
  // ---- (P) one thread per particle.  The window is anchored on the first particle of the block.
  for (int i = tid; i < FBINS; i += FT) bins[i] = 0;
  if (tid == 0) {
    const double* xa = MODE == 2 ? xh : x;  // the position the particles are sorted by
    const double xp = __ldg(xa + cr.first), yp = __ldg(xa + cap + cr.first), zp = __ldg(xa + 2 * cap + cr.first);
    const double lx0 = MODE == 2 ? leftX_J : g.leftX;
    const i64 ix = (i64)floor((xp - lx0) * g.dx_inv);
    const i64 ir = (i64)floor((sqrt(yp * yp + zp * zp) - g.r0) * g.dr_inv);
    i64 ax = ix - FBX / 2;
    if (sp.tile_w > 0 && sp.tile_w <= FBX) {
      i64 lx = ix - (i64)cr.chunk * sp.cs;
      lx = lx < 0 ? 0 : (lx > sp.cs - 1 ? sp.cs - 1 : lx);
      ax = (i64)cr.chunk * sp.cs + (lx / sp.tile_w) * sp.tile_w - (FBX - sp.tile_w) / 2;
    }
    sh.anchor[0] = (int)ax;
    sh.anchor[1] = (int)(ir - 3);
    sh.nslow[0] = 0;
    sh.nslow[1] = 0;
  }
  __syncthreads();
  const int ix0 = sh.anchor[0], ir0 = sh.anchor[1];
  const double dt_inv = 1.0 / dt;
  double* recJ = rec;                         // [7 | 6][FNPB]: fx, fr, ph.x, ph.y, amplitude(s)
  double* recR = rec + (ENV ? 6 : 7) * FSTR;  // [5 | 6][FNPB]
#pragma unroll 1
  for (int j = 0; j < FPPT; ++j) {
    const int li = tid + j * FT;
    if (li >= FNPB) continue;
    if (li >= cr.count) {
      skey[li] = 0xFFFFu;
      fast[li] = 0;
      continue;
    }
    const i64 ip = cr.first + li;
    const double wp = __ldg(w + ip);
    double px = mom[ip], py = mom[cap + ip], pz = mom[2 * cap + ip];
    double x1[3], xc[3], dt_gp;
    if (MODE == 2) {  // positions and momenta are final: the deposits of the step that was just re-binned
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        x1[c] = __ldg(x + c * cap + ip);
        xc[c] = __ldg(xh + c * cap + ip);
      }
      dt_gp = dt * rsqrt(1.0 + (px * px + py * py + pz * pz));
    } else {
      const double x0[3] = {__ldg(x + ip), __ldg(x + cap + ip), __ldg(x + 2 * cap + ip)};
      gather_push_one<ENV, NM>(g, Fld, und, x0[0], x0[1], x0[2], wp, dt_2, px, py, pz);
      mom[ip] = px; mom[cap + ip] = py; mom[2 * cap + ip] = pz;
      const double pp[3] = {px, py, pz};
      dt_gp = coords_one(x0, pp, dt, x1, xc);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        x[c * cap + ip] = x1[c];
        xh[c * cap + ip] = xc[c];
      }
    }
    unsigned short key = 0xFFFFu;
    unsigned char fl = 0;
    // Deposit records.  The J cell inside the window is the binning key; a charge deposit into the same cell goes
    // the register path of stage (F) with it, anything else is queued for stage (G).
    {  // current at the centred position
      Shape s;
      bool direct = false;
      if (wp != 0.0 && make_shape_at(g, leftX_J, xc[0], xc[1], xc[2], s) && fabs(px) + fabs(py) + fabs(pz) != 0.0 &&
          s.ix >= -1 && s.ix <= g.nxn - 1) {  // dep_curr skips w = 0, r >= rmax and particles at rest (grid_deps.f90:36-42)
        if (s.ix + 1 < (1 << 20) && s.ir < (1 << 12)) {
          recJ[li] = s.sx1;
          recJ[FSTR + li] = s.sr1;
          const double rinv = (s.rp > 0.0) ? rsqrt(s.rp * s.rp) : 0.0;  // deposit phase e^{-i theta}; 0 on the axis
          recJ[2 * FSTR + li] = xc[1] * rinv;
          recJ[3 * FSTR + li] = -xc[2] * rinv;
          const double ginv = dt_gp * dt_inv;  // 1 / gamma, from the position update's dt / gamma
          if (ENV) {
            double sn, cs;
            sincos(xc[0] * g.kx0, &sn, &cs);
            const cd base = cscale(pz * ginv, cmake(wp * cs, -wp * sn));
            recJ[4 * FSTR + li] = base.x;
            recJ[5 * FSTR + li] = base.y;
          } else {
            const double wg = wp * ginv;
            recJ[4 * FSTR + li] = px * wg;
            recJ[5 * FSTR + li] = py * wg;
            recJ[6 * FSTR + li] = pz * wg;
          }
          const i64 kx = s.ix - ix0, kr = s.ir - ir0;
          if (kx >= 0 && kx < FBX && kr >= 0 && kr < FBR) {
            key = (unsigned short)(kr * FBX + kx);
            fl |= 1;
          } else {
            cellJ[li] = (unsigned)(s.ix + 1) | ((unsigned)s.ir << 20);
            slowJ[atomicAdd(&sh.nslow[0], 1)] = (unsigned short)li;
          }
        } else direct = true;
      }
      if (direct) {
        GridGeom gj = g;
        gj.leftX = leftX_J;
        deposit_one<ENV, 1>(gj, ch, cr.chunk, J, xc[0], xc[1], xc[2], px, py, pz, wp);
      }
    }
    if (SC) {  // charge at the new position
      Shape s;
      bool direct = false;
      if (wp != 0.0 && make_shape_at(g, leftX_R, x1[0], x1[1], x1[2], s) && s.ix >= -1 && s.ix <= g.nxn - 1) {
        if (s.ix + 1 < (1 << 20) && s.ir < (1 << 12)) {
          recR[li] = s.sx1;
          recR[FSTR + li] = s.sr1;
          const double rinv = (s.rp > 0.0) ? rsqrt(s.rp * s.rp) : 0.0;
          recR[2 * FSTR + li] = x1[1] * rinv;
          recR[3 * FSTR + li] = -x1[2] * rinv;
          if (ENV) {
            double sn, cs;
            sincos(x1[0] * g.kx0, &sn, &cs);
            const cd wpc = cmake(wp * cs, -wp * sn);
            const cd base = cmul(wpc, wpc);  // Q2: the complex weight enters twice
            recR[4 * FSTR + li] = base.x;
            recR[5 * FSTR + li] = base.y;
          } else {
            recR[4 * FSTR + li] = wp;
          }
          const i64 kx = s.ix - ix0, kr = s.ir - ir0;
          if (key != 0xFFFFu && kx >= 0 && kx < FBX && kr >= 0 && kr < FBR && (int)(kr * FBX + kx) == (int)key) fl |= 2;
          else {
            cellR[li] = (unsigned)(s.ix + 1) | ((unsigned)s.ir << 20);
            slowR[atomicAdd(&sh.nslow[1], 1)] = (unsigned short)li;
          }
        } else direct = true;
      }
      if (direct) {
        GridGeom gr = g;
        gr.leftX = leftX_R;
        deposit_one<ENV, 0>(gr, ch, cr.chunk, Rho, x1[0], x1[1], x1[2], 0.0, 0.0, 0.0, wp);
      }
    }
    if (key != 0xFFFFu) atomicAdd(&bins[key], 1);
    skey[li] = key;
    fast[li] = fl;
  }
  __syncthreads();
  FPROF_MARK(0)
  // ---- (B) packed scan (low 16 bits: particles, high 16: segments) and the segment table
  {
    int items[FITEMS], cnt[FITEMS];
    int total = 0;
#pragma unroll
    for (int i = 0; i < FITEMS; ++i) {
      const int b = tid * FITEMS + i;
      cnt[i] = (b < FBINS) ? bins[b] : 0;
      items[i] = cnt[i] | (f_nseg(cnt[i]) << 16);
    }
    cub::BlockScan<int, FT>(sh.scan).ExclusiveSum(items, items, total);
#pragma unroll
    for (int i = 0; i < FITEMS; ++i) {
      const int b = tid * FITEMS + i;
      if (b < FBINS) {
        bins[b] = items[i] & 0xFFFF;
        int start = items[i] & 0xFFFF, t = items[i] >> 16, left = cnt[i];
        while (left > 0) {
          const int n = left <= FRUN + FSLACK ? left : FRUN;  // the tail joins the last full run
          tasks[t++] = b | (start << 11) | (n << 22);
          start += n;
          left -= n;
        }
      }
    }
    if (tid == 0) { sh.total = total & 0xFFFF; sh.ntask = total >> 16; }
  }
  __syncthreads();
  // ---- (C) local ids to sorted slots
#pragma unroll
  for (int j = 0; j < FPPT; ++j) {
    const int li = tid + j * FT;
    if (li < FNPB) {
      const int key = skey[li];
      if (key != 0xFFFF) order[atomicAdd(&bins[key], 1)] = (unsigned short)li;
    }
  }
  __syncthreads();
  FPROF_MARK(1)
  const i64 plane = g.nxn * g.nrn;
  const int ntask = sh.ntask;
  FPROF_MARK(2)
  FPROF_MARK(3)

  // ---- (F) deposit: one thread per (segment, unit), unit = J component(s) then rho
  constexpr int NU = NCJ + (SC ? 1 : 0);
  i64 klo, khi;
  f_keep_range(ch, cr.chunk, g.nxn, klo, khi);
  // Each (segment, unit) is shared by a lane pair: lane h takes the particles q = h, h + 2, ... of the run; the
  // halves are exchanged with one shuffle per accumulator (lane 0 ends up with the sums of the left x node pair,
  // lane 1 with the right pair) and each lane issues its half of the red.global.adds.
  const int nwork = ntask * NU * FSPLIT;
#pragma unroll 1
  for (int t = tid; t < ((nwork + 31) & ~31); t += FT) {
    const bool valid = t < nwork;
    const int hs = t % FSPLIT, h = hs & 1, tu = valid ? (t / FSPLIT) : 0;
    const int task = tu / NU, u = tu - task * NU;
    const bool isJ = u < NCJ;
    const int tw = tasks[task];
    const int key = tw & 0x7FF, start = (tw >> 11) & 0x7FF, n = valid ? (tw >> 22) : 0;
    const double* rb = isJ ? recJ : recR;
    const double* ra = rb + (4 + ((!ENV && isJ) ? u : 0)) * FSTR;
    const unsigned char bit = isJ ? 1 : 2;
    cd a[2][2][NM];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int m = 0; m < NM; ++m) a[i][k][m] = cmake(0.0, 0.0);
    int any = 0;
#pragma unroll 2
    for (int q = hs; q < n; q += FSPLIT) {
      const int li = order[start + q];
      if (!(fast[li] & bit)) continue;
      any = 1;
      const double fx = rb[li], fr = rb[FSTR + li];
      const cd ph1 = cmake(rb[2 * FSTR + li], rb[3 * FSTR + li]);
      const cd amp = ENV ? cmake(ra[li], ra[FSTR + li]) : cmake(ra[li], 0.0);
      const double w00 = (1.0 - fx) * (1.0 - fr), w01 = (1.0 - fx) * fr, w10 = fx * (1.0 - fr), w11 = fx * fr;
      cd ph = cmake(1.0, 0.0);
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
        for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
          const int m = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
          const cd pm = sgn ? cconj(ph) : ph;
          if (!ENV && iO == 0) {  // e^{-i 0 theta} = 1: real amplitude, the imaginary parts stay 0
            a[0][0][m].x += w00 * amp.x; a[0][1][m].x += w01 * amp.x;
            a[1][0][m].x += w10 * amp.x; a[1][1][m].x += w11 * amp.x;
            continue;
          }
          const cd f = ENV ? cmul(amp, pm) : cscale(amp.x, pm);
          a[0][0][m].x += w00 * f.x; a[0][0][m].y += w00 * f.y;
          a[0][1][m].x += w01 * f.x; a[0][1][m].y += w01 * f.y;
          a[1][0][m].x += w10 * f.x; a[1][0][m].y += w10 * f.y;
          a[1][1][m].x += w11 * f.x; a[1][1][m].y += w11 * f.y;
        }
      }
    }
    if (FSPLIT == 4) {  // lanes hs and hs ^ 2 first pool their partial sums
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            a[i][k][m].x += __shfl_xor_sync(0xffffffffu, a[i][k][m].x, 2);
            if (ENV || m > 0) a[i][k][m].y += __shfl_xor_sync(0xffffffffu, a[i][k][m].y, 2);
          }
      any |= __shfl_xor_sync(0xffffffffu, any, 2);
    }
    // lane h keeps x node h: send the other node's partial sums to the partner, add what it sends
    cd mine[2][NM];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const cd keep = h ? a[1][k][m] : a[0][k][m], give = h ? a[0][k][m] : a[1][k][m];
        mine[k][m].x = keep.x + __shfl_xor_sync(0xffffffffu, give.x, 1);
        mine[k][m].y = (!ENV && m == 0) ? 0.0 : keep.y + __shfl_xor_sync(0xffffffffu, give.y, 1);
      }
    any |= __shfl_xor_sync(0xffffffffu, any, 1);
    if (!valid || !any || hs >= 2) continue;
    const int kr = key / FBX, kx = key - kr * FBX;
    const i64 gx = (i64)ix0 + kx + h, gr = (i64)ir0 + kr;
    if (gx < klo || gx > khi) continue;
    const int l = isJ ? (ENV ? 2 : u) : 0;
    cd* pl = (isJ ? J : Rho) + plane * (g.nm * l) + gx + g.nxn * gr;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      if (!ENV && m == 0) {  // imaginary sums are exactly 0
        atomicAdd(&pl[0].x, mine[0][0].x);
        atomicAdd(&pl[g.nxn].x, mine[1][0].x);
        continue;
      }
      red_add(pl + plane * m, mine[0][m]);
      red_add(pl + plane * m + g.nxn, mine[1][m]);
    }
  }

#if CHB_FSYNC & 1
  __syncthreads();
#endif
  FPROF_MARK(4)
  // ---- (G) particles that changed cell: one thread per (particle, component, node), so that the few of
  // them cost a few warp-wide red.global.add instead of serialising inside divergent warps
#pragma unroll 1
  for (int pass = 0; pass < (SC ? 2 : 1); ++pass) {
    const bool isJ = pass == 0;
    const int per = (isJ ? NCJ : 1) * 4;
    const int nitems = sh.nslow[pass] * per;
    const double* rb = isJ ? recJ : recR;
    const unsigned short* lst = isJ ? slowJ : slowR;
    const unsigned* cells = isJ ? cellJ : cellR;
    cd* grid = isJ ? J : Rho;
#pragma unroll 1
    for (int t = tid; t < nitems; t += FT) {
      const int q = t / per, r = t - q * per;
      const int u = r >> 2, node = r & 3;
      const int i = node >> 1, k = node & 1;
      const int li = lst[q];
      const unsigned cell = cells[li];
      const i64 gx = (i64)(cell & 0xFFFFFu) - 1 + i, gr = (i64)(cell >> 20) + k;
      const bool keep = gx >= 0 && gx <= g.nxn - 1 && (!ch.on || chunk_keep(ch, cr.chunk, gx, g.nxn));
      if (!keep) continue;
      const double fx = rb[li], fr = rb[FSTR + li];
      const cd ph1 = cmake(rb[2 * FSTR + li], rb[3 * FSTR + li]);
      const double* ra = rb + (4 + ((!ENV && isJ) ? u : 0)) * FSTR;
      const cd amp = ENV ? cmake(ra[li], ra[FSTR + li]) : cmake(ra[li], 0.0);
      const double wgt = (i ? fx : 1.0 - fx) * (k ? fr : 1.0 - fr);
      const int l = isJ ? (ENV ? 2 : u) : 0;
      cd* pl = grid + plane * (g.nm * l) + gx + g.nxn * gr;
      cd ph = cmake(1.0, 0.0);
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
        for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
          const int m = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
          const cd pm = sgn ? cconj(ph) : ph;
          const cd f = ENV ? cmul(amp, pm) : cscale(amp.x, pm);
          red_add(pl + plane * m, cscale(wgt, f));
        }
      }
    }
  }
#if CHB_FSYNC & 2
  __syncthreads();
#endif
  FPROF_MARK(5)
  if (c_fprof_on && tid == 0) atomicAdd(&g_fprof[7], 1ull);
}

// ---- the step kernel of multi-step calls: per-CTA binning on the GATHER cell, gather with the node values of a run
// in registers (one thread per (run, field component)), push, position update, deposit records, deposit.  Measured
// against the one-thread-per-particle pass above (fused_pass_k<.., 0>, with and without a shared-memory field tile):
// 10.4 ms vs 11.0 - 12.6 ms per launch at 9.85e7 particles (profiles/r02_fused_variants.md), so this one stays.
template <int ENV, int NM, int SC>
__global__ void __launch_bounds__(FT, CHB_FMINB)
fused_particles_k(double* __restrict__ x, double* __restrict__ xh, double* __restrict__ mom, const double* __restrict__ w,
                  i64 cap, const cd* __restrict__ Fld, cd* __restrict__ J, cd* __restrict__ Rho, GridGeom g, ChunkSpec ch,
                  double dt_2, double dt, DeviceSet und, SortedSpec sp, double leftX_J, double leftX_R) {
  // leftX_J / leftX_R: node 0 of the grid the current / the charge is deposited on.  They equal g.leftX unless a
  // window moves every step (chimera_main.py:286-302: stage 1 before push_coords, stage 2 between dep_curr and
  // dep_dens for a 'Staged' frame); the gather always uses g.leftX, the window position of the closing step.
  constexpr int NKO = ENV ? (NM - 1) / 2 : NM - 1;
  constexpr int NCJ = ENV ? 1 : 3;  // Q1: the envelope current has l = 3 only
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* rec = reinterpret_cast<double*>(smem_raw);             // [FNF][FNPB], by local particle id
  int* bins = reinterpret_cast<int*>(rec + FNF * FSTR);          // [FBINS]
  int* tasks = bins + FBINS;                                     // [FMAXTASK]  key | start << 11 | n << 22
  unsigned short* skey = reinterpret_cast<unsigned short*>(tasks + FMAXTASK);  // [FNPB]
  unsigned short* order = skey + FNPB;                           // [FNPB]
  unsigned short* slowJ = order + FNPB;                          // [FNPB] local ids that changed cell (J)
  unsigned short* slowR = slowJ + FNPB;                          // [FNPB] ... (rho)
  unsigned* cellJ = reinterpret_cast<unsigned*>(slowR + FNPB);   // [FNPB] (ix + 1) | ir << 20 of the J deposit cell
  unsigned* cellR = cellJ + FNPB;                                // [FNPB]
  unsigned char* fast = reinterpret_cast<unsigned char*>(cellR + FNPB);        // [FNPB] bit0: J, bit1: rho
  __shared__ FShared sh;
  const int tid = threadIdx.x;
  long long fprof_t = c_fprof_on ? clock64() : 0;

  const FRange cr = f_range(sp, (int)blockIdx.x + sp.cta_base);
  if (cr.count <= 0) return;
  // ---- (A) loads (issued before the anchor barrier so that their latency overlaps it), gather records, histogram
  double xa[FPPT], ya[FPPT], za[FPPT], wa[FPPT];  // dead after stage (A)
#pragma unroll
  for (int j = 0; j < FPPT; ++j) {
    const int li = tid + j * FT;
    const bool in = li < cr.count;
    const i64 ip = cr.first + (in ? li : 0);
    xa[j] = __ldg(x + ip); ya[j] = __ldg(x + cap + ip); za[j] = __ldg(x + 2 * cap + ip);
    wa[j] = in ? __ldg(w + ip) : 0.0;
#if CHB_FPREFETCH
    // the momenta are first needed in stage (E): start their way from DRAM now
#if CHB_FPREFETCH == 2
#define CHB_PF "prefetch.global.L1 [%0];"
#else
#define CHB_PF "prefetch.global.L2 [%0];"
#endif
    asm volatile(CHB_PF ::"l"(mom + ip));
    asm volatile(CHB_PF ::"l"(mom + cap + ip));
    asm volatile(CHB_PF ::"l"(mom + 2 * cap + ip));
#endif
  }
  for (int i = tid; i < FBINS; i += FT) bins[i] = 0;
  if (tid == 0) {
    const double xp = xa[0], yp = ya[0], zp = za[0];  // li = 0: the first particle of the block
    const i64 ix = (i64)floor((xp - g.leftX) * g.dx_inv);
    const i64 ir = (i64)floor((sqrt(yp * yp + zp * zp) - g.r0) * g.dr_inv);
    i64 ax = ix - FBX / 2;
    if (sp.tile_w > 0 && sp.tile_w <= FBX) {
      i64 lx = ix - (i64)cr.chunk * sp.cs;
      lx = lx < 0 ? 0 : (lx > sp.cs - 1 ? sp.cs - 1 : lx);
      ax = (i64)cr.chunk * sp.cs + (lx / sp.tile_w) * sp.tile_w - (FBX - sp.tile_w) / 2;
    }
    sh.anchor[0] = (int)ax;
    sh.anchor[1] = (int)(ir - 3);
    sh.nslow[0] = 0;
    sh.nslow[1] = 0;
  }
  __syncthreads();
  const int ix0 = sh.anchor[0], ir0 = sh.anchor[1];
  double* fbuf = rec + 6 * FSTR;  // [6][FNPB] gathered field, phase (D)-(E)

  {
#pragma unroll
  for (int j = 0; j < FPPT; ++j) {
    const int li = tid + j * FT;
    if (li >= FNPB) continue;
    unsigned short key = 0xFFFFu;
    const double xp = xa[j], yp = ya[j], zp = za[j];
    double F[6] = {0, 0, 0, 0, 0, 0};
    Shape s;
    if (wa[j] != 0.0 && make_shape(g, xp, yp, zp, s) && s.ix >= 0 && s.ix <= g.nxn - 2) {
      const i64 kx = s.ix - ix0, kr = s.ir - ir0;
      if (kx >= 0 && kx < FBX && kr >= 0 && kr < FBR) {
        key = (unsigned short)(kr * FBX + kx);
        atomicAdd(&bins[key], 1);
        rec[li] = s.sx1;
        rec[FSTR + li] = s.sr1;
        const double rinv = (s.rp > 0.0) ? rsqrt(s.rp * s.rp) : 0.0;  // gather phase e^{+i theta}; axis: 0 | 1 (Q4)
        rec[2 * FSTR + li] = (s.rp > 0.0) ? yp * rinv : (ENV ? 1.0 : 0.0);
        rec[3 * FSTR + li] = zp * rinv;
        if (ENV) {
          double sn, cs;
          sincos(xp * g.kx0, &sn, &cs);
          rec[4 * FSTR + li] = cs;
          rec[5 * FSTR + li] = sn;
        }
      } else {
        gather_one<ENV>(g, Fld, xp, yp, zp, F);  // drifted out of the window: L2 path
      }
    }
    if (key == 0xFFFFu) {
#pragma unroll
      for (int l = 0; l < 6; ++l) fbuf[l * FSTR + li] = F[l];
    }
    skey[li] = key;
  }
  }
  __syncthreads();
  FPROF_MARK(0)
  // ---- (B) packed scan (low 16 bits: particles, high 16: segments) and the segment table
  {
    int items[FITEMS], cnt[FITEMS];
    int total = 0;
#pragma unroll
    for (int i = 0; i < FITEMS; ++i) {
      const int b = tid * FITEMS + i;
      cnt[i] = (b < FBINS) ? bins[b] : 0;
      items[i] = cnt[i] | (f_nseg(cnt[i]) << 16);
    }
    cub::BlockScan<int, FT>(sh.scan).ExclusiveSum(items, items, total);
#pragma unroll
    for (int i = 0; i < FITEMS; ++i) {
      const int b = tid * FITEMS + i;
      if (b < FBINS) {
        bins[b] = items[i] & 0xFFFF;
        int start = items[i] & 0xFFFF, t = items[i] >> 16, left = cnt[i];
        while (left > 0) {
          const int n = left <= FRUN + FSLACK ? left : FRUN;  // the tail joins the last full run
          tasks[t++] = b | (start << 11) | (n << 22);
          start += n;
          left -= n;
        }
      }
    }
    if (tid == 0) { sh.total = total & 0xFFFF; sh.ntask = total >> 16; }
  }
  __syncthreads();
  // ---- (C) local ids to sorted slots
#pragma unroll
  for (int j = 0; j < FPPT; ++j) {
    const int li = tid + j * FT;
    if (li < FNPB) {
      const int key = skey[li];
      if (key != 0xFFFF) order[atomicAdd(&bins[key], 1)] = (unsigned short)li;
    }
  }
  __syncthreads();
  FPROF_MARK(1)
  const i64 plane = g.nxn * g.nrn;
  const int ntask = sh.ntask;
  // ---- (D) gather: one thread per (segment, field component)
#pragma unroll 1
  for (int t = tid; t < ntask * 6 * FDSPLIT; t += FT) {
    const int hd = t % FDSPLIT, td = t / FDSPLIT;  // FDSPLIT lanes share a (segment, component): particle q = hd, hd + FDSPLIT, ..
    const int task = td / 6, l = td - task * 6;
    const int tw = tasks[task];
    const int key = tw & 0x7FF, start = (tw >> 11) & 0x7FF, n = tw >> 22;
    const int kr = key / FBX, kx = key - kr * FBX;
    const cd* pl = Fld + plane * g.nm * l + ((i64)ix0 + kx) + g.nxn * ((i64)ir0 + kr);
    // node values in difference form: value(fx, fr) = N0 + fx Nx + fr Nr + fx fr Nxr  (3 FMAs per interpolation)
    cd N[NM][4];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const cd n00 = __ldg(pl + plane * m), n10 = __ldg(pl + plane * m + 1);
      const cd n01 = __ldg(pl + plane * m + g.nxn), n11 = __ldg(pl + plane * m + g.nxn + 1);
      N[m][0] = n00;
      N[m][1] = csub(n10, n00);
      N[m][2] = csub(n01, n00);
      N[m][3] = csub(csub(n11, n01), N[m][1]);
    }
#pragma unroll 2
    for (int q = hd; q < n; q += FDSPLIT) {
      const int li = order[start + q];
      const double fx = rec[li], fr = rec[FSTR + li];
      const cd ph1 = cmake(rec[2 * FSTR + li], rec[3 * FSTR + li]);
      const double fxr = fx * fr;
      cd car = cmake(1.0, 0.0);
      if (ENV) car = cmake(rec[4 * FSTR + li], rec[5 * FSTR + li]);  // carrier e^{+i kx0 x}
      cd ph = cmake(1.0, 0.0);
      double Fv = 0.0;
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
        for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
          const int slot = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
          const double sx = fma(fxr, N[slot][3].x, fma(fr, N[slot][2].x, fma(fx, N[slot][1].x, N[slot][0].x)));
          if (!ENV && iO == 0) {  // mode 0 of the real solver: phase 1, only the real part enters
            Fv += sx;
          } else {
            const cd pm = ENV ? cmul(car, sgn ? cconj(ph) : ph) : ph;
            const double sy = fma(fxr, N[slot][3].y, fma(fr, N[slot][2].y, fma(fx, N[slot][1].y, N[slot][0].y)));
            Fv += pm.x * sx - pm.y * sy;
          }
        }
      }
      fbuf[l * FSTR + li] = Fv;
    }
  }
  __syncthreads();
  FPROF_MARK(2)

  // ---- (E) device field, Boris push, position update, deposit records.  The deposit records reuse the
  // whole record area, gathered field included: slot li of every plane belongs to the one thread that handles
  // particle li here, and it reads the particle's field values before it writes the particle's records.
  const double dt_inv = 1.0 / dt;
  double* recJ = rec;                         // [7 | 6][FNPB]: fx, fr, ph.x, ph.y, amplitude(s)
  double* recR = rec + (ENV ? 6 : 7) * FSTR;  // [5 | 6][FNPB]
  double xs[FPPT], ys[FPPT], zs[FPPT], ws[FPPT], pxs[FPPT], pys[FPPT], pzs[FPPT];
#pragma unroll
  for (int j = 0; j < FPPT; ++j) {  // x, w were read a moment ago by this CTA (L1/L2 hits); holding them in registers
    const int li = tid + j * FT;    // since stage (A) costs more in spills than the reload
    const bool in = li < cr.count;
    const i64 ip = cr.first + (in ? li : 0);
    xs[j] = __ldg(x + ip); ys[j] = __ldg(x + cap + ip); zs[j] = __ldg(x + 2 * cap + ip);
    ws[j] = in ? __ldg(w + ip) : 0.0;
    pxs[j] = mom[ip]; pys[j] = mom[cap + ip]; pzs[j] = mom[2 * cap + ip];
  }
#pragma unroll
  for (int j = 0; j < FPPT; ++j) {
    const int li = tid + j * FT;
    if (li >= cr.count) {
      if (li < FNPB) fast[li] = 0;
      continue;
    }
    const i64 ip = cr.first + li;
    double Fp[6];
#pragma unroll
    for (int l = 0; l < 6; ++l) Fp[l] = fbuf[l * FSTR + li];
    if (und.n) apply_devices(und, xs[j], ys[j], zs[j], Fp);
    double px = pxs[j], py = pys[j], pz = pzs[j];
    boris(px, py, pz, Fp[0], Fp[1], Fp[2], Fp[3], Fp[4], Fp[5], dt_2);
    mom[ip] = px; mom[cap + ip] = py; mom[2 * cap + ip] = pz;
    // push_coords (particle_tools.f90:58-82) in strict IEEE arithmetic, see push_coords_k
    const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz));
    const double dt_gp = __ddiv_rn(dt, __dsqrt_rn(__dadd_rn(1.0, p2)));
    const double x0[3] = {xs[j], ys[j], zs[j]}, pp[3] = {px, py, pz};
    double x1[3], xc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      x1[c] = __dadd_rn(x0[c], __dmul_rn(pp[c], dt_gp));
      xc[c] = __dmul_rn(0.5, __dadd_rn(x0[c], x1[c]));
      x[c * cap + ip] = x1[c];
      xh[c * cap + ip] = xc[c];
    }
    const int key = skey[li];
    const double wp = ws[j];
    unsigned char fl = 0;
    // Deposit records: a particle whose deposit cell is the cell it was binned under goes the register path
    // of stage (F); one that changed cell (or was outside the window) is queued for stage (G).
    {  // current at the centred position
      Shape s;
      bool direct = false;
      if (wp != 0.0 && make_shape_at(g, leftX_J, xc[0], xc[1], xc[2], s) && fabs(px) + fabs(py) + fabs(pz) != 0.0 &&
          s.ix >= -1 && s.ix <= g.nxn - 1) {  // dep_curr skips w = 0, r >= rmax and particles at rest (grid_deps.f90:36-42)
        if (s.ix + 1 < (1 << 20) && s.ir < (1 << 12)) {
          recJ[li] = s.sx1;
          recJ[FSTR + li] = s.sr1;
          const double rinv = (s.rp > 0.0) ? rsqrt(s.rp * s.rp) : 0.0;  // deposit phase e^{-i theta}; 0 on the axis
          recJ[2 * FSTR + li] = xc[1] * rinv;
          recJ[3 * FSTR + li] = -xc[2] * rinv;
          const double ginv = dt_gp * dt_inv;  // 1 / gamma, from the position update's dt / gamma
          if (ENV) {
            double sn, cs;
            sincos(xc[0] * g.kx0, &sn, &cs);
            const cd base = cscale(pz * ginv, cmake(wp * cs, -wp * sn));
            recJ[4 * FSTR + li] = base.x;
            recJ[5 * FSTR + li] = base.y;
          } else {
            const double wg = wp * ginv;
            recJ[4 * FSTR + li] = px * wg;
            recJ[5 * FSTR + li] = py * wg;
            recJ[6 * FSTR + li] = pz * wg;
          }
          const i64 kx = s.ix - ix0, kr = s.ir - ir0;
          if (key != 0xFFFF && kx >= 0 && kx < FBX && kr >= 0 && kr < FBR && (int)(kr * FBX + kx) == key) fl |= 1;
          else {
            cellJ[li] = (unsigned)(s.ix + 1) | ((unsigned)s.ir << 20);
            slowJ[atomicAdd(&sh.nslow[0], 1)] = (unsigned short)li;
          }
        } else direct = true;
      }
      if (direct) {
        GridGeom gj = g;
        gj.leftX = leftX_J;
        deposit_one<ENV, 1>(gj, ch, cr.chunk, J, xc[0], xc[1], xc[2], px, py, pz, wp);
      }
    }
    if (SC) {  // charge at the new position
      Shape s;
      bool direct = false;
      if (wp != 0.0 && make_shape_at(g, leftX_R, x1[0], x1[1], x1[2], s) && s.ix >= -1 && s.ix <= g.nxn - 1) {
        if (s.ix + 1 < (1 << 20) && s.ir < (1 << 12)) {
          recR[li] = s.sx1;
          recR[FSTR + li] = s.sr1;
          const double rinv = (s.rp > 0.0) ? rsqrt(s.rp * s.rp) : 0.0;
          recR[2 * FSTR + li] = x1[1] * rinv;
          recR[3 * FSTR + li] = -x1[2] * rinv;
          if (ENV) {
            double sn, cs;
            sincos(x1[0] * g.kx0, &sn, &cs);
            const cd wpc = cmake(wp * cs, -wp * sn);
            const cd base = cmul(wpc, wpc);  // Q2: the complex weight enters twice
            recR[4 * FSTR + li] = base.x;
            recR[5 * FSTR + li] = base.y;
          } else {
            recR[4 * FSTR + li] = wp;
          }
          const i64 kx = s.ix - ix0, kr = s.ir - ir0;
          if (key != 0xFFFF && kx >= 0 && kx < FBX && kr >= 0 && kr < FBR && (int)(kr * FBX + kx) == key) fl |= 2;
          else {
            cellR[li] = (unsigned)(s.ix + 1) | ((unsigned)s.ir << 20);
            slowR[atomicAdd(&sh.nslow[1], 1)] = (unsigned short)li;
          }
        } else direct = true;
      }
      if (direct) {
        GridGeom gr = g;
        gr.leftX = leftX_R;
        deposit_one<ENV, 0>(gr, ch, cr.chunk, Rho, x1[0], x1[1], x1[2], 0.0, 0.0, 0.0, wp);
      }
    }
    fast[li] = fl;
  }
  __syncthreads();
  FPROF_MARK(3)

  // ---- (F) deposit: one thread per (segment, unit), unit = J component(s) then rho
  constexpr int NU = NCJ + (SC ? 1 : 0);
  i64 klo, khi;
  f_keep_range(ch, cr.chunk, g.nxn, klo, khi);
  // Each (segment, unit) is shared by a lane pair: lane h takes the particles q = h, h + 2, ... of the run; the
  // halves are exchanged with one shuffle per accumulator (lane 0 ends up with the sums of the left x node pair,
  // lane 1 with the right pair) and each lane issues its half of the red.global.adds.
  const int nwork = ntask * NU * FSPLIT;
#pragma unroll 1
  for (int t = tid; t < ((nwork + 31) & ~31); t += FT) {
    const bool valid = t < nwork;
    const int hs = t % FSPLIT, h = hs & 1, tu = valid ? (t / FSPLIT) : 0;
    const int task = tu / NU, u = tu - task * NU;
    const bool isJ = u < NCJ;
    const int tw = tasks[task];
    const int key = tw & 0x7FF, start = (tw >> 11) & 0x7FF, n = valid ? (tw >> 22) : 0;
    const double* rb = isJ ? recJ : recR;
    const double* ra = rb + (4 + ((!ENV && isJ) ? u : 0)) * FSTR;
    const unsigned char bit = isJ ? 1 : 2;
    cd a[2][2][NM];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int m = 0; m < NM; ++m) a[i][k][m] = cmake(0.0, 0.0);
    int any = 0;
#pragma unroll 2
    for (int q = hs; q < n; q += FSPLIT) {
      const int li = order[start + q];
      if (!(fast[li] & bit)) continue;
      any = 1;
      const double fx = rb[li], fr = rb[FSTR + li];
      const cd ph1 = cmake(rb[2 * FSTR + li], rb[3 * FSTR + li]);
      const cd amp = ENV ? cmake(ra[li], ra[FSTR + li]) : cmake(ra[li], 0.0);
      const double w00 = (1.0 - fx) * (1.0 - fr), w01 = (1.0 - fx) * fr, w10 = fx * (1.0 - fr), w11 = fx * fr;
      cd ph = cmake(1.0, 0.0);
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
        for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
          const int m = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
          const cd pm = sgn ? cconj(ph) : ph;
          if (!ENV && iO == 0) {  // e^{-i 0 theta} = 1: real amplitude, the imaginary parts stay 0
            a[0][0][m].x += w00 * amp.x; a[0][1][m].x += w01 * amp.x;
            a[1][0][m].x += w10 * amp.x; a[1][1][m].x += w11 * amp.x;
            continue;
          }
          const cd f = ENV ? cmul(amp, pm) : cscale(amp.x, pm);
          a[0][0][m].x += w00 * f.x; a[0][0][m].y += w00 * f.y;
          a[0][1][m].x += w01 * f.x; a[0][1][m].y += w01 * f.y;
          a[1][0][m].x += w10 * f.x; a[1][0][m].y += w10 * f.y;
          a[1][1][m].x += w11 * f.x; a[1][1][m].y += w11 * f.y;
        }
      }
    }
    if (FSPLIT == 4) {  // lanes hs and hs ^ 2 first pool their partial sums
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            a[i][k][m].x += __shfl_xor_sync(0xffffffffu, a[i][k][m].x, 2);
            if (ENV || m > 0) a[i][k][m].y += __shfl_xor_sync(0xffffffffu, a[i][k][m].y, 2);
          }
      any |= __shfl_xor_sync(0xffffffffu, any, 2);
    }
    // lane h keeps x node h: send the other node's partial sums to the partner, add what it sends
    cd mine[2][NM];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const cd keep = h ? a[1][k][m] : a[0][k][m], give = h ? a[0][k][m] : a[1][k][m];
        mine[k][m].x = keep.x + __shfl_xor_sync(0xffffffffu, give.x, 1);
        mine[k][m].y = (!ENV && m == 0) ? 0.0 : keep.y + __shfl_xor_sync(0xffffffffu, give.y, 1);
      }
    any |= __shfl_xor_sync(0xffffffffu, any, 1);
    if (!valid || !any || hs >= 2) continue;
    const int kr = key / FBX, kx = key - kr * FBX;
    const i64 gx = (i64)ix0 + kx + h, gr = (i64)ir0 + kr;
    if (gx < klo || gx > khi) continue;
    const int l = isJ ? (ENV ? 2 : u) : 0;
    cd* pl = (isJ ? J : Rho) + plane * (g.nm * l) + gx + g.nxn * gr;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      if (!ENV && m == 0) {  // imaginary sums are exactly 0
        atomicAdd(&pl[0].x, mine[0][0].x);
        atomicAdd(&pl[g.nxn].x, mine[1][0].x);
        continue;
      }
      red_add(pl + plane * m, mine[0][m]);
      red_add(pl + plane * m + g.nxn, mine[1][m]);
    }
  }

#if CHB_FSYNC & 1
  __syncthreads();
#endif
  FPROF_MARK(4)
  // ---- (G) particles that changed cell: one thread per (particle, component, node), so that the few of
  // them cost a few warp-wide red.global.add instead of serialising inside divergent warps
#pragma unroll 1
  for (int pass = 0; pass < (SC ? 2 : 1); ++pass) {
    const bool isJ = pass == 0;
    const int per = (isJ ? NCJ : 1) * 4;
    const int nitems = sh.nslow[pass] * per;
    const double* rb = isJ ? recJ : recR;
    const unsigned short* lst = isJ ? slowJ : slowR;
    const unsigned* cells = isJ ? cellJ : cellR;
    cd* grid = isJ ? J : Rho;
#pragma unroll 1
    for (int t = tid; t < nitems; t += FT) {
      const int q = t / per, r = t - q * per;
      const int u = r >> 2, node = r & 3;
      const int i = node >> 1, k = node & 1;
      const int li = lst[q];
      const unsigned cell = cells[li];
      const i64 gx = (i64)(cell & 0xFFFFFu) - 1 + i, gr = (i64)(cell >> 20) + k;
      const bool keep = gx >= 0 && gx <= g.nxn - 1 && (!ch.on || chunk_keep(ch, cr.chunk, gx, g.nxn));
      if (!keep) continue;
      const double fx = rb[li], fr = rb[FSTR + li];
      const cd ph1 = cmake(rb[2 * FSTR + li], rb[3 * FSTR + li]);
      const double* ra = rb + (4 + ((!ENV && isJ) ? u : 0)) * FSTR;
      const cd amp = ENV ? cmake(ra[li], ra[FSTR + li]) : cmake(ra[li], 0.0);
      const double wgt = (i ? fx : 1.0 - fx) * (k ? fr : 1.0 - fr);
      const int l = isJ ? (ENV ? 2 : u) : 0;
      cd* pl = grid + plane * (g.nm * l) + gx + g.nxn * gr;
      cd ph = cmake(1.0, 0.0);
#pragma unroll
      for (int iO = 0; iO <= NKO; ++iO) {
        if (iO > 0) ph = cmul(ph, ph1);
#pragma unroll
        for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
          const int m = ENV ? (NKO + (sgn ? -iO : iO)) : iO;
          const cd pm = sgn ? cconj(ph) : ph;
          const cd f = ENV ? cmul(amp, pm) : cscale(amp.x, pm);
          red_add(pl + plane * m, cscale(wgt, f));
        }
      }
    }
  }
#if CHB_FSYNC & 2
  __syncthreads();
#endif
  FPROF_MARK(5)
  if (c_fprof_on && tid == 0) atomicAdd(&g_fprof[7], 1ull);
}

constexpr size_t F_SMEM = sizeof(double) * FNF * FSTR + sizeof(int) * (FBINS + FMAXTASK) +
                          4 * sizeof(unsigned short) * FNPB + 2 * sizeof(unsigned) * FNPB + FNPB;

template <int ENV, int SC, int MODE>
int launch_fused_nm(cudaStream_t st, double* x, double* xh, double* mom, const double* w, i64 cap, const cd* Fld, cd* J,
                    cd* Rho, const GridGeom& g, const ChunkSpec& ch, double dt_2, double dt, const DeviceSet& und,
                    const SortedSpec& sp, double leftX_J, double leftX_R) {
#define CHB_FUSED(NMV)                                                                                                 \
  case NMV: {                                                                                                          \
    static std::atomic<unsigned long long> attr{0}; /* the attribute is per device */                                  \
    int dev = 0;                                                                                                       \
    cudaGetDevice(&dev);                                                                                               \
    if (!(attr.load(std::memory_order_relaxed) & (1ull << (dev & 63)))) {                                              \
      CHB_CUDA(cudaFuncSetAttribute(fused_pass_k<ENV, NMV, SC, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    (int)F_SMEM));                                                                     \
      CHB_CUDA(cudaFuncSetAttribute(fused_particles_k<ENV, NMV, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                    (int)F_SMEM));                                                                     \
      attr.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);                                                    \
    }                                                                                                                  \
    if (MODE == 2)                                                                                                     \
      fused_pass_k<ENV, NMV, SC, 2><<<sp.ncta, FT, F_SMEM, st>>>(x, xh, mom, w, cap, Fld, J, Rho, g, ch, dt_2, dt, und, \
                                                                sp, leftX_J, leftX_R);                                 \
    else                                                                                                               \
      fused_particles_k<ENV, NMV, SC><<<sp.ncta, FT, F_SMEM, st>>>(x, xh, mom, w, cap, Fld, J, Rho, g, ch, dt_2, dt,    \
                                                                  und, sp, leftX_J, leftX_R);                          \
  } break;
  switch ((int)g.nm) {
    CHB_FUSED(1)
    CHB_FUSED(2)
    CHB_FUSED(3)
    CHB_FUSED(5)
    default: return -1;
  }
#undef CHB_FUSED
  CHB_LAUNCH_CHECK();
  return 0;
}

template <int ENV, int COORDS>
int launch_gpc_nm(cudaStream_t st, double* x, double* xh, double* mom, const double* w, i64 cap, const cd* Fld,
                  const GridGeom& g, double dt_2, double dt, const DeviceSet& und, i64 np) {
  // consecutive particles share cells: keep a CTA on consecutive particles (no grid stride across the whole array
  // unless the array is larger than the grid allows)
  const int grid = (int)std::min<i64>((np + 255) / 256, (i64)1 << 30);
  switch ((int)g.nm) {
    case 1: gather_push_coords_k<ENV, 1, COORDS><<<grid, 256, 0, st>>>(x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np); break;
    case 2: gather_push_coords_k<ENV, 2, COORDS><<<grid, 256, 0, st>>>(x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np); break;
    case 3: gather_push_coords_k<ENV, 3, COORDS><<<grid, 256, 0, st>>>(x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np); break;
    case 5: gather_push_coords_k<ENV, 5, COORDS><<<grid, 256, 0, st>>>(x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np); break;
    default: return -1;
  }
  CHB_LAUNCH_CHECK();
  return 0;
}
}  // namespace

void fused_profile_enable(int on) {
  cudaMemcpyToSymbol(c_fprof_on, &on, sizeof(int));
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_fprof, z, sizeof(z));
}
void fused_profile_read(unsigned long long out[8]) { cudaMemcpyFromSymbol(out, g_fprof, 8 * sizeof(unsigned long long)); }

// returns -1 when the mode count has no instantiation (the caller falls back to the separate kernels).
// deposit_only: the half after a re-binning step's sort -- J from the stored x_half and p, rho from the stored x
int launch_fused_particles(cudaStream_t st, int env, int space_charge, double* x, double* xh, double* mom,
                           const double* w, i64 cap, const cd* Fld, cd* J, cd* Rho, const GridGeom& g,
                           const ChunkSpec& ch, double push_dt, double dt, const DeviceSet& und, const SortedSpec& sp,
                           double leftX_J, double leftX_R, int deposit_only) {
  if (sp.ncta <= 0) return 0;
  if (env && (g.nm % 2) != 1) { set_error("envelope kernels need an odd number of mode slots"); return 2; }
  const double dt_2 = 0.5 * push_dt;
#define CHB_GO(E, S, M) return launch_fused_nm<E, S, M>(st, x, xh, mom, w, cap, Fld, J, Rho, g, ch, dt_2, dt, und, sp, leftX_J, leftX_R)
  if (deposit_only) {
    if (env) { if (space_charge) CHB_GO(1, 1, 2); CHB_GO(1, 0, 2); }
    if (space_charge) CHB_GO(0, 1, 2);
    CHB_GO(0, 0, 2);
  }
  if (env) { if (space_charge) CHB_GO(1, 1, 0); CHB_GO(1, 0, 0); }
  if (space_charge) CHB_GO(0, 1, 0);
  CHB_GO(0, 0, 0);
#undef CHB_GO
}

// gather + device + Boris push of step k and push_coords of step k+1, no deposit: the half of a re-binning step
// before its sort (and the tail of a step() call).  -1: no instantiation for this mode count.
int launch_gather_push_coords(cudaStream_t st, int env, double* x, double* xh, double* mom, const double* w, i64 cap,
                              const cd* Fld, const GridGeom& g, double push_dt, double dt, const DeviceSet& und, i64 np,
                              int coords) {
  if (np <= 0) return 0;
  if (env && (g.nm % 2) != 1) { set_error("envelope kernels need an odd number of mode slots"); return 2; }
  const double dt_2 = 0.5 * push_dt;
  if (coords) {
    if (env) return launch_gpc_nm<1, 1>(st, x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np);
    return launch_gpc_nm<0, 1>(st, x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np);
  }
  if (env) return launch_gpc_nm<1, 0>(st, x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np);
  return launch_gpc_nm<0, 0>(st, x, xh, mom, w, cap, Fld, g, dt_2, dt, und, np);
}

}  // namespace chb
