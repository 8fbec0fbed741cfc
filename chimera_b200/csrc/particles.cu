// particles.cu -- particle kernels of the PIC cycle: Boris push, leap-frog, field gather,
// charge/current deposition with azimuthal modes, binning and permutation.
//
// Replaces reference f90/particle_tools.f90 and f90/grid_deps{,_chnk,_env,_env_chnk}.f90.
// All kernels take particle views (PView/CPView) so the same code runs on the reference's
// (3,Np) layout (C-ABI host path) and on the engine's structure-of-arrays storage.
#include "common.cuh"
#include "kernels.cuh"

namespace chb {

// ------------------------------------------------------------------------------------------
// Boris push (particle_tools.f90:18-56).  96 B/particle of HBM traffic, no reuse: pure stream.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void boris(double& px, double& py, double& pz, double ex, double ey, double ez,
                                      double bx, double by, double bz, double dt_2) {
  const double umx = px + dt_2 * ex, umy = py + dt_2 * ey, umz = pz + dt_2 * ez;
  const double gamma = sqrt(1.0 + (umx * umx + umy * umy + umz * umz));
  const double tx = dt_2 * bx / gamma, ty = dt_2 * by / gamma, tz = dt_2 * bz / gamma;
  const double t2 = tx * tx + ty * ty + tz * tz;
  const double sx = 2 * tx / (1 + t2), sy = 2 * ty / (1 + t2), sz = 2 * tz / (1 + t2);
  const double u0x = umx + umy * tz - umz * ty;
  const double u0y = umy - umx * tz + umz * tx;
  const double u0z = umz + umx * ty - umy * tx;
  const double upx = umx + u0y * sz - u0z * sy;
  const double upy = umy - u0x * sz + u0z * sx;
  const double upz = umz + u0x * sy - u0y * sx;
  px = upx + dt_2 * ex;
  py = upy + dt_2 * ey;
  pz = upz + dt_2 * ez;
}

__global__ void __launch_bounds__(256) push_velocs_k(PView mom, CPView fld, double dt_2, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  double px = mom.at(0, ip), py = mom.at(1, ip), pz = mom.at(2, ip);
  boris(px, py, pz, fld.at(0, ip), fld.at(1, ip), fld.at(2, ip), fld.at(3, ip), fld.at(4, ip), fld.at(5, ip), dt_2);
  mom.at(0, ip) = px;
  mom.at(1, ip) = py;
  mom.at(2, ip) = pz;
}

int launch_push_velocs(cudaStream_t st, PView mom, CPView fld, double dt, i64 np) {
  if (np <= 0) return 0;
  push_velocs_k<<<grid_for(np, 256), 256, 0, st>>>(mom, fld, 0.5 * dt, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Leap-frog position update (particle_tools.f90:58-82)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) push_coords_k(PView x, CPView mom, PView xc, double dt, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const double px = mom.at(0, ip), py = mom.at(1, ip), pz = mom.at(2, ip);
  const double dt_gp = dt / sqrt(1.0 + (px * px + py * py + pz * pz));
  const double p[3] = {px, py, pz};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double x0 = x.at(c, ip);
    const double x1 = x0 + p[c] * dt_gp;
    x.at(c, ip) = x1;
    xc.at(c, ip) = 0.5 * (x0 + x1);
  }
}

int launch_push_coords(cudaStream_t st, PView x, CPView mom, PView xc, double dt, i64 np) {
  if (np <= 0) return 0;
  push_coords_k<<<grid_for(np, 256), 256, 0, st>>>(x, mom, xc, dt, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Shape factors shared by gather and deposit (grid_deps.f90:46-53).
// Returns false when the particle is skipped (outside r range / outside the grid).
// ------------------------------------------------------------------------------------------
struct Shape {
  i64 ix, ir;
  double sx0, sx1, sr0, sr1;
  double rp;
};

__device__ __forceinline__ bool make_shape(const GridGeom& g, double xp, double yp, double zp, Shape& s) {
  s.rp = sqrt(yp * yp + zp * zp);
  if (s.rp >= g.rmax) return false;
  const double xs = (xp - g.leftX) * g.dx_inv;
  s.ix = (i64)floor(xs);
  s.ir = (i64)floor((s.rp - g.r0) * g.dr_inv);
  if (s.ir < 0 || s.ir > g.nrn - 2) return false;
  s.sx1 = xs - (double)s.ix;
  s.sx0 = 1.0 - s.sx1;
  s.sr1 = (s.rp - __ldg(g.Rgrid + s.ir)) * g.dr_inv;
  s.sr0 = 1.0 - s.sr1;
  return true;
}

// ------------------------------------------------------------------------------------------
// Gather (proj_fld grid_deps.f90:149-217, proj_fld_env grid_deps_env.f90:164-238).
// One thread per particle; the 4 nodes x nm modes x 6 components are read through the
// read-only path (cell-sorted particles make these L1/L2 hits).
// ------------------------------------------------------------------------------------------
template <int ENV>
__device__ __forceinline__ bool gather_one(const GridGeom& g, const cd* __restrict__ Fld, double xp, double yp,
                                           double zp, double F[6]) {
  Shape s;
  if (!make_shape(g, xp, yp, zp, s)) return false;
  if (s.ix < 0 || s.ix > g.nxn - 2) return false;
  const int nko = ENV ? (int)(g.nm - 1) / 2 : (int)g.nm - 1;
  // Q4: phase at r = 0 is 0 (real solver) or 1 (envelope solver)
  cd ph1 = (s.rp > 0.0) ? cmake(yp / s.rp, zp / s.rp) : (ENV ? cmake(1.0, 0.0) : cmake(0.0, 0.0));
  cd car = cmake(1.0, 0.0);
  if (ENV) {
    double sn, cs;
    sincos(xp * g.kx0, &sn, &cs);
    car = cmake(cs, sn);
  }
  const double w00 = s.sr0 * s.sx0, w10 = s.sr0 * s.sx1, w01 = s.sr1 * s.sx0, w11 = s.sr1 * s.sx1;
  const i64 plane = g.nxn * g.nrn;
  const i64 node = s.ix + g.nxn * s.ir;
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
        const cd* pl = Fld + plane * (slot + g.nm * l) + node;
        const cd f00 = __ldg(pl), f10 = __ldg(pl + 1), f01 = __ldg(pl + g.nxn), f11 = __ldg(pl + g.nxn + 1);
        double acc = 0.0;
        acc += p00.x * f00.x - p00.y * f00.y;
        acc += p10.x * f10.x - p10.y * f10.y;
        acc += p01.x * f01.x - p01.y * f01.y;
        acc += p11.x * f11.x - p11.y * f11.y;
        F[l] += acc;
      }
    }
  }
  return true;
}

template <int ENV>
__global__ void __launch_bounds__(128) gather_k(CPView x, const double* __restrict__ w, const cd* __restrict__ Fld,
                                                PView out, GridGeom g, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  if (__ldg(w + ip) == 0.0) return;
  double F[6];
  if (!gather_one<ENV>(g, Fld, x.at(0, ip), x.at(1, ip), x.at(2, ip), F)) return;
#pragma unroll
  for (int l = 0; l < 6; ++l) out.at(l, ip) += F[l];
}

int launch_gather(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView out, const GridGeom& g,
                  i64 np) {
  if (np <= 0) return 0;
  if (env) gather_k<1><<<grid_for(np, 128), 128, 0, st>>>(x, w, Fld, out, g, np);
  else     gather_k<0><<<grid_for(np, 128), 128, 0, st>>>(x, w, Fld, out, g, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// Fused gather + Boris push for the resident engine: the per-particle field never goes to HBM.
// (gather at `x`, optional analytic undulator device, then push `mom`.)
template <int ENV>
__global__ void __launch_bounds__(128) gather_push_k(CPView x, const double* __restrict__ w,
                                                     const cd* __restrict__ Fld, PView mom, GridGeom g,
                                                     double dt_2, UndulParams und, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  double F[6] = {0, 0, 0, 0, 0, 0};
  const double xp = x.at(0, ip), yp = x.at(1, ip), zp = x.at(2, ip);
  if (__ldg(w + ip) != 0.0) {
    double G[6];
    if (gather_one<ENV>(g, Fld, xp, yp, zp, G)) {
#pragma unroll
      for (int l = 0; l < 6; ++l) F[l] = G[l];
    }
  }
  if (und.on) undul_field(und, xp, yp, F);
  double px = mom.at(0, ip), py = mom.at(1, ip), pz = mom.at(2, ip);
  boris(px, py, pz, F[0], F[1], F[2], F[3], F[4], F[5], dt_2);
  mom.at(0, ip) = px;
  mom.at(1, ip) = py;
  mom.at(2, ip) = pz;
}

int launch_gather_push(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView mom,
                       const GridGeom& g, double dt, const UndulParams& und, i64 np) {
  if (np <= 0) return 0;
  if (env) gather_push_k<1><<<grid_for(np, 128), 128, 0, st>>>(x, w, Fld, mom, g, 0.5 * dt, und, np);
  else     gather_push_k<0><<<grid_for(np, 128), 128, 0, st>>>(x, w, Fld, mom, g, 0.5 * dt, und, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// devices.f90:162-203 (NEXT-1 row): analytic planar undulator added to the per-particle field
__global__ void __launch_bounds__(256) undul_k(CPView x, PView fld, UndulParams und, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  double F[6] = {0, 0, 0, 0, 0, 0};
  undul_field(und, x.at(0, ip), x.at(1, ip), F);
  fld.at(3, ip) += F[3];
  fld.at(4, ip) += F[4];
}

int launch_undul(cudaStream_t st, CPView x, PView fld, const UndulParams& und, i64 np) {
  if (np <= 0) return 0;
  undul_k<<<grid_for(np, 256), 256, 0, st>>>(x, fld, und, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Deposition, direct variant: one thread per particle, FP64 red.global.add per node value.
// Covers dep_curr/dep_dens x {plain, _chnk, _env, _env_chnk}
// (grid_deps.f90:18-147, grid_deps_chnk.f90, grid_deps_env.f90:18-162, grid_deps_env_chnk.f90).
// The chunked variants differ from the plain ones only in which edge contributions are dropped
// (Q3: the left guard buffer of chunk 0 and the right one of the last chunk are discarded).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(cd* dst, cd v) {
  atomicAdd(&dst->x, v.x);
  atomicAdd(&dst->y, v.y);
}

// chunk-edge predicate: may the contribution of a particle of chunk `c` to global node gx be kept?
__device__ __forceinline__ bool chunk_keep(const ChunkSpec& ch, int c, i64 gx, i64 nxn) {
  const i64 nxleft = (i64)c * ch.cs;
  const i64 lx = gx - nxleft;
  if (lx <= 0) {
    if (lx < -ch.guards) return false;              // outside loc_left: undefined in the reference
    return nxleft - ch.guards >= 0;                  // grid_deps_chnk.f90:115
  } else if (lx >= ch.cs) {
    if (lx > ch.cs + ch.guards) return false;
    return nxleft + ch.cs + ch.guards <= nxn - 1;    // grid_deps_chnk.f90:110
  }
  return true;
}

template <int ENV, int CURR>
__global__ void __launch_bounds__(128) deposit_direct_k(CPView x, CPView mom, const double* __restrict__ w,
                                                        cd* __restrict__ grid, GridGeom g, ChunkSpec ch, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const double wp = __ldg(w + ip);
  if (wp == 0.0) return;
  const double xp = x.at(0, ip), yp = x.at(1, ip), zp = x.at(2, ip);
  Shape s;
  if (!make_shape(g, xp, yp, zp, s)) return;
  double v[3] = {1.0, 1.0, 1.0};
  if (CURR) {
    v[0] = mom.at(0, ip); v[1] = mom.at(1, ip); v[2] = mom.at(2, ip);
    if (fabs(v[0]) + fabs(v[1]) + fabs(v[2]) == 0.0) return;
    const double gp = sqrt(1.0 + v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
    for (int l = 0; l < 3; ++l) v[l] = ENV ? v[l] / gp : v[l] * wp / gp;
  }
  cd wpc = cmake(wp, 0.0);
  if (ENV) {
    double sn, cs;
    sincos(xp * g.kx0, &sn, &cs);
    wpc = cmake(wp * cs, -wp * sn);
  }
  int chunk = 0;
  if (ch.on) {  // chunk = the index range of IndInChunk this particle sits in
    int lo = 0, hi = ch.nchnk;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ip >= __ldg(ch.ind + mid)) lo = mid; else hi = mid;
    }
    chunk = lo;
    if (ip >= __ldg(ch.ind + ch.nchnk)) return;  // beyond the last chunk: not deposited by the reference
  }
  const int nko = ENV ? (int)(g.nm - 1) / 2 : (int)g.nm - 1;
  const cd ph1 = (s.rp > 0.0) ? cmake(yp / s.rp, -zp / s.rp) : cmake(0.0, 0.0);
  // cell weights (times the complex particle weight where the variant has one)
  cd cw[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double sh = (i ? s.sx1 : s.sx0) * (k ? s.sr1 : s.sr0);
      if (CURR) cw[i][k] = ENV ? cscale(sh, wpc) : cmake(sh, 0.0);
      else      cw[i][k] = ENV ? cmul(cscale(sh, wpc), wpc) : cmake(sh * wp, 0.0);  // Q2
    }
  bool keep[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const i64 gx = s.ix + i;
    keep[i] = ch.on ? chunk_keep(ch, chunk, gx, g.nxn) : (gx >= 0 && gx <= g.nxn - 1);
    if (gx < 0 || gx > g.nxn - 1) keep[i] = false;
  }
  const i64 plane = g.nxn * g.nrn;
  const int l0 = (CURR && ENV) ? 2 : 0;  // Q1
  const int l1 = CURR ? 3 : 1;
  cd ph = cmake(1.0, 0.0);
  for (int iO = 0; iO <= nko; ++iO) {
    if (iO > 0) ph = cmul(ph, ph1);
    for (int sgn = 0; sgn < ((ENV && iO > 0) ? 2 : 1); ++sgn) {
      const cd phs = sgn ? cconj(ph) : ph;
      const int slot = ENV ? (nko + (sgn ? -iO : iO)) : iO;
      for (int l = l0; l < l1; ++l) {
        const cd f = cscale(v[l], phs);
        cd* pl = grid + plane * (slot + g.nm * l) + s.ix + g.nxn * s.ir;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!keep[i]) continue;
#pragma unroll
          for (int k = 0; k < 2; ++k) red_add(pl + i + g.nxn * k, cmul(f, cw[i][k]));
        }
      }
    }
  }
}

// ghost-row fold after deposition: J(:,1) -= J(:,0); J(:,0) = 0   (grid_deps.f90:80-85)
__global__ void __launch_bounds__(256) ghost_fold_k(cd* __restrict__ grid, i64 nxn, i64 nrn, i64 nplanes) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nxn * nplanes) return;
  const i64 ix = t % nxn, q = t / nxn;
  cd* pl = grid + nxn * nrn * q;
  const cd a = pl[ix];
  cd b = pl[ix + nxn];
  b.x -= a.x;
  b.y -= a.y;
  pl[ix + nxn] = b;
  pl[ix] = cmake(0.0, 0.0);
}

int launch_ghost_fold(cudaStream_t st, cd* grid, i64 nxn, i64 nrn, i64 nplanes) {
  ghost_fold_k<<<grid_for(nxn * nplanes, 256), 256, 0, st>>>(grid, nxn, nrn, nplanes);
  CHB_LAUNCH_CHECK();
  return 0;
}

int launch_deposit_direct(cudaStream_t st, int env, int curr, CPView x, CPView mom, const double* w, cd* grid,
                          const GridGeom& g, const ChunkSpec& ch, i64 np, bool fold) {
  if (g.nm > 2 * kMaxModes) { set_error("too many azimuthal modes (%lld)", g.nm); return 3; }
  if (np > 0) {
    const unsigned nb = grid_for(np, 128);
    if (env && curr)       deposit_direct_k<1, 1><<<nb, 128, 0, st>>>(x, mom, w, grid, g, ch, np);
    else if (env && !curr) deposit_direct_k<1, 0><<<nb, 128, 0, st>>>(x, mom, w, grid, g, ch, np);
    else if (curr)         deposit_direct_k<0, 1><<<nb, 128, 0, st>>>(x, mom, w, grid, g, ch, np);
    else                   deposit_direct_k<0, 0><<<nb, 128, 0, st>>>(x, mom, w, grid, g, ch, np);
    CHB_LAUNCH_CHECK();
  }
  if (fold) CHB_TRY(launch_ghost_fold(st, grid, g.nxn, g.nrn, g.nm * (curr ? 3 : 1)));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Binning: chunk id per particle (int8, -2 = leaves the domain), per-chunk counts, GoOut.
// particle_tools.f90:155-208.  Warp-aggregated counting into shared memory, then global.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chunk_bin_k(CPView x, int8_t* __restrict__ chunked, int* __restrict__ counts,
                                                   int* __restrict__ goout, double x0, double inv, double l0,
                                                   double l1, double l2, double l3, int nchnk, i64 np) {
  extern __shared__ int s_cnt[];  // nchnk + 1 (last = out)
  for (int i = threadIdx.x; i <= nchnk; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip < np) {
    const double xp = x.at(0, ip), yp = x.at(1, ip), zp = x.at(2, ip);
    i64 c = (i64)floor((xp - x0) * inv);
    const double r2 = yp * yp + zp * zp;
    int8_t id = -2;
    if (xp >= l0 && xp <= l1 && r2 >= l2 && r2 <= l3) {
      if (c < 0) c = 0;
      if (c > nchnk - 1) c = nchnk - 1;
      id = (int8_t)c;
      atomicAdd(&s_cnt[c], 1);
    } else {
      atomicAdd(&s_cnt[nchnk], 1);
    }
    chunked[ip] = id;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= nchnk; i += blockDim.x) {
    const int v = s_cnt[i];
    if (v) atomicAdd(i < nchnk ? &counts[i] : goout, v);
  }
}

int launch_chunk_bin(cudaStream_t st, CPView x, int8_t* chunked, int* counts /*nchnk, zeroed*/, int* goout /*zeroed*/,
                     double x0, double inv, const double lims[4], int nchnk, i64 np) {
  if (np <= 0) return 0;
  chunk_bin_k<<<grid_for(np, 256), 256, (nchnk + 1) * sizeof(int), st>>>(x, chunked, counts, goout, x0, inv, lims[0],
                                                                          lims[1], lims[2], lims[3], nchnk, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Permutation gather dat(:,i) = dat(:,idx(i))  (align_data_vec/scl, particle_tools.f90:270-324)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) permute_k(PView dst, CPView src, const i64* __restrict__ idx, int ncomp, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const i64 j = __ldg(idx + ip);
  for (int c = 0; c < ncomp; ++c) dst.at(c, ip) = src.at(c, j);
}

int launch_permute(cudaStream_t st, PView dst, CPView src, const i64* idx, int ncomp, i64 np) {
  if (np <= 0) return 0;
  permute_k<<<grid_for(np, 256), 256, 0, st>>>(dst, src, idx, ncomp, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// keep-flag for sortpartsout (particle_tools.f90:130-153) / sortoutghosts (:326-347)
__global__ void __launch_bounds__(256) inside_flag_k(CPView x, int* __restrict__ flag, double l0, double l1, double l2,
                                                     double l3, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const double xp = x.at(0, ip), yp = x.at(1, ip), zp = x.at(2, ip);
  const double r2 = yp * yp + zp * zp;
  flag[ip] = (xp >= l0 && xp <= l1 && r2 >= l2 && r2 <= l3) ? 1 : 0;
}

int launch_inside_flag(cudaStream_t st, CPView x, int* flag, const double lims[4], i64 np) {
  if (np <= 0) return 0;
  inside_flag_k<<<grid_for(np, 256), 256, 0, st>>>(x, flag, lims[0], lims[1], lims[2], lims[3], np);
  CHB_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) nonzero_flag_k(const double* __restrict__ v, int* __restrict__ flag, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip < np) flag[ip] = (v[ip] != 0.0) ? 1 : 0;
}

int launch_nonzero_flag(cudaStream_t st, const double* v, int* flag, i64 np) {
  if (np <= 0) return 0;
  nonzero_flag_k<<<grid_for(np, 256), 256, 0, st>>>(v, flag, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// scatter the indices of flagged particles to their compacted positions (pos = exclusive scan of flag)
__global__ void __launch_bounds__(256) compact_index_k(const int* __restrict__ flag, const int* __restrict__ pos,
                                                       int* __restrict__ out, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip < np && flag[ip]) out[pos[ip]] = (int)ip;
}

int launch_compact_index(cudaStream_t st, const int* flag, const int* pos, int* out, i64 np) {
  if (np <= 0) return 0;
  compact_index_k<<<grid_for(np, 256), 256, 0, st>>>(flag, pos, out, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace chb
