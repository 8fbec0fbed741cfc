// particles.cu -- particle kernels of the PIC cycle: Boris push, leap-frog, field gather,
// charge/current deposition with azimuthal modes, binning and permutation.
//
// Replaces reference f90/particle_tools.f90 and f90/grid_deps{,_chnk,_env,_env_chnk}.f90.
// All kernels take particle views (PView/CPView) so the same code runs on the reference's
// (3,Np) layout (C-ABI host path) and on the engine's structure-of-arrays storage.
#include "common.cuh"
#include "kernels.cuh"
#include "particle_dev.cuh"

namespace chb {

// ------------------------------------------------------------------------------------------
// Boris push (particle_tools.f90:18-56).  96 B/particle of HBM traffic, no reuse: pure stream.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) push_velocs_k(PView mom, CPView fld, double dt_2, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  double px = mom.at(0, ip), py = mom.at(1, ip), pz = mom.at(2, ip);
  boris(px, py, pz, fld.at(0, ip), fld.at(1, ip), fld.at(2, ip), fld.at(3, ip), fld.at(4, ip), fld.at(5, ip), dt_2);
  mom.at(0, ip) = px;
  mom.at(1, ip) = py;
  mom.at(2, ip) = pz;
}

// The reference's (ncomp, Np) arrays (component fastest) through shared memory: a CTA moves the contiguous block of its
// 256 particles with coalesced accesses and each thread picks its components there -- a thread-per-particle walk of the
// array touches every 32-byte sector with three (six) separate instructions.
constexpr int kAosT = 256;
template <int NC>
__device__ __forceinline__ void aos_block_load(double* s, const double* g, i64 first, int n) {
  for (int i = threadIdx.x; i < NC * n; i += kAosT) s[i] = g[NC * first + i];  // plain loads: some of these arrays are in/out
}
template <int NC>
__device__ __forceinline__ void aos_block_store(double* __restrict__ g, const double* s, i64 first, int n) {
  for (int i = threadIdx.x; i < NC * n; i += kAosT) g[NC * first + i] = s[i];
}

__global__ void __launch_bounds__(kAosT) push_velocs_aos_k(double* __restrict__ mom, const double* __restrict__ fld,
                                                           double dt_2, i64 np) {
  __shared__ double sp[3 * kAosT], sf[6 * kAosT];
  const i64 first = (i64)blockIdx.x * kAosT;
  const int n = (int)(np - first < kAosT ? np - first : kAosT), t = threadIdx.x;
  aos_block_load<3>(sp, mom, first, n);
  aos_block_load<6>(sf, fld, first, n);
  __syncthreads();
  if (t < n) {
    double px = sp[3 * t], py = sp[3 * t + 1], pz = sp[3 * t + 2];
    boris(px, py, pz, sf[6 * t], sf[6 * t + 1], sf[6 * t + 2], sf[6 * t + 3], sf[6 * t + 4], sf[6 * t + 5], dt_2);
    sp[3 * t] = px; sp[3 * t + 1] = py; sp[3 * t + 2] = pz;
  }
  __syncthreads();
  aos_block_store<3>(mom, sp, first, n);
}

static inline bool is_aos(const double* p, i64 cs, i64 ps, int ncomp) { return p && cs == 1 && ps == ncomp; }

int launch_push_velocs(cudaStream_t st, PView mom, CPView fld, double dt, i64 np) {
  if (np <= 0) return 0;
  if (is_aos(mom.p, mom.cs, mom.ps, 3) && is_aos(fld.p, fld.cs, fld.ps, 6))
    push_velocs_aos_k<<<grid_for(np, kAosT), kAosT, 0, st>>>(mom.p, fld.p, 0.5 * dt, np);
  else
    push_velocs_k<<<grid_for(np, 256), 256, 0, st>>>(mom, fld, 0.5 * dt, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Leap-frog position update (particle_tools.f90:58-82)
// ------------------------------------------------------------------------------------------
// The position update is written with explicit round-to-nearest multiplies and adds (no FMA
// contraction) in the reference's operation order: the envelope kernels multiply positions by
// kx0 ~ 1e5..1e6 inside exp(+-i kx0 x), so a 1-ulp difference in x is a 1e-10 phase difference.
// With strict IEEE arithmetic the new positions are bit-identical to the CPU path.
__global__ void __launch_bounds__(256) push_coords_k(PView x, CPView mom, PView xc, double dt, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const double px = mom.at(0, ip), py = mom.at(1, ip), pz = mom.at(2, ip);
  const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz));
  const double dt_gp = __ddiv_rn(dt, __dsqrt_rn(__dadd_rn(1.0, p2)));
  const double p[3] = {px, py, pz};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double x0 = x.at(c, ip);
    const double x1 = __dadd_rn(x0, __dmul_rn(p[c], dt_gp));
    x.at(c, ip) = x1;
    xc.at(c, ip) = __dmul_rn(0.5, __dadd_rn(x0, x1));
  }
}

__global__ void __launch_bounds__(kAosT) push_coords_aos_k(double* __restrict__ x, const double* __restrict__ mom,
                                                           double* __restrict__ xc, double dt, i64 np) {
  __shared__ double sx[3 * kAosT], sp[3 * kAosT], sc[3 * kAosT];
  const i64 first = (i64)blockIdx.x * kAosT;
  const int n = (int)(np - first < kAosT ? np - first : kAosT), t = threadIdx.x;
  aos_block_load<3>(sx, x, first, n);
  aos_block_load<3>(sp, mom, first, n);
  __syncthreads();
  if (t < n) {  // the arithmetic of push_coords_k, operation for operation
    const double p[3] = {sp[3 * t], sp[3 * t + 1], sp[3 * t + 2]};
    const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(p[0], p[0]), __dmul_rn(p[1], p[1])), __dmul_rn(p[2], p[2]));
    const double dt_gp = __ddiv_rn(dt, __dsqrt_rn(__dadd_rn(1.0, p2)));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double x0 = sx[3 * t + c];
      const double x1 = __dadd_rn(x0, __dmul_rn(p[c], dt_gp));
      sx[3 * t + c] = x1;
      sc[3 * t + c] = __dmul_rn(0.5, __dadd_rn(x0, x1));
    }
  }
  __syncthreads();
  aos_block_store<3>(x, sx, first, n);
  aos_block_store<3>(xc, sc, first, n);
}

int launch_push_coords(cudaStream_t st, PView x, CPView mom, PView xc, double dt, i64 np) {
  if (np <= 0) return 0;
  if (is_aos(x.p, x.cs, x.ps, 3) && is_aos(mom.p, mom.cs, mom.ps, 3) && is_aos(xc.p, xc.cs, xc.ps, 3) && x.p != xc.p)
    push_coords_aos_k<<<grid_for(np, kAosT), kAosT, 0, st>>>(x.p, mom.p, xc.p, dt, np);
  else
    push_coords_k<<<grid_for(np, 256), 256, 0, st>>>(x, mom, xc, dt, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Shape factors shared by gather and deposit (grid_deps.f90:46-53).
// Returns false when the particle is skipped (outside r range / outside the grid).
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Gather (proj_fld grid_deps.f90:149-217, proj_fld_env grid_deps_env.f90:164-238).
// One thread per particle; the 4 nodes x nm modes x 6 components are read through the
// read-only path (cell-sorted particles make these L1/L2 hits).
// ------------------------------------------------------------------------------------------
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
                                                     double dt_2, DeviceSet und, i64 np) {
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
  if (und.n) apply_devices(und, xp, yp, zp, F);
  double px = mom.at(0, ip), py = mom.at(1, ip), pz = mom.at(2, ip);
  boris(px, py, pz, F[0], F[1], F[2], F[3], F[4], F[5], dt_2);
  mom.at(0, ip) = px;
  mom.at(1, ip) = py;
  mom.at(2, ip) = pz;
}

int launch_gather_push(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView mom,
                       const GridGeom& g, double dt, const DeviceSet& und, i64 np) {
  if (np <= 0) return 0;
  if (env) gather_push_k<1><<<grid_for(np, 128), 128, 0, st>>>(x, w, Fld, mom, g, 0.5 * dt, und, np);
  else     gather_push_k<0><<<grid_for(np, 128), 128, 0, st>>>(x, w, Fld, mom, g, 0.5 * dt, und, np);
  CHB_LAUNCH_CHECK();
  return 0;
}

// devices.f90:18-297 (NEXT-1 row): external-field devices added to the per-particle field
__global__ void __launch_bounds__(256) devices_k(CPView x, PView fld, DeviceSet und, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  double F[6] = {0, 0, 0, 0, 0, 0};
  apply_devices(und, x.at(0, ip), x.at(1, ip), x.at(2, ip), F);
  fld.at(2, ip) += F[2];
  fld.at(3, ip) += F[3];
  fld.at(4, ip) += F[4];
}

int launch_devices(cudaStream_t st, CPView x, PView fld, const DeviceSet& und, i64 np) {
  if (np <= 0) return 0;
  devices_k<<<grid_for(np, 256), 256, 0, st>>>(x, fld, und, np);
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
template <int ENV, int CURR>
__global__ void __launch_bounds__(128) deposit_direct_k(CPView x, CPView mom, const double* __restrict__ w,
                                                        cd* __restrict__ grid, GridGeom g, ChunkSpec ch, i64 np) {
  const i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const double wp = __ldg(w + ip);
  if (wp == 0.0) return;
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
  double p0 = 0, p1 = 0, p2 = 0;
  if (CURR) { p0 = mom.at(0, ip); p1 = mom.at(1, ip); p2 = mom.at(2, ip); }
  deposit_one<ENV, CURR>(g, ch, chunk, grid, x.at(0, ip), x.at(1, ip), x.at(2, ip), p0, p1, p2, wp);
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

// (ncomp, np) Fortran-ordered array (component fastest) -> ncomp planes of `cap` doubles: the layout the binned
// kernels read (used by the per-function entry points, whose arguments are the reference's (3, Np) arrays)
__global__ void __launch_bounds__(256) planes_from_aos_k(double* __restrict__ dst, const double* __restrict__ src, int ncomp,
                                                         i64 cap, i64 np) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= np * ncomp) return;
  const i64 ip = e / ncomp;
  dst[(e - ip * ncomp) * cap + ip] = src[e];
}
int launch_planes_from_aos(cudaStream_t st, double* dst, const double* src, int ncomp, i64 cap, i64 np) {
  if (np <= 0) return 0;
  planes_from_aos_k<<<grid_for(np * ncomp, 256), 256, 0, st>>>(dst, src, ncomp, cap, np);
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
