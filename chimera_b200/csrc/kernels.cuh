// kernels.cuh -- internal (device-pointer) launch API shared by the C-ABI host layer and the
// resident engine.  Everything here takes a stream and device pointers; no allocation.
#pragma once
#include <cmath>
#include "common.cuh"

namespace chb {

// geometry of the (x, r) grid as seen by gather/deposit (reference grid_deps.f90 dummy args)
struct GridGeom {
  double leftX, dx_inv, dr_inv, kx0;
  double r0, rmax;        // Rgrid(0), Rgrid(nr)
  const double* Rgrid;    // device, nrn entries
  i64 nxn, nrn, nm;       // x nodes, r nodes (incl. ghost), azimuthal-mode slots
};

// x-chunk bookkeeping of the *_chnk deposit variants (grid_deps_chnk.f90:38-47)
struct ChunkSpec {
  int on;
  const int* ind;  // device IndInChunk(0:nchnk)
  int nchnk, guards;
  i64 cs;          // chunk size in nodes = nxn / nchnk
};

// external-field devices (devices.f90:18-297), evaluated per particle between the field gather and the
// momentum push (species.py:258-277 make_device).  Up to kMaxDevices per launch; `t` = i_step * TimeStep.
enum DeviceKind {
  DEV_UNDUL_ANALYTIC = 1,        // devices.f90:162  p = a0, lambda, X0, Lx
  DEV_UNDUL_ANALYTIC_TAPER = 2,  // devices.f90:117  p = a0, lambda, X0, Lx, taper
  DEV_UNDUL_MAPPED = 3,          // devices.f90:18   p = lambda, Xleft, dx ; map a0(2, nx)
  DEV_UNDUL_MAPPED_TAP = 4,      // devices.f90:64   p = lambda, Xleft, dx, Lx, taper ; map a0(2, nx)
  DEV_PLANEWAVE = 5,             // devices.f90:205  p = a0, lambda, X0, Lx, ramp, theta, phi0
  DEV_GAUSSBEAM = 6              // devices.f90:254  a0 ; p = lambda, axis, x0, y0, z0, Lx, Ly, Lz
};
constexpr int kMaxDevices = 4;
struct DeviceSpec {
  int kind, nx;
  const double* map;  // device pointer, (2, nx) Fortran order (mapped undulators)
  double a0;
  double p[8];
  double c[3];        // host-side derived constants: wave number; sin, cos of the plane-wave angle
};
struct DeviceSet {
  int n;
  double t;
  DeviceSpec d[kMaxDevices];
};
// fills the derived constants of a spec from its parameters
static inline void device_finish(DeviceSpec& d) {
  const double pi = 4.0 * atan(1.0);
  const double lambda = (d.kind == DEV_UNDUL_MAPPED || d.kind == DEV_UNDUL_MAPPED_TAP || d.kind == DEV_GAUSSBEAM) ? d.p[0] : d.p[1];
  d.c[0] = 2.0 * pi / lambda;
  d.c[1] = d.kind == DEV_PLANEWAVE ? sin(d.p[5]) : 0.0;
  d.c[2] = d.kind == DEV_PLANEWAVE ? cos(d.p[5]) : 1.0;
}
static inline DeviceSet one_device(int kind, double a0, const double* params, int nparams, const double* map, int nx,
                                   double t) {
  DeviceSet s;
  memset(&s, 0, sizeof(s));
  s.n = 1;
  s.t = t;
  s.d[0].kind = kind; s.d[0].nx = nx; s.d[0].map = map; s.d[0].a0 = a0;
  for (int i = 0; i < nparams && i < 8; ++i) s.d[0].p[i] = params[i];
  device_finish(s.d[0]);
  return s;
}

__device__ __forceinline__ double ramp_ampl(double x, double X0, double Lx, double ramp) {
  if (x <= X0 || x >= X0 + Lx) return 0.0;
  if (x > X0 && x < X0 + ramp) return (x - X0) / ramp;
  if (x > X0 + Lx - ramp && x < X0 + Lx) return (X0 + Lx - x) / ramp;
  return 1.0;
}

__device__ __forceinline__ void apply_devices(const DeviceSet& ds, double x, double y, double z, double F[6]) {
  for (int i = 0; i < ds.n; ++i) {
    const DeviceSpec& d = ds.d[i];
    const double k = d.c[0];
    switch (d.kind) {
      case DEV_UNDUL_ANALYTIC:
      case DEV_UNDUL_ANALYTIC_TAPER: {
        const double X0 = d.p[2], Lx = d.p[3];
        double ampl = ramp_ampl(x, X0, Lx, d.p[1]);
        if (d.kind == DEV_UNDUL_ANALYTIC_TAPER) ampl = ampl * (1 + d.p[4] * (x - X0 - 0.5 * Lx) / (0.5 * Lx));
        ampl *= d.p[0];
        F[4] += ampl * sin(k * (x - X0)) * cosh(k * y);
        F[3] += ampl * cos(k * (x - X0)) * sinh(k * y);
      } break;
      case DEV_UNDUL_MAPPED:
      case DEV_UNDUL_MAPPED_TAP: {
        const double Xleft = d.p[1], dx = d.p[2], dx_inv = 1.0 / dx, Xright = Xleft + d.nx * dx;
        if (x < Xleft + dx || x > Xright - dx) break;
        const i64 ix = (i64)floor((x - Xleft) * dx_inv + 0.5);
        const double ddx = (x - Xleft) * dx_inv - (double)ix;
        const double S0[3] = {0.5 * (0.5 - ddx) * (0.5 - ddx), 0.75 - ddx * ddx, 0.5 * (0.5 + ddx) * (0.5 + ddx)};
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const i64 kk = ix - 1 + j;  // 1-based node; Q12: node 0 (read out of bounds by the reference) counts as 0
          if (kk < 1 || kk > d.nx) continue;
          s1 += S0[j] * __ldg(d.map + 2 * (kk - 1));
          s2 += S0[j] * __ldg(d.map + 2 * (kk - 1) + 1);
        }
        double amp = 1.0;
        if (d.kind == DEV_UNDUL_MAPPED_TAP) {
          const double Lx = d.p[3], x_shift = 0.5 * (d.nx * dx - Lx);
          amp = 1 + d.p[4] * (x - x_shift - Xleft - 0.5 * Lx) / (0.5 * Lx);
        }
        F[4] += amp * s1 * cosh(k * y);
        F[3] += amp * s2 * sinh(k * y);
      } break;
      case DEV_PLANEWAVE: {
        const double sinth = d.c[1], costh = d.c[2];
        const double ampl = d.p[0] * ramp_ampl(x, d.p[2], d.p[3], d.p[4]) * sin(k * (x * costh + y * sinth - ds.t) + d.p[6]);
        F[2] += ampl;
        F[3] += ampl * sinth;
        F[4] -= ampl * costh;
      } break;
      case DEV_GAUSSBEAM: {
        const double axis = d.p[1];
        const double xp = x - d.p[2] - axis * ds.t, yp = y - d.p[3], zp = z - d.p[4];
        const double E = d.a0 * exp(-xp * xp / (d.p[5] * d.p[5]) - yp * yp / (d.p[6] * d.p[6]) - zp * zp / (d.p[7] * d.p[7])) *
                         sin(k * xp);
        F[2] += E;
        F[4] -= axis * E;
      } break;
      default: break;
    }
  }
}

// ---- particles.cu
int launch_push_velocs(cudaStream_t st, PView mom, CPView fld, double dt, i64 np);
int launch_push_coords(cudaStream_t st, PView x, CPView mom, PView xc, double dt, i64 np);
int launch_gather(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView out, const GridGeom& g, i64 np);
int launch_gather_push(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView mom,
                       const GridGeom& g, double dt, const DeviceSet& und, i64 np);
int launch_devices(cudaStream_t st, CPView x, PView fld, const DeviceSet& und, i64 np);
int launch_planes_from_aos(cudaStream_t st, double* dst, const double* src, int ncomp, i64 cap, i64 np);
int launch_deposit_direct(cudaStream_t st, int env, int curr, CPView x, CPView mom, const double* w, cd* grid,
                          const GridGeom& g, const ChunkSpec& ch, i64 np, bool fold);
int launch_ghost_fold(cudaStream_t st, cd* grid, i64 nxn, i64 nrn, i64 nplanes);
// ---- particles_sorted.cu : kernels that exploit (chunk, x-tile, r-cell, x-cell)-sorted SoA particles
// run-accumulating deposit: a thread walks `run` consecutive particles, sums the contributions of
// those that share a cell in registers and issues one red.global.add per node value and cell run
int launch_deposit_runs(cudaStream_t st, int env, int curr, CPView x, CPView mom, const double* w, cd* grid,
                        const GridGeom& g, const ChunkSpec& ch, i64 np);
// binned deposit: a CTA takes kDepNPB consecutive particles of ONE x-chunk, counting-sorts them by
// cell in shared memory, then accumulates runs of same-cell particles in registers and issues one
// red.global.add per node value and run.  `SortedSpec` maps CTAs to chunks.
constexpr int kDepNPB = 1024;
struct SortedSpec {
  const int* ind;   // device IndInChunk(0:nchnk) (one chunk holding everything when not x-chunked)
  const int* cta;   // device prefix of CTA counts per chunk (0:nchnk), cta[c+1]-cta[c] = ceil(n_c / kDepNPB)
  int nchnk, ncta;
  i64 cs, tile_w;   // x cells per chunk and per re-binning tile (0: unknown / never re-binned)
  int cta_base;     // launch covers CTAs [cta_base, cta_base + ncta) of the table (partial launches)
};
int launch_deposit_binned(cudaStream_t st, int env, int curr, const double* x, const double* mom, const double* w,
                          i64 cap, cd* grid, const GridGeom& g, const ChunkSpec& ch, const SortedSpec& sp);
int launch_gather_push_binned(cudaStream_t st, int env, const double* x, const double* w, const cd* Fld, double* mom,
                              i64 cap, const GridGeom& g, double dt, const DeviceSet& und, const SortedSpec& sp);
int launch_gather_binned_out(cudaStream_t st, int env, const double* x, const double* w, const cd* Fld, double* fld_tot,
                             i64 cap, const GridGeom& g, const SortedSpec& sp);
// particles_fused.cu: gather + device + Boris push + position update + J / rho deposit in one kernel.
// `sp.cta` must be the CTA table for kFusedNPB particles per CTA.  push_dt = 2 pi q/m dt, dt = time step.
#ifndef CHB_FNPB
#define CHB_FNPB 512
#endif
constexpr int kFusedNPB = CHB_FNPB;  // particles per CTA of the fused kernel (tuning builds override it)
int launch_fused_particles(cudaStream_t st, int env, int space_charge, double* x, double* xh, double* mom,
                           const double* w, i64 cap, const cd* Fld, cd* J, cd* Rho, const GridGeom& g,
                           const ChunkSpec& ch, double push_dt, double dt, const DeviceSet& und, const SortedSpec& sp,
                           double leftX_J, double leftX_R,  // node 0 of the J / rho deposit grids (moving window)
                           int deposit_only = 0);           // 1: J and rho from the stored x_half / x / p (after a sort)
// gather + device + Boris push + push_coords in one streaming pass, no deposit (before a re-binning step's sort)
int launch_gather_push_coords(cudaStream_t st, int env, double* x, double* xh, double* mom, const double* w, i64 cap,
                              const cd* Fld, const GridGeom& g, double push_dt, double dt, const DeviceSet& und, i64 np,
                              int coords = 1);  // 0: gather + push only
void fused_profile_enable(int on);
void fused_profile_read(unsigned long long out[8]);
// field gather from a shared-memory tile of the EB grid + undulator + Boris push
int launch_gather_push_tiled(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView mom,
                             const GridGeom& g, double dt, const DeviceSet& und, i64 np);
int launch_chunk_bin(cudaStream_t st, CPView x, int8_t* chunked, int* counts, int* goout, double x0, double inv,
                     const double lims[4], int nchnk, i64 np);
int launch_permute(cudaStream_t st, PView dst, CPView src, const i64* idx, int ncomp, i64 np);
int launch_inside_flag(cudaStream_t st, CPView x, int* flag, const double lims[4], i64 np);
int launch_nonzero_flag(cudaStream_t st, const double* v, int* flag, i64 np);
int launch_compact_index(cudaStream_t st, const int* flag, const int* pos, int* out, i64 np);

// ---- diagnostics.cu : reductions behind moduls/diagnostics.py (NEXT-3)
int launch_field_energy(cudaStream_t st, const cd* EG, const double* fact, double* out, i64 nkx, i64 ncols);
int launch_beam_moments(cudaStream_t st, const double* x, const double* p, const double* w, i64 cap, i64 np, double* out16);
int launch_spectrum(cudaStream_t st, const double* p, const double* w, i64 cap, i64 np, int quantity, double lo, double hi,
                    int nbins, double* hist);
int launch_lineout(cudaStream_t st, const cd* A, cd* out, i64 nx, i64 offset);

// ---- gemm.cu : batched real GEMM on complex-interleaved data, FP64 tensor cores (DMMA)
//   C[M x N] = alpha * A[M x K] * B[K x N] + beta * C      (column-major, M = 2*Nx real rows)
// B must have been packed with gemm_pack_b (fragment-ordered tiles, zero padded).
struct GemmProblem {
  const double* A;   // lda in doubles
  const double* Bp;  // packed operator
  double* C;         // ldc in doubles
  double alpha, beta;
  const double* fact;  // optional real factor per complex element of C, (M/2 x N) column-major (epilogue, with the phase)
};
constexpr int kGemmMaxBatch = 48;
struct GemmBatch {
  GemmProblem p[kGemmMaxBatch];
  int count;
  // optional epilogue of every problem of the batch (beta must be 0): complex row i of C times
  // exp(i * phase_sign * phase_leftX * phase_kx[i]), then times fact -- the shiftX / shiftX_inv row scale of
  // fb_io.f90:52-57, :216-222 that otherwise costs a pass over C (launch_rowscale_phase)
  const double* phase_kx = nullptr;
  double phase_leftX = 0.0, phase_sign = 1.0;
};
i64 gemm_packed_size(i64 K, i64 N);  // doubles
int launch_gemm_pack_b(cudaStream_t st, double* Bp, const double* B, i64 K, i64 N, i64 ldb);
int launch_gemm(cudaStream_t st, const GemmBatch& batch, i64 M, i64 N, i64 K, i64 lda, i64 ldc);
void gemm_profile_enable(int on);
int gemm_profile_enabled();
void gemm_profile_read(double* ms, double* flops, long long* launches, int reset);

// ---- spectral.cu : elementwise kernels of the Fourier-Bessel PSATD update
int launch_rowscale_phase(cudaStream_t st, cd* a, const double* kx, double leftX, double sign, double scale,
                          const double* fact, i64 nkx, i64 ncols, i64 fact_cols);
// ncomp: number of (nxn, nrn, nm) component blocks at `eb` (6 = the whole EB grid; 3 = its E or B half)
int launch_eb_correction(cudaStream_t st, cd* eb, i64 nxn, i64 nrn, i64 nm, int env, int ncomp = 6);
int launch_maxwell_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const cd* gp, const void* C1,
                        const void* C2, int ncoef, int coef_complex, i64 P);
int launch_maxwell_init_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const cd* C1, const cd* C2, i64 P);
int launch_maxwell_static_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const double* w, const double* kx,
                               double beta0, i64 nkx, i64 P);
int launch_dt_stat(cudaStream_t st, cd* DT, const double* kx, double beta0, i64 nkx);
int launch_poiss_corr(cudaStream_t st, cd* J, const cd* gdj, const cd* gn, const cd* gp, double dt_inv,
                      const double* w2inv, i64 P);
int launch_poiss_corr_stat(cudaStream_t st, cd* J, const cd* gdj, const cd* gn, const cd* DT, const double* w2inv,
                           i64 nkx, i64 P);
int launch_field_drift(cudaStream_t st, cd* EG, const double* kx, double beta0, double dt, i64 nkx, i64 ncols);
int launch_mult_real(cudaStream_t st, cd* v, const double* A, i64 P, int ncomp);
int launch_add(cudaStream_t st, cd* v, const cd* A, i64 n);
int launch_window(cudaStream_t st, cd* a, const double* filtr, int modefilt, i64 nkx, i64 ncols, i64 nxfilt);
int launch_rowscale_cplx(cudaStream_t st, cd* a, const cd* s, i64 nkx, i64 ncols);

// linear combinations feeding the mode-coupling GEMMs (fb_math*.f90), see spectral.cu
//   out = ca * a + cb * b, with ca, cb in {0, +-1, +-i}; optional x-mirror-conjugate of the inputs
struct Unit { int re, im; };  // value = re + i*im, entries in {-1,0,1}
int launch_combine(cudaStream_t st, cd* out, const cd* a, Unit ca, const cd* b, Unit cb, int mirror, i64 nkx, i64 ncols);
//   out (+)= ca * a + cb * b   (accumulate when acc != 0)
int launch_axpby(cudaStream_t st, cd* out, const cd* a, Unit ca, const cd* b, Unit cb, int acc, i64 n);
int launch_fill(cudaStream_t st, double* y, i64 n, double vr, double vi, int is_complex);  // y[:] = value
int launch_add_f64(cudaStream_t st, double* y, const double* x, i64 n);                    // y += x
//   out (+)= i * kx * a   (sign = +-1)
int launch_ikx(cudaStream_t st, cd* out, const cd* a, const double* kx, double sign, int acc, i64 nkx, i64 ncols);
//   outm = i a - b, outp = i a + b in one pass
int launch_pm(cudaStream_t st, cd* outm, cd* outp, const cd* a, const cd* b, i64 n);
//   fused tails of fb_grad/fb_graddiv and fb_rot (i kx planes + the +-1/+-i recombination of the two GEMM results)
int launch_grad_tail(cudaStream_t st, cd* out, const cd* S, const cd* G1, const cd* G2, const double* kx, i64 nkx,
                     i64 Ps, i64 Pin, i64 nm);
//   fb_graddiv tail + poiss_corr in one pass: J_l += w2inv (gdj_l + dt_inv (gp_l - gn_l)), gdj never stored
int launch_grad_poiss_tail(cudaStream_t st, cd* J, const cd* S, const cd* G1, const cd* G2, const cd* gn, const cd* gp,
                           const double* kx, const double* w2inv, double dt_inv, i64 nkx, i64 n);
int launch_rot_tail(cudaStream_t st, cd* out2, cd* out3, const cd* v2, const cd* v3, const cd* GP, const cd* GM,
                    const double* kx, i64 nkx, i64 Ps, i64 Pv, i64 nm);


// ---- sr.cu : synchrotron-radiation spectra from stored tracks (SR.f90) and utils.f90 diagnostics helpers
// comp = 0: all three components (*_tot); 1..3: one component (*_comp)
int launch_sr_far(cudaStream_t st, double* spect, const double* coords, const double* mprv, const double* mnxt,
                  const double* wghts, int comp, double dt, const double* omega, const double* SinTh,
                  const double* CosTh, const double* SinPh, const double* CosPh, i64 nt, i64 np, i64 nom, i64 nth,
                  i64 nph);
// circ = 0: Cartesian screen, G1 = Xgrid(n1), G2s = Ygrid(n2); circ = 1: polar screen, G1 = Rgrid, G2s/G2c = Sin/CosPh
int launch_sr_near(cudaStream_t st, double* spect, const double* coords, const double* mom, const double* wghts,
                   int comp, double dt, const double* omega, const double* G1, const double* G2s, const double* G2c,
                   int circ, double z_scr, i64 nt, i64 np, i64 nom, i64 n1, i64 n2);
// phase(NO, nm): e^{i m theta} table; pwr(NO, nrn-1)
int launch_intens_profo(cudaStream_t st, double* pwr, const cd* fld, const cd* phase, int NO, i64 nxn, i64 nrn, i64 nm);
int launch_density_2x(cudaStream_t st, double* dens, const double* x, const double* y, const double* w,
                      const double grid4[4], int bx, int by, i64 n);

}  // namespace chb
