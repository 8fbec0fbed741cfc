// kernels.cuh -- internal (device-pointer) launch API shared by the C-ABI host layer and the
// resident engine.  Everything here takes a stream and device pointers; no allocation.
#pragma once
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

// analytic planar undulator (devices.f90:162-203)
struct UndulParams {
  int on;
  double a0, lambda, X0, Lx;
};

__device__ __forceinline__ void undul_field(const UndulParams& u, double x, double y, double F[6]) {
  const double ku = 2.0 * 3.14159265358979323846 / u.lambda;
  double ampl;
  if (x <= u.X0 || x >= u.X0 + u.Lx) ampl = 0.0;
  else if (x > u.X0 && x < u.X0 + u.lambda) ampl = (x - u.X0) / u.lambda;
  else if (x > u.X0 + u.Lx - u.lambda && x < u.X0 + u.Lx) ampl = (u.X0 + u.Lx - x) / u.lambda;
  else ampl = 1.0;
  ampl *= u.a0;
  F[4] += ampl * sin(ku * (x - u.X0)) * cosh(ku * y);
  F[3] += ampl * cos(ku * (x - u.X0)) * sinh(ku * y);
}

// ---- particles.cu
int launch_push_velocs(cudaStream_t st, PView mom, CPView fld, double dt, i64 np);
int launch_push_coords(cudaStream_t st, PView x, CPView mom, PView xc, double dt, i64 np);
int launch_gather(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView out, const GridGeom& g, i64 np);
int launch_gather_push(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView mom,
                       const GridGeom& g, double dt, const UndulParams& und, i64 np);
int launch_undul(cudaStream_t st, CPView x, PView fld, const UndulParams& und, i64 np);
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
                              i64 cap, const GridGeom& g, double dt, const UndulParams& und, const SortedSpec& sp);
// particles_fused.cu: gather + device + Boris push + position update + J / rho deposit in one kernel.
// `sp.cta` must be the CTA table for kFusedNPB particles per CTA.  push_dt = 2 pi q/m dt, dt = time step.
constexpr int kFusedNPB = 512;
int launch_fused_particles(cudaStream_t st, int env, int space_charge, double* x, double* xh, double* mom,
                           const double* w, i64 cap, const cd* Fld, cd* J, cd* Rho, const GridGeom& g,
                           const ChunkSpec& ch, double push_dt, double dt, const UndulParams& und, const SortedSpec& sp);
void fused_profile_enable(int on);
void fused_profile_read(unsigned long long out[8]);
// field gather from a shared-memory tile of the EB grid + undulator + Boris push
int launch_gather_push_tiled(cudaStream_t st, int env, CPView x, const double* w, const cd* Fld, PView mom,
                             const GridGeom& g, double dt, const UndulParams& und, i64 np);
int launch_chunk_bin(cudaStream_t st, CPView x, int8_t* chunked, int* counts, int* goout, double x0, double inv,
                     const double lims[4], int nchnk, i64 np);
int launch_permute(cudaStream_t st, PView dst, CPView src, const i64* idx, int ncomp, i64 np);
int launch_inside_flag(cudaStream_t st, CPView x, int* flag, const double lims[4], i64 np);
int launch_nonzero_flag(cudaStream_t st, const double* v, int* flag, i64 np);
int launch_compact_index(cudaStream_t st, const int* flag, const int* pos, int* out, i64 np);

// ---- gemm.cu : batched real GEMM on complex-interleaved data, FP64 tensor cores (DMMA)
//   C[M x N] = alpha * A[M x K] * B[K x N] + beta * C      (column-major, M = 2*Nx real rows)
// B must have been packed with gemm_pack_b (fragment-ordered tiles, zero padded).
struct GemmProblem {
  const double* A;   // lda in doubles
  const double* Bp;  // packed operator
  double* C;         // ldc in doubles
  double alpha, beta;
};
constexpr int kGemmMaxBatch = 48;
struct GemmBatch {
  GemmProblem p[kGemmMaxBatch];
  int count;
};
i64 gemm_packed_size(i64 K, i64 N);  // doubles
int launch_gemm_pack_b(cudaStream_t st, double* Bp, const double* B, i64 K, i64 N, i64 ldb);
int launch_gemm(cudaStream_t st, const GemmBatch& batch, i64 M, i64 N, i64 K, i64 lda, i64 ldc);
void gemm_profile_enable(int on);
void gemm_profile_read(double* ms, double* flops, long long* launches, int reset);

// ---- spectral.cu : elementwise kernels of the Fourier-Bessel PSATD update
int launch_rowscale_phase(cudaStream_t st, cd* a, const double* kx, double leftX, double sign, double scale,
                          const double* fact, i64 nkx, i64 ncols, i64 fact_cols);
int launch_eb_correction(cudaStream_t st, cd* eb, i64 nxn, i64 nrn, i64 nm, int env);
int launch_maxwell_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const cd* gp, const void* C1,
                        const void* C2, int ncoef, int coef_complex, i64 P);
int launch_maxwell_init_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const cd* C1, const cd* C2, i64 P);
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
//   out (+)= i * kx * a   (sign = +-1)
int launch_ikx(cudaStream_t st, cd* out, const cd* a, const double* kx, double sign, int acc, i64 nkx, i64 ncols);

}  // namespace chb
