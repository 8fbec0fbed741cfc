// spectral.cu -- bandwidth-bound elementwise kernels of the Fourier-Bessel PSATD field update.
//
// Replaces reference f90/maxwell_solvers.f90 (PSATD advance, Poisson correction, field drift,
// omp_* helpers), the phase/normalisation passes of f90/fb_io.f90, eb_correction[_env] of
// f90/grid_deps*.f90 and the linear combinations around the mode-coupling contractions of
// f90/fb_math*.f90.  All arrays are complex128 (double2) in Fortran order with x fastest, so a
// thread block walks contiguous x and every access is a coalesced 128-bit load/store.
#include "common.cuh"
#include "kernels.cuh"

namespace chb {

namespace {
constexpr int TPB = 256;
__device__ __forceinline__ cd ldg(const cd* p) { return __ldg(p); }
__device__ __forceinline__ cd unit_mul(Unit u, cd a) {  // (re + i im) * a with re,im in {-1,0,1}
  return make_double2(u.re * a.x - u.im * a.y, u.re * a.y + u.im * a.x);
}
}  // namespace

// a(:,col) *= scale * exp(i*sign*kx*leftX) [* fact(:, col % fact_cols)]      (fb_io.f90:40,52 / :121,129)
__global__ void __launch_bounds__(TPB) rowscale_phase_k(cd* __restrict__ a, const double* __restrict__ kx,
                                                        double leftX, double sign, double scale,
                                                        const double* __restrict__ fact, i64 nkx, i64 ncols,
                                                        i64 fact_cols, i64 cols_per_block) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nkx) return;
  double sn, cs;
  sincos(leftX * __ldg(kx + i), &sn, &cs);
  const cd ph = cmake(scale * cs, scale * sign * sn);
  const i64 c0 = (i64)blockIdx.y * cols_per_block;
  const i64 c1 = (c0 + cols_per_block < ncols) ? c0 + cols_per_block : ncols;
  for (i64 c = c0; c < c1; c += 4) {  // 4 independent loads in flight per thread
    cd v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c + k < c1) {
        v[k] = cmul(a[i + nkx * (c + k)], ph);
        if (fact) v[k] = cscale(__ldg(fact + i + nkx * ((c + k) % fact_cols)), v[k]);
      }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c + k < c1) a[i + nkx * (c + k)] = v[k];
  }
}

int launch_rowscale_phase(cudaStream_t st, cd* a, const double* kx, double leftX, double sign, double scale,
                          const double* fact, i64 nkx, i64 ncols, i64 fact_cols) {
  if (nkx <= 0 || ncols <= 0) return 0;
  const i64 cpb = (nkx * ncols < (i64)1 << 22) ? 4 : 16;  // short grids: more CTAs, one group of loads per thread
  dim3 grid(grid_for(nkx, TPB), (unsigned)((ncols + cpb - 1) / cpb));
  rowscale_phase_k<<<grid, TPB, 0, st>>>(a, kx, leftX, sign, scale, fact, nkx, ncols, fact_cols ? fact_cols : 1, cpb);
  CHB_LAUNCH_CHECK();
  return 0;
}

// a(:,col) *= s(:)   complex row scale
__global__ void __launch_bounds__(TPB) rowscale_cplx_k(cd* __restrict__ a, const cd* __restrict__ s, i64 nkx, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  a[e] = cmul(a[e], ldg(s + e % nkx));
}
int launch_rowscale_cplx(cudaStream_t st, cd* a, const cd* s, i64 nkx, i64 ncols) {
  const i64 n = nkx * ncols;
  if (n <= 0) return 0;
  rowscale_cplx_k<<<grid_for(n, TPB), TPB, 0, st>>>(a, s, nkx, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// eb_correction (grid_deps.f90:219-266) / eb_correction_env (grid_deps_env.f90:240-283):
// normalise by 1/2pi (m=0, real) or 1/pi, then fill the ghost row from row 1 (copy or negate).
// One thread per (ix, 8 rows): the rows of a thread are independent loads in flight together, and the grid has
// nrn/8 times more CTAs than a thread-per-column walk (which ran a 300-deep dependent chain on 24 CTAs at Nr = 301).
constexpr int kEbRows = 8;
__global__ void __launch_bounds__(TPB) eb_correction_k(cd* __restrict__ eb, i64 nxn, i64 nrn, i64 nm, int env) {
  const i64 ix = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ix >= nxn) return;
  const i64 q = blockIdx.y;  // plane = slot + nm*l
  const i64 r0 = 1 + (i64)blockIdx.z * kEbRows;
  const int slot = (int)(q % nm);
  const double pi_inv = 1.0 / 3.14159265358979323846;
  const double f = (!env && slot == 0) ? 0.5 * pi_inv : pi_inv;
  const bool negate = env ? (nm > 1) : (slot > 0);  // Q6
  cd* pl = eb + nxn * nrn * q + ix;
  cd v[kEbRows];
#pragma unroll
  for (int k = 0; k < kEbRows; ++k)
    if (r0 + k < nrn) v[k] = cscale(f, pl[nxn * (r0 + k)]);
#pragma unroll
  for (int k = 0; k < kEbRows; ++k)
    if (r0 + k < nrn) pl[nxn * (r0 + k)] = v[k];
  if (r0 == 1 && nrn > 1) pl[0] = negate ? cneg(v[0]) : v[0];
}
int launch_eb_correction(cudaStream_t st, cd* eb, i64 nxn, i64 nrn, i64 nm, int env, int ncomp) {
  if (nxn <= 0 || nrn <= 0) return 0;
  const i64 nrb = nrn > 1 ? (nrn - 1 + kEbRows - 1) / kEbRows : 1;
  if (nrb > 65535) { set_error("eb_correction: Nr too large"); return 4; }
  dim3 grid(grid_for(nxn, TPB), (unsigned)(nm * ncomp), (unsigned)nrb);
  eb_correction_k<<<grid, TPB, 0, st>>>(eb, nxn, nrn, nm, env);
  CHB_LAUNCH_CHECK();
  return 0;
}

// PSATD advance (maxwell_solvers.f90:18-96).  NCOEF = 5 with real tables (space charge),
// NCOEF = 3 with real or complex tables.  One thread per spectral point, 3 components each:
// the coefficient tables are read once and reused for the three components.
template <int NCOEF, typename CT>
__global__ void __launch_bounds__(TPB) maxwell_push_k(cd* __restrict__ EG, const cd* __restrict__ J,
                                                      const cd* __restrict__ gn, const cd* __restrict__ gp,
                                                      const CT* __restrict__ C1, const CT* __restrict__ C2, i64 P) {
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  CT c1[NCOEF], c2[NCOEF];
#pragma unroll
  for (int k = 0; k < NCOEF; ++k) { c1[k] = __ldg(C1 + p + P * k); c2[k] = __ldg(C2 + p + P * k); }
  auto mul = [](CT c, cd v) -> cd {
    if constexpr (sizeof(CT) == sizeof(double)) return cscale(*(const double*)&c, v);
    else return cmul(*(const cd*)&c, v);
  };
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const cd e = EG[p + P * l], g = EG[p + P * (l + 3)], j = ldg(J + p + P * l);
    cd en = cadd(cadd(mul(c1[0], e), mul(c1[1], g)), mul(c1[2], j));
    cd gw = cadd(cadd(mul(c2[0], e), mul(c2[1], g)), mul(c2[2], j));
    if constexpr (NCOEF == 5) {
      const cd a = ldg(gn + p + P * l), b = ldg(gp + p + P * l);
      en = cadd(cadd(en, mul(c1[3], a)), mul(c1[4], b));
      gw = cadd(cadd(gw, mul(c2[3], a)), mul(c2[4], b));
    }
    EG[p + P * (l + 3)] = gw;
    EG[p + P * l] = en;
  }
}
int launch_maxwell_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const cd* gp, const void* C1,
                        const void* C2, int ncoef, int coef_complex, i64 P) {
  if (P <= 0) return 0;
  const unsigned nb = grid_for(P, TPB);
  if (ncoef == 5 && !coef_complex)
    maxwell_push_k<5, double><<<nb, TPB, 0, st>>>(EG, J, gn, gp, (const double*)C1, (const double*)C2, P);
  else if (ncoef == 3 && !coef_complex)
    maxwell_push_k<3, double><<<nb, TPB, 0, st>>>(EG, J, nullptr, nullptr, (const double*)C1, (const double*)C2, P);
  else if (ncoef == 3 && coef_complex)
    maxwell_push_k<3, cd><<<nb, TPB, 0, st>>>(EG, J, nullptr, nullptr, (const cd*)C1, (const cd*)C2, P);
  else { set_error("maxwell_push: unsupported coefficient layout"); return 5; }
  CHB_LAUNCH_CHECK();
  return 0;
}

// maxwell_solvers.f90:98-129
__global__ void __launch_bounds__(TPB) maxwell_init_push_k(cd* __restrict__ EG, const cd* __restrict__ J,
                                                           const cd* __restrict__ gn, const cd* __restrict__ C1,
                                                           const cd* __restrict__ C2, i64 P) {
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const cd c10 = ldg(C1 + p), c11 = ldg(C1 + p + P), c20 = ldg(C2 + p), c21 = ldg(C2 + p + P);
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const cd j = ldg(J + p + P * l), a = ldg(gn + p + P * l);
    EG[p + P * l] = cadd(cadd(EG[p + P * l], cmul(c10, j)), cmul(c11, a));
    EG[p + P * (l + 3)] = cadd(cadd(EG[p + P * (l + 3)], cmul(c20, j)), cmul(c21, a));
  }
}
int launch_maxwell_init_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const cd* C1, const cd* C2, i64 P) {
  if (P <= 0) return 0;
  maxwell_init_push_k<<<grid_for(P, TPB), TPB, 0, st>>>(EG, J, gn, C1, C2, P);
  CHB_LAUNCH_CHECK();
  return 0;
}

// solvers.py:333-358 maxwell_solver_stat with the coefficient tables formed on the fly from w, kx and beta0
// (CPSATD1 = [i kx beta0 / den, -1 / den], CPSATD2 = [w^2 / den, i kx beta0 / den], den = w^2 - (kx beta0)^2),
// then the update of maxwell_init_push (maxwell_solvers.f90:98-129)
__global__ void __launch_bounds__(TPB) maxwell_static_push_k(cd* __restrict__ EG, const cd* __restrict__ J,
                                                             const cd* __restrict__ gn, const double* __restrict__ w,
                                                             const double* __restrict__ kx, double beta0, i64 nkx, i64 P) {
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double wv = __ldg(w + p), kxb = beta0 * __ldg(kx + p % nkx);
  const double w2 = wv * wv, den = w2 - kxb * kxb;
  const cd c10 = cmake(0.0, kxb / den), c11 = cmake(-1.0 / den, 0.0), c20 = cmake(w2 / den, 0.0), c21 = c10;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const cd j = ldg(J + p + P * l), a = ldg(gn + p + P * l);
    EG[p + P * l] = cadd(cadd(EG[p + P * l], cmul(c10, j)), cmul(c11, a));
    EG[p + P * (l + 3)] = cadd(cadd(EG[p + P * (l + 3)], cmul(c20, j)), cmul(c21, a));
  }
}
int launch_maxwell_static_push(cudaStream_t st, cd* EG, const cd* J, const cd* gn, const double* w, const double* kx,
                               double beta0, i64 nkx, i64 P) {
  if (P <= 0) return 0;
  maxwell_static_push_k<<<grid_for(P, TPB), TPB, 0, st>>>(EG, J, gn, w, kx, beta0, nkx, P);
  CHB_LAUNCH_CHECK();
  return 0;
}

// DT = -i beta0 kx of solvers.py:376 poiss_corr_stat
__global__ void __launch_bounds__(TPB) dt_stat_k(cd* __restrict__ DT, const double* __restrict__ kx, double beta0, i64 nkx) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nkx) DT[i] = cmake(0.0, -(beta0 * __ldg(kx + i)));
}
int launch_dt_stat(cudaStream_t st, cd* DT, const double* kx, double beta0, i64 nkx) {
  dt_stat_k<<<grid_for(nkx, TPB), TPB, 0, st>>>(DT, kx, beta0, nkx);
  CHB_LAUNCH_CHECK();
  return 0;
}

// maxwell_solvers.f90:131-164
__global__ void __launch_bounds__(TPB) poiss_corr_k(cd* __restrict__ J, const cd* __restrict__ gdj,
                                                    const cd* __restrict__ gn, const cd* __restrict__ gp,
                                                    double dt_inv, const double* __restrict__ w2inv, i64 P) {
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double wi = __ldg(w2inv + p);
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const i64 e = p + P * l;
    const cd d = cscale(dt_inv, csub(ldg(gp + e), ldg(gn + e)));
    J[e] = cadd(J[e], cscale(wi, cadd(ldg(gdj + e), d)));
  }
}
int launch_poiss_corr(cudaStream_t st, cd* J, const cd* gdj, const cd* gn, const cd* gp, double dt_inv,
                      const double* w2inv, i64 P) {
  if (P <= 0) return 0;
  poiss_corr_k<<<grid_for(P, TPB), TPB, 0, st>>>(J, gdj, gn, gp, dt_inv, w2inv, P);
  CHB_LAUNCH_CHECK();
  return 0;
}

// maxwell_solvers.f90:166-197
__global__ void __launch_bounds__(TPB) poiss_corr_stat_k(cd* __restrict__ J, const cd* __restrict__ gdj,
                                                         const cd* __restrict__ gn, const cd* __restrict__ DT,
                                                         const double* __restrict__ w2inv, i64 nkx, i64 P) {
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double wi = __ldg(w2inv + p);
  const cd dt = ldg(DT + p % nkx);
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const i64 e = p + P * l;
    J[e] = cadd(J[e], cscale(wi, cadd(ldg(gdj + e), cmul(dt, ldg(gn + e)))));
  }
}
int launch_poiss_corr_stat(cudaStream_t st, cd* J, const cd* gdj, const cd* gn, const cd* DT, const double* w2inv,
                           i64 nkx, i64 P) {
  if (P <= 0) return 0;
  poiss_corr_stat_k<<<grid_for(P, TPB), TPB, 0, st>>>(J, gdj, gn, DT, w2inv, nkx, P);
  CHB_LAUNCH_CHECK();
  return 0;
}

// maxwell_solvers.f90:199-226 : EG *= exp(-i dt/2 beta0 kx)
__global__ void __launch_bounds__(TPB) field_drift_k(cd* __restrict__ EG, const double* __restrict__ kx, double fac,
                                                     i64 nkx, i64 ncols, i64 cols_per_block) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nkx) return;
  double sn, cs;
  sincos(fac * __ldg(kx + i), &sn, &cs);
  const cd ph = cmake(cs, sn);
  const i64 c0 = (i64)blockIdx.y * cols_per_block;
  const i64 c1 = (c0 + cols_per_block < ncols) ? c0 + cols_per_block : ncols;
  for (i64 c = c0; c < c1; ++c) EG[i + nkx * c] = cmul(EG[i + nkx * c], ph);
}
int launch_field_drift(cudaStream_t st, cd* EG, const double* kx, double beta0, double dt, i64 nkx, i64 ncols) {
  if (nkx <= 0 || ncols <= 0) return 0;
  const i64 cpb = (nkx * ncols < (i64)1 << 22) ? 4 : 16;  // short grids: more CTAs, one group of loads per thread
  dim3 grid(grid_for(nkx, TPB), (unsigned)((ncols + cpb - 1) / cpb));
  field_drift_k<<<grid, TPB, 0, st>>>(EG, kx, -0.5 * dt * beta0, nkx, ncols, cpb);
  CHB_LAUNCH_CHECK();
  return 0;
}

// omp_mult_vec / omp_mult_scl (maxwell_solvers.f90:228-272): v(:,:,:,l) *= A  (real A shared by comps)
__global__ void __launch_bounds__(TPB) mult_real_k(cd* __restrict__ v, const double* __restrict__ A, i64 P, int ncomp) {
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double a = __ldg(A + p);
  for (int l = 0; l < ncomp; ++l) v[p + P * l] = cscale(a, v[p + P * l]);
}
int launch_mult_real(cudaStream_t st, cd* v, const double* A, i64 P, int ncomp) {
  if (P <= 0) return 0;
  mult_real_k<<<grid_for(P, TPB), TPB, 0, st>>>(v, A, P, ncomp);
  CHB_LAUNCH_CHECK();
  return 0;
}

// omp_add_vec / omp_add_scl (maxwell_solvers.f90:274-318)
__global__ void __launch_bounds__(TPB) add_k(cd* __restrict__ v, const cd* __restrict__ A, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) v[e] = cadd(v[e], ldg(A + e));
}
int launch_add(cudaStream_t st, cd* v, const cd* A, i64 n) {
  if (n <= 0) return 0;
  add_k<<<grid_for(n, TPB), TPB, 0, st>>>(v, A, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// x-space window of fb_filtr (fb_io.f90:257-303), applied between the two FFTs
__global__ void __launch_bounds__(TPB) window_k(cd* __restrict__ a, const double* __restrict__ filtr, int modefilt,
                                                i64 nkx, i64 nxfilt, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const i64 i = e % nkx;
  double f = 1.0;
  if ((modefilt == 0 || modefilt == 2) && i < nxfilt) f *= __ldg(filtr + i);
  if ((modefilt == 1 || modefilt == 2) && i >= nkx - nxfilt) f *= __ldg(filtr + (nkx - 1 - i));
  if (f != 1.0) a[e] = cscale(f, a[e]);
}
int launch_window(cudaStream_t st, cd* a, const double* filtr, int modefilt, i64 nkx, i64 ncols, i64 nxfilt) {
  const i64 n = nkx * ncols;
  if (n <= 0) return 0;
  window_k<<<grid_for(n, TPB), TPB, 0, st>>>(a, filtr, modefilt, nkx, nxfilt, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// out = ca*a + cb*b (b may be null); mirror != 0 applies the real solver's m=0 coupling source
//   ext(i) = -conj(src((nkx - i) mod nkx))      (fb_math.f90:35-36)
// to both inputs *before* the unit factors.
__global__ void __launch_bounds__(TPB) combine_k(cd* __restrict__ out, const cd* __restrict__ a, Unit ca,
                                                 const cd* __restrict__ b, Unit cb, int mirror, i64 nkx, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  i64 src = e;
  if (mirror) {
    const i64 i = e % nkx;  // mirror = 1 + shift: partner row (nkx - i - shift) mod nkx (kx slabs: chimera_b200/sharding.py)
    i64 j = nkx - i - (mirror - 1);
    if (j >= nkx) j -= nkx;
    src = e - i + j;
  }
  cd va = ldg(a + src);
  if (mirror) va = cmake(-va.x, va.y);
  cd r = unit_mul(ca, va);
  if (b) {
    cd vb = ldg(b + src);
    if (mirror) vb = cmake(-vb.x, vb.y);
    r = cadd(r, unit_mul(cb, vb));
  }
  out[e] = r;
}
int launch_combine(cudaStream_t st, cd* out, const cd* a, Unit ca, const cd* b, Unit cb, int mirror, i64 nkx,
                   i64 ncols) {
  const i64 n = nkx * ncols;
  if (n <= 0) return 0;
  combine_k<<<grid_for(n, TPB), TPB, 0, st>>>(out, a, ca, b, cb, mirror, nkx, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(TPB) axpby_k(cd* __restrict__ out, const cd* __restrict__ a, Unit ca,
                                               const cd* __restrict__ b, Unit cb, int acc, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  cd r = unit_mul(ca, ldg(a + e));
  if (b) r = cadd(r, unit_mul(cb, ldg(b + e)));
  if (acc) r = cadd(r, out[e]);
  out[e] = r;
}
// whole-array statements of the driver on resident arrays: y[:] = value, y += x
__global__ void __launch_bounds__(TPB) fill_k(double* __restrict__ y, i64 n, double vr, double vi, int is_complex) {
  const i64 tot = is_complex ? 2 * n : n;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (i64)gridDim.x * blockDim.x)
    y[i] = (is_complex && (i & 1)) ? vi : vr;
}
int launch_fill(cudaStream_t st, double* y, i64 n, double vr, double vi, int is_complex) {
  if (n <= 0) return 0;
  fill_k<<<grid_for(n, TPB), TPB, 0, st>>>(y, n, vr, vi, is_complex);
  CHB_LAUNCH_CHECK();
  return 0;
}
__global__ void __launch_bounds__(TPB) add_f64_k(double* __restrict__ y, const double* __restrict__ x, i64 n) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) y[i] += x[i];
}
int launch_add_f64(cudaStream_t st, double* y, const double* x, i64 n) {
  if (n <= 0) return 0;
  add_f64_k<<<grid_for(n, TPB), TPB, 0, st>>>(y, x, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

int launch_axpby(cudaStream_t st, cd* out, const cd* a, Unit ca, const cd* b, Unit cb, int acc, i64 n) {
  if (n <= 0) return 0;
  axpby_k<<<grid_for(n, TPB), TPB, 0, st>>>(out, a, ca, b, cb, acc, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// out (+)= sign * i * kx * a
__global__ void __launch_bounds__(TPB) ikx_k(cd* __restrict__ out, const cd* __restrict__ a,
                                             const double* __restrict__ kx, double sign, int acc, i64 nkx, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double k = sign * __ldg(kx + e % nkx);
  const cd v = ldg(a + e);
  cd r = cmake(-k * v.y, k * v.x);
  if (acc) r = cadd(r, out[e]);
  out[e] = r;
}
int launch_ikx(cudaStream_t st, cd* out, const cd* a, const double* kx, double sign, int acc, i64 nkx, i64 ncols) {
  const i64 n = nkx * ncols;
  if (n <= 0) return 0;
  ikx_k<<<grid_for(n, TPB), TPB, 0, st>>>(out, a, kx, sign, acc, nkx, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// outm = i a - b ; outp = i a + b: the two mode-coupling sources of fb_div (a = v3, b = v2) and fb_rot
// (a = v2, b = v3) in one pass over the inputs (same expressions as two launch_combine calls)
__global__ void __launch_bounds__(TPB) pm_k(cd* __restrict__ outm, cd* __restrict__ outp, const cd* __restrict__ a,
                                            const cd* __restrict__ b, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const Unit Ui{0, 1}, U1{1, 0}, Um1{-1, 0};
  const cd ia = unit_mul(Ui, ldg(a + e)), vb = ldg(b + e);
  outm[e] = cadd(ia, unit_mul(Um1, vb));
  outp[e] = cadd(ia, unit_mul(U1, vb));
}
int launch_pm(cudaStream_t st, cd* outm, cd* outp, const cd* a, const cd* b, i64 n) {
  if (n <= 0) return 0;
  pm_k<<<grid_for(n, TPB), TPB, 0, st>>>(outm, outp, a, b, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// tail of fb_grad / fb_graddiv (fb_math.f90:133-147, :262-291): out1 = i kx S, out2 = -G1 + G2, out3 = i G1 + i G2.
// out components are (Ps, nm) blocks; S planes are Pin apart (Ps <= Pin leading elements used)
__global__ void __launch_bounds__(TPB) grad_tail_k(cd* __restrict__ out, const cd* __restrict__ S,
                                                   const cd* __restrict__ G1, const cd* __restrict__ G2,
                                                   const double* __restrict__ kx, i64 nkx, i64 Ps, i64 Pin, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const Unit Ui{0, 1}, U1{1, 0}, Um1{-1, 0};
  const i64 pl = e / Ps, w = e - pl * Ps;
  const double k = __ldg(kx + w % nkx);
  const cd s = ldg(S + pl * Pin + w), g1 = ldg(G1 + e), g2 = ldg(G2 + e);
  out[e] = cmake(-k * s.y, k * s.x);
  out[e + n] = cadd(unit_mul(Um1, g1), unit_mul(U1, g2));
  out[e + 2 * n] = cadd(unit_mul(Ui, g1), unit_mul(Ui, g2));
}
int launch_grad_tail(cudaStream_t st, cd* out, const cd* S, const cd* G1, const cd* G2, const double* kx, i64 nkx,
                     i64 Ps, i64 Pin, i64 nm) {
  const i64 n = Ps * nm;
  if (n <= 0) return 0;
  grad_tail_k<<<grid_for(n, TPB), TPB, 0, st>>>(out, S, G1, G2, kx, nkx, Ps, Pin, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// tail of fb_rot (fb_math.f90:60-92): out2 = -i kx v3 + (i GP + i GM), out3 = +i kx v2 + (-GP + GM)
__global__ void __launch_bounds__(TPB) rot_tail_k(cd* __restrict__ out2, cd* __restrict__ out3,
                                                  const cd* __restrict__ v2, const cd* __restrict__ v3,
                                                  const cd* __restrict__ GP, const cd* __restrict__ GM,
                                                  const double* __restrict__ kx, i64 nkx, i64 Ps, i64 Pv, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const Unit Ui{0, 1}, U1{1, 0}, Um1{-1, 0};
  const i64 pl = e / Ps, w = e - pl * Ps;
  const double kp = __ldg(kx + w % nkx), km = -1.0 * kp;
  const cd a2 = ldg(v2 + pl * Pv + w), a3 = ldg(v3 + pl * Pv + w), gp = ldg(GP + e), gm = ldg(GM + e);
  out2[e] = cadd(cadd(unit_mul(Ui, gp), unit_mul(Ui, gm)), cmake(-km * a3.y, km * a3.x));
  out3[e] = cadd(cadd(unit_mul(Um1, gp), unit_mul(U1, gm)), cmake(-kp * a2.y, kp * a2.x));
}
int launch_rot_tail(cudaStream_t st, cd* out2, cd* out3, const cd* v2, const cd* v3, const cd* GP, const cd* GM,
                    const double* kx, i64 nkx, i64 Ps, i64 Pv, i64 nm) {
  const i64 n = Ps * nm;
  if (n <= 0) return 0;
  rot_tail_k<<<grid_for(n, TPB), TPB, 0, st>>>(out2, out3, v2, v3, GP, GM, kx, nkx, Ps, Pv, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

// one Poisson-correction iteration's tail (solvers.py:317-326 = fb_graddiv second half + poiss_corr,
// maxwell_solvers.f90:131-164) without materialising grad(div J):
//   gdj = (i kx S, -G1 + G2, i G1 + i G2);  J_l += w2inv * (gdj_l + dt_inv * (gp_l - gn_l))
__global__ void __launch_bounds__(TPB) grad_poiss_tail_k(cd* __restrict__ J, const cd* __restrict__ S,
                                                         const cd* __restrict__ G1, const cd* __restrict__ G2,
                                                         const cd* __restrict__ gn, const cd* __restrict__ gp,
                                                         const double* __restrict__ kx,
                                                         const double* __restrict__ w2inv, double dt_inv, i64 nkx,
                                                         i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const Unit Ui{0, 1}, U1{1, 0}, Um1{-1, 0};
  const double k = __ldg(kx + e % nkx), wi = __ldg(w2inv + e);
  const cd s = ldg(S + e), g1 = ldg(G1 + e), g2 = ldg(G2 + e);
  cd gdj[3];
  gdj[0] = cmake(-k * s.y, k * s.x);
  gdj[1] = cadd(unit_mul(Um1, g1), unit_mul(U1, g2));
  gdj[2] = cadd(unit_mul(Ui, g1), unit_mul(Ui, g2));
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const i64 q = e + n * l;
    const cd d = cscale(dt_inv, csub(ldg(gp + q), ldg(gn + q)));
    J[q] = cadd(J[q], cscale(wi, cadd(gdj[l], d)));
  }
}
int launch_grad_poiss_tail(cudaStream_t st, cd* J, const cd* S, const cd* G1, const cd* G2, const cd* gn, const cd* gp,
                           const double* kx, const double* w2inv, double dt_inv, i64 nkx, i64 n) {
  if (n <= 0) return 0;
  grad_poiss_tail_k<<<grid_for(n, TPB), TPB, 0, st>>>(J, S, G1, G2, gn, gp, kx, w2inv, dt_inv, nkx, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace chb
