// sr.cu -- synchrotron-radiation spectra from stored particle tracks (reference f90/SR.f90, called from
// moduls/SR.py:165-215) and the two diagnostics helpers of f90/utils.f90 the driver reaches through `fimera`
// (intens_profO :18, DENSITY_2x :210).  SURVEY.md section 8(f) row 4.
//
// The SR sum is  spect(om, pixel) += |w_p| * sum_comp | sum_it A_comp(it, p, pixel) e^{i om phi(it, p, pixel)} |^2 :
// O(np nt nom npix) sincos evaluations, compute bound.  Mapping: one CTA per (pixel, omega tile, particle slice).
// Per tile of SR_TT time steps the CTA's threads first build the omega-independent step record (phase, the
// Nyquist-like guard value, amplitudes) in shared memory -- one thread per time step, the strictly ordered
// IEEE operations of the reference for the phase because it is multiplied by omega afterwards -- then every
// thread owns one omega (W = 32/64/128 per tile) and one of the 128/W interleaved time sub-streams and
// accumulates its partial complex integrals in registers.  The sub-streams of a particle are summed in a fixed
// order through shared memory, |.|^2 is taken, weighted and kept per thread; one atomicAdd per (omega, pixel)
// and CTA at the end.  Tracks are (3, nt, np) Fortran order, read coalesced along the time axis.
#include "kernels.cuh"

namespace chb {

namespace {

constexpr int SR_T = 128;   // threads per CTA
constexpr int SR_TT = 128;  // time steps per tile

struct SRArgs {
  double* spect;
  const double *coords, *m1, *m2, *wghts, *omega;
  const double *g1a, *g1b, *g2a, *g2b;  // far: SinTh, CosTh, SinPh, CosPh; near: G1, -, G2 (Y or SinPh), CosPh
  int comp, circ;
  double dt, z_scr;
  i64 nt, np, nom, n1, n2;
  int W;  // omegas per tile (power of two, <= SR_T)
};

__device__ __forceinline__ double mulr(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double addr(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double subr(double a, double b) { return __dsub_rn(a, b); }

// MODE 0: far field (SR.f90:18-254); MODE 1: near field on a Cartesian or polar screen (:256-642)
// NC: components accumulated (3 = *_tot, 1 = *_comp)
template <int MODE, int NC>
__global__ void __launch_bounds__(SR_T) sr_k(SRArgs a) {
  constexpr int NQ = MODE == 0 ? 2 + NC : 2 + 2 * NC;  // record: phase, guard, amplitudes
  __shared__ double rec[NQ][SR_TT];
  __shared__ double part[2 * NC][SR_T];
  const int tid = threadIdx.x;
  const int W = a.W, G = SR_T / W;
  const int lane_om = tid & (W - 1), grp = tid / W;
  const i64 px = blockIdx.x, i1 = px % a.n1, i2 = px / a.n1;
  const i64 iom = (i64)blockIdx.y * W + lane_om;
  const bool live = iom < a.nom;
  const double omg = live ? a.omega[iom] : 0.0;
  const double pi = 3.141592653589793, pi2 = 2.0 * pi, pi2_inv = 1.0 / pi2;  // 4 atan(1)

  double sin_th = 0, cos_th = 0, sin_ph = 0, cos_ph = 0, x_scr = 0, y_scr = 0;
  if (MODE == 0) {
    sin_th = a.g1a[i1]; cos_th = a.g1b[i1]; sin_ph = a.g2a[i2]; cos_ph = a.g2b[i2];
  } else if (a.circ) {
    x_scr = mulr(a.g1a[i1], a.g2b[i2]); y_scr = mulr(a.g1a[i1], a.g2a[i2]);  // SR.f90:494-495
  } else {
    x_scr = a.g1a[i1]; y_scr = a.g2a[i2];
  }
  const double dt_inv = 1.0 / a.dt;

  double acc = 0.0;
  for (i64 ip = blockIdx.z; ip < a.np; ip += gridDim.z) {
    const double wp = fabs(a.wghts[ip]);
    double ire[NC], iim[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) ire[k] = iim[k] = 0.0;
    double carry = 0.0;  // C3_prev / phase_prv, 0 before the first step (SR.f90:65,302)
    for (i64 t0 = 0; t0 < a.nt; t0 += SR_TT) {
      const int cnt = (int)min((i64)SR_TT, a.nt - t0);
      double ph = 0.0;
      if (tid < cnt) {
        const i64 o = 3 * (t0 + tid + a.nt * ip);
        const double x0 = a.coords[o], x1 = a.coords[o + 1], x2 = a.coords[o + 2];
        const double tnow = mulr((double)(t0 + tid + 1), a.dt);  // it*dt, it 1-based
        if constexpr (MODE == 0) {
          double vp[3] = {a.m1[o], a.m1[o + 1], a.m1[o + 2]}, vn[3] = {a.m2[o], a.m2[o + 1], a.m2[o + 2]};
          double g = 1.0 / sqrt(1.0 + (vp[0] * vp[0] + vp[1] * vp[1] + vp[2] * vp[2]));
          vp[0] *= g; vp[1] *= g; vp[2] *= g;
          g = 1.0 / sqrt(1.0 + (vn[0] * vn[0] + vn[1] * vn[1] + vn[2] * vn[2]));
          vn[0] *= g; vn[1] *= g; vn[2] *= g;
          double ac[3], v[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) { ac[k] = (vn[k] - vp[k]) * dt_inv; v[k] = 0.5 * (vn[k] + vp[k]); }
          const double C2 = 1.0 - (v[2] * sin_th * cos_ph + v[1] * sin_th * sin_ph + v[0] * cos_th);
          const double C2_inv = 1.0 / C2, C2_inv2 = C2_inv * C2_inv;
          const double C1 = ac[2] * sin_th * cos_ph + ac[1] * sin_th * sin_ph + ac[0] * cos_th;
          // C3 = 2 pi (it dt - n.x), the reference's operation order, each operation rounded once
          const double nx = addr(addr(mulr(mulr(x2, sin_th), cos_ph), mulr(mulr(x1, sin_th), sin_ph)), mulr(x0, cos_th));
          ph = mulr(pi2, subr(tnow, nx));
          const double dirs[3] = {cos_th, sin_ph * sin_th, cos_ph * sin_th};
          if constexpr (NC == 3) {
#pragma unroll
            for (int k = 0; k < 3; ++k) rec[2 + k][tid] = (C1 * (dirs[k] - v[k]) - C2 * ac[k]) * C2_inv2 * a.dt;
          } else {
            const int c = a.comp - 1;
            rec[2][tid] = (c >= 0 && c < 3) ? (C1 * (dirs[c] - v[c]) - C2 * ac[c]) * C2_inv2 * a.dt : 0.0;
          }
        } else {
          double n[3] = {subr(a.z_scr, x0), subr(y_scr, x1), subr(x_scr, x2)};
          const double R_0 = __dsqrt_rn(addr(addr(mulr(n[0], n[0]), mulr(n[1], n[1])), mulr(n[2], n[2])));
          const double R_inv = 1.0 / R_0;
          n[0] *= R_inv; n[1] *= R_inv; n[2] *= R_inv;
          const double u0 = a.m1[o], u1 = a.m1[o + 1], u2 = a.m1[o + 2];
          const double g = 1.0 / sqrt(1.0 + (u0 * u0 + u1 * u1 + u2 * u2));
          const double v[3] = {u0 * g, u1 * g, u2 * g};
          ph = mulr(pi2, addr(tnow, R_0));  // Im(arg_phase), SR.f90:316
          if constexpr (NC == 3) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              rec[2 + k][tid] = a.dt * R_inv * (v[k] - n[k]);               // Im(arg_amp1)
              rec[5 + k][tid] = a.dt * R_inv * R_inv * pi2_inv * n[k];      // arg_amp2
            }
          } else {
            const int c = a.comp - 1;
            rec[2][tid] = a.dt * R_inv * (v[c] - n[c]);
            rec[3][tid] = a.dt * R_inv * R_inv * pi2_inv * n[c];
          }
        }
        rec[0][tid] = ph;
      }
      __syncthreads();
      if (tid < cnt) rec[1][tid] = fabs(subr(ph, tid > 0 ? rec[0][tid - 1] : carry));  // dPhase / kotelnikov
      __syncthreads();
      carry = rec[0][cnt - 1];
      if (live) {
        for (int j = grp; j < cnt; j += G) {
          const double guard = mulr(rec[1][j], omg);  // far: omg*dPhase < pi ; near: kotelnikov*omg < 2 pi
          if (guard < (MODE == 0 ? pi : pi2)) {
            double s, c;
            sincos(mulr(rec[0][j], omg), &s, &c);
            if constexpr (MODE == 0) {
#pragma unroll
              for (int k = 0; k < NC; ++k) {
                const double amp = rec[2 + k][j];
                ire[k] += amp * c;
                iim[k] += amp * s;
              }
            } else {
#pragma unroll
              for (int k = 0; k < NC; ++k) {
                const double ai = rec[2 + k][j] * omg, ar = rec[2 + NC + k][j];
                ire[k] += ar * c - ai * s;
                iim[k] += ar * s + ai * c;
              }
            }
          }
        }
      }
      __syncthreads();
    }
    // sum the G time sub-streams of each omega in a fixed order, then |.|^2
    if (G > 1) {
#pragma unroll
      for (int k = 0; k < NC; ++k) { part[2 * k][tid] = ire[k]; part[2 * k + 1][tid] = iim[k]; }
      __syncthreads();
      if (grp == 0) {
#pragma unroll
        for (int k = 0; k < NC; ++k)
          for (int g = 1; g < G; ++g) { ire[k] += part[2 * k][g * W + lane_om]; iim[k] += part[2 * k + 1][g * W + lane_om]; }
      }
      __syncthreads();
    }
    if (grp == 0) {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < NC; ++k) sum += ire[k] * ire[k] + iim[k] * iim[k];
      acc += wp * sum;
    }
  }
  if (grp == 0 && live) atomicAdd(a.spect + iom + a.nom * px, acc);
}

template <int MODE>
int launch_sr(cudaStream_t st, SRArgs a) {
  if (a.nt < 1 || a.np < 1 || a.nom < 1 || a.n1 < 1 || a.n2 < 1) return 0;
  a.W = a.nom <= 32 ? 32 : (a.nom <= 64 ? 64 : 128);
  const i64 npix = a.n1 * a.n2, ntile = (a.nom + a.W - 1) / a.W;
  if (npix > 2147483647LL || ntile > 65535) { set_error("sr_calc: screen or frequency grid too large"); return 2; }
  // particle slices: enough CTAs for ~8 waves of the 148 SMs x 8 resident CTAs, never more than particles
  i64 slices = (8 * 148 * 8 + npix * ntile - 1) / (npix * ntile);
  slices = slices < 1 ? 1 : (slices > a.np ? a.np : slices);
  if (slices > 65535) slices = 65535;
  dim3 grid((unsigned)npix, (unsigned)ntile, (unsigned)slices);
  if (a.comp == 0) sr_k<MODE, 3><<<grid, SR_T, 0, st>>>(a);
  else sr_k<MODE, 1><<<grid, SR_T, 0, st>>>(a);
  CHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int launch_sr_far(cudaStream_t st, double* spect, const double* coords, const double* mprv, const double* mnxt,
                  const double* wghts, int comp, double dt, const double* omega, const double* SinTh,
                  const double* CosTh, const double* SinPh, const double* CosPh, i64 nt, i64 np, i64 nom, i64 nth,
                  i64 nph) {
  SRArgs a{spect, coords, mprv, mnxt, wghts, omega, SinTh, CosTh, SinPh, CosPh, comp, 0, dt, 0.0, nt, np, nom, nth, nph, 0};
  return launch_sr<0>(st, a);
}

int launch_sr_near(cudaStream_t st, double* spect, const double* coords, const double* mom, const double* wghts,
                   int comp, double dt, const double* omega, const double* G1, const double* G2s, const double* G2c,
                   int circ, double z_scr, i64 nt, i64 np, i64 nom, i64 n1, i64 n2) {
  SRArgs a{spect, coords, mom, nullptr, wghts, omega, G1, nullptr, G2s, G2c, comp, circ, dt, z_scr, nt, np, nom, n1, n2, 0};
  return launch_sr<1>(st, a);
}

// ---------------------------------------------------------------------------------------------------------
// utils.f90:18-57 intens_profO.  phase(NO, nm) complex: e^{i m theta_iO} built on the host by the reference's
// repeated multiplication.  One CTA per (radial node, tile of IP_TILE angles); threads stride over (x, comp).
namespace {
constexpr int IP_T = 128, IP_TILE = 8;

__global__ void __launch_bounds__(IP_T) intens_prof_k(double* pwr, const cd* fld, const cd* phase, int NO, i64 nxn,
                                                      i64 nrn, i64 nm) {
  extern __shared__ cd ph_s[];  // [IP_TILE][nm]
  __shared__ double red[IP_TILE][IP_T / 32];
  const i64 ir = blockIdx.x + 1;
  const int o0 = blockIdx.y * IP_TILE;
  for (int i = threadIdx.x; i < IP_TILE * nm; i += IP_T) {
    const int o = o0 + i / (int)nm;
    ph_s[i] = o < NO ? phase[o + (i64)NO * (i % nm)] : cmake(0.0, 0.0);
  }
  __syncthreads();
  double acc[IP_TILE];
#pragma unroll
  for (int o = 0; o < IP_TILE; ++o) acc[o] = 0.0;
  for (i64 q = threadIdx.x; q < 3 * nxn; q += IP_T) {
    const i64 l = q / nxn, ix = q - l * nxn;
    cd s[IP_TILE];
#pragma unroll
    for (int o = 0; o < IP_TILE; ++o) s[o] = cmake(0.0, 0.0);
    for (i64 m = 0; m < nm; ++m) {
      const cd f = fld[ix + nxn * (ir + nrn * (m + nm * l))];
#pragma unroll
      for (int o = 0; o < IP_TILE; ++o) s[o] = cadd(s[o], cmul(ph_s[o * nm + m], f));
    }
#pragma unroll
    for (int o = 0; o < IP_TILE; ++o) acc[o] += s[o].x * s[o].x + s[o].y * s[o].y;
  }
#pragma unroll
  for (int o = 0; o < IP_TILE; ++o) {
    double v = acc[o];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) red[o][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < IP_TILE && o0 + threadIdx.x < NO) {
    double v = 0.0;
    for (int w = 0; w < IP_T / 32; ++w) v += red[threadIdx.x][w];
    pwr[(o0 + threadIdx.x) + (i64)NO * (ir - 1)] = v;
  }
}

// utils.f90:210-272 DENSITY_2x: one thread per particle, 5x5 nodes, FP64 atomics; dens is (bx+5, by+5)
__global__ void density2x_k(double* dens, const double* x, const double* y, const double* w, double origx,
                            double origy, double dlt_xg, double dlt_yg, int bx, int by, i64 n) {
  const i64 jp = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (jp >= n) return;
  const double dxi = 1.0 / dlt_xg, dyi = 1.0 / dlt_yg;
  const double x_max = origx + dlt_xg * bx, y_max = origy + dlt_yg * by;
  const double xp = x[jp], yp = y[jp];
  if (!(xp >= origx && xp <= x_max && yp >= origy && yp <= y_max)) return;
  const i64 kx = (i64)floor((xp - origx) * dxi + 0.5), ky = (i64)floor((yp - origy) * dyi + 0.5);
  const double d[2] = {(xp - addr(mulr(dlt_xg, (double)kx), origx)) * dxi, (yp - addr(mulr(dlt_yg, (double)ky), origy)) * dyi};
  double S[2][5];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const double t = d[a], t3 = t * t * t;
    S[a][0] = S[a][4] = 0.0;
    S[a][1] = 0.25 - 0.5 * t + t3 / 3.0;
    S[a][2] = 0.5 - fabs(t3) / 3.0;
    S[a][3] = 0.25 + 0.5 * t - t3 / 3.0;
    if (t >= 0.0) S[a][4] = fabs(t3) / 3.0; else S[a][0] = fabs(t3) / 3.0;
  }
  const i64 sx = bx + 5, sy = by + 5;
  const double wp = w[jp];
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const i64 gx = kx + i, gy = ky + j;  // (kx-2+i) + 2 storage offset
      const double v = wp * S[0][i] * S[1][j];
      if (v != 0.0 && gx >= 0 && gx < sx && gy >= 0 && gy < sy) atomicAdd(dens + gx + sx * gy, v);
    }
}

__global__ void scale2_k(double* v, double s1, double s2, i64 n) {
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) v[i] = v[i] * s1 * s2;  // dens * dlt_xg_inv * dlt_yg_inv, left to right (utils.f90:270)
}
}  // namespace

int launch_intens_profo(cudaStream_t st, double* pwr, const cd* fld, const cd* phase, int NO, i64 nxn, i64 nrn,
                        i64 nm) {
  if (nrn < 2 || NO < 1) return 0;
  dim3 grid((unsigned)(nrn - 1), (unsigned)((NO + IP_TILE - 1) / IP_TILE));
  intens_prof_k<<<grid, IP_T, sizeof(cd) * IP_TILE * (size_t)nm, st>>>(pwr, fld, phase, NO, nxn, nrn, nm);
  CHB_LAUNCH_CHECK();
  return 0;
}

int launch_density_2x(cudaStream_t st, double* dens, const double* x, const double* y, const double* w,
                      const double grid4[4], int bx, int by, i64 n) {
  const double dlt_xg = (grid4[1] - grid4[0]) / bx, dlt_yg = (grid4[3] - grid4[2]) / by;
  const i64 nd = (i64)(bx + 5) * (by + 5);
  CHB_CUDA(cudaMemsetAsync(dens, 0, sizeof(double) * (size_t)nd, st));
  if (n > 0) {
    density2x_k<<<grid_for(n, 256), 256, 0, st>>>(dens, x, y, w, grid4[0], grid4[2], dlt_xg, dlt_yg, bx, by, n);
    CHB_LAUNCH_CHECK();
  }
  scale2_k<<<grid_for(nd, 256), 256, 0, st>>>(dens, 1.0 / dlt_xg, 1.0 / dlt_yg, nd);
  CHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace chb
