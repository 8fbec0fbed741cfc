// sr_utils_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE), second translation unit
//
// C++/OpenMP restatement of the reference's synchrotron-radiation post-processing (f90/SR.f90) and of the
// two diagnostics helpers of f90/utils.f90 the Python driver calls through `fimera` (intens_profO,
// DENSITY_2x).  Same role and same rules as chimera_oracle.cpp: only tests/, __graft_entry__.smoke() and the
// cpu legs of bench.py may use it.  PARITY UNPINNED for the same reason (no Fortran compiler, no golden vectors
// in the reference); pinned by the numpy restatement in oracle/np_ref.py and analytic known answers
// (tests/test_sr.py).
//
// Conventions as in chimera_oracle.cpp: Fortran-ordered arrays, complex = (re,im) doubles, dims = numpy shape.
// Arithmetic follows the Fortran expression order, including how complex-by-real products with a purely
// imaginary factor evaluate (the real part stays an exact 0, so exp(ii*omg*C3) = (cos, sin)(omg*C3)).
#include <cmath>
#include <cstdint>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long long i64;

namespace {

const double kPi = 4.0 * std::atan(1.0);

// per-thread partial spectrum added to the shared one at the end (SR.f90:51-52,125-132)
struct LocalSpect {
  std::vector<double> v;
  explicit LocalSpect(i64 n) : v((size_t)n, 0.0) {}
  void flush(double* spect) {
    for (size_t i = 0; i < v.size(); ++i) {
#pragma omp atomic
      spect[i] += v[i];
    }
  }
};

// SR.f90:18-137 (comp = 0, all three components) and :139-254 (comp = 1..3; any other value adds zero, :225-226)
int sr_far(double* spect, const double* coords, const double* mprv, const double* mnxt, const double* wghts,
           int comp, double dt, const double* omega, const double* SinTh, const double* CosTh, const double* SinPh,
           const double* CosPh, i64 nt, i64 np, i64 nom, i64 nth, i64 nph) {
  const double dt_inv = 1.0 / dt;
  const bool tot = comp == 0;
#pragma omp parallel
  {
    LocalSpect loc(nom * nth * nph);
    std::vector<double> ire((size_t)(3 * nom)), iim((size_t)(3 * nom));
#pragma omp for schedule(static)
    for (i64 ip = 0; ip < np; ++ip) {
      const double wp = std::fabs(wghts[ip]);
      for (i64 iph = 0; iph < nph; ++iph) {
        const double sin_ph = SinPh[iph], cos_ph = CosPh[iph];
        for (i64 ith = 0; ith < nth; ++ith) {
          const double sin_th = SinTh[ith], cos_th = CosTh[ith];
          std::fill(ire.begin(), ire.end(), 0.0);
          std::fill(iim.begin(), iim.end(), 0.0);
          double C3_prev = 0.0, dPhase = 0.0;
          for (i64 it = 1; it <= nt; ++it) {
            const double* x = coords + 3 * ((it - 1) + nt * ip);
            const double* a = mprv + 3 * ((it - 1) + nt * ip);
            const double* b = mnxt + 3 * ((it - 1) + nt * ip);
            double vp[3], vn[3], acc[3], v[3];
            double g = 1.0 / std::sqrt(1.0 + (a[0] * a[0] + a[1] * a[1] + a[2] * a[2]));
            for (int k = 0; k < 3; ++k) vp[k] = a[k] * g;
            g = 1.0 / std::sqrt(1.0 + (b[0] * b[0] + b[1] * b[1] + b[2] * b[2]));
            for (int k = 0; k < 3; ++k) vn[k] = b[k] * g;
            for (int k = 0; k < 3; ++k) { acc[k] = (vn[k] - vp[k]) * dt_inv; v[k] = 0.5 * (vn[k] + vp[k]); }
            const double C2 = 1.0 - (v[2] * sin_th * cos_ph + v[1] * sin_th * sin_ph + v[0] * cos_th);
            const double C2_inv = 1.0 / C2, C2_inv2 = C2_inv * C2_inv;
            const double C1 = acc[2] * sin_th * cos_ph + acc[1] * sin_th * sin_ph + acc[0] * cos_th;
            const double C3 = 2.0 * kPi * ((double)it * dt - (x[2] * sin_th * cos_ph + x[1] * sin_th * sin_ph + x[0] * cos_th));
            dPhase = std::fabs(C3 - C3_prev);
            C3_prev = C3;
            double C4[3];
            C4[0] = (C1 * (cos_th - v[0]) - C2 * acc[0]) * C2_inv2;
            C4[1] = (C1 * (sin_ph * sin_th - v[1]) - C2 * acc[1]) * C2_inv2;
            C4[2] = (C1 * (cos_ph * sin_th - v[2]) - C2 * acc[2]) * C2_inv2;
            for (i64 iom = 0; iom < nom; ++iom) {
              const double omg = omega[iom];
              if (omg * dPhase < kPi) {
                const double ph = omg * C3, c = std::cos(ph), s = std::sin(ph);
                if (tot) {
                  for (int k = 0; k < 3; ++k) {
                    const double amp = C4[k] * dt;
                    ire[3 * iom + k] += amp * c;
                    iim[3 * iom + k] += amp * s;
                  }
                } else {
                  const double amp = (comp >= 1 && comp <= 3 ? C4[comp - 1] : 0.0) * dt;
                  ire[iom] += amp * c;
                  iim[iom] += amp * s;
                }
              }
            }
          }
          double* out = loc.v.data() + nom * (ith + nth * iph);
          for (i64 iom = 0; iom < nom; ++iom) {
            double sum = 0.0;
            if (tot) {
              for (int k = 0; k < 3; ++k) {
                const double m = std::hypot(ire[3 * iom + k], iim[3 * iom + k]);  // ABS(complex)
                sum += m * m;
              }
            } else {
              const double m = std::hypot(ire[iom], iim[iom]);
              sum = m * m;
            }
            out[iom] += wp * sum;
          }
        }
      }
    }
    loc.flush(spect);
  }
  return 0;
}

// SR.f90:256-350 / :352-447 (Cartesian screen) and :449-544 / :546-642 (polar screen, circ = 1): near-field
// Lienard-Wiechert integral.  Pixel (i1, i2): Cartesian x_scr = G1[i1], y_scr = G2[i2]; polar
// x_scr = G1[i1]*CosPh[i2], y_scr = G1[i1]*SinPh[i2] with the phi loop outermost (:490-495).
int sr_near(double* spect, const double* coords, const double* mom, const double* wghts, int comp, double dt,
            const double* omega, const double* G1, const double* G2s, const double* G2c, int circ, double z_scr,
            i64 nt, i64 np, i64 nom, i64 n1, i64 n2) {
  const double pi2 = 2.0 * kPi, pi2_inv = 1.0 / pi2;
  const bool tot = comp == 0;
  if (!tot && (comp < 1 || comp > 3)) return 2;
#pragma omp parallel
  {
    LocalSpect loc(nom * n1 * n2);
    std::vector<double> ire((size_t)(3 * nom)), iim((size_t)(3 * nom));
#pragma omp for schedule(static)
    for (i64 ip = 0; ip < np; ++ip) {
      const double wp = std::fabs(wghts[ip]);
      for (i64 i1 = 0; i1 < n1; ++i1) {
        for (i64 i2 = 0; i2 < n2; ++i2) {
          const double x_scr = circ ? G1[i1] * G2c[i2] : G1[i1];
          const double y_scr = circ ? G1[i1] * G2s[i2] : G2s[i2];
          std::fill(ire.begin(), ire.end(), 0.0);
          std::fill(iim.begin(), iim.end(), 0.0);
          double phase_prv = 0.0;
          for (i64 it = 1; it <= nt; ++it) {
            const double* x = coords + 3 * ((it - 1) + nt * ip);
            const double* u = mom + 3 * ((it - 1) + nt * ip);
            double n[3] = {z_scr - x[0], y_scr - x[1], x_scr - x[2]};
            const double R_0 = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const double R_inv = 1.0 / R_0;
            for (int k = 0; k < 3; ++k) n[k] *= R_inv;
            const double g = 1.0 / std::sqrt(1.0 + (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]));
            const double v[3] = {u[0] * g, u[1] * g, u[2] * g};
            const double phase = pi2 * ((double)it * dt + R_0);  // Im(arg_phase); Re is exactly 0
            const double kot = std::fabs(phase - phase_prv);
            phase_prv = phase;
            double a1[3], a2[3];
            for (int k = 0; k < 3; ++k) {
              a1[k] = dt * R_inv * (v[k] - n[k]);            // Im(arg_amp1)
              a2[k] = dt * R_inv * R_inv * pi2_inv * n[k];   // arg_amp2 (real)
            }
            for (i64 iom = 0; iom < nom; ++iom) {
              const double omg = omega[iom];
              if (kot * omg < pi2) {
                const double ph = phase * omg, c = std::cos(ph), s = std::sin(ph);
                if (tot) {
                  for (int k = 0; k < 3; ++k) {
                    const double ai = a1[k] * omg;
                    ire[3 * iom + k] += a2[k] * c - ai * s;
                    iim[3 * iom + k] += a2[k] * s + ai * c;
                  }
                } else {
                  const double ai = a1[comp - 1] * omg, ar = a2[comp - 1];
                  ire[iom] += ar * c - ai * s;
                  iim[iom] += ar * s + ai * c;
                }
              }
            }
          }
          double* out = loc.v.data() + nom * (i1 + n1 * i2);
          for (i64 iom = 0; iom < nom; ++iom) {
            double sum = 0.0;
            if (tot) {
              for (int k = 0; k < 3; ++k) {
                const double m = std::hypot(ire[3 * iom + k], iim[3 * iom + k]);
                sum += m * m;
              }
            } else {
              const double m = std::hypot(ire[iom], iim[iom]);
              sum = m * m;
            }
            out[iom] += wp * sum;
          }
        }
      }
    }
    loc.flush(spect);
  }
  return 0;
}

}  // namespace

extern "C" {

int oracle_sr_calc_far_tot(double* spect, const double* coords, const double* momenta_prv, const double* momenta_nxt,
                           const double* wghts, double dt, const double* omega, const double* SinTh,
                           const double* CosTh, const double* SinPh, const double* CosPh, i64 nt, i64 np, i64 nom,
                           i64 nth, i64 nph) {
  return sr_far(spect, coords, momenta_prv, momenta_nxt, wghts, 0, dt, omega, SinTh, CosTh, SinPh, CosPh, nt, np, nom, nth, nph);
}
int oracle_sr_calc_far_comp(double* spect, const double* coords, const double* momenta_prv, const double* momenta_nxt,
                            const double* wghts, int comp, double dt, const double* omega, const double* SinTh,
                            const double* CosTh, const double* SinPh, const double* CosPh, i64 nt, i64 np, i64 nom,
                            i64 nth, i64 nph) {
  return sr_far(spect, coords, momenta_prv, momenta_nxt, wghts, comp >= 1 && comp <= 3 ? comp : -1, dt, omega, SinTh, CosTh, SinPh, CosPh, nt, np, nom, nth, nph);
}
int oracle_sr_calc_near_tot(double* spect, const double* coords, const double* momenta, const double* wghts, double dt,
                            const double* omega, const double* Xgrid, const double* Ygrid, double z_scr, i64 nt,
                            i64 np, i64 nom, i64 nx, i64 ny) {
  return sr_near(spect, coords, momenta, wghts, 0, dt, omega, Xgrid, Ygrid, nullptr, 0, z_scr, nt, np, nom, nx, ny);
}
int oracle_sr_calc_near_comp(double* spect, const double* coords, const double* momenta, const double* wghts, int comp,
                             double dt, const double* omega, const double* Xgrid, const double* Ygrid, double z_scr,
                             i64 nt, i64 np, i64 nom, i64 nx, i64 ny) {
  if (comp < 1 || comp > 3) return 2;
  return sr_near(spect, coords, momenta, wghts, comp, dt, omega, Xgrid, Ygrid, nullptr, 0, z_scr, nt, np, nom, nx, ny);
}
int oracle_sr_calc_nearcirc_tot(double* spect, const double* coords, const double* momenta, const double* wghts,
                                double dt, const double* omega, const double* Rgrid, const double* SinPh,
                                const double* CosPh, double z_scr, i64 nt, i64 np, i64 nom, i64 nr, i64 nph) {
  return sr_near(spect, coords, momenta, wghts, 0, dt, omega, Rgrid, SinPh, CosPh, 1, z_scr, nt, np, nom, nr, nph);
}
int oracle_sr_calc_nearcirc_comp(double* spect, const double* coords, const double* momenta, const double* wghts,
                                 int comp, double dt, const double* omega, const double* Rgrid, const double* SinPh,
                                 const double* CosPh, double z_scr, i64 nt, i64 np, i64 nom, i64 nr, i64 nph) {
  if (comp < 1 || comp > 3) return 2;
  return sr_near(spect, coords, momenta, wghts, comp, dt, omega, Rgrid, SinPh, CosPh, 1, z_scr, nt, np, nom, nr, nph);
}

// utils.f90:18-57  PWR_RO(NO, nr) = sum over x and the 3 components of |sum_m e^{i m theta_iO} Fld(ix, ir, m, l)|^2
// on NO azimuthal angles theta_iO = 2 pi (iO-1)/(NO-1) (phases by repeated multiplication, :31-45); radial node 0
// (the ghost) is skipped.  Fld(nxn, nrn, nm = 2 nkO + 1, 3) complex, envelope slot order -nkO..nkO.
int oracle_intens_profo(double* pwr, const double* Fld_, int NO, i64 nxn, i64 nrn, i64 nm) {
  if (NO < 1 || nm < 1 || (nm % 2) == 0) return 2;
  const i64 nko = (nm - 1) / 2, nr = nrn - 1;
  for (i64 i = 0; i < (i64)NO * nr; ++i) pwr[i] = 0.0;
  std::vector<double> osr((size_t)NO), osi((size_t)NO);
  osr[0] = 1.0; osi[0] = 0.0;
  const double pr = std::cos(2.0 * kPi / (double)(NO - 1)), pi_ = std::sin(2.0 * kPi / (double)(NO - 1));
  for (int i = 1; i < NO; ++i) {
    osr[i] = osr[i - 1] * pr - osi[i - 1] * pi_;
    osi[i] = osr[i - 1] * pi_ + osi[i - 1] * pr;
  }
  std::vector<double> fr((size_t)nm), fi((size_t)nm);
  for (int iO = 0; iO < NO; ++iO) {
    const double ppr = osr[iO], ppi = osi[iO];
    const double d = ppr * ppr + ppi * ppi;
    const double pmr = ppr / d, pmi = -ppi / d;  // 1./Os(iO)
    fr[nko] = 1.0; fi[nko] = 0.0;
    for (i64 k = 1; k <= nko; ++k) {
      fr[nko + k] = fr[nko + k - 1] * ppr - fi[nko + k - 1] * ppi;
      fi[nko + k] = fr[nko + k - 1] * ppi + fi[nko + k - 1] * ppr;
      fr[nko - k] = fr[nko - k + 1] * pmr - fi[nko - k + 1] * pmi;
      fi[nko - k] = fr[nko - k + 1] * pmi + fi[nko - k + 1] * pmr;
    }
    for (int l = 0; l < 3; ++l)
      for (i64 ir = 1; ir <= nr; ++ir)
        for (i64 ix = 0; ix < nxn; ++ix) {
          double sr = 0.0, si = 0.0;
          for (i64 m = 0; m < nm; ++m) {
            const double* f = Fld_ + 2 * (ix + nxn * (ir + nrn * (m + nm * l)));
            sr += fr[m] * f[0] - fi[m] * f[1];
            si += fr[m] * f[1] + fi[m] * f[0];
          }
          const double a = std::hypot(sr, si);
          pwr[iO + (i64)NO * (ir - 1)] += a * a;
        }
  }
  return 0;
}

// utils.f90:210-272  2-D weighted histogram with a 5-node (third-order) shape; dens(bins_x+5, bins_y+5) with
// Fortran bounds -2:bins+2; particles outside [orig, max] in either coordinate are skipped
int oracle_density_2x(const double* x, const double* y, const double* wght, const double* grid, int bins_x,
                      int bins_y, double* dens, i64 n_part) {
  if (bins_x < 1 || bins_y < 1) return 2;
  const i64 sx = bins_x + 5, sy = bins_y + 5;
  for (i64 i = 0; i < sx * sy; ++i) dens[i] = 0.0;
  const double origx = grid[0], origy = grid[2];
  const double dlt_xg = (grid[1] - grid[0]) / bins_x, dlt_yg = (grid[3] - grid[2]) / bins_y;
  const double dxi = 1.0 / dlt_xg, dyi = 1.0 / dlt_yg;
  const double x_max = origx + dlt_xg * bins_x, y_max = origy + dlt_yg * bins_y;
  for (i64 jp = 0; jp < n_part; ++jp) {
    if (!(x[jp] >= origx && x[jp] <= x_max && y[jp] >= origy && y[jp] <= y_max)) continue;
    double S[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};  // S0(-2:2, 1:2)
    const i64 kx = (i64)std::floor((x[jp] - origx) * dxi + 0.5), ky = (i64)std::floor((y[jp] - origy) * dyi + 0.5);
    // xg(j) = dlt_xg*REAL(j) + origx: REAL() is single precision, exact for the small integers involved
    const double dx = (x[jp] - (dlt_xg * (double)(float)kx + origx)) * dxi;
    const double dy = (y[jp] - (dlt_yg * (double)(float)ky + origy)) * dyi;
    const double d[2] = {dx, dy};
    for (int a = 0; a < 2; ++a) {
      const double t = d[a], t3 = t * t * t;
      S[a][1] = 0.25 - 0.5 * t + t3 / 3.0;
      S[a][2] = 0.5 - std::fabs(t3) / 3.0;
      S[a][3] = 0.25 + 0.5 * t - t3 / 3.0;
      if (t >= 0.0) S[a][4] = std::fabs(t3) / 3.0; else S[a][0] = std::fabs(t3) / 3.0;
    }
    for (int j = -2; j <= 2; ++j)
      for (int i = -2; i <= 2; ++i) {
        const i64 gx = kx + i + 2, gy = ky + j + 2;  // shift to 0-based storage
        if (gx < 0 || gx >= sx || gy < 0 || gy >= sy) continue;
        dens[gx + sx * gy] += wght[jp] * S[0][i + 2] * S[1][j + 2];
      }
  }
  for (i64 i = 0; i < sx * sy; ++i) dens[i] = dens[i] * dxi * dyi;
  return 0;
}

}  // extern "C"
