// chimera_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++/OpenMP restatement of the reference's Fortran hot path
// (hightower8083/chimera, f90/*.f90), loop for loop, used ONLY as the checker in
// tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
// bench.py.  Nothing under chimera_b200/ may import, link or call this file.
//
// PARITY UNPINNED: the reference ships no golden vectors / known-answer tests for this
// path (doc/tests/*.py assert exit code only) and its Fortran cannot be compiled in the
// build container (no gfortran, no FFTW3).  This restatement is therefore pinned only by
// (a) analytic known-answer tests in tests/test_oracle_*.py, (b) an independent numpy
// restatement in oracle/np_ref.py, and (c) running the reference's *unmodified Python
// driver* on top of it (tools/gen_golden.py).  See DESIGN.md "Oracle".
//
// Conventions
//  * every array is Fortran (column-major) order, complex = interleaved (re,im) doubles;
//  * dims are passed as the numpy *shape* extents (nxn = number of x nodes, nrn = number of
//    r nodes incl. the r=-dr/2 ghost, nm = number of azimuthal-mode slots), not as the
//    Fortran upper bounds (nx = nxn-1 ...);
//  * every function cites the reference file:line it follows.
//  * FFTW3 (third-party, unpinned in the reference: README.md:20) is replaced by the
//    mixed-radix FFT below: unnormalised forward exp(-i..) / backward exp(+i..), the
//    published FFTW convention used at fb_io.f90:37,51,118,132.
//
// Out-of-range behaviour: the Fortran has no bounds checks.  A particle
// whose cell lies outside the grid is undefined behaviour there; here (and in the CUDA
// path) such a contribution is dropped.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long long i64;

namespace {

struct cd {
  double re, im;
  cd() : re(0), im(0) {}
  cd(double r, double i = 0.0) : re(r), im(i) {}
};
inline cd operator+(cd a, cd b) { return cd(a.re + b.re, a.im + b.im); }
inline cd operator-(cd a, cd b) { return cd(a.re - b.re, a.im - b.im); }
inline cd operator-(cd a) { return cd(-a.re, -a.im); }
inline cd operator*(cd a, cd b) { return cd(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline cd operator*(double s, cd a) { return cd(s * a.re, s * a.im); }
inline cd operator*(cd a, double s) { return cd(s * a.re, s * a.im); }
inline cd& operator+=(cd& a, cd b) { a.re += b.re; a.im += b.im; return a; }
inline cd& operator-=(cd& a, cd b) { a.re -= b.re; a.im -= b.im; return a; }
inline cd conj(cd a) { return cd(a.re, -a.im); }
inline cd mul_i(cd a) { return cd(-a.im, a.re); }  // i*a
inline cd cdiv(cd a, cd b) {
  double d = b.re * b.re + b.im * b.im;
  return cd((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}

// ---------------------------------------------------------------------------------------
// FFT: recursive mixed-radix decimation-in-time with an exact-angle twiddle table.
// sign = -1: forward (FFTW_FORWARD), sign = +1: backward (FFTW_BACKWARD), unnormalised.
// ---------------------------------------------------------------------------------------
struct FFTPlan {
  int n;
  std::vector<cd> tw;  // tw[k] = exp(-2 pi i k / n)
  std::vector<int> factors;
  explicit FFTPlan(int n_) : n(n_), tw(n_) {
    const long double tp = 6.283185307179586476925286766559L;
    for (int k = 0; k < n; ++k) {
      long double a = tp * (long double)k / (long double)n;
      tw[k] = cd((double)cosl(a), (double)(-sinl(a)));
    }
    int m = n;
    while (m % 4 == 0) { factors.push_back(4); m /= 4; }
    while (m % 2 == 0) { factors.push_back(2); m /= 2; }
    for (int p = 3; (i64)p * p <= m; p += 2)
      while (m % p == 0) { factors.push_back(p); m /= p; }
    if (m > 1) factors.push_back(m);
  }
  inline cd w(i64 k, int sign) const {  // exp(sign * 2 pi i k / n)
    cd t = tw[(int)(k % n)];
    return sign < 0 ? t : conj(t);
  }
  // out[0..len) = DFT of in[0], in[stride], ... ; len = n / tstride
  void rec(const cd* in, cd* out, int len, int stride, int fidx, int sign, cd* scratch) const {
    if (len == 1) { out[0] = in[0]; return; }
    const int p = factors[fidx];
    const int m = len / p;
    for (int q = 0; q < p; ++q) rec(in + (i64)q * stride, out + (i64)q * m, m, stride * p, fidx + 1, sign, scratch);
    // combine: X[k + m*j] = sum_q W_len^{q (k + m j)} Y_q[k]
    const int tstep = n / len;  // W_len = W_n^tstep
    if (p == 2) {
      for (int k = 0; k < m; ++k) {
        cd a = out[k], b = out[m + k] * w((i64)k * tstep, sign);
        out[k] = a + b;
        out[m + k] = a - b;
      }
    } else if (p == 4) {
      for (int k = 0; k < m; ++k) {
        cd a = out[k];
        cd b = out[m + k] * w((i64)k * tstep, sign);
        cd c = out[2 * m + k] * w((i64)2 * k * tstep, sign);
        cd d = out[3 * m + k] * w((i64)3 * k * tstep, sign);
        cd s0 = a + c, s1 = a - c, s2 = b + d, s3 = b - d;
        // W_4 = exp(sign * i pi/2) = sign * i
        cd is3 = sign < 0 ? cd(s3.im, -s3.re) : cd(-s3.im, s3.re);
        out[k] = s0 + s2;
        out[m + k] = s1 + is3;
        out[2 * m + k] = s0 - s2;
        out[3 * m + k] = s1 - is3;
      }
    } else {
      cd* tmp = scratch + p;
      for (int k = 0; k < m; ++k) {
        for (int q = 0; q < p; ++q) tmp[q] = out[(i64)q * m + k] * w((i64)q * k * tstep, sign);
        for (int j = 0; j < p; ++j) {
          cd acc = tmp[0];
          for (int q = 1; q < p; ++q) acc += tmp[q] * w((i64)((i64)q * j % p) * m * tstep, sign);
          scratch[j] = acc;
        }
        for (int j = 0; j < p; ++j) out[(i64)j * m + k] = scratch[j];
      }
    }
  }
  void exec(const cd* in, cd* out, int sign) const {
    std::vector<cd> scratch(2 * (size_t)n + 64);
    if (in == out) {
      std::vector<cd> tmp(in, in + n);
      rec(tmp.data(), out, n, 1, 0, sign, scratch.data());
    } else {
      rec(in, out, n, 1, 0, sign, scratch.data());
    }
  }
};

inline i64 ifloor(double v) { return (i64)std::floor(v); }

}  // namespace

extern "C" {

// test hook: 1-D complex DFT, unnormalised, sign=-1 forward / +1 backward
int oracle_fft1d(double* out, const double* in, i64 n, int sign) {
  FFTPlan plan((int)n);
  plan.exec((const cd*)in, (cd*)out, sign);
  return 0;
}

// =======================================================================================
// particle_tools.f90
// =======================================================================================

// particle_tools.f90:18-56  relativistic Boris push
int oracle_push_velocs(double* momenta, const double* Fld, double dt, i64 np) {
  const double dt_2 = 0.5 * dt;
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) {
    const double* f = Fld + 6 * ip;
    double* p = momenta + 3 * ip;
    double um[3], t[3], s[3], u0[3], up[3];
    for (int i = 0; i < 3; ++i) um[i] = p[i] + dt_2 * f[i];
    double gamma = std::sqrt(1.0 + (um[0] * um[0] + um[1] * um[1] + um[2] * um[2]));
    for (int i = 0; i < 3; ++i) t[i] = dt_2 * f[3 + i] / gamma;
    double t2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
    for (int i = 0; i < 3; ++i) s[i] = 2 * t[i] / (1 + t2);
    u0[0] = um[0] + um[1] * t[2] - um[2] * t[1];
    u0[1] = um[1] - um[0] * t[2] + um[2] * t[0];
    u0[2] = um[2] + um[0] * t[1] - um[1] * t[0];
    up[0] = um[0] + u0[1] * s[2] - u0[2] * s[1];
    up[1] = um[1] - u0[0] * s[2] + u0[2] * s[0];
    up[2] = um[2] + u0[0] * s[1] - u0[1] * s[0];
    for (int i = 0; i < 3; ++i) p[i] = up[i] + dt_2 * f[i];
  }
  return 0;
}

// particle_tools.f90:58-82  leap-frog position update + centred position
int oracle_push_coords(double* coord, const double* momenta, double* coord_cntr, double dt, i64 np) {
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) {
    double* x = coord + 3 * ip;
    const double* p = momenta + 3 * ip;
    double* xc = coord_cntr + 3 * ip;
    double dt_gp = dt / std::sqrt(1.0 + (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
    for (int i = 0; i < 3; ++i) {
      double x0 = x[i];
      double x1 = x0 + p[i] * dt_gp;
      x[i] = x1;
      xc[i] = 0.5 * (x0 + x1);
    }
  }
  return 0;
}

// particle_tools.f90:84-128  fill cells with a fixed pattern of particles
int oracle_genparts(double* coord, int* indPart, const double* Xgrid, const double* Rgrid,
                    const double* RandPackO, const double* PackX, const double* PackR,
                    const double* PackO, i64 np, i64 nx, i64 nr, i64 ppc) {
  const double pi = 4.0 * std::atan(1.0);
  std::memset(coord, 0, sizeof(double) * 4 * np);
  const double dr_2 = 0.5 * (Rgrid[1] - Rgrid[0]);
  i64 n = 0;
  const cd* po = (const cd*)PackO;
  for (i64 ir = 0; ir < nr - 1; ++ir)
    for (i64 ix = 0; ix < nx - 1; ++ix) {
      double r0 = Rgrid[ir] + dr_2, r1 = Rgrid[ir + 1] + dr_2;
      double x0 = Xgrid[ix], x1 = Xgrid[ix + 1];
      double a = 2.0 * pi * RandPackO[ix + nx * ir];
      cd osh(std::cos(a), std::sin(a));
      for (i64 ip = 0; ip < ppc; ++ip) {
        double xc = x0 + (x1 - x0) * PackX[ip];
        double rc = r0 + (r1 - r0) * PackR[ip];
        if (rc <= 0) continue;
        cd oc = po[ip] * osh;
        if (n >= np) return 1;
        coord[4 * n + 0] = xc;
        coord[4 * n + 1] = rc * oc.im;
        coord[4 * n + 2] = rc * oc.re;
        coord[4 * n + 3] = rc;
        ++n;
      }
    }
  *indPart = (int)n;
  return 0;
}

// particle_tools.f90:130-153
int oracle_sortpartsout(int* indx2stay, int* num2stay, const double* coord, const double* lims, i64 np) {
  int n = 0;
  for (i64 ip = 0; ip < np; ++ip) indx2stay[ip] = 0;
  for (i64 ip = 0; ip < np; ++ip) {
    double x = coord[3 * ip];
    double r2 = coord[3 * ip + 1] * coord[3 * ip + 1] + coord[3 * ip + 2] * coord[3 * ip + 2];
    if (x >= lims[0] && x <= lims[1] && r2 >= lims[2] && r2 <= lims[3]) indx2stay[n++] = (int)ip;
  }
  *num2stay = n;
  return 0;
}

// particle_tools.f90:155-208 ; nxg = len(Xgrid) (Fortran Xgrid(0:nx), nx = nxg-1)
int oracle_chunk_coords_boundaries(int8_t* chunked_indx, int* IndInChnk, int* GoOut, const double* coord,
                                   const double* lims, const double* Xgrid, int nchnk, i64 np, i64 nxg) {
  const i64 nx = nxg - 1;
  double inv = (nchnk > 1) ? 1.0 / (Xgrid[(nx + 1) / nchnk] - Xgrid[0]) : 1.0 / (Xgrid[nx] - Xgrid[0]);
  std::vector<i64> num(nchnk, 0);
  i64 out = 0;
  for (i64 ip = 0; ip < np; ++ip) {
    chunked_indx[ip] = -2;
    double x = coord[3 * ip];
    i64 ichnk = ifloor((x - Xgrid[0]) * inv);
    double r2 = coord[3 * ip + 1] * coord[3 * ip + 1] + coord[3 * ip + 2] * coord[3 * ip + 2];
    if (x >= lims[0] && x <= lims[1] && r2 >= lims[2] && r2 <= lims[3]) {
      // the Fortran indexes NumInChnk_loc(ichnk+1) unchecked; stay in range here
      if (ichnk < 0) ichnk = 0;
      if (ichnk > nchnk - 1) ichnk = nchnk - 1;
      num[ichnk] += 1;
      chunked_indx[ip] = (int8_t)ichnk;
    } else {
      out += 1;
    }
  }
  *GoOut = (int)out;
  IndInChnk[0] = 0;
  for (int c = 0; c < nchnk; ++c) IndInChnk[c + 1] = IndInChnk[c] + (int)num[c];
  return 0;
}

// particle_tools.f90:270-296 ; dat(3,np0), idx(np) 0-based
int oracle_align_data_vec(double* dat, const i64* idx, i64 np, i64 np0) {
  (void)np0;
  std::vector<double> tmp(3 * (size_t)np);
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip)
    for (int i = 0; i < 3; ++i) tmp[3 * ip + i] = dat[3 * idx[ip] + i];
  std::memcpy(dat, tmp.data(), sizeof(double) * 3 * np);
  return 0;
}

// particle_tools.f90:298-324
int oracle_align_data_scl(double* dat, const i64* idx, i64 np, i64 np0) {
  (void)np0;
  std::vector<double> tmp((size_t)np);
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) tmp[ip] = dat[idx[ip]];
  std::memcpy(dat, tmp.data(), sizeof(double) * np);
  return 0;
}

// particle_tools.f90:326-347
int oracle_sortoutghosts(int* indx2stay, int* num2stay, const double* coord, i64 np) {
  int n = 0;
  for (i64 ip = 0; ip < np; ++ip) indx2stay[ip] = 0;
  for (i64 ip = 0; ip < np; ++ip)
    if (coord[ip] != 0.0) indx2stay[n++] = (int)ip;
  *num2stay = n;
  return 0;
}

// =======================================================================================
// grid_deps*.f90 : deposition / gather.  One generic kernel covers the 8 deposit variants.
//   env  = 0: grid_deps.f90 / grid_deps_chnk.f90     (modes 0..nko, nm = nko+1)
//   env  = 1: grid_deps_env.f90 / grid_deps_env_chnk.f90 (modes -nko..nko, nm = 2 nko+1)
//   curr = 1: current (3 comps, env: only l=3 deposited, grid_deps_env.f90:76)
//   chunks != nullptr: chunked variants with guard semantics (grid_deps_chnk.f90:95-121)
// =======================================================================================
}  // extern "C"

namespace {

struct DepArgs {
  const double *coord, *momenta, *wghts;
  cd* grid;
  double leftX;
  const double* Rgrid;
  double dx_inv, dr_inv, kx0;
  i64 np, nxn, nrn, nm;
  int env, curr;
  const int* chunks;
  int guards, nchnk;
};

// Deposit the particles [ip0, ip1) of chunk `ichnk` (or of the whole array when chunks==0).
// `interior` receives direct writes; `left`/`right` (may be null) are the thread-private
// guard buffers of grid_deps_chnk.f90:46-47: left has x-extent guards+1 holding local nodes
// -guards..0, right holds local nodes cs..cs+guards.
void deposit_range(const DepArgs& a, i64 ip0, i64 ip1, int ichnk, i64 cs, cd* left, cd* right) {
  const i64 nxn = a.nxn, nrn = a.nrn, nm = a.nm;
  const int ncomp = a.curr ? 3 : 1;
  const i64 nko = a.env ? (nm - 1) / 2 : nm - 1;
  const double rmax = a.Rgrid[nrn - 1];
  const i64 nxleft = (i64)ichnk * cs;
  const i64 g = a.guards;
  std::vector<cd> phaseO(nko + 1);
  for (i64 ip = ip0; ip < ip1; ++ip) {
    double wp = a.wghts[ip];
    if (wp == 0.0) continue;
    double xp = a.coord[3 * ip], yp = a.coord[3 * ip + 1], zp = a.coord[3 * ip + 2];
    double rp = std::sqrt(yp * yp + zp * zp);
    if (rp >= rmax) continue;
    double veloc[3] = {0, 0, 0};
    if (a.curr) {
      for (int l = 0; l < 3; ++l) veloc[l] = a.momenta[3 * ip + l];
      if (std::fabs(veloc[0]) + std::fabs(veloc[1]) + std::fabs(veloc[2]) == 0.0) continue;
      double gp = std::sqrt(1.0 + veloc[0] * veloc[0] + veloc[1] * veloc[1] + veloc[2] * veloc[2]);
      if (a.env) for (int l = 0; l < 3; ++l) veloc[l] = veloc[l] / gp;          // grid_deps_env.f90:44
      else       for (int l = 0; l < 3; ++l) veloc[l] = veloc[l] * wp / gp;     // grid_deps.f90:44
    }
    cd wpc(wp, 0.0);
    if (a.env) wpc = cd(wp * std::cos(xp * a.kx0), -wp * std::sin(xp * a.kx0)); // grid_deps_env.f90:46

    i64 ix = ifloor((xp - a.leftX) * a.dx_inv);
    i64 ir = ifloor((rp - a.Rgrid[0]) * a.dr_inv);
    if (ir < 0 || ir > nrn - 2) continue;  // (never for 0<=rp<rmax on a uniform grid)
    double S0[2][2];  // S0[k][dim]
    S0[1][0] = (xp - a.leftX) * a.dx_inv - (double)ix;
    S0[0][0] = 1.0 - S0[1][0];
    S0[1][1] = (rp - a.Rgrid[ir]) * a.dr_inv;
    S0[0][1] = 1.0 - S0[1][1];

    cd phase_m = (rp > 0.0) ? cd(yp / rp, -zp / rp) : cd(0.0, 0.0);
    phaseO[0] = cd(1.0, 0.0);
    for (i64 iO = 1; iO <= nko; ++iO) phaseO[iO] = phaseO[iO - 1] * phase_m;

    // cell weights, [i][k] = Sx(i)*Sr(k) (times the complex weight per variant)
    cd cp[2][2];
    for (int k = 0; k < 2; ++k)
      for (int i = 0; i < 2; ++i) {
        double s = S0[i][0] * S0[k][1];
        if (a.curr) cp[i][k] = a.env ? cd(s, 0.0) * wpc : cd(s, 0.0);
        else        cp[i][k] = a.env ? (cd(s, 0.0) * wpc) * wpc   // Q2: weight applied twice, grid_deps_env.f90:145,147
                                     : cd(s * wp, 0.0);
      }

    const int l0 = (a.curr && a.env) ? 2 : 0;  // Q1: env current deposits l=3 only
    for (int l = l0; l < ncomp; ++l) {
      const double vl = a.curr ? veloc[l] : 1.0;
      for (i64 iO = 0; iO <= nko; ++iO) {
        for (int sgn = 0; sgn < (a.env && iO > 0 ? 2 : 1); ++sgn) {
          cd ph = sgn ? conj(phaseO[iO]) : phaseO[iO];
          i64 slot = a.env ? (nko + (sgn ? -iO : iO)) : iO;
          cd f = ph * vl;
          for (int i = 0; i < 2; ++i) {
            i64 gx = ix + i;           // global node
            i64 lx = gx - nxleft;      // node local to the chunk
            cd* dst = nullptr;
            i64 sx = 0, ext = 0;       // x index / x extent inside dst
            if (a.chunks) {
              if (lx <= 0) {           // grid_deps_chnk.f90:95
                if (lx < -g) continue; // outside loc_left: UB in the reference, dropped
                dst = left; sx = lx + g; ext = g + 1;
              } else if (lx >= cs) {   // grid_deps_chnk.f90:98
                if (lx > cs + g) continue;
                dst = right; sx = lx - cs; ext = g + 1;
              } else {
                dst = a.grid; sx = gx; ext = nxn;
              }
            } else {
              if (gx < 0 || gx > nxn - 1) continue;  // UB in the reference, dropped
              dst = a.grid; sx = gx; ext = nxn;
            }
            for (int k = 0; k < 2; ++k)
              dst[sx + ext * ((ir + k) + nrn * (slot + nm * l))] += f * cp[i][k];
          }
        }
      }
    }
  }
}

void ghost_fold(cd* grid, i64 nxn, i64 nrn, i64 nm, int ncomp) {
  // grid_deps.f90:80-85 : J(:,1) -= J(:,0); J(:,0) = 0
  for (i64 q = 0; q < nm * ncomp; ++q) {
    cd* pl = grid + nxn * nrn * q;
    for (i64 ix = 0; ix < nxn; ++ix) {
      pl[ix + nxn] -= pl[ix];
      pl[ix] = cd(0.0, 0.0);
    }
  }
}

int deposit(const DepArgs& a) {
  const int ncomp = a.curr ? 3 : 1;
  if (!a.chunks) {
    deposit_range(a, 0, a.np, 0, a.nxn, nullptr, nullptr);
  } else {
    const i64 cs = a.nxn / a.nchnk;  // chunk_size = (nx+1)/nchnk, grid_deps_chnk.f90:38
    const i64 g = a.guards;
    const i64 bufsz = (g + 1) * a.nrn * a.nm * ncomp;
    std::vector<std::vector<cd>> L(a.nchnk), R(a.nchnk);
    // one thread per chunk, as omp_set_num_threads(nchnk) does at grid_deps_chnk.f90:39
#pragma omp parallel for schedule(static, 1) num_threads(a.nchnk)
    for (int c = 0; c < a.nchnk; ++c) {
      L[c].assign(bufsz, cd());
      R[c].assign(bufsz, cd());
      deposit_range(a, a.chunks[c], a.chunks[c + 1], c, cs, L[c].data(), R[c].data());
    }
    // guard exchange (grid_deps_chnk.f90:110-119), done serially => deterministic and free
    // of the reference's pre-barrier race.
    for (int c = 0; c < a.nchnk; ++c) {
      const i64 nxleft = (i64)c * cs;
      if (nxleft + cs + g <= a.nxn - 1)
        for (i64 q = 0; q < a.nrn * a.nm * ncomp; ++q)
          for (i64 j = 0; j <= g; ++j) a.grid[nxleft + cs + j + a.nxn * q] += R[c][j + (g + 1) * q];
      if (nxleft - g >= 0)
        for (i64 q = 0; q < a.nrn * a.nm * ncomp; ++q)
          for (i64 j = 0; j <= g; ++j) a.grid[nxleft - g + j + a.nxn * q] += L[c][j + (g + 1) * q];
    }
  }
  ghost_fold(a.grid, a.nxn, a.nrn, a.nm, ncomp);
  return 0;
}

// grid_deps.f90:149-217 (env=0) and grid_deps_env.f90:164-238 (env=1)
int gather(const double* coord, const double* wghts, const cd* Fld, double* Fld_tot, double leftX,
           const double* Rgrid, double dx_inv, double dr_inv, double kx0, i64 np, i64 nxn, i64 nrn,
           i64 nm, int env) {
  const i64 nko = env ? (nm - 1) / 2 : nm - 1;
  const double rmax = Rgrid[nrn - 1];
#pragma omp parallel
  {
    std::vector<cd> phaseO(nko + 1);
#pragma omp for schedule(static)
    for (i64 ip = 0; ip < np; ++ip) {
      double wp = wghts[ip];
      if (wp == 0.0) continue;
      double xp = coord[3 * ip], yp = coord[3 * ip + 1], zp = coord[3 * ip + 2];
      double rp = std::sqrt(yp * yp + zp * zp);
      if (rp >= rmax) continue;
      i64 ix = ifloor((xp - leftX) * dx_inv);
      i64 ir = ifloor((rp - Rgrid[0]) * dr_inv);
      if (ix < 0 || ix > nxn - 2 || ir < 0 || ir > nrn - 2) continue;  // UB in the reference
      double S0[2][2];
      S0[1][0] = (xp - leftX) * dx_inv - (double)ix;
      S0[0][0] = 1.0 - S0[1][0];
      S0[1][1] = (rp - Rgrid[ir]) * dr_inv;
      S0[0][1] = 1.0 - S0[1][1];
      // Q4: phase at r=0 is 0 for proj_fld (grid_deps.f90:190), 1 for proj_fld_env (:205)
      cd phase_p = (rp > 0.0) ? cd(yp / rp, zp / rp) : (env ? cd(1.0, 0.0) : cd(0.0, 0.0));
      phaseO[0] = cd(1.0, 0.0);
      for (i64 iO = 1; iO <= nko; ++iO) phaseO[iO] = phaseO[iO - 1] * phase_p;
      cd car(1.0, 0.0);
      if (env) car = cd(std::cos(xp * kx0), std::sin(xp * kx0));
      double Fp[6] = {0, 0, 0, 0, 0, 0};
      for (i64 iO = 0; iO <= nko; ++iO)
        for (int sgn = 0; sgn < (env && iO > 0 ? 2 : 1); ++sgn) {
          cd ph = sgn ? conj(phaseO[iO]) : phaseO[iO];
          i64 slot = env ? (nko + (sgn ? -iO : iO)) : iO;
          for (int l = 0; l < 6; ++l) {
            const cd* pl = Fld + nxn * nrn * (slot + nm * l);
            double acc = 0.0;
            for (int k = 0; k < 2; ++k)
              for (int i = 0; i < 2; ++i) {
                cd pc = (cd(S0[k][1] * S0[i][0], 0.0) * car) * ph;
                cd f = pl[(ix + i) + nxn * (ir + k)];
                acc += pc.re * f.re - pc.im * f.im;  // DBLE(projcomp*Fld)
              }
            Fp[l] += acc;
          }
        }
      for (int l = 0; l < 6; ++l) Fld_tot[6 * ip + l] += Fp[l];
    }
  }
  return 0;
}

}  // namespace

extern "C" {

// grid_deps.f90:18-87
int oracle_dep_curr(const double* coord, const double* momenta, const double* wghts, double* curr,
                    double leftX, const double* Rgrid, double dx_inv, double dr_inv, i64 np, i64 nxn,
                    i64 nrn, i64 nm) {
  DepArgs a{coord, momenta, wghts, (cd*)curr, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, 0, 1, nullptr, 0, 1};
  return deposit(a);
}
// grid_deps.f90:89-147
int oracle_dep_dens(const double* coord, const double* wghts, double* dens, double leftX,
                    const double* Rgrid, double dx_inv, double dr_inv, i64 np, i64 nxn, i64 nrn, i64 nm) {
  DepArgs a{coord, nullptr, wghts, (cd*)dens, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, 0, 0, nullptr, 0, 1};
  return deposit(a);
}
// grid_deps_chnk.f90:18-130
int oracle_dep_curr_chnk(const double* coord, const double* momenta, const double* wghts, double* curr,
                         const int* IndInChunk, int guards, double leftX, const double* Rgrid,
                         double dx_inv, double dr_inv, i64 np, i64 nxn, i64 nrn, i64 nm, i64 nchnk) {
  DepArgs a{coord, momenta, wghts, (cd*)curr, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, 0, 1, IndInChunk, guards, (int)nchnk};
  return deposit(a);
}
// grid_deps_chnk.f90:132-234
int oracle_dep_dens_chnk(const double* coord, const double* wghts, double* dens, const int* IndInChunk,
                         int guards, double leftX, const double* Rgrid, double dx_inv, double dr_inv,
                         i64 np, i64 nxn, i64 nrn, i64 nm, i64 nchnk) {
  DepArgs a{coord, nullptr, wghts, (cd*)dens, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, 0, 0, IndInChunk, guards, (int)nchnk};
  return deposit(a);
}
// grid_deps_env.f90:18-95
int oracle_dep_curr_env(const double* coord, const double* momenta, const double* wghts, double* curr,
                        double leftX, const double* Rgrid, double dx_inv, double dr_inv, double kx0,
                        i64 np, i64 nxn, i64 nrn, i64 nm) {
  DepArgs a{coord, momenta, wghts, (cd*)curr, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, 1, 1, nullptr, 0, 1};
  return deposit(a);
}
// grid_deps_env.f90:97-162
int oracle_dep_dens_env(const double* coord, const double* wghts, double* dens, double leftX,
                        const double* Rgrid, double dx_inv, double dr_inv, double kx0, i64 np, i64 nxn,
                        i64 nrn, i64 nm) {
  DepArgs a{coord, nullptr, wghts, (cd*)dens, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, 1, 0, nullptr, 0, 1};
  return deposit(a);
}
// grid_deps_env_chnk.f90:18-145
int oracle_dep_curr_env_chnk(const double* coord, const double* momenta, const double* wghts,
                             double* curr, const int* IndInChunk, int guards, double leftX,
                             const double* Rgrid, double dx_inv, double dr_inv, double kx0, i64 np,
                             i64 nxn, i64 nrn, i64 nm, i64 nchnk) {
  DepArgs a{coord, momenta, wghts, (cd*)curr, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, 1, 1, IndInChunk, guards, (int)nchnk};
  return deposit(a);
}
// grid_deps_env_chnk.f90:147-261
int oracle_dep_dens_env_chnk(const double* coord, const double* wghts, double* dens,
                             const int* IndInChunk, int guards, double leftX, const double* Rgrid,
                             double dx_inv, double dr_inv, double kx0, i64 np, i64 nxn, i64 nrn, i64 nm,
                             i64 nchnk) {
  DepArgs a{coord, nullptr, wghts, (cd*)dens, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, 1, 0, IndInChunk, guards, (int)nchnk};
  return deposit(a);
}

// grid_deps.f90:149-217
int oracle_proj_fld(const double* coord, const double* wghts, const double* Fld, double* Fld_tot,
                    double leftX, const double* Rgrid, double dx_inv, double dr_inv, i64 np, i64 nxn,
                    i64 nrn, i64 nm) {
  return gather(coord, wghts, (const cd*)Fld, Fld_tot, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, 0);
}
// grid_deps_env.f90:164-238
int oracle_proj_fld_env(const double* coord, const double* wghts, const double* Fld, double* Fld_tot,
                        double leftX, const double* Rgrid, double dx_inv, double dr_inv, double kx0,
                        i64 np, i64 nxn, i64 nrn, i64 nm) {
  return gather(coord, wghts, (const cd*)Fld, Fld_tot, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, 1);
}

// grid_deps.f90:219-266 : m=0 * 1/2pi, m>0 * 1/pi ; ghost row: m=0 copy, m>0 negate
int oracle_eb_correction(double* eb_spc, i64 nxn, i64 nrn, i64 nm) {
  const double pi = 4.0 * std::atan(1.0);
  const double pi_inv = 1. / pi, pi2_inv = 0.5 * pi_inv;
  cd* eb = (cd*)eb_spc;
#pragma omp parallel for schedule(static) collapse(2)
  for (i64 l = 0; l < 6; ++l)
    for (i64 m = 0; m < nm; ++m) {
      cd* pl = eb + nxn * nrn * (m + nm * l);
      const double f = (m == 0) ? pi2_inv : pi_inv;
      for (i64 i = 0; i < nxn * nrn; ++i) pl[i] = f * pl[i];
      for (i64 ix = 0; ix < nxn; ++ix) pl[ix] = (m == 0) ? pl[ix + nxn] : -pl[ix + nxn];
    }
  return 0;
}
// grid_deps_env.f90:240-283 : all modes * 1/pi ; ghost: copy if nko==0 else negate (Q6)
int oracle_eb_correction_env(double* eb_spc, i64 nxn, i64 nrn, i64 nm) {
  const double pi = 4.0 * std::atan(1.0);
  const double pi_inv = 1. / pi;
  const i64 nko = (nm - 1) / 2;
  cd* eb = (cd*)eb_spc;
#pragma omp parallel for schedule(static) collapse(2)
  for (i64 l = 0; l < 6; ++l)
    for (i64 m = 0; m < nm; ++m) {
      cd* pl = eb + nxn * nrn * (m + nm * l);
      for (i64 i = 0; i < nxn * nrn; ++i) pl[i] = pi_inv * pl[i];
      for (i64 ix = 0; ix < nxn; ++ix) pl[ix] = (nko == 0) ? pl[ix + nxn] : -pl[ix + nxn];
    }
  return 0;
}

// =======================================================================================
// fb_io.f90 : DHT (over r) + FFT (over x)
// =======================================================================================
}  // extern "C"

namespace {

// forward: out(:,ik,m,l) = FFT_x[ sum_ir In(ir,ik,m) * in(:,1+ir,m,l) ] * exp(-i kx leftX)
// fb_io.f90:18-59 (ncomp=3), :61-98 (ncomp=1)
int fb_in(cd* out, const cd* in, double leftX, const double* kx, const double* In, i64 nkx, i64 nrn,
          i64 nm, i64 nkr, int ncomp) {
  const i64 nr = nrn - 1;
  FFTPlan plan((int)nkx);
  std::vector<cd> shiftX(nkx);
  for (i64 i = 0; i < nkx; ++i) shiftX[i] = cd(std::cos(leftX * kx[i]), -std::sin(leftX * kx[i]));
  for (int l = 0; l < ncomp; ++l)
    for (i64 m = 0; m < nm; ++m) {
#pragma omp parallel
      {
        std::vector<cd> A(nkx), B(nkx);
#pragma omp for schedule(static)
        for (i64 ik = 0; ik < nkr; ++ik) {
          std::fill(A.begin(), A.end(), cd());
          for (i64 ir = 0; ir < nr; ++ir) {
            const double c = In[ir + nr * (ik + nkr * m)];
            const cd* src = in + nkx * ((ir + 1) + nrn * (m + nm * l));
            for (i64 i = 0; i < nkx; ++i) { A[i].re += c * src[i].re; A[i].im += c * src[i].im; }
          }
          plan.exec(A.data(), B.data(), -1);
          cd* dst = out + nkx * (ik + nkr * (m + nm * l));
          for (i64 i = 0; i < nkx; ++i) dst[i] = B[i] * shiftX[i];
        }
      }
    }
  return 0;
}

// backward: out(:,1+ir,m) = IFFT_x[ ( sum_ik Out(ik,ir,m) * in(:,ik,m) ) * exp(+i kx leftX) ]
// (unnormalised), ghost row ir=0 zero.  fb_io.f90:100-140,142-180,182-228
void fb_out_plane(cd* out_pl /*(nkx,nrn)*/, const cd* in_pl /*(nkx,nkr)*/, const double* Out_m /*(nkr,nr)*/,
                  const std::vector<cd>& shiftX, const FFTPlan& plan, i64 nkx, i64 nrn, i64 nkr) {
  const i64 nr = nrn - 1;
  for (i64 i = 0; i < nkx; ++i) out_pl[i] = cd();
#pragma omp parallel
  {
    std::vector<cd> A(nkx), B(nkx);
#pragma omp for schedule(static)
    for (i64 ir = 0; ir < nr; ++ir) {
      std::fill(A.begin(), A.end(), cd());
      for (i64 ik = 0; ik < nkr; ++ik) {
        const double c = Out_m[ik + nkr * ir];
        const cd* src = in_pl + nkx * ik;
        for (i64 i = 0; i < nkx; ++i) { A[i].re += c * src[i].re; A[i].im += c * src[i].im; }
      }
      for (i64 i = 0; i < nkx; ++i) A[i] = A[i] * shiftX[i];
      plan.exec(A.data(), B.data(), +1);
      cd* dst = out_pl + nkx * (ir + 1);
      for (i64 i = 0; i < nkx; ++i) dst[i] = B[i];
    }
  }
}

}  // namespace

extern "C" {

int oracle_fb_vec_in(double* vec_fb, const double* vec, double leftX, const double* kx, const double* In,
                     i64 nkx, i64 nrn, i64 nm, i64 nkr) {
  return fb_in((cd*)vec_fb, (const cd*)vec, leftX, kx, In, nkx, nrn, nm, nkr, 3);
}
int oracle_fb_scl_in(double* scl_fb, const double* scl, double leftX, const double* kx, const double* In,
                     i64 nkx, i64 nrn, i64 nm, i64 nkr) {
  return fb_in((cd*)scl_fb, (const cd*)scl, leftX, kx, In, nkx, nrn, nm, nkr, 1);
}
static int fb_out_n(cd* out, const cd* in, double leftX, const double* kx, const double* Out, i64 nkx,
                    i64 nrn, i64 nm, i64 nkr, int ncomp) {
  const i64 nr = nrn - 1;
  FFTPlan plan((int)nkx);
  std::vector<cd> shiftX(nkx);
  for (i64 i = 0; i < nkx; ++i) shiftX[i] = cd(std::cos(leftX * kx[i]), std::sin(leftX * kx[i]));
  for (int l = 0; l < ncomp; ++l)
    for (i64 m = 0; m < nm; ++m)
      fb_out_plane(out + nkx * nrn * (m + nm * l), in + nkx * nkr * (m + nm * l), Out + nkr * nr * m,
                   shiftX, plan, nkx, nrn, nkr);
  return 0;
}
int oracle_fb_vec_out(double* vec, const double* vec_fb, double leftX, const double* kx, const double* Out,
                      i64 nkx, i64 nrn, i64 nm, i64 nkr) {
  return fb_out_n((cd*)vec, (const cd*)vec_fb, leftX, kx, Out, nkx, nrn, nm, nkr, 3);
}
int oracle_fb_scl_out(double* scl, const double* scl_fb, double leftX, const double* kx, const double* Out,
                      i64 nkx, i64 nrn, i64 nm, i64 nkr) {
  return fb_out_n((cd*)scl, (const cd*)scl_fb, leftX, kx, Out, nkx, nrn, nm, nkr, 1);
}
// fb_io.f90:182-228 : comps 1..3 from e_fb(...,1:3) (e_fb has 6 comps), 4..6 from b_fb
int oracle_fb_eb_out(double* eb_spc, const double* e_fb, const double* b_fb, double leftX, const double* kx,
                     const double* Out, i64 nkx, i64 nrn, i64 nm, i64 nkr) {
  fb_out_n((cd*)eb_spc, (const cd*)e_fb, leftX, kx, Out, nkx, nrn, nm, nkr, 3);
  fb_out_n((cd*)eb_spc + 3 * nkx * nrn * nm, (const cd*)b_fb, leftX, kx, Out, nkx, nrn, nm, nkr, 3);
  return 0;
}

// fb_io.f90:230-308 : x-space window applied to spectral fields.
// modefilt 0 = left (the only mode the driver uses, solvers.py:619).  Modes 1/2 multiply
// Aifft(nkx-nxfilt:nkx) (nxfilt+1 elements) by filtr(nxfilt:1:-1) (nxfilt elements): a shape
// mismatch in the reference; here the last nxfilt samples are used.
int oracle_fb_filtr(double* vec_, double leftX, const double* kx, const double* filtr, int modefilt,
                    i64 nkx, i64 nkr, i64 nm, i64 nxfilt) {
  cd* vec = (cd*)vec_;
  FFTPlan plan((int)nkx);
  std::vector<cd> shiftX(nkx), shiftX_inv(nkx);
  for (i64 i = 0; i < nkx; ++i) {
    shiftX[i] = cd(std::cos(leftX * kx[i]), std::sin(leftX * kx[i]));
    shiftX_inv[i] = cdiv(cd(1.0, 0.0), (double)nkx * shiftX[i]);
  }
#pragma omp parallel
  {
    std::vector<cd> A(nkx), B(nkx);
#pragma omp for schedule(static)
    for (i64 q = 0; q < nkr * nm * 3; ++q) {
      cd* v = vec + nkx * q;
      for (i64 i = 0; i < nkx; ++i) A[i] = v[i] * shiftX[i];
      plan.exec(A.data(), B.data(), +1);
      if (modefilt == 0 || modefilt == 2)
        for (i64 i = 0; i < nxfilt; ++i) B[i] = B[i] * filtr[i];
      if (modefilt == 1 || modefilt == 2)
        for (i64 i = 0; i < nxfilt; ++i) B[nkx - nxfilt + i] = B[nkx - nxfilt + i] * filtr[nxfilt - 1 - i];
      plan.exec(B.data(), A.data(), -1);
      for (i64 i = 0; i < nkx; ++i) v[i] = A[i] * shiftX_inv[i];
    }
  }
  return 0;
}

// =======================================================================================
// fb_math.f90 / fb_math_env.f90 : spectral-space vector calculus
// =======================================================================================
}  // extern "C"

namespace {

// acc(:) += sum_ik D(ik,ik_loc) * src(:,ik)      (real D, complex src)
inline void contract(cd* acc, const double* D /*column ik_loc: D + nkr*ik_loc*/, const cd* src, i64 nkx, i64 nkr) {
  for (i64 ik = 0; ik < nkr; ++ik) {
    const double c = D[ik];
    const cd* s = src + nkx * ik;
    for (i64 i = 0; i < nkx; ++i) { acc[i].re += c * s[i].re; acc[i].im += c * s[i].im; }
  }
}

struct FBDims {
  i64 nkx, nkr, nm, nkr_loc;
  int env;
  i64 nko() const { return env ? (nm - 1) / 2 : nm - 1; }
  // mode number of slot s, and slot of mode number
  i64 lo() const { return env ? -nko() : 0; }
  i64 hi() const { return nko(); }
  // D matrices have slots lo-? : real: 0..nko+1 ; env: -nko-1..nko+1
  i64 dslot(i64 mode) const { return env ? mode + nko() + 1 : mode; }
  i64 vslot(i64 mode) const { return mode - lo(); }
};

// mirrored mode: ext(1,:) = -conj(f(1,:)), ext(2:nkx,:) = -conj(f(nkx:2:-1,:))  fb_math.f90:35-36
void mirror(cd* ext, const cd* f, i64 nkx, i64 nkr) {
  for (i64 ik = 0; ik < nkr; ++ik) {
    ext[nkx * ik] = -conj(f[nkx * ik]);
    for (i64 i = 1; i < nkx; ++i) ext[i + nkx * ik] = -conj(f[(nkx - i) + nkx * ik]);
  }
}

// divergence-like pass shared by fb_div (fb_math.f90:151-199), fb_div_env (fb_math_env.f90:63-104)
// and the first half of fb_graddiv (fb_math.f90:226-256) / fb_graddiv_env (fb_math_env.f90:178-205).
// scl has mode slots [slo..shi] (graddiv: one extra on each open side).
void div_pass(cd* scl, i64 slo, i64 shi, const cd* vec, const double* Dp, const double* Dm,
              const double* kx, const FBDims& d, bool guard_nko0) {
  const i64 nkx = d.nkx, nkr = d.nkr, nkl = d.nkr_loc, nm = d.nm;
  const i64 pl_v = nkx * nkr;  // plane size of vec
  auto V = [&](i64 mode, int l) { return vec + pl_v * (d.vslot(mode) + nm * l); };
  std::vector<cd> ext;  // real solver: -conj(mirror) of mode 1, comps 2..3
  if (!d.env) {
    ext.assign(2 * pl_v, cd());
    // Q7: with nko=0 the reference reads mode slot 1 out of bounds (fb_math.f90:166); fb_graddiv
    // guards it (:217).  Both are treated as "missing mode = 0" here.
    if (d.nko() > 0) {
      mirror(ext.data(), V(1, 1), nkx, nkr);
      mirror(ext.data() + pl_v, V(1, 2), nkx, nkr);
    }
    (void)guard_nko0;
  }
  for (i64 mode = slo; mode <= shi; ++mode) {
    cd* out_m = scl + nkx * nkl * (mode - slo);
#pragma omp parallel
    {
      std::vector<cd> s(nkx), tmp(nkx);
#pragma omp for schedule(static)
      for (i64 ikl = 0; ikl < nkl; ++ikl) {
        cd* o = out_m + nkx * ikl;
        std::fill(s.begin(), s.end(), cd());
        if (mode >= d.lo() && mode <= d.hi()) {
          const cd* v1 = V(mode, 0) + nkx * ikl;
          for (i64 i = 0; i < nkx; ++i) o[i] += mul_i(v1[i]) * kx[i];
        }
        const double* dm = Dm + nkr * (ikl + nkl * d.dslot(mode));
        const double* dp = Dp + nkr * (ikl + nkl * d.dslot(mode));
        if (!d.env) {
          // real: + Dm(mode) * (i v3 - v2)[mode-1]  (mode 0: mirrored ext) ; + Dp(mode) * (i v3 + v2)[mode+1]
          for (i64 ik = 0; ik < nkr; ++ik) {
            const double c = dm[ik];
            const cd *a2, *a3;
            if (mode > 0) { a2 = V(mode - 1, 1) + nkx * ik; a3 = V(mode - 1, 2) + nkx * ik; }
            else          { a2 = ext.data() + nkx * ik;     a3 = ext.data() + pl_v + nkx * ik; }
            for (i64 i = 0; i < nkx; ++i) s[i] += c * (mul_i(a3[i]) - a2[i]);
          }
          if (mode < d.nko())
            for (i64 ik = 0; ik < nkr; ++ik) {
              const double c = dp[ik];
              const cd* a2 = V(mode + 1, 1) + nkx * ik;
              const cd* a3 = V(mode + 1, 2) + nkx * ik;
              for (i64 i = 0; i < nkx; ++i) s[i] += c * (mul_i(a3[i]) + a2[i]);
            }
        } else {
          // env: - Dm(mode) * (v2 - i v3)[mode-1] (mode > -nko) ; + Dp(mode) * (v2 + i v3)[mode+1] (mode < nko)
          if (mode > -d.nko())
            for (i64 ik = 0; ik < nkr; ++ik) {
              const double c = dm[ik];
              const cd* a2 = V(mode - 1, 1) + nkx * ik;
              const cd* a3 = V(mode - 1, 2) + nkx * ik;
              for (i64 i = 0; i < nkx; ++i) s[i] -= c * (a2[i] - mul_i(a3[i]));
            }
          if (mode < d.nko())
            for (i64 ik = 0; ik < nkr; ++ik) {
              const double c = dp[ik];
              const cd* a2 = V(mode + 1, 1) + nkx * ik;
              const cd* a3 = V(mode + 1, 2) + nkx * ik;
              for (i64 i = 0; i < nkx; ++i) s[i] += c * (a2[i] + mul_i(a3[i]));
            }
        }
        for (i64 i = 0; i < nkx; ++i) o[i] += s[i];
      }
    }
  }
}

// gradient-like pass shared by fb_grad (fb_math.f90:96-149), fb_grad_env (fb_math_env.f90:18-61) and the
// second half of fb_graddiv[_env].  scl has mode slots [slo..shi]; modes outside contribute nothing.
// `always_both` (graddiv second pass) applies both couplings for every output mode, reading the
// extra scl slots (fb_math.f90:283-287, fb_math_env.f90:216-227).
void grad_pass(cd* vec, const cd* scl, i64 slo, i64 shi, const double* Dp, const double* Dm,
               const double* kx, const FBDims& d, bool always_both) {
  const i64 nkx = d.nkx, nkr = d.nkr, nkl = d.nkr_loc, nm = d.nm;
  auto S = [&](i64 mode) { return scl + nkx * nkr * (mode - slo); };
  std::vector<cd> ext;
  if (!d.env) {
    ext.assign(nkx * nkr, cd());
    if (d.nko() > 0) mirror(ext.data(), S(1), nkx, nkr);  // Q7 again for nko = 0
  }
  for (i64 mode = d.lo(); mode <= d.hi(); ++mode) {
    cd* o1 = vec + nkx * nkl * (d.vslot(mode) + nm * 0);
    cd* o2 = vec + nkx * nkl * (d.vslot(mode) + nm * 1);
    cd* o3 = vec + nkx * nkl * (d.vslot(mode) + nm * 2);
#pragma omp parallel
    {
      std::vector<cd> s(nkx);
#pragma omp for schedule(static)
      for (i64 ikl = 0; ikl < nkl; ++ikl) {
        const cd* sc = S(mode) + nkx * ikl;
        cd* a1 = o1 + nkx * ikl;
        cd* a2 = o2 + nkx * ikl;
        cd* a3 = o3 + nkx * ikl;
        for (i64 i = 0; i < nkx; ++i) a1[i] += mul_i(sc[i]) * kx[i];
        const double* dm = Dm + nkr * (ikl + nkl * d.dslot(mode));
        const double* dp = Dp + nkr * (ikl + nkl * d.dslot(mode));
        // m-1 coupling
        const cd* lower = nullptr;
        if (!d.env) lower = (mode > 0) ? S(mode - 1) : ext.data();
        else if (always_both || mode > -d.nko()) lower = S(mode - 1);
        if (lower) {
          std::fill(s.begin(), s.end(), cd());
          contract(s.data(), dm, lower, nkx, nkr);
          for (i64 i = 0; i < nkx; ++i) { a2[i] -= s[i]; a3[i] += mul_i(s[i]); }
        }
        // m+1 coupling
        if (always_both || mode < d.nko()) {
          std::fill(s.begin(), s.end(), cd());
          contract(s.data(), dp, S(mode + 1), nkx, nkr);
          for (i64 i = 0; i < nkx; ++i) { a2[i] += s[i]; a3[i] += mul_i(s[i]); }
        }
      }
    }
  }
}

// fb_rot (fb_math.f90:18-94) and fb_rot_env (fb_math_env.f90:106-162)
void rot_pass(cd* out, const cd* vec, const double* Dp, const double* Dm, const double* kx, const FBDims& d) {
  const i64 nkx = d.nkx, nkr = d.nkr, nkl = d.nkr_loc, nm = d.nm;
  const i64 pl_v = nkx * nkr;
  auto V = [&](i64 mode, int l) { return vec + pl_v * (d.vslot(mode) + nm * l); };
  std::vector<cd> ext;
  if (!d.env) {
    ext.assign(3 * pl_v, cd());
    if (d.nko() > 0)
      for (int l = 0; l < 3; ++l) mirror(ext.data() + pl_v * l, V(1, l), nkx, nkr);
  }
  for (i64 mode = d.lo(); mode <= d.hi(); ++mode) {
    cd* o1 = out + nkx * nkl * (d.vslot(mode) + nm * 0);
    cd* o2 = out + nkx * nkl * (d.vslot(mode) + nm * 1);
    cd* o3 = out + nkx * nkl * (d.vslot(mode) + nm * 2);
#pragma omp parallel
    {
      std::vector<cd> s(nkx);
#pragma omp for schedule(static)
      for (i64 ikl = 0; ikl < nkl; ++ikl) {
        cd* a1 = o1 + nkx * ikl;
        cd* a2 = o2 + nkx * ikl;
        cd* a3 = o3 + nkx * ikl;
        const cd* v2 = V(mode, 1) + nkx * ikl;
        const cd* v3 = V(mode, 2) + nkx * ikl;
        for (i64 i = 0; i < nkx; ++i) {
          a2[i] -= mul_i(v3[i]) * kx[i];
          a3[i] += mul_i(v2[i]) * kx[i];
        }
        const double* dm = Dm + nkr * (ikl + nkl * d.dslot(mode));
        const double* dp = Dp + nkr * (ikl + nkl * d.dslot(mode));
        if (mode < d.nko()) {
          std::fill(s.begin(), s.end(), cd());
          for (i64 ik = 0; ik < nkr; ++ik) {
            const double c = dp[ik];
            const cd* b2 = V(mode + 1, 1) + nkx * ik;
            const cd* b3 = V(mode + 1, 2) + nkx * ik;
            for (i64 i = 0; i < nkx; ++i) s[i] -= c * (mul_i(b2[i]) - b3[i]);
          }
          for (i64 i = 0; i < nkx; ++i) a1[i] += s[i];
          std::fill(s.begin(), s.end(), cd());
          contract(s.data(), dp, V(mode + 1, 0), nkx, nkr);
          for (i64 i = 0; i < nkx; ++i) { a2[i] += mul_i(s[i]); a3[i] -= s[i]; }
        }
        const cd *l1 = nullptr, *l2 = nullptr, *l3 = nullptr;
        if (!d.env) {
          if (mode > 0) { l1 = V(mode - 1, 0); l2 = V(mode - 1, 1); l3 = V(mode - 1, 2); }
          else { l1 = ext.data(); l2 = ext.data() + pl_v; l3 = ext.data() + 2 * pl_v; }
        } else if (mode > -d.nko()) {
          l1 = V(mode - 1, 0); l2 = V(mode - 1, 1); l3 = V(mode - 1, 2);
        }
        if (l1) {
          if (!d.env) {  // Q5: fb_rot_env computes this term then discards it (fb_math_env.f90:146-151)
            std::fill(s.begin(), s.end(), cd());
            for (i64 ik = 0; ik < nkr; ++ik) {
              const double c = dm[ik];
              const cd* b2 = l2 + nkx * ik;
              const cd* b3 = l3 + nkx * ik;
              for (i64 i = 0; i < nkx; ++i) s[i] -= c * (mul_i(b2[i]) + b3[i]);
            }
            for (i64 i = 0; i < nkx; ++i) a1[i] += s[i];
          }
          std::fill(s.begin(), s.end(), cd());
          contract(s.data(), dm, l1, nkx, nkr);
          for (i64 i = 0; i < nkx; ++i) { a2[i] += mul_i(s[i]); a3[i] += s[i]; }
        }
      }
    }
  }
}

}  // namespace

extern "C" {

#define FB_DIMS(env_) FBDims d{nkx, nkr, nm, nkr_loc, env_}

int oracle_fb_rot(double* out, const double* vec, const double* Dp, const double* Dm, const double* kx,
                  i64 nkx, i64 nkr, i64 nm, i64 nkr_loc) {
  FB_DIMS(0);
  std::memset(out, 0, sizeof(cd) * nkx * nkr_loc * nm * 3);
  rot_pass((cd*)out, (const cd*)vec, Dp, Dm, kx, d);
  return 0;
}
int oracle_fb_rot_env(double* out, const double* vec, const double* Dp, const double* Dm, const double* kx,
                      i64 nkx, i64 nkr, i64 nm, i64 nkr_loc) {
  FB_DIMS(1);
  std::memset(out, 0, sizeof(cd) * nkx * nkr_loc * nm * 3);
  rot_pass((cd*)out, (const cd*)vec, Dp, Dm, kx, d);
  return 0;
}
int oracle_fb_grad(double* out, const double* scl, const double* Dp, const double* Dm, const double* kx,
                   i64 nkx, i64 nkr, i64 nm, i64 nkr_loc) {
  FB_DIMS(0);
  std::memset(out, 0, sizeof(cd) * nkx * nkr_loc * nm * 3);
  grad_pass((cd*)out, (const cd*)scl, d.lo(), d.hi(), Dp, Dm, kx, d, false);
  return 0;
}
int oracle_fb_grad_env(double* out, const double* scl, const double* Dp, const double* Dm, const double* kx,
                       i64 nkx, i64 nkr, i64 nm, i64 nkr_loc) {
  FB_DIMS(1);
  std::memset(out, 0, sizeof(cd) * nkx * nkr_loc * nm * 3);
  grad_pass((cd*)out, (const cd*)scl, d.lo(), d.hi(), Dp, Dm, kx, d, false);
  return 0;
}
int oracle_fb_div(double* out, const double* vec, const double* Dp, const double* Dm, const double* kx,
                  i64 nkx, i64 nkr, i64 nm, i64 nkr_loc) {
  FB_DIMS(0);
  std::memset(out, 0, sizeof(cd) * nkx * nkr_loc * nm);
  div_pass((cd*)out, d.lo(), d.hi(), (const cd*)vec, Dp, Dm, kx, d, false);
  return 0;
}
int oracle_fb_div_env(double* out, const double* vec, const double* Dp, const double* Dm, const double* kx,
                      i64 nkx, i64 nkr, i64 nm, i64 nkr_loc) {
  FB_DIMS(1);
  std::memset(out, 0, sizeof(cd) * nkx * nkr_loc * nm);
  div_pass((cd*)out, d.lo(), d.hi(), (const cd*)vec, Dp, Dm, kx, d, false);
  return 0;
}
// fb_math.f90:201-293 : internal scalar has modes 0..nko+1
int oracle_fb_graddiv(double* vec, const double* Dp, const double* Dm, const double* kx, i64 nkx, i64 nkr,
                      i64 nm, i64 nkr_loc) {
  FB_DIMS(0);
  if (nkr != nkr_loc) return 2;
  std::vector<cd> scl((size_t)(nkx * nkr_loc * (nm + 1)));
  div_pass(scl.data(), 0, d.nko() + 1, (const cd*)vec, Dp, Dm, kx, d, true);
  std::memset(vec, 0, sizeof(cd) * nkx * nkr * nm * 3);
  grad_pass((cd*)vec, scl.data(), 0, d.nko() + 1, Dp, Dm, kx, d, true);
  return 0;
}
// fb_math_env.f90:164-233 : internal scalar has modes -nko-1..nko+1
int oracle_fb_graddiv_env(double* vec, const double* Dp, const double* Dm, const double* kx, i64 nkx,
                          i64 nkr, i64 nm, i64 nkr_loc) {
  FB_DIMS(1);
  if (nkr != nkr_loc) return 2;
  std::vector<cd> scl((size_t)(nkx * nkr_loc * (nm + 2)));
  div_pass(scl.data(), -d.nko() - 1, d.nko() + 1, (const cd*)vec, Dp, Dm, kx, d, true);
  std::memset(vec, 0, sizeof(cd) * nkx * nkr * nm * 3);
  grad_pass((cd*)vec, scl.data(), -d.nko() - 1, d.nko() + 1, Dp, Dm, kx, d, true);
  return 0;
}

// =======================================================================================
// maxwell_solvers.f90
// =======================================================================================

// maxwell_solvers.f90:18-60 ; C1,C2 real (nkx,nkr,nm,5)
int oracle_maxwell_push_with_spchrg(double* EG_, const double* j_, const double* gn_, const double* gnp1_,
                                    const double* C1, const double* C2, i64 nkx, i64 nkr, i64 nm) {
  cd* EG = (cd*)EG_;
  const cd *J = (const cd*)j_, *gn = (const cd*)gn_, *gp = (const cd*)gnp1_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 q = 0; q < nkr * nm; ++q)
    for (int l = 0; l < 3; ++l)
      for (i64 i = 0; i < nkx; ++i) {
        const i64 p = i + nkx * q;
        const cd e = EG[p + P * l], g = EG[p + P * (l + 3)], j = J[p + P * l], a = gn[p + P * l], b = gp[p + P * l];
        cd en = C1[p] * e + C1[p + P] * g + C1[p + 2 * P] * j + C1[p + 3 * P] * a + C1[p + 4 * P] * b;
        cd gnw = C2[p] * e + C2[p + P] * g + C2[p + 2 * P] * j + C2[p + 3 * P] * a + C2[p + 4 * P] * b;
        EG[p + P * (l + 3)] = gnw;
        EG[p + P * l] = en;
      }
  return 0;
}
// maxwell_solvers.f90:62-96 ; C1,C2 complex (nkx,nkr,nm,3)
int oracle_maxwell_push_wo_spchrg(double* EG_, const double* j_, const double* C1_, const double* C2_,
                                  i64 nkx, i64 nkr, i64 nm) {
  cd* EG = (cd*)EG_;
  const cd *J = (const cd*)j_, *C1 = (const cd*)C1_, *C2 = (const cd*)C2_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 q = 0; q < nkr * nm; ++q)
    for (int l = 0; l < 3; ++l)
      for (i64 i = 0; i < nkx; ++i) {
        const i64 p = i + nkx * q;
        const cd e = EG[p + P * l], g = EG[p + P * (l + 3)], j = J[p + P * l];
        cd en = C1[p] * e + C1[p + P] * g + C1[p + 2 * P] * j;
        cd gnw = C2[p] * e + C2[p + P] * g + C2[p + 2 * P] * j;
        EG[p + P * (l + 3)] = gnw;
        EG[p + P * l] = en;
      }
  return 0;
}
// maxwell_solvers.f90:98-129 ; C1,C2 complex (nkx,nkr,nm,2)
int oracle_maxwell_init_push(double* EG_, const double* j_, const double* gn_, const double* C1_,
                             const double* C2_, i64 nkx, i64 nkr, i64 nm) {
  cd* EG = (cd*)EG_;
  const cd *J = (const cd*)j_, *gn = (const cd*)gn_, *C1 = (const cd*)C1_, *C2 = (const cd*)C2_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 q = 0; q < nkr * nm; ++q)
    for (int l = 0; l < 3; ++l)
      for (i64 i = 0; i < nkx; ++i) {
        const i64 p = i + nkx * q;
        const cd j = J[p + P * l], a = gn[p + P * l];
        EG[p + P * l] = EG[p + P * l] + C1[p] * j + C1[p + P] * a;
        EG[p + P * (l + 3)] = EG[p + P * (l + 3)] + C2[p] * j + C2[p + P] * a;
      }
  return 0;
}
// maxwell_solvers.f90:131-164
int oracle_poiss_corr(double* j_, const double* gdj_, const double* gn_, const double* gnp1_, double dt_inv,
                      const double* w2_inv, i64 nkx, i64 nkr, i64 nm) {
  cd* J = (cd*)j_;
  const cd *gdj = (const cd*)gdj_, *gn = (const cd*)gn_, *gp = (const cd*)gnp1_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 q = 0; q < nkr * nm; ++q)
    for (int l = 0; l < 3; ++l)
      for (i64 i = 0; i < nkx; ++i) {
        const i64 p = i + nkx * q;
        J[p + P * l] = J[p + P * l] + (gdj[p + P * l] + (gp[p + P * l] - gn[p + P * l]) * dt_inv) * w2_inv[p];
      }
  return 0;
}
// maxwell_solvers.f90:166-197 ; DT complex (nkx)
int oracle_poiss_corr_stat(double* j_, const double* gdj_, const double* gn_, const double* DT_,
                           const double* w2_inv, i64 nkx, i64 nkr, i64 nm) {
  cd* J = (cd*)j_;
  const cd *gdj = (const cd*)gdj_, *gn = (const cd*)gn_, *DT = (const cd*)DT_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 q = 0; q < nkr * nm; ++q)
    for (int l = 0; l < 3; ++l)
      for (i64 i = 0; i < nkx; ++i) {
        const i64 p = i + nkx * q;
        J[p + P * l] = J[p + P * l] + (gdj[p + P * l] + DT[i] * gn[p + P * l]) * w2_inv[p];
      }
  return 0;
}
// maxwell_solvers.f90:199-226
int oracle_field_drift(double* EG_, const double* kx, double beta0, double dt, i64 nkx, i64 nkr, i64 nm) {
  cd* EG = (cd*)EG_;
  std::vector<cd> prop(nkx);
  for (i64 i = 0; i < nkx; ++i) {
    const double a = -0.5 * dt * beta0 * kx[i];  // EXP(fact*kx), fact = -0.5 i dt beta0
    prop[i] = cd(std::cos(a), std::sin(a));
  }
#pragma omp parallel for schedule(static)
  for (i64 q = 0; q < nkr * nm * 6; ++q)
    for (i64 i = 0; i < nkx; ++i) EG[i + nkx * q] = EG[i + nkx * q] * prop[i];
  return 0;
}
// maxwell_solvers.f90:228-250
int oracle_omp_mult_vec(double* v_, const double* A, i64 nkx, i64 nkr, i64 nm) {
  cd* v = (cd*)v_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < P; ++p)
    for (int l = 0; l < 3; ++l) v[p + P * l] = v[p + P * l] * A[p];
  return 0;
}
// maxwell_solvers.f90:252-272
int oracle_omp_mult_scl(double* v_, const double* A, i64 nkx, i64 nkr, i64 nm) {
  cd* v = (cd*)v_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < P; ++p) v[p] = v[p] * A[p];
  return 0;
}
// maxwell_solvers.f90:274-296
int oracle_omp_add_vec(double* v_, const double* A_, i64 nkx, i64 nkr, i64 nm) {
  cd* v = (cd*)v_;
  const cd* A = (const cd*)A_;
  const i64 P = nkx * nkr * nm * 3;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < P; ++p) v[p] = v[p] + A[p];
  return 0;
}
// maxwell_solvers.f90:298-318
int oracle_omp_add_scl(double* v_, const double* A_, i64 nkx, i64 nkr, i64 nm) {
  cd* v = (cd*)v_;
  const cd* A = (const cd*)A_;
  const i64 P = nkx * nkr * nm;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < P; ++p) v[p] = v[p] + A[p];
  return 0;
}

// =======================================================================================
// devices.f90:162-203  analytic planar undulator with linear entry/exit tapers (NEXT-1 row)
// =======================================================================================
int oracle_undul_analytic(const double* coord, double* Fld, double t, const double* params, i64 np) {
  (void)t;
  const double pi = 4.0 * std::atan(1.0);
  const double a0 = params[0], lambda = params[1], X0 = params[2], Lx = params[3];
  const double ku = 2.0 * pi / lambda;
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) {
    const double x = coord[3 * ip], y = coord[3 * ip + 1];
    double ampl;
    if (x <= X0 || x >= X0 + Lx) ampl = 0.0;
    else if (x > X0 && x < X0 + lambda) ampl = (x - X0) / lambda;
    else if (x > X0 + Lx - lambda && x < X0 + Lx) ampl = (X0 + Lx - x) / lambda;
    else ampl = 1.0;
    ampl = ampl * a0;
    Fld[6 * ip + 4] += ampl * std::sin(ku * (x - X0)) * std::cosh(ku * y);
    Fld[6 * ip + 3] += ampl * std::cos(ku * (x - X0)) * std::sinh(ku * y);
  }
  return 0;
}

// devices.f90:117-160  the same with a linear amplitude taper along the undulator
int oracle_undul_analytic_taper(const double* coord, double* Fld, double t, const double* params, i64 np) {
  (void)t;
  const double pi = 4.0 * std::atan(1.0);
  const double a0 = params[0], lambda = params[1], X0 = params[2], Lx = params[3], taper = params[4];
  const double ku = 2.0 * pi / lambda;
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) {
    const double x = coord[3 * ip], y = coord[3 * ip + 1];
    double ampl;
    if (x <= X0 || x >= X0 + Lx) ampl = 0.0;
    else if (x > X0 && x < X0 + lambda) ampl = (x - X0) / lambda;
    else if (x > X0 + Lx - lambda && x < X0 + Lx) ampl = (X0 + Lx - x) / lambda;
    else ampl = 1.0;
    ampl = ampl * (1 + taper * (x - X0 - 0.5 * Lx) / (0.5 * Lx));
    ampl = ampl * a0;
    Fld[6 * ip + 4] += ampl * std::sin(ku * (x - X0)) * std::cosh(ku * y);
    Fld[6 * ip + 3] += ampl * std::cos(ku * (x - X0)) * std::sinh(ku * y);
  }
  return 0;
}

// devices.f90:18-62 (tap = 0) and :64-115 (tap = 1): undulator field from a tabulated on-axis map a0(2, nx) with
// node k (1-based) at Xleft + k dx, quadratic-spline weights around the nearest node.
// Q12: for Xleft + dx <= x < Xleft + 1.5 dx the nearest node is 1 and the reference reads a0(:, 0), one column
// before the array; that column is taken as zero here (and on the GPU).
static int undul_mapped_any(const double* coord, double* Fld, const double* a0, const double* params, i64 np, i64 nx,
                            int tap) {
  const double pi = 4.0 * std::atan(1.0);
  const double lambda = params[0], Xleft = params[1], dx = params[2];
  const double Lx = tap ? params[3] : 0.0, taper = tap ? params[4] : 0.0;
  const double ku = 2.0 * pi / lambda, dx_inv = 1.0 / dx, Xright = Xleft + nx * dx;
  const double x_shift = 0.5 * (nx * dx - Lx);
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) {
    const double xp = coord[3 * ip];
    if (xp < Xleft + dx || xp > Xright - dx) continue;
    const double yp = coord[3 * ip + 1];
    const i64 ix = (i64)std::floor((xp - Xleft) * dx_inv + 0.5);
    const double ddx = (xp - Xleft) * dx_inv - (double)ix;
    const double S0[3] = {0.5 * (0.5 - ddx) * (0.5 - ddx), 0.75 - ddx * ddx, 0.5 * (0.5 + ddx) * (0.5 + ddx)};
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < 3; ++j) {
      const i64 k = ix - 1 + j;  // 1-based node
      if (k < 1 || k > nx) continue;
      s1 += S0[j] * a0[2 * (k - 1)];
      s2 += S0[j] * a0[2 * (k - 1) + 1];
    }
    double amp = 1.0;
    if (tap) amp = 1 + taper * (xp - x_shift - Xleft - 0.5 * Lx) / (0.5 * Lx);
    Fld[6 * ip + 4] += amp * s1 * std::cosh(ku * yp);
    Fld[6 * ip + 3] += amp * s2 * std::sinh(ku * yp);
  }
  return 0;
}
int oracle_undul_mapped(const double* coord, double* Fld, double t, const double* a0, const double* params, i64 np,
                        i64 nx) {
  (void)t;
  return undul_mapped_any(coord, Fld, a0, params, np, nx, 0);
}
int oracle_undul_mapped_tap(const double* coord, double* Fld, double t, const double* a0, const double* params, i64 np,
                            i64 nx) {
  (void)t;
  return undul_mapped_any(coord, Fld, a0, params, np, nx, 1);
}

// devices.f90:205-251  linearly polarised plane wave at an angle theta in the x-y plane, with linear ramps
int oracle_planewave(const double* coord, double* Fld, double t, const double* params, i64 np) {
  const double pi = 4.0 * std::atan(1.0);
  const double a0 = params[0], lambda = params[1], X0 = params[2], Lx = params[3], ramp = params[4];
  const double theta = params[5], phi0 = params[6];
  const double sinth = std::sin(theta), costh = std::cos(theta), k0 = 2.0 * pi / lambda;
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) {
    const double x = coord[3 * ip], y = coord[3 * ip + 1];
    double ampl;
    if (x <= X0 || x >= X0 + Lx) ampl = 0.0;
    else if (x > X0 && x < X0 + ramp) ampl = (x - X0) / ramp;
    else if (x > X0 + Lx - ramp && x < X0 + Lx) ampl = (X0 + Lx - x) / ramp;
    else ampl = 1.0;
    ampl = a0 * ampl * std::sin(k0 * (x * costh + y * sinth - t) + phi0);
    Fld[6 * ip + 2] += ampl;
    Fld[6 * ip + 3] += ampl * sinth;
    Fld[6 * ip + 4] -= ampl * costh;
  }
  return 0;
}

// devices.f90:253-297  Gaussian wave packet travelling along +-x (axis = +-1)
int oracle_gaussbeam(const double* coord, double* Fld, double time, double a0, const double* params, i64 np) {
  const double pi = 4.0 * std::atan(1.0);
  const double lambda = params[0], axis = params[1], x0 = params[2], y0 = params[3], z0 = params[4];
  const double Lx2_inv = 1.0 / (params[5] * params[5]), Ly2_inv = 1.0 / (params[6] * params[6]),
               Lz2_inv = 1.0 / (params[7] * params[7]);
#pragma omp parallel for schedule(static)
  for (i64 ip = 0; ip < np; ++ip) {
    const double xp = coord[3 * ip] - x0 - axis * time, yp = coord[3 * ip + 1] - y0, zp = coord[3 * ip + 2] - z0;
    const double E = a0 * std::exp(-xp * xp * Lx2_inv - yp * yp * Ly2_inv - zp * zp * Lz2_inv) *
                     std::sin(2.0 * pi / lambda * xp);
    Fld[6 * ip + 2] += E;
    Fld[6 * ip + 4] -= axis * E;
  }
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// bench.py's CPU arm sets the team size itself: torchrun exports OMP_NUM_THREADS=1 to its workers
int oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

}  // extern "C"
