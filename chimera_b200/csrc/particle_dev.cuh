// particle_dev.cuh -- device helpers shared by the particle kernels (particles.cu, particles_sorted.cu):
// Boris rotation (particle_tools.f90:18-56), linear shape factors (grid_deps.f90:46-53), the
// per-particle field gather (grid_deps.f90:149-217, grid_deps_env.f90:164-238) and the chunk-edge
// rule of the *_chnk deposits (grid_deps_chnk.f90:95-115).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace chb {

// 1 / a for finite a >= 1: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~20 bits) and two Newton steps -> <= 1 ulp,
// 5 instructions instead of the ~30 of an IEEE division with its range fix-ups
__device__ __forceinline__ double rcp_ge1(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

// Boris rotation (particle_tools.f90:18-56).  t = dt/2 B / gamma and s = 2 t / (1 + t^2) are formed with ONE
// reciprocal each (the reference divides component by component; gfortran -ffast-math, the reference's own
// build flag, makes the same transformation), 1/gamma through rsqrt and 1/(1+t^2) through rcp_ge1.
__device__ __forceinline__ void boris(double& px, double& py, double& pz, double ex, double ey, double ez,
                                      double bx, double by, double bz, double dt_2) {
  const double umx = px + dt_2 * ex, umy = py + dt_2 * ey, umz = pz + dt_2 * ez;
  const double ginv = dt_2 * rsqrt(1.0 + (umx * umx + umy * umy + umz * umz));
  const double tx = bx * ginv, ty = by * ginv, tz = bz * ginv;
  const double sfac = 2.0 * rcp_ge1(1.0 + (tx * tx + ty * ty + tz * tz));
  const double sx = tx * sfac, sy = ty * sfac, sz = tz * sfac;
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


struct Shape {
  i64 ix, ir;
  double sx0, sx1, sr0, sr1;
  double rp;
};

// leftX: position of node 0 of the grid the shape is taken on (g.leftX, or the moved window's, chimera_main.py:286)
__device__ __forceinline__ bool make_shape_at(const GridGeom& g, double leftX, double xp, double yp, double zp, Shape& s) {
  s.rp = sqrt(yp * yp + zp * zp);
  if (s.rp >= g.rmax) return false;
  const double xs = (xp - leftX) * g.dx_inv;
  s.ix = (i64)floor(xs);
  s.ir = (i64)floor((s.rp - g.r0) * g.dr_inv);
  if (s.ir < 0 || s.ir > g.nrn - 2) return false;
  s.sx1 = xs - (double)s.ix;
  s.sx0 = 1.0 - s.sx1;
  s.sr1 = (s.rp - __ldg(g.Rgrid + s.ir)) * g.dr_inv;
  s.sr0 = 1.0 - s.sr1;
  return true;
}
__device__ __forceinline__ bool make_shape(const GridGeom& g, double xp, double yp, double zp, Shape& s) {
  return make_shape_at(g, g.leftX, xp, yp, zp, s);
}


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

// One particle deposited straight to the grid with red.global.add (grid_deps.f90:18-147 and the
// _chnk/_env variants): the body of the direct kernel and the fallback of the binned kernel for
// particles outside a CTA's cell box.  `chunk` is the x-chunk the particle is filed under.
template <int ENV, int CURR>
__device__ __forceinline__ void deposit_one(const GridGeom& g, const ChunkSpec& ch, int chunk, cd* __restrict__ grid,
                                            double xp, double yp, double zp, double p0, double p1, double p2,
                                            double wp) {
  Shape s;
  if (!make_shape(g, xp, yp, zp, s)) return;
  double v[3] = {1.0, 1.0, 1.0};
  if (CURR) {
    v[0] = p0; v[1] = p1; v[2] = p2;
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

}  // namespace chb
