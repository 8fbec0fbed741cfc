// diagnostics.cu -- reductions behind the reference's integrated diagnostics (moduls/diagnostics.py), NEXT-3 row:
//   nrg_out             diagnostics.py:109-124  field energy per kx from EG_fb and the EnergyFact table
//   get_beam_envelops   diagnostics.py:174-207  weighted first / second moments of a species
//   energy spectrum     weighted histogram of gamma or p_x (the demos' np.histogram(..., weights=w))
//   line-outs           one (r node, mode, component) column of a grid array (on-axis wake amplitude)
// All of them read arrays that live in HBM and return a few KB to the host instead of the GB-sized state.
#include "common.cuh"
#include "kernels.cuh"

namespace chb {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[kx] += sum over a slice of the (kr, m) columns of fact[kx, col] * sum_{c<3} |EG[kx, col, c]|^2.
// x is the fastest index: a warp reads 32 consecutive kx of one column (coalesced); blockIdx.y strides the columns.
__global__ void __launch_bounds__(256) field_energy_k(const cd* __restrict__ EG, const double* __restrict__ fact,
                                                      double* __restrict__ out, i64 nkx, i64 ncols) {
  const i64 ix = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ix >= nkx) return;
  const i64 plane = nkx * ncols;
  double acc = 0.0;
  for (i64 col = blockIdx.y; col < ncols; col += gridDim.y) {
    const i64 o = ix + nkx * col;
    const cd a = __ldg(EG + o), b = __ldg(EG + o + plane), c = __ldg(EG + o + 2 * plane);
    acc += __ldg(fact + o) * (a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y);
  }
  atomicAdd(out + ix, acc);
}

// 16 sums over the particles: w; then per axis c: w x, w x^2, w p^2, w x p, w p
__global__ void __launch_bounds__(256) beam_moments_k(const double* __restrict__ x, const double* __restrict__ p,
                                                      const double* __restrict__ w, i64 cap, i64 np,
                                                      double* __restrict__ out) {
  double s[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = 0.0;
  for (i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += (i64)gridDim.x * blockDim.x) {
    const double wp = __ldg(w + ip);
    s[0] += wp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double xc = __ldg(x + c * cap + ip), pc = __ldg(p + c * cap + ip);
      s[1 + 5 * c] += wp * xc;
      s[2 + 5 * c] += wp * xc * xc;
      s[3 + 5 * c] += wp * pc * pc;
      s[4 + 5 * c] += wp * xc * pc;
      s[5 + 5 * c] += wp * pc;
    }
  }
  __shared__ double red[8][16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const double v = warp_sum(s[i]);
    if (lane == 0) red[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0.0;
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    atomicAdd(out + threadIdx.x, v);
  }
}

// weighted histogram of gamma = sqrt(1 + p^2) (quantity 0) or p_x (1): bin = floor((q - lo) / (hi - lo) * nbins),
// values outside [lo, hi) are dropped except q == hi, which joins the last bin (numpy.histogram's rule)
__global__ void __launch_bounds__(256) spectrum_k(const double* __restrict__ p, const double* __restrict__ w, i64 cap,
                                                  i64 np, int quantity, double lo, double hi, int nbins,
                                                  double* __restrict__ hist) {
  extern __shared__ double sh[];
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const double scale = (double)nbins / (hi - lo);
  for (i64 ip = (i64)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += (i64)gridDim.x * blockDim.x) {
    const double px = __ldg(p + ip), py = __ldg(p + cap + ip), pz = __ldg(p + 2 * cap + ip);
    const double q = quantity == 0 ? sqrt(1.0 + (px * px + py * py + pz * pz)) : px;
    if (q < lo || q > hi) continue;
    int b = (int)floor((q - lo) * scale);
    if (b >= nbins) b = nbins - 1;
    atomicAdd(&sh[b], __ldg(w + ip));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh[i] != 0.0) atomicAdd(hist + i, sh[i]);
}

__global__ void __launch_bounds__(256) lineout_k(const cd* __restrict__ A, cd* __restrict__ out, i64 nx, i64 offset) {
  const i64 ix = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ix < nx) out[ix] = __ldg(A + offset + ix);
}

}  // namespace

int launch_field_energy(cudaStream_t st, const cd* EG, const double* fact, double* out, i64 nkx, i64 ncols) {
  CHB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * nkx, st));
  i64 gy = (148 * 8 * 256) / (nkx > 256 ? nkx : 256);  // ~8 CTAs per SM over the whole array
  gy = gy < 1 ? 1 : (gy > ncols ? ncols : gy);
  field_energy_k<<<dim3(grid_for(nkx, 256), (unsigned)gy), 256, 0, st>>>(EG, fact, out, nkx, ncols);
  CHB_LAUNCH_CHECK();
  return 0;
}

int launch_beam_moments(cudaStream_t st, const double* x, const double* p, const double* w, i64 cap, i64 np, double* out16) {
  CHB_CUDA(cudaMemsetAsync(out16, 0, sizeof(double) * 16, st));
  if (np <= 0) return 0;
  const i64 want = (np + 255) / 256;
  beam_moments_k<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(x, p, w, cap, np, out16);
  CHB_LAUNCH_CHECK();
  return 0;
}

int launch_spectrum(cudaStream_t st, const double* p, const double* w, i64 cap, i64 np, int quantity, double lo, double hi,
                    int nbins, double* hist) {
  CHB_CUDA(cudaMemsetAsync(hist, 0, sizeof(double) * nbins, st));
  if (np <= 0) return 0;
  const i64 want = (np + 255) / 256;
  spectrum_k<<<(unsigned)(want < 148 * 4 ? want : 148 * 4), 256, sizeof(double) * nbins, st>>>(p, w, cap, np, quantity, lo,
                                                                                              hi, nbins, hist);
  CHB_LAUNCH_CHECK();
  return 0;
}

int launch_lineout(cudaStream_t st, const cd* A, cd* out, i64 nx, i64 offset) {
  lineout_k<<<grid_for(nx, 256), 256, 0, st>>>(A, out, nx, offset);
  CHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace chb
