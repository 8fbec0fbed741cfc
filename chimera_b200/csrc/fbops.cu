// fbops.cu -- see fbops.cuh.  Reference: f90/fb_io.f90, f90/fb_math.f90, f90/fb_math_env.f90.
#include "fbops.cuh"

namespace chb {

// ------------------------------------------------------------------------------------------
static size_t round_mb(size_t b) { return (b + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1); }

int Scratch::reserve(size_t bytes) {
  for (auto& b : blocks)
    if (b.cap - b.used >= bytes + 256) return 0;
  Block nb{nullptr, round_mb(bytes + 256), 0};
  CHB_CUDA(cudaMalloc((void**)&nb.p, nb.cap));
  blocks.push_back(nb);
  ++gen;
  return 0;
}
void* Scratch::take(size_t bytes) {
  for (int pass = 0; pass < 2; ++pass) {
    if (!blocks.empty()) {
      Block& b = blocks.back();
      const size_t a = (b.used + 255) & ~size_t(255);
      if (a + bytes <= b.cap) {
        b.used = a + bytes;
        size_t tot = 0;
        for (auto& q : blocks) tot += q.used;
        if (tot > high) high = tot;
        return b.p + a;
      }
    }
    if (pass == 0) {
      size_t want = bytes + 256;
      if (want < (size_t(32) << 20)) want = size_t(32) << 20;
      Block nb{nullptr, round_mb(want), 0};
      if (cudaMalloc((void**)&nb.p, nb.cap) != cudaSuccess) {
        set_error("scratch: cudaMalloc of %zu bytes failed", nb.cap);
        return nullptr;
      }
      blocks.push_back(nb);
      ++gen;
    }
  }
  return nullptr;
}
void Scratch::reset() {
  if (blocks.size() > 1) {  // coalesce so the next round fits in one block
    size_t tot = 0;
    for (auto& b : blocks) { tot += b.cap; cudaFree(b.p); }
    blocks.clear();
    ++gen;
    Block nb{nullptr, round_mb(tot), 0};
    if (cudaMalloc((void**)&nb.p, nb.cap) == cudaSuccess) blocks.push_back(nb);
  }
  for (auto& b : blocks) b.used = 0;
}
void Scratch::destroy() {
  for (auto& b : blocks) cudaFree(b.p);
  blocks.clear();
}

int FFTCache::exec(cudaStream_t st, cd* data, i64 n, i64 batch, int dir) { return exec(st, data, data, n, batch, dir); }
int FFTCache::exec(cudaStream_t st, cd* in, cd* out, i64 n, i64 batch, int dir) {
  if (n <= 0 || batch <= 0) return 0;
  auto key = std::make_pair(n, batch);
  auto it = plans.find(key);
  if (it == plans.end()) {
    cufftHandle h;
    int nn[1] = {(int)n};
    cufftResult r = cufftPlanMany(&h, 1, nn, nullptr, 1, (int)n, nullptr, 1, (int)n, CUFFT_Z2Z, (int)batch);
    if (r != CUFFT_SUCCESS) { set_error("cufftPlanMany(n=%lld,batch=%lld) failed: %d", n, batch, (int)r); return 7; }
    it = plans.emplace(key, h).first;
  }
  cufftResult r = cufftSetStream(it->second, st);
  if (r == CUFFT_SUCCESS) r = cufftExecZ2Z(it->second, (cufftDoubleComplex*)in, (cufftDoubleComplex*)out, dir);
  if (r != CUFFT_SUCCESS) { set_error("cufftExecZ2Z failed: %d", (int)r); return 7; }
  return 0;
}
void FFTCache::destroy() {
  for (auto& kv : plans) cufftDestroy(kv.second);
  plans.clear();
}

size_t packed_ops_bytes(i64 K, i64 N, int nslots) { return (size_t)gemm_packed_size(K, N) * sizeof(double) * nslots; }

int pack_ops(cudaStream_t st, PackedOps& out, double* dst, const double* op, i64 K, i64 N, int nslots) {
  if (nslots > (int)(sizeof(out.slot) / sizeof(out.slot[0]))) { set_error("too many operator slots"); return 8; }
  out.K = K; out.N = N; out.nslots = nslots;
  const i64 sz = gemm_packed_size(K, N);
  for (int s = 0; s < nslots; ++s) {
    CHB_TRY(launch_gemm_pack_b(st, dst + sz * s, op + K * N * s, K, N, K));
    out.slot[s] = dst + sz * s;
  }
  return 0;
}

namespace {
struct Batcher {
  cudaStream_t st;
  i64 M, N, K, lda, ldc;
  GemmBatch b;
  int rc = 0;
  Batcher(cudaStream_t s, i64 M_, i64 N_, i64 K_, i64 lda_, i64 ldc_) : st(s), M(M_), N(N_), K(K_), lda(lda_), ldc(ldc_) { b.count = 0; }
  // every C of the batch leaves the kernel row-scaled by exp(i sign leftX kx) (and by the `fact` given to add)
  void phase(const double* kx, double leftX, double sign) { b.phase_kx = kx; b.phase_leftX = leftX; b.phase_sign = sign; }
  void add(const cd* A, const double* Bp, cd* C, double alpha, double beta, const double* fact = nullptr) {
    if (rc) return;
    if (b.count == kGemmMaxBatch) flush();
    b.p[b.count++] = GemmProblem{(const double*)A, Bp, (double*)C, alpha, beta, fact};
  }
  int flush() {
    if (!rc && b.count) rc = launch_gemm(st, b, M, N, K, lda, ldc);
    b.count = 0;
    return rc;
  }
};
}  // namespace

// ------------------------------------------------------------------------------------------
int fb_in_dev(FBCtx& c, cd* out_fb, const cd* in, double leftX, const double* kx, const PackedOps& In,
              const double* fact, i64 nkx, i64 nrn, i64 nm, i64 nkr, int ncomp) {
  // x-FFT first (out of place, into scratch), then the Hankel contraction with shiftX_inv and the quadrature factor
  // (fb_io.f90:52-57) in its epilogue: the two transforms act on different axes, and this order leaves no separate
  // pass over the spectral array
  const i64 nr = nrn - 1, ncols = nrn * nm * ncomp;
  cd* xf = c.scr->take_n<cd>(nkx * ncols);
  if (!xf) return 6;
  CHB_TRY(c.fft->exec(c.st, const_cast<cd*>(in), xf, nkx, ncols, CUFFT_FORWARD));
  Batcher gb(c.st, 2 * nkx, nkr, nr, 2 * nkx, 2 * nkx);
  gb.phase(kx, leftX, -1.0);
  for (int l = 0; l < ncomp; ++l)
    for (i64 m = 0; m < nm; ++m)
      gb.add(xf + nkx * (1 + nrn * (m + nm * l)), In.slot[m], out_fb + nkx * nkr * (m + nm * l), 1.0, 0.0,
             fact ? fact + nkx * nkr * m : nullptr);
  CHB_TRY(gb.flush());
  return 0;
}

int fb_out_dev(FBCtx& c, cd* out, const cd* const* srcs, int nsrc, int ncomp_each, double leftX, const double* kx,
               const PackedOps& Out, i64 nkx, i64 nrn, i64 nm, i64 nkr) {
  const i64 nr = nrn - 1;
  const int ncomp = nsrc * ncomp_each;
  // the contractions overwrite radial nodes 1..nr of every plane; only the ghost node 0 is left to zero (fb_io.f90:203)
  CHB_CUDA(cudaMemset2DAsync(out, sizeof(cd) * nkx * nrn, 0, sizeof(cd) * nkx, (size_t)(nm * ncomp), c.st));
  Batcher gb(c.st, 2 * nkx, nr, nkr, 2 * nkx, 2 * nkx);
  gb.phase(kx, leftX, +1.0);  // shiftX (fb_io.f90:216-222) in the epilogue; the ghost node stays zero either way
  for (int j = 0; j < nsrc; ++j)
    for (int l = 0; l < ncomp_each; ++l)
      for (i64 m = 0; m < nm; ++m)
        gb.add(srcs[j] + nkx * nkr * (m + nm * l), Out.slot[m],
               out + nkx * (1 + nrn * (m + nm * (l + ncomp_each * j))), 1.0, 0.0);
  CHB_TRY(gb.flush());
  CHB_TRY(c.fft->exec(c.st, out, nkx, nrn * nm * ncomp, CUFFT_INVERSE));
  return 0;
}

namespace {
__global__ void __launch_bounds__(256) take_rows_k(cd* __restrict__ dst, const cd* __restrict__ src,
                                                   const i64* __restrict__ rows, i64 nxs, i64 nkx, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const i64 col = e / nxs, j = e - col * nxs;
  dst[e] = __ldg(src + __ldg(rows + j) + nkx * col);
}
// out(i, col) = gathered[rank r][j, col] with map[i] = r * nxs + j; each rank's block is (nxs, ncols)
__global__ void __launch_bounds__(256) put_rows_k(cd* __restrict__ out, const cd* __restrict__ gath,
                                                  const i64* __restrict__ map, i64 nkx, i64 nxs, i64 ncols, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const i64 col = e / nkx, i = e - col * nkx;
  const i64 m = __ldg(map + i), r = m / nxs, j = m - r * nxs;
  out[e] = __ldg(gath + r * nxs * ncols + j + nxs * col);
}
}  // namespace

int fb_in_slab_dev(FBCtx& c, cd* out_fb, const cd* in, double leftX, const double* kx_slab, const PackedOps& In,
                   const double* fact_slab, const i64* rows, i64 nkx, i64 nxs, i64 nrn, i64 nm, i64 nkr, int ncomp) {
  const i64 nr = nrn - 1, ncols = nrn * nm * ncomp;
  cd* full = c.scr->take_n<cd>(nkx * ncols);
  cd* slab = c.scr->take_n<cd>(nxs * ncols);
  if (!full || !slab) return 6;
  CHB_CUDA(cudaMemcpyAsync(full, in, sizeof(cd) * nkx * ncols, cudaMemcpyDeviceToDevice, c.st));
  CHB_TRY(c.fft->exec(c.st, full, nkx, ncols, CUFFT_FORWARD));
  take_rows_k<<<grid_for(nxs * ncols, 256), 256, 0, c.st>>>(slab, full, rows, nxs, nkx, nxs * ncols);
  CHB_LAUNCH_CHECK();
  Batcher gb(c.st, 2 * nxs, nkr, nr, 2 * nxs, 2 * nxs);
  gb.phase(kx_slab, leftX, -1.0);  // shiftX_inv and the quadrature factor (fb_io.f90:52-57) in the epilogue
  for (int l = 0; l < ncomp; ++l)
    for (i64 m = 0; m < nm; ++m)
      gb.add(slab + nxs * (1 + nrn * (m + nm * l)), In.slot[m], out_fb + nxs * nkr * (m + nm * l), 1.0, 0.0,
             fact_slab ? fact_slab + nxs * nkr * m : nullptr);
  CHB_TRY(gb.flush());
  return 0;
}

// the second half of fb_in_slab_dev on a slab that is already x-transformed (column-block dataflow: the x-FFT ran on
// this rank's column block and the rows came in through the all-to-all)
int fb_in_slab_post_dev(FBCtx& c, cd* out_fb, const cd* slab, double leftX, const double* kx_slab, const PackedOps& In,
                        const double* fact_slab, i64 nxs, i64 nrn, i64 nm, i64 nkr, int ncomp) {
  const i64 nr = nrn - 1;
  Batcher gb(c.st, 2 * nxs, nkr, nr, 2 * nxs, 2 * nxs);
  gb.phase(kx_slab, leftX, -1.0);  // shiftX_inv and the quadrature factor (fb_io.f90:52-57) in the epilogue
  for (int l = 0; l < ncomp; ++l)
    for (i64 m = 0; m < nm; ++m)
      gb.add(slab + nxs * (1 + nrn * (m + nm * l)), In.slot[m], out_fb + nxs * nkr * (m + nm * l), 1.0, 0.0,
             fact_slab ? fact_slab + nxs * nkr * m : nullptr);
  CHB_TRY(gb.flush());
  return 0;
}

namespace {
// send[(r * ncols + col) * nxs + j] = blk[i + nkx * col] with map[i] = r * nxs + j: the rows of every rank's kx slab,
// contiguous per destination rank
__global__ void __launch_bounds__(256) pack_cols_k(cd* __restrict__ send, const cd* __restrict__ blk,
                                                   const i64* __restrict__ map, i64 nkx, i64 nxs, i64 ncols, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const i64 col = e / nkx, i = e - col * nkx;
  const i64 m = __ldg(map + i), r = m / nxs, j = m - r * nxs;
  send[(r * ncols + col) * nxs + j] = blk[e];
}
}  // namespace

// column-block dataflow, forward: x-FFT of this rank's column block (nkx, ncols) in place, then the rows sorted by
// destination rank for the all-to-all
int col_fwd_dev(FBCtx& c, cd* send, cd* blk, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols) {
  CHB_TRY(c.fft->exec(c.st, blk, nkx, ncols, CUFFT_FORWARD));
  pack_cols_k<<<grid_for(nkx * ncols, 256), 256, 0, c.st>>>(send, blk, gather_map, nkx, nxs, ncols, nkx * ncols);
  CHB_LAUNCH_CHECK();
  return 0;
}

int fb_out_slab_dev(FBCtx& c, cd* out_slab, const cd* const* srcs, int nsrc, int ncomp_each, double leftX,
                    const double* kx_slab, const PackedOps& Out, i64 nxs, i64 nrn, i64 nm, i64 nkr) {
  const i64 nr = nrn - 1;
  const int ncomp = nsrc * ncomp_each;
  CHB_CUDA(cudaMemset2DAsync(out_slab, sizeof(cd) * nxs * nrn, 0, sizeof(cd) * nxs, (size_t)(nm * ncomp), c.st));
  Batcher gb(c.st, 2 * nxs, nr, nkr, 2 * nxs, 2 * nxs);
  gb.phase(kx_slab, leftX, +1.0);
  for (int j = 0; j < nsrc; ++j)
    for (int l = 0; l < ncomp_each; ++l)
      for (i64 m = 0; m < nm; ++m)
        gb.add(srcs[j] + nxs * nkr * (m + nm * l), Out.slot[m],
               out_slab + nxs * (1 + nrn * (m + nm * (l + ncomp_each * j))), 1.0, 0.0);
  CHB_TRY(gb.flush());
  return 0;
}

// full(i, col) <- gathered[rank][j, col] (map[i] = rank * nxs + j) and slab(j, col) <- full(rows[j], col)
int rows_scatter_dev(FBCtx& c, cd* full, const cd* gathered, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols) {
  put_rows_k<<<grid_for(nkx * ncols, 256), 256, 0, c.st>>>(full, gathered, gather_map, nkx, nxs, ncols, nkx * ncols);
  CHB_LAUNCH_CHECK();
  return 0;
}
int rows_take_dev(FBCtx& c, cd* slab, const cd* full, const i64* rows, i64 nkx, i64 nxs, i64 ncols) {
  take_rows_k<<<grid_for(nxs * ncols, 256), 256, 0, c.st>>>(slab, full, rows, nxs, nkx, nxs * ncols);
  CHB_LAUNCH_CHECK();
  return 0;
}

namespace {
// put_rows_k for one column block of EB with the normalisation of eb_correction (grid_deps.f90:219-266: 1/2pi for mode 0
// of the real solver, 1/pi otherwise) folded in: the factor is per column and commutes with the x-FFT that follows
__global__ void __launch_bounds__(256) put_rows_norm_k(cd* __restrict__ out, const cd* __restrict__ gath,
                                                       const i64* __restrict__ map, i64 nkx, i64 nxs, i64 ncols, i64 col0,
                                                       i64 nrn, i64 nm, int env, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const i64 col = e / nkx, i = e - col * nkx;
  const i64 m = __ldg(map + i), r = m / nxs, j = m - r * nxs;
  const int slot = (int)(((col0 + col) / nrn) % nm);
  const double pi_inv = 1.0 / 3.14159265358979323846;
  const double f = (!env && slot == 0) ? 0.5 * pi_inv : pi_inv;
  const cd v = __ldg(gath + r * nxs * ncols + j + nxs * col);
  out[e] = cmake(f * v.x, f * v.y);
}
// ghost row of every (mode, component) plane from row 1 (copy, or negate: Q6)
__global__ void __launch_bounds__(256) eb_ghost_k(cd* __restrict__ eb, i64 nxn, i64 nrn, i64 nm, int env, i64 n) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const i64 q = e / nxn, ix = e - q * nxn;
  const int slot = (int)(q % nm);
  const bool negate = env ? (nm > 1) : (slot > 0);
  cd* pl = eb + nxn * nrn * q;
  const cd v = pl[ix + nxn];
  pl[ix] = negate ? cmake(-v.x, -v.y) : v;
}
}  // namespace

// column-block dataflow, backward: received (rank, column, row) blocks -> rows in place, normalised, inverse x-FFT
int col_bwd_dev(FBCtx& c, cd* blk, const cd* recv, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols, i64 col0, i64 nrn,
                i64 nm, int env) {
  put_rows_norm_k<<<grid_for(nkx * ncols, 256), 256, 0, c.st>>>(blk, recv, gather_map, nkx, nxs, ncols, col0, nrn, nm, env,
                                                                nkx * ncols);
  CHB_LAUNCH_CHECK();
  return c.fft->exec(c.st, blk, nkx, ncols, CUFFT_INVERSE);
}
int eb_ghost_dev(FBCtx& c, cd* eb, i64 nxn, i64 nrn, i64 nm, int env, int ncomp) {
  const i64 n = nxn * nm * ncomp;
  eb_ghost_k<<<grid_for(n, 256), 256, 0, c.st>>>(eb, nxn, nrn, nm, env, n);
  CHB_LAUNCH_CHECK();
  return 0;
}

int fb_out_finish_dev(FBCtx& c, cd* out, const cd* gathered, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols) {
  put_rows_k<<<grid_for(nkx * ncols, 256), 256, 0, c.st>>>(out, gathered, gather_map, nkx, nxs, ncols, nkx * ncols);
  CHB_LAUNCH_CHECK();
  return c.fft->exec(c.st, out, nkx, ncols, CUFFT_INVERSE);
}

int fb_filtr_dev(FBCtx& c, cd* vec, double leftX, const double* kx, const double* filtr, int modefilt, i64 nkx,
                 i64 nkr, i64 nm, i64 nxfilt) {
  const i64 ncols = nkr * nm * 3;
  CHB_TRY(launch_rowscale_phase(c.st, vec, kx, leftX, +1.0, 1.0, nullptr, nkx, ncols, 1));
  CHB_TRY(c.fft->exec(c.st, vec, nkx, ncols, CUFFT_INVERSE));
  CHB_TRY(launch_window(c.st, vec, filtr, modefilt, nkx, ncols, nxfilt));
  CHB_TRY(c.fft->exec(c.st, vec, nkx, ncols, CUFFT_FORWARD));
  // shiftX_inv = 1 / (nkx * shiftX) = conj(shiftX) / nkx
  CHB_TRY(launch_rowscale_phase(c.st, vec, kx, leftX, -1.0, 1.0 / (double)nkx, nullptr, nkx, ncols, 1));
  return 0;
}

// ------------------------------------------------------------------------------------------
namespace {
struct Modes {
  int env;
  i64 nm, nko, lo, hi;
  explicit Modes(const FBMathDims& d) : env(d.env), nm(d.nm) {
    nko = env ? (nm - 1) / 2 : nm - 1;
    lo = env ? -nko : 0;
    hi = nko;
  }
  i64 vslot(i64 mode) const { return mode - lo; }
  i64 dslot(i64 mode) const { return env ? mode + nko + 1 : mode; }
};
const Unit U1{1, 0}, Um1{-1, 0}, Ui{0, 1}, U0{0, 0};

// copy the leading nkr_loc columns of a (nkx, nkr) plane through i*kx: dst(:, 0:nkr_loc) (+)= sign*i*kx*src
int ikx_plane(cudaStream_t st, cd* dst, const cd* src, const double* kx, double sign, int acc, const FBMathDims& d) {
  return launch_ikx(st, dst, src, kx, sign, acc, d.nkx, d.nkr_loc);
}

// S[mode - slo] = [i kx v1[mode]] + Dm(mode).T1[mode-1] + Dp(mode).T2[mode+1],  T1 = i v3 - v2, T2 = i v3 + v2
// (fb_div fb_math.f90:151-199, fb_div_env fb_math_env.f90:63-104, first halves of fb_graddiv[_env])
int div_like(FBCtx& c, cd* S, i64 slo, i64 shi, const cd* vec, const PackedOps& Dp, const PackedOps& Dm,
             const double* kx, const FBMathDims& d) {
  const Modes mo(d);
  const i64 Pv = d.nkx * d.nkr, Ps = d.nkx * d.nkr_loc;
  const cd* v1 = vec;
  const cd* v2 = vec + Pv * d.nm;
  const cd* v3 = vec + Pv * d.nm * 2;
  cd* T1 = c.scr->take_n<cd>(Pv * d.nm);
  cd* T2 = c.scr->take_n<cd>(Pv * d.nm);
  cd* T1ext = nullptr;
  if (!T1 || !T2) return 6;
  CHB_TRY(launch_pm(c.st, T1, T2, v3, v2, Pv * d.nm));  // T1 = i v3 - v2, T2 = i v3 + v2
  if (!d.env && mo.nko > 0) {  // Q7: with nko = 0 the missing mode 1 is taken as zero
    T1ext = c.scr->take_n<cd>(Pv);
    if (!T1ext) return 6;
    CHB_TRY(launch_combine(c.st, T1ext, v3 + Pv * mo.vslot(1), Ui, v2 + Pv * mo.vslot(1), Um1, 1 + d.mirror_shift, d.nkx, d.nkr));
  }
  for (i64 mode = slo; mode <= shi; ++mode) {
    cd* o = S + Ps * (mode - slo);
    if (mode >= mo.lo && mode <= mo.hi) CHB_TRY(ikx_plane(c.st, o, v1 + Pv * mo.vslot(mode), kx, +1.0, 0, d));
    else CHB_CUDA(cudaMemsetAsync(o, 0, sizeof(cd) * Ps, c.st));
  }
  Batcher gm(c.st, 2 * d.nkx, d.nkr_loc, d.nkr, 2 * d.nkx, 2 * d.nkx);
  for (i64 mode = slo; mode <= shi; ++mode) {
    const cd* src = nullptr;
    if (!d.env) src = (mode > 0) ? T1 + Pv * mo.vslot(mode - 1) : T1ext;
    else if (mode > -mo.nko) src = T1 + Pv * mo.vslot(mode - 1);
    if (src) gm.add(src, Dm.slot[mo.dslot(mode)], S + Ps * (mode - slo), 1.0, 1.0);
  }
  CHB_TRY(gm.flush());
  Batcher gp(c.st, 2 * d.nkx, d.nkr_loc, d.nkr, 2 * d.nkx, 2 * d.nkx);
  for (i64 mode = slo; mode <= shi; ++mode)
    if (mode < mo.nko) gp.add(T2 + Pv * mo.vslot(mode + 1), Dp.slot[mo.dslot(mode)], S + Ps * (mode - slo), 1.0, 1.0);
  CHB_TRY(gp.flush());
  return 0;
}

// out(:,:,mode,1) = i kx S[mode];  G1 = Dm(mode).S[mode-1],  G2 = Dp(mode).S[mode+1];
// out(..,2) = -G1 + G2 ;  out(..,3) = i G1 + i G2
// (fb_grad fb_math.f90:96-149, fb_grad_env fb_math_env.f90:18-61, second halves of fb_graddiv[_env])
struct PoissTail {  // when given, grad_like ends in the Poisson-correction update of `out` (= J_fb) instead of storing
  const cd *gn, *gp;
  const double* w2inv;
  double dt_inv;
};
int grad_like(FBCtx& c, cd* out, const cd* S, i64 slo, i64 shi, const PackedOps& Dp, const PackedOps& Dm,
              const double* kx, const FBMathDims& d, bool always_both, const PoissTail* pt = nullptr) {
  const Modes mo(d);
  const i64 Pin = d.nkx * d.nkr, Ps = d.nkx * d.nkr_loc;
  auto Sp = [&](i64 mode) { return S + Pin * (mode - slo); };
  (void)shi;
  cd* G1 = c.scr->take_n<cd>(Ps * d.nm);
  cd* G2 = c.scr->take_n<cd>(Ps * d.nm);
  if (!G1 || !G2) return 6;
  cd* Sext = nullptr;
  if (!d.env && mo.nko > 0) {
    Sext = c.scr->take_n<cd>(Pin);
    if (!Sext) return 6;
    CHB_TRY(launch_combine(c.st, Sext, Sp(1), U1, nullptr, U0, 1 + d.mirror_shift, d.nkx, d.nkr));
  }
  Batcher gb(c.st, 2 * d.nkx, d.nkr_loc, d.nkr, 2 * d.nkx, 2 * d.nkx);
  for (i64 mode = mo.lo; mode <= mo.hi; ++mode) {
    const cd* lower = nullptr;
    if (!d.env) lower = (mode > 0) ? Sp(mode - 1) : Sext;
    else if (always_both || mode > -mo.nko) lower = Sp(mode - 1);
    // a slot no contraction writes (beta = 0 overwrites the others) has to read as zero in the tail
    if (lower) gb.add(lower, Dm.slot[mo.dslot(mode)], G1 + Ps * mo.vslot(mode), 1.0, 0.0);
    else CHB_CUDA(cudaMemsetAsync(G1 + Ps * mo.vslot(mode), 0, sizeof(cd) * Ps, c.st));
    if (always_both || mode < mo.nko)
      gb.add(Sp(mode + 1), Dp.slot[mo.dslot(mode)], G2 + Ps * mo.vslot(mode), 1.0, 0.0);
    else CHB_CUDA(cudaMemsetAsync(G2 + Ps * mo.vslot(mode), 0, sizeof(cd) * Ps, c.st));
  }
  CHB_TRY(gb.flush());
  if (pt)  // needs Ps == Pin (checked by the caller)
    return launch_grad_poiss_tail(c.st, out, Sp(mo.lo), G1, G2, pt->gn, pt->gp, kx, pt->w2inv, pt->dt_inv, d.nkx, Ps * d.nm);
  // out1 = i kx S, out2 = -G1 + G2, out3 = i G1 + i G2 in one pass
  return launch_grad_tail(c.st, out, Sp(mo.lo), G1, G2, kx, d.nkx, Ps, Pin, d.nm);
}
}  // namespace

size_t fb_math_scratch_bytes(const FBMathDims& d) {
  const size_t P = (size_t)d.nkx * (size_t)(d.nkr > d.nkr_loc ? d.nkr : d.nkr_loc) * sizeof(cd);
  return (6 * (size_t)d.nm + 12) * (P + 256);
}

int fb_grad_dev(FBCtx& c, cd* out, const cd* scl, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
                const FBMathDims& d) {
  const Modes mo(d);
  return grad_like(c, out, scl, mo.lo, mo.hi, Dp, Dm, kx, d, false);
}

int fb_div_dev(FBCtx& c, cd* out, const cd* vec, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
               const FBMathDims& d) {
  const Modes mo(d);
  return div_like(c, out, mo.lo, mo.hi, vec, Dp, Dm, kx, d);
}

// fb_graddiv (fb_math.f90:201-293) / fb_graddiv_env (fb_math_env.f90:164-233), in place
// `in` may alias `vec` (the reference's in-place form): the divergence is complete before the gradient is written
int fb_graddiv_dev(FBCtx& c, cd* vec, const cd* in, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
                   const FBMathDims& d) {
  if (d.nkr != d.nkr_loc) { set_error("fb_graddiv needs nkr == nkr_loc"); return 9; }
  const Modes mo(d);
  const i64 slo = d.env ? mo.lo - 1 : 0, shi = mo.hi + 1;
  cd* S = c.scr->take_n<cd>(d.nkx * d.nkr_loc * (shi - slo + 1));
  if (!S) return 6;
  CHB_TRY(div_like(c, S, slo, shi, in, Dp, Dm, kx, d));
  CHB_TRY(grad_like(c, vec, S, slo, shi, Dp, Dm, kx, d, true));
  return 0;
}

// one iteration of Solver.poiss_corr (solvers.py:317-326): J += PoissFact (grad div J + (gradRho_nxt - gradRho_prv)/dt)
int fb_poiss_iter_dev(FBCtx& c, cd* J, const cd* gn, const cd* gp, double dt_inv, const double* w2inv,
                      const PackedOps& Dp, const PackedOps& Dm, const double* kx, const FBMathDims& d) {
  if (d.nkr != d.nkr_loc) { set_error("fb_poiss_iter needs nkr == nkr_loc"); return 9; }
  const Modes mo(d);
  const i64 slo = d.env ? mo.lo - 1 : 0, shi = mo.hi + 1;
  cd* S = c.scr->take_n<cd>(d.nkx * d.nkr_loc * (shi - slo + 1));
  if (!S) return 6;
  CHB_TRY(div_like(c, S, slo, shi, J, Dp, Dm, kx, d));
  const PoissTail pt{gn, gp, w2inv, dt_inv};
  return grad_like(c, J, S, slo, shi, Dp, Dm, kx, d, true, &pt);
}

// fb_rot (fb_math.f90:18-94) / fb_rot_env (fb_math_env.f90:106-162)
//   out1 = -Dp.(i v2 - v3)[m+1] - Dm.(i v2 + v3)[m-1]        (env: the Dm term is dropped, Q5)
//   out2 = -i kx v3 + i GP + i GM ;  out3 = +i kx v2 - GP + GM ;  GP = Dp.v1[m+1], GM = Dm.v1[m-1]
int fb_rot_dev(FBCtx& c, cd* out, const cd* vec, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
               const FBMathDims& d) {
  const Modes mo(d);
  const i64 Pv = d.nkx * d.nkr, Ps = d.nkx * d.nkr_loc;
  const cd* v1 = vec;
  const cd* v2 = vec + Pv * d.nm;
  const cd* v3 = vec + Pv * d.nm * 2;
  cd* R1 = c.scr->take_n<cd>(Pv * d.nm);
  cd* R2 = c.scr->take_n<cd>(Pv * d.nm);
  cd* GP = c.scr->take_n<cd>(Ps * d.nm);
  cd* GM = c.scr->take_n<cd>(Ps * d.nm);
  if (!R1 || !R2 || !GP || !GM) return 6;
  if (!d.env) CHB_TRY(launch_pm(c.st, R1, R2, v2, v3, Pv * d.nm));  // R1 = i v2 - v3, R2 = i v2 + v3
  else CHB_TRY(launch_combine(c.st, R1, v2, Ui, v3, Um1, 0, d.nkx, d.nkr * d.nm));
  cd *R2ext = nullptr, *V1ext = nullptr;
  if (!d.env && mo.nko > 0) {
    R2ext = c.scr->take_n<cd>(Pv);
    V1ext = c.scr->take_n<cd>(Pv);
    if (!R2ext || !V1ext) return 6;
    CHB_TRY(launch_combine(c.st, R2ext, v2 + Pv * mo.vslot(1), Ui, v3 + Pv * mo.vslot(1), U1, 1 + d.mirror_shift, d.nkx, d.nkr));
    CHB_TRY(launch_combine(c.st, V1ext, v1 + Pv * mo.vslot(1), U1, nullptr, U0, 1 + d.mirror_shift, d.nkx, d.nkr));
  }
  // slots no beta = 0 contraction writes are zeroed; the others are overwritten
  auto zero = [&](cd* p, i64 slot) { return cudaMemsetAsync(p + Ps * slot, 0, sizeof(cd) * Ps, c.st); };
  Batcher g1(c.st, 2 * d.nkx, d.nkr_loc, d.nkr, 2 * d.nkx, 2 * d.nkx);
  for (i64 mode = mo.lo; mode <= mo.hi; ++mode) {
    const i64 s = mo.vslot(mode);
    if (mode < mo.nko) {
      g1.add(R1 + Pv * mo.vslot(mode + 1), Dp.slot[mo.dslot(mode)], out + Ps * s, -1.0, 0.0);
      g1.add(v1 + Pv * mo.vslot(mode + 1), Dp.slot[mo.dslot(mode)], GP + Ps * s, 1.0, 0.0);
    } else {
      CHB_CUDA(zero(out, s));
      CHB_CUDA(zero(GP, s));
    }
    const cd* l1 = nullptr;
    if (!d.env) l1 = (mode > 0) ? v1 + Pv * mo.vslot(mode - 1) : V1ext;
    else if (mode > -mo.nko) l1 = v1 + Pv * mo.vslot(mode - 1);
    if (l1) g1.add(l1, Dm.slot[mo.dslot(mode)], GM + Ps * s, 1.0, 0.0);
    else CHB_CUDA(zero(GM, s));
  }
  CHB_TRY(g1.flush());
  if (!d.env) {  // second term of component 1 accumulates on top of the first => separate launch
    Batcher g2(c.st, 2 * d.nkx, d.nkr_loc, d.nkr, 2 * d.nkx, 2 * d.nkx);
    for (i64 mode = mo.lo; mode <= mo.hi; ++mode) {
      const cd* l2 = (mode > 0) ? R2 + Pv * mo.vslot(mode - 1) : R2ext;
      if (l2) g2.add(l2, Dm.slot[mo.dslot(mode)], out + Ps * mo.vslot(mode), -1.0, 1.0);
    }
    CHB_TRY(g2.flush());
  }
  // out2 = -i kx v3 + i GP + i GM ; out3 = +i kx v2 - GP + GM in one pass
  return launch_rot_tail(c.st, out + Ps * d.nm, out + Ps * d.nm * 2, v2, v3, GP, GM, kx, d.nkx, Ps, Pv, d.nm);
}

}  // namespace chb
