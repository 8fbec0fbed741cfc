// fbops.cuh -- composite Fourier-Bessel operations on device pointers: DHT+FFT transforms
// (fb_io.f90) and spectral vector calculus (fb_math*.f90) assembled from the DMMA GEMM, cuFFT
// and the elementwise kernels.  Shared by the C-ABI host layer and the resident engine.
#pragma once
#include <cufft.h>
#include <map>
#include <vector>
#include <utility>
#include "kernels.cuh"

namespace chb {

// bump allocator over device memory for per-call temporaries.  take() never fails for lack of
// reserved space: it adds a block when the current one is exhausted; reset() (call it between
// top-level operations, after the stream has been synchronised or the temporaries are dead)
// coalesces the blocks into one, so a steady-state loop performs no cudaMalloc at all.
struct Scratch {
  struct Block { char* p; size_t cap, used; };
  std::vector<Block> blocks;
  size_t high = 0;
  unsigned long long gen = 0;          // bumped whenever a block is allocated or freed (captured graphs hold its addresses)
  int reserve(size_t bytes);            // make sure one block of at least `bytes` exists
  void* take(size_t bytes);             // 256-byte aligned; nullptr only if cudaMalloc fails
  template <typename T> T* take_n(i64 n) { return (T*)take(sizeof(T) * (size_t)(n > 0 ? n : 1)); }
  void reset();
  void destroy();
};

struct FFTCache {
  std::map<std::pair<i64, i64>, cufftHandle> plans;
  int exec(cudaStream_t st, cd* data, i64 n, i64 batch, int dir);  // in place, dir = CUFFT_FORWARD / CUFFT_INVERSE
  int exec(cudaStream_t st, cd* in, cd* out, i64 n, i64 batch, int dir);  // out of place (`in` is not written)
  void destroy();
};

// operator stack packed for the GEMM: one packed B per mode slot
struct PackedOps {
  const double* slot[2 * kMaxModes + 4];
  i64 K = 0, N = 0;
  int nslots = 0;
};
// bytes needed to pack an operator stack (K, N, nslots)
size_t packed_ops_bytes(i64 K, i64 N, int nslots);
// pack a device-resident operator stack op(K, N, nslots) (Fortran order) into `dst`
int pack_ops(cudaStream_t st, PackedOps& out, double* dst, const double* op, i64 K, i64 N, int nslots);

struct FBCtx {
  cudaStream_t st;
  Scratch* scr;
  FFTCache* fft;
};

// forward DHT + x-FFT + phase (fb_vec_in / fb_scl_in, fb_io.f90:18-98); fact (optional, (nkx,nkr,nm))
// is fused into the phase pass (the driver's omp_mult_vec(J_fb, DepFact), solvers.py:419)
int fb_in_dev(FBCtx& c, cd* out_fb, const cd* in, double leftX, const double* kx, const PackedOps& In,
              const double* fact, i64 nkx, i64 nrn, i64 nm, i64 nkr, int ncomp);
// backward DHT + phase + inverse x-FFT (fb_vec_out / fb_scl_out / fb_eb_out, fb_io.f90:100-228)
// srcs[j] is the spectral source of output component block j (ncomp_each comps each)
int fb_out_dev(FBCtx& c, cd* out, const cd* const* srcs, int nsrc, int ncomp_each, double leftX, const double* kx,
               const PackedOps& Out, i64 nkx, i64 nrn, i64 nm, i64 nkr);
// absorbing-layer window (fb_filtr, fb_io.f90:230-308)
int fb_filtr_dev(FBCtx& c, cd* vec, double leftX, const double* kx, const double* filtr, int modefilt, i64 nkx,
                 i64 nkr, i64 nm, i64 nxfilt);

struct FBMathDims {
  i64 nkx, nkr, nm, nkr_loc;
  int env;
  int mirror_shift = 0;  // rows are a kx slab of mirror pairs: partner of row i is (nkx - i - shift) mod nkx
};
// forward transform of a kx slab: x-FFT of the full grid first, then the DHT on the rows `rows` only
// (the two are linear maps on different axes and commute); out_fb has nxs rows
int fb_in_slab_dev(FBCtx& c, cd* out_fb, const cd* in, double leftX, const double* kx_slab, const PackedOps& In,
                   const double* fact_slab, const i64* rows, i64 nkx, i64 nxs, i64 nrn, i64 nm, i64 nkr, int ncomp);
// backward DHT + phase of a kx slab into out_slab (nxs, nrn, nm, ncomp); the inverse x-FFT follows the all-gather
int fb_out_slab_dev(FBCtx& c, cd* out_slab, const cd* const* srcs, int nsrc, int ncomp_each, double leftX,
                    const double* kx_slab, const PackedOps& Out, i64 nxs, i64 nrn, i64 nm, i64 nkr);
// rows of the gathered slabs back to natural kx order, inverse x-FFT
int fb_out_finish_dev(FBCtx& c, cd* out, const cd* gathered, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols);
// column-block dataflow of the multi-rank solve (engine.cu "colflow"): forward, x-FFT of a column block + rows sorted by
// destination rank; the DHT half of fb_in_slab_dev on a slab that arrived x-transformed.  Backward: fb_out_finish_dev
// on the received (rank, column, row) blocks is exactly the unpack + inverse x-FFT of a column block.
int col_fwd_dev(FBCtx& c, cd* send, cd* blk, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols);
int col_bwd_dev(FBCtx& c, cd* blk, const cd* recv, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols, i64 col0, i64 nrn,
                i64 nm, int env);  // col0: first column of the block; eb_correction's normalisation folded in
int eb_ghost_dev(FBCtx& c, cd* eb, i64 nxn, i64 nrn, i64 nm, int env, int ncomp);  // ghost rows of eb_correction
int fb_in_slab_post_dev(FBCtx& c, cd* out_fb, const cd* slab, double leftX, const double* kx_slab, const PackedOps& In,
                        const double* fact_slab, i64 nxs, i64 nrn, i64 nm, i64 nkr, int ncomp);
// row exchange between the all-gathered slab layout [rank][(nxs, ncols)], the full (nkx, ncols) array and one slab
int rows_scatter_dev(FBCtx& c, cd* full, const cd* gathered, const i64* gather_map, i64 nkx, i64 nxs, i64 ncols);
int rows_take_dev(FBCtx& c, cd* slab, const cd* full, const i64* rows, i64 nkx, i64 nxs, i64 ncols);
int fb_grad_dev(FBCtx& c, cd* out, const cd* scl, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
                const FBMathDims& d);
int fb_div_dev(FBCtx& c, cd* out, const cd* vec, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
               const FBMathDims& d);
int fb_rot_dev(FBCtx& c, cd* out, const cd* vec, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
               const FBMathDims& d);
int fb_poiss_iter_dev(FBCtx& c, cd* J, const cd* gn, const cd* gp, double dt_inv, const double* w2inv,
                      const PackedOps& Dp, const PackedOps& Dm, const double* kx, const FBMathDims& d);
int fb_graddiv_dev(FBCtx& c, cd* vec, const cd* in, const PackedOps& Dp, const PackedOps& Dm, const double* kx,
                   const FBMathDims& d);
// scratch bytes an fb_* call may take (upper bound), for reserve()
size_t fb_math_scratch_bytes(const FBMathDims& d);

}  // namespace chb
