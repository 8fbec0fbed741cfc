// api_host.cu -- the C ABI of include/chimera_b200.h with HOST buffers: each call stages its
// arguments to the device, runs the CUDA kernels and copies the results back (the drop-in
// boundary of SURVEY.md section 8b: what the reference's f2py layer binds).  Device-resident
// operation without the per-call copies is the engine's job (engine.cu).
#include <cub/device/device_scan.cuh>
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "../../include/chimera_b200.h"
#include "fbops.cuh"
#include "staging.cuh"

namespace chb {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

long long g_h2d_bytes = 0, g_d2h_bytes = 0;  // host<->device traffic of the host-buffer API
long long g_launches = 0;  // kernels launched by this library (bumped by CHB_LAUNCH_CHECK)

// process-wide context of the host-buffer API
struct HostCtx {
  bool ready = false;
  cudaStream_t st = nullptr;
  Scratch scr;
  FFTCache fft;
  Stager stage;  // pageable numpy buffers <-> device at PCIe speed (staging.cuh)
  std::mutex mu;
  int init() {
    if (ready) return 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      set_error("no CUDA device available (%s): libchimera_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
      return 1;
    }
    CHB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    ready = true;
    return 0;
  }
};
static HostCtx g_ctx;

// resident mode (below): does the managed range [p, p + bytes) need a prefetch to the device before a kernel uses it?
bool managed_needs_prefetch(const void* p, size_t bytes);

// RAII scope of one host call: staging helpers + final sync + scratch reset
struct Call {
  HostCtx& c;
  std::unique_lock<std::mutex> lock;
  int rc;
  explicit Call() : c(g_ctx), lock(g_ctx.mu), rc(g_ctx.init()) {}
  ~Call() {
    if (c.ready) {
      cudaStreamSynchronize(c.st);
      c.scr.reset();
    }
  }
  template <typename T> T* dev(i64 n) { return c.scr.take_n<T>(n); }
  // Resident mode (chimera_b200/resident.py): the caller's numpy arrays live in CUDA managed memory (numpy data
  // allocator, NEP 49), so the "host" pointer is a device pointer too -- no staging copy, only a prefetch that is a
  // no-op once the pages are on the device.  Plain device pointers pass through as well.
  static int accessible(const void* p) {  // 0: host memory, 1: managed, 2: device
    if (!p) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return at.type == cudaMemoryTypeManaged ? 1 : (at.type == cudaMemoryTypeDevice ? 2 : 0);
  }
  template <typename T> T* up(const T* h, i64 n) {
    if (n > 0) {
      const int kind = accessible(h);
      if (kind == 1 && managed_needs_prefetch(h, sizeof(T) * (size_t)n)) {
        int dev_id = 0;
        cudaGetDevice(&dev_id);
        if (cudaMemPrefetchAsync(h, sizeof(T) * (size_t)n, dev_id, c.st) != cudaSuccess) cudaGetLastError();
      }
      if (kind) return const_cast<T*>(h);
    }
    T* d = dev<T>(n);
    if (n > 0) g_h2d_bytes += (long long)(sizeof(T) * (size_t)n);
    if (d && n > 0 && c.stage.h2d(d, h, sizeof(T) * (size_t)n, c.st) != 0) return nullptr;
    return d;
  }
  template <typename T> int down(T* h, const T* d, i64 n) {
    if ((const void*)h == (const void*)d) return 0;  // computed in place in the caller's managed / device array
    if (n > 0 && accessible(h)) {
      CHB_CUDA(cudaMemcpyAsync(h, d, sizeof(T) * (size_t)n, cudaMemcpyDeviceToDevice, c.st));
      return 0;
    }
    if (n > 0) CHB_TRY(c.stage.d2h(h, d, sizeof(T) * (size_t)n, c.st));
    if (n > 0) g_d2h_bytes += (long long)(sizeof(T) * (size_t)n);
    return 0;
  }
  int sync() {
    CHB_CUDA(cudaStreamSynchronize(c.st));
    return 0;
  }
  FBCtx fb() { return FBCtx{c.st, &c.scr, &c.fft}; }
};

#define CALL_BEGIN() \
  Call call;         \
  if (call.rc) return call.rc
#define NEED(p) \
  if (!(p)) return 6

static GridGeom make_geom(const double* Rgrid_h, const double* Rgrid_d, double leftX, double dx_inv, double dr_inv,
                          double kx0, i64 nxn, i64 nrn, i64 nm) {
  GridGeom g;
  g.leftX = leftX; g.dx_inv = dx_inv; g.dr_inv = dr_inv; g.kx0 = kx0;
  g.r0 = Rgrid_h[0]; g.rmax = Rgrid_h[nrn - 1];
  g.Rgrid = Rgrid_d; g.nxn = nxn; g.nrn = nrn; g.nm = nm;
  return g;
}

static int deposit_host(int env, int curr, const double* coord, const double* momenta, const double* wghts,
                        double* grid, const int* ind, int guards, double leftX, const double* Rgrid, double dx_inv,
                        double dr_inv, double kx0, i64 np, i64 nxn, i64 nrn, i64 nm, i64 nchnk) {
  CALL_BEGIN();
  if (nxn < 2 || nrn < 2 || nm < 1) { set_error("deposit: bad grid shape"); return 2; }
  const i64 ng = nxn * nrn * nm * (curr ? 3 : 1);
  double* d_x = call.up(coord, 3 * np); NEED(d_x);
  double* d_p = nullptr;
  if (curr) { d_p = call.up(momenta, 3 * np); NEED(d_p); }
  double* d_w = call.up(wghts, np); NEED(d_w);
  cd* d_g = (cd*)call.up(grid, 2 * ng); NEED(d_g);
  double* d_r = call.up(Rgrid, nrn); NEED(d_r);
  ChunkSpec ch{0, nullptr, 1, 0, nxn};
  if (ind) {
    if (nchnk < 1) { set_error("deposit: nchnk < 1"); return 2; }
    int* d_ind = call.up(ind, nchnk + 1); NEED(d_ind);
    ch = ChunkSpec{1, d_ind, (int)nchnk, guards, nxn / nchnk};
  }
  GridGeom g = make_geom(Rgrid, d_r, leftX, dx_inv, dr_inv, kx0, nxn, nrn, nm);
  // Chunk-sorted particles of the driver keep the cell order they were generated in (genparts + a stable argsort), so
  // the CTA-binned deposit of the resident engine applies: particle planes in scratch, CTA table from the chunk table.
  // Whatever does not sit in a CTA's cell window takes the per-particle path inside that kernel, so any order is
  // correct; small calls and mode counts without an instantiation use the direct kernel.
  if (ind && np >= (i64)1 << 16 && Call::accessible(ind) != 2) {
    double* xs = call.dev<double>(3 * np); NEED(xs);
    CHB_TRY(launch_planes_from_aos(call.c.st, xs, d_x, 3, np, np));
    double* ps = nullptr;
    if (curr) {
      ps = call.dev<double>(3 * np); NEED(ps);
      CHB_TRY(launch_planes_from_aos(call.c.st, ps, d_p, 3, np, np));
    }
    std::vector<int> cta((size_t)nchnk + 1, 0);
    for (i64 c = 0; c < nchnk; ++c) {
      if (ind[c] < 0 || ind[c + 1] < ind[c] || ind[c + 1] > np) {  // the reference would read out of bounds
        set_error("deposit: IndInChunk is not a prefix table within 0..np (entry %lld)", (long long)c);
        return 2;
      }
      const int n = ind[c + 1] - ind[c];
      cta[c + 1] = cta[c] + (n > 0 ? (n + kDepNPB - 1) / kDepNPB : 0);
    }
    int* d_cta = call.dev<int>(nchnk + 1); NEED(d_cta);
    CHB_CUDA(cudaMemcpyAsync(d_cta, cta.data(), sizeof(int) * (size_t)(nchnk + 1), cudaMemcpyHostToDevice, call.c.st));
    CHB_CUDA(cudaStreamSynchronize(call.c.st));  // `cta` leaves scope
    SortedSpec sp{ch.ind, d_cta, (int)nchnk, cta[nchnk], nxn / nchnk, 0};
    const int rc = launch_deposit_binned(call.c.st, env, curr, xs, ps, d_w, np, d_g, g, ch, sp);
    if (rc != -1) {
      CHB_TRY(rc);
      CHB_TRY(launch_ghost_fold(call.c.st, d_g, nxn, nrn, nm * (curr ? 3 : 1)));
      CHB_TRY(call.down(grid, (double*)d_g, 2 * ng));
      return call.sync();
    }
  }
  CHB_TRY(launch_deposit_direct(call.c.st, env, curr, aos((const double*)d_x, 3), aos((const double*)d_p, 3), d_w, d_g,
                                g, ch, np, true));
  CHB_TRY(call.down(grid, (double*)d_g, 2 * ng));
  return call.sync();
}

static int gather_host(int env, const double* coord, const double* wghts, const double* Fld, double* Fld_tot,
                       double leftX, const double* Rgrid, double dx_inv, double dr_inv, double kx0, i64 np, i64 nxn,
                       i64 nrn, i64 nm) {
  CALL_BEGIN();
  if (nxn < 2 || nrn < 2 || nm < 1 || nm > 2 * kMaxModes) { set_error("proj_fld: bad grid shape"); return 2; }
  double* d_x = call.up(coord, 3 * np); NEED(d_x);
  double* d_w = call.up(wghts, np); NEED(d_w);
  cd* d_f = (cd*)call.up(Fld, 2 * nxn * nrn * nm * 6); NEED(d_f);
  double* d_o = call.up(Fld_tot, 6 * np); NEED(d_o);
  double* d_r = call.up(Rgrid, nrn); NEED(d_r);
  GridGeom g = make_geom(Rgrid, d_r, leftX, dx_inv, dr_inv, kx0, nxn, nrn, nm);
  // Large calls: the CTA-binned gather on particle planes in scratch (one "chunk" holding everything).  The driver's
  // particles keep the cell order they were generated / chunk-sorted in, so consecutive particles share cells; whatever
  // falls outside a CTA's cell window takes the per-particle path inside the kernel, so any order is correct.
  if (np >= (i64)1 << 16) {
    double* xs = call.dev<double>(3 * np); NEED(xs);
    CHB_TRY(launch_planes_from_aos(call.c.st, xs, d_x, 3, np, np));
    const int ncta = (int)((np + kDepNPB - 1) / kDepNPB);
    const int tab[4] = {0, (int)np, 0, ncta};  // IndInChunk(0:1), CTA prefix(0:1)
    int* d_tab = call.dev<int>(4); NEED(d_tab);
    CHB_CUDA(cudaMemcpyAsync(d_tab, tab, sizeof(tab), cudaMemcpyHostToDevice, call.c.st));
    CHB_CUDA(cudaStreamSynchronize(call.c.st));  // `tab` leaves scope
    SortedSpec sp{d_tab, d_tab + 2, 1, ncta, nxn, 0};
    const int rc = launch_gather_binned_out(call.c.st, env, xs, d_w, d_f, d_o, np, g, sp);
    if (rc != -1) {
      CHB_TRY(rc);
      CHB_TRY(call.down(Fld_tot, d_o, 6 * np));
      return call.sync();
    }
  }
  CHB_TRY(launch_gather(call.c.st, env, aos((const double*)d_x, 3), d_w, d_f, aos(d_o, 6), g, np));
  CHB_TRY(call.down(Fld_tot, d_o, 6 * np));
  return call.sync();
}

static int fb_in_host(double* out_fb, const double* in, double leftX, const double* kx, const double* In, i64 nkx,
                      i64 nrn, i64 nm, i64 nkr, int ncomp) {
  CALL_BEGIN();
  const i64 nr = nrn - 1;
  cd* d_in = (cd*)call.up(in, 2 * nkx * nrn * nm * ncomp); NEED(d_in);
  cd* d_out = call.dev<cd>(nkx * nkr * nm * ncomp); NEED(d_out);
  double* d_kx = call.up(kx, nkx); NEED(d_kx);
  double* d_op = call.up(In, nr * nkr * nm); NEED(d_op);
  double* d_pk = call.dev<double>(gemm_packed_size(nr, nkr) * nm); NEED(d_pk);
  PackedOps P;
  CHB_TRY(pack_ops(call.c.st, P, d_pk, d_op, nr, nkr, (int)nm));
  FBCtx fb = call.fb();
  CHB_TRY(fb_in_dev(fb, d_out, d_in, leftX, d_kx, P, nullptr, nkx, nrn, nm, nkr, ncomp));
  CHB_TRY(call.down(out_fb, (double*)d_out, 2 * nkx * nkr * nm * ncomp));
  return call.sync();
}

static int fb_out_host(double* out, const double* src0, const double* src1, int ncomp_each, double leftX,
                       const double* kx, const double* Out, i64 nkx, i64 nrn, i64 nm, i64 nkr, i64 src0_comps) {
  CALL_BEGIN();
  const i64 nr = nrn - 1;
  const int nsrc = src1 ? 2 : 1;
  const cd* srcs[2];
  cd* d_s0 = (cd*)call.up(src0, 2 * nkx * nkr * nm * src0_comps); NEED(d_s0);
  srcs[0] = d_s0;
  if (src1) { cd* d_s1 = (cd*)call.up(src1, 2 * nkx * nkr * nm * ncomp_each); NEED(d_s1); srcs[1] = d_s1; }
  cd* d_out = call.dev<cd>(nkx * nrn * nm * ncomp_each * nsrc); NEED(d_out);
  double* d_kx = call.up(kx, nkx); NEED(d_kx);
  double* d_op = call.up(Out, nkr * nr * nm); NEED(d_op);
  double* d_pk = call.dev<double>(gemm_packed_size(nkr, nr) * nm); NEED(d_pk);
  PackedOps P;
  CHB_TRY(pack_ops(call.c.st, P, d_pk, d_op, nkr, nr, (int)nm));
  FBCtx fb = call.fb();
  CHB_TRY(fb_out_dev(fb, d_out, srcs, nsrc, ncomp_each, leftX, d_kx, P, nkx, nrn, nm, nkr));
  CHB_TRY(call.down(out, (double*)d_out, 2 * nkx * nrn * nm * ncomp_each * nsrc));
  return call.sync();
}

enum FBMathOp { OP_ROT, OP_GRAD, OP_DIV, OP_GRADDIV };
static int fb_math_host(FBMathOp op, int env, double* out, const double* in, const double* Dp, const double* Dm,
                        const double* kx, i64 nkx, i64 nkr, i64 nm, i64 nkr_loc) {
  CALL_BEGIN();
  if (env && (nm % 2) != 1) { set_error("envelope operators need an odd number of mode slots"); return 2; }
  const int nd = (int)(env ? nm + 2 : nm + 1);
  FBMathDims d{nkx, nkr, nm, nkr_loc, env};
  const int in_comp = (op == OP_GRAD) ? 1 : 3;
  const int out_comp = (op == OP_DIV) ? 1 : 3;
  cd* d_in = (cd*)call.up(op == OP_GRADDIV ? out : in, 2 * nkx * nkr * nm * in_comp); NEED(d_in);
  cd* d_out = d_in;
  if (op != OP_GRADDIV) { d_out = call.dev<cd>(nkx * nkr_loc * nm * out_comp); NEED(d_out); }
  double* d_kx = call.up(kx, nkx); NEED(d_kx);
  double* d_dp = call.up(Dp, nkr * nkr_loc * nd); NEED(d_dp);
  double* d_dm = call.up(Dm, nkr * nkr_loc * nd); NEED(d_dm);
  double* pk_p = call.dev<double>(gemm_packed_size(nkr, nkr_loc) * nd); NEED(pk_p);
  double* pk_m = call.dev<double>(gemm_packed_size(nkr, nkr_loc) * nd); NEED(pk_m);
  PackedOps PP, PM;
  CHB_TRY(pack_ops(call.c.st, PP, pk_p, d_dp, nkr, nkr_loc, nd));
  CHB_TRY(pack_ops(call.c.st, PM, pk_m, d_dm, nkr, nkr_loc, nd));
  FBCtx fb = call.fb();
  switch (op) {
    case OP_ROT: CHB_TRY(fb_rot_dev(fb, d_out, d_in, PP, PM, d_kx, d)); break;
    case OP_GRAD: CHB_TRY(fb_grad_dev(fb, d_out, d_in, PP, PM, d_kx, d)); break;
    case OP_DIV: CHB_TRY(fb_div_dev(fb, d_out, d_in, PP, PM, d_kx, d)); break;
    case OP_GRADDIV: CHB_TRY(fb_graddiv_dev(fb, d_in, d_in, PP, PM, d_kx, d)); break;
  }
  CHB_TRY(call.down(out, (double*)d_out, 2 * nkx * nkr_loc * nm * out_comp));
  return call.sync();
}

// ---- stream compaction helper: out[] = indices with flag != 0, returns count in *num
static int compact_host(Call& call, const int* d_flag, int* h_idx, int* h_num, i64 np) {
  int* d_pos = call.dev<int>(np + 1); NEED(d_pos);
  int* d_idx = call.dev<int>(np + 1); NEED(d_idx);
  CHB_CUDA(cudaMemsetAsync(d_idx, 0, sizeof(int) * (size_t)(np + 1), call.c.st));
  size_t tmp_bytes = 0;
  CHB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag, d_pos, (int)np, call.c.st));
  void* d_tmp = call.c.scr.take(tmp_bytes + 16); NEED(d_tmp);
  CHB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flag, d_pos, (int)np, call.c.st));
  CHB_TRY(launch_compact_index(call.c.st, d_flag, d_pos, d_idx, np));
  int last_pos = 0, last_flag = 0;
  if (np > 0) {
    CHB_CUDA(cudaMemcpyAsync(&last_pos, d_pos + np - 1, sizeof(int), cudaMemcpyDeviceToHost, call.c.st));
    CHB_CUDA(cudaMemcpyAsync(&last_flag, d_flag + np - 1, sizeof(int), cudaMemcpyDeviceToHost, call.c.st));
  }
  CHB_TRY(call.down(h_idx, d_idx, np));
  CHB_TRY(call.sync());
  *h_num = last_pos + last_flag;
  return 0;
}

// genparts (particle_tools.f90:84-128): candidate (ir, ix, ip) in that nesting order
__global__ void __launch_bounds__(256) genparts_k(double* __restrict__ cand, int* __restrict__ flag,
                                                  const double* __restrict__ Xgrid, const double* __restrict__ Rgrid,
                                                  const double* __restrict__ RandPackO, const double* __restrict__ PackX,
                                                  const double* __restrict__ PackR, const cd* __restrict__ PackO,
                                                  i64 nx, i64 nr, i64 ppc, i64 ncand) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ncand) return;
  const i64 ip = e % ppc, cell = e / ppc;
  const i64 ix = cell % (nx - 1), ir = cell / (nx - 1);
  const double dr_2 = 0.5 * (Rgrid[1] - Rgrid[0]);
  const double r0 = Rgrid[ir] + dr_2, r1 = Rgrid[ir + 1] + dr_2, x0 = Xgrid[ix], x1 = Xgrid[ix + 1];
  const double xc = x0 + (x1 - x0) * PackX[ip], rc = r0 + (r1 - r0) * PackR[ip];
  double sn, cs;
  sincos(2.0 * 3.14159265358979323846 * RandPackO[ix + nx * ir], &sn, &cs);
  const cd oc = cmul(PackO[ip], cmake(cs, sn));
  cand[4 * e + 0] = xc;
  cand[4 * e + 1] = rc * oc.y;
  cand[4 * e + 2] = rc * oc.x;
  cand[4 * e + 3] = rc;
  flag[e] = (rc <= 0) ? 0 : 1;
}
__global__ void __launch_bounds__(256) genparts_scatter_k(double* __restrict__ coord, const double* __restrict__ cand,
                                                          const int* __restrict__ flag, const int* __restrict__ pos,
                                                          i64 ncand, i64 np) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ncand || !flag[e]) return;
  const i64 j = pos[e];
  if (j >= np) return;
  for (int c = 0; c < 4; ++c) coord[4 * j + c] = cand[4 * e + c];
}

// pseudo-random fill in [-1,1) for the microbenchmark (zeros would under-state the power draw)
__global__ void __launch_bounds__(256) fill_pattern_k(double* __restrict__ a, i64 n, unsigned long long seed) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  unsigned long long z = seed + (unsigned long long)e * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  a[e] = (double)(long long)(z >> 11) * (1.0 / 4503599627370496.0) - 1.0;
}


// ---- SR.f90 / utils.f90 (NEXT-4): host-buffer staging around the kernels of sr.cu
static int sr_far_host(double* spect, const double* coords, const double* mprv, const double* mnxt, const double* wghts,
                       int comp, double dt, const double* omega, const double* SinTh, const double* CosTh,
                       const double* SinPh, const double* CosPh, i64 nt, i64 np, i64 nom, i64 nth, i64 nph) {
  CALL_BEGIN();
  if (nt < 0 || np < 0 || nom < 0 || nth < 0 || nph < 0) { set_error("sr_calc_far: negative extent"); return 2; }
  if (!(dt != 0.0)) { set_error("sr_calc_far: dt must be non-zero"); return 2; }
  const i64 ns = nom * nth * nph, ntr = 3 * nt * np;
  if (ns == 0 || ntr == 0) return 0;
  double* d_s = call.up(spect, ns); NEED(d_s);
  double* d_x = call.up(coords, ntr); NEED(d_x);
  double* d_a = call.up(mprv, ntr); NEED(d_a);
  double* d_b = call.up(mnxt, ntr); NEED(d_b);
  double* d_w = call.up(wghts, np); NEED(d_w);
  double* d_o = call.up(omega, nom); NEED(d_o);
  double* d_st = call.up(SinTh, nth); NEED(d_st);
  double* d_ct = call.up(CosTh, nth); NEED(d_ct);
  double* d_sp = call.up(SinPh, nph); NEED(d_sp);
  double* d_cp = call.up(CosPh, nph); NEED(d_cp);
  CHB_TRY(launch_sr_far(call.c.st, d_s, d_x, d_a, d_b, d_w, comp, dt, d_o, d_st, d_ct, d_sp, d_cp, nt, np, nom, nth, nph));
  CHB_TRY(call.down(spect, d_s, ns));
  return call.sync();
}

static int sr_near_host(double* spect, const double* coords, const double* mom, const double* wghts, int comp,
                        double dt, const double* omega, const double* G1, const double* G2s, const double* G2c,
                        int circ, double z_scr, i64 nt, i64 np, i64 nom, i64 n1, i64 n2) {
  CALL_BEGIN();
  if (nt < 0 || np < 0 || nom < 0 || n1 < 0 || n2 < 0) { set_error("sr_calc_near: negative extent"); return 2; }
  if (comp < 0 || comp > 3) { set_error("sr_calc_near: comp must be 1, 2 or 3 (got %d)", comp); return 2; }
  const i64 ns = nom * n1 * n2, ntr = 3 * nt * np;
  if (ns == 0 || ntr == 0) return 0;
  double* d_s = call.up(spect, ns); NEED(d_s);
  double* d_x = call.up(coords, ntr); NEED(d_x);
  double* d_m = call.up(mom, ntr); NEED(d_m);
  double* d_w = call.up(wghts, np); NEED(d_w);
  double* d_o = call.up(omega, nom); NEED(d_o);
  double* d_1 = call.up(G1, n1); NEED(d_1);
  double* d_2s = call.up(G2s, n2); NEED(d_2s);
  double* d_2c = nullptr;
  if (circ) { d_2c = call.up(G2c, n2); NEED(d_2c); }
  CHB_TRY(launch_sr_near(call.c.st, d_s, d_x, d_m, d_w, comp, dt, d_o, d_1, d_2s, d_2c, circ, z_scr, nt, np, nom, n1, n2));
  CHB_TRY(call.down(spect, d_s, ns));
  return call.sync();
}

}  // namespace chb

using namespace chb;

extern "C" {

const char* chimera_last_error(void) { return get_error(); }
const char* chimera_version(void) { return "chimera_b200 0.1 (sm_100a)"; }
int chimera_device_count(int* n) {
  *n = 0;
  if (cudaGetDeviceCount(n) != cudaSuccess) { *n = 0; cudaGetLastError(); }
  return 0;
}
int chimera_set_device(int device) {
  CHB_CUDA(cudaSetDevice(device));
  return 0;
}
int chimera_sync(void) {
  CHB_CUDA(cudaDeviceSynchronize());
  return 0;
}
int chimera_kernel_launches(chb_i64* n) { *n = g_launches; return 0; }

// ---- resident mode: numpy's data allocator backed by CUDA managed memory, and the driver's whole-array statements
// (J[:] = 0, Rho += BckGrndRho, gradRho_prv[:] = gradRho_nxt, chimera_main.py:110-190, solvers.py:318) on the device
namespace {
std::mutex g_mm_mu;
struct MBlock { size_t bytes, cap; bool on_dev; unsigned uses; };  // what numpy asked for, what was allocated, see below
std::map<const void*, MBlock> g_mm;                           // managed blocks handed to numpy, by base address
std::unordered_map<size_t, std::vector<void*>> g_mm_free;     // freed blocks by capacity: numpy temporaries come in
size_t g_mm_cached = 0;                                       // repeating sizes, and cudaMallocManaged / cudaFree cost
constexpr size_t kMMCacheMax = size_t(8) << 30;               // milliseconds each (cudaFree also synchronises the device)
constexpr size_t kMMGran = size_t(2) << 20;
}  // namespace
void* chimera_managed_alloc(size_t bytes, int zero) {
  if (bytes == 0) bytes = 1;
  const size_t cap = (bytes + kMMGran - 1) / kMMGran * kMMGran;
  void* p = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mm_mu);
    auto it = g_mm_free.find(cap);
    if (it != g_mm_free.end() && !it->second.empty()) {
      p = it->second.back();
      it->second.pop_back();
      g_mm_cached -= cap;
    }
  }
  if (!p && cudaMallocManaged(&p, cap, cudaMemAttachGlobal) != cudaSuccess) {
    cudaGetLastError();
    chimera_managed_trim();  // give the cached blocks back and try once more
    if (cudaMallocManaged(&p, cap, cudaMemAttachGlobal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  }
  if (zero) {  // zero on the device: np.zeros arrays of the driver (J, Rho, EB, the spectral state) are used there first
    if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return nullptr; }
    cudaDeviceSynchronize();
  }
  std::lock_guard<std::mutex> lk(g_mm_mu);
  g_mm[p] = MBlock{bytes, cap, false, 0};  // np.zeros + a fill on the host is a common pattern: first use prefetches
  return p;
}

extern "C++" {
namespace chb {
// cudaMemPrefetchAsync on a range that already lives on the device is not free: ~50 us per call plus ~1 ms per GB of
// range walked -- 30 ms of the 70 ms per-call LWFA step, a quarter of the demo-size steps.  So a block is prefetched when
// it is first seen after its (re)allocation -- numpy arrays the driver built on the CPU: the re-ordered particle arrays of
// a re-binning step, tables made with np.zeros + a fill --, when the Python layer reports a host-side access
// (chimera_managed_touched), and at every 16th use (bounds what an unreported host access can cost: such a range is still
// correct, managed memory is coherent, but it comes in through page faults -- measured 12 MB in 50 ms); otherwise it is
// taken to be where the last call left it.  Blocks under 4 MB are prefetched at every call as before.
// CHIMERA_B200_PREFETCH=always | never overrides.
constexpr size_t kPrefetchAlwaysBelow = size_t(4) << 20;
bool managed_needs_prefetch(const void* p, size_t bytes) {
  static const int policy = [] {
    const char* e = getenv("CHIMERA_B200_PREFETCH");
    return !e ? 1 : (!strcmp(e, "always") ? 2 : (!strcmp(e, "never") ? 0 : 1));
  }();
  if (policy != 1) return policy == 2;
  std::lock_guard<std::mutex> lk(g_mm_mu);
  auto it = g_mm.upper_bound(p);
  if (it == g_mm.begin()) return true;  // not one of our blocks (someone else's managed memory)
  --it;
  const char* base = (const char*)it->first;
  if ((const char*)p + bytes > base + it->second.cap || it->second.cap < kPrefetchAlwaysBelow) return true;
  if (it->second.on_dev) return (++it->second.uses & 15u) == 0;
  // a partial range (a slice of the array) is prefetched without changing the block's state
  if ((const char*)p == base && bytes >= it->second.bytes) it->second.on_dev = true;
  return true;
}
}  // namespace chb
}  // extern "C++"

// host-side write to a managed array reported by the Python layer (chimera_b200/resident.py): prefetch it next time
void chimera_managed_touched(const void* p) {
  std::lock_guard<std::mutex> lk(g_mm_mu);
  auto it = g_mm.upper_bound(p);
  if (it == g_mm.begin()) return;
  --it;
  if ((const char*)p < (const char*)it->first + it->second.cap) it->second.on_dev = false;
}
int chimera_managed_owns(const void* p, size_t* bytes) {
  std::lock_guard<std::mutex> lk(g_mm_mu);
  auto it = g_mm.find(p);
  if (it == g_mm.end()) return 0;
  if (bytes) *bytes = it->second.bytes;
  return 1;
}
void chimera_managed_free(void* p) {
  size_t cap = 0;
  {
    std::lock_guard<std::mutex> lk(g_mm_mu);
    auto it = g_mm.find(p);
    if (it != g_mm.end()) { cap = it->second.cap; g_mm.erase(it); }
    if (cap && g_mm_cached + cap <= kMMCacheMax) {
      g_mm_free[cap].push_back(p);
      g_mm_cached += cap;
      return;
    }
  }
  if (cudaFree(p) != cudaSuccess) cudaGetLastError();  // at interpreter exit the context may be gone already
}
void chimera_managed_trim(void) {
  std::unordered_map<size_t, std::vector<void*>> drop;
  {
    std::lock_guard<std::mutex> lk(g_mm_mu);
    drop.swap(g_mm_free);
    g_mm_cached = 0;
  }
  for (auto& kv : drop)
    for (void* q : kv.second)
      if (cudaFree(q) != cudaSuccess) cudaGetLastError();
}
void* chimera_managed_realloc(void* p, size_t new_bytes) {
  size_t old = 0;
  {
    std::lock_guard<std::mutex> lk(g_mm_mu);
    auto it = g_mm.find(p);
    if (it == g_mm.end()) return nullptr;
    old = it->second.bytes;
    if (new_bytes <= it->second.cap) {  // fits the block: nothing moves
      it->second.bytes = new_bytes ? new_bytes : 1;
      it->second.on_dev = false;  // the driver fills the new tail on the host
      it->second.uses = 0;
      return p;
    }
  }
  void* q = chimera_managed_alloc(new_bytes, 0);
  if (!q) return nullptr;
  const size_t n = old < new_bytes ? old : new_bytes;
  if (n && cudaMemcpy(q, p, n, cudaMemcpyDefault) != cudaSuccess) { cudaGetLastError(); chimera_managed_free(q); return nullptr; }
  chimera_managed_free(p);
  return q;
}
int chimera_is_device_accessible(const void* p) { return Call::accessible(p); }
// y[0:n] = value (doubles; a complex array is 2n doubles with value_im in the odd slots)
int chimera_fill(double* y, chb_i64 n, double value_re, double value_im, int is_complex) {
  CALL_BEGIN();
  if (!Call::accessible(y)) { set_error("chimera_fill: not a managed / device array"); return 2; }
  if (value_re == 0.0 && value_im == 0.0) {
    CHB_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * (size_t)n * (is_complex ? 2 : 1), call.c.st));
    return call.sync();
  }
  CHB_TRY(launch_fill(call.c.st, y, n, value_re, value_im, is_complex));
  return call.sync();
}
int chimera_copy(void* dst, const void* src, chb_i64 bytes) {
  CALL_BEGIN();
  if (!Call::accessible(dst) || !Call::accessible(src)) { set_error("chimera_copy: not managed / device arrays"); return 2; }
  CHB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, call.c.st));
  return call.sync();
}
int chimera_add_inplace(double* y, const double* x, chb_i64 n) {  // y += x over n doubles
  CALL_BEGIN();
  if (!Call::accessible(y) || !Call::accessible(x)) { set_error("chimera_add_inplace: not managed / device arrays"); return 2; }
  CHB_TRY(launch_add_f64(call.c.st, y, x, n));
  return call.sync();
}
int chimera_host_traffic(chb_i64* h2d, chb_i64* d2h, int reset) {
  if (h2d) *h2d = g_h2d_bytes;
  if (d2h) *d2h = g_d2h_bytes;
  if (reset) g_h2d_bytes = g_d2h_bytes = 0;
  return 0;
}

// ------------------------------------------------------------------ particle_tools.f90
int chimera_push_velocs(double* momenta, const double* Fld, double dt, chb_i64 np) {
  CALL_BEGIN();
  double* d_p = call.up(momenta, 3 * np); NEED(d_p);
  double* d_f = call.up(Fld, 6 * np); NEED(d_f);
  CHB_TRY(launch_push_velocs(call.c.st, aos(d_p, 3), aos((const double*)d_f, 6), dt, np));
  CHB_TRY(call.down(momenta, d_p, 3 * np));
  return call.sync();
}

int chimera_push_coords(double* coord, const double* momenta, double* coord_cntr, double dt, chb_i64 np) {
  CALL_BEGIN();
  double* d_x = call.up(coord, 3 * np); NEED(d_x);
  double* d_p = call.up(momenta, 3 * np); NEED(d_p);
  double* d_c = call.dev<double>(3 * np); NEED(d_c);
  CHB_TRY(launch_push_coords(call.c.st, aos(d_x, 3), aos((const double*)d_p, 3), aos(d_c, 3), dt, np));
  CHB_TRY(call.down(coord, d_x, 3 * np));
  CHB_TRY(call.down(coord_cntr, d_c, 3 * np));
  return call.sync();
}

int chimera_genparts(double* coord, int* indPart, const double* Xgrid, const double* Rgrid, const double* RandPackO,
                     const double* PackX, const double* PackR, const double* PackO, chb_i64 np, chb_i64 nx,
                     chb_i64 nr, chb_i64 ppc) {
  CALL_BEGIN();
  *indPart = 0;
  const i64 ncand = (nx - 1) * (nr - 1) * ppc;
  double* d_coord = call.dev<double>(4 * np); NEED(d_coord);
  CHB_CUDA(cudaMemsetAsync(d_coord, 0, sizeof(double) * 4 * (size_t)np, call.c.st));
  if (ncand > 0) {
    double* d_xg = call.up(Xgrid, nx); NEED(d_xg);
    double* d_rg = call.up(Rgrid, nr); NEED(d_rg);
    double* d_ro = call.up(RandPackO, nx * nr); NEED(d_ro);
    double* d_px = call.up(PackX, ppc); NEED(d_px);
    double* d_pr = call.up(PackR, ppc); NEED(d_pr);
    cd* d_po = (cd*)call.up(PackO, 2 * ppc); NEED(d_po);
    double* d_cand = call.dev<double>(4 * ncand); NEED(d_cand);
    int* d_flag = call.dev<int>(ncand); NEED(d_flag);
    int* d_pos = call.dev<int>(ncand); NEED(d_pos);
    genparts_k<<<grid_for(ncand, 256), 256, 0, call.c.st>>>(d_cand, d_flag, d_xg, d_rg, d_ro, d_px, d_pr, d_po, nx, nr,
                                                           ppc, ncand);
    CHB_LAUNCH_CHECK();
    size_t tmp_bytes = 0;
    CHB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag, d_pos, (int)ncand, call.c.st));
    void* d_tmp = call.c.scr.take(tmp_bytes + 16); NEED(d_tmp);
    CHB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flag, d_pos, (int)ncand, call.c.st));
    genparts_scatter_k<<<grid_for(ncand, 256), 256, 0, call.c.st>>>(d_coord, d_cand, d_flag, d_pos, ncand, np);
    CHB_LAUNCH_CHECK();
    int lp = 0, lf = 0;
    CHB_CUDA(cudaMemcpyAsync(&lp, d_pos + ncand - 1, sizeof(int), cudaMemcpyDeviceToHost, call.c.st));
    CHB_CUDA(cudaMemcpyAsync(&lf, d_flag + ncand - 1, sizeof(int), cudaMemcpyDeviceToHost, call.c.st));
    CHB_TRY(call.sync());
    if (lp + lf > np) { set_error("genparts: coord holds %lld particles, %d generated", np, lp + lf); return 2; }
    *indPart = lp + lf;
  }
  CHB_TRY(call.down(coord, d_coord, 4 * np));
  return call.sync();
}

int chimera_sortpartsout(int* indx2stay, int* num2stay, const double* coord, const double* lims, chb_i64 np) {
  CALL_BEGIN();
  *num2stay = 0;
  if (np <= 0) return 0;
  double* d_x = call.up(coord, 3 * np); NEED(d_x);
  int* d_flag = call.dev<int>(np); NEED(d_flag);
  CHB_TRY(launch_inside_flag(call.c.st, aos((const double*)d_x, 3), d_flag, lims, np));
  return compact_host(call, d_flag, indx2stay, num2stay, np);
}

int chimera_sortoutghosts(int* indx2stay, int* num2stay, const double* coord, chb_i64 np) {
  CALL_BEGIN();
  *num2stay = 0;
  if (np <= 0) return 0;
  double* d_x = call.up(coord, np); NEED(d_x);
  int* d_flag = call.dev<int>(np); NEED(d_flag);
  CHB_TRY(launch_nonzero_flag(call.c.st, d_x, d_flag, np));
  return compact_host(call, d_flag, indx2stay, num2stay, np);
}

int chimera_chunk_coords_boundaries(int8_t* chunked_indx, int* IndInChnk, int* GoOut, const double* coord,
                                    const double* lims, const double* Xgrid, int nchnk, chb_i64 np, chb_i64 nxg) {
  CALL_BEGIN();
  if (nchnk < 1 || nchnk > 127) { set_error("nchnk must be in 1..127 (int8 chunk ids)"); return 2; }
  const i64 nx = nxg - 1;
  const double inv = (nchnk > 1) ? 1.0 / (Xgrid[(nx + 1) / nchnk] - Xgrid[0]) : 1.0 / (Xgrid[nx] - Xgrid[0]);
  double* d_x = call.up(coord, 3 * np); NEED(d_x);
  int8_t* d_id = call.dev<int8_t>(np); NEED(d_id);
  int* d_cnt = call.dev<int>(nchnk + 1); NEED(d_cnt);
  CHB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (nchnk + 1), call.c.st));
  CHB_TRY(launch_chunk_bin(call.c.st, aos((const double*)d_x, 3), d_id, d_cnt, d_cnt + nchnk, Xgrid[0], inv, lims,
                           nchnk, np));
  std::vector<int> cnt(nchnk + 1, 0);
  CHB_TRY(call.down(chunked_indx, d_id, np));
  CHB_TRY(call.down(cnt.data(), d_cnt, nchnk + 1));
  CHB_TRY(call.sync());
  IndInChnk[0] = 0;
  for (int c = 0; c < nchnk; ++c) IndInChnk[c + 1] = IndInChnk[c] + cnt[c];
  *GoOut = cnt[nchnk];
  return 0;
}

static int align_host(double* dat, const chb_i64* idx, i64 np, i64 np0, int ncomp) {
  CALL_BEGIN();
  if (np > np0) { set_error("align_data: np > np0"); return 2; }
  double* d_src = call.up(dat, ncomp * np0); NEED(d_src);
  i64* d_idx = call.up((const i64*)idx, np); NEED(d_idx);
  double* d_dst = call.dev<double>(ncomp * np); NEED(d_dst);
  CHB_TRY(launch_permute(call.c.st, aos(d_dst, ncomp), aos((const double*)d_src, ncomp), d_idx, ncomp, np));
  CHB_TRY(call.down(dat, d_dst, ncomp * np));
  return call.sync();
}
int chimera_align_data_vec(double* dat, const chb_i64* idx, chb_i64 np, chb_i64 np0) { return align_host(dat, idx, np, np0, 3); }
int chimera_align_data_scl(double* dat, const chb_i64* idx, chb_i64 np, chb_i64 np0) { return align_host(dat, idx, np, np0, 1); }

// ------------------------------------------------------------------ grid_deps*.f90
int chimera_dep_curr(const double* coord, const double* momenta, const double* wghts, double* curr, double leftX,
                     const double* Rgrid, double dx_inv, double dr_inv, chb_i64 np, chb_i64 nxn, chb_i64 nrn,
                     chb_i64 nm) {
  return deposit_host(0, 1, coord, momenta, wghts, curr, nullptr, 0, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, 1);
}
int chimera_dep_dens(const double* coord, const double* wghts, double* dens, double leftX, const double* Rgrid,
                     double dx_inv, double dr_inv, chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm) {
  return deposit_host(0, 0, coord, nullptr, wghts, dens, nullptr, 0, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, 1);
}
int chimera_dep_curr_chnk(const double* coord, const double* momenta, const double* wghts, double* curr,
                          const int* IndInChunk, int guards, double leftX, const double* Rgrid, double dx_inv,
                          double dr_inv, chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm, chb_i64 nchnk) {
  return deposit_host(0, 1, coord, momenta, wghts, curr, IndInChunk, guards, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, nchnk);
}
int chimera_dep_dens_chnk(const double* coord, const double* wghts, double* dens, const int* IndInChunk, int guards,
                          double leftX, const double* Rgrid, double dx_inv, double dr_inv, chb_i64 np, chb_i64 nxn,
                          chb_i64 nrn, chb_i64 nm, chb_i64 nchnk) {
  return deposit_host(0, 0, coord, nullptr, wghts, dens, IndInChunk, guards, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm, nchnk);
}
int chimera_dep_curr_env(const double* coord, const double* momenta, const double* wghts, double* curr, double leftX,
                         const double* Rgrid, double dx_inv, double dr_inv, double kx0, chb_i64 np, chb_i64 nxn,
                         chb_i64 nrn, chb_i64 nm) {
  return deposit_host(1, 1, coord, momenta, wghts, curr, nullptr, 0, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, 1);
}
int chimera_dep_dens_env(const double* coord, const double* wghts, double* dens, double leftX, const double* Rgrid,
                         double dx_inv, double dr_inv, double kx0, chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm) {
  return deposit_host(1, 0, coord, nullptr, wghts, dens, nullptr, 0, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, 1);
}
int chimera_dep_curr_env_chnk(const double* coord, const double* momenta, const double* wghts, double* curr,
                              const int* IndInChunk, int guards, double leftX, const double* Rgrid, double dx_inv,
                              double dr_inv, double kx0, chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm,
                              chb_i64 nchnk) {
  return deposit_host(1, 1, coord, momenta, wghts, curr, IndInChunk, guards, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, nchnk);
}
int chimera_dep_dens_env_chnk(const double* coord, const double* wghts, double* dens, const int* IndInChunk,
                              int guards, double leftX, const double* Rgrid, double dx_inv, double dr_inv, double kx0,
                              chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm, chb_i64 nchnk) {
  return deposit_host(1, 0, coord, nullptr, wghts, dens, IndInChunk, guards, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm, nchnk);
}
int chimera_proj_fld(const double* coord, const double* wghts, const double* Fld, double* Fld_tot, double leftX,
                     const double* Rgrid, double dx_inv, double dr_inv, chb_i64 np, chb_i64 nxn, chb_i64 nrn,
                     chb_i64 nm) {
  return gather_host(0, coord, wghts, Fld, Fld_tot, leftX, Rgrid, dx_inv, dr_inv, 0.0, np, nxn, nrn, nm);
}
int chimera_proj_fld_env(const double* coord, const double* wghts, const double* Fld, double* Fld_tot, double leftX,
                         const double* Rgrid, double dx_inv, double dr_inv, double kx0, chb_i64 np, chb_i64 nxn,
                         chb_i64 nrn, chb_i64 nm) {
  return gather_host(1, coord, wghts, Fld, Fld_tot, leftX, Rgrid, dx_inv, dr_inv, kx0, np, nxn, nrn, nm);
}
static int ebcorr_host(double* eb, i64 nxn, i64 nrn, i64 nm, int env) {
  CALL_BEGIN();
  const i64 n = nxn * nrn * nm * 6;
  cd* d = (cd*)call.up(eb, 2 * n); NEED(d);
  CHB_TRY(launch_eb_correction(call.c.st, d, nxn, nrn, nm, env));
  CHB_TRY(call.down(eb, (double*)d, 2 * n));
  return call.sync();
}
int chimera_eb_correction(double* eb, chb_i64 nxn, chb_i64 nrn, chb_i64 nm) { return ebcorr_host(eb, nxn, nrn, nm, 0); }
int chimera_eb_correction_env(double* eb, chb_i64 nxn, chb_i64 nrn, chb_i64 nm) { return ebcorr_host(eb, nxn, nrn, nm, 1); }

// ------------------------------------------------------------------ fb_io.f90
int chimera_fb_vec_in(double* vec_fb, const double* vec, double leftX, const double* kx, const double* In, chb_i64 nkx,
                      chb_i64 nrn, chb_i64 nm, chb_i64 nkr) {
  return fb_in_host(vec_fb, vec, leftX, kx, In, nkx, nrn, nm, nkr, 3);
}
int chimera_fb_scl_in(double* scl_fb, const double* scl, double leftX, const double* kx, const double* In, chb_i64 nkx,
                      chb_i64 nrn, chb_i64 nm, chb_i64 nkr) {
  return fb_in_host(scl_fb, scl, leftX, kx, In, nkx, nrn, nm, nkr, 1);
}
int chimera_fb_vec_out(double* vec, const double* vec_fb, double leftX, const double* kx, const double* Out,
                       chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr) {
  return fb_out_host(vec, vec_fb, nullptr, 3, leftX, kx, Out, nkx, nrn, nm, nkr, 3);
}
int chimera_fb_scl_out(double* scl, const double* scl_fb, double leftX, const double* kx, const double* Out,
                       chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr) {
  return fb_out_host(scl, scl_fb, nullptr, 1, leftX, kx, Out, nkx, nrn, nm, nkr, 1);
}
int chimera_fb_eb_out(double* eb_spc, const double* e_fb, const double* b_fb, double leftX, const double* kx,
                      const double* Out, chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr) {
  // only the first three components of e_fb (nkx,nkr,nm,6) are read (fb_io.f90:207)
  return fb_out_host(eb_spc, e_fb, b_fb, 3, leftX, kx, Out, nkx, nrn, nm, nkr, 3);
}
int chimera_fb_filtr(double* vec, double leftX, const double* kx, const double* filtr, int modefilt, chb_i64 nkx,
                     chb_i64 nkr, chb_i64 nm, chb_i64 nxfilt) {
  CALL_BEGIN();
  if (nxfilt > nkx) { set_error("fb_filtr: window longer than the grid"); return 2; }
  const i64 n = nkx * nkr * nm * 3;
  cd* d = (cd*)call.up(vec, 2 * n); NEED(d);
  double* d_kx = call.up(kx, nkx); NEED(d_kx);
  double* d_f = call.up(filtr, nxfilt); NEED(d_f);
  FBCtx fb = call.fb();
  CHB_TRY(fb_filtr_dev(fb, d, leftX, d_kx, d_f, modefilt, nkx, nkr, nm, nxfilt));
  CHB_TRY(call.down(vec, (double*)d, 2 * n));
  return call.sync();
}

// ------------------------------------------------------------------ fb_math*.f90
#define FBM(name, op, env, has_in)                                                                              \
  int chimera_##name(double* out, const double* in, const double* Dp, const double* Dm, const double* kx,      \
                     chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc) {                                   \
    return fb_math_host(op, env, out, in, Dp, Dm, kx, nkx, nkr, nm, nkr_loc);                                  \
  }
FBM(fb_rot, OP_ROT, 0, 1)
FBM(fb_grad, OP_GRAD, 0, 1)
FBM(fb_div, OP_DIV, 0, 1)
FBM(fb_rot_env, OP_ROT, 1, 1)
FBM(fb_grad_env, OP_GRAD, 1, 1)
FBM(fb_div_env, OP_DIV, 1, 1)
int chimera_fb_graddiv(double* vec_fb, const double* Dp, const double* Dm, const double* kx, chb_i64 nkx, chb_i64 nkr,
                       chb_i64 nm, chb_i64 nkr_loc) {
  return fb_math_host(OP_GRADDIV, 0, vec_fb, nullptr, Dp, Dm, kx, nkx, nkr, nm, nkr_loc);
}
int chimera_fb_graddiv_env(double* vec_fb, const double* Dp, const double* Dm, const double* kx, chb_i64 nkx,
                           chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc) {
  return fb_math_host(OP_GRADDIV, 1, vec_fb, nullptr, Dp, Dm, kx, nkx, nkr, nm, nkr_loc);
}

// ------------------------------------------------------------------ maxwell_solvers.f90
int chimera_maxwell_push_with_spchrg(double* EG, const double* j, const double* gn, const double* gp, const double* C1,
                                     const double* C2, chb_i64 nkx, chb_i64 nkr, chb_i64 nm) {
  CALL_BEGIN();
  const i64 P = nkx * nkr * nm;
  cd* d_eg = (cd*)call.up(EG, 2 * P * 6); NEED(d_eg);
  cd* d_j = (cd*)call.up(j, 2 * P * 3); NEED(d_j);
  cd* d_gn = (cd*)call.up(gn, 2 * P * 3); NEED(d_gn);
  cd* d_gp = (cd*)call.up(gp, 2 * P * 3); NEED(d_gp);
  double* d_c1 = call.up(C1, P * 5); NEED(d_c1);
  double* d_c2 = call.up(C2, P * 5); NEED(d_c2);
  CHB_TRY(launch_maxwell_push(call.c.st, d_eg, d_j, d_gn, d_gp, d_c1, d_c2, 5, 0, P));
  CHB_TRY(call.down(EG, (double*)d_eg, 2 * P * 6));
  return call.sync();
}
int chimera_maxwell_push_wo_spchrg(double* EG, const double* j, const double* C1, const double* C2, chb_i64 nkx,
                                   chb_i64 nkr, chb_i64 nm) {
  CALL_BEGIN();
  const i64 P = nkx * nkr * nm;
  cd* d_eg = (cd*)call.up(EG, 2 * P * 6); NEED(d_eg);
  cd* d_j = (cd*)call.up(j, 2 * P * 3); NEED(d_j);
  double* d_c1 = call.up(C1, 2 * P * 3); NEED(d_c1);
  double* d_c2 = call.up(C2, 2 * P * 3); NEED(d_c2);
  CHB_TRY(launch_maxwell_push(call.c.st, d_eg, d_j, nullptr, nullptr, d_c1, d_c2, 3, 1, P));
  CHB_TRY(call.down(EG, (double*)d_eg, 2 * P * 6));
  return call.sync();
}
int chimera_maxwell_init_push(double* EG, const double* j, const double* gn, const double* C1, const double* C2,
                              chb_i64 nkx, chb_i64 nkr, chb_i64 nm) {
  CALL_BEGIN();
  const i64 P = nkx * nkr * nm;
  cd* d_eg = (cd*)call.up(EG, 2 * P * 6); NEED(d_eg);
  cd* d_j = (cd*)call.up(j, 2 * P * 3); NEED(d_j);
  cd* d_gn = (cd*)call.up(gn, 2 * P * 3); NEED(d_gn);
  cd* d_c1 = (cd*)call.up(C1, 2 * P * 2); NEED(d_c1);
  cd* d_c2 = (cd*)call.up(C2, 2 * P * 2); NEED(d_c2);
  CHB_TRY(launch_maxwell_init_push(call.c.st, d_eg, d_j, d_gn, d_c1, d_c2, P));
  CHB_TRY(call.down(EG, (double*)d_eg, 2 * P * 6));
  return call.sync();
}
int chimera_poiss_corr(double* j, const double* gdj, const double* gn, const double* gp, double dt_inv,
                       const double* w2_inv, chb_i64 nkx, chb_i64 nkr, chb_i64 nm) {
  CALL_BEGIN();
  const i64 P = nkx * nkr * nm;
  cd* d_j = (cd*)call.up(j, 2 * P * 3); NEED(d_j);
  cd* d_gd = (cd*)call.up(gdj, 2 * P * 3); NEED(d_gd);
  cd* d_gn = (cd*)call.up(gn, 2 * P * 3); NEED(d_gn);
  cd* d_gp = (cd*)call.up(gp, 2 * P * 3); NEED(d_gp);
  double* d_w = call.up(w2_inv, P); NEED(d_w);
  CHB_TRY(launch_poiss_corr(call.c.st, d_j, d_gd, d_gn, d_gp, dt_inv, d_w, P));
  CHB_TRY(call.down(j, (double*)d_j, 2 * P * 3));
  return call.sync();
}
int chimera_poiss_corr_stat(double* j, const double* gdj, const double* gn, const double* DT, const double* w2_inv,
                            chb_i64 nkx, chb_i64 nkr, chb_i64 nm) {
  CALL_BEGIN();
  const i64 P = nkx * nkr * nm;
  cd* d_j = (cd*)call.up(j, 2 * P * 3); NEED(d_j);
  cd* d_gd = (cd*)call.up(gdj, 2 * P * 3); NEED(d_gd);
  cd* d_gn = (cd*)call.up(gn, 2 * P * 3); NEED(d_gn);
  cd* d_dt = (cd*)call.up(DT, 2 * nkx); NEED(d_dt);
  double* d_w = call.up(w2_inv, P); NEED(d_w);
  CHB_TRY(launch_poiss_corr_stat(call.c.st, d_j, d_gd, d_gn, d_dt, d_w, nkx, P));
  CHB_TRY(call.down(j, (double*)d_j, 2 * P * 3));
  return call.sync();
}
int chimera_field_drift(double* EG, const double* kx, double beta0, double dt, chb_i64 nkx, chb_i64 nkr, chb_i64 nm) {
  CALL_BEGIN();
  const i64 n = nkx * nkr * nm * 6;
  cd* d = (cd*)call.up(EG, 2 * n); NEED(d);
  double* d_kx = call.up(kx, nkx); NEED(d_kx);
  CHB_TRY(launch_field_drift(call.c.st, d, d_kx, beta0, dt, nkx, nkr * nm * 6));
  CHB_TRY(call.down(EG, (double*)d, 2 * n));
  return call.sync();
}
static int mult_host(double* v, const double* A, i64 P, int ncomp) {
  CALL_BEGIN();
  cd* d = (cd*)call.up(v, 2 * P * ncomp); NEED(d);
  double* d_a = call.up(A, P); NEED(d_a);
  CHB_TRY(launch_mult_real(call.c.st, d, d_a, P, ncomp));
  CHB_TRY(call.down(v, (double*)d, 2 * P * ncomp));
  return call.sync();
}
static int add_host(double* v, const double* A, i64 n) {
  CALL_BEGIN();
  cd* d = (cd*)call.up(v, 2 * n); NEED(d);
  cd* d_a = (cd*)call.up(A, 2 * n); NEED(d_a);
  CHB_TRY(launch_add(call.c.st, d, d_a, n));
  CHB_TRY(call.down(v, (double*)d, 2 * n));
  return call.sync();
}
int chimera_omp_mult_vec(double* v, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm) { return mult_host(v, A, nkx * nkr * nm, 3); }
int chimera_omp_mult_scl(double* v, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm) { return mult_host(v, A, nkx * nkr * nm, 1); }
int chimera_omp_add_vec(double* v, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm) { return add_host(v, A, nkx * nkr * nm * 3); }
int chimera_omp_add_scl(double* v, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm) { return add_host(v, A, nkx * nkr * nm); }

// ------------------------------------------------------------------ devices.f90
// one device over host arrays: coord (3,np), Fld (6,np) in-out, optional map a0(2,nx)
static int device_host(int kind, const double* coord, double* Fld, double t, double a0, const double* params, int nparams,
                       const double* map, chb_i64 nx, chb_i64 np) {
  CALL_BEGIN();
  double* d_x = call.up(coord, 3 * np); NEED(d_x);
  double* d_f = call.up(Fld, 6 * np); NEED(d_f);
  double* d_m = nullptr;
  if (map) { d_m = call.up(map, 2 * nx); NEED(d_m); }
  const DeviceSet u = one_device(kind, a0, params, nparams, d_m, (int)nx, t);
  CHB_TRY(launch_devices(call.c.st, aos((const double*)d_x, 3), aos(d_f, 6), u, np));
  CHB_TRY(call.down(Fld, d_f, 6 * np));
  return call.sync();
}
int chimera_undul_analytic(const double* coord, double* Fld, double t, const double* params, chb_i64 np) {
  return device_host(DEV_UNDUL_ANALYTIC, coord, Fld, t, 0.0, params, 4, nullptr, 0, np);
}
int chimera_undul_analytic_taper(const double* coord, double* Fld, double t, const double* params, chb_i64 np) {
  return device_host(DEV_UNDUL_ANALYTIC_TAPER, coord, Fld, t, 0.0, params, 5, nullptr, 0, np);
}
int chimera_undul_mapped(const double* coord, double* Fld, double t, const double* a0, const double* params, chb_i64 np,
                         chb_i64 nx) {
  if (nx < 3) { set_error("undul_mapped: nx >= 3 expected"); return 2; }
  return device_host(DEV_UNDUL_MAPPED, coord, Fld, t, 0.0, params, 3, a0, nx, np);
}
int chimera_undul_mapped_tap(const double* coord, double* Fld, double t, const double* a0, const double* params,
                             chb_i64 np, chb_i64 nx) {
  if (nx < 3) { set_error("undul_mapped_tap: nx >= 3 expected"); return 2; }
  return device_host(DEV_UNDUL_MAPPED_TAP, coord, Fld, t, 0.0, params, 5, a0, nx, np);
}
int chimera_planewave(const double* coord, double* Fld, double t, const double* params, chb_i64 np) {
  return device_host(DEV_PLANEWAVE, coord, Fld, t, 0.0, params, 7, nullptr, 0, np);
}
int chimera_gaussbeam(const double* coord, double* Fld, double time, double a0, const double* params, chb_i64 np) {
  return device_host(DEV_GAUSSBEAM, coord, Fld, time, a0, params, 8, nullptr, 0, np);
}

int chimera_gemm_profile(int on) { gemm_profile_enable(on); return 0; }
int chimera_fused_profile(int on) { fused_profile_enable(on); return 0; }
int chimera_fused_profile_read(unsigned long long* cycles8) {
  CHB_CUDA(cudaDeviceSynchronize());
  fused_profile_read(cycles8);
  return 0;
}
int chimera_gemm_profile_read(double* ms, double* flops, chb_i64* launches, int reset) {
  gemm_profile_read(ms, flops, launches, reset);
  return 0;
}

// ------------------------------------------------------------------ GEMM microbenchmark
int chimera_bench_gemm(chb_i64 nkx, chb_i64 K, chb_i64 N, int batch, int iters, double* ms) {
  CALL_BEGIN();
  if (batch < 1 || batch > kGemmMaxBatch) { set_error("bench_gemm: batch out of range"); return 2; }
  const i64 M = 2 * nkx;
  double* A = call.dev<double>(M * K * batch); NEED(A);
  double* C = call.dev<double>(M * N * batch); NEED(C);
  double* B = call.dev<double>(K * N); NEED(B);
  double* Bp = call.dev<double>(gemm_packed_size(K, N)); NEED(Bp);
  fill_pattern_k<<<grid_for(M * K * batch, 256), 256, 0, call.c.st>>>(A, M * K * batch, 0x9E3779B97F4A7C15ull);
  CHB_LAUNCH_CHECK();
  fill_pattern_k<<<grid_for(K * N, 256), 256, 0, call.c.st>>>(B, K * N, 0xD1B54A32D192ED03ull);
  CHB_LAUNCH_CHECK();
  CHB_TRY(launch_gemm_pack_b(call.c.st, Bp, B, K, N, K));
  GemmBatch gb;
  gb.count = batch;
  for (int b = 0; b < batch; ++b) gb.p[b] = GemmProblem{A + M * K * b, Bp, C + M * N * b, 1.0, 0.0};
  cudaEvent_t e0, e1;
  CHB_CUDA(cudaEventCreate(&e0));
  CHB_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) CHB_TRY(launch_gemm(call.c.st, gb, M, N, K, M, M));
  CHB_CUDA(cudaEventRecord(e0, call.c.st));
  for (int i = 0; i < iters; ++i) CHB_TRY(launch_gemm(call.c.st, gb, M, N, K, M, M));
  CHB_CUDA(cudaEventRecord(e1, call.c.st));
  CHB_CUDA(cudaEventSynchronize(e1));
  float t = 0;
  CHB_CUDA(cudaEventElapsedTime(&t, e0, e1));
  *ms = (double)t / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}


/* ---- f90/SR.f90 ---------------------------------------------------------------------------------- */
int chimera_sr_calc_far_tot(double* spect, const double* coords, const double* momenta_prv, const double* momenta_nxt,
                            const double* wghts, double dt, const double* omega, const double* SinTh,
                            const double* CosTh, const double* SinPh, const double* CosPh, chb_i64 nt, chb_i64 np,
                            chb_i64 nom, chb_i64 nth, chb_i64 nph) {
  return sr_far_host(spect, coords, momenta_prv, momenta_nxt, wghts, 0, dt, omega, SinTh, CosTh, SinPh, CosPh, nt, np, nom, nth, nph);
}
int chimera_sr_calc_far_comp(double* spect, const double* coords, const double* momenta_prv, const double* momenta_nxt,
                             const double* wghts, int comp, double dt, const double* omega, const double* SinTh,
                             const double* CosTh, const double* SinPh, const double* CosPh, chb_i64 nt, chb_i64 np,
                             chb_i64 nom, chb_i64 nth, chb_i64 nph) {
  /* SR.f90:219-227: any comp outside 1..3 integrates zero */
  return sr_far_host(spect, coords, momenta_prv, momenta_nxt, wghts, comp >= 1 && comp <= 3 ? comp : -1, dt, omega, SinTh, CosTh, SinPh, CosPh, nt, np, nom, nth, nph);
}
int chimera_sr_calc_near_tot(double* spect, const double* coords, const double* momenta, const double* wghts, double dt,
                             const double* omega, const double* Xgrid, const double* Ygrid, double z_scr, chb_i64 nt,
                             chb_i64 np, chb_i64 nom, chb_i64 nx, chb_i64 ny) {
  return sr_near_host(spect, coords, momenta, wghts, 0, dt, omega, Xgrid, Ygrid, nullptr, 0, z_scr, nt, np, nom, nx, ny);
}
int chimera_sr_calc_near_comp(double* spect, const double* coords, const double* momenta, const double* wghts, int comp,
                              double dt, const double* omega, const double* Xgrid, const double* Ygrid, double z_scr,
                              chb_i64 nt, chb_i64 np, chb_i64 nom, chb_i64 nx, chb_i64 ny) {
  if (comp < 1 || comp > 3) { set_error("sr_calc_near_comp: comp must be 1, 2 or 3 (got %d)", comp); return 2; }
  return sr_near_host(spect, coords, momenta, wghts, comp, dt, omega, Xgrid, Ygrid, nullptr, 0, z_scr, nt, np, nom, nx, ny);
}
int chimera_sr_calc_nearcirc_tot(double* spect, const double* coords, const double* momenta, const double* wghts,
                                 double dt, const double* omega, const double* Rgrid, const double* SinPh,
                                 const double* CosPh, double z_scr, chb_i64 nt, chb_i64 np, chb_i64 nom, chb_i64 nr,
                                 chb_i64 nph) {
  return sr_near_host(spect, coords, momenta, wghts, 0, dt, omega, Rgrid, SinPh, CosPh, 1, z_scr, nt, np, nom, nr, nph);
}
int chimera_sr_calc_nearcirc_comp(double* spect, const double* coords, const double* momenta, const double* wghts,
                                  int comp, double dt, const double* omega, const double* Rgrid, const double* SinPh,
                                  const double* CosPh, double z_scr, chb_i64 nt, chb_i64 np, chb_i64 nom, chb_i64 nr,
                                  chb_i64 nph) {
  if (comp < 1 || comp > 3) { set_error("sr_calc_nearcirc_comp: comp must be 1, 2 or 3 (got %d)", comp); return 2; }
  return sr_near_host(spect, coords, momenta, wghts, comp, dt, omega, Rgrid, SinPh, CosPh, 1, z_scr, nt, np, nom, nr, nph);
}

/* ---- f90/utils.f90 ------------------------------------------------------------------------------- */
int chimera_intens_profo(double* PWR_RO, const double* Fld, int NO, chb_i64 nxn, chb_i64 nrn, chb_i64 nm) {
  CALL_BEGIN();
  if (NO < 1 || nxn < 1 || nrn < 2 || nm < 1 || (nm % 2) == 0 || nm > 4096) {
    set_error("intens_profo: bad shape (NO=%d, Fld %lld x %lld x %lld x 3; mode slots must be 2*nkO+1)", NO, nxn, nrn, nm);
    return 2;
  }
  /* e^{i m theta_iO} by the reference's repeated multiplication (utils.f90:31-45) */
  const i64 nko = (nm - 1) / 2;
  const double pi = 4.0 * atan(1.0);
  std::vector<double> tab((size_t)(2 * NO * nm));
  auto T = [&](int o, i64 m) -> double* { return tab.data() + 2 * (o + (i64)NO * m); };
  double osr = 1.0, osi = 0.0;
  const double pr = cos(2.0 * pi / (double)(NO - 1)), pim = sin(2.0 * pi / (double)(NO - 1));
  for (int o = 0; o < NO; ++o) {
    if (o > 0) { const double r = osr * pr - osi * pim, i = osr * pim + osi * pr; osr = r; osi = i; }
    const double d = osr * osr + osi * osi, pmr = osr / d, pmi = -osi / d;
    T(o, nko)[0] = 1.0; T(o, nko)[1] = 0.0;
    for (i64 k = 1; k <= nko; ++k) {
      const double* a = T(o, nko + k - 1);
      T(o, nko + k)[0] = a[0] * osr - a[1] * osi; T(o, nko + k)[1] = a[0] * osi + a[1] * osr;
      const double* b = T(o, nko - k + 1);
      T(o, nko - k)[0] = b[0] * pmr - b[1] * pmi; T(o, nko - k)[1] = b[0] * pmi + b[1] * pmr;
    }
  }
  double* d_f = call.up(Fld, 2 * nxn * nrn * nm * 3); NEED(d_f);
  double* d_t = call.up(tab.data(), (i64)tab.size()); NEED(d_t);
  double* d_p = call.dev<double>((i64)NO * (nrn - 1)); NEED(d_p);
  CHB_TRY(launch_intens_profo(call.c.st, d_p, (const cd*)d_f, (const cd*)d_t, NO, nxn, nrn, nm));
  CHB_TRY(call.down(PWR_RO, d_p, (i64)NO * (nrn - 1)));
  return call.sync();
}
int chimera_density_2x(const double* x, const double* y, const double* wght, const double* grid, int bins_x,
                       int bins_y, double* dens, chb_i64 n_part) {
  CALL_BEGIN();
  if (bins_x < 1 || bins_y < 1 || n_part < 0) { set_error("density_2x: bins_x, bins_y >= 1 expected"); return 2; }
  double* d_x = call.up(x, n_part); NEED(d_x);
  double* d_y = call.up(y, n_part); NEED(d_y);
  double* d_w = call.up(wght, n_part); NEED(d_w);
  const i64 nd = (i64)(bins_x + 5) * (bins_y + 5);
  double* d_d = call.dev<double>(nd); NEED(d_d);
  CHB_TRY(launch_density_2x(call.c.st, d_d, d_x, d_y, d_w, grid, bins_x, bins_y, n_part));
  CHB_TRY(call.down(dens, d_d, nd));
  return call.sync();
}

}  // extern "C"
