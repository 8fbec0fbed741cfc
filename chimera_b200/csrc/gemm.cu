// gemm.cu -- the discrete-Hankel-transform / mode-coupling contraction on FP64 tensor cores.
//
// Every spectral contraction of the reference (fb_io.f90:46-50, :207-214; fb_math.f90:49-52 ...)
// is "complex slab times real operator": with x the fastest axis a complex [Nx x K] slab is a
// real column-major [2Nx x K] matrix, so the work is the real GEMM
//     C[M x N] = alpha * A[M x K] * B[K x N] + beta * C ,   M = 2 Nx, B = In | Out | DpS2S | DmS2S.
//
// sm_100a mapping.  tcgen05 has no f64 kind, so FP64 tensor work is mma.sync m8n8k4 (SASS DMMA).
//   * CTA tile 128 x 64 x 16, 8 warps as 4(M) x 2(N), warp tile 32 x 32 = 16 DMMA per k4 step,
//     accumulators in registers (64 per lane).  A 64 x 64 x 16 variant (warp tile 16 x 32) serves the launches whose
//     128-row grid would not fill the GPU twice over (the 512-row kx slabs of an 8-GPU run), a 32 x 64 x 16 variant
//     (warp tile 8 x 32) the demo-size grids (Nx = 120 .. 528) whose 64-row grid is under two CTAs per SM.
//   * A (the data) is staged by the TMA engine: one cp.async.bulk (UBLKCP) per k-row of the tile
//     into rows padded to 132 doubles, which makes the per-lane A-fragment reads (8 rows x 4 k)
//     bank-conflict free; completion is tracked with an mbarrier per stage (expect_tx).
//   * B (the operator, constant during a run) is pre-packed once into fragment order so a whole
//     16 x 64 tile is ONE contiguous 8 KB bulk copy and every B-fragment read is a conflict-free
//     LDS.64 at [tile][k4][n8][lane].
//   * warp-specialised ring: a ninth warp issues the copies and runs ahead by the ring depth (4 stages of the 128-row
//     tile, 6 of the 64-row tile); a "full" mbarrier (transaction count) per stage hands it to the 8 DMMA warps and an
//     "empty" mbarrier (one arrival per warp) hands it back, so there is no CTA-wide barrier in the k loop.  Matters
//     where a launch is one CTA per SM or less (the demo grids): a 304 x 300 x 300 launch of 50 CTAs went from 23.1 to 18.5 us, the 4096 x 512 x 512 launches from 33.2 to 34.8 TF/s.
#include <utility>
#include <vector>
#include "common.cuh"
#include <atomic>
#include "kernels.cuh"

namespace chb {

namespace {
constexpr int BN = 64, BK = 16;
constexpr int B_STAGE = BK * BN;                 // doubles
constexpr int GEMM_CONSUMERS = 256;              // 8 DMMA warps
#ifndef CHB_GEMM_NPROD_SMALL
#define CHB_GEMM_NPROD_SMALL 4
#endif
#ifndef CHB_GEMM_NS128
#define CHB_GEMM_NS128 4
#endif
#ifndef CHB_GEMM_NS64
#define CHB_GEMM_NS64 6
#endif
#ifndef CHB_GEMM_NS32
#define CHB_GEMM_NS32 4
#endif
template <int BM> struct Tile {
  static constexpr int STAGES = BM == 128 ? CHB_GEMM_NS128 : (BM == 64 ? CHB_GEMM_NS64 : CHB_GEMM_NS32);
  // copy warps: the bulk copies of a warp leave one lane at a time (UBLKCP takes uniform registers), ~50 cycles each, so
  // the 17 copies of a k-tile cost one warp ~800 cycles -- hidden behind the 2048 DMMA cycles of a 128-row tile with two
  // CTAs per SM, but longer than the DMMA work of a 32- or 64-row tile on a lightly loaded SM: those split the rows
  // over 4 copy warps (one per SM sub-partition)
  static constexpr int NPROD = BM == 128 ? 1 : CHB_GEMM_NPROD_SMALL;
  static constexpr int THREADS = GEMM_CONSUMERS + 32 * NPROD;
  static constexpr int SA = BM + 4;              // padded row stride of the A tile (doubles)
  static constexpr int A_STAGE = BK * SA;        // doubles
  static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + 2 * STAGES * sizeof(uint64_t);
  static constexpr int CTAS_PER_SM = (int)((227 * 1024) / (SMEM + 1024)) < 4 ? (int)((227 * 1024) / (SMEM + 1024)) : 4;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// pack B[K x N] (column-major, ldb) into [n-tile][k-tile][k4 (4)][n8 (8)][lane (32)], zero padded
__global__ void __launch_bounds__(256) gemm_pack_b_k(double* __restrict__ Bp, const double* __restrict__ B, i64 K,
                                                     i64 N, i64 ldb, i64 KT, i64 total) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int lane = (int)(e & 31);
  const int nb = (int)((e >> 5) & 7);
  const int ks = (int)((e >> 8) & 3);
  const i64 tile = e >> 10;
  const i64 kt = tile % KT, nt = tile / KT;
  const int g = lane >> 2, t = lane & 3;
  const i64 k = kt * BK + 4 * ks + t;
  const i64 n = nt * BN + 8 * nb + g;
  Bp[e] = (k < K && n < N) ? B[k + ldb * n] : 0.0;
}

template <int BM, bool PHASE>
__global__ void __launch_bounds__(Tile<BM>::THREADS, Tile<BM>::CTAS_PER_SM)
gemm_dmma_k(const __grid_constant__ GemmBatch batch, i64 M, i64 N, i64 K, i64 lda, i64 ldc, int KT) {
  constexpr int STAGES = Tile<BM>::STAGES, SA = Tile<BM>::SA, A_STAGE = Tile<BM>::A_STAGE;
  constexpr int NPROD = Tile<BM>::NPROD, GEMM_THREADS = Tile<BM>::THREADS;
  constexpr int WM = BM / 4, MF = WM / 8;  // warp tile rows, 8-row fragments per warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);  // copies of a stage have landed
  uint64_t* empty = full + STAGES;                                      // the 8 DMMA warps are done with a stage

  const GemmProblem& pr = batch.p[blockIdx.z];
  const i64 m0 = (i64)blockIdx.x * BM;
  const i64 nt = blockIdx.y;
  const i64 n0 = nt * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int rows = (int)((M - m0 < BM) ? (M - m0) : BM);  // valid rows of this tile (even)
  const uint32_t row_bytes = (uint32_t)rows * 8u;

  // zero the A stages once: m-rows past M are never written by the copies.  k-rows past K of the last k-tile keep
  // whatever an earlier tile left in the stage (finite data); they meet the zero padding of the packed B tile
  for (int i = tid; i < STAGES * A_STAGE; i += GEMM_THREADS) sA[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GEMM_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (warp >= GEMM_CONSUMERS / 32) {
    // the copy warps: run ahead of the DMMA warps by up to STAGES k-tiles, no CTA-wide barrier in the loop.  Copy warp
    // pw moves k-rows pw, pw + NPROD, ..; warp 0 also posts the transaction count and moves the B tile (a copy that
    // lands before the count is posted only drives the count negative until then)
    const int pw = warp - GEMM_CONSUMERS / 32;
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt % STAGES;
      if (kt >= STAGES) mbar_wait(&empty[s], (uint32_t)((kt / STAGES - 1) & 1));
      const i64 k0 = (i64)kt * BK;
      const int krows = (int)((K - k0 < BK) ? (K - k0) : BK);
      if (pw == 0 && lane == 0)
        mbar_expect_tx(&full[s], (uint32_t)krows * row_bytes + (uint32_t)(B_STAGE * sizeof(double)));
      __syncwarp();
      const int kr = pw + NPROD * lane;
      if (kr < krows) bulk_g2s(sA + s * A_STAGE + kr * SA, pr.A + m0 + (k0 + kr) * lda, row_bytes, &full[s]);
      if (pw == 0 && lane == 31)
        bulk_g2s(sB + s * B_STAGE, pr.Bp + (nt * KT + kt) * (i64)B_STAGE, (uint32_t)(B_STAGE * sizeof(double)), &full[s]);
    }
    return;
  }

  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  double acc[MF][4][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % STAGES;
    mbar_wait(&full[s], (uint32_t)((kt / STAGES) & 1));
    const double* a_s = sA + s * A_STAGE + wm * WM + g;
    const double* b_s = sB + s * B_STAGE + wn * 4 * 32 + lane;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      double a[MF], b[4];
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) a[mf] = a_s[(4 * ks + t) * SA + 8 * mf];
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) b[nf] = b_s[(ks * 8 + nf) * 32];
#pragma unroll
      for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) dmma(acc[mf][nf][0], acc[mf][nf][1], a[mf], b[nf]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // this warp has read everything it needs from stage s
  }

  // epilogue: lane (g,t) owns C[8mf+g][8nf+2t..2t+1]; real rows 2i, 2i+1 are the re / im of complex row i, so the other
  // half of a complex value sits in lane ^ 4
  const double alpha = pr.alpha, beta = pr.beta;
  constexpr bool phase = PHASE;
#pragma unroll
  for (int mf = 0; mf < MF; ++mf) {
    const i64 row = m0 + wm * WM + 8 * mf + g;
    const bool row_ok = row < M;
    double pc = 1.0, ps = 0.0;
    if (phase && row_ok) {
      sincos(batch.phase_leftX * __ldg(batch.phase_kx + (row >> 1)), &ps, &pc);
      ps *= batch.phase_sign;
    }
#pragma unroll
    for (int nf = 0; nf < 4; ++nf)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const i64 col = n0 + wn * 32 + 8 * nf + 2 * t + j;
        double v = alpha * acc[mf][nf][j];
        if (phase) {
          const double o = __shfl_xor_sync(0xffffffffu, v, 4);
          // (re + i im)(pc + i ps): this lane holds re (even row) or im (odd row), `o` is the other one
          v = (row & 1) ? fma(v, pc, o * ps) : fma(v, pc, -(o * ps));
          if (pr.fact && row_ok && col < N) v *= __ldg(pr.fact + (row >> 1) + (M >> 1) * col);
        }
        if (!row_ok || col >= N) continue;
        double* c = pr.C + row + ldc * col;
        if (beta != 0.0) v += beta * (*c);
        *c = v;
      }
  }
}

}  // namespace

i64 gemm_packed_size(i64 K, i64 N) {
  const i64 KT = (K + BK - 1) / BK, NT = (N + BN - 1) / BN;
  return KT * NT * (i64)B_STAGE;
}

int launch_gemm_pack_b(cudaStream_t st, double* Bp, const double* B, i64 K, i64 N, i64 ldb) {
  const i64 KT = (K + BK - 1) / BK;
  const i64 total = gemm_packed_size(K, N);
  gemm_pack_b_k<<<grid_for(total, 256), 256, 0, st>>>(Bp, B, K, N, ldb, KT, total);
  CHB_LAUNCH_CHECK();
  return 0;
}

// optional per-launch timing of the contraction kernel (CUDA events on the launching stream), read by
// bench.py for the FP64-tensor roofline of the DHT / mode-coupling GEMMs
namespace {
struct GemmProf {
  int on = 0;
  double flops = 0.0, ms = 0.0;
  long long launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void drain() {
    for (auto& pe : pending) {
      cudaEventSynchronize(pe.second);
      float t = 0;
      cudaEventElapsedTime(&t, pe.first, pe.second);
      ms += t;
      pool.push_back(pe.first);
      pool.push_back(pe.second);
    }
    pending.clear();
  }
} g_prof;
}  // namespace

void gemm_profile_enable(int on) {
  g_prof.drain();
  g_prof.on = on;
}
int gemm_profile_enabled() { return g_prof.on; }
void gemm_profile_read(double* ms, double* flops, long long* launches, int reset) {
  g_prof.drain();
  if (ms) *ms = g_prof.ms;
  if (flops) *flops = g_prof.flops;
  if (launches) *launches = g_prof.launches;
  if (reset) { g_prof.ms = 0; g_prof.flops = 0; g_prof.launches = 0; }
}

int launch_gemm(cudaStream_t st, const GemmBatch& batch, i64 M, i64 N, i64 K, i64 lda, i64 ldc) {
  if (batch.count <= 0 || M <= 0 || N <= 0) return 0;
  if (batch.count > kGemmMaxBatch) { set_error("gemm batch too large"); return 4; }
  if ((M & 1) || (lda & 1)) { set_error("gemm: M and lda must be even (complex-interleaved rows)"); return 4; }
  if (batch.phase_kx)
    for (int b = 0; b < batch.count; ++b)
      if (batch.p[b].beta != 0.0) { set_error("gemm: the phase epilogue needs beta = 0"); return 4; }
  // Tile height by a wave model of the launch: the busiest of the 148 SMs gets L = ceil(CTAs / 148) tiles and works
  // through them at the speed of its FP64 tensor pipe, so the launch costs about L x rows x f, f the relative cost of a
  // smaller tile (more operator traffic and fragment loads per DMMA); an SM holding a single CTA runs its pipe at about
  // 60 % (two DMMA warps per sub-partition do not cover the DMMA latency).  Measured (tools/gemm_sweep.py): the
  // 304 x 300 x 300 launches of the space-charge demo take 10.3 / 16.4 / 51 us (batch 1 / 3 / 12) on 32-row tiles
  // against 18.5 / 26.7 / 59 us on 64-row tiles; the 512-row slabs of an 8-GPU run stay on 64, the LWFA grid on 128.
  const i64 ncol_tiles = ((N + BN - 1) / BN) * batch.count;
  auto cost = [&](int rows_per_tile, double f) {
    const i64 ctas = ((M + rows_per_tile - 1) / rows_per_tile) * ncol_tiles;
    const i64 L = (ctas + 147) / 148;
    return (double)L * rows_per_tile * f * (L == 1 ? 1.6 : 1.0);
  };
  const double c128 = cost(128, 1.0), c64 = cost(64, 1.1), c32 = cost(32, 1.2);
  const bool tiny = c32 < c64 && c32 < c128;
  const bool small = !tiny && c64 < c128;
  {  // the attribute is per device: remember which devices have it (a process may drive several)
    static std::atomic<unsigned long long> attr_mask{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_mask.load(std::memory_order_relaxed) & bit)) {
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<128>::SMEM));
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<128>::SMEM));
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<64>::SMEM));
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<64>::SMEM));
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<32>::SMEM));
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<32>::SMEM));
      attr_mask.fetch_or(bit, std::memory_order_relaxed);
    }
  }
  const int KT = (int)((K + BK - 1) / BK);
  const int bm = tiny ? 32 : (small ? 64 : 128);
  dim3 grid((unsigned)((M + bm - 1) / bm), (unsigned)((N + BN - 1) / BN), (unsigned)batch.count);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.on) {
    e0 = g_prof.get(); e1 = g_prof.get();
    cudaEventRecord(e0, st);
  }
  const bool ph = batch.phase_kx != nullptr;
  if (tiny && ph) gemm_dmma_k<32, true><<<grid, Tile<32>::THREADS, Tile<32>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
  else if (tiny) gemm_dmma_k<32, false><<<grid, Tile<32>::THREADS, Tile<32>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
  else if (small && ph) gemm_dmma_k<64, true><<<grid, Tile<64>::THREADS, Tile<64>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
  else if (small) gemm_dmma_k<64, false><<<grid, Tile<64>::THREADS, Tile<64>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
  else if (ph) gemm_dmma_k<128, true><<<grid, Tile<128>::THREADS, Tile<128>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
  else gemm_dmma_k<128, false><<<grid, Tile<128>::THREADS, Tile<128>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
  if (g_prof.on) {
    cudaEventRecord(e1, st);
    g_prof.pending.push_back({e0, e1});
    g_prof.flops += 2.0 * (double)M * (double)N * (double)K * batch.count;
    g_prof.launches += 1;
    if (g_prof.pending.size() > 2048) g_prof.drain();
  }
  CHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace chb
