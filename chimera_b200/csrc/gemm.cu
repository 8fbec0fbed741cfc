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
//     128-row grid would not fill the GPU twice over: the demo-size grids (Nx = 120 .. 528) and the 512-row kx slabs
//     of an 8-GPU run.
//   * A (the data) is staged by the TMA engine: one cp.async.bulk (UBLKCP) per k-row of the tile
//     into rows padded to 132 doubles, which makes the per-lane A-fragment reads (8 rows x 4 k)
//     bank-conflict free; completion is tracked with an mbarrier per stage (expect_tx).
//   * B (the operator, constant during a run) is pre-packed once into fragment order so a whole
//     16 x 64 tile is ONE contiguous 8 KB bulk copy and every B-fragment read is a conflict-free
//     LDS.64 at [tile][k4][n8][lane].
//   * 3-stage ring; a stage is refilled right after the CTA-wide barrier that retires it.
#include <utility>
#include <vector>
#include "common.cuh"
#include <atomic>
#include "kernels.cuh"

namespace chb {

namespace {
constexpr int BN = 64, BK = 16, STAGES = 3;
constexpr int B_STAGE = BK * BN;                 // doubles
constexpr int GEMM_THREADS = 256;
template <int BM> struct Tile {
  static constexpr int SA = BM + 4;              // padded row stride of the A tile (doubles)
  static constexpr int A_STAGE = BK * SA;        // doubles
  static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + STAGES * sizeof(uint64_t);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

template <int BM>
__global__ void __launch_bounds__(GEMM_THREADS, BM == 128 ? 2 : 4)
gemm_dmma_k(const __grid_constant__ GemmBatch batch, i64 M, i64 N, i64 K, i64 lda, i64 ldc, int KT) {
  constexpr int SA = Tile<BM>::SA, A_STAGE = Tile<BM>::A_STAGE;
  constexpr int WM = BM / 4, MF = WM / 8;  // warp tile rows, 8-row fragments per warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);

  const GemmProblem& pr = batch.p[blockIdx.z];
  const i64 m0 = (i64)blockIdx.x * BM;
  const i64 nt = blockIdx.y;
  const i64 n0 = nt * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;

  const int rows = (int)((M - m0 < BM) ? (M - m0) : BM);  // valid rows of this tile (even)
  const uint32_t row_bytes = (uint32_t)rows * 8u;

  // zero the A stages once: m-rows past M are never written by the copies.  k-rows past K of the last k-tile keep
  // whatever an earlier tile left in the stage (finite data); they meet the zero padding of the packed B tile
  for (int i = tid; i < STAGES * A_STAGE; i += GEMM_THREADS) sA[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  auto issue = [&](int kt) {  // executed by warp 0
    const int s = kt % STAGES;
    const i64 k0 = (i64)kt * BK;
    const int krows = (int)((K - k0 < BK) ? (K - k0) : BK);
    if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)krows * row_bytes + (uint32_t)(B_STAGE * sizeof(double)));
    __syncwarp();
    if (lane < krows)
      bulk_g2s(sA + s * A_STAGE + lane * SA, pr.A + m0 + (k0 + lane) * lda, row_bytes, &full[s]);
    if (lane == 31)
      bulk_g2s(sB + s * B_STAGE, pr.Bp + (nt * KT + kt) * (i64)B_STAGE, (uint32_t)(B_STAGE * sizeof(double)), &full[s]);
  };

  if (warp == 0)
    for (int kt = 0; kt < STAGES && kt < KT; ++kt) issue(kt);

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
    __syncthreads();  // every warp is done with stage s
    if (warp == 0 && kt + STAGES < KT) issue(kt + STAGES);
  }

  // epilogue: lane (g,t) owns C[8mf+g][8nf+2t..2t+1]
  const double alpha = pr.alpha, beta = pr.beta;
#pragma unroll
  for (int mf = 0; mf < MF; ++mf) {
    const i64 row = m0 + wm * WM + 8 * mf + g;
    if (row >= M) continue;
#pragma unroll
    for (int nf = 0; nf < 4; ++nf)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const i64 col = n0 + wn * 32 + 8 * nf + 2 * t + j;
        if (col >= N) continue;
        double* c = pr.C + row + ldc * col;
        double v = alpha * acc[mf][nf][j];
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
  // 64-row tiles when the 128-row grid is short of two full waves of its 2 CTAs per SM (148 SMs)
  const i64 ctas128 = ((M + 127) / 128) * ((N + BN - 1) / BN) * batch.count;
  const bool small = ctas128 < 4 * 148;
  {  // the attribute is per device: remember which devices have it (a process may drive several)
    static std::atomic<unsigned long long> attr_mask{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_mask.load(std::memory_order_relaxed) & bit)) {
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<128>::SMEM));
      CHB_CUDA(cudaFuncSetAttribute(gemm_dmma_k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<64>::SMEM));
      attr_mask.fetch_or(bit, std::memory_order_relaxed);
    }
  }
  const int KT = (int)((K + BK - 1) / BK);
  const int bm = small ? 64 : 128;
  dim3 grid((unsigned)((M + bm - 1) / bm), (unsigned)((N + BN - 1) / BN), (unsigned)batch.count);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.on) {
    e0 = g_prof.get(); e1 = g_prof.get();
    cudaEventRecord(e0, st);
  }
  if (small) gemm_dmma_k<64><<<grid, GEMM_THREADS, Tile<64>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
  else gemm_dmma_k<128><<<grid, GEMM_THREADS, Tile<128>::SMEM, st>>>(batch, M, N, K, lda, ldc, KT);
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
