// common.cuh -- shared helpers for the chimera_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

typedef long long i64;

namespace chb {

// ---- error plumbing: every C-ABI entry returns int (0 = ok); message via chimera_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define CHB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      chb::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),   \
                     cudaGetErrorString(_e));                                               \
      return 100 + (int)_e;                                                                 \
    }                                                                                       \
  } while (0)

#define CHB_TRY(call)            \
  do {                           \
    int _rc = (call);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// every kernel launch of this library is followed by CHB_LAUNCH_CHECK(): it doubles as the launch counter
extern long long g_launches;
#define CHB_LAUNCH_CHECK()        \
  do {                            \
    ++chb::g_launches;            \
    CHB_CUDA(cudaGetLastError()); \
  } while (0)

// ---- complex helpers on double2 (x = re, y = im)
typedef double2 cd;
__host__ __device__ __forceinline__ cd cmake(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cd cmul(cd a, cd b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cd cscale(double s, cd a) { return make_double2(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ cd cconj(cd a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cd cmuli(cd a) { return make_double2(-a.y, a.x); }  // i*a
__host__ __device__ __forceinline__ cd cneg(cd a) { return make_double2(-a.x, -a.y); }

// ---- particle array view: element (component c, particle ip) lives at p[c*cs + ip*ps].
//   reference layout (3,Np) Fortran order ("AoS"):  cs = 1,  ps = ncomp
//   engine layout, structure of arrays:             cs = capacity, ps = 1
struct PView {
  double* p;
  i64 cs, ps;
  __device__ __forceinline__ double& at(int c, i64 ip) const { return p[c * cs + ip * ps]; }
};
struct CPView {
  const double* p;
  i64 cs, ps;
  __device__ __forceinline__ double at(int c, i64 ip) const { return __ldg(p + c * cs + ip * ps); }
};
static inline PView aos(double* p, int ncomp) { return PView{p, 1, ncomp}; }
static inline CPView aos(const double* p, int ncomp) { return CPView{p, 1, ncomp}; }
static inline PView soa(double* p, i64 cap) { return PView{p, cap, 1}; }
static inline CPView soa(const double* p, i64 cap) { return CPView{p, cap, 1}; }
static inline CPView cview(PView v) { return CPView{v.p, v.cs, v.ps}; }

static inline unsigned grid_for(i64 n, int block) { return (unsigned)((n + block - 1) / block); }

constexpr int kMaxModes = 8;  // azimuthal-mode slots held in registers by particle kernels (|m| <= 7)

}  // namespace chb
