// Shared helpers for the mmb200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/mmb200.h"

namespace mmb {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
// developer check (mmb_debug_smem_poison, include/mmb200_tools.h): after every launch
// fill the shared memory of every SM with NaN bit patterns on the legacy default stream
extern std::atomic<int> g_debug_poison;
void debug_poison_smem();

#define MMB_CHECK_CUDA(expr)                                                   \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      mmb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,              \
                     cudaGetErrorString(_e));                                  \
      return MMB_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

#define MMB_CHECK_LAUNCH()                                                     \
  do {                                                                         \
    mmb::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
    MMB_CHECK_CUDA(cudaGetLastError());                                        \
    if (mmb::g_debug_poison.load(std::memory_order_relaxed)) mmb::debug_poison_smem(); \
  } while (0)

#define MMB_REQUIRE(cond, msg)                                                 \
  do {                                                                         \
    if (!(cond)) {                                                             \
      mmb::set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, msg);   \
      return MMB_ERR_INVALID;                                                  \
    }                                                                          \
  } while (0)

// scipy.ndimage 'reflect' (d c b a | a b c d | d c b a): whole-sample symmetric
// extension with period 2n, valid for any distance outside the array.
__host__ __device__ __forceinline__ int reflect_index(int p, int n) {
  if (p >= 0 && p < n) return p;
  int m = 2 * n;
  p %= m;
  if (p < 0) p += m;
  return p < n ? p : m - 1 - p;
}

__host__ __device__ __forceinline__ int clamp_index(int p, int n) {
  return p < 0 ? 0 : (p >= n ? n - 1 : p);
}

// Workspace buffers are rewritten chunk after chunk at the same addresses and chunks
// may run on two streams at once, i.e. kernels of different chunks share an SM and its
// L1.  Readers of such buffers therefore load through L2 (ld.global.cg): a line cached
// in L1 by an earlier kernel must never be served to a later one.
__device__ __forceinline__ mmb_cand load_cand(const mmb_cand* p) {
  const int* q = reinterpret_cast<const int*>(p);
  mmb_cand c;
  c.z = __ldcg(q); c.y = __ldcg(q + 1); c.x = __ldcg(q + 2); c.s = __ldcg(q + 3);
  c.resp = __int_as_float(__ldcg(q + 4));
  return c;
}

constexpr int kMaxRadius = 64;   // templated fast paths cover radius <= 64

// Sampled Gaussian g[k] and second derivative h[k] for k = 0..r (both even),
// scipy.ndimage._filters._gaussian_kernel1d with truncate = 4.0.
struct LogWeights {
  float g[kMaxRadius + 1];
  float h[kMaxRadius + 1];
  float2 gh[kMaxRadius + 1];     // (g[k], h[k]) interleaved: one 64-bit constant load per tap
};

int gaussian_radius(double sigma);
// fills w (zero beyond r); returns r or -1 if r > kMaxRadius
int make_log_weights(double sigma, LogWeights* w);

// the sigma ladder travels in kernel parameter space (no device allocation, no
// host-buffer lifetime to manage on the asynchronous path)
constexpr int kMaxSigmas = 64;
struct SigmaLadder {
  double s[kMaxSigmas];
  int n;
};

// Cell grid of the pruning pair search (prune.cu): cubic cells of `cs` voxels, numbered
// z, y, x-major.
struct CellGrid {
  int cs;              // cell edge in voxels, >= the pair cut-off distance
  int ncz, ncy, ncx;
};

int num_sms();   // SM count of the current device (cached)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Optional per-kernel timing with CUDA events on the launching stream
// (mmb_profile_enable / mmb_profile_collect).  Kinds index mmb_profile_collect's
// output arrays.
enum ProfKind {
  PROF_TO_FLOAT = 0, PROF_PREPROCESS, PROF_LOG_X, PROF_LOG_Y, PROF_LOG_Z, PROF_LOCALMAX,
  PROF_PRUNE_EDGES, PROF_PRUNE_RESOLVE, PROF_COMPACT, PROF_SEAM, PROF_LOG_XY, PROF_NKINDS
};
bool prof_enabled();
void prof_begin(int kind, double units, cudaStream_t st);
void prof_end(cudaStream_t st);
// units of the scopes opened by this thread are multiplied by `f` until reset to 1: the
// sweeps work on pitch-padded rows but report TRUE voxels (X / pitch)
void prof_set_unit_scale(double f);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int kind, double units, cudaStream_t s) : st(s), on(prof_enabled()) {
    if (on) prof_begin(kind, units, st);
  }
  ~ProfScope() { if (on) prof_end(st); }
};

}  // namespace mmb
