// Library-wide state and the fused per-chunk driver.
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace mmb {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional event profiling ---------------------------------------------------
struct ProfRec { int kind; double units; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<bool> g_prof_on{false};
static thread_local ProfRec t_open;

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void prof_begin(int kind, double units, cudaStream_t st) {
  t_open.kind = kind; t_open.units = units;
  cudaEventCreate(&t_open.a); cudaEventCreate(&t_open.b);
  cudaEventRecord(t_open.a, st);
}
void prof_end(cudaStream_t st) {
  cudaEventRecord(t_open.b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(t_open);
}

// implemented in the other translation units
int to_float_impl(const void* in, int dtype, const int64_t st[3], int Z, int Y, int X,
                  float* out, int64_t pitch, double scale, cudaStream_t s);
int preprocess_impl(const void* in, int dtype, const int64_t st[3], int Z, int Y, int X, int bz,
                    int by, int bx, const mmb_preproc_params* p, float* out, int64_t pitch,
                    cudaStream_t s);
int log_scale_impl(const float* in, float* out, float* work, int Z, int Y, int X, int64_t pitch,
                   double sigma, cudaStream_t st);
int localmax_impl(const float* prev, const float* cur, const float* next, int Z, int Y, int X,
                  int64_t pitch, int s, float thr, int z_lo, int z_hi, mmb_cand* out,
                  int capacity, int* counter, cudaStream_t st);
int prune_within_impl(const mmb_cand* cand, int n, const double* sigmas_host, int num_sigma,
                      double overlap, int Y, int X, uint8_t* keep, cudaStream_t st);

// stable stream compaction of the survivors to the front of a second buffer
__global__ void compact_kernel(const mmb_cand* __restrict__ in, const uint8_t* __restrict__ keep,
                               int n, mmb_cand* __restrict__ out, int* __restrict__ counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool k = i < n && keep[i];
  const unsigned ballot = __ballot_sync(0xffffffffu, k);
  if (!ballot) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(counter, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (k) out[base + __popc(ballot & ((1u << lane) - 1u))] = in[i];
}

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_version(void) { return MMB_VERSION; }
extern "C" const char* mmb_last_error(void) { return g_err; }
extern "C" int64_t mmb_launch_count(void) { return g_launches.load(); }

extern "C" int mmb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on.store(on != 0);
  return MMB_OK;
}

extern "C" int mmb_profile_collect(double* ms, int64_t* launches, double* units) {
  MMB_REQUIRE(ms && launches && units, "null output");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int k = 0; k < PROF_NKINDS; ++k) { ms[k] = 0.0; launches[k] = 0; units[k] = 0.0; }
  for (auto& r : g_prof) {
    MMB_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    MMB_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.kind] += t; launches[r.kind] += 1; units[r.kind] += r.units;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return MMB_OK;
}

// work layout: [F][ring0][ring1][ring2][A][B][C][D][cand2 (capacity)][keep][counter]
extern "C" int64_t mmb_detect_work_bytes(int Z, int Y, int64_t pitch, int capacity) {
  const int64_t vol = align256((int64_t)Z * Y * pitch * (int64_t)sizeof(float));
  return 8 * vol + align256((int64_t)capacity * (int64_t)sizeof(mmb_cand)) +
         align256(capacity) + 256;
}

extern "C" int mmb_detect_chunk(const void* in, int dtype, const int64_t in_strides[3], int Z,
                                int Y, int X, int64_t pitch, double scale,
                                const mmb_preproc_params* pre, int bz, int by, int bx,
                                const double* sigmas, int num_sigma, double threshold,
                                double overlap, int z_lo, int z_hi, void* work, mmb_cand* cand,
                                int capacity, int* n_out, int* n_peaks, void* stream) {
  MMB_REQUIRE(in && in_strides && sigmas && work && cand && n_out, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(Y <= 65535 && Z <= 65535, "Y and Z must be <= 65535");
  MMB_REQUIRE(num_sigma > 0 && capacity > 0, "bad sizes");
  MMB_REQUIRE(z_lo >= 0 && z_hi <= Z && z_lo <= z_hi, "bad z range");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t vol_b = align256((int64_t)Z * Y * pitch * (int64_t)sizeof(float));
  char* base = (char*)work;
  float* F = (float*)base;
  float* ring[3] = {(float*)(base + vol_b), (float*)(base + 2 * vol_b), (float*)(base + 3 * vol_b)};
  float* lw = (float*)(base + 4 * vol_b);          // A,B,C,D contiguous (vol_b is 256-aligned)
  // log_scale_impl strides its work area by Z*Y*pitch floats, which fits in vol_b each
  char* tail = base + 8 * vol_b;
  mmb_cand* cand2 = (mmb_cand*)tail;
  uint8_t* keep = (uint8_t*)(tail + align256((int64_t)capacity * (int64_t)sizeof(mmb_cand)));
  int* counter = (int*)((char*)keep + align256(capacity));

  int rc;
  if (pre) rc = preprocess_impl(in, dtype, in_strides, Z, Y, X, bz, by, bx, pre, F, pitch, st);
  else rc = to_float_impl(in, dtype, in_strides, Z, Y, X, F, pitch, scale, st);
  if (rc) return rc;

  MMB_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  const float thr = (float)threshold;
  // ring slot of scale i is i % 3; local maxima of scale i-1 run once scale i exists
  for (int i = 0; i < num_sigma; ++i) {
    rc = log_scale_impl(F, ring[i % 3], lw, Z, Y, X, pitch, sigmas[i], st);
    if (rc) return rc;
    if (i >= 1) {
      const float* prev = i >= 2 ? ring[(i - 2) % 3] : nullptr;
      rc = localmax_impl(prev, ring[(i - 1) % 3], ring[i % 3], Z, Y, X, pitch, i - 1, thr, z_lo,
                         z_hi, cand2, capacity, counter, st);
      if (rc) return rc;
    }
  }
  {
    const int i = num_sigma - 1;
    const float* prev = i >= 1 ? ring[(i - 1) % 3] : nullptr;
    rc = localmax_impl(prev, ring[i % 3], nullptr, Z, Y, X, pitch, i, thr, z_lo, z_hi, cand2,
                       capacity, counter, st);
    if (rc) return rc;
  }
  int n = 0;
  MMB_CHECK_CUDA(cudaMemcpyAsync(&n, counter, sizeof(int), cudaMemcpyDeviceToHost, st));
  MMB_CHECK_CUDA(cudaStreamSynchronize(st));
  if (n_peaks) *n_peaks = n;
  if (n > capacity) {
    set_error("candidate buffer overflow: %d local maxima, capacity %d", n, capacity);
    *n_out = n;
    return MMB_ERR_OVERFLOW;
  }
  if (n == 0) { *n_out = 0; return MMB_OK; }
  rc = prune_within_impl(cand2, n, sigmas, num_sigma, overlap, Y, X, keep, st);
  if (rc) return rc;
  MMB_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  {
    ProfScope ps(PROF_COMPACT, n, st);
    compact_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cand2, keep, n, cand, counter);
  }
  MMB_CHECK_LAUNCH();
  MMB_CHECK_CUDA(cudaMemcpyAsync(n_out, counter, sizeof(int), cudaMemcpyDeviceToHost, st));
  MMB_CHECK_CUDA(cudaStreamSynchronize(st));
  return MMB_OK;
}
