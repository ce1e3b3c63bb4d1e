// Library-wide state and the fused per-chunk driver.
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <vector>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "tma.cuh"

namespace mmb {

int encode_tensor_map_f32(CUtensorMap* map, const float* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  });
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MMB_ERR_CUDA;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base,
                            d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1));
    return MMB_ERR_CUDA;
  }
  return MMB_OK;
}

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_debug_poison{0};

// one CTA per SM, 200 KB of dynamic shared memory each, filled with 0xFFFFFFFF (NaN as
// float, -1 as int): a kernel that consumes shared memory it never wrote then computes
// differently from a run whose shared memory still holds a predecessor's finite values
__global__ void __launch_bounds__(1024) poison_smem_kernel(unsigned pattern, unsigned* sink) {
  extern __shared__ unsigned poison_s[];
  const int n = 200 * 1024 / 4;
  for (int i = threadIdx.x; i < n; i += 1024) poison_s[i] = pattern;
  __syncthreads();
  if (pattern == 1u && poison_s[(threadIdx.x * 97) % n] != pattern) *sink = 1;   // keep the stores
}

void debug_poison_smem() {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         200 * 1024);
    configured = true;
  }
  poison_smem_kernel<<<num_sms(), 1024, 200 * 1024, 0>>>(0xFFFFFFFFu, nullptr);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

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

// events are recycled: creating two per launch would put thousands of
// cudaEventCreate calls inside a timed step
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t take_event() {
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_event_pool.empty()) {
      cudaEvent_t e = g_event_pool.back();
      g_event_pool.pop_back();
      return e;
    }
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
static thread_local double t_unit_scale = 1.0;
void prof_set_unit_scale(double f) { t_unit_scale = f; }
void prof_begin(int kind, double units, cudaStream_t st) {
  t_open.kind = kind; t_open.units = units * t_unit_scale;
  t_open.a = take_event(); t_open.b = take_event();
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
                    cudaStream_t s, void* scratch, int64_t scratch_bytes);
int log_scale_impl(const float* in, float* out, float* work, int Z, int Y, int X, int64_t pitch,
                   double sigma, cudaStream_t st);
int localmax_impl(const float* prev, const float* cur, const float* next, int Z, int Y, int X,
                  int64_t pitch, int s, float thr, int z_lo, int z_hi, mmb_cand* out,
                  int capacity, int* counter, cudaStream_t st);
int prune_within_impl(const mmb_cand* cand, int n, const double* sigmas_host, int num_sigma,
                      double overlap, int Y, int X, uint8_t* keep, cudaStream_t st, int z_sorted);
int make_ladder(const double* sigmas_host, int num_sigma, SigmaLadder* out);
CellGrid make_cell_grid(int Z, int Y, int X, int min_edge, int64_t max_cells);
int64_t cell_count(const CellGrid& g);
int prune_cell_edge(const SigmaLadder& ladder);
int prune_within_enqueue(const mmb_cand* cand, const int* n_ptr, int n_max,
                         const SigmaLadder& ladder, double overlap, int Y, int X, int2* edges,
                         int edge_cap, int* edge_count, unsigned char* state, uint8_t* keep,
                         cudaStream_t st, const CellGrid& grid, const int* cell_end,
                         int* od_count);
int sort_by_cell_enqueue(const mmb_cand* cand, const int* n_ptr, int n_max, const CellGrid& grid,
                         int* hist, mmb_cand* out, cudaStream_t st);

// stable stream compaction of the survivors to the front of a second buffer
__global__ void compact_kernel(const mmb_cand* __restrict__ in, const uint8_t* __restrict__ keep,
                               const int* __restrict__ n_ptr, int n_max,
                               mmb_cand* __restrict__ out, int* __restrict__ counter) {
  const int n = min(__ldcg(n_ptr), n_max);
  if (blockIdx.x * blockDim.x >= n) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool k = i < n && __ldcg(keep + i);
  const unsigned ballot = __ballot_sync(0xffffffffu, k);
  if (!ballot) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(counter, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (k) out[base + __popc(ballot & ((1u << lane) - 1u))] = load_cand(in + i);
}

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_version(void) { return MMB_VERSION; }
extern "C" int mmb_debug_smem_poison(int on) { g_debug_poison.store(on != 0); return MMB_OK; }
extern "C" const char* mmb_last_error(void) { return g_err; }
extern "C" int64_t mmb_launch_count(void) { return g_launches.load(); }

extern "C" int mmb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { g_event_pool.push_back(r.a); g_event_pool.push_back(r.b); }
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
    g_event_pool.push_back(r.a); g_event_pool.push_back(r.b);
  }
  g_prof.clear();
  return MMB_OK;
}

// ---- fused per-chunk driver -------------------------------------------------------
// work layout:
// [F][ring0][ring1][ring2][A][B][C][D][cand2][cand3][zhist][keep][edges][state][counters]
struct ChunkLayout {
  int64_t vol_b, cand_b, zhist_b, keep_b, edges_b, state_b, total;
  int edge_cap;
};

static ChunkLayout chunk_layout(int Z, int Y, int64_t pitch, int capacity) {
  ChunkLayout L;
  L.vol_b = align256((int64_t)Z * Y * pitch * (int64_t)sizeof(float));
  L.cand_b = align256((int64_t)capacity * (int64_t)sizeof(mmb_cand));
  // cell histogram of the pair search: cells are never narrower than 8 voxels
  L.zhist_b = align256((cdiv(Z, 8) * cdiv(Y, 8) * cdiv(pitch, 8) + 1) * (int64_t)sizeof(int));
  L.keep_b = align256(capacity);
  L.edge_cap = 4 * capacity + 4096;
  L.edges_b = align256((int64_t)L.edge_cap * (int64_t)sizeof(int2));
  L.state_b = align256(2 * ((int64_t)(capacity + 3) / 4 * 4 + 4));
  L.total = 8 * L.vol_b + 2 * L.cand_b + L.zhist_b + L.keep_b + L.edges_b + L.state_b + 256;
  return L;
}

extern "C" int64_t mmb_detect_work_bytes(int Z, int Y, int64_t pitch, int capacity) {
  return chunk_layout(Z, Y, pitch, capacity).total;
}

// status (device int32[4]): [0] local maxima found (may exceed capacity), [1] survivors
// written to `cand`, [2] kill edges found (may exceed mmb_detect_edge_capacity(capacity)),
// [3] size of the set whose survival depends on scikit-image's pair iteration order
extern "C" int mmb_detect_chunk_enqueue(const void* in, int dtype, const int64_t in_strides[3],
                                        int Z, int Y, int X, int64_t pitch, double scale,
                                        const mmb_preproc_params* pre, int bz, int by, int bx,
                                        const double* sigmas, int num_sigma, double threshold,
                                        double overlap, int z_lo, int z_hi, void* work,
                                        mmb_cand* cand, int capacity, int32_t* status,
                                        void* stream) {
  MMB_REQUIRE(in && in_strides && sigmas && work && cand && status, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(Y <= 65535 && Z <= 65535, "Y and Z must be <= 65535");
  MMB_REQUIRE(num_sigma > 0 && capacity > 0, "bad sizes");
  MMB_REQUIRE(z_lo >= 0 && z_hi <= Z && z_lo <= z_hi, "bad z range");
  cudaStream_t st = (cudaStream_t)stream;
  SigmaLadder ladder;
  int rc = make_ladder(sigmas, num_sigma, &ladder);
  if (rc) return rc;
  const ChunkLayout L = chunk_layout(Z, Y, pitch, capacity);
  char* base = (char*)work;
  float* F = (float*)base;
  float* ring[3] = {(float*)(base + L.vol_b), (float*)(base + 2 * L.vol_b),
                    (float*)(base + 3 * L.vol_b)};
  float* lw = (float*)(base + 4 * L.vol_b);       // A,B,C,D (log_scale_impl strides by Z*Y*pitch)
  char* tail = base + 8 * L.vol_b;
  mmb_cand* cand2 = (mmb_cand*)tail;                    tail += L.cand_b;
  mmb_cand* cand3 = (mmb_cand*)tail;                    tail += L.cand_b;
  int* zhist = (int*)tail;                              tail += L.zhist_b;
  uint8_t* keep = (uint8_t*)tail;                       tail += L.keep_b;
  int2* edges = (int2*)tail;                            tail += L.edges_b;
  unsigned char* state = (unsigned char*)tail;          tail += L.state_b;
  int* counters = (int*)tail;      // [0] peaks, [1] survivors, [2] edges, [3] order-dependent set

  // blocks above 32 voxels take the global-memory path, which borrows the (still idle)
  // sweep buffers A..D as scratch
  if (pre) rc = preprocess_impl(in, dtype, in_strides, Z, Y, X, bz, by, bx, pre, F, pitch, st, lw,
                                4 * L.vol_b);
  else rc = to_float_impl(in, dtype, in_strides, Z, Y, X, F, pitch, scale, st);
  if (rc) return rc;

  MMB_CHECK_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(int), st));
  const float thr = (float)threshold;
  // ring slot of scale i is i % 3; local maxima of scale i-1 run once scale i exists
  for (int i = 0; i < num_sigma; ++i) {
    rc = log_scale_impl(F, ring[i % 3], lw, Z, Y, X, pitch, sigmas[i], st);
    if (rc) return rc;
    if (i >= 1) {
      const float* prev = i >= 2 ? ring[(i - 2) % 3] : nullptr;
      rc = localmax_impl(prev, ring[(i - 1) % 3], ring[i % 3], Z, Y, X, pitch, i - 1, thr, z_lo,
                         z_hi, cand2, capacity, counters, st);
      if (rc) return rc;
    }
  }
  {
    const int i = num_sigma - 1;
    const float* prev = i >= 1 ? ring[(i - 1) % 3] : nullptr;
    rc = localmax_impl(prev, ring[i % 3], nullptr, Z, Y, X, pitch, i, thr, z_lo, z_hi, cand2,
                       capacity, counters, st);
    if (rc) return rc;
  }
  // list the local maxima cell by cell (cells as wide as the pruning cut-off) so that
  // the pair search visits the 27 cells around a candidate instead of every pair
  const int cell_edge = prune_cell_edge(ladder) < 8 ? 8 : prune_cell_edge(ladder);
  const CellGrid cgrid = make_cell_grid(Z, Y, X, cell_edge, L.zhist_b / (int64_t)sizeof(int) - 1);
  {
    ProfScope ps(PROF_COMPACT, capacity, st);
    rc = sort_by_cell_enqueue(cand2, counters, capacity, cgrid, zhist, cand3, st);
    if (rc) return rc;
  }
  rc = prune_within_enqueue(cand3, counters, capacity, ladder, overlap, Y, X, edges, L.edge_cap,
                            counters + 2, state, keep, st, cgrid, zhist, counters + 3);
  if (rc) return rc;
  {
    ProfScope ps(PROF_COMPACT, capacity, st);
    compact_kernel<<<(unsigned)cdiv(capacity, 256), 256, 0, st>>>(cand3, keep, counters, capacity,
                                                                  cand, counters + 1);
  }
  MMB_CHECK_LAUNCH();
  MMB_CHECK_CUDA(cudaMemcpyAsync(status, counters, 4 * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return MMB_OK;
}

extern "C" int mmb_detect_edge_capacity(int capacity) { return 4 * capacity + 4096; }

extern "C" int mmb_detect_chunk(const void* in, int dtype, const int64_t in_strides[3], int Z,
                                int Y, int X, int64_t pitch, double scale,
                                const mmb_preproc_params* pre, int bz, int by, int bx,
                                const double* sigmas, int num_sigma, double threshold,
                                double overlap, int z_lo, int z_hi, void* work, mmb_cand* cand,
                                int capacity, int* n_out, int* n_peaks, void* stream) {
  MMB_REQUIRE(n_out, "null buffer");
  MMB_REQUIRE(work && capacity > 0 && Z > 0 && Y > 0 && pitch > 0, "bad workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const ChunkLayout L = chunk_layout(Z, Y, pitch, capacity);
  // the driver's own counter block doubles as the status block of the synchronous form
  int32_t* status = (int32_t*)((char*)work + L.total - 256) + 8;
  int rc = mmb_detect_chunk_enqueue(in, dtype, in_strides, Z, Y, X, pitch, scale, pre, bz, by, bx,
                                    sigmas, num_sigma, threshold, overlap, z_lo, z_hi, work, cand,
                                    capacity, status, stream);
  if (rc) return rc;
  int32_t host[3] = {0, 0, 0};
  MMB_CHECK_CUDA(cudaMemcpyAsync(host, status, sizeof(host), cudaMemcpyDeviceToHost, st));
  MMB_CHECK_CUDA(cudaStreamSynchronize(st));
  if (n_peaks) *n_peaks = host[0];
  if (host[0] > capacity) {
    set_error("candidate buffer overflow: %d local maxima, capacity %d", host[0], capacity);
    *n_out = host[0];
    return MMB_ERR_OVERFLOW;
  }
  if (host[2] > L.edge_cap) {
    set_error("kill-edge buffer overflow: %d edges, capacity %d", host[2], L.edge_cap);
    *n_out = (host[2] - 4096) / 4 + 1;        // capacity whose edge buffer would fit
    return MMB_ERR_OVERFLOW;
  }
  *n_out = host[1];
  return MMB_OK;
}
