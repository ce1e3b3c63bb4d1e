// Exact percentiles of integer images, per z-plane or over a whole volume.
//
// Reference: the import metadata `near_min` / `near_max` that saturate_roi reads as
// config.near_max (magmap/plot/plot_3d.py:97-100) are, per channel, the minimum over
// z-planes of the plane's 0.5th percentile and the maximum of its 99.5th
// (magmap/io/importer.py:1368-1377, 571-583: np.percentile per plane, then
// calc_near_intensity_bounds :1447-1468); calc_intensity_bounds (:1415-1444) is the same
// percentile pair over a whole array.  np.percentile's default 'linear' method needs
// the two order statistics around the virtual index (n-1)q/100 and numpy's _lerp.
//
// uint8 / uint16 values are found exactly with two 256-bin histogram passes (high
// byte, then the low byte inside the bins that hold the wanted ranks), so an image is
// read twice (4 B/voxel for uint16) whatever the number of percentiles, with
// shared-memory histograms merged into global ones by 64-bit atomics.
#include "common.cuh"

namespace mmb {

constexpr int kPctThreads = 256;
constexpr int kPctMaxQ = 4;                    // percentiles per call
constexpr int kPctTargets = 2 * kPctMaxQ;      // order statistics per group

__device__ __forceinline__ double np_lerp_f64(double a, double b, double t) {
  const double d = b - a;                      // numpy _lerp (lib/_function_base_impl.py)
  return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

struct PctGeom {
  int64_t sz, sy, sx;       // element strides of the (Z, Y, X) view
  int Z, Y, X;
  int per_plane;            // 1: one group per z-plane, 0: the whole volume is one group
  int nq;
  double q[kPctMaxQ];       // percentiles, 0..100
};

template <typename T>
__device__ __forceinline__ unsigned load_key(const T* in, const PctGeom& g, int z, int64_t i) {
  const int y = (int)(i / g.X), x = (int)(i - (int64_t)y * g.X);
  return (unsigned)in[(int64_t)z * g.sz + (int64_t)y * g.sy + (int64_t)x * g.sx];
}

// pass 1: histogram of the high byte (the value itself for uint8).  grid = (chunks, Z)
template <typename T>
__global__ void __launch_bounds__(kPctThreads)
pct_hist_hi_kernel(const T* __restrict__ in, const __grid_constant__ PctGeom g,
                   unsigned long long* __restrict__ hist) {
  __shared__ unsigned s[256];
  s[threadIdx.x] = 0;
  __syncthreads();
  const int z = blockIdx.y;
  const int64_t n = (int64_t)g.Y * g.X;
  const int shift = sizeof(T) == 1 ? 0 : 8;
  for (int64_t i = (int64_t)blockIdx.x * kPctThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kPctThreads)
    atomicAdd(&s[load_key(in, g, z, i) >> shift], 1u);
  __syncthreads();
  const int grp = g.per_plane ? z : 0;
  if (s[threadIdx.x]) atomicAdd(&hist[(int64_t)grp * 256 + threadIdx.x], (unsigned long long)s[threadIdx.x]);
}

// ranks wanted per group: k = floor((n-1) q / 100) and k + 1 (clamped), located in the
// high-byte histogram: target t -> (bin, rank inside the bin).  One thread per target.
__global__ void pct_select_kernel(const __grid_constant__ PctGeom g, int n_groups,
                                  const unsigned long long* __restrict__ hist,
                                  int* __restrict__ tgt_bin,
                                  unsigned long long* __restrict__ tgt_rank) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_groups * 2 * g.nq) return;
  const int grp = idx / (2 * g.nq), t = idx - grp * 2 * g.nq;
  const long long n = (long long)g.Y * g.X * (g.per_plane ? 1 : g.Z);
  const double virt = (double)(n - 1) * (g.q[t >> 1] / 100.0);
  long long k = (long long)floor(virt);
  if (k < 0) k = 0;
  if (k > n - 1) k = n - 1;
  long long want = k + (t & 1);
  if (want > n - 1) want = n - 1;
  const unsigned long long* h = hist + (int64_t)grp * 256;
  unsigned long long acc = 0;
  int b = 0;
  for (; b < 255; ++b) {
    if (acc + h[b] > (unsigned long long)want) break;
    acc += h[b];
  }
  tgt_bin[(int64_t)grp * kPctTargets + t] = b;
  tgt_rank[(int64_t)grp * kPctTargets + t] = (unsigned long long)want - acc;
}

// pass 2 (uint16): low-byte histograms inside the target bins.  grid = (chunks, Z)
__global__ void __launch_bounds__(kPctThreads)
pct_hist_lo_kernel(const uint16_t* __restrict__ in, const __grid_constant__ PctGeom g,
                   const int* __restrict__ tgt_bin, unsigned long long* __restrict__ hist2) {
  __shared__ unsigned s[kPctTargets][256];
  __shared__ int bins[kPctTargets];
  const int z = blockIdx.y;
  const int grp = g.per_plane ? z : 0;
  const int nt = 2 * g.nq;
  for (int i = threadIdx.x; i < kPctTargets * 256; i += kPctThreads) (&s[0][0])[i] = 0;
  if (threadIdx.x < kPctTargets)
    bins[threadIdx.x] = threadIdx.x < nt ? tgt_bin[(int64_t)grp * kPctTargets + threadIdx.x] : -1;
  __syncthreads();
  const int64_t n = (int64_t)g.Y * g.X;
  for (int64_t i = (int64_t)blockIdx.x * kPctThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kPctThreads) {
    const unsigned v = load_key(in, g, z, i);
    const int hi = (int)(v >> 8);
#pragma unroll
    for (int t = 0; t < kPctTargets; ++t)
      if (hi == bins[t]) atomicAdd(&s[t][v & 255u], 1u);
  }
  __syncthreads();
  for (int t = 0; t < nt; ++t)
    if (s[t][threadIdx.x])
      atomicAdd(&hist2[((int64_t)grp * kPctTargets + t) * 256 + threadIdx.x],
                (unsigned long long)s[t][threadIdx.x]);
}

// order statistics -> np.percentile(..., method='linear').  One thread per (group, q).
__global__ void pct_finish_kernel(const __grid_constant__ PctGeom g, int n_groups, int is_u16,
                                  const int* __restrict__ tgt_bin,
                                  const unsigned long long* __restrict__ tgt_rank,
                                  const unsigned long long* __restrict__ hist2,
                                  double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_groups * g.nq) return;
  const int grp = idx / g.nq, qi = idx - grp * g.nq;
  double v[2];
  for (int w = 0; w < 2; ++w) {
    const int t = 2 * qi + w;
    const int hi = tgt_bin[(int64_t)grp * kPctTargets + t];
    if (!is_u16) { v[w] = (double)hi; continue; }
    const unsigned long long want = tgt_rank[(int64_t)grp * kPctTargets + t];
    const unsigned long long* h = hist2 + ((int64_t)grp * kPctTargets + t) * 256;
    unsigned long long acc = 0;
    int b = 0;
    for (; b < 255; ++b) {
      if (acc + h[b] > want) break;
      acc += h[b];
    }
    v[w] = (double)(hi * 256 + b);
  }
  const long long n = (long long)g.Y * g.X * (g.per_plane ? 1 : g.Z);
  const double virt = (double)(n - 1) * (g.q[qi] / 100.0);
  double k = floor(virt);
  double gamma = virt - k;
  if (virt >= (double)(n - 1) || virt < 0.0) gamma = 0.0;
  out[idx] = np_lerp_f64(v[0], v[1], gamma);
}

static inline int64_t align256p(int64_t x) { return (x + 255) / 256 * 256; }

}  // namespace mmb

using namespace mmb;

extern "C" int64_t mmb_percentiles_work_bytes(int n_groups) {
  const int64_t g = n_groups > 0 ? n_groups : 1;
  return align256p(g * 256 * 8) + align256p(g * kPctTargets * 4) + align256p(g * kPctTargets * 8) +
         align256p(g * kPctTargets * 256 * 8);
}

extern "C" int mmb_percentiles(const void* in, int dtype, const int64_t in_strides[3], int Z, int Y,
                               int X, int per_plane, const double* q_percent, int nq,
                               double* out_device, void* work, void* stream) {
  MMB_REQUIRE(in && in_strides && q_percent && out_device && work, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && Z <= 65535, "bad shape");
  MMB_REQUIRE(nq >= 1 && nq <= kPctMaxQ, "1..4 percentiles per call");
  if (dtype != MMB_U8 && dtype != MMB_U16) {
    set_error("mmb_percentiles: only uint8 and uint16 images (dtype %d given)", dtype);
    return MMB_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < nq; ++i) MMB_REQUIRE(q_percent[i] >= 0.0 && q_percent[i] <= 100.0, "percentile outside 0..100");
  cudaStream_t st = (cudaStream_t)stream;
  PctGeom g;
  g.sz = in_strides[0]; g.sy = in_strides[1]; g.sx = in_strides[2];
  g.Z = Z; g.Y = Y; g.X = X; g.per_plane = per_plane ? 1 : 0; g.nq = nq;
  for (int i = 0; i < kPctMaxQ; ++i) g.q[i] = i < nq ? q_percent[i] : 0.0;
  const int n_groups = per_plane ? Z : 1;
  char* base = (char*)work;
  unsigned long long* hist = (unsigned long long*)base;      base += align256p((int64_t)n_groups * 256 * 8);
  int* tgt_bin = (int*)base;                                  base += align256p((int64_t)n_groups * kPctTargets * 4);
  unsigned long long* tgt_rank = (unsigned long long*)base;  base += align256p((int64_t)n_groups * kPctTargets * 8);
  unsigned long long* hist2 = (unsigned long long*)base;
  MMB_CHECK_CUDA(cudaMemsetAsync(work, 0, (size_t)mmb_percentiles_work_bytes(n_groups), st));
  const int64_t n_plane = (int64_t)Y * X;
  // enough CTAs to fill the GPU, few enough that the global atomics stay negligible
  int chunks = (int)cdiv(n_plane, (int64_t)kPctThreads * 64);
  const int max_chunks = (int)cdiv(8 * (int64_t)num_sms(), Z);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)Z);
  if (dtype == MMB_U8)
    pct_hist_hi_kernel<uint8_t><<<grid, kPctThreads, 0, st>>>((const uint8_t*)in, g, hist);
  else
    pct_hist_hi_kernel<uint16_t><<<grid, kPctThreads, 0, st>>>((const uint16_t*)in, g, hist);
  MMB_CHECK_LAUNCH();
  const int n_t = n_groups * 2 * nq;
  pct_select_kernel<<<(unsigned)cdiv(n_t, 128), 128, 0, st>>>(g, n_groups, hist, tgt_bin, tgt_rank);
  MMB_CHECK_LAUNCH();
  if (dtype == MMB_U16) {
    pct_hist_lo_kernel<<<grid, kPctThreads, 0, st>>>((const uint16_t*)in, g, tgt_bin, hist2);
    MMB_CHECK_LAUNCH();
  }
  pct_finish_kernel<<<(unsigned)cdiv(n_groups * nq, 128), 128, 0, st>>>(
      g, n_groups, dtype == MMB_U16 ? 1 : 0, tgt_bin, tgt_rank, hist2, out_device);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}
