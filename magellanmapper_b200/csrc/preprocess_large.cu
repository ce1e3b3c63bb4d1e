// saturate_roi + denoise_roi for preprocessing blocks of ANY size (global-memory path).
//
// preprocess.cu keeps a whole block in the shared memory of one CTA, which stops at 32
// voxels a side.  Larger blocks - the whole ROI of the GUI path
// (magmap/gui/visualizer.py:2742-2743 calls plot_3d.saturate_roi / denoise_roi on the
// ROI as ONE block), the `lowres` profile (denoise_size 2000: one block per chunk,
// magmap/settings/roi_prof.py:194-), or fine resolutions where ceil(denoise_size /
// resolution) exceeds 32 - run here as a short sequence of kernels over the whole
// volume, every voxel finding its own block from its coordinates, so the number of
// launches does not depend on the number of blocks:
//   1. exact np.percentile (linear) per block: radix select over order-preserving keys,
//      8 bits per pass (1 pass uint8, 2 uint16, 4 float32, 8 float64), float64 _lerp;
//   2. stretch to [0, 1] in float64, block mean (fixed-order reduction), clip;
//   3. sigma = 8 Gaussian, mode 'nearest' at the BLOCK faces, truncate 4 (65 taps),
//      z then y then x as scipy.ndimage.gaussian_filter walks the axes;
//   4. unsharp mask, octahedron(1) erosion when the block mean exceeds the threshold.
// Same arithmetic, operation for operation, as the shared-memory kernel.
#include <math.h>
#include <type_traits>
#include "common.cuh"

namespace mmb {

constexpr int kLgThreads = 256;
constexpr int kLgTargets = 4;          // two percentiles x the two ranks around each
constexpr int kLgPartials = 64;        // CTAs per block in the reductions

struct LgGeom {
  int Z, Y, X;
  int bz, by, bx;
  int nbz, nby, nbx;
  int64_t sz, sy, sx;       // input element strides
  int64_t pitch;            // float volumes: [Z][Y][pitch]
};

struct LgBlock {            // per block, device memory
  unsigned long long prefix[kLgTargets];
  unsigned long long rank[kLgTargets];
  double vmin, vmax, mean;
  int degenerate, pad;
};

__device__ __forceinline__ void lg_block_box(const LgGeom& g, int b, int lo[3], int n[3]) {
  const int bxi = b % g.nbx; b /= g.nbx;
  const int byi = b % g.nby; b /= g.nby;
  lo[0] = b * g.bz; lo[1] = byi * g.by; lo[2] = bxi * g.bx;
  n[0] = min(g.bz, g.Z - lo[0]); n[1] = min(g.by, g.Y - lo[1]); n[2] = min(g.bx, g.X - lo[2]);
}

template <typename T>
__device__ __forceinline__ unsigned long long lg_key(T v) {
  if constexpr (std::is_integral<T>::value) {
    return (unsigned long long)v;
  } else if constexpr (sizeof(T) == 4) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? (unsigned)~u : (u | 0x80000000u);
  } else {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  }
}
template <typename T>
__device__ __forceinline__ double lg_value(unsigned long long k) {
  if constexpr (std::is_integral<T>::value) {
    return (double)k;
  } else if constexpr (sizeof(T) == 4) {
    const unsigned u = (unsigned)k;
    return (double)__uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
  } else {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
  }
}

__device__ __forceinline__ double lg_lerp(double a, double b, double t) {   // numpy _lerp
  const double d = b - a;
  return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

// ranks wanted per block (np.percentile 'linear': floor((n-1) q) and the next one)
__global__ void lg_init_kernel(const __grid_constant__ LgGeom g, int nblocks, double q0, double q1,
                               LgBlock* __restrict__ blk) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  int lo[3], n3[3];
  lg_block_box(g, b, lo, n3);
  const long long n = (long long)n3[0] * n3[1] * n3[2];
  for (int t = 0; t < kLgTargets; ++t) {
    const double virt = (double)(n - 1) * ((t < 2 ? q0 : q1) / 100.0);
    long long k = (long long)floor(virt);
    if (k < 0) k = 0;
    if (k > n - 1) k = n - 1;
    long long want = k + (t & 1);
    if (want > n - 1) want = n - 1;
    blk[b].prefix[t] = 0ull;
    blk[b].rank[t] = (unsigned long long)want;
  }
}

// one radix-select pass: per block and target, histogram of digit `pass` (from the most
// significant byte) over the elements whose higher digits equal the target's prefix.
// grid = (kLgPartials, nblocks)
template <typename T>
__global__ void __launch_bounds__(kLgThreads)
lg_hist_kernel(const T* __restrict__ in, const __grid_constant__ LgGeom g, int shift,
               const LgBlock* __restrict__ blk, unsigned long long* __restrict__ hist) {
  __shared__ unsigned s[kLgTargets][256];
  for (int i = threadIdx.x; i < kLgTargets * 256; i += kLgThreads) (&s[0][0])[i] = 0;
  const int b = blockIdx.y;
  int lo[3], n3[3];
  lg_block_box(g, b, lo, n3);
  unsigned long long pre[kLgTargets];
#pragma unroll
  for (int t = 0; t < kLgTargets; ++t) pre[t] = blk[b].prefix[t];
  __syncthreads();
  const long long n = (long long)n3[0] * n3[1] * n3[2];
  const int nyx = n3[1] * n3[2];
  const bool top = shift + 8 >= 8 * (int)sizeof(T) || shift + 8 >= 64;
  for (long long i = (long long)blockIdx.x * kLgThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kLgThreads) {
    const int z = (int)(i / nyx), r = (int)(i - (long long)z * nyx), y = r / n3[2],
              x = r - y * n3[2];
    const unsigned long long k = lg_key<T>(in[(int64_t)(lo[0] + z) * g.sz +
                                              (int64_t)(lo[1] + y) * g.sy +
                                              (int64_t)(lo[2] + x) * g.sx]);
    const unsigned d = (unsigned)(k >> shift) & 255u;
    const unsigned long long hi = top ? 0ull : (k >> (shift + 8));
#pragma unroll
    for (int t = 0; t < kLgTargets; ++t)
      if (top || hi == (pre[t] >> (shift + 8))) atomicAdd(&s[t][d], 1u);
  }
  __syncthreads();
  for (int t = 0; t < kLgTargets; ++t)
    if (s[t][threadIdx.x])
      atomicAdd(&hist[((int64_t)b * kLgTargets + t) * 256 + threadIdx.x],
                (unsigned long long)s[t][threadIdx.x]);
}

// locate each target's rank in its histogram, extend its prefix, clear the histogram
__global__ void lg_pick_kernel(int nblocks, int shift, LgBlock* __restrict__ blk,
                               unsigned long long* __restrict__ hist) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblocks * kLgTargets) return;
  const int b = idx / kLgTargets, t = idx - b * kLgTargets;
  unsigned long long* h = hist + (int64_t)idx * 256;
  const unsigned long long want = blk[b].rank[t];
  unsigned long long acc = 0;
  int bin = 0;
  for (; bin < 255; ++bin) {
    if (acc + h[bin] > want) break;
    acc += h[bin];
  }
  blk[b].prefix[t] |= (unsigned long long)bin << shift;
  blk[b].rank[t] = want - acc;
  for (int k = 0; k < 256; ++k) h[k] = 0ull;
}

template <typename T>
__global__ void lg_bounds_kernel(const __grid_constant__ LgGeom g, int nblocks, double q0,
                                 double q1, double max_thresh, int force_degenerate,
                                 LgBlock* __restrict__ blk) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  int lo[3], n3[3];
  lg_block_box(g, b, lo, n3);
  const long long n = (long long)n3[0] * n3[1] * n3[2];
  double v[2];
  for (int w = 0; w < 2; ++w) {
    const double virt = (double)(n - 1) * ((w == 0 ? q0 : q1) / 100.0);
    double gamma = virt - floor(virt);
    if (virt >= (double)(n - 1) || virt < 0.0) gamma = 0.0;
    v[w] = lg_lerp(lg_value<T>(blk[b].prefix[2 * w]), lg_value<T>(blk[b].prefix[2 * w + 1]), gamma);
  }
  double vmin = v[0], vmax = v[1];
  const int degenerate = force_degenerate || vmin == vmax;
  if (!degenerate && vmax < max_thresh) vmax = max_thresh;
  blk[b].vmin = vmin; blk[b].vmax = vmax; blk[b].degenerate = degenerate;
}

// stretch + clip -> den (float), partial sums of the stretched values per block.
// grid = (kLgPartials, nblocks)
template <typename T>
__global__ void __launch_bounds__(kLgThreads)
lg_stretch_kernel(const T* __restrict__ in, const __grid_constant__ LgGeom g, double clip_min,
                  double clip_max, const LgBlock* __restrict__ blk, float* __restrict__ den,
                  double* __restrict__ partial) {
  __shared__ double s_red[kLgThreads / 32];
  const int b = blockIdx.y;
  int lo[3], n3[3];
  lg_block_box(g, b, lo, n3);
  const double vmin = blk[b].vmin, vmax = blk[b].vmax;
  const bool degenerate = blk[b].degenerate != 0;
  const double den_d = 1.0 / (vmax - vmin);      // as the shared-memory kernel: one reciprocal
  const long long n = (long long)n3[0] * n3[1] * n3[2];
  const int nyx = n3[1] * n3[2];
  double psum = 0.0;
  for (long long i = (long long)blockIdx.x * kLgThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kLgThreads) {
    const int z = (int)(i / nyx), r = (int)(i - (long long)z * nyx), y = r / n3[2],
              x = r - y * n3[2];
    double s = (double)in[(int64_t)(lo[0] + z) * g.sz + (int64_t)(lo[1] + y) * g.sy +
                          (int64_t)(lo[2] + x) * g.sx];
    if (!degenerate) {
      s = fmin(fmax(s, vmin), vmax);
      s = (s - vmin) * den_d;
    }
    psum += s;
    den[((int64_t)(lo[0] + z) * g.Y + (lo[1] + y)) * g.pitch + (lo[2] + x)] =
        (float)fmin(fmax(s, clip_min), clip_max);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = psum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kLgThreads / 32; ++w) t += s_red[w];
    partial[(int64_t)b * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void lg_mean_kernel(const __grid_constant__ LgGeom g, int nblocks, int nparts,
                               const double* __restrict__ partial, LgBlock* __restrict__ blk) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  int lo[3], n3[3];
  lg_block_box(g, b, lo, n3);
  double t = 0.0;
  for (int k = 0; k < nparts; ++k) t += partial[(int64_t)b * nparts + k];
  blk[b].mean = t / (double)((long long)n3[0] * n3[1] * n3[2]);
}

struct LgWeights { float w[33]; int r; };     // sigma = 8: radius 32

// 1-D Gaussian along `axis`, 'nearest' at the faces of every block.  One thread per voxel.
__global__ void __launch_bounds__(kLgThreads)
lg_blur_kernel(const float* __restrict__ in, float* __restrict__ out,
               const __grid_constant__ LgGeom g, int axis, const __grid_constant__ LgWeights w) {
  const int x = blockIdx.x * kLgThreads + threadIdx.x;
  const int y = blockIdx.y, z = blockIdx.z;
  if (x >= g.X) return;
  const int c = axis == 0 ? z : (axis == 1 ? y : x);
  const int bsz = axis == 0 ? g.bz : (axis == 1 ? g.by : g.bx);
  const int n = axis == 0 ? g.Z : (axis == 1 ? g.Y : g.X);
  const int lo = c / bsz * bsz, hi = min(lo + bsz, n) - 1;
  const int64_t stride = axis == 0 ? (int64_t)g.Y * g.pitch : (axis == 1 ? g.pitch : 1);
  const int64_t here = ((int64_t)z * g.Y + y) * g.pitch + x;
  const float* base = in + (here - (int64_t)c * stride);
  // scipy.ndimage.correlate1d accumulates from the lowest tap upwards
  float acc = 0.f;
  for (int t = -w.r; t <= w.r; ++t) {
    const int q = min(max(c + t, lo), hi);
    acc = fmaf(w.w[t < 0 ? -t : t], __ldg(base + (int64_t)q * stride), acc);
  }
  out[here] = acc;
}

// unsharp mask: den + (den - strength * blurred), in place over `blur`
__global__ void __launch_bounds__(kLgThreads)
lg_unsharp_kernel(const float* __restrict__ den, float* __restrict__ blur,
                  const __grid_constant__ LgGeom g, float strength) {
  const int x = blockIdx.x * kLgThreads + threadIdx.x;
  if (x >= g.X) return;
  const int64_t here = ((int64_t)blockIdx.z * g.Y + blockIdx.y) * g.pitch + x;
  const float d = den[here];
  const float hp = d - strength * blur[here];
  blur[here] = d + hp;
}

// octahedron(1) erosion inside every block whose mean exceeds the threshold; else a copy
__global__ void __launch_bounds__(kLgThreads)
lg_erode_kernel(const float* __restrict__ in, float* __restrict__ out,
                const __grid_constant__ LgGeom g, double threshold,
                const LgBlock* __restrict__ blk) {
  const int x = blockIdx.x * kLgThreads + threadIdx.x;
  const int y = blockIdx.y, z = blockIdx.z;
  if (x >= g.X) return;
  const int bzi = z / g.bz, byi = y / g.by, bxi = x / g.bx;
  const int b = (bzi * g.nby + byi) * g.nbx + bxi;
  const int64_t here = ((int64_t)z * g.Y + y) * g.pitch + x;
  float v = in[here];
  if (threshold != 0.0 && blk[b].mean > threshold) {
    const int64_t plane = (int64_t)g.Y * g.pitch;
    const int z0 = bzi * g.bz, z1 = min(z0 + g.bz, g.Z) - 1;
    const int y0 = byi * g.by, y1 = min(y0 + g.by, g.Y) - 1;
    const int x0 = bxi * g.bx, x1 = min(x0 + g.bx, g.X) - 1;
    if (z > z0) v = fminf(v, in[here - plane]);
    if (z < z1) v = fminf(v, in[here + plane]);
    if (y > y0) v = fminf(v, in[here - g.pitch]);
    if (y < y1) v = fminf(v, in[here + g.pitch]);
    if (x > x0) v = fminf(v, in[here - 1]);
    if (x < x1) v = fminf(v, in[here + 1]);
  }
  out[here] = v;
}

static inline int64_t al(int64_t x) { return (x + 255) / 256 * 256; }

struct LgLayout { int64_t vol, blk, hist, partial, total; };
static LgLayout lg_layout(int Z, int Y, int64_t pitch, int64_t nblocks) {
  LgLayout L;
  L.vol = al((int64_t)Z * Y * pitch * 4);
  L.blk = al(nblocks * (int64_t)sizeof(LgBlock));
  L.hist = al(nblocks * kLgTargets * 256 * 8);
  L.partial = al(nblocks * kLgPartials * 8);
  L.total = 2 * L.vol + L.blk + L.hist + L.partial;
  return L;
}

int64_t preprocess_large_work_bytes(int Z, int Y, int64_t pitch, int bz, int by, int bx) {
  bz = bz < Z ? bz : Z; by = by < Y ? by : Y; bx = bx < (int)pitch ? bx : (int)pitch;
  if (bz < 1 || by < 1 || bx < 1) return 0;
  return lg_layout(Z, Y, pitch, cdiv(Z, bz) * cdiv(Y, by) * cdiv(pitch, bx)).total;
}

template <typename T>
static int run_large(const T* in, const LgGeom& g, const mmb_preproc_params& p, float* out,
                     char* work, cudaStream_t st) {
  const int64_t nblocks64 = (int64_t)g.nbz * g.nby * g.nbx;
  if (nblocks64 > 65535) {
    set_error("%lld preprocessing blocks above 32 voxels: more than 65535", (long long)nblocks64);
    return MMB_ERR_UNSUPPORTED;
  }
  const int nblocks = (int)nblocks64;
  const LgLayout L = lg_layout(g.Z, g.Y, g.pitch, nblocks);
  float* den = (float*)work;
  float* tmp = (float*)(work + L.vol);
  LgBlock* blk = (LgBlock*)(work + 2 * L.vol);
  unsigned long long* hist = (unsigned long long*)(work + 2 * L.vol + L.blk);
  double* partial = (double*)(work + 2 * L.vol + L.blk + L.hist);
  const unsigned nb = (unsigned)cdiv(nblocks, 128);
  ProfScope ps(PROF_PREPROCESS, (double)g.Z * g.Y * g.X, st);
  // equal percentiles select the same sample twice: vmin == vmax whatever the data,
  // i.e. the stretch is skipped (denoise_roi on its own arrives like this)
  const int force_degenerate = p.clip_vmin == p.clip_vmax;
  dim3 bgrid(kLgPartials, (unsigned)nblocks);
  if (!force_degenerate) {
    MMB_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)L.hist, st));
    lg_init_kernel<<<nb, 128, 0, st>>>(g, nblocks, p.clip_vmin, p.clip_vmax, blk);
    MMB_CHECK_LAUNCH();
    for (int shift = 8 * (int)sizeof(T) - 8; shift >= 0; shift -= 8) {
      lg_hist_kernel<T><<<bgrid, kLgThreads, 0, st>>>(in, g, shift, blk, hist);
      MMB_CHECK_LAUNCH();
      lg_pick_kernel<<<(unsigned)cdiv((int64_t)nblocks * kLgTargets, 128), 128, 0, st>>>(
          nblocks, shift, blk, hist);
      MMB_CHECK_LAUNCH();
    }
  }
  lg_bounds_kernel<T><<<nb, 128, 0, st>>>(g, nblocks, p.clip_vmin, p.clip_vmax, p.max_thresh,
                                          force_degenerate, blk);
  MMB_CHECK_LAUNCH();
  lg_stretch_kernel<T><<<bgrid, kLgThreads, 0, st>>>(in, g, p.clip_min, p.clip_max, blk, den,
                                                     partial);
  MMB_CHECK_LAUNCH();
  lg_mean_kernel<<<nb, 128, 0, st>>>(g, nblocks, kLgPartials, partial, blk);
  MMB_CHECK_LAUNCH();
  dim3 vgrid((unsigned)cdiv(g.X, kLgThreads), (unsigned)g.Y, (unsigned)g.Z);
  const float* cur = den;
  if (p.unsharp_strength != 0.0) {
    LgWeights w;
    const double sigma = 8.0;
    w.r = (int)(4.0 * sigma + 0.5);
    double sum = 0.0;
    for (int t = -w.r; t <= w.r; ++t) sum += exp(-0.5 / (sigma * sigma) * (double)(t * t));
    for (int t = 0; t <= w.r; ++t)
      w.w[t] = (float)(exp(-0.5 / (sigma * sigma) * (double)(t * t)) / sum);
    // z: den -> tmp, y: tmp -> out, x: out -> tmp; unsharp in place over tmp
    lg_blur_kernel<<<vgrid, kLgThreads, 0, st>>>(den, tmp, g, 0, w);
    MMB_CHECK_LAUNCH();
    lg_blur_kernel<<<vgrid, kLgThreads, 0, st>>>(tmp, out, g, 1, w);
    MMB_CHECK_LAUNCH();
    lg_blur_kernel<<<vgrid, kLgThreads, 0, st>>>(out, tmp, g, 2, w);
    MMB_CHECK_LAUNCH();
    lg_unsharp_kernel<<<vgrid, kLgThreads, 0, st>>>(den, tmp, g, (float)p.unsharp_strength);
    MMB_CHECK_LAUNCH();
    cur = tmp;
  }
  lg_erode_kernel<<<vgrid, kLgThreads, 0, st>>>(cur, out, g, p.erosion_threshold, blk);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

// `scratch` (may be NULL): preprocess_large_work_bytes() bytes the caller lends; otherwise
// the workspace is taken from the stream-ordered allocator for the duration of the call
int preprocess_large_impl(const void* in, int dtype, const int64_t strides[3], int Z, int Y,
                          int X, int bz, int by, int bx, const mmb_preproc_params* p, float* out,
                          int64_t pitch, void* scratch, cudaStream_t st) {
  MMB_REQUIRE(Y <= 65535 && Z <= 65535, "Y and Z must be <= 65535");
  LgGeom g;
  g.Z = Z; g.Y = Y; g.X = X; g.bz = bz; g.by = by; g.bx = bx;
  g.nbz = (int)cdiv(Z, bz); g.nby = (int)cdiv(Y, by); g.nbx = (int)cdiv(X, bx);
  g.sz = strides[0]; g.sy = strides[1]; g.sx = strides[2]; g.pitch = pitch;
  const int64_t nblocks = (int64_t)g.nbz * g.nby * g.nbx;
  const LgLayout L = lg_layout(Z, Y, pitch, nblocks);
  char* work = (char*)scratch;
  if (!work) MMB_CHECK_CUDA(cudaMallocAsync((void**)&work, (size_t)L.total, st));
  int rc;
  switch (dtype) {
    case MMB_U8:  rc = run_large<uint8_t>((const uint8_t*)in, g, *p, out, work, st); break;
    case MMB_U16: rc = run_large<uint16_t>((const uint16_t*)in, g, *p, out, work, st); break;
    case MMB_F32: rc = run_large<float>((const float*)in, g, *p, out, work, st); break;
    case MMB_F64: rc = run_large<double>((const double*)in, g, *p, out, work, st); break;
    default: set_error("unknown dtype %d", dtype); rc = MMB_ERR_INVALID;
  }
  if (!scratch) MMB_CHECK_CUDA(cudaFreeAsync(work, st));
  return rc;
}

}  // namespace mmb
