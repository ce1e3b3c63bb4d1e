// saturate_roi + denoise_roi on every preprocessing block, one CTA per block.
//
// Reference: magmap/cv/stack_detect.py:122-150 splits a chunk into
// denoise_max_shape blocks (25^3 at 1 um/px) and runs, per block,
//   plot_3d.saturate_roi (plot_3d.py:55-111): vmin,vmax = np.percentile(blk,
//     (clip_vmin, clip_vmax)); unchanged if equal; vmax = max(vmax, near_max *
//     max_thresh_factor); clip and stretch to [0,1];
//   plot_3d.denoise_roi (plot_3d.py:114-172): mean; clip to [clip_min, clip_max];
//     unsharp: x + (x - strength * gaussian(x, sigma=8, mode='nearest'));
//     octahedron(1) erosion when mean > erosion_threshold.
// The block lives in shared memory for the whole pipeline: exact order
// statistics by bisection over the key bits with the keys held in registers (both
// percentiles in the same passes), float64 for the percentile interpolation, the
// stretch and the block mean (they feed equality / threshold decisions), then the
// sigma=8 blur as three in-place matrix sweeps (radius 32 exceeds the block, so
// with 'nearest' padding each 1-D pass is a dense n x n matrix built on the
// host), unsharp, erosion, and one write of float32.
#include <math.h>
#include <stdlib.h>
#include <type_traits>
#include <vector>
#include <map>
#include <mutex>
#include "common.cuh"

namespace mmb {

int64_t preprocess_large_work_bytes(int Z, int Y, int64_t pitch, int bz, int by, int bx);

// CTA shapes: cubic blocks up to NL^3 run 640 threads (25 voxels per thread at 25^3); thin
// blocks - anisotropic volumes: 5 x 25 x 25 at 5 x 1 x 1 um - get their own (NZ, NL)
// instantiation with a CTA sized for ~20 voxels per thread, so that the per-thread loops
// and the line sweeps of the blur keep every thread busy
constexpr int kPreCubicThreads = 640;

struct PreGeom {
  int Z, Y, X;
  int bz, by, bx;
  int nbz, nby, nbx;
  int64_t sz, sy, sx;   // input element strides
  int64_t pitch;
};

__device__ __forceinline__ unsigned key_of(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_of(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// numpy _lerp (lib/_function_base_impl.py) in float64
__device__ __forceinline__ double np_lerp(double a, double b, double t) {
  const double d = b - a;
  return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

// FULL: the block is NZ x NL x NL, so every index decomposition divides by a
// compile-time constant; partial blocks at the chunk faces take the generic body.
template <typename T, int NZ, int NL, bool FULL, int kPreThreads>
__device__ __forceinline__ void preprocess_body(const T* __restrict__ in, const PreGeom& g,
                                                const mmb_preproc_params& p,
                                                const float* __restrict__ mats, int mat_pitch,
                                                float* __restrict__ out, int z0, int y0, int x0,
                                                int nz_rt, int ny_rt, int nx_rt) {
  constexpr int NVOX = NZ * NL * NL;
  constexpr int NPT = (NVOX + kPreThreads - 1) / kPreThreads;
  constexpr int NVOXP = (NVOX + 3) / 4 * 4;      // keeps M 16-byte aligned for float4 rows
  constexpr int NLP = (NL + 3) / 4 * 4;          // matrix row pitch (float4 rows)
  constexpr int NZP = (NZ + 3) / 4 * 4;
  // blur matrices live in shared memory with rows interleaved in pairs - entry (i / 2, j)
  // holds (m[i][j], m[i + 1][j]) - so that two outputs of a line share one packed FFMA2
  constexpr int MZ = 2 * ((NZ + 1) / 2) * NZP, MYX = 2 * ((NL + 1) / 2) * NLP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* vals = reinterpret_cast<float*>(smem_raw);                 // NVOX
  float* M = vals + NVOXP;                                           // MZ + 2 * MYX
  int* hist = reinterpret_cast<int*>(M + MZ + 2 * MYX);              // 2 * 256
  __shared__ unsigned s_prefix[2];
  __shared__ int s_k[2];
  __shared__ int s_cnt_le[2];
  __shared__ unsigned s_next[2];
  __shared__ double s_red[kPreThreads / 32];
  __shared__ double s_vmin, s_vmax, s_mean;
  __shared__ int s_degenerate;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int nz = FULL ? NZ : nz_rt, ny = FULL ? NL : ny_rt, nx = FULL ? NL : nx_rt;
  const int n = nz * ny * nx;
  const int nyx = ny * nx;

  // ---- load block, blur matrices --------------------------------------------
  for (int i = tid; i < n; i += kPreThreads) {
    const int z = i / nyx, r = i - z * nyx, y = r / nx, x = r - y * nx;
    vals[i] = (float)in[(int64_t)(z0 + z) * g.sz + (int64_t)(y0 + y) * g.sy +
                        (int64_t)(x0 + x) * g.sx];
  }
  if (p.unsharp_strength != 0.0) {
    const int lens[3] = {nz, ny, nx};
    for (int a = 0; a < 3; ++a) {
      const float* src = mats + (int64_t)(lens[a] - 1) * mat_pitch;   // matrix for length len
      const int mp = a == 0 ? NZP : NLP;
      float* Ma = M + (a == 0 ? 0 : MZ + (a - 1) * MYX);
      for (int i = tid; i < (a == 0 ? MZ : MYX); i += kPreThreads) {
        const int h = i & 1, e = i >> 1, pr = e / mp, c = e - pr * mp;
        const int r = 2 * pr + h;
        Ma[i] = (r < lens[a] && c < lens[a]) ? src[r * lens[a] + c] : 0.f;
      }
    }
  }
  if (tid < 2) {
    // np.percentile 'linear': virtual index (n-1)*q, q = pct/100 in float64
    const double q = (tid == 0 ? p.clip_vmin : p.clip_vmax) / 100.0;
    double virt = (double)(n - 1) * q;
    int k = (int)floor(virt);
    if (virt >= (double)(n - 1)) k = n - 1;
    if (k < 0) k = 0;
    s_k[tid] = k;
    s_prefix[tid] = 0u;
  }
  __syncthreads();

  // ---- exact order statistics of ranks s_k[0], s_k[1] ---------------------------
  // Bisection over the key bits, most significant first: K grows to the largest
  // value with count(key < K) <= rank, which is the element of that rank.  Every
  // thread keeps the keys of its voxels in registers, so a pass is NPT compares
  // per percentile, one warp reduction and one barrier - no shared-memory
  // histogram, no atomics.  Integer inputs need only as many passes as they have
  // bits (16 for uint16); floats use the order-preserving 32-bit key.
  constexpr bool INTKEY = std::is_integral<T>::value;
  constexpr int NBITS = INTKEY ? 8 * (int)sizeof(T) : 32;
  constexpr int NW = kPreThreads / 32;
  unsigned keys[NPT];
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = tid + k * kPreThreads;
    // padding never counts as "< K" (integer keys are below 2^16: the count below takes
    // the sign bit of key - K)
    keys[k] = INTKEY ? 0x7fffffffu : 0xffffffffu;
    if (i < n) keys[k] = INTKEY ? (unsigned)vals[i] : key_of(vals[i]);
  }
  unsigned K0 = 0u, K1 = 0u;
  {
    const int k0 = s_k[0], k1 = s_k[1];
#pragma unroll 1
    for (int bit = NBITS - 1; bit >= 0; --bit) {
      const unsigned T0 = K0 | (1u << bit), T1 = K1 | (1u << bit);
      int c0 = 0, c1 = 0;
#pragma unroll
      for (int k = 0; k < NPT; ++k) {
        if (INTKEY) {        // keys and thresholds < 2^31: the borrow is the sign of the difference
          c0 += (int)((keys[k] - T0) >> 31);
          c1 += (int)((keys[k] - T1) >> 31);
        } else {
          c0 += keys[k] < T0 ? 1 : 0;
          c1 += keys[k] < T1 ? 1 : 0;
        }
      }
      c0 = __reduce_add_sync(0xffffffffu, c0);
      c1 = __reduce_add_sync(0xffffffffu, c1);
      int* slot = hist + (bit & 1) * 2 * NW;   // double-buffered: one barrier per pass
      if (lane == 0) { slot[warp] = c0; slot[NW + warp] = c1; }
      __syncthreads();
      int t0 = lane < NW ? slot[lane] : 0;
      int t1 = lane < NW ? slot[NW + lane] : 0;
      t0 = __reduce_add_sync(0xffffffffu, t0);
      t1 = __reduce_add_sync(0xffffffffu, t1);
      if (t0 <= k0) K0 = T0;
      if (t1 <= k1) K1 = T1;
    }
  }
  // successor of each selected value: count(key <= K) and min(key > K)
  {
    if (tid < 2) { s_cnt_le[tid] = 0; s_next[tid] = 0xffffffffu; s_prefix[tid] = tid == 0 ? K0 : K1; }
    __syncthreads();
    int c0 = 0, c1 = 0;
    unsigned m0 = 0xffffffffu, m1 = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      if (tid + k * kPreThreads < n) {
        const unsigned key = keys[k];
        if (key <= K0) ++c0; else m0 = min(m0, key);
        if (key <= K1) ++c1; else m1 = min(m1, key);
      }
    }
    c0 = __reduce_add_sync(0xffffffffu, c0);
    c1 = __reduce_add_sync(0xffffffffu, c1);
    m0 = __reduce_min_sync(0xffffffffu, m0);
    m1 = __reduce_min_sync(0xffffffffu, m1);
    if (lane == 0) {
      atomicAdd(&s_cnt_le[0], c0); atomicAdd(&s_cnt_le[1], c1);
      atomicMin(&s_next[0], m0);   atomicMin(&s_next[1], m1);
    }
    __syncthreads();
  }
  if (tid == 0) {
    double v[2];
    for (int w = 0; w < 2; ++w) {
      const double q = (w == 0 ? p.clip_vmin : p.clip_vmax) / 100.0;
      const double virt = (double)(n - 1) * q;
      double prev = floor(virt);
      const double a = INTKEY ? (double)s_prefix[w] : (double)float_of(s_prefix[w]);
      double bb = a;
      double gamma = virt - prev;
      if (virt >= (double)(n - 1) || virt < 0.0) {
        gamma = 0.0;                       // both neighbours are the same sample
      } else {
        const int k = (int)prev;           // rank of a; rank k+1 is a again if duplicated
        if (k + 1 >= s_cnt_le[w])
          bb = INTKEY ? (double)s_next[w] : (double)float_of(s_next[w]);
      }
      v[w] = np_lerp(a, bb, gamma);
    }
    double vmin = v[0], vmax = v[1];
    const int degenerate = vmin == vmax;
    if (!degenerate && vmax < p.max_thresh) vmax = p.max_thresh;
    s_vmin = vmin; s_vmax = vmax; s_degenerate = degenerate;
  }
  __syncthreads();

  // ---- stretch, block mean, clip --------------------------------------------
  const double vmin = s_vmin, vmax = s_vmax;
  const bool degenerate = s_degenerate != 0;
  // one reciprocal per block instead of a float64 division per voxel (about 60 instructions
  // each): at most one unit in the last place of the float64 quotient, far inside the float32
  // rounding of the result
  const double inv_den = 1.0 / (vmax - vmin);
  float den[NPT];
  double psum = 0.0;
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = tid + k * kPreThreads;
    den[k] = 0.f;
    if (i < n) {
      double s = (double)vals[i];
      if (!degenerate) {
        s = fmin(fmax(s, vmin), vmax);
        s = (s - vmin) * inv_den;
      }
      psum += s;
      const double c = fmin(fmax(s, p.clip_min), p.clip_max);
      den[k] = (float)c;
      vals[i] = den[k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
  if (lane == 0) s_red[warp] = psum;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kPreThreads / 32; ++w) t += s_red[w];
    s_mean = t / (double)n;
  }
  __syncthreads();
  const bool erode = p.erosion_threshold != 0.0 && s_mean > p.erosion_threshold;

  // ---- sigma = 8 Gaussian, 'nearest', as three in-place matrix sweeps (z,y,x) --
  if (p.unsharp_strength != 0.0) {
    auto sweep = [&](auto lp_tag, int a, const float* Ma) {
      constexpr int LP = decltype(lp_tag)::value;        // padded line length of this axis
      const int len = a == 0 ? nz : (a == 1 ? ny : nx);
      const int nlines = n / len;
      for (int l = tid; l < nlines; l += kPreThreads) {
        int base, stride;
        if (a == 2) { base = l * nx; stride = 1; }
        else if (a == 1) { const int z = l / nx, x = l - z * nx; base = z * nyx + x; stride = nx; }
        else { base = l; stride = nyx; }
        float v[LP];
#pragma unroll
        for (int j = 0; j < LP; ++j) v[j] = j < len ? vals[base + j * stride] : 0.f;
#pragma unroll 1
        for (int i = 0; i < len; i += 2) {
          // outputs i and i + 1 in one packed accumulator; same order of accumulation per
          // output as a scalar loop over j
          const float4* row = reinterpret_cast<const float4*>(Ma + i * LP);   // (i / 2) * 2 LP
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int j2 = 0; j2 < LP / 2; ++j2) {
            const float4 m = row[j2];
            acc = __ffma2_rn(make_float2(m.x, m.y), make_float2(v[2 * j2], v[2 * j2]), acc);
            acc = __ffma2_rn(make_float2(m.z, m.w), make_float2(v[2 * j2 + 1], v[2 * j2 + 1]), acc);
          }
          vals[base + i * stride] = acc.x;
          if (i + 1 < len) vals[base + (i + 1) * stride] = acc.y;
        }
      }
      __syncthreads();
    };
    sweep(std::integral_constant<int, NZP>(), 0, M);
    sweep(std::integral_constant<int, NLP>(), 1, M + MZ);
    sweep(std::integral_constant<int, NLP>(), 2, M + MZ + MYX);
    // unsharp mask: den + (den - strength * blurred)
    const float us = (float)p.unsharp_strength;
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const int i = tid + k * kPreThreads;
      if (i < n) {
        const float hp = den[k] - us * vals[i];
        vals[i] = den[k] + hp;
      }
    }
    __syncthreads();
  }

  // ---- erosion (min over the 7-voxel octahedron) and store --------------------
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = tid + k * kPreThreads;
    if (i < n) {
      const int z = i / nyx, r = i - z * nyx, y = r / nx, x = r - y * nx;
      float v = vals[i];
      if (erode) {
        if (z > 0) v = fminf(v, vals[i - nyx]);
        if (z < nz - 1) v = fminf(v, vals[i + nyx]);
        if (y > 0) v = fminf(v, vals[i - nx]);
        if (y < ny - 1) v = fminf(v, vals[i + nx]);
        if (x > 0) v = fminf(v, vals[i - 1]);
        if (x < nx - 1) v = fminf(v, vals[i + 1]);
      }
      out[((int64_t)(z0 + z) * g.Y + (y0 + y)) * g.pitch + (x0 + x)] = v;
    }
  }
}

template <typename T, int NZ, int NL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
preprocess_kernel(const T* __restrict__ in, const __grid_constant__ PreGeom g,
                  const __grid_constant__ mmb_preproc_params p, const float* __restrict__ mats,
                  int mat_pitch, float* __restrict__ out) {
  int b = blockIdx.x;
  const int bxi = b % g.nbx; b /= g.nbx;
  const int byi = b % g.nby; b /= g.nby;
  const int bzi = b;
  const int z0 = bzi * g.bz, y0 = byi * g.by, x0 = bxi * g.bx;
  const int nz = min(g.bz, g.Z - z0), ny = min(g.by, g.Y - y0), nx = min(g.bx, g.X - x0);
  if (nz == NZ && ny == NL && nx == NL)       // CTA-uniform
    preprocess_body<T, NZ, NL, true, THREADS>(in, g, p, mats, mat_pitch, out, z0, y0, x0, nz, ny,
                                              nx);
  else
    preprocess_body<T, NZ, NL, false, THREADS>(in, g, p, mats, mat_pitch, out, z0, y0, x0, nz,
                                               ny, nx);
}

// dense matrices of scipy.ndimage.gaussian_filter1d(sigma=8, truncate=4,
// mode='nearest') acting on a length-n line, n = 1..32, float32, device resident
static const float* blur_matrices(int device, int* pitch_out) {
  static std::mutex mu;
  static std::map<int, float*> cache;
  constexpr int NMAX = 32;
  constexpr int PITCH = NMAX * NMAX;
  *pitch_out = PITCH;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(device);
  if (it != cache.end()) return it->second;
  const double sigma = 8.0;
  const int r = (int)(4.0 * sigma + 0.5);
  std::vector<double> w(2 * r + 1);
  double sum = 0.0;
  for (int t = -r; t <= r; ++t) { w[t + r] = exp(-0.5 / (sigma * sigma) * (double)(t * t)); sum += w[t + r]; }
  for (auto& x : w) x /= sum;
  std::vector<float> host((size_t)NMAX * PITCH, 0.f);
  for (int n = 1; n <= NMAX; ++n) {
    std::vector<double> m((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int t = -r; t <= r; ++t) {
        int j = i + t;
        j = j < 0 ? 0 : (j >= n ? n - 1 : j);
        m[(size_t)i * n + j] += w[t + r];
      }
    for (size_t k = 0; k < m.size(); ++k) host[(size_t)(n - 1) * PITCH + k] = (float)m[k];
  }
  float* d = nullptr;
  if (cudaMalloc((void**)&d, host.size() * sizeof(float)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) !=
      cudaSuccess) return nullptr;
  cache[device] = d;
  return d;
}

template <typename T, int NZ, int NL, int THREADS, int MINB>
static int launch_pre(const void* in, const PreGeom& g, const mmb_preproc_params& p,
                      const float* mats, int mat_pitch, float* out, cudaStream_t st) {
  constexpr int NLP = (NL + 3) / 4 * 4, NZP = (NZ + 3) / 4 * 4;
  const size_t smem = (size_t)((NZ * NL * NL + 3) / 4 * 4) * 4 +
                      (size_t)(2 * ((NZ + 1) / 2) * NZP + 2 * 2 * ((NL + 1) / 2) * NLP) * 4 +
                      512 * 4;
  static bool configured = false;
  auto kern = preprocess_kernel<T, NZ, NL, THREADS, MINB>;
  if (!configured) {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    configured = true;
  }
  const int64_t nblocks = (int64_t)g.nbz * g.nby * g.nbx;
  ProfScope ps(PROF_PREPROCESS, (double)g.Z * g.Y * g.X, st);
  kern<<<(unsigned)nblocks, THREADS, smem, st>>>((const T*)in, g, p, mats, mat_pitch, out);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

template <typename T>
static int dispatch_nl(const void* in, const PreGeom& g, const mmb_preproc_params& p,
                       const float* mats, int mat_pitch, float* out, cudaStream_t st) {
  const int m = g.bz > g.by ? (g.bz > g.bx ? g.bz : g.bx) : (g.by > g.bx ? g.by : g.bx);
  const int myx = g.by > g.bx ? g.by : g.bx;
  if (m <= 8) return launch_pre<T, 8, 8, kPreCubicThreads, 2>(in, g, p, mats, mat_pitch, out, st);
  // thin blocks of anisotropic volumes (z resolution coarser than y, x)
  if (g.bz <= 5 && myx > 8 && myx <= 25)
    return launch_pre<T, 5, 25, 160, 6>(in, g, p, mats, mat_pitch, out, st);
  if (g.bz <= 8 && myx > 8 && myx <= 25)
    return launch_pre<T, 8, 25, 256, 4>(in, g, p, mats, mat_pitch, out, st);
  if (m <= 16) return launch_pre<T, 16, 16, kPreCubicThreads, 2>(in, g, p, mats, mat_pitch, out, st);
  if (m <= 25) return launch_pre<T, 25, 25, kPreCubicThreads, 2>(in, g, p, mats, mat_pitch, out, st);
  return launch_pre<T, 32, 32, kPreCubicThreads, 1>(in, g, p, mats, mat_pitch, out, st);
}

int preprocess_large_impl(const void* in, int dtype, const int64_t strides[3], int Z, int Y,
                          int X, int bz, int by, int bx, const mmb_preproc_params* p, float* out,
                          int64_t pitch, void* scratch, cudaStream_t st);

// `scratch` / `scratch_bytes`: memory the caller lends to the large-block path (the fused
// chunk driver's sweep buffers are idle while it preprocesses); may be NULL / 0.
int preprocess_impl(const void* in, int dtype, const int64_t st[3], int Z, int Y, int X, int bz,
                    int by, int bx, const mmb_preproc_params* p, float* out, int64_t pitch,
                    cudaStream_t s, void* scratch, int64_t scratch_bytes) {
  bz = bz < Z ? bz : Z; by = by < Y ? by : Y; bx = bx < X ? bx : X;
  // float64 input: the shared-memory kernel holds the block as float32, which would take the
  // percentiles of ROUNDED samples; the global-memory path selects on the float64 keys
  if (bz > 32 || by > 32 || bx > 32 || dtype == MMB_F64) {
    // blocks that do not fit one CTA's shared memory: global-memory path
    if (scratch_bytes < preprocess_large_work_bytes(Z, Y, pitch, bz, by, bx)) scratch = nullptr;
    return preprocess_large_impl(in, dtype, st, Z, Y, X, bz, by, bx, p, out, pitch, scratch, s);
  }
  PreGeom g;
  g.Z = Z; g.Y = Y; g.X = X; g.bz = bz; g.by = by; g.bx = bx;
  g.nbz = (int)cdiv(Z, bz); g.nby = (int)cdiv(Y, by); g.nbx = (int)cdiv(X, bx);
  g.sz = st[0]; g.sy = st[1]; g.sx = st[2]; g.pitch = pitch;
  int dev = 0;
  MMB_CHECK_CUDA(cudaGetDevice(&dev));
  int mat_pitch = 0;
  const float* mats = blur_matrices(dev, &mat_pitch);
  if (!mats) { set_error("could not build blur matrices"); return MMB_ERR_CUDA; }
  switch (dtype) {
    case MMB_U8:  return dispatch_nl<uint8_t>(in, g, *p, mats, mat_pitch, out, s);
    case MMB_U16: return dispatch_nl<uint16_t>(in, g, *p, mats, mat_pitch, out, s);
    case MMB_F32: return dispatch_nl<float>(in, g, *p, mats, mat_pitch, out, s);
    case MMB_F64: return dispatch_nl<double>(in, g, *p, mats, mat_pitch, out, s);
  }
  set_error("unknown dtype %d", dtype);
  return MMB_ERR_INVALID;
}

}  // namespace mmb

extern "C" int mmb_preprocess_blocks(const void* in, int dtype, const int64_t in_strides[3], int Z,
                                     int Y, int X, int bz, int by, int bx,
                                     const mmb_preproc_params* p, float* out, int64_t pitch,
                                     void* stream) {
  MMB_REQUIRE(in && out && in_strides && p, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(bz > 0 && by > 0 && bx > 0, "bad block shape");
  return mmb::preprocess_impl(in, dtype, in_strides, Z, Y, X, bz, by, bx, p, out, pitch,
                              (cudaStream_t)stream, nullptr, 0);
}
